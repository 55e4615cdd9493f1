// pb_shim.h — TEST INFRASTRUCTURE.  The few lines of protobuf wire format that the stand-ins for the
// protoc-generated headers (parsimony.pb.h, sam.pb.h) need so that the reference's own loaders
// (src/mutation_annotated_tree.cpp:522-612, src/WEPP/sam2pb.cpp:111-151,489-549) compile and run without
// libprotobuf.  Written independently of the product's csrc/pbwire.h; both are checked against files
// produced by the real protobuf runtime (tests/test_formats.py).
#pragma once
#include <cstdint>
#include <istream>
#include <iterator>
#include <ostream>
#include <string>
#include <vector>

namespace pbshim {
struct In {
    const unsigned char* p;
    const unsigned char* e;
    bool get_varint(uint64_t& v) {
        v = 0;
        int s = 0;
        while (p < e) {
            unsigned char b = *p++;
            if (s < 64) v |= (uint64_t)(b & 127) << s;
            s += 7;
            if (b < 128) return true;
        }
        return false;
    }
    bool get_bytes(std::string& out) {
        uint64_t n;
        if (!get_varint(n) || n > (uint64_t)(e - p)) return false;
        out.assign((const char*)p, (size_t)n);
        p += n;
        return true;
    }
    bool skip(int wire) {
        uint64_t v;
        std::string s;
        if (wire == 0) return get_varint(v);
        if (wire == 2) return get_bytes(s);
        if (wire == 1) { p += 8; return p <= e; }
        if (wire == 5) { p += 4; return p <= e; }
        return false;
    }
};
inline void put_varint(std::string& o, uint64_t v) {
    while (v > 127) { o.push_back((char)(v | 128)); v >>= 7; }
    o.push_back((char)v);
}
inline void put_key(std::string& o, int field, int wire) { put_varint(o, (uint64_t)((field << 3) | wire)); }
inline void put_int(std::string& o, int field, int32_t v) { if (v) { put_key(o, field, 0); put_varint(o, (uint64_t)(int64_t)v); } }
inline void put_str(std::string& o, int field, const std::string& s, bool always) {
    if (s.empty() && !always) return;
    put_key(o, field, 2);
    put_varint(o, s.size());
    o += s;
}
}  // namespace pbshim

namespace google { namespace protobuf { namespace io {
class ZeroCopyInputStream {};
class IstreamInputStream : public ZeroCopyInputStream {
public:
    std::istream* in;
    explicit IstreamInputStream(std::istream* s) : in(s) {}
};
class CodedInputStream {
public:
    IstreamInputStream* src;
    explicit CodedInputStream(IstreamInputStream* s) : src(s) {}
    void SetTotalBytesLimit(int, int) {}
    std::string slurp() { return std::string(std::istreambuf_iterator<char>(*src->in), std::istreambuf_iterator<char>()); }
};
}}}  // namespace google::protobuf::io

// base of every stand-in message
struct PbShimMessage {
    virtual ~PbShimMessage() {}
    virtual bool parse(const std::string& bytes) = 0;
    virtual std::string bytes() const = 0;
    bool ParseFromCodedStream(google::protobuf::io::CodedInputStream* in) { return parse(in->slurp()); }
    bool SerializeToOstream(std::ostream* out) const { std::string b = bytes(); out->write(b.data(), (std::streamsize)b.size()); return true; }
};
