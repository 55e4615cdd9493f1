// pbwire.h — the subset of the protobuf wire format the WEPP file formats use (proto3 scalars int32 /
// string, nested messages, repeated fields packed or not): parsimony.proto:4-31 (the MAT) and
// sam.proto:4-18 (collapsed reads).  protobuf / protoc are not a dependency of this library: both schemas
// are eight tiny messages, read and written here directly.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>

namespace wepp::pb {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    Reader(const void* data, size_t n) : p((const uint8_t*)data), end((const uint8_t*)data + n) {}
    bool done() const { return p >= end || !ok; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 70; shift += 7) {
            if (p >= end) { ok = false; return 0; }
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7F) << (shift < 64 ? shift : 63);
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    // next field: returns false at the end of the message
    bool next(uint32_t& field, uint32_t& wire) {
        if (done()) return false;
        const uint64_t key = varint();
        if (!ok) return false;
        field = (uint32_t)(key >> 3);
        wire = (uint32_t)(key & 7);
        return true;
    }
    std::string_view bytes() {   // wire type 2 payload
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return {}; }
        std::string_view s((const char*)p, (size_t)n);
        p += n;
        return s;
    }
    void skip(uint32_t wire) {
        switch (wire) {
            case 0: varint(); break;
            case 1: if (end - p < 8) ok = false; else p += 8; break;
            case 2: bytes(); break;
            case 5: if (end - p < 4) ok = false; else p += 4; break;
            default: ok = false;
        }
    }
};

struct Writer {
    std::string out;
    void varint(uint64_t v) {
        while (v >= 0x80) { out.push_back((char)(v | 0x80)); v >>= 7; }
        out.push_back((char)v);
    }
    void key(uint32_t field, uint32_t wire) { varint(((uint64_t)field << 3) | wire); }
    // proto3: default values (0, "") are not serialised
    void int32(uint32_t field, int32_t v) {
        if (v == 0) return;
        key(field, 0);
        varint((uint64_t)(int64_t)v);   // negative int32 is sign-extended to 10 bytes
    }
    void str(uint32_t field, std::string_view s, bool always = false) {
        if (s.empty() && !always) return;
        key(field, 2);
        varint(s.size());
        out.append(s.data(), s.size());
    }
    void message(uint32_t field, const std::string& body) {
        key(field, 2);
        varint(body.size());
        out.append(body);
    }
};

}  // namespace wepp::pb
