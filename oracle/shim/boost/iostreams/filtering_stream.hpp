// Stand-in for boost::iostreams as the reference's loaders use it (TEST INFRASTRUCTURE):
// filtering_istream = push(gzip_decompressor()) optionally, then push(std::ifstream&); the whole file is read,
// inflated with zlib when asked, and served from memory.  filtering_streambuf<output> writes through
// (compression is not implemented: the reference's writers are never run with a .gz name here).
#pragma once
#include <zlib.h>

#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>

namespace boost { namespace iostreams {
struct gzip_error : std::runtime_error { gzip_error() : std::runtime_error("gzip error") {} };
struct gzip_decompressor {};
struct gzip_compressor {};
struct output {};
struct input {};

class filtering_istream : public std::istream {
    std::stringbuf buf_;
    bool gz_ = false;
public:
    filtering_istream() : std::istream(nullptr) { rdbuf(&buf_); }
    void push(const gzip_decompressor&) { gz_ = true; }
    void push(std::istream& in) {
        std::string raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        if (gz_) {
            z_stream zs{};
            if (inflateInit2(&zs, 15 + 16) != Z_OK) throw gzip_error();
            std::string out;
            zs.next_in = (Bytef*)raw.data();
            zs.avail_in = (uInt)raw.size();
            char chunk[1 << 16];
            int rc = Z_OK;
            while (rc == Z_OK) {
                zs.next_out = (Bytef*)chunk;
                zs.avail_out = sizeof(chunk);
                rc = inflate(&zs, Z_NO_FLUSH);
                if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); throw gzip_error(); }
                out.append(chunk, sizeof(chunk) - zs.avail_out);
            }
            inflateEnd(&zs);
            raw.swap(out);
        }
        buf_.str(raw);
    }
};

template <class Mode>
class filtering_streambuf : public std::stringbuf {
    std::ostream* sink_ = nullptr;
public:
    void push(const gzip_compressor&) { throw gzip_error(); }
    void push(std::ostream& o) { sink_ = &o; }
    void flush_to_sink() { if (sink_) { std::string s = str(); sink_->write(s.data(), (std::streamsize)s.size()); str(""); } }
};
template <class Mode>
inline void close(filtering_streambuf<Mode>& b) { b.flush_to_sink(); }
}}  // namespace boost::iostreams
