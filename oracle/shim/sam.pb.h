// Stand-in for the protoc-generated header of sam.proto:4-18 (TEST INFRASTRUCTURE): the accessors
// src/WEPP/sam2pb.cpp:111-151 and :489-549 use, over oracle/shim/pb_shim.h.
#pragma once
#include "parsimony.pb.h"

namespace Sam {
class read_info : public PbShimMessage {
    std::string read_, content_;
    int32_t start_ = 0, degree_ = 0;
public:
    const std::string& read() const { return read_; }
    const std::string& content() const { return content_; }
    int32_t start_idx() const { return start_; }
    int32_t degree() const { return degree_; }
    void set_read(const std::string& s) { read_ = s; }
    void set_content(const std::string& s) { content_ = s; }
    void set_start_idx(int32_t v) { start_ = v; }
    void set_degree(int32_t v) { degree_ = v; }
    bool parse(const std::string& b) override {
        pbshim::In in{(const unsigned char*)b.data(), (const unsigned char*)b.data() + b.size()};
        uint64_t key, v;
        while (in.p < in.e) {
            if (!in.get_varint(key)) return false;
            int f = (int)(key >> 3), w = (int)(key & 7);
            if (w == 2 && f == 1) { if (!in.get_bytes(read_)) return false; }
            else if (w == 0 && f == 3) { if (!in.get_varint(v)) return false; start_ = (int32_t)v; }
            else if (w == 0 && f == 5) { if (!in.get_varint(v)) return false; degree_ = (int32_t)v; }
            else if (w == 2 && f == 6) { if (!in.get_bytes(content_)) return false; }
            else if (!in.skip(w)) return false;
        }
        return true;
    }
    std::string bytes() const override {
        std::string o;
        pbshim::put_str(o, 1, read_, false);
        pbshim::put_int(o, 3, start_);
        pbshim::put_int(o, 5, degree_);
        pbshim::put_str(o, 6, content_, false);
        return o;
    }
};

class column_info : public Parsimony::strings_message {
public:
    column_info() : strings_message(1, 2) {}
    const std::string& column_name() const { return name_; }
    void set_column_name(const std::string& s) { name_ = s; }
    int input_columns_size() const { return (int)list_.size(); }
    const std::vector<std::string>& input_columns() const { return list_; }
    std::string* add_input_columns() { list_.emplace_back(); return &list_.back(); }
};

class sam : public PbShimMessage {
    std::vector<read_info> reads_;
    std::vector<column_info> cols_;
public:
    int reads_size() const { return (int)reads_.size(); }
    const std::vector<read_info>& reads() const { return reads_; }
    read_info* add_reads() { reads_.emplace_back(); return &reads_.back(); }
    int reverse_columns_size() const { return (int)cols_.size(); }
    const std::vector<column_info>& reverse_columns() const { return cols_; }
    column_info* add_reverse_columns() { cols_.emplace_back(); return &cols_.back(); }
    bool parse(const std::string& b) override {
        pbshim::In in{(const unsigned char*)b.data(), (const unsigned char*)b.data() + b.size()};
        uint64_t key;
        while (in.p < in.e) {
            if (!in.get_varint(key)) return false;
            int f = (int)(key >> 3), w = (int)(key & 7);
            if (w == 2 && f == 1) { if (!Parsimony::parse_repeated(in, reads_)) return false; }
            else if (w == 2 && f == 2) { if (!Parsimony::parse_repeated(in, cols_)) return false; }
            else if (!in.skip(w)) return false;
        }
        return true;
    }
    std::string bytes() const override {
        std::string o;
        for (const auto& m : reads_) pbshim::put_str(o, 1, m.bytes(), true);
        for (const auto& m : cols_) pbshim::put_str(o, 2, m.bytes(), true);
        return o;
    }
};
}  // namespace Sam
