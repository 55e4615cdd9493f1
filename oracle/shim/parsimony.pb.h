// Stand-in for the protoc-generated header of parsimony.proto:4-31 (TEST INFRASTRUCTURE): just the
// accessors src/mutation_annotated_tree.cpp:522-680 uses, over oracle/shim/pb_shim.h.
#pragma once
#include "pb_shim.h"

namespace Parsimony {
class mut : public PbShimMessage {
    int32_t position_ = 0, ref_ = 0, par_ = 0;
    std::vector<int32_t> nuc_;
    std::string chrom_;
public:
    int32_t position() const { return position_; }
    int32_t ref_nuc() const { return ref_; }
    int32_t par_nuc() const { return par_; }
    int mut_nuc_size() const { return (int)nuc_.size(); }
    int32_t mut_nuc(int i) const { return nuc_[i]; }
    const std::string& chromosome() const { return chrom_; }
    void set_position(int32_t v) { position_ = v; }
    void set_ref_nuc(int32_t v) { ref_ = v; }
    void set_par_nuc(int32_t v) { par_ = v; }
    void add_mut_nuc(int32_t v) { nuc_.push_back(v); }
    void clear_mut_nuc() { nuc_.clear(); }
    void set_chromosome(const std::string& s) { chrom_ = s; }
    bool parse(const std::string& b) override {
        pbshim::In in{(const unsigned char*)b.data(), (const unsigned char*)b.data() + b.size()};
        uint64_t key, v;
        while (in.p < in.e) {
            if (!in.get_varint(key)) return false;
            int f = (int)(key >> 3), w = (int)(key & 7);
            if (w == 0 && f == 1) { if (!in.get_varint(v)) return false; position_ = (int32_t)v; }
            else if (w == 0 && f == 2) { if (!in.get_varint(v)) return false; ref_ = (int32_t)v; }
            else if (w == 0 && f == 3) { if (!in.get_varint(v)) return false; par_ = (int32_t)v; }
            else if (w == 0 && f == 4) { if (!in.get_varint(v)) return false; nuc_.push_back((int32_t)v); }
            else if (w == 2 && f == 4) {
                std::string pk;
                if (!in.get_bytes(pk)) return false;
                pbshim::In q{(const unsigned char*)pk.data(), (const unsigned char*)pk.data() + pk.size()};
                while (q.p < q.e) { if (!q.get_varint(v)) return false; nuc_.push_back((int32_t)v); }
            } else if (w == 2 && f == 5) { if (!in.get_bytes(chrom_)) return false; }
            else if (!in.skip(w)) return false;
        }
        return true;
    }
    std::string bytes() const override {
        std::string o, pk;
        pbshim::put_int(o, 1, position_); pbshim::put_int(o, 2, ref_); pbshim::put_int(o, 3, par_);
        for (int32_t x : nuc_) pbshim::put_varint(pk, (uint64_t)(int64_t)x);
        pbshim::put_str(o, 4, pk, false);
        pbshim::put_str(o, 5, chrom_, false);
        return o;
    }
};

template <class T>
inline bool parse_repeated(pbshim::In& in, std::vector<T>& v) {
    std::string b;
    if (!in.get_bytes(b)) return false;
    v.emplace_back();
    return v.back().parse(b);
}

class mutation_list : public PbShimMessage {
    std::vector<mut> m_;
public:
    int mutation_size() const { return (int)m_.size(); }
    const mut& mutation(int i) const { return m_[i]; }
    mut* add_mutation() { m_.emplace_back(); return &m_.back(); }
    bool parse(const std::string& b) override {
        pbshim::In in{(const unsigned char*)b.data(), (const unsigned char*)b.data() + b.size()};
        uint64_t key;
        while (in.p < in.e) {
            if (!in.get_varint(key)) return false;
            if ((key & 7) == 2 && (key >> 3) == 1) { if (!parse_repeated(in, m_)) return false; }
            else if (!in.skip((int)(key & 7))) return false;
        }
        return true;
    }
    std::string bytes() const override {
        std::string o;
        for (const mut& m : m_) pbshim::put_str(o, 1, m.bytes(), true);
        return o;
    }
};

class strings_message : public PbShimMessage {   // a message of one optional string (field 1) + repeated strings (field 2 or 1)
protected:
    std::string name_;
    std::vector<std::string> list_;
    int name_field_, list_field_;
    strings_message(int nf, int lf) : name_field_(nf), list_field_(lf) {}
public:
    bool parse(const std::string& b) override {
        pbshim::In in{(const unsigned char*)b.data(), (const unsigned char*)b.data() + b.size()};
        uint64_t key;
        while (in.p < in.e) {
            if (!in.get_varint(key)) return false;
            int f = (int)(key >> 3), w = (int)(key & 7);
            if (w == 2 && f == list_field_) { list_.emplace_back(); if (!in.get_bytes(list_.back())) return false; }
            else if (w == 2 && f == name_field_) { if (!in.get_bytes(name_)) return false; }
            else if (!in.skip(w)) return false;
        }
        return true;
    }
    std::string bytes() const override {
        std::string o;
        if (name_field_ > 0) pbshim::put_str(o, name_field_, name_, false);
        for (const std::string& s : list_) pbshim::put_str(o, list_field_, s, true);
        return o;
    }
};

class condensed_node : public strings_message {
public:
    condensed_node() : strings_message(1, 2) {}
    const std::string& node_name() const { return name_; }
    int condensed_leaves_size() const { return (int)list_.size(); }
    const std::string& condensed_leaves(int i) const { return list_[i]; }
    void set_node_name(const std::string& s) { name_ = s; }
    void add_condensed_leaves(const std::string& s) { list_.push_back(s); }
};

class node_metadata : public strings_message {
public:
    node_metadata() : strings_message(-1, 1) {}
    int clade_annotations_size() const { return (int)list_.size(); }
    const std::string& clade_annotations(int i) const { return list_[i]; }
    void add_clade_annotations(const std::string& s) { list_.push_back(s); }
};

class data : public PbShimMessage {
    std::string newick_;
    std::vector<mutation_list> muts_;
    std::vector<condensed_node> cond_;
    std::vector<node_metadata> meta_;
public:
    const std::string& newick() const { return newick_; }
    void set_newick(const std::string& s) { newick_ = s; }
    int node_mutations_size() const { return (int)muts_.size(); }
    const mutation_list& node_mutations(int i) const { return muts_[i]; }
    mutation_list* add_node_mutations() { muts_.emplace_back(); return &muts_.back(); }
    int condensed_nodes_size() const { return (int)cond_.size(); }
    const condensed_node& condensed_nodes(int i) const { return cond_[i]; }
    condensed_node* add_condensed_nodes() { cond_.emplace_back(); return &cond_.back(); }
    int metadata_size() const { return (int)meta_.size(); }
    const node_metadata& metadata(int i) const { return meta_[i]; }
    node_metadata* add_metadata() { meta_.emplace_back(); return &meta_.back(); }
    bool parse(const std::string& b) override {
        pbshim::In in{(const unsigned char*)b.data(), (const unsigned char*)b.data() + b.size()};
        uint64_t key;
        while (in.p < in.e) {
            if (!in.get_varint(key)) return false;
            int f = (int)(key >> 3), w = (int)(key & 7);
            if (w == 2 && f == 1) { if (!in.get_bytes(newick_)) return false; }
            else if (w == 2 && f == 2) { if (!parse_repeated(in, muts_)) return false; }
            else if (w == 2 && f == 3) { if (!parse_repeated(in, cond_)) return false; }
            else if (w == 2 && f == 4) { if (!parse_repeated(in, meta_)) return false; }
            else if (!in.skip(w)) return false;
        }
        return true;
    }
    std::string bytes() const override {
        std::string o;
        pbshim::put_str(o, 1, newick_, false);
        for (const auto& m : muts_) pbshim::put_str(o, 2, m.bytes(), true);
        for (const auto& m : cond_) pbshim::put_str(o, 3, m.bytes(), true);
        for (const auto& m : meta_) pbshim::put_str(o, 4, m.bytes(), true);
        return o;
    }
};
}  // namespace Parsimony
