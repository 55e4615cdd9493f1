// host_io.h — the on-disk formats either side of the placement path:
//   * the UShER mutation-annotated tree, protobuf `Parsimony::data` optionally gzipped
//     (parsimony.proto:4-31; reference loader src/mutation_annotated_tree.cpp:415-508 Newick,
//     :522-612 load, :720-746 Node::add_mutation, :1224-1272 uncondense_leaves);
//   * the collapsed reads, protobuf `Sam::sam` (sam.proto:4-18; reference reader
//     src/WEPP/sam2pb.cpp:489-549, writer :111-151);
//   * the reference FASTA and mask.bed (src/WEPP/dataset.hpp:90-113,152-203).
// Everything is flat arrays (string pools + offsets) so that it crosses the C ABI unchanged.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

namespace wepp {

struct StringPool {
    std::vector<int64_t> off{0};
    std::string chars;
    size_t size() const { return off.size() - 1; }
    void push(std::string_view s) {
        chars.append(s.data(), s.size());
        off.push_back((int64_t)chars.size());
    }
    std::string_view at(size_t i) const { return std::string_view(chars.data() + off[i], (size_t)(off[i + 1] - off[i])); }
    std::string str(size_t i) const { return std::string(at(i)); }
};

// MAT::Tree as flat arrays.  Nodes are in creation order: Newick preorder first (the order
// Parsimony::data::node_mutations is given in), then the leaves added by uncondense_leaves.
// parent[v] < v; a node's children are its nodes with that parent in index order.
struct MatTree {
    std::vector<int32_t> parent;
    std::vector<float> branch_length;
    std::vector<std::string> id;               // Node::identifier
    std::vector<int64_t> mut_off;              // CSR over nodes, mutations sorted by position
    std::vector<int32_t> mut_pos;              // negative = masked mutation (kept, nucleotides 0)
    std::vector<uint8_t> mut_ref, mut_par, mut_nuc;
    int32_t n_annotations = 0;                 // Tree::get_num_annotations()
    std::vector<std::vector<std::string>> clade;   // per node, n_annotations strings (may be shorter for leaves added later)
    int64_t n_internal_ids = 0;                // Tree::curr_internal_node
    // condensed nodes as loaded (cleared by uncondense)
    std::vector<std::string> condensed_name;
    std::vector<std::vector<std::string>> condensed_leaves;
    int32_t n_nodes() const { return (int32_t)parent.size(); }
};

// Returns "" on success, else an error message (the reference prints it and exit(1)s).
std::string read_file_maybe_gz(const std::string& path, std::string& out);
std::string parse_mat(const std::string& pb_bytes, MatTree& out);
std::string load_mat(const std::string& path, bool uncondense, MatTree& out);
void uncondense_leaves(MatTree& t);
// Parsimony::data bytes of a tree (used by tests and by the synthetic-data writer)
std::string serialize_mat(const MatTree& t);

// std::vector<raw_read> (src/WEPP/read.hpp:6-12) + dataset::read_reverse_merge as flat arrays.
struct ReadSet {
    StringPool name;                           // raw_read::read
    std::vector<int32_t> start, end, degree;
    std::vector<int64_t> rm_off{0};
    std::vector<int32_t> rm_pos;
    std::vector<uint8_t> rm_nuc;               // 4-bit ids: 1,2,4,8 or 15 (N) — what get_nuc_id gives for the content char
    // reverse merge: column name -> raw read names (sam.proto:11-14)
    StringPool rev_key;
    std::vector<int64_t> rev_off{0};           // CSR over keys into rev_val
    StringPool rev_val;
    int64_t n_reads() const { return (int64_t)start.size(); }
};
std::string parse_reads(const std::string& pb_bytes, const std::string& reference, ReadSet& out, int n_threads);
std::string load_reads(const std::string& path, const std::string& reference, ReadSet& out, int n_threads);

// FASTA: first-token name of the header and the upper-cased concatenated sequence
std::string load_fasta(const std::string& path, std::string& name, std::string& seq);
// mask.bed: third whitespace-separated column of every line that has three (a missing file = no mask)
std::vector<int32_t> load_mask_bed(const std::string& path);

// nucleotide codec, src/mutation_annotated_tree.cpp:19-74 and :88-139
uint8_t nuc_id(char c);
char nuc_char(uint8_t id);

}  // namespace wepp
