// host_arena.h — MAT-level tree + reads -> the flattened arena the placement kernels consume.
//
// Replaces the reference's arena constructor (src/WEPP/arena.hpp:56-79): read masking
// (:62-72), the covered-site set (site_read_map, :157-175), create_condensed_tree
// (src/WEPP/util.cpp:79-133, BFS child order) and the preorder flatten arena::from_mat
// (src/WEPP/arena.cpp:3-56) including leaf_count (util.cpp:298-315).  Same results, different
// algorithms: coverage is a position difference array instead of per-read std::find scans, and
// nothing is O(N * depth) — stack_muts are not materialised (the device path never needs them;
// wepp_rescore rebuilds them for candidates only).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace wepp {

struct ArenaHost {
    int32_t genome_size = 0;
    // arena nodes in preorder
    std::vector<int32_t> parent, source, leaf_count;
    std::vector<uint8_t> is_leaf;
    std::vector<int64_t> mut_off;
    std::vector<int32_t> mut_pos;
    std::vector<uint8_t> mut_ref, mut_nuc;
    // condensed_node_mappings: MAT nodes folded into each arena node (first = its source)
    std::vector<int64_t> map_off;
    std::vector<int32_t> map_nodes;
    // reads after masking
    std::vector<int64_t> rm_off;
    std::vector<int32_t> rm_pos;
    std::vector<uint8_t> rm_nuc;
    std::vector<uint8_t> covered;  // [genome_size + 1]
};

std::string build_arena(int32_t n_mat_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                        const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int32_t n_masked,
                        const int32_t* masked, int64_t n_reads, const int32_t* start, const int32_t* end,
                        const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc, ArenaHost& out);

}  // namespace wepp
