// host_arena.cpp — see host_arena.h.
#include "host_arena.h"

#include <algorithm>

namespace wepp {

std::string build_arena(int32_t n_mat, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                        const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int32_t n_masked,
                        const int32_t* masked, int64_t n_reads, const int32_t* start, const int32_t* end,
                        const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc, ArenaHost& out) {
    if (n_mat < 1) return "tree has no nodes";
    if (parent[0] != -1) return "parent[0] must be -1 (node 0 is the root)";
    for (int32_t v = 1; v < n_mat; ++v)
        if (parent[v] < 0 || parent[v] >= v) return "parent[v] must satisfy 0 <= parent[v] < v";
    out = ArenaHost();
    out.genome_size = genome_size;
    const int32_t g = genome_size;

    // ---- masked sites, read masking (arena.hpp:62-72) ------------------------------------------
    std::vector<uint8_t> is_masked((size_t)g + 2, 0);
    for (int32_t i = 0; i < n_masked; ++i)
        if (masked[i] >= 1 && masked[i] <= g) is_masked[masked[i]] = 1;
    out.rm_off.assign((size_t)n_reads + 1, 0);
    out.rm_pos.reserve((size_t)rm_off[n_reads]);
    out.rm_nuc.reserve((size_t)rm_off[n_reads]);
    // ---- covered sites (arena.hpp:157-175): position p is covered if some read spans it with a
    //      base that is not N (after the erasure above) and p is not masked -----------------------
    std::vector<int64_t> span((size_t)g + 2, 0), n_at((size_t)g + 2, 0);
    for (int64_t r = 0; r < n_reads; ++r) {
        const int32_t s = start[r], e = end[r];
        if (s < 1 || e > g || e < s - 1) return "read window outside the genome";
        if (e >= s) {
            ++span[s];
            --span[e + 1];
        }
        for (int64_t k = rm_off[r]; k < rm_off[r + 1]; ++k) {
            const int32_t p = rm_pos[k];
            if (p < s || p > e) return "read mutation outside its window";
            if (is_masked[p]) continue;
            out.rm_pos.push_back(p);
            out.rm_nuc.push_back(rm_nuc[k]);
            if (rm_nuc[k] == 15) ++n_at[p];
        }
        out.rm_off[r + 1] = (int64_t)out.rm_pos.size();
    }
    out.covered.assign((size_t)g + 1, 0);
    {
        int64_t c = 0;
        for (int32_t p = 1; p <= g; ++p) {
            c += span[p];
            out.covered[p] = (c - n_at[p] > 0) && !is_masked[p];
        }
    }

    // ---- children lists of the MAT (index order = the loader's child order) ---------------------
    std::vector<int32_t> child_off((size_t)n_mat + 1, 0), child;
    for (int32_t v = 1; v < n_mat; ++v) ++child_off[parent[v] + 1];
    for (int32_t v = 0; v < n_mat; ++v) child_off[v + 1] += child_off[v];
    child.resize((size_t)std::max(n_mat - 1, 0));
    {
        std::vector<int32_t> cur(child_off.begin(), child_off.end() - 1);
        for (int32_t v = 1; v < n_mat; ++v) child[cur[parent[v]]++] = v;
    }
    auto has_covered = [&](int32_t v) {
        for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k) {
            if (mut_pos[k] < 1 || mut_pos[k] > g) continue;
            if (out.covered[mut_pos[k]]) return true;
        }
        return false;
    };

    // ---- condensed tree, BFS (util.cpp:79-133): keep a node iff it mutates a covered site; other
    //      nodes fold into the nearest kept ancestor; children in BFS pop order.  Flat arrays throughout
    //      (a vector per node costs seconds at 8 M nodes): the BFS order of the MAT, the condensed node every
    //      MAT node lands in, then children and folded nodes as CSR filled in pop order -----------------
    std::vector<int32_t> bfs((size_t)n_mat), land((size_t)n_mat, 0);   // land[v]: condensed id of v, or of the kept ancestor it folds into
    std::vector<int32_t> c_src, c_parent;                              // condensed nodes in creation (= pop) order
    c_src.reserve((size_t)n_mat);
    c_parent.reserve((size_t)n_mat);
    c_src.push_back(0);
    c_parent.push_back(-1);
    {
        int32_t head = 0, tail = 0;
        bfs[tail++] = 0;
        while (head < tail) {
            const int32_t v = bfs[head++];
            if (v != 0) {
                const int32_t np = land[parent[v]];
                if (has_covered(v)) {
                    land[v] = (int32_t)c_src.size();
                    c_src.push_back(v);
                    c_parent.push_back(np);
                } else {
                    land[v] = np;
                }
            }
            for (int32_t k = child_off[v]; k < child_off[v + 1]; ++k) bfs[tail++] = child[k];
        }
    }
    const int32_t nc = (int32_t)c_src.size();
    std::vector<int32_t> cc_off((size_t)nc + 1, 0), cc((size_t)std::max(nc - 1, 0));   // condensed children
    for (int32_t c = 1; c < nc; ++c) ++cc_off[c_parent[c] + 1];
    for (int32_t c = 0; c < nc; ++c) cc_off[c + 1] += cc_off[c];
    {
        std::vector<int32_t> cur(cc_off.begin(), cc_off.end() - 1);
        for (int32_t c = 1; c < nc; ++c) cc[cur[c_parent[c]]++] = c;   // creation order = pop order
    }
    std::vector<int64_t> cm_off((size_t)nc + 1, 0);                   // MAT nodes folded into each condensed node
    std::vector<int32_t> cm((size_t)n_mat);
    for (int32_t v = 0; v < n_mat; ++v) ++cm_off[land[v] + 1];
    for (int32_t c = 0; c < nc; ++c) cm_off[c + 1] += cm_off[c];
    {
        std::vector<int64_t> cur(cm_off.begin(), cm_off.end() - 1);
        for (int32_t i = 0; i < n_mat; ++i) cm[cur[land[bfs[i]]]++] = bfs[i];   // itself first (popped before what folds into it), then pop order
    }

    // ---- leaves under every MAT node (get_num_leaves, util.cpp:298-315) -------------------------
    std::vector<int32_t> leaves((size_t)n_mat, 0);
    for (int32_t v = n_mat - 1; v >= 0; --v) {
        if (child_off[v + 1] == child_off[v]) leaves[v] = 1;
        if (v > 0) leaves[parent[v]] += leaves[v];
    }

    // ---- preorder flatten (arena.cpp:3-56) ----------------------------------------------------
    out.parent.reserve(nc);
    out.source.reserve(nc);
    out.leaf_count.reserve(nc);
    out.is_leaf.reserve(nc);
    out.mut_off.reserve((size_t)nc + 1);
    out.map_off.reserve((size_t)nc + 1);
    out.map_nodes.reserve((size_t)n_mat);
    out.mut_off.push_back(0);
    out.map_off.push_back(0);
    std::vector<int32_t> arena_of((size_t)nc, -1);
    std::vector<std::pair<int32_t, int32_t>> stack;  // (condensed node, next child slot)
    std::vector<std::pair<int32_t, int64_t>> ms;
    auto emit = [&](int32_t c) {
        const int32_t idx = (int32_t)out.parent.size();
        arena_of[c] = idx;
        out.parent.push_back(c_parent[c] < 0 ? -1 : arena_of[c_parent[c]]);
        const int32_t v = c_src[c];
        out.source.push_back(v);
        out.leaf_count.push_back(leaves[v]);
        out.is_leaf.push_back(child_off[v + 1] == child_off[v]);
        // muts: the covered mutations, sorted by position (arena.cpp:45-46)
        ms.clear();
        bool sorted = true;
        for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k)
            if (mut_pos[k] >= 1 && mut_pos[k] <= g && out.covered[mut_pos[k]]) {
                sorted = sorted && (ms.empty() || ms.back().first <= mut_pos[k]);
                ms.emplace_back(mut_pos[k], k);
            }
        if (!sorted) std::stable_sort(ms.begin(), ms.end(), [](auto& a, auto& b) { return a.first < b.first; });
        for (auto& m : ms) {
            out.mut_pos.push_back(m.first);
            out.mut_ref.push_back(mut_ref[m.second]);
            out.mut_nuc.push_back(mut_nuc[m.second]);
        }
        out.mut_off.push_back((int64_t)out.mut_pos.size());
        out.map_nodes.insert(out.map_nodes.end(), cm.begin() + cm_off[c], cm.begin() + cm_off[c + 1]);
        out.map_off.push_back((int64_t)out.map_nodes.size());
    };
    emit(0);
    stack.emplace_back(0, cc_off[0]);
    while (!stack.empty()) {
        auto& [c, k] = stack.back();
        if (k < cc_off[c + 1]) {
            const int32_t ch = cc[k++];
            emit(ch);
            stack.emplace_back(ch, cc_off[ch]);
        } else {
            stack.pop_back();
        }
    }
    return "";
}

}  // namespace wepp
