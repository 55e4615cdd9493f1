// wepp_oracle.cpp — CPU restatement of WEPP's parsimonious read placement.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (wepp_b200/, include/) may
// import, link or execute this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, and only as the checker.
//
// Every function cites the reference lines (relative to /root/reference) it restates.
// The arithmetic is deliberately kept in the reference's own shape (sorted mismatch
// position vectors carried down a DFS, binary searches into the read's mutation
// list) so that it can be diffed against the reference by eye.  It is NOT the
// algorithm the CUDA path uses (that one is a signed-delta Euler-tour scan).
//
// Parity pin: this restatement is checked against the reference's own object code
// (oracle/_ref, shim-compiled from /root/reference/src/WEPP/*.cpp, see
// oracle/Makefile and oracle/ref_driver.cpp) by tests/golden/make_golden.py, whose
// outputs are committed under tests/golden/.  The reference itself ships no golden
// vectors for this path.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -pthread oracle/wepp_oracle.cpp -o oracle/libwepp_oracle.so

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <queue>
#include <thread>
#include <vector>

namespace {

constexpr uint8_t NUC_N = 0b1111;      // initial_filter.cpp:7
constexpr int NUM_RANGE_BINS = 50;     // config.hpp:13

struct Mut {  // MAT::Mutation, mutation_annotated_tree.hpp:44-53 (ordering = position only)
    int position;
    uint8_t ref_nuc;
    uint8_t mut_nuc;
};

struct Arena {  // the slice of `arena` / `haplotype` the hot path reads (haplotype.hpp:10-43)
    int n = 0;
    const int32_t* parent = nullptr;
    std::vector<std::vector<int>> children;   // haplotype::children, preorder order (arena.cpp:49-53)
    std::vector<std::vector<Mut>> muts;       // haplotype::muts, sorted by position (arena.cpp:46)
};

struct Read {  // raw_read, read.hpp:6-12
    int start, end, degree;
    std::vector<Mut> mutations;  // sorted by position by construction (sam2pb.cpp:519-534)
};

Arena make_arena(int n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                 const uint8_t* mut_ref, const uint8_t* mut_nuc) {
    Arena a;
    a.n = n_nodes;
    a.parent = parent;
    a.children.resize(n_nodes);
    a.muts.resize(n_nodes);
    for (int v = 0; v < n_nodes; ++v) {
        if (parent[v] >= 0) a.children[parent[v]].push_back(v);
        for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k)
            a.muts[v].push_back(Mut{mut_pos[k], mut_ref[k], mut_nuc[k]});
        std::sort(a.muts[v].begin(), a.muts[v].end(),
                  [](const Mut& x, const Mut& y) { return x.position < y.position; });
    }
    return a;
}

Read make_read(int64_t r, const int32_t* start, const int32_t* end, const int32_t* degree,
               const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc) {
    Read rd;
    rd.start = start[r];
    rd.end = end[r];
    rd.degree = degree[r];
    for (int64_t k = rm_off[r]; k < rm_off[r + 1]; ++k) rd.mutations.push_back(Mut{rm_pos[k], 0, rm_nuc[k]});
    return rd;
}

const Mut* read_lower_bound(const Read& read, int position) {
    auto it = std::lower_bound(read.mutations.begin(), read.mutations.end(), position,
                               [](const Mut& m, int p) { return m.position < p; });
    return it == read.mutations.end() ? nullptr : &*it;
}

// initial_filter.cpp:41-105 — recursive single_read_tree.  `parent_locations` are the
// positions where the parent haplotype differs from the read; `my_locations` the same
// for this node, produced by the reference's 3-way merge.  The range-tree compression
// (arena.cpp:68-169) is omitted: a haplotype without a mutation inside the read's
// range inherits its parent's set unchanged, which is exactly what the merge yields.
std::vector<int> node_locations(const Arena& arena, const std::vector<int>& parent_locations, int curr,
                                const Read& read) {
    std::vector<int> my_locations;
    my_locations.reserve(parent_locations.size());
    const std::vector<Mut>& curr_muts = arena.muts[curr];

    size_t i = 0;
    auto j = std::lower_bound(curr_muts.begin(), curr_muts.end(), read.start,
                              [](const Mut& m, int p) { return m.position < p; });  // :53-55
    while (i < parent_locations.size() || (j != curr_muts.end() && j->position <= read.end)) {  // :59
        bool parent_first = i < parent_locations.size() &&
                            (j == curr_muts.end() || j->position > read.end || parent_locations[i] < j->position);
        bool us_first = (j != curr_muts.end() && j->position <= read.end) &&
                        (i == parent_locations.size() || j->position < parent_locations[i]);
        if (us_first || !parent_first) {  // :62-71 and :76-86 share the evaluation
            const Mut* it = read_lower_bound(read, j->position);
            uint8_t read_nuc = (it == nullptr || it->position != j->position) ? j->ref_nuc : it->mut_nuc;  // :64-65
            if (read_nuc != NUC_N && read_nuc != j->mut_nuc) my_locations.push_back(j->position);          // :66
            ++j;
            if (!us_first) ++i;  // equal positions: the parent's entry is re-evaluated (:84-85)
        } else {
            my_locations.push_back(parent_locations[i]);  // :72-75
            ++i;
        }
    }
    return my_locations;
}

void single_read_tree_rec(const Arena& arena, const std::vector<int>& parent_locations, int curr,
                          const Read& read, std::vector<int>& max_nodes, int& max_val) {
    std::vector<int> my_locations = node_locations(arena, parent_locations, curr, read);
    int parsimony = (int)my_locations.size();  // :89-99
    if (parsimony < max_val) {
        max_val = parsimony;
        max_nodes.clear();
        max_nodes.push_back(curr);
    } else if (parsimony == max_val) {
        max_nodes.push_back(curr);
    }
    for (int child : arena.children[curr]) single_read_tree_rec(arena, my_locations, child, read, max_nodes, max_val);  // :101-104
}

// initial_filter.cpp:112-135 — wrapper: seed with the read's non-N mutation positions,
// then keep argmin haplotypes that are not `mapped`.
void single_read_tree(const Arena& arena, const Read& read, const uint8_t* mapped, std::vector<int>& max_indices,
                      int& max_val) {
    std::vector<int> root_mutations;
    for (const Mut& m : read.mutations)
        if (m.mut_nuc != NUC_N) root_mutations.push_back(m.position);  // :118-123
    std::vector<int> max_nodes;
    single_read_tree_rec(arena, root_mutations, 0, read, max_nodes, max_val);
    for (int v : max_nodes)
        if (!mapped || !mapped[v]) max_indices.push_back(v);  // :126-134
}

// ---- range trees (arena.cpp:68-169): an optimisation of the reference that does not change results — per genome
// range (<= NUM_RANGE_TREES of them, from read-start quantiles) a compressed tree of the haplotypes with at least one
// mutation inside the range; every other haplotype joins the `sources` of its nearest kept ancestor.
constexpr int NUM_RANGE_TREES = 25;   // config.hpp:14
struct RangedNode {                   // multi_haplotype, haplotype.hpp:245-260
    int root;
    std::vector<int> sources;
    std::vector<int> children;
};
struct RangedArena {
    std::vector<RangedNode> nodes;
    std::map<std::pair<int, int>, int> root_map;   // ranged_root_map
};

bool has_mutations_in_range(const Arena& arena, int v, int start, int end) {   // haplotype.hpp:57-64
    const std::vector<Mut>& m = arena.muts[v];
    auto it = std::lower_bound(m.begin(), m.end(), start, [](const Mut& x, int p) { return x.position < p; });
    return it != m.end() && it->position <= end;
}

int build_range_tree(const Arena& arena, RangedArena& ra, int parent, int curr, int start, int end) {   // arena.cpp:68-94
    int ret;
    if (parent == -1 || has_mutations_in_range(arena, curr, start, end)) {
        ret = (int)ra.nodes.size();
        ra.nodes.emplace_back();
        ra.nodes[ret].root = curr;
        ra.nodes[ret].sources = {curr};
    } else {
        ret = parent;
        ra.nodes[parent].sources.push_back(curr);
    }
    for (int child : arena.children[curr]) {
        const int sub = build_range_tree(arena, ra, ret, child, start, end);
        if (sub != ret) ra.nodes[ret].children.push_back(sub);
    }
    return ret;
}

void build_range_trees(const Arena& arena, RangedArena& ra, int64_t n_reads, const int32_t* start, const int32_t* end) {   // arena.cpp:96-136
    std::vector<std::pair<int, int>> read_ranges((size_t)n_reads);
    for (int64_t r = 0; r < n_reads; ++r) read_ranges[(size_t)r] = {start[r], end[r]};
    std::sort(read_ranges.begin(), read_ranges.end());
    const int num = (int)std::min<int64_t>(n_reads, NUM_RANGE_TREES);
    for (int i = 0; i < num; ++i) {
        const size_t s = read_ranges.size() * (size_t)i / num, e = read_ranges.size() * (size_t)(i + 1) / num;
        const int read_start = read_ranges[s].first;
        int read_end = 0;
        for (size_t j = s; j < e; ++j) read_end = std::max(read_end, read_ranges[j].second);
        const std::pair<int, int> r{read_start, read_end};
        if (ra.root_map.find(r) == ra.root_map.end()) ra.root_map[r] = build_range_tree(arena, ra, -1, 0, read_start, read_end);
    }
}

int find_range_tree_for(const RangedArena& ra, const Read& read) {   // arena.cpp:154-169
    auto it = ra.root_map.upper_bound({read.start, INT32_MAX});
    do {
        it = std::prev(it);
        if (read.start >= it->first.first && read.end <= it->first.second) return it->second;
    } while (it != ra.root_map.begin());
    return -1;
}

void single_read_tree_ranged_rec(const Arena& arena, const RangedArena& ra, const std::vector<int>& parent_locations, int curr,
                                 const Read& read, std::vector<int>& max_nodes, int& max_val) {   // initial_filter.cpp:41-105
    std::vector<int> my_locations = node_locations(arena, parent_locations, ra.nodes[curr].root, read);
    const int parsimony = (int)my_locations.size();
    if (parsimony < max_val) {
        max_val = parsimony;
        max_nodes.clear();
        max_nodes.push_back(curr);
    } else if (parsimony == max_val) {
        max_nodes.push_back(curr);
    }
    for (int child : ra.nodes[curr].children) single_read_tree_ranged_rec(arena, ra, my_locations, child, read, max_nodes, max_val);
}

void single_read_tree_ranged(const Arena& arena, const RangedArena& ra, const Read& read, const uint8_t* mapped,
                             std::vector<int>& max_indices, int& max_val) {   // initial_filter.cpp:112-135
    std::vector<int> root_mutations;
    for (const Mut& m : read.mutations)
        if (m.mut_nuc != NUC_N) root_mutations.push_back(m.position);
    std::vector<int> max_nodes;
    single_read_tree_ranged_rec(arena, ra, root_mutations, find_range_tree_for(ra, read), read, max_nodes, max_val);
    for (int rn : max_nodes)
        for (int src : ra.nodes[rn].sources)
            if (!mapped || !mapped[src]) max_indices.push_back(src);
}

}  // namespace

extern "C" {

// wepp_filter::cartesian_map, initial_filter.cpp:139-211 (+ node_score, initial_filter.hpp:54-57).
// Outputs: max_parsimony[R], multiplicity[R]; score[N] (double, summed in read order on
// one thread, or per-thread partials merged in thread order when n_threads>1);
// counts[N*50]; EPP lists (sorted arena indices) for reads with multiplicity <= epp_cap
// written to epp_nodes (capacity epp_capacity) with CSR offsets epp_off[R+1]; reads above
// the cap (or overflowing the buffer) get an empty range, as the reference's cache does
// (:189-196).  Returns 0, or -1 if the EPP buffer overflowed.
// range_trees != 0: score through the reference's range trees (arena.cpp:68-169), built — as the reference builds
// them, arena.cpp:96-136 — from the windows of range_n_reads reads (range_start / range_end; the whole sample's read
// set, so that a bounded timing sample meets the same trees the full run would).  Same results, less work per read.
// seconds_map (optional) receives the wall time of the read loop + merge alone: the reference's own "cartesian
// mapping took" boundary (initial_filter.cpp:144,238), without the arena / range-tree construction.
int oracle_cartesian_map_ex(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                            const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int64_t n_reads,
                            const int32_t* start, const int32_t* end, const int32_t* degree, const int64_t* rm_off,
                            const int32_t* rm_pos, const uint8_t* rm_nuc, const uint8_t* mapped, int32_t n_threads,
                            int32_t* max_parsimony, int32_t* multiplicity, double* score, int32_t* counts,
                            int32_t epp_cap, int64_t epp_capacity, int64_t* epp_off, int32_t* epp_nodes,
                            int32_t range_trees, int64_t range_n_reads, const int32_t* range_start, const int32_t* range_end,
                            double* seconds_map);

int oracle_cartesian_map(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                         const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int64_t n_reads,
                         const int32_t* start, const int32_t* end, const int32_t* degree, const int64_t* rm_off,
                         const int32_t* rm_pos, const uint8_t* rm_nuc, const uint8_t* mapped, int32_t n_threads,
                         int32_t* max_parsimony, int32_t* multiplicity, double* score, int32_t* counts,
                         int32_t epp_cap, int64_t epp_capacity, int64_t* epp_off, int32_t* epp_nodes) {
    return oracle_cartesian_map_ex(n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size, n_reads, start, end, degree,
                                   rm_off, rm_pos, rm_nuc, mapped, n_threads, max_parsimony, multiplicity, score, counts, epp_cap,
                                   epp_capacity, epp_off, epp_nodes, 0, 0, nullptr, nullptr, nullptr);
}

int oracle_cartesian_map_ex(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                            const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int64_t n_reads,
                            const int32_t* start, const int32_t* end, const int32_t* degree, const int64_t* rm_off,
                            const int32_t* rm_pos, const uint8_t* rm_nuc, const uint8_t* mapped, int32_t n_threads,
                            int32_t* max_parsimony, int32_t* multiplicity, double* score, int32_t* counts,
                            int32_t epp_cap, int64_t epp_capacity, int64_t* epp_off, int32_t* epp_nodes,
                            int32_t range_trees, int64_t range_n_reads, const int32_t* range_start, const int32_t* range_end,
                            double* seconds_map) {
    Arena arena = make_arena(n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc);
    RangedArena ranged;
    if (range_trees) {
        if (range_n_reads > 0 && range_start && range_end) build_range_trees(arena, ranged, range_n_reads, range_start, range_end);
        else build_range_trees(arena, ranged, n_reads, start, end);
    }
    const auto t_map0 = std::chrono::steady_clock::now();
    int bin_size = genome_size / NUM_RANGE_BINS;  // :146
    if (n_threads < 1) n_threads = 1;

    std::vector<std::vector<double>> t_score(n_threads);
    std::vector<std::vector<int32_t>> t_counts(n_threads);
    std::vector<std::vector<int32_t>> per_read_epp(epp_off ? n_reads : 0);

    auto worker = [&](int t) {
        if (score) t_score[t].assign(n_nodes, 0.0);
        if (counts) t_counts[t].assign((size_t)n_nodes * NUM_RANGE_BINS, 0);
        int64_t lo = n_reads * t / n_threads, hi = n_reads * (t + 1) / n_threads;
        for (int64_t r = lo; r < hi; ++r) {
            Read read = make_read(r, start, end, degree, rm_off, rm_pos, rm_nuc);
            std::vector<int> max_indices;
            int max_val = INT32_MAX;
            if (range_trees) single_read_tree_ranged(arena, ranged, read, mapped, max_indices, max_val);
            else single_read_tree(arena, read, mapped, max_indices, max_val);  // :165
            double delta = (double)read.degree / ((1 + max_val) * (double)max_indices.size());  // hpp:54-57
            int bucket = std::min(read.start / bin_size, NUM_RANGE_BINS - 1);                     // :169
            for (int v : max_indices) {                                                            // :171-177
                if (score) t_score[t][v] += delta;
                if (counts) t_counts[t][(size_t)v * NUM_RANGE_BINS + bucket] += read.degree;
            }
            max_parsimony[r] = max_val;                      // :189
            multiplicity[r] = (int32_t)max_indices.size();   // :190
            if (epp_off && (int)max_indices.size() <= epp_cap) {  // :191-196
                std::sort(max_indices.begin(), max_indices.end());
                per_read_epp[r].assign(max_indices.begin(), max_indices.end());
            }
        }
    };
    if (n_threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
        for (auto& x : th) x.join();
    }
    // :199-211 — merge of the dense per-chunk arrays (here: fixed thread order).
    if (score) {
        std::fill(score, score + n_nodes, 0.0);
        for (int t = 0; t < n_threads; ++t)
            for (int v = 0; v < n_nodes; ++v) score[v] += t_score[t][v];
    }
    if (counts) {
        std::fill(counts, counts + (size_t)n_nodes * NUM_RANGE_BINS, 0);
        for (int t = 0; t < n_threads; ++t)
            for (size_t k = 0; k < (size_t)n_nodes * NUM_RANGE_BINS; ++k) counts[k] += t_counts[t][k];
    }
    if (seconds_map) *seconds_map = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_map0).count();
    int rc = 0;
    if (epp_off) {
        int64_t off = 0;
        for (int64_t r = 0; r < n_reads; ++r) {
            epp_off[r] = off;
            if (off + (int64_t)per_read_epp[r].size() > epp_capacity) {
                rc = -1;
                continue;
            }
            if (!per_read_epp[r].empty())
                std::memcpy(epp_nodes + off, per_read_epp[r].data(), per_read_epp[r].size() * sizeof(int32_t));
            off += (int64_t)per_read_epp[r].size();
        }
        epp_off[n_reads] = off;
    }
    return rc;
}

// Per-(read,node) parsimony of ONE read against every arena node, for brute-force
// cross-checks (tests only): out[v] = |my_locations(v)|, same recursion as above.
int oracle_read_scores(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                       const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t r_start, int32_t r_end,
                       int32_t n_rm, const int32_t* rm_pos, const uint8_t* rm_nuc, int32_t* out) {
    Arena arena = make_arena(n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc);
    Read read;
    read.start = r_start;
    read.end = r_end;
    read.degree = 1;
    for (int k = 0; k < n_rm; ++k) read.mutations.push_back(Mut{rm_pos[k], 0, rm_nuc[k]});
    std::vector<int> seed;
    for (const Mut& m : read.mutations)
        if (m.mut_nuc != NUC_N) seed.push_back(m.position);  // initial_filter.cpp:118-123
    std::vector<std::pair<int, std::vector<int>>> stack;
    stack.emplace_back(0, seed);
    while (!stack.empty()) {
        auto [curr, parent_locations] = std::move(stack.back());
        stack.pop_back();
        std::vector<int> my_locations = node_locations(arena, parent_locations, curr, read);
        out[curr] = (int32_t)my_locations.size();
        for (int child : arena.children[curr]) stack.emplace_back(child, my_locations);
    }
    return 0;
}

// haplotype::stack_muts construction, arena.cpp:18-46: parent's stack minus positions this
// node mutates, plus this node's mutations whose mut_nuc differs from ref_nuc; sorted by
// position.  CSR out: caller passes capacity; returns total entries or -1 if too small.
int64_t oracle_stack_muts(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                          const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t n_sel, const int32_t* sel_nodes,
                          int64_t capacity, int64_t* out_off, int32_t* out_pos, uint8_t* out_nuc) {
    int64_t total = 0;
    for (int s = 0; s < n_sel; ++s) {
        // root → node path
        std::vector<int> path;
        for (int v = sel_nodes[s]; v >= 0; v = parent[v]) path.push_back(v);
        std::reverse(path.begin(), path.end());
        std::vector<Mut> stack;
        for (int v : path) {
            std::vector<Mut> next;
            for (const Mut& m : stack) {  // arena.cpp:20-35
                bool valid = true;
                for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k)
                    if (mut_pos[k] == m.position) {
                        valid = false;
                        break;
                    }
                if (valid) next.push_back(m);
            }
            for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k)  // :38-44
                if (mut_ref[k] != mut_nuc[k]) next.push_back(Mut{mut_pos[k], mut_ref[k], mut_nuc[k]});
            std::sort(next.begin(), next.end(), [](const Mut& x, const Mut& y) { return x.position < y.position; });
            stack.swap(next);
        }
        out_off[s] = total;
        if (total + (int64_t)stack.size() > capacity) return -1;
        for (const Mut& m : stack) {
            out_pos[total] = m.position;
            out_nuc[total] = m.mut_nuc;
            ++total;
        }
    }
    out_off[n_sel] = total;
    return total;
}

// haplotype::mutation_distance(comp, min_pos, max_pos), haplotype.hpp:123-173, on one
// (haplotype stack_muts, read) pair.
static int mutation_distance(const int32_t* s_pos, const uint8_t* s_nuc, int n_stack, const int32_t* c_pos,
                             const uint8_t* c_nuc, int n_comp, int min_pos, int max_pos) {
    int muts = 0;
    int i = (int)(std::lower_bound(s_pos, s_pos + n_stack, min_pos) - s_pos);       // :130-131
    int last_i = (int)(std::upper_bound(s_pos, s_pos + n_stack, max_pos) - s_pos);  // :132-133
    int j = 0;
    while (i < last_i || j < n_comp) {  // :136-170
        if (i == last_i) {
            if (c_nuc[j] != NUC_N) ++muts;
            ++j;
        } else if (s_pos[i] < min_pos) {
            ++i;
        } else if (s_pos[i] > max_pos) {
            return muts;
        } else if (j == n_comp) {
            ++muts;
            ++i;
        } else if (s_pos[i] < c_pos[j]) {
            ++muts;
            ++i;
        } else if (s_pos[i] > c_pos[j]) {
            if (c_nuc[j] != NUC_N) ++muts;
            ++j;
        } else if (s_pos[i] == c_pos[j] && s_nuc[i] != c_nuc[j] && c_nuc[j] != NUC_N) {
            ++muts;
            ++i;
            ++j;
        } else {
            ++i;
            ++j;
        }
    }
    return muts;
}

// EPP-over-candidates idiom, arena.cpp:614-625 / :846-857: for every read, the minimum
// mutation_distance over the candidate haplotypes and all candidates attaining it (in
// candidate order).  dist_out (optional) is the dense R x C distance matrix.
// argmin CSR: am_off[R+1], am_idx (indices INTO the candidate list), capacity R*C.
int oracle_rescore(int32_t n_cand, const int64_t* st_off, const int32_t* st_pos, const uint8_t* st_nuc,
                   int64_t n_reads, const int32_t* start, const int32_t* end, const int64_t* rm_off,
                   const int32_t* rm_pos, const uint8_t* rm_nuc, int32_t* min_dist, int32_t* dist_out,
                   int64_t* am_off, int32_t* am_idx) {
    int64_t off = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int best = INT_MAX;
        std::vector<int> epps;
        for (int c = 0; c < n_cand; ++c) {
            int d = mutation_distance(st_pos + st_off[c], st_nuc + st_off[c], (int)(st_off[c + 1] - st_off[c]),
                                      rm_pos + rm_off[r], rm_nuc + rm_off[r], (int)(rm_off[r + 1] - rm_off[r]),
                                      start[r], end[r]);
            if (dist_out) dist_out[r * n_cand + c] = d;
            if (d <= best) {  // arena.cpp:617-623
                if (d < best) {
                    best = d;
                    epps.clear();
                }
                epps.push_back(c);
            }
        }
        min_dist[r] = best;
        if (am_off) {
            am_off[r] = off;
            for (int c : epps) am_idx[off++] = c;
        }
    }
    if (am_off) am_off[n_reads] = off;
    return 0;
}

}  // extern "C"
