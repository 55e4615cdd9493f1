// delta_place.cuh — placement by SPARSE CORRECTIONS over the distinct window-restricted haplotypes ("states",
// state_place.cuh): the default for wepp_place over the whole read set with nothing mapped and no EPP lists
// (WEPP_DELTA_PLACE=0 keeps state_place_kernel, which also serves the lists / read sets that opt out here).
//
// A state s of a window list is a set of (position, allele) pairs — its net mutations inside the list's range — and
// a read r is a window [a, b] plus a few mutations M_r (allele classes A/C/G/T/N).  SURVEY Appendix A then reads
//
//     parsimony(r, s) = k_r + base_w(s) - red(r, s)
//       k_r       = non-N mutations of the read                       (the seed set, initial_filter.cpp:118-123)
//       base_w(s) = mutated positions of s inside the window [a, b]   (a read as the reference mismatches each)
//       red(r, s) = sum over the read's mutations (p, c) that s also mutates: 1, or 2 when the alleles agree
//                   (an N, or another allele, takes back the state's mismatch; the same allele also takes back
//                   the read's own seed mismatch)
//
// (each table of a state entry is "final allele X vs reference R": delta[ref] = (X != R), delta[c] = -(c == X),
// delta[N] = 0, checked entry by entry when the posting lists are built — any other table and the plan opts out; a
// reversion X == R keeps an entry with delta[ref] = 0 that only a read carrying the reference base as a "mutation"
// would see, so base_w sums delta[ref] and red = delta[ref] - delta[c] may be 0).  base_w is the
// same for every read of a window and red touches only the states that mutate one of the read's few positions:
// ~550-1,500 states per read on the 8 M-node bench tree instead of the 37 k entries of the list's 15 k states.
// Per read: start from the window's histogram "countable nodes per base score", move the touched states down by
// their red, read off the minimum and its node count (initial_filter.cpp:89-99, :126-134).  Per-node weights
// (:167-177): all states at base == min that the read does not touch get the read's weight through a per-(window,
// score) sum G that is expanded once per step; touched states that end at the minimum get it directly.  A touched
// state can never leave the minimum (red > 0 only lowers it), so there is nothing to subtract.
//
// What keeps the walk short (all exact, each explained where it is done): only the posting cells whose states can
// still end at or below the window's minimum are walked (post_pairs_kernel / delta_mrec_kernel); a lane that has seen
// a lower score stops tracking the hits above it, and a read's shortest posting list comes first (dp_read).
//
//   post_pairs_kernel      posting lists: per (list, position, core base score) the states that mutate the position
//                          {state | allele one-hot | delta[ref], nodes}
//   delta_keys_kernel      sort key (bucket, window, cost) per read -> cub sort + run-length encode = window groups
//   window_base_kernel     base_w(s) for every group of a list, and the groups' histograms
//   delta_mrec_kernel      per read mutation: the postings it walks (they depend on the window's minimum)
//   delta_place_kernel     persistent CTAs (8 warps, two per SM) pull work units (a window group, or 512 reads of a
//                          large one); a warp takes a read
//   delta_finalize_kernel  per-(bucket, state) accumulators += sum over the bucket's groups of G[group][base(s)]
// The peak loop places its removed reads through the same kernels over the subset's own work units
// (wepp_abi.cu: place_subset_by_delta).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels.cuh"
#include "rescore_tiles.cuh"
#include "state_place.cuh"

namespace wepp {

// CTA shapes (warps per CTA x CTAs per SM, the shared memory split between the CTAs): 8 x 2 where the widest list of
// the plan fits half of an SM's shared memory (measured 3.60 ms; 16 x 1: 4.12, 6 x 3: 4.2, 4 x 4: 5.3 — a heavy read at
// the end of a unit idles the other warps of its CTA, and fewer warps wait with smaller CTAs), else 16 x 1
constexpr int DP_WARPS_SM = 16;       // warps per SM in either shape
constexpr int DP_BINS = 64;           // score bins: bin = base - red + DP_VOFF
constexpr int DP_VOFF = SW_MAX_ACTIVE;   // base <= SW_MAX_ACTIVE and red <= 2 * base
constexpr int DP_CAND_MIN = 64;       // candidate queue entries per warp: at least this many (the rest of the shared memory is split)
constexpr int DP_FAST_MUTS = 7;       // reads with more mutations use byte scratch in global memory (nibbles hold <= 15)
constexpr int DP_UNIT = 512;          // reads per work unit (a group of up to 2 * DP_UNIT reads stays whole)
__host__ __device__ constexpr int dp_fixed(int warps) { return 16 + DP_BINS * 4 + warps * DP_BINS * 4; }   // ctrl, whist, mv
static_assert(DP_VOFF + SW_MAX_ACTIVE + 1 <= DP_BINS, "score bins");
constexpr uint32_t DP_X_NONE = 7u;    // allele class of a state entry that equals no read allele (IUPAC union)
constexpr int DP_MARGIN = 16;         // the core of a list of `width` positions is [DP_MARGIN, width - 1 - DP_MARGIN]: inside every window of the list
constexpr int DP_LEVELS = 8;          // posting cells per slot: core base score 0..6 and "7 or more"

struct DeltaGroup {     // reads of one bucket with the same window
    int64_t base_off;   // first byte of base_w in the base buffer (S_list bytes, 16-byte aligned)
    int32_t list, bucket;
    int32_t a_rel, b_rel;   // window relative to the list's first position
    int32_t m0;         // smallest occupied bin of the window's histogram
    int32_t prune;      // 1: the window holds the list's core (base_w >= core base score: the reads walk a prefix of every slot)
};
struct DeltaUnit {
    int32_t group, first, count, pad;   // reads order[first .. first + count)
};

// ---- posting lists -----------------------------------------------------------------------------------------------
// A state entry's table is "final allele X vs reference R": delta[ref] = (X != R), delta[c] = -(c == X), delta[N] = 0
// (host_prep.cpp mismatch_after / mismatch_seed).  Returns bit 3 = delta[ref], bits 0..2 = the class of X (1..4, or
// DP_X_NONE when X is an IUPAC union no read allele equals), or 0xFF when the table is of another form.
__device__ __forceinline__ uint32_t dp_table_class(uint32_t z, uint32_t w) {
    const uint32_t bref = z & 0xFFu;
    if (bref > 1u) return 0xFFu;
    const uint32_t b[4] = {(z >> 8) & 0xFFu, (z >> 16) & 0xFFu, z >> 24, w & 0xFFu};
    uint32_t xc = DP_X_NONE;
    int n_match = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (b[c] == 0xFFu) {
            xc = (uint32_t)c + 1u;
            ++n_match;
        } else if (b[c] != 0u) {
            return 0xFFu;
        }
    }
    return n_match <= 1 ? (xc | (bref << 3)) : 0xFFu;
}

// One thread per state: a (sort key, posting) pair per state entry, at the entry's own index, and the entries per
// posting cell.  A slot (list, position) has DP_LEVELS cells, by the state's CORE base score cb = its mismatching
// positions inside the list's core (capped): cb <= base_w(s) for every window w of the list that holds the core, so a
// read whose mutations can take back at most R mismatches only walks the cells cb <= (window minimum) + R of its slots
// — the other states cannot end at or below the window's minimum whatever the read hits (delta_mrec_kernel).  Sorting
// (stably) by cell keeps the postings in state order inside a cell: the states are numbered in Euler order of their
// first node (state_place.cuh), so a mutation carried by a clade posts runs of consecutive states — the 32 states a
// warp touches together then mostly lie next to each other (dp_nibble_word below).  The empty state's placeholder
// entry sorts to the end (key = all ones).
__global__ void post_pairs_kernel(const Entry* __restrict__ state_ent, const int64_t* __restrict__ state_eoff,
                                  const int32_t* __restrict__ state_list, const int32_t* __restrict__ state_first,
                                  const int32_t* __restrict__ lpos_base, const ListDesc* __restrict__ list_desc, int32_t n_states,
                                  uint32_t* __restrict__ cell_count, uint64_t* __restrict__ pkey, uint64_t* __restrict__ pval,
                                  int32_t* __restrict__ bad) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    const int l = state_list[s];
    const uint32_t local = (uint32_t)(s - state_first[l]);
    const int32_t lp = lpos_base[l];
    const int width = list_desc[l].width;
    const int64_t e0 = state_eoff[s], e1 = state_eoff[s + 1];
    uint32_t cb = 0u;
    for (int64_t k = e0; k < e1; ++k) {
        const Entry e = state_ent[k];
        const int q = (int)(e.w >> 16);
        if ((e.z & 0xFFu) == 1u && q >= DP_MARGIN && q <= width - 1 - DP_MARGIN) ++cb;   // delta[ref] = 1: a mismatch of the state
    }
    cb = min(cb, (uint32_t)DP_LEVELS - 1u);
    for (int64_t k = e0; k < e1; ++k) {
        const Entry e = state_ent[k];
        pkey[k] = ~0ull;
        pval[k] = 0ull;
        if (e.z == 0u && (e.w & 0xFFu) == 0u) continue;   // the empty state's placeholder entry
        const uint32_t xc = dp_table_class(e.z, e.w);
        if (xc == 0xFFu || local >= (1u << 24)) {
            *bad = 1;
            continue;
        }
        const uint32_t cell = ((uint32_t)lp + (e.w >> 16)) * DP_LEVELS + cb;
        atomicAdd(cell_count + cell, 1u);
        pkey[k] = (uint64_t)cell;
        // uint2 {state | allele of the state one-hot (A, C, G, T: bits 24..27; none for an IUPAC union) | delta[ref] << 28,
        // nodes}: a read mutation of class c takes back popc(x & (1 << (23 + c) | 1 << 28)) mismatches
        const uint32_t x_class = xc & 7u, one_hot = (x_class >= 1u && x_class <= 4u) ? (1u << (23u + x_class)) : 0u;
        pval[k] = (uint64_t)(local | one_hot | ((xc >> 3) << 28)) | ((uint64_t)e.y << 32);
    }
}

// ---- window groups -----------------------------------------------------------------------------------------------
constexpr int DP_COST_BITS = 8;
// one block per tile: key = (bucket << 24 | (start - b0) << 12 | (end - b0)) << 8 | 255 - min(255, touches / 64),
// value = read index.  Sorted ascending, the reads of a window come heaviest first (touches = states the read's
// mutations reach): the warps of a CTA pull reads from the front, so the last ones to finish are the cheapest.
__global__ void delta_keys_kernel(const TileDesc* __restrict__ tiles, const BucketDesc* __restrict__ buckets,
                                  const ListDesc* __restrict__ list_desc, const int64_t* __restrict__ perm,
                                  const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                                  const int64_t* __restrict__ rm_off, const int32_t* __restrict__ rm_pos,
                                  const uint8_t* __restrict__ rm_code, const uint32_t* __restrict__ post_off,
                                  const int32_t* __restrict__ lpos_base, uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
    const TileDesc td = tiles[blockIdx.x];
    const int32_t list = buckets[td.bucket].list;
    const int32_t b0 = list_desc[list].b0, lp = lpos_base[list];
    for (int i = threadIdx.x; i < td.count; i += blockDim.x) {
        const int64_t rid = perm[td.first + i];
        uint32_t cost = 0;
        // (the cells the read will walk if its window's minimum is the empty state's, the usual case: delta_mrec_kernel)
        int levels = 1;
        for (int64_t k = rm_off[rid]; k < rm_off[rid + 1]; ++k) levels += rm_code[k] <= 4u ? 2 : 1;
        levels = min(levels, DP_LEVELS);
        for (int64_t k = rm_off[rid]; k < rm_off[rid + 1]; ++k) {
            const size_t slot = (size_t)(lp + rm_pos[k] - b0);
            cost += post_off[slot * DP_LEVELS + levels] - post_off[slot * DP_LEVELS] + 16u;
        }
        const uint64_t w = ((uint64_t)td.bucket << 24) | ((uint64_t)(start[rid] - b0) << 12) | (uint64_t)(end[rid] - b0);
        key[td.first + i] = (w << DP_COST_BITS) | (uint64_t)(255u - min(255u, cost >> 6));
        val[td.first + i] = (uint32_t)rid;
    }
}

__global__ void delta_window_of_key_kernel(const uint64_t* __restrict__ key, int64_t n, uint64_t* __restrict__ window) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) window[i] = key[i] >> DP_COST_BITS;
}

// What the placement kernel reads per read, in sorted order: {read index, degree, first mutation, mutations | non-N
// mutations << 16}
__global__ void delta_records_kernel(const uint32_t* __restrict__ order, int64_t n, const int32_t* __restrict__ degree,
                                     const int64_t* __restrict__ rm_off, const uint8_t* __restrict__ rm_code, uint4* __restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t rid = order[i];
    const int64_t a = rm_off[rid], b = rm_off[rid + 1];
    uint32_t non_n = 0;
    for (int64_t k = a; k < b; ++k) non_n += rm_code[k] <= 4u;
    rec[i] = make_uint4(rid, (uint32_t)degree[rid], (uint32_t)a, (uint32_t)(b - a) | (non_n << 16));
}

// ... and per read mutation, at its index in the caller's mutation arrays, the postings it walks: {first posting,
// postings | allele class << 28} (one block per work unit, once the windows' minima are known).  A read with k
// allele mutations and n N's takes back at most R = 2 k + n mismatches of a state (2 where the alleles agree, 1 for
// another allele or an N), so a state must start at or below (window minimum) + R to end at or below the window
// minimum — and nothing above the window minimum can be the read's minimum.  base_w >= core base score, so the cells
// of core base score > (window minimum) + R are skipped: typically most of a slot for the reads with one or two
// mutations.  Windows that do not hold the list's core walk the whole slot.
__global__ void delta_mrec_kernel(const DeltaUnit* __restrict__ units, const DeltaGroup* __restrict__ groups,
                                  const ListDesc* __restrict__ list_desc, const int32_t* __restrict__ lpos_base,
                                  const uint32_t* __restrict__ post_off, const uint4* __restrict__ rec,
                                  const int32_t* __restrict__ rm_pos, const uint8_t* __restrict__ rm_code, int32_t prune_on,
                                  uint2* __restrict__ mrec) {
    const DeltaUnit du = units[blockIdx.x];
    const DeltaGroup dg = groups[du.group];
    const int32_t b0 = list_desc[dg.list].b0, lp = lpos_base[dg.list];
    for (int i = threadIdx.x; i < du.count; i += blockDim.x) {
        const uint4 r = rec[du.first + i];
        const int nm = (int)(r.w & 0xFFFFu), non_n = (int)(r.w >> 16);
        const int t = (dg.m0 - DP_VOFF) + nm + non_n;   // window minimum + R
        const int levels = (prune_on && dg.prune && t + 1 < DP_LEVELS) ? max(t + 1, 0) : DP_LEVELS;
        for (uint32_t k = r.z; k < r.z + (uint32_t)nm; ++k) {
            const size_t cell = (size_t)(lp + rm_pos[k] - b0) * DP_LEVELS;
            const uint32_t lo = post_off[cell], hi = post_off[cell + levels];
            uint2 m = make_uint2(lo, (hi - lo) | ((uint32_t)rm_code[k] << 28));
            // shortest posting list first (the order of a read's mutations is free): the states of the read's own
            // lineage carry its rare mutations, so they are low before the long lists of its clade-level mutations
            // are walked — and a lane that has seen a low score no longer tracks the hits above it
            uint32_t at = k;
            if (nm <= 32)
                for (; at > r.z && (mrec[at - 1].y & 0x0FFFFFFFu) > (m.y & 0x0FFFFFFFu); --at) mrec[at] = mrec[at - 1];
            mrec[at] = m;
        }
    }
}

struct WindowBaseParams {
    const Entry* state_ent;
    const int64_t* state_eoff;
    const int32_t* state_first;
    const int32_t* list_goff;     // [n_lists + 1] groups of each list ...
    const int32_t* list_gids;     // ... as indices into groups
    const DeltaGroup* groups;
    uint8_t* base;
    int32_t* whist;               // [n_groups][DP_BINS], zeroed
};

// grid (state chunks, lists): a thread takes one state and evaluates base_w for every group of its list; the
// countable nodes go to the groups' histograms warp-aggregated (the base scores of a warp take a handful of values)
__global__ void __launch_bounds__(256) window_base_kernel(const WindowBaseParams p) {
    const int l = blockIdx.y;
    const int s_lo = p.state_first[l], s_n = p.state_first[l + 1] - s_lo;
    const int g0 = p.list_goff[l], g1 = p.list_goff[l + 1];
    if ((int)(blockIdx.x * blockDim.x) >= s_n || g0 == g1) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < s_n;
    const unsigned FULL = 0xFFFFFFFFu;
    int64_t e0 = 0, e1 = 0;
    int ucnt = 0;
    if (valid) {
        e0 = p.state_eoff[s_lo + s];
        e1 = p.state_eoff[s_lo + s + 1];
        ucnt = (int)p.state_ent[e0].y;
    }
    for (int gi = g0; gi < g1; ++gi) {
        const int g = p.list_gids[gi];
        const DeltaGroup dg = p.groups[g];
        int cnt = 0;
        for (int64_t k = e0; k < e1; ++k) {
            const uint32_t z = __ldg(&p.state_ent[k].z), w = __ldg(&p.state_ent[k].w);
            const int pos = (int)(w >> 16);
            cnt += (pos >= dg.a_rel && pos <= dg.b_rel) ? (int)(z & 0xFFu) : 0;   // delta[ref] is 0 or 1
        }
        if (valid) p.base[dg.base_off + s] = (uint8_t)cnt;
        const int vmax = __reduce_max_sync(FULL, valid ? cnt : 0);
        for (int v = 0; v <= vmax; ++v) {
            const int sum = __reduce_add_sync(FULL, (valid && cnt == v) ? ucnt : 0);
            if (sum != 0 && (threadIdx.x & 31) == 0) atomicAdd(p.whist + (size_t)g * DP_BINS + DP_VOFF + v, sum);
        }
    }
}

// smallest occupied bin per group (one thread per group)
__global__ void window_m0_kernel(const int32_t* __restrict__ whist, int n_groups, DeltaGroup* __restrict__ groups) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int m0 = DP_BINS - 1;
    for (int v = DP_BINS - 1; v >= 0; --v)
        if (whist[(size_t)g * DP_BINS + v] > 0) m0 = v;
    groups[g].m0 = m0;
}

// ---- the placement kernel ----------------------------------------------------------------------------------------
struct DeltaPlaceParams {
    const DeltaUnit* units;
    int32_t n_units;
    int32_t smem_bytes;           // dynamic shared memory of the launch
    int32_t cand_cap;             // upper bound on the candidate queue entries per warp (tests shrink it to reach the re-walk)
    int* unit_counter;
    const DeltaGroup* groups;
    const uint8_t* base;
    const int32_t* whist;
    const uint2* post;
    const int32_t* state_first;
    const int64_t* sacc_off;
    const uint4* rec;             // per read in (bucket, window, heaviest first) order: delta_records_kernel
    const uint2* mrec;            // per read mutation: the posting range it touches
    int32_t* max_pars;
    int32_t* mult;
    double* saccS;
    int32_t* saccC;
    double* Gw;                   // [n_groups][DP_BINS]
    int32_t* Gc;
    uint32_t* gscratch;           // [grid][DP_WARPS][gscratch_words]: byte scratch of the reads with many mutations
    int64_t gscratch_words;
};

// Nibble scratch: state s lives in word (s / 256) * 32 + s % 32, nibble (s / 32) % 8 — 32 consecutive states are 32
// consecutive words (one per bank), where the plain layout s / 8 would put them in four words, eight lanes fighting
// over each.
// (measured: 4.16 ms against 4.21 ms for the plain layout, three instructions shorter per hit)
__device__ __forceinline__ uint32_t dp_nibble_word(uint32_t s) { return ((s >> 8) << 5) | (s & 31u); }
__device__ __forceinline__ int dp_nibble_shift(uint32_t s) { return (int)((s >> 5) & 7u) * 4; }

// One read by one warp.  FAST: nibble scratch in shared memory + candidate queue; else byte scratch in global memory
// and the postings are walked again for the weights.  rec / m = the read's record and this lane's mutation record
// (mutations 0..31; reads with more fetch the rest here).
constexpr int DP_U = 4;   // chunks of 32 postings per loop iteration (independent dependency chains)
// SINGLE (with FAST): a read with at most one mutation — every hit is its state's only one, so there is no earlier
// reduction to look up and no scratch to keep (nor to zero afterwards); the queue remembers d beside the state.
template <bool FAST, bool SINGLE = false>
__device__ __forceinline__ void dp_read(const DeltaPlaceParams& p, const DeltaGroup& dg, const uint4 rec, const uint2 m, int lane,
                                        const unsigned char* base_s, uint32_t* scr, int scr_words, int* mv, uint32_t* cand,
                                        int cand_cap, const int* whist_s, int64_t so, double (&gw)[2], int (&gc)[2]) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int nm = (int)(rec.w & 0xFFFFu), k_non_n = (int)(rec.w >> 16);
    int* n_cand_s = mv + (DP_BINS - 1);   // bins above DP_VOFF + SW_MAX_ACTIVE are never used: the last one counts candidates
    int lane_min = dg.m0;                 // the lowest score this lane has seen a state end at (tracked bins start at m0)
    for (int j0 = 0; j0 < nm; j0 += 32) {
        uint2 mm = m;
        if (j0 > 0) {
            mm = make_uint2(0u, 0u);
            if (j0 + lane < nm) mm = __ldg(p.mrec + rec.z + j0 + lane);
        }
        const int cnt = min(32, nm - j0);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t lo = __shfl_sync(FULL, mm.x, j), lc = __shfl_sync(FULL, mm.y, j);
            const uint32_t hi = lo + (lc & 0x0FFFFFFFu), c = lc >> 28;
            const uint32_t cmask = (1u << 28) | ((c >= 1u && c <= 4u) ? (1u << (23u + c)) : 0u);   // delta[ref], and the read's own allele
            // DP_U chunks of 32 postings per iteration, loaded one iteration ahead
            uint2 nxt[DP_U];
#pragma unroll
            for (int h = 0; h < DP_U; ++h) {
                nxt[h] = make_uint2(0u, 0u);
                if (lo + 32 * h + lane < hi) nxt[h] = __ldg(p.post + lo + 32 * h + lane);
            }
            lane_min = __reduce_min_sync(FULL, lane_min);   // what any lane has seen, once per posting list
            for (uint32_t i0 = lo; i0 < hi; i0 += 32 * DP_U) {
                uint2 e[DP_U];
                bool act[DP_U];
                uint32_t s[DP_U];
                int d[DP_U], v_new[DP_U];
#pragma unroll
                for (int h = 0; h < DP_U; ++h) {
                    e[h] = nxt[h];   // (a lane past the end holds x = 0: no allele bit, delta[ref] = 0, so d = 0)
                    nxt[h] = make_uint2(0u, 0u);
                    if (i0 + 32 * (DP_U + h) + lane < hi) nxt[h] = __ldg(p.post + i0 + 32 * (DP_U + h) + lane);
                    s[h] = e[h].x & 0xFFFFFFu;
                    d[h] = __popc(e[h].x & cmask);   // delta[ref] - delta[c]
                    act[h] = d[h] > 0;
                }
#pragma unroll
                for (int h = 0; h < DP_U; ++h) {
                    v_new[h] = DP_BINS;
                    if (act[h]) {
                        int oldred;
                        if (SINGLE) {
                            oldred = 0;
                        } else if (FAST) {
                            const int sh = dp_nibble_shift(s[h]);
                            oldred = (int)((atomicAdd(scr + dp_nibble_word(s[h]), (uint32_t)d[h] << sh) >> sh) & 15u);
                        } else {
                            const int sh = (int)(s[h] & 3u) * 8;
                            oldred = (int)((atomicAdd(scr + (s[h] >> 2), (uint32_t)d[h] << sh) >> sh) & 255u);
                        }
                        v_new[h] = (int)base_s[s[h]] + DP_VOFF - oldred - d[h];
                    }
                }
                // Only the bins at or below the window's own minimum m0 can hold the read's minimum (a touched state
                // only moves down), so only hits that end there are tracked: the state's nodes leave the tracked bin
                // they were in (if any) and enter the new one, and the state is remembered — the read's weight goes
                // to it once the minimum is known.  (Measured dead ends: keeping the top bins per lane in registers and
                // taking queue slots from a ballot instead of these same-address atomics — 5.26 vs 4.87 ms, the
                // unconditional instructions cost more than the serialised atomics of the ~10 % of hits that get here.)
                // A lane that has already seen a lower score skips a hit above it: the bins above the read's minimum
                // may then be off (an arrival not recorded, its departure later taken from a bin it never entered),
                // the minimum's bin and everything below it are not — a hit that ends at the final minimum is never
                // above anything its lane has seen.
#pragma unroll
                for (int h = 0; h < DP_U; ++h) {
                    if (v_new[h] <= lane_min) {
                        lane_min = v_new[h];
                        if (v_new[h] + d[h] <= dg.m0) atomicSub(&mv[v_new[h] + d[h]], (int)e[h].y);
                        atomicAdd(&mv[v_new[h]], (int)e[h].y);
                        if (FAST) {
                            const int slot = atomicAdd(n_cand_s, 1);
                            if (slot < cand_cap) cand[slot] = SINGLE ? (s[h] | ((uint32_t)d[h] << 24)) : s[h];
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    // occupied bins at or below m0 = the window's (m0 itself) + the tracked moves; minimum and its countable nodes
    int tot[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int v = lane + 32 * hf;
        tot[hf] = v <= dg.m0 ? whist_s[v] + mv[v] : 0;
    }
    const int n_cand = *n_cand_s;
    __syncwarp();
    mv[lane] = 0;
    mv[lane + 32] = 0;
    const uint32_t o0 = __ballot_sync(FULL, tot[0] > 0), o1 = __ballot_sync(FULL, tot[1] > 0);
    const int mV = o0 ? __ffs(o0) - 1 : (o1 ? 32 + __ffs(o1) - 1 : DP_VOFF);
    const int n_epp = (o0 | o1) ? __shfl_sync(FULL, mV < 32 ? tot[0] : tot[1], mV & 31) : 0;
    const int pars = k_non_n + mV - DP_VOFF;
    const int deg = (int)rec.y;
    double wgt = 0.0;
    if (n_epp > 0) wgt = (double)deg / ((double)(1 + pars) * (double)n_epp);   // node_score, initial_filter.hpp:54-57
    if (lane == 0) {
        p.max_pars[rec.x] = pars;
        p.mult[rec.x] = n_epp;
    }
    if (n_epp > 0 && lane == (mV & 31)) {
        // (selects, not gw[mV >> 5]: a dynamic index would put the two sums into local memory)
        const bool low = mV < 32;
        gw[0] += low ? wgt : 0.0;
        gw[1] += low ? 0.0 : wgt;
        gc[0] += low ? deg : 0;
        gc[1] += low ? 0 : deg;
    }
    // touched states at the minimum; scratch back to zero
    if (FAST && n_cand <= cand_cap) {
        for (int i = lane; i < n_cand; i += 32) {
            const uint32_t cs = cand[i], s = cs & 0xFFFFFFu;
            int red;
            if (SINGLE) {
                red = (int)(cs >> 24);
            } else {
                const int sh = dp_nibble_shift(s);
                red = (int)((atomicAnd(scr + dp_nibble_word(s), ~(15u << sh)) >> sh) & 15u);   // duplicates see 0
            }
            if (red && n_epp > 0 && (int)base_s[s] + DP_VOFF - red == mV) {
                atomicAdd(p.saccS + so + s, wgt);
                atomicAdd(p.saccC + so + s, deg);
            }
        }
        if (nm > 0 && !SINGLE) {
            __syncwarp();
            uint4* z = reinterpret_cast<uint4*>(scr);
            for (int i = lane; i < scr_words / 4; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
        for (int j0 = 0; j0 < nm; j0 += 32) {
            uint2 mm = make_uint2(0u, 0u);
            if (j0 + lane < nm) mm = __ldg(p.mrec + rec.z + j0 + lane);
            const int cnt = min(32, nm - j0);
            for (int j = 0; j < cnt; ++j) {
                const uint32_t lo = __shfl_sync(FULL, mm.x, j), hi = lo + (__shfl_sync(FULL, mm.y, j) & 0x0FFFFFFFu);
                uint32_t nx = 0u;
                if (lo + lane < hi) nx = __ldg(&p.post[lo + lane].x);
                const uint32_t c2 = __shfl_sync(FULL, mm.y, j) >> 28;
                const uint32_t cmask2 = (1u << 28) | ((c2 >= 1u && c2 <= 4u) ? (1u << (23u + c2)) : 0u);
                for (uint32_t i = lo + lane; i < hi; i += 32) {
                    const uint32_t x = nx, s = nx & 0xFFFFFFu;
                    if (i + 32 < hi) nx = __ldg(&p.post[i + 32].x);
                    int red;
                    if (SINGLE) {
                        red = __popc(x & cmask2);
                    } else if (FAST) {
                        const int sh = dp_nibble_shift(s);
                        red = (int)((atomicAnd(scr + dp_nibble_word(s), ~(15u << sh)) >> sh) & 15u);
                    } else {
                        const int sh = (int)(s & 3u) * 8;
                        red = (int)((atomicAnd(scr + (s >> 2), ~(255u << sh)) >> sh) & 255u);
                    }
                    if (red && n_epp > 0 && (int)base_s[s] + DP_VOFF - red == mV) {
                        atomicAdd(p.saccS + so + s, wgt);
                        atomicAdd(p.saccC + so + s, deg);
                    }
                }
            }
        }
    }
    __syncwarp();
}

template <int DP_WARPS>
__global__ void __launch_bounds__(DP_WARPS * 32, DP_WARPS_SM / DP_WARPS) delta_place_kernel(const DeltaPlaceParams p) {
    constexpr int DP_FIXED = dp_fixed(DP_WARPS);
    extern __shared__ __align__(16) unsigned char smem[];
    int* ctrl = reinterpret_cast<int*>(smem);                       // [0] unit, [1] next read of the unit
    int* whist_s = reinterpret_cast<int*>(smem + 16);
    int* mv_all = whist_s + DP_BINS;
    unsigned char* base_s = smem + DP_FIXED;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    int* mv = mv_all + warp * DP_BINS;
    for (int i = threadIdx.x; i < DP_WARPS * DP_BINS; i += blockDim.x) mv_all[i] = 0;
    uint32_t* gscr = p.gscratch + ((size_t)blockIdx.x * DP_WARPS + warp) * (size_t)p.gscratch_words;
    int scr_list = -1;   // the list this warp's nibble scratch is laid out (and zero) for
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            ctrl[0] = atomicAdd(p.unit_counter, 1);
            ctrl[1] = 0;
        }
        __syncthreads();
        const int u = ctrl[0];
        if (u >= p.n_units) break;
        const DeltaUnit du = p.units[u];
        const DeltaGroup dg = p.groups[du.group];
        const int s_n = p.state_first[dg.list + 1] - p.state_first[dg.list];
        const int s_pad = (s_n + 15) & ~15;
        const int stride = (s_n + 255) / 256 * 128;                 // nibble scratch bytes per warp (dp_nibble_word)
        const int aw = min(DP_WARPS, (p.smem_bytes - DP_FIXED - s_pad) / (stride + DP_CAND_MIN * 4));
        // the shared memory the scratch areas leave is the warps' candidate queues
        const int cand_cap = min(p.cand_cap, (p.smem_bytes - DP_FIXED - s_pad - aw * stride) / (aw * 16) * 4);
        {   // the window's base scores and histogram
            const uint4* src = reinterpret_cast<const uint4*>(p.base + dg.base_off);
            uint4* dst = reinterpret_cast<uint4*>(base_s);
            for (int i = threadIdx.x; i < s_pad / 16; i += blockDim.x) dst[i] = __ldg(src + i);
            if (threadIdx.x < DP_BINS) whist_s[threadIdx.x] = p.whist[(size_t)du.group * DP_BINS + threadIdx.x];
        }
        uint32_t* scr = reinterpret_cast<uint32_t*>(base_s + s_pad + (size_t)warp * stride);
        uint32_t* cand = reinterpret_cast<uint32_t*>(base_s + s_pad + (size_t)aw * stride) + (size_t)warp * cand_cap;
        if (warp < aw && dg.list != scr_list) {   // another list: the layout moved, its scratch area starts out zero
            uint4* z = reinterpret_cast<uint4*>(scr);
            for (int i = lane; i < stride / 16; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        scr_list = warp < aw ? dg.list : -1;
        __syncthreads();
        if (warp < aw) {
            const int64_t so = p.sacc_off[dg.bucket];
            double gw[2] = {0.0, 0.0};
            int gc[2] = {0, 0};
            // the warps claim the unit's reads (sorted heaviest first) from a shared counter, two reads ahead: the
            // record of the read after next and the mutation records of the next read are in flight while a read
            // is processed
            const uint4* rec = p.rec + du.first;
            auto claim = [&]() {
                int r = 0;
                if (lane == 0) r = atomicAdd(&ctrl[1], 1);
                return __shfl_sync(FULL, r, 0);
            };
            int i1 = claim(), i2 = claim();
            uint4 r1 = make_uint4(0u, 0u, 0u, 0u), r2 = r1;
            uint2 m1 = make_uint2(0u, 0u);
            if (i1 < du.count) r1 = __ldg(rec + i1);
            if (i2 < du.count) r2 = __ldg(rec + i2);
            if (lane < (int)(r1.w & 0xFFFFu)) m1 = __ldg(p.mrec + r1.z + lane);
            while (i1 < du.count) {
                const int i3 = claim();
                uint4 r3 = make_uint4(0u, 0u, 0u, 0u);
                if (i3 < du.count) r3 = __ldg(rec + i3);
                uint2 m2 = make_uint2(0u, 0u);
                if (lane < (int)(r2.w & 0xFFFFu)) m2 = __ldg(p.mrec + r2.z + lane);
                if ((int)(r1.w & 0xFFFFu) <= 1) dp_read<true, true>(p, dg, r1, m1, lane, base_s, scr, stride / 4, mv, cand, cand_cap, whist_s, so, gw, gc);
                else if ((int)(r1.w & 0xFFFFu) <= DP_FAST_MUTS) dp_read<true>(p, dg, r1, m1, lane, base_s, scr, stride / 4, mv, cand, cand_cap, whist_s, so, gw, gc);
                else dp_read<false>(p, dg, r1, m1, lane, base_s, gscr, 0, mv, cand, 0, whist_s, so, gw, gc);
                i1 = i2; r1 = r2; m1 = m2;
                i2 = i3; r2 = r3;
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (gc[hf] != 0) {
                    atomicAdd(p.Gw + (size_t)du.group * DP_BINS + lane + 32 * hf, gw[hf]);
                    atomicAdd(p.Gc + (size_t)du.group * DP_BINS + lane + 32 * hf, gc[hf]);
                }
            }
        }
    }
}

// per-(bucket, state) accumulators += the weights of the reads whose minimum is the state's base score in their
// window and that do not touch the state: sum over the bucket's groups of G[group][base_group(s)]
__global__ void delta_finalize_kernel(const DeltaGroup* __restrict__ groups, const int32_t* __restrict__ bucket_goff,
                                      const int32_t* __restrict__ state_first, const BucketDesc* __restrict__ buckets,
                                      const uint8_t* __restrict__ base, const double* __restrict__ Gw,
                                      const int32_t* __restrict__ Gc, const int64_t* __restrict__ sacc_off,
                                      double* __restrict__ saccS, int32_t* __restrict__ saccC) {
    const int b = blockIdx.y;
    const int g0 = bucket_goff[b], g1 = bucket_goff[b + 1];
    if (g0 == g1) return;
    const int l = buckets[b].list;
    const int s_n = state_first[l + 1] - state_first[l];
    const int64_t so = sacc_off[b];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < s_n; s += gridDim.x * blockDim.x) {
        double ws = 0.0;
        int cs = 0;
        for (int g = g0; g < g1; ++g) {
            const int v = (int)base[groups[g].base_off + s] + DP_VOFF;
            ws += Gw[(size_t)g * DP_BINS + v];
            cs += Gc[(size_t)g * DP_BINS + v];
        }
        if (cs != 0) {
            saccS[so + s] += ws;
            saccC[so + s] += cs;
        }
    }
}

}  // namespace wepp
