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
//   post_*_kernel          posting lists: per (list, position) the states that mutate it (state, allele class, nodes)
//   delta_keys_kernel      sort key (bucket, window) per read -> cub sort + run-length encode = window groups
//   window_base_kernel     base_w(s) for every group of a list, and the groups' histograms
//   delta_place_kernel     persistent CTAs pull work units (<= 512 reads of one group); a warp takes a read
//   delta_finalize_kernel  per-(bucket, state) accumulators += sum over the bucket's groups of G[group][base(s)] and of the
//                          W cells of the state's positions
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

constexpr int DP_BINS = 64;           // score bins: bin = base - red + DP_VOFF
constexpr int DP_VOFF = SW_MAX_ACTIVE;   // base <= SW_MAX_ACTIVE and red <= 2 * base
constexpr int DP_UNIT = 256;          // reads per work unit (a group of up to 2 * DP_UNIT reads stays whole)
static_assert(DP_VOFF + SW_MAX_ACTIVE + 1 <= DP_BINS, "score bins");
constexpr uint32_t DP_X_NONE = 7u;    // allele class of a state entry that equals no read allele (IUPAC union)

struct DeltaGroup {     // reads of one bucket with the same window
    int64_t base_off;   // first byte of base_w in the base buffer (S_list bytes, 16-byte aligned)
    int32_t list, bucket;
    int32_t a_rel, b_rel;   // window relative to the list's first position
    int32_t m0;         // smallest occupied bin of the window's histogram
    int32_t clip;       // min(a_rel, 16) | min(width - 1 - b_rel, 16) << 8: the margins a state must keep clear (see postings)
    int64_t w_off;      // first cell of the window's W rows: [b_rel - a_rel + 1][5 read alleles][3 bins]
};
struct DeltaW {         // weight and degree sums of the reads with one (window, position, allele, m0 - minimum)
    double w;
    int32_t c, pad;
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
// (list, position) slot.  Sorting by key = slot << 16 | allele class << 8 | entries of the state groups the postings by
// slot and, inside a slot, puts states with the same allele and the same number of mutations next to each other: the
// 32 states a warp touches together then mostly share one (base score, red), and their nodes move between the
// histogram bins with one warp sum.  The empty state's placeholder entry sorts to the end (key = all ones).
//
// What a posting says about its state beyond (state, allele class, delta[ref], countable nodes) lets the placement
// kernel decide most hits from registers alone:
//   margins   lo4 / hi4 = distance of the state's first / last mismatching position (delta[ref] = 1) from the list's
//             first / last position, saturated at 15: a window [a, b] with a <= lo4 and width - 1 - b <= hi4 holds
//             all of them, and base_w(s) is the state's full base score (5 bits) — no look-up
//   bloom     bit (q mod 13) and bit 13 + (q mod 14) for every position q the state mutates: a read whose OTHER
//             mutations set none of these pairs cannot touch the state a second time, so red(r, s) is this hit's alone
__device__ __forceinline__ uint32_t dp_bloom(uint32_t q) { return (1u << (q % 13u)) | (1u << (13u + q % 14u)); }
constexpr uint32_t DP_BLOOM1 = 0x1FFFu, DP_BLOOM2 = 0x7FFE000u;   // the two hash ranges
constexpr uint32_t DP_NODES_MASK = 0x0FFFFFFFu;                   // arena nodes < 2^28

__global__ void post_pairs_kernel(const Entry* __restrict__ state_ent, const int64_t* __restrict__ state_eoff,
                                  const int32_t* __restrict__ state_list, const int32_t* __restrict__ state_first,
                                  const int32_t* __restrict__ lpos_base, const ListDesc* __restrict__ list_desc, int32_t n_states,
                                  uint32_t* __restrict__ slot_count, uint64_t* __restrict__ pkey, uint64_t* __restrict__ pval,
                                  uint32_t* __restrict__ pmask, int32_t* __restrict__ bad) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    const int l = state_list[s];
    const uint32_t local = (uint32_t)(s - state_first[l]);
    const int32_t lp = lpos_base[l];
    const uint32_t width = (uint32_t)list_desc[l].width;
    const int64_t e0 = state_eoff[s], e1 = state_eoff[s + 1];
    const uint32_t len = (uint32_t)min((int64_t)255, e1 - e0);
    uint32_t lo = 15u, hi = 15u, full = 0u, bloom = 0u;
    for (int64_t k = e0; k < e1; ++k) {
        const Entry e = state_ent[k];
        if (e.z == 0u && (e.w & 0xFFu) == 0u) continue;
        const uint32_t q = e.w >> 16;
        bloom |= dp_bloom(q);
        if (e.z & 0xFFu) {
            ++full;
            lo = min(lo, q);
            hi = min(hi, width - 1u - q);
        }
    }
    if (full > 31u) *bad = 1;
    for (int64_t k = e0; k < e1; ++k) {
        const Entry e = state_ent[k];
        pkey[k] = ~0ull;
        pval[k] = 0ull;
        pmask[k] = 0u;
        if (e.z == 0u && (e.w & 0xFFu) == 0u) continue;   // the empty state's placeholder entry
        const uint32_t xc = dp_table_class(e.z, e.w);
        if (xc == 0xFFu || local >= (1u << 24) || (e.y & ~DP_NODES_MASK)) {
            *bad = 1;
            continue;
        }
        const uint32_t slot = (uint32_t)lp + (e.w >> 16);
        atomicAdd(slot_count + slot, 1u);
        pkey[k] = ((uint64_t)slot << 16) | ((uint64_t)xc << 8) | (uint64_t)len;
        // uint2 {state | class << 24 (bit 27 = delta[ref]) | lo4 << 28, nodes | hi4 << 28}; mask word = bloom | full base << 27
        pval[k] = (uint64_t)(local | (xc << 24) | (lo << 28)) | ((uint64_t)(e.y | (hi << 28)) << 32);
        pmask[k] = bloom | (full << 27);
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
                                  const uint32_t* __restrict__ post_off, const int32_t* __restrict__ lpos_base,
                                  uint64_t* __restrict__ key, uint32_t* __restrict__ val) {
    const TileDesc td = tiles[blockIdx.x];
    const int32_t list = buckets[td.bucket].list;
    const int32_t b0 = list_desc[list].b0, lp = lpos_base[list];
    for (int i = threadIdx.x; i < td.count; i += blockDim.x) {
        const int64_t rid = perm[td.first + i];
        uint32_t cost = 0;
        for (int64_t k = rm_off[rid]; k < rm_off[rid + 1]; ++k) {
            const int slot = lp + rm_pos[k] - b0;
            cost += post_off[slot + 1] - post_off[slot] + 16u;
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
// mutations << 16}; and per read mutation (at its index in the caller's mutation arrays) the posting range it
// touches: {first posting, postings | allele class << 28, bloom bits of the read's OTHER mutations, position - list start}.
__global__ void delta_records_kernel(const uint64_t* __restrict__ sorted_key, const uint32_t* __restrict__ order, int64_t n,
                                     const BucketDesc* __restrict__ buckets, const ListDesc* __restrict__ list_desc,
                                     const int32_t* __restrict__ lpos_base, const uint32_t* __restrict__ post_off,
                                     const int32_t* __restrict__ degree, const int64_t* __restrict__ rm_off,
                                     const int32_t* __restrict__ rm_pos, const uint8_t* __restrict__ rm_code,
                                     uint4* __restrict__ rec, uint4* __restrict__ mrec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t rid = order[i];
    const int32_t list = buckets[(int32_t)(sorted_key[i] >> (24 + DP_COST_BITS))].list;
    const int32_t b0 = list_desc[list].b0, lp = lpos_base[list];
    const int64_t a = rm_off[rid], b = rm_off[rid + 1];
    uint32_t non_n = 0, all = 0, twice = 0;   // bloom bits set by any / by two or more of the read's mutations
    for (int64_t k = a; k < b; ++k) {
        const uint32_t bits = dp_bloom((uint32_t)(rm_pos[k] - b0));
        twice |= all & bits;
        all |= bits;
    }
    for (int64_t k = a; k < b; ++k) {
        const uint32_t c = rm_code[k];
        non_n += c <= 4u;
        const uint32_t lo = post_off[lp + rm_pos[k] - b0], hi = post_off[lp + rm_pos[k] - b0 + 1];
        const uint32_t own = dp_bloom((uint32_t)(rm_pos[k] - b0));
        mrec[k] = make_uint4(lo, (hi - lo) | (c << 28), (all & ~own) | (own & twice), (uint32_t)(rm_pos[k] - b0));
    }
    rec[i] = make_uint4(rid, (uint32_t)degree[rid], (uint32_t)a, (uint32_t)(b - a) | (non_n << 16));
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
    int32_t cand_cap;             // candidate queue entries per warp, <= DP_CAND (tests shrink it to reach the slow path)
    int32_t tq_cap;               // touched-queue entries per warp, <= DP_TQ (tests shrink it to reach the scratch scan)
    int* unit_counter;
    const DeltaGroup* groups;
    const uint8_t* base;
    const int32_t* whist;
    const uint2* post;            // per posting: {state | class << 24 | delta[ref] << 27 | lo4 << 28, nodes | hi4 << 28}
    const uint32_t* postm;        // per posting: bloom bits of the state's positions | full base score << 27
    const int32_t* state_first;
    const int32_t* state_nodes;   // per state: countable nodes
    const int64_t* sacc_off;
    const uint4* rec;             // per read in (bucket, window, heaviest first) order: delta_records_kernel
    const uint4* mrec;            // per read mutation: the posting range it touches, bloom bits of the read's other mutations
    int32_t* max_pars;
    int32_t* mult;
    double* saccS;
    int32_t* saccC;
    double* Gw;                   // [n_groups][DP_BINS]
    int32_t* Gc;
    DeltaW* W;                    // per window group: [window positions][5][3], zeroed per launch
    uint32_t* gscratch;           // [grid][DP_WARPS][gscratch_words]: byte scratch of the reads that take the slow path
    int64_t gscratch_words;
    uint32_t* gspill;             // [grid][DP_WARPS][2][spill_cap]: what the two queues in shared memory cannot hold
    int32_t spill_cap;
    unsigned long long* stats;    // DP_STATS counters (WEPP_DELTA_STATS=1), else null
};
enum { DP_ST_HITS = 0, DP_ST_MULTI, DP_ST_CLIPPED, DP_ST_TRACKED, DP_ST_SLOW_READS, DP_ST_TAB_ENTRIES, DP_ST_TQ_OVER, DP_ST_CAND, DP_STATS };

// minimum and its countable nodes from the tracked bins, the read's weight (node_score, initial_filter.hpp:54-57);
// per-read results out; the weight of the untouched states at the minimum goes to the group's per-score sums
struct DpMin {
    int mV, n_epp, deg;
    double wgt;
};
__device__ __forceinline__ DpMin dp_minimum(int32_t* max_pars, int32_t* mult, int m0, const uint4 rec, int lane, int* mv,
                                            const int* whist_s, double& gw0, double& gw1, int& gc0, int& gc1) {
    const unsigned FULL = 0xFFFFFFFFu;
    int tot[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        const int v = lane + 32 * hf;
        tot[hf] = v <= m0 ? whist_s[v] + mv[v] : 0;
    }
    __syncwarp();
    mv[lane] = 0;
    mv[lane + 32] = 0;
    const uint32_t o0 = __ballot_sync(FULL, tot[0] > 0), o1 = __ballot_sync(FULL, tot[1] > 0);
    DpMin r;
    r.mV = o0 ? __ffs(o0) - 1 : (o1 ? 32 + __ffs(o1) - 1 : DP_VOFF);
    r.n_epp = (o0 | o1) ? __shfl_sync(FULL, r.mV < 32 ? tot[0] : tot[1], r.mV & 31) : 0;
    const int pars = (int)(rec.w >> 16) + r.mV - DP_VOFF;
    r.deg = (int)rec.y;
    r.wgt = 0.0;
    if (r.n_epp > 0) r.wgt = (double)r.deg / ((double)(1 + pars) * (double)r.n_epp);
    if (lane == 0) {
        max_pars[rec.x] = pars;
        mult[rec.x] = r.n_epp;
    }
    if (r.n_epp > 0 && lane == (r.mV & 31)) {
        if (r.mV < 32) {
            gw0 += r.wgt;
            gc0 += r.deg;
        } else {
            gw1 += r.wgt;
            gc1 += r.deg;
        }
    }
    return r;
}

// One read by one warp, the general way: every hit adds its red to a byte per state in global memory (the warp's own
// area, left zero), hits that end at or below the window's minimum m0 move their nodes between the tracked bins, and
// the postings are walked a second time for the weights.  Serves reads with any number of mutations and any number
// of multi-hit states; the fast path below hands over the few reads it cannot finish.
struct DpSlowArgs {   // by value: a reference to the kernel's parameter block would move it to local memory
    const uint2* post;
    const uint4* mrec;
    int32_t *max_pars, *mult;
    double* saccS;
    int32_t* saccC;
    double* Gw;          // this group's row
    int32_t* Gc;
    int m0;
};
__device__ __noinline__ void dp_read_slow(const DpSlowArgs p, const uint4 rec, int lane, const unsigned char* base_s, uint32_t* scr,
                                          int* mv, const int* whist_s) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int nm = (int)(rec.w & 0xFFFFu);
    double gw0 = 0.0, gw1 = 0.0;   // (its own: the caller's sums stay in registers)
    int gc0 = 0, gc1 = 0;
    for (int j0 = 0; j0 < nm; j0 += 32) {
        uint4 mm = make_uint4(0u, 0u, 0u, 0u);
        if (j0 + lane < nm) mm = __ldg(p.mrec + rec.z + j0 + lane);
        const int cnt = min(32, nm - j0);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t lo = __shfl_sync(FULL, mm.x, j), lc = __shfl_sync(FULL, mm.y, j);
            const uint32_t hi = lo + (lc & 0x0FFFFFFFu), c = lc >> 28;
            for (uint32_t i0 = lo + lane; i0 < hi; i0 += 128) {   // four chunks per iteration: the atomics' round trips overlap
                uint2 e[4];
                int d[4], oldred[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    e[h] = make_uint2(0u, 0u);
                    if (i0 + 32 * h < hi) e[h] = __ldg(p.post + i0 + 32 * h);
                    d[h] = i0 + 32 * h < hi ? (int)((e[h].x >> 27) & 1u) + (int)(((e[h].x >> 24) & 7u) == c) : 0;   // delta[ref] - delta[c]
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t s = e[h].x & 0xFFFFFFu;
                    const int sh = (int)(s & 3u) * 8;
                    oldred[h] = 0;
                    if (d[h] > 0) oldred[h] = (int)((atomicAdd(scr + (s >> 2), (uint32_t)d[h] << sh) >> sh) & 255u);
                }
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t s = e[h].x & 0xFFFFFFu;
                    const int v_new = (int)base_s[s] + DP_VOFF - oldred[h] - d[h];
                    if (d[h] > 0 && v_new <= p.m0) {
                        const int nodes = (int)(e[h].y & DP_NODES_MASK);
                        if (v_new + d[h] <= p.m0) atomicSub(&mv[v_new + d[h]], nodes);
                        atomicAdd(&mv[v_new], nodes);
                    }
                }
            }
            __syncwarp();   // a state hit again by the next mutation sees this one's red
        }
    }
    __syncwarp();
    const DpMin r = dp_minimum(p.max_pars, p.mult, p.m0, rec, lane, mv, whist_s, gw0, gw1, gc0, gc1);
    for (int j0 = 0; j0 < nm; j0 += 32) {
        uint4 mm = make_uint4(0u, 0u, 0u, 0u);
        if (j0 + lane < nm) mm = __ldg(p.mrec + rec.z + j0 + lane);
        const int cnt = min(32, nm - j0);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t lo = __shfl_sync(FULL, mm.x, j), hi = lo + (__shfl_sync(FULL, mm.y, j) & 0x0FFFFFFFu);
            for (uint32_t i0 = lo + lane; i0 < hi; i0 += 128) {
                uint32_t s[4];
                int red[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) s[h] = i0 + 32 * h < hi ? (__ldg(&p.post[i0 + 32 * h].x) & 0xFFFFFFu) : 0xFFFFFFFFu;
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int sh = (int)(s[h] & 3u) * 8;
                    red[h] = 0;
                    if (s[h] != 0xFFFFFFFFu) red[h] = (int)((atomicAnd(scr + (s[h] >> 2), ~(255u << sh)) >> sh) & 255u);   // a second visit sees 0
                }
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    if (red[h] && r.n_epp > 0 && (int)base_s[s[h]] + DP_VOFF - red[h] == r.mV) {
                        atomicAdd(p.saccS + s[h], r.wgt);
                        atomicAdd(p.saccC + s[h], r.deg);
                    }
            }
        }
    }
    if (gc0 != 0) {
        atomicAdd(p.Gw + lane, gw0);
        atomicAdd(p.Gc + lane, gc0);
    }
    if (gc1 != 0) {
        atomicAdd(p.Gw + lane + 32, gw1);
        atomicAdd(p.Gc + lane + 32, gc1);
    }
    __syncwarp();
}

// Per-warp shared memory of the fast path: the tracked bins, counters, the queue of possibly multi-hit states the read
// touched and the queue of the truly multi-hit ones that ended at or below m0 — and, after the window's base scores, a
// nibble per state (the red of the possibly multi-hit states only; all zero between reads).
constexpr int DP_WARPS = 8;           // warps per CTA
constexpr int DP_CTAS = 2;            // CTAs per SM the registers are bounded for
constexpr int DP_TQ = 1024;           // touched-queue entries per warp (more: the scratch is scanned instead)
constexpr int DP_CAND = 128;          // queue entries per warp: multi-hit state | final bin << 24
constexpr int DP_SPILL = 8192;         // spill entries per warp and queue in global memory (more: the slow path)
constexpr int DP_FAST_MUTS = 7;       // reads with more mutations take the slow path (a nibble holds red <= 15)
constexpr int DP_U = 4;               // chunks of 32 postings per loop iteration (independent dependency chains)
constexpr int DP_WARP_BYTES = DP_BINS * 4 + 16 + DP_TQ * 4 + DP_CAND * 4;
constexpr int DP_FIXED = 16 + DP_BINS * 4 + DP_WARPS * DP_WARP_BYTES;   // ctrl, whist, the warps' areas; then base_w, scratch
static_assert(DP_WARP_BYTES % 16 == 0, "per-warp areas stay 16-byte aligned");
__host__ __device__ inline int dp_scratch_stride(int s_n) { return ((s_n + 7) / 8 * 4 + 15) & ~15; }   // nibble scratch bytes per warp

// One read by one warp, mostly from registers.
//  * A hit (posting of one of the read's mutations (q, c)) whose state no OTHER mutation of the read can touch (bloom
//    test) has red = d, so its final bin b_w - d is known at once; base_w comes from the posting when the window holds
//    all of the state's mismatching positions (margins), else from shared memory.  b_w >= the window's minimum m0 and
//    d <= 2, so such a state ends in bin m0, m0 - 1 or m0 - 2 or is of no interest: three per-lane sums of countable
//    nodes, reduced over the warp after the walk, are all the walk keeps of them.  Their WEIGHTS are not handed out
//    state by state either: the read adds its weight once per mutation to W[window][q][c][m0 - min], and
//    delta_finalize_kernel gives every state s that mutates q the entries with b_w(s) - d(s, q, c) = min — exactly the
//    single-hit states at the read's minimum (a state the read hits twice has b_w - d above its final bin, hence above
//    the minimum, at each of the positions).
//  * The possibly multi-hit states accumulate red in the warp's nibble scratch, are remembered in the touched queue
//    (with the first hit's d: red > d tells a second hit) and are settled after the walk; the truly multi-hit ones at
//    the minimum get the weight directly.
// Returns false — everything left as it was found — when the multi-hit queue cannot hold the read: dp_read_slow then
// takes it (it hands out all weights directly and adds nothing to W).
template <bool STATS>
__device__ __forceinline__ bool dp_read(const DeltaPlaceParams& p, const DeltaGroup& dg, const uint4 rec, const uint4 m, int lane,
                                        const unsigned char* base_s, const int32_t* __restrict__ snodes, uint32_t* scr, int scr_words,
                                        int* mv, int* wctr, uint32_t* tq, uint32_t* cand, uint32_t* gsp, const int* whist_s, int64_t so,
                                        double& gw0, double& gw1, int& gc0, int& gc1) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int nm = (int)(rec.w & 0xFFFFu);   // <= DP_FAST_MUTS < 32: one mutation record per lane
    uint32_t* gsp_tq = gsp;
    uint32_t* gsp_cand = gsp + p.spill_cap;
    const uint32_t clipA = (uint32_t)dg.clip & 0xFFu, clipB = (uint32_t)dg.clip >> 8;
    const int m0 = dg.m0;
    unsigned st_hits = 0, st_multi = 0, st_clip = 0, st_track = 0;
    int acc0 = 0, acc1 = 0, acc2 = 0;   // countable nodes entering bins m0, m0 - 1, m0 - 2 (minus those leaving m0)
    for (int j = 0; j < nm; ++j) {
        const uint32_t lo = __shfl_sync(FULL, m.x, j), lc = __shfl_sync(FULL, m.y, j), rm = __shfl_sync(FULL, m.z, j);
        const uint32_t hi = lo + (lc & 0x0FFFFFFFu), c = lc >> 28;
        const uint32_t rm1 = rm & DP_BLOOM1, rm2 = rm & DP_BLOOM2;
        uint2 nxt[DP_U];
        uint32_t nxm[DP_U];
#pragma unroll
        for (int h = 0; h < DP_U; ++h) {
            nxt[h] = make_uint2(0u, 0u);
            nxm[h] = 0u;
            if (lo + 32 * h + lane < hi) {
                nxt[h] = __ldg(p.post + lo + 32 * h + lane);
                nxm[h] = __ldg(p.postm + lo + 32 * h + lane);
            }
        }
        for (uint32_t i0 = lo; i0 < hi; i0 += 32 * DP_U) {
            uint2 e[DP_U];
            uint32_t em[DP_U];
#pragma unroll
            for (int h = 0; h < DP_U; ++h) {
                e[h] = nxt[h];
                em[h] = nxm[h];
                nxt[h] = make_uint2(0u, 0u);
                nxm[h] = 0u;
                if (i0 + 32 * (DP_U + h) + lane < hi) {
                    nxt[h] = __ldg(p.post + i0 + 32 * (DP_U + h) + lane);
                    nxm[h] = __ldg(p.postm + i0 + 32 * (DP_U + h) + lane);
                }
            }
#pragma unroll
            for (int h = 0; h < DP_U; ++h) {
                const uint32_t s = e[h].x & 0xFFFFFFu;
                const int d = (int)((e[h].x >> 27) & 1u) + (int)(((e[h].x >> 24) & 7u) == c);   // delta[ref] - delta[c]
                const bool act = i0 + 32 * h + lane < hi && d > 0;
                const bool multi = act && (em[h] & rm1) != 0u && (em[h] & rm2) != 0u;
                const bool single = act && !multi;
                const bool clipped = (e[h].x >> 28) < clipA || (e[h].y >> 28) < clipB;
                int b = (int)(em[h] >> 27);
                if (single && clipped) b = (int)base_s[s];
                const int o = m0 - (b + DP_VOFF - d);   // how far below m0 the state ends
                const int nodes = single ? (int)(e[h].y & DP_NODES_MASK) : 0;
                acc0 += (o == 0 ? nodes : 0) - (o == d ? nodes : 0);   // (o == d: its base score is m0 — the only tracked bin it can have been in)
                acc1 += o == 1 ? nodes : 0;
                acc2 += o == 2 ? nodes : 0;
                if (STATS) {
                    st_hits += act;
                    st_multi += multi;
                    st_clip += single && clipped;
                    st_track += single && o >= 0;
                }
                if (multi) {
                    const int sh = (int)(s & 7u) * 4;
                    const uint32_t old = (atomicAdd(scr + (s >> 3), (uint32_t)d << sh) >> sh) & 15u;
                    if (old == 0u) {   // first touch
                        const int q = atomicAdd(&wctr[1], 1);
                        if (q < p.tq_cap) tq[q] = s | ((uint32_t)d << 24);
                        else if (q - p.tq_cap < p.spill_cap) gsp_tq[q - p.tq_cap] = s | ((uint32_t)d << 24);
                    }
                }
            }
        }
        __syncwarp();   // the next mutation's hits find this one's red
    }
    __syncwarp();
    // settle the possibly multi-hit states: their final red is known now
    const int n_tq = wctr[1];
    unsigned st_tab = 0;
    auto settle = [&](uint32_t s, int red, bool twice) {
        const int b = (int)base_s[s] + DP_VOFF, v = b - red;
        if (v <= m0) {
            const int nodes = __ldg(snodes + s);
            if (b <= m0) atomicSub(&mv[b], nodes);
            atomicAdd(&mv[v], nodes);
            if (twice) {   // hit by two or more of the read's mutations: W does not reach it
                const int q = atomicAdd(&wctr[0], 1);
                if (q < p.cand_cap) cand[q] = s | ((uint32_t)v << 24);
                else if (q - p.cand_cap < p.spill_cap) gsp_cand[q - p.cand_cap] = s | ((uint32_t)v << 24);
            }
        }
    };
    if (n_tq > p.tq_cap + p.spill_cap) {   // not even the spill area holds the touched states: the slow path takes the read
        for (int i = lane; i < scr_words; i += 32) scr[i] = 0u;
        __syncwarp();
        if (lane < 2) wctr[lane] = 0;
        if (STATS && lane == 0) atomicAdd(p.stats + DP_ST_TQ_OVER, 1ull);
        __syncwarp();
        return false;
    }
    for (int i = lane; i < n_tq; i += 32) {
        const uint32_t t = i < p.tq_cap ? tq[i] : gsp_tq[i - p.tq_cap], s = t & 0xFFFFFFu;
        const int sh = (int)(s & 7u) * 4;
        const int red = (int)((atomicAnd(scr + (s >> 3), ~(15u << sh)) >> sh) & 15u);
        settle(s, red, red > (int)(t >> 24));
        if (STATS) ++st_tab;
    }
    acc0 = __reduce_add_sync(FULL, acc0);
    acc1 = __reduce_add_sync(FULL, acc1);
    acc2 = __reduce_add_sync(FULL, acc2);
    if (lane == 0) {
        atomicAdd(&mv[m0], acc0);
        atomicAdd(&mv[m0 - 1], acc1);
        atomicAdd(&mv[m0 - 2], acc2);
    }
    __syncwarp();
    const int n_cand = wctr[0];
    __syncwarp();
    if (lane < 2) wctr[lane] = 0;
    if (STATS) {
        st_hits = __reduce_add_sync(FULL, st_hits);
        st_multi = __reduce_add_sync(FULL, st_multi);
        st_clip = __reduce_add_sync(FULL, st_clip);
        st_track = __reduce_add_sync(FULL, st_track);
        st_tab = __reduce_add_sync(FULL, st_tab);
        if (lane == 0) {
            atomicAdd(p.stats + DP_ST_HITS, (unsigned long long)st_hits);
            atomicAdd(p.stats + DP_ST_MULTI, (unsigned long long)st_multi);
            atomicAdd(p.stats + DP_ST_CLIPPED, (unsigned long long)st_clip);
            atomicAdd(p.stats + DP_ST_TRACKED, (unsigned long long)st_track);
            atomicAdd(p.stats + DP_ST_TAB_ENTRIES, (unsigned long long)st_tab);
            atomicAdd(p.stats + DP_ST_CAND, (unsigned long long)n_cand);
        }
    }
    if (n_cand > p.cand_cap + p.spill_cap) {   // the scratch is zero again; the tracked bins go back to zero too
        mv[lane] = 0;
        mv[lane + 32] = 0;
        __syncwarp();
        return false;
    }
    const DpMin r = dp_minimum(p.max_pars, p.mult, m0, rec, lane, mv, whist_s, gw0, gw1, gc0, gc1);
    if (r.n_epp > 0) {
        const int o = m0 - r.mV;
        if (o <= 2 && lane < nm) {   // single-hit states end at most two bins below m0
            DeltaW* w = p.W + dg.w_off + ((int64_t)((int)m.w - dg.a_rel) * 5 + (int64_t)((m.y >> 28) - 1u)) * 3 + o;
            atomicAdd(&w->w, r.wgt);
            atomicAdd(&w->c, r.deg);
        }
        for (int i = lane; i < n_cand; i += 32) {
            const uint32_t cs = i < p.cand_cap ? cand[i] : gsp_cand[i - p.cand_cap];
            if ((int)(cs >> 24) == r.mV) {
                atomicAdd(p.saccS + so + (cs & 0xFFFFFFu), r.wgt);
                atomicAdd(p.saccC + so + (cs & 0xFFFFFFu), r.deg);
            }
        }
    }
    __syncwarp();
    return true;
}

template <bool STATS>
__global__ void __launch_bounds__(DP_WARPS * 32, DP_CTAS) delta_place_kernel(const DeltaPlaceParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    int* ctrl = reinterpret_cast<int*>(smem);                       // [0] unit, [1] next read of the unit
    int* whist_s = reinterpret_cast<int*>(smem + 16);
    unsigned char* base_s = smem + DP_FIXED;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    unsigned char* wa = smem + 16 + DP_BINS * 4 + (size_t)warp * DP_WARP_BYTES;
    int* mv = reinterpret_cast<int*>(wa);
    int* wctr = mv + DP_BINS;
    uint32_t* tq = reinterpret_cast<uint32_t*>(wa + DP_BINS * 4 + 16);
    uint32_t* cand = tq + DP_TQ;
    mv[lane] = 0;
    mv[lane + 32] = 0;
    if (lane < 4) wctr[lane] = 0;
    uint32_t* gscr = p.gscratch + ((size_t)blockIdx.x * DP_WARPS + warp) * (size_t)p.gscratch_words;
    uint32_t* gsp = p.gspill + ((size_t)blockIdx.x * DP_WARPS + warp) * 2 * (size_t)p.spill_cap;
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
        const int s_first = p.state_first[dg.list];
        const int s_n = p.state_first[dg.list + 1] - s_first;
        const int s_pad = (s_n + 15) & ~15;
        const int stride = dp_scratch_stride(s_n);
        const int aw = min(DP_WARPS, (p.smem_bytes - DP_FIXED - s_pad) / stride);   // warps whose scratch fits (>= 1: host)
        {   // the window's base scores and histogram
            const uint4* src = reinterpret_cast<const uint4*>(p.base + dg.base_off);
            uint4* dst = reinterpret_cast<uint4*>(base_s);
            for (int i = threadIdx.x; i < s_pad / 16; i += blockDim.x) dst[i] = __ldg(src + i);
            if (threadIdx.x < DP_BINS) whist_s[threadIdx.x] = p.whist[(size_t)du.group * DP_BINS + threadIdx.x];
        }
        uint32_t* scr = reinterpret_cast<uint32_t*>(base_s + s_pad + (size_t)warp * stride);
        if (warp < aw && dg.list != scr_list) {   // another list: the layout moved, its scratch area starts out zero
            uint4* z = reinterpret_cast<uint4*>(scr);
            for (int i = lane; i < stride / 16; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        scr_list = warp < aw ? dg.list : -1;
        __syncthreads();
        if (warp >= aw) continue;
        const int64_t so = p.sacc_off[dg.bucket];
        const int32_t* snodes = p.state_nodes + s_first;
        double gw0 = 0.0, gw1 = 0.0;
        int gc0 = 0, gc1 = 0;
        // the warps claim the unit's reads (sorted heaviest first) from a shared counter, two reads ahead: the record
        // of the read after next and the mutation records of the next read are in flight while a read is processed
        const uint4* rec = p.rec + du.first;
        auto claim = [&]() {
            int r = 0;
            if (lane == 0) r = atomicAdd(&ctrl[1], 1);
            return __shfl_sync(FULL, r, 0);
        };
        int i1 = claim(), i2 = claim();
        uint4 r1 = make_uint4(0u, 0u, 0u, 0u), r2 = r1;
        uint4 m1 = make_uint4(0u, 0u, 0u, 0u);
        if (i1 < du.count) r1 = __ldg(rec + i1);
        if (i2 < du.count) r2 = __ldg(rec + i2);
        if (lane < (int)(r1.w & 0xFFFFu)) m1 = __ldg(p.mrec + r1.z + lane);
        while (i1 < du.count) {
            const int i3 = claim();
            uint4 r3 = make_uint4(0u, 0u, 0u, 0u);
            if (i3 < du.count) r3 = __ldg(rec + i3);
            uint4 m2 = make_uint4(0u, 0u, 0u, 0u);
            if (lane < (int)(r2.w & 0xFFFFu)) m2 = __ldg(p.mrec + r2.z + lane);
            if ((int)(r1.w & 0xFFFFu) > DP_FAST_MUTS ||
                !dp_read<STATS>(p, dg, r1, m1, lane, base_s, snodes, scr, stride / 4, mv, wctr, tq, cand, gsp, whist_s, so, gw0, gw1, gc0, gc1)) {
                const DpSlowArgs sa = {p.post, p.mrec, p.max_pars, p.mult, p.saccS + so, p.saccC + so,
                                       p.Gw + (size_t)du.group * DP_BINS, p.Gc + (size_t)du.group * DP_BINS, dg.m0};
                dp_read_slow(sa, r1, lane, base_s, gscr, mv, whist_s);
                if (STATS && lane == 0) atomicAdd(p.stats + DP_ST_SLOW_READS, 1ull);
            }
            i1 = i2; r1 = r2; m1 = m2;
            i2 = i3; r2 = r3;
        }
        if (gc0 != 0) {
            atomicAdd(p.Gw + (size_t)du.group * DP_BINS + lane, gw0);
            atomicAdd(p.Gc + (size_t)du.group * DP_BINS + lane, gc0);
        }
        if (gc1 != 0) {
            atomicAdd(p.Gw + (size_t)du.group * DP_BINS + lane + 32, gw1);
            atomicAdd(p.Gc + (size_t)du.group * DP_BINS + lane + 32, gc1);
        }
    }
}

// per-(bucket, state) accumulators += over the bucket's groups (windows):
//   G[group][base_group(s)]                     the reads whose minimum is the state's base score and that do not touch it
//   W[group][q][c][m0 - (base_group(s) - d)]    for every position q the state mutates inside the window and every read
//                                               allele c with d = delta[ref] - delta[c] > 0: the reads with mutation (q, c)
//                                               whose minimum is where that single hit brings the state
__global__ void delta_finalize_kernel(const DeltaGroup* __restrict__ groups, const int32_t* __restrict__ bucket_goff,
                                      const int32_t* __restrict__ state_first, const BucketDesc* __restrict__ buckets,
                                      const uint8_t* __restrict__ base, const double* __restrict__ Gw,
                                      const int32_t* __restrict__ Gc, const DeltaW* __restrict__ W,
                                      const Entry* __restrict__ state_ent, const int64_t* __restrict__ state_eoff,
                                      const int64_t* __restrict__ sacc_off, double* __restrict__ saccS, int32_t* __restrict__ saccC) {
    const int b = blockIdx.y;
    const int g0 = bucket_goff[b], g1 = bucket_goff[b + 1];
    if (g0 == g1) return;
    const int l = buckets[b].list;
    const int s_lo = state_first[l], s_n = state_first[l + 1] - s_lo;
    const int64_t so = sacc_off[b];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < s_n; s += gridDim.x * blockDim.x) {
        const int64_t e0 = state_eoff[s_lo + s], e1 = state_eoff[s_lo + s + 1];
        double ws = 0.0;
        int cs = 0;
        for (int g = g0; g < g1; ++g) {
            const DeltaGroup dg = groups[g];
            const int v = (int)base[dg.base_off + s] + DP_VOFF;
            ws += Gw[(size_t)g * DP_BINS + v];
            cs += Gc[(size_t)g * DP_BINS + v];
            if (v - 2 > dg.m0) continue;   // no single hit brings the state to a bin a read's minimum can be in
            for (int64_t k = e0; k < e1; ++k) {
                const uint32_t z = __ldg(&state_ent[k].z), w = __ldg(&state_ent[k].w);
                const int q = (int)(w >> 16);
                if (q < dg.a_rel || q > dg.b_rel || (z == 0u && (w & 0xFFu) == 0u)) continue;
                const uint32_t xc = dp_table_class(z, w);
                const int dref = (int)((xc >> 3) & 1u), x = (int)(xc & 7u);
                const DeltaW* wq = W + dg.w_off + (int64_t)(q - dg.a_rel) * 15;
#pragma unroll
                for (int c = 1; c <= 5; ++c) {
                    const int d = dref + (x == c ? 1 : 0);
                    const int o = dg.m0 - (v - d);
                    if (d > 0 && o >= 0 && o <= 2) {
                        const DeltaW t = wq[(c - 1) * 3 + o];
                        ws += t.w;
                        cs += t.c;
                    }
                }
            }
        }
        if (cs != 0) {
            saccS[so + s] += ws;
            saccC[so + s] += cs;
        }
    }
}

}  // namespace wepp
