// kernels.cuh — sm_100a kernels of the placement path.
//
// Data layout in HBM (all built by wepp_set_arena / wepp_set_reads):
//   stripes      Entry[]      Euler entries of all events, grouped by genome stripe
//                              (pos / stripe_width), sort key ascending inside a stripe.  An event of
//                              an internal node is a BOUNDARY pair (ENTER at the node, EXIT at its
//                              subtree end); an event of a leaf is ONE POINT entry (host_prep.h)
//   lists        Entry[]      one Euler list per distinct read-window stripe range: the k-way
//                              merge (by key) of the stripes it covers, entry 0 = a dummy boundary
//                              at index 0; each evaluated entry carries its countable-node count
//   prev_boundary int32[]     per list entry: the boundary entry whose state encloses it
//   chunk_start  int32[][W+1] per list: the W scan chunks (each starts on a boundary entry)
//   reads        SoA          start/end/degree + sparse (pos, class) mutations in caller order, and the
//                              bucket-sorted permutation `perm` the tiles index
//   accS/accC    double/int32 per bucket, per list entry: sum of read weights / degrees whose
//                              EPP set contains the entry's nodes
//   diff_lo/hi   uint64[N+1]  128-bit fixed-point difference array for the per-node score
//   counts       int32[(N+1)*50]  difference array, scanned in place into the result
//
// Kernels (reference lines they replace in src/WEPP/initial_filter.cpp):
//   build_lists_kernel / finalize_lists_kernel   per-window Euler CSR (replaces the range trees,
//                                                 arena.cpp:68-169)
//   place_kernel<K,ACC,EPP> K1 signed delta (:59-87) + K2 Euler prefix sum (:101-104) + K3
//                          min / multiplicity / EPP emission (:89-99, :126-134) + K3'
//                          per-entry weight accumulation (:167-177); one CTA = one tile of
//                          32*K reads, its warps split the list, K reads per lane
//   expand_kernel + scan kernels   entry accumulators -> per-node score / counts (:199-211)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_prep.h"

namespace wepp {

// bucket-list form of Entry::x
constexpr uint32_t IDX_MASK = 0x0FFFFFFFu;   // preorder index (n_nodes < 2^28)
constexpr uint32_t ENT_EVAL = 0x80000000u;   // scores are evaluated here: a boundary entry that owns >= 1 node, or
                                             // the last point entry of a leaf
constexpr uint32_t ENT_POINT = 0x40000000u;  // point entry (leaf event)
constexpr uint32_t ENT_SKIP = 0x20000000u;   // point entry that is not the last of its leaf: applied to the running
                                             // prefix like a boundary entry and undone once the leaf is evaluated
constexpr uint32_t KEY_MASK = IDX_MASK | ENT_POINT;
constexpr int NBINS = 50;
constexpr int FIX_SHIFT = 80;  // score fixed point: value * 2^80 in a signed 128-bit integer

struct PlaceParams {
    const Entry* lists;
    const ListDesc* list_desc;
    const int32_t* chunk_start;   // [n_lists][PLACE_WARPS + 1]
    const BucketDesc* buckets;
    const TileDesc* tiles;
    int32_t n_tiles;
    int32_t n_nodes;
    int* tile_counter;
    // reads, in caller order; perm = bucket-sorted position -> caller's read index
    const int32_t* start;
    const int32_t* end;
    const int32_t* degree;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_code;   // allele class 1..4 = A,C,G,T ; 5 = N
    const int64_t* perm;
    // mask
    const uint8_t* mapped;  // may be null
    // outputs
    int32_t* max_pars;  // caller order
    int32_t* mult;
    double* accS;
    int32_t* accC;
    int accumulate;  // 0 for place_subset
    // EPP lists
    int32_t epp_cap;
    unsigned long long epp_capacity;
    unsigned long long* epp_total;
    int64_t* epp_off;  // caller order, -1 = not cached
    int32_t* epp_nodes;
    int32_t smem_per_warp;  // unused (layout is compile-time); kept for ABI stability of the struct
};

// Per-lane selector storage of the read-allele code table: one byte per read per window position,
// (class | (class|8) << 4): fed to PRMT as two selector nibbles it yields the read's signed delta
// byte followed by its sign replicated, i.e. a sign-extended 16-bit half.  Two reads = one
// 16-bit selector = one packed s16x2 delta.  SHIFT turns Entry::w into the byte offset of the
// position's row (row = 32 lanes * K bytes; bits 8..15 of Entry::w are zero by construction).
template <int K> struct Sel;
template <> struct Sel<8> { using type = uint2;    static constexpr int SHIFT = 8; };
template <> struct Sel<4> { using type = uint32_t; static constexpr int SHIFT = 9; };
template <> struct Sel<2> { using type = uint16_t; static constexpr int SHIFT = 10; };

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// col = 32-bit shared-window address of the lane's column, off = byte offset of the position's row
template <int K>
__device__ __forceinline__ void load_sel(uint32_t col, uint32_t off, uint32_t (&s)[K / 2]) {
    const uint32_t a = col + off;
    if constexpr (K == 8) {
        uint32_t x, y;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
        s[0] = x; s[1] = x >> 16; s[2] = y; s[3] = y >> 16;   // PRMT reads selector bits 0..15 only
    } else if constexpr (K == 4) {
        uint32_t x;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a));
        s[0] = x; s[1] = x >> 16;
    } else {
        uint32_t x;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(x) : "r"(a));
        s[0] = x;
    }
}

// Packed 16x2 arithmetic (sm_90+ hardware: VIADD.16x2, VIMNMX.S16x2 with two predicate outputs).
// Pass-1 running sums are held as value + S_BIAS in each half so that halves stay non-negative
// and the sum of the packed minima is a strictly monotone detector of "some minimum decreased".
constexpr uint32_t S_BIAS = 0x1000u;
constexpr uint32_t S_BIAS2 = 0x10001000u;
constexpr uint32_t BEST_NONE = 0x3FFFu;
constexpr uint32_t BEST_NONE2 = 0x3FFF3FFFu;
constexpr int MAX_WINDOW = 4000;   // widest bucket the 16-bit halves are sized for (|score| < S_BIAS)

// One pair of reads at one evaluated entry (initial_filter.cpp:89-99): b = min(s, b) per half,
// and where s <= b the entry's countable nodes are added to the read's node count (a strict
// decrease is repaired by the caller).  VIMNMX.S16x2 with two predicate outputs + two predicated adds.
__device__ __forceinline__ void min_count2(uint32_t s, uint32_t& b, int& c_lo, int& c_hi, int ucnt) {
    asm("{\n\t"
        ".reg .pred ph, pl;\n\t"
        ".reg .u16 a0, a1, m0, m1;\n\t"
        "min.s16x2 %0, %3, %0;\n\t"
        "mov.b32 {m0, m1}, %0;\n\t"
        "mov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.s16 pl, m0, a0;\n\t"
        "setp.eq.s16 ph, m1, a1;\n\t"
        "@pl add.s32 %1, %1, %4;\n\t"
        "@ph add.s32 %2, %2, %4;\n\t"
        "}"
        : "+r"(b), "+r"(c_lo), "+r"(c_hi)
        : "r"(s), "r"(ucnt));
}

struct __align__(16) PatEntry {   // pass-2 pattern table entry
    double w;
    int32_t c;
    int32_t pad;
};

// Shared-memory layout of place_kernel (one CTA = one tile of 32*K reads, PLACE_WARPS warps).
//   per warp : 32 staged entries (512 B) + pass-2 reduction staging (double[8][36] + int[8][36])
//   per CTA  : tile id, EPP write bases, pattern tables, the read-allele selector table
// The per-warp staging areas double as the exchange buffer for the chunk summaries between
// pass 1 and pass 2.  The pattern tables must lie below 64 KB (their addresses are carried
// in 16-bit halves).
constexpr int PLACE_WARPS = 8;
constexpr int RED_G = 8;              // entries reduced together in pass 2
constexpr int G1 = 8;                 // entries per pass-1 group (their shared loads are issued together)
constexpr int RED_S_STRIDE = 36;      // doubles per staged row: 32 lanes + pad (conflict-free column sums)
constexpr int RED_C_STRIDE = 36;      // ints per staged row (er * 36 + part distinct mod 32: conflict-free column sums)
constexpr int SMEM_EBUF = 0;                                        // 32 staged entries (512 B)
constexpr int SMEM_REDS = 512;                                      // double[RED_G][RED_S_STRIDE]
constexpr int SMEM_REDC = SMEM_REDS + RED_G * RED_S_STRIDE * 8;     // int[RED_G][RED_C_STRIDE]
constexpr int SMEM_WARP = (SMEM_REDC + RED_G * RED_C_STRIDE * 4 + 15) & ~15;   // bytes per warp
constexpr int SMEM_CTRL = PLACE_WARPS * SMEM_WARP;                  // int tile id (16 B)
constexpr int SMEM_WPB = SMEM_CTRL + 16;                            // u64[256] EPP write bases
constexpr int SMEM_TBL = SMEM_WPB + 256 * 8;                        // PatEntry[2][16][32]: [even/odd reads][pattern][lane]
constexpr int TBL_PAT_STRIDE = 32 * 16;                             // bytes per pattern row
constexpr int TBL_HALF = 16 * TBL_PAT_STRIDE;                       // bytes per half table
constexpr int SMEM_CODES = SMEM_TBL + 2 * TBL_HALF;                 // selector table
static_assert(SMEM_CODES <= 65536, "pattern tables must be addressable with 16-bit offsets");

__device__ __forceinline__ uint4 ld_entry(const Entry* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}

// ---------------------------------------------------------------------------------------------
// List construction: rank-by-binary-search k-way merge of the stripes a list covers (stable in
// (key, stripe) order).  grid = (chunks, n_lists), one thread per source entry.
__global__ void build_lists_kernel(const Entry* __restrict__ stripes, const int64_t* __restrict__ stripe_off,
                                   const ListDesc* __restrict__ list_desc, Entry* __restrict__ out, int q) {
    const ListDesc ld = list_desc[blockIdx.y];
    const int64_t src0 = stripe_off[ld.qs];
    const int n_src = ld.n - 1;
    Entry* dst = out + ld.off;
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[0] = Entry{0u, 0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
        const uint4 e = ld_entry(stripes + src0 + i);
        const int s = (int)(e.y / (uint32_t)q);
        int64_t rank = 1 + (src0 + i - stripe_off[s]);
        for (int t = ld.qs; t <= ld.qe; ++t) {
            if (t == s) continue;
            int64_t lo = stripe_off[t], hi = stripe_off[t + 1];
            const int64_t base = lo;
            // stripes before mine: count key <= mine; after mine: count key < mine
            const uint32_t key = t < s ? e.x + 1u : e.x;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(&stripes[mid].x) < key) lo = mid + 1; else hi = mid;
            }
            rank += lo - base;
        }
        Entry o;
        o.x = (e.x >> 1) | ((e.x & 1u) ? ENT_POINT : 0u);
        o.y = 0;
        o.z = e.z;
        o.w = (e.w & 0xFFu) | ((e.y - (uint32_t)ld.b0) << 16);
        dst[rank] = o;
    }
}

// Stripe-neighbourhood rank table (built once per tree, on the first wepp_set_reads that can use it):
// for every stripe entry e of stripe s and d = 1..D,
//   tab[(d - 1) * E + e]       = entries of stripes s-d .. s-1 with key <= key(e)
//   tab[(D + d - 1) * E + e]   = entries of stripes s+1 .. s+d with key <  key(e)
// so that an entry's rank in ANY list [qs, qe] with qe - qs <= D is its rank in its own stripe plus two
// table look-ups (build_lists_ranked_kernel) instead of one binary search per other stripe.
__global__ void rank_table_kernel(const Entry* __restrict__ stripes, const int64_t* __restrict__ stripe_off,
                                  int n_stripes, int q, int D, int64_t E, int32_t* __restrict__ tab) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 e = ld_entry(stripes + i);
        const int s = (int)(e.y / (uint32_t)q);
        int32_t acc = 0;
        for (int d = 1; d <= D; ++d) {
            const int t = s - d;
            if (t >= 0) {
                int64_t lo = stripe_off[t], hi = stripe_off[t + 1];
                const int64_t base = lo;
                const uint32_t key = e.x + 1u;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (__ldg(&stripes[mid].x) < key) lo = mid + 1; else hi = mid;
                }
                acc += (int32_t)(lo - base);
            }
            tab[(int64_t)(d - 1) * E + i] = acc;
        }
        acc = 0;
        for (int d = 1; d <= D; ++d) {
            const int t = s + d;
            if (t < n_stripes) {
                int64_t lo = stripe_off[t], hi = stripe_off[t + 1];
                const int64_t base = lo;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (__ldg(&stripes[mid].x) < e.x) lo = mid + 1; else hi = mid;
                }
                acc += (int32_t)(lo - base);
            }
            tab[(int64_t)(D + d - 1) * E + i] = acc;
        }
    }
}

// build_lists_kernel with the ranks read from the table (same output, bit for bit).
__global__ void build_lists_ranked_kernel(const Entry* __restrict__ stripes, const int64_t* __restrict__ stripe_off,
                                          const ListDesc* __restrict__ list_desc, Entry* __restrict__ out, int q,
                                          const int32_t* __restrict__ tab, int D, int64_t E) {
    const ListDesc ld = list_desc[blockIdx.y];
    const int64_t src0 = stripe_off[ld.qs];
    const int n_src = ld.n - 1;
    Entry* dst = out + ld.off;
    if (blockIdx.x == 0 && threadIdx.x == 0) dst[0] = Entry{0u, 0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
        const int64_t gi = src0 + i;
        const uint4 e = ld_entry(stripes + gi);
        const int s = (int)(e.y / (uint32_t)q);
        int64_t rank = 1 + (gi - stripe_off[s]);
        if (s > ld.qs) rank += __ldg(&tab[(int64_t)(s - ld.qs - 1) * E + gi]);
        if (ld.qe > s) rank += __ldg(&tab[(int64_t)(D + ld.qe - s - 1) * E + gi]);
        Entry o;
        o.x = (e.x >> 1) | ((e.x & 1u) ? ENT_POINT : 0u);
        o.y = 0;
        o.z = e.z;
        o.w = (e.w & 0xFFu) | ((e.y - (uint32_t)ld.b0) << 16);
        dst[rank] = o;
    }
}

// Entry kinds, countable-node counts, enclosing boundary entries and scan chunks.
// grid = (chunks, n_lists).  Threads only ever change the ENT_EVAL / ENT_SKIP bits of x, and
// neighbours read x through KEY_MASK, so the kernel can be re-run in place when `mapped` changes.
//   boundary entry i : owns the nodes of [idx_i, idx of the next boundary entry) that are not
//                      evaluated by point entries in between; y = how many of them are not mapped
//   point entries    : the last one of a leaf evaluates the leaf (y = 1 - mapped | earlier entries
//                      of the same leaf << 8); earlier ones are ENT_SKIP
constexpr int FIN_THREADS = 256;

__global__ void finalize_lists_kernel(Entry* __restrict__ lists, const ListDesc* __restrict__ list_desc, int n_nodes,
                                      const uint8_t* __restrict__ mapped, const int32_t* __restrict__ mapped_prefix,
                                      int32_t* __restrict__ prev_boundary, int32_t* __restrict__ chunk_start) {
    const ListDesc ld = list_desc[blockIdx.y];
    Entry* e = lists + ld.off;
    int32_t* pb = prev_boundary + ld.off;
    const int n = ld.n;
    if (blockIdx.x == 0 && threadIdx.x <= PLACE_WARPS) {
        // chunk k nominally starts at k * ceil(n / W); moved forward to a boundary entry so that a
        // boundary entry and the point entries it encloses are scanned by one warp
        const int k = threadIdx.x;
        const int cs = (n + PLACE_WARPS - 1) / PLACE_WARPS;
        int c = k == PLACE_WARPS ? n : min(n, k * cs);
        while (c < n && (e[c].x & ENT_POINT)) ++c;
        chunk_start[blockIdx.y * (PLACE_WARPS + 1) + k] = c;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t x = e[i].x & KEY_MASK;
        const uint32_t idx = x & IDX_MASK;
        if (!(x & ENT_POINT)) {
            int j = i + 1, pts = 0, mpts = 0;
            while (j < n) {
                const uint32_t xj = e[j].x & KEY_MASK;
                if (!(xj & ENT_POINT)) break;
                const bool last = (j + 1 == n) || ((e[j + 1].x & KEY_MASK) != xj);
                if (last) {
                    ++pts;
                    if (mapped) mpts += mapped[xj & IDX_MASK];
                }
                pb[j] = i;
                ++j;
            }
            const uint32_t nidx = j < n ? (e[j].x & IDX_MASK) : (uint32_t)n_nodes;
            const int own = (int)(nidx - idx) - pts;
            int ucnt = own;
            if (mapped_prefix) ucnt -= (mapped_prefix[nidx] - mapped_prefix[idx]) - mpts;
            int b = i - 1;
            while (b >= 0 && (e[b].x & ENT_POINT)) --b;
            pb[i] = b;
            e[i].y = (uint32_t)ucnt;
            e[i].x = x | (own > 0 ? ENT_EVAL : 0u);
        } else {
            const bool last = (i + 1 == n) || ((e[i + 1].x & KEY_MASK) != x);
            if (last) {
                int g = 0;
                while (i - g - 1 >= 0 && (e[i - g - 1].x & KEY_MASK) == x) ++g;
                e[i].y = ((mapped && mapped[idx]) ? 0u : 1u) | ((uint32_t)g << 8);
                e[i].x = x | ENT_EVAL;
            } else {
                e[i].y = 0u;
                e[i].x = x | ENT_SKIP;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Read keying on the device (wepp_set_reads): the reads are uploaded once, in caller order, and
// stay there; these two kernels validate them, translate the allele codes, count the reads of
// every (stripe range, count bin) cell and scatter the read indices into bucket order.  The
// host only turns the (small) cell histogram into list / bucket / tile descriptors in between.
//   cell = (qs * span_cap + (qe - qs)) * bins_per_stripe + (bin - first bin of stripe qs)
struct ReadKeyParams {
    int64_t n_reads, n_muts;
    int32_t genome, q, bin_size, span_cap, bins_per_stripe;
    const int32_t* start;
    const int32_t* end;
    const int32_t* degree;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_nuc;           // 1,2,4,8 = A,C,G,T ; 15 = N (mutation_annotated_tree.cpp:19-74)
    uint8_t* rm_code;                // out: 1..5
    int32_t* cell;                   // out: per read
    int32_t* table;                  // out: reads per cell (zeroed by the caller)
    unsigned long long* true_counts; // out: degree-weighted reads per count bin (arena.cpp:138-151)
    int* status;                     // out: [0] = max RP_ERR_* code seen, [1] = 1 if a window exceeds span_cap
};

__global__ void read_keys_kernel(const ReadKeyParams p) {
    __shared__ unsigned long long tc[NBINS];
    for (int j = threadIdx.x; j < NBINS; j += blockDim.x) tc[j] = 0ull;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < p.n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = p.start[r], e = p.end[r], d = p.degree[r];
        const int64_t a = p.rm_off[r], b = p.rm_off[r + 1];
        int err = RP_OK;
        if (s < 1 || s > p.genome || e > p.genome || e < s - 1) err = RP_ERR_WINDOW;
        else if (d < 0) err = RP_ERR_DEGREE;
        else if (a < 0 || b < a || b > p.n_muts) err = RP_ERR_OFFSETS;
        else {
            int32_t prev = s - 1;
            for (int64_t k = a; k < b; ++k) {
                const int32_t pos = p.rm_pos[k];
                const uint32_t c = p.rm_nuc[k];
                if (pos <= prev || pos > e) err = max(err, (int)RP_ERR_MUT_ORDER);
                prev = pos;
                if (!(c == 1u || c == 2u || c == 4u || c == 8u || c == 15u)) err = max(err, (int)RP_ERR_MUT_CODE);
                p.rm_code[k] = (uint8_t)(c == 15u ? 5u : (uint32_t)__ffs((int)c));
            }
        }
        if (err != RP_OK) {
            atomicMax(&p.status[0], err);
            p.cell[r] = -1;
            continue;
        }
        const int32_t qs = s / p.q, qe = max(e, s) / p.q;
        const int32_t bin = min(s / p.bin_size, NBINS - 1);
        const int32_t bin0 = min((qs * p.q) / p.bin_size, NBINS - 1);
        atomicAdd(&tc[bin], (unsigned long long)d);
        if (qe - qs >= p.span_cap || bin - bin0 >= p.bins_per_stripe) {   // not in the table: the host keys this set
            p.status[1] = 1;
            p.cell[r] = -1;
            continue;
        }
        const int32_t cell = (qs * p.span_cap + (qe - qs)) * p.bins_per_stripe + (bin - bin0);
        p.cell[r] = cell;
        atomicAdd(&p.table[cell], 1);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < NBINS; j += blockDim.x)
        if (tc[j]) atomicAdd(&p.true_counts[j], tc[j]);
}

// The occupied cells of the histogram as (cell << 32 | reads) pairs, so that the host never walks the table
// (a few hundred of its ~10^6 cells are occupied).  counter = status[2]; pairs beyond `cap` are only counted.
__global__ void cells_compact_kernel(const int32_t* __restrict__ table, int64_t n_cells, unsigned long long* __restrict__ pairs,
                                     int cap, int* __restrict__ status) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = table[c];
        if (v != 0) {
            const int i = atomicAdd(&status[2], 1);
            if (i < cap) pairs[i] = ((unsigned long long)c << 32) | (unsigned long long)(uint32_t)v;
        }
    }
}

// bucket_of_cell[cell] = bucket for the occupied cells ((cell << 32 | bucket) pairs from the host)
__global__ void cells_assign_kernel(const unsigned long long* __restrict__ pairs, int n, int32_t* __restrict__ bucket_of_cell) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bucket_of_cell[pairs[i] >> 32] = (int32_t)(uint32_t)pairs[i];
}

// cursor[b] starts at the bucket's first sorted position.  The order inside a bucket is the order
// the atomics land in: it only decides which reads share a tile, never a result.
// shared plan (wepp_set_allreduce): this rank's reads per bucket, from its own cell histogram
__global__ void cells_local_count_kernel(const unsigned long long* __restrict__ pairs, int n, const int32_t* __restrict__ local_table,
                                         int32_t* __restrict__ bucket_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = local_table[pairs[i] >> 32];
    if (c) atomicAdd(bucket_count + (uint32_t)pairs[i], c);
}

__global__ void read_scatter_kernel(int64_t n_reads, const int32_t* __restrict__ cell,
                                    const int32_t* __restrict__ bucket_of_cell, unsigned long long* __restrict__ cursor,
                                    int64_t* __restrict__ perm) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t b = bucket_of_cell[cell[r]];
        perm[atomicAdd(&cursor[b], 1ull)] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Out-of-line (rare) EPP list emission: nodes of [a, b) that are not mapped.
__device__ __noinline__ void emit_range(int32_t* __restrict__ out, unsigned long long& wp, uint32_t a, uint32_t b,
                                        const uint8_t* __restrict__ mapped) {
    for (uint32_t v = a; v < b; ++v)
        if (!mapped || !mapped[v]) out[wp++] = (int32_t)v;
}

// ---- shared-window load/store helpers (32-bit shared addresses; immediate offsets fold into LDS/STS) ----
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }

// Pass-1 state of one lane: K reads as K/2 packed pairs.
template <int K>
struct Pass1 {
    static constexpr int P = K / 2;
    uint32_t S[P];    // running prefix (biased), boundary entries only
    uint32_t B[P];    // running minima
    uint32_t oB[P];   // minima before the last strict decrease (they only change when one happens)
    uint32_t bsum;    // sum of B: changes iff some minimum strictly decreased
    int cnt[K];       // countable nodes at the running minimum

    // scores X are evaluated for ucnt countable nodes (initial_filter.cpp:89-99)
    __device__ __forceinline__ void eval(const uint32_t (&X)[P], int ucnt) {
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            min_count2(X[q], B[q], cnt[2 * q], cnt[2 * q + 1], ucnt);
            sum += B[q];
        }
        if (sum != bsum) {  // rare: some read reached a new strict minimum here
#pragma unroll
            for (int q = 0; q < P; ++q) {
                if ((X[q] & 0xFFFFu) < (oB[q] & 0xFFFFu)) cnt[2 * q] = ucnt;
                if ((X[q] >> 16) < (oB[q] >> 16)) cnt[2 * q + 1] = ucnt;
                oB[q] = B[q];
            }
            bsum = sum;
        }
    }
};

// The placement kernel.  Persistent CTAs; each CTA pulls tiles (32*K reads of one window bucket)
// from a global counter.  The CTA's PLACE_WARPS warps share the tile's selector table and each
// scans one chunk of the bucket's Euler list for ALL reads of the tile (lane = K reads held as
// K/2 packed s16x2 registers), a two-level Euler-tour scan:
//   pass 1  per chunk: sum of deltas, min prefix, node count at the min  -> shared memory
//           per entry and pair of reads: PRMT (signed delta pair) + VIADD.16x2 (prefix sum) +
//           VIMNMX.S16x2 with its two predicates (running min, "<= min") + two predicated adds
//           (node count); a strict decrease of any minimum is caught once per entry by comparing
//           the sum of the packed minima and repaired out of line.
//   combine every warp folds the chunk summaries: global min, multiplicity, its own start offset
//   pass 2  per chunk: entries attaining the min -> weight/degree sums into the entry
//           accumulators (staged through shared memory, one atomic per entry per tile) and
//           explicit EPP lists for reads under the cache cap.
// A boundary entry moves the running prefix; a point entry (leaf event) is evaluated as prefix +
// delta and leaves the prefix alone.  Entries are staged 32 at a time through shared memory (zero
// entries pad the chunk's tail: a zero entry is a boundary entry that changes nothing and is never
// evaluated).  ACC = accumulate per-entry weights (wepp_place); EPP = emit explicit lists.
template <int K, bool ACC, bool EPP>
__global__ void __launch_bounds__(PLACE_WARPS * 32, 2) place_kernel(const PlaceParams p) {
    using ST = typename Sel<K>::type;
    constexpr int P = K / 2;           // packed pairs per lane
    constexpr int SHIFT = Sel<K>::SHIFT;
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t ebuf_s = smem_s + warp * SMEM_WARP + SMEM_EBUF;
    const uint32_t redS_s = smem_s + warp * SMEM_WARP + SMEM_REDS;
    const uint32_t redC_s = smem_s + warp * SMEM_WARP + SMEM_REDC;
    int* xch = reinterpret_cast<int*>(smem);  // exchange [PLACE_WARPS][3][32*K] ints, aliases the staging areas
    int* ctrl = reinterpret_cast<int*>(smem + SMEM_CTRL);
    unsigned long long* wpb = reinterpret_cast<unsigned long long*>(smem + SMEM_WPB);
    PatEntry* tbl = reinterpret_cast<PatEntry*>(smem + SMEM_TBL);
    unsigned char* codes = smem + SMEM_CODES;
    const uint32_t col = smem_s + SMEM_CODES + lane * K;  // this lane's column of the selector table
    const unsigned FULL = 0xFFFFFFFFu;
    constexpr int T = 32 * K;
    static_assert(PLACE_WARPS * 3 * T * 4 <= PLACE_WARPS * SMEM_WARP, "exchange buffer must fit the staging areas");
    // shared addresses of this lane's pattern-table column: low half = even reads, high half = odd
    // reads (the kernel is never launched in a cluster, so the shared window starts near 0)
    if (smem_s + SMEM_CODES > 0xFFFFu) __trap();
    const uint32_t tbase2 = (smem_s + SMEM_TBL + lane * 16) | ((smem_s + SMEM_TBL + TBL_HALF + lane * 16) << 16);
    const uint32_t tzero = tbase2 + (uint32_t)(((1 << P) - 1) * TBL_PAT_STRIDE) * 0x00010001u;  // pattern "nobody at min" -> 0

    for (;;) {
        __syncthreads();  // previous tile fully done (selector table, exchange buffer)
        if (threadIdx.x == 0) ctrl[0] = atomicAdd(p.tile_counter, 1);
        __syncthreads();
        const int t = ctrl[0];
        if (t >= p.n_tiles) break;
        const TileDesc td = p.tiles[t];
        const BucketDesc bd = p.buckets[td.bucket];
        const ListDesc ld = p.list_desc[bd.list];
        const Entry* ent = p.lists + ld.off;
        const int n = ld.n;
        // this warp's chunk of the list (starts on a boundary entry)
        const int c0 = p.chunk_start[bd.list * (PLACE_WARPS + 1) + warp];
        const int c1 = p.chunk_start[bd.list * (PLACE_WARPS + 1) + warp + 1];

        // ---- read tile -> shared selector table -------------------------------------------------
        int run0[K];
        int64_t rid[K];
        {
            int s_rel[K], e_rel[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int ti = lane * K + j;
                const bool valid = ti < td.count;
                rid[j] = valid ? p.perm[td.first + ti] : -1;   // caller's read index
                s_rel[j] = valid ? p.start[rid[j]] - ld.b0 : 1;
                e_rel[j] = valid ? p.end[rid[j]] - ld.b0 : 0;
                run0[j] = 0;
            }
            for (int pos = warp; pos < ld.width; pos += PLACE_WARPS) {
                uint32_t w[2] = {0u, 0u};
#pragma unroll
                for (int j = 0; j < K; ++j)   // class 0 (as reference) inside the window, class 5 outside
                    w[j >> 2] |= ((pos >= s_rel[j] && pos <= e_rel[j]) ? 0x80u : 0xD5u) << (8 * (j & 3));
                ST* row = reinterpret_cast<ST*>(codes) + pos * 32 + lane;
                if constexpr (K == 8) *row = make_uint2(w[0], w[1]);
                else *row = (ST)w[0];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (rid[j] >= 0) {
                const int64_t a = p.rm_off[rid[j]], b = p.rm_off[rid[j] + 1];
                for (int64_t k = a; k < b; ++k) {
                    const uint32_t c = p.rm_code[k];
                    if (warp == 0) {  // one owner per table column: no races
                        const int pr = p.rm_pos[k] - ld.b0;
                        codes[(pr * 32 + lane) * K + j] = (unsigned char)(c | ((c | 8u) << 4));
                    }
                    run0[j] += (c <= 4u);  // seed set: non-N mutations (initial_filter.cpp:118-123)
                }
            }
        }
        __syncthreads();

        // ---- pass 1: chunk-relative prefix sum of signed deltas, min prefix and its node count ----
        int run[K], best[K], cnt[K];
        {
            Pass1<K> st;
            st.bsum = (uint32_t)P * BEST_NONE2;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                st.S[q] = S_BIAS2;
                st.B[q] = st.oB[q] = BEST_NONE2;
            }
#pragma unroll
            for (int j = 0; j < K; ++j) st.cnt[j] = 0;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (c0 + lane < c1) nxt = ld_entry(ent + c0 + lane);
            for (int base = c0; base < c1; base += 32) {
                sts128(ebuf_s + lane * 16, nxt);
                const bool is_p = (nxt.x & (ENT_EVAL | ENT_POINT)) == (ENT_EVAL | ENT_POINT);
                const uint32_t em = __ballot_sync(FULL, (nxt.x & ENT_EVAL) != 0u);   // evaluated entries of this batch
                const uint32_t pm = __ballot_sync(FULL, is_p);                       // ... of which leaf evaluations
                const uint32_t mm = __ballot_sync(FULL, is_p && (nxt.y >> 8) != 0u); // ... of multi-event leaves
                __syncwarp();
                nxt = make_uint4(0, 0, 0, 0);
                if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
                for (int g = 0; g < 32 / G1; ++g) {
                    const uint32_t ea = ebuf_s + g * (G1 * 16);
                    const uint32_t eg = em >> (G1 * g), pg = pm >> (G1 * g);
                    if (((mm >> (G1 * g)) & ((1u << G1) - 1u)) == 0u) {
                        // all shared-memory loads of the group are issued before the first branch
                        uint4 e[G1];
                        uint32_t sel[G1][P];
#pragma unroll
                        for (int i = 0; i < G1; ++i) e[i] = lds128(ea + i * 16);
#pragma unroll
                        for (int i = 0; i < G1; ++i) load_sel<K>(col, e[i].w >> SHIFT, sel[i]);
#pragma unroll
                        for (int i = 0; i < G1; ++i) {
                            if (pg & (1u << i)) {          // warp-uniform: a leaf is evaluated, the prefix stays
                                uint32_t X[P];
#pragma unroll
                                for (int q = 0; q < P; ++q) X[q] = __vadd2(st.S[q], prmt(e[i].z, e[i].w, sel[i][q]));
                                st.eval(X, (int)(e[i].y & 0xFFu));
                            } else {
#pragma unroll
                                for (int q = 0; q < P; ++q) st.S[q] = __vadd2(st.S[q], prmt(e[i].z, e[i].w, sel[i][q]));
                                if (eg & (1u << i)) st.eval(st.S, (int)e[i].y);
                            }
                        }
                    } else {
                        // rare: the group holds a leaf with several events in this window.  Its earlier
                        // point entries were applied to the prefix; after the evaluation they are undone.
#pragma unroll 1
                        for (int i = 0; i < G1; ++i) {
                            const uint4 e = lds128(ea + i * 16);
                            uint32_t sel[P];
                            load_sel<K>(col, e.w >> SHIFT, sel);
                            if (pg & (1u << i)) {
                                uint32_t X[P];
#pragma unroll
                                for (int q = 0; q < P; ++q) X[q] = __vadd2(st.S[q], prmt(e.z, e.w, sel[q]));
                                st.eval(X, (int)(e.y & 0xFFu));
                                const int64_t at = (int64_t)base + g * G1 + i;
                                for (int k = 1; k <= (int)(e.y >> 8); ++k) {
                                    const uint4 u = ld_entry(ent + at - k);
                                    uint32_t us[P];
                                    load_sel<K>(col, u.w >> SHIFT, us);
#pragma unroll
                                    for (int q = 0; q < P; ++q) st.S[q] = __vsub2(st.S[q], prmt(u.z, u.w, us[q]));
                                }
                            } else {
#pragma unroll
                                for (int q = 0; q < P; ++q) st.S[q] = __vadd2(st.S[q], prmt(e.z, e.w, sel[q]));
                                if (eg & (1u << i)) st.eval(st.S, (int)e.y);
                            }
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < P; ++q) {
                run[2 * q] = (int)(st.S[q] & 0xFFFFu) - (int)S_BIAS;
                run[2 * q + 1] = (int)(st.S[q] >> 16) - (int)S_BIAS;
                const uint32_t bl = st.B[q] & 0xFFFFu, bh = st.B[q] >> 16;
                best[2 * q] = bl == BEST_NONE ? 0x3FFFFFFF : (int)bl - (int)S_BIAS;
                best[2 * q + 1] = bh == BEST_NONE ? 0x3FFFFFFF : (int)bh - (int)S_BIAS;
            }
#pragma unroll
            for (int j = 0; j < K; ++j) cnt[j] = st.cnt[j];
        }
        // ---- exchange chunk summaries, fold them ------------------------------------------------
        __syncthreads();  // all warps are done with their staging areas
#pragma unroll
        for (int j = 0; j < K; ++j) {
            xch[(warp * 3 + 0) * T + j * 32 + lane] = run[j];
            xch[(warp * 3 + 1) * T + j * 32 + lane] = best[j];
            xch[(warp * 3 + 2) * T + j * 32 + lane] = cnt[j];
        }
        __syncthreads();
        int epp_before[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            int off = run0[j], my_off = 0, gb = 0x3FFFFFFF;
#pragma unroll
            for (int w = 0; w < PLACE_WARPS; ++w) {
                if (w == warp) my_off = off;
                const int b = xch[(w * 3 + 1) * T + j * 32 + lane];
                if (b != 0x3FFFFFFF) gb = min(gb, off + b);
                off += xch[(w * 3 + 0) * T + j * 32 + lane];
            }
            int gc = 0, before = 0;
            off = run0[j];
#pragma unroll
            for (int w = 0; w < PLACE_WARPS; ++w) {
                const int b = xch[(w * 3 + 1) * T + j * 32 + lane];
                const int c = xch[(w * 3 + 2) * T + j * 32 + lane];
                if (b != 0x3FFFFFFF && off + b == gb) {
                    gc += c;
                    if (w < warp) before += c;
                }
                off += xch[(w * 3 + 0) * T + j * 32 + lane];
            }
            run[j] = my_off;      // pass 2 starts from this warp's absolute offset
            best[j] = gb;
            cnt[j] = gc;
            epp_before[j] = before;
        }
        __syncthreads();  // exchange buffer consumed; staging areas are free again

        // ---- per-read results (warp 0 writes; EPP space is allocated by warp 0) ------------------
        double wgt[K];
        int deg[K];
        unsigned long long wp[K];
        uint32_t small_mask = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            wgt[j] = 0.0;
            deg[j] = 0;
            wp[j] = 0;
            if (rid[j] >= 0) {
                const int d = p.degree[rid[j]];
                if (cnt[j] > 0) {
                    // node_score, initial_filter.hpp:54-57
                    wgt[j] = (double)d / ((double)(1 + best[j]) * (double)cnt[j]);
                    deg[j] = d;
                }
                if (warp == 0) {
                    const int64_t orig = rid[j];
                    p.max_pars[orig] = best[j];
                    p.mult[orig] = cnt[j];
                    if (EPP && p.epp_off) {
                        long long off = -1;
                        unsigned long long base = ~0ull;
                        if (cnt[j] > 0 && cnt[j] <= p.epp_cap) {
                            const unsigned long long o = atomicAdd(p.epp_total, (unsigned long long)cnt[j]);
                            if (o + (unsigned long long)cnt[j] <= p.epp_capacity) {
                                off = (long long)o;
                                base = o;
                            }
                        } else if (cnt[j] == 0) {
                            off = 0;  // empty but known
                        }
                        p.epp_off[orig] = off;
                        wpb[j * 32 + lane] = base;
                    }
                }
            } else if (EPP && warp == 0 && p.epp_off) {
                wpb[j * 32 + lane] = ~0ull;
            }
        }
        if (EPP && p.epp_off) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const unsigned long long base = wpb[j * 32 + lane];
                if (base != ~0ull) {
                    wp[j] = base + (unsigned long long)epp_before[j];
                    small_mask |= 1u << j;
                }
            }
        }
        if (!ACC) {
            if (!__syncthreads_or(small_mask != 0)) continue;   // nothing to emit: no pass 2
        }

        // ---- pattern tables: for the lane's even reads (low halves) and odd reads (high halves),
        //      the sum of weights / degrees of the reads whose NOT-at-min bit is clear, indexed by
        //      the P-bit pattern (bit q = read of pair q is above its min).  All warps hold the same
        //      per-read results, so they split the 2 x 2^P patterns. --------------------------------
        if (ACC) {
            for (int hp = warp; hp < 2 * (1 << P); hp += PLACE_WARPS) {
                const int h = hp >> P, pat = hp & ((1 << P) - 1);
                double ws = 0.0;
                int ds = 0;
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    if (!((pat >> q) & 1)) {
                        ws += h ? wgt[2 * q + 1] : wgt[2 * q];
                        ds += h ? deg[2 * q + 1] : deg[2 * q];
                    }
                }
                PatEntry pe;
                pe.w = ws;
                pe.c = ds;
                pe.pad = 0;
                tbl[(h * 16 + pat) * 32 + lane] = pe;
            }
        }
        __syncthreads();

        // ---- pass 2: which entries attain the min -> weights into the entry accumulators,
        //      EPP node lists for reads under the cache cap.  rel = running score - min >= 0 at
        //      every evaluated entry, so min(rel, 1) is the read's NOT-at-min bit. ------------------
        double* accS = p.accS + bd.acc_off;
        int32_t* accC = p.accC + bd.acc_off;
        uint32_t rel[P];
        uint32_t small_lo = 0, small_hi = 0;
        uint32_t enc_lo = 0, enc_hi = 0;   // EPP: reads at their min in the enclosing boundary entry's state
#pragma unroll
        for (int q = 0; q < P; ++q) {
            rel[q] = ((uint32_t)(run[2 * q] - best[2 * q]) & 0xFFFFu) | ((uint32_t)(run[2 * q + 1] - best[2 * q + 1]) << 16);
            small_lo |= ((small_mask >> (2 * q)) & 1u) << q;
            small_hi |= ((small_mask >> (2 * q + 1)) & 1u) << q;
        }
        {
            const int er = lane >> 2, part = lane & 3;   // reduction role: entry er of the group, quarter `part` of the lanes
            const uint32_t rS = redS_s + (er * RED_S_STRIDE + part) * 8;
            const uint32_t rC = redC_s + (er * RED_C_STRIDE + part) * 4;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (c0 + lane < c1) nxt = ld_entry(ent + c0 + lane);
            for (int base = c0; base < c1; base += 32) {
                sts128(ebuf_s + lane * 16, nxt);
                const bool is_p = (nxt.x & (ENT_EVAL | ENT_POINT)) == (ENT_EVAL | ENT_POINT);
                const uint32_t em = __ballot_sync(FULL, (nxt.x & ENT_EVAL) != 0u);
                const uint32_t pm = __ballot_sync(FULL, is_p);
                const uint32_t mm = __ballot_sync(FULL, is_p && (nxt.y >> 8) != 0u);
                const uint32_t qm = EPP ? __ballot_sync(FULL, (nxt.x & ENT_POINT) != 0u) : 0u;   // any point entry
                __syncwarp();
                nxt = make_uint4(0, 0, 0, 0);
                if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    const uint32_t ea = ebuf_s + g * 128;
                    const uint32_t eg = em >> (8 * g), pg = pm >> (8 * g);
                    uint32_t tt[RED_G];
                    if (((mm >> (8 * g)) & 0xFFu) == 0u) {
#pragma unroll
                        for (int i = 0; i < RED_G; ++i) {
                            const uint2 e = lds64(ea + i * 16 + 8);   // delta bytes + position
                            uint32_t sel[P];
                            load_sel<K>(col, e.y >> SHIFT, sel);
                            uint32_t ti = tbase2;
                            const bool pt = (pg & (1u << i)) != 0u;
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                const uint32_t x = __vadd2(rel[q], prmt(e.x, e.y, sel[q]));
                                ti += __vmins2(x, 0x00010001u) * (uint32_t)(TBL_PAT_STRIDE << q);
                                rel[q] = pt ? rel[q] : x;   // a leaf evaluation leaves the prefix alone
                            }
                            tt[i] = (eg & (1u << i)) ? ti : tzero;
                        }
                    } else {
                        // rare: a leaf with several events in this window (see pass 1)
#pragma unroll
                        for (int i = 0; i < RED_G; ++i) {
                            const uint4 e = lds128(ea + i * 16);
                            uint32_t sel[P];
                            load_sel<K>(col, e.w >> SHIFT, sel);
                            uint32_t ti = tbase2;
                            const bool pt = (pg & (1u << i)) != 0u;
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                const uint32_t x = __vadd2(rel[q], prmt(e.z, e.w, sel[q]));
                                ti += __vmins2(x, 0x00010001u) * (uint32_t)(TBL_PAT_STRIDE << q);
                                rel[q] = pt ? rel[q] : x;
                            }
                            tt[i] = (eg & (1u << i)) ? ti : tzero;
                            if (pt && (e.y >> 8) != 0u) {
                                const int64_t at = (int64_t)base + g * 8 + i;
                                for (int k = 1; k <= (int)(e.y >> 8); ++k) {
                                    const uint4 u = ld_entry(ent + at - k);
                                    uint32_t us[P];
                                    load_sel<K>(col, u.w >> SHIFT, us);
#pragma unroll
                                    for (int q = 0; q < P; ++q) rel[q] = __vsub2(rel[q], prmt(u.z, u.w, us[q]));
                                }
                            }
                        }
                    }
                    if (EPP && small_mask) {   // explicit EPP lists (sorted: the list is in preorder)
#pragma unroll 1
                        for (int i = 0; i < RED_G; ++i) {
                            uint32_t tti = tt[0];
#pragma unroll
                            for (int k = 1; k < RED_G; ++k) tti = (i == k) ? tt[k] : tti;
                            const uint32_t dd = tti - tbase2;   // per half: pattern * TBL_PAT_STRIDE
                            const uint32_t hit_lo = ~((dd & 0xFFFFu) / TBL_PAT_STRIDE) & small_lo;
                            const uint32_t hit_hi = ~((dd >> 16) / TBL_PAT_STRIDE) & small_hi;
                            const bool any_point = ((qm >> (8 * g + i)) & 1u) != 0u;
                            if (!any_point) {   // boundary entry: its state encloses the following point entries
                                enc_lo = hit_lo;
                                enc_hi = hit_hi;
                            }
                            if (hit_lo | hit_hi | (any_point ? (enc_lo | enc_hi) : 0u)) {
                                const int64_t at = (int64_t)base + g * 8 + i;
                                const uint32_t v = __ldg(&ent[at].x) & IDX_MASK;
                                const uint32_t vn = at + 1 < n ? (__ldg(&ent[at + 1].x) & IDX_MASK) : (uint32_t)p.n_nodes;
#pragma unroll
                                for (int q = 0; q < P; ++q) {
                                    if (!any_point) {
                                        if (hit_lo & (1u << q)) emit_range(p.epp_nodes, wp[2 * q], v, vn, p.mapped);
                                        if (hit_hi & (1u << q)) emit_range(p.epp_nodes, wp[2 * q + 1], v, vn, p.mapped);
                                    } else {   // the leaf itself, then the enclosing state's nodes up to the next entry
                                        if (hit_lo & (1u << q)) emit_range(p.epp_nodes, wp[2 * q], v, v + 1, p.mapped);
                                        if (enc_lo & (1u << q)) emit_range(p.epp_nodes, wp[2 * q], v + 1, vn, p.mapped);
                                        if (hit_hi & (1u << q)) emit_range(p.epp_nodes, wp[2 * q + 1], v, v + 1, p.mapped);
                                        if (enc_hi & (1u << q)) emit_range(p.epp_nodes, wp[2 * q + 1], v + 1, vn, p.mapped);
                                    }
                                }
                            }
                        }
                    }
                    if (ACC) {
                        // this lane's partial sums for the 8 entries -> staging, 4 entries at a time
#pragma unroll
                        for (int h4 = 0; h4 < RED_G; h4 += 4) {
                            // one 16-byte load per lookup: a quarter-warp covers all 32 banks (the 8+4-byte
                            // pair of loads on this 16-byte stride would be 4-way conflicted)
                            uint4 pa[4], pb[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                pa[i] = lds128(tt[h4 + i] & 0xFFFFu);
                                pb[i] = lds128(tt[h4 + i] >> 16);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const double wa = __hiloint2double((int)pa[i].y, (int)pa[i].x);
                                const double wb = __hiloint2double((int)pb[i].y, (int)pb[i].x);
                                sts_f64(redS_s + ((h4 + i) * RED_S_STRIDE + lane) * 8, wa + wb);
                                sts32(redC_s + ((h4 + i) * RED_C_STRIDE + lane) * 4, pa[i].z + pb[i].z);
                            }
                        }
                        // column sums: lane (er, part) adds 8 of the 32 staged values of entry er
                        __syncwarp();
                        double sv[8];
                        uint32_t cv[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            sv[k] = lds_f64(rS + 4 * k * 8);
                            cv[k] = lds32(rC + 4 * k * 4);
                        }
                        // pairwise tree: three dependent adds instead of eight
                        double s = ((sv[0] + sv[1]) + (sv[2] + sv[3])) + ((sv[4] + sv[5]) + (sv[6] + sv[7]));
                        uint32_t c = ((cv[0] + cv[1]) + (cv[2] + cv[3])) + ((cv[4] + cv[5]) + (cv[6] + cv[7]));
                        s += __shfl_xor_sync(FULL, s, 1);
                        c += __shfl_xor_sync(FULL, c, 1);
                        s += __shfl_xor_sync(FULL, s, 2);
                        c += __shfl_xor_sync(FULL, c, 2);
                        if (part == 0) {   // padded entries stage zeros: nothing is added out of range
                            if (s != 0.0) atomicAdd(accS + base + g * RED_G + er, s);
                            if (c != 0) atomicAdd(accC + base + g * RED_G + er, (int)c);
                        }
                        __syncwarp();
                    }
                }
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Entry accumulators -> per-node difference arrays.  grid = (chunks, n_buckets).
//   boundary entry: its nodes' value differs from the previous boundary entry's by (cur - prv)
//                   from its index on (an entry that owns no node has value 0);
//   point entry   : the leaf's value differs from the enclosing boundary entry's at exactly one node.
__device__ __forceinline__ void dbl_to_fix(double d, unsigned long long& lo, long long& hi) {
    lo = 0;
    hi = 0;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    const int ex = (int)((bits >> 52) & 0x7FF);
    if (ex == 0) return;  // zero / denormal
    const unsigned long long man = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    const int sh = ex - 1075 + FIX_SHIFT;
    if (sh >= 64) {
        hi = (long long)(man << (sh - 64));
    } else if (sh > 0) {
        lo = man << sh;
        hi = (long long)(man >> (64 - sh));
    } else if (sh == 0) {
        lo = man;
    } else if (sh > -53) {
        lo = man >> (-sh);
    }
}

__device__ __forceinline__ void atomic_add128(unsigned long long* dlo, unsigned long long* dhi, unsigned long long lo,
                                              long long hi) {
    const unsigned long long old = atomicAdd(dlo, lo);
    const unsigned long long carry = (old + lo < old) ? 1ull : 0ull;
    const unsigned long long h = (unsigned long long)hi + carry;
    if (h) atomicAdd(dhi, h);
}

__global__ void expand_kernel(const Entry* __restrict__ lists, const ListDesc* __restrict__ list_desc,
                              const BucketDesc* __restrict__ buckets, const int32_t* __restrict__ prev_boundary,
                              const double* __restrict__ accS, const int32_t* __restrict__ accC,
                              unsigned long long* __restrict__ diff_lo, unsigned long long* __restrict__ diff_hi,
                              int32_t* __restrict__ counts) {
    const BucketDesc bd = buckets[blockIdx.y];
    const ListDesc ld = list_desc[bd.list];
    const Entry* e = lists + ld.off;
    const int32_t* pb = prev_boundary + ld.off;
    const double* s = accS + bd.acc_off;
    const int32_t* c = accC + bd.acc_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const uint32_t x = __ldg(&e[i].x);
        if (x & ENT_SKIP) continue;
        const int pi = pb[i];
        const double cur = s[i], prv = pi >= 0 ? s[pi] : 0.0;
        const int32_t ccur = c[i], cprv = pi >= 0 ? c[pi] : 0;
        if (cur == prv && ccur == cprv) continue;
        const uint32_t idx = x & IDX_MASK;
        const bool point = (x & ENT_POINT) != 0u;
        if (cur != prv) {
            unsigned long long alo, blo;
            long long ahi, bhi;
            dbl_to_fix(cur, alo, ahi);
            dbl_to_fix(prv, blo, bhi);
            const unsigned long long lo = alo - blo;
            const long long hi = ahi - bhi - (alo < blo ? 1 : 0);
            atomic_add128(diff_lo + idx, diff_hi + idx, lo, hi);
            if (point) {   // back to the enclosing value right after the leaf
                const unsigned long long nlo = 0ull - lo;
                const long long nhi = ~hi + (lo == 0ull ? 1 : 0);
                atomic_add128(diff_lo + idx + 1, diff_hi + idx + 1, nlo, nhi);
            }
        }
        if (counts && ccur != cprv) {
            atomicAdd(counts + (size_t)idx * NBINS + bd.bin, ccur - cprv);
            if (point) atomicAdd(counts + (size_t)(idx + 1) * NBINS + bd.bin, cprv - ccur);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 128-bit inclusive prefix sum over nodes -> double score.  Three phases, CHUNK nodes per block.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

struct U128 {
    unsigned long long lo, hi;
};
__device__ __forceinline__ U128 add128(U128 a, U128 b) {
    U128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
__device__ __forceinline__ U128 shfl_up128(U128 v, int d) {
    U128 r;
    r.lo = __shfl_up_sync(0xFFFFFFFFu, v.lo, d);
    r.hi = __shfl_up_sync(0xFFFFFFFFu, v.hi, d);
    return r;
}

// block-wide exclusive scan of per-thread totals; returns this thread's exclusive prefix and the
// block total in `total`.
__device__ __forceinline__ U128 block_exclusive128(U128 v, U128& total) {
    __shared__ U128 warp_tot[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    U128 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        U128 o = shfl_up128(inc, d);
        if (lane >= d) inc = add128(inc, o);
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    U128 off = {0, 0};
    U128 tot = {0, 0};
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) off = add128(off, warp_tot[w]);
        tot = add128(tot, warp_tot[w]);
    }
    total = tot;
    U128 exc;  // inclusive - own
    exc.lo = inc.lo - v.lo;
    exc.hi = inc.hi - v.hi - (inc.lo < v.lo ? 1ull : 0ull);
    __syncthreads();
    return add128(off, exc);
}

__global__ void __launch_bounds__(SCAN_THREADS) score_chunk_sum_kernel(const unsigned long long* __restrict__ lo,
                                                                        const unsigned long long* __restrict__ hi, int n,
                                                                        U128* __restrict__ chunk_tot) {
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    U128 acc = {0, 0};
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) acc = add128(acc, U128{lo[base + k], hi[base + k]});
    U128 total;
    block_exclusive128(acc, total);
    if (threadIdx.x == 0) chunk_tot[blockIdx.x] = total;
}

__global__ void score_chunk_scan_kernel(U128* __restrict__ chunk_tot, int n_chunks) {
    // single block; exclusive scan in place
    __shared__ U128 carry_s;
    if (threadIdx.x == 0) carry_s = U128{0, 0};
    __syncthreads();
    for (int base = 0; base < n_chunks; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        U128 v = i < n_chunks ? chunk_tot[i] : U128{0, 0};
        U128 total;
        U128 exc = block_exclusive128(v, total);
        const U128 carry = carry_s;
        if (i < n_chunks) chunk_tot[i] = add128(carry, exc);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = add128(carry, total);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) score_apply_kernel(const unsigned long long* __restrict__ lo,
                                                                    const unsigned long long* __restrict__ hi, int n,
                                                                    const U128* __restrict__ chunk_off,
                                                                    const uint8_t* __restrict__ mapped,
                                                                    double* __restrict__ score) {
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    U128 v[SCAN_ITEMS];
    U128 acc = {0, 0};
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? U128{lo[base + k], hi[base + k]} : U128{0, 0};
        acc = add128(acc, v[k]);
    }
    U128 total;
    U128 run = add128(block_exclusive128(acc, total), chunk_off[blockIdx.x]);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        run = add128(run, v[k]);
        if (base + k < n) {
            // fixed point -> double (value is non-negative up to rounding of the inputs)
            const double d = ((double)(long long)run.hi * 18446744073709551616.0 + (double)run.lo) *
                             8.271806125530277e-25;  // 2^-80
            score[base + k] = (mapped && mapped[base + k]) ? 0.0 : d;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// counts: in-place inclusive prefix along nodes of int32[(N+1)][50] (HBM streaming: the matrix is
// read twice and written once).  A block owns CNT_CHUNK consecutive nodes; thread b < 50 owns bin b
// (a warp reads 200 contiguous bytes per node) and keeps CNT_BATCH independent loads in flight.
constexpr int CNT_CHUNK = 512;   // nodes per block
constexpr int CNT_BATCH = 16;    // rows loaded together

__global__ void __launch_bounds__(64) counts_chunk_sum_kernel(const int32_t* __restrict__ counts, int n,
                                                               int32_t* __restrict__ chunk_tot) {
    const int b = threadIdx.x;
    if (b >= NBINS) return;
    const int v0 = blockIdx.x * CNT_CHUNK, v1 = min(n, v0 + CNT_CHUNK);
    int32_t acc = 0;
    for (int v = v0; v < v1; v += CNT_BATCH) {
        int32_t x[CNT_BATCH];
#pragma unroll
        for (int k = 0; k < CNT_BATCH; ++k) x[k] = (v + k < v1) ? __ldg(counts + (size_t)(v + k) * NBINS + b) : 0;
#pragma unroll
        for (int k = 0; k < CNT_BATCH; ++k) acc += x[k];
    }
    chunk_tot[(size_t)blockIdx.x * NBINS + b] = acc;
}

// exclusive scan over chunks: one block per bin, each thread sums a contiguous run of chunks,
// the run totals are scanned through shared memory
constexpr int CSCAN_THREADS = 256;
__global__ void __launch_bounds__(CSCAN_THREADS) counts_chunk_scan_kernel(const int32_t* __restrict__ chunk_tot,
                                                                           int32_t* __restrict__ chunk_off, int n_chunks) {
    __shared__ int32_t run_tot[CSCAN_THREADS];
    const int b = blockIdx.x;
    const int per = (n_chunks + CSCAN_THREADS - 1) / CSCAN_THREADS;
    const int c0 = min(n_chunks, (int)threadIdx.x * per), c1 = min(n_chunks, c0 + per);
    int32_t acc = 0;
    for (int c = c0; c < c1; ++c) acc += chunk_tot[(size_t)c * NBINS + b];
    run_tot[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t run = 0;
        for (int t = 0; t < CSCAN_THREADS; ++t) {
            const int32_t x = run_tot[t];
            run_tot[t] = run;
            run += x;
        }
    }
    __syncthreads();
    acc = run_tot[threadIdx.x];
    for (int c = c0; c < c1; ++c) {
        chunk_off[(size_t)c * NBINS + b] = acc;
        acc += chunk_tot[(size_t)c * NBINS + b];
    }
}

__global__ void __launch_bounds__(64) counts_apply_kernel(int32_t* __restrict__ counts, int n,
                                                           const int32_t* __restrict__ chunk_off,
                                                           const uint8_t* __restrict__ mapped) {
    const int b = threadIdx.x;
    if (b >= NBINS) return;
    const int v0 = blockIdx.x * CNT_CHUNK, v1 = min(n, v0 + CNT_CHUNK);
    int32_t acc = chunk_off[(size_t)blockIdx.x * NBINS + b];
    for (int v = v0; v < v1; v += CNT_BATCH) {
        int32_t x[CNT_BATCH];
        uint8_t m[CNT_BATCH];
#pragma unroll
        for (int k = 0; k < CNT_BATCH; ++k) {
            x[k] = (v + k < v1) ? counts[(size_t)(v + k) * NBINS + b] : 0;
            m[k] = (mapped && v + k < v1) ? mapped[v + k] : 0;
        }
#pragma unroll
        for (int k = 0; k < CNT_BATCH; ++k) {
            acc += x[k];
            if (v + k < v1) counts[(size_t)(v + k) * NBINS + b] = m[k] ? 0 : acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dist_divergence (initial_filter.cpp:214-231): the share of read-count bins in which the node
// collects more than READ_DIST_FACTOR_THRESHOLD of the sample's reads.  One thread per node.
struct BinCounts {
    int32_t v[NBINS];
};
// The same as a count of bins (0..50) in one byte per node: what wepp_get_node_summary moves over PCIe
// (the host divides by bins_active: the same IEEE division, bit-identical).
__global__ void divergence_count_kernel(const int32_t* __restrict__ counts, int n, BinCounts true_counts, double threshold,
                                        uint8_t* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int32_t* row = counts + (size_t)v * NBINS;
    int divergence = 0;
#pragma unroll 10
    for (int j = 0; j < NBINS; ++j) {
        const double proportion = (double)row[j] / (double)true_counts.v[j];
        divergence += proportion > threshold;
    }
    out[v] = (uint8_t)divergence;
}

__global__ void divergence_kernel(const int32_t* __restrict__ counts, int n, BinCounts true_counts, int bins_active,
                                  double threshold, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int32_t* row = counts + (size_t)v * NBINS;
    int divergence = 0;
#pragma unroll 10
    for (int j = 0; j < NBINS; ++j) {
        const double proportion = (double)row[j] / (double)true_counts.v[j];   // 0/0 = NaN compares false, as on the host
        divergence += proportion > threshold;
    }
    out[v] = (double)divergence / (double)bins_active;
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU exchange step fused with the dist_divergence evaluation, over NVLink peer memory
// (the GPU form of the reference's chunk merge, initial_filter.cpp:199-211, followed by :214-231).
// Every rank owns a slice of the nodes.  For its slice it loads the per-node read counts and
// scores of ALL ranks straight from their HBM (peer pointers opened with CUDA IPC), sums them,
// keeps the merged counts in its own counts slice, evaluates dist_divergence from the merged rows
// while they are in shared memory, and stores the merged score and dist_divergence (as its bin count, one
// byte) of the slice into every rank's output arrays.  Per rank: (W-1)/W x N x 208 B in over NVLink, N/W x 9 B x W out
// — against 2 x (W-1)/W x N x 208 B each way for a ring all-reduce of the same arrays — and the
// counts matrix never makes the second (all-gather) trip, because nothing reads it after the
// divergence is known.  Ranks synchronise around the kernel with a stream-ordered barrier
// (the caller's tiny NCCL all-reduce): peers' scans done before, peers' loads done after.
constexpr int MAX_PEERS = 8;
constexpr int PM_NODES = 64;      // nodes per block iteration (slice starts are multiples of it: 16-byte aligned rows)
constexpr int PM_THREADS = 256;

struct PeerMergeParams {
    int32_t world, rank, lo, hi;             // this rank's node slice [lo, hi)
    const int32_t* counts_in[MAX_PEERS];     // every rank's scanned counts[N][50]
    const double* score_in[MAX_PEERS];       // every rank's score[N]
    double* score_out[MAX_PEERS];            // every rank's merged score[N]
    uint8_t* div_out[MAX_PEERS];             // every rank's dist_divergence[N] as the count of bins over the threshold
    int32_t* counts_own;                     // this rank's counts: the slice is overwritten with the merged rows
    BinCounts true_counts;                   // degree-weighted reads per bin over ALL ranks (arena.cpp:138-151)
    int32_t bins_active;
    double threshold;
};

__global__ void __launch_bounds__(PM_THREADS) peer_merge_kernel(const PeerMergeParams p) {
    __shared__ __align__(16) int32_t rows[PM_NODES * NBINS];
    for (int64_t v0 = (int64_t)p.lo + (int64_t)blockIdx.x * PM_NODES; v0 < p.hi; v0 += (int64_t)gridDim.x * PM_NODES) {
        const int nv = (int)min((int64_t)PM_NODES, (int64_t)p.hi - v0);
        const int64_t base = v0 * NBINS;
        if (nv == PM_NODES) {   // 128-bit loads over NVLink, all ranks' loads of one chunk in flight together
            constexpr int N4 = PM_NODES * NBINS / 4;                 // 800 16-byte words per rank
            constexpr int U = (N4 + PM_THREADS - 1) / PM_THREADS;    // 4 per thread: all issued before the first use
            int4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = make_int4(0, 0, 0, 0);
            // two ranks' loads in flight at a time, starting from the next rank up so that the ranks do not all
            // pull from the same GPU at once (integer sums: the order does not matter)
            for (int k = 0; k < p.world; k += 2) {
                int g0 = p.rank + 1 + k, g1 = g0 + 1;
                g0 -= g0 >= p.world ? p.world : 0;
                g1 -= g1 >= p.world ? p.world : 0;
                const bool two = k + 1 < p.world;
                const int4* src0 = reinterpret_cast<const int4*>(p.counts_in[g0] + base);
                const int4* src1 = reinterpret_cast<const int4*>(p.counts_in[g1] + base);
                int4 v[U], w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = threadIdx.x + u * PM_THREADS;
                    v[u] = i < N4 ? src0[i] : make_int4(0, 0, 0, 0);
                    w[u] = (two && i < N4) ? src1[i] : make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    s[u].x += v[u].x + w[u].x; s[u].y += v[u].y + w[u].y;
                    s[u].z += v[u].z + w[u].z; s[u].w += v[u].w + w[u].w;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = threadIdx.x + u * PM_THREADS;
                if (i < N4) {
                    reinterpret_cast<int4*>(rows)[i] = s[u];
                    reinterpret_cast<int4*>(p.counts_own + base)[i] = s[u];
                }
            }
        } else {
            for (int i = threadIdx.x; i < nv * NBINS; i += PM_THREADS) {
                int32_t s = 0;
#pragma unroll
                for (int g = 0; g < MAX_PEERS; ++g)
                    if (g < p.world) s += p.counts_in[g][base + i];
                rows[i] = s;
                p.counts_own[base + i] = s;
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < nv) {
            const int64_t v = v0 + threadIdx.x;
            const int32_t* row = rows + threadIdx.x * NBINS;
            int divergence = 0;
#pragma unroll 10
            for (int j = 0; j < NBINS; ++j) {
                const double proportion = (double)row[j] / (double)p.true_counts.v[j];   // as divergence_kernel
                divergence += proportion > p.threshold;
            }
            double sc = 0.0;   // fixed rank order: every rank receives the same bits
#pragma unroll
            for (int g = 0; g < MAX_PEERS; ++g)
                if (g < p.world) sc += p.score_in[g][v];
#pragma unroll
            for (int g = 0; g < MAX_PEERS; ++g)
                if (g < p.world) {
                    p.score_out[g][v] = sc;
                    p.div_out[g][v] = (uint8_t)divergence;   // wepp_get_node_summary divides by the bins with reads
                }
        }
        __syncthreads();
    }
}

}  // namespace wepp
