// kernels.cuh — sm_100a kernels of the placement path.
//
// Data layout in HBM (all built by wepp_set_arena / wepp_set_reads):
//   stripes      Entry[2E']   Euler entries of all events, grouped by genome stripe
//                              (pos / stripe_width), preorder index ascending inside a stripe
//   lists        Entry[]      one Euler list per distinct read-window stripe range: the k-way
//                              merge (by preorder index) of the stripes it covers, entry 0 = a
//                              dummy at index 0; each entry carries its segment's node count
//   reads        SoA          start/end/degree + sparse (pos, code) mutations, bucket-sorted
//   accS/accC    double/int32 per bucket, per list segment: sum of read weights / degrees whose
//                              EPP set contains the segment
//   diff_lo/hi   uint64[N+1]  128-bit fixed-point difference array for the per-node score
//   counts       int32[(N+1)*50]  difference array, scanned in place into the result
//
// Kernels (reference lines they replace in src/WEPP/initial_filter.cpp):
//   build_lists_kernel / finalize_lists_kernel   per-window Euler CSR (replaces the range trees,
//                                                 arena.cpp:68-169)
//   place_kernel<K>        K1 signed delta (:59-87) + K2 Euler prefix sum (:101-104) + K3
//                          min / multiplicity / EPP emission (:89-99, :126-134) + K3'
//                          per-segment weight accumulation (:167-177); one warp = one tile of
//                          32*K reads in lock step over the list, K reads per lane
//   expand_kernel + scan kernels   segment accumulators -> per-node score / counts (:199-211)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_prep.h"

namespace wepp {

constexpr uint32_t IDX_MASK = 0x3FFFFFFFu;
constexpr uint32_t SEG_FLAG = 0x80000000u;
constexpr int NBINS = 50;
constexpr int FIX_SHIFT = 80;  // score fixed point: value * 2^80 in a signed 128-bit integer

struct PlaceParams {
    const Entry* lists;
    const ListDesc* list_desc;
    const BucketDesc* buckets;
    const TileDesc* tiles;
    int32_t n_tiles;
    int32_t n_nodes;
    int* tile_counter;
    // reads (bucket-sorted)
    const int32_t* start;
    const int32_t* end;
    const int32_t* degree;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_code;
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

// One pair of reads at one non-empty segment (initial_filter.cpp:89-99): b = min(s, b) per half,
// and where s <= b the segment's countable nodes are added to the read's node count (a strict
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
//   per warp : 32 staged entries (512 B) + pass-2 reduction staging (double[8][36] + int[8][33])
//   per CTA  : tile id, EPP write bases, pattern tables, the read-allele selector table
// The per-warp staging areas double as the exchange buffer for the chunk summaries between
// pass 1 and pass 2.  The pattern tables must lie below 64 KB (their byte offsets are carried
// in 16-bit halves).
constexpr int PLACE_WARPS = 8;
constexpr int RED_G = 8;              // entries reduced together in pass 2
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
// List construction: rank-by-binary-search k-way merge of the stripes a list covers.
// grid = (chunks, n_lists), one thread per source entry.
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
            // stripes before mine: count idx <= mine; after mine: count idx < mine
            const uint32_t key = t < s ? e.x + 1u : e.x;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(&stripes[mid].x) < key) lo = mid + 1; else hi = mid;
            }
            rank += lo - base;
        }
        Entry o;
        o.x = e.x;
        o.y = 0;
        o.z = e.z;
        o.w = (e.w & 0xFFu) | ((e.y - (uint32_t)ld.b0) << 16);
        dst[rank] = o;
    }
}

// Segment lengths and countable-node counts.  grid = (chunks, n_lists).
__global__ void finalize_lists_kernel(Entry* __restrict__ lists, const ListDesc* __restrict__ list_desc,
                                      int n_nodes, const int32_t* __restrict__ mapped_prefix) {
    const ListDesc ld = list_desc[blockIdx.y];
    Entry* e = lists + ld.off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const uint32_t idx = e[i].x & IDX_MASK;
        const uint32_t nxt = (i + 1 < ld.n) ? (e[i + 1].x & IDX_MASK) : (uint32_t)n_nodes;
        const uint32_t len = nxt - idx;
        uint32_t ucnt = len;
        if (mapped_prefix) ucnt -= (uint32_t)(mapped_prefix[nxt] - mapped_prefix[idx]);
        e[i].y = ucnt;
        e[i].x = idx | (len ? SEG_FLAG : 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// Out-of-line (rare) EPP list emission: nodes of one argmin segment that are not mapped.
__device__ __noinline__ void emit_segment(int32_t* __restrict__ out, unsigned long long& wp, uint32_t v, uint32_t u,
                                          const uint8_t* __restrict__ mapped) {
    while (u) {
        if (!mapped || !mapped[v]) {
            out[wp++] = (int32_t)v;
            --u;
        }
        ++v;
    }
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

// The placement kernel.  Persistent CTAs; each CTA pulls tiles (32*K reads of one window bucket)
// from a global counter.  The CTA's PLACE_WARPS warps share the tile's selector table and each
// scans one contiguous chunk of the bucket's Euler list for ALL reads of the tile (lane = K
// reads held as K/2 packed s16x2 registers), a two-level Euler-tour scan:
//   pass 1  per chunk: sum of deltas, min prefix, node count at the min  -> shared memory
//           per entry and pair of reads: PRMT (signed delta pair) + VIADD.16x2 (prefix sum) +
//           VIMNMX.S16x2 with its two predicates (running min, "<= min") + two predicated adds
//           (node count); a strict decrease of any minimum is caught once per entry by comparing
//           the sum of the packed minima and repaired out of line.
//   combine every warp folds the chunk summaries: global min, multiplicity, its own start offset
//   pass 2  per chunk: segments attaining the min -> weight/degree sums into the segment
//           accumulators (staged through shared memory, one atomic per segment per tile) and
//           explicit EPP lists for reads under the cache cap.
// Entries are staged 32 at a time through shared memory (zero entries pad the chunk's tail: a
// zero entry changes nothing and is never a segment end), so both passes run branch-free over
// whole batches.  ACC = accumulate per-segment weights (wepp_place); EPP = emit explicit lists.
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
        // this warp's chunk of the list (multiple of 32 entries)
        const int cs = (((n + PLACE_WARPS - 1) / PLACE_WARPS) + 31) & ~31;
        const int c0 = min(n, warp * cs), c1 = min(n, c0 + cs);

        // ---- read tile -> shared selector table -------------------------------------------------
        int run0[K];
        int64_t rid[K];
        {
            int s_rel[K], e_rel[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int ti = lane * K + j;
                const bool valid = ti < td.count;
                rid[j] = valid ? td.first + ti : -1;
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
            uint32_t S[P], B[P], oB[P];
            uint32_t bsum = (uint32_t)P * BEST_NONE2;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                S[q] = S_BIAS2;
                B[q] = oB[q] = BEST_NONE2;
            }
#pragma unroll
            for (int j = 0; j < K; ++j) cnt[j] = 0;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (c0 + lane < c1) nxt = ld_entry(ent + c0 + lane);
            for (int base = c0; base < c1; base += 32) {
                sts128(ebuf_s + lane * 16, nxt);
                const uint32_t fm = __ballot_sync(FULL, (nxt.x & SEG_FLAG) != 0u);  // segment ends of this batch
                __syncwarp();
                nxt = make_uint4(0, 0, 0, 0);
                if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    const uint32_t ea = ebuf_s + g * 128;
                    const uint32_t fg = fm >> (8 * g);
                    // all shared-memory loads of the group are issued before the first branch
                    uint4 e[8];
                    uint32_t sel[8][P];
#pragma unroll
                    for (int i = 0; i < 8; ++i) e[i] = lds128(ea + i * 16);
#pragma unroll
                    for (int i = 0; i < 8; ++i) load_sel<K>(col, e[i].w >> SHIFT, sel[i]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int q = 0; q < P; ++q) S[q] = __vadd2(S[q], prmt(e[i].z, e[i].w, sel[i][q]));
                        if (fg & (1u << i)) {   // warp-uniform: non-empty segment ends here
                            const int ucnt = (int)e[i].y;
                            uint32_t sum = 0;
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                min_count2(S[q], B[q], cnt[2 * q], cnt[2 * q + 1], ucnt);
                                sum += B[q];
                            }
                            if (sum != bsum) {  // rare: some read reached a new strict minimum here
#pragma unroll
                                for (int q = 0; q < P; ++q) {   // oB = minima before this entry (they only change here)
                                    if ((S[q] & 0xFFFFu) < (oB[q] & 0xFFFFu)) cnt[2 * q] = ucnt;
                                    if ((S[q] >> 16) < (oB[q] >> 16)) cnt[2 * q + 1] = ucnt;
                                    oB[q] = B[q];
                                }
                                bsum = sum;
                            }
                        }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < P; ++q) {
                run[2 * q] = (int)(S[q] & 0xFFFFu) - (int)S_BIAS;
                run[2 * q + 1] = (int)(S[q] >> 16) - (int)S_BIAS;
                const uint32_t bl = B[q] & 0xFFFFu, bh = B[q] >> 16;
                best[2 * q] = bl == BEST_NONE ? 0x3FFFFFFF : (int)bl - (int)S_BIAS;
                best[2 * q + 1] = bh == BEST_NONE ? 0x3FFFFFFF : (int)bh - (int)S_BIAS;
            }
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
                    const int64_t orig = p.perm[rid[j]];
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

        // ---- pass 2: which segments attain the min -> weights into the segment accumulators,
        //      EPP node lists for reads under the cache cap.  rel = running score - min >= 0 at
        //      every non-empty segment, so min(rel, 1) is the read's NOT-at-min bit. ---------------
        double* accS = p.accS + bd.acc_off;
        int32_t* accC = p.accC + bd.acc_off;
        uint32_t rel[P];
        uint32_t small_lo = 0, small_hi = 0;
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
                const uint32_t fm = __ballot_sync(FULL, (nxt.x & SEG_FLAG) != 0u);
                __syncwarp();
                nxt = make_uint4(0, 0, 0, 0);
                if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    const uint32_t ea = ebuf_s + g * 128;
                    const uint32_t fg = fm >> (8 * g);
                    uint32_t tt[RED_G];
#pragma unroll
                    for (int i = 0; i < RED_G; ++i) {
                        const uint2 e = lds64(ea + i * 16 + 8);   // delta bytes + position
                        uint32_t sel[P];
                        load_sel<K>(col, e.y >> SHIFT, sel);
                        uint32_t ti = tbase2;
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            rel[q] = __vadd2(rel[q], prmt(e.x, e.y, sel[q]));
                            ti += __vmins2(rel[q], 0x00010001u) * (uint32_t)(TBL_PAT_STRIDE << q);
                        }
                        tt[i] = (fg & (1u << i)) ? ti : tzero;
                    }
                    if (EPP && small_mask) {   // explicit EPP lists (sorted: the list is in preorder)
#pragma unroll
                        for (int i = 0; i < RED_G; ++i) {
                            const uint32_t dd = tt[i] - tbase2;   // per half: pattern * TBL_PAT_STRIDE
                            const uint32_t hit_lo = ~((dd & 0xFFFFu) / TBL_PAT_STRIDE) & small_lo;
                            const uint32_t hit_hi = ~((dd >> 16) / TBL_PAT_STRIDE) & small_hi;
                            if (hit_lo | hit_hi) {
                                const uint2 e = lds64(ea + i * 16);   // idx | flag, countable nodes
#pragma unroll
                                for (int q = 0; q < P; ++q) {
                                    if (hit_lo & (1u << q)) emit_segment(p.epp_nodes, wp[2 * q], e.x & IDX_MASK, e.y, p.mapped);
                                    if (hit_hi & (1u << q)) emit_segment(p.epp_nodes, wp[2 * q + 1], e.x & IDX_MASK, e.y, p.mapped);
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
                        double s = 0.0;
                        uint32_t c = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            s += lds_f64(rS + 4 * k * 8);
                            c += lds32(rC + 4 * k * 4);
                        }
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
// Segment accumulators -> per-node difference arrays.  grid = (chunks, n_buckets).
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
                              const BucketDesc* __restrict__ buckets, const double* __restrict__ accS,
                              const int32_t* __restrict__ accC, unsigned long long* __restrict__ diff_lo,
                              unsigned long long* __restrict__ diff_hi, int32_t* __restrict__ counts) {
    const BucketDesc bd = buckets[blockIdx.y];
    const ListDesc ld = list_desc[bd.list];
    const Entry* e = lists + ld.off;
    const double* s = accS + bd.acc_off;
    const int32_t* c = accC + bd.acc_off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const double cur = s[i], prv = i ? s[i - 1] : 0.0;
        const int32_t ccur = c[i], cprv = i ? c[i - 1] : 0;
        if (cur == prv && ccur == cprv) continue;
        const uint32_t idx = __ldg(&e[i].x) & IDX_MASK;
        if (cur != prv) {
            unsigned long long alo, blo;
            long long ahi, bhi;
            dbl_to_fix(cur, alo, ahi);
            dbl_to_fix(prv, blo, bhi);
            const unsigned long long lo = alo - blo;
            const long long hi = ahi - bhi - (alo < blo ? 1 : 0);
            atomic_add128(diff_lo + idx, diff_hi + idx, lo, hi);
        }
        if (ccur != cprv) atomicAdd(counts + (size_t)idx * NBINS + bd.bin, ccur - cprv);
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
// counts: in-place inclusive prefix along nodes of int32[(N+1)][50].
constexpr int CNT_CHUNK = 1024;  // nodes per block

__global__ void __launch_bounds__(64) counts_chunk_sum_kernel(const int32_t* __restrict__ counts, int n,
                                                               int32_t* __restrict__ chunk_tot) {
    const int b = threadIdx.x;
    if (b >= NBINS) return;
    const int v0 = blockIdx.x * CNT_CHUNK, v1 = min(n, v0 + CNT_CHUNK);
    int32_t acc = 0;
#pragma unroll 8
    for (int v = v0; v < v1; ++v) acc += counts[(size_t)v * NBINS + b];
    chunk_tot[(size_t)blockIdx.x * NBINS + b] = acc;
}

// exclusive scan over chunks, one thread per bin (out of place so the loads pipeline)
__global__ void counts_chunk_scan_kernel(const int32_t* __restrict__ chunk_tot, int32_t* __restrict__ chunk_off,
                                         int n_chunks) {
    const int b = threadIdx.x;
    if (b >= NBINS) return;
    int32_t acc = 0;
#pragma unroll 8
    for (int c = 0; c < n_chunks; ++c) {
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
    for (int v = v0; v < v1; ++v) {
        acc += counts[(size_t)v * NBINS + b];
        counts[(size_t)v * NBINS + b] = (mapped && mapped[v]) ? 0 : acc;
    }
}

}  // namespace wepp
