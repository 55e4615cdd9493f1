// rescore_tiles.cuh — K4 over the RESIDENT reads: read x candidate-haplotype mutation distance with
// min / argmin, on the placement kernel's tile machinery.
//
// Replaces haplotype::mutation_distance(const raw_read&) (reference src/WEPP/haplotype.hpp:123-177)
// applied over a candidate set with the "<= / <" argmin idiom of src/WEPP/arena.cpp:614-625 and
// :846-857.  The sorted merge of the reference counts, for a read r with window [s, e] and a
// candidate c with net root->node mutations stack_muts (arena.cpp:18-46):
//     +1 for a stack mutation in the window where the read has no entry,
//     +1 for a read entry (not N) where the candidate has none,
//     +1 where both have one, the alleles differ and the read's is not N,
// which is   dist(r, c) = k(r) + sum over stack mutations (p, a) of c with s <= p <= e of h(read at p, a)
// with k(r) = the read's non-N entries and h = +1 (read as reference), -1 (read allele == a), 0 (N, or
// another allele: already counted in k).  h is a signed delta selected by the read's allele class —
// exactly what the placement kernel's shared-memory selector table and PRMT evaluate for 256 reads at
// once.  So: per window list (the reads' buckets, wepp_set_reads) the candidates' in-window stack
// mutations are laid out as 16-byte entries (the last one of a candidate carries RT_END; a candidate
// without any contributes one zero entry), a CTA takes one read tile, its 8 warps split the candidate
// range, every entry costs one selector load + K/2 PRMT + K/2 VIADD.16x2 per lane, and at a candidate's
// last entry the packed distances go through the same VIMNMX min / count as placement.
//   mode 0: min distance, number of argmins (and, optionally, the dense R x C matrix), plus the number
//           of argmins in the warps' candidate ranges before each warp's own (for mode 1)
//   mode 1: argmin lists in candidate order at CSR offsets the host derives from the counts
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_scan.cuh>

#include "kernels.cuh"

namespace wepp {

constexpr uint32_t RT_END = 0x80000000u;   // Entry::x: last entry of its candidate (x & ~RT_END = position in the candidate list)
constexpr int RT_XCH = PLACE_WARPS * 512;  // after the warps' 32-entry staging buffers
template <int K> struct RtLayout { static constexpr int CODES = RT_XCH + PLACE_WARPS * 2 * 32 * K * 4; };

struct RescoreTileParams {
    const Entry* cent;        // candidate entries of all lists, (list, candidate)-major
    const int64_t* coff;      // [n_lists * n_cand + 1] first entry of (list, candidate)
    int32_t n_cand;
    const ListDesc* list_desc;
    const BucketDesc* buckets;
    const TileDesc* tiles;
    const int32_t* start;
    const int32_t* end;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_code;   // allele class 1..4 = A,C,G,T ; 5 = N
    const int64_t* perm;
    int32_t* min_dist;        // caller order
    int32_t* n_argmin;
    int32_t* before;          // [R][PLACE_WARPS]
    int32_t* dist;            // optional dense [R][n_cand]
    const int64_t* am_off;    // mode 1
    int32_t* am_idx;
};

// entries of (list, candidate): its stack mutations inside the list's window, at least one
__global__ void cand_count_kernel(const ListDesc* __restrict__ ld, int n_lists, int n_cand,
                                  const int64_t* __restrict__ st_off, const int32_t* __restrict__ st_pos,
                                  int64_t* __restrict__ cnt, int min_one) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)n_lists * n_cand;
    if (i > n) return;
    if (i == n) {   // the scan's total lands here
        cnt[i] = 0;
        return;
    }
    const int l = (int)(i / n_cand), c = (int)(i % n_cand);
    const int b0 = ld[l].b0, b1 = b0 + ld[l].width;
    const int64_t a = st_off[c], b = st_off[c + 1];
    int64_t lo = a, hi = b;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (st_pos[mid] < b0) lo = mid + 1; else hi = mid;
    }
    const int64_t first = lo;
    hi = b;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (st_pos[mid] < b1) lo = mid + 1; else hi = mid;
    }
    cnt[i] = max(lo - first, (int64_t)min_one);   // min_one = 0: candidates with nothing in the window get no entry
}

__global__ void cand_fill_kernel(const ListDesc* __restrict__ ld, int n_lists, int n_cand,
                                 const int64_t* __restrict__ st_off, const int32_t* __restrict__ st_pos,
                                 const uint8_t* __restrict__ st_nuc, const int64_t* __restrict__ coff,
                                 Entry* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_lists * n_cand) return;
    const int l = (int)(i / n_cand), c = (int)(i % n_cand);
    const int b0 = ld[l].b0;
    const int64_t a = st_off[c], b = st_off[c + 1];
    int64_t lo = a, hi = b;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (st_pos[mid] < b0) lo = mid + 1; else hi = mid;
    }
    const int64_t o = coff[i], n = coff[i + 1] - o;
    if (n == 0) return;   // compact lists (min / count only): the kernel evaluates such candidates in bulk
    if (lo == b || st_pos[lo] >= b0 + ld[l].width) {   // nothing in the window: one zero entry
        out[o] = Entry{(uint32_t)c | RT_END, 0u, 0u, 0u};
        return;
    }
    for (int64_t k = 0; k < n; ++k) {
        const uint32_t nuc = st_nuc[lo + k];
        // delta bytes by read class: ref +1 ; the class equal to the candidate's allele -1 ; others 0
        // (an IUPAC allele equals no read allele: mutation_distance compares the 4-bit codes, haplotype.hpp:150-160)
        Entry e;
        e.x = (uint32_t)c | (k + 1 == n ? RT_END : 0u);
        e.y = 0u;
        e.z = 0x01u | (nuc == 1u ? 0x0000FF00u : 0u) | (nuc == 2u ? 0x00FF0000u : 0u) | (nuc == 4u ? 0xFF000000u : 0u);
        e.w = (nuc == 8u ? 0xFFu : 0u) | ((uint32_t)(st_pos[lo + k] - b0) << 16);
        out[o + k] = e;
    }
}

// ---------------------------------------------------------------------------------------------
// Candidate stack_muts on the device (haplotype::stack_muts, arena.cpp:18-46: the last event per position on the
// root path, kept when it differs from the reference allele, sorted by position).  One warp per candidate walks
// from the node up to the root; a per-warp bitmap over the genome in shared memory records the positions already
// decided by a deeper event, kept (position, allele) pairs are collected in shared memory and bitonic-sorted.
// The root-path walk is a chain of dependent loads — ~0.1 ms for hundreds of candidates at once here, against
// ~1 ms of cache misses on the host.  Candidates with more than SB_CAP kept mutations report -1 (host builds those).
constexpr int SB_CAP = 1024;
constexpr int SB_WARPS = 4;
constexpr int SB_MAX_GENOME = 65536 * 2;   // bitmap words are sized by the launch; this bounds the shared memory

struct StackBuildParams {
    const int32_t* parent;
    const int64_t* mut_off;
    const int32_t* mut_pos;
    const uint8_t* mut_ref;
    const uint8_t* mut_nuc;
    const int32_t* nodes;
    int32_t n, bitmap_words;
    uint32_t* out;      // [n][SB_CAP]: position << 4 | allele, ascending
    int32_t* count;     // [n]
};

__global__ void __launch_bounds__(SB_WARPS * 32) cand_stack_kernel(const StackBuildParams p) {
    extern __shared__ __align__(16) uint32_t sb_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * SB_WARPS + warp;
    if (c >= p.n) return;   // whole warps only: no block-wide barrier below
    uint32_t* bitmap = sb_smem + (size_t)warp * (p.bitmap_words + SB_CAP);
    uint32_t* buf = bitmap + p.bitmap_words;
    const unsigned FULL = 0xFFFFFFFFu;
    for (int i = lane; i < p.bitmap_words; i += 32) bitmap[i] = 0u;
    __syncwarp();
    int cnt = 0;
    for (int32_t v = p.nodes[c]; v >= 0; v = p.parent[v]) {
        const int64_t a = p.mut_off[v], b = p.mut_off[v + 1];
        for (int64_t k0 = a; k0 < b; k0 += 32) {
            const int64_t k = k0 + lane;
            bool keep = false;
            uint32_t packed = 0;
            if (k < b) {
                const uint32_t pos = (uint32_t)p.mut_pos[k];
                const uint32_t bit = 1u << (pos & 31u);
                const uint32_t old = atomicOr(&bitmap[pos >> 5], bit);   // a node has at most one event per position
                const uint32_t nuc = p.mut_nuc[k];
                keep = !(old & bit) && p.mut_ref[k] != nuc;
                packed = (pos << 4) | nuc;
            }
            const uint32_t m = __ballot_sync(FULL, keep);
            const int idx = cnt + __popc(m & ((1u << lane) - 1u));
            if (keep && idx < SB_CAP) buf[idx] = packed;
            cnt += __popc(m);
        }
        __syncwarp();
    }
    if (cnt > SB_CAP) {
        if (lane == 0) p.count[c] = -1;
        return;
    }
    int size = 32;
    while (size < cnt) size <<= 1;
    for (int i = cnt + lane; i < size; i += 32) buf[i] = 0xFFFFFFFFu;
    __syncwarp();
    for (int k = 2; k <= size; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < size; i += 32) {
                const int partner = i ^ j;
                if (partner > i) {
                    const uint32_t x = buf[i], y = buf[partner];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        buf[i] = y;
                        buf[partner] = x;
                    }
                }
            }
            __syncwarp();
        }
    uint32_t* dst = p.out + (size_t)c * SB_CAP;
    for (int i = lane; i < cnt; i += 32) dst[i] = buf[i];
    if (lane == 0) p.count[c] = cnt;
}

// scratch rows -> CSR arrays (offsets from an exclusive scan of the counts, overflowed candidates count 0)
__global__ void cand_stack_compact_kernel(const uint32_t* __restrict__ rows, const int32_t* __restrict__ count,
                                          const int64_t* __restrict__ off, int n, int32_t* __restrict__ pos,
                                          uint8_t* __restrict__ nuc) {
    const int c = blockIdx.x;
    if (c >= n) return;
    const int m = max(count[c], 0);
    const uint32_t* src = rows + (size_t)c * SB_CAP;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        pos[off[c] + i] = (int32_t)(src[i] >> 4);
        nuc[off[c] + i] = (uint8_t)(src[i] & 15u);
    }
}

__global__ void clamp_counts_kernel(const int32_t* __restrict__ count, int n, int64_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? (int64_t)max(count[i], 0) : 0;
}

template <int K, int MODE>
__global__ void __launch_bounds__(PLACE_WARPS * 32, 2) rescore_tile_kernel(const RescoreTileParams p) {
    using ST = typename Sel<K>::type;
    constexpr int P = K / 2;
    constexpr int SHIFT = Sel<K>::SHIFT;
    constexpr int T = 32 * K;
    constexpr int CODES = RtLayout<K>::CODES;
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t ebuf_s = smem_s + warp * 512;
    int* xch = reinterpret_cast<int*>(smem + RT_XCH);   // [PLACE_WARPS][2][T]
    unsigned char* codes = smem + CODES;
    const uint32_t col = smem_s + CODES + lane * K;
    const unsigned FULL = 0xFFFFFFFFu;

    const TileDesc td = p.tiles[blockIdx.x];
    const BucketDesc bd = p.buckets[td.bucket];
    const ListDesc ld = p.list_desc[bd.list];

    // ---- read tile -> shared selector table (as place_kernel) --------------------------------------
    int k_non_n[K];
    int64_t rid[K];
    {
        int s_rel[K], e_rel[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int ti = lane * K + j;
            const bool valid = ti < td.count;
            rid[j] = valid ? p.perm[td.first + ti] : -1;
            s_rel[j] = valid ? p.start[rid[j]] - ld.b0 : 1;
            e_rel[j] = valid ? p.end[rid[j]] - ld.b0 : 0;
            k_non_n[j] = 0;
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
                if (warp == 0) {
                    const int pr = p.rm_pos[k] - ld.b0;
                    codes[(pr * 32 + lane) * K + j] = (unsigned char)(c | ((c | 8u) << 4));
                }
                k_non_n[j] += (c <= 4u);
            }
        }
    }
    __syncthreads();

    // ---- this warp's candidates ---------------------------------------------------------------------
    const int C = p.n_cand;
    const int cw0 = (int)((int64_t)warp * C / PLACE_WARPS), cw1 = (int)((int64_t)(warp + 1) * C / PLACE_WARPS);
    const int64_t lbase = (int64_t)bd.list * C;
    const int64_t c0 = p.coff[lbase + cw0], c1 = p.coff[lbase + cw1];
    const Entry* ent = p.cent;

    uint32_t S0[P];
#pragma unroll
    for (int q = 0; q < P; ++q) S0[q] = S_BIAS2 + ((uint32_t)k_non_n[2 * q] | ((uint32_t)k_non_n[2 * q + 1] << 16));

    Pass1<K> st;
    uint32_t bestp[P];                 // mode 1: the reads' packed biased minima
    unsigned long long wp[K];          // mode 1: write positions
    st.bsum = (uint32_t)P * BEST_NONE2;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        st.S[q] = S0[q];
        st.B[q] = st.oB[q] = BEST_NONE2;
        bestp[q] = 0xFFFFFFFFu;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
        st.cnt[j] = 0;
        wp[j] = 0;
    }
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            uint32_t b = 0xFFFFu;   // never equals a distance: padding lanes emit nothing
            if (rid[j] >= 0) {
                b = (uint32_t)p.min_dist[rid[j]] + S_BIAS;
                wp[j] = (unsigned long long)p.am_off[rid[j]] + (unsigned long long)p.before[rid[j] * PLACE_WARPS + warp];
            }
            if (j & 1) bestp[j >> 1] = (bestp[j >> 1] & 0x0000FFFFu) | (b << 16);
            else bestp[j >> 1] = (bestp[j >> 1] & 0xFFFF0000u) | b;
        }
    }

    int n_end = 0;   // candidates with an entry run in this warp's range (mode 0)
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (c0 + lane < c1) nxt = ld_entry(ent + c0 + lane);
    for (int64_t base = c0; base < c1; base += 32) {
        sts128(ebuf_s + lane * 16, nxt);
        const uint32_t em = __ballot_sync(FULL, (nxt.x & RT_END) != 0u);
        n_end += __popc(em);
        __syncwarp();
        nxt = make_uint4(0, 0, 0, 0);   // zero entries pad the tail: no delta, no RT_END
        if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            const uint32_t ea = ebuf_s + g * 128;
            const uint32_t eg = em >> (8 * g);
            uint4 e[8];
            uint32_t sel[8][P];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = lds128(ea + i * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i) load_sel<K>(col, e[i].w >> SHIFT, sel[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int q = 0; q < P; ++q) st.S[q] = __vadd2(st.S[q], prmt(e[i].z, e[i].w, sel[i][q]));
                if (eg & (1u << i)) {   // warp-uniform: the candidate is complete
                    const uint32_t c = e[i].x & ~RT_END;
                    if (MODE == 0) {
                        st.eval(st.S, 1);
                        if (p.dist) {
#pragma unroll
                            for (int j = 0; j < K; ++j)
                                if (rid[j] >= 0)
                                    p.dist[rid[j] * C + c] = (int)((st.S[j >> 1] >> (16 * (j & 1))) & 0xFFFFu) - (int)S_BIAS;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            const uint32_t x = st.S[q] ^ bestp[q];
                            if ((x & 0xFFFFu) == 0u) p.am_idx[wp[2 * q]++] = (int32_t)c;
                            if ((x >> 16) == 0u) p.am_idx[wp[2 * q + 1]++] = (int32_t)c;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < P; ++q) st.S[q] = S0[q];
                }
            }
        }
        __syncwarp();
    }
    if constexpr (MODE == 0) {
        // compact lists: the candidates of this warp's range without an entry have no stack mutation in the window,
        // so their distance is the read's own k: one evaluation for all of them
        if (cw1 - cw0 - n_end > 0) st.eval(S0, cw1 - cw0 - n_end);

        // ---- fold the warps' (min, count) ------------------------------------------------------------------
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const uint32_t bl = st.B[q] & 0xFFFFu, bh = st.B[q] >> 16;
            xch[(warp * 2 + 0) * T + (2 * q) * 32 + lane] = bl == BEST_NONE ? 0x3FFFFFFF : (int)bl - (int)S_BIAS;
            xch[(warp * 2 + 0) * T + (2 * q + 1) * 32 + lane] = bh == BEST_NONE ? 0x3FFFFFFF : (int)bh - (int)S_BIAS;
        }
#pragma unroll
        for (int j = 0; j < K; ++j) xch[(warp * 2 + 1) * T + j * 32 + lane] = st.cnt[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (rid[j] < 0) continue;
            int gb = 0x3FFFFFFF;
#pragma unroll
            for (int w = 0; w < PLACE_WARPS; ++w) gb = min(gb, xch[(w * 2 + 0) * T + j * 32 + lane]);
            int gc = 0, before = 0;
#pragma unroll
            for (int w = 0; w < PLACE_WARPS; ++w) {
                if (xch[(w * 2 + 0) * T + j * 32 + lane] == gb) {
                    const int c = xch[(w * 2 + 1) * T + j * 32 + lane];
                    gc += c;
                    if (w < warp) before += c;
                }
            }
            p.before[rid[j] * PLACE_WARPS + warp] = before;
            if (warp == 0) {
                p.min_dist[rid[j]] = gb;
                p.n_argmin[rid[j]] = gc;
            }
        }
    }
}

}  // namespace wepp
