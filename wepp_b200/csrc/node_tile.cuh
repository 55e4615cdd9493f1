// node_tile.cuh — per-node results in ONE pass over the nodes: the reference's chunk merge (initial_filter.cpp:199-211)
// and, from the finished rows while they are still on chip, the divergence bin count (:214-231).
//
// A bucket (window list x count bin) gives every node a value: the accumulators of the boundary entry whose range
// encloses the node, or of the leaf's own point entry — a piecewise constant function over preorder indices.
// score[v] and mapped_read_counts[v][bin] are the sums of these functions over the buckets.  A block owns NT_TILE
// consecutive nodes: for every bucket it starts from the value of the entry enclosing the tile's first node (one
// look-up through prev_boundary) and adds the changes of the bucket's entries inside the tile to a difference tile in
// shared memory; a scan down the 512 rows finishes the tile, which is written once — coalesced, 200 B rows back to
// back.  Nothing is zeroed, scattered into or re-read in HBM: the counts matrix (N x 50 x 4 B = 1.6 GB at 8 M nodes)
// is written exactly once (the previous formulation — memset, atomic scatter, chunk sums, in-place apply — moved it
// four times).  The score uses the same 128-bit fixed point differences as before (a double difference array would
// cancel catastrophically inside a tile just as well).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace wepp {

constexpr int NT_TILE = 256;                   // nodes per block
constexpr int NT_THREADS = 512;
constexpr int NT_SEG = 16;                     // rows per scan segment (a thread takes a column piece through registers)
constexpr int NT_NSEG = NT_TILE / NT_SEG;
constexpr int NT_SEG_WORDS = NT_SEG * NBINS + 4;   // + 4 words: the segments' columns fall into different banks, rows stay 16-byte aligned
constexpr int NT_SMEM_CNT = NT_NSEG * NT_SEG_WORDS * 4;
constexpr int NT_BPT = 2;                      // buckets per thread and round
constexpr int NT_BATCH = NT_THREADS * NT_BPT;  // buckets staged per round
struct NtBucket {                              // what the entry phase needs of a staged bucket
    uint32_t ent_off;                          // first entry of the bucket's list in the per-entry arrays
    uint32_t acc_off;                          // first accumulator: per (bucket, entry), or per (bucket, state)
    uint32_t e_lo_bin;                         // first entry inside the tile | count bin << 26
    int32_t first_state;
};
struct __align__(16) SAccPacked {              // a (bucket, state) accumulator pair in one 16-byte sector piece
    double w;
    int32_t c;
    int32_t pad;
};
constexpr int NT_U = 4;                        // entries in flight per thread in the entry phase
constexpr int NT_SUPER = 8;                    // consecutive tiles per block
constexpr int NT_SMEM = NT_SMEM_CNT + 4 * NT_TILE * 4 + NT_NSEG * NBINS * 4 + (NT_THREADS / 32) * 16 +
                        NT_BATCH * (int)sizeof(NtBucket) + (NT_BATCH + 4) * 4 + 16 * 4 + 52 * 4 + 16 + 52 * 4 + 4 * 4;

// Per (tile, list), tile-major so that a block reads its row with coalesced loads: ptr = first entry of the list at
// or after the tile's first node; enc = the boundary entry whose range encloses that node (the last boundary entry
// before ptr).  Also compact copies of what the tile kernel reads per entry: idx | flags.
__global__ void tile_ptr_kernel(const Entry* __restrict__ lists, const ListDesc* __restrict__ list_desc,
                                const int32_t* __restrict__ prev_boundary, int n_tiles, int n_lists,
                                int32_t* __restrict__ ptr, int32_t* __restrict__ enc, uint32_t* __restrict__ ent_x) {
    const int l = blockIdx.y;
    const ListDesc ld = list_desc[l];
    const Entry* e = lists + ld.off;
    const int32_t* pb = prev_boundary + ld.off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const uint32_t x = __ldg(&e[i].x);
        ent_x[ld.off + i] = x;
        const int c = (int)((x & IDX_MASK) / NT_TILE);
        const int cp = i > 0 ? (int)((__ldg(&e[i - 1].x) & IDX_MASK) / NT_TILE) : -1;
        for (int t = cp + 1; t <= c; ++t) {
            ptr[(size_t)t * n_lists + l] = i;
            enc[(size_t)t * n_lists + l] = pb[i];
        }
        if (i == ld.n - 1) {
            const int last_boundary = (x & ENT_POINT) ? pb[i] : i;
            for (int t = c + 1; t <= n_tiles; ++t) {
                ptr[(size_t)t * n_lists + l] = ld.n;
                enc[(size_t)t * n_lists + l] = last_boundary;
            }
        }
    }
}

constexpr uint32_t NT_REC_POINT = 0x200u, NT_REC_SKIP = 0x400u;

// entries per (tile, bucket), tile-major: the input of the scan that gives rec_off
__global__ void tile_count_kernel(const int32_t* __restrict__ tile_ptr, const BucketDesc* __restrict__ buckets, int n_tiles,
                                  int n_lists, int n_buckets, uint32_t* __restrict__ count) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > (int64_t)n_tiles * n_buckets) return;
    if (k == (int64_t)n_tiles * n_buckets) {
        count[k] = 0u;
        return;
    }
    const int t = (int)(k / n_buckets), b = (int)(k % n_buckets);
    const int l = buckets[b].list;
    count[k] = (uint32_t)(tile_ptr[(size_t)(t + 1) * n_lists + l] - tile_ptr[(size_t)t * n_lists + l]);
}

// one thread per (bucket, entry): the entry's record at its place in the tile-major order
template <bool BY_STATE>
__global__ void tile_records_kernel(const uint32_t* __restrict__ ent_x, const int32_t* __restrict__ prev_boundary,
                                    const ListDesc* __restrict__ list_desc, const BucketDesc* __restrict__ buckets,
                                    const int32_t* __restrict__ tile_ptr, const uint32_t* __restrict__ rec_off,
                                    const int32_t* __restrict__ sid, const int32_t* __restrict__ state_first,
                                    const int64_t* __restrict__ sacc_off, int n_lists, int n_buckets,
                                    uint32_t* __restrict__ rec_x, uint32_t* __restrict__ rec_cur, uint32_t* __restrict__ rec_prv) {
    const int b = blockIdx.y;
    const BucketDesc bd = buckets[b];
    const ListDesc ld = list_desc[bd.list];
    const uint32_t base = BY_STATE ? (uint32_t)sacc_off[b] : (uint32_t)bd.acc_off;
    const int32_t first = BY_STATE ? state_first[bd.list] : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const uint32_t x = ent_x[ld.off + i];
        const int idx = (int)(x & IDX_MASK), t = idx / NT_TILE;
        const uint32_t slot = rec_off[(size_t)t * n_buckets + b] + (uint32_t)(i - tile_ptr[(size_t)t * n_lists + bd.list]);
        const int pbi = prev_boundary[ld.off + i];
        uint32_t cur = 0xFFFFFFFFu, prv = 0xFFFFFFFFu;
        if (BY_STATE) {
            const int32_t sc = sid[ld.off + i], sp = pbi >= 0 ? sid[ld.off + pbi] : -1;
            if (sc >= 0) cur = base + (uint32_t)(sc - first);
            if (sp >= 0) prv = base + (uint32_t)(sp - first);
        } else {
            cur = base + (uint32_t)i;
            if (pbi >= 0) prv = base + (uint32_t)pbi;
        }
        rec_x[slot] = (uint32_t)(idx - t * NT_TILE) | ((x & ENT_POINT) ? NT_REC_POINT : 0u) | ((x & ENT_SKIP) ? NT_REC_SKIP : 0u) |
                      ((uint32_t)bd.bin << 12);
        rec_cur[slot] = cur;
        rec_prv[slot] = prv;
    }
}

struct NodeTileParams {
    const ListDesc* list_desc;
    const BucketDesc* buckets;
    int32_t n_buckets, n_lists, n_nodes, n_tiles;
    const uint32_t* ent_x;         // per list entry: idx | flags
    const int32_t* prev_boundary;
    const int32_t* tile_ptr;       // [n_tiles + 1][n_lists]
    const int32_t* tile_enc;       // [n_tiles + 1][n_lists]
    // accumulators: per (bucket, entry) ...
    const double* accS;
    const int32_t* accC;
    // ... or per (bucket, state) through the entries' state index (sid < 0: the entry is not evaluated, value 0)
    const int32_t* sid;
    const int32_t* state_first;
    const int64_t* sacc_off;
    const SAccPacked* sacc;
    // tile-major entry records (optional; nullptr: the per-bucket tables above are walked): rec_off[tile][bucket] = first
    // record of the bucket's entries inside the tile; per record x = row | flags | bin << 12 and the indices of the
    // entry's own and its predecessor's accumulator (0xFFFFFFFF: none, value 0)
    const uint32_t* rec_off;
    const uint32_t* rec_x;
    const uint32_t* rec_cur;
    const uint32_t* rec_prv;
    const uint8_t* mapped;
    double* score;
    int32_t* counts;               // [n_nodes][NBINS] or nullptr
    uint8_t* div_count;            // bins with counts / true_counts over the threshold, or nullptr
    BinCounts min_count;           // per bin: the smallest count over the threshold (INT32_MAX: none)
};

__global__ void sacc_pack_kernel(const double* __restrict__ saccS, const int32_t* __restrict__ saccC, int64_t n,
                                 SAccPacked* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = SAccPacked{saccS[i], saccC[i], 0};
}

__device__ __forceinline__ int nt_off(int row, int col) { return (row / NT_SEG) * NT_SEG_WORDS + (row % NT_SEG) * NBINS + col; }

// 128-bit add into four 32-bit limbs with native shared-memory atomics (a 64-bit shared atomic add is a compare-and-swap
// loop): every adder propagates exactly the carries its own additions produce, so concurrent adds commute.
template <int STRIDE = NT_TILE>
__device__ __forceinline__ void nt_add128(uint32_t* limb, int row, unsigned long long lo, unsigned long long hi) {
    const uint32_t a[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
    unsigned long long carry = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned long long t = (unsigned long long)a[k] + carry;
        const uint32_t add = (uint32_t)t;
        carry = t >> 32;
        if (add) {
            const uint32_t old = atomicAdd(limb + k * STRIDE + row, add);
            carry += ((unsigned long long)old + add) >> 32;
        }
    }
}

template <bool BY_STATE>
__global__ void __launch_bounds__(NT_THREADS, 2) node_tile_kernel(const NodeTileParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    int* cnt = reinterpret_cast<int*>(smem);
    uint32_t* limb = reinterpret_cast<uint32_t*>(smem + NT_SMEM_CNT);            // [4][NT_TILE]
    int* segsum = reinterpret_cast<int*>(limb + 4 * NT_TILE);
    U128* warp_tot = reinterpret_cast<U128*>(segsum + NT_NSEG * NBINS);
    NtBucket* stage = reinterpret_cast<NtBucket*>(warp_tot + NT_THREADS / 32);
    int* pre = reinterpret_cast<int*>(stage + NT_BATCH);                         // [NT_BATCH + 1] entries before each staged bucket
    int* wsum = pre + NT_BATCH + 4;                                              // [16]
    int* carry_cnt = wsum + 16;                                                  // [NBINS (+ pad)] last finished row of the counts
    U128* carry_score = reinterpret_cast<U128*>(carry_cnt + 52);                 // ... and of the score
    int* spill_cnt = reinterpret_cast<int*>(carry_score + 1);                    // [NBINS (+ pad)] changes at the row after the tile's last:
    uint32_t* spill_limb = reinterpret_cast<uint32_t*>(spill_cnt + 52);          // [4]  the way back from a leaf in the last row
    const bool with_counts = p.counts != nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto value_at = [&](uint32_t acc_off, int32_t first, int32_t st_or_i, double& s, int32_t& c) {
        // st_or_i: the entry's state (BY_STATE; < 0 = not evaluated) or the entry index (< 0 = no entry)
        s = 0.0;
        c = 0;
        if (st_or_i < 0) return;
        if (BY_STATE) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.sacc + (size_t)acc_off + (st_or_i - first)));
            s = __hiloint2double((int)v.y, (int)v.x);
            c = (int32_t)v.z;
        } else {
            s = __ldg(p.accS + (size_t)acc_off + st_or_i);
            c = __ldg(p.accC + (size_t)acc_off + st_or_i);
        }
    };
    // a block finishes NT_SUPER consecutive tiles: only the first one looks up the buckets' values at its first node,
    // the others start from the last row of the tile before
    for (int sub = 0; sub < NT_SUPER; ++sub) {
        const int tile = blockIdx.x * NT_SUPER + sub;
        if (tile >= p.n_tiles) break;
        const int c0 = tile * NT_TILE;
        const int rows = min(NT_TILE, p.n_nodes - c0);
        __syncthreads();   // the tile before has left shared memory
        if (with_counts) {
            uint4* z = reinterpret_cast<uint4*>(cnt);
            for (int i = threadIdx.x; i < NT_SMEM_CNT / 16; i += NT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int i = threadIdx.x; i < 4 * NT_TILE; i += NT_THREADS) limb[i] = 0u;
        if (threadIdx.x < 56) spill_cnt[threadIdx.x] = 0;   // spill_cnt[52] + spill_limb[4]
        __syncthreads();
        if (sub > 0) {
            if (with_counts && threadIdx.x < NBINS) cnt[nt_off(0, threadIdx.x)] = carry_cnt[threadIdx.x];
            if (threadIdx.x == 64) {
                const U128 cs = *carry_score;
                limb[0] = (uint32_t)cs.lo;
                limb[NT_TILE] = (uint32_t)(cs.lo >> 32);
                limb[2 * NT_TILE] = (uint32_t)cs.hi;
                limb[3 * NT_TILE] = (uint32_t)(cs.hi >> 32);
            }
        }
        U128 carry = {0ull, 0ull};   // this thread's share of the score at the tile's first node
        if (p.rec_x != nullptr) {
            // ---- tile-major records (tile_records_kernel): the tile's entries of all buckets lie back to back, each with
            //      the indices of its own and its predecessor's accumulator: coalesced loads, one look-up level ----------
            auto value_abs = [&](uint32_t idx, double& sv, int32_t& cv) {
                sv = 0.0;
                cv = 0;
                if (idx == 0xFFFFFFFFu) return;
                if (BY_STATE) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.sacc + idx));
                    sv = __hiloint2double((int)v.y, (int)v.x);
                    cv = (int32_t)v.z;
                } else {
                    sv = __ldg(p.accS + idx);
                    cv = __ldg(p.accC + idx);
                }
            };
            if (sub == 0) {   // the buckets' values at the tile's first node
                for (int b = threadIdx.x; b < p.n_buckets; b += NT_THREADS) {
                    const BucketDesc bd = p.buckets[b];
                    const int enc = __ldg(p.tile_enc + (size_t)tile * p.n_lists + bd.list);
                    int32_t key = enc;
                    uint32_t acc_off = (uint32_t)bd.acc_off;
                    int32_t first = 0;
                    if (BY_STATE) {
                        acc_off = (uint32_t)p.sacc_off[b];
                        first = p.state_first[bd.list];
                        if (enc >= 0) key = __ldg(p.sid + p.list_desc[bd.list].off + enc);
                    }
                    double sv;
                    int32_t cv;
                    value_at(acc_off, first, key, sv, cv);
                    if (sv != 0.0) {
                        unsigned long long lo;
                        long long hi;
                        dbl_to_fix(sv, lo, hi);
                        carry = add128(carry, U128{lo, (unsigned long long)hi});
                    }
                    if (with_counts && cv != 0) atomicAdd(&cnt[nt_off(0, bd.bin)], cv);
                }
            }
            const uint32_t r0 = __ldg(p.rec_off + (size_t)tile * p.n_buckets), r1 = __ldg(p.rec_off + (size_t)(tile + 1) * p.n_buckets);
            for (uint32_t k0 = r0 + threadIdx.x; k0 < r1; k0 += NT_U * NT_THREADS) {
                uint32_t x[NT_U], ic[NT_U], ip[NT_U];
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    const uint32_t k = k0 + u * NT_THREADS;
                    x[u] = NT_REC_SKIP;
                    ic[u] = ip[u] = 0xFFFFFFFFu;
                    if (k < r1) {
                        x[u] = __ldg(p.rec_x + k);
                        ic[u] = __ldg(p.rec_cur + k);
                        ip[u] = __ldg(p.rec_prv + k);
                    }
                }
                double cur[NT_U], prv[NT_U];
                int32_t ccur[NT_U], cprv[NT_U];
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    value_abs((x[u] & NT_REC_SKIP) ? 0xFFFFFFFFu : ic[u], cur[u], ccur[u]);
                    value_abs((x[u] & NT_REC_SKIP) ? 0xFFFFFFFFu : ip[u], prv[u], cprv[u]);
                }
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    if ((x[u] & NT_REC_SKIP) || (cur[u] == prv[u] && ccur[u] == cprv[u])) continue;
                    const int row = (int)(x[u] & 0x1FFu);
                    const int bn = (int)((x[u] >> 12) & 63u);
                    const bool point = (x[u] & NT_REC_POINT) != 0u;
                    const bool back = point && row + 1 < rows, spill = point && row + 1 == rows;
                    if (cur[u] != prv[u]) {
                        unsigned long long alo, blo;
                        long long ahi, bhi;
                        dbl_to_fix(cur[u], alo, ahi);
                        dbl_to_fix(prv[u], blo, bhi);
                        const unsigned long long lo = alo - blo;
                        const long long hi = ahi - bhi - (alo < blo ? 1 : 0);
                        nt_add128(limb, row, lo, (unsigned long long)hi);
                        if (back) nt_add128(limb, row + 1, 0ull - lo, (unsigned long long)(~hi + (lo == 0ull ? 1 : 0)));
                        if (spill) nt_add128<1>(spill_limb, 0, 0ull - lo, (unsigned long long)(~hi + (lo == 0ull ? 1 : 0)));
                    }
                    if (with_counts && ccur[u] != cprv[u]) {
                        atomicAdd(&cnt[nt_off(row, bn)], ccur[u] - cprv[u]);
                        if (back) atomicAdd(&cnt[nt_off(row + 1, bn)], cprv[u] - ccur[u]);
                        if (spill) atomicAdd(&spill_cnt[bn], cprv[u] - ccur[u]);
                    }
                }
            }
        } else
        for (int b0 = 0; b0 < p.n_buckets; b0 += NT_BATCH) {
            if (b0 > 0) __syncthreads();
            // ---- NT_BPT buckets per thread: their entries inside the tile, and (first tile of the block) their
            //      values at the tile's first node ------------------------------------------------------------------
            int n_in[NT_BPT], enc[NT_BPT], bin[NT_BPT];
            NtBucket nbk[NT_BPT];
#pragma unroll
            for (int u = 0; u < NT_BPT; ++u) {
                const int b = b0 + threadIdx.x * NT_BPT + u;
                n_in[u] = 0;
                enc[u] = -1;
                bin[u] = 0;
                nbk[u] = NtBucket{0u, 0u, 0u, 0};
                if (b < p.n_buckets) {
                    const BucketDesc bd = p.buckets[b];
                    const ListDesc ld = p.list_desc[bd.list];
                    const size_t at = (size_t)tile * p.n_lists + bd.list;
                    const int e_lo = __ldg(p.tile_ptr + at), e_hi = __ldg(p.tile_ptr + at + p.n_lists);
                    if (sub == 0) enc[u] = __ldg(p.tile_enc + at);
                    bin[u] = bd.bin;
                    nbk[u].ent_off = (uint32_t)ld.off;
                    nbk[u].acc_off = (uint32_t)(BY_STATE ? p.sacc_off[b] : bd.acc_off);
                    nbk[u].first_state = BY_STATE ? p.state_first[bd.list] : 0;
                    nbk[u].e_lo_bin = (uint32_t)e_lo | ((uint32_t)bd.bin << 26);
                    n_in[u] = e_hi - e_lo;
                }
            }
            if (sub == 0) {
#pragma unroll
                for (int u = 0; u < NT_BPT; ++u)
                    if (BY_STATE && enc[u] >= 0) enc[u] = __ldg(p.sid + (size_t)nbk[u].ent_off + enc[u]);
            }
            int tsum = 0;
#pragma unroll
            for (int u = 0; u < NT_BPT; ++u) {
                stage[threadIdx.x * NT_BPT + u] = nbk[u];
                if (sub == 0) {
                    double s;
                    int32_t c;
                    value_at(nbk[u].acc_off, nbk[u].first_state, enc[u], s, c);
                    if (s != 0.0) {
                        unsigned long long lo;
                        long long hi;
                        dbl_to_fix(s, lo, hi);
                        carry = add128(carry, U128{lo, (unsigned long long)hi});
                    }
                    if (with_counts && c != 0) atomicAdd(&cnt[nt_off(0, bin[u])], c);
                }
                tsum += n_in[u];
            }
            // exclusive prefix of the buckets' entry counts
            int inc = tsum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            int run = inc - tsum;
            for (int w = 0; w < warp; ++w) run += wsum[w];
#pragma unroll
            for (int u = 0; u < NT_BPT; ++u) {
                pre[threadIdx.x * NT_BPT + u] = run;
                run += n_in[u];
            }
            if (threadIdx.x == NT_THREADS - 1) pre[NT_BATCH] = run;
            __syncthreads();
            // ---- the staged buckets' entries, flattened over the threads; NT_U of them in flight per thread, the loads
            //      of one dependency level issued together (entry -> its and its predecessor's state -> accumulators) ----
            const int total = pre[NT_BATCH];
            for (int t0 = threadIdx.x; t0 < total; t0 += NT_U * NT_THREADS) {
                NtBucket nb[NT_U];
                int ei[NT_U];
                uint32_t x[NT_U];
                int32_t k_cur[NT_U], k_prv[NT_U];
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    const int t = t0 + u * NT_THREADS;
                    ei[u] = -1;
                    x[u] = ENT_SKIP;
                    k_prv[u] = -1;
                    if (t < total) {
                        int lo_b = 0, hi_b = NT_BATCH;   // last bucket with pre <= t
                        while (hi_b - lo_b > 1) {
                            const int mid = (lo_b + hi_b) >> 1;
                            if (pre[mid] <= t) lo_b = mid;
                            else hi_b = mid;
                        }
                        nb[u] = stage[lo_b];
                        ei[u] = (int)(nb[u].e_lo_bin & 0x03FFFFFFu) + (t - pre[lo_b]);
                    }
                }
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    if (ei[u] >= 0) {
                        x[u] = __ldg(p.ent_x + (size_t)nb[u].ent_off + ei[u]);
                        k_prv[u] = __ldg(p.prev_boundary + (size_t)nb[u].ent_off + ei[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    k_cur[u] = ei[u];
                    if (BY_STATE) {
                        const bool live = !(x[u] & ENT_SKIP);
                        k_cur[u] = live ? __ldg(p.sid + (size_t)nb[u].ent_off + ei[u]) : -1;
                        k_prv[u] = (live && k_prv[u] >= 0) ? __ldg(p.sid + (size_t)nb[u].ent_off + k_prv[u]) : -1;
                    }
                }
                double cur[NT_U], prv[NT_U];
                int32_t ccur[NT_U], cprv[NT_U];
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    cur[u] = prv[u] = 0.0;
                    ccur[u] = cprv[u] = 0;
                    if (!(x[u] & ENT_SKIP)) {
                        value_at(nb[u].acc_off, nb[u].first_state, k_cur[u], cur[u], ccur[u]);
                        value_at(nb[u].acc_off, nb[u].first_state, k_prv[u], prv[u], cprv[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < NT_U; ++u) {
                    if ((x[u] & ENT_SKIP) || (cur[u] == prv[u] && ccur[u] == cprv[u])) continue;
                    const int row = (int)(x[u] & IDX_MASK) - c0;
                    const int bn = (int)(nb[u].e_lo_bin >> 26);
                    // back to the enclosing value right after a leaf: the next row, or — from the tile's last row — the
                    // spill row that joins the carry into the block's next tile
                    const bool point = (x[u] & ENT_POINT) != 0u;
                    const bool back = point && row + 1 < rows, spill = point && row + 1 == rows;
                    if (cur[u] != prv[u]) {
                        unsigned long long alo, blo;
                        long long ahi, bhi;
                        dbl_to_fix(cur[u], alo, ahi);
                        dbl_to_fix(prv[u], blo, bhi);
                        const unsigned long long lo = alo - blo;
                        const long long hi = ahi - bhi - (alo < blo ? 1 : 0);
                        nt_add128(limb, row, lo, (unsigned long long)hi);
                        if (back) nt_add128(limb, row + 1, 0ull - lo, (unsigned long long)(~hi + (lo == 0ull ? 1 : 0)));
                        if (spill) nt_add128<1>(spill_limb, 0, 0ull - lo, (unsigned long long)(~hi + (lo == 0ull ? 1 : 0)));
                    }
                    if (with_counts && ccur[u] != cprv[u]) {
                        atomicAdd(&cnt[nt_off(row, bn)], ccur[u] - cprv[u]);
                        if (back) atomicAdd(&cnt[nt_off(row + 1, bn)], cprv[u] - ccur[u]);
                        if (spill) atomicAdd(&spill_cnt[bn], cprv[u] - ccur[u]);
                    }
                }
            }
        }
        if (sub == 0) {   // the looked-up values go to row 0 once per warp
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                U128 o;
                o.lo = __shfl_xor_sync(0xFFFFFFFFu, carry.lo, d);
                o.hi = __shfl_xor_sync(0xFFFFFFFFu, carry.hi, d);
                carry = add128(carry, o);
            }
            if (lane == 0 && (carry.lo | carry.hi) != 0ull) nt_add128(limb, 0, carry.lo, carry.hi);
        }
        __syncthreads();
        // ---- score: inclusive 128-bit scan down the rows -----------------------------------------------------------
        {
            U128 inc = {0ull, 0ull};
            if (threadIdx.x < NT_TILE) {
                inc.lo = (unsigned long long)limb[threadIdx.x] | ((unsigned long long)limb[NT_TILE + threadIdx.x] << 32);
                inc.hi = (unsigned long long)limb[2 * NT_TILE + threadIdx.x] | ((unsigned long long)limb[3 * NT_TILE + threadIdx.x] << 32);
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const U128 o = shfl_up128(inc, d);
                if (lane >= d) inc = add128(inc, o);
            }
            if (lane == 31) warp_tot[warp] = inc;
            __syncthreads();
            if ((int)threadIdx.x < rows) {   // (only the first NT_TILE threads hold a row: the other warps have nothing to add up)
                U128 off = {0, 0};
                for (int w = 0; w < warp; ++w) off = add128(off, warp_tot[w]);
                const U128 run = add128(off, inc);
                const double d = ((double)(long long)run.hi * 18446744073709551616.0 + (double)run.lo) * 8.271806125530277e-25;   // 2^-80
                p.score[c0 + threadIdx.x] = (p.mapped && p.mapped[c0 + threadIdx.x]) ? 0.0 : d;
                if ((int)threadIdx.x == rows - 1)
                    *carry_score = add128(run, U128{(unsigned long long)spill_limb[0] | ((unsigned long long)spill_limb[1] << 32),
                                                    (unsigned long long)spill_limb[2] | ((unsigned long long)spill_limb[3] << 32)});
            }
        }
        if (!with_counts) continue;
        // ---- counts: scan down the rows, column by column: segment sums, then segment prefixes + rescan in place ----
        for (int pr = threadIdx.x; pr < NT_NSEG * NBINS; pr += NT_THREADS) {
            const int seg = pr / NBINS, col = pr % NBINS;
            const int* c = cnt + seg * NT_SEG_WORDS + col;
            int v[NT_SEG];   // independent loads, then the adds
#pragma unroll
            for (int i = 0; i < NT_SEG; ++i) v[i] = c[i * NBINS];
            int sum = 0;
#pragma unroll
            for (int i = 0; i < NT_SEG; ++i) sum += v[i];
            segsum[seg * NBINS + col] = sum;
        }
        __syncthreads();
        for (int pr = threadIdx.x; pr < NT_NSEG * NBINS; pr += NT_THREADS) {
            const int seg = pr / NBINS, col = pr % NBINS;
            int* c = cnt + seg * NT_SEG_WORDS + col;
            int v[NT_SEG];
#pragma unroll
            for (int i = 0; i < NT_SEG; ++i) v[i] = c[i * NBINS];
            int run = 0;
            for (int sg = 0; sg < seg; ++sg) run += segsum[sg * NBINS + col];
#pragma unroll
            for (int i = 0; i < NT_SEG; ++i) {
                run += v[i];
                c[i * NBINS] = run;
            }
            // the tile's last row is the next tile's start (taken before mapped rows are blanked): its segment's thread
            if (seg == (rows - 1) / NT_SEG) carry_cnt[col] = c[((rows - 1) % NT_SEG) * NBINS] + spill_cnt[col];
        }
        __syncthreads();
        // ---- mapped nodes hold nothing; divergence bin count per node; the tile goes out once -----------------------
        if ((int)threadIdx.x < rows) {
            int* row = cnt + nt_off(threadIdx.x, 0);
            if (p.mapped && p.mapped[c0 + threadIdx.x]) {
#pragma unroll 10
                for (int j = 0; j < NBINS; ++j) row[j] = 0;
            }
            if (p.div_count) {
                // counts / true_counts > threshold, as integers: min_count[j] is the smallest count whose IEEE quotient
                // exceeds the threshold (found on the host with the very division the reference does, :214-231)
                int divergence = 0;
#pragma unroll 10
                for (int j = 0; j < NBINS; ++j) divergence += row[j] >= p.min_count.v[j];
                p.div_count[c0 + threadIdx.x] = (uint8_t)divergence;
            }
        }
        __syncthreads();
        uint4* out = reinterpret_cast<uint4*>(p.counts + (size_t)c0 * NBINS);
        const int words = rows * NBINS;
        for (int i = threadIdx.x; i < NT_TILE * NBINS / 4; i += NT_THREADS) {
            const int w = i * 4, sg = w / (NT_SEG * NBINS), in = w % (NT_SEG * NBINS);
            if (w + 3 < words) {
                out[i] = *reinterpret_cast<const uint4*>(cnt + sg * NT_SEG_WORDS + in);
            } else {
                for (int k = 0; k < 4; ++k)
                    if (w + k < words) p.counts[(size_t)c0 * NBINS + w + k] = cnt[sg * NT_SEG_WORDS + in + k];
            }
        }
    }
}

}  // namespace wepp
