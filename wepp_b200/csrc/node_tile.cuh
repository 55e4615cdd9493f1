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

constexpr int NT_TILE = 512;                   // nodes per block
constexpr int NT_THREADS = 512;
constexpr int NT_SEG = 64;                     // rows per scan segment
constexpr int NT_NSEG = NT_TILE / NT_SEG;
constexpr int NT_SEG_WORDS = NT_SEG * NBINS + 4;   // + 4 words: the segments' columns fall into different banks, rows stay 16-byte aligned
constexpr int NT_SMEM_CNT = NT_NSEG * NT_SEG_WORDS * 4;
constexpr int NT_BATCH = 512;                  // buckets staged per round
struct NtBucket {                              // what the entry phase needs of a staged bucket
    int64_t ent_off;                           // first entry of the bucket's list in the per-entry arrays
    int64_t acc_off;                           // first accumulator: per (bucket, entry), or per (bucket, state)
    int32_t e_lo;                              // first entry inside the tile
    int32_t first_state;
    int32_t bin;
    int32_t pad;
};
constexpr int NT_SMEM = NT_SMEM_CNT + 4 * NT_TILE * 4 + NT_NSEG * NBINS * 4 + (NT_THREADS / 32) * 16 +
                        NT_BATCH * (int)sizeof(NtBucket) + (NT_BATCH + 1) * 4 + 16 * 4;

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
    const double* saccS;
    const int32_t* saccC;
    const uint8_t* mapped;
    double* score;
    int32_t* counts;               // [n_nodes][NBINS] or nullptr
    uint8_t* div_count;            // bins with counts / true_counts over the threshold, or nullptr
    BinCounts true_counts;
    double threshold;
};

__device__ __forceinline__ int nt_off(int row, int col) { return (row / NT_SEG) * NT_SEG_WORDS + (row % NT_SEG) * NBINS + col; }

// 128-bit add into four 32-bit limbs with native shared-memory atomics (a 64-bit shared atomic add is a compare-and-swap
// loop): every adder propagates exactly the carries its own additions produce, so concurrent adds commute.
__device__ __forceinline__ void nt_add128(uint32_t* limb, int row, unsigned long long lo, unsigned long long hi) {
    const uint32_t a[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
    unsigned long long carry = 0ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned long long t = (unsigned long long)a[k] + carry;
        const uint32_t add = (uint32_t)t;
        carry = t >> 32;
        if (add) {
            const uint32_t old = atomicAdd(limb + k * NT_TILE + row, add);
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
    int* wsum = pre + NT_BATCH + 1;                                              // [16]
    const int tile = blockIdx.x;
    const int c0 = tile * NT_TILE;
    const int rows = min(NT_TILE, p.n_nodes - c0);
    const bool with_counts = p.counts != nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (with_counts) {
        uint4* z = reinterpret_cast<uint4*>(cnt);
        for (int i = threadIdx.x; i < NT_SMEM_CNT / 16; i += NT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) limb[k * NT_TILE + threadIdx.x] = 0u;
    U128 carry = {0ull, 0ull};   // this thread's share of the score at the tile's first node
    auto value = [&](int64_t ent_off, int64_t acc_off, int32_t first, int i, double& s, int32_t& c) {
        s = 0.0;
        c = 0;
        if (i < 0) return;
        if (BY_STATE) {
            const int32_t st = __ldg(p.sid + ent_off + i);
            if (st >= 0) {
                s = __ldg(p.saccS + acc_off + (st - first));
                c = __ldg(p.saccC + acc_off + (st - first));
            }
        } else {
            s = __ldg(p.accS + acc_off + i);
            c = __ldg(p.accC + acc_off + i);
        }
    };
    for (int b0 = 0; b0 < p.n_buckets; b0 += NT_BATCH) {
        __syncthreads();
        // ---- a thread per bucket: its entries inside the tile, and its value at the tile's first node ---------------
        const int b = b0 + threadIdx.x;
        int n_in = 0;
        if (threadIdx.x < NT_BATCH && b < p.n_buckets) {
            const BucketDesc bd = p.buckets[b];
            const ListDesc ld = p.list_desc[bd.list];
            const size_t at = (size_t)tile * p.n_lists + bd.list;
            const int e_lo = __ldg(p.tile_ptr + at), e_hi = __ldg(p.tile_ptr + at + p.n_lists), enc = __ldg(p.tile_enc + at);
            NtBucket nb;
            nb.ent_off = ld.off;
            nb.acc_off = BY_STATE ? p.sacc_off[b] : bd.acc_off;
            nb.first_state = BY_STATE ? p.state_first[bd.list] : 0;
            nb.e_lo = e_lo;
            nb.bin = bd.bin;
            nb.pad = 0;
            stage[threadIdx.x] = nb;
            n_in = e_hi - e_lo;
            double s;
            int32_t c;
            value(nb.ent_off, nb.acc_off, nb.first_state, enc, s, c);
            if (s != 0.0) {
                unsigned long long lo;
                long long hi;
                dbl_to_fix(s, lo, hi);
                carry = add128(carry, U128{lo, (unsigned long long)hi});
            }
            if (with_counts && c != 0) atomicAdd(&cnt[nt_off(0, bd.bin)], c);
        }
        // exclusive prefix of the buckets' entry counts (NT_BATCH == NT_THREADS: one value per thread)
        int inc = n_in;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        pre[threadIdx.x] = woff + inc - n_in;
        if (threadIdx.x == NT_THREADS - 1) pre[NT_BATCH] = woff + inc;
        __syncthreads();
        // ---- the staged buckets' entries, flattened over the threads ------------------------------------------------
        const int total = pre[NT_BATCH];
        for (int t = threadIdx.x; t < total; t += NT_THREADS) {
            int lo_b = 0, hi_b = NT_BATCH;   // last bucket with pre <= t
            while (hi_b - lo_b > 1) {
                const int mid = (lo_b + hi_b) >> 1;
                if (pre[mid] <= t) lo_b = mid;
                else hi_b = mid;
            }
            const NtBucket nb = stage[lo_b];
            const int i = nb.e_lo + (t - pre[lo_b]);
            const uint32_t x = __ldg(p.ent_x + nb.ent_off + i);
            if (x & ENT_SKIP) continue;
            const int pbi = __ldg(p.prev_boundary + nb.ent_off + i);
            double cur, prv;
            int32_t ccur, cprv;
            value(nb.ent_off, nb.acc_off, nb.first_state, i, cur, ccur);
            value(nb.ent_off, nb.acc_off, nb.first_state, pbi, prv, cprv);
            if (cur == prv && ccur == cprv) continue;
            const int row = (int)(x & IDX_MASK) - c0;
            const bool back = (x & ENT_POINT) != 0u && row + 1 < rows;   // back to the enclosing value right after a leaf
            if (cur != prv) {
                unsigned long long alo, blo;
                long long ahi, bhi;
                dbl_to_fix(cur, alo, ahi);
                dbl_to_fix(prv, blo, bhi);
                const unsigned long long lo = alo - blo;
                const long long hi = ahi - bhi - (alo < blo ? 1 : 0);
                nt_add128(limb, row, lo, (unsigned long long)hi);
                if (back) nt_add128(limb, row + 1, 0ull - lo, (unsigned long long)(~hi + (lo == 0ull ? 1 : 0)));
            }
            if (with_counts && ccur != cprv) {
                atomicAdd(&cnt[nt_off(row, nb.bin)], ccur - cprv);
                if (back) atomicAdd(&cnt[nt_off(row + 1, nb.bin)], cprv - ccur);
            }
        }
    }
    {   // the carries go to row 0 once per warp
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
    // ---- score: inclusive 128-bit scan down the rows ---------------------------------------------------------------
    {
        U128 inc;
        inc.lo = (unsigned long long)limb[threadIdx.x] | ((unsigned long long)limb[NT_TILE + threadIdx.x] << 32);
        inc.hi = (unsigned long long)limb[2 * NT_TILE + threadIdx.x] | ((unsigned long long)limb[3 * NT_TILE + threadIdx.x] << 32);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const U128 o = shfl_up128(inc, d);
            if (lane >= d) inc = add128(inc, o);
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        U128 off = {0, 0};
        for (int w = 0; w < warp; ++w) off = add128(off, warp_tot[w]);
        const U128 run = add128(off, inc);
        if ((int)threadIdx.x < rows) {
            const double d = ((double)(long long)run.hi * 18446744073709551616.0 + (double)run.lo) * 8.271806125530277e-25;   // 2^-80
            p.score[c0 + threadIdx.x] = (p.mapped && p.mapped[c0 + threadIdx.x]) ? 0.0 : d;
        }
    }
    if (!with_counts) return;
    // ---- counts: scan down the rows, column by column: segment sums, then segment prefixes + rescan in place --------
    const int seg = threadIdx.x / NBINS, col = threadIdx.x % NBINS;
    if (seg < NT_NSEG) {
        const int* c = cnt + seg * NT_SEG_WORDS + col;
        int sum = 0;
#pragma unroll 8
        for (int i = 0; i < NT_SEG; ++i) sum += c[i * NBINS];
        segsum[seg * NBINS + col] = sum;
    }
    __syncthreads();
    if (seg < NT_NSEG) {
        int run = 0;
        for (int s = 0; s < seg; ++s) run += segsum[s * NBINS + col];
        int* c = cnt + seg * NT_SEG_WORDS + col;
#pragma unroll 8
        for (int i = 0; i < NT_SEG; ++i) {
            run += c[i * NBINS];
            c[i * NBINS] = run;
        }
    }
    __syncthreads();
    // ---- mapped nodes hold nothing; divergence bin count per node; the tile goes out once ---------------------------
    if ((int)threadIdx.x < rows) {
        int* row = cnt + nt_off(threadIdx.x, 0);
        if (p.mapped && p.mapped[c0 + threadIdx.x]) {
#pragma unroll 10
            for (int j = 0; j < NBINS; ++j) row[j] = 0;
        }
        if (p.div_count) {
            int divergence = 0;
#pragma unroll 10
            for (int j = 0; j < NBINS; ++j) {
                const double proportion = (double)row[j] / (double)p.true_counts.v[j];   // 0/0 = NaN compares false, as on the host
                divergence += proportion > p.threshold;
            }
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

}  // namespace wepp
