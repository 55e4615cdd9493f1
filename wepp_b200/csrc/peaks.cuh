// peaks.cuh — device side of the greedy peak selection (reference: wepp_filter::step / singular_step /
// find_correspondents / remove_read, src/WEPP/initial_filter.cpp:241-453).
//
// The host loop (wepp_filter_peaks in wepp_abi.cu) keeps the per-node working score, the `mapped` flags and
// the per-read `removed` flags on the device and drives four small kernels per step:
//   peak_max_kernel      max of full_score = score * sqrt(dist_divergence) (haplotype.hpp:183-185) over the nodes
//                        still in `current` (not mapped, score > SCORE_EPSILON; initial_filter.cpp:441-449)
//   peak_collect_kernel  the nodes within SCORE_EPSILON of that maximum — the only ones `step` can look at
//                        (:412: it walks the sorted list while |full_score - top| < SCORE_EPSILON)
//   correspond_kernel    find_correspondents for the chosen peaks at once: a remaining read corresponds to a
//                        peak iff its mutation distance to it equals the read's minimum parsimony
//                        (:270-276; equal to membership in the read's EPP set), i.e. iff the minimum of the
//                        distances to the chosen peaks equals max_parsimony[read]
//   subtract_kernel      remove_read for all those reads at once: their weights were re-accumulated over their
//                        EPP sets by the placement kernel (wepp_place on the subset, no mask: "use ORIGINAL
//                        size", :314) and are subtracted from the working score
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rescore.cuh"

namespace wepp {

constexpr double PEAK_SCORE_EPSILON = 1e-9;   // src/WEPP/config.hpp:15

__global__ void peak_max_kernel(const double* __restrict__ score, const double* __restrict__ divergence,
                                const uint8_t* __restrict__ mapped, int n, unsigned long long* __restrict__ out_max) {
    __shared__ unsigned long long warp_max[8];
    unsigned long long best = 0ull;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
        const double s = score[v];
        if (!mapped[v] && s > PEAK_SCORE_EPSILON) {
            const double f = s * sqrt(divergence[v]);
            // non-negative doubles order like their bit patterns
            const unsigned long long b = f > 0.0 ? (unsigned long long)__double_as_longlong(f) : 0ull;
            best = b > best ? b : best;
        }
    }
    for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, d);
        best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = warp_max[w] > best ? warp_max[w] : best;
        if (best) atomicMax(out_max, best);
    }
}

__global__ void peak_collect_kernel(const double* __restrict__ score, const double* __restrict__ divergence,
                                    const uint8_t* __restrict__ mapped, int n, double top, int* __restrict__ count,
                                    int32_t* __restrict__ nodes, double* __restrict__ fulls, int capacity) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const double s = score[v];
    if (mapped[v] || !(s > PEAK_SCORE_EPSILON)) return;
    const double f = s * sqrt(divergence[v]);
    if (fabs(f - top) < PEAK_SCORE_EPSILON) {
        const int i = atomicAdd(count, 1);
        if (i < capacity) {
            nodes[i] = v;
            fulls[i] = f;
        }
    }
}

struct CorrespondParams {
    int64_t n_reads;
    const int32_t* start;      // caller order
    const int32_t* end;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_nuc;
    const int32_t* max_pars;
    uint8_t* removed;
    int32_t n_cand;
    const int64_t* st_off;     // candidate stack_muts CSR
    const int32_t* st_pos;
    const uint8_t* st_nuc;
    int* count;
    int64_t* list;             // newly removed reads (caller indices), unordered
};

__global__ void correspond_kernel(const CorrespondParams p) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads || p.removed[r]) return;
    const int64_t a = p.rm_off[r];
    const int n_comp = (int)(p.rm_off[r + 1] - a);
    const int s = p.start[r], e = p.end[r], want = p.max_pars[r];
    for (int c = 0; c < p.n_cand; ++c) {
        const int64_t so = p.st_off[c];
        const int d = mutation_distance_dev(p.st_pos + so, p.st_nuc + so, (int)(p.st_off[c + 1] - so), p.rm_pos + a,
                                            p.rm_nuc + a, n_comp, s, e);
        if (d == want) {
            p.removed[r] = 1;
            p.list[atomicAdd(p.count, 1)] = r;
            return;
        }
    }
}

__global__ void subtract_kernel(double* __restrict__ cur, const double* __restrict__ contrib, int n) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) cur[v] -= contrib[v];
}

__global__ void mark_kernel(uint8_t* __restrict__ mapped, const int32_t* __restrict__ nodes, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) mapped[nodes[i]] = 1;
}

}  // namespace wepp
