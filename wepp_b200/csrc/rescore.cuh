// rescore.cuh — K4: read x candidate-haplotype mutation distance with min / argmin.
//
// Replaces haplotype::mutation_distance(const raw_read&) (reference src/WEPP/haplotype.hpp:123-177)
// applied over a candidate set with the "<= / <" argmin idiom of src/WEPP/arena.cpp:614-625 and
// :846-857 (also find_correspondents, src/WEPP/initial_filter.cpp:270-276).
//
// Host: each candidate's stack_muts (net root->node mutations, arena.cpp:18-46) is rebuilt by
// walking its root path — only candidates need it, not all N nodes as in the reference.
// Device: one thread per read, all lanes of a warp walk the same candidate so the candidate's
// mutation list is a broadcast load; the distance is the sorted-merge count of the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace wepp {

struct RescoreParams {
    int64_t n_reads;
    const int32_t* start;
    const int32_t* end;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_nuc;
    int32_t n_cand;
    const int64_t* st_off;
    const int32_t* st_pos;
    const uint8_t* st_nuc;
    int32_t* min_dist;
    int32_t* n_argmin;
    int32_t* dist;          // optional dense R x C
    const int64_t* am_off;  // fill pass
    int32_t* am_idx;
};

__device__ __forceinline__ int mutation_distance_dev(const int32_t* __restrict__ s_pos, const uint8_t* __restrict__ s_nuc,
                                                     int n_stack, const int32_t* __restrict__ c_pos,
                                                     const uint8_t* __restrict__ c_nuc, int n_comp, int min_pos,
                                                     int max_pos) {
    // first stack entry with position >= min_pos
    int lo = 0, hi = n_stack;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_pos[mid] < min_pos) lo = mid + 1; else hi = mid;
    }
    int i = lo, j = 0, muts = 0;
    while (true) {
        const bool s_ok = i < n_stack && s_pos[i] <= max_pos;
        const bool c_ok = j < n_comp;
        if (!s_ok && !c_ok) break;
        const int sp = s_ok ? s_pos[i] : 0x7FFFFFFF;
        const int cp = c_ok ? c_pos[j] : 0x7FFFFFFF;
        if (sp < cp) {            // haplotype-only position
            ++muts;
            ++i;
        } else if (cp < sp) {     // read-only position: counts unless N
            muts += c_nuc[j] != 15;
            ++j;
        } else {                  // both: counts if the alleles differ and the read is not N
            muts += (s_nuc[i] != c_nuc[j]) && (c_nuc[j] != 15);
            ++i;
            ++j;
        }
    }
    return muts;
}

// Candidate stack_muts (arena.cpp:18-46): the last event per position on the root path, kept when
// mut != ref, sorted by position.  Only candidates need them, not all N nodes as in the reference.
inline bool build_candidate_stacks(int32_t n_nodes, int32_t genome, const int32_t* parent, const int64_t* mut_off,
                                   const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t n_cand,
                                   const int32_t* cand, std::vector<int64_t>& st_off, std::vector<int32_t>& st_pos,
                                   std::vector<uint8_t>& st_nuc, std::string& err) {
    st_off.assign((size_t)n_cand + 1, 0);
    st_pos.clear();
    st_nuc.clear();
    for (int32_t c = 0; c < n_cand; ++c)
        if (cand[c] < 0 || cand[c] >= n_nodes) {
            err = "candidate node index out of range";
            return false;
        }
    // The root-path walks are latency-bound pointer chasing (every step misses the cache on an 8M-node tree):
    // spread the candidates over host threads, and inside a thread walk LANES paths in lockstep so that
    // their independent misses overlap; the event arrays of the nodes ahead are prefetched.
    const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), 32, n_cand / 16}));
    std::vector<std::vector<int32_t>> t_pos((size_t)n_thr);
    std::vector<std::vector<uint8_t>> t_nuc((size_t)n_thr);
    auto work = [&](int t) {
        constexpr int LANES = 16, AHEAD = 6;
        const int32_t c_lo = (int32_t)((int64_t)n_cand * t / n_thr), c_hi = (int32_t)((int64_t)n_cand * (t + 1) / n_thr);
        std::vector<int32_t> path[LANES], touched;
        std::vector<int64_t> last((size_t)genome + 1, -1);
        for (int32_t cb = c_lo; cb < c_hi; cb += LANES) {
            const int nb = std::min<int32_t>(LANES, c_hi - cb);
            int32_t cur[LANES];
            for (int b = 0; b < nb; ++b) {
                path[b].clear();
                cur[b] = cand[cb + b];
            }
            for (bool any = true; any;) {
                any = false;
                for (int b = 0; b < nb; ++b) {
                    const int32_t v = cur[b];
                    if (v < 0) continue;
                    __builtin_prefetch(&mut_off[v]);
                    path[b].push_back(v);
                    cur[b] = parent[v];
                    any = true;
                }
            }
            for (int b = 0; b < nb; ++b) {
                const std::vector<int32_t>& pth = path[b];
                const int np = (int)pth.size();
                touched.clear();
                for (int i = np - 1; i >= 0; --i) {   // root first: the deepest event at a position wins
                    if (i - AHEAD >= 0) {
                        const int64_t ka = mut_off[pth[(size_t)(i - AHEAD)]];
                        __builtin_prefetch(&mut_pos[ka]);
                        __builtin_prefetch(&mut_ref[ka]);
                        __builtin_prefetch(&mut_nuc[ka]);
                    }
                    const int32_t v = pth[(size_t)i];
                    for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k) {
                        if (last[mut_pos[k]] < 0) touched.push_back(mut_pos[k]);
                        last[mut_pos[k]] = k;
                    }
                }
                std::sort(touched.begin(), touched.end());
                int64_t kept = 0;
                for (int32_t pos : touched) {
                    const int64_t k = last[pos];
                    if (mut_ref[k] != mut_nuc[k]) {
                        t_pos[t].push_back(pos);
                        t_nuc[t].push_back(mut_nuc[k]);
                        ++kept;
                    }
                    last[pos] = -1;
                }
                st_off[(size_t)(cb + b) + 1] = kept;   // per-candidate count for now
            }
        }
    };
    if (n_thr == 1) {
        work(0);
    } else {
        std::vector<std::thread> thr;
        for (int t = 0; t < n_thr; ++t) thr.emplace_back(work, t);
        for (auto& th : thr) th.join();
    }
    for (int32_t c = 0; c < n_cand; ++c) st_off[(size_t)c + 1] += st_off[(size_t)c];
    st_pos.reserve((size_t)st_off[(size_t)n_cand]);
    st_nuc.reserve((size_t)st_off[(size_t)n_cand]);
    for (int t = 0; t < n_thr; ++t) {
        st_pos.insert(st_pos.end(), t_pos[t].begin(), t_pos[t].end());
        st_nuc.insert(st_nuc.end(), t_nuc[t].begin(), t_nuc[t].end());
    }
    return true;
}

template <bool FILL>
__global__ void rescore_kernel(const RescoreParams p) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.n_reads) return;
    const int64_t a = p.rm_off[r];
    const int n_comp = (int)(p.rm_off[r + 1] - a);
    const int s = p.start[r], e = p.end[r];
    int best = FILL ? p.min_dist[r] : 0x7FFFFFFF;
    int n_best = 0;
    int64_t wp = FILL ? p.am_off[r] : 0;
    for (int c = 0; c < p.n_cand; ++c) {
        const int64_t so = p.st_off[c];
        const int d = mutation_distance_dev(p.st_pos + so, p.st_nuc + so, (int)(p.st_off[c + 1] - so), p.rm_pos + a,
                                            p.rm_nuc + a, n_comp, s, e);
        if (FILL) {
            if (d == best) p.am_idx[wp++] = c;
        } else {
            if (p.dist) p.dist[r * p.n_cand + c] = d;
            if (d < best) {
                best = d;
                n_best = 1;
            } else if (d == best) {
                ++n_best;
            }
        }
    }
    if (!FILL) {
        p.min_dist[r] = best;
        p.n_argmin[r] = n_best;
    }
}

#define RS_CU(call)                                                            \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) {                                               \
            err = std::string(#call) + ": " + cudaGetErrorString(_e);          \
            cleanup();                                                         \
            return -2;                                                         \
        }                                                                      \
    } while (0)

inline int rescore_run(int device, cudaStream_t stream, int32_t n_nodes, int32_t genome, const int32_t* parent,
                       const int64_t* mut_off, const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc,
                       int64_t n_reads, const int32_t* start, const int32_t* end, const int64_t* rm_off,
                       const int32_t* rm_pos, const uint8_t* rm_nuc, int32_t n_cand, const int32_t* cand,
                       int32_t* min_dist, int32_t* dist, int64_t* am_off, int32_t* am_idx, int64_t am_capacity,
                       std::string& err) {
    (void)device;
    std::vector<int64_t> st_off;
    std::vector<int32_t> st_pos;
    std::vector<uint8_t> st_nuc;
    if (!build_candidate_stacks(n_nodes, genome, parent, mut_off, mut_pos, mut_ref, mut_nuc, n_cand, cand, st_off, st_pos,
                                st_nuc, err))
        return -1;
    int32_t *d_start = nullptr, *d_end = nullptr, *d_rm_pos = nullptr, *d_st_pos = nullptr, *d_min = nullptr,
            *d_nbest = nullptr, *d_dist = nullptr, *d_am_idx = nullptr;
    int64_t *d_rm_off = nullptr, *d_st_off = nullptr, *d_am_off = nullptr;
    uint8_t *d_rm_nuc = nullptr, *d_st_nuc = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_start); cudaFree(d_end); cudaFree(d_rm_pos); cudaFree(d_st_pos); cudaFree(d_min); cudaFree(d_nbest);
        cudaFree(d_dist); cudaFree(d_am_idx); cudaFree(d_rm_off); cudaFree(d_st_off); cudaFree(d_am_off);
        cudaFree(d_rm_nuc); cudaFree(d_st_nuc);
    };
    const int64_t nm = rm_off[n_reads];
    const size_t R = (size_t)std::max<int64_t>(n_reads, 1);
    RS_CU(cudaMalloc(&d_start, R * 4));
    RS_CU(cudaMalloc(&d_end, R * 4));
    RS_CU(cudaMalloc(&d_rm_off, (R + 1) * 8));
    RS_CU(cudaMalloc(&d_rm_pos, (size_t)std::max<int64_t>(nm, 1) * 4));
    RS_CU(cudaMalloc(&d_rm_nuc, (size_t)std::max<int64_t>(nm, 1)));
    RS_CU(cudaMalloc(&d_st_off, st_off.size() * 8));
    RS_CU(cudaMalloc(&d_st_pos, std::max<size_t>(st_pos.size(), 1) * 4));
    RS_CU(cudaMalloc(&d_st_nuc, std::max<size_t>(st_nuc.size(), 1)));
    RS_CU(cudaMalloc(&d_min, R * 4));
    RS_CU(cudaMalloc(&d_nbest, R * 4));
    if (dist) RS_CU(cudaMalloc(&d_dist, R * (size_t)n_cand * 4));
    RS_CU(cudaMemcpyAsync(d_start, start, (size_t)n_reads * 4, cudaMemcpyHostToDevice, stream));
    RS_CU(cudaMemcpyAsync(d_end, end, (size_t)n_reads * 4, cudaMemcpyHostToDevice, stream));
    RS_CU(cudaMemcpyAsync(d_rm_off, rm_off, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, stream));
    if (nm) {
        RS_CU(cudaMemcpyAsync(d_rm_pos, rm_pos, (size_t)nm * 4, cudaMemcpyHostToDevice, stream));
        RS_CU(cudaMemcpyAsync(d_rm_nuc, rm_nuc, (size_t)nm, cudaMemcpyHostToDevice, stream));
    }
    RS_CU(cudaMemcpyAsync(d_st_off, st_off.data(), st_off.size() * 8, cudaMemcpyHostToDevice, stream));
    if (!st_pos.empty()) {
        RS_CU(cudaMemcpyAsync(d_st_pos, st_pos.data(), st_pos.size() * 4, cudaMemcpyHostToDevice, stream));
        RS_CU(cudaMemcpyAsync(d_st_nuc, st_nuc.data(), st_nuc.size(), cudaMemcpyHostToDevice, stream));
    }
    RescoreParams p = {};
    p.n_reads = n_reads;
    p.start = d_start; p.end = d_end; p.rm_off = d_rm_off; p.rm_pos = d_rm_pos; p.rm_nuc = d_rm_nuc;
    p.n_cand = n_cand;
    p.st_off = d_st_off; p.st_pos = d_st_pos; p.st_nuc = d_st_nuc;
    p.min_dist = d_min; p.n_argmin = d_nbest; p.dist = d_dist;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_reads + threads - 1) / threads);
    if (n_reads > 0) {
        rescore_kernel<false><<<blocks, threads, 0, stream>>>(p);
        RS_CU(cudaGetLastError());
    }
    RS_CU(cudaMemcpyAsync(min_dist, d_min, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, stream));
    if (dist) RS_CU(cudaMemcpyAsync(dist, d_dist, (size_t)n_reads * (size_t)n_cand * 4, cudaMemcpyDeviceToHost, stream));
    if (am_off) {
        std::vector<int32_t> nbest((size_t)n_reads);
        RS_CU(cudaMemcpyAsync(nbest.data(), d_nbest, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, stream));
        RS_CU(cudaStreamSynchronize(stream));
        int64_t tot = 0;
        for (int64_t r = 0; r < n_reads; ++r) {
            am_off[r] = tot;
            tot += nbest[r];
        }
        am_off[n_reads] = tot;
        if (am_idx) {
            if (tot > am_capacity) {
                err = "am_idx capacity too small";
                cleanup();
                return -4;
            }
            RS_CU(cudaMalloc(&d_am_off, ((size_t)n_reads + 1) * 8));
            RS_CU(cudaMalloc(&d_am_idx, (size_t)std::max<int64_t>(tot, 1) * 4));
            RS_CU(cudaMemcpyAsync(d_am_off, am_off, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, stream));
            p.am_off = d_am_off;
            p.am_idx = d_am_idx;
            if (n_reads > 0) {
                rescore_kernel<true><<<blocks, threads, 0, stream>>>(p);
                RS_CU(cudaGetLastError());
            }
            RS_CU(cudaMemcpyAsync(am_idx, d_am_idx, (size_t)tot * 4, cudaMemcpyDeviceToHost, stream));
        }
    }
    RS_CU(cudaStreamSynchronize(stream));
    cleanup();
    return 0;
}

}  // namespace wepp
