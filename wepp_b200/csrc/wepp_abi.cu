// wepp_abi.cu — the C ABI declared in include/wepp_b200.h: context, device memory, launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <climits>
#include <cmath>
#include <cstring>
#include <set>
#include <unordered_map>
#include <string>
#include <thread>
#include <atomic>
#include <memory>
#include <vector>

#include "../../include/wepp_b200.h"
#include "host_arena.h"
#include "host_prep.h"
#include "kernels.cuh"
#include "rescore.cuh"
#include "rescore_tiles.cuh"
#include "state_place.cuh"
#include "delta_place.cuh"
#include "node_tile.cuh"
#include "peaks.cuh"

using namespace wepp;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
}  // namespace

namespace wepp {
int abi_fail(int code, const std::string& msg) { return fail(code, msg); }   // for the other ABI translation units
}

namespace {

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(WEPP_E_CUDA, std::string(#call) + " (wepp_abi.cu:" + std::to_string(__LINE__) + "): " + cudaGetErrorString(_e)); \
    } while (0)

// Grow-only device buffer.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }   // wepp_destroy selects the device before the handle goes
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        // a buffer that has to grow again gets 25 % headroom: sizes that creep up call after call (candidate
        // sets, subsets) must not pay a cudaFree + cudaMalloc (a device synchronisation) every time
        const size_t want = p ? n + n / 4 : n;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(want, 1) * sizeof(T));
        if (e != cudaSuccess && want > n) {
            (void)cudaGetLastError();
            e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
            if (e == cudaSuccess) cap = n;
            return e;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Stream-ordered temporary from the device's memory pool (kept warm: wepp_create raises the pool's release
// threshold): the builds that run once per read set must not pay a cudaMalloc + cudaFree (a device synchronisation)
// per scratch array.
template <typename T>
struct TmpBuf {
    T* p = nullptr;
    cudaStream_t st = nullptr;
    explicit TmpBuf(cudaStream_t s) : st(s) {}
    TmpBuf(const TmpBuf&) = delete;
    TmpBuf& operator=(const TmpBuf&) = delete;
    ~TmpBuf() {
        if (p) cudaFreeAsync(p, st);
    }
    cudaError_t ensure(size_t n) {
        if (p) {
            cudaFreeAsync(p, st);
            p = nullptr;
        }
        return cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st);
    }
};

// WEPP_TIMING=2: wall time of the phases of a build on stderr (development aid; adds synchronisations)
struct Laps {
    bool on;
    cudaStream_t st;
    const char* who;
    std::chrono::steady_clock::time_point t;
    Laps(const char* w, cudaStream_t s) : on(getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) == 2), st(s), who(w), t(std::chrono::steady_clock::now()) {}
    void operator()(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[%s] %-34s %8.3f ms\n", who, what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

}  // namespace

struct wepp_handle {
    int device = 0;
    int n_sms = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    bool stats_pending = false;
    cudaEvent_t ev[6] = {};
    int32_t opt_q = 16, opt_k = 0;

    // arena
    bool has_arena = false;
    int32_t n_nodes = 0, genome = 0;
    std::vector<int32_t> parent;  // kept for rescore (root paths)
    std::vector<int64_t> mut_off;
    std::vector<int32_t> mut_pos;
    std::vector<uint8_t> mut_ref, mut_nuc;
    EulerStripes es;
    DevBuf<Entry> d_stripes;
    DevBuf<int64_t> d_stripe_off;
    DevBuf<int32_t> d_rank_tab;   // stripe-neighbourhood rank table (rank_table_kernel), built on first use
    int32_t rank_tab_d = 0;       // neighbour stripes per side it covers (0 = not built)

    // mask
    bool has_mask = false;
    DevBuf<uint8_t> d_mapped;
    DevBuf<int32_t> d_mapped_prefix;
    bool lists_final = false;

    // reads: resident on the device in caller order (uploaded once by wepp_set_reads); the host copy is
    // fetched back only when a host-side consumer needs it (subset plans, rescore) — ensure_host_reads()
    bool has_reads = false;
    bool host_reads = false;
    int64_t n_reads = 0, n_read_muts = 0;
    std::vector<int32_t> r_start, r_end, r_degree;
    std::vector<int64_t> r_off;
    std::vector<int32_t> r_pos;
    std::vector<uint8_t> r_nuc;
    DevBuf<int32_t> d_rstart, d_rend, d_rdegree, d_rpos;
    DevBuf<int64_t> d_roff;
    DevBuf<uint8_t> d_rnuc, d_rcode;
    // device keying scratch (wepp_set_reads)
    DevBuf<int32_t> d_cell, d_table, d_bucket_of_cell;
    DevBuf<unsigned long long> d_cursor, d_true_counts, d_cell_pairs, d_cell_assign;
    DevBuf<int> d_key_status;
    void* h_stage = nullptr;   // pinned: key status + true counts + cell table
    size_t h_stage_cap = 0;
    uint8_t* h_div_stage = nullptr;   // pinned: per-node divergence bin counts (wepp_get_node_summary)
    size_t h_div_cap = 0;
    DevBuf<uint8_t> d_div_count;
    bool div_count_valid = false;     // d_div_count holds the bin counts of the last place (node_tile_kernel)

    struct DevPlan {
        ReadPlan plan;
        DevBuf<int64_t> perm;
        DevBuf<ListDesc> lists;
        DevBuf<BucketDesc> buckets;
        DevBuf<TileDesc> tiles;
        DevBuf<Entry> entries;
        DevBuf<int32_t> prev_boundary;   // per list entry: enclosing / previous boundary entry
        DevBuf<int32_t> chunk_start;     // [n_lists][PLACE_WARPS + 1]
        bool final_for_mask = false;
        bool lists_built = false;        // entries hold the lists of list_ranges (for the tree of the handle)
        std::vector<std::pair<int32_t, int32_t>> list_ranges;
        DevBuf<int32_t> tile_ptr, tile_enc;   // [n_tiles + 1][n_lists]: node_tile.cuh
        DevBuf<uint32_t> ent_x;               // per list entry: idx | flags
        // tile-major entry records of the node tile kernel (full plan): valid for rec_buckets / rec_mode
        DevBuf<uint32_t> rec_off, rec_x, rec_cur, rec_prv;
        bool rec_ready = false;
        int rec_mode = -1;                    // 1: accumulators per (bucket, state); 0: per (bucket, entry)
        std::vector<BucketDesc> rec_buckets;
        bool tile_ptr_ready = false;
        // distinct window-restricted haplotypes of the lists (state_place.cuh; built on demand, no mask)
        bool states_ready = false, states_usable = false;
        std::vector<std::pair<int32_t, int32_t>> state_ranges;   // (qs, qe) of the lists the states were built for
        std::vector<std::pair<int32_t, int32_t>> state_buckets;  // (list, bin) of the buckets their accumulators were laid out for
        int32_t n_states = 0;
        int64_t sacc_total = 0;
        DevBuf<int32_t> sid, state_first;
        DevBuf<int64_t> state_eoff, sacc_off;
        DevBuf<Entry> state_ent;
        // sparse corrections over the states (delta_place.cuh): posting lists per (list, position), kept with the
        // states; window groups of the current read set, rebuilt when the reads change
        bool delta_usable = false, delta_groups_ready = false, delta_groups_usable = false;
        std::vector<int32_t> h_state_first;
        int32_t max_list_states = 0;
        DevBuf<int32_t> state_list, lpos_base;
        DevBuf<uint32_t> post_off;
        DevBuf<uint2> post;
        int32_t n_groups = 0, n_units = 0;
        int64_t delta_touch_est = 0;
        DevBuf<uint32_t> order;
        DevBuf<uint4> rec;
        DevBuf<uint2> mrec;
        DevBuf<DeltaGroup> groups;
        DevBuf<DeltaUnit> units;
        DevBuf<uint8_t> base;
        DevBuf<int32_t> whist, list_goff, list_gids, bucket_goff, Gc;
        DevBuf<double> Gw;
        DevBuf<uint32_t> gscratch;
        int64_t gscratch_words = 0;
        // a subset of the resident reads through the same path (the peak loop's remove_read): where each read sits in the
        // sorted order, the groups' first positions, and the subset's own records / work units
        std::vector<int64_t> h_group_first;
        DevBuf<uint32_t> pos_of_read, sub_pos, sub_pos_sorted;
        bool pos_ready = false;
        DevBuf<uint4> rec_sub;
        DevBuf<DeltaUnit> units_sub;
        void release() {
            pos_of_read.release(); sub_pos.release(); sub_pos_sorted.release(); rec_sub.release(); units_sub.release();
            perm.release(); lists.release(); buckets.release(); tiles.release(); entries.release();
            prev_boundary.release(); chunk_start.release(); tile_ptr.release(); tile_enc.release(); ent_x.release(); rec_off.release(); rec_x.release(); rec_cur.release(); rec_prv.release();
            sid.release(); state_first.release(); state_eoff.release(); sacc_off.release(); state_ent.release();
            state_list.release(); lpos_base.release(); post_off.release(); post.release(); order.release();
            rec.release(); mrec.release(); groups.release(); units.release(); base.release(); whist.release(); list_goff.release();
            list_gids.release(); bucket_goff.release(); Gc.release(); Gw.release(); gscratch.release();
        }
    };
    DevPlan full, sub;

    // accumulators and outputs
    DevBuf<SAccPacked> d_sacc_packed;
    DevBuf<double> d_accS, d_saccS;
    DevBuf<int32_t> d_accC, d_saccC;
    DevBuf<int32_t> d_maxpars, d_mult;
    DevBuf<double> d_score, d_divergence;
    DevBuf<int32_t> d_counts;
    int32_t true_counts[NBINS] = {};   // arena::true_read_counts, src/WEPP/arena.cpp:138-151
    DevBuf<unsigned long long> d_diff_lo, d_diff_hi;
    DevBuf<U128> d_chunk128;
    DevBuf<int32_t> d_cchunk_tot, d_cchunk_off;
    DevBuf<int64_t> d_epp_off;
    DevBuf<int32_t> d_epp_nodes;
    DevBuf<unsigned long long> d_epp_total;
    DevBuf<int> d_tile_counter;
    bool has_results = false;
    bool has_epp = false;
    int64_t epp_capacity = 0;

    // multi-GPU exchange over peer memory (peer_merge_kernel): 0 score, 1 counts, 2 merged score, 3 dist_divergence
    DevBuf<double> d_score_merged;
    void* peer_ptr[MAX_PEERS][4] = {};
    int32_t peer_rank = -1, peer_world = 0;
    int64_t peer_true_counts[NBINS] = {};
    bool peer_merged = false;   // d_score_merged / d_div_count hold the merged results of the last place
    bool peer_stale = false;    // the reads changed since wepp_peer_open: the summed true read counts are out of date
    // read-sharded ranks that agree on one plan and exchange the per-(bucket, state) accumulators: wepp_set_allreduce
    wepp_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    DevBuf<int32_t> d_table_local, d_bucket_count;
    bool shared_plan = false;   // the plan of the resident reads was derived from the all-reduced cell histogram
    bool exchange_timed = false;   // ev[4] was recorded after the exchange of the last place

    // K4 over the resident reads (rescore_tiles.cuh): candidate stacks, per-window candidate entries, results
    DevBuf<int64_t> d_st_off, d_ccnt, d_coff, d_am_off;
    DevBuf<int32_t> d_st_pos, d_rs_min, d_rs_nbest, d_rs_before, d_rs_dist, d_am_idx;
    DevBuf<uint8_t> d_st_nuc, d_cub_tmp;
    DevBuf<Entry> d_cent;
    // the tree on the device for the candidates' root-path walks (cand_stack_kernel); uploaded on first use
    bool tree_on_device = false;
    DevBuf<int32_t> d_parent, d_mut_pos, d_sb_nodes, d_sb_count, d_sb_pos;
    DevBuf<int64_t> d_mut_off, d_sb_cnt64, d_sb_off;
    DevBuf<uint8_t> d_mut_ref, d_mut_nuc, d_sb_nuc;
    DevBuf<uint32_t> d_sb_rows;
    // candidate stack_muts by node, kept across calls (the iterative loops re-score mostly the same candidates)
    std::unordered_map<int32_t, std::pair<int64_t, int32_t>> st_cache;
    std::vector<int32_t> st_cache_pos;
    std::vector<uint8_t> st_cache_nuc;

    wepp_stats stats = {};
};

namespace {

template <typename T>
cudaError_t upload(DevBuf<T>& b, const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = b.ensure(v.size());
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

// Descriptors (and, for a host-keyed plan, the permutation) to the device, then the per-window
// Euler lists.  A device-keyed plan has already written dp.perm itself.
int upload_plan(wepp_handle* h, wepp_handle::DevPlan& dp, bool host_perm) {
    ReadPlan& pl = dp.plan;
    if (host_perm) CU(upload(dp.perm, pl.perm, h->stream));
    CU(upload(dp.lists, pl.lists, h->stream));
    CU(upload(dp.buckets, pl.buckets, h->stream));
    CU(upload(dp.tiles, pl.tiles, h->stream));
    // The Euler lists are a function of the tree and of the lists' stripe ranges only: a read set that maps to the
    // same sequence of window lists as the one before (the next sample of the same protocol) finds them — and what
    // was derived from them: finalisation under the same mask, tile tables — already on the device.
    bool same_lists = dp.lists_built && dp.list_ranges.size() == pl.lists.size();
    for (size_t i = 0; same_lists && i < pl.lists.size(); ++i)
        same_lists = dp.list_ranges[i].first == pl.lists[i].qs && dp.list_ranges[i].second == pl.lists[i].qe;
    if (getenv("WEPP_NO_LIST_REUSE") && atoi(getenv("WEPP_NO_LIST_REUSE")) != 0) same_lists = false;
    if (same_lists) {
        dp.delta_groups_ready = false;
    } else {
        dp.lists_built = false;
    CU(dp.entries.ensure((size_t)pl.list_entries_total));
    CU(dp.prev_boundary.ensure((size_t)pl.list_entries_total));
    CU(dp.chunk_start.ensure(pl.lists.size() * (PLACE_WARPS + 1)));
    if (!pl.lists.empty()) {
        int max_n = 0;
        for (const ListDesc& l : pl.lists) max_n = std::max(max_n, l.n);
        dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 4096), (unsigned)pl.lists.size());
        // Entry ranks from the per-tree rank table when it fits (<= 16 GiB), else one binary search per stripe.
        int span = 0;
        for (const ListDesc& l : pl.lists) span = std::max(span, l.qe - l.qs);
        const int64_t E = (int64_t)h->es.entries.size();
        const char* no_tab = getenv("WEPP_NO_RANK_TABLE");
        const bool use_tab = span >= 1 && E > 0 && !(no_tab && atoi(no_tab) != 0) &&
                             (double)E * 2.0 * span * sizeof(int32_t) <= 16.0 * 1024 * 1024 * 1024;
        if (use_tab && h->rank_tab_d < span) {
            h->rank_tab_d = 0;
            CU(h->d_rank_tab.ensure((size_t)E * 2 * (size_t)span));
            rank_table_kernel<<<(unsigned)std::min<int64_t>((E + 255) / 256, 1 << 20), 256, 0, h->stream>>>(
                h->d_stripes.p, h->d_stripe_off.p, h->es.n_stripes, h->es.stripe_width, span, E, h->d_rank_tab.p);
            CU(cudaGetLastError());
            h->rank_tab_d = span;
        }
        if (use_tab)
            build_lists_ranked_kernel<<<grid, 256, 0, h->stream>>>(h->d_stripes.p, h->d_stripe_off.p, dp.lists.p, dp.entries.p,
                                                                   h->es.stripe_width, h->d_rank_tab.p, h->rank_tab_d, E);
        else
            build_lists_kernel<<<grid, 256, 0, h->stream>>>(h->d_stripes.p, h->d_stripe_off.p, dp.lists.p, dp.entries.p,
                                                            h->es.stripe_width);
        CU(cudaGetLastError());
    }
    dp.list_ranges.clear();
    for (const ListDesc& l : pl.lists) dp.list_ranges.emplace_back(l.qs, l.qe);
    dp.lists_built = true;
    dp.final_for_mask = false;
    dp.tile_ptr_ready = false;
    dp.rec_ready = false;
    dp.delta_groups_ready = false;   // the window groups (delta_place.cuh) belong to the read set
    }
    // the states (state_place.cuh) are a function of the tree and of the lists' stripe ranges only: they stay
    // valid while consecutive read sets map to the same sequence of window lists
    // (and to the same (list, bin) buckets: the per-(bucket, state) accumulator offsets were laid out for them)
    if (dp.states_ready) {
        bool same = dp.state_ranges.size() == pl.lists.size() && dp.state_buckets.size() == pl.buckets.size();
        for (size_t i = 0; same && i < pl.lists.size(); ++i)
            same = dp.state_ranges[i].first == pl.lists[i].qs && dp.state_ranges[i].second == pl.lists[i].qe;
        for (size_t i = 0; same && i < pl.buckets.size(); ++i)
            same = dp.state_buckets[i].first == pl.buckets[i].list && dp.state_buckets[i].second == pl.buckets[i].bin;
        if (!same) dp.states_ready = false;
    }
    return WEPP_OK;
}

// Host copy of the resident reads, for the host-side consumers (subset plans, rescore).
int ensure_host_reads(wepp_handle* h) {
    if (h->host_reads) return WEPP_OK;
    const size_t n = (size_t)h->n_reads, nm = (size_t)h->n_read_muts;
    h->r_start.resize(n); h->r_end.resize(n); h->r_degree.resize(n); h->r_off.resize(n + 1);
    h->r_pos.resize(nm); h->r_nuc.resize(nm);
    if (n) {
        CU(cudaMemcpyAsync(h->r_start.data(), h->d_rstart.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->r_end.data(), h->d_rend.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->r_degree.data(), h->d_rdegree.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaMemcpyAsync(h->r_off.data(), h->d_roff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    if (nm) {
        CU(cudaMemcpyAsync(h->r_pos.data(), h->d_rpos.p, nm * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->r_nuc.data(), h->d_rnuc.p, nm, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    h->host_reads = true;
    return WEPP_OK;
}

// The dynamic shared memory a kernel may be launched with is per-device state shared by every host thread (the ranks
// of a wepp_group launch the same kernels at the same time), so it is not set to what the launch at hand needs —
// another thread's smaller value could land between this thread's call and its launch — but once and for all to the
// device's opt-in maximum (less the kernel's static shared memory).
template <typename Kern>
cudaError_t allow_max_smem(Kern kernel, const wepp_handle* h) {
    cudaFuncAttributes fa;
    const cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(h->smem_optin - fa.sharedSizeBytes));
}

int finalize_plan(wepp_handle* h, wepp_handle::DevPlan& dp) {
    if (dp.final_for_mask || dp.plan.lists.empty()) return WEPP_OK;
    int max_n = 0;
    for (const ListDesc& l : dp.plan.lists) max_n = std::max(max_n, l.n);
    dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 4096), (unsigned)dp.plan.lists.size());
    finalize_lists_kernel<<<grid, FIN_THREADS, 0, h->stream>>>(dp.entries.p, dp.lists.p, h->n_nodes,
                                                               h->has_mask ? h->d_mapped.p : nullptr,
                                                               h->has_mask ? h->d_mapped_prefix.p : nullptr,
                                                               dp.prev_boundary.p, dp.chunk_start.p);
    CU(cudaGetLastError());
    dp.final_for_mask = true;
    return WEPP_OK;
}

static_assert(PLACE_TABLE_BYTES == 232448 - SMEM_CODES, "host_prep.h: PLACE_TABLE_BYTES out of step with the kernel's layout");

template <int K, bool ACC, bool EPP>
int launch_place(wepp_handle* h, const PlaceParams& pp, int width) {
    PlaceParams p = pp;
    p.smem_per_warp = SMEM_WARP;
    const size_t smem = (size_t)SMEM_CODES + (((size_t)width * 32 * K + 15) & ~(size_t)15);
    if (smem > h->smem_optin || width > MAX_WINDOW)
        return fail(WEPP_E_INVALID, "read window too wide for shared memory (" + std::to_string(width) + " bases)");
    CU(allow_max_smem(place_kernel<K, ACC, EPP>, h));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, place_kernel<K, ACC, EPP>, PLACE_WARPS * 32, smem));
    per_sm = std::max(per_sm, 1);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(p.n_tiles, (int64_t)per_sm * h->n_sms));
    place_kernel<K, ACC, EPP><<<grid, PLACE_WARPS * 32, smem, h->stream>>>(p);
    CU(cudaGetLastError());
    return WEPP_OK;
}

template <int K>
int launch_place_k(wepp_handle* h, const PlaceParams& pp, int width) {
    if (pp.accumulate) return pp.epp_off ? launch_place<K, true, true>(h, pp, width) : launch_place<K, true, false>(h, pp, width);
    return launch_place<K, false, true>(h, pp, width);
}

// dynamic shared memory of one delta_place_kernel CTA: DP_CTAS of them share an SM (each also costs 1 KiB of system use)
inline int delta_smem_bytes(const wepp_handle* h, int ctas) { return (int)((((h->smem_optin + 1024) / ctas) - 1024) & ~(size_t)15); }
// shared memory a CTA of `warps` warps needs for a list of s states: the fixed areas, the base scores, and at least four
// warps' nibble scratch + candidate queues
inline int64_t delta_smem_need(int64_t s, int warps) {
    return dp_fixed(warps) + ((s + 15) & ~15ll) + 4 * ((s + 255) / 256 * 128 + DP_CAND_MIN * 4);
}

// The distinct restricted haplotypes ("states") of every list of the plan: state_place.cuh.
int build_states(wepp_handle* h, wepp_handle::DevPlan& dp) {
    if (dp.states_ready) return WEPP_OK;
    dp.states_ready = true;
    dp.states_usable = false;
    dp.rec_ready = false;
    const ReadPlan& pl = dp.plan;
    dp.state_ranges.clear();
    for (const ListDesc& l : pl.lists) dp.state_ranges.emplace_back(l.qs, l.qe);
    dp.state_buckets.clear();
    for (const BucketDesc& b : pl.buckets) dp.state_buckets.emplace_back(b.list, b.bin);
    const int n_lists = (int)pl.lists.size();
    const int64_t E = pl.list_entries_total;
    if (n_lists == 0 || n_lists > SW_MAX_LISTS || E <= 0 || E > 0x7FFFFFFFll) return WEPP_OK;
    cudaStream_t st = h->stream;
    const auto t_build = std::chrono::steady_clock::now();
    Laps lap("build_states", st);
    TmpBuf<uint64_t> key(st), key2(st), h2(st);
    TmpBuf<uint32_t> val(st), val2(st);
    TmpBuf<int32_t> overflow(st), flag(st), incl(st), rep_state(st), state_ucnt(st), state_rep(st);
    DevBuf<int32_t>& state_list = dp.state_list;
    TmpBuf<int64_t> state_len(st);
    dp.delta_usable = false;
    CU(key.ensure((size_t)E)); CU(key2.ensure((size_t)E)); CU(h2.ensure((size_t)E));
    CU(val.ensure((size_t)E)); CU(val2.ensure((size_t)E));
    CU(overflow.ensure((size_t)n_lists)); CU(flag.ensure((size_t)E)); CU(incl.ensure((size_t)E));
    CU(dp.sid.ensure((size_t)E));
    TmpBuf<ChunkNet> nets(st), ctx(st);
    CU(nets.ensure((size_t)n_lists * SW_CHUNKS));
    CU(ctx.ensure((size_t)n_lists * SW_CHUNKS));
    CU(cudaMemsetAsync(overflow.p, 0, (size_t)n_lists * 4, st));
    StateWalkParams wp = {};
    wp.lists = dp.entries.p; wp.list_desc = dp.lists.p; wp.n_lists = n_lists;
    wp.key = key.p; wp.h2 = h2.p; wp.overflow = overflow.p; wp.nets = nets.p;
    const unsigned walk_blocks = (unsigned)(((int64_t)n_lists * SW_CHUNKS * 32 + 127) / 128);
    wp.pass = -1;
    state_walk_kernel<<<walk_blocks, 128, 0, st>>>(wp);
    CU(cudaGetLastError());
    state_ctx_kernel<<<(n_lists + 31) / 32, 32, 0, st>>>(nets.p, n_lists, ctx.p, overflow.p);
    CU(cudaGetLastError());
    lap("walk: chunk nets + contexts");
    wp.ctx = ctx.p;
    wp.pass = 0;
    state_walk_kernel<<<walk_blocks, 128, 0, st>>>(wp);
    CU(cudaGetLastError());
    std::vector<int32_t> ov((size_t)n_lists);
    CU(cudaMemcpyAsync(ov.data(), overflow.p, (size_t)n_lists * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int32_t o : ov)
        if (o) return WEPP_OK;   // a list with too many active positions: place_kernel serves this plan
    lap("walk: hashes");
    iota_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(val.p, E);
    CU(cudaGetLastError());
    size_t tmp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, key.p, key2.p, val.p, val2.p, (int)E, 0, 64, st));
    CU(h->d_cub_tmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.p, tmp, key.p, key2.p, val.p, val2.p, (int)E, 0, 64, st));
    state_flag_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(key2.p, val2.p, h2.p, E, flag.p);
    CU(cudaGetLastError());
    CU(cub::DeviceScan::InclusiveSum(nullptr, tmp, flag.p, incl.p, (int)E, st));
    CU(h->d_cub_tmp.ensure(tmp));
    CU(cub::DeviceScan::InclusiveSum(h->d_cub_tmp.p, tmp, flag.p, incl.p, (int)E, st));
    int32_t n_states = 0;
    CU(cudaMemcpyAsync(&n_states, incl.p + (E - 1), 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    lap("sort + flags + scan");
    if (n_states <= 0) return WEPP_OK;
    const size_t S = (size_t)n_states;
    if (!(getenv("WEPP_STATE_ORDER") && atoi(getenv("WEPP_STATE_ORDER")) == 0)) {   // (0: keep the hash order — a development switch)
        TmpBuf<uint32_t> first_entry(st), first_sorted(st), ids(st), order(st);
        TmpBuf<int32_t> newid(st);
        CU(first_entry.ensure(S)); CU(first_sorted.ensure(S)); CU(ids.ensure(S)); CU(order.ensure(S)); CU(newid.ensure(S));
        CU(cudaMemsetAsync(first_entry.p, 0xFF, S * 4, st));
        state_first_entry_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(key2.p, val2.p, incl.p, E, first_entry.p);
        iota_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(ids.p, (int64_t)S);
        CU(cudaGetLastError());
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, first_entry.p, first_sorted.p, ids.p, order.p, (int)S, 0, 32, st));
        CU(h->d_cub_tmp.ensure(tmp));
        CU(cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.p, tmp, first_entry.p, first_sorted.p, ids.p, order.p, (int)S, 0, 32, st));
        state_newid_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(order.p, n_states, newid.p);
        state_renumber_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(key2.p, newid.p, E, incl.p);
        CU(cudaGetLastError());
        lap("states in Euler order");
    }
    CU(state_ucnt.ensure(S)); CU(state_rep.ensure(S)); CU(state_list.ensure(S)); CU(state_len.ensure(S + 1));
    CU(dp.state_eoff.ensure(S + 1)); CU(dp.state_first.ensure((size_t)n_lists + 1)); CU(rep_state.ensure((size_t)E));
    CU(cudaMemsetAsync(state_ucnt.p, 0, S * 4, st));
    CU(cudaMemsetAsync(state_rep.p, 0x7F, S * 4, st));
    CU(cudaMemsetAsync(state_len.p, 0, (S + 1) * 8, st));
    CU(cudaMemsetAsync(rep_state.p, 0xFF, (size_t)E * 4, st));
    state_assign_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(key2.p, val2.p, incl.p, dp.entries.p, h2.p, E, dp.sid.p,
                                                                     state_ucnt.p, state_rep.p, state_len.p, state_list.p);
    CU(cudaGetLastError());
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, state_len.p, dp.state_eoff.p, (int)(S + 1), st));
    CU(h->d_cub_tmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, state_len.p, dp.state_eoff.p, (int)(S + 1), st));
    state_first_kernel<<<(unsigned)((S + 1 + 255) / 256), 256, 0, st>>>(state_list.p, n_states, n_lists, dp.state_first.p);
    CU(cudaGetLastError());
    state_rep_mark_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(state_rep.p, n_states, rep_state.p);
    CU(cudaGetLastError());
    int64_t total_ent = 0;
    std::vector<int32_t> first((size_t)n_lists + 1);
    CU(cudaMemcpyAsync(&total_ent, dp.state_eoff.p + S, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(first.data(), dp.state_first.p, ((size_t)n_lists + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(dp.state_ent.ensure((size_t)std::max<int64_t>(total_ent, 1)));
    lap("assign + offsets");
    wp.pass = 1;
    wp.rep_state = rep_state.p; wp.state_eoff = dp.state_eoff.p; wp.state_ucnt = state_ucnt.p;
    wp.state_first = dp.state_first.p; wp.state_ent = dp.state_ent.p;
    state_walk_kernel<<<walk_blocks, 128, 0, st>>>(wp);
    CU(cudaGetLastError());
    std::vector<int64_t> sacc((size_t)pl.buckets.size());
    int64_t acc = 0;
    for (size_t b = 0; b < pl.buckets.size(); ++b) {
        sacc[b] = acc;
        acc += first[(size_t)pl.buckets[b].list + 1] - first[(size_t)pl.buckets[b].list];
    }
    CU(upload(dp.sacc_off, sacc, st));
    lap("walk: representatives");
    if (getenv("WEPP_STATE_VERIFY") && atoi(getenv("WEPP_STATE_VERIFY")) != 0) {
        // the states were told apart by two 64-bit hashes + size: compare every evaluated entry's actual state with the
        // entries stored for the state it was assigned to (a third walk; a development / audit switch)
        TmpBuf<unsigned long long> bad_states(st);
        CU(bad_states.ensure(1));
        CU(cudaMemsetAsync(bad_states.p, 0, 8, st));
        wp.pass = 2;
        wp.sid = dp.sid.p;
        wp.mismatches = bad_states.p;
        state_walk_kernel<<<walk_blocks, 128, 0, st>>>(wp);
        CU(cudaGetLastError());
        unsigned long long n_bad = 0;
        CU(cudaMemcpyAsync(&n_bad, bad_states.p, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (n_bad) return fail(WEPP_E_STATE, "state verification: " + std::to_string(n_bad) + " list entries differ from the state they were hashed to");
        fprintf(stderr, "[wepp] state verification: %lld list entries, %d states, every entry equals its state's representative\n",
                (long long)E, n_states);
    }
    dp.h_state_first = first;
    dp.max_list_states = 0;
    for (int l = 0; l < n_lists; ++l) dp.max_list_states = std::max(dp.max_list_states, first[(size_t)l + 1] - first[(size_t)l]);
    {   // posting lists per (list, position): delta_place.cuh
        std::vector<int32_t> lpos((size_t)n_lists + 1);
        int64_t slots = 0;
        for (int l = 0; l < n_lists; ++l) {
            lpos[(size_t)l] = (int32_t)slots;
            slots += pl.lists[(size_t)l].width;
        }
        lpos[(size_t)n_lists] = (int32_t)slots;
        const int64_t cells = slots * DP_LEVELS;
        const bool fits = cells < (1ll << 30) && total_ent < (1ll << 31);
        if (fits) {
            TmpBuf<uint32_t> slot_count(st);
            TmpBuf<uint64_t> pkey(st), pkey2(st), pval(st);
            TmpBuf<int32_t> bad(st);
            const size_t TE = (size_t)std::max<int64_t>(total_ent, 1);
            CU(upload(dp.lpos_base, lpos, st));
            CU(slot_count.ensure((size_t)cells + 1)); CU(bad.ensure(1));
            CU(pkey.ensure(TE)); CU(pkey2.ensure(TE)); CU(pval.ensure(TE));
            CU(dp.post_off.ensure((size_t)cells + 1));
            CU(dp.post.ensure(TE));
            CU(cudaMemsetAsync(slot_count.p, 0, ((size_t)cells + 1) * 4, st));
            CU(cudaMemsetAsync(bad.p, 0, 4, st));
            post_pairs_kernel<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(dp.state_ent.p, dp.state_eoff.p, state_list.p, dp.state_first.p,
                                                                          dp.lpos_base.p, dp.lists.p, n_states, slot_count.p, pkey.p, pval.p, bad.p);
            CU(cudaGetLastError());
            CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, slot_count.p, dp.post_off.p, (int)(cells + 1), st));
            CU(h->d_cub_tmp.ensure(tmp));
            CU(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, slot_count.p, dp.post_off.p, (int)(cells + 1), st));
            int slot_bits = 1;
            while ((1ll << slot_bits) < cells + 1) ++slot_bits;
            static_assert(sizeof(uint2) == sizeof(uint64_t), "postings are sorted as 64-bit values");
            CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, pkey.p, pkey2.p, pval.p, reinterpret_cast<uint64_t*>(dp.post.p),
                                               (int)total_ent, 0, slot_bits, st));
            CU(h->d_cub_tmp.ensure(tmp));
            CU(cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.p, tmp, pkey.p, pkey2.p, pval.p, reinterpret_cast<uint64_t*>(dp.post.p),
                                               (int)total_ent, 0, slot_bits, st));
            int32_t h_bad = 0;
            CU(cudaMemcpyAsync(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            // shared memory of delta_place_kernel: the widest list's base scores + at least 4 warps' nibble scratch
            const int64_t s_max = dp.max_list_states;
            const int64_t need = delta_smem_need(s_max, DP_WARPS_SM);
            dp.delta_usable = h_bad == 0 && need <= (int64_t)delta_smem_bytes(h, 1);
            if (getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) != 0)
                fprintf(stderr, "[wepp timing] postings: %lld slots, %lld entries, tables %s, widest list %lld states, shared memory %lld of %lld -> %s\n",
                        (long long)slots, (long long)total_ent, h_bad ? "NOT of the allele form" : "ok", (long long)s_max, (long long)need,
                        (long long)h->smem_optin, dp.delta_usable ? "usable" : "unusable");
        }
    }
    CU(cudaStreamSynchronize(st));   // the host vectors above go out of scope
    lap("posting lists");
    dp.sacc_total = acc;
    dp.n_states = n_states;
    dp.states_usable = true;
    if (getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) != 0)
        fprintf(stderr, "[wepp timing] states: %d lists, %lld list entries -> %d distinct restricted haplotypes, %lld state entries, built in %.1f ms\n",
                n_lists, (long long)E, n_states, (long long)total_ent,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_build).count());
    return WEPP_OK;
}

// Window groups of the current read set for delta_place_kernel: the reads sorted by (bucket, window), one group per
// distinct key, base scores + histogram per group, work units of <= DP_UNIT reads.
int build_delta_groups(wepp_handle* h, wepp_handle::DevPlan& dp) {
    if (dp.delta_groups_ready) return WEPP_OK;
    dp.delta_groups_ready = true;
    dp.delta_groups_usable = false;
    const ReadPlan& pl = dp.plan;
    const int64_t R = pl.n_reads;
    const int n_lists = (int)pl.lists.size(), n_buckets = (int)pl.buckets.size(), n_tiles = (int)pl.tiles.size();
    if (!dp.delta_usable || R <= 0 || R > 0x7FFFFFFFll || n_tiles == 0 || n_buckets >= (1 << 20)) return WEPP_OK;
    cudaStream_t st = h->stream;
    Laps lap("build_delta_groups", st);
    TmpBuf<uint64_t> key(st), key2(st), ukey(st);
    TmpBuf<uint32_t> val(st);
    TmpBuf<int32_t> ucount(st), nruns(st);
    CU(key.ensure((size_t)R)); CU(key2.ensure((size_t)R)); CU(ukey.ensure((size_t)R));
    CU(val.ensure((size_t)R)); CU(dp.order.ensure((size_t)R)); CU(ucount.ensure((size_t)R)); CU(nruns.ensure(1));
    delta_keys_kernel<<<n_tiles, 256, 0, st>>>(dp.tiles.p, dp.buckets.p, dp.lists.p, dp.perm.p, h->d_rstart.p, h->d_rend.p, h->d_roff.p,
                                               h->d_rpos.p, h->d_rcode.p, dp.post_off.p, dp.lpos_base.p, key.p, val.p);
    CU(cudaGetLastError());
    int key_bits = 24 + DP_COST_BITS;
    while ((1ll << (key_bits - 24 - DP_COST_BITS)) < n_buckets) ++key_bits;
    size_t tmp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, key.p, key2.p, val.p, dp.order.p, (int)R, 0, key_bits, st));
    CU(h->d_cub_tmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortPairs(h->d_cub_tmp.p, tmp, key.p, key2.p, val.p, dp.order.p, (int)R, 0, key_bits, st));
    if (h->n_read_muts >= (1ll << 32)) return WEPP_OK;   // 32-bit mutation offsets in the read records
    CU(dp.rec.ensure((size_t)R));
    CU(dp.mrec.ensure((size_t)std::max<int64_t>(h->n_read_muts, 1)));
    delta_records_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(dp.order.p, R, h->d_rdegree.p, h->d_roff.p, h->d_rcode.p, dp.rec.p);
    CU(cudaGetLastError());
    delta_window_of_key_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(key2.p, R, key.p);   // the sort is done with key
    CU(cudaGetLastError());
    CU(cub::DeviceRunLengthEncode::Encode(nullptr, tmp, key.p, ukey.p, ucount.p, nruns.p, (int)R, st));
    CU(h->d_cub_tmp.ensure(tmp));
    CU(cub::DeviceRunLengthEncode::Encode(h->d_cub_tmp.p, tmp, key.p, ukey.p, ucount.p, nruns.p, (int)R, st));
    int32_t n_groups = 0;
    CU(cudaMemcpyAsync(&n_groups, nruns.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    lap("keys + sort + records + run lengths");
    // many distinct windows per read: the per-group work (base scores of every state) outweighs the sparse reads
    const bool force = getenv("WEPP_DELTA_PLACE") && atoi(getenv("WEPP_DELTA_PLACE")) == 2;   // tests
    if (n_groups <= 0 || (!force && (int64_t)n_groups * 2 > R + 64)) return WEPP_OK;
    std::vector<uint64_t> hk((size_t)n_groups);
    std::vector<int32_t> hc((size_t)n_groups);
    CU(cudaMemcpyAsync(hk.data(), ukey.p, (size_t)n_groups * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hc.data(), ucount.p, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<DeltaGroup> groups((size_t)n_groups);
    std::vector<DeltaUnit> units;
    std::vector<int32_t> bucket_goff((size_t)n_buckets + 1, 0), list_goff((size_t)n_lists + 1, 0), list_gids((size_t)n_groups);
    int64_t base_total = 0, first = 0;
    dp.h_group_first.assign((size_t)n_groups + 1, 0);
    dp.pos_ready = false;
    for (int g = 0; g < n_groups; ++g) {
        dp.h_group_first[(size_t)g] = first;
        DeltaGroup& dg = groups[(size_t)g];
        dg.bucket = (int32_t)(hk[(size_t)g] >> 24);
        dg.list = pl.buckets[(size_t)dg.bucket].list;
        dg.a_rel = (int32_t)((hk[(size_t)g] >> 12) & 0xFFFu);
        dg.b_rel = (int32_t)(hk[(size_t)g] & 0xFFFu);
        dg.m0 = 0;
        dg.prune = dg.a_rel <= DP_MARGIN && dg.b_rel >= pl.lists[(size_t)dg.list].width - 1 - DP_MARGIN;   // the window holds the list's core
        dg.base_off = base_total;
        const int64_t s_n = dp.h_state_first[(size_t)dg.list + 1] - dp.h_state_first[(size_t)dg.list];
        base_total += (s_n + 15) & ~15ll;
        ++bucket_goff[(size_t)dg.bucket + 1];
        ++list_goff[(size_t)dg.list + 1];
        for (int32_t o = 0; o < hc[(size_t)g];) {   // units of DP_UNIT reads; a remainder of up to 2 * DP_UNIT stays whole
            const int32_t left = hc[(size_t)g] - o;
            const int32_t take = left <= 2 * DP_UNIT ? left : DP_UNIT;
            units.push_back(DeltaUnit{g, (int32_t)(first + o), take, 0});
            o += take;
        }
        first += hc[(size_t)g];
    }
    dp.h_group_first[(size_t)n_groups] = first;
    if (base_total > (8ll << 30)) return WEPP_OK;
    for (int b = 0; b < n_buckets; ++b) bucket_goff[(size_t)b + 1] += bucket_goff[(size_t)b];
    for (int l = 0; l < n_lists; ++l) list_goff[(size_t)l + 1] += list_goff[(size_t)l];
    {
        std::vector<int32_t> cur(list_goff.begin(), list_goff.end() - 1);
        for (int g = 0; g < n_groups; ++g) list_gids[(size_t)cur[(size_t)groups[(size_t)g].list]++] = g;
    }
    lap("group descriptors (host)");
    CU(upload(dp.groups, groups, st));
    CU(upload(dp.units, units, st));
    CU(upload(dp.bucket_goff, bucket_goff, st));
    CU(upload(dp.list_goff, list_goff, st));
    CU(upload(dp.list_gids, list_gids, st));
    CU(dp.base.ensure((size_t)std::max<int64_t>(base_total, 16)));
    CU(dp.whist.ensure((size_t)n_groups * DP_BINS));
    CU(dp.Gw.ensure((size_t)n_groups * DP_BINS));
    CU(dp.Gc.ensure((size_t)n_groups * DP_BINS));
    CU(cudaMemsetAsync(dp.whist.p, 0, (size_t)n_groups * DP_BINS * 4, st));
    CU(cudaMemsetAsync(dp.base.p, 0, (size_t)std::max<int64_t>(base_total, 16), st));
    WindowBaseParams wb = {};
    wb.state_ent = dp.state_ent.p; wb.state_eoff = dp.state_eoff.p; wb.state_first = dp.state_first.p;
    wb.list_goff = dp.list_goff.p; wb.list_gids = dp.list_gids.p; wb.groups = dp.groups.p; wb.base = dp.base.p; wb.whist = dp.whist.p;
    dim3 grid((unsigned)((dp.max_list_states + 255) / 256), (unsigned)n_lists);
    window_base_kernel<<<grid, 256, 0, st>>>(wb);
    CU(cudaGetLastError());
    window_m0_kernel<<<(n_groups + 255) / 256, 256, 0, st>>>(dp.whist.p, n_groups, dp.groups.p);
    CU(cudaGetLastError());
    // the postings every read mutation walks: they depend on the windows' minima (WEPP_DELTA_PRUNE=0: whole slots)
    delta_mrec_kernel<<<(unsigned)units.size(), 128, 0, st>>>(dp.units.p, dp.groups.p, dp.lists.p, dp.lpos_base.p, dp.post_off.p, dp.rec.p,
                                                              h->d_rpos.p, h->d_rcode.p,
                                                              !(getenv("WEPP_DELTA_PRUNE") && atoi(getenv("WEPP_DELTA_PRUNE")) == 0), dp.mrec.p);
    CU(cudaGetLastError());
    // byte scratch in global memory for the reads with many mutations: one area per warp of the persistent grid
    const int64_t words = ((int64_t)dp.max_list_states + 3) / 4 + 4;
    const size_t need = (size_t)h->n_sms * DP_WARPS_SM * (size_t)words;
    if (need > dp.gscratch.cap) {
        CU(dp.gscratch.ensure(need));
        CU(cudaMemsetAsync(dp.gscratch.p, 0, dp.gscratch.cap * 4, st));   // the kernel leaves it zero
    }
    dp.gscratch_words = words;
    CU(cudaStreamSynchronize(st));   // the host vectors above go out of scope
    lap("base scores + histograms");
    dp.n_groups = n_groups;
    dp.n_units = (int32_t)units.size();
    dp.delta_groups_usable = true;
    if (getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) != 0)
        fprintf(stderr, "[wepp timing] window groups: %d groups, %d units, %.1f MB of base scores\n", n_groups, dp.n_units, base_total / 1e6);
    return WEPP_OK;
}

template <int K>
int launch_state_place(wepp_handle* h, const StatePlaceParams& p, int n_tiles, int width) {
    const size_t smem = (size_t)RtLayout<K>::CODES + 2 * TBL_HALF + (((size_t)width * 32 * K + 15) & ~(size_t)15);
    if (smem > h->smem_optin || width > MAX_WINDOW)
        return fail(WEPP_E_INVALID, "read window too wide for shared memory (" + std::to_string(width) + " bases)");
    CU(allow_max_smem(state_place_kernel<K>, h));
    state_place_kernel<K><<<n_tiles, PLACE_WARPS * 32, smem, h->stream>>>(p);
    CU(cudaGetLastError());
    return WEPP_OK;
}

// (`sub`: only these reads of the FULL plan, through its sparse-correction path — work units over the subset's own
// records; the states, posting lists and window groups are those of the whole read set.  Per-node weights go to
// score_out through the difference arrays, nothing else of the handle's results changes.)
struct DeltaSubset {
    const DeltaUnit* units;
    int32_t n_units;
    const uint4* rec;
};
int run_place(wepp_handle* h, wepp_handle::DevPlan& dp, bool accumulate, int32_t epp_cap, int64_t epp_capacity,
              bool with_counts = true, double* score_out = nullptr, const DeltaSubset* sub = nullptr) {
    ReadPlan& pl = dp.plan;
    int rc = finalize_plan(h, dp);
    if (rc) return rc;
    const int n = h->n_nodes;
    int64_t launches = 0;
    h->peer_merged = false;

    CU(h->d_maxpars.ensure((size_t)h->n_reads));
    CU(h->d_mult.ensure((size_t)h->n_reads));
    CU(h->d_tile_counter.ensure(1));
    CU(cudaMemsetAsync(h->d_tile_counter.p, 0, sizeof(int), h->stream));
    CU(h->d_epp_total.ensure(1));
    CU(cudaMemsetAsync(h->d_epp_total.p, 0, sizeof(unsigned long long), h->stream));
    const bool want_epp = epp_cap > 0 && epp_capacity > 0;
    if (want_epp) {
        CU(h->d_epp_off.ensure((size_t)h->n_reads));
        CU(h->d_epp_nodes.ensure((size_t)epp_capacity));
        CU(cudaMemsetAsync(h->d_epp_off.p, 0xFF, (size_t)h->n_reads * sizeof(int64_t), h->stream));
    }
    // per-node results: one pass over node tiles (node_tile.cuh); WEPP_NODE_TILES=0 keeps the difference arrays in
    // HBM (expand_kernel + scans), as does a tile pointer table over 2 GiB
    const int n_node_tiles = (n + NT_TILE - 1) / NT_TILE;
    const bool tiles_env = !(getenv("WEPP_NODE_TILES") && atoi(getenv("WEPP_NODE_TILES")) == 0);
    const bool node_tiles = accumulate && tiles_env && !sub && !pl.lists.empty() && pl.acc_total < (1ll << 32) &&
                            pl.list_entries_total < (1ll << 32) && pl.buckets.size() < (1u << 24) &&
                            (double)pl.lists.size() * (n_node_tiles + 1) * 8.0 <= 2.0 * 1024 * 1024 * 1024;
    Laps lap_place("run_place", h->stream);
    if (node_tiles && !dp.tile_ptr_ready) {
        CU(dp.tile_ptr.ensure(pl.lists.size() * ((size_t)n_node_tiles + 1)));
        CU(dp.tile_enc.ensure(pl.lists.size() * ((size_t)n_node_tiles + 1)));
        CU(dp.ent_x.ensure((size_t)pl.list_entries_total));
        int max_n = 0;
        for (const ListDesc& l : pl.lists) max_n = std::max(max_n, l.n);
        dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 4096), (unsigned)pl.lists.size());
        tile_ptr_kernel<<<grid, 256, 0, h->stream>>>(dp.entries.p, dp.lists.p, dp.prev_boundary.p, n_node_tiles, (int)pl.lists.size(),
                                                     dp.tile_ptr.p, dp.tile_enc.p, dp.ent_x.p);
        CU(cudaGetLastError());
        dp.tile_ptr_ready = true;
        lap_place("tile tables");
    }
    if (accumulate) {
        CU(h->d_accS.ensure((size_t)pl.acc_total));
        CU(h->d_accC.ensure((size_t)pl.acc_total));
        CU(h->d_score.ensure((size_t)n));
        CU(h->d_counts.ensure(((size_t)n + 1) * NBINS));
        if (!node_tiles) {
            CU(h->d_diff_lo.ensure((size_t)n + 1));
            CU(h->d_diff_hi.ensure((size_t)n + 1));
            if (with_counts) CU(cudaMemsetAsync(h->d_counts.p, 0, ((size_t)n + 1) * NBINS * sizeof(int32_t), h->stream));
            CU(cudaMemsetAsync(h->d_diff_lo.p, 0, ((size_t)n + 1) * sizeof(unsigned long long), h->stream));
            CU(cudaMemsetAsync(h->d_diff_hi.p, 0, ((size_t)n + 1) * sizeof(unsigned long long), h->stream));
        }
    }
    bool acc_zeroed = false;
    auto zero_acc = [&]() -> cudaError_t {   // the per-(bucket, entry) accumulators, for the paths that add into them
        if (acc_zeroed) return cudaSuccess;
        acc_zeroed = true;
        cudaError_t e = cudaMemsetAsync(h->d_accS.p, 0, (size_t)pl.acc_total * sizeof(double), h->stream);
        if (e != cudaSuccess) return e;
        return cudaMemsetAsync(h->d_accC.p, 0, (size_t)pl.acc_total * sizeof(int32_t), h->stream);
    };

    PlaceParams pp = {};
    pp.lists = dp.entries.p;
    pp.list_desc = dp.lists.p;
    pp.chunk_start = dp.chunk_start.p;
    pp.buckets = dp.buckets.p;
    pp.tiles = dp.tiles.p;
    pp.n_tiles = (int32_t)pl.tiles.size();
    pp.n_nodes = n;
    pp.tile_counter = h->d_tile_counter.p;
    pp.start = h->d_rstart.p;
    pp.end = h->d_rend.p;
    pp.degree = h->d_rdegree.p;
    pp.rm_off = h->d_roff.p;
    pp.rm_pos = h->d_rpos.p;
    pp.rm_code = h->d_rcode.p;
    pp.perm = dp.perm.p;
    pp.mapped = h->has_mask ? h->d_mapped.p : nullptr;
    pp.max_pars = h->d_maxpars.p;
    pp.mult = h->d_mult.p;
    pp.accS = h->d_accS.p;
    pp.accC = h->d_accC.p;
    pp.accumulate = accumulate ? 1 : 0;
    pp.epp_cap = want_epp ? epp_cap : 0;
    pp.epp_capacity = want_epp ? (unsigned long long)epp_capacity : 0ull;
    pp.epp_total = h->d_epp_total.p;
    pp.epp_off = want_epp ? h->d_epp_off.p : nullptr;
    pp.epp_nodes = want_epp ? h->d_epp_nodes.p : nullptr;

    // score the distinct window-restricted haplotypes instead of scanning the Euler lists (state_place.cuh) when the
    // whole read set is placed with nothing mapped and no explicit EPP lists; WEPP_STATE_PLACE=0 keeps place_kernel
    const bool state_env = !(getenv("WEPP_STATE_PLACE") && atoi(getenv("WEPP_STATE_PLACE")) == 0);
    bool by_states = false;
    if (sub) {
        if (!dp.states_usable || !dp.delta_groups_usable) return fail(WEPP_E_STATE, "subset placement without the read set's window groups");
        by_states = true;
    } else if (state_env && accumulate && !want_epp && !h->has_mask && &dp == &h->full && pp.n_tiles > 0) {
        rc = build_states(h, dp);
        if (rc) return rc;
        by_states = dp.states_usable;
    }
    // ... and, when the reads share few distinct windows, by sparse corrections per read (delta_place.cuh);
    // WEPP_DELTA_PLACE=0 keeps state_place_kernel
    const bool delta_env = !(getenv("WEPP_DELTA_PLACE") && atoi(getenv("WEPP_DELTA_PLACE")) == 0);
    bool by_delta = false;
    if (sub) {
        by_delta = true;
    } else if (by_states && delta_env && dp.delta_usable) {
        rc = build_delta_groups(h, dp);
        if (rc) return rc;
        by_delta = dp.delta_groups_usable;
    }
    lap_place("states / window groups");
    if (!sub) {
        h->stats.place_path = by_delta ? 2 : (by_states ? 1 : 0);
        h->stats.n_states = by_states ? dp.n_states : 0;
        h->stats.n_window_groups = by_delta ? dp.n_groups : 0;
    }
    if (accumulate && !by_states) CU(zero_acc());   // place_kernel adds into the per-(bucket, entry) accumulators
    CU(cudaEventRecord(h->ev[0], h->stream));
    if (by_states) {
        CU(h->d_saccS.ensure((size_t)std::max<int64_t>(dp.sacc_total, 1)));
        CU(h->d_saccC.ensure((size_t)std::max<int64_t>(dp.sacc_total, 1)));
        CU(cudaMemsetAsync(h->d_saccS.p, 0, (size_t)dp.sacc_total * sizeof(double), h->stream));
        CU(cudaMemsetAsync(h->d_saccC.p, 0, (size_t)dp.sacc_total * sizeof(int32_t), h->stream));
    }
    if (by_delta) {
        CU(cudaMemsetAsync(dp.Gw.p, 0, (size_t)dp.n_groups * DP_BINS * sizeof(double), h->stream));
        CU(cudaMemsetAsync(dp.Gc.p, 0, (size_t)dp.n_groups * DP_BINS * sizeof(int32_t), h->stream));
        DeltaPlaceParams dq = {};
        dq.units = sub ? sub->units : dp.units.p; dq.n_units = sub ? sub->n_units : dp.n_units; dq.unit_counter = h->d_tile_counter.p;
        dq.groups = dp.groups.p; dq.base = dp.base.p; dq.whist = dp.whist.p; dq.post = dp.post.p;
        dq.state_first = dp.state_first.p; dq.sacc_off = dp.sacc_off.p; dq.rec = sub ? sub->rec : dp.rec.p; dq.mrec = dp.mrec.p;
        dq.max_pars = h->d_maxpars.p; dq.mult = h->d_mult.p; dq.saccS = h->d_saccS.p; dq.saccC = h->d_saccC.p;
        dq.Gw = dp.Gw.p; dq.Gc = dp.Gc.p; dq.gscratch = dp.gscratch.p; dq.gscratch_words = dp.gscratch_words;
        // shared memory: fixed areas + the widest list's base scores + 16 warps' nibble scratch, as far as it fits
        const int64_t s_max = dp.max_list_states;
        // (one CTA per SM: all of it — what the scratch areas leave is the warps' candidate queues)
        // two CTAs of 8 warps per SM where the widest list leaves room in half of the shared memory, else one of 16
        // (WEPP_DELTA_CTAS=1 forces the latter: tests)
        const bool two = delta_smem_need(s_max, DP_WARPS_SM / 2) <= (int64_t)delta_smem_bytes(h, 2) &&
                         !(getenv("WEPP_DELTA_CTAS") && atoi(getenv("WEPP_DELTA_CTAS")) == 1);
        const int ctas = two ? 2 : 1, warps = DP_WARPS_SM / ctas;
        const int smem = delta_smem_bytes(h, ctas);
        dq.smem_bytes = smem;
        dq.cand_cap = 1 << 20;
        if (getenv("WEPP_DELTA_CAND")) dq.cand_cap = std::max(0, atoi(getenv("WEPP_DELTA_CAND")));
        const int grid_dp = std::max(1, std::min((int)dq.n_units, h->n_sms * ctas));
        if (two) {
            CU(allow_max_smem(delta_place_kernel<DP_WARPS_SM / 2>, h));
            delta_place_kernel<DP_WARPS_SM / 2><<<grid_dp, warps * 32, smem, h->stream>>>(dq);
        } else {
            CU(allow_max_smem(delta_place_kernel<DP_WARPS_SM>, h));
            delta_place_kernel<DP_WARPS_SM><<<grid_dp, warps * 32, smem, h->stream>>>(dq);
        }
        CU(cudaGetLastError());
        dim3 fgrid((unsigned)std::min<int64_t>((s_max + 255) / 256, 1024), (unsigned)pl.buckets.size());
        delta_finalize_kernel<<<fgrid, 256, 0, h->stream>>>(dp.groups.p, dp.bucket_goff.p, dp.state_first.p, dp.buckets.p, dp.base.p,
                                                           dp.Gw.p, dp.Gc.p, dp.sacc_off.p, h->d_saccS.p, h->d_saccC.p);
        CU(cudaGetLastError());
        launches += 2;
    } else if (by_states) {
        StatePlaceParams sp = {};
        sp.state_ent = dp.state_ent.p; sp.state_eoff = dp.state_eoff.p; sp.state_first = dp.state_first.p;
        sp.sacc_off = dp.sacc_off.p; sp.list_desc = dp.lists.p; sp.buckets = dp.buckets.p; sp.tiles = dp.tiles.p;
        sp.start = h->d_rstart.p; sp.end = h->d_rend.p; sp.degree = h->d_rdegree.p; sp.rm_off = h->d_roff.p;
        sp.rm_pos = h->d_rpos.p; sp.rm_code = h->d_rcode.p; sp.perm = dp.perm.p;
        sp.max_pars = h->d_maxpars.p; sp.mult = h->d_mult.p; sp.saccS = h->d_saccS.p; sp.saccC = h->d_saccC.p;
        const int k = pl.reads_per_tile / 32;
        if (k == 8) rc = launch_state_place<8>(h, sp, pp.n_tiles, pl.max_width);
        else if (k == 4) rc = launch_state_place<4>(h, sp, pp.n_tiles, pl.max_width);
        else rc = launch_state_place<2>(h, sp, pp.n_tiles, pl.max_width);
        if (rc) return rc;
        ++launches;
    } else if (pp.n_tiles > 0) {
        const int k = pl.reads_per_tile / 32;
        if (k == 8) rc = launch_place_k<8>(h, pp, pl.max_width);
        else if (k == 4) rc = launch_place_k<4>(h, pp, pl.max_width);
        else rc = launch_place_k<2>(h, pp, pl.max_width);
        if (rc) return rc;
        ++launches;
    }
    CU(cudaEventRecord(h->ev[1], h->stream));
    h->exchange_timed = false;
    // read-sharded ranks: the accumulators of all ranks line up (one plan) — sum them, then finish per node
    if (accumulate && h->allreduce && h->shared_plan && &dp == &h->full && !sub) {
        int rc_x;
        if (by_states) {
            rc_x = h->allreduce(h->allreduce_user, h->d_saccS.p, dp.sacc_total, WEPP_DTYPE_F64, (void*)h->stream);
            if (!rc_x) rc_x = h->allreduce(h->allreduce_user, h->d_saccC.p, dp.sacc_total, WEPP_DTYPE_I32, (void*)h->stream);
        } else {
            rc_x = h->allreduce(h->allreduce_user, h->d_accS.p, pl.acc_total, WEPP_DTYPE_F64, (void*)h->stream);
            if (!rc_x) rc_x = h->allreduce(h->allreduce_user, h->d_accC.p, pl.acc_total, WEPP_DTYPE_I32, (void*)h->stream);
        }
        if (rc_x) return fail(WEPP_E_STATE, "the all-reduce hook failed");
        CU(cudaEventRecord(h->ev[4], h->stream));
        h->exchange_timed = true;
    }
    if (!sub) h->div_count_valid = false;
    if (accumulate && by_states && !node_tiles && !sub) {
        // the difference-array node path reads per-(bucket, entry) accumulators: spread the (merged) per-(bucket, state) ones
        int max_n = 0;
        for (const ListDesc& l : pl.lists) max_n = std::max(max_n, l.n);
        dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 4096), (unsigned)pl.buckets.size());
        state_scatter_kernel<<<grid, 256, 0, h->stream>>>(dp.lists.p, dp.buckets.p, dp.sid.p, dp.state_first.p, dp.sacc_off.p,
                                                          h->d_saccS.p, h->d_saccC.p, h->d_accS.p, h->d_accC.p);
        CU(cudaGetLastError());
        ++launches;
    }
    if (accumulate && node_tiles) {
        const bool by_state_acc = by_states;
        NodeTileParams np = {};
        np.list_desc = dp.lists.p; np.buckets = dp.buckets.p;
        np.n_buckets = (int32_t)pl.buckets.size(); np.n_lists = (int32_t)pl.lists.size(); np.n_nodes = n; np.n_tiles = n_node_tiles;
        np.ent_x = dp.ent_x.p; np.prev_boundary = dp.prev_boundary.p; np.tile_ptr = dp.tile_ptr.p; np.tile_enc = dp.tile_enc.p;
        np.accS = h->d_accS.p; np.accC = h->d_accC.p;
        np.sid = dp.sid.p; np.state_first = dp.state_first.p; np.sacc_off = dp.sacc_off.p;
        if (by_state_acc) {   // (weight, degree) pairs side by side: one sector per look-up in the tile kernel
            CU(h->d_sacc_packed.ensure((size_t)std::max<int64_t>(dp.sacc_total, 1)));
            sacc_pack_kernel<<<(unsigned)((dp.sacc_total + 255) / 256), 256, 0, h->stream>>>(h->d_saccS.p, h->d_saccC.p, dp.sacc_total,
                                                                                             h->d_sacc_packed.p);
            CU(cudaGetLastError());
            ++launches;
        }
        np.sacc = h->d_sacc_packed.p;
        // tile-major entry records for the full plan (kept while lists, buckets, states and the accumulator kind stay)
        const bool rec_env = !(getenv("WEPP_TILE_RECORDS") && atoi(getenv("WEPP_TILE_RECORDS")) == 0);
        if (rec_env && &dp == &h->full && (double)n_node_tiles * (double)pl.buckets.size() < 1.0e9) {
            const int mode = by_state_acc ? 1 : 0;
            bool same = dp.rec_ready && dp.rec_mode == mode && dp.rec_buckets.size() == pl.buckets.size();
            for (size_t b = 0; same && b < pl.buckets.size(); ++b)
                same = dp.rec_buckets[b].list == pl.buckets[b].list && dp.rec_buckets[b].bin == pl.buckets[b].bin &&
                       dp.rec_buckets[b].acc_off == pl.buckets[b].acc_off;
            if (!same) {
                const size_t n_tb = (size_t)n_node_tiles * pl.buckets.size() + 1;
                TmpBuf<uint32_t> tcount(h->stream);
                CU(tcount.ensure(n_tb));
                CU(dp.rec_off.ensure(n_tb));
                CU(dp.rec_x.ensure((size_t)pl.acc_total)); CU(dp.rec_cur.ensure((size_t)pl.acc_total)); CU(dp.rec_prv.ensure((size_t)pl.acc_total));
                tile_count_kernel<<<(unsigned)((n_tb + 255) / 256), 256, 0, h->stream>>>(dp.tile_ptr.p, dp.buckets.p, n_node_tiles,
                                                                                         (int)pl.lists.size(), (int)pl.buckets.size(), tcount.p);
                CU(cudaGetLastError());
                size_t tmp = 0;
                CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp, tcount.p, dp.rec_off.p, (int)n_tb, h->stream));
                CU(h->d_cub_tmp.ensure(tmp));
                CU(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp, tcount.p, dp.rec_off.p, (int)n_tb, h->stream));
                int max_n = 0;
                for (const ListDesc& l : pl.lists) max_n = std::max(max_n, l.n);
                dim3 rgrid((unsigned)std::min<int64_t>((max_n + 255) / 256, 4096), (unsigned)pl.buckets.size());
                if (mode == 1)
                    tile_records_kernel<true><<<rgrid, 256, 0, h->stream>>>(dp.ent_x.p, dp.prev_boundary.p, dp.lists.p, dp.buckets.p, dp.tile_ptr.p,
                                                                           dp.rec_off.p, dp.sid.p, dp.state_first.p, dp.sacc_off.p,
                                                                           (int)pl.lists.size(), (int)pl.buckets.size(), dp.rec_x.p, dp.rec_cur.p, dp.rec_prv.p);
                else
                    tile_records_kernel<false><<<rgrid, 256, 0, h->stream>>>(dp.ent_x.p, dp.prev_boundary.p, dp.lists.p, dp.buckets.p, dp.tile_ptr.p,
                                                                            dp.rec_off.p, nullptr, nullptr, nullptr,
                                                                            (int)pl.lists.size(), (int)pl.buckets.size(), dp.rec_x.p, dp.rec_cur.p, dp.rec_prv.p);
                CU(cudaGetLastError());
                dp.rec_ready = true;
                dp.rec_mode = mode;
                dp.rec_buckets = pl.buckets;
            }
            np.rec_off = dp.rec_off.p; np.rec_x = dp.rec_x.p; np.rec_cur = dp.rec_cur.p; np.rec_prv = dp.rec_prv.p;
        }
        np.mapped = h->has_mask ? h->d_mapped.p : nullptr;
        np.score = score_out ? score_out : h->d_score.p;
        np.counts = with_counts ? h->d_counts.p : nullptr;
        np.div_count = nullptr;
        if (with_counts && !score_out && h->peer_world == 0) {   // dist_divergence's bin count from the finished rows
            CU(h->d_div_count.ensure((size_t)n));
            np.div_count = h->d_div_count.p;
            // counts / true_counts > 0.5 % (initial_filter.cpp:214-231) as an integer threshold per bin: the smallest
            // count whose IEEE quotient exceeds it, found with the very division (monotone in the count)
            const double threshold = 0.5 / 100;
            for (int j = 0; j < NBINS; ++j) {
                const int32_t t = h->true_counts[j];
                int32_t lo = 0, hi = t;   // (double)hi / t = 1 > threshold when t > 0
                if (t <= 0) {
                    np.min_count.v[j] = INT32_MAX;
                    continue;
                }
                while (lo < hi) {
                    const int32_t mid = lo + (hi - lo) / 2;
                    if ((double)mid / (double)t > threshold) hi = mid;
                    else lo = mid + 1;
                }
                np.min_count.v[j] = lo;
            }
            h->div_count_valid = true;
        }
        if (by_state_acc) {
            CU(cudaFuncSetAttribute(node_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM));
            node_tile_kernel<true><<<(n_node_tiles + NT_SUPER - 1) / NT_SUPER, NT_THREADS, NT_SMEM, h->stream>>>(np);
        } else {
            CU(cudaFuncSetAttribute(node_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM));
            node_tile_kernel<false><<<(n_node_tiles + NT_SUPER - 1) / NT_SUPER, NT_THREADS, NT_SMEM, h->stream>>>(np);
        }
        CU(cudaGetLastError());
        ++launches;
    } else if (accumulate) {
        if (!pl.buckets.empty()) {
            int max_n = 0;
            for (const ListDesc& l : pl.lists) max_n = std::max(max_n, l.n);
            dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 4096), (unsigned)pl.buckets.size());
            if (sub)
                expand_states_score_kernel<<<grid, 256, 0, h->stream>>>(dp.entries.p, dp.lists.p, dp.buckets.p, dp.prev_boundary.p, dp.sid.p,
                                                                        dp.state_first.p, dp.sacc_off.p, h->d_saccS.p, h->d_diff_lo.p,
                                                                        h->d_diff_hi.p);
            else
                expand_kernel<<<grid, 256, 0, h->stream>>>(dp.entries.p, dp.lists.p, dp.buckets.p, dp.prev_boundary.p,
                                                           h->d_accS.p, h->d_accC.p, h->d_diff_lo.p, h->d_diff_hi.p,
                                                           with_counts ? h->d_counts.p : nullptr);
            CU(cudaGetLastError());
            ++launches;
        }
        const int n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
        CU(h->d_chunk128.ensure((size_t)n_chunks));
        score_chunk_sum_kernel<<<n_chunks, SCAN_THREADS, 0, h->stream>>>(h->d_diff_lo.p, h->d_diff_hi.p, n,
                                                                         h->d_chunk128.p);
        score_chunk_scan_kernel<<<1, SCAN_THREADS, 0, h->stream>>>(h->d_chunk128.p, n_chunks);
        score_apply_kernel<<<n_chunks, SCAN_THREADS, 0, h->stream>>>(h->d_diff_lo.p, h->d_diff_hi.p, n,
                                                                     h->d_chunk128.p,
                                                                     h->has_mask ? h->d_mapped.p : nullptr,
                                                                     score_out ? score_out : h->d_score.p);
        launches += 3;
        if (with_counts) {
            const int c_chunks = (n + CNT_CHUNK - 1) / CNT_CHUNK;
            CU(h->d_cchunk_tot.ensure((size_t)c_chunks * NBINS));
            CU(h->d_cchunk_off.ensure((size_t)c_chunks * NBINS));
            counts_chunk_sum_kernel<<<c_chunks, 64, 0, h->stream>>>(h->d_counts.p, n, h->d_cchunk_tot.p);
            counts_chunk_scan_kernel<<<NBINS, CSCAN_THREADS, 0, h->stream>>>(h->d_cchunk_tot.p, h->d_cchunk_off.p, c_chunks);
            counts_apply_kernel<<<c_chunks, 64, 0, h->stream>>>(h->d_counts.p, n, h->d_cchunk_off.p,
                                                                h->has_mask ? h->d_mapped.p : nullptr);
            launches += 3;
        }
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(h->ev[2], h->stream));
    if (sub) return WEPP_OK;   // the handle's results and statistics stay those of the last whole placement
    h->stats_pending = true;
    wepp_stats& st = h->stats;
    st.n_nodes = n;
    st.n_events = h->es.n_events;
    st.n_euler_entries = (int64_t)h->es.entries.size();
    st.n_reads = pl.n_reads;
    st.n_buckets = (int64_t)pl.buckets.size();
    st.n_lists = (int64_t)pl.lists.size();
    st.n_tiles = (int64_t)pl.tiles.size();
    st.list_entries_total = pl.list_entries_total;
    st.scanned_entries = pl.scanned_entries;
    st.scanned_read_entries = pl.scanned_read_entries;
    st.kernel_launches = launches;
    st.reads_per_tile = pl.reads_per_tile;
    st.stripe_width = h->es.stripe_width;
    // Algorithmic bytes of one place (DESIGN.md "Roofline"): two passes over each tile's Euler
    // list (16 B entries), the packed reads once, per-read results, the segment accumulators
    // written and read once, and the per-node difference arrays / results written and scanned.
    const int64_t rm = pl.n_read_muts;
    int64_t bytes = pl.scanned_entries * 16 * (accumulate ? 2 : 1);
    bytes += pl.n_reads * (12 + 8 + 8) + rm * 5;
    bytes += pl.n_reads * 8;
    if (accumulate) {
        bytes += pl.acc_total * 12 * 2;
        bytes += (int64_t)n * (16 + 8 + 200) * 2;
    }
    st.algorithmic_bytes = bytes;
    h->has_results = true;
    h->has_epp = want_epp;
    h->epp_capacity = want_epp ? epp_capacity : 0;
    return WEPP_OK;
}

}  // namespace

template <typename T>
static void copy_out(T* dst, const std::vector<T>& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
}

int set_arena_with(wepp_handle* h, int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                   const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, const EulerStripes* flattened);

extern "C" {

const char* wepp_last_error(void) { return g_err.c_str(); }
int wepp_abi_version(void) { return 1; }

int wepp_create(int device, wepp_handle** out) {
    if (!out) return fail(WEPP_E_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(WEPP_E_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") +
                                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return fail(WEPP_E_INVALID, "device ordinal out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(WEPP_E_CUDA, std::string("device ") + prop.name + " is not sm_100 class; this library is built for sm_100a only");
    wepp_handle* h = new wepp_handle();
    h->device = device;
    h->n_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // keep freed stream-ordered temporaries (TmpBuf) in the pool instead of handing them back at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
    }
    for (auto& ev : h->ev) CU(cudaEventCreate(&ev));
    *out = h;
    return WEPP_OK;
}

void wepp_destroy(wepp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int g = 0; g < MAX_PEERS; ++g)
        for (int b = 0; b < 4; ++b)
            if (h->peer_ptr[g][b] && g != h->peer_rank) cudaIpcCloseMemHandle(h->peer_ptr[g][b]);
    if (h->h_div_stage) cudaFreeHost(h->h_div_stage);
    h->d_div_count.release();
    h->d_stripes.release(); h->d_stripe_off.release(); h->d_mapped.release(); h->d_mapped_prefix.release();
    h->full.release(); h->sub.release();
    h->d_rstart.release(); h->d_rend.release(); h->d_rdegree.release(); h->d_rpos.release(); h->d_roff.release();
    h->d_rnuc.release(); h->d_rcode.release(); h->d_cell.release(); h->d_table.release(); h->d_bucket_of_cell.release();
    h->d_cursor.release(); h->d_true_counts.release(); h->d_key_status.release(); h->d_cell_pairs.release(); h->d_cell_assign.release();
    if (h->h_stage) cudaFreeHost(h->h_stage);
    h->d_accS.release(); h->d_accC.release(); h->d_maxpars.release(); h->d_mult.release(); h->d_score.release();
    h->d_counts.release(); h->d_divergence.release(); h->d_diff_lo.release(); h->d_diff_hi.release(); h->d_chunk128.release();
    h->d_cchunk_tot.release(); h->d_cchunk_off.release(); h->d_epp_off.release(); h->d_epp_nodes.release();
    h->d_epp_total.release(); h->d_tile_counter.release();
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int wepp_set_options(wepp_handle* h, int32_t stripe_width, int32_t reads_per_lane) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (h->has_arena) return fail(WEPP_E_STATE, "wepp_set_options must precede wepp_set_arena");
    if (stripe_width < 1 || stripe_width > 4096) return fail(WEPP_E_INVALID, "stripe_width must be in 1..4096");
    if (!(reads_per_lane == 0 || reads_per_lane == 2 || reads_per_lane == 4 || reads_per_lane == 8))
        return fail(WEPP_E_INVALID, "reads_per_lane must be 0, 2, 4 or 8");
    h->opt_q = stripe_width;
    h->opt_k = reads_per_lane;
    return WEPP_OK;
}

int wepp_set_arena(wepp_handle* h, int32_t n_nodes, const int32_t* parent, const int64_t* mut_off,
                   const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size) {
    return set_arena_with(h, n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size, nullptr);
}

}  // extern "C"

// wepp_set_arena; `flattened` = the tree's Euler stripes already built for this stripe width (the ranks of a group
// share one host flatten)
int set_arena_with(wepp_handle* h, int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                   const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, const EulerStripes* flattened) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!parent || !mut_off) return fail(WEPP_E_INVALID, "parent / mut_off is NULL");
    if (n_nodes >= 1 && mut_off[n_nodes] > 0 && (!mut_pos || !mut_ref || !mut_nuc))
        return fail(WEPP_E_INVALID, "mutation arrays are NULL");
    CU(cudaSetDevice(h->device));
    if (flattened) {
        h->es = *flattened;
    } else {
        std::string err = build_euler_stripes(n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size, h->opt_q, h->es);
        if (!err.empty()) return fail(WEPP_E_INVALID, err);
    }
    h->n_nodes = n_nodes;
    h->genome = genome_size;
    h->parent.assign(parent, parent + n_nodes);
    h->mut_off.assign(mut_off, mut_off + n_nodes + 1);
    const int64_t nm = mut_off[n_nodes];
    h->mut_pos.assign(mut_pos, mut_pos + nm);
    h->mut_ref.assign(mut_ref, mut_ref + nm);
    h->mut_nuc.assign(mut_nuc, mut_nuc + nm);
    CU(upload(h->d_stripes, h->es.entries, h->stream));
    CU(upload(h->d_stripe_off, h->es.stripe_off, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->has_arena = true;
    h->has_reads = false;
    h->has_mask = false;
    h->has_results = false;
    h->rank_tab_d = 0;
    h->full.states_ready = false;
    h->sub.states_ready = false;
    h->full.lists_built = false;
    h->sub.lists_built = false;
    if (h->peer_world > 0) {   // another tree may reallocate the exported per-node buffers: the peers' mappings must go
        for (int g = 0; g < MAX_PEERS; ++g)
            for (int b = 0; b < 4; ++b) {
                if (h->peer_ptr[g][b] && g != h->peer_rank) cudaIpcCloseMemHandle(h->peer_ptr[g][b]);
                h->peer_ptr[g][b] = nullptr;
            }
        h->peer_rank = -1;
        h->peer_world = 0;
        h->peer_merged = false;
    }
    h->tree_on_device = false;
    h->st_cache.clear();
    h->st_cache_pos.clear();
    h->st_cache_nuc.clear();
    return WEPP_OK;
}

extern "C" {

int wepp_set_reads(wepp_handle* h, int64_t n_reads, const int32_t* start, const int32_t* end,
                   const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_arena) return fail(WEPP_E_STATE, "wepp_set_arena must be called first");
    if (n_reads < 0) return fail(WEPP_E_INVALID, "n_reads < 0");
    if (n_reads > 0 && (!start || !end || !degree || !rm_off)) return fail(WEPP_E_INVALID, "read arrays are NULL");
    CU(cudaSetDevice(h->device));
    static const int64_t zero_off[1] = {0};
    if (n_reads == 0) rm_off = zero_off;
    const int64_t nm = rm_off[n_reads];
    if (nm < 0 || (nm > 0 && (!rm_pos || !rm_nuc))) return fail(WEPP_E_INVALID, "read mutation arrays are NULL / rm_off is negative");
    h->has_reads = false;
    h->has_results = false;
    h->host_reads = false;
    if (h->peer_world > 0) h->peer_stale = true;   // the ranks' true read counts were summed at wepp_peer_open
    cudaStream_t st = h->stream;
    const size_t n = (size_t)n_reads;
    // WEPP_TIMING=2: phase times on stderr (development aid; adds synchronisations)
    const bool timing = getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) == 2;
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[wepp_set_reads] %-34s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_mark).count());
        t_mark = t;
    };

    // ---- the reads go to the device once, as they are (caller order) ------------------------------
    CU(h->d_rstart.ensure(n)); CU(h->d_rend.ensure(n)); CU(h->d_rdegree.ensure(n)); CU(h->d_roff.ensure(n + 1));
    CU(h->d_rpos.ensure((size_t)nm)); CU(h->d_rnuc.ensure((size_t)nm)); CU(h->d_rcode.ensure((size_t)nm));
    if (n) {
        CU(cudaMemcpyAsync(h->d_rstart.p, start, n * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(h->d_rend.p, end, n * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(h->d_rdegree.p, degree, n * 4, cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpyAsync(h->d_roff.p, rm_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nm) {
        CU(cudaMemcpyAsync(h->d_rpos.p, rm_pos, (size_t)nm * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(h->d_rnuc.p, rm_nuc, (size_t)nm, cudaMemcpyHostToDevice, st));
    }

    lap("reads to the device");
    // ---- keying kernel: validation, allele classes, true read counts, reads per (window, bin) cell ----
    const int32_t q = h->es.stripe_width, n_stripes = h->es.n_stripes;
    const int32_t bin_size = h->genome / NBINS;
    const int64_t span_cap = std::min<int64_t>(n_stripes, MAX_WINDOW / q + 2);
    const int64_t bins_per_stripe = std::min<int64_t>(NBINS, q / std::max(bin_size, 1) + 2);
    const int64_t n_cells = (int64_t)n_stripes * span_cap * bins_per_stripe;
    const bool device_keys = n_reads > 0 && n_cells <= (int64_t)(1 << 22);   // else: host keying below
    constexpr int CELL_PAIR_CAP = 1 << 16;   // occupied histogram cells handed back as pairs (else the whole table)
    const size_t stage_bytes = 16 + NBINS * 8 + (device_keys ? std::max((size_t)n_cells * 4, (size_t)CELL_PAIR_CAP * 8) : 0);
    if (stage_bytes > h->h_stage_cap) {
        if (h->h_stage) cudaFreeHost(h->h_stage);
        h->h_stage = nullptr;
        h->h_stage_cap = 0;
        CU(cudaMallocHost(&h->h_stage, stage_bytes));
        h->h_stage_cap = stage_bytes;
    }
    int* st_status = reinterpret_cast<int*>(h->h_stage);
    unsigned long long* st_true = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(h->h_stage) + 16);
    int32_t* st_table = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(h->h_stage) + 16 + NBINS * 8);
    unsigned long long* st_pairs = reinterpret_cast<unsigned long long*>(st_table);   // same area: pairs, or the table on overflow
    CU(h->d_key_status.ensure(4)); CU(h->d_true_counts.ensure(NBINS)); CU(h->d_cell.ensure(n));
    CU(h->d_table.ensure((size_t)(device_keys ? n_cells : 1)));
    CU(cudaMemsetAsync(h->d_key_status.p, 0, 4 * sizeof(int), st));
    CU(cudaMemsetAsync(h->d_true_counts.p, 0, NBINS * 8, st));
    if (device_keys) CU(cudaMemsetAsync(h->d_table.p, 0, (size_t)n_cells * 4, st));
    if (n_reads > 0) {
        ReadKeyParams kp = {};
        kp.n_reads = n_reads; kp.n_muts = nm; kp.genome = h->genome; kp.q = q; kp.bin_size = bin_size;
        kp.span_cap = device_keys ? (int32_t)span_cap : 0;   // 0: every read reports "window exceeds the table"
        kp.bins_per_stripe = (int32_t)bins_per_stripe;
        kp.start = h->d_rstart.p; kp.end = h->d_rend.p; kp.degree = h->d_rdegree.p; kp.rm_off = h->d_roff.p;
        kp.rm_pos = h->d_rpos.p; kp.rm_nuc = h->d_rnuc.p; kp.rm_code = h->d_rcode.p; kp.cell = h->d_cell.p;
        kp.table = h->d_table.p; kp.true_counts = h->d_true_counts.p; kp.status = h->d_key_status.p;
        const int blocks = (int)std::min<int64_t>((n_reads + 255) / 256, (int64_t)h->n_sms * 8);
        read_keys_kernel<<<blocks, 256, 0, st>>>(kp);
        CU(cudaGetLastError());
    }
    // Read-sharded ranks (wepp_set_allreduce): the cell histogram and the true read counts are summed over the ranks,
    // so that every rank derives the SAME window lists, buckets and (from them) distinct states — the per-(bucket,
    // state) accumulators of a placement then line up across ranks and are what the ranks exchange.  The rank's own
    // histogram is kept for its bucket sizes.
    h->shared_plan = false;
    if (h->allreduce && device_keys) {
        CU(h->d_table_local.ensure((size_t)n_cells));
        CU(cudaMemcpyAsync(h->d_table_local.p, h->d_table.p, (size_t)n_cells * 4, cudaMemcpyDeviceToDevice, st));
        if (h->allreduce(h->allreduce_user, h->d_table.p, n_cells, WEPP_DTYPE_I32, (void*)st) != 0 ||
            h->allreduce(h->allreduce_user, h->d_true_counts.p, NBINS, WEPP_DTYPE_I64, (void*)st) != 0)
            return fail(WEPP_E_STATE, "the all-reduce hook failed");
        h->shared_plan = true;
    } else if (h->allreduce) {
        return fail(WEPP_E_INVALID, "read-sharded ranks need the device keying path (stripe geometry too fine for the cell table)");
    }
    if (n_reads > 0 || h->shared_plan) {
        if (device_keys) {
            CU(h->d_cell_pairs.ensure(CELL_PAIR_CAP));
            cells_compact_kernel<<<(unsigned)std::min<int64_t>((n_cells + 255) / 256, (int64_t)h->n_sms * 8), 256, 0, st>>>(
                h->d_table.p, n_cells, h->d_cell_pairs.p, CELL_PAIR_CAP, h->d_key_status.p);
            CU(cudaGetLastError());
        }
    }
    CU(cudaMemcpyAsync(st_status, h->d_key_status.p, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(st_true, h->d_true_counts.p, NBINS * 8, cudaMemcpyDeviceToHost, st));
    if (device_keys) CU(cudaMemcpyAsync(st_pairs, h->d_cell_pairs.p, (size_t)CELL_PAIR_CAP * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));   // also: the caller's buffers are consumed from here on
    lap("keying kernel + histogram out");
    if (st_status[0] != RP_OK) return fail(WEPP_E_INVALID, read_plan_error(st_status[0]));
    if (h->shared_plan && st_status[1] != 0)
        return fail(WEPP_E_INVALID, "read-sharded ranks need every read window inside the cell table's span");
    h->n_reads = n_reads;
    h->n_read_muts = nm;

    ReadPlan& pl = h->full.plan;
    bool host_perm = false;
    if (device_keys && st_status[1] == 0) {
        // ---- descriptors from the cell histogram; the reads are scattered into bucket order on the device ----
        pl = ReadPlan();
        pl.n_reads = n_reads;
        pl.n_read_muts = nm;
        // the occupied cells, ascending (= by first stripe, then span, then bin)
        std::vector<unsigned long long> occ;
        if (st_status[2] <= CELL_PAIR_CAP) {
            occ.assign(st_pairs, st_pairs + st_status[2]);
        } else {   // more occupied cells than pairs fit: fetch the table itself
            CU(cudaMemcpyAsync(st_table, h->d_table.p, (size_t)n_cells * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (int64_t c = 0; c < n_cells; ++c)
                if (st_table[c]) occ.push_back(((unsigned long long)c << 32) | (unsigned long long)(uint32_t)st_table[c]);
        }
        std::sort(occ.begin(), occ.end());
        std::vector<unsigned long long> assign;   // (cell << 32 | bucket)
        assign.reserve(occ.size());
        std::vector<int64_t> bucket_count;
        // bucket coarsening (merge_span_chain, host_prep.h): per (first stripe, count bin) the occupied spans are
        // grouped; a group's reads share the list of its widest span
        const char* no_merge = getenv("WEPP_NO_BUCKET_MERGE");
        const bool merge = !(no_merge && atoi(no_merge) != 0);
        const int64_t cells_per_qs = span_cap * bins_per_stripe;
        int32_t max_sp = 0;
        for (unsigned long long pr : occ) max_sp = std::max(max_sp, (int32_t)(((int64_t)(pr >> 32) % cells_per_qs) / bins_per_stripe));
        const int64_t tile = 32 * (int64_t)((h->opt_k == 2 || h->opt_k == 4 || h->opt_k == 8) ? h->opt_k
                                                                                              : reads_per_lane_for_width((max_sp + 1) * q));
        std::vector<std::vector<std::pair<int32_t, int64_t>>> chains((size_t)bins_per_stripe);
        std::vector<int32_t> tgt;
        std::vector<std::pair<int32_t, int32_t>> lists_here;   // (span, list) of the current first stripe
        for (size_t a0 = 0; a0 < occ.size();) {
            const int32_t qs = (int32_t)((int64_t)(occ[a0] >> 32) / cells_per_qs);
            const int32_t bin0 = std::min((qs * q) / std::max(bin_size, 1), NBINS - 1);
            for (auto& ch : chains) ch.clear();
            size_t a1 = a0;
            for (; a1 < occ.size() && (int32_t)((int64_t)(occ[a1] >> 32) / cells_per_qs) == qs; ++a1) {
                const int64_t in_qs = (int64_t)(occ[a1] >> 32) % cells_per_qs;
                chains[(size_t)(in_qs % bins_per_stripe)].emplace_back((int32_t)(in_qs / bins_per_stripe), (int64_t)(uint32_t)occ[a1]);
            }
            lists_here.clear();
            for (int32_t bb = 0; bb < bins_per_stripe; ++bb) {
                const auto& chain = chains[(size_t)bb];
                if (chain.empty()) continue;
                if (merge) merge_span_chain(chain, tile, tgt);
                else {
                    tgt.resize(chain.size());
                    for (size_t j = 0; j < chain.size(); ++j) tgt[j] = (int32_t)j;
                }
                int32_t cur_bucket = -1, cur_target = -1;
                for (size_t j = 0; j < chain.size(); ++j) {
                    if (tgt[j] != cur_target) {   // a new group: its bucket uses the list of the group's widest span
                        cur_target = tgt[j];
                        const int32_t sp_t = chain[(size_t)cur_target].first;
                        int32_t list = -1;
                        for (const auto& lh : lists_here)
                            if (lh.first == sp_t) list = lh.second;
                        if (list < 0) {
                            list = (int32_t)pl.lists.size();
                            lists_here.emplace_back(sp_t, list);
                            pl.lists.push_back(make_list_desc(h->es, qs, qs + sp_t));
                        }
                        cur_bucket = (int32_t)pl.buckets.size();
                        pl.buckets.push_back(BucketDesc{0, list, bin0 + bb});
                        bucket_count.push_back(0);
                    }
                    const int64_t cell = ((int64_t)qs * span_cap + chain[j].first) * bins_per_stripe + bb;
                    assign.push_back(((unsigned long long)cell << 32) | (unsigned long long)(uint32_t)cur_bucket);
                    bucket_count[(size_t)cur_bucket] += chain[j].second;
                }
            }
            a0 = a1;
        }
        CU(h->d_bucket_of_cell.ensure((size_t)n_cells));   // only the occupied cells are ever read
        CU(upload(h->d_cell_assign, assign, st));
        if (!assign.empty()) {
            cells_assign_kernel<<<(unsigned)((assign.size() + 255) / 256), 256, 0, st>>>(h->d_cell_assign.p, (int)assign.size(),
                                                                                        h->d_bucket_of_cell.p);
            CU(cudaGetLastError());
        }
        if (h->shared_plan) {   // the structure came from the summed histogram; the bucket sizes are this rank's own
            std::vector<int32_t> local(bucket_count.size(), 0);
            CU(h->d_bucket_count.ensure(bucket_count.size()));
            CU(cudaMemsetAsync(h->d_bucket_count.p, 0, bucket_count.size() * 4, st));
            if (!assign.empty()) {
                cells_local_count_kernel<<<(unsigned)((assign.size() + 255) / 256), 256, 0, st>>>(
                    h->d_cell_assign.p, (int)assign.size(), h->d_table_local.p, h->d_bucket_count.p);
                CU(cudaGetLastError());
            }
            CU(cudaMemcpyAsync(local.data(), h->d_bucket_count.p, local.size() * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (size_t b = 0; b < local.size(); ++b) bucket_count[b] = local[b];
        }
        std::vector<int64_t> first;
        std::string err = finish_read_plan(h->es, h->opt_k, bucket_count, pl, first);
        if (!err.empty()) return fail(WEPP_E_INVALID, err);
        lap("descriptors (host)");
        std::vector<unsigned long long> cursor(first.begin(), first.end());
        CU(upload(h->d_cursor, cursor, st));
        CU(h->full.perm.ensure(n));
        const int blocks = (int)std::min<int64_t>((n_reads + 255) / 256, (int64_t)h->n_sms * 8);
        read_scatter_kernel<<<blocks, 256, 0, st>>>(n_reads, h->d_cell.p, h->d_bucket_of_cell.p, h->d_cursor.p, h->full.perm.p);
        CU(cudaGetLastError());
    } else {
        // ---- host keying: stripe geometries whose cell table would be too large, windows wider than the
        //      table's span, and the empty read set ----
        std::string err = build_read_plan(h->es, h->genome, n_reads, start, end, degree, rm_off, rm_pos, rm_nuc, h->opt_k,
                                          nullptr, 0, pl);
        if (!err.empty()) return fail(WEPP_E_INVALID, err);
        host_perm = true;
    }
    for (int j = 0; j < NBINS; ++j) h->true_counts[j] = (int32_t)st_true[j];
    lap("scatter into bucket order");
    int rc = upload_plan(h, h->full, host_perm);
    if (rc) return rc;
    lap("descriptors up + Euler lists");
    h->has_reads = true;
    return WEPP_OK;
}

int wepp_set_allreduce(wepp_handle* h, wepp_allreduce_fn fn, void* user) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    h->allreduce = fn;
    h->allreduce_user = user;
    h->has_reads = false;   // the plan has to be derived again, with or without the other ranks
    h->has_results = false;
    h->shared_plan = false;
    return WEPP_OK;
}

int wepp_set_mapped(wepp_handle* h, const uint8_t* mapped) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_arena) return fail(WEPP_E_STATE, "wepp_set_arena must be called first");
    CU(cudaSetDevice(h->device));
    bool any = false;
    if (mapped)
        for (int32_t v = 0; v < h->n_nodes && !any; ++v) any = mapped[v] != 0;
    if (!any && !h->has_mask) return WEPP_OK;   // nothing mapped before, nothing now: the finalised lists stay valid
    h->has_mask = any;
    if (any) {
        std::vector<uint8_t> m(mapped, mapped + h->n_nodes);
        for (auto& x : m) x = x ? 1 : 0;
        std::vector<int32_t> pre((size_t)h->n_nodes + 1, 0);
        for (int32_t v = 0; v < h->n_nodes; ++v) pre[v + 1] = pre[v] + m[v];
        CU(upload(h->d_mapped, m, h->stream));
        CU(upload(h->d_mapped_prefix, pre, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    h->full.final_for_mask = false;
    h->sub.final_for_mask = false;
    return WEPP_OK;
}

int wepp_place(wepp_handle* h, int32_t epp_cap, int64_t epp_capacity) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_reads) return fail(WEPP_E_STATE, "wepp_set_reads must be called first");
    CU(cudaSetDevice(h->device));
    return run_place(h, h->full, true, epp_cap, epp_capacity);
}

int wepp_place_subset(wepp_handle* h, int64_t n_sel, const int64_t* read_idx, int32_t epp_cap, int64_t epp_capacity) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_reads) return fail(WEPP_E_STATE, "wepp_set_reads must be called first");
    if (n_sel < 0 || (n_sel > 0 && !read_idx)) return fail(WEPP_E_INVALID, "bad subset");
    CU(cudaSetDevice(h->device));
    int rc = ensure_host_reads(h);
    if (rc) return rc;
    std::string err = build_read_plan(h->es, h->genome, h->n_reads, h->r_start.data(), h->r_end.data(), h->r_degree.data(),
                                      h->r_off.data(), h->r_pos.data(), h->r_nuc.data(), h->opt_k, read_idx, n_sel,
                                      h->sub.plan);
    if (!err.empty()) return fail(WEPP_E_INVALID, err);
    rc = upload_plan(h, h->sub, true);
    if (rc) return rc;
    if (!h->has_results) {  // per-read arrays may not exist yet
        CU(h->d_maxpars.ensure((size_t)h->n_reads));
        CU(h->d_mult.ensure((size_t)h->n_reads));
    }
    return run_place(h, h->sub, false, epp_cap, epp_capacity);
}

int wepp_get_read_results(wepp_handle* h, int32_t* max_parsimony, int32_t* multiplicity) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_results) return fail(WEPP_E_STATE, "no placement results yet");
    CU(cudaSetDevice(h->device));
    if (max_parsimony) CU(cudaMemcpyAsync(max_parsimony, h->d_maxpars.p, (size_t)h->n_reads * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (multiplicity) CU(cudaMemcpyAsync(multiplicity, h->d_mult.p, (size_t)h->n_reads * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return WEPP_OK;
}

int wepp_get_node_results(wepp_handle* h, double* score, int32_t* counts) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_results || !h->d_score.p) return fail(WEPP_E_STATE, "no per-node results yet (call wepp_place)");
    CU(cudaSetDevice(h->device));
    if (score) CU(cudaMemcpyAsync(score, h->d_score.p, (size_t)h->n_nodes * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (counts) CU(cudaMemcpyAsync(counts, h->d_counts.p, (size_t)h->n_nodes * NBINS * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return WEPP_OK;
}

int wepp_get_node_summary(wepp_handle* h, double* score, double* dist_divergence) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_results || !h->d_score.p) return fail(WEPP_E_STATE, "no per-node results yet (call wepp_place)");
    CU(cudaSetDevice(h->device));
    const int n = h->n_nodes;
    const bool merged = h->peer_merged;   // wepp_peer_merge already evaluated both over all ranks
    const double* d_score_src = merged ? h->d_score_merged.p : h->d_score.p;
    if (dist_divergence && n > 0) {
        // dist_divergence = (bins over the threshold) / (bins with reads): the device counts the bins into one byte
        // per node, only those bytes cross PCIe (1 instead of 8 per node), and host threads do the division while
        // the score array is still in flight — the same IEEE division as on the device, bit for bit.
        CU(h->d_div_count.ensure((size_t)n));
        if ((size_t)n > h->h_div_cap) {
            if (h->h_div_stage) cudaFreeHost(h->h_div_stage);
            h->h_div_stage = nullptr;
            h->h_div_cap = 0;
            CU(cudaMallocHost(&h->h_div_stage, (size_t)n));
            h->h_div_cap = (size_t)n;
        }
        BinCounts tc;
        int active = 0;
        for (int j = 0; j < NBINS; ++j) {
            tc.v[j] = merged ? (int32_t)h->peer_true_counts[j] : h->true_counts[j];
            active += tc.v[j] != 0;
        }
        if (!merged && !h->div_count_valid) {
            divergence_count_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_counts.p, n, tc, 0.5 / 100, h->d_div_count.p);
            CU(cudaGetLastError());
        }
        CU(cudaMemcpyAsync(h->h_div_stage, h->d_div_count.p, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(h->ev[3], h->stream));
        if (score) CU(cudaMemcpyAsync(score, d_score_src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventSynchronize(h->ev[3]));
        double table[NBINS + 1];
        for (int k = 0; k <= NBINS; ++k) table[k] = (double)k / (double)active;
        const uint8_t* src = h->h_div_stage;
        const int n_thr = (int)std::max(1u, std::min(16u, std::min(std::thread::hardware_concurrency(), (unsigned)(n / 65536 + 1))));
        auto expand = [&](int t) {
            const int64_t a = (int64_t)n * t / n_thr, b = (int64_t)n * (t + 1) / n_thr;
            for (int64_t v = a; v < b; ++v) dist_divergence[v] = table[src[v]];
        };
        if (n_thr == 1) {
            expand(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < n_thr; ++t) pool.emplace_back(expand, t);
            for (auto& th : pool) th.join();
        }
        CU(cudaStreamSynchronize(h->stream));
        return WEPP_OK;
    }
    if (score) CU(cudaMemcpyAsync(score, d_score_src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return WEPP_OK;
}

int wepp_get_epp(wepp_handle* h, int64_t* epp_off, int32_t* epp_nodes, int64_t capacity, int64_t* n_epp) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_results || !h->has_epp) return fail(WEPP_E_STATE, "the last place call did not request EPP lists");
    if (!epp_off) return fail(WEPP_E_INVALID, "epp_off is NULL");
    CU(cudaSetDevice(h->device));
    const int64_t r = h->n_reads;
    std::vector<int64_t> off((size_t)r);
    std::vector<int32_t> mult((size_t)r);
    CU(cudaMemcpyAsync(off.data(), h->d_epp_off.p, (size_t)r * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(mult.data(), h->d_mult.p, (size_t)r * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    unsigned long long used = 0;
    CU(cudaMemcpyAsync(&used, h->d_epp_total.p, sizeof(used), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    used = std::min<unsigned long long>(used, (unsigned long long)h->epp_capacity);
    std::vector<int32_t> dev_nodes((size_t)used);
    if (used) CU(cudaMemcpyAsync(dev_nodes.data(), h->d_epp_nodes.p, (size_t)used * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // device lists are in allocation order; hand them back in read order
    int64_t total = 0;
    for (int64_t i = 0; i < r; ++i) {
        epp_off[i] = total;
        if (off[i] >= 0 && mult[i] > 0) total += mult[i];
    }
    epp_off[r] = total;
    if (n_epp) *n_epp = total;
    if (epp_nodes) {
        if (total > capacity) return fail(WEPP_E_CAPACITY, "epp_nodes capacity too small");
        for (int64_t i = 0; i < r; ++i)
            if (off[i] >= 0 && mult[i] > 0)
                std::memcpy(epp_nodes + epp_off[i], dev_nodes.data() + off[i], (size_t)mult[i] * sizeof(int32_t));
    }
    return WEPP_OK;
}

int wepp_cartesian_map(wepp_handle* h, int64_t n_reads, const int32_t* start, const int32_t* end,
                       const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc,
                       const uint8_t* mapped, int32_t* max_parsimony, int32_t* multiplicity, double* score,
                       int32_t* counts) {
    int rc = wepp_set_reads(h, n_reads, start, end, degree, rm_off, rm_pos, rm_nuc);
    if (rc) return rc;
    rc = wepp_set_mapped(h, mapped);
    if (rc) return rc;
    rc = wepp_place(h, 0, 0);
    if (rc) return rc;
    rc = wepp_get_read_results(h, max_parsimony, multiplicity);
    if (rc) return rc;
    return wepp_get_node_results(h, score, counts);
}

int wepp_device_buffer(wepp_handle* h, int32_t which, void** dev_ptr, int64_t* n_bytes) {
    if (!h || !dev_ptr) return fail(WEPP_E_INVALID, "NULL argument");
    if (!h->has_results) return fail(WEPP_E_STATE, "no placement results yet");
    void* p = nullptr;
    int64_t b = 0;
    switch (which) {
        case WEPP_BUF_SCORE: p = h->d_score.p; b = (int64_t)h->n_nodes * 8; break;
        case WEPP_BUF_COUNTS: p = h->d_counts.p; b = (int64_t)h->n_nodes * NBINS * 4; break;
        case WEPP_BUF_MAX_PARS: p = h->d_maxpars.p; b = h->n_reads * 4; break;
        case WEPP_BUF_MULT: p = h->d_mult.p; b = h->n_reads * 4; break;
        default: return fail(WEPP_E_INVALID, "unknown buffer id");
    }
    if (!p) return fail(WEPP_E_STATE, "buffer not allocated");
    *dev_ptr = p;
    if (n_bytes) *n_bytes = b;
    return WEPP_OK;
}

// ---- multi-GPU exchange step over NVLink peer memory ---------------------------------------------
namespace {
struct PeerBlob {   // what one rank publishes (WEPP_PEER_BLOB_BYTES)
    cudaIpcMemHandle_t mem[4];
    int64_t true_counts[NBINS];
    int64_t n_nodes;
};
static_assert(sizeof(PeerBlob) == WEPP_PEER_BLOB_BYTES, "WEPP_PEER_BLOB_BYTES out of date");

void peer_release(wepp_handle* h) {
    for (int g = 0; g < MAX_PEERS; ++g)
        for (int b = 0; b < 4; ++b) {
            if (h->peer_ptr[g][b] && g != h->peer_rank) cudaIpcCloseMemHandle(h->peer_ptr[g][b]);
            h->peer_ptr[g][b] = nullptr;
        }
    h->peer_rank = -1;
    h->peer_world = 0;
}
}  // namespace

int wepp_peer_export(wepp_handle* h, void* blob) {
    if (!h || !blob) return fail(WEPP_E_INVALID, "NULL argument");
    if (!h->has_arena || !h->has_reads) return fail(WEPP_E_STATE, "wepp_set_arena and wepp_set_reads must be called first");
    CU(cudaSetDevice(h->device));
    const size_t n = (size_t)h->n_nodes;
    CU(h->d_score.ensure(n));
    CU(h->d_counts.ensure((n + 1) * NBINS));
    CU(h->d_score_merged.ensure(n));
    CU(h->d_div_count.ensure(n));
    PeerBlob pb = {};
    CU(cudaIpcGetMemHandle(&pb.mem[0], h->d_score.p));
    CU(cudaIpcGetMemHandle(&pb.mem[1], h->d_counts.p));
    CU(cudaIpcGetMemHandle(&pb.mem[2], h->d_score_merged.p));
    CU(cudaIpcGetMemHandle(&pb.mem[3], h->d_div_count.p));
    for (int j = 0; j < NBINS; ++j) pb.true_counts[j] = h->true_counts[j];
    pb.n_nodes = h->n_nodes;
    std::memcpy(blob, &pb, sizeof(pb));
    return WEPP_OK;
}

int wepp_peer_open(wepp_handle* h, int32_t rank, int32_t world, const void* blobs) {
    if (!h || !blobs) return fail(WEPP_E_INVALID, "NULL argument");
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) return fail(WEPP_E_INVALID, "bad rank / world (at most 8 ranks)");
    if (!h->d_score.p || !h->d_counts.p || !h->d_score_merged.p || !h->d_div_count.p)
        return fail(WEPP_E_STATE, "wepp_peer_export must be called first");
    CU(cudaSetDevice(h->device));
    peer_release(h);
    const PeerBlob* pb = static_cast<const PeerBlob*>(blobs);
    for (int j = 0; j < NBINS; ++j) h->peer_true_counts[j] = 0;
    h->peer_rank = rank;
    h->peer_world = world;
    h->peer_stale = false;
    for (int g = 0; g < world; ++g) {
        if (pb[g].n_nodes != h->n_nodes) {
            peer_release(h);
            return fail(WEPP_E_INVALID, "ranks hold different trees");
        }
        for (int j = 0; j < NBINS; ++j) h->peer_true_counts[j] += pb[g].true_counts[j];
        if (g == rank) {
            h->peer_ptr[g][0] = h->d_score.p;
            h->peer_ptr[g][1] = h->d_counts.p;
            h->peer_ptr[g][2] = h->d_score_merged.p;
            h->peer_ptr[g][3] = h->d_div_count.p;
            continue;
        }
        for (int b = 0; b < 4; ++b) {
            cudaError_t e = cudaIpcOpenMemHandle(&h->peer_ptr[g][b], pb[g].mem[b], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                h->peer_ptr[g][b] = nullptr;
                peer_release(h);
                return fail(WEPP_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
        }
    }
    return WEPP_OK;
}

int wepp_peer_merge(wepp_handle* h) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (h->peer_world < 1) return fail(WEPP_E_STATE, "wepp_peer_open must be called first");
    if (h->peer_stale)
        return fail(WEPP_E_STATE, "the reads changed since wepp_peer_open (the ranks' true read counts are summed there): "
                                  "wepp_peer_export / wepp_peer_open again after wepp_set_reads");
    if (!h->has_results) return fail(WEPP_E_STATE, "no per-node results yet (call wepp_place)");
    CU(cudaSetDevice(h->device));
    PeerMergeParams p = {};
    const int64_t n = h->n_nodes;
    const int w = h->peer_world, r = h->peer_rank;
    auto cut = [&](int g) { return g >= w ? n : (n * g / w) / PM_NODES * PM_NODES; };
    p.world = w;
    p.rank = r;
    p.lo = (int32_t)cut(r);
    p.hi = (int32_t)cut(r + 1);
    for (int g = 0; g < w; ++g) {
        p.score_in[g] = static_cast<const double*>(h->peer_ptr[g][0]);
        p.counts_in[g] = static_cast<const int32_t*>(h->peer_ptr[g][1]);
        p.score_out[g] = static_cast<double*>(h->peer_ptr[g][2]);
        p.div_out[g] = static_cast<uint8_t*>(h->peer_ptr[g][3]);
    }
    p.counts_own = h->d_counts.p;
    p.bins_active = 0;
    for (int j = 0; j < NBINS; ++j) {
        if (h->peer_true_counts[j] > 0x7FFFFFFFll) return fail(WEPP_E_INVALID, "degree-weighted read count of a bin exceeds int32");
        p.true_counts.v[j] = (int32_t)h->peer_true_counts[j];
        p.bins_active += p.true_counts.v[j] != 0;
    }
    p.threshold = 0.5 / 100;
    if (p.hi > p.lo) {
        const int blocks = (int)std::min<int64_t>(((int64_t)p.hi - p.lo + PM_NODES - 1) / PM_NODES, (int64_t)h->n_sms * 8);
        CU(cudaEventRecord(h->ev[4], h->stream));
        peer_merge_kernel<<<blocks, PM_THREADS, 0, h->stream>>>(p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(h->ev[5], h->stream));
    }
    h->peer_merged = true;
    return WEPP_OK;
}

int wepp_peer_close(wepp_handle* h) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) != 0 && h->peer_merged) {   // development aid
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]) == cudaSuccess)
            fprintf(stderr, "[wepp timing] rank %d: last peer_merge_kernel %.3f ms\n", h->peer_rank, ms);
    }
    peer_release(h);
    h->peer_merged = false;
    return WEPP_OK;
}

int wepp_sync(wepp_handle* h) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return WEPP_OK;
}

int wepp_set_stream(wepp_handle* h, void* cuda_stream) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return WEPP_OK;
}

int wepp_get_stats(wepp_handle* h, wepp_stats* out) {
    if (!h || !out) return fail(WEPP_E_INVALID, "NULL argument");
    if (h->stats_pending) {
        CU(cudaSetDevice(h->device));
        CU(cudaEventSynchronize(h->ev[2]));
        float ms_scan = 0, ms_node = 0;
        CU(cudaEventElapsedTime(&ms_scan, h->ev[0], h->ev[1]));
        CU(cudaEventElapsedTime(&ms_node, h->ev[1], h->ev[2]));
        h->stats.ms_exchange = 0.f;
        if (h->exchange_timed) {   // read-sharded ranks: the accumulator all-reduce sits between placement and node kernels
            float ms_x = 0;
            CU(cudaEventElapsedTime(&ms_x, h->ev[1], h->ev[4]));
            h->stats.ms_exchange = ms_x;
            ms_node -= ms_x;
        }
        h->stats.ms_scan_kernel = ms_scan;
        h->stats.ms_node_kernels = ms_node;
        h->stats.ms_place_total = ms_scan + ms_node;
        h->stats_pending = false;
    }
    *out = h->stats;
    return WEPP_OK;
}

// ---- arena builder (host) ---------------------------------------------------------------------
struct wepp_arena {
    wepp::ArenaHost a;
};

int wepp_arena_build(int32_t n_mat_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                     const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int32_t n_masked,
                     const int32_t* masked, int64_t n_reads, const int32_t* start, const int32_t* end,
                     const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc, wepp_arena** out) {
    if (!out || !parent || !mut_off || !rm_off) return fail(WEPP_E_INVALID, "NULL argument");
    wepp_arena* a = new wepp_arena();
    std::string err = build_arena(n_mat_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size, n_masked, masked,
                                  n_reads, start, end, rm_off, rm_pos, rm_nuc, a->a);
    if (!err.empty()) {
        delete a;
        return fail(WEPP_E_INVALID, err);
    }
    *out = a;
    return WEPP_OK;
}

void wepp_arena_free(wepp_arena* a) { delete a; }

int wepp_arena_dims(const wepp_arena* a, int32_t* n_nodes, int64_t* n_muts, int64_t* n_read_muts, int64_t* n_mapped) {
    if (!a) return fail(WEPP_E_INVALID, "arena is NULL");
    if (n_nodes) *n_nodes = (int32_t)a->a.parent.size();
    if (n_muts) *n_muts = (int64_t)a->a.mut_pos.size();
    if (n_read_muts) *n_read_muts = (int64_t)a->a.rm_pos.size();
    if (n_mapped) *n_mapped = (int64_t)a->a.map_nodes.size();
    return WEPP_OK;
}

int wepp_arena_get(const wepp_arena* a, int32_t* parent, int32_t* source, int32_t* leaf_count, int64_t* mut_off,
                   int32_t* mut_pos, uint8_t* mut_ref, uint8_t* mut_nuc, int64_t* map_off, int32_t* map_nodes) {
    if (!a) return fail(WEPP_E_INVALID, "arena is NULL");
    copy_out(parent, a->a.parent); copy_out(source, a->a.source); copy_out(leaf_count, a->a.leaf_count);
    copy_out(mut_off, a->a.mut_off); copy_out(mut_pos, a->a.mut_pos); copy_out(mut_ref, a->a.mut_ref);
    copy_out(mut_nuc, a->a.mut_nuc); copy_out(map_off, a->a.map_off); copy_out(map_nodes, a->a.map_nodes);
    return WEPP_OK;
}

int wepp_arena_get_reads(const wepp_arena* a, int64_t* rm_off, int32_t* rm_pos, uint8_t* rm_nuc) {
    if (!a) return fail(WEPP_E_INVALID, "arena is NULL");
    copy_out(rm_off, a->a.rm_off); copy_out(rm_pos, a->a.rm_pos); copy_out(rm_nuc, a->a.rm_nuc);
    return WEPP_OK;
}

int wepp_set_arena_from(wepp_handle* h, const wepp_arena* a) {
    if (!h || !a) return fail(WEPP_E_INVALID, "NULL argument");
    const wepp::ArenaHost& x = a->a;
    return wepp_set_arena(h, (int32_t)x.parent.size(), x.parent.data(), x.mut_off.data(), x.mut_pos.data(),
                          x.mut_ref.data(), x.mut_nuc.data(), x.genome_size);
}

// Host-only view of the Euler stripes (no GPU needed): used by the CPU test-suite to check the
// signed-delta tables against the oracle.  entries = 4 x uint32 per entry (Entry master form).
int64_t wepp_host_euler_stripes(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                                const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size,
                                int32_t stripe_width, uint32_t* entries, int64_t capacity, int64_t* stripe_off,
                                int32_t stripe_off_len) {
    EulerStripes es;
    std::string err = build_euler_stripes(n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size, stripe_width, es);
    if (!err.empty()) return fail(WEPP_E_INVALID, err);
    const int64_t n = (int64_t)es.entries.size();
    if (entries) {
        if (capacity < n) return fail(WEPP_E_CAPACITY, "entries capacity too small");
        std::memcpy(entries, es.entries.data(), (size_t)n * sizeof(Entry));
    }
    if (stripe_off) {
        if (stripe_off_len < es.n_stripes + 1) return fail(WEPP_E_CAPACITY, "stripe_off too small");
        std::memcpy(stripe_off, es.stripe_off.data(), ((size_t)es.n_stripes + 1) * sizeof(int64_t));
    }
    return n;
}

// Host-only view of the read plan: bucket-sorted permutation and, per sorted read, the list's
// stripe range; returns the number of tiles (or a negative error).
int64_t wepp_host_read_plan(int32_t genome_size, int32_t stripe_width, int32_t reads_per_lane, int64_t n_reads,
                            const int32_t* start, const int32_t* end, const int32_t* degree, const int64_t* rm_off,
                            const int32_t* rm_pos, const uint8_t* rm_nuc, int64_t* perm, int32_t* qs, int32_t* qe,
                            int32_t* bin, int32_t* reads_per_tile) {
    EulerStripes es;
    es.stripe_width = stripe_width;
    es.n_stripes = genome_size / stripe_width + 1;
    es.stripe_off.assign((size_t)es.n_stripes + 1, 0);
    ReadPlan pl;
    std::string err = build_read_plan(es, genome_size, n_reads, start, end, degree, rm_off, rm_pos, rm_nuc,
                                      reads_per_lane, nullptr, 0, pl);
    if (!err.empty()) return fail(WEPP_E_INVALID, err);
    for (const TileDesc& t : pl.tiles)
        for (int64_t i = t.first; i < t.first + t.count; ++i) {
            const BucketDesc& b = pl.buckets[t.bucket];
            if (perm) perm[i] = pl.perm[i];
            if (qs) qs[i] = pl.lists[b.list].qs;
            if (qe) qe[i] = pl.lists[b.list].qe;
            if (bin) bin[i] = b.bin;
        }
    if (reads_per_tile) *reads_per_tile = pl.reads_per_tile;
    return (int64_t)pl.tiles.size();
}

}  // extern "C"

namespace {

template <int K>
int launch_rescore_tiles(wepp_handle* h, const RescoreTileParams& p, int n_tiles, int width, int mode) {
    const size_t smem = (size_t)RtLayout<K>::CODES + (((size_t)width * 32 * K + 15) & ~(size_t)15);
    if (smem > h->smem_optin || width > MAX_WINDOW)
        return fail(WEPP_E_INVALID, "read window too wide for shared memory (" + std::to_string(width) + " bases)");
    if (mode == 0) {
        CU(allow_max_smem(rescore_tile_kernel<K, 0>, h));
        rescore_tile_kernel<K, 0><<<n_tiles, PLACE_WARPS * 32, smem, h->stream>>>(p);
    } else {
        CU(allow_max_smem(rescore_tile_kernel<K, 1>, h));
        rescore_tile_kernel<K, 1><<<n_tiles, PLACE_WARPS * 32, smem, h->stream>>>(p);
    }
    CU(cudaGetLastError());
    return WEPP_OK;
}

int launch_rescore_tiles_k(wepp_handle* h, const RescoreTileParams& p, int n_tiles, int width, int k, int mode) {
    if (k == 8) return launch_rescore_tiles<8>(h, p, n_tiles, width, mode);
    if (k == 4) return launch_rescore_tiles<4>(h, p, n_tiles, width, mode);
    return launch_rescore_tiles<2>(h, p, n_tiles, width, mode);
}

// K4 over the resident reads: the candidates' in-window stack mutations become per-list entry runs
// (count -> scan -> fill), then one tile kernel for min / argmin count and one for the argmin lists.
int rescore_resident(wepp_handle* h, int32_t n_cand, const int32_t* cand_nodes, int32_t* min_dist, int32_t* dist,
                     int64_t* am_off, int32_t* am_idx, int64_t am_capacity) {
    wepp_handle::DevPlan& dp = h->full;
    const ReadPlan& pl = dp.plan;
    const int64_t R = h->n_reads;
    if (R == 0 || pl.tiles.empty()) {
        if (am_off) am_off[0] = 0;
        return WEPP_OK;
    }
    std::string err;
    static const bool timing = getenv("WEPP_RESCORE_TIMING") && atoi(getenv("WEPP_RESCORE_TIMING")) != 0;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {   // development aid: phase times on stderr (serialises the stream)
        if (!timing) return;
        cudaStreamSynchronize(h->stream);
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[wepp_rescore] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    std::vector<int64_t> st_off((size_t)n_cand + 1, 0);
    std::vector<int32_t> st_pos;
    std::vector<uint8_t> st_nuc;
    {
        if (h->st_cache_pos.size() > ((size_t)64 << 20)) {   // bound the cache
            h->st_cache.clear();
            h->st_cache_pos.clear();
            h->st_cache_nuc.clear();
        }
        std::vector<int32_t> missing;
        for (int32_t c = 0; c < n_cand; ++c) {
            if (cand_nodes[c] < 0 || cand_nodes[c] >= h->n_nodes) return fail(WEPP_E_INVALID, "candidate node index out of range");
            if (h->st_cache.emplace(cand_nodes[c], std::make_pair((int64_t)-1, 0)).second) missing.push_back(cand_nodes[c]);
        }
        if (!missing.empty()) {
            std::vector<int64_t> m_off;
            std::vector<int32_t> m_pos;
            std::vector<uint8_t> m_nuc;
            bool on_device = false;
            const int bitmap_words = (h->genome + 1 + 31) / 32;
            const size_t sb_smem = (size_t)SB_WARPS * ((size_t)bitmap_words + SB_CAP) * 4;
            static const bool host_stacks = getenv("WEPP_HOST_STACKS") && atoi(getenv("WEPP_HOST_STACKS")) != 0;
            if (!host_stacks && missing.size() >= 16 && sb_smem <= 48 * 1024) {
                // root-path walks on the device (cand_stack_kernel); the tree arrays go up on first use
                const int nmiss = (int)missing.size();
                if (!h->tree_on_device) {
                    CU(upload(h->d_parent, h->parent, h->stream));
                    CU(upload(h->d_mut_off, h->mut_off, h->stream));
                    CU(upload(h->d_mut_pos, h->mut_pos, h->stream));
                    CU(upload(h->d_mut_ref, h->mut_ref, h->stream));
                    CU(upload(h->d_mut_nuc, h->mut_nuc, h->stream));
                    h->tree_on_device = true;
                }
                CU(upload(h->d_sb_nodes, missing, h->stream));
                CU(h->d_sb_rows.ensure((size_t)nmiss * SB_CAP));
                CU(h->d_sb_count.ensure((size_t)nmiss));
                CU(h->d_sb_cnt64.ensure((size_t)nmiss + 1));
                CU(h->d_sb_off.ensure((size_t)nmiss + 1));
                StackBuildParams sp = {};
                sp.parent = h->d_parent.p; sp.mut_off = h->d_mut_off.p; sp.mut_pos = h->d_mut_pos.p;
                sp.mut_ref = h->d_mut_ref.p; sp.mut_nuc = h->d_mut_nuc.p; sp.nodes = h->d_sb_nodes.p;
                sp.n = nmiss; sp.bitmap_words = bitmap_words; sp.out = h->d_sb_rows.p; sp.count = h->d_sb_count.p;
                cand_stack_kernel<<<(nmiss + SB_WARPS - 1) / SB_WARPS, SB_WARPS * 32, sb_smem, h->stream>>>(sp);
                CU(cudaGetLastError());
                clamp_counts_kernel<<<(nmiss + 1 + 255) / 256, 256, 0, h->stream>>>(h->d_sb_count.p, nmiss, h->d_sb_cnt64.p);
                CU(cudaGetLastError());
                size_t tmp_bytes = 0;
                CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_sb_cnt64.p, h->d_sb_off.p, nmiss + 1, h->stream));
                CU(h->d_cub_tmp.ensure(tmp_bytes));
                CU(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp_bytes, h->d_sb_cnt64.p, h->d_sb_off.p, nmiss + 1, h->stream));
                std::vector<int32_t> cnt((size_t)nmiss);
                m_off.resize((size_t)nmiss + 1);
                CU(cudaMemcpyAsync(cnt.data(), h->d_sb_count.p, (size_t)nmiss * 4, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaMemcpyAsync(m_off.data(), h->d_sb_off.p, ((size_t)nmiss + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                on_device = std::find(cnt.begin(), cnt.end(), -1) == cnt.end();   // a stack over SB_CAP: the host builds them all
                if (on_device) {
                    const size_t total = (size_t)m_off[(size_t)nmiss];
                    m_pos.resize(total);
                    m_nuc.resize(total);
                    if (total) {
                        CU(h->d_sb_pos.ensure(total));
                        CU(h->d_sb_nuc.ensure(total));
                        cand_stack_compact_kernel<<<nmiss, 64, 0, h->stream>>>(h->d_sb_rows.p, h->d_sb_count.p, h->d_sb_off.p, nmiss,
                                                                              h->d_sb_pos.p, h->d_sb_nuc.p);
                        CU(cudaGetLastError());
                        CU(cudaMemcpyAsync(m_pos.data(), h->d_sb_pos.p, total * 4, cudaMemcpyDeviceToHost, h->stream));
                        CU(cudaMemcpyAsync(m_nuc.data(), h->d_sb_nuc.p, total, cudaMemcpyDeviceToHost, h->stream));
                        CU(cudaStreamSynchronize(h->stream));
                    }
                }
            }
            if (!on_device &&
                !build_candidate_stacks(h->n_nodes, h->genome, h->parent.data(), h->mut_off.data(), h->mut_pos.data(),
                                        h->mut_ref.data(), h->mut_nuc.data(), (int32_t)missing.size(), missing.data(), m_off,
                                        m_pos, m_nuc, err))
                return fail(WEPP_E_INVALID, err);
            const int64_t base = (int64_t)h->st_cache_pos.size();
            h->st_cache_pos.insert(h->st_cache_pos.end(), m_pos.begin(), m_pos.end());
            h->st_cache_nuc.insert(h->st_cache_nuc.end(), m_nuc.begin(), m_nuc.end());
            for (size_t i = 0; i < missing.size(); ++i)
                h->st_cache[missing[i]] = std::make_pair(base + m_off[i], (int32_t)(m_off[i + 1] - m_off[i]));
        }
        std::vector<std::pair<int64_t, int32_t>> where((size_t)n_cand);
        for (int32_t c = 0; c < n_cand; ++c) {
            where[(size_t)c] = h->st_cache[cand_nodes[c]];
            st_off[(size_t)c + 1] = st_off[(size_t)c] + where[(size_t)c].second;
        }
        st_pos.resize((size_t)st_off[(size_t)n_cand]);
        st_nuc.resize((size_t)st_off[(size_t)n_cand]);
        for (int32_t c = 0; c < n_cand; ++c) {
            const auto& w = where[(size_t)c];
            std::copy_n(h->st_cache_pos.begin() + w.first, w.second, st_pos.begin() + st_off[(size_t)c]);
            std::copy_n(h->st_cache_nuc.begin() + w.first, w.second, st_nuc.begin() + st_off[(size_t)c]);
        }
    }
    lap("candidate stacks (host)");
    const int n_lists = (int)pl.lists.size();
    const int64_t n_lc = (int64_t)n_lists * n_cand;
    if (n_lc + 1 > 0x7FFFFFFFll) return fail(WEPP_E_INVALID, "too many (window list, candidate) pairs: split the candidate set");
    // capacity of the entry buffer: one entry per (list, candidate) + every stack mutation once per list covering it
    int64_t cap = n_lc;
    {
        std::vector<int32_t> cover((size_t)h->genome + 2, 0);
        for (const ListDesc& l : pl.lists) {
            cover[(size_t)std::min(l.b0, h->genome + 1)] += 1;
            cover[(size_t)std::min(l.b0 + l.width, h->genome + 1)] -= 1;
        }
        for (size_t i = 1; i < cover.size(); ++i) cover[i] += cover[i - 1];
        for (int32_t pos : st_pos) cap += cover[(size_t)pos];
    }
    CU(upload(h->d_st_off, st_off, h->stream));
    CU(upload(h->d_st_pos, st_pos, h->stream));
    CU(upload(h->d_st_nuc, st_nuc, h->stream));
    CU(h->d_ccnt.ensure((size_t)n_lc + 1));
    CU(h->d_coff.ensure((size_t)n_lc + 1));
    CU(h->d_cent.ensure((size_t)cap));
    CU(h->d_rs_min.ensure((size_t)R));
    CU(h->d_rs_nbest.ensure((size_t)R));
    CU(h->d_rs_before.ensure((size_t)R * PLACE_WARPS));
    if (dist) CU(h->d_rs_dist.ensure((size_t)R * (size_t)n_cand));
    const unsigned blocks = (unsigned)((n_lc + 1 + 255) / 256);
    // min / count only: candidates with nothing in a window are evaluated in bulk and get no entry there; the dense
    // matrix and the argmin lists need every candidate's own evaluation
    const int min_one = (dist || (am_off && am_idx)) ? 1 : 0;
    cand_count_kernel<<<blocks, 256, 0, h->stream>>>(dp.lists.p, n_lists, n_cand, h->d_st_off.p, h->d_st_pos.p, h->d_ccnt.p, min_one);
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->d_ccnt.p, h->d_coff.p, (int)(n_lc + 1), h->stream));
    CU(h->d_cub_tmp.ensure(tmp_bytes));
    CU(cub::DeviceScan::ExclusiveSum(h->d_cub_tmp.p, tmp_bytes, h->d_ccnt.p, h->d_coff.p, (int)(n_lc + 1), h->stream));
    cand_fill_kernel<<<blocks, 256, 0, h->stream>>>(dp.lists.p, n_lists, n_cand, h->d_st_off.p, h->d_st_pos.p, h->d_st_nuc.p,
                                                    h->d_coff.p, h->d_cent.p);
    CU(cudaGetLastError());
    lap("upload + entry lists");

    RescoreTileParams p = {};
    p.cent = h->d_cent.p;
    p.coff = h->d_coff.p;
    p.n_cand = n_cand;
    p.list_desc = dp.lists.p;
    p.buckets = dp.buckets.p;
    p.tiles = dp.tiles.p;
    p.start = h->d_rstart.p;
    p.end = h->d_rend.p;
    p.rm_off = h->d_roff.p;
    p.rm_pos = h->d_rpos.p;
    p.rm_code = h->d_rcode.p;
    p.perm = dp.perm.p;
    p.min_dist = h->d_rs_min.p;
    p.n_argmin = h->d_rs_nbest.p;
    p.before = h->d_rs_before.p;
    p.dist = dist ? h->d_rs_dist.p : nullptr;
    const int n_tiles = (int)pl.tiles.size(), k = pl.reads_per_tile / 32;
    int rc = launch_rescore_tiles_k(h, p, n_tiles, pl.max_width, k, 0);
    if (rc) return rc;
    lap("tile kernel (min / count)");
    CU(cudaMemcpyAsync(min_dist, h->d_rs_min.p, (size_t)R * 4, cudaMemcpyDeviceToHost, h->stream));
    if (dist) CU(cudaMemcpyAsync(dist, h->d_rs_dist.p, (size_t)R * (size_t)n_cand * 4, cudaMemcpyDeviceToHost, h->stream));
    if (am_off) {
        std::vector<int32_t> nbest((size_t)R);
        CU(cudaMemcpyAsync(nbest.data(), h->d_rs_nbest.p, (size_t)R * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        int64_t tot = 0;
        for (int64_t r = 0; r < R; ++r) {
            am_off[r] = tot;
            tot += nbest[(size_t)r];
        }
        am_off[R] = tot;
        if (am_idx) {
            if (tot > am_capacity) return fail(WEPP_E_CAPACITY, "am_idx capacity too small");
            CU(h->d_am_off.ensure((size_t)R + 1));
            CU(h->d_am_idx.ensure((size_t)std::max<int64_t>(tot, 1)));
            CU(cudaMemcpyAsync(h->d_am_off.p, am_off, ((size_t)R + 1) * 8, cudaMemcpyHostToDevice, h->stream));
            p.am_off = h->d_am_off.p;
            p.am_idx = h->d_am_idx.p;
            rc = launch_rescore_tiles_k(h, p, n_tiles, pl.max_width, k, 1);
            if (rc) return rc;
            CU(cudaMemcpyAsync(am_idx, h->d_am_idx.p, (size_t)tot * 4, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CU(cudaStreamSynchronize(h->stream));
    lap("results out (+ argmin lists)");
    return WEPP_OK;
}

}  // namespace

extern "C" {

int wepp_rescore(wepp_handle* h, int32_t n_cand, const int32_t* cand_nodes, int32_t* min_dist, int32_t* dist,
                 int64_t* am_off, int32_t* am_idx, int64_t am_capacity) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_reads) return fail(WEPP_E_STATE, "wepp_set_reads must be called first");
    if (n_cand < 1 || !cand_nodes || !min_dist) return fail(WEPP_E_INVALID, "bad candidate set / min_dist is NULL");
    CU(cudaSetDevice(h->device));
    std::string err;
    // the tile kernel over the resident reads; WEPP_RESCORE_GENERIC=1 selects the generic one-thread-per-read
    // kernel of wepp_rescore_reads instead (development: A/B timing)
    static const bool generic = getenv("WEPP_RESCORE_GENERIC") && atoi(getenv("WEPP_RESCORE_GENERIC")) != 0;
    if (!generic) return rescore_resident(h, n_cand, cand_nodes, min_dist, dist, am_off, am_idx, am_capacity);
    int rc = ensure_host_reads(h);
    if (rc) return rc;
    rc = rescore_run(h->device, h->stream, h->n_nodes, h->genome, h->parent.data(), h->mut_off.data(), h->mut_pos.data(),
                         h->mut_ref.data(), h->mut_nuc.data(), h->n_reads, h->r_start.data(), h->r_end.data(),
                         h->r_off.data(), h->r_pos.data(), h->r_nuc.data(), n_cand, cand_nodes, min_dist, dist, am_off,
                         am_idx, am_capacity, err);
    if (rc) return fail(rc, err);
    return WEPP_OK;
}

int wepp_rescore_reads(wepp_handle* h, int64_t n_reads, const int32_t* start, const int32_t* end, const int64_t* rm_off,
                       const int32_t* rm_pos, const uint8_t* rm_nuc, int32_t n_cand, const int32_t* cand_nodes,
                       int32_t* min_dist, int32_t* dist, int64_t* am_off, int32_t* am_idx, int64_t am_capacity) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_arena) return fail(WEPP_E_STATE, "wepp_set_arena must be called first");
    if (n_reads < 0 || (n_reads > 0 && (!start || !end || !rm_off))) return fail(WEPP_E_INVALID, "bad read arrays");
    if (n_cand < 1 || !cand_nodes || !min_dist) return fail(WEPP_E_INVALID, "bad candidate set / min_dist is NULL");
    CU(cudaSetDevice(h->device));
    static const int64_t zero_off[1] = {0};
    std::string err;
    int rc = rescore_run(h->device, h->stream, h->n_nodes, h->genome, h->parent.data(), h->mut_off.data(), h->mut_pos.data(),
                         h->mut_ref.data(), h->mut_nuc.data(), n_reads, start, end, n_reads ? rm_off : zero_off, rm_pos, rm_nuc,
                         n_cand, cand_nodes, min_dist, dist, am_off, am_idx, am_capacity, err);
    if (rc) return fail(rc, err);
    return WEPP_OK;
}

}  // extern "C"

// ---- greedy peak selection (wepp_filter::filter, initial_filter.cpp:455-506) ---------------------
namespace {

constexpr int MAX_PEAK_PEAK_MUTATION = 2;   // src/WEPP/config.hpp:19
constexpr int FREYJA_PEAKS_LIMIT = 5000;    // :20
constexpr int TOP_N = 10;                   // :21
constexpr int MAX_PEAKS = 300;              // :22
constexpr int MAX_NEIGHBORS_WEPP = 50;      // :23

// Host view of the arena for the haplotype-to-haplotype work of the peak loop.
struct PeakHost {
    const wepp_handle* h;
    std::vector<int64_t> own_child_off;
    std::vector<int32_t> own_child;
    const std::vector<int64_t>& child_off;   // the children lists: this object's own, or another's (worker threads)
    const std::vector<int32_t>& child;
    std::unordered_map<int32_t, std::vector<std::pair<int32_t, uint8_t>>> cache;
    std::vector<int64_t> last;

    // a worker's view: shares the children lists, has its own stack cache and scratch
    PeakHost(const wepp_handle* hh, const PeakHost& share) : h(hh), child_off(share.child_off), child(share.child) {
        last.assign((size_t)h->genome + 1, -1);
    }
    explicit PeakHost(const wepp_handle* hh) : h(hh), child_off(own_child_off), child(own_child) {
        std::vector<int64_t>& child_off = own_child_off;
        std::vector<int32_t>& child = own_child;
        const int32_t n = h->n_nodes;
        child_off.assign((size_t)n + 1, 0);
        for (int32_t v = 1; v < n; ++v) ++child_off[h->parent[v] + 1];
        for (int32_t v = 0; v < n; ++v) child_off[v + 1] += child_off[v];
        child.resize((size_t)std::max(n - 1, 0));
        std::vector<int64_t> cur(child_off.begin(), child_off.end() - 1);
        for (int32_t v = 1; v < n; ++v) child[cur[h->parent[v]]++] = v;   // preorder = the reference's creation order
        last.assign((size_t)h->genome + 1, -1);
    }

    // haplotype::stack_muts (arena.cpp:18-46): the last event per position on the root path, kept when it differs
    // from the reference allele, sorted by position
    const std::vector<std::pair<int32_t, uint8_t>>& stack(int32_t v) {
        auto it = cache.find(v);
        if (it != cache.end()) return it->second;
        std::vector<int32_t> path, touched;
        for (int32_t u = v; u >= 0; u = h->parent[u]) path.push_back(u);
        for (auto p = path.rbegin(); p != path.rend(); ++p)
            for (int64_t k = h->mut_off[*p]; k < h->mut_off[*p + 1]; ++k) {
                if (last[h->mut_pos[k]] < 0) touched.push_back(h->mut_pos[k]);
                last[h->mut_pos[k]] = k;
            }
        std::sort(touched.begin(), touched.end());
        std::vector<std::pair<int32_t, uint8_t>> st;
        for (int32_t pos : touched) {
            const int64_t k = last[pos];
            if (h->mut_ref[k] != h->mut_nuc[k]) st.emplace_back(pos, h->mut_nuc[k]);
            last[pos] = -1;
        }
        return cache.emplace(v, std::move(st)).first->second;
    }

    // a->mutation_distance(b) (haplotype.hpp:123-181 with comp = b->stack_muts, [0, INT_MAX])
    int dist(int32_t a, int32_t b) {
        const auto& A = stack(a);
        const auto B = stack(b);   // copy: stack(a) may rehash the cache
        const auto& A2 = stack(a);
        size_t i = 0, j = 0;
        int m = 0;
        while (i < A2.size() || j < B.size()) {
            if (i == A2.size()) { m += B[j].second != 15; ++j; }
            else if (j == B.size()) { ++m; ++i; }
            else if (A2[i].first < B[j].first) { ++m; ++i; }
            else if (A2[i].first > B[j].first) { m += B[j].second != 15; ++j; }
            else { m += (A2[i].second != B[j].second) && (B[j].second != 15); ++i; ++j; }
        }
        (void)A;
        return m;
    }

    // The same walk at the largest radius of the neighbour expansion, once: every node of the radius-`radius`
    // neighbourhood with the smallest radius at which neighbors() reaches it (the largest distance from the pivot over
    // the tree path between them — the climb stops at the first ancestor beyond the radius and the descent does not
    // pass a node beyond it).  neighbors(pivot, r, mapped) = the nodes here with need <= r that are not mapped.
    void neighborhood(int32_t pivot, int radius, std::vector<std::pair<int32_t, int>>& out) {
        out.clear();
        std::unordered_map<int32_t, int> chain;   // the pivot's ancestors within reach: need along the climb
        int32_t curr = pivot;
        int need = 0;
        chain.emplace(pivot, 0);
        while (h->parent[curr] >= 0) {
            const int d = dist(pivot, h->parent[curr]);
            if (d > radius) break;
            need = std::max(need, d);
            curr = h->parent[curr];
            chain.emplace(curr, need);
        }
        std::vector<std::pair<int32_t, int>> st{{curr, 0}};
        while (!st.empty()) {
            const int32_t v = st.back().first;
            const int above = st.back().second;
            st.pop_back();
            int nd;
            const auto it = chain.find(v);
            if (it != chain.end()) {
                nd = it->second;
            } else {
                const int d = dist(pivot, v);
                if (d > radius) continue;
                nd = std::max(above, d);
            }
            out.emplace_back(v, nd);
            for (int64_t k = child_off[v + 1] - 1; k >= child_off[v]; --k) st.emplace_back(child[k], nd);
        }
    }

    // arena::highest_scoring_neighbors(pivot, include_mapped = false, radius, INT_MAX) (arena.cpp:209-249),
    // in the order the reference's recursion inserts them
    void neighbors(int32_t pivot, int radius, const std::vector<uint8_t>& mapped, std::vector<int32_t>& out) {
        out.clear();
        int32_t curr = pivot;
        while (h->parent[curr] >= 0 && dist(pivot, h->parent[curr]) <= radius) curr = h->parent[curr];
        std::vector<int32_t> st{curr};
        while (!st.empty()) {
            const int32_t v = st.back();
            st.pop_back();
            if (dist(pivot, v) > radius) continue;
            if (!mapped[v]) out.push_back(v);
            for (int64_t k = child_off[v + 1] - 1; k >= child_off[v]; --k) st.push_back(child[k]);
        }
    }
};

}  // namespace

namespace {

__global__ void inverse_order_kernel(const uint32_t* __restrict__ order, int64_t n, uint32_t* __restrict__ pos_of_read) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos_of_read[order[i]] = (uint32_t)i;
}
__global__ void gather_pos_kernel(const int64_t* __restrict__ read_idx, int n, const uint32_t* __restrict__ pos_of_read, uint32_t* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = pos_of_read[read_idx[i]];
}
__global__ void gather_rec_kernel(const uint32_t* __restrict__ pos, int n, const uint4* __restrict__ rec, uint4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rec[pos[i]];
}

// remove_read for a batch (initial_filter.cpp:284-342) through the sparse corrections of the whole read set's plan:
// the per-node weights node_score(max_parsimony, multiplicity, degree) of the reads d_read_idx[0 .. n) — resident reads,
// caller order — at their minimum-parsimony nodes, into score_out[N].  The reads are looked up in the plan's sorted
// order, their records gathered, one work unit made per window group present (the host sees only the n sorted
// positions), and delta_place_kernel / delta_finalize_kernel run over those units; the per-node sums come from the
// difference arrays (no counts matrix).  Nothing is rebuilt per call — the window lists of a subset plan were.
int place_subset_by_delta(wepp_handle* h, const int64_t* d_read_idx, int n, double* score_out) {
    wepp_handle::DevPlan& dp = h->full;
    cudaStream_t st = h->stream;
    const int64_t R = dp.plan.n_reads;
    if (!dp.pos_ready) {
        CU(dp.pos_of_read.ensure((size_t)R));
        inverse_order_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(dp.order.p, R, dp.pos_of_read.p);
        CU(cudaGetLastError());
        dp.pos_ready = true;
    }
    CU(dp.sub_pos.ensure((size_t)n)); CU(dp.sub_pos_sorted.ensure((size_t)n)); CU(dp.rec_sub.ensure((size_t)n));
    gather_pos_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_read_idx, n, dp.pos_of_read.p, dp.sub_pos.p);
    CU(cudaGetLastError());
    int bits = 1;
    while ((1ll << bits) < R) ++bits;
    size_t tmp = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp, dp.sub_pos.p, dp.sub_pos_sorted.p, n, 0, bits, st));
    CU(h->d_cub_tmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortKeys(h->d_cub_tmp.p, tmp, dp.sub_pos.p, dp.sub_pos_sorted.p, n, 0, bits, st));
    gather_rec_kernel<<<(n + 255) / 256, 256, 0, st>>>(dp.sub_pos_sorted.p, n, dp.rec.p, dp.rec_sub.p);
    CU(cudaGetLastError());
    std::vector<uint32_t> pos((size_t)n);
    CU(cudaMemcpyAsync(pos.data(), dp.sub_pos_sorted.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<DeltaUnit> units;
    const std::vector<int64_t>& gf = dp.h_group_first;
    for (int i = 0; i < n;) {
        const int g = (int)(std::upper_bound(gf.begin(), gf.end(), (int64_t)pos[(size_t)i]) - gf.begin()) - 1;
        int j = i;
        while (j < n && (int64_t)pos[(size_t)j] < gf[(size_t)g + 1]) ++j;
        for (int o = i; o < j;) {   // units of DP_UNIT reads; a remainder of up to 2 * DP_UNIT stays whole
            const int take = j - o <= 2 * DP_UNIT ? j - o : DP_UNIT;
            units.push_back(DeltaUnit{g, o, take, 0});
            o += take;
        }
        i = j;
    }
    CU(upload(dp.units_sub, units, st));
    const DeltaSubset sub = {dp.units_sub.p, (int32_t)units.size(), dp.rec_sub.p};
    return run_place(h, dp, true, 0, 0, /*with_counts*/ false, score_out, &sub);
}

}  // namespace

extern "C" int wepp_filter_peaks(wepp_handle* h, const int32_t* leaf_count, const int32_t* id_rank, int32_t* out_nodes,
                                 int32_t capacity, int32_t* n_peaks_out, int32_t* n_out) {
    if (!h) return fail(WEPP_E_INVALID, "handle is NULL");
    if (!h->has_reads) return fail(WEPP_E_STATE, "wepp_set_reads must be called first");
    if (!leaf_count || !id_rank || !n_out) return fail(WEPP_E_INVALID, "NULL argument");
    CU(cudaSetDevice(h->device));
    const int n = h->n_nodes;
    const int64_t R = h->n_reads;
    cudaStream_t st = h->stream;

    const auto t_enter = std::chrono::steady_clock::now();
    // reset_haplotype_state + cartesian_map (initial_filter.cpp:458-466)
    h->has_mask = false;
    h->full.final_for_mask = false;
    h->sub.final_for_mask = false;
    int rc = run_place(h, h->full, true, 0, 0);
    if (rc) return rc;
    CU(h->d_divergence.ensure((size_t)n));
    {
        BinCounts tc;
        int active = 0;
        for (int j = 0; j < NBINS; ++j) {
            tc.v[j] = h->true_counts[j];
            active += tc.v[j] != 0;
        }
        divergence_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->d_counts.p, n, tc, active, 0.5 / 100, h->d_divergence.p);
        CU(cudaGetLastError());
    }
    DevBuf<double> d_cur, d_orig, d_contrib, d_fulls;
    DevBuf<uint8_t> d_pmapped, d_removed, d_stnuc;
    DevBuf<int32_t> d_nodes, d_stpos, d_marks;
    DevBuf<int64_t> d_list, d_stoff;
    DevBuf<unsigned long long> d_max;
    DevBuf<int> d_count;
    auto release = [&]() {
        d_cur.release(); d_orig.release(); d_contrib.release(); d_fulls.release(); d_pmapped.release();
        d_removed.release(); d_stnuc.release(); d_nodes.release(); d_stpos.release(); d_marks.release();
        d_list.release(); d_stoff.release(); d_max.release(); d_count.release();
    };
    (void)0;
#define FCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            release();                                                                             \
            return fail(WEPP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));         \
        }                                                                                          \
    } while (0)
    int cand_cap = 1 << 20;   // nodes within SCORE_EPSILON of the top score that fit the buffers (grown on demand)
    FCU(d_cur.ensure((size_t)n)); FCU(d_orig.ensure((size_t)n)); FCU(d_contrib.ensure((size_t)n));
    FCU(d_pmapped.ensure((size_t)n)); FCU(d_removed.ensure((size_t)std::max<int64_t>(R, 1)));
    FCU(d_nodes.ensure((size_t)cand_cap)); FCU(d_fulls.ensure((size_t)cand_cap)); FCU(d_max.ensure(1)); FCU(d_count.ensure(1));
    FCU(d_list.ensure((size_t)std::max<int64_t>(R, 1)));
    FCU(cudaMemcpyAsync(d_cur.p, h->d_score.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    FCU(cudaMemcpyAsync(d_orig.p, h->d_score.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    FCU(cudaMemsetAsync(d_pmapped.p, 0, (size_t)n, st));
    FCU(cudaMemsetAsync(d_removed.p, 0, (size_t)std::max<int64_t>(R, 1), st));
    // find_correspondents reads the resident reads (caller order).  The removed reads' per-node weights come from the
    // sparse corrections of the whole read set's plan where the cartesian_map above took that path (WEPP_PEAK_DELTA=0:
    // never), else from a placement of a subset plan over its own window lists, which needs the host copy of the reads
    const bool subset_by_delta = h->stats.place_path == 2 && h->full.delta_groups_usable &&
                                 !(getenv("WEPP_PEAK_DELTA") && atoi(getenv("WEPP_PEAK_DELTA")) == 0);
    if (!subset_by_delta && (rc = ensure_host_reads(h)) != 0) {
        release();
        return rc;
    }

    // read-sharded ranks (wepp_set_allreduce with one plan: wepp_group_filter_peaks, or a caller's own ranks in step):
    // the cartesian_map above merged the accumulators, so score / counts / divergence are those of the whole read set
    // on every rank, bit for bit; every step below exchanges the number of reads removed and their per-node weights
    const bool sharded = h->allreduce && h->shared_plan;
    DevBuf<long long> d_xchg;
    auto sum_over_ranks = [&](int64_t& v) -> int {
        if (!sharded) return WEPP_OK;
        long long x = v;
        CU(d_xchg.ensure(1));
        CU(cudaMemcpyAsync(d_xchg.p, &x, 8, cudaMemcpyHostToDevice, st));
        if (h->allreduce(h->allreduce_user, d_xchg.p, 1, WEPP_DTYPE_I64, (void*)st) != 0) return fail(WEPP_E_STATE, "the all-reduce hook failed");
        CU(cudaMemcpyAsync(&x, d_xchg.p, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        v = x;
        return WEPP_OK;
    };
    int64_t R_all = R;
    if ((rc = sum_over_ranks(R_all)) != 0) {
        release();
        return rc;
    }

    PeakHost ph(h);
    std::vector<uint8_t> mapped((size_t)n, 0);
    std::set<int32_t> peaks;
    int64_t remaining = R_all;
    auto cmp_tie = [&](int32_t l, int32_t r) {   // score_comparator's tie-breaks (arena.hpp:24-29)
        if (leaf_count[l] != leaf_count[r]) return leaf_count[l] > leaf_count[r];
        return id_rank[l] > id_rank[r];
    };

    std::vector<int32_t> cand_nodes, consideration, nb, marks;
    std::vector<double> cand_full;
    std::vector<int64_t> removed_now;
    // WEPP_TIMING=1: accumulated phase times of the loop on stderr (development aid; adds synchronisations)
    const bool timing = getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) != 0;
    double t_phase[6] = {};
    int n_steps = 0;
    int64_t n_removed_total = 0;
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](int k) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const auto t = std::chrono::steady_clock::now();
        t_phase[k] += std::chrono::duration<double, std::milli>(t - t_mark).count();
        t_mark = t;
    };
    if (timing) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "[wepp timing] initial filter: cartesian_map + set-up of the loop %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count());
        t_mark = std::chrono::steady_clock::now();
    }
    while (remaining > 0 && (int)peaks.size() < MAX_PEAKS) {
        ++n_steps;
        lap(5);
        // ---- the head of the sorted `current` list (initial_filter.cpp:396-417) -------------------
        FCU(cudaMemsetAsync(d_max.p, 0, 8, st));
        peak_max_kernel<<<1184, 256, 0, st>>>(d_cur.p, h->d_divergence.p, d_pmapped.p, n, d_max.p);
        unsigned long long top_bits = 0;
        FCU(cudaMemcpyAsync(&top_bits, d_max.p, 8, cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        double top;
        std::memcpy(&top, &top_bits, 8);
        if (!(top >= PEAK_SCORE_EPSILON)) break;   // no available peaks (:401-404) / current is empty
        int n_cand = 0;
        for (;;) {
            FCU(cudaMemsetAsync(d_count.p, 0, sizeof(int), st));
            peak_collect_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_cur.p, h->d_divergence.p, d_pmapped.p, n, top, d_count.p,
                                                                 d_nodes.p, d_fulls.p, cand_cap);
            FCU(cudaMemcpyAsync(&n_cand, d_count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            FCU(cudaStreamSynchronize(st));
            if (n_cand <= cand_cap) break;
            // more nodes tie for the top score than the buffers hold (one uninformative read alone can put millions of
            // nodes at the same score; the reference walks its sorted list and never fails here): grow and collect again
            cand_cap = n_cand;
            FCU(d_nodes.ensure((size_t)cand_cap)); FCU(d_fulls.ensure((size_t)cand_cap));
        }
        cand_nodes.resize((size_t)n_cand);
        cand_full.resize((size_t)n_cand);
        FCU(cudaMemcpyAsync(cand_nodes.data(), d_nodes.p, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, st));
        FCU(cudaMemcpyAsync(cand_full.data(), d_fulls.p, (size_t)n_cand * 8, cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        // all candidates are within SCORE_EPSILON of the maximum, hence of each other: the comparator orders
        // them by leaf_count, then id (descending)
        std::vector<int32_t> order(cand_nodes.begin(), cand_nodes.end());
        std::sort(order.begin(), order.end(), cmp_tie);

        consideration.clear();
        marks.clear();
        for (int32_t v : order) {
            if ((int)consideration.size() >= TOP_N || (int)consideration.size() + (int)peaks.size() >= MAX_PEAKS) break;
            bool valid = true;
            for (int32_t old : consideration)
                if (!(ph.dist(old, v) > MAX_PEAK_PEAK_MUTATION)) valid = false;   // valid_two_tops
            if (valid) {
                consideration.push_back(v);
                mapped[v] = 1;
                marks.push_back(v);
            }
        }
        lap(0);
        // ---- clear_neighbors (:368-385) ------------------------------------------------------------
        for (int32_t pivot : consideration) {
            peaks.insert(pivot);
            ph.neighbors(pivot, MAX_PEAK_PEAK_MUTATION, mapped, nb);
            for (int32_t v : nb) {
                mapped[v] = 1;
                marks.push_back(v);
            }
        }
        lap(1);
        FCU(upload(d_marks, marks, st));
        if (!marks.empty()) mark_kernel<<<((int)marks.size() + 255) / 256, 256, 0, st>>>(d_pmapped.p, d_marks.p, (int)marks.size());
        // ---- singular_step for every chosen peak (:344-366): correspondents, then their removal -----
        std::vector<int64_t> so{0};
        std::vector<int32_t> sp;
        std::vector<uint8_t> sn;
        for (int32_t pivot : consideration) {
            for (const auto& m : ph.stack(pivot)) {
                sp.push_back(m.first);
                sn.push_back(m.second);
            }
            so.push_back((int64_t)sp.size());
        }
        FCU(upload(d_stoff, so, st)); FCU(upload(d_stpos, sp, st)); FCU(upload(d_stnuc, sn, st));
        FCU(cudaMemsetAsync(d_count.p, 0, sizeof(int), st));
        CorrespondParams cp = {};
        cp.n_reads = R; cp.start = h->d_rstart.p; cp.end = h->d_rend.p; cp.rm_off = h->d_roff.p; cp.rm_pos = h->d_rpos.p;
        cp.rm_nuc = h->d_rnuc.p; cp.max_pars = h->d_maxpars.p; cp.removed = d_removed.p;
        cp.n_cand = (int32_t)consideration.size(); cp.st_off = d_stoff.p; cp.st_pos = d_stpos.p; cp.st_nuc = d_stnuc.p;
        cp.count = d_count.p; cp.list = d_list.p;
        if (R > 0 && !consideration.empty()) correspond_kernel<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(cp);
        int n_rem = 0;
        FCU(cudaMemcpyAsync(&n_rem, d_count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        lap(2);
        int64_t n_rem_all = n_rem;
        if ((rc = sum_over_ranks(n_rem_all)) != 0) {
            release();
            return rc;
        }
        n_removed_total += n_rem_all;
        if (n_rem_all > 0) {
            if (n_rem > 0 && subset_by_delta) {
                if ((rc = place_subset_by_delta(h, d_list.p, n_rem, d_contrib.p)) != 0) {
                    release();
                    return rc;
                }
            } else if (n_rem > 0) {
                removed_now.resize((size_t)n_rem);
                FCU(cudaMemcpyAsync(removed_now.data(), d_list.p, (size_t)n_rem * 8, cudaMemcpyDeviceToHost, st));
                FCU(cudaStreamSynchronize(st));
                std::sort(removed_now.begin(), removed_now.end());
                std::string err = build_read_plan(h->es, h->genome, h->n_reads, h->r_start.data(), h->r_end.data(),
                                                  h->r_degree.data(), h->r_off.data(), h->r_pos.data(), h->r_nuc.data(),
                                                  h->opt_k, removed_now.data(), n_rem, h->sub.plan);
                if (!err.empty()) {
                    release();
                    return fail(WEPP_E_INVALID, err);
                }
                lap(3);
                rc = upload_plan(h, h->sub, true);
                if (!rc) rc = run_place(h, h->sub, true, 0, 0, /*with_counts*/ false, d_contrib.p);
                if (rc) {
                    release();
                    return rc;
                }
            } else {
                FCU(cudaMemsetAsync(d_contrib.p, 0, (size_t)n * 8, st));   // none of this rank's reads: it still takes part
            }
            if (sharded && h->allreduce(h->allreduce_user, d_contrib.p, n, WEPP_DTYPE_F64, (void*)st) != 0) {
                release();
                return fail(WEPP_E_STATE, "the all-reduce hook failed");
            }
            subtract_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_cur.p, d_contrib.p, n);
            remaining -= n_rem_all;
            lap(4);
        }
        if (consideration.empty()) break;   // nothing selectable: the reference would spin on the same head
    }

    if (timing)
        fprintf(stderr, "[wepp timing] peak loop: %d steps, %lld reads removed | top score + ties %.1f ms | neighbourhoods (host) %.1f ms | "
                        "correspondents %.1f ms | subset plan (host) %.1f ms | subset lists + place + subtract %.1f ms | other %.1f ms\n",
                n_steps, (long long)n_removed_total, t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5]);
    // ---- neighbours of the peaks (:476-503) ---------------------------------------------------------
    const auto t_nb = std::chrono::steady_clock::now();
    std::vector<double> full((size_t)n), dv((size_t)n);
    FCU(cudaMemcpyAsync(full.data(), d_orig.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    FCU(cudaMemcpyAsync(dv.data(), h->d_divergence.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    FCU(cudaMemcpyAsync(h->d_score.p, d_orig.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));   // recover_state
    FCU(cudaStreamSynchronize(st));
    for (int v = 0; v < n; ++v) full[v] = full[v] * std::sqrt(dv[v]);   // haplotype::full_score with score = orig_score
    auto cmp = [&](int32_t l, int32_t r) {   // score_comparator (arena.hpp:16-31)
        if (std::fabs(full[l] - full[r]) > PEAK_SCORE_EPSILON) return full[l] > full[r];
        return cmp_tie(l, r);
    };
    // the five radii are nested: one walk per peak at the largest radius, on the host threads (each with its own stack
    // cache), records the radius each node is first reached at; the rounds below filter it
    const std::vector<int32_t> peak_list(peaks.begin(), peaks.end());
    std::vector<std::vector<std::pair<int32_t, int>>> hood(peak_list.size());
    {
        const int n_thr = (int)std::max<size_t>(1, std::min<size_t>({(size_t)16, (size_t)std::thread::hardware_concurrency(), peak_list.size()}));
        std::atomic<size_t> next{0};
        auto work = [&](PeakHost* me) {
            for (size_t i = next.fetch_add(1); i < peak_list.size(); i = next.fetch_add(1))
                me->neighborhood(peak_list[i], MAX_PEAK_PEAK_MUTATION + 4, hood[i]);
        };
        if (n_thr <= 1) {
            work(&ph);
        } else {
            std::vector<std::unique_ptr<PeakHost>> views;
            std::vector<std::thread> pool;
            for (int t = 0; t < n_thr; ++t) views.emplace_back(new PeakHost(h, ph));
            for (int t = 0; t < n_thr; ++t) pool.emplace_back(work, views[(size_t)t].get());
            for (auto& th : pool) th.join();
        }
    }
    std::set<int32_t> nbrs;
    for (int k = 0; k < 5; ++k) {
        std::fill(mapped.begin(), mapped.end(), 0);
        std::set<int32_t> curr;
        for (size_t pi = 0; pi < peak_list.size(); ++pi) {
            nb.clear();
            for (const auto& vn : hood[pi])
                if (vn.second <= MAX_PEAK_PEAK_MUTATION + k && !mapped[vn.first]) nb.push_back(vn.first);
            std::set<int32_t, decltype(cmp)> ordered(cmp);
            for (int32_t v : nb) ordered.insert(v);
            int i = 0;
            for (int32_t v : ordered) {
                if (peaks.count(v) || curr.count(v)) continue;
                mapped[v] = 1;
                curr.insert(v);
                if (++i == MAX_NEIGHBORS_WEPP) break;
            }
        }
        if (std::abs(FREYJA_PEAKS_LIMIT - ((int)curr.size() + (int)peaks.size())) <
            std::abs(FREYJA_PEAKS_LIMIT - ((int)nbrs.size() + (int)peaks.size())))
            nbrs = curr;
    }
    if (timing)
        fprintf(stderr, "[wepp timing] neighbour expansion (host) %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_nb).count());
    release();
#undef FCU
    const int total = (int)peaks.size() + (int)nbrs.size();
    if (n_peaks_out) *n_peaks_out = (int32_t)peaks.size();
    *n_out = total;
    if (out_nodes) {
        if (capacity < total) return fail(WEPP_E_CAPACITY, "out_nodes capacity too small");
        int i = 0;
        for (int32_t v : peaks) out_nodes[i++] = v;
        for (int32_t v : nbrs) out_nodes[i++] = v;
    }
    h->has_results = true;
    return WEPP_OK;
}

#include "peer_group.cuh"
