// peer_group.cuh — several B200s of one box driven by ONE process (the product driver's multi-GPU mode; the
// torchrun / torch.distributed route of bench.py lends the library NCCL instead): a handle and a host thread per
// rank, reads dealt round-robin, the tree replicated, and the exchanges of wepp_set_allreduce served by a kernel over
// NVLink peer memory instead of a collective library.
//
// peer_allreduce_kernel: rank r owns slice r of the buffer.  It loads slice r of EVERY rank's buffer (peer loads,
// 16 bytes per lane, all ranks' loads of an element in flight together), adds them in rank order and stores the sum
// into slice r of every rank's buffer (peer stores) — reduce-scatter and all-gather in one pass, each element crossing
// the switch once in and once out per peer.  The sum order is fixed and each element is computed by exactly one rank,
// so every rank ends up with bit-identical buffers: the ranks of the peak loop then take the same decisions without
// talking to each other.  Ordering between the ranks' streams is by CUDA events (recorded before / waited for after a
// host-thread barrier), never by a device-side spin.
//
// Included at the end of wepp_abi.cu (it needs wepp_handle).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace wepp {

constexpr int PG_MAX_RANKS = 16;
struct PeerPtrs {
    void* p[PG_MAX_RANKS];
};

template <typename T>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerPtrs bufs, const int G, const int rank, const int64_t count) {
    constexpr int V = 16 / (int)sizeof(T);
    union Vec {
        int4 q;
        T t[V];
    };
    const int64_t n_vec = count / V;
    const int64_t v_lo = n_vec * rank / G, v_hi = n_vec * (rank + 1) / G;
    for (int64_t i = v_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_hi; i += (int64_t)gridDim.x * blockDim.x) {
        Vec x[PG_MAX_RANKS];
#pragma unroll
        for (int p = 0; p < PG_MAX_RANKS; ++p)
            if (p < G) x[p].q = reinterpret_cast<const int4*>(bufs.p[p])[i];
        Vec acc = x[0];
#pragma unroll
        for (int p = 1; p < PG_MAX_RANKS; ++p)
            if (p < G) {
#pragma unroll
                for (int k = 0; k < V; ++k) acc.t[k] += x[p].t[k];
            }
#pragma unroll
        for (int p = 0; p < PG_MAX_RANKS; ++p)
            if (p < G) reinterpret_cast<int4*>(bufs.p[p])[i] = acc.q;
    }
    // the elements past the last whole vector: the last rank's
    if (rank == G - 1 && blockIdx.x == 0) {
        for (int64_t i = n_vec * V + threadIdx.x; i < count; i += blockDim.x) {
            T acc = reinterpret_cast<const T*>(bufs.p[0])[i];
            for (int p = 1; p < G; ++p) acc += reinterpret_cast<const T*>(bufs.p[p])[i];
            for (int p = 0; p < G; ++p) reinterpret_cast<T*>(bufs.p[p])[i] = acc;
        }
    }
}

}  // namespace wepp

struct wepp_group {
    int G = 0;
    std::vector<int> dev;
    std::vector<wepp_handle*> h;
    struct RankRef {
        wepp_group* g;
        int rank;
    };
    std::vector<RankRef> refs;
    void* bufs[wepp::PG_MAX_RANKS] = {};
    cudaEvent_t ready[wepp::PG_MAX_RANKS] = {}, done[wepp::PG_MAX_RANKS] = {};
    // host-thread barrier; a rank that fails outside an exchange breaks it so that the others do not wait for ever
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    uint64_t generation = 0;
    bool broken = false;
    std::string err;
    // the reads as dealt: rank r holds the caller's reads r, r + G, r + 2G, ...
    int64_t n_reads = 0;
    // one host thread per rank for the life of the group (a thread spawned per call costs more than an exchange)
    std::vector<std::thread> workers;
    std::mutex tm;
    std::condition_variable task_cv, done_cv;
    std::function<int(int)> task;
    uint64_t task_gen = 0;
    int n_done = 0;
    bool quit = false;
    std::vector<int> rc;
    std::vector<std::string> msg;

    bool barrier() {
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        const uint64_t gen = generation;
        if (++waiting == G) {
            waiting = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        cv.wait(lk, [&] { return generation != gen || broken; });
        return generation != gen;
    }
    void break_barrier() {
        std::lock_guard<std::mutex> lk(m);
        broken = true;
        cv.notify_all();
    }
};

namespace {

// the hook handed to wepp_set_allreduce for every rank of a group
int peer_group_allreduce(void* user, void* dev_ptr, int64_t count, int32_t dtype, void* cuda_stream) {
    auto* ref = static_cast<wepp_group::RankRef*>(user);
    wepp_group* g = ref->g;
    const int r = ref->rank, G = g->G;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (count <= 0) return g->barrier() && g->barrier() ? 0 : 1;
    g->bufs[r] = dev_ptr;
    if (cudaEventRecord(g->ready[r], st) != cudaSuccess) {
        g->break_barrier();
        return 1;
    }
    if (!g->barrier()) return 1;   // every rank's buffer is published and its producer's work is behind `ready`
    wepp::PeerPtrs pp = {};
    for (int p = 0; p < G; ++p) {
        pp.p[p] = g->bufs[p];
        if (p != r && cudaStreamWaitEvent(st, g->ready[p], 0) != cudaSuccess) {
            g->break_barrier();
            return 1;
        }
    }
    const int64_t per_rank = count / G + 1;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((per_rank + 1023) / 1024, (int64_t)g->h[(size_t)r]->n_sms * 4));
    if (dtype == WEPP_DTYPE_F64) wepp::peer_allreduce_kernel<double><<<blocks, 256, 0, st>>>(pp, G, r, count);
    else if (dtype == WEPP_DTYPE_I64) wepp::peer_allreduce_kernel<long long><<<blocks, 256, 0, st>>>(pp, G, r, count);
    else wepp::peer_allreduce_kernel<int32_t><<<blocks, 256, 0, st>>>(pp, G, r, count);
    if (cudaGetLastError() != cudaSuccess || cudaEventRecord(g->done[r], st) != cudaSuccess) {
        g->break_barrier();
        return 1;
    }
    if (!g->barrier()) return 1;   // every rank's slice kernel is enqueued: nobody reads its buffer before all of them ran
    for (int p = 0; p < G; ++p)
        if (p != r && cudaStreamWaitEvent(st, g->done[p], 0) != cudaSuccess) {
            g->break_barrier();
            return 1;
        }
    return 0;
}

void group_worker(wepp_group* g, int r) {
    cudaSetDevice(g->dev[(size_t)r]);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(int)> fn;
        {
            std::unique_lock<std::mutex> lk(g->tm);
            g->task_cv.wait(lk, [&] { return g->quit || g->task_gen != seen; });
            if (g->quit) return;
            seen = g->task_gen;
            fn = g->task;
        }
        const int rc = fn(r);
        if (rc) {
            g->msg[(size_t)r] = wepp_last_error();
            g->break_barrier();
        }
        std::lock_guard<std::mutex> lk(g->tm);
        g->rc[(size_t)r] = rc;
        if (++g->n_done == g->G) g->done_cv.notify_all();
    }
}

// fn(rank) on the ranks' threads, concurrently (the exchanges are rendezvous); the first failure is reported
template <typename F>
int group_run(wepp_group* g, F fn) {
    {
        std::lock_guard<std::mutex> lk(g->m);
        g->broken = false;
        g->waiting = 0;
    }
    {
        std::unique_lock<std::mutex> lk(g->tm);
        g->task = fn;
        g->n_done = 0;
        std::fill(g->rc.begin(), g->rc.end(), 0);
        ++g->task_gen;
        g->task_cv.notify_all();
        g->done_cv.wait(lk, [&] { return g->n_done == g->G; });
        g->task = nullptr;
    }
    // a rank that failed on its own comes before the ranks that only saw the broken exchange
    int first = -1;
    for (int r = 0; r < g->G; ++r)
        if (g->rc[(size_t)r] && (first < 0 || (g->msg[(size_t)first].find("all-reduce hook") != std::string::npos &&
                                               g->msg[(size_t)r].find("all-reduce hook") == std::string::npos)))
            first = r;
    if (first < 0) return WEPP_OK;
    return fail(g->rc[(size_t)first], "rank " + std::to_string(first) + ": " + g->msg[(size_t)first]);
}

}  // namespace

extern "C" {

int wepp_group_create(int32_t n_ranks, const int32_t* devices, wepp_group** out) {
    if (!out || n_ranks < 1 || n_ranks > wepp::PG_MAX_RANKS) return fail(WEPP_E_INVALID, "a group has 1.." + std::to_string(wepp::PG_MAX_RANKS) + " ranks");
    auto* g = new wepp_group();
    g->G = n_ranks;
    g->refs.resize((size_t)n_ranks);
    for (int r = 0; r < n_ranks; ++r) g->dev.push_back(devices ? devices[r] : r);
    auto bail = [&](int rc) {
        const std::string keep = wepp_last_error();
        wepp_group_destroy(g);
        return fail(rc, keep);
    };
    for (int r = 0; r < n_ranks; ++r) {
        wepp_handle* h = nullptr;
        const int rc = wepp_create(g->dev[(size_t)r], &h);
        if (rc) return bail(rc);
        g->h.push_back(h);
        // the other ranks' devices read and write this rank's buffers (and the other way round)
        for (int p = 0; p < n_ranks; ++p) {
            if (g->dev[(size_t)p] == g->dev[(size_t)r]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, g->dev[(size_t)r], g->dev[(size_t)p]) != cudaSuccess || !can)
                return bail(fail(WEPP_E_CUDA, "device " + std::to_string(g->dev[(size_t)r]) + " cannot access device " +
                                                  std::to_string(g->dev[(size_t)p]) + " as a peer"));
            const cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[(size_t)p], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(fail(WEPP_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)));
            (void)cudaGetLastError();
        }
        if (cudaEventCreateWithFlags(&g->ready[r], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g->done[r], cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(WEPP_E_CUDA, "cudaEventCreate failed"));
        g->refs[(size_t)r] = {g, r};
        if (n_ranks > 1) wepp_set_allreduce(h, peer_group_allreduce, &g->refs[(size_t)r]);
    }
    g->rc.assign((size_t)n_ranks, 0);
    g->msg.assign((size_t)n_ranks, std::string());
    for (int r = 0; r < n_ranks; ++r) g->workers.emplace_back(group_worker, g, r);
    *out = g;
    return WEPP_OK;
}

void wepp_group_destroy(wepp_group* g) {
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->tm);
        g->quit = true;
        g->task_cv.notify_all();
    }
    for (auto& t : g->workers) t.join();
    for (size_t r = 0; r < g->h.size(); ++r) {
        if (g->h[r]) wepp_destroy(g->h[r]);
        else cudaSetDevice(g->dev[r]);
        if (g->ready[r]) cudaEventDestroy(g->ready[r]);
        if (g->done[r]) cudaEventDestroy(g->done[r]);
    }
    delete g;
}

int32_t wepp_group_size(const wepp_group* g) { return g ? g->G : 0; }

wepp_handle* wepp_group_handle(wepp_group* g, int32_t rank) {
    return g && rank >= 0 && rank < g->G ? g->h[(size_t)rank] : nullptr;
}

wepp_handle* wepp_group_take(wepp_group* g, int32_t rank) {
    if (!g || rank < 0 || rank >= g->G || !g->h[(size_t)rank]) return nullptr;
    wepp_handle* h = g->h[(size_t)rank];
    g->h[(size_t)rank] = nullptr;
    wepp_set_allreduce(h, nullptr, nullptr);
    return h;
}

// group-wide calls need every rank: a handle taken out (wepp_group_take) ends the group's collective life
static int group_whole(const wepp_group* g) {
    if (!g) return fail(WEPP_E_INVALID, "group is NULL");
    for (const wepp_handle* h : g->h)
        if (!h) return fail(WEPP_E_STATE, "a handle was taken out of the group");
    return WEPP_OK;
}

int wepp_group_run(wepp_group* g, wepp_group_fn fn, void* user) {
    if (!fn) return fail(WEPP_E_INVALID, "NULL argument");
    if (int rc = group_whole(g)) return rc;
    return group_run(g, [&](int r) { return fn(r, g->h[(size_t)r], user); });
}

int wepp_group_set_arena(wepp_group* g, int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                         const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size) {
    if (int rc = group_whole(g)) return rc;
    if (!parent || !mut_off) return fail(WEPP_E_INVALID, "NULL argument");
    // one host flatten for all ranks (it already uses every core); each rank copies and uploads it
    wepp::EulerStripes es;
    const std::string err = wepp::build_euler_stripes(n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size, g->h[0]->opt_q, es);
    if (!err.empty()) return fail(WEPP_E_INVALID, err);
    return group_run(g, [&](int r) {
        return set_arena_with(g->h[(size_t)r], n_nodes, parent, mut_off, mut_pos, mut_ref, mut_nuc, genome_size,
                              g->h[(size_t)r]->opt_q == g->h[0]->opt_q ? &es : nullptr);
    });
}

int wepp_group_set_reads(wepp_group* g, int64_t n_reads, const int32_t* start, const int32_t* end, const int32_t* degree,
                         const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc) {
    if (int rc = group_whole(g)) return rc;
    if (n_reads < 0 || (n_reads > 0 && (!start || !end || !degree || !rm_off))) return fail(WEPP_E_INVALID, "NULL argument");
    g->n_reads = n_reads;
    const int G = g->G;
    return group_run(g, [&](int r) {
        // the caller's reads r, r + G, r + 2G, ...
        std::vector<int32_t> s, e, d, mp;
        std::vector<int64_t> off{0};
        std::vector<uint8_t> mn;
        for (int64_t i = r; i < n_reads; i += G) {
            s.push_back(start[i]); e.push_back(end[i]); d.push_back(degree[i]);
            for (int64_t k = rm_off[i]; k < rm_off[i + 1]; ++k) {
                mp.push_back(rm_pos[k]);
                mn.push_back(rm_nuc[k]);
            }
            off.push_back((int64_t)mp.size());
        }
        return wepp_set_reads(g->h[(size_t)r], (int64_t)s.size(), s.data(), e.data(), d.data(), off.data(), mp.data(), mn.data());
    });
}

int wepp_group_place(wepp_group* g) {
    if (int rc = group_whole(g)) return rc;
    return group_run(g, [&](int r) {
        const int rc = wepp_place(g->h[(size_t)r], 0, 0);
        return rc ? rc : wepp_sync(g->h[(size_t)r]);
    });
}

int wepp_group_get_read_results(wepp_group* g, int32_t* max_parsimony, int32_t* multiplicity) {
    if (int rc = group_whole(g)) return rc;
    const int G = g->G;
    const int64_t n = g->n_reads;
    return group_run(g, [&](int r) {
        const int64_t mine = n > r ? (n - r + G - 1) / G : 0;
        std::vector<int32_t> mp((size_t)mine), mu((size_t)mine);
        const int rc = wepp_get_read_results(g->h[(size_t)r], mp.data(), mu.data());
        if (rc) return rc;
        for (int64_t k = 0; k < mine; ++k) {
            if (max_parsimony) max_parsimony[r + k * G] = mp[(size_t)k];
            if (multiplicity) multiplicity[r + k * G] = mu[(size_t)k];
        }
        return 0;
    });
}

int wepp_group_filter_peaks(wepp_group* g, const int32_t* leaf_count, const int32_t* id_rank, int32_t* out_nodes, int32_t capacity,
                            int32_t* n_peaks_out, int32_t* n_out) {
    if (!n_out) return fail(WEPP_E_INVALID, "NULL argument");
    if (int rc = group_whole(g)) return rc;
    std::vector<std::vector<int32_t>> outs((size_t)g->G);
    std::vector<int32_t> np((size_t)g->G, 0), no((size_t)g->G, 0);
    const int32_t n = g->h[0]->n_nodes;
    int rc = group_run(g, [&](int r) {
        outs[(size_t)r].resize((size_t)std::max(n, 1));
        return wepp_filter_peaks(g->h[(size_t)r], leaf_count, id_rank, outs[(size_t)r].data(), n, &np[(size_t)r], &no[(size_t)r]);
    });
    if (rc) return rc;
    for (int r = 1; r < g->G; ++r)   // the ranks decide on bit-identical merged scores: anything else is a bug, never a result
        if (np[(size_t)r] != np[0] || no[(size_t)r] != no[0] || !std::equal(outs[0].begin(), outs[0].begin() + no[0], outs[(size_t)r].begin()))
            return fail(WEPP_E_STATE, "the ranks of the group chose different peaks");
    if (n_peaks_out) *n_peaks_out = np[0];
    *n_out = no[0];
    if (out_nodes) {
        if (capacity < no[0]) return fail(WEPP_E_CAPACITY, "out_nodes capacity too small");
        std::copy(outs[0].begin(), outs[0].begin() + no[0], out_nodes);
    }
    return WEPP_OK;
}

}  // extern "C"
