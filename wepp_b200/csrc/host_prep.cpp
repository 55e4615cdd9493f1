// host_prep.cpp — see host_prep.h.
#include "host_prep.h"

#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <thread>
#include <unordered_map>

namespace wepp {

namespace {

constexpr int NUM_RANGE_BINS = 50;  // reference: src/WEPP/config.hpp:13

// Mismatch state of a haplotype whose last event at a position has (mut, ref), for a read
// whose allele class there is c: 0 = "as reference" (no entry in the read's mutation list),
// 1..4 = the read carries A/C/G/T.  Reference: src/WEPP/initial_filter.cpp:64-66 —
//   read_nuc = (read has a mutation here) ? its mut_nuc : event.ref_nuc;
//   mismatch = read_nuc != N && read_nuc != event.mut_nuc.
inline int mismatch_after(int c, int mut, int ref) {
    if (c == 0) return mut != ref;
    return (1 << (c - 1)) != mut;
}
// Before the first event on the path the state is the seed set (initial_filter.cpp:118-123):
// a mismatch iff the read carries a non-N mutation at the position.
inline int mismatch_seed(int c) { return c != 0; }

inline uint32_t pack4(const int* d) {
    return (uint32_t)(uint8_t)(int8_t)d[0] | ((uint32_t)(uint8_t)(int8_t)d[1] << 8) |
           ((uint32_t)(uint8_t)(int8_t)d[2] << 16) | ((uint32_t)(uint8_t)(int8_t)d[3] << 24);
}

}  // namespace

std::string build_euler_stripes(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off,
                                const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc,
                                int32_t genome_size, int32_t stripe_width, EulerStripes& out) {
    if (n_nodes < 1) return "arena has no nodes";
    if (n_nodes >= (1 << 28)) return "arena too large (n_nodes must be < 2^28)";
    if (genome_size < NUM_RANGE_BINS) return "genome_size must be >= 50";
    if (stripe_width < 1) return "stripe_width must be >= 1";
    if (parent[0] != -1) return "parent[0] must be -1";
    if (mut_off[0] != 0) return "mut_off[0] must be 0";
    const int64_t n_mut = mut_off[n_nodes];

    // subtree ends from the preorder parent array (validates parent[v] < v)
    std::vector<int32_t> sub_end(n_nodes);
    {
        std::vector<int32_t> size(n_nodes, 1);
        for (int32_t v = n_nodes - 1; v >= 1; --v) {
            int32_t p = parent[v];
            if (p < 0 || p >= v) return "parent[v] must satisfy 0 <= parent[v] < v (preorder)";
            size[p] += size[v];
        }
        for (int32_t v = 0; v < n_nodes; ++v) sub_end[v] = v + size[v];
        // preorder validity: v must lie inside its parent's interval, which the sizes above
        // only guarantee if the numbering really is a DFS order.
        for (int32_t v = 1; v < n_nodes; ++v)
            if (sub_end[v] > sub_end[parent[v]]) return "node numbering is not a preorder of the tree";
    }

    auto _t0 = std::chrono::steady_clock::now(); auto _lap=[&](const char* w){ if(getenv("WEPP_TIMING")){auto t=std::chrono::steady_clock::now(); fprintf(stderr,"[flatten] %-20s %.1f ms\n", w, std::chrono::duration<double,std::milli>(t-_t0).count()); _t0=t;} };
    // ---- events grouped by genome position, node order inside a position: a parallel stable counting sort -----------
    // The nearest ancestor event of an event — the "state before it" — only involves events at the SAME position, so
    // positions are independent: one sequential DFS over the tree becomes a stack walk per position, and the host
    // threads take stripes of positions.  Every event is copied next to its position's other events together with what
    // the walk needs of its node (index, subtree end, alleles): the walks then read memory front to back.
    struct PosEvent {
        int64_t k;          // event index (emission rank)
        int32_t v, v_end;   // node, subtree end
        uint8_t nuc, ref;
    };
    if (n_mut >= (1ll << 33)) return "too many mutation events";
    const int nt_all = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    const int nt_sort = n_mut < (1 << 16) ? 1 : nt_all;
    std::vector<std::string> t_error((size_t)nt_sort);
    std::vector<std::vector<int64_t>> t_count((size_t)nt_sort, std::vector<int64_t>((size_t)genome_size + 2, 0));
    auto node_range = [&](int t) { return std::make_pair((int32_t)((int64_t)n_nodes * t / nt_sort), (int32_t)((int64_t)n_nodes * (t + 1) / nt_sort)); };
    auto run_threads = [&](int nt, auto&& fn) {
        if (nt == 1) {
            fn(0);
            return;
        }
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(fn, t);
        for (auto& x : th) x.join();
    };
    run_threads(nt_sort, [&](int t) {   // validation + per-thread histogram
        const auto [v0, v1] = node_range(t);
        std::vector<int64_t>& cnt = t_count[(size_t)t];
        for (int32_t v = v0; v < v1; ++v) {
            if (mut_off[v + 1] < mut_off[v]) { t_error[(size_t)t] = "mut_off must be non-decreasing"; return; }
            for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k) {
                const int32_t p = mut_pos[k];
                if (p < 1 || p > genome_size) { t_error[(size_t)t] = "mutation position outside [1, genome_size]"; return; }
                if (mut_nuc[k] < 1 || mut_nuc[k] > 15 || mut_ref[k] < 1 || mut_ref[k] > 15) {
                    t_error[(size_t)t] = "mutation nucleotide code outside 1..15";
                    return;
                }
                ++cnt[(size_t)p];
            }
        }
    });
    for (const std::string& e : t_error)
        if (!e.empty()) return e;
    _lap("validate");
    std::vector<int64_t> pos_off((size_t)genome_size + 2, 0);
    {
        int64_t run = 0;
        for (int32_t p = 0; p <= genome_size; ++p) {   // thread t's events at p start after those of the threads before it
            pos_off[(size_t)p] = run;
            for (int t = 0; t < nt_sort; ++t) {
                const int64_t c = t_count[(size_t)t][(size_t)p];
                t_count[(size_t)t][(size_t)p] = run;
                run += c;
            }
        }
        pos_off[(size_t)genome_size + 1] = run;
    }
    std::vector<PosEvent> ev_at((size_t)n_mut);
    run_threads(nt_sort, [&](int t) {
        const auto [v0, v1] = node_range(t);
        std::vector<int64_t>& cur = t_count[(size_t)t];
        for (int32_t v = v0; v < v1; ++v)
            for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k)
                ev_at[(size_t)cur[(size_t)mut_pos[k]]++] = PosEvent{k, v, sub_end[v], mut_nuc[k], mut_ref[k]};
    });
    _lap("by position");
    const int32_t q = stripe_width;
    const int32_t n_stripes = genome_size / q + 1;
    struct Keyed {
        uint64_t order;   // key (idx << 1 | point) << 34 | emission rank (2 * event + exit): the master order within a stripe
        Entry e;
    };
    std::vector<std::vector<Entry>> stripe_entries((size_t)n_stripes);
    std::vector<int64_t> stripe_events((size_t)n_stripes, 0);
    std::vector<std::string> stripe_error((size_t)n_stripes);
    auto do_stripe = [&](int32_t s, std::vector<Keyed>& keyed, std::vector<int64_t>& stack) {
        keyed.clear();
        const int32_t p_lo = std::max(1, s * q), p_hi = std::min(genome_size, s * q + q - 1);
        for (int32_t p = p_lo; p <= p_hi; ++p) {
            const int64_t a0 = pos_off[(size_t)p], a1 = pos_off[(size_t)p + 1];
            stack.clear();
            for (int64_t j = a0; j < a1; ++j) {   // events at p in node (= preorder) order
                const PosEvent& ev = ev_at[(size_t)j];
                const int64_t k = ev.k;
                const int32_t v = ev.v;
                while (!stack.empty() && ev_at[(size_t)stack.back()].v_end <= v) stack.pop_back();
                if (!stack.empty() && ev_at[(size_t)stack.back()].v == v) {
                    stripe_error[(size_t)s] = "a node has two mutations at the same position";
                    return;
                }
                const PosEvent* pe = stack.empty() ? nullptr : &ev_at[(size_t)stack.back()];
                stack.push_back(j);
                int d[5];
                bool any = false;
                for (int c = 0; c < 5; ++c) {
                    const int after = mismatch_after(c, ev.nuc, ev.ref);
                    const int before = pe == nullptr ? mismatch_seed(c) : mismatch_after(c, pe->nuc, pe->ref);
                    d[c] = after - before;
                    any |= d[c] != 0;
                }
                if (!any) continue;  // event changes nothing for any read: drop
                ++stripe_events[(size_t)s];
                if (ev.v_end == v + 1) {
                    // leaf: one POINT entry — node v's own score is the enclosing state plus this delta;
                    // the running prefix is not changed, so no EXIT entry is needed
                    const uint32_t key = ((uint32_t)v << 1) | 1u;
                    keyed.push_back(Keyed{((uint64_t)key << 34) | (uint64_t)(2 * k),
                                          Entry{key, (uint32_t)p, pack4(d), (uint32_t)(uint8_t)(int8_t)d[4]}});
                    continue;
                }
                const uint32_t key = (uint32_t)v << 1;
                keyed.push_back(Keyed{((uint64_t)key << 34) | (uint64_t)(2 * k), Entry{key, (uint32_t)p, pack4(d), (uint32_t)(uint8_t)(int8_t)d[4]}});
                if (ev.v_end < n_nodes) {
                    int nd[5];
                    for (int c = 0; c < 5; ++c) nd[c] = -d[c];
                    const uint32_t xkey = (uint32_t)ev.v_end << 1;
                    keyed.push_back(Keyed{((uint64_t)xkey << 34) | (uint64_t)(2 * k + 1),
                                          Entry{xkey, (uint32_t)p, pack4(nd), (uint32_t)(uint8_t)(int8_t)nd[4]}});
                }
            }
        }
        // master order: by key; at one key in emission order (node, event, enter before exit) — what a stable sort by key
        // of the sequential emission gives: exits of subtrees ending at a node, its own enters, then its point entries
        std::sort(keyed.begin(), keyed.end(), [](const Keyed& x, const Keyed& y) { return x.order < y.order; });
        std::vector<Entry>& dst = stripe_entries[(size_t)s];
        dst.resize(keyed.size());
        for (size_t i = 0; i < keyed.size(); ++i) dst[i] = keyed[i].e;
    };
    {
        int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        if (n_mut < (1 << 16)) nt = 1;
        std::atomic<int32_t> next{0};
        auto worker = [&]() {
            std::vector<Keyed> keyed;
            std::vector<int64_t> stack;
            for (;;) {
                const int32_t s0 = next.fetch_add(8);
                if (s0 >= n_stripes) break;
                for (int32_t s = s0; s < std::min(n_stripes, s0 + 8); ++s) do_stripe(s, keyed, stack);
            }
        };
        if (nt == 1) {
            worker();
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back(worker);
            for (auto& x : th) x.join();
        }
    }
    _lap("stripes");
    for (const std::string& e : stripe_error)
        if (!e.empty()) return e;
    out.stripe_width = q;
    out.n_stripes = n_stripes;
    out.n_events = 0;
    out.stripe_off.assign((size_t)n_stripes + 1, 0);
    for (int32_t s = 0; s < n_stripes; ++s) {
        out.n_events += stripe_events[(size_t)s];
        out.stripe_off[(size_t)s + 1] = out.stripe_off[(size_t)s] + (int64_t)stripe_entries[(size_t)s].size();
    }
    out.entries.resize((size_t)out.stripe_off[(size_t)n_stripes]);
    {
        int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        if (out.entries.size() < (1u << 16)) nt = 1;
        std::atomic<int32_t> next{0};
        auto copier = [&]() {
            for (;;) {
                const int32_t s0 = next.fetch_add(16);
                if (s0 >= n_stripes) break;
                for (int32_t s = s0; s < std::min(n_stripes, s0 + 16); ++s) {
                    const std::vector<Entry>& src = stripe_entries[(size_t)s];
                    if (!src.empty()) std::memcpy(out.entries.data() + out.stripe_off[(size_t)s], src.data(), src.size() * sizeof(Entry));
                }
            }
        };
        if (nt == 1) {
            copier();
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back(copier);
            for (auto& x : th) x.join();
        }
    }
    _lap("copy");
    return "";
}

// Blocked-range parallel loop over [0, n) on the host (the reference uses TBB for its loaders).
template <typename F>
static void parallel_for(int64_t n, F&& body) {
    int nt = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
    if (n < (1 << 15)) nt = 1;
    if (nt == 1) {
        body(0, n, 0);
        return;
    }
    std::vector<std::thread> th;
    const int64_t step = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const int64_t a = t * step, b = std::min(n, a + step);
        if (a >= b) break;
        th.emplace_back([&, a, b, t] { body(a, b, t); });
    }
    for (auto& x : th) x.join();
}

// Lists, accumulator offsets, reads per tile and tiles from per-bucket read counts (shared by the
// host keying below and the device keying of wepp_set_reads).  `first` gets nb + 1 offsets into
// the bucket-sorted read order.
std::string finish_read_plan(const EulerStripes& es, int32_t reads_per_lane, const std::vector<int64_t>& bucket_count,
                             ReadPlan& out, std::vector<int64_t>& first) {
    for (const ListDesc& l : out.lists) {
        if (l.width > 65535) return "read window wider than 65535 bases is not supported";
        out.max_width = std::max(out.max_width, l.width);
    }
    (void)es;
    int64_t off = 0;
    for (ListDesc& l : out.lists) {
        l.off = off;
        off += l.n;
    }
    out.list_entries_total = off;
    int64_t acc = 0;
    for (BucketDesc& b : out.buckets) {
        b.acc_off = acc;
        acc += out.lists[b.list].n;
    }
    out.acc_total = acc;

    // reads per tile
    int32_t k = reads_per_lane;
    if (k != 2 && k != 4 && k != 8) {
        // selector table = width * 32 lanes * K bytes next to ~49 KB of staging and pattern tables.  The largest K
        // whose table fits the 227 KB of a B200 CTA wins even when only one CTA is then resident per SM: measured
        // on 8M nodes, K=8 at one CTA/SM beats K=4 at two by 1.6x on 490-base windows, K=4 at one beats K=2 at
        // two by 1.8x on 1.1 kb (ONT) windows (profiles/other_configs.py) — the per-entry work is shared by 2x the reads.
        k = reads_per_lane_for_width(out.max_width);
        auto tiles_for = [&](int kk) {
            int64_t t = 0;
            for (int64_t c : bucket_count) t += (c + 32 * kk - 1) / (32 * kk);
            return t;
        };
        while (k > 2 && tiles_for(k) < 592) k >>= 1;  // keep >= 2 tiles per resident CTA when reads are few
    }
    out.reads_per_tile = 32 * k;

    const size_t nb = out.buckets.size();
    first.assign(nb + 1, 0);
    for (size_t b = 0; b < nb; ++b) first[b + 1] = first[b] + bucket_count[b];
    // tiles, longest lists first (LPT order for the persistent-CTA scheduler)
    for (size_t b = 0; b < nb; ++b) {
        const int64_t c = bucket_count[b];
        for (int64_t o = 0; o < c; o += out.reads_per_tile) {
            TileDesc t;
            t.first = first[b] + o;
            t.count = (int32_t)std::min<int64_t>(out.reads_per_tile, c - o);
            t.bucket = (int32_t)b;
            out.tiles.push_back(t);
            out.scanned_entries += out.lists[out.buckets[b].list].n;
        }
        out.scanned_read_entries += c * (int64_t)out.lists[out.buckets[b].list].n;
    }
    std::stable_sort(out.tiles.begin(), out.tiles.end(), [&](const TileDesc& a, const TileDesc& b) {
        return out.lists[out.buckets[a.bucket].list].n > out.lists[out.buckets[b.bucket].list].n;
    });
    return "";
}

int32_t reads_per_lane_for_width(int32_t width) {
    return width <= PLACE_TABLE_BYTES / (32 * 8) ? 8 : (width <= PLACE_TABLE_BYTES / (32 * 4) ? 4 : 2);
}

void merge_span_chain(const std::vector<std::pair<int32_t, int64_t>>& spans, int64_t reads_per_tile, std::vector<int32_t>& target) {
    const size_t m = spans.size();
    target.assign(m, 0);
    auto tiles = [&](int64_t c) { return (c + reads_per_tile - 1) / reads_per_tile; };
    size_t group = 0;      // first span of the group being carried upwards
    int64_t carry = 0;
    for (size_t i = 0; i < m; ++i) {
        const int64_t total = carry + spans[i].second;
        bool move = false;
        if (i + 1 < m) {
            const int64_t w = spans[i].first + 1, w2 = spans[i + 1].first + 1, c2 = spans[i + 1].second;
            move = tiles(total + c2) * w2 <= tiles(total) * w + tiles(c2) * w2;
        }
        if (move) {
            carry = total;
        } else {
            for (size_t j = group; j <= i; ++j) target[j] = (int32_t)i;
            group = i + 1;
            carry = 0;
        }
    }
}

ListDesc make_list_desc(const EulerStripes& es, int32_t qs, int32_t qe) {
    const int32_t q = es.stripe_width;
    ListDesc ld;
    ld.qs = qs;
    ld.qe = qe;
    ld.b0 = qs * q;
    ld.width = (qe - qs + 1) * q;
    ld.n = (int32_t)(1 + es.stripe_off[qe + 1] - es.stripe_off[qs]);
    ld.off = 0;
    ld.pad = 0;
    return ld;
}

std::string build_read_plan(const EulerStripes& es, int32_t genome_size, int64_t n_reads, const int32_t* start,
                            const int32_t* end, const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos,
                            const uint8_t* rm_nuc, int32_t reads_per_lane, const int64_t* subset, int64_t n_subset,
                            ReadPlan& out) {
    const int32_t q = es.stripe_width;
    const int32_t bin_size = genome_size / NUM_RANGE_BINS;
    const int64_t n_sel = subset ? n_subset : n_reads;
    out = ReadPlan();
    out.n_reads = n_sel;

    // validation and window keys, in parallel
    std::vector<int32_t> bucket_of((size_t)n_sel);
    std::vector<int32_t> key_qs((size_t)n_sel), key_qe((size_t)n_sel);
    std::vector<uint8_t> key_bin((size_t)n_sel);
    std::vector<const char*> errs(64, nullptr);
    std::vector<int64_t> n_mut(64, 0);
    parallel_for(n_sel, [&](int64_t a, int64_t b, int t) {
        for (int64_t i = a; i < b; ++i) {
            const int64_t r = subset ? subset[i] : i;
            if (r < 0 || r >= n_reads) { errs[t] = "read index out of range"; return; }
            const int32_t s = start[r], e = end[r];
            if (s < 1 || s > genome_size || e > genome_size || e < s - 1) {
                errs[t] = read_plan_error(RP_ERR_WINDOW);
                return;
            }
            if (degree[r] < 0) { errs[t] = read_plan_error(RP_ERR_DEGREE); return; }
            if (rm_off[r + 1] < rm_off[r]) { errs[t] = read_plan_error(RP_ERR_OFFSETS); return; }
            int32_t prev = s - 1;
            for (int64_t k = rm_off[r]; k < rm_off[r + 1]; ++k) {
                if (rm_pos[k] <= prev || rm_pos[k] > e) {
                    errs[t] = read_plan_error(RP_ERR_MUT_ORDER);
                    return;
                }
                prev = rm_pos[k];
                const uint8_t c = rm_nuc[k];
                if (!(c == 1 || c == 2 || c == 4 || c == 8 || c == 15)) {
                    errs[t] = read_plan_error(RP_ERR_MUT_CODE);
                    return;
                }
            }
            n_mut[t] += rm_off[r + 1] - rm_off[r];
            key_qs[i] = s / q;
            key_qe[i] = std::max(e, s) / q;
            key_bin[i] = (uint8_t)std::min(s / bin_size, NUM_RANGE_BINS - 1);
        }
    });
    for (const char* e : errs)
        if (e) return e;
    for (int64_t m : n_mut) out.n_read_muts += m;

    // bucket coarsening (merge_span_chain): per (first stripe, count bin) the spans that occur, then every
    // read's qe becomes its group's
    if (n_sel > 0 && !(std::getenv("WEPP_NO_BUCKET_MERGE") && std::atoi(std::getenv("WEPP_NO_BUCKET_MERGE")) != 0)) {
        int32_t max_span = 0;
        for (int64_t i = 0; i < n_sel; ++i) max_span = std::max(max_span, key_qe[i] - key_qs[i]);
        const int64_t tile = 32 * (int64_t)((reads_per_lane == 2 || reads_per_lane == 4 || reads_per_lane == 8)
                                                ? reads_per_lane : reads_per_lane_for_width((max_span + 1) * q));
        const uint64_t S = (uint64_t)max_span + 1;
        std::vector<uint64_t> keys((size_t)n_sel);
        for (int64_t i = 0; i < n_sel; ++i)
            keys[(size_t)i] = (((uint64_t)key_qs[i] * NUM_RANGE_BINS + key_bin[i]) * S) + (uint64_t)(key_qe[i] - key_qs[i]);
        std::vector<uint64_t> sorted(keys);
        std::sort(sorted.begin(), sorted.end());
        std::vector<uint64_t> uniq;          // distinct keys, ascending
        std::vector<int32_t> uniq_target;    // the span each one is served by
        std::vector<std::pair<int32_t, int64_t>> chain;
        std::vector<int32_t> tgt;
        for (size_t a = 0; a < sorted.size();) {
            const uint64_t head = sorted[a] / S;
            chain.clear();
            size_t b = a;
            while (b < sorted.size() && sorted[b] / S == head) {
                size_t c = b;
                while (c < sorted.size() && sorted[c] == sorted[b]) ++c;
                chain.emplace_back((int32_t)(sorted[b] % S), (int64_t)(c - b));
                uniq.push_back(sorted[b]);
                b = c;
            }
            merge_span_chain(chain, tile, tgt);
            for (size_t j = 0; j < chain.size(); ++j) uniq_target.push_back(chain[(size_t)tgt[j]].first);
            a = b;
        }
        for (int64_t i = 0; i < n_sel; ++i) {
            const size_t u = (size_t)(std::lower_bound(uniq.begin(), uniq.end(), keys[(size_t)i]) - uniq.begin());
            key_qe[i] = key_qs[i] + uniq_target[u];
        }
    }

    // list and bucket ids in first-appearance order (flat tables: one short vector of (qe, list) per qs)
    std::vector<std::vector<std::pair<int32_t, int32_t>>> lists_of_qs((size_t)es.n_stripes + 1);
    std::vector<int32_t> bucket_id;   // [list * NUM_RANGE_BINS + bin]
    std::vector<int64_t> bucket_count;
    for (int64_t i = 0; i < n_sel; ++i) {
        const int32_t qs = key_qs[i], qe = key_qe[i];
        int32_t l = -1;
        for (const auto& pr : lists_of_qs[qs])
            if (pr.first == qe) {
                l = pr.second;
                break;
            }
        if (l < 0) {
            l = (int32_t)out.lists.size();
            lists_of_qs[qs].emplace_back(qe, l);
            out.lists.push_back(make_list_desc(es, qs, qe));
            bucket_id.resize(out.lists.size() * NUM_RANGE_BINS, -1);
        }
        int32_t& b = bucket_id[(size_t)l * NUM_RANGE_BINS + key_bin[i]];
        if (b < 0) {
            b = (int32_t)out.buckets.size();
            out.buckets.push_back(BucketDesc{0, l, (int32_t)key_bin[i]});
            bucket_count.push_back(0);
        }
        bucket_of[i] = b;
        ++bucket_count[b];
    }
    std::vector<int64_t> first;
    std::string err = finish_read_plan(es, reads_per_lane, bucket_count, out, first);
    if (!err.empty()) return err;

    // stable counting sort of the selected reads by bucket: sorted position -> caller's index.  The
    // reads themselves stay where they are (the device keeps them in caller order and the placement
    // kernel goes through this permutation).
    out.perm.resize((size_t)n_sel);
    {
        std::vector<int64_t> cur(first.begin(), first.end() - 1);
        for (int64_t i = 0; i < n_sel; ++i) out.perm[cur[bucket_of[i]]++] = subset ? subset[i] : i;
    }
    return "";
}

const char* read_plan_error(int code) {
    switch (code) {
        case RP_ERR_WINDOW: return "read window must satisfy 1 <= start <= genome_size, start-1 <= end <= genome_size";
        case RP_ERR_DEGREE: return "read degree must be >= 0";
        case RP_ERR_OFFSETS: return "rm_off must be non-decreasing";
        case RP_ERR_MUT_ORDER: return "read mutations must be sorted, unique and inside [start,end]";
        case RP_ERR_MUT_CODE: return "read allele code must be one of 1,2,4,8,15";
        default: return "invalid reads";
    }
}

}  // namespace wepp
