// host_prep.cpp — see host_prep.h.
#include "host_prep.h"

#include <cstdlib>
#include <algorithm>
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

    // nearest-ancestor event per position, maintained along the DFS
    std::vector<int64_t> last_ev((size_t)genome_size + 1, -1);
    std::vector<int64_t> prev_ev((size_t)n_mut, -1);
    {
        std::vector<int32_t> stack;
        for (int32_t v = 0; v < n_nodes; ++v) {
            while (!stack.empty() && sub_end[stack.back()] <= v) {
                int32_t u = stack.back();
                stack.pop_back();
                for (int64_t k = mut_off[u + 1] - 1; k >= mut_off[u]; --k) last_ev[mut_pos[k]] = prev_ev[k];
            }
            if (mut_off[v + 1] < mut_off[v]) return "mut_off must be non-decreasing";
            for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k) {
                int32_t p = mut_pos[k];
                if (p < 1 || p > genome_size) return "mutation position outside [1, genome_size]";
                if (mut_nuc[k] < 1 || mut_nuc[k] > 15 || mut_ref[k] < 1 || mut_ref[k] > 15)
                    return "mutation nucleotide code outside 1..15";
                if (last_ev[p] >= mut_off[v]) return "a node has two mutations at the same position";
                prev_ev[k] = last_ev[p];
                last_ev[p] = k;
            }
            stack.push_back(v);
        }
    }

    // entries (enter + exit), then counting sort by (stripe, idx)
    const int32_t q = stripe_width;
    const int32_t n_stripes = genome_size / q + 1;
    std::vector<Entry> raw;
    raw.reserve((size_t)n_mut * 2);
    int64_t n_events = 0;
    for (int32_t v = 0; v < n_nodes; ++v) {
        for (int64_t k = mut_off[v]; k < mut_off[v + 1]; ++k) {
            int d[5];
            bool any = false;
            const int64_t pk = prev_ev[k];
            for (int c = 0; c < 5; ++c) {
                int after = mismatch_after(c, mut_nuc[k], mut_ref[k]);
                int before = pk < 0 ? mismatch_seed(c) : mismatch_after(c, mut_nuc[pk], mut_ref[pk]);
                d[c] = after - before;
                any |= d[c] != 0;
            }
            if (!any) continue;  // event changes nothing for any read: drop
            ++n_events;
            if (sub_end[v] == v + 1) {
                // leaf: one POINT entry — node v's own score is the enclosing state plus this delta;
                // the running prefix is not changed, so no EXIT entry is needed
                raw.push_back(Entry{((uint32_t)v << 1) | 1u, (uint32_t)mut_pos[k], pack4(d), (uint32_t)(uint8_t)(int8_t)d[4]});
                continue;
            }
            raw.push_back(Entry{(uint32_t)v << 1, (uint32_t)mut_pos[k], pack4(d), (uint32_t)(uint8_t)(int8_t)d[4]});
            if (sub_end[v] < n_nodes) {
                int nd[5];
                for (int c = 0; c < 5; ++c) nd[c] = -d[c];
                raw.push_back(Entry{(uint32_t)sub_end[v] << 1, (uint32_t)mut_pos[k], pack4(nd), (uint32_t)(uint8_t)(int8_t)nd[4]});
            }
        }
    }
    // pass 1: stable counting sort by key = (idx << 1 | point): at one preorder index the
    // boundary entries (exits of subtrees ending here, the node's own enters) precede its point entries
    std::vector<Entry> by_idx(raw.size());
    {
        std::vector<int64_t> cnt((size_t)2 * n_nodes + 1, 0);
        for (const Entry& e : raw) ++cnt[e.x + 1];
        for (int64_t v = 0; v < (int64_t)2 * n_nodes; ++v) cnt[v + 1] += cnt[v];
        for (const Entry& e : raw) by_idx[cnt[e.x]++] = e;
    }
    raw.clear();
    raw.shrink_to_fit();
    // pass 2: stable counting sort by stripe
    out.stripe_width = q;
    out.n_stripes = n_stripes;
    out.n_events = n_events;
    out.stripe_off.assign((size_t)n_stripes + 1, 0);
    for (const Entry& e : by_idx) ++out.stripe_off[e.y / q + 1];
    for (int32_t s = 0; s < n_stripes; ++s) out.stripe_off[s + 1] += out.stripe_off[s];
    out.entries.resize(by_idx.size());
    {
        std::vector<int64_t> cur(out.stripe_off.begin(), out.stripe_off.end() - 1);
        for (const Entry& e : by_idx) out.entries[cur[e.y / q]++] = e;
    }
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
