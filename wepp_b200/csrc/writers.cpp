// writers.cpp — the files `wepp detectPeaks` leaves for Freyja, sam_generation.py and the Dashboard
// (SURVEY §8f rank 2, Appendix C).  Reference: arena::print_full_report (src/WEPP/arena.cpp:389-444),
// dump_haplotype_proportion (:446-492), dump_haplotype_uncertainty (:494-528), dump_lineage_proportion
// (:530-588), dump_read2haplotype_mapping (:590-696), resolve_unaccounted_mutations (:698-904),
// dump_haplotypes (:906-931), called in the order of pipeline.cpp:53,70-75.
// Every "EPP of a read over the selected haplotypes" (arena.cpp:614-625, :846-857) is one call of the K4
// kernel through the C ABI (wepp_rescore / wepp_rescore_reads); the host only groups and formats.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <thread>
#include <unordered_map>

#include "../../include/wepp_b200.h"
#include "pipeline.h"

namespace wepp {

namespace {

template <class F>
void parallel_for(size_t n, int n_threads, F fn) {
    n_threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, n));
    if (n_threads == 1) {
        fn((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    const size_t per = (n + (size_t)n_threads - 1) / (size_t)n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const size_t a = std::min(n, (size_t)t * per), b = std::min(n, a + per);
        if (a < b) pool.emplace_back([=]() { fn(a, b); });
    }
    for (auto& th : pool) th.join();
}

// get_mutations(T, sample), util.cpp:271-296: net root->node mutations of a MAT node (the deepest event per
// position wins, reversions to the reference dropped), here sorted by position
struct MatMut {
    int32_t pos;
    uint8_t ref, nuc;
};
std::vector<MatMut> get_mutations(const MatTree& t, int32_t node) {
    struct Ev { int32_t pos; int32_t order; int64_t k; };
    std::vector<Ev> ev;
    int32_t order = 0;
    for (int32_t u = node; u >= 0; u = t.parent[(size_t)u], ++order)
        for (int64_t k = t.mut_off[(size_t)u]; k < t.mut_off[(size_t)u + 1]; ++k) ev.push_back({t.mut_pos[(size_t)k], order, k});
    std::sort(ev.begin(), ev.end(), [](const Ev& x, const Ev& y) { return x.pos != y.pos ? x.pos < y.pos : x.order < y.order; });
    std::vector<MatMut> out;
    for (size_t i = 0; i < ev.size(); ++i) {
        if (i && ev[i].pos == ev[i - 1].pos) continue;
        const int64_t k = ev[i].k;
        if (t.mut_ref[(size_t)k] != t.mut_nuc[(size_t)k]) out.push_back({ev[i].pos, t.mut_ref[(size_t)k], t.mut_nuc[(size_t)k]});
    }
    return out;
}

// mutation_distance(node1_mutations, node2_mutations), util.cpp:188-222 (inputs sorted, positions unique)
int mutation_distance(const std::vector<MatMut>& a, const std::vector<MatMut>& b) {
    size_t i = 0, j = 0;
    int d = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i].pos == b[j].pos) {
            d += a[i].nuc != b[j].nuc;
            ++i;
            ++j;
        } else if (a[i].pos < b[j].pos) {
            ++d;
            ++i;
        } else {
            ++d;
            ++j;
        }
    }
    return d + (int)(a.size() - i) + (int)(b.size() - j);
}

// the lineage label of a haplotype (arena.cpp:408-431, :461-484, :546-569): the first non-empty annotation at
// clade_idx on the root path of its source MAT node, then "/<clade>" for every other MAT node folded into it
bool lineage_enabled(const Pipeline& p) {
    if (p.ds.clade_idx() < 0) return false;
    const int max_idx = (int)p.mat.clade[0].size() - 1;
    if (p.ds.clade_idx() > max_idx) {
        std::fprintf(stderr, "\n\nERROR: CLADE_IDX = %d exceeds Max CLADE_IDX = %d in the MAT!!!\n\n", p.ds.clade_idx(), max_idx);
        return false;
    }
    return true;
}
std::string lineage_name(const Pipeline& p, int32_t hap) {
    const size_t ci = (size_t)p.ds.clade_idx();
    auto clade_of = [&](int32_t u) -> const std::string& {
        static const std::string empty;
        const auto& c = p.mat.clade[(size_t)u];
        return ci < c.size() ? c[ci] : empty;
    };
    std::string name;
    const int64_t m0 = p.arena.map_off[(size_t)hap], m1 = p.arena.map_off[(size_t)hap + 1];
    for (int32_t u = p.arena.map_nodes[(size_t)m0]; u >= 0; u = p.mat.parent[(size_t)u])
        if (!clade_of(u).empty()) {
            name = clade_of(u);
            break;
        }
    for (int64_t k = m0 + 1; k < m1; ++k) {
        const std::string& c = clade_of(p.arena.map_nodes[(size_t)k]);
        if (!c.empty()) name += "/" + c;
    }
    return name;
}

std::string write_file(const std::string& path, const std::string& body) {
    std::ofstream f(path, std::ios::binary);
    if (!f) return "cannot write " + path;
    f << body;
    return f ? "" : "cannot write " + path;
}

// EPP sets of `n_reads` reads over the haplotypes of `abundance` (in that order): CSR of positions in the list
std::string epp_over(Pipeline& p, const std::vector<int32_t>& cand, bool handle_reads, int64_t n_reads, const int32_t* start,
                     const int32_t* end, const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc,
                     std::vector<int64_t>& am_off, std::vector<int32_t>& am_idx) {
    am_off.assign((size_t)n_reads + 1, 0);
    am_idx.clear();
    if (cand.empty() || n_reads == 0) return "";
    std::vector<int32_t> min_dist((size_t)n_reads);
    auto call = [&](int32_t* idx, int64_t cap) {
        return handle_reads ? wepp_rescore(p.h, (int32_t)cand.size(), cand.data(), min_dist.data(), nullptr, am_off.data(), idx, cap)
                            : wepp_rescore_reads(p.h, n_reads, start, end, rm_off, rm_pos, rm_nuc, (int32_t)cand.size(), cand.data(),
                                                 min_dist.data(), nullptr, am_off.data(), idx, cap);
    };
    if (call(nullptr, 0) != WEPP_OK) return wepp_last_error();
    am_idx.resize((size_t)std::max<int64_t>(am_off[(size_t)n_reads], 1));
    if (call(am_idx.data(), (int64_t)am_idx.size()) != WEPP_OK) return wepp_last_error();
    am_idx.resize((size_t)am_off[(size_t)n_reads]);
    return "";
}

}  // namespace

std::string pipeline_write_results(Pipeline& p, const std::vector<Abundance>& full) {
    const Dataset& ds = p.ds;
    const ArenaHost& a = p.arena;
    const size_t G = p.reference.size();
    std::vector<int32_t> cand;
    for (const Abundance& ab : full) cand.push_back(ab.hap);
    std::string err;

    // ---- print_full_report (arena.cpp:389-444) ------------------------------------------------------
    std::cout << "----- [final report] -----" << std::endl << std::endl;
    const bool with_lineage = lineage_enabled(p);
    std::vector<std::string> lineage(full.size());
    if (with_lineage)
        for (size_t i = 0; i < full.size(); ++i) lineage[i] = lineage_name(p, full[i].hap);
    std::unordered_map<std::string, double> a_map;   // same container and insertion order as the reference
    if (with_lineage) {
        for (size_t i = 0; i < full.size(); ++i) a_map[lineage[i]] += full[i].value;
        std::cout << "--- lineage abundance " << std::endl;
        for (const auto& kv : a_map) std::printf("* lineage: %s abundance: %.6f\n", kv.first.c_str(), kv.second);
    }

    // ---- dump_haplotype_proportion (:446-492) -------------------------------------------------------
    {
        std::string body;
        for (size_t i = 0; i < full.size(); ++i) body += p.hap_id(full[i].hap) + "," + lineage[i] + "," + std::to_string(full[i].value) + "\n";
        if (!(err = write_file(ds.result("_haplotype_abundance.csv"), body)).empty()) return err;
    }
    // ---- dump_haplotype_uncertainty (:494-528): max pairwise distance of the MAT nodes folded together --
    {
        std::string body;
        for (const Abundance& ab : full) {
            const int64_t m0 = a.map_off[(size_t)ab.hap], m1 = a.map_off[(size_t)ab.hap + 1];
            const size_t k = (size_t)(m1 - m0);
            std::vector<std::vector<MatMut>> muts(k);
            parallel_for(k, p.n_threads, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i) muts[i] = get_mutations(p.mat, a.map_nodes[(size_t)m0 + i]);
            });
            std::atomic<int> max_dist{0};
            parallel_for(k, p.n_threads, [&](size_t lo, size_t hi) {
                int best = 0;
                for (size_t i = lo; i < hi; ++i)
                    for (size_t j = i + 1; j < k; ++j) best = std::max(best, mutation_distance(muts[i], muts[j]));
                int cur = max_dist.load();
                while (best > cur && !max_dist.compare_exchange_weak(cur, best)) {}
            });
            body += std::to_string(max_dist.load());
            for (size_t i = 0; i < k; ++i) body += "," + p.mat.id[(size_t)a.map_nodes[(size_t)m0 + i]];
            body += "\n";
        }
        if (!(err = write_file(ds.result("_haplotype_uncertainty.csv"), body)).empty()) return err;
    }
    // ---- dump_lineage_proportion (:530-588) -----------------------------------------------------------
    {
        std::string body;
        if (with_lineage) {
            std::vector<std::pair<std::string, double>> sorted(a_map.begin(), a_map.end());
            std::sort(sorted.begin(), sorted.end(), [](const auto& x, const auto& y) { return x.second > y.second; });
            for (const auto& lp : sorted) body += lp.first + "," + std::to_string(lp.second) + "\n";
        }
        if (!(err = write_file(ds.result("_lineage_abundance.csv"), body)).empty()) return err;
    }

    const int64_t R = p.reads.n_reads();
    // reverse_merge[read name] -> raw read names (dataset::read_reverse_merge)
    std::unordered_map<std::string_view, int64_t> rev_index;
    rev_index.reserve(p.reads.rev_key.size() * 2);
    for (size_t k = 0; k < p.reads.rev_key.size(); ++k) rev_index.emplace(p.reads.rev_key.at(k), (int64_t)k);
    auto append_raw_names = [&](int64_t read, std::string& row) {
        auto it = rev_index.find(p.reads.name.at((size_t)read));
        if (it == rev_index.end()) return;
        for (int64_t v = p.reads.rev_off[(size_t)it->second]; v < p.reads.rev_off[(size_t)it->second + 1]; ++v) {
            row += ',';
            row += p.reads.rev_val.at((size_t)v);
        }
    };

    // ---- resolve_unaccounted_mutations (:698-904) -----------------------------------------------------
    {
        struct Residual { int pos; char nuc; std::string value; };
        std::vector<Residual> residual;
        {
            std::ifstream file(ds.residual_mutations_path());
            if (file.is_open()) {
                std::string line;
                while (std::getline(file, line)) {
                    const size_t c = line.find(',');
                    const std::string part = line.substr(0, c);
                    std::string rest = c == std::string::npos ? line : line.substr(c + 1);   // npos + 1 == 0
                    std::replace(rest.begin(), rest.end(), ',', ':');
                    if (part.size() < 2) continue;
                    residual.push_back({std::atoi(part.substr(0, part.size() - 1).c_str()), part.back(), rest});
                }
                std::fprintf(stderr, "Residual mutations: %ld\n", (long)residual.size());
            } else {
                std::cerr << "Unable to open file: " << ds.residual_mutations_path() << std::endl;
            }
        }
        std::vector<std::string> keys(residual.size());
        for (size_t m = 0; m < residual.size(); ++m) keys[m] = std::to_string(residual[m].pos) + residual[m].nuc + ":" + residual[m].value;
        // per read: its copy with the residual alleles masked to N (:736-785), and which residual mutations it
        // covers / masks.  Only reads that touch a residual site are materialised.
        struct Touched {
            int64_t read;
            std::vector<int32_t> pos;
            std::vector<uint8_t> nuc;
            std::vector<int32_t> covered, masked;   // indices into `residual`
        };
        std::vector<std::vector<Touched>> per_thread((size_t)std::max(p.n_threads, 1));
        std::atomic<int> slot{0};
        parallel_for((size_t)R, p.n_threads, [&](size_t lo, size_t hi) {
            std::vector<Touched>& mine = per_thread[(size_t)slot.fetch_add(1)];
            for (size_t r = lo; r < hi; ++r) {
                const int32_t s = p.reads.start[r], e = p.reads.end[r];
                Touched t;
                bool any = false;
                for (size_t m = 0; m < residual.size(); ++m) {
                    if (residual[m].pos < s || residual[m].pos > e) continue;
                    if (!any) {
                        any = true;
                        t.read = (int64_t)r;
                        t.pos.assign(a.rm_pos.begin() + a.rm_off[r], a.rm_pos.begin() + a.rm_off[r + 1]);
                        t.nuc.assign(a.rm_nuc.begin() + a.rm_off[r], a.rm_nuc.begin() + a.rm_off[r + 1]);
                    }
                    bool site_found = false;
                    for (size_t k = 0; k < t.pos.size(); ++k) {
                        if (t.pos[k] != residual[m].pos) continue;
                        if (t.nuc[k] != 15) {
                            if (nuc_char(t.nuc[k]) == residual[m].nuc) {
                                t.nuc[k] = 15;
                                t.covered.push_back((int32_t)m);
                            }
                        } else {
                            t.masked.push_back((int32_t)m);
                        }
                        site_found = true;
                        break;
                    }
                    if (!site_found && p.reference[(size_t)residual[m].pos - 1] == residual[m].nuc) {
                        const size_t at = (size_t)(std::upper_bound(t.pos.begin(), t.pos.end(), residual[m].pos) - t.pos.begin());
                        t.pos.insert(t.pos.begin() + (long)at, residual[m].pos);
                        t.nuc.insert(t.nuc.begin() + (long)at, (uint8_t)15);
                        t.covered.push_back((int32_t)m);
                    }
                }
                if (any && (!t.covered.empty() || !t.masked.empty())) mine.push_back(std::move(t));
            }
        });
        std::vector<Touched> touched;
        for (auto& v : per_thread)
            for (auto& t : v) touched.push_back(std::move(t));
        std::sort(touched.begin(), touched.end(), [](const Touched& x, const Touched& y) { return x.read < y.read; });
        // mutations_read_map / masked_mutations_read_map
        std::vector<std::vector<int32_t>> covered_by(residual.size()), masked_by(residual.size());   // indices into touched
        for (size_t i = 0; i < touched.size(); ++i) {
            for (int32_t m : touched[i].covered) covered_by[(size_t)m].push_back((int32_t)i);
            for (int32_t m : touched[i].masked) masked_by[(size_t)m].push_back((int32_t)i);
        }
        // identical keys (a residual line repeated) share one map entry in the reference
        std::unordered_map<std::string, size_t> first_of;
        for (size_t m = 0; m < residual.size(); ++m) {
            auto ins = first_of.emplace(keys[m], m);
            if (!ins.second) {
                auto& c = covered_by[ins.first->second];
                c.insert(c.end(), covered_by[m].begin(), covered_by[m].end());
                covered_by[m].clear();
                auto& k = masked_by[ins.first->second];
                k.insert(k.end(), masked_by[m].begin(), masked_by[m].end());
                masked_by[m].clear();
            }
        }
        std::string body_reads;
        for (size_t m = 0; m < residual.size(); ++m) {
            if (first_of[keys[m]] != m || covered_by[m].empty()) continue;
            body_reads += keys[m];
            for (int32_t i : covered_by[m]) append_raw_names(touched[(size_t)i].read, body_reads);
            body_reads += "\n";
        }
        if (!(err = write_file(ds.result("_mutation_reads.csv"), body_reads)).empty()) return err;
        // EPPs of the masked copies over the selected haplotypes: one K4 launch for all of them
        std::vector<int32_t> t_start(touched.size()), t_end(touched.size());
        std::vector<int64_t> t_off(touched.size() + 1, 0);
        std::vector<int32_t> t_pos;
        std::vector<uint8_t> t_nuc;
        for (size_t i = 0; i < touched.size(); ++i) {
            t_start[i] = p.reads.start[(size_t)touched[i].read];
            t_end[i] = p.reads.end[(size_t)touched[i].read];
            t_pos.insert(t_pos.end(), touched[i].pos.begin(), touched[i].pos.end());
            t_nuc.insert(t_nuc.end(), touched[i].nuc.begin(), touched[i].nuc.end());
            t_off[i + 1] = (int64_t)t_pos.size();
        }
        std::vector<int64_t> am_off;
        std::vector<int32_t> am_idx;
        err = epp_over(p, cand, false, (int64_t)touched.size(), t_start.data(), t_end.data(), t_off.data(), t_pos.data(), t_nuc.data(), am_off, am_idx);
        if (!err.empty()) return err;
        std::string body_haps;
        for (size_t m = 0; m < residual.size(); ++m) {
            if (first_of[keys[m]] != m || (covered_by[m].empty() && masked_by[m].empty())) continue;
            std::vector<int64_t> count(cand.size(), 0);
            std::vector<uint8_t> seen(cand.size(), 0);
            auto add = [&](int32_t i) {
                for (int64_t k = am_off[(size_t)i]; k < am_off[(size_t)i + 1]; ++k) {
                    count[(size_t)am_idx[(size_t)k]] += p.reads.degree[(size_t)touched[(size_t)i].read];
                    seen[(size_t)am_idx[(size_t)k]] = 1;
                }
            };
            for (int32_t i : covered_by[m]) add(i);
            for (int32_t i : masked_by[m]) add(i);
            int64_t max_reads = 0;
            bool any = false;
            for (size_t c = 0; c < cand.size(); ++c)
                if (seen[c]) {
                    max_reads = any ? std::max(max_reads, count[c]) : count[c];
                    any = true;
                }
            if (!any) {
                std::fprintf(stderr, "\nThere are ZERO reads mapping to the selected peaks\n\n");
                max_reads = 0;
            }
            body_haps += keys[m];
            for (size_t c = 0; c < cand.size(); ++c)
                if (seen[c] && count[c] == max_reads) body_haps += "," + p.hap_id(cand[c]);
            body_haps += "\n";
        }
        if (!(err = write_file(ds.result("_mutation_haplotypes.csv"), body_haps)).empty()) return err;
    }

    // ---- dump_haplotypes (:906-931) -------------------------------------------------------------------
    {
        std::ofstream tsv(ds.result("_haplotypes.tsv"), std::ios::binary);
        if (!tsv) return "cannot write " + ds.result("_haplotypes.tsv");
        const std::string quals(G, '?');
        for (const Abundance& ab : full) {
            const std::vector<MatMut> muts = get_mutations(p.mat, a.source[(size_t)ab.hap]);
            int start_idx = 1;
            std::string md = "MD:Z:", seq = p.reference;
            for (const MatMut& m : muts) {
                if (m.pos < 1 || (size_t)m.pos > G) continue;
                seq[(size_t)m.pos - 1] = nuc_char(m.nuc);
                md += std::to_string(m.pos - start_idx) + nuc_char(m.ref);
                start_idx = m.pos + 1;
            }
            if (seq.size() - (size_t)start_idx + 1) md += std::to_string(seq.size() - (size_t)start_idx + 1);
            tsv << p.hap_id(ab.hap) << "\t0\t" << p.ref_name << "\t1\t60\t" << G << "M\t*\t0\t0\t" << seq << "\t" << quals << "\t" << md << "\tRG:Z:group\n";
        }
    }

    // ---- dump_read2haplotype_mapping (:590-696) -------------------------------------------------------
    {
        std::vector<int64_t> am_off;
        std::vector<int32_t> am_idx;
        err = epp_over(p, cand, true, R, nullptr, nullptr, nullptr, nullptr, nullptr, am_off, am_idx);
        if (!err.empty()) return err;
        // reads of every haplotype, in read order
        std::vector<int64_t> h_off(cand.size() + 1, 0);
        for (int32_t c : am_idx) ++h_off[(size_t)c + 1];
        for (size_t c = 0; c < cand.size(); ++c) h_off[c + 1] += h_off[c];
        std::vector<int64_t> h_reads((size_t)h_off[cand.size()]);
        {
            std::vector<int64_t> cur(h_off.begin(), h_off.end() - 1);
            for (int64_t r = 0; r < R; ++r)
                for (int64_t k = am_off[(size_t)r]; k < am_off[(size_t)r + 1]; ++k) h_reads[(size_t)cur[(size_t)am_idx[(size_t)k]]++] = r;
        }
        // two haplotypes with one id (it cannot happen: ids are MAT node names) would share a map entry
        std::vector<std::string> rows(cand.size()), cov_rows(cand.size());
        parallel_for(cand.size(), p.n_threads, [&](size_t lo, size_t hi) {
            std::vector<int32_t> cov(G + 2);
            for (size_t c = lo; c < hi; ++c) {
                // coverage (:637-665): positions spanned by an EPP read with a base that is not N
                std::fill(cov.begin(), cov.end(), 0);
                std::string& row = rows[c];
                if (h_off[c + 1] > h_off[c]) row = p.hap_id(cand[c]);
                for (int64_t k = h_off[c]; k < h_off[c + 1]; ++k) {
                    const int64_t r = h_reads[(size_t)k];
                    append_raw_names(r, row);
                    const int32_t s = std::max(p.reads.start[(size_t)r], 1), e = std::min<int32_t>(p.reads.end[(size_t)r], (int32_t)G);
                    if (e >= s) {
                        ++cov[(size_t)s];
                        --cov[(size_t)e + 1];
                    }
                }
                int64_t run = 0, covered = 0;
                std::vector<int32_t> n_at;   // positions where reads carry N: covered only if some read has a base there
                for (int64_t k = h_off[c]; k < h_off[c + 1]; ++k) {
                    const int64_t r = h_reads[(size_t)k];
                    for (int64_t q = a.rm_off[(size_t)r]; q < a.rm_off[(size_t)r + 1]; ++q)
                        if (a.rm_nuc[(size_t)q] == 15) n_at.push_back(a.rm_pos[(size_t)q]);
                }
                std::sort(n_at.begin(), n_at.end());
                size_t q = 0;
                for (size_t pos = 1; pos <= G; ++pos) {
                    run += cov[pos];
                    int64_t ns = 0;
                    while (q < n_at.size() && (size_t)n_at[q] == pos) {
                        ++ns;
                        ++q;
                    }
                    covered += (run - ns) > 0;
                }
                if (!row.empty()) row += "\n";
                cov_rows[c] = p.hap_id(cand[c]) + "," + std::to_string((double)covered / (double)G) + "\n";
            }
        });
        {
            std::ofstream csv(ds.result("_haplotype_reads.csv"), std::ios::binary);
            if (!csv) return "cannot write " + ds.result("_haplotype_reads.csv");
            for (const std::string& row : rows) csv << row;
        }
        {
            std::ofstream csv(ds.result("_haplotype_coverage.csv"), std::ios::binary);
            if (!csv) return "cannot write " + ds.result("_haplotype_coverage.csv");
            for (const std::string& row : cov_rows) csv << row;
        }
        // the script that turns these files into the Dashboard's BAMs (:692-695)
        const std::string command = "python " + ds.o.working_directory + "/src/WEPP/sam_generation.py '" + ds.results_directory() + "' '" +
                                    ds.intermediate_directory() + "' " + ds.o.file_prefix;
        if (std::system(command.c_str())) std::fprintf(stderr, "\nCannot run sam_generation.py\n");
    }
    return "";
}

}  // namespace wepp
