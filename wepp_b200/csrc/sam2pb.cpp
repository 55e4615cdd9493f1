// sam2pb.cpp — `wepp sam2PB`: SAM alignments -> quality-masked, collapsed reads -> <P>_reads.pb
// (reference src/WEPP/sam2pb.cpp:54-109 sam2PB, :155-278 add_reads (CIGAR walk, Phred
// masking, allele counts), :280-331 read_correction (depth / allele-frequency masking), :333-361
// merge_duplicates, :363-455 subsample, :457-477 build, :111-151 dump_proto; sam.proto:4-18).
//
// Same algorithm, restructured: lines are parsed by a thread pool into per-thread buffers that are
// concatenated in FILE order, so the output does not depend on scheduling (the reference appends in
// completion order and sorts with an unstable sort; which raw read names a collapsed read is named after is
// then thread-dependent — here it is always the first in the file).  Two reference quirks are decided on
// knowingly:
//   * header / unmapped lines are skipped with `continue`; the reference `return`s (sam2pb.cpp:165-168) and
//     so also drops every later read of the same TBB sub-range — an accident of chunking;
//   * the subsample (only when there are more reads than --max-reads) draws from std::random_device like the
//     reference (:373-374) unless WEPP_SEED is set.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <mutex>
#include <numeric>
#include <random>
#include <sstream>
#include <thread>

#include "pbwire.h"
#include "pipeline.h"

namespace wepp {

namespace {

const std::string GENOME_STRING{"ACGTN_"};                  // sam2pb.hpp:10
using SubTable = std::vector<std::array<int, 6>>;            // sam2pb.hpp:11
constexpr double SCORE_EPSILON = 1e-9;                       // config.hpp:15
constexpr int SUBSAMPLE_ITERS = 1000;                        // config.hpp:6

struct SamRead {   // sam2pb.hpp:14-50
    std::string raw_name;
    int start_idx;   // zero-based
    int degree;
    std::string aligned;
    std::string degree_name() const {
        return raw_name + "_READ_" + std::to_string(start_idx + 1) + "_" + std::to_string(start_idx + 1 + (long)aligned.size() - 1) + "_" +
               std::to_string(degree);
    }
    bool operator<(const SamRead& r) const {
        if (start_idx != r.start_idx) return start_idx < r.start_idx;
        if (aligned.size() != r.aligned.size()) return aligned.size() < r.aligned.size();
        return aligned < r.aligned;
    }
    bool operator==(const SamRead& r) const { return start_idx == r.start_idx && aligned == r.aligned; }
};

// one SAM line -> aligned string (sam2pb.cpp:158-257).  Returns false for header / unmapped / short lines.
bool parse_line(const std::string& line, int phred_cutoff, SamRead& out) {
    std::vector<std::string> tok;
    {
        std::istringstream ss(line);
        std::string w;
        while (ss >> w) tok.push_back(std::move(w));
    }
    if (tok.empty() || tok[0][0] == '@' || tok.size() < 11) return false;
    if (std::atoi(tok[1].c_str()) & 4) return false;
    const int start_idx = std::atoi(tok[3].c_str());
    const std::string& seq = tok[9];
    const std::string& phred = tok[10];
    const std::string& cigar = tok[5];
    std::string build;
    size_t seq_idx = 0;
    // chunks "<digits><letter>" (the reference's regex \d+[A-Za-z]; anything else, e.g. '=', is not a chunk)
    for (size_t i = 0; i < cigar.size();) {
        if (cigar[i] < '0' || cigar[i] > '9') { ++i; continue; }
        size_t j = i;
        long len = 0;
        while (j < cigar.size() && cigar[j] >= '0' && cigar[j] <= '9') len = len * 10 + (cigar[j++] - '0');
        if (j >= cigar.size()) break;
        const char op = cigar[j];
        i = j + 1;
        if (!((op >= 'A' && op <= 'Z') || (op >= 'a' && op <= 'z')) || len <= 0) continue;
        switch (op) {
            case 'I': seq_idx += (size_t)len; break;
            case 'D': build.append((size_t)len, '_'); break;
            case 'N': build.append((size_t)len, 'N'); seq_idx += (size_t)len; break;
            case 'H': break;
            case 'S': seq_idx += (size_t)len; break;
            default:
                for (long k = 0; k < len; ++k, ++seq_idx) {
                    const int q = (seq_idx < phred.size() ? (int)phred[seq_idx] : 0) - 33;
                    char c = q < phred_cutoff ? 'N' : (seq_idx < seq.size() ? seq[seq_idx] : '\0');
                    if (c == '\0' || GENOME_STRING.find(c) == std::string::npos) c = 'N';   // non ACGTN -> N
                    build.push_back(c);
                }
        }
    }
    out.raw_name = std::move(tok[0]);
    out.start_idx = start_idx - 1;
    out.degree = 1;
    out.aligned = std::move(build);
    return true;
}

template <class F>
void run_threads(int n_threads, F fn) {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back([=]() { fn(t); });
    for (auto& th : pool) th.join();
}

void count_alleles(const std::vector<SamRead>& reads, size_t lo, size_t hi, SubTable& table) {
    for (size_t r = lo; r < hi; ++r) {
        const SamRead& rd = reads[r];
        for (size_t i = 0; i < rd.aligned.size(); ++i) {
            const char c = rd.aligned[i];
            if (c == 'N') continue;   // sam2pb.cpp:266-268
            const long pos = (long)i + rd.start_idx;
            if (pos < 0 || (size_t)pos >= table.size()) continue;
            table[(size_t)pos][GENOME_STRING.find(c)] += rd.degree;
        }
    }
}

// sam::subsample, sam2pb.cpp:363-455: of SUBSAMPLE_ITERS random subsets the one whose allele-frequency vector
// has the smallest KL divergence from the full sample's
void subsample(std::vector<SamRead>& reads, const SubTable& collapsed, size_t genome, int subsampled_reads, int n_threads) {
    if ((long)reads.size() <= (long)subsampled_reads) return;
    auto frequency_vector = [&](const SubTable& t) {
        std::vector<double> p;
        p.reserve(genome * 6);
        for (size_t i = 0; i < genome; ++i) {
            const int sum = std::accumulate(t[i].begin(), t[i].end(), 0);
            for (size_t j = 0; j < 6; ++j) p.push_back(sum ? (double)t[i][j] / sum : 0.0);
        }
        const double tot = std::accumulate(p.begin(), p.end(), 0.0);
        for (double& x : p) x /= tot;
        return p;
    };
    const std::vector<double> p = frequency_vector(collapsed);
    std::mt19937 g;
    if (const char* s = std::getenv("WEPP_SEED")) g.seed((unsigned)std::strtoul(s, nullptr, 10));
    else g.seed(std::random_device{}());
    std::mutex mu;
    double best_score = std::numeric_limits<double>::max();
    std::vector<int> best_set;
    std::atomic<int> next{0};
    run_threads(n_threads, [&](int) {
        std::vector<int> index_set(reads.size());
        for (;;) {
            if (next.fetch_add(1) >= SUBSAMPLE_ITERS) break;
            std::iota(index_set.begin(), index_set.end(), 0);
            {
                std::lock_guard<std::mutex> lock(mu);
                std::shuffle(index_set.begin(), index_set.end(), g);
            }
            SubTable af(genome);
            for (int j = 0; j < subsampled_reads; ++j) {
                const SamRead& rd = reads[(size_t)index_set[(size_t)j]];
                for (size_t k = 0; k < rd.aligned.size(); ++k) {
                    const long pos = (long)k + rd.start_idx;
                    if (pos < 0 || (size_t)pos >= genome) continue;
                    af[(size_t)pos][GENOME_STRING.find(rd.aligned[k])] += rd.degree;   // N counted here (:414-421)
                }
            }
            const std::vector<double> q = frequency_vector(af);
            double divergence = 0;
            for (size_t j = 0; j < p.size(); ++j) {
                if (p[j] == 0) continue;
                divergence += p[j] * (std::log(p[j]) - std::log(std::max(q[j], 1e-10)));
            }
            std::lock_guard<std::mutex> lock(mu);
            if (divergence < best_score) {
                best_score = divergence;
                best_set.assign(index_set.begin(), index_set.begin() + subsampled_reads);
            }
        }
    });
    std::vector<SamRead> chosen;
    chosen.reserve(best_set.size());
    for (int i : best_set) chosen.push_back(std::move(reads[(size_t)i]));
    reads = std::move(chosen);
}

}  // namespace

int sam2pb(const Dataset& ds) {
    const auto t0 = std::chrono::steady_clock::now();
    const int phred_cutoff = (int)ds.o.min_phred, depth_cutoff = (int)ds.o.min_depth;
    const double freq_cutoff = ds.min_af();
    const int n_threads = (int)std::max(1u, ds.o.threads);
    std::string ref_name, ref;
    std::string err = load_fasta(ds.ref_path(), ref_name, ref);
    if (!err.empty()) {
        std::cerr << "Error: " << err << std::endl;
        return 1;
    }
    std::vector<std::string> lines;
    {
        std::ifstream f(ds.sam_path());
        std::string s;
        while (std::getline(f, s)) lines.emplace_back(std::move(s));
    }
    // ---- add_reads: parse in parallel, keep file order -------------------------------------------------
    std::vector<std::vector<SamRead>> part((size_t)n_threads);
    run_threads(n_threads, [&](int t) {
        const size_t per = (lines.size() + (size_t)n_threads - 1) / (size_t)n_threads;
        const size_t lo = std::min(lines.size(), (size_t)t * per), hi = std::min(lines.size(), lo + per);
        for (size_t i = lo; i < hi; ++i) {
            SamRead r;
            if (parse_line(lines[i], phred_cutoff, r)) part[(size_t)t].push_back(std::move(r));
        }
    });
    std::vector<SamRead> reads;
    for (auto& v : part)
        for (auto& r : v) reads.push_back(std::move(r));
    part.clear();
    lines.clear();
    SubTable frequency(ref.size());
    {
        std::vector<SubTable> local((size_t)n_threads, SubTable(ref.size()));
        run_threads(n_threads, [&](int t) {
            const size_t per = (reads.size() + (size_t)n_threads - 1) / (size_t)n_threads;
            const size_t lo = std::min(reads.size(), (size_t)t * per), hi = std::min(reads.size(), lo + per);
            count_alleles(reads, lo, hi, local[(size_t)t]);
        });
        for (const SubTable& l : local)
            for (size_t i = 0; i < ref.size(); ++i)
                for (int j = 0; j < 6; ++j) frequency[i][(size_t)j] += l[i][(size_t)j];
    }
    // ---- build (sam2pb.cpp:457-477) ---------------------------------------------------------------------
    const SubTable& collapsed = frequency;
    subsample(reads, collapsed, ref.size(), (int)std::min<uint32_t>(ds.o.max_reads, 0x7FFFFFFFu), n_threads);
    // read_correction (:280-331): mask bases at low-depth sites and alleles below the frequency cut-off; gaps -> N
    {
        std::vector<int> total(ref.size());
        for (size_t i = 0; i < ref.size(); ++i) total[i] = std::accumulate(collapsed[i].begin(), collapsed[i].end(), 0);
        run_threads(n_threads, [&](int t) {
            const size_t per = (reads.size() + (size_t)n_threads - 1) / (size_t)n_threads;
            const size_t lo = std::min(reads.size(), (size_t)t * per), hi = std::min(reads.size(), lo + per);
            for (size_t r = lo; r < hi; ++r) {
                std::string& al = reads[r].aligned;
                for (size_t j = 0; j < al.size(); ++j) {
                    const long indx = (long)j + reads[r].start_idx;
                    if (indx >= 0 && (size_t)indx < ref.size()) {
                        const size_t curr = GENOME_STRING.find(al[j]);
                        if (depth_cutoff > total[(size_t)indx]) al[j] = 'N';
                        else if (freq_cutoff - (double)collapsed[(size_t)indx][curr] / total[(size_t)indx] > SCORE_EPSILON) al[j] = 'N';
                    }
                    if (al[j] == '_') al[j] = 'N';
                }
            }
        });
    }
    std::stable_sort(reads.begin(), reads.end());
    // merge_duplicates (:333-361)
    if (reads.empty()) {
        std::cerr << "Zero reads; likely did not find input sam file\n";
        return 1;
    }
    std::map<std::string, std::vector<std::string>> reverse_merge;
    std::vector<SamRead> merged;
    {
        size_t i = 0;
        while (i < reads.size()) {
            size_t j = i + 1;
            while (j < reads.size() && reads[j] == reads[i]) ++j;
            SamRead m = reads[i];
            m.degree = (int)(j - i);
            auto& names = reverse_merge[m.degree_name()];
            for (size_t k = i; k < j; ++k) names.push_back(reads[k].raw_name);
            merged.push_back(std::move(m));
            i = j;
        }
    }
    // dump_proto (:111-151), Sam::sam wire format, fields in number order like protobuf's own serialiser
    {
        std::ofstream out(ds.pb_path(), std::ios::out | std::ios::binary);
        if (!out) {
            std::fprintf(stderr, "ERROR: cannot write %s\n", ds.pb_path().c_str());
            return 1;
        }
        pb::Writer file;
        for (const SamRead& rd : merged) {
            pb::Writer m;
            m.str(1, rd.degree_name());
            m.int32(3, rd.start_idx + 1);
            m.int32(5, rd.degree);
            m.str(6, rd.aligned);
            file.message(1, m.out);
            if (file.out.size() > (1u << 24)) {
                out << file.out;
                file.out.clear();
            }
        }
        for (const auto& kv : reverse_merge) {
            pb::Writer m;
            m.str(1, kv.first);
            for (const std::string& s : kv.second) m.str(2, s, true);
            file.message(2, m.out);
            if (file.out.size() > (1u << 24)) {
                out << file.out;
                file.out.clear();
            }
        }
        out << file.out;
    }
    const long secs = (long)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "\nFiles generated in %ld sec \n\n", secs);
    return 0;
}

}  // namespace wepp
