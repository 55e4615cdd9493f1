// pipeline.cpp — `wepp detectPeaks`: options, loading, the GPU initial filter, the iterative Freyja post
// filter.  See pipeline.h for the reference lines each stage replaces.  The result writers are in
// writers.cpp, `wepp sam2PB` in sam2pb.cpp.
#include "pipeline.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>
#include <thread>

#include "../../include/wepp_b200.h"

namespace wepp {

namespace {

constexpr double SCORE_EPSILON = 1e-9;          // src/WEPP/config.hpp:15
constexpr int MAX_NEIGHBORS_FREYJA = 500;       // :24
constexpr int MAX_NEIGHBOR_MUTATION = 2;        // :25
constexpr int MAX_NEIGHBOR_ITERATIONS = 10;     // :27

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    long seconds() const { return (long)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - t0).count(); }
};

// ---- option table, util.cpp:145-158 ---------------------------------------------------------------
struct OptSpec {
    const char* long_name;
    char short_name;
    const char* help;
};
const OptSpec OPT_TABLE[] = {
    {"working-directory", 'w', "WEPP's working directory."},
    {"input-mat", 'i', "Input mutation-annotated tree."},
    {"dataset", 'd', "Data folder containing reads."},
    {"max-reads", 'm', "Maximum number of reads."},
    {"file-prefix", 'p', "Prefix for intermediate files."},
    {"ref-fasta", 'f', "Reference sequence."},
    {"min-af", 'a', "Allele Frequency threshold for masking errorneous alleles."},
    {"min-depth", 'c', "Depth threshold for masking low coverage sites."},
    {"min-phred", 'q', "Phred score threshold for masking low quality alleles."},
    {"min-prop", 'r', "Minimum haplotype abundance."},
    {"clade-idx", 'n', "Index used for inferring lineage proportions from haplotypes."},
    {"threads", 'T', "Number of threads to use when possible [DEFAULT uses all available cores]"},
    {"help", 'h', "Print help messages"},
};

// boost::lexical_cast<uint32_t>: digits with an optional sign, a leading '-' negates modulo 2^32
// (this is how the documented `-n -1` reaches clade_idx() < 0, dataset.hpp:34-36, arena.hpp:119-121)
bool parse_u32(const std::string& s, uint32_t& out) {
    if (s.empty()) return false;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        i = 1;
    }
    if (i >= s.size()) return false;
    uint64_t v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (uint64_t)(s[i] - '0');
        if (v > 0xFFFFFFFFull) return false;
    }
    out = neg ? (uint32_t)(0u - (uint32_t)v) : (uint32_t)v;
    return true;
}

}  // namespace

std::string usage_text() {
    std::ostringstream o;
    o << "Arguments:\n";
    for (const OptSpec& s : OPT_TABLE) {
        std::string flag = std::string("  -") + s.short_name + " [ --" + s.long_name + " ]" + (std::strcmp(s.long_name, "help") ? " arg" : "");
        if (flag.size() < 38) flag.resize(38, ' ');
        o << flag << s.help << "\n";
    }
    return o.str();
}

std::string parse_options(const std::vector<std::string>& args, Options& out) {
    auto assign = [&](const std::string& name, const std::string& v) -> std::string {
        uint32_t u = 0;
        auto want_u32 = [&](uint32_t& dst) -> std::string {
            if (!parse_u32(v, u)) return "the argument ('" + v + "') for option '--" + name + "' is invalid";
            dst = u;
            return "";
        };
        if (name == "working-directory") out.working_directory = v;
        else if (name == "input-mat") out.input_mat = v;
        else if (name == "dataset") out.dataset = v;
        else if (name == "file-prefix") out.file_prefix = v;
        else if (name == "ref-fasta") out.ref_fasta = v;
        else if (name == "min-af") out.min_af = v;
        else if (name == "min-prop") out.min_prop = v;
        else if (name == "max-reads") return want_u32(out.max_reads);
        else if (name == "min-depth") return want_u32(out.min_depth);
        else if (name == "min-phred") return want_u32(out.min_phred);
        else if (name == "clade-idx") return want_u32(out.clade_idx);
        else if (name == "threads") return want_u32(out.threads);
        return "";
    };
    if (out.threads == 0) out.threads = std::max(1u, std::thread::hardware_concurrency());
    for (size_t i = 0; i < args.size(); ++i) {
        const std::string& a = args[i];
        std::string name, value;
        bool has_value = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            const size_t eq = a.find('=');
            std::string key = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            // boost::program_options accepts unambiguous prefixes of long names
            int hits = 0;
            for (const OptSpec& s : OPT_TABLE) {
                if (key == s.long_name) { name = s.long_name; hits = 1; break; }
                if (std::string(s.long_name).compare(0, key.size(), key) == 0) { name = s.long_name; ++hits; }
            }
            if (hits == 0) return "unrecognised option '" + a + "'";
            if (hits > 1) return "option '" + a + "' is ambiguous";
            if (eq != std::string::npos) { value = a.substr(eq + 1); has_value = true; }
        } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
            for (const OptSpec& s : OPT_TABLE)
                if (s.short_name == a[1]) name = s.long_name;
            if (name.empty()) return "unrecognised option '" + a + "'";
            if (a.size() > 2) { value = a.substr(2); has_value = true; }
        } else {
            return "too many positional options have been specified on the command line";
        }
        if (name == "help") { out.help = true; continue; }
        if (!has_value) {
            if (i + 1 >= args.size()) return "the required argument for option '--" + name + "' is missing";
            value = args[++i];
        }
        std::string err = assign(name, value);
        if (!err.empty()) return err;
    }
    return "";
}

// ---- haplotype stacks -----------------------------------------------------------------------------
std::shared_ptr<const Stack> HapStacks::get(int32_t v) {
    const int s = (int)((uint32_t)v % SHARDS);
    {
        std::lock_guard<std::mutex> g(mu_[s]);
        auto it = cache_[s].find(v);
        if (it != cache_[s].end()) return it->second;
    }
    // arena.cpp:18-46: the last event per position on the root path, kept when it differs from the reference
    // allele.  Walk up from v: the first event seen at a position is the deepest one.
    struct Ev { int32_t pos; int32_t order; int64_t k; };
    std::vector<Ev> ev;
    int32_t order = 0;
    for (int32_t u = v; u >= 0; u = a_.parent[(size_t)u], ++order)
        for (int64_t k = a_.mut_off[(size_t)u]; k < a_.mut_off[(size_t)u + 1]; ++k) ev.push_back({a_.mut_pos[(size_t)k], order, k});
    std::sort(ev.begin(), ev.end(), [](const Ev& x, const Ev& y) { return x.pos != y.pos ? x.pos < y.pos : x.order < y.order; });
    auto st = std::make_shared<Stack>();
    for (size_t i = 0; i < ev.size(); ++i) {
        if (i && ev[i].pos == ev[i - 1].pos) continue;
        const int64_t k = ev[i].k;
        if (a_.mut_ref[(size_t)k] != a_.mut_nuc[(size_t)k]) st->push_back({ev[i].pos, a_.mut_ref[(size_t)k], a_.mut_nuc[(size_t)k]});
    }
    std::lock_guard<std::mutex> g(mu_[s]);
    auto ins = cache_[s].emplace(v, st);
    return ins.first->second;
}

int HapStacks::distance(int32_t a, int32_t b) {
    const auto A = get(a), B = get(b);
    const Stack& s = *A;
    const Stack& c = *B;
    size_t i = 0, j = 0;
    int m = 0;
    while (i < s.size() || j < c.size()) {
        if (i == s.size()) { m += c[j].nuc != 15; ++j; }
        else if (j == c.size()) { ++m; ++i; }
        else if (s[i].pos < c[j].pos) { ++m; ++i; }
        else if (s[i].pos > c[j].pos) { m += c[j].nuc != 15; ++j; }
        else { m += (s[i].nuc != c[j].nuc) && (c[j].nuc != 15); ++i; ++j; }
    }
    return m;
}

Pipeline::~Pipeline() {
    if (h) wepp_destroy(h);
}

bool Pipeline::score_less(int32_t l, int32_t r) const {
    const double el = full_score[(size_t)l], er = full_score[(size_t)r];
    if (std::fabs(el - er) > SCORE_EPSILON) return el > er;
    if (arena.leaf_count[(size_t)l] != arena.leaf_count[(size_t)r]) return arena.leaf_count[(size_t)l] > arena.leaf_count[(size_t)r];
    return id_rank[(size_t)l] > id_rank[(size_t)r];
}

// ---- loading: dataset::mat / masked_sites / reads + the arena constructor ----------------------------
std::string pipeline_load(Pipeline& p) {
    const Dataset& ds = p.ds;
    p.n_threads = (int)std::max(1u, ds.o.threads);
    std::string err = load_mat(ds.mat_path(), true, p.mat);
    if (!err.empty()) return err;
    p.masked = load_mask_bed(ds.mask_path());
    err = load_fasta(ds.ref_path(), p.ref_name, p.reference);
    if (!err.empty()) return "Error: " + err;
    Timer t;
    err = load_reads(ds.pb_path(), p.reference, p.reads, p.n_threads);
    if (!err.empty()) return "ERROR: " + err;
    std::printf("--- parsed %s containing %d merged reads in %ld sec\n\n", ds.pb_path().c_str(), (int)p.reads.n_reads(), t.seconds());
    err = build_arena(p.mat.n_nodes(), p.mat.parent.data(), p.mat.mut_off.data(), p.mat.mut_pos.data(), p.mat.mut_ref.data(),
                      p.mat.mut_nuc.data(), (int32_t)p.reference.size(), (int32_t)p.masked.size(), p.masked.data(),
                      p.reads.n_reads(), p.reads.start.data(), p.reads.end.data(), p.reads.rm_off.data(), p.reads.rm_pos.data(),
                      p.reads.rm_nuc.data(), p.arena);
    if (!err.empty()) return err;
    const int32_t n = (int32_t)p.arena.parent.size();
    // rank of haplotype::id in std::string order (score_comparator's last tie-break, arena.hpp:27-29)
    std::vector<int32_t> order((size_t)n);
    for (int32_t v = 0; v < n; ++v) order[(size_t)v] = v;
    {   // (ids are unique, so the order does not depend on how the sort is cut up: sorted runs on the host threads, merged pairwise)
        auto less = [&](int32_t a, int32_t b) { return p.hap_id(a) < p.hap_id(b); };
        int runs = 1;
        const size_t min_run = std::getenv("WEPP_SORT_RUN") ? (size_t)std::max(1, std::atoi(std::getenv("WEPP_SORT_RUN"))) : 65536;   // (tests shrink it)
        while (runs * 2 <= std::min(p.n_threads, 64) && (size_t)n / (size_t)(runs * 2) >= min_run) runs *= 2;
        auto cut = [&](int r) { return order.begin() + (ptrdiff_t)((int64_t)n * r / runs); };
        std::vector<std::thread> pool;
        for (int r = 0; r < runs; ++r) pool.emplace_back([&, r]() { std::sort(cut(r), cut(r + 1), less); });
        for (auto& t : pool) t.join();
        for (int w = 1; w < runs; w *= 2) {
            pool.clear();
            for (int r = 0; r + w < runs; r += 2 * w)
                pool.emplace_back([&, r, w]() { std::inplace_merge(cut(r), cut(r + w), cut(std::min(r + 2 * w, runs)), less); });
            for (auto& t : pool) t.join();
        }
    }
    p.id_rank.assign((size_t)n, 0);
    for (int32_t k = 0; k < n; ++k) p.id_rank[(size_t)order[(size_t)k]] = k;
    p.child_off.assign((size_t)n + 1, 0);
    for (int32_t v = 1; v < n; ++v) ++p.child_off[(size_t)p.arena.parent[(size_t)v] + 1];
    for (int32_t v = 0; v < n; ++v) p.child_off[(size_t)v + 1] += p.child_off[(size_t)v];
    p.child.assign((size_t)std::max(n - 1, 0), 0);
    std::vector<int64_t> cur(p.child_off.begin(), p.child_off.end() - 1);
    for (int32_t v = 1; v < n; ++v) p.child[(size_t)cur[(size_t)p.arena.parent[(size_t)v]]++] = v;
    p.stacks = std::make_unique<HapStacks>(p.arena);
    return "";
}

// ---- initial filter on the GPU (pipeline.cpp:24-41; wepp_filter::filter, initial_filter.cpp:455-506) ----
std::string pipeline_initial_filter(Pipeline& p, std::vector<int32_t>& running) {
    const ArenaHost& a = p.arena;
    const int32_t n = (int32_t)a.parent.size();
    // WEPP_DEVICES="0,1,2,3" (or WEPP_GPUS=4 for devices 0..3; WEPP_DEVICE for a single one): the initial filter —
    // cartesian_map and the peak loop — runs read-sharded on all of them (wepp_group, include/wepp_b200.h); the stages
    // after it work on candidate sets and stay on the first device
    std::vector<int32_t> devices;
    if (const char* e = std::getenv("WEPP_DEVICES")) {
        for (const char* c = e; *c;) {
            char* end = nullptr;
            const long d = std::strtol(c, &end, 10);
            if (end == c) break;
            devices.push_back((int32_t)d);
            c = *end == ',' ? end + 1 : end;
        }
    } else if (const char* e = std::getenv("WEPP_GPUS")) {
        for (int d = 0; d < std::atoi(e); ++d) devices.push_back(d);
    }
    if (devices.empty()) devices.push_back(std::getenv("WEPP_DEVICE") ? std::atoi(std::getenv("WEPP_DEVICE")) : 0);
    std::vector<int32_t> out((size_t)n);
    int32_t n_peaks = 0, n_out = 0;
    Timer t;
    // haplotype::full_score with score = orig_score (recover_haplotype_state, haplotype.hpp:51-55,183-185)
    auto summary = [&](wepp_handle* h) {
        std::vector<double> score((size_t)n), dd((size_t)n);
        if (wepp_get_node_summary(h, score.data(), dd.data()) != WEPP_OK) return false;
        p.full_score.resize((size_t)n);
        for (int32_t v = 0; v < n; ++v) p.full_score[(size_t)v] = score[(size_t)v] * std::sqrt(dd[(size_t)v]);
        return true;
    };
    if (devices.size() == 1) {
        if (wepp_create(devices[0], &p.h) != WEPP_OK) return std::string("no usable B200 device: ") + wepp_last_error();
        if (wepp_set_arena(p.h, n, a.parent.data(), a.mut_off.data(), a.mut_pos.data(), a.mut_ref.data(), a.mut_nuc.data(), a.genome_size) != WEPP_OK)
            return wepp_last_error();
        if (wepp_set_reads(p.h, p.reads.n_reads(), p.reads.start.data(), p.reads.end.data(), p.reads.degree.data(), a.rm_off.data(),
                           a.rm_pos.data(), a.rm_nuc.data()) != WEPP_OK)
            return wepp_last_error();
        t = Timer();
        if (wepp_filter_peaks(p.h, a.leaf_count.data(), p.id_rank.data(), out.data(), n, &n_peaks, &n_out) != WEPP_OK) return wepp_last_error();
        std::cout << "--- cartesian mapping + peak selection on the GPU took " << t.seconds() << " seconds " << std::endl;
        if (!summary(p.h)) return wepp_last_error();
    } else {
        wepp_group* g = nullptr;
        if (wepp_group_create((int32_t)devices.size(), devices.data(), &g) != WEPP_OK) return std::string("no usable group of B200 devices: ") + wepp_last_error();
        struct Guard {
            wepp_group* g;
            ~Guard() { wepp_group_destroy(g); }
        } guard{g};
        if (wepp_group_set_arena(g, n, a.parent.data(), a.mut_off.data(), a.mut_pos.data(), a.mut_ref.data(), a.mut_nuc.data(), a.genome_size) != WEPP_OK)
            return wepp_last_error();
        if (wepp_group_set_reads(g, p.reads.n_reads(), p.reads.start.data(), p.reads.end.data(), p.reads.degree.data(), a.rm_off.data(),
                                 a.rm_pos.data(), a.rm_nuc.data()) != WEPP_OK)
            return wepp_last_error();
        t = Timer();
        if (wepp_group_filter_peaks(g, a.leaf_count.data(), p.id_rank.data(), out.data(), n, &n_peaks, &n_out) != WEPP_OK) return wepp_last_error();
        std::cout << "--- cartesian mapping + peak selection on " << devices.size() << " GPUs took " << t.seconds() << " seconds " << std::endl;
        if (!summary(wepp_group_handle(g, 0))) return wepp_last_error();
        // the first rank's handle lives on for the later stages (rescoring over candidate sets, the writers), which see
        // the whole read set again; the other ranks go with the group
        p.h = wepp_group_take(g, 0);
        if (wepp_set_reads(p.h, p.reads.n_reads(), p.reads.start.data(), p.reads.end.data(), p.reads.degree.data(), a.rm_off.data(),
                           a.rm_pos.data(), a.rm_nuc.data()) != WEPP_OK)
            return wepp_last_error();
    }
    out.resize((size_t)n_out);
    running = std::move(out);
    return "";
}

// ---- post filter ------------------------------------------------------------------------------------
namespace {

struct ScoreCmp {
    const Pipeline* p;
    bool operator()(int32_t l, int32_t r) const { return p->score_less(l, r); }
};
using ScoreSet = std::set<int32_t, ScoreCmp>;

// arena::closest_neighbors (arena.cpp:171-207): breadth-first over tree edges, a node is kept (and expanded)
// while it lies within max_radius mutations of the target; the best num_limit by score_comparator survive.
// The std::set with the reference's comparator is kept on purpose: its membership test is the comparator's
// equivalence, insertion order is the reference's.
ScoreSet closest_neighbors(Pipeline& p, int32_t target, int max_radius, int num_limit) {
    ScoreSet all(ScoreCmp{&p}), ret(ScoreCmp{&p});
    std::deque<int32_t> q;
    q.push_back(target);
    while (!q.empty()) {
        const int32_t curr = q.front();
        q.pop_front();
        if (all.find(curr) != all.end() || p.stacks->distance(curr, target) > max_radius) continue;
        all.insert(curr);
        if (p.arena.parent[(size_t)curr] >= 0) q.push_back(p.arena.parent[(size_t)curr]);
        for (int64_t k = p.child_off[(size_t)curr]; k < p.child_off[(size_t)curr + 1]; ++k) q.push_back(p.child[(size_t)k]);
    }
    int taken = 0;
    for (int32_t v : all) {
        if (taken++ == num_limit) break;
        ret.insert(v);
    }
    return ret;
}

// freyja_post_filter::dump_barcode, post_filter.cpp:7-54
std::string dump_barcode(Pipeline& p, const std::vector<int32_t>& haps) {
    const std::string& ref = p.reference;
    auto mut_string = [&](uint8_t ref_nuc, int32_t pos, uint8_t nuc) {
        return std::string(1, nuc_char(ref_nuc)) + std::to_string(pos) + std::string(1, nuc_char(nuc));
    };
    std::set<std::string> mutations;
    std::vector<std::set<std::string>> node_muts;
    node_muts.reserve(haps.size());
    for (int32_t v : haps) {
        std::set<std::string> mine;
        const auto st = p.stacks->get(v);
        for (const StackMut& m : *st) {
            std::string s = mut_string(m.ref, m.pos, m.nuc);
            mutations.insert(s);
            mine.insert(std::move(s));
        }
        node_muts.push_back(std::move(mine));
    }
    // every non-N read mutation (after masking): de-duplicated as (position, allele) before the strings are made
    {
        const size_t g = ref.size();
        std::vector<uint8_t> seen((g + 2) * 16, 0);   // one slot per 4-bit code: a foreign .pb may carry IUPAC codes
        const ArenaHost& a = p.arena;
        for (size_t k = 0; k < a.rm_pos.size(); ++k) {
            const uint8_t nuc = a.rm_nuc[k];
            if (nuc == 15) continue;
            uint8_t& s = seen[(size_t)a.rm_pos[k] * 16 + (size_t)(nuc & 15)];
            if (s) continue;
            s = 1;
            mutations.insert(mut_string(nuc_id(ref[(size_t)a.rm_pos[k] - 1]), a.rm_pos[k], nuc));
        }
    }
    std::ofstream out(p.ds.barcodes_path(), std::ios::binary);
    if (!out) return "cannot write " + p.ds.barcodes_path();
    std::string buf;
    for (const std::string& m : mutations) {
        buf += ',';
        buf += m;
    }
    buf += '\n';
    out << buf;
    for (size_t i = 0; i < haps.size(); ++i) {
        buf.clear();
        buf += 'N';
        buf += std::to_string(haps[i]);
        auto it = node_muts[i].begin();
        for (const std::string& m : mutations) {   // both sets are sorted: one merge pass
            bool contains = false;
            if (it != node_muts[i].end() && *it == m) {
                contains = true;
                ++it;
            }
            buf += contains ? ",1" : ",0";
        }
        buf += '\n';
        out << buf;
    }
    return "";
}

// freyja_post_filter::filter, post_filter.cpp:56-124
std::vector<Abundance> freyja_filter(Pipeline& p, const std::vector<int32_t>& input) {
    std::fprintf(stderr, "%ld peaks selected for Freyja!\n\n", (long)input.size());
    std::string err = dump_barcode(p, input);
    if (!err.empty()) {
        std::cerr << err << std::endl;
        return {};
    }
    const Dataset& ds = p.ds;
    double af_thresh = 0.0;
    if (ds.min_af() > 0.01) af_thresh = ds.min_af();
    namespace fs = std::filesystem;
    const fs::path idir = fs::current_path() / ds.intermediate_directory();
    const std::string variants = (idir / (ds.o.file_prefix + "_corrected_variants.tsv")).string();
    const std::string depth = (idir / (ds.o.file_prefix + "_depth.tsv")).string();
    const std::string barcodes = (idir / (ds.o.file_prefix + "_barcodes.csv")).string();
    const std::string output = (idir / "freyja_output_latest.txt").string();
    const std::string command = "bash -c \"cd " + ds.o.working_directory + "/src/Freyja/ && freyja demix '" + variants + "' '" + depth +
                                "' --barcodes '" + barcodes + "' --output '" + output + "' --eps " + std::to_string(ds.min_prop()) +
                                " --af " + std::to_string(af_thresh) + "\"";
    if (std::system(command.c_str()) != 0) {
        std::cerr << "Failed to run freyja" << std::endl;
        return {};
    }
    std::vector<Abundance> nodes;
    std::ifstream fin(ds.intermediate_directory() + "freyja_output_latest.txt");
    std::string tmp;
    std::getline(fin, tmp);
    std::getline(fin, tmp);
    fin >> tmp;
    std::getline(fin, tmp);   // all of the selected ids
    const int32_t n = (int32_t)p.arena.parent.size();
    {
        std::stringstream ss{tmp};
        std::string index;
        while (ss >> index) {
            const int ind = std::atoi(index.c_str() + 1);   // skip past the 'N'
            if (ind < 0 || ind >= n) {
                std::cerr << "freyja returned an unknown haplotype " << index << std::endl;
                return {};
            }
            nodes.push_back({ind, 0.0});
        }
    }
    fin >> tmp;
    std::getline(fin, tmp);
    std::stringstream ss{tmp};
    double sum = 0;
    for (Abundance& a : nodes) {
        double ab = 0;
        ss >> ab;
        a.value = ab;
        sum += ab;
    }
    for (Abundance& a : nodes) a.value /= sum;
    return nodes;
}

}  // namespace

// post_filter::iterative_filter, post_filter.hpp:19-68 (haplotype pointers order like arena indices)
std::string pipeline_post_filter(Pipeline& p, std::vector<int32_t> input, std::vector<Abundance>& out) {
    const int num_filter_rounds = MAX_NEIGHBOR_ITERATIONS, freeze_round = 1;
    std::set<int32_t> frozen, last_round;
    // WEPP_TIMING=1: stage times on stderr (development aid)
    const bool timing = std::getenv("WEPP_TIMING") && std::atoi(std::getenv("WEPP_TIMING")) != 0;
    auto t_last = std::chrono::steady_clock::now();
    auto stage = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (timing && what) std::fprintf(stderr, "[wepp timing] %-28s %9.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    out.clear();
    for (int i = 0; i < num_filter_rounds; ++i) {
        std::vector<int32_t> full_input = input;
        for (int32_t hap : frozen)
            if (std::find(full_input.begin(), full_input.end(), hap) == full_input.end()) full_input.push_back(hap);
        stage(nullptr);
        std::vector<Abundance> filtered = freyja_filter(p, full_input);
        stage("barcodes + freyja demix");
        std::vector<int32_t> this_round;
        for (const Abundance& a : filtered) this_round.push_back(a.hap);
        std::sort(this_round.begin(), this_round.end());
        if (i == num_filter_rounds - 1 || std::includes(last_round.begin(), last_round.end(), this_round.begin(), this_round.end())) {
            out = std::move(filtered);
            return "";
        }
        if (i >= freeze_round)
            std::set_intersection(this_round.begin(), this_round.end(), last_round.begin(), last_round.end(), std::inserter(frozen, frozen.end()));
        last_round = std::set<int32_t>(this_round.begin(), this_round.end());
        // add neighbours: every haplotype's neighbourhood is searched in parallel (closest_neighbors is a pure
        // function of the tree and the scores); the union is built serially, in the reference's order
        std::vector<std::vector<int32_t>> nbrs(this_round.size());
        {
            std::vector<std::thread> pool;
            std::atomic<size_t> next{0};
            for (int t = 0; t < p.n_threads; ++t)
                pool.emplace_back([&]() {
                    for (size_t k; (k = next.fetch_add(1)) < this_round.size();) {
                        const ScoreSet s = closest_neighbors(p, this_round[k], MAX_NEIGHBOR_MUTATION, MAX_NEIGHBORS_FREYJA);
                        nbrs[k].assign(s.begin(), s.end());
                    }
                });
            for (auto& th : pool) th.join();
        }
        ScoreSet build(ScoreCmp{&p});
        for (const std::vector<int32_t>& nb : nbrs) build.insert(nb.begin(), nb.end());
        input.assign(build.begin(), build.end());
        stage("neighbour expansion");
    }
    return "";
}

// ---- detect_peaks / pipeline::run / run_from_last_initial (pipeline.cpp:5-80) ---------------------------
int detect_peaks(const Dataset& ds) {
    Pipeline p{ds};
    std::string err = pipeline_load(p);
    if (!err.empty()) {
        std::fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    std::vector<int32_t> running;
    {
        std::cout << "----- [running initial filter] -----" << std::endl;
        std::cout << "--- in: " << p.arena.parent.size() << " haplotypes" << std::endl;
        Timer t;
        err = pipeline_initial_filter(p, running);
        if (!err.empty()) {
            std::fprintf(stderr, "%s\n", err.c_str());
            return 1;
        }
        std::cout << "--- initial filter took " << t.seconds() << " seconds " << std::endl << std::endl;
    }
    {   // pipeline::save / recover, pipeline.hpp:25-42
        std::ofstream fout(ds.checkpoint_path());
        for (int32_t v : running) fout << v << std::endl;
    }
    {
        std::ifstream fin(ds.checkpoint_path());
        running.clear();
        int index;
        while (fin >> index) running.push_back(index);
    }
    std::cout << "----- [running post filter] -----" << std::endl;
    std::cout << "--- in: " << running.size() << " haplotypes" << std::endl;
    Timer t;
    std::vector<Abundance> full;
    err = pipeline_post_filter(p, running, full);
    if (!err.empty()) {
        std::fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    const bool timing = std::getenv("WEPP_TIMING") && std::atoi(std::getenv("WEPP_TIMING")) != 0;
    const auto t_w = std::chrono::steady_clock::now();
    err = pipeline_write_results(p, full);
    if (!err.empty()) {
        std::fprintf(stderr, "%s\n", err.c_str());
        return 1;
    }
    if (timing)
        std::fprintf(stderr, "[wepp timing] %-28s %9.1f ms\n", "result files",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_w).count());
    std::cout << "--- post filter + result files took " << t.seconds() << " seconds " << std::endl;
    std::cout << "--- RUN COMPLETED" << std::endl;
    return 0;
}

// ---- main.cpp:14-70 ------------------------------------------------------------------------------------
int cli_main(int argc, const char* const* argv) {
    static const char* cnames[] = {"COMMAND", "detectPeaks", "sam2PB"};
    static const char* chelp[] = {"DESCRIPTION\n\n", "Detects Peaks from the MAT\n\n", "Applies QC before running WEPP\n\n"};
    auto print_help = [&]() {
        for (int i = 0; i < 3; ++i) std::fprintf(stderr, "%-15s\t%s", cnames[i], chelp[i]);
        std::cerr << "\n" << usage_text() << "\n";
    };
    // the command is the first positional token (boost positional "command", main.cpp:21-22)
    std::vector<std::string> rest;
    std::string cmd;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (cmd.empty() && !a.empty() && a[0] != '-') {   // the first positional token, wherever it stands
            cmd = a;
            continue;
        }
        rest.push_back(a);
    }
    if (cmd.empty()) {
        std::fprintf(stderr, "\nNo command selected. Help follows:\n\n");
        print_help();
        return 0;   // 0 when no command is selected (main.cpp:44-45; CI depends on it)
    }
    if (cmd == "detectPeaks" || cmd == "sam2PB") {
        Options o;
        const std::string err = parse_options(rest, o);
        if (o.help) {
            std::cout << usage_text() << std::endl;
            return 0;
        }
        if (!err.empty()) {
            std::cerr << err << "\n" << usage_text() << std::endl;
            return 1;
        }
        Dataset ds{o};
        return cmd == "detectPeaks" ? detect_peaks(ds) : sam2pb(ds);
    }
    if (cmd == "help") {
        std::fprintf(stderr, "\n");
        print_help();
        return 0;
    }
    std::fprintf(stderr, "\nInvalid command. Help follows:\n\n");
    print_help();
    return 1;
}

}  // namespace wepp

extern "C" int wepp_cli_main(int argc, const char* const* argv) { return wepp::cli_main(argc, argv); }
