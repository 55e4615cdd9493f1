// pipeline.h — the `wepp` command line around the placement path: what Snakemake calls
// (SURVEY §8b "Process/CLI boundary"; reference src/WEPP/main.cpp:14-70, util.cpp:137-186,
// dataset.hpp:10-229, pipeline.cpp:5-80).
//
//   wepp detectPeaks ...   load MAT / reads / FASTA / mask -> arena -> initial filter on the GPU
//                          (wepp_filter_peaks) -> <P>_checkpoint.txt -> iterative Freyja post filter
//                          (post_filter.hpp:19-68, post_filter.cpp:7-124) -> the result files
//                          (arena.cpp:446-931), whose per-read work is the K4 kernel (wepp_rescore_reads)
//   wepp sam2PB ...        SAM -> collapsed reads -> <P>_reads.pb (sam2pb.cpp:54-477)
//   wepp help
// The host logic is C++ like the reference's; everything per-read x per-haplotype runs through the C ABI
// in include/wepp_b200.h.  There is no CPU fallback: detectPeaks fails when wepp_create finds no sm_100 GPU.
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "host_arena.h"
#include "host_io.h"

struct wepp_handle;

namespace wepp {

// command-line options, util.cpp:145-158 (same names, short flags and defaults)
struct Options {
    std::string working_directory = "./";
    std::string input_mat, dataset, file_prefix, ref_fasta;
    std::string min_af = "0.005", min_prop = "0.005";
    uint32_t max_reads = 1000000000u, min_depth = 10, min_phred = 20, clade_idx = 1, threads = 0;
    bool help = false;
};
// returns "" on success; on failure the message (usage is printed by the caller)
std::string parse_options(const std::vector<std::string>& args, Options& out);
std::string usage_text();

// dataset.hpp:22-211: every path the stages use, relative to the current directory
struct Dataset {
    Options o;
    explicit Dataset(Options opt) : o(std::move(opt)) {}
    double min_af() const { return (double)std::stof(o.min_af); }       // dataset.hpp:38-40
    double min_prop() const { return (double)std::stof(o.min_prop); }   // :46-48
    int32_t clade_idx() const { return (int32_t)o.clade_idx; }          // arena.hpp:119-121 (uint32 -> int32)
    std::string data_directory() const { return "./data/" + o.dataset + "/"; }
    std::string intermediate_directory() const { return "./intermediate/" + o.dataset + "/"; }
    std::string results_directory() const { return "./results/" + o.dataset + "/"; }
    std::string ref_path() const { return data_directory() + o.ref_fasta; }
    std::string mat_path() const { return data_directory() + o.input_mat; }
    std::string mask_path() const { return data_directory() + "/mask.bed"; }
    std::string pb_path() const { return intermediate_directory() + o.file_prefix + "_reads.pb"; }
    std::string sam_path() const { return intermediate_directory() + o.file_prefix + "_alignment.sam"; }
    std::string checkpoint_path() const { return intermediate_directory() + o.file_prefix + "_checkpoint.txt"; }
    std::string barcodes_path() const { return intermediate_directory() + o.file_prefix + "_barcodes.csv"; }
    std::string residual_mutations_path() const { return intermediate_directory() + "residual_mutations.txt"; }
    std::string result(const char* suffix) const { return results_directory() + o.file_prefix + suffix; }
};

// one stack_muts entry (arena.cpp:18-46): position, the MAT mutation's ref_nuc, the net allele
struct StackMut {
    int32_t pos;
    uint8_t ref, nuc;
};
using Stack = std::vector<StackMut>;

// haplotype::stack_muts on demand: the reference materialises them for all N nodes (O(N * depth) memory);
// the post filter and the writers touch a few thousand haplotypes, so they are rebuilt from the root path
// when first asked for and cached (thread-safe).
class HapStacks {
public:
    explicit HapStacks(const ArenaHost& a) : a_(a) {}
    std::shared_ptr<const Stack> get(int32_t v);
    // a->mutation_distance(b): haplotype.hpp:123-173 with comp = b->stack_muts over [0, INT_MAX] (:179-181)
    int distance(int32_t a, int32_t b);
private:
    static constexpr int SHARDS = 64;
    const ArenaHost& a_;
    std::mutex mu_[SHARDS];
    std::unordered_map<int32_t, std::shared_ptr<const Stack>> cache_[SHARDS];
};

struct Abundance {
    int32_t hap;      // arena index
    double value;
};

// everything detectPeaks holds (the reference's `pipeline` + `arena` objects, pipeline.hpp:13-24)
struct Pipeline {
    Dataset ds;
    int n_threads = 1;
    std::string ref_name, reference;
    MatTree mat;
    std::vector<int32_t> masked;
    ReadSet reads;           // as loaded; the masked mutation lists live in arena.rm_*
    ArenaHost arena;
    std::vector<int32_t> id_rank;      // rank of haplotype::id in std::string order
    std::vector<int64_t> child_off;    // arena children CSR (preorder = creation order)
    std::vector<int32_t> child;
    wepp_handle* h = nullptr;
    std::vector<double> full_score;    // haplotype::full_score() after recover_haplotype_state
    std::unique_ptr<HapStacks> stacks;

    explicit Pipeline(Dataset d) : ds(std::move(d)) {}
    ~Pipeline();
    const std::string& hap_id(int32_t v) const { return mat.id[(size_t)arena.source[(size_t)v]]; }
    // score_comparator, arena.hpp:16-31
    bool score_less(int32_t l, int32_t r) const;
};

// stages; each returns "" or an error message
std::string pipeline_load(Pipeline& p);                                        // arena::arena, arena.hpp:56-79
std::string pipeline_initial_filter(Pipeline& p, std::vector<int32_t>& running);   // pipeline.cpp:24-41
std::string pipeline_post_filter(Pipeline& p, std::vector<int32_t> input, std::vector<Abundance>& out);   // post_filter.hpp:19-68
std::string pipeline_write_results(Pipeline& p, const std::vector<Abundance>& full);   // pipeline.cpp:53,70-75
int detect_peaks(const Dataset& ds);   // pipeline.cpp:5-22; returns the process exit code
int sam2pb(const Dataset& ds);         // sam2pb.cpp:54-109
int cli_main(int argc, const char* const* argv);   // main.cpp:14-70

}  // namespace wepp
