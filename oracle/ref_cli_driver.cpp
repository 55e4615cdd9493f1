// ref_cli_driver.cpp — the REFERENCE'S WHOLE `wepp` BINARY under the stand-in headers (TEST INFRASTRUCTURE).
//
// oracle/Makefile target `refcli` compiles the reference's own translation units from where they lie —
// src/WEPP/{arena,util,dataset,initial_filter,pipeline,post_filter,sam2pb}.cpp and
// src/mutation_annotated_tree.cpp — against oracle/shim (oneTBB, Boost, protobuf are absent from this
// image) and links them with this file into oracle/_ref/wepp_ref.  Only the reference's main.cpp is
// replaced: its boost::program_options command line (src/WEPP/main.cpp:14-71, util.cpp:137-186) is parsed
// here by hand into the same variables_map (same flags, same defaults), after which the reference's
// detect_peaks(ds) / sam2PB(ds) run unmodified: MAT + read loaders, arena, placement, peak loop,
// post filter (it shells out to whatever `freyja` is on PATH), all result writers.
// `wepp_ref loadmat <file> <uncondense 0|1>` additionally dumps the loaded tree (ids, parents, mutations,
// annotations) as text for the loader parity test.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "tbb/tbb.h"
#include <boost/program_options.hpp>

#include "src/usher_graph.hpp"
#include "src/WEPP/pipeline.hpp"
#include "src/WEPP/sam2pb.hpp"
#include "src/WEPP/util.hpp"

Timer timer;   // src/WEPP/main.cpp:12

static int dump_mat(const char* path, bool uncondense) {
    MAT::Tree T = MAT::load_mutation_annotated_tree(path);
    if (uncondense) T.uncondense_leaves();
    for (MAT::Node* n : T.depth_first_expansion()) {
        std::printf("%s\t%s\t", n->identifier.c_str(), n->parent ? n->parent->identifier.c_str() : "");
        for (auto& m : n->mutations) std::printf("%d:%d:%d:%d,", m.position, (int)m.ref_nuc, (int)m.par_nuc, (int)m.mut_nuc);
        std::printf("\t");
        for (auto& c : n->clade_annotations) std::printf("%s|", c.c_str());
        std::printf("\n");
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: wepp_ref detectPeaks|sam2PB|loadmat ...\n");
        return 0;
    }
    const std::string cmd = argv[1];
    if (cmd == "loadmat") return dump_mat(argv[2], argc > 3 && std::atoi(argv[3]) != 0);
    boost::program_options::variables_map vm;
    vm.set<std::string>("working-directory", "./");
    vm.set<std::string>("input-mat", "");
    vm.set<std::string>("dataset", "");
    vm.set<uint32_t>("max-reads", 1000000000u);
    vm.set<std::string>("file-prefix", "");
    vm.set<std::string>("ref-fasta", "");
    vm.set<std::string>("min-af", "0.005");
    vm.set<uint32_t>("min-depth", 10u);
    vm.set<uint32_t>("min-phred", 20u);
    vm.set<std::string>("min-prop", "0.005");
    vm.set<uint32_t>("clade-idx", 1u);
    vm.set<uint32_t>("threads", 4u);
    for (int i = 2; i + 1 < argc; i += 2) {
        const std::string f = argv[i], v = argv[i + 1];
        if (f == "-w") vm.set<std::string>("working-directory", v);
        else if (f == "-i") vm.set<std::string>("input-mat", v);
        else if (f == "-d") vm.set<std::string>("dataset", v);
        else if (f == "-m") vm.set<uint32_t>("max-reads", (uint32_t)std::stoul(v));
        else if (f == "-p") vm.set<std::string>("file-prefix", v);
        else if (f == "-f") vm.set<std::string>("ref-fasta", v);
        else if (f == "-a") vm.set<std::string>("min-af", v);
        else if (f == "-c") vm.set<uint32_t>("min-depth", (uint32_t)std::stoul(v));
        else if (f == "-q") vm.set<uint32_t>("min-phred", (uint32_t)std::stoul(v));
        else if (f == "-r") vm.set<std::string>("min-prop", v);
        else if (f == "-n") vm.set<uint32_t>("clade-idx", (uint32_t)std::stol(v));
        else if (f == "-T") vm.set<uint32_t>("threads", (uint32_t)std::stoul(v));
        else { std::fprintf(stderr, "unknown flag %s\n", f.c_str()); return 1; }
    }
    dataset ds{vm};
    if (cmd == "detectPeaks") detect_peaks(ds);
    else if (cmd == "sam2PB") sam2PB(ds);
    else return 1;
    return 0;
}
