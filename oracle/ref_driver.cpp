// ref_driver.cpp — drives the REFERENCE'S OWN placement code, shim-compiled from where it lies.
//
// TEST INFRASTRUCTURE ONLY (see oracle/Makefile).  This translation unit textually includes
// /root/reference/src/WEPP/initial_filter.cpp so that its file-static `single_read_tree`
// (initial_filter.cpp:112-135) and wepp_filter's private state become reachable; arena.cpp,
// util.cpp and dataset.cpp are compiled unmodified as separate objects.  oneTBB, Boost and the
// protoc-generated header are replaced by the small stand-ins under oracle/shim/ (they are absent
// from this image); the protobuf / gzip loaders (mutation_annotated_tree.cpp, sam2pb.cpp) are NOT
// compiled — this file provides the handful of MAT:: helpers the placement path links against
// and injects the tree and the reads directly.  Everything between "inject" and "fetch results"
// — masking, site_read_map, create_condensed_tree, arena::from_mat, build_range_trees,
// cartesian_map, single_read_tree, mutation_distance, wepp_filter::filter — is the reference's
// object code.
//
// No reference source is copied into the repository; the build reads it from /root/reference.

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "tbb/tbb.h"
#include <boost/program_options.hpp>

// compiled with -fno-access-control (oracle/Makefile): wepp_filter keeps its per-read state private
#include "src/WEPP/initial_filter.cpp"

// ------------------------------------------------------------------------------------------------
// Minimal MAT:: helpers (own code; behaviour per the cited reference lines).
namespace Mutation_Annotated_Tree {

// src/mutation_annotated_tree.cpp:19-74 (note: 'V' falls through to N there; kept)
int8_t get_nuc_id(char nuc) {
    switch (nuc) {
        case 'a': case 'A': return 1;
        case 'c': case 'C': return 2;
        case 'g': case 'G': return 4;
        case 't': case 'T': return 8;
        case 'R': return 5;
        case 'Y': return 10;
        case 'S': return 6;
        case 'W': return 9;
        case 'K': return 12;
        case 'M': return 3;
        case 'B': return 14;
        case 'D': return 13;
        case 'H': return 11;
        default: return 15;
    }
}
// src/mutation_annotated_tree.cpp:88-139
char get_nuc(int8_t id) {
    static const char* t = "NACMGRSVTWYHKDBN";
    return (id >= 1 && id <= 14) ? t[id] : 'N';
}
void string_split(std::string const& s, char delim, std::vector<std::string>& words) {
    size_t a = 0, b;
    while ((b = s.find(delim, a)) != std::string::npos) {
        words.emplace_back(s.substr(a, b - a));
        a = b + 1;
    }
    words.emplace_back(s.substr(a));
}
void string_split(std::string s, std::vector<std::string>& words) {
    std::istringstream ss(s);
    std::string w;
    while (ss >> w) words.push_back(std::move(w));
}
Node::Node() : level(0), branch_length(-1), parent(nullptr) {}
Node::Node(std::string id, float l) : level(1), branch_length(l), identifier(id), parent(nullptr) {}
Node::Node(std::string id, Node* p, float l) : level(p->level + 1), branch_length(l), identifier(id), parent(p) {}
bool Node::is_leaf() { return children.empty(); }
bool Node::is_root() { return parent == nullptr; }
size_t Tree::get_num_annotations() const { return root ? root->clade_annotations.size() : 0; }
// src/mutation_annotated_tree.cpp:854-878
Node* Tree::create_node(std::string const& identifier, float branch_len, size_t num_annotations) {
    all_nodes.clear();
    Node* n = new Node(identifier, branch_len);
    for (size_t k = 0; k < num_annotations; k++) n->clade_annotations.emplace_back("");
    root = n;
    all_nodes[identifier] = root;
    return n;
}
Node* Tree::create_node(std::string const& identifier, Node* par, float branch_len) {
    if (all_nodes.find(identifier) != all_nodes.end()) {
        fprintf(stderr, "Error: %s already in the tree!\n", identifier.c_str());
        exit(1);
    }
    Node* n = new Node(identifier, par, branch_len);
    size_t na = get_num_annotations();
    for (size_t k = 0; k < na; k++) n->clade_annotations.emplace_back("");
    all_nodes[identifier] = n;
    par->children.push_back(n);
    return n;
}
Node* Tree::get_node(std::string nid) const {
    auto it = all_nodes.find(nid);
    return it == all_nodes.end() ? nullptr : it->second;
}
// src/mutation_annotated_tree.cpp:904-921
std::vector<Node*> Tree::rsearch(const std::string& nid, bool include_self) const {
    std::vector<Node*> anc;
    Node* node = get_node(nid);
    if (!node) return anc;
    if (include_self) anc.push_back(node);
    while (node->parent) {
        anc.push_back(node->parent);
        node = node->parent;
    }
    return anc;
}
void Tree::uncondense_leaves() {}  // injected trees carry no condensed nodes

}  // namespace Mutation_Annotated_Tree

// ------------------------------------------------------------------------------------------------
// Injection points the reference calls through dataset::mat() / dataset::reads().
namespace {
struct Injected {
    int32_t n_nodes = 0;
    const int32_t* parent = nullptr;
    const int64_t* mut_off = nullptr;
    const int32_t* mut_pos = nullptr;
    const uint8_t* mut_ref = nullptr;
    const uint8_t* mut_par = nullptr;
    const uint8_t* mut_nuc = nullptr;
    int64_t n_reads = 0;
    const int32_t* start = nullptr;
    const int32_t* end = nullptr;
    const int32_t* degree = nullptr;
    const int64_t* rm_off = nullptr;
    const int32_t* rm_pos = nullptr;
    const uint8_t* rm_nuc = nullptr;
} g_in;
}  // namespace

MAT::Tree Mutation_Annotated_Tree::load_mutation_annotated_tree(std::string) {
    MAT::Tree T;
    std::vector<MAT::Node*> nodes((size_t)g_in.n_nodes);
    for (int32_t v = 0; v < g_in.n_nodes; ++v) {
        std::string id = "n" + std::to_string(v);
        nodes[v] = v == 0 ? T.create_node(id, -1.0f, 0) : T.create_node(id, nodes[g_in.parent[v]], -1.0f);
        for (int64_t k = g_in.mut_off[v]; k < g_in.mut_off[v + 1]; ++k) {
            MAT::Mutation m;
            m.position = g_in.mut_pos[k];
            m.ref_nuc = (int8_t)g_in.mut_ref[k];
            m.par_nuc = (int8_t)g_in.mut_par[k];
            m.mut_nuc = (int8_t)g_in.mut_nuc[k];
            nodes[v]->mutations.push_back(m);  // already sorted and unique per node
        }
    }
    return T;
}

// What src/WEPP/sam2pb.cpp:508-536 produces per read, built from the injected sparse form.
std::vector<raw_read> load_reads_from_proto(std::string const& reference, std::string const&,
                                            std::unordered_map<std::string, std::vector<std::string>>& reverse_merge) {
    std::vector<raw_read> reads((size_t)g_in.n_reads);
    for (int64_t r = 0; r < g_in.n_reads; ++r) {
        raw_read& out = reads[r];
        out.start = g_in.start[r];
        out.end = g_in.end[r];
        out.degree = g_in.degree[r];
        out.read = "r" + std::to_string(r);
        for (int64_t k = g_in.rm_off[r]; k < g_in.rm_off[r + 1]; ++k) {
            MAT::Mutation m;
            m.is_missing = g_in.rm_nuc[k] == 15;
            m.chrom = "NC_045512v2";
            m.position = g_in.rm_pos[k];
            m.ref_nuc = m.par_nuc = MAT::get_nuc_id(reference[m.position - 1]);
            m.mut_nuc = (int8_t)g_in.rm_nuc[k];
            out.mutations.push_back(m);
        }
        reverse_merge[out.read].push_back(out.read);
    }
    return reads;
}

// dataset::read_reverse_merge lives in dataset.cpp (compiled); nothing else needed here.

// ------------------------------------------------------------------------------------------------
namespace {
struct Session {
    dataset* ds = nullptr;
    arena* ar = nullptr;
    wepp_filter* filt = nullptr;
};
Session g_s;  // the reference caches per-process statics (arena.hpp:158, dataset.hpp:91,180): one session per process

int hap_index(haplotype* h) { return (int)(h - &g_s.ar->haplotypes()[0]); }
}  // namespace

extern "C" {

// Creates <workdir>/data/ds/{ref.fa,mask.bed}, chdirs into workdir, injects tree and reads and runs the
// reference's arena constructor (arena.hpp:56-79).  ref_seq has genome_size characters.  Returns the
// number of arena (condensed) nodes, or -1.
int ref_open(const char* workdir, int32_t n_threads, const char* ref_seq, int32_t n_masked, const int32_t* masked,
             int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
             const uint8_t* mut_ref, const uint8_t* mut_par, const uint8_t* mut_nuc, int64_t n_reads,
             const int32_t* start, const int32_t* end, const int32_t* degree, const int64_t* rm_off,
             const int32_t* rm_pos, const uint8_t* rm_nuc) {
    if (g_s.ar) return -1;
    std::string wd(workdir);
    mkdir(wd.c_str(), 0755);
    mkdir((wd + "/data").c_str(), 0755);
    mkdir((wd + "/data/ds").c_str(), 0755);
    mkdir((wd + "/intermediate").c_str(), 0755);
    mkdir((wd + "/intermediate/ds").c_str(), 0755);
    mkdir((wd + "/results").c_str(), 0755);
    mkdir((wd + "/results/ds").c_str(), 0755);
    {
        std::ofstream fa(wd + "/data/ds/ref.fa");
        fa << ">NC_045512v2 synthetic\n" << ref_seq << "\n";
    }
    if (n_masked > 0) {
        std::ofstream mb(wd + "/data/ds/mask.bed");
        for (int32_t i = 0; i < n_masked; ++i) mb << "NC_045512v2\t" << (masked[i] - 1) << "\t" << masked[i] << "\n";
    } else {
        unlink((wd + "/data/ds/mask.bed").c_str());
    }
    if (chdir(wd.c_str()) != 0) return -1;
    g_in.n_nodes = n_nodes; g_in.parent = parent; g_in.mut_off = mut_off; g_in.mut_pos = mut_pos;
    g_in.mut_ref = mut_ref; g_in.mut_par = mut_par; g_in.mut_nuc = mut_nuc;
    g_in.n_reads = n_reads; g_in.start = start; g_in.end = end; g_in.degree = degree;
    g_in.rm_off = rm_off; g_in.rm_pos = rm_pos; g_in.rm_nuc = rm_nuc;

    boost::program_options::variables_map vm;
    vm.set<std::string>("working-directory", "./");
    vm.set<std::string>("input-mat", "tree.pb");
    vm.set<std::string>("dataset", "ds");
    vm.set<uint32_t>("max-reads", 1000000000u);
    vm.set<std::string>("file-prefix", "x");
    vm.set<std::string>("ref-fasta", "ref.fa");
    vm.set<std::string>("min-af", "0.005");
    vm.set<uint32_t>("min-depth", 10u);
    vm.set<uint32_t>("min-phred", 20u);
    vm.set<std::string>("min-prop", "0.005");
    vm.set<uint32_t>("clade-idx", 1u);
    vm.set<uint32_t>("threads", (uint32_t)std::max(1, n_threads));
    g_s.ds = new dataset(vm);
    g_s.ar = new arena(*g_s.ds);
    g_s.filt = new wepp_filter();
    return (int)g_s.ar->haplotypes().size();
}

int64_t ref_arena_mut_count(void) {
    int64_t t = 0;
    for (auto& h : g_s.ar->haplotypes()) t += (int64_t)h.muts.size();
    return t;
}
int64_t ref_arena_stack_count(void) {
    int64_t t = 0;
    for (auto& h : g_s.ar->haplotypes()) t += (int64_t)h.stack_muts.size();
    return t;
}

// The reference's flattened arena (arena.cpp:3-56): per node parent index, source MAT node index,
// leaf_count, muts CSR and stack_muts CSR.
void ref_arena_get(int32_t* parent, int32_t* source, int32_t* leaf_count, int64_t* mut_off, int32_t* mut_pos,
                   uint8_t* mut_ref, uint8_t* mut_nuc, int64_t* st_off, int32_t* st_pos, uint8_t* st_nuc) {
    auto& nodes = g_s.ar->haplotypes();
    int64_t mo = 0, so = 0;
    for (size_t i = 0; i < nodes.size(); ++i) {
        haplotype& h = nodes[i];
        parent[i] = h.parent ? hap_index(h.parent) : -1;
        source[i] = std::atoi(h.id.c_str() + 1);
        leaf_count[i] = h.leaf_count;
        mut_off[i] = mo;
        for (auto& m : h.muts) { mut_pos[mo] = m.position; mut_ref[mo] = (uint8_t)m.ref_nuc; mut_nuc[mo] = (uint8_t)m.mut_nuc; ++mo; }
        st_off[i] = so;
        for (auto& m : h.stack_muts) { st_pos[so] = m.position; st_nuc[so] = (uint8_t)m.mut_nuc; ++so; }
    }
    mut_off[nodes.size()] = mo;
    st_off[nodes.size()] = so;
}

// Reads after the reference's masking (arena.hpp:62-72): mutation counts per read and the lists.
int64_t ref_reads_mut_count(void) {
    int64_t t = 0;
    for (auto& r : g_s.ar->reads()) t += (int64_t)r.mutations.size();
    return t;
}
void ref_reads_get(int64_t* rm_off, int32_t* rm_pos, uint8_t* rm_nuc) {
    int64_t o = 0;
    const auto& reads = g_s.ar->reads();
    for (size_t r = 0; r < reads.size(); ++r) {
        rm_off[r] = o;
        for (auto& m : reads[r].mutations) { rm_pos[o] = m.position; rm_nuc[o] = (uint8_t)m.mut_nuc; ++o; }
    }
    rm_off[reads.size()] = o;
}

// wepp_filter::cartesian_map (initial_filter.cpp:139-239) on the first n_sel reads (all if < 0).
// Any output may be NULL.  epp lists: CSR over reads, arena indices sorted (the cache, :189-196).
// Returns milliseconds spent inside cartesian_map.
double ref_cartesian_map(int64_t n_sel, const uint8_t* mapped, int32_t* max_parsimony, int32_t* multiplicity,
                         double* score, int32_t* counts, double* dist_divergence, int64_t* epp_off,
                         int32_t* epp_nodes, int64_t epp_capacity) {
    arena& ar = *g_s.ar;
    wepp_filter& f = *g_s.filt;
    ar.reset_haplotype_state();
    ar.build_range_trees();
    if (mapped)
        for (size_t i = 0; i < ar.haplotypes().size(); ++i) ar.haplotypes()[i].mapped = mapped[i] != 0;
    const std::vector<raw_read>& all = ar.reads();
    std::vector<raw_read> subset;
    const std::vector<raw_read>* reads = &all;
    if (n_sel >= 0 && (size_t)n_sel < all.size()) {
        subset.assign(all.begin(), all.begin() + n_sel);
        reads = &subset;
    }
    f.reset(reads->size());
    std::vector<haplotype*> haps = ar.haplotype_pointers();
    Timer t;
    t.Start();
    auto t0 = std::chrono::steady_clock::now();
    f.cartesian_map(ar, haps, *reads);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const size_t r = reads->size(), n = ar.haplotypes().size();
    if (max_parsimony) std::copy(f.max_parismony.begin(), f.max_parismony.end(), max_parsimony);
    if (multiplicity) std::copy(f.parsimony_multiplicity.begin(), f.parsimony_multiplicity.end(), multiplicity);
    for (size_t i = 0; i < n; ++i) {
        haplotype& h = ar.haplotypes()[i];
        if (score) score[i] = h.score;
        if (dist_divergence) dist_divergence[i] = h.dist_divergence;
        if (counts)
            for (int b = 0; b < NUM_RANGE_BINS; ++b) counts[i * NUM_RANGE_BINS + b] = h.mapped_read_counts[b];
    }
    if (epp_off) {
        int64_t o = 0;
        for (size_t i = 0; i < r; ++i) {
            epp_off[i] = o;
            for (haplotype* h : f.epp_positions_cache[i]) {
                if (o < epp_capacity && epp_nodes) epp_nodes[o] = hap_index(h);
                ++o;
            }
        }
        epp_off[r] = o;
    }
    return ms;
}

// static single_read_tree (initial_filter.cpp:112-135) for one read under a mapped mask.
int ref_single_read_tree(int64_t read_idx, const uint8_t* mapped, int32_t* max_val_out, int32_t* nodes_out,
                         int32_t capacity) {
    arena& ar = *g_s.ar;
    ar.build_range_trees();
    for (size_t i = 0; i < ar.haplotypes().size(); ++i) ar.haplotypes()[i].mapped = mapped ? mapped[i] != 0 : false;
    std::vector<haplotype*> idx;
    int max_val = INT32_MAX;
    single_read_tree(ar, ar.reads()[read_idx], idx, max_val);
    std::sort(idx.begin(), idx.end());
    *max_val_out = max_val;
    int n = 0;
    for (haplotype* h : idx) {
        if (n < capacity) nodes_out[n] = hap_index(h);
        ++n;
    }
    return n;
}

// haplotype::mutation_distance(const raw_read&) (haplotype.hpp:175-177): dense R x C matrix.
void ref_mutation_distance(int32_t n_cand, const int32_t* cand, int64_t n_reads, int32_t* dist) {
    arena& ar = *g_s.ar;
    for (int64_t r = 0; r < n_reads; ++r)
        for (int32_t c = 0; c < n_cand; ++c)
            dist[r * n_cand + c] = ar.haplotypes()[cand[c]].mutation_distance(ar.reads()[r]);
}

void ref_flush(void) { std::cout.flush(); fflush(stdout); }

// wepp_filter::filter (initial_filter.cpp:455-506): cartesian map + peak loop + neighbour expansion.
int ref_filter(int32_t* selected, int32_t capacity) {
    std::vector<haplotype*> res = g_s.filt->filter(*g_s.ar);
    int n = 0;
    for (haplotype* h : res) {
        if (n < capacity) selected[n] = hap_index(h);
        ++n;
    }
    return n;
}

}  // extern "C"
