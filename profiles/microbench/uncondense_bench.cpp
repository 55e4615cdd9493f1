#include "../../wepp_b200/csrc/host_io.h"
#include <chrono>
#include <cstdio>
#include <random>
using namespace wepp;
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4000000;
    MatTree t;
    std::mt19937 rng(3);
    t.parent.resize(n); t.id.resize(n); t.branch_length.assign(n, 1.0f); t.mut_off.assign(n + 1, 0); t.clade.assign(n, {});
    std::vector<char> has_child(n, 0);
    for (int v = 0; v < n; ++v) { t.parent[v] = v ? (int)(rng() % v) : -1; if (v) has_child[t.parent[v]] = 1; }
    for (int v = 0; v < n; ++v) t.id[v] = has_child[v] ? "node_" + std::to_string(v) : "S" + std::to_string(v) + "|hap/" + std::to_string(v) + "|2021-01-01";
    // a mutation on every third node
    for (int v = 0; v < n; ++v) { if (v % 3 == 0) { t.mut_pos.push_back(1 + v % 29000); t.mut_ref.push_back(1); t.mut_par.push_back(1); t.mut_nuc.push_back(2); } t.mut_off[v + 1] = (int64_t)t.mut_pos.size(); }
    int nc = 0;
    for (int v = 0; v < n; ++v) if (!has_child[v] && rng() % 5 == 0) {
        t.condensed_name.push_back(t.id[v]);
        std::vector<std::string> l;
        const int k = 1 + rng() % 4;
        for (int j = 0; j < k; ++j) l.push_back(t.id[v] + "_c" + std::to_string(j));
        t.condensed_leaves.push_back(std::move(l));
        ++nc;
    }
    t.n_internal_ids = n;
    printf("%d nodes, %d condensed nodes\n", n, nc);
    auto t0 = std::chrono::steady_clock::now();
    uncondense_leaves(t);
    printf("uncondense_leaves %.2f s -> %zu nodes\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), t.parent.size());
    unsigned long long h = 0;
    for (size_t v = 0; v < t.id.size(); ++v) h = h * 1000003ull + std::hash<std::string>()(t.id[v]) + (unsigned)t.parent[v];
    printf("checksum %llx\n", h);
}
