// host_io.cpp — MAT / collapsed-read / FASTA / mask.bed loaders (see host_io.h for the reference lines).
#include "host_io.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <cctype>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <queue>
#include <sstream>
#include <thread>

#include <chrono>
#include <sys/stat.h>
#include <unistd.h>

#include "pbwire.h"

namespace wepp {

// ---- host threads ---------------------------------------------------------------------------------
// WEPP_THREADS bounds the loaders' worker threads (default: the hardware concurrency, at most 64).
int io_threads() {
    int t = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("WEPP_THREADS")) t = atoi(e);
    return std::max(1, std::min(t, 64));
}
// f(thread, lo, hi) over [0, n) cut into equal ranges, one per thread
template <typename F>
static void parallel_ranges(size_t n, int n_threads, F f) {
    const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_threads, (n + 4095) / 4096));
    if (T == 1) {
        f(0, (size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t]() { f(t, n * (size_t)t / T, n * (size_t)(t + 1) / T); });
    for (auto& x : th) x.join();
}
namespace {
struct IoLaps {   // WEPP_TIMING=1: wall time of the loader's phases on stderr
    const char* what;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit IoLaps(const char* w) : what(w), on(getenv("WEPP_TIMING") && atoi(getenv("WEPP_TIMING")) != 0), t0(std::chrono::steady_clock::now()) {}
    void operator()(const char* phase) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[wepp timing] %s: %-28s %8.1f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

uint8_t nuc_id(char c) {   // src/mutation_annotated_tree.cpp:19-74 ('V' falls through to N there)
    switch (c) {
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

char nuc_char(uint8_t id) {   // src/mutation_annotated_tree.cpp:88-139
    static const char t[] = "NACMGRSVTWYHKDBN";
    return id < 16 ? t[id] : 'N';
}

// ---- files --------------------------------------------------------------------------------------
static std::string slurp(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f) return "could not open " + path;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    f.seekg(0, std::ios::beg);
    out.resize((size_t)std::max<std::streamoff>(n, 0));
    if (n > 0) f.read(&out[0], n);
    return f ? "" : "could not read " + path;
}

// gzip is detected by the file NAME containing ".gz" (src/mutation_annotated_tree.cpp:530)
static std::string inflate_if_gz(const std::string& path, std::string& raw, std::string& out);
std::string read_file_maybe_gz(const std::string& path, std::string& out) {
    std::string raw;
    std::string err = slurp(path, raw);
    if (!err.empty()) return err;
    return inflate_if_gz(path, raw, out);
}
static std::string inflate_if_gz(const std::string& path, std::string& raw, std::string& out) {
    if (path.find(".gz") == std::string::npos) {
        out.swap(raw);
        return "";
    }
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 15 + 16) != Z_OK) return "zlib: inflateInit2 failed";
    out.clear();
    out.reserve(raw.size() * 4);
    zs.next_in = (Bytef*)raw.data();
    size_t left = raw.size();
    std::vector<char> buf(1 << 20);
    int rc = Z_OK;
    for (;;) {
        if (zs.avail_in == 0 && left > 0) {
            const size_t take = std::min<size_t>(left, 1u << 30);
            zs.avail_in = (uInt)take;
            left -= take;
        }
        zs.next_out = (Bytef*)buf.data();
        zs.avail_out = (uInt)buf.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) break;
        out.append(buf.data(), buf.size() - zs.avail_out);
        if (rc == Z_STREAM_END) {
            if (zs.avail_in == 0 && left == 0) break;
            if (inflateReset(&zs) != Z_OK) break;   // concatenated gzip members
            rc = Z_OK;
        } else if (zs.avail_in == 0 && left == 0 && zs.avail_out != 0) {
            break;
        }
    }
    inflateEnd(&zs);
    if (rc != Z_STREAM_END && rc != Z_OK) return "gzip: corrupt stream in " + path;
    return "";
}

// ---- Newick -------------------------------------------------------------------------------------
// The same token machine as create_tree_from_newick_string (src/mutation_annotated_tree.cpp:415-508):
// the string is cut at commas; in each token '(' opens an internal node, the characters before the
// first ':' or ')' are the leaf name, labels after ')' are ignored (internal nodes are renamed
// node_1, node_2, ... in the order they open) and the branch-length queues per level are kept
// exactly (including the stale-length quirk of a ')' that is not followed by ':').
namespace {
// Set of the identifiers seen so far (the reference's all_nodes map only has to refuse duplicates here): open
// addressing over node indices, strings compared on equal hashes.
struct IdSet {
    std::vector<uint64_t> slot;   // hash tag << 32 | node index + 1; 0 = free
    uint64_t mask = 0;
    explicit IdSet(size_t n) {
        size_t cap = 16;
        while (cap < n * 2 + 2) cap <<= 1;
        slot.assign(cap, 0ull);
        mask = cap - 1;
    }
    static uint64_t hash(std::string_view s) {
        uint64_t h = 0x9E3779B97F4A7C15ull ^ s.size();
        size_t i = 0;
        for (; i + 8 <= s.size(); i += 8) {
            uint64_t x;
            std::memcpy(&x, s.data() + i, 8);
            h = (h ^ x) * 0xFF51AFD7ED558CCDull;
            h ^= h >> 32;
        }
        uint64_t x = 0;
        if (i < s.size()) std::memcpy(&x, s.data() + i, s.size() - i);
        h = (h ^ x) * 0xC4CEB9FE1A85EC53ull;
        return h ^ (h >> 29);
    }
    // false when `s` is already there; otherwise remembers it as node `idx` of `ids`
    bool insert(std::string_view s, int32_t idx, const std::vector<std::string>& ids) {
        const uint64_t h = hash(s), tag = h >> 32;
        for (uint64_t k = h & mask;; k = (k + 1) & mask) {
            if (slot[k] == 0ull) {
                slot[k] = (tag << 32) | (uint64_t)(uint32_t)(idx + 1);
                return true;
            }
            if ((slot[k] >> 32) == tag && ids[(size_t)(uint32_t)slot[k] - 1] == s) return false;
        }
    }
};
inline bool blen_char(char c) { return std::isdigit((unsigned char)c) || c == '.' || c == 'e' || c == 'E' || c == '-' || c == '+'; }
inline float blen_value(const std::string& s) { return s.empty() ? -1.0f : std::strtof(s.c_str(), nullptr); }
}  // namespace

// The token machine, one character at a time and in one thread: any Newick string (the fast path below hands the
// unusual ones back to it).
static std::string parse_newick_general(std::string_view nw, MatTree& t) {
    struct Tok { std::string leaf; int no, nc; };
    std::vector<Tok> toks;
    std::vector<std::queue<float>> blen(128);
    size_t level = 0;
    size_t a = 0;
    for (;;) {
        size_t b = nw.find(',', a);
        const size_t e = b == std::string_view::npos ? nw.size() : b;
        Tok tk{std::string(), 0, 0};
        bool stop = false, bstart = false;
        std::string branch;
        for (size_t i = a; i < e; ++i) {
            const char c = nw[i];
            if (c == ':') {
                stop = true;
                branch.clear();
                bstart = true;
            } else if (c == '(') {
                ++tk.no;
                ++level;
                if (blen.size() <= level) blen.resize(level * 2);
            } else if (c == ')') {
                stop = true;
                ++tk.nc;
                if (level == 0) return "incorrect Newick format";
                blen[level].push(blen_value(branch));
                --level;
                bstart = false;
            } else if (!stop) {
                tk.leaf += c;
                bstart = false;
            } else if (bstart) {
                if (blen_char(c)) branch += c;
            }
        }
        toks.push_back(std::move(tk));
        blen[level].push(blen_value(branch));
        if (b == std::string_view::npos) break;
        a = b + 1;
    }
    if (level != 0) return "incorrect Newick format";

    std::vector<int32_t> stack;
    IdSet seen(toks.size() * 2);
    auto create = [&](std::string&& id, float len) -> std::string {
        t.id.push_back(std::move(id));
        if (!seen.insert(t.id.back(), (int32_t)t.id.size() - 1, t.id)) return t.id.back() + " already in the tree";
        t.parent.push_back(stack.empty() ? -1 : stack.back());
        t.branch_length.push_back(len);
        return "";
    };
    for (Tok& tk : toks) {
        for (int j = 0; j < tk.no; ++j) {
            if (blen[level].empty()) return "incorrect Newick format";
            std::string err = create("node_" + std::to_string(++t.n_internal_ids), blen[level].front());
            if (!err.empty()) return err;
            blen[level].pop();
            ++level;
            stack.push_back((int32_t)t.parent.size() - 1);
        }
        if (stack.empty()) return "Newick tree without an internal node";
        if (blen[level].empty()) return "incorrect Newick format";
        std::string err = create(std::move(tk.leaf), blen[level].front());
        if (!err.empty()) return err;
        blen[level].pop();
        for (int j = 0; j < tk.nc; ++j) {
            if (stack.empty()) return "incorrect Newick format";
            stack.pop_back();
            --level;
        }
    }
    return "";
}

// Fast path for the usual shape of a token — '('s, then the leaf name, then ":length" / ")" / "):length" groups:
// the tokens are scanned by all host threads (name as a view into the string, one length per ')' and one for the
// token's end), then one thread replays the levels.  The result is that of parse_newick_general.
static std::string parse_newick(std::string_view nw, MatTree& t) {
    IoLaps lap("parse_newick");
    std::vector<size_t> cut;   // token i = [cut[i], cut[i + 1] - 1)
    cut.push_back(0);
    for (size_t a = 0;;) {
        const void* c = a < nw.size() ? std::memchr(nw.data() + a, ',', nw.size() - a) : nullptr;
        if (!c) break;
        a = (size_t)((const char*)c - nw.data()) + 1;
        cut.push_back(a);
    }
    const size_t n_tok = cut.size();
    cut.push_back(nw.size() + 1);
    lap("commas");
    struct Tok { uint32_t leaf_len, no, nc; size_t leaf_off, f_off; };
    std::vector<Tok> toks(n_tok);
    const int T = io_threads();
    std::vector<std::vector<float>> lens((size_t)T);
    std::vector<char> unusual((size_t)T, 0);
    parallel_ranges(n_tok, T, [&](int ti, size_t lo, size_t hi) {
        std::vector<float>& fl = lens[(size_t)ti];
        fl.reserve((hi - lo) * 2);
        std::string branch;
        for (size_t k = lo; k < hi; ++k) {
            Tok tk{0u, 0u, 0u, cut[k], fl.size()};
            bool stop = false, bstart = false, leaf_open = false;
            branch.clear();
            for (size_t i = cut[k]; i < cut[k + 1] - 1; ++i) {
                const char c = nw[i];
                if (c == ':') {
                    stop = true;
                    branch.clear();
                    bstart = true;
                } else if (c == '(') {
                    if (leaf_open || stop) unusual[(size_t)ti] = 1;   // a '(' after the name or after a ')': general path
                    ++tk.no;
                } else if (c == ')') {
                    stop = true;
                    ++tk.nc;
                    fl.push_back(blen_value(branch));
                    bstart = false;
                } else if (!stop) {
                    if (!leaf_open) {
                        leaf_open = true;
                        tk.leaf_off = i;
                    }
                    ++tk.leaf_len;
                    bstart = false;
                } else if (bstart) {
                    if (blen_char(c)) branch += c;
                }
            }
            fl.push_back(blen_value(branch));
            toks[k] = tk;
        }
    });
    lap("token scan");
    for (char u : unusual)
        if (u) return parse_newick_general(nw, t);
    // token k's lengths: lens[thread of k][f_off ...] — resolve the thread once per range
    const int Tn = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, (n_tok + 4095) / 4096));
    auto thread_of = [&](size_t k) {
        int ti = (int)((k * (size_t)Tn) / n_tok);
        while (ti + 1 < Tn && n_tok * (size_t)(ti + 1) / Tn <= k) ++ti;
        while (ti > 0 && n_tok * (size_t)ti / Tn > k) --ti;
        return ti;
    };
    // levels: per level a FIFO of branch lengths, filled over the whole string, then drained while the nodes are made
    std::vector<std::vector<float>> blen(128);
    size_t level = 0;
    for (size_t k = 0; k < n_tok; ++k) {
        const Tok& tk = toks[k];
        const float* f = lens[(size_t)thread_of(k)].data() + tk.f_off;
        level += tk.no;
        if (blen.size() <= level) blen.resize(level * 2);
        for (uint32_t j = 0; j < tk.nc; ++j) {
            if (level == 0) return "incorrect Newick format";
            blen[level].push_back(f[j]);
            --level;
        }
        blen[level].push_back(f[tk.nc]);
    }
    if (level != 0) return "incorrect Newick format";
    lap("levels");
    std::vector<size_t> head(blen.size(), 0);
    size_t n_nodes = n_tok;
    for (const Tok& tk : toks) n_nodes += tk.no;
    // structure first, by one thread and with integers only: parent, branch length and where each node's name comes
    // from (internal node number k > 0, or token -1 - index)
    t.parent.resize(n_nodes); t.branch_length.resize(n_nodes);
    std::vector<int64_t> name_src(n_nodes);
    std::vector<int32_t> stack;
    size_t made = 0;
    auto create = [&](int64_t src, size_t lvl) -> bool {
        if (head[lvl] >= blen[lvl].size()) return false;
        name_src[made] = src;
        t.parent[made] = stack.empty() ? -1 : stack.back();
        t.branch_length[made] = blen[lvl][head[lvl]++];
        ++made;
        return true;
    };
    for (size_t k = 0; k < n_tok; ++k) {
        const Tok& tk = toks[k];
        for (uint32_t j = 0; j < tk.no; ++j) {
            if (!create(++t.n_internal_ids, level)) return "incorrect Newick format";
            ++level;
            stack.push_back((int32_t)made - 1);
        }
        if (stack.empty()) return "Newick tree without an internal node";
        if (!create(-1 - (int64_t)k, level)) return "incorrect Newick format";
        for (uint32_t j = 0; j < tk.nc; ++j) {
            if (stack.empty()) return "incorrect Newick format";
            stack.pop_back();
            --level;
        }
    }
    lap("structure");
    // names and the duplicate check by all threads: every name goes into one lock-free table (hash tag | node); two
    // equal names anywhere raise the flag and the one-thread machine is run for the reference's message
    t.id.resize(n_nodes);
    size_t cap = 16;
    while (cap < n_nodes * 2 + 2) cap <<= 1;
    std::unique_ptr<std::atomic<uint64_t>[]> slot(new std::atomic<uint64_t>[cap]);
    parallel_ranges(cap, T, [&](int, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) slot[i].store(0ull, std::memory_order_relaxed);
    });
    parallel_ranges(n_nodes, T, [&](int, size_t lo, size_t hi) {
        char name[32] = "node_";
        for (size_t v = lo; v < hi; ++v) {
            if (name_src[v] > 0) {
                char* last = std::to_chars(name + 5, name + sizeof(name), (long long)name_src[v]).ptr;
                t.id[v].assign(name, (size_t)(last - name));
            } else {
                const Tok& tk = toks[(size_t)(-1 - name_src[v])];
                t.id[v].assign(nw.data() + tk.leaf_off, tk.leaf_len);
            }
        }
    });
    lap("names");
    std::atomic<int> duplicate{0};
    parallel_ranges(n_nodes, T, [&](int, size_t lo, size_t hi) {
        for (size_t v = lo; v < hi; ++v) {
            const uint64_t h = IdSet::hash(t.id[v]), mine = ((h >> 32) << 32) | (uint64_t)(uint32_t)(v + 1);
            for (uint64_t k = h & (cap - 1);; k = (k + 1) & (cap - 1)) {
                uint64_t cur = slot[k].load(std::memory_order_acquire);
                if (cur == 0ull && slot[k].compare_exchange_strong(cur, mine, std::memory_order_acq_rel)) break;
                // (a failed exchange left the winner in cur; its name was written before the barrier above)
                if ((cur >> 32) == (h >> 32) && t.id[(size_t)(uint32_t)cur - 1] == t.id[v]) {
                    duplicate.store(1);
                    break;
                }
            }
        }
    });
    if (duplicate.load()) {
        MatTree again;
        std::string err = parse_newick_general(nw, again);
        return err.empty() ? "duplicate node identifier" : err;
    }
    lap("nodes");
    return "";
}

// ---- Parsimony::data ----------------------------------------------------------------------------
namespace {
struct Mut { int32_t pos; uint8_t ref, par, nuc; };

// Node::add_mutation, src/mutation_annotated_tree.cpp:720-746
void add_mutation(std::vector<Mut>& muts, const Mut& m) {
    auto it = std::lower_bound(muts.begin(), muts.end(), m, [](const Mut& x, const Mut& y) { return x.pos < y.pos; });
    if (it != muts.end() && it->pos == m.pos) {
        if (it->par != m.nuc) it->nuc = m.nuc;   // update to the new allele
        else muts.erase(it);                     // reversal: the position drops out
    } else {
        muts.insert(it, m);
    }
}
}  // namespace

// One node's mutation_list -> the node's mutations (Node::add_mutation order rules, :556-596); false = malformed
static bool parse_node_mutations(std::string_view bytes, std::vector<Mut>& cur, const char*& err) {
    cur.clear();
    uint32_t f, w;
    pb::Reader r(bytes.data(), bytes.size());
    while (r.next(f, w)) {
        if (!(w == 2 && f == 1)) { r.skip(w); continue; }
        std::string_view mb = r.bytes();
        pb::Reader m(mb.data(), mb.size());
        int32_t pos = 0, ref = 0, par = 0;
        int nuc = 0;
        uint32_t mf, mw;
        while (m.next(mf, mw)) {
            if (mw == 0 && mf == 1) pos = (int32_t)m.varint();
            else if (mw == 0 && mf == 2) ref = (int32_t)m.varint();
            else if (mw == 0 && mf == 3) par = (int32_t)m.varint();
            else if (mw == 0 && mf == 4) nuc += 1 << ((int32_t)m.varint() & 7);          // unpacked repeated
            else if (mw == 2 && mf == 4) {                                               // packed repeated
                std::string_view pk = m.bytes();
                pb::Reader q(pk.data(), pk.size());
                while (!q.done()) nuc += 1 << ((int32_t)q.varint() & 7);
            } else m.skip(mw);
        }
        if (!m.ok) { err = "malformed mut"; return false; }
        Mut mu;
        mu.pos = pos;
        if (pos >= 0) {   // :566-580
            mu.ref = (uint8_t)(1 << (ref & 7));
            mu.par = (uint8_t)(1 << (par & 7));
            mu.nuc = (uint8_t)nuc;
            if (mu.nuc != mu.par) add_mutation(cur, mu);
        } else {          // masked mutation, :581-587
            mu.ref = mu.par = mu.nuc = 0;
            add_mutation(cur, mu);
        }
    }
    if (!r.ok) { err = "malformed mutation_list"; return false; }
    return true;
}

// The reference fills the nodes' mutations with a tbb::parallel_for over the preorder (:556-596); here the nodes are
// cut into one range per host thread, each range parsed into its own arrays and the arrays concatenated.
std::string parse_mat(const std::string& bytes, MatTree& t) {
    t = MatTree();
    IoLaps lap("parse_mat");
    pb::Reader top(bytes.data(), bytes.size());
    std::string_view newick;
    std::vector<std::string_view> node_muts, meta, condensed;
    uint32_t f, w;
    while (top.next(f, w)) {
        if (w == 2 && f == 1) newick = top.bytes();
        else if (w == 2 && f == 2) node_muts.push_back(top.bytes());
        else if (w == 2 && f == 3) condensed.push_back(top.bytes());
        else if (w == 2 && f == 4) meta.push_back(top.bytes());
        else top.skip(w);
    }
    if (!top.ok) return "malformed Parsimony::data";
    lap("top-level fields");
    std::string err = parse_newick(newick, t);
    if (!err.empty()) return err;
    lap("newick");
    const size_t n = t.parent.size();
    if (node_muts.size() < n) return "Parsimony::data has fewer node_mutations than Newick nodes";
    const bool hasmeta = !meta.empty();
    if (hasmeta && meta.size() < n) return "Parsimony::data has fewer metadata entries than Newick nodes";
    t.clade.assign(n, {});
    t.mut_off.assign(n + 1, 0);
    const int T = io_threads();
    struct Part {
        std::vector<Mut> muts;
        const char* err = nullptr;
    };
    std::vector<Part> parts((size_t)T);
    parallel_ranges(n, T, [&](int ti, size_t lo, size_t hi) {
        Part& pt = parts[(size_t)ti];
        std::vector<Mut> cur;
        uint32_t ff, ww;
        pt.muts.reserve((hi - lo) * 2);
        for (size_t v = lo; v < hi; ++v) {
            if (hasmeta) {
                pb::Reader r(meta[v].data(), meta[v].size());
                while (r.next(ff, ww)) {
                    if (ww == 2 && ff == 1) t.clade[v].emplace_back(r.bytes());
                    else r.skip(ww);
                }
                if (!r.ok) { pt.err = "malformed node_metadata"; return; }
            }
            if (!parse_node_mutations(node_muts[v], cur, pt.err)) return;
            pt.muts.insert(pt.muts.end(), cur.begin(), cur.end());
            t.mut_off[v + 1] = (int64_t)cur.size();   // counts now, offsets below
        }
    });
    for (const Part& pt : parts)
        if (pt.err) return pt.err;
    lap("node mutations + metadata");
    for (size_t v = 0; v < n; ++v) t.mut_off[v + 1] += t.mut_off[v];
    const size_t nm = (size_t)t.mut_off[n];
    t.mut_pos.resize(nm); t.mut_ref.resize(nm); t.mut_par.resize(nm); t.mut_nuc.resize(nm);
    {
        // the ranges parallel_ranges handed out, again: part ti starts at the offset of its first node
        const int Tn = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, (n + 4095) / 4096));
        std::vector<std::thread> th;
        for (int ti = 0; ti < Tn; ++ti)
            th.emplace_back([&, ti]() {
                size_t k = (size_t)t.mut_off[n * (size_t)ti / Tn];
                for (const Mut& mu : parts[(size_t)ti].muts) {
                    t.mut_pos[k] = mu.pos; t.mut_ref[k] = mu.ref; t.mut_par[k] = mu.par; t.mut_nuc[k] = mu.nuc;
                    ++k;
                }
            });
        for (auto& x : th) x.join();
    }
    lap("concatenate");
    t.n_annotations = n ? (int32_t)t.clade[0].size() : 0;
    for (std::string_view cb : condensed) {
        pb::Reader r(cb.data(), cb.size());
        std::string name;
        std::vector<std::string> leaves;
        while (r.next(f, w)) {
            if (w == 2 && f == 1) name = std::string(r.bytes());
            else if (w == 2 && f == 2) leaves.emplace_back(r.bytes());
            else r.skip(w);
        }
        if (!r.ok) return "malformed condensed_node";
        t.condensed_name.push_back(std::move(name));
        t.condensed_leaves.push_back(std::move(leaves));
    }
    lap("condensed nodes");
    return "";
}

// Tree::uncondense_leaves, src/mutation_annotated_tree.cpp:1224-1272.  The reference walks a
// tbb::concurrent_unordered_map, so the order in which condensed nodes are expanded — and with it the
// node_<k> ids handed to condensed nodes that carry mutations, and the order of the new siblings — is
// unspecified there; here it is the file order.
namespace {
// identifier -> node index for uncondense_leaves: the reference's all_nodes map restricted to what this pass does with
// it (look a name up, rename the node).  Open addressing over node indices, the keys are the strings in t.id
// themselves (std::unordered_map<std::string, int32_t> over 8 M ids costs several seconds).
struct IdIndex {
    std::vector<int32_t> slot;   // node index, -1 free, -2 deleted
    uint64_t mask = 0;
    const std::vector<std::string>& ids;
    IdIndex(const std::vector<std::string>& id, size_t expect) : ids(id) {
        size_t cap = 16;
        while (cap < expect * 2 + 2) cap <<= 1;
        slot.assign(cap, -1);
        mask = cap - 1;
    }
    int64_t find_slot(std::string_view name) const {
        for (uint64_t k = IdSet::hash(name) & mask;; k = (k + 1) & mask) {
            if (slot[k] == -1) return -1;
            if (slot[k] >= 0 && ids[(size_t)slot[k]] == name) return (int64_t)k;
        }
    }
    void put(int32_t v) {   // index[ids[v]] = v (a later node with the same identifier takes the name over)
        const std::string& name = ids[(size_t)v];
        int64_t first_free = -1;
        for (uint64_t k = IdSet::hash(name) & mask;; k = (k + 1) & mask) {
            if (slot[k] == -1) {
                slot[first_free >= 0 ? (uint64_t)first_free : k] = v;
                return;
            }
            if (slot[k] == -2) {
                if (first_free < 0) first_free = (int64_t)k;
            } else if (ids[(size_t)slot[k]] == name) {
                slot[k] = v;
                return;
            }
        }
    }
};
}  // namespace

void uncondense_leaves(MatTree& t) {
    if (t.condensed_name.empty()) return;
    IoLaps lap("uncondense_leaves");
    IdIndex index(t.id, t.id.size() + t.condensed_name.size());
    {   // all identifiers, by the host threads (a free slot is claimed with a compare-and-swap); two nodes with one
        // identifier — the later one owns the name — are left to one thread
        std::atomic<int> duplicate{0};
        parallel_ranges(t.id.size(), io_threads(), [&](int, size_t lo, size_t hi) {
            for (size_t v = lo; v < hi; ++v) {
                const std::string& name = t.id[v];
                for (uint64_t k = IdSet::hash(name) & index.mask;; k = (k + 1) & index.mask) {
                    const int32_t cur = __atomic_load_n(&index.slot[k], __ATOMIC_ACQUIRE);
                    if (cur == -1) {
                        int32_t expect = -1;
                        if (__atomic_compare_exchange_n(&index.slot[k], &expect, (int32_t)v, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) break;
                        if (t.id[(size_t)expect] == name) {
                            duplicate.store(1);
                            break;
                        }
                    } else if (t.id[(size_t)cur] == name) {
                        duplicate.store(1);
                        break;
                    }
                }
            }
        });
        if (duplicate.load()) {
            std::fill(index.slot.begin(), index.slot.end(), -1);
            for (size_t v = 0; v < t.id.size(); ++v) index.put((int32_t)v);
        }
    }
    lap("identifier index");
    struct NewNode { int32_t parent; std::string id; float len; };
    std::vector<NewNode> added;
    auto rename = [&](int64_t at, int32_t n, std::string id) {   // all_nodes.erase(old name); node->identifier = id; all_nodes[id] = node
        index.slot[(size_t)at] = -2;
        t.id[n] = std::move(id);
        index.put(n);
    };
    for (size_t c = 0; c < t.condensed_name.size(); ++c) {
        const int64_t at = index.find_slot(t.condensed_name[c]);
        if (at < 0) continue;
        const int32_t n = index.slot[(size_t)at];
        const int32_t par = t.parent[n] >= 0 ? t.parent[n] : n;
        const std::vector<std::string>& s = t.condensed_leaves[c];
        const bool has_muts = t.mut_off[n + 1] > t.mut_off[n];
        if (s.size() > 1 && has_muts) {
            rename(at, n, "node_" + std::to_string(++t.n_internal_ids));
            for (const std::string& leaf : s) added.push_back({n, leaf, -1.0f});
        } else if (s.size() > 1) {
            rename(at, n, s[0]);
            for (size_t k = 1; k < s.size(); ++k) added.push_back({par, s[k], t.branch_length[n]});
        } else if (s.size() == 1) {
            rename(at, n, s[0]);
        }
    }
    lap("condensed nodes");
    t.parent.reserve(t.parent.size() + added.size());
    t.id.reserve(t.id.size() + added.size());
    for (NewNode& nn : added) {
        t.parent.push_back(nn.parent);
        t.id.push_back(std::move(nn.id));
        t.branch_length.push_back(nn.len);
        t.clade.emplace_back((size_t)t.n_annotations, std::string());
        t.mut_off.push_back(t.mut_off.back());
    }
    t.condensed_name.clear();
    t.condensed_leaves.clear();
    lap("new leaves");
}

// ---- flattened-tree sidecar ----------------------------------------------------------------------
// A parsed MAT (before uncondense_leaves) written next to the file it came from as "<file>.wepp_flat" — or into
// $WEPP_SIDECAR_DIR — and read back instead of inflating and parsing the protobuf again: the same tree serves
// every sample of a run (workflow/rules/filter.smk:21-36 passes one MAT to all of them).  The sidecar names its
// source by size and a 64-bit hash of the source's bytes; any mismatch, a short file or another layout version and the
// source is parsed (and the sidecar rewritten).  WEPP_SIDECAR=0 neither reads nor writes one;
// a directory that cannot be written to is not an error.
namespace {
constexpr char SIDECAR_MAGIC[8] = {'W', 'E', 'P', 'P', 'F', 'L', 'T', '2'};
struct SidecarHeader {
    char magic[8];
    uint64_t src_size, src_hash;
    int64_t src_mtime_ns;
    uint64_t n_nodes, n_muts, n_internal_ids, n_condensed;
    int64_t n_annotations;
    uint64_t n_blobs;
};
uint64_t hash_bytes(const std::string& raw) {   // chunks hashed by the host threads, the chunk hashes mixed in order
    const size_t CH = 8u << 20, nch = (raw.size() + CH - 1) / CH;
    std::vector<uint64_t> hs(nch);
    parallel_ranges(nch, io_threads(), [&](int, size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; ++c) hs[c] = IdSet::hash(std::string_view(raw.data() + c * CH, std::min(CH, raw.size() - c * CH)));
    });
    uint64_t h = 0x243F6A8885A308D3ull ^ raw.size();
    for (uint64_t x : hs) h = (h ^ x) * 0x9E3779B97F4A7C15ull + (h >> 31);
    return h;
}
std::string sidecar_path(const std::string& path) {
    if (const char* d = getenv("WEPP_SIDECAR_DIR")) {
        const size_t sl = path.find_last_of('/');
        return std::string(d) + "/" + (sl == std::string::npos ? path : path.substr(sl + 1)) + ".wepp_flat";
    }
    return path + ".wepp_flat";
}
bool sidecar_enabled() { return !(getenv("WEPP_SIDECAR") && atoi(getenv("WEPP_SIDECAR")) == 0); }

struct BlobWriter {
    std::string out;
    uint64_t n = 0;
    void raw(const void* p, size_t bytes) {
        const uint64_t b = bytes;
        out.append((const char*)&b, 8);
        if (bytes) out.append((const char*)p, bytes);
        out.append((8 - bytes % 8) % 8, '\0');
        ++n;
    }
    template <typename T> void vec(const std::vector<T>& v) { raw(v.data(), v.size() * sizeof(T)); }
    // strings as an offset array and the characters back to back
    template <typename It> void strings(It first, It last) {
        std::vector<int64_t> off{0};
        std::string chars;
        for (It it = first; it != last; ++it) {
            chars.append(*it);
            off.push_back((int64_t)chars.size());
        }
        vec(off);
        raw(chars.data(), chars.size());
    }
};
struct BlobReader {
    const char* p;
    const char* end;
    bool ok = true;
    std::string_view next() {
        if (!ok || end - p < 8) { ok = false; return {}; }
        uint64_t b;
        std::memcpy(&b, p, 8);
        p += 8;
        const uint64_t padded = b + (8 - b % 8) % 8;
        if (padded > (uint64_t)(end - p)) { ok = false; return {}; }
        std::string_view v(p, (size_t)b);
        p += padded;
        return v;
    }
    template <typename T> bool vec(std::vector<T>& v, size_t expect) {
        std::string_view b = next();
        if (!ok || b.size() != expect * sizeof(T)) return ok = false;
        v.resize(expect);
        if (expect) std::memcpy(v.data(), b.data(), b.size());
        return true;
    }
    // `count` strings; f(i, string_view) is called by the host threads
    template <typename F> bool strings(size_t count, F f) {
        std::vector<int64_t> off;
        if (!vec(off, count + 1)) return false;
        std::string_view chars = next();
        if (!ok || off[0] != 0 || (size_t)off[count] != chars.size()) return ok = false;
        for (size_t i = 0; i < count; ++i)
            if (off[i + 1] < off[i]) return ok = false;
        parallel_ranges(count, io_threads(), [&](int, size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) f(i, chars.substr((size_t)off[i], (size_t)(off[i + 1] - off[i])));
        });
        return true;
    }
};

void write_sidecar(const std::string& file, const SidecarHeader& key, const MatTree& t) {
    BlobWriter w;
    w.vec(t.parent); w.vec(t.branch_length); w.vec(t.mut_off); w.vec(t.mut_pos); w.vec(t.mut_ref); w.vec(t.mut_par); w.vec(t.mut_nuc);
    w.strings(t.id.begin(), t.id.end());
    std::vector<int64_t> coff{0};
    std::vector<std::string_view> flat;
    for (const auto& c : t.clade) {
        for (const std::string& x : c) flat.push_back(x);
        coff.push_back((int64_t)flat.size());
    }
    w.vec(coff);
    w.strings(flat.begin(), flat.end());
    w.strings(t.condensed_name.begin(), t.condensed_name.end());
    std::vector<int64_t> loff{0};
    flat.clear();
    for (const auto& c : t.condensed_leaves) {
        for (const std::string& x : c) flat.push_back(x);
        loff.push_back((int64_t)flat.size());
    }
    w.vec(loff);
    w.strings(flat.begin(), flat.end());
    SidecarHeader h = key;
    std::memcpy(h.magic, SIDECAR_MAGIC, 8);
    h.n_nodes = t.parent.size(); h.n_muts = t.mut_pos.size(); h.n_internal_ids = (uint64_t)t.n_internal_ids;
    h.n_condensed = t.condensed_name.size(); h.n_annotations = t.n_annotations; h.n_blobs = w.n;
    const std::string tmp = file + "." + std::to_string((long long)getpid()) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;   // read-only data directory: no sidecar
    const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(w.out.data(), 1, w.out.size(), f) == w.out.size();
    if (fclose(f) != 0 || !ok || rename(tmp.c_str(), file.c_str()) != 0) remove(tmp.c_str());
}

bool read_sidecar(const std::string& file, const SidecarHeader& key, MatTree& t) {
    std::string bytes;
    if (!slurp(file, bytes).empty() || bytes.size() < sizeof(SidecarHeader)) return false;
    SidecarHeader h;
    std::memcpy(&h, bytes.data(), sizeof(h));
    // (the source's modification time is recorded but not compared: a copied data directory keeps a valid sidecar)
    if (std::memcmp(h.magic, SIDECAR_MAGIC, 8) != 0 || h.src_size != key.src_size || h.src_hash != key.src_hash ||
        h.n_nodes >= (1ull << 31) || h.n_annotations < 0)
        return false;
    t = MatTree();
    BlobReader r{bytes.data() + sizeof(h), bytes.data() + bytes.size()};
    const size_t n = (size_t)h.n_nodes, nm = (size_t)h.n_muts, nc = (size_t)h.n_condensed;
    if (!r.vec(t.parent, n) || !r.vec(t.branch_length, n) || !r.vec(t.mut_off, n + 1) || !r.vec(t.mut_pos, nm) ||
        !r.vec(t.mut_ref, nm) || !r.vec(t.mut_par, nm) || !r.vec(t.mut_nuc, nm))
        return false;
    if (t.mut_off[0] != 0 || (size_t)t.mut_off[n] != nm) return false;
    for (size_t v = 0; v < n; ++v)
        if (t.mut_off[v + 1] < t.mut_off[v] || t.parent[v] >= (int32_t)v || (v > 0 && t.parent[v] < 0)) return false;
    t.id.resize(n);
    if (!r.strings(n, [&](size_t i, std::string_view x) { t.id[i].assign(x); })) return false;
    std::vector<int64_t> coff;
    if (!r.vec(coff, n + 1) || coff[0] != 0) return false;
    for (size_t v = 0; v < n; ++v)
        if (coff[v + 1] < coff[v]) return false;
    std::vector<std::string> flat((size_t)coff[n]);
    if (!r.strings(flat.size(), [&](size_t i, std::string_view x) { flat[i].assign(x); })) return false;
    t.clade.resize(n);
    parallel_ranges(n, io_threads(), [&](int, size_t lo, size_t hi) {
        for (size_t v = lo; v < hi; ++v)
            for (int64_t k = coff[v]; k < coff[v + 1]; ++k) t.clade[v].push_back(std::move(flat[(size_t)k]));
    });
    t.condensed_name.resize(nc);
    if (!r.strings(nc, [&](size_t i, std::string_view x) { t.condensed_name[i].assign(x); })) return false;
    std::vector<int64_t> loff;
    if (!r.vec(loff, nc + 1) || loff[0] != 0) return false;
    for (size_t c = 0; c < nc; ++c)
        if (loff[c + 1] < loff[c]) return false;
    flat.assign((size_t)loff[nc], std::string());
    if (!r.strings(flat.size(), [&](size_t i, std::string_view x) { flat[i].assign(x); })) return false;
    t.condensed_leaves.resize(nc);
    for (size_t c = 0; c < nc; ++c)
        for (int64_t k = loff[c]; k < loff[c + 1]; ++k) t.condensed_leaves[c].push_back(std::move(flat[(size_t)k]));
    t.n_annotations = (int32_t)h.n_annotations;
    t.n_internal_ids = (int64_t)h.n_internal_ids;
    return r.ok && h.n_blobs == 17 && r.p == r.end;
}
}  // namespace

std::string load_mat(const std::string& path, bool uncondense, MatTree& out) {
    IoLaps lap("load_mat");
    std::string raw, bytes;
    std::string err = slurp(path, raw);
    if (!err.empty()) return "Could not load the mutation-annotated tree object from file: " + path + " (" + err + ")";
    lap("read file");
    SidecarHeader key = {};
    const bool cache = sidecar_enabled();
    bool from_sidecar = false;
    if (cache) {
        struct stat sb;
        if (stat(path.c_str(), &sb) == 0) key.src_mtime_ns = (int64_t)sb.st_mtim.tv_sec * 1000000000ll + sb.st_mtim.tv_nsec;
        key.src_size = raw.size();
        key.src_hash = hash_bytes(raw);
        lap("hash of the source");
        from_sidecar = read_sidecar(sidecar_path(path), key, out);
        lap(from_sidecar ? "sidecar read" : "no usable sidecar");
    }
    if (!from_sidecar) {
        err = inflate_if_gz(path, raw, bytes);
        if (!err.empty()) return "Could not load the mutation-annotated tree object from file: " + path + " (" + err + ")";
        std::string().swap(raw);
        lap("inflate");
        err = parse_mat(bytes, out);
        if (!err.empty()) return err;
        if (cache) {
            write_sidecar(sidecar_path(path), key, out);
            lap("sidecar written");
        }
    }
    if (uncondense) uncondense_leaves(out);
    lap("uncondense");
    return "";
}

// Newick of the tree in index order with leaf names only (what the loader needs back)
static void newick_rec(const MatTree& t, const std::vector<std::vector<int32_t>>& ch, int32_t v, std::string& out) {
    // iterative to survive deep trees
    struct Frame { int32_t v; size_t k; };
    std::vector<Frame> st{{v, 0}};
    while (!st.empty()) {
        Frame& fr = st.back();
        const auto& c = ch[fr.v];
        if (c.empty()) {
            out += t.id[fr.v];
            st.pop_back();
            continue;
        }
        if (fr.k == 0) out += '(';
        if (fr.k == c.size()) {
            out += ')';
            st.pop_back();
            continue;
        }
        if (fr.k > 0) out += ',';
        const int32_t child = c[fr.k++];
        st.push_back({child, 0});
    }
}

std::string serialize_mat(const MatTree& t) {
    const int32_t n = t.n_nodes();
    std::vector<std::vector<int32_t>> ch((size_t)n);
    for (int32_t v = 1; v < n; ++v) ch[t.parent[v]].push_back(v);
    std::string nw;
    if (n) newick_rec(t, ch, 0, nw);
    nw += ';';
    // preorder (children in index order) — node_mutations / metadata are given in this order
    std::vector<int32_t> order;
    order.reserve(n);
    std::vector<int32_t> st;
    if (n) st.push_back(0);
    while (!st.empty()) {
        const int32_t v = st.back();
        st.pop_back();
        order.push_back(v);
        for (size_t k = ch[v].size(); k-- > 0;) st.push_back(ch[v][k]);
    }
    pb::Writer top;
    top.str(1, nw);
    for (int32_t v : order) {
        pb::Writer ml;
        for (int64_t k = t.mut_off[v]; k < t.mut_off[v + 1]; ++k) {
            pb::Writer m;
            auto bit = [](uint8_t x) { int b = 0; while (b < 3 && !((x >> b) & 1)) ++b; return b; };
            m.int32(1, t.mut_pos[k]);
            m.int32(2, bit(t.mut_ref[k]));
            m.int32(3, bit(t.mut_par[k]));
            pb::Writer pk;
            for (int b = 0; b < 4; ++b)
                if ((t.mut_nuc[k] >> b) & 1) pk.varint((uint64_t)b);
            m.str(4, pk.out);
            ml.message(1, m.out);
        }
        top.message(2, ml.out);
    }
    for (size_t c = 0; c < t.condensed_name.size(); ++c) {
        pb::Writer cn;
        cn.str(1, t.condensed_name[c]);
        for (const std::string& s : t.condensed_leaves[c]) cn.str(2, s, true);
        top.message(3, cn.out);
    }
    if (t.n_annotations > 0 || !t.clade.empty()) {
        for (int32_t v : order) {
            pb::Writer md;
            if ((size_t)v < t.clade.size())
                for (const std::string& s : t.clade[v]) md.str(1, s, true);
            top.message(4, md.out);
        }
    }
    return std::move(top.out);
}

// ---- Sam::sam -----------------------------------------------------------------------------------
std::string parse_reads(const std::string& bytes, const std::string& reference, ReadSet& out, int n_threads) {
    out = ReadSet();
    pb::Reader top(bytes.data(), bytes.size());
    std::vector<std::string_view> reads, cols;
    uint32_t f, w;
    while (top.next(f, w)) {
        if (w == 2 && f == 1) reads.push_back(top.bytes());
        else if (w == 2 && f == 2) cols.push_back(top.bytes());
        else top.skip(w);
    }
    if (!top.ok) return "malformed Sam::sam";
    const size_t n = reads.size();
    std::vector<std::string_view> name(n), content(n);
    out.start.assign(n, 0);
    out.end.assign(n, 0);
    out.degree.assign(n, 0);
    std::vector<int64_t> cnt(n + 1, 0);
    std::string err;
    // pass 1 (threaded): fields and the number of mutations per read (src/WEPP/sam2pb.cpp:508-536)
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, n_threads), n / 4096 + 1));
    auto span = [&](int t, size_t& lo, size_t& hi) { lo = n * (size_t)t / nt; hi = n * (size_t)(t + 1) / nt; };
    std::vector<std::string> errs((size_t)nt);
    auto pass1 = [&](int t) {
        size_t lo, hi;
        span(t, lo, hi);
        for (size_t i = lo; i < hi; ++i) {
            pb::Reader r(reads[i].data(), reads[i].size());
            uint32_t rf, rw;
            int32_t start = 0, degree = 0;
            while (r.next(rf, rw)) {
                if (rw == 2 && rf == 1) name[i] = r.bytes();
                else if (rw == 0 && rf == 3) start = (int32_t)r.varint();
                else if (rw == 2 && rf == 6) content[i] = r.bytes();
                else if (rw == 0 && rf == 5) degree = (int32_t)r.varint();
                else r.skip(rw);
            }
            if (!r.ok) { errs[t] = "malformed read_info"; return; }
            out.start[i] = start;
            out.end[i] = start + (int32_t)content[i].size() - 1;
            out.degree[i] = degree;
            if (start < 1 || (size_t)out.end[i] > reference.size()) {
                if (!content[i].empty()) { errs[t] = "read " + std::string(name[i]) + " lies outside the reference"; return; }
            }
            int64_t c = 0;
            const char* ref = reference.data() + start - 1;
            for (size_t k = 0; k < content[i].size(); ++k) c += content[i][k] != ref[k] && content[i][k] != '_';
            cnt[i + 1] = c;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(pass1, t);
        pass1(0);
        for (auto& x : th) x.join();
    }
    for (const std::string& e : errs)
        if (!e.empty()) return e;
    for (size_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    out.rm_off = cnt;
    out.rm_pos.assign((size_t)cnt[n], 0);
    out.rm_nuc.assign((size_t)cnt[n], 0);
    auto pass2 = [&](int t) {
        size_t lo, hi;
        span(t, lo, hi);
        for (size_t i = lo; i < hi; ++i) {
            int64_t o = out.rm_off[i];
            const char* ref = reference.data() + out.start[i] - 1;
            for (size_t k = 0; k < content[i].size(); ++k) {
                const char c = content[i][k];
                if (c != ref[k] && c != '_') {
                    out.rm_pos[o] = out.start[i] + (int32_t)k;
                    out.rm_nuc[o] = nuc_id(c);
                    ++o;
                }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(pass2, t);
        pass2(0);
        for (auto& x : th) x.join();
    }
    out.name.off.reserve(n + 1);
    for (size_t i = 0; i < n; ++i) out.name.push(name[i]);
    // reverse merge table: repeated keys append (reverse_merge[name].push_back, :539-544)
    std::unordered_map<std::string, std::vector<std::string_view>> rev;
    std::vector<std::string> order;
    for (std::string_view cb : cols) {
        pb::Reader r(cb.data(), cb.size());
        std::string key;
        std::vector<std::string_view> vals;
        while (r.next(f, w)) {
            if (w == 2 && f == 1) key = std::string(r.bytes());
            else if (w == 2 && f == 2) vals.push_back(r.bytes());
            else r.skip(w);
        }
        if (!r.ok) return "malformed column_info";
        if (vals.empty()) continue;   // operator[] is only reached inside the input_columns loop
        auto ins = rev.emplace(key, std::vector<std::string_view>());
        if (ins.second) order.push_back(key);
        ins.first->second.insert(ins.first->second.end(), vals.begin(), vals.end());
    }
    for (const std::string& k : order) {
        out.rev_key.push(k);
        for (std::string_view v : rev[k]) out.rev_val.push(v);
        out.rev_off.push_back((int64_t)out.rev_val.size());
    }
    return "";
}

std::string load_reads(const std::string& path, const std::string& reference, ReadSet& out, int n_threads) {
    std::string bytes;
    std::string err = slurp(path, bytes);
    if (!err.empty()) return "Could not load the read protobuf from file: " + path;
    return parse_reads(bytes, reference, out, n_threads);
}

// ---- FASTA / mask.bed ---------------------------------------------------------------------------
std::string load_fasta(const std::string& path, std::string& name, std::string& seq) {
    std::ifstream f(path);
    if (!f.is_open()) return "Unable to open file " + path;
    std::string line;
    std::getline(f, line);
    if (line.empty() || line[0] != '>') return "Fasta format NOT correct " + path;
    std::istringstream iss(line.substr(1));
    name.clear();
    iss >> name;
    seq.clear();
    while (std::getline(f, line)) {
        for (char& c : line) c = (char)std::toupper((unsigned char)c);
        seq += line;
    }
    return "";
}

std::vector<int32_t> load_mask_bed(const std::string& path) {
    std::vector<int32_t> out;
    std::ifstream f(path);
    if (!f.is_open()) return out;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ls(line);
        std::string c1, c2;
        int c3;
        if (ls >> c1 >> c2 >> c3) out.push_back(c3);
    }
    return out;
}

}  // namespace wepp
