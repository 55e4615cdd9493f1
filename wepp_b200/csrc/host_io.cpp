// host_io.cpp — MAT / collapsed-read / FASTA / mask.bed loaders (see host_io.h for the reference lines).
#include "host_io.h"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <queue>
#include <sstream>
#include <thread>

#include "pbwire.h"

namespace wepp {

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
std::string read_file_maybe_gz(const std::string& path, std::string& out) {
    std::string raw;
    std::string err = slurp(path, raw);
    if (!err.empty()) return err;
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
static std::string parse_newick(const std::string& nw, MatTree& t) {
    struct Tok { std::string leaf; int no, nc; };
    std::vector<Tok> toks;
    std::vector<std::queue<float>> blen(128);
    size_t level = 0;
    size_t a = 0;
    auto to_f = [](const std::string& s) { return s.empty() ? -1.0f : std::strtof(s.c_str(), nullptr); };
    for (;;) {
        size_t b = nw.find(',', a);
        const size_t e = b == std::string::npos ? nw.size() : b;
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
                blen[level].push(to_f(branch));
                --level;
                bstart = false;
            } else if (!stop) {
                tk.leaf += c;
                bstart = false;
            } else if (bstart) {
                if (std::isdigit((unsigned char)c) || c == '.' || c == 'e' || c == 'E' || c == '-' || c == '+') branch += c;
            }
        }
        toks.push_back(std::move(tk));
        blen[level].push(to_f(branch));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    if (level != 0) return "incorrect Newick format";

    std::unordered_map<std::string, int32_t> seen;
    seen.reserve(toks.size() * 2);
    std::vector<int32_t> stack;
    auto create = [&](const std::string& id, float len) -> std::string {
        if (!seen.emplace(id, (int32_t)t.parent.size()).second) return id + " already in the tree";
        t.parent.push_back(stack.empty() ? -1 : stack.back());
        t.id.push_back(id);
        t.branch_length.push_back(len);
        return "";
    };
    for (const Tok& tk : toks) {
        for (int j = 0; j < tk.no; ++j) {
            const std::string nid = "node_" + std::to_string(++t.n_internal_ids);
            if (blen[level].empty()) return "incorrect Newick format";
            std::string err = create(nid, blen[level].front());
            if (!err.empty()) return err;
            blen[level].pop();
            ++level;
            stack.push_back((int32_t)t.parent.size() - 1);
        }
        if (stack.empty()) return "Newick tree without an internal node";
        if (blen[level].empty()) return "incorrect Newick format";
        std::string err = create(tk.leaf, blen[level].front());
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

std::string parse_mat(const std::string& bytes, MatTree& t) {
    t = MatTree();
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
    std::string err = parse_newick(std::string(newick), t);
    if (!err.empty()) return err;
    const size_t n = t.parent.size();
    if (node_muts.size() < n) return "Parsimony::data has fewer node_mutations than Newick nodes";
    const bool hasmeta = !meta.empty();
    if (hasmeta && meta.size() < n) return "Parsimony::data has fewer metadata entries than Newick nodes";
    t.clade.assign(n, {});
    t.mut_off.assign(n + 1, 0);
    std::vector<Mut> cur;
    for (size_t v = 0; v < n; ++v) {
        if (hasmeta) {
            pb::Reader r(meta[v].data(), meta[v].size());
            while (r.next(f, w)) {
                if (w == 2 && f == 1) t.clade[v].emplace_back(r.bytes());
                else r.skip(w);
            }
            if (!r.ok) return "malformed node_metadata";
        }
        cur.clear();
        pb::Reader r(node_muts[v].data(), node_muts[v].size());
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
            if (!m.ok) return "malformed mut";
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
        if (!r.ok) return "malformed mutation_list";
        for (const Mut& mu : cur) {
            t.mut_pos.push_back(mu.pos);
            t.mut_ref.push_back(mu.ref);
            t.mut_par.push_back(mu.par);
            t.mut_nuc.push_back(mu.nuc);
        }
        t.mut_off[v + 1] = (int64_t)t.mut_pos.size();
    }
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
    return "";
}

// Tree::uncondense_leaves, src/mutation_annotated_tree.cpp:1224-1272.  The reference walks a
// tbb::concurrent_unordered_map, so the order in which condensed nodes are expanded — and with it the
// node_<k> ids handed to condensed nodes that carry mutations, and the order of the new siblings — is
// unspecified there; here it is the file order.
void uncondense_leaves(MatTree& t) {
    if (t.condensed_name.empty()) return;
    std::unordered_map<std::string, int32_t> index;
    index.reserve(t.id.size() * 2);
    for (size_t v = 0; v < t.id.size(); ++v) index[t.id[v]] = (int32_t)v;
    struct NewNode { int32_t parent; std::string id; float len; };
    std::vector<NewNode> added;
    for (size_t c = 0; c < t.condensed_name.size(); ++c) {
        auto it = index.find(t.condensed_name[c]);
        if (it == index.end()) continue;
        const int32_t n = it->second;
        const int32_t par = t.parent[n] >= 0 ? t.parent[n] : n;
        const std::vector<std::string>& s = t.condensed_leaves[c];
        const bool has_muts = t.mut_off[n + 1] > t.mut_off[n];
        if (s.size() > 1 && has_muts) {
            index.erase(it);
            t.id[n] = "node_" + std::to_string(++t.n_internal_ids);
            index[t.id[n]] = n;
            for (const std::string& leaf : s) added.push_back({n, leaf, -1.0f});
        } else if (s.size() > 1) {
            index.erase(it);
            t.id[n] = s[0];
            index[t.id[n]] = n;
            for (size_t k = 1; k < s.size(); ++k) added.push_back({par, s[k], t.branch_length[n]});
        } else if (s.size() == 1) {
            index.erase(it);
            t.id[n] = s[0];
            index[t.id[n]] = n;
        }
    }
    for (NewNode& nn : added) {
        t.parent.push_back(nn.parent);
        t.id.push_back(std::move(nn.id));
        t.branch_length.push_back(nn.len);
        t.clade.emplace_back((size_t)t.n_annotations, std::string());
        t.mut_off.push_back(t.mut_off.back());
    }
    t.condensed_name.clear();
    t.condensed_leaves.clear();
}

std::string load_mat(const std::string& path, bool uncondense, MatTree& out) {
    std::string bytes;
    std::string err = read_file_maybe_gz(path, bytes);
    if (!err.empty()) return "Could not load the mutation-annotated tree object from file: " + path + " (" + err + ")";
    err = parse_mat(bytes, out);
    if (!err.empty()) return err;
    if (uncondense) uncondense_leaves(out);
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
