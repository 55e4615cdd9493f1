// wepp_abi_io.cpp — C ABI of the file-format loaders (include/wepp_b200.h, "File formats" section).
#include <cstring>

#include "../../include/wepp_b200.h"
#include "abi_internal.h"
#include "host_io.h"

using namespace wepp;

struct wepp_mat {
    MatTree t;
    StringPool ids, clades;
};
struct wepp_readset {
    ReadSet r;
};

namespace {
int finish_mat(wepp_mat* m, const std::string& err, int32_t uncondense, wepp_mat** out) {
    if (!err.empty()) {
        delete m;
        return abi_fail(WEPP_E_INVALID, err);
    }
    if (uncondense) uncondense_leaves(m->t);
    for (const std::string& s : m->t.id) m->ids.push(s);
    const int32_t na = m->t.n_annotations;
    for (size_t v = 0; v < m->t.id.size(); ++v)
        for (int32_t k = 0; k < na; ++k)
            m->clades.push(v < m->t.clade.size() && (size_t)k < m->t.clade[v].size() ? m->t.clade[v][k] : std::string());
    *out = m;
    return WEPP_OK;
}
template <typename T>
void put(T* dst, const std::vector<T>& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
}
void put_pool(int64_t* off, char* chars, const StringPool& p) {
    put(off, p.off);
    if (chars && !p.chars.empty()) std::memcpy(chars, p.chars.data(), p.chars.size());
}
}  // namespace

extern "C" {

int wepp_mat_load(const char* path, int32_t uncondense, wepp_mat** out) {
    if (!path || !out) return abi_fail(WEPP_E_INVALID, "NULL argument");
    wepp_mat* m = new wepp_mat();
    return finish_mat(m, load_mat(path, false, m->t), uncondense, out);
}

int wepp_mat_parse(const void* pb_bytes, int64_t n_bytes, int32_t uncondense, wepp_mat** out) {
    if (!pb_bytes || n_bytes < 0 || !out) return abi_fail(WEPP_E_INVALID, "NULL argument");
    wepp_mat* m = new wepp_mat();
    return finish_mat(m, parse_mat(std::string((const char*)pb_bytes, (size_t)n_bytes), m->t), uncondense, out);
}

void wepp_mat_free(wepp_mat* m) { delete m; }

int wepp_mat_dims(const wepp_mat* m, int32_t* n_nodes, int64_t* n_muts, int32_t* n_annotations, int64_t* id_chars,
                  int64_t* clade_chars) {
    if (!m) return abi_fail(WEPP_E_INVALID, "mat is NULL");
    if (n_nodes) *n_nodes = m->t.n_nodes();
    if (n_muts) *n_muts = (int64_t)m->t.mut_pos.size();
    if (n_annotations) *n_annotations = m->t.n_annotations;
    if (id_chars) *id_chars = (int64_t)m->ids.chars.size();
    if (clade_chars) *clade_chars = (int64_t)m->clades.chars.size();
    return WEPP_OK;
}

int wepp_mat_get(const wepp_mat* m, int32_t* parent, int64_t* mut_off, int32_t* mut_pos, uint8_t* mut_ref,
                 uint8_t* mut_par, uint8_t* mut_nuc, int64_t* id_off, char* id_chars) {
    if (!m) return abi_fail(WEPP_E_INVALID, "mat is NULL");
    put(parent, m->t.parent); put(mut_off, m->t.mut_off); put(mut_pos, m->t.mut_pos); put(mut_ref, m->t.mut_ref);
    put(mut_par, m->t.mut_par); put(mut_nuc, m->t.mut_nuc);
    put_pool(id_off, id_chars, m->ids);
    return WEPP_OK;
}

int wepp_mat_get_clades(const wepp_mat* m, int64_t* clade_off, char* clade_chars) {
    if (!m) return abi_fail(WEPP_E_INVALID, "mat is NULL");
    put_pool(clade_off, clade_chars, m->clades);
    return WEPP_OK;
}

int64_t wepp_mat_serialize(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                           const uint8_t* mut_ref, const uint8_t* mut_par, const uint8_t* mut_nuc, const int64_t* id_off,
                           const char* id_chars, void* out, int64_t capacity) {
    if (n_nodes < 1 || !parent || !mut_off || !id_off || !id_chars) return abi_fail(WEPP_E_INVALID, "NULL argument");
    MatTree t;
    t.parent.assign(parent, parent + n_nodes);
    t.branch_length.assign((size_t)n_nodes, -1.0f);
    t.mut_off.assign(mut_off, mut_off + n_nodes + 1);
    const int64_t nm = mut_off[n_nodes];
    if (nm > 0 && (!mut_pos || !mut_ref || !mut_par || !mut_nuc)) return abi_fail(WEPP_E_INVALID, "mutation arrays are NULL");
    t.mut_pos.assign(mut_pos, mut_pos + nm);
    t.mut_ref.assign(mut_ref, mut_ref + nm);
    t.mut_par.assign(mut_par, mut_par + nm);
    t.mut_nuc.assign(mut_nuc, mut_nuc + nm);
    for (int32_t v = 0; v < n_nodes; ++v) {
        if (v > 0 && (parent[v] < 0 || parent[v] >= v)) return abi_fail(WEPP_E_INVALID, "parent[v] must be < v");
        t.id.emplace_back(id_chars + id_off[v], (size_t)(id_off[v + 1] - id_off[v]));
    }
    const std::string bytes = serialize_mat(t);
    if (out && capacity > 0) std::memcpy(out, bytes.data(), (size_t)std::min<int64_t>(capacity, (int64_t)bytes.size()));
    return (int64_t)bytes.size();
}

int wepp_reads_load(const char* path, const char* ref_seq, int64_t ref_len, int32_t n_threads, wepp_readset** out) {
    if (!path || !ref_seq || ref_len < 0 || !out) return abi_fail(WEPP_E_INVALID, "NULL argument");
    wepp_readset* r = new wepp_readset();
    const std::string err = load_reads(path, std::string(ref_seq, (size_t)ref_len), r->r, n_threads);
    if (!err.empty()) {
        delete r;
        return abi_fail(WEPP_E_INVALID, err);
    }
    *out = r;
    return WEPP_OK;
}

int wepp_reads_parse(const void* pb_bytes, int64_t n_bytes, const char* ref_seq, int64_t ref_len, int32_t n_threads,
                     wepp_readset** out) {
    if (!pb_bytes || n_bytes < 0 || !ref_seq || ref_len < 0 || !out) return abi_fail(WEPP_E_INVALID, "NULL argument");
    wepp_readset* r = new wepp_readset();
    const std::string bytes((const char*)pb_bytes, (size_t)n_bytes);
    const std::string err = parse_reads(bytes, std::string(ref_seq, (size_t)ref_len), r->r, n_threads);
    if (!err.empty()) {
        delete r;
        return abi_fail(WEPP_E_INVALID, err);
    }
    *out = r;
    return WEPP_OK;
}

void wepp_reads_free(wepp_readset* r) { delete r; }

int wepp_reads_dims(const wepp_readset* r, int64_t* n_reads, int64_t* n_muts, int64_t* name_chars, int64_t* n_rev_keys,
                    int64_t* n_rev_vals, int64_t* rev_key_chars, int64_t* rev_val_chars) {
    if (!r) return abi_fail(WEPP_E_INVALID, "readset is NULL");
    if (n_reads) *n_reads = r->r.n_reads();
    if (n_muts) *n_muts = (int64_t)r->r.rm_pos.size();
    if (name_chars) *name_chars = (int64_t)r->r.name.chars.size();
    if (n_rev_keys) *n_rev_keys = (int64_t)r->r.rev_key.size();
    if (n_rev_vals) *n_rev_vals = (int64_t)r->r.rev_val.size();
    if (rev_key_chars) *rev_key_chars = (int64_t)r->r.rev_key.chars.size();
    if (rev_val_chars) *rev_val_chars = (int64_t)r->r.rev_val.chars.size();
    return WEPP_OK;
}

int wepp_reads_get(const wepp_readset* r, int32_t* start, int32_t* end, int32_t* degree, int64_t* rm_off,
                   int32_t* rm_pos, uint8_t* rm_nuc, int64_t* name_off, char* name_chars) {
    if (!r) return abi_fail(WEPP_E_INVALID, "readset is NULL");
    put(start, r->r.start); put(end, r->r.end); put(degree, r->r.degree); put(rm_off, r->r.rm_off);
    put(rm_pos, r->r.rm_pos); put(rm_nuc, r->r.rm_nuc);
    put_pool(name_off, name_chars, r->r.name);
    return WEPP_OK;
}

int wepp_reads_get_reverse(const wepp_readset* r, int64_t* key_off, char* key_chars, int64_t* rev_off, int64_t* val_off,
                           char* val_chars) {
    if (!r) return abi_fail(WEPP_E_INVALID, "readset is NULL");
    put_pool(key_off, key_chars, r->r.rev_key);
    put(rev_off, r->r.rev_off);
    put_pool(val_off, val_chars, r->r.rev_val);
    return WEPP_OK;
}

}  // extern "C"
