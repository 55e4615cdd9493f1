// state_place.cuh — placement over the DISTINCT window-restricted haplotypes of every read window: the default for
// wepp_place over the whole read set with nothing mapped and no explicit EPP lists (WEPP_STATE_PLACE=0 selects
// place_kernel, kernels.cuh, which also serves masks, EPP lists, subsets and lists that opt out here).
//
// A read's parsimony score at a node depends on the node only through the node's haplotype restricted to the
// read's window (the last event per window position on the root path).  A window of ~170 bases
// sees 12-16 k distinct restricted haplotypes on the 8 M-node bench tree while its Euler list has 61-78 k entries
// (profiles/distinct_haplotypes.py), so scoring each distinct haplotype ("state") once — weighted by the number
// of countable nodes in it — does a quarter of the work of scanning the Euler list, for the same integers
// (initial_filter.cpp:41-135 min / multiplicity, :167-177 per-node weights).
//
//   state_walk_kernel     one thread per window list walks the list in order, keeping the active net delta table
//                         per position (ENTER adds, EXIT carries the negated table, a leaf's point entries are
//                         applied, evaluated and undone): pass 0 gives every evaluated entry two independent
//                         commutative hashes of its state; pass 1 writes, for each state's representative entry,
//                         the state as K4-style entries (position, five signed bytes)
//   dedupe (host-driven)  cub radix sort of (list, hash) keys; neighbours that differ in key / second hash / size
//                         start a new state; countable nodes are summed per state
//   state_place_kernel    the tile machinery of rescore_tile_kernel: pass 1 = min and node count over the states
//                         (each evaluated with its countable-node total), pass 2 = weight / degree sums of the
//                         reads at their minimum into per-(bucket, state) accumulators
//   state_scatter_kernel  per-(bucket, entry) accumulators from the states' — expand_kernel then runs unchanged
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels.cuh"
#include "rescore_tiles.cuh"

namespace wepp {

constexpr int SW_MAX_ACTIVE = 24;          // positions with a non-zero net table at once (more: the list opts out)
constexpr uint64_t SW_NOT_EVAL = ~0ull;

__device__ __forceinline__ uint64_t sw_mix(uint64_t x) {   // splitmix64 finaliser
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

constexpr int SW_CHUNKS = 64;              // chunks a list is walked in (one warp's lane 0 each)
struct ChunkNet {
    int32_t n;                              // active positions at the chunk's end when started from an empty context
    uint32_t pos[SW_MAX_ACTIVE];
    uint64_t tab[SW_MAX_ACTIVE];
};

// chunk k of a list of n entries starts at k * ceil(n / SW_CHUNKS), moved forward to a boundary entry so that a
// leaf's point entries are never split (the same rule in every pass)
__device__ __forceinline__ int sw_chunk_begin(const Entry* e, int n, int k) {
    if (k >= SW_CHUNKS) return n;
    const int cs = (n + SW_CHUNKS - 1) / SW_CHUNKS;
    int c = min(n, k * cs);
    while (c < n && (__ldg(&e[c].x) & ENT_POINT)) ++c;
    return c;
}

struct StateWalkParams {
    const Entry* lists;
    const ListDesc* list_desc;
    int32_t n_lists;
    int32_t pass;                 // -1: chunk nets ; 0: hashes ; 1: representatives' entries ; 2: verify (WEPP_STATE_VERIFY)
    // pass 0 out (per list entry)
    uint64_t* key;                // (list << 51 | hash >> 13), SW_NOT_EVAL for entries that are not evaluated
    uint64_t* h2;                 // second hash (size in the low byte)
    int32_t* overflow;            // per list: 1 = more than SW_MAX_ACTIVE active positions (state path unusable)
    ChunkNet* nets;               // [n_lists][SW_CHUNKS]: net effect of each chunk on the context (pass -1 out)
    const ChunkNet* ctx;          // [n_lists][SW_CHUNKS]: context at each chunk's start (passes 0 / 1 in)
    // pass 1 in / out
    const int32_t* rep_state;     // per list entry: global state index it represents, or -1
    const int64_t* state_eoff;    // [S + 1]
    const int32_t* state_ucnt;    // [S]
    const int32_t* state_first;   // [n_lists + 1] first state of each list
    Entry* state_ent;
    // pass 2 in: the state index of every entry; out: entries whose state differs from their state's stored entries
    const int32_t* sid;
    unsigned long long* mismatches;
};

// five signed bytes packed as b0..b3 in z and b4 in the low byte of w
__device__ __forceinline__ uint64_t sw_pack(uint32_t z, uint32_t w) { return (uint64_t)z | ((uint64_t)(w & 0xFFu) << 32); }

__device__ __forceinline__ uint64_t sw_add_bytes(uint64_t a, uint64_t b) {   // per-byte wrapping add of 5 bytes
    const uint64_t H = 0x8080808080ull;
    return (((a & ~H) + (b & ~H)) ^ ((a ^ b) & H)) & 0xFFFFFFFFFFull;
}
__device__ __forceinline__ uint64_t sw_neg_bytes(uint64_t a) {               // per-byte negation
    return sw_add_bytes(~a & 0xFFFFFFFFFFull, 0x0101010101ull);
}

constexpr int SW_MAX_LISTS = 8190;   // 13 bits of the sort key

__global__ void iota_kernel(uint32_t* __restrict__ v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (uint32_t)i;
}

// one warp per (list, chunk), lane 0 walks: the walk is sequential and branchy, tens of thousands of independent
// walks keep the schedulers busy.  Pass -1 walks every chunk from an empty context and stores its net effect;
// passes 0 and 1 start from the sum of the earlier chunks' nets.
__global__ void state_walk_kernel(const StateWalkParams p) {
    const int wid = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int l = wid / SW_CHUNKS, ck = wid % SW_CHUNKS;
    if (l >= p.n_lists || (threadIdx.x & 31) != 0) return;
    const ListDesc ld = p.list_desc[l];
    const Entry* e = p.lists + ld.off;
    const int i_begin = sw_chunk_begin(e, ld.n, ck), i_end = sw_chunk_begin(e, ld.n, ck + 1);
    uint32_t pos[SW_MAX_ACTIVE];
    uint64_t tab[SW_MAX_ACTIVE];
    int n_act = 0;
    uint64_t H1 = 0, H2 = 0;
    bool over = false;
    auto apply = [&](uint32_t ps, uint64_t d) {
        if (d == 0ull) return;
        int k = 0;
        while (k < n_act && pos[k] != ps) ++k;
        if (k < n_act) {
            const uint64_t old = tab[k], now = sw_add_bytes(old, d);
            H1 -= sw_mix(((uint64_t)ps << 40) | old);
            H2 -= sw_mix((((uint64_t)ps << 40) | old) ^ 0x5851F42D4C957F2Dull);
            if (now == 0ull) {
                --n_act;
                pos[k] = pos[n_act];
                tab[k] = tab[n_act];
            } else {
                tab[k] = now;
                H1 += sw_mix(((uint64_t)ps << 40) | now);
                H2 += sw_mix((((uint64_t)ps << 40) | now) ^ 0x5851F42D4C957F2Dull);
            }
        } else {
            if (n_act == SW_MAX_ACTIVE) {
                over = true;
                return;
            }
            pos[n_act] = ps;
            tab[n_act] = d;
            ++n_act;
            H1 += sw_mix(((uint64_t)ps << 40) | d);
            H2 += sw_mix((((uint64_t)ps << 40) | d) ^ 0x5851F42D4C957F2Dull);
        }
    };
    if (p.pass >= 0) {   // context at the chunk's start (state_ctx_kernel)
        const ChunkNet& cn = p.ctx[(size_t)l * SW_CHUNKS + ck];
        for (int a = 0; a < cn.n; ++a) apply(cn.pos[a], cn.tab[a]);
    }
    for (int i = i_begin; i < i_end && !over; ++i) {
        const uint4 en = ld_entry(e + i);
        const uint64_t d = sw_pack(en.z, en.w);
        const uint32_t ps = en.w >> 16;
        apply(ps, d);
        if (en.x & ENT_EVAL) {
            if (p.pass == 0) {
                p.key[ld.off + i] = ((uint64_t)l << 51) | (H1 >> 13);
                p.h2[ld.off + i] = (H2 & ~0xFFull) | (uint64_t)n_act;
            } else if (p.pass == 1) {
                const int s = p.rep_state[ld.off + i];
                if (s >= 0) {
                    // the state as entries sorted by position (insertion sort of <= SW_MAX_ACTIVE items)
                    int order[SW_MAX_ACTIVE];
                    for (int a = 0; a < n_act; ++a) {
                        int b = a;
                        while (b > 0 && pos[order[b - 1]] > pos[a]) {
                            order[b] = order[b - 1];
                            --b;
                        }
                        order[b] = a;
                    }
                    const uint32_t local = (uint32_t)(s - p.state_first[l]);
                    const uint32_t uc = (uint32_t)p.state_ucnt[s];
                    Entry* out = p.state_ent + p.state_eoff[s];
                    if (n_act == 0) {
                        out[0] = Entry{local | RT_END, uc, 0u, 0u};
                    } else {
                        for (int a = 0; a < n_act; ++a) {
                            const uint64_t t = tab[order[a]];
                            Entry o;
                            o.x = local | (a + 1 == n_act ? RT_END : 0u);
                            o.y = uc;
                            o.z = (uint32_t)t;
                            o.w = (uint32_t)(t >> 32) | (pos[order[a]] << 16);
                            out[a] = o;
                        }
                    }
                }
            }
            else if (p.pass == 2) {
                // The states were identified by two 64-bit hashes (+ size): here every evaluated entry's actual state is
                // compared, position by position, with the entries stored for the state it was assigned to.
                const int s = p.sid[ld.off + i];
                bool ok = s >= 0;
                if (ok) {
                    const int64_t e0 = p.state_eoff[s], e1 = p.state_eoff[s + 1];
                    if (n_act == 0) {
                        ok = e1 - e0 == 1 && p.state_ent[e0].z == 0u && (p.state_ent[e0].w & 0xFFu) == 0u;
                    } else {
                        ok = e1 - e0 == n_act;
                        for (int a = 0; ok && a < n_act; ++a) {   // stored sorted by position: find each active position
                            bool found = false;
                            for (int64_t k = e0; k < e1; ++k) {
                                const Entry se = p.state_ent[k];
                                if ((se.w >> 16) == pos[a]) found = sw_pack(se.z, se.w) == tab[a];
                            }
                            ok = found;
                        }
                    }
                }
                if (!ok) atomicAdd(p.mismatches, 1ull);
            }
        } else if (p.pass == 0) {
            p.key[ld.off + i] = SW_NOT_EVAL;
        }
        if ((en.x & ENT_POINT) && !(en.x & ENT_SKIP)) {
            // the leaf has been evaluated: its point entries leave the context again
            const int g = (int)(en.y >> 8);
            for (int k = 0; k <= g; ++k) {
                const uint4 u = ld_entry(e + i - k);
                apply(u.w >> 16, sw_neg_bytes(sw_pack(u.z, u.w)));
            }
        }
    }
    if (p.pass == -1) {
        ChunkNet& cn = p.nets[(size_t)l * SW_CHUNKS + ck];
        cn.n = n_act;
        for (int a = 0; a < n_act; ++a) {
            cn.pos[a] = pos[a];
            cn.tab[a] = tab[a];
        }
    }
    if (over) p.overflow[l] = 1;   // zeroed by the host before pass -1; any pass may raise it
}

// context at every chunk's start = sum of the earlier chunks' nets (one thread per list: 64 small sparse adds)
__global__ void state_ctx_kernel(const ChunkNet* __restrict__ nets, int n_lists, ChunkNet* __restrict__ ctx, int32_t* __restrict__ overflow) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lists) return;
    uint32_t pos[SW_MAX_ACTIVE];
    uint64_t tab[SW_MAX_ACTIVE];
    int n_act = 0;
    for (int k = 0; k < SW_CHUNKS; ++k) {
        ChunkNet& out = ctx[(size_t)l * SW_CHUNKS + k];
        out.n = n_act;
        for (int a = 0; a < n_act; ++a) {
            out.pos[a] = pos[a];
            out.tab[a] = tab[a];
        }
        const ChunkNet& cn = nets[(size_t)l * SW_CHUNKS + k];
        for (int a = 0; a < cn.n; ++a) {
            int j = 0;
            while (j < n_act && pos[j] != cn.pos[a]) ++j;
            if (j < n_act) {
                tab[j] = sw_add_bytes(tab[j], cn.tab[a]);
                if (tab[j] == 0ull) {
                    --n_act;
                    pos[j] = pos[n_act];
                    tab[j] = tab[n_act];
                }
            } else if (n_act < SW_MAX_ACTIVE) {
                pos[n_act] = cn.pos[a];
                tab[n_act] = cn.tab[a];
                ++n_act;
            } else {
                overflow[l] = 1;
            }
        }
    }
}

// sorted (key, entry) pairs -> "starts a new state" flags
__global__ void state_flag_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ ent, const uint64_t* __restrict__ h2,
                                  int64_t n, int32_t* __restrict__ flag) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int f = 0;
    if (key[j] != SW_NOT_EVAL) f = (j == 0) || key[j] != key[j - 1] || h2[ent[j]] != h2[ent[j - 1]];
    flag[j] = f;
}

// States are renumbered in the order of their first evaluated entry (Euler order, list by list): the states first
// met inside a subtree get consecutive numbers, so the posting list of a mutation carried by a whole clade is a few
// runs of consecutive states (delta_place.cuh: conflict-free scratch banks, coalesced base scores).
__global__ void state_first_entry_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ ent, const int32_t* __restrict__ incl,
                                         int64_t n, uint32_t* __restrict__ first_entry) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || key[j] == SW_NOT_EVAL) return;
    atomicMin(&first_entry[incl[j] - 1], ent[j]);
}
__global__ void state_newid_kernel(const uint32_t* __restrict__ order, int32_t n_states, int32_t* __restrict__ newid) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_states) newid[order[k]] = k;
}
__global__ void state_renumber_kernel(const uint64_t* __restrict__ key, const int32_t* __restrict__ newid, int64_t n, int32_t* __restrict__ incl) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || key[j] == SW_NOT_EVAL) return;
    incl[j] = newid[incl[j] - 1] + 1;
}

// state index per entry, countable nodes / representative / size per state
__global__ void state_assign_kernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ ent, const int32_t* __restrict__ incl,
                                    const Entry* __restrict__ lists, const uint64_t* __restrict__ h2, int64_t n,
                                    int32_t* __restrict__ sid, int32_t* __restrict__ state_ucnt, int32_t* __restrict__ state_rep,
                                    int64_t* __restrict__ state_len, int32_t* __restrict__ state_list) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = ent[j];
    if (key[j] == SW_NOT_EVAL) {
        sid[i] = -1;
        return;
    }
    const int32_t s = incl[j] - 1;
    sid[i] = s;
    const uint32_t x = lists[i].x, y = lists[i].y;
    const int32_t uc = (x & ENT_POINT) ? (int32_t)(y & 0xFFu) : (int32_t)y;
    if (uc) atomicAdd(&state_ucnt[s], uc);
    atomicMin(&state_rep[s], (int32_t)i);
    if (j == 0 || incl[j - 1] != incl[j]) {   // first of its state in sorted order
        const int len = (int)(h2[i] & 0xFFull);
        state_len[s] = len > 0 ? len : 1;
        state_list[s] = (int32_t)(key[j] >> 51);
    }
}

__global__ void state_first_kernel(const int32_t* __restrict__ state_list, int32_t n_states, int32_t n_lists, int32_t* __restrict__ first) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_states) return;
    const int cur = s < n_states ? state_list[s] : n_lists;
    const int prv = s > 0 ? state_list[s - 1] : -1;
    for (int l = prv + 1; l <= cur; ++l) first[l] = s;   // lists without a state get an empty range
}

__global__ void state_rep_mark_kernel(const int32_t* __restrict__ state_rep, int32_t n_states, int32_t* __restrict__ rep_state) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_states) rep_state[state_rep[s]] = s;
}

struct StatePlaceParams {
    const Entry* state_ent;
    const int64_t* state_eoff;
    const int32_t* state_first;   // [n_lists + 1]
    const int64_t* sacc_off;      // per bucket: first accumulator
    const ListDesc* list_desc;
    const BucketDesc* buckets;
    const TileDesc* tiles;
    const int32_t* start;
    const int32_t* end;
    const int32_t* degree;
    const int64_t* rm_off;
    const int32_t* rm_pos;
    const uint8_t* rm_code;
    const int64_t* perm;
    int32_t* max_pars;
    int32_t* mult;
    double* saccS;
    int32_t* saccC;
};

// One CTA per read tile; its warps split the list's states.  Pass 1: min and countable-node count; pass 2: the
// weights of the reads at their minimum, summed over the warp and added to the (bucket, state) accumulators.
template <int K>
__global__ void __launch_bounds__(PLACE_WARPS * 32, 2) state_place_kernel(const StatePlaceParams p) {
    using ST = typename Sel<K>::type;
    constexpr int P = K / 2;
    constexpr int SHIFT = Sel<K>::SHIFT;
    constexpr int T = 32 * K;
    constexpr int TBL = RtLayout<K>::CODES;            // pass-2 pattern tables PatEntry[2][16][32] (as place_kernel)
    constexpr int CODES = TBL + 2 * TBL_HALF;          // the selector table follows them
    static_assert(CODES <= 65536, "pattern tables must be addressable with 16-bit offsets");
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t ebuf_s = smem_s + warp * 512;
    int* xch = reinterpret_cast<int*>(smem + RT_XCH);
    PatEntry* tbl = reinterpret_cast<PatEntry*>(smem + TBL);
    unsigned char* codes = smem + CODES;
    const uint32_t col = smem_s + CODES + lane * K;
    const unsigned FULL = 0xFFFFFFFFu;
    if (smem_s + CODES > 0xFFFFu) __trap();
    const uint32_t tbase2 = (smem_s + TBL + lane * 16) | ((smem_s + TBL + TBL_HALF + lane * 16) << 16);

    const TileDesc td = p.tiles[blockIdx.x];
    const BucketDesc bd = p.buckets[td.bucket];
    const ListDesc ld = p.list_desc[bd.list];

    int k_non_n[K];
    int64_t rid[K];
    {
        int s_rel[K], e_rel[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int ti = lane * K + j;
            const bool valid = ti < td.count;
            rid[j] = valid ? p.perm[td.first + ti] : -1;
            s_rel[j] = valid ? p.start[rid[j]] - ld.b0 : 1;
            e_rel[j] = valid ? p.end[rid[j]] - ld.b0 : 0;
            k_non_n[j] = 0;
        }
        for (int pos = warp; pos < ld.width; pos += PLACE_WARPS) {
            uint32_t w[2] = {0u, 0u};
#pragma unroll
            for (int j = 0; j < K; ++j)
                w[j >> 2] |= ((pos >= s_rel[j] && pos <= e_rel[j]) ? 0x80u : 0xD5u) << (8 * (j & 3));
            ST* row = reinterpret_cast<ST*>(codes) + pos * 32 + lane;
            if constexpr (K == 8) *row = make_uint2(w[0], w[1]);
            else *row = (ST)w[0];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (rid[j] >= 0) {
            const int64_t a = p.rm_off[rid[j]], b = p.rm_off[rid[j] + 1];
            for (int64_t k = a; k < b; ++k) {
                const uint32_t c = p.rm_code[k];
                if (warp == 0) {
                    const int pr = p.rm_pos[k] - ld.b0;
                    codes[(pr * 32 + lane) * K + j] = (unsigned char)(c | ((c | 8u) << 4));
                }
                k_non_n[j] += (c <= 4u);   // seed set: non-N mutations (initial_filter.cpp:118-123)
            }
        }
    }
    __syncthreads();

    const int s_lo = p.state_first[bd.list], s_n = p.state_first[bd.list + 1] - s_lo;
    const int w0 = (int)((int64_t)warp * s_n / PLACE_WARPS), w1 = (int)((int64_t)(warp + 1) * s_n / PLACE_WARPS);
    const int64_t c0 = p.state_eoff[s_lo + w0], c1 = p.state_eoff[s_lo + w1];
    const Entry* ent = p.state_ent;

    uint32_t S0[P];
#pragma unroll
    for (int q = 0; q < P; ++q) S0[q] = S_BIAS2 + ((uint32_t)k_non_n[2 * q] | ((uint32_t)k_non_n[2 * q + 1] << 16));

    // ---- pass 1 -----------------------------------------------------------------------------------------
    Pass1<K> st;
    st.bsum = (uint32_t)P * BEST_NONE2;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        st.S[q] = S0[q];
        st.B[q] = st.oB[q] = BEST_NONE2;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) st.cnt[j] = 0;
    {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (c0 + lane < c1) nxt = ld_entry(ent + c0 + lane);
        for (int64_t base = c0; base < c1; base += 32) {
            sts128(ebuf_s + lane * 16, nxt);
            const uint32_t em = __ballot_sync(FULL, (nxt.x & RT_END) != 0u);
            __syncwarp();
            nxt = make_uint4(0, 0, 0, 0);
            if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                const uint32_t ea = ebuf_s + g * 128;
                const uint32_t eg = em >> (8 * g);
                uint4 e[8];
                uint32_t sel[8][P];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = lds128(ea + i * 16);
#pragma unroll
                for (int i = 0; i < 8; ++i) load_sel<K>(col, e[i].w >> SHIFT, sel[i]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int q = 0; q < P; ++q) st.S[q] = __vadd2(st.S[q], prmt(e[i].z, e[i].w, sel[i][q]));
                    if (eg & (1u << i)) {
                        st.eval(st.S, (int)e[i].y);
#pragma unroll
                        for (int q = 0; q < P; ++q) st.S[q] = S0[q];
                    }
                }
            }
            __syncwarp();
        }
    }
    // ---- fold the warps' (min, count) ---------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const uint32_t bl = st.B[q] & 0xFFFFu, bh = st.B[q] >> 16;
        xch[(warp * 2 + 0) * T + (2 * q) * 32 + lane] = bl == BEST_NONE ? 0x3FFFFFFF : (int)bl - (int)S_BIAS;
        xch[(warp * 2 + 0) * T + (2 * q + 1) * 32 + lane] = bh == BEST_NONE ? 0x3FFFFFFF : (int)bh - (int)S_BIAS;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) xch[(warp * 2 + 1) * T + j * 32 + lane] = st.cnt[j];
    __syncthreads();
    double wgt[K];
    int deg[K];
    uint32_t bestp[P];
#pragma unroll
    for (int q = 0; q < P; ++q) bestp[q] = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        wgt[j] = 0.0;
        deg[j] = 0;
        uint32_t b = 0xFFFFu;
        if (rid[j] >= 0) {
            int gb = 0x3FFFFFFF;
#pragma unroll
            for (int w = 0; w < PLACE_WARPS; ++w) gb = min(gb, xch[(w * 2 + 0) * T + j * 32 + lane]);
            int gc = 0;
#pragma unroll
            for (int w = 0; w < PLACE_WARPS; ++w)
                if (xch[(w * 2 + 0) * T + j * 32 + lane] == gb) gc += xch[(w * 2 + 1) * T + j * 32 + lane];
            if (gb == 0x3FFFFFFF) gb = 0;   // no state at all (cannot happen: every list has its dummy entry)
            const int d = p.degree[rid[j]];
            if (gc > 0) {
                wgt[j] = (double)d / ((double)(1 + gb) * (double)gc);   // node_score, initial_filter.hpp:54-57
                deg[j] = d;
                b = (uint32_t)gb + S_BIAS;
            }
            if (warp == 0) {
                p.max_pars[rid[j]] = gb;
                p.mult[rid[j]] = gc;
            }
        }
        if (j & 1) bestp[j >> 1] = (bestp[j >> 1] & 0x0000FFFFu) | (b << 16);
        else bestp[j >> 1] = (bestp[j >> 1] & 0xFFFF0000u) | b;
    }
    // pattern tables: for the lane's even reads (low halves) and odd reads (high halves), the weight / degree sums
    // of the reads whose NOT-at-min bit is clear, indexed by the P-bit pattern; the warps split the 2 x 2^P patterns
    for (int hp = warp; hp < 2 * (1 << P); hp += PLACE_WARPS) {
        const int hh = hp >> P, pat = hp & ((1 << P) - 1);
        double ws = 0.0;
        int ds = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            if (!((pat >> q) & 1)) {
                ws += hh ? wgt[2 * q + 1] : wgt[2 * q];
                ds += hh ? deg[2 * q + 1] : deg[2 * q];
            }
        }
        PatEntry pe;
        pe.w = ws;
        pe.c = ds;
        pe.pad = 0;
        tbl[(hh * 16 + pat) * 32 + lane] = pe;
    }
    __syncthreads();
    // ---- pass 2 -----------------------------------------------------------------------------------------
    double* accS = p.saccS + p.sacc_off[td.bucket];
    int32_t* accC = p.saccC + p.sacc_off[td.bucket];
    uint32_t S[P];
#pragma unroll
    for (int q = 0; q < P; ++q) S[q] = S0[q];
    {
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (c0 + lane < c1) nxt = ld_entry(ent + c0 + lane);
        for (int64_t base = c0; base < c1; base += 32) {
            sts128(ebuf_s + lane * 16, nxt);
            const uint32_t em = __ballot_sync(FULL, (nxt.x & RT_END) != 0u);
            __syncwarp();
            nxt = make_uint4(0, 0, 0, 0);
            if (base + 32 + lane < c1) nxt = ld_entry(ent + base + 32 + lane);
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                const uint32_t ea = ebuf_s + g * 128;
                const uint32_t eg = em >> (8 * g);
#pragma unroll 2
                for (int i = 0; i < 8; ++i) {
                    const uint4 e = lds128(ea + i * 16);
                    uint32_t sel[P];
                    load_sel<K>(col, e.w >> SHIFT, sel);
#pragma unroll
                    for (int q = 0; q < P; ++q) S[q] = __vadd2(S[q], prmt(e.z, e.w, sel[q]));
                    if (eg & (1u << i)) {
                        uint32_t ti = tbase2;
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            // per half: 0 = the read is at its minimum here, 1 = it is not
                            ti += __vminu2(S[q] ^ bestp[q], 0x00010001u) * (uint32_t)(TBL_PAT_STRIDE << q);
                            S[q] = S0[q];
                        }
                        const uint4 pa = lds128(ti & 0xFFFFu), pb = lds128(ti >> 16);
                        double ws = __hiloint2double((int)pa.y, (int)pa.x) + __hiloint2double((int)pb.y, (int)pb.x);
                        // degrees first (one REDUX): a zero total means no read of the tile is at its minimum here
                        // (a read's weight is non-zero only with its degree)
                        const int ds = __reduce_add_sync(FULL, (int)(pa.z + pb.z));
                        if (ds != 0) {
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) ws += __shfl_xor_sync(FULL, ws, o);
                            if (lane == 0) {
                                const uint32_t s = e.x & ~RT_END;
                                if (ws != 0.0) atomicAdd(accS + s, ws);
                                atomicAdd(accC + s, ds);
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// per-(bucket, entry) accumulators from the (bucket, state) ones; entries that are not evaluated hold 0
__global__ void state_scatter_kernel(const ListDesc* __restrict__ list_desc, const BucketDesc* __restrict__ buckets,
                                     const int32_t* __restrict__ sid, const int32_t* __restrict__ state_first,
                                     const int64_t* __restrict__ sacc_off, const double* __restrict__ saccS,
                                     const int32_t* __restrict__ saccC, double* __restrict__ accS, int32_t* __restrict__ accC) {
    const BucketDesc bd = buckets[blockIdx.y];
    const ListDesc ld = list_desc[bd.list];
    const int32_t first = state_first[bd.list];
    const int64_t so = sacc_off[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const int32_t s = sid[ld.off + i];
        accS[bd.acc_off + i] = s >= 0 ? saccS[so + (s - first)] : 0.0;
        accC[bd.acc_off + i] = s >= 0 ? saccC[so + (s - first)] : 0;
    }
}

// expand_kernel (kernels.cuh) straight from the per-(bucket, state) weight sums, score only: what a list entry holds is
// its state's sum, the step at an entry is that minus the enclosing boundary entry's.  The per-step node pass of the
// peak loop (no counts, no per-entry copy of the accumulators).
__global__ void expand_states_score_kernel(const Entry* __restrict__ lists, const ListDesc* __restrict__ list_desc,
                                           const BucketDesc* __restrict__ buckets, const int32_t* __restrict__ prev_boundary,
                                           const int32_t* __restrict__ sid, const int32_t* __restrict__ state_first,
                                           const int64_t* __restrict__ sacc_off, const double* __restrict__ saccS,
                                           unsigned long long* __restrict__ diff_lo, unsigned long long* __restrict__ diff_hi) {
    const BucketDesc bd = buckets[blockIdx.y];
    const ListDesc ld = list_desc[bd.list];
    const Entry* e = lists + ld.off;
    const int32_t* pb = prev_boundary + ld.off;
    const int32_t* sd = sid + ld.off;
    const double* s = saccS + sacc_off[blockIdx.y] - state_first[bd.list];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ld.n; i += gridDim.x * blockDim.x) {
        const uint32_t x = __ldg(&e[i].x);
        if (x & ENT_SKIP) continue;
        const int pi = pb[i];
        const int32_t si = sd[i], sp = pi >= 0 ? sd[pi] : -1;
        if (si == sp) continue;
        const double cur = si >= 0 ? s[si] : 0.0, prv = sp >= 0 ? s[sp] : 0.0;
        if (cur == prv) continue;
        const uint32_t idx = x & IDX_MASK;
        unsigned long long alo, blo;
        long long ahi, bhi;
        dbl_to_fix(cur, alo, ahi);
        dbl_to_fix(prv, blo, bhi);
        const unsigned long long lo = alo - blo;
        const long long hi = ahi - bhi - (alo < blo ? 1 : 0);
        atomic_add128(diff_lo + idx, diff_hi + idx, lo, hi);
        if (x & ENT_POINT) {   // back to the enclosing value right after the leaf
            const unsigned long long nlo = 0ull - lo;
            const long long nhi = ~hi + (lo == 0ull ? 1 : 0);
            atomic_add128(diff_lo + idx + 1, diff_hi + idx + 1, nlo, nhi);
        }
    }
}

}  // namespace wepp
