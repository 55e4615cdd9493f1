/* wepp_b200.h — C ABI of the B200-native WEPP read-placement library (libwepp_b200.so).
 *
 * This is the drop-in boundary for the reference's placement hot path.  Citations are
 * file:line relative to the TurakhiaLab/WEPP tree (commit 6177aced):
 *
 *   wepp_set_arena        <- the flattened condensed tree the reference's `arena` holds:
 *                            std::vector<haplotype> nodes, src/WEPP/arena.cpp:3-56
 *                            (haplotype::parent / ::muts, src/WEPP/haplotype.hpp:10-43)
 *   wepp_set_reads        <- std::vector<raw_read>, src/WEPP/read.hpp:6-12, after masking
 *                            (src/WEPP/arena.hpp:62-72)
 *   wepp_set_mapped       <- haplotype::mapped, src/WEPP/haplotype.hpp:36 (read at
 *                            src/WEPP/initial_filter.cpp:130)
 *   wepp_place            <- wepp_filter::cartesian_map, src/WEPP/initial_filter.cpp:139-211
 *                            (which calls single_read_tree, :41-135, for every read)
 *   wepp_place_subset     <- single_read_tree on chosen reads under the current `mapped`
 *                            mask, as wepp_filter::remove_read does, :298-302
 *   wepp_filter_peaks     <- wepp_filter::filter, src/WEPP/initial_filter.cpp:455-506 (the peak loop
 *                            :284-453 and the neighbour expansion :476-503)
 *   wepp_rescore          <- haplotype::mutation_distance(const raw_read&),
 *                            src/WEPP/haplotype.hpp:123-177, over a candidate set with the
 *                            min / argmin idiom of src/WEPP/arena.cpp:614-625 and :846-857
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer, the
 * library owns all device memory; every function returns 0 on success and a negative
 * WEPP_E_* code on failure, with a message available from wepp_last_error() (the
 * reference itself reports errors with fprintf(stderr)+exit(1), e.g. src/WEPP/sam2pb.cpp:497-500).
 * One handle drives one GPU and is used from one host thread at a time.  There is no CPU
 * fallback: without a usable CUDA device wepp_create fails.
 *
 * Nucleotide codes are the reference's 4-bit one-hot / IUPAC ids
 * (src/mutation_annotated_tree.cpp:19-74): A=1 C=2 G=4 T=8, unions for ambiguity, N=15.
 * Tree mutations may carry any code 1..15 in mut_nuc; read alleles must be one of
 * 1,2,4,8,15 (what sam2PB emits, src/WEPP/sam2pb.cpp:245-249) — anything else is
 * rejected with WEPP_E_INVALID.
 */
#ifndef WEPP_B200_H
#define WEPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WEPP_NUM_RANGE_BINS 50      /* src/WEPP/config.hpp:13 */
#define WEPP_MAX_CACHED_EPP 2048    /* src/WEPP/config.hpp:9  */

#define WEPP_OK           0
#define WEPP_E_INVALID   -1   /* bad argument / malformed input            */
#define WEPP_E_CUDA      -2   /* CUDA runtime error (message has details)  */
#define WEPP_E_STATE     -3   /* call order violated (e.g. place before set_reads) */
#define WEPP_E_CAPACITY  -4   /* caller-provided buffer too small          */

typedef struct wepp_handle wepp_handle;

/* Life cycle.  `device` is a CUDA ordinal.  */
int         wepp_create(int device, wepp_handle** out);
void        wepp_destroy(wepp_handle* h);
const char* wepp_last_error(void);
int         wepp_abi_version(void);

/* Stream control.  By default the handle owns a private non-blocking stream.  wepp_set_stream
 * makes every later launch and copy use the caller's cudaStream_t (passed as void*), so the
 * caller can order its own work (NCCL collectives, CUDA events) against the placement.
 * wepp_place / wepp_place_subset only enqueue work; wepp_sync (or any wepp_get_*) waits.  */
int wepp_set_stream(wepp_handle* h, void* cuda_stream);
int wepp_sync(wepp_handle* h);

/* Tunables (optional, before wepp_set_arena): stripe width in bases used to bucket read
 * windows (default 16) and reads per warp lane K in {2,4,8} (0 = choose per bucket). */
int wepp_set_options(wepp_handle* h, int32_t stripe_width, int32_t reads_per_lane);

/* The flattened tree.  Node v is the v-th haplotype in preorder (arena index):
 * parent[0] = -1 and parent[v] < v.  Node v's mutations are
 * mut_pos/mut_ref/mut_nuc[mut_off[v] .. mut_off[v+1]) with 1 <= pos <= genome_size and at
 * most one mutation per position per node.  Builds the device-resident Euler-tour event
 * stripes.  */
int wepp_set_arena(wepp_handle* h, int32_t n_nodes, const int32_t* parent, const int64_t* mut_off,
                   const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size);

/* The collapsed reads.  Window [start,end] is 1-based and closed; read r's mutation list
 * is rm_pos/rm_nuc[rm_off[r] .. rm_off[r+1]), sorted by position, inside the window.
 * Packs the reads, groups them into window buckets and builds the per-bucket Euler lists. */
int wepp_set_reads(wepp_handle* h, int64_t n_reads, const int32_t* start, const int32_t* end,
                   const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc);

/* haplotype::mapped for every node (NULL = none mapped).  */
int wepp_set_mapped(wepp_handle* h, const uint8_t* mapped);

/* cartesian_map over all reads.  Results stay on the device until fetched.
 * epp_cap: reads whose multiplicity is <= epp_cap get their sorted EPP node list stored
 * (reference: MAX_CACHED_EPP_SIZE); epp_capacity: total int32 slots reserved for them
 * (lists that do not fit are dropped and reported as uncached).  */
int wepp_place(wepp_handle* h, int32_t epp_cap, int64_t epp_capacity);

/* single_read_tree for a subset of reads (indices into the set_reads order) under the
 * current mapped mask; per-node accumulators are NOT touched.  Per-read results are
 * written for those reads only.  */
int wepp_place_subset(wepp_handle* h, int64_t n_sel, const int64_t* read_idx, int32_t epp_cap, int64_t epp_capacity);

/* Fetch results of the last place call (any pointer may be NULL to skip it).
 * max_parsimony[R], multiplicity[R]; score[N] (double), counts[N*50] (int32, row = node).
 * EPP lists: epp_off[R+1] CSR into epp_nodes (sorted arena indices); reads that were not
 * cached have an empty range; *n_epp receives the number of slots used.  */
int wepp_get_read_results(wepp_handle* h, int32_t* max_parsimony, int32_t* multiplicity);
int wepp_get_node_results(wepp_handle* h, double* score, int32_t* counts);
int wepp_get_epp(wepp_handle* h, int64_t* epp_off, int32_t* epp_nodes, int64_t capacity, int64_t* n_epp);

/* What the later stages actually read from the per-node state: score[N] and dist_divergence[N]
 * (src/WEPP/initial_filter.cpp:214-231 — the share of the 50 read-count bins in which the node
 * collects more than READ_DIST_FACTOR_THRESHOLD = 0.5 % of the sample's degree-weighted reads,
 * arena::read_counts(), src/WEPP/arena.cpp:138-151).  Computed on the device from the counts
 * matrix, so the 200-byte-per-node matrix itself need not cross PCIe (mapped_read_counts has no
 * reader after cartesian_map in the reference).  Either pointer may be NULL.  */
int wepp_get_node_summary(wepp_handle* h, double* score, double* dist_divergence);

/* The whole reference call in one go with host buffers in and out
 * (set_reads + set_mapped + place + get_*): what the cgo/ctypes/C++ binding calls.  */
int wepp_cartesian_map(wepp_handle* h, int64_t n_reads, const int32_t* start, const int32_t* end,
                       const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc,
                       const uint8_t* mapped, int32_t* max_parsimony, int32_t* multiplicity, double* score,
                       int32_t* counts);

/* Candidate re-scoring (K4).  Candidates are arena indices; for every read the minimum
 * mutation distance over the candidates and, optionally, the dense distance matrix
 * dist[R*n_cand] and the argmin sets as CSR (am_off[R+1], am_idx = positions in the
 * candidate list, in candidate order; capacity am_capacity).  */
int wepp_rescore(wepp_handle* h, int32_t n_cand, const int32_t* cand_nodes, int32_t* min_dist, int32_t* dist,
                 int64_t* am_off, int32_t* am_idx, int64_t am_capacity);

/* The same over reads handed in directly (not the set_reads ones): the result writers score copies of
 * the reads whose residual alleles were masked to N (arena::resolve_unaccounted_mutations,
 * src/WEPP/arena.cpp:736-785, :846-857).  Needs only wepp_set_arena.  */
int wepp_rescore_reads(wepp_handle* h, int64_t n_reads, const int32_t* start, const int32_t* end, const int64_t* rm_off,
                       const int32_t* rm_pos, const uint8_t* rm_nuc, int32_t n_cand, const int32_t* cand_nodes,
                       int32_t* min_dist, int32_t* dist, int64_t* am_off, int32_t* am_idx, int64_t am_capacity);

/* Multi-GPU exchange step (one process per GPU, reads sharded, tree replicated): the GPU form of the
 * reference's chunk merge of the per-thread node arrays (src/WEPP/initial_filter.cpp:199-211) fused with the
 * dist_divergence evaluation that follows it (:214-231), over NVLink peer memory instead of a collective.
 *   wepp_peer_export  after wepp_set_reads: this rank's blob (CUDA IPC handles of its per-node arrays + its
 *                     degree-weighted reads per count bin, arena::read_counts, src/WEPP/arena.cpp:138-151)
 *   wepp_peer_open    all ranks' blobs, in rank order (the caller moves them: torch.distributed / MPI / files)
 *   wepp_peer_merge   enqueued after wepp_place on the handle's stream.  Each rank sums, for its slice of the
 *                     nodes, the counts and scores of all ranks straight from their HBM, evaluates
 *                     dist_divergence and stores the merged score / dist_divergence of the slice into every
 *                     rank; wepp_get_node_summary then returns the merged arrays.  The caller brackets it with
 *                     stream-ordered barriers across the ranks (all placements done before; all merges done
 *                     before the next wepp_place).  The merged mapped_read_counts stay distributed (each rank
 *                     holds its slice); use an all-reduce of WEPP_BUF_COUNTS if every rank needs all of them.
 * At most 8 ranks (one NVSwitch domain).  Re-export after wepp_set_arena or when the read set changes. */
#define WEPP_PEER_BLOB_BYTES 664
int wepp_peer_export(wepp_handle* h, void* blob);
int wepp_peer_open(wepp_handle* h, int32_t rank, int32_t world, const void* blobs);
int wepp_peer_merge(wepp_handle* h);
int wepp_peer_close(wepp_handle* h);

/* The whole initial filter: wepp_filter::filter (src/WEPP/initial_filter.cpp:455-506) — cartesian_map over
 * the current reads with nothing mapped, then the greedy peak loop (step / clear_neighbors / singular_step /
 * find_correspondents / remove_read, :241-453: pick the top full_score = score * sqrt(dist_divergence) nodes
 * (ties within SCORE_EPSILON by leaf_count, then id), map their radius-MAX_PEAK_PEAK_MUTATION neighbourhoods,
 * take the reads they explain out of every node's score, until MAX_PEAKS peaks or no reads / scores are left)
 * and the neighbour expansion (:476-503).  The per-read work runs on the GPU: correspondents are found with the
 * mutation-distance kernel, their weights are re-accumulated by the placement kernel and subtracted.
 * leaf_count[N] is haplotype::leaf_count, id_rank[N] the rank of haplotype::id in ascending std::string order
 * (the comparator's last tie-break, src/WEPP/arena.hpp:16-31).  out_nodes receives the peaks (ascending arena
 * index) followed by the chosen neighbours (ascending), as the reference returns them; *n_peaks / *n_out their
 * numbers.  On return the per-node score / counts are those of the cartesian_map (recover_haplotype_state).  */
int wepp_filter_peaks(wepp_handle* h, const int32_t* leaf_count, const int32_t* id_rank, int32_t* out_nodes,
                      int32_t capacity, int32_t* n_peaks, int32_t* n_out);

/* Device-resident views for callers that keep results on the GPU (NCCL all-reduce of the
 * per-node arrays in the multi-GPU driver).  Pointers stay valid until the next
 * set_arena / destroy.  */
#define WEPP_BUF_SCORE      1   /* double[N]            */
#define WEPP_BUF_COUNTS     2   /* int32[N*50]          */
#define WEPP_BUF_MAX_PARS   3   /* int32[R], read order */
#define WEPP_BUF_MULT       4   /* int32[R], read order */
int wepp_device_buffer(wepp_handle* h, int32_t which, void** dev_ptr, int64_t* n_bytes);

/* ---- Read-sharded ranks (one process per GPU): one plan, one exchange of accumulators ------------------
 * The reference merges its threads' dense per-node arrays under a mutex (src/WEPP/initial_filter.cpp:199-211).
 * Here ranks hold disjoint read shards and the same tree, and the caller lends the library ONE collective — an
 * in-place sum all-reduce over the ranks, stream-ordered on `cuda_stream` (ncclAllReduce in a C++ driver;
 * torch.distributed.all_reduce in wepp_b200/multigpu.py).  With the hook set,
 *   wepp_set_reads  sums the (window, bin) cell histogram and the true read counts over the ranks before the plan is
 *                   derived, so every rank builds the SAME window lists, buckets and distinct states (only the bucket
 *                   sizes and tiles are the rank's own);
 *   wepp_place      sums the per-(bucket, state) accumulators (~100 MB at 8 M nodes; per-(bucket, entry) ones when the
 *                   states opt out) over the ranks — instead of the N x 208 B per-node arrays — and then finishes the
 *                   per-node score / read counts / divergence from the merged accumulators on every rank.
 * Per-read results stay on the rank that owns the read.  All ranks must issue the same calls in the same order.
 * The hook returns 0 on success.  Pass fn = NULL to go back to single-rank behaviour.  */
#define WEPP_DTYPE_I32 0
#define WEPP_DTYPE_I64 1
#define WEPP_DTYPE_F64 2
typedef int (*wepp_allreduce_fn)(void* user, void* dev_ptr, int64_t count, int32_t dtype, void* cuda_stream);
int wepp_set_allreduce(wepp_handle* h, wepp_allreduce_fn fn, void* user);

/* ---- Several GPUs of one box from ONE process (the product driver's multi-GPU mode, WEPP_GPUS) -----------
 * Replaces the chunk merge of the reference's TBB read decomposition (src/WEPP/initial_filter.cpp:152,
 * :199-211) across devices: a group owns one handle and one host thread per rank, deals the reads round-robin
 * (rank r holds the caller's reads r, r + G, ...), replicates the tree and serves the exchanges of
 * wepp_set_allreduce itself — peer_allreduce_kernel reads and writes the other ranks' buffers over NVLink peer
 * memory (rank r sums slice r of every buffer in rank order and stores it to every rank: bit-identical results
 * on all ranks), ordered by CUDA events between the ranks' streams.  No collective library is involved.
 * `devices` may be NULL (ranks on devices 0..n-1) and may name a device more than once (ranks sharing a GPU:
 * how the single-GPU test box covers this path).
 *   wepp_group_set_arena / set_reads / place   = wepp_set_arena / wepp_set_reads (dealt) / wepp_place(h, 0, 0) on
 *     every rank at once; afterwards every rank holds the merged per-node results (wepp_get_node_results on
 *     wepp_group_handle(g, 0)), wepp_group_get_read_results gathers the per-read ones in the caller's order.
 *   wepp_group_filter_peaks = wepp_filter_peaks with the reads sharded: the cartesian_map exchanges the
 *     per-(bucket, state) accumulators, every step of the peak loop the number of reads removed and the removed
 *     reads' per-node weights; the ranks take the same decisions on identical merged scores.
 *   wepp_group_run calls fn(rank, handle, user) on the ranks' threads concurrently (anything else that must run
 *     in step); wepp_group_take removes a rank's handle from the group (hook cleared, caller owns it).  */
typedef struct wepp_group wepp_group;
typedef int (*wepp_group_fn)(int32_t rank, wepp_handle* h, void* user);
int  wepp_group_create(int32_t n_ranks, const int32_t* devices, wepp_group** out);
void wepp_group_destroy(wepp_group* g);
int32_t wepp_group_size(const wepp_group* g);
wepp_handle* wepp_group_handle(wepp_group* g, int32_t rank);
wepp_handle* wepp_group_take(wepp_group* g, int32_t rank);
int  wepp_group_run(wepp_group* g, wepp_group_fn fn, void* user);
int  wepp_group_set_arena(wepp_group* g, int32_t n_nodes, const int32_t* parent, const int64_t* mut_off,
                          const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size);
int  wepp_group_set_reads(wepp_group* g, int64_t n_reads, const int32_t* start, const int32_t* end,
                          const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc);
int  wepp_group_place(wepp_group* g);
int  wepp_group_get_read_results(wepp_group* g, int32_t* max_parsimony, int32_t* multiplicity);
int  wepp_group_filter_peaks(wepp_group* g, const int32_t* leaf_count, const int32_t* id_rank, int32_t* out_nodes,
                             int32_t capacity, int32_t* n_peaks, int32_t* n_out);

/* Introspection for benchmarks: numbers describing the last wepp_place.  */
typedef struct wepp_stats {
    int64_t n_nodes, n_events, n_euler_entries;     /* tree */
    int64_t n_reads, n_buckets, n_lists, n_tiles;   /* read bucketing */
    int64_t list_entries_total;                     /* sum of per-list entries */
    int64_t scanned_entries;                        /* sum over tiles of list length (one pass) */
    int64_t scanned_read_entries;                   /* sum over reads of their list length */
    int64_t algorithmic_bytes;                      /* DESIGN.md §roofline definition, per place */
    int64_t kernel_launches;                        /* kernels launched by the last place */
    float   ms_place_total;                         /* CUDA events, whole place */
    float   ms_scan_kernel;                         /* the dominant placement kernel */
    float   ms_node_kernels;                        /* expand + prefix scans */
    int32_t reads_per_tile;
    int32_t stripe_width;
    int32_t place_path;                             /* 0 Euler-list scan, 1 distinct states, 2 sparse corrections over the states */
    int32_t n_states;                               /* distinct window-restricted haplotypes (paths 1, 2) */
    int32_t n_window_groups;                        /* (bucket, window) groups of the read set (path 2) */
    float   ms_exchange;                            /* read-sharded ranks: the accumulator all-reduce (not in ms_node_kernels) */
} wepp_stats;
int wepp_get_stats(wepp_handle* h, wepp_stats* out);

/* The arena builder (host side; no GPU needed).  Replaces the reference's arena constructor,
 * src/WEPP/arena.hpp:56-79: masks read mutations at the masked sites (:62-72), derives the covered
 * sites (site_read_map, :157-175), condenses the MAT to nodes that mutate a covered site
 * (create_condensed_tree, src/WEPP/util.cpp:79-133) and flattens it in preorder (arena::from_mat,
 * src/WEPP/arena.cpp:3-56).  Input tree: node 0 is the root, parent[v] < v, a node's children are
 * taken in index order (the order the MAT loader created them), mutations sorted by position.
 * Outputs (sizes from wepp_arena_dims): per arena node its parent, source MAT node, leaf_count,
 * mutation CSR, and the CSR of MAT nodes folded into it (first = the source;
 * condensed_node_mappings); and the masked reads' mutation CSR.  */
typedef struct wepp_arena wepp_arena;
int  wepp_arena_build(int32_t n_mat_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                      const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size, int32_t n_masked,
                      const int32_t* masked, int64_t n_reads, const int32_t* start, const int32_t* end,
                      const int64_t* rm_off, const int32_t* rm_pos, const uint8_t* rm_nuc, wepp_arena** out);
void wepp_arena_free(wepp_arena* a);
int  wepp_arena_dims(const wepp_arena* a, int32_t* n_nodes, int64_t* n_muts, int64_t* n_read_muts, int64_t* n_mapped);
int  wepp_arena_get(const wepp_arena* a, int32_t* parent, int32_t* source, int32_t* leaf_count, int64_t* mut_off,
                    int32_t* mut_pos, uint8_t* mut_ref, uint8_t* mut_nuc, int64_t* map_off, int32_t* map_nodes);
int  wepp_arena_get_reads(const wepp_arena* a, int64_t* rm_off, int32_t* rm_pos, uint8_t* rm_nuc);
int  wepp_set_arena_from(wepp_handle* h, const wepp_arena* a);

/* ---- File formats either side of the path (host side; no GPU needed) ---------------------------
 * The MAT loader replaces MAT::load_mutation_annotated_tree + Tree::uncondense_leaves as dataset::mat()
 * calls them (src/WEPP/dataset.hpp:213-220; src/mutation_annotated_tree.cpp:415-508 Newick, :522-612
 * protobuf `Parsimony::data` with gzip detected by a ".gz" in the file name, :720-746 add_mutation,
 * :1224-1272 uncondense).  Nodes come back in creation order (Newick preorder, then un-condensed
 * leaves): parent[v] < v, children of a node = its nodes in index order — exactly the input
 * wepp_arena_build takes.  Mutations keep the loader's 4-bit ids; a negative position is a masked
 * mutation.  Strings are returned as one character pool plus offsets (n + 1 entries).
 * The Newick token scan and the per-node mutation fill (the reference's tbb::parallel_for, :556-596) run
 * on WEPP_THREADS host threads (default: all).  wepp_mat_load keeps a flattened-tree sidecar
 * "<path>.wepp_flat" (or in $WEPP_SIDECAR_DIR): the parsed tree as flat arrays, keyed by the source's
 * size and a 64-bit hash of its bytes, read back on the next load of the same file instead of
 * inflating and parsing it; WEPP_SIDECAR=0 turns it off, an unwritable directory is not an error.  */
typedef struct wepp_mat wepp_mat;
int  wepp_mat_load(const char* path, int32_t uncondense, wepp_mat** out);
int  wepp_mat_parse(const void* pb_bytes, int64_t n_bytes, int32_t uncondense, wepp_mat** out);
void wepp_mat_free(wepp_mat* m);
int  wepp_mat_dims(const wepp_mat* m, int32_t* n_nodes, int64_t* n_muts, int32_t* n_annotations, int64_t* id_chars,
                   int64_t* clade_chars);
int  wepp_mat_get(const wepp_mat* m, int32_t* parent, int64_t* mut_off, int32_t* mut_pos, uint8_t* mut_ref,
                  uint8_t* mut_par, uint8_t* mut_nuc, int64_t* id_off, char* id_chars);
/* clade_annotations: entry v * n_annotations + k (src/mutation_annotated_tree.cpp:560-564) */
int  wepp_mat_get_clades(const wepp_mat* m, int64_t* clade_off, char* clade_chars);
/* Parsimony::data bytes of a tree given as flat arrays (leaf names = ids of childless nodes; internal
 * nodes are renamed node_<k> by any loader).  Returns the size; writes at most `capacity` bytes.  */
int64_t wepp_mat_serialize(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                           const uint8_t* mut_ref, const uint8_t* mut_par, const uint8_t* mut_nuc, const int64_t* id_off,
                           const char* id_chars, void* out, int64_t capacity);

/* The collapsed reads: load_reads_from_proto (src/WEPP/sam2pb.cpp:489-549) over `Sam::sam`
 * (sam.proto:4-18).  A read's mutations are the positions where its content differs from the reference
 * and is not '_' (N included, id 15); end = start + len - 1.  reverse columns: key k owns values
 * rev_off[k] .. rev_off[k+1].  */
typedef struct wepp_readset wepp_readset;
int  wepp_reads_load(const char* path, const char* ref_seq, int64_t ref_len, int32_t n_threads, wepp_readset** out);
int  wepp_reads_parse(const void* pb_bytes, int64_t n_bytes, const char* ref_seq, int64_t ref_len, int32_t n_threads,
                      wepp_readset** out);
void wepp_reads_free(wepp_readset* r);
int  wepp_reads_dims(const wepp_readset* r, int64_t* n_reads, int64_t* n_muts, int64_t* name_chars, int64_t* n_rev_keys,
                     int64_t* n_rev_vals, int64_t* rev_key_chars, int64_t* rev_val_chars);
int  wepp_reads_get(const wepp_readset* r, int32_t* start, int32_t* end, int32_t* degree, int64_t* rm_off,
                    int32_t* rm_pos, uint8_t* rm_nuc, int64_t* name_off, char* name_chars);
int  wepp_reads_get_reverse(const wepp_readset* r, int64_t* key_off, char* key_chars, int64_t* rev_off,
                            int64_t* val_off, char* val_chars);

/* ---- The `wepp` command line (process boundary) -----------------------------------------
 * Everything `build/wepp <command> ...` does, argv as main() gets it: `detectPeaks` (src/WEPP/main.cpp:47-49,
 * pipeline.cpp:5-80: loaders -> arena -> initial filter on the GPU -> <P>_checkpoint.txt -> iterative Freyja
 * post filter -> the result files of src/WEPP/arena.cpp:446-931), `sam2PB` (main.cpp:50-52, sam2pb.cpp:54-109),
 * `help`.  Same flags and defaults as src/WEPP/util.cpp:145-158, same files, same exit codes (0; 1 on a fatal
 * error or an invalid command; 0 for help / no command, main.cpp:36-45).  The executable built at build/wepp
 * is a three-line main() around this call.  */
int wepp_cli_main(int argc, const char* const* argv);

/* Host-only introspection (no GPU needed; used by the CPU test-suite): the Euler-tour event
 * stripes built from an arena (4 x uint32 per entry: preorder idx, position, signed-delta bytes
 * for read allele ref/A/C/G, delta byte for T) and the read bucketing plan.  Both return a
 * count (>= 0) or a negative WEPP_E_* code.  */
int64_t wepp_host_euler_stripes(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off, const int32_t* mut_pos,
                                const uint8_t* mut_ref, const uint8_t* mut_nuc, int32_t genome_size,
                                int32_t stripe_width, uint32_t* entries, int64_t capacity, int64_t* stripe_off,
                                int32_t stripe_off_len);
int64_t wepp_host_read_plan(int32_t genome_size, int32_t stripe_width, int32_t reads_per_lane, int64_t n_reads,
                            const int32_t* start, const int32_t* end, const int32_t* degree, const int64_t* rm_off,
                            const int32_t* rm_pos, const uint8_t* rm_nuc, int64_t* perm, int32_t* qs, int32_t* qe,
                            int32_t* bin, int32_t* reads_per_tile);

#ifdef __cplusplus
}
#endif
#endif /* WEPP_B200_H */
