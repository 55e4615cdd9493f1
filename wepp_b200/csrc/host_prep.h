// host_prep.h — host-side flattening for the placement path.
//
// (1) Tree side: turns the preorder arena (reference: arena::from_mat, src/WEPP/arena.cpp:3-56)
//     into Euler-tour "entries": every mutation event of node v becomes an ENTER entry at
//     preorder index v and an EXIT entry at v's subtree end, each carrying a signed-delta
//     table — the change in mismatch count the event causes for a read whose allele at that
//     position is ref / A / C / G / T (N and "outside the read" never change anything).
//     The table folds in the reference's merge semantics (src/WEPP/initial_filter.cpp:59-87):
//     the state after an event depends only on the read allele and the event's mut/ref
//     nucleotides; the state before it is that of the nearest ancestor event at the same
//     position, or the seed set (:118-123) when there is none.
// (2) Read side: sorts reads into window buckets (stripe-aligned genome intervals x count
//     bin) and cuts the buckets into warp tiles.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace wepp {

// Device/host shared 16-byte Euler entry.
//  master (stripe) form : x = sort key (preorder idx << 1 | point), y = absolute position, z = delta
//                         bytes for codes 0..3 (ref,A,C,G), w = byte0 delta for code 4 (T), other bytes 0.
//                         A BOUNDARY entry (point = 0) changes the running prefix from its index on: an
//                         event of an internal node v is an ENTER at v and an EXIT (negated) at v's
//                         subtree end.  A POINT entry (point = 1) is an event of a leaf v: v's own score
//                         is the running prefix plus the deltas of v's point entries; the prefix itself
//                         is untouched, so leaves cost one entry per event instead of two.
//  bucket-list form     : x = idx | flags (ENT_EVAL, ENT_POINT, ENT_SKIP, see kernels.cuh), y = countable
//                         nodes evaluated at this entry (point entries: bits 8.. = preceding entries of
//                         the same leaf), z as above, w = byte0 delta T, byte1 = 0 (code 5: N / outside),
//                         bytes 2..3 = position - bucket start.
struct Entry {
    uint32_t x, y, z, w;
};

struct EulerStripes {
    int32_t stripe_width = 32;
    int32_t n_stripes = 0;
    int64_t n_events = 0;           // events with a non-zero delta table
    std::vector<int64_t> stripe_off;  // n_stripes + 1
    std::vector<Entry> entries;       // grouped by stripe (= pos / stripe_width), idx ascending within
};

// Returns "" on success, else an error message.
std::string build_euler_stripes(int32_t n_nodes, const int32_t* parent, const int64_t* mut_off,
                                const int32_t* mut_pos, const uint8_t* mut_ref, const uint8_t* mut_nuc,
                                int32_t genome_size, int32_t stripe_width, EulerStripes& out);

struct ListDesc {      // one Euler list per distinct stripe range
    int64_t off;       // first entry in the concatenated list buffer (entry 0 is the dummy at idx 0)
    int32_t n;         // entries including the dummy
    int32_t qs, qe;    // stripe range, inclusive
    int32_t b0;        // first position covered = qs * stripe_width
    int32_t width;     // positions covered
    int32_t pad;
};
struct BucketDesc {    // list x count bin: owns one accumulator row
    int64_t acc_off;   // offset of this bucket's per-segment accumulators
    int32_t list;
    int32_t bin;       // min(start / (G/50), 49)  (initial_filter.cpp:146,169)
};
struct TileDesc {
    int64_t first;     // first read (sorted order)
    int32_t count;     // reads in tile (<= 32*K)
    int32_t bucket;
};

struct ReadPlan {
    int64_t n_reads = 0;
    int64_t n_read_muts = 0;        // mutations of the selected reads
    int32_t reads_per_tile = 256;
    int32_t max_width = 0;
    std::vector<int64_t> perm;      // bucket-sorted position -> caller's read index (the reads stay in caller order)
    std::vector<ListDesc> lists;
    std::vector<BucketDesc> buckets;
    std::vector<TileDesc> tiles;    // longest lists first
    int64_t list_entries_total = 0;
    int64_t acc_total = 0;
    int64_t scanned_entries = 0;       // sum over tiles of list length
    int64_t scanned_read_entries = 0;  // sum over reads of list length
};

// Shared memory left for place_kernel's selector table on a B200 CTA (232,448 B opt-in maximum minus the kernel's
// fixed areas; kept in step with kernels.cuh by a static_assert in wepp_abi.cu).
constexpr int32_t PLACE_TABLE_BYTES = 232448 - 50192;

// Read validation errors (shared by the host keying and the device keying kernel).
enum { RP_OK = 0, RP_ERR_WINDOW = 1, RP_ERR_DEGREE = 2, RP_ERR_OFFSETS = 3, RP_ERR_MUT_ORDER = 4, RP_ERR_MUT_CODE = 5 };
const char* read_plan_error(int code);

// Host keying (place_subset, the peak loop, and the fallback of wepp_set_reads): validates the
// reads, keys them by (stripe range, count bin) and produces the permutation + descriptors.
// reads_per_lane: 0 = pick from the widest bucket, else 2/4/8.
std::string build_read_plan(const EulerStripes& es, int32_t genome_size, int64_t n_reads, const int32_t* start,
                            const int32_t* end, const int32_t* degree, const int64_t* rm_off, const int32_t* rm_pos,
                            const uint8_t* rm_nuc, int32_t reads_per_lane, const int64_t* subset, int64_t n_subset,
                            ReadPlan& out);

// Bucket coarsening.  A read may use ANY list whose stripe range contains its window (positions outside the
// window select class "outside" and change nothing), at the price of a longer list.  For one (first stripe,
// count bin), `spans` holds the (qe - qs, reads) pairs in ascending span order; target[i] is the index of the
// span whose list the reads of span i use.  Greedy, smallest span first: a group moves up to the next span when
// that does not cost more tile-entries (tiles x list width) — sparse buckets stop paying for nearly empty tiles.
void merge_span_chain(const std::vector<std::pair<int32_t, int64_t>>& spans, int64_t reads_per_tile, std::vector<int32_t>& target);
int32_t reads_per_lane_for_width(int32_t width);

// Descriptor half of the plan, from per-bucket read counts (out.lists / out.buckets already filled).
ListDesc make_list_desc(const EulerStripes& es, int32_t qs, int32_t qe);
std::string finish_read_plan(const EulerStripes& es, int32_t reads_per_lane, const std::vector<int64_t>& bucket_count,
                             ReadPlan& out, std::vector<int64_t>& first);

}  // namespace wepp
