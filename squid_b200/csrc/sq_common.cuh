// Shared device-side vocabulary of the B200 segment-graph path: the HBM-resident record batch,
// the segment (node) table and the per-record helpers every kernel uses.
// The element-level rules are `SQ_HD` (host+device) so that tests/emul can step the very same
// functions on the CPU while no GPU is attached; the product only ever runs them inside kernels.
#ifndef SQ_COMMON_CUH
#define SQ_COMMON_CUH
#include <stdint.h>

#if defined(__CUDACC__)
#define SQ_HD __host__ __device__ __forceinline__
#define SQ_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SQ_HD inline
#define SQ_HD_NOINLINE inline
#endif

namespace sq {

constexpr int kMaxBlocks = 16;     // per mate (packer rejects more; keeps sorts insertion-only like std::sort's n<=16 path)
constexpr int kLocateTol = 5;      // LocateRead overhang tolerance   (SegmentGraph.cpp:1209)
constexpr int kSeedThresh = 3;     // BuildNode_STAR `thresh`          (SegmentGraph.cpp:286)
constexpr int kMateBlockLen = 15;  // synthetic mate block             (SegmentGraph.cpp:308)

// record class bits written by the classify kernel
enum : uint8_t {
    CLS_GATE = 1,      // passes the record gate (SegmentGraph.cpp:302 / 1584 / 3136)
    CLS_KEEP = 2,      // ... and is not Equal() to the previous gate-passing record (:315-318, 1597-1600)
    CLS_CONC = 4,      // kept, proper FR pair within 750 kb (:651-654) and owns >= 1 block (:655)
    CLS_PART = 8,      // CLS_CONC and partially aligned (:668-683) => PartialAlignCluster, else ConcordantCluster
    CLS_HASBLK = 16,   // kept and owns >= 1 block (feeds ReadsMain, :320-333)
    CLS_DISPL = 32,    // CLS_HASBLK and the first kept block does not start at the record position (leading block dropped)
    CLS_REST = 64,     // CLS_CONC with a mate flag and >= 2 blocks: its later blocks feed ConcordRest (:690-699)
};

// HBM-resident SoA batch (include/squid_b200.h: sqg_batch), device pointers.
struct DevBatch {
    int64_t n_rec = 0, n_blk = 0;
    const int32_t *ref_id = nullptr, *pos = nullptr, *mate_ref_id = nullptr, *mate_pos = nullptr, *end_pos = nullptr;
    const uint16_t *flag = nullptr, *total_len = nullptr, *lowphred_run = nullptr;
    const uint8_t *mapq = nullptr, *aux = nullptr;
    const uint32_t *blk_off = nullptr;
    const int32_t *blk_ref_pos = nullptr, *blk_match_ref = nullptr;
    const uint16_t *blk_read_pos = nullptr, *blk_match_read = nullptr;
};

// The same batch seen through a tile staged in shared memory: arrays indexed by the GLOBAL record / block index, backed by
// staging buffers that start at record `rec0` / block `blk0`.  The element rules are templates over the batch type, so the
// very same code classifies from HBM (DevBatch) and from a staged tile (TileBatch).
template <class T> struct OffPtr {
    const T *p; uint32_t off;  // indices stay below 2^32 (records: 2^31 per context, blocks: uint32 offsets)
    SQ_HD T operator[](int64_t i) const { return p[(uint32_t)i - off]; }
};
template <class T> SQ_HD const T *elem_ptr(const T *p, int64_t i) { return p + i; }
template <class T> SQ_HD const T *elem_ptr(const OffPtr<T> &q, int64_t i) { return q.p + ((uint32_t)i - q.off); }
struct TileBatch {
    int64_t n_rec = 0, n_blk = 0;
    OffPtr<int32_t> ref_id, pos, mate_ref_id, mate_pos, end_pos;
    OffPtr<uint16_t> flag, total_len, lowphred_run;
    OffPtr<uint8_t> mapq, aux;
    OffPtr<uint32_t> blk_off;
    OffPtr<int32_t> blk_ref_pos, blk_match_ref;
    OffPtr<uint16_t> blk_read_pos, blk_match_read;
};

struct Params {
    int32_t min_mapq, max_lowphred_len, concord_dist_pos, concord_dist_idx, read_len, n_ref;
};

// Segment table: segments tile every chromosome; chr_first[c]..chr_first[c+1] are the segments of c.
// bin_seg is a coarse position index over the tiling: for chromosome c, bin k (positions [k<<bin_shift, (k+1)<<bin_shift))
// starts inside segment bin_seg[bin_off[c] + k].  Every "which segment holds position x" question of the path (LocateRead's
// scans, the depth cursor, the spanning-block fallback) is then one table read plus a step or two along the tiling instead
// of a binary search over the whole table.  With bin_seg == nullptr the lookups fall back to binary searches.
struct NodeTable {
    int32_t n = 0, n_ref = 0;
    const int32_t *chr = nullptr, *pos = nullptr, *end = nullptr;  // end = Position + Length
    const int32_t *chr_first = nullptr;                            // n_ref + 1
    const int32_t *bin_seg = nullptr, *bin_off = nullptr;          // bin_off: n_ref + 1
    int32_t bin_shift = 12;
};

// An aligned block in registers/local memory (SingleBamRec_t).
struct Blk {
    int32_t ref_id, ref_pos, match_ref, read_pos, match_read;
    bool rev;
};

SQ_HD bool flag_mapped(uint16_t f) { return !(f & 0x4); }
SQ_HD bool flag_mate_mapped(uint16_t f) { return !(f & 0x8); }
SQ_HD bool flag_rev(uint16_t f) { return f & 0x10; }
SQ_HD bool flag_mate_rev(uint16_t f) { return f & 0x20; }
SQ_HD bool flag_first(uint16_t f) { return f & 0x40; }
SQ_HD bool flag_second(uint16_t f) { return f & 0x80; }
SQ_HD bool flag_dup(uint16_t f) { return f & 0x400; }
SQ_HD bool flag_proper(uint16_t f) { return f & 0x2; }

// Record gate shared by the three passes (SegmentGraph.cpp:302, 1584, 3136).  A mapped record with
// RefID -1 is rejected at load time, so the three gates coincide.
SQ_HD bool record_gate(uint16_t flag, uint8_t mapq, uint8_t aux, int32_t ref_id, int32_t min_mapq) {
    return !((aux & (1u | 2u | 8u)) || (int32_t)mapq < min_mapq || flag_dup(flag) || !flag_mapped(flag) || ref_id < 0);
}
SQ_HD bool has_mate_block(uint16_t flag, int32_t mate_ref_id) { return flag_mate_mapped(flag) && mate_ref_id != -1; }

// first index in [lo,hi) with a[i] >= v
SQ_HD int32_t lower_bound_i32(const int32_t *a, int32_t lo, int32_t hi, int32_t v) {
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [lo,hi) with a[i] > v
SQ_HD int32_t upper_bound_i32(const int32_t *a, int32_t lo, int32_t hi, int32_t v) {
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Segment of chromosome c (c0 = chr_first[c] <= result < c1 = chr_first[c+1], c1 > c0) that holds position x,
// clamped to the first / last segment of the chromosome when x lies outside [0, chromosome length).
SQ_HD int32_t seg_at(const NodeTable &nt, int32_t c, int32_t c0, int32_t c1, int32_t x) {
    if (x < nt.pos[c0]) return c0;
    if (x >= nt.end[c1 - 1]) return c1 - 1;
    if (nt.bin_seg) {
        const int32_t bo = nt.bin_off[c] + (x >> nt.bin_shift);
        int32_t j = nt.bin_seg[bo];
        if (nt.end[j] > x) return j;
        int32_t hi = bo + 1 < nt.bin_off[c + 1] ? nt.bin_seg[bo + 1] : c1 - 1;  // segment holding the start of the next bin
        if (hi - j <= 4) { do { j++; } while (nt.end[j] <= x); return j; }
        return upper_bound_i32(nt.end, j + 1, hi + 1, x);
    }
    return upper_bound_i32(nt.end, c0, c1, x);
}
// value of bin k of chromosome c: the segment holding position k << bin_shift (the last segment of c at the chromosome end)
SQ_HD int32_t bin_seg_value(const NodeTable &nt, int32_t c, int32_t k) {
    const int32_t c0 = nt.chr_first[c], c1 = nt.chr_first[c + 1];
    if (c1 <= c0) return c0;
    const int32_t j = upper_bound_i32(nt.end, c0, c1, k << nt.bin_shift);
    return j < c1 ? j : c1 - 1;
}
// first segment of c with End > v   (== upper_bound_i32(nt.end, c0, c1, v))
SQ_HD int32_t seg_first_end_gt(const NodeTable &nt, int32_t c, int32_t c0, int32_t c1, int32_t v) {
    if (c1 <= c0) return c0;
    if (v >= nt.end[c1 - 1]) return c1;
    return seg_at(nt, c, c0, c1, v);
}
// first segment of c with End >= v  (== lower_bound_i32(nt.end, c0, c1, v))
SQ_HD int32_t seg_first_end_ge(const NodeTable &nt, int32_t c, int32_t c0, int32_t c1, int32_t v) {
    return seg_first_end_gt(nt, c, c0, c1, v - 1);
}
// last segment of c with Position <= v, c0-1 if none  (== upper_bound_i32(nt.pos, c0, c1, v) - 1)
SQ_HD int32_t seg_last_pos_le(const NodeTable &nt, int32_t c, int32_t c0, int32_t c1, int32_t v) {
    if (c1 <= c0 || v < nt.pos[c0]) return c0 - 1;
    return seg_at(nt, c, c0, c1, v);
}

// Own blocks of record r sorted by read position (SortbyReadPos: std::sort on <=16 elements is an
// insertion sort, i.e. stable).  Returns the count.
template <class B>
SQ_HD int load_sorted_blocks(const B &b, int64_t r, Blk *out) {
    const uint32_t o = b.blk_off[r], n = b.blk_off[r + 1] - o;
    const int32_t rid = b.ref_id[r];
    const bool rev = flag_rev(b.flag[r]);
    int cnt = (int)n;
    if (cnt > kMaxBlocks) cnt = kMaxBlocks;
    for (int k = 0; k < cnt; k++) {
        Blk x;
        x.ref_id = rid; x.ref_pos = b.blk_ref_pos[o + k]; x.match_ref = b.blk_match_ref[o + k];
        x.read_pos = b.blk_read_pos[o + k]; x.match_read = b.blk_match_read[o + k]; x.rev = rev;
        int j = k;
        while (j > 0 && x.read_pos < out[j - 1].read_pos) { out[j] = out[j - 1]; j--; }
        out[j] = x;
    }
    return cnt;
}

}  // namespace sq
#endif
