// Shared device-side vocabulary of the B200 segment-graph path: the HBM-resident record batch,
// the segment (node) table and the per-record helpers every kernel uses.
// The element-level rules are `SQ_HD` (host+device) so that tests/emul can step the very same
// functions on the CPU while no GPU is attached; the product only ever runs them inside kernels.
#ifndef SQ_COMMON_CUH
#define SQ_COMMON_CUH
#include <stdint.h>

#if defined(__CUDACC__)
#define SQ_HD __host__ __device__ __forceinline__
#else
#define SQ_HD inline
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

struct Params {
    int32_t min_mapq, max_lowphred_len, concord_dist_pos, concord_dist_idx, read_len, n_ref;
};

// Segment table: segments tile every chromosome; chr_first[c]..chr_first[c+1] are the segments of c.
struct NodeTable {
    int32_t n = 0, n_ref = 0;
    const int32_t *chr = nullptr, *pos = nullptr, *end = nullptr;  // end = Position + Length
    const int32_t *chr_first = nullptr;                            // n_ref + 1
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

// Own blocks of record r sorted by read position (SortbyReadPos: std::sort on <=16 elements is an
// insertion sort, i.e. stable).  Returns the count.
SQ_HD int load_sorted_blocks(const DevBatch &b, int64_t r, Blk *out) {
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
