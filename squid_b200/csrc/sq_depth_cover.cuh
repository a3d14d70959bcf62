// Per-segment depth (Support / AvgDepth numerators) and breakpoint coverage rules.
// Reference behaviour: BuildNode_STAR part D (SegmentGraph.cpp:765-826) and
// ExactBPConcordantSupport's BAM pass (SegmentGraph.cpp:3129-3166).
#ifndef SQ_DEPTH_COVER_CUH
#define SQ_DEPTH_COVER_CUH
#include "sq_common.cuh"

namespace sq {

constexpr int32_t kNoNode = 0x7fffffff;

SQ_HD bool depth_contained(const NodeTable &nt, int32_t j, int32_t chr, int32_t start, int32_t len) {  // :789, :811
    return nt.chr[j] == chr && start >= nt.pos[j] - kSeedThresh && start + len <= nt.end[j] + kSeedThresh;
}

// The merge loop at :784-803 moves a segment cursor forward only: a read is tested against the
// segment max(cursor, n(start)) where n(start) is the first segment of its chromosome that ends
// right of `start`; blocks of <= 3 bp can already be "contained" (+-3) in one of the <= 3
// segments just left of n(start), which then is where the cursor stops.  This returns that
// per-read target m; the cursor is the running maximum of m over the stream.
SQ_HD int32_t depth_target(const NodeTable &nt, int32_t chr, int32_t start, int32_t len) {
    if (chr < 0 || chr >= nt.n_ref) return kNoNode;
    const int32_t c0 = nt.chr_first[chr], c1 = nt.chr_first[chr + 1];
    const int32_t n = seg_first_end_gt(nt, chr, c0, c1, start);  // first segment with End > start
    if (len <= kSeedThresh) {
        for (int32_t j = (n - 3 > c0 ? n - 3 : c0); j < n; j++)
            if (depth_contained(nt, j, chr, start, len)) return j;
    }
    return n < c1 ? n : kNoNode;  // a start at/after the chromosome end walks the cursor off the table
}

// ---- breakpoint coverage ------------------------------------------------------------------------

// Pass-3 record filter (:3136-3142): gate, then keep only the right-hand record of a same-chromosome pair.
SQ_HD bool cover_qualifies(uint8_t cls, uint16_t f, int32_t rid, int32_t pos, int32_t mrid, int32_t mpos) {
    if (!(cls & CLS_GATE)) return false;
    if (flag_mate_mapped(f) && mrid == rid) {
        if (mpos > pos) return false;
        if (mpos == pos && flag_second(f)) return false;
    }
    return true;
}
SQ_HD int32_t cover_start(uint16_t f, int32_t rid, int32_t pos, int32_t mrid, int32_t mpos) {  // :3148-3155
    return (flag_mate_mapped(f) && mrid == rid) ? mpos : pos;
}
SQ_HD uint64_t chrpos_key(int32_t chr, int32_t pos) { return ((uint64_t)(uint32_t)(chr + 1) << 32) | (uint32_t)pos; }

// first index in [lo,hi) with a[i] >= v (uint64 keys)
SQ_HD int64_t lower_bound_u64(const uint64_t *a, int64_t lo, int64_t hi, uint64_t v) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
SQ_HD int64_t upper_bound_u64(const uint64_t *a, int64_t lo, int64_t hi, uint64_t v) {
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace sq
#endif
