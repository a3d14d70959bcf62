// Filter rules: record gate, consecutive-duplicate test, concordance and partial-alignment class.
// One record in, one class byte out; the -mq / -pl / 750 kb / proper-pair rules of the reference's
// per-record preamble (SegmentGraph.cpp:297-318, 651-689; ReadRec.cpp:119-141).
#ifndef SQ_CLASSIFY_CUH
#define SQ_CLASSIFY_CUH
#include "sq_common.cuh"

namespace sq {

// own(l) == [synthetic mate block of r]   (lists compared on RefID, RefPos, MatchRef; ReadRec.cpp:125)
template <class B>
SQ_HD bool own_equals_mate_of(const B &b, int64_t l, int64_t r) {
    const uint32_t ol = b.blk_off[l], nl = b.blk_off[l + 1] - ol;
    if (!has_mate_block(b.flag[r], b.mate_ref_id[r])) return nl == 0;
    return nl == 1 && b.ref_id[l] == b.mate_ref_id[r] && b.blk_ref_pos[ol] == b.mate_pos[r] && b.blk_match_ref[ol] == kMateBlockLen;
}

// Two block lists of the same length n > 1 (CIGAR order), each sorted by read position (SortbyReadPos: std::sort on <= 16
// elements is an insertion sort, i.e. stable), compared on (RefPos, MatchRef).  Kept out of line and fed with plain pointers:
// consecutive records with equal mate fields AND several blocks are rare, and the sort needs stack arrays.
SQ_HD_NOINLINE bool sorted_blocks_equal(const int32_t *lp, const int32_t *lm, const uint16_t *lr, const int32_t *rp, const int32_t *rm, const uint16_t *rr, int n) {
    int32_t xp[kMaxBlocks], xm[kMaxBlocks], xr[kMaxBlocks], yp[kMaxBlocks], ym[kMaxBlocks], yr[kMaxBlocks];
    if (n > kMaxBlocks) n = kMaxBlocks;
    for (int k = 0; k < n; k++) {
        int j = k;
        const int32_t q = lr[k];
        while (j > 0 && q < xr[j - 1]) { xp[j] = xp[j - 1]; xm[j] = xm[j - 1]; xr[j] = xr[j - 1]; j--; }
        xp[j] = lp[k]; xm[j] = lm[k]; xr[j] = q;
        j = k;
        const int32_t q2 = rr[k];
        while (j > 0 && q2 < yr[j - 1]) { yp[j] = yp[j - 1]; ym[j] = ym[j - 1]; yr[j] = yr[j - 1]; j--; }
        yp[j] = rp[k]; ym[j] = rm[k]; yr[j] = q2;
    }
    for (int k = 0; k < n; k++)
        if (xp[k] != yp[k] || xm[k] != ym[k]) return false;
    return true;
}

// ReadRec_t::Equal(lastreadrec, tmpreadrec) for two alignment records, each carrying its own
// blocks (sorted by read position) plus the synthetic 15-bp mate block (SegmentGraph.cpp:305-315).
// Equal = direct match or mate-swapped match; written out it does not depend on which record is
// the first mate:  [own==own && mate==mate] || [own(l)==mate(r) && mate(l)==own(r)].
template <class B>
SQ_HD bool records_equal(const B &b, int64_t l, int64_t r) {
    const uint32_t ol = b.blk_off[l], nl = b.blk_off[l + 1] - ol;
    const uint32_t orr = b.blk_off[r], nr = b.blk_off[r + 1] - orr;
    const bool ml = has_mate_block(b.flag[l], b.mate_ref_id[l]), mr = has_mate_block(b.flag[r], b.mate_ref_id[r]);
    bool direct = (nl == nr) && (ml == mr) && (b.ref_id[l] == b.ref_id[r] || nl == 0);
    if (direct && ml) direct = b.mate_ref_id[l] == b.mate_ref_id[r] && b.mate_pos[l] == b.mate_pos[r];
    if (direct && nl == 1) direct = b.blk_ref_pos[ol] == b.blk_ref_pos[orr] && b.blk_match_ref[ol] == b.blk_match_ref[orr];
    else if (direct && nl > 1)
        direct = sorted_blocks_equal(elem_ptr(b.blk_ref_pos, ol), elem_ptr(b.blk_match_ref, ol), elem_ptr(b.blk_read_pos, ol),
                                     elem_ptr(b.blk_ref_pos, orr), elem_ptr(b.blk_match_ref, orr), elem_ptr(b.blk_read_pos, orr), (int)nl);
    if (direct) return true;
    return own_equals_mate_of(b, l, r) && own_equals_mate_of(b, r, l);
}

// Equal(default-constructed ReadRec_t, r): both lists of r empty.
template <class B>
SQ_HD bool record_equals_empty(const B &b, int64_t r) {
    return b.blk_off[r + 1] == b.blk_off[r] && !has_mate_block(b.flag[r], b.mate_ref_id[r]);
}

// SegmentGraph.cpp:651-654
SQ_HD bool is_concordant_pair(uint16_t f, int32_t rid, int32_t pos, int32_t mrid, int32_t mpos) {
    if (!(flag_mapped(f) && flag_mate_mapped(f) && mrid != -1 && rid == mrid && flag_proper(f))) return false;
    if (flag_rev(f) && !flag_mate_rev(f)) return pos >= mpos && pos - mpos <= 750000;
    if (!flag_rev(f) && flag_mate_rev(f)) return mpos >= pos && mpos - pos <= 750000;
    return false;
}

struct ClassifyOut {
    uint8_t cls;
    uint64_t other_key;  // (chr+1)<<32 | end of the CIGAR-first block when the record updates otherrightmost, else 0
    int32_t first_len;   // length of the first kept block of a CLS_CONC record, else 0
    int32_t cc_end;      // end of the first kept block of a ConcordantCluster entry (CLS_CONC, not CLS_PART), else kNoCcEnd
};
constexpr int32_t kNoCcEnd = -(1 << 30);
constexpr int32_t kCcWalkTile = 0x7fffffff;  // per-tile maximum of cc_end: "this tile holds several chromosomes, walk it"

// `prev` = index of the previous gate-passing record, or -1.
template <class B>
SQ_HD ClassifyOut classify_record(const B &b, const Params &p, int64_t r, int64_t prev) {
    ClassifyOut o;
    o.cls = 0; o.other_key = 0; o.first_len = 0; o.cc_end = kNoCcEnd;
    const uint16_t f = b.flag[r];
    const int32_t rid = b.ref_id[r];
    if (!record_gate(f, b.mapq[r], b.aux[r], rid, p.min_mapq)) return o;
    o.cls = CLS_GATE;
    const bool dup = prev < 0 ? record_equals_empty(b, r) : records_equal(b, prev, r);
    if (dup) return o;
    o.cls |= CLS_KEEP;
    const uint32_t off = b.blk_off[r], nb = b.blk_off[r + 1] - off;
    if (nb == 0) return o;
    o.cls |= CLS_HASBLK;
    if (b.blk_ref_pos[off] != b.pos[r]) o.cls |= CLS_DISPL;
    if (!is_concordant_pair(f, rid, b.pos[r], b.mate_ref_id[r], b.mate_pos[r])) return o;
    o.cls |= CLS_CONC;
    o.first_len = b.blk_match_ref[off];
    const bool fm = flag_first(f), sm = flag_second(f);
    if (fm || sm) {
        if (nb > 1) o.cls |= CLS_REST;
        o.other_key = ((uint64_t)(uint32_t)(rid + 1) << 32) | (uint32_t)(b.blk_ref_pos[off] + b.blk_match_ref[off]);
        // partial alignment: > 15 unaligned bases at either end of the read and no low-phred run (:668-683)
        const bool low = (int32_t)b.lowphred_run[r] > p.max_lowphred_len;
        if (!low) {
            int32_t front_rp = 0x7fffffff, back_rp = -1, back_mr = 0;
            for (uint32_t k = 0; k < nb; k++) {  // sorted-by-read-pos front/back; ties keep CIGAR order (stable)
                const int32_t rp = b.blk_read_pos[off + k];
                if (rp < front_rp) front_rp = rp;
                if (rp >= back_rp) { back_rp = rp; back_mr = b.blk_match_read[off + k]; }
            }
            if (front_rp > 15 || (int32_t)b.total_len[r] - back_rp - back_mr > 15) o.cls |= CLS_PART;
        }
    }
    if (!(o.cls & CLS_PART)) o.cc_end = b.blk_ref_pos[off] + b.blk_match_ref[off];
    return o;
}

}  // namespace sq
#endif
