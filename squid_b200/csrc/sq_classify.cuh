// Filter rules: record gate, consecutive-duplicate test, concordance and partial-alignment class.
// One record in, one class byte out; the -mq / -pl / 750 kb / proper-pair rules of the reference's
// per-record preamble (SegmentGraph.cpp:297-318, 651-689; ReadRec.cpp:119-141).
#ifndef SQ_CLASSIFY_CUH
#define SQ_CLASSIFY_CUH
#include "sq_common.cuh"

namespace sq {

// own(l) == [synthetic mate block of r]   (lists compared on RefID, RefPos, MatchRef; ReadRec.cpp:125)
SQ_HD bool own_equals_mate_of(const DevBatch &b, int64_t l, int64_t r) {
    const uint32_t ol = b.blk_off[l], nl = b.blk_off[l + 1] - ol;
    if (!has_mate_block(b.flag[r], b.mate_ref_id[r])) return nl == 0;
    return nl == 1 && b.ref_id[l] == b.mate_ref_id[r] && b.blk_ref_pos[ol] == b.mate_pos[r] && b.blk_match_ref[ol] == kMateBlockLen;
}

// ReadRec_t::Equal(lastreadrec, tmpreadrec) for two alignment records, each carrying its own
// blocks (sorted by read position) plus the synthetic 15-bp mate block (SegmentGraph.cpp:305-315).
// Equal = direct match or mate-swapped match; written out it does not depend on which record is
// the first mate:  [own==own && mate==mate] || [own(l)==mate(r) && mate(l)==own(r)].
SQ_HD bool records_equal(const DevBatch &b, int64_t l, int64_t r) {
    const uint32_t ol = b.blk_off[l], nl = b.blk_off[l + 1] - ol;
    const uint32_t orr = b.blk_off[r], nr = b.blk_off[r + 1] - orr;
    const bool ml = has_mate_block(b.flag[l], b.mate_ref_id[l]), mr = has_mate_block(b.flag[r], b.mate_ref_id[r]);
    bool direct = (nl == nr) && (ml == mr) && (b.ref_id[l] == b.ref_id[r] || nl == 0);
    if (direct && ml) direct = b.mate_ref_id[l] == b.mate_ref_id[r] && b.mate_pos[l] == b.mate_pos[r];
    if (direct && nl > 0) {
        // both lists are sorted by read position; same strand => same permutation of CIGAR order,
        // different strand => compare through the sorted views
        Blk x[kMaxBlocks], y[kMaxBlocks];
        const int cx = load_sorted_blocks(b, l, x), cy = load_sorted_blocks(b, r, y);
        for (int k = 0; k < cx && k < cy; k++)
            if (x[k].ref_pos != y[k].ref_pos || x[k].match_ref != y[k].match_ref) { direct = false; break; }
    }
    if (direct) return true;
    return own_equals_mate_of(b, l, r) && own_equals_mate_of(b, r, l);
}

// Equal(default-constructed ReadRec_t, r): both lists of r empty.
SQ_HD bool record_equals_empty(const DevBatch &b, int64_t r) {
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
};

// `prev` = index of the previous gate-passing record, or -1.
SQ_HD ClassifyOut classify_record(const DevBatch &b, const Params &p, int64_t r, int64_t prev) {
    ClassifyOut o;
    o.cls = 0; o.other_key = 0; o.first_len = 0;
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
    return o;
}

}  // namespace sq
#endif
