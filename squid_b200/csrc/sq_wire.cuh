// Widening of the wire form of a record batch (include/squid_b200.h: sqg_wire) into the resident layout, tile by tile while the
// later chunks of the upload are still on the bus.  One CTA per wire tile (512 records): positions are a CTA scan of the deltas
// that restarts at every record listed in full (and at the tile head), the block offsets a CTA scan of the block counts on top of
// the tile's first block index (one packed scan gives the offsets in the batch and in the wire block arrays: a record with an
// implied block owns a block of the batch but none on the wire), and every block finds its record in the scanned offsets.
// Escaped values are looked up in the tile's slice of the exception lists (sorted by index).
// Pure data movement: ~18 B read and ~44 B written per record, far below the PCIe time of the chunk it overlaps.
#ifndef SQ_WIRE_CUH
#define SQ_WIRE_CUH
#include <cub/block/block_scan.cuh>

#include "sq_common.cuh"
#include "squid_b200.h"

namespace sq {

struct WireDev {  // device copies of the wire arrays that need widening (the others are uploaded straight into the batch)
    int64_t n_rec, n_blk, n_tiles;
    const int32_t *tile_ref_id, *tile_pos;
    const uint32_t *tile_blk_off, *tile_rec_exc_off, *tile_blk_exc_off, *tile_wblk_off;
    const uint16_t *dpos, *span; const int16_t *dmate;
    const uint8_t *lowphred_run, *aux_nblk;
    const uint16_t *blk_dref, *blk_match_ref16, *blk_read_pos, *blk_match_read;  // wire block arrays (explicit blocks only)
    int64_t n_wblk;
    const sqg_wire_rec_exc *rec_exc; const sqg_wire_blk_exc *blk_exc;
};
struct WireOut {
    int32_t *ref_id, *pos, *mate_ref_id, *mate_pos, *end_pos;
    uint16_t *lowphred_run; uint8_t *aux; uint32_t *blk_off;
    int32_t *blk_ref_pos, *blk_match_ref; uint16_t *blk_read_pos, *blk_match_read;
    int32_t *bad;  // set when the wire batch contradicts itself
};

constexpr int kWireTile = SQG_WIRE_TILE;

struct WirePos { int32_t pos, ref, abs; };  // abs: pos/ref are absolute (a restart of the running sum)
struct WirePosOp {
    __device__ __forceinline__ WirePos operator()(const WirePos &a, const WirePos &b) const { return b.abs ? b : WirePos{a.pos + b.pos, a.ref, a.abs}; }
};

__global__ void __launch_bounds__(kWireTile) k_wire_decode(WireDev w, WireOut o, int64_t tile0) {
    typedef cub::BlockScan<uint32_t, kWireTile> Scan;
    typedef cub::BlockScan<WirePos, kWireTile> ScanPos;
    __shared__ union { typename Scan::TempStorage cnt; typename ScanPos::TempStorage pos; } s_scan;
    __shared__ uint32_t s_off[kWireTile + 1];  // low 16 bits: blocks of the batch before the record, high 16 bits: blocks on the wire
    __shared__ int32_t s_pos[kWireTile], s_span[kWireTile];  // s_span < 0: the record's blocks are explicit
    const int64_t t = tile0 + blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t i = t * kWireTile + tid;
    const bool valid = i < w.n_rec;
    uint32_t nb = 0;
    bool implied = false;
    WirePos wp{0, 0, 0};
    int32_t mref = 0, mpos = 0, end = 0, dm = 0, sp = 0;
    uint16_t lpr = 0;
    uint8_t an = 0;
    bool listed = false;
    if (valid) {
        const uint16_t dp = w.dpos[i], sp16 = w.span[i];
        const int16_t dm16 = w.dmate[i];
        const uint8_t lp = w.lowphred_run[i];
        an = w.aux_nblk[i];
        lpr = lp; nb = an >> 4; dm = dm16; sp = sp16;
        if (nb == 14) { nb = 1; implied = true; }
        wp.pos = dp;
        if (tid == 0) { wp.pos = w.tile_pos[t] + (int32_t)dp; wp.ref = w.tile_ref_id[t]; wp.abs = 1; }
        if (dp == 0xFFFF || sp16 == 0xFFFF || dm16 == (int16_t)-32768 || lp == 255 || nb == 15) {
            uint32_t lo = w.tile_rec_exc_off[t], hi = w.tile_rec_exc_off[t + 1];
            const uint32_t end_ = hi;
            while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (w.rec_exc[m].idx < (uint32_t)i) lo = m + 1; else hi = m; }
            listed = true;
            if (!implied && lo < end_ && w.rec_exc[lo].idx == (uint32_t)i) {
                const sqg_wire_rec_exc e = w.rec_exc[lo];
                wp.pos = e.pos; wp.ref = e.ref_id; wp.abs = 1; mref = e.mate_ref_id; mpos = e.mate_pos; end = e.end_pos; lpr = e.lowphred_run; nb = e.n_blk;
            } else {  // an escape without its entry (or on a record that claims an implied block, which never needs one)
                atomicOr(o.bad, 1);
                wp.pos = w.tile_pos[t]; wp.ref = w.tile_ref_id[t]; wp.abs = 1; mref = wp.ref; mpos = wp.pos; end = wp.pos; nb = 0; implied = false;
            }
            if (nb > 64u) { atomicOr(o.bad, 8); nb = 0; }  // (keeps the packed scan inside its 16-bit halves; the path takes at most 16 blocks per record)
        }
    }
    ScanPos(s_scan.pos).InclusiveScan(wp, wp, WirePosOp());
    __syncthreads();  // (the two scans share their scratch)
    const int32_t pos = wp.pos;
    if (valid) {
        if (!listed) { mref = wp.ref; mpos = pos + dm; end = pos + sp; }
        o.ref_id[i] = wp.ref; o.pos[i] = pos; o.mate_ref_id[i] = mref; o.mate_pos[i] = mpos; o.end_pos[i] = end;
        o.lowphred_run[i] = lpr; o.aux[i] = an & 15;
    }
    uint32_t excl, total;
    Scan(s_scan.cnt).ExclusiveSum(nb | ((implied ? 0u : nb) << 16), excl, total);
    s_off[tid] = excl; s_pos[tid] = pos; s_span[tid] = implied ? sp : -1;
    if (tid == 0) s_off[kWireTile] = total;
    const uint32_t b0 = w.tile_blk_off[t], wb0 = w.tile_wblk_off[t];
    const uint32_t tot_b = total & 0xFFFFu, tot_w = total >> 16;
    if (valid) o.blk_off[i] = b0 + (excl & 0xFFFFu);
    if (tid == 0) {
        if (b0 + tot_b != w.tile_blk_off[t + 1] || (uint64_t)b0 + tot_b > (uint64_t)w.n_blk || wb0 + tot_w != w.tile_wblk_off[t + 1] || (uint64_t)wb0 + tot_w > (uint64_t)w.n_wblk) atomicOr(o.bad, 2);
        if (t == w.n_tiles - 1) o.blk_off[w.n_rec] = w.tile_blk_off[t + 1];
    }
    __syncthreads();
    const bool sane = b0 + tot_b == w.tile_blk_off[t + 1] && (uint64_t)b0 + tot_b <= (uint64_t)w.n_blk && wb0 + tot_w == w.tile_wblk_off[t + 1] && (uint64_t)wb0 + tot_w <= (uint64_t)w.n_wblk;
    const uint32_t nblk = sane ? tot_b : 0u;  // (an inconsistent tile writes no block at all: the load is refused)
    for (uint32_t k = tid; k < nblk; k += kWireTile) {
        int lo = 0, hi = kWireTile;  // the record whose block range holds k: last j with (s_off[j] & 0xFFFF) <= k
        while (hi - lo > 1) { const int m = (lo + hi) >> 1; if ((s_off[m] & 0xFFFFu) <= k) lo = m; else hi = m; }
        const uint32_t g = b0 + k;
        int32_t rp, ml;
        uint16_t rpos, mread;
        const int32_t isp = s_span[lo];
        if (isp >= 0) {  // the block the record implies
            rp = s_pos[lo]; ml = isp; rpos = 0; mread = (uint16_t)isp;
        } else {
            const uint32_t wk = wb0 + (s_off[lo] >> 16) + (k - (s_off[lo] & 0xFFFFu));
            const uint16_t dr = w.blk_dref[wk], mr = w.blk_match_ref16[wk];
            rpos = w.blk_read_pos[wk]; mread = w.blk_match_read[wk];
            if (dr == 0xFFFF || mr == 0xFFFF) {
                uint32_t l2 = w.tile_blk_exc_off[t], h2 = w.tile_blk_exc_off[t + 1];
                const uint32_t end_ = h2;
                while (l2 < h2) { const uint32_t m = (l2 + h2) >> 1; if (w.blk_exc[m].idx < wk) l2 = m + 1; else h2 = m; }
                if (l2 < end_ && w.blk_exc[l2].idx == wk) { rp = w.blk_exc[l2].ref_pos; ml = w.blk_exc[l2].match_ref; }
                else { atomicOr(o.bad, 4); rp = s_pos[lo]; ml = 0; }
            } else {
                rp = s_pos[lo] + (int32_t)dr; ml = (int32_t)mr;
            }
        }
        o.blk_ref_pos[g] = rp; o.blk_match_ref[g] = ml; o.blk_read_pos[g] = rpos; o.blk_match_read[g] = mread;
    }
}

}  // namespace sq
#endif
