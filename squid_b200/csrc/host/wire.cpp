// Packer of the wire form of a record batch (include/squid_b200.h: sqg_wire): what a front end hands to
// sqg_load_concordant_wire() instead of the 32 + 12 B resident layout, because the PCIe transfer is what an end-to-end call of the
// path (the replacement of the reference's three BamReader passes, SegmentGraph.cpp:293-296, 1570-1577, 3126-3129) waits for.
// Tile-parallel: pass 1 counts the entries of the two exception lists per tile, pass 2 fills everything.
#include <cuda_runtime.h>
#include <omp.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "squid_b200_host.h"

namespace {
struct WireOwner {
    sqg_wire w;  // first member: the handle handed out is &w
    bool pinned;
    std::vector<void *> mem;
};
void *wire_alloc(WireOwner *o, size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) bytes = 8;
    if (o->pinned) { if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; } }
    else p = malloc(bytes);
    if (p) o->mem.push_back(p);
    return p;
}
void wire_release(WireOwner *o) {
    for (void *p : o->mem) { if (o->pinned) cudaFreeHost(p); else free(p); }
    delete o;
}
struct RecCode { uint16_t dpos, span; int16_t dmate; uint8_t lp, nb; bool exc, implied; };  // nb = block code (0..13 explicit, 14 implied, 15 escape)
inline RecCode code_record(const sqg_batch &b, int64_t i, int32_t tref, int32_t tpos) {  // (tref, tpos) = the previous record, or the record itself at the head of a tile
    RecCode c;
    const int64_t dp = (int64_t)b.pos[i] - tpos, sp = (int64_t)b.end_pos[i] - b.pos[i], dm = (int64_t)b.mate_pos[i] - b.pos[i];
    const uint32_t nb = b.blk_off[i + 1] - b.blk_off[i];
    const bool pos_ok = b.ref_id[i] == tref && dp >= 0 && dp < 0xFFFF;
    const bool span_ok = sp >= 0 && sp < 0xFFFF;
    const bool mate_ok = b.mate_ref_id[i] == b.ref_id[i] && dm > -32768 && dm <= 32767;
    const bool lp_ok = b.lowphred_run[i] < 255, nb_ok = nb < 14;
    c.exc = !(pos_ok && span_ok && mate_ok && lp_ok && nb_ok);
    // one block that the record implies: an unclipped, unspliced read (only for records that need no exception entry)
    c.implied = false;
    if (!c.exc && nb == 1) {
        const uint32_t k = b.blk_off[i];
        c.implied = b.blk_ref_pos[k] == b.pos[i] && (int64_t)b.blk_match_ref[k] == sp && b.blk_read_pos[k] == 0 && (int64_t)b.blk_match_read[k] == sp;
    }
    c.dpos = pos_ok ? (uint16_t)dp : 0xFFFF;
    c.span = span_ok ? (uint16_t)sp : 0xFFFF;
    c.dmate = mate_ok ? (int16_t)dm : (int16_t)-32768;
    c.lp = lp_ok ? (uint8_t)b.lowphred_run[i] : 255;
    c.nb = c.implied ? 14 : (nb_ok ? (uint8_t)nb : 15);
    return c;
}
// a block is listed in full when its record's position is (its dref is relative to it) or when one of its own values does not fit
inline bool block_exc(const sqg_batch &b, int64_t i, int64_t k, bool rec_pos_exc) {
    const int64_t dr = (int64_t)b.blk_ref_pos[k] - b.pos[i];
    return rec_pos_exc || dr < 0 || dr >= 0xFFFF || b.blk_match_ref[k] < 0 || b.blk_match_ref[k] >= 0xFFFF;
}
}  // namespace

extern "C" {

int sqh_pack_wire(const sqg_batch *bp, int32_t pinned, sqg_wire **out) {
    if (!bp || !out || bp->n_rec < 0 || bp->n_blk < 0) return SQG_EINVAL;
    *out = nullptr;
    const sqg_batch &b = *bp;
    const int64_t n = b.n_rec, nb = b.n_blk, T = SQG_WIRE_TILE, nt = (n + T - 1) / T;
    if (n > 0 && (b.blk_off[0] != 0 || (int64_t)b.blk_off[n] != nb)) return SQG_EINVAL;
    WireOwner *o = new (std::nothrow) WireOwner();
    if (!o) return SQG_ENOMEM;
    o->pinned = pinned != 0;
    memset(&o->w, 0, sizeof(o->w));
    std::vector<uint32_t> n_rexc((size_t)nt + 1, 0), n_bexc((size_t)nt + 1, 0), n_wb((size_t)nt + 1, 0);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t t = 0; t < nt; t++) {
        const int64_t i0 = t * T, i1 = i0 + T < n ? i0 + T : n;
        uint32_t re = 0, be = 0, wb = 0;
        for (int64_t i = i0; i < i1; i++) {
            if (b.blk_off[i + 1] < b.blk_off[i] || b.blk_off[i + 1] - b.blk_off[i] > 0xFFFFu || (int64_t)b.blk_off[i + 1] > nb || b.aux[i] >= 16) { bad = 1; continue; }
            const int64_t pv = i > i0 ? i - 1 : i;
            const RecCode c = code_record(b, i, b.ref_id[pv], b.pos[pv]);
            re += c.exc;
            if (c.implied) continue;
            wb += b.blk_off[i + 1] - b.blk_off[i];
            for (int64_t k = b.blk_off[i]; k < b.blk_off[i + 1]; k++) be += block_exc(b, i, k, c.dpos == 0xFFFF);
        }
        n_rexc[(size_t)t + 1] = re; n_bexc[(size_t)t + 1] = be; n_wb[(size_t)t + 1] = wb;
    }
    if (bad) { wire_release(o); return SQG_EINVAL; }
    uint64_t tr = 0, tb = 0, tw = 0;
    for (int64_t t = 0; t < nt; t++) { tr += n_rexc[(size_t)t + 1]; tb += n_bexc[(size_t)t + 1]; tw += n_wb[(size_t)t + 1]; }
    if (tr > 0xFFFFFFFFull || tb > 0xFFFFFFFFull) { wire_release(o); return SQG_EUNSUPPORTED; }
    sqg_wire &w = o->w;
    w.n_rec = n; w.n_blk = nb; w.n_tiles = nt; w.n_rec_exc = (int64_t)tr; w.n_blk_exc = (int64_t)tb; w.n_wblk = (int64_t)tw;
    const int64_t nwb = (int64_t)tw;
#define A(type, field, cnt) type *field = (type *)wire_alloc(o, sizeof(type) * (size_t)(cnt)); if (!field) { wire_release(o); return SQG_ENOMEM; } w.field = field
    A(int32_t, tile_ref_id, nt); A(int32_t, tile_pos, nt);
    A(uint32_t, tile_blk_off, nt + 1); A(uint32_t, tile_rec_exc_off, nt + 1); A(uint32_t, tile_blk_exc_off, nt + 1); A(uint32_t, tile_wblk_off, nt + 1);
    A(uint16_t, dpos, n); A(uint16_t, span, n); A(int16_t, dmate, n); A(uint16_t, flag, n); A(uint16_t, total_len, n);
    A(uint8_t, lowphred_run, n); A(uint8_t, mapq, n); A(uint8_t, aux_nblk, n);
    A(uint16_t, blk_dref, nwb); A(uint16_t, blk_match_ref, nwb); A(uint16_t, blk_read_pos, nwb); A(uint16_t, blk_match_read, nwb);
    A(sqg_wire_rec_exc, rec_exc, tr); A(sqg_wire_blk_exc, blk_exc, tb);
#undef A
    tile_rec_exc_off[0] = 0; tile_blk_exc_off[0] = 0; tile_wblk_off[0] = 0;
    for (int64_t t = 0; t < nt; t++) {
        tile_rec_exc_off[t + 1] = tile_rec_exc_off[t] + n_rexc[(size_t)t + 1];
        tile_blk_exc_off[t + 1] = tile_blk_exc_off[t] + n_bexc[(size_t)t + 1];
        tile_wblk_off[t + 1] = tile_wblk_off[t] + n_wb[(size_t)t + 1];
    }
    tile_blk_off[nt] = (uint32_t)nb;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < nt; t++) {
        const int64_t i0 = t * T, i1 = i0 + T < n ? i0 + T : n;
        tile_ref_id[t] = b.ref_id[i0]; tile_pos[t] = b.pos[i0]; tile_blk_off[t] = b.blk_off[i0];
        uint32_t re = tile_rec_exc_off[t], be = tile_blk_exc_off[t], wk = tile_wblk_off[t];
        for (int64_t i = i0; i < i1; i++) {
            const int64_t pv = i > i0 ? i - 1 : i;
            const RecCode c = code_record(b, i, b.ref_id[pv], b.pos[pv]);
            dpos[i] = c.dpos; span[i] = c.span; dmate[i] = c.dmate; lowphred_run[i] = c.lp; aux_nblk[i] = (uint8_t)(b.aux[i] | (c.nb << 4));
            if (c.exc) {
                sqg_wire_rec_exc e;
                e.idx = (uint32_t)i; e.ref_id = b.ref_id[i]; e.pos = b.pos[i]; e.mate_ref_id = b.mate_ref_id[i]; e.mate_pos = b.mate_pos[i]; e.end_pos = b.end_pos[i];
                e.lowphred_run = b.lowphred_run[i]; e.n_blk = (uint16_t)(b.blk_off[i + 1] - b.blk_off[i]);
                rec_exc[re++] = e;
            }
            if (c.implied) continue;
            for (int64_t k = b.blk_off[i]; k < b.blk_off[i + 1]; k++, wk++) {
                blk_read_pos[wk] = b.blk_read_pos[k]; blk_match_read[wk] = b.blk_match_read[k];
                if (block_exc(b, i, k, c.dpos == 0xFFFF)) {
                    blk_dref[wk] = 0xFFFF; blk_match_ref[wk] = 0xFFFF;
                    blk_exc[be++] = sqg_wire_blk_exc{wk, b.blk_ref_pos[k], b.blk_match_ref[k]};
                } else {
                    blk_dref[wk] = (uint16_t)(b.blk_ref_pos[k] - b.pos[i]); blk_match_ref[wk] = (uint16_t)b.blk_match_ref[k];
                }
            }
        }
    }
    // the fields shipped as they are
    const int nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static)
    for (int p = 0; p < nthreads; p++) {
        const int64_t r0 = n * p / nthreads, r1 = n * (p + 1) / nthreads;
        memcpy(flag + r0, b.flag + r0, (size_t)(r1 - r0) * 2); memcpy(total_len + r0, b.total_len + r0, (size_t)(r1 - r0) * 2);
        memcpy(mapq + r0, b.mapq + r0, (size_t)(r1 - r0));
    }
    *out = &o->w;
    return SQG_OK;
}

void sqh_free_wire(sqg_wire *w) {
    if (w) wire_release(reinterpret_cast<WireOwner *>(w));
}

int64_t sqh_wire_bytes(const sqg_wire *w) {
    if (!w) return 0;
    return w->n_rec * 13 + w->n_wblk * 8 + w->n_tiles * 8 + (w->n_tiles + 1) * 16 + w->n_rec_exc * (int64_t)sizeof(sqg_wire_rec_exc) + w->n_blk_exc * (int64_t)sizeof(sqg_wire_blk_exc);
}

}  // extern "C"
