// Range-shard planner of the multi-GPU path: cuts the sorted concordant stream into contiguous record ranges at
// CLEAN cuts -- record indices at which every piece of state that BuildNode_STAR, BuildEdges' depth cursor and the
// 0-coverage rules carry from record to record is provably reset, so that each shard can run the path on its own records and
// the per-shard results (seed ops, depth numerators, edge tables, coverage counts) combine to the single-stream result
// bit for bit.  A record c is a clean cut when
//   A. it starts more than ReadLen + 270 bp right of the end of every earlier alignment on its chromosome (GetEndPosition,
//      so spliced reads that jump over the gap count), or it is the first record of a chromosome: a 0-coverage record for
//      whatever group is pending (SegmentGraph.cpp:616-620), windows emptied (:633-636), pending segment closed (:621-630);
//      no block of an earlier read lies at or right of it (ConcordRest :690-699, depth cursor :784-803, LocateRead hints);
//   B. it is itself a kept record with an aligned block at its position (gate :297-303, not Equal to its predecessor :315-318):
//      the shard that starts with it sees it exactly as the whole stream does;
//   C. every discordant group lies either more than ReadLen + 200 bp left of the second-to-last kept record before c (the
//      group is triggered and the depth streams' break index (:338-339) found inside the earlier shard) or more than
//      ReadLen + 200 bp right of c (nothing the group looks at lies left of the cut);
//   D. the earlier shard holds at least one record that updates otherrightmost (:684-689), so its running maximum at the
//      cut does not depend on the shards before it.
// The chimeric reads (and with them the groups) are replicated on every shard.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../sq_classify.cuh"
#include "prepass.h"
#include "squid_b200.h"

namespace {
using namespace sq;

struct Planner {
    DevBatch b;  // host pointers
    Params p;
    const std::vector<Group> *G;
    const std::vector<DiscBlock> *D;
    int32_t margin;

    bool gate(int64_t r) const { return record_gate(b.flag[r], b.mapq[r], b.aux[r], b.ref_id[r], p.min_mapq); }
    int64_t prev_gate(int64_t r, int64_t floor) const {
        for (int64_t q = r - 1; q >= floor; q--) if (gate(q)) return q;
        return -1;
    }
    ClassifyOut classify(int64_t r, int64_t floor) const { return classify_record(b, p, r, prev_gate(r, floor)); }

    // conditions B-D for a cut at record c (condition A holds); `lo` = first record of the shard that would end at c
    bool clean(int64_t c, int64_t lo) const {
        const int32_t RL = p.read_len;
        // B
        if (b.blk_off[c + 1] == b.blk_off[c] || b.blk_ref_pos[b.blk_off[c]] != b.pos[c]) return false;
        if (!(classify(c, 0).cls & CLS_KEEP)) return false;
        // the two last kept records before c, and an otherrightmost update, inside [lo, c)
        int64_t kept[2] = {-1, -1};
        int nk = 0;
        bool upd = false;
        for (int64_t q = c - 1; q >= lo && (nk < 2 || !upd) && c - q < (1 << 20); q--) {
            if (!gate(q)) continue;
            const ClassifyOut o = classify(q, 0);
            if ((o.cls & CLS_KEEP) && nk < 2) kept[nk++] = q;
            if (o.other_key != 0) upd = true;
        }
        if (nk < 2 || !upd) return false;
        const int32_t c2 = b.ref_id[kept[1]], p2 = b.pos[kept[1]], ci = b.ref_id[c], pi = b.pos[c];
        // C: first group that is not clear of the cut on the left must be clear of it on the right
        const std::vector<Group> &g = *G;
        size_t lo_g = 0, hi_g = g.size();
        while (lo_g < hi_g) {  // right ends are increasing along the sorted groups
            const size_t m = (lo_g + hi_g) >> 1;
            const bool left = g[m].chr < c2 || (g[m].chr == c2 && (int64_t)g[m].right + RL + margin < p2);
            if (left) lo_g = m + 1; else hi_g = m;
        }
        if (lo_g < g.size()) {
            const Group &x = g[lo_g];
            const int64_t start = (*D)[x.ds].pos;
            const bool right = x.chr > ci || (x.chr == ci && start - RL - margin > pi);
            if (!right) return false;
        }
        return true;
    }
};
}  // namespace

extern "C" int sqg_plan_shards(const sqg_batch *hb, const sqg_chimeric *chim, const sqg_config *cfg, int32_t n_ref, int32_t n_shards,
                               int64_t *cuts, int32_t *n_planned) {
    if (!hb || !chim || !cfg || !cuts || !n_planned || n_shards < 1 || hb->n_rec < 0) return SQG_EINVAL;
    if (!cfg->using_star) return SQG_EUNSUPPORTED;
    const int64_t n = hb->n_rec;
    sqh::ChimPrepass pre;
    sqh::chimeric_prepass(*chim, n_ref, cfg->read_len, pre);
    Planner P;
    P.b.n_rec = hb->n_rec; P.b.n_blk = hb->n_blk;
    P.b.ref_id = hb->ref_id; P.b.pos = hb->pos; P.b.mate_ref_id = hb->mate_ref_id; P.b.mate_pos = hb->mate_pos; P.b.end_pos = hb->end_pos;
    P.b.flag = hb->flag; P.b.total_len = hb->total_len; P.b.lowphred_run = hb->lowphred_run; P.b.mapq = hb->mapq; P.b.aux = hb->aux;
    P.b.blk_off = hb->blk_off; P.b.blk_ref_pos = hb->blk_ref_pos; P.b.blk_match_ref = hb->blk_match_ref;
    P.b.blk_read_pos = hb->blk_read_pos; P.b.blk_match_read = hb->blk_match_read;
    P.p.min_mapq = cfg->min_mapq; P.p.max_lowphred_len = cfg->max_lowphred_len; P.p.concord_dist_pos = cfg->concord_dist_pos;
    P.p.concord_dist_idx = cfg->concord_dist_idx; P.p.read_len = cfg->read_len; P.p.n_ref = n_ref;
    P.G = &pre.groups; P.D = &pre.disc; P.margin = 200;
    const int64_t gap = (int64_t)cfg->read_len + kIslandSlack + P.margin;

    int32_t made = 0;
    cuts[0] = 0;
    int64_t target = n_shards > 1 ? n / n_shards : n;
    int32_t cur_chr = -2;
    int64_t run_end = 0;
    for (int64_t i = 0; i < n && made + 1 < n_shards; i++) {
        const int32_t c = hb->ref_id[i];
        if (c < 0) break;  // unmapped tail stays with the last shard
        const bool new_chr = c != cur_chr;
        if (new_chr) { cur_chr = c; run_end = -(1ll << 40); }
        if (i >= target && i > cuts[made] && (new_chr || (int64_t)hb->pos[i] - run_end > gap) && P.clean(i, cuts[made])) {
            cuts[++made] = i;
            target = std::max<int64_t>(i + 1, n / n_shards * (int64_t)(made + 1));
        }
        int64_t e = hb->end_pos[i];
        if (hb->blk_off[i + 1] > hb->blk_off[i]) {
            const uint32_t k = hb->blk_off[i + 1] - 1;
            e = std::max<int64_t>(e, (int64_t)hb->blk_ref_pos[k] + hb->blk_match_ref[k]);
        }
        e = std::max<int64_t>(e, hb->pos[i]);
        if (e > run_end) run_end = e;
    }
    cuts[made + 1] = n;
    *n_planned = made + 1;
    return SQG_OK;
}
