// The binding INTEGRATION.md describes, compiled for real: the four seams of SegmentGraph_t that make up the hot path (and one
// of the next rows),
// redefined on top of libsquid_b200.so (include/squid_b200.h, include/squid_b200_host.h).  This file is compiled against the
// reference's OWN headers and linked with the reference's OWN, unmodified objects (ReadRec.cpp, SegmentGraph.cpp, WriteIO.cpp,
// Config.cpp, built in place from /root/reference by integration/Makefile); the reference's definitions of these four member
// functions are demoted to weak symbols with objcopy, so every caller inside the reference -- the constructor
// (src/SegmentGraph.cpp:104-109), main (src/main.cpp:58-60) -- now lands here.  Nothing of the reference is edited or copied.
//
//   SegmentGraph_t::BuildNode_STAR            src/SegmentGraph.cpp:192   -> sqg_load_concordant + sqg_load_chimeric + sqg_build_nodes
//   SegmentGraph_t::BuildEdges                src/SegmentGraph.cpp:1932  -> sqg_build_edges (+ UpdateNodeLink, the reference's own)
//   SegmentGraph_t::ExactBreakpoint           src/SegmentGraph.cpp:3019  -> sqh_exact_breakpoint (host twin)
//   SegmentGraph_t::ExactBPConcordantSupport  src/SegmentGraph.cpp:3083  -> sqg_bp_coverage between the reference's own glue
//   SegmentGraph_t::ConnectedComponent        src/SegmentGraph.cpp:2986  -> sqg_connected_components
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "SegmentGraph.h"
#include "squid_b200.h"
#include "squid_b200_host.h"

namespace {

struct Binding {
    sqg_ctx *ctx = nullptr;
    sqh_case *conc = nullptr;  // packed concordant BAM (host twin of the per-record decode, one pass instead of three)
    // flat copy of Chimrecord (sqg_chimeric)
    std::vector<uint32_t> read_off;
    std::vector<uint16_t> n_first;
    std::vector<int32_t> first_total, second_total, b_ref_id, b_ref_pos, b_read_pos, b_match_ref, b_match_read;
    std::vector<uint8_t> first_low, second_low, multi, b_rev;
    sqg_chimeric view;
} B;

[[noreturn]] void fail(const char *what, const char *msg) {
    fprintf(stderr, "squid_b200 binding: %s: %s\n", what, msg ? msg : "");
    exit(3);
}

// Chimrecord -> sqg_chimeric: FirstRead blocks, then SecondMate blocks, per read
void pack_chimrecord(const SBamrecord_t &C) {
    B.read_off.assign(1, 0); B.n_first.clear(); B.first_total.clear(); B.second_total.clear(); B.first_low.clear(); B.second_low.clear(); B.multi.clear();
    B.b_ref_id.clear(); B.b_ref_pos.clear(); B.b_read_pos.clear(); B.b_match_ref.clear(); B.b_match_read.clear(); B.b_rev.clear();
    for (const ReadRec_t &r : C) {
        for (int m = 0; m < 2; m++)
            for (const SingleBamRec_t &s : (m ? r.SecondMate : r.FirstRead)) {
                B.b_ref_id.push_back(s.RefID); B.b_ref_pos.push_back(s.RefPos); B.b_read_pos.push_back(s.ReadPos);
                B.b_match_ref.push_back(s.MatchRef); B.b_match_read.push_back(s.MatchRead); B.b_rev.push_back(s.IsReverse ? 1 : 0);
            }
        B.read_off.push_back((uint32_t)B.b_ref_id.size());
        B.n_first.push_back((uint16_t)r.FirstRead.size());
        // the reference leaves the totals / flags of a mate it never saw uninitialised; nothing reads them
        B.first_total.push_back(r.FirstRead.empty() ? 0 : r.FirstTotalLen); B.second_total.push_back(r.SecondMate.empty() ? 0 : r.SecondTotalLen);
        B.first_low.push_back(!r.FirstRead.empty() && r.FirstLowPhred); B.second_low.push_back(!r.SecondMate.empty() && r.SecondLowPhred);
        B.multi.push_back(0);
    }
    sqg_chimeric &v = B.view;
    v.n_reads = (int64_t)C.size(); v.n_blk = (int64_t)B.b_ref_id.size();
    v.read_off = B.read_off.data(); v.n_first = B.n_first.data(); v.first_total_len = B.first_total.data(); v.second_total_len = B.second_total.data();
    v.first_lowphred = B.first_low.data(); v.second_lowphred = B.second_low.data(); v.multi_filter = B.multi.data();
    v.blk_ref_id = B.b_ref_id.data(); v.blk_ref_pos = B.b_ref_pos.data(); v.blk_read_pos = B.b_read_pos.data();
    v.blk_match_ref = B.b_match_ref.data(); v.blk_match_read = B.b_match_read.data(); v.blk_is_reverse = B.b_rev.data();
}
// the trimmed blocks back into Chimrecord (LocateRead trims in place, src/SegmentGraph.cpp:1229-1248)
void unpack_blocks(SBamrecord_t &C) {
    size_t k = 0;
    for (ReadRec_t &r : C)
        for (int m = 0; m < 2; m++)
            for (SingleBamRec_t &s : (m ? r.SecondMate : r.FirstRead)) {
                s.RefPos = B.b_ref_pos[k]; s.ReadPos = B.b_read_pos[k]; s.MatchRef = B.b_match_ref[k]; s.MatchRead = B.b_match_read[k];
                k++;
            }
}

}  // namespace

void SegmentGraph_t::BuildNode_STAR(const vector<int> &RefLength, SBamrecord_t &Chimrecord, string bamfile) {
    sqh_options o;
    sqh_default_options(&o);
    o.phred33 = Phred_Type; o.max_lowphred_len = Max_LowPhred_Len; o.min_phred = Min_Phred; o.min_mapq = Min_MapQual;
    o.concord_dist_pos = Concord_Dist_Pos; o.concord_dist_idx = Concord_Dist_Idx;
    std::vector<const char *> names;
    for (const ReadRec_t &r : Chimrecord) names.push_back(r.Qname.c_str());  // ChimName (:196-201)
    char err[512] = "";
    if (sqh_open_concordant(bamfile.c_str(), names.data(), (int64_t)names.size(), &o, &B.conc, err, sizeof err)) fail("sqh_open_concordant", err);
    sqg_config cfg = {1, (int32_t)Max_LowPhred_Len, (int32_t)Min_MapQual, Concord_Dist_Pos, Concord_Dist_Idx, (int32_t)ReadLen};
    const int dev = getenv("SQUID_B200_DEVICE") ? atoi(getenv("SQUID_B200_DEVICE")) : 0;
    if (sqg_create(&B.ctx, &cfg, RefLength.data(), (int32_t)RefLength.size(), dev)) fail("sqg_create", B.ctx ? sqg_last_error(B.ctx) : "no CUDA device");
    if (sqg_load_concordant(B.ctx, sqh_case_batch(B.conc), 0)) fail("sqg_load_concordant", sqg_last_error(B.ctx));
    pack_chimrecord(Chimrecord);
    if (sqg_load_chimeric(B.ctx, &B.view)) fail("sqg_load_chimeric", sqg_last_error(B.ctx));
    int32_t *chr, *pos, *len, *cnt3, *sum3, other;
    int64_t n;
    if (sqg_build_nodes(B.ctx, &chr, &pos, &len, &n, &cnt3, &sum3, &other)) fail("sqg_build_nodes", sqg_last_error(B.ctx));
    vNodes.clear();
    vNodes.reserve((size_t)n);
    for (int64_t i = 0; i < n; i++) {  // Support / AvgDepth as :773-779, 785-801, 807-824 assemble them
        Node_t t(chr[i], pos[i], len[i], cnt3[i] + cnt3[n + i] + (other ? cnt3[2 * n + i] : 0));
        t.AvgDepth = sum3[i];
        t.AvgDepth += sum3[n + i];
        if (other) { t.AvgDepth += sum3[2 * n + i]; t.AvgDepth = 1.0 * t.AvgDepth / t.Length; }
        vNodes.push_back(t);
    }
}

void SegmentGraph_t::BuildEdges(SBamrecord_t &Chimrecord, string bamfile) {
    (void)bamfile;
    int32_t *i1, *i2, *w;
    uint8_t *hd;
    int64_t m;
    if (sqg_build_edges(B.ctx, &i1, &i2, &hd, &w, &m, &B.view)) fail("sqg_build_edges", sqg_last_error(B.ctx));
    vEdges.clear();
    vEdges.reserve((size_t)m);
    for (int64_t k = 0; k < m; k++) vEdges.push_back(Edge_t(i1[k], hd[k] & 1, i2[k], (hd[k] >> 1) & 1, w[k]));
    unpack_blocks(Chimrecord);
    UpdateNodeLink();  // :1960, the reference's own
}

void SegmentGraph_t::ExactBreakpoint(SBamrecord_t &Chimrecord, map<Edge_t, vector<pair<int, int> > > &ExactBP) {
    ExactBP.clear();
    pack_chimrecord(Chimrecord);
    std::vector<int32_t> c(vNodes.size()), p(vNodes.size()), l(vNodes.size());
    for (size_t i = 0; i < vNodes.size(); i++) { c[i] = vNodes[i].Chr; p[i] = vNodes[i].Position; l[i] = vNodes[i].Length; }
    int32_t *rows = nullptr;
    int64_t n = 0;
    if (sqh_exact_breakpoint(c.data(), p.data(), l.data(), (int64_t)c.size(), &B.view, Concord_Dist_Pos, Concord_Dist_Idx, &rows, &n)) fail("sqh_exact_breakpoint", "");
    for (int64_t k = 0; k < n; k++) {
        const int32_t *r = rows + 6 * k;
        ExactBP[Edge_t(r[0], r[2] != 0, r[1], r[3] != 0, 1)].push_back(make_pair(r[4], r[5]));
    }
    sqh_free(rows);
    unpack_blocks(Chimrecord);
}

void SegmentGraph_t::ExactBPConcordantSupport(string Input_BAM, SBamrecord_t &Chimrecord, const map<Edge_t, vector<pair<int, int> > > &ExactBP,
                                               map<Edge_t, vector<pair<int, int> > > &ExactBP_concord_support) {
    (void)Input_BAM; (void)Chimrecord;
    ExactBP_concord_support.clear();
    auto less = [](pair<int, int> a, pair<int, int> b) { return a.first != b.first ? a.first < b.first : a.second < b.second; };
    // the breakpoints of every edge (:3091-3109)
    auto bps_of = [&](const Edge_t &e, vector<pair<int, int> > &out) {
        map<Edge_t, vector<pair<int, int> > >::const_iterator it = ExactBP.find(e);
        if (it != ExactBP.cend() && it->second.size() != 0) {
            for (const pair<int, int> &q : it->second) { out.push_back(make_pair(vNodes[it->first.Ind1].Chr, q.first)); out.push_back(make_pair(vNodes[it->first.Ind2].Chr, q.second)); }
        } else {
            out.push_back(make_pair(vNodes[e.Ind1].Chr, vNodes[e.Ind1].Position + (e.Head1 ? 0 : vNodes[e.Ind1].Length)));
            out.push_back(make_pair(vNodes[e.Ind2].Chr, vNodes[e.Ind2].Position + (e.Head2 ? 0 : vNodes[e.Ind2].Length)));
        }
    };
    vector<pair<int, int> > BPs;
    for (const Edge_t &e : vEdges) bps_of(e, BPs);
    sort(BPs.begin(), BPs.end(), less);
    // the BAM pass (:3124-3168) on the device
    std::vector<int32_t> bc(BPs.size()), bp(BPs.size()), Coverages(BPs.size(), 0);
    for (size_t k = 0; k < BPs.size(); k++) { bc[k] = BPs[k].first; bp[k] = BPs[k].second; }
    if (sqg_bp_coverage(B.ctx, bc.data(), bp.data(), (int64_t)BPs.size(), Coverages.data())) fail("sqg_bp_coverage", sqg_last_error(B.ctx));
    // coverage of the breakpoint pairs of each edge (:3170-3211)
    for (const Edge_t &e : vEdges) {
        vector<pair<int, int> > mine, supports;
        bps_of(e, mine);
        for (size_t k = 0; k + 1 < mine.size(); k += 2) {
            const size_t a = (size_t)(lower_bound(BPs.begin(), BPs.end(), mine[k], less) - BPs.begin());
            const size_t b = (size_t)(lower_bound(BPs.begin(), BPs.end(), mine[k + 1], less) - BPs.begin());
            supports.push_back(make_pair(Coverages[a], Coverages[b]));
        }
        ExactBP_concord_support[e] = supports;
    }
}

void SegmentGraph_t::ConnectedComponent() {
    std::vector<int32_t> a(vEdges.size()), b(vEdges.size()), lab(vNodes.size());
    for (size_t k = 0; k < vEdges.size(); k++) { a[k] = vEdges[k].Ind1; b[k] = vEdges[k].Ind2; }
    const int dev = getenv("SQUID_B200_DEVICE") ? atoi(getenv("SQUID_B200_DEVICE")) : 0;
    int32_t nc = 0;
    if (sqg_connected_components(dev, (int64_t)vNodes.size(), a.data(), b.data(), (int64_t)a.size(), lab.data(), &nc)) fail("sqg_connected_components", "");
    Label.assign(lab.begin(), lab.end());
}
