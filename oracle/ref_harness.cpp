// TEST INFRASTRUCTURE (oracle side).  Not part of the shipped product; only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may execute the binary built from this file.
//
// Drives the reference's OWN, UNMODIFIED sources (compiled in place from /root/reference/src by
// oracle/Makefile against the shims in oracle/shim/) through the segment-graph construction
// path and dumps the state at every seam of SURVEY.md §8b:
//   BuildNode_STAR (SegmentGraph.cpp:192)  -> nodes.bin
//   BuildEdges     (SegmentGraph.cpp:1932) -> edges.bin, chim_after_edges.bin (LocateRead trims)
//   the host-side filters of the constructor (SegmentGraph.cpp:111-122) -> final_nodes.bin, final_edges.bin
//   ExactBreakpoint (:3019) -> exactbp.bin ; ExactBPConcordantSupport (:3083) -> support.bin
// Ordering()/GLPK is never reached (main.cpp:41 is skipped on purpose: off the hot path).
//
// Determinism note: BuildNode_STAR dereferences bamdiscordant.cend() after
// the last discordant group (SegmentGraph.cpp:606,620,633,640,644).  To pin that read, this
// harness replaces global operator new with a zero-filling allocator with 64 bytes of zeroed
// slack, so the one-past-the-end element always reads as {RefID 0, RefPos 0, MatchRef 0}.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include "SegmentGraph.h"
#include "Config.h"
#include "WriteIO.h"

void *operator new(size_t n) {
    void *p = calloc(1, n + 64);
    if (!p) throw std::bad_alloc();
    return p;
}
void *operator new[](size_t n) {
    void *p = calloc(1, n + 64);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void *p) noexcept { free(p); }
void operator delete[](void *p) noexcept { free(p); }
void operator delete(void *p, size_t) noexcept { free(p); }
void operator delete[](void *p, size_t) noexcept { free(p); }

static std::string outdir;
static void dump_i32(const std::string &name, const std::vector<int32_t> &v) {
    FILE *f = fopen((outdir + "/" + name).c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", name.c_str()); exit(2); }
    if (!v.empty()) fwrite(v.data(), 4, v.size(), f);
    fclose(f);
}
static void dump_f64(const std::string &name, const std::vector<double> &v) {
    FILE *f = fopen((outdir + "/" + name).c_str(), "wb");
    if (!v.empty()) fwrite(v.data(), 8, v.size(), f);
    fclose(f);
}
static void dump_nodes(const std::string &stem, const SegmentGraph_t &g) {
    std::vector<int32_t> a;
    std::vector<double> d;
    for (const Node_t &n : g.vNodes) {
        a.push_back(n.Chr); a.push_back(n.Position); a.push_back(n.Length); a.push_back(n.Support);
        d.push_back(n.AvgDepth);
    }
    dump_i32(stem + "_i32.bin", a);
    dump_f64(stem + "_f64.bin", d);
}
static void dump_edges(const std::string &name, const std::vector<Edge_t> &e) {
    std::vector<int32_t> a;
    for (const Edge_t &x : e) {
        a.push_back(x.Ind1); a.push_back(x.Ind2); a.push_back(x.Head1); a.push_back(x.Head2); a.push_back(x.Weight);
    }
    dump_i32(name, a);
}
static void dump_chim(const std::string &name, const SBamrecord_t &c) {
    std::vector<int32_t> a;
    for (size_t i = 0; i < c.size(); i++) {
        for (int m = 0; m < 2; m++) {
            const std::vector<SingleBamRec_t> &v = m ? c[i].SecondMate : c[i].FirstRead;
            for (const SingleBamRec_t &b : v) {
                a.push_back((int32_t)i); a.push_back(m); a.push_back(b.RefID); a.push_back(b.RefPos); a.push_back(b.ReadPos);
                a.push_back(b.MatchRef); a.push_back(b.MatchRead); a.push_back(b.IsReverse);
            }
        }
    }
    dump_i32(name, a);
    std::vector<int32_t> t;
    for (size_t i = 0; i < c.size(); i++) {
        // *LowPhred of a mate that was never seen is uninitialised in the reference;
        // dump it only where the mate owns blocks or a total length.
        t.push_back(c[i].FirstTotalLen); t.push_back(c[i].SecondTotalLen);
        t.push_back(c[i].FirstTotalLen ? (int)c[i].FirstLowPhred : -1);
        t.push_back(c[i].SecondTotalLen ? (int)c[i].SecondLowPhred : -1);
    }
    dump_i32(name + ".meta", t);
}
static void dump_bpmap(const std::string &name, const map<Edge_t, vector<pair<int, int> > > &m) {
    std::vector<int32_t> a;
    for (auto it = m.begin(); it != m.end(); ++it)
        for (const pair<int, int> &p : it->second) {
            a.push_back(it->first.Ind1); a.push_back(it->first.Ind2); a.push_back(it->first.Head1); a.push_back(it->first.Head2);
            a.push_back(p.first); a.push_back(p.second);
        }
    dump_i32(name, a);
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    // BWA mode (src/main.cpp:31-36 with an empty -c): `--bwa`, the second path is ignored ("-"):
    // one BAM carries concordant and discordant alignments; BuildNode_BWA + RawEdges run instead of the STAR functions.
    // usage: squid_ref <concordant.sqmb> <chimeric.sqmb> <outdir> [-mq N] [-pl N] [-pm N] [-pt 0|1] [-dp N] [-di N] [-w N] [-r X] [-a N] [--stop-after nodes|edges|filters]
    if (argc < 4) { fprintf(stderr, "usage: %s conc.sqmb chim.sqmb outdir [opts]\n", argv[0]); return 2; }
    std::string conc = argv[1], chim = argv[2];
    outdir = argv[3];
    UsingSTAR = true;
    Min_MapQual = 255;  // Config.cpp:221-222 (STAR, no -mq)
    std::string stop = "";
    bool quiet = false, bwa = false, mq_given = false, write_outputs = false;
    for (int i = 4; i < argc; i++) {
        std::string a = argv[i];
        auto nxt = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return std::string(argv[++i]); };
        if (a == "-mq") { Min_MapQual = (uint16_t)atoi(nxt().c_str()); mq_given = true; }
        else if (a == "--bwa") bwa = true;
        else if (a == "-pl") Max_LowPhred_Len = (uint16_t)atoi(nxt().c_str());
        else if (a == "-pm") Min_Phred = (uint8_t)atoi(nxt().c_str());
        else if (a == "-pt") Phred_Type = atoi(nxt().c_str()) != 0;
        else if (a == "-dp") Concord_Dist_Pos = atoi(nxt().c_str());
        else if (a == "-di") Concord_Dist_Idx = atoi(nxt().c_str());
        else if (a == "-w") Min_Edge_Weight = atoi(nxt().c_str());
        else if (a == "-r") DiscordantRatio = atof(nxt().c_str());
        else if (a == "-a") MaxAllowedDegree = atoi(nxt().c_str());
        else if (a == "--stop-after") stop = nxt();
        else if (a == "--quiet") quiet = true;
        else if (a == "--write-outputs") write_outputs = true;  // ref_graph.txt (OutputGraph) and ref_sv.txt (WriteBEDPE under a stand-in ordering)
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (bwa) { UsingSTAR = false; if (!mq_given) Min_MapQual = 1; }  // Config.cpp:22 default; 255 is the STAR override (:221-222)
    FILE *saved_stdout = nullptr;
    (void)saved_stdout;
    if (quiet) { if (!freopen("/dev/null", "w", stdout)) return 2; }

    map<string, int> RefTable;
    vector<string> RefName;
    vector<int> RefLength;
    BuildRefName(conc, RefName, RefTable, RefLength);
    if (RefLength.empty()) { fprintf(stderr, "cannot read %s\n", conc.c_str()); return 2; }
    SBamrecord_t Chimrecord;
    double t0 = now();
    if (!bwa) BuildChimericSBamRecord(Chimrecord, RefName, chim);  // main.cpp:32-34: only with -c
    double t_chim = now() - t0;
    dump_chim("chim_loaded.bin", Chimrecord);
    { std::vector<int32_t> v{(int32_t)ReadLen}; dump_i32("readlen.bin", v); }

    SegmentGraph_t g;
    t0 = now();
    if (bwa) g.BuildNode_BWA(RefLength, conc); else g.BuildNode_STAR(RefLength, Chimrecord, conc);  // SegmentGraph.cpp:105-108
    double t_nodes = now() - t0;
    if (bwa) { std::vector<int32_t> v{(int32_t)ReadLen}; dump_i32("readlen.bin", v); }  // BuildNode_BWA infers ReadLen itself (:857-864)
    dump_nodes("nodes", g);
    double t_edges = 0, t_filters = 0, t_exactbp = 0, t_cov = 0;
    size_t n_final_nodes = 0, n_final_edges = 0;
    if (stop != "nodes") {
        t0 = now();
        g.BuildEdges(Chimrecord, conc);
        t_edges = now() - t0;
        dump_edges("edges_i32.bin", g.vEdges);
        dump_chim("chim_after_edges.bin", Chimrecord);
        if (stop != "edges") {
            t0 = now();
            g.FilterbyWeight();
            vector<bool> KeepEdge;
            g.FilterbyInterleaving(KeepEdge);
            g.FilterEdges(KeepEdge);
            g.CompressNode();
            g.FurtherCompressNode();
            g.ConnectedComponent();
            g.MultiplyDisEdges();
            t_filters = now() - t0;
            dump_nodes("final_nodes", g);
            dump_edges("final_edges_i32.bin", g.vEdges);
            dump_i32("labels_i32.bin", std::vector<int32_t>(g.Label.begin(), g.Label.end()));
            n_final_nodes = g.vNodes.size(); n_final_edges = g.vEdges.size();
            if (stop != "filters") {
                map<Edge_t, vector<pair<int, int> > > ExactBP, Support;
                t0 = now();
                g.ExactBreakpoint(Chimrecord, ExactBP);
                t_exactbp = now() - t0;
                dump_bpmap("exactbp_i32.bin", ExactBP);
                dump_chim("chim_after_exactbp.bin", Chimrecord);
                t0 = now();
                g.ExactBPConcordantSupport(conc, Chimrecord, ExactBP, Support);
                t_cov = now() - t0;
                dump_bpmap("support_i32.bin", Support);
                if (write_outputs) {
                    // main.cpp:37-38, 61-62 without the GLPK ordering in between (out of scope, not installed): the component
                    // ordering is a deterministic stand-in (below), dumped so that the writer twin is fed the very same one.
                    g.OutputGraph(outdir + "/ref_graph.txt");
                    int n_lab = 0;
                    for (size_t i = 0; i < g.Label.size(); i++) n_lab = std::max(n_lab, g.Label[i] + 1);
                    vector<vector<int> > Components(n_lab);
                    // orientations: every discordant edge, in vEdges order, orients its two nodes the way that makes it consistent
                    // with index order (WriteIO.cpp:61) unless one of them is already oriented; every third component is then
                    // reverse-complemented as a whole, which moves its edges to the mirrored test (WriteIO.cpp:63)
                    std::vector<int> sign(g.vNodes.size(), 0);
                    for (size_t e = 0; e < g.vEdges.size(); e++) {
                        const Edge_t &ed = g.vEdges[e];
                        if (!g.IsDiscordant((int)e) || sign[ed.Ind1] || sign[ed.Ind2] || ed.Ind1 == ed.Ind2) continue;
                        sign[ed.Ind1] = ed.Head1 ? -1 : 1; sign[ed.Ind2] = ed.Head2 ? 1 : -1;
                    }
                    for (size_t i = 0; i < g.vNodes.size(); i++) Components[g.Label[i]].push_back(sign[i] < 0 ? -(int)(i + 1) : (int)(i + 1));
                    for (size_t c = 0; c < Components.size(); c += 3) {
                        std::reverse(Components[c].begin(), Components[c].end());
                        for (size_t j = 0; j < Components[c].size(); j++) Components[c][j] = -Components[c][j];
                    }
                    {
                        std::vector<int32_t> flat;
                        flat.push_back((int32_t)Components.size());
                        for (size_t c = 0; c < Components.size(); c++) { flat.push_back((int32_t)Components[c].size()); for (size_t j = 0; j < Components[c].size(); j++) flat.push_back(Components[c][j]); }
                        dump_i32("components_i32.bin", flat);
                    }
                    vector<pair<int, int> > Node_NewChr; Node_NewChr.resize(g.vNodes.size());
                    for (unsigned int i = 0; i < Components.size(); i++)
                        for (unsigned int j = 0; j < Components[i].size(); j++) Node_NewChr[abs(Components[i][j]) - 1] = make_pair(i, j);
                    dump_edges("edges_before_demultiply_i32.bin", g.vEdges);
                    g.DeMultiplyDisEdges();
                    WriteBEDPE(outdir + "/ref_sv.txt", g, Components, Node_NewChr, RefName, ExactBP, Support);
                }
            }
        }
    }
    FILE *f = fopen((outdir + "/timings.json").c_str(), "w");
    fprintf(f, "{\"load_chimeric_s\": %.6f, \"build_nodes_s\": %.6f, \"build_edges_s\": %.6f, \"host_filters_s\": %.6f, \"exact_breakpoint_s\": %.6f, \"bp_coverage_s\": %.6f, \"read_len\": %d, \"n_chim\": %zu, \"n_final_nodes\": %zu, \"n_final_edges\": %zu}\n",
            t_chim, t_nodes, t_edges, t_filters, t_exactbp, t_cov, (int)ReadLen, Chimrecord.size(), n_final_nodes, n_final_edges);
    fclose(f);
    return 0;
}
