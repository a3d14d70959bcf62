// Host twin of the two output formats of the path:
//   <prefix>_graph.txt  SegmentGraph_t::OutputGraph  (src/SegmentGraph.cpp:3223-3234)
//   <prefix>_sv.txt     WriteBEDPE                   (src/WriteIO.cpp:45-124) with DeMultiplyDisEdges (src/SegmentGraph.cpp:3012-3017)
// Same text, byte for byte: the same iostream formatting (default precision for AvgDepth, `endl` / '\n' as the reference
// uses them) and, for the BEDPE, the same unstable std::sort by weight on the same edge order (the tie order of equal
// weights is libstdc++'s introsort permutation; reproduced by making the same call on an element of the same shape).
#include <algorithm>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "squid_b200_host.h"

namespace {
struct EdgeRow {  // Edge_t (src/BPEdge.h:24-29)
    int Ind1, Ind2;
    bool Head1, Head2;
    int Weight, GroupWeight;
};
struct EdgeLess {  // Edge_t::operator< (src/BPEdge.h:59-70)
    bool operator()(const EdgeRow &a, const EdgeRow &b) const {
        if (a.Ind1 != b.Ind1) return a.Ind1 < b.Ind1;
        if (a.Ind2 != b.Ind2) return a.Ind2 < b.Ind2;
        if (a.Head1 != b.Head1) return (int)a.Head1 < (int)b.Head1;
        if (a.Head2 != b.Head2) return (int)a.Head2 < (int)b.Head2;
        return false;
    }
};
typedef std::map<EdgeRow, std::vector<std::pair<int, int>>, EdgeLess> BpMap;
void rows_to_map(const int32_t *rows6, int64_t n, BpMap &m) {
    for (int64_t i = 0; i < n; i++) {
        const int32_t *r = rows6 + 6 * i;
        m[EdgeRow{r[0], r[1], r[2] != 0, r[3] != 0, 0, 0}].emplace_back(r[4], r[5]);
    }
}
}  // namespace

extern "C" {

int sqh_write_graph(const char *path, const int32_t *chr, const int32_t *pos, const int32_t *len, const int32_t *support, const double *avg_depth,
                    const int32_t *label, int64_t n_nodes, const int32_t *ind1, const int32_t *ind2, const uint8_t *head1, const uint8_t *head2,
                    const int32_t *weight, int64_t n_edges) {
    if (!path || n_nodes < 0 || n_edges < 0) return SQG_EINVAL;
    std::ofstream output(path, std::ios::out);
    if (!output) return SQG_EINVAL;
    output << "# type=node\tid\tChr\tPosition\tEnd\tSupport\tAvgDepth\tLabel\n";
    output << "# type=edge\tid\tInd1\tHead1\tInd2\tHead2\tWeight\n";
    for (int64_t i = 0; i < n_nodes; i++)
        output << "node\t" << (int)i << '\t' << chr[i] << '\t' << pos[i] << '\t' << (pos[i] + len[i]) << '\t' << support[i] << '\t' << avg_depth[i] << '\t' << label[i] << '\n';
    for (int64_t i = 0; i < n_edges; i++)
        output << "edge\t" << (int)i << '\t' << ind1[i] << '\t' << (head1[i] ? "H\t" : "T\t") << ind2[i] << '\t' << (head2[i] ? "H\t" : "T\t") << weight[i] << std::endl;
    output.close();
    return output.fail() ? SQG_EINVAL : SQG_OK;
}

int sqh_write_bedpe(const char *path, const char *const *ref_name, int32_t n_ref, const int32_t *chr, const int32_t *pos, const int32_t *len, int64_t n_nodes,
                    const int32_t *ind1, const int32_t *ind2, const uint8_t *head1, const uint8_t *head2, const int32_t *weight, int64_t n_edges,
                    const int64_t *comp_off, const int32_t *comp_nodes, int64_t n_comp, const int32_t *exactbp_rows6, int64_t n_exactbp,
                    const int32_t *support_rows6, int64_t n_support, double discordant_ratio, int32_t concord_dist_pos, int32_t concord_dist_idx) {
    if (!path || !ref_name || n_nodes < 0 || n_edges < 0 || n_comp < 0) return SQG_EINVAL;
    auto discordant = [&](const EdgeRow &e) {  // IsDiscordant(Edge_t), src/SegmentGraph.cpp:179-189
        if (chr[e.Ind1] != chr[e.Ind2]) return true;
        if (pos[e.Ind2] - pos[e.Ind1] - len[e.Ind1] > concord_dist_pos && e.Ind2 - e.Ind1 > concord_dist_idx) return true;
        return e.Head1 != false || e.Head2 != true;
    };
    std::vector<EdgeRow> E((size_t)n_edges);
    for (int64_t i = 0; i < n_edges; i++) {
        if (ind1[i] < 0 || ind1[i] >= n_nodes || ind2[i] < 0 || ind2[i] >= n_nodes) return SQG_EINVAL;
        E[(size_t)i] = EdgeRow{ind1[i], ind2[i], head1[i] != 0, head2[i] != 0, weight[i], 0};
    }
    for (EdgeRow &e : E)  // DeMultiplyDisEdges: int / double, truncated back to int
        if (discordant(e) && discordant_ratio != 1) e.Weight = (int)e.Weight / discordant_ratio;
    // Node_NewChr (src/main.cpp:46-50)
    std::vector<std::pair<int, int>> where((size_t)n_nodes, std::make_pair(0, 0));
    for (int64_t c = 0; c < n_comp; c++)
        for (int64_t j = comp_off[c]; j < comp_off[c + 1]; j++) {
            const int64_t id = std::abs(comp_nodes[j]) - 1;
            if (id < 0 || id >= n_nodes) return SQG_EINVAL;
            where[(size_t)id] = std::make_pair((int)c, (int)(j - comp_off[c]));
        }
    auto comp_at = [&](const std::pair<int, int> &p) { return comp_nodes[comp_off[p.first] + p.second]; };
    BpMap exact, support;
    rows_to_map(exactbp_rows6, n_exactbp, exact);
    rows_to_map(support_rows6, n_support, support);

    std::sort(E.begin(), E.end(), [](EdgeRow a, EdgeRow b) { return a.Weight > b.Weight; });  // WriteIO.cpp:48
    std::ofstream output(path, std::ios::out);
    if (!output) return SQG_EINVAL;
    output << "# chrom1\tstart1\tend1\tchrom2\tstart2\tend2\tname\tscore\tstrand1\tstrand2\tnum_concordantfrag_bp1\tnum_concordantfrag_bp2\n";
    for (const EdgeRow &e : E) {
        const int a = e.Ind1, b = e.Ind2;
        const bool flag_chr = chr[a] == chr[b];
        const bool flag_ori = e.Head1 == false && e.Head2 == true;
        const bool flag_dist = pos[b] - pos[a] - len[a] <= concord_dist_pos || b - a <= concord_dist_idx;
        if (flag_chr && flag_ori && flag_dist) continue;
        const std::pair<int, int> p1 = where[(size_t)a], p2 = where[(size_t)b];
        bool consistent = false;  // the edge agrees with the ordering of its component (WriteIO.cpp:57-64)
        if (p1.first == p2.first && p1.second < p2.second && e.Head1 == (comp_at(p1) < 0) && e.Head2 == (comp_at(p2) > 0)) consistent = true;
        else if (p1.first == p2.first && p1.second > p2.second && e.Head2 == (comp_at(p2) < 0) && e.Head1 == (comp_at(p1) > 0)) consistent = true;
        if (!consistent) continue;
        if (chr[a] < 0 || chr[a] >= n_ref || chr[b] < 0 || chr[b] >= n_ref) return SQG_EINVAL;
        const auto itsup = support.find(e);
        if (itsup == support.end()) return SQG_ESTATE;  // the reference asserts (WriteIO.cpp:80)
        const auto itbp = exact.find(e);
        std::vector<std::pair<int, int>> BP;
        if (itbp == exact.end() || itbp->second.empty()) BP.emplace_back(e.Head1 ? pos[a] : pos[a] + len[a], e.Head2 ? pos[b] : pos[b] + len[b]);
        else BP = itbp->second;
        if (BP.size() != itsup->second.size()) return SQG_ESTATE;  // asserted at WriteIO.cpp:92
        for (size_t k = 0; k < BP.size(); k++) {
            output << ref_name[chr[a]] << '\t';
            if (e.Head1) output << BP[k].first << '\t' << (pos[a] + len[a]) << '\t';
            else output << pos[a] << '\t' << BP[k].first << '\t';
            output << ref_name[chr[b]] << '\t';
            if (e.Head2) output << BP[k].second << '\t' << (pos[b] + len[b]) << '\t';
            else output << pos[b] << '\t' << BP[k].second << '\t';
            output << ".\t" << e.Weight << "\t" << (e.Head1 ? "-\t" : "+\t") << (e.Head2 ? "-\t" : "+\t") << itsup->second[k].first << "\t" << itsup->second[k].second << std::endl;
        }
    }
    output.close();
    return output.fail() ? SQG_EINVAL : SQG_OK;
}

}  // extern "C"
