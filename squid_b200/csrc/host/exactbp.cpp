// Host twin of SegmentGraph_t::ExactBreakpoint + CountTop (src/SegmentGraph.cpp:3019-3081, 51-102): the chimeric reads are
// located once more, on the FINAL graph (after the host filters and node compression the nodes no longer tile the genome, so
// this is the literal hinted scan of LocateRead, src/SegmentGraph.cpp:1207-1293, not the tiled closed form of the device
// path), every discordant split junction contributes a (bp1, bp2) pair to its edge, and each edge keeps at most five
// representative pairs.  Small: chimeric reads only, one pass.  Written from scratch against the cited lines.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <new>
#include <vector>

#include "squid_b200_host.h"

namespace {

struct Nodes {
    const int32_t *chr, *pos, *len;
    int64_t n;
};
struct BlkRef {  // one aligned block of a chimeric read, mutable in place (LocateRead trims)
    int32_t ref_id;
    int32_t *ref_pos, *read_pos, *match_ref, *match_read;
    bool rev;
};
struct EdgeKey {  // Edge_t's canonicalising constructor and operator< (src/BPEdge.h:31-52, 59-70)
    int32_t i1, i2;
    bool h1, h2;
    EdgeKey(int32_t a, bool ha, int32_t b, bool hb) {
        if (a > b) { i1 = b; h1 = hb; i2 = a; h2 = ha; } else { i1 = a; h1 = ha; i2 = b; h2 = hb; }
    }
    bool operator<(const EdgeKey &o) const {
        if (i1 != o.i1) return i1 < o.i1;
        if (i2 != o.i2) return i2 < o.i2;
        if (h1 != o.h1) return (int)h1 < (int)o.h1;
        if (h2 != o.h2) return (int)h2 < (int)o.h2;
        return false;
    }
};

inline bool fits(const Nodes &N, int64_t i, const BlkRef &b) {  // :1213
    const int thresh = 5;
    return N.chr[i] == b.ref_id && *b.ref_pos >= N.pos[i] - thresh && *b.ref_pos + *b.match_ref <= N.pos[i] + N.len[i] + thresh;
}

// LocateRead(int initialguess, ReadRec_t&): blocks of the first mate, then of the second; the cursor carries over from block to
// block and is reset to the hint when it has left the node vector (:1211-1212).
void locate_read(const Nodes &N, int64_t hint, std::vector<BlkRef> &blocks, std::vector<int64_t> &node_of) {
    node_of.assign(blocks.size(), 0);
    int64_t i = hint;
    for (size_t k = 0; k < blocks.size(); k++) {
        BlkRef &b = blocks[k];
        if (i < 0 || i >= N.n) i = hint;
        if (!fits(N, i, b)) {
            if (N.chr[i] < b.ref_id || (N.chr[i] == b.ref_id && N.pos[i] <= *b.ref_pos)) {
                for (; i < N.n && N.chr[i] <= b.ref_id; i++) if (fits(N, i, b)) break;
            } else {
                for (; i > -1 && N.chr[i] >= b.ref_id; i--) if (fits(N, i, b)) break;
            }
        }
        if (i < 0 || i >= N.n || N.chr[i] != b.ref_id) { node_of[k] = -1; continue; }
        node_of[k] = i;
        if (*b.ref_pos < N.pos[i]) {  // left overhang (:1229-1238)
            const int32_t d = N.pos[i] - *b.ref_pos;
            if (!b.rev) *b.read_pos += d;
            *b.match_ref -= d; *b.match_read -= d;
            *b.ref_pos = N.pos[i];
        }
        if (*b.ref_pos + *b.match_ref > N.pos[i] + N.len[i]) {  // right overhang (:1239-1247)
            const int32_t d = *b.ref_pos + *b.match_ref - N.pos[i] - N.len[i];
            if (b.rev) *b.read_pos += d;
            *b.match_ref -= d; *b.match_read -= d;
        }
    }
}

// CountTop (:51-102): representatives of the (bp1, bp2) pairs of one edge
void count_top(bool head1, bool head2, std::vector<std::pair<int, int>> &x) {
    std::sort(x.begin(), x.end(), [](std::pair<int, int> a, std::pair<int, int> b) { return a.first != b.first ? a.first < b.first : a.second < b.second; });
    std::vector<std::pair<int, int>> y = x;
    y.erase(std::unique(y.begin(), y.end()), y.end());
    std::vector<double> count(y.size(), 0);
    for (size_t i = 0; i < y.size(); i++)
        for (size_t j = 0; j < x.size(); j++) {
            if (y[i] == x[j]) count[i] += 1;
            else if (std::abs(y[i].first - x[j].first) + std::abs(y[i].second - x[j].second) < 10) count[i] += 0.5;
        }
    x.clear();
    while (x.size() < 5) {
        const size_t at = (size_t)(std::max_element(count.begin(), count.end()) - count.begin());
        if (!(count[at] > 3)) break;
        bool far = true;
        for (const auto &q : x) if (std::abs(q.first - y[at].first) + std::abs(q.second - y[at].second) < 50) far = false;
        if (far) x.push_back(y[at]);
        count[at] = 0;
    }
    if (x.empty()) {  // no pair seen more than three times: the extreme positions on the edge's sides
        int max1 = 0, max2 = 0, min1 = std::numeric_limits<int>::max(), min2 = std::numeric_limits<int>::max();
        for (const auto &q : y) { min1 = std::min(min1, q.first); max1 = std::max(max1, q.first); min2 = std::min(min2, q.second); max2 = std::max(max2, q.second); }
        x.emplace_back(head1 ? min1 : max1, head2 ? min2 : max2);
    }
}

}  // namespace

extern "C" {

void sqh_free(void *p) { free(p); }

int sqh_exact_breakpoint(const int32_t *node_chr, const int32_t *node_pos, const int32_t *node_len, int64_t n_nodes, sqg_chimeric *chim,
                         int32_t concord_dist_pos, int32_t concord_dist_idx, int32_t **rows6, int64_t *n_rows) {
    if (!node_chr || !node_pos || !node_len || n_nodes <= 0 || !chim || !rows6 || !n_rows) return SQG_EINVAL;
    *rows6 = nullptr; *n_rows = 0;
    const Nodes N{node_chr, node_pos, node_len, n_nodes};
    auto discordant = [&](const EdgeKey &e) {  // IsDiscordant(Edge_t), :179-189
        if (N.chr[e.i1] != N.chr[e.i2]) return true;
        if (N.pos[e.i2] - N.pos[e.i1] - N.len[e.i1] > concord_dist_pos && e.i2 - e.i1 > concord_dist_idx) return true;
        return e.h1 != false || e.h2 != true;
    };
    std::map<EdgeKey, std::vector<std::pair<int, int>>> bp;
    int64_t first_front = 0;
    std::vector<BlkRef> blocks;
    std::vector<int64_t> node_of;
    for (int64_t r = 0; r < chim->n_reads; r++) {
        const uint32_t o = chim->read_off[r], e = chim->read_off[r + 1];
        const uint32_t nf = chim->n_first[r], ns = (e - o) - nf;
        if (nf <= 1 && ns <= 1) continue;  // :3024
        blocks.clear();
        for (uint32_t k = o; k < e; k++)
            blocks.push_back(BlkRef{chim->blk_ref_id[k], &chim->blk_ref_pos[k], &chim->blk_read_pos[k], &chim->blk_match_ref[k], &chim->blk_match_read[k], chim->blk_is_reverse[k] != 0});
        locate_read(N, first_front, blocks, node_of);
        if (node_of[0] != -1) first_front = node_of[0];
        for (int mate = 0; mate < 2; mate++) {
            const uint32_t base = mate ? nf : 0, cnt = mate ? ns : nf;
            for (uint32_t k = 0; k + 1 < cnt; k++) {
                const int64_t i = node_of[base + k], j = node_of[base + k + 1];
                if (i == j || i == -1 || j == -1) continue;
                const BlkRef &a = blocks[base + k], &b = blocks[base + k + 1];
                const EdgeKey key((int32_t)i, a.rev, (int32_t)j, !b.rev);
                if (!discordant(key)) continue;
                int b1 = a.rev ? *a.ref_pos : *a.ref_pos + *a.match_ref;
                int b2 = b.rev ? *b.ref_pos + *b.match_ref : *b.ref_pos;
                const bool a_after_b = a.ref_id != b.ref_id ? a.ref_id > b.ref_id : *a.ref_pos > *b.ref_pos;  // SingleBamRec_t::operator>
                if (a_after_b) std::swap(b1, b2);
                bp[key].emplace_back(b1, b2);
            }
        }
    }
    int64_t total = 0;
    for (auto &kv : bp) { count_top(kv.first.h1, kv.first.h2, kv.second); total += (int64_t)kv.second.size(); }
    int32_t *out = (int32_t *)malloc(sizeof(int32_t) * 6 * (size_t)(total > 0 ? total : 1));
    if (!out) return SQG_ENOMEM;
    int64_t w = 0;
    for (const auto &kv : bp)
        for (const auto &q : kv.second) {
            int32_t *row = out + 6 * w++;
            row[0] = kv.first.i1; row[1] = kv.first.i2; row[2] = kv.first.h1; row[3] = kv.first.h2; row[4] = q.first; row[5] = q.second;
        }
    *rows6 = out; *n_rows = total;
    return SQG_OK;
}

}  // extern "C"
