// Host twin of SQUID's read model: CIGAR -> aligned blocks, discordance predicates.
// Follows the behaviour of src/ReadRec.cpp (cited per function); written from scratch.
#include "readrec.h"

#include <algorithm>

namespace sqh {

namespace {
enum : uint32_t { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
inline uint32_t op_of(uint32_t c) { return c & 15u; }
inline int len_of(uint32_t c) { return (int)(c >> 4); }
}  // namespace

int32_t Alignment::end_pos() const {
    int32_t e = pos;
    for (uint32_t i = 0; i < n_cigar; i++) {
        uint32_t op = op_of(cigar[i]);
        if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) e += len_of(cigar[i]);
    }
    return e;
}

// src/ReadRec.cpp:10-88.  Quirks kept on purpose:
//  * total_len counts M,S,H,I,=,X;
//  * a block opens at M or = and swallows every op up to the next S, H or N (D adds to the
//    reference span only, I to the read span only, P and X to both);
//  * I, D, X, P met outside a block are skipped without advancing anything;
//  * the poly-A/T test reads the bases at [read_pos - hard_clip, +span) and rejects >= 75 %;
//  * reverse-strand blocks get read_pos = total_len - read_pos - span.
void decode_alignment(const Alignment &a, const HostConfig &cfg, Decoded &out) {
    out.blocks.clear();
    int total = 0;
    int32_t end = a.pos;
    for (uint32_t i = 0; i < a.n_cigar; i++) {
        uint32_t op = op_of(a.cigar[i]);
        if (op == OP_M || op == OP_S || op == OP_H || op == OP_I || op == OP_EQ || op == OP_X) total += len_of(a.cigar[i]);
        if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) end += len_of(a.cigar[i]);
    }
    out.total_len = total;
    out.end_pos = end;

    const int thr = (cfg.phred33 ? 33 : 64) + cfg.min_phred;
    int best = 0;
    if (a.qual) {
        int run = 0;
        for (uint32_t i = 0; i < a.l_seq; i++) {
            run = ((int)(signed char)a.qual[i] < (int)(signed char)(char)thr) ? run + 1 : 0;
            if (run > best) best = run;
        }
    } else {
        // synthesised qualities: `synth_lowrun` times '#', then 'I'
        uint32_t lseq = 0;
        for (uint32_t i = 0; i < a.n_cigar; i++) {
            uint32_t op = op_of(a.cigar[i]);
            if (op == OP_M || op == OP_I || op == OP_S || op == OP_EQ || op == OP_X) lseq += len_of(a.cigar[i]);
        }
        const int lowc = '#', highc = 'I', t = (int)(signed char)(char)thr;
        if (highc < t) best = (int)lseq;
        else if (lowc < t) best = (int)std::min<uint32_t>(a.synth_lowrun, lseq);
    }
    out.lowphred_run = best;

    int read_pos = 0, ref_pos = a.pos, hard = 0, blk_index = 0;
    const bool rev = a.is_reverse(), first = a.is_first();
    for (uint32_t i = 0; i < a.n_cigar; i++) {
        const uint32_t op = op_of(a.cigar[i]);
        if (op == OP_S || op == OP_H) {
            read_pos += len_of(a.cigar[i]);
            if (op == OP_H) hard += len_of(a.cigar[i]);
        } else if (op == OP_M || op == OP_EQ) {
            int span_read = 0, span_ref = 0;
            uint32_t j = i;
            for (; j < a.n_cigar; j++) {
                const uint32_t o = op_of(a.cigar[j]);
                if (o == OP_S || o == OP_H || o == OP_N) break;
                if (o != OP_D) span_read += len_of(a.cigar[j]);
                if (o != OP_I) span_ref += len_of(a.cigar[j]);
            }
            int n_a = 0, n_t = 0;
            if (a.seq) {
                for (int k = read_pos - hard; k < read_pos + span_read - hard; k++) {
                    if (k < 0 || k >= (int)a.l_seq) continue;  // the reference asserts instead
                    const char c = a.seq[k];
                    if (c == 'a' || c == 'A') n_a++;
                    else if (c == 't' || c == 'T') n_t++;
                }
            } else if (blk_index < 4) {
                if (a.synth_polya >> (4 + blk_index) & 1) n_t = span_read;       // painted last => wins
                else if (a.synth_polya >> blk_index & 1) n_a = span_read;
            }
            // 1.0*n/span < 0.75  <=>  4n < 3span for the integer ranges that occur
            if (4 * n_a < 3 * span_read && 4 * n_t < 3 * span_read) {
                Block b;
                b.ref_id = a.ref_id; b.ref_pos = ref_pos; b.read_pos = rev ? total - read_pos - span_read : read_pos;
                b.match_ref = span_ref; b.match_read = span_read; b.mapq = a.mapq; b.is_reverse = rev; b.is_first = first;
                out.blocks.push_back(b);
            }
            read_pos += span_read;
            ref_pos += span_ref;
            blk_index++;
            i = j - 1;
        } else if (op == OP_N) {
            ref_pos += len_of(a.cigar[i]);
        }
    }
}

void Read::sort_by_read_pos() {
    auto by_read_pos = [](const Block &x, const Block &y) { return x.read_pos < y.read_pos; };
    std::sort(first.begin(), first.end(), by_read_pos);
    std::sort(second.begin(), second.end(), by_read_pos);
}

bool Read::single_anchored() const { return (first.empty() || second.empty()) && !multi_filter; }

static bool mate_is_discordant(const std::vector<Block> &v) {
    for (size_t i = 0; i + 1 < v.size(); i++) {
        const Block &x = v[i], &y = v[i + 1];
        if (x.ref_id != y.ref_id || x.is_reverse != y.is_reverse) return true;
        const bool ref_fwd = x.ref_pos < y.ref_pos, read_fwd = x.read_pos < y.read_pos;
        if (!x.is_reverse && ref_fwd != read_fwd) return true;
        if (x.is_reverse && ref_fwd == read_fwd) return true;
    }
    return false;
}
bool Read::end_discordant(bool first_mate) const { return mate_is_discordant(first_mate ? first : second); }

bool Read::pair_discordant(bool check_ends) const {
    if (first.empty() || second.empty()) return false;
    if (check_ends && (end_discordant(true) || end_discordant(false))) return true;
    const Block &ff = first.front(), &fb = first.back(), &sf = second.front(), &sb = second.back();
    if (ff.ref_id != sb.ref_id || ff.is_reverse == sb.is_reverse) return true;
    if (!ff.is_reverse && ff.ref_pos - ff.read_pos > sb.ref_pos - (second_total - sb.read_pos - sb.match_read)) return true;
    if (!sf.is_reverse && sf.ref_pos - sf.read_pos > fb.ref_pos - (first_total - fb.read_pos - fb.match_read)) return true;
    return false;
}

static bool lists_match(const std::vector<Block> &x, const std::vector<Block> &y) {
    if (x.size() != y.size()) return false;
    for (size_t i = 0; i < x.size(); i++)
        if (x[i].ref_id != y[i].ref_id || x[i].ref_pos != y[i].ref_pos || x[i].match_ref != y[i].match_ref) return false;
    return true;
}
bool Read::equal(const Read &a, const Read &b) {
    return (lists_match(a.first, b.first) && lists_match(a.second, b.second)) ||
           (lists_match(a.first, b.second) && lists_match(a.second, b.first));
}

bool Read::front_smaller(const Read &a, const Read &b) {
    const Block *x = nullptr, *y = nullptr;
    if (!a.first.empty() && !b.first.empty()) { x = &a.first.front(); y = &b.first.front(); }
    else if (!a.second.empty() && !b.second.empty()) { x = &a.second.front(); y = &b.second.front(); }
    else if (!a.first.empty() && !b.second.empty()) { x = &a.first.front(); y = &b.second.front(); }
    else if (!a.second.empty() && !b.first.empty()) { x = &a.second.front(); y = &b.first.front(); }
    else return false;
    return x->ref_id != y->ref_id ? x->ref_id < y->ref_id : x->ref_pos < y->ref_pos;
}

}  // namespace sqh
