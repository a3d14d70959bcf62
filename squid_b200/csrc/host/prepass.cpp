#include "prepass.h"

#include <algorithm>
#include <cstdlib>

namespace sqh {
namespace {
struct B {  // one chimeric block with the fields the pre-pass reads
    int32_t chr, pos, rpos, mref, mread;
    bool rev, first;
};
struct R {
    std::vector<B> F, S;
    int32_t ft, st_;
    bool fl, sl, mf;
};
bool end_disc(const std::vector<B> &v) {  // ReadRec.cpp:178-209
    for (size_t i = 0; i + 1 < v.size(); i++) {
        if (v[i].chr != v[i + 1].chr || v[i].rev != v[i + 1].rev) return true;
        const bool a = v[i].pos < v[i + 1].pos, r = v[i].rpos < v[i + 1].rpos;
        if (!v[i].rev && a != r) return true;
        if (v[i].rev && a == r) return true;
    }
    return false;
}
bool pair_disc(const R &r) {  // ReadRec.cpp:211-228 with needcheck=true
    if (r.F.empty() || r.S.empty()) return false;
    if (end_disc(r.F) || end_disc(r.S)) return true;
    const B &ff = r.F.front(), &fb = r.F.back(), &sf = r.S.front(), &sb = r.S.back();
    if (ff.chr != sb.chr || ff.rev == sb.rev) return true;
    if (!ff.rev && ff.pos - ff.rpos > sb.pos - (r.st_ - sb.rpos - sb.mread)) return true;
    if (!sf.rev && sf.pos - sf.rpos > fb.pos - (r.ft - fb.rpos - fb.mread)) return true;
    return false;
}
sq::DiscBlock as_disc(const B &b) { return sq::DiscBlock{b.chr, b.pos, b.mref, b.rev ? 1 : 0}; }
}  // namespace

void chimeric_prepass(const sqg_chimeric &c, int32_t n_ref, int32_t read_len, ChimPrepass &out) {
    out = ChimPrepass();
    struct DB { sq::DiscBlock d; int32_t rpos, mread; bool first; };
    std::vector<DB> dis;
    std::vector<std::pair<int, int>> part((size_t)n_ref, std::make_pair(0, 0));  // resize()d then appended (:203-204)
    auto push_dis = [&](const B &b) { dis.push_back(DB{as_disc(b), b.rpos, b.mread, b.first}); };
    for (int64_t i = 0; i < c.n_reads; i++) {
        R r;
        const uint32_t o = c.read_off[i], e = c.read_off[i + 1], nf = c.n_first[i];
        for (uint32_t k = o; k < e; k++) {
            B b{c.blk_ref_id[k], c.blk_ref_pos[k], c.blk_read_pos[k], c.blk_match_ref[k], c.blk_match_read[k], c.blk_is_reverse[k] != 0, k - o < nf};
            (k - o < nf ? r.F : r.S).push_back(b);
        }
        r.ft = c.first_total_len[i]; r.st_ = c.second_total_len[i];
        r.fl = c.first_lowphred[i]; r.sl = c.second_lowphred[i]; r.mf = c.multi_filter[i];
        const bool single = (r.F.empty() || r.S.empty()) && !r.mf;
        if (end_disc(r.F) || end_disc(r.S) || single || pair_disc(r)) {  // :208-213
            for (const B &b : r.F) push_dis(b);
            for (const B &b : r.S) push_dis(b);
            continue;
        }
        bool fin = false, sin = false;
        for (int m = 0; m < 2; m++) {  // blocks of one mate more than 750 kb apart (:217-239)
            const std::vector<B> &v = m ? r.S : r.F;
            int prev = -1;
            for (int k = 0; k + 1 < (int)v.size(); k++)
                if (std::abs(v[k].pos - v[k + 1].pos) > 750000) {
                    if (prev != k) push_dis(v[k]);
                    push_dis(v[k + 1]);
                    prev = k + 1;
                    if (k + 1 == (int)v.size() - 1) (m ? sin : fin) = true;
                }
        }
        if (!r.F.empty() && !r.S.empty() && std::abs(r.F.back().pos - r.S.back().pos) > 750000) {  // :240-249
            if (!fin) { push_dis(r.F.back()); fin = true; }
            if (!sin) { push_dis(r.S.back()); sin = true; }
        }
        if (!fin && !sin) {  // soft-clipped ends of otherwise concordant chimeric reads (:250-259)
            if (!r.F.empty() && r.F.front().rpos > 15 && !r.fl)
                part.push_back({r.F[0].chr, r.F[0].rev ? r.F[0].pos + r.F[0].mref : r.F[0].pos});
            if (!r.F.empty() && r.ft - r.F.back().rpos - r.F.back().mread > 15 && !r.fl)
                part.push_back({r.F.back().chr, r.F.back().rev ? r.F.back().pos : r.F.back().pos + r.F.back().mref});
            if (!r.S.empty() && r.S.front().rpos > 15 && !r.sl)
                part.push_back({r.S[0].chr, r.S[0].rev ? r.S[0].pos + r.S[0].mref : r.S[0].pos});
            if (!r.S.empty() && r.st_ - r.S.back().rpos - r.S.back().mread > 15 && !r.sl) {
                // `!bamdiscordant.back().Same(SecondMate.back())` (:257); back() of an empty vector is UB in the
                // reference, we read it as "not the same"
                const B &sb = r.S.back();
                bool same = false;
                if (!dis.empty()) {
                    const DB &l = dis.back();
                    same = l.d.chr == sb.chr && l.d.pos == sb.pos && l.rpos == sb.rpos && l.mread == sb.mread && l.d.len == sb.mref &&
                           (l.d.rev != 0) == sb.rev && l.first == sb.first;
                }
                if (!same) part.push_back({sb.chr, sb.rev ? sb.pos : sb.pos + sb.mref});
            }
        }
    }
    std::sort(part.begin(), part.end(), [](std::pair<int, int> a, std::pair<int, int> b) { return a.first == b.first ? a.second < b.second : a.first < b.first; });
    // same unstable sort, same key, same sequence as :264 => same order among equal (RefID,RefPos)
    std::sort(dis.begin(), dis.end(), [](const DB &a, const DB &b) { return a.d.chr != b.d.chr ? a.d.chr < b.d.chr : a.d.pos < b.d.pos; });
    for (auto &p : part) { out.part_chr.push_back(p.first); out.part_pos.push_back(p.second); }
    out.disc.reserve(dis.size() + 1);
    for (const DB &d : dis) out.disc.push_back(d.d);
    const int32_t n = (int32_t)dis.size();
    out.disc.push_back(sq::DiscBlock{0, 0, 0, 0});
    for (int32_t s = 0; s < n;) {  // :341-348 chain while the next block starts within ReadLen of the running right end
        int32_t right = out.disc[s].pos + out.disc[s].len, e = s;
        for (; e < n && out.disc[e].chr == out.disc[s].chr && out.disc[e].pos < right + read_len; e++)
            right = std::max(right, out.disc[e].pos + out.disc[e].len);
        out.groups.push_back(sq::Group{s, e, out.disc[s].chr, right});
        s = e;
    }
}
}  // namespace sqh
