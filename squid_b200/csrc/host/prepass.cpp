#include "prepass.h"

#include <algorithm>
#include <cstdlib>
#include <thread>
#include <chrono>
#include <cstdio>

namespace sqh {
namespace {
// one chimeric read as index ranges over the flat block arrays of sqg_chimeric: FirstRead = [f0,f1), SecondMate = [f1,s1)
struct View {
    const sqg_chimeric &c;
    uint32_t f0, f1, s1;
    int32_t chr(uint32_t k) const { return c.blk_ref_id[k]; }
    int32_t pos(uint32_t k) const { return c.blk_ref_pos[k]; }
    int32_t rpos(uint32_t k) const { return c.blk_read_pos[k]; }
    int32_t mref(uint32_t k) const { return c.blk_match_ref[k]; }
    int32_t mread(uint32_t k) const { return c.blk_match_read[k]; }
    bool rev(uint32_t k) const { return c.blk_is_reverse[k] != 0; }
};
bool end_disc(const View &v, uint32_t a, uint32_t b) {  // ReadRec.cpp:178-209 on blocks [a,b)
    for (uint32_t i = a; i + 1 < b; i++) {
        if (v.chr(i) != v.chr(i + 1) || v.rev(i) != v.rev(i + 1)) return true;
        const bool x = v.pos(i) < v.pos(i + 1), r = v.rpos(i) < v.rpos(i + 1);
        if (!v.rev(i) && x != r) return true;
        if (v.rev(i) && x == r) return true;
    }
    return false;
}
bool pair_disc(const View &v, int32_t ft, int32_t st_) {  // ReadRec.cpp:211-228 with needcheck=true
    if (v.f0 == v.f1 || v.f1 == v.s1) return false;
    if (end_disc(v, v.f0, v.f1) || end_disc(v, v.f1, v.s1)) return true;
    const uint32_t ff = v.f0, fb = v.f1 - 1, sf = v.f1, sb = v.s1 - 1;
    if (v.chr(ff) != v.chr(sb) || v.rev(ff) == v.rev(sb)) return true;
    if (!v.rev(ff) && v.pos(ff) - v.rpos(ff) > v.pos(sb) - (st_ - v.rpos(sb) - v.mread(sb))) return true;
    if (!v.rev(sf) && v.pos(sf) - v.rpos(sf) > v.pos(fb) - (ft - v.rpos(fb) - v.mread(fb))) return true;
    return false;
}

// ---- std::sort, run on several threads -----------------------------------------------------------------------------
// The order in which equal (RefID,RefPos) blocks leave the reference's `sort(bamdiscordant)` (SegmentGraph.cpp:264) is
// observable, so the pre-pass must produce exactly libstdc++'s std::sort permutation.  std::sort
// is introsort: quicksort partitions (median of first+1 / middle / last-1 moved to the front, unguarded Hoare partition
// around it, recursion on the right part, loop on the left) down to ranges of 16, heap sort when the depth budget
// 2*floor(log2 n) runs out, and a final insertion sort.  After a partition the two parts never interact again, so they can
// be sorted by different threads without changing a single comparison; the final insertion pass is stable and never
// moves an element across a partition boundary.  tests/test_cpu_host_twin.py checks this routine against std::sort.
struct SortKey { uint64_t key; uint32_t k; };
inline bool sk_lt(const SortKey &x, const SortKey &y) { return x.key < y.key; }
void sort_loop(SortKey *first, SortKey *last, int depth, int fanout) {
    std::vector<std::thread> kids;
    while (last - first > 16) {
        if (depth == 0) { std::partial_sort(first, last, last, sk_lt); break; }
        --depth;
        SortKey *mid = first + (last - first) / 2, *a = first + 1, *c = last - 1;
        // median of (*a, *mid, *c) to *first
        if (sk_lt(*a, *mid)) {
            if (sk_lt(*mid, *c)) std::swap(*first, *mid);
            else if (sk_lt(*a, *c)) std::swap(*first, *c);
            else std::swap(*first, *a);
        } else if (sk_lt(*a, *c)) std::swap(*first, *a);
        else if (sk_lt(*mid, *c)) std::swap(*first, *c);
        else std::swap(*first, *mid);
        // unguarded partition of [first+1, last) around *first
        SortKey *lo = first + 1, *hi = last;
        for (;;) {
            while (sk_lt(*lo, *first)) ++lo;
            --hi;
            while (sk_lt(*first, *hi)) --hi;
            if (!(lo < hi)) break;
            std::swap(*lo, *hi);
            ++lo;
        }
        SortKey *cut = lo;
        if (fanout > 0 && last - cut > 4096) {
            --fanout;
            kids.emplace_back(sort_loop, cut, last, depth, fanout);
        } else sort_loop(cut, last, depth, 0);
        last = cut;
    }
    for (std::thread &t : kids) t.join();
}
void sort_like_std(SortKey *first, SortKey *last, int fanout) {
    if (first == last) return;
    int lg = 0;
    for (size_t n = (size_t)(last - first); n > 1; n >>= 1) lg++;
    sort_loop(first, last, 2 * lg, fanout);
    for (SortKey *i = first + 1; i < last; ++i) {  // final insertion sort
        const SortKey v = *i;
        SortKey *j = i;
        while (j > first && sk_lt(v, *(j - 1))) { *j = *(j - 1); --j; }
        *j = v;
    }
}
}  // namespace

// test hook: does sort_like_std reproduce std::sort's permutation (payload included) on n keys drawn from [0,range)?
extern "C" int sqh_selftest_sort(int64_t n, uint64_t seed, uint64_t range, int pattern, int fanout) {
    std::vector<SortKey> a((size_t)n), b;
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    for (int64_t i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        uint64_t v = range ? x % range : 0;
        if (pattern == 1) v = (uint64_t)i / 3;                 // sorted with ties
        else if (pattern == 2) v = (uint64_t)(n - i) / 3;      // reversed with ties
        else if (pattern == 3) v = (uint64_t)std::min(i, n - 1 - i);  // organ pipe
        a[(size_t)i] = SortKey{v, (uint32_t)i};
    }
    b = a;
    std::sort(a.begin(), a.end(), sk_lt);
    sort_like_std(b.data(), b.data() + b.size(), fanout);
    for (size_t i = 0; i < a.size(); i++) if (a[i].key != b[i].key || a[i].k != b[i].k) return 0;
    return 1;
}

void chimeric_prepass(const sqg_chimeric &c, int32_t n_ref, int32_t read_len, ChimPrepass &out) {
    const bool timing = getenv("SQH_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *w) { if (timing) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[prepass] %s %.1f ms\n", w, 1e3 * std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    out = ChimPrepass();
    typedef SortKey DB;  // key = (RefID,RefPos) packed, k = block index
    std::vector<DB> dis;
    dis.reserve((size_t)c.n_blk);
    std::vector<std::pair<int, int>> part((size_t)n_ref, std::make_pair(0, 0));  // resize()d then appended (:203-204)
    auto push_dis = [&](uint32_t k) { dis.push_back(DB{((uint64_t)(uint32_t)c.blk_ref_id[k] << 32) | (uint32_t)c.blk_ref_pos[k], k}); };
    for (int64_t i = 0; i < c.n_reads; i++) {
        const uint32_t o = c.read_off[i], e = c.read_off[i + 1], nf = c.n_first[i];
        const View v{c, o, o + nf, e};
        const int32_t ft = c.first_total_len[i], st_ = c.second_total_len[i];
        const bool fl = c.first_lowphred[i], sl = c.second_lowphred[i], mf = c.multi_filter[i];
        const bool fempty = v.f0 == v.f1, sempty = v.f1 == v.s1;
        const bool single = (fempty || sempty) && !mf;
        if (end_disc(v, v.f0, v.f1) || end_disc(v, v.f1, v.s1) || single || pair_disc(v, ft, st_)) {  // :208-213
            for (uint32_t k = o; k < e; k++) push_dis(k);
            continue;
        }
        bool fin = false, sin = false;
        for (int m = 0; m < 2; m++) {  // blocks of one mate more than 750 kb apart (:217-239)
            const uint32_t a = m ? v.f1 : v.f0, b = m ? v.s1 : v.f1;
            int64_t prev = -1;
            for (uint32_t k = a; k + 1 < b; k++)
                if (std::abs(v.pos(k) - v.pos(k + 1)) > 750000) {
                    if (prev != (int64_t)k) push_dis(k);
                    push_dis(k + 1);
                    prev = k + 1;
                    if (k + 1 == b - 1) (m ? sin : fin) = true;
                }
        }
        if (!fempty && !sempty && std::abs(v.pos(v.f1 - 1) - v.pos(v.s1 - 1)) > 750000) {  // :240-249
            if (!fin) { push_dis(v.f1 - 1); fin = true; }
            if (!sin) { push_dis(v.s1 - 1); sin = true; }
        }
        if (!fin && !sin) {  // soft-clipped ends of otherwise concordant chimeric reads (:250-259)
            if (!fempty && v.rpos(v.f0) > 15 && !fl) part.push_back({v.chr(v.f0), v.rev(v.f0) ? v.pos(v.f0) + v.mref(v.f0) : v.pos(v.f0)});
            if (!fempty) { const uint32_t b = v.f1 - 1; if (ft - v.rpos(b) - v.mread(b) > 15 && !fl) part.push_back({v.chr(b), v.rev(b) ? v.pos(b) : v.pos(b) + v.mref(b)}); }
            if (!sempty && v.rpos(v.f1) > 15 && !sl) part.push_back({v.chr(v.f1), v.rev(v.f1) ? v.pos(v.f1) + v.mref(v.f1) : v.pos(v.f1)});
            if (!sempty) {
                const uint32_t b = v.s1 - 1;
                if (st_ - v.rpos(b) - v.mread(b) > 15 && !sl) {
                    // `!bamdiscordant.back().Same(SecondMate.back())` (:257); back() of an empty vector is UB in the
                    // reference, we read it as "not the same".  Same() compares every field incl. IsFirstRead.
                    bool same = false;
                    if (!dis.empty()) {
                        const uint32_t l = dis.back().k;
                        // which read does block l belong to? only its IsFirstRead matters: l is a SecondMate block iff it
                        // lies at/after the first SecondMate block of its own read; find that read by binary search
                        const uint32_t *ro = c.read_off;
                        int64_t lo = 0, hi = c.n_reads;
                        while (lo < hi) { const int64_t m2 = (lo + hi) >> 1; if (ro[m2 + 1] <= l) lo = m2 + 1; else hi = m2; }
                        const bool l_first = l - ro[lo] < c.n_first[lo];
                        same = v.chr(l) == v.chr(b) && v.pos(l) == v.pos(b) && v.rpos(l) == v.rpos(b) && v.mread(l) == v.mread(b) &&
                               v.mref(l) == v.mref(b) && v.rev(l) == v.rev(b) && l_first == false;
                    }
                    if (!same) part.push_back({v.chr(b), v.rev(b) ? v.pos(b) : v.pos(b) + v.mref(b)});
                }
            }
        }
    }
    lap("reads");
    std::sort(part.begin(), part.end(), [](std::pair<int, int> a, std::pair<int, int> b) { return a.first == b.first ? a.second < b.second : a.first < b.first; });
    // Same unstable std::sort, same ordering relation, same input sequence as :264 => the same permutation, including
    // the order among equal (RefID,RefPos) which the sub-cluster walk observes.  The packed key
    // compares exactly like operator< of SingleBamRec_t (RefID, RefPos are non-negative here).
    lap("part sort");
    sort_like_std(dis.data(), dis.data() + dis.size(), 6);
    lap("disc sort");
    for (auto &p : part) { out.part_chr.push_back(p.first); out.part_pos.push_back(p.second); }
    out.disc.resize(dis.size() + 1);
    {   // gather the sorted blocks (random access into the caller's arrays): split across a few threads
        const size_t nd = dis.size();
        const unsigned nt = nd > (1u << 16) ? 8 : 1;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                for (size_t i = nd * t / nt; i < nd * (t + 1) / nt; i++) {
                    const uint32_t k = dis[i].k;
                    out.disc[i] = sq::DiscBlock{c.blk_ref_id[k], c.blk_ref_pos[k], c.blk_match_ref[k], c.blk_is_reverse[k] ? 1 : 0};
                }
            });
        for (auto &x : th) x.join();
    }
    const int32_t n = (int32_t)dis.size();
    out.disc[dis.size()] = sq::DiscBlock{0, 0, 0, 0};  // what *cend() reads
    for (int32_t s = 0; s < n;) {  // :341-348 chain while the next block starts within ReadLen of the running right end
        int32_t right = out.disc[s].pos + out.disc[s].len, e = s;
        for (; e < n && out.disc[e].chr == out.disc[s].chr && out.disc[e].pos < right + read_len; e++)
            right = std::max(right, out.disc[e].pos + out.disc[e].len);
        out.groups.push_back(sq::Group{s, e, out.disc[s].chr, right});
        s = e;
    }
    lap("groups");
}
}  // namespace sqh
