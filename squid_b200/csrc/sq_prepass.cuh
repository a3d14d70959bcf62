// The chimeric pre-pass of BuildNode_STAR (SegmentGraph.cpp:196-264, 341-348) on the device: which blocks of Chimrecord are
// "discordant" (bamdiscordant, in push order), the soft-clip positions of the other reads (PartAlignPos), the sort of both and
// the chaining of the discordant blocks into groups.  host/prepass.cpp is the same computation on the cores (it still serves
// the shard planner and the CPU stepping harness); on the device it costs the host nothing, which is what limits several
// processes -- one per GPU -- on one host.
//   k_pre_count   one thread per read: how many blocks it pushes, and which block it pushed last
//   (scans)       push offsets; for every read the nearest earlier read that pushed (`bamdiscordant.back()`, :257)
//   k_pre_write   the (key, block) pairs in push order -- the input sequence of the reference's unstable std::sort, whose
//                 permutation sq_gpusort.cuh replays -- and the PartAlignPos entries (any order: sorted by their full value)
//   k_pre_gather  sorted block indices -> DiscBlock array (+ the zeroed element *cend() reads)
//   k_pre_opens   a block opens a group iff it is the first of its chromosome or starts at/after (maximum end over all earlier
//                 blocks of the chromosome) + ReadLen (:341-348; host/prepass.cpp explains why the prefix maximum is enough)
//   k_pre_groups  group table from the compacted opening indices
#ifndef SQ_PREPASS_CUH
#define SQ_PREPASS_CUH
#include "sq_seed.cuh"

namespace sq {

struct PreChim {  // device copy of sqg_chimeric
    int64_t n_reads;
    const uint32_t *read_off; const uint16_t *n_first;
    const int32_t *first_total, *second_total;
    const uint8_t *first_low, *second_low, *multi;
    const int32_t *chr, *pos, *rpos, *mref, *mread;
    const uint8_t *rev;
};

__device__ __forceinline__ bool pre_end_disc(const PreChim &c, uint32_t a, uint32_t b) {  // ReadRec.cpp:178-209 on blocks [a,b)
    for (uint32_t i = a; i + 1 < b; i++) {
        if (c.chr[i] != c.chr[i + 1] || (c.rev[i] != 0) != (c.rev[i + 1] != 0)) return true;
        const bool x = c.pos[i] < c.pos[i + 1], r = c.rpos[i] < c.rpos[i + 1];
        if (!c.rev[i] && x != r) return true;
        if (c.rev[i] && x == r) return true;
    }
    return false;
}
__device__ __forceinline__ bool pre_pair_disc(const PreChim &c, uint32_t f0, uint32_t f1, uint32_t s1, int32_t ft, int32_t st) {  // ReadRec.cpp:211-228
    if (f0 == f1 || f1 == s1) return false;
    const uint32_t ff = f0, fb = f1 - 1, sf = f1, sb = s1 - 1;
    if (c.chr[ff] != c.chr[sb] || (c.rev[ff] != 0) == (c.rev[sb] != 0)) return true;
    if (!c.rev[ff] && c.pos[ff] - c.rpos[ff] > c.pos[sb] - (st - c.rpos[sb] - c.mread[sb])) return true;
    if (!c.rev[sf] && c.pos[sf] - c.rpos[sf] > c.pos[fb] - (ft - c.rpos[fb] - c.mread[fb])) return true;
    return false;
}
__device__ __forceinline__ int32_t pre_absdiff(int32_t a, int32_t b) { const int64_t d = (int64_t)a - b; return (int32_t)(d < 0 ? -d : d); }

// The read loop (:206-262) for read i.  EMIT = false: counts only.  Returns the number of discordant pushes; *last_k = block of the
// last push (-1: none); pushes go to (keys, idx) from `at` on.  part entries are appended through `n_part`.
template <bool EMIT>
__device__ __forceinline__ int32_t pre_read(const PreChim &c, int64_t i, int32_t *last_k, uint64_t *keys, uint32_t *idx, int64_t at, int64_t back_k,
                                            uint64_t *part, unsigned long long *n_part) {
    const uint32_t o = c.read_off[i], e = c.read_off[i + 1], f0 = o, f1 = o + c.n_first[i], s1 = e;
    const int32_t ft = c.first_total[i], st = c.second_total[i];
    const bool fl = c.first_low[i] != 0, sl = c.second_low[i] != 0, mf = c.multi[i] != 0;
    const bool fempty = f0 == f1, sempty = f1 == s1;
    int32_t n = 0, lk = -1;
    auto push = [&](uint32_t k) {
        if (EMIT) { keys[at + n] = ((uint64_t)(uint32_t)c.chr[k] << 32) | (uint32_t)c.pos[k]; idx[at + n] = k; }
        n++; lk = (int32_t)k;
    };
    const bool ed = pre_end_disc(c, f0, f1) || pre_end_disc(c, f1, s1);
    if (ed || ((fempty || sempty) && !mf) || pre_pair_disc(c, f0, f1, s1, ft, st)) {  // :208-213
        for (uint32_t k = o; k < e; k++) push(k);
        *last_k = lk;
        return n;
    }
    bool fin = false, sin = false;
    for (int m = 0; m < 2; m++) {  // blocks of one mate more than 750 kb apart (:217-239)
        const uint32_t a = m ? f1 : f0, b = m ? s1 : f1;
        int64_t prev = -1;
        for (uint32_t k = a; k + 1 < b; k++)
            if (pre_absdiff(c.pos[k], c.pos[k + 1]) > 750000) {
                if (prev != (int64_t)k) push(k);
                push(k + 1);
                prev = k + 1;
                if (k + 1 == b - 1) { if (m) sin = true; else fin = true; }
            }
    }
    if (!fempty && !sempty && pre_absdiff(c.pos[f1 - 1], c.pos[s1 - 1]) > 750000) {  // :240-249
        if (!fin) { push(f1 - 1); fin = true; }
        if (!sin) { push(s1 - 1); sin = true; }
    }
    if (EMIT && !fin && !sin) {  // soft-clipped ends of otherwise concordant chimeric reads (:250-259)
        auto add = [&](int32_t chr, int32_t p) { part[atomicAdd(n_part, 1ull)] = ((uint64_t)(uint32_t)chr << 32) | (uint32_t)p; };
        if (!fempty && c.rpos[f0] > 15 && !fl) add(c.chr[f0], c.rev[f0] ? c.pos[f0] + c.mref[f0] : c.pos[f0]);
        if (!fempty) { const uint32_t b = f1 - 1; if (ft - c.rpos[b] - c.mread[b] > 15 && !fl) add(c.chr[b], c.rev[b] ? c.pos[b] : c.pos[b] + c.mref[b]); }
        if (!sempty && c.rpos[f1] > 15 && !sl) add(c.chr[f1], c.rev[f1] ? c.pos[f1] + c.mref[f1] : c.pos[f1]);
        if (!sempty) {
            const uint32_t b = s1 - 1;
            if (st - c.rpos[b] - c.mread[b] > 15 && !sl) {
                // `!bamdiscordant.back().Same(SecondMate.back())` (:257): back() = the last push so far, of this read or of the
                // nearest earlier read that pushed (an empty vector is UB in the reference: read as "not the same")
                const int64_t l = lk >= 0 ? (int64_t)lk : back_k;
                bool same = false;
                if (l >= 0) {
                    int64_t lo = 0, hi = c.n_reads;  // the read that owns block l: is l one of its FirstRead blocks?
                    while (lo < hi) { const int64_t m2 = (lo + hi) >> 1; if (c.read_off[m2 + 1] <= (uint32_t)l) lo = m2 + 1; else hi = m2; }
                    const bool l_first = (uint32_t)l - c.read_off[lo] < c.n_first[lo];
                    same = c.chr[l] == c.chr[b] && c.pos[l] == c.pos[b] && c.rpos[l] == c.rpos[b] && c.mread[l] == c.mread[b] && c.mref[l] == c.mref[b] &&
                           (c.rev[l] != 0) == (c.rev[b] != 0) && !l_first;  // SingleBamRec_t::Same(): every field incl. IsFirstRead
                }
                if (!same) add(c.chr[b], c.rev[b] ? c.pos[b] : c.pos[b] + c.mref[b]);
            }
        }
    }
    *last_k = lk;
    return n;
}

__global__ void k_pre_count(PreChim c, int32_t *n_dis, int32_t *pusher, int32_t *last_k) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n_reads) return;
    int32_t lk;
    const int32_t n = pre_read<false>(c, i, &lk, nullptr, nullptr, 0, -1, nullptr, nullptr);
    n_dis[i] = n;
    last_k[i] = lk;
    pusher[i] = n > 0 ? (int32_t)i : -1;  // (after the exclusive max scan: the nearest earlier read that pushed)
}
__global__ void k_pre_write(PreChim c, const int64_t *off, const int32_t *prev_pusher, const int32_t *last_k, uint64_t *keys, uint32_t *idx, uint64_t *part, unsigned long long *n_part) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n_reads) return;
    const int32_t j = prev_pusher[i];
    int32_t lk;
    pre_read<true>(c, i, &lk, keys, idx, off[i], j >= 0 ? (int64_t)last_k[j] : -1, part, n_part);
}
__global__ void k_pre_split_part(const uint64_t *part, int64_t n, int32_t *pchr, int32_t *ppos) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) { pchr[i] = (int32_t)(part[i] >> 32); ppos[i] = (int32_t)(uint32_t)part[i]; }
}
__global__ void k_pre_gather(PreChim c, const uint32_t *idx, int64_t n, DiscBlock *D, uint64_t *endkey) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e > n) return;
    if (e == n) { D[e] = DiscBlock{0, 0, 0, 0}; return; }
    const uint32_t k = idx[e];
    const DiscBlock d{c.chr[k], c.pos[k], c.mref[k], c.rev[k] ? 1 : 0};
    D[e] = d;
    endkey[e] = ((uint64_t)(uint32_t)d.chr << 32) | (uint32_t)(d.pos + d.len);  // running maximum of (chr, end): chromosomes ascend
}
// excl[e] = maximum (chr, end) key over the blocks before e (0 for e = 0)
__global__ void k_pre_opens(const DiscBlock *D, const uint64_t *excl, int64_t n, int32_t read_len, uint8_t *opens) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const DiscBlock d = D[e];
    const uint64_t p = excl[e];
    opens[e] = (e == 0 || (int32_t)(p >> 32) != d.chr || d.pos >= (int32_t)(uint32_t)p + read_len) ? 1 : 0;
}
struct PreOpenOp {
    const uint8_t *opens;
    __device__ bool operator()(int32_t e) const { return opens[e] != 0; }
};
// group g = blocks [open[g], open[g+1]) (the last one up to n); its right end = the running maximum when the next group opens
__global__ void k_pre_groups(const DiscBlock *D, const int32_t *open, const int32_t *n_groups, const uint64_t *excl, const uint64_t *incl, int64_t n, Group *G) {
    const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t nG = *n_groups;
    if (g >= nG) return;
    const int32_t ds = open[g], de = g + 1 < nG ? open[g + 1] : (int32_t)n;
    const uint64_t r = g + 1 < nG ? excl[de] : incl[n - 1];
    G[g] = Group{ds, de, D[ds].chr, (int32_t)(uint32_t)r};
}

}  // namespace sq
#endif
