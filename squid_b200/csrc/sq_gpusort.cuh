// libstdc++'s std::sort permutation, computed on the GPU.
//
// The order in which equal (RefID,RefPos) discordant blocks leave the reference's `sort(bamdiscordant)`
// (SegmentGraph.cpp:264) is observable, so the chimeric pre-pass has to produce exactly the
// permutation of libstdc++'s introsort.  host/prepass.cpp does that on the CPU cores (sort_like_std); this is the same
// computation on the device, so that the host cores stay free (one process per GPU shares them):
//   * big ranges: one introsort level at a time over ALL active ranges at once.  Per range: median of three to the front
//     (std::__move_median_to_first), then the unguarded Hoare partition in closed form -- L = ascending positions whose element
//     is not < pivot, R = descending positions whose element is not > pivot, swap L[i] <-> R[i] for i < m = #{i : L[i] < R[i]},
//     cut = L[m] if that lies left of R[m-1] else R[m-1] (derivation and CPU twin: host/prepass.cpp partition_parallel).  L and R
//     come from one packed prefix sum over the whole array; ranges are found by binary search over the sorted range list.
//   * ranges of at most kLeafMax elements: one block each, staged in shared memory, where one lane runs the literal
//     std::__introsort_loop and the insertion pass of std::__final_insertion_sort (which never moves an element across a
//     partition boundary, so it can be run per range).
//   * a range whose depth budget 2*floor(log2 n) runs out would switch to heap sort in std::sort: reported (status 2), the
//     caller then runs the CPU twin.  (Never seen on alignment data.)
// Verified against std::sort on the host: tests/test_gpu_sort.py (random keys with many ties, sorted, reversed, organ pipe).
#ifndef SQ_GPUSORT_CUH
#define SQ_GPUSORT_CUH
#include <cub/cub.cuh>
#include <cstdint>

namespace sq {
namespace gsort {

constexpr int kLeafMax = 1024;   // ranges up to this size are finished by one block
constexpr int kSmall = 16;       // std::sort's _S_threshold

struct Range { uint32_t first, last; int32_t depth; };

__device__ __forceinline__ void swap_at(uint64_t *keys, uint32_t *idx, uint32_t a, uint32_t b) {
    const uint64_t k = keys[a]; keys[a] = keys[b]; keys[b] = k;
    const uint32_t v = idx[a]; idx[a] = idx[b]; idx[b] = v;
}
// std::__move_median_to_first(first, first+1, mid, last-1)
template <class K, class V>
__device__ __forceinline__ void median_to_first(K *keys, V *idx, uint32_t first, uint32_t last) {
    const uint32_t mid = first + (last - first) / 2, a = first + 1, c = last - 1;
    uint32_t w;
    if (keys[a] < keys[mid]) {
        if (keys[mid] < keys[c]) w = mid;
        else if (keys[a] < keys[c]) w = c;
        else w = a;
    } else if (keys[a] < keys[c]) w = a;
    else if (keys[mid] < keys[c]) w = c;
    else w = mid;
    const K k = keys[first]; keys[first] = keys[w]; keys[w] = k;
    const V v = idx[first]; idx[first] = idx[w]; idx[w] = v;
}

__global__ void k_median(uint64_t *keys, uint32_t *idx, const Range *rg, const int32_t *n_rg) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_rg) return;
    median_to_first(keys, idx, rg[r].first, rg[r].last);
}
// per element: its active range (or -1) and the packed flags (low word: belongs to L, high word: belongs to R)
__global__ void k_flags(const uint64_t *keys, uint32_t n, const Range *rg, const int32_t *n_rg, int32_t *rid, uint64_t *flags) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    uint64_t f = 0;
    int32_t r = -1;
    const int32_t nr = *n_rg;
    if (p < n && nr > 0) {
        int32_t lo = 0, hi = nr;  // last range with first <= p
        while (lo < hi) { const int32_t m = (lo + hi) >> 1; if (rg[m].first <= p) lo = m + 1; else hi = m; }
        if (lo > 0 && p < rg[lo - 1].last) {
            r = lo - 1;
            if (p > rg[r].first) {
                const uint64_t k = keys[p], piv = keys[rg[r].first];
                if (!(k < piv)) f |= 1ull;
                if (!(piv < k)) f |= 1ull << 32;
            }
        }
    }
    if (p < n) rid[p] = r;
    flags[p] = f;
}
__global__ void k_scatter(uint32_t n, const Range *rg, const int32_t *rid, const uint64_t *flags, const uint64_t *scan, uint32_t *Lpos, uint32_t *Rpos) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t r = rid[p];
    if (r < 0) return;
    const uint64_t f = flags[p];
    if (!f) return;
    const uint32_t base = rg[r].first + 1;
    const uint64_t s = scan[p], s0 = scan[base];
    if (f & 1ull) Lpos[base + ((uint32_t)s - (uint32_t)s0)] = p;
    if (f >> 32) Rpos[base + ((uint32_t)(s >> 32) - (uint32_t)(s0 >> 32))] = p;
}
// per range: m, the cut, nR
__global__ void k_cut(const Range *rg, const int32_t *n_rg, const uint64_t *scan, const uint32_t *Lpos, const uint32_t *Rpos, uint32_t *m_out, uint32_t *cut_out, uint32_t *nR_out) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= *n_rg) return;
    const uint32_t base = rg[r].first + 1, last = rg[r].last;
    const uint64_t s0 = scan[base], s1 = scan[last];
    const uint32_t nL = (uint32_t)s1 - (uint32_t)s0, nR = (uint32_t)(s1 >> 32) - (uint32_t)(s0 >> 32);
    const uint32_t *L = Lpos + base, *R = Rpos + base;  // both ascending; R[i] of the closed form is R[nR-1-i]
    uint32_t lo = 0, hi = nL < nR ? nL : nR;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (L[mid] < R[nR - 1 - mid]) lo = mid + 1; else hi = mid; }
    const uint32_t m = lo;
    uint32_t cut;
    if (m == 0) cut = L[0];
    else cut = (m < nL && L[m] < R[nR - m]) ? L[m] : R[nR - m];
    m_out[r] = m; cut_out[r] = cut; nR_out[r] = nR;
}
__global__ void k_swap(uint64_t *keys, uint32_t *idx, uint32_t n, const Range *rg, const int32_t *rid, const uint32_t *Lpos, const uint32_t *Rpos, const uint32_t *m_in, const uint32_t *nR_in) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t r = rid[p];
    if (r < 0) return;
    const uint32_t base = rg[r].first + 1;
    if (p < base) return;
    const uint32_t i = p - base;
    if (i >= m_in[r]) return;
    swap_at(keys, idx, Lpos[base + i], Rpos[base + nR_in[r] - 1 - i]);
}
// children of every range: still-big ones are counted (cnt) for the next level, the others go to the leaf list
__global__ void k_children_count(const Range *rg, const int32_t *n_rg, const uint32_t *cut_in, int32_t *cnt, int32_t cnt_len, Range *leaves, int32_t *n_leaves, int32_t *status) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t nr = *n_rg;
    if (r >= cnt_len) return;
    if (r >= nr) { cnt[r] = 0; return; }  // (the scan below runs over the whole array: the host does not know nr)
    const Range g = rg[r];
    const uint32_t cut = cut_in[r];
    const Range ch[2] = {Range{g.first, cut, g.depth - 1}, Range{cut, g.last, g.depth - 1}};
    int32_t c = 0;
    for (int k = 0; k < 2; k++) {
        const uint32_t sz = ch[k].last - ch[k].first;
        if (sz > (uint32_t)kLeafMax) { c++; if (ch[k].depth == 0) atomicMax(status, 2); }
        else if (sz > 1) leaves[atomicAdd(n_leaves, 1)] = ch[k];
    }
    cnt[r] = c;
}
__global__ void k_children_write(const Range *rg, const int32_t *n_rg, const uint32_t *cut_in, const int32_t *off, Range *next, int32_t *n_next) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t nr = *n_rg;
    if (r >= nr) return;
    const Range g = rg[r];
    const uint32_t cut = cut_in[r];
    int32_t o = off[r];
    if (cut - g.first > (uint32_t)kLeafMax) next[o++] = Range{g.first, cut, g.depth - 1};
    if (g.last - cut > (uint32_t)kLeafMax) next[o++] = Range{cut, g.last, g.depth - 1};
    if (r == nr - 1) *n_next = o;
}
// One block per leaf range: lane 0 runs the literal introsort loop and the insertion pass in shared memory.
__global__ void __launch_bounds__(32) k_leaves(uint64_t *keys, uint32_t *idx, const Range *leaves, const int32_t *n_leaves, int32_t *status) {
    __shared__ uint64_t sk[kLeafMax];
    __shared__ uint32_t sv[kLeafMax];
    for (int32_t q = blockIdx.x; q < *n_leaves; q += gridDim.x) {
        const Range g = leaves[q];
        const int n = (int)(g.last - g.first);
        for (int i = threadIdx.x; i < n; i += 32) { sk[i] = keys[g.first + i]; sv[i] = idx[g.first + i]; }
        __syncwarp();
        if (threadIdx.x == 0) {
            // std::__introsort_loop with an explicit stack: (first, last, depth); the loop keeps the left part, the right is pushed
            uint32_t st_f[64], st_l[64]; int32_t st_d[64];
            int sp = 0;
            st_f[0] = 0; st_l[0] = (uint32_t)n; st_d[0] = g.depth; sp = 1;
            while (sp > 0) {
                sp--;
                uint32_t first = st_f[sp], last = st_l[sp];
                int32_t depth = st_d[sp];
                while (last - first > (uint32_t)kSmall) {
                    if (depth == 0) { atomicMax(status, 2); break; }
                    --depth;
                    median_to_first(sk, sv, first, last);
                    uint32_t lo = first + 1, hi = last;
                    const uint64_t piv = sk[first];
                    for (;;) {
                        while (sk[lo] < piv) ++lo;
                        --hi;
                        while (piv < sk[hi]) --hi;
                        if (!(lo < hi)) break;
                        const uint64_t k = sk[lo]; sk[lo] = sk[hi]; sk[hi] = k;
                        const uint32_t v = sv[lo]; sv[lo] = sv[hi]; sv[hi] = v;
                        ++lo;
                    }
                    if (sp < 64) { st_f[sp] = lo; st_l[sp] = last; st_d[sp] = depth; sp++; } else { atomicMax(status, 3); }
                    last = lo;
                }
            }
            for (int i = 1; i < n; i++) {  // the insertion pass, per range
                const uint64_t k = sk[i]; const uint32_t v = sv[i];
                int j = i;
                while (j > 0 && k < sk[j - 1]) { sk[j] = sk[j - 1]; sv[j] = sv[j - 1]; --j; }
                sk[j] = k; sv[j] = v;
            }
        }
        __syncwarp();
        for (int i = threadIdx.x; i < n; i += 32) { keys[g.first + i] = sk[i]; idx[g.first + i] = sv[i]; }
        __syncwarp();
    }
}

// Device scratch of one sort (sizes in elements for n keys): see bytes_needed().
struct Scratch {
    int32_t *rid; uint64_t *flags, *scan; uint32_t *Lpos, *Rpos, *m, *cut, *nR; int32_t *cnt, *off;
    Range *ra, *rb, *leaves; int32_t *counters;  // counters: [0] n_a, [1] n_b, [2] n_leaves, [3] status
    void *cub_temp; size_t cub_bytes;
};
inline size_t max_ranges(size_t n) { return n / (kLeafMax / 2) + 8; }
inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t cub_bytes_needed(size_t n) {
    size_t a = 0, b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, a, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)(n + 1));
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t *)nullptr, (int32_t *)nullptr, (int)(max_ranges(n) + 1));
    return (a > b ? a : b) + 256;
}
inline size_t bytes_needed(size_t n) {
    const size_t R = max_ranges(n), Lv = n / 2 + 8;
    return align_up(4 * n) + 2 * align_up(8 * (n + 1)) + 2 * align_up(4 * (n + 1)) + 3 * align_up(4 * R) + 2 * align_up(4 * (R + 1)) + 2 * align_up(sizeof(Range) * R) +
           align_up(sizeof(Range) * Lv) + 256 + align_up(cub_bytes_needed(n));
}
inline Scratch carve(unsigned char *base, size_t n) {
    const size_t R = max_ranges(n), Lv = n / 2 + 8;
    Scratch s;
    auto take = [&](size_t bytes) { unsigned char *q = base; base += align_up(bytes); return (void *)q; };
    s.rid = (int32_t *)take(4 * n); s.flags = (uint64_t *)take(8 * (n + 1)); s.scan = (uint64_t *)take(8 * (n + 1));
    s.Lpos = (uint32_t *)take(4 * (n + 1)); s.Rpos = (uint32_t *)take(4 * (n + 1));
    s.m = (uint32_t *)take(4 * R); s.cut = (uint32_t *)take(4 * R); s.nR = (uint32_t *)take(4 * R);
    s.cnt = (int32_t *)take(4 * (R + 1)); s.off = (int32_t *)take(4 * (R + 1));
    s.ra = (Range *)take(sizeof(Range) * R); s.rb = (Range *)take(sizeof(Range) * R); s.leaves = (Range *)take(sizeof(Range) * Lv);
    s.counters = (int32_t *)take(256);
    s.cub_bytes = cub_bytes_needed(n); s.cub_temp = take(s.cub_bytes);
    return s;
}

// Sorts (keys, idx)[n] on `stream` into std::sort's permutation.  Returns 0, a CUDA error (< 0 as -(int)cudaError_t), or 2 / 3
// when std::sort would have left the quicksort path (depth budget exhausted): the arrays are then in an unspecified order and the
// caller must redo the sort on the CPU from its own copy.  `launches` is incremented per kernel.
inline int sort_like_std_device(uint64_t *keys, uint32_t *idx, size_t n, unsigned char *scratch, cudaStream_t stream, int64_t *launches) {
    if (n < 2) return 0;
    Scratch s = carve(scratch, n);
    int lg = 0;
    for (size_t v = n; v > 1; v >>= 1) lg++;
    int32_t h[4] = {0, 0, 0, 0};
    Range root{0, (uint32_t)n, 2 * lg};
    cudaError_t e;
#define GS_CK(x) do { e = (x); if (e != cudaSuccess) return -(int)e; } while (0)
    GS_CK(cudaMemsetAsync(s.counters, 0, 256, stream));
    if (n > (size_t)kLeafMax) { h[0] = 1; GS_CK(cudaMemcpyAsync(s.ra, &root, sizeof(Range), cudaMemcpyHostToDevice, stream)); }
    else { h[2] = 1; GS_CK(cudaMemcpyAsync(s.leaves, &root, sizeof(Range), cudaMemcpyHostToDevice, stream)); }
    GS_CK(cudaMemcpyAsync(s.counters, h, 16, cudaMemcpyHostToDevice, stream));
    Range *cur = s.ra, *nxt = s.rb;
    int32_t *n_cur = s.counters, *n_nxt = s.counters + 1;
    int32_t n_active = h[0];
    const unsigned eb = (unsigned)((n + 1 + 255) / 256);
    // The levels are enqueued back to back: every kernel reads the number of active ranges from device memory, so the host
    // only looks (status, ranges left) every few levels instead of paying a round trip per level.
    const int32_t R = (int32_t)max_ranges(n);
    const unsigned rb = (unsigned)((R + 1 + 127) / 128);
    const int kCheckEvery = 4;
    for (int level = 0; n_active > 0 && level < 4 * lg + 8; level++) {
        k_median<<<rb, 128, 0, stream>>>(keys, idx, cur, n_cur);
        k_flags<<<eb, 256, 0, stream>>>(keys, (uint32_t)n, cur, n_cur, s.rid, s.flags);
        size_t tb = s.cub_bytes;
        GS_CK(cub::DeviceScan::ExclusiveSum(s.cub_temp, tb, s.flags, s.scan, (int)(n + 1), stream));
        k_scatter<<<eb, 256, 0, stream>>>((uint32_t)n, cur, s.rid, s.flags, s.scan, s.Lpos, s.Rpos);
        k_cut<<<rb, 128, 0, stream>>>(cur, n_cur, s.scan, s.Lpos, s.Rpos, s.m, s.cut, s.nR);
        k_swap<<<eb, 256, 0, stream>>>(keys, idx, (uint32_t)n, cur, s.rid, s.Lpos, s.Rpos, s.m, s.nR);
        k_children_count<<<rb, 128, 0, stream>>>(cur, n_cur, s.cut, s.cnt, R + 1, s.leaves, s.counters + 2, s.counters + 3);
        tb = s.cub_bytes;
        GS_CK(cub::DeviceScan::ExclusiveSum(s.cub_temp, tb, s.cnt, s.off, R + 1, stream));
        GS_CK(cudaMemsetAsync(n_nxt, 0, 4, stream));
        k_children_write<<<rb, 128, 0, stream>>>(cur, n_cur, s.cut, s.off, nxt, n_nxt);
        if (launches) *launches += 9;
        Range *t = cur; cur = nxt; nxt = t;
        int32_t *tn = n_cur; n_cur = n_nxt; n_nxt = tn;
        if ((level + 1) % kCheckEvery == 0) {
            GS_CK(cudaMemcpyAsync(h, s.counters, 16, cudaMemcpyDeviceToHost, stream));
            GS_CK(cudaStreamSynchronize(stream));
            if (h[3]) return h[3];
            n_active = *(n_cur == s.counters ? &h[0] : &h[1]);
        }
    }
    GS_CK(cudaMemcpyAsync(h, s.counters, 16, cudaMemcpyDeviceToHost, stream));
    GS_CK(cudaStreamSynchronize(stream));
    if (h[3]) return h[3];
    n_active = *(n_cur == s.counters ? &h[0] : &h[1]);
    if (n_active > 0) return 2;
    if (h[2] > 0) {
        k_leaves<<<(unsigned)(h[2] < 148 * 32 ? h[2] : 148 * 32), 32, 0, stream>>>(keys, idx, s.leaves, s.counters + 2, s.counters + 3);
        if (launches) *launches += 1;
        GS_CK(cudaMemcpyAsync(h, s.counters, 16, cudaMemcpyDeviceToHost, stream));
        GS_CK(cudaStreamSynchronize(stream));
        if (h[3]) return h[3];
    }
    GS_CK(cudaGetLastError());
#undef GS_CK
    return 0;
}

}  // namespace gsort
}  // namespace sq
#endif
