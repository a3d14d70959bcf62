// SegmentGraph_t::ConnectedComponent (SegmentGraph.cpp:2986-3003, DFS :2911-2935) on the device.
// The reference restarts its node scan for every component (O(#components x N)) and labels each component, in the order of its
// smallest node index, by a DFS over Head/TailEdges.  Which node a DFS visits first does not matter for the labels: Label[i] =
// rank of i's component among all components ordered by their smallest node.  That is a union-find whose links always point to
// the smaller index -- the root of every tree is the component's smallest node -- followed by a prefix count of the roots:
//   k_cc_init      parent[i] = i
//   k_cc_hook      one thread per edge: find both roots (path halving), link the larger root under the smaller with a CAS, retry
//                  when another thread got there first (lock-free; every link strictly decreases the parent index, so no cycle)
//   k_cc_flatten   parent[i] = root(i);  is_root[i] = parent[i] == i
//   (scan)         rank of every root among the roots, exclusive prefix sum
//   k_cc_label     Label[i] = rank[parent[i]]
// Self-loops (Ind1 == Ind2) and duplicate edges change nothing, as in the DFS.
#ifndef SQ_CC_CUH
#define SQ_CC_CUH
#include <cstdint>

namespace sq {

__device__ __forceinline__ int32_t cc_find(int32_t *parent, int32_t x) {
    for (;;) {
        const int32_t p = ((volatile int32_t *)parent)[x];
        if (p == x) return x;
        const int32_t g = ((volatile int32_t *)parent)[p];
        if (g != p) parent[x] = g;  // path halving (a benign race: any ancestor is a valid parent)
        x = p;
    }
}
__global__ void k_cc_init(int32_t *parent, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (int32_t)i;
}
__global__ void k_cc_hook(int32_t *parent, const int32_t *ind1, const int32_t *ind2, int64_t n_edges, int64_t n_nodes, int32_t *bad) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    int32_t u = ind1[e], v = ind2[e];
    if (u < 0 || v < 0 || u >= n_nodes || v >= n_nodes) { atomicOr(bad, 1); return; }
    for (;;) {
        u = cc_find(parent, u); v = cc_find(parent, v);
        if (u == v) return;
        const int32_t hi = u > v ? u : v, lo = u > v ? v : u;
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;
        // someone linked `hi` meanwhile: look again from where we are
    }
}
__global__ void k_cc_flatten(int32_t *parent, int64_t n, int32_t *is_root) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t r = (int32_t)i;
    while (parent[r] != r) r = parent[r];  // links are final here: plain walk
    is_root[i] = r == (int32_t)i ? 1 : 0;
    __syncwarp();
    parent[i] = r;  // (racing readers above see either an ancestor or the root: both lead to the root)
}
__global__ void k_cc_label(const int32_t *parent, const int32_t *rank, int64_t n, int32_t *label) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) label[i] = rank[parent[i]];
}

}  // namespace sq
#endif
