// TEST INFRASTRUCTURE (oracle side): the sliver of Boost.Graph the reference's MincutRecursion
// touches (SegmentGraph.cpp:3316-3325) — off the hot path, present only so the unmodified
// reference sources compile.  stoer_wagner_min_cut is a plain O(V^3) Stoer-Wagner.
#ifndef SHIM_BOOST_GRAPH_HPP
#define SHIM_BOOST_GRAPH_HPP
#include <climits>
#include <cstddef>
#include <utility>
#include <vector>
namespace boost {
struct vecS {};
struct undirectedS {};
struct no_property {};
struct edge_weight_t {};
struct vertex_index_t {};
static const edge_weight_t edge_weight = edge_weight_t();
static const vertex_index_t vertex_index = vertex_index_t();
template <class Tag, class T, class Next = no_property> struct property {};

template <class OutS, class VS, class Dir, class VP, class EP> class adjacency_list {
public:
    size_t n;
    std::vector<std::pair<size_t, size_t>> edges;
    std::vector<int> weights;
    template <class EdgeIt, class WIt> adjacency_list(EdgeIt b, EdgeIt e, WIt w, size_t nv, size_t = 0) : n(nv) {
        for (; b != e; ++b, ++w) {
            edges.push_back(std::make_pair((size_t)b->first, (size_t)b->second));
            weights.push_back((int)*w);
        }
    }
};
template <class G> size_t num_vertices(const G &g) { return g.n; }
struct shim_weight_map { const std::vector<int> *w; };
struct shim_index_map {};
template <class G> shim_weight_map get(edge_weight_t, G &g) { return shim_weight_map{&g.weights}; }
template <class G> shim_index_map get(vertex_index_t, G &) { return shim_index_map(); }
template <class G, class Tag> struct property_map { typedef shim_weight_map type; };
template <class M> struct property_traits { typedef int value_type; };
struct shim_parity_map { std::vector<bool> *bits; };
struct shim_color_holder {
    std::vector<bool> bits;
};
inline shim_color_holder make_one_bit_color_map(size_t n, shim_index_map) { shim_color_holder h; h.bits.assign(n, false); return h; }
inline bool get(const shim_color_holder &h, size_t i) { return h.bits[i]; }
inline shim_parity_map parity_map(shim_color_holder &h) { return shim_parity_map{&h.bits}; }
template <class G> int stoer_wagner_min_cut(const G &g, shim_weight_map, shim_parity_map pm) {
    const size_t n = g.n;
    if (n < 2) return 0;
    std::vector<std::vector<long>> w(n, std::vector<long>(n, 0));
    for (size_t i = 0; i < g.edges.size(); i++) {
        size_t a = g.edges[i].first, b = g.edges[i].second;
        if (a == b) continue;
        w[a][b] += g.weights[i]; w[b][a] += g.weights[i];
    }
    std::vector<std::vector<size_t>> members(n);
    for (size_t i = 0; i < n; i++) members[i].push_back(i);
    std::vector<size_t> active(n);
    for (size_t i = 0; i < n; i++) active[i] = i;
    long best = LONG_MAX;
    std::vector<size_t> bestSide;
    while (active.size() > 1) {
        std::vector<long> key(n, 0);
        std::vector<bool> added(n, false);
        size_t prev = active[0], last = active[0];
        for (size_t it = 0; it < active.size(); it++) {
            size_t sel = (size_t)-1;
            for (size_t v : active) if (!added[v] && (sel == (size_t)-1 || key[v] > key[sel])) sel = v;
            added[sel] = true;
            prev = last; last = sel;
            for (size_t v : active) if (!added[v]) key[v] += w[sel][v];
        }
        if (key[last] < best) { best = key[last]; bestSide = members[last]; }
        for (size_t v : active) if (v != last && v != prev) { w[prev][v] += w[last][v]; w[v][prev] = w[prev][v]; }
        members[prev].insert(members[prev].end(), members[last].begin(), members[last].end());
        for (size_t i = 0; i < active.size(); i++) if (active[i] == last) { active.erase(active.begin() + i); break; }
    }
    if (pm.bits) { pm.bits->assign(n, false); for (size_t v : bestSide) (*pm.bits)[v] = true; }
    return (int)best;
}
}  // namespace boost
#endif
