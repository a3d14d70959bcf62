"""SegmentGraph_t::ConnectedComponent on the device (sqg_connected_components, squid_b200/csrc/sq_cc.cuh):
labels equal to the reference's own on its final graphs (golden dumps of oracle/_ref), on fresh reference runs, and -- at the
~2 M-node size the row is about, where the reference's O(#components x N) rescan is out of reach -- equal to an independent
labelling by scipy ordered the same way."""
import os

import numpy as np
import pytest

from squid_b200 import api, synth
from tests import common

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def by_smallest_node(n, a, b):
    """Independent labelling: scipy's components renumbered in the order of their smallest node."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    m = coo_matrix((np.ones(len(a), np.int8), (a, b)), shape=(n, n))
    _, lab = connected_components(m, directed=False)
    first = np.full(lab.max() + 1 if n else 0, n, np.int64)
    np.minimum.at(first, lab, np.arange(n))
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order); rank[order] = np.arange(len(order))
    return rank[lab].astype(np.int32)


@pytest.mark.parametrize("case", ["chr17_3k", "fourchr_6k"])
def test_labels_match_reference_golden(case):
    d = os.path.join(GOLD, case, "ref")
    nodes = np.fromfile(os.path.join(d, "final_nodes_i32.bin"), np.int32).reshape(-1, 4)
    edges = np.fromfile(os.path.join(d, "final_edges_i32.bin"), np.int32).reshape(-1, 5)
    want = np.fromfile(os.path.join(d, "labels_i32.bin"), np.int32)
    got = api.ConnectedComponent(nodes.shape[0], edges[:, 0], edges[:, 1])
    assert np.array_equal(got, want)


def test_labels_match_reference_run(ref_oracle, tmp_path):
    cp, hp, conc, chim, info = common.write_case(str(tmp_path), 200000, seed=23, disc_frac=0.03, ref_len=synth.GRCH38_LEN, fusion_support=8)
    ref = ref_oracle.run(cp, hp, str(tmp_path / "ref"))
    want = np.fromfile(str(tmp_path / "ref" / "labels_i32.bin"), np.int32)
    fe = ref["final_edges"]
    got = api.ConnectedComponent(ref["final_nodes"].shape[0], fe[:, 0], fe[:, 1])
    assert want.shape[0] > 50 and np.array_equal(got, want)


@pytest.mark.parametrize("n,m,seed", [(1, 0, 1), (7, 0, 2), (50, 200, 3), (5000, 3000, 4), (2_000_000, 1_500_000, 5), (2_000_000, 6_000_000, 6)])
def test_labels_on_random_graphs(n, m, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, n, m).astype(np.int32); b = rng.integers(0, n, m).astype(np.int32)  # self-loops and duplicates included
    if m > 10:
        a[:5] = b[:5]
        a[5:10], b[5:10] = a[0:5], b[0:5]
    got = api.ConnectedComponent(n, a, b)
    assert np.array_equal(got, by_smallest_node(n, a, b))


def test_chain_and_star_graphs():
    n = 300000
    i = np.arange(n - 1, dtype=np.int32)
    assert np.array_equal(api.ConnectedComponent(n, i + 1, i), np.zeros(n, np.int32))                            # one long path, edges reversed
    assert np.array_equal(api.ConnectedComponent(n, np.full(n - 1, n - 1, np.int32), i), np.zeros(n, np.int32))  # star around the LAST node
    j = np.arange(n - 2, dtype=np.int32)
    assert np.array_equal(api.ConnectedComponent(n, j + 2, j), (np.arange(n) % 2).astype(np.int32))              # evens and odds
    k = j[::2]
    got = api.ConnectedComponent(n, k + 2, k)                                                                    # evens chained, odds isolated
    assert np.array_equal(got, by_smallest_node(n, k + 2, k)) and got[0] == 0 and got[1] == 1 and got[3] == 2


def test_bad_edge_is_refused():
    with pytest.raises(api.SquidB200Error):
        api.ConnectedComponent(10, np.array([3], np.int32), np.array([10], np.int32))
