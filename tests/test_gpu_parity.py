"""GPU parity tests proper: the CUDA path through the C ABI against the oracle (the reference's own sources
built against shims, oracle/_ref) on the same seeded synthetic inputs.  Bit-exact: everything is integer work
(AvgDepth is an int sum divided once on the host, so it is compared with == as well)."""
import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu

CASES = [
    # n_pairs, seed, disc_frac, ref_len, kwargs
    (3000, 3, 0.05, None, {}),
    (20000, 17, 0.02, None, {}),
    (30000, 29, 0.005, [3000000, 2000000, 500000, 16569], {"n_genes": 20}),
    (100000, 1003, 0.02, "grch38", {}),
    (200000, 1022, 0.005, [30000000, 20000000, 5000000, 16569], {"n_genes": 300}),
    (400000, 1020, 0.05, "grch38", {"fusion_support": 10}),
]


@pytest.mark.parametrize("n_pairs,seed,disc,ref_len,kw", CASES)
def test_hot_path_matches_reference(tmp_path, built_lib, ref_oracle, monkeypatch, n_pairs, seed, disc, ref_len, kw):
    from oracle import pyref
    from squid_b200 import synth
    if seed in (17, 1003):  # force the chunked (multi-threaded) read loop of the chimeric pre-pass on a small input
        monkeypatch.setenv("SQH_PREPASS_CHUNKS", "13")
    rl = synth.GRCH38_LEN if ref_len == "grch38" else ref_len
    cp, hp, conc, chim, info = common.write_case(str(tmp_path), n_pairs, seed, disc, rl, **kw)
    ref = ref_oracle.run(cp, hp, str(tmp_path / "ref"))
    got = common.run_cuda(cp, hp, do_cov_with=ref)
    assert got["read_len"] == ref["read_len"]
    common.assert_same(ref, got)
    assert got["support"] == pyref.support_map(ref)
    g = got["graph"]
    assert g.launch_count() > 0
    for ph in ("classify", "seed", "depth_edges", "edge_sort", "coverage"):
        assert g.phase_ms(ph) >= 0.0


def test_set_nodes_then_edges(tmp_path, built_lib, ref_oracle):
    """Edges and coverage on an injected segment table (the graph-reload seam, SegmentGraph.cpp:126-157)."""
    from squid_b200 import api
    cp, hp, *_ = common.write_case(str(tmp_path), 20000, 5, 0.02)
    ref = ref_oracle.run(cp, hp, str(tmp_path / "ref"))
    case = api.HostCase(cp, hp)
    g = api.SegmentGraph(case.config, case.ref_len)
    g.load_concordant(case.batch); g.load_chimeric(case.chimeric)
    g.set_nodes(ref["nodes"][:, 0], ref["nodes"][:, 1], ref["nodes"][:, 2])
    e = g.BuildEdges()
    assert np.array_equal(e.table(), ref["edges"])
    assert np.array_equal(case.chimeric.block_table(), ref["chim_after_edges"])


def test_errors_are_loud(tmp_path, built_lib):
    from squid_b200 import api
    cp, hp, *_ = common.write_case(str(tmp_path), 2000, 9, 0.05)
    case = api.HostCase(cp, hp)
    g = api.SegmentGraph(case.config, case.ref_len)
    with pytest.raises(api.SquidB200Error):
        g.BuildEdges()  # nothing loaded
    bad = case.batch.slice(0, case.batch.n_rec)
    bad.a["pos"] = bad.a["pos"][::-1].copy()
    g.load_concordant(bad)  # unsorted: the validation kernel's verdict surfaces at the next synchronising call
    g.load_chimeric(case.chimeric)
    with pytest.raises(api.SquidB200Error) as ei:
        g.BuildNode_STAR()
    assert "sorted" in str(ei.value)
