"""GPU tests on the benchmark-shaped workload (squid_b200.synth_gpu): a reduced-size cross-check against the reference
build, and size-independent properties at a size the CPU oracle would not finish in seconds."""
import os

import numpy as np
import pytest

from tests import common

pytestmark = pytest.mark.gpu


def _run_device_batch(batch, tx, prob, n_pairs, seed, tmp):
    from squid_b200 import api, sqmb, synth, synth_gpu
    chim_tab, _ = synth.make_chimeric(tx, prob, n_pairs, seed, 0.005, adversarial=False)
    sqmb.write_sqmb(tmp + "/chim.sqmb", chim_tab)
    sqmb.write_sqmb(tmp + "/empty.sqmb", sqmb.empty(synth.GRCH38_LEN, 0))
    case = api.HostCase(tmp + "/empty.sqmb", tmp + "/chim.sqmb")
    g = api.SegmentGraph(case.config, case.ref_len)
    g.attach_concordant_device(synth_gpu.batch_struct(batch), keepalive=batch)
    g.load_chimeric(case.chimeric)
    nodes = g.BuildNode_STAR()
    edges = g.BuildEdges()
    return g, case, nodes, edges


def test_bench_generator_matches_reference_at_reduced_size(tmp_path, built_lib, ref_oracle):
    from squid_b200 import sqmb, synth, synth_gpu
    n_pairs, seed = 300_000, 100
    batch, tx, prob = synth_gpu.make_bench_batch(n_pairs, seed=seed, device="cuda", n_genes=2000)
    g, case, nodes, edges = _run_device_batch(batch, tx, prob, n_pairs, seed, str(tmp_path))
    sqmb.write_sqmb(str(tmp_path / "conc.sqmb"), synth_gpu.to_alntable(batch, synth.GRCH38_LEN))
    ref = ref_oracle.run(str(tmp_path / "conc.sqmb"), str(tmp_path / "chim.sqmb"), str(tmp_path / "ref"))
    got_nodes = np.stack([nodes.Chr, nodes.Position, nodes.Length], axis=1)
    assert np.array_equal(got_nodes, ref["nodes"][:, :3])
    assert np.array_equal(edges.table(), ref["edges"])
    assert np.array_equal(nodes.Support, ref["nodes"][:, 3]) and np.array_equal(nodes.AvgDepth, ref["avgdepth"])
    assert np.array_equal(case.chimeric.block_table(), ref["chim_after_edges"])
    from oracle import pyref
    sup = g.ExactBPConcordantSupport(ref["final_nodes"], ref["final_edges"], pyref.exactbp_map(ref))
    assert sup == pyref.support_map(ref)


def test_properties_at_scale(tmp_path, built_lib):
    from squid_b200 import api, shard, synth, synth_gpu
    n_pairs, seed = 5_000_000, 7
    batch, tx, prob = synth_gpu.make_bench_batch(n_pairs, seed=seed, device="cuda")
    g, case, nodes, edges = _run_device_batch(batch, tx, prob, n_pairs, seed, str(tmp_path))
    ref_len = np.asarray(synth.GRCH38_LEN)
    # segments tile every chromosome
    end = nodes.Position + nodes.Length
    same = nodes.Chr[1:] == nodes.Chr[:-1]
    assert np.all(nodes.Length > 0) and np.all(end[:-1][same] == nodes.Position[1:][same])
    first = np.r_[True, ~same]; last = np.r_[~same, True]
    assert np.all(nodes.Position[first] == 0) and np.array_equal(end[last], ref_len[nodes.Chr[last]])
    assert np.array_equal(np.unique(nodes.Chr), np.arange(len(ref_len)))
    # edges: strictly increasing keys, canonical (Ind1 <= Ind2), positive weights that add up to the raw edge count
    keys = shard.pack_edge_keys(edges.Ind1, edges.Ind2, edges.Head1, edges.Head2)
    assert np.all(keys[1:] > keys[:-1]) and np.all(edges.Ind1 <= edges.Ind2) and np.all(edges.Weight > 0)
    assert int(edges.Weight.astype(np.int64).sum()) >= g.stat("raw_edges") >= edges.Weight.shape[0]  # raw (key, count) pairs of the tiles
    # every kept read contributes to at most one segment's Support stream; Support never exceeds the records + blocks
    assert 0 < int(nodes.Support.astype(np.int64).sum()) <= int(batch["blk_ref_pos"].shape[0]) + g.stat("disc_blocks")
    # idempotence and the graph-reload seam: same nodes injected -> same edges
    g2 = api.SegmentGraph(case.config, case.ref_len)
    g2.attach_concordant_device(synth_gpu.batch_struct(batch), keepalive=batch)
    case2 = api.HostCase(str(tmp_path / "empty.sqmb"), str(tmp_path / "chim.sqmb"))
    g2.load_chimeric(case2.chimeric)
    g2.set_nodes(nodes.Chr, nodes.Position, nodes.Length)
    e2 = g2.BuildEdges()
    assert np.array_equal(e2.table(), edges.table())
    assert np.array_equal(case2.chimeric.block_table(), case.chimeric.block_table())
    # coverage: monotone under adding breakpoints is not guaranteed (indBP lag), but a breakpoint list and the same list
    # queried twice must agree, and coverage of a position is bounded by the qualifying records
    bc = nodes.Chr[1:][same][:5000].astype(np.int32); bp = nodes.Position[1:][same][:5000].astype(np.int32)
    c1 = g.BPCoverage(bc, bp); c2 = g2.BPCoverage(bc, bp)
    assert np.array_equal(c1, c2) and c1.min() >= 0 and c1.max() <= batch["ref_id"].shape[0]
