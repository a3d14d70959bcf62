"""BASELINE.json configs[2..4] at a prefix size the reference finishes in a minute: the benchmark generator
(squid_b200.synth_gpu, App. C block mix) with
  C3  STAR split mode where 1 % of the chimeric read names also appear in the concordant stream (the ChimName gate,
      SegmentGraph.cpp:302, on raw names at scale),
  C4  tumour-like input, 5 % discordant pairs,
  C5  boundary-heavy cohort shape: low-support fusions (lambda = 9), four times the segments per read of C2,
each compared `==` with the reference build on the same records: segments, Support, AvgDepth, edges, trimmed chimeric blocks and
the breakpoint support of the reference's own final graph."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_PAIRS = int(os.environ.get("SQUID_CONFIG_TEST_PAIRS", 4_000_000))


def _run(tmp, n_pairs, seed, disc_frac, fusion_support, chimname_frac, ref_oracle):
    import torch
    from oracle import pyref
    from squid_b200 import api, sqmb, synth, synth_gpu
    import bench
    batch, tx, prob = synth_gpu.make_bench_batch(n_pairs, seed=seed, device="cuda", exon_len=bench.BENCH_EXON_LEN, min_block=bench.BENCH_MIN_BLOCK)
    chim_tab, _ = synth.make_chimeric(tx, prob, n_pairs, seed, disc_frac, fusion_support=fusion_support, adversarial=False)
    tab = synth_gpu.to_alntable(batch, synth.GRCH38_LEN)
    n_gate = 0
    if chimname_frac > 0:  # give some concordant records the NAME of a chimeric read (suffix-free, so the raw-name gate fires)
        rng = np.random.Generator(np.random.PCG64(seed + 1))
        names = np.unique(chim_tab.name_id)
        take = rng.choice(names, size=max(1, int(chimname_frac * names.shape[0])), replace=False)
        rec = rng.choice(tab.n, size=take.shape[0], replace=False)
        tab.name_id[rec] = take
        aux = batch["aux"].clone()
        aux[torch.as_tensor(rec, device=aux.device)] |= 8  # SQG_AUX_CHIMNAME: what the host packer's name probe would set
        batch["aux"] = aux
        n_gate = int(take.shape[0])
    cp, hp = tmp + "/conc.sqmb", tmp + "/chim.sqmb"
    sqmb.write_sqmb(cp, tab); sqmb.write_sqmb(hp, chim_tab)
    sqmb.write_sqmb(tmp + "/empty.sqmb", sqmb.empty(synth.GRCH38_LEN, 0))
    case = api.HostCase(tmp + "/empty.sqmb", hp)
    g = api.SegmentGraph(case.config, case.ref_len)
    g.attach_concordant_device(synth_gpu.batch_struct(batch), keepalive=batch)
    g.load_chimeric(case.chimeric)
    nodes = g.BuildNode_STAR()
    edges = g.BuildEdges()
    ref = ref_oracle.run(cp, hp, tmp + "/ref", timeout=3000)
    got_nodes = np.stack([nodes.Chr, nodes.Position, nodes.Length, nodes.Support], axis=1).astype(np.int32)
    assert np.array_equal(got_nodes, ref["nodes"]), "segments / Support differ"
    assert np.array_equal(nodes.AvgDepth, ref["avgdepth"])
    assert np.array_equal(edges.table(), ref["edges"])
    assert np.array_equal(case.chimeric.block_table(), ref["chim_after_edges"])
    sup = g.ExactBPConcordantSupport(ref["final_nodes"], ref["final_edges"], pyref.exactbp_map(ref))
    assert sup == pyref.support_map(ref)
    return nodes, edges, n_gate, int(case.chimeric.n_reads)


def test_c3_chimname_overlap(tmp_path, built_lib, ref_oracle):
    nodes, edges, n_gate, n_chim = _run(str(tmp_path), N_PAIRS, 200, 0.005, 20.0, 0.01, ref_oracle)
    assert n_gate >= 100 and n_chim > 10000


def test_c4_five_percent_discordant(tmp_path, built_lib, ref_oracle):
    nodes, edges, _, n_chim = _run(str(tmp_path), N_PAIRS, 404, 0.05, 20.0, 0.0, ref_oracle)
    assert n_chim > 0.03 * N_PAIRS


def test_c5_boundary_heavy(tmp_path, built_lib, ref_oracle):
    nodes, edges, _, _ = _run(str(tmp_path), N_PAIRS, 1000, 0.02, 9.0, 0.0, ref_oracle)
    assert nodes.Chr.shape[0] > N_PAIRS // 500  # at least the segment density of configs[4] (2 M segments per 1 B pairs)
