"""Shared helpers of the parity tests: make a synthetic case, run the oracle, run the CUDA path, compare."""
import os
import tempfile

import numpy as np

from squid_b200 import sqmb, synth


def write_case(tmpdir, n_pairs, seed, disc_frac=0.02, ref_len=None, **kw):
    conc, chim, info = synth.make_case(n_pairs, ref_len=ref_len if ref_len is not None else synth.CHR17_LEN, seed=seed, disc_frac=disc_frac, **kw)
    cp, hp = os.path.join(tmpdir, "conc.sqmb"), os.path.join(tmpdir, "chim.sqmb")
    sqmb.write_sqmb(cp, conc); sqmb.write_sqmb(hp, chim)
    return cp, hp, conc, chim, info


def run_cuda(cp, hp, device=0, do_cov_with=None):
    """Full hot path through the C ABI.  Returns dict with nodes/avgdepth/edges/chim_after_edges (+coverage map)."""
    from squid_b200 import api
    case = api.HostCase(cp, hp)
    g = api.SegmentGraph(case.config, case.ref_len, device=device)
    nodes = g.BuildNode_STAR(case.chimeric, case.batch)
    edges = g.BuildEdges()
    out = {
        "nodes": np.stack([nodes.Chr, nodes.Position, nodes.Length, nodes.Support], axis=1).astype(np.int32),
        "avgdepth": nodes.AvgDepth, "edges": edges.table(), "chim_after_edges": case.chimeric.block_table(),
        "read_len": case.config.ReadLen, "graph": g, "case": case,
    }
    if do_cov_with is not None:
        ref = do_cov_with
        from oracle import pyref
        out["support"] = g.ExactBPConcordantSupport(ref["final_nodes"], ref["final_edges"], pyref.exactbp_map(ref))
    return out


def assert_same(ref, got, what=("nodes", "avgdepth", "edges", "chim_after_edges")):
    for k in what:
        a, b = ref[k], got[k]
        assert a.shape == b.shape, "%s: shape %s vs %s" % (k, a.shape, b.shape)
        if not np.array_equal(a, b):
            bad = np.flatnonzero((a != b).reshape(a.shape[0], -1).any(axis=1))[:5]
            raise AssertionError("%s differs at rows %s: ref %s got %s" % (k, bad, a[bad], b[bad]))
