"""Diagnostic run of the exact range-sharded path on one GPU (several contexts in one process): for every case and shard
count, which outputs differ from the reference build.  Keeps going after a mismatch; exit code = number of failing runs."""
import os
import sys
import tempfile
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from squid_b200 import api, sharded, synth  # noqa: E402
from tests import common  # noqa: E402

CASES = [
    (20000, 17, 0.02, None, {}),
    (30000, 29, 0.005, [3000000, 2000000, 500000, 16569], {"n_genes": 20}),
    (100000, 1003, 0.02, "grch38", {}),
    (200000, 1022, 0.005, [30000000, 20000000, 5000000, 16569], {"n_genes": 300}),
    (400000, 1020, 0.05, "grch38", {"fusion_support": 10}),
]


def run_sharded(case, cuts, ref=None, bps=None, **kw):
    ns = len(cuts) - 1
    sg = sharded.ShardedSegmentGraph(case.config, case.ref_len, ns, list(range(ns)), **kw)
    sg.load([case.batch.slice(cuts[i], cuts[i + 1]) for i in range(ns)], lambda: api.ChimericReads(case.chimeric.a))
    nodes = sg.BuildNode_STAR()
    edges = sg.BuildEdges()
    out = {"nodes": np.stack([nodes.Chr, nodes.Position, nodes.Length, nodes.Support], axis=1).astype(np.int32), "avgdepth": nodes.AvgDepth,
           "edges": edges.table(), "chim_after_edges": sg.Chimrecord.block_table(), "rounds": dict(sg.rounds), "sg": sg}
    if ref is not None:
        out["support"] = api.SegmentGraph.ExactBPConcordantSupport(sg, ref["final_nodes"], ref["final_edges"], pyref.exactbp_map(ref))
    if bps is not None:
        out["cov"] = sg.BPCoverage(bps[:, 0], bps[:, 1])
    out["rounds"] = dict(sg.rounds)
    return out


def dense_bps(case):
    a = case.batch.a
    keys = np.unique((a["ref_id"].astype(np.int64) << 32) | a["pos"].astype(np.int64))
    keys = keys[keys >= 0]
    dense = np.unique(np.concatenate([keys, keys + 1, keys + 2, keys + 37]))
    bp = np.stack([(dense >> 32).astype(np.int32), (dense & 0xffffffff).astype(np.int32)], axis=1)
    return bp[bp[:, 1] < np.asarray(case.ref_len)[bp[:, 0]]]


def main():
    pyref.build()
    fails = 0
    shard_counts = [int(x) for x in os.environ.get("SHARDS", "2,3,5,8").split(",")]
    for n_pairs, seed, disc, ref_len, kw in CASES:
        with tempfile.TemporaryDirectory() as d:
            rl = synth.GRCH38_LEN if ref_len == "grch38" else ref_len
            cp, hp, *_ = common.write_case(d, n_pairs, seed, disc, rl, **kw)
            ref = pyref.run(cp, hp, os.path.join(d, "ref")) if pyref.available() else None
            case = api.HostCase(cp, hp)
            bps = dense_bps(case)
            g1 = api.SegmentGraph(case.config, case.ref_len)
            g1.load_concordant(case.batch); g1.load_chimeric(api.ChimericReads(case.chimeric.a))
            g1.BuildNode_STAR()
            cov1 = g1.BPCoverage(bps[:, 0], bps[:, 1])
            for ns in shard_counts:
                cuts = api.plan_shards(case.batch, case.chimeric, case.config, len(case.ref_len), ns)
                try:
                    got = run_sharded(case, cuts, ref, bps)
                except Exception:
                    fails += 1
                    print("seed %d shards %d cuts %s: EXCEPTION" % (seed, ns, cuts)); traceback.print_exc()
                    continue
                bad = []
                if ref is not None:
                    for k in ("nodes", "avgdepth", "edges", "chim_after_edges"):
                        if ref[k].shape != got[k].shape or not np.array_equal(ref[k], got[k]):
                            bad.append(k)
                            if ref[k].shape == got[k].shape:
                                rows = np.flatnonzero((ref[k] != got[k]).reshape(ref[k].shape[0], -1).any(axis=1))
                                print("   %s: %d rows differ, first %s ref %s got %s" % (k, rows.size, rows[:3], ref[k][rows[:3]].tolist(), got[k][rows[:3]].tolist()))
                            else:
                                print("   %s: shape ref %s got %s" % (k, ref[k].shape, got[k].shape))
                    if got["support"] != pyref.support_map(ref):
                        bad.append("support")
                if not np.array_equal(cov1, got["cov"]):
                    rows = np.flatnonzero(cov1 != got["cov"])
                    print("   dense cov: %d of %d differ, first %s single %s sharded %s" % (rows.size, cov1.size, rows[:5], cov1[rows[:5]], got["cov"][rows[:5]]))
                    bad.append("dense_cov")
                fails += 1 if bad else 0
                print("seed %d shards %d (planned %d) cuts %s rounds %s: %s" % (seed, ns, len(cuts) - 1, cuts, got["rounds"], "OK" if not bad else "DIFF " + ",".join(bad)), flush=True)
                got["sg"].close()
    print("failing runs:", fails)
    return fails


if __name__ == "__main__":
    sys.exit(min(main(), 100))
