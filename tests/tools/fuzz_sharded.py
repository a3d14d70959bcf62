"""Fuzz of the exact range-sharded path on one GPU: random case shapes x random shard counts, every run compared with the
reference build on the whole stream (segments, Support, AvgDepth, edges, trimmed chimeric blocks, breakpoint support) and,
for the dense breakpoint list, with the single-context run.  usage: fuzz_sharded.py SEED_LO SEED_HI"""
import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from squid_b200 import api, synth  # noqa: E402
from tests import common  # noqa: E402
from tests.tools import gpu_sharded_diag as diag  # noqa: E402


def main():
    pyref.build()
    bad = runs = 0
    for seed in range(int(sys.argv[1]), int(sys.argv[2])):
        rnd = random.Random(seed * 104729)
        n = rnd.choice([5000, 20000, 60000, 150000])
        d = rnd.choice([0.005, 0.02, 0.05, 0.1])
        ref = rnd.choice([synth.CHR17_LEN, [30000000, 20000000, 5000000, 16569], synth.GRCH38_LEN, [3000000, 2000000, 500000, 16569]])
        kw = {"n_genes": rnd.choice([None, 30, 300, 2000]), "fusion_support": rnd.choice([5, 10, 20, 100])}
        if rnd.random() < 0.25:
            kw.update(exon_len=(20, 170), intron_len=(60, 400))
        with tempfile.TemporaryDirectory() as td:
            cp, hp, *_ = common.write_case(td, n, seed, d, ref, **kw)
            try:
                refd = pyref.run(cp, hp, os.path.join(td, "ref"))
            except RuntimeError as e:  # the reference itself crashes on this input (undefined behaviour there): not a parity case
                print("seed", seed, "reference build failed:", str(e)[:80], flush=True)
                continue
            case = api.HostCase(cp, hp)
            bps = diag.dense_bps(case)[:: rnd.choice([1, 3, 17])]
            g1 = api.SegmentGraph(case.config, case.ref_len)
            g1.load_concordant(case.batch); g1.load_chimeric(api.ChimericReads(case.chimeric.a))
            try:
                g1.BuildNode_STAR()
            except api.SquidB200Error as e:
                print("seed", seed, "single context refused:", e, flush=True)
                continue
            cov1 = g1.BPCoverage(bps[:, 0], bps[:, 1]) if bps.shape[0] else np.zeros(0, np.int32)
            for ns in rnd.sample([2, 3, 4, 6, 8, 16], 2):
                cuts = api.plan_shards(case.batch, case.chimeric, case.config, len(case.ref_len), ns)
                if len(cuts) - 1 < 2:
                    print("seed", seed, "shards", ns, "no clean cut", flush=True)
                    continue
                runs += 1
                try:
                    got = diag.run_sharded(case, cuts, refd, bps if bps.shape[0] else None)
                    what = [k for k in ("nodes", "avgdepth", "edges", "chim_after_edges") if refd[k].shape != got[k].shape or not np.array_equal(refd[k], got[k])]
                    if got["support"] != pyref.support_map(refd):
                        what.append("support")
                    if bps.shape[0] and not np.array_equal(cov1, got["cov"]):
                        what.append("dense_cov")
                    got["sg"].close()
                except Exception as e:  # noqa: BLE001
                    what = ["EXC %r" % (e,)]
                bad += bool(what)
                print("seed", seed, "n", n, "d", d, "nref", len(ref), kw, "shards", len(cuts) - 1, "rounds", got.get("rounds") if not what or "EXC" not in what[0] else None,
                      "OK" if not what else "DIFF " + ",".join(what), flush=True)
    print("runs", runs, "bad", bad)
    return bad


if __name__ == "__main__":
    sys.exit(min(main(), 100))
