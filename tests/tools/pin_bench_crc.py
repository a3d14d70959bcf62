"""Pins the benchmark workload to the reference: generates bench.py's exact configs[1] input on the GPU,
writes it as SQMB, runs the reference's own sources (oracle/_ref/squid_ref) on it -- one core, minutes -- and the CPU
restatement for the breakpoint coverage of the bench's stand-in breakpoint list, and stores the CRC32s of every output in
tests/golden/bench_crc.json.  bench.py compares its own outputs with these on every run.

    python tests/tools/pin_bench_crc.py [--pairs 100000000] [--seed 100] [--out gpurun_out/bench_crc_new.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xffffffff


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=bench.DEFAULT_PAIRS)
    ap.add_argument("--seed", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_crc_new.json"))
    ap.add_argument("--tmp", default=None)
    args = ap.parse_args()
    import torch
    from oracle import pyref
    from squid_b200 import api, sqmb, synth, synth_gpu
    t0 = time.time()
    batch, tx, prob = bench.make_workload(args.pairs, args.seed, "cuda")
    chim_tab, _ = synth.make_chimeric(tx, prob, args.pairs, args.seed, bench.DISC_FRAC, adversarial=False)
    import shutil
    shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 60e9
    tmp = args.tmp or tempfile.mkdtemp(prefix="sqpin_", dir="/dev/shm" if shm_ok else None)
    os.makedirs(tmp, exist_ok=True)
    cp, hp = tmp + "/conc.sqmb", tmp + "/chim.sqmb"
    sqmb.write_sqmb(hp, chim_tab)
    sqmb.write_sqmb(tmp + "/empty.sqmb", sqmb.empty(synth.GRCH38_LEN, 0))
    R = int(batch["ref_id"].shape[0])
    print("generated %d records in %.1f s" % (R, time.time() - t0), flush=True)
    # ---- ours
    case = api.HostCase(tmp + "/empty.sqmb", hp)
    g = api.SegmentGraph(case.config, case.ref_len)
    g.attach_concordant_device(synth_gpu.batch_struct(batch), keepalive=batch)
    g.load_chimeric(case.chimeric)
    nodes = g.BuildNode_STAR()
    edges = g.BuildEdges()
    bc, bp = bench.bps_from_graph(nodes, edges)
    cov = g.BPCoverage(bc, bp)
    ours = bench.output_crcs(nodes, edges, case.chimeric.block_table(), cov)
    print("ours", ours, flush=True)
    # ---- the reference's own sources on the same records
    t0 = time.time()
    sqmb.write_sqmb(cp, synth_gpu.to_alntable(batch, synth.GRCH38_LEN))
    print("SQMB written in %.1f s (%.1f GB)" % (time.time() - t0, os.path.getsize(cp) / 1e9), flush=True)
    del batch
    torch.cuda.empty_cache()
    t0 = time.time()
    ref = pyref.run(cp, hp, tmp + "/ref", extra_args=("--stop-after", "edges"), timeout=7200)
    t_ref = time.time() - t0
    print("reference: %.1f s, %s" % (t_ref, ref["timings"]), flush=True)
    want = {"nodes": crc(ref["nodes"]), "avgdepth": crc(ref["avgdepth"]), "edges": crc(ref["edges"]), "chim_after_edges": crc(ref["chim_after_edges"]),
            "n_nodes": int(ref["nodes"].shape[0]), "n_edges": int(ref["edges"].shape[0])}
    # ---- breakpoint coverage of the bench's stand-in breakpoint list: the CPU restatement (pinned to the reference on the
    #      test cases; the reference itself only counts for the breakpoints of its own final graph)
    t0 = time.time()
    me = pyref.run_restate(cp, hp, tmp + "/me", bps=np.stack([bc, bp], axis=1))
    print("restatement: %.1f s" % (time.time() - t0), flush=True)
    want["coverage"] = crc(me["cov"]); want["n_bp"] = int(bc.shape[0])
    restate_same = all(crc(me[k]) == want[k] for k in ("nodes", "avgdepth", "edges", "chim_after_edges"))
    same = {k: ours[k] == want[k] for k in want}
    rec = {"workload": bench.workload_id(args.pairs, args.seed), "records": R, "reference_seconds": t_ref, "reference_timings": ref["timings"],
           "reference_crc32": want, "restatement_equals_reference": restate_same, "cuda_path_crc32": ours, "cuda_equals_reference": same}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rec, open(args.out, "w"), indent=1)
    print(json.dumps(rec), flush=True)
    for f in (cp, hp):
        os.remove(f)
    assert all(same.values()) and restate_same, "MISMATCH against the reference at full size"


if __name__ == "__main__":
    main()
