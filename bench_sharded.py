#!/usr/bin/env python
"""Strong-scaling run of the EXACT range-sharded path: ONE sorted stream of the configs[1]
shape is cut at clean cuts (sqg_plan_shards) and every rank runs the path on its record range; seed ops, depth numerators,
edge tables, hints and the coverage chain are exchanged through squid_b200.sharded (torch.distributed over NCCL).  The
outputs are bit-identical for every N, which this script proves by printing checksums of segments, Support, depth
numerators, edges and breakpoint coverage: run it at N = 1, 2, 4 and compare the "checksums" objects.

    python bench_sharded.py --pairs 100000000                                        # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench_sharded.py --pairs 100000000

bench.py stays the contract benchmark (weak scaling over independent streams); this one measures what sharding ONE genome
costs: the chimeric pre-pass is replicated on every rank and bounds the speed-up.
Rank 0 generates the stream on its GPU and hands the ranges over through /dev/shm."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402  (ClockSampler, bps_from_graph)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20_000_000, help="read pairs of the whole stream")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    from squid_b200 import api, build, sharded, sqmb, synth, synth_gpu
    build.build(verbose=False)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = None
    if world > 1:
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)  # NCCL prints on stdout
        dist.init_process_group("nccl", device_id=dev)
    comm = sharded.DistComm() if world > 1 else sharded.LocalComm()
    shm = "/dev/shm/sq_sharded_%s" % os.environ.get("MASTER_PORT", str(os.getpid()))

    # ---- the chimeric reads: every rank builds the same set (seeded) ----------------------------------------------------------
    P = args.pairs
    rng = np.random.Generator(np.random.PCG64(100))
    tx = synth.Transcriptome(rng, np.asarray(synth.GRCH38_LEN, dtype=np.int64), 20000)
    prob = tx.g_expr / tx.g_expr.sum()
    chim_tab, _ = synth.make_chimeric(tx, prob, P, 100, bench.DISC_FRAC, adversarial=False)
    with tempfile.TemporaryDirectory() as d:
        sqmb.write_sqmb(d + "/chim.sqmb", chim_tab)
        sqmb.write_sqmb(d + "/conc.sqmb", sqmb.empty(synth.GRCH38_LEN, 0))
        case = api.HostCase(d + "/conc.sqmb", d + "/chim.sqmb")
    cfg, chim0 = case.config, case.chimeric

    # ---- the stream: generated once (rank 0), planned on the host, handed over as record ranges -----------------------------
    t_plan = 0.0
    if rank == 0:
        batch, _, _ = synth_gpu.make_bench_batch(P, seed=100, device=str(dev))
        os.makedirs(shm, exist_ok=True)
        host = {k: v.cpu().numpy() for k, v in batch.items()}
        del batch
        torch.cuda.empty_cache()
        hb = api.RecordBatch({k: (v.view(np.uint16) if v.dtype == np.int16 else v.view(np.uint32) if (k == "blk_off") else v) for k, v in host.items()})
        t0 = time.perf_counter()
        cuts = api.plan_shards(hb, chim0, cfg, len(case.ref_len), world)
        t_plan = time.perf_counter() - t0
        if len(cuts) - 1 != world:
            raise SystemExit("the planner found %d shards for %d ranks" % (len(cuts) - 1, world))
        for s in range(world):
            sl = hb.slice(cuts[s], cuts[s + 1])
            for k, v in sl.a.items():
                np.save("%s/%d_%s.npy" % (shm, s, k), v)
        json.dump(cuts, open(shm + "/cuts.json", "w"))
        del hb, host
    if world > 1:
        dist.barrier()
    cuts = json.load(open(shm + "/cuts.json"))
    mine = {k: np.load("%s/%d_%s.npy" % (shm, rank, k)) for k in api.BATCH_DTYPES}
    dbatch = {k: torch.from_numpy(v.view(np.int16) if v.dtype == np.uint16 else v.view(np.int32) if v.dtype == np.uint32 else v).to(dev) for k, v in mine.items()}
    R_mine = int(mine["ref_id"].shape[0])
    del mine
    if world > 1:
        dist.barrier()
    if rank == 0:
        for f in os.listdir(shm):
            os.unlink(os.path.join(shm, f))
        os.rmdir(shm)
    dstruct = synth_gpu.batch_struct(dbatch)
    R = cuts[-1]

    sg = sharded.ShardedSegmentGraph(cfg, case.ref_len, world, [rank], comm=comm, devices=[local])
    g = sg.g[0]
    state = {}

    def step():
        pool = state.setdefault("pool", [])
        chim = pool.pop() if pool else api.ChimericReads(chim0.a)
        state["chim"] = chim
        g.attach_concordant_device(dstruct, keepalive=dbatch)
        g.load_chimeric(chim)
        t0 = time.perf_counter()
        nodes = sg.BuildNode_STAR()
        t1 = time.perf_counter()
        edges = sg.BuildEdges(gather_chimeric=False)
        t2 = time.perf_counter()
        if "bps" not in state:
            state["bps"] = bench.bps_from_graph(nodes, edges)
        bc, bp = state["bps"]
        cov = sg.BPCoverage(bc, bp)
        t3 = time.perf_counter()
        tl = state.setdefault("tl", {})
        for k, v in (("BuildNode_STAR", t1 - t0), ("BuildEdges", t2 - t1), ("BPCoverage", t3 - t2)):
            tl[k] = tl.get(k, 0.0) + 1e3 * v
        return nodes, edges, cov

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        nodes, edges, cov = step()
    state["pool"] = [api.ChimericReads(chim0.a) for _ in range(args.steps)]
    state["tl"] = {}
    sg.rounds = {"seeds": 0, "hints": 0, "chain": 0}
    barrier()
    l0 = g.launch_count()
    sampler = bench.ClockSampler(local)
    sampler.start()
    t = time.perf_counter()
    phases = {}
    for _ in range(args.steps):
        nodes, edges, cov = step()
        for ph in ("classify", "seed", "tile", "depth_edges", "edge_sort", "coverage", "k_cov_compact"):
            v = g.phase_ms(ph)
            if v >= 0:
                phases[ph] = phases.get(ph, 0.0) + v / args.steps
    barrier()
    sec = (time.perf_counter() - t) / args.steps
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())
    crc = lambda *arrs: "%08x" % (zlib.crc32(b"".join(np.ascontiguousarray(a).tobytes() for a in arrs)) & 0xFFFFFFFF)
    sums = {"segments": crc(nodes.Chr, nodes.Position, nodes.Length), "support": crc(nodes.Support), "depth_numerators": crc(nodes.count3, nodes.sumlen3),
            "edges": crc(edges.table()), "coverage": crc(cov), "n_segments": int(nodes.Chr.shape[0]), "n_edges": int(edges.Weight.shape[0]),
            "n_breakpoints": int(cov.shape[0]), "coverage_sum": int(cov.sum())}
    per_rank = comm.allgather([{"rank": rank, "records": R_mine, "phases_ms": {k: round(v, 2) for k, v in phases.items()},
                                "host_timeline_ms": {k: round(v / args.steps, 2) for k, v in state["tl"].items()}, "launches_per_step": (g.launch_count() - l0) // args.steps}])
    if rank == 0:
        line = {"metric": "read pairs/s through segment-graph build", "mode": "exact range shards of one stream", "value": (R // 2) / sec, "unit": "read pairs/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong", "dtype": "int32", "data": "synthetic",
                "config": {"workload": "synthetic GRCh38-layout sorted alignment records, %d read pairs in ONE stream, ~0.5%% discordant (configs[1] shape)" % (R // 2),
                           "records": R, "cuts": cuts, "chimeric_reads": int(chim0.n_reads), "planner_s": round(t_plan, 3)},
                "checksums": sums, "exchange_rounds_per_step": {k: v / args.steps for k, v in sg.rounds.items()}, "per_rank": per_rank, "clocks": clocks}
        if real_stdout is not None:
            os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
