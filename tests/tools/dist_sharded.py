"""One shard per torch.distributed rank / GPU over NCCL: the exact range-sharded path of squid_b200.sharded against the
reference build, on a seeded synthetic stream.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tests/tools/dist_sharded.py [n_pairs]
Every rank builds the same case (seeded), plans the same cuts, loads only its own record range, and must end with the
reference's segments, Support, AvgDepth, edges, trimmed chimeric blocks and breakpoint support.  Rank 0 prints one line."""
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402  (test infrastructure: the checker)
from squid_b200 import api, sharded, synth  # noqa: E402
from tests import common  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = tempfile.mkdtemp(prefix="sq_dist_%d_" % rank)
    cp, hp, *_ = common.write_case(d, n_pairs, 1020, 0.05, synth.GRCH38_LEN, fusion_support=10)
    case = api.HostCase(cp, hp)
    cuts = api.plan_shards(case.batch, case.chimeric, case.config, len(case.ref_len), world)
    assert len(cuts) - 1 == world, "planner found %d shards for %d ranks" % (len(cuts) - 1, world)
    comm = sharded.DistComm()
    sg = sharded.ShardedSegmentGraph(case.config, case.ref_len, world, [rank], comm=comm, devices=[local])
    sg.load([case.batch.slice(cuts[rank], cuts[rank + 1])], lambda: api.ChimericReads(case.chimeric.a))
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    nodes = sg.BuildNode_STAR()
    edges = sg.BuildEdges()
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.perf_counter()
    got = {"nodes": np.stack([nodes.Chr, nodes.Position, nodes.Length, nodes.Support], axis=1).astype(np.int32), "avgdepth": nodes.AvgDepth,
           "edges": edges.table(), "chim_after_edges": sg.Chimrecord.block_table()}
    verdict = "unchecked"
    if rank == 0:
        pyref.build()
        ref = pyref.run(cp, hp, os.path.join(d, "ref"))
    objs = [ref if rank == 0 else None]
    dist.broadcast_object_list(objs, src=0)
    ref = objs[0]
    common.assert_same(ref, got)
    sup = api.SegmentGraph.ExactBPConcordantSupport(sg, ref["final_nodes"], ref["final_edges"], pyref.exactbp_map(ref))
    assert sup == pyref.support_map(ref)
    verdict = "bit-exact"
    oks = [None] * world
    dist.all_gather_object(oks, verdict)
    if rank == 0:
        print({"world": world, "records": case.batch.n_rec, "cuts": cuts, "segments": int(nodes.Chr.shape[0]), "edges": int(edges.Weight.shape[0]),
               "verdicts": oks, "rounds": sg.rounds, "build_nodes_edges_s": round(t1 - t0, 4), "backend": dist.get_backend()}, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
