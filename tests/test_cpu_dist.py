"""world_size-2 gloo test of the N>1 plumbing on CPU: every rank holds a partial sparse edge table, the tables are
all-gathered and merge-reduced exactly as bench.py does on the device (allgather + sort + reduce-by-key)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, edges_path, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from squid_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    e = np.fromfile(edges_path, dtype=np.int32).reshape(-1, 5)
    rng = np.random.default_rng(7)
    # split every edge's weight between the two ranks (an edge seen by both shards), some edges on one rank only
    w0 = rng.integers(0, e[:, 4] + 1)
    mine = w0 if rank == 0 else e[:, 4] - w0
    keep = mine > 0
    keys = shard.pack_edge_keys(e[keep, 0], e[keep, 1], e[keep, 2], e[keep, 3]).astype(np.int64)
    ws = mine[keep].astype(np.int32)
    n = torch.tensor([keys.shape[0]], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(s.item() for s in sizes))
    kt = torch.zeros(mx, dtype=torch.int64); wt = torch.zeros(mx, dtype=torch.int32)
    kt[: keys.shape[0]] = torch.from_numpy(keys); wt[: keys.shape[0]] = torch.from_numpy(ws)
    gk = [torch.empty_like(kt) for _ in range(world)]; gw = [torch.empty_like(wt) for _ in range(world)]
    dist.all_gather(gk, kt); dist.all_gather(gw, wt)
    tabs = [(gk[r][: int(sizes[r].item())].numpy().astype(np.uint64), gw[r][: int(sizes[r].item())].numpy()) for r in range(world)]
    mk, mw = shard.merge_edge_tables(tabs)
    i1, i2, h1, h2 = shard.unpack_edge_keys(mk)
    np.stack([i1, i2, h1.astype(np.int32), h2.astype(np.int32), mw], axis=1).astype(np.int32).tofile(os.path.join(out_dir, "merged_%d.bin" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_edge_table_merge(tmp_path):
    edges = os.path.join(ROOT, "tests", "golden", "fourchr_6k", "ref", "edges_i32.bin")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, edges, str(tmp_path)), nprocs=2, join=True)
    want = np.fromfile(edges, dtype=np.int32).reshape(-1, 5)
    for r in range(2):
        got = np.fromfile(str(tmp_path / ("merged_%d.bin" % r)), dtype=np.int32).reshape(-1, 5)
        assert np.array_equal(got, want)
