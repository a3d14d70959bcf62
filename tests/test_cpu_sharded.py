"""CPU side of the exact range-sharded path: the clean-cut planner (host code of the library) checked
against its own contract on synthetic streams, the exchange rules of squid_b200.sharded on hand-made cases, and the
all-gather plumbing over a world_size-2 gloo group (two shards per rank, results in shard order on every rank)."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_cut(case, c, margin=200):
    """Conditions A and B of host/plan.cpp, restated with numpy on the packed batch."""
    a = case.batch.a
    rl = case.config.ReadLen
    same = a["ref_id"][:c] == a["ref_id"][c]
    off = a["blk_off"].astype(np.int64)
    last_blk_end = np.where(off[1:c + 1] > off[:c], a["blk_ref_pos"][np.maximum(off[1:c + 1] - 1, 0)] + a["blk_match_ref"][np.maximum(off[1:c + 1] - 1, 0)], 0)
    ends = np.maximum(a["end_pos"][:c], last_blk_end)
    if same.any():
        assert int(a["pos"][c]) - int(ends[same].max()) > rl + 70 + margin
    assert a["ref_id"][c] >= 0 and off[c + 1] > off[c] and a["blk_ref_pos"][off[c]] == a["pos"][c]
    assert a["mapq"][c] >= case.config.Min_MapQual and not (a["flag"][c] & 0x404) and not (a["aux"][c] & 0xB)


def test_planner_contract(tmp_path, built_lib):
    from squid_b200 import api, synth
    for n_pairs, seed, disc, rl, kw in [(20000, 17, 0.02, None, {}), (60000, 5, 0.02, synth.GRCH38_LEN, {})]:
        d = tmp_path / ("c%d" % seed)
        d.mkdir()
        cp, hp, *_ = common.write_case(str(d), n_pairs, seed, disc, rl, **kw)
        case = api.HostCase(cp, hp)
        n = case.batch.n_rec
        assert api.plan_shards(case.batch, case.chimeric, case.config, len(case.ref_len), 1) == [0, n]
        for ns in (2, 4, 8):
            cuts = api.plan_shards(case.batch, case.chimeric, case.config, len(case.ref_len), ns)
            assert cuts[0] == 0 and cuts[-1] == n and 2 <= len(cuts) <= ns + 1
            assert all(x < y for x, y in zip(cuts, cuts[1:]))
            for c in cuts[1:-1]:
                _check_cut(case, c)
            if len(cuts) == ns + 1:  # balanced when the stream offers enough cuts: no shard starts before its nominal share
                assert all(cuts[i] >= (n // ns) * i for i in range(1, ns))


def test_planner_refuses_dirty_stream(tmp_path, built_lib):
    """A stream without a single coverage gap (every record overlaps the next) has no clean cut: one shard."""
    from squid_b200 import api
    cp, hp, *_ = common.write_case(str(tmp_path), 3000, 3, 0.05)
    case = api.HostCase(cp, hp)
    b = case.batch.slice(0, case.batch.n_rec)
    b.a["end_pos"][:] = np.int32(2**30)  # every alignment spans everything to its right
    assert api.plan_shards(b, case.chimeric, case.config, len(case.ref_len), 4) == [0, b.n_rec]


def test_exchange_rules():
    from squid_b200 import sharded as sh
    # seeds: shard 0 never assumes an earlier emission; a shard behind empty shards must rerun without it
    assert sh.first_wrong_prior([False, True, True], [5, 0, 3]) is None
    assert sh.first_wrong_prior([False, True, True], [0, 4, 3]) == (1, False)
    assert sh.first_wrong_prior([False, False, True], [0, 0, 3]) == (2, False)
    assert sh.first_wrong_prior([False, False, False], [0, 0, 3]) is None
    # hints: the nearest earlier shard that located a read supplies firstfrontindex, 0 at the start of the stream
    assert sh.incoming_hints([(False, 7), (True, -1), (False, 9), (False, -1)]) == [0, 7, 7, 9]
    assert sh.incoming_hints([(False, -1), (True, 4)]) == [0, 0]
    # coverage chain: the guess for shard s is the number of breakpoints some earlier shard passes
    assert sh.chain_k_in_guess([3, 3, 10, 7]) == [0, 3, 3, 10]
    # depth numerators add in the reference's int arithmetic
    big = np.full((3, 2), 2**31 - 1, np.int32)
    c, s, o = sh.merge_depth([(big, big, 0), (np.ones((3, 2), np.int32), np.zeros((3, 2), np.int32), 1)])
    assert c.dtype == np.int32 and int(c[0, 0]) == -2**31 and int(s[0, 0]) == 2**31 - 1 and o == 1


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from squid_b200 import sharded as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sh.DistComm()
    ids = [2 * rank, 2 * rank + 1]  # two consecutive shards per rank
    got = comm.allgather([np.full(3, s, np.int32) for s in ids])
    assert [int(x[0]) for x in got] == list(range(2 * world))
    # the seed-op hand-over: shard 0 and 1 came back empty, so shard 1 and then shard 2 must rerun without prior emission
    counts = {0: 0, 1: 0, 2: 4, 3: 1}
    assumed = [s > 0 for s in range(2 * world)]
    reran = []
    while True:
        allc = comm.allgather([counts[s] for s in ids])
        w = sh.first_wrong_prior(assumed, allc)
        if w is None:
            break
        assumed[w[0]] = w[1]
        reran.append(w[0])
    assert reran == [1, 2] and assumed == [False, False, False, True]
    # t of the breakpoints each rank's shards own, max-combined; depth numerators summed with int wrap-around
    t = np.full(6, -1, np.int64)
    t[3 * rank:3 * rank + 3] = 100 * rank + np.arange(3)
    assert comm.allreduce(t, "max").tolist() == [0, 1, 2, 100, 101, 102]
    big = np.array([2**31 - 1, 5], np.int32) if rank == 0 else np.array([1, 7], np.int32)
    assert comm.allreduce(big, "sum").tolist() == [-2**31, 12]
    # ragged arrays (edge keys are uint64 with the top bit in use), two per rank, back in shard order and dtype
    mine = [np.arange(s + 1, dtype=np.uint64) + np.uint64(2**63 + 10 * s) for s in ids]
    got = comm.allgather_arrays(mine)
    assert [g.dtype for g in got] == [np.uint64] * (2 * world)
    assert [g.tolist() for g in got] == [[2**63 + 10 * s + i for i in range(s + 1)] for s in range(2 * world)]
    assert [g.shape[0] for g in comm.allgather_arrays([np.zeros(0, np.int32), np.zeros(rank, np.int32)])] == [0, 0, 0, 1]
    open(os.path.join(out_dir, "ok_%d" % rank), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(str(tmp_path / ("ok_%d" % r))) for r in range(2))
