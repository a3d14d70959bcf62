"""The wire form of a record batch (include/squid_b200.h: sqg_wire; packer squid_b200/csrc/host/wire.cpp): packing a batch and
widening it again -- here with an independent numpy decoder, on the device with k_wire_decode (tests/test_gpu_wire.py) -- must
give back every array bit for bit, including the values that need the exception lists."""
import os

import numpy as np
import pytest

from squid_b200 import api, synth
from tests import common

T = 512


def decode_wire(a: dict, n_rec: int, n_blk: int) -> dict:
    """Numpy restatement of the format's definition."""
    lp = a["lowphred_run"].astype(np.uint16)
    code = (a["aux_nblk"] >> 4).astype(np.int64)
    implied = code == 14
    nb = np.where(implied, 1, code)
    aux = (a["aux_nblk"] & 15).astype(np.uint8)
    esc = (a["dpos"] == 0xFFFF) | (a["span"] == 0xFFFF) | (a["dmate"] == -32768) | (a["lowphred_run"] == 255) | (code == 15)
    e = a["rec_exc"]
    assert np.array_equal(np.flatnonzero(esc), e["idx"].astype(np.int64)), "every escaped record, and only those, is in rec_exc (sorted)"
    assert not (esc & implied).any(), "a record with an implied block needs no exception entry"
    i = e["idx"].astype(np.int64)
    # positions: the running sum of dpos restarts at every tile head (tile_pos) and at every listed record (its own pos)
    ref = np.zeros(n_rec, np.int32); pos = np.zeros(n_rec, np.int64)
    exc_at = dict(zip(i.tolist(), range(len(i))))
    for r in range(n_rec):
        if r in exc_at:
            ref[r], pos[r] = e["ref_id"][exc_at[r]], e["pos"][exc_at[r]]
        elif r % T == 0:
            ref[r], pos[r] = a["tile_ref_id"][r // T], a["tile_pos"][r // T] + int(a["dpos"][r])
        else:
            ref[r], pos[r] = ref[r - 1], pos[r - 1] + int(a["dpos"][r])
    end = pos + a["span"]
    mpos = pos + a["dmate"]
    mref = ref.copy()
    mref[i], mpos[i], end[i], lp[i], nb[i] = e["mate_ref_id"], e["mate_pos"], e["end_pos"], e["lowphred_run"], e["n_blk"]
    off = np.zeros(n_rec + 1, np.int64)
    np.cumsum(nb, out=off[1:])
    assert off[-1] == n_blk
    nw = np.where(implied, 0, nb)  # blocks on the wire
    woff = np.zeros(n_rec + 1, np.int64)
    np.cumsum(nw, out=woff[1:])
    n_wblk = int(woff[-1])
    assert n_wblk == a["blk_dref"].shape[0]
    nt = len(a["tile_blk_off"]) - 1
    assert np.array_equal(off[::T][:nt], a["tile_blk_off"][:-1].astype(np.int64)) and a["tile_blk_off"][-1] == n_blk
    assert np.array_equal(woff[::T][:nt], a["tile_wblk_off"][:-1].astype(np.int64)) and a["tile_wblk_off"][-1] == n_wblk
    assert np.array_equal(np.searchsorted(e["idx"], np.arange(nt + 1) * T), a["tile_rec_exc_off"])
    # explicit blocks
    wrec = np.repeat(np.arange(n_rec), nw)
    wrp = pos[wrec] + a["blk_dref"]
    wml = a["blk_match_ref"].astype(np.int64)
    besc = (a["blk_dref"] == 0xFFFF) | (a["blk_match_ref"] == 0xFFFF)
    be = a["blk_exc"]
    assert np.array_equal(np.flatnonzero(besc), be["idx"].astype(np.int64))
    assert np.array_equal(np.searchsorted(be["idx"], a["tile_wblk_off"]), a["tile_blk_exc_off"])
    k = be["idx"].astype(np.int64)
    wrp[k], wml[k] = be["ref_pos"], be["match_ref"]
    # blocks of the batch: explicit ones from the wire, implied ones from their record
    rp = np.zeros(n_blk, np.int64); ml = np.zeros(n_blk, np.int64); rpos = np.zeros(n_blk, np.uint16); mread = np.zeros(n_blk, np.uint16)
    exp_rec = np.flatnonzero(~implied)
    dst = np.concatenate([np.arange(off[r], off[r + 1]) for r in exp_rec]) if n_wblk else np.zeros(0, np.int64)
    rp[dst], ml[dst], rpos[dst], mread[dst] = wrp, wml, a["blk_read_pos"], a["blk_match_read"]
    imp = np.flatnonzero(implied)
    rp[off[imp]] = pos[imp]; ml[off[imp]] = a["span"][imp]; rpos[off[imp]] = 0; mread[off[imp]] = a["span"][imp]
    return {"ref_id": ref, "pos": pos.astype(np.int32), "mate_ref_id": mref, "mate_pos": mpos.astype(np.int32), "end_pos": end.astype(np.int32), "flag": a["flag"],
            "total_len": a["total_len"], "lowphred_run": lp, "mapq": a["mapq"], "aux": aux, "blk_off": off.astype(np.uint32), "blk_ref_pos": rp.astype(np.int32),
            "blk_match_ref": ml.astype(np.int32), "blk_read_pos": rpos, "blk_match_read": mread}


def adversarial_batch(seed: int, n: int = 3000):
    """A sorted batch whose values sit on every escape boundary of the format."""
    rng = np.random.default_rng(seed)
    ref = np.sort(rng.integers(0, 3, n)).astype(np.int32)
    pos = np.zeros(n, np.int64)
    for c in range(3):
        m = ref == c
        steps = rng.choice([0, 1, 7, 130, 65534, 65535, 65536, 200000], size=int(m.sum()), p=[.2, .3, .3, .14, .02, .02, .01, .01])
        pos[m] = np.cumsum(steps)
    nb = rng.choice([0, 1, 2, 3, 14, 15, 16], size=n, p=[.05, .6, .2, .1, .02, .02, .01])
    off = np.zeros(n + 1, np.int64); np.cumsum(nb, out=off[1:])
    nblk = int(off[-1])
    rec_of = np.repeat(np.arange(n), nb)
    span = rng.choice([0, 100, 65534, 65535, 65536, 1 << 20], size=n, p=[.05, .75, .05, .05, .05, .05])
    dm = rng.choice([-40000, -32768, -32767, -1, 0, 250, 32767, 32768, 1 << 21], size=n, p=[.03, .03, .03, .1, .1, .6, .04, .04, .03])
    mref = np.where(rng.random(n) < 0.05, rng.integers(-1, 3, n), ref).astype(np.int32)
    a = {"ref_id": ref, "pos": pos.astype(np.int32), "mate_ref_id": mref, "mate_pos": (pos + dm).astype(np.int32), "end_pos": (pos + span).astype(np.int32),
         "flag": rng.integers(0, 1 << 12, n).astype(np.uint16), "total_len": rng.integers(30, 65535, n).astype(np.uint16),
         "lowphred_run": rng.choice([0, 3, 254, 255, 256, 40000], size=n, p=[.5, .3, .05, .05, .05, .05]).astype(np.uint16),
         "mapq": rng.integers(0, 256, n).astype(np.uint8), "aux": rng.choice([0, 1, 2, 8, 11], size=n).astype(np.uint8), "blk_off": off.astype(np.uint32),
         "blk_ref_pos": (pos[rec_of] + rng.choice([0, 5, 65534, 65535, 70000], size=nblk, p=[.4, .5, .04, .03, .03])).astype(np.int32),
         "blk_match_ref": rng.choice([1, 3, 100, 65534, 65535, 90000], size=nblk, p=[.1, .1, .7, .04, .03, .03]).astype(np.int32),
         "blk_read_pos": rng.integers(0, 65535, nblk).astype(np.uint16), "blk_match_read": rng.integers(0, 65535, nblk).astype(np.uint16)}
    # a share of the single-block records as plain unclipped reads (the implied block of the format), some of them off by one field
    one = np.flatnonzero(nb == 1)
    plain = one[rng.random(one.shape[0]) < 0.6]
    k = off[plain]
    sp = np.minimum(span[plain], 60000)
    a["end_pos"][plain] = (pos[plain] + sp).astype(np.int32)
    a["blk_ref_pos"][k] = pos[plain].astype(np.int32); a["blk_match_ref"][k] = sp.astype(np.int32)
    a["blk_read_pos"][k] = 0; a["blk_match_read"][k] = sp.astype(np.uint16)
    near = plain[rng.random(plain.shape[0]) < 0.15]
    kn = off[near]
    which = rng.integers(0, 4, near.shape[0])
    a["blk_ref_pos"][kn[which == 0]] += 1; a["blk_match_ref"][kn[which == 1]] -= 1; a["blk_read_pos"][kn[which == 2]] = 5; a["blk_match_read"][kn[which == 3]] += 1
    return api.RecordBatch(a)


def roundtrip(batch):
    w = api.WireBatch(batch)
    d = decode_wire(w.arrays(), batch.n_rec, batch.n_blk)
    for k, v in batch.a.items():
        assert np.array_equal(np.asarray(d[k]).astype(v.dtype), v), k
    return w


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_wire_roundtrip_escape_boundaries(seed):
    b = adversarial_batch(seed)
    w = roundtrip(b)
    assert w.struct.n_rec_exc > 0 and w.struct.n_blk_exc > 0
    assert 0 < w.struct.n_wblk < w.struct.n_blk  # some blocks are implied, some are not


def test_wire_roundtrip_synthetic_case(tmp_path):
    cp, hp, conc, chim, info = common.write_case(str(tmp_path), 20000, seed=5, disc_frac=0.02, ref_len=synth.CHR17_LEN)
    case = api.HostCase(cp, hp)
    w = roundtrip(case.batch)
    s = w.struct
    assert w.nbytes < 0.6 * (32 * s.n_rec + 12 * s.n_blk)  # the point of the format
    assert s.n_rec_exc < 0.2 * s.n_rec


def test_wire_empty_and_ragged():
    e = api.RecordBatch({k: np.zeros(1 if k == "blk_off" else 0, dt) for k, dt in api.BATCH_DTYPES.items()})
    w = api.WireBatch(e)
    assert w.struct.n_tiles == 0 and w.struct.n_rec == 0
    for n in (1, 511, 512, 513, 1025):
        roundtrip(adversarial_batch(7, n))


def test_wire_rejects_wide_aux():
    b = adversarial_batch(4, 100)
    b.a["aux"][50] = 16
    with pytest.raises(api.SquidB200Error):
        api.WireBatch(b)
