"""CPU tests of the host twin (no GPU needed): alignment decoder, chimeric loader, SoA packer, C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(built_lib):
    from squid_b200 import api
    L = api.lib()
    declared = set()
    for h in ("squid_b200.h", "squid_b200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        declared |= set(re.findall(r"\b(sq[gh]_[a-z0-9_]+)\s*\(", src))
    declared -= {"sqg_ctx", "sqh_case"}
    assert len(declared) >= 20
    for s in sorted(declared):
        assert hasattr(L, s), "include/*.h declares %s but the library does not export it" % s


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the context cannot be created: the product path fails loudly."""
    import torch
    from squid_b200 import api
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.SquidB200Error) as e:
        api.SegmentGraph(api.Config(ReadLen=100), [1000000])
    assert e.value.code == api.SQG_ENODEVICE


@pytest.mark.parametrize("case", ["chr17_3k", "fourchr_6k", "kat_decode"])
def test_chimeric_loader_matches_reference(case, built_lib):
    """sqh_open_case (twin of BuildChimericSBamRecord) against the reference's Chimrecord dump."""
    from squid_b200 import api
    g = pyref.load_dumps(os.path.join(GOLD, case, "ref"))
    hc = api.HostCase(os.path.join(GOLD, case, "conc.sqmb"), os.path.join(GOLD, case, "chim.sqmb"))
    assert hc.config.ReadLen == g["read_len"]
    assert np.array_equal(hc.chimeric.block_table(), g["chim_loaded"])
    meta = g["chim_loaded_meta"]
    assert np.array_equal(hc.chimeric.a["first_total_len"], meta[:, 0]) and np.array_equal(hc.chimeric.a["second_total_len"], meta[:, 1])
    for col, key in ((2, "first_lowphred"), (3, "second_lowphred")):
        known = meta[:, col] >= 0  # the reference leaves *LowPhred of an unseen mate uninitialised
        assert np.array_equal(hc.chimeric.a[key][known], meta[known, col])


def test_decoder_known_answers(built_lib, tmp_path):
    """SURVEY.md App. E, through the concordant packer (sqh_case_blocks)."""
    from squid_b200 import api, sqmb
    F1, REV, P = 0x40, 0x10, 0x1
    q = lambda n, lo=0: "#" * lo + "I" * (n - lo)
    recs = [
        dict(ref_id=0, pos=1000, cigar="10S50M1000N40M", flag=P | F1, seq="C" * 100, qual=q(100)),            # E1
        dict(ref_id=0, pos=1000, cigar="10S50M1000N40M", flag=P | F1 | REV, seq="C" * 100, qual=q(100)),      # E2
        dict(ref_id=0, pos=1100, cigar="30M2I20M3D48M", flag=P | F1, seq="C" * 100, qual=q(100)),             # E3
        dict(ref_id=0, pos=1200, cigar="5H95M", flag=P | F1, seq="C" * 95, qual=q(95)),                       # E4
        dict(ref_id=0, pos=1300, cigar="40M1000N60M", flag=P | F1, seq="A" * 30 + "C" * 70, qual=q(100)),     # E5 dropped
        dict(ref_id=0, pos=1300, cigar="40M1000N60M", flag=P | F1, seq="A" * 29 + "C" * 71, qual=q(100)),     # E5 kept
        dict(ref_id=0, pos=1400, cigar="100M", flag=P | F1, seq="C" * 100, qual=q(100, 11)),                  # E6
        dict(ref_id=0, pos=1500, cigar="5X95M", flag=P | F1, seq="C" * 100, qual=q(100)),                     # E7
    ]
    t = sqmb.from_records([3000000], recs)
    sqmb.write_sqmb(str(tmp_path / "c.sqmb"), t)
    sqmb.write_sqmb(str(tmp_path / "h.sqmb"), sqmb.empty([3000000], 0))
    hc = api.HostCase(str(tmp_path / "c.sqmb"), str(tmp_path / "h.sqmb"))
    B = lambda r: [tuple(int(v) for v in row) for row in hc.blocks(r)[0]]  # (ref_pos, match_ref, read_pos, match_read)
    assert B(0) == [(1000, 50, 10, 50), (2050, 40, 60, 40)] and hc.blocks(0)[1] == 100
    assert B(1) == [(1000, 50, 40, 50), (2050, 40, 0, 40)]
    assert B(2) == [(1100, 101, 0, 100)]
    assert B(3) == [(1200, 95, 5, 95)] and hc.blocks(3)[1] == 100
    assert B(4) == [(2340, 60, 40, 60)]
    assert B(5) == [(1300, 40, 0, 40), (2340, 60, 40, 60)]
    assert hc.blocks(6)[2] == 11
    assert B(7) == [(1500, 95, 0, 95)] and hc.blocks(7)[1] == 100
    assert int(hc.batch.a["end_pos"][0]) == 1000 + 50 + 1000 + 40


def test_packer_gate_bits_and_roundtrip(built_lib):
    """aux bits (XA / IH>1 / ChimName) and the SoA layout on a golden case."""
    from squid_b200 import api, sqmb
    case = os.path.join(GOLD, "fourchr_6k")
    hc = api.HostCase(case + "/conc.sqmb", case + "/chim.sqmb")
    b = hc.batch
    assert b.a["blk_off"][-1] == b.n_blk and np.all(np.diff(b.a["blk_off"].astype(np.int64)) >= 0)
    key = b.a["ref_id"].astype(np.int64) * (1 << 32) + b.a["pos"]
    assert np.all(np.diff(key[b.a["ref_id"] >= 0]) >= 0)
    assert (b.a["aux"] & 8).any(), "some concordant records must carry a chimeric read name (ChimName gate)"
    assert (b.a["aux"] & 2).any() and (b.a["aux"] & 1).any()


def test_shard_plan_and_edge_merge():
    from squid_b200 import shard
    ref = np.repeat(np.arange(5), [10, 50, 5, 30, 5]).astype(np.int32)
    plan = shard.plan_shards(ref, 2)
    assert plan[0][0] == 0 and plan[-1][1] == 100 and plan[0][1] == plan[1][0]
    cut = plan[0][1]
    assert cut in (0, 10, 60, 65, 95, 100)  # chromosome boundaries only
    k = shard.pack_edge_keys([1, 1, 7], [2, 5, 9], [0, 1, 1], [1, 0, 1])
    a = (k[[0, 2]], np.array([3, 4], np.int32)); b2 = (k[[0, 1]], np.array([5, 6], np.int32))
    mk, mw = shard.merge_edge_tables([a, b2])
    i1, i2, h1, h2 = shard.unpack_edge_keys(mk)
    assert list(zip(i1, i2, h1.astype(int), h2.astype(int), mw)) == [(1, 2, 0, 1, 8), (1, 5, 1, 0, 6), (7, 9, 1, 1, 4)]


def test_parallel_sort_reproduces_std_sort(built_lib):
    """The chimeric pre-pass sorts with a multi-threaded restatement of libstdc++'s introsort; it must leave equal keys in
    exactly std::sort's order (that order is observable)."""
    import ctypes as C
    from squid_b200 import api
    L = api.lib()
    L.sqh_selftest_sort.argtypes = [C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.sqh_selftest_sort.restype = C.c_int
    for n in (0, 1, 16, 17, 1000, 4097, 40000, 200000, 1000003):
        for rng in (1, 3, 100, 1 << 40):
            for pat in range(4):
                for fan in (0, 3, 8):  # threads: 0 = the sequential introsort loop alone
                    if n > 500000 and (fan != 8 or rng not in (3, 1 << 40)):
                        continue  # the largest size only on the fully parallel path
                    assert L.sqh_selftest_sort(n, n * 31 + rng + pat, rng, pat, fan) == 1, (n, rng, pat, fan)


def test_chimname_gate_bits_match_brute_force(tmp_path):
    """SQG_AUX_CHIMNAME of the packed batch (SegmentGraph.cpp:302: the RAW record name looked up in the suffix-stripped Qnames of
    Chimrecord, plus the empty strings ChimName is pre-sized with) against a plain Python set -- the packer answers through a Bloom
    filter first, so every bit is checked, hits and misses."""
    from squid_b200 import api, sqmb, synth
    conc, chim, info = synth.make_case(30000, ref_len=synth.CHR17_LEN, seed=77, disc_frac=0.05)
    # give some concordant records the name of a chimeric read (with and without the /1 /2 suffix)
    rng = np.random.default_rng(3)
    take = rng.choice(conc.n, size=400, replace=False)
    conc.name_id[take] = rng.choice(chim.name_id, size=400)
    cp, hp = str(tmp_path / "c.sqmb"), str(tmp_path / "h.sqmb")
    sqmb.write_sqmb(cp, conc); sqmb.write_sqmb(hp, chim)

    def raw_name(t, r):
        s = "q%d" % t.name_id[r]
        if t.aux[r] & sqmb.AUX_NAME_SUFFIX:
            s += "/2" if t.flag[r] & 0x80 else "/1"
        return s
    names = {""}
    for r in range(chim.n):
        if (chim.flag[r] & 0x4) or (chim.flag[r] & 0x400):
            continue
        s = raw_name(chim, r)
        names.add(s[:-2] if s.endswith(("/1", "/2")) else s)
    want = np.array([raw_name(conc, r) in names for r in range(conc.n)])
    case = api.HostCase(cp, hp)
    got = (case.batch.a["aux"] & 8) != 0
    assert want.sum() > 50 and (~want).sum() > 1000
    assert np.array_equal(got, want)
