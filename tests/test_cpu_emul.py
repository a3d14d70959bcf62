"""The device-side rules (squid_b200/csrc/*.cuh, `SQ_HD`) stepped on the CPU by tests/emul and compared with the golden
dumps: the event-driven seed machine with island cuts, the closed-form LocateRead with the hint fix-up, the depth
cursor and the coverage chain are all exercised here without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyref
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emul_bin():
    out = os.path.join(ROOT, "tests", "emul", "_build", "emul")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = ["tests/emul/emul_main.cpp", "squid_b200/csrc/host/readrec.cpp", "squid_b200/csrc/host/chimeric.cpp", "squid_b200/csrc/host/prepass.cpp"]
    cmd = ["g++", "-std=c++17", "-O2", "-fopenmp", "-I", "include", "-I", "squid_b200/csrc", "-o", out] + srcs + ["-lpthread"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


@pytest.mark.parametrize("case", ["chr17_3k", "fourchr_6k"])
@pytest.mark.parametrize("one_island,bin_shift,chunks", [(False, 12, 1), (True, 12, 1), (False, 4, 7), (False, 0, 64)])
def test_stepped_rules_match_golden(case, one_island, bin_shift, chunks, emul_bin, tmp_path):
    g = pyref.load_dumps(os.path.join(GOLD, case, "ref"))
    bps = pyref.breakpoints_of(g)
    bps.tofile(str(tmp_path / "bps.bin"))
    env = dict(os.environ)
    env["SQH_PREPASS_CHUNKS"] = str(chunks)    # chimeric pre-pass: read loop split into this many chunks
    env["SQ_EMUL_BIN_SHIFT"] = str(bin_shift)  # segment-table position index: 0 = plain binary searches, 4 = many tiny bins
    if one_island:
        env["SQ_EMUL_ONE_ISLAND"] = "1"
        env["SQ_EMUL_NO_FAST_EDGES"] = "1"  # this variant runs every read through the generic read_edges
    r = subprocess.run([emul_bin, os.path.join(GOLD, case, "conc.sqmb"), os.path.join(GOLD, case, "chim.sqmb"), str(tmp_path), str(tmp_path / "bps.bin")],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    me = pyref.load_dumps(str(tmp_path))
    common.assert_same(g, me, ("nodes", "avgdepth", "edges", "chim_after_edges"))
    cov = np.fromfile(str(tmp_path / "cov_i32.bin"), dtype=np.int32)
    assert pyref.support_from_cov(g, bps, cov) == pyref.support_map(g)


@pytest.mark.parametrize("n,seed,disc,ref_len,kw", [
    (600, 303, 0.3, [3000000, 2000000], dict(n_genes=40, fusion_support=8)),
    (600, 302, 0.3, [3000000, 2000000], dict(n_genes=40, fusion_support=8)),
    (10000, 36, 0.02, None, dict(n_genes=20, fusion_support=20)),
    (30000, 38, 0.05, [3000000, 2000000, 500000, 16569], dict(n_genes=100, fusion_support=60, exon_len=(20, 170), intron_len=(60, 400))),
])
@pytest.mark.parametrize("dense", [False, True])
def test_stepped_rules_with_short_blocks(n, seed, disc, ref_len, kw, dense, emul_bin, ref_oracle, tmp_path):
    """Aligned blocks of 1-3 bp (min_block=1: nothing filtered out): the ReadsOther start masks, the replayed tie order of
    sort(ReadsOther), the short-block branches of the seed windows and of LocateRead, stepped on the CPU against the reference."""
    from squid_b200 import synth
    cp, hp, *_ = common.write_case(str(tmp_path), n, seed, disc, ref_len if ref_len is not None else synth.CHR17_LEN, min_block=1, **kw)
    g = ref_oracle.run(cp, hp, str(tmp_path / "ref"))
    bps = pyref.breakpoints_of(g)
    bps.tofile(str(tmp_path / "bps.bin"))
    env = dict(os.environ)
    if dense:
        env["SQ_EMUL_DENSE"] = "100000"; env["SQ_EMUL_CC_TILE"] = "8"
    os.makedirs(str(tmp_path / "emu"))
    r = subprocess.run([emul_bin, cp, hp, str(tmp_path / "emu"), str(tmp_path / "bps.bin")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "short ReadsOther blocks" in r.stderr and " 0 short ReadsOther" not in r.stderr
    me = pyref.load_dumps(str(tmp_path / "emu"))
    common.assert_same(g, me, ("nodes", "avgdepth", "edges", "chim_after_edges"))
    cov = np.fromfile(str(tmp_path / "emu" / "cov_i32.bin"), dtype=np.int32)
    assert pyref.support_from_cov(g, bps, cov) == pyref.support_map(g)
