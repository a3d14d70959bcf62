"""BGZF/BAM front end of the host packer (squid_b200/csrc/host/bam.cpp, sqh_open_bam_case) against the
SQMB path that the oracle pins: the same alignment table written as a real BAM file (squid_b200.bamio, an independent
implementation of the SAM/BAM specification) must give byte-identical packed batches, chimeric reads, ReadLen and
reference lengths.  Plus known-answer records with explicit bases / qualities / clips, and loud failures on damaged files."""
import os
import zlib

import numpy as np
import pytest

from tests import common


def _same_case(a, b):
    for k in a.batch.a:
        assert np.array_equal(a.batch.a[k], b.batch.a[k]), k
    for k in a.chimeric.a:
        assert np.array_equal(a.chimeric.a[k], b.chimeric.a[k]), k
    assert a.config == b.config and np.array_equal(a.ref_len, b.ref_len)


def _probe(path):
    import ctypes as C
    from squid_b200 import api
    L = api.lib()
    L.sqh_probe_bam.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    n, pm = C.c_int64(), C.c_int32()
    assert L.sqh_probe_bam(path.encode(), C.byref(n), C.byref(pm)) == 0
    return n.value, bool(pm.value)


@pytest.mark.parametrize("block,compressed,htslib", [(0xFF00, True, False), (997, True, False), (0, False, False), (0xFF00, True, True), (1500, True, True)])
def test_bam_equals_sqmb(tmp_path, built_lib, block, compressed, htslib):
    from squid_b200 import api, bamio, synth
    cp, hp, conc, chim, info = common.write_case(str(tmp_path), 6000, 17, 0.05, synth.GRCH38_LEN)
    assert (conc.aux & 1).any() and (conc.aux & 2).any() and conc.lowrun.any()  # XA, IH and low-quality runs are exercised
    cb, hb = str(tmp_path / "conc.bam"), str(tmp_path / "chim.bam")
    bamio.write_bam(cb, conc, block=block or 0xFF00, compressed=compressed, htslib_blocks=htslib)  # small blocks: records straddle BGZF members
    bamio.write_bam(hb, chim, block=block or 0xFF00, compressed=compressed, htslib_blocks=htslib)
    _same_case(api.HostCase(cp, hp), api.HostCase(cb, hb, bam=True))
    # members that end at record boundaries (htslib) are walked in parallel -- after that has been proven for the file; any
    # other layout takes the sequential walk
    n, per_member = _probe(cb)
    assert n == conc.n and per_member == (compressed and htslib)


def test_bam_known_answers(tmp_path, built_lib):
    """Explicit bases and qualities, hard and soft clips, I/D/N/=/X, a poly-A block, tags of every integer width."""
    from squid_b200 import api, bamio, sqmb
    F = sqmb
    recs = [
        dict(ref_id=0, pos=100, cigar="5H10S40M5I20M2D10M1000N15M", flag=F.FLAG_PAIRED | F.FLAG_FIRST | F.FLAG_MATE_REVERSE, mate_ref_id=0, mate_pos=1300,
             seq="C" * 10 + "G" * 75 + "A" * 15, qual="!" * 13 + "I" * 87, name_id=7, ih=1),
        dict(ref_id=0, pos=1300, cigar="30=5X30M35S", flag=F.FLAG_PAIRED | F.FLAG_SECOND | F.FLAG_REVERSE, mate_ref_id=0, mate_pos=100,
             seq="ACGT" * 25, qual="I" * 100, name_id=7, xa=True, ih=3, name_suffix=True),
        dict(ref_id=1, pos=5, cigar="100M", flag=F.FLAG_PAIRED | F.FLAG_FIRST, mate_ref_id=1, mate_pos=400, lowrun=12, polya=1, name_id=9, mapq=3),
        # odd length: the last base sits alone in the high nibble of the last packed byte, and here it decides the poly-A rule
        # (72 of 95 bases are A: 4*72 >= 3*95, dropped; with 71 it would be kept)
        dict(ref_id=1, pos=900, cigar="95M", flag=F.FLAG_PAIRED | F.FLAG_FIRST, mate_ref_id=1, mate_pos=1400, seq="A" * 71 + "C" * 23 + "A", qual="I" * 95, name_id=11),
        dict(ref_id=1, pos=950, cigar="95M", flag=F.FLAG_PAIRED | F.FLAG_FIRST, mate_ref_id=1, mate_pos=1400, seq="A" * 71 + "C" * 23 + "G", qual="I" * 95, name_id=12),
    ]
    t = F.from_records([5000, 3000], recs)
    ch = F.from_records([5000, 3000], [dict(ref_id=0, pos=10, cigar="60M40S", flag=F.FLAG_PAIRED | F.FLAG_FIRST, mate_ref_id=1, mate_pos=50, name_id=1),
                                       dict(ref_id=1, pos=700, cigar="60S40M", flag=F.FLAG_PAIRED | F.FLAG_FIRST | F.FLAG_REVERSE, mate_ref_id=1, mate_pos=50, name_id=1),
                                       dict(ref_id=1, pos=50, cigar="100M", flag=F.FLAG_PAIRED | F.FLAG_SECOND | F.FLAG_REVERSE, mate_ref_id=0, mate_pos=10, name_id=1)])
    cp, hp, cb, hb = (str(tmp_path / n) for n in ("c.sqmb", "h.sqmb", "c.bam", "h.bam"))
    F.write_sqmb(cp, t); F.write_sqmb(hp, ch); bamio.write_bam(cb, t, block=64); bamio.write_bam(hb, ch)
    a, b = api.HostCase(cp, hp), api.HostCase(cb, hb, bam=True)
    _same_case(a, b)
    blk, total, low = b.blocks(0)
    # 40M5I20M2D10M is one block (I adds to the read span, D to the reference span), N splits; the 15-bp poly-A block is dropped
    assert blk.tolist() == [[100, 72, 15, 75]] and total == 105 and low == 13
    blk, total, low = b.blocks(1)
    assert blk.tolist() == [[1300, 65, 35, 65]] and total == 100  # reverse strand: read_pos = 100 - 0 - 65
    assert b.batch.a["aux"].tolist() == [0, 3, 0, 0, 0]  # XA + IH>1 on record 1; the suffixed name never matches ChimName
    assert b.blocks(3)[0].shape[0] == 0 and b.blocks(4)[0].tolist() == [[950, 95, 0, 95]]
    assert b.blocks(2)[0].shape[0] == 0 and b.blocks(2)[2] == 12  # poly-A block removed, 12 low qualities


def test_damaged_files_fail_loudly(tmp_path, built_lib):
    from squid_b200 import api, bamio
    cp, hp, conc, chim, info = common.write_case(str(tmp_path), 500, 3, 0.05)
    good = bamio.bgzf_compress(bamio.bam_bytes(conc), 4096)
    hb = str(tmp_path / "chim.bam")
    bamio.write_bam(hb, chim)

    def opens(data):
        p = str(tmp_path / "x.bam")
        open(p, "wb").write(data)
        return api.HostCase(p, hb, bam=True)
    opens(good)
    for bad, what in ((good[: len(good) // 2], "truncated"), (good[:200] + bytes([good[200] ^ 0x55]) + good[201:], "flipped byte"),
                      (zlib.compress(bamio.bam_bytes(conc)), "zlib, not BGZF"), (b"BAM\1" + b"\xff" * 40, "garbage header")):
        with pytest.raises(api.SquidB200Error):
            opens(bad)
    with pytest.raises(api.SquidB200Error):
        api.HostCase(str(tmp_path / "missing.bam"), hb, bam=True)
