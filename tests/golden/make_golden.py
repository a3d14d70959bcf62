"""Regenerates the committed golden fixtures.  Needs oracle/_ref/squid_ref (i.e. /root/reference present once):
every case is a pair of SQMB inputs plus the seam dumps of the reference's own sources run on them.

    python tests/golden/make_golden.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import pyref  # noqa: E402
from squid_b200 import sqmb, synth  # noqa: E402

CASES = {
    "chr17_3k": dict(n_pairs=3000, seed=3, disc_frac=0.05, ref_len=synth.CHR17_LEN),
    "fourchr_6k": dict(n_pairs=6000, seed=11, disc_frac=0.03, ref_len=[3000000, 2000000, 500000, 16569], n_genes=12),
}
# BWA mode: one merged BAM, `squid_ref --bwa`.  Not implemented on the device yet (round 1 reports
# SQG_EUNSUPPORTED for using_star = 0); the reference's outputs are pinned here so that the next round starts from a checker.
BWA_CASES = {
    "bwa_chr17_3k": dict(n_pairs=3000, seed=3, disc_frac=0.05, ref_len=synth.CHR17_LEN),
    "bwa_fourchr_6k": dict(n_pairs=6000, seed=11, disc_frac=0.03, ref_len=[3000000, 2000000, 500000, 16569], n_genes=12),
}
KEEP = ["nodes_i32.bin", "nodes_f64.bin", "edges_i32.bin", "chim_loaded.bin", "chim_loaded.bin.meta", "chim_after_edges.bin", "final_nodes_i32.bin",
        "final_edges_i32.bin", "exactbp_i32.bin", "support_i32.bin", "readlen.bin",
        # --write-outputs: OutputGraph / WriteBEDPE of the reference (under the harness's stand-in ordering) and what they were fed
        "final_nodes_f64.bin", "labels_i32.bin", "chim_after_exactbp.bin", "components_i32.bin", "edges_before_demultiply_i32.bin", "ref_graph.txt", "ref_sv.txt"]

F1, F2, REV, MREV, PAIRED = 0x40, 0x80, 0x10, 0x20, 0x1


def kat_records():
    """Hand-made alignments for the decoder rules of SURVEY.md App. E (E1-E7); they go into the CHIMERIC file so that the
    reference's chim_loaded dump shows how ReadRec_t decodes each of them."""
    q = lambda n, lo=0: "#" * lo + "I" * (n - lo)
    recs = [
        dict(name_id=1, ref_id=0, pos=1000, cigar="10S50M1000N40M", flag=PAIRED | F1, seq="C" * 100, qual=q(100)),                      # E1
        dict(name_id=2, ref_id=0, pos=1000, cigar="10S50M1000N40M", flag=PAIRED | F1 | REV, seq="C" * 100, qual=q(100)),                # E2
        dict(name_id=3, ref_id=0, pos=500, cigar="30M2I20M3D48M", flag=PAIRED | F1, seq="C" * 100, qual=q(100)),                        # E3
        dict(name_id=4, ref_id=0, pos=700, cigar="5H95M", flag=PAIRED | F1, seq="C" * 95, qual=q(95)),                                  # E4
        dict(name_id=5, ref_id=0, pos=900, cigar="40M1000N60M", flag=PAIRED | F1, seq="A" * 30 + "C" * 10 + "C" * 60, qual=q(100)),     # E5: 30/40 A -> dropped
        dict(name_id=6, ref_id=0, pos=900, cigar="40M1000N60M", flag=PAIRED | F1, seq="a" * 29 + "C" * 11 + "C" * 60, qual=q(100)),     # E5: 29/40 kept
        dict(name_id=7, ref_id=0, pos=1200, cigar="100M", flag=PAIRED | F1, seq="C" * 100, qual=q(100, 11)),                             # E6: run 11 -> low
        dict(name_id=8, ref_id=0, pos=1200, cigar="100M", flag=PAIRED | F2, seq="C" * 100, qual=q(100, 10)),                             # E6: run 10 -> not low
        dict(name_id=9, ref_id=0, pos=1500, cigar="5X95M", flag=PAIRED | F1, seq="C" * 100, qual=q(100)),                                # E7
        dict(name_id=10, ref_id=0, pos=1600, cigar="2I98M", flag=PAIRED | F1, seq="C" * 100, qual=q(100)),                               # E7
        dict(name_id=11, ref_id=0, pos=1700, cigar="20M5P30M10S", flag=PAIRED | F2 | REV, seq="T" * 45 + "C" * 15, qual=q(60)),          # P inside a block, poly-T
        dict(name_id=12, ref_id=1, pos=100, cigar="50=50X", flag=PAIRED | F1, name_suffix=True, seq="C" * 100, qual=q(100)),             # '=' opens a block; /1 suffix
        dict(name_id=12, ref_id=1, pos=5000, cigar="100M", flag=PAIRED | F2, name_suffix=True, seq="C" * 100, qual=q(100)),              # mate "/2" merges by stripped name
    ]
    return recs


def make_bwa():
    pyref.build()
    for name, kw in BWA_CASES.items():
        d = os.path.join(HERE, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        kw = dict(kw)
        tab, info = synth.make_bwa_case(kw.pop("n_pairs"), **kw)
        sqmb.write_sqmb(d + "/all.sqmb", tab)
        pyref.run(d + "/all.sqmb", "-", d + "/ref", extra_args=("--bwa",))
        for f in os.listdir(d + "/ref"):
            if f not in KEEP:
                os.remove(os.path.join(d, "ref", f))
        print(name, "records", tab.n)


def main():
    if "--bwa-only" in sys.argv:
        return make_bwa()
    make_bwa()
    pyref.build()
    for name, kw in CASES.items():
        d = os.path.join(HERE, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        kw = dict(kw)
        conc, chim, info = synth.make_case(kw.pop("n_pairs"), **kw)
        sqmb.write_sqmb(d + "/conc.sqmb", conc); sqmb.write_sqmb(d + "/chim.sqmb", chim)
        pyref.run(d + "/conc.sqmb", d + "/chim.sqmb", d + "/ref", extra_args=("--write-outputs",))
        for f in os.listdir(d + "/ref"):
            if f not in KEEP:
                os.remove(os.path.join(d, "ref", f))
        print(name, "records", conc.n, "chimeric records", chim.n)
    # decoder known-answer case
    d = os.path.join(HERE, "kat_decode")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    ref_len = [3000000, 2000000]
    chim = sqmb.from_records(ref_len, kat_records())
    conc = synth.make_case(300, ref_len=ref_len, seed=5, disc_frac=0.1, n_genes=6)[0]
    sqmb.write_sqmb(d + "/conc.sqmb", conc); sqmb.write_sqmb(d + "/chim.sqmb", chim)
    # only the chimeric loader's view is pinned here (BuildNode on hand-made reads is not meaningful)
    r = os.path.join(d, "ref")
    os.makedirs(r)
    import subprocess
    subprocess.run([pyref.REF_BIN, d + "/conc.sqmb", d + "/chim.sqmb", r, "--quiet", "--stop-after", "nodes"], capture_output=True)
    for f in os.listdir(r):
        if f not in ("chim_loaded.bin", "chim_loaded.bin.meta", "readlen.bin"):
            os.remove(os.path.join(r, f))
    print("kat_decode", np.fromfile(r + "/chim_loaded.bin", dtype=np.int32).reshape(-1, 8))


if __name__ == "__main__":
    main()
