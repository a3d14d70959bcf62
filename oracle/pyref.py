"""TEST INFRASTRUCTURE.  Python driver of the oracle: builds and runs oracle/_ref/squid_ref (the reference's own
sources against shims) and loads its seam dumps.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product (squid_b200/) never does.
"""
from __future__ import annotations

import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "squid_ref")
REFERENCE_SRC = "/root/reference"


def build(verbose: bool = False) -> None:
    """(Re)builds oracle/_ref when the reference sources are present (development container only);
    on the GPU box the prebuilt binary that travelled with the snapshot is used as is."""
    if os.path.isdir(os.path.join(REFERENCE_SRC, "src")):
        srcs = [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "Makefile")] + [os.path.join(HERE, "shim", d, f) for d, _, fs in os.walk(os.path.join(HERE, "shim")) for f in fs]
        if os.path.exists(REF_BIN) and all(os.path.getmtime(s) <= os.path.getmtime(REF_BIN) for s in srcs if os.path.exists(s)):
            return
        r = subprocess.run(["make", "-C", HERE, "ref"], capture_output=True, text=True)
        if verbose:
            print(r.stdout[-2000:], r.stderr[-2000:])
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + r.stderr[-4000:])


def available() -> bool:
    return os.path.exists(REF_BIN)


def _i32(path: str, cols: int) -> np.ndarray:
    return np.fromfile(path, dtype=np.int32).reshape(-1, cols) if os.path.exists(path) else np.zeros((0, cols), np.int32)


def load_dumps(outdir: str) -> dict:
    d = {
        "nodes": _i32(os.path.join(outdir, "nodes_i32.bin"), 4),
        "avgdepth": np.fromfile(os.path.join(outdir, "nodes_f64.bin")) if os.path.exists(os.path.join(outdir, "nodes_f64.bin")) else np.zeros(0),
        "edges": _i32(os.path.join(outdir, "edges_i32.bin"), 5),
        "chim_loaded": _i32(os.path.join(outdir, "chim_loaded.bin"), 8),
        "chim_loaded_meta": _i32(os.path.join(outdir, "chim_loaded.bin.meta"), 4),
        "chim_after_edges": _i32(os.path.join(outdir, "chim_after_edges.bin"), 8),
        "final_nodes": _i32(os.path.join(outdir, "final_nodes_i32.bin"), 4),
        "final_edges": _i32(os.path.join(outdir, "final_edges_i32.bin"), 5),
        "exactbp": _i32(os.path.join(outdir, "exactbp_i32.bin"), 6),
        "support": _i32(os.path.join(outdir, "support_i32.bin"), 6),
    }
    p = os.path.join(outdir, "timings.json")
    d["timings"] = json.load(open(p)) if os.path.exists(p) else {}
    rl = os.path.join(outdir, "readlen.bin")
    d["read_len"] = int(np.fromfile(rl, dtype=np.int32)[0]) if os.path.exists(rl) else 0
    return d


def run(conc_sqmb: str, chim_sqmb: str, outdir: str, extra_args=(), timeout: float = 3600.0) -> dict:
    if not available():
        raise RuntimeError("oracle/_ref/squid_ref is missing (run oracle.pyref.build() where /root/reference exists)")
    os.makedirs(outdir, exist_ok=True)
    r = subprocess.run([REF_BIN, conc_sqmb, chim_sqmb, outdir, "--quiet", *extra_args], capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("reference harness failed rc=%d: %s" % (r.returncode, r.stderr[-2000:]))
    return load_dumps(outdir)


def exactbp_map(d: dict) -> dict:
    m = {}
    for row in d["exactbp"]:
        m.setdefault(tuple(int(v) for v in row[:4]), []).append((int(row[4]), int(row[5])))
    return m


def support_map(d: dict) -> dict:
    m = {}
    for row in d["support"]:
        m.setdefault(tuple(int(v) for v in row[:4]), []).append((int(row[4]), int(row[5])))
    return m
