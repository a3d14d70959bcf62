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


RESTATE_LIB = os.path.join(HERE, "liboracle.so")


def build_restate() -> None:
    """Compiles the CPU restatement (oracle/restate/squid_oracle.cpp -> oracle/liboracle.so)."""
    src = os.path.join(HERE, "restate", "squid_oracle.cpp")
    if os.path.exists(RESTATE_LIB) and os.path.getmtime(RESTATE_LIB) >= os.path.getmtime(src):
        return
    r = subprocess.run(["make", "-C", HERE, "restate"], capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(RESTATE_LIB):
        raise RuntimeError("oracle restatement build failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])


def run_restate(conc_sqmb: str, chim_sqmb: str, outdir: str, bps=None, opts=None, stop_after: int = 0) -> dict:
    """Runs the CPU restatement; `bps` = sorted (chr,pos) int array for the coverage pass.  Same dump layout as run()."""
    import ctypes as C
    build_restate()
    lib = C.CDLL(RESTATE_LIB)
    lib.sqo_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.c_char_p, C.c_int]
    os.makedirs(outdir, exist_ok=True)
    bf = None
    if bps is not None:
        bf = os.path.join(outdir, "bps_in.bin")
        np.ascontiguousarray(bps, dtype=np.int32).reshape(-1, 2).tofile(bf)
    o = (C.c_int * 6)(*(opts if opts is not None else (1, 10, 4, -1, 50000, 20)))
    rc = lib.sqo_run(conc_sqmb.encode(), chim_sqmb.encode(), outdir.encode(), o, bf.encode() if bf else None, stop_after)
    if rc != 0:
        raise RuntimeError("sqo_run failed rc=%d" % rc)
    d = load_dumps(outdir)
    cp = os.path.join(outdir, "cov_i32.bin")
    d["cov"] = np.fromfile(cp, dtype=np.int32) if os.path.exists(cp) else None
    return d


def breakpoints_of(d: dict) -> np.ndarray:
    """The sorted BPs vector ExactBPConcordantSupport assembles (SegmentGraph.cpp:3091-3109) from a dump of the
    reference harness (final graph + ExactBP map)."""
    fn = d["final_nodes"].astype(np.int64)
    xm = exactbp_map(d)
    bps = []
    for e in d["final_edges"]:
        k = tuple(int(v) for v in e[:4])
        if k in xm:
            for b1, b2 in xm[k]:
                bps.append((int(fn[k[0], 0]), b1)); bps.append((int(fn[k[1], 0]), b2))
        else:
            bps.append((int(fn[k[0], 0]), int(fn[k[0], 1] + (0 if k[2] else fn[k[0], 2]))))
            bps.append((int(fn[k[1], 0]), int(fn[k[1], 1] + (0 if k[3] else fn[k[1], 2]))))
    bps.sort()
    return np.array(bps, dtype=np.int32).reshape(-1, 2)


def support_from_cov(d: dict, bps: np.ndarray, cov: np.ndarray) -> dict:
    """Maps a Coverages vector back to {edge: [(cov1,cov2)]} as SegmentGraph.cpp:3171-3211 does (lower_bound lookups)."""
    fn = d["final_nodes"].astype(np.int64)
    xm = exactbp_map(d)
    key = bps[:, 0].astype(np.int64) * (1 << 32) + bps[:, 1].astype(np.int64)
    look = lambda c, p: int(cov[np.searchsorted(key, int(c) * (1 << 32) + int(p))])
    out = {}
    for e in d["final_edges"]:
        k = tuple(int(v) for v in e[:4])
        if k in xm:
            out[k] = [(look(fn[k[0], 0], b1), look(fn[k[1], 0], b2)) for b1, b2 in xm[k]]
        else:
            out[k] = [(look(fn[k[0], 0], fn[k[0], 1] + (0 if k[2] else fn[k[0], 2])), look(fn[k[1], 0], fn[k[1], 1] + (0 if k[3] else fn[k[1], 2])))]
    return out


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
