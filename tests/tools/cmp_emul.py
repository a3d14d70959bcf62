"""Dev tool: run oracle/_ref/squid_ref and tests/emul on a synthetic case and compare every seam."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from squid_b200 import synth, sqmb

def load(d, name, cols):
    p = os.path.join(d, name)
    return np.fromfile(p, dtype=np.int32).reshape(-1, cols) if os.path.exists(p) else None

def run_case(n_pairs, seed, disc_frac, ref_len, tmp, verbose=True, **kw):
    os.makedirs(tmp + "/ref", exist_ok=True); os.makedirs(tmp + "/emu", exist_ok=True)
    conc, chim, info = synth.make_case(n_pairs, ref_len=ref_len, seed=seed, disc_frac=disc_frac, **kw)
    sqmb.write_sqmb(tmp + "/conc.sqmb", conc); sqmb.write_sqmb(tmp + "/chim.sqmb", chim)
    t = time.time()
    r = subprocess.run(["oracle/_ref/squid_ref", tmp + "/conc.sqmb", tmp + "/chim.sqmb", tmp + "/ref", "--quiet"], capture_output=True, text=True)
    t_ref = time.time() - t
    if r.returncode != 0:
        return "ref-fail rc=%d %s" % (r.returncode, r.stderr[-300:])
    # BPs as ExactBPConcordantSupport assembles them (SegmentGraph.cpp:3091-3109)
    fn = load(tmp + "/ref", "final_nodes_i32.bin", 4).astype(np.int64); fe = load(tmp + "/ref", "final_edges_i32.bin", 5); xb = load(tmp + "/ref", "exactbp_i32.bin", 6)
    bps = []
    xmap = {}
    for row in xb:
        xmap.setdefault(tuple(int(v) for v in row[:4]), []).append((int(row[4]), int(row[5])))
    for e in fe:
        k = tuple(int(v) for v in e[:4])
        if k in xmap:
            for b1, b2 in xmap[k]:
                bps.append((fn[e[0], 0], b1)); bps.append((fn[e[1], 0], b2))
        else:
            bps.append((fn[e[0], 0], fn[e[0], 1] + (0 if e[2] else fn[e[0], 2])))
            bps.append((fn[e[1], 0], fn[e[1], 1] + (0 if e[3] else fn[e[1], 2])))
    bps.sort()
    np.array(bps, dtype=np.int32).reshape(-1, 2).tofile(tmp + "/bps.bin")
    t = time.time()
    r = subprocess.run(["tests/emul/_build/emul", tmp + "/conc.sqmb", tmp + "/chim.sqmb", tmp + "/emu", tmp + "/bps.bin"], capture_output=True, text=True)
    t_emu = time.time() - t
    if r.returncode != 0:
        return "emu-fail rc=%d %s" % (r.returncode, r.stderr[-300:])
    msgs = []
    for name, cols in (("nodes_i32.bin", 4), ("edges_i32.bin", 5), ("chim_after_edges.bin", 8)):
        a, b = load(tmp + "/ref", name, cols), load(tmp + "/emu", name, cols)
        if a.shape != b.shape or not np.array_equal(a, b):
            msgs.append("%s differs: ref %s emu %s" % (name, a.shape, b.shape))
            if verbose and a.shape == b.shape:
                bad = np.flatnonzero((a != b).any(axis=1))[:5]
                for i in bad: msgs.append("   row %d ref %s emu %s" % (i, a[i], b[i]))
            elif verbose:
                sa = set(map(tuple, a[:, :3])); sb = set(map(tuple, b[:, :3]))
                msgs.append("   only ref: %s" % sorted(sa - sb)[:6]); msgs.append("   only emu: %s" % sorted(sb - sa)[:6])
    a = np.fromfile(tmp + "/ref/nodes_f64.bin"); b = np.fromfile(tmp + "/emu/nodes_f64.bin")
    if a.shape != b.shape or not np.array_equal(a, b): msgs.append("avgdepth differs")
    # coverage: support map rows are (edge, cov1, cov2) in map order with bp lookup by lower_bound
    sup = load(tmp + "/ref", "support_i32.bin", 6); cov = np.fromfile(tmp + "/emu/cov_i32.bin", dtype=np.int32)
    bpa = np.array(bps, dtype=np.int64).reshape(-1, 2)
    bkey = bpa[:, 0] * (1 << 32) + bpa[:, 1]
    # rebuild expected pairs in the same order as the support dump (map<Edge_t> order = sorted edge key; per edge the ExactBP order)
    exp = []
    for e in sorted(tuple(int(v) for v in x) for x in fe[:, :4]):
        pairs = xmap.get(e)
        if pairs:
            for b1, b2 in pairs:
                exp.append((fn[e[0], 0] * (1 << 32) + b1, fn[e[1], 0] * (1 << 32) + b2))
        else:
            exp.append((fn[e[0], 0] * (1 << 32) + fn[e[0], 1] + (0 if e[2] else fn[e[0], 2]), fn[e[1], 0] * (1 << 32) + fn[e[1], 1] + (0 if e[3] else fn[e[1], 2])))
    got = np.array([(cov[np.searchsorted(bkey, k1)], cov[np.searchsorted(bkey, k2)]) for k1, k2 in exp], dtype=np.int32).reshape(-1, 2)
    if got.shape != sup[:, 4:].shape or not np.array_equal(got, sup[:, 4:]):
        msgs.append("coverage differs (%d of %d)" % ((got != sup[:, 4:]).sum() if got.shape == sup[:, 4:].shape else -1, got.size))
    st = "OK" if not msgs else "MISMATCH\n" + "\n".join(msgs)
    return "%s  [pairs=%d recs=%d chim=%d nodes=%d edges=%d bps=%d ref %.2fs emu %.2fs]" % (st, n_pairs, conc.n, chim.n, len(load(tmp + "/ref", "nodes_i32.bin", 4)), len(load(tmp + "/ref", "edges_i32.bin", 5)), len(bps), t_ref, t_emu)

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 17
    d = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
    ref = synth.CHR17_LEN if (len(sys.argv) <= 4 or sys.argv[4] == "chr17") else synth.GRCH38_LEN
    print(run_case(n, seed, d, ref, "/tmp/cmp_%d_%d" % (n, seed)))
