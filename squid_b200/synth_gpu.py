"""Benchmark-scale synthetic workload (SURVEY.md App. C shape), generated directly as a struct-of-arrays record
batch with torch on the GPU: 100 M read pairs would take minutes (and tens of GB) in numpy.

Same transcriptome model as squid_b200.synth (the gene/exon tables are built by the numpy code and shared), same
read shapes (100-bp FR proper pairs, spliced blocks, 2 % soft clips, 0.5 % low-quality runs, 1 % multi-mappers,
0.3 % exact duplicates).  The chimeric reads (0.5 % of pairs) are small and come from squid_b200.synth.make_chimeric
through the host loader, exactly as in the parity tests.  torch is used here as plumbing (device memory, sort, RNG).
"""
from __future__ import annotations

import numpy as np
import torch

from . import synth

READ_LEN = synth.READ_LEN


def _map_interval(tx_t, a, b, max_blocks):
    ex_tstart, ex_len, ex_gstart = tx_t["ex_tstart"], tx_t["ex_len"], tx_t["ex_gstart"]
    k0 = torch.searchsorted(ex_tstart, a, right=True) - 1
    k1 = torch.searchsorted(ex_tstart, b - 1, right=True) - 1
    nblk = (k1 - k0 + 1).clamp_(max=max_blocks)
    gs, gl = [], []
    for j in range(max_blocks):
        k = torch.minimum(k0 + j, k1)
        lo = torch.maximum(a, ex_tstart[k])
        hi = torch.minimum(b, ex_tstart[k] + ex_len[k])
        valid = j < nblk
        gs.append(torch.where(valid, ex_gstart[k] + (lo - ex_tstart[k]), torch.zeros_like(lo)).to(torch.int32))
        gl.append(torch.where(valid, hi - lo, torch.zeros_like(lo)).to(torch.int32))
    return nblk.to(torch.int32), gs, gl


def make_bench_batch(n_pairs: int, seed: int = 100, device: str = "cuda", ref_len=None, n_genes: int = 20000, max_blocks: int = 3,
                     exon_len=(60, 1500), min_block: int = 4):
    """Returns (dict of device tensors in sqg_batch layout, Transcriptome, expression probabilities).
    Pairs with an aligned block shorter than `min_block` are dropped (as in synth.make_case: aligners do not emit 1-3 bp
    overhangs by default, and such blocks make the reference's Support depend on an unstable sort), so the
    batch holds slightly fewer than n_pairs pairs; use ref_id.shape[0] // 2."""
    ref_len = np.asarray(synth.GRCH38_LEN if ref_len is None else ref_len, dtype=np.int64)
    rng = np.random.Generator(np.random.PCG64(seed))
    tx = synth.Transcriptome(rng, ref_len, n_genes, exon_len=exon_len)
    p = tx.g_expr / tx.g_expr.sum()
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    T = lambda x, dt=torch.int64: torch.as_tensor(np.ascontiguousarray(x), device=dev).to(dt)
    tx_t = {"ex_tstart": T(tx.ex_tstart), "ex_len": T(tx.ex_len), "ex_gstart": T(tx.ex_gstart)}
    g_tstart, g_tlen, g_chr = T(tx.g_tstart), T(tx.g_tlen), T(tx.g_chr, torch.int32)
    cum = torch.as_tensor(np.cumsum(p), device=dev, dtype=torch.float64)
    n = n_pairs
    # 0.3 % exact duplicates: draw n0 pairs, then repeat some of them
    n_dup = int(0.003 * n)
    n0 = n - n_dup
    gene = torch.searchsorted(cum, torch.rand(n0, generator=g, device=dev, dtype=torch.float64)).clamp_(max=n_genes - 1)
    tlen = g_tlen[gene]
    ins = torch.minimum(torch.randint(150, 401, (n0,), generator=g, device=dev), tlen)
    t0 = g_tstart[gene] + (torch.rand(n0, generator=g, device=dev, dtype=torch.float64) * (tlen - ins + 1).to(torch.float64)).to(torch.int64)
    chr_ = g_chr[gene]
    del tlen
    clip = torch.rand(n0, generator=g, device=dev) < 0.02
    amt = torch.randint(16, 41, (n0,), generator=g, device=dev)
    which = torch.randint(0, 4, (n0,), generator=g, device=dev)
    z = torch.zeros_like(amt)
    lcl = torch.where(clip & (which == 0), amt, z); lcr = torch.where(clip & (which == 1), amt, z)
    rcl = torch.where(clip & (which == 2), amt, z); rcr = torch.where(clip & (which == 3), amt, z)
    del clip, amt, which
    first_left = torch.rand(n0, generator=g, device=dev) < 0.5
    if n_dup:
        pick = torch.randint(0, n0, (n_dup,), generator=g, device=dev)
        cat = lambda x: torch.cat([x, x[pick]])
        t0, ins, chr_, lcl, lcr, rcl, rcr, first_left = map(cat, (t0, ins, chr_, lcl, lcr, rcl, rcr, first_left))
    # left (forward) and right (reverse) records
    ln, lgs, lgl = _map_interval(tx_t, t0 + lcl, t0 + READ_LEN - lcr, max_blocks)
    rn, rgs, rgl = _map_interval(tx_t, t0 + ins - READ_LEN + rcl, t0 + ins - rcr, max_blocks)
    del t0, ins
    if min_block > 1:
        big = 1 << 30
        ok = torch.ones_like(ln, dtype=torch.bool)
        for j in range(max_blocks):
            ok &= torch.where(ln > j, lgl[j], torch.full_like(lgl[j], big)) >= min_block
            ok &= torch.where(rn > j, rgl[j], torch.full_like(rgl[j], big)) >= min_block
        keep = torch.nonzero(ok).squeeze(1)
        sel = lambda x: x[keep]
        ln, rn, chr_, lcl, lcr, rcl, rcr, first_left = map(sel, (ln, rn, chr_, lcl, lcr, rcl, rcr, first_left))
        lgs = [x[keep] for x in lgs]; lgl = [x[keep] for x in lgl]; rgs = [x[keep] for x in rgs]; rgl = [x[keep] for x in rgl]
        n = int(keep.shape[0])
        del ok, keep
    FP, FPR, FR, FMR, F1, F2 = 0x1, 0x2, 0x10, 0x20, 0x40, 0x80
    lflag = torch.where(first_left, FP | FPR | FMR | F1, FP | FPR | FMR | F2).to(torch.int32)
    rflag = torch.where(first_left, FP | FPR | FR | F2, FP | FPR | FR | F1).to(torch.int32)
    del first_left

    def rec(nblk, gs, gl, cl, cr, flag, mate_pos, reverse):
        pos = gs[0]
        last = torch.zeros_like(pos)
        for j in range(max_blocks):
            last = torch.where(nblk > j, gs[j] + gl[j], last)
        return {"ref_id": chr_, "pos": pos, "mate_ref_id": chr_, "mate_pos": mate_pos, "end_pos": last, "flag": flag, "nblk": nblk,
                "gs": gs, "gl": gl, "cl": cl.to(torch.int32), "cr": cr.to(torch.int32), "rev": reverse}

    L = rec(ln, lgs, lgl, lcl, lcr, lflag, rgs[0], False)
    R = rec(rn, rgs, rgl, rcl, rcr, rflag, lgs[0], True)
    del lcl, lcr, rcl, rcr
    cat2 = lambda k: torch.cat([L[k], R[k]])
    ref_id, pos = cat2("ref_id"), cat2("pos")
    key = (ref_id.to(torch.int64) << 32) | pos.to(torch.int64)
    order = torch.sort(key, stable=True).indices
    del key
    N = 2 * n
    out = {}
    out["ref_id"] = ref_id[order]; out["pos"] = pos[order]
    del ref_id, pos
    out["mate_ref_id"] = cat2("mate_ref_id")[order]; out["mate_pos"] = cat2("mate_pos")[order]; out["end_pos"] = cat2("end_pos")[order]
    flag = cat2("flag")[order]
    nblk = cat2("nblk")[order]
    cl = cat2("cl")[order]
    rev = torch.cat([torch.zeros(n, dtype=torch.bool, device=dev), torch.ones(n, dtype=torch.bool, device=dev)])[order]
    # decorations of the gate / low-phred rules
    mapq = torch.full((N,), 255, dtype=torch.int32, device=dev)
    aux = torch.zeros(N, dtype=torch.int32, device=dev)
    u = torch.rand(N, generator=g, device=dev)
    mm = u < 0.01
    mapq = torch.where(mm, torch.full_like(mapq, 3), mapq); aux = torch.where(mm, aux | 2, aux)
    aux = torch.where((u >= 0.01) & (u < 0.012), aux | 1, aux)
    flag = torch.where((u >= 0.012) & (u < 0.014), flag | 0x400, flag)
    lowrun = torch.where((u >= 0.014) & (u < 0.019), torch.full_like(aux, 20), torch.zeros_like(aux))
    out["flag"] = flag.to(torch.int16); out["mapq"] = mapq.to(torch.uint8); out["aux"] = aux.to(torch.uint8)
    out["lowphred_run"] = lowrun.to(torch.int16)
    out["total_len"] = torch.full((N,), READ_LEN, dtype=torch.int16, device=dev)
    del u, mm, mapq, aux, lowrun, flag
    blk_off = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(nblk.to(torch.int64), 0, out=blk_off[1:])
    nb = int(blk_off[-1].item())
    b_ref_pos = torch.empty(nb, dtype=torch.int32, device=dev); b_match_ref = torch.empty(nb, dtype=torch.int32, device=dev)
    b_read_pos = torch.empty(nb, dtype=torch.int16, device=dev); b_match_read = torch.empty(nb, dtype=torch.int16, device=dev)
    rp = cl.clone()  # read offset of the current block in sequencing order of the CIGAR
    for j in range(max_blocks):
        gsj = torch.cat([L["gs"][j], R["gs"][j]])[order]
        glj = torch.cat([L["gl"][j], R["gl"][j]])[order]
        v = nblk > j
        idx = (blk_off[:-1] + j)[v]
        b_ref_pos[idx] = gsj[v]; b_match_ref[idx] = glj[v]; b_match_read[idx] = glj[v].to(torch.int16)
        rpos = torch.where(rev, READ_LEN - rp - glj, rp)  # ReadRec.cpp:74-75
        b_read_pos[idx] = rpos[v].to(torch.int16)
        rp = rp + glj
        del gsj, glj, v, idx, rpos
    out["blk_off"] = blk_off.to(torch.int32)
    out["blk_ref_pos"] = b_ref_pos; out["blk_match_ref"] = b_match_ref; out["blk_read_pos"] = b_read_pos; out["blk_match_read"] = b_match_read
    out = {k: v.contiguous() for k, v in out.items()}
    return out, tx, p


def batch_struct(t: dict):
    """sqg_batch over device (or pinned host) tensors."""
    from . import api
    s = api.sqg_batch()
    s.n_rec = int(t["ref_id"].shape[0]); s.n_blk = int(t["blk_ref_pos"].shape[0])
    for k in api.BATCH_DTYPES:
        setattr(s, k, t[k].data_ptr())
    return s


def batch_bytes(t: dict) -> int:
    return int(sum(v.numel() * v.element_size() for v in t.values()))


def to_alntable(t: dict, ref_len) -> "synth.AlnTable":
    """SoA batch (any device) -> alignment-level table, so that the same workload can be written as SQMB and fed to
    the oracle (used by the tests to cross-check the benchmark generator at reduced size)."""
    from . import sqmb
    a = {k: v.detach().cpu().numpy() for k, v in t.items()}
    n = a["ref_id"].shape[0]
    off = a["blk_off"].astype(np.int64)
    nblk = np.diff(off)
    rev = (a["flag"].astype(np.int64) & 0x10) != 0
    tot = a["total_len"].astype(np.int64)
    first = off[:-1]
    mlen = a["blk_match_read"].astype(np.int64)
    rpos = a["blk_read_pos"].astype(np.int64)
    aligned = np.add.reduceat(mlen, first) if n else np.zeros(0, np.int64)
    lc = np.where(rev, tot - rpos[first] - mlen[first], rpos[first])
    rc = tot - lc - aligned
    maxb = int(nblk.max(initial=1))
    gs = np.zeros((n, maxb), np.int64); gl = np.zeros((n, maxb), np.int64)
    for j in range(maxb):
        v = nblk > j
        gs[v, j] = a["blk_ref_pos"][first[v] + j]; gl[v, j] = a["blk_match_ref"][first[v] + j]
    coff, cig = synth._cigars_from_blocks(nblk, gs, gl, lc, rc)
    tb = sqmb.empty(ref_len, n)
    tb.ref_id = a["ref_id"].astype(np.int32); tb.pos = a["pos"].astype(np.int32)
    tb.mate_ref_id = a["mate_ref_id"].astype(np.int32); tb.mate_pos = a["mate_pos"].astype(np.int32)
    tb.flag = a["flag"].astype(np.uint16); tb.mapq = a["mapq"].astype(np.uint8)
    ax = a["aux"].astype(np.uint8)
    tb.aux = ((ax & 1) | (ax & 2)).astype(np.uint8); tb.ih = np.where(ax & 2, 2, 0).astype(np.uint8)
    tb.lowrun = a["lowphred_run"].astype(np.uint16)
    tb.name_id = np.arange(n, dtype=np.uint64)
    tb.cigar_off, tb.cigar = coff, cig
    return tb
