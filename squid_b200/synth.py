"""Deterministic synthetic RNA-seq alignment generator, numpy, BamAlignment level.

Produces a coordinate-sorted concordant AlnTable and a chimeric AlnTable.  It is used (a) by the
parity tests, where the same tables feed the reference-built oracle (as SQMB files) and the CUDA
path (through the host packer), and (b) by bench.py for the bounded CPU-baseline sample.  The
100 M-pair bench workload is generated directly on the GPU by squid_b200.synth_gpu with the same
statistical shape.

The model is transcriptome-like: gene loci with exons/introns, log-normal expression, 100-bp FR
pairs, spliced CIGARs, a few soft clips / low-quality runs / duplicates / multi-mappers, planted
fusions supported by split reads and discordant pairs, plus uniform noise chimeras.
"""
from __future__ import annotations

import numpy as np

from . import sqmb
from .sqmb import AlnTable

GRCH38_LEN = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
              133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
              58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
CHR17_LEN = [83257441]
READ_LEN = 100


class Transcriptome:
    def __init__(self, rng: np.random.Generator, ref_len, n_genes: int, exon_len=(100, 1500), intron_len=(200, 20000), n_exons=(4, 12)):
        ref_len = np.asarray(ref_len, dtype=np.int64)
        self.ref_len = ref_len
        usable = np.where(ref_len > 400000, ref_len, 0).astype(np.float64)
        if usable.sum() == 0:
            usable = ref_len.astype(np.float64)
        g_chr = rng.choice(len(ref_len), size=n_genes, p=usable / usable.sum())
        g_nex = rng.integers(n_exons[0], n_exons[1] + 1, size=n_genes)
        ex_gene = np.repeat(np.arange(n_genes), g_nex)
        ex_len = rng.integers(exon_len[0], exon_len[1] + 1, size=ex_gene.shape[0]).astype(np.int64)
        in_len = rng.integers(intron_len[0], intron_len[1] + 1, size=ex_gene.shape[0]).astype(np.int64)
        first = np.r_[True, ex_gene[1:] != ex_gene[:-1]]
        in_len[first] = 0  # "intron before exon k"
        step = ex_len + in_len
        cs = np.cumsum(step)
        g_first = np.flatnonzero(first)
        base = np.repeat(cs[g_first] - step[g_first], g_nex)
        ex_rel_end = cs - base  # end of exon k relative to gene start
        ex_rel_start = ex_rel_end - ex_len
        g_span = ex_rel_end[np.r_[g_first[1:] - 1, ex_gene.shape[0] - 1]]
        room = np.maximum(ref_len[g_chr] - g_span - 2000, 1)
        g_start = 1000 + (rng.random(n_genes) * room).astype(np.int64)
        self.g_chr = g_chr.astype(np.int64)
        self.g_first_exon = g_first
        self.g_nex = g_nex
        self.ex_gene = ex_gene
        self.ex_len = ex_len
        self.ex_gstart = np.repeat(g_start, g_nex) + ex_rel_start
        self.ex_tstart = np.cumsum(ex_len) - ex_len  # global transcript coordinate
        self.g_tstart = self.ex_tstart[g_first]
        self.g_tlen = np.add.reduceat(ex_len, g_first)
        ok = (g_start + g_span + 1000 < ref_len[g_chr])
        self.g_expr = np.exp(rng.normal(0.0, 2.0, size=n_genes)) * ok * (self.g_tlen >= 450)

    def map_interval(self, a: np.ndarray, b: np.ndarray, max_blocks: int = 6):
        """Global-transcript interval [a,b) -> genomic blocks.  Returns (nblk, gstart[n,max], glen[n,max])."""
        k0 = np.searchsorted(self.ex_tstart, a, side="right") - 1
        k1 = np.searchsorted(self.ex_tstart, b - 1, side="right") - 1
        nblk = (k1 - k0 + 1).astype(np.int64)
        assert nblk.max(initial=1) <= max_blocks and nblk.min(initial=1) >= 1
        gs = np.zeros((a.shape[0], max_blocks), dtype=np.int64)
        gl = np.zeros((a.shape[0], max_blocks), dtype=np.int64)
        for j in range(max_blocks):
            k = np.minimum(k0 + j, k1)
            lo = np.maximum(a, self.ex_tstart[k])
            hi = np.minimum(b, self.ex_tstart[k] + self.ex_len[k])
            valid = j < nblk
            gs[:, j] = np.where(valid, self.ex_gstart[k] + (lo - self.ex_tstart[k]), 0)
            gl[:, j] = np.where(valid, hi - lo, 0)
        return nblk, gs, gl


def _cigars_from_blocks(nblk, gs, gl, lclip, rclip):
    """[lclip S] M (N M)* [rclip S] for every row; returns (cigar_off uint32[n+1], cigar uint32)."""
    n = nblk.shape[0]
    nops = 2 * nblk - 1 + (lclip > 0) + (rclip > 0)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(nops, out=off[1:])
    cig = np.zeros(int(off[-1]), dtype=np.uint32)
    cur = off[:-1].copy()
    has = lclip > 0
    cig[cur[has]] = (lclip[has].astype(np.uint32) << 4) | sqmb.OP_S
    cur = cur + has
    for j in range(gs.shape[1]):
        v = j < nblk
        if not v.any():
            break
        if j > 0:
            gap = gs[:, j] - (gs[:, j - 1] + gl[:, j - 1])
            cig[cur[v]] = (gap[v].astype(np.uint32) << 4) | sqmb.OP_N
            cur = cur + v
        cig[cur[v]] = (gl[v, j].astype(np.uint32) << 4) | sqmb.OP_M
        cur = cur + v
    has = rclip > 0
    cig[cur[has]] = (rclip[has].astype(np.uint32) << 4) | sqmb.OP_S
    return off.astype(np.uint32), cig


def _table(ref_len, ref_id, pos, mate_ref_id, mate_pos, flag, name_id, cigar_off, cigar, mapq=None) -> AlnTable:
    n = ref_id.shape[0]
    t = sqmb.empty(ref_len, n)
    t.ref_id = ref_id.astype(np.int32); t.pos = pos.astype(np.int32)
    t.mate_ref_id = mate_ref_id.astype(np.int32); t.mate_pos = mate_pos.astype(np.int32)
    t.flag = flag.astype(np.uint16); t.name_id = name_id.astype(np.uint64)
    t.mapq = (np.full(n, 255) if mapq is None else mapq).astype(np.uint8)
    t.cigar_off = cigar_off; t.cigar = cigar
    return t


def _pairs_from_transcripts(rng, tx: Transcriptome, gene: np.ndarray, name_base: int, clip_frac: float, min_block: int = 4):
    """FR proper pairs sampled from transcripts: two records per pair (left = forward, right = reverse)."""
    n = gene.shape[0]
    ins = np.minimum(rng.integers(150, 401, size=n), tx.g_tlen[gene])
    t0 = tx.g_tstart[gene] + (rng.random(n) * (tx.g_tlen[gene] - ins + 1)).astype(np.int64)
    la, lb = t0, t0 + READ_LEN
    rb = t0 + ins
    ra = rb - READ_LEN
    # soft clips: shorten the aligned interval on one side
    lcl = np.zeros(n, np.int64); lcr = np.zeros(n, np.int64); rcl = np.zeros(n, np.int64); rcr = np.zeros(n, np.int64)
    c = rng.random(n) < clip_frac
    amt = rng.integers(16, 41, size=n)
    which = rng.integers(0, 4, size=n)
    lcl = np.where(c & (which == 0), amt, 0); lcr = np.where(c & (which == 1), amt, 0)
    rcl = np.where(c & (which == 2), amt, 0); rcr = np.where(c & (which == 3), amt, 0)
    ln, lgs, lgl = tx.map_interval(la + lcl, lb - lcr)
    rn, rgs, rgl = tx.map_interval(ra + rcl, rb - rcr)
    chr_ = tx.g_chr[gene]
    lpos, rpos = lgs[:, 0], rgs[:, 0]
    first_is_left = rng.random(n) < 0.5
    lflag = sqmb.FLAG_PAIRED | sqmb.FLAG_PROPER | sqmb.FLAG_MATE_REVERSE | np.where(first_is_left, sqmb.FLAG_FIRST, sqmb.FLAG_SECOND)
    rflag = sqmb.FLAG_PAIRED | sqmb.FLAG_PROPER | sqmb.FLAG_REVERSE | np.where(first_is_left, sqmb.FLAG_SECOND, sqmb.FLAG_FIRST)
    names = name_base + np.arange(n)
    loff, lcig = _cigars_from_blocks(ln, lgs, lgl, lcl, lcr)
    roff, rcig = _cigars_from_blocks(rn, rgs, rgl, rcl, rcr)
    left = _table(tx.ref_len, chr_, lpos, chr_, rpos, lflag, names, loff, lcig)
    right = _table(tx.ref_len, chr_, rpos, chr_, lpos, rflag, names, roff, rcig)
    if min_block > 1:
        # Pairs with an aligned block shorter than `min_block` are dropped.  Blocks of <= 3 bp make the
        # reference's per-segment Support depend on the tie order of an unstable std::sort
        # (SegmentGraph.cpp:781; DESIGN.md "known divergence"), so the default parity workload avoids them.
        big = np.int64(1 << 40)
        ok = (np.where(lgl > 0, lgl, big).min(axis=1) >= min_block) & (np.where(rgl > 0, rgl, big).min(axis=1) >= min_block)
        keep = np.flatnonzero(ok)
        left, right = left.take(keep), right.take(keep)
    return left, right


def _simple_cigar(n, lclip, m, rclip):
    """[lclip S] m M [rclip S]"""
    nblk = np.ones(n, np.int64)
    gs = np.zeros((n, 1), np.int64)
    gl = m.reshape(n, 1).astype(np.int64)
    return _cigars_from_blocks(nblk, gs, gl, lclip.astype(np.int64), rclip.astype(np.int64))


def make_chimeric(tx: Transcriptome, p: np.ndarray, n_pairs: int, seed: int, disc_frac: float, fusion_support: float = 20.0, adversarial: bool = True):
    """Chimeric AlnTable: planted fusions (split reads + discordant pairs), uniform noise chimeras and (adversarial)
    non-discordant chimeric reads that feed PartAlignPos / the 750 kb rule.  Own RNG stream (seed)."""
    rng = np.random.Generator(np.random.PCG64([seed, 0xC41]))
    ref_len = tx.ref_len
    n_genes = tx.g_chr.shape[0]
    # ---- chimeric reads -------------------------------------------------------------------------
    n_chim = int(round(disc_frac * n_pairs))
    n_fus = max(1, int(round(0.9 * n_chim / fusion_support))) if n_chim > 0 else 0
    chim_tabs = []
    fusions = []
    name_ctr = 1_000_000_000
    ex_w = p[tx.ex_gene] * (tx.ex_len >= 150)
    ex_w = ex_w / ex_w.sum()
    for f in range(n_fus):
        ea, eb = rng.choice(tx.ex_gene.shape[0], size=2, p=ex_w)
        chrA, chrB = int(tx.g_chr[tx.ex_gene[ea]]), int(tx.g_chr[tx.ex_gene[eb]])
        bpA = int(tx.ex_gstart[ea] + tx.ex_len[ea])  # reads lie left of bpA, forward
        inv = rng.random() < 0.4
        bpB = int(tx.ex_gstart[eb] + tx.ex_len[eb]) if inv else int(tx.ex_gstart[eb])  # inverted: reads lie left of bpB, reverse
        k = int(rng.poisson(fusion_support))
        if k == 0:
            continue
        fusions.append((chrA, bpA, chrB, bpB, inv, k))
        split = rng.random(k) < 0.6
        x = rng.integers(20, 81, size=k)  # bases of the split mate on the A side
        jit = np.where(rng.random(k) < 0.15, rng.integers(-2, 3, size=k), 0)  # alignment jitter at the junction
        off1 = rng.integers(0, 200, size=k); off2 = rng.integers(30, 250, size=k)
        names = name_ctr + np.arange(k); name_ctr += k
        first_is_A = rng.random(k) < 0.5
        fA = np.where(first_is_A, sqmb.FLAG_FIRST, sqmb.FLAG_SECOND) | sqmb.FLAG_PAIRED
        fB = np.where(first_is_A, sqmb.FLAG_SECOND, sqmb.FLAG_FIRST) | sqmb.FLAG_PAIRED
        ks = np.flatnonzero(split); kd = np.flatnonzero(~split)
        if ks.size:
            m = ks.size
            # split mate, A part: x M (100-x) S, forward, ends at bpA(+jit)
            posA = bpA + jit[ks] - x[ks]
            offA, cigA = _simple_cigar(m, np.zeros(m), x[ks], READ_LEN - x[ks])
            # split mate, B part
            if not inv:
                posB = np.full(m, bpB) + jit[ks]
                offB, cigB = _simple_cigar(m, x[ks], READ_LEN - x[ks], np.zeros(m))
                flagBpart = fA[ks]
                matepos = bpB + off2[ks] + READ_LEN
                mateflag = fB[ks] | sqmb.FLAG_REVERSE
            else:
                posB = bpB + jit[ks] - (READ_LEN - x[ks])
                offB, cigB = _simple_cigar(m, np.zeros(m), READ_LEN - x[ks], x[ks])
                flagBpart = fA[ks] | sqmb.FLAG_REVERSE
                matepos = bpB - off2[ks] - 2 * READ_LEN
                mateflag = fB[ks]
            offM, cigM = _simple_cigar(m, np.zeros(m), np.full(m, READ_LEN), np.zeros(m))
            cA = np.full(m, chrA); cB = np.full(m, chrB)
            chim_tabs.append(_table(ref_len, cA, posA, cB, matepos, fA[ks] | np.where(inv, 0, sqmb.FLAG_MATE_REVERSE), names[ks], offA, cigA))
            chim_tabs.append(_table(ref_len, cB, posB, cB, matepos, flagBpart | np.where(inv, 0, sqmb.FLAG_MATE_REVERSE), names[ks], offB, cigB))
            chim_tabs.append(_table(ref_len, cB, matepos, cA, posA, mateflag, names[ks], offM, cigM))
        if kd.size:
            m = kd.size
            posA = bpA - off1[kd] - READ_LEN
            if not inv:
                posB = bpB + off2[kd]
                flB = fB[kd] | sqmb.FLAG_REVERSE
                flA = fA[kd] | sqmb.FLAG_MATE_REVERSE
            else:
                posB = bpB - off2[kd] - READ_LEN
                flB = fB[kd]
                flA = fA[kd]
            offM, cigM = _simple_cigar(m, np.zeros(m), np.full(m, READ_LEN), np.zeros(m))
            cA = np.full(m, chrA); cB = np.full(m, chrB)
            chim_tabs.append(_table(ref_len, cA, posA, cB, posB, flA, names[kd], offM, cigM))
            chim_tabs.append(_table(ref_len, cB, posB, cA, posA, flB, names[kd], offM.copy(), cigM.copy()))
    # noise chimeras: two random 100M mates anywhere; plus (adversarial) non-discordant "chimeric" reads:
    # same chromosome, FR, with a soft clip > 15 (feeds PartAlignPos) or blocks > 750 kb apart
    n_noise = max(0, n_chim - sum(t.n for t in chim_tabs) // 2) if n_chim else 0
    n_noise = int(min(n_noise, max(2, 0.1 * n_chim))) if n_chim else 0
    if n_noise:
        m = n_noise
        w = ref_len / ref_len.sum()
        c1 = rng.choice(len(ref_len), size=m, p=w); c2 = rng.choice(len(ref_len), size=m, p=w)
        p1 = (rng.random(m) * (ref_len[c1] - 1000)).astype(np.int64) + 100
        p2 = (rng.random(m) * (ref_len[c2] - 1000)).astype(np.int64) + 100
        names = name_ctr + np.arange(m); name_ctr += m
        s1 = np.where(rng.random(m) < 0.5, sqmb.FLAG_REVERSE, 0); s2 = np.where(rng.random(m) < 0.5, sqmb.FLAG_REVERSE, 0)
        offM, cigM = _simple_cigar(m, np.zeros(m), np.full(m, READ_LEN), np.zeros(m))
        chim_tabs.append(_table(ref_len, c1, p1, c2, p2, sqmb.FLAG_PAIRED | sqmb.FLAG_FIRST | s1 | (s2 << 1), names, offM, cigM))
        chim_tabs.append(_table(ref_len, c2, p2, c1, p1, sqmb.FLAG_PAIRED | sqmb.FLAG_SECOND | s2 | (s1 << 1), names, offM.copy(), cigM.copy()))
    if adversarial and n_chim:
        m = max(2, n_chim // 20)
        g3 = rng.choice(n_genes, size=m, p=p)
        l3, r3 = _pairs_from_transcripts(rng, tx, g3, name_ctr, 1.0, 1)  # every pair has one clipped mate
        name_ctr += m
        chim_tabs += [l3, r3]
        m = max(2, n_chim // 40)
        g4 = rng.choice(n_genes, size=m, p=p)
        l4, r4 = _pairs_from_transcripts(rng, tx, g4, name_ctr, 0.0, 1)
        name_ctr += m
        # move the right mate > 750 kb away on the same chromosome, keep FR orientation
        np4 = np.minimum(r4.pos.astype(np.int64) + 800_000, ref_len[r4.ref_id] - 2000)
        r4.pos = np4.astype(np.int32); l4.mate_pos = r4.pos.copy()
        offM, cigM = _simple_cigar(m, np.zeros(m), np.full(m, READ_LEN), np.zeros(m))
        r4.cigar_off, r4.cigar = offM, cigM
        chim_tabs += [l4, r4]
    if chim_tabs:
        chim = sqmb.concat(chim_tabs)
        chim.pos = np.maximum(chim.pos, 0); chim.mate_pos = np.maximum(chim.mate_pos, 0)
        # shuffle: the chimeric file is not coordinate sorted in general
        chim = chim.take(rng.permutation(chim.n))
    else:
        chim = sqmb.empty(ref_len, 0)
    return chim, fusions


def make_case(n_pairs: int, ref_len=None, seed: int = 17, disc_frac: float = 0.005, n_genes: int | None = None,
              fusion_support: float = 20.0, clip_frac: float = 0.02, adversarial: bool = True, min_block: int = 4,
              exon_len=(100, 1500), intron_len=(200, 20000)):
    """Returns (concordant AlnTable sorted by coordinate, chimeric AlnTable, info dict).  Short exons make every read span
    several of them (many aligned blocks per record)."""
    ref_len = np.asarray(GRCH38_LEN if ref_len is None else ref_len, dtype=np.int64)
    rng = np.random.Generator(np.random.PCG64(seed))
    if n_genes is None:
        n_genes = int(max(8, min(20000, n_pairs // 400)))
    tx = Transcriptome(rng, ref_len, n_genes, exon_len=exon_len, intron_len=intron_len)
    p = tx.g_expr / tx.g_expr.sum()
    gene = rng.choice(n_genes, size=n_pairs, p=p)
    left, right = _pairs_from_transcripts(rng, tx, gene, 0, clip_frac, min_block)
    n = left.n
    parts = []
    if adversarial and n >= 50:
        # per-record decorations exercising the gate / low-phred / poly-A / duplicate rules
        for t in (left, right):
            r = rng.random(n)
            t.lowrun[r < 0.005] = 20
            r = rng.random(n)
            mm = r < 0.01
            t.mapq[mm] = 3; t.aux[mm] |= sqmb.AUX_IH; t.ih[mm] = 2
            r = rng.random(n)
            t.aux[r < 0.002] |= sqmb.AUX_XA
            r = rng.random(n)
            t.flag[r < 0.002] |= sqmb.FLAG_DUP
            r = rng.random(n)
            t.polya[r < 0.002] = 1
            r = rng.random(n)
            t.aux[(r < 0.002) & ((t.aux & sqmb.AUX_IH) == 0)] |= sqmb.AUX_IH  # IH:1 is kept
            t.ih[(t.aux & sqmb.AUX_IH != 0) & (t.ih == 0)] = 1
        # exact consecutive duplicates (different names)
        d = np.flatnonzero(rng.random(n) < 0.003)
        if d.size:
            dl, dr = left.take(d), right.take(d)
            dl.name_id = (np.uint64(500_000_000) + np.arange(d.size).astype(np.uint64)); dr.name_id = dl.name_id.copy()
            parts += [dl, dr]
        # mate-unmapped singletons
        s = rng.random(n) < 0.004
        left.flag[s] = (left.flag[s] | sqmb.FLAG_MATE_UNMAPPED) & ~np.uint16(sqmb.FLAG_PROPER | sqmb.FLAG_MATE_REVERSE)
        left.mate_pos[s] = left.pos[s]
        keep_right = ~s
        right = right.take(np.flatnonzero(keep_right))
        # long same-chromosome fragments (> -dp) and cross-chromosome pairs living in the concordant file
        m = max(4, n // 1000)
        g2 = rng.choice(n_genes, size=m, p=p)
        l2, r2 = _pairs_from_transcripts(rng, tx, g2, 600_000_000, 0.0, min_block)
        m = l2.n
        far = rng.integers(60_000, 900_000, size=m)
        newpos = np.minimum(l2.pos.astype(np.int64) + far, ref_len[l2.ref_id] - 200)
        off, cig = _simple_cigar(m, np.zeros(m), np.full(m, READ_LEN), np.zeros(m))
        r2 = _table(ref_len, l2.ref_id, newpos, l2.ref_id, l2.pos, r2.flag & ~np.uint16(sqmb.FLAG_PROPER), l2.name_id, off, cig)
        l2.mate_pos = r2.pos.copy(); l2.flag &= ~np.uint16(sqmb.FLAG_PROPER)
        x = rng.random(m) < 0.3  # a third of them land on another chromosome
        if len(ref_len) > 1:
            oc = (r2.ref_id + 1 + rng.integers(0, len(ref_len) - 1, size=m)) % len(ref_len)
            r2.ref_id = np.where(x, oc, r2.ref_id).astype(np.int32)
            r2.pos = np.where(x, np.minimum(r2.pos, ref_len[r2.ref_id] - 200), r2.pos).astype(np.int32)
            l2.mate_ref_id = r2.ref_id.copy(); l2.mate_pos = r2.pos.copy()
        parts += [l2, r2]

    chim, fusions = make_chimeric(tx, p, n_pairs, seed, disc_frac, fusion_support, adversarial)
    # 1 % of chimeric names also occur in the concordant file (ChimName gate, SegmentGraph.cpp:302);
    # half of those carry a /1,/2 suffix there, which defeats the gate
    if chim.n and adversarial:
        un = np.unique(chim.name_id)
        pick = un[rng.random(un.shape[0]) < 0.01]
        if pick.size < 2 and un.size >= 4:
            pick = un[:: max(1, un.size // 4)][:4]  # small cases: still exercise the gate
        if pick.size:
            m = pick.size
            g5 = rng.choice(n_genes, size=m, p=p)
            l5, r5 = _pairs_from_transcripts(rng, tx, g5, 0, 0.0, min_block)
            pick = pick[: l5.n]; m = l5.n
            l5.name_id = pick.copy(); r5.name_id = pick.copy()
            sfx = (np.arange(m) % 2) == 1  # every other one carries the suffix
            l5.aux[sfx] |= sqmb.AUX_NAME_SUFFIX; r5.aux[sfx] |= sqmb.AUX_NAME_SUFFIX
            parts += [l5, r5]
    conc = sqmb.concat([left, right] + parts).sorted_by_coordinate()
    info = {"n_pairs": n_pairs, "n_records": conc.n, "n_chim_records": chim.n, "fusions": fusions, "seed": seed}
    return conc, chim, info


def make_bwa_case(n_pairs: int, **kw):
    """BWA-style input (`squid --bwa -b all.bam`): ONE coordinate-sorted table that carries the
    concordant alignments and the split / discordant ones of make_case together, as bwa mem writes them (the chimeric parts
    keep their read names, which RawEdges joins on, SegmentGraph.cpp:1873-1926).  Returns (AlnTable, info)."""
    conc, chim, info = make_case(n_pairs, **kw)
    return sqmb.concat([conc, chim]).sorted_by_coordinate(), info
