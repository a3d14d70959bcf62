"""SQMB: uncompressed, mmap-able stand-in for a coordinate-sorted BAM (layout: oracle/shim/sqmb_format.h).

An `AlnTable` carries, per alignment record, exactly the BamAlignment members SQUID reads on the
segment-graph path (reference: src/ReadRec.cpp:10-88, src/SegmentGraph.cpp:297-314, 651-654, 3131-3155).
It is the input of the host packer (squid_b200.host: CIGAR -> aligned blocks, the twin of
ReadRec_t::ReadRec_t) and, written to disk, of the oracle's BamReader shim.
"""
from __future__ import annotations

import dataclasses
import numpy as np

CIGAR_OPS = "MIDNSHP=X"
OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X = range(9)

FLAG_PAIRED, FLAG_PROPER, FLAG_UNMAPPED, FLAG_MATE_UNMAPPED = 0x1, 0x2, 0x4, 0x8
FLAG_REVERSE, FLAG_MATE_REVERSE, FLAG_FIRST, FLAG_SECOND, FLAG_DUP = 0x10, 0x20, 0x40, 0x80, 0x400

AUX_XA, AUX_IH, AUX_NAME_SUFFIX = 1, 2, 4


@dataclasses.dataclass
class AlnTable:
    ref_len: np.ndarray  # int32[n_ref]
    ref_id: np.ndarray  # int32[n]
    pos: np.ndarray
    mate_ref_id: np.ndarray
    mate_pos: np.ndarray
    flag: np.ndarray  # uint16
    mapq: np.ndarray  # uint8
    aux: np.ndarray  # uint8 (AUX_*)
    ih: np.ndarray  # uint8
    polya: np.ndarray  # uint8
    lowrun: np.ndarray  # uint16
    name_id: np.ndarray  # uint64
    seq_off: np.ndarray  # int64 (-1 = synthesised bases/qualities)
    cigar_off: np.ndarray  # uint32[n+1]
    cigar: np.ndarray  # uint32
    blob: np.ndarray  # uint8

    @property
    def n(self) -> int:
        return int(self.ref_id.shape[0])

    def take(self, idx: np.ndarray) -> "AlnTable":
        """Rows `idx` in that order (CIGARs re-packed)."""
        idx = np.asarray(idx, dtype=np.int64)
        lens = (self.cigar_off[1:].astype(np.int64) - self.cigar_off[:-1].astype(np.int64))[idx]
        new_off = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        np.cumsum(lens, out=new_off[1:])
        src = np.repeat(self.cigar_off[:-1].astype(np.int64)[idx] - new_off[:-1], lens) + np.arange(int(new_off[-1]), dtype=np.int64)
        return AlnTable(
            ref_len=self.ref_len,
            ref_id=self.ref_id[idx], pos=self.pos[idx], mate_ref_id=self.mate_ref_id[idx], mate_pos=self.mate_pos[idx],
            flag=self.flag[idx], mapq=self.mapq[idx], aux=self.aux[idx], ih=self.ih[idx], polya=self.polya[idx],
            lowrun=self.lowrun[idx], name_id=self.name_id[idx], seq_off=self.seq_off[idx],
            cigar_off=new_off.astype(np.uint32), cigar=self.cigar[src] if src.size else np.zeros(0, np.uint32), blob=self.blob,
        )

    def sorted_by_coordinate(self) -> "AlnTable":
        key = (self.ref_id.astype(np.int64) << 32) | (self.pos.astype(np.int64) & 0xFFFFFFFF)
        # unmapped (ref_id -1) sort last, as samtools does
        key = np.where(self.ref_id < 0, np.int64(1) << 62, key)
        return self.take(np.argsort(key, kind="stable"))


def concat(tables: list[AlnTable]) -> AlnTable:
    offs = [0]
    for t in tables:
        offs.append(offs[-1] + int(t.cigar.shape[0]))
    cig_off = np.concatenate([t.cigar_off[:-1].astype(np.int64) + o for t, o in zip(tables, offs[:-1])] + [np.array([offs[-1]], np.int64)])
    assert all(t.blob.size == 0 for t in tables[1:]) or len(tables) == 1, "concat supports one blob only"
    cat = lambda f: np.concatenate([getattr(t, f) for t in tables])
    return AlnTable(
        ref_len=tables[0].ref_len, ref_id=cat("ref_id"), pos=cat("pos"), mate_ref_id=cat("mate_ref_id"), mate_pos=cat("mate_pos"),
        flag=cat("flag"), mapq=cat("mapq"), aux=cat("aux"), ih=cat("ih"), polya=cat("polya"), lowrun=cat("lowrun"),
        name_id=cat("name_id"), seq_off=cat("seq_off"), cigar_off=cig_off.astype(np.uint32), cigar=cat("cigar"), blob=tables[0].blob,
    )


def from_records(ref_len, recs: list[dict]) -> AlnTable:
    """Small explicit tables for known-answer tests.  Each rec: ref_id,pos,cigar(str),flag and optional
    mate_ref_id,mate_pos,mapq,xa,ih,name_id,name_suffix,seq,qual,lowrun,polya."""
    import re

    n = len(recs)
    t = empty(ref_len, n)
    cig, off, blob = [], [0], bytearray()
    for i, r in enumerate(recs):
        t.ref_id[i] = r["ref_id"]; t.pos[i] = r["pos"]
        t.mate_ref_id[i] = r.get("mate_ref_id", -1); t.mate_pos[i] = r.get("mate_pos", -1)
        t.flag[i] = r["flag"]; t.mapq[i] = r.get("mapq", 255)
        a = 0
        if r.get("xa"): a |= AUX_XA
        if "ih" in r: a |= AUX_IH; t.ih[i] = r["ih"]
        if r.get("name_suffix"): a |= AUX_NAME_SUFFIX
        t.aux[i] = a
        t.name_id[i] = r.get("name_id", i)
        t.lowrun[i] = r.get("lowrun", 0); t.polya[i] = r.get("polya", 0)
        for ln, op in re.findall(r"(\d+)([MIDNSHP=X])", r["cigar"]):
            cig.append((int(ln) << 4) | CIGAR_OPS.index(op))
        off.append(len(cig))
        if "seq" in r:
            s, q = r["seq"].encode(), r["qual"].encode()
            assert len(s) == len(q)
            t.seq_off[i] = len(blob)
            blob += np.uint32(len(s)).tobytes() + s + q
            blob += b"\0" * ((-len(blob)) % 4)
    t.cigar_off = np.array(off, dtype=np.uint32)
    t.cigar = np.array(cig, dtype=np.uint32)
    t.blob = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
    return t


def empty(ref_len, n: int) -> AlnTable:
    z = lambda dt: np.zeros(n, dtype=dt)
    return AlnTable(
        ref_len=np.asarray(ref_len, dtype=np.int32), ref_id=z(np.int32), pos=z(np.int32), mate_ref_id=z(np.int32) - 1, mate_pos=z(np.int32) - 1,
        flag=z(np.uint16), mapq=z(np.uint8), aux=z(np.uint8), ih=z(np.uint8), polya=z(np.uint8), lowrun=z(np.uint16),
        name_id=z(np.uint64), seq_off=z(np.int64) - 1, cigar_off=np.zeros(n + 1, np.uint32), cigar=np.zeros(0, np.uint32), blob=np.zeros(0, np.uint8),
    )


def write_sqmb(path: str, t: AlnTable) -> None:
    def pad(b: bytes) -> bytes:
        return b + b"\0" * ((-len(b)) % 8)

    with open(path, "wb") as f:
        f.write(b"SQMB0002")
        f.write(np.array([t.ref_len.shape[0], t.n, t.cigar.shape[0], t.blob.shape[0]], dtype=np.uint64).tobytes())
        for arr, dt in (
            (t.ref_len, np.int32), (t.ref_id, np.int32), (t.pos, np.int32), (t.mate_ref_id, np.int32), (t.mate_pos, np.int32),
            (t.flag, np.uint16), (t.mapq, np.uint8), (t.aux, np.uint8), (t.ih, np.uint8), (t.polya, np.uint8), (t.lowrun, np.uint16),
            (t.name_id, np.uint64), (t.seq_off, np.int64), (t.cigar_off, np.uint32), (t.cigar, np.uint32), (t.blob, np.uint8),
        ):
            f.write(pad(np.ascontiguousarray(arr, dtype=dt).tobytes()))
