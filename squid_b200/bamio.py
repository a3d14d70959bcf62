"""BAM / BGZF writer for AlnTable (tests and examples): the same alignment table that squid_b200.sqmb writes as SQMB, as a
real coordinate-sorted BAM file following the SAM/BAM specification.  The C++ front end (squid_b200/csrc/host/bam.cpp,
sqh_open_bam_case) reads it back; the two implementations share nothing but the specification, so a round trip
SQMB-packed == BAM-packed checks both.  Names are "q<name_id>[/1|/2]", references "chr<i>"; bases and qualities that an
AlnTable only summarises (polya / lowrun) are materialised exactly as the host decoder synthesises them."""
from __future__ import annotations

import struct
import zlib

import numpy as np

from . import sqmb

_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def _materialise(cig, polya: int, lowrun: int):
    """(seq, qual) strings for a record whose bases/qualities are summarised: 'C' everywhere, the k-th aligned block painted
    'A' (bit k) or 'T' (bit 4+k, wins); `lowrun` times '#', then 'I' (host/readrec.cpp decode_alignment)."""
    ops = [(int(c) & 15, int(c) >> 4) for c in cig]
    lseq = sum(l for o, l in ops if o in (sqmb.OP_M, sqmb.OP_I, sqmb.OP_S, sqmb.OP_EQ, sqmb.OP_X))
    seq = bytearray(b"C" * lseq)
    read_pos = hard = blk = 0
    i = 0
    while i < len(ops):
        o, l = ops[i]
        if o in (sqmb.OP_S, sqmb.OP_H):
            read_pos += l
            if o == sqmb.OP_H:
                hard += l
        elif o in (sqmb.OP_M, sqmb.OP_EQ):
            span = 0
            j = i
            while j < len(ops) and ops[j][0] not in (sqmb.OP_S, sqmb.OP_H, sqmb.OP_N):
                if ops[j][0] != sqmb.OP_D:
                    span += ops[j][1]
                j += 1
            if blk < 4:
                a, b = max(0, read_pos - hard), min(lseq, read_pos - hard + span)
                if (polya >> (4 + blk)) & 1:
                    seq[a:b] = b"T" * (b - a)
                elif (polya >> blk) & 1:
                    seq[a:b] = b"A" * (b - a)
            read_pos += span
            blk += 1
            i = j - 1
        i += 1
    low = min(lowrun, lseq)
    return bytes(seq), b"#" * low + b"I" * (lseq - low)


def bam_bytes(t: sqmb.AlnTable) -> bytes:
    return b"".join(bam_parts(t))


def bam_parts(t: sqmb.AlnTable) -> list:
    """[header bytes, record 0 bytes, record 1 bytes, ...] (each record with its block_size prefix)."""
    parts = []
    out = bytearray()
    text = b"@HD\tVN:1.6\tSO:coordinate\n" + b"".join(b"@SQ\tSN:chr%d\tLN:%d\n" % (i, int(l)) for i, l in enumerate(t.ref_len))
    out += b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", int(t.ref_len.shape[0]))
    for i, l in enumerate(t.ref_len):
        nm = b"chr%d\0" % i
        out += struct.pack("<i", len(nm)) + nm + struct.pack("<i", int(l))
    parts.append(bytes(out))
    coff = t.cigar_off.astype(np.int64)
    blob = t.blob.tobytes()
    for r in range(t.n):
        cig = t.cigar[coff[r]:coff[r + 1]]
        name = b"q%d" % int(t.name_id[r])
        if t.aux[r] & sqmb.AUX_NAME_SUFFIX:
            name += b"/2" if (int(t.flag[r]) & 0x80) else b"/1"
        name += b"\0"
        so = int(t.seq_off[r])
        if so >= 0:
            (l,) = struct.unpack_from("<I", blob, so)
            seq, qual = blob[so + 4:so + 4 + l], blob[so + 4 + l:so + 4 + 2 * l]
        else:
            seq, qual = _materialise(cig, int(t.polya[r]), int(t.lowrun[r]))
        lseq = len(seq)
        codes = [_SEQ_CODE.get(chr(c).upper(), 15) for c in seq] + [0]
        packed = bytes((codes[2 * k] << 4) | codes[2 * k + 1] for k in range((lseq + 1) // 2))
        phred = bytes((c - 33) & 0xFF for c in qual)
        aux = b""
        if t.aux[r] & sqmb.AUX_XA:
            aux += b"XAZchr1,+100,50M,0;\0"
        if t.aux[r] & sqmb.AUX_IH:  # every integer type BamTools' GetTag<int> accepts gets its turn
            v = int(t.ih[r])
            aux += [b"IHC" + struct.pack("<B", v), b"IHS" + struct.pack("<H", v), b"IHi" + struct.pack("<i", v), b"IHc" + struct.pack("<b", min(v, 127))][r % 4]
        aux += b"NHC\x01" + b"XSA+" + b"MDZ50\0" + b"ZBBS" + struct.pack("<i", 2) + struct.pack("<HH", 7, 9)  # tags the path ignores
        ref_span = sum(int(c) >> 4 for c in cig if (int(c) & 15) in (sqmb.OP_M, sqmb.OP_D, sqmb.OP_N, sqmb.OP_EQ, sqmb.OP_X))
        pos = int(t.pos[r])
        body = struct.pack("<iiBBHHHIiii", int(t.ref_id[r]), pos, len(name), int(t.mapq[r]), _reg2bin(max(pos, 0), max(pos, 0) + max(ref_span, 1)),
                           len(cig), int(t.flag[r]), lseq, int(t.mate_ref_id[r]), int(t.mate_pos[r]), 0)
        body += name + np.ascontiguousarray(cig, np.uint32).tobytes() + packed + phred + aux
        parts.append(struct.pack("<i", len(body)) + body)
    return parts


def bgzf_compress(data: bytes, block: int = 0xFF00, level: int = 6) -> bytes:
    out = bytearray()
    for o in range(0, len(data), block):
        chunk = data[o:o + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cd = c.compress(chunk) + c.flush()
        bsize = len(cd) + 25  # 12 header + 6 extra + cdata + 8 trailer, minus 1
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + cd + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))
    return bytes(out) + _BGZF_EOF


def _member(chunk: bytes, level: int) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    cd = c.compress(chunk) + c.flush()
    return b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cd) + 25) + cd + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))


def bgzf_compress_records(parts: list, block: int = 0xFF00, level: int = 6) -> bytes:
    """As htslib writes BAM: the header is flushed on its own and no record straddles two BGZF members (bgzf_flush_try) unless
    it is larger than a member."""
    out = bytearray()
    hdr = parts[0]
    for o in range(0, len(hdr), block):
        out += _member(hdr[o:o + block], level)
    cur = bytearray()
    for rec in parts[1:]:
        if cur and len(cur) + len(rec) > block:
            out += _member(bytes(cur), level)
            cur = bytearray()
        cur += rec
        while len(cur) > block:  # an oversized record spills over
            out += _member(bytes(cur[:block]), level)
            cur = cur[block:]
    if cur:
        out += _member(bytes(cur), level)
    return bytes(out) + _BGZF_EOF


def write_bam(path: str, t: sqmb.AlnTable, block: int = 0xFF00, level: int = 6, compressed: bool = True, htslib_blocks: bool = False) -> None:
    """htslib_blocks=True: BGZF members end at record boundaries (what samtools / aligners write); False: the stream is cut
    every `block` bytes wherever that falls (records straddle members -- legal BGZF, and the harder case for a reader)."""
    with open(path, "wb") as f:
        if not compressed:
            f.write(bam_bytes(t))
        elif htslib_blocks:
            f.write(bgzf_compress_records(bam_parts(t), block, level))
        else:
            f.write(bgzf_compress(bam_bytes(t), block, level))
