#include "bam.h"

#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sqh {
namespace {
inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }

struct Member { size_t cdata, clen, off; uint32_t isize; };  // deflate payload [cdata, cdata+clen) -> out[off, off+isize)

// size in bytes of one aux value of type `t` starting at p (p points behind the type byte); 0 on malformed input
size_t aux_size(char t, const uint8_t *p, const uint8_t *end) {
    switch (t) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'Z': case 'H': { const uint8_t *q = p; while (q < end && *q) q++; return q < end ? (size_t)(q - p) + 1 : 0; }
        case 'B': {
            if (p + 5 > end) return 0;
            const char st = (char)p[0];
            const uint32_t cnt = rd32(p + 1);
            const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
            return es ? 5 + (size_t)cnt * es : 0;
        }
        default: return 0;
    }
}
}  // namespace

bool bgzf_inflate_all(const uint8_t *d, size_t len, std::vector<uint8_t> &out, std::string &err, int threads) {
    std::vector<Member> mem;
    size_t o = 0, total = 0;
    while (o < len) {
        if (o + 18 > len || d[o] != 0x1f || d[o + 1] != 0x8b || d[o + 2] != 8 || !(d[o + 3] & 4)) { err = "not a BGZF member at offset " + std::to_string(o); return false; }
        const uint16_t xlen = rd16(d + o + 10);
        size_t x = o + 12, xe = x + xlen;
        if (xe > len) { err = "truncated BGZF header"; return false; }
        int64_t bsize = -1;
        while (x + 4 <= xe) {  // extra subfields: SI1 SI2 SLEN data
            const uint16_t slen = rd16(d + x + 2);
            if (d[x] == 'B' && d[x + 1] == 'C' && slen == 2 && x + 6 <= xe) bsize = rd16(d + x + 4);
            x += 4 + slen;
        }
        if (bsize < 0) { err = "gzip member without a BC subfield (plain gzip is not BGZF)"; return false; }
        const size_t msize = (size_t)bsize + 1;
        if (o + msize > len || msize < (size_t)xlen + 20) { err = "truncated BGZF block"; return false; }
        Member m;
        m.cdata = o + 12 + xlen; m.clen = msize - xlen - 20; m.isize = rd32(d + o + msize - 4); m.off = total;
        if (m.isize > 65536) { err = "BGZF block larger than 64 KiB"; return false; }
        total += m.isize;
        mem.push_back(m);
        o += msize;
    }
    out.resize(total);
    const int T = threads > 0 ? threads : omp_get_num_procs();
    int bad = 0;
#pragma omp parallel for num_threads(T) schedule(dynamic, 16)
    for (long long i = 0; i < (long long)mem.size(); i++) {
        const Member &m = mem[(size_t)i];
        if (m.isize == 0) continue;  // the EOF marker
        z_stream z;
        memset(&z, 0, sizeof(z));
        if (inflateInit2(&z, -15) != Z_OK) { bad = 1; continue; }
        z.next_in = const_cast<Bytef *>(d + m.cdata); z.avail_in = (uInt)m.clen;
        z.next_out = out.data() + m.off; z.avail_out = m.isize;
        const int rc = inflate(&z, Z_FINISH);
        if (rc != Z_STREAM_END || z.avail_out != 0) bad = 1;
        else if (crc32(crc32(0L, Z_NULL, 0), out.data() + m.off, m.isize) != rd32(d + m.cdata + m.clen)) bad = 2;
        inflateEnd(&z);
    }
    if (bad) { err = bad == 2 ? "BGZF block fails its CRC32" : "corrupt deflate stream in a BGZF block"; return false; }
    return true;
}

bool BamTable::open(const std::string &path, std::string &err, int threads) {
    *this = BamTable();
    const bool timing = getenv("SQH_TIMING") != nullptr;
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *w) { if (timing) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[bam] %s %.1f ms\n", w, 1e3 * std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 4) { ::close(fd); err = path + " is empty"; return false; }
    const size_t flen = (size_t)st.st_size;
    void *map = mmap(nullptr, flen, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) { err = "cannot map " + path; return false; }
    const uint8_t *f = (const uint8_t *)map;
    std::vector<uint8_t> raw;
    const uint8_t *b; size_t n;
    if (memcmp(f, "BAM\1", 4) == 0) { b = f; n = flen; }  // uncompressed BAM stream
    else {
        if (!bgzf_inflate_all(f, flen, raw, err, threads)) { munmap(map, flen); return false; }
        b = raw.data(); n = raw.size();
    }
    lap("inflate");
    bool ok = false;
    do {
        if (n < 12 || memcmp(b, "BAM\1", 4) != 0) { err = "not a BAM file (magic)"; break; }
        size_t o = 4;
        const int32_t l_text = rdi32(b + o); o += 4;
        if (l_text < 0 || o + (size_t)l_text + 4 > n) { err = "truncated BAM header"; break; }
        o += (size_t)l_text;
        const int32_t n_ref = rdi32(b + o); o += 4;
        bool hdr_ok = n_ref >= 0;
        for (int32_t i = 0; hdr_ok && i < n_ref; i++) {
            if (o + 4 > n) { hdr_ok = false; break; }
            const int32_t l_name = rdi32(b + o); o += 4;
            if (l_name < 1 || o + (size_t)l_name + 4 > n) { hdr_ok = false; break; }
            ref_name.emplace_back((const char *)b + o, (size_t)l_name - 1); o += (size_t)l_name;
            ref_len.push_back(rdi32(b + o)); o += 4;
        }
        if (!hdr_ok) { err = "truncated BAM reference list"; break; }
        // record boundaries (sequential hop over block_size), then the fields in parallel
        std::vector<size_t> rec;
        bool rec_ok = true;
        while (o < n) {
            if (o + 4 > n) { rec_ok = false; break; }
            const int32_t bs = rdi32(b + o);
            if (bs < 32 || o + 4 + (size_t)bs > n) { rec_ok = false; break; }
            rec.push_back(o + 4);
            o += 4 + (size_t)bs;
        }
        if (!rec_ok) { err = "truncated BAM alignment record"; break; }
        lap("record boundaries");
        const size_t R = rec.size();
        ref_id.resize(R); pos.resize(R); mate_ref_id.resize(R); mate_pos.resize(R); ih.assign(R, 0); flag.resize(R); mapq.resize(R); tags.assign(R, 0);
        name_off.assign(R + 1, 0); cigar_off.assign(R + 1, 0); seq_off.assign(R + 1, 0);
        for (size_t r = 0; r < R; r++) {
            const uint8_t *p = b + rec[r];
            const uint32_t l_name = p[8], n_cig = rd16(p + 12), l_seq = rd32(p + 16);
            name_off[r + 1] = name_off[r] + (l_name ? l_name - 1 : 0);
            cigar_off[r + 1] = cigar_off[r] + n_cig;
            seq_off[r + 1] = seq_off[r] + l_seq;
        }
        if (!names.resize(name_off[R]) || !cigar.resize(cigar_off[R]) || !seq.resize(seq_off[R]) || !qual.resize(seq_off[R])) { err = "out of memory"; break; }
        const int T = threads > 0 ? threads : omp_get_num_procs();
        int bad = 0;
        lap("offsets + allocation");
#pragma omp parallel for num_threads(T) schedule(static)
        for (long long rr = 0; rr < (long long)R; rr++) {
            const size_t r = (size_t)rr;
            const uint8_t *p = b + rec[r];
            const size_t bs = (size_t)rdi32(p - 4);
            const uint8_t *end = p + bs;
            ref_id[r] = rdi32(p); pos[r] = rdi32(p + 4);
            const uint32_t l_name = p[8];
            mapq[r] = p[9];
            const uint32_t n_cig = rd16(p + 12);
            flag[r] = rd16(p + 14);
            const uint32_t l_seq = rd32(p + 16);
            mate_ref_id[r] = rdi32(p + 20); mate_pos[r] = rdi32(p + 24);
            const uint8_t *q = p + 32;
            if (q + l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq > end) { bad = 1; continue; }
            if (l_name) memcpy(names.data() + name_off[r], q, l_name - 1);
            q += l_name;
            memcpy(cigar.data() + cigar_off[r], q, 4ull * n_cig);
            q += 4ull * n_cig;
            static const char code[] = "=ACMGRSVTWYHKDBN";
            char *sq = seq.data() + seq_off[r], *ql = qual.data() + seq_off[r];
            for (uint32_t k = 0; k < l_seq; k++) sq[k] = code[(q[k >> 1] >> ((k & 1) ? 0 : 4)) & 15];
            q += (l_seq + 1) / 2;
            for (uint32_t k = 0; k < l_seq; k++) ql[k] = (char)(uint8_t)(q[k] + 33);
            q += l_seq;
            while (q + 3 <= end) {  // aux: tag[2] type value
                const char t0 = (char)q[0], t1 = (char)q[1], ty = (char)q[2];
                const size_t sz = aux_size(ty, q + 3, end);
                if (sz == 0 || q + 3 + sz > end) { bad = 1; break; }
                if (t0 == 'X' && t1 == 'A') tags[r] |= 1;
                else if (t0 == 'I' && t1 == 'H') {
                    tags[r] |= 2;
                    const uint8_t *v = q + 3;
                    switch (ty) {
                        case 'c': ih[r] = (int8_t)v[0]; break;
                        case 'C': ih[r] = v[0]; break;
                        case 's': ih[r] = (int16_t)rd16(v); break;
                        case 'S': ih[r] = rd16(v); break;
                        case 'i': ih[r] = rdi32(v); break;
                        case 'I': ih[r] = (int32_t)rd32(v); break;
                        default: break;  // GetTag<int> fails on a non-integer tag: the value stays 0
                    }
                }
                q += 3 + sz;
            }
        }
        if (bad) { err = "malformed BAM alignment record"; break; }
        lap("fields");
        ok = true;
    } while (false);
    munmap(map, flen);
    if (!ok) *this = BamTable();
    return ok;
}

AlnSource BamTable::source() const {
    AlnSource s;
    s.n_rec = n_rec();
    const BamTable *t = this;
    s.at = [t](uint64_t r) {
        Alignment a;
        a.ref_id = t->ref_id[r]; a.pos = t->pos[r]; a.mate_ref_id = t->mate_ref_id[r]; a.mate_pos = t->mate_pos[r];
        a.flag = t->flag[r]; a.mapq = t->mapq[r];
        a.tag_xa = t->tags[r] & 1; a.tag_ih = t->tags[r] & 2; a.ih_value = t->ih[r];
        a.cigar = t->cigar.data() + t->cigar_off[r]; a.n_cigar = (uint32_t)(t->cigar_off[r + 1] - t->cigar_off[r]);
        a.l_seq = (uint32_t)(t->seq_off[r + 1] - t->seq_off[r]);
        a.seq = t->seq.data() + t->seq_off[r]; a.qual = t->qual.data() + t->seq_off[r];
        return a;
    };
    s.name = [t](uint64_t r) { return std::string(t->names.data() + t->name_off[r], (size_t)(t->name_off[r + 1] - t->name_off[r])); };
    return s;
}

}  // namespace sqh
