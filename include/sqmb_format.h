// SQMB = an uncompressed, memory-mappable stand-in for a coordinate-sorted BAM: the alignment-level
// input format of the host packer (squid_b200/csrc/host) and of the test oracle's BamReader shim.
// "SQMB" = an uncompressed, memory-mappable stand-in for a coordinate-sorted BAM.
// It carries exactly the BamAlignment members the reference reads on the
// segment-graph path (BamTools row), so the reference's own sources can be
// fed without BamTools/BGZF.  Written by squid_b200/sqmb.py (numpy), read by the
// BamReader shim (api/BamReader.h) through this header.
//
// Layout (little endian, every section 8-byte aligned):
//   char     magic[8]   = "SQMB0002"
//   uint64   n_ref, n_rec, n_cigar, blob_bytes
//   int32    ref_len[n_ref]                      reference i is named "chr<i>"
//   int32    ref_id[n_rec], pos[n_rec], mate_ref_id[n_rec], mate_pos[n_rec]
//   uint16   flag[n_rec]                         BAM FLAG bits
//   uint8    mapq[n_rec]
//   uint8    aux[n_rec]                          bit0: XA tag present, bit1: IH tag present (value in ih[]),
//                                                bit2: Name carries a "/1" or "/2" suffix (by mate flag)
//   uint8    ih[n_rec]                           IH tag value (when aux bit1)
//   uint8    polya[n_rec]                        bit k (k<4): k-th aligned CIGAR block is all 'A'; bit 4+k: all 'T'
//   uint16   lowrun[n_rec]                       length of a low-quality run synthesised at the start of Qualities
//   uint64   name_id[n_rec]                      Name = "q<name_id>" (+suffix)
//   int64    seq_off[n_rec]                      -1: QueryBases/Qualities are synthesised; else offset into blob of
//                                                uint32 l_seq, then l_seq base chars, then l_seq quality chars
//   uint32   cigar_off[n_rec+1]
//   uint32   cigar[n_cigar]                      BAM encoding: len<<4 | op, op index into "MIDNSHP=X"
//   uint8    blob[blob_bytes]
#ifndef SQMB_FORMAT_H
#define SQMB_FORMAT_H
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

struct SqmbView {
    uint64_t n_ref = 0, n_rec = 0, n_cigar = 0, blob_bytes = 0;
    const int32_t *ref_len = nullptr, *ref_id = nullptr, *pos = nullptr, *mate_ref_id = nullptr, *mate_pos = nullptr;
    const uint16_t *flag = nullptr;
    const uint8_t *mapq = nullptr, *aux = nullptr, *ih = nullptr, *polya = nullptr;
    const uint16_t *lowrun = nullptr;
    const uint64_t *name_id = nullptr;
    const int64_t *seq_off = nullptr;
    const uint32_t *cigar_off = nullptr, *cigar = nullptr;
    const uint8_t *blob = nullptr;
    void *map_base = nullptr;
    size_t map_len = 0;

    static size_t pad8(size_t x) { return (x + 7) & ~size_t(7); }

    bool open(const std::string &path) {
        close();
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 40) { ::close(fd); return false; }
        map_len = (size_t)st.st_size;
        map_base = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (map_base == MAP_FAILED) { map_base = nullptr; return false; }
        const uint8_t *p = (const uint8_t *)map_base;
        if (memcmp(p, "SQMB0002", 8) != 0) { close(); return false; }
        const uint64_t *h = (const uint64_t *)(p + 8);
        n_ref = h[0]; n_rec = h[1]; n_cigar = h[2]; blob_bytes = h[3];
        size_t o = 40;
        auto take = [&](size_t bytes) { const uint8_t *q = p + o; o += pad8(bytes); return q; };
        ref_len = (const int32_t *)take(4 * n_ref);
        ref_id = (const int32_t *)take(4 * n_rec);
        pos = (const int32_t *)take(4 * n_rec);
        mate_ref_id = (const int32_t *)take(4 * n_rec);
        mate_pos = (const int32_t *)take(4 * n_rec);
        flag = (const uint16_t *)take(2 * n_rec);
        mapq = (const uint8_t *)take(n_rec);
        aux = (const uint8_t *)take(n_rec);
        ih = (const uint8_t *)take(n_rec);
        polya = (const uint8_t *)take(n_rec);
        lowrun = (const uint16_t *)take(2 * n_rec);
        name_id = (const uint64_t *)take(8 * n_rec);
        seq_off = (const int64_t *)take(8 * n_rec);
        cigar_off = (const uint32_t *)take(4 * (n_rec + 1));
        cigar = (const uint32_t *)take(4 * n_cigar);
        blob = (const uint8_t *)take(blob_bytes);
        if (o > map_len) { close(); return false; }
        return true;
    }
    void close() {
        if (map_base) munmap(map_base, map_len);
        map_base = nullptr; map_len = 0; n_rec = 0;
    }
    ~SqmbView() { close(); }
};
#endif
