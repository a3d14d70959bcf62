// TEST INFRASTRUCTURE (oracle side): API-compatible stand-in for BamTools 2.4.0 `BamReader`,
// serving alignments from an in-memory SQMB file (sqmb_format.h) so that BGZF inflate is
// excluded on the reference side exactly as it is on the GPU side.
#ifndef SHIM_BAMREADER_H
#define SHIM_BAMREADER_H
#include <cassert>
#include <string>
#include <vector>
#include "BamAlignment.h"
#include "sqmb_format.h"
namespace BamTools {
struct SamSequence {
    std::string Name, Length;
};
typedef std::vector<SamSequence>::iterator SamSequenceIterator;
struct SamSequenceDictionary {
    std::vector<SamSequence> v;
    SamSequenceIterator Begin() { return v.begin(); }
    SamSequenceIterator End() { return v.end(); }
};
struct SamHeader {
    SamSequenceDictionary Sequences;
};
class BamReader {
    SqmbView view;
    bool opened = false;
    uint64_t cur = 0;

public:
    bool Open(const std::string &fn) {
        opened = view.open(fn);
        cur = 0;
        return opened;
    }
    bool IsOpen() const { return opened; }
    bool Close() {
        view.close();
        opened = false;
        return true;
    }
    SamHeader GetHeader() const {
        SamHeader h;
        for (uint64_t i = 0; i < view.n_ref; i++) h.Sequences.v.push_back({"chr" + std::to_string(i), std::to_string(view.ref_len[i])});
        return h;
    }
    bool GetNextAlignment(BamAlignment &a) {
        if (!opened || cur >= view.n_rec) return false;
        const uint64_t r = cur++;
        static const char OPS[] = "MIDNSHP=X";
        a.RefID = view.ref_id[r];
        a.Position = view.pos[r];
        a.MateRefID = view.mate_ref_id[r];
        a.MatePosition = view.mate_pos[r];
        a.AlignmentFlag = view.flag[r];
        a.MapQuality = view.mapq[r];
        a.TagXA = view.aux[r] & 1;
        a.TagIH = view.aux[r] & 2;
        a.TagIHValue = view.ih[r];
        a.Name = "q" + std::to_string(view.name_id[r]);
        if (view.aux[r] & 4) a.Name += (a.AlignmentFlag & 0x80) ? "/2" : "/1";
        a.CigarData.clear();
        uint32_t lseq = 0;
        for (uint32_t c = view.cigar_off[r]; c < view.cigar_off[r + 1]; c++) {
            uint32_t op = view.cigar[c] & 15, len = view.cigar[c] >> 4;
            assert(op < 9);
            a.CigarData.push_back(CigarOp(OPS[op], len));
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) lseq += len;
        }
        if (view.seq_off[r] >= 0) {
            const uint8_t *b = view.blob + view.seq_off[r];
            uint32_t l;
            memcpy(&l, b, 4);
            a.QueryBases.assign((const char *)b + 4, l);
            a.Qualities.assign((const char *)b + 4 + l, l);
        } else {
            a.QueryBases.assign(lseq, 'C');
            a.Qualities.assign(lseq, 'I');
            uint32_t lr = view.lowrun[r] < lseq ? view.lowrun[r] : lseq;
            for (uint32_t i = 0; i < lr; i++) a.Qualities[i] = '#';
            if (view.polya[r]) {
                // locate the read span of each top-level aligned block (M or = opens a block that runs
                // until S, H or N; inside it every op but D consumes read) and paint it A or T
                int rp = 0, blk = 0;  // rp: offset in QueryBases (hard clips are not in SEQ)
                for (size_t c = 0; c < a.CigarData.size(); c++) {
                    char t = a.CigarData[c].Type;
                    if (t == 'S') rp += a.CigarData[c].Length;
                    else if (t == 'M' || t == '=') {
                        int span = 0;
                        size_t d = c;
                        for (; d < a.CigarData.size() && a.CigarData[d].Type != 'S' && a.CigarData[d].Type != 'H' && a.CigarData[d].Type != 'N'; d++)
                            if (a.CigarData[d].Type != 'D') span += a.CigarData[d].Length;
                        if (blk < 4 && (view.polya[r] >> blk & 1))
                            for (int i = rp; i < rp + span && i < (int)lseq; i++) a.QueryBases[i] = 'A';
                        if (blk < 4 && (view.polya[r] >> (4 + blk) & 1))
                            for (int i = rp; i < rp + span && i < (int)lseq; i++) a.QueryBases[i] = 'T';
                        rp += span;
                        blk++;
                        c = d - 1;
                    }
                }
            }
        }
        return true;
    }
};
}  // namespace BamTools
#endif
