// TEST INFRASTRUCTURE (oracle side): API-compatible stand-in for BamTools 2.4.0 `BamAlignment`,
// exposing only the members the reference reads on the segment-graph path.
// Semantics follow BamTools 2.4.0's published behaviour (source not available here):
//   GetEndPosition() = Position + sum(len of M,D,N,=,X), half-open; flag bits per the SAM spec.
#ifndef SHIM_BAMALIGNMENT_H
#define SHIM_BAMALIGNMENT_H
#include <cstdint>
#include <string>
#include <vector>
namespace BamTools {
struct CigarOp {
    char Type;
    uint32_t Length;
    CigarOp(char t = '\0', uint32_t l = 0) : Type(t), Length(l) {}
};
struct BamAlignment {
    std::string Name, QueryBases, Qualities;
    std::vector<CigarOp> CigarData;
    int32_t RefID = -1, Position = -1, MateRefID = -1, MatePosition = -1;
    uint16_t MapQuality = 0;
    uint32_t AlignmentFlag = 0;
    bool TagXA = false, TagIH = false;
    int TagIHValue = 0;
    bool IsPaired() const { return AlignmentFlag & 0x1; }
    bool IsProperPair() const { return AlignmentFlag & 0x2; }
    bool IsMapped() const { return !(AlignmentFlag & 0x4); }
    bool IsMateMapped() const { return !(AlignmentFlag & 0x8); }
    bool IsReverseStrand() const { return AlignmentFlag & 0x10; }
    bool IsMateReverseStrand() const { return AlignmentFlag & 0x20; }
    bool IsFirstMate() const { return AlignmentFlag & 0x40; }
    bool IsSecondMate() const { return AlignmentFlag & 0x80; }
    bool IsDuplicate() const { return AlignmentFlag & 0x400; }
    bool HasTag(const std::string &t) const {
        if (t == "XA") return TagXA;
        if (t == "IH") return TagIH;
        return false;
    }
    template <typename T> bool GetTag(const std::string &t, T &dst) const {
        if (t == "IH" && TagIH) { dst = (T)TagIHValue; return true; }
        return false;
    }
    int GetEndPosition(bool usePadded = false, bool closedInterval = false) const {
        int e = Position;
        for (const CigarOp &c : CigarData)
            if (c.Type == 'M' || c.Type == 'D' || c.Type == 'N' || c.Type == '=' || c.Type == 'X') e += (int)c.Length;
            else if (usePadded && c.Type == 'I') e += (int)c.Length;
        return closedInterval ? e - 1 : e;
    }
};
}  // namespace BamTools
#endif
