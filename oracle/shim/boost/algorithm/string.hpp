// TEST INFRASTRUCTURE (oracle side): the two Boost.StringAlgo entry points the reference calls
// (boost::split with boost::is_any_of; ReadRec.cpp:295, SegmentGraph.cpp:134, WriteIO.cpp:22), off the hot path.
#ifndef SHIM_BOOST_STRING_HPP
#define SHIM_BOOST_STRING_HPP
#include <string>
#include <vector>
namespace boost {
struct is_any_of_t {
    std::string set;
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline is_any_of_t is_any_of(const std::string &s) { return is_any_of_t{s}; }
template <class Seq, class Pred> Seq &split(Seq &out, const std::string &in, Pred p) {
    out.clear();
    std::string cur;
    for (char c : in) {
        if (p(c)) { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}
}  // namespace boost
#endif
