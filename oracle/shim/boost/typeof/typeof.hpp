// TEST INFRASTRUCTURE: BOOST_AUTO as used at SegmentGraph.cpp:3324.
#ifndef SHIM_BOOST_TYPEOF_HPP
#define SHIM_BOOST_TYPEOF_HPP
#define BOOST_AUTO(var, expr) auto var = (expr)
#endif
