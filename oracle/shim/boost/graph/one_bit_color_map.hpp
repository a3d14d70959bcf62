// TEST INFRASTRUCTURE: placeholder; the declarations the reference needs live in boost/graph/adjacency_list.hpp.
