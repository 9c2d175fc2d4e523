// stand-in, see oracle/shim/README.md: everything lives in adjacency_list.hpp
#include <algorithm>
#include <boost/graph/adjacency_list.hpp>
