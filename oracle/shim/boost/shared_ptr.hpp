// stand-in for <boost/shared_ptr.hpp> (PCL 1.8 clouds are boost::shared_ptr): see oracle/shim/README.md
#ifndef IFTWT_SHIM_BOOST_SHARED_PTR_HPP
#define IFTWT_SHIM_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
}
#endif
