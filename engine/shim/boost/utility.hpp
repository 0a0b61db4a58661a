/* boost::next stand-in (Book.cpp:101, Network.cpp:296). Build aid where boost is not installed. */
#pragma once
#include <iterator>
namespace boost { template <class It> It next(It it) { return std::next(it); } }
