// TEST INFRASTRUCTURE (oracle build only) -- Boost.Unordered stand-in
// (genetics/individual/individual_collection.h:35).
#ifndef ORACLE_SHIM_BOOST_UNORDERED_MAP_HPP
#define ORACLE_SHIM_BOOST_UNORDERED_MAP_HPP
#include <unordered_map>
namespace boost { using std::unordered_map; }
#endif
