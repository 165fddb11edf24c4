// TEST INFRASTRUCTURE (oracle build only) -- boost::algorithm::trim stand-in
// (genetics/individual/individual_genotype_file.h:35).
#ifndef ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP
#define ORACLE_SHIM_BOOST_ALGORITHM_STRING_HPP
#include <cctype>
#include <string>
namespace boost {
namespace algorithm {
inline void trim(std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) ++b;
    while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
    s = s.substr(b, e - b);
}
}
using algorithm::trim;
}
#endif
