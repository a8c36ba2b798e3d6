// oracle/ref_shim/boost/algorithm/string.hpp — TEST INFRASTRUCTURE ONLY
// (limbo/tools/macros.hpp:50 includes it for boost::replace_all in BO_PARAMS).
#pragma once
#include <string>
namespace boost {
inline void replace_all(std::string& s, const std::string& from, const std::string& to)
{
    size_t pos = 0;
    while ((pos = s.find(from, pos)) != std::string::npos) { s.replace(pos, from.size(), to); pos += to.size(); }
}
} // namespace boost
