/* boost::algorithm stand-ins used by SGFTree.cpp:186-213. Build aid where boost is not installed. */
#pragma once
#include <string>
namespace boost { namespace algorithm {
inline bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }
inline bool find_first(const std::string& s, const std::string& p) { return s.find(p) != std::string::npos; }
inline void trim(std::string& s) {
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    s = (b == std::string::npos) ? std::string() : s.substr(b, e - b + 1);
}
} using algorithm::starts_with; using algorithm::find_first; using algorithm::trim; }
