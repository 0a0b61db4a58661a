/* printf-style boost::format stand-in covering the reference's uses
 * (Network.cpp:855, Book.cpp:100, SGFTree.cpp:414-472): one directive per operator%.
 * Build aid for compiling the reference's sources where boost is not installed (engine/Makefile and
 * oracle/ref/Makefile); a build with real boost headers does not need it. */
#pragma once
#include <cstdio>
#include <ostream>
#include <string>
namespace boost {
class format {
    std::string m_fmt, m_out;
    size_t m_pos = 0;
    void flush_literal() {
        while (m_pos < m_fmt.size()) {
            if (m_fmt[m_pos] == '%') {
                if (m_pos + 1 < m_fmt.size() && m_fmt[m_pos + 1] == '%') { m_out += '%'; m_pos += 2; continue; }
                return;
            }
            m_out += m_fmt[m_pos++];
        }
    }
    std::string take_directive() {
        flush_literal();
        size_t b = m_pos;
        if (b >= m_fmt.size()) return std::string();
        size_t e = b + 1;
        while (e < m_fmt.size() && !std::isalpha((unsigned char)m_fmt[e])) e++;
        if (e < m_fmt.size()) e++;
        m_pos = e;
        return m_fmt.substr(b, e - b);
    }
    template <class T> format& put(const char* lenmod, T v) {
        std::string d = take_directive();
        if (d.empty()) return *this;
        char conv = d.back();
        std::string spec = d.substr(0, d.size() - 1) + lenmod + conv;
        char buf[128];
        std::snprintf(buf, sizeof buf, spec.c_str(), v);
        m_out += buf; flush_literal();
        return *this;
    }
public:
    explicit format(const std::string& f) : m_fmt(f) { flush_literal(); }
    format& operator%(int v) { return put("", v); }
    format& operator%(unsigned v) { return put("", v); }
    format& operator%(long v) { return put("l", v); }
    format& operator%(unsigned long v) { return put("l", v); }
    format& operator%(long long v) { return put("ll", v); }
    format& operator%(unsigned long long v) { return put("ll", v); }
    format& operator%(double v) { return put("", v); }
    format& operator%(float v) { return put("", (double)v); }
    std::string str() const { return m_out; }
};
inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}
