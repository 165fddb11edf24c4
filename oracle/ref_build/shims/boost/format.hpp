// TEST INFRASTRUCTURE (oracle build only) -- printf-subset stand-in for Boost.Format (Boost is an
// un-vendored dependency of the reference: src/util/time/timing.h:35, algorithms/epistasis_func.h:45,
// used at algorithms/epistasis_func.cpp:502 with "%7d\t%7d\t%7d\t%f\t%f\t%f\t%f").
#ifndef ORACLE_SHIM_BOOST_FORMAT_HPP
#define ORACLE_SHIM_BOOST_FORMAT_HPP
#include <cstdio>
#include <ostream>
#include <string>
#include <vector>
namespace boost {
class format {
public:
    explicit format(const char *f) : fmt_(f) { split(); }
    explicit format(const std::string &f) : fmt_(f) { split(); }
    template <class T> format &operator%(const T &v) { feed(v); return *this; }
    std::string str() const {
        std::string s = out_;
        for (size_t i = next_; i < pieces_.size(); ++i) s += pieces_[i].lit;
        return s + tail_;
    }
private:
    struct piece { std::string lit, spec; };
    std::string fmt_, out_, tail_;
    std::vector<piece> pieces_;
    size_t next_ = 0;
    void split() {
        std::string lit;
        size_t i = 0;
        while (i < fmt_.size()) {
            if (fmt_[i] == '%' && i + 1 < fmt_.size() && fmt_[i + 1] == '%') { lit += '%'; i += 2; continue; }
            if (fmt_[i] == '%') {
                size_t j = i + 1;
                while (j < fmt_.size() && std::string("diuoxXfFeEgGsc").find(fmt_[j]) == std::string::npos) ++j;
                piece p; p.lit = lit; p.spec = fmt_.substr(i, j - i + 1);
                pieces_.push_back(p); lit.clear(); i = j + 1; continue;
            }
            lit += fmt_[i++];
        }
        tail_ = lit;
    }
    template <class T> void emit(const std::string &spec, T v) {
        char buf[512];
        std::snprintf(buf, sizeof buf, spec.c_str(), v);
        out_ += buf;
    }
    void one(const std::string &spec, double v) {
        char c = spec[spec.size() - 1];
        if (c == 'd' || c == 'i') emit(spec, (long long)v); else emit(spec, v);
    }
    void one(const std::string &spec, long long v) {
        char c = spec[spec.size() - 1];
        std::string s;
        for (size_t q = 0; q < spec.size(); ++q)   // drop C length modifiers; we always pass long long
            if (spec[q] != 'l' && spec[q] != 'h' && spec[q] != 'z') s += spec[q];
        if (c == 'd' || c == 'i' || c == 'u' || c == 'x' || c == 'X' || c == 'o') {
            s.insert(s.size() - 1, "ll"); emit(s, v);
        } else if (c == 's') { emit(std::string("%lld"), v); }
        else emit(spec, (double)v);
    }
    void one(const std::string &spec, const std::string &v) { (void)spec; out_ += v; }
    template <class T> void feed(const T &v) {
        if (next_ >= pieces_.size()) return;
        out_ += pieces_[next_].lit;
        dispatch(pieces_[next_].spec, v);
        ++next_;
    }
    void dispatch(const std::string &s, double v) { one(s, v); }
    void dispatch(const std::string &s, float v) { one(s, (double)v); }
    void dispatch(const std::string &s, int v) { one(s, (long long)v); }
    void dispatch(const std::string &s, unsigned v) { one(s, (long long)v); }
    void dispatch(const std::string &s, long v) { one(s, (long long)v); }
    void dispatch(const std::string &s, unsigned long v) { one(s, (long long)v); }
    void dispatch(const std::string &s, long long v) { one(s, v); }
    void dispatch(const std::string &s, const std::string &v) { one(s, v); }
    void dispatch(const std::string &s, const char *v) { one(s, std::string(v)); }
};
inline std::ostream &operator<<(std::ostream &o, const format &f) { return o << f.str(); }
inline std::string str(const format &f) { return f.str(); }
}
#endif
