// .NET number formatting as the reference's text output needs it (the host layer stands in for a
// C# host, so it must print what the BCL would print):
//   ToString("0.###")          15-significant-digit decimal image, then half-up at 3 decimals,
//                              trailing zeros dropped, negative zero keeps its sign
//   ToString("F3"/"F6")        correctly rounded fixed notation
//   Math.Round(v, d)           round-half-even of v * 10^d, divided back
//   double.ToString()          shortest round-trip, scientific above 1e15 / below 1e-5
// Written independently of the test oracle (integer arithmetic on the 15-digit significand).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>

namespace lpr381 {
namespace text {

inline const char* inf_symbol(bool neg) { return neg ? "-\xE2\x88\x9E" : "\xE2\x88\x9E"; }

// value = sig * 10^(exp10 - 14) with sig a 15-digit integer (or 0)
inline void significand15(double av, uint64_t& sig, int& exp10) {
    if (av == 0) {
        sig = 0;
        exp10 = 0;
        return;
    }
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof buf, av, std::chars_format::scientific, 14);
    *r.ptr = 0;
    sig = 0;
    const char* p = buf;
    for (; *p && *p != 'e'; ++p)
        if (*p != '.') sig = sig * 10 + (uint64_t)(*p - '0');
    exp10 = std::atoi(p + 1);
}

inline std::string custom_hash(double v, int decimals = 3) {  // "0.###"
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return inf_symbol(v < 0);
    uint64_t sig;
    int e;
    significand15(std::fabs(v), sig, e);
    // keep `decimals` digits after the point: drop k = 14 - e - decimals low digits, half-up
    int drop = 14 - e - decimals;
    uint64_t q = sig;  // integer count of 10^-decimals units
    if (drop > 15) q = 0;
    else if (drop > 0) {
        uint64_t p10 = 1;
        for (int i = 0; i < drop; i++) p10 *= 10;
        const uint64_t rem = sig % p10;
        q = sig / p10 + ((rem * 2 >= p10) ? 1 : 0);
    }
    // q in units of 10^-decimals, as a digit string (for drop < 0 the 15 digits are followed by zeros)
    std::string digits = drop >= 0 ? std::to_string(q) : std::to_string(sig) + std::string((size_t)(-drop), '0');
    if ((int)digits.size() <= decimals) digits.insert(0, (size_t)(decimals + 1 - (int)digits.size()), '0');
    std::string ip = digits.substr(0, digits.size() - decimals);
    std::string fp = digits.substr(digits.size() - decimals);
    while (!fp.empty() && fp.back() == '0') fp.pop_back();
    std::string out = std::signbit(v) ? "-" : "";
    out += ip;
    if (!fp.empty()) out += "." + fp;
    return out;
}

inline double math_round(double v, int digits) {
    static const double p[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    if (std::fabs(v) < 1e16) {
        v *= p[digits];
        v = std::nearbyint(v);
        v /= p[digits];
    }
    return v;
}

inline std::string round_trip(double v) {  // double.ToString()
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return inf_symbol(v < 0);
    if (v == 0) return std::signbit(v) ? "-0" : "0";
    char buf[48];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    *r.ptr = 0;
    std::string dig;
    const char* p = buf;
    for (; *p != 'e'; ++p)
        if (*p != '.') dig.push_back(*p);
    const int e = std::atoi(p + 1);
    while (dig.size() > 1 && dig.back() == '0') dig.pop_back();
    std::string out = std::signbit(v) ? "-" : "";
    if (e >= 15 || e < -4) {
        out += dig.substr(0, 1);
        if (dig.size() > 1) out += "." + dig.substr(1);
        char eb[16];
        std::snprintf(eb, sizeof eb, "E%c%02d", e < 0 ? '-' : '+', e < 0 ? -e : e);
        return out + eb;
    }
    if (e < 0) return out + "0." + std::string((size_t)(-e - 1), '0') + dig;
    if ((int)dig.size() <= e + 1) return out + dig + std::string((size_t)(e + 1 - (int)dig.size()), '0');
    return out + dig.substr(0, (size_t)e + 1) + "." + dig.substr((size_t)e + 1);
}

inline std::string fixed(double v, int k) {  // ToString("F<k>", InvariantCulture)
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed, k);
    return std::string(buf, r.ptr);
}

inline std::string pad_left(const std::string& s, int width) {
    int units = 0;
    for (unsigned char ch : s)
        if ((ch & 0xC0) != 0x80) units++;
    return units >= width ? s : std::string((size_t)(width - units), ' ') + s;
}

}  // namespace text
}  // namespace lpr381
