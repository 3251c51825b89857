// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, link or call anything under oracle/.
//
// Restatement of the .NET 8 number-formatting behaviour the reference relies on
// (it has no formatting code of its own; it calls the BCL):
//   d.ToString("0.###")                R/Models/PrimalSimplex.cs:280, DualSimplex.cs:255,
//                                      BranchAndBoundKnapsack.cs:93,142,214
//   Math.Round(x, k)                   R/Models/PrimalSimplex.cs:140-146
//   double.ToString() (shortest R/T)   R/Models/PrimalSimplex.cs:146 (string.Join over doubles)
//   d.ToString("F3"/"F6", Invariant)   R/Models/Branch&Bound.cs:40,45,81,172,206
// PARITY UNPINNED: no .NET runtime exists in this environment, so these rules are written
// from knowledge of the .NET Core 3.0+ formatter (15-significant-digit Dragon4 buffer followed
// by half-up decimal rounding for custom formats; correctly-rounded "F"; negative zero keeps
// its sign).  They cannot be checked against the real runtime here.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <charconv>

namespace orc {

// Decimal digit buffer "0.DIGITS x 10^scale", the shape of .NET's NumberBuffer.
struct DigitBuf {
    std::string dig;   // no leading zeros, no trailing zeros
    int scale = 0;
    bool neg = false;
};

// Double -> 15 significant digits, correctly rounded (what Dragon4Double(value, 15, true) yields).
inline DigitBuf digits15(double v) {
    DigitBuf d;
    d.neg = std::signbit(v);
    if (v == 0.0) return d;
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.14e", std::fabs(v));
    d.dig.push_back(buf[0]);
    d.dig.append(buf + 2, 14);
    d.scale = std::atoi(std::strchr(buf, 'e') + 1) + 1;
    while (!d.dig.empty() && d.dig.back() == '0') d.dig.pop_back();
    return d;
}

// Number.RoundNumber(ref number, pos, isCorrectlyRounded:false): half-up on the digit string.
inline void round_number(DigitBuf& d, int pos) {
    int nd = (int)d.dig.size();
    int i = 0;
    while (i < pos && i < nd) i++;
    if (i == pos && i < nd && d.dig[i] >= '5') {
        while (i > 0 && d.dig[i - 1] == '9') i--;
        if (i > 0) {
            d.dig[i - 1]++;
        } else {
            d.scale++;
            d.dig.assign(1, '1');
            i = 1;
        }
        d.dig.resize(i);
    } else {
        if (i > nd) i = nd;
        while (i > 0 && d.dig[i - 1] == '0') i--;
        d.dig.resize(i);
    }
    if (d.dig.empty()) d.scale = 0;  // FloatingPoint kind: the sign is kept
}

// d.ToString("0.###") and the general "0.###...#" family (decimals = number of '#').
inline std::string fmt_custom(double v, int decimals = 3) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "\xE2\x88\x9E" : "-\xE2\x88\x9E";
    DigitBuf d = digits15(v);
    round_number(d, d.scale + decimals);
    std::string out;
    if (d.neg) out.push_back('-');
    int nd = (int)d.dig.size();
    if (d.scale <= 0) {
        out.push_back('0');
    } else {
        for (int k = 0; k < d.scale; k++) out.push_back(k < nd ? d.dig[k] : '0');
    }
    if (nd > d.scale && nd > 0) {
        out.push_back('.');
        for (int idx = d.scale; idx < nd; idx++) out.push_back(idx < 0 ? '0' : d.dig[idx]);
    }
    return out;
}

// Math.Round(value, digits): MidpointRounding.ToEven on value*10^digits.
inline double math_round(double v, int digits) {
    static const double p10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8,
                                 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    if (std::fabs(v) < 1e16) {
        double p = p10[digits];
        v *= p;
        v = std::nearbyint(v);
        v /= p;
    }
    return v;
}
inline double math_round0(double v) { return std::nearbyint(v); }

// double.ToString(): shortest round-trippable digits, "G" layout (scientific when the decimal
// exponent is > 15 or < -4 digits).
inline std::string fmt_roundtrip(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "\xE2\x88\x9E" : "-\xE2\x88\x9E";
    if (v == 0.0) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    *r.ptr = 0;
    // buf = d[.ddd]e[+-]XX
    std::string dig;
    const char* e = std::strchr(buf, 'e');
    for (const char* p = buf; p < e; ++p)
        if (*p != '.') dig.push_back(*p);
    int scale = std::atoi(e + 1) + 1;
    while (dig.size() > 1 && dig.back() == '0') dig.pop_back();
    std::string out;
    if (std::signbit(v)) out.push_back('-');
    int nd = (int)dig.size();
    if (scale > 15 || scale < -3) {
        out.push_back(dig[0]);
        if (nd > 1) {
            out.push_back('.');
            out.append(dig, 1, std::string::npos);
        }
        int ex = scale - 1;
        char eb[16];
        std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
        out += eb;
        return out;
    }
    if (scale <= 0) {
        out += "0.";
        for (int k = scale; k < 0; k++) out.push_back('0');
        out += dig;
    } else {
        for (int k = 0; k < scale; k++) out.push_back(k < nd ? dig[k] : '0');
        if (nd > scale) {
            out.push_back('.');
            out.append(dig, scale, std::string::npos);
        }
    }
    return out;
}

// d.ToString("F<k>", InvariantCulture): correctly rounded from the exact binary value.
inline std::string fmt_fixed(double v, int k) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "Infinity" : "-Infinity";
    char buf[400];
    std::snprintf(buf, sizeof buf, "%.*f", k, v);
    return buf;
}

inline std::string pad_left(const std::string& s, int width) {
    // .NET PadLeft counts UTF-16 code units; "∞" is one unit but three UTF-8 bytes.
    int units = 0;
    for (unsigned char ch : s)
        if ((ch & 0xC0) != 0x80) units++;
    if (units >= width) return s;
    return std::string(width - units, ' ') + s;
}

}  // namespace orc
