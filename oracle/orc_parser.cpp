// ORACLE — TEST INFRASTRUCTURE ONLY.  Text model parser, following R/Models/LPParser.cs:9-79.
#include "orc_model.hpp"

#include <cctype>
#include <cstdlib>
#include <cmath>

namespace orc {

std::string g_newline = "\n";

static bool is_ws(char ch) { return ch == ' ' || ch == '\t' || ch == '\v' || ch == '\f' || ch == '\r' || ch == '\n'; }

static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && is_ws(s[a])) a++;
    while (b > a && is_ws(s[b - 1])) b--;
    return s.substr(a, b - a);
}

// double.TryParse(s, out v): NumberStyles.Float | AllowThousands, '.' decimal separator.
static bool dotnet_try_parse(const std::string& raw, double& out) {
    std::string s = trim(raw);
    if (s.empty()) return false;
    if (s == "NaN") { out = std::strtod("nan", nullptr); return true; }
    if (s == "Infinity" || s == "+Infinity" || s == "\xE2\x88\x9E") { out = HUGE_VAL; return true; }
    if (s == "-Infinity" || s == "-\xE2\x88\x9E") { out = -HUGE_VAL; return true; }
    std::string clean;
    size_t i = 0;
    if (s[i] == '+' || s[i] == '-') clean.push_back(s[i++]);
    bool digits = false;
    while (i < s.size() && (std::isdigit((unsigned char)s[i]) || s[i] == ',')) {
        if (s[i] != ',') { clean.push_back(s[i]); digits = true; }
        i++;
    }
    if (i < s.size() && s[i] == '.') {
        clean.push_back('.');
        i++;
        while (i < s.size() && std::isdigit((unsigned char)s[i])) { clean.push_back(s[i++]); digits = true; }
    }
    if (!digits) return false;
    if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
        size_t j = i + 1;
        std::string ex = "e";
        if (j < s.size() && (s[j] == '+' || s[j] == '-')) ex.push_back(s[j++]);
        bool ed = false;
        while (j < s.size() && std::isdigit((unsigned char)s[j])) { ex.push_back(s[j++]); ed = true; }
        if (!ed) return false;
        clean += ex;
        i = j;
    }
    if (i != s.size()) return false;
    out = std::strtod(clean.c_str(), nullptr);
    return true;
}

// LPParser.ParseCoefficients (LPParser.cs:61-79): positional coefficients; the subscript after
// 'x' is matched but never used.
static std::vector<double> parse_coefficients(const std::string& expr_in) {
    std::string expr;
    for (char ch : expr_in) {           // Replace("-", "+-").Replace(" ", "")
        if (ch == '-') expr += "+-";
        else if (ch != ' ') expr.push_back(ch);
    }
    std::vector<double> coef;
    size_t pos = 0;
    while (pos <= expr.size()) {
        size_t nxt = expr.find('+', pos);
        if (nxt == std::string::npos) nxt = expr.size();
        std::string part = expr.substr(pos, nxt - pos);
        pos = nxt + 1;
        if (part.empty()) continue;     // StringSplitOptions.RemoveEmptyEntries
        std::string t = trim(part);
        // ^([-]?\d*\.?\d*)x\d+$
        size_t k = 0;
        if (k < t.size() && t[k] == '-') k++;
        while (k < t.size() && std::isdigit((unsigned char)t[k])) k++;
        if (k < t.size() && t[k] == '.') k++;
        while (k < t.size() && std::isdigit((unsigned char)t[k])) k++;
        size_t vend = k;
        bool ok = k < t.size() && t[k] == 'x';
        if (ok) {
            k++;
            size_t d0 = k;
            while (k < t.size() && std::isdigit((unsigned char)t[k])) k++;
            ok = k > d0 && k == t.size();
        }
        if (!ok) throw SolveError(ERR_PARSE, "Cannot parse coefficient: " + part);
        std::string val = t.substr(0, vend);
        double v;
        if (val.empty()) v = 1;
        else if (val == "-") v = -1;
        else if (!dotnet_try_parse(val, v))
            throw SolveError(ERR_PARSE, "The input string '" + val + "' was not in a correct format.");
        coef.push_back(v);
    }
    return coef;
}

static bool starts_ci(const std::string& s, const char* w) {
    for (size_t i = 0; w[i]; i++)
        if (i >= s.size() || std::tolower((unsigned char)s[i]) != w[i]) return false;
    return true;
}

Problem parse_text(const std::string& input) {
    std::vector<std::string> lines;
    size_t pos = 0;
    while (pos <= input.size()) {
        size_t nxt = input.find_first_of("\r\n", pos);
        if (nxt == std::string::npos) nxt = input.size();
        std::string l = trim(input.substr(pos, nxt - pos));
        if (!l.empty()) lines.push_back(l);
        pos = nxt + 1;
    }
    if (lines.size() < 2) throw SolveError(ERR_PARSE, "Input must contain an objective and at least one constraint.");

    // ^(max|min)\s*:\s*(.+)$, case-insensitive
    const std::string& obj = lines[0];
    bool is_max = starts_ci(obj, "max"), is_min = starts_ci(obj, "min");
    size_t k = 3;
    bool ok = is_max || is_min;
    if (ok) {
        while (k < obj.size() && is_ws(obj[k])) k++;
        ok = k < obj.size() && obj[k] == ':';
        k++;
        while (ok && k < obj.size() && is_ws(obj[k])) k++;
        ok = ok && k < obj.size();
    }
    if (!ok) throw SolveError(ERR_PARSE, "Objective format incorrect. Example: Max: 3x1 + 5x2");

    Problem p;
    p.sense = is_max ? MAX : MIN;
    p.c = parse_coefficients(obj.substr(k));

    for (size_t i = 1; i < lines.size(); i++) {
        const std::string& cl = lines[i];
        // ^(.+?)(<=|>=|=)(.+)$ : shortest non-empty LHS followed by a relation and a non-empty RHS
        size_t lhs_len = 0, rel_len = 0;
        bool found = false;
        for (size_t L = 1; L < cl.size() && !found; L++) {
            size_t rl = 0;
            if (cl.compare(L, 2, "<=") == 0 || cl.compare(L, 2, ">=") == 0) rl = 2;
            else if (cl[L] == '=') rl = 1;
            if (rl && L + rl < cl.size()) { lhs_len = L; rel_len = rl; found = true; }
        }
        if (!found) throw SolveError(ERR_PARSE, "Constraint format incorrect: " + cl);
        std::string lhs = trim(cl.substr(0, lhs_len));
        std::string rel = cl.substr(lhs_len, rel_len);
        std::string rhs = trim(cl.substr(lhs_len + rel_len));
        Row r;
        r.a = parse_coefficients(lhs);
        r.rel = rel == "<=" ? LE : rel == ">=" ? GE : EQ;
        if (!dotnet_try_parse(rhs, r.b)) throw SolveError(ERR_PARSE, "Invalid RHS number: " + rhs);
        p.rows.push_back(std::move(r));
    }
    return p;
}

}  // namespace orc
