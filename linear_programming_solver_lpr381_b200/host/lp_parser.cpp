// LPParser.ParseFromText / ParseCoefficients (R/Models/LPParser.cs:9-79) for the host layer.
// Format: "Max:"/"Min:" objective line, then one constraint per line; coefficients are
// positional (the digits after 'x' are matched but ignored), no exponent notation in
// coefficients, RHS parsed like double.TryParse.
#include <cctype>
#include <cmath>
#include <cstdlib>

#include "lp_model.hpp"

namespace lpr381 {

std::string& NewLine() {
    static std::string nl = "\n";
    return nl;
}

namespace {

bool space(char ch) { return ch == ' ' || (ch >= '\t' && ch <= '\r'); }

std::string Trim(const std::string& s) {
    size_t lo = 0, hi = s.size();
    while (lo < hi && space(s[lo])) ++lo;
    while (hi > lo && space(s[hi - 1])) --hi;
    return s.substr(lo, hi - lo);
}

// double.TryParse with NumberStyles.Float | AllowThousands
bool TryParseDouble(const std::string& text, double& value) {
    const std::string s = Trim(text);
    if (s.empty()) return false;
    if (s == "NaN") return value = std::nan(""), true;
    if (s == "Infinity" || s == "+Infinity" || s == "\xE2\x88\x9E") return value = INFINITY, true;
    if (s == "-Infinity" || s == "-\xE2\x88\x9E") return value = -INFINITY, true;
    std::string num;
    size_t k = 0;
    if (s[k] == '-' || s[k] == '+') num += s[k++];
    size_t mantissa_digits = 0;
    for (; k < s.size() && (std::isdigit((unsigned char)s[k]) || s[k] == ','); ++k)
        if (s[k] != ',') num += s[k], ++mantissa_digits;
    if (k < s.size() && s[k] == '.') {
        num += s[k++];
        for (; k < s.size() && std::isdigit((unsigned char)s[k]); ++k) num += s[k], ++mantissa_digits;
    }
    if (mantissa_digits == 0) return false;
    if (k < s.size() && (s[k] == 'e' || s[k] == 'E')) {
        std::string ex(1, 'e');
        size_t j = k + 1;
        if (j < s.size() && (s[j] == '-' || s[j] == '+')) ex += s[j++];
        size_t exp_digits = 0;
        for (; j < s.size() && std::isdigit((unsigned char)s[j]); ++j) ex += s[j], ++exp_digits;
        if (exp_digits == 0) return false;
        num += ex;
        k = j;
    }
    if (k != s.size()) return false;
    value = std::strtod(num.c_str(), nullptr);
    return true;
}

// one "+"-separated term must match ^([-]?\d*\.?\d*)x\d+$
bool SplitTerm(const std::string& term, std::string& coefficient) {
    size_t k = 0;
    if (k < term.size() && term[k] == '-') ++k;
    while (k < term.size() && std::isdigit((unsigned char)term[k])) ++k;
    if (k < term.size() && term[k] == '.') ++k;
    while (k < term.size() && std::isdigit((unsigned char)term[k])) ++k;
    const size_t coef_end = k;
    if (k >= term.size() || term[k] != 'x') return false;
    const size_t first_index_digit = ++k;
    while (k < term.size() && std::isdigit((unsigned char)term[k])) ++k;
    if (k == first_index_digit || k != term.size()) return false;
    coefficient = term.substr(0, coef_end);
    return true;
}

std::vector<double> ParseCoefficients(const std::string& expression) {
    std::string expr;
    for (char ch : expression) {
        if (ch == '-') expr += "+-";
        else if (ch != ' ') expr += ch;
    }
    std::vector<double> coefficients;
    size_t begin = 0;
    while (begin <= expr.size()) {
        size_t end = expr.find('+', begin);
        if (end == std::string::npos) end = expr.size();
        const std::string part = expr.substr(begin, end - begin);
        begin = end + 1;
        if (part.empty()) continue;
        std::string valStr;
        if (!SplitTerm(Trim(part), valStr)) throw LpException("Cannot parse coefficient: " + part);
        double val;
        if (valStr.empty()) val = 1;
        else if (valStr == "-") val = -1;
        else if (!TryParseDouble(valStr, val))
            throw LpException("The input string '" + valStr + "' was not in a correct format.");
        coefficients.push_back(val);
    }
    return coefficients;
}

}  // namespace

LPProblem LPParser::ParseFromText(const std::string& input) {
    std::vector<std::string> lines;
    for (size_t begin = 0; begin <= input.size();) {
        size_t end = input.find_first_of("\r\n", begin);
        if (end == std::string::npos) end = input.size();
        const std::string line = Trim(input.substr(begin, end - begin));
        if (!line.empty()) lines.push_back(line);
        begin = end + 1;
    }
    if (lines.size() < 2) throw LpException("Input must contain an objective and at least one constraint.");

    // ^(max|min)\s*:\s*(.+)$  (IgnoreCase)
    const std::string& head = lines[0];
    std::string word;
    for (size_t k = 0; k < 3 && k < head.size(); ++k) word += (char)std::tolower((unsigned char)head[k]);
    size_t k = 3;
    bool ok = word == "max" || word == "min";
    while (ok && k < head.size() && space(head[k])) ++k;
    ok = ok && k < head.size() && head[k] == ':';
    ++k;
    while (ok && k < head.size() && space(head[k])) ++k;
    ok = ok && k < head.size();
    if (!ok) throw LpException("Objective format incorrect. Example: Max: 3x1 + 5x2");

    LPProblem problem;
    problem.ObjectiveSense = word == "max" ? Sense::Max : Sense::Min;
    problem.C = ParseCoefficients(head.substr(k));

    for (size_t i = 1; i < lines.size(); ++i) {
        const std::string& consLine = lines[i];
        // ^(.+?)(<=|>=|=)(.+)$: the earliest relation sign that leaves a non-empty right side
        size_t at = std::string::npos, width = 0;
        for (size_t pos = 1; pos < consLine.size() && at == std::string::npos; ++pos) {
            size_t w = 0;
            if (consLine[pos] == '=') w = 1;
            else if ((consLine[pos] == '<' || consLine[pos] == '>') && pos + 1 < consLine.size() && consLine[pos + 1] == '=')
                w = 2;
            if (w && pos + w < consLine.size()) at = pos, width = w;
        }
        if (at == std::string::npos) throw LpException("Constraint format incorrect: " + consLine);
        const std::string relStr = consLine.substr(at, width);
        const std::string rhsStr = Trim(consLine.substr(at + width));
        Constraint cons;
        cons.A = ParseCoefficients(Trim(consLine.substr(0, at)));
        cons.Relation = relStr == "<=" ? Rel::LE : relStr == ">=" ? Rel::GE : Rel::EQ;
        if (!TryParseDouble(rhsStr, cons.B)) throw LpException("Invalid RHS number: " + rhsStr);
        problem.Constraints.push_back(cons);
    }
    return problem;
}

}  // namespace lpr381
