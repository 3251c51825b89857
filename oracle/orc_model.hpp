// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_dotnet.hpp header).  PARITY UNPINNED by the
// reference (it ships no tests or fixtures); pinned by the builder's known-answer tests in
// tests/golden/.
//
// Model types and text parser, following
//   R/Models/PrimalSimplex.cs:8-49   (Sense, Rel, Constraint, LPProblem, SimplexResult)
//   R/Models/LPParser.cs:9-79        (ParseFromText, ParseCoefficients)
// where R = /root/reference/Linear_Programming_Solver.
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

enum Sense { MAX = 0, MIN = 1 };
enum Rel { LE = 0, GE = 1, EQ = 2 };

struct Row {
    std::vector<double> a;
    int rel = LE;
    double b = 0;
};

struct Problem {
    int sense = MAX;
    std::vector<double> c;
    std::vector<Row> rows;
    int nvars() const { return (int)c.size(); }
};

// bool[,] highlight mask handed to the updatePivot callback (null -> rows == 0).
struct Mask {
    int rows = 0, cols = 0;
    std::vector<uint8_t> bits;
};
using Sink = std::function<void(const std::string&, const Mask&)>;

struct Outcome {  // SimplexResult
    std::string report, summary;
    double z = 0;
    bool has_x = false, has_tableau = false;
    std::vector<double> x;
    std::vector<double> T;  // row-major rows x cols
    int rows = 0, cols = 0;
    std::vector<int> basis;
    std::vector<std::string> names;
};

// What the arithmetic did, for bit-exact comparison with the CUDA path.
struct Trace {
    std::vector<int> enter, leave;        // one entry per Pivot call
    bool keep_history = false;
    std::vector<std::vector<double>> history;  // tableau after iteration k (k = 0 .. pivots)
    int status = 0;                        // 0 OPTIMAL, 1 UNBOUNDED, 2 INFEASIBLE (dual only)
    int silent_pivots = 0;                 // DualSimplex.ForceDualFeasibility pivots
};

struct SolveError : std::runtime_error {
    int code;
    SolveError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
// negative status codes shared with include/lpx.h
enum { ERR_GE_ROW = -1, ERR_NEG_RHS = -2, ERR_ITER_LIMIT = -3, ERR_BAD_ARGS = -4, ERR_PARSE = -6,
       ERR_UNSUPPORTED_ALGO = -7, ERR_REV_UNSUPPORTED = -10, ERR_SINGULAR = -11 };

extern std::string g_newline;  // Environment.NewLine ("\n" here; "\r\n" on the reference's Windows)

Problem parse_text(const std::string& input);

}  // namespace orc
