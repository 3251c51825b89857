// Internal helpers shared by the host-layer translation units.
#pragma once
#include <string>
#include <vector>

#include "lp_model.hpp"

namespace lpr381 {

struct Flat {  // LPProblem as the C ABI wants it
    int m = 0, n = 0, sense = 0;
    std::vector<double> A, b, c;
    std::vector<int> rel;
};
Flat flatten(const LPProblem& p);
void throw_on(int rc);
std::vector<std::string> var_names(int n, int m);
std::string tableau_text(const char* title, const double* T, int rows, int cols, const std::vector<int>& basis,
                         const std::vector<std::string>& names, int iter);
Highlight cross(int rows, int cols, int row, int col);

}  // namespace lpr381
