// lpx_knap.hpp — internal declarations of the knapsack Branch & Bound engines (not part of the C ABI).
#pragma once
#include "../../include/lpx.h"

namespace lpx {

// Device-resident search (lpx_knap_dev.cu): one warp per instance runs the whole best-first loop.
int knapsack_search_device(int count, int n, const double* profit, const double* weight, const double* capacity,
                           const lpx_options& opt, int* found, double* best_value, int* best_x, long long* n_evals,
                           long long* n_pops, int* rank_order, lpx_knap_pop_fn on_pop, void* user);
void knapsack_dev_release_cache();

}  // namespace lpx
