// lpx_stream.hpp — internal cross-file declarations (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include "../../include/lpx.h"

namespace lpx {

// whole-GPU streaming solver (lpx_stream.cu), host buffers in / host buffers out
int stream_solve_host(int m, int n, int sense, const double* A, const int* rel, const double* b, const double* c,
                      const lpx_options* opt, int* status, int* n_pivots, int* pivots, int pivots_cap, int* basis,
                      double* x, double* z, double* tableau, double* history, int history_cap);

// register-resident batched kernel (lpx_reg.cu)
bool reg_kernel_supports(int m, int n, int m_expanded, bool has_rel);
int reg_launch_batched(int count, int m, int n, int sense, const double* A, const double* b, const double* c,
                       const lpx_options& opt, int* status, int* n_pivots, int* basis, double* x, double* z,
                       double* tableau, unsigned long long* total_pivots, cudaStream_t stream);

void set_bnb_instance(int k);

}  // namespace lpx
