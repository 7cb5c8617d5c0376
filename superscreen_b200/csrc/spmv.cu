// CSR sparse mat-vec (K14): grad_y @ g, grad_x @ g, Gy @ g (solver/solve_film.py:556-559).
// ~7 nnz per row; HBM-bound (12*nnz + 16*n bytes).  One thread per (row, rhs); rows are summed in
// stored (sorted-column) order -> deterministic.
#include "scb_common.cuh"

namespace scb {
__global__ void spmv_kernel(int64_t nrows, const int32_t* __restrict__ indptr,
                            const int32_t* __restrict__ indices, const double* __restrict__ data,
                            int64_t nrhs, const double* __restrict__ x, double alpha, double beta,
                            double* __restrict__ y) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= nrows * nrhs) return;
  const int64_t i = idx / nrhs, k = idx % nrhs;
  double acc = 0.0;
  for (int32_t q = indptr[i]; q < indptr[i + 1]; q++) acc += data[q] * x[(int64_t)indices[q] * nrhs + k];
  y[idx] = beta == 0.0 ? alpha * acc : alpha * acc + beta * y[idx];
}
}  // namespace scb

using namespace scb;

extern "C" int scb_spmv(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                        const double* data, int64_t nrhs, const double* x, double alpha, double beta,
                        double* y, scb_stream_t stream) {
  SCB_CHECK_ARG(nrows >= 0 && nrhs > 0, "bad sizes");
  if (nrows == 0) return SCB_OK;
  spmv_kernel<<<(unsigned)ceil_div(nrows * nrhs, 256), 256, 0, (cudaStream_t)stream>>>(
      nrows, indptr, indices, data, nrhs, x, alpha, beta, y);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}
