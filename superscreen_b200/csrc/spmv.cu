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

// ---- fused pieces of one film solve (solver/solve_film.py:526-531,556): each replaces a chain of
// ---- gather / add / scale / scatter steps by one pass

// B[r, c] = (applied[ix[r], c] + other[ix[r], c] - ha_eff[ix[r], c]) * scale[r];  padding rows = 0
__global__ void solve_rhs_kernel(int64_t n_int, int64_t n_pad, const int64_t* __restrict__ ix, int64_t nrhs,
                                 const double* __restrict__ applied, const double* __restrict__ other,
                                 const double* __restrict__ ha_eff, const double* __restrict__ scale,
                                 double* __restrict__ B) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n_pad * nrhs) return;
  const int64_t r = idx / nrhs, c = idx % nrhs;
  double v = 0.0;
  if (r < n_int) {
    const int64_t i = ix[r] * nrhs + c;
    v = applied[i];
    if (other) v += other[i];
    if (ha_eff) v -= ha_eff[i];
    if (scale) v *= scale[r];
  }
  B[idx] = v;
}

// g[i, c] = g0[i, c] + (pos[i] >= 0 ? X[pos[i], c] / scale[pos[i]] : 0)
__global__ void solve_stream_kernel(int64_t n, int64_t nrhs, const int32_t* __restrict__ pos,
                                    const double* __restrict__ X, const double* __restrict__ scale,
                                    const double* __restrict__ g0, double* __restrict__ g) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * nrhs) return;
  const int64_t i = idx / nrhs, c = idx % nrhs;
  const int32_t r = pos[i];
  double v = g0 ? g0[idx] : 0.0;
  if (r >= 0) {
    const double x = X[(int64_t)r * nrhs + c];
    v += scale ? x / scale[r] : x;
  }
  g[idx] = v;
}

// J[i, c, 0] = (grad_y g)[i, c],  J[i, c, 1] = -(grad_x g)[i, c]; both operators share the pattern
__global__ void current_density_kernel(int64_t n, const int32_t* __restrict__ indptr,
                                       const int32_t* __restrict__ indices, const double* __restrict__ gx,
                                       const double* __restrict__ gy, int64_t nrhs, const double* __restrict__ g,
                                       double* __restrict__ J) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * nrhs) return;
  const int64_t i = idx / nrhs, k = idx % nrhs;
  double ax = 0.0, ay = 0.0;
  for (int32_t q = indptr[i]; q < indptr[i + 1]; q++) {
    const double v = g[(int64_t)indices[q] * nrhs + k];
    ax += gx[q] * v;
    ay += gy[q] * v;
  }
  J[2 * idx] = 1.0 * ay;
  J[2 * idx + 1] = -1.0 * ax;
}
}  // namespace scb

using namespace scb;

extern "C" int scb_solve_rhs(int64_t n_int, int64_t n_pad, const int64_t* ix, int64_t nrhs, const double* applied,
                             const double* other, const double* ha_eff, const double* scale, double* B,
                             scb_stream_t stream) {
  SCB_CHECK_ARG(n_int >= 0 && n_pad >= n_int && nrhs > 0, "bad sizes");
  if (n_pad == 0) return SCB_OK;
  solve_rhs_kernel<<<(unsigned)ceil_div(n_pad * nrhs, 256), 256, 0, (cudaStream_t)stream>>>(
      n_int, n_pad, ix, nrhs, applied, other, ha_eff, scale, B);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int scb_solve_stream(int64_t n, int64_t nrhs, const int32_t* pos, const double* X, const double* scale,
                                const double* g0, double* g, scb_stream_t stream) {
  SCB_CHECK_ARG(n >= 0 && nrhs > 0, "bad sizes");
  if (n == 0) return SCB_OK;
  solve_stream_kernel<<<(unsigned)ceil_div(n * nrhs, 256), 256, 0, (cudaStream_t)stream>>>(n, nrhs, pos, X, scale,
                                                                                            g0, g);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int scb_current_density(int64_t n, const int32_t* indptr, const int32_t* indices, const double* gradient_x,
                                   const double* gradient_y, int64_t nrhs, const double* g, double* J,
                                   scb_stream_t stream) {
  SCB_CHECK_ARG(n >= 0 && nrhs > 0, "bad sizes");
  if (n == 0) return SCB_OK;
  current_density_kernel<<<(unsigned)ceil_div(n * nrhs, 256), 256, 0, (cudaStream_t)stream>>>(
      n, indptr, indices, gradient_x, gradient_y, nrhs, g, J);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

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

extern "C" int scb_solve_step(int64_t n, int64_t n_int, int64_t n_pad, int64_t nrhs, const int64_t* rhs_ix,
                              const double* applied, const double* other, const double* ha_eff, const double* scale,
                              const double* LU, const double* dinv, double* B, const int32_t* pos, const double* g0,
                              double* g, const int32_t* op_indptr, const int32_t* op_indices,
                              const double* gradient_x, const double* gradient_y, double* J, scb_stream_t stream) {
  if (int rc = scb_solve_rhs(n_int, n_pad, rhs_ix, nrhs, applied, other, ha_eff, scale, B, stream)) return rc;
  if (int rc = scb_getrs_nopiv(n_pad, LU, dinv, nrhs, B, stream)) return rc;
  if (int rc = scb_solve_stream(n, nrhs, pos, B, scale, g0, g, stream)) return rc;
  return scb_current_density(n, op_indptr, op_indices, gradient_x, gradient_y, nrhs, g, J, stream);
}
