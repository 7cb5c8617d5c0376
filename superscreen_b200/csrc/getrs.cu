// Multi-RHS triangular solves with the no-pivot LU (K12): scipy.linalg.lu_solve at
// solver/solve_film.py:367,388,530,545.  Right-looking blocked substitution: per 128-row block
// one kernel solves the diagonal block with its stored inverse (redundantly in every CTA, from
// L2) and applies the rank-128 update to the rows still to be solved.  nrhs = 1 is HBM-bound
// (one sweep over L and one over U = 8 n^2 bytes).
#include "scb_common.cuh"

namespace scb {

constexpr int NB = SCB_LU_BLOCK;
constexpr int RT = 8;  // rhs columns per pass

// forward (lower = 1):  y_k = invL_kk b_k ;  b[i] -= L[i, kblock] y_k   for rows i below block k
// backward (lower = 0): x_k = invU_kk b_k ;  b[i] -= U[i, kblock] x_k   for rows i above block k
// grid.x CTAs split the remaining rows; every CTA recomputes the 128 x nrhs diagonal solve into
// shared memory; CTA 0 writes it back.
__global__ void __launch_bounds__(256)
getrs_step_kernel(const double* __restrict__ LU, int64_t ld, const double* __restrict__ dblk, int64_t k,
                  int lower, int64_t nrhs, double* __restrict__ B, int64_t row_lo, int64_t row_hi) {
  __shared__ double xs[NB][RT + 1];
  __shared__ double bs[NB][RT + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t o = k * NB;
  for (int64_t r0 = 0; r0 < nrhs; r0 += RT) {
    const int nr = (int)((nrhs - r0) < RT ? (nrhs - r0) : RT);
    __syncthreads();
    for (int idx = tid; idx < NB * nr; idx += 256) {
      const int r = idx / nr, c = idx % nr;
      bs[r][c] = B[(o + r) * nrhs + r0 + c];
    }
    __syncthreads();
    // x = dblk (128x128, triangular incl. zeros) @ bs : thread -> (row r = tid/2, half h = tid%2)
    {
      const int r = tid >> 1, h = tid & 1;
      double accv[RT];
#pragma unroll
      for (int c = 0; c < RT; c++) accv[c] = 0.0;
      const int klo = lower ? 0 : r, khi = lower ? r + 1 : NB;
      for (int kk = klo + h; kk < khi; kk += 2) {
        const double a = dblk[r * NB + kk];
#pragma unroll
        for (int c = 0; c < RT; c++) accv[c] += a * bs[kk][c];
      }
#pragma unroll
      for (int c = 0; c < RT; c++) {
        accv[c] += __shfl_xor_sync(0xffffffffu, accv[c], 1);
        if (h == 0) xs[r][c] = accv[c];
      }
    }
    __syncthreads();
    if (blockIdx.x == 0)
      for (int idx = tid; idx < NB * nr; idx += 256) {
        const int r = idx / nr, c = idx % nr;
        B[(o + r) * nrhs + r0 + c] = xs[r][c];
      }
    // rank-128 update of this CTA's share of the remaining rows: one warp per row
    const int64_t nrows = row_hi - row_lo;
    const int64_t per = (nrows + gridDim.x - 1) / gridDim.x;
    const int64_t lo = row_lo + blockIdx.x * per;
    const int64_t hi = lo + per < row_hi ? lo + per : row_hi;
    for (int64_t i = lo + warp; i < hi; i += 8) {
      const double* Lrow = LU + i * ld + o;
      double a[4];
#pragma unroll
      for (int q = 0; q < 4; q++) a[q] = Lrow[lane + 32 * q];
      double accv[RT];
#pragma unroll
      for (int c = 0; c < RT; c++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) s += a[q] * xs[lane + 32 * q][c];
        accv[c] = s;
      }
#pragma unroll
      for (int c = 0; c < RT; c++) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) accv[c] += __shfl_xor_sync(0xffffffffu, accv[c], off);
      }
      if (lane < nr) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < RT; c++) v = (lane == c) ? accv[c] : v;
        B[i * nrhs + r0 + lane] -= v;
      }
    }
  }
}

}  // namespace scb

using namespace scb;

extern "C" int scb_getrs_nopiv(int64_t n_pad, const double* LU, const double* dinv, int64_t nrhs,
                               double* B, scb_stream_t stream) {
  SCB_CHECK_ARG(n_pad > 0 && n_pad % NB == 0, "n_pad must be a positive multiple of 128");
  SCB_CHECK_ARG(nrhs > 0, "nrhs must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nb = n_pad / NB;
  for (int64_t k = 0; k < nb; k++) {
    const int64_t lo = (k + 1) * NB, hi = n_pad;
    int grid = (int)((hi - lo + 63) / 64);
    grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
    getrs_step_kernel<<<grid, 256, 0, s>>>(LU, n_pad, dinv + k * 2 * NB * NB, k, 1, nrhs, B, lo, hi);
    SCB_LAUNCH_CHECK();
  }
  for (int64_t k = nb - 1; k >= 0; k--) {
    const int64_t lo = 0, hi = k * NB;
    int grid = (int)((hi - lo + 63) / 64);
    grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
    getrs_step_kernel<<<grid, 256, 0, s>>>(LU, n_pad, dinv + k * 2 * NB * NB + NB * NB, k, 0, nrhs, B, lo, hi);
    SCB_LAUNCH_CHECK();
  }
  return SCB_OK;
}
