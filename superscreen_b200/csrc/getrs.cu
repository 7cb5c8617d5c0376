// Multi-RHS triangular solves with the no-pivot LU (K12): scipy.linalg.lu_solve at
// solver/solve_film.py:367,388,530,545.
//
// Right-looking blocked substitution over the 128-row blocks of the factorization.  One kernel per
// block step k:  every CTA applies the rank-128 update  B[i] -= F[i, block k] x_k  to its share of
// the rows still to be solved (one warp per row, 1 KB coalesced reads of the factor row), and
// CTA 0 -- which owns the 128 rows of the NEXT block -- finishes that block with the stored inverse
// of its diagonal block (x_next = inv(F_next,next) b_next), so the next kernel starts from a
// finished x.  The chain per step is therefore launch + one tall GEMV + one 128x128 GEMV by a
// single CTA; nrhs = 1 is HBM-bound in the limit (8 n^2 bytes for both sweeps).
#include "scb_common.cuh"

namespace scb {

constexpr int NB = SCB_LU_BLOCK;

// one warp: acc[c] = sum_q a[q] * xs[lane + 32 q][c], reduced over the warp
template <int RT>
__device__ __forceinline__ void row_dot(const double (&a)[4], const double (*xs)[RT + 1], int lane,
                                        double (&acc)[RT]) {
#pragma unroll
  for (int c = 0; c < RT; c++) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) s += a[q] * xs[lane + 32 * q][c];
    acc[c] = s;
  }
#pragma unroll
  for (int c = 0; c < RT; c++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
  }
}

// step kernel.  x_k (final) is stored in rows [k*128, k*128+128) of B.
//   do_update : apply B[i] -= F[i, block k] x_k for rows in [row_lo, row_hi) (minus the next block)
//   next      : index of the block to finish (x_next = dnext * b_next), or -1
template <int RT>  // rhs columns per pass (1 for the single right-hand side of solve_film, else 8)
__global__ void __launch_bounds__(256)
getrs_step_kernel(const double* __restrict__ F, int64_t ld, const double* __restrict__ dnext, int64_t k,
                  int64_t next, int do_update, int64_t nrhs, double* __restrict__ B, int64_t row_lo,
                  int64_t row_hi) {
  __shared__ double xs[NB][RT + 1];
  __shared__ double bs[NB][RT + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t o = k * NB;
  const int64_t on = next * NB;
  for (int64_t r0 = 0; r0 < nrhs; r0 += RT) {
    const int nr = (int)((nrhs - r0) < RT ? (nrhs - r0) : RT);
    __syncthreads();
    if (do_update) {
      for (int idx = tid; idx < NB * RT; idx += 256) {
        const int r = idx / RT, c = idx % RT;
        xs[r][c] = c < nr ? B[(o + r) * nrhs + r0 + c] : 0.0;
      }
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      if (next >= 0) {
        // The 128 rows of the next block; each warp owns 16 rows.  The rows of the stored inverse
        // are prefetched into L2 first, and the 64 loads of each phase are issued up front, so a
        // step costs two short memory round trips instead of 32 dependent ones.
        {
          const char* base = reinterpret_cast<const char*>(dnext);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int line = (warp * 4 + q) * 32 + lane;  // 1024 lines of 128 B = 128 KB
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)line * 128));
          }
        }
        double bval[16];
#pragma unroll
        for (int rr = 0; rr < 16; rr++) {
          const int r = warp + 8 * rr;
          bval[rr] = lane < nr ? B[(on + r) * nrhs + r0 + lane] : 0.0;
        }
        if (do_update) {
          double fa[16][4];
#pragma unroll
          for (int rr = 0; rr < 16; rr++) {
            const double* Frow = F + (on + warp + 8 * rr) * ld + o;
#pragma unroll
            for (int q = 0; q < 4; q++) fa[rr][q] = Frow[lane + 32 * q];
          }
#pragma unroll
          for (int rr = 0; rr < 16; rr++) {
            double accv[RT];
            row_dot<RT>(fa[rr], xs, lane, accv);
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < RT; c++) v = (lane == c) ? accv[c] : v;
            bval[rr] -= v;
          }
        }
#pragma unroll
        for (int rr = 0; rr < 16; rr++)
          if (lane < RT) bs[warp + 8 * rr][lane] = bval[rr];
        __syncthreads();
        {
          double da[16][4];
#pragma unroll
          for (int rr = 0; rr < 16; rr++) {
            const double* Drow = dnext + (int64_t)(warp + 8 * rr) * NB;
#pragma unroll
            for (int q = 0; q < 4; q++) da[rr][q] = Drow[lane + 32 * q];
          }
#pragma unroll
          for (int rr = 0; rr < 16; rr++) {
            double accv[RT];
            row_dot<RT>(da[rr], bs, lane, accv);
            if (lane < nr) {
              double v = 0.0;
#pragma unroll
              for (int c = 0; c < RT; c++) v = (lane == c) ? accv[c] : v;
              B[(on + warp + 8 * rr) * nrhs + r0 + lane] = v;
            }
          }
        }
      }
    } else if (do_update) {
      // rank-128 update of the remaining rows (next block excluded): 2 rows per warp, all loads
      // (factor rows and the right-hand-side entries) issued before the first use
      const int64_t i0 = row_lo + ((int64_t)(blockIdx.x - 1) * 8 + warp) * 2;
      double a[2][4], bv[2];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int64_t i = i0 + u;
        ok[u] = i < row_hi && !(next >= 0 && i >= on && i < on + NB);
        if (ok[u]) {
          const double* Frow = F + i * ld + o;
#pragma unroll
          for (int q = 0; q < 4; q++) a[u][q] = Frow[lane + 32 * q];
          bv[u] = lane < nr ? B[i * nrhs + r0 + lane] : 0.0;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (!ok[u]) continue;
        double accv[RT];
        row_dot<RT>(a[u], xs, lane, accv);
        if (lane < nr) {
          double v = 0.0;
#pragma unroll
          for (int c = 0; c < RT; c++) v = (lane == c) ? accv[c] : v;
          B[(i0 + u) * nrhs + r0 + lane] = bv[u] - v;
        }
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
  auto grid_for = [](int64_t rows) { return (int)((rows + 15) / 16) + 1; };  // CTA 0 + 16 rows per CTA
  auto launch = [&](int grid, const double* dnext, int64_t k, int64_t next, int upd, int64_t lo, int64_t hi) {
    if (nrhs == 1)
      getrs_step_kernel<1><<<grid, 256, 0, s>>>(LU, n_pad, dnext, k, next, upd, nrhs, B, lo, hi);
    else
      getrs_step_kernel<8><<<grid, 256, 0, s>>>(LU, n_pad, dnext, k, next, upd, nrhs, B, lo, hi);
  };
  const double* invL = dinv;            // block k: dinv + k*2*128*128
  const double* invU = dinv + NB * NB;  // block k: dinv + k*2*128*128 + 128*128
  const int64_t bs = 2 * NB * NB;
  // forward: L y = b.  prologue finishes block 0, step k updates rows below and finishes block k+1
  launch(1, invL, 0, 0, 0, 0, 0);
  SCB_LAUNCH_CHECK();
  for (int64_t k = 0; k + 1 < nb; k++) {
    const int64_t lo = (k + 1) * NB, hi = n_pad;
    launch(grid_for(hi - lo), invL + (k + 1) * bs, k, k + 1, 1, lo, hi);
    SCB_LAUNCH_CHECK();
  }
  // backward: U x = y
  launch(1, invU + (nb - 1) * bs, 0, nb - 1, 0, 0, 0);
  SCB_LAUNCH_CHECK();
  for (int64_t k = nb - 1; k >= 1; k--) {
    const int64_t lo = 0, hi = k * NB;
    launch(grid_for(hi - lo), invU + (k - 1) * bs, k, k - 1, 1, lo, hi);
    SCB_LAUNCH_CHECK();
  }
  return SCB_OK;
}
