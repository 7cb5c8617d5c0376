// Multi-RHS triangular solves with the no-pivot LU (K12): scipy.linalg.lu_solve at
// solver/solve_film.py:367,388,530,545.
//
// One PERSISTENT kernel per sweep (forward L y = b, backward U x = y).  The factor is processed in
// 128-row blocks; CTAs grab blocks in sweep order from an atomic work counter (so a CTA only ever
// waits for blocks that are already owned by running or finished CTAs: deadlock-free without
// co-residency assumptions).  For its block i a CTA streams the block row of the factor,
//     acc_i = b_i - sum_{j before i} F_ij x_j ,
// waiting on a per-block ready flag before it touches x_j, then finishes the block with the stored
// inverse of the diagonal block, x_i = inv(F_ii) acc_i, publishes x_i (threadfence + flag) and grabs
// the next block.  The critical chain per block is flag -> one 128x128 GEMV whose factor rows are
// already in registers -> one 128x128 GEMV from L2-prefetched data -> flag; all the HBM traffic
// (8 n^2 bytes for both sweeps at nrhs <= 8) streams off the chain.
#include "scb_common.cuh"

namespace scb {

int64_t lu_flags_offset(int64_t n_pad);  // getrf.cu

constexpr int NB = SCB_LU_BLOCK;

__device__ __forceinline__ void dmma_rhs(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

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

__device__ __forceinline__ void prefetch_tile_l2(const double* tile, int64_t ld_bytes, int warp, int lane) {
  // 128 rows x 1 KB: warp w covers rows 16w..16w+15, lane -> (row, 128-byte line)
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int idx = q * 32 + lane;  // 0..127 -> 16 rows x 8 lines
    const int r = warp * 16 + (idx >> 3), line = idx & 7;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(tile) + r * ld_bytes + line * 128));
  }
}

// RT: right-hand sides per pass.  lower = 1: forward sweep with L (unit diagonal handled through the
// stored inverse), lower = 0: backward sweep with U.
template <int RT>
__global__ void __launch_bounds__(256)
trsv_sweep_kernel(const double* __restrict__ F, int64_t ld, const double* __restrict__ dinv, int64_t nb,
                  int lower, int64_t nrhs, int64_t r0, double* __restrict__ B, int* __restrict__ flags,
                  int* __restrict__ counter, int epoch) {
  __shared__ double xs[NB][RT + 1];
  __shared__ double bs[NB][RT + 1];
  __shared__ int s_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nr = (int)((nrhs - r0) < RT ? (nrhs - r0) : RT);
  const int64_t ldb = ld * (int64_t)sizeof(double);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_p = atomicAdd(counter, 1);
    __syncthreads();
    const int p = s_p;  // position in sweep order
    if (p >= nb) break;
    const int64_t i = lower ? p : nb - 1 - p;
    const double* dblk = dinv + i * 2 * NB * NB + (lower ? 0 : NB * NB);
    prefetch_tile_l2(dblk, NB * sizeof(double), warp, lane);
    // right-hand side entries of my 16 rows (lane c < nr holds column c)
    double bval[16];
#pragma unroll
    for (int rr = 0; rr < 16; rr++) bval[rr] = lane < nr ? B[(i * NB + warp + 8 * rr) * nrhs + r0 + lane] : 0.0;
    if (p > 0) {
      const int64_t j0 = lower ? 0 : nb - 1;
      prefetch_tile_l2(F + i * NB * ld + j0 * NB, ldb, warp, lane);
    }
    for (int q = 0; q < p; q++) {
      const int64_t j = lower ? q : nb - 1 - q;
      // factor rows first (independent of x_j), next tile into L2, then wait for x_j
      double fa[16][4];
#pragma unroll
      for (int rr = 0; rr < 16; rr++) {
        const double* Frow = F + (i * NB + warp + 8 * rr) * ld + j * NB;
#pragma unroll
        for (int k = 0; k < 4; k++) fa[rr][k] = Frow[lane + 32 * k];
      }
      if (q + 1 < p) {
        const int64_t jn = lower ? q + 1 : nb - 2 - q;
        prefetch_tile_l2(F + i * NB * ld + jn * NB, ldb, warp, lane);
      }
      if (tid == 0) {
        while (*reinterpret_cast<volatile int*>(flags + j) != epoch) {
        }
        __threadfence();
      }
      __syncthreads();
      for (int idx = tid; idx < NB * RT; idx += 256) {
        const int r = idx / RT, c = idx % RT;
        xs[r][c] = c < nr ? __ldcg(&B[(j * NB + r) * nrhs + r0 + c]) : 0.0;
      }
      __syncthreads();
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
    // finish: x_i = inv(F_ii) acc_i
#pragma unroll
    for (int rr = 0; rr < 16; rr++)
      if (lane < RT) bs[warp + 8 * rr][lane] = bval[rr];
    double da[16][4];
#pragma unroll
    for (int rr = 0; rr < 16; rr++) {
      const double* Drow = dblk + (int64_t)(warp + 8 * rr) * NB;
#pragma unroll
      for (int k = 0; k < 4; k++) da[rr][k] = Drow[lane + 32 * k];
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 16; rr++) {
      double accv[RT];
      row_dot<RT>(da[rr], bs, lane, accv);
      if (lane < nr) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < RT; c++) v = (lane == c) ? accv[c] : v;
        B[(i * NB + warp + 8 * rr) * nrhs + r0 + lane] = v;
      }
    }
    // publish: the barrier orders every thread's x_i stores before thread 0's fence (cumulative at
    // gpu scope), which orders them before the flag store -- one fence per block instead of 257
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      *reinterpret_cast<volatile int*>(flags + i) = epoch;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Many right-hand sides (nrhs > 16): blocked right-looking substitution on the fp64 tensor cores.
// One launch per 128-row block step k; CTA (column chunk of 16 right-hand sides, row tile i):
//     B_i -= F_ik X_k                                  (128x128 times 128x16, DMMA)
// and the CTA of the NEXT block of the sweep goes on to solve it with the stored inverse of its
// diagonal block, X_i = inv(F_ii) B_i, so the next launch finds X_{k+1} in place.  Step k = -1
// only solves the first block.  The flops (4 n^2 nrhs) run at tensor rate and every factor tile
// is read once per column chunk; the chain per block is one launch + two 128x128x16 products.
// ---------------------------------------------------------------------------------------
constexpr int RC = 16;        // right-hand sides per CTA
constexpr int XLD = RC + 4;   // row stride of the shared X / Y tiles (conflict-free B fragments)


// acc[rt][ct] (rows warp*16 + rt*8 + g, columns ct*8 + 2t, +1) += sign * A[128 x 128] * Xs[128 x RC]
__device__ __forceinline__ void tile_product(const double* __restrict__ A, int64_t lda, const double* __restrict__ Xs,
                                             double sign, double (&acc)[2][2][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double* A0 = A + (int64_t)(warp * 16 + g) * lda + t;
  const double* A1 = A0 + 8 * lda;
  // all A fragments of half the K range are requested before the first DMMA (latency-bound loads)
#pragma unroll 1
  for (int kh = 0; kh < NB / 4; kh += 16) {
    double a0[16], a1[16];
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0[u] = __ldg(A0 + (kh + u) * 4);
      a1[u] = __ldg(A1 + (kh + u) * 4);
    }
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int ks = kh + u;
      const double b0 = Xs[(ks * 4 + t) * XLD + g];
      const double b1 = Xs[(ks * 4 + t) * XLD + 8 + g];
      const double x0 = sign * a0[u], x1 = sign * a1[u];
      dmma_rhs(acc[0][0][0], acc[0][0][1], x0, b0);
      dmma_rhs(acc[0][1][0], acc[0][1][1], x0, b1);
      dmma_rhs(acc[1][0][0], acc[1][0][1], x1, b0);
      dmma_rhs(acc[1][1][0], acc[1][1][1], x1, b1);
    }
  }
}

__global__ void __launch_bounds__(256)
trsm_rhs_step_kernel(const double* __restrict__ F, int64_t ld, const double* __restrict__ dinv, int64_t nb,
                     int lower, int64_t k, int64_t nrhs, double* __restrict__ B) {
  __shared__ double Xs[NB * XLD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t c0 = (int64_t)blockIdx.x * RC;
  // row tile of this CTA: the blocks after k in sweep order (k < 0: only the first block is solved)
  const int64_t first = lower ? 0 : nb - 1;
  const int64_t i = k < 0 ? first : (lower ? k + 1 + blockIdx.y : k - 1 - (int64_t)blockIdx.y);
  const bool solve_here = blockIdx.y == 0;
  double acc[2][2][2];
  double* Bi = B + i * NB * nrhs;
  if (solve_here) {
    // the two tiles on the critical path of this and of the next step: into L2 ahead of use
    prefetch_tile_l2(dinv + i * 2 * NB * NB + (lower ? 0 : NB * NB), NB * sizeof(double), warp, lane);
    const int64_t inext = lower ? i + 1 : i - 1;
    if (inext >= 0 && inext < nb)
      prefetch_tile_l2(F + inext * NB * ld + i * NB, ld * (int64_t)sizeof(double), warp, lane);
  }
  // C fragments = B_i chunk (columns beyond nrhs read as zero and are never stored)
#pragma unroll
  for (int rt = 0; rt < 2; rt++)
#pragma unroll
    for (int ct = 0; ct < 2; ct++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int64_t c = c0 + ct * 8 + 2 * t + e;
        acc[rt][ct][e] = c < nrhs ? Bi[(int64_t)(warp * 16 + rt * 8 + g) * nrhs + c] : 0.0;
      }
  if (k >= 0) {
    const double* Bk = B + k * NB * nrhs;
    for (int idx = tid; idx < NB * RC; idx += 256) {
      const int r = idx / RC, c = idx % RC;
      Xs[r * XLD + c] = (c0 + c) < nrhs ? Bk[(int64_t)r * nrhs + c0 + c] : 0.0;
    }
    __syncthreads();
    tile_product(F + i * NB * ld + k * NB, ld, Xs, -1.0, acc);
  }
  if (solve_here) {
    __syncthreads();  // everyone is done reading Xs
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int ct = 0; ct < 2; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          Xs[(warp * 16 + rt * 8 + g) * XLD + ct * 8 + 2 * t + e] = acc[rt][ct][e];
          acc[rt][ct][e] = 0.0;
        }
    __syncthreads();
    tile_product(dinv + i * 2 * NB * NB + (lower ? 0 : NB * NB), NB, Xs, 1.0, acc);
  }
#pragma unroll
  for (int rt = 0; rt < 2; rt++)
#pragma unroll
    for (int ct = 0; ct < 2; ct++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int64_t c = c0 + ct * 8 + 2 * t + e;
        if (c < nrhs) Bi[(int64_t)(warp * 16 + rt * 8 + g) * nrhs + c] = acc[rt][ct][e];
      }
}

// ---------------------------------------------------------------------------------------
// 2..8 right-hand sides: the same persistent flag-driven sweep, but the two 128x128 products per
// block run on the fp64 tensor cores (mma.sync m8n8k4: the 8 right-hand sides are exactly one n
// tile), so there is no warp-shuffle reduction on the critical chain.  Warp w owns rows
// 16 w .. 16 w + 15 of the block; the A fragments of the factor tile are loaded straight from
// global memory before the ready flag of x_j is awaited.
// ---------------------------------------------------------------------------------------
constexpr int XP = 12;  // row stride of the shared x tile: conflict-free B fragments

__global__ void __launch_bounds__(256)
trsv_sweep_dmma_kernel(const double* __restrict__ F, int64_t ld, const double* __restrict__ dinv, int64_t nb,
                       int lower, int64_t nrhs, int64_t r0, double* __restrict__ B, int* __restrict__ flags,
                       int* __restrict__ counter, int epoch) {
  __shared__ double xs[NB * XP];
  __shared__ int s_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nr = (int)((nrhs - r0) < 8 ? (nrhs - r0) : 8);
  const int64_t ldb = ld * (int64_t)sizeof(double);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_p = atomicAdd(counter, 1);
    __syncthreads();
    const int p = s_p;  // position in sweep order
    if (p >= nb) break;
    const int64_t i = lower ? p : nb - 1 - p;
    const double* dblk = dinv + i * 2 * NB * NB + (lower ? 0 : NB * NB);
    prefetch_tile_l2(dblk, NB * sizeof(double), warp, lane);
    // C fragments: rows 16 warp + 8 rt + g, columns 2t, 2t + 1 of the right-hand side block
    double acc[2][2];
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int e = 0; e < 2; e++)
        acc[rt][e] = (2 * t + e) < nr ? B[(i * NB + warp * 16 + rt * 8 + g) * nrhs + r0 + 2 * t + e] : 0.0;
    if (p > 0) {
      const int64_t j0 = lower ? 0 : nb - 1;
      prefetch_tile_l2(F + i * NB * ld + j0 * NB, ldb, warp, lane);
    }
    for (int q = 0; q < p; q++) {
      const int64_t j = lower ? q : nb - 1 - q;
      // A fragments of the factor tile first (independent of x_j), next tile into L2, then wait
      double a[2][NB / 4];
#pragma unroll
      for (int rt = 0; rt < 2; rt++) {
        const double* Frow = F + (i * NB + warp * 16 + rt * 8 + g) * ld + j * NB + t;
#pragma unroll
        for (int ks = 0; ks < NB / 4; ks++) a[rt][ks] = Frow[ks * 4];
      }
      if (q + 1 < p) {
        const int64_t jn = lower ? q + 1 : nb - 2 - q;
        prefetch_tile_l2(F + i * NB * ld + jn * NB, ldb, warp, lane);
      }
      if (tid == 0) {
        while (*reinterpret_cast<volatile int*>(flags + j) != epoch) {
        }
        __threadfence();
      }
      __syncthreads();
      for (int idx = tid; idx < NB * 8; idx += 256) {
        const int r = idx >> 3, c = idx & 7;
        xs[r * XP + c] = c < nr ? __ldcg(&B[(j * NB + r) * nrhs + r0 + c]) : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < NB / 4; ks++) {
        const double b = xs[(ks * 4 + t) * XP + g];
        dmma_rhs(acc[0][0], acc[0][1], -a[0][ks], b);
        dmma_rhs(acc[1][0], acc[1][1], -a[1][ks], b);
      }
    }
    // finish: x_i = inv(F_ii) acc_i
    double da[2][NB / 4];
#pragma unroll
    for (int rt = 0; rt < 2; rt++) {
      const double* Drow = dblk + (int64_t)(warp * 16 + rt * 8 + g) * NB + t;
#pragma unroll
      for (int ks = 0; ks < NB / 4; ks++) da[rt][ks] = Drow[ks * 4];
    }
    __syncthreads();  // everyone is done with the x tile of the last product
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int e = 0; e < 2; e++) xs[(warp * 16 + rt * 8 + g) * XP + 2 * t + e] = acc[rt][e];
    __syncthreads();
    double out[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int ks = 0; ks < NB / 4; ks++) {
      const double b = xs[(ks * 4 + t) * XP + g];
      dmma_rhs(out[0][0], out[0][1], da[0][ks], b);
      dmma_rhs(out[1][0], out[1][1], da[1][ks], b);
    }
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int e = 0; e < 2; e++)
        if ((2 * t + e) < nr) B[(i * NB + warp * 16 + rt * 8 + g) * nrhs + r0 + 2 * t + e] = out[rt][e];
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      *reinterpret_cast<volatile int*>(flags + i) = epoch;
    }
  }
}

static int g_epoch = 0;
static int g_capacity[64][2] = {};  // resident CTAs per device for RT = 1 / 8

}  // namespace scb

using namespace scb;

extern "C" int scb_getrs_nopiv(int64_t n_pad, const double* LU, const double* dinv, int64_t nrhs,
                               double* B, scb_stream_t stream) {
  SCB_CHECK_ARG(n_pad > 0 && n_pad % NB == 0, "n_pad must be a positive multiple of 128");
  SCB_CHECK_ARG(nrhs > 0, "nrhs must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nb = n_pad / NB;
  int dev = 0;
  SCB_CUDA(cudaGetDevice(&dev));
  const int which = nrhs == 1 ? 0 : 1;
  int& cap = g_capacity[dev & 63][which];
  if (cap == 0) {
    int per_sm = 0, sms = 0;
    if (which == 0)
      SCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_sweep_kernel<1>, 256, 0));
    else
      SCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_sweep_dmma_kernel, 256, 0));
    SCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cap = per_sm * sms;
    if (cap < 1) cap = 1;
  }
  if (nrhs > 16) {
    // (up to 16 right-hand sides two passes of the flag-driven DMMA sweep are faster)
    // tensor-core path: one launch per block step and sweep direction
    const unsigned ncc = (unsigned)((nrhs + RC - 1) / RC);
    for (int lower = 1; lower >= 0; lower--) {
      trsm_rhs_step_kernel<<<dim3(ncc, 1), 256, 0, s>>>(LU, n_pad, dinv, nb, lower, -1, nrhs, B);
      SCB_LAUNCH_CHECK();
      for (int64_t q = 0; q + 1 < nb; q++) {
        const int64_t k = lower ? q : nb - 1 - q;
        trsm_rhs_step_kernel<<<dim3(ncc, (unsigned)(nb - 1 - q)), 256, 0, s>>>(LU, n_pad, dinv, nb, lower, k, nrhs, B);
        SCB_LAUNCH_CHECK();
      }
    }
    return SCB_OK;
  }
  // integer scratch behind the packed panels of the factorization workspace
  int* flags = reinterpret_cast<int*>(const_cast<double*>(dinv) + lu_flags_offset(n_pad));
  int* counters = flags + nb;  // [0]: work counter
  const int grid = (int)(nb < cap ? nb : cap);
  const int RTv = which == 0 ? 1 : 8;
  for (int lower = 1; lower >= 0; lower--) {
    for (int64_t r0 = 0; r0 < nrhs; r0 += RTv) {
      const int epoch = ++g_epoch;
      SCB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int), s));
      if (which == 0)
        trsv_sweep_kernel<1><<<grid, 256, 0, s>>>(LU, n_pad, dinv, nb, lower, nrhs, r0, B, flags, counters, epoch);
      else
        trsv_sweep_dmma_kernel<<<grid, 256, 0, s>>>(LU, n_pad, dinv, nb, lower, nrhs, r0, B, flags, counters, epoch);
      SCB_LAUNCH_CHECK();
    }
  }
  return SCB_OK;
}
