// Two-level blocked right-looking fp64 LU without pivoting (K11) for the row-diagonally-dominant
// systems of this path, row-major, padded to a multiple of NB = 128.
//
// Outer panels of KB = q*128 columns (q = 8 by default); inside an outer panel, per inner block
// step k (offset o = 128 k):
//   1. diag_kernel   (1 CTA)        LU of the 128x128 diagonal block held in registers, fused with
//                                   the explicit inverses inv(L_kk), inv(U_kk) (kept for getrs);
//   2. trsm_kernel   (4*rem/128 CTAs) L21 = A21 inv(U_kk), U12 = inv(L_kk) A12 as DMMA GEMMs;
//                                   results are written in place AND as "fragment-major" packed
//                                   copies (-L21 and U12) laid out exactly as the mma.sync m8n8k4
//                                   A/B register fragments, into chunk slot k of the outer panel;
//   3. update_kernel on the L-shaped strip of the outer panel only (K = 128).
// Once per outer panel: update_kernel on the big trailing block with K = KB -- the only O(n^3)
// contraction.  Operand chunks arrive by TMA bulk copies (cp.async.bulk + mbarrier, one 32 KB +
// one 16 KB copy per stage) and are consumed with conflict-free LDS.64 straight into DMMA.8x8x4;
// C tiles are read into the accumulators and written back with 128-bit accesses.  Two resident
// CTAs per SM overlap one CTA's C traffic with the other's DMMA work; K = KB amortises the C
// traffic and CTA prologue/epilogue over q times more math than a plain rank-128 update.
//
// Symmetric variant (scb_getrf_sym_nopiv): for a constant Lambda the system is diagonally similar
// (D = W^1/2) to a symmetric matrix S.  Every Schur complement of S is symmetric, so only the tiles
// that intersect the lower triangle are updated (update_kernel_t<true>: CTAs of upper tiles exit
// at once), only the column panel is solved (trsm_sym_kernel), and the row panel follows from
// U12 = diag(U11) L21^T -- half the flops of step 3 / the bulk update.  The output is still an
// ordinary (L, U) pair in M, so scb_getrs_nopiv and lu_piv are unchanged.
//
// Reference semantics: scipy.linalg.lu_factor(-A) at solver/solve_film.py:232,253,279 (LAPACK
// dgetrf).  Pivoting is unnecessary here; parity is on the solution (1e-8 rel-L2), see DESIGN.md.
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "scb_common.cuh"

namespace scb {

constexpr int NB = SCB_LU_BLOCK;  // 128
constexpr int BM = 128;           // update tile rows
constexpr int BN = 64;            // update tile cols
constexpr int KC = 32;            // k chunk
constexpr int NCHUNK = NB / KC;   // 4
constexpr int A_CHUNK = BM * KC;  // doubles per packed A chunk (32 KB)
constexpr int B_CHUNK = KC * BN;  // doubles per packed B chunk (16 KB)

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
  }
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// packed ("fragment-major") operand layouts
//   Lpack tile (128 rows x 128 k):  [chunk c(4)][row block rb(16)][k4 step s(8)][lane(32)]
//        lane = (row%8)*4 + k%4 holds  -L[rb*8 + row%8][c*32 + s*4 + k%4]
//   Upack tile (128 k x 64 cols):   [chunk c(4)][k4 step s(8)][col block nb(8)][lane(32)]
//        lane = (col%8)*4 + k%4 holds   U[c*32 + s*4 + k%4][nb*8 + col%8]
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t lpack_index(int r /*0..127*/, int k /*0..127*/) {
  return (int64_t)(k >> 5) * A_CHUNK + (((r >> 3) * 8 + ((k & 31) >> 2)) * 32 + (r & 7) * 4 + (k & 3));
}
__device__ __forceinline__ int64_t upack_index(int k /*0..127*/, int c /*0..63*/) {
  return (int64_t)(k >> 5) * B_CHUNK + ((((k & 31) >> 2) * 8 + (c >> 3)) * 32 + (c & 7) * 4 + (k & 3));
}

// ---------------------------------------------------------------------------------------
// 3. trailing update
// ---------------------------------------------------------------------------------------
struct __align__(128) UpdateStage {
  double a[A_CHUNK];
  double b[B_CHUNK];
};

template <bool TRI>
__global__ void __launch_bounds__(256, 2)
update_kernel_t(double* __restrict__ M, int64_t ld, int64_t row0, int64_t col0,
                const double* __restrict__ Lpack, const double* __restrict__ Upack, int tile_chunks,
                int chunk0, int nchunks, int kBand) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  UpdateStage* stage = reinterpret_cast<UpdateStage*>(smem_raw);
  __shared__ uint64_t bars[2];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;  // 4 x 2 warps, 32x32 warp tiles
  const int g = lane >> 2, t = lane & 3;

  // tile coordinates.  TRI: the region is square and only the tiles that intersect its lower
  // triangle are updated (row tile by owns column tiles 0 .. 2 by + 1); the others exit at once.
  // Raster order: the hardware issues CTAs with blockIdx.x fastest, i.e. one row tile (one L tile)
  // against ALL column tiles -- at 20k that is 280 U tiles = 143 MB per row, more than the L2 holds, so
  // every U tile came from DRAM once per row tile (10.1 GB per launch for 2.9 GB algorithmic).  The linear
  // CTA index is therefore re-mapped to bands of kBand (16) row tiles, column-major inside a band: the
  // resident wave works on kBand L tiles (16 MB) and ~20 U tiles, and a U tile is fetched from DRAM once
  // per band.  Same tiles, same arithmetic per tile: results are bit-identical.
  int by = blockIdx.y, bx = blockIdx.x;
  if (kBand > 1 && (int)gridDim.y > 1) {
    const int gx = gridDim.x, gy = gridDim.y;
    const int64_t id = (int64_t)blockIdx.y * gx + blockIdx.x;
    const int band = (int)(id / ((int64_t)kBand * gx));
    const int band_rows = (gy - band * kBand) < kBand ? (gy - band * kBand) : kBand;
    const int rem = (int)(id - (int64_t)band * kBand * gx);
    bx = rem / band_rows;
    by = band * kBand + rem % band_rows;
  }
  if (TRI && bx > 2 * by + 1) return;
  // packed operands are indexed by ABSOLUTE 128-row / 64-column tile and by chunk slot
  const double* Ltile = Lpack + ((row0 >> 7) + by) * ((int64_t)tile_chunks * A_CHUNK) + (int64_t)chunk0 * A_CHUNK;
  const double* Utile = Upack + ((col0 >> 6) + bx) * ((int64_t)tile_chunks * B_CHUNK) + (int64_t)chunk0 * B_CHUNK;
  constexpr uint32_t kStageBytes = (A_CHUNK + B_CHUNK) * sizeof(double);

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      mbar_expect_tx(&bars[c], kStageBytes);
      bulk_g2s(stage[c].a, Ltile + (int64_t)c * A_CHUNK, A_CHUNK * sizeof(double), &bars[c]);
      bulk_g2s(stage[c].b, Utile + (int64_t)c * B_CHUNK, B_CHUNK * sizeof(double), &bars[c]);
    }
  }

  // accumulators start as the C tile
  double acc[4][4][2];
  double* Cbase = M + (row0 + (int64_t)by * BM + wm * 32 + g) * ld + col0 + (int64_t)bx * BN + wn * 32 + 2 * t;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double2 v = *reinterpret_cast<const double2*>(Cbase + (int64_t)(i * 8) * ld + j * 8);
      acc[i][j][0] = v.x;
      acc[i][j][1] = v.y;
    }

#pragma unroll 1
  for (int c = 0; c < nchunks; c++) {
    const int st = c & 1;
    mbar_wait(&bars[st], (c >> 1) & 1);
    const double* As = stage[st].a + (wm * 4 * 8) * 32 + lane;
    const double* Bs = stage[st].b + (wn * 4) * 32 + lane;
#pragma unroll
    for (int s = 0; s < 8; s++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[(i * 8 + s) * 32];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[(s * 8 + j) * 32];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    // The stage that was just read through the generic proxy (LDS) is about to be overwritten through
    // the async proxy (the next bulk copy): accesses through different proxies are NOT ordered by
    // bar.sync alone.  Every reader therefore issues a generic->async proxy fence before the barrier.
    // Without it the kernel was not reproducible: rarely (a few tiles per launch at 20k, more often for
    // the short K = 128 launches) the operand fragments of a warp came out wrong in units of 32-byte
    // shared-memory sectors, i.e. LU entries off by 1e-8 .. 1e-5 relative and run-to-run different
    // factors (tools/lu_kernel_determinism.cu: identical inputs, bitwise comparison; 0 differences
    // in every configuration with the fence, hundreds to thousands of words without it).
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (tid == 0 && c + 2 < nchunks) {
      mbar_expect_tx(&bars[st], kStageBytes);
      bulk_g2s(stage[st].a, Ltile + (int64_t)(c + 2) * A_CHUNK, A_CHUNK * sizeof(double), &bars[st]);
      bulk_g2s(stage[st].b, Utile + (int64_t)(c + 2) * B_CHUNK, B_CHUNK * sizeof(double), &bars[st]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      *reinterpret_cast<double2*>(Cbase + (int64_t)(i * 8) * ld + j * 8) =
          make_double2(acc[i][j][0], acc[i][j][1]);
}

// ---------------------------------------------------------------------------------------
// 2. panel solves as GEMMs with the explicit block inverses (K = 128, one pass per CTA)
//    blockIdx.x <  ncol : L21 tile (64 rows x 128)  = A21 tile * invU      warps 2 x 4
//    blockIdx.x >= ncol : U12 tile (128 x 64 cols)  = invL * A12 tile      warps 4 x 2
//    Each CTA reads only the rows (col panel) / columns (row panel) it overwrites, so the
//    in-place update is race-free.
// ---------------------------------------------------------------------------------------
constexpr int TA_LD = KC + 4;  // 36: A chunk [TM][36]   (ld % 16 == 4 -> conflict-free fragments)

// TRI = 1: B is upper triangular (B[k][n] = 0 for k > n): a warp skips chunk c when c > wn.
// TRI = 2: A is lower triangular (A[m][k] = 0 for k > m): a warp skips chunk c when c > wm.
template <int TM, int TN, int WMW, int WNW, int TRI>
__device__ __forceinline__ void gemm_k128(const double* __restrict__ A, int64_t lda,
                                          const double* __restrict__ B, int64_t ldb, double* As,
                                          double* Bs, double (&acc)[4][4][2]) {
  constexpr int TB_LD = TN + 4;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WNW, wn = warp % WNW;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  // The operands of chunk c + 1 are on their way while chunk c feeds the tensor cores (one round trip to
  // L2 / HBM per CTA instead of four back to back): the A rows through registers, the B rows (the block
  // inverse, shared by all CTAs of the launch) by cp.async into the second of two shared buffers.  Same
  // fragments, same k order: bit-identical results.  Bs points to 2 x KC x TB_LD doubles.
  constexpr int QA = TM * 16 / 256;        // double2 of an A chunk per thread
  constexpr int QB = KC * (TN / 2) / 256;  // 16-byte granules of a B chunk per thread
  double2 ra[QA];
  auto fetch_a = [&](int c) {
#pragma unroll
    for (int q = 0; q < QA; q++) {
      const int idx = q * 256 + tid;
      const int r = idx >> 4, kk = (idx & 15) * 2;
      ra[q] = *reinterpret_cast<const double2*>(A + (int64_t)r * lda + c * KC + kk);
    }
  };
  auto issue_b = [&](int c) {
    double* dst = Bs + (c & 1) * (KC * TB_LD);
#pragma unroll
    for (int q = 0; q < QB; q++) {
      const int idx = q * 256 + tid;
      const int kr = idx / (TN / 2), cc = (idx % (TN / 2)) * 2;
      cp_async16(dst + kr * TB_LD + cc, B + (int64_t)(c * KC + kr) * ldb + cc);
    }
    cp_async_commit();
  };
  issue_b(0);
  fetch_a(0);
#pragma unroll 1
  for (int c = 0; c < NCHUNK; c++) {
    __syncthreads();  // the previous chunk has been consumed: As and the other B buffer are free
#pragma unroll
    for (int q = 0; q < QA; q++) {
      const int idx = q * 256 + tid;
      const int r = idx >> 4, kk = (idx & 15) * 2;
      *reinterpret_cast<double2*>(&As[r * TA_LD + kk]) = ra[q];
    }
    if (c + 1 < NCHUNK) issue_b(c + 1);
    cp_async_wait_pending(c + 1 < NCHUNK ? 1 : 0);
    __syncthreads();
    if (c + 1 < NCHUNK) fetch_a(c + 1);
    if ((TRI == 1 && c > wn) || (TRI == 2 && c > wm)) continue;  // structurally zero block
    const double* Bc = Bs + (c & 1) * (KC * TB_LD);
#pragma unroll
    for (int s = 0; s < 8; s++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[(wm * 32 + i * 8 + g) * TA_LD + s * 4 + t];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bc[(s * 4 + t) * TB_LD + wn * 32 + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
}

constexpr int kTrsmSmemDoubles = 128 * TA_LD + 2 * KC * (64 + 4) > 64 * TA_LD + 2 * KC * (128 + 4)
                                     ? 128 * TA_LD + 2 * KC * (64 + 4)
                                     : 64 * TA_LD + 2 * KC * (128 + 4);  // A chunk + two B chunk buffers

__global__ void __launch_bounds__(256)
trsm_kernel(double* __restrict__ M, int64_t ld, int64_t o, int ncol, const double* __restrict__ invL,
            const double* __restrict__ invU, double* __restrict__ Lpack, double* __restrict__ Upack,
            int tile_chunks, int chunk0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t o2 = o + NB;
  double acc[4][4][2];
  if ((int)blockIdx.x < ncol) {
    // ---- column panel: 64 rows of L21 ----
    const int tile = blockIdx.x;
    double* Atile = M + (o2 + (int64_t)tile * 64) * ld + o;
    double* Bs = As + 64 * TA_LD;
    gemm_k128<64, 128, 2, 4, 1>(Atile, ld, invU, NB, As, Bs, acc);
    const int wm = warp / 4, wn = warp % 4;
    double* P = Lpack + ((o2 >> 7) + (tile >> 1)) * ((int64_t)tile_chunks * A_CHUNK) + (int64_t)chunk0 * A_CHUNK;
    const int rbase = (tile & 1) * 64;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = wm * 32 + i * 8 + g;
        const int cc = wn * 32 + j * 8 + 2 * t;
        *reinterpret_cast<double2*>(Atile + (int64_t)r * ld + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
        P[lpack_index(rbase + r, cc)] = -acc[i][j][0];
        P[lpack_index(rbase + r, cc + 1)] = -acc[i][j][1];
      }
  } else {
    // ---- row panel: 64 columns of U12 ----
    const int tile = blockIdx.x - ncol;
    double* Btile = M + o * ld + o2 + (int64_t)tile * 64;
    double* Bs = As + 128 * TA_LD;
    gemm_k128<128, 64, 4, 2, 2>(invL, NB, Btile, ld, As, Bs, acc);
    const int wm = warp / 2, wn = warp % 2;
    double* P = Upack + ((o2 >> 6) + tile) * ((int64_t)tile_chunks * B_CHUNK) + (int64_t)chunk0 * B_CHUNK;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = wm * 32 + i * 8 + g;
        const int cc = wn * 32 + j * 8 + 2 * t;
        *reinterpret_cast<double2*>(Btile + (int64_t)r * ld + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
        P[upack_index(r, cc)] = acc[i][j][0];
        P[upack_index(r, cc + 1)] = acc[i][j][1];
      }
  }
}

// Symmetric variant: only the column panel is solved (64 rows per CTA); the row panel follows from
// symmetry, U12 = diag(U11) L21^T, and is emitted here as the packed B operand of the update kernel
// and, transposed, into the upper triangle of M (so getrs and lu_piv see an ordinary LU).
__global__ void __launch_bounds__(256)
trsm_sym_kernel(double* __restrict__ M, int64_t ld, int64_t o, const double* __restrict__ invU,
                double* __restrict__ Lpack, double* __restrict__ Upack, int tile_chunks, int chunk0, int tile0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);
  __shared__ double dU[NB];  // diagonal of U11
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t o2 = o + NB;
  if (tid < NB) dU[tid] = M[(o + tid) * ld + o + tid];
  double acc[4][4][2];
  const int tile = blockIdx.x + tile0;  // 64-row tile below the diagonal block
  double* Atile = M + (o2 + (int64_t)tile * 64) * ld + o;
  double* Bs = As + 64 * TA_LD;
  gemm_k128<64, 128, 2, 4, 1>(Atile, ld, invU, NB, As, Bs, acc);  // ends after a __syncthreads: dU visible
  const int wm = warp / 4, wn = warp % 4;
  double* PL = Lpack + ((o2 >> 7) + (tile >> 1)) * ((int64_t)tile_chunks * A_CHUNK) + (int64_t)chunk0 * A_CHUNK;
  double* PU = Upack + ((o2 >> 6) + tile) * ((int64_t)tile_chunks * B_CHUNK) + (int64_t)chunk0 * B_CHUNK;
  const int rbase = (tile & 1) * 64;
  double* Urow0 = M + o * ld + o2 + (int64_t)tile * 64;  // U12[k][c]: row o + k, column o2 + 64 tile + c
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = wm * 32 + i * 8 + g;       // row inside this 64-row tile
      const int cc = wn * 32 + j * 8 + 2 * t;  // k index 0..127
      const double v0 = acc[i][j][0], v1 = acc[i][j][1];
      *reinterpret_cast<double2*>(Atile + (int64_t)r * ld + cc) = make_double2(v0, v1);
      PL[lpack_index(rbase + r, cc)] = -v0;
      PL[lpack_index(rbase + r, cc + 1)] = -v1;
      const double u0 = dU[cc] * v0, u1 = dU[cc + 1] * v1;
      PU[upack_index(cc, r)] = u0;
      PU[upack_index(cc + 1, r)] = u1;
      Urow0[(int64_t)cc * ld + r] = u0;
      Urow0[(int64_t)(cc + 1) * ld + r] = u1;
    }
}

// ---------------------------------------------------------------------------------------
// Latency variants of the two GEMM kernels for the launches that sit on the per-block critical chain
// (the tiles inside an outer panel's diagonal square: at most 14 panel tiles and 56 update tiles of
// 128 x 64, i.e. far fewer CTAs than SMs, each of them a K = 128 pass that the full-size kernels run as
// one un-overlapped load -> DMMA -> store sequence per CTA).  Same arithmetic, finer tiles:
//   * update_lat_kernel: C tiles of 64 x 32 (four per tile of update_kernel_t, 128 threads, warp tile
//     16 x 32), the packed operand chunks streamed through a 4-stage cp.async ring;
//   * trsm_sym_lat_kernel: 16 rows of a 64-row panel tile per CTA (row split: a CTA reads and overwrites
//     only its own rows, so the in-place update stays race-free), operands by cp.async in one commit
//     group per k chunk.
// Every accumulator sees the same fragments in the same k order as in the full-size kernels: the factors
// are bit-identical (checked by tests/test_gpu_round2.py::test_lu_latency_kernels_bit_identical).
// ---------------------------------------------------------------------------------------
constexpr int LAT_STAGES = 4;
constexpr int LAT_A = (A_CHUNK / 2);  // doubles of half an A chunk: 8 row blocks x 8 k4 steps x 32 lanes
constexpr int LAT_B = (B_CHUNK / 2);  // doubles of half a B chunk: 8 k4 steps x 4 column blocks x 32 lanes
constexpr int kUpdLatSmem = LAT_STAGES * (LAT_A + LAT_B) * (int)sizeof(double);  // 96 KB

template <bool TRI>
__global__ void __launch_bounds__(128)
update_lat_kernel(double* __restrict__ M, int64_t ld, int64_t row0, int64_t col0, const double* __restrict__ Lpack,
                  const double* __restrict__ Upack, int tile_chunks, int chunk0, int nchunks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);  // [stage][8 rb][8 s][32]
  double* Bs = As + LAT_STAGES * LAT_A;               // [stage][8 s][4 nb][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ty = blockIdx.y, tx = blockIdx.x;  // 64-row / 32-column tile
  const int by = ty >> 1, bx = tx >> 1;        // the 128 x 64 tile of update_kernel_t it belongs to
  if (TRI && bx > 2 * by + 1) return;          // same element set as update_kernel_t<true>
  const double* Ltile = Lpack + ((row0 >> 7) + by) * ((int64_t)tile_chunks * A_CHUNK) + (int64_t)chunk0 * A_CHUNK +
                        (ty & 1) * LAT_A;
  const double* Utile = Upack + ((col0 >> 6) + bx) * ((int64_t)tile_chunks * B_CHUNK) + (int64_t)chunk0 * B_CHUNK +
                        (tx & 1) * 4 * 32;
  auto issue = [&](int c) {
    double* a = As + (c % LAT_STAGES) * LAT_A;
    double* b = Bs + (c % LAT_STAGES) * LAT_B;
    const double* ga = Ltile + (int64_t)c * A_CHUNK;
    const double* gb = Utile + (int64_t)c * B_CHUNK;
#pragma unroll
    for (int q = 0; q < LAT_A / 2 / 128; q++) {  // 8 granules of 16 bytes per thread
      const int idx = q * 128 + tid;
      cp_async16(a + 2 * idx, ga + 2 * idx);
    }
#pragma unroll
    for (int q = 0; q < LAT_B / 2 / 128; q++) {  // 4 per thread: 8 segments (one per k4 step) of 1 KB
      const int idx = q * 128 + tid;
      const int sseg = idx >> 6, w2 = (idx & 63) * 2;
      cp_async16(b + sseg * 128 + w2, gb + sseg * 256 + w2);
    }
    cp_async_commit();
  };
  const int pre = nchunks < LAT_STAGES ? nchunks : LAT_STAGES;
  for (int c = 0; c < pre; c++) issue(c);

  double acc[2][4][2];
  double* Cbase = M + (row0 + (int64_t)ty * 64 + warp * 16 + g) * ld + col0 + (int64_t)tx * 32 + 2 * t;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double2 v = *reinterpret_cast<const double2*>(Cbase + (int64_t)(i * 8) * ld + j * 8);
      acc[i][j][0] = v.x;
      acc[i][j][1] = v.y;
    }
#pragma unroll 1
  for (int c = 0; c < nchunks; c++) {
    const int issued = (c + LAT_STAGES) < nchunks ? (c + LAT_STAGES) : nchunks;
    cp_async_wait_pending(issued - c - 1);
    __syncthreads();
    const double* a_s = As + (c % LAT_STAGES) * LAT_A + (warp * 2 * 8) * 32 + lane;
    const double* b_s = Bs + (c % LAT_STAGES) * LAT_B + lane;
#pragma unroll
    for (int sk = 0; sk < 8; sk++) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; i++) a[i] = a_s[(i * 8 + sk) * 32];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = b_s[(sk * 4 + j) * 32];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (c + LAT_STAGES < nchunks) {
      __syncthreads();  // every warp has read the stage that is refilled now
      issue(c + LAT_STAGES);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      *reinterpret_cast<double2*>(Cbase + (int64_t)(i * 8) * ld + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
}

constexpr int LT_ALD = NB + 4;  // 132: row stride of the shared A rows / inv(U) rows (== 4 mod 16: conflict-free)
constexpr int kTrsmLatSmem = (16 * LT_ALD + NB * LT_ALD) * (int)sizeof(double);  // 152 KB

__global__ void __launch_bounds__(128)
trsm_sym_lat_kernel(double* __restrict__ M, int64_t ld, int64_t o, const double* __restrict__ invU,
                    double* __restrict__ Lpack, double* __restrict__ Upack, int tile_chunks, int chunk0, int tile0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* As = reinterpret_cast<double*>(smem_raw);  // [16][132]   rows of A21
  double* Bs = As + 16 * LT_ALD;                      // [128][132]  inv(U11) (upper triangle by 32-blocks)
  __shared__ double dU[NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t o2 = o + NB;
  const int tile = tile0 + (int)(blockIdx.x >> 2);  // 64-row tile below the diagonal block
  const int q = blockIdx.x & 3;                      // 16-row quarter of it
  double* Atile = M + (o2 + (int64_t)tile * 64) * ld + o;
  const double* Arows = Atile + (int64_t)(q * 16) * ld;
  dU[tid] = M[(o + tid) * ld + o + tid];
  // one commit group per k chunk c: columns 32c .. 32c+31 of the 16 A rows, rows 32c .. 32c+31 of inv(U)
  // right of their diagonal 32-block (the rest of inv(U) is structurally zero)
#pragma unroll
  for (int c = 0; c < NCHUNK; c++) {
#pragma unroll
    for (int k = 0; k < 2; k++) {  // A: 16 rows x 16 granules
      const int idx = k * 128 + tid;
      const int r = idx >> 4, cc = c * KC + (idx & 15) * 2;
      cp_async16(As + r * LT_ALD + cc, Arows + (int64_t)r * ld + cc);
    }
    const int gran_per_row = (NB - c * KC) / 2;  // 64, 48, 32, 16
    for (int idx = tid; idx < KC * gran_per_row; idx += 128) {
      const int r = c * KC + idx / gran_per_row, cc = c * KC + (idx % gran_per_row) * 2;
      cp_async16(Bs + r * LT_ALD + cc, invU + r * NB + cc);
    }
    cp_async_commit();
  }
  const int wn = warp;  // column group: k index 32 wn .. 32 wn + 31 of the result
  double acc[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
  for (int c = 0; c < NCHUNK; c++) {
    cp_async_wait_pending(NCHUNK - 1 - c);
    __syncthreads();
    if (c > wn) continue;  // structurally zero block of inv(U)
#pragma unroll
    for (int sk = 0; sk < 8; sk++) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; i++) a[i] = As[(i * 8 + g) * LT_ALD + c * KC + sk * 4 + t];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[(c * KC + sk * 4 + t) * LT_ALD + wn * 32 + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  __syncthreads();  // (all reads of this CTA's rows are complete before any of them is overwritten)
  double* PL = Lpack + ((o2 >> 7) + (tile >> 1)) * ((int64_t)tile_chunks * A_CHUNK) + (int64_t)chunk0 * A_CHUNK;
  double* PU = Upack + ((o2 >> 6) + tile) * ((int64_t)tile_chunks * B_CHUNK) + (int64_t)chunk0 * B_CHUNK;
  const int rbase = (tile & 1) * 64;
  double* Urow0 = M + o * ld + o2 + (int64_t)tile * 64;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = q * 16 + i * 8 + g;        // row inside the 64-row tile
      const int cc = wn * 32 + j * 8 + 2 * t;  // k index 0..127
      const double v0 = acc[i][j][0], v1 = acc[i][j][1];
      *reinterpret_cast<double2*>(Atile + (int64_t)r * ld + cc) = make_double2(v0, v1);
      PL[lpack_index(rbase + r, cc)] = -v0;
      PL[lpack_index(rbase + r, cc + 1)] = -v1;
      const double u0 = dU[cc] * v0, u1 = dU[cc + 1] * v1;
      PU[upack_index(cc, r)] = u0;
      PU[upack_index(cc + 1, r)] = u1;
      Urow0[(int64_t)cc * ld + r] = u0;
      Urow0[(int64_t)(cc + 1) * ld + r] = u1;
    }
}

// ---------------------------------------------------------------------------------------
// 1. diagonal block: LU (no pivoting) + explicit inverses of both factors in ONE sweep.
//    The 128x128 block lives in REGISTERS, distributed 2-D cyclically over 512 threads
//    (thread (ti, tj) = (warp, lane) owns rows ti + 16a, a < 8, and columns tj + 32b, b < 4).
//    Step j broadcasts register row j and register column j through double-buffered shared
//    vectors (one __syncthreads per step) and applies rank-1 updates.  Register slots are
//    re-used as soon as their LU value is final:
//      slot (r, c), r > c : A -> L[r][c] (final at step c, copied out) -> inv(L)[r][c]
//      slot (r, c), r < c : A -> U[r][c] (final at step r, copied out) -> Z[c][r], where Z is
//                           the unit lower factor of U^T = Z^-1 D, so inv(U) = Z^T D^-1.
//    Forward elimination applied to the identity yields inv(L); the same elimination applied
//    to U^T (multipliers U[j][c] / U[j][j], i.e. the pivot row already being broadcast) yields Z.
// ---------------------------------------------------------------------------------------
constexpr int DLD = NB + 1;  // padded row stride of the shared copy of the factors

__device__ __forceinline__ double fast_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

// seed r0 (relative error e = 1 - a r0, |e| ~ 2^-23) and one third-order step: 1/a = r0 (1 + e + e^2 + O(e^3)),
// three dependent FMAs instead of four
__device__ __forceinline__ double fast_rcp_halley(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// two Newton steps (what the compiler's own fp64 division uses on the MUFU.RCP64H seed)
__device__ __forceinline__ double fast_rcp2(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
}

__global__ void __launch_bounds__(512, 1)
diag_kernel(double* __restrict__ M, int64_t ld, int64_t o, double* __restrict__ invL,
            double* __restrict__ invU, int32_t* __restrict__ info, int block_index) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* D = reinterpret_cast<double*>(smem_raw);  // [128][129] finished L / U entries
  __shared__ double rowb[2][NB];                    // register row j
  __shared__ double colb[2][NB];                    // register column j
  __shared__ double rdiag[NB];                      // 1 / U[j][j]
  const int tid = threadIdx.x;
  const int ti = tid >> 5, tj = tid & 31;
  double* blk = M + o * ld + o;
  double x[8][4];
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) x[a][b] = blk[(int64_t)(ti + 16 * a) * ld + tj + 32 * b];
  if (ti == 0) {
#pragma unroll
    for (int b = 0; b < 4; b++) rowb[0][tj + 32 * b] = x[0][b];
  }
  if (tj == 0) {
#pragma unroll
    for (int a = 0; a < 8; a++) colb[0][ti + 16 * a] = x[a][0];
  }
  __syncthreads();
  int bad = 0;
  // j = 16*ap + jj with the outer index unrolled: every register-array index below is a
  // compile-time constant (dynamic indexing would push the block into local memory).
#pragma unroll
  for (int ap = 0; ap < 8; ap++) {
#pragma unroll 1
    for (int jj = 0; jj < 16; jj++) {
      const int j = 16 * ap + jj;
      const int cur = j & 1, nxt = cur ^ 1;
      const double piv = rowb[cur][j];
      if (bad == 0 && !(fabs(piv) > 0.0 && isfinite(piv))) bad = j + 1;
      const double rp = fast_rcp(piv);
      if (tid == 0) rdiag[j] = rp;
      const int bj = ap >> 1;  // column block of column j (compile-time after unrolling)
      const bool own_col = tj == (j & 31);
      double v[4], vs[4];
#pragma unroll
      for (int b = 0; b < 4; b++) {
        v[b] = rowb[cur][tj + 32 * b];  // c > j: U[j][c];  c < j: inv(L)[j][c]
        vs[b] = v[b] * rp;              // multipliers of the U^T elimination (c > j)
      }
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int r = ti + 16 * a;
        const double w = colb[cur][r];  // r > j: A[r][j];  r < j: Z[j][r]
        if (a > ap || (a == ap && r > j)) {
          // rows below the pivot: A update (c > j), inv(L) update (c < j), L output (c == j)
          const double l = w * rp;
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (b == bj && own_col) {
              D[r * DLD + j] = l;
              x[a][b] = -l;
            } else {
              x[a][b] = fma(-l, v[b], x[a][b]);
            }
          }
        } else if (a < ap || (a == ap && r < j)) {
          // rows above the pivot: Z update on the columns right of the pivot
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (b < bj) continue;
            if (b > bj || tj + 32 * b > j) x[a][b] = fma(-w, vs[b], x[a][b]);
          }
        } else {
          // the pivot row itself: U output, then seed Z[c][j] = -U[j][c] / U[j][j]
#pragma unroll
          for (int b = 0; b < 4; b++) {
            if (b < bj) continue;
            const int c = tj + 32 * b;
            if (b > bj || c >= j) D[j * DLD + c] = x[a][b];
            if (b > bj || c > j) x[a][b] = -vs[b];
          }
        }
      }
      const int jn = j + 1;
      if (jn < NB) {
        if (ti == (jn & 15)) {
          if (jj < 15) {
#pragma unroll
            for (int b = 0; b < 4; b++) rowb[nxt][tj + 32 * b] = x[ap][b];
          } else {
#pragma unroll
            for (int b = 0; b < 4; b++) rowb[nxt][tj + 32 * b] = x[ap + 1 < 8 ? ap + 1 : 7][b];
          }
        }
        if (tj == (jn & 31)) {
          if (jj < 15 || ((ap + 1) >> 1) == bj) {
#pragma unroll
            for (int a = 0; a < 8; a++) colb[nxt][ti + 16 * a] = x[a][bj];
          } else {
#pragma unroll
            for (int a = 0; a < 8; a++) colb[nxt][ti + 16 * a] = x[a][bj + 1 < 4 ? bj + 1 : 3];
          }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0 && bad) atomicCAS(info, 0, block_index * NB + bad);
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int r = ti + 16 * a, c = tj + 32 * b;
      blk[(int64_t)r * ld + c] = D[r * DLD + c];
      const double v = x[a][b];
      invL[r * NB + c] = c < r ? v : (c == r ? 1.0 : 0.0);
      invU[r * NB + c] = c > r ? v * rdiag[c] : (c == r ? rdiag[c] : 0.0);
    }
}

// ---------------------------------------------------------------------------------------
// 1b. diagonal block, small-footprint version (256 threads, 3 x 34 KB shared): same outputs as
//     diag_kernel, but organised as a 2x2 recursion over 64x64 quadrants so that one CTA fits in
//     the slot of an update CTA (256 threads, <= 128 registers, < 113 KB shared).  This is what
//     allows the panel factorization of the next outer panel to run on a high-priority stream
//     CONCURRENTLY with the big trailing update (look-ahead): the hardware block scheduler only
//     places a high-priority CTA into a slot it fits in.
//       A11 -> sweep64 -> L11, U11, inv(L11), inv(U11)
//       U12 = inv(L11) A12,  L21 = A21 inv(U11),  A22 -= L21 U12           (64^3 DMMA GEMMs)
//       A22 -> sweep64 -> L22, U22, inv(L22), inv(U22)
//       inv(L)21 = -inv(L22) L21 inv(L11),  inv(U)12 = -inv(U11) U12 inv(U22)
// ---------------------------------------------------------------------------------------
#ifndef SCB_STAMP
#define SCB_STAMP(i)  // phase time stamps, only defined by tools/diag_phases.cu
#endif
constexpr int QN = 64;        // quadrant size
constexpr int QLD = QN + 4;   // 68: conflict-free DMMA fragment loads (ld % 16 == 4)

// fused LU + inv(L) + inv(U)^T-elimination sweep on a 64x64 block held in registers:
// thread (ti = warp 0..7, tj = lane) owns rows ti + 8a (a < 8) and columns tj + 32b (b < 2).
// Dout[64][QLD] receives the finished L (strictly lower) and U (upper) entries.
// SYM: the block is symmetric, so inv(U) = inv(L)^T D^-1 is derived afterwards and the U^T
// elimination (rows above the pivot) is skipped; slots above the diagonal then keep U.
template <bool SYM>
__device__ __forceinline__ int sweep64(double (&x)[8][2], double* __restrict__ Dout, double (*rowb)[QN],
                                       double (*colb)[QN], double* __restrict__ rdiag) {
  const int tid = threadIdx.x;
  const int ti = tid >> 5, tj = tid & 31;
  int bad = 0;
  if (ti == 0) {
#pragma unroll
    for (int b = 0; b < 2; b++) rowb[0][tj + 32 * b] = x[0][b];
    if (tj == 0) rdiag[0] = fast_rcp(x[0][0]);
  }
  if (tj == 0) {
#pragma unroll
    for (int a = 0; a < 8; a++) colb[0][ti + 8 * a] = x[a][0];
  }
  __syncthreads();
#pragma unroll
  for (int ap = 0; ap < 8; ap++) {
#pragma unroll 1
    for (int jj = 0; jj < 8; jj++) {
      const int j = 8 * ap + jj;
      const int cur = j & 1, nxt = cur ^ 1;
      const double piv = rowb[cur][j];
      if (bad == 0 && !(fabs(piv) > 0.0 && isfinite(piv))) bad = j + 1;
      const double rp = rdiag[j];
      const int bj = ap >> 2;  // column block of column j (compile-time after unrolling)
      const bool own_col = tj == (j & 31);
      double v[2], vs[2];
#pragma unroll
      for (int b = 0; b < 2; b++) {
        v[b] = rowb[cur][tj + 32 * b];
        vs[b] = v[b] * rp;
      }
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int r = ti + 8 * a;
        const double w = colb[cur][r];
        if (a > ap || (a == ap && r > j)) {
          const double l = w * rp;
#pragma unroll
          for (int b = 0; b < 2; b++) {
            if (b == bj && own_col) {
              Dout[r * QLD + j] = l;
              x[a][b] = -l;
            } else {
              x[a][b] = fma(-l, v[b], x[a][b]);
            }
          }
        } else if (a < ap || (a == ap && r < j)) {
          if (!SYM) {
#pragma unroll
            for (int b = 0; b < 2; b++) {
              if (b < bj) continue;
              if (b > bj || tj + 32 * b > j) x[a][b] = fma(-w, vs[b], x[a][b]);
            }
          }
        } else {
#pragma unroll
          for (int b = 0; b < 2; b++) {
            if (b < bj) continue;
            const int c = tj + 32 * b;
            if (b > bj || c >= j) Dout[j * QLD + c] = x[a][b];
            if (!SYM && (b > bj || c > j)) x[a][b] = -vs[b];
          }
        }
      }
      const int jn = j + 1;
      if (jn < QN) {
        if (ti == (jn & 7)) {
          // publish register row jn (row block ap, or ap + 1 when jj == 7) and 1 / pivot
          const int bn = jn >> 5;
#pragma unroll
          for (int b = 0; b < 2; b++) {
            const double val = (jj < 7) ? x[ap][b] : x[ap + 1 < 8 ? ap + 1 : 7][b];
            rowb[nxt][tj + 32 * b] = val;
            if (b == bn && tj == (jn & 31)) rdiag[jn] = fast_rcp(val);
          }
        }
        if (tj == (jn & 31)) {
          const bool next_block = (jj == 7) && ((ap & 3) == 3);  // jn crosses into column block bj + 1
#pragma unroll
          for (int a = 0; a < 8; a++) colb[nxt][ti + 8 * a] = next_block ? x[a][bj + 1 < 2 ? bj + 1 : 1] : x[a][bj];
        }
      }
      __syncthreads();
    }
  }
  return bad;
}

// acc (warp tile 16 x 32, warps 4 x 2) = A[64][QLD] * B[64][QLD]
// (Skipping the 8 x 8 x 4 products that only add structural zeros of the triangular operands was measured:
//  no gain -- the phases around these products are bound by global-memory latency and store issue.)
__device__ __forceinline__ void gemm64(const double* __restrict__ As, const double* __restrict__ Bs,
                                       double (&acc)[2][4][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
  for (int s = 0; s < QN / 4; s++) {
    double a[2], b[4];
#pragma unroll
    for (int i = 0; i < 2; i++) a[i] = As[(wm * 16 + i * 8 + g) * QLD + s * 4 + t];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = Bs[(s * 4 + t) * QLD + wn * 32 + j * 8 + g];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// fragment (i, j) of this thread covers row frag_row(i), columns frag_col(j), frag_col(j) + 1
__device__ __forceinline__ int frag_row(int i) { return ((threadIdx.x >> 5) >> 1) * 16 + i * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int frag_col(int j) { return ((threadIdx.x >> 5) & 1) * 32 + j * 8 + 2 * (threadIdx.x & 3); }

__device__ __forceinline__ void load_quadrant(double* __restrict__ S, const double* __restrict__ G, int64_t ldg) {
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int idx = q * 256 + threadIdx.x;
    const int r = idx >> 5, c = (idx & 31) * 2;
    *reinterpret_cast<double2*>(&S[r * QLD + c]) = *reinterpret_cast<const double2*>(G + (int64_t)r * ldg + c);
  }
}
__device__ __forceinline__ void store_quadrant(double* __restrict__ G, int64_t ldg, const double* __restrict__ S) {
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const int idx = q * 256 + threadIdx.x;
    const int r = idx >> 5, c = (idx & 31) * 2;
    *reinterpret_cast<double2*>(G + (int64_t)r * ldg + c) = *reinterpret_cast<const double2*>(&S[r * QLD + c]);
  }
}
// sweep registers -> inv(L) and inv(U) of the quadrant, to shared buffers and to global
__device__ __forceinline__ void emit_inverses(const double (&x)[8][2], const double* __restrict__ rdiag,
                                              double* __restrict__ SL, double* __restrict__ SU,
                                              double* __restrict__ GL, double* __restrict__ GU) {
  const int ti = threadIdx.x >> 5, tj = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int r = ti + 8 * a, c = tj + 32 * b;
      const double v = x[a][b];
      const double il = c < r ? v : (c == r ? 1.0 : 0.0);
      const double iu = c > r ? v * rdiag[c] : (c == r ? rdiag[c] : 0.0);
      if (SL) SL[r * QLD + c] = il;
      if (SU) SU[r * QLD + c] = iu;
      GL[r * NB + c] = il;
      GU[r * NB + c] = iu;
    }
}

// symmetric block: inv(L) from the sweep registers, inv(U) = inv(L)^T D^-1 through the shared copy
__device__ __forceinline__ void emit_inverses_sym(const double (&x)[8][2], const double* __restrict__ rdiag,
                                                  double* __restrict__ SL, double* __restrict__ SU,
                                                  double* __restrict__ GL, double* __restrict__ GU) {
  const int ti = threadIdx.x >> 5, tj = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int r = ti + 8 * a, c = tj + 32 * b;
      const double il = c < r ? x[a][b] : (c == r ? 1.0 : 0.0);
      SL[r * QLD + c] = il;
      GL[r * NB + c] = il;
    }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int r = ti + 8 * a, c = tj + 32 * b;
      const double iu = c > r ? SL[c * QLD + r] * rdiag[c] : (c == r ? rdiag[c] : 0.0);
      if (SU) SU[r * QLD + c] = iu;
      GU[r * NB + c] = iu;
    }
}

template <bool SYM>
__global__ void __launch_bounds__(256, 2)
diag_kernel_small(double* __restrict__ M, int64_t ld, int64_t o, double* __restrict__ invL,
                  double* __restrict__ invU, int32_t* __restrict__ info, int block_index) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* S0 = reinterpret_cast<double*>(smem_raw);
  double* S1 = S0 + QN * QLD;
  double* S2 = S1 + QN * QLD;
  __shared__ double rowb[2][QN], colb[2][QN], rdiag[QN], dvec[QN];
  const int tid = threadIdx.x, ti = tid >> 5, tj = tid & 31;
  double* A11 = M + o * ld + o;
  double* A12 = A11 + QN;
  double* A21 = A11 + (int64_t)QN * ld;
  double* A22 = A21 + QN;
  double x[8][2];
  double acc[2][4][2];

  // ---- phase 1: A11 ----
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) x[a][b] = A11[(int64_t)(ti + 8 * a) * ld + tj + 32 * b];
  SCB_STAMP(0);
  int bad = sweep64<SYM>(x, S0, rowb, colb, rdiag);
  SCB_STAMP(1);
  if (SYM) {
    emit_inverses_sym(x, rdiag, S1, S2, invL, invU);  // S1 = inv(L11), S2 = inv(U11)
    if (tid < QN) dvec[tid] = S0[tid * QLD + tid];    // diagonal of U11
  } else {
    emit_inverses(x, rdiag, S1, S2, invL, invU);
  }
  store_quadrant(A11, ld, S0);                   // L11 \ U11 in place
  __syncthreads();
  SCB_STAMP(2);
  if (!SYM) {
    // ---- phase 2a: U12 = inv(L11) A12 ----
    load_quadrant(S0, A12, ld);
    __syncthreads();
    gemm64(S1, S0, acc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = frag_row(i), c = frag_col(j);
        const double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
        *reinterpret_cast<double2*>(&S0[r * QLD + c]) = v;            // S0 = U12
        *reinterpret_cast<double2*>(A12 + (int64_t)r * ld + c) = v;
      }
  }
  // ---- phase 2b: L21 = A21 inv(U11)   (symmetric: U12 = diag(U11) L21^T) ----
  load_quadrant(S1, A21, ld);
  __syncthreads();
  gemm64(S1, S2, acc);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = frag_row(i), c = frag_col(j);
      const double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<double2*>(&S1[r * QLD + c]) = v;            // S1 = L21
      *reinterpret_cast<double2*>(A21 + (int64_t)r * ld + c) = v;
      if (SYM) {
        const double u0 = dvec[c] * v.x, u1 = dvec[c + 1] * v.y;
        S0[c * QLD + r] = u0;                                       // S0 = U12
        S0[(c + 1) * QLD + r] = u1;
        A12[(int64_t)c * ld + r] = u0;
        A12[(int64_t)(c + 1) * ld + r] = u1;
      }
    }
  __syncthreads();
  SCB_STAMP(3);
  // ---- phase 3: A22 -= L21 U12  -> S2 ----
  gemm64(S1, S0, acc);
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = frag_row(i), c = frag_col(j);
      const double2 cin = *reinterpret_cast<const double2*>(A22 + (int64_t)r * ld + c);
      *reinterpret_cast<double2*>(&S2[r * QLD + c]) = make_double2(cin.x - acc[i][j][0], cin.y - acc[i][j][1]);
    }
  __syncthreads();
  // ---- phase 4: A22 ----
#pragma unroll
  for (int a = 0; a < 8; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) x[a][b] = S2[(ti + 8 * a) * QLD + tj + 32 * b];
  __syncthreads();
  SCB_STAMP(4);
  const int bad2 = sweep64<SYM>(x, S2, rowb, colb, rdiag);
  SCB_STAMP(5);
  if (bad == 0 && bad2) bad = QN + bad2;
  if (tid == 0 && bad) atomicCAS(info, 0, block_index * NB + bad);
  store_quadrant(A22, ld, S2);                   // L22 \ U22 in place
  __syncthreads();
  double* iL22 = invL + QN * NB + QN;
  double* iU22 = invU + QN * NB + QN;
  if (SYM)
    emit_inverses_sym(x, rdiag, S2, nullptr, iL22, iU22);  // S2 = inv(L22)
  else
    emit_inverses(x, rdiag, S2, nullptr, iL22, iU22);
  __syncthreads();
  SCB_STAMP(6);
  // ---- phase 5a: inv(L)21 = -(inv(L22) L21) inv(L11) ----
  gemm64(S2, S1, acc);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      *reinterpret_cast<double2*>(&S1[frag_row(i) * QLD + frag_col(j)]) = make_double2(acc[i][j][0], acc[i][j][1]);
  load_quadrant(S2, invL, NB);                   // inv(L11) (written in phase 1 by this CTA)
  __syncthreads();
  gemm64(S1, S2, acc);
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = frag_row(i), c = frag_col(j);
      *reinterpret_cast<double2*>(invL + (int64_t)(QN + r) * NB + c) = make_double2(-acc[i][j][0], -acc[i][j][1]);
      *reinterpret_cast<double2*>(invL + (int64_t)r * NB + QN + c) = make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(invU + (int64_t)(QN + r) * NB + c) = make_double2(0.0, 0.0);
      if (SYM) {  // inv(U)12 = (inv(L)21)^T D2^-1   (rdiag now holds 1 / diag(U22))
        invU[(int64_t)c * NB + QN + r] = -acc[i][j][0] * rdiag[r];
        invU[(int64_t)(c + 1) * NB + QN + r] = -acc[i][j][1] * rdiag[r];
      }
    }
  if (!SYM) {
    __syncthreads();
    // ---- phase 5b: inv(U)12 = -inv(U11) (U12 inv(U22)) ----
    load_quadrant(S2, iU22, NB);
    __syncthreads();
    gemm64(S0, S2, acc);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        *reinterpret_cast<double2*>(&S0[frag_row(i) * QLD + frag_col(j)]) = make_double2(acc[i][j][0], acc[i][j][1]);
    load_quadrant(S2, invU, NB);                 // inv(U11)
    __syncthreads();
    gemm64(S2, S0, acc);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int r = frag_row(i), c = frag_col(j);
        *reinterpret_cast<double2*>(invU + (int64_t)r * NB + QN + c) = make_double2(-acc[i][j][0], -acc[i][j][1]);
      }
  }
  SCB_STAMP(7);
}

// ---------------------------------------------------------------------------------------
// 1c. diagonal block of the SYMMETRIC factorization, blocked (the default in symmetric mode).
//     Same outputs as diag_kernel_small<true> (L \ U in place with U = D L^T, inv(L), inv(U) =
//     inv(L)^T D^-1), same 2x2 quadrant organisation and footprint, but each 64x64 quadrant is
//     factored by a right-looking LDL^T with 8-wide panels instead of 64 rank-1 sweeps:
//       (A) warp 0: LDL^T of the 8x8 diagonal tile in registers (one row per lane, shuffles);
//       (B) 64 threads: panel rows below by forward substitution, written as L and, scaled by
//           D, mirrored into the upper triangle as U; one more warp: inverse of the 8x8 tile;
//       (C) all warps: trailing lower-triangle 8x8 tiles, C -= L U, one DMMA pair per tile;
//     and inv(L) is assembled from the 8x8 tile inverses by three doubling levels of small DMMA
//     products (X21 = -X22 L21 X11).  Only 128 pivots remain sequential.
// ---------------------------------------------------------------------------------------
constexpr int TLD = 36;  // row stride of the 32 x 32 scratch used by the doubling products

// (i, j), j <= i, of the e-th tile of a lower-triangular tile grid (row-major enumeration)
__device__ __forceinline__ void tri_tile(int e, int& i, int& j) {
  // e < 28: rows start at 0, 1, 3, 6, 10, 15, 21
  i = (e >= 1) + (e >= 3) + (e >= 6) + (e >= 10) + (e >= 15) + (e >= 21);
  j = e - i * (i + 1) / 2;
}

// S: 64x64 symmetric block (lower triangle + diagonal are read).  On return S = L \ U,
// SL = inv(L) (dense 64x64, unit lower), dd = diag(U), rd = 1 / dd.  tmp: 32 x TLD scratch.
__device__ __forceinline__ int ldl64_blocked(double* __restrict__ S, double* __restrict__ SL,
                                             double* __restrict__ tmp, double* __restrict__ dd,
                                             double* __restrict__ rd, int* __restrict__ bad_s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int idx = tid; idx < QN * QLD; idx += 256) SL[idx] = 0.0;
  if (tid == 0) *bad_s = 0x7fffffff;
  int bad = 0;  // first bad pivot seen by this lane (warp 0 only)
  __syncthreads();
#pragma unroll 1
  for (int kb = 0; kb < 8; kb++) {
    const int k0 = 8 * kb;
    if (kb == 0) SCB_STAMP(8);
    // ---- (A) 8x8 diagonal tile: every lane of warp 0 holds the whole lower triangle in registers
    //      and runs the same straight-line LDL^T (no shuffles).  Every lane also STORES every result --
    //      the same value to the same address, one wavefront per store: lane-predicated stores made the
    //      compiler emit dozens of divergent blocks per tile, which cost more than the arithmetic
    //      (tools/phase_a_bench.cu: 1950 -> 1185 cycles per tile, identical results)
    if (warp == 0) {
      double a[8][8];  // a[r][c], c <= r
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) a[r][c] = S[(k0 + r) * QLD + k0 + c];
      __syncwarp();  // every lane has read the tile before entries of it are overwritten
#pragma unroll
      for (int j = 0; j < 8; j++) {
        // The 8 pivots are a chain of dependent fp64 operations: reciprocal (seed + 3 dependent FMAs,
        // 41 cycles), the multipliers l_r = a_rj / d_j, then every trailing entry is ONE independent FMA
        // a_rc -= l_r a_cj.
        const double djj = a[j][j];
        const double rjj = fast_rcp_halley(djj);
        double l[8];
#pragma unroll
        for (int r = j + 1; r < 8; r++) l[r] = a[r][j] * rjj;
#pragma unroll
        for (int r = j + 1; r < 8; r++)
#pragma unroll
          for (int c = j + 1; c <= r; c++) a[r][c] = fma(-l[r], a[c][j], a[r][c]);
        // column j is final (stores are off the dependency chain)
#pragma unroll
        for (int r = j + 1; r < 8; r++) {
          S[(k0 + r) * QLD + k0 + j] = l[r];     // L[r][j]
          S[(k0 + j) * QLD + k0 + r] = a[r][j];  // U[j][r] = d_j L[r][j]
        }
        S[(k0 + j) * QLD + k0 + j] = djj;
        dd[k0 + j] = djj;
        rd[k0 + j] = rjj;
        bad = (bad == 0 && !(fabs(djj) > 0.0 && isfinite(djj))) ? k0 + j + 1 : bad;
      }
    }
    if (kb == 0) SCB_STAMP(9);
    __syncthreads();
    if (kb == 0) SCB_STAMP(10);
    // ---- (B) panel rows below the tile (threads 0..63) and the tile inverse (warp 7) ----
    {
      const int r = k0 + 8 + tid;
      if (r < QN) {
        double u[8];
#pragma unroll
        for (int j = 0; j < 8; j++) u[j] = S[r * QLD + k0 + j];
#pragma unroll
        for (int j = 0; j < 8; j++) {
#pragma unroll
          for (int pp = 0; pp < j; pp++) u[j] = fma(-u[pp], S[(k0 + j) * QLD + k0 + pp], u[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          S[r * QLD + k0 + j] = u[j] * rd[k0 + j];  // L[r][k0 + j]
          S[(k0 + j) * QLD + r] = u[j];             // U[k0 + j][r] = d_j L[r][k0 + j]
        }
      }
      if (warp == 7 && lane < 8) {
        // column `lane` of the inverse of the unit lower 8x8 tile
        double xv[8];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
          double acc = (rr == lane) ? 1.0 : 0.0;
#pragma unroll
          for (int pp = 0; pp < rr; pp++)
            if (pp >= lane) acc = fma(-S[(k0 + rr) * QLD + k0 + pp], xv[pp], acc);
          xv[rr] = (rr >= lane) ? acc : 0.0;
          SL[(k0 + rr) * QLD + k0 + lane] = xv[rr];
        }
      }
    }
    if (kb == 0) SCB_STAMP(11);
    __syncthreads();
    if (kb == 0) SCB_STAMP(12);
    // ---- (C) trailing lower-triangle tiles: C -= L_panel U_panel, two tiles per warp in flight ----
    // (Measured alternatives that were slower: four tiles in flight (spills), and letting warp 0 run ahead
    //  into the next tile factorization while the other warps finish this phase.)
    const int nt = 7 - kb;
    const int ntiles = nt * (nt + 1) / 2;
#pragma unroll 1
    for (int e0 = warp; e0 < ntiles; e0 += 16) {
      const bool two = e0 + 8 < ntiles;
      int ti, tj, ui = 0, uj = 0;
      tri_tile(e0, ti, tj);
      if (two) tri_tile(e0 + 8, ui, uj);
      const int R0 = k0 + 8 + 8 * ti, C0 = k0 + 8 + 8 * tj;
      const int R1 = k0 + 8 + 8 * ui, C1 = k0 + 8 + 8 * uj;
      double2 c0 = *reinterpret_cast<const double2*>(&S[(R0 + g) * QLD + C0 + 2 * t]);
      double2 c1 = make_double2(0.0, 0.0);
      if (two) c1 = *reinterpret_cast<const double2*>(&S[(R1 + g) * QLD + C1 + 2 * t]);
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        const double av0 = -S[(R0 + g) * QLD + k0 + ks * 4 + t];
        const double bv0 = S[(k0 + ks * 4 + t) * QLD + C0 + g];
        dmma(c0.x, c0.y, av0, bv0);
        if (two) {
          const double av1 = -S[(R1 + g) * QLD + k0 + ks * 4 + t];
          const double bv1 = S[(k0 + ks * 4 + t) * QLD + C1 + g];
          dmma(c1.x, c1.y, av1, bv1);
        }
      }
      *reinterpret_cast<double2*>(&S[(R0 + g) * QLD + C0 + 2 * t]) = c0;
      if (two) *reinterpret_cast<double2*>(&S[(R1 + g) * QLD + C1 + 2 * t]) = c1;
    }
    if (kb == 0) SCB_STAMP(13);
    __syncthreads();
    if (kb == 0) SCB_STAMP(14);
  }
  SCB_STAMP(15);
  if (bad) atomicMin(bad_s, bad);  // (rare) the smallest bad index wins
  // ---- inv(L) by doubling: X21 = -X22 (L21 X11) for block sizes 8, 16, 32 ----
#pragma unroll 1
  for (int sz = 8; sz < QN; sz *= 2) {
    const int tps = sz / 8;               // tiles per side of one off-diagonal block
    const int tpp = tps * tps;            // tiles per pair
    const int total = (QN / (2 * sz)) * tpp;
    for (int e = warp; e < total; e += 8) {  // T = L21 X11
      const int pair = e / tpp, q = e % tpp, i = q / tps, j = q % tps;
      const int a0 = pair * 2 * sz, b0 = a0 + sz;
      double c0 = 0.0, c1 = 0.0;
      for (int k = 0; k < sz; k += 4) {
        const double av = S[(b0 + 8 * i + g) * QLD + a0 + k + t];
        const double bv = SL[(a0 + k + t) * QLD + a0 + 8 * j + g];
        dmma(c0, c1, av, bv);
      }
      *reinterpret_cast<double2*>(&tmp[(pair * sz + 8 * i + g) * TLD + 8 * j + 2 * t]) = make_double2(c0, c1);
    }
    __syncthreads();
    for (int e = warp; e < total; e += 8) {  // X21 = -X22 T
      const int pair = e / tpp, q = e % tpp, i = q / tps, j = q % tps;
      const int a0 = pair * 2 * sz, b0 = a0 + sz;
      double c0 = 0.0, c1 = 0.0;
      for (int k = 0; k < sz; k += 4) {
        const double av = -SL[(b0 + 8 * i + g) * QLD + b0 + k + t];
        const double bv = tmp[(pair * sz + k + t) * TLD + 8 * j + g];
        dmma(c0, c1, av, bv);
      }
      *reinterpret_cast<double2*>(&SL[(b0 + 8 * i + g) * QLD + a0 + 8 * j + 2 * t]) = make_double2(c0, c1);
    }
    __syncthreads();
  }
  return *bad_s == 0x7fffffff ? 0 : *bad_s;
}

// inv(L) (shared, dense) -> global inv(L) and inv(U) = inv(L)^T D^-1 (global, optionally shared)
__device__ __forceinline__ void emit_inverses_symb(const double* __restrict__ SL, const double* __restrict__ rd,
                                                   double* __restrict__ SU, double* __restrict__ GL,
                                                   double* __restrict__ GU) {
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const int idx = q * 256 + threadIdx.x;
    const int r = idx >> 6, c = idx & 63;
    GL[r * NB + c] = SL[r * QLD + c];
    const double iu = c > r ? SL[c * QLD + r] * rd[c] : (c == r ? rd[c] : 0.0);
    if (SU) SU[r * QLD + c] = iu;
    GU[r * NB + c] = iu;
  }
}

__global__ void __launch_bounds__(256, 2)
diag_kernel_symb(double* __restrict__ M, int64_t ld, int64_t o, double* __restrict__ invL,
                 double* __restrict__ invU, int32_t* __restrict__ info, int block_index) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* S0 = reinterpret_cast<double*>(smem_raw);
  double* S1 = S0 + QN * QLD;
  double* S2 = S1 + QN * QLD;
  double* tmp = S2 + QN * QLD;  // 32 x TLD
  __shared__ double dd[QN], rd[QN];
  __shared__ int bad_s;
  const int tid = threadIdx.x;
  double* A11 = M + o * ld + o;
  double* A12 = A11 + QN;
  double* A21 = A11 + (int64_t)QN * ld;
  double* A22 = A21 + QN;
  double acc[2][4][2];

  // ---- phase 1: A11 = L11 D1 L11^T ----
  load_quadrant(S0, A11, ld);
  __syncthreads();
  SCB_STAMP(0);
  int bad = ldl64_blocked(S0, S1, tmp, dd, rd, &bad_s);  // S0 = L11 \ U11, S1 = inv(L11)
  SCB_STAMP(1);
  emit_inverses_symb(S1, rd, S2, invL, invU);            // S2 = inv(U11)
  store_quadrant(A11, ld, S0);
  __syncthreads();
  SCB_STAMP(2);
  // ---- phase 2: L21 = A21 inv(U11), U12 = D1 L21^T ----
  load_quadrant(S1, A21, ld);
  __syncthreads();
  gemm64(S1, S2, acc);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = frag_row(i), c = frag_col(j);
      const double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<double2*>(&S1[r * QLD + c]) = v;  // S1 = L21
      *reinterpret_cast<double2*>(A21 + (int64_t)r * ld + c) = v;
      const double u0 = dd[c] * v.x, u1 = dd[c + 1] * v.y;
      S0[c * QLD + r] = u0;  // S0 = U12
      S0[(c + 1) * QLD + r] = u1;
      A12[(int64_t)c * ld + r] = u0;
      A12[(int64_t)(c + 1) * ld + r] = u1;
    }
  __syncthreads();
  SCB_STAMP(3);
  // ---- phase 3: A22 -= L21 U12 -> S2 ----
  gemm64(S1, S0, acc);
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = frag_row(i), c = frag_col(j);
      const double2 cin = *reinterpret_cast<const double2*>(A22 + (int64_t)r * ld + c);
      *reinterpret_cast<double2*>(&S2[r * QLD + c]) = make_double2(cin.x - acc[i][j][0], cin.y - acc[i][j][1]);
    }
  __syncthreads();
  SCB_STAMP(4);
  // ---- phase 4: A22 = L22 D2 L22^T ----
  const int bad2 = ldl64_blocked(S2, S0, tmp, dd, rd, &bad_s);  // S2 = L22 \ U22, S0 = inv(L22); dd, rd: block 2
  SCB_STAMP(5);
  if (bad == 0 && bad2) bad = QN + bad2;
  if (tid == 0 && bad) atomicCAS(info, 0, block_index * NB + bad);
  store_quadrant(A22, ld, S2);
  emit_inverses_symb(S0, rd, nullptr, invL + QN * NB + QN, invU + QN * NB + QN);
  __syncthreads();
  SCB_STAMP(6);
  // ---- phase 5: inv(L)21 = -(inv(L22) L21) inv(L11), inv(U)12 = (inv(L)21)^T D2^-1 ----
  gemm64(S0, S1, acc);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      *reinterpret_cast<double2*>(&S1[frag_row(i) * QLD + frag_col(j)]) = make_double2(acc[i][j][0], acc[i][j][1]);
  load_quadrant(S2, invL, NB);  // inv(L11) (written in phase 1 by this CTA)
  __syncthreads();
  gemm64(S1, S2, acc);
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int r = frag_row(i), c = frag_col(j);
      *reinterpret_cast<double2*>(invL + (int64_t)(QN + r) * NB + c) = make_double2(-acc[i][j][0], -acc[i][j][1]);
      *reinterpret_cast<double2*>(invL + (int64_t)r * NB + QN + c) = make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(invU + (int64_t)(QN + r) * NB + c) = make_double2(0.0, 0.0);
      invU[(int64_t)c * NB + QN + r] = -acc[i][j][0] * rd[r];
      invU[(int64_t)(c + 1) * NB + QN + r] = -acc[i][j][1] * rd[r];
    }
  SCB_STAMP(7);
}

// ---------------------------------------------------------------------------------------
// Partial (row) pivoting -- the fallback for systems that are not provably safe without it
// (reference: scipy.linalg.lu_factor = LAPACK dgetrf at solver/solve_film.py:232,253,279 always
// pivots).  The pivot rows of a 128-column panel are found on a scratch copy of the panel
// (rows o .. n, ping-pong buffers X -> Y, one launch per column so that every column's pivot is the
// maximum over ALL rows below the diagonal, exactly LAPACK's choice up to ties); the interchanges
// are then applied to the full rows of M and the panel is factored by the same unpivoted kernels
// as above -- on the permuted rows that is the identical elimination.
// ---------------------------------------------------------------------------------------
constexpr int PR = 64;  // panel rows per CTA of the pivot search

struct PivCand {
  double a;   // |value|
  int32_t r;  // row relative to the panel origin
};

// larger magnitude wins; ties (and NaN-free equality) go to the smaller row: deterministic
__device__ __forceinline__ bool piv_better(double a, int32_t r, double b, int32_t q) {
  return a > b || (a == b && r < q);
}

// CTA-wide argmax of (a, r) over 256 threads; result valid in thread 0
__device__ __forceinline__ void piv_block_argmax(double& a, int32_t& r, double* sa, int32_t* sr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double b = __shfl_xor_sync(0xffffffffu, a, off);
    const int32_t q = __shfl_xor_sync(0xffffffffu, r, off);
    if (piv_better(b, q, a, r)) { a = b; r = q; }
  }
  if (lane == 0) { sa[warp] = a; sr[warp] = r; }
  __syncthreads();
  if (tid == 0) {
    for (int w8 = 1; w8 < 8; w8++)
      if (piv_better(sa[w8], sr[w8], a, r)) { a = sa[w8]; r = sr[w8]; }
  }
}

// step -1: Y <- panel of M (rows o .., columns o .. o+127), candidates of column 0
__global__ void __launch_bounds__(256)
piv_init_kernel(const double* __restrict__ M, int64_t ld, int64_t o, int64_t nrows, double* __restrict__ Y,
                PivCand* __restrict__ cand_out) {
  __shared__ double sa[8];
  __shared__ int32_t sr[8];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * PR;
  double best = -1.0;
  int32_t brow = 0x7fffffff;
  for (int idx = tid; idx < PR * 64; idx += 256) {  // 64 double2 per row
    const int64_t r = r0 + (idx >> 6);
    const int c = (idx & 63) * 2;
    if (r < nrows) {
      const double2 v = *reinterpret_cast<const double2*>(M + (o + r) * ld + o + c);
      *reinterpret_cast<double2*>(Y + r * NB + c) = v;
      if (c == 0 && piv_better(fabs(v.x), (int32_t)r, best, brow)) { best = fabs(v.x); brow = (int32_t)r; }
    }
  }
  piv_block_argmax(best, brow, sa, sr);
  if (tid == 0) cand_out[blockIdx.x] = PivCand{best, brow};
}

// step j (0 .. 127): pivot of column j from the candidates of the previous launch; rows j+1 .. of
// X with rows (j, p) interchanged are eliminated into Y (columns j+1 ..), candidates of column j+1
__global__ void __launch_bounds__(256)
piv_step_kernel(const double* __restrict__ X, double* __restrict__ Y, int64_t nrows, int j, int ncand,
                const PivCand* __restrict__ cand_in, PivCand* __restrict__ cand_out, int64_t o,
                int32_t* __restrict__ piv, int32_t* __restrict__ info) {
  __shared__ double sa[8];
  __shared__ int32_t sr[8];
  __shared__ double prow[NB];
  __shared__ int32_t s_p;
  __shared__ double s_rp;
  const int tid = threadIdx.x;
  // ---- pivot of column j: reduce the per-CTA candidates (every CTA does it redundantly) ----
  double a = -1.0;
  int32_t p = 0x7fffffff;
  for (int k = tid; k < ncand; k += 256) {
    const PivCand c = cand_in[k];
    if (piv_better(c.a, c.r, a, p)) { a = c.a; p = c.r; }
  }
  piv_block_argmax(a, p, sa, sr);
  if (tid == 0) {
    s_p = p;
    const bool bad = !(a > 0.0 && isfinite(a));
    if (blockIdx.x == 0) {
      piv[o + j] = (int32_t)(o + (bad ? j : p));
      if (bad) atomicCAS(info, 0, (int32_t)(o + j + 1));
    }
    if (bad) s_p = j;  // singular column: no interchange, the diagonal kernel reports it as well
  }
  __syncthreads();
  p = s_p;
  if (j == NB - 1) return;
  // pivot row (row p of X) -> shared, and 1 / pivot
  if (tid < NB) prow[tid] = X[(int64_t)p * NB + tid];
  __syncthreads();
  if (tid == 0) s_rp = 1.0 / prow[j];
  __syncthreads();
  const double rp = s_rp;
  // ---- eliminate my rows: thread (tr, tc) -> rows tr + 8 k, columns 4 tc .. 4 tc + 3 ----
  const int tr = tid >> 5, tc = tid & 31;
  const int c0 = 4 * tc;
  const int64_t r0 = (int64_t)blockIdx.x * PR;
  double best = -1.0;
  int32_t brow = 0x7fffffff;
  if (c0 + 3 > j) {  // this thread owns at least one column right of j
    for (int k = 0; k < PR / 8; k++) {
      const int64_t r = r0 + tr + 8 * k;
      if (r <= j || r >= nrows) continue;
      const int64_t src = (r == p) ? j : r;  // after the interchange position p holds old row j
      const double l = X[src * NB + j] * rp;
      const double4 x = *reinterpret_cast<const double4*>(X + src * NB + c0);
      double v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int c = c0 + e;
        if (c > j) {
          v[e] = fma(-l, prow[c], v[e]);
          if (c == j + 1 && piv_better(fabs(v[e]), (int32_t)r, best, brow)) { best = fabs(v[e]); brow = (int32_t)r; }
        }
      }
      *reinterpret_cast<double4*>(Y + r * NB + c0) = make_double4(v[0], v[1], v[2], v[3]);
    }
  }
  piv_block_argmax(best, brow, sa, sr);
  if (tid == 0) cand_out[blockIdx.x] = PivCand{best, brow};
}

// perm <- identity
__global__ void perm_init_kernel(int64_t n, int32_t* __restrict__ perm) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) perm[i] = (int32_t)i;
}

// applies the 128 interchanges of panel o to the full rows of M (one thread per column, the
// interchanges in order) and to the running permutation
__global__ void __launch_bounds__(256)
laswp_kernel(double* __restrict__ M, int64_t ld, int64_t o, const int32_t* __restrict__ piv,
             int32_t* __restrict__ perm) {
  __shared__ int32_t sp[NB];
  const int tid = threadIdx.x;
  if (tid < NB) sp[tid] = piv[o + tid];
  __syncthreads();
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + tid;
  if (c < ld) {
    for (int j = 0; j < NB; j++) {
      const int64_t p = sp[j];
      if (p != o + j) {
        const double a = M[(o + j) * ld + c], b = M[p * ld + c];
        M[(o + j) * ld + c] = b;
        M[p * ld + c] = a;
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    for (int j = 0; j < NB; j++) {
      const int64_t p = sp[j];
      if (p != o + j) {
        const int32_t t = perm[o + j];
        perm[o + j] = perm[p];
        perm[p] = t;
      }
    }
  }
}

static bool g_attr_set[64] = {};  // kernel attributes are per device
static int g_diag_small = 1;
static int g_diag_symb = 1;  // blocked LDL^T diagonal kernel in symmetric mode (SCB_DIAG_SYMB=0: sweep version)
static int g_lookahead = 1;
static int g_recursive_strips = 1;  // binary-tree schedule of the inner strip updates (SCB_LU_RECURSIVE=0: eager)
static int g_split_panel = 1;  // symmetric LU: square of the outer panel on the chain, rows below on a 2nd stream (SCB_LU_SPLIT=0: off)
static int g_lazy_strips = 0;  // left-looking inner strips: same flops, measured no faster (narrow grids)
static int g_inner_la = 0;     // split panels: block-level look-ahead inside the square (SCB_LU_INNER_LA=1; measured: no gain)
static int g_tail_q = 0;       // outer-panel width (in 128-blocks) used for the last g_tail_blocks blocks (0: same q)
static int g_tail_blocks = 0;
static int g_band = 16;        // raster order of the update kernel: row tiles per band (SCB_LU_BAND=0: hardware order)
static int g_lat = 3;          // latency variants of the GEMM kernels on the per-block chain: bit 0 in-square panel solve + K = 128 update, bit 1 the K = 1024 update of the next panel's square, bit 2 (off: measured slower, 5.3k film 3.57 -> 3.69 ms, 20k 80.1 -> 80.5 ms) that update split so that its first block column releases the chain (SCB_LU_LAT=0: all off)

}  // namespace scb

using namespace scb;

static int lu_outer_blocks() {
  static int q = -1;
  if (q < 0) {
    q = 8;
    if (const char* e = getenv("SCB_LU_Q")) q = atoi(e);
    if (q < 1) q = 1;
    if (q > 8) q = 8;
  }
  return q;
}

// helper stream (highest priority) and event pool for the look-ahead, one set per (device, caller
// stream): factorizations issued on different streams (independent films) overlap on the GPU
struct LuStreams {
  cudaStream_t panel = nullptr;
  cudaStream_t rest = nullptr;  // split panels: rows below the outer panel's diagonal square
  cudaStream_t square = nullptr;  // split panels: the columns of the square that are not next on the chain
  std::vector<cudaEvent_t> events;
  cudaEvent_t event(size_t i) {
    while (events.size() <= i) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      events.push_back(e);
    }
    return events[i];
  }
};
static std::map<std::pair<int, cudaStream_t>, LuStreams> g_lu_streams;
static std::mutex g_lu_streams_mutex;

// offset (in doubles) of the small integer scratch used by scb_getrs_nopiv (ready flags + work counter)
namespace scb {
int64_t lu_flags_offset(int64_t n_pad) {
  const int64_t nb = n_pad / NB;
  return nb * 2 * NB * NB + 4 * n_pad * NB * lu_outer_blocks();
}
}  // namespace scb

extern "C" int64_t scb_getrf_dinv_bytes(int64_t n_pad) {
  const int64_t nb = n_pad / NB;
  // [nb][2][128][128] block inverses + 2 x (Lpack [n_pad x KB] + Upack [KB x n_pad]), KB = q * 128
  // (two pack sets: the next outer panel is factored while the previous one is still being applied)
  // + getrs scratch (nb ready flags + counters)
  return (lu_flags_offset(n_pad) + nb + 16) * (int64_t)sizeof(double);
}

// Two-level right-looking LU with look-ahead.
//  * Inner panels of NB = 128 columns are factored and applied only to the L-shaped strip of the
//    current outer panel (KB = q*128 columns); the big trailing block is updated once per outer
//    panel with K = KB, which amortises the C-tile traffic and the CTA prologue/epilogue of the
//    DMMA update kernel over q times more math.
//  * Look-ahead: the trailing update of outer panel P is split into the L-shaped strip that the
//    next panel needs (done first) and the rest; the factorization of panel P+1 (a latency-bound
//    chain of small kernels) then runs on a high-priority stream concurrently with "the rest".
//    Every kernel of that chain fits into the SM slot of an update CTA, so the hardware block
//    scheduler interleaves them as slots free up.
static int getrf_impl(int64_t n_pad, double* M, double* dinv, int32_t* info, scb_stream_t stream, bool sym) {
  SCB_CHECK_ARG(n_pad > 0 && n_pad % NB == 0, "n_pad must be a positive multiple of 128");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nb = n_pad / NB;
  const int q = lu_outer_blocks();
  const int tile_chunks = q * NCHUNK;
  double* pack_base = dinv + nb * 2 * NB * NB;
  const int64_t pack_set = 2 * n_pad * NB * q;  // doubles per (Lpack + Upack) set
  const int diag_smem = NB * DLD * sizeof(double);
  const int diag_small_smem = 3 * QN * QLD * sizeof(double);
  const int diag_symb_smem = (3 * QN * QLD + 32 * TLD) * sizeof(double);
  const int trsm_smem = kTrsmSmemDoubles * sizeof(double);
  const int upd_smem = 2 * sizeof(UpdateStage);
  int dev = 0;
  SCB_CUDA(cudaGetDevice(&dev));
  if (!g_attr_set[dev & 63]) {
    SCB_CUDA(cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
    SCB_CUDA(cudaFuncSetAttribute(diag_kernel_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_small_smem));
    SCB_CUDA(cudaFuncSetAttribute(diag_kernel_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_small_smem));
    SCB_CUDA(cudaFuncSetAttribute(diag_kernel_symb, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_symb_smem));
    SCB_CUDA(cudaFuncSetAttribute(update_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem));
    SCB_CUDA(cudaFuncSetAttribute(update_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem));
    SCB_CUDA(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem));
    SCB_CUDA(cudaFuncSetAttribute(trsm_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem));
    SCB_CUDA(cudaFuncSetAttribute(update_kernel_t<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    SCB_CUDA(cudaFuncSetAttribute(update_kernel_t<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (const char* e = getenv("SCB_DIAG_SMALL")) g_diag_small = atoi(e);
    if (const char* e = getenv("SCB_DIAG_SYMB")) g_diag_symb = atoi(e);
    if (const char* e = getenv("SCB_LU_LOOKAHEAD")) g_lookahead = atoi(e);
    if (const char* e = getenv("SCB_LU_LAZY")) g_lazy_strips = atoi(e);
    if (const char* e = getenv("SCB_LU_RECURSIVE")) g_recursive_strips = atoi(e);
    if (const char* e = getenv("SCB_LU_SPLIT")) g_split_panel = atoi(e);
    if (const char* e = getenv("SCB_LU_INNER_LA")) g_inner_la = atoi(e);
    if (const char* e = getenv("SCB_LU_TAIL_Q")) g_tail_q = atoi(e);
    if (const char* e = getenv("SCB_LU_TAIL_BLOCKS")) g_tail_blocks = atoi(e);
    if (const char* e = getenv("SCB_LU_BAND")) g_band = atoi(e);
    if (const char* e = getenv("SCB_LU_LAT")) g_lat = atoi(e);
    SCB_CUDA(cudaFuncSetAttribute(update_lat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdLatSmem));
    SCB_CUDA(cudaFuncSetAttribute(update_lat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdLatSmem));
    SCB_CUDA(cudaFuncSetAttribute(trsm_sym_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmLatSmem));
    g_attr_set[dev & 63] = true;
  }
  LuStreams* lsp;
  {
    std::lock_guard<std::mutex> lock(g_lu_streams_mutex);
    lsp = &g_lu_streams[std::make_pair(dev, s)];  // (std::map nodes are address-stable)
  }
  LuStreams& ls = *lsp;
  const bool lookahead = g_lookahead && g_diag_small;
  cudaStream_t sp = s;
  if (lookahead) {
    if (!ls.panel) {
      int lo = 0, hi = 0;
      SCB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      SCB_CUDA(cudaStreamCreateWithPriority(&ls.panel, cudaStreamNonBlocking, hi));
    }
    sp = ls.panel;
  }
  const bool split = sym && lookahead && g_split_panel && g_recursive_strips && !g_lazy_strips;
  cudaStream_t sr = sp;
  if (split) {
    if (!ls.rest) {
      int lo = 0, hi = 0;
      SCB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      SCB_CUDA(cudaStreamCreateWithPriority(&ls.rest, cudaStreamNonBlocking, hi));
    }
    sr = ls.rest;
  }
  const bool inner_la = split && g_inner_la;
  cudaStream_t sq = sp;
  if (inner_la) {
    if (!ls.square) {
      int lo = 0, hi = 0;
      SCB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      SCB_CUDA(cudaStreamCreateWithPriority(&ls.square, cudaStreamNonBlocking, hi));
    }
    sq = ls.square;
  }
  size_t ev = 0;
  SCB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
  // ready flags / counters of scb_getrs_nopiv live behind the packed panels: start from zero
  SCB_CUDA(cudaMemsetAsync(dinv + lu_flags_offset(n_pad), 0, (nb + 16) * sizeof(double), s));

  // Outer panels: q blocks each; optionally narrower panels (SCB_LU_TAIL_Q) for the last
  // SCB_LU_TAIL_BLOCKS blocks, where the trailing block is too small to hide a q-block chain.
  std::vector<int64_t> pk;  // first 128-block of every outer panel
  std::vector<int> pq;      // blocks in it
  {
    const int qt = (g_tail_q > 0 && g_tail_q < q) ? g_tail_q : q;
    const int64_t tail0 = (qt < q && g_tail_blocks > 0 && g_tail_blocks < nb) ? nb - g_tail_blocks : nb;
    int64_t k = 0;
    while (k < nb) {
      int w = k < tail0 ? q : qt;
      if (k < tail0 && k + w > tail0) w = (int)(tail0 - k);
      if (k + w > nb) w = (int)(nb - k);
      pk.push_back(k);
      pq.push_back(w);
      k += w;
    }
  }
  const int64_t np = (int64_t)pk.size();

  // factorization of outer panel P (inner blocks kb .. kb+q_eff-1) on stream st, packs -> set (P & 1)
  // rest_ready: event after which the rows below the panel's square are up to date (split panels)
  // sq_ready: event after which the part of the panel's square right of its first block column is up to date
  // (split square update, see below); nullptr: everything was ready when `st` was released
  auto factor_panel = [&](int64_t P, cudaStream_t st, cudaEvent_t rest_ready, cudaEvent_t sq_ready) -> int {
    if (P >= np) return SCB_OK;
    const int64_t kb = pk[P];
    const int q_eff = pq[P];
    const int64_t panel_end = (kb + q_eff) * NB;
    double* Lpack = pack_base + (P & 1) * pack_set;
    double* Upack = Lpack + n_pad * NB * q;
    if (split) {
      // Split panel (symmetric): the chain on `st` factors only the panel's diagonal square
      // (diag -> the few column-panel tiles inside the square -> their strip tiles); the rows below
      // the square follow on the second stream, block by block, without ever delaying the chain.
      cudaEvent_t e_in = ls.event(ev++);
      SCB_CUDA(cudaEventRecord(e_in, st));
      SCB_CUDA(cudaStreamWaitEvent(sr, e_in, 0));
      if (rest_ready) SCB_CUDA(cudaStreamWaitEvent(sr, rest_ready, 0));
      const int nt_below = (int)(nb - kb - q_eff);  // 128-row tiles below the outer panel
      cudaEvent_t e_sq_prev = nullptr;  // look-ahead: the square's other columns have received inner panel i - 1
      for (int i = 0; i < q_eff; i++) {
        const int64_t k = kb + i;
        const int64_t o = k * NB;
        double* invL = dinv + k * 2 * NB * NB;
        double* invU = invL + NB * NB;
        const int inner_rem = q_eff - 1 - i;
        if (g_diag_symb)
          diag_kernel_symb<<<1, 256, diag_symb_smem, st>>>(M, n_pad, o, invL, invU, info, (int)k);
        else
          diag_kernel_small<true><<<1, 256, diag_small_smem, st>>>(M, n_pad, o, invL, invU, info, (int)k);
        SCB_LAUNCH_CHECK();
        if (inner_rem == 0 && nt_below == 0) break;
        cudaEvent_t e_d = ls.event(ev++);
        SCB_CUDA(cudaEventRecord(e_d, st));
        if (inner_rem > 0) {
          if (g_lat)
            trsm_sym_lat_kernel<<<8 * inner_rem, 128, kTrsmLatSmem, st>>>(M, n_pad, o, invU, Lpack, Upack,
                                                                         tile_chunks, i * NCHUNK, 0);
          else
            trsm_sym_kernel<<<2 * inner_rem, 256, trsm_smem, st>>>(M, n_pad, o, invU, Lpack, Upack, tile_chunks,
                                                                   i * NCHUNK, 0);
          SCB_LAUNCH_CHECK();
        }
        cudaEvent_t e_t = ls.event(ev++);
        SCB_CUDA(cudaEventRecord(e_t, st));
        if (nt_below > 0) {
          SCB_CUDA(cudaStreamWaitEvent(sr, e_d, 0));
          trsm_sym_kernel<<<2 * nt_below, 256, trsm_smem, sr>>>(M, n_pad, o, invU, Lpack, Upack, tile_chunks,
                                                                i * NCHUNK, 2 * inner_rem);
          SCB_LAUNCH_CHECK();
        }
        if (inner_rem > 0) {
          // Inside the square (the latency-critical chain): eager right-looking updates, K = 128 -- every
          // launch is short (4 k-chunks per CTA) and the next diagonal block is ready right after it.
          const int64_t r1 = (kb + i + 1) * NB;
          if (i == 0 && sq_ready) SCB_CUDA(cudaStreamWaitEvent(st, sq_ready, 0));  // (split square update)
          if (inner_la) {
            // Block-level look-ahead: the chain updates only the block column it factors next (2 x inner_rem
            // CTAs, so it never queues behind its own wide launches for SM slots held by bulk CTAs); the
            // other columns of the square receive inner panel i on a third stream and are awaited one
            // step later.  Per tile the inner panels are still applied in ascending order.
            if (e_sq_prev) SCB_CUDA(cudaStreamWaitEvent(st, e_sq_prev, 0));
            update_kernel_t<false><<<dim3(2, inner_rem), 256, upd_smem, st>>>(
                M, n_pad, r1, r1, Lpack, Upack, tile_chunks, i * NCHUNK, NCHUNK, g_band);
            SCB_LAUNCH_CHECK();
            e_sq_prev = nullptr;
            if (inner_rem > 1) {
              SCB_CUDA(cudaStreamWaitEvent(sq, e_t, 0));
              update_kernel_t<true><<<dim3(2 * (inner_rem - 1), inner_rem - 1), 256, upd_smem, sq>>>(
                  M, n_pad, r1 + NB, r1 + NB, Lpack, Upack, tile_chunks, i * NCHUNK, NCHUNK, g_band);
              SCB_LAUNCH_CHECK();
              e_sq_prev = ls.event(ev++);
              SCB_CUDA(cudaEventRecord(e_sq_prev, sq));
            }
          } else if (g_lat) {
            update_lat_kernel<true><<<dim3(4 * inner_rem, 2 * inner_rem), 128, kUpdLatSmem, st>>>(
                M, n_pad, r1, r1, Lpack, Upack, tile_chunks, i * NCHUNK, NCHUNK);
            SCB_LAUNCH_CHECK();
          } else {
            update_kernel_t<true><<<dim3(2 * inner_rem, inner_rem), 256, upd_smem, st>>>(
                M, n_pad, r1, r1, Lpack, Upack, tile_chunks, i * NCHUNK, NCHUNK, g_band);
            SCB_LAUNCH_CHECK();
          }
          // Rows below the square (throughput work on the second stream): recursive (binary-tree)
          // schedule -- after inner block i the w = 2^tz(i+1) block columns that follow receive the last
          // w inner panels at once, so most of these flops run with K = 256 / 512.
          if (nt_below > 0) {
            int w = 1;
            while (((i + 1) & w) == 0) w <<= 1;
            const int k_lo = i + 1 - w;
            const int c_lo = i + 1, c_hi = (i + 1 + w) < q_eff ? (i + 1 + w) : q_eff;
            const int ncb = c_hi - c_lo;
            const int64_t r0 = (kb + c_lo) * NB;
            SCB_CUDA(cudaStreamWaitEvent(sr, e_t, 0));
            update_kernel_t<false><<<dim3(2 * ncb, nt_below), 256, upd_smem, sr>>>(
                M, n_pad, panel_end, r0, Lpack, Upack, tile_chunks, k_lo * NCHUNK, w * NCHUNK, g_band);
            SCB_LAUNCH_CHECK();
          }
        }
      }
      cudaEvent_t e_out = ls.event(ev++);
      SCB_CUDA(cudaEventRecord(e_out, sr));
      SCB_CUDA(cudaStreamWaitEvent(st, e_out, 0));
      return SCB_OK;
    }
    for (int i = 0; i < q_eff; i++) {
      const int64_t k = kb + i;
      const int64_t o = k * NB;
      double* invL = dinv + k * 2 * NB * NB;
      double* invU = invL + NB * NB;
      const int nt = (int)(nb - k - 1);  // 128-tiles after this block
      if (i > 0 && g_lazy_strips) {
        // left-looking inside the outer panel: bring block column i and block row i up to date
        // with ALL previous inner panels at once (K = 128 i) right before they are factored
        dim3 gc(2, nt + 1);  // rows [o, n) x cols [o, o+128)
        update_kernel_t<false><<<gc, 256, upd_smem, st>>>(M, n_pad, o, o, Lpack, Upack, tile_chunks, 0, i * NCHUNK, g_band);
        SCB_LAUNCH_CHECK();
        if (nt > 0 && !sym) {  // (symmetric: the row panel is mirrored from the column panel)
          dim3 gr(2 * nt, 1);  // rows [o, o+128) x cols [o+128, n)
          update_kernel_t<false><<<gr, 256, upd_smem, st>>>(M, n_pad, o, o + NB, Lpack, Upack, tile_chunks, 0, i * NCHUNK, g_band);
          SCB_LAUNCH_CHECK();
        }
      }
      if (g_diag_small && sym && g_diag_symb)
        diag_kernel_symb<<<1, 256, diag_symb_smem, st>>>(M, n_pad, o, invL, invU, info, (int)k);
      else if (g_diag_small && sym)
        diag_kernel_small<true><<<1, 256, diag_small_smem, st>>>(M, n_pad, o, invL, invU, info, (int)k);
      else if (g_diag_small)
        diag_kernel_small<false><<<1, 256, diag_small_smem, st>>>(M, n_pad, o, invL, invU, info, (int)k);
      else
        diag_kernel<<<1, 512, diag_smem, st>>>(M, n_pad, o, invL, invU, info, (int)k);
      SCB_LAUNCH_CHECK();
      if (nt == 0) break;
      if (sym)
        trsm_sym_kernel<<<2 * nt, 256, trsm_smem, st>>>(M, n_pad, o, invU, Lpack, Upack, tile_chunks, i * NCHUNK, 0);
      else
        trsm_kernel<<<4 * nt, 256, trsm_smem, st>>>(M, n_pad, o, 2 * nt, invL, invU, Lpack, Upack, tile_chunks,
                                                    i * NCHUNK);
      SCB_LAUNCH_CHECK();
      const int inner_rem = q_eff - 1 - i;  // inner blocks still to factor in this outer panel
      if (inner_rem > 0 && g_recursive_strips && !g_lazy_strips) {
        // Recursive (binary-tree) schedule of the updates inside the outer panel: after inner block
        // i the w = 2^tz(i+1) block columns (and, unsymmetric, block rows) that follow receive the
        // contribution of the last w inner panels at once.  Same flops as the eager variant, but 57 %
        // of them run with K = 512 and 29 % with K = 256 instead of all with K = 128.
        int w = 1;
        while (((i + 1) & w) == 0) w <<= 1;
        const int k_lo = i + 1 - w;
        const int c_lo = i + 1, c_hi = (i + 1 + w) < q_eff ? (i + 1 + w) : q_eff;
        const int ncb = c_hi - c_lo;
        const int64_t r0 = (kb + c_lo) * NB;              // first row / column of the band
        const int nrt = (int)(nb - kb - c_lo);            // 128-row tiles from the band to the end
        dim3 ga(2 * ncb, nrt);
        if (sym)
          update_kernel_t<true><<<ga, 256, upd_smem, st>>>(M, n_pad, r0, r0, Lpack, Upack, tile_chunks, k_lo * NCHUNK,
                                                           w * NCHUNK, g_band);
        else
          update_kernel_t<false><<<ga, 256, upd_smem, st>>>(M, n_pad, r0, r0, Lpack, Upack, tile_chunks, k_lo * NCHUNK,
                                                            w * NCHUNK, g_band);
        SCB_LAUNCH_CHECK();
        const int ncright = nrt - ncb;                    // 128-column blocks right of the band
        if (ncright > 0 && !sym) {
          dim3 gb(2 * ncright, ncb);
          update_kernel_t<false><<<gb, 256, upd_smem, st>>>(M, n_pad, r0, r0 + (int64_t)ncb * NB, Lpack, Upack,
                                                            tile_chunks, k_lo * NCHUNK, w * NCHUNK, g_band);
          SCB_LAUNCH_CHECK();
        }
      } else if (inner_rem > 0 && !g_lazy_strips) {
        // right-looking variant: apply inner panel i to the rest of the outer panel's L-shaped strip
        dim3 ga(2 * inner_rem, nt);
        if (sym)  // (tiles above the block diagonal of the panel's own square are never read)
          update_kernel_t<true><<<ga, 256, upd_smem, st>>>(M, n_pad, o + NB, o + NB, Lpack, Upack, tile_chunks,
                                                           i * NCHUNK, NCHUNK, g_band);
        else
          update_kernel_t<false><<<ga, 256, upd_smem, st>>>(M, n_pad, o + NB, o + NB, Lpack, Upack, tile_chunks,
                                                            i * NCHUNK, NCHUNK, g_band);
        SCB_LAUNCH_CHECK();
        const int nright = nt - inner_rem;
        if (nright > 0 && !sym) {  // (symmetric: the row strip is never read)
          dim3 gb(2 * nright, inner_rem);
          update_kernel_t<false><<<gb, 256, upd_smem, st>>>(M, n_pad, o + NB, panel_end, Lpack, Upack, tile_chunks,
                                                   i * NCHUNK, NCHUNK, g_band);
          SCB_LAUNCH_CHECK();
        }
      }
    }
    return SCB_OK;
  };

  // SCB_LU_TRACE=1: time stamps (CUDA events) of the panel chain / strip / bulk phases, printed per call
  static const bool trace = getenv("SCB_LU_TRACE") != nullptr;
  struct Stamp { const char* what; int64_t panel; cudaEvent_t e; };
  std::vector<Stamp> stamps;
  auto stamp = [&](const char* what, int64_t panel, cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    stamps.push_back({what, panel, e});
  };
  stamp("start", -1, s);
  if (lookahead) {
    cudaEvent_t e0 = ls.event(ev++);
    SCB_CUDA(cudaEventRecord(e0, s));
    SCB_CUDA(cudaStreamWaitEvent(sp, e0, 0));
  }
  if (int rc = factor_panel(0, sp, nullptr, nullptr)) return rc;
  stamp("chain_end", 0, sp);
  for (int64_t P = 0; P < np; P++) {
    const int64_t kb = pk[P];
    const int q_eff = pq[P];
    const int64_t e0 = (kb + q_eff) * NB;  // first row/col after panel P
    if (e0 >= n_pad) break;
    const int64_t e1 = e0 + (int64_t)pq[P + 1] * NB;  // end of panel P+1
    const double* Lpack = pack_base + (P & 1) * pack_set;
    const double* Upack = Lpack + n_pad * NB * q;
    const int nchunks = q_eff * NCHUNK;
    if (lookahead) {  // the main stream waits for the factorization of panel P
      cudaEvent_t e = ls.event(ev++);
      SCB_CUDA(cudaEventRecord(e, sp));
      SCB_CUDA(cudaStreamWaitEvent(s, e, 0));
    }
    // A: the L-shaped strip that panel P+1 lives in (rows e0..e1 x all columns, rows below x cols e0..e1)
    const int nt0 = (int)((n_pad - e0) / NB), ntp = (int)((e1 - e0) / NB), nt1 = (int)((n_pad - e1) / NB);
    cudaEvent_t rest_ready = nullptr, sq_ready = nullptr;
    if (sym && split) {
      // symmetric, split panels: the square of panel P+1 first -- its factorization (a latency-bound
      // chain) starts as soon as these 72 tiles are done and runs concurrently with the update of
      // the rows below the square, which only the second (rows-below) stream of the panel waits for
      bool chain_released = false;
      if ((g_lat & 6) == 6 && ntp > 1 && nt1 > 0) {
        // Split square (SCB_LU_LAT=7; off by default, measured slower: the released diagonal block then shares the
        // GPU with the rest of the square): the first diagonal block of the next panel and its panel solve need only the FIRST block
        // column of the square (K = 1024 over 8 row tiles: ~17 us instead of ~44 for the whole square) -- the
        // chain is released behind it; the rest of the square (a triangular region again) is updated beside the
        // first diagonal block and awaited before the chain's first in-square update.  Same element set, one
        // K = 1024 pass per element: bit-identical.
        update_lat_kernel<false><<<dim3(4, 2 * ntp), 128, kUpdLatSmem, s>>>(M, n_pad, e0, e0, Lpack, Upack, tile_chunks,
                                                                           0, nchunks);
        SCB_LAUNCH_CHECK();
        cudaEvent_t e_sq = ls.event(ev++);
        SCB_CUDA(cudaEventRecord(e_sq, s));
        SCB_CUDA(cudaStreamWaitEvent(sp, e_sq, 0));
        chain_released = true;
        update_lat_kernel<true><<<dim3(4 * (ntp - 1), 2 * (ntp - 1)), 128, kUpdLatSmem, s>>>(
            M, n_pad, e0 + NB, e0 + NB, Lpack, Upack, tile_chunks, 0, nchunks);
        SCB_LAUNCH_CHECK();
        sq_ready = ls.event(ev++);
        SCB_CUDA(cudaEventRecord(sq_ready, s));
      } else if (g_lat & 2) {  // the square of the next panel (on the chain as well): 64 x 32 tiles over all SMs
        update_lat_kernel<true><<<dim3(4 * ntp, 2 * ntp), 128, kUpdLatSmem, s>>>(M, n_pad, e0, e0, Lpack, Upack,
                                                                                tile_chunks, 0, nchunks);
        SCB_LAUNCH_CHECK();
      } else {
        update_kernel_t<true><<<dim3(2 * ntp, ntp), 256, upd_smem, s>>>(M, n_pad, e0, e0, Lpack, Upack, tile_chunks, 0,
                                                                         nchunks, g_band);
        SCB_LAUNCH_CHECK();
      }
      if (nt1 > 0) {
        if (!chain_released) {
          cudaEvent_t e_sq = ls.event(ev++);
          SCB_CUDA(cudaEventRecord(e_sq, s));
          SCB_CUDA(cudaStreamWaitEvent(sp, e_sq, 0));
        }
        update_kernel_t<false><<<dim3(2 * ntp, nt1), 256, upd_smem, s>>>(M, n_pad, e1, e0, Lpack, Upack, tile_chunks, 0,
                                                                          nchunks, g_band);
        SCB_LAUNCH_CHECK();
        rest_ready = ls.event(ev++);
        SCB_CUDA(cudaEventRecord(rest_ready, s));
      }
    } else if (sym) {
      // symmetric: one launch for the whole column strip of panel P+1 (its own square and everything below)
      dim3 g12(2 * ntp, nt0);
      update_kernel_t<true><<<g12, 256, upd_smem, s>>>(M, n_pad, e0, e0, Lpack, Upack, tile_chunks, 0, nchunks, g_band);
      SCB_LAUNCH_CHECK();
    } else {
      // rows of panel P+1: all columns to the right
      dim3 g1(2 * nt0, ntp);
      update_kernel_t<false><<<g1, 256, upd_smem, s>>>(M, n_pad, e0, e0, Lpack, Upack, tile_chunks, 0, nchunks, g_band);
      SCB_LAUNCH_CHECK();
      if (nt1 > 0) {
        dim3 g2(2 * ntp, nt1);
        update_kernel_t<false><<<g2, 256, upd_smem, s>>>(M, n_pad, e1, e0, Lpack, Upack, tile_chunks, 0, nchunks, g_band);
        SCB_LAUNCH_CHECK();
      }
    }
    stamp("strip_end", P + 1, s);
    if (lookahead && !rest_ready) {  // panel P+1 can be factored as soon as its strip is up to date
      cudaEvent_t e = ls.event(ev++);
      SCB_CUDA(cudaEventRecord(e, s));
      SCB_CUDA(cudaStreamWaitEvent(sp, e, 0));
    }
    if (int rc = factor_panel(P + 1, sp, rest_ready, sq_ready)) return rc;
    stamp("chain_end", P + 1, sp);
    // B: the rest of the trailing block, concurrently with the factorization of panel P+1
    if (nt1 > 0) {
      if (sym) {  // tiles on / below the diagonal only
        update_kernel_t<true><<<dim3(2 * nt1, nt1), 256, upd_smem, s>>>(M, n_pad, e1, e1, Lpack, Upack, tile_chunks, 0,
                                                                         nchunks, g_band);
      } else {
        dim3 g3(2 * nt1, nt1);
        update_kernel_t<false><<<g3, 256, upd_smem, s>>>(M, n_pad, e1, e1, Lpack, Upack, tile_chunks, 0, nchunks, g_band);
      }
      SCB_LAUNCH_CHECK();
      stamp("bulk_end", P, s);
    }
  }
  if (lookahead) {
    cudaEvent_t e = ls.event(ev++);
    SCB_CUDA(cudaEventRecord(e, sp));
    SCB_CUDA(cudaStreamWaitEvent(s, e, 0));
  }
  if (trace) {
    stamp("end", -1, s);
    cudaStreamSynchronize(s);
    for (size_t k = 1; k < stamps.size(); k++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, stamps[0].e, stamps[k].e);
      fprintf(stderr, "[scb lu trace] %-10s panel %3lld  t = %8.3f ms\n", stamps[k].what, (long long)stamps[k].panel, ms);
    }
    for (auto& st : stamps) cudaEventDestroy(st.e);
  }
  return SCB_OK;
}

// Right-looking LU with partial pivoting, one level of blocking (NB = 128): per panel the pivot
// search (129 small launches), the row interchanges, then the unpivoted diagonal / panel-solve /
// trailing-update kernels of the fast path with K = 128.  A fallback: ~25 TFLOP/s instead of ~32.
static int getrf_piv_impl(int64_t n_pad, double* M, double* dinv, int32_t* piv, int32_t* perm, int32_t* info,
                          scb_stream_t stream) {
  SCB_CHECK_ARG(n_pad > 0 && n_pad % NB == 0, "n_pad must be a positive multiple of 128");
  SCB_CHECK_ARG(piv != nullptr && perm != nullptr, "piv and perm are required");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nb = n_pad / NB;
  const int q = lu_outer_blocks();
  const int tile_chunks = q * NCHUNK;
  double* pack_base = dinv + nb * 2 * NB * NB;
  const int64_t pack_set = 2 * n_pad * NB * q;
  double* Lpack = pack_base;
  double* Upack = Lpack + n_pad * NB * q;
  // scratch of the pivot search: the second pack set (unused by the one-level algorithm)
  double* X = pack_base + pack_set;
  double* Y = X + n_pad * NB;
  PivCand* cand0 = reinterpret_cast<PivCand*>(Y + n_pad * NB);
  const int64_t max_cand = (n_pad + PR - 1) / PR;
  PivCand* cand1 = cand0 + max_cand;
  SCB_CHECK_ARG(2 * n_pad * NB + 4 * max_cand <= pack_set, "factorization workspace too small for pivoting");
  const int diag_small_smem = 3 * QN * QLD * sizeof(double);
  const int trsm_smem = kTrsmSmemDoubles * sizeof(double);
  const int upd_smem = 2 * sizeof(UpdateStage);
  int dev = 0;
  SCB_CUDA(cudaGetDevice(&dev));
  static bool attr_set[64] = {};
  if (!attr_set[dev & 63]) {
    SCB_CUDA(cudaFuncSetAttribute(diag_kernel_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_small_smem));
    SCB_CUDA(cudaFuncSetAttribute(update_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem));
    SCB_CUDA(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem));
    SCB_CUDA(cudaFuncSetAttribute(update_kernel_t<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set[dev & 63] = true;
  }
  SCB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
  SCB_CUDA(cudaMemsetAsync(dinv + lu_flags_offset(n_pad), 0, (nb + 16) * sizeof(double), s));
  perm_init_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, s>>>(n_pad, perm);
  SCB_LAUNCH_CHECK();
  for (int64_t k = 0; k < nb; k++) {
    const int64_t o = k * NB;
    const int64_t nrows = n_pad - o;
    const int ncta = (int)((nrows + PR - 1) / PR);
    // 1. pivot rows of this panel
    piv_init_kernel<<<ncta, 256, 0, s>>>(M, n_pad, o, nrows, X, cand0);
    SCB_LAUNCH_CHECK();
    double *src = X, *dst = Y;
    PivCand *cin = cand0, *cout = cand1;
    for (int j = 0; j < NB; j++) {
      piv_step_kernel<<<ncta, 256, 0, s>>>(src, dst, nrows, j, ncta, cin, cout, o, piv, info);
      SCB_LAUNCH_CHECK();
      double* t = src; src = dst; dst = t;
      PivCand* c = cin; cin = cout; cout = c;
    }
    // 2. interchanges on the full rows
    laswp_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, s>>>(M, n_pad, o, piv, perm);
    SCB_LAUNCH_CHECK();
    // 3. the panel, unpivoted, on the permuted rows
    double* invL = dinv + k * 2 * NB * NB;
    double* invU = invL + NB * NB;
    diag_kernel_small<false><<<1, 256, diag_small_smem, s>>>(M, n_pad, o, invL, invU, info, (int)k);
    SCB_LAUNCH_CHECK();
    const int nt = (int)(nb - k - 1);
    if (nt == 0) break;
    trsm_kernel<<<4 * nt, 256, trsm_smem, s>>>(M, n_pad, o, 2 * nt, invL, invU, Lpack, Upack, tile_chunks, 0);
    SCB_LAUNCH_CHECK();
    update_kernel_t<false><<<dim3(2 * nt, nt), 256, upd_smem, s>>>(M, n_pad, o + NB, o + NB, Lpack, Upack, tile_chunks,
                                                                  0, NCHUNK, g_band);
    SCB_LAUNCH_CHECK();
  }
  return SCB_OK;
}

extern "C" int scb_getrf_piv(int64_t n_pad, double* M, double* dinv, int32_t* piv, int32_t* perm, int32_t* info,
                             scb_stream_t stream) {
  return getrf_piv_impl(n_pad, M, dinv, piv, perm, info, stream);
}

// ---------------------------------------------------------------------------------------
// CUDA-graph replay of a factorization.  The launch sequence of getrf_impl (hundreds of small kernels,
// event records and waits over two or three streams) depends only on (n_pad, sym) and its pointer
// arguments.  A small film factors in a few milliseconds of latency-bound kernels, and enqueueing that
// sequence costs the host about as long as the GPU needs to run it -- several independent films of one
// model (C3 / C4: one stream per film) then start one after the other instead of together.  The second
// time a (n_pad, M, dinv, info, sym) combination is seen (the caching allocator of the host side hands
// the same blocks back in a steady-state loop) the sequence is captured once into a graph and replayed
// from then on with a single cudaGraphLaunch; the fork / join structure over the helper streams
// becomes the branches of the graph.  Identical kernels, identical arguments: bit-identical factors.
// SCB_LU_GRAPH=0 disables it; SCB_LU_GRAPH_MAX_N (default 12288) bounds the matrix size -- large
// factorizations are not latency-bound and rely on stream priorities for their look-ahead.
// ---------------------------------------------------------------------------------------
namespace {
struct LuGraphKey {
  int dev;
  int64_t n_pad;
  double* M;
  double* dinv;
  int32_t* info;
  bool sym;
  bool operator<(const LuGraphKey& o) const {
    if (dev != o.dev) return dev < o.dev;
    if (n_pad != o.n_pad) return n_pad < o.n_pad;
    if (M != o.M) return M < o.M;
    if (dinv != o.dinv) return dinv < o.dinv;
    if (info != o.info) return info < o.info;
    return sym < o.sym;
  }
};
struct LuGraphEntry {
  cudaGraphExec_t exec = nullptr;
  int sightings = 0;
  int nodes = 0;  // kernel launches inside the graph (for scb_launch_count)
  uint64_t last_use = 0;
};
std::map<LuGraphKey, LuGraphEntry> g_lu_graphs;
std::mutex g_lu_graphs_mutex;
int g_lu_graph_enabled = -1;
int64_t g_lu_graph_max_n = 12288;
constexpr size_t kMaxLuGraphs = 96;

uint64_t g_lu_graph_clock = 0;

// table full: forget the keys that were seen once and never again; if every entry holds a graph, drop
// the least recently used one
void lu_graphs_evict_locked() {
  bool dropped = false;
  for (auto it = g_lu_graphs.begin(); it != g_lu_graphs.end();) {
    if (!it->second.exec) {
      it = g_lu_graphs.erase(it);
      dropped = true;
    } else {
      ++it;
    }
  }
  if (dropped || g_lu_graphs.empty()) return;
  auto lru = g_lu_graphs.begin();
  for (auto it = g_lu_graphs.begin(); it != g_lu_graphs.end(); ++it)
    if (it->second.last_use < lru->second.last_use) lru = it;
  cudaGraphExecDestroy(lru->second.exec);
  g_lu_graphs.erase(lru);
}
}  // namespace

static int getrf_graphed(int64_t n_pad, double* M, double* dinv, int32_t* info, scb_stream_t stream, bool sym) {
  if (g_lu_graph_enabled < 0) {
    int en = 1;
    if (const char* e = getenv("SCB_LU_GRAPH")) en = atoi(e);
    if (getenv("SCB_LU_TRACE")) en = 0;
    if (const char* e = getenv("SCB_LU_GRAPH_MAX_N")) g_lu_graph_max_n = atoll(e);
    g_lu_graph_enabled = en;
  }
  if (!g_lu_graph_enabled || n_pad > g_lu_graph_max_n || n_pad <= 0 || n_pad % NB != 0)
    return getrf_impl(n_pad, M, dinv, info, stream, sym);
  cudaStream_t s = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return getrf_impl(n_pad, M, dinv, info, stream, sym);  // the caller is capturing: just take part in it
  }
  int dev = 0;
  SCB_CUDA(cudaGetDevice(&dev));
  const LuGraphKey key{dev, n_pad, M, dinv, info, sym};
  cudaGraphExec_t exec = nullptr;
  int nodes = 0;
  bool capture = false;
  {
    std::lock_guard<std::mutex> lock(g_lu_graphs_mutex);
    if (g_lu_graphs.size() >= kMaxLuGraphs && g_lu_graphs.find(key) == g_lu_graphs.end()) lu_graphs_evict_locked();
    LuGraphEntry& e = g_lu_graphs[key];
    e.sightings++;
    e.last_use = ++g_lu_graph_clock;
    exec = e.exec;
    nodes = e.nodes;
    capture = exec == nullptr && e.sightings >= 2;
  }
  if (exec) {
    SCB_CUDA(cudaGraphLaunch(exec, s));
    scb::count_launch(nodes);
    return SCB_OK;
  }
  if (!capture) return getrf_impl(n_pad, M, dinv, info, stream, sym);
  // Capture the sequence (nothing runs), instantiate, launch.  The capture runs on a stream of the
  // library's own (one per device, under a lock): the caller's stream may be the legacy default stream,
  // which cannot be captured, and is never put into capture mode.
  static std::mutex capture_mutex;
  static cudaStream_t capture_streams[64] = {};
  std::lock_guard<std::mutex> capture_lock(capture_mutex);
  cudaStream_t& cs = capture_streams[dev & 63];
  if (cs == nullptr && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    cs = nullptr;
    return getrf_impl(n_pad, M, dinv, info, stream, sym);
  }
  const int64_t before = scb_launch_count();
  if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return getrf_impl(n_pad, M, dinv, info, stream, sym);
  }
  const int rc = getrf_impl(n_pad, M, dinv, info, (scb_stream_t)cs, sym);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
  const int captured = (int)(scb_launch_count() - before);
  scb::count_launch(-captured);  // nothing was launched yet
  if (rc != SCB_OK || ce != cudaSuccess || graph == nullptr) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    g_lu_graph_enabled = 0;  // do not try again in this process
    return getrf_impl(n_pad, M, dinv, info, stream, sym);
  }
  cudaGraphExec_t new_exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&new_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess || new_exec == nullptr) {
    cudaGetLastError();
    g_lu_graph_enabled = 0;
    return getrf_impl(n_pad, M, dinv, info, stream, sym);
  }
  {
    std::lock_guard<std::mutex> lock(g_lu_graphs_mutex);
    LuGraphEntry& e = g_lu_graphs[key];
    if (e.exec) {  // (another thread was faster)
      cudaGraphExecDestroy(new_exec);
      new_exec = e.exec;
    } else {
      e.exec = new_exec;
      e.nodes = captured;
    }
  }
  SCB_CUDA(cudaGraphLaunch(new_exec, s));
  scb::count_launch(captured);
  return SCB_OK;
}

extern "C" int scb_getrf_nopiv(int64_t n_pad, double* M, double* dinv, int32_t* info, scb_stream_t stream) {
  return getrf_graphed(n_pad, M, dinv, info, stream, false);
}

extern "C" int scb_getrf_sym_nopiv(int64_t n_pad, double* M, double* dinv, int32_t* info,
                                   scb_stream_t stream) {
  return getrf_graphed(n_pad, M, dinv, info, stream, true);
}
