// Multi-RHS triangular solves with the LU factors (K12): scipy.linalg.lu_solve at
// solver/solve_film.py:367,388,530,545.
//
// One PERSISTENT kernel per sweep (forward L y = b, backward U x = y).  The factor is processed in
// 128-row blocks; CTAs grab (block, column-chain) work items in sweep order from an atomic counter (so
// a CTA only ever waits for items that are already owned by running or finished CTAs: deadlock-free
// without co-residency assumptions).  For its block i a CTA streams the block row of the factor,
//     acc_i = b_i - sum_{j before i} F_ij x_j ,
// waiting on a per-block ready flag before it touches x_j, then finishes the block with the stored
// inverse of the diagonal block, x_i = inv(F_ii) acc_i, publishes x_i (threadfence + flag) and grabs
// the next item.  What sits on the critical chain of a block is kept minimal:
//   * the factor tile F_ij is in registers BEFORE the flag of x_j is awaited, and the inverse of the
//     diagonal block was staged into shared memory when the block was grabbed;
//   * one right-hand side: the 16 row sums of a warp are reduced together by a transposing butterfly
//     (16 shuffles instead of 16 x 5);
//   * several right-hand sides: both 128x128 products of a block run on the fp64 tensor cores
//     (mma.sync m8n8k4, SASS DMMA), 8 or 16 columns per chain, and all the column chains of a
//     launch advance CONCURRENTLY (work item = (block, chain), chains of the same block row are
//     grabbed back to back so that they share the factor tiles in L2).
//   * no memory fence on the chain: a finished x entry is published as a 16-byte packet {value, tag}
//     written with ONE 128-bit store and read with 128-bit volatile loads (value and tag arrive
//     together, as in decoupled look-back scans), so the data validates itself.  A per-block flag,
//     stored right behind the packets WITHOUT a fence, only keeps the CTAs that have caught up with
//     the head of the chain from polling with all their threads (one thread per CTA spins on it).
// Packets and flags live in a stream-ordered scratch allocation that is zeroed per sweep (no epochs:
// the launch sequence is replayable, e.g. from a CUDA graph).
#include "scb_common.cuh"

namespace scb {

constexpr int NB = SCB_LU_BLOCK;
constexpr int DLDS = NB + 4;  // row stride of the shared copy of inv(F_ii): conflict-free A fragments

__device__ __forceinline__ void dmma_rhs(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
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

// inv(F_ii) (128 x 128, row stride 128) -> shared [128][DLDS]
__device__ __forceinline__ void stage_inverse(double* __restrict__ dsm, const double* __restrict__ dblk) {
#pragma unroll 8
  for (int idx = threadIdx.x; idx < NB * NB / 2; idx += 256) {
    const int r = idx >> 6, c = (idx & 63) * 2;
    const double2 v = *reinterpret_cast<const double2*>(dblk + r * NB + c);
    *reinterpret_cast<double2*>(&dsm[r * DLDS + c]) = v;
  }
}

// a published solution entry: value and ready tag in one naturally aligned 16-byte word
struct __align__(16) Packet {
  double value;
  unsigned long long tag;
};

__device__ __forceinline__ void publish(Packet* p, double v) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};\n" ::"l"(p), "l"(__double_as_longlong(v)), "l"(1ull)
               : "memory");
}

// one thread of the CTA spins on the (unfenced) block flag; the packets behind it validate themselves
__device__ __forceinline__ void wait_block(const int* flag) {
  if (threadIdx.x == 0) {
    while (*reinterpret_cast<const volatile int*>(flag) == 0) {
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void signal_block(int* flag) {
  __syncthreads();  // every thread has issued its packet stores
  if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(flag) = 1;
}

// spins until the packet has been published and returns its value (the 128-bit load delivers value
// and tag together)
__device__ __forceinline__ double consume(const Packet* p) {
  unsigned long long v, tag;
  do {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(v), "=l"(tag) : "l"(p) : "memory");
  } while (tag == 0ull);
  return __longlong_as_double((long long)v);
}

// Sums 16 per-lane values over the 32 lanes of a warp with a transposing butterfly: 16 shuffles.
// On return lane L holds the total of value (L >> 1) & 15 (both lanes of a pair hold it).
__device__ __forceinline__ double reduce16(double (&a)[16], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double send = up ? a[k] : a[k + 8];
      const double keep = up ? a[k + 8] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double send = up ? a[k] : a[k + 4];
      const double keep = up ? a[k + 4] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const double send = up ? a[k] : a[k + 2];
      const double keep = up ? a[k + 2] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    const double send = up ? a[0] : a[1];
    const double keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// ---------------------------------------------------------------------------------------
// one right-hand side.  Warp w owns rows w + 8 k (k < 16) of the block; after reduce16 lane L holds
// the value of row w + 8 (L >> 1).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
trsv_sweep_kernel(const double* __restrict__ F, int64_t ld, const double* __restrict__ dinv, int64_t nb, int lower,
                  double* __restrict__ B, Packet* __restrict__ packets, int* __restrict__ flags,
                  int* __restrict__ counter) {
  extern __shared__ __align__(16) double dsm[];  // [NB][DLDS]
  __shared__ double xs[NB];
  __shared__ int s_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int myrow = warp + 8 * (lane >> 1);
  const int64_t ldb = ld * (int64_t)sizeof(double);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_p = atomicAdd(counter, 1);
    __syncthreads();
    const int p = s_p;  // position in sweep order
    if (p >= nb) break;
    const int64_t i = lower ? p : nb - 1 - p;
    if (p > 0) prefetch_tile_l2(F + i * NB * ld + (lower ? 0 : nb - 1) * NB, ldb, warp, lane);
    stage_inverse(dsm, dinv + i * 2 * NB * NB + (lower ? 0 : NB * NB));
    double bval = B[i * NB + myrow];
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
      if (q + 1 < p) prefetch_tile_l2(F + i * NB * ld + (lower ? q + 1 : nb - 2 - q) * NB, ldb, warp, lane);
      wait_block(flags + j);
      if (tid < NB) xs[tid] = consume(packets + j * NB + tid);
      __syncthreads();
      double s[16];
      const double x0 = xs[lane], x1 = xs[lane + 32], x2 = xs[lane + 64], x3 = xs[lane + 96];
#pragma unroll
      for (int rr = 0; rr < 16; rr++) s[rr] = fma(fa[rr][3], x3, fma(fa[rr][2], x2, fma(fa[rr][1], x1, fa[rr][0] * x0)));
      bval -= reduce16(s, lane);
      __syncthreads();  // xs is rewritten for the next tile
    }
    // finish: x_i = inv(F_ii) acc_i with the inverse already in shared memory
    if ((lane & 1) == 0) xs[myrow] = bval;
    __syncthreads();  // (also orders stage_inverse's stores before the reads below)
    {
      double s[16];
      const double x0 = xs[lane], x1 = xs[lane + 32], x2 = xs[lane + 64], x3 = xs[lane + 96];
#pragma unroll
      for (int rr = 0; rr < 16; rr++) {
        const double* Drow = dsm + (warp + 8 * rr) * DLDS;
        s[rr] = fma(Drow[lane + 96], x3, fma(Drow[lane + 64], x2, fma(Drow[lane + 32], x1, Drow[lane] * x0)));
      }
      const double v = reduce16(s, lane);
      if ((lane & 1) == 0) {
        publish(packets + i * NB + myrow, v);
        B[i * NB + myrow] = v;
      }
    }
    signal_block(flags + i);
  }
}

// ---------------------------------------------------------------------------------------
// RC = 8 or 16 right-hand sides per chain on the fp64 tensor cores.  Work item p -> block p / nchains,
// chain p % nchains (columns c_base + RC * chain ..).  Warp w owns rows 16 w .. 16 w + 15 of the block:
// C fragments (rows 16 w + 8 rt + g, columns 8 ct + 2 t, + 1), A fragments of the factor tile loaded
// straight from global memory before the ready flag of x_j is awaited.
// ---------------------------------------------------------------------------------------
// PACKETS: x travels in self-validating packets behind an unfenced flag (single chain: measured faster);
// otherwise the consumers read x from B behind a fenced flag (several concurrent chains: the packet
// traffic of all the chains at the head of the sweep measured slower than two fences per block).
template <int RC, bool PACKETS>
__global__ void __launch_bounds__(256, 1)
trsm_sweep_dmma_kernel(const double* __restrict__ F, int64_t ld, const double* __restrict__ dinv, int64_t nb,
                       int lower, int64_t nrhs, int64_t c_base, int nchains, double* __restrict__ B,
                       Packet* __restrict__ packets, int* __restrict__ flags, int* __restrict__ counter) {
  constexpr int XP = RC + 4;
  constexpr int PK = NB * RC / 256;  // packets consumed per thread and tile   // row stride of the shared x tile: conflict-free B fragments
  constexpr int CT = RC / 8;   // 8-column tiles per chain
  extern __shared__ __align__(16) double smem[];
  double* dsm = smem;              // [NB][DLDS]
  double* xs = smem + NB * DLDS;   // [NB][XP]
  __shared__ int s_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t ldb = ld * (int64_t)sizeof(double);
  const int64_t nitems = nb * nchains;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_p = atomicAdd(counter, 1);
    __syncthreads();
    const int64_t item = s_p;
    if (item >= nitems) break;
    const int p = (int)(item / nchains);  // position of the block in sweep order
    const int ch = (int)(item % nchains);
    const int64_t i = lower ? p : nb - 1 - p;
    const int64_t c0 = c_base + (int64_t)ch * RC;
    const int nr = (int)((nrhs - c0) < RC ? (nrhs - c0) : RC);
    Packet* cpackets = packets + (int64_t)ch * nb * (NB * RC);  // [nb][NB][RC] of this chain
    int* cflags = flags + (int64_t)ch * nb;
    if (p > 0) prefetch_tile_l2(F + i * NB * ld + (lower ? 0 : nb - 1) * NB, ldb, warp, lane);
    stage_inverse(dsm, dinv + i * 2 * NB * NB + (lower ? 0 : NB * NB));
    double acc[2][CT][2];
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int ct = 0; ct < CT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = ct * 8 + 2 * t + e;
          acc[rt][ct][e] = c < nr ? B[(i * NB + warp * 16 + rt * 8 + g) * nrhs + c0 + c] : 0.0;
        }
    // The A fragments of a factor tile are fetched in two halves (k < 64, k >= 64) that are software
    // pipelined across tiles: while one half feeds the tensor cores the other one is in flight, so the
    // global-memory latency of a tile (the CTA is alone on its SM) hides behind the previous tile's DMMAs
    // and the wait for x_j instead of preceding every tile.
    constexpr int HK = NB / 8;  // k4 steps per half
    double aA[2][HK], aB[2][HK];
    const double* Frow0 = F + (i * NB + warp * 16 + g) * ld + t;  // row of rt = 0; rt = 1: + 8 ld
    auto load_half = [&](double (&dst)[2][HK], int64_t jt, int half) {
#pragma unroll
      for (int rt = 0; rt < 2; rt++) {
        const double* Fr = Frow0 + (int64_t)rt * 8 * ld + jt * NB + half * (NB / 2);
#pragma unroll
        for (int ks = 0; ks < HK; ks++) dst[rt][ks] = Fr[ks * 4];
      }
    };
    if (p > 0) load_half(aA, lower ? 0 : nb - 1, 0);
    for (int q = 0; q < p; q++) {
      const int64_t j = lower ? q : nb - 1 - q;
      load_half(aB, j, 1);
      if (q + 1 < p) prefetch_tile_l2(F + i * NB * ld + (lower ? q + 1 : nb - 2 - q) * NB, ldb, warp, lane);
      if (PACKETS) {
        wait_block(cflags + j);
        // all of this thread's packets are requested together; (rare) retry until every tag is set
        const Packet* pk = cpackets + j * (NB * RC) + tid;
        unsigned long long v[PK], tg[PK];
        bool ready;
        do {
          ready = true;
#pragma unroll
          for (int k = 0; k < PK; k++) {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(v[k]), "=l"(tg[k]) : "l"(pk + 256 * k) : "memory");
          }
#pragma unroll
          for (int k = 0; k < PK; k++) ready = ready && (tg[k] != 0ull);
        } while (!ready);
#pragma unroll
        for (int k = 0; k < PK; k++) {
          const int idx = tid + 256 * k;
          xs[(idx / RC) * XP + idx % RC] = __longlong_as_double((long long)v[k]);
        }
      } else {
        if (tid == 0) {
          while (*reinterpret_cast<const volatile int*>(cflags + j) == 0) {
          }
          __threadfence();
        }
        __syncthreads();
        for (int idx = tid; idx < NB * RC; idx += 256) {
          const int r = idx / RC, c = idx % RC;
          xs[r * XP + c] = c < nr ? __ldcg(&B[(j * NB + r) * nrhs + c0 + c]) : 0.0;
        }
      }
      __syncthreads();
      // The two k halves of the tile accumulate into separate registers (acc / acc2): on the critical tile of a
      // block step (x_j just published) the dependent DMMA chain per accumulator is 16 long instead of 32.
      double acc2[2][CT][2];
#pragma unroll
      for (int rt = 0; rt < 2; rt++)
#pragma unroll
        for (int ct = 0; ct < CT; ct++) acc2[rt][ct][0] = acc2[rt][ct][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < HK; ks++) {
#pragma unroll
        for (int ct = 0; ct < CT; ct++) {
          const double b = xs[(ks * 4 + t) * XP + ct * 8 + g];
          const double b2 = xs[((HK + ks) * 4 + t) * XP + ct * 8 + g];
          dmma_rhs(acc[0][ct][0], acc[0][ct][1], -aA[0][ks], b);
          dmma_rhs(acc[1][ct][0], acc[1][ct][1], -aA[1][ks], b);
          dmma_rhs(acc2[0][ct][0], acc2[0][ct][1], -aB[0][ks], b2);
          dmma_rhs(acc2[1][ct][0], acc2[1][ct][1], -aB[1][ks], b2);
        }
      }
#pragma unroll
      for (int rt = 0; rt < 2; rt++)
#pragma unroll
        for (int ct = 0; ct < CT; ct++) {
          acc[rt][ct][0] += acc2[rt][ct][0];
          acc[rt][ct][1] += acc2[rt][ct][1];
        }
      if (q + 1 < p) load_half(aA, lower ? q + 1 : nb - 2 - q, 0);  // first half of the next tile
      __syncthreads();  // xs is rewritten for the next tile
    }
    // finish: x_i = inv(F_ii) acc_i, A fragments from the shared copy of the inverse
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int ct = 0; ct < CT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) xs[(warp * 16 + rt * 8 + g) * XP + ct * 8 + 2 * t + e] = acc[rt][ct][e];
    __syncthreads();
    double out[2][CT][2], out2[2][CT][2];  // (even / odd k4 steps: two dependent chains of 16 instead of one of 32)
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int ct = 0; ct < CT; ct++) out[rt][ct][0] = out[rt][ct][1] = out2[rt][ct][0] = out2[rt][ct][1] = 0.0;
    const double* D0 = dsm + (warp * 16 + g) * DLDS + t;
    const double* D1 = D0 + 8 * DLDS;
#pragma unroll 4
    for (int ks = 0; ks < NB / 4; ks += 2) {
      const double a0 = D0[ks * 4], a1 = D1[ks * 4];
      const double c0 = D0[(ks + 1) * 4], c1 = D1[(ks + 1) * 4];
#pragma unroll
      for (int ct = 0; ct < CT; ct++) {
        const double b = xs[(ks * 4 + t) * XP + ct * 8 + g];
        const double b2 = xs[((ks + 1) * 4 + t) * XP + ct * 8 + g];
        dmma_rhs(out[0][ct][0], out[0][ct][1], a0, b);
        dmma_rhs(out[1][ct][0], out[1][ct][1], a1, b);
        dmma_rhs(out2[0][ct][0], out2[0][ct][1], c0, b2);
        dmma_rhs(out2[1][ct][0], out2[1][ct][1], c1, b2);
      }
    }
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int ct = 0; ct < CT; ct++) {
        out[rt][ct][0] += out2[rt][ct][0];
        out[rt][ct][1] += out2[rt][ct][1];
      }
#pragma unroll
    for (int rt = 0; rt < 2; rt++)
#pragma unroll
      for (int ct = 0; ct < CT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = ct * 8 + 2 * t + e;
          const int r = warp * 16 + rt * 8 + g;
          if (PACKETS) publish(cpackets + i * (NB * RC) + r * RC + c, out[rt][ct][e]);  // (columns >= nr: zeros)
          if (c < nr) B[(i * NB + r) * nrhs + c0 + c] = out[rt][ct][e];
        }
    if (PACKETS) {
      signal_block(cflags + i);
    } else {
      // the barrier orders every thread's x_i stores before thread 0's fence (cumulative at gpu scope),
      // which orders them before the flag store
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        *reinterpret_cast<volatile int*>(cflags + i) = 1;
      }
    }
  }
}

constexpr int kMaxChains = 8;  // column chains advanced concurrently by one launch
static int g_sms[64] = {};
static bool g_attr[64] = {};

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
  const int smem1 = NB * DLDS * sizeof(double);
  const int smem8 = (NB * DLDS + NB * (8 + 4)) * sizeof(double);
  const int smem16 = (NB * DLDS + NB * (16 + 4)) * sizeof(double);
  if (!g_attr[dev & 63]) {
    SCB_CUDA(cudaFuncSetAttribute(trsv_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    SCB_CUDA(cudaFuncSetAttribute(trsm_sweep_dmma_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8));
    SCB_CUDA(cudaFuncSetAttribute(trsm_sweep_dmma_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    SCB_CUDA(cudaDeviceGetAttribute(&g_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    cudaMemPool_t pool;  // keep the freed flag scratch in the stream-ordered pool
    SCB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = ~0ull;
    SCB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    g_attr[dev & 63] = true;
  }
  const int sms = g_sms[dev & 63] > 0 ? g_sms[dev & 63] : 148;  // one CTA per SM (shared-memory bound)
  // column chains: one right-hand side, one chain of 8, or chains of 16 (at most kMaxChains per pass)
  // (measured at 20k, 64 right-hand sides: 4 fenced chains of 16 columns 5.2 ms; 8 packet chains of 8 columns
  //  6.5 ms; 16-column chains streamed through a cp.async shared-memory pipeline 5.7 ms)
  const int rc = nrhs == 1 ? 1 : (nrhs <= 8 ? 8 : 16);
  const int64_t chains_total = (nrhs + rc - 1) / rc;
  const int64_t max_chains = chains_total < kMaxChains ? chains_total : kMaxChains;
  // [max_chains][nb][128][rc] solution packets + [max_chains][nb] block flags + work counter
  const size_t packet_bytes = rc == 16 ? 0 : sizeof(Packet) * (size_t)(max_chains * nb * NB * rc);
  const size_t scratch_bytes = packet_bytes + sizeof(int) * (size_t)(max_chains * nb + 4);
  char* scratch = nullptr;
  SCB_CUDA(cudaMallocAsync(&scratch, scratch_bytes, s));
  Packet* packets = reinterpret_cast<Packet*>(scratch);
  int* flags = reinterpret_cast<int*>(scratch + packet_bytes);
  int* counter = flags + max_chains * nb;
  for (int lower = 1; lower >= 0; lower--) {
    for (int64_t ch0 = 0; ch0 < chains_total; ch0 += kMaxChains) {
      const int nchains = (int)((chains_total - ch0) < kMaxChains ? (chains_total - ch0) : kMaxChains);
      SCB_CUDA(cudaMemsetAsync(scratch, 0, scratch_bytes, s));
      const int64_t items = nb * nchains;
      const int grid = (int)(items < sms ? items : sms);
      if (rc == 1)
        trsv_sweep_kernel<<<grid, 256, smem1, s>>>(LU, n_pad, dinv, nb, lower, B, packets, flags, counter);
      else if (rc == 8)
        trsm_sweep_dmma_kernel<8, true><<<grid, 256, smem8, s>>>(LU, n_pad, dinv, nb, lower, nrhs, ch0 * rc, nchains, B,
                                                           packets, flags, counter);
      else
        trsm_sweep_dmma_kernel<16, false><<<grid, 256, smem16, s>>>(LU, n_pad, dinv, nb, lower, nrhs, ch0 * rc, nchains,
                                                                    B, packets, flags, counter);
      SCB_LAUNCH_CHECK();
    }
  }
  SCB_CUDA(cudaFreeAsync(scratch, s));
  return SCB_OK;
}
