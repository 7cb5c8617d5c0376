// Variants of the 8x8 LDL^T tile factorization that sits on the critical chain of diag_kernel_symb
// (one warp, every lane holds the whole lower triangle), timed with clock64 inside the kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/phase_a_bench.cu -o build/phase_a_bench
#include <cstdio>
#include <cmath>
__device__ __forceinline__ double rcp_halley(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  const double e = fma(-a, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
constexpr int LD = 68;

// v0: current production code (branchy stores)
__device__ __forceinline__ void tile_v0(double* S, double* dd, double* rd, int lane, int& bad) {
  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) a[r][c] = S[r * LD + c];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const double djj = a[j][j];
    const double rjj = rcp_halley(djj);
    double l[8];
#pragma unroll
    for (int r = j + 1; r < 8; r++) l[r] = a[r][j] * rjj;
#pragma unroll
    for (int r = j + 1; r < 8; r++) {
#pragma unroll
      for (int c = j + 1; c <= r; c++) a[r][c] = fma(-l[r], a[c][j], a[r][c]);
      if (lane == 8 + r) { S[r * LD + j] = l[r]; S[j * LD + r] = a[r][j]; }
    }
    if (lane == j) {
      dd[j] = djj; rd[j] = rjj; S[j * LD + j] = djj;
      if (bad == 0 && !(fabs(djj) > 0.0 && isfinite(djj))) bad = j + 1;
    }
  }
}
// v1: no stores inside the pivot loop; results written afterwards by distinct lanes (predicated, no branches)
__device__ __forceinline__ void tile_v1(double* S, double* dd, double* rd, int lane, int& bad) {
  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) a[r][c] = S[r * LD + c];
  __syncwarp();
  double L[8][8], d[8], rdv[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    d[j] = a[j][j];
    rdv[j] = rcp_halley(d[j]);
#pragma unroll
    for (int r = j + 1; r < 8; r++) L[r][j] = a[r][j] * rdv[j];
#pragma unroll
    for (int r = j + 1; r < 8; r++)
#pragma unroll
      for (int c = j + 1; c <= r; c++) a[r][c] = fma(-L[r][j], a[c][j], a[r][c]);
  }
  int b = 0;
#pragma unroll
  for (int j = 7; j >= 0; j--) b = !(fabs(d[j]) > 0.0 && isfinite(d[j])) ? j + 1 : b;
  bad = bad ? bad : b;
  // lane r (< 8) stores row r of L and column r of U; lane 8 + j stores d_j, 1 / d_j
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < r; c++)
      if (lane == r) { S[r * LD + c] = L[r][c]; S[c * LD + r] = a[r][c]; }
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (lane == 8 + j) { dd[j] = d[j]; rd[j] = rdv[j]; S[j * LD + j] = d[j]; }
}
// v2: fraction-free (Bareiss) elimination: the division of step j is by the pivot of step j - 1, whose
// reciprocal is off the chain.  m[r][c] after step j holds the (j+1) x (j+1) bordered minors.
__device__ __forceinline__ void tile_v2(double* S, double* dd, double* rd, int lane, int& bad) {
  double m[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) m[r][c] = S[r * LD + c];
  __syncwarp();
  double piv[8], rpiv[8], col[8][8];
  double rprev = 1.0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    piv[j] = m[j][j];
#pragma unroll
    for (int r = j + 1; r < 8; r++) col[r][j] = m[r][j];
#pragma unroll
    for (int r = j + 1; r < 8; r++)
#pragma unroll
      for (int c = j + 1; c <= r; c++) m[r][c] = fma(piv[j], m[r][c], -(col[r][j] * col[c][j])) * rprev;
    rpiv[j] = rcp_halley(piv[j]);  // needed by the NEXT step only
    rprev = rpiv[j];
  }
  // d_j = piv_j / piv_{j-1};  L[r][j] = col[r][j] / piv_j;  U[j][r] = d_j L[r][j] = col[r][j] / piv_{j-1}
  int b = 0;
#pragma unroll
  for (int j = 7; j >= 0; j--) b = !(fabs(piv[j]) > 0.0 && isfinite(piv[j])) ? j + 1 : b;
  bad = bad ? bad : b;
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < r; c++)
      if (lane == r) { S[r * LD + c] = col[r][c] * rpiv[c]; S[c * LD + r] = col[r][c] * (c ? rpiv[c - 1] : 1.0); }
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (lane == 8 + j) {
      const double dj = piv[j] * (j ? rpiv[j - 1] : 1.0);
      dd[j] = dj; rd[j] = rpiv[j] * (j ? piv[j - 1] : 1.0); S[j * LD + j] = dj;
    }
}

// v3: every lane stores every result (same value to the same address: one wavefront, no predicates, no
// divergent blocks); v4: the U row of a pivot additionally as 128-bit stores
template <bool VEC, bool LANE0 = false>
__device__ __forceinline__ void tile_v3(double* S, double* dd, double* rd, int lane, int& bad) {
  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) a[r][c] = S[r * LD + c];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const double djj = a[j][j];
    const double rjj = rcp_halley(djj);
    double l[8];
#pragma unroll
    for (int r = j + 1; r < 8; r++) l[r] = a[r][j] * rjj;
#pragma unroll
    for (int r = j + 1; r < 8; r++)
#pragma unroll
      for (int c = j + 1; c <= r; c++) a[r][c] = fma(-l[r], a[c][j], a[r][c]);
    if (LANE0 && lane != 0) {
      bad = (bad == 0 && !(fabs(djj) > 0.0 && isfinite(djj))) ? j + 1 : bad;
      continue;
    }
#pragma unroll
    for (int r = j + 1; r < 8; r++) S[r * LD + j] = l[r];
    if (VEC) {
      if (j & 1) {
        S[j * LD + j] = djj;
#pragma unroll
        for (int c = j + 1; c < 8; c += 2) *reinterpret_cast<double2*>(&S[j * LD + c]) = make_double2(a[c][j], a[c + 1][j]);
      } else {
        *reinterpret_cast<double2*>(&S[j * LD + j]) = make_double2(djj, a[j + 1][j]);
#pragma unroll
        for (int c = j + 2; c < 8; c += 2) *reinterpret_cast<double2*>(&S[j * LD + c]) = make_double2(a[c][j], a[c + 1][j]);
      }
    } else {
      S[j * LD + j] = djj;
#pragma unroll
      for (int r = j + 1; r < 8; r++) S[j * LD + r] = a[r][j];
    }
    dd[j] = djj;
    rd[j] = rjj;
    bad = (bad == 0 && !(fabs(djj) > 0.0 && isfinite(djj))) ? j + 1 : bad;
  }
}

template <int V>
__global__ void bench(const double* A0, double* out, long long* cyc, int reps) {
  __shared__ double S[8 * LD], dd[8], rd[8];
  const int lane = threadIdx.x;
  int bad = 0;
  long long total = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = lane; i < 64; i += 32) S[(i / 8) * LD + i % 8] = A0[i] + 1e-9 * rep;
    __syncwarp();
    const long long t0 = clock64();
    if (V == 0) tile_v0(S, dd, rd, lane, bad);
    if (V == 1) tile_v1(S, dd, rd, lane, bad);
    if (V == 2) tile_v2(S, dd, rd, lane, bad);
    if (V == 3) tile_v3<false>(S, dd, rd, lane, bad);
    if (V == 4) tile_v3<true>(S, dd, rd, lane, bad);
    if (V == 5) tile_v3<false, true>(S, dd, rd, lane, bad);
    __syncwarp();
    total += clock64() - t0;
  }
  if (lane == 0) cyc[V] = total / reps;
  for (int i = lane; i < 64; i += 32) out[V * 80 + i] = S[(i / 8) * LD + i % 8];
  if (lane < 8) { out[V * 80 + 64 + lane] = dd[lane]; out[V * 80 + 72 + lane] = rd[lane]; }
  if (lane == 0 && bad) printf("bad %d\n", bad);
}

// dependent-issue latencies, measured so that the compiler cannot drop the chain
__global__ void lat(double* out, long long* cyc, double x0, double y0, int iters) {
  double x = x0, y = y0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
  long long t1 = clock64();
  out[0] = x; cyc[0] = (t1 - t0) / (4 * iters);
  x = x0; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) { x = rcp_halley(x); x = rcp_halley(x); }
  t1 = clock64();
  out[1] = x; cyc[1] = (t1 - t0) / (2 * iters);
  x = x0; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.0; }
  t1 = clock64();
  out[2] = x; cyc[2] = (t1 - t0) / iters;
}

int main() {
  double h[64];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) h[i * 8 + j] = (i == j) ? 200.0 + i : 1.0 / (1.0 + abs(i - j));
  double *A0, *out; long long* cyc;
  cudaMalloc(&A0, sizeof(h)); cudaMalloc(&out, 6 * 80 * 8); cudaMalloc(&cyc, 64);
  cudaMemcpy(A0, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int pass = 0; pass < 2; pass++) {
    bench<0><<<1, 32>>>(A0, out, cyc, 200);
    bench<1><<<1, 32>>>(A0, out, cyc, 200);
    bench<2><<<1, 32>>>(A0, out, cyc, 200);
    bench<3><<<1, 32>>>(A0, out, cyc, 200);
    bench<4><<<1, 32>>>(A0, out, cyc, 200);
    bench<5><<<1, 32>>>(A0, out, cyc, 200);
    cudaDeviceSynchronize();
  }
  long long hc[8]; double ho[480];
  cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
  printf("8x8 LDL^T tile, one warp: v0 (production) %lld cycles, v1 (stores after the loop) %lld, v2 (fraction-free) %lld, v3 (uniform stores) %lld, v4 (uniform + 128-bit) %lld, v5 (lane 0 stores) %lld\n", hc[0], hc[1], hc[2], hc[3], hc[4], hc[5]);
  double e1 = 0, e2 = 0;
  for (int i = 0; i < 80; i++) { e1 = fmax(e1, fabs(ho[80 + i] - ho[i]) / fmax(1e-300, fabs(ho[i]))); e2 = fmax(e2, fabs(ho[160 + i] - ho[i]) / fmax(1e-300, fabs(ho[i]))); }
  printf("max relative difference of the outputs: v1 vs v0 %.3e, v2 vs v0 %.3e\n", e1, e2);
  double e3 = 0, e4 = 0;
  for (int i = 0; i < 80; i++) { e3 = fmax(e3, fabs(ho[240 + i] - ho[i])); e4 = fmax(e4, fabs(ho[320 + i] - ho[i])); }
  printf("max absolute difference: v3 vs v0 %.3e, v4 vs v0 %.3e\n", e3, e4);
  lat<<<1, 32>>>(out, cyc, 1.5, 1.0000001, 2000); cudaDeviceSynchronize();
  lat<<<1, 32>>>(out, cyc, 1.5, 1.0000001, 2000); cudaDeviceSynchronize();
  cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
  printf("dependent-issue latency (cycles): DFMA %lld, rcp (seed + 3 FMA) %lld, MUFU.RCP64H + DADD %lld\n", hc[0], hc[1], hc[2]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
