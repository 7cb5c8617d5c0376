// Dependent-issue latencies (cycles) of the instructions on the LU pivot chain, one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/latency_fp64.cu -o build/latency_fp64
#include <cstdio>
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
__global__ void lat(double* out, long long* cyc, double x0, int iters) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  // DFMA
  t0 = clk();
  for (int i = 0; i < iters; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(y), "d"(1e-9));
  t1 = clk(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DMUL
  t0 = clk();
  for (int i = 0; i < iters; i++) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(y));
  t1 = clk(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // SHFL (64-bit = 2 x SHFL)
  t0 = clk();
  for (int i = 0; i < iters; i++) { x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); asm volatile("" : "+d"(x)); }
  t1 = clk(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // MUFU.RCP64H seed only
  t0 = clk();
  for (int i = 0; i < iters; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  t1 = clk(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // seed + 2 Newton
  x = x0;
  t0 = clk();
  for (int i = 0; i < iters; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); x = fma(r, e, r); }
  t1 = clk(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // shared memory load -> use
  __shared__ double sm[64];
  if (threadIdx.x < 32) { sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y; }
  __syncthreads();
  int idx = threadIdx.x & 31;
  t0 = clk();
  for (int i = 0; i < iters; i++) { double v = sm[idx]; idx = (idx + (v > 1e300 ? 1 : 0) + 1) & 63; }
  t1 = clk(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // DMMA dependent accumulate
  double c0 = 0, c1 = 0;
  t0 = clk();
  for (int i = 0; i < iters; i++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(x), "d"(y));
  t1 = clk(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // __syncthreads (256 threads launched in second config)
  t0 = clk();
  for (int i = 0; i < iters; i++) __syncthreads();
  t1 = clk(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = x + c0 + c1 + idx;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 64);
  const int iters = 4096;
  for (int threads : {32, 256}) {
    lat<<<1, threads>>>(out, cyc, 1.5, iters); cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, cyc, 1.5, iters); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("%d threads: DFMA %.1f  DMUL %.1f  SHFL64 %.1f  MUFU.RCP64H %.1f  rcp(seed+2NR) %.1f  LDS->use %.1f  DMMA(dep) %.1f  BAR %.1f cycles\n", threads,
           h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters, h[5] / (double)iters, h[6] / (double)iters, h[7] / (double)iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
