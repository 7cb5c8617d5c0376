// Measures the fp64 roofline denominators on the box: DMMA (mma.sync m8n8k4 f64) issue rate,
// DFMA issue rate, cuBLAS DGEMM and cuSOLVER DGETRF (comparators only, never on the product path).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_fp64 tools/peak_fp64.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
// every kernel launch is checked: a configuration that cannot launch (too many registers for the
// block size) must not be reported as a throughput
#define LAUNCH_OK() CK(cudaGetLastError())

template <int NACC>
__global__ void dmma_kernel(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {4, 8, 16}) {  // (32 warps x 16 accumulator pairs do not fit the register file)
    int iters = 20000;
    dmma_kernel<16><<<nsm, warps * 32>>>(out, 100);
    LAUNCH_OK();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0); dmma_kernel<16><<<nsm, warps * 32>>>(out, iters); LAUNCH_OK(); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * warps * nsm;
    printf("DMMA m8n8k4 warps/SM=%d: %.2f TFLOP/s (%.3f ms)\n", warps, flops / best * 1e-9, best);
  }
  for (int warps : {8, 16, 32}) {
    int iters = 20000;
    dfma_kernel<16><<<nsm, warps * 32>>>(out, 100);
    LAUNCH_OK();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
      cudaEventRecord(e0); dfma_kernel<16><<<nsm, warps * 32>>>(out, iters); LAUNCH_OK(); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 2.0 * 32 * 16.0 * iters * warps * nsm;
    printf("DFMA warps/SM=%d: %.2f TFLOP/s (%.3f ms)\n", warps, flops / best * 1e-9, best);
  }
  // sustained DMMA for ~2 s (power-capped clocks)
  {
    cudaEventRecord(e0);
    int launches = 0;
    for (; launches < 40; launches++) { dmma_kernel<16><<<nsm, 512>>>(out, 200000); LAUNCH_OK(); }
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 16.0 * 200000 * 16 * nsm * launches;
    printf("DMMA sustained (%.0f ms): %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  // cuBLAS DGEMM
  {
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192, 16384}) {
      double *A, *B, *C; size_t bytes = sizeof(double) * n * (size_t)n;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
      CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
      double one = 1.0, zero = 0.0;
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n);
      CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("cuBLAS DGEMM n=%d: %.2f TFLOP/s (%.3f ms)\n", n, 2.0 * n * (double)n * n / best * 1e-9, best);
      // rank-128 update shape (the LU trailing update): C[n x n] -= A[n x 128] B[128 x n]
      double mone = -1.0;
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, 128, &mone, A, n, B, n, &one, C, n);
      CK(cudaDeviceSynchronize());
      best = 1e30f;
      for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, 128, &mone, A, n, B, n, &one, C, n); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      printf("cuBLAS DGEMM rank-128 update n=%d: %.2f TFLOP/s (%.3f ms)\n", n, 2.0 * n * (double)n * 128 / best * 1e-9, best);
      cudaFree(A); cudaFree(B); cudaFree(C);
    }
    cublasDestroy(h);
  }
  // cuSOLVER DGETRF comparator at the C2 size
  {
    cusolverDnHandle_t h; cusolverDnCreate(&h);
    for (int n : {9856, 19712}) {
      double* A; size_t bytes = sizeof(double) * n * (size_t)n; CK(cudaMalloc(&A, bytes));
      std::vector<double> hA((size_t)n * 64);
      CK(cudaMemset(A, 0, bytes));
      // diagonally dominant matrix: small off-diagonals from a cheap pattern, big diagonal
      std::vector<double> col(n);
      for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) col[i] = 1.0 / (1.0 + ((i * 131 + j * 71) % 1000));
        col[j] = 2.0 * n;
        CK(cudaMemcpy(A + (size_t)j * n, col.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
      }
      int lwork; cusolverDnDgetrf_bufferSize(h, n, n, A, n, &lwork);
      double* work; CK(cudaMalloc(&work, sizeof(double) * lwork));
      int *ipiv, *info; CK(cudaMalloc(&ipiv, sizeof(int) * n)); CK(cudaMalloc(&info, sizeof(int)));
      for (int piv = 1; piv >= 0; piv--) {
        cudaEventRecord(e0); cusolverDnDgetrf(h, n, n, A, n, work, piv ? ipiv : nullptr, info); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("cuSOLVER DGETRF n=%d pivot=%d: %.2f TFLOP/s (%.1f ms) [matrix is already LU on 2nd pass]\n", n, piv, 2.0 / 3.0 * n * (double)n * n / ms * 1e-9, ms);
      }
      cudaFree(A); cudaFree(work); cudaFree(ipiv); cudaFree(info);
    }
    cusolverDnDestroy(h);
  }
  return 0;
}
