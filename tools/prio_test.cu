// Does a high-priority kernel whose single CTA needs a WHOLE SM get scheduled while a low-priority
// kernel with half-SM CTAs keeps every SM busy?  (decides how LU look-ahead can be organised)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(256, 2) busy_small(unsigned long long ns, unsigned long long* first_start) {
  extern __shared__ double sm[];
  unsigned long long t0 = gtime();
  if (blockIdx.x == 0 && threadIdx.x == 0) *first_start = t0;
  ns = ns / 2 + (ns * ((blockIdx.x * 2654435761u) % 1024u)) / 1024u;  // desynchronise CTAs: 0.5..1.5 x ns
  while (gtime() - t0 < ns) { sm[threadIdx.x] += 1.0; }
}
__global__ void __launch_bounds__(512, 1) busy_big(unsigned long long ns, unsigned long long* times) {
  extern __shared__ double sm[];
  unsigned long long t0 = gtime();
  while (gtime() - t0 < ns) { sm[threadIdx.x] += 1.0; }
  if (threadIdx.x == 0) { times[0] = t0; times[1] = gtime(); }
}
__global__ void __launch_bounds__(256, 2) busy_half(unsigned long long ns, unsigned long long* times) {
  extern __shared__ double sm[];
  unsigned long long t0 = gtime();
  while (gtime() - t0 < ns) { sm[threadIdx.x] += 1.0; }
  if (threadIdx.x == 0) { times[0] = t0; times[1] = gtime(); }
}
int main() {
  int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  printf("priority range: lowest %d highest %d\n", lo, hi);
  cudaStream_t s_lo, s_hi; CK(cudaStreamCreateWithPriority(&s_lo, cudaStreamNonBlocking, lo)); CK(cudaStreamCreateWithPriority(&s_hi, cudaStreamNonBlocking, hi));
  unsigned long long *d; CK(cudaMalloc(&d, 8 * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(busy_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CK(cudaFuncSetAttribute(busy_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
  CK(cudaFuncSetAttribute(busy_half, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int variant = 0; variant < 2; variant++) {
    CK(cudaMemset(d, 0, 64));
    // low priority: 148*2*40 CTAs x 20 us -> ~800 us
    busy_small<<<148 * 2 * 40, 256, 96 * 1024, s_lo>>>(20000ull, d + 4);
    // give it time to fill the GPU
    cudaEvent_t e; cudaEventCreate(&e);
    // host sleep ~100us
    for (volatile int i = 0; i < 200000; i++) {}
    unsigned long long *hp; cudaHostAlloc(&hp, 8, cudaHostAllocDefault);
    if (variant == 0) busy_big<<<1, 512, 132 * 1024, s_hi>>>(80000ull, d);
    else busy_half<<<1, 256, 100 * 1024, s_hi>>>(80000ull, d);
    CK(cudaDeviceSynchronize());
    unsigned long long h[8]; CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
    printf("%s CTA: started %.1f us after the low-priority kernel started, ran %.1f us\n", variant == 0 ? "full-SM (512 thr, 132 KB)" : "half-SM (256 thr, 100 KB)",
           (double)(h[0] - h[4]) * 1e-3, (double)(h[1] - h[0]) * 1e-3);
  }
  return 0;
}
