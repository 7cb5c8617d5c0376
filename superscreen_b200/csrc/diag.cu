// Roofline denominators measured in place (bench.py): the issue rate of the fp64 tensor-core atom
// DMMA.8x8x4 (mma.sync m8n8k4 f64) and of DFMA, as register-resident dependent chains long enough to
// hide the pipe latency.  Diagnostics only -- nothing on the solve path calls these.
#include "scb_common.cuh"

namespace scb {

template <int NACC>
__global__ void __launch_bounds__(512) dmma_issue_kernel(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(512) dfma_issue_kernel(double* out, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * (int64_t)blockDim.x + threadIdx.x] = s;
}

}  // namespace scb

using namespace scb;

// kind 0: DMMA, kind 1: DFMA.  One CTA of 512 threads per SM, 16 independent accumulator chains per
// thread.  *flop receives the floating-point operations the launch executes; scratch: double[sms * 512].
extern "C" int scb_diag_issue_rate(int kind, int64_t iters, double* scratch, double* flop_host,
                                   scb_stream_t stream) {
  SCB_CHECK_ARG(kind == 0 || kind == 1, "kind must be 0 (DMMA) or 1 (DFMA)");
  SCB_CHECK_ARG(iters > 0 && iters < (1ll << 30), "bad iteration count");
  int dev = 0, sms = 0;
  SCB_CUDA(cudaGetDevice(&dev));
  SCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t s = (cudaStream_t)stream;
  const double warps = 16.0 * sms;
  if (kind == 0) {
    dmma_issue_kernel<16><<<sms, 512, 0, s>>>(scratch, (int)iters);
    if (flop_host) *flop_host = 2.0 * 8 * 8 * 4 * 16.0 * (double)iters * warps;
  } else {
    dfma_issue_kernel<16><<<sms, 512, 0, s>>>(scratch, (int)iters);
    if (flop_host) *flop_host = 2.0 * 32 * 16.0 * (double)iters * warps;
  }
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int64_t scb_diag_scratch_elems(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return (int64_t)sms * 512;
}
