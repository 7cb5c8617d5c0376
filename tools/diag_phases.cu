// Phase timing of the diagonal-block kernel (clock64 stamps by thread 0).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Iinclude -Isuperscreen_b200/csrc tools/diag_phases.cu -o /tmp/diag_phases
#include <cstdio>
#include <cstdint>
__device__ long long g_stamps[16];
#define SCB_STAMP(i) do { if (threadIdx.x == 0) g_stamps[i] = clock64(); } while (0)
#include "../superscreen_b200/csrc/getrf.cu"

__global__ void rcp_test(const double* x, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { out[i] = scb::fast_rcp2(x[i]); out[n + i] = scb::fast_rcp(x[i]); double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); out[2 * n + i] = r; }
}

int main() {
  {
    const int m = 1 << 20;
    double* hx = (double*)malloc(sizeof(double) * m); double* ho = (double*)malloc(sizeof(double) * 3 * m);
    unsigned long long st = 88172645463325252ull;
    for (int i = 0; i < m; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; double u = (st >> 11) * (1.0 / 9007199254740992.0); hx[i] = ldexp(1.0 + u, (int)(st % 41) - 20) * ((st >> 5) & 1 ? 1 : -1); }
    double *dx, *dout; cudaMalloc(&dx, sizeof(double) * m); cudaMalloc(&dout, sizeof(double) * 3 * m);
    cudaMemcpy(dx, hx, sizeof(double) * m, cudaMemcpyHostToDevice);
    rcp_test<<<m / 256, 256>>>(dx, dout, m);
    cudaMemcpy(ho, dout, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost);
    double e2 = 0, e3 = 0, e0 = 0;
    for (int i = 0; i < m; i++) { const double ex = 1.0 / hx[i]; e2 = fmax(e2, fabs(ho[i] - ex) / fabs(ex)); e3 = fmax(e3, fabs(ho[m + i] - ex) / fabs(ex)); e0 = fmax(e0, fabs(ho[2 * m + i] - ex) / fabs(ex)); }
    printf("reciprocal max rel error: seed %.3e, 2 Newton steps %.3e, 3 Newton steps %.3e\n", e0, e2, e3);
  }
  const int n = 1024;
  double* M; double* dinv; int32_t* info;
  cudaMalloc(&M, sizeof(double) * n * n);
  cudaMalloc(&dinv, 2 * 128 * 128 * sizeof(double));
  cudaMalloc(&info, 4);
  double* h = (double*)malloc(sizeof(double) * n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) h[i * n + j] = (i == j) ? 200.0 : 1.0 / (1.0 + abs(i - j));
  cudaMemcpy(M, h, sizeof(double) * n * n, cudaMemcpyHostToDevice);
  const int smem = 3 * scb::QN * scb::QLD * sizeof(double);
  cudaFuncSetAttribute(scb::diag_kernel_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(scb::diag_kernel_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int smem_b = (3 * scb::QN * scb::QLD + 32 * scb::TLD) * sizeof(double);
  cudaFuncSetAttribute(scb::diag_kernel_symb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 9; rep++) {
    cudaMemcpy(M, h, sizeof(double) * n * n, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    if (rep < 3)
      scb::diag_kernel_small<false><<<1, 256, smem>>>(M, n, 0, dinv, dinv + 128 * 128, info, 0);
    else if (rep < 6)
      scb::diag_kernel_small<true><<<1, 256, smem>>>(M, n, 0, dinv, dinv + 128 * 128, info, 0);
    else
      scb::diag_kernel_symb<<<1, 256, smem_b>>>(M, n, 0, dinv, dinv + 128 * 128, info, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long st[16];
    cudaMemcpyFromSymbol(st, g_stamps, sizeof(st));
    printf("%s rep %d: %.1f us total (events);", rep < 3 ? "general" : (rep < 6 ? "symmetric sweep" : "symmetric blocked"), rep, ms * 1e3);
    const char* names[] = {"sweep A11", "emit+store", "U12,L21 gemms", "A22 update", "sweep A22", "emit+store", "inverse coupling"};
    for (int k = 0; k < 7; k++) printf(" %s=%lld", names[k], st[k + 1] - st[k]);
    printf(" cycles; stamped span %lld\n", st[7] - st[0]);
    if (rep >= 6) printf("   blocked kb=0 (2nd quadrant): A=%lld sync=%lld B=%lld sync=%lld C=%lld sync=%lld ; all 8 steps=%lld doubling=%lld\n",
                         st[9] - st[8], st[10] - st[9], st[11] - st[10], st[12] - st[11], st[13] - st[12], st[14] - st[13], st[15] - st[8], st[5] - st[15]);
  }
  {  // check the last (blocked symmetric) result against a host LU without pivoting of the 128x128 block
    static double lu[128][128], got[128 * 1024], il[128 * 128], iu[128 * 128];
    for (int i = 0; i < 128; i++) for (int j = 0; j < 128; j++) lu[i][j] = h[i * n + j];
    for (int k = 0; k < 128; k++) for (int i = k + 1; i < 128; i++) {
      lu[i][k] /= lu[k][k];
      for (int j = k + 1; j < 128; j++) lu[i][j] -= lu[i][k] * lu[k][j];
    }
    cudaMemcpy(got, M, sizeof(double) * 128 * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(il, dinv, sizeof(il), cudaMemcpyDeviceToHost);
    cudaMemcpy(iu, dinv + 128 * 128, sizeof(iu), cudaMemcpyDeviceToHost);
    double e = 0, eil = 0, eiu = 0;
    for (int i = 0; i < 128; i++) for (int j = 0; j < 128; j++) e = fmax(e, fabs(got[i * n + j] - lu[i][j]));
    // inv(L) L = I, U inv(U) = I
    for (int i = 0; i < 128; i++) for (int j = 0; j < 128; j++) {
      double sl = 0, su = 0;
      for (int k = 0; k < 128; k++) {
        const double Lkj = k > j ? lu[k][j] : (k == j ? 1.0 : 0.0);
        const double Uik = k >= i ? lu[i][k] : 0.0;
        sl += il[i * 128 + k] * Lkj;
        su += Uik * iu[k * 128 + j];
      }
      eil = fmax(eil, fabs(sl - (i == j))); eiu = fmax(eiu, fabs(su - (i == j)));
    }
    printf("blocked symmetric vs host LU: max |LU diff| %.3e  |inv(L) L - I| %.3e  |U inv(U) - I| %.3e\n", e, eil, eiu);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
