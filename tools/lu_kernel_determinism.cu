// Determinism harness for the LU kernels: runs each kernel several times on IDENTICAL inputs and
// compares the outputs bit by bit (a kernel whose CTAs are independent must reproduce itself).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/lu_kernel_determinism tools/lu_kernel_determinism.cu
//   tools/lu_kernel_determinism [n_pad = 19968] [reps = 4]
#include "../superscreen_b200/csrc/api.cu"
#include "../superscreen_b200/csrc/getrf.cu"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));   \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__global__ void fill_kernel(double* p, int64_t n, uint64_t seed, double scale, double shift) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    p[i] = shift + scale * ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5);
  }
}
__global__ void diag_boost_kernel(double* M, int64_t n, double v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) M[i * n + i] = v;
}
// counts differing 8-byte words; remembers the smallest differing index
__global__ void compare_kernel(const unsigned long long* a, const unsigned long long* b, int64_t n,
                               unsigned long long* count, unsigned long long* first) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (a[i] != b[i]) {
      atomicAdd(count, 1ull);
      atomicMin(first, (unsigned long long)i);
    }
}
static unsigned long long* g_cmp = nullptr;
static void compare(const char* what, const double* a, const double* b, int64_t n, int64_t ld) {
  if (!g_cmp) CK(cudaMalloc(&g_cmp, 16));
  unsigned long long init[2] = {0ull, ~0ull}, res[2];
  CK(cudaMemcpy(g_cmp, init, 16, cudaMemcpyHostToDevice));
  compare_kernel<<<1024, 256>>>((const unsigned long long*)a, (const unsigned long long*)b, n, g_cmp, g_cmp + 1);
  CK(cudaMemcpy(res, g_cmp, 16, cudaMemcpyDeviceToHost));
  if (res[0] == 0)
    printf("    %-28s identical\n", what);
  else if (ld > 0)
    printf("    %-28s %llu words differ, first at row %lld col %lld (block %lld, %lld)\n", what, res[0],
           (long long)(res[1] / ld), (long long)(res[1] % ld), (long long)(res[1] / ld / 128), (long long)(res[1] % ld / 128));
  else
    printf("    %-28s %llu words differ, first at %llu\n", what, res[0], res[1]);
}

int main(int argc, char** argv) {
  using namespace scb;
  const int64_t n = argc > 1 ? atoll(argv[1]) : 19968;
  const int reps = argc > 2 ? atoi(argv[2]) : 4;
  const int64_t nb = n / NB;
  const int q = 8, tile_chunks = q * NCHUNK;
  const int64_t pack_elems = n * NB * q;
  double *M0, *M, *Mref, *Lp, *Up, *Lp2, *Up2, *inv;
  CK(cudaMalloc(&M0, n * n * 8));
  CK(cudaMalloc(&M, n * n * 8));
  CK(cudaMalloc(&Mref, n * n * 8));
  CK(cudaMalloc(&Lp, pack_elems * 8));
  CK(cudaMalloc(&Up, pack_elems * 8));
  CK(cudaMalloc(&Lp2, pack_elems * 8));
  CK(cudaMalloc(&Up2, pack_elems * 8));
  CK(cudaMalloc(&inv, 2 * NB * NB * 8));
  fill_kernel<<<2048, 256>>>(M0, n * n, 1, 1.0, 0.0);
  fill_kernel<<<2048, 256>>>(Lp, pack_elems, 2, 1e-2, 0.0);
  fill_kernel<<<2048, 256>>>(Up, pack_elems, 3, 1e-2, 0.0);
  CK(cudaDeviceSynchronize());
  const int upd_smem = 2 * sizeof(UpdateStage);
  const int trsm_smem = kTrsmSmemDoubles * sizeof(double);
  CK(cudaFuncSetAttribute(update_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem));
  CK(cudaFuncSetAttribute(update_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem));
  CK(cudaFuncSetAttribute(update_kernel_t<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(update_kernel_t<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(trsm_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem));
  CK(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, trsm_smem));

  // ---- A. trailing update, triangular and rectangular, K = 1024 and K = 128, both raster orders ----
  for (int tri = 1; tri >= 0; tri--)
    for (int nchunks : {32, 4})
      for (int band : {0, 16}) {
        printf("update_kernel_t<%d> grid (%lld, %lld) nchunks %d band %d\n", tri, (long long)(2 * nb), (long long)nb, nchunks, band);
        for (int rep = 0; rep <= reps; rep++) {
          double* out = rep == 0 ? Mref : M;
          CK(cudaMemcpy(out, M0, n * n * 8, cudaMemcpyDeviceToDevice));
          dim3 g((unsigned)(2 * nb), (unsigned)nb);
          if (tri)
            update_kernel_t<true><<<g, 256, upd_smem>>>(out, n, 0, 0, Lp, Up, tile_chunks, 0, nchunks, band);
          else
            update_kernel_t<false><<<g, 256, upd_smem>>>(out, n, 0, 0, Lp, Up, tile_chunks, 0, nchunks, band);
          CK(cudaGetLastError());
          CK(cudaDeviceSynchronize());
          if (rep > 0) compare("C", Mref, M, n * n, n);
        }
      }

  // ---- B. symmetric panel solve: every 64-row tile below the first diagonal block ----
  {
    fill_kernel<<<64, 256>>>(inv, 2 * NB * NB, 5, 1e-1, 0.0);
    diag_boost_kernel<<<(unsigned)((n + 255) / 256), 256>>>(M0, n, 3.0);
    CK(cudaDeviceSynchronize());
    const int nt = (int)(nb - 1);
    printf("trsm_sym_kernel grid %d\n", 2 * nt);
    for (int rep = 0; rep <= reps; rep++) {
      double* out = rep == 0 ? Mref : M;
      double* lp = rep == 0 ? Lp2 : Lp;
      double* up = rep == 0 ? Up2 : Up;
      CK(cudaMemcpy(out, M0, n * n * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemset(lp, 0, pack_elems * 8));
      CK(cudaMemset(up, 0, pack_elems * 8));
      trsm_sym_kernel<<<2 * nt, 256, trsm_smem>>>(out, n, 0, inv + NB * NB, lp, up, tile_chunks, 0, 0);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      if (rep > 0) {
        compare("M (L21 in place, U12^T)", Mref, M, n * n, n);
        compare("Lpack", Lp2, Lp, pack_elems, 0);
        compare("Upack", Up2, Up, pack_elems, 0);
      }
    }
    printf("trsm_kernel (general) grid %d\n", 4 * nt);
    for (int rep = 0; rep <= reps; rep++) {
      double* out = rep == 0 ? Mref : M;
      double* lp = rep == 0 ? Lp2 : Lp;
      double* up = rep == 0 ? Up2 : Up;
      CK(cudaMemcpy(out, M0, n * n * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemset(lp, 0, pack_elems * 8));
      CK(cudaMemset(up, 0, pack_elems * 8));
      trsm_kernel<<<4 * nt, 256, trsm_smem>>>(out, n, 0, 2 * nt, inv, inv + NB * NB, lp, up, tile_chunks, 0);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      if (rep > 0) {
        compare("M (L21, U12 in place)", Mref, M, n * n, n);
        compare("Lpack", Lp2, Lp, pack_elems, 0);
        compare("Upack", Up2, Up, pack_elems, 0);
      }
    }
  }
  printf("done\n");
  return 0;
}
