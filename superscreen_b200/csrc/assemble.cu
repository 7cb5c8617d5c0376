// Fused system assembly (K1-K3, K8-K10, K19): writes M = -A straight into the padded LU
// workspace with one HBM write of 8*n_pad^2 bytes; Q and the dense Laplacian are never
// materialised (reference: distance.py:87-115, device/mesh.py:434-458,
// solver/utils.py:290-297, solver/solve_film.py:181-185,285-305).
#include "scb_common.cuh"

namespace scb {

int nbody_kernel_sum(int64_t m, const double* tgt, int64_t n, const double* src,
                     const int64_t* src_idx, const double* w, const double* v, int64_t ldv,
                     int64_t nrhs, double prefactor, double* out, int accumulate, cudaStream_t s);

__global__ void fill_i32_kernel(int64_t n, int32_t* p, int32_t v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void pos_kernel(int64_t n_int, const int64_t* __restrict__ ix, int32_t* pos) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n_int) pos[ix[r]] = (int32_t)r;
}

// Dense part: tile of TR rows x TC cols per CTA; a thread owns 2 adjacent columns (one 128-bit
// store per row) and TR/4 rows; column data lives in registers, row coordinates in shared
// memory (broadcast reads).  A warp writes 512 contiguous bytes per row.
constexpr int TR = 64;
constexpr int TC = 128;

__global__ void __launch_bounds__(256) assemble_dense_kernel(
    const double* __restrict__ sites, const double* __restrict__ weights,
    const double* __restrict__ qdw, const double* __restrict__ sym_scale, int64_t n_int,
    const int64_t* __restrict__ ix, int64_t n_pad, double* __restrict__ M) {
  __shared__ double rx[TR], ry[TR], rd[TR], rs[TR];
  const int tid = threadIdx.x;
  const int64_t row0 = blockIdx.y * (int64_t)TR;
  const int64_t col0 = blockIdx.x * (int64_t)TC;
  if (tid < TR) {
    const int64_t r = row0 + tid;
    if (r < n_int) {
      const int64_t i = ix[r];
      rx[tid] = sites[2 * i];
      ry[tid] = sites[2 * i + 1];
      rd[tid] = -qdw[i];
      rs[tid] = sym_scale ? sym_scale[i] : 1.0;
    } else {
      rx[tid] = 0.0; ry[tid] = 0.0; rd[tid] = 1.0;  // identity padding
      rs[tid] = 1.0;
    }
  }
  const int cp = tid & 63;   // column pair
  const int rg = tid >> 6;   // row group 0..3
  const int64_t c = col0 + 2 * cp;
  double cx[2], cy[2], cw[2];
  bool cvalid[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    cvalid[k] = (c + k) < n_int;
    if (cvalid[k]) {
      const int64_t j = ix[c + k];
      cx[k] = sites[2 * j];
      cy[k] = sites[2 * j + 1];
      // general: q w_c.  symmetrised (S = W^1/2 (-A) W^-1/2): q sqrt(w_r) sqrt(w_c)
      cw[k] = (sym_scale ? sym_scale[j] : weights[j]) * kOneOver4Pi;
    } else {
      cx[k] = 0.0; cy[k] = 0.0; cw[k] = 0.0;
    }
  }
  __syncthreads();
#pragma unroll 4
  for (int rr = 0; rr < TR / 4; rr++) {
    const int lr = rg * (TR / 4) + rr;
    const int64_t r = row0 + lr;
    const double x = rx[lr], y = ry[lr], sr = rs[lr];
    double2 v;
    double* vv = &v.x;
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const double dx = x - cx[k], dy = y - cy[k];
      const double r2 = dx * dx + dy * dy;
      double val = inv_r3(r2) * (cw[k] * sr);
      // padding rows/cols and the (overwritten) diagonal never see inf: mask them
      val = (r2 > 0.0 && cvalid[k] && r < n_int) ? val : 0.0;
      if (r == c + k) val = rd[lr];
      vv[k] = val;
    }
    *reinterpret_cast<double2*>(M + r * n_pad + c) = v;
  }
}

// Sparse part: one thread per interior row adds Lambda_c*lap_rc + T_rc on the <= ~8 pattern
// entries of that row and evaluates a lower bound of the row-dominance margin.
__global__ void assemble_sparse_kernel(const double* __restrict__ sites,
                                       const double* __restrict__ weights,
                                       const double* __restrict__ qdw,
                                       const double* __restrict__ Lambda,
                                       const int32_t* __restrict__ indptr,
                                       const int32_t* __restrict__ indices,
                                       const double* __restrict__ lap, const double* __restrict__ T,
                                       int64_t n_int, const int64_t* __restrict__ ix,
                                       const int32_t* __restrict__ pos, int64_t n_pad,
                                       double* __restrict__ M, double* __restrict__ margin,
                                       const double* __restrict__ C, const double* __restrict__ sym_scale) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_int) return;
  const int64_t i = ix[r];
  const double xi = sites[2 * i], yi = sites[2 * i + 1];
  double diag = -qdw[i];
  double extra = 0.0;  // sum over sparse off-diagonals of (|M_rc| - q w)
  for (int32_t q = indptr[i]; q < indptr[i + 1]; q++) {
    const int32_t j = indices[q];
    const int32_t c = pos[j];
    if (c < 0) continue;
    const double s = Lambda[j] * lap[q] + (T ? T[q] : 0.0);
    if (c == r) {
      diag += s;
    } else {
      const double dx = xi - sites[2 * j], dy = yi - sites[2 * j + 1];
      const double qw = inv_r3(dx * dx + dy * dy) * (weights[j] * kOneOver4Pi);
      const double val = qw + s;
      // symmetrised storage: entry scaled by sqrt(w_r) / sqrt(w_c)
      M[r * n_pad + c] = sym_scale ? val * (sym_scale[i] / sym_scale[j]) : val;
      extra += fabs(val) - qw;
    }
  }
  M[r * n_pad + r] = diag;
  if (margin) {
    // sum_{c in ix, c != r} q w <= sum_{all j != i} q w = qdw_i - C_i  (lower bound of margin)
    margin[r] = fabs(diag) - ((qdw[i] - C[i]) + extra);
  }
}

__global__ void grad_lambda_kernel(int64_t n, const int32_t* __restrict__ indptr,
                                   const int32_t* __restrict__ indices,
                                   const double* __restrict__ gx, const double* __restrict__ gy,
                                   const double* __restrict__ Lambda, double* __restrict__ T) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double lx = 0.0, ly = 0.0;
  for (int32_t q = indptr[i]; q < indptr[i + 1]; q++) {
    const double l = Lambda[indices[q]];
    lx += gx[q] * l;
    ly += gy[q] * l;
  }
  // einsum("ijk,ijk->jk") over the two gradient components (solve_film.py:183)
  for (int32_t q = indptr[i]; q < indptr[i + 1]; q++) T[q] = lx * gx[q] + ly * gy[q];
}

// out[i,:] (+)= qdw_i v_i - sum_j (Lambda_j lap_ij + T_ij) v_j   (diagonal + sparse part of A@v)
__global__ void apply_sparse_kernel(int64_t n, const double* __restrict__ qdw,
                                    const double* __restrict__ Lambda,
                                    const int32_t* __restrict__ indptr,
                                    const int32_t* __restrict__ indices,
                                    const double* __restrict__ lap, const double* __restrict__ T,
                                    int64_t nrhs, const double* __restrict__ v,
                                    double* __restrict__ out) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * nrhs) return;
  const int64_t i = idx / nrhs, k = idx % nrhs;
  double acc = qdw[i] * v[i * nrhs + k];
  if (Lambda) {
    for (int32_t q = indptr[i]; q < indptr[i + 1]; q++) {
      const int32_t j = indices[q];
      acc -= (Lambda[j] * lap[q] + (T ? T[q] : 0.0)) * v[(int64_t)j * nrhs + k];
    }
  }
  out[idx] += acc;
}

}  // namespace scb

using namespace scb;

extern "C" int scb_grad_lambda_term(int64_t n, const int32_t* op_indptr, const int32_t* op_indices,
                                    const double* gradient_x, const double* gradient_y,
                                    const double* Lambda, double* T, scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0, "n must be positive");
  grad_lambda_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
      n, op_indptr, op_indices, gradient_x, gradient_y, Lambda, T);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int scb_system_assemble(int64_t n, const double* sites, const double* weights,
                                   const double* qdw, const double* C, const double* Lambda,
                                   const int32_t* op_indptr, const int32_t* op_indices,
                                   const double* laplacian, const double* T, int64_t n_int,
                                   const int64_t* ix, int32_t* pos_scratch, int64_t n_pad,
                                   double* negA, double* margin, const double* sym_scale,
                                   scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0 && n_int > 0 && n_int <= n, "bad sizes");
  SCB_CHECK_ARG(sym_scale == nullptr || T == nullptr, "the symmetrised form needs homogeneous Lambda (T == NULL)");
  SCB_CHECK_ARG(n_pad >= n_int && n_pad % SCB_LU_BLOCK == 0, "n_pad must be a multiple of 128 >= n_int");
  SCB_CHECK_ARG(margin == nullptr || C != nullptr, "margin needs C");
  cudaStream_t s = (cudaStream_t)stream;
  fill_i32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(n, pos_scratch, -1);
  SCB_LAUNCH_CHECK();
  pos_kernel<<<(unsigned)ceil_div(n_int, 256), 256, 0, s>>>(n_int, ix, pos_scratch);
  SCB_LAUNCH_CHECK();
  dim3 grid((unsigned)(n_pad / TC), (unsigned)(n_pad / TR));
  assemble_dense_kernel<<<grid, 256, 0, s>>>(sites, weights, qdw, sym_scale, n_int, ix, n_pad, negA);
  SCB_LAUNCH_CHECK();
  assemble_sparse_kernel<<<(unsigned)ceil_div(n_int, 128), 128, 0, s>>>(
      sites, weights, qdw, Lambda, op_indptr, op_indices, laplacian, T, n_int, ix, pos_scratch,
      n_pad, negA, margin, C, sym_scale);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int scb_apply_operator(int64_t n, const double* sites, const double* weights,
                                  const double* qdw, const double* Lambda,
                                  const int32_t* op_indptr, const int32_t* op_indices,
                                  const double* laplacian, const double* T, int64_t n_src,
                                  const int64_t* src_idx, int64_t nrhs, const double* v, double* out,
                                  int accumulate, scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0 && nrhs > 0, "bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t ns = src_idx ? n_src : n;
  // dense part: -sum_{j != i} q_ij w_j v_j
  int rc = nbody_kernel_sum(n, sites, ns, sites, src_idx, weights, v, nrhs, nrhs, -kOneOver4Pi, out,
                            accumulate, s);
  if (rc) return rc;
  apply_sparse_kernel<<<(unsigned)ceil_div(n * nrhs, 256), 256, 0, s>>>(
      n, qdw, Lambda, op_indptr, op_indices, laplacian, T, nrhs, v, out);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}
