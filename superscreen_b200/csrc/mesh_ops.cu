// Mesh topology + FEM operator assembly on the device (rows a4-a8 of SURVEY.md section 8).
//
// Everything is built from one integer structure, the vertex star (vertex -> incident
// (triangle, local index) pairs, ascending by triangle id), obtained by a counting sort on
// integer keys.  All float sums run over a star or a CSR row in a fixed order, so results are
// run-to-run deterministic (no float atomics).  Integer atomics are used only for counting /
// slot assignment, and every slot-ordered segment is sorted afterwards.
//
// Reference semantics restated here (file:line under /root/reference/superscreen):
//   device/utils.py:139-152 get_edges        device/utils.py:230-273 triangle/vertex areas
//   device/mesh.py:157-170 boundary indices  device/mesh.py:400-432 C_vector
//   device/edge_mesh.py:38-63                fem.py:70-121 adjacency, directed-edge map
//   fem.py:124-296 weights + laplacian       fem.py:299-402 gradients
#include <stdarg.h>

#include "scb_common.cuh"

namespace scb {

// ---- workspace layout (int32 elements) -------------------------------------------------
struct MeshWs {
  int32_t* v2t_ptr;    // [n+1] star offsets
  int32_t* v2t_ent;    // [3m]  tri*4 + local index, ascending per vertex
  int32_t* cursor;     // [n]
  int32_t* nbr;        // [6m]  neighbour multiset per vertex (2 per star entry), sorted
  int32_t* adj_ptr;    // [n+1] unique-neighbour offsets
  int32_t* edge_ptr;   // [n+1] offsets of edges (i<j) owned by vertex i
  int32_t* bvert_ptr;  // [n+1] boundary-vertex offsets
  int32_t* scan_tmp;   // [1024]
};

__host__ __device__ inline MeshWs carve(int32_t* ws, int64_t n, int64_t m) {
  MeshWs w;
  int32_t* p = ws;
  w.v2t_ptr = p;   p += n + 1;
  w.v2t_ent = p;   p += 3 * m;
  w.cursor = p;    p += n;
  w.nbr = p;       p += 6 * m;
  w.adj_ptr = p;   p += n + 1;
  w.edge_ptr = p;  p += n + 1;
  w.bvert_ptr = p; p += n + 1;
  w.scan_tmp = p;  p += 1024;
  return w;
}

// ---- single-block exclusive scan (n is O(1e4..1e6); latency-trivial) -------------------
__global__ void scan_exclusive_kernel(int32_t* data, int64_t count) {
  __shared__ int32_t partial[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (count + blockDim.x - 1) / blockDim.x;
  const int64_t lo = t * chunk;
  const int64_t hi = lo + chunk < count ? lo + chunk : count;
  int32_t s = 0;
  for (int64_t i = lo; i < hi; i++) s += data[i];
  partial[t] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan on the 1024 partials
  for (int off = 1; off < blockDim.x; off <<= 1) {
    int32_t v = (t >= off) ? partial[t - off] : 0;
    __syncthreads();
    partial[t] += v;
    __syncthreads();
  }
  int32_t run = partial[t] - s;
  for (int64_t i = lo; i < hi; i++) {
    int32_t v = data[i];
    data[i] = run;
    run += v;
  }
}

__global__ void count_star_kernel(int64_t m, const int64_t* __restrict__ el, int32_t* cnt) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * m) return;
  atomicAdd(&cnt[el[i]], 1);
}

__global__ void fill_star_kernel(int64_t m, const int64_t* __restrict__ el, const int32_t* ptr,
                                 int32_t* cursor, int32_t* ent) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * m) return;
  int64_t v = el[i];
  int32_t slot = atomicAdd(&cursor[v], 1);
  int32_t t = (int32_t)(i / 3), k = (int32_t)(i % 3);
  ent[ptr[v] + slot] = t * 4 + k;
}

__device__ inline void insertion_sort(int32_t* a, int len) {
  for (int i = 1; i < len; i++) {
    int32_t x = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > x) {
      a[j + 1] = a[j];
      j--;
    }
    a[j + 1] = x;
  }
}

// per vertex: sort star, build sorted neighbour multiset, count structures
__global__ void star_analyze_kernel(int64_t n, const int64_t* __restrict__ el, MeshWs w,
                                    int64_t* counts, int32_t* flags) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int32_t s0 = w.v2t_ptr[v], s1 = w.v2t_ptr[v + 1];
  const int deg = s1 - s0;
  int32_t* ent = w.v2t_ent + s0;
  insertion_sort(ent, deg);
  int32_t* nb = w.nbr + 2 * (int64_t)s0;
  for (int e = 0; e < deg; e++) {
    int32_t t = ent[e] >> 2, k = ent[e] & 3;
    nb[2 * e] = (int32_t)el[3 * (int64_t)t + (k + 1) % 3];      // head of directed edge v->.
    nb[2 * e + 1] = (int32_t)el[3 * (int64_t)t + (k + 2) % 3];  // tail of directed edge .->v
  }
  // duplicated directed edge (inconsistent orientation) check: heads must be distinct
  int bad = 0;
  for (int a = 0; a < deg; a++)
    for (int b = a + 1; b < deg; b++)
      if (nb[2 * a] == nb[2 * b] || nb[2 * a + 1] == nb[2 * b + 1]) bad = 1;
  insertion_sort(nb, 2 * deg);
  int uniq = 0, upper = 0, bflag = 0, bedges = 0;
  for (int e = 0; e < 2 * deg;) {
    int f = e + 1;
    while (f < 2 * deg && nb[f] == nb[e]) f++;
    int mult = f - e;
    uniq++;
    if (nb[e] == (int32_t)v) bad = 1;  // degenerate triangle
    if (mult > 2) bad = 1;
    if (mult == 1) bflag = 1;
    if (nb[e] > (int32_t)v) {
      upper++;
      if (mult == 1) bedges++;
    }
    e = f;
  }
  w.adj_ptr[v] = uniq;
  w.edge_ptr[v] = upper;
  w.bvert_ptr[v] = bflag;
  if (bedges) atomicAdd((unsigned long long*)&counts[3], (unsigned long long)bedges);
  if (bad) atomicOr(flags, 1);
  if (v == 0) {
    w.adj_ptr[n] = 0;
    w.edge_ptr[n] = 0;
    w.bvert_ptr[n] = 0;
  }
}

__global__ void totals_kernel(int64_t n, MeshWs w, int64_t* counts) {
  counts[0] = w.adj_ptr[n];
  counts[1] = w.edge_ptr[n];
  counts[2] = w.bvert_ptr[n];
}

// ---- build phase -----------------------------------------------------------------------
struct P2 {
  double x, y;
};
__device__ __forceinline__ P2 ldp(const double* __restrict__ s, int64_t i) {
  const double2 v = *reinterpret_cast<const double2*>(s + 2 * i);
  return P2{v.x, v.y};
}

__device__ __forceinline__ double tri_area(P2 p0, P2 p1, P2 p2) {
  // s = [p2-p1, p0-p2]; area = det(s)/2   (device/utils.py:240-248)
  double s00 = p2.x - p1.x, s01 = p2.y - p1.y;
  double s10 = p0.x - p2.x, s11 = p0.y - p2.y;
  return 0.5 * (s00 * s11 - s01 * s10);
}

__global__ void triangle_kernel(int64_t m, const double* __restrict__ sites,
                                const int64_t* __restrict__ el, scb_mesh_out o) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  int64_t v[3] = {el[3 * t], el[3 * t + 1], el[3 * t + 2]};
  P2 p[3] = {ldp(sites, v[0]), ldp(sites, v[1]), ldp(sites, v[2])};
  double a = tri_area(p[0], p[1], p[2]);
  o.triangle_areas[t] = a;
  o.centroids[2 * t] = ((p[0].x + p[1].x) + p[2].x) / 3.0;
  o.centroids[2 * t + 1] = ((p[0].y + p[1].y) + p[2].y) / 3.0;
  // gradient_triangles (fem.py:322-346): e_k = p_{k+1} - p_{k+2}; Gx = e_y/(2a), Gy = -e_x/(2a)
  double gx[3], gy[3];
  const double two_a = 2.0 * a;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    P2 pa = p[(k + 1) % 3], pb = p[(k + 2) % 3];
    gx[k] = (pa.y - pb.y) / two_a;
    gy[k] = -(pa.x - pb.x) / two_a;
  }
  // sort the three columns ascending (scipy canonical CSR)
  int ord[3] = {0, 1, 2};
#pragma unroll
  for (int a2 = 0; a2 < 2; a2++)
#pragma unroll
    for (int b = 0; b < 2 - a2; b++)
      if (v[ord[b]] > v[ord[b + 1]]) {
        int tmp = ord[b];
        ord[b] = ord[b + 1];
        ord[b + 1] = tmp;
      }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.gtri_indices[3 * t + k] = (int32_t)v[ord[k]];
    o.gtri_x[3 * t + k] = gx[ord[k]];
    o.gtri_y[3 * t + k] = gy[ord[k]];
  }
}

// angle helper following the reference's arccos(dot / (|v1||v2|)) formulation
__device__ __forceinline__ double angle_between(double x1, double y1, double x2, double y2) {
  double dot = __dadd_rn(__dmul_rn(x1, x2), __dmul_rn(y1, y2));
  double n1 = sqrt(__dadd_rn(__dmul_rn(x1, x1), __dmul_rn(y1, y1)));
  double n2 = sqrt(__dadd_rn(__dmul_rn(x2, x2), __dmul_rn(y2, y2)));
  return acos(dot / (n1 * n2));
}

// per vertex: areas, adjacency, edges, boundary, star (head order), operator pattern, C is separate
__global__ void vertex_structure_kernel(int64_t n, const double* __restrict__ sites,
                                        const int64_t* __restrict__ el, MeshWs w, scb_mesh_out o) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int32_t s0 = w.v2t_ptr[v], s1 = w.v2t_ptr[v + 1];
  const int deg = s1 - s0;
  const int32_t* ent = w.v2t_ent + s0;
  // vertex area: triangles in ascending id order == the reference loop order (device/utils.py:268-273)
  double area = 0.0;
  for (int e = 0; e < deg; e++) area += o.triangle_areas[ent[e] >> 2] / 3.0;
  o.vertex_areas[v] = area;
  // star ordered by head vertex (fem.py:386-391 adj_tri.data[i])
  {
    int32_t* heads = o.star_heads + s0;
    int32_t* tris = o.star_tris + s0;
    for (int e = 0; e < deg; e++) {
      int32_t t = ent[e] >> 2, k = ent[e] & 3;
      int32_t h = (int32_t)el[3 * (int64_t)t + (k + 1) % 3];
      int j = e - 1;
      while (j >= 0 && heads[j] > h) {
        heads[j + 1] = heads[j];
        tris[j + 1] = tris[j];
        j--;
      }
      heads[j + 1] = h;
      tris[j + 1] = t;
    }
    o.star_indptr[v] = s0;
    if (v == n - 1) o.star_indptr[n] = s1;
  }
  // adjacency / edges / operator pattern from the sorted neighbour multiset
  const int32_t* nb = w.nbr + 2 * (int64_t)s0;
  int32_t ap = w.adj_ptr[v], ep = w.edge_ptr[v];
  int32_t op = ap + (int32_t)v;  // pattern adjacency + I: row v starts at adj_ptr[v] + v
  o.adj_indptr[v] = ap;
  o.op_indptr[v] = op;
  if (v == n - 1) {
    o.adj_indptr[n] = w.adj_ptr[n];
    o.op_indptr[n] = w.adj_ptr[n] + (int32_t)n;
  }
  const P2 pv = ldp(sites, v);
  bool self_done = false;
  for (int e = 0; e < 2 * deg;) {
    int f = e + 1;
    while (f < 2 * deg && nb[f] == nb[e]) f++;
    const int32_t j = nb[e];
    o.adj_indices[ap++] = j;
    if (!self_done && j > (int32_t)v) {
      o.op_indices[op++] = (int32_t)v;
      self_done = true;
    }
    o.op_indices[op++] = j;
    if (j > (int32_t)v) {
      o.edges[2 * (int64_t)ep] = v;
      o.edges[2 * (int64_t)ep + 1] = j;
      o.edge_is_boundary[ep] = (f - e) == 1 ? 1 : 0;
      const P2 pj = ldp(sites, j);
      // edge_mesh.py:49-56: centers = mean of the two endpoints, directions = p[j]-p[i]
      o.edge_centers[2 * (int64_t)ep] = (pv.x + pj.x) / 2.0;
      o.edge_centers[2 * (int64_t)ep + 1] = (pv.y + pj.y) / 2.0;
      double dx = pj.x - pv.x, dy = pj.y - pv.y;
      o.edge_directions[2 * (int64_t)ep] = dx;
      o.edge_directions[2 * (int64_t)ep + 1] = dy;
      o.edge_lengths[ep] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      ep++;
    }
    e = f;
  }
  if (!self_done) o.op_indices[op++] = (int32_t)v;
  if (w.bvert_ptr[v + 1] > w.bvert_ptr[v]) o.boundary_indices[w.bvert_ptr[v]] = v;
}

// Laplacian row (fem.py:259-296) and vertex-gradient row (fem.py:350-402) for vertex v.
__global__ void vertex_operator_kernel(int64_t n, const double* __restrict__ sites,
                                       const int64_t* __restrict__ el, MeshWs w, int method,
                                       scb_mesh_out o) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int32_t s0 = w.v2t_ptr[v], s1 = w.v2t_ptr[v + 1];
  const int deg = s1 - s0;
  const int32_t* ent = w.v2t_ent + s0;
  const int32_t r0 = o.op_indptr[v], r1 = o.op_indptr[v + 1];
  const P2 pv = ldp(sites, v);

  // ---------------- Laplacian ----------------
  double rowsum = 0.0;
  int diag_pos = -1;
  for (int32_t q = r0; q < r1; q++) {
    const int32_t j = o.op_indices[q];
    if (j == (int32_t)v) {
      diag_pos = q;
      continue;
    }
    double wij = 0.0;
    if (method == 1) {
      wij = 1.0;  // uniform: adjacency (fem.py:246-248)
    } else if (method == 2) {
      // inv_euclidean (fem.py:148-160): 1/|p_a - p_b| with a < b in local-vertex order of the
      // assigning triangle; direction does not matter for the norm.
      const P2 pj = ldp(sites, j);
      double dx = pj.x - pv.x, dy = pj.y - pv.y;
      wij = 1.0 / sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    } else {
      // half cotangent (fem.py:188-222): for every triangle containing edge (v, j) add
      // 0.5/tan(angle at the third vertex)
      for (int e = 0; e < deg; e++) {
        const int32_t t = ent[e] >> 2, k = ent[e] & 3;
        const int64_t a = el[3 * (int64_t)t + (k + 1) % 3], b = el[3 * (int64_t)t + (k + 2) % 3];
        int64_t c;  // third vertex
        if (a == j) c = b;
        else if (b == j) c = a;
        else continue;
        const P2 pc = ldp(sites, c), pj = ldp(sites, j);
        // the reference orders (vec1, vec2) by local index of the two far vertices; the dot
        // product and the norm product are commutative, so only the values matter
        double th = angle_between(pv.x - pc.x, pv.y - pc.y, pj.x - pc.x, pj.y - pc.y);
        wij += 0.5 / tan(th);
      }
    }
    rowsum += wij;
    o.laplacian[q] = wij;  // scaled below
  }
  const double inv_mass = 1.0 / o.vertex_areas[v];
  for (int32_t q = r0; q < r1; q++) {
    if (q == diag_pos) o.laplacian[q] = inv_mass * (-rowsum);
    else o.laplacian[q] = inv_mass * o.laplacian[q];
  }

  // ---------------- vertex gradient ----------------
  const int32_t* stris = o.star_tris + s0;  // head-ordered star
  double tot = 0.0;
  for (int e = 0; e < deg; e++) {
    const int64_t t = stris[e];
    const P2 p0 = ldp(sites, el[3 * t]), p1 = ldp(sites, el[3 * t + 1]), p2 = ldp(sites, el[3 * t + 2]);
    // quirk Q1: angle at the triangle's LOCAL vertex 0 (fem.py:393-398)
    tot += angle_between(p1.x - p0.x, p1.y - p0.y, p2.x - p0.x, p2.y - p0.y);
  }
  for (int32_t q = r0; q < r1; q++) {
    const int32_t j = o.op_indices[q];
    double gx = 0.0, gy = 0.0;
    for (int e = 0; e < deg; e++) {
      const int64_t t = stris[e];
      const int64_t tv[3] = {el[3 * t], el[3 * t + 1], el[3 * t + 2]};
      int kk = tv[0] == j ? 0 : (tv[1] == j ? 1 : (tv[2] == j ? 2 : -1));
      if (kk < 0) continue;
      const P2 p0 = ldp(sites, tv[0]), p1 = ldp(sites, tv[1]), p2 = ldp(sites, tv[2]);
      const double th = angle_between(p1.x - p0.x, p1.y - p0.y, p2.x - p0.x, p2.y - p0.y);
      const double omega = th / tot;
      const P2 pp[3] = {p0, p1, p2};
      const P2 pa = pp[(kk + 1) % 3], pb = pp[(kk + 2) % 3];
      const double two_a = 2.0 * o.triangle_areas[t];
      gx += omega * ((pa.y - pb.y) / two_a);
      gy += omega * (-(pa.x - pb.x) / two_a);
    }
    o.gradient_x[q] = gx;
    o.gradient_y[q] = gy;
  }
}

// C_vector (device/mesh.py:400-432).  The four reductions (mean x/y, min/max x/y) are done by
// one block in a fixed order -> deterministic.
__global__ void c_vector_stats_kernel(int64_t n, const double* __restrict__ sites, double* stats) {
  __shared__ double sx[1024], sy[1024], mnx[1024], mxx[1024], mny[1024], mxy[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  const int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
  double ax = 0, ay = 0, a0 = 1e300, a1 = -1e300, b0 = 1e300, b1 = -1e300;
  for (int64_t i = lo; i < hi; i++) {
    double x = sites[2 * i], y = sites[2 * i + 1];
    ax += x; ay += y;
    a0 = fmin(a0, x); a1 = fmax(a1, x);
    b0 = fmin(b0, y); b1 = fmax(b1, y);
  }
  sx[t] = ax; sy[t] = ay; mnx[t] = a0; mxx[t] = a1; mny[t] = b0; mxy[t] = b1;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (t < off) {
      sx[t] += sx[t + off]; sy[t] += sy[t + off];
      mnx[t] = fmin(mnx[t], mnx[t + off]); mxx[t] = fmax(mxx[t], mxx[t + off]);
      mny[t] = fmin(mny[t], mny[t + off]); mxy[t] = fmax(mxy[t], mxy[t + off]);
    }
    __syncthreads();
  }
  if (t == 0) {
    stats[0] = sx[0] / (double)n;  // mean x
    stats[1] = sy[0] / (double)n;
    stats[2] = mnx[0]; stats[3] = mxx[0]; stats[4] = mny[0]; stats[5] = mxy[0];
  }
}

__global__ void c_vector_kernel(int64_t n, const double* __restrict__ sites,
                                const double* __restrict__ stats, double* C) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mx = stats[0], my = stats[1];
  // ptp of the centred coordinates: (max - mean) - (min - mean), as numpy evaluates it
  const double a = ((stats[3] - mx) - (stats[2] - mx)) / 2.0;
  const double b = ((stats[5] - my) - (stats[4] - my)) / 2.0;
  const double x = sites[2 * i] - mx, y = sites[2 * i + 1] - my;
  double c = 0.0;
#pragma unroll
  for (int p = -1; p <= 1; p += 2)
#pragma unroll
    for (int q = -1; q <= 1; q += 2) {
      double u = a - p * x, w2 = b - q * y;
      c += sqrt(1.0 / (u * u) + 1.0 / (w2 * w2));
    }
  if (isinf(c)) c = 1e30;
  C[i] = c / (4.0 * 3.141592653589793);
}

}  // namespace scb

using namespace scb;

extern "C" int64_t scb_mesh_workspace_elems(int64_t n, int64_t m) {
  return 4 * (n + 1) + n + 9 * m + 1024 + 8 + 16 /* stats as doubles */;
}

extern "C" int scb_mesh_analyze(int64_t n, int64_t m, const int64_t* elements, int32_t* workspace,
                                int64_t* counts, int32_t* flags, scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0 && m > 0, "empty mesh");
  SCB_CHECK_ARG(n < (1ll << 29) && m < (1ll << 28), "mesh too large for int32 internals");
  cudaStream_t s = (cudaStream_t)stream;
  MeshWs w = carve(workspace, n, m);
  SCB_CUDA(cudaMemsetAsync(workspace, 0, sizeof(int32_t) * scb_mesh_workspace_elems(n, m), s));
  SCB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * 4, s));
  SCB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), s));
  const int T = 256;
  count_star_kernel<<<(unsigned)ceil_div(3 * m, T), T, 0, s>>>(m, elements, w.v2t_ptr);
  SCB_LAUNCH_CHECK();
  scan_exclusive_kernel<<<1, 1024, 0, s>>>(w.v2t_ptr, n + 1);
  SCB_LAUNCH_CHECK();
  fill_star_kernel<<<(unsigned)ceil_div(3 * m, T), T, 0, s>>>(m, elements, w.v2t_ptr, w.cursor, w.v2t_ent);
  SCB_LAUNCH_CHECK();
  star_analyze_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, s>>>(n, elements, w, counts, flags);
  SCB_LAUNCH_CHECK();
  scan_exclusive_kernel<<<1, 1024, 0, s>>>(w.adj_ptr, n + 1);
  SCB_LAUNCH_CHECK();
  scan_exclusive_kernel<<<1, 1024, 0, s>>>(w.edge_ptr, n + 1);
  SCB_LAUNCH_CHECK();
  scan_exclusive_kernel<<<1, 1024, 0, s>>>(w.bvert_ptr, n + 1);
  SCB_LAUNCH_CHECK();
  totals_kernel<<<1, 1, 0, s>>>(n, w, counts);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int scb_mesh_build(int64_t n, int64_t m, const double* sites, const int64_t* elements,
                              const int32_t* workspace, int weight_method, const scb_mesh_out* out,
                              scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0 && m > 0, "empty mesh");
  SCB_CHECK_ARG(out != nullptr, "null output struct");
  SCB_CHECK_ARG(weight_method >= 0 && weight_method <= 2, "unknown weight method");
  cudaStream_t s = (cudaStream_t)stream;
  MeshWs w = carve(const_cast<int32_t*>(workspace), n, m);
  scb_mesh_out o = *out;
  const int T = 128;
  triangle_kernel<<<(unsigned)ceil_div(m, T), T, 0, s>>>(m, sites, elements, o);
  SCB_LAUNCH_CHECK();
  vertex_structure_kernel<<<(unsigned)ceil_div(n, T), T, 0, s>>>(n, sites, elements, w, o);
  SCB_LAUNCH_CHECK();
  vertex_operator_kernel<<<(unsigned)ceil_div(n, T), T, 0, s>>>(n, sites, elements, w, weight_method, o);
  SCB_LAUNCH_CHECK();
  // stats live (as doubles) in the tail of the workspace, 8-byte aligned
  int64_t off = 4 * (n + 1) + n + 9 * m + 1024;
  off = (off + 1) & ~1ll;
  double* stats = reinterpret_cast<double*>(const_cast<int32_t*>(workspace) + off);
  c_vector_stats_kernel<<<1, 1024, 0, s>>>(n, sites, stats);
  SCB_LAUNCH_CHECK();
  c_vector_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(n, sites, stats, o.C);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

extern "C" int scb_c_vector(int64_t n, const double* points, double* scratch, double* C, scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0, "n must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  c_vector_stats_kernel<<<1, 1024, 0, s>>>(n, points, scratch);
  SCB_LAUNCH_CHECK();
  c_vector_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(n, points, scratch, C);
  SCB_LAUNCH_CHECK();
  return SCB_OK;
}

// ---------------------------------------------------------------------------------------
// Laplacian smoothing sweep (reference device/mesh.py:172-211): every vertex moves to the mean of
// its neighbours; boundary vertices are restored afterwards.  The reference accumulates, per
// coordinate, first the neighbours j > i and then the neighbours j < i, each in ascending order
// (two np.bincount passes over the lexicographically sorted edge list), and divides by the
// neighbour count: the same order is used here so the result is bit-identical.
// ---------------------------------------------------------------------------------------
namespace scb {
__global__ void smooth_kernel(int64_t n, const double* __restrict__ sites, const int32_t* __restrict__ indptr,
                              const int32_t* __restrict__ indices, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = indptr[i], e = indptr[i + 1];
  double lox = 0.0, loy = 0.0, hix = 0.0, hiy = 0.0;
  for (int p = b; p < e; p++) {  // rows are sorted: j < i first, then j > i
    const int j = indices[p];
    const double2 sj = *reinterpret_cast<const double2*>(sites + 2 * (int64_t)j);
    if (j < i) {
      lox += sj.x;
      loy += sj.y;
    } else {
      hix += sj.x;
      hiy += sj.y;
    }
  }
  const double cnt = (double)(e - b);
  out[2 * i] = (hix + lox) / cnt;
  out[2 * i + 1] = (hiy + loy) / cnt;
}
__global__ void restore_boundary_kernel(int64_t nb, const int64_t* __restrict__ boundary,
                                        const double* __restrict__ sites, double* __restrict__ out) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= nb) return;
  const int64_t i = boundary[k];
  out[2 * i] = sites[2 * i];
  out[2 * i + 1] = sites[2 * i + 1];
}
}  // namespace scb

extern "C" int scb_mesh_smooth(int64_t n, const double* sites, const int32_t* adj_indptr,
                               const int32_t* adj_indices, int64_t n_boundary, const int64_t* boundary_indices,
                               double* out_sites, scb_stream_t stream) {
  SCB_CHECK_ARG(n > 0, "empty mesh");
  SCB_CHECK_ARG(sites != out_sites, "smoothing cannot run in place");
  cudaStream_t s = (cudaStream_t)stream;
  scb::smooth_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, sites, adj_indptr, adj_indices, out_sites);
  SCB_LAUNCH_CHECK();
  if (n_boundary > 0) {
    scb::restore_boundary_kernel<<<(unsigned)((n_boundary + 127) / 128), 128, 0, s>>>(n_boundary, boundary_indices,
                                                                                     sites, out_sites);
    SCB_LAUNCH_CHECK();
  }
  return SCB_OK;
}
