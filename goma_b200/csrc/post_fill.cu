// What surrounds matrix_fill_full in solve_nonlinear_problem, on the device-resident system: the PSPG global norms
// before it (src/mm_fill_aux.c:1128,612), row-sum scaling and residual norms after it (src/sl_matrix_util.c:441,
// src/mm_sol_nonlinear.c:3177-3375), and the CSR hand-off to a GPU solver.  See include/goma_gpu_fill.h.
#include <cstring>

#include "ctx.h"

using namespace goma_b200;

// ------------------------------------------------------------------ PSPG global norms
// h_elem_siz (mm_fill_aux.c:844-1070): squared distances between opposite face centroids, from the vertex nodes
__global__ void global_h_U_kernel(const int *__restrict__ conn, int npe, int dim, int num_elems,
                                  const double *__restrict__ cx, const double *__restrict__ cy,
                                  const double *__restrict__ cz, const unsigned char *__restrict__ elem_owned,
                                  const int *__restrict__ first_unknown, const unsigned char *__restrict__ node_kind,
                                  int slot_u0, int slot_u1, int slot_u2, int k0u, int k1u, int k2u, int k3u,
                                  int num_owned_nodes, const double *__restrict__ x, double *__restrict__ sums) {
  double h = 0.0, cnt = 0.0, vv = 0.0, nv = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < num_elems; e += stride) {
    if (elem_owned && !elem_owned[e]) continue;
    const int *c = conn + (size_t)e * npe;
    double hsq = 0.0;
    for (int a = 0; a < dim; a++) {
      const double *X = a == 0 ? cx : (a == 1 ? cy : cz);
      if (dim == 2) {
        const double x0 = X[c[0]], x1 = X[c[1]], x2 = X[c[2]], x3 = X[c[3]];
        const double h0 = 0.5 * (x1 + x2) - 0.5 * (x0 + x3), h1 = 0.5 * (x0 + x1) - 0.5 * (x2 + x3);
        hsq += h0 * h0 + h1 * h1;
      } else {
        double v[8];
        for (int k = 0; k < 8; k++) v[k] = X[c[k]];
        const double p1 = 0.25 * (v[0] + v[1] + v[2] + v[3]), p2 = 0.25 * (v[1] + v[2] + v[5] + v[6]);
        const double p3 = 0.25 * (v[2] + v[3] + v[6] + v[7]), p4 = 0.25 * (v[0] + v[1] + v[4] + v[5]);
        const double p5 = 0.25 * (v[0] + v[3] + v[4] + v[7]), p6 = 0.25 * (v[4] + v[5] + v[6] + v[7]);
        hsq += (p2 - p5) * (p2 - p5) + (p3 - p4) * (p3 - p4) + (p1 - p6) * (p1 - p6);
      }
    }
    h += sqrt(hsq / (double)dim);
    cnt += 1.0;
  }
  const int ku[4][3] = {{k0u, k0u + 1, k0u + 2}, {k1u, k1u + 1, k1u + 2}, {k2u, k2u + 1, k2u + 2}, {k3u, k3u + 1, k3u + 2}};
  (void)slot_u0; (void)slot_u1; (void)slot_u2;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < num_owned_nodes; n += stride) {
    const int kd = node_kind[n];
    if (ku[kd][0] < 0) continue;
    const int f = first_unknown[n];
    for (int a = 0; a < dim; a++) {
      const double v = x[f + ku[kd][a]];
      vv += v * v;
      nv += 1.0;
    }
  }
  // block partials, summed by the host in block order: the same bits run to run (the PSPG tau, and with it the
  // whole equal-order fill, depends on these sums)
  __shared__ double sh[4][8];
  double vals[4] = {h, cnt, vv, nv};
  for (int q = 0; q < 4; q++) {
    double v = vals[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += sh[threadIdx.x][w];
    sums[4 * blockIdx.x + threadIdx.x] = v;
  }
}

extern "C" int goma_gpu_global_h_U(goma_gpu_ctx *c, const unsigned char *elem_owned, double sums_out[4]) {
  if (!c || !sums_out) return fail(-2, "null argument");
  CU(cudaSetDevice(c->device));
  const goma_gpu_problem &p = c->prob;
  unsigned char *d_owned = nullptr;
  if (elem_owned) {  // the mask travels with every call (num_elems bytes); its buffer is kept
    if (!c->d_elem_owned) CU(cudaMalloc((void **)&c->d_elem_owned, std::max(1, p.num_elems)));
    d_owned = c->d_elem_owned;
    CU(cudaMemcpyAsync(d_owned, elem_owned, p.num_elems, cudaMemcpyHostToDevice, c->stream));
  }
  constexpr int HU_BLOCKS_MAX = 148 * 8;
  if (!c->d_sums) CU(cudaMalloc((void **)&c->d_sums, 4 * HU_BLOCKS_MAX * sizeof(double)));
  int ku[4] = {-1, -1, -1, -1};  // offset of U inside a node of each kind (V, W follow it)
  for (int k = 0; k < p.num_kinds && k < 4; k++) ku[k] = p.kind_slot[k][GOMA_SLOT_U];
  const int threads = 256, blocks = std::max(1, std::min(148 * 8, (std::max(p.num_elems, p.num_owned_nodes) + threads - 1) / threads));
  global_h_U_kernel<<<blocks, threads, 0, c->stream>>>(c->d_conn, p.elem_type, p.dim, p.num_elems, c->d_coord[0],
                                                       c->d_coord[1], c->d_coord[2], d_owned, c->d_first, c->d_kind, 0, 1, 2,
                                                       ku[0], ku[1], ku[2], ku[3], p.num_owned_nodes, c->d_x, c->d_sums);
  CU(cudaGetLastError());
  std::vector<double> part(4 * (size_t)blocks);
  CU(cudaMemcpyAsync(part.data(), c->d_sums, part.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int q = 0; q < 4; q++) sums_out[q] = 0.0;
  for (int b = 0; b < blocks; b++)
    for (int q = 0; q < 4; q++) sums_out[q] += part[4 * (size_t)b + q];
  return 0;
}

// ------------------------------------------------------------------ after the fill: row-sum scaling, norms
// row_sum_scale_MSR (src/sl_matrix_util.c:507-600).  Consecutive rows are contiguous in memory (MSR off-diagonals as
// well as CSR rows), so a CTA stages a BATCH of rows in shared memory with asynchronous copies (cp.async: the whole
// batch in flight at once), sums and scales it there (one warp per row) and streams it back with 16-byte stores:
// exactly one read and one write of every value (ncu: 45.1 + 45.0 GB of DRAM traffic for 89.3 GB algorithmic,
// profiles/r2r_row_sum_scale_4rows.txt), no reliance on L1/L2 for the second sweep.
// The copies move 16-byte pairs: the buffer is shifted by the parity of the batch's first index so that aligned pairs
// of a[] land on aligned slots.  ROWS rows per batch, one warp per row; the buffer holds ROWS x the longest row of
// THIS matrix (dynamic shared memory).  Measured at 1M hex27 elements (profiles/r2n_row_sum_scale_variants.txt,
// r2s_row_sum_scale_latency.txt): 22.9 ms with 8 rows and 8-byte copies -> 19.4 ms with 4 rows and 16-byte copies ->
// 17.8-18.4 ms (0.74-0.77 of the copy roof; CSR layout of C3 0.82-0.83) once the three dependent DRAM latencies of a
// batch (row starts, values, per-row scalars) were folded into one.  Slower: 2 rows, a two-buffer pipeline, one warp
// per row with the row held in registers, batches cut by entry count.
#define RSS_THREADS (32 * ROWS)
#define RSS_ROWS ROWS

__device__ __forceinline__ void rss_cp_async8(void *dst, const void *src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void rss_cp_async16(void *dst, const void *src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

// longest owned row (entries staged per row: MSR off-diagonals, or the CSR row with its diagonal), once per context
__global__ void max_row_len_kernel(int nrows, const long long *__restrict__ rowstart, int *__restrict__ out) {
  int m = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) m = max(m, (int)(rowstart[r + 1] - rowstart[r]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__device__ __forceinline__ void rss_cp_async4(void *dst, const void *src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

// One DRAM latency per batch instead of three: the row starts of the NEXT batch are fetched while the current one is
// in flight, and the per-row scalars (diagonal a[row] of the MSR layout or its offset in the CSR row, b[row]) travel
// with the same asynchronous copy group as the values instead of being loaded by lane 0 after the sums.
template <bool CSR, int ROWS>
__global__ void __launch_bounds__(RSS_THREADS) row_sum_scale_kernel(int nrows, const long long *__restrict__ rowstart, long long msr0,
                                                                    const int *__restrict__ dpos, double *__restrict__ a,
                                                                    double *__restrict__ b, double *__restrict__ scale,
                                                                    int *__restrict__ zero_rows, int cap) {
  extern __shared__ __align__(16) double buf_[];  // cap + 2 doubles: ROWS x the longest row of this matrix
  __shared__ long long rs_[2][RSS_ROWS + 1];
  __shared__ double dg[RSS_ROWS], bb[RSS_ROWS];
  __shared__ int dp[RSS_ROWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbatch = (nrows + RSS_ROWS - 1) / RSS_ROWS;
  auto row_start = [&](int r) -> long long { return CSR ? rowstart[r] - msr0 + r : rowstart[r]; };
  if ((int)blockIdx.x < nbatch && tid <= min(RSS_ROWS, nrows - (int)blockIdx.x * RSS_ROWS)) rs_[0][tid] = row_start(blockIdx.x * RSS_ROWS + tid);
  __syncthreads();
  int it = 0;
  for (int bt = blockIdx.x; bt < nbatch; bt += gridDim.x, it++) {
    const long long *rs = rs_[it & 1];
    const int r0 = bt * RSS_ROWS, nr = min(RSS_ROWS, nrows - r0);
    const int bt2 = bt + (int)gridDim.x;
    const bool fetch_next = bt2 < nbatch && tid <= min(RSS_ROWS, nrows - bt2 * RSS_ROWS);
    long long rs_next = 0;
    if (fetch_next) rs_next = row_start(bt2 * RSS_ROWS + tid);  // lands while the copies below are in flight
    const long long k0 = rs[0];
    const int len = (int)(rs[nr] - k0);
    const bool staged = len <= cap;  // (always, unless the longest row outgrows the shared memory of an SM)
    // the buffer is shifted by the parity of k0 so that 16-byte-aligned pairs of a[] land on 16-byte-aligned slots;
    // the first and last entries of the batch are copied alone when they do not fill a pair
    const int sh = (int)(k0 & 1);
    double *buf = buf_ + sh;
    const int p_lo = sh, p_hi = (len - sh) >> 1;  // entries [p_lo, p_lo + 2 * p_hi) go as pairs
    if (tid < nr) {
      rss_cp_async8(&bb[tid], &b[r0 + tid]);
      if (CSR)
        rss_cp_async4(&dp[tid], &dpos[r0 + tid]);
      else
        rss_cp_async8(&dg[tid], &a[r0 + tid]);
    }
    if (staged) {
      for (int q = tid; q < p_hi; q += RSS_THREADS) rss_cp_async16(&buf[p_lo + 2 * q], &a[k0 + p_lo + 2 * q]);
      if (tid == 0 && sh && len > 0) rss_cp_async8(&buf[0], &a[k0]);
      if (tid == 32 % RSS_THREADS && p_lo + 2 * p_hi < len) rss_cp_async8(&buf[len - 1], &a[k0 + len - 1]);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (fetch_next) rs_[(it & 1) ^ 1][tid] = rs_next;
    __syncthreads();
    if (warp < nr) {
      const int row = r0 + warp;
      const int o0 = (int)(rs[warp] - k0), o1 = (int)(rs[warp + 1] - k0);
      double sum = 0.0;
      if (staged)
        for (int k = o0 + lane; k < o1; k += 32) sum += fabs(buf[k]);
      else
        for (long long k = rs[warp] + lane; k < rs[warp + 1]; k += 32) sum += fabs(a[k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      // MSR: the diagonal lives apart in a[row]; CSR: it is one of the staged entries (already in the sum)
      const double diag = CSR ? (staged ? buf[o0 + dp[warp]] : a[rs[warp] + dp[warp]]) : dg[warp];
      double row_sum = CSR ? sum : fabs(diag) + sum;
      if (fabs(diag) > 1.0e-200) row_sum = diag >= 0.0 ? row_sum : -row_sum;  // keep the diagonal positive (:547-549)
      // one reciprocal per row and a multiply per entry differ from the reference's divide by at most 1 ulp (parity
      // tolerance 1e-12) and keep the fp64 divide sequence off an HBM-bound pass
      const double inv = 1.0 / row_sum;
      __syncwarp();  // (every lane has read the diagonal before it is scaled)
      if (staged)
        for (int k = o0 + lane; k < o1; k += 32) buf[k] *= inv;
      else
        for (long long k = rs[warp] + lane; k < rs[warp + 1]; k += 32) a[k] *= inv;
      if (lane == 0) {
        scale[row] = row_sum;
        if (row_sum == 0.0) atomicAdd(zero_rows, 1);
        if (!CSR) a[row] = diag / row_sum;
        b[row] = bb[warp] / row_sum;
      }
    }
    __syncthreads();
    if (staged) {
      for (int q = tid; q < p_hi; q += RSS_THREADS)
        *reinterpret_cast<double2 *>(&a[k0 + p_lo + 2 * q]) = *reinterpret_cast<const double2 *>(&buf[p_lo + 2 * q]);
      if (tid == 0 && sh && len > 0) a[k0] = buf[0];
      if (tid == 32 % RSS_THREADS && p_lo + 2 * p_hi < len) a[k0 + len - 1] = buf[len - 1];
    }
    __syncthreads();
  }
}

// CSR layout: offset of the diagonal inside every owned row, from the node-node lists (init only)
__global__ void csr_dpos_kernel(int num_owned_nodes, const long long *__restrict__ nn_ptr, const int *__restrict__ nn_list,
                                const unsigned short *__restrict__ cum_full, const unsigned short *__restrict__ cum_p,
                                const int *__restrict__ first_unknown, const unsigned char *__restrict__ node_kind,
                                const __grid_constant__ KindInfo K, int *__restrict__ dpos) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= num_owned_nodes) return;
  const long long b = nn_ptr[nd];
  int lo = 0, hi = (int)(nn_ptr[nd + 1] - b);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nn_list[b + mid] < nd)
      lo = mid + 1;
    else
      hi = mid;
  }
  const int kd = node_kind[nd], fu = first_unknown[nd];
  for (int s = 0; s < K.nunk[kd]; s++)
    dpos[fu + s] = cum_full[b + lo] + s - ((K.tslot[kd] >= 0 && s == K.tslot[kd]) ? cum_p[b + lo] : 0);
}

int goma_b200::build_csr_dpos(goma_gpu_ctx *c) {
  const int no = c->num_owned_unknowns, nown = c->prob.num_owned_nodes;
  if (c->d_dpos) return 0;
  CU(cudaMalloc((void **)&c->d_dpos, std::max<size_t>(no, 1) * sizeof(int)));
  c->device_bytes += (size_t)no * sizeof(int);
  if (nown > 0) {
    const KindInfo K = make_kind_info(c->prob);
    csr_dpos_kernel<<<(nown + 127) / 128, 128, 0, c->stream>>>(nown, c->dpat.nn_ptr, c->dpat.nn_list, c->dpat.cum_full, c->dpat.cum_p,
                                                                c->d_first, c->d_kind, K, c->d_dpos);
    CU(cudaGetLastError());
  }
  return 0;
}

extern "C" int goma_gpu_row_sum_scale(goma_gpu_ctx *c, double *scale_out, int *zero_rows_out) {
  if (!c) return fail(-2, "null context");
  CU(cudaSetDevice(c->device));
  const int n = c->num_owned_unknowns;
  if (!c->d_scale) CU(cudaMalloc((void **)&c->d_scale, std::max(1, c->prob.num_unknowns) * sizeof(double)));
  if (!c->d_zero_rows) CU(cudaMalloc((void **)&c->d_zero_rows, sizeof(int)));
  CU(cudaMemsetAsync(c->d_zero_rows, 0, sizeof(int), c->stream));
  if (n > 0) {
    const bool csr = c->layout == GOMA_GPU_LAYOUT_CSR;
    const long long msr0 = (long long)c->prob.num_unknowns + 1;
    static const int rss_rows = getenv("GOMA_GPU_RSS_ROWS") ? atoi(getenv("GOMA_GPU_RSS_ROWS")) : 4;  // (2 and 8 kept for A/B runs)
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (c->rss_max_row < 0) {  // the staging buffer is sized for this matrix, not for the worst case: more CTAs per SM
      CU(cudaMemsetAsync(c->d_zero_rows, 0, sizeof(int), c->stream));
      max_row_len_kernel<<<std::min(1024, (n + 255) / 256), 256, 0, c->stream>>>(n, c->d_rowstart, c->d_zero_rows);
      CU(cudaMemcpyAsync(&c->rss_max_row, c->d_zero_rows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      CU(cudaMemsetAsync(c->d_zero_rows, 0, sizeof(int), c->stream));
      if (csr) c->rss_max_row += 1;
    }
    auto launch = [&](auto kern, int rows) -> int {
      int cap = (rows * c->rss_max_row + 1) & ~1;
      size_t dyn = (size_t)(cap + 2) * sizeof(double);
      if (dyn > 200 * 1024) {  // a row of more than 200 KB / rows: those batches take the two-pass path
        cap = 2048;
        dyn = (size_t)(cap + 2) * sizeof(double);
      }
      if (dyn > 48 * 1024) CU(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      int per_sm = 0;  // a whole number of resident waves: the batches are handed out grid-stride
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, 32 * rows, dyn));
      const int blocks = std::max(1, std::min(sms * std::max(per_sm, 1), (n + rows - 1) / rows));
      kern<<<blocks, 32 * rows, dyn, c->stream>>>(n, c->d_rowstart, msr0, csr ? c->d_dpos : nullptr, c->d_a, c->d_resid, c->d_scale, c->d_zero_rows,
                                                  cap);
      return 0;
    };
    int lrc = 0;
    if (csr)
      lrc = rss_rows == 2 ? launch(row_sum_scale_kernel<true, 2>, 2) : rss_rows == 4 ? launch(row_sum_scale_kernel<true, 4>, 4)
            : launch(row_sum_scale_kernel<true, 8>, 8);
    else
      lrc = rss_rows == 2 ? launch(row_sum_scale_kernel<false, 2>, 2) : rss_rows == 4 ? launch(row_sum_scale_kernel<false, 4>, 4)
            : launch(row_sum_scale_kernel<false, 8>, 8);
    if (lrc) return lrc;
    CU(cudaGetLastError());
  }
  int zr = 0;
  CU(cudaMemcpyAsync(&zr, c->d_zero_rows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (scale_out && n > 0) CU(cudaMemcpyAsync(scale_out, c->d_scale, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (zero_rows_out) *zero_rows_out = zr;
  // a zero row sum turned its never-touched (zero) slots into 0 * inf = NaN, as in the reference -- but there the
  // next fill starts from re-zeroed storage: make the next first-touch fill do the same
  if (zr > 0) c->rezero = true;
  return 0;
}

extern "C" int goma_gpu_scale_buffer(goma_gpu_ctx *c, double **d_scale, int *num_owned_unknowns) {
  if (!c) return fail(-2, "null context");
  if (d_scale) *d_scale = c->d_scale;
  if (num_owned_unknowns) *num_owned_unknowns = c->num_owned_unknowns;
  return 0;
}

constexpr int NORM_BLOCKS = 592, NORM_THREADS = 256;
__global__ void vector_norms_kernel(const double *__restrict__ v, int n, double *__restrict__ partials) {
  __shared__ double sh[4][NORM_THREADS / 32];
  double mx = -1.0, l1 = 0.0, l2 = 0.0, idx = -1.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double t = v[i], w = fabs(t);
    l1 += w;
    l2 += t * t;
    if (w > mx) { mx = w; idx = (double)i; }  // first occurrence of the maximum, as the reference's strict '>'
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    l2 += __shfl_xor_sync(0xffffffffu, l2, o);
    const double m2 = __shfl_xor_sync(0xffffffffu, mx, o), i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (m2 > mx || (m2 == mx && i2 >= 0.0 && (idx < 0.0 || i2 < idx))) { mx = m2; idx = i2; }
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh[0][w] = mx; sh[1][w] = l1; sh[2][w] = l2; sh[3][w] = idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < NORM_THREADS / 32; k++) {
      l1 += sh[1][k];
      l2 += sh[2][k];
      if (sh[0][k] > mx || (sh[0][k] == mx && sh[3][k] >= 0.0 && (idx < 0.0 || sh[3][k] < idx))) { mx = sh[0][k]; idx = sh[3][k]; }
    }
    double *o = partials + 4 * blockIdx.x;
    o[0] = mx; o[1] = l1; o[2] = l2; o[3] = idx;
  }
}

extern "C" int goma_gpu_vector_norms(goma_gpu_ctx *c, int which, double out[4]) {
  if (!c || !out) return fail(-2, "null argument");
  if (which < 0 || which > 2) return fail(-2, "which must be 0 (resid), 1 (x) or 2 (xdot)");
  CU(cudaSetDevice(c->device));
  const double *v = which == 0 ? c->d_resid : (which == 1 ? c->d_x : c->d_xdot);
  if (!c->d_partials) CU(cudaMalloc((void **)&c->d_partials, 4 * NORM_BLOCKS * sizeof(double)));
  vector_norms_kernel<<<NORM_BLOCKS, NORM_THREADS, 0, c->stream>>>(v, c->num_owned_unknowns, c->d_partials);
  CU(cudaGetLastError());
  std::vector<double> h(4 * NORM_BLOCKS);
  CU(cudaMemcpyAsync(h.data(), c->d_partials, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  double mx = -1.0, l1 = 0.0, l2 = 0.0, idx = -1.0;
  for (int b = 0; b < NORM_BLOCKS; b++) {  // fixed order: reproducible run to run
    l1 += h[4 * b + 1];
    l2 += h[4 * b + 2];
    if (h[4 * b] > mx || (h[4 * b] == mx && h[4 * b + 3] >= 0.0 && (idx < 0.0 || h[4 * b + 3] < idx))) { mx = h[4 * b]; idx = h[4 * b + 3]; }
  }
  out[0] = mx; out[1] = l1; out[2] = l2; out[3] = idx;
  return 0;
}

// ------------------------------------------------------------------ CSR hand-off to a GPU solver

// one thread per owned node: the rows of its unknowns share the node-node list (exo_conn.c build_node_node);
// columns = the unknowns of the neighbour nodes in increasing node id (find_MSR_problem_graph), diagonal included
__global__ void csr_structure_kernel(int num_owned_nodes, const long long *__restrict__ nn_ptr,
                                     const int *__restrict__ nn_list, const int *__restrict__ first_unknown,
                                     const unsigned char *__restrict__ node_kind, const __grid_constant__ KindInfo K,
                                     const long long *__restrict__ rowstart, long long msr0,
                                     long long *__restrict__ rowptr, int *__restrict__ colind, int *__restrict__ dpos,
                                     int num_rows) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= num_owned_nodes) return;
  const int kd = node_kind[nd], fu = first_unknown[nd];
  for (int s = 0; s < K.nunk[kd]; s++) {
    const int row = fu + s;
    const long long base = rowstart[row] - msr0 + row;  // every earlier row adds its diagonal
    rowptr[row] = base;
    if (row == num_rows - 1) rowptr[num_rows] = rowstart[row + 1] - msr0 + row + 1;
    const bool nop = K.tslot[kd] >= 0 && s == K.tslot[kd];
    long long pos = base;
    for (long long q = nn_ptr[nd]; q < nn_ptr[nd + 1]; q++) {
      const int m = nn_list[q], km = node_kind[m], fm = first_unknown[m];
      const int ncol = K.nunk[km] - (nop ? K.npress[km] : 0);
      for (int c = 0; c < ncol; c++) {
        if (fm + c == row) dpos[row] = (int)(pos - base);
        colind[pos++] = fm + c;
      }
    }
  }
}

// one warp per row: MSR row (diagonal apart) -> CSR row (diagonal at dpos)
__global__ void csr_values_kernel(int num_rows, const long long *__restrict__ rowstart, const long long *__restrict__ rowptr,
                                  const int *__restrict__ dpos, const double *__restrict__ a, double *__restrict__ v) {
  const int lane = threadIdx.x & 31;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= num_rows) return;
  // the row descriptors of the NEXT row are loaded under the copy of the current one (one DRAM latency per row, not two)
  long long k0 = rowstart[row], c0 = rowptr[row], c1 = rowptr[row + 1];
  int d = dpos[row];
  double dg = a[row];
  while (true) {
    const int nxt = row + nwarp;
    const bool more = nxt < num_rows;
    long long k0n = 0, c0n = 0, c1n = 0;
    int dn = 0;
    double dgn = 0.0;
    if (more) {
      k0n = rowstart[nxt];
      c0n = rowptr[nxt];
      c1n = rowptr[nxt + 1];
      dn = dpos[nxt];
      dgn = a[nxt];
    }
    const int len = (int)(c1 - c0);
    for (int t = lane; t < len; t += 32) v[c0 + t] = t < d ? a[k0 + t] : (t == d ? dg : a[k0 + t - 1]);
    if (!more) break;
    row = nxt; k0 = k0n; c0 = c0n; c1 = c1n; d = dn; dg = dgn;
  }
}

extern "C" int goma_gpu_csr_structure(goma_gpu_ctx *c, const goma_gpu_problem *p, goma_gpu_csr *out) {
  if (!c || !p || !out) return fail(-2, "null argument");
  CU(cudaSetDevice(c->device));
  const int nrows = c->num_owned_unknowns;
  if (!c->d_csr_colind) {
    if (!c->dpat.nn_ptr || !c->dpat.nn_list) return fail(-2, "node-node lists are not available");
    long long h_rs[2] = {0, 0};
    CU(cudaMemcpy(&h_rs[0], c->d_rowstart, sizeof(long long), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&h_rs[1], c->d_rowstart + nrows, sizeof(long long), cudaMemcpyDeviceToHost));
    const long long msr0 = h_rs[0];
    c->csr_nnz = nrows > 0 ? (h_rs[1] - msr0) + nrows : 0;
    const KindInfo K = make_kind_info(*p);
    long long *d_nn_ptr = c->dpat.nn_ptr;
    int *d_nn_list = c->dpat.nn_list;
    if (!c->d_csr_rowptr) CU(cudaMalloc((void **)&c->d_csr_rowptr, ((size_t)nrows + 1) * sizeof(long long)));
    CU(cudaMalloc((void **)&c->d_csr_colind, std::max<size_t>((size_t)c->csr_nnz, 1) * sizeof(int)));
    CU(cudaMalloc((void **)&c->d_csr_dpos, std::max<size_t>((size_t)nrows, 1) * sizeof(int)));
    if (c->layout == GOMA_GPU_LAYOUT_CSR)
      c->d_csr_values = c->d_a;  // the fill scatters straight into the CSR values: no second copy of the matrix
    else
      CU(cudaMalloc((void **)&c->d_csr_values, std::max<size_t>((size_t)c->csr_nnz, 1) * sizeof(double)));
    c->device_bytes += (size_t)c->csr_nnz * (c->layout == GOMA_GPU_LAYOUT_CSR ? 4 : 12) + (size_t)nrows * 12;
    CU(cudaMemset(c->d_csr_rowptr, 0, ((size_t)nrows + 1) * sizeof(long long)));
    const int nown = c->prob.num_owned_nodes;
    if (nown > 0 && nrows > 0) {
      csr_structure_kernel<<<(nown + 127) / 128, 128, 0, c->stream>>>(nown, d_nn_ptr, d_nn_list, c->d_first, c->d_kind, K,
                                                                      c->d_rowstart, msr0, c->d_csr_rowptr,
                                                                      c->d_csr_colind, c->d_csr_dpos, nrows);
      CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(c->stream));
  }
  out->num_rows = nrows;
  out->nnz = c->csr_nnz;
  out->d_rowptr = c->d_csr_rowptr;
  out->d_colind = c->d_csr_colind;
  out->d_values = c->d_csr_values;
  return 0;
}

extern "C" int goma_gpu_csr_values(goma_gpu_ctx *c) {
  if (!c) return fail(-2, "null context");
  if (c->layout == GOMA_GPU_LAYOUT_CSR) return 0;  // the values are assembled in place
  if (!c->d_csr_rowptr) return fail(-2, "call goma_gpu_csr_structure first");
  CU(cudaSetDevice(c->device));
  const int nrows = c->num_owned_unknowns;
  if (nrows > 0) {
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int blocks = std::max(1, std::min(sms * 8, (nrows + 7) / 8));
    csr_values_kernel<<<blocks, 256, 0, c->stream>>>(nrows, c->d_rowstart, c->d_csr_rowptr, c->d_csr_dpos, c->d_a, c->d_csr_values);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

__global__ void csr_rowptr_kernel(int nrows, const long long *__restrict__ rowstart, long long msr0, long long *__restrict__ rowptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= nrows) rowptr[r] = rowstart[r] - msr0 + r;
}

extern "C" int goma_gpu_csr_rows(goma_gpu_ctx *c, goma_gpu_csr *out) {
  if (!c || !out) return fail(-2, "null argument");
  if (c->layout != GOMA_GPU_LAYOUT_CSR)
    return fail(-2, "goma_gpu_csr_rows needs matrix_layout = GOMA_GPU_LAYOUT_CSR (use goma_gpu_csr_structure / _values for MSR)");
  CU(cudaSetDevice(c->device));
  const int nrows = c->num_owned_unknowns;
  if (!c->d_csr_rowptr) {
    CU(cudaMalloc((void **)&c->d_csr_rowptr, ((size_t)nrows + 1) * sizeof(long long)));
    csr_rowptr_kernel<<<(nrows + 256) / 256, 256, 0, c->stream>>>(nrows, c->d_rowstart, (long long)c->prob.num_unknowns + 1, c->d_csr_rowptr);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
  }
  out->num_rows = nrows;
  out->nnz = c->csr_nnz;
  out->d_rowptr = c->d_csr_rowptr;
  out->d_colind = c->d_csr_colind;  // NULL unless goma_gpu_csr_structure ran
  out->d_values = c->d_a;
  return 0;
}

// ------------------------------------------------------------------ w = A v on the device-resident matrix
// The product the Newton line search takes after a fill (mm_sol_nonlinear.c:442-449: AZ_MSR_matvec_mult for "msr",
// GomaSparseMatrix::matrix_vector_mult otherwise).  No column-index array is read: all rows of a node share the node's
// sorted neighbour list (exo_conn.c build_node_node) and their columns are the unknowns of those neighbours in node
// order -- energy rows skip the pressure unknowns (Inter_Mask) -- so the indices cost 4 bytes per NODE pair instead of
// 4 bytes per entry, and the pass streams the values once (8 bytes per entry).  One warp per row, lanes over the
// neighbour nodes (column offsets by a warp scan of the neighbours' unknown counts); MSR: the diagonal lives in a[row]
// and the entries behind it sit one slot earlier.
#define MV_WARPS 4    // warps per CTA, one node each
#define MV_CAP 1024   // columns of a node's rows staged per warp (hex27 NS + energy: 532); longer lists take the slow path
template <bool CSR>
__global__ void __launch_bounds__(32 * MV_WARPS, 12) node_graph_matvec_kernel(int num_owned_nodes, const long long *__restrict__ nn_ptr,
                                                                          const int *__restrict__ nn_list,
                                                                          const int *__restrict__ first_unknown,
                                                                          const unsigned char *__restrict__ node_kind,
                                                                          const __grid_constant__ KindInfo K,
                                                                          const long long *__restrict__ rowstart, long long msr0,
                                                                          const double *__restrict__ a, const double *__restrict__ v,
                                                                          double *__restrict__ w, int cap, int nlist) {
  // the column list of the node's rows, expanded once per node into shared memory (all its rows share it; the energy
  // row has its own, without the pressure unknowns): the values are then read in storage order, every byte once
  extern __shared__ int cols_[];  // [MV_WARPS][2 (1 without an energy equation)][cap]
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int nwarp = gridDim.x * MV_WARPS;
  int *colf = cols_ + (size_t)wl * nlist * cap, *coln = colf + (nlist > 1 ? cap : 0);
  for (int nd = blockIdx.x * MV_WARPS + wl; nd < num_owned_nodes; nd += nwarp) {
    const int kd = node_kind[nd], fu = first_unknown[nd], nu = K.nunk[kd], ts = K.tslot[kd];
    const long long b = nn_ptr[nd], e = nn_ptr[nd + 1];
    int totf = 0, totn = 0;  // columns of a full row / of the energy row
    bool fits = true;
    for (long long q0 = b; q0 < e; q0 += 32) {
      const long long q = q0 + lane;
      int fm = 0, nf = 0, np = 0;
      if (q < e) {
        const int m = nn_list[q], km = node_kind[m];
        fm = first_unknown[m];
        nf = K.nunk[km];
        np = ts >= 0 ? K.npress[km] : 0;
      }
      int incf = nf, incp = np;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incf, o), u = __shfl_up_sync(0xffffffffu, incp, o);
        if (lane >= o) {
          incf += t;
          incp += u;
        }
      }
      const int of = totf + incf - nf, on = totn + (incf - incp) - (nf - np);
      totf += __shfl_sync(0xffffffffu, incf, 31);
      totn += __shfl_sync(0xffffffffu, incf - incp, 31);
      fits = fits && totf <= cap;
      if (fits) {
        for (int c = 0; c < nf; c++) colf[of + c] = fm + c;
        if (ts >= 0)
          for (int c = 0; c < nf - np; c++) coln[on + c] = fm + c;
      }
    }
    __syncwarp();
    if (fits) {
      // four rows of the node per sweep over the column list: one shared-memory read and one gather of v per column,
      // four independent loads of a in flight per lane (the rows of a node are neighbours in memory)
      for (int s0 = 0; s0 < nu; s0 += 4) {
        long long base[4];
        int rw[4];
        bool on[4];
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int sr = s0 + r;
          on[r] = sr < nu && sr != ts;  // (the energy row has its own list, below)
          rw[r] = fu + min(sr, nu - 1);
          base[r] = CSR ? rowstart[rw[r]] - msr0 + rw[r] : rowstart[rw[r]];
        }
#pragma unroll 2
        for (int t = lane; t < totf; t += 32) {
          const int col = colf[t];
          const double x = v[col];
#pragma unroll
          for (int r = 0; r < 4; r++) {
            const long long idx = CSR ? base[r] + t : (col == rw[r] ? (long long)rw[r] : base[r] + t - (col > rw[r] ? 1 : 0));
            if (on[r]) acc[r] += a[idx] * x;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
          double sum = acc[r];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (lane == 0 && on[r]) w[rw[r]] = sum;
        }
      }
      if (ts >= 0) {  // energy row: no pressure columns
        const int row = fu + ts;
        const long long base = CSR ? rowstart[row] - msr0 + row : rowstart[row];
        double sum = 0.0;
#pragma unroll 4
        for (int t = lane; t < totn; t += 32) {
          const int col = coln[t];
          const long long idx = CSR ? base + t : (col == row ? (long long)row : base + t - (col > row ? 1 : 0));
          sum += a[idx] * v[col];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) w[row] = sum;
      }
      __syncwarp();
      continue;
    }
    for (int s = 0; s < nu; s++) {
      const int row = fu + s;
      const bool nop = ts >= 0 && s == ts;  // energy row: no pressure columns
      const long long base = CSR ? rowstart[row] - msr0 + row : rowstart[row];
      double sum = 0.0;
      {  // slow path: lanes over the neighbour nodes, offsets by a warp scan per row
        int carry = 0;
        for (long long q0 = b; q0 < e; q0 += 32) {
          const long long q = q0 + lane;
          int fm = 0, ncol = 0;
          if (q < e) {
            const int m = nn_list[q], km = node_kind[m];
            fm = first_unknown[m];
            ncol = K.nunk[km] - (nop ? K.npress[km] : 0);
          }
          int incl = ncol;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          const int off = carry + incl - ncol;
          carry += __shfl_sync(0xffffffffu, incl, 31);
          for (int c = 0; c < ncol; c++) {
            const int col = fm + c;
            const double val = CSR ? a[base + off + c] : (col == row ? a[row] : a[base + off + c - (col > row ? 1 : 0)]);
            sum += val * v[col];
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) w[row] = sum;
    }
    __syncwarp();
  }
}

extern "C" int goma_gpu_matvec(goma_gpu_ctx *c, const double *d_v, double *d_w) {
  if (!c || !d_v || !d_w) return fail(-2, "null argument");
  if (!c->dpat.nn_ptr || !c->dpat.nn_list) return fail(-2, "node-node lists are not resident");
  CU(cudaSetDevice(c->device));
  const int nown = c->prob.num_owned_nodes;
  if (nown > 0) {
    const KindInfo K = make_kind_info(c->prob);
    const long long msr0 = (long long)c->prob.num_unknowns + 1;
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    // staging buffer per warp: the longest column list of this matrix (an energy problem keeps a second list without
    // the pressure unknowns); lists beyond MV_CAP columns take the path without shared memory
    if (c->rss_max_row < 0) {
      if (!c->d_zero_rows) CU(cudaMalloc((void **)&c->d_zero_rows, sizeof(int)));
      CU(cudaMemsetAsync(c->d_zero_rows, 0, sizeof(int), c->stream));
      max_row_len_kernel<<<std::min(1024, (c->num_owned_unknowns + 255) / 256), 256, 0, c->stream>>>(c->num_owned_unknowns, c->d_rowstart, c->d_zero_rows);
      CU(cudaMemcpyAsync(&c->rss_max_row, c->d_zero_rows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      if (c->layout == GOMA_GPU_LAYOUT_CSR) c->rss_max_row += 1;
    }
    const int longest = c->rss_max_row + (c->layout == GOMA_GPU_LAYOUT_CSR ? 0 : 1);  // columns of a row, diagonal included
    int cap = std::min(MV_CAP, (longest + 3) & ~3);
    if (c->matvec_cap > 0) cap = std::min(cap, c->matvec_cap);  // (option "matvec_cap": tests of the slow path)
    const int nlist = c->prob.energy ? 2 : 1;
    const size_t dyn = (size_t)MV_WARPS * nlist * cap * sizeof(int);
    auto launch = [&](auto kern) -> int {
      int per_sm = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, 32 * MV_WARPS, dyn));
      const int blocks = std::max(1, std::min(sms * std::max(per_sm, 1), (nown + MV_WARPS - 1) / MV_WARPS));
      kern<<<blocks, 32 * MV_WARPS, dyn, c->stream>>>(nown, c->dpat.nn_ptr, c->dpat.nn_list, c->d_first, c->d_kind, K, c->d_rowstart, msr0,
                                                       c->d_a, d_v, d_w, cap, nlist);
      return 0;
    };
    if (int lrc = c->layout == GOMA_GPU_LAYOUT_CSR ? launch(node_graph_matvec_kernel<true>) : launch(node_graph_matvec_kernel<false>)) return lrc;
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int goma_gpu_node_graph(goma_gpu_ctx *c, long long **d_nn_ptr, int **d_nn_list) {
  if (!c) return fail(-2, "null context");
  if (!c->dpat.nn_ptr) return fail(-2, "node-node lists are not resident");
  if (d_nn_ptr) *d_nn_ptr = c->dpat.nn_ptr;
  if (d_nn_list) *d_nn_list = c->dpat.nn_list;
  return 0;
}
