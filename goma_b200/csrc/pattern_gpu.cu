// Sparsity skeleton, element ordering and first-touch masks built ON THE DEVICE.
//
// What the reference does serially on the host at start-up, and pattern.cpp restates multi-threaded for the host-only
// entry point goma_gpu_pattern_msr, is done here with the mesh already in HBM:
//   src/exo_conn.c:135-200   build_node_elem      -> node_elem_lists   (radix sort of (node, element) pairs)
//   src/exo_conn.c:204-376   build_node_node      -> node_node_kernel  (one warp per node: gather, bitonic sort,
//                                                    unique; + face-neighbour centroids for centroid nodes :315-347)
//   src/mm_fill_util.c:3229-3445 find_MSR_problem_graph -> cumulative unknown counts along each list + row starts
//                                                    (rows x Inter_Mask columns; energy rows skip pressure columns)
//   element colouring (no counterpart: the reference's loop is serial) -> Jones-Plassmann rounds with hashed
//                                                    priorities, deterministic; border elements (those touching an
//                                                    external node) form their own classes AFTER the interior ones so
//                                                    that the ghost exchange overlaps the interior assembly
//   first-touch masks of the write-once scatter     -> first_touch_kernel (earlier class = earlier writer)
// The per-element column offsets that replace load_lec's in_list search (src/mm_fill.c:5461) are looked up by
// build_records_kernel straight from these lists (binary search), so no per-pair table ever exists.
#include <cub/cub.cuh>

#include <climits>
#include <cstring>

#include "ctx.h"

namespace goma_b200 {

namespace {

template <class T>
int dev_alloc(T **p, size_t n, goma_gpu_ctx *c) {
  CU(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
  if (c) c->device_bytes += n * sizeof(T);
  return 0;
}

__global__ void iota_div_kernel(int *out, long long n, int div) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    out[k] = (int)(k / div);
}
__global__ void count_kernel(const int *__restrict__ conn, long long n, int *__restrict__ cnt) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    atomicAdd(&cnt[conn[k]], 1);
}

__device__ __forceinline__ unsigned hash32(unsigned x) {  // element priority of the colouring (fixed: deterministic)
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

// ---- build_node_node: one warp per node.  PASS 0 counts the distinct neighbours, PASS 1 writes the sorted list,
//      the running unknown counts along it (column offset of each neighbour's first unknown) and the row sums.
template <int CAP, int PASS>
__global__ void __launch_bounds__(128) node_node_kernel(int nn, int npe, int cen, int f0, int fc, const int *__restrict__ conn,
                                                        const int *__restrict__ ne_ptr, const int *__restrict__ ne_list,
                                                        const unsigned char *__restrict__ kind, const KindInfo K, int need_p,
                                                        int *__restrict__ nn_cnt, const long long *__restrict__ nn_ptr,
                                                        int *__restrict__ nn_list, unsigned short *__restrict__ cum_full,
                                                        unsigned short *__restrict__ cum_p, int *__restrict__ row_full,
                                                        int *__restrict__ row_p, int *__restrict__ err) {
  extern __shared__ int sbuf[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  int *buf = sbuf + wib * CAP;
  for (int nd = blockIdx.x * wpb + wib; nd < nn; nd += gridDim.x * wpb) {
    const int q0 = ne_ptr[nd], q1 = ne_ptr[nd + 1];
    int total = (q1 - q0) * npe;
    if (total > CAP - 8) {  // room for the face-neighbour centroids below
      if (lane == 0) atomicMax(err, 1);
      if (PASS == 0 && lane == 0) nn_cnt[nd] = 0;
      continue;
    }
    for (int q = q0; q < q1; q++) {
      const int e = ne_list[q];
      if (lane < npe) buf[(q - q0) * npe + lane] = conn[(size_t)e * npe + lane];
    }
    // a centroid node sees the centroid nodes of the face neighbours of its element (exo_conn.c:315-347)
    if (cen >= 0 && q1 - q0 == 1) {
      const int e = ne_list[q0];
      if (conn[(size_t)e * npe + cen] == nd) {
        if (lane == 0) {
          for (int f = f0; f < f0 + fc; f++) {
            const int fnode = conn[(size_t)e * npe + f];
            for (int q = ne_ptr[fnode]; q < ne_ptr[fnode + 1]; q++)
              if (ne_list[q] != e && total < CAP) buf[total++] = conn[(size_t)ne_list[q] * npe + cen];
          }
        }
        total = __shfl_sync(0xffffffffu, total, 0);
      }
    }
    int S = 32;
    while (S < total) S <<= 1;
    for (int t = total + lane; t < S; t += 32) buf[t] = INT_MAX;
    __syncwarp();
    for (int k = 2; k <= S; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < S; t += 32) {
          const int x = t ^ j;
          if (x > t) {
            const int a = buf[t], b = buf[x];
            const bool up = (t & k) == 0;
            if ((a > b) == up) {
              buf[t] = b;
              buf[x] = a;
            }
          }
        }
        __syncwarp();
      }
    int ucount = 0, cf = 0, cp = 0;
    const long long base_out = PASS == 1 ? nn_ptr[nd] : 0;
    for (int base = 0; base < S; base += 32) {
      const int t = base + lane;
      const int v = buf[t];
      const int prev = t > 0 ? buf[t - 1] : -1;
      const bool flag = v != INT_MAX && v != prev;
      const unsigned bal = __ballot_sync(0xffffffffu, flag);
      if (PASS == 1) {
        const int kd = flag ? kind[v] : 0;
        int wf = flag ? K.nunk[kd] : 0, wp = flag ? K.npress[kd] : 0;
        int sf = wf, sp = wp;  // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int uf = __shfl_up_sync(0xffffffffu, sf, o), up = __shfl_up_sync(0xffffffffu, sp, o);
          if (lane >= o) {
            sf += uf;
            sp += up;
          }
        }
        if (flag) {
          const int pos = ucount + __popc(bal & ((1u << lane) - 1u));
          nn_list[base_out + pos] = v;
          const int off = cf + sf - wf;
          if (off > 65535) atomicMax(err, 2);
          cum_full[base_out + pos] = (unsigned short)off;
          if (need_p) cum_p[base_out + pos] = (unsigned short)(cp + sp - wp);
        }
        cf += __shfl_sync(0xffffffffu, sf, 31);
        cp += __shfl_sync(0xffffffffu, sp, 31);
      }
      ucount += __popc(bal);
    }
    if (lane == 0) {
      if (PASS == 0)
        nn_cnt[nd] = ucount;
      else {
        row_full[nd] = cf;
        row_p[nd] = cp;
      }
    }
    __syncwarp();
  }
}

// ---- find_MSR_problem_graph: length of the rows of one node, then (after a scan over the nodes) the row starts
__global__ void node_row_len_kernel(int nn, const unsigned char *__restrict__ kind, const KindInfo K,
                                    const int *__restrict__ row_full, const int *__restrict__ row_p,
                                    long long *__restrict__ node_len) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= nn) return;
  const int kd = kind[nd];
  long long len = (long long)K.nunk[kd] * (row_full[nd] - 1);
  if (K.tslot[kd] >= 0) len -= row_p[nd];
  node_len[nd] = len;
}
__global__ void rowstart_kernel(int nn, int N, const unsigned char *__restrict__ kind, const KindInfo K,
                                const int *__restrict__ first_unknown, const int *__restrict__ row_full,
                                const int *__restrict__ row_p, const long long *__restrict__ node_off,
                                const long long *__restrict__ node_len, long long *__restrict__ rowstart) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= nn) return;
  const int kd = kind[nd], fu = first_unknown[nd];
  long long pos = (long long)N + 1 + node_off[nd];
  for (int s = 0; s < K.nunk[kd]; s++) {
    rowstart[fu + s] = pos;
    pos += (s == K.tslot[kd] && K.tslot[kd] >= 0) ? row_full[nd] - row_p[nd] - 1 : row_full[nd] - 1;
  }
  if (nd == nn - 1) rowstart[N] = (long long)N + 1 + node_off[nd] + node_len[nd];
}

// ---- colouring: one round.  An element takes the smallest colour none of its coloured neighbours holds once every
//      neighbour that precedes it is coloured.  BY_ID: "precedes" = lower element number, which reproduces the serial
//      greedy colouring in element order exactly (8 colours on a structured hex mesh, 4 on quads) -- the number of
//      rounds is the depth of that dependence (a diagonal wavefront, nx + 2 ny + 4 nz on a lattice).  Otherwise
//      "precedes" = higher hashed priority (Jones-Plassmann, O(log n) rounds): the finish for meshes whose numbering
//      makes the greedy dependence too deep.  In place: a neighbour that must come later can never be coloured
//      before this element, so the colours an element sees are final and the result does not depend on timing.
template <bool BY_ID>
__global__ void colour_round_kernel(int ne, int npe, const int *__restrict__ conn, const int *__restrict__ ne_ptr,
                                    const int *__restrict__ ne_list, volatile signed char *col, int *__restrict__ remaining,
                                    int count, int *__restrict__ err) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne || col[e] >= 0) return;
  const unsigned pe = hash32((unsigned)e);
  unsigned long long used = 0ull;
  for (int i = 0; i < npe; i++) {
    const int nd = conn[(size_t)e * npe + i];
    for (int q = ne_ptr[nd]; q < ne_ptr[nd + 1]; q++) {
      const int e2 = ne_list[q];
      if (e2 == e) continue;
      const int c2 = col[e2];
      if (c2 >= 0) {
        used |= 1ull << c2;
        continue;
      }
      bool before;
      if (BY_ID)
        before = e2 < e;
      else {
        const unsigned p2 = hash32((unsigned)e2);
        before = p2 > pe || (p2 == pe && e2 > e);
      }
      if (before) {  // wait for it
        if (count) *remaining = 1;  // (the host only asks whether anybody is still waiting)
        return;
      }
    }
  }
  int c = 0;
  while (c < 63 && ((used >> c) & 1ull)) c++;
  if (c >= 63) atomicMax(err, 3);
  col[e] = (signed char)c;
}

// class of an element = colour, or colour + ncol when it touches an external node (needs the ghost exchange)
__global__ void max_colour_kernel(int ne, const signed char *__restrict__ col, int *__restrict__ maxc) {
  int m = -1;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) m = max(m, (int)col[e]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(maxc, m);
}
__global__ void class_kernel(int ne, int npe, int ncol, int nmat, const int *__restrict__ elem_mat, int nchunk, int chunk_elems,
                             int num_owned_nodes, int split_border, const int *__restrict__ conn,
                             const signed char *__restrict__ col, unsigned *__restrict__ cls, int *__restrict__ ids,
                             int *__restrict__ hist) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  bool border = false;
  if (split_border)
    for (int i = 0; i < npe; i++) border = border || conn[(size_t)e * npe + i] >= num_owned_nodes;
  // processing order: interior elements before border elements (the ghost exchange overlaps the former); inside each,
  // chunks of consecutive elements whose matrix rows fit the L2 cache, one colour after the other
  // ... and inside a colour one material after the other: a launch covers one class, so it runs with the constants
  // of ONE material (mp_glob[Matilda[ebn]], mm_fill.c:224-235) in its kernel parameters
  const int k = (((border ? nchunk : 0) + e / chunk_elems) * ncol + col[e]) * nmat + (elem_mat ? elem_mat[e] : 0);
  cls[e] = (unsigned)k;
  ids[e] = e;
  atomicAdd(&hist[k], 1);
}

// ---- first-touch masks: the pair (i, j) of element e is NOT a first touch iff an element of a lower class holds
//      both nodes (elements of one class share no node: classes are colours)
__global__ void first_touch_kernel(int ne, int npe, const int *__restrict__ conn, const int *__restrict__ ne_ptr,
                                   const int *__restrict__ ne_list, const unsigned *__restrict__ cls,
                                   unsigned *__restrict__ pair_first, unsigned *__restrict__ node_first,
                                   int *__restrict__ err) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  constexpr int MAXNB = 96;
  int ids[MAXNB];
  unsigned masks[MAXNB];
  int nnb = 0;
  const unsigned ce = cls[e];
  for (int i = 0; i < npe; i++) {
    const int nd = conn[(size_t)e * npe + i];
    for (int q = ne_ptr[nd]; q < ne_ptr[nd + 1]; q++) {
      const int e2 = ne_list[q];
      if (e2 == e || cls[e2] >= ce) continue;
      int k = 0;
      while (k < nnb && ids[k] != e2) k++;
      if (k == nnb) {
        if (nnb == MAXNB) {
          atomicMax(err, 4);
          continue;
        }
        ids[nnb] = e2;
        masks[nnb] = 0u;
        nnb++;
      }
      masks[k] |= 1u << i;
    }
  }
  unsigned notfirst[32];
  for (int i = 0; i < npe; i++) notfirst[i] = 0u;
  unsigned node_nf = 0u;
  for (int k = 0; k < nnb; k++) {
    const unsigned sh = masks[k];
    node_nf |= sh;
    for (int i = 0; i < npe; i++)
      if ((sh >> i) & 1u) notfirst[i] |= sh;
  }
  const unsigned all = npe == 32 ? 0xffffffffu : ((1u << npe) - 1u);
  for (int i = 0; i < npe; i++) pair_first[(size_t)e * npe + i] = all & ~notfirst[i];
  node_first[e] = all & ~node_nf;
}

int grid_for(long long n, int threads, int cap = 148 * 16) {
  return (int)std::max<long long>(1, std::min<long long>(cap, (n + threads - 1) / threads));
}

}  // namespace

KindInfo make_kind_info(const goma_gpu_problem &p) {
  KindInfo K;
  memset(&K, 0, sizeof(K));
  for (int k = 0; k < GOMA_GPU_MAX_KINDS; k++) {
    K.tslot[k] = -1;
    if (k >= p.num_kinds) continue;
    K.nunk[k] = p.kind_num_unknowns[k];
    K.npress[k] = p.kind_slot[k][GOMA_SLOT_P] < 0 ? 0 : (p.pressure_interp == GOMA_PRESSURE_P1 ? p.dim + 1 : 1);
    K.tslot[k] = p.energy ? p.kind_slot[k][GOMA_SLOT_T] : -1;
  }
  return K;
}

void free_device_pattern(goma_gpu_ctx *c, bool keep_node_node) {
  DevPattern &d = c->dpat;
  void *drop[] = {d.ne_ptr, d.ne_list, d.cum_full, d.cum_p, d.row_full, d.row_p, d.pair_first, d.node_first, d.cls};
  for (void *q : drop)
    if (q) cudaFree(q);
  d.ne_ptr = d.ne_list = d.row_full = d.row_p = nullptr;
  d.cum_full = d.cum_p = nullptr;
  d.pair_first = d.node_first = nullptr;
  d.cls = nullptr;
  if (!keep_node_node) {
    if (d.nn_ptr) cudaFree(d.nn_ptr);
    if (d.nn_list) cudaFree(d.nn_list);
    d.nn_ptr = nullptr;
    d.nn_list = nullptr;
  }
}

// Everything goma_gpu_fill_init needs from the mesh topology, from the arrays already uploaded to the context
// (d_conn, d_kind, d_first).  On return: c->dpat (lists), c->d_rowstart, c->nnz_plus, c->d_elem_list,
// c->colour_begin (class boundaries) and c->first_border_class.
int build_pattern_device(goma_gpu_ctx *c) {
  const goma_gpu_problem &p = c->prob;
  const int nn = p.num_nodes, ne = p.num_elems, npe = p.elem_type, N = p.num_unknowns;
  const long long nconn = (long long)ne * npe;
  if (nconn > 2147483647LL) return fail(-2, "more than 2^31 connectivity entries");
  DevPattern &d = c->dpat;
  cudaStream_t st = c->stream;
  const KindInfo K = make_kind_info(p);
  int *d_err = nullptr;
  CU(cudaMalloc((void **)&d_err, 4 * sizeof(int)));
  CU(cudaMemsetAsync(d_err, 0, 4 * sizeof(int), st));
  void *d_temp = nullptr;
  size_t temp_bytes = 0;
  auto temp = [&](size_t need) -> int {
    if (need > temp_bytes) {
      if (d_temp) cudaFree(d_temp);
      d_temp = nullptr;
      CU(cudaMalloc(&d_temp, need));
      temp_bytes = need;
    }
    return 0;
  };
  struct Cleanup {
    void **t;
    int **e;
    ~Cleanup() {
      if (*t) cudaFree(*t);
      if (*e) cudaFree(*e);
    }
  } cleanup{&d_temp, &d_err};

  // ---- node -> elements: stable radix sort of the (node, element) pairs keeps the elements of a node ascending
  if (dev_alloc(&d.ne_ptr, (size_t)nn + 1, c) || dev_alloc(&d.ne_list, (size_t)nconn, c)) return -3;
  {
    int *d_cnt = nullptr, *d_elem = nullptr, *d_keys_out = nullptr;
    CU(cudaMalloc((void **)&d_cnt, ((size_t)nn + 1) * sizeof(int)));
    CU(cudaMalloc((void **)&d_elem, std::max<size_t>(nconn, 1) * sizeof(int)));
    CU(cudaMalloc((void **)&d_keys_out, std::max<size_t>(nconn, 1) * sizeof(int)));
    CU(cudaMemsetAsync(d_cnt, 0, ((size_t)nn + 1) * sizeof(int), st));
    if (nconn > 0) {
      count_kernel<<<grid_for(nconn, 256), 256, 0, st>>>(c->d_conn, nconn, d_cnt);
      iota_div_kernel<<<grid_for(nconn, 256), 256, 0, st>>>(d_elem, nconn, npe);
    }
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_cnt, d.ne_ptr, nn + 1, st);
    if (temp(need)) return -3;
    cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_cnt, d.ne_ptr, nn + 1, st);
    if (nconn > 0) {
      int bits = 1;
      while ((1LL << bits) < (long long)std::max(nn, 2)) bits++;
      need = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, need, c->d_conn, d_keys_out, d_elem, d.ne_list, (int)nconn, 0, bits, st);
      if (temp(need)) return -3;
      size_t tb = temp_bytes;
      cub::DeviceRadixSort::SortPairs(d_temp, tb, c->d_conn, d_keys_out, d_elem, d.ne_list, (int)nconn, 0, bits, st);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    cudaFree(d_cnt);
    cudaFree(d_elem);
    cudaFree(d_keys_out);
  }

  // ---- node-node lists
  const int cen = npe == GOMA_GPU_QUAD9 ? 8 : (npe == GOMA_GPU_HEX27 ? 20 : -1);
  const int f0 = npe == GOMA_GPU_QUAD9 ? 4 : 21, fc = npe == GOMA_GPU_QUAD9 ? 4 : (npe == GOMA_GPU_HEX27 ? 6 : 0);
  const int need_p = p.energy ? 1 : 0;
  int *d_nn_cnt = nullptr;
  CU(cudaMalloc((void **)&d_nn_cnt, ((size_t)nn + 1) * sizeof(int)));
  CU(cudaMemsetAsync(d_nn_cnt, 0, ((size_t)nn + 1) * sizeof(int), st));
  if (dev_alloc(&d.nn_ptr, (size_t)nn + 1, c) || dev_alloc(&d.row_full, (size_t)nn, c) || dev_alloc(&d.row_p, (size_t)nn, c)) return -3;
  constexpr int CAP = 1024;  // candidates per node: up to 37 hex27 / 127 hex8 elements round one node
  const int nn_grid = grid_for((long long)nn * 32, 128, 148 * 12);
  const size_t nn_smem = 4 * CAP * sizeof(int);
  if (nn > 0) node_node_kernel<CAP, 0><<<nn_grid, 128, nn_smem, st>>>(nn, npe, cen, f0, fc, c->d_conn, d.ne_ptr, d.ne_list, c->d_kind, K, need_p,
                                                                      d_nn_cnt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, d_err);
  {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, d_nn_cnt, d.nn_ptr, nn + 1, st);
    if (temp(need)) return -3;
    cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_nn_cnt, d.nn_ptr, nn + 1, st);
  }
  long long nn_total = 0;
  CU(cudaMemcpyAsync(&nn_total, d.nn_ptr + nn, sizeof(long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  cudaFree(d_nn_cnt);
  d.nn_total = nn_total;
  if (dev_alloc(&d.nn_list, (size_t)nn_total, c) || dev_alloc(&d.cum_full, (size_t)nn_total, c)) return -3;
  if (need_p && dev_alloc(&d.cum_p, (size_t)nn_total, c)) return -3;
  if (nn > 0) node_node_kernel<CAP, 1><<<nn_grid, 128, nn_smem, st>>>(nn, npe, cen, f0, fc, c->d_conn, d.ne_ptr, d.ne_list, c->d_kind, K, need_p,
                                                                      nullptr, d.nn_ptr, d.nn_list, d.cum_full, d.cum_p, d.row_full, d.row_p, d_err);
  CU(cudaGetLastError());

  // ---- row starts (== ija[0..N] of the MSR graph, 64-bit)
  if (dev_alloc(&c->d_rowstart, (size_t)N + 1, c)) return -3;
  {
    long long *d_len = nullptr, *d_off = nullptr;
    CU(cudaMalloc((void **)&d_len, std::max<size_t>(nn, 1) * sizeof(long long)));
    CU(cudaMalloc((void **)&d_off, std::max<size_t>(nn, 1) * sizeof(long long)));
    long long h_start_only = (long long)N + 1;
    if (nn > 0) {
      node_row_len_kernel<<<(nn + 255) / 256, 256, 0, st>>>(nn, c->d_kind, K, d.row_full, d.row_p, d_len);
      size_t need = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, need, d_len, d_off, nn, st);
      if (temp(need)) return -3;
      cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_len, d_off, nn, st);
      rowstart_kernel<<<(nn + 255) / 256, 256, 0, st>>>(nn, N, c->d_kind, K, c->d_first, d.row_full, d.row_p, d_off, d_len, c->d_rowstart);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(&c->nnz_plus, c->d_rowstart + N, sizeof(long long), cudaMemcpyDeviceToHost, st));
    } else {
      CU(cudaMemcpyAsync(c->d_rowstart, &h_start_only, sizeof(long long), cudaMemcpyHostToDevice, st));
      c->nnz_plus = h_start_only;
    }
    CU(cudaStreamSynchronize(st));
    cudaFree(d_len);
    cudaFree(d_off);
  }

  // ---- colouring, classes, processing order
  signed char *d_colour = nullptr;
  int *d_counters = nullptr;  // [0] remaining, [1] max colour
  CU(cudaMalloc((void **)&d_colour, std::max(ne, 1)));
  CU(cudaMalloc((void **)&d_counters, 2 * sizeof(int)));
  CU(cudaMemsetAsync(d_colour, 0xff, std::max(ne, 1), st));
  {
    // rounds are enqueued in groups; the host looks at the count of waiting elements once per group
    const int group = 64, greedy_rounds = 4096, max_rounds = 4096 + 4096;
    int rounds = 0, remaining = ne > 0 ? 1 : 0;
    while (remaining > 0) {
      if (rounds >= max_rounds) return fail(-3, "element colouring did not converge");
      const bool by_id = rounds < greedy_rounds;
      for (int k = 0; k < group; k++, rounds++) {
        if (k == group - 1) CU(cudaMemsetAsync(d_counters, 0, sizeof(int), st));
        if (by_id)
          colour_round_kernel<true><<<(ne + 127) / 128, 128, 0, st>>>(ne, npe, c->d_conn, d.ne_ptr, d.ne_list, d_colour, d_counters,
                                                                      k == group - 1, d_err);
        else
          colour_round_kernel<false><<<(ne + 127) / 128, 128, 0, st>>>(ne, npe, c->d_conn, d.ne_ptr, d.ne_list, d_colour, d_counters,
                                                                       k == group - 1, d_err);
      }
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(&remaining, d_counters, sizeof(int), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
    }
  }
  int maxc = -1;
  {
    int init = -1;
    CU(cudaMemcpyAsync(d_counters + 1, &init, sizeof(int), cudaMemcpyHostToDevice, st));
    if (ne > 0) max_colour_kernel<<<grid_for(ne, 256), 256, 0, st>>>(ne, d_colour, d_counters + 1);
    CU(cudaMemcpyAsync(&maxc, d_counters + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  const int ncol = maxc + 1;
  const int split_border = p.num_owned_nodes < p.num_nodes ? 1 : 0;
  // Chunks: when many elements contribute to the same matrix slots (hex8 / quad4: 2-8 elements per node pair), a
  // colour sweep over the whole mesh reads and writes every shared sector once per contribution from DRAM.  Sweeping
  // the colours chunk by chunk, with the rows of a chunk small enough to stay in L2, lets the contributions meet there.
  // Only worth it while a (chunk, colour) launch still fills the GPU: not for hex27 (45-75 KB of matrix per element).
  int chunk_elems = std::max(ne, 1);
  {
    int l2 = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, c->device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const double bytes_per_elem = ne > 0 ? 8.0 * (double)c->nnz_plus / ne : 1.0;
    long long want = c->chunk_elems_option > 0 ? c->chunk_elems_option : (long long)(0.5 * l2 / bytes_per_elem);
    const bool many_writers = npe == GOMA_GPU_HEX8 || npe == GOMA_GPU_QUAD4;
    // measured (profiles/r2d_c5_chunks.txt): on B200 the extra launches cost more than the L2 hits save -- off unless asked for
    const bool on = c->chunk_elems_option > 0;
    (void)many_writers;
    (void)sms;
    if (on && want < ne) chunk_elems = (int)std::max<long long>(want, 1);
  }
  const int nchunk = ne > 0 ? (ne + chunk_elems - 1) / chunk_elems : 1;
  const int nmat = std::max(1, c->num_mats);
  const long long ncls_ll = (long long)ncol * nmat * nchunk * (split_border ? 2 : 1);
  if (ncol > 63) return fail(-2, "element colouring needs more than 63 colours");
  if (ncls_ll > (1LL << 24)) return fail(-2, "too many (chunk, colour) classes");
  const int ncls = (int)ncls_ll;
  int *d_ids = nullptr, *d_hist = nullptr;
  if (dev_alloc(&d.cls, (size_t)ne, c)) return -3;
  CU(cudaMalloc((void **)&d_ids, std::max(ne, 1) * sizeof(int)));
  CU(cudaMalloc((void **)&d_hist, std::max(ncls, 1) * sizeof(int)));
  CU(cudaMemsetAsync(d_hist, 0, std::max(ncls, 1) * sizeof(int), st));
  if (dev_alloc(&c->d_elem_list, (size_t)ne, c)) return -3;
  std::vector<int> hist(std::max(ncls, 1), 0);
  if (ne > 0) {
    class_kernel<<<(ne + 255) / 256, 256, 0, st>>>(ne, npe, ncol, nmat, c->d_elem_mat, nchunk, chunk_elems, p.num_owned_nodes, split_border, c->d_conn, d_colour,
                                                   d.cls, d_ids, d_hist);
    unsigned *d_cls_out = nullptr;
    CU(cudaMalloc((void **)&d_cls_out, (size_t)ne * sizeof(unsigned)));
    int bits = 1;
    while ((1LL << bits) < std::max(ncls, 2)) bits++;
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, d.cls, d_cls_out, d_ids, c->d_elem_list, ne, 0, bits, st);
    if (temp(need)) return -3;
    size_t tb = temp_bytes;
    cub::DeviceRadixSort::SortPairs(d_temp, tb, d.cls, d_cls_out, d_ids, c->d_elem_list, ne, 0, bits, st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hist.data(), d_hist, (size_t)ncls * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_cls_out);
  }
  cudaFree(d_hist);
  c->colour_begin.assign(ncls + 1, 0);
  for (int k = 0; k < ncls; k++) c->colour_begin[k + 1] = c->colour_begin[k] + hist[k];
  c->first_border_class = split_border ? ncol * nmat * nchunk : ncls;
  c->num_colours = ncol * nmat;  // classes per chunk: (colour, material)
  c->num_chunks = nchunk;
  c->chunk_elems = chunk_elems;

  // ---- first-touch masks
  if (npe > 32) return fail(-2, "first-touch masks need <= 32 nodes per element");
  if (dev_alloc(&d.pair_first, (size_t)nconn, c) || dev_alloc(&d.node_first, (size_t)ne, c)) return -3;
  if (ne > 0) first_touch_kernel<<<(ne + 63) / 64, 64, 0, st>>>(ne, npe, c->d_conn, d.ne_ptr, d.ne_list, d.cls, d.pair_first, d.node_first, d_err);
  CU(cudaGetLastError());
  int h_err[4] = {0, 0, 0, 0};
  CU(cudaMemcpyAsync(h_err, d_err, sizeof(h_err), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  cudaFree(d_colour);
  cudaFree(d_counters);
  cudaFree(d_ids);
  switch (h_err[0]) {
    case 0: break;
    case 1: return fail(-2, "sparsity pattern: more than 1016 candidate neighbours round one node");
    case 2: return fail(-2, "sparsity pattern: row longer than 65535 columns");
    case 3: return fail(-2, "element colouring needs more than 63 colours");
    default: return fail(-2, "first-touch masks: more than 96 lower-class neighbours of one element");
  }
  return 0;
}

// host copy of the node-node lists and row starts (export / validation at test sizes)
int download_pattern(goma_gpu_ctx *c, Pattern &out) {
  const goma_gpu_problem &p = c->prob;
  const DevPattern &d = c->dpat;
  if (!d.nn_ptr || !d.nn_list) return fail(-2, "node-node lists are not resident any more");
  out.num_nodes = p.num_nodes;
  out.num_unknowns = p.num_unknowns;
  out.npe = p.elem_type;
  out.nn_ptr.resize((size_t)p.num_nodes + 1);
  out.nn_list.resize((size_t)d.nn_total);
  out.rowstart.resize((size_t)p.num_unknowns + 1);
  static_assert(sizeof(long long) == sizeof(int64_t), "64-bit");
  CU(cudaMemcpy(out.nn_ptr.data(), d.nn_ptr, out.nn_ptr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  if (d.nn_total) CU(cudaMemcpy(out.nn_list.data(), d.nn_list, out.nn_list.size() * sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out.rowstart.data(), c->d_rowstart, out.rowstart.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  out.nnz_plus = c->nnz_plus;
  return 0;
}

}  // namespace goma_b200
