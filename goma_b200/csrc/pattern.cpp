// See pattern.h.  Pure host C++ (no CUDA): runs once per problem in goma_gpu_fill_init.
#include "pattern.h"

#include <algorithm>
#include <cstring>
#include <thread>

namespace goma_b200 {

namespace {

int num_chunks(int64_t n, int num_threads) {
  int64_t want = std::min<int64_t>(num_threads, (n + 4095) / 4096);
  return (int)std::max<int64_t>(1, want);
}

template <class F>
void parallel_chunks(int64_t n, int num_threads, F f) {
  int T = num_chunks(n, num_threads);
  if (T == 1) {
    f(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) {
    int64_t lo = n * t / T, hi = n * (t + 1) / T;
    th.emplace_back([=] { f(t, lo, hi); });
  }
  for (auto &x : th) x.join();
}

// local index of the centroid node (reference: el_elm_info.c:4095 centroid_node)
int centroid_local(int elem_type) {
  if (elem_type == GOMA_GPU_QUAD9) return 8;
  if (elem_type == GOMA_GPU_HEX27) return 20;
  return -1;
}
// local indices of the face-centre nodes: the other element holding one is the face neighbour
void face_nodes(int elem_type, int &first, int &count) {
  if (elem_type == GOMA_GPU_QUAD9) {
    first = 4;
    count = 4;
  } else if (elem_type == GOMA_GPU_HEX27) {
    first = 21;
    count = 6;
  } else {
    first = 0;
    count = 0;
  }
}

}  // namespace

int kind_num_pressure(const goma_gpu_problem &p, int kind) {
  if (p.kind_slot[kind][GOMA_SLOT_P] < 0) return 0;
  return p.pressure_interp == GOMA_PRESSURE_P1 ? p.dim + 1 : 1;
}

std::string build_pattern(const goma_gpu_problem &p, Pattern &out, int num_threads) {
  const int nn = p.num_nodes, ne = p.num_elems, npe = p.elem_type;
  const int *conn = p.elem_connect;
  out.num_nodes = nn;
  out.num_unknowns = p.num_unknowns;
  out.npe = npe;

  // ---- node -> element adjacency (counting sort) : exo_conn.c build_node_elem
  std::vector<int64_t> ne_ptr(nn + 1, 0);
  for (int64_t k = 0; k < (int64_t)ne * npe; k++) {
    int nd = conn[k];
    if (nd < 0 || nd >= nn) return "element connectivity entry out of range";
    ne_ptr[nd + 1]++;
  }
  for (int i = 0; i < nn; i++) ne_ptr[i + 1] += ne_ptr[i];
  std::vector<int> ne_list(ne_ptr[nn]);
  {
    std::vector<int64_t> fill(ne_ptr.begin(), ne_ptr.end() - 1);
    for (int e = 0; e < ne; e++)
      for (int k = 0; k < npe; k++) ne_list[fill[conn[(int64_t)e * npe + k]]++] = e;
  }

  // ---- node-node lists, sorted & unique (+ face-neighbour centroids for centroid nodes)
  const int cen = centroid_local(p.elem_type);
  int f0, fc;
  face_nodes(p.elem_type, f0, fc);
  int T = std::max(1, num_threads);
  std::vector<std::vector<int>> chunk_list(T);
  std::vector<std::vector<int>> chunk_cnt(T);
  std::vector<int64_t> chunk_lo(T, 0), chunk_hi(T, 0);
  const int used_threads = num_chunks(nn, T);
  parallel_chunks(nn, T, [&](int t, int64_t lo, int64_t hi) {
    chunk_lo[t] = lo;
    chunk_hi[t] = hi;
    std::vector<int> &L = chunk_list[t];
    std::vector<int> &C = chunk_cnt[t];
    C.resize(hi - lo);
    std::vector<int> buf;
    for (int64_t nd = lo; nd < hi; nd++) {
      buf.clear();
      for (int64_t q = ne_ptr[nd]; q < ne_ptr[nd + 1]; q++) {
        const int *c = conn + (int64_t)ne_list[q] * npe;
        buf.insert(buf.end(), c, c + npe);
      }
      if (cen >= 0 && ne_ptr[nd + 1] - ne_ptr[nd] == 1) {
        int e = ne_list[ne_ptr[nd]];
        if (conn[(int64_t)e * npe + cen] == nd) {
          for (int f = f0; f < f0 + fc; f++) {
            int fnode = conn[(int64_t)e * npe + f];
            for (int64_t q = ne_ptr[fnode]; q < ne_ptr[fnode + 1]; q++)
              if (ne_list[q] != e) buf.push_back(conn[(int64_t)ne_list[q] * npe + cen]);
          }
        }
      }
      std::sort(buf.begin(), buf.end());
      buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
      C[nd - lo] = (int)buf.size();
      L.insert(L.end(), buf.begin(), buf.end());
    }
  });
  out.nn_ptr.assign(nn + 1, 0);
  for (int t = 0; t < used_threads; t++)
    for (int64_t nd = chunk_lo[t]; nd < chunk_hi[t]; nd++) out.nn_ptr[nd + 1] = chunk_cnt[t][nd - chunk_lo[t]];
  for (int i = 0; i < nn; i++) out.nn_ptr[i + 1] += out.nn_ptr[i];
  out.nn_list.resize(out.nn_ptr[nn]);
  for (int t = 0; t < used_threads; t++) {
    if (!chunk_list[t].empty())
      std::memcpy(out.nn_list.data() + out.nn_ptr[chunk_lo[t]], chunk_list[t].data(),
                  chunk_list[t].size() * sizeof(int));
    std::vector<int>().swap(chunk_list[t]);
  }

  // ---- cumulative unknown counts along each list, row starts (find_MSR_problem_graph)
  int kind_np[GOMA_GPU_MAX_KINDS];
  for (int k = 0; k < p.num_kinds; k++) kind_np[k] = kind_num_pressure(p, k);
  out.need_pair_p = p.energy != 0;  // only energy rows mask out pressure columns
  out.cum_full.resize(out.nn_list.size());
  out.cum_p.resize(out.need_pair_p ? out.nn_list.size() : 0);
  std::vector<int> row_full(nn), row_p(nn);
  bool overflow = false;
  parallel_chunks(nn, T, [&](int, int64_t lo, int64_t hi) {
    for (int64_t nd = lo; nd < hi; nd++) {
      int cf = 0, cp = 0;
      for (int64_t q = out.nn_ptr[nd]; q < out.nn_ptr[nd + 1]; q++) {
        int kd = p.node_kind[out.nn_list[q]];
        if (cf > 65535) overflow = true;
        out.cum_full[q] = (uint16_t)cf;
        if (out.need_pair_p) out.cum_p[q] = (uint16_t)cp;
        cf += p.kind_num_unknowns[kd];
        cp += kind_np[kd];
      }
      row_full[nd] = cf;
      row_p[nd] = cp;
    }
  });
  if (overflow) return "row longer than 65535 columns";
  const int N = p.num_unknowns;
  out.rowstart.assign(N + 1, 0);
  int64_t pos = (int64_t)N + 1;
  for (int nd = 0; nd < nn; nd++) {
    int kd = p.node_kind[nd];
    int fu = p.first_unknown[nd];
    int tslot = p.kind_slot[kd][GOMA_SLOT_T];
    for (int s = 0; s < p.kind_num_unknowns[kd]; s++) {
      if (fu + s >= N) return "first_unknown inconsistent with num_unknowns";
      out.rowstart[fu + s] = pos;
      pos += (s == tslot && tslot >= 0) ? row_full[nd] - row_p[nd] - 1 : row_full[nd] - 1;
    }
  }
  out.rowstart[N] = pos;
  out.nnz_plus = pos;

  // ---- per element pair offsets (replaces load_lec's in_list search, mm_fill.c:5461)
  const int64_t npairs = (int64_t)ne * npe * npe;
  out.pair_full.resize(npairs);
  out.pair_p.resize(out.need_pair_p ? npairs : 0);
  bool missing = false;
  parallel_chunks(ne, T, [&](int, int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; e++) {
      const int *c = conn + e * npe;
      for (int i = 0; i < npe; i++) {
        const int *b = out.nn_list.data() + out.nn_ptr[c[i]];
        const int *en = out.nn_list.data() + out.nn_ptr[c[i] + 1];
        for (int j = 0; j < npe; j++) {
          const int *it = std::lower_bound(b, en, c[j]);
          if (it == en || *it != c[j]) {
            missing = true;
            continue;
          }
          int64_t q = it - out.nn_list.data();
          out.pair_full[(e * npe + i) * npe + j] = out.cum_full[q];
          if (out.need_pair_p) out.pair_p[(e * npe + i) * npe + j] = out.cum_p[q];
        }
      }
    }
  });
  if (missing) return "Could not find vbl in sparse matrix";  // mm_fill.c:5462 wording

  // ---- greedy colouring in element order (structured hex meshes: 8 colours, quads: 4)
  std::vector<int> colour(ne);
  {
    std::vector<unsigned long long> node_mask(nn, 0ull);
    int ncol = 0;
    for (int e = 0; e < ne; e++) {
      unsigned long long used = 0;
      for (int k = 0; k < npe; k++) used |= node_mask[conn[(int64_t)e * npe + k]];
      int c = 0;
      while (c < 63 && ((used >> c) & 1ull)) c++;
      if (c >= 63) return "element colouring needs more than 63 colours";
      colour[e] = c;
      ncol = std::max(ncol, c + 1);
      for (int k = 0; k < npe; k++) node_mask[conn[(int64_t)e * npe + k]] |= 1ull << c;
    }
    out.colour_begin.assign(ncol + 1, 0);
    for (int e = 0; e < ne; e++) out.colour_begin[colour[e] + 1]++;
    for (int c = 0; c < ncol; c++) out.colour_begin[c + 1] += out.colour_begin[c];
    out.colour_order.resize(ne);
    std::vector<int> fill(out.colour_begin.begin(), out.colour_begin.end() - 1);
    for (int e = 0; e < ne; e++) out.colour_order[fill[colour[e]]++] = e;
  }
  // ---- first-touch masks: pair (i,j) of e is NOT first iff an element of lower colour holds both nodes
  if (npe > 32) return "first-touch masks need <= 32 nodes per element";
  out.pair_first.assign((size_t)ne * npe, 0u);
  out.node_first.assign(ne, 0u);
  parallel_chunks(ne, T, [&](int, int64_t lo, int64_t hi) {
    std::vector<int> nb;
    for (int64_t e = lo; e < hi; e++) {
      const int *c = conn + e * npe;
      nb.clear();
      for (int i = 0; i < npe; i++)
        for (int64_t q = ne_ptr[c[i]]; q < ne_ptr[c[i] + 1]; q++)
          if (colour[ne_list[q]] < colour[e]) nb.push_back(ne_list[q]);
      std::sort(nb.begin(), nb.end());
      nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
      uint32_t notfirst[32] = {0};
      uint32_t node_nf = 0;
      for (int e2 : nb) {
        uint32_t shared = 0;  // local nodes of e that e2 also holds
        for (int i = 0; i < npe; i++)
          for (int64_t q = ne_ptr[c[i]]; q < ne_ptr[c[i] + 1]; q++)
            if (ne_list[q] == e2) shared |= 1u << i;
        node_nf |= shared;
        for (int i = 0; i < npe; i++)
          if ((shared >> i) & 1u) notfirst[i] |= shared;
      }
      const uint32_t all = npe == 32 ? 0xffffffffu : ((1u << npe) - 1u);
      for (int i = 0; i < npe; i++) out.pair_first[e * npe + i] = all & ~notfirst[i];
      out.node_first[e] = all & ~node_nf;
    }
  });
  return "";
}

void emit_msr_columns(const goma_gpu_problem &p, const Pattern &pat, int *ija) {
  const int N = p.num_unknowns;
  for (int r = 0; r <= N; r++) ija[r] = (int)pat.rowstart[r];
  for (int nd = 0; nd < p.num_nodes; nd++) {
    int kd = p.node_kind[nd];
    int fu = p.first_unknown[nd];
    int tslot = p.kind_slot[kd][GOMA_SLOT_T];
    for (int s = 0; s < p.kind_num_unknowns[kd]; s++) {
      int64_t pos = pat.rowstart[fu + s];
      bool nop = (s == tslot && tslot >= 0);
      for (int64_t q = pat.nn_ptr[nd]; q < pat.nn_ptr[nd + 1]; q++) {
        int m = pat.nn_list[q];
        int km = p.node_kind[m];
        int ncol = p.kind_num_unknowns[km] - (nop ? kind_num_pressure(p, km) : 0);
        for (int c = 0; c < ncol; c++) {
          int col = p.first_unknown[m] + c;
          if (col != fu + s) ija[pos++] = col;
        }
      }
    }
  }
}

}  // namespace goma_b200
