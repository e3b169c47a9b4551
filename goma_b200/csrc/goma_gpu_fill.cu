// C ABI of the sm_100a matrix_fill path (see include/goma_gpu_fill.h for the contract and
// the reference call sites each entry point replaces).  No torch, no CPU fallback: every
// failure is an error code plus goma_gpu_last_error().
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/goma_gpu_fill.h"
#include "fill_kernel.cuh"
#include "pattern.h"
#include "tables.h"

using namespace goma_b200;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(-3, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                          std::to_string(__LINE__));                                           \
  } while (0)

struct goma_gpu_ctx {
  goma_gpu_problem prob;  // scalar members + kind tables only; pointers are not retained
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Pattern pat;  // host copy kept for get_msr (nn lists) -- sized for test/bench meshes
  // device arrays
  int *d_conn = nullptr, *d_first = nullptr;
  double *d_coord[3] = {nullptr, nullptr, nullptr};
  unsigned char *d_kind = nullptr, *d_dbc_flag = nullptr;
  double *d_dbc_value = nullptr;
  long long *d_rowstart = nullptr;
  unsigned short *d_pair_full = nullptr, *d_pair_p = nullptr;
  unsigned *d_pair_first = nullptr, *d_node_first = nullptr;
  double *d_tables = nullptr;
  unsigned char *d_erec = nullptr;  // per-element gather records (ElemRec<C>)
  double *d_x = nullptr, *d_x_old = nullptr, *d_x_older = nullptr, *d_xdot = nullptr, *d_xdot_old = nullptr;
  double *d_a = nullptr, *d_resid = nullptr;
  int *d_flags = nullptr;
  long long *d_prof = nullptr;     // phase cycle counters (GOMA_GPU_PROFILE=1)
  int *d_elem_list = nullptr;      // colour-ordered element list
  // peer-memory exchange_dof
  unsigned long long *d_xflags = nullptr;  // [3][GOMA_GPU_MAX_NEIGHBORS] epochs published by the neighbours
  int num_neighbors = 0;
  void *peer_vec[3][GOMA_GPU_MAX_NEIGHBORS] = {};
  unsigned long long *peer_flags[GOMA_GPU_MAX_NEIGHBORS] = {};
  int my_slot_at[GOMA_GPU_MAX_NEIGHBORS] = {};
  int *d_recv_list = nullptr;
  std::vector<int> recv_ptr;
  int tail_begin = 0;
  unsigned long long epoch[3] = {0, 0, 0};
  double *d_sums = nullptr;  // goma_gpu_global_h_U
  unsigned char *d_elem_owned = nullptr;
  long long *d_csr_rowptr = nullptr;  // CSR hand-off
  int *d_csr_colind = nullptr, *d_csr_dpos = nullptr;
  double *d_csr_values = nullptr;
  long long csr_nnz = 0;
  double *d_scale = nullptr;      // row-sum scale vector
  double *d_partials = nullptr;   // per-block partial norms
  int *d_zero_rows = nullptr;
  int num_owned_unknowns = 0;
  std::vector<int> colour_begin;   // [ncolours+1]
  int num_sms = 0, blocks_per_sm = 0;  // cached launch geometry (cudaGetDeviceProperties is slow)
  int scatter_mode = 2;            // 0 fp64 atomics, 1 coloured load+add+store, 2 coloured first-touch stores
  int grid_limit = 0;
  double last_ms = 0.0;
  int last_launches = 0;
  size_t device_bytes = 0;
};

template <class T>
static int upload(T **dst, const T *src, size_t n, goma_gpu_ctx *c) {
  CU(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  c->device_bytes += n * sizeof(T);
  if (n) CU(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
template <class T>
static int dalloc(T **dst, size_t n, goma_gpu_ctx *c) {
  CU(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  c->device_bytes += n * sizeof(T);
  CU(cudaMemset(*dst, 0, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}

extern "C" const char *goma_gpu_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ kernel dispatch
namespace {

struct KernelEntry {
  void (*fn)(const FillParams);
  void (*build_records)(const FillParams, int);
  int tpe;
  size_t smem;
  int tbl_pad;
  size_t rec_bytes;
};

template <class C>
KernelEntry entry() {
  if constexpr (C::WS)
    return {fill_kernel_ws<C>, build_records_kernel<C>, C::TPE + C::NMUL, sizeof(Smem<C>), C::TBL_PAD, sizeof(ElemRec<C>)};
  else
    return {fill_kernel<C>, build_records_kernel<C>, C::TPE, sizeof(Smem<C>), C::TBL_PAD, sizeof(ElemRec<C>)};
}

// the instantiated physics/element combinations
//                      DIM NN NGP  P1    ENERGY NSPEC ALE  TPE TI MINB
int pick_kernel(const goma_gpu_problem &p, KernelEntry &k) {
  const bool p1 = p.pressure_interp == GOMA_PRESSURE_P1;
  if (p.ale) {  // pseudo-solid ARBITRARY mesh motion (config C4): Q2/P1, NS (+ energy) (+ species)
    if (!p1 || p.pspg) return fail(-2, "ALE is instantiated for Q2/P1 (QUAD9/HEX27) without PSPG only");
    const int fs = (p.energy ? 10 : 0) + p.num_species;  // field set: NS | NS+T | NS+Y | NS+T+2Y
    if (fs != 0 && fs != 10 && fs != 1 && fs != 12)
      return fail(-2, "ALE: instantiated field sets are NS, NS+T, NS+1 species, NS+T+2 species");
    if (p.elem_type == GOMA_GPU_QUAD9) {
      k = fs == 0    ? entry<Cfg<2, 9, 9, true, false, 0, true, 96, 1, 4>>()
          : fs == 10 ? entry<Cfg<2, 9, 9, true, true, 0, true, 96, 1, 4>>()
          : fs == 1  ? entry<Cfg<2, 9, 9, true, false, 1, true, 96, 1, 4>>()
                     : entry<Cfg<2, 9, 9, true, true, 2, true, 96, 1, 2>>();
      return 0;
    }
    if (p.elem_type == GOMA_GPU_HEX27) {
      k = fs == 0    ? entry<Cfg<3, 27, 27, true, false, 0, true, 256, 1, 1>>()
          : fs == 10 ? entry<Cfg<3, 27, 27, true, true, 0, true, 256, 1, 1>>()
          : fs == 1  ? entry<Cfg<3, 27, 27, true, false, 1, true, 256, 1, 1>>()
                     : entry<Cfg<3, 27, 27, true, true, 2, true, 256, 1, 1>>();
      return 0;
    }
    return fail(-2, "ALE needs QUAD9 or HEX27 elements");
  }
  if (p.pspg && p1) return fail(-2, "PSPG with P1 pressure is not supported by the GPU fill");
  if (!p1 && !p.pspg) return fail(-2, "equal-order velocity/pressure needs Pressure Stabilization (PSPG)");
  if (p1 && p.num_species) {  // Q2/P1 with species (Fickian): generic block path, one node pair per thread
    const int fs = (p.energy ? 10 : 0) + p.num_species;
    if (fs != 1 && fs != 12) return fail(-2, "Q2/P1 with species: instantiated field sets are NS+1 species, NS+T+2 species");
    if (p.elem_type == GOMA_GPU_QUAD9) {
      k = fs == 1 ? entry<Cfg<2, 9, 9, true, false, 1, false, 96, 1, 4>>() : entry<Cfg<2, 9, 9, true, true, 2, false, 96, 1, 2>>();
      return 0;
    }
    if (p.elem_type == GOMA_GPU_HEX27) {
      k = fs == 1 ? entry<Cfg<3, 27, 27, true, false, 1, false, 256, 1, 1>>()
                  : entry<Cfg<3, 27, 27, true, true, 2, false, 256, 1, 1>>();
      return 0;
    }
  }
  if (p1 && p.elem_type == GOMA_GPU_QUAD9) {
    k = p.energy ? entry<Cfg<2, 9, 9, true, true, 0, false, 32, 3, 8>>()
                 : entry<Cfg<2, 9, 9, true, false, 0, false, 32, 3, 8>>();
    return 0;
  }
  if (p1 && p.elem_type == GOMA_GPU_HEX27) {
    // config C2 (Q2/P1 Navier-Stokes): warp-specialised kernel, 192 builder + 128 multiplier threads, one CTA per SM
    static const bool ws = getenv("GOMA_GPU_WS") ? atoi(getenv("GOMA_GPU_WS")) != 0 : false;
    k = p.energy ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2>>()
        : ws     ? entry<Cfg<3, 27, 27, true, false, 0, false, 192, 3, 1, true>>()
                 : entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2>>();
    return 0;
  }
  if (!p1 && p.elem_type == GOMA_GPU_HEX8) {  // Q1/Q1 PSPG (config C5 and its sub-cases)
    if (p.energy && p.num_species == 2) { k = entry<Cfg<3, 8, 8, false, true, 2, false, 64, 1, 8>>(); return 0; }
    if (p.energy && p.num_species == 0) { k = entry<Cfg<3, 8, 8, false, true, 0, false, 64, 1, 4>>(); return 0; }
    if (!p.energy && p.num_species == 0) { k = entry<Cfg<3, 8, 8, false, false, 0, false, 64, 1, 4>>(); return 0; }
    return fail(-2, "hex8 Q1/Q1: instantiated field sets are NS, NS+T, NS+T+2 species");
  }
  if (!p1 && p.elem_type == GOMA_GPU_QUAD4) {
    if (p.energy && p.num_species == 2) { k = entry<Cfg<2, 4, 4, false, true, 2, false, 32, 1, 8>>(); return 0; }
    if (!p.energy && p.num_species == 0) { k = entry<Cfg<2, 4, 4, false, false, 0, false, 32, 1, 8>>(); return 0; }
    return fail(-2, "quad4 Q1/Q1: instantiated field sets are NS, NS+T+2 species");
  }
  return fail(-2, "element type / interpolation combination not supported by the GPU fill");
}

}  // namespace

// ------------------------------------------------------------------ init / destroy
static void static_params(const goma_gpu_ctx *c, FillParams &P);
static int validate(const goma_gpu_problem &p) {
  if (p.dim != 2 && p.dim != 3) return fail(-2, "dim must be 2 or 3");
  const int et = p.elem_type;
  if (et != GOMA_GPU_QUAD4 && et != GOMA_GPU_QUAD9 && et != GOMA_GPU_HEX8 && et != GOMA_GPU_HEX27)
    return fail(-2, "element types other than QUAD4/QUAD9/HEX8/HEX27 are not supported");
  if ((p.dim == 2) != (et == GOMA_GPU_QUAD4 || et == GOMA_GPU_QUAD9)) return fail(-2, "dim / element type mismatch");
  if (p.num_kinds < 1 || p.num_kinds > GOMA_GPU_MAX_KINDS) return fail(-2, "num_kinds out of range");
  if (p.num_species < 0 || p.num_species > 4) return fail(-2, "num_species out of range (MAX_CONC = 4)");
  if (!p.elem_connect || !p.first_unknown || !p.node_kind || !p.dbc_flag || !p.dbc_value)
    return fail(-2, "null array in goma_gpu_problem");
  for (int d = 0; d < p.dim; d++)
    if (!p.coord[d]) return fail(-2, "null coordinate array");
  if (p.pressure_interp == GOMA_PRESSURE_P1 && !(et == GOMA_GPU_QUAD9 || et == GOMA_GPU_HEX27))
    return fail(-2, "P1 pressure needs a centroid node (QUAD9/HEX27)");
  if (p.num_owned_nodes < 0 || p.num_owned_nodes > p.num_nodes) return fail(-2, "num_owned_nodes out of range");
  return 0;
}

extern "C" int goma_gpu_fill_init(const goma_gpu_problem *problem, int device, goma_gpu_ctx **out) {
  if (!problem || !out) return fail(-2, "null argument");
  *out = nullptr;
  const goma_gpu_problem &p = *problem;
  if (int rc = validate(p)) return rc;
  KernelEntry ke;
  if (int rc = pick_kernel(p, ke)) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(-3, "no CUDA device: the goma_gpu_fill path has no CPU fallback");
  CU(cudaSetDevice(device));

  goma_gpu_ctx *c = new goma_gpu_ctx();
  c->prob = p;
  c->device = device;
  c->num_owned_unknowns = p.num_owned_nodes < p.num_nodes ? p.first_unknown[p.num_owned_nodes] : p.num_unknowns;
  int nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::string perr = build_pattern(p, c->pat, std::min(nthreads, 32));
  if (!perr.empty()) {
    delete c;
    return fail(-2, "sparsity pattern: " + perr);
  }
  // optional bit-exact check against the host's own MSR graph
  if (p.ija) {
    const int N = p.num_unknowns;
    if (c->pat.nnz_plus > 2147483647LL) {
      delete c;
      return fail(-2, "host ija given but nnz exceeds the 32-bit MSR limit");
    }
    std::vector<int> mine((size_t)c->pat.nnz_plus + 1, 0);
    emit_msr_columns(p, c->pat, mine.data());
    (void)N;
    for (long long k = 0; k < c->pat.nnz_plus; k++) {
      if (mine[k] != p.ija[k]) {
        delete c;
        return fail(-2, "host MSR graph differs from the derived one at ija[" + std::to_string(k) + "]");
      }
    }
  }

  const int nn = p.num_nodes, ne = p.num_elems, npe = p.elem_type, N = p.num_unknowns;
  int rc = 0;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  rc |= upload(&c->d_conn, p.elem_connect, (size_t)ne * npe, c);
  rc |= upload(&c->d_first, p.first_unknown, nn, c);
  for (int d = 0; d < p.dim; d++) rc |= upload(&c->d_coord[d], p.coord[d], nn, c);
  rc |= upload(&c->d_kind, p.node_kind, nn, c);
  rc |= upload(&c->d_dbc_flag, p.dbc_flag, N, c);
  rc |= upload(&c->d_dbc_value, p.dbc_value, N, c);
  static_assert(sizeof(long long) == sizeof(int64_t), "64-bit row pointers");
  rc |= upload(&c->d_rowstart, (const long long *)c->pat.rowstart.data(), (size_t)N + 1, c);
  rc |= upload(&c->d_pair_full, (const unsigned short *)c->pat.pair_full.data(), c->pat.pair_full.size(), c);
  rc |= upload(&c->d_pair_p, (const unsigned short *)c->pat.pair_p.data(), c->pat.pair_p.size(), c);
  rc |= upload(&c->d_pair_first, (const unsigned *)c->pat.pair_first.data(), c->pat.pair_first.size(), c);
  rc |= upload(&c->d_node_first, (const unsigned *)c->pat.node_first.data(), c->pat.node_first.size(), c);
  rc |= upload(&c->d_elem_list, c->pat.colour_order.data(), c->pat.colour_order.size(), c);
  c->colour_begin = c->pat.colour_begin;
  if (rc) {
    goma_gpu_fill_destroy(c);
    return -3;
  }
  // pair tables live on the device from here on
  std::vector<uint16_t>().swap(c->pat.pair_full);
  std::vector<uint16_t>().swap(c->pat.pair_p);
  std::vector<uint32_t>().swap(c->pat.pair_first);
  std::vector<uint32_t>().swap(c->pat.node_first);
  std::vector<int>().swap(c->pat.colour_order);

  // quadrature / basis tables, packed in the order Smem<C>::tbl expects
  ElemTables t = make_tables(p.elem_type);
  std::vector<double> packed;
  packed.insert(packed.end(), t.wt.begin(), t.wt.end());
  packed.insert(packed.end(), t.phi.begin(), t.phi.end());
  packed.insert(packed.end(), t.dphi.begin(), t.dphi.end());
  packed.insert(packed.end(), t.psi.begin(), t.psi.end());
  packed.resize(ke.tbl_pad, 0.0);
  rc |= upload(&c->d_tables, packed.data(), packed.size(), c);

  rc |= dalloc(&c->d_x, N, c);
  rc |= dalloc(&c->d_x_old, N, c);
  rc |= dalloc(&c->d_x_older, N, c);
  rc |= dalloc(&c->d_xdot, N, c);
  rc |= dalloc(&c->d_xdot_old, N, c);
  rc |= dalloc(&c->d_resid, N, c);
  rc |= dalloc(&c->d_a, (size_t)c->pat.nnz_plus + 1, c);
  rc |= dalloc(&c->d_flags, 4, c);
  if (rc) {
    goma_gpu_fill_destroy(c);
    return -3;
  }

  // per-element gather records: everything load_elem_dofptr would recompute per element and per
  // iteration, gathered once; afterwards the pair tables they were built from are dropped
  {
    CU(cudaMalloc((void **)&c->d_erec, std::max<size_t>((size_t)ne * ke.rec_bytes, 16)));
    c->device_bytes += (size_t)ne * ke.rec_bytes;
    FillParams P;
    static_params(c, P);
    if (ne > 0) ke.build_records<<<(ne + 127) / 128, 128, 0, c->stream>>>(P, ne);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    void *drop[] = {c->d_pair_full, c->d_pair_p, c->d_pair_first, c->d_node_first};
    for (void *q : drop)
      if (q) cudaFree(q);
    c->d_pair_full = c->d_pair_p = nullptr;
    c->d_pair_first = c->d_node_first = nullptr;
  }

  if (getenv("GOMA_GPU_PROFILE")) rc |= dalloc(&c->d_prof, 2 * 8 * 4096, c);
  if (ke.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute((const void *)ke.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ke.smem);
    if (e != cudaSuccess) {
      goma_gpu_fill_destroy(c);
      return fail(-3, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    }
  }
  *out = c;
  return 0;
}

extern "C" void goma_gpu_fill_destroy(goma_gpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  void *ptrs[] = {c->d_conn, c->d_first, c->d_coord[0], c->d_coord[1], c->d_coord[2], c->d_kind, c->d_dbc_flag,
                  c->d_dbc_value, c->d_rowstart, c->d_pair_full, c->d_pair_p, c->d_pair_first, c->d_node_first, c->d_prof, c->d_tables, c->d_erec, c->d_x, c->d_x_old,
                  c->d_x_older, c->d_xdot, c->d_xdot_old, c->d_a, c->d_resid, c->d_flags, c->d_elem_list};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  for (int k = 0; k < c->num_neighbors; k++) {
    for (int v = 0; v < 3; v++)
      if (c->peer_vec[v][k]) cudaIpcCloseMemHandle(c->peer_vec[v][k]);
    if (c->peer_flags[k]) cudaIpcCloseMemHandle(c->peer_flags[k]);
  }
  if (c->d_xflags) cudaFree(c->d_xflags);
  if (c->d_recv_list) cudaFree(c->d_recv_list);
  if (c->d_sums) cudaFree(c->d_sums);
  if (c->d_elem_owned) cudaFree(c->d_elem_owned);
  if (c->d_csr_rowptr) cudaFree(c->d_csr_rowptr);
  if (c->d_csr_colind) cudaFree(c->d_csr_colind);
  if (c->d_csr_dpos) cudaFree(c->d_csr_dpos);
  if (c->d_csr_values) cudaFree(c->d_csr_values);
  if (c->d_scale) cudaFree(c->d_scale);
  if (c->d_partials) cudaFree(c->d_partials);
  if (c->d_zero_rows) cudaFree(c->d_zero_rows);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int goma_gpu_fill_get_msr(goma_gpu_ctx *c, long long *nnz_plus) {
  if (!c) return fail(-2, "null context");
  if (nnz_plus) *nnz_plus = c->pat.nnz_plus;
  return 0;
}

// ija export needs the host arrays again (they are not retained in the context)
extern "C" int goma_gpu_fill_export_msr(goma_gpu_ctx *c, const goma_gpu_problem *p, int *ija_out) {
  if (!c || !p || !ija_out) return fail(-2, "null argument");
  if (c->pat.nnz_plus > 2147483647LL) return fail(-2, "nnz exceeds the 32-bit MSR limit of ija");
  emit_msr_columns(*p, c->pat, ija_out);
  return 0;
}

// Host-only: derive the MSR graph (no device needed).  ija_out may be NULL to query nnz_plus.
extern "C" int goma_gpu_pattern_msr(const goma_gpu_problem *p, long long *nnz_plus, int *ija_out) {
  if (!p) return fail(-2, "null argument");
  if (int rc = validate(*p)) return rc;
  Pattern pat;
  int nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::string perr = build_pattern(*p, pat, std::min(nthreads, 32));
  if (!perr.empty()) return fail(-2, "sparsity pattern: " + perr);
  if (nnz_plus) *nnz_plus = pat.nnz_plus;
  if (ija_out) {
    if (pat.nnz_plus > 2147483647LL) return fail(-2, "nnz exceeds the 32-bit MSR limit of ija");
    emit_msr_columns(*p, pat, ija_out);
  }
  return 0;
}

extern "C" int goma_gpu_fill_set_option(goma_gpu_ctx *c, const char *name, int value) {
  if (!c || !name) return fail(-2, "null argument");
  if (!strcmp(name, "scatter")) {
    if (value < 0 || value > 2) return fail(-2, "scatter must be 0 (atomic), 1 (coloured) or 2 (first-touch)");
    c->scatter_mode = value;
    return 0;
  }
  if (!strcmp(name, "grid_limit")) {
    c->grid_limit = value;
    return 0;
  }
  return fail(-2, std::string("unknown option ") + name);
}

extern "C" int goma_gpu_fill_device_buffers(goma_gpu_ctx *c, goma_gpu_device_buffers *o) {
  if (!c || !o) return fail(-2, "null argument");
  o->d_x = c->d_x;
  o->d_x_old = c->d_x_old;
  o->d_x_older = c->d_x_older;
  o->d_xdot = c->d_xdot;
  o->d_xdot_old = c->d_xdot_old;
  o->d_a = c->d_a;
  o->d_resid = c->d_resid;
  o->stream = (void *)c->stream;
  return 0;
}

// mesh / map part of the kernel arguments (also what build_records_kernel reads)
static void static_params(const goma_gpu_ctx *c, FillParams &P) {
  const goma_gpu_problem &p = c->prob;
  memset(&P, 0, sizeof(P));
  P.conn = c->d_conn;
  for (int d = 0; d < 3; d++) P.coord[d] = c->d_coord[d];
  P.first_unknown = c->d_first;
  P.node_kind = c->d_kind;
  memcpy(P.kind_slot, p.kind_slot, sizeof(P.kind_slot));
  P.rowstart = c->d_rowstart;
  P.pair_full = c->d_pair_full;
  P.pair_p = c->d_pair_p;
  P.pair_first = c->d_pair_first;
  P.node_first = c->d_node_first;
  P.dbc_flag = c->d_dbc_flag;
  P.dbc_value = c->d_dbc_value;
  P.num_owned_nodes = p.num_owned_nodes;
  P.erec = c->d_erec;
}

// ------------------------------------------------------------------ the fill
static int launch_fill(goma_gpu_ctx *c, double delta_t, double theta, double time_value, double h_elem_avg,
                       double U_norm, int assemble_residual, int assemble_jacobian) {
  const goma_gpu_problem &p = c->prob;
  KernelEntry ke;
  if (int rc = pick_kernel(p, ke)) return rc;
  FillParams P;
  static_params(c, P);
  P.x = c->d_x;
  P.x_old = c->d_x_old;
  P.xdot = c->d_xdot;
  P.a = c->d_a;
  P.resid = c->d_resid;
  P.flags = c->d_flags;
  P.tables = c->d_tables;
  P.assemble_residual = assemble_residual;
  P.assemble_jacobian = assemble_jacobian;
  P.transient = p.transient;
  memcpy(P.etm_mom, p.etm_momentum, sizeof(P.etm_mom));
  memcpy(P.etm_cont, p.etm_continuity, sizeof(P.etm_cont));
  memcpy(P.etm_energy, p.etm_energy, sizeof(P.etm_energy));
  memcpy(P.etm_species, p.etm_species, sizeof(P.etm_species));
  memcpy(P.etm_mesh, p.etm_mesh, sizeof(P.etm_mesh));
  if (!p.transient) {  // steady: the *_dot terms are skipped
    P.etm_mom[0] = 0.0;
    P.etm_energy[0] = 0.0;
    P.etm_species[0] = 0.0;
  }
  P.rho = p.rho;
  P.mu = p.mu;
  P.k = p.conductivity;
  P.Cp = p.heat_capacity;
  P.beta = p.volume_expansion;
  P.Tref = p.reference_temperature;
  P.heat_source = p.heat_source;
  for (int d = 0; d < 3; d++) P.g[d] = p.momentum_source[d];
  P.source_model = p.momentum_source_model;
  for (int w = 0; w < 4; w++) P.diffusivity[w] = p.diffusivity[w];
  P.delta_t = delta_t;
  P.theta = theta;
  P.time_value = time_value;
  P.h_elem_avg = h_elem_avg;
  P.U_norm = U_norm;
  P.lame_mu = p.lame_mu;
  P.lame_lambda = p.lame_lambda;
  P.pspg = p.pspg;
  P.ps_scaling = p.ps_scaling;
  P.prof = c->d_prof;
  P.debug = getenv("GOMA_GPU_DEBUG") ? atoi(getenv("GOMA_GPU_DEBUG")) : 0;
  if (p.transient && !(delta_t > 0.0)) return fail(-2, "transient fill needs delta_t > 0");

  if (c->num_sms == 0) {
    CU(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->blocks_per_sm, (const void *)ke.fn, ke.tpe, ke.smem));
    if (c->blocks_per_sm < 1) return fail(-3, "fill kernel does not fit on an SM");
  }
  int max_grid = c->num_sms * c->blocks_per_sm;
  if (c->grid_limit > 0) max_grid = std::min(max_grid, c->grid_limit);

  c->last_launches = 0;
  CU(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->stream));
  CU(cudaEventRecord(c->ev0, c->stream));
  P.scatter_mode = c->scatter_mode;
  if (c->scatter_mode != 2) {
    // accumulate-into semantics need zeroed storage; the first-touch mode overwrites every slot the
    // elements touch and never writes the others (zeroed once at init), so it needs no memset
    if (assemble_residual) CU(cudaMemsetAsync(c->d_resid, 0, (size_t)p.num_unknowns * sizeof(double), c->stream));
    if (assemble_jacobian) CU(cudaMemsetAsync(c->d_a, 0, ((size_t)c->pat.nnz_plus + 1) * sizeof(double), c->stream));
  }
  if (c->scatter_mode == 0) {
    P.elem_list = nullptr;
    P.elem_begin = 0;
    P.elem_end = p.num_elems;
    int grid = std::max(1, std::min(max_grid, p.num_elems));
    ke.fn<<<grid, ke.tpe, ke.smem, c->stream>>>(P);
    c->last_launches++;
  } else {
    // colour classes in increasing order, one launch each: stream order is the inter-colour barrier
    P.elem_list = c->d_elem_list;
    for (size_t col = 0; col + 1 < c->colour_begin.size(); col++) {
      P.elem_begin = c->colour_begin[col];
      P.elem_end = c->colour_begin[col + 1];
      int n = P.elem_end - P.elem_begin;
      if (n <= 0) continue;
      int grid = std::max(1, std::min(max_grid, n));
      ke.fn<<<grid, ke.tpe, ke.smem, c->stream>>>(P);
      c->last_launches++;
    }
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev1, c->stream));
  return 0;
}

static int finish_fill(goma_gpu_ctx *c, int flags_out[3]) {
  int h_flags[4] = {0, 0, 0, 0};
  CU(cudaMemcpyAsync(h_flags, c->d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (c->d_prof) {
    std::vector<long long> h(2 * 8 * 4096);
    CU(cudaMemcpy(h.data(), c->d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double s[8] = {0}, s2[8] = {0};
    int nb = 0;
    for (int b = 0; b < 4096; b++)
      if (h[b * 8 + 6] > 0) {
        nb++;
        for (int k = 0; k < 6; k++) s[k] += (double)h[b * 8 + k] / (double)h[b * 8 + 6];
        for (int k = 0; k < 6; k++) s2[k] += (double)h[(4096 + b) * 8 + k] / (double)h[b * 8 + 6];
      }
    if (nb)
      fprintf(stderr, "[goma_gpu profile] build split: ale-x %.0f J %.0f inv %.0f grads %.0f fields %.0f gp+vg %.0f\n",
              s2[0] / nb, s2[1] / nb, s2[2] / nb, s2[3] / nb, s2[4] / nb, s2[5] / nb);
    if (nb)
      fprintf(stderr, "[goma_gpu profile] cycles/element/CTA (mean over %d CTAs of the last launch): build %.0f rows %.0f "
                      "gauss loop + write-out %.0f\n", nb, s[0] / nb, s[1] / nb, s[2] / nb);
  }
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  if (flags_out) {
    flags_out[0] = h_flags[0];  // neg_elem_volume
    flags_out[1] = h_flags[1];  // neg_lub_height
    flags_out[2] = h_flags[2];  // zero_detJ
  }
  // matrix_fill_full's return convention (mm_fill.c:285-311)
  return (h_flags[0] || h_flags[1] || h_flags[2]) ? -1 : 0;
}

extern "C" int goma_gpu_fill_device(goma_gpu_ctx *c, double delta_t, double theta, double time_value,
                                    double h_elem_avg, double U_norm, int assemble_residual, int assemble_jacobian,
                                    int flags_out[3]) {
  if (!c) return fail(-2, "null context");
  CU(cudaSetDevice(c->device));
  if (int rc = launch_fill(c, delta_t, theta, time_value, h_elem_avg, U_norm, assemble_residual, assemble_jacobian))
    return rc;
  return finish_fill(c, flags_out);
}

extern "C" int goma_gpu_fill(goma_gpu_ctx *c, const double *x, const double *x_old, const double *x_older,
                             const double *xdot, const double *xdot_old, double delta_t, double theta,
                             double time_value, double h_elem_avg, double U_norm, int assemble_residual,
                             int assemble_jacobian, double *a, double *resid_vector, int flags_out[3]) {
  if (!c) return fail(-2, "null context");
  if (!x) return fail(-2, "x is null");
  if (assemble_jacobian && !a) return fail(-2, "a is null");
  if (assemble_residual && !resid_vector) return fail(-2, "resid_vector is null");
  CU(cudaSetDevice(c->device));
  const size_t nb = (size_t)c->prob.num_unknowns * sizeof(double);
  CU(cudaMemcpyAsync(c->d_x, x, nb, cudaMemcpyHostToDevice, c->stream));
  if (c->prob.transient) {
    if (!xdot) return fail(-2, "transient fill needs xdot");
    CU(cudaMemcpyAsync(c->d_xdot, xdot, nb, cudaMemcpyHostToDevice, c->stream));
    if (x_old) CU(cudaMemcpyAsync(c->d_x_old, x_old, nb, cudaMemcpyHostToDevice, c->stream));
    if (x_older) CU(cudaMemcpyAsync(c->d_x_older, x_older, nb, cudaMemcpyHostToDevice, c->stream));
    if (xdot_old) CU(cudaMemcpyAsync(c->d_xdot_old, xdot_old, nb, cudaMemcpyHostToDevice, c->stream));
  }
  if (int rc = launch_fill(c, delta_t, theta, time_value, h_elem_avg, U_norm, assemble_residual, assemble_jacobian))
    return rc;
  // one copy: the PCIe link is saturated by it (≈47 GB/s measured; two concurrent copy streams gave the same)
  if (assemble_jacobian)
    CU(cudaMemcpyAsync(a, c->d_a, ((size_t)c->pat.nnz_plus + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (assemble_residual) CU(cudaMemcpyAsync(resid_vector, c->d_resid, nb, cudaMemcpyDeviceToHost, c->stream));
  return finish_fill(c, flags_out);
}

extern "C" int goma_gpu_fill_last_stats(goma_gpu_ctx *c, double *kernel_ms, int *launches) {
  if (!c) return fail(-2, "null context");
  if (kernel_ms) *kernel_ms = c->last_ms;
  if (launches) *launches = c->last_launches;
  return 0;
}

// ------------------------------------------------------------------ exchange_dof halves
__global__ void pack_dofs_kernel(const double *__restrict__ v, const int *__restrict__ list, int n,
                                 double *__restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[list[i]];
}
__global__ void unpack_dofs_kernel(double *__restrict__ v, const int *__restrict__ list, int n,
                                   const double *__restrict__ buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[list[i]] = buf[i];
}
extern "C" int goma_gpu_pack_dofs(goma_gpu_ctx *c, const double *d_vec, const int *d_list, int n, double *d_buf) {
  if (!c) return fail(-2, "null context");
  if (n <= 0) return 0;
  pack_dofs_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_vec, d_list, n, d_buf);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int goma_gpu_unpack_dofs(goma_gpu_ctx *c, double *d_vec, const int *d_list, int n, const double *d_buf) {
  if (!c) return fail(-2, "null context");
  if (n <= 0) return 0;
  unpack_dofs_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_vec, d_list, n, d_buf);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

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
  double vals[4] = {h, cnt, vv, nv};
  for (int q = 0; q < 4; q++) {
    double v = vals[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&sums[q], v);
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
  if (!c->d_sums) CU(cudaMalloc((void **)&c->d_sums, 4 * sizeof(double)));
  CU(cudaMemsetAsync(c->d_sums, 0, 4 * sizeof(double), c->stream));
  int ku[4] = {-1, -1, -1, -1};  // offset of U inside a node of each kind (V, W follow it)
  for (int k = 0; k < p.num_kinds && k < 4; k++) ku[k] = p.kind_slot[k][GOMA_SLOT_U];
  const int threads = 256, blocks = std::max(1, std::min(148 * 8, (std::max(p.num_elems, p.num_owned_nodes) + threads - 1) / threads));
  global_h_U_kernel<<<blocks, threads, 0, c->stream>>>(c->d_conn, p.elem_type, p.dim, p.num_elems, c->d_coord[0],
                                                       c->d_coord[1], c->d_coord[2], d_owned, c->d_first, c->d_kind, 0, 1, 2,
                                                       ku[0], ku[1], ku[2], ku[3], p.num_owned_nodes, c->d_x, c->d_sums);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(sums_out, c->d_sums, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------ exchange_dof over peer memory
struct ExchangeArgs {
  int nn;
  unsigned long long epoch;
  unsigned long long *peer_ready[GOMA_GPU_MAX_NEIGHBORS];  // the slot of this rank in each neighbour's flag block
  const unsigned long long *my_ready;                      // this rank's flag block, row of the vector
  const double *peer_vec[GOMA_GPU_MAX_NEIGHBORS];
  int recv_ptr[GOMA_GPU_MAX_NEIGHBORS + 1];
  const int *recv_list;
  double *tail;
};

__global__ void exchange_dof_kernel(const __grid_constant__ ExchangeArgs A) {
  // publish: everything written to this rank's vector before this kernel (stream order) is visible to the
  // neighbours once they observe the epoch
  if (blockIdx.x == 0 && threadIdx.x < A.nn) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.peer_ready[threadIdx.x]), "l"(A.epoch) : "memory");
  }
  const int total = A.recv_ptr[A.nn];
  for (int base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
    const int k = base + threadIdx.x;
    int nb = 0;
    if (k < total) {
      while (k >= A.recv_ptr[nb + 1]) nb++;
      unsigned long long seen;
      do {  // the neighbour's vector of this epoch is complete
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(A.my_ready + nb) : "memory");
      } while (seen < A.epoch);
      A.tail[k] = A.peer_vec[nb][A.recv_list[k]];
    }
  }
}

extern "C" int goma_gpu_exchange_export(goma_gpu_ctx *c, goma_gpu_exchange_handles *out) {
  if (!c || !out) return fail(-2, "null argument");
  CU(cudaSetDevice(c->device));
  if (!c->d_xflags) {
    CU(cudaMalloc((void **)&c->d_xflags, 3 * GOMA_GPU_MAX_NEIGHBORS * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_xflags, 0, 3 * GOMA_GPU_MAX_NEIGHBORS * sizeof(unsigned long long)));
    CU(cudaDeviceSynchronize());
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == GOMA_GPU_IPC_HANDLE_BYTES, "IPC handle size");
  memset(out, 0, sizeof(*out));
  double *vecs[3] = {c->d_x, c->d_xdot, c->d_x_old};
  for (int v = 0; v < 3; v++) CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->vec[v], vecs[v]));
  CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out->flags, c->d_xflags));
  out->device = c->device;
  return 0;
}

extern "C" int goma_gpu_exchange_setup(goma_gpu_ctx *c, int num_neighbors, const goma_gpu_exchange_handles *nh,
                                       const int *my_slot_at_neighbor, const int *recv_ptr, const int *recv_list,
                                       int tail_begin) {
  if (!c) return fail(-2, "null context");
  if (num_neighbors < 0 || num_neighbors > GOMA_GPU_MAX_NEIGHBORS) return fail(-2, "too many neighbours");
  if (num_neighbors && (!nh || !my_slot_at_neighbor || !recv_ptr || !recv_list)) return fail(-2, "null argument");
  if (!c->d_xflags) return fail(-2, "call goma_gpu_exchange_export first");
  CU(cudaSetDevice(c->device));
  for (int k = 0; k < c->num_neighbors; k++) {  // a second set-up replaces the first
    for (int v = 0; v < 3; v++)
      if (c->peer_vec[v][k]) cudaIpcCloseMemHandle(c->peer_vec[v][k]), c->peer_vec[v][k] = nullptr;
    if (c->peer_flags[k]) cudaIpcCloseMemHandle(c->peer_flags[k]), c->peer_flags[k] = nullptr;
  }
  c->num_neighbors = num_neighbors;
  c->tail_begin = tail_begin;
  c->recv_ptr.assign(recv_ptr, recv_ptr + num_neighbors + 1);
  if (tail_begin + c->recv_ptr[num_neighbors] > c->prob.num_unknowns) return fail(-2, "external tail exceeds the vector");
  for (int k = 0; k < num_neighbors; k++) {
    if (my_slot_at_neighbor[k] < 0 || my_slot_at_neighbor[k] >= GOMA_GPU_MAX_NEIGHBORS) return fail(-2, "bad neighbour slot");
    c->my_slot_at[k] = my_slot_at_neighbor[k];
    for (int v = 0; v < 3; v++)
      CU(cudaIpcOpenMemHandle(&c->peer_vec[v][k], *(const cudaIpcMemHandle_t *)nh[k].vec[v], cudaIpcMemLazyEnablePeerAccess));
    void *pf = nullptr;
    CU(cudaIpcOpenMemHandle(&pf, *(const cudaIpcMemHandle_t *)nh[k].flags, cudaIpcMemLazyEnablePeerAccess));
    c->peer_flags[k] = (unsigned long long *)pf;
  }
  if (c->d_recv_list) cudaFree(c->d_recv_list);
  c->d_recv_list = nullptr;
  const int total = c->recv_ptr[num_neighbors];
  CU(cudaMalloc((void **)&c->d_recv_list, std::max(1, total) * sizeof(int)));
  if (total) CU(cudaMemcpy(c->d_recv_list, recv_list, (size_t)total * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int goma_gpu_exchange_dof(goma_gpu_ctx *c, int which) {
  if (!c) return fail(-2, "null context");
  if (which < 0 || which > 2) return fail(-2, "which must be 0 (x), 1 (xdot) or 2 (x_old)");
  if (c->num_neighbors == 0) return 0;
  CU(cudaSetDevice(c->device));
  ExchangeArgs A;
  memset(&A, 0, sizeof(A));
  A.nn = c->num_neighbors;
  A.epoch = ++c->epoch[which];
  double *vecs[3] = {c->d_x, c->d_xdot, c->d_x_old};
  for (int k = 0; k < A.nn; k++) {
    A.peer_ready[k] = c->peer_flags[k] + which * GOMA_GPU_MAX_NEIGHBORS + c->my_slot_at[k];
    A.peer_vec[k] = (const double *)c->peer_vec[which][k];
    A.recv_ptr[k] = c->recv_ptr[k];
  }
  A.recv_ptr[A.nn] = c->recv_ptr[A.nn];
  A.my_ready = c->d_xflags + which * GOMA_GPU_MAX_NEIGHBORS;
  A.recv_list = c->d_recv_list;
  A.tail = vecs[which] + c->tail_begin;
  const int total = c->recv_ptr[A.nn];
  const int threads = 256, blocks = std::max(1, std::min(148, (total + threads - 1) / threads));
  exchange_dof_kernel<<<blocks, threads, 0, c->stream>>>(A);
  CU(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ after the fill: row-sum scaling, norms
// One warp per owned row: the off-diagonal run of an MSR row is contiguous, lanes stride over it (coalesced),
// the second sweep over the row (the division) hits L1/L2.  HBM-bound: reads and writes every value once.
__global__ void __launch_bounds__(256, 6) row_sum_scale_kernel(int nrows, const long long *__restrict__ rowstart, double *__restrict__ a,
                                     double *__restrict__ b, double *__restrict__ scale, int *__restrict__ zero_rows) {
  const int lane = threadIdx.x & 31;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  long long k0n = 0, k1n = 0;
  if (row < nrows) {
    k0n = rowstart[row];
    k1n = rowstart[row + 1];
  }
  for (; row < nrows; row += nwarp) {
    const long long k0 = k0n, k1 = k1n;
    if (row + nwarp < nrows) {  // the next row's extent is on its way while this row streams
      k0n = rowstart[row + nwarp];
      k1n = rowstart[row + nwarp + 1];
    }
    double sum = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long long k = k0 + lane;
    for (; k + 96 < k1; k += 128) {  // four independent 256-byte requests in flight per warp
      const double v0 = a[k], v1 = a[k + 32], v2 = a[k + 64], v3 = a[k + 96];
      sum += fabs(v0);
      s1 += fabs(v1);
      s2 += fabs(v2);
      s3 += fabs(v3);
    }
    for (; k < k1; k += 32) sum += fabs(a[k]);
    sum = (sum + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double diag = a[row];
    double row_sum = fabs(diag) + sum;
    if (fabs(diag) > 1.0e-200) row_sum = diag >= 0.0 ? row_sum : -row_sum;  // keep the diagonal positive (:547-549)
    // the reference divides (a[k] /= row_sum); one reciprocal per row and a multiply per entry differ from that by
    // at most 1 ulp (the parity tolerance is 1e-12) and take the fp64 divide sequence (~20 instructions per entry)
    // off an HBM-bound pass
    const double inv = 1.0 / row_sum;
    k = k0 + lane;
    for (; k + 96 < k1; k += 128) {
      const double v0 = a[k], v1 = a[k + 32], v2 = a[k + 64], v3 = a[k + 96];
      a[k] = v0 * inv;
      a[k + 32] = v1 * inv;
      a[k + 64] = v2 * inv;
      a[k + 96] = v3 * inv;
    }
    for (; k < k1; k += 32) a[k] = a[k] * inv;
    if (lane == 0) {
      scale[row] = row_sum;
      if (row_sum == 0.0) atomicAdd(zero_rows, 1);
      a[row] = diag / row_sum;
      b[row] = b[row] / row_sum;
    }
  }
}

extern "C" int goma_gpu_row_sum_scale(goma_gpu_ctx *c, double *scale_out, int *zero_rows_out) {
  if (!c) return fail(-2, "null context");
  CU(cudaSetDevice(c->device));
  const int n = c->num_owned_unknowns;
  if (!c->d_scale) CU(cudaMalloc((void **)&c->d_scale, std::max(1, c->prob.num_unknowns) * sizeof(double)));
  if (!c->d_zero_rows) CU(cudaMalloc((void **)&c->d_zero_rows, sizeof(int)));
  CU(cudaMemsetAsync(c->d_zero_rows, 0, sizeof(int), c->stream));
  if (n > 0) {
    const int threads = 256;
    int per_sm = 0;  // a whole number of resident waves: the rows are handed out grid-stride
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)row_sum_scale_kernel, threads, 0));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int blocks = std::max(1, std::min(sms * std::max(per_sm, 1), (n + 7) / 8));
    row_sum_scale_kernel<<<blocks, threads, 0, c->stream>>>(n, c->d_rowstart, c->d_a, c->d_resid, c->d_scale, c->d_zero_rows);
    CU(cudaGetLastError());
  }
  int zr = 0;
  CU(cudaMemcpyAsync(&zr, c->d_zero_rows, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (scale_out && n > 0) CU(cudaMemcpyAsync(scale_out, c->d_scale, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (zero_rows_out) *zero_rows_out = zr;
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
struct CsrKinds {
  int num_unknowns[GOMA_GPU_MAX_KINDS];  // unknowns of a node of each kind
  int num_pressure[GOMA_GPU_MAX_KINDS];  // ... of which pressure (last in the node)
  int tslot[GOMA_GPU_MAX_KINDS];         // offset of T inside the node (-1: none): energy rows carry no P columns
};

// one thread per owned node: the rows of its unknowns share the node-node list (exo_conn.c build_node_node);
// columns = the unknowns of the neighbour nodes in increasing node id (find_MSR_problem_graph), diagonal included
__global__ void csr_structure_kernel(int num_owned_nodes, const long long *__restrict__ nn_ptr,
                                     const int *__restrict__ nn_list, const int *__restrict__ first_unknown,
                                     const unsigned char *__restrict__ node_kind, const __grid_constant__ CsrKinds K,
                                     const long long *__restrict__ rowstart, long long msr0,
                                     long long *__restrict__ rowptr, int *__restrict__ colind, int *__restrict__ dpos,
                                     int num_rows) {
  const int nd = blockIdx.x * blockDim.x + threadIdx.x;
  if (nd >= num_owned_nodes) return;
  const int kd = node_kind[nd], fu = first_unknown[nd];
  for (int s = 0; s < K.num_unknowns[kd]; s++) {
    const int row = fu + s;
    const long long base = rowstart[row] - msr0 + row;  // every earlier row adds its diagonal
    rowptr[row] = base;
    if (row == num_rows - 1) rowptr[num_rows] = rowstart[row + 1] - msr0 + row + 1;
    const bool nop = K.tslot[kd] >= 0 && s == K.tslot[kd];
    long long pos = base;
    for (long long q = nn_ptr[nd]; q < nn_ptr[nd + 1]; q++) {
      const int m = nn_list[q], km = node_kind[m], fm = first_unknown[m];
      const int ncol = K.num_unknowns[km] - (nop ? K.num_pressure[km] : 0);
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
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < num_rows; row += nwarp) {
    const long long k0 = rowstart[row], c0 = rowptr[row];
    const int len = (int)(rowptr[row + 1] - c0), d = dpos[row];
    for (int t = lane; t < len; t += 32) v[c0 + t] = t < d ? a[k0 + t] : (t == d ? a[row] : a[k0 + t - 1]);
  }
}

extern "C" int goma_gpu_csr_structure(goma_gpu_ctx *c, const goma_gpu_problem *p, goma_gpu_csr *out) {
  if (!c || !p || !out) return fail(-2, "null argument");
  CU(cudaSetDevice(c->device));
  const int nrows = c->num_owned_unknowns;
  if (!c->d_csr_rowptr) {
    if (c->pat.nn_ptr.empty()) return fail(-2, "node-node lists are not available");
    const long long msr0 = c->pat.rowstart[0];
    c->csr_nnz = nrows > 0 ? (long long)(c->pat.rowstart[nrows] - msr0) + nrows : 0;
    CsrKinds K;
    memset(&K, 0, sizeof(K));
    for (int k = 0; k < GOMA_GPU_MAX_KINDS; k++) {
      K.tslot[k] = -1;
      if (k >= p->num_kinds) continue;
      K.num_unknowns[k] = p->kind_num_unknowns[k];
      K.num_pressure[k] = kind_num_pressure(*p, k);
      K.tslot[k] = p->energy ? p->kind_slot[k][GOMA_SLOT_T] : -1;
    }
    long long *d_nn_ptr = nullptr;
    int *d_nn_list = nullptr;
    CU(cudaMalloc((void **)&d_nn_ptr, c->pat.nn_ptr.size() * sizeof(long long)));
    CU(cudaMalloc((void **)&d_nn_list, std::max<size_t>(c->pat.nn_list.size(), 1) * sizeof(int)));
    CU(cudaMemcpy(d_nn_ptr, c->pat.nn_ptr.data(), c->pat.nn_ptr.size() * sizeof(long long), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_nn_list, c->pat.nn_list.data(), c->pat.nn_list.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMalloc((void **)&c->d_csr_rowptr, ((size_t)nrows + 1) * sizeof(long long)));
    CU(cudaMalloc((void **)&c->d_csr_colind, std::max<size_t>((size_t)c->csr_nnz, 1) * sizeof(int)));
    CU(cudaMalloc((void **)&c->d_csr_dpos, std::max<size_t>((size_t)nrows, 1) * sizeof(int)));
    CU(cudaMalloc((void **)&c->d_csr_values, std::max<size_t>((size_t)c->csr_nnz, 1) * sizeof(double)));
    c->device_bytes += (size_t)c->csr_nnz * 12 + (size_t)nrows * 12;
    CU(cudaMemset(c->d_csr_rowptr, 0, ((size_t)nrows + 1) * sizeof(long long)));
    const int nown = c->prob.num_owned_nodes;
    if (nown > 0 && nrows > 0) {
      csr_structure_kernel<<<(nown + 127) / 128, 128, 0, c->stream>>>(nown, d_nn_ptr, d_nn_list, c->d_first, c->d_kind, K,
                                                                      c->d_rowstart, msr0, c->d_csr_rowptr,
                                                                      c->d_csr_colind, c->d_csr_dpos, nrows);
      CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_nn_ptr);
    cudaFree(d_nn_list);
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
