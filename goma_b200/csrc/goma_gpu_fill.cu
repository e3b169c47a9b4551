// C ABI of the sm_100a matrix_fill path (see include/goma_gpu_fill.h for the contract and
// the reference call sites each entry point replaces).  No torch, no CPU fallback: every
// failure is an error code plus goma_gpu_last_error().
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/goma_gpu_fill.h"
#include "fill_kernel.cuh"
#include "ctx.h"
#include "pattern.h"
#include "tables.h"

using namespace goma_b200;

static thread_local std::string g_err;
int goma_b200::fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

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
    return {fill_kernel_ws<C>, build_records_kernel<C>, C::TPE + C::NMUL, sizeof(Smem<C>), C::TBL_GLOBAL, sizeof(ElemRec<C>)};
  else
    return {fill_kernel<C>, build_records_kernel<C>, C::TPE, sizeof(Smem<C>), C::TBL_GLOBAL, sizeof(ElemRec<C>)};
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
    // the tensor-core write-out stores the fields of a node as one run: U, V, W(, T) must be consecutive unknowns
    // (they are in the reference's ordering: variables by increasing id, include/rf_fem_const.h:174-200)
    for (int k = 0; k < p.num_kinds; k++) {
      const int u = p.kind_slot[k][GOMA_SLOT_U];
      if (u < 0) continue;
      if (p.kind_slot[k][GOMA_SLOT_V] != u + 1 || p.kind_slot[k][GOMA_SLOT_W] != u + 2 ||
          (p.energy && p.kind_slot[k][GOMA_SLOT_T] != u + 3))
        return fail(-2, "velocity (and temperature) unknowns must be consecutive inside a node");
    }
    // default: 16 padded node blocks on the tensor cores, tensor-core set-up phases, two CTAs per SM (the fastest of
    // the A/B runs under profiles/r2*_variants*); GOMA_GPU_VARIANT selects the others for experiments
    static const int var = getenv("GOMA_GPU_VARIANT") ? atoi(getenv("GOMA_GPU_VARIANT")) : 0;
    static const bool ws = getenv("GOMA_GPU_WS") ? atoi(getenv("GOMA_GPU_WS")) != 0 : false;
    // (variant numbers: see Cfg::VAR; 1xx = two CTAs per SM at 128 registers instead of three at 80)
    if (p.energy)
      k = var == 1     ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 3, false, 1>>()
          : var == 100 ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2, false, 0>>()
          : var == 101 ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2, false, 1>>()
          : var == 105 ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2, false, 5>>()
          : var == 104 ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2, false, 4>>()
          : var == 108 ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2, false, 8>>()
          : var == 3   ? entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 3, false, 0>>()
                       : entry<Cfg<3, 27, 27, true, true, 0, false, 256, 3, 2, false, 1>>();
    else if (ws)
      k = entry<Cfg<3, 27, 27, true, false, 0, false, 192, 3, 1, true>>();
    else
      k = var == 1     ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 3, false, 1>>()
          : var == 100 ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2, false, 0>>()
          : var == 101 ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2, false, 1>>()
          : var == 105 ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2, false, 5>>()
          : var == 104 ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2, false, 4>>()
          : var == 108 ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2, false, 8>>()
          : var == 3   ? entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 3, false, 0>>()
                       : entry<Cfg<3, 27, 27, true, false, 0, false, 256, 3, 2, false, 1>>();
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
  if (p.num_nodes < 0 || p.num_elems < 0 || p.num_unknowns < 0) return fail(-2, "negative size in goma_gpu_problem");
  // several element blocks / materials: one element type, one set of equations, constants per material (the header);
  // a host with more than one block or material must say which element has which material
  if (p.num_materials > 64) return fail(-2, "more than 64 materials");
  if (p.num_materials > 1 && (!p.elem_material || !p.materials))
    return fail(-2, "num_materials > 1 needs elem_material and materials (blocks of different element types or equations are not supported)");
  if (p.num_elem_blocks > 1 && p.num_materials <= 1 && !p.elem_material)
    return fail(-2, "several element blocks: state num_materials, elem_material and materials (Matilda[ebn], mm_fill.c:224-235)");
  if (p.num_materials > 1)
    for (int e = 0; e < p.num_elems; e++)
      if (p.elem_material[e] < 0 || p.elem_material[e] >= p.num_materials)
        return fail(-2, "elem_material[" + std::to_string(e) + "] out of range");
  if (p.matrix_layout != GOMA_GPU_LAYOUT_MSR && p.matrix_layout != GOMA_GPU_LAYOUT_CSR)
    return fail(-2, "matrix_layout must be GOMA_GPU_LAYOUT_MSR or GOMA_GPU_LAYOUT_CSR");

  // ---- unknown map: everything build_pattern / build_records_kernel index with must be consistent
  const int dim = p.dim, npe = p.elem_type;
  for (int k = 0; k < p.num_kinds; k++) {
    const int nu = p.kind_num_unknowns[k];
    if (nu < 0 || nu > 64) return fail(-2, "kind_num_unknowns out of range");
    const int np_ = p.pressure_interp == GOMA_PRESSURE_P1 ? dim + 1 : 1;
    for (int sl = 0; sl < GOMA_NSLOT; sl++) {
      const int off = p.kind_slot[k][sl];
      if (off < -1 || off + (sl == GOMA_SLOT_P && off >= 0 ? np_ : 1) > nu + (off < 0 ? 1 : 0))
        return fail(-2, "kind_slot does not fit kind_num_unknowns");
    }
  }
  bool used[GOMA_GPU_MAX_KINDS] = {false, false, false, false};
  for (int n = 0; n < p.num_nodes; n++) {
    const int kd = p.node_kind[n];
    if (kd >= p.num_kinds) return fail(-2, "node_kind[" + std::to_string(n) + "] >= num_kinds");
    used[kd] = true;
    const int fu = p.first_unknown[n];
    if (fu < 0 || (long long)fu + p.kind_num_unknowns[kd] > p.num_unknowns)
      return fail(-2, "first_unknown[" + std::to_string(n) + "] + its unknowns exceed num_unknowns");
    if (n + 1 < p.num_nodes && p.first_unknown[n + 1] != fu + p.kind_num_unknowns[kd])
      return fail(-2, "first_unknown is not the running sum of the nodal unknown counts at node " + std::to_string(n));
  }
  for (int u = 0; u < p.num_unknowns; u++)
    if (p.dbc_flag[u] > 2) return fail(-2, "dbc_flag must be 0, 1 or 2");
  // every field the selected kernel gathers must exist on every node kind in use (a node kind without T or Y,
  // i.e. a variable not defined on all nodes, would be gathered from the neighbouring unknown)
  const bool p1 = p.pressure_interp == GOMA_PRESSURE_P1;
  for (int k = 0; k < p.num_kinds; k++) {
    if (!used[k]) continue;
    auto need = [&](int sl, const char *what) -> int {
      return p.kind_slot[k][sl] < 0 ? fail(-2, std::string("node kind ") + std::to_string(k) + " carries no " + what +
                                                   " unknown: variables must live on every node of the block")
                                    : 0;
    };
    for (int d = 0; d < dim; d++)
      if (int rc = need(GOMA_SLOT_U + d, "velocity")) return rc;
    if (p.energy)
      if (int rc = need(GOMA_SLOT_T, "temperature")) return rc;
    for (int w = 0; w < p.num_species; w++)
      if (int rc = need(GOMA_SLOT_Y0 + w, "species")) return rc;
    if (p.ale)
      for (int d = 0; d < dim; d++)
        if (int rc = need(GOMA_SLOT_DX + d, "mesh displacement")) return rc;
    if (!p1)
      if (int rc = need(GOMA_SLOT_P, "pressure")) return rc;
  }
  for (long long k = 0; k < (long long)p.num_elems * npe; k++)
    if (p.elem_connect[k] < 0 || p.elem_connect[k] >= p.num_nodes) return fail(-2, "element connectivity entry out of range");
  if (p1) {
    const int cen = npe == GOMA_GPU_QUAD9 ? 8 : 20;
    for (int e = 0; e < p.num_elems; e++)
      if (p.kind_slot[p.node_kind[p.elem_connect[(long long)e * npe + cen]]][GOMA_SLOT_P] < 0)
        return fail(-2, "P1 pressure: the centroid node of element " + std::to_string(e) + " carries no pressure unknowns");
  }
  return 0;
}

// host_stream_chunks: which rows are final after each chunk.  A row is final once no later chunk holds an element
// with a node at or before it: chunk_done_row[k] = min over the chunks after k of the lowest unknown they touch.
__global__ void chunk_min_row_kernel(int ne, int npe, int chunk_elems, const int *__restrict__ conn, const int *__restrict__ first_unknown,
                                     int *__restrict__ chunk_min) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int m = 0x7fffffff;
  for (int k = 0; k < npe; k++) m = min(m, first_unknown[conn[(size_t)e * npe + k]]);
  atomicMin(&chunk_min[e / chunk_elems], m);
}

static int setup_host_streaming(goma_gpu_ctx *c) {
  const goma_gpu_problem &p = c->prob;
  const int K = c->num_chunks, N = p.num_unknowns;
  int *d_min = nullptr;
  CU(cudaMalloc((void **)&d_min, K * sizeof(int)));
  CU(cudaMemsetAsync(d_min, 0x7f, K * sizeof(int), c->stream));
  chunk_min_row_kernel<<<(p.num_elems + 255) / 256, 256, 0, c->stream>>>(p.num_elems, p.elem_type, c->chunk_elems, c->d_conn, c->d_first, d_min);
  std::vector<int> h(K);
  cudaError_t e1 = cudaMemcpyAsync(h.data(), d_min, K * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e2 = cudaStreamSynchronize(c->stream);
  cudaFree(d_min);
  CU(e1);
  CU(e2);
  const long long msr0 = (long long)N + 1;
  const bool csr = c->layout == GOMA_GPU_LAYOUT_CSR;
  c->chunk_done_row.assign(K, c->num_owned_unknowns);
  c->chunk_done_off.assign(K, c->a_len);
  long long suffix = c->num_owned_unknowns;
  for (int k = K - 1; k >= 0; k--) {  // (the last chunk completes everything)
    c->chunk_done_row[k] = suffix;
    suffix = std::min<long long>(suffix, h[k]);
  }
  for (int k = 0; k + 1 < K; k++) {
    long long rs = 0;
    CU(cudaMemcpy(&rs, c->d_rowstart + c->chunk_done_row[k], sizeof(long long), cudaMemcpyDeviceToHost));
    c->chunk_done_off[k] = csr ? rs - msr0 + c->chunk_done_row[k] : rs;
  }
  CU(cudaStreamCreateWithFlags(&c->cstream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
  c->ev_chunk.assign(K, nullptr);
  for (int k = 0; k < K; k++) CU(cudaEventCreateWithFlags(&c->ev_chunk[k], cudaEventDisableTiming));
  c->stream_chunks = K;
  return 0;
}

extern "C" int goma_gpu_fill_init(const goma_gpu_problem *problem, int device, goma_gpu_ctx **out) {
  if (!problem || !out) return fail(-2, "null argument");
  *out = nullptr;
  const auto t_init0 = std::chrono::steady_clock::now();
  const goma_gpu_problem &p = *problem;
  if (int rc = validate(p)) return rc;
  KernelEntry ke;
  if (int rc = pick_kernel(p, ke)) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(-3, "no CUDA device: the goma_gpu_fill path has no CPU fallback");
  CU(cudaSetDevice(device));

  struct Guard {  // every early return below destroys the half-built context
    goma_gpu_ctx *c;
    ~Guard() { if (c) goma_gpu_fill_destroy(c); }
  } guard{new goma_gpu_ctx()};
  goma_gpu_ctx *c = guard.c;
  c->prob = p;
  c->device = device;
  if (const char *ce = getenv("GOMA_GPU_CHUNK_ELEMS")) c->chunk_elems_option = atoi(ce);  // experiments: 0 auto, < 0 off
  const bool want_stream = p.host_stream_chunks > 1 && p.num_owned_nodes >= p.num_nodes && p.num_elems > 0;  // (materials: fine, a chunk ends after its last (colour, material) class)
  if (want_stream) c->chunk_elems_option = (p.num_elems + p.host_stream_chunks - 1) / p.host_stream_chunks;
  c->num_owned_unknowns = p.num_owned_nodes < p.num_nodes ? p.first_unknown[p.num_owned_nodes] : p.num_unknowns;
  const int nn = p.num_nodes, ne = p.num_elems, npe = p.elem_type, N = p.num_unknowns;
  int rc = 0;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const auto t_upload = now();
  c->setup_s[1] = secs(t_init0, t_upload);  // validate()
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  CU(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_pre, cudaEventDisableTiming));
  if (p.num_materials > 1) {
    c->num_mats = p.num_materials;
    c->mats.assign(p.materials, p.materials + p.num_materials);
    rc |= upload(&c->d_elem_mat, p.elem_material, (size_t)ne, c);
    c->prob.elem_material = nullptr;  // (host arrays are not kept)
    c->prob.materials = nullptr;
  }
  rc |= upload(&c->d_conn, p.elem_connect, (size_t)ne * npe, c);
  rc |= upload(&c->d_first, p.first_unknown, nn, c);
  for (int d = 0; d < p.dim; d++) rc |= upload(&c->d_coord[d], p.coord[d], nn, c);
  rc |= upload(&c->d_kind, p.node_kind, nn, c);
  rc |= upload(&c->d_dbc_flag, p.dbc_flag, N, c);
  rc |= upload(&c->d_dbc_value, p.dbc_value, N, c);
  if (rc) return -3;
  // node-element / node-node lists, MSR row starts, element colouring and first-touch masks: built on the device
  // from the arrays just uploaded (pattern_gpu.cu)
  const auto t_pat = now();
  c->setup_s[2] = secs(t_upload, t_pat);
  if (int prc = build_pattern_device(c)) return prc;
  const auto t_pat1 = now();
  c->setup_s[3] = secs(t_pat, t_pat1);
  // optional bit-exact check against the host's own MSR graph
  if (p.ija) {
    if (c->nnz_plus > 2147483647LL) return fail(-2, "host ija given but nnz exceeds the 32-bit MSR limit");
    Pattern hp;
    if (int drc = download_pattern(c, hp)) return drc;
    std::vector<int> mine((size_t)c->nnz_plus + 1, 0);
    emit_msr_columns(p, hp, mine.data());
    for (long long k = 0; k < c->nnz_plus; k++) {
      if (mine[k] != p.ija[k]) return fail(-2, "host MSR graph differs from the derived one at ija[" + std::to_string(k) + "]");
    }
  }

  // quadrature / basis tables, packed in the order Smem<C>::tbl expects
  ElemTables t = make_tables(p.elem_type);
  std::vector<double> packed;
  packed.insert(packed.end(), t.wt.begin(), t.wt.end());
  packed.insert(packed.end(), t.psi.begin(), t.psi.end());
  packed.insert(packed.end(), t.l1d.begin(), t.l1d.end());
  packed.insert(packed.end(), t.dphi.begin(), t.dphi.end());
  packed.insert(packed.end(), t.phi.begin(), t.phi.end());
  packed.resize(ke.tbl_pad, 0.0);
  rc |= upload(&c->d_tables, packed.data(), packed.size(), c);

  rc |= dalloc(&c->d_x, N, c);
  rc |= dalloc(&c->d_x_old, N, c);
  rc |= dalloc(&c->d_x_older, N, c);
  rc |= dalloc(&c->d_xdot, N, c);
  rc |= dalloc(&c->d_xdot_old, N, c);
  rc |= dalloc(&c->d_resid, N, c);
  c->layout = p.matrix_layout;
  {
    // CSR of the owned rows: every row gains its diagonal, rows of external unknowns do not exist
    long long rs_owned = 0;
    const int no = c->num_owned_unknowns;
    CU(cudaMemcpy(&rs_owned, c->d_rowstart + no, sizeof(long long), cudaMemcpyDeviceToHost));
    c->csr_nnz = (rs_owned - ((long long)N + 1)) + no;
  }
  c->a_len = c->layout == GOMA_GPU_LAYOUT_CSR ? c->csr_nnz : c->nnz_plus + 1;
  rc |= dalloc(&c->d_a, (size_t)c->a_len, c);
  rc |= dalloc(&c->d_flags, 4, c);
  if (rc) return -3;

  // per-element gather records: everything load_elem_dofptr would recompute per element and per
  // iteration, gathered once; afterwards the pair tables they were built from are dropped
  {
    CU(cudaMalloc((void **)&c->d_erec, std::max<size_t>((size_t)ne * ke.rec_bytes, 16)));
    c->device_bytes += (size_t)ne * ke.rec_bytes;
    FillParams P;
    static_params(c, P);
    P.flags = c->d_flags;
    if (c->layout == GOMA_GPU_LAYOUT_CSR) {
      if (int drc = build_csr_dpos(c)) return drc;
    }
    if (ne > 0) ke.build_records<<<(ne + 127) / 128, 128, 0, c->stream>>>(P, ne);
    CU(cudaGetLastError());
    if (want_stream && c->num_chunks > 1)
      if (int src = setup_host_streaming(c)) return src;
    int h_flags[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(h_flags, c->d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h_flags[3]) return fail(-2, "sparsity pattern: Could not find vbl in sparse matrix");  // mm_fill.c:5462 wording
    // the node-node lists stay (CSR structure, ija export); everything else of the topology is dropped
    free_device_pattern(c, true);
  }

  if (getenv("GOMA_GPU_PROFILE")) rc |= dalloc(&c->d_prof, 2 * 8 * 4096, c);
  if (ke.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute((const void *)ke.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ke.smem);
    if (e != cudaSuccess) return fail(-3, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  }
  if (rc) return -3;
  c->setup_s[4] = secs(t_pat1, now());
  c->setup_s[0] = secs(t_init0, now());
  guard.c = nullptr;
  *out = c;
  return 0;
}

extern "C" void goma_gpu_fill_destroy(goma_gpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  void *ptrs[] = {c->d_conn, c->d_first, c->d_coord[0], c->d_coord[1], c->d_coord[2], c->d_kind, c->d_dbc_flag,
                  c->d_dbc_value, c->d_rowstart, c->d_prof, c->d_tables, c->d_erec, c->d_x, c->d_x_old,
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
  if (c->d_csr_values && c->d_csr_values != c->d_a) cudaFree(c->d_csr_values);
  if (c->d_dpos) cudaFree(c->d_dpos);
  if (c->d_elem_mat) cudaFree(c->d_elem_mat);
  if (c->d_work) cudaFree(c->d_work);
  if (c->d_scale) cudaFree(c->d_scale);
  if (c->d_partials) cudaFree(c->d_partials);
  if (c->d_zero_rows) cudaFree(c->d_zero_rows);
  free_device_pattern(c, false);
  for (cudaEvent_t e : c->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->cstream) cudaStreamDestroy(c->cstream);
  if (c->ev_x) cudaEventDestroy(c->ev_x);
  if (c->ev_pre) cudaEventDestroy(c->ev_pre);
  if (c->xstream) cudaStreamDestroy(c->xstream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int goma_gpu_fill_get_msr(goma_gpu_ctx *c, long long *nnz_plus) {
  if (!c) return fail(-2, "null context");
  if (nnz_plus) *nnz_plus = c->nnz_plus;
  return 0;
}

extern "C" int goma_gpu_fill_value_count(goma_gpu_ctx *c, long long *count) {
  if (!c || !count) return fail(-2, "null argument");
  *count = c->a_len;
  return 0;
}

// ija export needs the host arrays again (they are not retained in the context)
extern "C" int goma_gpu_fill_export_msr(goma_gpu_ctx *c, const goma_gpu_problem *p, int *ija_out) {
  if (!c || !p || !ija_out) return fail(-2, "null argument");
  if (c->nnz_plus > 2147483647LL) return fail(-2, "nnz exceeds the 32-bit MSR limit of ija");
  CU(cudaSetDevice(c->device));
  Pattern hp;
  if (int drc = download_pattern(c, hp)) return drc;
  emit_msr_columns(*p, hp, ija_out);
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
  if (!strcmp(name, "exchange_timeout_ms")) {  // bound of the wait inside goma_gpu_exchange_dof (default ~20 s)
    c->exchange_spin_limit = (long long)std::max(1, value) * 2000000LL;
    return 0;
  }
  if (!strcmp(name, "matvec_cap")) {  // goma_gpu_matvec: longest column list staged in shared memory (tests)
    c->matvec_cap = value;
    return 0;
  }
  if (!strcmp(name, "rezero")) {  // the next first-touch fill zeroes the whole matrix / residual storage first
    c->rezero = value != 0;
    return 0;
  }
  if (!strcmp(name, "accumulate")) {  // goma_gpu_fill adds into the caller's a / resid_vector (the reference's +=)
    c->accumulate = value != 0;
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
  P.nn_ptr = c->dpat.nn_ptr;
  P.nn_list = c->dpat.nn_list;
  P.cum_full = c->dpat.cum_full;
  P.cum_p = c->dpat.cum_p;
  P.pair_first = c->dpat.pair_first;
  P.node_first = c->dpat.node_first;
  P.dbc_flag = c->d_dbc_flag;
  P.dbc_value = c->d_dbc_value;
  P.num_owned_nodes = p.num_owned_nodes;
  P.erec = c->d_erec;
  P.csr = c->layout == GOMA_GPU_LAYOUT_CSR ? 1 : 0;
  P.msr0 = (long long)p.num_unknowns + 1;
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
#ifdef GOMA_PROFILE_PHASES
  CU(cudaMemcpyToSymbolAsync(g_store_debug, &P.debug, sizeof(int), 0, cudaMemcpyHostToDevice, c->stream));
#endif
  if (p.transient && !(delta_t > 0.0)) return fail(-2, "transient fill needs delta_t > 0");

  if (c->num_sms == 0) {
    int sms = 0, bps = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, (const void *)ke.fn, ke.tpe, ke.smem));
    if (bps < 1 || sms < 1) return fail(-3, "fill kernel does not fit on an SM");
    c->blocks_per_sm = bps;
    c->num_sms = sms;  // cached only once the geometry is known to be valid
  }
  int max_grid = c->num_sms * c->blocks_per_sm;
  if (c->grid_limit > 0) max_grid = std::min(max_grid, c->grid_limit);

  c->last_launches = 0;
  CU(cudaMemsetAsync(c->d_flags, 0, 4 * sizeof(int), c->stream));
  CU(cudaEventRecord(c->ev0, c->stream));
  int mode = c->scatter_mode;
  if (c->preloaded) mode = mode == 2 ? 1 : mode;  // accumulate into what the host uploaded: no store-over, no memset
  P.scatter_mode = mode;
  if (mode == 2 && c->rezero) {
    // slots no element touches are zero from init; something (a row-sum scaling that met a zero row, a solver
    // working in place) may have changed them since: zero everything once
    CU(cudaMemsetAsync(c->d_resid, 0, (size_t)p.num_unknowns * sizeof(double), c->stream));
    CU(cudaMemsetAsync(c->d_a, 0, (size_t)c->a_len * sizeof(double), c->stream));
    c->rezero = false;
  }
  if (mode != 2 && !c->preloaded) {
    // accumulate-into semantics need zeroed storage; the first-touch mode overwrites every slot the
    // elements touch and never writes the others (zeroed once at init), so it needs no memset
    if (assemble_residual) CU(cudaMemsetAsync(c->d_resid, 0, (size_t)p.num_unknowns * sizeof(double), c->stream));
    if (assemble_jacobian) CU(cudaMemsetAsync(c->d_a, 0, (size_t)c->a_len * sizeof(double), c->stream));
  }
  auto wait_for_exchange = [&]() -> int {  // the ghost values must have landed before an element reads them
    if (c->exchange_in_flight) {
      CU(cudaStreamWaitEvent(c->stream, c->ev_x, 0));
      c->exchange_in_flight = false;
    }
    return 0;
  };
  // dynamic hand-out of the elements inside a launch (fill_kernel): one counter per class, zeroed here
  static const bool static_stride = getenv("GOMA_GPU_STATIC") && atoi(getenv("GOMA_GPU_STATIC")) != 0;  // A/B runs
  // share of a launch handed out by block index before the counter takes over (per cent)
  static const int static_pct = getenv("GOMA_GPU_STATIC_PCT") ? std::max(0, std::min(100, atoi(getenv("GOMA_GPU_STATIC_PCT")))) : 75;
  const size_t nctr = std::max<size_t>(c->colour_begin.size(), 2);
  if (!static_stride) {
    if (!c->d_work) CU(cudaMalloc((void **)&c->d_work, nctr * sizeof(int)));
    CU(cudaMemsetAsync(c->d_work, 0, nctr * sizeof(int), c->stream));
  }
  P.work = nullptr;
  auto set_material = [&](int m) {  // the constants of material m (mp_glob[mn], elc_glob[mn]) for the next launch
    if (c->num_mats <= 1) return;
    const goma_gpu_material &M = c->mats[m];
    P.rho = M.rho;
    P.mu = M.mu;
    P.k = M.conductivity;
    P.Cp = M.heat_capacity;
    P.beta = M.volume_expansion;
    P.Tref = M.reference_temperature;
    P.heat_source = M.heat_source;
    for (int d = 0; d < 3; d++) P.g[d] = M.momentum_source[d];
    P.source_model = M.momentum_source_model;
    for (int w = 0; w < 4; w++) P.diffusivity[w] = M.diffusivity[w];
    P.lame_mu = M.lame_mu;
    P.lame_lambda = M.lame_lambda;
  };
  if (mode == 0 && c->num_mats <= 1) {
    if (int wrc = wait_for_exchange()) return wrc;
    P.elem_list = nullptr;
    P.elem_begin = 0;
    P.elem_end = p.num_elems;
    int grid = std::max(1, std::min(max_grid, p.num_elems));
    P.work = static_stride ? nullptr : c->d_work;
    P.static_rounds = std::max(1, (int)((long long)static_pct * p.num_elems / 100 / grid));
    ke.fn<<<grid, ke.tpe, ke.smem, c->stream>>>(P);
    c->last_launches++;
  } else {
    // classes in increasing order, one launch each: stream order is the inter-class barrier.  Classes
    // [0, first_border_class) hold the elements with owned nodes only: they run while exchange_dof is in flight
    // on its own stream; the border classes wait for it (hide the exchange behind the interior)
    P.elem_list = c->d_elem_list;
    for (size_t col = 0; col + 1 < c->colour_begin.size(); col++) {
      if ((int)col == c->first_border_class)
        if (int wrc = wait_for_exchange()) return wrc;
      P.elem_begin = c->colour_begin[col];
      P.elem_end = c->colour_begin[col + 1];
      int n = P.elem_end - P.elem_begin;
      if (n > 0) {
        set_material((int)(col % (size_t)c->num_mats));
        P.work = static_stride ? nullptr : c->d_work + col;
        int grid = std::max(1, std::min(max_grid, n));
        P.static_rounds = std::max(1, (int)((long long)static_pct * n / 100 / grid));
        ke.fn<<<grid, ke.tpe, ke.smem, c->stream>>>(P);
        c->last_launches++;
      }
      // host streaming: the last colour of a chunk is done -- its finished rows may leave for the host
      if (c->stream_chunks > 1 && (int)((col + 1) % c->num_colours) == 0 && (int)(col / c->num_colours) < c->stream_chunks)
        CU(cudaEventRecord(c->ev_chunk[col / c->num_colours], c->stream));
    }
  }
  if (int wrc = wait_for_exchange()) return wrc;  // (no border class: nothing read the tail, keep the order anyway)
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev1, c->stream));
  return 0;
}

static int finish_fill(goma_gpu_ctx *c, int flags_out[3]) {
  int h_flags[4] = {0, 0, 0, 0};
  unsigned long long xerr = 0;
  CU(cudaMemcpyAsync(h_flags, c->d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, c->stream));
  if (c->d_xflags && c->num_neighbors)
    CU(cudaMemcpyAsync(&xerr, c->d_xflags + 6 * GOMA_GPU_MAX_NEIGHBORS, sizeof(xerr), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->fill_pending = false;
  if (xerr) return fail(-4, "exchange_dof before this fill: neighbour slot " + std::to_string(xerr - 1) + " never published its vector");
  if (c->d_prof) {
    std::vector<long long> h(2 * 8 * 4096);
    CU(cudaMemcpy(h.data(), c->d_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double s[8] = {0}, s2[8] = {0};
    int nb = 0;
    for (int b = 0; b < 4096; b++)
      if (h[b * 8 + 6] > 0) {
        nb++;
        for (int k = 0; k < 6; k++) s[k] += (double)h[b * 8 + k] / (double)h[b * 8 + 6];
        for (int k = 0; k < 7; k++) s2[k] += (double)h[(4096 + b) * 8 + k] / (double)h[b * 8 + 6];
      }
    if (nb)
      fprintf(stderr, "[goma_gpu profile] build split: ale-x %.0f J %.0f inv %.0f grads %.0f fields %.0f gp+vg %.0f | DMMA part of the loop (warp 0) %.0f\n",
              s2[0] / nb, s2[1] / nb, s2[2] / nb, s2[3] / nb, s2[4] / nb, s2[5] / nb, s2[6] / nb);
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

// The same without the host synchronisation: everything is enqueued on the context's stream and the call returns.
// *done_event (a cudaEvent_t owned by the context) is recorded behind the last assembly kernel: a solver on another
// stream orders itself with cudaStreamWaitEvent.  goma_gpu_fill_wait collects the return code and the flags.
extern "C" int goma_gpu_fill_device_async(goma_gpu_ctx *c, double delta_t, double theta, double time_value,
                                          double h_elem_avg, double U_norm, int assemble_residual,
                                          int assemble_jacobian, void **done_event) {
  if (!c) return fail(-2, "null context");
  CU(cudaSetDevice(c->device));
  if (int rc = launch_fill(c, delta_t, theta, time_value, h_elem_avg, U_norm, assemble_residual, assemble_jacobian))
    return rc;
  c->fill_pending = true;
  if (done_event) *done_event = (void *)c->ev1;
  return 0;
}

extern "C" int goma_gpu_fill_wait(goma_gpu_ctx *c, int flags_out[3]) {
  if (!c) return fail(-2, "null context");
  if (!c->fill_pending) return fail(-2, "no asynchronous fill is pending");
  CU(cudaSetDevice(c->device));
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
  if (c->accumulate) {
    // the reference adds into caller-owned storage (mm_fill.c:5463 a[ja] +=, :5390 resid +=): upload what the
    // caller holds and accumulate on top of it
    if (assemble_jacobian)
      CU(cudaMemcpyAsync(c->d_a, a, (size_t)c->a_len * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (assemble_residual) CU(cudaMemcpyAsync(c->d_resid, resid_vector, nb, cudaMemcpyHostToDevice, c->stream));
    c->preloaded = true;
  }
  int lrc = launch_fill(c, delta_t, theta, time_value, h_elem_avg, U_norm, assemble_residual, assemble_jacobian);
  if (c->accumulate) {
    c->preloaded = false;
    c->rezero = true;  // the uploaded values sit in slots the first-touch mode never rewrites
  }
  if (lrc) return lrc;
  if (assemble_jacobian && c->stream_chunks > 1 && c->scatter_mode != 0) {
    // rows no later chunk touches go to the host while the later chunks are still being assembled (the launches above
    // are all enqueued; every copy waits for the event behind the last colour of its chunk)
    const bool csr = c->layout == GOMA_GPU_LAYOUT_CSR;
    const long long msr0 = (long long)c->prob.num_unknowns + 1;
    long long r_prev = 0, o_prev = csr ? 0 : msr0;
    for (int k = 0; k < c->stream_chunks; k++) {
      const bool last = k + 1 == c->stream_chunks;
      const long long r = last ? msr0 : c->chunk_done_row[k], o = last ? c->a_len : c->chunk_done_off[k];
      CU(cudaStreamWaitEvent(c->cstream, c->ev_chunk[k], 0));
      if (!csr && r > r_prev)  // MSR: the diagonal entries of those rows (and, at the end, the unused a[N])
        CU(cudaMemcpyAsync(a + r_prev, c->d_a + r_prev, (size_t)(r - r_prev) * sizeof(double), cudaMemcpyDeviceToHost, c->cstream));
      if (o > o_prev)
        CU(cudaMemcpyAsync(a + o_prev, c->d_a + o_prev, (size_t)(o - o_prev) * sizeof(double), cudaMemcpyDeviceToHost, c->cstream));
      r_prev = std::max(r_prev, r);
      o_prev = std::max(o_prev, o);
    }
    CU(cudaEventRecord(c->ev_copy, c->cstream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
  } else if (assemble_jacobian) {
    // one copy: the PCIe link is saturated by it (≈47 GB/s measured; two concurrent copy streams gave the same)
    CU(cudaMemcpyAsync(a, c->d_a, (size_t)c->a_len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  if (assemble_residual) CU(cudaMemcpyAsync(resid_vector, c->d_resid, nb, cudaMemcpyDeviceToHost, c->stream));
  return finish_fill(c, flags_out);
}

extern "C" int goma_gpu_fill_setup_stats(goma_gpu_ctx *c, double out[5]) {
  if (!c || !out) return fail(-2, "null argument");
  for (int k = 0; k < 5; k++) out[k] = c->setup_s[k];
  return 0;
}

extern "C" int goma_gpu_fill_last_stats(goma_gpu_ctx *c, double *kernel_ms, int *launches) {
  if (!c) return fail(-2, "null context");
  if (kernel_ms) *kernel_ms = c->last_ms;
  if (launches) *launches = c->last_launches;
  return 0;
}
