// Private to the library: the context behind goma_gpu_ctx, error plumbing and small device-memory helpers shared
// by the translation units (goma_gpu_fill.cu: init + assembly; exchange.cu: exchange_dof over peer memory;
// post_fill.cu: PSPG norms, row-sum scaling, residual norms, CSR hand-off).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/goma_gpu_fill.h"
#include "pattern.h"

namespace goma_b200 {
int fail(int code, const std::string &msg);  // records the message for goma_gpu_last_error(), returns code
}
using goma_b200::fail;

#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(-3, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                          std::to_string(__LINE__));                                           \
  } while (0)

namespace goma_b200 {
struct KindInfo {  // per node kind: unknowns, of which pressure (last in the node), offset of T (-1: none)
  int nunk[GOMA_GPU_MAX_KINDS], npress[GOMA_GPU_MAX_KINDS], tslot[GOMA_GPU_MAX_KINDS];
};
// device-resident topology built by build_pattern_device (pattern_gpu.cu)
struct DevPattern {
  int *ne_ptr = nullptr, *ne_list = nullptr;            // node -> elements, ascending
  long long *nn_ptr = nullptr;                          // node-node lists (exo_conn.c build_node_node), sorted
  int *nn_list = nullptr;
  unsigned short *cum_full = nullptr, *cum_p = nullptr;  // per list entry: unknowns / pressure unknowns of earlier neighbours
  int *row_full = nullptr, *row_p = nullptr;            // per node: unknowns / pressure unknowns of all neighbours
  unsigned *pair_first = nullptr, *node_first = nullptr;  // first-touch masks
  unsigned *cls = nullptr;                              // class ((border, chunk, colour) key) of each element
  long long nn_total = 0;
};
}  // namespace goma_b200
struct goma_gpu_ctx;
namespace goma_b200 {
KindInfo make_kind_info(const goma_gpu_problem &p);
int build_pattern_device(goma_gpu_ctx *c);
void free_device_pattern(goma_gpu_ctx *c, bool keep_node_node);
int download_pattern(goma_gpu_ctx *c, Pattern &out);
int build_csr_dpos(goma_gpu_ctx *c);  // post_fill.cu: diagonal offsets of the CSR layout (needs the init-time lists)
}  // namespace goma_b200

struct goma_gpu_ctx {
  goma_gpu_problem prob;  // scalar members + kind tables only; pointers are not retained
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  goma_b200::DevPattern dpat;
  long long nnz_plus = 0;          // == ija[N] of the MSR graph
  int layout = 0;                  // GOMA_GPU_LAYOUT_*: how d_a is laid out
  long long a_len = 0;             // doubles in d_a: nnz_plus + 1 (MSR) or csr_nnz (CSR)
  int *d_work = nullptr;           // one hand-out counter per class launch (dynamic element distribution)
  int matvec_cap = 0;              // option "matvec_cap": columns staged per node in goma_gpu_matvec (0 = default)
  int rss_max_row = -1;            // longest owned row (row-sum scaling sizes its staging buffer with it); -1 = not yet known
  int *d_dpos = nullptr;           // CSR layout: offset of the diagonal inside each owned row
  int num_colours = 0;             // classes per chunk = colours x materials; class = (((border ? nchunk : 0) + chunk) * ncol + colour) * nmat + material
  int num_mats = 1;                // materials (goma_gpu_problem::num_materials, at least 1)
  int *d_elem_mat = nullptr;       // [num_elems] material of each element (NULL with one material)
  std::vector<goma_gpu_material> mats;  // host copy of goma_gpu_problem::materials
  int num_chunks = 1;              // chunks of consecutive elements swept one after the other (L2-sized; 1 = off)
  int chunk_elems = 0;             // elements per chunk in effect (pattern build)
  int chunk_elems_option = 0;      // "chunk_elems" option at init: 0 auto, > 0 elements per chunk, < 0 off
  // host_stream_chunks: rows [0, chunk_done_row[k]) are complete once chunk k has been assembled; chunk_done_off[k] is
  // where the entries of that row start in d_a.  Copies run on cstream behind ev_chunk[k].
  int stream_chunks = 0;
  std::vector<long long> chunk_done_row, chunk_done_off;
  std::vector<cudaEvent_t> ev_chunk;
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_copy = nullptr;
  int first_border_class = 0;      // classes from here on touch external nodes: they wait for the ghost exchange
  cudaStream_t xstream = nullptr;  // exchange_dof runs here, overlapped with the interior classes
  cudaEvent_t ev_x = nullptr, ev_pre = nullptr;
  bool exchange_in_flight = false;
  // device arrays
  int *d_conn = nullptr, *d_first = nullptr;
  double *d_coord[3] = {nullptr, nullptr, nullptr};
  unsigned char *d_kind = nullptr, *d_dbc_flag = nullptr;
  double *d_dbc_value = nullptr;
  long long *d_rowstart = nullptr;
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
  long long exchange_spin_limit = 40000000000LL;  // ~20 s of SM clock: bounded wait in exchange_dof_kernel
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
  bool rezero = false;       // zero the whole storage before the next first-touch fill
  bool accumulate = false;   // goma_gpu_fill adds into the caller's a / resid_vector
  bool preloaded = false;    // (internal) d_a / d_resid hold uploaded host values for this fill
  bool fill_pending = false; // goma_gpu_fill_device_async launched, goma_gpu_fill_wait not yet called
  double last_ms = 0.0;
  int last_launches = 0;
  size_t device_bytes = 0;
  double setup_s[5] = {0, 0, 0, 0, 0};  // goma_gpu_fill_init: total, validation, uploads, pattern (device), tables + state + records
};
