// exchange_dof() (src/dp_comm.c:48-102) on the device: the peer-memory pull kernel over CUDA IPC handles, and the
// pack / unpack halves for a host that drives the transport itself.  See include/goma_gpu_fill.h.
#include <cstring>

#include "ctx.h"

using namespace goma_b200;

// ------------------------------------------------------------------ exchange_dof over peer memory
// flag block of a rank: [3][MAX_NEIGHBORS] epochs "neighbour k's vector is complete" (written by the neighbours),
// [3][MAX_NEIGHBORS] epochs "neighbour k has pulled my vector" (written by the neighbours), one error word
constexpr int XFLAG_DONE = 3 * GOMA_GPU_MAX_NEIGHBORS, XFLAG_ERR = 6 * GOMA_GPU_MAX_NEIGHBORS, XFLAG_WORDS = XFLAG_ERR + 1;

struct ExchangeArgs {
  int nn;
  unsigned long long epoch;
  unsigned long long *peer_ready[GOMA_GPU_MAX_NEIGHBORS];  // the slot of this rank in each neighbour's flag block
  const unsigned long long *my_ready;                      // this rank's flag block, row of the vector
  const double *peer_vec[GOMA_GPU_MAX_NEIGHBORS];
  int recv_ptr[GOMA_GPU_MAX_NEIGHBORS + 1];
  const int *recv_list;
  double *tail;
  unsigned long long *error;    // set when a neighbour never publishes this epoch (mismatched collective)
  long long spin_limit;         // clock64 ticks a thread waits before it gives up
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
      const long long t0 = clock64();
      bool ok = true;
      do {  // the neighbour's vector of this epoch is complete
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(A.my_ready + nb) : "memory");
        if (seen >= A.epoch) break;
        if (clock64() - t0 > A.spin_limit) {  // bounded wait: report instead of hanging the device
          ok = false;
          *A.error = 1ull + (unsigned long long)nb;
          break;
        }
      } while (true);
      if (ok) A.tail[k] = A.peer_vec[nb][A.recv_list[k]];
    }
  }
}

// after the pull (stream order): tell every neighbour that its vector of this epoch has been read
__global__ void exchange_done_kernel(const __grid_constant__ ExchangeArgs A) {
  if (threadIdx.x < A.nn) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.peer_ready[threadIdx.x] + XFLAG_DONE), "l"(A.epoch) : "memory");
  }
}
// the owner's side: wait until every neighbour has pulled the vector of the current epoch
__global__ void exchange_fence_kernel(int nn, const unsigned long long *done, unsigned long long epoch, long long spin_limit,
                                      unsigned long long *error) {
  if (threadIdx.x >= nn) return;
  const long long t0 = clock64();
  unsigned long long seen;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(done + threadIdx.x) : "memory");
    if (seen >= epoch) return;
  } while (clock64() - t0 <= spin_limit);
  *error = 1ull + (unsigned long long)threadIdx.x;
}

extern "C" int goma_gpu_exchange_export(goma_gpu_ctx *c, goma_gpu_exchange_handles *out) {
  if (!c || !out) return fail(-2, "null argument");
  CU(cudaSetDevice(c->device));
  if (!c->d_xflags) {  // [3][MAX_NEIGHBORS] epochs published by the neighbours + one error word
    CU(cudaMalloc((void **)&c->d_xflags, XFLAG_WORDS * sizeof(unsigned long long)));
    CU(cudaMemset(c->d_xflags, 0, XFLAG_WORDS * sizeof(unsigned long long)));
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
  c->num_neighbors = 0;
  // a (re-)set-up starts a new generation on every rank: epochs and the flag block restart from zero.  The host
  // must put a barrier between the set-up of all ranks and the first exchange (dp_comm.setup_peer_exchange does),
  // as it must between the reference's set_dof_communication and the first exchange_dof.
  for (int v = 0; v < 3; v++) c->epoch[v] = 0;
  CU(cudaMemset(c->d_xflags, 0, XFLAG_WORDS * sizeof(unsigned long long)));
  if (tail_begin < 0 || tail_begin > c->prob.num_unknowns) return fail(-2, "tail_begin outside the vector");
  if (num_neighbors && recv_ptr[0] != 0) return fail(-2, "recv_ptr[0] must be 0");
  for (int k = 0; k < num_neighbors; k++)
    if (recv_ptr[k + 1] < recv_ptr[k]) return fail(-2, "recv_ptr must be non-decreasing");
  c->tail_begin = tail_begin;
  c->recv_ptr.assign(recv_ptr, recv_ptr + num_neighbors + 1);
  if (num_neighbors == 0) c->recv_ptr.assign(1, 0);
  if ((long long)tail_begin + c->recv_ptr[num_neighbors] > c->prob.num_unknowns) return fail(-2, "external tail exceeds the vector");
  for (int k = 0; k < c->recv_ptr[num_neighbors]; k++)
    if (recv_list[k] < 0) return fail(-2, "negative dof index in recv_list");
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
  c->num_neighbors = num_neighbors;
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
  A.error = c->d_xflags + XFLAG_ERR;
  A.spin_limit = c->exchange_spin_limit;
  const int total = c->recv_ptr[A.nn];
  // The pull runs on its own stream, behind everything enqueued so far on the context's stream (the solver update
  // that produced the vector); the next fill assembles its interior classes meanwhile and makes its border classes
  // wait for ev_x.  Few blocks: a block waiting for a late neighbour holds registers a fill CTA could use.
  const int threads = 256, blocks = std::max(1, std::min(32, (total + threads - 1) / threads));
  CU(cudaEventRecord(c->ev_pre, c->stream));
  CU(cudaStreamWaitEvent(c->xstream, c->ev_pre, 0));
  exchange_dof_kernel<<<blocks, threads, 0, c->xstream>>>(A);
  CU(cudaGetLastError());
  exchange_done_kernel<<<1, GOMA_GPU_MAX_NEIGHBORS, 0, c->xstream>>>(A);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev_x, c->xstream));
  c->exchange_in_flight = true;
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
extern "C" int goma_gpu_exchange_fence(goma_gpu_ctx *c, int which) {
  if (!c) return fail(-2, "null context");
  if (which < 0 || which > 2) return fail(-2, "which must be 0 (x), 1 (xdot) or 2 (x_old)");
  if (c->num_neighbors == 0 || c->epoch[which] == 0) return 0;
  CU(cudaSetDevice(c->device));
  exchange_fence_kernel<<<1, GOMA_GPU_MAX_NEIGHBORS, 0, c->stream>>>(c->num_neighbors, c->d_xflags + XFLAG_DONE + which * GOMA_GPU_MAX_NEIGHBORS,
                                                                  c->epoch[which], c->exchange_spin_limit, c->d_xflags + XFLAG_ERR);
  CU(cudaGetLastError());
  return 0;
}

extern "C" int goma_gpu_exchange_status(goma_gpu_ctx *c) {
  if (!c) return fail(-2, "null context");
  if (!c->d_xflags) return 0;
  CU(cudaSetDevice(c->device));
  unsigned long long e = 0;
  CU(cudaStreamSynchronize(c->xstream));
  CU(cudaMemcpyAsync(&e, c->d_xflags + XFLAG_ERR, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (e) return fail(-4, "exchange_dof: neighbour slot " + std::to_string(e - 1) + " never published its vector (timed out)");
  return 0;
}

extern "C" int goma_gpu_pack_dofs(goma_gpu_ctx *c, const double *d_vec, const int *d_list, int n, double *d_buf) {
  if (!c) return fail(-2, "null context");
  if (n <= 0) return 0;
  CU(cudaSetDevice(c->device));
  pack_dofs_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_vec, d_list, n, d_buf);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int goma_gpu_unpack_dofs(goma_gpu_ctx *c, double *d_vec, const int *d_list, int n, const double *d_buf) {
  if (!c) return fail(-2, "null context");
  if (n <= 0) return 0;
  CU(cudaSetDevice(c->device));
  unpack_dofs_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_vec, d_list, n, d_buf);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
