/*
 * goma_gpu_fill.h -- C ABI of the B200 (sm_100a) implementation of Goma's
 * per-Newton-iteration assembly hot path.
 *
 * What each entry point replaces in the reference (paths under /root/reference):
 *
 *   goma_gpu_fill_init      one-off snapshot of the host globals the element loop reads
 *                           (src/mm_fill.c:317-3785 reads pd/mp/upd/ei/Nodes/Coor/
 *                           Proc_Elem_Connect; src/mm_fill_ptrs.c:170,1136 load_ei /
 *                           load_elem_dofptr rebuild the gather tables per element, per
 *                           iteration -- here they are built once) and of the MSR graph
 *                           (src/mm_fill_util.c:2865 alloc_MSR_sparse_arrays, :3229
 *                           find_MSR_problem_graph; src/exo_conn.c:204 build_node_node).
 *   goma_gpu_fill           int matrix_fill_full(struct GomaLinearSolverData *ams, double x[],
 *                           double resid_vector[], double x_old[], double x_older[],
 *                           double xdot[], double xdot_old[], double x_update[],
 *                           double *ptr_delta_t, double *ptr_theta, ..., double *ptr_time_value,
 *                           Exo_DB *exo, Dpi *dpi, int *ptr_num_total_nodes,
 *                           dbl *ptr_h_elem_avg, dbl *ptr_U_norm, dbl *estifm)
 *                           include/mm_fill.h:41-58, src/mm_fill.c:158-312 (element loop,
 *                           flags) + :317 matrix_fill (one element) + :5175 load_lec (scatter).
 *   goma_gpu_fill_device    same, device-resident operands (hand-off to a GPU solver; the
 *                           `value` leg of bench.py).
 *   goma_gpu_global_h_U     global_h_elem_siz / global_velocity_norm,
 *                           src/mm_fill_aux.c:1128 and :612 (PSPG only,
 *                           src/mm_sol_nonlinear.c:1184-1192).
 *   goma_gpu_exchange_*     exchange_dof() (src/dp_comm.c:48-102): gather of list_dof_send (:77-80) and
 *                           the contiguous receive tail (:86-96), as one kernel that pulls the ghost
 *                           values out of the neighbours' HBM over NVLink (CUDA IPC peer pointers).
 *   goma_gpu_pack_dofs /    the two halves on their own, for a host that wants to drive the transport
 *   goma_gpu_unpack_dofs    itself (ncclSend/ncclRecv on the packed buffer).
 *   goma_gpu_fill_destroy   (no counterpart: the reference never frees lec/ei.)
 *
 * Conventions kept from the reference: return 0 on success, -1 on "domain failure"
 * (neg_elem_volume / neg_lub_height / zero_detJ, src/mm_fill.c:285-311) with the
 * three flags reported; any other problem is an error (<= -2) with a message from
 * goma_gpu_last_error() -- there is NO CPU fallback.  resid_vector and the matrix
 * values are what matrix_fill_full leaves in caller-zeroed storage
 * (src/mm_sol_nonlinear.c:1109-1121): this library overwrites them.
 * All indices are 0-based ints as in the reference; matrix positions are 64-bit.
 */
#ifndef GOMA_GPU_FILL_H
#define GOMA_GPU_FILL_H

#ifdef __cplusplus
extern "C" {
#endif

/* element types (values = nodes per element; reference: include/el_elm.h) */
#define GOMA_GPU_QUAD4 4
#define GOMA_GPU_QUAD9 9
#define GOMA_GPU_HEX8 8
#define GOMA_GPU_HEX27 27

/* per-node unknown "slots", in the order variables appear inside a node
 * (increasing variable id, include/rf_fem_const.h:174-200; species expanded) */
enum {
  GOMA_SLOT_U = 0, GOMA_SLOT_V, GOMA_SLOT_W, GOMA_SLOT_T,
  GOMA_SLOT_Y0, GOMA_SLOT_Y1, GOMA_SLOT_Y2, GOMA_SLOT_Y3,
  GOMA_SLOT_DX, GOMA_SLOT_DY, GOMA_SLOT_DZ, GOMA_SLOT_P,
  GOMA_NSLOT
};

#define GOMA_GPU_MAX_KINDS 4

/* pressure interpolation */
#define GOMA_PRESSURE_P1 1 /* discontinuous {1,s,t[,u]} on the centroid node (I_P1) */
#define GOMA_PRESSURE_EQ 2 /* same basis as velocity (Q1/Q1, needs PSPG)            */

/* mp_glob[mn] / elc_glob[mn] constants of one material (CONSTANT models only), see the fields of the same names below */
struct goma_gpu_material {
  double rho, mu, conductivity, heat_capacity, volume_expansion, reference_temperature;
  double diffusivity[4];
  double momentum_source[3];
  int momentum_source_model;
  double heat_source;
  double lame_mu, lame_lambda;
};

/* Plain-C snapshot of the host state the element loop depends on. */
struct goma_gpu_problem {
  /* mesh: Exo_DB / rd_mesh.c:397-507 globals */
  int dim;
  int elem_type; /* GOMA_GPU_* */
  int num_nodes; /* dpi->num_universe_nodes (owned + external)       */
  int num_owned_nodes; /* internal + boundary; rows of nodes >= this are never written */
  int num_elems; /* all local elements, owned and ghost (mm_fill.c:224) */
  const int *elem_connect; /* Proc_Elem_Connect, [num_elems * npe] */
  const double *coord[3]; /* Coor[dim][node] */

  /* unknown map: Nodes[n]->First_Unknown and the nodal variable layout */
  int num_unknowns; /* NumUnknowns + NumExtUnknowns */
  const int *first_unknown; /* [num_nodes] */
  int num_kinds; /* distinct nodal layouts (<= GOMA_GPU_MAX_KINDS) */
  const unsigned char *node_kind; /* [num_nodes] */
  /* kind_slot[k][s] = offset of slot s inside a node of kind k, or -1;
   * for GOMA_PRESSURE_P1 the P slot is the first of dim+1 consecutive dofs */
  int kind_slot[GOMA_GPU_MAX_KINDS][GOMA_NSLOT];
  int kind_num_unknowns[GOMA_GPU_MAX_KINDS];

  /* optional host MSR graph (ams->bindx); when given it must equal the one this
   * library derives (checked bit-exactly) -- pass NULL to let the library build it */
  const int *ija;

  /* physics switches: pd->e / pd->v, upd */
  int pressure_interp; /* GOMA_PRESSURE_* */
  int energy; /* R_ENERGY active */
  int num_species; /* upd->Max_Num_Species_Eqn */
  int ale; /* mesh1..dim active, pd->MeshMotion == ARBITRARY */
  int transient; /* pd->TimeIntegration != STEADY */
  int pspg; /* 0 off, 1 global ("yes"), 2 local */
  double ps_scaling;

  /* pd->etm[imtrx][eqn][LOG2_*]; a zero multiplier means the term's bit is off */
  double etm_momentum[6]; /* mass advection boundary diffusion source porous */
  double etm_continuity[2]; /* advection(div) source */
  double etm_energy[5]; /* mass advection boundary diffusion source */
  double etm_species[5];
  double etm_mesh[5];

  /* mp / elc constants (CONSTANT models only) */
  double rho, mu, conductivity, heat_capacity, volume_expansion, reference_temperature;
  double diffusivity[4];
  double momentum_source[3]; /* Navier-Stokes Source vector */
  int momentum_source_model; /* 0 CONSTANT, 1 BOUSS (hydrostatic part kept), 2 BOUSSINESQ (mm_std_models.c:125-360) */
  double heat_source;
  double lame_mu, lame_lambda;

  /* Dirichlet table: Nodes[]->DBC, BC_Types[].BC_Data_Float[0], BC_relax (bc_dirich.c:86-140) */
  const unsigned char *dbc_flag; /* [num_unknowns] 0 none, 1 residual = x - value, 2 hard set (residual 0) */
  const double *dbc_value; /* [num_unknowns] */

  /* exo->num_elem_blocks and upd->Num_Mat of the host.  Several element blocks / materials are assembled when
   * they share ONE element type and ONE set of active equations and differ in the material constants only (the
   * reference's loop picks mp = mp_glob[Matilda[ebn]] per element block, src/mm_fill.c:224-235, 621-640):
   * `materials[m]` then overrides the scalar constants above for the elements with elem_material[e] == m.
   * With num_materials <= 1 (or 0 = not stated) both pointers may be NULL.  Anything else (blocks of different
   * element types, materials with different equations) is refused by goma_gpu_fill_init. */
  int num_elem_blocks;
  int num_materials;
  const int *elem_material;                 /* [num_elems] material index = Matilda[block of the element]; NULL = all 0 */
  const struct goma_gpu_material *materials; /* [num_materials]; NULL = the scalar constants above for every element */

  /* Layout of the assembled values.  GOMA_GPU_LAYOUT_MSR (0): the reference's ams->val -- diagonal a[0..N), a[N]
   * unused, off-diagonals of row r at a[ija[r]..ija[r+1]) (src/mm_fill_util.c:2865-3031).  GOMA_GPU_LAYOUT_CSR (1):
   * CSR of the OWNED rows with the diagonal at its sorted position -- the element blocks are scattered straight into
   * the array a GPU solver or an Epetra/Tpetra CRS matrix consumes (the target of SumIntoGlobalValues,
   * src/linalg/sparse_matrix_epetra.cpp:111-118), no second copy of the matrix: goma_gpu_fill's `a` and d_a then
   * hold goma_gpu_csr::nnz values, goma_gpu_csr_structure's d_values IS d_a and goma_gpu_csr_values is a no-op. */
  int matrix_layout;

  /* Host-buffer calls (goma_gpu_fill with a in host memory) are bound by the copy of the matrix to the host.  With
   * host_stream_chunks = K > 1 the elements are swept in K chunks of consecutive elements (all colours of a chunk
   * before the next chunk) and the rows no later chunk touches leave for the host while the later chunks are still
   * being assembled (the part of mm_fill.c:224 the host waits for shrinks from the whole loop to one chunk).
   * Effective when element and node numbering advance together (any mesh generator's default order); always
   * correct, the values are bit-identical.  0 or 1 = one sweep, one copy.  Ignored on sub-domains with ghost
   * nodes (their classes are split interior / border for the halo exchange instead). */
  int host_stream_chunks;
};
#define GOMA_GPU_LAYOUT_MSR 0
#define GOMA_GPU_LAYOUT_CSR 1

typedef struct goma_gpu_ctx goma_gpu_ctx;

int goma_gpu_fill_init(const struct goma_gpu_problem *problem, int device, goma_gpu_ctx **ctx);
void goma_gpu_fill_destroy(goma_gpu_ctx *ctx);

/* MSR graph as built by the library: nnz_plus = ija[num_unknowns].  export_msr writes the
 * whole ija[0..nnz_plus) (needs the host arrays of `problem` again; fails beyond 2^31-1). */
int goma_gpu_fill_get_msr(goma_gpu_ctx *ctx, long long *nnz_plus);
/* doubles in the value array (`a` of goma_gpu_fill, d_a): nnz_plus + 1 for the MSR layout, goma_gpu_csr::nnz for CSR */
int goma_gpu_fill_value_count(goma_gpu_ctx *ctx, long long *count);
int goma_gpu_fill_export_msr(goma_gpu_ctx *ctx, const struct goma_gpu_problem *problem, int *ija_out);

/* Host-only (no device): the MSR graph the library derives from mesh + unknown map, i.e. what
 * find_MSR_problem_graph (src/mm_fill_util.c:3229) builds; ija_out may be NULL to size it. */
int goma_gpu_pattern_msr(const struct goma_gpu_problem *problem, long long *nnz_plus, int *ija_out);

/* Host-buffer call: H2D of the state vectors, assembly, D2H of a[0..nnz_plus] and resid. */
int goma_gpu_fill(goma_gpu_ctx *ctx, const double *x, const double *x_old, const double *x_older,
                  const double *xdot, const double *xdot_old, double delta_t, double theta,
                  double time_value, double h_elem_avg, double U_norm, int assemble_residual,
                  int assemble_jacobian, double *a, double *resid_vector, int flags_out[3]);

/* Device-resident call.  Pointers returned by goma_gpu_fill_device_buffers stay valid for the
 * life of the context; the caller fills d_x (and friends) and reads d_a / d_resid. */
struct goma_gpu_device_buffers {
  double *d_x, *d_x_old, *d_x_older, *d_xdot, *d_xdot_old;
  double *d_a; /* [nnz_plus + 1] MSR values */
  double *d_resid; /* [num_unknowns] */
  void *stream; /* cudaStream_t all work is ordered on */
};
int goma_gpu_fill_device_buffers(goma_gpu_ctx *ctx, struct goma_gpu_device_buffers *out);
int goma_gpu_fill_device(goma_gpu_ctx *ctx, double delta_t, double theta, double time_value,
                         double h_elem_avg, double U_norm, int assemble_residual,
                         int assemble_jacobian, int flags_out[3]);

/* Asynchronous form for a GPU solver pipeline: enqueue the assembly on the context's stream and return.
 * *done_event receives a cudaEvent_t (owned by the context) recorded behind the last assembly kernel;
 * goma_gpu_fill_wait synchronises, returns what goma_gpu_fill_device would have and reports the flags. */
int goma_gpu_fill_device_async(goma_gpu_ctx *ctx, double delta_t, double theta, double time_value,
                               double h_elem_avg, double U_norm, int assemble_residual,
                               int assemble_jacobian, void **done_event);
int goma_gpu_fill_wait(goma_gpu_ctx *ctx, int flags_out[3]);

/* global_h_elem_siz / global_velocity_norm (src/mm_fill_aux.c:1128-1207, :612-680) local sums over the
 * elements listed as owned and the owned nodes, from the device-resident x:
 *   sums_out = { sum_e sqrt(sum_p hsquared[p] / dim), number of elements summed,
 *                sum of squared velocity unknowns on owned nodes, number of those unknowns }.
 * The host all-reduces (SUM) and divides (mm_fill_aux.c:1194-1204).  elem_owned may be NULL (all). */
int goma_gpu_global_h_U(goma_gpu_ctx *ctx, const unsigned char *elem_owned, double sums_out[4]);

/* exchange_dof halves: gather x[list[k]] into buf (device), and the ghost tail pointer */
int goma_gpu_pack_dofs(goma_gpu_ctx *ctx, const double *d_vec, const int *d_list, int n, double *d_buf);
int goma_gpu_unpack_dofs(goma_gpu_ctx *ctx, double *d_vec, const int *d_list, int n, const double *d_buf);

/* The two O(nnz) passes that follow matrix_fill_full in solve_nonlinear_problem, on the device-resident
 * system (so the 45 GB matrix need not cross PCIe to be scaled):
 *   goma_gpu_row_sum_scale   row_sum_scaling_scale -> row_sum_scale_MSR (src/sl_matrix_util.c:441,507-600,
 *                            call site src/mm_sol_nonlinear.c:1317): over the owned rows, scale = sum |a_row|
 *                            with the sign of the diagonal; a_row /= scale, resid /= scale.  scale_out (host,
 *                            [owned unknowns]) may be NULL; the device copy is at goma_gpu_scale_buffer().
 *                            zero_rows_out counts rows whose sum is 0 (the reference warns and goes on).
 *   goma_gpu_vector_norms    Loo_norm / L1_norm / L2_norm (src/mm_sol_nonlinear.c:3320,3275,3177; call sites
 *                            :1451-1453) of the owned part of a device vector: which 0 = resid, 1 = x,
 *                            2 = xdot.  out = { max |v|, sum |v|, sum v^2, index of the max }: local values,
 *                            the host applies MPI_MAXLOC / MPI_SUM and the square root as the reference does. */
int goma_gpu_row_sum_scale(goma_gpu_ctx *ctx, double *scale_out, int *zero_rows_out);
int goma_gpu_scale_buffer(goma_gpu_ctx *ctx, double **d_scale, int *num_owned_unknowns);
int goma_gpu_vector_norms(goma_gpu_ctx *ctx, int which, double out[4]);

/* Device-resident hand-off to a GPU linear solver: a CSR view of the owned rows of the system --
 * 64-bit row pointers, sorted 32-bit column indices with the diagonal in place (what cuSPARSE / AmgX / Ginkgo /
 * an Epetra-Tpetra CRS view take; the reference's own plug-in point is the GomaSparseMatrix table,
 * include/linalg/sparse_matrix.h:35-83).  goma_gpu_csr_structure builds rowptr / colind on the device from the
 * node-node lists, exactly the column order of find_MSR_problem_graph (src/mm_fill_util.c:3229-3445) with the
 * diagonal merged in, and allocates the value array; goma_gpu_csr_values re-gathers the values from the MSR
 * storage after a fill (one read + one write of the matrix, HBM-bound).  Pointers stay valid for the context. */
struct goma_gpu_csr {
  int num_rows;        /* owned unknowns */
  long long nnz;       /* entries of those rows, diagonal included */
  long long *d_rowptr; /* [num_rows + 1] */
  int *d_colind;       /* [nnz] local unknown numbers (external columns included); NULL from goma_gpu_csr_rows */
  double *d_values;    /* [nnz] */
};
int goma_gpu_csr_structure(goma_gpu_ctx *ctx, const struct goma_gpu_problem *problem, struct goma_gpu_csr *out);
/* The same without the 4-byte-per-entry column array (C3 at 2M elements: 75 GB that do not fit beside the 150 GB
 * of values): row pointers and values only.  The columns of a row are then described by the node-level lists every
 * row of a node shares -- goma_gpu_node_graph: the sorted neighbour nodes of each node (exo_conn.c build_node_node),
 * whose unknowns, in node order, are the row's columns (energy rows skip pressure unknowns). */
int goma_gpu_csr_rows(goma_gpu_ctx *ctx, struct goma_gpu_csr *out);
int goma_gpu_node_graph(goma_gpu_ctx *ctx, long long **d_nn_ptr, int **d_nn_list);
int goma_gpu_csr_values(goma_gpu_ctx *ctx);
/* w = A v with the device-resident matrix of the last fill (either layout): the product the Newton line search takes
 * right after a fill (src/mm_sol_nonlinear.c:442-449, AZ_MSR_matvec_mult / GomaSparseMatrix::matrix_vector_mult,
 * include/linalg/sparse_matrix.h:76-78).  d_v: device vector over all local unknowns (owned + external columns),
 * d_w: device vector, the entries of the owned rows are written.  Needs no column-index array: the columns come
 * from the node-level neighbour lists (goma_gpu_node_graph). */
int goma_gpu_matvec(goma_gpu_ctx *ctx, const double *d_v, double *d_w);

/* exchange_dof() (src/dp_comm.c:48-102) over NVLink peer memory, one rank per GPU of one node.
 * Every rank exports CUDA IPC handles of its state vectors and of a small flag block
 * (goma_gpu_exchange_export); the host passes them round once (any transport) and each rank opens its
 * neighbours' (goma_gpu_exchange_setup), handing over, per neighbour, the dof indices IN THE NEIGHBOUR'S
 * numbering whose values fill this rank's contiguous external tail -- i.e. the neighbour's
 * list_dof_send block for this rank (src/dp_map_comm_vec.c:224-461), in its order.
 * goma_gpu_exchange_dof then is ONE kernel on the context's stream: it publishes "my vector of this epoch
 * is complete" in the neighbours' flag blocks, waits for theirs and pulls the ghost values straight out of
 * the neighbours' HBM.  No host synchronisation, no staging buffer.  The caller must not overwrite a
 * vector again before every rank has finished the fill that follows the exchange (in Goma the linear
 * solve, a collective, sits in between).  The call is collective over the neighbourhood, like the reference's:
 * a rank whose neighbour never calls it waits in the kernel.  which: 0 = x, 1 = xdot, 2 = x_old. */
#define GOMA_GPU_IPC_HANDLE_BYTES 64
#define GOMA_GPU_MAX_NEIGHBORS 32
struct goma_gpu_exchange_handles {
  unsigned char vec[3][GOMA_GPU_IPC_HANDLE_BYTES]; /* x, xdot, x_old */
  unsigned char flags[GOMA_GPU_IPC_HANDLE_BYTES];
  int device; /* CUDA device ordinal of the exporting rank */
};
int goma_gpu_exchange_export(goma_gpu_ctx *ctx, struct goma_gpu_exchange_handles *out);
int goma_gpu_exchange_setup(goma_gpu_ctx *ctx, int num_neighbors,
                            const struct goma_gpu_exchange_handles *neighbor_handles, /* [num_neighbors] */
                            const int *my_slot_at_neighbor, /* [num_neighbors] my index in that rank's neighbour list */
                            const int *recv_ptr,  /* [num_neighbors + 1] offsets into recv_list */
                            const int *recv_list, /* neighbour-local dof indices, tail order */
                            int tail_begin /* first external unknown = num owned unknowns */);
int goma_gpu_exchange_dof(goma_gpu_ctx *ctx, int which);
/* The wait inside the exchange kernel is bounded (about 20 s of SM clock): a neighbour that never publishes its
 * epoch -- a mismatched collective -- raises an error word instead of hanging the device.  This call (and every
 * goma_gpu_fill* that follows an exchange) returns -4 with a message when that happened, 0 otherwise. */
int goma_gpu_exchange_status(goma_gpu_ctx *ctx);
/* Write-after-read fence for the OWNER of the values: enqueues, on the context's stream, a wait until every
 * neighbour has pulled this rank's vector `which` of the last exchange.  Call it before anything on that stream
 * overwrites the vector when no collective (in Goma: the linear solve) separates the exchange from the update.
 * Bounded like the exchange itself; a time-out surfaces through goma_gpu_exchange_status / the next fill. */
int goma_gpu_exchange_fence(goma_gpu_ctx *ctx, int which);

/* timing / accounting of the last goma_gpu_fill*: device ms of the assembly kernel(s)
 * (CUDA events on the context's stream) and number of kernel launches */
int goma_gpu_fill_last_stats(goma_gpu_ctx *ctx, double *kernel_ms, int *launches);
/* wall seconds of goma_gpu_fill_init: out = { total, validation of the snapshot (host), uploads, sparsity pattern +
 * colouring + first-touch masks (device), tables + state / matrix allocation + gather records } */
int goma_gpu_fill_setup_stats(goma_gpu_ctx *ctx, double out[5]);

/* options: "scatter" = 0 fp64 atomics into zeroed storage | 1 coloured load+add+store | 2 coloured first-touch
 * stores (default; no memset of the matrix, bit-reproducible);  "grid_limit" = cap on resident CTAs (tests);
 * "rezero" = 1: the next first-touch fill zeroes the whole device storage first -- needed after anything other
 * than this library changed slots of d_a that no element touches (a solver factorising in place); the library sets
 * it itself after a row-sum scaling that met a zero row;  "accumulate" = 1: goma_gpu_fill uploads the caller's a /
 * resid_vector and ADDS the assembly to them, the literal semantics of the reference (src/mm_fill.c:5390,5463) for
 * a host that pre-loads the residual; the default (0) overwrites, which equals the reference for the caller-zeroed
 * storage of src/mm_sol_nonlinear.c:1109-1121 and saves a 45 GB upload;  "exchange_timeout_ms" = bound of the wait
 * inside goma_gpu_exchange_dof, in milliseconds of a 2 GHz SM clock. */
int goma_gpu_fill_set_option(goma_gpu_ctx *ctx, const char *name, int value);

const char *goma_gpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
