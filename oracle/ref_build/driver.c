/* goma_ref_fill -- drives the UNMODIFIED reference C sources (compiled from
 * /root/reference/src by oracle/ref_build/Makefile) through one or more
 * matrix_fill_full() calls on an in-memory mesh.
 *
 * TEST INFRASTRUCTURE ONLY: this is the parity oracle and the CPU baseline.
 * Nothing in goma_b200/ links, loads or executes it.
 *
 * The call sequence mirrors the reference's own main() (src/main.c:431-777:
 * *_alloc, read_input_file, read_mesh_exoII, setup_pd, assembly_alloc,
 * pre_process, bf_init, setup_problem) and the parts of solve_problem()
 * (src/rf_solve.c:655-816, 917) and solve_nonlinear_problem()
 * (src/mm_sol_nonlinear.c:1108-1281) that bracket the fill.
 *
 * usage: goma_ref_fill <workdir> map            -> map.bin
 *        goma_ref_fill <workdir> fill [nrep]    -> fill_out.bin   (reads state.bin)
 * <workdir> holds: input (Goma deck), *.mat, mesh.bin.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "mpi.h"
#include "std.h"
#include "dp_types.h"
#include "dp_utils.h"
#include "dpi.h"
#include "exo_struct.h"
#include "el_elm.h"
#include "mm_as.h"
#include "mm_as_alloc.h"
#include "mm_as_structs.h"
#include "mm_bc.h"
#include "mm_eh.h"
#include "mm_fill.h"
#include "mm_fill_aux.h"
#include "mm_fill_util.h"
#include "mm_input.h"
#include "mm_mp.h"
#include "mm_prob_def.h"
#include "mm_sol_nonlinear.h"
#include "sl_matrix_util.h"
#include "mm_unknown_map.h"
#include "rd_exo.h"
#include "rd_mesh.h"
#include "rf_allo.h"
#include "rf_bc.h"
#include "rf_bc_const.h"
#include "rf_fem.h"
#include "rf_fem_const.h"
#include "rf_io.h"
#include "rf_io_const.h"
#include "rf_masks.h"
#include "rf_node_const.h"
#include "rf_pre_proc.h"
#include "rf_solve.h"
#include "rf_solver.h"
#include "rf_util.h"
#include "rf_vars_const.h"
#include "sl_util_structs.h"

extern void fake_exodus_set_mesh(int, int, const double *, const double *, const double *, int, int,
                                 const char *, const int *, int, const int *, const int *, const int *);
extern void fake_exodus_set_blocks(int, const int *);
extern void fake_exodus_set_side_sets(int, const int *, const int *, const int *, const int *);
extern char **Argv;
extern int Argc;
extern double time_goma_started;
extern Comm_Ex **cx;
extern int PSPG;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void *xread(FILE *f, size_t n, size_t sz) {
  void *p = malloc(n * sz + 8);
  if (n && fread(p, sz, n, f) != n) {
    fprintf(stderr, "goma_ref_fill: short read\n");
    exit(2);
  }
  return p;
}

static void load_mesh(const char *fn) {
  FILE *f = fopen(fn, "rb");
  if (!f) {
    perror(fn);
    exit(2);
  }
  int h[6];
  char et[32];
  if (fread(h, sizeof(int), 6, f) != 6 || fread(et, 1, 32, f) != 32)
    exit(2);
  int dim = h[0], nn = h[1], ne = h[2], npe = h[3], nns = h[4], nsl = h[5];
  double *x = xread(f, nn, 8), *y = dim > 1 ? xread(f, nn, 8) : NULL, *z = dim > 2 ? xread(f, nn, 8) : NULL;
  int *conn = xread(f, (size_t)ne * npe, 4);
  int *ns_ids = xread(f, nns, 4), *ns_ptr = xread(f, nns + 1, 4), *ns_nodes = xread(f, nsl, 4);
  /* optional trailer: element blocks (consecutive elements), one material each */
  int nb = 0, *counts = NULL;
  if (fread(&nb, sizeof(int), 1, f) == 1) {
    if (nb > 0) counts = xread(f, nb, 4);
  } else
    nb = 0;
  /* second optional trailer: side sets -- count, ids, ptr[count+1], then 1-based element and EXODUS side numbers */
  int nss = 0, *ss_ids = NULL, *ss_ptr = NULL, *ss_elem = NULL, *ss_side = NULL;
  if (fread(&nss, sizeof(int), 1, f) == 1 && nss > 0) {
    ss_ids = xread(f, nss, 4);
    ss_ptr = xread(f, nss + 1, 4);
    ss_elem = xread(f, ss_ptr[nss], 4);
    ss_side = xread(f, ss_ptr[nss], 4);
  } else
    nss = 0;
  fclose(f);
  fake_exodus_set_mesh(dim, nn, x, y, z, ne, npe, et, conn, nns, ns_ids, ns_ptr, ns_nodes);
  if (nb > 0)
    fake_exodus_set_blocks(nb, counts);
  if (nss > 0)
    fake_exodus_set_side_sets(nss, ss_ids, ss_ptr, ss_elem, ss_side);
}

static Exo_DB *exo;
static Dpi *dpi;

static void reference_setup(void) {
  static char *argv0[] = {"goma_ref_fill", NULL};
  int argc = 1;
  char **argv = argv0;
  MPI_Init(&argc, &argv);
  time_goma_started = MPI_Wtime();
  Argv = argv0;
  Argc = 1;
  MPI_Comm_size(MPI_COMM_WORLD, &Num_Proc);
  MPI_Comm_rank(MPI_COMM_WORLD, &ProcID);
  Dim = 0;

  strcpy(Input_File, "input");
  strcpy(Echo_Input_File, "echo_input");
  ECHO("OPEN", Echo_Input_File);

  GOMA_EH(pd_alloc(), "pd_alloc");
  GOMA_EH(mp_alloc(), "mp_alloc");
  GOMA_EH(gn_alloc(), "gn_alloc");
  GOMA_EH(ve_alloc(), "ve_alloc");
  GOMA_EH(elc_alloc(), "elc_alloc");
  GOMA_EH(elc_rs_alloc(), "elc_rs_alloc");
  GOMA_EH(cr_alloc(), "cr_alloc");
  GOMA_EH(evp_alloc(), "evp_alloc");
  GOMA_EH(tran_alloc(), "tran_alloc");
  GOMA_EH(eigen_alloc(), "eigen_alloc");
  GOMA_EH(cont_alloc(), "cont_alloc");
  GOMA_EH(loca_alloc(), "loca_alloc");
  GOMA_EH(efv_alloc(), "efv_alloc");

  read_input_file(NULL, 0);

  EXO_ptr = alloc_struct_1(Exo_DB, 1);
  init_exo_struct(EXO_ptr);
  DPI_ptr = alloc_struct_1(Dpi, 1);
  init_dpi_struct(DPI_ptr);
  exo = EXO_ptr;
  dpi = DPI_ptr;
  if (read_mesh_exoII(exo, dpi) < 0) {
    fprintf(stderr, "read_mesh_exoII failed\n");
    exit(3);
  }
  GOMA_EH(setup_pd(), "setup_pd");
  GOMA_EH(evp_tensor_alloc(exo), "evp_tensor_alloc");
  GOMA_EH(assembly_alloc(exo), "assembly_alloc");
  pre_process(exo);
  GOMA_EH(bf_init(exo), "bf_init");
  cx = malloc(sizeof(Comm_Ex *) * upd->Total_Num_Matrices);
  for (int i = 0; i < upd->Total_Num_Matrices; i++)
    cx[i] = NULL;
  (void)setup_problem(exo, dpi);
  pg->imtrx = 0;
}

static void wr(FILE *f, const void *p, size_t n, size_t sz) {
  if (n && fwrite(p, sz, n, f) != n) {
    fprintf(stderr, "goma_ref_fill: short write\n");
    exit(2);
  }
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <workdir> map|fill [nrep]\n", argv[0]);
    return 2;
  }
  if (chdir(argv[1]) != 0) {
    perror(argv[1]);
    return 2;
  }
  const char *mode = argv[2];
  int nrep = argc > 3 ? atoi(argv[3]) : 1;
  load_mesh("mesh.bin");
  double t0 = now_s();
  reference_setup();

  const int imtrx = 0;
  int numProcUnknowns = NumUnknowns[imtrx] + NumExtUnknowns[imtrx];
  int num_total_nodes = dpi->num_universe_nodes;

  /* matrix allocation, MSR only (rf_solve.c:770-800) */
  if (strcmp(Matrix_Format, "msr") != 0) {
    fprintf(stderr, "oracle driver supports 'Matrix storage format = msr' only (got %s)\n", Matrix_Format);
    return 3;
  }
  struct GomaLinearSolverData *ams = alloc_struct_1(struct GomaLinearSolverData, 1);
  int *ija = NULL, *ija_attic = NULL;
  double *a = NULL, *a_old = NULL;
  int *node_to_fill = alloc_int_1(num_total_nodes, 0);
  alloc_MSR_sparse_arrays(&ija, &a, &a_old, 0, node_to_fill, exo, dpi);
  alloc_extern_ija_buffer(num_universe_dofs[imtrx], num_internal_dofs[imtrx] + num_boundary_dofs[imtrx], ija,
                          &ija_attic);
  ams->GomaMatrixData = NULL;
  ams->bindx = ija;
  ams->val = a;
  ams->belfry = ija_attic;
  ams->val_old = a_old;
  ams->indx = ams->bpntr = ams->rpntr = ams->cpntr = NULL;
  ams->npn = dpi->num_internal_nodes + dpi->num_boundary_nodes;
  ams->npn_plus = ams->npn + dpi->num_external_nodes;
  ams->npu = num_internal_dofs[imtrx] + num_boundary_dofs[imtrx];
  ams->npu_plus = num_universe_dofs[imtrx];
  ams->nnz = ija[ams->npu] - 1;
  ams->nnz_plus = ija[num_universe_dofs[imtrx]];
  int N = num_universe_dofs[imtrx];
  int nnz_plus = ams->nnz_plus;
  double t_setup = now_s() - t0;

  double *x = alloc_dbl_1(numProcUnknowns, 0.0), *x_old = alloc_dbl_1(numProcUnknowns, 0.0);
  double *x_older = alloc_dbl_1(numProcUnknowns, 0.0), *xdot = alloc_dbl_1(numProcUnknowns, 0.0);
  double *xdot_old = alloc_dbl_1(numProcUnknowns, 0.0), *x_update = alloc_dbl_1(2 * numProcUnknowns, 0.0);
  double *resid = alloc_dbl_1(numProcUnknowns, 0.0), *scale = alloc_dbl_1(numProcUnknowns, 0.0);
  pg->matrices = malloc(sizeof(struct Matrix_Data));
  pg->matrices[imtrx].ams = ams;
  pg->matrices[imtrx].x = x;
  pg->matrices[imtrx].x_old = x_old;
  pg->matrices[imtrx].x_older = x_older;
  pg->matrices[imtrx].xdot = xdot;
  pg->matrices[imtrx].xdot_old = xdot_old;
  pg->matrices[imtrx].x_update = x_update;
  pg->matrices[imtrx].scale = scale;
  pg->matrices[imtrx].resid_vector = resid;

  /* solve_problem always runs find_and_set_Dirichlet before the first fill (rf_solve.c:917): it
   * builds Nodes[]->DBC, which put_dirichlet_in_matrix consumes.  Done here on scratch vectors so
   * the caller's state is only touched when it asks for the preset. */
  {
    double *xs = alloc_dbl_1(numProcUnknowns, 0.0), *xds = alloc_dbl_1(numProcUnknowns, 0.0);
    find_and_set_Dirichlet(xs, xds, exo, dpi);
    free(xs);
    free(xds);
  }

  if (strcmp(mode, "map") == 0) {
    /* unknown map + sparsity + Dirichlet table: the bit-exact contract */
    FILE *f = fopen("map.bin", "wb");
    int nn = exo->num_nodes;
    int hdr[8] = {numProcUnknowns, N, nnz_plus, nn, exo->num_elems, Num_BC, PSPG, upd->Max_Num_Species_Eqn};
    wr(f, hdr, 8, 4);
    int *fu = malloc((nn + 1) * sizeof(int));
    for (int i = 0; i < nn; i++)
      fu[i] = Nodes[i]->First_Unknown[imtrx];
    fu[nn] = numProcUnknowns;
    wr(f, fu, nn + 1, 4);
    wr(f, ija, nnz_plus + 1, 4);
    /* per-unknown (node, variable type, sub-index) as Index_Solution sees them */
    int *idv_out = malloc(3 * (size_t)numProcUnknowns * sizeof(int));
    for (int i = 0; i < numProcUnknowns; i++) {
      idv_out[3 * i] = idv[imtrx][i][0];
      idv_out[3 * i + 1] = idv[imtrx][i][1];
      idv_out[3 * i + 2] = idv[imtrx][i][2];
    }
    wr(f, idv_out, 3 * (size_t)numProcUnknowns, 4);
    /* Dirichlet: x after find_and_set_Dirichlet on a NaN-marked vector */
    double *xm = alloc_dbl_1(numProcUnknowns, 0.0), *xd = alloc_dbl_1(numProcUnknowns, 0.0);
    for (int i = 0; i < numProcUnknowns; i++)
      xm[i] = -7.77e77;
    find_and_set_Dirichlet(xm, xd, exo, dpi);
    wr(f, xm, numProcUnknowns, 8);
    /* per-unknown DBC index (-1 = none) from Nodes[]->DBC (bc_dirich.c:86-100) */
    int *dbc = malloc((size_t)numProcUnknowns * sizeof(int));
    for (int i = 0; i < numProcUnknowns; i++)
      dbc[i] = -1;
    for (int n = 0; n < nn; n++) {
      NODE_INFO_STRUCT *node = Nodes[n];
      if (node->DBC[imtrx]) {
        int nu = (n + 1 < nn ? Nodes[n + 1]->First_Unknown[imtrx] : numProcUnknowns) - node->First_Unknown[imtrx];
        for (int o = 0; o < nu; o++)
          dbc[node->First_Unknown[imtrx] + o] = node->DBC[imtrx][o];
      }
    }
    wr(f, dbc, numProcUnknowns, 4);
    /* Inter_Mask rows/cols for variable ids 0..9 (v, T, Y, d, -, P) */
    int im[100];
    for (int r = 0; r < 10; r++)
      for (int c = 0; c < 10; c++)
        im[10 * r + c] = Inter_Mask[imtrx][r][c];
    wr(f, im, 100, 4);
    fclose(f);
    printf("map: unknowns=%d N=%d nnz_plus=%d setup_s=%.3f\n", numProcUnknowns, N, nnz_plus, t_setup);
    return 0;
  }

  /* ---- fill mode ---- */
  FILE *fs = fopen("state.bin", "rb");
  if (!fs) {
    perror("state.bin");
    return 2;
  }
  int sh[4]; /* n_unknowns, n_states, assemble_jacobian, apply_dirichlet_preset */
  double sp[5]; /* delta_t, theta, time, h_elem_avg (<0: compute), U_norm (<0: compute) */
  if (fread(sh, 4, 4, fs) != 4 || fread(sp, 8, 5, fs) != 5)
    return 2;
  if (sh[0] != numProcUnknowns) {
    fprintf(stderr, "state.bin has %d unknowns, problem has %d\n", sh[0], numProcUnknowns);
    return 3;
  }
  FILE *fo = fopen("fill_out.bin", "wb");
  int oh[4] = {numProcUnknowns, nnz_plus, sh[1], N};
  wr(fo, oh, 4, 4);
  double delta_t = sp[0], theta = sp[1], time_value = sp[2];
  tran->time_value = time_value;
  tran->delta_t = delta_t;
  tran->theta = theta;
  for (int s = 0; s < sh[1]; s++) {
    if (fread(x, 8, numProcUnknowns, fs) != (size_t)numProcUnknowns ||
        fread(x_old, 8, numProcUnknowns, fs) != (size_t)numProcUnknowns ||
        fread(x_older, 8, numProcUnknowns, fs) != (size_t)numProcUnknowns ||
        fread(xdot, 8, numProcUnknowns, fs) != (size_t)numProcUnknowns ||
        fread(xdot_old, 8, numProcUnknowns, fs) != (size_t)numProcUnknowns)
      return 2;
    if (sh[3])
      find_and_set_Dirichlet(x, xdot, exo, dpi);
    double h_elem_avg = 0., U_norm = 0.;
    double best = 1e300, total = 0.;
    int err = 0;
    for (int rep = 0; rep < nrep; rep++) {
      init_vec_value(resid, 0.0, numProcUnknowns);
      init_vec_value(a, 0.0, nnz_plus + 1);
      /* mm_sol_nonlinear.c:1184-1192 */
      if (upd->matrix_index[VELOCITY1] == pg->imtrx && (PSPG && Num_Var_In_Type[pg->imtrx][PRESSURE])) {
        h_elem_avg = sp[3] >= 0 ? sp[3] : global_h_elem_siz(x, x_old, xdot, resid, exo, dpi);
        U_norm = sp[4] >= 0 ? sp[4] : global_velocity_norm(x, exo, dpi);
      }
      af->Assemble_Residual = TRUE;
      af->Assemble_Jacobian = sh[2] ? TRUE : FALSE;
      af->Assemble_LSA_Jacobian_Matrix = FALSE;
      af->Assemble_LSA_Mass_Matrix = FALSE;
      double t1 = now_s();
      err = matrix_fill_full(ams, x, resid, x_old, x_older, xdot, xdot_old, x_update, &delta_t, &theta,
                             First_Elem_Side_BC_Array[pg->imtrx], &time_value, exo, dpi, &num_total_nodes,
                             &h_elem_avg, &U_norm, NULL);
      double dt = now_s() - t1;
      total += dt;
      if (dt < best)
        best = dt;
    }
    int flags[4] = {err, neg_elem_volume, neg_lub_height, zero_detJ};
    double tm[4] = {best, total / nrep, h_elem_avg, U_norm};
    wr(fo, flags, 4, 4);
    wr(fo, tm, 4, 8);
    wr(fo, x, numProcUnknowns, 8);
    wr(fo, a, nnz_plus + 1, 8);
    wr(fo, resid, numProcUnknowns, 8);
    printf("fill[%d]: err=%d elems=%d best_s=%.6f mean_s=%.6f elems_per_s=%.1f\n", s, err, exo->num_elems, best,
           total / nrep, exo->num_elems / best);
    if (argc > 4 && !strcmp(argv[4], "post")) {
      /* what solve_nonlinear_problem does next: row-sum scaling (mm_sol_nonlinear.c:1317), then the norms of
       * the scaled residual (:1451-1453) */
      double *scale = (double *)calloc(numProcUnknowns, sizeof(double));
      char dofname_r[80];
      int num_unk_r = -1;
      row_sum_scaling_scale(ams, resid, scale);
      double nrm[4];
      nrm[0] = Loo_norm(resid, NumUnknowns[pg->imtrx], &num_unk_r, dofname_r);
      nrm[1] = L1_norm(resid, NumUnknowns[pg->imtrx]);
      nrm[2] = L2_norm(resid, NumUnknowns[pg->imtrx]);
      nrm[3] = (double)num_unk_r;
      FILE *fp = fopen("post_out.bin", s == 0 ? "wb" : "ab");
      wr(fp, nrm, 4, 8);
      wr(fp, scale, numProcUnknowns, 8);
      wr(fp, a, nnz_plus + 1, 8);
      wr(fp, resid, numProcUnknowns, 8);
      fclose(fp);
      free(scale);
    }
  }
  fclose(fo);
  fclose(fs);
  return 0;
}
