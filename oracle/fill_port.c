/* fill_port.c -- CPU restatement (the "port" oracle) of the reference's matrix_fill hot path.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline leg as a checker.  The product (goma_b200/) never links or calls it.
 *
 * Parity pinning: this restatement is checked against outputs of the reference itself
 * (oracle/_ref = the reference's unmodified C sources, see oracle/ref_build/) through the
 * committed fixtures tests/golden/*.npz and, where oracle/_ref is present, live.
 *
 * Structure follows the reference one element at a time, with a dense element block `lec`
 * exactly as src/mm_fill.c:317 matrix_fill does:
 *   gather             load_elem_dofptr        src/mm_fill_ptrs.c:1136
 *   Gauss loop         src/mm_fill.c:1253      find_stu/Gq_weight src/el_elm_info.c:1615,3469
 *   basis              shape()                 src/rf_shape.c:185,361,698,1105; P1 src/mm_fill_util.c:3802
 *   map                beer_belly              src/mm_fill_util.c:258-276 (J), :386-480 (detJ, B)
 *   gradients          load_bf_grad            src/mm_fill_util.c:1765-1776
 *   fields             load_fv, load_fv_grads  src/load_field_variables.c:128,2049
 *   momentum           assemble_momentum       src/mm_fill_momentum.c:534-662 (R), :1564-1735 (J_m_v),
 *                                              :746-915 (J_m_T), :2052-2117 (J_m_P); fluid_stress :3268-3271,
 *                                              :3458-3469, :3704; sources :3738, src/mm_std_models.c:125
 *   continuity         assemble_continuity     src/mm_fill_continuity.c:435-444 (R), :665-761 (J_c_v)
 *   energy             assemble_energy         src/mm_fill_energy.c:322-381 (R), :425-487 (J_e_T), :628-692 (J_e_v)
 *   species            assemble_mass_transport src/mm_fill_species.c:194- (Fickian, constant D; concentration
 *                                              form: coeff_rho = 1), get_continuous_species_terms :9739
 *   PSPG               calc_pspg               src/mm_fill_stabilization.c:852-1591 (tau :1030-1096, momentum
 *                                              residual :1281-1319, d_pspg :1321-1524); continuity terms
 *                                              src/mm_fill_continuity.c:604-613,746-756,~870,~1200;
 *                                              h_elem_siz src/mm_fill_aux.c:844, element_velocity :759
 *   mesh (ALE)         assemble_mesh           src/mm_fill_terms.c:421-428 (R, ARBITRARY), :529-576 (J_d_d);
 *                                              belly_flop src/mm_fill_solid.c:77-1120 (grad_d, Eulerian strain,
 *                                              volume change, neg_elem_volume :659-815); mesh_stress_tensor :3208-3287;
 *                                              mesh sensitivities J_m_d src/mm_fill_momentum.c:2200-2442,
 *                                              J_c_d src/mm_fill_continuity.c:1004-1148, in the closed forms
 *                                              d(grad_phi_i[p])/d(d_bj) = -grad_phi_j[p] grad_phi_i[b],
 *                                              d(detJ)/d(d_bj) = detJ grad_phi_j[b]  
 *   Dirichlet          put_dirichlet_in_matrix src/bc_dirich.c:44-151
 *   scatter            load_lec (MSR)          src/mm_fill.c:5241-5483 (in_list search :5461)
 * Cartesian coordinates only (h3 = 1, grad_phi_e[i][a][p][q] = delta_aq grad_phi[i][p],
 * src/mm_fill_util.c:1838-1871).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/goma_gpu_fill.h"

#define MAXN 27
#define MAXF 12 /* phi-interpolated fields per node */
#define MAXD (MAXN * MAXF + 4)

static void lagrange1d(int order, double s, double *L, double *dL) {
  if (order == 1) {
    L[0] = 0.5 * (1.0 - s); L[1] = 0.5 * (1.0 + s);
    dL[0] = -0.5; dL[1] = 0.5;
  } else {
    L[0] = -0.5 * s * (1.0 - s); L[1] = 1.0 - s * s; L[2] = 0.5 * s * (1.0 + s);
    dL[0] = -0.5 * (1.0 - 2.0 * s); dL[1] = -2.0 * s; dL[2] = 0.5 * (1.0 + 2.0 * s);
  }
}

/* lattice position (0..order per direction) of each local node, Exodus/PATRAN order */
static const int LAT4[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
static const int LAT9[9][3] = {{0, 0, 0}, {2, 0, 0}, {2, 2, 0}, {0, 2, 0}, {1, 0, 0}, {2, 1, 0}, {1, 2, 0}, {0, 1, 0}, {1, 1, 0}};
static const int LAT8[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int LAT27[27][3] = {
    {0, 0, 0}, {2, 0, 0}, {2, 2, 0}, {0, 2, 0}, {0, 0, 2}, {2, 0, 2}, {2, 2, 2}, {0, 2, 2}, {1, 0, 0},
    {2, 1, 0}, {1, 2, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 1}, {2, 2, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 2},
    {1, 2, 2}, {0, 1, 2}, {1, 1, 1}, {1, 1, 0}, {1, 1, 2}, {0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}};

static void basis(int et, const double xi[3], double *phi, double (*dphi)[3]) {
  int dim = (et == 4 || et == 9) ? 2 : 3, order = (et == 4 || et == 8) ? 1 : 2;
  const int(*lat)[3] = et == 4 ? LAT4 : et == 9 ? LAT9 : et == 8 ? LAT8 : LAT27;
  double L[3][3] = {{1, 0, 0}, {1, 0, 0}, {1, 0, 0}}, dL[3][3] = {{0}};
  for (int d = 0; d < dim; d++) lagrange1d(order, xi[d], L[d], dL[d]);
  for (int i = 0; i < et; i++) {
    double l0 = L[0][lat[i][0]], l1 = L[1][lat[i][1]], l2 = dim == 3 ? L[2][lat[i][2]] : 1.0;
    phi[i] = l0 * l1 * l2;
    dphi[i][0] = dL[0][lat[i][0]] * l1 * l2;
    dphi[i][1] = l0 * dL[1][lat[i][1]] * l2;
    dphi[i][2] = dim == 3 ? l0 * l1 * dL[2][lat[i][2]] : 0.0;
  }
}

static void gauss_point(int et, int ip, double xi[3], double *wt) {
  /* find_stu / Gq_weight: s fastest, then t, then u; positive abscissa first */
  static const double F1 = 0.57735026918962584208, F2 = 0.77459666924148340428;
  static const double W1 = 0.55555555555555555556, W2 = 0.88888888888888888888;
  int dim = (et == 4 || et == 9) ? 2 : 3, n1 = (et == 4 || et == 8) ? 2 : 3;
  int idx[3] = {ip % n1, (ip / n1) % n1, ip / (n1 * n1)};
  double w = 1.0;
  xi[2] = 0.0;
  for (int d = 0; d < dim; d++) {
    if (n1 == 2) {
      xi[d] = idx[d] == 0 ? F1 : -F1;
    } else {
      xi[d] = idx[d] == 0 ? F2 : idx[d] == 1 ? 0.0 : -F2;
      w *= idx[d] == 1 ? W2 : W1;
    }
  }
  *wt = w;
}

/* in_list over one MSR row, as load_lec does (mm_fill.c:5461) */
static long msr_pos(const int *ija, int ie, int je) {
  if (ie == je) return ie;
  for (int k = ija[ie]; k < ija[ie + 1]; k++)
    if (ija[k] == je) return k;
  return -1;
}

/* Returns 0, or -2 if a (row,col) the element block touches is missing from ija.  `only_mat` >= 0 restricts the loop
 * to the elements of that material (p then carries its constants). */
static int port_fill_elems(const struct goma_gpu_problem *p, int only_mat, const int *elem_mat, const int *ija, const double *x,
                           const double *x_old, const double *xdot, double delta_t, double theta, double time_value,
                           double h_elem_avg, double U_norm, int assemble_residual, int assemble_jacobian, double *a,
                           double *resid) {
  (void)x_old; (void)time_value;
  const int dim = p->dim, et = p->elem_type, nn = et;
  const int n1 = (et == 4 || et == 8) ? 2 : 3;
  const int ngp = dim == 2 ? n1 * n1 : n1 * n1 * n1;
  const int p1 = p->pressure_interp == GOMA_PRESSURE_P1;
  const int cen = et == 9 ? 8 : et == 27 ? 20 : 0;
  const int np = p1 ? dim + 1 : 0;
  /* fields interpolated with phi, in nodal order */
  int fslot[MAXF], nf = 0, fT = -1, fP = -1, fY = -1, fD = -1;
  for (int d = 0; d < dim; d++) fslot[nf++] = GOMA_SLOT_U + d;
  if (p->energy) { fT = nf; fslot[nf++] = GOMA_SLOT_T; }
  if (p->num_species) fY = nf;
  for (int w = 0; w < p->num_species; w++) fslot[nf++] = GOMA_SLOT_Y0 + w;
  if (p->ale) { fD = nf; for (int d = 0; d < dim; d++) fslot[nf++] = GOMA_SLOT_DX + d; }
  if (!p1) { fP = nf; fslot[nf++] = GOMA_SLOT_P; }
  const int ns = p->num_species;
  const int ndof = nf * nn + np; /* element block size; P1 dofs last */
  const int transient = p->transient;
  double em[6], ee[5], es[5];
  memcpy(em, p->etm_momentum, sizeof(em));
  memcpy(ee, p->etm_energy, sizeof(ee));
  memcpy(es, p->etm_species, sizeof(es));
  if (!transient) em[0] = ee[0] = es[0] = 0.0;
  const double ed3 = p->etm_mesh[3];
  const double lam = p->lame_lambda, mus = p->lame_mu;
  const int nonlinear_mesh = 1; /* Solid Constitutive Equation = NONLINEAR (the deck this repo writes) */
  const double ec0 = p->etm_continuity[0];
  const double tfac = transient ? (1.0 + 2.0 * theta) / delta_t : 0.0;
  const double rho = p->rho, mu = p->mu, rcp = p->rho * p->heat_capacity, kc = p->conductivity;

  double *R = (double *)malloc(sizeof(double) * ndof);
  double *J = (double *)malloc(sizeof(double) * ndof * ndof);
  int *gun = (int *)malloc(sizeof(int) * ndof);
  int *lnode = (int *)malloc(sizeof(int) * ndof);
  int rc = 0;

  for (int e = 0; e < p->num_elems; e++) {
    if (only_mat >= 0 && elem_mat[e] != only_mat) continue;
    const int *c = p->elem_connect + (size_t)e * nn;
    double X[3][MAXN], U[MAXF][MAXN], Ud[MAXF][MAXN], Pd[4] = {0, 0, 0, 0};
    /* local dof numbering: dof(f,i) = f*nn + i ; P1 dofs = nf*nn + q */
    for (int i = 0; i < nn; i++) {
      int kd = p->node_kind[c[i]];
      for (int d = 0; d < dim; d++) X[d][i] = p->coord[d][c[i]];
      for (int f = 0; f < nf; f++) {
        int g = p->first_unknown[c[i]] + p->kind_slot[kd][fslot[f]];
        gun[f * nn + i] = g;
        lnode[f * nn + i] = i;
        U[f][i] = x[g];
        Ud[f][i] = transient ? xdot[g] : 0.0;
      }
    }
    for (int q = 0; q < np; q++) {
      int g = p->first_unknown[c[cen]] + p->kind_slot[p->node_kind[c[cen]]][GOMA_SLOT_P] + q;
      gun[nf * nn + q] = g;
      lnode[nf * nn + q] = cen;
      Pd[q] = x[g];
    }
    memset(R, 0, sizeof(double) * ndof);
    memset(J, 0, sizeof(double) * ndof * ndof);
    if (p->ale) /* the map uses the displaced coordinates x = X + d (beer_belly, mm_fill_util.c:258-276) */
      for (int d = 0; d < dim; d++)
        for (int i = 0; i < nn; i++) X[d][i] += U[fD + d][i];

    /* BLOCK 1.5 (mm_fill.c:754-787): element-level PSPG data */
    double tau = 0.0, tau1 = 0.0, hh_siz = 0.0, v_avg[3] = {0, 0, 0}, dtau_dv[3] = {0, 0, 0};
    if (p->pspg) {
      const double mu_avg = mu, rho_avg = rho;
      if (p->pspg == 1) { /* global: Re from the global norms, no Jacobian dependence */
        double Re = rho * U_norm * h_elem_avg / (2.0 * mu_avg);
        tau = Re <= 3.0 ? p->ps_scaling * h_elem_avg * h_elem_avg / (12.0 * mu_avg)
                        : p->ps_scaling * h_elem_avg / (2.0 * rho * U_norm);
      } else { /* local: h_elem_siz (face-centroid differences) and element_velocity */
        double hsq[3] = {0, 0, 0};
        if (dim == 2) {
          for (int a_ = 0; a_ < 2; a_++) {
            double h0 = 0.5 * (X[a_][1] + X[a_][2]) - 0.5 * (X[a_][0] + X[a_][3]);
            double h1 = 0.5 * (X[a_][0] + X[a_][1]) - 0.5 * (X[a_][2] + X[a_][3]);
            hsq[0] += h0 * h0; hsq[1] += h1 * h1;
          }
        } else {
          for (int a_ = 0; a_ < 3; a_++) {
            const double *xx = X[a_];
            double p1 = 0.25 * (xx[0] + xx[1] + xx[2] + xx[3]), p2 = 0.25 * (xx[1] + xx[2] + xx[5] + xx[6]);
            double p3 = 0.25 * (xx[2] + xx[3] + xx[6] + xx[7]), p4 = 0.25 * (xx[0] + xx[1] + xx[4] + xx[5]);
            double p5 = 0.25 * (xx[0] + xx[3] + xx[4] + xx[7]), p6 = 0.25 * (xx[4] + xx[5] + xx[6] + xx[7]);
            hsq[0] += (p2 - p5) * (p2 - p5); hsq[1] += (p3 - p4) * (p3 - p4); hsq[2] += (p1 - p6) * (p1 - p6);
          }
        }
        for (int a_ = 0; a_ < dim; a_++) hh_siz += hsq[a_];
        hh_siz /= (double)dim;
        double vv = 0.0;
        const int q2 = (et == 9 || et == 27);
        for (int a_ = 0; a_ < dim; a_++) {
          if (q2) v_avg[a_] = U[a_][nn - 1]; /* I_Q2: "centroid_node = dofs - 1" (mm_fill_aux.c:809-817) */
          else for (int k = 0; k < nn; k++) v_avg[a_] += U[a_][k] / (double)nn;
          vv += v_avg[a_] * v_avg[a_];
        }
        tau1 = rho_avg * rho_avg * vv / hh_siz + 9.0 * mu_avg * mu_avg / (hh_siz * hh_siz);
        if (transient) tau1 += 4.0 / (delta_t * delta_t);
        tau = p->ps_scaling / sqrt(tau1);
        for (int b = 0; b < dim; b++) dtau_dv[b] = -tau / tau1 * rho_avg * rho_avg / hh_siz * v_avg[b];
      }
    }

    for (int ip = 0; ip < ngp; ip++) {
      double xi[3], wt, phi[MAXN], dphi[MAXN][3], g[MAXN][3], psi[4];
      gauss_point(et, ip, xi, &wt);
      basis(et, xi, phi, dphi);
      psi[0] = 1.0;
      for (int d = 0; d < dim; d++) psi[1 + d] = xi[d];
      /* beer_belly */
      double Jm[3][3] = {{0}}, B[3][3] = {{0}}, det;
      for (int a_ = 0; a_ < dim; a_++)
        for (int b = 0; b < dim; b++)
          for (int k = 0; k < nn; k++) Jm[a_][b] += X[b][k] * dphi[k][a_];
      if (dim == 2) {
        det = Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0];
        B[0][0] = Jm[1][1] / det; B[0][1] = -Jm[0][1] / det;
        B[1][0] = -Jm[1][0] / det; B[1][1] = Jm[0][0] / det;
      } else {
        det = Jm[0][0] * (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) - Jm[0][1] * (Jm[1][0] * Jm[2][2] - Jm[2][0] * Jm[1][2]) +
              Jm[0][2] * (Jm[1][0] * Jm[2][1] - Jm[2][0] * Jm[1][1]);
        B[0][0] = (Jm[1][1] * Jm[2][2] - Jm[2][1] * Jm[1][2]) / det;
        B[0][1] = -(Jm[0][1] * Jm[2][2] - Jm[2][1] * Jm[0][2]) / det;
        B[0][2] = (Jm[0][1] * Jm[1][2] - Jm[1][1] * Jm[0][2]) / det;
        B[1][0] = -(Jm[1][0] * Jm[2][2] - Jm[2][0] * Jm[1][2]) / det;
        B[1][1] = (Jm[0][0] * Jm[2][2] - Jm[2][0] * Jm[0][2]) / det;
        B[1][2] = -(Jm[0][0] * Jm[1][2] - Jm[1][0] * Jm[0][2]) / det;
        B[2][0] = (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]) / det;
        B[2][1] = -(Jm[0][0] * Jm[2][1] - Jm[2][0] * Jm[0][1]) / det;
        B[2][2] = (Jm[0][0] * Jm[1][1] - Jm[1][0] * Jm[0][1]) / det;
      }
      /* zero_detJ (mm_fill_util.c:335-343) is raised only inside beer_belly's SHELL / TRISHELL branch (:312-344):
       * continuum elements are assembled whatever |detJ| is */
      const double d_area = det * wt; /* h3 = 1 */
      for (int i = 0; i < nn; i++)
        for (int q = 0; q < dim; q++) {
          g[i][q] = 0.0;
          for (int r = 0; r < dim; r++) g[i][q] += B[q][r] * dphi[i][r];
        }
      /* load_fv / load_fv_grads: grad[f][q] = d f / d x_q */
      double val[MAXF], dot[MAXF], grad[MAXF][3];
      for (int f = 0; f < nf; f++) {
        val[f] = dot[f] = 0.0;
        grad[f][0] = grad[f][1] = grad[f][2] = 0.0;
        for (int k = 0; k < nn; k++) {
          val[f] += U[f][k] * phi[k];
          dot[f] += Ud[f][k] * phi[k];
          for (int q = 0; q < dim; q++) grad[f][q] += U[f][k] * g[k][q];
        }
      }
      double Pr = 0.0;
      if (p1) for (int q = 0; q < np; q++) Pr += Pd[q] * psi[q];
      else Pr = val[fP];
      const double T = p->energy ? val[fT] : 0.0;
      /* momentum_source_term */
      double fs[3] = {0, 0, 0}, dfdT[3] = {0, 0, 0};
      if (em[4] != 0.0) {
        for (int a_ = 0; a_ < dim; a_++) {
          if (p->momentum_source_model == 0) {
            fs[a_] = p->momentum_source[a_];
          } else if (p->energy) {
            double d = -p->volume_expansion * (T - p->reference_temperature);
            fs[a_] = rho * p->momentum_source[a_] * (p->momentum_source_model == 1 ? 1.0 + d : d);
            dfdT[a_] = -p->momentum_source[a_] * rho * p->volume_expansion;
          }
        }
      }
      double div_v = 0.0;
      for (int q = 0; q < dim; q++) div_v += grad[q][q];
      /* convection velocity v - xdot_mesh (get_convection_velocity, mm_fill_species.c:9479-9492;
       * assemble_momentum's x_dot, mm_fill_momentum.c:414-416) */
      double vcv[3] = {0, 0, 0};
      for (int q = 0; q < dim; q++) vcv[q] = val[q] - (p->ale && transient ? dot[fD + q] : 0.0);

      for (int i = 0; i < nn; i++) {
        const double phi_i = phi[i];
        /* ---- momentum rows */
        for (int a_ = 0; a_ < dim; a_++) {
          const int row = a_ * nn + i;
          if (assemble_residual) {
            double adv = 0.0, diff = 0.0;
            for (int q = 0; q < dim; q++) adv += vcv[q] * grad[a_][q]; /* (v - xdot)_q d_q v_a */
            for (int q = 0; q < dim; q++) {
              double Pi = mu * (grad[q][a_] + grad[a_][q]) - (q == a_ ? Pr : 0.0);
              diff += g[i][q] * Pi;
            }
            R[row] += -em[0] * rho * phi_i * dot[a_] * d_area - em[1] * rho * phi_i * adv * d_area -
                      em[3] * diff * d_area + em[4] * phi_i * fs[a_] * d_area;
          }
          if (assemble_jacobian) {
            for (int j = 0; j < nn; j++) {
              double gij = 0.0, vgj = 0.0;
              for (int q = 0; q < dim; q++) { gij += g[i][q] * g[j][q]; vgj += vcv[q] * g[j][q]; }
              for (int b = 0; b < dim; b++) {
                double mass = (a_ == b) ? -em[0] * rho * phi_i * phi[j] * tfac * d_area : 0.0;
                double adv = -em[1] * rho * phi_i * (phi[j] * grad[a_][b] + (a_ == b ? vgj : 0.0)) * d_area;
                double dif = -em[3] * mu * (g[i][b] * g[j][a_] + (a_ == b ? gij : 0.0)) * d_area;
                J[row * ndof + b * nn + j] += mass + adv + dif;
              }
              if (p->energy) J[row * ndof + fT * nn + j] += em[4] * phi_i * dfdT[a_] * phi[j] * d_area;
            }
            for (int q = 0; q < np; q++) J[row * ndof + nf * nn + q] += em[3] * g[i][a_] * psi[q] * d_area;
            if (!p1) for (int j = 0; j < nn; j++) J[row * ndof + fP * nn + j] += em[3] * g[i][a_] * phi[j] * d_area;
          }
        }
        /* ---- species rows (Fickian, constant D, concentration form) */
        for (int w = 0; w < ns; w++) {
          const int fw = fY + w, row = fw * nn + i;
          const double D = p->diffusivity[w];
          if (assemble_residual) {
            double adv = 0.0, diff = 0.0;
            for (int q = 0; q < dim; q++) { adv += vcv[q] * grad[fw][q]; diff += g[i][q] * (-D * grad[fw][q]); }
            R[row] += -es[0] * phi_i * dot[fw] * d_area - es[1] * phi_i * adv * d_area + es[3] * diff * d_area;
          }
          if (assemble_jacobian) {
            for (int j = 0; j < nn; j++) {
              double gij = 0.0, vgj = 0.0;
              for (int q = 0; q < dim; q++) { gij += g[i][q] * g[j][q]; vgj += vcv[q] * g[j][q]; }
              J[row * ndof + fw * nn + j] += (-es[0] * phi_i * phi[j] * tfac - es[1] * phi_i * vgj - es[3] * D * gij) * d_area;
              for (int b = 0; b < dim; b++) J[row * ndof + b * nn + j] += -es[1] * phi_i * phi[j] * grad[fw][b] * d_area;
            }
          }
        }
        /* ---- continuity rows, equal-order pressure (+ PSPG) */
        if (!p1) {
          const int row = fP * nn + i;
          double mom[3] = {0, 0, 0}, pspg[3] = {0, 0, 0};
          if (p->pspg) {
            for (int a_ = 0; a_ < dim; a_++) {
              double adv = 0.0;
              for (int q = 0; q < dim; q++) adv += val[q] * grad[a_][q];
              mom[a_] = em[0] * rho * dot[a_] + em[1] * rho * adv + em[3] * grad[fP][a_] - em[4] * fs[a_];
              pspg[a_] = tau * mom[a_];
            }
          }
          if (assemble_residual) {
            double ps = 0.0;
            for (int a_ = 0; a_ < dim; a_++) ps += g[i][a_] * pspg[a_];
            R[row] += ec0 * phi_i * div_v * d_area + ps * d_area;
          }
          if (assemble_jacobian) {
            for (int j = 0; j < nn; j++) {
              double vgj = 0.0;
              for (int q = 0; q < dim; q++) vgj += val[q] * g[j][q];
              for (int b = 0; b < dim; b++) {
                double ps = 0.0;
                if (p->pspg)
                  for (int a_ = 0; a_ < dim; a_++) {
                    double d = tau * (em[0] * rho * tfac * phi[j] * (a_ == b) +
                                      em[1] * rho * (phi[j] * grad[a_][b] + (a_ == b ? vgj : 0.0))) +
                               dtau_dv[b] * (et == 9 || et == 27 ? (j == nn - 1 ? 1.0 : 0.0) : 1.0 / nn) * mom[a_];
                    ps += g[i][a_] * d;
                  }
                J[row * ndof + b * nn + j] += ec0 * phi_i * g[j][b] * d_area + ps * d_area;
              }
              if (p->pspg) {
                double psP = 0.0, psT = 0.0;
                for (int a_ = 0; a_ < dim; a_++) {
                  psP += g[i][a_] * tau * em[3] * g[j][a_];
                  psT += g[i][a_] * tau * (-em[4] * dfdT[a_] * phi[j]);
                }
                J[row * ndof + fP * nn + j] += psP * d_area;
                if (p->energy) J[row * ndof + fT * nn + j] += psT * d_area;
              }
            }
          }
        }
        /* ---- energy row */
        if (p->energy) {
          const int row = fT * nn + i;
          if (assemble_residual) {
            double adv = 0.0, diff = 0.0;
            for (int q = 0; q < dim; q++) { adv += vcv[q] * grad[fT][q]; diff += g[i][q] * (-kc * grad[fT][q]); }
            R[row] += -ee[0] * rcp * phi_i * dot[fT] * d_area - ee[1] * rcp * phi_i * adv * d_area +
                      ee[3] * diff * d_area + ee[4] * phi_i * p->heat_source * d_area;
          }
          if (assemble_jacobian) {
            for (int j = 0; j < nn; j++) {
              double gij = 0.0, vgj = 0.0;
              for (int q = 0; q < dim; q++) { gij += g[i][q] * g[j][q]; vgj += vcv[q] * g[j][q]; }
              J[row * ndof + fT * nn + j] += (-ee[0] * rcp * phi_i * phi[j] * tfac - ee[1] * rcp * phi_i * vgj -
                                              ee[3] * kc * gij) * d_area;
              for (int b = 0; b < dim; b++)
                J[row * ndof + b * nn + j] += -ee[1] * rcp * phi_i * phi[j] * grad[fT][b] * d_area;
            }
          }
        }
      }
      /* ---- pseudo-solid mesh equations and the mesh sensitivities of the fluid equations */
      if (p->ale) {
        double G[3][3] = {{0}}, E[3][3], F[3][3], TT[3][3], vc, vs;
        for (int pp_ = 0; pp_ < dim; pp_++)
          for (int q = 0; q < dim; q++) G[pp_][q] = grad[fD + q][pp_]; /* grad_d[p][q] = d_p d_q */
        for (int pp_ = 0; pp_ < dim; pp_++)
          for (int q = 0; q < dim; q++) {
            E[pp_][q] = 0.5 * (G[pp_][q] + G[q][pp_]);
            if (nonlinear_mesh)
              for (int a_ = 0; a_ < dim; a_++) E[pp_][q] -= 0.5 * G[pp_][a_] * G[q][a_];
            F[pp_][q] = (pp_ == q ? 1.0 : 0.0) - G[pp_][q];
          }
        double cof[3][3] = {{0}}; /* d(det F)/dF[p][q] */
        double detF;
        if (dim == 2) {
          detF = F[0][0] * F[1][1] - F[0][1] * F[1][0];
          cof[0][0] = F[1][1]; cof[0][1] = -F[1][0]; cof[1][0] = -F[0][1]; cof[1][1] = F[0][0];
        } else {
          cof[0][0] = F[1][1] * F[2][2] - F[1][2] * F[2][1]; cof[0][1] = F[1][2] * F[2][0] - F[1][0] * F[2][2];
          cof[0][2] = F[1][0] * F[2][1] - F[1][1] * F[2][0]; cof[1][0] = F[0][2] * F[2][1] - F[0][1] * F[2][2];
          cof[1][1] = F[0][0] * F[2][2] - F[0][2] * F[2][0]; cof[1][2] = F[0][1] * F[2][0] - F[0][0] * F[2][1];
          cof[2][0] = F[0][1] * F[1][2] - F[0][2] * F[1][1]; cof[2][1] = F[0][2] * F[1][0] - F[0][0] * F[1][2];
          cof[2][2] = F[0][0] * F[1][1] - F[0][1] * F[1][0];
          detF = F[0][0] * cof[0][0] + F[0][1] * cof[0][1] + F[0][2] * cof[0][2];
        }
        if (detF <= 0.0) { rc = -1; goto done; } /* neg_elem_volume (mm_fill_solid.c:659-663, :811-815) */
        vc = 1.0 / detF;
        vs = 3.0 * (pow(vc, 1.0 / 3.0) - 1.0);
        for (int pp_ = 0; pp_ < dim; pp_++)
          for (int q = 0; q < dim; q++) TT[pp_][q] = lam * vs * (pp_ == q) + 2.0 * mus * E[pp_][q];
        /* fluid stress at this point, for J_m_d */
        double Pi[3][3], advv[3];
        for (int a_ = 0; a_ < dim; a_++) {
          advv[a_] = 0.0;
          for (int q = 0; q < dim; q++) {
            advv[a_] += vcv[q] * grad[a_][q];
            Pi[a_][q] = mu * (grad[q][a_] + grad[a_][q]) - (q == a_ ? Pr : 0.0);
          }
        }
        for (int i = 0; i < nn; i++) {
          double giTT[3], giPi[3];
          for (int a_ = 0; a_ < dim; a_++) {
            giTT[a_] = giPi[a_] = 0.0;
            for (int q = 0; q < dim; q++) { giTT[a_] += g[i][q] * TT[a_][q]; giPi[a_] += g[i][q] * Pi[a_][q]; }
            if (assemble_residual) R[(fD + a_) * nn + i] += -ed3 * giTT[a_] * d_area;
          }
          if (!assemble_jacobian) continue;
          for (int j = 0; j < nn; j++) {
            double vgj = 0.0;
            for (int q = 0; q < dim; q++) vgj += vcv[q] * g[j][q];
            for (int b = 0; b < dim; b++) {
              /* d grad_d[p][q] / d d_bj = grad_phi_j[p] (delta_qb - grad_d[b][q]) */
              double dG[3][3], dE[3][3], ddet = 0.0;
              for (int pp_ = 0; pp_ < dim; pp_++)
                for (int q = 0; q < dim; q++) dG[pp_][q] = g[j][pp_] * ((q == b ? 1.0 : 0.0) - G[b][q]);
              for (int pp_ = 0; pp_ < dim; pp_++)
                for (int q = 0; q < dim; q++) {
                  dE[pp_][q] = 0.5 * (dG[pp_][q] + dG[q][pp_]);
                  if (nonlinear_mesh)
                    for (int a_ = 0; a_ < dim; a_++) dE[pp_][q] -= 0.5 * (dG[pp_][a_] * G[q][a_] + G[pp_][a_] * dG[q][a_]);
                  ddet += cof[pp_][q] * (-dG[pp_][q]); /* d det(F), F = I - grad_d */
                }
              const double dvc = -ddet * vc * vc, dvs = dvc * pow(vc, -2.0 / 3.0);
              for (int a_ = 0; a_ < dim; a_++) {
                /* J_d_d = diff_a + diff_b + diff_c (mm_fill_terms.c:529-576) */
                double gjTT = 0.0, gidTT = 0.0;
                for (int q = 0; q < dim; q++) {
                  gjTT += g[j][q] * TT[a_][q];
                  gidTT += g[i][q] * (lam * dvs * (a_ == q) + 2.0 * mus * dE[a_][q]);
                }
                J[((fD + a_) * nn + i) * ndof + (fD + b) * nn + j] +=
                    ed3 * (g[i][b] * gjTT - gidTT - g[j][b] * giTT[a_]) * d_area;
                /* J_m_d (mm_fill_momentum.c:2200-2442), steady */
                double gjPi = 0.0, gidPi = 0.0;
                for (int q = 0; q < dim; q++) {
                  gjPi += g[j][q] * Pi[a_][q];
                  gidPi += g[i][q] * mu * (-g[j][a_] * grad[q][b] - g[j][q] * grad[a_][b]);
                }
                J[(a_ * nn + i) * ndof + (fD + b) * nn + j] +=
                    (em[1] * rho * phi[i] * (vgj * grad[a_][b] - advv[a_] * g[j][b]) +
                     em[3] * (g[i][b] * gjPi - gidPi - g[j][b] * giPi[a_]) + em[4] * phi[i] * fs[a_] * g[j][b]) * d_area;
                /* transient: mass x d|J| (:2213-2220) and advection_c = d(v - xdot)/d d_bj . grad v, which the
                 * reference adds only when the mass term is on (:2300-2318) */
                if (em[0] != 0.0)
                  J[(a_ * nn + i) * ndof + (fD + b) * nn + j] +=
                      (-em[0] * rho * phi[i] * dot[a_] * g[j][b] + em[1] * rho * phi[i] * tfac * phi[j] * grad[a_][b]) * d_area;
              }
              /* J_e_d (mm_fill_energy.c:758-925) and J_s_d (mm_fill_species.c:1103-1330): same three-part pattern */
              for (int s_ = 0; s_ < (p->energy ? 1 : 0) + ns; s_++) {
                const int fs_ = (p->energy && s_ == 0) ? fT : fY + s_ - (p->energy ? 1 : 0);
                const int isT = (fs_ == fT);
                const double cm = isT ? ee[0] * rcp : es[0], ca = isT ? ee[1] * rcp : es[1];
                const double cd = isT ? ee[3] * kc : es[3] * p->diffusivity[fs_ - fY];
                const double src = isT ? ee[4] * p->heat_source : 0.0;
                double vgs = 0.0, gjgs = 0.0, gigs = 0.0, gij = 0.0;
                for (int q = 0; q < dim; q++) {
                  vgs += vcv[q] * grad[fs_][q]; gjgs += g[j][q] * grad[fs_][q];
                  gigs += g[i][q] * grad[fs_][q]; gij += g[i][q] * g[j][q];
                }
                double v = ca * phi[i] * (vgj * grad[fs_][b] - vgs * g[j][b]) +
                           cd * (gjgs * g[i][b] + gij * grad[fs_][b] - gigs * g[j][b]) + src * phi[i] * g[j][b];
                if (cm != 0.0) v += -cm * phi[i] * dot[fs_] * g[j][b] + ca * phi[i] * tfac * phi[j] * grad[fs_][b];
                J[(fs_ * nn + i) * ndof + (fD + b) * nn + j] += v * d_area;
              }
              /* J_c_d (mm_fill_continuity.c:1004-1148): d(div v) + div v d|J| */
              double ddiv = 0.0;
              for (int q = 0; q < dim; q++) ddiv -= g[j][q] * grad[q][b];
              if (i == 0)
                for (int q = 0; q < np; q++)
                  J[(nf * nn + q) * ndof + (fD + b) * nn + j] += ec0 * psi[q] * (ddiv + div_v * g[j][b]) * d_area;
            }
          }
        }
      }
      /* ---- continuity rows (P1) */
      for (int q = 0; q < np; q++) {
        const int row = nf * nn + q;
        if (assemble_residual) R[row] += ec0 * psi[q] * div_v * d_area;
        if (assemble_jacobian)
          for (int j = 0; j < nn; j++)
            for (int b = 0; b < dim; b++) J[row * ndof + b * nn + j] += ec0 * psi[q] * g[j][b] * d_area;
      }
    }

    /* put_dirichlet_in_matrix */
    for (int r = 0; r < ndof; r++) {
      int fl = p->dbc_flag[gun[r]];
      if (!fl) continue;
      for (int cidx = 0; cidx < ndof; cidx++) J[r * ndof + cidx] = 0.0;
      J[r * ndof + r] = 1.0;
      R[r] = fl == 1 ? x[gun[r]] - p->dbc_value[gun[r]] : 0.0;
    }
    /* load_lec, MSR */
    for (int r = 0; r < ndof; r++) {
      if (c[lnode[r]] >= p->num_owned_nodes) continue;
      int ie = gun[r];
      if (assemble_residual) resid[ie] += R[r];
      if (!assemble_jacobian) continue;
      int rowT = p->energy && r >= fT * nn && r < (fT + 1) * nn;
      for (int cidx = 0; cidx < ndof; cidx++) {
        if (rowT && cidx >= nf * nn) continue;           /* Inter_Mask[T][P] = 0 */
        if (rowT && !p1 && cidx / nn == fP) continue;
        long pos = msr_pos(ija, ie, gun[cidx]);
        if (pos < 0) { rc = -2; continue; }
        a[pos] += J[r * ndof + cidx];
      }
    }
  }
done:
  free(R); free(J); free(gun); free(lnode);
  return rc;
}

/* The reference's element loop takes mp = mp_glob[Matilda[ebn]] per element block (mm_fill.c:224-235, 621-640).  Element
 * blocks hold consecutive elements, so sweeping material by material adds the element contributions in the same order. */
int goma_port_fill(const struct goma_gpu_problem *p, const int *ija, const double *x, const double *x_old,
                   const double *xdot, double delta_t, double theta, double time_value, double h_elem_avg,
                   double U_norm, int assemble_residual, int assemble_jacobian, double *a, double *resid) {
  if (p->num_materials <= 1 || !p->materials || !p->elem_material)
    return port_fill_elems(p, -1, NULL, ija, x, x_old, xdot, delta_t, theta, time_value, h_elem_avg, U_norm,
                           assemble_residual, assemble_jacobian, a, resid);
  for (int m = 0; m < p->num_materials; m++) {
    struct goma_gpu_problem pm = *p;
    const struct goma_gpu_material *M = &p->materials[m];
    pm.rho = M->rho; pm.mu = M->mu; pm.conductivity = M->conductivity; pm.heat_capacity = M->heat_capacity;
    pm.volume_expansion = M->volume_expansion; pm.reference_temperature = M->reference_temperature;
    memcpy(pm.diffusivity, M->diffusivity, sizeof(pm.diffusivity));
    memcpy(pm.momentum_source, M->momentum_source, sizeof(pm.momentum_source));
    pm.momentum_source_model = M->momentum_source_model;
    pm.heat_source = M->heat_source; pm.lame_mu = M->lame_mu; pm.lame_lambda = M->lame_lambda;
    int rc = port_fill_elems(&pm, m, p->elem_material, ija, x, x_old, xdot, delta_t, theta, time_value, h_elem_avg, U_norm,
                             assemble_residual, assemble_jacobian, a, resid);
    if (rc) return rc;
  }
  return 0;
}
