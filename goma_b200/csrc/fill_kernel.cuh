// Element assembly kernel for sm_100a: persistent, warp-specialised CTAs.  Producer warps build the
// per-element operand tables (geometry, basis gradients, field values, residual) for element n+1
// into one half of a double-buffered shared-memory arena while consumer warps accumulate and scatter
// the Jacobian blocks of element n from the other half.
//
// What it replaces in the reference, per element (src/mm_fill.c:317 matrix_fill):
//   BLOCK 1   load_elem_dofptr / load_ei          -> phase 0 gather through prebuilt tables
//   per Gauss point (mm_fill.c:1253-2665):
//     load_basis_functions (mm_fill_util.c:2608)  -> constant tables staged once per CTA by TMA
//     beer_belly           (mm_fill_util.c:140)   -> phases 1-2 (J, detJ, B = J^-1 by cofactors)
//     load_bf_grad         (mm_fill_util.c:1634)  -> phase 3  (grad_phi = B . dphi/dxi)
//     load_fv / load_fv_grads (load_field_variables.c:128,2049) -> phase 4
//     assemble_momentum    (mm_fill_momentum.c:100)   residual :534-662, J_m_v :1564-1735,
//                                                     J_m_T :746-915, J_m_P :2052-2117
//     assemble_continuity  (mm_fill_continuity.c:119) residual :435-613, J_c_v :665-761
//     assemble_energy      (mm_fill_energy.c:109)     residual :322-381, J_e_T :425-487, J_e_v :628-692
//   BLOCK 8   put_dirichlet_in_matrix (bc_dirich.c:44)
//   load_lec  (mm_fill.c:5175, MSR branch :5241-5483) -> slot-mapped fp64 atomic scatter
//
// The Jacobian uses the Cartesian closed forms of SURVEY.md App. A instead of the reference's
// zero-padded grad_phi_e / d_Pi tensors; the per-(i,j) node-pair block is accumulated in
// registers over the Gauss points and written once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/goma_gpu_fill.h"

namespace goma_b200 {

struct FillParams {
  // mesh / maps
  const int *conn;
  const double *coord[3];
  const int *first_unknown;
  const unsigned char *node_kind;
  int kind_slot[GOMA_GPU_MAX_KINDS][GOMA_NSLOT];
  const long long *rowstart;
  const unsigned short *pair_full;
  const unsigned short *pair_p;
  const unsigned char *dbc_flag;
  const double *dbc_value;
  const int *elem_list;  // optional indirection (colour classes); nullptr = identity
  int elem_begin, elem_end;
  int num_owned_nodes;
  // state
  const double *x, *x_old, *xdot;
  double *a;
  double *resid;
  int *flags;
  const double *tables;
  // switches
  int assemble_residual, assemble_jacobian, transient, use_atomics;
  // constants
  double etm_mom[6], etm_cont[2], etm_energy[5], etm_species[5], etm_mesh[5];
  double rho, mu, k, Cp, beta, Tref, heat_source;
  double g[3];
  int source_model;  // 0 CONSTANT, 1 BOUSS (hydrostatic), 2 BOUSSINESQ
  double diffusivity[4];
  double delta_t, theta, time_value, h_elem_avg, U_norm;
  double lame_mu, lame_lambda;
  long long *prof;  // optional per-CTA phase cycle counters (debug/profiling), 8 per CTA
  int debug;        // experiments only: bit0 = drop the matrix stores, bit1 = skip the Gauss loop
};

template <int DIM_, int NN_, int NGP_, bool P1_, bool ENERGY_, int NSPEC_, bool ALE_, int NCT_, int NPT_, int TI_,
          bool SPEC_ = true>
struct Cfg {
  static constexpr int DIM = DIM_, NN = NN_, NGP = NGP_, NSPEC = NSPEC_, TI = TI_;
  // SPEC: warp-specialised (NPT producer + NCT consumer threads, double-buffered element data);
  // !SPEC: every thread plays both roles in turn (NCT == NPT == CTA size), several CTAs per SM
  static constexpr bool SPEC = SPEC_;
  static constexpr int NCT = NCT_;        // consumer threads (Jacobian tiles + scatter)
  static constexpr int NPT = NPT_;        // producer threads (everything else)
  static constexpr int TPE = SPEC ? NCT + NPT : NCT;   // CTA size
  static constexpr int NBUF = SPEC ? 2 : 1;
  static constexpr int MINB = SPEC ? 1 : 2;
  static_assert(SPEC || NCT == NPT, "role-less variant: all threads do everything");
  static constexpr bool P1 = P1_, ENERGY = ENERGY_, ALE = ALE_;
  static constexpr int F_V = 0;
  static constexpr int F_T = DIM;
  static constexpr int F_Y = DIM + (ENERGY ? 1 : 0);
  static constexpr int F_D = F_Y + NSPEC;
  static constexpr int F_P = F_D + (ALE ? DIM : 0);
  static constexpr int NF = F_P + (P1 ? 0 : 1);
  static constexpr int NP = P1 ? DIM + 1 : 0;
  static constexpr int CEN = NN == 9 ? 8 : (NN == 27 ? 20 : 0);
  static constexpr int NTILE = (NN / TI) * NN;  // register tiles (TI rows x 1 column of node pairs) per element
  static_assert(NN % TI == 0, "row tile must divide the node count");
  static_assert(NCT % 32 == 0 && NPT % 32 == 0, "warp-granular roles");
  static constexpr int TBL = NGP + NGP * NN + NGP * NN * DIM + NGP * (DIM + 1);
  static constexpr int TBL_PAD = (TBL + 1) & ~1;  // 16-byte multiple for the bulk copy
  static constexpr int T_WT = 0, T_PHI = NGP, T_DPHI = T_PHI + NGP * NN, T_PSI = T_DPHI + NGP * NN * DIM;
  // per-Gauss-point derived data (doubles): see phase 4b
  static constexpr int G_GV = 0;                                   // c_adv d_b v_a            [a][b]
  static constexpr int G_GT = G_GV + DIM * DIM;                    // ce_adv d_b T             [b]
  static constexpr int G_RQ = G_GT + (ENERGY ? DIM : 0);           // momentum residual, multiplies w phi_i   [a]
  static constexpr int G_RP = G_RQ + DIM;                          // -e3 Pi[a][p], multiplies w grad_phi_i[p]
  static constexpr int G_RE = G_RP + DIM * DIM;                    // energy residual scalar
  static constexpr int G_RF = G_RE + (ENERGY ? 1 : 0);             // e3 q[p]
  static constexpr int G_DIV = G_RF + (ENERGY ? DIM : 0);          // ec0 div v
  static constexpr int GPD = (G_DIV + 1 + 1) & ~1;
  __host__ __device__ static constexpr int slot(int f) {
    return f < DIM                       ? GOMA_SLOT_U + f
           : (ENERGY && f == F_T)        ? GOMA_SLOT_T
           : (f >= F_Y && f < F_Y + NSPEC) ? GOMA_SLOT_Y0 + (f - F_Y)
           : (ALE && f >= F_D && f < F_D + DIM) ? GOMA_SLOT_DX + (f - F_D)
                                         : GOMA_SLOT_P;
  }
};

// Everything the consumers need about one element; double-buffered.
template <class C>
struct alignas(16) ElemBuf {
  double SI[C::NGP][C::NN][4];  // (w phi_i, w grad_phi_i[p])   test-function side (read as broadcast)
  // trial-function side (phi_j, grad_phi_j[p]), split in two 16-byte-stride arrays so that the
  // consumers' per-lane LDS.128 are bank-conflict free
  double2 SJa[C::NGP][C::NN];   // (phi_j, g_j[0])
  double2 SJb[C::NGP][C::NN];   // (g_j[1], g_j[2])
  double VG[C::NGP][C::NN];     // v . grad_phi_j
  double GP[C::NGP][C::GPD];    // per-Gauss-point derived quantities
  long long rs[C::NF][C::NN];   // MSR row start of (field,node); -1 = not written here (ghost/Dirichlet)
  long long rsP[C::NP > 0 ? C::NP : 1];
  int node[C::NN];
  int fu[C::NN];
  int kind[C::NN];
  int gun[C::NF][C::NN];
  int elem;
  int pad_;
};

template <class C>
struct alignas(16) Smem {
  double tbl[C::TBL_PAD];
  ElemBuf<C> eb[C::NBUF];
  // producer scratch
  double F[C::NGP][C::NF][C::DIM + 2];  // value, grad[DIM], time derivative
  double X[C::DIM][C::NN];
  double U[C::NF][C::NN];
  double Udot[C::NF][C::NN];
  double Pd[C::NP > 0 ? C::NP : 1];
  double w[C::NGP];
  double B[C::NGP][C::DIM * C::DIM];
  double Pgp[C::NGP];
  unsigned long long mbar;
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(phase)
      : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(src), "r"(bytes), "r"(b)
               : "memory");
}
// named barriers (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// per-role cycle counters: compiled in only with -DGOMA_PROFILE_PHASES (they cost registers)
#ifdef GOMA_PROFILE_PHASES
#define GOMA_CLOCK() clock64()
#else
#define GOMA_CLOCK() 0LL
#endif
enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_EMPTY0 = 3, BAR_EMPTY1 = 4, BAR_PROD = 5 };

__device__ __forceinline__ void acc_add(const FillParams &P, double *addr, double val) {
  if (P.debug & 1) return;
  if (P.use_atomics)
    atomicAdd(addr, val);  // result unused -> RED.E.ADD.F64
  else
    *addr += val;
}

// One matrix entry through the slot map: row (MSR row start `rowstart`, global id `row`) of local node
// li, column `col_off` inside local node lj.  `po` = column offset of node lj's first unknown in a row
// of node li (already net of masked-out pressure columns for energy rows).
template <class C>
__device__ __forceinline__ void mat_add(const FillParams &P, const ElemBuf<C> &e, long long rowstart, int row, int lj,
                                        int po, int col_off, double val) {
  const int col = e.fu[lj] + col_off;
  const long long pos = (row == col) ? (long long)row : rowstart + po + col_off - (col > row ? 1 : 0);
  acc_add(P, &P.a[pos], val);
}

// momentum_source_term (mm_fill_momentum.c:3738) CONSTANT branch and bouss_momentum_source
// (mm_std_models.c:125-360), temperature piece
template <class C>
__device__ __forceinline__ void momentum_source(const FillParams &P, double T, double f[3], double dfdT[3]) {
#pragma unroll
  for (int a = 0; a < 3; a++) {
    f[a] = 0.0;
    dfdT[a] = 0.0;
  }
  if (P.etm_mom[4] == 0.0) return;
#pragma unroll
  for (int a = 0; a < C::DIM; a++) {
    if (P.source_model == 0) {
      f[a] = P.g[a];
    } else if (C::ENERGY) {
      double d = -P.beta * (T - P.Tref);
      f[a] = P.rho * P.g[a] * (P.source_model == 1 ? (1.0 + d) : d);
      dfdT[a] = -P.g[a] * P.rho * P.beta;
    }
  }
}

// length-N dot product with three independent accumulation chains (hides DFMA latency in the
// low-parallelism producer phases)
template <int N>
__device__ __forceinline__ double dot3(const double *__restrict__ a, int sa, const double *__restrict__ b, int sb) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  constexpr int M = N / 3;
#pragma unroll
  for (int k = 0; k < M; k++) {
    s0 += a[k * sa] * b[k * sb];
    s1 += a[(k + M) * sa] * b[(k + M) * sb];
    s2 += a[(k + 2 * M) * sa] * b[(k + 2 * M) * sb];
  }
#pragma unroll
  for (int k = 3 * M; k < N; k++) s0 += a[k * sa] * b[k * sb];
  return (s0 + s1) + s2;
}

// =====================================================================================
// producer: phases 0-5 and 7 for one element, into buffer `e`
// =====================================================================================
template <class C>
__device__ __forceinline__ void produce_element(const FillParams &P, Smem<C> &s, ElemBuf<C> &e, int elem, int tid) {
  constexpr int DIM = C::DIM, NN = C::NN, NGP = C::NGP, NF = C::NF, NP = C::NP, NPT = C::NPT;
  const double *t_wt = s.tbl + C::T_WT;
  const double *t_phi = s.tbl + C::T_PHI;    // [gp][NN]
  const double *t_dphi = s.tbl + C::T_DPHI;  // [gp][NN][DIM]
  const double *t_psi = s.tbl + C::T_PSI;    // [gp][DIM+1]
  const double rcp = P.rho * P.Cp;
  const double c_adv = -P.etm_mom[1] * P.rho;
  const double ce_adv = -P.etm_energy[1] * rcp;

  // ---- phase 0: connectivity, coordinates, unknown indices, nodal values (load_elem_dofptr)
  if (tid == 0) e.elem = elem;
  for (int k = tid; k < NN; k += NPT) {
    int nd = P.conn[(size_t)elem * NN + k];
    e.node[k] = nd;
    e.fu[k] = P.first_unknown[nd];
    e.kind[k] = P.node_kind[nd];
#pragma unroll
    for (int d = 0; d < DIM; d++) s.X[d][k] = P.coord[d][nd];
  }
  bar_sync(BAR_PROD, NPT);
  for (int idx = tid; idx < NF * NN; idx += NPT) {
    int f = idx / NN, k = idx - f * NN;
    int gun = e.fu[k] + P.kind_slot[e.kind[k]][C::slot(f)];
    e.gun[f][k] = gun;
    s.U[f][k] = P.x[gun];
    s.Udot[f][k] = P.transient ? P.xdot[gun] : 0.0;
    bool owned = e.node[k] < P.num_owned_nodes;
    e.rs[f][k] = (owned && P.dbc_flag[gun] == 0) ? P.rowstart[gun] : -1;
  }
  if (C::P1 && tid < NP) {
    int gun = e.fu[C::CEN] + P.kind_slot[e.kind[C::CEN]][GOMA_SLOT_P] + tid;
    s.Pd[tid] = P.x[gun];
    bool owned = e.node[C::CEN] < P.num_owned_nodes;
    e.rsP[tid] = (owned && P.dbc_flag[gun] == 0) ? P.rowstart[gun] : -1;
  }
  // ---- phase 1: J[a][b] = sum_k x_b,k dphi_k/dxi_a   (beer_belly, mm_fill_util.c:258-276)
  for (int idx = tid; idx < NGP * DIM * DIM; idx += NPT) {
    int gp = idx / (DIM * DIM), ab = idx - gp * DIM * DIM;
    int a = ab / DIM, b = ab - a * DIM;
    s.B[gp][ab] = dot3<NN>(s.X[b], 1, &t_dphi[gp * NN * DIM + a], DIM);
  }
  bar_sync(BAR_PROD, NPT);
  // ---- phase 2: detJ, B = J^-1 by cofactors (mm_fill_util.c:386-391, :450-480)
  if (tid < NGP) {
    double *J = s.B[tid];
    double det;
    if (DIM == 2) {
      double j00 = J[0], j01 = J[1], j10 = J[2], j11 = J[3];
      det = j00 * j11 - j01 * j10;
      double rd = 1.0 / det;
      J[0] = j11 * rd;
      J[1] = -j01 * rd;
      J[2] = -j10 * rd;
      J[3] = j00 * rd;
    } else {
      double j00 = J[0], j01 = J[1], j02 = J[2], j10 = J[3], j11 = J[4], j12 = J[5], j20 = J[6], j21 = J[7],
             j22 = J[8];
      det = j00 * (j11 * j22 - j12 * j21) - j01 * (j10 * j22 - j20 * j12) + j02 * (j10 * j21 - j20 * j11);
      double rd = 1.0 / det;
      J[0] = (j11 * j22 - j21 * j12) * rd;
      J[1] = -(j01 * j22 - j21 * j02) * rd;
      J[2] = (j01 * j12 - j11 * j02) * rd;
      J[3] = -(j10 * j22 - j20 * j12) * rd;
      J[4] = (j00 * j22 - j20 * j02) * rd;
      J[5] = -(j00 * j12 - j10 * j02) * rd;
      J[6] = (j10 * j21 - j11 * j20) * rd;
      J[7] = -(j00 * j21 - j20 * j01) * rd;
      J[8] = (j00 * j11 - j10 * j01) * rd;
    }
    s.w[tid] = det * t_wt[tid];  // d_area = detJ * wt * h3, h3 = 1 (Cartesian)
  }
  bar_sync(BAR_PROD, NPT);
  // ---- phase 3: grad_phi[i][p] = sum_q B[p][q] dphi_i/dxi_q  (load_bf_grad, mm_fill_util.c:1765-1776)
  for (int idx = tid; idx < NGP * NN; idx += NPT) {
    int gp = idx / NN, i = idx - gp * NN;
    const double *B = s.B[gp];
    const double *dp = &t_dphi[(gp * NN + i) * DIM];
    const double w = s.w[gp], ph = t_phi[gp * NN + i];
    double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int p = 0; p < DIM; p++) {
#pragma unroll
      for (int q = 0; q < DIM; q++) g[p] += B[p * DIM + q] * dp[q];
    }
    e.SJa[gp][i] = make_double2(ph, g[0]);
    e.SJb[gp][i] = make_double2(g[1], g[2]);
    *reinterpret_cast<double2 *>(&e.SI[gp][i][0]) = make_double2(w * ph, w * g[0]);
    *reinterpret_cast<double2 *>(&e.SI[gp][i][2]) = make_double2(w * g[1], w * g[2]);
  }
  bar_sync(BAR_PROD, NPT);
  // ---- phase 4: field values, gradients, time derivatives at the Gauss points (load_fv, load_fv_grads)
  //      one thread per (Gauss point, field): 2 vector loads + 2 scalar loads feed DIM+2 FMAs per node
  for (int idx = tid; idx < NGP * NF; idx += NPT) {
    int gp = idx / NF, f = idx - gp * NF;
    double val[3] = {0.0, 0.0, 0.0}, dot[3] = {0.0, 0.0, 0.0}, gr[3][3] = {{0.0}};
    constexpr int M = NN / 3;
#pragma unroll
    for (int k = 0; k < M; k++) {
#pragma unroll
      for (int c = 0; c < 3; c++) {  // three independent chains
        const int kk = k + c * M;
        const double2 a = e.SJa[gp][kk], b = e.SJb[gp][kk];
        const double u = s.U[f][kk];
        val[c] += u * a.x;
        gr[c][0] += u * a.y;
        gr[c][1] += u * b.x;
        gr[c][2] += u * b.y;
        dot[c] += s.Udot[f][kk] * a.x;
      }
    }
#pragma unroll
    for (int kk = 3 * M; kk < NN; kk++) {
      const double2 a = e.SJa[gp][kk], b = e.SJb[gp][kk];
      const double u = s.U[f][kk];
      val[0] += u * a.x;
      gr[0][0] += u * a.y;
      gr[0][1] += u * b.x;
      gr[0][2] += u * b.y;
      dot[0] += s.Udot[f][kk] * a.x;
    }
    s.F[gp][f][0] = (val[0] + val[1]) + val[2];
#pragma unroll
    for (int p = 0; p < DIM; p++) s.F[gp][f][1 + p] = (gr[0][p] + gr[1][p]) + gr[2][p];
    s.F[gp][f][1 + DIM] = (dot[0] + dot[1]) + dot[2];
  }
  if (C::P1) {
    for (int gp = tid; gp < NGP; gp += NPT) {
      double v = 0.0;
#pragma unroll
      for (int p = 0; p < NP; p++) v += s.Pd[p] * t_psi[gp * (DIM + 1) + p];
      s.Pgp[gp] = v;
    }
  }
  bar_sync(BAR_PROD, NPT);
  // ---- phase 4b: per-Gauss-point terms shared by every row/column of the element
  for (int gp = tid; gp < NGP; gp += NPT) {
    double *G = e.GP[gp];
    double v[DIM], vdot[DIM], gv[DIM][DIM];  // gv[a][b] = d_b v_a
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      v[a] = s.F[gp][C::F_V + a][0];
      vdot[a] = s.F[gp][C::F_V + a][1 + DIM];
#pragma unroll
      for (int b = 0; b < DIM; b++) gv[a][b] = s.F[gp][C::F_V + a][1 + b];
    }
    const double T = C::ENERGY ? s.F[gp][C::F_T][0] : 0.0;
    const double Pr = C::P1 ? s.Pgp[gp] : s.F[gp][C::F_P][0];
    double fs[3], dfdT[3];
    momentum_source<C>(P, T, fs, dfdT);
    double div = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      div += gv[a][a];
      double adv = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; p++) adv += v[p] * gv[a][p];
      // residual, momentum a (mm_fill_momentum.c:534-662): mass + advection + source multiply phi_i
      G[C::G_RQ + a] = -P.etm_mom[0] * P.rho * vdot[a] - P.etm_mom[1] * P.rho * adv + P.etm_mom[4] * fs[a];
#pragma unroll
      for (int p = 0; p < DIM; p++) {
        // Pi[a][p] = -P delta + mu (d_a v_p + d_p v_a)  (fluid_stress, mm_fill_momentum.c:3268-3271)
        double Pi = P.mu * (gv[p][a] + gv[a][p]) - (p == a ? Pr : 0.0);
        G[C::G_RP + a * DIM + p] = -P.etm_mom[3] * Pi;
        G[C::G_GV + a * DIM + p] = c_adv * gv[a][p];
      }
    }
    G[C::G_DIV] = P.etm_cont[0] * div;
    if (C::ENERGY) {
      double adv = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; p++) {
        double gT = s.F[gp][C::F_T][1 + p];
        adv += v[p] * gT;
        G[C::G_GT + p] = ce_adv * gT;
        G[C::G_RF + p] = P.etm_energy[3] * (-P.k * gT);  // + grad_phi_i . q, q = -k grad T
      }
      G[C::G_RE] = -P.etm_energy[0] * rcp * s.F[gp][C::F_T][1 + DIM] - P.etm_energy[1] * rcp * adv +
                   P.etm_energy[4] * P.heat_source;
    }
  }
  for (int idx = tid; idx < NGP * NN; idx += NPT) {
    int gp = idx / NN, j = idx - gp * NN;
    const double2 ja = e.SJa[gp][j], jb = e.SJb[gp][j];
    const double gj[3] = {ja.y, jb.x, jb.y};
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < DIM; p++) acc += s.F[gp][C::F_V + p][0] * gj[p];
    e.VG[gp][j] = acc;
  }
  // the consumers may start on this buffer now: everything they read is written above
  __threadfence_block();
  bar_sync(BAR_PROD, NPT);
}

// residual rows + Dirichlet rows (bc_dirich.c:130-140) and the P1 pressure coupling; producer side,
// runs after the buffer has been handed to the consumers (reads it only)
template <class C>
__device__ __forceinline__ void produce_rows(const FillParams &P, Smem<C> &s, const ElemBuf<C> &e, int tid) {
  constexpr int DIM = C::DIM, NN = C::NN, NGP = C::NGP, NF = C::NF, NP = C::NP, NPT = C::NPT;
  const double *t_psi = s.tbl + C::T_PSI;
  // ---- phase 5
  for (int idx = tid; idx < NF * NN + NP; idx += NPT) {
    const bool prow = idx >= NF * NN;  // P1 continuity row
    const int f = prow ? 0 : idx / NN;
    const int i = prow ? C::CEN : idx - f * NN;
    const int gun = prow ? e.fu[C::CEN] + P.kind_slot[e.kind[C::CEN]][GOMA_SLOT_P] + (idx - NF * NN) : e.gun[f][i];
    if (e.node[i] >= P.num_owned_nodes) continue;
    const int dbc = P.dbc_flag[gun];
    if (dbc) {
      if (P.assemble_residual) acc_add(P, &P.resid[gun], dbc == 1 ? P.x[gun] - P.dbc_value[gun] : 0.0);
      if (P.assemble_jacobian) acc_add(P, &P.a[gun], 1.0);
      continue;
    }
    if (!P.assemble_residual) continue;
    double R = 0.0;
    if (prow) {
      const int p = idx - NF * NN;
      double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
      for (int gp = 0; gp + 2 < NGP; gp += 3) {
        r0 += s.w[gp] * t_psi[gp * (DIM + 1) + p] * e.GP[gp][C::G_DIV];
        r1 += s.w[gp + 1] * t_psi[(gp + 1) * (DIM + 1) + p] * e.GP[gp + 1][C::G_DIV];
        r2 += s.w[gp + 2] * t_psi[(gp + 2) * (DIM + 1) + p] * e.GP[gp + 2][C::G_DIV];
      }
      for (int gp = NGP - NGP % 3; gp < NGP; gp++) r0 += s.w[gp] * t_psi[gp * (DIM + 1) + p] * e.GP[gp][C::G_DIV];
      R = (r0 + r1) + r2;
    } else {
      const bool isT = C::ENERGY && f == C::F_T;
      if (f >= DIM && !isT) continue;
      const int q0 = isT ? C::G_RE : C::G_RQ + f, q1 = isT ? C::G_RF : C::G_RP + f * DIM;
      double r[3] = {0.0, 0.0, 0.0};
#pragma unroll 3
      for (int gp = 0; gp < NGP; gp++) {
        const double *si = e.SI[gp][i], *G = e.GP[gp];
        double t = si[0] * G[q0];
#pragma unroll
        for (int p = 0; p < DIM; p++) t += si[1 + p] * G[q1 + p];
        r[gp % 3] += t;
      }
      R = (r[0] + r[1]) + r[2];
    }
    acc_add(P, &P.resid[gun], R);
  }
  // ---- phase 7: P1 pressure coupling, S[i][a][p] = sum_gp w grad_phi_i[a] psi_p
  //      J_m_P (mm_fill_momentum.c:2091-2104) and J_c_v (mm_fill_continuity.c:686-716) share it
  if (C::P1 && P.assemble_jacobian) {
    const int elem = e.elem;
    const int poff = P.kind_slot[e.kind[C::CEN]][GOMA_SLOT_P];
    for (int idx = tid; idx < NN * DIM * NP; idx += NPT) {
      int i = idx / (DIM * NP), r = idx - i * DIM * NP;
      int a = r / NP, p = r - a * NP;
      const double S = dot3<NGP>(&e.SI[0][i][1 + a], NN * 4, &t_psi[p], DIM + 1);
      if (e.rs[a][i] >= 0) {
        size_t pq = ((size_t)elem * NN + i) * NN + C::CEN;
        mat_add<C>(P, e, e.rs[a][i], e.gun[a][i], C::CEN, (int)P.pair_full[pq], poff + p, P.etm_mom[3] * S);
      }
      if (e.rsP[p] >= 0) {
        size_t pq = ((size_t)elem * NN + C::CEN) * NN + i;
        mat_add<C>(P, e, e.rsP[p], e.fu[C::CEN] + poff + p, i, (int)P.pair_full[pq],
                   P.kind_slot[e.kind[i]][GOMA_SLOT_U + a], P.etm_cont[0] * S);
      }
    }
  }
}

// =====================================================================================
// consumer: phase 6, node-pair blocks.  Thread = (row tile of TI nodes, one column node j); the
// TI x 1 tile of DIMxDIM (+energy) blocks is accumulated in registers over the Gauss points.
// =====================================================================================
template <class C>
__device__ __forceinline__ void consume_element(const FillParams &P, const ElemBuf<C> &e, int tid) {
  constexpr int DIM = C::DIM, NN = C::NN, NGP = C::NGP, TI = C::TI, NCT = C::NCT;
  const double tfac = P.transient ? (1.0 + 2.0 * P.theta) / P.delta_t : 0.0;
  const double rcp = P.rho * P.Cp;
  const double c_adv = -P.etm_mom[1] * P.rho, c_diff = -P.etm_mom[3] * P.mu, c_mass = -P.etm_mom[0] * P.rho * tfac;
  const double ce_adv = -P.etm_energy[1] * rcp, ce_diff = -P.etm_energy[3] * P.k,
               ce_mass = -P.etm_energy[0] * rcp * tfac;
  const int elem = e.elem;
  for (int t = tid; t < C::NTILE; t += NCT) {
    const int it = t / NN, j = t - it * NN, i0 = it * TI;
    // slot-map offsets of the TI pairs, fetched before the Gauss loop so their latency is hidden
    int po[TI], pp_[TI];
#pragma unroll
    for (int ii = 0; ii < TI; ii++) {
      size_t pq = ((size_t)elem * NN + i0 + ii) * NN + j;
      po[ii] = (int)P.pair_full[pq];
      pp_[ii] = C::ENERGY ? (int)P.pair_p[pq] : 0;
    }
    double A[TI][DIM][DIM], S1[TI], S2[TI], S3[TI], ET[TI][DIM];
#pragma unroll
    for (int ii = 0; ii < TI; ii++) {
      S1[ii] = S2[ii] = S3[ii] = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        ET[ii][a] = 0.0;
#pragma unroll
        for (int b = 0; b < DIM; b++) A[ii][a][b] = 0.0;
      }
    }
    const int ngp_run = (P.debug & 2) ? 1 : NGP;
#pragma unroll 1
    for (int gp = 0; gp < ngp_run; gp++) {
      const double2 j01 = e.SJa[gp][j];
      const double2 j23 = e.SJb[gp][j];
      const double phi_j = j01.x;
      const double gj[3] = {j01.y, j23.x, j23.y};
      const double vgj = e.VG[gp][j];
      double gjs[DIM], GV[DIM][DIM], GT[DIM];
      const double *G = e.GP[gp];
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        gjs[a] = c_diff * gj[a];
        if (C::ENERGY) GT[a] = G[C::G_GT + a];
#pragma unroll
        for (int b = 0; b < DIM; b++) GV[a][b] = G[C::G_GV + a * DIM + b];
      }
#pragma unroll
      for (int ii = 0; ii < TI; ii++) {
        const double2 i01 = *reinterpret_cast<const double2 *>(&e.SI[gp][i0 + ii][0]);
        const double2 i23 = *reinterpret_cast<const double2 *>(&e.SI[gp][i0 + ii][2]);
        const double wphi = i01.x;
        const double wg[3] = {i01.y, i23.x, i23.y};
        const double pp = wphi * phi_j;
        S1[ii] += wphi * vgj;
        S3[ii] += pp;
#pragma unroll
        for (int p = 0; p < DIM; p++) S2[ii] += wg[p] * gj[p];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
#pragma unroll
          for (int b = 0; b < DIM; b++) {
            // J_m_v (mm_fill_momentum.c:1629-1712, d_Pi->v :3458-3469):
            //   -rho phi_i phi_j d_b v_a  - mu grad_phi_i[b] grad_phi_j[a]   (+ delta_ab terms below)
            A[ii][a][b] += pp * GV[a][b];
            A[ii][a][b] += wg[b] * gjs[a];
          }
          if (C::ENERGY) ET[ii][a] += pp * GT[a];  // J_e_v (mm_fill_energy.c:640)
        }
      }
    }
    // ---- scatter the tile through the slot map (load_lec, MSR branch)
    double dfdT[3] = {0.0, 0.0, 0.0};
    if (C::ENERGY && P.source_model != 0 && P.etm_mom[4] != 0.0) {
#pragma unroll
      for (int a = 0; a < DIM; a++) dfdT[a] = -P.g[a] * P.rho * P.beta * P.etm_mom[4];
    }
    const int kj = e.kind[j];
    const int colU = P.kind_slot[kj][GOMA_SLOT_U];  // velocity components are contiguous in a node
    const int colT = C::ENERGY ? P.kind_slot[kj][GOMA_SLOT_T] : 0;
#pragma unroll
    for (int ii = 0; ii < TI; ii++) {
      const int i = i0 + ii;
      const double dm = c_adv * S1[ii] + c_diff * S2[ii] + c_mass * S3[ii];
      if (i != j) {
        // distinct nodes: no diagonal inside the block, one shift for the whole block
        const int sh = e.node[j] > e.node[i] ? 1 : 0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          const long long rstart = e.rs[a][i];
          if (rstart < 0) continue;
          double *base = &P.a[rstart + po[ii] - sh];
#pragma unroll
          for (int b = 0; b < DIM; b++) acc_add(P, base + colU + b, A[ii][a][b] + (a == b ? dm : 0.0));
          if (C::ENERGY) acc_add(P, base + colT, dfdT[a] * S3[ii]);  // J_m_T (mm_std_models.c:337)
        }
        if (C::ENERGY) {
          const long long rstart = e.rs[C::F_T][i];
          if (rstart >= 0) {
            double *base = &P.a[rstart + po[ii] - pp_[ii] - sh];
#pragma unroll
            for (int b = 0; b < DIM; b++) acc_add(P, base + colU + b, ET[ii][b]);
            acc_add(P, base + colT, ce_adv * S1[ii] + ce_diff * S2[ii] + ce_mass * S3[ii]);  // J_e_T
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          const long long rstart = e.rs[a][i];
          if (rstart < 0) continue;
          const int row = e.gun[a][i];
#pragma unroll
          for (int b = 0; b < DIM; b++)
            mat_add<C>(P, e, rstart, row, j, po[ii], colU + b, A[ii][a][b] + (a == b ? dm : 0.0));
          if (C::ENERGY) mat_add<C>(P, e, rstart, row, j, po[ii], colT, dfdT[a] * S3[ii]);
        }
        if (C::ENERGY) {
          const long long rstart = e.rs[C::F_T][i];
          if (rstart >= 0) {
            const int row = e.gun[C::F_T][i];
#pragma unroll
            for (int b = 0; b < DIM; b++) mat_add<C>(P, e, rstart, row, j, po[ii] - pp_[ii], colU + b, ET[ii][b]);
            mat_add<C>(P, e, rstart, row, j, po[ii] - pp_[ii], colT,
                       ce_adv * S1[ii] + ce_diff * S2[ii] + ce_mass * S3[ii]);
          }
        }
      }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::TPE, C::MINB) fill_kernel(const __grid_constant__ FillParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<C> &s = *reinterpret_cast<Smem<C> *>(smem_raw);
  const int tid = threadIdx.x;
  constexpr int NT = C::TPE;

  // ---- stage the quadrature/basis tables once per CTA with a TMA bulk copy
  if (tid == 0) {
    mbar_init(&s.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s.mbar, C::TBL_PAD * 8);
    tma_bulk_g2s(s.tbl, P.tables, C::TBL_PAD * 8, &s.mbar);
  }
  mbar_wait(&s.mbar, 0);

  // elements of this CTA: ee = elem_begin + blockIdx.x + n * gridDim.x
  const int first = P.elem_begin + blockIdx.x;
  const int count = first < P.elem_end ? (P.elem_end - first + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (!C::SPEC) {
    // ---- role-less variant: the CTA walks producer and consumer work in turn
    long long t_elem = 0, t_cons = 0, t_rows = 0;
    for (int n = 0; n < count; n++) {
      const int ee = first + n * (int)gridDim.x;
      const int elem = P.elem_list ? P.elem_list[ee] : ee;
      long long c0 = GOMA_CLOCK();
      produce_element<C>(P, s, s.eb[0], elem, tid);
      long long c1 = GOMA_CLOCK();
      if (P.assemble_jacobian) consume_element<C>(P, s.eb[0], tid);
      long long c2 = GOMA_CLOCK();
      produce_rows<C>(P, s, s.eb[0], tid);
      bar_sync(BAR_PROD, NT);
      long long c3 = GOMA_CLOCK();
      t_elem += c1 - c0; t_cons += c2 - c1; t_rows += c3 - c2;
    }
    if (P.prof && tid == 0) {
      P.prof[blockIdx.x * 8 + 1] = t_elem; P.prof[blockIdx.x * 8 + 2] = t_rows; P.prof[blockIdx.x * 8 + 4] = t_cons;
      P.prof[blockIdx.x * 8 + 0] = 0; P.prof[blockIdx.x * 8 + 3] = 0; P.prof[blockIdx.x * 8 + 6] = count;
    }
    return;
  }

  if (tid >= C::NCT) {
    // ================= producer warps =================
    const int ptid = tid - C::NCT;
    long long t_wait = 0, t_elem = 0, t_rows = 0;
    for (int n = 0; n < count; n++) {
      const int b = n & 1;
      const int ee = first + n * (int)gridDim.x;
      const int elem = P.elem_list ? P.elem_list[ee] : ee;
      long long c0 = GOMA_CLOCK();
      if (n >= 2) bar_sync(BAR_EMPTY0 + b, NT);  // consumers are done with what was in this buffer
      long long c1 = GOMA_CLOCK();
      produce_element<C>(P, s, s.eb[b & (C::NBUF - 1)], elem, ptid);
      bar_arrive(BAR_FULL0 + b, NT);
      long long c2 = GOMA_CLOCK();
      produce_rows<C>(P, s, s.eb[b & (C::NBUF - 1)], ptid);
      bar_sync(BAR_PROD, C::NPT);  // scratch (w, F, ...) is reused by the next element
      long long c3 = GOMA_CLOCK();
      t_wait += c1 - c0; t_elem += c2 - c1; t_rows += c3 - c2;
    }
    if (P.prof && ptid == 0) {
      P.prof[blockIdx.x * 8 + 0] = t_wait; P.prof[blockIdx.x * 8 + 1] = t_elem; P.prof[blockIdx.x * 8 + 2] = t_rows;
      P.prof[blockIdx.x * 8 + 6] = count;
    }
  } else {
    // ================= consumer warps =================
    long long t_wait = 0, t_work = 0;
    for (int n = 0; n < count; n++) {
      const int b = n & 1;
      long long c0 = GOMA_CLOCK();
      bar_sync(BAR_FULL0 + b, NT);
      long long c1 = GOMA_CLOCK();
      if (P.assemble_jacobian) consume_element<C>(P, s.eb[b & (C::NBUF - 1)], tid);
      if (n + 2 < count) bar_arrive(BAR_EMPTY0 + b, NT);
      long long c2 = GOMA_CLOCK();
      t_wait += c1 - c0; t_work += c2 - c1;
    }
    if (P.prof && tid == 0) { P.prof[blockIdx.x * 8 + 3] = t_wait; P.prof[blockIdx.x * 8 + 4] = t_work; }
  }
}

}  // namespace goma_b200
