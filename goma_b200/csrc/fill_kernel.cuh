// Element assembly kernels for sm_100a: persistent CTAs, one element at a time per CTA, software-pipelined over
// the elements a CTA owns.  Per element: the gather record (built once at init) arrives by a TMA bulk copy one
// element ahead, the unknowns by cp.async under the previous element's Gauss loop; operand tables (geometry,
// basis gradients, field values, per-Gauss-point terms) are built in shared memory; the node-pair Jacobian
// blocks are accumulated in registers over the Gauss points and go straight from registers to their matrix
// slots (slot map: no search, no staging).  fill_kernel is the default; fill_kernel_ws is a warp-specialised
// variant (builder / multiplier warps over mbarrier-handed operand buffers) for the Q2/P1 Navier-Stokes block.
//
// What it replaces in the reference, per element (src/mm_fill.c:317 matrix_fill):
//   BLOCK 1   load_elem_dofptr / load_ei          -> ElemRec (build_records_kernel) + gather_state
//   per Gauss point (mm_fill.c:1253-2665):
//     load_basis_functions (mm_fill_util.c:2608)  -> constant tables staged once per CTA by TMA
//     beer_belly           (mm_fill_util.c:140)   -> phases 1-2 (J, detJ, B = J^-1 by cofactors)
//     load_bf_grad         (mm_fill_util.c:1634)  -> phase 3  (grad_phi = B . dphi/dxi)
//     load_fv / load_fv_grads (load_field_variables.c:128,2049) -> phase 4
//     assemble_momentum    (mm_fill_momentum.c:100)   residual :534-662, J_m_v :1564-1735,
//                                                     J_m_T :746-915, J_m_P :2052-2117, J_m_d :2200-2442
//     assemble_continuity  (mm_fill_continuity.c:119) residual :435-613, J_c_v :665-761, J_c_d :1004-1148,
//                                                     PSPG calc_pspg (mm_fill_stabilization.c:852)
//     assemble_energy      (mm_fill_energy.c:109)     residual :322-381, J_e_T :425-487, J_e_v :628-692, J_e_d :758-925
//     assemble_mass_transport (mm_fill_species.c:194) Fickian: J_s_s, J_s_v, J_s_d :1103-1330
//     assemble_mesh        (mm_fill_terms.c:111)      ARBITRARY / NONLINEAR: residual :421-428, J_d_d :529-576
//   BLOCK 8   put_dirichlet_in_matrix (bc_dirich.c:44)
//   load_lec  (mm_fill.c:5175, MSR branch :5241-5483) -> slot-mapped scatter (atomic | coloured | first-touch)
//
// The Jacobian uses the Cartesian closed forms of SURVEY.md App. A instead of the reference's
// zero-padded grad_phi_e / d_Pi tensors.
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
  const long long *nn_ptr;         // node-node lists with the running unknown counts along them (init only:
  const int *nn_list;              // build_records_kernel turns them into the per-element slot map)
  const unsigned short *cum_full;
  const unsigned short *cum_p;
  const unsigned *pair_first;  // first-touch masks (write-once scatter)
  const unsigned *node_first;
  const unsigned char *dbc_flag;
  const double *dbc_value;
  const unsigned char *erec;  // per-element gather records (ElemRec<C>), built once by build_records_kernel
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
  int assemble_residual, assemble_jacobian, transient;
  int scatter_mode;  // 0 fp64 atomics (any order), 1 coloured load+add+store, 2 coloured first-touch stores
  // value layout: 0 = MSR (diagonal in a[0..N), off-diagonals of row r at a[ija[r]..ija[r+1]), the reference's
  // ams->val); 1 = CSR of the owned rows with the diagonal in place (what Epetra's SumIntoGlobalValues target or a
  // GPU solver takes): position = rowptr[r] + offset of the column in the row's sorted list, no diagonal shift
  int csr;
  long long msr0;  // ija[0] = N + 1: rowptr_csr[r] = rowstart[r] - msr0 + r
  // constants
  double etm_mom[6], etm_cont[2], etm_energy[5], etm_species[5], etm_mesh[5];
  double rho, mu, k, Cp, beta, Tref, heat_source;
  double g[3];
  int source_model;  // 0 CONSTANT, 1 BOUSS (hydrostatic), 2 BOUSSINESQ
  double diffusivity[4];
  double delta_t, theta, time_value, h_elem_avg, U_norm;
  double lame_mu, lame_lambda;
  int pspg;  // 0 off, 1 global, 2 local (calc_pspg)
  double ps_scaling;
  long long *prof;  // optional per-CTA phase cycle counters (profiling build), 8 per CTA
  int *work;        // per-launch counter: the elements after the first static_rounds * gridDim.x are handed out dynamically (NULL: grid-stride)
  int static_rounds;  // >= 1
  int debug;        // experiments only: bit0 = drop the matrix stores, bit1 = skip the Gauss loop
};

template <int DIM_, int NN_, int NGP_, bool P1_, bool ENERGY_, int NSPEC_, bool ALE_, int TPE_, int TI_, int MINB_, bool WS_ = false, int VAR_ = 0>
struct Cfg {
  // variants of the hex27 Q2/P1 kernels (A/B runs): bit 0 = all 16 padded node blocks on the tensor cores
  // (no scalar remainder); bit 2 = scalar set-up phases (Jacobian of the map, field interpolation, row sums) instead of
  // the tensor-core ones; bit 3 = no tensor cores at all (the round-1 scalar 3 x 1 register-tile kernel)
  static constexpr int VAR = VAR_;
  static constexpr bool MMA_SETUP = !(VAR_ & 4);
  static constexpr int DIM = DIM_, NN = NN_, NGP = NGP_, NSPEC = NSPEC_, TI = TI_, TPE = TPE_, MINB = MINB_;
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
  // WS: warp-specialised kernel (fill_kernel_ws): TPE builder threads prepare element e+1 while NMUL multiplier
  // threads run the Gauss loop of element e; operands double-buffered, records in a ring of four
  static constexpr bool WS = WS_;
  static constexpr int NOPB = WS ? 2 : 1, NRECB = WS ? 4 : 2;
  static constexpr int NMUL = 256, NJP = (NN + 1) / 2, NTILE2 = (NN / TI) * NJP;
  static_assert(!WS || (NTILE <= NMUL && !ENERGY && NSPEC == 0 && !ALE && P1), "the warp-specialised kernel is the Q2/P1 NS block");
  static_assert(NN % TI == 0, "row tile must divide the node count");
  static_assert(TPE % 32 == 0, "whole warps");
  static constexpr int NWARP = TPE / 32;
  // constant tables in global memory: weights, dphi/dxi, P1 basis, phi.  The tensor-core configurations keep phi in
  // their operand table (filled once per CTA), so only the part before T_PHI is staged into s.tbl
  // order: weights, P1 basis, the 1-D Lagrange factors L[point][node], dL[point][node] (18), dphi/dxi, phi.  The
  // tensor-core configurations form dphi/dxi from the 1-D factors on the fly (two multiplies per value instead of
  // 17 KB of shared memory) and stage only what precedes T_DPHI
  static constexpr int T_WT = 0, T_PSI = NGP, T_L1D = T_PSI + NGP * (DIM + 1), T_DPHI = T_L1D + 18,
                       T_PHI = T_DPHI + NGP * NN * DIM;
  static constexpr int TBL = T_PHI + NGP * NN;
  static constexpr bool PHI_IN_OPERANDS = NN == 27 && NGP == 27 && P1_ && NSPEC_ == 0 && !ALE_ && !WS_ && !(VAR_ & 8);  // == MMA below
  static constexpr int TBL_PAD = ((PHI_IN_OPERANDS ? T_DPHI : TBL) + 1) & ~1;  // staged doubles, 16-byte multiple for the bulk copy
  static constexpr int TBL_GLOBAL = TBL + 2;
  // per-Gauss-point derived data (doubles): see phase 4b
  static constexpr int G_GV = 0;                                   // c_adv d_b v_a            [a][b]
  static constexpr int G_GT = G_GV + DIM * DIM;                    // ce_adv d_b T             [b]
  static constexpr int G_RQ = G_GT + (ENERGY ? DIM : 0);           // momentum residual, multiplies w phi_i   [a]
  static constexpr int G_RP = G_RQ + DIM;                          // -e3 Pi[a][p], multiplies w grad_phi_i[p]
  static constexpr int G_RE = G_RP + DIM * DIM;                    // energy residual scalar
  static constexpr int G_RF = G_RE + (ENERGY ? 1 : 0);             // e3 q[p]
  static constexpr int G_DIV = G_RF + (ENERGY ? DIM : 0);          // ec0 div v
  static constexpr int G_GY = G_DIV + 1;                           // cs_adv d_b Y_w           [w][b]
  static constexpr int G_RY = G_GY + NSPEC * DIM;                  // species residual scalar  [w]
  static constexpr int G_RFY = G_RY + NSPEC;                       // e3 j_w[p]                [w][p]
  static constexpr int G_HB = G_RFY + NSPEC * DIM;                 // PSPG: tau e1 rho d_b v_a [a][b]
  static constexpr int G_MOM = G_HB + (P1 ? 0 : DIM * DIM);        // PSPG: momentum residual  [a]
  static constexpr int G_PS = G_MOM + (P1 ? 0 : DIM);              // PSPG: tau * momentum     [a]
  // ALE (pseudo-solid mesh): mesh residual operand and the per-Gauss-point pieces of J_d_d
  static constexpr int G_ZERO = G_PS + (P1 ? 0 : DIM);             // 0.0 (rows without a phi_i term)
  static constexpr int G_RD = G_ZERO + (ALE ? 1 : 0);              // -ed3 TT[a][p], multiplies w grad_phi_i[p]
  static constexpr int G_GM = G_RD + (ALE ? DIM * DIM : 0);        // sum_c grad_d[a][c] M[b][c],  M = I - grad_d^T
  static constexpr int G_CM = G_GM + (ALE ? DIM * DIM : 0);        // sum_q cof(F)[p][q] M[b][q]   [p][b]
  static constexpr int G_C1 = G_CM + (ALE ? DIM * DIM : 0);        // ed3 lambda vc^2 vc^(-2/3)
  static constexpr int GPD = (G_C1 + (ALE ? 1 : 0) + 1) & ~1;
  static constexpr int NROWS = NF * NN + NP;  // rows of the element block
  static constexpr bool GENERAL = !P1 || NSPEC > 0 || ALE;  // generic block accumulation instead of the NS(+T) fast path
  // Gauss-point thirds of the row sums (the tensor-core configurations sum the rows in one product)
  static constexpr int NPART = ((PHI_IN_OPERANDS && !(VAR_ & 4)) || NGP < 3) ? 1 : 3;
  // hex27 Q2/P1 NS(+T): the node-pair blocks run on the FP64 tensor cores (mma.sync m8n8k4.f64, SASS DMMA): 8 x 8
  // node blocks, K = Gauss points (27 -> 28); operand tables component-major with bank-conflict-free strides
  static constexpr bool MMA = NN == 27 && NGP == 27 && P1 && !GENERAL && !WS_ && !(VAR_ & 8);
  static_assert(MMA == PHI_IN_OPERANDS, "table staging and operand layout must agree");
  static constexpr int NGK = MMA ? 28 : NGP;  // rows of the per-Gauss-point tables (row 27 = the zero K padding)
  __host__ __device__ static constexpr int slot(int f) {
    return f < DIM                       ? GOMA_SLOT_U + f
           : (ENERGY && f == F_T)        ? GOMA_SLOT_T
           : (f >= F_Y && f < F_Y + NSPEC) ? GOMA_SLOT_Y0 + (f - F_Y)
           : (ALE && f >= F_D && f < F_D + DIM) ? GOMA_SLOT_DX + (f - F_D)
                                         : GOMA_SLOT_P;
  }
};

template <class C, bool MMA = C::MMA>
struct Operands {
  // test-function side, component-major: SI[gp][0][i] = w phi_i, SI[gp][1+p][i] = w grad_phi_i[p].  A warp of the
  // Gauss loop spans at most two row tiles: each of these 8-byte loads is one shared-memory wavefront (a broadcast
  // 16-byte load costs two); the set-up phases read them with the lanes along i (conflict-free)
  double SI[C::NGP][4][C::NN];
  // trial-function side (phi_j, grad_phi_j[p]), split in two 16-byte-stride arrays so that the
  // per-lane LDS.128 of the Gauss loop are bank-conflict free
  double2 SJa[C::NGP][C::NN];  // (phi_j, g_j[0])
  double2 SJb[C::NGP][C::NN];  // (g_j[1], g_j[2])
  double VG[C::NGP][C::NN];    // v . grad_phi_j
};
// Tensor-core layout: ONE table SJ[gp][c * 28 + node], c = 0: phi (the same for every element, written once per
// CTA), 1..3: grad_phi[p], 4: v . grad_phi.  Test- and trial-function side read the same values: the quadrature
// weight w(gp) is folded into the other operand of every product (the lane that forms a B fragment multiplies it by
// the w of its Gauss point).  A fragment of mma.m8n8k4 is (8 consecutive nodes) x (4 consecutive Gauss points): with
// a Gauss-point stride of 4 (mod 8) doubles the 16 lanes of a half-warp hit 16 different bank pairs.  Column 27 of
// every component and row 27 (the K padding) stay zero from the kernel prologue on.  31 KB instead of the 52 KB of
// separate w-scaled and raw tables: three CTAs fit on an SM.
template <class C>
struct Operands<C, true> {
  static constexpr int NNP = 28, SJ_S = 5 * NNP;
  static_assert(SJ_S % 8 == 4, "conflict-free fragment loads");
  double SJ[C::NGK][SJ_S];
};
template <class C>
__device__ __forceinline__ double op_si(const Operands<C> &op, int gp, int c, int i) {
  if constexpr (C::MMA)
    return op.SJ[gp][c * 28 + i];  // (unweighted: the tensor-core configurations fold w into the other operand)
  else
    return op.SI[gp][c][i];
}

// Everything about one element that does not depend on the state vector, gathered once at init
// (what load_ei / load_elem_dofptr, mm_fill_ptrs.c:170,1136, recompute per element and per Newton
// iteration) and laid out so that ONE TMA bulk copy brings it into shared memory.
template <class C>
struct alignas(16) ElemRec {
  double X[C::DIM][C::NN];               // nodal coordinates (undeformed)
  long long rs[C::NF][C::NN];            // MSR row start of (field, node); -1 = not written here (ghost / Dirichlet)
  long long rsP[C::NP > 0 ? C::NP : 1];  // ... of the P1 pressure rows on the centroid node
  int gun[C::NF][C::NN];                 // global unknown number (gun_list)
  int gunP;                              // first P1 pressure unknown of the centroid node
  unsigned node_first;                   // bit i: this element is the first writer of node i's residual / diagonal
  unsigned first[C::NN];                 // bit j of word i: first writer of the node pair (i, j)
  unsigned short po[C::NN][C::NN];       // slot map: column offset of node j's first unknown in a row of node i
  unsigned short pp[C::ENERGY ? C::NN : 1][C::ENERGY ? C::NN : 1];  // pressure columns before j (energy rows skip them)
  unsigned char rank[C::NN];             // position of the local node in increasing global id (= column order)
  unsigned char cs[C::NN][C::NF + 1];    // offset of field f inside node j
  unsigned char flag[C::NF][C::NN];      // bits 0-1: Dirichlet flag of the unknown, bit 2: row is owned by this rank
  unsigned char flagP[4];
  unsigned char poffP;                   // offset of the first P1 unknown inside the centroid node
};

template <class C>
struct alignas(16) Smem {
  double tbl[C::TBL_PAD];
  ElemRec<C> rec[C::NRECB];  // ring: the record of the next element lands while this one is assembled
  Operands<C> op[C::NOPB];
  double GP[C::NOPB][C::NGK][C::GPD];   // per-Gauss-point derived quantities
  double F[C::NGP][C::NF][C::DIM + 2];  // value, grad[DIM], time derivative
  double X[C::ALE ? C::DIM : 1][C::ALE ? C::NN : 1];  // ALE: displaced coordinates x = X + d
  double U[C::NRECB][C::NF][C::NN];     // nodal unknowns, buffered like rec (cp.async gather)
  double Udot[C::NRECB][C::NF][C::NN];
  double Pd[C::NRECB][C::NP > 0 ? C::NP : 2];
  double w[C::NGK];  // w(gp) = detJ * weight; entry 27 (K padding of the tensor-core tables) stays zero
  double B[C::NGP][C::DIM * C::DIM];
  double Pgp[C::NGP];
  double tau, dtau[3];  // PSPG tau and d tau / d v_avg[b] (element level)
  double redR[C::NPART][C::NROWS];  // partial row sums (element_rows)
  double redS[C::NPART][C::P1 ? C::DIM * C::NN : 1][C::NP > 0 ? C::NP : 1];
  unsigned char lat[32];  // tensor-core configurations: lattice position o0 + 3 o1 + 9 o2 of the local nodes
  unsigned long long mbar;
  unsigned long long mbar_rec[C::NRECB];
  unsigned long long full[2], empty[2];  // WS: operand buffer hand-off between builders and multipliers
  int next_ee;  // index of the element this CTA takes next (dynamic hand-out)
};

static_assert(sizeof(double2) == 16, "double2 layout");

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

// per-thread asynchronous 8-byte global -> shared copy (SASS: LDGSTS), the indirect gather of x[gun]
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// barrier over the threads that run the set-up phases: the whole CTA, or the builder warps of the WS kernel
template <class C>
__device__ __forceinline__ void cta_sync() {
  if constexpr (C::WS)
    asm volatile("bar.sync 1, %0;" ::"n"(C::TPE) : "memory");
  else
    __syncthreads();
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}

#ifdef GOMA_PROFILE_PHASES
#define GOMA_CLOCK() clock64()
#else
#define GOMA_CLOCK() 0LL
#endif

// One contribution to a matrix / residual slot.  `first` = this element is the first (lowest colour)
// writer of the slot in this fill: a plain store then replaces zeroing + accumulation.
__device__ __forceinline__ void slot_add(const FillParams &P, double *addr, double val, bool first) {
  if (P.debug & 1) return;
  if (P.scatter_mode == 2) {
    // colour-ordered launches: no concurrent writer of this slot exists, so both forms are race-free;
    // both are fire-and-forget (no load round trip in the issuing warp)
    if (first)
      *addr = val;
    else
      atomicAdd(addr, val);
  } else if (P.scatter_mode == 0) {
    atomicAdd(addr, val);  // result unused -> no-return reduction
  } else {
    *addr += val;
  }
}
#ifdef GOMA_PROFILE_PHASES
__constant__ int g_store_debug;  // profiling build only: bit 2 (4) drops the reductions, bit 3 (8) the first-touch stores
#endif
// the same with the scatter mode fixed at compile time (the node-pair write-out issues thousands per element)
template <int MODE>
__device__ __forceinline__ void slot_add_m(double *addr, double val, bool first) {
#ifdef GOMA_PROFILE_PHASES
  if ((g_store_debug & 4) && !first) return;
  if ((g_store_debug & 8) && first) return;
#endif
  if (MODE == 2) {
    if (first)
      *addr = val;
    else
      atomicAdd(addr, val);
  } else if (MODE == 0) {
    atomicAdd(addr, val);
  } else {
    *addr += val;
  }
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

// D (8x8) += A (8x4, row) * B (4x8, col) in fp64 on the tensor cores (SASS: DMMA.8x8x4).  Lane l holds
// A[l/4][l%4], B[l%4][l/4] and D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d[0]), "+d"(d[1])
      : "d"(a), "d"(b));
}

// HEX27 local node -> lattice position o0 + 3 o1 + 9 o2 (Exodus / PATRAN order, rf_shape.c:1105-1200; tables.h)
__device__ const unsigned char HEX27_LATTICE[27] = {0, 2, 8, 6, 18, 20, 26, 24, 1, 5, 7, 3, 9, 11, 17, 15, 19, 23, 25, 21, 13, 4, 22, 12, 14, 10, 16};

// dphi_k / dxi_a at Gauss point gp from the 1-D Lagrange factors: the product make_tables() evaluates
// ((x0 * x1) * x2, x_d = dL or L of direction d), so the value has the bits of the table entry
__device__ __forceinline__ double dphi_from_1d(const double *__restrict__ l1d, const unsigned char *__restrict__ lat, int gp, int k, int a) {
  const int o = lat[k];
  const int o0 = o % 3, o1 = (o / 3) % 3, o2 = o / 9;
  const int p0 = gp % 3, p1 = (gp / 3) % 3, p2 = gp / 9;
  const double x0 = l1d[(a == 0 ? 9 : 0) + p0 * 3 + o0];
  const double x1 = l1d[(a == 1 ? 9 : 0) + p1 * 3 + o1];
  const double x2 = l1d[(a == 2 ? 9 : 0) + p2 * 3 + o2];
  return (x0 * x1) * x2;
}

// length-N dot product with three independent accumulation chains (hides DFMA latency in the
// low-parallelism set-up phases)
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
// phases 0-4: gather + operand tables for one element
// =====================================================================================
#ifdef GOMA_PROFILE_PHASES
#define GOMA_STAMP(k) do { long long t_ = clock64(); if (stamps) stamps[k] += t_ - last_; last_ = t_; } while (0)
#else
#define GOMA_STAMP(k) do { } while (0)
#endif

// One-off (init): fill the gather record of every element.  Restates load_ei / load_elem_dofptr
// (mm_fill_ptrs.c:170-1050,1136-1530) for the in-scope variables plus the Dirichlet / ownership
// tests of load_lec (mm_fill.c:5374) and put_dirichlet_in_matrix (bc_dirich.c:86-140).
template <class C>
__global__ void build_records_kernel(const FillParams P, int num_elems) {
  constexpr int DIM = C::DIM, NN = C::NN, NF = C::NF, NP = C::NP;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= num_elems) return;
  ElemRec<C> &r = reinterpret_cast<ElemRec<C> *>(const_cast<unsigned char *>(P.erec))[e];
  int node[NN], fu[NN], kind[NN];
  for (int k = 0; k < NN; k++) {
    const int nd = P.conn[(size_t)e * NN + k];
    node[k] = nd;
    fu[k] = P.first_unknown[nd];
    kind[k] = P.node_kind[nd];
    r.first[k] = P.pair_first ? P.pair_first[(size_t)e * NN + k] : 0u;
    for (int d = 0; d < DIM; d++) r.X[d][k] = P.coord[d][nd];
  }
  r.node_first = P.node_first ? P.node_first[e] : 0u;
  // slot map: column offset of node j's first unknown in a row of node i = the running unknown count at node j's
  // position in node i's sorted neighbour list (replaces the in_list search of load_lec, mm_fill.c:5461)
  for (int i = 0; i < NN; i++) {
    const long long b = P.nn_ptr[node[i]];
    const int len = (int)(P.nn_ptr[node[i] + 1] - b);
    for (int j = 0; j < NN; j++) {
      int lo = 0, hi = len;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (P.nn_list[b + mid] < node[j])
          lo = mid + 1;
        else
          hi = mid;
      }
      if (lo >= len || P.nn_list[b + lo] != node[j]) {
        P.flags[3] = 1;  // "Could not find vbl in sparse matrix"
        lo = 0;
      }
      r.po[i][j] = P.cum_full[b + lo];
      if (C::ENERGY) r.pp[i][j] = P.cum_p[b + lo];
    }
  }
  for (int k = 0; k < NN; k++) {
    int rk = 0;
    for (int m = 0; m < NN; m++) rk += node[m] < node[k] ? 1 : 0;
    r.rank[k] = (unsigned char)rk;
    const bool owned = node[k] < P.num_owned_nodes;
    for (int f = 0; f < NF; f++) {
      const int cs = P.kind_slot[kind[k]][C::slot(f)];
      const int gun = fu[k] + cs;
      const int dbc = P.dbc_flag[gun];
      r.cs[k][f] = (unsigned char)cs;
      r.gun[f][k] = gun;
      r.flag[f][k] = (unsigned char)(dbc | (owned ? 4 : 0));
      r.rs[f][k] = (owned && dbc == 0) ? (P.csr ? P.rowstart[gun] - P.msr0 + gun : P.rowstart[gun]) : -1;
    }
    r.cs[k][NF] = 0;
  }
  r.gunP = 0;
  r.poffP = 0;
  for (int p = 0; p < 4; p++) r.flagP[p] = 0;
  r.rsP[0] = -1;
  if (C::P1) {
    const int poff = P.kind_slot[kind[C::CEN]][GOMA_SLOT_P];
    const bool owned = node[C::CEN] < P.num_owned_nodes;
    r.poffP = (unsigned char)poff;
    r.gunP = fu[C::CEN] + poff;
    for (int p = 0; p < NP; p++) {
      const int gun = r.gunP + p;
      const int dbc = P.dbc_flag[gun];
      r.flagP[p] = (unsigned char)(dbc | (owned ? 4 : 0));
      r.rsP[p] = (owned && dbc == 0) ? (P.csr ? P.rowstart[gun] - P.msr0 + gun : P.rowstart[gun]) : -1;
    }
  }
}

// indirect gather of the state for one element: x[gun] (and xdot[gun]) -> shared memory, asynchronously
template <class C>
__device__ __forceinline__ void gather_state(const FillParams &P, Smem<C> &s, int buf, int tid, int NT = C::TPE) {
  constexpr int NN = C::NN, NF = C::NF;
  const ElemRec<C> &r = s.rec[buf];
  for (int idx = tid; idx < NF * NN; idx += NT) {
    const int gun = (&r.gun[0][0])[idx];
    cp_async8(&(&s.U[buf][0][0])[idx], &P.x[gun]);
    if (P.transient) cp_async8(&(&s.Udot[buf][0][0])[idx], &P.xdot[gun]);
  }
  if (C::P1 && tid < C::NP) cp_async8(&s.Pd[buf][tid], &P.x[r.gunP + tid]);
  cp_async_commit();
}

template <class C>
__device__ __forceinline__ void build_element(const FillParams &P, Smem<C> &s, int buf, int bo, int tid, long long *stamps) {
#ifdef GOMA_PROFILE_PHASES
  long long last_ = clock64();
#endif
  (void)stamps;
  constexpr int DIM = C::DIM, NN = C::NN, NGP = C::NGP, NF = C::NF, NP = C::NP, NT = C::TPE;
  const double *t_wt = s.tbl + C::T_WT;
  const double *t_phi = s.tbl + (C::MMA ? 0 : C::T_PHI);  // [gp][NN] (tensor-core configurations: in op.SJ)
  (void)t_phi;
  const double *t_dphi = s.tbl + (C::MMA ? 0 : C::T_DPHI);  // [gp][NN][DIM] (tensor-core configurations: dphi_from_1d)
  (void)t_dphi;
  const double *t_psi = s.tbl + C::T_PSI;    // [gp][DIM+1]
  const double rcp = P.rho * P.Cp;
  const double c_adv = -P.etm_mom[1] * P.rho;
  const double ce_adv = -P.etm_energy[1] * rcp;
  Operands<C> &op = s.op[bo];
  const ElemRec<C> &rec = s.rec[buf];
  const double (*U)[NN] = s.U[buf];
  const double (*Udot)[NN] = s.Udot[buf];

  // ---- ALE: the map is built on the displaced coordinates x = X + d (beer_belly, mm_fill_util.c:258-276)
  if (C::ALE) {
    for (int idx = tid; idx < DIM * NN; idx += NT) {
      const int d = idx / NN, k = idx - d * NN;
      s.X[d][k] = rec.X[d][k] + U[C::F_D + d][k];
    }
    cta_sync<C>();
  }
  GOMA_STAMP(0);
  // ---- phase 1: J[a][b] = sum_k x_b,k dphi_k/dxi_a   (beer_belly, mm_fill_util.c:258-276)
  if constexpr (C::MMA && C::MMA_SETUP) {
    // tensor cores: rows m = (Gauss point, a) (81 -> 88), columns b (3 -> 8), K = nodes (27 -> 28)
    const int warp = tid >> 5, lane = tid & 31, r = lane >> 2, kq = lane & 3;
    for (int mt = warp; mt < 11; mt += C::NWARP) {
      const int m = min(mt * 8 + r, 80);
      double acc[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < 7; ks++) {
        const int k = 4 * ks + kq;
        const bool live = k < NN;
        const double av = live ? dphi_from_1d(s.tbl + C::T_L1D, s.lat, m / 3, k, m % 3) : 0.0;
        const double bv = (live && r < DIM) ? rec.X[r][k] : 0.0;
        dmma884(acc, av, bv);
      }
      const int mo = mt * 8 + r;
      if (mo < 81) {
#pragma unroll
        for (int c = 0; c < 2; c++)
          if (2 * kq + c < DIM) (&s.B[0][0])[mo * DIM + 2 * kq + c] = acc[c];
      }
    }
  } else
  //      one thread per (Gauss point, a, third of the nodes): each dphi load feeds DIM FMAs
  {
    constexpr int NITEM = NGP * DIM, Q = (NN + 3) / 4;
    constexpr int NROUND = (NITEM * 4 + NT - 1) / NT;
#pragma unroll 1
    for (int rnd = 0; rnd < NROUND; rnd++) {
      const int idx = rnd * NT + tid;
      const int item = idx >> 2, c = idx & 3;
      const bool live = item < NITEM;
      const int gp = live ? item / DIM : 0, a = live ? item - gp * DIM : 0;
      const int k0 = c * Q, k1 = (k0 + Q < NN) ? k0 + Q : NN;
      double acc[3] = {0.0, 0.0, 0.0};
      if (live) {
        const double *dp = &t_dphi[gp * NN * DIM + a];
#pragma unroll 4
        for (int k = k0; k < k1; k++) {
          const double d = C::MMA ? dphi_from_1d(s.tbl + C::T_L1D, s.lat, gp, k, a) : dp[k * DIM];
#pragma unroll
          for (int b = 0; b < DIM; b++) acc[b] += (C::ALE ? s.X[b][k] : rec.X[b][k]) * d;
        }
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
        for (int b = 0; b < DIM; b++) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
      if (live && c == 0) {
#pragma unroll
        for (int b = 0; b < DIM; b++) s.B[gp][a * DIM + b] = acc[b];
      }
    }
  }
  cta_sync<C>();
  GOMA_STAMP(1);
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
    // (zero_detJ, mm_fill_util.c:335-343, is only ever raised inside beer_belly's SHELL / TRISHELL branch (:312-344):
    //  for the continuum elements of this path the reference assembles whatever |detJ| is, and so does this kernel)
    s.w[tid] = det * t_wt[tid];  // d_area = detJ * wt * h3, h3 = 1 (Cartesian)
  }
  if (!C::P1 && tid == NT - 1) {
    // BLOCK 1.5 of matrix_fill (mm_fill.c:754-787) + the tau of calc_pspg (mm_fill_stabilization.c:1030-1096)
    double tau = 0.0, dtau[3] = {0.0, 0.0, 0.0};
    if (P.pspg == 1) {  // global: element Reynolds number from the global norms, no Jacobian dependence
      const double Re = P.rho * P.U_norm * P.h_elem_avg / (2.0 * P.mu);
      tau = Re <= 3.0 ? P.ps_scaling * P.h_elem_avg * P.h_elem_avg / (12.0 * P.mu)
                      : P.ps_scaling * P.h_elem_avg / (2.0 * P.rho * P.U_norm);
    } else if (P.pspg == 2) {  // local: h_elem_siz (mm_fill_aux.c:844) and element_velocity (:759)
      double hh = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        const double *xx = C::ALE ? s.X[a] : rec.X[a];
        if (DIM == 2) {
          const double h0 = 0.5 * (xx[1] + xx[2]) - 0.5 * (xx[0] + xx[3]);
          const double h1 = 0.5 * (xx[0] + xx[1]) - 0.5 * (xx[2] + xx[3]);
          hh += h0 * h0 + h1 * h1;
        } else {
          const double p1 = 0.25 * (xx[0] + xx[1] + xx[2] + xx[3]), p2 = 0.25 * (xx[1] + xx[2] + xx[5] + xx[6]);
          const double p3 = 0.25 * (xx[2] + xx[3] + xx[6] + xx[7]), p4 = 0.25 * (xx[0] + xx[1] + xx[4] + xx[5]);
          const double p5 = 0.25 * (xx[0] + xx[3] + xx[4] + xx[7]), p6 = 0.25 * (xx[4] + xx[5] + xx[6] + xx[7]);
          hh += (p2 - p5) * (p2 - p5) + (p3 - p4) * (p3 - p4) + (p1 - p6) * (p1 - p6);
        }
      }
      hh /= (double)DIM;
      double vavg[3] = {0.0, 0.0, 0.0}, vv = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        for (int k = 0; k < NN; k++) vavg[a] += U[C::F_V + a][k] / (double)NN;  // I_Q1: mean of the nodes
        vv += vavg[a] * vavg[a];
      }
      double tau1 = P.rho * P.rho * vv / hh + 9.0 * P.mu * P.mu / (hh * hh);
      if (P.transient) tau1 += 4.0 / (P.delta_t * P.delta_t);
      tau = P.ps_scaling / sqrt(tau1);
#pragma unroll
      for (int b = 0; b < DIM; b++) dtau[b] = -tau / tau1 * P.rho * P.rho / hh * vavg[b] / (double)NN;
    }
    s.tau = tau;
#pragma unroll
    for (int b = 0; b < 3; b++) s.dtau[b] = dtau[b];
  }
  cta_sync<C>();
  GOMA_STAMP(2);
  // ---- phase 3: grad_phi[i][p] = sum_q B[p][q] dphi_i/dxi_q  (load_bf_grad, mm_fill_util.c:1765-1776)
  for (int idx = tid; idx < NGP * NN; idx += NT) {
    int gp = idx / NN, i = idx - gp * NN;
    const double *B = s.B[gp];
    double dpv[3] = {0.0, 0.0, 0.0};
    if constexpr (C::MMA) {
#pragma unroll
      for (int q = 0; q < DIM; q++) dpv[q] = dphi_from_1d(s.tbl + C::T_L1D, s.lat, gp, i, q);
    } else {
#pragma unroll
      for (int q = 0; q < DIM; q++) dpv[q] = t_dphi[(gp * NN + i) * DIM + q];
    }
    const double *dp = dpv;
    double ph;
    if constexpr (C::MMA)
      ph = op.SJ[gp][i];  // written once in the kernel prologue
    else
      ph = t_phi[gp * NN + i];
    const double w = s.w[gp];
    double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int p = 0; p < DIM; p++) {
#pragma unroll
      for (int q = 0; q < DIM; q++) g[p] += B[p * DIM + q] * dp[q];
    }
    if constexpr (C::MMA) {
#pragma unroll
      for (int p = 0; p < DIM; p++) op.SJ[gp][(1 + p) * 28 + i] = g[p];
      (void)w;
      (void)ph;
    } else {
      op.SJa[gp][i] = make_double2(ph, g[0]);
      op.SJb[gp][i] = make_double2(g[1], g[2]);
      op.SI[gp][0][i] = w * ph;
      op.SI[gp][1][i] = w * g[0];
      op.SI[gp][2][i] = w * g[1];
      op.SI[gp][3][i] = w * g[2];
    }
  }
  cta_sync<C>();
  GOMA_STAMP(3);
  // ---- phase 4: field values, gradients, time derivatives at the Gauss points (load_fv, load_fv_grads)
  if constexpr (C::MMA && C::MMA_SETUP) {
    // tensor cores: F_q[gp][f] = sum_k SJ_q[gp][k] U[f][k], q = value | grad_0..2 | time derivative (phi with Udot):
    // rows = Gauss points (27 -> 32), columns = fields (NF -> 8), K = nodes (27 -> 28)
    const int warp = tid >> 5, lane = tid & 31, r = lane >> 2, kq = lane & 3;
    const int nq = P.transient ? 5 : 4;
    for (int job = warp; job < 4 * nq; job += C::NWARP) {
      const int mt = job & 3, q = job >> 2;
      const int gpr = min(mt * 8 + r, 27);  // row 27 of the tables is zero
      const double *arow = &op.SJ[gpr][(q == 4 ? 0 : q) * 28];
      const double (*Usrc)[NN] = q == 4 ? Udot : U;
      double acc[2] = {0.0, 0.0};
#pragma unroll
      for (int ks = 0; ks < 7; ks++) {
        const int k = 4 * ks + kq;
        const double av = arow[k];  // column 27 is zero
        const double bv = (k < NN && r < NF) ? Usrc[r][k] : 0.0;
        dmma884(acc, av, bv);
      }
      const int gp = mt * 8 + r;
      if (gp < NGP) {
#pragma unroll
        for (int c = 0; c < 2; c++)
          if (2 * kq + c < NF) s.F[gp][2 * kq + c][q == 4 ? 1 + DIM : q] = acc[c];
      }
    }
    if (!P.transient)
      for (int idx = tid; idx < NGP * NF; idx += NT) s.F[idx / NF][idx % NF][1 + DIM] = 0.0;
  } else
  //      four lanes per Gauss point, each over a quarter of the nodes for ALL fields (one load of the basis
  //      functions feeds NF x (DIM+2) FMAs), combined by shuffles
  {
    constexpr int Q = (NN + 3) / 4;
    constexpr int NROUND = (NGP * 4 + NT - 1) / NT;
#pragma unroll 1
    for (int rnd = 0; rnd < NROUND; rnd++) {
      const int idx = rnd * NT + tid;
      const int gpi = idx >> 2, c = idx & 3;
      const bool live = gpi < NGP;
      const int gp = live ? gpi : 0;
      double acc[NF][DIM + 2];
#pragma unroll
      for (int f = 0; f < NF; f++)
#pragma unroll
        for (int q = 0; q < DIM + 2; q++) acc[f][q] = 0.0;
      const int k0 = c * Q, k1 = (k0 + Q < NN) ? k0 + Q : NN;
      if (live) {
#pragma unroll 2
        for (int kk = k0; kk < k1; kk++) {
          double2 a, b;
          if constexpr (C::MMA) {
            a = make_double2(op.SJ[gp][kk], op.SJ[gp][28 + kk]);
            b = make_double2(op.SJ[gp][56 + kk], op.SJ[gp][84 + kk]);
          } else {
            a = op.SJa[gp][kk];
            b = op.SJb[gp][kk];
          }
#pragma unroll
          for (int f = 0; f < NF; f++) {
            const double u = U[f][kk];
            acc[f][0] += u * a.x;
            acc[f][1] += u * a.y;
            acc[f][2] += u * b.x;
            if (DIM == 3) acc[f][3] += u * b.y;
            if (P.transient) acc[f][DIM + 1] += Udot[f][kk] * a.x;
          }
        }
      }
#pragma unroll
      for (int f = 0; f < NF; f++)
#pragma unroll
        for (int q = 0; q < DIM + 2; q++) {
          double v = acc[f][q];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          acc[f][q] = v;
        }
      if (live) {
        // lane c of the group stores fields c, c+4, ...
#pragma unroll
        for (int f = 0; f < NF; f++)
          if ((f & 3) == c) {
#pragma unroll
            for (int q = 0; q < DIM + 2; q++) s.F[gp][f][q] = acc[f][q];
          }
      }
    }
  }
  if (C::P1) {
    for (int gp = tid; gp < NGP; gp += NT) {
      double v = 0.0;
#pragma unroll
      for (int p = 0; p < NP; p++) v += s.Pd[buf][p] * t_psi[gp * (DIM + 1) + p];
      s.Pgp[gp] = v;
    }
  }
  cta_sync<C>();
  GOMA_STAMP(4);
  // ---- phase 4b: per-Gauss-point terms shared by every row/column of the element
  //      one thread per (Gauss point, job): momentum component a | energy | species | PSPG | ALE
  constexpr int J_EN = DIM, J_SP = J_EN + (C::ENERGY ? 1 : 0), J_PS = J_SP + (C::NSPEC > 0 ? 1 : 0),
                J_AL = J_PS + (C::P1 ? 0 : 1), NJOB = J_AL + (C::ALE ? 1 : 0);
  for (int idx = tid; idx < NGP * NJOB; idx += NT) {
    const int gp = idx / NJOB, job = idx - gp * NJOB;
    double *G = s.GP[bo][gp];
    double v[DIM], vdot[DIM], gv[DIM][DIM];  // gv[a][b] = d_b v_a
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      v[a] = s.F[gp][C::F_V + a][0];
      vdot[a] = s.F[gp][C::F_V + a][1 + DIM];
#pragma unroll
      for (int b = 0; b < DIM; b++) gv[a][b] = s.F[gp][C::F_V + a][1 + b];
    }
    // convection velocity v - xdot_mesh (get_convection_velocity, mm_fill_species.c:9479-9492;
    // x_dot of assemble_momentum, mm_fill_momentum.c:414-416); Udot is zero in steady runs
    double vc[DIM];
#pragma unroll
    for (int a = 0; a < DIM; a++) vc[a] = v[a] - (C::ALE ? s.F[gp][C::F_D + a][1 + DIM] : 0.0);
    const double T = C::ENERGY ? s.F[gp][C::F_T][0] : 0.0;
    const double Pr = C::P1 ? s.Pgp[gp] : s.F[gp][C::F_P][0];
    double fs[3], dfdT[3];
    momentum_source<C>(P, T, fs, dfdT);
    double div = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; a++) div += gv[a][a];
    if (job == 0) G[C::G_DIV] = P.etm_cont[0] * div;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      if (job != a) continue;
      double adv = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; p++) adv += vc[p] * gv[a][p];
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
    if (C::NSPEC > 0 && job == J_SP) {
#pragma unroll
    for (int w = 0; w < C::NSPEC; w++) {
      // assemble_mass_transport, Fickian constant-D, concentration form (mm_fill_species.c:527-640)
      double adv = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; p++) {
        const double gY = s.F[gp][C::F_Y + w][1 + p];
        adv += vc[p] * gY;
        G[C::G_GY + w * DIM + p] = -P.etm_species[1] * gY;
        G[C::G_RFY + w * DIM + p] = P.etm_species[3] * (-P.diffusivity[w] * gY);
      }
      G[C::G_RY + w] = -P.etm_species[0] * s.F[gp][C::F_Y + w][1 + DIM] - P.etm_species[1] * adv;
    }
    }
    if (!C::P1 && job == J_PS) {
      // calc_pspg (mm_fill_stabilization.c:1281-1319): momentum residual without the viscous term
      const double tau = s.tau;
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        double adv = 0.0;
#pragma unroll
        for (int p = 0; p < DIM; p++) {
          adv += v[p] * gv[a][p];
          G[C::G_HB + a * DIM + p] = tau * P.etm_mom[1] * P.rho * gv[a][p];
        }
        const double mom = P.etm_mom[0] * P.rho * vdot[a] + P.etm_mom[1] * P.rho * adv +
                           P.etm_mom[3] * s.F[gp][C::F_P][1 + a] - P.etm_mom[4] * fs[a];
        G[C::G_MOM + a] = mom;
        G[C::G_PS + a] = tau * mom;
      }
    }
    if (C::ENERGY && job == J_EN) {
      double adv = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; p++) {
        double gT = s.F[gp][C::F_T][1 + p];
        adv += vc[p] * gT;
        G[C::G_GT + p] = ce_adv * gT;
        G[C::G_RF + p] = P.etm_energy[3] * (-P.k * gT);  // + grad_phi_i . q, q = -k grad T
      }
      G[C::G_RE] = -P.etm_energy[0] * rcp * s.F[gp][C::F_T][1 + DIM] - P.etm_energy[1] * rcp * adv +
                   P.etm_energy[4] * P.heat_source;
    }
    if (C::ALE && job == J_AL) {
      // belly_flop (mm_fill_solid.c:77-1120): grad_d, Eulerian strain of the NONLINEAR model, volume change;
      // mesh_stress_tensor (:3208-3287): TT = lambda vs I + 2 mu E,  vs = 3 (vc^(1/3) - 1)
      double Gd[DIM][DIM], M[DIM][DIM], Fm[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}, cof[3][3];
#pragma unroll
      for (int p = 0; p < DIM; p++)
#pragma unroll
        for (int q = 0; q < DIM; q++) {
          Gd[p][q] = s.F[gp][C::F_D + q][1 + p];  // grad_d[p][q] = d_p d_q
          Fm[p][q] = (p == q ? 1.0 : 0.0) - Gd[p][q];
        }
#pragma unroll
      for (int b = 0; b < DIM; b++)
#pragma unroll
        for (int q = 0; q < DIM; q++) M[b][q] = (q == b ? 1.0 : 0.0) - Gd[b][q];
      cof[0][0] = Fm[1][1] * Fm[2][2] - Fm[1][2] * Fm[2][1];
      cof[0][1] = Fm[1][2] * Fm[2][0] - Fm[1][0] * Fm[2][2];
      cof[0][2] = Fm[1][0] * Fm[2][1] - Fm[1][1] * Fm[2][0];
      cof[1][0] = Fm[0][2] * Fm[2][1] - Fm[0][1] * Fm[2][2];
      cof[1][1] = Fm[0][0] * Fm[2][2] - Fm[0][2] * Fm[2][0];
      cof[1][2] = Fm[0][1] * Fm[2][0] - Fm[0][0] * Fm[2][1];
      cof[2][0] = Fm[0][1] * Fm[1][2] - Fm[0][2] * Fm[1][1];
      cof[2][1] = Fm[0][2] * Fm[1][0] - Fm[0][0] * Fm[1][2];
      cof[2][2] = Fm[0][0] * Fm[1][1] - Fm[0][1] * Fm[1][0];
      const double detF = Fm[0][0] * cof[0][0] + Fm[0][1] * cof[0][1] + Fm[0][2] * cof[0][2];
      if (detF <= 0.0) P.flags[0] = 1;  // neg_elem_volume (mm_fill_solid.c:659-663, :811-815)
      const double vch = 1.0 / detF, cb = cbrt(vch);
      const double vs = 3.0 * (cb - 1.0);
      const double ed3 = P.etm_mesh[3];
      G[C::G_ZERO] = 0.0;
      G[C::G_C1] = ed3 * P.lame_lambda * vch * vch / (cb * cb);
#pragma unroll
      for (int a = 0; a < DIM; a++)
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          double E = 0.5 * (Gd[a][b] + Gd[b][a]), gm = 0.0, cm = 0.0;
#pragma unroll
          for (int c = 0; c < DIM; c++) {
            E -= 0.5 * Gd[a][c] * Gd[b][c];
            gm += Gd[a][c] * M[b][c];
            cm += cof[a][c] * M[b][c];
          }
          G[C::G_RD + a * DIM + b] = -ed3 * (P.lame_lambda * vs * (a == b ? 1.0 : 0.0) + 2.0 * P.lame_mu * E);
          G[C::G_GM + a * DIM + b] = gm;
          G[C::G_CM + a * DIM + b] = cm;
        }
    }
  }
  for (int idx = tid; idx < NGP * NN; idx += NT) {
    int gp = idx / NN, j = idx - gp * NN;
    double gj[3];
    if constexpr (C::MMA) {
#pragma unroll
      for (int p = 0; p < 3; p++) gj[p] = op.SJ[gp][(1 + p) * 28 + j];
    } else {
      const double2 ja = op.SJa[gp][j], jb = op.SJb[gp][j];
      gj[0] = ja.y, gj[1] = jb.x, gj[2] = jb.y;
    }
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < DIM; p++)
      acc += (s.F[gp][C::F_V + p][0] - (C::ALE ? s.F[gp][C::F_D + p][1 + DIM] : 0.0)) * gj[p];
    if constexpr (C::MMA)
      op.SJ[gp][4 * 28 + j] = acc;
    else
      op.VG[gp][j] = acc;
  }
  cta_sync<C>();
  GOMA_STAMP(5);
}

// =====================================================================================
// phases 5 and 7: residual rows + Dirichlet rows (bc_dirich.c:130-140), P1 pressure coupling
// =====================================================================================
template <class C>
__device__ __forceinline__ void element_rows(const FillParams &P, Smem<C> &s, int buf, int bo, int tid) {
  constexpr int DIM = C::DIM, NN = C::NN, NGP = C::NGP, NF = C::NF, NP = C::NP, NT = C::TPE;
  constexpr int NROW = C::NROWS, NPART = C::NPART;
  const double *t_psi = s.tbl + C::T_PSI;
  const Operands<C> &op = s.op[bo];
  const ElemRec<C> &rec = s.rec[buf];
  // ---- part A: every row (field f of node i | P1 continuity row p) is summed over a third of the Gauss
  //      points by one thread; the velocity rows accumulate the P1 pressure coupling on the way:
  //      S[i][a][p] = sum_gp w grad_phi_i[a] psi_p, shared by J_m_P (mm_fill_momentum.c:2091-2104) and
  //      J_c_v (mm_fill_continuity.c:686-716)
  if constexpr (C::MMA && C::MMA_SETUP) {
    // tensor cores.  Warps 0-3: the velocity (and temperature) rows of 8 nodes each,
    //   R[i][n] = sum_gp sum_c SI_c[gp][i] Gc[gp][n],  n = a: (G_RQ+a | G_RP+3a+p),  n = 3: (G_RE | G_RF+p)
    // warps 4-7: the P1 coupling sums S[a][i][p] = sum_gp SI_{1+a}[gp][i] psi[gp][p] (12 tiles of 8 nodes x 4)
    const int warp = tid >> 5, lane = tid & 31, r = lane >> 2, kq = lane & 3;
    if (warp < 4) {
      const int ia = min(warp * 8 + r, 27);
      const int n = r;  // column of the B fragment this lane feeds
      const bool nv = n < DIM, nT = C::ENERGY && n == C::F_T;
      const int q0 = nv ? C::G_RQ + n : C::G_RE, q1 = nv ? C::G_RP + n * DIM : C::G_RF;
      double acc[2] = {0.0, 0.0};
#pragma unroll 1
      for (int ks = 0; ks < 7; ks++) {
        const int gp = 4 * ks + kq;  // row 27 of the tables and w[27] are zero
        const double *si = &op.SJ[gp][ia];
        const double *G = s.GP[bo][gp];
        const double wv = (nv || nT) ? s.w[gp] : 0.0;  // the quadrature weight rides on the B operand
        dmma884(acc, si[0], wv * G[q0]);
#pragma unroll
        for (int p = 0; p < DIM; p++) dmma884(acc, si[28 * (1 + p)], wv * G[q1 + p]);
      }
      const int i = warp * 8 + r;
      if (i < NN) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int col = 2 * kq + c;
          if (col < DIM || (C::ENERGY && col == C::F_T)) s.redR[0][col * NN + i] = acc[c];
        }
      }
    } else {
      for (int job = warp - 4; job < 4 * DIM; job += 4) {
        const int mt = job & 3, a = job >> 2;
        const int ia = min(mt * 8 + r, 27);
        double acc[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < 7; ks++) {
          const int gp = 4 * ks + kq;
          const double bv = (gp < NGP && r < NP) ? s.w[gp] * t_psi[gp * (DIM + 1) + r] : 0.0;
          dmma884(acc, op.SJ[gp][(1 + a) * 28 + ia], bv);
        }
        const int i = mt * 8 + r;
        if (i < NN) {
#pragma unroll
          for (int c = 0; c < 2; c++)
            if (2 * kq + c < NP) s.redS[0][a * NN + i][2 * kq + c] = acc[c];
        }
      }
    }
  } else
  //      velocity rows: one thread per (node i, third) does all DIM components off one load of the test functions
  for (int idx = tid; idx < NN * NPART; idx += NT) {
    const int c = idx / NN, i = idx - c * NN;
    const int gp0 = c * NGP / NPART, gp1 = (c + 1) * NGP / NPART;
    double R[DIM], S[DIM][NP > 0 ? NP : 1];
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      R[a] = 0.0;
#pragma unroll
      for (int p = 0; p < NP; p++) S[a][p] = 0.0;
    }
    for (int gp = gp0; gp < gp1; gp++) {
      const double wq = C::MMA ? s.w[gp] : 1.0;  // (the tensor-core table is unweighted)
      const double2 s01 = make_double2(wq * op_si<C>(op, gp, 0, i), wq * op_si<C>(op, gp, 1, i));
      const double2 s23 = make_double2(wq * op_si<C>(op, gp, 2, i), wq * op_si<C>(op, gp, 3, i));
      const double sg[3] = {s01.y, s23.x, s23.y};
      const double *G = s.GP[bo][gp];
      double ps[NP > 0 ? NP : 1];
#pragma unroll
      for (int p = 0; p < NP; p++) ps[p] = t_psi[gp * (DIM + 1) + p];
#pragma unroll
      for (int a = 0; a < DIM; a++) {
        double t = s01.x * G[C::G_RQ + a];
#pragma unroll
        for (int p = 0; p < DIM; p++) t += sg[p] * G[C::G_RP + a * DIM + p];
        R[a] += t;
#pragma unroll
        for (int p = 0; p < NP; p++) S[a][p] += sg[a] * ps[p];
      }
    }
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      s.redR[c][a * NN + i] = R[a];
#pragma unroll
      for (int p = 0; p < NP; p++) s.redS[c][a * NN + i][p] = S[a][p];
    }
  }
  //      the other rows: one thread per (row, third)
  for (int idx = tid; idx < (NROW - DIM * NN) * NPART; idx += NT) {
    const int c = idx / (NROW - DIM * NN), r = DIM * NN + idx - c * (NROW - DIM * NN);
    if (C::MMA && C::MMA_SETUP && r < NF * NN) continue;  // temperature rows: summed with the velocity rows on the tensor cores
    const int gp0 = c * NGP / NPART, gp1 = (c + 1) * NGP / NPART;
    const bool prow = r >= NF * NN;  // P1 continuity row
    const int f = prow ? 0 : r / NN;
    const int i = prow ? C::CEN : r - f * NN;
    double R = 0.0;
    if (prow) {
      const int p = r - NF * NN;
      for (int gp = gp0; gp < gp1; gp++) R += s.w[gp] * t_psi[gp * (DIM + 1) + p] * s.GP[bo][gp][C::G_DIV];
    } else {
      const bool isT = C::ENERGY && f == C::F_T;
      const bool isY = f >= C::F_Y && f < C::F_Y + C::NSPEC;
      const bool isP = !C::P1 && f == C::F_P;
      // row = sum_gp  w phi_i * G[q0]  +  w grad_phi_i[p] * G[q1 + p]; the last case is the mesh rows
      // (assemble_mesh residual, mm_fill_terms.c:421-428)
      const int q0 = isT ? C::G_RE : isY ? C::G_RY + (f - C::F_Y) : isP ? C::G_DIV : C::G_ZERO;
      const int q1 = isT ? C::G_RF : isY ? C::G_RFY + (f - C::F_Y) * DIM : isP ? C::G_PS : C::G_RD + (f - C::F_D) * DIM;
      for (int gp = gp0; gp < gp1; gp++) {
        const double wq = C::MMA ? s.w[gp] : 1.0;
        const double2 s01 = make_double2(wq * op_si<C>(op, gp, 0, i), wq * op_si<C>(op, gp, 1, i));
        const double2 s23 = make_double2(wq * op_si<C>(op, gp, 2, i), wq * op_si<C>(op, gp, 3, i));
        const double sg[3] = {s01.y, s23.x, s23.y};
        const double *G = s.GP[bo][gp];
        double t = s01.x * G[q0];
#pragma unroll
        for (int p = 0; p < DIM; p++) t += sg[p] * G[q1 + p];
        R += t;
      }
    }
    s.redR[c][r] = R;
  }
  cta_sync<C>();
  // ---- part B: residual rows and Dirichlet rows (put_dirichlet_in_matrix, bc_dirich.c:86-140)
  for (int r = tid; r < NROW; r += NT) {
    const bool prow = r >= NF * NN;
    const int f = prow ? 0 : r / NN;
    const int i = prow ? C::CEN : r - f * NN;
    const int gun = prow ? rec.gunP + (r - NF * NN) : rec.gun[f][i];
    const int flag = prow ? rec.flagP[r - NF * NN] : rec.flag[f][i];
    if (!(flag & 4)) continue;  // row of an external node (load_lec, mm_fill.c:5374)
    const bool first = (rec.node_first >> i) & 1u;
    const int dbc = flag & 3;
    if (dbc) {
      if (P.assemble_residual) slot_add(P, &P.resid[gun], dbc == 1 ? P.x[gun] - P.dbc_value[gun] : 0.0, first);
      if (P.assemble_jacobian) {
        long long dpos = gun;  // MSR: the diagonal lives in a[0..N)
        if (P.csr) {           // CSR: at its sorted position inside the row
          const int before = prow ? rec.po[C::CEN][C::CEN] + rec.poffP + (r - NF * NN)
                                  : rec.po[i][i] + rec.cs[i][f] - ((C::ENERGY && f == C::F_T) ? rec.pp[i][i] : 0);
          dpos = P.rowstart[gun] - P.msr0 + gun + before;
        }
        slot_add(P, &P.a[dpos], 1.0, first);
      }
      continue;
    }
    if (!P.assemble_residual) continue;
    if (!prow) {
      const bool known = f < DIM || (C::ENERGY && f == C::F_T) || (f >= C::F_Y && f < C::F_Y + C::NSPEC) ||
                         (!C::P1 && f == C::F_P) || (C::ALE && f >= C::F_D && f < C::F_D + DIM);
      if (!known) continue;
    }
    double R = s.redR[0][r];
#pragma unroll
    for (int c = 1; c < NPART; c++) R += s.redR[c][r];
    slot_add(P, &P.resid[gun], R, first);
  }
  // ---- part C (phase 7): P1 pressure coupling.  The centroid node belongs to this element only: these
  //      slots have a single writer.
  if constexpr (C::P1) if (P.assemble_jacobian) {
    const int poff = rec.poffP;
    for (int idx = tid; idx < NN * DIM * NP; idx += NT) {
      int i = idx / (DIM * NP), r = idx - i * DIM * NP;
      int a = r / NP, p = r - a * NP;
      double S = s.redS[0][a * NN + i][p];
#pragma unroll
      for (int c = 1; c < NPART; c++) S += s.redS[c][a * NN + i][p];
      // (row velocity a of node i, column pressure p of the centroid node)
      if (rec.rs[a][i] >= 0) {
        const int row = rec.gun[a][i], col = rec.gunP + p;
        const long long pos = rec.rs[a][i] + rec.po[i][C::CEN] + poff + p - ((!P.csr && col > row) ? 1 : 0);
        slot_add(P, &P.a[pos], P.etm_mom[3] * S, true);
      }
    }
    // (row pressure p of the centroid node, column velocity a of node i): a runs fastest over the lanes, so that the
    // three entries of a column node leave in one store request
    for (int idx = tid; idx < NN * DIM * NP; idx += NT) {
      const int p = idx / (NN * DIM), r = idx - p * NN * DIM;
      const int i = r / DIM, a = r - i * DIM;
      if (rec.rsP[p] < 0) continue;
      double S = s.redS[0][a * NN + i][p];
#pragma unroll
      for (int c = 1; c < NPART; c++) S += s.redS[c][a * NN + i][p];
      const int row = rec.gunP + p, col = rec.gun[a][i];
      const long long pos = rec.rsP[p] + rec.po[C::CEN][i] + rec.cs[i][a] - ((!P.csr && col > row) ? 1 : 0);
      slot_add(P, &P.a[pos], P.etm_cont[0] * S, true);
    }
    if constexpr (C::ALE) {
      // J_c_d (mm_fill_continuity.c:1004-1148): d(div v)/d d_bj + div v d|J|/d d_bj, with
      // d(grad_phi_k[q])/d d_bj = -grad_phi_j[q] grad_phi_k[b] and d|J|/d d_bj = |J| grad_phi_j[b]
      for (int idx = tid; idx < NN * DIM * NP; idx += NT) {
        const int j = idx / (DIM * NP), r = idx - j * DIM * NP;
        const int b = r / NP, p = r - b * NP;
        if (rec.rsP[p] < 0) continue;
        double acc = 0.0;
        for (int gp = 0; gp < NGP; gp++) {
          const double2 ja = op.SJa[gp][j], jb = op.SJb[gp][j];
          const double gj[3] = {ja.y, jb.x, jb.y};
          double ddiv = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; q++) ddiv -= gj[q] * s.F[gp][C::F_V + q][1 + b];
          acc += s.w[gp] * t_psi[gp * (DIM + 1) + p] * (P.etm_cont[0] * ddiv + s.GP[bo][gp][C::G_DIV] * gj[b]);
        }
        const int row = rec.gunP + p, col = rec.gun[C::F_D + b][j];
        const long long pos = rec.rsP[p] + rec.po[C::CEN][j] + rec.cs[j][C::F_D + b] - ((!P.csr && col > row) ? 1 : 0);
        slot_add(P, &P.a[pos], acc, true);
      }
    }
  }
}

// =====================================================================================
// phase 6: node-pair blocks.  Thread = (row tile of TI nodes, one column node j); the TI x 1 tile of
// NF x NF blocks is accumulated in registers over the Gauss points, then written to its matrix slots.
// =====================================================================================
template <class C>
struct Tile {
  double V[C::TI][C::NF][C::NF];
};

template <class C>
__device__ __forceinline__ void gauss_loop(const FillParams &P, const Smem<C> &s, int i0, int j, Tile<C> &out, int bo = 0) {
  constexpr int DIM = C::DIM, NGP = C::NGP, TI = C::TI;
  const Operands<C> &op = s.op[bo];
  const double tfac = P.transient ? (1.0 + 2.0 * P.theta) / P.delta_t : 0.0;
  const double rcp = P.rho * P.Cp;
  const double c_adv = -P.etm_mom[1] * P.rho, c_diff = -P.etm_mom[3] * P.mu, c_mass = -P.etm_mom[0] * P.rho * tfac;
  const double ce_adv = -P.etm_energy[1] * rcp, ce_diff = -P.etm_energy[3] * P.k,
               ce_mass = -P.etm_energy[0] * rcp * tfac;
  // D: the delta_ab part of J_m_v, sum_gp w phi_i (c_adv v.grad_phi_j + c_mass phi_j) + c_diff w grad_phi_i.grad_phi_j;
  // the energy configurations also need the three sums separately (J_e_T has other coefficients, J_m_T wants S3)
  double A[TI][DIM][DIM], D[TI], S1[TI], S2[TI], S3[TI], ET[TI][DIM];
#pragma unroll
  for (int ii = 0; ii < TI; ii++) {
    D[ii] = S1[ii] = S2[ii] = S3[ii] = 0.0;
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
    const double2 j01 = op.SJa[gp][j];
    const double2 j23 = op.SJb[gp][j];
    const double phi_j = j01.x;
    const double gj[3] = {j01.y, j23.x, j23.y};
    const double vgj = op.VG[gp][j];
    const double qj = c_adv * vgj + c_mass * phi_j;
    double gjs[DIM], GV[DIM][DIM], GT[DIM];
    const double *G = s.GP[bo][gp];
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      gjs[a] = c_diff * gj[a];
      if (C::ENERGY) GT[a] = G[C::G_GT + a];
#pragma unroll
      for (int b = 0; b < DIM; b++) GV[a][b] = G[C::G_GV + a * DIM + b];
    }
#pragma unroll
    for (int ii = 0; ii < TI; ii++) {
      const double2 i01 = make_double2(op_si<C>(op, gp, 0, i0 + ii), op_si<C>(op, gp, 1, i0 + ii));
      const double2 i23 = make_double2(op_si<C>(op, gp, 2, i0 + ii), op_si<C>(op, gp, 3, i0 + ii));
      const double wphi = i01.x;
      const double wg[3] = {i01.y, i23.x, i23.y};
      const double pp = wphi * phi_j;
      if (C::ENERGY) {
        S1[ii] += wphi * vgj;
        S3[ii] += pp;
#pragma unroll
        for (int p = 0; p < DIM; p++) S2[ii] += wg[p] * gj[p];
      } else {
        double d = D[ii] + wphi * qj;
#pragma unroll
        for (int p = 0; p < DIM; p++) d += wg[p] * gjs[p];
        D[ii] = d;
      }
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
  double dfdT[3] = {0.0, 0.0, 0.0};
  if (C::ENERGY && P.source_model != 0 && P.etm_mom[4] != 0.0) {
#pragma unroll
    for (int a = 0; a < DIM; a++) dfdT[a] = -P.g[a] * P.rho * P.beta * P.etm_mom[4];
  }
#pragma unroll
  for (int ii = 0; ii < TI; ii++) {
    const double dm = C::ENERGY ? c_adv * S1[ii] + c_diff * S2[ii] + c_mass * S3[ii] : D[ii];
#pragma unroll
    for (int a = 0; a < DIM; a++) {
#pragma unroll
      for (int b = 0; b < DIM; b++) out.V[ii][a][b] = A[ii][a][b] + (a == b ? dm : 0.0);
      if (C::ENERGY) {
        out.V[ii][a][C::F_T] = dfdT[a] * S3[ii];  // J_m_T (mm_std_models.c:337)
        out.V[ii][C::F_T][a] = ET[ii][a];
      }
    }
    if (C::ENERGY) out.V[ii][C::F_T][C::F_T] = ce_adv * S1[ii] + ce_diff * S2[ii] + ce_mass * S3[ii];  // J_e_T
  }
}

// Generic block accumulation: every active field pair, straight into the NF x NF tile.  Used for the
// equal-order (PSPG) and species configurations; the NS(+T) P1 configurations use gauss_loop above.
template <class C>
__device__ __forceinline__ void gauss_loop_general(const FillParams &P, const Smem<C> &s, int i0, int j, Tile<C> &out, int bo = 0) {
  constexpr int DIM = C::DIM, NGP = C::NGP, TI = C::TI, NF = C::NF, NN = C::NN;
  const Operands<C> &op = s.op[bo];
  const double tfac = P.transient ? (1.0 + 2.0 * P.theta) / P.delta_t : 0.0;
  const double rcp = P.rho * P.Cp;
  const double c_adv = -P.etm_mom[1] * P.rho, c_diff = -P.etm_mom[3] * P.mu, c_mass = -P.etm_mom[0] * P.rho * tfac;
  const double ce_adv = -P.etm_energy[1] * rcp, ce_diff = -P.etm_energy[3] * P.k,
               ce_mass = -P.etm_energy[0] * rcp * tfac;
  const double cs_adv = -P.etm_species[1], cs_mass = -P.etm_species[0] * tfac;
  const double tau = C::P1 ? 0.0 : s.tau;
  double dfdT[3] = {0.0, 0.0, 0.0};
  if (C::ENERGY && P.source_model != 0 && P.etm_mom[4] != 0.0) {
#pragma unroll
    for (int a = 0; a < DIM; a++) dfdT[a] = -P.g[a] * P.rho * P.beta;
  }
#pragma unroll
  for (int ii = 0; ii < TI; ii++)
#pragma unroll
    for (int r = 0; r < NF; r++)
#pragma unroll
      for (int c = 0; c < NF; c++) out.V[ii][r][c] = 0.0;
  const int ngp_run = (P.debug & 2) ? 1 : NGP;
#pragma unroll 1
  for (int gp = 0; gp < ngp_run; gp++) {
    const double2 j01 = op.SJa[gp][j], j23 = op.SJb[gp][j];
    const double phi_j = j01.x;
    const double gj[3] = {j01.y, j23.x, j23.y};
    const double vgj = op.VG[gp][j];
    const double *G = s.GP[bo][gp];
#pragma unroll
    for (int ii = 0; ii < TI; ii++) {
      const double2 i01 = make_double2(op.SI[gp][0][i0 + ii], op.SI[gp][1][i0 + ii]);
      const double2 i23 = make_double2(op.SI[gp][2][i0 + ii], op.SI[gp][3][i0 + ii]);
      const double wphi = i01.x;
      const double wg[3] = {i01.y, i23.x, i23.y};
      const double pp = wphi * phi_j;
      double gij = 0.0;
#pragma unroll
      for (int p = 0; p < DIM; p++) gij += wg[p] * gj[p];
      const double s1 = wphi * vgj;
      double (&V)[NF][NF] = out.V[ii];
      const double dm = c_adv * s1 + c_diff * gij + c_mass * pp;
#pragma unroll
      for (int a = 0; a < DIM; a++) {
#pragma unroll
        for (int b = 0; b < DIM; b++) V[a][b] += pp * G[C::G_GV + a * DIM + b] + c_diff * wg[b] * gj[a];
        V[a][a] += dm;
        if (C::ENERGY) {
          V[a][C::F_T] += P.etm_mom[4] * dfdT[a] * pp;              // J_m_T
          V[C::F_T][a] += pp * G[C::G_GT + a];                      // J_e_v
        }
#pragma unroll
        for (int w = 0; w < C::NSPEC; w++) V[C::F_Y + w][a] += pp * G[C::G_GY + w * DIM + a];  // J_s_v
        if (!C::P1) V[a][C::F_P] += P.etm_mom[3] * wg[a] * phi_j;  // J_m_P (mm_fill_momentum.c:2091-2104)
      }
      if (C::ENERGY) V[C::F_T][C::F_T] += ce_adv * s1 + ce_diff * gij + ce_mass * pp;  // J_e_T
#pragma unroll
      for (int w = 0; w < C::NSPEC; w++)
        V[C::F_Y + w][C::F_Y + w] += cs_adv * s1 - P.etm_species[3] * P.diffusivity[w] * gij + cs_mass * pp;
      if (!C::P1) {
        // continuity row of node i: div term + PSPG (mm_fill_continuity.c:686-756, d_pspg mm_fill_stabilization.c:1321-)
        double rmom = 0.0, rdf = 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          rmom += wg[a] * G[C::G_MOM + a];
          rdf += wg[a] * dfdT[a];
        }
        const double t1 = tau * P.rho * (P.etm_mom[0] * tfac * phi_j + P.etm_mom[1] * vgj);
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          double q = 0.0;
#pragma unroll
          for (int a = 0; a < DIM; a++) q += wg[a] * G[C::G_HB + a * DIM + b];
          V[C::F_P][b] += P.etm_cont[0] * wphi * gj[b] + phi_j * q + wg[b] * t1 + s.dtau[b] * rmom;
        }
        V[C::F_P][C::F_P] += tau * P.etm_mom[3] * gij;
        if (C::ENERGY) V[C::F_P][C::F_T] += -tau * P.etm_mom[4] * phi_j * rdf;
      }
      if (C::ALE) {
        // Mesh-sensitivity blocks.  Every term of a residual integrand  phi_i q0 + grad_phi_i . q1  depends on
        // d_bj through (i) |J| -> integrand x grad_phi_j[b], (ii) grad_phi_i -> -grad_phi_i[b] (grad_phi_j . q1),
        // (iii) the gradients inside q0/q1 -> -grad_phi_j[.] (d_b field), (iv) in transient runs the mesh
        // velocity in v - xdot (which the reference differentiates only when the mass term is on).
        // J_m_d mm_fill_momentum.c:2200-2442, J_e_d mm_fill_energy.c:758-925, J_s_d mm_fill_species.c:1103-1330,
        // J_d_d mm_fill_terms.c:529-576 with d(TT)/d d_bj from mesh_stress_tensor (mm_fill_solid.c:3288-3430).
        const double ed3 = P.etm_mesh[3];
        double Gd[DIM][DIM], gvu[DIM][DIM];  // grad_d[p][q] = d_p d_q ; gvu[a][b] = d_b v_a
#pragma unroll
        for (int p = 0; p < DIM; p++)
#pragma unroll
          for (int q = 0; q < DIM; q++) {
            Gd[p][q] = s.F[gp][C::F_D + q][1 + p];
            gvu[p][q] = s.F[gp][C::F_V + p][1 + q];
          }
        double iM[DIM], iG[DIM], MiG[DIM], igv[DIM], CMj[DIM];
#pragma unroll
        for (int c = 0; c < DIM; c++) {
          iG[c] = 0.0;
          igv[c] = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; q++) {
            iG[c] += wg[q] * Gd[q][c];
            igv[c] += wg[q] * gvu[q][c];
          }
        }
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          iM[b] = wg[b];
          CMj[b] = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; q++) {
            iM[b] -= wg[q] * Gd[b][q];
            CMj[b] += gj[q] * G[C::G_CM + q * DIM + b];
          }
        }
#pragma unroll
        for (int b = 0; b < DIM; b++) {
          MiG[b] = iG[b];
#pragma unroll
          for (int c = 0; c < DIM; c++) MiG[b] -= Gd[b][c] * iG[c];
        }
        const double mass_on_m = P.etm_mom[0] != 0.0 ? tfac : 0.0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
          // momentum a and mesh a integrands, and grad_phi_j . q1
          double Rm = wphi * G[C::G_RQ + a], Rd = 0.0, jm = 0.0, jd = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; q++) {
            Rm += wg[q] * G[C::G_RP + a * DIM + q];
            Rd += wg[q] * G[C::G_RD + a * DIM + q];
            jm += gj[q] * G[C::G_RP + a * DIM + q];
            jd += gj[q] * G[C::G_RD + a * DIM + q];
          }
#pragma unroll
          for (int b = 0; b < DIM; b++) {
            const double GVab = G[C::G_GV + a * DIM + b];
            V[a][C::F_D + b] += gj[b] * Rm - wg[b] * jm - wphi * vgj * GVab - c_diff * (gj[a] * igv[b] + gij * gvu[a][b]) -
                                mass_on_m * pp * GVab;
            // sum_q grad_phi_i[q] dE[a][q] of the NONLINEAR (Eulerian) strain
            const double idE = 0.5 * (gj[a] * iM[b] + gij * ((a == b ? 1.0 : 0.0) - Gd[b][a])) -
                               0.5 * (gj[a] * MiG[b] + gij * G[C::G_GM + a * DIM + b]);
            V[C::F_D + a][C::F_D + b] +=
                gj[b] * Rd - wg[b] * jd - (G[C::G_C1] * CMj[b] * wg[a] + ed3 * 2.0 * P.lame_mu * idE);
          }
        }
        if (C::ENERGY) {
          double Re = wphi * G[C::G_RE], je = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; q++) {
            Re += wg[q] * G[C::G_RF + q];
            je += gj[q] * G[C::G_RF + q];
          }
          const double mass_on = P.etm_energy[0] != 0.0 ? tfac : 0.0;
#pragma unroll
          for (int b = 0; b < DIM; b++)
            V[C::F_T][C::F_D + b] += gj[b] * Re - wg[b] * je - gij * G[C::G_RF + b] -
                                     (wphi * vgj + mass_on * pp) * G[C::G_GT + b];
        }
#pragma unroll
        for (int w = 0; w < C::NSPEC; w++) {
          double Ry = wphi * G[C::G_RY + w], jy = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; q++) {
            Ry += wg[q] * G[C::G_RFY + w * DIM + q];
            jy += gj[q] * G[C::G_RFY + w * DIM + q];
          }
          const double mass_on = P.etm_species[0] != 0.0 ? tfac : 0.0;
#pragma unroll
          for (int b = 0; b < DIM; b++)
            V[C::F_Y + w][C::F_D + b] += gj[b] * Ry - wg[b] * jy - gij * G[C::G_RFY + w * DIM + b] -
                                         (wphi * vgj + mass_on * pp) * G[C::G_GY + w * DIM + b];
        }
      }
    }
  }
  (void)NN;
}

// first-touch store of the NF consecutive entries of one (row, column node) run, as 16-byte vectors where possible
template <int NF>
__device__ __forceinline__ void store_run(double *dst, const double *v) {
  const bool al = (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0;
  if (NF == 3) {
    if (al) {
      *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]);
      dst[2] = v[2];
    } else {
      dst[0] = v[0];
      *reinterpret_cast<double2 *>(dst + 1) = make_double2(v[1], v[2]);
    }
  } else {
    if (al) {
      *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]);
    } else {
      dst[0] = v[0];
      *reinterpret_cast<double2 *>(dst + 1) = make_double2(v[1], v[2]);
      dst[3] = v[3];
    }
  }
}

// the same for a run of N doubles (equal-order / species / ALE blocks: N = 5 .. 10)
template <int N>
__device__ __forceinline__ void store_run_n(double *dst, const double *v) {
  if ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) *reinterpret_cast<double2 *>(dst + k) = make_double2(v[k], v[k + 1]);
    if (N & 1) dst[N - 1] = v[N - 1];
  } else {
    dst[0] = v[0];
#pragma unroll
    for (int k = 1; k + 1 < N; k += 2) *reinterpret_cast<double2 *>(dst + k) = make_double2(v[k], v[k + 1]);
    if (!(N & 1)) dst[N - 1] = v[N - 1];
  }
}

// write-out of one node-pair tile straight from registers: one thread per (row tile, column node j).  In a
// row (i, fr) the entries of column node j are contiguous; when the tile is the first writer of the pair and
// the fields of node j sit next to each other, the NF doubles go out as 16-byte + 8-byte stores (fewer LSU
// sector operations than NF scalar stores -- the LSU data path is the busiest unit of this kernel).
template <class C, int MODE>
__device__ __forceinline__ void write_tile_direct(const FillParams &P, const ElemRec<C> &s, int i, int j, const Tile<C> &t) {
  constexpr int NF = C::NF;
  const int rj = s.rank[j];
  bool packed = true;  // cs[j][f] == cs[j][0] + f: node j's fields occupy consecutive columns
#pragma unroll
  for (int f = 1; f < NF; f++) packed = packed && (s.cs[j][f] == s.cs[j][0] + f);
#pragma unroll
  for (int ii = 0; ii < C::TI; ii++, i++) {
    const bool first = (s.first[i] >> j) & 1u;
    const int ri = s.rank[i];
#pragma unroll
    for (int fr = 0; fr < NF; fr++) {
      const long long rstart = s.rs[fr][i];
      if (rstart < 0) continue;
      const bool rowT = C::ENERGY && fr == C::F_T;
      const int row = s.gun[fr][i];
      double *arow = P.a + rstart;
      if (MODE == 2 && (NF == 3 || (NF == 4 && C::P1)) && first && packed && (rj != ri || P.csr)) {
        double *dst = arow + s.po[i][j] + s.cs[j][0] - ((!P.csr && rj > ri) ? 1 : 0) - ((C::ENERGY && rowT) ? s.pp[i][j] : 0);
        store_run<NF>(dst, &t.V[ii][fr][0]);
        continue;
      }
      if (MODE == 2 && NF > 4 && first && packed && (rj != ri || P.csr)) {
        // general field sets: the entries of (row, column node) are one run; energy rows of the equal-order
        // configurations end before the pressure column (the last field of the node)
        double *dst = arow + s.po[i][j] + s.cs[j][0] - ((!P.csr && rj > ri) ? 1 : 0) - ((C::ENERGY && rowT) ? s.pp[i][j] : 0);
        if (rowT && !C::P1)
          store_run_n<NF - 1>(dst, &t.V[ii][fr][0]);
        else
          store_run_n<NF>(dst, &t.V[ii][fr][0]);
        continue;
      }
#pragma unroll
      for (int fc = 0; fc < NF; fc++) {
        if (rowT && !C::P1 && fc == C::F_P) continue;
        int off = s.po[i][j] + s.cs[j][fc];
        if (C::ENERGY && rowT) off -= s.pp[i][j];
        double *dst;
        if (P.csr)
          dst = arow + off;
        else if (rj != ri)
          dst = arow + off - (rj > ri ? 1 : 0);
        else
          dst = (fc == fr) ? P.a + row : arow + off - (fc > fr ? 1 : 0);
        slot_add_m<MODE>(dst, t.V[ii][fr][fc], first);
      }
    }
  }
}


// the same for ONE node pair (i, j): V[fr][fc] of the tensor-core path
template <class C, int MODE>
__device__ __forceinline__ void write_pair(const FillParams &P, const ElemRec<C> &s, int i, int j, const double (&V)[C::NF][C::NF]) {
  constexpr int NF = C::NF;
  const int rj = s.rank[j], ri = s.rank[i];
  bool packed = true;
#pragma unroll
  for (int f = 1; f < NF; f++) packed = packed && (s.cs[j][f] == s.cs[j][0] + f);
  const bool first = (s.first[i] >> j) & 1u;
#pragma unroll
  for (int fr = 0; fr < NF; fr++) {
    const long long rstart = s.rs[fr][i];
    if (rstart < 0) continue;
    const bool rowT = C::ENERGY && fr == C::F_T;
    const int row = s.gun[fr][i];
    double *arow = P.a + rstart;
    if (MODE == 2 && (NF == 3 || (NF == 4 && C::P1)) && first && packed && (rj != ri || P.csr)) {
      double *dst = arow + s.po[i][j] + s.cs[j][0] - ((!P.csr && rj > ri) ? 1 : 0) - ((C::ENERGY && rowT) ? s.pp[i][j] : 0);
      store_run<NF>(dst, &V[fr][0]);
      continue;
    }
#pragma unroll
    for (int fc = 0; fc < NF; fc++) {
      int off = s.po[i][j] + s.cs[j][fc];
      if (C::ENERGY && rowT) off -= s.pp[i][j];
      double *dst;
      if (P.csr)
        dst = arow + off;
      else if (rj != ri)
        dst = arow + off - (rj > ri ? 1 : 0);
      else
        dst = (fc == fr) ? P.a + row : arow + off - (fc > fr ? 1 : 0);
      slot_add_m<MODE>(dst, V[fr][fc], first);
    }
  }
}


// =====================================================================================
// phase 6 on the FP64 tensor cores (hex27 Q2/P1 NS and NS+T).  For every component pair (a, b) the node-pair
// block of J_m_v (mm_fill_momentum.c:1629-1712) is a 27 x 27 x (2 * 27) matrix product over the Gauss points:
//     V_ab[i][j] = sum_gp  (w phi_i) . (phi_j c_adv d_b v_a  [+ delta_ab q_j])  +  (w grad_phi_i[b]) . (c_diff grad_phi_j[a])
// with q_j = c_adv v.grad_phi_j + c_mass phi_j.  The delta_ab diffusion term is the trace over the three diagonal
// products, kept in separate accumulators (KD) so that it costs no extra multiply.  A warp owns 8 x 8 node
// blocks (27 -> 32 in both directions, K 27 -> 28): 18 DMMA per block and K step (23 with energy: J_m_T, J_e_v,
// J_e_T from the products S1 = w phi_i . v.grad_phi_j, S3 = w phi_i . phi_j and w phi_i . phi_j ce_adv d_a T).
// Shared-memory traffic per DMMA is one 8-byte fragment load per ~1 tensor instruction (a lane's fragment feeds
// 8 FMAs) against one load per 2.9 FMAs of the scalar 3 x 1 register tile this replaces.
// =====================================================================================
template <class C, int MODE>
__device__ __forceinline__ void gauss_blocks_mma(const FillParams &P, const Smem<C> &s, const ElemRec<C> &rec, int warp, int lane,
                                                 long long *t_mma = nullptr) {
  static_assert(C::MMA && C::DIM == 3, "tensor-core path: hex27 Q2/P1");
  constexpr int NF = C::NF;
  const Operands<C> &op = s.op[0];
  const double tfac = P.transient ? (1.0 + 2.0 * P.theta) / P.delta_t : 0.0;
  const double rcp = P.rho * P.Cp;
  const double c_adv = -P.etm_mom[1] * P.rho, c_diff = -P.etm_mom[3] * P.mu, c_mass = -P.etm_mom[0] * P.rho * tfac;
  const double ce_adv = -P.etm_energy[1] * rcp, ce_diff = -P.etm_energy[3] * P.k, ce_mass = -P.etm_energy[0] * rcp * tfac;
  const int r = lane >> 2, kq = lane & 3;
  const int nks = (P.debug & 2) ? 1 : 7;
  // 27 = 3 * 8 + 3: the nine full 8 x 8 node blocks run on the tensor cores (warps 0 .. NWARP-3); padding the
  // remaining 3 nodes to a fourth block row / column would spend 16/9 of that work on the pipe the kernel is bound by,
  // so the 153 node pairs with a node >= 24 are done by the last two warps as scalar 3 x 1 register tiles
  // (gauss_loop): 51 tiles, 20-23 FMAs per pair and Gauss point, about as long as two tensor-core blocks
  constexpr bool REM = !(C::VAR & 1);
  constexpr int NMW = REM ? C::NWARP - 2 : C::NWARP, NBLK = REM ? 9 : 16, NBS = REM ? 3 : 4;
  if (REM && warp >= NMW) {
    // 153 pairs with a node >= 24 over 64 threads: (i >= 24, any j) first, then (i < 24, j >= 24)
#pragma unroll 1
    for (int t = (warp - NMW) * 32 + lane; t < 81 + 72; t += 64) {
      const int i = t < 81 ? 24 + t / 27 : (t - 81) / 3, j = t < 81 ? t % 27 : 24 + (t - 81) % 3;
      double A[3][3], D = 0.0, S1 = 0.0, S2 = 0.0, S3 = 0.0, ET[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) A[a][b] = 0.0;
      const int ngp_run = (P.debug & 2) ? 1 : 27;
#pragma unroll 1
      for (int gp = 0; gp < ngp_run; gp++) {
        const double *G = s.GP[0][gp];
        const double wv = s.w[gp];
        const double phi_i = op.SJ[gp][i], gi[3] = {op.SJ[gp][28 + i], op.SJ[gp][56 + i], op.SJ[gp][84 + i]};
        const double phj = wv * op.SJ[gp][j], vgj = wv * op.SJ[gp][112 + j];
        const double gj[3] = {wv * op.SJ[gp][28 + j], wv * op.SJ[gp][56 + j], wv * op.SJ[gp][84 + j]};
        const double pp = phi_i * phj;
        double gij = 0.0;
#pragma unroll
        for (int p = 0; p < 3; p++) gij += gi[p] * gj[p];
        if (C::ENERGY) {
          S1 += phi_i * vgj;
          S2 += gij;
          S3 += pp;
        } else {
          D += phi_i * (c_adv * vgj + c_mass * phj) + c_diff * gij;
        }
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const double gjs = c_diff * gj[a];
#pragma unroll
          for (int b = 0; b < 3; b++) A[a][b] += pp * G[C::G_GV + a * 3 + b] + gi[b] * gjs;
          if (C::ENERGY) ET[a] += pp * G[C::G_GT + a];
        }
      }
      if (P.debug & 1) continue;
      double dfdT[3] = {0.0, 0.0, 0.0};
      if (C::ENERGY && P.source_model != 0 && P.etm_mom[4] != 0.0) {
#pragma unroll
        for (int a = 0; a < 3; a++) dfdT[a] = -P.g[a] * P.rho * P.beta * P.etm_mom[4];
      }
      const double dm = C::ENERGY ? c_adv * S1 + c_diff * S2 + c_mass * S3 : D;
      double V[NF][NF];
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) V[a][b] = A[a][b] + (a == b ? dm : 0.0);
        if constexpr (C::ENERGY) {
          V[a][C::F_T] = dfdT[a] * S3;
          V[C::F_T][a] = ET[a];
        }
      }
      if constexpr (C::ENERGY) V[C::F_T][C::F_T] = ce_adv * S1 + ce_diff * S2 + ce_mass * S3;
      write_pair<C, MODE>(P, rec, i, j, V);
    }
    return;
  }
#pragma unroll 1
  for (int blk = warp; blk < NBLK; blk += NMW) {
    const int I0 = (blk / NBS) * 8, J0 = (blk % NBS) * 8;
    const int ia = min(I0 + r, 27), jb = min(J0 + r, 27);  // rows / columns >= 27 read the zero padding
#ifdef GOMA_PROFILE_PHASES
    const long long t_blk0 = clock64();
#endif
    double T[3][3][2], KD[3][2], S1[2] = {0.0, 0.0}, S3[2] = {0.0, 0.0}, ET[3][2];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      KD[a][0] = KD[a][1] = 0.0;
      ET[a][0] = ET[a][1] = 0.0;
#pragma unroll
      for (int b = 0; b < 3; b++) T[a][b][0] = T[a][b][1] = 0.0;
    }
#pragma unroll 1
    for (int ks = 0; ks < nks; ks++) {
      const int gp = 4 * ks + kq;
      const double *si = &op.SJ[gp][ia], *sj = &op.SJ[gp][jb];
      const double *G = s.GP[0][gp];
      const double wv = s.w[gp];  // the weight of this lane's Gauss point rides on the B fragments (w[27] = 0)
      const double aphi = si[0];
      const double ag[3] = {si[28], si[56], si[84]};
      const double bphi = wv * sj[0];
      const double bg[3] = {wv * sj[28], wv * sj[56], wv * sj[84]};
      const double bvg = wv * sj[112];
      const double bq = C::ENERGY ? 0.0 : c_adv * bvg + c_mass * bphi;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double bgs = c_diff * bg[a];
#pragma unroll
        for (int b = 0; b < 3; b++) {
          const double gv = G[C::G_GV + a * 3 + b];
          const double bt = (a == b && !C::ENERGY) ? fma(bphi, gv, bq) : bphi * gv;
          dmma884(T[a][b], aphi, bt);
          if (a != b) dmma884(T[a][b], ag[b], bgs);
        }
        dmma884(KD[a], ag[a], bg[a]);
        if (C::ENERGY) dmma884(ET[a], aphi, bphi * G[C::G_GT + a]);  // J_e_v (mm_fill_energy.c:640)
      }
      if (C::ENERGY) {
        dmma884(S1, aphi, bvg);
        dmma884(S3, aphi, bphi);
      }
    }
#ifdef GOMA_PROFILE_PHASES
    if (t_mma) *t_mma += clock64() - t_blk0;
#endif
    if (P.debug & 1) continue;
    const int i = I0 + r;
    if (i >= 27) continue;
    double dfdT[3] = {0.0, 0.0, 0.0};
    if (C::ENERGY && P.source_model != 0 && P.etm_mom[4] != 0.0) {
#pragma unroll
      for (int a = 0; a < 3; a++) dfdT[a] = -P.g[a] * P.rho * P.beta * P.etm_mom[4];
    }
    // write-out: the node's fields are consecutive unknowns (checked at init), so the NF entries of (row, column
    // node j) are contiguous; everything that depends on the row node only is fetched once per block
    long long rsv[NF];
#pragma unroll
    for (int fr = 0; fr < NF; fr++) rsv[fr] = rec.rs[fr][i];
    const int ri = rec.rank[i];
    const unsigned fi = rec.first[i];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int j = J0 + 2 * kq + c;
      if (j >= 27) continue;
      const double S2 = (KD[0][c] + KD[1][c]) + KD[2][c];  // sum_gp w grad_phi_i . grad_phi_j
      const double dm = C::ENERGY ? c_adv * S1[c] + c_diff * S2 + c_mass * S3[c] : c_diff * S2;
      double V[NF][NF];
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) V[a][b] = (a == b) ? T[a][a][c] + (c_diff * KD[a][c] + dm) : T[a][b][c];
        if constexpr (C::ENERGY) {
          V[a][C::F_T] = dfdT[a] * S3[c];  // J_m_T (mm_std_models.c:337)
          V[C::F_T][a] = ET[a][c];
        }
      }
      if constexpr (C::ENERGY) V[C::F_T][C::F_T] = ce_adv * S1[c] + ce_diff * S2 + ce_mass * S3[c];  // J_e_T
      const int rj = rec.rank[j];
      if (rj == ri) {  // the diagonal node pair: the diagonal entries live in a[0..N), the rest shifts
        write_pair<C, MODE>(P, rec, i, j, V);
        continue;
      }
      const bool first = (fi >> j) & 1u;
      const int off = rec.po[i][j] + rec.cs[j][0] - ((!P.csr && rj > ri) ? 1 : 0);
      const int offT = C::ENERGY ? off - rec.pp[i][j] : off;  // energy rows carry no pressure columns
      // one branch on `first` per pair (not per entry): a warp whose lanes disagree runs each side once
      if (MODE == 2 && first) {
        // first touch: plain stores, as 16-byte vectors where the run allows (the store path -- L1 tag stage, one
        // crossbar packet per instruction and sector -- is what this phase is bound by: fewer, wider requests)
#pragma unroll
        for (int fr = 0; fr < NF; fr++) {
          if (rsv[fr] < 0) continue;
          double *dst = P.a + rsv[fr] + ((C::ENERGY && fr == C::F_T) ? offT : off);
#ifdef GOMA_PROFILE_PHASES
          if (g_store_debug & 8) continue;
#endif
          store_run<NF>(dst, &V[fr][0]);
        }
      } else {
#pragma unroll
        for (int fr = 0; fr < NF; fr++) {
          if (rsv[fr] < 0) continue;
          double *dst = P.a + rsv[fr] + ((C::ENERGY && fr == C::F_T) ? offT : off);
#pragma unroll
          for (int fc = 0; fc < NF; fc++) slot_add_m<MODE>(dst + fc, V[fr][fc], false);
        }
      }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::TPE, C::MINB) fill_kernel(const __grid_constant__ FillParams P) {
  constexpr int NN = C::NN, NF = C::NF, TI = C::TI, NT = C::TPE;
  constexpr unsigned REC_BYTES = (unsigned)sizeof(ElemRec<C>);
  static_assert(REC_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<C> &s = *reinterpret_cast<Smem<C> *>(smem_raw);
  const int tid = threadIdx.x;
  int ee = P.elem_begin + blockIdx.x;

  // ---- prologue: quadrature/basis tables and the record of the first element, one TMA bulk copy each
  if (tid == 0) {
    mbar_init(&s.mbar, 1);
    mbar_init(&s.mbar_rec[0], 1);
    mbar_init(&s.mbar_rec[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!P.transient)
    for (int idx = tid; idx < 2 * NF * NN; idx += NT) (&s.Udot[0][0][0])[idx] = 0.0;
  if constexpr (C::MMA) {  // the padding row / column of the tensor-core operand tables is zero for the whole launch
    for (int idx = tid; idx < (int)(sizeof(s.op) / 8); idx += NT) reinterpret_cast<double *>(&s.op)[idx] = 0.0;
    for (int idx = tid; idx < (int)(sizeof(s.GP) / 8); idx += NT) (&s.GP[0][0][0])[idx] = 0.0;
    if (tid == 0) s.w[C::NGK - 1] = 0.0;
    if (tid < 27) s.lat[tid] = HEX27_LATTICE[tid];
    __syncthreads();
    for (int idx = tid; idx < C::NGP * NN; idx += NT) {  // phi_j at the Gauss points: the same for every element
      const int gp = idx / NN, j = idx - gp * NN;
      s.op[0].SJ[gp][j] = P.tables[C::T_PHI + idx];
    }
  }
  __syncthreads();
  if (ee >= P.elem_end) return;
  if (tid == 0) {
    mbar_expect_tx(&s.mbar, C::TBL_PAD * 8);
    tma_bulk_g2s(s.tbl, P.tables, C::TBL_PAD * 8, &s.mbar);
    const int elem = P.elem_list ? P.elem_list[ee] : ee;
    mbar_expect_tx(&s.mbar_rec[0], REC_BYTES);
    tma_bulk_g2s(&s.rec[0], P.erec + (size_t)elem * REC_BYTES, REC_BYTES, &s.mbar_rec[0]);
  }
  mbar_wait(&s.mbar_rec[0], 0);
  gather_state<C>(P, s, 0, tid);
  cp_async_wait_all();
  mbar_wait(&s.mbar, 0);
  __syncthreads();

  long long t_build = 0, t_rows = 0, t_loop = 0;
  long long stamps_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long *stamps = (P.prof && tid == 0) ? stamps_ : nullptr;
  int count = 0;
  int pending = 0;  // (thread 0) the index the counter handed out for the element after the next
  if (tid == 0 && P.work) pending = P.elem_begin + P.static_rounds * (int)gridDim.x + atomicAdd(P.work, 1);
#pragma unroll 1
  for (; ee < P.elem_end; count++) {
    const int buf = count & 1;
    // Which element comes next: the first static_rounds * gridDim.x elements of the launch are taken by block index,
    // the rest from a counter, so that a CTA on a slower SM takes fewer elements and the launch ends without a tail (elements of one
    // launch share no slot: the result does not depend on who assembles which).  The record of that element starts
    // its way into the other buffer (free since the barrier that closed the previous iteration).
    if (tid == 0) {
      int nx = ee + (int)gridDim.x;              // the first static_rounds elements of a CTA go by block index ...
      if (P.work && count + 1 >= P.static_rounds) {  // ... the later ones come from the counter, asked for one element
        nx = pending;                                // ahead: the round trip of the atomic is over when it is needed
        pending = P.elem_begin + P.static_rounds * (int)gridDim.x + atomicAdd(P.work, 1);
      }
      s.next_ee = nx;
      if (nx < P.elem_end) {
        const int nxt = P.elem_list ? P.elem_list[nx] : nx;
        mbar_expect_tx(&s.mbar_rec[buf ^ 1], REC_BYTES);
        tma_bulk_g2s(&s.rec[buf ^ 1], P.erec + (size_t)nxt * REC_BYTES, REC_BYTES, &s.mbar_rec[buf ^ 1]);
      }
    }
    long long c0 = GOMA_CLOCK();
    build_element<C>(P, s, buf, 0, tid, stamps);
    long long c1 = GOMA_CLOCK();
    element_rows<C>(P, s, buf, 0, tid);
    long long c2 = GOMA_CLOCK();
    const int ee_next = s.next_ee;  // (written before the barriers of the phases above)
    const bool has_next = ee_next < P.elem_end;
    // ... and, once it has landed, the gather of the next element's unknowns runs under the Gauss loop
    if (has_next) {
      mbar_wait(&s.mbar_rec[buf ^ 1], ((count + 1) >> 1) & 1);
      gather_state<C>(P, s, buf ^ 1, tid);
    }
    if (P.assemble_jacobian) {
      const ElemRec<C> &rec = s.rec[buf];
      if constexpr (C::MMA) {
        long long *tm = stamps ? &stamps_[6] : nullptr;  // profiling build: cycles of the DMMA part (warp 0)
        if (P.scatter_mode == 2)
          gauss_blocks_mma<C, 2>(P, s, rec, tid >> 5, tid & 31, tm);
        else if (P.scatter_mode == 0)
          gauss_blocks_mma<C, 0>(P, s, rec, tid >> 5, tid & 31, tm);
        else
          gauss_blocks_mma<C, 1>(P, s, rec, tid >> 5, tid & 31, tm);
      } else {
#pragma unroll 1
        for (int t = tid; t < C::NTILE; t += NT) {
          Tile<C> tile;
          const int it = t / NN, j = t - it * NN, i = it * TI;
          if (C::GENERAL)
            gauss_loop_general<C>(P, s, i, j, tile);
          else
            gauss_loop<C>(P, s, i, j, tile);
          if (P.debug & 1) continue;
          if (P.scatter_mode == 2)
            write_tile_direct<C, 2>(P, rec, i, j, tile);
          else if (P.scatter_mode == 0)
            write_tile_direct<C, 0>(P, rec, i, j, tile);
          else
            write_tile_direct<C, 1>(P, rec, i, j, tile);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    ee = ee_next;
    t_loop += GOMA_CLOCK() - c2;
    t_build += c1 - c0;
    t_rows += c2 - c1;
  }
  if (P.prof && tid == 0 && count) {
    long long *o = P.prof + blockIdx.x * 8;
    o[0] = t_build; o[1] = t_rows; o[2] = t_loop; o[3] = 0; o[6] = count;
    long long *o2 = P.prof + (4096 + blockIdx.x) * 8;
    for (int k = 0; k < 7; k++) o2[k] = stamps_[k];
  }
}

// =====================================================================================
// Warp-specialised variant for the Q2/P1 Navier-Stokes block: one CTA per SM.  C::TPE "builder" threads run
// the set-up phases and the residual rows of element k while C::NMUL "multiplier" threads run the Gauss loop
// (TI x 1 register tiles) and the write-out of element k-1.  Hand-off through mbarriers:
//   full[b]  : operands of buffer b complete (builders -> multipliers)
//   empty[b] : multipliers are done with buffer b and with the record of that element (-> builders)
// Records travel by TMA two elements ahead (ring of four), the state gather by cp.async one element ahead.
// =====================================================================================
template <class C>
__global__ void __launch_bounds__(C::TPE + C::NMUL, 1) fill_kernel_ws(const __grid_constant__ FillParams P) {
  constexpr int NN = C::NN, NF = C::NF, TI = C::TI, NB = C::TPE, NR = C::NRECB;
  constexpr unsigned REC_BYTES = (unsigned)sizeof(ElemRec<C>);
  static_assert(C::WS && NR == 4, "ring of four records");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<C> &s = *reinterpret_cast<Smem<C> *>(smem_raw);
  const int tid = threadIdx.x;
  const int ee0 = P.elem_begin + blockIdx.x, stride = gridDim.x;
  const int nel = ee0 < P.elem_end ? (P.elem_end - ee0 + stride - 1) / stride : 0;
  if (tid == 0) {
    mbar_init(&s.mbar, 1);
    for (int k = 0; k < NR; k++) mbar_init(&s.mbar_rec[k], 1);
    for (int k = 0; k < 2; k++) {
      mbar_init(&s.full[k], 1);
      mbar_init(&s.empty[k], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!P.transient)
    for (int idx = tid; idx < NR * NF * NN; idx += NB + C::NMUL) (&s.Udot[0][0][0])[idx] = 0.0;
  __syncthreads();
  if (nel == 0) return;

  if (tid < NB) {
    // ------------------------------------------------------------------ builders
    auto elem_of = [&](int k) { const int ee = ee0 + k * stride; return P.elem_list ? P.elem_list[ee] : ee; };
    auto fetch_record = [&](int k) {  // one thread: TMA bulk copy of the record of the CTA's k-th element
      mbar_expect_tx(&s.mbar_rec[k % NR], REC_BYTES);
      tma_bulk_g2s(&s.rec[k % NR], P.erec + (size_t)elem_of(k) * REC_BYTES, REC_BYTES, &s.mbar_rec[k % NR]);
    };
    if (tid == 0) {
      mbar_expect_tx(&s.mbar, C::TBL_PAD * 8);
      tma_bulk_g2s(s.tbl, P.tables, C::TBL_PAD * 8, &s.mbar);
      fetch_record(0);
      if (nel > 1) fetch_record(1);
    }
    mbar_wait(&s.mbar_rec[0], 0);
    gather_state<C>(P, s, 0, tid, NB);
    cp_async_wait_all();
    mbar_wait(&s.mbar, 0);
    cta_sync<C>();
    long long tw = 0, tb = 0, tr = 0, tt = 0;
#pragma unroll 1
    for (int k = 0; k < nel; k++) {
      const int br = k % NR, bo = k & 1;
      long long c0 = GOMA_CLOCK();
      // buffer bo and the ring slots of elements <= k-2 are free once the multipliers have finished element k-2
      if (k >= 2) mbar_wait(&s.empty[bo], ((k - 2) >> 1) & 1);
      if (k + 2 < nel && tid == 0) fetch_record(k + 2);
      if (k + 1 < nel) {  // record k+1 was requested one iteration ago: its state gather runs under this build
        mbar_wait(&s.mbar_rec[(k + 1) % NR], ((k + 1) / NR) & 1);
        gather_state<C>(P, s, (k + 1) % NR, tid, NB);
      }
      long long c1 = GOMA_CLOCK();
      build_element<C>(P, s, br, bo, tid, nullptr);
      long long c2 = GOMA_CLOCK();
      element_rows<C>(P, s, br, bo, tid);
      long long c3 = GOMA_CLOCK();
      cp_async_wait_all();
      cta_sync<C>();  // operands of k and the unknowns of k+1 are in shared memory, written by all builders
      if (tid == 0) mbar_arrive(&s.full[bo]);
      tw += c1 - c0; tb += c2 - c1; tr += c3 - c2; tt += GOMA_CLOCK() - c3;
    }
    if (P.prof && tid == 0) {
      long long *o = P.prof + blockIdx.x * 8;
      o[0] = tw; o[1] = tb; o[2] = tr; o[3] = tt; o[6] = nel;
    }
  } else {
    // ------------------------------------------------------------------ multipliers
    const int mt = tid - NB;
    const int it = mt / NN, j = mt - it * NN, i0 = it * TI;
    long long mw = 0, ml = 0, mo = 0;
#pragma unroll 1
    for (int k = 0; k < nel; k++) {
      const int br = k % NR, bo = k & 1;
      long long c0 = GOMA_CLOCK();
      mbar_wait(&s.full[bo], (k >> 1) & 1);
      long long c1 = GOMA_CLOCK(), c2 = c1;
      if (P.assemble_jacobian && mt < C::NTILE) {
        Tile<C> tile;
        gauss_loop<C>(P, s, i0, j, tile, bo);
        c2 = GOMA_CLOCK();
        if (!(P.debug & 1)) {
          if (P.scatter_mode == 2)
            write_tile_direct<C, 2>(P, s.rec[br], i0, j, tile);
          else if (P.scatter_mode == 0)
            write_tile_direct<C, 0>(P, s.rec[br], i0, j, tile);
          else
            write_tile_direct<C, 1>(P, s.rec[br], i0, j, tile);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(C::NMUL) : "memory");  // every multiplier is done reading buffer bo
      if (mt == 0) mbar_arrive(&s.empty[bo]);
      mw += c1 - c0; ml += c2 - c1; mo += GOMA_CLOCK() - c2;
    }
    if (P.prof && mt == 0) {
      long long *o = P.prof + (4096 + blockIdx.x) * 8;
      o[0] = mw; o[1] = ml; o[2] = mo; o[3] = 0; o[4] = 0; o[5] = 0;
    }
  }
}

}  // namespace goma_b200
