// Quadrature rules and reference-element basis tables, evaluated once on the host.
//
// Restates (tensor-product form) what the reference evaluates per Gauss point, per node,
// per element through a switch statement:
//   src/el_elm_info.c:1615-1830  find_stu   (point order: s fastest, then t, then u; +a,0,-a)
//   src/el_elm_info.c:3469-3640  Gq_weight  (5/9, 8/9 products; 1.0 for the 2-point rule)
//   src/rf_shape.c:185,361,698,1105  shape() for BILINEAR_QUAD, BIQUAD_QUAD, TRILINEAR_HEX,
//                                    TRIQUAD_HEX in Exodus/PATRAN node order
//   src/mm_fill_util.c:3802-3855  P1 pressure basis {1, s, t[, u]}
#pragma once
#include <vector>

#include "../../include/goma_gpu_fill.h"

namespace goma_b200 {

struct ElemTables {
  int dim = 0, nn = 0, ngp = 0;
  std::vector<double> xi;    // [ngp][3]
  std::vector<double> wt;    // [ngp]
  std::vector<double> phi;   // [ngp][nn]
  std::vector<double> dphi;  // [ngp][nn][dim]
  std::vector<double> psi;   // [ngp][dim+1]  P1 basis
  std::vector<double> l1d;   // [2][3][3]: 1-D Lagrange factors L[point][node] and dL[point][node] at the 1-D Gauss points
  // basis at the element centroid xi = 0 (PSPG block 1.5, mm_fill.c:754-787)
  std::vector<double> phi0, dphi0;
};

inline void lagrange1d(int order, double s, double *L, double *dL) {
  if (order == 1) {
    L[0] = 0.5 * (1.0 - s);
    L[1] = 0.5 * (1.0 + s);
    dL[0] = -0.5;
    dL[1] = 0.5;
  } else {
    L[0] = -0.5 * s * (1.0 - s);
    L[1] = (1.0 - s * s);
    L[2] = 0.5 * s * (1.0 + s);
    dL[0] = -0.5 * (1.0 - 2.0 * s);
    dL[1] = -2.0 * s;
    dL[2] = 0.5 * (1.0 + 2.0 * s);
  }
}

// lattice offsets of the local nodes (same tables as goma_b200/mesh.py)
inline const int (*node_lattice(int elem_type))[3] {
  static const int q4[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
  static const int q9[9][3] = {{0, 0, 0}, {2, 0, 0}, {2, 2, 0}, {0, 2, 0}, {1, 0, 0},
                               {2, 1, 0}, {1, 2, 0}, {0, 1, 0}, {1, 1, 0}};
  static const int h8[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                               {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  static const int h27[27][3] = {
      {0, 0, 0}, {2, 0, 0}, {2, 2, 0}, {0, 2, 0}, {0, 0, 2}, {2, 0, 2}, {2, 2, 2}, {0, 2, 2}, {1, 0, 0},
      {2, 1, 0}, {1, 2, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 1}, {2, 2, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 2},
      {1, 2, 2}, {0, 1, 2}, {1, 1, 1}, {1, 1, 0}, {1, 1, 2}, {0, 1, 1}, {2, 1, 1}, {1, 0, 1}, {1, 2, 1}};
  switch (elem_type) {
    case GOMA_GPU_QUAD4: return q4;
    case GOMA_GPU_QUAD9: return q9;
    case GOMA_GPU_HEX8: return h8;
    default: return h27;
  }
}

inline void eval_basis(int elem_type, const double xi[3], double *phi, double *dphi) {
  const int dim = (elem_type == GOMA_GPU_QUAD4 || elem_type == GOMA_GPU_QUAD9) ? 2 : 3;
  const int order = (elem_type == GOMA_GPU_QUAD4 || elem_type == GOMA_GPU_HEX8) ? 1 : 2;
  const int nn = elem_type;
  const int(*lat)[3] = node_lattice(elem_type);
  double L[3][3], dL[3][3];
  for (int d = 0; d < 3; d++) {
    L[d][0] = 1.0;
    dL[d][0] = 0.0;
  }
  for (int d = 0; d < dim; d++) lagrange1d(order, xi[d], L[d], dL[d]);
  for (int i = 0; i < nn; i++) {
    const int *o = lat[i];
    double l0 = L[0][o[0]], l1 = L[1][o[1]], l2 = dim == 3 ? L[2][o[2]] : 1.0;
    phi[i] = l0 * l1 * l2;
    dphi[i * dim + 0] = dL[0][o[0]] * l1 * l2;
    dphi[i * dim + 1] = l0 * dL[1][o[1]] * l2;
    if (dim == 3) dphi[i * dim + 2] = l0 * l1 * dL[2][o[2]];
  }
}

inline ElemTables make_tables(int elem_type) {
  ElemTables t;
  t.dim = (elem_type == GOMA_GPU_QUAD4 || elem_type == GOMA_GPU_QUAD9) ? 2 : 3;
  const int order = (elem_type == GOMA_GPU_QUAD4 || elem_type == GOMA_GPU_HEX8) ? 1 : 2;
  t.nn = elem_type;
  const int n1 = order + 1;
  t.ngp = t.dim == 2 ? n1 * n1 : n1 * n1 * n1;
  static const double F1 = 0.57735026918962584208, F2 = 0.77459666924148340428;
  static const double W1 = 0.55555555555555555556, W2 = 0.88888888888888888888;
  const double pts2[2] = {F1, -F1}, wts2[2] = {1.0, 1.0};
  const double pts3[3] = {F2, 0.0, -F2}, wts3[3] = {W1, W2, W1};
  const double *pts = order == 1 ? pts2 : pts3, *wts = order == 1 ? wts2 : wts3;
  t.xi.assign(t.ngp * 3, 0.0);
  t.wt.resize(t.ngp);
  t.phi.resize(t.ngp * t.nn);
  t.dphi.resize(t.ngp * t.nn * t.dim);
  t.psi.resize(t.ngp * (t.dim + 1));
  for (int g = 0; g < t.ngp; g++) {
    int is = g % n1, it = (g / n1) % n1, iu = g / (n1 * n1);
    double *xi = &t.xi[g * 3];
    xi[0] = pts[is];
    xi[1] = pts[it];
    xi[2] = t.dim == 3 ? pts[iu] : 0.0;
    // Gq_weight multiplies weight_s * weight_t (* weight_u) in that order
    t.wt[g] = t.dim == 3 ? wts[is] * wts[it] * wts[iu] : wts[is] * wts[it];
    eval_basis(elem_type, xi, &t.phi[g * t.nn], &t.dphi[g * t.nn * t.dim]);
    t.psi[g * (t.dim + 1) + 0] = 1.0;
    for (int d = 0; d < t.dim; d++) t.psi[g * (t.dim + 1) + 1 + d] = xi[d];
  }
  t.l1d.assign(18, 0.0);
  for (int pt = 0; pt < n1; pt++) {
    double L[3] = {0, 0, 0}, dL[3] = {0, 0, 0};
    lagrange1d(order, pts[pt], L, dL);
    for (int n = 0; n < n1; n++) {
      t.l1d[pt * 3 + n] = L[n];
      t.l1d[9 + pt * 3 + n] = dL[n];
    }
  }
  t.phi0.resize(t.nn);
  t.dphi0.resize(t.nn * t.dim);
  const double zero[3] = {0, 0, 0};
  eval_basis(elem_type, zero, t.phi0.data(), t.dphi0.data());
  return t;
}

}  // namespace goma_b200
