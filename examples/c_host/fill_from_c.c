/*
 * A C host driving the GPU fill through the C ABI only (include/goma_gpu_fill.h) -- what the shim at the top
 * of Goma's matrix_fill_full does, stripped of Goma: read a problem snapshot, call
 * goma_gpu_fill_init / goma_gpu_fill, write ams->val and resid_vector.
 *
 *   gcc -std=c99 -I include examples/c_host/fill_from_c.c -L goma_b200 -lgoma_gpu_fill -Wl,-rpath,$PWD/goma_b200 -o fill_from_c
 *   ./fill_from_c problem.bin out.bin
 *
 * problem.bin (written by tests/test_c_host.py): the scalar members of struct goma_gpu_problem as the struct
 * itself (pointer members ignored), followed by the arrays in declaration order.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "goma_gpu_fill.h"

static void *slurp(FILE *f, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (bytes && fread(p, 1, bytes, f) != bytes) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
  return p;
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s problem.bin out.bin\n", argv[0]);
    return 2;
  }
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  struct goma_gpu_problem p;
  if (fread(&p, sizeof(p), 1, f) != 1) return 2;
  const int npe = p.elem_type;
  p.elem_connect = (const int *)slurp(f, sizeof(int) * (size_t)p.num_elems * npe);
  for (int d = 0; d < 3; d++) p.coord[d] = d < p.dim ? (const double *)slurp(f, sizeof(double) * p.num_nodes) : NULL;
  p.first_unknown = (const int *)slurp(f, sizeof(int) * p.num_nodes);
  p.node_kind = (const unsigned char *)slurp(f, p.num_nodes);
  p.ija = NULL; /* let the library derive the MSR graph */
  p.dbc_flag = (const unsigned char *)slurp(f, p.num_unknowns);
  p.dbc_value = (const double *)slurp(f, sizeof(double) * p.num_unknowns);
  double *x = (double *)slurp(f, sizeof(double) * p.num_unknowns);
  fclose(f);

  goma_gpu_ctx *ctx = NULL;
  if (goma_gpu_fill_init(&p, 0, &ctx) != 0) {
    fprintf(stderr, "goma_gpu_fill_init: %s\n", goma_gpu_last_error());
    return 1;
  }
  long long nnz_plus = 0;
  goma_gpu_fill_get_msr(ctx, &nnz_plus);
  double *a = (double *)calloc((size_t)nnz_plus + 1, sizeof(double));
  double *resid = (double *)calloc(p.num_unknowns, sizeof(double));
  int flags[3] = {0, 0, 0};
  /* int matrix_fill_full(ams, x, resid_vector, x_old, x_older, xdot, xdot_old, x_update, &delta_t, &theta, ...) */
  int err = goma_gpu_fill(ctx, x, NULL, NULL, NULL, NULL, 0.0, 0.0, 0.0, 0.0, 0.0, 1, 1, a, resid, flags);
  if (err < -1) {
    fprintf(stderr, "goma_gpu_fill: %s\n", goma_gpu_last_error());
    return 1;
  }
  FILE *o = fopen(argv[2], "wb");
  fwrite(&err, sizeof(int), 1, o);
  fwrite(flags, sizeof(int), 3, o);
  fwrite(&nnz_plus, sizeof(long long), 1, o);
  fwrite(a, sizeof(double), (size_t)nnz_plus + 1, o);
  fwrite(resid, sizeof(double), p.num_unknowns, o);
  fclose(o);
  goma_gpu_fill_destroy(ctx);
  printf("matrix_fill_full (GPU) returned %d; %lld MSR values, %d unknowns\n", err, nnz_plus + 1, p.num_unknowns);
  return 0;
}
