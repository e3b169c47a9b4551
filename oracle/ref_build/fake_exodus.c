/* In-memory stand-in for the handful of EXODUS II read calls the reference's
 * rd_exo() makes (rd_exo.c:202-728 in /root/reference/src).  The oracle driver
 * registers a structured mesh here, then the reference's own read_mesh_exoII()
 * runs unmodified on top of it.  TEST INFRASTRUCTURE ONLY. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "exodusII.h"

struct fake_mesh {
  int dim, num_nodes, num_elems, npe;
  char elem_type[33];
  const double *x, *y, *z;
  const int *conn; /* 1-based, num_elems*npe */
  int num_ns;
  const int *ns_ids, *ns_ptr, *ns_nodes; /* 1-based node ids */
  int num_eb;          /* element blocks: block b (id b+1) = eb_count[b] consecutive elements */
  const int *eb_count;
  int num_ss;          /* side sets: ss_ptr[i]..ss_ptr[i+1] index ss_elem (1-based element) / ss_side (1-based EXODUS side) */
  const int *ss_ids, *ss_ptr, *ss_elem, *ss_side;
} g_fake;

void fake_exodus_set_side_sets(int num_ss, const int *ids, const int *ptr, const int *elem, const int *side) {
  g_fake.num_ss = num_ss; g_fake.ss_ids = ids; g_fake.ss_ptr = ptr; g_fake.ss_elem = elem; g_fake.ss_side = side;
}
/* nodes of one element side in EXODUS II order (corners counter-clockwise seen from outside, then the mid-side /
 * mid-face nodes), as local 0-based node numbers; returns the count */
static int side_nodes(int side, int *ln) {
  const int s = side - 1;
  if (g_fake.dim == 2) {
    ln[0] = s; ln[1] = (s + 1) % 4;
    if (g_fake.npe == 9) { ln[2] = 4 + s; return 3; }
    return 2;
  }
  static const int hexc[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
  static const int hexm[6][5] = {{8, 13, 16, 12, 25}, {9, 14, 17, 13, 24}, {10, 15, 18, 14, 26},
                                 {12, 19, 15, 11, 23}, {11, 10, 9, 8, 21}, {16, 17, 18, 19, 22}};
  for (int k = 0; k < 4; k++) ln[k] = hexc[s][k];
  if (g_fake.npe == 27) { for (int k = 0; k < 5; k++) ln[4 + k] = hexm[s][k]; return 9; }
  return 4;
}
static int ss_index(int ss_id) {
  for (int i = 0; i < g_fake.num_ss; i++) if (g_fake.ss_ids[i] == ss_id) return i;
  return -1;
}
int ex_get_side_set_node_list_len(int id, int ss_id, int *len) {
  (void)id;
  int ln[9], i = ss_index(ss_id);
  *len = i < 0 ? 0 : (g_fake.ss_ptr[i + 1] - g_fake.ss_ptr[i]) * side_nodes(1, ln);
  return 0;
}
int ex_get_side_set_node_list(int id, int ss_id, int *cnt, int *list) {
  (void)id;
  int i = ss_index(ss_id), pos = 0;
  if (i < 0) return -1;
  for (int k = g_fake.ss_ptr[i]; k < g_fake.ss_ptr[i + 1]; k++) {
    int ln[9], n = side_nodes(g_fake.ss_side[k], ln);
    const int *c = g_fake.conn + (size_t)(g_fake.ss_elem[k] - 1) * g_fake.npe;
    cnt[k - g_fake.ss_ptr[i]] = n;
    for (int q = 0; q < n; q++) list[pos++] = c[ln[q]];
  }
  return 0;
}

void fake_exodus_set_blocks(int num_eb, const int *counts) {
  g_fake.num_eb = num_eb;
  g_fake.eb_count = counts;
}
static int eb_first(int blk_id) { /* first element of the block with id blk_id (1-based ids) */
  int first = 0;
  for (int b = 0; b + 1 < blk_id && b < g_fake.num_eb; b++) first += g_fake.eb_count[b];
  return first;
}
static int eb_size(int blk_id) { return g_fake.num_eb ? g_fake.eb_count[blk_id - 1] : g_fake.num_elems; }

void fake_exodus_set_mesh(int dim, int num_nodes, const double *x, const double *y, const double *z,
                          int num_elems, int npe, const char *elem_type, const int *conn,
                          int num_ns, const int *ns_ids, const int *ns_ptr, const int *ns_nodes) {
  g_fake.dim = dim; g_fake.num_nodes = num_nodes; g_fake.x = x; g_fake.y = y; g_fake.z = z;
  g_fake.num_elems = num_elems; g_fake.npe = npe; g_fake.conn = conn;
  strncpy(g_fake.elem_type, elem_type, 32);
  g_fake.num_ns = num_ns; g_fake.ns_ids = ns_ids; g_fake.ns_ptr = ns_ptr; g_fake.ns_nodes = ns_nodes;
}

int ex_open_int(const char *path, int mode, int *comp_ws, int *io_ws, float *version, int run_version) {
  (void)path; (void)mode; (void)run_version;
  *comp_ws = 8; *io_ws = 8; *version = 8.03f;
  return 1;
}
int ex_open(const char *path, int mode, int *comp_ws, int *io_ws, float *version) {
  return ex_open_int(path, mode, comp_ws, io_ws, version, 0);
}
int ex_close(int id) { (void)id; return 0; }
int ex_opts(int o) { (void)o; return 0; }
int ex_get_init(int id, char *title, int *num_dim, int *num_nodes, int *num_elems, int *num_eb,
                int *num_ns, int *num_ss) {
  (void)id;
  strcpy(title, "goma_b200 oracle in-memory mesh");
  *num_dim = g_fake.dim; *num_nodes = g_fake.num_nodes; *num_elems = g_fake.num_elems;
  *num_eb = g_fake.num_eb ? g_fake.num_eb : 1; *num_ns = g_fake.num_ns; *num_ss = g_fake.num_ss;
  return 0;
}
int ex_inquire(int id, int what, int *ri, float *rf, char *rc) {
  (void)id; (void)rc;
  *ri = 0; *rf = 0.f;
  switch (what) {
  case EX_INQ_API_VERS: case EX_INQ_DB_VERS: *rf = 8.03f; break;
  case EX_INQ_NS_NODE_LEN: *ri = g_fake.num_ns ? g_fake.ns_ptr[g_fake.num_ns] : 0; break;
  case EX_INQ_SS_ELEM_LEN: *ri = g_fake.num_ss ? g_fake.ss_ptr[g_fake.num_ss] : 0; break;
  case EX_INQ_SS_NODE_LEN: { int ln[9]; *ri = g_fake.num_ss ? g_fake.ss_ptr[g_fake.num_ss] * side_nodes(1, ln) : 0; break; }
  case EX_INQ_SS_DF_LEN: *ri = 0; break;
  default: break; /* no QA, info, dist-factors, side sets, properties, time planes */
  }
  return 0;
}
int ex_get_coord(int id, double *x, double *y, double *z) {
  (void)id;
  size_t n = (size_t)g_fake.num_nodes * sizeof(double);
  if (x && g_fake.x) memcpy(x, g_fake.x, n);
  if (y && g_fake.y) memcpy(y, g_fake.y, n);
  if (z && g_fake.z) memcpy(z, g_fake.z, n);
  return 0;
}
int ex_get_coord_names(int id, char **names) {
  (void)id;
  const char *nm[3] = {"x", "y", "z"};
  for (int i = 0; i < g_fake.dim; i++) strcpy(names[i], nm[i]);
  return 0;
}
int ex_get_ids(int id, int type, int *ids) {
  (void)id;
  if (type == EX_ELEM_BLOCK) for (int b = 0; b < (g_fake.num_eb ? g_fake.num_eb : 1); b++) ids[b] = b + 1;
  else if (type == EX_NODE_SET) for (int i = 0; i < g_fake.num_ns; i++) ids[i] = g_fake.ns_ids[i];
  else if (type == EX_SIDE_SET) for (int i = 0; i < g_fake.num_ss; i++) ids[i] = g_fake.ss_ids[i];
  return 0;
}
int ex_get_block(int id, int type, int blk, char *etype, int *nel, int *npe, int *nedge, int *nface, int *nattr) {
  (void)id; (void)type; (void)nedge; (void)nface;
  strcpy(etype, g_fake.elem_type);
  *nel = eb_size(blk); *npe = g_fake.npe; *nattr = 0;
  return 0;
}
int ex_get_conn(int id, int type, int blk, int *conn, int *e, int *f) {
  (void)id; (void)type; (void)e; (void)f;
  memcpy(conn, g_fake.conn + (size_t)eb_first(blk) * g_fake.npe, (size_t)eb_size(blk) * g_fake.npe * sizeof(int));
  return 0;
}
int ex_get_concat_sets(int id, int type, ex_set_specs *s) {
  (void)id;
  if (type == EX_SIDE_SET) {
    int *ids = s->sets_ids, *cnt = s->num_entries_per_set, *ndf = s->num_dist_per_set;
    int *idx = s->sets_entry_index, *dfi = s->sets_dist_index, *lst = s->sets_entry_list, *ext = s->sets_extra_list;
    for (int i = 0; i < g_fake.num_ss; i++) {
      ids[i] = g_fake.ss_ids[i];
      cnt[i] = g_fake.ss_ptr[i + 1] - g_fake.ss_ptr[i];
      ndf[i] = 0; idx[i] = g_fake.ss_ptr[i]; dfi[i] = 0;
    }
    if (g_fake.num_ss) {
      memcpy(lst, g_fake.ss_elem, (size_t)g_fake.ss_ptr[g_fake.num_ss] * sizeof(int));
      memcpy(ext, g_fake.ss_side, (size_t)g_fake.ss_ptr[g_fake.num_ss] * sizeof(int));
    }
    return 0;
  }
  if (type != EX_NODE_SET) return 0;
  int *ids = s->sets_ids, *cnt = s->num_entries_per_set, *ndf = s->num_dist_per_set;
  int *idx = s->sets_entry_index, *dfi = s->sets_dist_index, *lst = s->sets_entry_list;
  for (int i = 0; i < g_fake.num_ns; i++) {
    ids[i] = g_fake.ns_ids[i];
    cnt[i] = g_fake.ns_ptr[i + 1] - g_fake.ns_ptr[i];
    ndf[i] = 0; idx[i] = g_fake.ns_ptr[i]; dfi[i] = 0;
  }
  if (g_fake.num_ns) memcpy(lst, g_fake.ns_nodes, (size_t)g_fake.ns_ptr[g_fake.num_ns] * sizeof(int));
  return 0;
}
int ex_get_variable_param(int id, int type, int *n) { (void)id; (void)type; *n = 0; return 0; }
