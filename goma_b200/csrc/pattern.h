// Host-side construction of the sparsity skeleton the scatter kernel writes through.
//
// Restates, 64-bit clean and multi-threaded, what the reference does serially in
//   src/exo_conn.c:204-376          build_node_node (sorted node-node lists, plus
//                                   face-neighbour centroids for centroid nodes :315-347)
//   src/mm_fill_util.c:3229-3445    find_MSR_problem_graph (rows x Inter_Mask columns)
// and replaces the per-entry `in_list` search of load_lec (src/mm_fill.c:5461) by a
// per-element table of column offsets computed once.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/goma_gpu_fill.h"

namespace goma_b200 {

struct Pattern {
  int num_nodes = 0, num_unknowns = 0, npe = 0;
  std::vector<int64_t> nn_ptr;   // [num_nodes+1]
  std::vector<int> nn_list;      // sorted neighbours of each node
  std::vector<uint16_t> cum_full;  // per nn_list entry: unknowns carried by earlier neighbours
  std::vector<uint16_t> cum_p;     // ... of which pressure unknowns
  std::vector<int64_t> rowstart;   // [num_unknowns+1] == MSR ija[0..N] (64-bit)
  int64_t nnz_plus = 0;            // == ija[N]
  // per element, per local node pair (i,j): column offset (all variables) of node j's first
  // unknown inside a row of node i, and the number of pressure unknowns before it
  std::vector<uint16_t> pair_full, pair_p;  // [num_elems*npe*npe]
  bool need_pair_p = false;
  // Element colouring (no two elements of a colour share a node) and first-touch masks for the
  // write-once scatter: colours are processed in increasing order, so the element of lowest colour
  // among those sharing a node pair writes the slot with a plain store and the later ones add to it.
  std::vector<int> colour_order;      // elements sorted by colour
  std::vector<int> colour_begin;      // [ncolours+1] into colour_order
  std::vector<uint32_t> pair_first;   // [num_elems*npe] bit j of word (e,i): e is the first writer of pair (i,j)
  std::vector<uint32_t> node_first;   // [num_elems]     bit i: e is the first element touching node i's rows
};

// returns "" on success, else an error message
std::string build_pattern(const goma_gpu_problem &p, Pattern &out, int num_threads);

// number of pressure unknowns of a node kind
int kind_num_pressure(const goma_gpu_problem &p, int kind);

// column ids of the MSR graph (ija[N+1 .. nnz_plus)), for export / validation at test sizes
void emit_msr_columns(const goma_gpu_problem &p, const Pattern &pat, int *ija_out);

}  // namespace goma_b200
