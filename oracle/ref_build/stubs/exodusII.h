/* stand-in for exodusII.h: constants and prototypes only (every ex_* aborts if
 * called; the oracle driver builds the Exo_DB in memory instead of reading a file) */
#ifndef GOMA_B200_ORACLE_EXODUS_STUB_H
#define GOMA_B200_ORACLE_EXODUS_STUB_H
#include <stddef.h>
#include <stdint.h>
#define MAX_STR_LENGTH 32L
#define MAX_NAME_LENGTH 32L
#define MAX_LINE_LENGTH 80L
#define MAX_ERR_LENGTH 512
#define EX_READ 0x0002
#define EX_WRITE 0x0001
#define EX_CLOBBER 0x0008
#define EX_NOCLOBBER 0x0004
#define EX_VERBOSE 1
#define EX_DEBUG 2
#define EX_ABORT 4
#define EX_NOERR 0
#define EX_WARN 1
#define EX_FATAL (-1)
#define EX_MSG (-1000)
#define EX_LASTERR (-1003)
#define EX_NOCLASSIC 0x0020
#define EX_LARGE_MODEL 0x0010
#define EX_NETCDF4 0x0040
#define EX_ALL_INT64_API 0x1C000
typedef int64_t ex_entity_id;
typedef enum {
  EX_NODAL = 14, EX_NODE_BLOCK = 14, EX_NODE_SET = 2, EX_EDGE_BLOCK = 6, EX_EDGE_SET = 7,
  EX_FACE_BLOCK = 8, EX_FACE_SET = 9, EX_ELEM_BLOCK = 1, EX_ELEM_SET = 10, EX_SIDE_SET = 3,
  EX_ELEM_MAP = 4, EX_NODE_MAP = 5, EX_EDGE_MAP = 11, EX_FACE_MAP = 12, EX_GLOBAL = 13,
  EX_COORDINATE = 15, EX_INVALID = -1
} ex_entity_type;
typedef enum {
  EX_INQ_FILE_TYPE = 1, EX_INQ_API_VERS = 2, EX_INQ_DB_VERS = 3, EX_INQ_TITLE = 4, EX_INQ_DIM = 5,
  EX_INQ_NODES = 6, EX_INQ_ELEM = 7, EX_INQ_ELEM_BLK = 8, EX_INQ_NODE_SETS = 9, EX_INQ_NS_NODE_LEN = 10,
  EX_INQ_SIDE_SETS = 11, EX_INQ_SS_NODE_LEN = 12, EX_INQ_SS_ELEM_LEN = 13, EX_INQ_QA = 14, EX_INQ_INFO = 15,
  EX_INQ_TIME = 16, EX_INQ_EB_PROP = 17, EX_INQ_NS_PROP = 18, EX_INQ_SS_PROP = 19, EX_INQ_NS_DF_LEN = 20,
  EX_INQ_SS_DF_LEN = 21, EX_INQ_LIB_VERS = 22, EX_INQ_EM_PROP = 23, EX_INQ_NM_PROP = 24,
  EX_INQ_ELEM_MAP = 25, EX_INQ_NODE_MAP = 26, EX_INQ_DB_MAX_USED_NAME_LENGTH = 50, EX_INQ_INVALID = -1
} ex_inquiry;
typedef struct ex_block {
  int64_t id; ex_entity_type type; char topology[33]; int64_t num_entry; int64_t num_nodes_per_entry;
  int64_t num_edges_per_entry; int64_t num_faces_per_entry; int64_t num_attribute;
} ex_block;
typedef struct ex_set {
  int64_t id; ex_entity_type type; int64_t num_entry; int64_t num_distribution_factor;
  void *entry_list; void *extra_list; void *distribution_factor_list;
} ex_set;
typedef struct ex_set_specs {
  void *sets_ids, *num_entries_per_set, *num_dist_per_set, *sets_entry_index, *sets_dist_index,
      *sets_entry_list, *sets_extra_list, *sets_dist_fact;
} ex_set_specs;
#endif
