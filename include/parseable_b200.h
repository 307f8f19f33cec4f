/*
 * parseable_b200.h — C ABI of the B200-native columnar query hot path for Parseable.
 *
 * This is the drop-in boundary of SURVEY.md §8(b).  Each entry point names the
 * reference interface it replaces (paths relative to /root/reference):
 *
 *   pq_init / pq_shutdown     QUERY_SESSION / QUERY_RUNTIME singletons        src/query/mod.rs:86-98
 *   pq_query_open             StandardTableProvider::scan +
 *                             create_parquet_physical_plan (+ the FilterExec /
 *                             AggregateExec DataFusion stacks on top)         src/query/stream_schema_provider.rs:114-189, 526-659
 *   pq_query_next             SendableRecordBatchStream::poll_next driven by
 *                             collect_partitioned / execute_stream_partitioned src/query/mod.rs:287, 310-334
 *   pq_query_stream           the same stream as ONE Arrow C stream object:
 *                             execute_stream_partitioned's merged
 *                             SendableRecordBatchStream (arrow-rs imports it
 *                             with ArrowArrayStreamReader)                     src/query/mod.rs:310-343
 *   pq_query_metrics          get_total_bytes_scanned ("bytes_scanned")        src/query/mod.rs:437-452
 *   pq_last_error             ExecuteError / DataFusionError::External         src/query/mod.rs:904-917
 *   pq_query_close            dropping the stream (cancellation)               src/query/mod.rs:300-340
 *   pq_table_*                the hot tier (local Parquet cache), here in HBM  src/hottier.rs:1541,1609
 *   pq_comm_*                 Partial -> RepartitionExec(Hash) -> Final merge,
 *                             here one NCCL all-reduce of partial tables       (DataFusion AggregateExec)
 *
 * Plain C: pointers and sizes only.  Results leave through the Arrow C Data
 * Interface (ArrowArray / ArrowSchema below, ABI-identical to arrow/c/abi.h).
 * Inputs are borrowed for the duration of the call.  Every function is
 * re-entrant and thread-safe; a PqQuery may be driven from any thread but by one
 * thread at a time.  There is no CPU fallback: without a CUDA device every
 * compute entry point returns PQ_ERR_CUDA.
 */
#ifndef PARSEABLE_B200_H
#define PARSEABLE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif
/* ---- Arrow C Stream Interface (https://arrow.apache.org/docs/format/CStreamInterface.html) ---- */
#ifndef ARROW_C_STREAM_INTERFACE
#define ARROW_C_STREAM_INTERFACE
struct ArrowArrayStream {
  int (*get_schema)(struct ArrowArrayStream*, struct ArrowSchema* out);
  int (*get_next)(struct ArrowArrayStream*, struct ArrowArray* out); /* out->release == NULL: end of stream */
  const char* (*get_last_error)(struct ArrowArrayStream*);
  void (*release)(struct ArrowArrayStream*);
  void* private_data;
};
#endif

/* ---- status codes (SURVEY §8b "Error convention") ---- */
enum {
  PQ_OK = 0,
  PQ_END_OF_STREAM = 1,
  PQ_ERR_INVALID_ARG = -1,
  PQ_ERR_UNSUPPORTED = -2, /* plan/encoding not handled: the shim raises an error, it never falls back */
  PQ_ERR_IO = -3,
  PQ_ERR_CORRUPT = -4,
  PQ_ERR_CUDA = -5, /* CUDA or NCCL */
  PQ_ERR_OOM = -6
};

/* ---- literal / column value types ---- */
typedef enum {
  PQ_T_NULL = 0,
  PQ_T_BOOL = 1,
  PQ_T_I64 = 2,   /* Int64 */
  PQ_T_F64 = 3,   /* Float64 */
  PQ_T_UTF8 = 4,  /* Utf8 */
  PQ_T_TS_MS = 5  /* Timestamp(Millisecond, None) */
} PqType;

/* ---- predicate: flat postfix program (SURVEY §8a "Predicate vocabulary") ---- */
typedef enum {
  PQ_OP_CMP = 1,      /* push  col <cmp> literal          (NULL input -> NULL)        */
  PQ_OP_IS_NULL = 2,  /* push  col IS NULL                                            */
  PQ_OP_IS_NOT_NULL = 3,
  PQ_OP_LIKE = 4,     /* push  col LIKE pattern  ESCAPE '\'  (flags: NOT, case-insens) */
  PQ_OP_AND = 5,      /* pop 2, push Kleene AND                                       */
  PQ_OP_OR = 6,       /* pop 2, push Kleene OR                                        */
  PQ_OP_NOT = 7,      /* pop 1, push Kleene NOT                                       */
  PQ_OP_CONST = 8     /* push literal TRUE/FALSE/NULL (lit.type BOOL or NULL)         */
} PqOpKind;

typedef enum { PQ_EQ = 0, PQ_NE = 1, PQ_LT = 2, PQ_LE = 3, PQ_GT = 4, PQ_GE = 5 } PqCmp;

#define PQ_LIKE_NEGATED 1u
#define PQ_LIKE_CASE_INSENSITIVE 2u

typedef struct {
  int32_t type; /* PqType */
  int32_t _pad;
  int64_t i64;  /* BOOL (0/1), I64, TS_MS */
  double f64;   /* F64 */
  const char* str; /* UTF8 / LIKE pattern, not NUL-terminated */
  uint64_t str_len;
} PqLiteral;

typedef struct {
  int32_t kind;  /* PqOpKind */
  int32_t col;   /* index into PqQueryDesc.columns for leaf ops */
  int32_t cmp;   /* PqCmp for PQ_OP_CMP */
  uint32_t flags;
  PqLiteral lit;
} PqPredOp;

/* ---- aggregates (alert_enums.rs:216-223: Avg, Count, CountDistinct, Min, Max, Sum) ---- */
typedef enum {
  PQ_AGG_COUNT_STAR = 0,
  PQ_AGG_COUNT = 1,
  PQ_AGG_SUM = 2,
  PQ_AGG_MIN = 3,
  PQ_AGG_MAX = 4,
  PQ_AGG_AVG = 5
} PqAggFn;

typedef struct {
  int32_t fn;  /* PqAggFn */
  int32_t col; /* index into PqQueryDesc.columns; ignored for COUNT_STAR */
} PqAgg;

/* ---- computed GROUP BY keys: the counts / histogram API groups by DATE_BIN(<width>, p_timestamp, origin)
 *      (src/query/mod.rs:623-680) ---- */
typedef enum { PQ_KEY_COLUMN = 0, PQ_KEY_DATE_BIN = 1 } PqKeyKind;
typedef struct {
  int32_t kind;       /* PqKeyKind */
  int32_t _pad;
  int64_t width_ms;   /* DATE_BIN: stride in milliseconds (> 0) */
  int64_t origin_ms;  /* DATE_BIN: origin, milliseconds since the epoch (the reference passes 1970-01-01) */
} PqKeyExpr;

/* ---- inputs ---- */
typedef struct {
  const char* path;   /* file to read, or NULL when buf is given */
  const uint8_t* buf; /* whole Parquet file image in host memory, or NULL */
  uint64_t size;      /* bytes of buf (ignored for path) */
} PqFile;

typedef struct {
  const char* name; /* matched against the Parquet schema BY NAME (streams.rs:1024-1037) */
  int32_t type;     /* PqType expected by the plan; a missing column reads as all-NULL */
  int32_t _pad;
} PqColumn;

typedef struct PqTable PqTable; /* HBM-resident encoded column chunks */
typedef struct PqQuery PqQuery;

typedef struct {
  /* scan inputs: either a resident table, or a file list (host buffers / paths) */
  const PqTable* table;
  const PqFile* files;
  uint32_t n_files;

  /* every column the plan references; all indices below point into this array */
  const PqColumn* columns;
  uint32_t n_columns;

  /* output projection for non-aggregate queries (TableProvider::scan `projection`) */
  const int32_t* projection;
  uint32_t n_projection;

  /* WHERE: postfix program; n_pred == 0 means no filter */
  const PqPredOp* pred;
  uint32_t n_pred;

  /* GROUP BY + aggregates; n_aggs == 0 means a filter/projection scan */
  const int32_t* group_by;
  uint32_t n_group_by;
  const PqAgg* aggs;
  uint32_t n_aggs;

  int64_t limit;       /* < 0: none (TableProvider::scan `limit`) */
  uint32_t batch_size; /* rows per output batch; 0 -> 20000 (stream_schema_provider.rs:160) */

  /* row-group sharding for multi-GPU: this process scans row groups g with
   * g % shard_count == shard_index (mirrors partitioned_files round-robin, :351-364) */
  uint32_t shard_index;
  uint32_t shard_count; /* 0 or 1: no sharding */
  uint32_t flags;

  /* NULL, or n_group_by entries: how group_by[k] becomes a key (plain column | DATE_BIN of a Timestamp / Int64 column);
   * a DATE_BIN key comes back as a Timestamp(ms) column named date_bin(<column>) holding the bin start */
  const PqKeyExpr* group_exprs;
} PqQueryDesc;

#define PQ_QUERY_COUNT_ONLY 1u    /* filter scan: only rows_selected is wanted, emit no batches */
#define PQ_QUERY_ALLREDUCE 2u     /* aggregate: all-reduce partial tables over the pq_comm communicator */
#define PQ_QUERY_EMIT_ROW_IDS 4u  /* filter scan: append a UInt64 `__row_id` column (global row ordinal) */

typedef struct {
  uint64_t bytes_scanned;   /* compressed bytes of the column chunks read (plan metric "bytes_scanned") */
  uint64_t rows_scanned;    /* rows in the row groups that survived pruning */
  uint64_t rows_selected;   /* rows passing the predicate */
  uint64_t row_groups_total;
  uint64_t row_groups_pruned;
  uint64_t algorithmic_bytes; /* uncompressed encoded bytes of the pages read + bitmap bytes written */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  uint64_t kernel_launches;
  double device_ms;         /* CUDA-event time of the device work of this query */
  double scan_kernel_ms;    /* CUDA-event time of the fused scan kernel alone */
  uint64_t groups;          /* output groups (aggregate queries) */
  double host_ms;           /* wall time of pq_query_open (planning + uploads + device work + result copy) */
  double upload_ms;         /* of which: footer parse, page walk and H2D of the column chunks (file-list queries) */
  double allreduce_ms;      /* CUDA-event time of the NCCL all-reduce of the partial tables (PQ_QUERY_ALLREDUCE) */
} PqMetrics;

/* ---- lifecycle ---- */
int pq_init(const int* device_ids, int n); /* n == 0: current device / device 0 */
void pq_shutdown(void);
const char* pq_version(void);
int pq_device_count(void);

/* ---- HBM-resident table (hot tier) ---- */
int pq_table_open(const PqFile* files, uint32_t n_files, const char* const* columns, uint32_t n_columns,
                  uint32_t shard_index, uint32_t shard_count, PqTable** out);
uint64_t pq_table_rows(const PqTable*);
uint64_t pq_table_device_bytes(const PqTable*);
void pq_table_close(PqTable*);

/* ---- query ---- */
int pq_query_open(const PqQueryDesc* desc, PqQuery** out);
int pq_query_next(PqQuery* q, int partition, struct ArrowArray* out, struct ArrowSchema* out_schema);
/* All remaining batches of a partition as one Arrow C stream.  The stream borrows the query: release
 * it (or drain it) before pq_query_close; the batches it produced stay valid on their own. */
int pq_query_stream(PqQuery* q, int partition, struct ArrowArrayStream* out);
int pq_query_metrics(PqQuery* q, PqMetrics* out);
/* The whole result (every batch, whatever pq_query_next already handed out) as JSON text formatted on the GPU: what
 * QueryResponse::to_json builds from the batches on the CPU (src/response.rs:31-58, src/utils/arrow/mod.rs:49-64:
 * arrow_json::ArrayWriter conventions -- NULL values leave their key out, Timestamp(ms) as ISO-8601, floats shortest
 * round-trip).  flags 0: one JSON array `[{...},{...}]`; PQ_JSON_LINES: one object per line (NDJSON).  *out stays valid
 * until the next pq_query_json call on q or pq_query_close. */
#define PQ_JSON_LINES 1u
int pq_query_json(PqQuery* q, uint32_t flags, const char** out, uint64_t* len);
const char* pq_last_error(PqQuery* q); /* q == NULL: last error of the calling thread */
void pq_query_close(PqQuery* q);

/* ---- host helpers (no GPU work) ---- */
/* Page-locked host memory for file images: PqFile.buf inside such a block is DMA'd straight
 * to HBM, other host memory is staged through the library's own pinned slices. */
void* pq_host_alloc(uint64_t bytes);
void pq_host_free(void* p);
/* JSON description of a Parquet file as the host metadata layer parsed it (schema leaves, row
 * groups, column chunks, every page header).  Returns the JSON length (excluding NUL), or a
 * negative status; writes at most cap bytes.  Replaces nothing at run time: it exists so the
 * footer/page-header reader can be checked against an independent reader without a GPU. */
int64_t pq_file_describe(const PqFile* file, char* out, uint64_t cap);

/* ---- scan planning on the C side (no GPU work): which manifests / files reach the scan ----
 * What StandardTableProvider::scan does above create_parquet_physical_plan:
 *   time bounds of the filters      src/query/stream_schema_provider.rs:884-940 (extract_timestamp_bound, PartialTimeFilter)
 *   Snapshot::manifests             src/catalog/snapshot.rs:40-71
 *   is_overlapping_query            src/query/stream_schema_provider.rs:750-775
 *   is_within_staging_window        :842-864
 *   can_be_pruned / satisfy_constraints   :955-1043 (TypedStatistics, src/catalog/column.rs:52-68)
 *   collect_from_snapshot           :449-510 (newest first, pruning, LIMIT truncation)
 *   partitioned_files + statistics merge   :351-446, src/catalog/column.rs:70-198
 *   supports_filters_pushdown       :665-683, 866-882
 * The Rust host may keep doing this itself; these entry points let a shim hand the manifest over instead. */
#define PQ_T_TS_NS 6 /* planning only: Timestamp(Nanosecond) literal in PqLiteral.i64 */

typedef struct {
  const char* column; /* NULL: the filter is not `column <cmp> literal` (never prunes, never a time bound) */
  int32_t cmp;        /* PqCmp */
  int32_t _pad;
  PqLiteral lit;
} PqPlanFilter;

typedef enum { PQ_STAT_NONE = 0, PQ_STAT_BOOL = 1, PQ_STAT_INT = 2, PQ_STAT_FLOAT = 3, PQ_STAT_STRING = 4 } PqStatKind;
typedef struct {
  const char* column;
  int32_t kind; /* PqStatKind; PQ_STAT_NONE: the manifest holds no statistics for the column */
  int32_t _pad;
  int64_t min_i, max_i; /* BOOL (0/1), INT */
  double min_f, max_f;  /* FLOAT */
  const char* min_s;    /* STRING, not NUL-terminated */
  uint64_t min_s_len;
  const char* max_s;
  uint64_t max_s_len;
} PqColumnStat;

typedef struct {
  const char* path;
  uint64_t num_rows;
  uint64_t file_size;
  const PqColumnStat* stats;
  uint32_t n_stats;
  uint32_t _pad;
} PqManifestFile;

typedef struct {
  int64_t time_lower_ns; /* naive UTC, nanoseconds since the epoch */
  int64_t time_upper_ns;
} PqManifestItem;

typedef enum { PQ_BOUND_LOW = 0, PQ_BOUND_HIGH = 1, PQ_BOUND_EQ = 2 } PqBoundKind;
typedef struct {
  int32_t kind;     /* PqBoundKind */
  int32_t included; /* LOW / HIGH: the bound itself belongs to the range */
  int64_t time_ns;
} PqTimeBound;

/* One bound per filter that is `column <cmp> timestamp literal` (a Utf8 literal only on `time_partition`).
 * Returns the number of bounds written (<= n), or a negative status. */
int32_t pq_plan_time_bounds(const PqPlanFilter* filters, uint32_t n, const char* time_partition, PqTimeBound* out);
/* keep[i] = 1 when manifest item i can hold rows inside every bound */
int32_t pq_plan_manifests(const PqManifestItem* items, uint32_t n, const PqTimeBound* bounds, uint32_t n_bounds, uint8_t* keep);
int32_t pq_plan_is_overlapping_query(const PqManifestItem* items, uint32_t n, const PqTimeBound* bounds, uint32_t n_bounds);
int32_t pq_plan_within_staging_window(const PqTimeBound* bounds, uint32_t n_bounds, int64_t now_ns);
/* files in manifest order (oldest first) -> out_index: the files to scan, newest first, without those whose
 * statistics rule a filter out, cut once `limit` rows are covered (limit < 0: none).  Returns how many. */
int64_t pq_plan_collect_files(const PqManifestFile* files, uint32_t n_files, const PqPlanFilter* filters, uint32_t n_filters,
                              int64_t limit, uint32_t* out_index);
/* a [min, max] of one column merged over two files; returns 1 and fills *out, or 0 when the ranges cannot be merged
 * (different kinds, an inverted or NaN float range): the planner then skips min / max for that column */
int32_t pq_plan_merge_stat(const PqColumnStat* a, const PqColumnStat* b, PqColumnStat* out);
/* exact[i] = 1: the scan alone answers filter i (minute-aligned time comparison), 0: Inexact */
int32_t pq_plan_pushdown(const PqPlanFilter* filters, uint32_t n, uint8_t* exact);

/* ---- multi-GPU: one process per GPU, NCCL communicator owned by the library ---- */
#define PQ_COMM_ID_BYTES 128
int pq_comm_unique_id(uint8_t id[PQ_COMM_ID_BYTES]);
int pq_comm_init_rank(const uint8_t id[PQ_COMM_ID_BYTES], int nranks, int rank);
int pq_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* PARSEABLE_B200_H */
