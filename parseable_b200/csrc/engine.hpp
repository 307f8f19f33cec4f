// Host side of the B200 query path: HBM-resident tables of encoded column
// chunks, query planning (row-group pruning, work items, predicate/aggregate
// compilation) and the kernel launch sequence.  Mirrors, on the host, what
// StandardTableProvider::scan + create_parquet_physical_plan do before
// DataFusion's operators run (/root/reference/src/query/stream_schema_provider.rs:114-189, 526-659).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/parseable_b200.h"
#include "device_structs.hpp"
#include "parquet_meta.hpp"

namespace pqb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PQB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw ::pqb::Error(_e == cudaErrorMemoryAllocation ? PQ_ERR_OOM : PQ_ERR_CUDA,            \
                         std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
  } while (0)

// ---- process-wide state ----
class Context {
 public:
  static Context& get();
  void init(const int* devices, int n);
  void shutdown();
  void ensure();  // throws PQ_ERR_CUDA when no device is usable; binds the device to the calling thread
  int device() const { return device_; }
  int sm_count() const { return sm_count_; }
  size_t smem_optin() const { return smem_optin_; }
  // pinned staging buffers, grow-only cache
  uint8_t* pinned_acquire(size_t bytes);
  void pinned_release(uint8_t* p);
  bool is_pinned(const void* p);

 private:
  std::mutex mu_;
  bool inited_ = false;
  int device_ = 0;
  int sm_count_ = 0;
  size_t smem_optin_ = 0;
  struct Pinned { uint8_t* p; size_t cap; bool busy; };
  std::vector<Pinned> pinned_;
};

// ---- host view of a Parquet file ----
struct HostFile {
  std::string path;
  const uint8_t* data = nullptr;  // whole file image (mmap or caller buffer)
  uint64_t size = 0;
  bool mapped = false;
  FileMeta meta;
  ~HostFile();
};

struct TablePageRef {
  uint32_t first_page = 0;  // into Table::pages (data pages only)
  uint32_t n_pages = 0;
};

struct TableChunk {
  bool present = false;
  int leaf = -1;
  const ColumnChunkMeta* meta = nullptr;
  uint64_t arena_off = 0;   // where the chunk's bytes start in the arena
  uint64_t file_off = 0;
  uint64_t bytes = 0;       // compressed == uncompressed (codec NONE)
  uint64_t dict_off = 0;    // arena offset of the dictionary payload
  uint32_t dict_len = 0, dict_n = 0;
  TablePageRef pages;
  bool has_dict_pages = false, has_plain_pages = false, has_delta_pages = false;
  uint32_t max_bw = 0;
  uint64_t dict8_off = ~0ull;   // flat store: 8-byte aligned copy of a numeric dictionary
};

struct TableColumn {
  std::string name;
  uint8_t kind = 0;      // DevKind
  bool is_ts = false;
  uint8_t max_def = 0;
};

struct TableRowGroup {
  uint32_t file = 0, rg_in_file = 0;
  uint32_t num_rows = 0;
  uint64_t global_row0 = 0;         // ordinal over ALL row groups of the file list (before sharding)
  std::vector<TableChunk> chunks;   // per table column
  bool pages_aligned = false;       // every present column has the same page boundaries (Parseable's writer: 20 000-row pages)
};

// Query-independent side tables of one table column, built on first use and kept with the table:
// string dictionary entry offsets, and (GROUP BY) the interned group ids of every dictionary entry.
struct KeyDict { std::vector<uint32_t> offs; std::vector<uint8_t> bytes; };
struct ColSide {
  std::vector<uint32_t> base_per_rg;   // dictionary entries of this column in the row groups before g
  uint32_t total_entries = 0;
  uint32_t max_dict_n = 0;
  uint64_t* d_ent_off = nullptr;       // arena offset of every dictionary entry (strings: behind the length prefix)
  bool ent_ready = false;
  // group-key interning (local numbering, hot-first)
  bool key_ready = false;
  uint32_t* d_gid = nullptr;           // group id per dictionary entry
  uint32_t card = 0;
  KeyDict kd;                          // distinct values in group-id order
  uint32_t max_ent_len = 0;            // longest dictionary entry (sizes projected string buffers)
  uint32_t max_plain_len = 0;          // longest value of the column's PLAIN byte-array pages
  bool has_delta = false;              // some page of the column is DELTA_BINARY_PACKED
  bool delta_ready = false;            // ... and its pages have aligned 8-byte copies in d_delta_flat (ensure_plain8)
  uint8_t* d_delta_flat = nullptr;
  uint32_t* d_kd_offs = nullptr;       // the same on the device (result assembly)
  uint8_t* d_kd_bytes = nullptr;
  uint32_t kd_max_len = 0;
  // multi-GPU: the numbering every rank of the communicator agreed on (unify_key), tagged with the communicator's epoch
  bool glob_ready = false;
  uint64_t glob_epoch = 0;
  uint32_t glob_card = 0, glob_max_len = 0;
  KeyDict glob_kd;
  uint32_t* d_glob_gid = nullptr;
  uint32_t* d_glob_kd_offs = nullptr;
  uint8_t* d_glob_kd_bytes = nullptr;
  uint64_t* d_key_hash = nullptr;
  // GROUP BY on a column with pages that have no dictionary (PLAIN fallback, PLAIN / DELTA numerics): every ROW of those
  // pages is an entry behind the dictionary entries; d_gid / d_glob_gid then hold, per such page, one group id per row --
  // the page's "id page", staged by the aggregate kernel like 32-bit dictionary indices (FK_IDS)
  uint32_t key_entries = 0;            // entries d_gid covers (== total_entries when the column has only dictionary pages)
  uint32_t n_dict_pad = 0;             // first row entry (total_entries rounded up to 4)
  uint64_t* d_row_ent = nullptr;       // where every row entry's value sits (relative to the arena, ~0: NULL / padding)
  struct KeyRowPage { uint32_t page; uint32_t ebase; };   // page index, first entry of its rows relative to n_dict_pad
  std::vector<KeyRowPage> key_row_pages;
};

// What a query needs per SET of referenced columns, built once per (table, column set): the chunk
// table, the work items (row ranges between page boundaries common to the columns) and what the
// kernels' shared-memory layouts depend on.
struct Shape {
  std::vector<int> tcols;
  std::vector<DevItem> items;
  DevChunk* d_chunks = nullptr;        // [table row group * ncols + slot]
  DevItem* d_items = nullptr;
  uint32_t n_flat = 0, n_general = 0, n_slab_fast = 0;
  uint32_t bitmap_words = 0;
  std::vector<uint32_t> max_bw, flat_max_bw;          // per slot
  std::vector<uint8_t> has_dict, has_plain, has_delta, flat_plain8, flat_nullable;   // flat_nullable: some flat page of the slot carries a validity bitmap
  std::string why_general;             // first reason an item could not go to the flat kernels (diagnostics)
  std::atomic<unsigned long long> last_total{~0ull};   // rows the last filter scan of this shape selected (sizes the next result)
  ~Shape();
};

// Encoded column chunks of a set of files, resident in HBM ("hot tier in HBM").
class Table {
 public:
  Table() = default;
  ~Table();
  Table(const Table&) = delete;
  void open(const PqFile* files, uint32_t n_files, const std::vector<std::string>& columns, uint32_t shard_index,
            uint32_t shard_count, cudaStream_t stream);

  std::vector<std::unique_ptr<HostFile>> files;
  std::vector<TableColumn> columns;
  std::vector<TableRowGroup> row_groups;
  std::vector<DevPage> pages;       // host copy
  uint8_t* d_arena = nullptr;
  uint64_t arena_bytes = 0;
  DevPage* d_pages = nullptr;
  // slab index (k_slab_index): per page, per kSlabRows rows, the run directory and window start the
  // scan kernel would otherwise derive by walking the run headers; see DESIGN.md
  DevSlabRec* d_slab_recs = nullptr;
  DirEntry* d_slab_dirs = nullptr;
  uint8_t* d_strmat = nullptr;            // DELTA_BYTE_ARRAY pages rewritten as PLAIN BYTE_ARRAY pages (outside the arena: DevPage.off wraps)
  uint8_t* d_slab_flat = nullptr;         // flat bit-packed copies of run-heavy pages (k_flatten_pages)
  uint64_t slab_flat_bytes = 0;
  uint64_t total_slabs = 0;
  std::vector<uint32_t> col_valwin_cap;   // per table column: staged window bytes the index was built for
  uint64_t total_rows = 0;
  uint64_t h2d_bytes = 0;
  uint64_t chunk_bytes = 0;
  int find_column(const std::string& name) const;

  // flat store (flat_store.cuh): header-less bit-packed copies of the NULL-free dictionary-index pages,
  // aligned copies of PLAIN 8-byte pages and of numeric dictionaries
  uint8_t* d_flat = nullptr;
  uint64_t flat_bytes = 0;
  mutable std::vector<FlatPageRec> flat_pages;   // host copy, parallel to pages (ensure_plain8 adds entries later)
  mutable FlatPageRec* d_flat_pages = nullptr;
  uint64_t flat_page_count = 0;
  bool nulls_classified = false;         // build_flat_store looked at every page's definition levels (voff == ~0 then means: no NULLs)

  // lazily built, query independent (the table is immutable once opened); guarded by side_mu
  mutable std::mutex side_mu;
  mutable std::vector<ColSide> sides;
  mutable std::map<std::vector<int>, std::shared_ptr<Shape>> shapes;
  std::shared_ptr<Shape> shape_for(const std::vector<int>& tcols, cudaStream_t stream) const;
  void ensure_ent_off(int tcol, cudaStream_t stream) const;
  void ensure_key(int tcol, cudaStream_t stream) const;
  void unify_key(int tcol, cudaStream_t stream) const;        // collective over the pq_comm communicator
  void ensure_plain8(int tcol, cudaStream_t stream) const;   // DELTA_BINARY_PACKED pages -> row-addressable 8-byte values

 private:
  void build_flat_store(cudaStream_t stream);
};

// staged window bytes for one slab of a dictionary-index stream of the given bit width
inline uint32_t valwin_cap_for_bw(uint32_t max_bw) { return ((kSlabRows * max_bw / 8 + kSlabRows / 8 + 64) + 15u) & ~15u; }
// launches k_slab_index (defined next to k_scan, query.cu)
void launch_slab_index(const uint8_t* arena, const DevPage* pages, uint32_t n_pages, const uint32_t* col_caps, DevSlabRec* recs,
                       DirEntry* dirs, uint8_t* page_fast, cudaStream_t stream);
void launch_flatten_pages(const uint8_t* arena, const DevPage* pages, const void* jobs, uint32_t n_jobs, uint8_t* side,
                          DevSlabRec* recs, DirEntry* dirs, uint8_t* page_fast, cudaStream_t stream);
// launches k_flat_store (flat_store.cuh); jobs are FlatStoreJob records on the device
void launch_flat_store(const uint8_t* arena, const DevPage* pages, const void* jobs, uint32_t n_jobs, uint8_t* flat, uint8_t* ok,
                       uint32_t* maxlen, cudaStream_t stream);
// side-table builders (prep_kernels.cuh), defined in query.cu
void launch_entry_offsets(const Table& t, int tcol, uint64_t* d_out, uint32_t* max_len, cudaStream_t stream);
void launch_dba_lengths(const uint8_t* arena, const DevPage* pages, const DbaJob* jobs, uint32_t n_jobs, uint8_t* scratch, DbaInfo* info, cudaStream_t stream);
void launch_dba_materialise(const uint8_t* arena, const DevPage* pages, const DbaJob* jobs, const DbaInfo* info, uint32_t n_jobs,
                            const uint8_t* scratch, uint8_t* mat, cudaStream_t stream);
void launch_check_flat_indices(const uint8_t* flat, const FlatPageRec* fpages, const uint32_t* dict_n, uint32_t n_pages, uint32_t* first_bad, cudaStream_t stream);
void launch_page_has_nulls(const uint8_t* arena, const DevPage* pages, uint32_t n_pages, uint8_t* out, cudaStream_t stream);
void launch_delta_to_plain8(const uint8_t* arena, const DevPage* pages, const void* jobs, uint32_t n_jobs, uint8_t* flat_base, uint8_t* ok,
                            cudaStream_t stream);
void build_key_side(const Table& t, int tcol, ColSide& side, cudaStream_t stream);
void unify_key_side(const Table& t, int tcol, ColSide& side, cudaStream_t stream);

// page-locked host block that result batches can alias (zero copy); returns to the pool when
// the last batch that references it is released by the consumer
struct PinnedBlock {
  uint8_t* p = nullptr;
  size_t bytes = 0;
  uint8_t* dev = nullptr;   // the device block this one was copied from, while the query keeps it (JSON egress reads it there)
  ~PinnedBlock();
};

struct OutColumn {
  std::string name;
  int type = PQ_T_I64;               // PqType
  std::vector<uint8_t> values;       // 8-byte values, bit-packed bools, or utf8 bytes
  std::shared_ptr<PinnedBlock> ext;  // when set, the values are ext->p + ext_off (not `values`)
  size_t ext_off = 0;
  std::vector<int32_t> offsets;      // utf8
  std::vector<uint8_t> validity;     // empty when null_count == 0
  int64_t null_count = 0;
  // device-assembled results: every buffer of the column lives in `ext`
  bool ext_all = false;
  size_t ext_validity_off = 0;       // validity bitmap (used when null_count != 0)
  size_t ext_offsets_off = 0;        // utf8: int32 offsets of this batch's first row (n + 1 entries follow)
};

struct OutBatch {
  int64_t rows = 0;
  std::vector<OutColumn> cols;
};

class Query {
 public:
  explicit Query(const PqQueryDesc& d);
  ~Query();
  int next(int partition, ArrowArray* out, ArrowSchema* schema);
  void schema(ArrowSchema* out) const;   // of the result batches (an empty struct when the query produced none)
  // every batch of the result as JSON text formatted on the device (json_egress.cuh); the bytes stay valid until the
  // next call or the query is closed
  void json(uint32_t flags, const char** out, uint64_t* len);
  PqMetrics metrics{};
  std::string error;

 private:
  void run(const PqQueryDesc& d);
  std::unique_ptr<Table> owned_table_;
  std::vector<OutBatch> batches_;
  std::vector<std::shared_ptr<PinnedBlock>> dev_blocks_;   // result blocks whose device copy is kept until the query closes
  std::shared_ptr<PinnedBlock> json_block_;
  uint32_t batch_rows_ = 20000;
  size_t next_batch_ = 0;
  bool schema_only_done_ = false;
};

void export_batch(const OutBatch& b, ArrowArray* out, ArrowSchema* schema);
std::string describe_file(const PqFile& f);

// NCCL communicator owned by the library (one process per GPU)
int comm_unique_id(uint8_t* id);
int comm_init_rank(const uint8_t* id, int nranks, int rank);
int comm_destroy();
bool comm_active();
uint64_t comm_epoch();        // changes with every communicator
void comm_group_begin();      // ncclGroupStart / End: the collectives in between launch as one
void comm_group_end();
int comm_nranks();
int comm_rank();
void comm_allreduce_u64(void* buf, size_t count, int op /*0 sum,1 min(s64),2 max(s64),3 sum f64*/, cudaStream_t s);
void comm_allgather_bytes(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);

}  // namespace pqb
