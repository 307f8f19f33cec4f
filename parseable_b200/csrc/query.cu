// Query planning and execution: the C-ABI PqQueryDesc becomes a DevPlan, work
// items and side tables; then  k_dict_entry_offsets -> k_leaf_luts ->
// k_key_intern x2 -> k_scan -> k_agg_compact  run on one stream.
//
// Reference behaviour restated here (all /root/reference paths):
//   * predicate pushed into the scan AND re-applied (Inexact pushdown,
//     src/query/stream_schema_provider.rs:665-683): one fused evaluation gives the
//     same rows;
//   * row-group pruning from footer min/max (ParquetFormat::with_enable_pruning,
//     :146): pruning never changes results, only rows_scanned;
//   * SQL three-valued logic, NULL group keys, COUNT -> Int64, SUM(Int64) wrapping,
//     float totalOrder.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

#include "engine.hpp"
#include "prep_kernels.cuh"
#include "scan_kernel.cuh"

namespace pqb {

namespace {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  void alloc(size_t count, cudaStream_t stream) {
    n = count;
    s = stream;
    if (count) PQB_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), stream));
  }
  void upload(const std::vector<T>& v, cudaStream_t stream) {
    alloc(v.size(), stream);
    if (!v.empty()) PQB_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  }
  void zero() { if (n) PQB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  ~DevBuf() { if (p) cudaFreeAsync(p, s); }
};

struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

// ---- LIKE pattern classification (arrow-string like.rs fast paths) ----
struct LikePlan { uint32_t kind; std::string needle; };
LikePlan classify_like(const std::string& pat) {
  // unescaped structure: sequence of literal chars and wildcards
  std::string lit;
  std::vector<int> tokens;  // -1 '%', -2 '_', >=0 literal byte
  for (size_t i = 0; i < pat.size(); i++) {
    char c = pat[i];
    if (c == '\\' && i + 1 < pat.size()) { tokens.push_back((unsigned char)pat[++i]); }
    else if (c == '%') tokens.push_back(-1);
    else if (c == '_') tokens.push_back(-2);
    else tokens.push_back((unsigned char)c);
  }
  bool has_us = false;
  int npct = 0;
  for (int t : tokens) { has_us |= t == -2; npct += t == -1; }
  auto literal_of = [&](size_t b, size_t e) { std::string s; for (size_t i = b; i < e; i++) s.push_back(char(tokens[i])); return s; };
  if (!has_us) {
    if (npct == 0) return {LIKE_EQ, literal_of(0, tokens.size())};
    // collapse leading / trailing runs of '%'
    size_t b = 0, e = tokens.size();
    while (b < e && tokens[b] == -1) b++;
    while (e > b && tokens[e - 1] == -1) e--;
    bool inner = false;
    for (size_t i = b; i < e; i++) inner |= tokens[i] == -1;
    if (!inner) {
      bool lead = b > 0, trail = e < tokens.size();
      std::string n = literal_of(b, e);
      if (lead && trail) return {LIKE_CONTAINS, n};
      if (trail) return {LIKE_PREFIX, n};
      if (lead) return {LIKE_SUFFIX, n};
    }
  }
  return {LIKE_GENERAL, pat};
}

uint64_t f64_bits(double d) { uint64_t b; std::memcpy(&b, &d, 8); return b; }
double bits_f64(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; }

// tri-state for pruning
enum Tri { TRI_FALSE = 0, TRI_TRUE = 1, TRI_MAYBE = 2 };

struct HostLeaf {
  DevLeaf d{};
  int qcol = -1;
  std::string str;
};

// Can `col <cmp> lit` be decided for a whole chunk from its min/max statistics?
Tri leaf_from_stats(const HostLeaf& lf, uint8_t kind, const TableChunk& ch, uint32_t rg_rows) {
  if (!ch.present) {
    // column absent: all NULL
    if (lf.d.kind == LK_IS_NULL) return TRI_TRUE;
    return TRI_FALSE;  // NULL compare is never TRUE; IS NOT NULL is FALSE
  }
  const ColumnStats& st = ch.meta->stats;
  bool no_nulls = st.null_count == 0;
  bool all_nulls = st.null_count >= 0 && uint64_t(st.null_count) == uint64_t(ch.meta->num_values);
  if (lf.d.kind == LK_IS_NULL) return no_nulls ? TRI_FALSE : (all_nulls ? TRI_TRUE : TRI_MAYBE);
  if (lf.d.kind == LK_IS_NOT_NULL) return no_nulls ? TRI_TRUE : (all_nulls ? TRI_FALSE : TRI_MAYBE);
  if (all_nulls) return TRI_FALSE;
  if (lf.d.kind != LK_CMP || !st.has_min || !st.has_max) return TRI_MAYBE;
  int lo_c, hi_c;  // sign of compare(min, lit), compare(max, lit)
  if (kind == DK_I64) {
    if (st.min.size() != 8 || st.max.size() != 8) return TRI_MAYBE;
    int64_t mn, mx;
    std::memcpy(&mn, st.min.data(), 8);
    std::memcpy(&mx, st.max.data(), 8);
    lo_c = mn < lf.d.lit_i64 ? -1 : (mn > lf.d.lit_i64 ? 1 : 0);
    hi_c = mx < lf.d.lit_i64 ? -1 : (mx > lf.d.lit_i64 ? 1 : 0);
  } else if (kind == DK_F64) {
    if (st.min.size() != 8 || st.max.size() != 8) return TRI_MAYBE;
    double mn, mx, lit = bits_f64(uint64_t(lf.d.lit_i64));
    std::memcpy(&mn, st.min.data(), 8);
    std::memcpy(&mx, st.max.data(), 8);
    // footer statistics ignore NaN and may fold -0.0/+0.0: only decide on clean finite bounds
    if (std::isnan(mn) || std::isnan(mx) || std::isnan(lit)) return TRI_MAYBE;
    if (mn == 0.0 || mx == 0.0 || lit == 0.0) return TRI_MAYBE;
    lo_c = mn < lit ? -1 : (mn > lit ? 1 : 0);
    hi_c = mx < lit ? -1 : (mx > lit ? 1 : 0);
    // a chunk may still hold NaN (greater than everything in totalOrder) outside [min,max]
    // -> never claim TRI_TRUE/FALSE on the upper side
    switch (lf.d.cmp) {
      case PQ_LT: case PQ_LE: return (lf.d.cmp == PQ_LT ? lo_c >= 0 : lo_c > 0) ? TRI_FALSE : TRI_MAYBE;
      case PQ_EQ: return (lo_c > 0) ? TRI_FALSE : TRI_MAYBE;
      default: return TRI_MAYBE;
    }
  } else if (kind == DK_STR) {
    auto cmpb = [&](const std::string& a) {
      int c = cmp_bytes((const uint8_t*)a.data(), uint32_t(a.size()), (const uint8_t*)lf.str.data(), uint32_t(lf.str.size()));
      return c;
    };
    lo_c = cmpb(st.min);
    hi_c = cmpb(st.max);
    // string max statistics may be truncated upper bounds: only use them to rule rows OUT
    switch (lf.d.cmp) {
      case PQ_EQ: return (lo_c > 0 || hi_c < 0) ? TRI_FALSE : TRI_MAYBE;
      case PQ_LT: return lo_c >= 0 ? TRI_FALSE : TRI_MAYBE;
      case PQ_LE: return lo_c > 0 ? TRI_FALSE : TRI_MAYBE;
      case PQ_GT: return hi_c <= 0 ? TRI_FALSE : TRI_MAYBE;
      case PQ_GE: return hi_c < 0 ? TRI_FALSE : TRI_MAYBE;
      default: return TRI_MAYBE;
    }
  } else {
    return TRI_MAYBE;
  }
  (void)rg_rows;
  Tri r = TRI_MAYBE;
  switch (lf.d.cmp) {
    case PQ_EQ: r = (lo_c > 0 || hi_c < 0) ? TRI_FALSE : ((lo_c == 0 && hi_c == 0) ? TRI_TRUE : TRI_MAYBE); break;
    case PQ_NE: r = (lo_c > 0 || hi_c < 0) ? TRI_TRUE : ((lo_c == 0 && hi_c == 0) ? TRI_FALSE : TRI_MAYBE); break;
    case PQ_LT: r = hi_c < 0 ? TRI_TRUE : (lo_c >= 0 ? TRI_FALSE : TRI_MAYBE); break;
    case PQ_LE: r = hi_c <= 0 ? TRI_TRUE : (lo_c > 0 ? TRI_FALSE : TRI_MAYBE); break;
    case PQ_GT: r = lo_c > 0 ? TRI_TRUE : (hi_c <= 0 ? TRI_FALSE : TRI_MAYBE); break;
    case PQ_GE: r = lo_c >= 0 ? TRI_TRUE : (hi_c < 0 ? TRI_FALSE : TRI_MAYBE); break;
  }
  if (r == TRI_TRUE && !no_nulls) r = TRI_MAYBE;  // NULL rows evaluate to NULL, not TRUE
  return r;
}

Tri tri_and(Tri a, Tri b) { return (a == TRI_FALSE || b == TRI_FALSE) ? TRI_FALSE : ((a == TRI_TRUE && b == TRI_TRUE) ? TRI_TRUE : TRI_MAYBE); }
Tri tri_or(Tri a, Tri b) { return (a == TRI_TRUE || b == TRI_TRUE) ? TRI_TRUE : ((a == TRI_FALSE && b == TRI_FALSE) ? TRI_FALSE : TRI_MAYBE); }
// NOT of "certainly not TRUE" is not "certainly TRUE" under NULLs: keep MAYBE
Tri tri_not(Tri a) { (void)a; return TRI_MAYBE; }

const char* type_name(int t) {
  switch (t) { case PQ_T_BOOL: return "Boolean"; case PQ_T_I64: return "Int64"; case PQ_T_F64: return "Float64";
    case PQ_T_UTF8: return "Utf8"; case PQ_T_TS_MS: return "Timestamp(ms)"; default: return "Null"; }
}

uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) & ~(a - 1); }

}  // namespace

Query::Query(const PqQueryDesc& d) {
  run(d);
}
Query::~Query() = default;

void launch_slab_index(const uint8_t* arena, const DevPage* pages, uint32_t n_pages, const uint32_t* col_caps, DevSlabRec* recs,
                       DirEntry* dirs, uint8_t* page_fast, cudaStream_t stream) {
  if (!n_pages) return;
  k_slab_index<<<(n_pages + 63) / 64, 64, 0, stream>>>(arena, pages, n_pages, col_caps, recs, dirs, page_fast);
  PQB_CUDA(cudaGetLastError());
}

void launch_flatten_pages(const uint8_t* arena, const DevPage* pages, const void* jobs, uint32_t n_jobs, uint8_t* side,
                          DevSlabRec* recs, DirEntry* dirs, uint8_t* page_fast, cudaStream_t stream) {
  if (!n_jobs) return;
  k_flatten_pages<<<(n_jobs + 31) / 32, 32, 0, stream>>>(arena, pages, static_cast<const FlatJob*>(jobs), n_jobs, side, recs, dirs, page_fast);
  PQB_CUDA(cudaGetLastError());
}

void Query::run(const PqQueryDesc& d) {
  const auto t_begin = std::chrono::steady_clock::now();
  const bool verbose = getenv("PQB_VERBOSE") != nullptr;
  auto mark = [&](const char* what) {   // PQB_VERBOSE: host timeline of this query
    if (verbose)
      fprintf(stderr, "[pqb] +%.3f ms %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), what);
  };
  struct HostTimer {
    std::chrono::steady_clock::time_point t0;
    PqMetrics* m;
    ~HostTimer() { m->host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
  } host_timer{t_begin, &metrics};
  Context& ctx = Context::get();
  ctx.ensure();
  if (d.n_columns > (uint32_t)kMaxCols) throw Error(PQ_ERR_UNSUPPORTED, "too many referenced columns");
  if (d.n_aggs > (uint32_t)kMaxAggs) throw Error(PQ_ERR_UNSUPPORTED, "too many aggregates");
  if (d.n_group_by > (uint32_t)kMaxKeys) throw Error(PQ_ERR_UNSUPPORTED, "too many GROUP BY columns");
  if (d.n_pred > (uint32_t)kMaxPredOps) throw Error(PQ_ERR_UNSUPPORTED, "predicate program too long");
  if (d.n_group_by && !d.n_aggs) throw Error(PQ_ERR_INVALID_ARG, "GROUP BY without aggregates");

  cudaStream_t stream;
  PQB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{stream};

  // ---- input: resident table, or upload the referenced columns of a file list ----
  const Table* table = reinterpret_cast<const Table*>(d.table);
  std::vector<int> tcol(d.n_columns, -1);  // query column -> table column
  if (!table) {
    if (!d.files || !d.n_files) throw Error(PQ_ERR_INVALID_ARG, "query has neither a table nor files");
    std::vector<std::string> names;
    for (uint32_t c = 0; c < d.n_columns; c++) names.push_back(d.columns[c].name ? d.columns[c].name : "");
    owned_table_ = std::make_unique<Table>();
    owned_table_->open(d.files, d.n_files, names, d.shard_index, d.shard_count, stream);
    metrics.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    table = owned_table_.get();
    metrics.h2d_bytes += table->h2d_bytes;
  }
  for (uint32_t c = 0; c < d.n_columns; c++) {
    if (!d.columns[c].name) throw Error(PQ_ERR_INVALID_ARG, "column without a name");
    tcol[c] = table->find_column(d.columns[c].name);
    if (tcol[c] < 0) throw Error(PQ_ERR_INVALID_ARG, std::string("column '") + d.columns[c].name + "' is not part of the resident table");
  }

  // ---- column kinds vs the plan's expectation ----
  DevPlan plan{};
  plan.ncols = d.n_columns;
  std::vector<bool> col_all_null(d.n_columns, false);
  for (uint32_t c = 0; c < d.n_columns; c++) {
    const TableColumn& tc = table->columns[tcol[c]];
    uint8_t kind = tc.kind;
    int want = d.columns[c].type;
    if (kind == 0xfe) {  // in no file: all NULL, take the plan's type
      col_all_null[c] = true;
      kind = want == PQ_T_F64 ? DK_F64 : want == PQ_T_UTF8 ? DK_STR : want == PQ_T_BOOL ? DK_BOOL : DK_I64;
    } else {
      bool ok = (want == PQ_T_I64 && kind == DK_I64 && !tc.is_ts) || (want == PQ_T_TS_MS && kind == DK_I64 && tc.is_ts) ||
                (want == PQ_T_F64 && kind == DK_F64) || (want == PQ_T_UTF8 && kind == DK_STR) ||
                (want == PQ_T_BOOL && kind == DK_BOOL) || want == PQ_T_NULL;
      // Int64 plan type also accepts a timestamp column and vice versa (same physical values)
      if (!ok && kind == DK_I64 && (want == PQ_T_I64 || want == PQ_T_TS_MS)) ok = true;
      if (!ok) throw Error(PQ_ERR_INVALID_ARG, std::string("column '") + tc.name + "' is not " + type_name(want) + " in the Parquet files");
    }
    plan.cols[c].kind = kind;
    plan.cols[c].max_def = tc.max_def;
  }
  auto out_type_of = [&](uint32_t c) -> int {
    const TableColumn& tc = table->columns[tcol[c]];
    if (d.columns[c].type != PQ_T_NULL) return d.columns[c].type;
    switch (plan.cols[c].kind) { case DK_F64: return PQ_T_F64; case DK_STR: return PQ_T_UTF8; case DK_BOOL: return PQ_T_BOOL;
      default: return tc.is_ts ? PQ_T_TS_MS : PQ_T_I64; }
  };

  // ---- predicate compile ----
  std::vector<HostLeaf> leaves;
  std::vector<DevPredOp> prog;
  std::vector<uint8_t> lit_pool(16, 0);
  {
    int depth = 0;
    for (uint32_t i = 0; i < d.n_pred; i++) {
      const PqPredOp& op = d.pred[i];
      switch (op.kind) {
        case PQ_OP_CMP: case PQ_OP_IS_NULL: case PQ_OP_IS_NOT_NULL: case PQ_OP_LIKE: {
          if (op.col < 0 || uint32_t(op.col) >= d.n_columns) throw Error(PQ_ERR_INVALID_ARG, "predicate column out of range");
          if (leaves.size() >= (size_t)kMaxLeaves) throw Error(PQ_ERR_UNSUPPORTED, "too many leaf predicates");
          HostLeaf lf;
          lf.qcol = op.col;
          lf.d.col = uint8_t(op.col);
          uint8_t kind = plan.cols[op.col].kind;
          if (op.kind == PQ_OP_IS_NULL) lf.d.kind = LK_IS_NULL;
          else if (op.kind == PQ_OP_IS_NOT_NULL) lf.d.kind = LK_IS_NOT_NULL;
          else if (op.kind == PQ_OP_LIKE) {
            if (kind != DK_STR) throw Error(PQ_ERR_INVALID_ARG, "LIKE needs a Utf8 column");
            if (op.lit.type != PQ_T_UTF8) throw Error(PQ_ERR_INVALID_ARG, "LIKE needs a Utf8 pattern");
            LikePlan lp = classify_like(std::string(op.lit.str ? op.lit.str : "", op.lit.str_len));
            lf.d.kind = LK_LIKE;
            lf.d.cmp = uint8_t(lp.kind);
            lf.d.flags = op.flags;
            lf.str = lp.needle;
          } else {
            lf.d.kind = LK_CMP;
            if (op.cmp < PQ_EQ || op.cmp > PQ_GE) throw Error(PQ_ERR_INVALID_ARG, "bad comparison operator");
            lf.d.cmp = uint8_t(op.cmp);
            // literal coercion as DataFusion's type coercion does for column-vs-literal
            switch (kind) {
              case DK_I64:
                if (op.lit.type == PQ_T_I64 || op.lit.type == PQ_T_TS_MS) lf.d.lit_i64 = op.lit.i64;
                else if (op.lit.type == PQ_T_F64 && std::nearbyint(op.lit.f64) == op.lit.f64 && std::fabs(op.lit.f64) < 9.2e18)
                  lf.d.lit_i64 = int64_t(op.lit.f64);
                else throw Error(PQ_ERR_UNSUPPORTED, "Int64 column compared with a non-integer literal");
                break;
              case DK_F64:
                if (op.lit.type == PQ_T_F64) lf.d.lit_i64 = int64_t(f64_bits(op.lit.f64));
                else if (op.lit.type == PQ_T_I64) lf.d.lit_i64 = int64_t(f64_bits(double(op.lit.i64)));  // `status = 200` on a Float64 column
                else throw Error(PQ_ERR_INVALID_ARG, "Float64 column compared with a non-numeric literal");
                break;
              case DK_BOOL:
                if (op.lit.type != PQ_T_BOOL) throw Error(PQ_ERR_INVALID_ARG, "Boolean column compared with a non-boolean literal");
                lf.d.lit_i64 = op.lit.i64 ? 1 : 0;
                break;
              case DK_STR:
                if (op.lit.type != PQ_T_UTF8) throw Error(PQ_ERR_INVALID_ARG, "Utf8 column compared with a non-string literal");
                lf.str.assign(op.lit.str ? op.lit.str : "", op.lit.str_len);
                break;
              default: throw Error(PQ_ERR_UNSUPPORTED, "comparison on this column type");
            }
          }
          if (kind == DK_STR && (lf.d.kind == LK_CMP || lf.d.kind == LK_LIKE)) {
            lf.d.str_off = uint32_t(lit_pool.size());
            lf.d.str_len = uint32_t(lf.str.size());
            lit_pool.insert(lit_pool.end(), lf.str.begin(), lf.str.end());
            lit_pool.resize(align_up(uint32_t(lit_pool.size()) + 8, 8), 0);
          }
          prog.push_back({PK_LEAF, uint8_t(leaves.size())});
          leaves.push_back(lf);
          depth++;
          break;
        }
        case PQ_OP_AND: case PQ_OP_OR:
          if (depth < 2) throw Error(PQ_ERR_INVALID_ARG, "predicate program underflow");
          prog.push_back({uint8_t(op.kind == PQ_OP_AND ? PK_AND : PK_OR), 0});
          depth--;
          break;
        case PQ_OP_NOT:
          if (depth < 1) throw Error(PQ_ERR_INVALID_ARG, "predicate program underflow");
          prog.push_back({PK_NOT, 0});
          break;
        case PQ_OP_CONST:
          prog.push_back({PK_CONST, uint8_t(op.lit.type == PQ_T_NULL ? 2 : (op.lit.i64 ? 1 : 0))});
          depth++;
          break;
        default: throw Error(PQ_ERR_INVALID_ARG, "unknown predicate op");
      }
      if (depth > kPredStack) throw Error(PQ_ERR_UNSUPPORTED, "predicate nesting too deep");
    }
    if (d.n_pred && depth != 1) throw Error(PQ_ERR_INVALID_ARG, "predicate program does not reduce to one value");
  }

  // ---- row-group pruning + constant folding of leaves that statistics decide everywhere ----
  std::vector<uint32_t> rgs;  // surviving table row groups
  std::vector<int> leaf_const(leaves.size(), -1);  // -1 unknown; else Tri over all survivors
  metrics.row_groups_total = table->row_groups.size();
  for (uint32_t g = 0; g < table->row_groups.size(); g++) {
    const TableRowGroup& rg = table->row_groups[g];
    std::vector<Tri> lt(leaves.size());
    for (size_t l = 0; l < leaves.size(); l++) {
      const TableChunk& ch = rg.chunks[tcol[leaves[l].qcol]];
      lt[l] = leaf_from_stats(leaves[l], plan.cols[leaves[l].qcol].kind, ch, rg.num_rows);
    }
    Tri root = TRI_TRUE;
    if (!prog.empty()) {
      std::vector<Tri> st;
      for (const DevPredOp& op : prog) {
        if (op.kind == PK_LEAF) st.push_back(lt[op.arg]);
        else if (op.kind == PK_CONST) st.push_back(op.arg == 1 ? TRI_TRUE : TRI_FALSE);
        else if (op.kind == PK_NOT) st.back() = tri_not(st.back());
        else { Tri b = st.back(); st.pop_back(); st.back() = op.kind == PK_AND ? tri_and(st.back(), b) : tri_or(st.back(), b); }
      }
      root = st[0];
    }
    if (root == TRI_FALSE) { metrics.row_groups_pruned++; continue; }
    for (size_t l = 0; l < leaves.size(); l++) {
      if (leaf_const[l] == -1) leaf_const[l] = lt[l];
      else if (leaf_const[l] != lt[l]) leaf_const[l] = TRI_MAYBE;
    }
    rgs.push_back(g);
    metrics.rows_scanned += rg.num_rows;
  }
  // a leaf that is TRUE in every surviving row group is replaced by a constant: the injected
  // p_timestamp range filter (src/query/mod.rs:774-833) usually disappears here and its column
  // is then never read.  (A NOT above it is fine: TRUE means "TRUE for every row, no NULLs".)
  std::vector<bool> leaf_live(leaves.size(), true);
  for (DevPredOp& op : prog)
    if (op.kind == PK_LEAF && leaf_const[op.arg] == TRI_TRUE) { leaf_live[op.arg] = false; op = {PK_CONST, 1}; }

  // ---- which columns does the kernel really read? ----
  std::vector<bool> col_used(d.n_columns, false);
  for (size_t l = 0; l < leaves.size(); l++) if (leaf_live[l]) col_used[leaves[l].qcol] = true;
  for (uint32_t k = 0; k < d.n_group_by; k++) {
    if (d.group_by[k] < 0 || uint32_t(d.group_by[k]) >= d.n_columns) throw Error(PQ_ERR_INVALID_ARG, "group-by column out of range");
    col_used[d.group_by[k]] = true;
  }
  for (uint32_t a = 0; a < d.n_aggs; a++) {
    if (d.aggs[a].fn == PQ_AGG_COUNT_STAR) continue;
    if (d.aggs[a].col < 0 || uint32_t(d.aggs[a].col) >= d.n_columns) throw Error(PQ_ERR_INVALID_ARG, "aggregate column out of range");
    col_used[d.aggs[a].col] = true;
  }
  // compact to kernel column slots
  std::vector<int> slot_of(d.n_columns, -1);
  std::vector<uint32_t> qcol_of_slot;
  for (uint32_t c = 0; c < d.n_columns; c++)
    if (col_used[c]) { slot_of[c] = int(qcol_of_slot.size()); qcol_of_slot.push_back(c); }
  const uint32_t ncols = uint32_t(qcol_of_slot.size());
  {
    DevPlan p2 = plan;
    for (uint32_t s = 0; s < ncols; s++) p2.cols[s] = plan.cols[qcol_of_slot[s]];
    plan = p2;
    plan.ncols = ncols;
  }
  // renumber live leaves
  std::vector<int> leaf_slot(leaves.size(), -1);
  uint32_t nleaves = 0;
  for (size_t l = 0; l < leaves.size(); l++) {
    if (!leaf_live[l]) continue;
    leaf_slot[l] = int(nleaves);
    plan.leaves[nleaves] = leaves[l].d;
    plan.leaves[nleaves].col = uint8_t(slot_of[leaves[l].qcol]);
    nleaves++;
  }
  plan.nleaves = nleaves;
  // per column: the leaves a dictionary LUT answers (the kernel fuses up to two into the unpack)
  for (uint32_t c = 0; c < (uint32_t)kMaxCols; c++) { plan.col_nlut[c] = 0; plan.col_l0[c] = -1; plan.col_l1[c] = -1; }
  for (uint32_t l = 0; l < nleaves; l++) {
    const DevLeaf& lf = plan.leaves[l];
    if (lf.kind != LK_CMP && lf.kind != LK_LIKE) continue;
    if (plan.col_nlut[lf.col] == 0) plan.col_l0[lf.col] = int8_t(l);
    else if (plan.col_nlut[lf.col] == 1) plan.col_l1[lf.col] = int8_t(l);
    plan.col_nlut[lf.col]++;
  }
  {
    const char* rm = getenv("PQB_ROW_MAJOR");  // experiment switch: register-only row-major pass for no-NULL slabs
    plan.row_major = rm && rm[0] == '1';
  }
  {
    // conjunction of 1-4 CMP/LIKE leaves (folded TRUE constants are neutral): specialised row pass
    bool conj = nleaves >= 1 && nleaves <= 4;
    uint32_t leaves_seen = 0;
    for (const DevPredOp& op : prog) {
      if (op.kind == PK_LEAF) { leaves_seen++; conj &= plan.leaves[leaf_slot[op.arg]].kind == LK_CMP || plan.leaves[leaf_slot[op.arg]].kind == LK_LIKE; }
      else if (op.kind == PK_AND) {}
      else if (op.kind == PK_CONST && op.arg == 1) {}
      else conj = false;
    }
    const char* fa = getenv("PQB_FAST_AND");
    plan.fast_and = conj && leaves_seen == nleaves && !(fa && fa[0] == '0');
  }
  plan.npred = uint32_t(prog.size());
  for (size_t i = 0; i < prog.size(); i++) {
    plan.pred[i] = prog[i];
    if (prog[i].kind == PK_LEAF) plan.pred[i].arg = uint8_t(leaf_slot[prog[i].arg]);
  }

  // ---- aggregates ----
  const bool has_aggs = d.n_aggs > 0;
  bool only_count_star = has_aggs && d.n_group_by == 0;
  for (uint32_t a = 0; a < d.n_aggs; a++) only_count_star &= d.aggs[a].fn == PQ_AGG_COUNT_STAR;
  const bool agg_kernel = has_aggs && !only_count_star;
  plan.mode = agg_kernel ? SM_AGG : SM_FILTER;
  std::vector<int> agg_out_type(d.n_aggs, PQ_T_I64);
  if (has_aggs) {
    std::map<int, int> nn_of_col;
    uint32_t n_acc = 0;
    plan.naggs = d.n_aggs;
    for (uint32_t a = 0; a < d.n_aggs; a++) {
      DevAgg& ag = plan.aggs[a];
      ag = DevAgg{};
      ag.fn = uint8_t(d.aggs[a].fn);
      if (ag.fn == AG_COUNT_STAR) { agg_out_type[a] = PQ_T_I64; continue; }
      if (ag.fn > AG_AVG) throw Error(PQ_ERR_INVALID_ARG, "unknown aggregate function");
      uint32_t qc = uint32_t(d.aggs[a].col);
      ag.col = uint8_t(slot_of[qc]);
      ag.kind = plan.cols[ag.col].kind;
      if (ag.fn != AG_COUNT && ag.kind != DK_I64 && ag.kind != DK_F64)
        throw Error(PQ_ERR_UNSUPPORTED, std::string("SUM/MIN/MAX/AVG over ") + type_name(out_type_of(qc)) + " is not on the GPU path");
      auto it = nn_of_col.find(int(qc));
      if (it == nn_of_col.end()) { it = nn_of_col.emplace(int(qc), int(nn_of_col.size())).first; ag.update_nn = 1; }
      ag.nn_slot = uint8_t(it->second);
      if (ag.fn == AG_COUNT) { agg_out_type[a] = PQ_T_I64; continue; }
      ag.acc_slot = uint8_t(n_acc);
      uint8_t how = 0;
      if (ag.fn == AG_SUM) how = ag.kind == DK_F64 ? 1 : 0;
      else if (ag.fn == AG_AVG) how = 1;
      else how = ag.fn == AG_MIN ? 2 : 3;
      plan.acc_init[n_acc++] = how;
      agg_out_type[a] = ag.fn == AG_AVG ? PQ_T_F64 : out_type_of(qc);
    }
    plan.n_acc = n_acc;
    plan.n_nn = uint32_t(nn_of_col.size());
  }

  mark("plan compiled");
  // ---- per query chunk table, work items ----
  const uint32_t nrg = uint32_t(rgs.size());
  std::vector<DevChunk> chunks(size_t(nrg) * std::max<uint32_t>(ncols, 1));
  std::vector<DevItem> items;
  uint32_t n_fast_items = 0;   // items the table's slab index covers: no in-kernel run-header walk
  const bool use_slab_index = ncols > 0 && table->d_slab_recs != nullptr;
  std::vector<uint8_t> col_needs_ent(ncols, 0);  // entry offsets (string leaf / any key column)
  std::vector<uint8_t> col_has_lut(ncols, 0);
  for (uint32_t l = 0; l < nleaves; l++) {
    const DevLeaf& lf = plan.leaves[l];
    if (lf.kind == LK_CMP || lf.kind == LK_LIKE) {
      col_has_lut[lf.col] = 1;
      if (plan.cols[lf.col].kind == DK_STR) col_needs_ent[lf.col] = 1;
    }
  }
  for (uint32_t k = 0; k < d.n_group_by; k++) {
    uint32_t s = uint32_t(slot_of[d.group_by[k]]);
    if (plan.cols[s].kind != DK_BOOL) { col_needs_ent[s] = 1; col_has_lut[s] = 1; }
  }
  uint64_t total_entries = 0;
  uint32_t bitmap_words = 0;
  uint64_t algo_bytes = 0, scanned_bytes = 0;
  std::vector<uint8_t> col_has_nulls(std::max<uint32_t>(ncols, 1), 0);   // statistics cannot rule NULLs out
  std::vector<std::vector<uint32_t>> bounds;   // reused across row groups
  std::vector<uint32_t> common;
  size_t bounds_n = 0;
  items.reserve(size_t(nrg) * 16);
  for (uint32_t gi = 0; gi < nrg; gi++) {
    const TableRowGroup& rg = table->row_groups[rgs[gi]];
    bounds_n = 0;
    int first_present = -1;
    for (uint32_t s = 0; s < ncols; s++) {
      const TableChunk& tc = rg.chunks[tcol[qcol_of_slot[s]]];
      DevChunk& dc = chunks[size_t(gi) * ncols + s];
      dc.present = tc.present ? 1 : 0;
      if (!tc.present || tc.meta->stats.null_count != 0) col_has_nulls[s] = 1;   // absent column: every row NULL
      if (!tc.present) continue;
      const uint8_t kind = plan.cols[s].kind;
      dc.dict_off = tc.dict_off;
      dc.dict_len = tc.dict_len;
      dc.dict_n = tc.dict_n;
      dc.first_page = tc.pages.first_page;
      dc.n_pages = tc.pages.n_pages;
      if (col_has_lut[s] || col_needs_ent[s]) {
        dc.lut_base = uint32_t(total_entries);
        total_entries += tc.dict_n;
        if (total_entries > 0xfffffff0ull) throw Error(PQ_ERR_UNSUPPORTED, "too many dictionary entries for one query");
      }
      if (tc.has_delta_pages && kind != DK_I64)
        throw Error(PQ_ERR_UNSUPPORTED, "column '" + table->columns[tcol[qcol_of_slot[s]]].name + "': DELTA_BINARY_PACKED is decoded for INT64 columns only");
      plan.cols[s].has_delta |= tc.has_delta_pages;
      if (tc.has_plain_pages && kind == DK_STR)
        throw Error(PQ_ERR_UNSUPPORTED, "column '" + table->columns[tcol[qcol_of_slot[s]]].name + "': PLAIN (dictionary-fallback) string pages are not decoded on the GPU yet");
      if (kind == DK_STR && tc.dict_n == 0 && tc.has_dict_pages && false) {}
      plan.cols[s].max_bw = std::max(plan.cols[s].max_bw, tc.max_bw);
      plan.cols[s].has_dict |= tc.has_dict_pages;
      plan.cols[s].has_plain |= tc.has_plain_pages;
      scanned_bytes += tc.bytes;
      algo_bytes += uint64_t(tc.meta->total_uncompressed_size);
      if (first_present < 0) first_present = int(s);
      if (!rg.pages_aligned) {
        if (bounds.size() <= bounds_n) bounds.emplace_back();
        std::vector<uint32_t>& b = bounds[bounds_n++];
        b.clear();
        for (uint32_t p = 0; p < tc.pages.n_pages; p++) b.push_back(table->pages[tc.pages.first_page + p].first_row);
      }
    }
    // boundaries common to every present column
    common.clear();
    const bool aligned = rg.pages_aligned && first_present >= 0;
    if (aligned) {   // the pages ARE the items
      const TableChunk& tc0 = rg.chunks[tcol[qcol_of_slot[first_present]]];
      for (uint32_t p = 0; p < tc0.pages.n_pages; p++) common.push_back(table->pages[tc0.pages.first_page + p].first_row);
    } else if (bounds_n == 0) common.push_back(0);
    else {
      common = bounds[0];
      for (size_t i = 1; i < bounds_n; i++) {
        std::vector<uint32_t> t;
        std::set_intersection(common.begin(), common.end(), bounds[i].begin(), bounds[i].end(), std::back_inserter(t));
        common.swap(t);
      }
    }
    if (common.empty() || common[0] != 0) throw Error(PQ_ERR_CORRUPT, "row group pages do not start at row 0");
    for (size_t i = 0; i < common.size(); i++) {
      DevItem it{};
      it.rg = gi;
      it.row0 = common[i];
      it.nrows = (i + 1 < common.size() ? common[i + 1] : rg.num_rows) - common[i];
      it.global_row0 = rg.global_row0 + common[i];
      it.bitmap_word0 = bitmap_words;
      bitmap_words += (it.nrows + 31) / 32 + 1;
      bool fast = use_slab_index;
      for (uint32_t s = 0; s < ncols; s++) {
        const TableChunk& tc = rg.chunks[tcol[qcol_of_slot[s]]];
        if (!tc.present) continue;
        // page whose first_row == row0
        uint32_t lo = aligned ? uint32_t(i) : 0, hi = aligned ? uint32_t(i) + 1 : tc.pages.n_pages;
        while (hi - lo > 1) {
          uint32_t mid = (lo + hi) / 2;
          if (table->pages[tc.pages.first_page + mid].first_row <= it.row0) lo = mid; else hi = mid;
        }
        it.page[s] = tc.pages.first_page + lo;
        const DevPage& pg = table->pages[it.page[s]];
        if (pg.first_row != it.row0 || pg.num_rows != it.nrows || !(pg.flags & 1u)) fast = false;
      }
      it.fast = fast && it.nrows ? 1u : 0u;
      n_fast_items += it.fast;
      items.push_back(it);
    }
  }
  plan.n_items = uint32_t(items.size());
  metrics.bytes_scanned = scanned_bytes;
  // an aggregated column whose footers promise null_count == 0 in every row group read: its
  // non-null counter equals the group's row count, so the scan skips that atomic
  std::vector<uint8_t> nn_is_rows(kMaxAggs, 0);
  for (uint32_t a = 0; a < d.n_aggs; a++) {
    DevAgg& ag = plan.aggs[a];
    if (ag.fn == AG_COUNT_STAR || col_has_nulls[ag.col]) continue;
    ag.update_nn = 0;
    nn_is_rows[a] = 1;
  }

  // columns whose dictionary indices the row phase needs (GROUP BY keys, aggregate inputs)
  for (uint32_t k = 0; k < d.n_group_by; k++) plan.cols[slot_of[d.group_by[k]]].need_idx = 1;
  for (uint32_t a = 0; a < d.n_aggs; a++)
    if (d.aggs[a].fn != PQ_AGG_COUNT_STAR) plan.cols[slot_of[d.aggs[a].col]].need_idx = 1;

  // GROUP BY keys
  plan.nkeys = d.n_group_by;
  for (uint32_t k = 0; k < d.n_group_by; k++) {
    DevKey& key = plan.keys[k];
    key.col = uint8_t(slot_of[d.group_by[k]]);
    uint8_t kind = plan.cols[key.col].kind;
    key.kind = kind == DK_BOOL ? KK_BOOL : KK_DICT_LUT;
    if (key.kind == KK_DICT_LUT && (plan.cols[key.col].has_plain || plan.cols[key.col].has_delta))
      throw Error(PQ_ERR_UNSUPPORTED, "GROUP BY column '" + table->columns[tcol[d.group_by[k]]].name + "' has PLAIN (dictionary-fallback) pages; only dictionary-encoded keys are on the GPU path");
  }

  mark("chunks + items built");
  // ---- shared-memory layout ----
  SmemLayout L{};
  uint32_t off = align_up(uint32_t(sizeof(ScanCtl)), 128);
  for (uint32_t s = 0; s < ncols; s++) {
    // window = bytes of one slab at the widest index + one header per 8 values + alignment slop;
    // anything denser makes the kernel shrink the slab (always correct, only slower)
    L.defwin_cap[s] = plan.cols[s].max_def ? align_up(kSlabRows / 8 + kSlabRows / 16 + 64, 16) : 0;
    L.valwin_cap[s] = plan.cols[s].has_dict ? valwin_cap_for_bw(plan.cols[s].max_bw) : 0;
    // the slab index holds window-relative bit offsets: stage at least the window it was built for
    if (n_fast_items && plan.cols[s].has_dict) L.valwin_cap[s] = std::max(L.valwin_cap[s], table->col_valwin_cap[tcol[qcol_of_slot[s]]]);
    if (plan.cols[s].has_delta) L.valwin_cap[s] = std::max<uint32_t>(L.valwin_cap[s], align_up(kDeltaWindowBytes, 16));
    for (int b = 0; b < 2; b++) { L.defwin[s][b] = off; off += align_up(L.defwin_cap[s] + 16, 128); }
    for (int b = 0; b < 2; b++) { L.valwin[s][b] = off; off += align_up(L.valwin_cap[s] + 16, 128); }
    L.valid[s] = off; off += align_up((kSlabWords + 2) * 4, 16);
    L.rank[s] = off; off += kSlabWords * 4;
    // staging: u32 dictionary indices, or i64 values of DELTA_BINARY_PACKED pages
    L.idx[s] = (plan.cols[s].has_dict || plan.cols[s].has_delta) ? off : 0;
    off += plan.cols[s].has_delta ? kSlabRows * 8 : (plan.cols[s].has_dict ? kSlabRows * 4 : 0);
    L.defdir[s] = off; off += kMaxDirEntries * sizeof(DirEntry);
    for (int b = 0; b < 2; b++) {   // bulk-copy destination for prebuilt directories: 16-byte aligned
      off = align_up(off, 16);
      L.valdir[s][b] = off;
      off += std::max<uint32_t>(kMaxDirEntries * sizeof(DirEntry), plan.cols[s].has_delta ? kMaxDeltaEntries * sizeof(DeltaEntry) : 0);
    }
  }
  off = align_up(off, 16);
  L.recs = off; off += uint32_t(kRecBatch * std::max<uint32_t>(ncols, 1) * sizeof(DevSlabRec));
  L.leafT = off; off += std::max<uint32_t>(nleaves, 1) * kLeafWords * 4;
  L.sel = off; off += kSlabWords * 4;
  L.lutc = off; if (plan.fast_and) off += nleaves * kLutCacheBytes;
  off = align_up(off, 128);
  L.acc = off;
  const uint32_t smem_fixed = off;

  // ---- device side tables ----
  Timer t_all, t_scan;
  PQB_CUDA(cudaEventRecord(t_all.a, stream));
  DevBuf<DevChunk> d_chunks; d_chunks.upload(chunks, stream);
  DevBuf<DevItem> d_items; d_items.upload(items, stream);
  DevBuf<uint8_t> d_lit; d_lit.upload(lit_pool, stream);
  DevBuf<uint64_t> d_ent; d_ent.alloc(std::max<uint64_t>(total_entries, 1), stream);
  for (uint32_t l = 0; l < nleaves; l++) plan.leaves[l].lut_off = uint32_t(uint64_t(l) * total_entries);
  if (uint64_t(nleaves) * total_entries > 0xfffffff0ull) throw Error(PQ_ERR_UNSUPPORTED, "leaf LUTs too large");
  DevBuf<uint8_t> d_luts; d_luts.alloc(std::max<uint64_t>(uint64_t(nleaves) * total_entries, 16), stream);
  DevBuf<uint32_t> d_gid; d_gid.alloc(std::max<uint64_t>(uint64_t(d.n_group_by) * total_entries, 4), stream);
  DevBuf<unsigned long long> d_counters; d_counters.alloc(8, stream); d_counters.zero();
  DevBuf<uint8_t> d_colkind, d_colneeds;
  {
    std::vector<uint8_t> ck(std::max<uint32_t>(ncols, 1)), cn(std::max<uint32_t>(ncols, 1));
    for (uint32_t s = 0; s < ncols; s++) { ck[s] = plan.cols[s].kind; cn[s] = col_needs_ent[s]; }
    d_colkind.upload(ck, stream);
    d_colneeds.upload(cn, stream);
  }
  metrics.h2d_bytes += chunks.size() * sizeof(DevChunk) + items.size() * sizeof(DevItem) + lit_pool.size();

  DevPrepArgs pa{};
  pa.arena = table->d_arena;
  pa.chunks = d_chunks.p;
  pa.n_chunks = nrg * ncols;
  pa.ncols = ncols;
  pa.ent_off = d_ent.p;
  pa.luts = d_luts.p;
  pa.gid_luts = d_gid.p;
  pa.lit_pool = d_lit.p;
  pa.counters = d_counters.p;

  uint64_t launches = 0;
  bool any_ent = false;
  for (uint32_t s = 0; s < ncols; s++) any_ent |= col_needs_ent[s] != 0;
  uint32_t max_dict_n = 1;
  for (const DevChunk& c : chunks) max_dict_n = std::max(max_dict_n, c.dict_n);
  if (nrg && ncols && any_ent) {
    k_dict_entry_offsets<<<(pa.n_chunks + 3) / 4, 128, 0, stream>>>(pa, d_colkind.p, d_colneeds.p);
    launches++;
  }
  if (nrg && ncols && nleaves) {
    bool any_lut = false;
    for (uint32_t l = 0; l < nleaves; l++) any_lut |= plan.leaves[l].kind == LK_CMP || plan.leaves[l].kind == LK_LIKE;
    if (any_lut) {
      dim3 grid(pa.n_chunks, std::min<uint32_t>((max_dict_n + 255) / 256, 64));
      k_leaf_luts<<<grid, 256, 0, stream>>>(pa, plan);
      launches++;
    }
  }

  // ---- GROUP BY key interning ----
  // Local: every dictionary entry of a key column is interned into a device hash table and gets a
  // dense id.  The distinct values are then packed to the host (they are also the output key
  // dictionary).  Multi-GPU: the packed sets are all-gathered and numbered identically on every
  // rank (rank order, first occurrence), local ids are remapped, so partial tables are slot-aligned
  // for one ncclAllReduce.
  struct KeyBufs { DevBuf<unsigned long long> slots; DevBuf<uint32_t> gid_of_slot, rep, counter; uint32_t cap = 0; };
  struct KeyDict { std::vector<uint32_t> offs; std::vector<uint8_t> bytes; };
  std::vector<std::unique_ptr<KeyBufs>> keybufs(d.n_group_by);
  std::vector<uint32_t> key_card(d.n_group_by, 0);
  std::vector<KeyDict> kd(d.n_group_by);
  const bool multi = agg_kernel && (d.flags & PQ_QUERY_ALLREDUCE) && comm_active() && comm_nranks() > 1;
  if ((d.flags & PQ_QUERY_ALLREDUCE) && !comm_active()) throw Error(PQ_ERR_INVALID_ARG, "PQ_QUERY_ALLREDUCE without pq_comm_init_rank");
  for (uint32_t k = 0; agg_kernel && k < d.n_group_by; k++) {
    DevKey& key = plan.keys[k];
    key.gid_off = uint32_t(uint64_t(k) * total_entries);
    if (key.kind == KK_BOOL) { key_card[k] = 2; continue; }
    const uint8_t kkind = plan.cols[key.col].kind;
    uint32_t card_l = 0;
    uint32_t maxn = 1;
    if (nrg) {
      uint64_t sumn = 0;
      for (uint32_t gi = 0; gi < nrg; gi++) { uint32_t n = chunks[size_t(gi) * ncols + key.col].dict_n; maxn = std::max(maxn, n); sumn += n; }
      uint64_t cap = 64;
      while (cap < 4ull * maxn) cap <<= 1;
      for (;;) {
        auto kb = std::make_unique<KeyBufs>();
        kb->cap = uint32_t(cap);
        kb->slots.alloc(cap, stream); kb->slots.zero();
        kb->gid_of_slot.alloc(cap, stream);
        kb->rep.alloc(cap, stream);
        kb->counter.alloc(2, stream); kb->counter.zero();
        DevKeyTable t{kb->slots.p, kb->gid_of_slot.p, kb->rep.p, kb->counter.p, uint32_t(cap - 1), key.col, key.gid_off, kkind};
        dim3 grid(nrg, std::min<uint32_t>((maxn + 255) / 256, 64));
        k_key_intern<<<grid, 256, 0, stream>>>(pa, t, 0);
        k_key_intern<<<grid, 256, 0, stream>>>(pa, t, 1);
        launches += 2;
        uint32_t cnt[2];
        PQB_CUDA(cudaMemcpyAsync(cnt, kb->counter.p, 8, cudaMemcpyDeviceToHost, stream));
        PQB_CUDA(cudaStreamSynchronize(stream));
        metrics.d2h_bytes += 8;
        if (cnt[1] == 1 || cnt[0] * 2ull > cap) {  // table too full: grow and redo
          if (cap > (1ull << 28) || cap > 4 * sumn + 64) throw Error(PQ_ERR_UNSUPPORTED, "GROUP BY key table overflow");
          cap <<= 2;
          continue;
        }
        if (cnt[1]) throw Error(PQ_ERR_CUDA, "group key lookup failed");
        card_l = cnt[0];
        keybufs[k] = std::move(kb);
        break;
      }
    }
    // pack the local distinct values: lengths -> offsets (host) -> bytes
    KeyDict loc;
    loc.offs.assign(size_t(card_l) + 1, 0);
    if (card_l) {
      DevBuf<uint32_t> lens;
      lens.alloc(card_l, stream);
      k_key_lens<<<(card_l + 255) / 256, 256, 0, stream>>>(table->d_arena, d_ent.p, keybufs[k]->rep.p, card_l, kkind, lens.p);
      std::vector<uint32_t> hl(card_l);
      PQB_CUDA(cudaMemcpyAsync(hl.data(), lens.p, card_l * 4ull, cudaMemcpyDeviceToHost, stream));
      PQB_CUDA(cudaStreamSynchronize(stream));
      uint64_t tot = 0;
      for (uint32_t g = 0; g < card_l; g++) { loc.offs[g] = uint32_t(tot); tot += hl[g]; }
      loc.offs[card_l] = uint32_t(tot);
      if (tot > 0x7fffffffull) throw Error(PQ_ERR_UNSUPPORTED, "group key strings exceed 2 GiB");
      DevBuf<uint32_t> doffs;
      doffs.upload(loc.offs, stream);
      DevBuf<uint8_t> dbytes;
      dbytes.alloc(std::max<uint64_t>(tot, 1), stream);
      k_key_bytes<<<card_l, 64, 0, stream>>>(table->d_arena, d_ent.p, keybufs[k]->rep.p, card_l, kkind, doffs.p, dbytes.p);
      launches += 2;
      loc.bytes.resize(tot);
      if (tot) PQB_CUDA(cudaMemcpyAsync(loc.bytes.data(), dbytes.p, tot, cudaMemcpyDeviceToHost, stream));
      PQB_CUDA(cudaStreamSynchronize(stream));
      metrics.d2h_bytes += card_l * 4ull + tot;
    }
    if (!multi) {
      key_card[k] = card_l;
      kd[k] = std::move(loc);
      continue;
    }
    // ---- multi-GPU: agree on one numbering ----
    const int nr = comm_nranks(), me = comm_rank();
    std::vector<unsigned long long> sizes(size_t(nr) * 2);
    {
      unsigned long long mine[2] = {card_l, loc.bytes.size()};
      DevBuf<unsigned long long> dsend, drecv;
      dsend.alloc(2, stream);
      drecv.alloc(size_t(nr) * 2, stream);
      PQB_CUDA(cudaMemcpyAsync(dsend.p, mine, 16, cudaMemcpyHostToDevice, stream));
      comm_allgather_bytes(dsend.p, drecv.p, 16, stream);
      PQB_CUDA(cudaMemcpyAsync(sizes.data(), drecv.p, size_t(nr) * 16, cudaMemcpyDeviceToHost, stream));
      PQB_CUDA(cudaStreamSynchronize(stream));
    }
    unsigned long long cardmax = 0, bytesmax = 0;
    for (int r = 0; r < nr; r++) { cardmax = std::max(cardmax, sizes[2 * r]); bytesmax = std::max(bytesmax, sizes[2 * r + 1]); }
    const size_t per_rank = ((4 * (cardmax + 1) + bytesmax) + 15) & ~size_t(15);
    std::vector<uint8_t> sendbuf(per_rank, 0), recvbuf(per_rank * nr);
    std::memcpy(sendbuf.data(), loc.offs.data(), loc.offs.size() * 4);
    if (!loc.bytes.empty()) std::memcpy(sendbuf.data() + 4 * (cardmax + 1), loc.bytes.data(), loc.bytes.size());
    {
      DevBuf<uint8_t> dsend, drecv;
      dsend.upload(sendbuf, stream);
      drecv.alloc(per_rank * nr, stream);
      comm_allgather_bytes(dsend.p, drecv.p, per_rank, stream);
      PQB_CUDA(cudaMemcpyAsync(recvbuf.data(), drecv.p, per_rank * nr, cudaMemcpyDeviceToHost, stream));
      PQB_CUDA(cudaStreamSynchronize(stream));
      metrics.d2h_bytes += per_rank * nr;
      metrics.h2d_bytes += per_rank;
    }
    std::map<std::string, uint32_t> ids;  // identical content + identical insertion order on every rank
    std::vector<uint32_t> remap(std::max<uint32_t>(card_l, 1), 0);
    KeyDict glob;
    glob.offs.push_back(0);
    for (int r = 0; r < nr; r++) {
      const uint8_t* base = recvbuf.data() + per_rank * r;
      const uint32_t* offs = reinterpret_cast<const uint32_t*>(base);
      const uint8_t* bytes = base + 4 * (cardmax + 1);
      for (unsigned long long i = 0; i < sizes[2 * r]; i++) {
        std::string v(reinterpret_cast<const char*>(bytes + offs[i]), offs[i + 1] - offs[i]);
        auto it = ids.find(v);
        if (it == ids.end()) {
          it = ids.emplace(v, uint32_t(ids.size())).first;
          glob.bytes.insert(glob.bytes.end(), v.begin(), v.end());
          glob.offs.push_back(uint32_t(glob.bytes.size()));
        }
        if (r == me) remap[i] = it->second;
      }
    }
    if (card_l) {
      DevBuf<uint32_t> dremap;
      dremap.upload(remap, stream);
      dim3 grid(nrg, std::min<uint32_t>((maxn + 255) / 256, 64));
      k_gid_remap<<<grid, 256, 0, stream>>>(pa, key.col, key.gid_off, dremap.p, card_l);
      launches++;
      PQB_CUDA(cudaStreamSynchronize(stream));
    }
    key_card[k] = uint32_t(ids.size());
    kd[k] = std::move(glob);
  }
  uint64_t nslots64 = 1;
  for (uint32_t k = 0; k < d.n_group_by; k++) {
    plan.keys[k].card = key_card[k];
    plan.keys[k].stride = uint32_t(nslots64);
    nslots64 *= uint64_t(key_card[k]) + 1;
    if (nslots64 > (1ull << 26)) throw Error(PQ_ERR_UNSUPPORTED, "GROUP BY key space too large for the dense accumulator table");
  }
  plan.nslots = uint32_t(nslots64);
  const uint32_t cells = 1 + plan.n_acc + plan.n_nn;

  // ---- accumulators ----
  DevBuf<unsigned long long> d_acc;
  size_t smem_total = smem_fixed;
  if (agg_kernel) {
    d_acc.alloc(size_t(plan.nslots) * cells, stream);
    k_acc_init<<<std::min<uint32_t>(1024, (plan.nslots * cells + 255) / 256), 256, 0, stream>>>(d_acc.p, plan.nslots, plan.n_acc, cells, plan);
    launches++;
    size_t acc_bytes = size_t(plan.nslots) * cells * 8;
    if (smem_fixed + acc_bytes + 1024 <= ctx.smem_optin()) { plan.smem_acc = 1; smem_total = smem_fixed + acc_bytes; }
  }
  L.total = uint32_t(smem_total);
  if (smem_total > ctx.smem_optin()) throw Error(PQ_ERR_UNSUPPORTED, "query needs more shared memory than one SM has");

  // ---- selection bitmap / counts ----
  const bool want_rows = !has_aggs && !(d.flags & PQ_QUERY_COUNT_ONLY);
  plan.write_bitmap = want_rows ? 1 : 0;
  DevBuf<uint32_t> d_bitmap, d_item_counts;
  if (want_rows) { d_bitmap.alloc(std::max<uint32_t>(bitmap_words, 1), stream); d_bitmap.zero(); }
  d_item_counts.alloc(std::max<size_t>(items.size(), 1), stream);
  if (want_rows) algo_bytes += metrics.rows_scanned / 8;
  metrics.algorithmic_bytes = algo_bytes;

  mark("prep kernels queued");
  // ---- the fused scan ----
  DevScanArgs sa{};
  sa.arena = table->d_arena;
  sa.pages = table->d_pages;
  sa.chunks = d_chunks.p;
  sa.items = d_items.p;
  sa.luts = d_luts.p;
  sa.gid_luts = d_gid.p;
  sa.lit_pool = d_lit.p;
  sa.bitmap = d_bitmap.p;
  sa.item_counts = d_item_counts.p;
  sa.acc = d_acc.p;
  sa.counters = d_counters.p;
  sa.slab_recs = table->d_slab_recs;
  sa.slab_dirs = table->d_slab_dirs;
  if (!items.empty()) {
    PQB_CUDA(cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ctx.smem_optin())));
    int occ = 1;
    PQB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scan, kScanThreads, smem_total));
    if (occ < 1) occ = 1;
    uint32_t grid = std::min<uint32_t>(uint32_t(items.size()), uint32_t(ctx.sm_count() * occ));
    if (const char* g = getenv("PQB_GRID")) grid = std::max(1, atoi(g));   // debugging aid: forces several items per CTA
    if (getenv("PQB_VERBOSE"))
      fprintf(stderr, "[pqb] k_scan: %u CTAs x %d threads, %zu B smem/CTA, %d CTAs/SM, %zu items\n", grid, kScanThreads,
              size_t(smem_total), occ, items.size());
    if (verbose) fprintf(stderr, "[pqb] %u of %zu items covered by the slab index\n", n_fast_items, items.size());
    PQB_CUDA(cudaEventRecord(t_scan.a, stream));
    k_scan<<<grid, kScanThreads, smem_total, stream>>>(plan, L, sa);
    PQB_CUDA(cudaEventRecord(t_scan.b, stream));
    PQB_CUDA(cudaGetLastError());
    launches++;
  }

  if (getenv("PQB_DEBUG_ITEMS")) {
    std::vector<uint32_t> ic(items.size());
    PQB_CUDA(cudaMemcpyAsync(ic.data(), d_item_counts.p, ic.size() * 4, cudaMemcpyDeviceToHost, stream));
    PQB_CUDA(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < ic.size(); i++)
      fprintf(stderr, "item %zu rg %u row0 %u nrows %u g0 %llu count %u\n", i, items[i].rg, items[i].row0, items[i].nrows,
              (unsigned long long)items[i].global_row0, ic[i]);
  }
  mark("scan queued");
  unsigned long long h_counters[4] = {0, 0, 0, 0};
  PQB_CUDA(cudaMemcpyAsync(h_counters, d_counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, stream));
  metrics.d2h_bytes += sizeof(h_counters);

  // ---- results ----
  const uint32_t batch_rows = d.batch_size ? d.batch_size : 20000;
  if (agg_kernel) {
    DevBuf<uint32_t> d_out_count, d_out_slot;
    DevBuf<unsigned long long> d_out_cells;
    uint32_t out_cap = plan.nslots;
    d_out_count.alloc(1, stream); d_out_count.zero();
    d_out_slot.alloc(out_cap, stream);
    d_out_cells.alloc(size_t(out_cap) * cells, stream);
    // multi-GPU: one all-reduce of the partial tables
    if (d.flags & PQ_QUERY_ALLREDUCE) {
      comm_allreduce_u64(d_acc.p, plan.nslots, 0, stream);
      for (uint32_t a = 0; a < plan.n_acc; a++) {
        uint8_t how = plan.acc_init[a];
        comm_allreduce_u64(d_acc.p + size_t(1 + a) * plan.nslots, plan.nslots, how == 0 ? 0 : how == 1 ? 3 : how == 2 ? 1 : 2, stream);
      }
      for (uint32_t k = 0; k < plan.n_nn; k++) comm_allreduce_u64(d_acc.p + size_t(1 + plan.n_acc + k) * plan.nslots, plan.nslots, 0, stream);
    }
    k_agg_compact<<<std::min<uint32_t>(512, (plan.nslots + 255) / 256), 256, 0, stream>>>(d_acc.p, plan.nslots, cells, d_out_count.p, d_out_slot.p, d_out_cells.p, out_cap);
    launches++;
    uint32_t n_out = 0;
    PQB_CUDA(cudaMemcpyAsync(&n_out, d_out_count.p, 4, cudaMemcpyDeviceToHost, stream));
    PQB_CUDA(cudaStreamSynchronize(stream));
    if (h_counters[1]) throw Error(PQ_ERR_CORRUPT, "corrupt or unsupported page encoding met on the device (code " + std::to_string(h_counters[1]) + ")");
    std::vector<uint32_t> out_slot(n_out);
    std::vector<unsigned long long> out_cells(size_t(n_out) * cells);
    if (n_out) {
      PQB_CUDA(cudaMemcpyAsync(out_slot.data(), d_out_slot.p, n_out * 4, cudaMemcpyDeviceToHost, stream));
      for (uint32_t c = 0; c < cells; c++)
        PQB_CUDA(cudaMemcpyAsync(out_cells.data() + size_t(c) * n_out, d_out_cells.p + size_t(c) * out_cap, n_out * 8ull, cudaMemcpyDeviceToHost, stream));
    }
    PQB_CUDA(cudaEventRecord(t_all.b, stream));
    PQB_CUDA(cudaStreamSynchronize(stream));
    metrics.d2h_bytes += uint64_t(n_out) * (4 + 8ull * cells);

    // SQL: a global aggregate over zero rows still yields one row
    bool synth_empty = d.n_group_by == 0 && n_out == 0;
    uint32_t n_rows = synth_empty ? 1 : n_out;
    metrics.groups = n_rows;
    unsigned long long rows_sel = 0;
    for (uint32_t i = 0; i < n_out; i++) rows_sel += out_cells[i];
    metrics.rows_selected = rows_sel;

    // deterministic order: ascending dense slot (= mixed radix of group ids)
    std::vector<uint32_t> order(n_out);
    for (uint32_t i = 0; i < n_out; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return out_slot[a] < out_slot[b]; });

    for (uint32_t r0 = 0; r0 < n_rows; r0 += batch_rows) {
      uint32_t nb = std::min(batch_rows, n_rows - r0);
      OutBatch ob;
      ob.rows = nb;
      for (uint32_t k = 0; k < d.n_group_by; k++) {
        OutColumn oc;
        uint32_t qc = uint32_t(d.group_by[k]);
        oc.name = d.columns[qc].name;
        oc.type = out_type_of(qc);
        uint8_t kind = plan.cols[plan.keys[k].col].kind;
        oc.validity.assign((nb + 7) / 8, 0);
        if (kind == DK_STR) oc.offsets.push_back(0);
        else if (kind == DK_BOOL) oc.values.assign((nb + 7) / 8, 0);
        else oc.values.resize(size_t(nb) * 8);
        for (uint32_t i = 0; i < nb; i++) {
          uint32_t slot = out_slot[order[r0 + i]];
          uint32_t gid = (slot / plan.keys[k].stride) % (key_card[k] + 1);
          bool valid = gid != key_card[k];
          if (valid) oc.validity[i >> 3] |= uint8_t(1u << (i & 7)); else oc.null_count++;
          if (kind == DK_STR) {
            if (valid) oc.values.insert(oc.values.end(), kd[k].bytes.begin() + kd[k].offs[gid], kd[k].bytes.begin() + kd[k].offs[gid + 1]);
            oc.offsets.push_back(int32_t(oc.values.size()));
          } else if (kind == DK_BOOL) {
            if (valid && gid) oc.values[i >> 3] |= uint8_t(1u << (i & 7));
          } else if (valid) {
            std::memcpy(oc.values.data() + size_t(i) * 8, kd[k].bytes.data() + kd[k].offs[gid], 8);
          }
        }
        if (!oc.null_count) oc.validity.clear();
        ob.cols.push_back(std::move(oc));
      }
      for (uint32_t a = 0; a < d.n_aggs; a++) {
        const DevAgg& ag = plan.aggs[a];
        OutColumn oc;
        static const char* fn_names[] = {"count(*)", "count", "sum", "min", "max", "avg"};
        oc.name = ag.fn == AG_COUNT_STAR ? "count(*)" : std::string(fn_names[ag.fn]) + "(" + d.columns[d.aggs[a].col].name + ")";
        oc.type = agg_out_type[a];
        oc.values.resize(size_t(nb) * 8);
        oc.validity.assign((nb + 7) / 8, 0);
        for (uint32_t i = 0; i < nb; i++) {
          unsigned long long rows = 0, nn = 0, cell = 0;
          if (!synth_empty) {
            uint32_t o = order[r0 + i];
            rows = out_cells[o];
            if (ag.fn != AG_COUNT_STAR) nn = nn_is_rows[a] ? rows : out_cells[size_t(1 + plan.n_acc + ag.nn_slot) * n_out + o];
            if (ag.fn >= AG_SUM) cell = out_cells[size_t(1 + ag.acc_slot) * n_out + o];
          }
          bool valid = true;
          uint64_t v = 0;
          switch (ag.fn) {
            case AG_COUNT_STAR: v = rows; break;
            case AG_COUNT: v = nn; break;
            case AG_SUM: valid = nn > 0; v = cell; break;
            case AG_AVG: valid = nn > 0; if (valid) v = f64_bits(bits_f64(cell) / double(nn)); break;
            default:  // MIN / MAX
              valid = nn > 0;
              v = ag.kind == DK_F64 ? f64_from_order_key((int64_t)cell) : cell;
          }
          if (valid) { oc.validity[i >> 3] |= uint8_t(1u << (i & 7)); std::memcpy(oc.values.data() + size_t(i) * 8, &v, 8); }
          else oc.null_count++;
        }
        if (!oc.null_count) oc.validity.clear();
        ob.cols.push_back(std::move(oc));
      }
      batches_.push_back(std::move(ob));
    }
  } else {
    // ---- filter / COUNT(*) ----
    // bitmap-driven stream compaction on the device: per-item prefix, then one CTA per item
    DevBuf<unsigned long long> d_item_base, d_total, d_ids;
    std::shared_ptr<PinnedBlock> ids_block;   // selected row ordinals land in page-locked memory, batches alias it
    unsigned long long n_ids = 0;
    if (want_rows && !items.empty()) {
      if (d.n_projection && !(d.flags & PQ_QUERY_EMIT_ROW_IDS))
        throw Error(PQ_ERR_UNSUPPORTED, "projection of column values is not on the GPU path yet: ask for PQ_QUERY_EMIT_ROW_IDS or PQ_QUERY_COUNT_ONLY");
      d_item_base.alloc(items.size(), stream);
      d_total.alloc(1, stream);
      k_item_prefix<<<1, 1024, 0, stream>>>(d_item_counts.p, uint32_t(items.size()), d_item_base.p, d_total.p);
      launches++;
      unsigned long long total = 0;
      PQB_CUDA(cudaMemcpyAsync(&total, d_total.p, 8, cudaMemcpyDeviceToHost, stream));
      PQB_CUDA(cudaStreamSynchronize(stream));
      mark("scan done, selected-row total on host");
      metrics.d2h_bytes += 8;
      unsigned long long keep = total;
      if (d.limit >= 0 && (unsigned long long)d.limit < keep) keep = (unsigned long long)d.limit;
      if (keep) {
        d_ids.alloc(keep, stream);
        uint32_t grid = std::min<uint32_t>(uint32_t(items.size()), uint32_t(ctx.sm_count() * 8));
        k_compact_row_ids<<<grid, 256, 0, stream>>>(d_bitmap.p, d_items.p, d_item_counts.p, d_item_base.p, uint32_t(items.size()), d_ids.p, keep);
        launches++;
        ids_block = std::make_shared<PinnedBlock>();
        ids_block->p = ctx.pinned_acquire(keep * 8);
        ids_block->bytes = keep * 8;
        n_ids = keep;
        PQB_CUDA(cudaMemcpyAsync(ids_block->p, d_ids.p, keep * 8, cudaMemcpyDeviceToHost, stream));
        metrics.d2h_bytes += keep * 8;
      }
    }
    PQB_CUDA(cudaEventRecord(t_all.b, stream));
    PQB_CUDA(cudaStreamSynchronize(stream));
    mark("results on host");
    if (h_counters[1]) throw Error(PQ_ERR_CORRUPT, "corrupt or unsupported page encoding met on the device (code " + std::to_string(h_counters[1]) + ")");
    metrics.rows_selected = h_counters[0];
    if (has_aggs) {  // SELECT COUNT(*) [, COUNT(*)...] WHERE ...
      unsigned long long total = h_counters[0];
      if (d.flags & PQ_QUERY_ALLREDUCE) {
        if (!comm_active()) throw Error(PQ_ERR_INVALID_ARG, "PQ_QUERY_ALLREDUCE without pq_comm_init_rank");
        comm_allreduce_u64(d_counters.p, 1, 0, stream);
        PQB_CUDA(cudaMemcpyAsync(&total, d_counters.p, 8, cudaMemcpyDeviceToHost, stream));
        PQB_CUDA(cudaStreamSynchronize(stream));
      }
      OutBatch ob;
      ob.rows = 1;
      for (uint32_t a = 0; a < d.n_aggs; a++) {
        OutColumn oc;
        oc.name = "count(*)";
        oc.type = PQ_T_I64;
        oc.values.resize(8);
        std::memcpy(oc.values.data(), &total, 8);
        ob.cols.push_back(std::move(oc));
      }
      metrics.groups = 1;
      batches_.push_back(std::move(ob));
    } else if (want_rows) {
      // selected row ordinals, ascending (projection of column VALUES is the next widening step)
      for (size_t r0 = 0; r0 < n_ids || (r0 == 0 && n_ids == 0); r0 += batch_rows) {
        size_t nb = std::min<size_t>(batch_rows, n_ids - r0);
        OutBatch ob;
        ob.rows = int64_t(nb);
        OutColumn oc;
        oc.name = "__row_id";
        oc.type = PQ_T_I64;
        if (nb) { oc.ext = ids_block; oc.ext_off = r0 * 8; }
        ob.cols.push_back(std::move(oc));
        batches_.push_back(std::move(ob));
        if (n_ids == 0) break;
      }
    }
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, t_all.a, t_all.b);
  metrics.device_ms = ms;
  if (!items.empty()) { cudaEventElapsedTime(&ms, t_scan.a, t_scan.b); metrics.scan_kernel_ms = ms; }
  metrics.kernel_launches = launches;
}

void Query::schema(ArrowSchema* out) const {
  if (!batches_.empty()) { export_batch(batches_[0], nullptr, out); return; }
  OutBatch none;
  export_batch(none, nullptr, out);
}

int Query::next(int partition, ArrowArray* out, ArrowSchema* schema) {
  (void)partition;
  if (next_batch_ >= batches_.size()) return PQ_END_OF_STREAM;
  export_batch(batches_[next_batch_++], out, schema);
  return PQ_OK;
}

}  // namespace pqb
