/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the
 * semantics of Parseable's query hot path, used solely as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs.  Nothing in parseable_b200/ links, imports or calls this.
 *
 * The arithmetic of the reference path lives in third-party crates that are
 * NOT under /root/reference (Cargo.lock: datafusion 53.1.0 :1510-1511,
 * arrow-ord / arrow-string / arrow-select 58.1.0 :670-722, parquet 58.1.0
 * :3843-3844), and there is no Rust toolchain here, so this file restates their
 * published behaviour and is anchored on the reference's call sites:
 *   - FilterExec over the pushed predicate ........ src/query/stream_schema_provider.rs:122-151
 *   - AggregateExec driven by collect_partitioned .. src/query/mod.rs:287-291
 *   - GROUP BY + COUNT(*) known answers ............ src/storage/field_stats.rs:298-330, 927-1327
 * PARITY PINNING: grouped COUNT(*) (incl. the NULL group and Int64 counts) is
 * pinned by the field_stats.rs fixtures re-created in tests/golden/.  SUM / MIN /
 * MAX / f64 accumulation / comparison and LIKE results are "parity unpinned" by
 * any reference test; for those this oracle is cross-checked
 * against an independent engine (pyarrow/Acero) in tests/test_oracle.py.
 *
 * Rules restated:
 *   - SQL three-valued logic: a NULL input makes a comparison NULL; AND/OR are
 *     Kleene; only TRUE rows are selected.
 *   - Float compare and MIN/MAX use IEEE-754 totalOrder (NaN == NaN, NaN greatest,
 *     -0.0 < +0.0).
 *   - COUNT -> Int64; COUNT(col) skips NULLs; SUM(Int64) wraps; SUM(Float64) adds
 *     sequentially in row order; MIN/MAX/SUM over no non-NULL input -> NULL;
 *     AVG = SUM(as f64)/COUNT.
 *   - NULL is its own GROUP BY key; float keys group by bit pattern.
 *   - LIKE: '%' any run, '_' exactly one character, ESCAPE '\'; ILIKE folds ASCII case
 *     (documented limitation: the reference folds full Unicode).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- leaf predicates: write T[i] (1 = TRUE) and N[i] (1 = NULL) ---- */
static int cmp_res(int c, int op) {
  switch (op) { case 0: return c == 0; case 1: return c != 0; case 2: return c < 0; case 3: return c <= 0;
                case 4: return c > 0; default: return c >= 0; }
}

void or_cmp_i64(const int64_t* v, const uint8_t* valid, int64_t n, int op, int64_t lit, uint8_t* T, uint8_t* N) {
  for (int64_t i = 0; i < n; i++) {
    if (valid && !valid[i]) { T[i] = 0; N[i] = 1; continue; }
    int c = v[i] < lit ? -1 : (v[i] > lit ? 1 : 0);
    T[i] = (uint8_t)cmp_res(c, op); N[i] = 0;
  }
}

/* totalOrder key: flip the magnitude bits of negative values (arrow-ord's f64 total_cmp) */
static int64_t total_key(double d) {
  int64_t b; memcpy(&b, &d, 8);
  return b ^ (int64_t)(((uint64_t)(b >> 63)) >> 1);
}

void or_cmp_f64(const double* v, const uint8_t* valid, int64_t n, int op, double lit, uint8_t* T, uint8_t* N) {
  int64_t kl = total_key(lit);
  for (int64_t i = 0; i < n; i++) {
    if (valid && !valid[i]) { T[i] = 0; N[i] = 1; continue; }
    int64_t k = total_key(v[i]);
    int c = k < kl ? -1 : (k > kl ? 1 : 0);
    T[i] = (uint8_t)cmp_res(c, op); N[i] = 0;
  }
}

void or_cmp_str(const int32_t* off, const uint8_t* data, const uint8_t* valid, int64_t n, int op,
                const uint8_t* lit, int32_t litlen, uint8_t* T, uint8_t* N) {
  for (int64_t i = 0; i < n; i++) {
    if (valid && !valid[i]) { T[i] = 0; N[i] = 1; continue; }
    int32_t len = off[i + 1] - off[i];
    int32_t m = len < litlen ? len : litlen;
    int c = memcmp(data + off[i], lit, (size_t)m);
    if (c == 0) c = len < litlen ? -1 : (len > litlen ? 1 : 0);
    T[i] = (uint8_t)cmp_res(c, op); N[i] = 0;
  }
}

/* LIKE by dynamic programming over (pattern token, string position) — deliberately a different
 * algorithm from the device's two-pointer matcher. */
static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

static int like_dp(const uint8_t* s, int32_t n, const uint8_t* p, int32_t m, int ci) {
  /* tokenise the pattern */
  int16_t* tok = (int16_t*)malloc(sizeof(int16_t) * (size_t)(m + 1));
  int32_t nt = 0;
  for (int32_t i = 0; i < m; i++) {
    if (p[i] == '\\' && i + 1 < m) tok[nt++] = p[++i];
    else if (p[i] == '%') tok[nt++] = -1;
    else if (p[i] == '_') tok[nt++] = -2;
    else tok[nt++] = p[i];
  }
  /* character start positions of the utf-8 string */
  int32_t* cs = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 2));
  int32_t nc = 0;
  for (int32_t i = 0; i < n; i++) if ((s[i] & 0xc0) != 0x80) cs[nc++] = i;
  cs[nc] = n;
  /* reach[j] : pattern tokens [0,t) can match bytes [0,j) ; positions are byte offsets */
  uint8_t* cur = (uint8_t*)calloc((size_t)n + 2, 1);
  uint8_t* nxt = (uint8_t*)calloc((size_t)n + 2, 1);
  cur[0] = 1;
  for (int32_t t = 0; t < nt; t++) {
    memset(nxt, 0, (size_t)n + 2);
    if (tok[t] == -1) {
      int seen = 0;
      for (int32_t j = 0; j <= n; j++) { if (cur[j]) seen = 1; if (seen) nxt[j] = 1; }
    } else if (tok[t] == -2) {
      for (int32_t c = 0; c < nc; c++) if (cur[cs[c]]) nxt[cs[c + 1]] = 1;
    } else {
      for (int32_t j = 0; j < n; j++) {
        if (!cur[j]) continue;
        int a = s[j], b = tok[t];
        if (ci) { a = lower(a); b = lower(b); }
        if (a == b) nxt[j + 1] = 1;
      }
    }
    uint8_t* tmp = cur; cur = nxt; nxt = tmp;
  }
  int r = cur[n];
  free(tok); free(cs); free(cur); free(nxt);
  return r;
}

/* flags: 1 negated, 2 case-insensitive */
void or_like(const int32_t* off, const uint8_t* data, const uint8_t* valid, int64_t n, const uint8_t* pat,
             int32_t plen, uint32_t flags, uint8_t* T, uint8_t* N) {
  for (int64_t i = 0; i < n; i++) {
    if (valid && !valid[i]) { T[i] = 0; N[i] = 1; continue; }
    int r = like_dp(data + off[i], off[i + 1] - off[i], pat, plen, (flags & 2) != 0);
    if (flags & 1) r = !r;
    T[i] = (uint8_t)r; N[i] = 0;
  }
}

void or_is_null(const uint8_t* valid, int64_t n, int negate, uint8_t* T, uint8_t* N) {
  for (int64_t i = 0; i < n; i++) {
    int isnull = valid ? !valid[i] : 0;
    T[i] = (uint8_t)(negate ? !isnull : isnull); N[i] = 0;
  }
}

/* Kleene connectives, in place on (Ta, Na) */
void or_and(uint8_t* Ta, uint8_t* Na, const uint8_t* Tb, const uint8_t* Nb, int64_t n) {
  for (int64_t i = 0; i < n; i++) {
    int fa = !Ta[i] && !Na[i], fb = !Tb[i] && !Nb[i];
    int t = Ta[i] && Tb[i];
    int nul = !fa && !fb && (Na[i] || Nb[i]);
    Ta[i] = (uint8_t)t; Na[i] = (uint8_t)nul;
  }
}
void or_or(uint8_t* Ta, uint8_t* Na, const uint8_t* Tb, const uint8_t* Nb, int64_t n) {
  for (int64_t i = 0; i < n; i++) {
    int t = Ta[i] || Tb[i];
    int nul = !t && (Na[i] || Nb[i]);
    Ta[i] = (uint8_t)t; Na[i] = (uint8_t)nul;
  }
}
void or_not(uint8_t* T, const uint8_t* N, int64_t n) {
  for (int64_t i = 0; i < n; i++) T[i] = (uint8_t)(!T[i] && !N[i]);
}
int64_t or_count(const uint8_t* T, int64_t n) {
  int64_t c = 0;
  for (int64_t i = 0; i < n; i++) c += T[i] != 0;
  return c;
}

/* ---- GROUP BY ----
 * keys: nkeys arrays of int64 codes + validity; aggs described by parallel arrays.
 * Output rows are in first-seen order.  Returns the number of groups, or -1 on OOM.
 * agg fn: 0 COUNT(*), 1 COUNT(col), 2 SUM, 3 MIN, 4 MAX, 5 AVG ; agg type: 0 int64, 1 float64 */
typedef struct {
  int64_t* keys;      /* ngroups * nkeys */
  uint8_t* keynull;   /* ngroups * nkeys */
  int64_t ngroups, cap;
  int64_t* table;     /* open addressing: group index + 1 */
  int64_t tcap;
} GroupTable;

static uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

int64_t or_groupby(int64_t nrows, const uint8_t* sel, int nkeys, const int64_t* const* key_codes,
                   const uint8_t* const* key_valid, int naggs, const int* agg_fn, const int* agg_type,
                   const void* const* agg_vals, const uint8_t* const* agg_valid, int64_t max_groups,
                   int64_t* out_keys, uint8_t* out_keynull, int64_t* out_i64, double* out_f64,
                   uint8_t* out_valid /* [naggs][max_groups] */) {
  GroupTable g;
  g.ngroups = 0;
  g.tcap = 1024;
  while (g.tcap < max_groups * 2 + 2) g.tcap <<= 1;
  g.table = (int64_t*)calloc((size_t)g.tcap, sizeof(int64_t));
  if (!g.table) return -1;
  int64_t* nn = (int64_t*)calloc((size_t)(naggs > 0 ? naggs : 1) * (size_t)max_groups, sizeof(int64_t));
  if (!nn) { free(g.table); return -1; }
  for (int a = 0; a < naggs; a++)
    for (int64_t j = 0; j < max_groups; j++) { out_i64[a * max_groups + j] = 0; out_f64[a * max_groups + j] = 0.0; out_valid[a * max_groups + j] = 0; }
  int64_t kbuf[16]; uint8_t nbuf[16];
  for (int64_t i = 0; i < nrows; i++) {
    if (sel && !sel[i]) continue;
    uint64_t h = 0x9e3779b97f4a7c15ULL;
    for (int k = 0; k < nkeys; k++) {
      int isnull = key_valid[k] ? !key_valid[k][i] : 0;
      kbuf[k] = isnull ? 0 : key_codes[k][i];
      nbuf[k] = (uint8_t)isnull;
      h = mix(h ^ (uint64_t)kbuf[k] ^ ((uint64_t)isnull << 63)) + (uint64_t)k;
    }
    int64_t slot = (int64_t)(h & (uint64_t)(g.tcap - 1)), gi = -1;
    for (;;) {
      int64_t e = g.table[slot];
      if (e == 0) {
        if (g.ngroups >= max_groups) { free(g.table); free(nn); return -2; }
        gi = g.ngroups++;
        g.table[slot] = gi + 1;
        for (int k = 0; k < nkeys; k++) { out_keys[gi * nkeys + k] = kbuf[k]; out_keynull[gi * nkeys + k] = nbuf[k]; }
        break;
      }
      int same = 1;
      for (int k = 0; k < nkeys && same; k++)
        same = out_keynull[(e - 1) * nkeys + k] == nbuf[k] && (nbuf[k] || out_keys[(e - 1) * nkeys + k] == kbuf[k]);
      if (same) { gi = e - 1; break; }
      slot = (slot + 1) & (g.tcap - 1);
    }
    for (int a = 0; a < naggs; a++) {
      int64_t o = (int64_t)a * max_groups + gi;
      if (agg_fn[a] == 0) { out_i64[o]++; out_valid[o] = 1; continue; }
      if (agg_valid[a] && !agg_valid[a][i]) continue;
      int first = nn[o] == 0;
      nn[o]++;
      if (agg_fn[a] == 1) { out_i64[o]++; continue; }
      if (agg_type[a] == 0) {
        int64_t v = ((const int64_t*)agg_vals[a])[i];
        switch (agg_fn[a]) {
          case 2: out_i64[o] = (int64_t)((uint64_t)out_i64[o] + (uint64_t)v); break; /* wrapping */
          case 3: if (first || v < out_i64[o]) out_i64[o] = v; break;
          case 4: if (first || v > out_i64[o]) out_i64[o] = v; break;
          case 5: out_f64[o] += (double)v; break;
        }
      } else {
        double v = ((const double*)agg_vals[a])[i];
        switch (agg_fn[a]) {
          case 2: case 5: out_f64[o] += v; break; /* sequential, row order */
          case 3: if (first || total_key(v) < total_key(out_f64[o])) out_f64[o] = v; break;
          case 4: if (first || total_key(v) > total_key(out_f64[o])) out_f64[o] = v; break;
        }
      }
    }
  }
  for (int a = 0; a < naggs; a++)
    for (int64_t j = 0; j < g.ngroups; j++) {
      int64_t o = (int64_t)a * max_groups + j;
      if (agg_fn[a] == 0) { out_valid[o] = 1; continue; }
      if (agg_fn[a] == 1) { out_valid[o] = 1; continue; }
      out_valid[o] = nn[o] > 0;
      if (agg_fn[a] == 5 && nn[o] > 0) out_f64[o] = out_f64[o] / (double)nn[o];
    }
  free(g.table); free(nn);
  return g.ngroups;
}
