// Result assembly on the device: the accumulator table of an aggregate query becomes the buffers of
// its Arrow result batches (values, validity bitmaps, string offsets + bytes of the group keys) in
// ONE device block that is copied to page-locked host memory once; the batches alias that block.
// Replaces, for the reference, AggregateExec(Final)'s output + the RecordBatch construction
// (DataFusion; results are what Query::execute returns, /root/reference/src/query/mod.rs:287-291).
#pragma once
#include <cuda_runtime.h>

#include "decode_core.cuh"
#include "device_structs.hpp"

namespace pqb {

constexpr int kSlotTile = 1024;

// non-empty groups per tile of kSlotTile slots
__global__ void k_slot_tile_counts(const unsigned long long* __restrict__ rows, uint32_t nslots, uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t ws[8];
  const uint32_t s0 = blockIdx.x * kSlotTile;
  uint32_t c = 0;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kSlotTile; i += blockDim.x) c += (s0 + i < nslots && rows[s0 + i] != 0) ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; w++) t += ws[w];
    tile_counts[blockIdx.x] = t;
  }
}

// ascending slot order: out_slot[base[tile] + rank inside the tile]; 256 threads x 4 consecutive slots
__global__ void __launch_bounds__(256) k_slot_compact(const unsigned long long* __restrict__ rows, uint32_t nslots,
                                                      const unsigned long long* __restrict__ tile_base, uint32_t* __restrict__ out_slot) {
  __shared__ uint32_t ws[8];
  const uint32_t s0 = blockIdx.x * kSlotTile + threadIdx.x * 4;
  uint32_t f[4], c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { f[k] = (s0 + k < nslots && rows[s0 + k] != 0) ? 1u : 0u; c += f[k]; }
  uint32_t incl = c;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (uint32_t w = 0; w < warp; w++) wbase += ws[w];
  unsigned long long pos = tile_base[blockIdx.x] + wbase + incl - c;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (f[k]) out_slot[pos++] = s0 + k;
}

struct FinishKey {
  const uint32_t* kd_offs;     // key dictionary: offsets (group-id order)
  const uint8_t* kd_bytes;
  uint64_t val_off;            // 8-byte values | bit-packed booleans (per batch words) | int32 string offsets (n_out + 1)
  uint64_t valid_off;          // validity words, per batch
  uint64_t len_off;            // strings: u32 length per row (scratch, device only)
  uint64_t data_off;           // strings: bytes
  uint32_t stride, card;
  uint64_t wstride;            // stride in 64 bits (hashed group-by decodes the wide id)
  uint32_t kind;               // DevKind
  uint32_t is_bin;             // DATE_BIN key: value = bin_base + group id * bin_width
  int64_t bin_base, bin_width;
};
struct FinishArgs {
  const unsigned long long* acc;
  const unsigned long long* wide;   // hashed group-by: wide group id per slot (nullptr: the slot is the id)
  const uint32_t* out_slot;
  uint8_t* out;                // the result block
  uint32_t* nulls;             // [(nkeys + naggs) * nbatches] inside the block
  uint32_t n_out, nslots, n_acc, naggs, nkeys;
  uint32_t batch_rows, words_per_batch, nbatches;
  DevAgg aggs[kMaxAggs];
  uint8_t nn_is_rows[kMaxAggs];
  uint64_t val_off[kMaxAggs], valid_off[kMaxAggs];
  FinishKey keys[kMaxKeys];
};

// one thread per output row: aggregate values + validity, numeric / boolean key values, string key lengths
__global__ void k_agg_finish(const __grid_constant__ FinishArgs f) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= f.n_out) return;
  const uint32_t slot = f.out_slot[i];
  const uint32_t batch = i / f.batch_rows, pos = i - batch * f.batch_rows;
  const uint32_t word = batch * f.words_per_batch + (pos >> 5), bit = 1u << (pos & 31);
  const unsigned long long rows = f.acc[slot];
  for (uint32_t a = 0; a < f.naggs; a++) {
    const DevAgg ag = f.aggs[a];
    unsigned long long nn = 0, cell = 0;
    if (ag.fn != AG_COUNT_STAR) nn = f.nn_is_rows[a] ? rows : f.acc[size_t(1 + f.n_acc + ag.nn_slot) * f.nslots + slot];
    if (ag.fn >= AG_SUM) cell = f.acc[size_t(1 + ag.acc_slot) * f.nslots + slot];
    bool valid = true;
    unsigned long long v = 0;
    switch (ag.fn) {
      case AG_COUNT_STAR: v = rows; break;
      case AG_COUNT: v = nn; break;
      case AG_SUM: valid = nn > 0; v = cell; break;
      case AG_AVG:
        valid = nn > 0;
        if (valid) v = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)cell) / double(nn));
        break;
      default:
        valid = nn > 0;
        v = ag.kind == DK_F64 ? f64_from_order_key((int64_t)cell) : cell;
    }
    reinterpret_cast<unsigned long long*>(f.out + f.val_off[a])[i] = valid ? v : 0ull;
    if (valid) atomicOr(reinterpret_cast<uint32_t*>(f.out + f.valid_off[a]) + word, bit);
    else atomicAdd(&f.nulls[(f.nkeys + a) * f.nbatches + batch], 1u);
  }
  for (uint32_t k = 0; k < f.nkeys; k++) {
    const FinishKey& key = f.keys[k];
    const uint32_t gid = uint32_t(((f.wide ? f.wide[slot] : uint64_t(slot)) / key.wstride) % (key.card + 1));
    const bool valid = gid != key.card;   // NULL is its own group (field_stats.rs:1009-1037)
    if (valid) atomicOr(reinterpret_cast<uint32_t*>(f.out + key.valid_off) + word, bit);
    else atomicAdd(&f.nulls[k * f.nbatches + batch], 1u);
    if (key.kind == DK_STR) {
      reinterpret_cast<uint32_t*>(f.out + key.len_off)[i] = valid ? key.kd_offs[gid + 1] - key.kd_offs[gid] : 0u;
    } else if (key.kind == DK_BOOL) {
      if (valid && gid) atomicOr(reinterpret_cast<uint32_t*>(f.out + key.val_off) + word, bit);
    } else {
      unsigned long long v = 0;
      if (valid && key.is_bin) v = (unsigned long long)(key.bin_base + (long long)gid * key.bin_width);
      else if (valid) {
        const uint8_t* p = key.kd_bytes + key.kd_offs[gid];
        for (int b = 0; b < 8; b++) v |= (unsigned long long)p[b] << (8 * b);
      }
      reinterpret_cast<unsigned long long*>(f.out + key.val_off)[i] = v;
    }
  }
}

// exclusive scan of u32 lengths into int32 Arrow offsets (n + 1 entries); one block
__global__ void k_offsets_scan(const uint32_t* __restrict__ lens, uint32_t n, int32_t* __restrict__ offs) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    unsigned long long v = i < n ? lens[i] : 0, incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < nwarps ? warp_sums[lane] : 0, wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
        if ((int)lane >= o) wi += t;
      }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long excl = carry + warp_sums[warp] + incl - v;
    if (i < n) offs[i] = int32_t(excl);
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offs[n] = int32_t(carry);
}

// string key bytes: one warp per output row
__global__ void k_key_gather(const __grid_constant__ FinishArgs f, uint32_t k) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= f.n_out) return;
  const FinishKey& key = f.keys[k];
  const uint32_t gid = uint32_t(((f.wide ? f.wide[f.out_slot[i]] : uint64_t(f.out_slot[i])) / key.wstride) % (key.card + 1));
  if (gid == key.card) return;
  const uint32_t a = key.kd_offs[gid], n = key.kd_offs[gid + 1] - a;
  uint8_t* dst = f.out + key.data_off + reinterpret_cast<const int32_t*>(f.out + key.val_off)[i];
  for (uint32_t b = lane; b < n; b += 32) dst[b] = key.kd_bytes[a + b];
}


// ---- projection: TableProvider::scan(projection, ...) (stream_schema_provider.rs:526-659, :114-189) ----
// Late materialisation: the filter kernels leave a selection bitmap; only the selected rows of the
// projected columns are decoded, straight out of the flat store (row r of a page is bits
// [r*bw, (r+1)*bw) / 8-byte slot r), dictionary values through the chunk's dictionary.  The buffers
// of every result batch are assembled in one device block, like the aggregate results above.
struct ProjCol {
  const uint64_t* ent;   // strings: arena offset of every dictionary entry of the column
  uint64_t val_off;      // 8-byte values | per-batch bit words (bool) | int32 offsets (n + 1) (strings)
  uint64_t valid_off;    // per-batch validity words
  uint64_t src_off;      // strings: u64 arena offset of the bytes per output row (device-only scratch)
  uint64_t len_off;      // strings: u32 length per output row (device-only scratch)
  uint64_t data_off;     // strings: bytes
  uint32_t slot;         // column slot of the plan (chunk table, item.page, item.poff); 0xffffffff: the __row_id column
  uint32_t kind;         // DevKind
};
struct ProjArgs {
  const uint8_t* arena;
  const uint8_t* flat;
  const FlatPageRec* fpages;
  const DevChunk* chunks;
  const DevItem* items;
  const uint32_t* bitmap;
  const uint32_t* item_counts;
  const unsigned long long* item_base;
  uint8_t* out;
  uint32_t* nulls;       // [ncols * nbatches]
  unsigned long long n_out;   // output rows kept (LIMIT)
  uint32_t n_items, plan_ncols, ncols;
  uint32_t batch_rows, words_per_batch, nbatches;
  ProjCol cols[kMaxCols + 1];
};

__device__ __forceinline__ uint32_t flat_bits_at(const uint8_t* flat, uint64_t off, uint64_t bit, uint32_t bw) {
  if (bw == 0) return 0;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(flat + off) + (bit >> 5);
  const uint32_t sh = uint32_t(bit & 31);
  const uint32_t v = __funnelshift_r(w[0], w[1], sh);
  return bw >= 32 ? v : (v & ((1u << bw) - 1u));
}

// one CTA per work item (grid-strided): bitmap words -> output positions -> values of every projected column
__global__ void __launch_bounds__(256) k_project(const __grid_constant__ ProjArgs f) {
  __shared__ uint32_t warp_sums[8];
  __shared__ uint32_t carry;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (uint32_t it = blockIdx.x; it < f.n_items; it += gridDim.x) {
    if (f.item_counts[it] == 0) continue;   // uniform per block
    const DevItem& item = f.items[it];
    const uint32_t nwords = (item.nrows + 31) >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t w0 = 0; w0 < nwords; w0 += blockDim.x) {
      const uint32_t w = w0 + threadIdx.x;
      uint32_t word = w < nwords ? f.bitmap[item.bitmap_word0 + w] : 0;
      const uint32_t c = __popc(word);
      uint32_t incl = c;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += t;
      }
      if (lane == 31) warp_sums[warp] = incl;
      __syncthreads();
      if (warp == 0) {
        uint32_t s = lane < nwarps ? warp_sums[lane] : 0, si = s;
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
          if ((int)lane >= o) si += t;
        }
        if (lane < nwarps) warp_sums[lane] = si - s;
      }
      __syncthreads();
      unsigned long long pos = f.item_base[it] + carry + warp_sums[warp] + incl - c;
      while (word) {
        const uint32_t b = __ffs(word) - 1;
        word &= word - 1;
        if (pos < f.n_out) {
          const uint32_t r = w * 32 + b;   // row inside the item
          const uint32_t batch = uint32_t(pos / f.batch_rows), bp = uint32_t(pos - uint64_t(batch) * f.batch_rows);
          const uint32_t vword = batch * f.words_per_batch + (bp >> 5), vbit = 1u << (bp & 31);
          for (uint32_t ci = 0; ci < f.ncols; ci++) {
            const ProjCol& pc = f.cols[ci];
            if (pc.slot == 0xffffffffu) {   // __row_id: ordinal of the row in the scanned table
              reinterpret_cast<unsigned long long*>(f.out + pc.val_off)[pos] = item.global_row0 + r;
              continue;
            }
            if ((item.absent >> pc.slot) & 1u) {   // column missing from this file: NULL
              atomicAdd(&f.nulls[ci * f.nbatches + batch], 1u);
              continue;
            }
            const FlatPageRec fp = f.fpages[item.page[pc.slot]];
            const uint64_t row = uint64_t(item.poff[pc.slot]) + r;
            if (fp.voff != ~0ull && !((reinterpret_cast<const uint32_t*>(f.flat + fp.voff)[row >> 5] >> (row & 31)) & 1u)) {
              atomicAdd(&f.nulls[ci * f.nbatches + batch], 1u);   // NULL row: value slot stays 0, string length 0
              continue;
            }
            atomicOr(reinterpret_cast<uint32_t*>(f.out + pc.valid_off) + vword, vbit);
            if (fp.fkind == FK_BYTES) {   // PLAIN byte array: the row's bytes inside the page
              const uint64_t e = fp.base + reinterpret_cast<const uint32_t*>(f.flat + fp.off)[row];
              reinterpret_cast<unsigned long long*>(f.out + pc.src_off)[pos] = e;
              reinterpret_cast<uint32_t*>(f.out + pc.len_off)[pos] = load_u32_unaligned(f.arena + e - 4);
            } else if (fp.fkind == FK_PLAIN8) {
              reinterpret_cast<unsigned long long*>(f.out + pc.val_off)[pos] = reinterpret_cast<const unsigned long long*>(f.flat + fp.off)[row];
            } else if (fp.fkind == FK_BITS) {
              const uint32_t v = (reinterpret_cast<const uint32_t*>(f.flat + fp.off)[row >> 5] >> (row & 31)) & 1u;
              if (v) atomicOr(reinterpret_cast<uint32_t*>(f.out + pc.val_off) + vword, vbit);
            } else {
              const DevChunk& ch = f.chunks[item.rg * f.plan_ncols + pc.slot];
              uint32_t idx = flat_bits_at(f.flat, fp.off, row * fp.bw, fp.bw);
              idx = idx < ch.dict_n ? idx : (ch.dict_n ? ch.dict_n - 1 : 0);
              if (pc.kind == DK_STR) {
                const uint64_t e = pc.ent[ch.lut_base + idx];
                reinterpret_cast<unsigned long long*>(f.out + pc.src_off)[pos] = e;
                reinterpret_cast<uint32_t*>(f.out + pc.len_off)[pos] = load_u32_unaligned(f.arena + e - 4);
              } else {
                reinterpret_cast<unsigned long long*>(f.out + pc.val_off)[pos] = reinterpret_cast<const unsigned long long*>(f.flat + ch.dict8_off)[idx];
              }
            }
          }
        }
        pos++;
      }
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry += warp_sums[warp] + incl;
      __syncthreads();
    }
  }
}

// string bytes of one projected column: one warp per output row
__global__ void k_project_bytes(const __grid_constant__ ProjArgs f, uint32_t ci, unsigned long long n_rows) {
  const unsigned long long i = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (i >= n_rows) return;
  const ProjCol& pc = f.cols[ci];
  const uint32_t n = reinterpret_cast<const uint32_t*>(f.out + pc.len_off)[i];
  const uint8_t* src = f.arena + reinterpret_cast<const unsigned long long*>(f.out + pc.src_off)[i];
  uint8_t* dst = f.out + pc.data_off + reinterpret_cast<const int32_t*>(f.out + pc.val_off)[i];
  for (uint32_t b = lane; b < n; b += 32) dst[b] = src[b];
}

}  // namespace pqb
