// Page decompression on the GPU: LZ4_RAW (Parseable's default codec,
// /root/reference/src/cli.rs:441-448), SNAPPY (what the reference's CI pins,
// docker-compose-test.yaml:45), ZSTD and GZIP (zstd_decode.cuh, inflate_decode.cuh; k_decompress_zstd below).  The reference gets these from the lz4_flex 0.13 /
// snap 1.1 crates through parquet 58.1.0; here one warp
// decodes one page: every lane parses the (tiny) sequence headers redundantly — the
// loads broadcast — and the 32 lanes share the literal / match copies.
//
// Output goes into the same HBM arena the scan kernel reads, so a compressed file
// costs one extra HBM write + read of the decoded pages and nothing else changes.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_structs.hpp"
#include "inflate_decode.cuh"
#include "zstd_decode.cuh"

namespace pqb {

struct DecompJob {
  uint64_t src_off;   // compressed page payload inside the staging buffer
  uint64_t dst_off;   // decoded page payload inside the arena
  uint32_t src_len;
  uint32_t dst_len;   // uncompressed_page_size from the page header
  uint32_t codec;     // parquet CompressionCodec: 7 LZ4_RAW, 1 SNAPPY, 0 plain copy
  uint32_t _pad;
};

// literal run: `len` bytes that do not overlap, any alignment of source and destination.  A page of incompressible
// bit-packed indices or PLAIN doubles is ONE literal run of 40-160 KB handled by one warp, so the copy must keep many
// bytes in flight: the destination is walked in aligned 16-byte chunks, every lane builds its chunk from five aligned
// source words with a funnel shift (the source sits at an arbitrary byte phase), four chunks per lane and trip.
__device__ __forceinline__ void warp_literal_copy(uint8_t* __restrict__ d, const uint8_t* __restrict__ s, uint32_t len, uint32_t lane) {
  uint32_t head = uint32_t(-reinterpret_cast<uintptr_t>(d)) & 15u;   // bytes until d is 16-byte aligned
  if (head > len) head = len;
  if (lane < head) d[lane] = s[lane];
  d += head; s += head; len -= head;
  const uint32_t n16 = len >> 4;
  if (n16) {
    const uint32_t sh = (uint32_t(reinterpret_cast<uintptr_t>(s)) & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(s) & ~uintptr_t(3));
    uint4* dw = reinterpret_cast<uint4*>(d);
    uint32_t c = lane;
    for (; c + 3 * 32 < n16; c += 4 * 32) {
      uint32_t w[4][5];
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < 5; j++) w[k][j] = (j < 4 || sh) ? sw[(c + k * 32) * 4 + j] : 0u;   // the fifth word only when the phase needs it (it may lie past the source)
#pragma unroll
      for (int k = 0; k < 4; k++)
        dw[c + k * 32] = make_uint4(__funnelshift_r(w[k][0], w[k][1], sh), __funnelshift_r(w[k][1], w[k][2], sh),
                                    __funnelshift_r(w[k][2], w[k][3], sh), __funnelshift_r(w[k][3], w[k][4], sh));
    }
    for (; c < n16; c += 32) {
      uint32_t w[5];
#pragma unroll
      for (int j = 0; j < 5; j++) w[j] = (j < 4 || sh) ? sw[c * 4 + j] : 0u;
      dw[c] = make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh));
    }
  }
  const uint32_t done = n16 << 4;
  if (done + lane < len) d[done + lane] = s[done + lane];   // < 16 bytes left
}

// match copy with LZ77 overlap semantics: the source pattern [dp-off, dp) already exists, bytes
// beyond it repeat with period `off`
__device__ __forceinline__ void warp_match_copy(uint8_t* d, uint32_t dp, uint32_t off, uint32_t len, uint32_t lane) {
  if (off >= len) {
    for (uint32_t i = lane; i < len; i += 32) d[dp + i] = d[dp - off + i];
  } else {
    for (uint32_t i = lane; i < len; i += 32) d[dp + i] = d[dp - off + (i % off)];
  }
}

__global__ void k_decompress_pages(const DecompJob* __restrict__ jobs, uint32_t njobs, const uint8_t* __restrict__ src_base,
                                   uint8_t* __restrict__ arena, unsigned long long* __restrict__ counters) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= njobs) return;
  const DecompJob job = jobs[j];
  const uint8_t* s = src_base + job.src_off;
  uint8_t* d = arena + job.dst_off;
  const uint32_t sn = job.src_len, dn = job.dst_len;
  uint32_t sp = 0, dp = 0;
  bool bad = false;
  if (job.codec == 0) {
    warp_literal_copy(d, s, sn < dn ? sn : dn, lane);
    return;
  }
  if (job.codec == 7) {
    // ---- LZ4 block format: token | literal length ext | literals | offset(2) | match length ext ----
    while (sp < sn) {
      const uint32_t token = s[sp++];
      uint32_t lit = token >> 4;
      if (lit == 15) {
        uint32_t b;
        do { if (sp >= sn) { bad = true; break; } b = s[sp++]; lit += b; } while (b == 255);
      }
      if (bad || sp + lit > sn || dp + lit > dn) { bad = true; break; }
      warp_literal_copy(d + dp, s + sp, lit, lane);
      sp += lit;
      dp += lit;
      if (sp >= sn) break;  // the last sequence carries literals only
      if (sp + 2 > sn) { bad = true; break; }
      const uint32_t off = uint32_t(s[sp]) | (uint32_t(s[sp + 1]) << 8);
      sp += 2;
      uint32_t ml = token & 15;
      if (ml == 15) {
        uint32_t b;
        do { if (sp >= sn) { bad = true; break; } b = s[sp++]; ml += b; } while (b == 255);
      }
      ml += 4;
      if (bad || off == 0 || off > dp || dp + ml > dn) { bad = true; break; }
      __syncwarp();  // the literals just written may be the match source
      warp_match_copy(d, dp, off, ml, lane);
      dp += ml;
      __syncwarp();
    }
  } else {
    // ---- Snappy: varint uncompressed length, then tagged elements ----
    uint32_t ulen = 0, shift = 0;
    while (sp < sn) {
      uint32_t b = s[sp++];
      ulen |= (b & 0x7f) << shift;
      shift += 7;
      if (!(b & 0x80) || shift > 28) break;
    }
    if (ulen != dn) bad = true;
    while (!bad && sp < sn) {
      const uint32_t tag = s[sp++];
      const uint32_t kind = tag & 3;
      if (kind == 0) {
        uint32_t len = (tag >> 2) + 1;
        if (len > 60) {
          const uint32_t nb = len - 60;
          if (sp + nb > sn) { bad = true; break; }
          len = 0;
          for (uint32_t k = 0; k < nb; k++) len |= uint32_t(s[sp + k]) << (8 * k);
          len += 1;
          sp += nb;
        }
        if (sp + len > sn || dp + len > dn) { bad = true; break; }
        warp_literal_copy(d + dp, s + sp, len, lane);
        sp += len;
        dp += len;
        __syncwarp();
      } else {
        uint32_t len, off;
        if (kind == 1) {
          if (sp + 1 > sn) { bad = true; break; }
          len = 4 + ((tag >> 2) & 7);
          off = ((tag >> 5) << 8) | s[sp];
          sp += 1;
        } else if (kind == 2) {
          if (sp + 2 > sn) { bad = true; break; }
          len = (tag >> 2) + 1;
          off = uint32_t(s[sp]) | (uint32_t(s[sp + 1]) << 8);
          sp += 2;
        } else {
          if (sp + 4 > sn) { bad = true; break; }
          len = (tag >> 2) + 1;
          off = uint32_t(s[sp]) | (uint32_t(s[sp + 1]) << 8) | (uint32_t(s[sp + 2]) << 16) | (uint32_t(s[sp + 3]) << 24);
          sp += 4;
        }
        if (off == 0 || off > dp || dp + len > dn) { bad = true; break; }
        __syncwarp();
        warp_match_copy(d, dp, off, len, lane);
        dp += len;
        __syncwarp();
      }
    }
  }
  if ((bad || dp != dn) && lane == 0) atomicExch(&counters[0], 1ull);
}

// ZSTD (codec 6) and GZIP (codec 2) pages: persistent warps draw pages from a counter; every warp owns one workspace
// (tables + the literals of one zstd block) in global memory.  zstd_decode.cuh / inflate_decode.cuh hold the formats.
union HeavyWs {
  ZstdWs z;
  InflateWs g;
};
__global__ void __launch_bounds__(128) k_decompress_zstd(const DecompJob* __restrict__ jobs, uint32_t njobs, const uint8_t* __restrict__ src_base,
                                                         uint8_t* __restrict__ arena, unsigned long long* __restrict__ counters,
                                                         HeavyWs* __restrict__ ws, unsigned int* __restrict__ next) {
  const uint32_t lane = threadIdx.x & 31;
  HeavyWs& w = ws[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)];
  for (;;) {
    uint32_t j = 0;
    if (lane == 0) j = atomicAdd(next, 1u);
    j = __shfl_sync(0xffffffffu, j, 0);
    if (j >= njobs) break;
    const DecompJob job = jobs[j];
    const bool ok = job.codec == 2u ? gzip_decode(w.g, src_base + job.src_off, job.src_len, arena + job.dst_off, job.dst_len)
                                    : zstd_decode(w.z, src_base + job.src_off, job.src_len, arena + job.dst_off, job.dst_len);
    if (!ok && lane == 0) atomicExch(&counters[0], 1ull);
    __syncwarp();
  }
}

// After decompression the host still does not know two bytes it normally reads from the file:
// the definition-level byte count (first 4 bytes of a v1 page) and the dictionary index bit width.
// flags: bit 0 page has a v1 definition-level block, bits 8.. DevEnc.
__global__ void k_page_fixup(DevPage* __restrict__ pages, const uint32_t* __restrict__ which, uint32_t n,
                             const uint8_t* __restrict__ arena, unsigned long long* __restrict__ counters) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  DevPage p = pages[which[i]];
  const uint8_t* payload = arena + p.off;
  uint32_t pos = p.val_off;   // v2 pages: the level bytes in front of the values (lengths from the page header); v1: 0
  if (p.def_len == 0xffffffffu) {  // v1 page of a nullable column: 4-byte length + RLE definition levels
    uint32_t dl = 0;
    if (p.len >= 4) dl = uint32_t(payload[0]) | (uint32_t(payload[1]) << 8) | (uint32_t(payload[2]) << 16) | (uint32_t(payload[3]) << 24);
    if (p.len < 4 || uint64_t(dl) + 4 > p.len) { atomicExch(&counters[0], 2ull); dl = 0; }
    p.def_off = 4;
    p.def_len = dl;
    pos = 4 + dl;
  }
  p.val_off = pos;
  if (p.enc == DE_DICT) {
    if (pos < p.len && p.num_rows) { p.bit_width = payload[pos]; p.val_off = pos + 1; }
    else p.bit_width = 0;
  } else if (p.enc == DE_RLE_BOOL) {
    p.bit_width = 1;
    if (p.num_rows && pos + 4 <= p.len) p.val_off = pos + 4;
  }
  pages[which[i]] = p;
}

}  // namespace pqb
