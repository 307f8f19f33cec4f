// Shortest round-trip decimal form of a double (Ryu: Ulf Adams, "Ryu: fast float-to-string conversion", PLDI 2018),
// written from the paper's algorithm for the JSON egress kernels: serde_json prints f64 with the ryu crate
// (/root/reference/src/response.rs:31-58 -> serde_json::Value), so the digits and the notation below are what the
// reference's HTTP response carries.  Host + device; tools/json_host.cpp exposes it to tests/test_json_egress.py,
// which checks digits and round trip against Python's repr (also shortest round-trip) on the CPU.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PQB_RYU_FN __host__ __device__ inline
#else
#define PQB_RYU_FN inline
#endif

namespace pqb {

// the tables exist twice under nvcc: one copy in device memory, one for host callers
#if defined(__CUDACC__)
namespace ryu_dev {
#define PQB_RYU_TABLE __device__ const
#include "ryu_tables.inc"
}  // namespace ryu_dev
#undef PQB_RYU_TABLE
#endif
namespace ryu_host {
#define PQB_RYU_TABLE static const
#include "ryu_tables.inc"
}  // namespace ryu_host
#undef PQB_RYU_TABLE

PQB_RYU_FN const uint64_t* ryu_pow5_inv(uint32_t i) {
#if defined(__CUDA_ARCH__)
  return ryu_dev::kRyuPow5Inv[i];
#else
  return ryu_host::kRyuPow5Inv[i];
#endif
}
PQB_RYU_FN const uint64_t* ryu_pow5(uint32_t i) {
#if defined(__CUDA_ARCH__)
  return ryu_dev::kRyuPow5[i];
#else
  return ryu_host::kRyuPow5[i];
#endif
}

PQB_RYU_FN uint64_t ryu_umulh(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return uint64_t((unsigned __int128)a * b >> 64);
#endif
}
// (m * mul) >> j for a 128-bit mul = {lo, hi}, 64 <= j < 128 + 64
PQB_RYU_FN uint64_t ryu_mulshift(uint64_t m, const uint64_t* mul, uint32_t j) {
  // m * lo = (h0 : l0), m * hi = (h1 : l1); sum = (h1 : l1 + h0 carry) above bit 64
  const uint64_t h0 = ryu_umulh(m, mul[0]);
  const uint64_t l1 = m * mul[1], h1 = ryu_umulh(m, mul[1]);
  const uint64_t mid = h0 + l1;
  const uint64_t hi = h1 + (mid < h0 ? 1u : 0u);
  const uint32_t s = j - 64;   // shift of the 128-bit value (hi : mid)
  if (s == 0) return mid;
  if (s < 64) return (mid >> s) | (hi << (64 - s));
  return hi >> (s - 64);
}
PQB_RYU_FN uint32_t ryu_pow5bits(int32_t e) { return uint32_t(((uint32_t(e) * 1217359u) >> 19) + 1); }
PQB_RYU_FN uint32_t ryu_log10pow2(int32_t e) { return (uint32_t(e) * 78913u) >> 18; }
PQB_RYU_FN uint32_t ryu_log10pow5(int32_t e) { return (uint32_t(e) * 732923u) >> 20; }
PQB_RYU_FN bool ryu_mult_pow5(uint64_t v, uint32_t p) {
  uint32_t c = 0;
  while (v && v % 5 == 0) { v /= 5; c++; if (c >= p) return true; }
  return c >= p;
}
PQB_RYU_FN bool ryu_mult_pow2(uint64_t v, uint32_t p) { return (v & ((1ull << p) - 1ull)) == 0; }

// finite, non-zero |value| -> shortest digits (as an integer) and the decimal exponent of its last digit
PQB_RYU_FN void ryu_d2d(uint64_t mant, uint32_t expo, uint64_t& digits, int32_t& exp10) {
  int32_t e2;
  uint64_t m2;
  if (expo == 0) { e2 = 1 - 1023 - 52 - 2; m2 = mant; }
  else { e2 = int32_t(expo) - 1023 - 52 - 2; m2 = (1ull << 52) | mant; }
  const bool accept = (m2 & 1u) == 0;
  const uint64_t mv = 4 * m2;
  const uint32_t mm_shift = (mant != 0 || expo <= 1) ? 1u : 0u;
  uint64_t vr, vp, vm;
  int32_t e10;
  bool vm_tz = false, vr_tz = false;
  if (e2 >= 0) {
    const uint32_t q = ryu_log10pow2(e2) - (e2 > 3 ? 1u : 0u);
    e10 = int32_t(q);
    const int32_t k = 125 + int32_t(ryu_pow5bits(int32_t(q))) - 1;
    const int32_t i = -e2 + int32_t(q) + k;
    const uint64_t* mul = ryu_pow5_inv(q);
    vr = ryu_mulshift(4 * m2, mul, uint32_t(i));
    vp = ryu_mulshift(4 * m2 + 2, mul, uint32_t(i));
    vm = ryu_mulshift(4 * m2 - 1 - mm_shift, mul, uint32_t(i));
    if (q <= 21) {
      if (mv % 5 == 0) vr_tz = ryu_mult_pow5(mv, q);
      else if (accept) vm_tz = ryu_mult_pow5(mv - 1 - mm_shift, q);
      else vp -= ryu_mult_pow5(mv + 2, q) ? 1u : 0u;
    }
  } else {
    const uint32_t q = ryu_log10pow5(-e2) - (-e2 > 1 ? 1u : 0u);
    e10 = int32_t(q) + e2;
    const int32_t i = -e2 - int32_t(q);
    const int32_t k = int32_t(ryu_pow5bits(i)) - 125;
    const int32_t j = int32_t(q) - k;
    const uint64_t* mul = ryu_pow5(uint32_t(i));
    vr = ryu_mulshift(4 * m2, mul, uint32_t(j));
    vp = ryu_mulshift(4 * m2 + 2, mul, uint32_t(j));
    vm = ryu_mulshift(4 * m2 - 1 - mm_shift, mul, uint32_t(j));
    if (q <= 1) {
      vr_tz = true;
      if (accept) vm_tz = mm_shift == 1;
      else --vp;
    } else if (q < 63) {
      vr_tz = ryu_mult_pow2(mv, q);
    }
  }
  int32_t removed = 0;
  uint32_t last = 0;
  uint64_t out;
  if (vm_tz || vr_tz) {
    while (vp / 10 > vm / 10) {
      vm_tz = vm_tz && vm % 10 == 0;
      vr_tz = vr_tz && last == 0;
      last = uint32_t(vr % 10);
      vr /= 10; vp /= 10; vm /= 10;
      ++removed;
    }
    if (vm_tz) {
      while (vm % 10 == 0) {
        vr_tz = vr_tz && last == 0;
        last = uint32_t(vr % 10);
        vr /= 10; vp /= 10; vm /= 10;
        ++removed;
      }
    }
    if (vr_tz && last == 5 && vr % 2 == 0) last = 4;   // exactly half: round to even
    out = vr + (((vr == vm && (!accept || !vm_tz)) || last >= 5) ? 1u : 0u);
  } else {
    bool up = false;
    while (vp / 10 > vm / 10) {
      up = vr % 10 >= 5;
      vr /= 10; vp /= 10; vm /= 10;
      ++removed;
    }
    out = vr + ((vr == vm || up) ? 1u : 0u);
  }
  digits = out;
  exp10 = e10 + removed;
}

PQB_RYU_FN uint32_t ryu_declen(uint64_t v) {
  uint32_t n = 1;
  while (v >= 10) { v /= 10; n++; }
  return n;
}

// JSON number text of a FINITE double, the way serde_json (ryu crate, "pretty" notation) prints it; returns the length
// (at most 24 bytes).  Non-finite values are the caller's business (JSON has none: arrow-json writes null).
PQB_RYU_FN uint32_t ryu_format_f64(double value, char* out) {
  uint64_t bits;
  memcpy(&bits, &value, 8);
  const bool neg = (bits >> 63) != 0;
  const uint64_t mant = bits & ((1ull << 52) - 1);
  const uint32_t expo = uint32_t((bits >> 52) & 0x7ffu);
  uint32_t n = 0;
  if (neg) out[n++] = '-';
  if (expo == 0 && mant == 0) { out[n++] = '0'; out[n++] = '.'; out[n++] = '0'; return n; }
  uint64_t digits;
  int32_t k;
  ryu_d2d(mant, expo, digits, k);
  const int32_t len = int32_t(ryu_declen(digits));
  const int32_t kk = len + k;
  char d[20];
  { uint64_t v = digits; for (int32_t i = len - 1; i >= 0; i--) { d[i] = char('0' + v % 10); v /= 10; } }
  if (0 <= k && kk <= 16) {                       // 1234e7 -> 12340000000.0
    for (int32_t i = 0; i < len; i++) out[n++] = d[i];
    for (int32_t i = 0; i < k; i++) out[n++] = '0';
    out[n++] = '.'; out[n++] = '0';
  } else if (0 < kk && kk <= 16) {                // 1234e-2 -> 12.34
    for (int32_t i = 0; i < kk; i++) out[n++] = d[i];
    out[n++] = '.';
    for (int32_t i = kk; i < len; i++) out[n++] = d[i];
  } else if (-5 < kk && kk <= 0) {                // 1234e-6 -> 0.001234
    out[n++] = '0'; out[n++] = '.';
    for (int32_t i = 0; i < -kk; i++) out[n++] = '0';
    for (int32_t i = 0; i < len; i++) out[n++] = d[i];
  } else {                                        // 1e30, 1.234e33
    out[n++] = d[0];
    if (len > 1) { out[n++] = '.'; for (int32_t i = 1; i < len; i++) out[n++] = d[i]; }
    out[n++] = 'e';
    int32_t e = kk - 1;
    if (e < 0) { out[n++] = '-'; e = -e; }
    if (e >= 100) { out[n++] = char('0' + e / 100); e %= 100; out[n++] = char('0' + e / 10); out[n++] = char('0' + e % 10); }
    else if (e >= 10) { out[n++] = char('0' + e / 10); out[n++] = char('0' + e % 10); }
    else out[n++] = char('0' + e);
  }
  return n;
}

}  // namespace pqb
