// TEST HARNESS ONLY: exposes the pure host/device functions of decode_core.cuh to
// tests/test_decode_core.py so the run-walker / bit-unpacker are exercised on the
// CPU.  Never linked into libparseable_b200.so.
#include <cstdint>
#include <cstring>
#include <vector>
#include "decode_core.cuh"
using namespace pqb;

extern "C" {
// Decode `n` values of an RLE/bit-packed hybrid stream exactly as the scan kernel does: windows of
// `win_cap` bytes, slabs of `slab` values, directory of `max_ent` entries.  Returns values decoded.
int64_t dc_decode_hybrid(const uint8_t* stream, uint64_t len, uint32_t bw, uint32_t n, uint32_t slab,
                         uint32_t win_cap, uint32_t max_ent, uint32_t* out) {
  std::vector<uint8_t> padded(len + win_cap + 64, 0);
  std::memcpy(padded.data() + 16, stream, len);   // arena offset 16: unaligned start on purpose? no, 16-aligned
  StreamState st;
  stream_init(st, 16, 16 + len, bw);
  uint32_t done = 0;
  std::vector<DirEntry> dir(max_ent);
  std::vector<uint32_t> win((win_cap + 16) / 4 + 4);
  int guard = 0;
  while (done < n) {
    uint64_t s = stream_window_start(st) & ~15ull;
    std::memset(win.data(), 0, win.size() * 4);
    uint64_t avail = padded.size() - s;
    std::memcpy(win.data(), padded.data() + s, avail < win_cap ? avail : win_cap);
    Window w{reinterpret_cast<const uint8_t*>(win.data()), s, win_cap};
    uint32_t need = n - done < slab ? n - done : slab;
    uint32_t nent = 0;
    uint32_t got = walk_stream(st, w, need, dir.data(), nent, max_ent);
    if (got == 0) { if (++guard > 2) return -int64_t(done) - 1; continue; }
    guard = 0;
    for (uint32_t e = 0; e < nent; e++) {
      const DirEntry& d = dir[e];
      for (uint32_t j = 0; j < d.count; j++)
        out[done + d.start + j] = d.kind ? bp_get(win.data(), d.payload, bw, j) : d.payload;
    }
    done += got;
  }
  return done;
}
// hybrid stream -> flat bit-packed words (the slab index's copy of run-heavy pages), `slab` values per call
int64_t dc_transcode(const uint8_t* stream, uint64_t len, uint32_t bw, uint32_t n, uint32_t slab, uint32_t* out_words) {
  std::vector<uint8_t> padded(len + 64 + 16, 0);
  std::memcpy(padded.data() + 16, stream, len);
  StreamState st;
  stream_init(st, 16, 16 + len, bw);
  BitWriter b{out_words, 0, 0};
  uint32_t done = 0;
  while (done < n) {
    uint32_t need = n - done < slab ? n - done : slab;
    uint32_t got = transcode_values(st, padded.data(), need, b);
    done += got;
    if (got < need) return -int64_t(done) - 1;
  }
  bitwriter_flush(b);
  return done;
}
// DELTA_BINARY_PACKED page payload -> int64 values, through the kernel's window / directory geometry
int64_t dc_decode_delta(const uint8_t* stream, uint64_t len, uint32_t n, uint32_t slab, uint32_t win_cap,
                        uint32_t max_ent, int64_t* out) {
  std::vector<uint8_t> padded(len + win_cap + 64, 0);
  std::memcpy(padded.data() + 16, stream, len);
  DeltaState st;
  delta_init(st, 16, 16 + len);
  uint32_t done = 0;
  std::vector<DeltaEntry> dir(max_ent);
  std::vector<uint32_t> win((win_cap + 16) / 4 + 4);
  int guard = 0;
  while (done < n) {
    uint64_t s = delta_window_start(st) & ~15ull;
    std::memset(win.data(), 0, win.size() * 4);
    uint64_t avail = padded.size() - s;
    std::memcpy(win.data(), padded.data() + s, avail < win_cap ? avail : win_cap);
    Window w{reinterpret_cast<const uint8_t*>(win.data()), s, win_cap};
    uint32_t need = n - done < slab ? n - done : slab;
    uint32_t nent = 0;
    uint32_t got = walk_delta(st, w, need, dir.data(), nent, max_ent);
    if (got == 0) { if (++guard > 2 || st.bad) return -int64_t(done) - 1; continue; }
    guard = 0;
    for (uint32_t e = 0; e < nent; e++) {
      const DeltaEntry& d = dir[e];
      for (uint32_t j = 0; j < d.count; j++) {
        int64_t delta = d.kind ? d.min_delta : int64_t(uint64_t(d.min_delta) + bp_get64(win.data(), d.bitoff, d.bw, j));
        st.last_value = int64_t(uint64_t(st.last_value) + uint64_t(delta));
        out[done + d.start + j] = st.last_value;
      }
    }
    done += got;
  }
  return done;
}
int64_t dc_f64_key(uint64_t bits) { return f64_order_key(bits); }
uint64_t dc_f64_from_key(int64_t k) { return f64_from_order_key(k); }
int dc_like(const uint8_t* s, uint32_t n, const uint8_t* p, uint32_t m, uint32_t kind, int ci) {
  return like_match(s, n, p, m, kind, ci != 0);
}
uint64_t dc_load_u64(const uint8_t* base, uint32_t off) { return load_u64_unaligned(base + off); }
}
