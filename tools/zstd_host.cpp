// TEST HARNESS ONLY: the ZSTD and GZIP page decoders of parseable_b200/csrc/zstd_decode.cuh / inflate_decode.cuh
// compiled for the host (one "lane"), so that tests/test_zstd.py can check them against pyarrow's codecs on the CPU.
// Never linked into libparseable_b200.so.
#include <cstdint>
#include <cstdlib>
#include <new>
#include "inflate_decode.cuh"
#include "zstd_decode.cuh"

extern "C" int zs_host_decode(const uint8_t* src, uint32_t sn, uint8_t* dst, uint64_t dn) {
  pqb::ZstdWs* w = new (std::nothrow) pqb::ZstdWs();
  if (!w) return -1;
  const bool ok = pqb::zstd_decode(*w, src, sn, dst, dn);
  delete w;
  return ok ? 1 : 0;
}
extern "C" int gz_host_decode(const uint8_t* src, uint32_t sn, uint8_t* dst, uint64_t dn) {
  pqb::InflateWs* w = new (std::nothrow) pqb::InflateWs();
  if (!w) return -1;
  const bool ok = pqb::gzip_decode(*w, src, sn, dst, dn);
  delete w;
  return ok ? 1 : 0;
}
