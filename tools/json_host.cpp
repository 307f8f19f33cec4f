// TEST HARNESS ONLY: host build of the number / string formatting the JSON egress kernels use
// (parseable_b200/csrc/ryu_f64.cuh, json_format.cuh), for tests/test_json_egress.py.  Never linked into the product.
#include <cstdint>
#include "ryu_f64.cuh"

extern "C" uint32_t jh_format_f64(double v, char* out) { return pqb::ryu_format_f64(v, out); }
extern "C" void jh_format_f64_many(const double* v, uint64_t n, char* out /* 32 bytes each, NUL padded */) {
  for (uint64_t i = 0; i < n; i++) {
    char* o = out + i * 32;
    const uint32_t k = pqb::ryu_format_f64(v[i], o);
    for (uint32_t j = k; j < 32; j++) o[j] = 0;
  }
}
