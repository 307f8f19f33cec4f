// Minimal Thrift compact-protocol reader: just enough to walk Parquet's
// FileMetaData and PageHeader structures.  Replaces what the `parquet` crate's
// thrift module does for the reference (Cargo.lock:3843; not in /root/reference).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace pqb {

struct ThriftError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

enum ThriftType : uint8_t {
  T_STOP = 0, T_TRUE = 1, T_FALSE = 2, T_BYTE = 3, T_I16 = 4, T_I32 = 5, T_I64 = 6,
  T_DOUBLE = 7, T_BINARY = 8, T_LIST = 9, T_SET = 10, T_MAP = 11, T_STRUCT = 12
};

class ThriftReader {
 public:
  ThriftReader(const uint8_t* p, size_t n) : p_(p), end_(p + n), begin_(p) {}
  size_t consumed() const { return size_t(p_ - begin_); }

  uint8_t byte() {
    if (p_ >= end_) throw ThriftError("thrift: truncated");
    return *p_++;
  }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      uint8_t b = byte();
      v |= uint64_t(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    throw ThriftError("thrift: varint too long");
  }
  int64_t zigzag() {
    uint64_t v = varint();
    return int64_t(v >> 1) ^ -int64_t(v & 1);
  }
  // Field header; returns false on STOP.  last_id carries the delta state.
  bool field(int16_t& last_id, int16_t& id, uint8_t& type) {
    uint8_t h = byte();
    if (h == 0) return false;
    type = h & 0x0f;
    uint8_t delta = h >> 4;
    id = delta ? int16_t(last_id + delta) : int16_t(zigzag());
    last_id = id;
    return true;
  }
  void list_header(uint32_t& size, uint8_t& elem) {
    uint8_t h = byte();
    elem = h & 0x0f;
    size = h >> 4;
    if (size == 15) {
      const uint64_t n = varint();
      // every element takes at least one byte: a count beyond the remaining bytes is a crafted / corrupt footer
      // (callers reserve() on it)
      if (n > uint64_t(end_ - p_)) throw ThriftError("thrift: list longer than the buffer");
      size = uint32_t(n);
    }
  }
  std::string binary() {
    uint64_t n = varint();
    if (uint64_t(end_ - p_) < n) throw ThriftError("thrift: truncated binary");
    std::string s(reinterpret_cast<const char*>(p_), size_t(n));
    p_ += n;
    return s;
  }
  double f64() {
    if (end_ - p_ < 8) throw ThriftError("thrift: truncated double");
    double d;
    std::memcpy(&d, p_, 8);
    p_ += 8;
    return d;
  }
  void skip(uint8_t type) {
    struct Depth { int& d; Depth(int& x) : d(x) { if (++d > kMaxDepth) throw ThriftError("thrift: nesting too deep"); } ~Depth() { --d; } } guard(depth_);
    switch (type) {
      case T_TRUE: case T_FALSE: break;
      case T_BYTE: byte(); break;
      case T_I16: case T_I32: case T_I64: varint(); break;
      case T_DOUBLE: f64(); break;
      case T_BINARY: {
        uint64_t n = varint();
        if (uint64_t(end_ - p_) < n) throw ThriftError("thrift: truncated");
        p_ += n;
        break;
      }
      case T_LIST: case T_SET: {
        uint32_t n; uint8_t et;
        list_header(n, et);
        for (uint32_t i = 0; i < n; i++) skip_elem(et);
        break;
      }
      case T_MAP: {
        uint64_t n = varint();
        if (n > uint64_t(end_ - p_)) throw ThriftError("thrift: map longer than the buffer");
        if (n) {
          uint8_t kv = byte();
          for (uint64_t i = 0; i < n; i++) { skip_elem(kv >> 4); skip_elem(kv & 0x0f); }
        }
        break;
      }
      case T_STRUCT: skip_struct(); break;
      default: throw ThriftError("thrift: bad type");
    }
  }
  void skip_struct() {
    int16_t last = 0, id; uint8_t t;
    while (field(last, id, t)) skip(t);
  }

 private:
  // list/map elements of bool type occupy one byte each
  void skip_elem(uint8_t t) {
    if (t == T_TRUE || t == T_FALSE) byte(); else skip(t);
  }
  static constexpr int kMaxDepth = 64;   // skip() / skip_struct() recurse on nested containers
  int depth_ = 0;
  const uint8_t* p_;
  const uint8_t* end_;
  const uint8_t* begin_;
};

}  // namespace pqb
