"""The formatting the JSON egress kernels run per value (csrc/json_egress.cuh, csrc/ryu_f64.cuh), compiled for the
host: shortest round-trip doubles against Python's repr (the same digits serde_json's ryu prints; the reference's
response goes through serde_json: src/response.rs:31-58), Int64, Timestamp(ms) the way chrono prints a NaiveDateTime,
string escapes the way serde_json writes them."""
import ctypes as C
import datetime as dt
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def jh():
    so = os.path.join(ROOT, "tools", "libjson_host.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", ROOT, "tools"])
    lib = C.CDLL(so)
    lib.jh_format_f64_many.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    lib.jh_format_i64.argtypes = [C.c_int64, C.c_char_p]
    lib.jh_format_ts_ms.argtypes = [C.c_int64, C.c_char_p]
    lib.jh_escape.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
    return lib


def _digits_exp(s: str):
    s = s.lstrip("-")
    m, e = (s.split("e") + ["0"])[:2] if "e" in s else (s, "0")
    a, b = (m.split(".") + [""])[:2]
    d, e = (a + b).lstrip("0") or "0", int(e) - len(b)
    while len(d) > 1 and d.endswith("0"):
        d, e = d[:-1], e + 1
    return d, e


def _format_many(lib, vals):
    v = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.zeros(len(v) * 32, np.uint8)
    lib.jh_format_f64_many(v.ctypes.data, len(v), out.ctypes.data)
    return [bytes(out[i * 32:(i + 1) * 32]).split(b"\0")[0].decode() for i in range(len(v))]


def test_f64_shortest_round_trip_digits(jh):
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2**64 - 1, 120_000, dtype=np.uint64).view(np.float64)
    sets = [bits[np.isfinite(bits)], rng.standard_normal(50_000), rng.random(50_000) * 10.0 ** rng.integers(-30, 30, 50_000),
            np.floor(rng.lognormal(3.5, 1.2, 50_000)) / 1000.0, rng.integers(-10**15, 10**15, 50_000).astype(np.float64),
            np.array([0.1, 0.2, 0.3, 1e15, 1e16, 1e17, 1e21, 1e22, 1e23, 1e-5, 1e-6, 1e-7, 5e-324, 2.2250738585072014e-308, 2.225073858507201e-308,
                      1.7976931348623157e308, 9007199254740993.0, 4.35, 0.5, 1 / 3, 2 / 3, 1e100, 9.999999999999999e22, 123456789012345678.0])]
    for vals in sets:
        for x, s in zip(vals, _format_many(jh, vals)):
            assert float(s) == float(x) and _digits_exp(s) == _digits_exp(repr(float(x))), (repr(float(x)), s)
            assert json.loads(s) == float(x)


def test_f64_notation_is_the_ryu_crates(jh):
    """serde_json -> ryu::Buffer::format_finite: decimal notation for 1e-5 <= |x| < 1e16, exponent form outside, always a
    fraction or an exponent (an f64 never prints like an integer)."""
    want = {-0.0: "-0.0", 1.0: "1.0", -1.5: "-1.5", 1e15: "1000000000000000.0", 1e16: "1e16", 1.2345678901234568e17: "1.2345678901234568e17",
            1e-5: "0.00001", 1.234e-5: "0.00001234", 1e-6: "1e-6", 1.5e-7: "1.5e-7", 123.456: "123.456", 5e-324: "5e-324", 1e100: "1e100", 0.3: "0.3",
            1.7976931348623157e308: "1.7976931348623157e308"}
    got = _format_many(jh, list(want) + [0.0])
    assert got == list(want.values()) + ["0.0"]


def test_i64_and_timestamps(jh):
    buf = C.create_string_buffer(64)
    for v in (0, 1, -1, 42, -9223372036854775808, 9223372036854775807, 1000000, -987654321012):
        n = jh.jh_format_i64(v, buf)
        assert buf.raw[:n].decode() == str(v)
    epoch = dt.datetime(1970, 1, 1)
    rng = np.random.default_rng(1)
    cases = [0, 1, 999, 1000, -1, -1000, 86_399_999, 86_400_000, 1_700_000_000_000, 1_700_000_000_500, 951_782_400_000, 951_868_800_000,
             4_102_444_800_000, -2_208_988_800_000, 253_402_300_799_999] + [int(x) for x in rng.integers(-6 * 10**13, 2 * 10**14, 3000)]
    for ms in cases:
        n = jh.jh_format_ts_ms(ms, buf)
        t = epoch + dt.timedelta(milliseconds=ms)
        want = t.strftime("%Y-%m-%dT%H:%M:%S") if t.year >= 1000 else f"{t.year:04d}" + t.strftime("-%m-%dT%H:%M:%S")
        if ms % 1000:
            want += f".{ms % 1000:03d}"
        assert buf.raw[:n].decode() == want, ms


def test_string_escapes_match_json(jh):
    rng = np.random.default_rng(2)
    samples = ["", "plain", 'quote " and \\ backslash', "tab\tnl\ncr\r", "\b\f\x00\x01\x1f", "δέλτα ünï ✓ 😀", "a/b", "\x7f"]
    samples += ["".join(chr(int(c)) for c in rng.integers(0, 128, int(rng.integers(0, 40)))) for _ in range(500)]
    out = C.create_string_buffer(4096)
    for s in samples:
        b = s.encode()
        n = jh.jh_escape(b, len(b), out)
        assert n != 0xffffffff
        txt = '"' + out.raw[:n].decode() + '"'
        assert json.loads(txt) == s
        assert txt == json.dumps(s, ensure_ascii=False)      # serde_json and Python agree on the short escapes and \\u00XX
