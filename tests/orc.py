"""ctypes binding of the parity oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

The oracle restates the reference CPU path (see oracle/huff_oracle.cpp); only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liborc.so")


def build(force=False):
    src = os.path.join(_ROOT, "oracle", "huff_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u32p, i32p, u64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_uint64))
        L.orc_java_random_bytes.argtypes = [C.c_int64, u8p, C.c_size_t]
        L.orc_histogram.argtypes = [u8p, C.c_size_t, u64p]
        L.orc_build_code_lengths.argtypes = [u64p, i32p]
        L.orc_canonical_codes.argtypes = [i32p, u32p]
        for f in (L.orc_encode_literal, L.orc_encode_fast):
            f.argtypes = [u8p, C.c_size_t, i32p, u32p, u8p, C.c_size_t]
            f.restype = C.c_int64
        for f in (L.orc_decode_literal, L.orc_decode_fast):
            f.argtypes = [u8p, C.c_size_t, i32p, u8p, C.c_size_t]
            f.restype = C.c_int64
        L.orc_sha256.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_compress_buffer.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_char_p, C.c_int64,
                                          C.c_int, C.c_int, C.POINTER(u8p), u64p]
        L.orc_decompress_buffer.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.POINTER(u8p), u64p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_encode_chunks_mt.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, u8p, u32p, i32p]
        L.orc_decode_chunks_mt.argtypes = [u8p, u32p, i32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, u8p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(data):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


def java_random_bytes(seed, n):
    out = np.empty(n, dtype=np.uint8)
    lib().orc_java_random_bytes(seed, _p(out, C.c_uint8), n)
    return out


def histogram(data):
    d = _u8(data)
    h = np.zeros(256, dtype=np.uint64)
    lib().orc_histogram(_p(d, C.c_uint8), d.size, _p(h, C.c_uint64))
    return h


def code_lengths(freq):
    f = np.ascontiguousarray(freq, dtype=np.uint64)
    ln = np.zeros(256, dtype=np.int32)
    mx = lib().orc_build_code_lengths(_p(f, C.c_uint64), _p(ln, C.c_int32))
    return ln, mx


def canonical_codes(lengths):
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    code = np.zeros(256, dtype=np.uint32)
    mx = lib().orc_canonical_codes(_p(ln, C.c_int32), _p(code, C.c_uint32))
    return code, mx


def encode(data, lengths, codes, literal=False):
    d = _u8(data)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    cd = np.ascontiguousarray(codes, dtype=np.uint32)
    out = np.empty(d.size * 4 + 8, dtype=np.uint8)
    f = lib().orc_encode_literal if literal else lib().orc_encode_fast
    n = f(_p(d, C.c_uint8), d.size, _p(ln, C.c_int32), _p(cd, C.c_uint32), _p(out, C.c_uint8), out.size)
    assert n >= 0
    return out[:n].copy()


def decode(comp, lengths, out_size, literal=False):
    c = _u8(comp)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    out = np.empty(max(out_size, 1), dtype=np.uint8)
    f = lib().orc_decode_literal if literal else lib().orc_decode_fast
    rc = f(_p(c, C.c_uint8), c.size, _p(ln, C.c_int32), _p(out, C.c_uint8), out_size)
    return out[:out_size], rc


def sha256(data):
    d = _u8(data)
    out = np.empty(32, dtype=np.uint8)
    lib().orc_sha256(_p(d, C.c_uint8), d.size, _p(out, C.c_uint8))
    return out.tobytes()


def encode_chunk(data, literal=False):
    """histogram -> codebook -> encode of ONE chunk; returns (payload, lengths[256], codes[256])."""
    ln, mx = code_lengths(histogram(data))
    assert mx >= 0
    cd, _ = canonical_codes(ln)
    return encode(data, ln, cd, literal), ln, cd


def compress(data, chunk_bytes, name="x.bin", mtime_ms=0, literal=False, threads=0):
    d = _u8(data)
    out = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    rc = lib().orc_compress_buffer(_p(d, C.c_uint8), d.size, chunk_bytes, name.encode(), mtime_ms,
                                   int(literal), threads, C.byref(out), C.byref(n))
    if rc != 0:
        raise RuntimeError("oracle compress failed rc=%d" % rc)
    res = bytes(C.cast(out, C.POINTER(C.c_uint8 * n.value)).contents) if n.value else b""
    lib().orc_free(out)
    return res


def decompress(blob, literal=False, threads=0):
    d = _u8(blob)
    out = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    rc = lib().orc_decompress_buffer(_p(d, C.c_uint8), d.size, int(literal), threads, C.byref(out), C.byref(n))
    if rc != 0:
        raise RuntimeError("oracle decompress failed rc=%d" % rc)
    res = bytes(C.cast(out, C.POINTER(C.c_uint8 * n.value)).contents) if n.value else b""
    lib().orc_free(out)
    return res


def encode_chunks_mt(data, chunk_bytes, literal=True, threads=0):
    d = _u8(data)
    K = (d.size + chunk_bytes - 1) // chunk_bytes
    comp = np.empty(K * (chunk_bytes + 8), dtype=np.uint8)
    sizes = np.zeros(K, dtype=np.uint32)
    lens = np.zeros(K * 256, dtype=np.int32)
    T = lib().orc_encode_chunks_mt(_p(d, C.c_uint8), d.size, chunk_bytes, int(literal), threads,
                                   _p(comp, C.c_uint8), _p(sizes, C.c_uint32), _p(lens, C.c_int32))
    assert T > 0
    return comp, sizes, lens.reshape(K, 256), T


def decode_chunks_mt(comp, sizes, lens, n, chunk_bytes, literal=True, threads=0):
    out = np.empty(n, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    T = lib().orc_decode_chunks_mt(_p(comp, C.c_uint8), _p(sizes, C.c_uint32), _p(lens, C.c_int32), n,
                                   chunk_bytes, int(literal), threads, _p(out, C.c_uint8))
    assert T > 0
    return out, T
