"""huffb200 — B200-native Huffman codec behind DataComp's compressor API.

This package is a thin ctypes binding of ``libhuffb200.so`` (hand-written sm_100a CUDA kernels
behind the C ABI of ``include/huffb200.h``) plus Python mirrors of the reference's two service
interfaces (``service/CompressionService.java:11-66``, ``service/FrequencyService.java:6-27``).
It is test/bench plumbing: the product is the shared library, which a Java host binds directly
(INTEGRATION.md).  There is NO CPU fallback — without the CUDA library and a GPU every
compute call raises.

The directory name contains hyphens, so load it with ``__graft_entry__.load_package()``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HZ_LIB") or os.path.join(_HERE, "libhuffb200.so")   # HZ_LIB: developer knob (A/B builds)

# status codes (include/huffb200.h)
HZ_OK = 0
HZ_ERR_ARG, HZ_ERR_CUDA, HZ_ERR_NOMEM, HZ_ERR_CODE_TOO_LONG, HZ_ERR_OUT_TOO_SMALL = -1, -2, -3, -4, -5
HZ_ERR_DECODE, HZ_ERR_BAD_LENGTHS, HZ_ERR_IO, HZ_ERR_FORMAT, HZ_ERR_CHECKSUM, HZ_ERR_UNSUPPORTED = -6, -7, -8, -9, -10, -11

SEG_BYTES = 57344

_EXPORTS = [
    "hz_create", "hz_destroy", "hz_last_error", "hz_strerror", "hz_set_stream", "hz_sync", "hz_device_count",
    "hz_version", "hz_num_chunks", "hz_histogram", "hz_build_codebooks", "hz_codes_from_lengths", "hz_encode",
    "hz_encode_with_lengths", "hz_decode", "hz_sha256_chunks", "hz_compress_file", "hz_decompress_file",
    "hz_verify_file", "hz_compress_buffer", "hz_decompress_buffer", "hz_free", "hz_prof_enable", "hz_prof_reset",
    "hz_prof_count", "hz_prof_get", "hz_launch_count", "hz_host_alloc", "hz_host_free",
    "hz_comm_unique_id", "hz_comm_init", "hz_comm_destroy", "hz_encode_global", "hz_stage_metrics", "hz_stage_name",
]


class HzError(IOError):
    """Raised for every non-zero status; mirrors the reference's IOException surface."""

    def __init__(self, status, message):
        super().__init__("%s (status %d)" % (message, status))
        self.status = status


def build_library(force=False):
    """Compile libhuffb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "huffb200.h"))
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-j8", "libhuffb200.so", "datacomp"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libhuffb200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64
        L.hz_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.hz_destroy.argtypes = [vp]
        L.hz_destroy.restype = None
        L.hz_last_error.argtypes = [vp]
        L.hz_last_error.restype = C.c_char_p
        L.hz_strerror.argtypes = [C.c_int]
        L.hz_strerror.restype = C.c_char_p
        L.hz_set_stream.argtypes = [vp, vp]
        L.hz_sync.argtypes = [vp]
        L.hz_version.restype = u32
        L.hz_num_chunks.argtypes = [u64, u32]
        L.hz_num_chunks.restype = u64
        L.hz_histogram.argtypes = [vp, vp, u64, u32, vp]
        L.hz_build_codebooks.argtypes = [vp, vp, u32, vp, vp]
        L.hz_codes_from_lengths.argtypes = [vp, vp, u32, vp]
        L.hz_encode.argtypes = [vp, vp, u64, u32, vp, u64, vp, vp, vp]
        L.hz_encode_with_lengths.argtypes = [vp, vp, u64, u32, vp, vp, u64, vp]
        L.hz_decode.argtypes = [vp, vp, u64, vp, vp, vp, vp, vp, u32, vp, u64]
        L.hz_encode_global.argtypes = [vp, vp, u64, u32, vp, u64, vp, vp]
        L.hz_comm_unique_id.argtypes = [vp]
        L.hz_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.hz_comm_destroy.argtypes = [vp]
        L.hz_stage_metrics.argtypes = [vp, vp]
        L.hz_stage_name.argtypes = [C.c_int]
        L.hz_stage_name.restype = C.c_char_p
        L.hz_sha256_chunks.argtypes = [vp, vp, u64, u32, vp]
        L.hz_compress_file.argtypes = [vp, C.c_char_p, C.c_char_p, u32, C.c_char_p, i64, vp, vp]
        L.hz_decompress_file.argtypes = [vp, C.c_char_p, C.c_char_p, vp, vp]
        L.hz_verify_file.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int)]
        L.hz_compress_buffer.argtypes = [vp, vp, u64, u32, C.c_char_p, i64, C.POINTER(vp), C.POINTER(u64)]
        L.hz_decompress_buffer.argtypes = [vp, vp, u64, C.POINTER(vp), C.POINTER(u64)]
        L.hz_free.argtypes = [vp]
        L.hz_free.restype = None
        L.hz_prof_enable.argtypes = [vp, C.c_int]
        L.hz_prof_reset.argtypes = [vp]
        L.hz_prof_count.argtypes = [vp]
        L.hz_prof_get.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(u64)]
        L.hz_launch_count.argtypes = [vp]
        L.hz_launch_count.restype = u64
        L.hz_synth_fill.argtypes = [vp, vp, u64, u64, u64, vp]
        L.hz_dev_reload_knobs.argtypes = [vp]
        L.hz_host_sha256.argtypes = [vp, u64, vp]
        L.hz_host_sha256.restype = None
        _lib = L
    return _lib


PROGRESS_FN = C.CFUNCTYPE(None, C.c_double, C.c_void_p)


def _ptr(x):
    """Raw address of a numpy array, a torch tensor (host or device), bytes, an int, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    if isinstance(x, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(x)), C.c_void_p).value
    raise TypeError(type(x))


class Codec:
    """One codec context on one GPU (hz_ctx).  Array arguments may be numpy arrays (host memory)
    or torch tensors on the context's device (used in place)."""

    def __init__(self, device=0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.hz_create(device, C.byref(h))
        if rc != HZ_OK:
            raise HzError(rc, "hz_create(device=%d) failed: no usable CUDA device; there is no CPU fallback" % device)
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.hz_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != HZ_OK:
            raise HzError(rc, self._L.hz_last_error(self._h).decode(errors="replace"))

    # -- plumbing --------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self._L.hz_set_stream(self._h, cuda_stream))

    def sync(self):
        self._check(self._L.hz_sync(self._h))

    def launch_count(self):
        return int(self._L.hz_launch_count(self._h))

    def prof_enable(self, on=True):
        self._L.hz_prof_enable(self._h, int(on))

    def prof_reset(self):
        self._L.hz_prof_reset(self._h)

    def prof(self):
        out = {}
        for i in range(self._L.hz_prof_count(self._h)):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_uint64()
            self._L.hz_prof_get(self._h, i, C.byref(name), C.byref(ms), C.byref(n))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    # -- raw stage-level calls (pointers: numpy arrays / torch tensors / ints) --------------
    def histogram_raw(self, d_in, n, chunk_bytes, hist):
        self._check(self._L.hz_histogram(self._h, _ptr(d_in), n, chunk_bytes, _ptr(hist)))

    def encode_raw(self, d_in, n, chunk_bytes, out, out_cap, comp_off, len_out=None, hist_out=None):
        self._check(self._L.hz_encode(self._h, _ptr(d_in), n, chunk_bytes, _ptr(out), out_cap, _ptr(comp_off),
                                      _ptr(len_out), _ptr(hist_out)))

    def encode_with_lengths_raw(self, d_in, n, chunk_bytes, len256, out, out_cap, comp_off):
        self._check(self._L.hz_encode_with_lengths(self._h, _ptr(d_in), n, chunk_bytes, _ptr(len256), _ptr(out),
                                                   out_cap, _ptr(comp_off)))

    def decode_raw(self, comp, comp_bytes, comp_off, comp_size, orig_size, orig_off, lens, K, out, out_cap):
        self._check(self._L.hz_decode(self._h, _ptr(comp), comp_bytes, _ptr(comp_off), _ptr(comp_size),
                                      _ptr(orig_size), _ptr(orig_off), _ptr(lens), K, _ptr(out), out_cap))

    # -- global-codebook extension (one logical file over several GPUs) ----------------------
    @staticmethod
    def comm_unique_id():
        """128-byte NCCL unique id (call on ONE rank, give the bytes to every rank's comm_init)."""
        buf = C.create_string_buffer(128)
        st = lib().hz_comm_unique_id(buf)
        if st != HZ_OK:
            raise HzError(st, "hz_comm_unique_id failed (libnccl.so.2 missing?)")
        return buf.raw

    def comm_init(self, unique_id, nranks, rank):
        self._check(self._L.hz_comm_init(self._h, C.c_char_p(bytes(unique_id)), nranks, rank))

    def comm_destroy(self):
        self._check(self._L.hz_comm_destroy(self._h))

    def encode_global_raw(self, src, n, chunk_bytes, out, out_cap, comp_off, len256_out):
        self._check(self._L.hz_encode_global(self._h, _ptr(src), n, chunk_bytes, _ptr(out), out_cap, _ptr(comp_off), _ptr(len256_out)))

    def encode_global(self, data, chunk_bytes):
        """-> (payload, comp_off[K+1], len256): ONE codebook over all chunks (and all ranks after comm_init)."""
        d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        K = int(self._L.hz_num_chunks(d.size, chunk_bytes))
        cap = d.size * 4 + 16
        out = np.zeros(cap, dtype=np.uint8)
        off = np.zeros(K + 1, dtype=np.uint64)
        l256 = np.zeros(256, dtype=np.uint8)
        self.encode_global_raw(d if d.size else None, d.size, chunk_bytes, out, cap, off, l256)
        return out[: int(off[K])].copy(), off, l256

    def reload_knobs(self):
        """Developer knobs (HZ_* environment variables) are read when the context is created; re-read them."""
        self._check(self._L.hz_dev_reload_knobs(self._h))

    def synth_fill(self, d_out, n, stream_offset, seed, qtable):
        self._check(self._L.hz_synth_fill(self._h, _ptr(d_out), n, stream_offset, seed, _ptr(qtable)))

    # -- numpy conveniences (host buffers in, host buffers out) ------------------------------
    def histogram(self, data, chunk_bytes):
        d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        K = int(self._L.hz_num_chunks(d.size, chunk_bytes))
        hist = np.zeros((K, 256), dtype=np.uint32)
        if K:
            self.histogram_raw(d, d.size, chunk_bytes, hist)
        return hist

    def build_codebooks(self, hist):
        h = np.ascontiguousarray(hist, dtype=np.uint32).reshape(-1, 256)
        K = h.shape[0]
        lens = np.zeros((K, 256), dtype=np.uint8)
        codes = np.zeros((K, 256), dtype=np.uint32)
        self._check(self._L.hz_build_codebooks(self._h, _ptr(h), K, _ptr(lens), _ptr(codes)))
        return lens, codes

    def codes_from_lengths(self, lens):
        ln = np.ascontiguousarray(lens, dtype=np.uint8).reshape(-1, 256)
        codes = np.zeros(ln.shape, dtype=np.uint32)
        self._check(self._L.hz_codes_from_lengths(self._h, _ptr(ln), ln.shape[0], _ptr(codes)))
        return codes

    def encode(self, data, chunk_bytes, want_hist=False):
        """-> (payload bytes, comp_off[K+1], lens[K,256][, hist[K,256]])"""
        d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        K = int(self._L.hz_num_chunks(d.size, chunk_bytes))
        out = np.zeros(d.size + 16, dtype=np.uint8)
        off = np.zeros(K + 1, dtype=np.uint64)
        lens = np.zeros((K, 256), dtype=np.uint8)
        hist = np.zeros((K, 256), dtype=np.uint32) if want_hist else None
        self.encode_raw(d, d.size, chunk_bytes, out, d.size, off, lens, hist)
        res = (out[: int(off[K])].copy(), off, lens)
        return res + (hist,) if want_hist else res

    def encode_with_lengths(self, data, chunk_bytes, len256):
        d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        K = int(self._L.hz_num_chunks(d.size, chunk_bytes))
        cap = d.size * 4 + 16
        out = np.zeros(cap, dtype=np.uint8)
        off = np.zeros(K + 1, dtype=np.uint64)
        l256 = np.ascontiguousarray(len256, dtype=np.uint8)
        self.encode_with_lengths_raw(d, d.size, chunk_bytes, l256, out, cap, off)
        return out[: int(off[K])].copy(), off

    def decode(self, payload, comp_off, comp_size, orig_size, lens):
        p = np.ascontiguousarray(payload, dtype=np.uint8)
        co = np.ascontiguousarray(comp_off, dtype=np.uint64)
        cs = np.ascontiguousarray(comp_size, dtype=np.uint32)
        os_ = np.ascontiguousarray(orig_size, dtype=np.uint32)
        ln = np.ascontiguousarray(lens, dtype=np.uint8)
        K = cs.size
        total = int(os_.astype(np.uint64).sum())
        out = np.zeros(max(total, 1), dtype=np.uint8)
        self.decode_raw(p, p.size, co, cs, os_, None, ln, K, out, total)
        return out[:total]

    def sha256_chunks(self, data, chunk_bytes):
        d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        K = int(self._L.hz_num_chunks(d.size, chunk_bytes))
        dig = np.zeros((K, 32), dtype=np.uint8)
        if K:
            self._check(self._L.hz_sha256_chunks(self._h, _ptr(d), d.size, chunk_bytes, _ptr(dig)))
        return dig

    # -- container level ----------------------------------------------------------------------
    def compress_buffer(self, data, chunk_bytes, name="x.bin", mtime_ms=0):
        d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        out, n = C.c_void_p(), C.c_uint64()
        self._check(self._L.hz_compress_buffer(self._h, _ptr(d) if d.size else None, d.size, chunk_bytes,
                                               name.encode(), mtime_ms, C.byref(out), C.byref(n)))
        res = C.string_at(out, n.value)
        self._L.hz_free(out)
        return res

    def decompress_buffer(self, blob):
        d = np.frombuffer(blob, dtype=np.uint8)
        out, n = C.c_void_p(), C.c_uint64()
        self._check(self._L.hz_decompress_buffer(self._h, _ptr(d), d.size, C.byref(out), C.byref(n)))
        res = C.string_at(out, n.value)
        self._L.hz_free(out)
        return res

    def compress_file(self, in_path, out_path, chunk_bytes, name=None, mtime_ms=-1, progress=None):
        cb = PROGRESS_FN(lambda f, u: progress(f)) if progress else None
        self._check(self._L.hz_compress_file(self._h, os.fsencode(in_path), os.fsencode(out_path), chunk_bytes,
                                             name.encode() if name else None, mtime_ms,
                                             C.cast(cb, C.c_void_p) if cb else None, None))

    def decompress_file(self, in_path, out_path, progress=None):
        cb = PROGRESS_FN(lambda f, u: progress(f)) if progress else None
        self._check(self._L.hz_decompress_file(self._h, os.fsencode(in_path), os.fsencode(out_path),
                                               C.cast(cb, C.c_void_p) if cb else None, None))

    def stage_metrics(self):
        """{stage name: (ms, count, bytes)} of the last file-/buffer-level call, in StageMetrics.Stage order."""
        class M(C.Structure):
            _fields_ = [("ms", C.c_double), ("count", C.c_uint64), ("bytes", C.c_uint64)]
        arr = (M * 8)()
        self._check(self._L.hz_stage_metrics(self._h, arr))
        return {self._L.hz_stage_name(i).decode(): (arr[i].ms, int(arr[i].count), int(arr[i].bytes)) for i in range(8)}

    def verify_file(self, path):
        ok = C.c_int()
        self._check(self._L.hz_verify_file(self._h, os.fsencode(path), C.byref(ok)))
        return bool(ok.value)


class B200FrequencyService:
    """Mirror of service/FrequencyService.java:6-27 (computeHistogram / getServiceName / isAvailable)."""

    def __init__(self, codec=None, device=0):
        self._codec = codec or Codec(device)

    def compute_histogram(self, data, offset, length):
        d = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        window = np.ascontiguousarray(d[offset: offset + length])
        if length == 0:
            return np.zeros(256, dtype=np.int64)
        return self._codec.histogram(window, max(length, 1))[0].astype(np.int64)     # long[256]

    def get_service_name(self):
        return "B200 (CUDA sm_100a)"

    def is_available(self):
        return lib().hz_device_count() > 0


class B200CompressionService:
    """Mirror of service/CompressionService.java:11-66 with the constructor of
    cpu/CpuCompressionService.java:36 (chunk size in whole MiB) plus a bytes-granular factory
    for the 64 KB - 4 MB sweep of BASELINE.json."""

    def __init__(self, chunk_size_mb=16, device=0, _chunk_bytes=None):
        self.chunk_bytes = int(_chunk_bytes) if _chunk_bytes else int(chunk_size_mb) * 1024 * 1024
        self._codec = Codec(device)

    @classmethod
    def with_chunk_bytes(cls, chunk_bytes, device=0):
        return cls(device=device, _chunk_bytes=chunk_bytes)

    def compress(self, input_path, output_path, progress_callback=None):
        self._codec.compress_file(input_path, output_path, self.chunk_bytes, progress=progress_callback)

    def decompress(self, input_path, output_path, progress_callback=None):
        self._codec.decompress_file(input_path, output_path, progress=progress_callback)

    def resume_compression(self, input_path, output_path, last_completed_chunk, progress_callback=None):
        raise NotImplementedError("Resume not yet implemented")      # cpu/CpuCompressionService.java:636-641

    def verify_integrity(self, compressed_path):
        return self._codec.verify_file(compressed_path)

    def get_last_stage_metrics(self):
        """getLastStageMetrics() of the service classes (cpu/CpuCompressionService.java:52)."""
        return self._codec.stage_metrics()

    def get_service_name(self):
        return "B200 Compression"

    def is_available(self):
        return lib().hz_device_count() > 0

    def close(self):
        self._codec.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
