"""The drop-in boundary without a GPU: libhuffb200.so loads, exports every entry point include/*.h declares, its
device-free helpers answer, and every compute path FAILS LOUDLY when there is no CUDA device (no CPU fallback) —
the library, the Python mirror of the reference's service interfaces, and the CLI (cli/DataCompCLI.java:24-91,155-169:
usage text and exit code 1).  No compute call is made here."""
import ctypes as C
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "data-compression-implementing-gpu-driven-huffman-encoding-in-java_b200")
DECL = re.compile(r"^\s*(?:const\s+char\s*\*|int|void|uint32_t|uint64_t)\s+(hz_\w+)\s*\(", re.M)


def _declared():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        names += DECL.findall(open(h).read())
    return sorted(set(names))


@pytest.fixture(scope="module")
def so(hz):
    return C.CDLL(hz.build_library())


def _no_gpu(so):
    return so.hz_device_count() <= 0


def test_headers_declare_the_boundary():
    names = _declared()
    assert len(names) >= 30
    for must in ("hz_create", "hz_destroy", "hz_histogram", "hz_build_codebooks", "hz_encode", "hz_decode",
                 "hz_compress_file", "hz_decompress_file", "hz_verify_file", "hz_last_error", "hz_sync"):
        assert must in names


def test_library_exports_every_declared_symbol(so):
    missing = [n for n in _declared() if not hasattr(so, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing


def test_python_binding_uses_only_declared_symbols(hz):
    assert set(hz._EXPORTS) <= set(_declared())


def test_device_free_helpers(so, hz):
    so.hz_version.restype = C.c_uint32
    assert so.hz_version() >= 0x100
    so.hz_num_chunks.restype = C.c_uint64
    so.hz_num_chunks.argtypes = [C.c_uint64, C.c_uint32]
    for n, chunk, want in ((0, 1024, 0), (1, 1024, 1), (1024, 1024, 1), (1025, 1024, 2), (1 << 34, 1 << 24, 1024)):
        assert so.hz_num_chunks(n, chunk) == want               # (n + chunk - 1) / chunk, cpu/CpuCompressionService.java:64
    so.hz_strerror.restype = C.c_char_p
    texts = {so.hz_strerror(code) for code in range(0, -12, -1)}
    assert len(texts) == 12 and all(texts)                      # every status has its own message
    assert so.hz_strerror(hz.HZ_OK) == b"ok"


def test_no_cuda_device_is_an_error_not_a_fallback(so, hz):
    if not _no_gpu(so):
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    assert so.hz_create(0, C.byref(ctx)) == hz.HZ_ERR_CUDA and not ctx.value
    with pytest.raises(hz.HzError):
        hz.Codec(0)


def test_cli_usage_and_errors(tmp_path, hz, so):
    hz.build_library()
    exe = os.path.join(PKG, "datacomp")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stdout + r.stderr and "decompress" in r.stdout + r.stderr
    r = subprocess.run([exe, "compress", str(tmp_path / "missing.bin"), str(tmp_path / "o.dcz")], capture_output=True, text=True)
    assert r.returncode == 1 and "Error:" in r.stdout + r.stderr
    if _no_gpu(so):
        src = tmp_path / "in.bin"
        src.write_bytes(b"A" * 2048)
        r = subprocess.run([exe, "c", str(src), str(tmp_path / "o.dcz")], capture_output=True, text=True)
        assert r.returncode == 1 and "Error:" in r.stdout + r.stderr and not (tmp_path / "o.dcz").exists()
