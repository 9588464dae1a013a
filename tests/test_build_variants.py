"""Build-level guards that need no GPU: the compile-time A/B switches of the kernels still compile for sm_100a, and
the shipped objects hold the instructions DESIGN.md claims (TMA bulk copies + mbarriers in the decoder, 256-bit
streaming loads in the encoder, shared-memory atomics in the histogram, ballots in the codebook replay).  SASS
mnemonics as listed in the B200 profiling recipe; `cuobjdump` reads the objects `make` left in build/."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "data-compression-implementing-gpu-driven-huffman-encoding-in-java_b200")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(NVCC) or not os.path.exists(CUOBJDUMP), reason="CUDA toolkit not installed")


def _sass(obj):
    return subprocess.check_output([CUOBJDUMP, "-sass", obj], text=True)


@pytest.fixture(scope="module")
def objects(hz):
    hz.build_library()
    return os.path.join(PKG, "build")


def test_shipped_objects_hold_the_claimed_instructions(objects):
    for unit in ("hz_decode.o", "hz_decode_fused.o"):
        dec = _sass(os.path.join(objects, unit))
        assert "UBLKCP.S.G" in dec and "UBLKCP.G.S" in dec          # cp.async.bulk global -> shared (staging) and back (windows)
        assert "SYNCS.ARRIVE.TRANS64" in dec and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in dec    # mbarrier expect_tx / try_wait
    enc = _sass(os.path.join(objects, "hz_encode.o"))
    assert "256.CONSTANT" in enc and "STG.E.128" in enc         # 256-bit streaming loads, 128-bit stores
    hist = _sass(os.path.join(objects, "hz_hist.o"))
    assert "ATOMS" in hist and "128" in hist                    # shared-memory atomics fed by 128-bit loads
    cb = _sass(os.path.join(objects, "hz_codebook.o"))
    assert cb.count("VOTE") >= 20 and "POPC" in cb              # warp_heap_replay: ballots + popcount
    # the chained encode: every histogram CTA releases the dependent launch (griddepcontrol.launch_dependents = PREEXIT),
    # look-back and ready flags are acquire loads / release stores at GPU scope, the encoder polls its chunk's flag
    assert "PREEXIT" in cb and "LDG.E.64.STRONG.GPU" in cb and "NANOSLEEP" in cb
    assert "LDG.E.64.STRONG.GPU" in enc and "NANOSLEEP" in enc and "PREEXIT" not in enc
    fused = _sass(os.path.join(objects, "hz_decode_fused.o"))
    for shape in ("ILi24ELi1ELi1E", "ILi8ELi2ELi1E", "ILi5ELi3ELi1E", "ILi24ELi1ELi2E"):   # CTA shapes of the fused decoder: 24 x 1, 8 x 2,
        assert "dec_fused_kernel" + shape in fused                                          # 5 x 3 warps, and 24 x 1 as cluster pairs
    assert fused.count("UCGABAR_ARV") == 2 and fused.count("UCGABAR_WAIT") == 2           # the pair's two cluster barriers (start, exit)
    for text in (dec, enc, hist, cb):
        assert "sm_100a" in text or "SM100" in text.upper() or "EF_CUDA_SM100" in text


def _spills(ptxas_log):
    """entry function -> bytes of spill stores, from `-Xptxas -v` output"""
    out, name = {}, None
    for line in ptxas_log.splitlines():
        if "Compiling entry function" in line:
            name = line.split("'")[1]
        elif "spill stores" in line and name:
            out[name] = int(line.split("stack frame,")[1].split("bytes spill stores")[0])
    return out


@pytest.mark.parametrize("src,flag,hot,expect", [
    ("hz_codebook.cu", "-DHZ_CB_PLAIN", ("codebook_kernel", "codebook_warp_kernel"), None),     # the literal one-thread sift loops
    ("hz_decode_fused.cu", "-DFU_CHECK", (), None),                                              # own bounds assertions
    ("hz_decode_fused.cu", "-DFU_TIMING", (), None),                                             # per-phase clocks
])
def test_ab_switches_compile(tmp_path, src, flag, hot, expect):
    obj = str(tmp_path / (src + ".o"))
    r = subprocess.run([NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", flag,
                        "-Xptxas", "-v", "-c", os.path.join(PKG, "csrc", src), "-o", obj], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    spills = _spills(r.stderr)
    for kernel in hot:
        hits = [v for k, v in spills.items() if kernel in k]
        assert hits and all(v == 0 for v in hits), "%s spills: %s" % (kernel, spills)
    if expect:
        assert _sass(obj).count(expect) >= 20


def test_hot_kernels_of_the_shipped_build_do_not_spill(objects):
    hot = {"hz_hist": ("hist_seg_lanes",), "hz_codebook": ("codebook_kernel", "codebook_warp_kernel", "codebook_lane_kernel"),   # (hist_chain_kernel is capped at 40 registers for 6 CTAs per SM: its codebook tail spills one word)
           "hz_encode": ("encode_kernel",), "hz_decode": ("dec_sync_kernel", "dec_write_kernel")}
    # (dec_fused_kernel is register-capped at 80 by its 768-thread CTA; its compaction phase spills a few words)
    for unit, kernels in hot.items():
        spills = _spills(open(os.path.join(objects, unit + ".ptxas.log")).read())
        for kernel in kernels:
            hits = [v for k, v in spills.items() if kernel in k]
            assert hits and all(v == 0 for v in hits), "%s spills: %s" % (kernel, spills)
