"""Developer tool: GPU encode -> GPU decode on Zipf streams, reports the first mismatching positions per chunk.
python tools/dbg_decode.py n:chunk:H [...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import datasets
hz = ge.load_package()
c = hz.Codec(0)
for a in sys.argv[1:]:
    n, chunk, H = (int(x) for x in a.split(":"))
    data = datasets.zipf_stream(n, H, seed=H + 1)
    payload, off, lens = c.encode(data, chunk)[:3]
    K = len(off) - 1
    sizes = np.diff(off).astype(np.uint32)
    orig = np.array([min(chunk, n - k * chunk) for k in range(K)], dtype=np.uint32)
    try:
        back = c.decode(payload, off[:-1], sizes, orig, lens)
    except Exception as e:
        print(a, "EXC", e); continue
    bad = np.nonzero(back != data)[0]
    print(a, "K=%d" % K, "mismatches=%d" % bad.size, "b/sym=%.3f" % (8 * off[-1] / max(1, n)))
    if bad.size:
        ks = np.unique(bad // chunk)
        for k in ks[:6]:
            b = bad[(bad // chunk) == k] - k * chunk
            print("   chunk %d csize %d: %d bad, first %s last %d" % (k, sizes[k], b.size, b[:8], b[-1]))
    if bad.size and os.environ.get("DBG_DETAIL"):
        k = int(bad[0] // chunk)
        ln = lens[k]
        b0 = int(bad[0])
        print("   first bad at %d (chunk %d + %d): expected syms %s lens %s" % (b0, k, b0 - k * chunk, data[b0 - 4:b0 + 8], ln[data[b0 - 4:b0 + 8]]))
        print("   got %s lens %s" % (back[b0 - 4:b0 + 8], ln[back[b0 - 4:b0 + 8]]))
        # runs of bad
        d = np.diff(bad); starts = np.concatenate([[0], np.nonzero(d > 1)[0] + 1])
        for s in starts[:10]:
            e = s
            while e + 1 < bad.size and bad[e + 1] == bad[e] + 1: e += 1
            p = int(bad[s])
            print("   run at %d len %d: exp lens %s got lens %s" % (p - k * chunk, e - s + 1, ln[data[p - 2:p + 4]], ln[back[p - 2:p + 4]]))
