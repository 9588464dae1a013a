"""Developer tool: per-subsequence (entry, exit, count) of the fused decoder against the Python model for ONE chunk."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
import datasets, fused_model as fm
n, chunk, H, k = (int(x) for x in sys.argv[1].split(":"))
data = datasets.zipf_stream(n, H, seed=H + 1)[k * chunk:(k + 1) * chunk]
os.environ["HZ_FU_DUMP"] = "/tmp/fu_dump.bin"
hz = ge.load_package()
c = hz.Codec(0)
payload, off, lens = c.encode(data, chunk)[:3]
back = c.decode(payload, off[:-1], np.diff(off).astype(np.uint32), np.array([data.size], dtype=np.uint32), lens)
print("mismatches", int((back != data).sum()))
d = np.fromfile("/tmp/fu_dump.bin", dtype=np.uint32)
S = min(17, max(4, payload.size * 32 // data.size)); lead = min(4, max(2, -(-payload.size * 7 // data.size)))
A = fm.canonical(lens[0]); tab = fm.build_table(A); rd = fm.Bits(payload)
sub = S * 32; nsub = max(1, (payload.size * 8 + sub - 1) // sub)
prev_exit = 0; shown = 0
for i in range(nsub):
    nominal = i * sub
    b, s = fm.walk(A, tab, rd, nominal + prev_exit, nominal + sub)
    ex = b - (nominal + sub)
    g = int(d[i]); ge_, gx, gc = g & 0xFF, (g >> 8) & 0xFF, g >> 16
    if (ge_, gx, gc) != (prev_exit, ex, len(s)) and shown < 12:
        lb = fm.walk(A, tab, rd, nominal - lead * 32, nominal)[0] - nominal if i else 0
        print("sub %d (unit %d lane %d): gpu entry/exit/count %d %d %d  model %d %d %d  model guess %d" % (i, i // 32, i % 32, ge_, gx, gc, prev_exit, ex, len(s), lb))
        shown += 1
    prev_exit = ex
# trace the first bad subsequence lookup by lookup
prev_exit = 0; obase = 0
for i in range(nsub):
    nominal = i * sub
    b, s = fm.walk(A, tab, rd, nominal + prev_exit, nominal + sub)
    got = back[obase:obase + len(s)]
    if not np.array_equal(got, np.array(s, dtype=np.uint8)):
        print("first bad sub %d obase %d entry %d" % (i, obase, prev_exit))
        pos = nominal + prev_exit; o = obase
        while pos < nominal + sub:
            v = rd.peek32(pos)
            syms, l, inv = fm.lookup(A, tab, v)
            kind = tab[v >> 20][0]
            g = back[o:o + len(syms)]
            mark = "" if list(g) == syms else "   <-- got %s" % list(g)
            print("  pos %4d (mod32 %2d) kind %d l %2d syms %s%s" % (pos - nominal, pos % 32, kind, l, syms, mark))
            if mark and kind == 2:
                print("     entry lmin/lmax", tab[v >> 20][4:], "v=%08x" % v)
            pos += l; o += len(syms)
        break
    obase += len(s); prev_exit = b - (nominal + sub)
