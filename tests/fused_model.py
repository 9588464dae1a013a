"""Python model of the fused decoder's algorithm (csrc/hz_decode_fused.cu): FMT-1 lookup table, the packed-counter
walks, fu_settle, the per-subsequence lead-in guess and the chain of entry == previous exit links.  It is the
CPU-side pin of the ALGORITHM (table contents, settle rule, chain repair); the kernel itself is checked against
the oracle on the GPU."""
import numpy as np

LUTB = 12


def canonical(lens):
    lens = np.asarray(lens, dtype=np.int64)
    count = np.bincount(lens[lens > 0], minlength=34)
    first = np.zeros(34, dtype=np.int64); offs = np.zeros(34, dtype=np.int64); lim = np.zeros(34, dtype=np.int64)
    symbase = np.zeros(34, dtype=np.int64)
    c = 0; o = 0
    for L in range(1, 33):
        c = (c + (count[L - 1] if L > 1 else 0)) << 1
        first[L] = c; offs[L] = o; o += count[L]
        symbase[L] = offs[L] - c
        lim[L] = (c + count[L]) << (32 - L)
    sorted_syms = np.array(sorted(np.nonzero(lens)[0], key=lambda s: (lens[s], s)), dtype=np.int64)
    maxlen = int(lens.max())
    return dict(lens=lens, count=count, first=first, lim=lim, symbase=symbase, sorted=sorted_syms, maxlen=maxlen)


def long_len(A, v, lmin, lmax):
    if lmin == 0:
        return 0
    l = lmin
    while l < lmax and v >= A["lim"][l]:
        l += 1
    return l if v < A["lim"][l] else 0


def build_table(A):
    """-> list of (kind, syms, ltot, n, lmin, lmax); kind 0 normal, 1 single-length long, 2 rare."""
    lens = A["lens"]
    base = [None] * (1 << LUTB)
    lim12 = [int(A["lim"][L] >> (32 - LUTB)) for L in range(0, LUTB + 1)]
    for x in range(1 << LUTB):
        li = 1
        for L in range(1, LUTB + 1):
            li += x >= lim12[L]
        if li <= LUTB:
            base[x] = (int(A["sorted"][A["symbase"][li] + (x >> (LUTB - li))]), li)
    tab = []
    for x in range(1 << LUTB):
        e0 = base[x]
        if e0:
            syms = [e0[0]]; used = e0[1]; wtot = used; cur = x; lprev = e0[1]
            while True:
                cur = (cur << lprev) & ((1 << LUTB) - 1)
                e = base[cur]
                if not e or used + e[1] > LUTB:
                    break
                if len(syms) < 4:
                    syms.append(e[0]); wtot = used + e[1]
                used += e[1]; lprev = e[1]
            tab.append((0, syms, wtot, len(syms), 0, 0))
        else:
            kind = (2, [], 0, 0, 0, 0)
            if A["maxlen"] > LUTB:
                vlo = x << (32 - LUTB); vhi = vlo | ((1 << (32 - LUTB)) - 1)
                lmin = long_len(A, vlo, LUTB + 1, A["maxlen"])
                if lmin:
                    lmax = long_len(A, vhi, lmin, A["maxlen"]) or A["maxlen"]
                    kind = (1, [], lmin, 1, lmin, lmin) if (lmin == lmax and lmin <= 24) else (2, [], 0, 0, lmin, lmax)
            tab.append(kind)
    return tab


class Bits:
    def __init__(self, comp):
        self.bits = np.unpackbits(np.asarray(comp, dtype=np.uint8))

    def peek32(self, pos):
        b = self.bits[pos:pos + 32] if pos >= 0 else np.zeros(0, np.uint8)
        v = 0
        for x in b:
            v = (v << 1) | int(x)
        return v << (32 - len(b))


def lookup(A, tab, v):
    """one table lookup at the 32 stream bits v -> (symbols, bits consumed, invalid)"""
    kind, syms, ltot, n, lmin, lmax = tab[v >> (32 - LUTB)]
    if kind == 0:
        return syms, ltot, False
    if kind == 1:
        return [int(A["sorted"][A["symbase"][ltot] + (v >> (32 - ltot))])], ltot, False
    l = long_len(A, v, lmin, lmax)
    if not l:
        return [0], 1, True
    return [int(A["sorted"][A["symbase"][l] + (v >> (32 - l))])], l, False


def walk(A, tab, rd, pos, limit):
    """kernel's fu_walk / fu_skim + fu_settle: decode from pos until the first codeword boundary >= limit.
    -> (boundary, symbols that begin before limit)"""
    out = []
    while True:
        syms, l, _ = lookup(A, tab, rd.peek32(pos))
        if pos + l >= limit:
            if len(syms) > 1 and pos + l > limit:           # fu_settle
                q = pos; j = 0
                while True:
                    q += int(A["lens"][syms[j]]); j += 1
                    if not (q < limit and j < len(syms)):
                        break
                return q, out + syms[:j]
            return pos + l, out + syms
        out += syms; pos += l


def decode_chunk(comp, lens, osize, S, lead):
    """subsequences of S words, lead-in of `lead` words; -> decoded bytes, number of chain repairs"""
    A = canonical(lens); tab = build_table(A); rd = Bits(comp)
    sub = S * 32
    nsub = max(1, (len(comp) * 8 + sub - 1) // sub)
    entry = [0] * nsub; exitv = [0] * nsub; syms = [None] * nsub
    for i in range(nsub):
        nominal = i * sub
        if i:
            b, _ = walk(A, tab, rd, nominal - lead * 32, nominal)
            entry[i] = b - nominal
    repairs = 0
    end = lambda i: min(i * sub + sub, len(comp) * 8)         # the chunk's last subsequence ends with the chunk
    for i in range(nsub):
        nominal = i * sub
        b, s = walk(A, tab, rd, nominal + entry[i], end(i))
        exitv[i] = b - end(i); syms[i] = s
    for i in range(1, nsub):                                 # chain repair, in stream order
        if entry[i] != exitv[i - 1]:
            repairs += 1
            entry[i] = exitv[i - 1]
            nominal = i * sub
            b, s = walk(A, tab, rd, nominal + entry[i], end(i))
            exitv[i] = b - end(i); syms[i] = s
    out = [x for s in syms for x in s]
    if len(out) < osize:
        out += [int(A["sorted"][0])] * (osize - len(out))
    return np.array(out[:osize], dtype=np.uint8), repairs
