"""Emulation (numpy) of the two-stage row median of csrc/dstr_rows_mma.cuh (coarse bisection on fp16 images of the scaled values,
exact finish on the float keys) against np.median of the zero-filled row."""
import numpy as np

def f2key(f):
    b = np.float32(f).view(np.uint32)
    return np.uint32(~b) if b & 0x80000000 else np.uint32(b | 0x80000000)
def keys_of(x):
    b = x.view(np.uint32)
    return np.where(b & 0x80000000, ~b, b | np.uint32(0x80000000)).astype(np.uint32)
def key2f(k):
    k = np.uint32(k)
    b = (k & np.uint32(0x7fffffff)) if (k & 0x80000000) else np.uint32(~k)
    return np.uint32(b).view(np.float32)
def t_of(k16):
    bits = (k16 & 0x7fff) if (k16 & 0x8000) else ((~k16) & 0xffff)
    return np.uint16(bits).view(np.float16)

def median_two_stage(c, thr_q, thr, stats):
    n = c.size
    m = (c * c) > thr_q
    x = np.where(m, np.float32(0), c + np.float32(0)).astype(np.float32)
    key = keys_of(x)
    KZ = np.uint32(0x80000000)
    k1, k2 = (n - 1) >> 1, n >> 1
    cneg = int((key < KZ).sum()); cle0 = int((key <= KZ).sum())
    if cneg <= k1 and k2 < cle0:
        stats['zero'] += 1
        return np.float32(0)
    e = np.frexp(np.float32(thr))[1]
    scale = np.float32(np.ldexp(1.0, max(-24, min(14 - e, 40))))
    with np.errstate(over='ignore'):
        h = (x * scale).astype(np.float16)
    res16, lo_cnt, hi16, hi_cnt = 0, 0, 0xFC00, n
    steps = 0
    for b in range(15, -1, -1):
        if hi_cnt - lo_cnt == 1: break
        trial = res16 | (1 << b)
        t = t_of(trial)
        assert not np.isnan(t), hex(trial)
        cnt = int((h < t).sum()); steps += 1
        if cnt <= k1: res16, lo_cnt = trial, cnt
        else: hi16, hi_cnt = trial, cnt
    stats['steps'] += steps
    lo_t, hi_t = t_of(res16), t_of(hi16)
    cand = (h >= lo_t) & (h < hi_t)
    mcnt = int(cand.sum()); assert mcnt == hi_cnt - lo_cnt and mcnt >= 1, (mcnt, hi_cnt, lo_cnt)
    r = k1 - lo_cnt
    assert 0 <= r < mcnt
    cmin, cmax = key[cand].min(), key[cand].max()
    if r == 0: kk1 = cmin
    elif r == mcnt - 1: kk1 = cmax
    else:
        stats['bisect'] += 1
        lo, hi = int(cmin), int(cmax)
        while lo < hi:
            mid = lo + (hi - lo) // 2
            if int((key <= np.uint32(mid)).sum()) > k1: hi = mid
            else: lo = mid + 1
        kk1 = np.uint32(lo)
    stats['m'] = max(stats['m'], mcnt)
    med = key2f(kk1)
    if k2 != k1:
        cle = int((key <= kk1).sum())
        nxt = key[key > kk1].min() if (key > kk1).any() else np.uint32(0xffffffff)
        kk2 = kk1 if cle >= k1 + 2 else nxt
        med = (key2f(kk1) + key2f(kk2)) * np.float32(0.5)
    return np.float32(med), x

rng = np.random.default_rng(1)
stats = dict(zero=0, steps=0, bisect=0, m=0)
nrows = 0
for trial in range(4000):
    n = int(rng.choice([16, 17, 33, 64, 129, 260, 515, 1002, 1026]))
    kind = trial % 8
    if kind == 0: c = rng.standard_normal(n).astype(np.float32) * 0.05
    elif kind == 1: c = (rng.standard_normal(n) * 0.05 + 0.02).astype(np.float32)
    elif kind == 2: c = (rng.standard_normal(n) * 1e-6 - 3e-3).astype(np.float32)          # many values inside one fp16 ulp
    elif kind == 3: c = rng.choice(np.array([-0.01, 0.01, 0.02, 0.0200001, 0.03], np.float32), n)  # ties
    elif kind == 4: c = (rng.standard_normal(n) * 1e-9).astype(np.float32) + np.float32(1e-7)  # far below the scale
    elif kind == 5: c = np.abs(rng.standard_normal(n)).astype(np.float32) * 0.04
    elif kind == 6: c = -np.abs(rng.standard_normal(n)).astype(np.float32) * 0.04 + np.float32(1e-8)
    else: c = rng.standard_cauchy(n).astype(np.float32) * 0.01
    thr = np.float32(abs(rng.standard_normal()) * 0.05 + 0.02) if kind != 7 else np.float32(0.5)
    thr_q = np.float32(thr * thr)
    out = median_two_stage(c, thr_q, thr, stats)
    if isinstance(out, tuple): med, x = out
    else:
        med = out; x = np.where((c * c) > thr_q, np.float32(0), c).astype(np.float32)
    ref = np.median(x.astype(np.float32))
    s = np.sort(x); ref2 = (s[(n - 1) >> 1] + s[n >> 1]) * np.float32(0.5) if n % 2 == 0 else s[n >> 1]
    assert med == np.float32(ref2), (trial, kind, n, med, ref2)
    nrows += 1
print("ok", nrows, stats)
