"""CPU checks of the arithmetic ALGORITHMS behind two device shortcuts (the device code itself is checked on the GPU:
test_unchecked_division_is_ieee, the two-layer parity cases):

* dvd_nr / sqrt_nr (csrc/hb_device.cuh): the IEEE fast paths without the range check. Restated here with exact rational
  arithmetic (every fma rounded once, in the stated mode) on top of ANY reciprocal / reciprocal square root within one
  ulp of the true value -- the accuracy MUFU.RCP / MUFU.RSQ guarantee -- and compared with the correctly rounded
  quotient / root, signed zeros included.
* feistel_domain: the count-leading-zeros form of the reference's counting loop (pcg_shared.h:550-603)."""
import math
import struct
from fractions import Fraction

import numpy as np


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def bits(x):
    return struct.unpack("I", struct.pack("f", x))[0]


def from_bits(b):
    return struct.unpack("f", struct.pack("I", b & 0xFFFFFFFF))[0]


def round_fraction(q, mode):
    """Exact rational (non-zero, well inside the normal range) -> binary32, 'rn' (ties to even) or 'rm' (toward -inf)."""
    sign = -1 if q < 0 else 1
    a = abs(q)
    e = math.floor(math.log2(a))
    if Fraction(2) ** e > a:
        e -= 1
    if Fraction(2) ** (e + 1) <= a:
        e += 1
    ulp = Fraction(2) ** (e - 23)
    k = a / ulp                      # in [2^23, 2^24)
    lo = k.numerator // k.denominator
    rem = k - lo
    if mode == "rn":
        up = rem > Fraction(1, 2) or (rem == Fraction(1, 2) and lo % 2 == 1)
    else:                            # toward -inf: magnitudes grow for negative values
        up = rem > 0 and sign < 0
    return sign * float((lo + (1 if up else 0)) * ulp)


def fma(a, b, c, mode="rn"):
    """One correctly rounded a*b + c on binary32 values, IEEE signed zeros."""
    exact = Fraction(a) * Fraction(b) + Fraction(c)
    if exact != 0:
        return round_fraction(exact, mode)
    prod_neg = (math.copysign(1.0, a) * math.copysign(1.0, b)) < 0
    c_neg = math.copysign(1.0, c) < 0
    if Fraction(a) * Fraction(b) != 0 or Fraction(c) != 0:      # x + (-x): +0, or -0 when rounding down
        return -0.0 if mode == "rm" else 0.0
    if prod_neg == c_neg:                                        # (+-0) + (+-0) of like sign
        return -0.0 if c_neg else 0.0
    return -0.0 if mode == "rm" else 0.0


def mul(a, b):
    exact = Fraction(a) * Fraction(b)
    if exact == 0:
        return math.copysign(0.0, math.copysign(1.0, a) * math.copysign(1.0, b))
    return round_fraction(exact, "rn")


def dvd_nr(a, b, r0):
    """hb_device.cuh dvd_nr with r0 standing for rcp.approx(b)."""
    r = fma(r0, fma(-b, r0, 1.0), r0)
    q = mul(a, r)
    return fma(r, fma(-b, q, a, "rm"), q)


def sqrt_nr(x, y0):
    s, h = mul(x, y0), mul(y0, 0.5)
    return fma(fma(-s, s, x), h, s)


def nudge(v, ulps):
    return from_bits(bits(v) + ulps)


def test_unchecked_division_algorithm_is_correctly_rounded():
    rng = np.random.default_rng(2026)
    n = 4000
    mant = rng.integers(0, 1 << 23, size=(n, 2))
    expo = rng.integers(127 - 30, 127 + 30, size=(n, 2))
    for i in range(n):
        a = from_bits((int(expo[i, 0]) << 23) | int(mant[i, 0]))
        b = from_bits((int(expo[i, 1]) << 23) | int(mant[i, 1]))
        if i % 3 == 0:
            a = -a
        if i % 7 == 0:                                           # the neighbourhood of exact quotients
            a = f32(b * float(rng.integers(1, 1 << 12)))
        want = round_fraction(Fraction(a) / Fraction(b), "rn")
        r_true = round_fraction(1 / Fraction(b), "rn")
        for err in (-1, 0, 1):                                   # any reciprocal within one ulp
            got = dvd_nr(a, b, nudge(r_true, err))
            assert bits(got) == bits(want), (a, b, err, got, want)


def test_unchecked_division_keeps_signed_zero_numerators():
    for b in (1e-5, 0.3, 1.0, 7.25e3):
        b = f32(b)
        r0 = round_fraction(1 / Fraction(b), "rn")
        for a in (0.0, -0.0):
            got = dvd_nr(a, b, r0)
            assert got == 0.0 and math.copysign(1.0, got) == math.copysign(1.0, a), (a, b, got)


def test_unchecked_sqrt_algorithm_is_correctly_rounded():
    rng = np.random.default_rng(7)
    for i in range(3000):
        x = from_bits((int(rng.integers(127 - 40, 127 + 40)) << 23) | int(rng.integers(0, 1 << 23)))
        if i % 5 == 0:                                           # perfect squares and their neighbours
            k = float(rng.integers(1, 1 << 11))
            x = nudge(f32(k * k), int(rng.integers(-1, 2)))
        want = f32(math.sqrt(x))                                 # binary64 sqrt rounded to binary32 is correctly rounded
        y_true = round_fraction(1 / Fraction(math.sqrt(x)), "rn")
        for err in (-1, 0, 1):
            got = sqrt_nr(x, nudge(y_true, err))
            assert bits(got) == bits(want), (x, err, got, want)


def test_feistel_domain_bits_by_clz_equal_the_counting_loop():
    def loop(n):
        b = 0
        while b < 30 and (1 << b) < n:
            b += 1
        return b + (b & 1)

    def clz32(v):
        return 32 - v.bit_length()

    def by_clz(n):
        b = 0 if n <= 1 else min(32 - clz32(n - 1), 30)
        return b + (b & 1)

    for n in list(range(0, 5000)) + [(1 << k) + d for k in range(2, 32) for d in (-1, 0, 1)] + [0xFFFFFFF0]:
        assert loop(n) == by_clz(n), n


def test_entry_face_pick_by_counting_equals_the_sequential_scan():
    """pick_entry_triangle (csrc/hb_kernels.cuh): 'first group whose cumulative weight exceeds the target' found by
    counting the cumulative weights <= target (descending pass with predicated moves) gives the group, residual and
    group weight of the sequential scan -- zero-weight groups, fewer than 8 groups and target == total included."""
    rng = np.random.default_rng(11)
    f = np.float32

    def sequential(w, ng, u):
        total = f(0)
        for g in range(ng):
            total = f(total + w[g])
        target = f(u * total)
        sel, resid, w_sel, cum = ng - 1, f(0), f(0), f(0)
        for g in range(ng):
            c1 = f(cum + w[g])
            if c1 > target:
                return g, f(target - cum), w[g]
            cum = c1
        return sel, resid, w_sel

    def counting(w, ng, u):
        c, total = [], f(0)
        for g in range(8):
            if g < ng:
                total = f(total + w[g])
            c.append(total)
        target = f(u * total)
        below, lo, w_sel = 0, f(0), f(0)
        for g in range(7, -1, -1):
            le = c[g] <= target
            w_sel = w_sel if le else (w[g] if g < ng else f(0))
            lo = max(lo, c[g]) if le else lo
            below += 1 if le else 0
        if below < ng:
            return min(below, ng - 1), f(target - lo), w_sel
        return ng - 1, f(0), f(0)

    for trial in range(20000):
        ng = int(rng.integers(1, 9))
        w = rng.random(8).astype(np.float32) * (rng.random(8) > 0.35)        # ~a third of the groups face away
        w = np.where(np.arange(8) < ng, w, 0).astype(np.float32)
        if not w[:ng].sum() > 0:
            continue
        u = f(rng.random()) if trial % 50 else f(1.0 - 2.0 ** -24)          # the largest uniform the RNG produces
        a, b = sequential(w, ng, u), counting(w, ng, u)
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2], (w, ng, u, a, b)


def test_warp_pool_feistel_resolution_equals_per_item_walks():
    """resolve_sources (csrc/hb_kernels.cuh) restated lane by lane: a warp's pool of 32 x batch items, lanes drawing the
    next unresolved item whenever theirs lands below n, resolves every item exactly once and to the value of the plain
    per-item cycle walk (feistel(): at most 64 walks, then modulo) -- ragged last iteration and tiny pools included."""
    M32 = 0xFFFFFFFF

    def pcg_hash(x):
        x = (x * 747796405 + 2891336453) & M32
        x = ((((x >> ((x >> 28) + 4)) ^ x) & M32) * 277803737) & M32
        return ((x >> 22) ^ x) & M32

    def domain(n):
        b = 0 if n <= 1 else min((n - 1).bit_length(), 30)
        b += b & 1
        return b >> 1, (1 << (b >> 1)) - 1

    def walk(cur, hb, hm, seed):
        L, R = (cur >> hb) & hm, cur & hm
        for rc in (0x9E3779B9, 0x85EBCA6B, 0xC2B2AE35, 0x27D4EB2F):
            L, R = R, L ^ (pcg_hash(seed ^ R ^ rc) & hm)
        return ((L << hb) | R) & M32

    def feistel(i, n, seed):
        hb, hm = domain(n)
        cur = i
        for _ in range(64):
            cur = walk(cur, hb, hm, seed)
            if cur < n:
                return cur
        return cur % n

    def resolve(k_warp, stride, count, cont_first, n, seed, batch):
        hb, hm = domain(n)
        total = sum(min(32, count - (k_warp + c * stride)) for c in range(batch) if k_warp + c * stride < count)
        out, nxt = {}, 0
        item, cur, walks = [None] * 32, [0] * 32, [0] * 32
        while True:
            need = [lane for lane in range(32) if item[lane] is None]
            avail = total - nxt
            for r, lane in enumerate(need):
                if r < avail:
                    item[lane] = nxt + r
                    cur[lane] = cont_first + k_warp + (item[lane] >> 5) * stride + (item[lane] & 31)
                    walks[lane] = 0
            nxt += min(len(need), avail)
            if all(it is None for it in item):
                return out, total
            for lane in range(32):
                if item[lane] is not None:
                    cur[lane] = walk(cur[lane], hb, hm, seed)
                    walks[lane] += 1
                    if cur[lane] < n or walks[lane] == 64:
                        assert item[lane] not in out
                        out[item[lane]] = cur[lane] if cur[lane] < n else cur[lane] % n
                        item[lane] = None

    for (count, n, first, stride, batch) in [(1000, 1000, 0, 256, 4), (777, 5000, 4000, 64, 16), (33, 40, 3, 32, 16),
                                             (5, 3, 0, 32, 4), (4096, 70000, 1234, 512, 16)]:
        seed = 0xB17CA3D9 ^ count
        for k_warp in range(0, min(count, stride), 32):
            base = k_warp
            while base < count:
                out, total = resolve(base, stride, count, first, n, seed, batch)
                assert len(out) == total
                for it, src in out.items():
                    k = base + (it >> 5) * stride + (it & 31)
                    assert k < count and src == feistel(first + k, n, seed), (count, n, k)
                base += batch * stride
