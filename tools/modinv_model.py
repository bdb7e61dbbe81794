"""Model of csrc/modinv.cuh — modular inversion by Bernstein-Yang divsteps ("safegcd"), variable time, 32-bit words.

Inverts modulo an odd modulus M < 2^256 (p = 2^255 - 19 for field elements, l for scalars) in at most 20 batches of 30
divsteps instead of a 254-squaring Fermat chain.  The arithmetic mirrors the device code limb for limb (nine signed
30-bit limbs, 64-bit accumulators) and asserts every range the C code relies on; the multi-step batch routine
(count-trailing-zeros for runs of even steps, up to eight low bits of g cancelled at once while delta <= 0) is checked
against the step-by-step divstep recurrence.

    python tools/modinv_model.py
"""
import random

M30 = (1 << 30) - 1
M32 = (1 << 32) - 1
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
NL = 9


def to_limbs(x):
    return [(x >> (30 * i)) & M30 for i in range(NL)]


def from_limbs(v):
    """limbs 0..7 in [0, 2^30), limb 8 signed"""
    return sum(v[i] << (30 * i) for i in range(NL))


def s32(x):
    x &= M32
    return x - (1 << 32) if x >> 31 else x


def s64(x):
    assert -(1 << 63) <= x < (1 << 63), "int64 overflow"
    return x


def divsteps_30_plain(delta, f0, g0):
    """30 divsteps one at a time on the low 32 bits; 2^30 [f'; g'] = t [f; g]"""
    u, v, q, r = 1, 0, 0, 1
    f, g = f0 & M32, g0 & M32
    for _ in range(30):
        assert f & 1
        if g & 1:
            if delta > 0:
                delta, f, g, u, v, q, r = 1 - delta, g, (g - f) & M32, q, r, q - u, r - v
            else:
                delta, g, q, r = 1 + delta, (g + f) & M32, q + u, r + v
        else:
            delta = 1 + delta
        g >>= 1
        u, v = 2 * u, 2 * v
        assert max(abs(u), abs(v), abs(q), abs(r)) <= 1 << 30
    return delta, (u, v, q, r)


def divsteps_30(delta, f0, g0):
    """the device routine: runs of even steps by ctz, up to 8 low bits of g cancelled at once while delta <= 0"""
    u, v, q, r = 1, 0, 0, 1
    f, g = f0 & M32, g0 & M32
    i = 30
    while True:
        x = g | (1 << i)
        zeros = (x & -x).bit_length() - 1
        g >>= zeros
        u, v = (u << zeros) & M32, (v << zeros) & M32
        delta += zeros
        i -= zeros
        if i == 0:
            break
        if delta > 0:
            delta, f, g, u, v, q, r = -delta, g, (-f) & M32, q, r, (-u) & M32, (-v) & M32
        Lb = min(1 - delta, i, 8)
        finv = f
        finv = (finv * (2 - f * finv)) & M32
        finv = (finv * (2 - f * finv)) & M32
        w = (-(g * finv)) & ((1 << Lb) - 1)
        g = (g + f * w) & M32
        q = (q + u * w) & M32
        r = (r + v * w) & M32
    return delta, (s32(u), s32(v), s32(q), s32(r))


def update_fg(f, g, t):
    u, v, q, r = t
    cf = s64(u * f[0] + v * g[0])
    cg = s64(q * f[0] + r * g[0])
    assert cf & M30 == 0 and cg & M30 == 0
    cf >>= 30
    cg >>= 30
    for i in range(1, NL):
        cf = s64(cf + u * f[i] + v * g[i])
        cg = s64(cg + q * f[i] + r * g[i])
        f[i - 1], g[i - 1] = cf & M30, cg & M30
        cf >>= 30
        cg >>= 30
    assert -(1 << 31) <= cf < (1 << 31) and -(1 << 31) <= cg < (1 << 31)
    f[NL - 1], g[NL - 1] = cf, cg


def update_de(d, e, t, mod, inv30):
    """(d, e) <- t (d, e) / 2^30 mod M; inputs and outputs in (-2M, M)"""
    u, v, q, r = t
    sd, se = (-1 if d[NL - 1] < 0 else 0), (-1 if e[NL - 1] < 0 else 0)
    md = (u & sd) + (v & se)
    me = (q & sd) + (r & se)
    cd = s64(u * d[0] + v * e[0])
    ce = s64(q * d[0] + r * e[0])
    md -= (inv30 * (cd & M32) + md) & M30
    me -= (inv30 * (ce & M32) + me) & M30
    assert abs(md) < 1 << 31 and abs(me) < 1 << 31
    cd = s64(cd + mod[0] * md)
    ce = s64(ce + mod[0] * me)
    assert cd & M30 == 0 and ce & M30 == 0
    cd >>= 30
    ce >>= 30
    for i in range(1, NL):
        cd = s64(cd + u * d[i] + v * e[i] + mod[i] * md)
        ce = s64(ce + q * d[i] + r * e[i] + mod[i] * me)
        d[i - 1], e[i - 1] = cd & M30, ce & M30
        cd >>= 30
        ce >>= 30
    assert -(1 << 31) <= cd < (1 << 31) and -(1 << 31) <= ce < (1 << 31)
    d[NL - 1], e[NL - 1] = cd, ce


def modinv(x, m):
    """x^-1 mod m (0 for x = 0 mod m), x < 2^256"""
    mod = to_limbs(m)
    inv30 = pow(m, -1, 1 << 30)
    f, g = to_limbs(m), to_limbs(x)  # x may exceed m, as on the device
    d, e = [0] * NL, [1] + [0] * (NL - 1)
    delta = 1
    batches = 0
    while any(g):
        delta, t = divsteps_30(delta, f[0] | ((f[1] & 3) << 30), g[0] | ((g[1] & 3) << 30))
        update_de(d, e, t, mod, inv30)
        update_fg(f, g, t)
        assert -2 * m < from_limbs(d) < m and -2 * m < from_limbs(e) < m
        batches += 1
        assert batches <= 24
    fv = from_limbs(f)
    if fv not in (1, -1):
        assert x % m == 0
        return 0, batches
    res = from_limbs(d)
    if fv < 0:
        res = -res
    return res % m, batches


def selftest(rounds=400, seed=3):
    rng = random.Random(seed)
    for _ in range(rounds * 20):  # the multi-step batch is the step-by-step recurrence
        dl, f, g = rng.randint(-40, 40), rng.getrandbits(32) | 1, rng.getrandbits(32)
        if rng.random() < 0.2:
            g = rng.getrandbits(6) << rng.randrange(27)
        assert divsteps_30(dl, f, g) == divsteps_30_plain(dl, f, g)
    worst = 0
    for m in (P, L):
        edge = [0, 1, 2, m - 1, m - 2, (m + 1) // 2, 2**255, 2**256 - 1, 2**128, 2**64 - 1, 3, m, m + 1, 2 * m, 2**30, 2**30 - 1]
        for k in range(rounds):
            x = edge[k] if k < len(edge) else rng.getrandbits(256)
            got, b = modinv(x, m)
            worst = max(worst, b)
            assert got == (pow(x % m, -1, m) if x % m else 0), (hex(x), hex(m))
    return worst


if __name__ == "__main__":
    print("modinv model ok, at most", selftest(), "batches of 30 divsteps")
