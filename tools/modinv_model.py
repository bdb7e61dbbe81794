"""Model of csrc/modinv.cuh — modular inversion by Bernstein-Yang divsteps ("safegcd"), variable time.

Inverts modulo an odd modulus M < 2^256 (p = 2^255 - 19 for field elements, l for scalars) in ~9 batches of 62
divsteps instead of a 254-squaring Fermat chain.  The arithmetic below mirrors the device code limb for limb
(five signed 62-bit limbs, 128-bit accumulators) and asserts every range the C code relies on.

    python tools/modinv_model.py
"""
import random

M62 = (1 << 62) - 1
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


def to_limbs(x):
    """non-negative x < 2^310 -> five 62-bit limbs (little-endian)"""
    return [(x >> (62 * i)) & M62 for i in range(5)]


def from_limbs(v):
    """limbs 0..3 in [0, 2^62), limb 4 signed"""
    return sum(v[i] << (62 * i) for i in range(5))


def s64(x):
    """wrap to a signed 64-bit integer"""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >> 63 else x


def s128(x):
    assert -(1 << 127) <= x < (1 << 127), "int128 overflow"
    return x


def divsteps_62(delta, f0, g0):
    """62 divsteps on the low 64 bits of f (odd) and g.  Returns delta and the matrix t = (u, v, q, r) with
    2^62 [f_62; g_62] = t [f_0; g_0]; every entry has absolute value <= 2^62."""
    u, v, q, r = 1, 0, 0, 1
    f, g = f0 & ((1 << 64) - 1), g0 & ((1 << 64) - 1)
    for _ in range(62):
        assert f & 1
        if g & 1:
            if delta > 0:
                delta, f, g, u, v, q, r = 1 - delta, g, (g - f) & ((1 << 64) - 1), q, r, q - u, r - v
            else:
                delta, g, q, r = 1 + delta, (g + f) & ((1 << 64) - 1), q + u, r + v
        else:
            delta = 1 + delta
        # g is even now: halve it (the low 64 bits suffice for the remaining steps), double the first row
        g >>= 1
        u, v = 2 * u, 2 * v
        assert max(abs(u), abs(v), abs(q), abs(r)) <= 1 << 62
    return delta, (u, v, q, r)


def update_fg(f, g, t):
    u, v, q, r = t
    cf = s128(u * f[0] + v * g[0])
    cg = s128(q * f[0] + r * g[0])
    assert cf & M62 == 0 and cg & M62 == 0
    cf >>= 62
    cg >>= 62
    for i in range(1, 5):
        cf = s128(cf + u * f[i] + v * g[i])
        cg = s128(cg + q * f[i] + r * g[i])
        f[i - 1], g[i - 1] = cf & M62, cg & M62
        cf >>= 62
        cg >>= 62
    f[4], g[4] = s64(cf), s64(cg)
    assert f[4] == cf and g[4] == cg


def update_de(d, e, t, mod, mod_inv62):
    """(d, e) <- t (d, e) / 2^62 mod M; inputs and outputs in (-2M, M)"""
    u, v, q, r = t
    sd, se = (-1 if d[4] < 0 else 0), (-1 if e[4] < 0 else 0)
    md = (u & sd) + (v & se)   # add M once per negative input: keeps the sums in range
    me = (q & sd) + (r & se)
    cd = s128(u * d[0] + v * e[0])
    ce = s128(q * d[0] + r * e[0])
    md -= (mod_inv62 * (cd & M62) + md) & M62
    me -= (mod_inv62 * (ce & M62) + me) & M62
    assert abs(md) < 1 << 63 and abs(me) < 1 << 63
    cd = s128(cd + mod[0] * md)
    ce = s128(ce + mod[0] * me)
    assert cd & M62 == 0 and ce & M62 == 0
    cd >>= 62
    ce >>= 62
    for i in range(1, 5):
        cd = s128(cd + u * d[i] + v * e[i] + mod[i] * md)
        ce = s128(ce + q * d[i] + r * e[i] + mod[i] * me)
        d[i - 1], e[i - 1] = cd & M62, ce & M62
        cd >>= 62
        ce >>= 62
    d[4], e[4] = s64(cd), s64(ce)
    assert d[4] == cd and e[4] == ce


def modinv(x, m):
    """x^-1 mod m (0 for x = 0 mod m), x < 2^256"""
    mod = to_limbs(m)
    mod_inv62 = pow(m, -1, 1 << 62)
    f, g = to_limbs(m), to_limbs(x % m)  # the device reduces x below m first (one conditional subtraction chain)
    d, e = [0] * 5, [1, 0, 0, 0, 0]
    delta = 1
    batches = 0
    while any(g):
        delta, t = divsteps_62(delta, f[0], g[0] | ((g[1] & 3) << 62))
        update_de(d, e, t, mod, mod_inv62)
        update_fg(f, g, t)
        assert -2 * m < from_limbs(d) < m and -2 * m < from_limbs(e) < m
        batches += 1
        assert batches <= 12
    fv = from_limbs(f)
    assert fv in (1, -1) or (fv in (m, -m) and x % m == 0)
    res = from_limbs(d)
    if fv < 0:
        res = -res
    res %= m
    if x % m == 0:
        res = 0
    return res, batches


def selftest(rounds=400, seed=3):
    rng = random.Random(seed)
    worst = 0
    for m in (P, L):
        edge = [0, 1, 2, m - 1, m - 2, (m + 1) // 2, 2**255, 2**256 - 1, 2**128, 2**64 - 1, 3, m, m + 1]
        for k in range(rounds):
            x = edge[k] if k < len(edge) else rng.getrandbits(256)
            got, b = modinv(x, m)
            worst = max(worst, b)
            assert got == (pow(x % m, -1, m) if x % m else 0), (hex(x), hex(m))
    return worst


if __name__ == "__main__":
    print("modinv model ok, at most", selftest(), "batches of 62 divsteps")
