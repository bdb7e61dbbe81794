"""Lane-level model of csrc/fe8.cuh — GF(2^255-19) arithmetic with ONE 32-bit word per lane, 8 lanes per element.

The CUDA code was written from this model; every intermediate is masked to the register width it has on the
device (32 / 64 bits) and every bound the comments in fe8.cuh claim is asserted here, on random and edge inputs:

    python tools/fe8_model.py          # self-test (also run by tests/test_fe8_model.py)

Representation ("tight"): value = sum w_j 2^(32 j), w_j < 2^32, w_7 <= 2^31 + 2^8  (so value < 2^255 + 2^233).
"""
import random

P = 2**255 - 19
M32, M64 = 2**32 - 1, 2**64 - 1
W7_MAX = 2**31 + 2**8


def to_lanes(v):
    return [(v >> (32 * j)) & M32 for j in range(8)]


def from_lanes(w):
    return sum(x << (32 * j) for j, x in enumerate(w))


def is_tight(w):
    return all(0 <= x <= M32 for x in w) and w[7] <= W7_MAX


def shfl(vals, src):
    """lane j reads vals[src[j]] (within the octet)"""
    return [vals[src[j]] for j in range(8)]


def ripple(lo, c):
    """lo[j] < 2^32 words, c[j] in {0,1} = carry out of lane j into lane j+1 (c[7] == 0); lanes whose word is all ones
    propagate.  Resolved with two ballots and an 8-bit addition, exactly as on the device."""
    assert c[7] == 0 and all(x in (0, 1) for x in c)
    g = sum(c[j] << j for j in range(8))          # ballot(c)
    p = sum((1 << j) for j in range(8) if lo[j] == M32)  # ballot(lo == 0xffffffff)
    assert g & p == 0                              # a lane that overflowed holds a small word
    # Adding X = p | g and Y = g as 8-bit integers: at bit j the carry-out is (x_j & y_j) | ((x_j ^ y_j) & carry_in_j)
    # = g_j | (p_j & carry_in_j) — the carry lane j hands to lane j + 1.  So the carry INTO bit j of X + Y is the carry
    # into lane j, and (X + Y) ^ X ^ Y reads it off for all lanes at once.
    xx, yy = p | g, g
    cin = ((xx + yy) ^ xx ^ yy) & 0xFF
    out = [(lo[j] + ((cin >> j) & 1)) & M32 for j in range(8)]
    # lane 7 never overflows: its word is < 2^31 + small
    assert lo[7] + ((cin >> 7) & 1) <= M32
    return out


def normalize(s, max_hi_bits=26):
    """s[j]: lane sums < 2^64 (a lazy linear combination of tight values) -> tight words of the same value mod p.
    Lane 7 keeps 31 bits and its excess wraps to lane 0 times 19 (2^255 = 19); the other lanes pass bits >= 32 up."""
    assert all(0 <= x <= M64 for x in s)
    hi = [(s[j] >> (31 if j == 7 else 32)) for j in range(8)]
    lo = [(s[j] & (0x7FFFFFFF if j == 7 else M32)) for j in range(8)]
    assert all(h < 2**max_hi_bits for h in hi), hi
    r = shfl(hi, [(j - 1) & 7 for j in range(8)])
    v = [lo[j] + r[j] * (19 if j == 0 else 1) for j in range(8)]
    assert all(x < 2**33 for x in v)
    vlo = [x & M32 for x in v]
    c = [x >> 32 for x in v]
    assert c[7] == 0
    out = ripple(vlo, c)
    assert is_tight(out), out
    return out


def mul(a, b):
    """a, b: 32-bit words per lane (any value < 2^256) -> tight product.  Lane j accumulates the folded column
    T_j = sum_i a_i * b_((j-i) mod 8) * (1 if i <= j else 38) with b split into 16-bit halves so that plain 64-bit
    accumulators never overflow, then two shuffle passes and the ripple."""
    assert all(0 <= x <= M32 for x in a + b)
    acc0, acc1 = [0] * 8, [0] * 8
    for i in range(8):
        ai = shfl(a, [i] * 8)
        br = shfl(b, [(j - i) & 7 for j in range(8)])
        for j in range(8):
            f = 1 if i <= j else 38
            blo, bhi = (br[j] & 0xFFFF) * f, (br[j] >> 16) * f
            assert blo <= M32 and bhi <= M32
            acc0[j] += ai[j] * blo
            acc1[j] += ai[j] * bhi
            assert acc0[j] <= M64 and acc1[j] <= M64
    # T = acc0 + acc1 * 2^16 as three words
    t0, t1, t2 = [0] * 8, [0] * 8, [0] * 8
    for j in range(8):
        T = acc0[j] + (acc1[j] << 16)
        t0[j], t1[j], t2[j] = T & M32, (T >> 32) & M32, T >> 64
        assert t2[j] < 2**10
    # pass 1: word j += t1 of lane j-1 and t2 of lane j-2; what leaves the top wraps around times 38 (2^256 = 38)
    r1 = shfl(t1, [(j - 1) & 7 for j in range(8)])
    r2 = shfl(t2, [(j - 2) & 7 for j in range(8)])
    w = [t0[j] + r1[j] * (38 if j == 0 else 1) + r2[j] * (38 if j < 2 else 1) for j in range(8)]
    assert all(x < 2**39 for x in w)
    return normalize(w, max_hi_bits=8)


def add(a, b):
    return normalize([a[j] + b[j] for j in range(8)], 3)


K1 = to_lanes(P - 37)             # a - b       = a + ~b + K1          (mod p):  ~b = 2^256 - 1 - b = -b + 37
K2 = to_lanes(2**256 - 38 - 74)   # a - b - c   = a + ~b + ~c + K2    (mod p):  2p - 74


def sub(a, b):
    return normalize([a[j] + (b[j] ^ M32) + K1[j] for j in range(8)], 3)


def lincomb(plus, minus):
    """sum of the tight values in `plus` minus those in `minus` (at most two negated), one normalisation"""
    assert len(minus) <= 2 and len(plus) <= 4
    k = [0] * 8 if not minus else (K1 if len(minus) == 1 else K2)
    s = [sum(x[j] for x in plus) + sum(x[j] ^ M32 for x in minus) + k[j] for j in range(8)]
    return normalize(s, 4)


def canon_value(w):
    return from_lanes(w) % P


def selftest(rounds=3000, seed=1):
    rng = random.Random(seed)
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2**255, 2**255 + 2**233 - 1, 2**256 - 1, 2**256 - 38, 2**224 - 1,
            2**255 - 1, (2**255 - 1) ^ (2**32 - 1), int("ffffffff" * 7, 16), (W7_MAX << 224) | (2**224 - 1)]

    def rnd():
        t = rng.random()
        if t < 0.2:
            return rng.choice(edge)
        if t < 0.3:  # long runs of all-ones words (ripple paths)
            v = rng.getrandbits(256)
            lo, hi = sorted((rng.randrange(9), rng.randrange(9)))
            for j in range(lo, hi):
                v |= M32 << (32 * j)
            return v
        return rng.getrandbits(256)

    def tight(v):
        w = to_lanes(v)
        if w[7] > W7_MAX:
            w[7] &= 0x7FFFFFFF
        return w

    for _ in range(rounds):
        x, y, z = rnd(), rnd(), rnd()
        a, b = to_lanes(x), to_lanes(y)
        assert canon_value(mul(a, b)) == x * y % P
        ta, tb, tc = tight(x), tight(y), tight(z)
        va, vb, vc = from_lanes(ta), from_lanes(tb), from_lanes(tc)
        assert canon_value(add(ta, tb)) == (va + vb) % P
        assert canon_value(sub(ta, tb)) == (va - vb) % P
        assert canon_value(lincomb([ta, ta, tb], [tc])) == (2 * va + vb - vc) % P
        assert canon_value(lincomb([ta], [tb, tc])) == (va - vb - vc) % P
        assert canon_value(lincomb([ta, tb, tc, ta], [])) == (2 * va + vb + vc) % P
    return True


if __name__ == "__main__":
    selftest()
    print("fe8 model ok")
