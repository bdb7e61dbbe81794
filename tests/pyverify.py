"""An independent range-proof verifier in plain Python integers — TEST INFRASTRUCTURE.

Written from the protocol (Bunz et al., Bulletproofs, section 4.2 eqs. (65)-(67) and Protocol 2) and the reference's
Fiat-Shamir byte layouts (bulletproof_challenge.cu:6-77, bulletproof_vectors.cu:448-465), NOT from
oracle/ref_corrected.c or the CUDA code: its own projective Edwards addition law on Python integers (checked against
the affine law of oracle/pyref.py at import), generators folded literally round by round, every equation checked as
written.  It gives the
protocol-level oracle an anchor outside the C restatement: the C oracle's prover, the device prover and the two
verifiers must all agree with it (tests/test_protocol_independent.py)."""
import hashlib

from oracle import pyref

P, L = pyref.P, pyref.L


class _Pt:
    """projective twisted Edwards point (X : Y : Z), a = -1; add-2008-bbjlp, complete on this curve"""
    __slots__ = ("X", "Y", "Z")

    def __init__(self, x, y, z=1):
        self.X, self.Y, self.Z = x % P, y % P, z % P

    def __add__(self, o):
        A = self.Z * o.Z % P
        B = A * A % P
        C = self.X * o.X % P
        D = self.Y * o.Y % P
        E = pyref.D * C * D % P
        F, G = (B - E) % P, (B + E) % P
        X3 = A * F * ((self.X + self.Y) * (o.X + o.Y) - C - D) % P
        Y3 = A * G * (D + C) % P  # a = -1: D - a C
        return _Pt(X3, Y3, F * G % P)

    def __rmul__(self, k):
        r, q = _Pt(0, 1), self
        while k:
            if k & 1:
                r = r + q
            q = q + q
            k >>= 1
        return r

    def __eq__(self, o):
        return (self.X * o.Z - o.X * self.Z) % P == 0 and (self.Y * o.Z - o.Y * self.Z) % P == 0

    def affine(self):
        zi = pow(self.Z, P - 2, P)
        return (self.X * zi % P, self.Y * zi % P)


def _msm(scalars, points):
    r = _Pt(0, 1)
    for k, p in zip(scalars, points):
        r = r + (k % L) * p
    return r


assert (7 * _Pt(*pyref.B) + _Pt(*pyref.pt_mul(5, pyref.B))).affine() == pyref.pt_mul(12, pyref.B)


def _challenge(domain, data):
    h = bytearray(hashlib.sha256(domain + data).digest())
    h[31] &= 0x7F  # bulletproof_challenge.cu:19
    return bytes(h)


def _affine(words):
    """(X, Y, Z, T) as 16 little-endian uint64 -> affine (x, y), or None when off the curve / inconsistent"""
    c = [int(w) for w in words]
    X, Y, Z, T = (sum(c[4 * k + i] << (64 * i) for i in range(4)) % P for k in range(4))
    if Z == 0:
        return None
    zi = pow(Z, P - 2, P)
    pt = (X * zi % P, Y * zi % P)
    if not pyref.on_curve(pt) or (X * Y - Z * T) % P != 0:
        return None
    return pt


def _int(words):
    return sum(int(w) << (64 * i) for i, w in enumerate(words))


def _xy(pt):
    return pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little")


def parse_record(rec, n):
    """flat proof record (include/bpk.h) as uint64 words -> dict"""
    k = n.bit_length() - 1
    w = [int(x) for x in rec]
    pts = [w[16 * i:16 * (i + 1)] for i in range(5)]
    sc = lambda o: _int(w[o:o + 4])  # noqa: E731
    out = dict(V=pts[0], A=pts[1], S=pts[2], T1=pts[3], T2=pts[4], taux=sc(80), mu=sc(84), t=sc(88), a=sc(92), b=sc(96),
               c=sc(100), x=sc(104))
    out["L"] = [w[108 + 16 * j:108 + 16 * (j + 1)] for j in range(k)]
    out["R"] = [w[108 + 16 * (k + j):108 + 16 * (k + j + 1)] for j in range(k)]
    return out


def verify(rec, n, G, H, g, h, V_ext=None):
    """accept / reject of one n-bit range proof; G, H: lists of affine points, g, h: affine points"""
    pr = parse_record(rec, n)
    k = n.bit_length() - 1
    pts = {name: _affine(pr[name]) for name in ("V", "A", "S", "T1", "T2")}
    Ls, Rs = [_affine(p) for p in pr["L"]], [_affine(p) for p in pr["R"]]
    if any(p is None for p in list(pts.values()) + Ls + Rs):
        return False
    if V_ext is not None and _affine(V_ext) != pts["V"]:
        return False
    yb = _challenge(b"BulletproofYChal", _xy(pts["V"]) + _xy(pts["A"]) + _xy(pts["S"]) + b"y_ch")
    zb = _challenge(b"BulletproofZChal", yb + b"z_ch")
    xb = _challenge(b"BulletproofXChal", _xy(pts["T1"]) + _xy(pts["T2"]) + b"xcha")
    y, z, x = (int.from_bytes(b, "little") % L for b in (yb, zb, xb))
    t, taux, mu, a, b = (pr[f] % L for f in ("t", "taux", "mu", "a", "b"))
    if pr["c"] % L != t:
        return False
    # (65): t g + taux h == z^2 V + delta(y, z) g + x T1 + x^2 T2
    Gp, Hq, gp, hp = [_Pt(*q) for q in G], [_Pt(*q) for q in H], _Pt(*g), _Pt(*h)
    Vp, Ap, Sp, T1p, T2p = (_Pt(*pts[name]) for name in ("V", "A", "S", "T1", "T2"))
    delta = ((z - z * z) * sum(pow(y, i, L) for i in range(n)) - pow(z, 3, L) * (2**n - 1)) % L
    if not (_msm([t, taux], [gp, hp]) == _msm([z * z, delta, x, x * x], [Vp, gp, T1p, T2p])):
        return False
    # (66)-(67) with H'_i = y^-i H_i; the argument's extra generator is Q = h, so P also carries t Q - mu h
    yinv = pow(y, -1, L)
    Hp = [pow(yinv, i, L) * Hq[i] for i in range(n)]
    Pt = Ap + x * Sp + _msm([-z] * n, Gp) + _msm([z * pow(y, i, L) + z * z * pow(2, i, L) for i in range(n)], Hp)
    Pt = Pt + ((t - mu) % L) * hp
    # Protocol 2, verifier side: fold G, H' and P with the round challenges
    tr = _challenge(b"BulletproofIP", t.to_bytes(32, "little") + taux.to_bytes(32, "little") + mu.to_bytes(32, "little"))
    Gs, Hs = Gp, Hp
    for j in range(k):
        tr = _challenge(b"InnerProductChal", tr + Ls[j][0].to_bytes(32, "little") + Rs[j][0].to_bytes(32, "little"))
        if j == 0 and pr["x"] % P != int.from_bytes(tr, "little"):  # the proof stores the first raw challenge
            return False
        u = int.from_bytes(tr, "little") % L
        ui = pow(u, -1, L)
        half = len(Gs) // 2
        Gs = [ui * Gs[i] + u * Gs[i + half] for i in range(half)]
        Hs = [u * Hs[i] + ui * Hs[i + half] for i in range(half)]
        Pt = Pt + (u * u % L) * _Pt(*Ls[j]) + (ui * ui % L) * _Pt(*Rs[j])
    return Pt == _msm([a, b, a * b], [Gs[0], Hs[0], hp])
