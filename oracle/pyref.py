"""oracle/pyref.py — TEST INFRASTRUCTURE.  Independent Python big-int model used to pin
oracle/ref_corrected.c.  Nothing in cudabulletproof_b200/ imports this.

Constants follow RFC 8032 §5.1 and SURVEY.md Appendix B; formulas are textbook affine Edwards
arithmetic (deliberately a different algorithm from the extended-coordinate code under test)."""
import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
D2 = (2 * D) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
BY = (4 * pow(5, P - 2, P)) % P


def fe_inv(x):
    return pow(x, P - 2, P)


def recover_x(y, sign):
    if y >= P:
        return None
    u = (y * y - 1) % P
    v = (D * y * y + 1) % P
    x = (u * pow(v, 3, P) * pow(u * pow(v, 7, P) % P, (P - 5) // 8, P)) % P
    if (v * x * x - u) % P != 0:
        if (v * x * x + u) % P != 0:
            return None
        x = (x * SQRT_M1) % P
    if x == 0 and sign:
        return None
    if (x & 1) != sign:
        x = P - x
    return x


BX = recover_x(BY, 0)
B = (BX, BY)
IDENT = (0, 1)


def pt_add(p, q):
    """affine twisted Edwards addition, a = -1 (complete on this curve)"""
    x1, y1 = p
    x2, y2 = q
    k = D * x1 * x2 * y1 * y2 % P
    x3 = (x1 * y2 + y1 * x2) * fe_inv(1 + k) % P
    y3 = (y1 * y2 + x1 * x2) * fe_inv(1 - k) % P
    return (x3, y3)


def pt_neg(p):
    return ((-p[0]) % P, p[1])


def pt_mul(k, p):
    r = IDENT
    while k:
        if k & 1:
            r = pt_add(r, p)
        p = pt_add(p, p)
        k >>= 1
    return r


def on_curve(p):
    x, y = p
    return (-x * x + y * y - 1 - D * x * x * y * y) % P == 0


def encode(p):
    x, y = p
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def decode(b):
    v = int.from_bytes(b, "little")
    sign = v >> 255
    y = v & ((1 << 255) - 1)
    x = recover_x(y, sign)
    return None if x is None else (x, y)


def msm(scalars, points):
    r = IDENT
    for k, p in zip(scalars, points):
        r = pt_add(r, pt_mul(k, p))
    return r


def rfc8032_public_key(secret: bytes) -> bytes:
    h = hashlib.sha512(secret).digest()
    a = bytearray(h[:32])
    a[0] &= 248
    a[31] &= 127
    a[31] |= 64
    return encode(pt_mul(int.from_bytes(a, "little"), B))


def rfc8032_scalar(secret: bytes) -> int:
    h = hashlib.sha512(secret).digest()
    a = bytearray(h[:32])
    a[0] &= 248
    a[31] &= 127
    a[31] |= 64
    return int.from_bytes(a, "little")


# RFC 8032 §7.1 test vectors (secret key, public key)
RFC8032_VECTORS = [
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60",
     "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb",
     "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7",
     "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025"),
    ("f5e5767cf153319517630f226876b86c8160cc583bc013744c6bf255f5cc0ee5",
     "278117fc144c72340f67d0f2316e8386ceffbf2b2428c9c51fef7c597f1d426e"),
    ("833fe62409237b9d62ec77587520911e9a759cec1d19755b7da901b96dca3d42",
     "ec172b93ad5e563bf4932c70e1245034c35467ef2efd4d64ebf819683467e2bf"),
]
