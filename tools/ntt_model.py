"""Index-math model of the multi-pass NTT in multilinear_b200/csrc/ntt.cu (design aid, not product code).

N = R_0 * R_1 * ... * R_{P-1}.  Pass p views the array as [A][R][B] (A = product of earlier radices,
B = product of later ones), runs an R-point in-place DIF over the middle index (natural in,
bit-reversed out), un-reverses on store and multiplies by the inter-pass twiddle w_N^(k*b*A).
The final pass (B = 1) writes natural order: k = k_0 + R_0*k_1 + ... .
"""
import sys, os, random
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import pyref as P

M = P.M


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def tile_dif(col, R, wN_pow, N):
    """in-place DIF over list col (len R), natural in -> bit-reversed out; twiddle w_R^e = w_N^(e*N/R)"""
    q, L = 0, R
    while L > 1:
        half = L // 2
        for blk in range(0, R, L):
            for i in range(half):
                u, v = col[blk + i], col[blk + i + half]
                col[blk + i] = (u + v) % M
                col[blk + i + half] = (u - v) * wN_pow((i * (R // L)) * (N // R)) % M
        L //= 2


def ntt_multipass(x, gen, radices, zero_padded=False, inverse=False):
    N = 1
    for r in radices:
        N *= r
    logs = [r.bit_length() - 1 for r in radices]
    if inverse:
        gen = P.inv(gen)
    wN_pow = lambda e: pow(gen, e % N, M)
    data = list(x) + [0] * (N - len(x)) if zero_padded else list(x)
    A = 1
    for p, R in enumerate(radices):
        B = N // (A * R)
        out = [None] * N
        last = p == len(radices) - 1
        for a in range(A):
            for b in range(B):
                col = [data[a * R * B + m * B + b] for m in range(R)]
                tile_dif(col, R, wN_pow, N)
                for pos in range(R):
                    k = bitrev(pos, logs[p])
                    if not last:
                        out[a * R * B + k * B + b] = col[pos] * wN_pow(k * b * A) % M
                    else:
                        # a = k_0*(A/R_0) + a' ; a' = digits k_1..k_{P-2} (k_1 most significant)
                        R0 = radices[0] if len(radices) > 1 else 1
                        k0, ap = divmod(a, A // R0) if len(radices) > 1 else (0, 0)
                        mid, mul, rem_span = 0, 1, A // R0
                        for d in range(1, len(radices) - 1):
                            rem_span //= radices[d]
                            kd, ap = divmod(ap, rem_span)
                            mid += kd * mul
                            mul *= radices[d]
                        out[k0 + R0 * mid + A * k] = col[pos]
        data = out
        A *= R
    if inverse:
        ninv = P.inv(N % M)
        data = [v * ninv % M for v in data]
    return data


if __name__ == "__main__":
    random.seed(3)
    for radices in ([8], [4, 8], [8, 4], [4, 4, 8], [2, 8, 4], [4, 2, 4, 8], [8, 8, 16]):
        N = 1
        for r in radices:
            N *= r
        g = P.pow_2_generator(N.bit_length() - 1)
        x = [random.randrange(M) for _ in range(N)]
        assert ntt_multipass(x, g, radices) == P.ntt(x, g), radices
        assert ntt_multipass(P.ntt(x, g), g, radices, inverse=True) == x, radices
        h = [random.randrange(M) for _ in range(N // 2)]
        assert ntt_multipass(h, g, radices, zero_padded=True) == P.reed_solomon(h, g), radices
        print("ok", radices)
