"""Generates the log2 / exp2 tables and polynomial coefficients of include/rt_math.h (rt_powf).

    python tools/gen_math_tables.py

log2 table: z in [0.70703125, 1.4140625) is cut into 32 intervals by the top 5 mantissa bits after
subtracting the bit pattern 0x3f350000; entry i holds 1/c_i (c_i = interval centre, rounded to binary64)
and log2 c_i := -log2(the rounded 1/c_i), so that r = z * (1/c_i) - 1 and log2 c_i pair exactly.
exp2 table: 2^(i/32).  60-digit decimal arithmetic, results rounded to binary64 and printed as C hex floats.
"""
import math
import struct
from decimal import Decimal, getcontext

getcontext().prec = 60
LN2 = Decimal(2).ln()
N = 32
OFF = 0x3F350000


def u2f(u):
    return struct.unpack("<f", struct.pack("<I", u))[0]


def main():
    print("/* log2 table: { 1/c_i, log2 c_i } */")
    worst = 0.0
    for i in range(N):
        lo, hi = u2f(OFF + (i << 18)), u2f(OFF + ((i + 1) << 18))
        invc = float(1 / ((Decimal(lo) + Decimal(hi)) / 2))
        logc = float(-(Decimal(invc).ln() / LN2))
        worst = max(worst, abs(lo * invc - 1), abs(hi * invc - 1))
        print(f"  {invc.hex()}, {logc.hex()},")
    print(f"/* max |r| = {worst:.6f} */")
    print("/* exp2 table: 2^(i/32) */")
    for i in range(N):
        print(f"  {float(Decimal(2) ** (Decimal(i) / N)).hex()},")
    print("/* log2(1+r) = r * (L1 + L2 r + ... ), L_k = (-1)^(k+1) / (k ln 2) */")
    for k in range(1, 6):
        print(f"  L{k} = {float(Decimal((-1) ** (k + 1)) / (k * LN2)).hex()}")
    print("/* 2^r = 1 + E1 r + E2 r^2 + ..., E_k = ln2^k / k! */")
    for k in range(1, 5):
        print(f"  E{k} = {float(LN2 ** k / Decimal(math.factorial(k))).hex()}")


def asin_fit(degree=7):
    """P(t) with asin(sqrt t) = sqrt t (1 + t P(t)) on [0, 1/4]: least squares on 400 Chebyshev nodes;
    the target comes from the Taylor series in 60-digit decimals."""
    import numpy as np

    def series(t):
        c, s, tk = Decimal(1), Decimal(0), Decimal(1)
        for k in range(1, 400):
            c = c * Decimal((2 * k - 1) ** 2) / Decimal(2 * k * (2 * k + 1))
            term = c * tk
            s += term
            tk *= t
            if term < Decimal(10) ** -50:
                break
        return s

    n = 400
    ts = [Decimal(0.125) * (1 + Decimal(math.cos(math.pi * (i + 0.5) / n))) for i in range(n)]
    fit = np.polynomial.chebyshev.Chebyshev.fit(np.array([float(t) for t in ts]), np.array([float(series(t)) for t in ts]),
                                                degree, domain=[0, 0.25])
    coefs = fit.convert(kind=np.polynomial.Polynomial).coef
    worst = Decimal(0)
    for i in range(1, 2001):
        t, p = Decimal(i) / Decimal(8000), Decimal(0)
        for c in reversed(coefs):
            p = p * t + Decimal(float(c))
        worst = max(worst, abs((1 + t * p) - (1 + t * series(t))) / (1 + t * series(t)))
    print(f"/* asin: P(t), degree {degree}, max relative error {float(worst):.3g}; highest power first */")
    for c in reversed(coefs):
        print(f"  {float(c).hex()},")


if __name__ == "__main__":
    main()
    asin_fit()
