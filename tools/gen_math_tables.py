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


if __name__ == "__main__":
    main()
