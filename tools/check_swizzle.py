"""Brute-force bank-conflict check of the x-stage ("column" mapping) shared-memory exchanges.

Model: 32 banks x 4 bytes; a warp-wide access of 16-byte elements is served in 4 phases of 8
threads, of 8-byte elements in 2 phases of 16 threads; a phase is conflict free when its threads
touch distinct 16/8-byte slots modulo 128 bytes (or the same address). Reports the worst number of
wavefronts per phase for every exchange write/read of a length-N plan under a swizzle f(n).
"""
import itertools
import sys


def plan(n):
    log2n = n.bit_length() - 1
    r0 = 8 if log2n % 3 == 0 else (2 if log2n % 3 == 1 else 4)
    stages = (log2n + 2) // 3
    radix = [r0] + [8] * (stages - 1)
    ns = [1]
    for s in range(1, stages):
        ns.append(ns[-1] * radix[s - 1])
    return radix, ns


def worst(n, log2v, elem_bytes, f):
    v = 1 << log2v
    t = n // 8
    radix, ns = plan(n)
    phase = 8 if elem_bytes == 16 else 16
    slots = 128 // elem_bytes
    res = 1
    # thread id -> j = tid % t, lane = tid // t   (column mapping)
    nthreads = t * v
    def addr(nn, lane):
        return (nn << log2v) + (lane ^ (f(nn) & (v - 1)))
    for s in range(len(radix) - 1):
        r, nss = radix[s], ns[s]
        m = 8 // r
        for i in range(m):
            for q in range(r):
                for w0 in range(0, nthreads, phase):
                    seen = {}
                    for tid in range(w0, min(w0 + phase, nthreads)):
                        j, lane = tid % t, tid // t
                        b = j + i * t
                        k = b % nss
                        a = addr((b - k) * r + k + q * nss, lane)
                        seen.setdefault(a % slots, set()).add(a)
                    res = max(res, max(len(x) for x in seen.values()))
        for mm in range(8):
            for w0 in range(0, nthreads, phase):
                seen = {}
                for tid in range(w0, min(w0 + phase, nthreads)):
                    j, lane = tid % t, tid // t
                    a = addr(j + t * mm, lane)
                    seen.setdefault(a % slots, set()).add(a)
                res = max(res, max(len(x) for x in seen.values()))
    return res


if __name__ == "__main__":
    for elem, log2v in ((16, 3), (8, 4)):
        for shifts in ((0,), (0, 3), (0, 3, 6), (0, 3, 6, 9), (0, 4, 8), (0, 2, 4, 6, 8), (0, 3, 4, 6, 7, 8)):
            f = lambda nn, sh=shifts: __import__("functools").reduce(lambda a, b: a ^ b, [(nn >> s) for s in sh])
            out = []
            for n in (16, 32, 64, 128, 256, 512, 1024, 2048):
                out.append(worst(n, log2v, elem, f))
            print(f"elem {elem:2d}B V={1<<log2v:2d} shifts {shifts}: worst wavefronts/phase for N=16..2048: {out}")
