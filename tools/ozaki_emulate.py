#!/usr/bin/env python
"""CPU emulation (numpy, exact integer arithmetic) of the split-operand cross term that
csrc/clike_i8_kernel.cu runs on the tcgen05 INT8 tensor path: FP64 rows and spectra are cut into
S signed 7-bit digits after scaling each row by a power of two; the digit planes are contracted
exactly (int8 x int8 -> int32) and the pairs (s, t) with s + t <= P are recombined in FP64.
Prints the measured error of chi2 = Syy - 2 Sym + Smm against the FP64 direct form on horns data
and on a cancellation fixture, next to the a-priori bound the guard uses.

    python tools/ozaki_emulate.py [--S 7 --P 8]
"""
import argparse
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massivedatans_b200 import synth  # noqa: E402


def digits(a, S):
    """a[rows, C] -> (q[S, rows, C] int8-valued, exponent e[rows]) with a = 2^e sum_s q_s 2^-(7s-1) + tail"""
    amax = numpy.abs(a).max(axis=1)
    e = numpy.where(amax > 0, numpy.ceil(numpy.log2(numpy.where(amax > 0, amax, 1.0))) + 1, 0.0)
    r = a / (2.0 ** e)[:, None]            # |r| <= 0.5
    q = numpy.empty((S,) + a.shape)
    r = r * 64.0
    for s in range(S):
        q[s] = numpy.rint(r)
        r = (r - q[s]) * 128.0
    assert numpy.abs(q).max() <= 64
    return q, e


def cross_term(Y, M, S, P):
    qy, ey = digits(Y, S)
    qm, em = digits(M, S)
    out = numpy.zeros((M.shape[0], Y.shape[0]))
    for p in range(P, 1, -1):                       # smallest weights first
        G = numpy.zeros((M.shape[0], Y.shape[0]), dtype=numpy.int64)
        for s in range(1, S + 1):
            t = p - s
            if 1 <= t <= S:
                G += qm[t - 1].astype(numpy.int64) @ qy[s - 1].astype(numpy.int64).T
        assert numpy.abs(G).max() < 2 ** 31
        out += G.astype(float) * 2.0 ** (-(7 * p - 2))
    return out * (2.0 ** em)[:, None] * (2.0 ** ey)[None, :]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--S', type=int, default=7)
    ap.add_argument('--P', type=int, default=8)
    ap.add_argument('--n', type=int, default=4000)
    args = ap.parse_args()
    S, P = args.S, args.P
    for name in ('horns', 'cancellation'):
        if name == 'horns':
            x, y, _ = synth.horns(args.n, legacy=False, seed=5)
            pts = synth.parameter_points(64, seed=2)
            M = numpy.array([p[0] * numpy.exp(-0.5 * ((p[1] - x) / p[2]) ** 2) for p in pts])
            Y = numpy.ascontiguousarray(y.T)
        else:
            nx = 200
            x = numpy.linspace(400, 800, nx)
            rs = numpy.random.RandomState(5)
            pts = synth.parameter_points(8, seed=3)
            pts[:, 0] = 50.0
            pts[:, 2] = rs.uniform(20, 60, size=8)
            M = numpy.array([p[0] * numpy.exp(-0.5 * ((p[1] - x) / p[2]) ** 2) for p in pts])
            Y = rs.normal(0, 1e-2, size=(args.n, nx))
            fit = numpy.arange(args.n) % 3 == 0
            Y[fit] = M[numpy.arange(args.n)[fit] % 8] + rs.normal(0, 1e-6, size=(fit.sum(), nx))
        C = Y.shape[1]
        direct = ((M[:, None, :] - Y[None, :, :]) ** 2).sum(axis=2)
        Syy = (Y ** 2).sum(axis=1)
        Smm = (M ** 2).sum(axis=1)
        Sym = cross_term(Y, M, S, P)
        chi = Syy[None, :] - 2 * Sym + Smm[:, None]
        rel = numpy.abs(chi - direct) / direct
        scale = (Syy[None, :] + Smm[:, None])
        err_vs_scale = numpy.abs(chi - direct) / scale
        bound = 4.0 * C * (S * 2.0 ** (-7 * (P - 1)) + 2.0 ** (1 - 7 * S))      # relative to Syy + Smm
        for tol in (1e-9, 1e-10):
            guard = bound / tol
            kept = chi >= guard * scale
            print('%-13s S=%d P=%d C=%d tol=%g: guard %.3g  kept %.4f of %d  max rel err of kept %.3g  '
                  '(max |err|/(Syy+Smm) %.3g, bound %.3g)'
                  % (name, S, P, C, tol, guard, kept.mean(), kept.size, rel[kept].max() if kept.any() else 0,
                     err_vs_scale.max(), bound))


if __name__ == '__main__':
    main()
