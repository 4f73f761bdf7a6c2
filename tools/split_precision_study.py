"""Numerical study (CPU, numpy): how accurate are the time-loop products of LstmProjectedStreams when the fp32
operands are split into 2-3 low-precision pieces for the tensor cores?  Emulates the operand splitting exactly
(round-to-nearest bf16 / truncated tf32), accumulates in float64 (the hardware accumulates in fp32, whose own error
is ~1e-7 relative and common to every variant) and compares out / in_diff / gradients with an unsplit float64 run,
using the metric of tests/parity_util.py (max|a-b| / max|b|).

    python tools/split_precision_study.py
"""
import sys

import numpy as np


def bf16_rn(x):
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def tf32_trunc(x):
    x = np.asarray(x, np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x, kind, parts):
    x = np.asarray(x, np.float32)
    out, rem = [], x.copy()
    for _ in range(parts):
        p = bf16_rn(rem) if kind == "bf16" else tf32_trunc(rem)
        out.append(p.astype(np.float64))
        rem = (rem - p).astype(np.float32)
    return out


def make_mm(scheme):
    """returns mm(a, w) = a @ w.T under the operand-splitting scheme."""
    if scheme == "exact":
        return lambda a, w: np.asarray(a, np.float64) @ np.asarray(w, np.float64).T
    kind, pa, pw, terms = scheme

    def mm(a, w):
        A, W = split(a, kind, pa), split(w, kind, pw)
        acc = 0.0
        for i in range(pa):
            for j in range(pw):
                if terms is None or (i, j) in terms:
                    acc = acc + A[i] @ W[j].T
        return acc
    return mm


def sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def run(I, C, R, S, T, scale, seed, mm, nchunks=2):
    rng = np.random.RandomState(seed)
    n = 4 * C * I + 4 * C * R + 4 * C + 3 * C + R * C
    flat = ((rng.random_sample(n) - 0.5) * 2 * scale).astype(np.float32).astype(np.float64)
    o = 0
    def take(k, shape):
        nonlocal o
        v = flat[o:o + k].reshape(shape)
        o += k
        return v
    Wx, Wr, b = take(4 * C * I, (4 * C, I)), take(4 * C * R, (4 * C, R)), take(4 * C, (4 * C,))
    pi, pf, po, Wm = take(C, (C,)), take(C, (C,)), take(C, (C,)), take(R * C, (R, C))
    c0, r0 = np.zeros((S, C)), np.zeros((S, R))
    res = []
    for ch in range(nchunks):
        x = rng.randn(T, S, I).astype(np.float32).astype(np.float64)
        od = (rng.randn(T, S, R) * 0.1).astype(np.float32).astype(np.float64)
        pre = x @ Wx.T + b
        g = np.zeros((T, S, C)); i_ = g.copy(); f = g.copy(); og = g.copy(); c = np.zeros((T + 1, S, C)); h = g.copy(); m = g.copy()
        r = np.zeros((T + 1, S, R)); c[0] = c0; r[0] = r0
        for t in range(T):
            # activations cross the chip as fp32 (what the engine stores), so round them
            a = pre[t] + mm(r[t].astype(np.float32), Wr)
            g[t] = np.tanh(a[:, :C]); i_[t] = sig(a[:, C:2 * C] + pi * c[t]); f[t] = sig(a[:, 2 * C:3 * C] + pf * c[t])
            c[t + 1] = np.clip(f[t] * c[t] + i_[t] * g[t], -50, 50)
            og[t] = sig(a[:, 3 * C:] + po * c[t + 1]); h[t] = np.tanh(c[t + 1]); m[t] = og[t] * h[t]
            r[t + 1] = mm(m[t].astype(np.float32), Wm)
        out = r[1:].copy()
        dG = np.zeros((T, S, 4 * C)); dr = np.zeros((T, S, R)); dc_next = np.zeros((S, C))
        dgn = np.zeros((S, 4 * C))
        for t in range(T - 1, -1, -1):
            dr[t] = od[t] + (mm(dgn.astype(np.float32), Wr.T) if t < T - 1 else 0.0)
            dm = mm(dr[t].astype(np.float32), Wm.T)
            dh = dm * og[t]; do = dm * h[t] * og[t] * (1 - og[t])
            dc = dh * (1 - h[t] ** 2) + do * po + dc_next
            if t < T - 1:
                dc = dc + dgn[:, C:2 * C] * pi + dgn[:, 2 * C:3 * C] * pf  # peephole i,f of t+1 read c(t)
                # (dc_next carries f(t+1)*dc(t+1))
            di = dc * g[t] * i_[t] * (1 - i_[t]); df = dc * c[t] * f[t] * (1 - f[t]); dg = dc * i_[t] * (1 - g[t] ** 2)
            dgn = np.concatenate([dg, di, df, do], 1); dG[t] = dgn
            dc_next = dc * f[t]
        in_diff = dG @ Wx
        gWr = np.einsum("tsk,tsr->kr", dG, r[:-1]); gWm = np.einsum("tsr,tsc->rc", dr, m)
        res.append((out, in_diff, gWr, gWm))
        c0, r0 = c[T], r[T]
    return res


def main():
    cases = [("cfg3 layer1 S=64 T=20 scale .05", (40, 800, 512, 64, 20, 0.05)),
             ("cfg2 S=4 T=20 scale .05", (40, 800, 512, 4, 20, 0.05)),
             ("small 256/128 S=64 scale .08", (40, 256, 128, 64, 6, 0.08)),
             ("small 96/64 S=24 scale .3", (16, 96, 64, 24, 5, 0.3))]
    schemes = {"tf32 hi/lo (4 terms)": ("tf32", 2, 2, None),
               "bf16 x2 / x2 (4 terms)": ("bf16", 2, 2, None),
               "bf16 x3 act / x2 w (6 terms)": ("bf16", 3, 2, None),
               "bf16 x3 / x3 (9 terms)": ("bf16", 3, 3, None),
               "bf16 x3 / x3 minus 3 smallest": ("bf16", 3, 3, {(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)}),
               "single tf32": ("tf32", 1, 1, None)}
    only = sys.argv[1:] or None
    for cname, shp in cases:
        I, C, R, S, T, scale = shp
        ref = run(I, C, R, S, T, scale, 6, make_mm("exact"))
        print("== %s" % cname)
        for sname, sch in schemes.items():
            if only and not any(o in sname for o in only):
                continue
            got = run(I, C, R, S, T, scale, 6, make_mm(sch))
            errs = []
            for ch in range(len(ref)):
                errs.append(["%.1e" % (np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(got[ch], ref[ch])])
            print("  %-32s out,in_diff,gWr,gWm chunk0 %s  chunk1 %s" % (sname, errs[0], errs[1]))


if __name__ == "__main__":
    main()
