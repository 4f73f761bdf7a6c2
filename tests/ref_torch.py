"""Independent torch restatement of the LSTMP-streams equations, used ONLY to pin the C
oracle (tests/test_oracle.py, tests/golden/make_golden.py).

Written from the equations (misc/LSTM_DIAG_EQUATION.jpg eqs 1-5,7 with diagonal peepholes;
google/nnet/bd-nnet-lstm-projected-streams.h:261-325), not from the oracle's C code: the
backward pass comes from torch.autograd, so it checks the reference's hand-derived BPTT
(LPS.h:369-454) rather than repeating it.  Two reference quirks are modelled explicitly:

* the cell clamp to +-50 (LPS.h:296-297) is NOT differentiated by the reference's backward
  (no mask on d_c, LPS.h:424-428)  -> straight-through clamp;
* BPTT is truncated at the chunk boundary: the carried state (c_0, r_0) is a constant
  (backpropagate_buf_ row-block T+1 / history stay zero, LPS.h:351-352).
"""
import torch


class _ClampSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lo, hi):
        return x.clamp(lo, hi)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def unpack(flat, I, C, R):
    off = 0

    def take(n, shape):
        nonlocal off
        t = flat[off:off + n].reshape(shape)
        off += n
        return t

    w_x = take(4 * C * I, (4 * C, I))
    w_r = take(4 * C * R, (4 * C, R))
    b = take(4 * C, (4 * C,))
    p_i = take(C, (C,))
    p_f = take(C, (C,))
    p_o = take(C, (C,))
    w_m = take(R * C, (R, C))
    return w_x, w_r, b, p_i, p_f, p_o, w_m


def forward(flat, x, c0, r0, S, I, C, R):
    """x: (T*S, I) time-major rows t*S+s.  Returns out (T*S, R), c_T, r_T and a dict of per-step
    activations (each (T, S, C) / (T, S, R))."""
    w_x, w_r, b, p_i, p_f, p_o, w_m = unpack(flat, I, C, R)
    T = x.shape[0] // S
    xs = x.reshape(T, S, I)
    c, r = c0, r0
    outs, acts = [], {k: [] for k in "gifochmr"}
    for t in range(T):
        pre = xs[t] @ w_x.t() + b + r @ w_r.t()
        pg, pi, pf, po = pre[:, :C], pre[:, C:2 * C], pre[:, 2 * C:3 * C], pre[:, 3 * C:]
        i = torch.sigmoid(pi + c * p_i)
        f = torch.sigmoid(pf + c * p_f)
        g = torch.tanh(pg)
        c = _ClampSTE.apply(g * i + c * f, -50.0, 50.0)
        h = torch.tanh(c)
        o = torch.sigmoid(po + c * p_o)
        m = h * o
        r = m @ w_m.t()
        outs.append(r)
        for k, v in zip("gifochmr", (g, i, f, o, c, h, m, r)):
            acts[k].append(v)
    out = torch.stack(outs).reshape(T * S, R)
    return out, c, r, {k: torch.stack(v) for k, v in acts.items()}


def fwd_bwd(flat_np, x_np, out_diff_np, c0_np, r0_np, S, I, C, R, dtype=torch.float64):
    """Returns out, in_diff, grad(flat) (plain sums over all T*S rows), c_T, r_T as numpy."""
    flat = torch.tensor(flat_np, dtype=dtype, requires_grad=True)
    x = torch.tensor(x_np, dtype=dtype, requires_grad=True)
    c0 = torch.tensor(c0_np, dtype=dtype)
    r0 = torch.tensor(r0_np, dtype=dtype)
    out, cT, rT, _ = forward(flat, x, c0, r0, S, I, C, R)
    od = torch.tensor(out_diff_np, dtype=dtype)
    (out * od).sum().backward()
    return (out.detach().numpy(), x.grad.numpy(), flat.grad.numpy(), cT.detach().numpy(), rT.detach().numpy())
