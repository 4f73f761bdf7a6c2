"""Generates tests/golden/lstmp_small.npz from the fp32 CPU oracle.

The reference has no golden vectors for this path and cannot be executed here (SURVEY.md
section 8c), so these come from the oracle restatement AFTER it has been pinned against torch
autograd / finite differences (tests/test_oracle.py).  Committed so that (a) the GPU engine is
checked against fixed bytes and (b) a silent change of the oracle is caught
(tests/test_oracle.py::test_oracle_reproduces_golden).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402


def main():
    I, C, R, S, T, nchunks = 8, 16, 8, 3, 4, 3
    lr, mmt = 1e-2, 0.9
    rng = np.random.RandomState(20240924)
    params = oracle_py.init_params(I, C, R, 0.3, 4321)
    x = rng.randn(nchunks, T * S, I).astype(np.float32)
    od = (rng.randn(nchunks, T * S, R) * 0.1).astype(np.float32)
    flags = np.array([[0, 0, 0], [0, 1, 0], [1, 0, 1]], np.int32)
    o = oracle_py.Oracle(I, C, R, S, np.float32)
    o.set_params(params)
    out, in_diff, corr, pafter, state = [], [], [], [], []
    for n in range(nchunks):
        o.reset(flags[n])
        out.append(o.propagate(x[n]))
        in_diff.append(o.backpropagate(x[n], od[n], mmt))
        o.update(lr)
        corr.append(o.get_grads())
        pafter.append(o.get_params())
        state.append(o.get_state())
    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.abspath(__file__)), "lstmp_small.npz"),
        dims=np.array([I, C, R, S, T]), nchunks=nchunks, lr=lr, momentum=mmt, params=params, x=x, out_diff=od,
        flags=flags, out=np.stack(out), in_diff=np.stack(in_diff), corr=np.stack(corr),
        params_after=np.stack(pafter), state=np.stack(state))


if __name__ == "__main__":
    main()
