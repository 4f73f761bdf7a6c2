"""Generates tests/golden/lstmp_small.npz by RUNNING THE REFERENCE ITSELF: oracle/_ref is the reference's unmodified
google/nnet/bd-nnet-lstm-projected-streams.h compiled from /root/reference against a CPU Kaldi surface
(`make -C oracle ref`, oracle/ref_build/).  The reference ships no golden vectors of its own; these
bytes are its outputs on seeded inputs, committed because /root/reference does not exist on the GPU box.  They pin
(a) the restated oracle (tests/test_oracle.py::test_oracle_reproduces_golden) and (b) the CUDA engine
(tests/test_parity_gpu.py::test_golden_fixture).  `source` in the file records what produced it.

    python tests/golden/make_golden.py        # needs /root/reference (or a prebuilt oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py, ref_py  # noqa: E402


def main():
    I, C, R, S, T, nchunks = 8, 16, 8, 3, 4, 3
    lr, mmt = 1e-2, 0.9
    rng = np.random.RandomState(20240924)
    params = oracle_py.init_params(I, C, R, 0.3, 4321)
    x = rng.randn(nchunks, T * S, I).astype(np.float32)
    od = (rng.randn(nchunks, T * S, R) * 0.1).astype(np.float32)
    flags = np.array([[0, 0, 0], [0, 1, 0], [1, 0, 1]], np.int32)
    if not ref_py.build():
        raise SystemExit("oracle/_ref cannot be built here (no /root/reference): the golden file must come from the reference")
    ref_py.use_builtin_gemm()  # plain triple-loop sgemm: no dependence on a BLAS build
    o = ref_py.RefLstm(I, C, R, S)
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
        params_after=np.stack(pafter), state=np.stack(state),
        source="oracle/_ref: " + ref_py.lib().lstmp_ref_sources().decode())


if __name__ == "__main__":
    main()
