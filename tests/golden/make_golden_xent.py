"""Generates tests/golden/xent_small.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref: Xent::EvalMasked extracted
from google/nnet/nnet-loss.cc:76-164 at build time, compiled against a CPU Kaldi surface, `make -C oracle ref`).
The reference ships no golden vectors; these are its outputs on seeded inputs.  Run from the repo root:
    python tests/golden/make_golden_xent.py        # needs /root/reference (or a prebuilt oracle/_ref)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py, xent_oracle  # noqa: E402


def main():
    rows, num_pdf = 24, 37
    if not ref_py.build():
        raise SystemExit("oracle/_ref cannot be built here (no /root/reference)")
    o = ref_py.RefXent()
    out = {"source": "oracle/_ref: google/nnet/nnet-loss.cc Xent::EvalMasked"}
    for n, (soft, seed) in enumerate([(False, 1), (True, 2)]):
        mask, y, post = xent_oracle.random_case(rows, num_pdf, seed=seed, soft=soft, empty_every=7, dup_every=5)
        diff = o.eval_masked(mask, y, post)
        rp = np.zeros(rows + 1, np.int32)
        pdf, w = [], []
        for t, lst in enumerate(post):
            rp[t + 1] = rp[t] + len(lst)
            for p_, w_ in lst:
                pdf.append(p_)
                w.append(w_)
        out.update({"mask%d" % n: mask, "y%d" % n: y, "row_ptr%d" % n: rp, "pdf%d" % n: np.array(pdf, np.int32),
                    "weight%d" % n: np.array(w, np.float32), "diff%d" % n: diff,
                    "stats%d" % n: np.array([o.stats()[k] for k in ("loss", "entropy", "correct", "frames")],
                                            np.float64)})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "xent_small.npz"), **out)
    print("wrote xent_small.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
