"""Generates tests/golden/xent_small.npz from the CPU oracle (oracle/xent_oracle.py): the committed golden
vectors for the masked cross-entropy (the reference ships none).  Run from the repo root:
    python tests/golden/make_golden_xent.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import xent_oracle  # noqa: E402


def main():
    rows, num_pdf = 24, 37
    o = xent_oracle.XentOracle()
    out = {}
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
                    "stats%d" % n: np.array([o.loss, o.entropy, o.correct, o.frames], np.float64)})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "xent_small.npz"), **out)
    print("wrote xent_small.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
