"""Shared helpers for the GPU parity tests: run the CUDA engine (through the C ABI via the
host-side mirror) and the CPU oracle on identical seeded inputs and compare.

Tolerance (BASELINE.json north_star): outputs within 1e-4 relative, fp32.  Measured both as
max|a-b| / max|b| and Kaldi-style ||a-b||_F <= tol * ||b||_F (cu-matrix.h:142-143 ApproxEqual)."""
import numpy as np

TOL = 1e-4


def rel_max(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def rel_fro(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def assert_close(a, b, what, tol=TOL):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.isfinite(a).all(), "%s: non-finite values from the engine" % what
    rm, rf = rel_max(a, b), rel_fro(a, b)
    assert rm <= tol and rf <= tol, "%s: rel_max=%.3e rel_fro=%.3e (tol %.0e)" % (what, rm, rf, tol)
    return rm


def make_inputs(I, R, S, T, nchunks, seed, od_scale=0.1):
    rng = np.random.RandomState(seed)
    xs = [rng.randn(T * S, I).astype(np.float32) for _ in range(nchunks)]
    ods = [(rng.randn(T * S, R) * od_scale).astype(np.float32) for _ in range(nchunks)]
    return xs, ods


def strided_cuda(arr, pad):
    """Device copy of arr inside a wider allocation, so stride(0) = cols + pad (a pitched CuMatrix)."""
    import torch
    rows, cols = arr.shape
    buf = torch.full((rows, cols + pad), float("nan"), dtype=torch.float32, device="cuda")
    view = buf[:, :cols]
    view.copy_(torch.from_numpy(arr))
    return view


def run_pair(klb, oracle_py, I, C, R, S, T, nchunks=1, momentum=0.9, lr=1e-3, scale=0.1, seed=0, resets=None,
             pad=0, init_state=False, want_in_diff=True, check_record=False, od_scale=0.1, flat=None, Tmax=None):
    """Runs nchunks of propagate/backpropagate/update on both sides; asserts parity after every call.
    resets: optional list (per chunk) of flag vectors applied before the chunk.  Returns max error seen."""
    import torch
    if flat is None:
        flat = oracle_py.init_params(I, C, R, scale, seed + 100)
    xs, ods = make_inputs(I, R, S, T, nchunks, seed, od_scale)
    comp = klb.LstmProjectedStreams(I, R, max_frames=Tmax or T)
    comp.InitData("<CellDim> %d <NumStream> %d <ParamScale> %g" % (C, S, scale))
    comp.SetParams(flat)
    comp.SetTrainOptions(klb.NnetTrainOptions(learn_rate=lr, momentum=momentum))
    o = oracle_py.Oracle(I, C, R, S, np.float32)
    o.set_params(flat)
    if init_state:
        rng = np.random.RandomState(seed + 7)
        st = np.zeros((S, 7 * C + R), np.float32)
        st[:, 4 * C:5 * C] = rng.randn(S, C) * 0.5
        st[:, 7 * C:] = rng.randn(S, R) * 0.5
        o.set_state(st)
        comp.engine.set_state(st[:, 4 * C:5 * C], st[:, 7 * C:])
    worst = 0.0
    for n in range(nchunks):
        if resets is not None and resets[n] is not None:
            comp.Reset(list(resets[n]))
            o.reset(resets[n])
        x, od = xs[n], ods[n]
        xd = strided_cuda(x, pad) if pad else torch.from_numpy(x).cuda()
        odd = strided_cuda(od, pad) if pad else torch.from_numpy(od).cuda()
        if pad:
            outbuf = torch.zeros((T * S, R + pad), dtype=torch.float32, device="cuda")
            out = outbuf[:, :R]
            comp.PropagateFnc(xd, out)
        else:
            out = comp.Propagate(xd)
        ref_out = o.propagate(x)
        worst = max(worst, assert_close(out.cpu().numpy(), ref_out, "chunk %d out" % n))
        if check_record:
            rec = comp.engine.get_record(False)
            ref_rec = o.prop_buf()[S:(T + 1) * S]
            for k, name in enumerate("gifochm"):
                assert_close(rec[:, k * C:(k + 1) * C], ref_rec[:, k * C:(k + 1) * C], "chunk %d act %s" % (n, name))
            assert_close(rec[:, 7 * C:], ref_rec[:, 7 * C:], "chunk %d act r" % n)
        if want_in_diff:
            if pad:
                idbuf = torch.zeros((T * S, I + pad), dtype=torch.float32, device="cuda")
                in_diff = idbuf[:, :I]
                comp.BackpropagateFnc(xd, out, odd, in_diff)
            else:
                in_diff = comp.Backpropagate(xd, out, odd)
        else:
            comp.BackpropagateFnc(xd, out, odd, None)
            in_diff = None
        ref_in_diff = o.backpropagate(x, od, momentum, want_in_diff=True)
        if check_record:
            rec = comp.engine.get_record(True)
            ref_rec = o.bprop_buf()[S:(T + 1) * S]
            for k, name in enumerate("gifo"):
                assert_close(rec[:, k * C:(k + 1) * C], ref_rec[:, k * C:(k + 1) * C], "chunk %d diff %s" % (n, name))
            assert_close(rec[:, 7 * C:], ref_rec[:, 7 * C:], "chunk %d diff r" % n)
        if in_diff is not None:
            worst = max(worst, assert_close(in_diff.cpu().numpy(), ref_in_diff, "chunk %d in_diff" % n))
        comp.Update()
        o.update(lr)
        sl = oracle_py.param_slices(I, C, R)
        g, rg = comp.GetGradients(), o.get_grads()
        pp, rp = comp.GetParams(), o.get_params()
        for name, (a, b, _) in sl.items():
            worst = max(worst, assert_close(g[a:b], rg[a:b], "chunk %d %s_corr" % (n, name)))
            assert_close(pp[a:b], rp[a:b], "chunk %d %s" % (n, name))
        c, r = comp.engine.get_state()
        st = o.get_state()
        assert_close(c, st[:, 4 * C:5 * C], "chunk %d state c" % n)
        assert_close(r, st[:, 7 * C:], "chunk %d state r" % n)
    return worst, comp, o
