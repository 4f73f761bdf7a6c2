"""GPU parity: the sm_100a engine, called through the C ABI, against the CPU oracle on the
same seeded inputs (tolerance 1e-4 relative fp32, BASELINE.json north_star)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def klb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import kaldi_lstm_b200 as k
    k.load_library()  # must exist: no fallback
    return k


@pytest.fixture(scope="module")
def oracle_blas(oracle_mod):
    oracle_mod.use_openblas()
    yield oracle_mod
    oracle_mod.use_builtin_gemm()


def test_tiny_with_record(klb, oracle_mod):
    from parity_util import run_pair
    run_pair(klb, oracle_mod, I=8, C=12, R=8, S=3, T=5, nchunks=2, check_record=True, init_state=True, scale=0.3)


def test_single_frame_chunks(klb, oracle_mod):
    from parity_util import run_pair
    run_pair(klb, oracle_mod, I=8, C=16, R=8, S=2, T=1, nchunks=3, check_record=True, scale=0.3)


def test_cfg2_recipe_default(klb, oracle_blas):
    """BASELINE.json configs[1]: 40-in 800-cell 512-proj, NumStream=4, 20-frame BPTT; two chunks with
    state carry-over, a Reset on a subset of streams, momentum 0.9."""
    from parity_util import run_pair
    resets = [None, np.array([0, 1, 0, 1], np.int32), np.array([1, 0, 0, 0], np.int32)]
    run_pair(klb, oracle_blas, I=40, C=800, R=512, S=4, T=20, nchunks=3, momentum=0.9, lr=1e-3, scale=0.05,
             resets=resets, seed=3)


def test_cfg2_paramscale_001(klb, oracle_blas):
    """ParamScale 0.01 and lr 1e-5 exactly as the shipped recipe (google/nnet.proto:3, train_lstm_streams.sh:3-8)."""
    from parity_util import run_pair
    run_pair(klb, oracle_blas, I=40, C=800, R=512, S=4, T=20, nchunks=2, momentum=0.9, lr=1e-5, scale=0.01, seed=4,
             od_scale=0.01)


def test_cfg1_single_stream_100_frames(klb, oracle_blas):
    """BASELINE.json configs[0] shape: S=1, one utterance x 100 frames (the standard/ LstmProjected case,
    standard/nnet/nnet-lstm-projected.h:222-466, is the S=1 special case of the streams component)."""
    from parity_util import run_pair
    run_pair(klb, oracle_blas, I=40, C=800, R=512, S=1, T=100, nchunks=1, momentum=0.0, scale=0.05, seed=5)


def test_cfg3_layer1_s64(klb, oracle_blas):
    from parity_util import run_pair
    run_pair(klb, oracle_blas, I=40, C=800, R=512, S=64, T=20, nchunks=2, momentum=0.9, scale=0.05, seed=6,
             resets=[None, (np.arange(64) % 3 == 0).astype(np.int32)])


def test_cfg3_layer2_s64_input512(klb, oracle_blas):
    from parity_util import run_pair
    run_pair(klb, oracle_blas, I=512, C=800, R=512, S=64, T=20, nchunks=1, momentum=0.9, scale=0.03, seed=7)


def test_cfg4_per_gpu_shape_s32(klb, oracle_blas):
    from parity_util import run_pair
    run_pair(klb, oracle_blas, I=40, C=800, R=512, S=32, T=20, nchunks=1, scale=0.05, seed=8)


def test_pitched_matrices_and_null_in_diff(klb, oracle_mod):
    """in/out/out_diff/in_diff live in Nnet-owned pitched CuMatrix buffers (stride_ > num_cols_,
    cu-matrix.cc:67-73); in_diff may be absent for the first trainable layer."""
    from parity_util import run_pair
    run_pair(klb, oracle_mod, I=40, C=64, R=32, S=4, T=6, nchunks=2, pad=12, scale=0.2, seed=9)
    run_pair(klb, oracle_mod, I=40, C=64, R=32, S=4, T=6, nchunks=1, want_in_diff=False, scale=0.2, seed=10)


def test_cell_clamp_saturation(klb, oracle_mod):
    """|c| driven past 50: forward clamps (LPS.h:296-297), backward ignores the clamp (LPS.h:424-428)."""
    from parity_util import run_pair
    I, C, R, S, T = 8, 16, 8, 2, 8
    flat = oracle_mod.init_params(I, C, R, 0.3, 77)
    a, b, _ = oracle_mod.param_slices(I, C, R)["bias"]
    flat[a:b] = 4.0
    worst, comp, o = run_pair(klb, oracle_mod, I, C, R, S, T, nchunks=1, flat=flat, init_state=True, seed=11,
                              check_record=True)
    # make the state large and run again: cells must hit the clamp on both sides
    st = np.zeros((S, 7 * C + R), np.float32)
    st[:, 4 * C:5 * C] = 49.9
    o.set_state(st)
    comp.engine.set_state(st[:, 4 * C:5 * C], st[:, 7 * C:])
    import torch
    x = np.random.RandomState(1).randn(T * S, I).astype(np.float32)
    out = comp.Propagate(torch.from_numpy(x).cuda())
    ref = o.propagate(x)
    from parity_util import assert_close
    assert_close(out.cpu().numpy(), ref, "saturated out")
    rec = comp.engine.get_record(False)
    assert np.abs(rec[:, 4 * C:5 * C]).max() == 50.0


@pytest.mark.parametrize("ngroups", [1, 2, 4])
def test_stream_group_decompositions(klb, oracle_blas, ngroups, monkeypatch):
    """The same maths under every work decomposition the engine can pick."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_NGROUPS", str(ngroups))
    monkeypatch.setenv("LSTMP_B200_REC", "0")  # the FP32 FFMA kernels under every decomposition
    _, comp, _ = run_pair(klb, oracle_blas, I=40, C=256, R=128, S=64, T=6, nchunks=2, scale=0.08, seed=12 + ngroups)
    assert comp.engine.info()["ngroups"] == ngroups
    assert comp.engine.info()["fwd_tensor_core"] == 0


@pytest.mark.parametrize("shape", [(40, 256, 128, 64, 6), (16, 96, 64, 24, 5), (8, 32, 32, 8, 3), (40, 800, 512, 64, 4),
                                   (40, 800, 512, 32, 4), (8, 16, 8, 3, 4), (40, 800, 512, 5, 3)])
def test_tma_tensor_core_loops(klb, oracle_blas, shape, monkeypatch):
    """The TMA-fed tcgen05 time loops (num_stream <= 64, cell/recur dims multiples of 8), forward and backward, against
    the oracle with both activation records checked, and against the FP32 FFMA kernels on the same inputs."""
    import torch
    from parity_util import run_pair
    I, C, R, S, T = shape
    _, comp, _ = run_pair(klb, oracle_blas, I=I, C=C, R=R, S=S, T=T, nchunks=3, scale=0.08, seed=40 + S,
                          check_record=True, init_state=True,
                          resets=[None, (np.arange(S) % 2 == 0).astype(np.int32), None])
    info = comp.engine.info()
    assert info["fwd_tensor_core"] == 2 and info["bwd_tensor_core"] == 2
    x = torch.randn(T * S, I, device="cuda")
    od = torch.randn(T * S, R, device="cuda") * 0.1
    a = comp.Copy()
    monkeypatch.setenv("LSTMP_B200_REC", "0")
    b = comp.Copy()
    assert a.engine.info()["fwd_tensor_core"] == 2 and b.engine.info()["fwd_tensor_core"] == 0
    assert b.engine.info()["bwd_tensor_core"] == 0
    oa, ob = a.Propagate(x), b.Propagate(x)
    assert (oa - ob).abs().max().item() <= 2e-5 * ob.abs().max().item()
    ra, rb = a.engine.get_record(False), b.engine.get_record(False)
    assert np.abs(ra - rb).max() <= 2e-5 * np.abs(rb).max()
    da, db = a.Backpropagate(x, oa, od), b.Backpropagate(x, ob, od)
    assert (da - db).abs().max().item() <= 2e-5 * db.abs().max().item()
    ga, gb = a.engine.get_flat(2), b.engine.get_flat(2)
    assert np.abs(ga - gb).max() <= 2e-5 * np.abs(gb).max()
    ra, rb = a.engine.get_record(True), b.engine.get_record(True)
    assert np.abs(ra - rb).max() <= 2e-5 * np.abs(rb).max()


@pytest.mark.parametrize("kp", [1, 2, 4, 8])
def test_backward_cluster_sizes(klb, oracle_blas, kp, monkeypatch):
    """The backward kernel's d_r product is K-split over clusters of kp CTAs (partials summed through distributed shared
    memory): same numbers for every cluster size."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_BWD_KP", str(kp))
    _, comp, _ = run_pair(klb, oracle_blas, I=40, C=256, R=128, S=64, T=6, nchunks=2, scale=0.08, seed=70 + kp,
                          check_record=True)
    info = comp.engine.info()
    assert info["bwd_tensor_core"] == 2 and info["bwd_cluster"] == kp and info["bwd_ctas"] % kp == 0


@pytest.mark.parametrize("groups", [1, 2])
def test_tma_stream_groups(klb, oracle_blas, groups, monkeypatch):
    """One or two independent stream groups (own grid barrier, own rows of the hi/lo exchange arrays)."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_TMA_GROUPS", str(groups))
    _, comp, _ = run_pair(klb, oracle_blas, I=40, C=256, R=128, S=64, T=6, nchunks=2, scale=0.08, seed=80 + groups,
                          check_record=True, resets=[None, (np.arange(64) % 5 == 0).astype(np.int32)])
    info = comp.engine.info()
    assert info["fwd_tensor_core"] == 2 and info["ngroups"] == groups and info["streams_per_group"] == 64 // groups


def test_tma_two_slot_ring(klb, oracle_blas, monkeypatch):
    """A 2-slot operand ring (what the tightest shapes get): fewer TMA producer threads than usual, slot reuse on
    every chunk."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_TMA_SLOTS", "2")
    _, comp, _ = run_pair(klb, oracle_blas, I=40, C=800, R=512, S=64, T=3, nchunks=2, scale=0.05, seed=91,
                          check_record=True)
    assert comp.engine.info()["fwd_tensor_core"] == 2


def test_tma_loops_few_ctas(klb, oracle_mod, monkeypatch):
    """Few CTAs with many cells / columns each: multi-round elementwise loops, wide MMA N, one cluster."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_MAX_CTAS", "6")
    _, comp, _ = run_pair(klb, oracle_mod, I=16, C=96, R=40, S=20, T=4, nchunks=2, scale=0.2, seed=23, check_record=True)
    info = comp.engine.info()
    assert info["fwd_tensor_core"] == 2 and info["bwd_ctas"] == 4


def test_few_ctas(klb, oracle_mod, monkeypatch):
    """Multi-pass tile loops: few CTAs with many cells each."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_MAX_CTAS", "5")
    run_pair(klb, oracle_mod, I=16, C=100, R=36, S=20, T=4, nchunks=2, scale=0.2, seed=21, check_record=True)


def test_weights_streamed_mode_forced(klb, oracle_mod, monkeypatch):
    """The per-timestep weights-streamed path (used when slices do not fit on chip) gives the same numbers."""
    from parity_util import run_pair
    monkeypatch.setenv("LSTMP_B200_FORCE_STREAMED", "1")
    _, comp, _ = run_pair(klb, oracle_mod, I=40, C=64, R=32, S=4, T=6, nchunks=3, scale=0.2, seed=31, check_record=True,
                          init_state=True, resets=[None, np.array([0, 1, 0, 0], np.int32), None])
    assert comp.engine.info()["weights_streamed"] == 1


def test_cfg5_shape_2048_1024(klb, oracle_blas):
    """BASELINE.json configs[4] layer shape (40-in, 2048-cell, 1024-proj): 43 MB of weights do not fit in
    shared memory -> weights-streamed mode.  Few streams/frames so the CPU oracle stays fast."""
    from parity_util import run_pair
    _, comp, _ = run_pair(klb, oracle_blas, I=40, C=2048, R=1024, S=8, T=4, nchunks=2, scale=0.03, seed=32)
    assert comp.engine.info()["weights_streamed"] == 1


def test_cfg5_full_shape_s64_t20(klb, oracle_blas):
    """BASELINE.json configs[4] at the per-GPU shape the config names: 40-in, 2048-cell, 1024-proj, NumStream=64, 20-frame
    BPTT (weights-streamed mode), one chunk with momentum against the oracle (about 85 GFLOP on the host)."""
    from parity_util import run_pair
    _, comp, _ = run_pair(klb, oracle_blas, I=40, C=2048, R=1024, S=64, T=20, nchunks=1, scale=0.02, seed=33,
                          resets=[(np.arange(64) % 4 == 1).astype(np.int32)])
    assert comp.engine.info()["weights_streamed"] == 1


def test_cfg3_stacked_two_layers(klb, oracle_blas):
    """BASELINE.json configs[2] as the STACK the benchmark times (bench.py compute()): 40->800/512->800/512, NumStream=64,
    T=20, two chunks with carried state and a Reset on a subset of streams; layer 2's in_diff is layer 1's out_diff.
    Both layers against two chained oracles: top output, bottom in_diff path (layer-1 gradients), parameters after
    Update of both layers."""
    import torch
    from parity_util import assert_close
    S, T, lr, mmt = 64, 20, 1e-3, 0.9
    shapes = [(40, 800, 512), (512, 800, 512)]
    layers, oracles = [], []
    for li, (I, C, R) in enumerate(shapes):
        flat = oracle_blas.init_params(I, C, R, 0.05, 200 + li)
        c = klb.LstmProjectedStreams(I, R, max_frames=T)
        c.InitData("<CellDim> %d <NumStream> %d" % (C, S))
        c.SetParams(flat)
        c.SetTrainOptions(klb.NnetTrainOptions(lr, mmt))
        o = oracle_blas.Oracle(I, C, R, S, np.float32)
        o.set_params(flat)
        layers.append(c)
        oracles.append(o)
    rng = np.random.RandomState(17)
    outs = [torch.empty(S * T, R, device="cuda") for (_, _, R) in shapes]
    in_diff2 = torch.empty(S * T, 512, device="cuda")
    for n in range(2):
        x = rng.randn(S * T, 40).astype(np.float32)
        od = (rng.randn(S * T, 512) * 0.1).astype(np.float32)
        flags = ((np.arange(S) + n) % 3 == 0).astype(np.int32) if n else np.zeros(S, np.int32)
        xd, odd = torch.from_numpy(x).cuda(), torch.from_numpy(od).cuda()
        h = xd
        for li, c in enumerate(layers):           # exactly bench.py compute()
            c.Reset(list(flags))
            c.PropagateFnc(h, outs[li])
            h = outs[li]
        layers[1].BackpropagateFnc(outs[0], outs[1], odd, in_diff2)
        layers[0].BackpropagateFnc(xd, outs[0], in_diff2, None)
        for c in layers:
            c.Update()
        # two chained oracles
        for o in oracles:
            o.reset(flags)
        h1 = oracles[0].propagate(x)
        h2 = oracles[1].propagate(h1)
        d1 = oracles[1].backpropagate(h1, od, mmt, want_in_diff=True)
        oracles[0].backpropagate(x, d1, mmt, want_in_diff=False)
        for o in oracles:
            o.update(lr)
        assert_close(outs[0].cpu().numpy(), h1, "chunk %d layer-1 out" % n)
        assert_close(outs[1].cpu().numpy(), h2, "chunk %d layer-2 out" % n)
        assert_close(in_diff2.cpu().numpy(), d1, "chunk %d layer-2 in_diff" % n)
        for li in range(2):
            assert_close(layers[li].GetGradients(), oracles[li].get_grads(), "chunk %d layer-%d corr" % (n, li + 1))
            assert_close(layers[li].GetParams(), oracles[li].get_params(), "chunk %d layer-%d params" % (n, li + 1))


def test_growing_chunk_length(klb, oracle_mod):
    """The reference resizes its buffers per call (LPS.h:230); the mirror re-creates the engine."""
    from parity_util import run_pair
    run_pair(klb, oracle_mod, I=8, C=16, R=8, S=2, T=9, nchunks=1, Tmax=4, scale=0.3, seed=22)


def test_golden_fixture(klb):
    """Committed golden vectors (tests/golden/make_golden.py): engine vs stored oracle outputs."""
    import torch
    from parity_util import assert_close
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lstmp_small.npz"))
    I, C, R, S, T = [int(v) for v in g["dims"]]
    comp = klb.LstmProjectedStreams(I, R, max_frames=T)
    comp.InitData("<CellDim> %d <NumStream> %d" % (C, S))
    comp.SetParams(g["params"])
    comp.SetTrainOptions(klb.NnetTrainOptions(float(g["lr"]), float(g["momentum"])))
    for n in range(int(g["nchunks"])):
        comp.Reset(list(g["flags"][n]))
        x = torch.from_numpy(g["x"][n]).cuda()
        out = comp.Propagate(x)
        in_diff = comp.Backpropagate(x, out, torch.from_numpy(g["out_diff"][n]).cuda())
        comp.Update()
        assert_close(out.cpu().numpy(), g["out"][n], "golden out %d" % n)
        assert_close(in_diff.cpu().numpy(), g["in_diff"][n], "golden in_diff %d" % n)
        assert_close(comp.GetGradients(), g["corr"][n], "golden corr %d" % n)
        assert_close(comp.GetParams(), g["params_after"][n], "golden params %d" % n)


def test_error_behaviour(klb):
    import torch
    comp = klb.LstmProjectedStreams(8, 8)
    with pytest.raises(RuntimeError):  # KALDI_ERR on unknown token (LPS.h:70)
        comp.InitData("<CellDim> 16 <NumStreams> 2")
    comp.InitData("<CellDim> 16 <NumStream> 3")
    with pytest.raises(AssertionError):  # LPS.h:225
        comp.Propagate(torch.zeros((4, 8), device="cuda"))
    with pytest.raises(AssertionError):  # LPS.h:214
        comp.Reset([0, 1])
    with pytest.raises(klb.EngineError) as ei:
        comp.engine.backpropagate(torch.zeros((3, 8), device="cuda"), torch.zeros((3, 8), device="cuda"))
    assert ei.value.code == -4  # ESTATE: no propagate yet
    with pytest.raises(klb.EngineError):
        klb.Engine(10, 16, 8, 2, 4)  # input_dim % 4 != 0
    with pytest.raises(klb.EngineError):
        comp.engine.propagate(torch.zeros((3, 8)), torch.zeros((3, 8), device="cuda"))  # host tensor


def test_copy_is_deep(klb, oracle_mod):
    import torch
    comp = klb.LstmProjectedStreams(8, 8)
    comp.InitData("<CellDim> 16 <NumStream> 2 <ParamScale> 0.3", seed=5)
    x = torch.randn(6, 8, device="cuda")
    comp.Propagate(x)
    twin = comp.Copy()
    a = comp.Propagate(x).cpu().numpy()
    b = twin.Propagate(x).cpu().numpy()
    np.testing.assert_array_equal(a, b)  # same params AND same carried state
    twin.SetParams(np.zeros(twin.NumParams(), np.float32))
    assert np.abs(comp.GetParams()).max() > 0


# ---- size-independent properties at BASELINE.json's full shapes -----------------------------
def _full(klb, S=64, I=40, seed=0):
    import torch
    comp = klb.LstmProjectedStreams(I, 512)
    comp.InitData("<CellDim> 800 <NumStream> %d <ParamScale> 0.05" % S, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(20 * S, I, device="cuda", generator=g)
    od = torch.randn(20 * S, 512, device="cuda", generator=g) * 0.1
    return comp, x, od


def test_full_size_determinism_and_linearity(klb):
    """Run-to-run bit determinism (no atomics anywhere) and linearity of the backward pass in out_diff."""
    import torch
    comp, x, od = _full(klb)
    twin = comp.Copy()
    out1 = comp.Propagate(x)
    d1 = comp.Backpropagate(x, out1, od)
    g1 = comp.engine.get_flat(2)
    out2 = twin.Propagate(x)
    d2 = twin.Backpropagate(x, out2, od * 2.0)
    g2 = twin.engine.get_flat(2)
    assert torch.equal(out1, out2)
    assert np.abs(g2 - 2.0 * g1).max() <= 2e-5 * np.abs(g2).max()
    assert (d2 - 2.0 * d1).abs().max().item() <= 2e-5 * d2.abs().max().item()
    # bitwise repeatability
    third = comp.Copy()
    third.engine.set_state(*twin.engine.get_state())
    comp2, _, _ = _full(klb)
    o_a = comp2.Propagate(x)
    da = comp2.Backpropagate(x, o_a, od)
    assert torch.equal(o_a, out1) and torch.equal(da, d1)
    np.testing.assert_array_equal(comp2.engine.get_flat(2), g1)


def test_full_size_zero_out_diff_and_reset(klb):
    import torch
    comp, x, od = _full(klb, S=4)
    out = comp.Propagate(x)
    d = comp.Backpropagate(x, out, torch.zeros_like(od))
    assert d.abs().max().item() == 0.0
    assert np.abs(comp.engine.get_flat(2)).max() == 0.0
    # Reset(all ones) == fresh component: second chunk equals the first
    comp.Reset([1] * 4)
    out2 = comp.Propagate(x)
    assert torch.equal(out, out2)
    # without reset the carried state changes the result
    out3 = comp.Propagate(x)
    assert not torch.equal(out3, out2)
    # zero momentum, one update: params move by exactly -lr * G
    comp.SetTrainOptions(klb.NnetTrainOptions(0.5, 0.0))
    before = comp.GetParams()
    comp.Backpropagate(x, out3, od)
    G = comp.engine.get_flat(2)
    comp.Update()
    np.testing.assert_allclose(comp.GetParams(), before - 0.5 * G, rtol=0, atol=1e-6)
