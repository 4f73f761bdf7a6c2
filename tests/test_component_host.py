"""CPU tests of the Python component mirror's HOST logic (kaldi-lstm_b200/component.py) with a stand-in engine:
config parsing and error behaviour as the reference (google/nnet/bd-nnet-lstm-projected-streams.h:55-99, 212-225),
re-creation of the engine when a longer chunk arrives (the reference resizes per call, LPS.h:230), deep Copy()."""
import numpy as np
import pytest
import torch

import kaldi_lstm_b200 as klb
from kaldi_lstm_b200 import component


class FakeEngine:
    created = []

    def __init__(self, I, C, R, S, T, device=0):
        self.I, self.C, self.R, self.S, self.Tmax = I, C, R, S, T
        self.num_params = 4 * C * I + 4 * C * R + 4 * C + 3 * C + R * C
        self.flat = {0: np.zeros(self.num_params, np.float32), 1: np.zeros(self.num_params, np.float32)}
        self.state = (np.zeros((S, C), np.float32), np.zeros((S, R), np.float32))
        self.calls = []
        self.closed = False
        FakeEngine.created.append(self)

    def set_flat(self, which, a):
        self.flat[which] = np.asarray(a, np.float32).copy()

    def get_flat(self, which):
        return self.flat[which].copy()

    def get_state(self):
        return self.state[0].copy(), self.state[1].copy()

    def set_state(self, c, r):
        self.state = (np.asarray(c, np.float32).copy(), np.asarray(r, np.float32).copy())

    def reset(self, flags):
        self.calls.append(("reset", list(flags)))

    def propagate(self, x, out):
        assert x.shape[0] // self.S <= self.Tmax
        self.calls.append(("propagate", x.shape[0]))

    def backpropagate(self, x, od, ind):
        self.calls.append(("backpropagate", ind is not None))

    def update(self, lr, mmt):
        self.calls.append(("update", lr, mmt))

    def close(self):
        self.closed = True


@pytest.fixture
def fake(monkeypatch):
    FakeEngine.created = []
    monkeypatch.setattr(component, "Engine", FakeEngine)
    return FakeEngine


def test_init_data_parsing_and_errors(fake):
    c = klb.LstmProjectedStreams(40, 512)
    c.InitData("<CellDim> 800 <NumStream> 4 <ParamScale> 0.01", seed=3)             # google/nnet.proto:3
    assert (c.ncell_, c.nstream_, c.nrecur_, c.InputDim(), c.OutputDim()) == (800, 4, 512, 40, 512)
    assert c.NumParams() == 2181600 and c.IsUpdatable() and c.GetType() == "kLstmProjectedStreams"
    p = c.GetParams()
    assert p.dtype == np.float32 and np.abs(p).max() <= 0.01 and np.abs(p).max() > 0.009          # U(-scale, scale), LPS.h:41-53
    with pytest.raises(RuntimeError) as ei:                                          # KALDI_ERR, LPS.h:70
        klb.LstmProjectedStreams(40, 512).InitData("<CellDim> 800 <NumStreams> 4")
    assert "Unknown token <NumStreams>" in str(ei.value)
    assert "w_gifo_x_" in c.Info() and "peephole_o_c_" in c.Info() and "w_r_m__corr_" not in c.Info()
    assert "w_gifo_r__corr_" in c.InfoGradient() or "w_gifo_r_corr" in c.InfoGradient()


def test_shape_assertions_and_call_forwarding(fake):
    c = klb.LstmProjectedStreams(8, 6, max_frames=5)
    c.InitData("<CellDim> 4 <NumStream> 3")
    c.SetTrainOptions(klb.NnetTrainOptions(learn_rate=0.5, momentum=0.25))
    with pytest.raises(AssertionError):       # KALDI_ASSERT(prev_nnet_state_.NumRows() == stream_reset_flag.size()), LPS.h:214
        c.Reset([1, 0])
    with pytest.raises(AssertionError):       # KALDI_ASSERT(in.NumRows() % nstream_ == 0), LPS.h:225
        c.Propagate(torch.zeros(7, 8))
    c.Reset([1, 0, 1])
    x = torch.zeros(15, 8)
    out = c.Propagate(x)
    assert out.shape == (15, 6)
    ind = c.Backpropagate(x, out, torch.zeros(15, 6))
    assert ind.shape == x.shape
    assert c.Backpropagate(x, out, torch.zeros(15, 6), want_in_diff=False) is None
    c.Update()
    assert c.engine.calls == [("reset", [1, 0, 1]), ("propagate", 15), ("backpropagate", True), ("backpropagate", False),
                              ("update", 0.5, 0.25)]


def test_longer_chunk_recreates_the_engine_and_keeps_everything(fake):
    c = klb.LstmProjectedStreams(8, 6, max_frames=4)
    c.InitData("<CellDim> 4 <NumStream> 2", seed=1)
    first = c.engine
    first.set_flat(1, np.arange(first.num_params, dtype=np.float32))
    first.set_state(np.full((2, 4), 3.0), np.full((2, 6), 4.0))
    params = c.GetParams()
    c.Propagate(torch.zeros(2 * 9, 8))        # T = 9 > max_frames = 4
    assert c.engine is not first and first.closed and c.engine.Tmax == 9
    np.testing.assert_array_equal(c.GetParams(), params)
    np.testing.assert_array_equal(c.GetGradients(), np.arange(first.num_params, dtype=np.float32))
    assert c.engine.get_state()[0][0, 0] == 3.0 and c.engine.get_state()[1][0, 0] == 4.0


def test_copy_is_deep(fake):
    c = klb.LstmProjectedStreams(8, 6)
    c.InitData("<CellDim> 4 <NumStream> 2 <ParamScale> 0.3", seed=2)
    c.SetTrainOptions(klb.NnetTrainOptions(0.1, 0.9))
    c.engine.set_state(np.ones((2, 4)), np.ones((2, 6)))
    twin = c.Copy()
    assert twin.engine is not c.engine and twin.GetTrainOptions() is not c.GetTrainOptions()
    np.testing.assert_array_equal(twin.GetParams(), c.GetParams())
    assert twin.engine.get_state()[0].sum() == 8.0 and twin.GetTrainOptions().momentum == 0.9
    twin.SetParams(np.zeros(twin.NumParams(), np.float32))
    assert np.abs(c.GetParams()).max() > 0


def test_time_shift_mirror_config_parsing():
    """TimeShift::InitData (standard/nnet/nnet-time-shift.h:21-30): <Shift> n, anything else is a KALDI_ERR."""
    import pytest
    import kaldi_lstm_b200 as klb
    ts = klb.TimeShift(40)
    ts.InitData("<Shift> 5")
    assert ts.shift_ == 5 and ts.GetType() == "TimeShift" and ts.input_dim_ == ts.output_dim_ == 40
    ts.InitData("<Shift> -3")
    assert ts.shift_ == -3
    with pytest.raises(RuntimeError):
        ts.InitData("<Shfit> 3")
    assert ts.BackpropagateFnc(None, None, None, None) is None   # meaningless in the reference too (:53-56)
