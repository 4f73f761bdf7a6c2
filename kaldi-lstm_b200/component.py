"""Host-side mirror of the reference component's interface.

Same method names, argument meaning and error behaviour as
``kaldi::nnet1::LstmProjectedStreams`` (google/nnet/bd-nnet-lstm-projected-streams.h), so the
parity tests read like tests of the reference.  All arithmetic happens in the sm_100a engine
behind the C ABI (include/lstmp_b200.h); this class only parses config, owns the handle and
forwards calls.  Matrices are torch CUDA float32 tensors standing in for CuMatrix<BaseFloat>
(row-major, row stride = stride(0), like CuMatrixBase::Stride()).
"""
import numpy as np

from . import nnet_io
from .engine import Engine, EngineError, EINVAL


class NnetTrainOptions:
    """nnet/nnet-trnopts.h (upstream Kaldi): the two fields the component reads (LPS.h:465,502)."""

    def __init__(self, learn_rate=0.008, momentum=0.0):
        self.learn_rate = float(learn_rate)
        self.momentum = float(momentum)


def _moment_statistics(a):
    a = np.asarray(a, np.float64).ravel()
    if a.size == 0:
        return "( empty )"
    mean = a.mean()
    var = a.var()
    std = np.sqrt(var) if var > 0 else 0.0
    skew = ((a - mean) ** 3).mean() / std ** 3 if std > 0 else 0.0
    kurt = ((a - mean) ** 4).mean() / std ** 4 - 3.0 if std > 0 else 0.0
    return "( min %g, max %g, mean %g, variance %g, skewness %g, kurtosis %g )" % (a.min(), a.max(), mean, var, skew, kurt)


class LstmProjectedStreams:
    MARKER = "<LstmProjectedStreams>"  # google/nnet.proto:3

    def __init__(self, input_dim, output_dim, device=0, max_frames=20):
        # LPS.h:27-33
        self.input_dim_ = int(input_dim)
        self.output_dim_ = int(output_dim)
        self.ncell_ = 0
        self.nrecur_ = int(output_dim)
        self.nstream_ = 0
        self.opts_ = NnetTrainOptions()
        self._device = device
        self._max_frames = int(max_frames)
        self._engine = None
        # element-wise clip of the accumulated gradient in Update: None for the streams component, 50.0 for the
        # standard single-stream LstmProjected (standard/nnet/nnet-lstm-projected.h:480-493)
        self.max_grad_ = None

    # ---- Component surface ------------------------------------------------------------
    def GetType(self):
        return "kLstmProjectedStreams"

    def InputDim(self):
        return self.input_dim_

    def OutputDim(self):
        return self.output_dim_

    def IsUpdatable(self):
        return True

    def SetTrainOptions(self, opts):
        self.opts_ = opts

    def GetTrainOptions(self):
        return self.opts_

    def _make_engine(self, max_frames=None):
        if max_frames is not None:
            self._max_frames = int(max_frames)
        self._engine = Engine(self.input_dim_, self.ncell_, self.nrecur_, self.nstream_, self._max_frames, self._device)

    def InitData(self, config, seed=0):
        """Parses ``<CellDim> C <NumStream> S <ParamScale> x`` (LPS.h:55-99); unknown tokens raise
        like KALDI_ERR (LPS.h:70).  Parameters ~ U(-scale, +scale) (LPS.h:41-53), own RNG."""
        param_scale = 0.02
        toks = config.split()
        i = 0
        while i < len(toks):
            t = toks[i]
            if t == "<CellDim>":
                self.ncell_ = int(toks[i + 1])
            elif t == "<NumStream>":
                self.nstream_ = int(toks[i + 1])
            elif t == "<ParamScale>":
                param_scale = float(toks[i + 1])
            else:
                raise RuntimeError("Unknown token %s, a typo in config? (CellDim|NumStream|ParamScale)" % t)
            i += 2
        self._make_engine()
        rng = np.random.RandomState(seed)
        flat = ((rng.random_sample(self._engine.num_params) - 0.5) * 2.0 * param_scale).astype(np.float32)
        self._engine.set_flat(0, flat)

    def Copy(self):
        other = LstmProjectedStreams(self.input_dim_, self.output_dim_, self._device, self._max_frames)
        other.ncell_, other.nstream_ = self.ncell_, self.nstream_
        other.opts_ = NnetTrainOptions(self.opts_.learn_rate, self.opts_.momentum)
        other.max_grad_ = self.max_grad_
        other._make_engine()
        other._engine.set_flat(0, self._engine.get_flat(0))
        other._engine.set_flat(1, self._engine.get_flat(1))
        other._engine.set_state(*self._engine.get_state())
        return other

    # ---- model files (text form of ReadData / WriteData, LPS.h:101-150) ---------------------------------
    def ToNnetComponent(self):
        """This layer as an nnet_io.NnetComponent `<LstmProjectedStreams> out in <CellDim> C <NumStream> S [...]`."""
        return nnet_io.lstm_component_from_flat(self.GetParams(), self.output_dim_, self.input_dim_, self.ncell_,
                                                num_stream=self.nstream_)

    @classmethod
    def FromNnetComponent(cls, comp, device=0, max_frames=20, num_stream=None):
        """Builds the layer from a parsed `<LstmProjectedStreams>` component, or from the standard version's
        `<LstmProjected>` (standard/nnet/nnet-lstm-projected.h:111-123 -- same seven blocks, no <NumStream>), which runs
        as the num_stream = 1 case of the streams engine unless `num_stream` says otherwise."""
        if comp.type not in nnet_io.LSTM_TYPES:
            raise RuntimeError("not an LSTM component: %s" % comp.type)
        c = cls(comp.input_dim, comp.output_dim, device, max_frames)
        c.ncell_ = int(comp.attr("<CellDim>"))
        c.nstream_ = int(num_stream or comp.attr("<NumStream>") or 1)
        if comp.type == "<LstmProjected>":
            c.max_grad_ = 50.0   # the standard version clips in Update (nnet-lstm-projected.h:482)
        c._make_engine()
        c.SetParams(nnet_io.lstm_flat_params(comp))
        return c

    def NumParams(self):
        return self._engine.num_params  # LPS.h:152-160

    def GetParams(self):
        return self._engine.get_flat(0)  # LPS.h:162-189

    def SetParams(self, flat):
        self._engine.set_flat(0, flat)

    def GetGradients(self):
        """The *_corr_ buffers (what InfoGradient reports, LPS.h:201-210), flat."""
        return self._engine.get_flat(1)

    def Info(self):
        return self._info(self.GetParams(), "")

    def InfoGradient(self):
        return self._info(self.GetGradients(), "_corr")

    def _info(self, flat, sfx):
        I, C, R = self.input_dim_, self.ncell_, self.nrecur_
        names = [("w_gifo_x_", 4 * C * I), ("w_gifo_r_", 4 * C * R), ("bias_", 4 * C), ("peephole_i_c_", C),
                 ("peephole_f_c_", C), ("peephole_o_c_", C), ("w_r_m_", R * C)]
        out, off = "    ", 0
        for n, l in names:
            out += "\n  %s%s  %s" % (n, sfx + ("_" if sfx else ""), _moment_statistics(flat[off:off + l]))
            off += l
        return out

    def Reset(self, stream_reset_flag):
        # LPS.h:212-220; KALDI_ASSERT(prev_nnet_state_.NumRows() == stream_reset_flag.size())
        if len(stream_reset_flag) != self.nstream_:
            raise AssertionError("KALDI_ASSERT: prev_nnet_state_.NumRows() == stream_reset_flag.size()")
        self._engine.reset(stream_reset_flag)

    def _ensure_frames(self, rows):
        if rows % self.nstream_ != 0:
            raise AssertionError("KALDI_ASSERT: in.NumRows() % nstream_ == 0")  # LPS.h:225
        T = rows // self.nstream_
        if T > self._max_frames:
            old = self._engine
            params, corr, state = old.get_flat(0), old.get_flat(1), old.get_state()
            old.close()
            self._make_engine(T)
            self._engine.set_flat(0, params)
            self._engine.set_flat(1, corr)
            self._engine.set_state(*state)

    def PropagateFnc(self, in_, out):
        self._ensure_frames(in_.shape[0])
        self._engine.propagate(in_, out)

    def BackpropagateFnc(self, in_, out, out_diff, in_diff):
        # `out` is unused, as in the reference (LPS.h:334-349 reads propagate_buf_).
        self._engine.backpropagate(in_, out_diff, in_diff)

    def Update(self, input_=None, diff=None):
        # both arguments are unused by the reference too (LPS.h:501-512)
        if self.max_grad_:
            self._engine.update_clipped(self.opts_.learn_rate, self.opts_.momentum, self.max_grad_)
        else:
            self._engine.update(self.opts_.learn_rate, self.opts_.momentum)

    # Component::Propagate / Backpropagate (upstream nnet-component.h): size the output, then *Fnc
    def Propagate(self, in_):
        import torch
        out = torch.empty((in_.shape[0], self.output_dim_), dtype=torch.float32, device=in_.device)
        self.PropagateFnc(in_, out)
        return out

    def Backpropagate(self, in_, out, out_diff, want_in_diff=True):
        import torch
        in_diff = torch.empty_like(in_) if want_in_diff else None
        self.BackpropagateFnc(in_, out, out_diff, in_diff)
        return in_diff

    def fresh_gradient(self):
        """Zero-copy tensor view of the fresh-gradient arena (what a data-parallel caller all-reduces)."""
        return self._engine.arena_tensor(2)

    @property
    def engine(self):
        return self._engine


class TimeShift:
    """Mirror of the standard version's TimeShift component (standard/nnet/nnet-time-shift.h): out row dst = in row
    clamp(dst + shift, 0, rows - 1) as one device gather; Backpropagate is a no-op there (:53-56) and here."""
    MARKER = "<TimeShift>"

    def __init__(self, dim, shift=0):
        self.input_dim_ = self.output_dim_ = int(dim)
        self.shift_ = int(shift)

    def GetType(self):
        return "TimeShift"

    def InitData(self, config):
        toks = config.split()
        i = 0
        while i < len(toks):
            if toks[i] == "<Shift>":
                self.shift_ = int(toks[i + 1])
                i += 2
            else:
                raise RuntimeError("Unknown token %s, a typo in config? (Shift)" % toks[i])   # nnet-time-shift.h:27-28

    def PropagateFnc(self, in_, out):
        from .engine import time_shift
        time_shift(in_, out, self.shift_)

    def Propagate(self, in_):
        import torch
        out = torch.empty_like(in_)
        self.PropagateFnc(in_, out)
        return out

    def BackpropagateFnc(self, in_, out, out_diff, in_diff):
        return None
