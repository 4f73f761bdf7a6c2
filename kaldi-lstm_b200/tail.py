"""Host-side mirror of the network's output tail as ONE component:
`<AffineTransform> num_pdf input_dim` + `<Softmax>` (google/nnet.proto:4-5) + the objective the trainer evaluates on
them (`Xent::EvalMasked`, google/nnet/nnet-loss.cc:76-164; google/nnetbin/bd-nnet-train-lstm-streams.cc:215-228), over
the C ABI's lstmp_b200_tail_* entry points.  Parameters are the flat vector [linearity_ (num_pdf x input_dim) | bias_],
the order of AffineTransform::GetParams upstream."""
import ctypes

import numpy as np

from . import engine as _e
from .component import NnetTrainOptions
from .loss import posterior_to_csr


class TailEngine:
    def __init__(self, input_dim, num_pdf, max_frames, device=0):
        L = _e.load_library()
        vp, ci, sz, fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
        L.lstmp_b200_tail_create.argtypes = [ci, ci, ci, ci, ctypes.POINTER(vp)]
        L.lstmp_b200_tail_destroy.argtypes = [vp]
        L.lstmp_b200_tail_arena.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(sz)]
        L.lstmp_b200_tail_set_flat.argtypes = [vp, ci, vp, vp]
        L.lstmp_b200_tail_get_flat.argtypes = [vp, ci, vp, vp]
        L.lstmp_b200_tail_propagate_eval.argtypes = [vp, vp, sz, ci, vp, vp, vp, vp, vp, sz, vp]
        L.lstmp_b200_tail_backpropagate.argtypes = [vp, vp, sz, vp, sz, ci, vp]
        L.lstmp_b200_tail_update.argtypes = [vp, fp, fp, vp]
        L.lstmp_b200_tail_allreduce_grads_nccl.argtypes = [vp, vp, vp]
        L.lstmp_b200_tail_get_diff.argtypes = [vp, vp, sz, vp]
        L.lstmp_b200_tail_get_stats.argtypes = [vp, vp, vp]
        L.lstmp_b200_tail_reset_stats.argtypes = [vp, vp]
        self._L = L
        h = vp()
        _e._chk(L.lstmp_b200_tail_create(int(input_dim), int(num_pdf), int(max_frames), device, ctypes.byref(h)))
        self._h = h
        self.I, self.P, self.max_frames, self.device = int(input_dim), int(num_pdf), int(max_frames), device
        self.num_params = self.P * self.I + self.P
        self._rows = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.lstmp_b200_tail_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return _e._cur_stream(self.device)

    def set_flat(self, which, arr):
        a = np.ascontiguousarray(arr, np.float32)
        assert a.size == self.num_params
        _e._chk(self._L.lstmp_b200_tail_set_flat(self._h, which, ctypes.c_void_p(a.ctypes.data), self._stream()))

    def get_flat(self, which):
        out = np.empty(self.num_params, np.float32)
        _e._chk(self._L.lstmp_b200_tail_get_flat(self._h, which, ctypes.c_void_p(out.ctypes.data), self._stream()))
        return out

    def arena_tensor(self, which):
        """Zero-copy torch view of arena `which` (0 params, 1 momentum-accumulated, 2 fresh grads)."""
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _e._chk(self._L.lstmp_b200_tail_arena(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        import torch
        return torch.as_tensor(_e._CudaArray(p.value, n.value, self), device="cuda:%d" % self.device)

    def propagate_eval(self, x, frame_mask_host, row_ptr, pdf, weight, post_out=None):
        px, ldx = _e.Engine._mat(x, self.I, "in")
        mask = np.ascontiguousarray(frame_mask_host, np.float32)
        rp = np.ascontiguousarray(row_ptr, np.int32)
        pd = np.ascontiguousarray(pdf, np.int32)
        wt = np.ascontiguousarray(weight, np.float32)
        rows = x.shape[0]
        if mask.size != rows or rp.size != rows + 1 or pd.size != wt.size:
            raise _e.EngineError(_e.EINVAL, "mask / posterior sizes do not match the %d frames" % rows)
        if post_out is not None:
            pp, ldp = _e.Engine._mat(post_out, self.P, "post_out")
        else:
            pp, ldp = None, 0
        vp = ctypes.c_void_p
        _e._chk(self._L.lstmp_b200_tail_propagate_eval(
            self._h, px, ldx, rows, vp(mask.ctypes.data), vp(rp.ctypes.data), vp(pd.ctypes.data) if pd.size else None,
            vp(wt.ctypes.data) if wt.size else None, pp, ldp, self._stream()))
        self._rows = rows

    def backpropagate(self, x, in_diff=None):
        px, ldx = _e.Engine._mat(x, self.I, "in")
        if in_diff is not None:
            pi, ldi = _e.Engine._mat(in_diff, self.I, "in_diff")
        else:
            pi, ldi = None, 0
        _e._chk(self._L.lstmp_b200_tail_backpropagate(self._h, px, ldx, pi, ldi, x.shape[0], self._stream()))

    def update(self, learn_rate, momentum):
        _e._chk(self._L.lstmp_b200_tail_update(self._h, float(learn_rate), float(momentum), self._stream()))

    def allreduce_grads_nccl(self, comm_ptr, stream_ptr=None):
        st = ctypes.c_void_p(stream_ptr) if stream_ptr is not None else self._stream()
        _e._chk(self._L.lstmp_b200_tail_allreduce_grads_nccl(self._h, ctypes.c_void_p(comm_ptr), st))

    def get_diff(self):
        out = np.empty((self._rows, self.P), np.float32)
        _e._chk(self._L.lstmp_b200_tail_get_diff(self._h, ctypes.c_void_p(out.ctypes.data), self.P, self._stream()))
        return out

    def stats(self):
        st = _e.XentStats()
        _e._chk(self._L.lstmp_b200_tail_get_stats(self._h, ctypes.byref(st), self._stream()))
        return {"loss": st.loss, "entropy": st.entropy, "correct": int(st.correct), "frames": int(st.frames),
                "kernel_launches": int(st.kernel_launches)}

    def reset_stats(self):
        _e._chk(self._L.lstmp_b200_tail_reset_stats(self._h, self._stream()))


class AffineSoftmaxXent:
    """AffineTransform + Softmax + Xent::EvalMasked.  Propagate of the pair and the loss evaluation are one call
    (PropagateEval), Backpropagate + the gradient part of AffineTransform::Update another, Update the momentum step."""
    MARKER = "<AffineTransform>"

    def __init__(self, input_dim, output_dim, device=0, max_frames=640):
        self.input_dim_, self.output_dim_ = int(input_dim), int(output_dim)
        self.opts_ = NnetTrainOptions()
        self._device, self._max_frames = device, int(max_frames)
        self._engine = TailEngine(self.input_dim_, self.output_dim_, self._max_frames, device)

    def InitData(self, config="", seed=0):
        """[upstream] AffineTransform::InitData: <ParamStddev> (0.1), <BiasMean> (-2.0), <BiasRange> (2.0):
        linearity ~ N(0, 1) * stddev, bias ~ mean + (U(0,1) - 0.5) * range."""
        stddev, bmean, brange = 0.1, -2.0, 2.0
        toks = config.split()
        i = 0
        while i < len(toks):
            if toks[i] == "<ParamStddev>":
                stddev = float(toks[i + 1])
            elif toks[i] == "<BiasMean>":
                bmean = float(toks[i + 1])
            elif toks[i] == "<BiasRange>":
                brange = float(toks[i + 1])
            elif toks[i] in ("<LearnRateCoef>", "<BiasLearnRateCoef>", "<MaxNorm>"):
                if float(toks[i + 1]) not in (1.0, 0.0):
                    raise RuntimeError("%s other than the default is not supported by the fused tail" % toks[i])
            else:
                raise RuntimeError("Unknown token %s, a typo in config?" % toks[i])
            i += 2
        rng = np.random.RandomState(seed)
        W = (rng.randn(self.output_dim_, self.input_dim_) * stddev).astype(np.float32)
        b = (bmean + (rng.rand(self.output_dim_) - 0.5) * brange).astype(np.float32)
        self.SetParams(np.concatenate([W.ravel(), b]))

    def SetTrainOptions(self, opts):
        self.opts_ = opts

    def NumParams(self):
        return self._engine.num_params

    def GetParams(self):
        return self._engine.get_flat(0)

    def SetParams(self, flat):
        self._engine.set_flat(0, flat)

    def GetGradients(self):
        return self._engine.get_flat(1)

    def PropagateEval(self, in_, frame_mask_host, post, want_posteriors=False):
        import torch
        rows = in_.shape[0]
        if rows > self._engine.max_frames:
            prm, corr, st = self._engine.get_flat(0), self._engine.get_flat(1), self._engine.stats()
            assert st["frames"] == 0, "AffineSoftmaxXent created for fewer frames than this call needs"
            self._engine = TailEngine(self.input_dim_, self.output_dim_, rows, self._device)
            self._engine.set_flat(0, prm)
            self._engine.set_flat(1, corr)
        row_ptr, pdf, weight = post if isinstance(post, tuple) else posterior_to_csr(post)
        assert rows == len(row_ptr) - 1                                              # KALDI_ASSERT nnet-loss.cc:80
        y = torch.empty((rows, self.output_dim_), dtype=torch.float32, device=in_.device) if want_posteriors else None
        try:
            self._engine.propagate_eval(in_, frame_mask_host, row_ptr, pdf, weight, y)
        except Exception as e:  # KALDI_ERR on a pdf-id outside the network output       nnet-loss.cc:88-91
            if "pdf-id" in str(e):
                raise RuntimeError(str(e))
            raise
        return y

    def Backpropagate(self, in_, want_in_diff=True):
        import torch
        in_diff = torch.empty_like(in_) if want_in_diff else None
        self._engine.backpropagate(in_, in_diff)
        return in_diff

    def BackpropagateFnc(self, in_, in_diff):
        self._engine.backpropagate(in_, in_diff)

    def Update(self):
        self._engine.update(self.opts_.learn_rate, self.opts_.momentum)

    def Stats(self):
        return self._engine.stats()

    def Report(self):
        s = self.Stats()
        f = s["frames"] if s["frames"] else float("nan")
        return ("AvgLoss: %g (Xent), [AvgXent: %g, AvgTargetEnt: %g]\n\nFRAME_ACCURACY >> %g%% <<"
                % ((s["loss"] - s["entropy"]) / f, s["loss"] / f, s["entropy"] / f, 100.0 * s["correct"] / f))

    def fresh_gradient(self):
        return self._engine.arena_tensor(2)

    @property
    def engine(self):
        return self._engine
