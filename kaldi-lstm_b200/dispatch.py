"""Multi-stream chunk filling -- mirror of the trainer's stream dispatch
(google/nnetbin/bd-nnet-train-lstm-streams.cc:128-209).

Packs ``num_stream`` utterances into time-major BPTT chunks of ``batch_size`` frames
(row = t*S + s, :187-206), applies the targets delay by reading the feature row
``curt + targets_delay`` clamped to the last frame (:198-202), pads exhausted streams with
mask 0 and the last target (:190-196) and raises the per-stream reset flag when a stream takes
a new utterance (:170).  ``shard(rank, world)`` gives rank r of N every N-th utterance so that
N processes (one per GPU) run the same loop on disjoint data.
"""
import numpy as np


class StreamDispatcher:
    def __init__(self, num_stream, batch_size, targets_delay, feat_dim):
        self.S, self.T, self.delay, self.D = int(num_stream), int(batch_size), int(targets_delay), int(feat_dim)
        self.keys = [""] * self.S
        self.feats = [None] * self.S
        self.targets = [None] * self.S
        self.curt = np.zeros(self.S, np.int64)
        self.lent = np.zeros(self.S, np.int64)
        self.new_utt_flags = np.zeros(self.S, np.int32)
        self.num_done = 0
        self.num_skipped = 0
        self._it = None

    @staticmethod
    def shard(utterances, rank, world):
        for n, u in enumerate(utterances):
            if n % world == rank:
                yield u

    def _on_new_utterance(self, s):
        """Hook: stream s just took a new utterance (the device dispatcher uploads it here)."""

    def open(self, utterances):
        """utterances: iterable of (key, feats[frames x D] float32, targets[frames])."""
        self._it = iter(utterances)

    def _refill(self):
        # :146-174
        for s in range(self.S):
            if self.curt[s] < self.lent[s]:
                self.new_utt_flags[s] = 0
                continue
            while True:
                try:
                    key, feats, targets = next(self._it)
                except StopIteration:
                    break
                if targets is None:  # missing targets (:156-161)
                    self.num_skipped += 1
                    continue
                if feats.shape[0] != len(targets):  # length mismatch (:163-167)
                    self.num_skipped += 1
                    continue
                self.keys[s] = key
                self.feats[s] = np.asarray(feats, np.float32)
                self.targets[s] = np.asarray(targets)
                self.curt[s] = 0
                self.lent[s] = feats.shape[0]
                self.new_utt_flags[s] = 1
                self._on_new_utterance(s)
                break

    def next_chunk(self):
        """Returns (feat [T*S x D], frame_mask [T*S], target [T*S], new_utt_flags [S]) or None
        once every stream is exhausted (:177-181)."""
        self._refill()
        if not np.any(self.curt < self.lent):
            return None
        S, T = self.S, self.T
        feat = np.zeros((T * S, self.D), np.float32)
        mask = np.zeros(T * S, np.float32)
        target = np.zeros(T * S, np.int64)
        t_idx = np.arange(T)
        for s in range(S):
            L = int(self.lent[s])
            if L == 0:
                continue  # stream never got an utterance: rows stay zero, mask 0
            cur = self.curt[s] + t_idx                       # curt at each t (:204 increments every t)
            valid = cur < L                                   # :190
            rows = t_idx * S + s
            mask[rows] = valid.astype(np.float32)
            tgt_idx = np.where(valid, cur, L - 1)             # :192,:195
            target[rows] = self.targets[s][tgt_idx]
            f_idx = np.where(cur + self.delay < L, cur + self.delay, L - 1)  # :198-202
            feat[rows] = self.feats[s][f_idx]
            self.curt[s] += T
        self.num_done += int(self.new_utt_flags.sum())
        return feat, mask, target, self.new_utt_flags.copy()


class DeviceStreamDispatcher(StreamDispatcher):
    """The same bookkeeping with the chunk's FEATURE matrix assembled on the GPU (lstmp_b200_dispatch_*, SURVEY.md
    section 8(f) rank 3): an utterance is copied host -> device once, when a stream takes it (TRAIN.cc:152-170); per
    chunk one gather kernel applies the targets-delay shift, the last-frame padding and the AddShift + Rescale feature
    transform (google/feature_transform.nnet.txt:2-5) straight into the time-major chunk matrix.  frame_mask, target
    and new_utt_flags stay host arrays exactly as in the reference (Xent::EvalMasked takes them from the host)."""

    def __init__(self, num_stream, batch_size, targets_delay, feat_dim, max_utt_frames=4096, device=0, shift=None,
                 scale=None):
        super().__init__(num_stream, batch_size, targets_delay, feat_dim)
        import ctypes
        from . import engine as _e
        self._ct, self._e = ctypes, _e
        L = _e.load_library()
        vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.lstmp_b200_dispatch_create.argtypes = [ci, ci, ci, ci, ci, ci, ctypes.POINTER(vp)]
        L.lstmp_b200_dispatch_destroy.argtypes = [vp]
        L.lstmp_b200_dispatch_set_transform.argtypes = [vp, vp, vp]
        L.lstmp_b200_dispatch_load_utt.argtypes = [vp, ci, vp, sz, ci]
        L.lstmp_b200_dispatch_assemble.argtypes = [vp, vp, vp, vp, sz, vp]
        L.lstmp_b200_dispatch_get_stats.argtypes = [vp, vp]
        self._L = L
        h = vp()
        _e._chk(L.lstmp_b200_dispatch_create(self.S, self.T, self.delay, self.D, int(max_utt_frames), device,
                                             ctypes.byref(h)))
        self._h = h
        self.device = device
        if shift is not None or scale is not None:
            sh = None if shift is None else np.ascontiguousarray(shift, np.float32)
            sc = None if scale is None else np.ascontiguousarray(scale, np.float32)
            assert (sh is None or sh.size == self.D) and (sc is None or sc.size == self.D)
            _e._chk(L.lstmp_b200_dispatch_set_transform(h, None if sh is None else vp(sh.ctypes.data),
                                                        None if sc is None else vp(sc.ctypes.data)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.lstmp_b200_dispatch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _on_new_utterance(self, s):
        f = self.feats[s]
        self._e._chk(self._L.lstmp_b200_dispatch_load_utt(self._h, s, self._ct.c_void_p(f.ctypes.data), f.shape[1],
                                                          f.shape[0]))

    def stats(self):
        class _S(self._ct.Structure):
            _fields_ = [(n, self._ct.c_ulonglong) for n in ("kernel_launches", "h2d_bytes", "utterances_loaded",
                                                            "chunks_assembled")]
        st = _S()
        self._e._chk(self._L.lstmp_b200_dispatch_get_stats(self._h, self._ct.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in _S._fields_}

    def next_chunk(self, feat_out=None):
        """Returns (feat [T*S x D] CUDA tensor, frame_mask [T*S] host, target [T*S] host, new_utt_flags [S] host) or None
        once every stream is exhausted.  feat_out: optional preallocated CUDA float32 [T*S x D] matrix."""
        import torch
        self._refill()
        if not np.any(self.curt < self.lent):
            return None
        S, T = self.S, self.T
        if feat_out is None:
            feat_out = torch.empty((T * S, self.D), dtype=torch.float32, device="cuda:%d" % self.device)
        curt = np.ascontiguousarray(self.curt, np.int32)
        lent = np.ascontiguousarray(self.lent, np.int32)
        vp = self._ct.c_void_p
        self._e._chk(self._L.lstmp_b200_dispatch_assemble(self._h, vp(curt.ctypes.data), vp(lent.ctypes.data),
                                                          vp(feat_out.data_ptr()), feat_out.stride(0),
                                                          self._e._cur_stream(self.device)))
        mask = np.zeros(T * S, np.float32)
        target = np.zeros(T * S, np.int64)
        t_idx = np.arange(T)
        for s in range(S):
            L = int(self.lent[s])
            if L == 0:
                continue
            cur = self.curt[s] + t_idx
            valid = cur < L
            rows = t_idx * S + s
            mask[rows] = valid.astype(np.float32)
            target[rows] = self.targets[s][np.where(valid, cur, L - 1)]
            self.curt[s] += T
        self.num_done += int(self.new_utt_flags.sum())
        return feat_out, mask, target, self.new_utt_flags.copy()
