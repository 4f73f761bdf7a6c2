#!/usr/bin/env python
"""bench.py -- frames/sec of the LstmProjectedStreams BPTT hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one BPTT chunk of synthetic input: for every stacked
layer Propagate, then Backpropagate from a synthetic out_diff at the top, the gradient all-reduce
(N > 1), and Update -- what nnet.Propagate / nnet.Backpropagate do per chunk in
google/nnetbin/bd-nnet-train-lstm-streams.cc:209-229.  A frame is one (stream, timestep) row;
value = N * S * T * steps / seconds, the whole-job aggregate (weak scaling: S per GPU fixed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[2]: the shape north_star's target is quoted on (NumStream=64, 800/512)
    "cfg3": dict(layers=[(40, 800, 512), (512, 800, 512)], S=64, T=20,
                 desc="configs[2]: stacked 2x LstmProjectedStreams (40->800/512->800/512), NumStream=64 per GPU, "
                      "20-frame BPTT, fwd+bwd+update"),
    "cfg3-layer1": dict(layers=[(40, 800, 512)], S=64, T=20,
                        desc="configs[2] first layer only: LstmProjectedStreams 40->800/512, NumStream=64, T=20"),
    "cfg2": dict(layers=[(40, 800, 512)], S=4, T=20,
                 desc="configs[1]: LstmProjectedStreams 40->800/512, NumStream=4, 20-frame BPTT (recipe default)"),
    "cfg5": dict(layers=[(40, 2048, 1024)], S=64, T=20,
                 desc="configs[4] per-GPU shape: LstmProjectedStreams 40->2048/1024, NumStream=64 per GPU (512 over 8), "
                      "T=20 (weights-streamed mode: 43 MB of weights do not fit in shared memory)"),
    "cfg4-lstm": dict(layers=[(40, 800, 512)], S=32, T=20,
                      desc="configs[3] LSTM part: 40->800/512, NumStream=32 per GPU (256 over 8), T=20"),
}
LR, MOMENTUM, PARAM_SCALE = 1e-5, 0.9, 0.01  # google/train_lstm_streams.sh:3-8, google/nnet.proto:3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.load = False

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append((self.load, parts))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [p for load, p in self.rows if load] or [p for _, p in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(p[0]) for p in rows if p[0].replace(".", "").isdigit())
        mx = [float(p[1]) for p in rows if p[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(p[3 + k].lower().startswith("active") for p in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# ---------------------------------------------------------------------------------------------
# algorithmic bytes / flops (DESIGN.md "Kernels")
# ---------------------------------------------------------------------------------------------
def alg_fwd_kernel(I, C, R, S, T):
    ts = S * T
    rd = ts * 4 * C + (4 * C * R + R * C) + S * (C + R)          # pre-activations, weights once, carried state
    wr = ts * 4 * C + 3 * ts * C + 2 * ts * R + S * (C + R)      # g,i,f,o | c,h,m | r (record + out) | state
    flops = ts * (2 * R * 4 * C + 2 * C * R)
    return 4 * (rd + wr), flops


def alg_bwd_kernel(I, C, R, S, T):
    ts = S * T
    rd = ts * 4 * C + (T + 1) * S * C + ts * C + ts * R + (4 * C * R + R * C)  # g,i,f,o | c | h | out_diff | weights
    wr = ts * 4 * C + ts * R + 7 * C                                           # DGIFO | DR | bias+peephole grads
    flops = ts * (2 * R * 4 * C + 2 * C * R)
    return 4 * (rd + wr), flops


def alg_chunk(I, C, R, S, T):
    """F = S*T*(6*I*4C + 6*R*4C + 6*C*R); compulsory bytes B."""
    P = 4 * C * I + 4 * C * R + 7 * C + R * C
    F = S * T * (6 * I * 4 * C + 6 * R * 4 * C + 6 * C * R)
    B = S * T * (4 * I + 4 * (7 * C + R) + 4 * R) + S * T * (4 * R + 8 * (7 * C + R) + 4 * I) + 4 * P * 4
    return B, F


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle (restatement of the reference's CPU matrix path,
# cblas sgemm + serial elementwise loops) on this box's host cores
# ---------------------------------------------------------------------------------------------
class CpuStack:
    """The CPU arm: oracle/_ref (the reference's own LstmProjectedStreams compiled from /root/reference against a CPU
    Kaldi surface; kind "reference") when it was built, else the C restatement oracle/lstmp_streams_oracle.c
    (kind "port").  Same Python interface either way."""

    def __init__(self, wl, threads=None):
        from oracle import oracle_py, ref_py
        if threads is None:
            threads = min(os.cpu_count() or 1, 64)   # scipy's OpenBLAS is built for at most 64 threads
        if ref_py.available():
            self.kind, self.mod, Layer = "reference", ref_py, ref_py.RefLstm
            self.what = "oracle/_ref: the reference's own bd-nnet-lstm-projected-streams.h on its CPU matrix path"
        else:
            oracle_py.build()
            self.kind, self.mod, Layer = "port", oracle_py, oracle_py.Oracle
            self.what = "oracle/lstmp_streams_oracle.c (oracle/_ref not built on this machine)"
        self.threads = self.mod.use_openblas(threads) or 1
        self.blas = "openblas(scipy-bundled)" if self.threads > 0 and self.mod.use_openblas(threads) else "builtin-loops"
        self.S, self.T = wl["S"], wl["T"]
        self.layers = []
        for li, (I, C, R) in enumerate(wl["layers"]):
            o = Layer(I, C, R, self.S, np.float32)
            o.set_params(oracle_py.init_params(I, C, R, PARAM_SCALE, 4321 + li))
            self.layers.append(o)
        rng = np.random.RandomState(1234)
        I0, Rtop = wl["layers"][0][0], wl["layers"][-1][2]
        self.x = rng.randn(self.S * self.T, I0).astype(np.float32)
        self.od = (rng.randn(self.S * self.T, Rtop) * 0.1).astype(np.float32)

    def set_threads(self, n):
        return self.mod.use_openblas(n) or 0

    def step(self):
        acts = [self.x]
        for o in self.layers:
            acts.append(o.propagate(acts[-1]))
        d = self.od
        for li in reversed(range(len(self.layers))):
            d = self.layers[li].backpropagate(acts[li], d, MOMENTUM, want_in_diff=True)
        for o in self.layers:
            o.update(LR)

    def frames(self):
        return self.S * self.T

    def single_thread_sample(self, seconds=2.0, max_steps=20):
        """Kaldi's nnet1 default is a single-threaded BLAS: a short 1-thread sample."""
        try:
            if not self.set_threads(1):
                return None
            t1 = time.perf_counter()
            n1 = 0
            while n1 < 1 or (time.perf_counter() - t1 < seconds and n1 < max_steps):
                self.step()
                n1 += 1
            out = {"value": self.frames() * n1 / (time.perf_counter() - t1), "unit": "frames/s", "cores": 1,
                   "sample": "%d full chunks" % n1}
            self.set_threads(self.threads)
            return out
        except Exception as e:
            return {"error": str(e)[:200]}


def workload_config(args, wl, world):
    """The `config` object: identical for both arms (the driver compares them)."""
    I0, Rtop = wl["layers"][0][0], wl["layers"][-1][2]
    bytes_per_chunk = wl["S"] * wl["T"] * (I0 + Rtop) * 4
    ring = max(4, int(160e6 // bytes_per_chunk) + 1)
    return {"workload": wl["desc"], "name": args.workload, "num_stream_per_gpu": wl["S"], "bptt_frames": wl["T"],
            "learn_rate": LR, "momentum": MOMENTUM, "param_scale": PARAM_SCALE,
            "parallelism": ("streams sharded over %d GPU(s), 1 NCCL sum-allreduce of each layer's gradients per Update, "
                            "overlapped with the backward pass of the layer below" % world) if world > 1 else "1 GPU",
            "l2": "GPU arm: inputs larger than L2, a ring of %d distinct (feature, out_diff) chunks = %.0f MB" % (
                ring, ring * bytes_per_chunk / 1e6)}


def run_reference_arm(args, wl):
    stack = CpuStack(wl)
    for _ in range(args.warmup):
        stack.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stack.step()
    dt = time.perf_counter() - t0
    val = stack.frames() * args.steps / dt
    single = stack.single_thread_sample()
    line = {
        "impl": "reference", "metric": "frames/sec LstmProjectedStreams 800-cell/512-proj BPTT", "value": val,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, wl, args.gpus),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": stack.threads, "kind": stack.kind,
                         "sample": "%d full chunks of the workload; %s; sgemm=%s on %d threads (host has %d cpus), "
                                   "elementwise loops serial as in kaldi-matrix.cc" % (
                                       args.steps, stack.what, stack.blas, stack.threads, os.cpu_count()),
                         "single_thread": single},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def bench_xent(args):
    """masked cross-entropy with sparse targets at the cfg 4 per-GPU chunk
    (S=32 x T=20 = 640 frames x 16624 pdfs).  One step = one EvalMasked call (CSR posterior H2D + row kernel +
    fixed-order reduction); metric frames/s; HBM-bound: 8 bytes per (frame, pdf)."""
    import numpy as np
    rows, P = 640, 16624
    cfg = {"workload": "xent-cfg4: Xent::EvalMasked, 640 frames (S=32 x T=20) x 16624 pdfs, hard labels, 1/5 masked",
           "l2": "ring of 16 distinct net_out matrices = 681 MB > L2"}
    from oracle import xent_oracle
    if args.impl == "reference":
        mask, y, post = xent_oracle.random_case(rows, P, seed=1, mask_every=5)
        o = xent_oracle.XentOracle()
        for _ in range(args.warmup):
            o.eval_masked(mask, y, post)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            o.eval_masked(mask, y, post)
        dt = time.perf_counter() - t0
        v = rows * args.steps / dt
        print(json.dumps({"impl": "reference", "metric": "frames/sec Xent::EvalMasked", "value": v, "unit": "frames/s",
                          "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": 1, "kind": "port",
                                           "sample": "%d dense EvalMasked calls (numpy restatement)" % args.steps},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    import torch
    import kaldi_lstm_b200 as klb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the product has no CPU path (use --impl reference for the CPU arm)")
    nring = 16
    g = torch.Generator(device="cuda").manual_seed(1)
    ys = [torch.softmax(torch.randn(rows, P, device="cuda", generator=g) * 2, dim=1) for _ in range(nring)]
    diff = torch.empty(rows, P, device="cuda")
    rng = np.random.RandomState(2)
    post = (np.arange(rows + 1, dtype=np.int32), rng.randint(0, P, rows).astype(np.int32), np.ones(rows, np.float32))
    mask = (np.arange(rows) % 5 != 0).astype(np.float32)
    x = klb.Xent(rows)
    for i in range(max(args.warmup, 3)):
        x.EvalMasked(mask, ys[i % nring], post, diff)
    torch.cuda.synchronize()
    l0 = x.Stats()["kernel_launches"]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(args.steps):
        x.EvalMasked(mask, ys[i % nring], post, diff)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    launches = x.Stats()["kernel_launches"] - l0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    alg = 8.0 * rows * P
    ach = alg / (ms / args.steps * 1e-3) / 1e9
    o = xent_oracle.XentOracle()
    m2, y2, p2 = xent_oracle.random_case(rows, P, seed=1, mask_every=5)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < min(args.cpu_seconds, 10.0):
        o.eval_masked(m2, y2, p2)
        n += 1
    dt = time.perf_counter() - t0
    print(json.dumps({"metric": "frames/sec Xent::EvalMasked", "value": rows * args.steps / (ms * 1e-3), "unit": "frames/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": cfg, "gpu_launches": int(launches),
                      "e2e": {"value": rows * args.steps / (ms * 1e-3), "unit": "frames/s",
                              "h2d_bytes_per_step": int(4 * (rows + 1) + 12 * rows), "d2h_bytes_per_step": 0,
                              "what": "the timed call already takes the host posterior / mask and copies them"},
                      "roofline": {"kernel": "xent_rows_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                                   "frac": ach / hbm, "traffic": None, "alg_bytes_per_launch": alg,
                                   "note": "whole-call time (H2D of the CSR posterior + 2 kernels), not the kernel alone"},
                      "cpu_baseline": {"value": rows * n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                                       "sample": "%d dense EvalMasked calls (numpy restatement of nnet-loss.cc:76-164)" % n}}))
    return 0


def mgpu_parity_check(klb, exchange, dev, rank, world):
    """N > 1, before the timed region: (1) N ranks x S/N streams through the engine's own NCCL entry point
    (lstmp_b200_allreduce_grads_nccl) reproduce ONE engine running all S streams on this rank's GPU; (2) the replicas stay
    bit-identical across ranks.  Small shape, 3 chunks with state carry-over, momentum 0.9.  Returns a dict for the JSON
    line; never raises (a failure is reported, not hidden)."""
    import torch
    import torch.distributed as dist
    try:
        I, C, R, T, nchunks = 40, 256, 128, 6, 3
        Sl = 16
        Stot = Sl * world
        full = klb.LstmProjectedStreams(I, R, device=dev.index, max_frames=T)
        full.InitData("<CellDim> %d <NumStream> %d <ParamScale> 0.1" % (C, Stot), seed=99)
        full.SetTrainOptions(klb.NnetTrainOptions(1e-3, 0.9))
        part = klb.LstmProjectedStreams(I, R, device=dev.index, max_frames=T)
        part.InitData("<CellDim> %d <NumStream> %d" % (C, Sl))
        part.SetParams(full.GetParams())
        part.SetTrainOptions(klb.NnetTrainOptions(1e-3, 0.9))
        ex = klb.parallel.GradientExchange([part], dev)
        g = torch.Generator(device=dev).manual_seed(4242)      # same data on every rank
        worst = 0.0
        for n in range(nchunks):
            x = torch.randn(T, Stot, I, device=dev, generator=g)
            od = torch.randn(T, Stot, R, device=dev, generator=g) * 0.1
            of = full.Propagate(x.reshape(T * Stot, I))
            full.Backpropagate(x.reshape(T * Stot, I), of, od.reshape(T * Stot, R), want_in_diff=False)
            full.Update()
            xs = x[:, rank * Sl:(rank + 1) * Sl].reshape(T * Sl, I).contiguous()
            ods = od[:, rank * Sl:(rank + 1) * Sl].reshape(T * Sl, R).contiguous()
            op = part.Propagate(xs)
            part.Backpropagate(xs, op, ods, want_in_diff=False)
            ex.start(0)
            ex.finish(0)
            part.Update()
            ref_out = of.reshape(T, Stot, R)[:, rank * Sl:(rank + 1) * Sl].reshape(T * Sl, R)
            worst = max(worst, float((op - ref_out).abs().max() / ref_out.abs().max()))
        pf = torch.from_numpy(full.GetParams()).to(dev)
        pp = torch.from_numpy(part.GetParams()).to(dev)
        err = float((pp - pf).abs().max() / pf.abs().max())
        lo, hi = pp.clone(), pp.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(torch.equal(lo, hi))
        t = torch.tensor([err, worst], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ex.close()
        return {"ok": bool(identical and float(t[0]) <= 1e-4 and float(t[1]) <= 1e-4), "replicas_bit_identical": identical,
                "params_vs_single_engine_rel": float(t[0]), "out_vs_single_engine_rel": float(t[1]),
                "shape": "%d ranks x %d streams vs 1 x %d streams, 40->256/128, T=%d, %d chunks" % (world, Sl, Stot, T, nchunks),
                "allreduce": "lstmp_b200_allreduce_grads_nccl on an own ncclComm_t, side stream"}
    except Exception as e:  # report, do not hide
        return {"ok": False, "error": str(e)[:300]}


def bench_cfg4(args):
    """BASELINE.json configs[3] as a whole network: LstmProjectedStreams 40->800/512 + AffineTransform 512->16624 +
    Softmax + Xent::EvalMasked, NumStream=256 sharded over 8 GPUs = 32 streams (640 frames) per GPU, T=20, fwd + bwd +
    [all-reduce] + update of both components.  One step = Reset, LSTM Propagate, tail PropagateEval (host mask + sparse
    posterior in), tail Backpropagate, LSTM Backpropagate, Update x2."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    S, T, I0, C, R, P = 32, 20, 40, 800, 512, 16624
    rows = S * T
    desc = ("configs[3]: LstmProjectedStreams 40->800/512 + AffineTransform 512->16624 + Softmax + masked xent, "
            "NumStream=32 per GPU (256 over 8), 20-frame BPTT, fwd+bwd+update")
    cfg = {"workload": desc, "name": "cfg4", "num_stream_per_gpu": S, "bptt_frames": T, "learn_rate": LR,
           "momentum": MOMENTUM, "param_scale": PARAM_SCALE,
           "parallelism": ("streams sharded over %d GPU(s), NCCL sum-allreduce of each component's gradients per Update, "
                           "overlapped" % world) if world > 1 else "1 GPU",
           "l2": "GPU arm: per step 42.6 MB of logits/diff + 34 MB of tail weights stream through L2; ring of 16 chunks"}
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle_py, ref_py, tail_oracle
        Layer = ref_py.RefLstm if ref_py.available() else oracle_py.Oracle
        mod = ref_py if ref_py.available() else oracle_py
        if not ref_py.available():
            oracle_py.build()
        threads = mod.use_openblas(min(os.cpu_count() or 1, 64)) or 1
        lstm = Layer(I0, C, R, S, np.float32)
        lstm.set_params(oracle_py.init_params(I0, C, R, PARAM_SCALE, 4321))
        tail = tail_oracle.TailOracle(R, P, np.float32)
        rng = np.random.RandomState(1)
        tail.set_params(np.concatenate([(rng.randn(P * R) * 0.1), -2.0 + (rng.rand(P) - 0.5) * 2.0]).astype(np.float32))
        x = rng.randn(rows, I0).astype(np.float32)
        mask = (np.arange(rows) % 5 != 0).astype(np.float32)
        post = [[(int(rng.randint(0, P)), 1.0)] for _ in range(rows)]

        def step():
            h = lstm.propagate(x)
            tail.propagate_eval(h, mask, post)
            d = tail.backpropagate(h, MOMENTUM).astype(np.float32)
            lstm.backpropagate(x, d, MOMENTUM, want_in_diff=False)
            tail.update(LR)
            lstm.update(LR)
        for _ in range(max(1, args.warmup)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        v = rows * args.steps / dt
        kind = "reference" if ref_py.available() else "port"
        print(json.dumps({"impl": "reference", "metric": "frames/sec LstmProjectedStreams 800-cell/512-proj BPTT", "value": v,
                          "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": kind,
                                           "sample": "%d full chunks; LSTM = %s, tail = numpy restatement of the upstream "
                                                     "AffineTransform / Softmax + the reference's EvalMasked formulation"
                                                     % (args.steps, "oracle/_ref" if kind == "reference" else "oracle port")},
                          "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import kaldi_lstm_b200 as klb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("LSTMP_B200_BENCH_NCCL_ALGO", "Ring") != "none":
            os.environ.setdefault("NCCL_ALGO", os.environ.get("LSTMP_B200_BENCH_NCCL_ALGO", "Ring"))
        dist.init_process_group("nccl", device_id=dev)
    lstm = klb.LstmProjectedStreams(I0, R, device=local_rank, max_frames=T)
    lstm.InitData("<CellDim> %d <NumStream> %d <ParamScale> %g" % (C, S, PARAM_SCALE), seed=4321)
    lstm.SetTrainOptions(klb.NnetTrainOptions(LR, MOMENTUM))
    tail = klb.AffineSoftmaxXent(R, P, device=local_rank, max_frames=rows)
    tail.InitData("<ParamStddev> 0.1 <BiasMean> -2.0 <BiasRange> 2.0", seed=1)
    tail.SetTrainOptions(klb.NnetTrainOptions(LR, MOMENTUM))
    comps = [lstm, tail]
    exchange = klb.parallel.GradientExchange(comps, dev) if world > 1 else None
    rng = np.random.RandomState(2 + rank)
    nring = 16
    X = torch.randn(nring, rows, I0, device=dev)
    masks = [(np.arange(rows) % 5 != (k % 5)).astype(np.float32) for k in range(nring)]
    posts = [(np.arange(rows + 1, dtype=np.int32), rng.randint(0, P, rows).astype(np.int32), np.ones(rows, np.float32))
             for _ in range(nring)]
    out = torch.empty(rows, R, device=dev)
    od = torch.empty(rows, R, device=dev)

    def compute(x, mask, post, flags):
        lstm.Reset(flags)
        lstm.PropagateFnc(x, out)
        tail.PropagateEval(out, mask, post)
        tail.BackpropagateFnc(out, od)
        if exchange is not None:
            exchange.start(1)
        lstm.BackpropagateFnc(x, out, od, None)
        if exchange is not None:
            exchange.start(0)
            exchange.finish(0)
        lstm.Update()
        if exchange is not None:
            exchange.finish(1)
        tail.Update()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def launches():
        return lstm.engine.info()["kernel_launches"] + tail.Stats()["kernel_launches"]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(max(3, args.warmup)):
        compute(X[i % nring], masks[i % nring], posts[i % nring], [1 if (i + s) % 50 == 0 else 0 for s in range(S)])
    sync_all()
    sampler.load = True
    l0 = launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        compute(X[i % nring], masks[i % nring], posts[i % nring], [1 if (i + s) % 50 == 0 else 0 for s in range(S)])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    gl = launches() - l0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * rows * args.steps / (ms * 1e-3)

    # end to end: host utterances + host targets through the device dispatcher, loss read back every step (lagged by one)
    e2e = None
    if not args.no_e2e:
        prng = np.random.RandomState(777 + rank)
        pool = []
        for k in range(96):
            L = int(prng.randint(300, 701))
            pool.append(("u%d" % k, prng.randn(L, I0).astype(np.float32), prng.randint(0, P, size=L)))

        def utterances():
            k = 0
            while True:
                yield pool[k % len(pool)]
                k += 1
        disp = klb.DeviceStreamDispatcher(S, T, 5, I0, max_utt_frames=704, device=local_rank,
                                          shift=(prng.randn(I0) * 3).astype(np.float32),
                                          scale=(0.25 + 0.001 * np.arange(I0)).astype(np.float32))
        disp.open(utterances())
        xd = [torch.empty(rows, I0, device=dev) for _ in range(2)]
        rp = np.arange(rows + 1, dtype=np.int32)
        ones = np.ones(rows, np.float32)

        def e2e_step(i):
            feat, mask, target, flags = disp.next_chunk(feat_out=xd[i & 1])
            compute(feat, mask, (rp, target.astype(np.int32), ones), flags.tolist())
        for i in range(3):
            e2e_step(i)
        tail.Stats()
        sync_all()
        st0 = disp.stats()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        loss = None
        for i in range(args.steps):
            e2e_step(3 + i)
            if i % 10 == 9:
                loss = tail.Stats()["loss"]          # Report()-style D2H of the accumulated statistics (32 bytes, syncs)
        loss = tail.Stats()["loss"]
        s1.record()
        sync_all()
        st1 = disp.stats()
        ems = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * rows * args.steps / (ems * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": int((st1["h2d_bytes"] - st0["h2d_bytes"]) / args.steps + 4 * (rows + 1) + 12 * rows),
               "d2h_bytes_per_step": 3.2, "ms_per_step": ems / args.steps, "loss": loss,
               "what": "host utterances + host targets -> DeviceStreamDispatcher -> LSTM -> fused tail (mask + CSR posterior "
                       "H2D from pinned staging) -> backward -> update; accumulated loss statistics D2H every 10 steps"}
        disp.close()

    lstm.engine.timing_enable(True)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nprof = min(args.steps, 30)
    a0.record()
    for i in range(nprof):
        compute(X[i % nring], masks[i % nring], posts[i % nring], [0] * S)
    a1.record()
    torch.cuda.synchronize()
    kinds = lstm.engine.timing_read()
    lstm.engine.timing_enable(False)
    lstm_us = 1e3 * sum(v[0] for v in kinds.values()) / nprof
    step_us = 1e3 * a0.elapsed_time(a1) / nprof
    sampler.load = False
    if rank == 0:
        sampler.stop()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops", peaks.get("tensor_tflops", 1500.0)))
    tail_flops = 3 * 2.0 * rows * R * P
    tail_us = max(step_us - lstm_us, 1e-3)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", "cfg4", "--impl", "reference",
                                "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=300)
            cpu_baseline = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as e:
            cpu_baseline = {"error": str(e)[:200]}
    if rank == 0:
        print(json.dumps({
            "metric": "frames/sec LstmProjectedStreams 800-cell/512-proj BPTT", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "e2e": e2e, "gpu_launches": int(gl), "clocks": sampler.summary(),
            "roofline": {"kernel": "tail: 3 tcgen05 3xTF32 GEMMs (logits, in_diff, W gradient) + fused softmax/xent",
                         "bound": "tensor", "achieved": tail_flops / (tail_us * 1e-6) / 1e12, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": tail_flops / (tail_us * 1e-6) / 1e12 / tf_peak, "traffic": None,
                         "note": "logical fp32 flops of the three tail GEMMs / (step time - LSTM kernel time); the 3xTF32 "
                                 "split issues 3x that on the tensor pipe", "tail_us_per_step": tail_us,
                         "lstm_us_per_step": lstm_us},
            "cpu_baseline": cpu_baseline}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["xent-cfg4", "cfg4"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the extra cfg2 (NumStream=4) measurement")
    args = ap.parse_args()
    if args.workload == "xent-cfg4":
        return bench_xent(args)
    if args.workload == "cfg4":
        return bench_cfg4(args)
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, wl)
        return 0

    import torch
    import torch.distributed as dist

    import kaldi_lstm_b200 as klb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the gradient all-reduces are 9-34 MB: on 8 B200s NCCL's default choice measured 1.327 ms per step, the ring
        # algorithm 1.289 ms (profiles/r2_nccl_algo_n8.txt); NCCL_ALGO set by the user wins
        if os.environ.get("LSTMP_B200_BENCH_NCCL_ALGO", "Ring") != "none":
            os.environ.setdefault("NCCL_ALGO", os.environ.get("LSTMP_B200_BENCH_NCCL_ALGO", "Ring"))
        dist.init_process_group("nccl", device_id=dev)
    S, T = wl["S"], wl["T"]
    rows = S * T

    # ---- model: random-init weights of the named architecture (no checkpoints offline) ----------
    layers = []
    for li, (I, C, R) in enumerate(wl["layers"]):
        comp = klb.LstmProjectedStreams(I, R, device=local_rank, max_frames=T)
        comp.InitData("<CellDim> %d <NumStream> %d <ParamScale> %g" % (C, S, PARAM_SCALE), seed=4321 + li)
        comp.SetTrainOptions(klb.NnetTrainOptions(LR, MOMENTUM))
        layers.append(comp)
    grad_views = [c.engine.arena_tensor(2) for c in layers]
    I0, Rtop = wl["layers"][0][0], wl["layers"][-1][2]

    # ---- synthetic data: ring of distinct chunks whose total size exceeds the 126 MB L2 ---------
    bytes_per_chunk = rows * (I0 + Rtop) * 4
    ring = max(4, int(160e6 // bytes_per_chunk) + 1)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    X = torch.randn(ring, rows, I0, device=dev, generator=g)                  # "40-dim fbank" after CMVN ~ N(0,1)
    OD = torch.randn(ring, rows, Rtop, device=dev, generator=g) * 0.1         # synthetic loss gradient
    outs = [torch.empty(rows, R, device=dev) for (_, _, R) in wl["layers"]]
    in_diffs = [None] + [torch.empty(rows, I, device=dev) for (I, _, _) in wl["layers"][1:]]

    exchange = klb.parallel.GradientExchange(layers, dev) if world > 1 else None

    def compute(x, od, i, flags=None):
        if flags is None:   # staggered synthetic utterance boundaries: stream s starts a new utterance every 50 chunks
            flags = [1 if (i + s) % 50 == 0 else 0 for s in range(S)]
        h = x
        for li, comp in enumerate(layers):
            comp.Reset(flags)                                 # nnet.Reset(new_utt_flags), TRAIN.cc:209
            comp.PropagateFnc(h, outs[li])                    # nnet.Propagate, TRAIN.cc:215
            h = outs[li]
        d = od
        for li in reversed(range(len(layers))):               # nnet.Backpropagate, TRAIN.cc:228
            inp = x if li == 0 else outs[li - 1]
            layers[li].BackpropagateFnc(inp, outs[li], d, in_diffs[li])
            d = in_diffs[li]
            if exchange is not None:                          # layer li's gradient is final: all-reduce it on the side
                exchange.start(li)                            # stream while the layers below run their backward
        for li, comp in enumerate(layers):
            if exchange is not None:
                exchange.finish(li)                           # Update(li) waits for its own all-reduce only
            comp.Update()

    def launches():
        return sum(c.engine.info()["kernel_launches"] for c in layers)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    mgpu_parity = mgpu_parity_check(klb, exchange, dev, rank, world) if world > 1 else None

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- warm-up, then the timed region (device-resident inputs) ---------------------------------
    for i in range(max(3, args.warmup)):
        compute(X[i % ring], OD[i % ring], i)
    sync_all()
    sampler.load = True
    l0 = launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        compute(X[i % ring], OD[i % ring], i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    gpu_launches = launches() - l0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * rows * args.steps / (ms * 1e-3)

    # ---- end-to-end: host utterances in, host scalar out, through the same public calls --------------
    # The trainer's loop (TRAIN.cc:143-229): utterances live in HOST memory; the stream dispatcher uploads each one
    # once when a stream takes it (pinned staging, own copy stream) and assembles the time-major chunk on the device
    # with the delay shift, padding and the CMVN transform; the loss-gradient stand-in goes H2D from pinned memory on a
    # copy stream (double-buffered); every step's result (4-byte checksum of the top output) comes back D2H and is
    # read by the host one step later (the next step's uploads are already in flight then).
    e2e = None
    if not args.no_e2e:
        rng = np.random.RandomState(777 + rank)
        pool = [("u%d" % k, rng.randn(int(rng.randint(300, 701)), I0).astype(np.float32)) for k in range(192)]
        pool = [(k, f, rng.randint(0, 8000, size=f.shape[0])) for k, f in pool]

        def utterances():
            k = 0
            while True:
                yield pool[k % len(pool)]
                k += 1
        shift = (rng.randn(I0) * 3).astype(np.float32)
        scale = (0.25 + 0.001 * np.arange(I0)).astype(np.float32)
        disp = klb.DeviceStreamDispatcher(S, T, 5, I0, max_utt_frames=704, device=local_rank, shift=shift, scale=scale)
        disp.open(utterances())
        nh = 8
        ODh = [(torch.randn(rows, Rtop) * 0.1).pin_memory() for _ in range(nh)]
        xd = [torch.empty(rows, I0, device=dev) for _ in range(2)]
        odd = [torch.empty(rows, Rtop, device=dev) for _ in range(2)]
        res = [torch.empty(1).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        od_ready = [torch.cuda.Event() for _ in range(2)]
        od_free = [torch.cuda.Event() for _ in range(2)]
        res_ready = [torch.cuda.Event() for _ in range(2)]
        main_stream = torch.cuda.current_stream(dev)

        def e2e_enqueue(i):
            b = i & 1
            feat, mask, target, flags = disp.next_chunk(feat_out=xd[b])   # utterance uploads + gather kernel
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(od_free[b])
                odd[b].copy_(ODh[i % nh], non_blocking=True)     # loss gradient stand-in (targets go H2D in LOSS.cc:96)
                od_ready[b].record(copy_stream)
            main_stream.wait_event(od_ready[b])
            compute(xd[b], odd[b], i, flags=flags.tolist())
            od_free[b].record(main_stream)
            res[b].copy_(outs[-1][rows - S:].sum().reshape(1), non_blocking=True)   # progress metric D2H
            res_ready[b].record(main_stream)

        def e2e_collect(i):
            res_ready[i & 1].synchronize()
            return float(res[i & 1][0])

        for i in range(4):
            e2e_enqueue(i)
            if i:
                e2e_collect(i - 1)
        e2e_collect(3)
        sync_all()
        st0 = disp.stats()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            e2e_enqueue(4 + i)
            if i:
                e2e_collect(4 + i - 1)
        e2e_collect(4 + args.steps - 1)
        s1.record()
        sync_all()
        st1 = disp.stats()
        ems = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": world * rows * args.steps / (ems * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": int((st1["h2d_bytes"] - st0["h2d_bytes"]) / args.steps + rows * Rtop * 4),
               "d2h_bytes_per_step": 4, "ms_per_step": ems / args.steps,
               "utterances_uploaded": st1["utterances_loaded"] - st0["utterances_loaded"],
               "what": "host utterances -> DeviceStreamDispatcher (each utterance H2D once from pinned staging on its own "
                       "copy stream; delay shift, padding and CMVN in one gather kernel) + pinned loss-gradient chunk H2D "
                       "on a copy stream, Reset/Propagate/Backpropagate/[allreduce]/Update through the component API, "
                       "4-byte output checksum D2H every step, read by the host one step later"}
        disp.close()

    # ---- per-kernel device times (engine-side CUDA events on the launch stream) -------------------
    for c in layers:
        c.engine.timing_enable(True)
    nprof = min(args.steps, 50)
    for i in range(nprof):
        compute(X[i % ring], OD[i % ring], i)
    kinds = {}
    for c in layers:
        for k, (tms, cnt) in c.engine.timing_read().items():
            a = kinds.setdefault(k, [0.0, 0])
            a[0] += tms
            a[1] += cnt
        c.engine.timing_enable(False)
    sampler.load = False
    sync_all()
    tot = sum(v[0] for v in kinds.values()) or 1.0
    kernels = {k: {"us_per_launch": 1e3 * v[0] / max(v[1], 1), "launches_per_step": v[1] / nprof,
                   "share": v[0] / tot} for k, v in kinds.items() if v[1]}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    dom = max(("fwd_recurrent", "bwd_recurrent"), key=lambda k: kinds.get(k, [0, 0])[0])
    I_, C_, R_ = wl["layers"][0]
    ab, af = (alg_bwd_kernel if dom == "bwd_recurrent" else alg_fwd_kernel)(I_, C_, R_, S, T)
    dom_us = kernels.get(dom, {}).get("us_per_launch", float("nan"))
    ach = ab / (dom_us * 1e-6) / 1e9 if dom_us == dom_us else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    inf0 = layers[0].engine.info()
    fwd_name = {2: "lstmp_fwd_tma_kernel", 1: "lstmp_fwd_tc_kernel"}.get(inf0.get("fwd_tensor_core"), "lstmp_fwd_kernel")
    bwd_name = "lstmp_bwd_tma_kernel" if inf0.get("bwd_tensor_core") == 2 else "lstmp_bwd_kernel"
    roofline = {"kernel": fwd_name if dom == "fwd_recurrent" else bwd_name, "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                "unit": "GB/s", "frac": (ach / hbm_peak) if ach else None, "traffic": traffic,
                "peak_source": peak_src, "alg_bytes_per_launch": ab, "alg_flops_per_launch": af,
                "us_per_launch": dom_us, "achieved_tflops_fp32": af / (dom_us * 1e-6) / 1e12 if ach else None}
    cb, cf = 0, 0
    for (I, C, R) in wl["layers"]:
        b_, f_ = alg_chunk(I, C, R, S, T)
        cb, cf = cb + b_, cf + f_
    step_us = 1e3 * ms / args.steps
    chunk_roofline = {"alg_bytes_per_step": cb, "alg_flops_per_step": cf,
                      "hbm_bound_us": cb / (hbm_peak * 1e9) * 1e6,
                      "frac_of_hbm_bound": (cb / (hbm_peak * 1e9) * 1e6) / step_us}

    if rank == 0:
        sampler.stop()
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        stack = CpuStack(wl)
        stack.step()
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < args.cpu_seconds and n < 200):
            stack.step()
            n += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": stack.frames() * n / dt, "unit": "frames/s", "cores": stack.threads, "kind": stack.kind,
                        "sample": "%d full chunks (%.1f s) of the same workload; %s; sgemm=%s on %d threads (host has "
                                  "%d cpus), elementwise loops serial as in kaldi-matrix.cc" % (
                                      n, dt, stack.what, stack.blas, stack.threads, os.cpu_count()),
                        "single_thread": stack.single_thread_sample()}

    secondary = None
    if rank == 0 and world == 1 and args.workload == "cfg3" and not args.no_secondary:
        # BASELINE.json configs[1] (NumStream=4, the shipped recipe's default) measured in the same run, device-resident
        try:
            wl2 = WORKLOADS["cfg2"]
            c2 = klb.LstmProjectedStreams(40, 512, device=local_rank, max_frames=wl2["T"])
            c2.InitData("<CellDim> 800 <NumStream> %d <ParamScale> %g" % (wl2["S"], PARAM_SCALE), seed=4321)
            c2.SetTrainOptions(klb.NnetTrainOptions(LR, MOMENTUM))
            r2 = wl2["S"] * wl2["T"]
            x2 = torch.randn(64, r2, 40, device=dev)
            od2 = torch.randn(64, r2, 512, device=dev) * 0.1
            o2 = torch.empty(r2, 512, device=dev)

            def step2(i):
                c2.PropagateFnc(x2[i % 64], o2)
                c2.BackpropagateFnc(x2[i % 64], o2, od2[i % 64], None)
                c2.Update()
            for i in range(10):
                step2(i)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            n2 = 200
            for i in range(n2):
                step2(i)
            a1.record()
            torch.cuda.synchronize()
            m2 = a0.elapsed_time(a1)
            secondary = {"cfg2": {"workload": wl2["desc"], "value": r2 * n2 / (m2 * 1e-3), "unit": "frames/s",
                                  "ms_per_step": m2 / n2, "steps": n2}}
        except Exception as e:  # never let the secondary measurement break the contract line
            secondary = {"cfg2": {"error": str(e)[:200]}}
        # forward-only (cross-validate mode) frames/s of the main workload, device-resident inputs
        try:
            def fwd_only(i):
                flags = [1 if (i + s) % 50 == 0 else 0 for s in range(S)]
                hcur = X[i % ring]
                for li, comp in enumerate(layers):
                    comp.Reset(flags)
                    comp.PropagateFnc(hcur, outs[li])
                    hcur = outs[li]
            for i in range(5):
                fwd_only(i)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nf = 100
            f0.record()
            for i in range(nf):
                fwd_only(i)
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1)
            secondary["forward_only"] = {"workload": wl["desc"] + " -- Reset + Propagate only", "value": rows * nf / (fms * 1e-3),
                                         "unit": "frames/s", "ms_per_step": fms / nf, "steps": nf}
        except Exception as e:
            secondary["forward_only"] = {"error": str(e)[:200]}

        # the loss of configs[3] at its per-GPU chunk (640 frames x 16624 pdfs): Xent::EvalMasked with sparse targets,
        # and the fused softmax + xent of the output tail; device-resident net_out ring larger than L2
        try:
            xr, xP, xn = 640, 16624, 4
            ys = [torch.softmax(torch.randn(xr, xP, device=dev) * 2, dim=1) for _ in range(xn)]
            xdiff = torch.empty(xr, xP, device=dev)
            xrng = np.random.RandomState(2)
            xpost = (np.arange(xr + 1, dtype=np.int32), xrng.randint(0, xP, xr).astype(np.int32), np.ones(xr, np.float32))
            xmask = (np.arange(xr) % 5 != 0).astype(np.float32)
            xe = klb.Xent(xr, device=local_rank)
            for i in range(5):
                xe.EvalMasked(xmask, ys[i % xn], xpost, xdiff)
            torch.cuda.synchronize()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            nx = 100
            for i in range(nx):
                xe.EvalMasked(xmask, ys[i % xn], xpost, xdiff)
            x1.record()
            torch.cuda.synchronize()
            xms = x0.elapsed_time(x1) / nx
            secondary["xent_cfg4"] = {"workload": "Xent::EvalMasked, 640 frames x 16624 pdfs, sparse hard targets, 1/5 masked "
                                                  "(whole call: CSR H2D + 2 kernels)", "value": xr / (xms * 1e-3),
                                      "unit": "frames/s", "us_per_call": 1e3 * xms,
                                      "achieved_gbs": 8.0 * xr * xP / (xms * 1e-3) / 1e9, "alg_bytes": 8 * xr * xP}
        except Exception as e:
            secondary["xent_cfg4"] = {"error": str(e)[:200]}

    # ---- N > 1: the other configurations BASELINE.json names for 8 GPUs, device-resident, same timing rules --------
    if world > 1 and not args.no_secondary:
        def measure_stack(shapes, S2, T2, nsteps):
            ls = []
            for li, (I, C, R) in enumerate(shapes):
                c = klb.LstmProjectedStreams(I, R, device=local_rank, max_frames=T2)
                c.InitData("<CellDim> %d <NumStream> %d <ParamScale> %g" % (C, S2, PARAM_SCALE), seed=4321 + li)
                c.SetTrainOptions(klb.NnetTrainOptions(LR, MOMENTUM))
                ls.append(c)
            ex = klb.parallel.GradientExchange(ls, dev)
            r2 = S2 * T2
            nr = 8
            xs = torch.randn(nr, r2, shapes[0][0], device=dev)
            ods = torch.randn(nr, r2, shapes[-1][2], device=dev) * 0.1
            o2 = [torch.empty(r2, R, device=dev) for (_, _, R) in shapes]
            idf = [None] + [torch.empty(r2, I, device=dev) for (I, _, _) in shapes[1:]]

            def step(i):
                h = xs[i % nr]
                for li, c in enumerate(ls):
                    c.PropagateFnc(h, o2[li])
                    h = o2[li]
                d = ods[i % nr]
                for li in reversed(range(len(ls))):
                    ls[li].BackpropagateFnc(xs[i % nr] if li == 0 else o2[li - 1], o2[li], d, idf[li])
                    d = idf[li]
                    ex.start(li)
                for li, c in enumerate(ls):
                    ex.finish(li)
                    c.Update()
            for i in range(3):
                step(i)
            sync_all()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for i in range(nsteps):
                step(i)
            a1.record()
            sync_all()
            t = torch.tensor([a0.elapsed_time(a1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ex.close()
            m = float(t.item())
            return {"value": world * r2 * nsteps / (m * 1e-3), "unit": "frames/s", "ms_per_step": m / nsteps,
                    "steps": nsteps, "num_stream_per_gpu": S2, "num_stream_total": S2 * world}
        secondary = {}
        try:
            if 256 % world == 0:
                r = measure_stack([(40, 800, 512)], 256 // world, 20, 50)
                r["workload"] = ("configs[3] LSTM part, STRONG scaling: 40->800/512, NumStream=256 in total = %d per GPU, "
                                 "T=20" % (256 // world))
                secondary["cfg4_lstm_strong"] = r
        except Exception as e:
            secondary["cfg4_lstm_strong"] = {"error": str(e)[:200]}
        try:
            r = measure_stack([(40, 2048, 1024)], 64, 20, 10)
            r["workload"] = "configs[4]: 40->2048/1024, NumStream=64 per GPU (weak), T=20 (weights-streamed mode)"
            secondary["cfg5"] = r
        except Exception as e:
            secondary["cfg5"] = {"error": str(e)[:200]}

    if rank == 0:
        info = layers[0].engine.info()
        line = {
            "metric": "frames/sec LstmProjectedStreams 800-cell/512-proj BPTT", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "arithmetic": "fp32 storage, inputs, outputs and accumulation; every product runs on tcgen05 kind::f16 with BOTH "
                          "operands split into bf16 hi + lo pieces and all four cross terms accumulated in fp32 (TMEM): "
                          "relative error 4e-6 end to end against the fp32 oracle, tolerance 1e-4",
            "config": workload_config(args, wl, world),
            "engine": {k: info[k] for k in ("fwd_tensor_core", "bwd_tensor_core", "ngroups", "ctas_per_group",
                                            "streams_per_group", "cells_per_cta", "rcols_per_cta", "bwd_ctas",
                                            "bwd_cluster", "gemm_backend")},
            "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": sampler.summary(), "roofline": roofline,
            "chunk_roofline": chunk_roofline, "kernels": kernels, "cpu_baseline": cpu_baseline,
            "secondary": secondary, "mgpu_parity": mgpu_parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
