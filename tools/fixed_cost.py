"""Per-launch fixed cost vs per-timestep cost of the recurrent kernels: engine event times at several chunk lengths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kaldi_lstm_b200 as klb
S = 64
res = {}
for T in (2, 4, 10, 20):
    comp = klb.LstmProjectedStreams(40, 512, max_frames=T)
    comp.InitData("<CellDim> 800 <NumStream> %d <ParamScale> 0.01" % S)
    x = torch.randn(8, T * S, 40, device="cuda")
    od = torch.randn(8, T * S, 512, device="cuda") * 0.1
    out = torch.empty(T * S, 512, device="cuda")
    for i in range(5):
        comp.PropagateFnc(x[i % 8], out); comp.BackpropagateFnc(x[i % 8], out, od[i % 8], None); comp.Update()
    comp.engine.timing_enable(True)
    n = 40
    for i in range(n):
        comp.PropagateFnc(x[i % 8], out); comp.BackpropagateFnc(x[i % 8], out, od[i % 8], None); comp.Update()
    t = comp.engine.timing_read()
    res[T] = {k: 1e3 * v[0] / max(v[1], 1) for k, v in t.items() if v[1]}
    print(T, {k: round(v, 1) for k, v in res[T].items()})
for k in ("fwd_recurrent", "bwd_recurrent"):
    per = (res[20][k] - res[4][k]) / 16.0
    print(k, "per step %.2f us, fixed %.1f us (T=4), %.1f us (T=20)" % (per, res[4][k] - 4 * per, res[20][k] - 20 * per))
