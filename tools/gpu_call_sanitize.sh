#!/bin/bash
# compute-sanitizer on small cases of the new kernels (memcheck + racecheck + synccheck)
set -u
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
import kaldi_lstm_b200 as klb
for (I, C, R, S, T) in [(8, 32, 16, 8, 3), (40, 256, 128, 64, 2)]:
    comp = klb.LstmProjectedStreams(I, R, max_frames=T)
    comp.InitData("<CellDim> %d <NumStream> %d <ParamScale> 0.1" % (C, S), seed=1)
    x = torch.randn(T * S, I, device="cuda"); od = torch.randn(T * S, R, device="cuda") * 0.1
    for _ in range(2):
        out = comp.Propagate(x); d = comp.Backpropagate(x, out, od); comp.Update()
    torch.cuda.synchronize()
    print("ok", (I, C, R, S, T), comp.engine.info()["fwd_tensor_core"], float(out.abs().sum()), float(d.abs().sum()))
tail = klb.AffineSoftmaxXent(16, 40, max_frames=8)
tail.InitData("", seed=1)
xx = torch.randn(8, 16, device="cuda")
tail.PropagateEval(xx, np.ones(8, np.float32), [[(3, 1.0)]] * 8); tail.Backpropagate(xx); tail.Update()
torch.cuda.synchronize(); print("tail ok", tail.Stats())
PY
for tool in memcheck racecheck synccheck; do
  LSTMP_B200_BWD_COOP=0 timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|tail ok|Hazard|Invalid|error" gpurun_out/sanitizer_$tool.log | head -12
done
