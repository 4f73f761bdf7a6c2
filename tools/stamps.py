import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LSTMP_B200_DEBUG"] = os.environ.get("LSTMP_B200_DEBUG", "4")
os.environ["LSTMP_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kaldi-lstm_b200", "_lib", "liblstmp_b200_stamps.so")
import torch
import kaldi_lstm_b200 as klb
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
comp = klb.LstmProjectedStreams(40, 512, max_frames=T)
comp.InitData("<CellDim> 800 <NumStream> %d <ParamScale> 0.01" % S)
x = torch.randn(T * S, 40, device="cuda")
od = torch.randn(T * S, 512, device="cuda") * 0.1
for _ in range(3):
    out = comp.Propagate(x)
    comp.BackpropagateFnc(x, out, od, None)
torch.cuda.synchronize()
print(comp.engine.info())
comp.engine.close()
