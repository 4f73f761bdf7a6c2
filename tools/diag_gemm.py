"""GPU diagnostic for the tcgen05 GEMM: which K blocks / layouts contribute correctly."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kaldi_lstm_b200 import engine
engine.load_library()
torch.manual_seed(0)

def run(M, N, K, tA, tB, A, B):
    C = torch.zeros(M, N, device="cuda")
    engine.debug_gemm(1, C, M, N, K, 1.0, A, tA, B, tB, 0.0, None)
    torch.cuda.synchronize()
    return C

def ref(A, tA, B, tB):
    return ((A.t() if tA else A).double() @ (B.t() if tB else B).double())

print("== per-K-block contribution, K-major/K-major, M=128 N=128")
for K in (32, 64, 96, 128, 160, 256):
    M = N = 128
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    full = run(M, N, K, 0, 1, A, B)
    r = ref(A, 0, B, 1)
    print("K=%d total rel err %.3e" % (K, (full.double() - r).abs().max().item() / r.abs().max().item()))
    for kb in range(K // 32):
        Am = torch.zeros_like(A); Am[:, kb*32:(kb+1)*32] = A[:, kb*32:(kb+1)*32]
        c = run(M, N, K, 0, 1, Am, B); rr = ref(Am, 0, B, 1)
        print("   kb=%d rel err %.3e  (|c|max %.3f |ref|max %.3f)" % (kb, (c.double()-rr).abs().max().item()/rr.abs().max().item(), c.abs().max().item(), rr.abs().max().item()))

print("== repeat same launch 3x K=128 (determinism)")
A = torch.randn(128, 128, device="cuda"); B = torch.randn(128, 128, device="cuda")
outs = [run(128, 128, 128, 0, 1, A, B) for _ in range(3)]
print([float((o - outs[0]).abs().max()) for o in outs], float((outs[0].double()-ref(A,0,B,1)).abs().max()))

print("== layout decode: A MN-major (stored KxM), B K-major identity; K=8")
M, N, K = 128, 128, 8
A = torch.zeros(K, M, device="cuda")
for k in range(K):
    for m in range(M):
        A[k, m] = m + 1000 * k
B = torch.zeros(N, K, device="cuda")
for k in range(K):
    B[k, k] = 1.0
C = run(M, N, K, 1, 1, A, B)
print("expect C[m][n] = m + 1000 n for n<8")
print(C[:12, :8].cpu().numpy().astype(int))
print(C[60:68, :8].cpu().numpy().astype(int))

print("== layout decode: A K-major identity-ish, B MN-major (stored KxN); K=8")
A = torch.zeros(M, K, device="cuda")
for k in range(K):
    A[k, k] = 1.0
B = torch.zeros(K, N, device="cuda")
for k in range(K):
    for n in range(N):
        B[k, n] = n + 1000 * k
C = run(M, N, K, 0, 0, A, B)
print("expect C[m][n] = n + 1000 m for m<8")
print(C[:8, :12].cpu().numpy().astype(int))
print(C[:8, 60:68].cpu().numpy().astype(int))

print("== MN/MN random K=8,16,32,64")
for K in (8, 16, 32, 64, 96):
    A = torch.randn(K, 128, device="cuda"); B = torch.randn(K, 128, device="cuda")
    c = run(128, 128, K, 1, 0, A, B); r = ref(A, 1, B, 0)
    print("K=%d rel err %.3e" % (K, (c.double()-r).abs().max().item()/r.abs().max().item()))
print("== K/K BN=64 (N=40) K=32,128")
for K in (32, 128):
    A = torch.randn(128, K, device="cuda"); B = torch.randn(40, K, device="cuda")
    c = run(128, 40, K, 0, 1, A, B); r = ref(A, 0, B, 1)
    print("K=%d rel err %.3e" % (K, (c.double()-r).abs().max().item()/r.abs().max().item()))

print("== shape1 variants (K/K 1280x3200x512)")
import itertools
M, N, K = 1280, 3200, 512
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C0 = torch.randn(M, N, device="cuda"); bias = torch.randn(N, device="cuda")
for alpha, beta, useb in ((1.0, 0.0, False), (0.7, 0.0, False), (1.0, 0.3, False), (1.0, 0.0, True), (0.7, 0.3, True)):
    C = C0.clone()
    engine.debug_gemm(1, C, M, N, K, alpha, A, 0, B, 1, beta, bias if useb else None)
    torch.cuda.synchronize()
    r = alpha * ref(A, 0, B, 1) + beta * C0.double() + (bias.double() if useb else 0)
    e = (C.double() - r).abs()
    bad = (e > 1e-3 * r.abs().max()).nonzero()
    print("alpha=%.1f beta=%.1f bias=%d relerr %.3e nbad %d" % (alpha, beta, useb, e.max().item()/r.abs().max().item(), bad.shape[0]), bad[:5].tolist())
