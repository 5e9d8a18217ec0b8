"""One-hot probe of the tcgen05 wgrad kernel: where does D[r][p] * h_prev[r][k] land?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200")); sys.path.insert(0, ROOT)
import torch
from b200rl import ops as O
dev = torch.device("cuda:0")
ops = O.default()
B, T, Fin, NH = 4, 8, 4, 3
L = O.NetLayout(Fin, NH)
rows = B * T
def run(D, Hm, variant):
    flat = torch.zeros(L.P, device=dev)
    ws = dict(gates=D.view(B, T, 64, 4).contiguous(), h=Hm.view(B, T, 64).contiguous(), c=torch.zeros(B, T, 64, device=dev))
    hx0 = torch.zeros(B, 64, device=dev); u = torch.zeros(B, T, Fin, device=dev); dhead = torch.zeros(B, T, NH, device=dev)
    partials = torch.zeros(1, L.P, device=dev)
    ops.lstm_wgrad(L, flat, ws, hx0, u, dhead, partials, variant=variant)
    torch.cuda.synchronize()
    o, n = L.off["W_hh"]
    return partials[0, o:o + n].view(256, 64).clone()
torch.manual_seed(0)
for (r, p, k) in [(1, 0, 0), (1, 5, 0), (1, 0, 5), (2, 0, 0), (9, 0, 0), (1, 4, 0), (1, 0, 4), (1, 37, 21), (17, 255, 63), (3, 1, 1)]:
    D = torch.zeros(rows, 256, device=dev); Hm = torch.zeros(rows, 64, device=dev)
    D[r, p] = 1.0; Hm[r - 1, k] = 1.0          # h_prev[r] = h[r-1] (r % T != 0)
    g1 = run(D, Hm, 1); g0 = run(D, Hm, 0)
    nz1 = g1.nonzero().tolist(); nz0 = g0.nonzero().tolist()
    print(f"r={r} p={p} (n={(p&3)*64+(p>>2)}) k={k}: simt {nz0} {[float(g0[i,j]) for i,j in nz0]} | tc {nz1[:8]} {[round(float(g1[i,j]),4) for i,j in nz1[:8]]}")
# random full test
D = torch.randn(rows, 256, device=dev); Hm = torch.randn(rows, 64, device=dev)
g1 = run(D, Hm, 1); g0 = run(D, Hm, 0)
print("random: max|simt|", float(g0.abs().max()), "max diff", float((g1 - g0).abs().max()))
