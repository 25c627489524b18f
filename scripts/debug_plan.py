import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lav_b200 import heads, ops
torch.manual_seed(0)
dev = torch.device("cuda:0")
gru = torch.nn.GRU(4, 512, batch_first=True).to(dev)
mlp = torch.nn.Linear(512, 2).to(dev)
for B in (2, 32, 64):
    embd = torch.randn(B, 512, device=dev) * 0.5
    nxp = torch.tensor([[0.0, -20.0]] * B, device=dev)
    cast = torch.randn(B, 6, 20, 2, device=dev) * 0.1
    with torch.no_grad():
        a = heads._plan_rollout(gru, mlp, 6, 20, 5, embd, nxp, cast, 4, 192, h16_ok=False)
        b = heads._plan_rollout(gru, mlp, 6, 20, 5, embd, nxp, cast, 4, 192, h16_ok=True)
    d = (a - b).abs()
    print(B, "plan rollout kernel-vs-cudnn max abs", float(d.max()), "scale", float(a.abs().max()), "per-frame max", [round(float(x), 4) for x in d.flatten(1).max(1)[0][:4]])
