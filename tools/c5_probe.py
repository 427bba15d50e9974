import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from simple_tip_b200 import engine as E
dev = E.require_cuda()
n_train, n_test, d, classes = 160000, 5000, 2048, 1000
g = torch.Generator(device=dev).manual_seed(5)
centres = torch.randn((classes, d), generator=g, device=dev) * 0.5
ytr = torch.arange(n_train, device=dev) % classes
order = torch.argsort(ytr, stable=True)
t_sorted = (centres[ytr[order]] + torch.randn((n_train, d), generator=g, device=dev))
class_off = np.arange(classes + 1, dtype=np.int64) * (n_train // classes)
eng = E.NnEngine(t_sorted, class_off, order.to(torch.int32))
yte = torch.arange(n_test, device=dev) % classes
x = centres[yte] + torch.randn((n_test, d), generator=g, device=dev)
q_order = torch.argsort(yte, stable=True)
x = x[q_order].contiguous(); q_class = yte[q_order].to(torch.int32).contiguous()
q_off = np.arange(classes + 1, dtype=np.int64) * (n_test // classes)
for _ in range(2):
    E.dsa_distances(eng, x, q_class, q_off)
torch.cuda.synchronize()
print("stats", eng.stats.cpu().numpy())
for mode, (cnt, idx) in eng.last_cand_cnt_by_mode.items():
    c = cnt.cpu().numpy(); print(mode, "entries/query mean", c.mean(), "max", c.max(), "overflow", (c > eng.cap).sum())
