"""One launch of every hot kernel of round 2, for `ncu --set full` (see profiles/README.md for the command):
  A  512^3 transport symbol: Fx, Fy, recurrence middle pass (tile kernel), By, Bx
  B  512^3 same symbol, CPC_OPT_Z_RECURRENCE = 0: the fused forward-FFT / division / backward-FFT middle pass
  C  512^3 non-separable Diag: the table form of the fused middle pass
  D  512 x 512 x 1024, line form of the recurrence: end-value sweep + thread-per-line solve
  E  256^3 x 4 wave block: narrow x pass, y pass, fused z pass with the 4 x 4 arrow solve
No warm-up applies: every compute kernel below is launched exactly once per section."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import circulantpreconditioner_b200 as cpc

lam = (55.5556,) * 3
n = 512
b = torch.randn(n ** 3, dtype=torch.float64, device="cuda").to(torch.complex128)
x = torch.empty_like(b)
with cpc.CirculantPlan(n, n, n) as p:
    p.set_symbol_transport(*lam)
    p.apply(b, x)                                   # A
    p.set_option("z_recurrence", 0)
    p.apply(b, x)                                   # B
    d = torch.from_numpy(p.get_diag()).cuda()
    d[5] += 0.25
    p.set_symbol_diag(d)
    assert p.info()["symbol_kind"] == 2
    p.apply(b, x)                                   # C
    del d
torch.cuda.synchronize()
del b, x
b = torch.randn(512 * 512 * 1024, dtype=torch.float64, device="cuda").to(torch.complex128)
x = torch.empty_like(b)
with cpc.CirculantPlan(512, 512, 1024) as p:
    p.set_symbol_transport(*lam)
    p.set_option("z_line_form", 1)
    p.apply(b, x)                                   # D
torch.cuda.synchronize()
del b, x
m = 256
b = torch.randn(4 * m ** 3, dtype=torch.float64, device="cuda").to(torch.complex128)
x = torch.empty_like(b)
with cpc.CirculantPlan(m, m, m, ncomp=4) as p:
    p.set_symbol_wave(700.0, 0.0793651, 0.0793651, 0.0793651)
    p.apply(b, x)                                   # E
torch.cuda.synchronize()
print("ncu_targets done")
