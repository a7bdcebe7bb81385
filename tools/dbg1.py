import faulthandler, sys, os
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import circulantpreconditioner_b200 as cpc
print("create", flush=True)
p = cpc.CirculantPlan(4, 1, 1)
print("created", p.info(), flush=True)
p.set_symbol_transport(1.0)
print("symbol", flush=True)
b = torch.arange(4, dtype=torch.complex128, device="cuda")
x = p.apply(b)
torch.cuda.synchronize()
print(x, flush=True)
p2 = cpc.CirculantPlan(32, 32, 32)
p2.set_symbol_transport(1.0, 1.0, 1.0)
b = torch.ones(32**3, dtype=torch.complex128, device="cuda")
x = p2.apply(b); torch.cuda.synchronize(); print(x[:4], flush=True)
