import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc
n = 512
b = torch.randn(n ** 3, dtype=torch.float64, device="cuda").to(torch.complex128)
x = torch.empty_like(b)
for sg in [0, 300, 600, 1000, 1500, 2500, 4000, 8000, 0]:
    os.environ["CPC_STAGGER"] = str(sg)
    with cpc.CirculantPlan(n, n, n) as p:
        p.set_symbol_transport(55.5556, 55.5556, 55.5556)
        for _ in range(3):
            p.apply(b, x)
        torch.cuda.synchronize()
        reps, acc = 10, None
        for _ in range(reps):
            ms = p.apply_profiled(b, x)
            acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
        print(f"stagger {sg:5d}: passes " + " ".join(f"{a/reps:.3f}" for a in acc), flush=True)
