"""cProfile of the small-N loss+grad step (host-side overhead hunt). Dev tool."""
import cProfile, pstats, sys, os, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gptorch_b200 import kernels
from gptorch_b200.models import GPR
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
np.random.seed(42)
x = np.linspace(0, 1, n).reshape((-1, 1)); y = np.sin(6 * x) + 0.1 * np.random.randn(n, 1)
model = GPR(x, y, kernels.Linear(1) + kernels.Rbf(1) + kernels.Constant(1))
params = [p for p in model.parameters() if p.requires_grad]
def step():
    for p in params: p.grad = None
    loss = model.loss(); loss.backward(); return loss
for _ in range(30): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue()[:9000])
