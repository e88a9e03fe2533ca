"""Run the strain-rate / vorticity kernel on a synthetic n^3 float32 field (for ncu / timing)."""
import sys
import torch
sys.path.insert(0, ".")
from ptv_interpolation_b200.engine import PTVEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
g = torch.Generator(device=dev); g.manual_seed(0)
u, v, w = (torch.randn((n, n, n), device=dev, dtype=torch.float32, generator=g) for _ in range(3))
m = (torch.rand((n, n, n), device=dev, generator=g) > 0.6).view(torch.uint8)
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = eng.strain_vorticity(u, v, w, 1.0, 1.0, 1.0, mask=m)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"n={n} strain+vorticity {ms:.3f} ms  {21.0 * n**3 / ms / 1e6:.0f} GB/s algorithmic")
