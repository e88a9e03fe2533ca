import sys, json, torch
sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic, _cabi
from ptv_interpolation_b200.engine import PTVEngine, set_tuning, _ptr
n = 1024
dev = torch.device("cuda", 0)
eng = PTVEngine(dev)
g = torch.Generator(device=dev); g.manual_seed(0)
u, v, w = (torch.randn((n, n, n), device=dev, dtype=torch.float32, generator=g) for _ in range(3))
m = synthetic.fcc_sphere_pack_mask(n, device=dev).view(torch.uint8)
div = torch.empty_like(u)
acc = torch.zeros(2 + 3 * n, dtype=torch.float64, device=dev)
lib = eng.lib
def run(flux, h):
    q = (_ptr(acc[2:2+n]), _ptr(acc[2+n:2+2*n]), _ptr(acc[2+2*n:])) if flux else (None, None, None)
    _cabi.check(lib.ptv_divergence_flux(_ptr(u), _ptr(v), _ptr(w), _ptr(m), n, n, n, h, h, h, None, None, None, 0,
                                        _ptr(div), _ptr(acc[:2]), q[0], q[1], q[2], None))
for bulk, rows, flux, h in [(0, 0, 1, 1.0), (1, 0, 1, 1.0), (1, 0, 0, 1.0), (1, 64, 1, 1.0), (1, 32, 1, 1.0), (0, 0, 1, 2.00625), (1, 0, 1, 2.00625)]:
    set_tuning(stencil_bulk=bulk, stencil_rows=rows)
    b = 1e9
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(flux, h); e1.record(); torch.cuda.synchronize()
        if it: b = min(b, e0.elapsed_time(e1))
    print(json.dumps({"bulk": bulk, "rows": rows, "flux": flux, "h": h, "ms": b, "frac": 17.0 * n**3 / b / 1e6 / 6466.5}))
