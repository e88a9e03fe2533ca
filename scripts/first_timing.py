"""Ad-hoc timing of the hot path on the synthetic configs (development aid, not the bench)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from ptv_interpolation_b200 import synthetic
from ptv_interpolation_b200.engine import PTVEngine, set_tuning


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2"]
    variants = [dict()] + [eval("dict(%s)" % a) for a in sys.argv[2:]]
    dev = torch.device("cuda", 0)
    eng = PTVEngine(dev)
    for name in names:
        cfg = synthetic.make_config(name, device=dev)
        n = cfg["n"]
        mask = cfg["mask"].view(torch.uint8)
        ax = torch.linspace(0, n - 1, n, dtype=torch.float64, device=dev)
        pore = int(cfg["mask"].sum())
        for var in variants:
            set_tuning(tile=128, r0=1, ppc=0.5, stream=1, stream_tile=128, stats=0, rscale=1.3)
            set_tuning(**var)
            for masked in (True, False):
                if not masked and n > 512:
                    continue
                ts = []
                for it in range(3):
                    torch.cuda.synchronize()
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                    eng.build(cfg["points"], cfg["values"])
                    e1.record()
                    out = eng.interpolate(ax, ax, ax, mask=mask if masked else None, method=cfg["method"], k=cfg["k"])
                    e2.record()
                    torch.cuda.synchronize()
                    ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
                tb, ti = min(t[0] for t in ts), min(t[1] for t in ts)
                set_tuning(stats=1)
                eng.interpolate(ax, ax, ax, mask=mask if masked else None, method=cfg["method"], k=cfg["k"])
                st = eng.knn_stats()
                set_tuning(stats=0)
                nv = pore if masked else n ** 3
                print(f"{name} {var} masked={masked}: build {tb:.2f} ms, interp {ti:.1f} ms, "
                      f"{nv / ti / 1e3:.1f} Mvox/s ({'pore' if masked else 'all'}), stats {st}", flush=True)
                del out


if __name__ == "__main__":
    main()
