"""Run the reference's UNMODIFIED scripts on the CUDA path by aliasing the drop-in modules under the
names the scripts import (``interpolator``, ``physics``, ``filtering``, ``velocity_analysis``):

    python -m ptv_interpolation_b200.compat /path/to/ptv_interpolation/main.py --ptv data.csv --mask m.tif ...

Everything the drop-in modules do not provide (viewers, sparse variational cleaning, ...) is taken from
the script's own directory; a function that exists on neither side raises as usual.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types

_ALIASES = ("interpolator", "physics", "filtering", "velocity_analysis")


def install(fallback_dir: str | None = None) -> None:
    """Put merged modules into ``sys.modules``: names defined by the CUDA drop-ins win, every other
    name falls through to the reference's own module of that name (if ``fallback_dir`` holds one)."""
    for name in _ALIASES:
        ours = importlib.import_module(f"ptv_interpolation_b200.{name}")
        merged = types.ModuleType(name)
        merged.__dict__["__ptv_b200__"] = True
        ref_path = os.path.join(fallback_dir, name + ".py") if fallback_dir else None
        if ref_path and os.path.exists(ref_path):
            spec = importlib.util.spec_from_file_location(f"_ptv_reference_{name}", ref_path)
            ref = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(ref)
                merged.__dict__.update({k: v for k, v in ref.__dict__.items() if not k.startswith("__")})
            except ImportError:  # e.g. tifffile / matplotlib missing: the drop-in names still work
                pass
        public = getattr(ours, "__all__", [k for k in ours.__dict__ if not k.startswith("_")])
        merged.__dict__.update({k: getattr(ours, k) for k in public})
        sys.modules[name] = merged


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    sys.path.insert(0, os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
