"""Drop-in mirror of the reference's ``filtering.py``: global speed threshold and
the kNN median/MAD outlier filter, the latter on the CUDA path (self-query of the spatial hash with
k+1 neighbours + per-particle median / MAD of the neighbour speeds in the same kernel)."""
from __future__ import annotations

import numpy as np

from .engine import default_engine

__all__ = ["remove_outliers_knn", "remove_outliers_threshold", "apply_filters"]


def _drop_rows(df, keep, removed_message):
    """Rows where ``keep`` is False are dropped and the index renumbered; an untouched frame is returned
    as is (callers rely on identity when nothing was removed, filtering.py:53-58,69-73)."""
    dropped = int(keep.size - np.count_nonzero(keep))
    if dropped == 0:
        return df, 0
    print(removed_message(dropped))
    return df.loc[keep].reset_index(drop=True), dropped


def remove_outliers_knn(df, k=25, threshold=3.0, device=None):
    """filtering.py:5-58 -- median/MAD test of every particle's speed against its k nearest neighbours.
    One self-query of the spatial hash with k+1 neighbours; median, MAD and the decision are taken in the
    same kernel (ptv_outlier_filter)."""
    import torch
    n = len(df)
    if n <= k:
        print(f"  Warning: DataFrame too small ({n}) for k-NN filter (k={k}). Skipping.")
        return df
    eng = default_engine(device)
    xyz, uvw = (torch.from_numpy(np.ascontiguousarray(df[list(cols)].values, dtype=np.float64)).to(eng.device)
                for cols in ("xyz", "uvw"))
    eng.build(xyz, uvw)
    keep_dev, kth_dev = eng.outlier_filter(k=k, threshold=threshold)
    radius = float(np.median(kth_dev.cpu().numpy()))
    print(f"  Filtering radius: median voxel distance to {k}-th neighbor = {radius:.4f}")
    out, dropped = _drop_rows(df, keep_dev.cpu().numpy() != 0,
                              lambda m: f"  Outlier Filter: Removed {m} points ({m/n*100:.2f}%).")
    if dropped == 0:
        print("  Outlier Filter: No outliers detected.")
    return out


def remove_outliers_threshold(df, max_speed=10.0):
    """filtering.py:60-73 -- global cap on the velocity magnitude (host side: one pass over Np rows)."""
    uvw = df[["u", "v", "w"]].to_numpy()
    speed = np.sqrt(uvw[:, 0] ** 2 + uvw[:, 1] ** 2 + uvw[:, 2] ** 2)  # the reference's summation order
    out, _ = _drop_rows(df, speed <= max_speed,
                        lambda m: f"  Threshold Filter: Removed {m} points with speed > {max_speed}.")
    return out


def apply_filters(df, args):
    """filtering.py:75-89 -- what main.py:145-147 calls: nothing unless --filter-outliers, else the speed cap
    followed (if rows are left) by the kNN median/MAD filter, configured from the argparse namespace."""
    if args.filter_outliers:
        df = remove_outliers_threshold(df, max_speed=args.filter_max_speed)
        if len(df):
            df = remove_outliers_knn(df, k=args.filter_neighbors, threshold=args.filter_threshold)
    return df
