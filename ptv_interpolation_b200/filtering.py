"""Drop-in mirror of the reference's ``filtering.py``: global speed threshold and
the kNN median/MAD outlier filter, the latter on the CUDA path (self-query of the spatial hash with
k+1 neighbours + per-particle median / MAD of the neighbour speeds in the same kernel)."""
from __future__ import annotations

import numpy as np

from .engine import default_engine

__all__ = ["remove_outliers_knn", "remove_outliers_threshold", "apply_filters"]


def remove_outliers_knn(df, k=25, threshold=3.0, device=None):
    """filtering.py:5-58."""
    import torch
    if len(df) <= k:
        print(f"  Warning: DataFrame too small ({len(df)}) for k-NN filter (k={k}). Skipping.")
        return df
    eng = default_engine(device)
    pts = torch.from_numpy(np.ascontiguousarray(df[["x", "y", "z"]].values, dtype=np.float64)).to(eng.device)
    vals = torch.from_numpy(np.ascontiguousarray(df[["u", "v", "w"]].values, dtype=np.float64)).to(eng.device)
    eng.build(pts, vals)
    keep, kth = eng.outlier_filter(k=k, threshold=threshold)
    keep_mask = keep.cpu().numpy().astype(bool)
    median_filter_radius = float(np.median(kth.cpu().numpy()))
    print(f"  Filtering radius: median voxel distance to {k}-th neighbor = {median_filter_radius:.4f}")
    n_removed = int(np.sum(~keep_mask))
    if n_removed > 0:
        print(f"  Outlier Filter: Removed {n_removed} points ({n_removed/len(df)*100:.2f}%).")
        return df[keep_mask].reset_index(drop=True)
    print("  Outlier Filter: No outliers detected.")
    return df


def remove_outliers_threshold(df, max_speed=10.0):
    """filtering.py:60-73 (host; O(Np) elementwise)."""
    u, v, w = df["u"].values, df["v"].values, df["w"].values
    speed = np.sqrt(u**2 + v**2 + w**2)
    keep_mask = speed <= max_speed
    n_removed = np.sum(~keep_mask)
    if n_removed > 0:
        print(f"  Threshold Filter: Removed {n_removed} points with speed > {max_speed}.")
        return df[keep_mask].reset_index(drop=True)
    return df


def apply_filters(df, args):
    """filtering.py:75-89."""
    if not args.filter_outliers:
        return df
    df = remove_outliers_threshold(df, max_speed=args.filter_max_speed)
    if len(df) > 0:
        df = remove_outliers_knn(df, k=args.filter_neighbors, threshold=args.filter_threshold)
    return df
