import sys; sys.path.insert(0,'.')
import numpy as np, pandas as pd, torch
from ptv_interpolation_b200 import synthetic, interpolator as gi
from ptv_interpolation_b200.engine import set_tuning, default_engine
n=48
mask = synthetic.hex6_sphere_pack_mask(n)
pts = synthetic.sample_pore_particles(mask, 12000, seed=5)
vals = synthetic.sphere_pack_flow(pts, n)
pts, vals, mask = pts.numpy(), vals.numpy(), mask.numpy()
grid,_ = gi.create_grid(((0,n),(0,n),(0,n)), n)
df = pd.DataFrame({"x": pts[:,0],"y":pts[:,1],"z":pts[:,2],"u":vals[:,0],"v":vals[:,1],"w":vals[:,2]})
for stream in (2,1):
    set_tuning(stream=stream, stats=1)
    gi.interpolate_field(df, grid, mask=mask, method="idw", out_dtype=np.float64)
    print(stream, int(mask.sum()), default_engine().knn_stats())
