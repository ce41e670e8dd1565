import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pathfinder_b200 import api, scenes
flat = scenes.random_paths(100000, 8192, 0x5EED0004); size = 8192
r = api.CudaRenderer((size, size), background_color=(1, 1, 1, 1)); r.set_debug_lists_enabled(True)
scene = api.Scene.from_flat(flat); scene.build_and_render(r, api.BuildOptions())
tiles = r.debug_tiles(); z, rect = r.debug_z_buffer()
w = rect[2] - rect[0]
fb = tiles["tile_y"].astype(np.int64) * w + tiles["tile_x"].astype(np.int64)
live = tiles["path_id"].astype(np.int64) >= z.reshape(-1)[fb]
t = tiles[live]; fb = fb[live]
alpha = t["alpha_tile_id"] != 0xFFFFFFFF
n = np.bincount(fb, minlength=z.size); na = np.bincount(fb, weights=alpha, minlength=z.size).astype(np.int64)
print("fb tiles", z.size, "entries", len(t), "alpha entries", int(alpha.sum()))
print("tiles with n=0:", int((n == 0).sum()), " uniform (n>0, no alpha):", int(((n > 0) & (na == 0)).sum()), " with alpha:", int((na > 0).sum()))
print("n histogram:", np.bincount(np.minimum(n, 12))[:13].tolist())
print("alpha-per-tile histogram:", np.bincount(np.minimum(na, 8))[:9].tolist())
