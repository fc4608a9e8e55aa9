"""BASELINE config 3 end to end on one GPU: deterministic primary projection of a 512^3 label phantom
onto a 1024x768 detector for 720 views, then FDK (TEXTBOOK weights) back to 512^3.  Host buffers
through the C ABI; prints the two wall times and the reconstructed attenuation of water."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monte_b200 import _abi, api, scenes  # noqa: E402

api.init(0)
n, pitch = 512, 0.05
lab = scenes.cylinder_phantom(n, pitch)
vol = scenes.volume_for(lab, pitch)
mg = scenes.mc_geom(0, 32.5 / 1024, n_views=720, ny=1024, nx=768)
mg.angle_step_deg = 0.5
xs = scenes.make_xs()
import torch
line_t = torch.zeros((720, 1024, 768), dtype=torch.float32).pin_memory()
lab_t = torch.from_numpy(lab).pin_memory()
api.project_primary(mg, vol, lab_t.numpy(), xs, 140.0, views=(0, 8), out=line_t.numpy())     # warm-up
t = time.perf_counter()
line = api.project_primary(mg, vol, lab_t.numpy(), xs, 140.0, out=line_t.numpy())
t_proj = time.perf_counter() - t
g = _abi.generic_fdk_geom(720, 1024, 768, 512, textbook=True)
g.du = g.dv = mg.pixel
g.half_u, g.half_v = mg.half, 0.5 * 768 * mg.pixel
mg_half_v = g.half_v
rec_t = torch.empty((512, 512, 512), dtype=torch.float32).pin_memory()
api.fdk(g, line, want_filtered=False, out=rec_t.numpy())                                       # warm-up (allocations)
t = time.perf_counter()
_, rec, _, st = api.fdk(g, line, want_filtered=False, out=rec_t.numpy())
t_fdk = time.perf_counter() - t
c = n // 2
zz, tt, ss = np.ogrid[:n, :n, :n]
r = np.hypot((ss - c + 0.5) * pitch, (tt - c + 0.5) * pitch)
water = (np.abs(zz - c) < 20) & (r > 7.5) & (r < 9.0)
mu = float(rec[np.broadcast_to(water, rec.shape)].mean())
print(json.dumps({"project_s": t_proj, "rays": 720 * 1024 * 768, "fdk_s": t_fdk, "fdk_stats_ms": st["ms_total"],
                  "mu_water_reconstructed": mu, "mu_water_table": float(xs.total[0][140])}))
