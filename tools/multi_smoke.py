import sys, time, numpy as np
sys.path.insert(0, '.')
from pixelforge_b200 import load_product_scenes
t0=time.time()
p = load_product_scenes()
c, d, r = p.render("gears", 400, 300)
print("rendered", r.pixels_shaded, r.triangles_submitted, (c & 0xffffff != 0).sum(), "in", round(time.time()-t0,2), "s", flush=True)
np.save("/tmp/multi_out.npy", c)
