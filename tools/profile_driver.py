"""Small driver for ncu captures: one DMRecon::start batch on a workload. usage: profile_driver.py [C2] [n_steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mve_b200 import dmrecon, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s = synth.make_scene(name, device="cuda")
g = dmrecon.Scene.from_synth(s)
st = dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
for i in range(steps):
    t = time.time()
    _, stats = g.reconstruct(st, list(range(s.n_views)), download=False)
    print("step", i, "%.3fs" % (time.time() - t), stats.as_dict(), flush=True)
