"""Kernel-variant micro-benchmark: times DMRecon::start (HBM-resident) on a cached scene with the library named by
$B200MVS_LIB.  usage: kbench.py C2 [steps]   (scene cached in /tmp/kbench_<name>.npz)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mve_b200 import dmrecon, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cache = "/tmp/kbench_%s.npz" % name
if not os.path.exists(cache):
    synth.save_scene_npz(synth.make_scene(name, device="cuda"), cache)
s = synth.load_scene_npz(cache)
g = dmrecon.Scene.from_synth(s)
st = dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
best = None
for i in range(steps + 1):
    t = time.time()
    _, stats = g.reconstruct(st, list(range(s.n_views)), download=False)
    dt = time.time() - t
    d = stats.as_dict()
    d["wall_ms"] = 1e3 * dt
    if i > 0 and (best is None or d["ms_patch_kernel"] < best["ms_patch_kernel"]):
        best = d
print(json.dumps({"lib": os.environ.get("B200MVS_LIB", "default"), "kernel_ms": best["ms_patch_kernel"], "device_ms": best["ms_total_device"],
                  "wall_ms": best["wall_ms"], "n_opt": best["n_opt"], "n_filled": best["n_filled"], "sets": best["n_sample_sets"]}))
