"""Runs one BASELINE config (or a shard of it) on one GPU and prints throughput + counters.
usage: run_config.py C3 [first_ref n_refs]     (all views are uploaded, only the shard is reconstructed)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mve_b200 import dmrecon, synth  # noqa: E402

name = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nref = int(sys.argv[3]) if len(sys.argv) > 3 else 8
t = time.time()
s = synth.make_scene(name, device="cuda")
t_gen = time.time() - t
t = time.time()
g = dmrecon.Scene.from_synth(s)
t_up = time.time() - t
st = dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
refs = list(range(first, min(s.n_views, first + nref)))
out = None
for i in range(2):
    t = time.time()
    maps, stats = g.reconstruct(st, refs, download=(i == 1), want=("depth", "conf"))
    dt = time.time() - t
    d = stats.as_dict()
    out = dict(config=name, refs=refs, wall_s=dt, px_per_s=d["n_filled"] / dt, fill=d["n_filled"] / (len(refs) * maps[0]["depth"].size) if maps else None,
               gen_s=t_gen, upload_s=t_up, **d)
print(json.dumps(out))
if maps:
    depth = maps[0]["depth"]
    print("view %d: fill %.3f, depth range %.3f..%.3f, gvs %s" % (refs[0], (depth > 0).mean(), depth[depth > 0].min(), depth.max(),
                                                                  g.global_view_selection(st, refs[0])), file=sys.stderr)
