"""bench.py contract on CPU: the reference arm runs here (it is the reference's own CPU path) and must print exactly one
JSON line with the agreed keys; the weak-scaling workload keeps the per-rank geometry."""
import json
import os
import subprocess
import sys

import numpy as np

from tests.util import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "T0",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mvs_depth_pixels_per_second" and d["unit"] == "depth-pixels/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_weak_scaling_workload_keeps_block_geometry():
    """At N GPUs the scene is N copies of the 4x4 camera block; relative to its block centre every block has the cameras
    of the single-GPU scene (up to the random distance jitter) and rank r owns block r."""
    sys.path.insert(0, ROOT)
    import bench
    from mve_b200 import sharding, synth
    c1 = bench.workload_cfg("C2", 1)
    c3 = bench.workload_cfg("C2", 3)
    assert c3["views"] == 48 and c3["blocks"] == 3 and c3["features"] == 3 * c1["features"]
    small = dict(width=96, height=54)
    s1 = synth.make_scene(dict(c1, **small), only_views=[])
    s3 = synth.make_scene(dict(c3, **small), only_views=[])

    def centres(s):
        return np.stack([-(s.rot[v].reshape(3, 3).T @ s.trans[v]) for v in range(s.n_views)])
    a, b = centres(s1), centres(s3)
    for r in range(3):
        own = sharding.owned_views(48, r, 3)
        assert own == list(range(16 * r, 16 * r + 16))
        blk = b[own]
        off = blk[:, :2] - blk[:, :2].mean(0)
        assert np.allclose(off, a[:, :2] - a[:, :2].mean(0), atol=1e-5)
        # viewing directions (third row of R) match the single-block scene up to the jitter of the camera height
        assert np.allclose(s3.rot[own][:, 6:9], s1.rot[:, 6:9], atol=0.02)
    per_view = np.zeros(48, int)
    for refs in s3.feat_refs:
        per_view[refs] += 1
    assert per_view.min() > 0.5 * per_view.max()          # no view sees a disproportionate part of the scene
