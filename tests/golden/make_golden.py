"""Mints the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref, built by oracle/Makefile from
/root/reference).  Run in the build container only:   python tests/golden/make_golden.py

Outputs (all under tests/golden/):
  srgb2lin.npy            the 256-entry table of libs/dmrecon/mvs_tools.cc:30-95, parsed from the source
  <S>_scene.npz           the synthetic scene (images, cameras, features) so that fixtures are self-contained
  <S>_ref.npz             reference results for that scene:
      gvs_default / gvs_n3     "Global View Selection:" line of the reference per view (default and -n 3)
      patch_in / patch_out     inputs and mvs::PatchOptimization results through oracle/_ref/ref_harness
      depth_v / conf_v / dz_v  maps written by oracle/_ref/dmrecon for views v (apps/dmrecon CLI, unmodified)
      undist_v                 pyramid level `scale` written by the reference (scale != 0 only)
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mve_b200 import synth            # noqa: E402
from oracle import oracle_py as O     # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")


def parse_lut():
    src = open("/root/reference/libs/dmrecon/mvs_tools.cc").read()
    a = src.index("srgb2lin[256] = {")
    body = src[a:src.index("};", a)]
    vals = [np.float32(x.rstrip("f")) for x in re.findall(r"[0-9.]+(?:e-?[0-9]+)?f", body)]
    assert len(vals) == 256
    return np.asarray(vals, np.float32)


def run_cli(scene_dir, scale, nrn, extra=()):
    cmd = [os.path.join(REF, "dmrecon"), "-s%d" % scale, "--local-neighbors=%d" % nrn, "--keep-conf", "--keep-dz",
           "--progress=silent", "--force"] + list(extra) + [scene_dir]
    return subprocess.run(cmd, capture_output=True, text=True, check=True, env=dict(os.environ, OMP_NUM_THREADS="1")).stdout


def gvs_lines(scene_dir, scale, nrn, n_views, extra=()):
    out = {}
    for v in range(n_views):
        txt = subprocess.run([os.path.join(REF, "dmrecon"), "-s%d" % scale, "--local-neighbors=%d" % nrn,
                              "--progress=simple", "--force", "-l%d" % v] + list(extra) + [scene_dir],
                             capture_output=True, text=True).stdout
        m = re.search(r"Global View Selection:([ 0-9]*)", txt)
        out[v] = np.asarray([int(x) for x in m.group(1).split()], np.int32) if m else np.zeros(0, np.int32)
    return out


def mint_gvs_only(name):
    """Only the printed global view selections (many-candidate scene)."""
    s = synth.make_scene(name)
    synth.save_scene_npz(s, os.path.join(GOLD, "%s_scene.npz" % name))
    s = synth.load_scene_npz(os.path.join(GOLD, "%s_scene.npz" % name))
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        synth.write_mve_scene(s, tmp)
        data = {}
        for tag, extra in (("gvs_default", ()), ("gvs_n3", ("-n3",))):
            for v, ids in gvs_lines(tmp, s.scale, s.nr_recon_neighbors, s.n_views, extra).items():
                data["%s_%d" % (tag, v)] = ids
        np.savez_compressed(os.path.join(GOLD, "%s_ref.npz" % name), **data)
        print(name, "gvs only;", "view 0 ->", data["gvs_default_0"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def mint(name, map_views, n_patches=1500):
    s = synth.make_scene(name)
    synth.save_scene_npz(s, os.path.join(GOLD, "%s_scene.npz" % name))
    s = synth.load_scene_npz(os.path.join(GOLD, "%s_scene.npz" % name))
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        synth.write_mve_scene(s, tmp)
        data = {}
        for tag, extra in (("gvs_default", ()), ("gvs_n3", ("-n3",))):
            g = gvs_lines(tmp, s.scale, s.nr_recon_neighbors, s.n_views, extra)
            for v, ids in g.items():
                data["%s_%d" % (tag, v)] = ids
        run_cli(tmp, s.scale, s.nr_recon_neighbors)
        for v in map_views:
            vd = os.path.join(tmp, "views", "view_%04d.mve" % v)
            data["depth_%d" % v] = synth.read_mvei(os.path.join(vd, "depth-L%d.mvei" % s.scale))[:, :, 0]
            data["conf_%d" % v] = synth.read_mvei(os.path.join(vd, "conf-L%d.mvei" % s.scale))[:, :, 0]
            data["dz_%d" % v] = synth.read_mvei(os.path.join(vd, "dz-L%d.mvei" % s.scale))
            if s.scale:
                data["undist_%d" % v] = synth.read_mvei(os.path.join(vd, "undist-L%d.png" % s.scale))
        # patch-level vectors: realistic inputs = a slice of the oracle's own execution trace (seeds + queue)
        osc = O.OracleScene(s)
        st = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
        ref = map_views[0]
        r = osc.reconstruct(st, ref, trace_cap=200000)
        tin = r["trace_in"]
        seeds = np.nonzero(tin["n_local"] == 0)[0]
        rest = np.nonzero(tin["n_local"] != 0)[0]
        rng = np.random.default_rng(7)
        pick = np.sort(np.concatenate([seeds, rng.choice(rest, size=min(n_patches, len(rest)), replace=False)]))
        pin = np.ascontiguousarray(tin[pick])
        # a few hostile inputs: image border, negative depth slope, far-off depth
        extra = np.zeros(6, O.PATCH_IN)
        extra["local_ids"] = -1
        extra[0] = (1, 1, 5.0, 0, 0, 0, [-1] * 4)
        extra[1] = (s.width // (2 ** s.scale) - 2, 10, 5.0, 0, 0, 0, [-1] * 4)
        extra[2] = (40, 40, 5.0, -3.0, 0.0, 0, [-1] * 4)
        extra[3] = (40, 40, 50.0, 0, 0, 0, [-1] * 4)
        extra[4] = (40, 40, 0.5, 0, 0, 0, [-1] * 4)
        extra[5] = (40, 40, -1.0, 0, 0, 0, [-1] * 4)
        pin = np.concatenate([pin, extra])
        fin, fout = os.path.join(tmp, "pin.bin"), os.path.join(tmp, "pout.bin")
        pin.tofile(fin)
        txt = subprocess.run([os.path.join(REF, "ref_harness"), "patches", tmp, str(ref), str(s.scale),
                              str(s.nr_recon_neighbors), fin, fout], capture_output=True, text=True, check=True).stdout
        m = re.search(r"Global View Selection:([ 0-9]*)", txt)
        data["patch_gvs"] = np.asarray([int(x) for x in m.group(1).split()], np.int32)
        data["patch_ref_view"] = np.int32(ref)
        data["patch_in"] = pin
        data["patch_out"] = np.fromfile(fout, dtype=O.PATCH_OUT)
        np.savez_compressed(os.path.join(GOLD, "%s_ref.npz" % name), **data)
        print(name, "patches", len(pin), "maps", map_views)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    np.save(os.path.join(GOLD, "srgb2lin.npy"), parse_lut())
    mint("T0", [0, 3])
    mint("T1", [4])
    mint("T2", [0])
    mint("T4", [1])
    mint_gvs_only("T3")
