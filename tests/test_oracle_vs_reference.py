"""Pins the CPU restatement (oracle/mvs_oracle.cc) against outputs of the REFERENCE ITSELF.

The reference has no tests or golden vectors for libs/dmrecon, so the pins are fixtures minted
from the unmodified reference compiled into oracle/_ref (tests/golden/make_golden.py): the sRGB table, pyramid
levels, printed global view selections, per-patch mvs::PatchOptimization results and whole depth/conf/dz maps.

Float tolerances: the reference is built with -funsafe-math-optimizations and FMA contraction (Makefile.inc:4-5),
the restatement with strict IEEE evaluation; the reference differs from ITSELF by the same amounts across
compiler flags (depth rel p99 3.2e-4, max 3.1e-3 at map level).  Integer results must be equal.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle_py as O
from tests.util import GOLD, ROOT, golden_ref, golden_scene, map_stats, patch_compare


@pytest.fixture(scope="module")
def osc():
    cache = {}

    def get(name):
        if name not in cache:
            s = golden_scene(name)
            cache[name] = (s, O.OracleScene(s))
        return cache[name]
    return get


def test_srgb_table_matches_reference():
    """mvs_tools.cc:30-95: the formula used by oracle and product reproduces the 256 floats bit for bit."""
    ref = np.load(os.path.join(GOLD, "srgb2lin.npy"))
    i = np.arange(256)
    x = i / 255.0
    mine = np.where(i <= 0.04045 * 255.0, x / 12.92, ((x + 0.055) / 1.055) ** 2.4).astype(np.float32)
    assert (mine == ref).all()


def test_pyramid_bit_exact(osc):
    """rescale_half_size_gaussian<uint8>(img, 1.f) (image_tools.h:617-694): level `scale` as saved by the reference."""
    s, sc = osc("T1")
    ref = golden_ref("T1")
    assert (sc.level(4, s.scale) == ref["undist_4"]).all()
    # odd dimensions (161x121 -> 81x61): clamped taps at the right / bottom edge
    s, sc = osc("T4")
    assert (sc.level(1, s.scale) == golden_ref("T4")["undist_1"]).all()


@pytest.mark.parametrize("name", ["T0", "T1", "T2", "T3", "T4"])
def test_global_view_selection_exact(osc, name):
    """Integer result of GlobalViewSelection (global_view_selection.cc:34-101) for default and -n 3."""
    s, sc = osc(name)
    ref = golden_ref(name)
    for tag, gmax in (("gvs_default", 20), ("gvs_n3", 3)):
        st = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors, global_vs_max=gmax)
        for v in range(s.n_views):
            assert sc.global_view_selection(st, v) == ref["%s_%d" % (tag, v)].tolist(), (name, tag, v)


@pytest.mark.parametrize("name", ["T0", "T1", "T2", "T4"])
def test_patch_optimization_vs_reference(osc, name):
    """mvs::PatchOptimization through ref_harness: same inputs -> same view ids, floats within FP noise."""
    s, sc = osc(name)
    ref = golden_ref(name)
    st = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
    got = sc.optimize_patches(st, int(ref["patch_ref_view"]), ref["patch_gvs"].tolist(), ref["patch_in"])
    c = patch_compare(got, ref["patch_out"])
    n = c["n"]
    # success/failure and selected views are discrete decisions: allow 0.2 % threshold flips from FP noise
    assert c["ok_mismatch"] <= max(1, 0.002 * n), c["ok_mismatch"]
    assert c["ids_mismatch"] <= max(1, 0.002 * n), c["ids_mismatch"]
    assert np.percentile(c["rel"], 99) < 2e-5
    assert np.percentile(c["rel"], 99.9) < 1e-3
    assert np.percentile(c["conf_abs"], 99) < 1e-4
    assert np.percentile(c["dz_abs"], 99) < 1e-4


# T2 (orbit around a sphere, 60 degree view spacing, 24 % fill) is poorly conditioned on purpose: region growing
# amplifies FP noise through its thresholded decisions, so the reference's OWN maps move by this much when its
# compiler flags change.  T0/T1 are the well-conditioned cases.
MAP_TOL = {"T0": dict(iou=0.995, p99=2e-3, mx=2e-2, conf=5e-3, dz=5e-3),
           "T4": dict(iou=0.995, p99=2e-3, mx=2e-2, conf=5e-3, dz=5e-3),   # odd sizes: principal point moves per level
           "T1": dict(iou=0.995, p99=2e-3, mx=2e-2, conf=5e-3, dz=5e-3),
           "T2": dict(iou=0.98, p99=1e-2, mx=5e-2, conf=1e-1, dz=1e-2)}


@pytest.mark.parametrize("name,views", [("T0", [0, 3]), ("T1", [4]), ("T2", [0]), ("T4", [1])])
def test_maps_vs_reference_cli(osc, name, views):
    """Whole depth/conf/dz maps of the unmodified apps/dmrecon CLI vs the restatement in strict priority order."""
    s, sc = osc(name)
    ref = golden_ref(name)
    tol = MAP_TOL[name]
    st = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
    for v in views:
        r = sc.reconstruct(st, v)
        iou, rel, both = map_stats(ref["depth_%d" % v], r["depth"])
        assert iou > tol["iou"], iou
        assert np.percentile(rel, 50) < 1e-4
        assert np.percentile(rel, 99) < tol["p99"]
        assert rel.max() < tol["mx"]
        assert np.percentile(np.abs(ref["conf_%d" % v] - r["conf"])[both], 99) < tol["conf"]
        assert np.percentile(np.abs(ref["dz_%d" % v] - r["dz"])[both], 99) < tol["dz"]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_harness")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_patches_on_fresh_scene():
    """Live run of the compiled reference on a scene that is NOT in the fixtures."""
    from mve_b200 import synth
    s = synth.make_scene("T0", seed=77, features=200)
    sc = O.OracleScene(s)
    st = O.default_settings(scale=0, nr_recon_neighbors=4)
    r = sc.reconstruct(st, 1, trace_cap=3000)
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_mve_scene(s, tmp)
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        r["trace_in"].tofile(fin)
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "patches", tmp, "1", "0", "4", fin, fout],
                       check=True, capture_output=True)
        ref_out = np.fromfile(fout, dtype=O.PATCH_OUT)
    c = patch_compare(r["trace_out"], ref_out)
    assert c["ok_mismatch"] <= 3 and c["ids_mismatch"] <= 3
    assert np.percentile(c["rel"], 99) < 2e-5
