"""Depth-map consumers on the device (-m gpu) against the REFERENCE's own functions (libs/mve/depthmap.cc) run live through
oracle/_ref/ref_harness dmops on the same buffers: depthmap_confidence_clean, depthmap_cleanup (bit-exact) and
depthmap_triangulate - vertex ids, faces and vertex count exact, vertices / colours exact up to the reference binary's own
-funsafe-math contraction (<= 1e-6 relative)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests.util import ROOT, golden_ref, golden_scene

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
needs_ref = pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/ref_harness not built")


def _depth_case(kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "golden":
        ref = golden_ref("T0")
        return np.ascontiguousarray(ref["depth_0"], np.float32), np.ascontiguousarray(ref["conf_0"], np.float32)
    h, w = (97, 131) if kind == "ragged" else (270, 480)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d = (5.0 + 0.4 * np.sin(xx / 17.0) + 0.3 * np.cos(yy / 11.0)).astype(np.float32)
    d[(xx > w * 0.6) & (yy > h * 0.3)] += 1.5                      # a depth discontinuity
    hole = rng.random((h, w)) < (0.45 if kind == "ragged" else 0.08)   # ragged: many small islands
    d[hole] = 0.0
    d[:, :3] = 0.0
    conf = rng.random((h, w)).astype(np.float32) - 0.2
    return d, conf


@needs_ref
@pytest.mark.parametrize("kind", ["golden", "ragged", "large"])
def test_confidence_clean_and_cleanup_bit_exact(kind):
    from mve_b200 import depthmap as D
    dm, cm = _depth_case(kind)
    h, w = dm.shape
    with tempfile.TemporaryDirectory() as tmp:
        dm.tofile(os.path.join(tmp, "dm.f32")); cm.tofile(os.path.join(tmp, "cm.f32"))
        subprocess.run([HARNESS, "dmops", "confclean", str(w), str(h), os.path.join(tmp, "dm.f32"), os.path.join(tmp, "cm.f32"),
                        os.path.join(tmp, "cc.f32")], check=True)
        want = np.fromfile(os.path.join(tmp, "cc.f32"), np.float32).reshape(h, w)
        got = dm.copy()
        D.depthmap_confidence_clean(got, cm)
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
        for thres in (1, 7, 50, 2000):
            subprocess.run([HARNESS, "dmops", "cleanup", str(w), str(h), str(thres), os.path.join(tmp, "dm.f32"),
                            os.path.join(tmp, "cl.f32")], check=True)
            want = np.fromfile(os.path.join(tmp, "cl.f32"), np.float32).reshape(h, w)
            got = D.depthmap_cleanup(dm, thres)
            assert (got.view(np.uint32) == want.view(np.uint32)).all(), thres
    # empty and full maps
    z = np.zeros((5, 7), np.float32)
    assert (D.depthmap_cleanup(z, 3) == 0).all()
    o = np.ones((5, 7), np.float32)
    assert (D.depthmap_cleanup(o, 35) == 1).all() and (D.depthmap_cleanup(o, 36) == 0).all()


@needs_ref
@pytest.mark.parametrize("kind,dd,color", [("golden", 5.0, True), ("ragged", 5.0, False), ("large", 0.0, True), ("large", 2.0, False)])
def test_triangulate_matches_reference(kind, dd, color):
    from mve_b200 import depthmap as D
    dm, _ = _depth_case(kind, seed=3)
    h, w = dm.shape
    ax = float(max(w, h))
    invproj = np.array([1 / ax, 0, -0.5 * w / ax, 0, 1 / ax, -0.5 * h / ax, 0, 0, 1], np.float32)
    ci = None
    if color:
        ci = np.random.default_rng(1).integers(0, 255, size=(h, w, 3), dtype=np.uint8)
    with tempfile.TemporaryDirectory() as tmp:
        dm.tofile(os.path.join(tmp, "dm.f32"))
        cpath = "-"
        if ci is not None:
            cpath = os.path.join(tmp, "ci.u8"); ci.tofile(cpath)
        subprocess.run([HARNESS, "dmops", "triangulate", str(w), str(h), repr(dd), os.path.join(tmp, "dm.f32"), cpath, "3"] +
                       [repr(float(v)) for v in invproj] + [os.path.join(tmp, "out")], check=True)
        vids = np.fromfile(os.path.join(tmp, "out.vids"), np.uint32).reshape(h, w)
        verts = np.fromfile(os.path.join(tmp, "out.verts"), np.float32).reshape(-1, 3)
        faces = np.fromfile(os.path.join(tmp, "out.faces"), np.uint32).reshape(-1, 3)
        cols = np.fromfile(os.path.join(tmp, "out.colors"), np.float32)
        cols = cols.reshape(-1, 4) if cols.size else np.zeros((0, 4), np.float32)
        nrm = np.fromfile(os.path.join(tmp, "out.normals"), np.float32).reshape(-1, 3)
        cfs = np.fromfile(os.path.join(tmp, "out.confs"), np.float32)
        scl = np.fromfile(os.path.join(tmp, "out.scales"), np.float32)
    got = D.depthmap_triangulate(dm, invproj, dd_factor=dd, color=ci)
    assert len(verts) > 100 and len(faces) > 100
    # the rest of scene2pset's per-view work: vertex normals (angle-weighted), boundary confidences (exact: ring / 4), scale values
    ps = D.depthmap_pointset(dm, invproj, dd_factor=dd, color=ci, with_normals=True, conf_iterations=4, scale_factor=2.5)
    assert (ps["faces"] == faces).all() and (ps["vertex_ids"] == vids).all()
    assert (ps["confidences"] == cfs).all()
    assert set(np.unique(cfs)).issubset({0.0, 0.25, 0.5, 0.75, 1.0}) and (cfs == 0).any()
    assert kind == "ragged" or (cfs == 1).any()        # a ragged map may have no vertex further than 4 rings from a boundary
    dn = np.abs(ps["normals"] - nrm).max(-1)
    # angle weights are acos() of float dot products (device acosf vs the host's libm): measured p99.9 3.4e-5, max 6.7e-5
    assert np.percentile(dn, 99.9) <= 1e-4 and dn.max() <= 2e-3, (np.percentile(dn, 99.9), dn.max())
    ds = np.abs(ps["scales"] - scl) / np.abs(scl).max()
    assert ds.max() <= 3e-5, ds.max()          # float sums over <= 9 neighbours, the reference binary contracts to FMAs
    assert (got["vertex_ids"] == vids).all()
    assert got["faces"].shape == faces.shape and (got["faces"] == faces).all()
    assert got["vertices"].shape == verts.shape
    assert np.abs(got["vertices"] - verts).max() <= 1e-6 * np.abs(verts).max()
    if color:
        assert np.abs(got["colors"] - cols).max() <= 1e-6
    # world transform = the reference's mesh_transform of the same vertices
    ctw = np.eye(4, dtype=np.float32)
    ctw[:3, :3] = np.array([[0.36, 0.48, -0.8], [-0.8, 0.6, 0.0], [0.48, 0.64, 0.6]], np.float32)
    ctw[:3, 3] = [1.5, -2.0, 0.25]
    gw = D.depthmap_triangulate(dm, invproj, dd_factor=dd, cam_to_world=ctw)
    want = verts @ ctw[:3, :3].T + ctw[:3, 3]
    assert np.abs(gw["vertices"] - want).max() <= 2e-6 * np.abs(want).max()
    assert (gw["faces"] == faces).all()
