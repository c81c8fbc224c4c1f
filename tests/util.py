"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def golden_scene(name):
    from mve_b200 import synth
    return synth.load_scene_npz(os.path.join(GOLD, "%s_scene.npz" % name))


def golden_ref(name):
    return np.load(os.path.join(GOLD, "%s_ref.npz" % name))


def map_stats(a_depth, b_depth):
    """Fill-mask IoU and relative depth difference percentiles on commonly filled pixels."""
    m1, m2 = a_depth > 0, b_depth > 0
    both = m1 & m2
    iou = both.sum() / max(1, (m1 | m2).sum())
    rel = np.abs(a_depth - b_depth)[both] / a_depth[both]
    return iou, rel, both


def patch_compare(got, ref):
    """Compares PatchOptimization results. Returns dict of mismatch counts / error percentiles."""
    ok_r, ok_g = ref["conf"] > 0, got["conf"] > 0
    both = ok_r & ok_g
    rel = np.abs(got["depth"] - ref["depth"])[both] / np.abs(ref["depth"][both])
    return dict(n=len(ref), ok_mismatch=int((ok_r != ok_g).sum()),
                ids_mismatch=int((got["local_ids"] != ref["local_ids"]).any(-1)[both].sum()),
                rel=rel, conf_abs=np.abs(got["conf"] - ref["conf"])[both],
                dz_abs=np.maximum(np.abs(got["dz_i"] - ref["dz_i"]), np.abs(got["dz_j"] - ref["dz_j"]))[both],
                nrm_abs=np.abs(got["normal"] - ref["normal"]).max(-1)[both], both=both)
