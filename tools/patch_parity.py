"""Patch-level parity figures of both device implementations of PatchOptimization (runs on the GPU box):
GPU vs the reference's own mvs::PatchOptimization results (golden, minted by oracle/_ref/ref_harness) on identical inputs.
Writes gpurun_out/patch_parity.json.  Test infrastructure."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mve_b200 import dmrecon                      # noqa: E402
from tests.util import golden_ref, golden_scene, patch_compare   # noqa: E402


def main():
    out = {}
    for name in ("T0", "T1", "T2", "T4"):
        s = golden_scene(name)
        ref = golden_ref(name)
        g = dmrecon.Scene.from_synth(s)
        gs = dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
        for mode, tag in ((1, "warp"), (2, "thread")):
            g.set_patch_mode(mode)
            got = g.optimize_patches(gs, int(ref["patch_ref_view"]), ref["patch_gvs"].tolist(), ref["patch_in"])
            c = patch_compare(got, ref["patch_out"])
            rel, cf, dz = c["rel"], c["conf_abs"], c["dz_abs"]
            out["%s_%s" % (name, tag)] = dict(
                n=int(c["n"]), both=int(c["both"].sum()), ok_mismatch=c["ok_mismatch"], ids_mismatch=c["ids_mismatch"],
                depth_rel_p99=float(np.percentile(rel, 99)), depth_rel_p999=float(np.percentile(rel, 99.9)), depth_rel_max=float(rel.max()),
                depth_rel_frac_le_1e5=float((rel <= 1e-5).mean()), depth_bit_identical=float((rel == 0).mean()),
                conf_abs_p99=float(np.percentile(cf, 99)), conf_abs_p999=float(np.percentile(cf, 99.9)),
                conf_frac_le_1e5=float((cf <= 1e-5).mean()),
                dz_abs_p99=float(np.percentile(dz, 99)), dz_abs_p999=float(np.percentile(dz, 99.9)), dz_frac_le_1e6=float((dz <= 1e-6).mean()))
            print(name, tag, json.dumps(out["%s_%s" % (name, tag)]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "patch_parity.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
