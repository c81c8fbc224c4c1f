"""Diagnostics on the GPU box: GPU path vs oracle at pyramid / patch / map level, with timings.
Writes gpurun_out/gpu_check_<scene>.json. Test infrastructure (uses oracle/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mve_b200 import dmrecon, synth  # noqa: E402
from oracle import oracle_py as O    # noqa: E402


def cmp_maps(a, b):
    m1, m2 = a["depth"] > 0, b["depth"] > 0
    both = m1 & m2
    rel = np.abs(a["depth"] - b["depth"])[both] / a["depth"][both]
    pct = np.percentile(rel, [50, 90, 99, 99.9, 100]) if both.any() else [0] * 5
    return dict(fill_a=int(m1.sum()), fill_b=int(m2.sum()), iou=float(both.sum() / max(1, (m1 | m2).sum())),
                bit_identical=float((a["depth"][both] == b["depth"][both]).mean()) if both.any() else 0.0,
                rel_p50=float(pct[0]), rel_p90=float(pct[1]), rel_p99=float(pct[2]), rel_p999=float(pct[3]), rel_max=float(pct[4]),
                conf_p99=float(np.percentile(np.abs(a["conf"] - b["conf"])[both], 99)) if both.any() else 0.0,
                dz_p99=float(np.percentile(np.abs(a["dz"] - b["dz"])[both], 99)) if both.any() else 0.0,
                ids_same=float((a["view_ids"] == b["view_ids"]).all(-1)[both].mean()) if both.any() else 0.0)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "T0"
    nref = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = {"scene": name}
    s = synth.make_scene(name)
    osc = O.OracleScene(s)
    t = time.time()
    gsc = dmrecon.Scene.from_synth(s)
    out["upload_s"] = time.time() - t
    ost = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
    gst = dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
    # pyramid
    mism = 0
    for v in range(s.n_views):
        assert gsc.num_levels(v) == osc.num_levels(v)
        for l in range(osc.num_levels(v)):
            mism += int((gsc.level(v, l) != osc.level(v, l)).sum())
    out["pyramid_mismatch_bytes"] = mism
    # global view selection
    out["gvs_equal"] = all(gsc.global_view_selection(gst, r) == osc.global_view_selection(ost, r) for r in range(s.n_views))
    for ref in range(min(nref, s.n_views)):
        g = osc.global_view_selection(ost, ref)
        r0 = osc.reconstruct(ost, ref, trace_cap=60000)
        tin, tout = r0["trace_in"], r0["trace_out"]
        st = dmrecon.Stats()
        t = time.time()
        gout = gsc.optimize_patches(gst, ref, g, tin, st)
        dt = time.time() - t
        ok_o, ok_g = tout["conf"] > 0, gout["conf"] > 0
        both = ok_o & ok_g
        rel = np.abs(gout["depth"] - tout["depth"])[both] / tout["depth"][both]
        pd = dict(n=len(tin), n_seed=int((tin["n_local"] == 0).sum()), ok_oracle=int(ok_o.sum()), ok_gpu=int(ok_g.sum()),
                  ok_mismatch=int((ok_o != ok_g).sum()),
                  ids_mismatch=int((gout["local_ids"] != tout["local_ids"]).any(-1)[both].sum()),
                  iter_mismatch=int((gout["iterations"] != tout["iterations"])[both].sum()),
                  depth_bit_identical=float((gout["depth"] == tout["depth"])[both].mean()),
                  depth_rel_p50=float(np.percentile(rel, 50)), depth_rel_p99=float(np.percentile(rel, 99)),
                  depth_rel_p999=float(np.percentile(rel, 99.9)), depth_rel_max=float(rel.max()),
                  frac_rel_gt_1e5=float((rel > 1e-5).mean()), frac_rel_gt_1e4=float((rel > 1e-4).mean()),
                  conf_abs_p99=float(np.percentile(np.abs(gout["conf"] - tout["conf"])[both], 99)),
                  conf_abs_max=float(np.abs(gout["conf"] - tout["conf"])[both].max()),
                  dz_abs_p99=float(np.percentile(np.abs(gout["dz_i"] - tout["dz_i"])[both], 99)),
                  normal_abs_max=float(np.abs(gout["normal"] - tout["normal"])[both].max()),
                  gpu_call_s=dt, gpu_kernel_ms=st.ms_patch_kernel, gpu_sets=int(st.n_sample_sets),
                  oracle_pse=int(r0["stats"]["n_pse_deriv"] + r0["stats"]["n_pse_color"]), oracle_opts=int(r0["stats"]["n_opt"]))
        out["patch_ref%d" % ref] = pd
        # maps
        t = time.time()
        maps, st2 = gsc.reconstruct(gst, [ref])
        dt = time.time() - t
        rw = osc.reconstruct_wavefront(ost, ref, 0.0)
        out["map_ref%d" % ref] = dict(gpu_vs_wavefront=cmp_maps(rw, maps[0]), gpu_vs_strict=cmp_maps(r0, maps[0]),
                                      wavefront_vs_strict=cmp_maps(r0, rw), gpu_s=dt, stats=st2.as_dict(),
                                      oracle_wavefront_opts=int(rw["stats"]["n_opt"]), oracle_wavefront_rounds=int(rw["stats"]["n_spec_rounds"]))
    # all views in one batch
    t = time.time()
    maps, st3 = gsc.reconstruct(gst, list(range(s.n_views)))
    out["batch_all"] = dict(s=time.time() - t, stats=st3.as_dict())
    t = time.time()
    maps, st3 = gsc.reconstruct(gst, list(range(s.n_views)), download=False)
    out["batch_all_resident"] = dict(s=time.time() - t, stats=st3.as_dict())
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check_%s.json" % name), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
