"""Map-level parity of the GPU path against the reference CLI at BASELINE size, per frontier mode (runs on the GPU box):

    python tools/parity_modes.py C2 5 [out.json]

Reference: oracle/_ref/dmrecon (the unmodified reference, strict priority order) on the same scene directory.
Modes: default (every queued entry runs each round), frontier_band, frontier_topk (include/b200mvs.h).
Prints SURVEY 8c's figures and the device time of each mode."""
import json
import sys
import time

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

from mve_b200 import dmrecon, synth          # noqa: E402
from tests.util import map_parity, reference_cli_maps   # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    views = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "5").split(",")]
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    import torch
    s = synth.make_scene(name, device="cuda" if torch.cuda.is_available() else None)
    t = time.time()
    ref = reference_cli_maps(s, views)
    t_ref = time.time() - t
    g = dmrecon.Scene.from_synth(s)
    res = dict(workload=name, views=views, reference_cli_seconds=t_ref, modes={})
    modes = [("default", {}), ("band_0.01", dict(frontier_band=0.01)), ("band_0.003", dict(frontier_band=0.003)),
             ("topk_4096", dict(frontier_topk=4096)), ("topk_1024", dict(frontier_topk=1024)),
             ("topk_256", dict(frontier_topk=256)), ("topk_64", dict(frontier_topk=64))]
    for mname, kw in modes:
        st_ = dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors, **kw)
        g.reconstruct(st_, views)                                   # warm-up
        maps, st = g.reconstruct(st_, views)
        per = [map_parity(ref[v], maps[k]) for k, v in enumerate(views)]
        res["modes"][mname] = dict(parity=per, device_ms=st.ms_total_device, optimise_ms=st.ms_optimise_phases,
                                   rounds=int(st.n_rounds), n_opt=int(st.n_opt), n_filled=int(st.n_filled),
                                   barriers=int(st.n_grid_barriers))
        print(mname, json.dumps(res["modes"][mname]), flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
