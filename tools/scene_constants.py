"""Scene constants of SURVEY.md §8(d): the ORACLE's work counts (strict reference order) for a workload, which
define the algorithmic bytes of the roofline figure:
    bytes_alg = 300 * N_PSE + 75 * N_opt + 28 * N_filled
Test infrastructure (runs oracle/).  Writes profiles/scene_constants_<name>.json.
usage: python tools/scene_constants.py C2 [view ...]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mve_b200 import synth            # noqa: E402
from oracle import oracle_py as O     # noqa: E402


def main():
    name = sys.argv[1]
    views = [int(x) for x in sys.argv[2:]] or [0]
    s = synth.make_scene(name)
    sc = O.OracleScene(s)
    st = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
    res = {}

    def work(v):
        t = time.time()
        r = sc.reconstruct(st, v)
        res[v] = dict({k: int(r["stats"][k]) for k in r["stats"].dtype.names}, seconds=time.time() - t,
                      pixels=int(r["depth"].size))
    th = [threading.Thread(target=work, args=(v,)) for v in views]
    [t.start() for t in th]
    [t.join() for t in th]
    tot = {k: sum(res[v][k] for v in views) for k in ("n_opt", "n_pse_deriv", "n_pse_color", "n_filled", "pixels")}
    n_pse = tot["n_pse_deriv"] + tot["n_pse_color"]
    bytes_alg = 300 * n_pse + 75 * tot["n_opt"] + 28 * tot["n_filled"]
    out = dict(workload=name, views=views, per_view=res, totals=tot,
               pse_per_filled_px=n_pse / tot["n_filled"], opt_per_filled_px=tot["n_opt"] / tot["n_filled"],
               bytes_alg_per_filled_px=bytes_alg / tot["n_filled"], fill_ratio=tot["n_filled"] / tot["pixels"],
               cpu_px_per_s_per_core=tot["n_filled"] / sum(res[v]["seconds"] for v in views),
               definition="bytes_alg = 300*N_PSE + 75*N_opt + 28*N_filled with the oracle's strict-order counts (SURVEY.md 8d)")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "scene_constants_%s.json" % name), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "per_view"}, indent=1))


if __name__ == "__main__":
    main()
