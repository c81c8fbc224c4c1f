"""Agreement of the per-pixel local view sets (runs on the GPU box):

    python tests/measure_view_ids.py C2 5 [out.json]

The reference CLI does not save which 4 neighbour views a pixel was reconstructed from, so the map-level agreement of the
local view ids is measured against the oracle's strict-priority-order run (oracle/mvs_oracle.cc, pinned against the
reference), for the default frontier mode and for frontier_topk = 64."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mve_b200 import dmrecon, synth          # noqa: E402
from oracle import oracle_py                 # noqa: E402  (test infrastructure: a measurement script under tests/: the oracle is test infrastructure)
from tests.util import map_parity            # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    view = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    s = synth.make_scene(name, device="cuda")
    o = oracle_py.OracleScene(s)
    t = time.time()
    ref = o.reconstruct(oracle_py.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors), view)
    res = dict(workload=name, view=view, oracle_seconds=time.time() - t, modes={})
    g = dmrecon.Scene.from_synth(s)
    for mname, kw in (("default", {}), ("topk_64", dict(frontier_topk=64))):
        maps, _ = g.reconstruct(dmrecon.Settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors, **kw), [view])
        res["modes"][mname] = map_parity(ref, maps[0])
        print(mname, json.dumps(res["modes"][mname]), flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
