"""The frontier ("wavefront") schedule the GPU runs vs the reference's strict priority order, both evaluated by the
CPU restatement: quantifies what the change of processing ORDER alone does to the maps (DESIGN.md "Frontier
schedule").  The tolerances are the map-level parity tolerances stated for the product."""
import numpy as np
import pytest

from oracle import oracle_py as O
from tests.util import golden_scene, map_stats


@pytest.mark.parametrize("name,view", [("T0", 0), ("T1", 4)])
def test_wavefront_close_to_strict(name, view):
    s = golden_scene(name)
    sc = O.OracleScene(s)
    st = O.default_settings(scale=s.scale, nr_recon_neighbors=s.nr_recon_neighbors)
    a = sc.reconstruct(st, view)
    b = sc.reconstruct_wavefront(st, view, 0.0)
    iou, rel, both = map_stats(a["depth"], b["depth"])
    assert iou > 0.99
    assert abs(int((a["depth"] > 0).sum()) - int((b["depth"] > 0).sum())) <= 0.01 * (a["depth"] > 0).sum()
    assert np.percentile(rel, 50) < 5e-4
    assert np.percentile(rel, 99) < 5e-3
    assert rel.max() < 2e-2
    assert np.percentile(np.abs(a["conf"] - b["conf"])[both], 99) < 2e-2
    # the schedule must not cost more optimisations than the strict order needs
    assert b["stats"]["n_opt"] <= 1.1 * a["stats"]["n_opt"]


def test_wavefront_is_deterministic():
    s = golden_scene("T0")
    sc = O.OracleScene(s)
    st = O.default_settings(scale=s.scale)
    a = sc.reconstruct_wavefront(st, 2, 0.0)
    b = sc.reconstruct_wavefront(st, 2, 0.0)
    for k in ("depth", "conf", "dz", "normal", "view_ids"):
        assert (a[k] == b[k]).all()
