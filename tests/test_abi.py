"""The C-ABI library loads and exports every symbol include/b200mvs.h declares; struct layouts match; the product
fails loudly without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from tests.util import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "b200mvs.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mvs_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported():
    from mve_b200 import dmrecon
    L = dmrecon.lib()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(L, n), n
    assert sorted(dmrecon.EXPORTS) == names


def test_struct_layouts_match_header():
    from mve_b200 import dmrecon
    assert C.sizeof(dmrecon.Settings) == 4 * 10 + 4 * 6 + 4 + 4
    assert dmrecon.PATCH_IN.itemsize == 40 and dmrecon.PATCH_OUT.itemsize == 60
    assert C.sizeof(dmrecon.Progress) == 32
    assert C.sizeof(dmrecon.Stats) == 8 * 15


def test_default_settings_match_reference():
    """libs/dmrecon/settings.h:22-52."""
    from mve_b200 import dmrecon
    s = dmrecon.Settings()
    assert (s.filter_width, s.max_iterations, s.nr_recon_neighbors, s.global_vs_max, s.scale, s.use_color_scale) == (5, 20, 4, 20, 0, 1)
    assert abs(s.min_ncc - 0.3) < 1e-7 and abs(s.accept_ncc - 0.6) < 1e-7 and abs(s.min_refine_diff - 0.001) < 1e-9
    assert s.min_parallax == 10.0
    assert s.aabb_min[0] < -3e38 and s.aabb_max[2] > 3e38


def test_no_cpu_fallback():
    """Without a usable CUDA device creating a scene must raise, never silently compute on the CPU."""
    import torch
    from mve_b200 import dmrecon
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(dmrecon.B200MVSError) as e:
        dmrecon.Scene(2)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """Nothing under mve_b200/, include/ or shim/ may import, include or link oracle/ (shim/Makefile links the
    REFERENCE's own libmve.a / libmve_util.a, which oracle/Makefile compiles into oracle/_ref - that is the reference, not the
    oracle restatement)."""
    bad = []
    for base in ("mve_b200", "include", "shim"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) and "_build" not in d:
                    txt = open(os.path.join(d, f)).read()
                    if re.search(r"(import|from)\s+oracle|mvs_oracle|oracle_py|libmvs_oracle", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad
