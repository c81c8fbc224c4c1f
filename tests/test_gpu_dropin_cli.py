"""The drop-in itself (-m gpu): the reference's UNMODIFIED CLI driver (apps/dmrecon/dmrecon.cc, compiled from the
reference tree) linked against our mvs::DMRecon shim + libb200mvs.so instead of libmve_dmrecon.a, run on an MVE scene
directory, must leave the same embeddings on disk as the reference binary did (golden maps minted by
tests/golden/make_golden.py from oracle/_ref/dmrecon).  Skipped when shim/_build/dmrecon_b200 was not built."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests.util import ROOT, golden_ref, golden_scene, map_stats

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "shim", "_build", "dmrecon_b200")


@pytest.mark.skipif(not os.path.exists(CLI), reason="shim/_build/dmrecon_b200 not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name,views", [("T0", [0, 3]), ("T1", [4])])
def test_cli_writes_reference_layout(name, views):
    from mve_b200 import synth
    s = golden_scene(name)
    ref = golden_ref(name)
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_mve_scene(s, tmp)
        cmd = [CLI, "-s%d" % s.scale, "--local-neighbors=%d" % s.nr_recon_neighbors, "--keep-conf", "--keep-dz",
               "--progress=silent", "--force", "-l" + ",".join(str(v) for v in views), tmp]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        for v in views:
            vd = os.path.join(tmp, "views", "view_%04d.mve" % v)
            depth = synth.read_mvei(os.path.join(vd, "depth-L%d.mvei" % s.scale))
            conf = synth.read_mvei(os.path.join(vd, "conf-L%d.mvei" % s.scale))
            dz = synth.read_mvei(os.path.join(vd, "dz-L%d.mvei" % s.scale))
            assert depth.shape[2] == 1 and conf.shape[2] == 1 and dz.shape[2] == 2 and depth.dtype == np.float32
            iou, rel, both = map_stats(ref["depth_%d" % v], depth[:, :, 0])
            assert iou > 0.99
            assert np.percentile(rel, 50) < 5e-4 and np.percentile(rel, 99) < 5e-3
            assert np.percentile(np.abs(ref["conf_%d" % v] - conf[:, :, 0])[both], 99) < 2e-2
            if s.scale:
                und = [f for f in os.listdir(vd) if f.startswith("undist-L%d" % s.scale)]
                assert und, "undist-L<s> must be saved for scale != 0 (dmrecon.cc:138-143)"
                assert (synth.read_mvei(os.path.join(vd, und[0])) == ref["undist_%d" % v]).all()
        # resume semantics live in the unchanged driver: a second run without --force skips finished views
        out2 = subprocess.run([c for c in cmd if c != "--force"], capture_output=True, text=True, timeout=600)
        assert out2.returncode == 0
