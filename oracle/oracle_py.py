"""TEST INFRASTRUCTURE - ctypes driver for oracle/libmvs_oracle.so (see mvs_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmvs_oracle.so")


class Settings(C.Structure):
    _fields_ = [("filter_width", C.c_uint32), ("min_ncc", C.c_float), ("min_parallax", C.c_float),
                ("accept_ncc", C.c_float), ("min_refine_diff", C.c_float), ("max_iterations", C.c_uint32),
                ("nr_recon_neighbors", C.c_uint32), ("global_vs_max", C.c_uint32), ("scale", C.c_int32),
                ("use_color_scale", C.c_int32), ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3)]


def default_settings(scale: int = 0, nr_recon_neighbors: int = 4, **kw) -> Settings:
    """Defaults of mvs::Settings (libs/dmrecon/settings.h:22-52)."""
    fmax = float(np.finfo(np.float32).max)
    s = Settings(5, 0.3, 10.0, 0.6, 0.001, 20, nr_recon_neighbors, 20, scale, 1,
                 (C.c_float * 3)(-fmax, -fmax, -fmax), (C.c_float * 3)(fmax, fmax, fmax))
    for k, v in kw.items():
        setattr(s, k, v)
    return s


PATCH_IN = np.dtype([("x", "<i4"), ("y", "<i4"), ("depth", "<f4"), ("dz_i", "<f4"), ("dz_j", "<f4"),
                     ("n_local", "<i4"), ("local_ids", "<i4", (4,))])
PATCH_OUT = np.dtype([("conf", "<f4"), ("depth", "<f4"), ("dz_i", "<f4"), ("dz_j", "<f4"),
                      ("normal", "<f4", (3,)), ("n_local", "<i4"), ("local_ids", "<i4", (4,)),
                      ("iterations", "<i4"), ("converged", "<i4"), ("opti_success", "<i4")])
STATS = np.dtype([(n, "<u8") for n in ("n_opt", "n_pse_deriv", "n_pse_color", "n_ncc", "n_update", "n_pops",
                                        "n_stale", "n_filled", "n_seeds_processed", "n_seeds_success",
                                        "n_spec_rounds", "n_spec_wasted")])


def build() -> str:
    """Compile the restatement (and, when /root/reference is present, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def _lib():
    if not os.path.exists(_LIB_PATH):
        build()
    lib = C.CDLL(_LIB_PATH)
    lib.mvs_oracle_create.restype = C.c_void_p
    lib.mvs_oracle_create.argtypes = [C.c_int]
    lib.mvs_oracle_destroy.argtypes = [C.c_void_p]
    lib.mvs_oracle_set_view.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvs_oracle_set_features.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvs_oracle_num_levels.argtypes = [C.c_void_p, C.c_int]
    lib.mvs_oracle_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mvs_oracle_get_level_calib.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.mvs_oracle_global_view_selection.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.mvs_oracle_optimize_patches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.mvs_oracle_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 8 + \
                                          [C.c_int64, C.c_void_p, C.c_double]
    lib.mvs_oracle_reconstruct_wavefront.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 6
    return lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleScene:
    def __init__(self, scene):
        """`scene` is a mve_b200.synth.Scene (or anything with the same fields)."""
        self.lib = _lib()
        self.n_views = scene.n_views
        self.h = self.lib.mvs_oracle_create(self.n_views)
        if not self.h:
            raise RuntimeError("mvs_oracle_create failed")
        for v in range(self.n_views):
            img = np.ascontiguousarray(scene.images[v], dtype=np.uint8)
            hh, ww = img.shape[:2]
            pp = np.ascontiguousarray(scene.ppoint[v], np.float32)
            rot = np.ascontiguousarray(scene.rot[v], np.float32)
            tr = np.ascontiguousarray(scene.trans[v], np.float32)
            rc = self.lib.mvs_oracle_set_view(self.h, v, _p(img), ww, hh, float(scene.flen[v]),
                                              float(scene.paspect[v]), _p(pp), _p(rot), _p(tr))
            if rc != 0:
                raise RuntimeError("mvs_oracle_set_view failed")
        off = np.zeros(len(scene.feat_refs) + 1, np.int32)
        off[1:] = np.cumsum([len(r) for r in scene.feat_refs])
        ids = np.concatenate(scene.feat_refs).astype(np.int32) if len(scene.feat_refs) else np.zeros(0, np.int32)
        pos = np.ascontiguousarray(scene.feat_pos, np.float32)
        self.lib.mvs_oracle_set_features(self.h, len(scene.feat_refs), _p(pos), _p(off), _p(ids))

    def __del__(self):
        try:
            if self.h:
                self.lib.mvs_oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def num_levels(self, view: int) -> int:
        return self.lib.mvs_oracle_num_levels(self.h, view)

    def level(self, view: int, level: int) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        if self.lib.mvs_oracle_get_level(self.h, view, level, C.byref(w), C.byref(h), None) != 0:
            raise IndexError(level)
        out = np.empty((h.value, w.value, 3), np.uint8)
        self.lib.mvs_oracle_get_level(self.h, view, level, C.byref(w), C.byref(h), _p(out))
        return out

    def level_calib(self, view: int, level: int):
        k, ki = np.empty(9, np.float32), np.empty(9, np.float32)
        self.lib.mvs_oracle_get_level_calib(self.h, view, level, _p(k), _p(ki))
        return k, ki

    def global_view_selection(self, st: Settings, ref: int) -> List[int]:
        out = np.empty(512, np.int32)
        n = self.lib.mvs_oracle_global_view_selection(self.h, C.byref(st), ref, _p(out), 512)
        return out[:n].tolist()

    def optimize_patches(self, st: Settings, ref: int, global_ids: Sequence[int], patches: np.ndarray,
                         stats: Optional[np.ndarray] = None) -> np.ndarray:
        patches = np.ascontiguousarray(patches, dtype=PATCH_IN)
        out = np.zeros(len(patches), PATCH_OUT)
        g = np.asarray(global_ids, np.int32)
        rc = self.lib.mvs_oracle_optimize_patches(self.h, C.byref(st), ref, _p(g), len(g), _p(patches),
                                                  len(patches), _p(out), _p(stats))
        if rc != 0:
            raise RuntimeError("mvs_oracle_optimize_patches rc=%d" % rc)
        return out

    def reconstruct(self, st: Settings, ref: int, trace_cap: int = 0, max_seconds: float = 0.0):
        w, h = C.c_int(), C.c_int()
        self.lib.mvs_oracle_get_level(self.h, ref, st.scale, C.byref(w), C.byref(h), None)
        W, H = w.value, h.value
        depth = np.zeros((H, W), np.float32)
        conf = np.zeros((H, W), np.float32)
        dz = np.zeros((H, W, 2), np.float32)
        normal = np.zeros((H, W, 3), np.float32)
        vids = np.zeros((H, W, 4), np.int32)
        stats = np.zeros(1, STATS)
        tin = np.zeros(trace_cap, PATCH_IN) if trace_cap else None
        tout = np.zeros(trace_cap, PATCH_OUT) if trace_cap else None
        tn = C.c_int64(0)
        rc = self.lib.mvs_oracle_reconstruct(self.h, C.byref(st), ref, _p(depth), _p(conf), _p(dz), _p(normal),
                                             _p(vids), _p(stats), _p(tin), _p(tout), trace_cap, C.byref(tn),
                                             max_seconds)
        if rc != 0:
            raise RuntimeError("mvs_oracle_reconstruct rc=%d" % rc)
        res = dict(depth=depth, conf=conf, dz=dz, normal=normal, view_ids=vids, stats=stats[0], n_trace=tn.value)
        if trace_cap:
            n = min(trace_cap, tn.value)
            res["trace_in"], res["trace_out"] = tin[:n], tout[:n]
        return res

    def reconstruct_wavefront(self, st: Settings, ref: int, band: float = 0.0, topk: int = 0):
        w, h = C.c_int(), C.c_int()
        self.lib.mvs_oracle_get_level(self.h, ref, st.scale, C.byref(w), C.byref(h), None)
        W, H = w.value, h.value
        depth = np.zeros((H, W), np.float32)
        conf = np.zeros((H, W), np.float32)
        dz = np.zeros((H, W, 2), np.float32)
        normal = np.zeros((H, W, 3), np.float32)
        vids = np.zeros((H, W, 4), np.int32)
        stats = np.zeros(1, STATS)
        rc = self.lib.mvs_oracle_reconstruct_wavefront(self.h, C.byref(st), ref, band, int(topk), _p(depth), _p(conf), _p(dz),
                                                       _p(normal), _p(vids), _p(stats))
        if rc != 0:
            raise RuntimeError("mvs_oracle_reconstruct_wavefront rc=%d" % rc)
        return dict(depth=depth, conf=conf, dz=dz, normal=normal, view_ids=vids, stats=stats[0])
