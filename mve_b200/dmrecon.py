"""Host-side mirror of the reference's dmrecon interface over the C ABI of libb200mvs.so.

The reference's boundary for this path is the C++ class `mvs::DMRecon(scene, settings).start()`
(libs/dmrecon/dmrecon.h:40-68) fed by `mve::Scene`/`mve::View`; the compiled drop-in for that is
shim/ (see INTEGRATION.md).  This module is the same surface for Python callers (tests, bench.py):

    Settings  <-> mvs::Settings           (libs/dmrecon/settings.h:22-52, same names and defaults)
    Scene     <-> mve::Scene              (views with `undistorted` images + cameras, bundle features)
    DMRecon   <-> mvs::DMRecon            (ctor validation messages of dmrecon.cc:30-87; start())

There is no CPU fallback: importing works without a GPU, creating a Scene does not.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import build as _build

_LIB = None


class B200MVSError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("b200mvs error %d: %s" % (code, msg))
        self.code = code


class Settings(C.Structure):
    """mvs::Settings (settings.h:22-52). Field names follow the C ABI (snake case of the reference's)."""
    _fields_ = [("filter_width", C.c_uint32), ("min_ncc", C.c_float), ("min_parallax", C.c_float),
                ("accept_ncc", C.c_float), ("min_refine_diff", C.c_float), ("max_iterations", C.c_uint32),
                ("nr_recon_neighbors", C.c_uint32), ("global_vs_max", C.c_uint32), ("scale", C.c_int32),
                ("use_color_scale", C.c_int32), ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3),
                ("frontier_band", C.c_float), ("frontier_topk", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__()
        lib().b200mvs_default_settings(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class _Maps(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("conf", C.c_void_p), ("dz", C.c_void_p), ("normal", C.c_void_p),
                ("view_ids", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32)]


class Progress(C.Structure):
    """mvs::Progress (progress.h:27-43)."""
    _fields_ = [("status", C.c_int32), ("cancelled", C.c_int32), ("filled", C.c_uint64),
                ("queue_size", C.c_uint64), ("start_time", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("n_opt", C.c_uint64), ("n_sample_sets", C.c_uint64), ("n_rounds", C.c_uint64),
                ("n_filled", C.c_uint64), ("n_seeds_processed", C.c_uint64), ("n_seeds_success", C.c_uint64),
                ("n_entries_peak", C.c_uint64), ("ms_patch_kernel", C.c_double), ("ms_total_device", C.c_double),
                ("n_patch_launches", C.c_uint64), ("n_kernel_launches", C.c_uint64),
                ("ms_optimise_phases", C.c_double), ("n_grid_barriers", C.c_uint64),
                ("ms_optimise_thread_phases", C.c_double), ("ms_sort_phases", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


PATCH_IN = np.dtype([("x", "<i4"), ("y", "<i4"), ("depth", "<f4"), ("dz_i", "<f4"), ("dz_j", "<f4"),
                     ("n_local", "<i4"), ("local_ids", "<i4", (4,))])
PATCH_OUT = np.dtype([("conf", "<f4"), ("depth", "<f4"), ("dz_i", "<f4"), ("dz_j", "<f4"),
                      ("normal", "<f4", (3,)), ("n_local", "<i4"), ("local_ids", "<i4", (4,)),
                      ("iterations", "<i4"), ("converged", "<i4"), ("opti_success", "<i4")])

EXPORTS = ["b200mvs_default_settings", "b200mvs_create", "b200mvs_destroy", "b200mvs_last_error", "b200mvs_version",
           "b200mvs_upload_view", "b200mvs_upload_view_device", "b200mvs_set_view_camera", "b200mvs_set_features", "b200mvs_num_levels",
           "b200mvs_get_level", "b200mvs_global_view_selection", "b200mvs_optimize_patches", "b200mvs_reconstruct",
           "b200mvs_plan_views", "b200mvs_set_patch_mode", "b200mvs_depthmap_last_error", "b200mvs_depthmap_confidence_clean",
           "b200mvs_depthmap_cleanup", "b200mvs_depthmap_triangulate", "b200mvs_depthmap_pointset"]


def lib():
    """Loads libb200mvs.so (building it when sources are newer). Fails loudly when it cannot."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("B200MVS_LIB") or _build.LIB      # B200MVS_LIB: kernel-variant experiments (tools/kbench.py)
    if path == _build.LIB and _build.needs_build():
        path = _build.build()
    L = C.CDLL(path)
    L.b200mvs_last_error.restype = C.c_char_p
    L.b200mvs_last_error.argtypes = [C.c_void_p]
    L.b200mvs_version.restype = C.c_char_p
    L.b200mvs_default_settings.argtypes = [C.c_void_p]
    L.b200mvs_create.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.b200mvs_destroy.argtypes = [C.c_void_p]
    L.b200mvs_upload_view.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200mvs_upload_view_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200mvs_set_features.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200mvs_set_view_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200mvs_num_levels.argtypes = [C.c_void_p, C.c_int]
    L.b200mvs_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200mvs_global_view_selection.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.b200mvs_optimize_patches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_void_p]
    L.b200mvs_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]
    L.b200mvs_plan_views.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.b200mvs_set_patch_mode.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    _LIB = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """Device-resident counterpart of mve::Scene for this path: views (image + camera) and bundle features."""

    def __init__(self, n_views: int, device: int = 0):
        self._lib = lib()
        self.n_views = n_views
        self.device = device
        h = C.c_void_p()
        rc = self._lib.b200mvs_create(device, n_views, C.byref(h))
        if rc != 0:
            raise B200MVSError(rc, self._lib.b200mvs_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200mvs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise B200MVSError(rc, self._lib.b200mvs_last_error(self._h).decode())
        return rc

    @classmethod
    def from_synth(cls, s, device: int = 0, views: Optional[Sequence[int]] = None) -> "Scene":
        """Uploads a mve_b200.synth.Scene (host images -> device, pyramids built on the device)."""
        sc = cls(s.n_views, device)
        for v in (range(s.n_views) if views is None else views):
            sc.set_view(v, s.images[v], s.flen[v], s.paspect[v], s.ppoint[v], s.rot[v], s.trans[v])
        sc.set_features(s.feat_pos, s.feat_refs)
        return sc

    def set_view(self, view_id: int, image: np.ndarray, flen, paspect, ppoint, rot, trans):
        """mve::View with an `undistorted` uint8 image + mve::CameraInfo (camera.h:23-170)."""
        img = np.ascontiguousarray(image, dtype=np.uint8)
        if img.ndim == 2:
            img = img[:, :, None]
        h, w, ch = img.shape
        pp = np.ascontiguousarray(ppoint, np.float32)
        r = np.ascontiguousarray(rot, np.float32).reshape(9)
        t = np.ascontiguousarray(trans, np.float32).reshape(3)
        self._check(self._lib.b200mvs_upload_view(self._h, view_id, _p(img), w, h, ch, float(flen), float(paspect),
                                                  _p(pp), _p(r), _p(t)))

    def set_view_device(self, view_id: int, dev_ptr: int, w: int, h: int, flen, paspect, ppoint, rot, trans, stream: int = 0):
        """Same, the image (h x w x 3 uint8) already being in this device's memory (e.g. a torch tensor's data_ptr())."""
        pp = np.ascontiguousarray(ppoint, np.float32)
        r = np.ascontiguousarray(rot, np.float32).reshape(9)
        t = np.ascontiguousarray(trans, np.float32).reshape(3)
        self._check(self._lib.b200mvs_upload_view_device(self._h, view_id, C.c_void_p(dev_ptr), w, h, float(flen),
                                                         float(paspect), _p(pp), _p(r), _p(t), C.c_void_p(stream)))

    def set_view_camera(self, view_id: int, w: int, h: int, flen, paspect, ppoint, rot, trans):
        """SingleView::create: camera + image size, the colour image is loaded later (or never, if not needed)."""
        pp = np.ascontiguousarray(ppoint, np.float32)
        r = np.ascontiguousarray(rot, np.float32).reshape(9)
        t = np.ascontiguousarray(trans, np.float32).reshape(3)
        self._check(self._lib.b200mvs_set_view_camera(self._h, view_id, w, h, float(flen), float(paspect), _p(pp), _p(r), _p(t)))

    def set_features(self, pos: np.ndarray, refs: Sequence[np.ndarray]):
        """mve::Bundle::Features (bundle.h:51-60)."""
        off = np.zeros(len(refs) + 1, np.int32)
        if len(refs):
            off[1:] = np.cumsum([len(r) for r in refs])
        ids = np.concatenate(refs).astype(np.int32) if len(refs) else np.zeros(0, np.int32)
        p = np.ascontiguousarray(pos, np.float32)
        self._check(self._lib.b200mvs_set_features(self._h, len(refs), _p(p), _p(off), _p(ids)))

    def num_levels(self, view_id: int) -> int:
        return self._check(self._lib.b200mvs_num_levels(self._h, view_id))

    def level(self, view_id: int, level: int) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        self._check(self._lib.b200mvs_get_level(self._h, view_id, level, C.byref(w), C.byref(h), None))
        out = np.empty((h.value, w.value, 3), np.uint8)
        self._check(self._lib.b200mvs_get_level(self._h, view_id, level, C.byref(w), C.byref(h), _p(out)))
        return out

    def global_view_selection(self, settings: Settings, ref_view: int) -> List[int]:
        """DMRecon::globalViewSelection (dmrecon.cc:211-241)."""
        out = np.empty(64, np.int32)
        n = self._check(self._lib.b200mvs_global_view_selection(self._h, C.byref(settings), ref_view, _p(out), 64))
        return out[:n].tolist()

    def set_patch_mode(self, mode: int = 0, thread_min: int = -1):
        """Engine knob: 1 = one warp per patch, 2 = one thread per patch, 0 = by size (see include/b200mvs.h)."""
        self._check(self._lib.b200mvs_set_patch_mode(self._h, mode, thread_min))

    def plan_views(self, settings: Settings, ref_views: Sequence[int]):
        """Global view selection + seed lists of these reference views ahead of their reconstruct() call; safe to call from
        another thread while a reconstruct() of a previous batch is running (ctypes releases the GIL)."""
        refs = np.asarray(ref_views, np.int32)
        self._check(self._lib.b200mvs_plan_views(self._h, C.byref(settings), len(refs), _p(refs)))

    def optimize_patches(self, settings: Settings, ref_view: int, global_ids: Sequence[int], patches: np.ndarray,
                         stats: Optional[Stats] = None) -> np.ndarray:
        """Batch of independent mvs::PatchOptimization runs (patch-level parity entry)."""
        patches = np.ascontiguousarray(patches, dtype=PATCH_IN)
        out = np.zeros(len(patches), PATCH_OUT)
        g = np.asarray(global_ids, np.int32)
        self._check(self._lib.b200mvs_optimize_patches(self._h, C.byref(settings), ref_view, _p(g), len(g), _p(patches),
                                                       len(patches), _p(out), C.byref(stats) if stats is not None else None))
        return out

    def reconstruct(self, settings: Settings, ref_views: Sequence[int], download: bool = True,
                    want=("depth", "conf", "dz", "normal", "view_ids"), out=None, progress=None):
        """DMRecon::start for a batch of reference views. Returns (list of map dicts or None, Stats).
        out: optional list (one dict per view) of preallocated host arrays (e.g. pinned) to receive the maps.
        progress: optional (Progress * n) array, updated live; setting .cancelled from another thread cancels the run."""
        refs = np.asarray(ref_views, np.int32)
        n = len(refs)
        stats = Stats()
        failed = C.c_int32(-1)
        maps_arr = None
        results = None
        if download:
            maps_arr = (_Maps * n)()
            results = []
            for j, r in enumerate(refs):
                w, h = C.c_int(), C.c_int()
                if self._lib.b200mvs_get_level(self._h, int(r), settings.scale, C.byref(w), C.byref(h), None) != 0:
                    # let b200mvs_reconstruct report it with the reference's own message (dmrecon.cc:37-75)
                    maps_arr, results = None, None
                    break
                W, H = w.value, h.value
                if out is not None:
                    d = out[j]
                    for k, a in d.items():
                        if a.shape[:2] != (H, W) or not a.flags["C_CONTIGUOUS"]:
                            raise ValueError("out[%d][%s] has the wrong shape" % (j, k))
                    results.append(d)
                    for k in ("depth", "conf", "dz", "normal", "view_ids"):
                        setattr(maps_arr[j], k, d[k].ctypes.data if k in d else None)
                    continue
                d = dict(depth=np.empty((H, W), np.float32))
                if "conf" in want:
                    d["conf"] = np.empty((H, W), np.float32)
                if "dz" in want:
                    d["dz"] = np.empty((H, W, 2), np.float32)
                if "normal" in want:
                    d["normal"] = np.empty((H, W, 3), np.float32)
                if "view_ids" in want:
                    d["view_ids"] = np.empty((H, W, 4), np.int32)
                results.append(d)
                for k in ("depth", "conf", "dz", "normal", "view_ids"):
                    setattr(maps_arr[j], k, d[k].ctypes.data if k in d else None)
        rc = self._lib.b200mvs_reconstruct(self._h, C.byref(settings), n, _p(refs), maps_arr, progress, C.byref(stats),
                                           C.byref(failed))
        if rc != 0:
            msg = self._lib.b200mvs_last_error(self._h).decode()
            if failed.value >= 0:
                msg += " (view %d)" % failed.value
            raise B200MVSError(rc, msg)
        return results, stats


class DMRecon:
    """mvs::DMRecon (dmrecon.h:40-68): construct with a scene and settings, call start()."""

    def __init__(self, scene: Scene, settings: Settings, ref_view_nr: int):
        if ref_view_nr < 0 or ref_view_nr >= scene.n_views:
            raise ValueError("Master view index out of bounds")          # dmrecon.cc:37-38
        if settings.scale < 0:
            raise ValueError("Invalid scale factor")                     # dmrecon.cc:41-42
        self.scene, self.settings, self.ref_view_nr = scene, settings, ref_view_nr
        self.maps = None
        self.stats = None

    def getRefViewNr(self) -> int:
        return self.ref_view_nr

    def start(self):
        maps, stats = self.scene.reconstruct(self.settings, [self.ref_view_nr])
        self.maps, self.stats = maps[0], stats
        return self.maps
