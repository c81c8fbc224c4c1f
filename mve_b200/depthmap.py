"""Host-side mirror of the reference's depth-map consumers over the C ABI (include/b200mvs.h):
mve::image::depthmap_confidence_clean / depthmap_cleanup and mve::geom::depthmap_triangulate (libs/mve/depthmap.{h,cc}),
the per-view work of apps/scene2pset.  No CPU fallback: the calls fail without a CUDA device."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import dmrecon

DD_FACTOR_DEFAULT = 5.0      # mve::geom::DD_FACTOR_DEFAULT (libs/mve/depthmap.h)


def _lib():
    L = dmrecon.lib()
    if not getattr(L, "_dm_ready", False):
        L.b200mvs_depthmap_last_error.restype = C.c_char_p
        L.b200mvs_depthmap_confidence_clean.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.b200mvs_depthmap_cleanup.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
        L.b200mvs_depthmap_triangulate.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]
        L.b200mvs_depthmap_pointset.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_float, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L._dm_ready = True
    return L


def _check(rc):
    if rc < 0:
        raise dmrecon.B200MVSError(rc, _lib().b200mvs_depthmap_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def depthmap_confidence_clean(dm: np.ndarray, cm: np.ndarray, device: int = 0) -> None:
    """In place: dm = 0 where cm <= 0 (depthmap.cc:118-131)."""
    if dm.shape != cm.shape:
        raise ValueError("Image dimensions do not match")
    assert dm.dtype == np.float32 and dm.flags["C_CONTIGUOUS"]
    cm = np.ascontiguousarray(cm, np.float32)
    _check(_lib().b200mvs_depthmap_confidence_clean(device, _p(dm), _p(cm), dm.shape[1], dm.shape[0]))


def depthmap_cleanup(dm: np.ndarray, thres: int, device: int = 0) -> np.ndarray:
    """Islands of dm != 0 smaller than thres pixels removed (depthmap.cc:25-113)."""
    dm = np.ascontiguousarray(dm, np.float32)
    out = np.empty_like(dm)
    _check(_lib().b200mvs_depthmap_cleanup(device, _p(dm), dm.shape[1], dm.shape[0], int(thres), _p(out)))
    return out


def depthmap_triangulate(dm: np.ndarray, invproj: np.ndarray, dd_factor: float = DD_FACTOR_DEFAULT,
                         cam_to_world: Optional[np.ndarray] = None, color: Optional[np.ndarray] = None, device: int = 0):
    """mve::geom::depthmap_triangulate (depthmap.cc:196-375). Returns dict(vertex_ids [H,W] uint32, vertices [V,3], colors [V,4]
    or None, faces [F,3] uint32, device_ms)."""
    dm = np.ascontiguousarray(dm, np.float32)
    h, w = dm.shape
    ip = np.ascontiguousarray(invproj, np.float32).reshape(9)
    ctw = None if cam_to_world is None else np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
    cch = 0
    if color is not None:
        color = np.ascontiguousarray(color, np.uint8)
        if color.shape[:2] != (h, w):
            raise ValueError("Color image dimension mismatch")
        cch = 1 if color.ndim == 2 else color.shape[2]
    cap_v, cap_f = w * h, 2 * (w - 1) * (h - 1)
    vids = np.empty((h, w), np.uint32)
    verts = np.empty((cap_v, 3), np.float32)
    cols = np.empty((cap_v, 4), np.float32) if color is not None else None
    faces = np.empty((cap_f, 3), np.uint32)
    nv, nf, ms = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    _check(_lib().b200mvs_depthmap_triangulate(device, _p(dm), w, h, _p(ip), float(dd_factor), _p(ctw), _p(color), cch, _p(vids), _p(verts),
                                               _p(cols), _p(faces), cap_v, cap_f, C.byref(nv), C.byref(nf), C.byref(ms)))
    return dict(vertex_ids=vids, vertices=verts[:nv.value].copy(), colors=None if cols is None else cols[:nv.value].copy(),
                faces=faces[:nf.value].copy(), device_ms=ms.value)


def depthmap_pointset(dm: np.ndarray, invproj: np.ndarray, dd_factor: float = DD_FACTOR_DEFAULT,
                      cam_to_world: Optional[np.ndarray] = None, color: Optional[np.ndarray] = None,
                      with_normals: bool = True, conf_iterations: int = 4, scale_factor: Optional[float] = 2.5, device: int = 0):
    """The per-view work of apps/scene2pset (scene2pset.cc:264-358): triangulation + vertex normals + boundary confidences +
    scale values. Returns the dict of depthmap_triangulate plus normals [V,3], confidences [V], scales [V] (None when skipped)."""
    dm = np.ascontiguousarray(dm, np.float32)
    h, w = dm.shape
    ip = np.ascontiguousarray(invproj, np.float32).reshape(9)
    ctw = None if cam_to_world is None else np.ascontiguousarray(cam_to_world, np.float32).reshape(16)
    cch = 0
    if color is not None:
        color = np.ascontiguousarray(color, np.uint8)
        cch = 1 if color.ndim == 2 else color.shape[2]
    cap_v, cap_f = w * h, 2 * (w - 1) * (h - 1)
    vids = np.empty((h, w), np.uint32)
    verts = np.empty((cap_v, 3), np.float32)
    cols = np.empty((cap_v, 4), np.float32) if color is not None else None
    faces = np.empty((cap_f, 3), np.uint32)
    nrm = np.empty((cap_v, 3), np.float32) if with_normals else None
    cf = np.empty(cap_v, np.float32) if conf_iterations > 0 else None
    sc = np.empty(cap_v, np.float32) if scale_factor is not None else None
    nv, nf, ms = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    _check(_lib().b200mvs_depthmap_pointset(device, _p(dm), w, h, _p(ip), float(dd_factor), _p(ctw), _p(color), cch, _p(vids), _p(verts),
                                            _p(cols), _p(faces), _p(nrm), _p(cf), int(conf_iterations), _p(sc),
                                            float(scale_factor if scale_factor is not None else 0.0), cap_v, cap_f,
                                            C.byref(nv), C.byref(nf), C.byref(ms)))
    n = nv.value
    return dict(vertex_ids=vids, vertices=verts[:n].copy(), colors=None if cols is None else cols[:n].copy(), faces=faces[:nf.value].copy(),
                normals=None if nrm is None else nrm[:n].copy(), confidences=None if cf is None else cf[:n].copy(),
                scales=None if sc is None else sc[:n].copy(), device_ms=ms.value)
