"""Host-side mirror of the reference's public sub-solver classes (registration.h:107-317) and of
RobustRegistrationSolver::computeTIMs (registration.cc:471-505), the calls the reference's own unit tests make
(scale-solver-test.cc, rotation-solver-test.cc, translation-solver-test.cc).  Same class names, constructor
arguments and method names; arrays are 3xN numpy (column = point, like the Eigen matrices), results are returned
instead of written through pointers.  Every method is one call into libpsulvsb_b200.so (CUDA; no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import capi


def _cm(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != 3:
        raise ValueError("expected a 3xN array")
    return np.asfortranarray(a)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def computeTIMs(v):
    """registration.cc:471-505 -> (3 x N(N-1)/2 line vectors v_j - v_i, 2 x L index map)."""
    v = _cm(v)
    n = v.shape[1]
    L = n * (n - 1) // 2
    tims = np.zeros((3, L), order="F")
    vmap = np.zeros((2, L), dtype=np.int32, order="F")
    capi.check(capi.lib().psulvsb_compute_tims_host(_ptr(v), n, _ptr(tims), _ptr(vmap)))
    return tims, vmap


class ScaleInliersSelector:  # registration.h:203-215, registration.cc:418-434
    def __init__(self, noise_bound: float, cbar2: float):
        self.noise_bound, self.cbar2 = float(noise_bound), float(cbar2)

    def solveForScale(self, src, dst):
        src, dst = _cm(src), _cm(dst)
        inl = np.zeros(src.shape[1], dtype=np.uint8)
        capi.check(capi.lib().psulvsb_scale_inliers_host(_ptr(src), _ptr(dst), src.shape[1], self.noise_bound,
                                                         self.cbar2, _ptr(inl)))
        return 1.0, inl.astype(bool)


class TLSScaleSolver:  # registration.h:180-198, registration.cc:397-415
    """The reference draws its RANSAC candidates with rand(); here they come from the replayable stream
    philox(seed; DOMAIN_SCALE, event, k), event = number of earlier calls on this object."""

    def __init__(self, noise_bound: float, cbar2: float, seed: int = 0):
        self.noise_bound, self.cbar2, self.seed = float(noise_bound), float(cbar2), int(seed)
        self.calls = 0

    def solveForScale(self, src, dst, last_best=None):
        src, dst = _cm(src), _cm(dst)
        inl = np.zeros(src.shape[1], dtype=np.uint8)
        s = C.c_double(0.0)
        lb = C.byref(C.c_double(float(last_best))) if last_best is not None else None
        capi.check(capi.lib().psulvsb_tls_scale_host(_ptr(src), _ptr(dst), src.shape[1], self.noise_bound, self.cbar2,
                                                     self.seed, self.calls, lb, C.byref(s), _ptr(inl)))
        self.calls += 1
        return s.value, inl.astype(bool)


class GNCTLSRotationSolver:  # registration.h:263-290, registration.cc:1563-1692
    @dataclasses.dataclass
    class Params:  # registration.h:246-251
        max_iterations: int = 100
        cost_threshold: float = 1e-6
        gnc_factor: float = 1.4
        noise_bound: float = 0.01

    def __init__(self, params: "GNCTLSRotationSolver.Params"):
        self.params_ = params
        self.cost_ = 0.0
        self.iterations_ = 0

    def getParams(self):
        return self.params_

    def setParams(self, params):
        self.params_ = params

    def getCostAtTermination(self):
        return self.cost_

    def solveForRotation(self, src, dst, last_best=None):
        src, dst = _cm(src), _cm(dst)
        p = self.params_
        R = np.zeros(9)
        inl = np.zeros(src.shape[1], dtype=np.uint8)
        cost = C.c_double(0.0)
        its = C.c_int(0)
        lb = None if last_best is None else np.asfortranarray(np.asarray(last_best, dtype=np.float64))
        capi.check(capi.lib().psulvsb_gnc_tls_rotation_host(_ptr(src), _ptr(dst), src.shape[1], p.noise_bound,
                                                            int(p.max_iterations), p.gnc_factor, p.cost_threshold,
                                                            _ptr(lb), _ptr(R), _ptr(inl), C.byref(cost), C.byref(its)))
        self.cost_, self.iterations_ = cost.value, its.value
        return R.reshape(3, 3, order="F"), inl.astype(bool)


class TLSTranslationSolver:  # registration.h:221-239, registration.cc:436-463
    def __init__(self, noise_bound: float, cbar2: float):
        self.noise_bound, self.cbar2 = float(noise_bound), float(cbar2)

    def solveForTranslation(self, src, dst, last_best=None):
        src, dst = _cm(src), _cm(dst)
        t = np.zeros(3)
        inl = np.zeros(src.shape[1], dtype=np.uint8)
        lb = None if last_best is None else np.ascontiguousarray(last_best, dtype=np.float64)
        capi.check(capi.lib().psulvsb_tls_translation_host(_ptr(src), _ptr(dst), src.shape[1], self.noise_bound,
                                                           self.cbar2, _ptr(lb), _ptr(t), _ptr(inl)))
        return t, inl.astype(bool)
