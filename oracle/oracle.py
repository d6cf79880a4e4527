"""ctypes binding of the CPU parity oracle (oracle/psulvsb_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpsulvsb_oracle.so")

DOMAIN_L_SAMPLED = 1
DOMAIN_BASIC = 2
DOMAIN_UNIFORM = 3
DOMAIN_SCALE = 4


class Params(C.Structure):
    _fields_ = [
        ("noise_bound", C.c_double),
        ("cbar2", C.c_double),
        ("estimate_scaling", C.c_int),
        ("rotation_max_iterations", C.c_int),
        ("rotation_gnc_factor", C.c_double),
        ("rotation_cost_threshold", C.c_double),
        ("inlier_selection_mode", C.c_int),
        ("kcore_heuristic_threshold", C.c_double),
        ("score_noise_bound", C.c_double),
        ("inloop_noise_bound", C.c_double),
        ("inloop_cbar2", C.c_double),
        ("inloop_max_iterations", C.c_int),
        ("inloop_gnc_factor", C.c_double),
        ("inloop_cost_threshold", C.c_double),
        ("rotation_similar", C.c_double),
        ("local_max_iter", C.c_int),
        ("tpro_host", C.c_double),
        ("tpro_local", C.c_double),
        ("host_round_limit", C.c_int),
        ("wallclock_cap_s", C.c_double),
        ("self_update", C.c_int),
        ("seed", C.c_uint64),
    ]


class Solution(C.Structure):
    _fields_ = [
        ("valid", C.c_int),
        ("scale", C.c_double),
        ("final_inlier_count", C.c_int),
        ("translation", C.c_double * 3),
        ("rotation", C.c_double * 9),
        ("host_rounds", C.c_int),
        ("local_iters", C.c_int),
        ("n_line_vectors", C.c_longlong),
        ("n_reduced", C.c_longlong),
        ("final_C", C.c_int),
        ("refined", C.c_int),
        ("escalations", C.c_int),
    ]


class LocalTrace(C.Structure):
    _fields_ = [
        ("host_round", C.c_int),
        ("local_iter", C.c_int),
        ("n_sampled_lines", C.c_int),
        ("n_sampled_points", C.c_int),
        ("basic_choose", C.c_int),
        ("gnc_iterations", C.c_int),
        ("rot_inliers", C.c_int),
        ("n_rot_points", C.c_int),
        ("similar", C.c_int),
        ("curr_count", C.c_int),
        ("best_count", C.c_int),
        ("local_r", C.c_int),
        ("p_local", C.c_double),
        ("l_rate", C.c_double),
        ("b_rate", C.c_double),
        ("scale", C.c_double),
        ("R", C.c_double * 9),
        ("t", C.c_double * 3),
    ]


class HostTrace(C.Structure):
    _fields_ = [
        ("host_round", C.c_int),
        ("curr_count", C.c_int),
        ("best_host", C.c_int),
        ("new_corr_count", C.c_int),
        ("inlier_map_size", C.c_int),
        ("host_r", C.c_int),
        ("p_host", C.c_double),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("local", C.POINTER(LocalTrace)),
        ("local_cap", C.c_int),
        ("local_n", C.c_int),
        ("host", C.POINTER(HostTrace)),
        ("host_cap", C.c_int),
        ("host_n", C.c_int),
        ("final_inliers", C.POINTER(C.c_int)),
        ("inlier_counter", C.POINTER(C.c_int)),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only, no external dependency)."""
    src = os.path.join(_HERE, "psulvsb_oracle.cpp")
    hdr = os.path.join(_HERE, "psulvsb_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpsulvsb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


def _declare(L: C.CDLL) -> None:
    L.oracle_default_params.argtypes = [C.POINTER(Params)]
    L.oracle_default_params.restype = None
    L.oracle_solve.argtypes = [C.POINTER(Params), _dp, _dp, C.c_int, _dp, _dp, C.c_int, _ip, _ip,
                               C.POINTER(Solution), C.POINTER(Trace)]
    L.oracle_solve.restype = C.c_int
    L.oracle_consistency_mask.argtypes = [_dp, _dp, C.c_int, C.c_double, _u8p, _dp]
    L.oracle_consistency_mask.restype = None
    L.oracle_scale_inliers.argtypes = [_dp, _dp, C.c_longlong, C.c_double, _u8p]
    L.oracle_scale_inliers.restype = None
    L.oracle_reduced_set.argtypes = [_dp, _dp, C.c_int, C.c_double, _ip, _ip, C.c_longlong]
    L.oracle_reduced_set.restype = C.c_longlong
    L.oracle_svd_rot.argtypes = [_dp, _dp, _dp, C.c_longlong, _dp]
    L.oracle_svd_rot.restype = None
    L.oracle_svd3.argtypes = [_dp, _dp, _dp, _dp]
    L.oracle_svd3.restype = None
    L.oracle_gnc_tls.argtypes = [_dp, _dp, C.c_longlong, C.c_double, C.c_int, C.c_double, C.c_double, _dp, _dp,
                                 _u8p, _dp]
    L.oracle_gnc_tls.restype = C.c_int
    L.oracle_tls_translation.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double, _dp, _dp, _u8p]
    L.oracle_tls_translation.restype = None
    L.oracle_tls_scale.argtypes = [_dp, _dp, C.c_longlong, C.c_double, C.c_double, _dp, C.c_uint64, C.c_uint32,
                                   _dp, _u8p]
    L.oracle_tls_scale.restype = C.c_int
    L.oracle_score.argtypes = [_dp, _dp, C.c_int, C.c_double, _dp, _dp, C.c_double, _u8p, _dp]
    L.oracle_score.restype = C.c_int
    L.oracle_weighted_svd.argtypes = [_dp, _dp, _ip, C.c_int, _dp, _dp]
    L.oracle_weighted_svd.restype = None
    L.oracle_rmse.argtypes = [_dp, _dp, _ip, C.c_int, _dp]
    L.oracle_rmse.restype = C.c_double
    L.oracle_inlier_probability.argtypes = [C.c_double, C.c_double]
    L.oracle_inlier_probability.restype = C.c_double
    L.oracle_philox4x32.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32)]
    L.oracle_philox4x32.restype = None
    L.oracle_rand31.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    L.oracle_rand31.restype = C.c_uint32
    L.oracle_uniform01.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    L.oracle_uniform01.restype = C.c_double
    L.oracle_sample_without_replacement.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_longlong,
                                                    C.c_longlong, C.POINTER(C.c_int64)]
    L.oracle_sample_without_replacement.restype = C.c_longlong
    L.oracle_max_clique.argtypes = [C.c_int, _ip, _ip, C.c_longlong, _ip]
    L.oracle_max_clique.restype = C.c_int


def _cm(a) -> np.ndarray:
    """3xN array -> contiguous column-major buffer (Eigen layout)."""
    a = np.asarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == 3, a.shape
    return np.asfortranarray(a)


def _p(a: np.ndarray, t=_dp):
    return a.ctypes.data_as(t)


def default_params(**kw) -> Params:
    p = Params()
    lib().oracle_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def solve(params: Params, src, dst, ori_src=None, ori_dst=None, keep_mask=None, reduce_map=None,
          trace_cap: int = 4096):
    """Returns (Solution, dict(trace)).  src/dst are 3xC; ori_* default to src/dst (C == M)."""
    src = _cm(src)
    dst = _cm(dst)
    Cn = src.shape[1]
    if ori_src is None:
        ori_src, ori_dst = src, dst
    ori_src = _cm(ori_src)
    ori_dst = _cm(ori_dst)
    M = ori_src.shape[1]
    if keep_mask is None:
        assert Cn == M
        keep_mask = np.ones(M, dtype=np.int32)
        reduce_map = np.arange(M, dtype=np.int32)
    keep_mask = np.ascontiguousarray(keep_mask, dtype=np.int32)
    reduce_map = np.ascontiguousarray(reduce_map, dtype=np.int32)
    sol = Solution()
    loc = (LocalTrace * trace_cap)()
    hst = (HostTrace * trace_cap)()
    fin = np.zeros(M, dtype=np.int32)
    cnt = np.zeros(M, dtype=np.int32)
    tr = Trace(C.cast(loc, C.POINTER(LocalTrace)), trace_cap, 0, C.cast(hst, C.POINTER(HostTrace)), trace_cap, 0,
               _p(fin, _ip), _p(cnt, _ip))
    rc = lib().oracle_solve(C.byref(params), _p(src), _p(dst), Cn, _p(ori_src), _p(ori_dst), M, _p(keep_mask, _ip),
                            _p(reduce_map, _ip), C.byref(sol), C.byref(tr))
    if rc != 0:
        raise RuntimeError(f"oracle_solve failed rc={rc}")
    trace = {
        "local": [loc[i] for i in range(tr.local_n)],
        "host": [hst[i] for i in range(tr.host_n)],
        "final_inliers": fin,
        "inlier_counter": cnt,
    }
    return sol, trace


def solution_R(sol) -> np.ndarray:
    return np.array(sol.rotation[:]).reshape(3, 3, order="F")


def solution_t(sol) -> np.ndarray:
    return np.array(sol.translation[:])


def consistency_mask(src, dst, beta: float, want_margin: bool = False):
    src = _cm(src)
    dst = _cm(dst)
    n = src.shape[1]
    mask = np.zeros((n, n), dtype=np.uint8)
    margin = np.zeros((n, n), dtype=np.float64) if want_margin else None
    lib().oracle_consistency_mask(_p(src), _p(dst), n, beta, _p(mask, _u8p), _p(margin) if want_margin else None)
    return (mask, margin) if want_margin else mask


def scale_inliers(sv, tv, beta: float) -> np.ndarray:
    sv = _cm(sv)
    tv = _cm(tv)
    K = sv.shape[1]
    mask = np.zeros(K, dtype=np.uint8)
    lib().oracle_scale_inliers(_p(sv), _p(tv), K, beta, _p(mask, _u8p))
    return mask


def reduced_set(src, dst, beta: float):
    src = _cm(src)
    dst = _cm(dst)
    n = src.shape[1]
    cnt = lib().oracle_reduced_set(_p(src), _p(dst), n, beta, None, None, 0)
    pi = np.zeros(max(cnt, 1), dtype=np.int32)
    pj = np.zeros(max(cnt, 1), dtype=np.int32)
    lib().oracle_reduced_set(_p(src), _p(dst), n, beta, _p(pi, _ip), _p(pj, _ip), cnt)
    return pi[:cnt], pj[:cnt]


def svd_rot(X, Y, W) -> np.ndarray:
    X = _cm(X)
    Y = _cm(Y)
    W = np.ascontiguousarray(W, dtype=np.float64)
    R = np.zeros(9)
    lib().oracle_svd_rot(_p(X), _p(Y), _p(W), X.shape[1], _p(R))
    return R.reshape(3, 3, order="F")


def svd3(A):
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    U = np.zeros(9)
    S = np.zeros(3)
    V = np.zeros(9)
    lib().oracle_svd3(_p(A), _p(U), _p(S), _p(V))
    return U.reshape(3, 3, order="F"), S, V.reshape(3, 3, order="F")


def gnc_tls(sv, tv, noise_bound, max_iterations=100, gnc_factor=1.4, cost_threshold=0.005, R_init=None):
    sv = _cm(sv)
    tv = _cm(tv)
    K = sv.shape[1]
    R = np.zeros(9)
    inl = np.zeros(K, dtype=np.uint8)
    cost = C.c_double(0)
    ri = None
    if R_init is not None:
        ri = np.asfortranarray(np.asarray(R_init, dtype=np.float64))
    its = lib().oracle_gnc_tls(_p(sv), _p(tv), K, noise_bound, max_iterations, gnc_factor, cost_threshold,
                               _p(ri) if ri is not None else None, _p(R), _p(inl, _u8p), C.byref(cost))
    return R.reshape(3, 3, order="F"), inl, its, cost.value


def tls_translation(src, dst, noise_bound, cbar2=1.0, last_best=None):
    src = _cm(src)
    dst = _cm(dst)
    N = src.shape[1]
    t = np.zeros(3)
    inl = np.zeros(N, dtype=np.uint8)
    lb = None if last_best is None else np.ascontiguousarray(last_best, dtype=np.float64)
    lib().oracle_tls_translation(_p(src), _p(dst), N, noise_bound, cbar2, _p(lb) if lb is not None else None,
                                 _p(t), _p(inl, _u8p))
    return t, inl


def tls_scale(sv, tv, noise_bound, cbar2=1.0, last_best=None, seed=0, event=0):
    sv = _cm(sv)
    tv = _cm(tv)
    K = sv.shape[1]
    s = C.c_double(0)
    inl = np.zeros(K, dtype=np.uint8)
    lb = None if last_best is None else C.byref(C.c_double(last_best))
    its = lib().oracle_tls_scale(_p(sv), _p(tv), K, noise_bound, cbar2, lb, seed, event, C.byref(s), _p(inl, _u8p))
    return s.value, inl, its


def score(P, Q, scale, R, t, tau):
    P = _cm(P)
    Q = _cm(Q)
    N = P.shape[1]
    Rc = np.asfortranarray(np.asarray(R, dtype=np.float64))
    tc = np.ascontiguousarray(t, dtype=np.float64)
    inl = np.zeros(N, dtype=np.uint8)
    res = np.zeros(N)
    cnt = lib().oracle_score(_p(P), _p(Q), N, scale, _p(Rc), _p(tc), tau, _p(inl, _u8p), _p(res))
    return cnt, inl, res


def weighted_svd(src, tgt, w, T_init) -> np.ndarray:
    src = _cm(src)
    tgt = _cm(tgt)
    w = np.ascontiguousarray(w, dtype=np.int32)
    Ti = np.asfortranarray(np.asarray(T_init, dtype=np.float64))
    To = np.zeros(16)
    lib().oracle_weighted_svd(_p(src), _p(tgt), _p(w, _ip), src.shape[1], _p(Ti), _p(To))
    return To.reshape(4, 4, order="F")


def rmse(src, tgt, mask, T) -> float:
    src = _cm(src)
    tgt = _cm(tgt)
    mask = np.ascontiguousarray(mask, dtype=np.int32)
    Tc = np.asfortranarray(np.asarray(T, dtype=np.float64))
    return lib().oracle_rmse(_p(src), _p(tgt), _p(mask, _ip), src.shape[1], _p(Tc))


def inlier_probability(r: float, sigma: float) -> float:
    return lib().oracle_inlier_probability(r, sigma)


def philox(seed: int, domain: int, event: int, block: int) -> np.ndarray:
    out = (C.c_uint32 * 4)()
    lib().oracle_philox4x32(seed, domain, event, block, out)
    return np.array(out[:], dtype=np.uint32)


def rand31(seed, domain, event, k) -> int:
    return lib().oracle_rand31(seed, domain, event, k)


def uniform01(seed, domain, event, k) -> float:
    return lib().oracle_uniform01(seed, domain, event, k)


def sample_without_replacement(seed, domain, event, n, count):
    out = np.zeros(max(count, 1), dtype=np.int64)
    consumed = lib().oracle_sample_without_replacement(seed, domain, event, n, count,
                                                       out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out[:count], consumed


def max_clique(n, edges) -> np.ndarray:
    e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
    u = np.ascontiguousarray(e[:, 0])
    v = np.ascontiguousarray(e[:, 1])
    out = np.zeros(max(n, 1), dtype=np.int32)
    k = lib().oracle_max_clique(n, _p(u, _ip), _p(v, _ip), len(e), _p(out, _ip))
    return out[:k]
