"""numpy restatement of the reference driver's normal-angle histogram pre-filter.

TEST INFRASTRUCTURE ONLY.  Follows examples/teaser_cpp_ply/PSULVSB.cc:87-172 (histogram_outlier_removal)
and :174-188 (mask_filter) statement by statement; the reference has no fixture for it, so parity for the
pre-filter is "product vs this restatement" (unpinned by the reference's own tests).
Positions taken where the reference is undefined (same as the product, csrc/host_io.cpp): sigma == 0 -> one
bin; an angle exactly on the upper edge of the last bin goes into the last bin.
"""
import math

import numpy as np


def histogram_outlier_removal(src_normals, tgt_normals):
    n = src_normals.shape[1]
    keep = np.zeros(n, dtype=np.int32)
    all_angles = np.full(n, -1.0)
    remain = []
    o_max, o_min, s = 0.0, float(2**31 - 1), 0.0
    for i in range(n):
        a, b = src_normals[:, i], tgt_normals[:, i]
        za, zb = (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2], (b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]
        if za > 0:
            a = a / math.sqrt(za)
        if zb > 0:
            b = b / math.sqrt(zb)
        c = (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
        # std::max(-1.0, std::min(1.0, c)) (PSULVSB.cc:100): std::min(1.0, NaN) returns its FIRST argument, so a
        # NaN normal (PCL's answer for a degenerate neighbourhood) yields cos = 1, angle 0 -- it is NOT skipped
        # by the isnan test that follows in the reference
        c = 1.0 if math.isnan(c) else max(-1.0, min(1.0, c))
        ang = math.acos(c) * 180.0 / math.pi
        if math.isnan(ang):
            continue
        remain.append(ang)
        all_angles[i] = ang
        o_min, o_max = min(ang, o_min), max(ang, o_max)
        s += ang
    if not remain:
        return keep, 0
    cnt = float(len(remain))
    mean = s / cnt
    sq = 0.0
    for d in remain:
        sq += (d - mean) ** 2
    sd = math.sqrt(sq / cnt)
    width = 3.49 * sd / cnt ** (1.0 / 3.0)
    ok = width > 0 and math.isfinite(width)
    size = max(1, int(math.ceil((o_max - o_min) / width))) if ok else 1
    hist = [[] for _ in range(size)]
    peak_id, peak_h = 0, 0
    for i in range(n):
        if all_angles[i] == -1:
            continue
        b = int((all_angles[i] - o_min) / width) if ok else 0
        b = min(max(b, 0), size - 1)
        hist[b].append(i)
        if len(hist[b]) > peak_h:
            peak_h, peak_id = len(hist[b]), b
    hmean = sum(float(len(h)) for h in hist) / size
    hvar = 0.0
    for h in hist:
        hvar += (len(h) - hmean) ** 2
    thr = hmean + math.sqrt(hvar / size)
    rem = 0
    for i, h in enumerate(hist):
        if abs(i - peak_id) > 2:
            for j in h:
                keep[j] = -1
        if len(h) > thr:
            for j in h:
                keep[j] = 1
                rem += 1
    return keep, rem


def mask_filter(src, tgt, keep):
    idx = np.flatnonzero(keep == 1)
    rm = np.full(src.shape[1], -1, dtype=np.int32)
    rm[idx] = np.arange(idx.size, dtype=np.int32)
    return np.asfortranarray(src[:, idx]), np.asfortranarray(tgt[:, idx]), rm


def knn_pca_normals(points, k=20, viewpoint=(0.0, 0.0, 0.0)):
    """Restatement of what PSULVSB.cc:35-85 asks PCL for: for every point the k nearest neighbours (itself
    included), the eigenvector of the smallest eigenvalue of their covariance, flipped towards the viewpoint."""
    P = np.asarray(points, dtype=np.float64)
    n = P.shape[1]
    Pf = P.astype(np.float32)
    out = np.zeros((3, n))
    vp = np.asarray(viewpoint, dtype=np.float64)
    for i in range(n):
        d = ((Pf - Pf[:, i:i + 1]) ** 2).sum(axis=0)
        idx = np.argsort(d, kind="stable")[:k]
        nb = P[:, idx]
        c = np.cov(nb, bias=True) * nb.shape[1]
        w, v = np.linalg.eigh(c)
        nv = v[:, 0]
        if (vp - P[:, i]) @ nv < 0:
            nv = -nv
        out[:, i] = nv
    return out
