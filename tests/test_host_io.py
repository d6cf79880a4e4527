"""Helpers around the hot path (include/psulvsb_io.h): the device pre-filter vs its numpy restatement (-m gpu),
PLY and correspondence-file readers on files written by the test (CPU)."""
import struct

import numpy as np
import pytest

import psulvsb_b200  # noqa: F401
from oracle import prefilter as OP
from psulvsb_b200 import capi, io


def test_io_symbols_exported():
    L = capi.lib()
    for name in io.IO_SYMBOLS:
        assert hasattr(L, name)


def _normals(n, seed, agree=0.6):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((3, n))
    b = a + 0.15 * rng.standard_normal((3, n))
    bad = rng.uniform(0, 1, n) > agree
    b[:, bad] = rng.standard_normal((3, int(bad.sum())))
    return a, b


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(50, 0), (500, 1), (5000, 2), (1889, 3), (20000, 4)])
def test_histogram_outlier_removal_matches_restatement(n, seed):
    a, b = _normals(n, seed)
    if seed == 1:
        a[:, 7] = np.nan       # PCL's NaN normals: std::min(1.0, NaN) == 1.0 in the reference -> angle 0, not skipped
        b[:, 11] = 0.0         # zero vector: Eigen's normalized() returns it unchanged
    keep_g, rem_g = io.histogram_outlier_removal(a, b)
    keep_o, rem_o = OP.histogram_outlier_removal(a, b)
    assert np.array_equal(keep_g, keep_o) and rem_g == rem_o
    assert set(np.unique(keep_g)) <= {-1, 0, 1}
    assert rem_g == int((keep_g == 1).sum()) and rem_g > 0


@pytest.mark.gpu
def test_histogram_degenerate_inputs():
    a = np.tile(np.array([[0.0], [0.0], [1.0]]), (1, 20))
    keep, rem = io.histogram_outlier_removal(a, a.copy())        # all angles equal: one bin, nothing above mean+sigma
    keep_o, rem_o = OP.histogram_outlier_removal(a, a.copy())
    assert np.array_equal(keep, keep_o) and rem == rem_o
    keep, rem = io.histogram_outlier_removal(np.zeros((3, 0)), np.zeros((3, 0)))
    assert keep.size == 0 and rem == 0


@pytest.mark.gpu
def test_prefilter_reduce_is_histogram_then_mask_filter():
    a, b = _normals(3000, 9)
    rng = np.random.default_rng(10)
    src, tgt = rng.standard_normal((3, 3000)), rng.standard_normal((3, 3000))
    keep, sr, tr, rm, rem = io.prefilter_reduce(a, b, src, tgt)
    keep_o, rem_o = OP.histogram_outlier_removal(a, b)
    sro, tro, rmo = OP.mask_filter(src, tgt, keep_o)
    assert np.array_equal(keep, keep_o) and rem == rem_o == sr.shape[1]
    assert np.array_equal(sr, sro) and np.array_equal(tr, tro) and np.array_equal(rm, rmo)


@pytest.mark.gpu
def test_prefilter_reduce_batch_equals_single_calls():
    sets = []
    for k, n in enumerate([3000, 17, 1200, 0, 800]):
        a, b = _normals(max(n, 1), 20 + k)
        rng = np.random.default_rng(40 + k)
        sets.append((a[:, :n], b[:, :n], rng.standard_normal((3, n)), rng.standard_normal((3, n))))
    got = io.prefilter_reduce_batch([s[0] for s in sets], [s[1] for s in sets], [s[2] for s in sets], [s[3] for s in sets])
    for (a, b, src, tgt), (keep, sr, tr, rm, rem) in zip(sets, got):
        if src.shape[1] == 0:
            assert keep.size == 0 and sr.shape[1] == 0 and rem == 0
            continue
        k1, s1, t1, r1, rem1 = io.prefilter_reduce(a, b, src, tgt)
        assert np.array_equal(keep, k1) and rem == rem1
        assert np.array_equal(sr, s1) and np.array_equal(tr, t1) and np.array_equal(rm, r1)


def test_prefilter_has_no_cpu_fallback():
    if capi.lib().psulvsb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(Exception):
        io.histogram_outlier_removal(np.ones((3, 8)), np.ones((3, 8)))


@pytest.mark.gpu
def test_mask_filter_matches_restatement():
    rng = np.random.default_rng(5)
    src, tgt = rng.standard_normal((3, 300)), rng.standard_normal((3, 300))
    keep = rng.integers(-1, 2, 300).astype(np.int32)
    sr, tr, rm = io.mask_filter(src, tgt, keep)
    sro, tro, rmo = OP.mask_filter(src, tgt, keep)
    assert np.array_equal(sr, sro) and np.array_equal(tr, tro) and np.array_equal(rm, rmo)


def _write_ply(path, pts, fmt, dtype, extra=True, faces=True):
    n = pts.shape[1]
    tname = {"f4": "float", "f8": "double"}[dtype]
    hdr = ["ply", f"format {fmt} 1.0", "comment written by tests/test_host_io.py", f"element vertex {n}"]
    if extra:
        hdr.append("property uchar red")
    hdr += [f"property {tname} x", f"property {tname} y", f"property {tname} z"]
    if extra:
        hdr.append("property float confidence")
    if faces:
        hdr += ["element face 2", "property list uchar int vertex_indices"]
    hdr.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode())
        if fmt == "ascii":
            for i in range(n):
                row = ([str(i % 256)] if extra else []) + [repr(float(v)) for v in pts[:, i]] + (["0.5"] if extra else [])
                f.write((" ".join(row) + "\n").encode())
            if faces:
                f.write(b"3 0 1 2\n3 1 2 3\n")
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            c = "f" if dtype == "f4" else "d"
            for i in range(n):
                if extra:
                    f.write(struct.pack(e + "B", i % 256))
                f.write(struct.pack(e + 3 * c, *[float(v) for v in pts[:, i]]))
                if extra:
                    f.write(struct.pack(e + "f", 0.5))
            if faces:
                f.write(struct.pack(e + "Biii", 3, 0, 1, 2) + struct.pack(e + "Biii", 3, 1, 2, 3))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("dtype", ["f4", "f8"])
def test_ply_reader_round_trip(tmp_path, fmt, dtype):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-2, 2, (3, 257))
    p = str(tmp_path / "cloud.ply")
    _write_ply(p, pts, fmt, dtype)
    got = io.read_ply_xyz(p)
    assert got.shape == (3, 257)
    assert np.array_equal(got, pts.astype(np.float32).astype(np.float64))     # PointXYZ holds floats
    with pytest.raises(capi.PsulvsbError):
        io.read_ply_xyz(str(tmp_path / "missing.ply"))


def test_correspondence_files(tmp_path):
    rng = np.random.default_rng(2)
    src, dst = rng.uniform(-3, 3, (3, 40)), rng.uniform(-3, 3, (3, 40))
    lines = [" ".join(f"{v:.6f}" for v in list(src[:, i]) + list(dst[:, i])) for i in range(40)]
    plain = tmp_path / "cloud_bin_0+cloud_bin_1@corr.txt"
    plain.write_text("\n".join(lines) + "\n")
    counted = tmp_path / "pair.txt"
    counted.write_text("40\n" + "\n".join(lines[:20]) + "\nnot a correspondence line\n" + "\n".join(lines[20:]) + "\n")
    for path in (plain, counted):
        s, d = io.read_correspondences(str(path))
        assert s.shape == (3, 40)
        assert np.allclose(s, src, atol=5e-7) and np.allclose(d, dst, atol=5e-7)
    T = np.arange(16, dtype=np.float64).reshape(4, 4) * 0.25
    gt = tmp_path / "cloud_bin_0+cloud_bin_1@GTmat.txt"
    gt.write_text("\n".join(" ".join(repr(float(v)) for v in row) for row in T) + "\n")
    assert np.array_equal(io.read_gtmat(str(gt)), T)
    log = tmp_path / "gt.log"
    log.write_text("0 1 37\n0 2 37\nbad line\n5 9 37\n")
    assert io.read_gtlog(str(log)) == [(0, 1), (0, 2), (5, 9)]
