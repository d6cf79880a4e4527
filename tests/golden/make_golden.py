"""Collects the reference's own test fixtures for the PSULVSB hot path into tests/golden/.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Only DATA files and known-answer constants from the reference's tests are gathered; no reference
source code is copied.  Every fixture records where it came from (golden.json).

Sources (all under /root/reference/TEASER-plusplus/test):
  teaser/data/registration_test/{objectIn,sceneIn}.csv        3x168 correspondences
  teaser/data/registration_test/fixed_scale_inliers.csv       168*167 booleans (K1 known answer)
  teaser/data/registration_test/rotation_only_src.csv         200x3 (GNC-TLS known answer)
  teaser/data/registration_test/translation_test_v{1,2}_inliers.csv  3x34
  benchmark/data/benchmark_{1..6}/{src,dst}.ply + {R,t,s}_ref.csv + parameters.txt
  teaser/data/registration_test/bun_zipper_res3.ply           1889 bunny vertices (config 1 of BASELINE.json;
                                                              the full bun_zipper.ply is not in the reference tree)
Known-answer constants:
  rotation-solver-test.cc:232-234 expected_R; translation-solver-test.cc:110 expected_t;
  registration-test.cc:229-308 expected R,t; tls-test.cc:21-86 scalar TLS answers.
"""
import json
import os
import re

import numpy as np

REF = "/root/reference/TEASER-plusplus/test"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_csv_matrix(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip().rstrip(",")
            if line:
                rows.append([float(x) for x in line.split(",")])
    return np.array(rows, dtype=np.float64)


def read_ascii_ply(path):
    with open(path) as f:
        lines = f.read().splitlines()
    n = None
    start = None
    for i, l in enumerate(lines):
        if l.startswith("element vertex"):
            n = int(l.split()[-1])
        if l.strip() == "end_header":
            start = i + 1
            break
    pts = np.array([[float(x) for x in lines[start + k].split()[:3]] for k in range(n)])
    return pts.T.copy()  # 3xN


def main():
    d = os.path.join(REF, "teaser/data/registration_test")
    obj = read_csv_matrix(os.path.join(d, "objectIn.csv"))
    scn = read_csv_matrix(os.path.join(d, "sceneIn.csv"))
    assert obj.shape == (3, 168) and scn.shape == (3, 168)
    fixed = read_csv_matrix(os.path.join(d, "fixed_scale_inliers.csv")).astype(np.uint8).ravel()
    assert fixed.size == 168 * 167
    rot_src = read_csv_matrix(os.path.join(d, "rotation_only_src.csv"))  # 200x3
    tv1 = read_csv_matrix(os.path.join(d, "translation_test_v1_inliers.csv"))
    tv2 = read_csv_matrix(os.path.join(d, "translation_test_v2_inliers.csv"))
    np.savez_compressed(
        os.path.join(OUT, "registration_test.npz"),
        objectIn=obj, sceneIn=scn, fixed_scale_inliers=fixed,
        rotation_only_src=rot_src.T.copy(), translation_v1=tv1, translation_v2=tv2,
    )
    bench = {}
    for k in range(1, 7):
        bd = os.path.join(REF, f"benchmark/data/benchmark_{k}")
        par = open(os.path.join(bd, "parameters.txt")).read()
        nb = float(re.search(r"Noise Bound:\s*([0-9.eE+-]+)", par).group(1))
        bench[f"b{k}_src"] = read_ascii_ply(os.path.join(bd, "src.ply"))
        bench[f"b{k}_dst"] = read_ascii_ply(os.path.join(bd, "dst.ply"))
        bench[f"b{k}_R"] = read_csv_matrix(os.path.join(bd, "R_ref.csv"))
        bench[f"b{k}_t"] = read_csv_matrix(os.path.join(bd, "t_ref.csv")).ravel()
        bench[f"b{k}_s"] = read_csv_matrix(os.path.join(bd, "s_ref.csv")).ravel()
        bench[f"b{k}_noise_bound"] = np.array([nb])
    np.savez_compressed(os.path.join(OUT, "benchmark.npz"), **bench)
    meta = {
        "generated_by": "tests/golden/make_golden.py",
        "source_root": REF,
        "fixed_scale_beta": 2 * 0.0067364,
        "fixed_scale_note": "registration-test.cc:286-291 (commented-out check) against getScaleInliersMask(); "
                            "ordered pairs (i, j != i) row-major; noise_bound 0.0067364, cbar2 1",
        "gnc_tls_params": {"max_iterations": 100, "cost_threshold": 1e-12, "gnc_factor": 1.4, "noise_bound": 1e-3,
                           "source": "rotation-solver-test.cc:239"},
        "expected_R_rotation_only": [[0.997379773225804, -0.019905935977315, -0.069551000516966],
                                     [0.013777311189888, 0.996068297974922, -0.087510750572249],
                                     [0.071019530105605, 0.086323226782879, 0.993732623426126]],
        "expected_t_translation": [-0.098430131086161, 0.008679113091532, 0.197317864174211],
        "translation_noise_bound": 0.00673642835,
        "registration_expected_R": [[0.9974, -0.0199, -0.0696], [0.0138, 0.9961, -0.0875], [0.0710, 0.0863, 0.9937]],
        "registration_expected_t": [-0.1011, 0.0908, 0.1344],
        "registration_tolerance": {"rot_rad": 0.2, "trans": 0.1, "source": "registration-test.cc:304-307"},
        "tls_cases": [
            {"x": [0.5, 1, 0.6, 0.7, 1.2], "ranges": [0.9, 0.9, 0.4, 0.5, 0.4], "estimate": 0.8383,
             "inliers": [1, 1, 1, 1, 1]},
            {"x": [0.5, 1, 0.6, 0.7, 1.2, 10], "ranges": [0.9, 0.9, 0.4, 0.5, 0.4, 0.5], "estimate": 0.8383,
             "inliers": [1, 1, 1, 1, 1, 0]},
            {"x": [0.5, 1, 0.6, 20, 16, 10], "ranges": [0.9, 0.9, 0.4, 0.5, 0.4, 0.5], "estimate": 0.6425,
             "inliers": [1, 1, 1, 0, 0, 0]},
        ],
    }
    bunny = read_ascii_ply(os.path.join(REF, "teaser", "data", "registration_test", "bun_zipper_res3.ply"))
    np.savez_compressed(os.path.join(OUT, "bunny_res3.npz"), vertices=bunny.astype(np.float32))
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
