"""The reference's own sub-solver unit tests (scale-solver-test.cc, rotation-solver-test.cc,
translation-solver-test.cc, registration-test.cc:286-291) replayed against the CUDA sub-solver classes, next to
the oracle on the same inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import psulvsb_b200  # noqa: F401
    from psulvsb_b200 import capi, subsolvers

    if capi.lib().psulvsb_device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return subsolvers


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    return oracle


def ang_err(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def test_compute_tims_and_fixed_scale_inliers(S, O, golden):
    """registration-test.cc:286-291: computeTIMs + ScaleInliersSelector on objectIn / sceneIn against
    fixed_scale_inliers.csv (bit-exact, 4016 ordered = 2008 unordered consistent pairs)."""
    reg, meta = golden["reg"], golden["meta"]
    src, dst = reg["objectIn"], reg["sceneIn"]
    n = src.shape[1]
    sv, smap = S.computeTIMs(src)
    tv, tmap = S.computeTIMs(dst)
    assert sv.shape == (3, n * (n - 1) // 2) and np.array_equal(smap, tmap)
    ii, jj = np.triu_indices(n, 1)
    assert np.array_equal(smap[0], ii) and np.array_equal(smap[1], jj)
    assert np.array_equal(sv, src[:, jj] - src[:, ii])                   # bit-exact differences
    nb = meta["fixed_scale_beta"] / 2
    scale, inl = S.ScaleInliersSelector(nb, 1.0).solveForScale(sv, tv)
    assert scale == 1.0
    full = np.zeros((n, n), dtype=bool)
    full[ii, jj] = inl
    full |= full.T
    gold = reg["fixed_scale_inliers"]
    assert np.array_equal(full[~np.eye(n, dtype=bool)], gold.astype(bool))
    assert np.array_equal(inl, O.scale_inliers(sv, tv, meta["fixed_scale_beta"]).astype(bool))


def test_scale_inliers_selector_cases(S, golden):
    """scale-solver-test.cc:71-130 (FixedScale)."""
    obj = golden["reg"]["objectIn"]
    sel = S.ScaleInliersSelector(1.0, 1.0)
    assert sel.solveForScale(obj, obj)[1].all()
    assert not sel.solveForScale(obj, obj * 3 + 10)[1].any()
    shifted = obj.copy()
    shifted[:, 0] *= 10
    m = sel.solveForScale(obj, shifted)[1]
    assert not m[0] and m[1:].all()


def test_gnc_tls_rotation_known_answers(S, O, golden):
    """rotation-solver-test.cc:137-251."""
    meta = golden["meta"]
    gp = meta["gnc_tls_params"]
    P = S.GNCTLSRotationSolver.Params(max_iterations=gp["max_iterations"], cost_threshold=gp["cost_threshold"],
                                      gnc_factor=gp["gnc_factor"], noise_bound=gp["noise_bound"])
    solver = S.GNCTLSRotationSolver(P)
    rng = np.random.default_rng(0)
    src = rng.uniform(-1, 1, (3, 10))
    R, inl = solver.solveForRotation(src, src)
    assert np.linalg.norm(R - np.eye(3)) < 1e-5
    th = 1.2345
    for Rref in (
        np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]]),
        np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]),
        np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]),
    ):
        R, _ = solver.solveForRotation(src, Rref @ src)
        assert ang_err(Rref, R) < 1e-5
    expected_R = np.array(meta["expected_R_rotation_only"])
    s = golden["reg"]["rotation_only_src"]
    R, inl = solver.solveForRotation(s, expected_R @ s)
    assert ang_err(expected_R, R) < 1e-5 and inl.all()
    Ro, inl_o, its_o, cost_o = O.gnc_tls(s, expected_R @ s, gp["noise_bound"], gp["max_iterations"], gp["gnc_factor"],
                                         gp["cost_threshold"])
    assert np.abs(R - Ro).max() < 1e-9 and solver.iterations_ == its_o
    c = solver.getCostAtTermination()                      # noise-free data: the loop stops on its first pass, cost = inf
    assert c == cost_o or abs(c - cost_o) <= 1e-9 * max(1.0, abs(cost_o))


def test_gnc_tls_rotation_with_outliers_and_warm_start(S, O):
    rng = np.random.default_rng(5)
    k = 400
    sv = rng.normal(size=(3, k))
    ax = np.array([0.3, -0.5, 0.8])
    ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rt = np.eye(3) + np.sin(0.9) * K + (1 - np.cos(0.9)) * K @ K
    tv = Rt @ sv + rng.uniform(-0.01, 0.01, (3, k))
    bad = rng.permutation(k)[: k // 2]
    tv[:, bad] = rng.normal(size=(3, len(bad))) * 2
    P = S.GNCTLSRotationSolver.Params(max_iterations=100, cost_threshold=0.005, gnc_factor=1.4, noise_bound=0.1)
    solver = S.GNCTLSRotationSolver(P)
    for last in (None, Rt):
        R, inl = solver.solveForRotation(sv, tv, last_best=last)
        Ro, inl_o, its_o, _ = O.gnc_tls(sv, tv, 0.1, 100, 1.4, 0.005, R_init=last)
        assert np.abs(R - Ro).max() < 1e-9 and solver.iterations_ == its_o
        assert np.array_equal(inl, inl_o.astype(bool))
        assert ang_err(R, Rt) < 0.02


def test_translation_known_answers(S, O, golden):
    """translation-solver-test.cc:21-113 (loose pin: the fork rewrote the estimator as max-stabbing)."""
    reg, meta = golden["reg"], golden["meta"]
    v1, v2 = reg["translation_v1"], reg["translation_v2"]
    solver = S.TLSTranslationSolver(0.025, 1.0)
    t, inl = solver.solveForTranslation(v1, v1)
    assert np.linalg.norm(t) < 1e-5 and inl.all()
    for axis in range(3):
        sh = v1.copy()
        sh[axis] += 1
        t, _ = solver.solveForTranslation(v1, sh)
        e = np.zeros(3)
        e[axis] = 1
        assert np.linalg.norm(t - e) < 1e-5
    nb = meta["translation_noise_bound"]
    t, inl = S.TLSTranslationSolver(nb, 1.0).solveForTranslation(v1, v2)
    assert np.linalg.norm(t - np.array(meta["expected_t_translation"])) < 5e-3
    to, inl_o = O.tls_translation(v1, v2, nb)
    assert np.abs(t - to).max() < 1e-12
    # the reference ANDs the per-axis masks it filled (registration.cc:196-202, :457-462)
    want = np.all(np.abs((v2 - v1) - to[:, None]) <= nb, axis=0)
    assert np.array_equal(inl, want)
    # pseudo-measurement of the last best translation (registration.cc:136-161)
    x = np.array([[0.0, 0.01, 5.0, 5.01], [0] * 4, [0] * 4])
    z = np.zeros_like(x)
    s2 = S.TLSTranslationSolver(0.05, 1.0)
    assert abs(s2.solveForTranslation(z, x)[0][0] - 0.005) < 1e-12
    assert abs(s2.solveForTranslation(z, x, last_best=[5.0, 0.0, 0.0])[0][0] - (5.0 + 5.01 + 5.0) / 3) < 1e-12


def test_tls_scale_matches_oracle_on_the_same_draws(S, O, golden):
    """tls-test.cc:21-86 inputs (as line vectors giving X = x, alpha = ranges) and random line vectors: the same
    Philox draws give the same consensus set, so estimate and mask agree with the oracle."""
    for case in golden["meta"]["tls_cases"]:
        x = np.array(case["x"], float)
        rg = np.array(case["ranges"], float)
        nb = 0.5
        sv = np.zeros((3, len(x)))
        tv = np.zeros((3, len(x)))
        sv[0] = 2 * nb / rg
        tv[0] = x * sv[0]
        for seed in range(4):
            est_o, inl_o, _ = O.tls_scale(sv, tv, nb, 1.0, None, seed=seed, event=0)
            est, inl = S.TLSScaleSolver(nb, 1.0, seed=seed).solveForScale(sv, tv)
            assert abs(est - est_o) <= 1e-12 * abs(est_o) and np.array_equal(inl, inl_o.astype(bool))
    rng = np.random.default_rng(8)
    k = 3000
    sv = rng.normal(size=(3, k))
    tv = 1.7 * sv + rng.uniform(-0.01, 0.01, (3, k))
    bad = rng.permutation(k)[: int(0.7 * k)]
    tv[:, bad] = rng.normal(size=(3, len(bad))) * 3
    solver = S.TLSScaleSolver(0.02, 1.0, seed=11)
    for event, last in enumerate([None, 1.69, 1.2]):
        est_o, inl_o, _ = O.tls_scale(sv, tv, 0.02, 1.0, last, seed=11, event=event)
        est, inl = solver.solveForScale(sv, tv, last_best=last)
        assert abs(est - est_o) <= 1e-12 * abs(est_o) and np.array_equal(inl, inl_o.astype(bool))
    assert abs(est - 1.7) < 0.01
