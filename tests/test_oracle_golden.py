"""Pins the CPU oracle against the reference's own fixtures / known answers.

CPU only.  The oracle is test infrastructure; these tests are what makes it trustworthy.
"""
import numpy as np
import pytest

from oracle import oracle as O


def ang_err(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return abs(np.arccos(np.clip(c, -1, 1)))


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10.  Our counter layout is (block_lo, block_hi, event, domain).
    out = O.philox(0, 0, 0, 0)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = O.philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    # ctr = 243f6a88 85a308d3 13198a2e 03707344, key = a4093822 299f31d0
    out = O.philox(0x299F31D0A4093822, 0x03707344, 0x13198A2E, 0x85A308D3243F6A88)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_fixed_scale_inliers_bit_exact(golden):
    """K1 known answer: registration-test.cc:286-291 data, 168*167 booleans."""
    reg, meta = golden["reg"], golden["meta"]
    src, dst = reg["objectIn"], reg["sceneIn"]
    mask = O.consistency_mask(src, dst, meta["fixed_scale_beta"])
    n = src.shape[1]
    off = ~np.eye(n, dtype=bool)
    flat = mask[off]  # row-major, j != i
    gold = reg["fixed_scale_inliers"]
    assert flat.size == gold.size == 168 * 167
    assert int(gold.sum()) == 4016
    assert np.array_equal(flat, gold)
    pi, pj = O.reduced_set(src, dst, meta["fixed_scale_beta"])
    assert len(pi) == 2008
    assert np.all(pi < pj)
    assert np.all(mask[pi, pj] == 1)
    # row-major order of the upper triangle
    key = pi.astype(np.int64) * n + pj
    assert np.all(np.diff(key) > 0)


def test_scale_inliers_selector_cases(golden):
    """scale-solver-test.cc:71-130 (FixedScale)."""
    obj = golden["reg"]["objectIn"]
    beta = 2 * 1 * np.sqrt(1)
    assert O.scale_inliers(obj, obj, beta).all()
    assert not O.scale_inliers(obj, obj * 3 + 10, beta).any()
    shifted = obj.copy()
    shifted[:, 0] *= 10
    m = O.scale_inliers(obj, shifted, beta)
    assert m[0] == 0 and m[1:].all()


def test_svd3_against_lapack():
    rng = np.random.default_rng(7)
    for k in range(200):
        A = rng.standard_normal((3, 3)) * 10 ** rng.uniform(-3, 3)
        if k % 10 == 0:
            A[:, 2] = A[:, 0] * 0.5  # rank deficient
        U, S, V = O.svd3(A)
        assert np.allclose(U @ np.diag(S) @ V.T, A, rtol=0, atol=1e-13 * np.abs(A).max())
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-13)
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-13)
        assert np.all(np.diff(S) <= 0) and np.all(S >= 0)
        assert np.allclose(S, np.linalg.svd(A, compute_uv=False), rtol=1e-12, atol=1e-13 * np.abs(A).max())


def test_svd_rot_against_numpy():
    rng = np.random.default_rng(3)
    for _ in range(50):
        X = rng.standard_normal((3, 40))
        Y = rng.standard_normal((3, 40))
        W = rng.uniform(0, 1, 40)
        R = O.svd_rot(X, Y, W)
        H = (X * W) @ Y.T
        U, _, Vt = np.linalg.svd(H)
        V = Vt.T
        if np.linalg.det(U) * np.linalg.det(V) < 0:
            V[:, 2] *= -1
        assert np.allclose(R, V @ U.T, atol=1e-11)
        assert abs(np.linalg.det(R) - 1) < 1e-12


def test_gnc_tls_known_answer(golden):
    """rotation-solver-test.cc:137-251."""
    meta = golden["meta"]
    gp = meta["gnc_tls_params"]
    rng = np.random.default_rng(0)
    src = rng.uniform(-1, 1, (3, 10))
    R, inl, its, cost = O.gnc_tls(src, src, gp["noise_bound"], gp["max_iterations"], gp["gnc_factor"],
                                  gp["cost_threshold"])
    assert np.linalg.norm(R - np.eye(3)) < 1e-5
    th = 1.2345
    for Rref in (
        np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]]),
        np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]),
        np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]),
    ):
        R, *_ = O.gnc_tls(src, Rref @ src, gp["noise_bound"], gp["max_iterations"], gp["gnc_factor"],
                          gp["cost_threshold"])
        assert ang_err(Rref, R) < 1e-5
    expected_R = np.array(meta["expected_R_rotation_only"])
    s = golden["reg"]["rotation_only_src"]
    R, inl, its, cost = O.gnc_tls(s, expected_R @ s, gp["noise_bound"], gp["max_iterations"], gp["gnc_factor"],
                                  gp["cost_threshold"])
    assert ang_err(expected_R, R) < 1e-5
    assert inl.all()


def test_gnc_tls_rejects_outliers():
    rng = np.random.default_rng(11)
    K = 400
    sv = rng.uniform(-1, 1, (3, K))
    ax = np.array([0.3, -0.5, 0.8])
    ax /= np.linalg.norm(ax)
    a = 0.9
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rgt = np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx
    tv = Rgt @ sv + rng.uniform(-0.01, 0.01, (3, K))
    out = rng.permutation(K)[: K // 2]
    tv[:, out] = rng.uniform(-1, 1, (3, len(out)))
    # tight cost threshold: runs GNC to (near) binary weights
    R, inl, its, cost = O.gnc_tls(sv, tv, 0.03, 100, 1.4, 1e-9)
    assert ang_err(Rgt, R) < 0.01
    good = np.ones(K, bool)
    good[out] = False
    assert inl[good].mean() > 0.95 and inl[~good].mean() < 0.1
    # the driver's loose threshold (PSULVSB.cc:299, REG:945) stops GNC early: still a usable R
    R2, inl2, its2, _ = O.gnc_tls(sv, tv, 0.1, 100, 1.4, 0.005)
    assert its2 < its and ang_err(Rgt, R2) < 0.05


def test_translation_known_answers(golden):
    """translation-solver-test.cc:21-113.  The estimator was rewritten (max-stabbing), so the
    arbitrary-translation case is a loose pin."""
    reg, meta = golden["reg"], golden["meta"]
    v1, v2 = reg["translation_v1"], reg["translation_v2"]
    t, _ = O.tls_translation(v1, v1, 0.025)
    assert np.linalg.norm(t) < 1e-5
    for axis in range(3):
        sh = v1.copy()
        sh[axis] += 1
        t, _ = O.tls_translation(v1, sh, 0.025)
        e = np.zeros(3)
        e[axis] = 1
        assert np.linalg.norm(t - e) < 1e-5
    t, inl = O.tls_translation(v1, v2, meta["translation_noise_bound"])
    assert np.linalg.norm(t - np.array(meta["expected_t_translation"])) < 5e-3
    # max-stabbing invariant: the estimate is the mean of a maximal set of mutually
    # overlapping intervals on each axis
    nb = meta["translation_noise_bound"]
    for axis in range(3):
        x = (v2 - v1)[axis]
        depth = [(np.abs(x - c) <= 2 * nb).sum() for c in x]
        members = np.abs(x - t[axis]) <= nb + 1e-12
        assert members.sum() >= 1
        best = 0
        for xk in x:
            best = max(best, ((x >= xk) & (x <= xk + 2 * nb)).sum())
        win = [xk for xk in x if ((x >= xk) & (x <= xk + 2 * nb)).sum() == best][0]
        sel = (x >= win) & (x <= win + 2 * nb)
        assert abs(t[axis] - x[sel].mean()) < 1e-9


def test_translation_pseudo_measurement():
    x = np.array([[0.0, 0.01, 5.0, 5.01], [0] * 4, [0] * 4])
    z = np.zeros_like(x)
    t0, _ = O.tls_translation(z, x, 0.05)
    assert abs(t0[0] - 0.005) < 1e-12  # equal depth: the first (smallest) stabbing set wins (strict > at REG:181)
    # a last-best pseudo-measurement at 5.0 tips the balance (REG:136-161)
    t1, _ = O.tls_translation(z, x, 0.05, last_best=[5.0, 0.0, 0.0])
    assert abs(t1[0] - (5.0 + 5.01 + 5.0) / 3) < 1e-12


def test_scalar_tls_scale_cases(golden):
    """tls-test.cc:21-86.  Upstream's adaptive-voting TLS was replaced by a 1-D RANSAC consensus
    (REG:66-120) whose answer depends on the draws, so the upstream answers are a loose pin: whenever
    the consensus set equals the upstream inlier set the estimate must match (1e-3, as in the test);
    in every case the estimate is the inverse-variance weighted mean of its own consensus set."""
    hit = 0
    for case in golden["meta"]["tls_cases"]:
        x = np.array(case["x"], float)
        rg = np.array(case["ranges"], float)
        # line vectors with |sv| = beta/ranges and |tv| = x*|sv| give X = x, alpha = ranges (REG:402-412)
        nb = 0.5
        beta = 2 * nb
        sv = np.zeros((3, len(x)))
        tv = np.zeros((3, len(x)))
        sv[0] = beta / rg
        tv[0] = x * sv[0]
        for seed in range(8):
            est, inl, its = O.tls_scale(sv, tv, nb, 1.0, None, seed=seed, event=0)
            sel = inl.astype(bool)
            assert sel.any()
            wm = (x[sel] / rg[sel] ** 2).sum() / (1 / rg[sel] ** 2).sum()
            assert abs(est - wm) < 1e-12
            if list(inl) == case["inliers"]:
                hit += 1
                assert abs(est - case["estimate"]) < 1e-3
    assert hit >= 3


def test_inlier_probability_against_scipy():
    from scipy.special import gammainc

    for r in [0.0, 1e-4, 0.003, 0.01, 0.02, 0.035, 0.04, 0.1]:
        z = r * r / (2 * 0.01 ** 2)
        assert abs(O.inlier_probability(r, 0.01) - (1 - gammainc(1.5, z))) < 1e-13


def test_weighted_svd_and_rmse():
    rng = np.random.default_rng(5)
    M = 60
    P = rng.uniform(-1, 1, (3, M))
    th = 0.4
    Rgt = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    tgt = Rgt @ P + np.array([[0.3], [-0.2], [0.1]])
    w = rng.integers(0, 4, M)
    w[:5] = 2
    T0 = np.eye(4)
    T0[:3, :3] = np.array([[np.cos(0.35), -np.sin(0.35), 0], [np.sin(0.35), np.cos(0.35), 0], [0, 0, 1]])
    T0[:3, 3] = [0.25, -0.15, 0.05]
    T = O.weighted_svd(P, tgt, w, T0)
    assert np.allclose(T[:3, :3], Rgt, atol=1e-10)
    assert np.allclose(T[:3, 3], [0.3, -0.2, 0.1], atol=1e-10)
    mask = (w > 0).astype(np.int32)
    assert O.rmse(P, tgt, mask, T) < 1e-10
    assert O.rmse(P, tgt, mask, T0) > 1e-3
    assert np.isnan(O.rmse(P, tgt, np.zeros(M, np.int32), T))


def test_sample_without_replacement_semantics():
    idx, consumed = O.sample_without_replacement(42, O.DOMAIN_BASIC, 3, 1000, 300)
    assert len(set(idx.tolist())) == 300 and idx.min() >= 0 and idx.max() < 1000
    # replay by hand: rejection sampling over rand31 % n
    seen, out, k = set(), [], 0
    while len(out) < 300:
        r = O.rand31(42, O.DOMAIN_BASIC, 3, k) % 1000
        k += 1
        if r not in seen:
            seen.add(r)
            out.append(r)
    assert out == idx.tolist() and k == consumed
    full, _ = O.sample_without_replacement(1, O.DOMAIN_L_SAMPLED, 0, 64, 64)
    assert sorted(full.tolist()) == list(range(64))


def test_max_clique_small():
    edges = [(0, 1), (0, 2), (1, 2), (2, 3), (3, 4), (4, 5), (3, 5), (3, 6), (4, 6), (5, 6)]
    c = O.max_clique(7, edges)
    assert sorted(c.tolist()) == [3, 4, 5, 6]
    rng = np.random.default_rng(2)
    n = 60
    A = rng.random((n, n)) < 0.2
    A = np.triu(A, 1)
    planted = [3, 9, 17, 22, 31, 40, 41, 55]
    for a in planted:
        for b in planted:
            if a < b:
                A[a, b] = True
    e = np.argwhere(A)
    c = O.max_clique(n, e)
    assert len(c) >= len(planted)
    S = A | A.T
    for a in c:
        for b in c:
            if a != b:
                assert S[a, b]


def test_end_to_end_object_scene(golden):
    """registration-test.cc:229-308.  Loose pin: upstream's tolerances (0.2 rad / 0.1) were set for
    upstream TEASER++ (max clique + GNC-TLS); the PSULVSB solver maximises consensus at
    tau = 2*NOISE_BOUND*(1+C/M) and on this 168-point, ~30 %-inlier fixture lands ~0.2 rad from the
    upstream answer while explaining MORE correspondences than it.  Asserted: within 0.25 rad / 0.12 of
    the upstream answer, and at least as many inliers as the upstream transform at the same tau."""
    reg, meta = golden["reg"], golden["meta"]
    src, dst = reg["objectIn"], reg["sceneIn"]
    Rexp = np.array(meta["registration_expected_R"])
    texp = np.array(meta["registration_expected_t"])
    nb = 0.0067364
    tau = 2 * nb * 2
    exp_inl = (np.linalg.norm(dst - (Rexp @ src + texp[:, None]), axis=0) <= tau).sum()
    for seed in range(4):
        p = O.default_params(noise_bound=nb, cbar2=1.0, estimate_scaling=0, rotation_cost_threshold=1e-6,
                             inloop_noise_bound=nb, inloop_cost_threshold=1e-6, score_noise_bound=nb,
                             wallclock_cap_s=0.0, seed=seed)
        sol, tr = O.solve(p, src, dst)
        assert sol.valid == 1
        R, t = O.solution_R(sol), O.solution_t(sol)
        assert abs(np.linalg.det(R) - 1) < 1e-9
        assert ang_err(Rexp, R) < 0.25 and np.linalg.norm(t - texp) < 0.12
        ours = (np.linalg.norm(dst - (R @ src + t[:, None]), axis=0) <= tau).sum()
        assert ours >= exp_inl
        assert sol.n_reduced > 0 and sol.host_rounds >= 1


def test_rank_one_rotation_is_roundoff_determined_in_the_reference_algorithm():
    """Why no parity is defined when a basic subset is ONE line vector (tiny reduced sets, benchmark_1): H = x y^T
    has rank 1, and the null-space completion of Eigen's two-sided Jacobi sweep (restated in the oracle's svd3) is
    decided by entries of size ~ 1 ulp crossing the 2 eps threshold.  Moving the inputs by one ulp flips R = V U^T by
    O(1) in a large share of cases -- the reference's own answer changes with the compiler's FMA / vector flags
    there.  With two line vectors (rank 2) the rotation is unique and stable."""
    rng = np.random.default_rng(0)

    def flipped(k, trials=400):
        bad = 0
        for _ in range(trials):
            x, y = rng.normal(size=(3, k)), rng.normal(size=(3, k))
            r0 = O.svd_rot(x, y, np.ones(k))
            x2 = np.nextafter(x, x + rng.choice([-1.0, 1.0], x.shape))
            y2 = np.nextafter(y, y + rng.choice([-1.0, 1.0], y.shape))
            bad += int(np.abs(r0 - O.svd_rot(x2, y2, np.ones(k))).max() > 1e-6)
        return bad / trials

    assert flipped(1) > 0.2
    assert flipped(2) == 0.0 and flipped(3) == 0.0
