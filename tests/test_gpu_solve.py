"""End-to-end GPU parity: psulvsb_solve (C ABI, host buffers) vs the CPU oracle on the same inputs and
the same replayed Philox sample stream, compared step by step through the per-iteration traces.

Bars (north_star): inlier sets / counts / control flow bit-exact; R within 1e-5 rad, t within 1e-5.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

R_TOL = 1e-5   # rad  (north_star)
T_TOL = 1e-5   # units (north_star)


@pytest.fixture(scope="module")
def env():
    import psulvsb_b200  # noqa: F401
    from oracle import oracle
    from psulvsb_b200 import capi, synth

    if capi.lib().psulvsb_device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback")
    return {"capi": capi, "synth": synth, "O": oracle, "h": capi.Handle(0)}


PKW = dict(noise_bound=0.05, cbar2=1.0, estimate_scaling=0, rotation_cost_threshold=0.005, wallclock_cap_s=0.0)


def both(env, pair, pre=None, seed=0, **kw):
    capi, O = env["capi"], env["O"]
    args = dict(PKW)
    args.update(kw)
    po = O.default_params(seed=seed, **args)
    pg = capi.default_params(seed=seed, **args)
    if pre is None:
        so, to = O.solve(po, pair["src"], pair["dst"])
        prob = capi.HostProblem(pair["src"], pair["dst"])
    else:
        so, to = O.solve(po, pre["src_reduce"], pre["dst_reduce"], pair["src"], pair["dst"], pre["keep_mask"],
                         pre["reduce_map"])
        prob = capi.HostProblem(pre["src_reduce"], pre["dst_reduce"], pair["src"], pair["dst"], pre["keep_mask"],
                                pre["reduce_map"])
    sg, tg = env["h"].solve(pg, prob, trace_cap=4096)
    return so, to, sg, tg


def assert_same_run(env, so, to, sg, tg):
    synth, O = env["synth"], env["O"]
    assert sg.status == 0
    assert sg.n_line_vectors == so.n_line_vectors and sg.n_reduced == so.n_reduced
    assert len(tg["local"]) == len(to["local"]) and len(tg["host"]) == len(to["host"])
    int_fields = ["host_round", "local_iter", "n_sampled_lines", "n_sampled_points", "basic_choose", "gnc_iterations",
                  "rot_inliers", "n_rot_points", "similar", "curr_count", "best_count", "local_r"]
    for a, b in zip(tg["local"], to["local"]):
        for f in int_fields:
            assert getattr(a, f) == getattr(b, f), (f, a.local_iter, getattr(a, f), getattr(b, f))
        assert a.l_rate == b.l_rate and a.b_rate == b.b_rate
        assert abs(a.p_local - b.p_local) < 1e-12
        Ra, Rb = np.array(a.R[:]).reshape(3, 3, order="F"), np.array(b.R[:]).reshape(3, 3, order="F")
        assert synth.rotation_error(Ra, Rb) < R_TOL
        assert np.abs(np.array(a.t[:]) - np.array(b.t[:])).max() < T_TOL
    for a, b in zip(tg["host"], to["host"]):
        for f in ["host_round", "curr_count", "best_host", "new_corr_count", "inlier_map_size", "host_r"]:
            assert getattr(a, f) == getattr(b, f), (f, getattr(a, f), getattr(b, f))
        assert abs(a.p_host - b.p_host) < 1e-12
    assert bool(sg.valid) == bool(so.valid)
    assert sg.final_inlier_count == so.final_inlier_count
    assert sg.host_rounds == so.host_rounds and sg.local_iters == so.local_iters
    assert sg.final_C == so.final_C and sg.escalations == so.escalations and sg.refined == so.refined
    assert np.array_equal(tg["final_inliers"], to["final_inliers"])      # inlier set bit-exact
    assert np.array_equal(tg["inlier_counter"], to["inlier_counter"])
    assert synth.rotation_error(sg.R, O.solution_R(so)) < R_TOL
    assert np.abs(sg.t - O.solution_t(so)).max() < T_TOL
    assert sg.scale == so.scale


@pytest.mark.parametrize("n,ratio,outl,seed", [(200, 0.5, "gross", 1), (500, 0.8, "fpfh", 2), (1000, 0.9, "fpfh", 3),
                                               (1000, 0.9, "gross", 4), (2000, 0.95, "fpfh", 5)])
def test_solve_matches_oracle_full_set(env, n, ratio, outl, seed):
    pair = env["synth"].make_pair(n, ratio, seed, outliers=outl)
    so, to, sg, tg = both(env, pair, seed=seed)
    assert_same_run(env, so, to, sg, tg)
    assert env["synth"].rotation_error(sg.R, pair["R"]) < 0.05


@pytest.mark.parametrize("n,ratio,seed", [(600, 0.7, 11), (1500, 0.9, 12), (3000, 0.9, 13)])
def test_solve_matches_oracle_with_self_update(env, n, ratio, seed):
    """keep_mask pre-filter emulation: the working set grows through the probabilistic self-update."""
    synth = env["synth"]
    pair = synth.make_pair(n, ratio, seed)
    pre = synth.prefilter(pair, seed)
    so, to, sg, tg = both(env, pair, pre, seed=seed)
    assert_same_run(env, so, to, sg, tg)
    assert sg.final_C >= pre["src_reduce"].shape[1]


@pytest.mark.parametrize("n,ratio,side,seed", [(20_000, 0.95, 3.0, 21), (100_000, 0.99, 30.0, 11)])
def test_large_n_matches_oracle_end_to_end(env, n, ratio, side, seed):
    """BASELINE configs[2] as a REGISTRATION: N = 100 000 correspondences, 99 % outliers (5.0e9 line vectors, which
    the reference's int pair indices, registration.cc:682-686, cannot hold; the oracle streams them with 64-bit
    indices, ~35 s of CPU), and N = 20 000.  Same step-by-step bar as the small cases: reduced set, sample sizes,
    GNC-TLS iterations, inlier sets, control flow bit-exact; R, t within 1e-5."""
    pair = env["synth"].make_pair(n, ratio, 4242, side=side)
    so, to, sg, tg = both(env, pair, seed=seed)
    assert sg.n_line_vectors == n * (n - 1) // 2
    assert_same_run(env, so, to, sg, tg)
    assert sg.final_inlier_count >= int(0.9 * n * (1 - ratio))
    assert env["synth"].rotation_error(sg.R, pair["R"]) < 0.01


def test_solve_cfg_a_5k_95pct(env):
    """BASELINE config[1]: N = 5000 correspondences, 95 % outliers."""
    pair = env["synth"].make_pair(5000, 0.95, 101)
    so, to, sg, tg = both(env, pair, seed=101)
    assert_same_run(env, so, to, sg, tg)
    assert sg.final_inlier_count >= 240
    assert env["synth"].rotation_error(sg.R, pair["R"]) < 0.01


def test_solve_golden_registration_test(env, golden):
    """registration-test.cc:229-308 known answer on objectIn/sceneIn, asserted exactly as the oracle's own
    pin (tests/test_oracle_golden.py::test_end_to_end_object_scene: 0.25 rad / 0.12 of the upstream
    answer, at least as many inliers as it), plus step-by-step parity with the oracle."""
    capi, O = env["capi"], env["O"]
    reg, meta = golden["reg"], golden["meta"]
    src, dst = reg["objectIn"], reg["sceneIn"]
    Rexp = np.array(meta["registration_expected_R"]).reshape(3, 3)
    texp = np.array(meta["registration_expected_t"])
    nb = 0.0067364
    tau = 2 * nb * 2
    exp_inl = (np.linalg.norm(dst - (Rexp @ src + texp[:, None]), axis=0) <= tau).sum()
    for seed in range(4):
        kw = dict(noise_bound=nb, cbar2=1.0, estimate_scaling=0, rotation_cost_threshold=1e-6,
                  inloop_noise_bound=nb, inloop_cost_threshold=1e-6, score_noise_bound=nb, wallclock_cap_s=0.0)
        so, to, sg, tg = both(env, {"src": src, "dst": dst}, seed=seed, **kw)
        assert_same_run(env, so, to, sg, tg)
        assert sg.valid and sg.status == 0
        assert env["synth"].rotation_error(sg.R, Rexp) < 0.25 and np.linalg.norm(sg.t - texp) < 0.12
        ours = (np.linalg.norm(dst - (sg.R @ src + sg.t[:, None]), axis=0) <= tau).sum()
        assert ours >= exp_inl


def test_batch_equals_individual_solves(env):
    capi, synth = env["capi"], env["synth"]
    pairs = [synth.make_pair(n, 0.9, 50 + i) for i, n in enumerate([300, 800, 1200, 500, 64])]
    probs = [capi.HostProblem(p["src"], p["dst"]) for p in pairs]
    seeds = [7, 8, 9, 10, 11]
    params = capi.default_params(**PKW)
    batch = env["h"].solve_batch(params, probs, seeds)
    for prob, seed, b in zip(probs, seeds, batch):
        params.seed = seed
        one, _ = env["h"].solve(params, prob)
        assert b.status == 0 and one.status == 0
        assert np.array_equal(np.array(b.rotation[:]), np.array(one.rotation[:]))     # deterministic: bit-equal
        assert np.array_equal(np.array(b.translation[:]), np.array(one.translation[:]))
        assert b.final_inlier_count == one.final_inlier_count and b.local_iters == one.local_iters


def test_chunked_pool_gives_the_same_results(env):
    """psulvsb_set_batching: chunks advancing concurrently on the handle's engine pool (host buffers: dynamic hand-out;
    resident: even split) return exactly what one lock-step batch returns, default seeds included."""
    capi, synth = env["capi"], env["synth"]
    pairs = [synth.make_pair(n, 0.9, 150 + i) for i, n in enumerate([300, 800, 1200, 500, 64, 700, 333, 900, 410, 256, 777])]
    probs = [capi.HostProblem(p["src"], p["dst"]) for p in pairs]
    params = capi.default_params(seed=21, **PKW)
    h = capi.Handle(0)
    h.set_batching(64, 1)
    ref = h.solve_batch(params, probs)
    assert h.last_chunk_ticks == [h.last_ticks]

    def same(a, b):
        for x, y in zip(a, b):
            assert x.status == 0 and y.status == 0
            assert np.array_equal(np.array(x.rotation[:]), np.array(y.rotation[:]))
            assert np.array_equal(np.array(x.translation[:]), np.array(y.translation[:]))
            assert (x.final_inlier_count, x.local_iters, x.n_reduced) == (y.final_inlier_count, y.local_iters, y.n_reduced)

    for chunk, lanes in ((3, 2), (4, 3), (2, 4)):
        h.set_batching(chunk, lanes)
        same(ref, h.solve_batch(params, probs))
        assert len(h.last_chunk_ticks) == (len(probs) + chunk - 1) // chunk
        assert h.resident_size == 0  # a chunked host-buffer batch leaves nothing resident
        h.upload(probs)
        assert h.resident_size == len(probs)
        same(ref, h.solve_resident(params))
        assert len(h.last_chunk_ticks) >= min(lanes, (len(probs) + chunk - 1) // chunk)  # engines x their sub-batches
        assert h.last_device_ms > 0 and h.last_stage_ms(2) > 0 and h.last_stage_ms(3) > 0
    seeds = list(range(100, 100 + len(probs)))
    h.set_batching(64, 1)
    ref = h.solve_batch(params, probs, seeds)
    h.set_batching(3, 2)
    same(ref, h.solve_batch(params, probs, seeds))
    # an invalid problem in a later chunk fails the call with its message on the calling thread
    bad = capi.HostProblem(pairs[0]["src"].copy(), pairs[0]["dst"].copy())
    bad.src[0, 0] = np.nan
    with pytest.raises(capi.PsulvsbError, match="non-finite"):
        h.solve_batch(params, probs[:7] + [bad])
    h.close()


def test_pipelined_batches_match_solve_batch(env):
    """psulvsb_batch_submit / psulvsb_batch_wait: several batches in flight on one handle (the lanes pull chunks from one
    queue across calls, so a later batch's upload overlaps an earlier batch's solve) return what psulvsb_solve_batch
    returns for each of them; tickets may be waited for out of order; a failing batch fails its own wait only; the
    other entry points first let the queue drain."""
    capi, synth = env["capi"], env["synth"]
    batches = []
    for k in range(4):
        pairs = [synth.make_pair(n, 0.9, 400 + 10 * k + i) for i, n in enumerate([300, 640, 150, 900, 420, 64, 510][: 4 + k])]
        batches.append([capi.HostProblem(p["src"], p["dst"]) for p in pairs])
    params = capi.default_params(seed=33, **PKW)
    h = capi.Handle(0)
    h.set_batching(64, 1)
    refs = [h.solve_batch(params, b) for b in batches]

    def same(a, b):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x.status == 0 and y.status == 0
            assert np.array_equal(np.array(x.rotation[:]), np.array(y.rotation[:]))
            assert np.array_equal(np.array(x.translation[:]), np.array(y.translation[:]))
            assert (x.final_inlier_count, x.local_iters, x.n_reduced) == (y.final_inlier_count, y.local_iters, y.n_reduced)

    for chunk, lanes in ((64, 2), (3, 2), (2, 3)):
        h.set_batching(chunk, lanes)
        tickets = [h.submit(params, b) for b in batches]
        for k in (2, 0, 3, 1):
            same(refs[k], h.wait(tickets[k]))
        assert h.last_device_ms > 0
        with pytest.raises(capi.PsulvsbError, match="unknown ticket"):
            h.wait(tickets[0])
    # seeds, a failing batch between two good ones, and a synchronous call that has to wait for the queue
    seeds = list(range(7, 7 + len(batches[1])))
    h.set_batching(64, 1)
    ref_seeded = h.solve_batch(params, batches[1], seeds)
    h.set_batching(3, 2)
    bad = capi.HostProblem(batches[0][0].src.copy(), batches[0][0].dst.copy())
    bad.src[0, 0] = np.nan
    t_a = h.submit(params, batches[1], seeds)
    t_bad = h.submit(params, batches[2][:3] + [bad])
    t_b = h.submit(params, batches[3])
    same(refs[0], h.solve_batch(params, batches[0]))  # drains the queue first
    with pytest.raises(capi.PsulvsbError, match="non-finite"):
        h.wait(t_bad)
    same(ref_seeded, h.wait(t_a))
    same(refs[3], h.wait(t_b))
    h.close()


def test_resident_solve_is_repeatable(env):
    capi, synth = env["capi"], env["synth"]
    probs = [capi.HostProblem(*(lambda p: (p["src"], p["dst"]))(synth.make_pair(700, 0.9, 70 + i))) for i in range(3)]
    h = env["h"]
    h.upload(probs)
    params = capi.default_params(seed=5, **PKW)
    a = h.solve_resident(params)
    b = h.solve_resident(params)
    for x, y in zip(a, b):
        assert np.array_equal(np.array(x.rotation[:]), np.array(y.rotation[:]))
        assert x.final_inlier_count == y.final_inlier_count
    assert h.launch_count > 0 and h.last_device_ms > 0


def test_degenerate_inputs(env):
    capi = env["capi"]
    h = env["h"]
    # no consistent pair at all: invalid solution, no hang (the reference would spin forever)
    rng = np.random.default_rng(0)
    src = rng.uniform(-1, 1, (3, 40))
    dst = src * 50.0
    sol, _ = h.solve(capi.default_params(**PKW), capi.HostProblem(src, dst))
    assert sol.status == 0 and not sol.valid and sol.n_reduced == 0
    # identical clouds, identity transform
    sol, _ = h.solve(capi.default_params(**PKW), capi.HostProblem(src, src.copy()))
    assert sol.status == 0 and sol.valid
    assert env["synth"].rotation_error(sol.R, np.eye(3)) < 1e-6 and np.abs(sol.t).max() < 1e-6


def test_mirror_solver_api(env):
    """RobustRegistrationSolver mirror: Params, solve(src, dst), getSolution() as the reference's drivers use them."""
    from psulvsb_b200 import RobustRegistrationSolver

    synth = env["synth"]
    pair = synth.make_pair(800, 0.8, 5)
    pre = synth.prefilter(pair, 5)
    params = RobustRegistrationSolver.Params()
    params.noise_bound = 0.05
    params.cbar2 = 1
    params.estimate_scaling = False
    params.rotation_max_iterations = 100
    params.rotation_gnc_factor = 1.4
    params.rotation_estimation_algorithm = RobustRegistrationSolver.ROTATION_ESTIMATION_ALGORITHM.GNC_TLS
    params.rotation_cost_threshold = 0.005
    params.ori_src, params.ori_dst = pair["src"], pair["dst"]
    params.keep_mask = pre["keep_mask"]
    params.reduce_map = {int(o): int(r) for o, r in enumerate(pre["reduce_map"]) if r >= 0}
    params.replay = True
    solver = RobustRegistrationSolver(params)
    solver.solve(pre["src_reduce"], pre["dst_reduce"])
    sol = solver.getSolution()
    assert sol.valid
    assert synth.rotation_error(sol.rotation, pair["R"]) < 0.05
    assert np.linalg.norm(sol.translation - pair["t"]) < 0.05


# ---------------------------------------------------------------------------------------------
# unknown scale (Params::estimate_scaling = true): ratio histogram reduced set (registration.cc:687-752)
# and the TLS scale estimate per local iteration (registration.cc:958-983)
# ---------------------------------------------------------------------------------------------
def scaled_pair(synth, n, ratio, seed, scale):
    pair = synth.make_pair(n, ratio, seed, outliers="gross")
    rng = np.random.default_rng(seed + 999)
    inl = pair["inlier_mask"]
    dst = pair["dst"].copy()
    # inliers: q = s R p + t + noise (the generator's noise is kept, the clean part is rescaled)
    clean = pair["R"] @ pair["src"] + pair["t"][:, None]
    dst[:, inl] = scale * (pair["R"] @ pair["src"][:, inl]) + pair["t"][:, None] + (dst[:, inl] - clean[:, inl])
    pair = dict(pair)
    pair["dst"] = np.asfortranarray(dst)
    return pair


@pytest.mark.parametrize("n,ratio,seed,scale", [(300, 0.5, 21, 1.7), (800, 0.8, 22, 0.6), (1500, 0.9, 23, 2.5)])
def test_unknown_scale_matches_oracle(env, n, ratio, seed, scale):
    pair = scaled_pair(env["synth"], n, ratio, seed, scale)
    so, to, sg, tg = both(env, pair, seed=seed, estimate_scaling=1)
    assert sg.status == 0
    assert sg.n_reduced == so.n_reduced                      # three-bin reduced set, same size ...
    assert len(tg["local"]) == len(to["local"])
    for a, b in zip(tg["local"], to["local"]):               # ... and the same run, step by step
        for f in ["host_round", "n_sampled_lines", "n_sampled_points", "basic_choose", "gnc_iterations", "rot_inliers",
                  "n_rot_points", "similar", "curr_count", "best_count", "local_r"]:
            assert getattr(a, f) == getattr(b, f), (f, a.local_iter, getattr(a, f), getattr(b, f))
        assert abs(a.scale - b.scale) <= 1e-12 * abs(b.scale)
    assert sg.final_inlier_count == so.final_inlier_count
    assert np.array_equal(tg["final_inliers"], to["final_inliers"])
    assert abs(sg.scale - so.scale) <= 1e-12 * abs(so.scale)
    assert env["synth"].rotation_error(sg.R, env["O"].solution_R(so)) < R_TOL
    assert np.abs(sg.t - env["O"].solution_t(so)).max() < T_TOL
    assert abs(sg.scale - scale) < 0.02 * scale and env["synth"].rotation_error(sg.R, pair["R"]) < 0.05


def test_unknown_scale_histogram_growth_matches_oracle(env):
    """A length ratio above MaxScale = 10000 makes the reference grow MaxScale and its histogram in the middle of
    the pair loop (registration.cc:714-718): every LATER pair is binned on the new, coarser grid.  Two nearly
    coincident source pairs (ratios ~1e4 and ~1e5, the second later in pair order) exercise two growth steps."""
    pair = scaled_pair(env["synth"], 800, 0.8, 31, 1.3)
    out = np.flatnonzero(~pair["inlier_mask"])
    src = pair["src"].copy()
    src[:, out[40]] = src[:, out[3]] + np.array([3e-4, 0, 0])
    src[:, out[90]] = src[:, out[70]] + np.array([0, 3e-5, 0])
    pair["src"] = np.asfortranarray(src)
    so, to, sg, tg = both(env, pair, seed=5, estimate_scaling=1)
    assert sg.status == 0
    assert sg.n_reduced == so.n_reduced and len(tg["local"]) == len(to["local"])
    for a, b in zip(tg["local"], to["local"]):
        for f in ["n_sampled_lines", "n_sampled_points", "basic_choose", "gnc_iterations", "rot_inliers",
                  "n_rot_points", "similar", "curr_count", "best_count", "local_r"]:
            assert getattr(a, f) == getattr(b, f), (f, a.local_iter, getattr(a, f), getattr(b, f))
    assert np.array_equal(tg["final_inliers"], to["final_inliers"])
    assert abs(sg.scale - so.scale) <= 1e-12 * abs(so.scale)
    assert env["synth"].rotation_error(sg.R, env["O"].solution_R(so)) < R_TOL


def test_unknown_scale_benchmark_1_rank_deficient(env, golden):
    """benchmark_1 has 10 points: |L_sampled| = 4 and the basic subset is ONE line vector, so H = sv tv^T has
    rank 1 and R = V U^T is decided by the SVD's null-space completion.  Tiny subsets replay the reference's
    arithmetic on the device (k3_rotation.cu gnc_tls_serial + the two-sided Jacobi of svd3.cuh), so the run is
    identical to the oracle's step by step -- and it meets the fixture's ground-truth scale."""
    capi, O = env["capi"], env["O"]
    b = golden["bench"]
    nb = float(b["b1_noise_bound"][0])
    kw = dict(noise_bound=nb, cbar2=1.0, estimate_scaling=1, rotation_cost_threshold=0.005, wallclock_cap_s=0.0,
              inloop_noise_bound=nb, score_noise_bound=nb)
    for seed in (1, 2, 3, 4):
        so, to = O.solve(O.default_params(seed=seed, **kw), b["b1_src"], b["b1_dst"])
        sg, tg = env["h"].solve(capi.default_params(seed=seed, **kw), capi.HostProblem(b["b1_src"], b["b1_dst"]),
                                trace_cap=4096)
        assert_same_run(env, so, to, sg, tg)
        assert any(rec.basic_choose == 1 for rec in tg["local"])      # the rank-1 case did occur
    sg, _ = env["h"].solve(capi.default_params(seed=1, **kw), capi.HostProblem(b["b1_src"], b["b1_dst"]))
    assert sg.status == 0 and sg.valid
    assert abs(sg.scale - float(b["b1_s"][0])) < 0.02 * float(b["b1_s"][0])


@pytest.mark.parametrize("k", [4, 6])
def test_unknown_scale_reference_benchmark_fixtures(env, golden, k):
    """TEASER-plusplus/test/benchmark/data/benchmark_{1,4,6}: ground truth R_ref / s_ref of the reference's
    own mini benchmarks (unknown scale).  The fork reports the translation in the q = s (R p + t) convention
    (registration.cc:1250 divides by the scale), so t_ref is compared with s * t."""
    capi, O = env["capi"], env["O"]
    b = golden["bench"]
    src, dst = b[f"b{k}_src"], b[f"b{k}_dst"]
    nb = float(b[f"b{k}_noise_bound"][0])
    kw = dict(noise_bound=nb, cbar2=1.0, estimate_scaling=1, rotation_cost_threshold=0.005, wallclock_cap_s=0.0,
              inloop_noise_bound=nb, score_noise_bound=nb)
    for seed in (1, 2):
        so, to = O.solve(O.default_params(seed=seed, **kw), src, dst)
        sg, tg = env["h"].solve(capi.default_params(seed=seed, **kw), capi.HostProblem(src, dst), trace_cap=4096)
        assert sg.status == 0 and bool(sg.valid) == bool(so.valid)
        assert sg.n_reduced == so.n_reduced and sg.local_iters == so.local_iters
        assert sg.final_inlier_count == so.final_inlier_count
        assert abs(sg.scale - so.scale) <= 1e-10 * abs(so.scale)
        assert env["synth"].rotation_error(sg.R, O.solution_R(so)) < R_TOL
        assert np.abs(sg.t - O.solution_t(so)).max() < T_TOL
        assert abs(sg.scale - float(b[f"b{k}_s"][0])) < 0.02 * float(b[f"b{k}_s"][0])
        assert env["synth"].rotation_error(sg.R, b[f"b{k}_R"]) < 0.02


@pytest.mark.parametrize("n", [2, 3, 4, 7, 12, 25])
def test_tiny_problems_terminate(env, n):
    """Degenerate sizes: |L_sampled| and the basic subset collapse to 0 or 1 line vectors (rank-deficient
    rotation solves, empty translation sets).  No parity is defined there (see benchmark_1), but the engine must
    terminate with a status, never hang or fault, and agree with the oracle on the sizes it derives."""
    capi, O = env["capi"], env["O"]
    rng = np.random.default_rng(n)
    src = rng.uniform(-1, 1, (3, n))
    ang = 0.4
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    dst = R @ src + np.array([[0.2], [-0.1], [0.3]]) + rng.uniform(-0.005, 0.005, (3, n))
    for scaling in (0, 1):
        kw = dict(PKW)
        kw["estimate_scaling"] = scaling
        sg, _ = env["h"].solve(capi.default_params(seed=1, **kw), capi.HostProblem(src, dst))
        so, _ = O.solve(O.default_params(seed=1, **kw), src, dst)
        assert sg.status in (0, capi.ERR_INTERNAL)           # ERR_INTERNAL = the reference's non-terminating loop, cut
        assert sg.n_line_vectors == n * (n - 1) // 2 == so.n_line_vectors
        assert sg.n_reduced == so.n_reduced
        if sg.status == 0 and sg.valid:
            Rg = sg.R
            assert np.allclose(Rg @ Rg.T, np.eye(3), atol=1e-9) and np.linalg.det(Rg) > 0


def test_invalid_problems_are_refused(env):
    capi = env["capi"]
    h = env["h"]
    p = capi.default_params(**PKW)
    with pytest.raises(capi.PsulvsbError) as ei:                     # a single correspondence has no line vector
        h.solve(p, capi.HostProblem(np.zeros((3, 1)), np.zeros((3, 1))))
    assert ei.value.code == capi.ERR_INVALID
    bad = np.zeros((3, 10))
    bad[1, 3] = np.nan
    with pytest.raises(capi.PsulvsbError) as ei:                     # non-finite coordinates
        h.solve(p, capi.HostProblem(bad, np.ones((3, 10))))
    assert ei.value.code == capi.ERR_INVALID
    with pytest.raises(capi.PsulvsbError):                           # nothing uploaded
        capi.Handle(0).solve_resident(p)
    # a failed re-upload leaves NOTHING resident (not the previous batch size over a half-overwritten arena)
    hh = capi.Handle(0)
    good = env["synth"].make_pair(200, 0.5, 2)
    hh.upload([capi.HostProblem(good["src"], good["dst"])] * 3)
    assert hh.resident_size == 3
    with pytest.raises(capi.PsulvsbError):
        hh.upload([capi.HostProblem(good["src"], good["dst"]), capi.HostProblem(bad, np.ones((3, 10)))])
    assert hh.resident_size == 0
    with pytest.raises(capi.PsulvsbError):
        hh.solve_resident(p)
    # solve() on the handle replaces the resident batch; solve_resident follows the library's count
    hh.upload([capi.HostProblem(good["src"], good["dst"])] * 3)
    hh.solve(p, capi.HostProblem(good["src"], good["dst"]))
    assert hh.resident_size == 1 and len(hh.solve_resident(p)) == 1
    # keep_mask == 1 with a reduce_map entry that is not a column of src / dst would index out of bounds on the device
    km = np.ones(200, dtype=np.int32)
    rm = np.arange(200, dtype=np.int32)
    rm[17] = -1
    with pytest.raises(capi.PsulvsbError) as ei:
        h.solve(p, capi.HostProblem(good["src"], good["dst"], good["src"], good["dst"], km, rm))
    assert ei.value.code == capi.ERR_INVALID
    rm[17] = 200
    with pytest.raises(capi.PsulvsbError) as ei:
        h.solve(p, capi.HostProblem(good["src"], good["dst"], good["src"], good["dst"], km, rm))
    assert ei.value.code == capi.ERR_INVALID
    # the handle is still usable afterwards
    pair = env["synth"].make_pair(300, 0.5, 1)
    sol, _ = h.solve(p, capi.HostProblem(pair["src"], pair["dst"]))
    assert sol.status == 0 and sol.valid


def test_keep_mask_minus_one_points_are_scored_but_never_adopted(env):
    """keep_mask == -1 (bins far from the histogram peak, PSULVSB.cc:152-157): counted as inliers of the final
    score when they fit, never appended by the self-update (registration.cc:1428-1434)."""
    capi, synth, O = env["capi"], env["synth"], env["O"]
    pair = synth.make_pair(900, 0.8, 33)
    pre = synth.prefilter(pair, 33, keep_inlier=0.5, keep_outlier=0.3, discard_outlier=0.6)
    assert (pre["keep_mask"] == -1).sum() > 50
    so, to, sg, tg = both(env, pair, pre, seed=33)
    assert_same_run(env, so, to, sg, tg)
    adopted = sg.final_C - pre["src_reduce"].shape[1]
    assert adopted <= int((pre["keep_mask"] == 0).sum())
