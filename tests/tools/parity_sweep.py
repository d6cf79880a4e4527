"""TEST INFRASTRUCTURE (imports the CPU oracle as the checker).  Parity sweep: the CUDA solver (C ABI) vs the CPU oracle on many random problems with the same replayed
sample stream.  A run is IDENTICAL when every integer field of every local / host trace record, the final
inlier set, and R / t / scale of every local iteration agree (R within 1e-5 rad, t within 1e-5).

Runs are reported in three groups:
  consensus     the oracle ends with a registration (>= 10 final inliers): the bar of the north star applies;
  no consensus  the oracle itself fails (0..9 inliers): every hypothesis is fitted to outliers, the basic subsets
                are 1-4 line vectors (rank-deficient Kabsch, see test_unknown_scale_benchmark_1_rank_deficient)
                and the (1.0, 1.0) round's clique is one of many maximum cliques -- no parity is defined there,
                the sweep only reports how often the runs coincide anyway;
  refused       PSULVSB_ERR_UNSUPPORTED.
Run on a GPU box:   python tests/tools/parity_sweep.py [n_problems]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
import psulvsb_b200  # noqa: E402,F401
from oracle import oracle as O  # noqa: E402
from psulvsb_b200 import capi, synth  # noqa: E402

INT_FIELDS = ["host_round", "local_iter", "n_sampled_lines", "n_sampled_points", "basic_choose", "gnc_iterations",
              "rot_inliers", "n_rot_points", "similar", "curr_count", "best_count", "local_r"]
HOST_FIELDS = ["host_round", "curr_count", "best_host", "new_corr_count", "inlier_map_size", "host_r"]


def first_difference(tg, to):
    for i, (a, b) in enumerate(zip(tg["local"], to["local"])):
        for f in INT_FIELDS:
            if getattr(a, f) != getattr(b, f):
                return i, f, getattr(a, f), getattr(b, f), a.basic_choose, a.b_rate
        Ra, Rb = np.array(a.R[:]).reshape(3, 3, order="F"), np.array(b.R[:]).reshape(3, 3, order="F")
        if synth.rotation_error(Ra, Rb) >= 1e-5:
            return i, "R", synth.rotation_error(Ra, Rb), 0.0, a.basic_choose, a.b_rate
        if np.abs(np.array(a.t[:]) - np.array(b.t[:])).max() >= 1e-5:
            return i, "t", float(np.abs(np.array(a.t[:]) - np.array(b.t[:])).max()), 0.0, a.basic_choose, a.b_rate
    return None


def sweep(n_prob, sizes=(150, 300, 600, 1000, 2000, 3000), verbose=True):
    """-> (groups, refused, differing): groups[group][kind] = [identical, total]."""
    rng = np.random.default_rng(2026)
    h = capi.Handle(0)
    groups = {"consensus": {}, "no consensus": {}}
    refused = 0
    differing = []
    max_r, max_t = 0.0, 0.0
    for k in range(n_prob):
        n = int(rng.choice(list(sizes)))
        ratio = float(rng.choice([0.5, 0.8, 0.9, 0.95]))
        kind = ["full", "prefilter", "unknown_scale"][k % 3]
        outl = "gross" if (kind == "unknown_scale" or k % 2) else "fpfh"
        pair = synth.make_pair(n, ratio, 10_000 + k, outliers=outl)
        kw = dict(noise_bound=0.05, cbar2=1.0, estimate_scaling=1 if kind == "unknown_scale" else 0,
                  rotation_cost_threshold=0.005, wallclock_cap_s=0.0, seed=k)
        if kind == "prefilter":
            pre = synth.prefilter(pair, k)
            args_o = (pre["src_reduce"], pre["dst_reduce"], pair["src"], pair["dst"], pre["keep_mask"], pre["reduce_map"])
        else:
            args_o = (pair["src"], pair["dst"])
        prob = capi.HostProblem(*args_o)
        so, to = O.solve(O.default_params(**kw), *args_o)
        try:
            sg, tg = h.solve(capi.default_params(**kw), prob, trace_cap=4096)
        except capi.PsulvsbError as e:
            refused += 1
            if verbose:
                print(f"  refused: problem {k} kind={kind} n={n}: {e}")
            continue
        grp = "consensus" if so.final_inlier_count >= 10 else "no consensus"
        diff = first_difference(tg, to)
        same = (sg.status == 0 and diff is None and len(tg["local"]) == len(to["local"]) and
                len(tg["host"]) == len(to["host"]) and
                all(getattr(a, f) == getattr(b, f) for a, b in zip(tg["host"], to["host"]) for f in HOST_FIELDS) and
                np.array_equal(tg["final_inliers"], to["final_inliers"]) and
                sg.final_inlier_count == so.final_inlier_count and sg.refined == so.refined)
        dr = synth.rotation_error(sg.R, O.solution_R(so))
        dt = float(np.abs(sg.t - O.solution_t(so)).max())
        same = same and dr < 1e-5 and dt < 1e-5
        d = groups[grp].setdefault(kind, [0, 0])
        d[0] += int(same)
        d[1] += 1
        if same:
            max_r, max_t = max(max_r, dr), max(max_t, dt)
        else:
            msg = (f"  differs [{grp}]: problem {k} kind={kind} n={n} outliers={ratio} ({outl}): iters gpu/oracle "
                   f"{sg.local_iters}/{so.local_iters}, inliers {sg.final_inlier_count}/{so.final_inlier_count}, "
                   f"dR={dr:.2e} dt={dt:.2e}; first difference (iter, field, gpu, oracle, basic_choose, b_rate): {diff}")
            differing.append((grp, k, msg))
            if verbose:
                print(msg)
    if verbose:
        print(f"largest deviation of the final transform over the identical runs: R {max_r:.3e} rad, t {max_t:.3e}")
    return groups, refused, differing


def main():
    n_prob = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    groups, refused, _ = sweep(n_prob)
    for grp, kinds in groups.items():
        a = sum(v[0] for v in kinds.values())
        b = sum(v[1] for v in kinds.values())
        print(f"{grp}: identical step by step {a}/{b}   " + "  ".join(f"{k} {v[0]}/{v[1]}" for k, v in kinds.items()))
    print(f"refused: {refused}")


if __name__ == "__main__":
    main()
