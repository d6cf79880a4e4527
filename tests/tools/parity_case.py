"""TEST INFRASTRUCTURE (imports the CPU oracle as the checker).  Dump the local traces (GPU vs oracle) of single problems of parity_sweep.py around their first difference.
    python tests/tools/parity_case.py 183 55 ..."""
import sys

import numpy as np

sys.path.insert(0, ".")
import psulvsb_b200  # noqa: E402,F401
from oracle import oracle as O  # noqa: E402
from psulvsb_b200 import capi, synth  # noqa: E402
sys.path.insert(0, "tests/tools")
from parity_sweep import first_difference  # noqa: E402


def main():
    want = [int(a) for a in sys.argv[1:]]
    rng = np.random.default_rng(2026)
    h = capi.Handle(0)
    for k in range(max(want) + 1):
        n = int(rng.choice([150, 300, 600, 1000, 2000, 3000]))
        ratio = float(rng.choice([0.5, 0.8, 0.9, 0.95]))
        if k not in want:
            continue
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
        so, to = O.solve(O.default_params(**kw), *args_o)
        sg, tg = h.solve(capi.default_params(**kw), capi.HostProblem(*args_o), trace_cap=4096)
        d = first_difference(tg, to)
        print(f"problem {k} kind={kind} n={n} outliers={ratio}: n_reduced {sg.n_reduced}/{so.n_reduced} first difference {d}")
        i0 = max(0, (d[0] if d else 0) - 1)
        for i in range(i0, min(i0 + 3, len(tg["local"]), len(to["local"]))):
            for name, a in (("gpu", tg["local"][i]), ("ora", to["local"][i])):
                R = np.array(a.R[:]).reshape(3, 3, order="F")
                print(f"  {name} it {a.local_iter} L {a.n_sampled_lines} basic {a.basic_choose} gnc {a.gnc_iterations} "
                      f"rot_inl {a.rot_inliers} n_rot_pts {a.n_rot_points} curr {a.curr_count} best {a.best_count} "
                      f"R0 {R[0].round(6)} t {np.array(a.t[:]).round(6)}")


if __name__ == "__main__":
    main()
