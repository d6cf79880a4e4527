"""Shape of the bench workload (cfg-A: N = 5000, 95 % outliers): reduced-set size, basic-subset size K and GNC
iterations of every local iteration of a few registrations -- the numbers the GNC / sampler cost models use."""
import sys

import numpy as np

sys.path.insert(0, ".")
import psulvsb_b200  # noqa: E402,F401
from psulvsb_b200 import capi, synth  # noqa: E402


def main():
    h = capi.Handle(0)
    for seed in range(4):
        pair = synth.make_pair(5000, 0.95, 1000 + seed, outliers="fpfh" if seed % 2 == 0 else "gross")
        kw = dict(noise_bound=0.05, cbar2=1.0, estimate_scaling=0, rotation_cost_threshold=0.005, wallclock_cap_s=0.0,
                  seed=seed)
        sg, tg = h.solve(capi.default_params(**kw), capi.HostProblem(pair["src"], pair["dst"]), trace_cap=4096)
        loc = tg["local"]
        print(f"seed {seed}: n_reduced {sg.n_reduced} local_iters {sg.local_iters} host_rounds {sg.host_rounds} "
              f"inliers {sg.final_inlier_count}")
        print("   L      :", [a.n_sampled_lines for a in loc])
        print("   basic K:", [a.basic_choose for a in loc])
        print("   gnc its:", [a.gnc_iterations for a in loc])
        print("   points :", [a.n_sampled_points for a in loc])


if __name__ == "__main__":
    main()
