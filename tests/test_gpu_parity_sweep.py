"""Randomised end-to-end parity, driver-run: a 60-problem cut of tests/tools/parity_sweep.py (the full 300-problem
sweep is recorded under profiles/).  Product (C ABI, CUDA) vs CPU oracle on the same replayed sample stream, every
integer field and R, t of every local iteration, the final inlier set and the final transform.

Bar: wherever the oracle reaches a registration (>= 10 final inliers) the runs must be IDENTICAL step by step.  Failed
registrations (the oracle itself ends below 10 inliers: every hypothesis is fitted to outliers) are compared too;
since tiny basic subsets replay the reference's arithmetic and both sides return the canonical maximum clique they
coincide as well, except where a weighted Kabsch over > 32 line vectors degenerates to rank 1 in mid-GNC."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))


def test_sixty_random_problems_are_identical_step_by_step():
    import parity_sweep

    groups, refused, differing = parity_sweep.sweep(60, sizes=(150, 300, 600, 1000, 2000), verbose=False)
    assert refused == 0
    cons = groups["consensus"]
    assert sum(v[1] for v in cons.values()) >= 40                       # the cut does exercise the consensus path
    bad = [m for g, _, m in differing if g == "consensus"]
    assert not bad, "\n".join(bad)
    nocons = groups["no consensus"]
    same = sum(v[0] for v in nocons.values())
    total = sum(v[1] for v in nocons.values())
    assert total == 0 or same >= 0.8 * total, "\n".join(m for _, _, m in differing)
