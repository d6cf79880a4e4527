"""Host-side mirror of teaser::RobustRegistrationSolver for the PSULVSB path.

Same names and argument meaning as the reference interface (registration.h:326-832):
`RobustRegistrationSolver.Params` with the PSULVSB additions `ori_src, ori_dst, keep_mask,
reduce_map` (registration.h:469-472), `solve(src, dst)` on 3xN correspondences
(registration.cc:622), `getSolution()` (registration.h:553), `reset(params)` (registration.h:747).
All computation goes through the C ABI (libpsulvsb_b200.so); the C++ facade with the identical
class layout for the reference's C++ drivers is include/teaser/registration.h.
"""
from __future__ import annotations

import dataclasses
import enum

import numpy as np

from . import capi


class ROTATION_ESTIMATION_ALGORITHM(enum.IntEnum):  # registration.h:342-346
    GNC_TLS = 0
    FGR = 1


class INLIER_GRAPH_FORMULATION(enum.IntEnum):  # registration.h:351-355
    CHAIN = 0
    COMPLETE = 1


class INLIER_SELECTION_MODE(enum.IntEnum):  # registration.h:365-370
    PMC_EXACT = 0
    PMC_HEU = 1
    KCORE_HEU = 2
    NONE = 3


@dataclasses.dataclass
class RegistrationSolution:  # registration.h:34-41
    valid: bool = True
    scale: float = 1.0
    translation: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))
    rotation: np.ndarray = dataclasses.field(default_factory=lambda: np.eye(3))
    final_inlier_count: int = 0


@dataclasses.dataclass
class Params:  # registration.h:378-473 (defaults as there)
    noise_bound: float = 0.01
    cbar2: float = 1.0
    estimate_scaling: bool = True
    rotation_estimation_algorithm: ROTATION_ESTIMATION_ALGORITHM = ROTATION_ESTIMATION_ALGORITHM.GNC_TLS
    rotation_gnc_factor: float = 1.4
    rotation_max_iterations: int = 100
    rotation_cost_threshold: float = 1e-6
    rotation_tim_graph: INLIER_GRAPH_FORMULATION = INLIER_GRAPH_FORMULATION.CHAIN
    inlier_selection_mode: INLIER_SELECTION_MODE = INLIER_SELECTION_MODE.PMC_EXACT
    kcore_heuristic_threshold: float = 0.5
    use_max_clique: bool = True
    max_clique_exact_solution: bool = True
    max_clique_time_limit: float = 3600.0
    ori_src: np.ndarray | None = None
    ori_dst: np.ndarray | None = None
    keep_mask: np.ndarray | None = None
    reduce_map: dict | np.ndarray | None = None
    # not in the reference: key of the replayable sample stream (the reference seeds with time(NULL))
    seed: int = 0
    replay: bool = False  # True disables the 60 s wall-clock rule (registration.cc:1475)


class RobustRegistrationSolver:
    Params = Params
    ROTATION_ESTIMATION_ALGORITHM = ROTATION_ESTIMATION_ALGORITHM
    INLIER_SELECTION_MODE = INLIER_SELECTION_MODE
    INLIER_GRAPH_FORMULATION = INLIER_GRAPH_FORMULATION

    def __init__(self, params: Params | None = None, device: int = 0):
        self._handle = capi.Handle(device)
        self._solution = RegistrationSolution()
        self._raw = None
        self.reset(params if params is not None else Params())

    def reset(self, params: Params) -> None:
        self._params = params
        self._solution = RegistrationSolution()

    def getParams(self) -> Params:
        return self._params

    def _c_params(self) -> capi.Params:
        p = self._params
        if p.rotation_estimation_algorithm != ROTATION_ESTIMATION_ALGORITHM.GNC_TLS:
            raise capi.PsulvsbError(capi.ERR_UNSUPPORTED, "only GNC_TLS is on the PSULVSB path (registration.cc:1111)")
        # deprecated fields first, exactly as registration.cc:628-637 (the second one wins when both are cleared)
        mode = p.inlier_selection_mode
        if not p.use_max_clique:
            mode = INLIER_SELECTION_MODE.NONE
        if not p.max_clique_exact_solution:
            mode = INLIER_SELECTION_MODE.PMC_HEU
        return capi.default_params(
            noise_bound=p.noise_bound, cbar2=p.cbar2, estimate_scaling=int(bool(p.estimate_scaling)),
            rotation_max_iterations=p.rotation_max_iterations, rotation_gnc_factor=p.rotation_gnc_factor,
            rotation_cost_threshold=p.rotation_cost_threshold, inlier_selection_mode=int(mode),
            kcore_heuristic_threshold=p.kcore_heuristic_threshold, seed=p.seed,
            wallclock_cap_s=0.0 if p.replay else 60.0)

    def solve(self, src, dst) -> RegistrationSolution:
        """solve(src, dst) on 3xN correspondence matrices (registration.cc:622).  Like the reference,
        ori_src / ori_dst / keep_mask / reduce_map come from Params; when they are absent the
        reduced set is taken to be the full set."""
        p = self._params
        rm = p.reduce_map
        if isinstance(rm, dict):  # std::map<int,int> original -> reduced
            M = np.asarray(p.ori_src).shape[1]
            dense = np.full(M, -1, dtype=np.int32)
            for k, v in rm.items():
                dense[int(k)] = int(v)
            rm = dense
        prob = capi.HostProblem(src, dst, p.ori_src, p.ori_dst, p.keep_mask, rm)
        sol, _ = self._handle.solve(self._c_params(), prob)
        self._raw = sol
        if sol.status != capi.OK:
            self._solution = RegistrationSolution(valid=False)
            return self._solution
        self._solution = RegistrationSolution(valid=bool(sol.valid), scale=sol.scale, translation=sol.t,
                                              rotation=sol.R, final_inlier_count=sol.final_inlier_count)
        return self._solution

    def getSolution(self) -> RegistrationSolution:
        return self._solution

    @property
    def diagnostics(self):
        return self._raw
