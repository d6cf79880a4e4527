#!/usr/bin/env python
"""bench.py -- registrations/s of the PSULVSB hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--batch B] [--impl reference]

Workload (config.workload): BASELINE configs[1], "3DMatch-shaped synthetic fragment pair: N = 5000
FPFH-style correspondences, 95 % outliers".  One STEP = one pass of the hot path
(RobustRegistrationSolver::solve, registration.cc:622-1535) over a batch of B independent
fragment pairs of that shape per GPU (synthetic, fixed seeds).  Independent pairs shard across
GPUs with no data-path collective (weak scaling: B pairs per GPU).

  value : registrations/s, inputs resident in HBM (psulvsb_batch_solve_resident), device time
          from CUDA events on the engine's stream, max over ranks.
  e2e   : the same through psulvsb_solve_batch with HOST buffers: staging + H2D + solve + D2H of
          the solutions inside the timed region.
  roofline     : the consistency-mask kernel (stage 1), FP32-pipe bound:
                 16 issue slots per unordered pair; peak = SMs x 128 lanes x SM clock under load.
  cpu_baseline : the CPU oracle (a port of the reference's algorithm; the reference itself cannot be
                 built in this image) on a bounded sample of the same problems, 1 core.
  --impl reference : the oracle on all host cores (process pool), same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CORR = 5000
OUTLIER_RATIO = 0.95
PARAM_KW = dict(noise_bound=0.05, cbar2=1.0, estimate_scaling=0, rotation_max_iterations=100,
                rotation_gnc_factor=1.4, rotation_cost_threshold=0.005, wallclock_cap_s=0.0)
K1_SLOTS_PER_PAIR = 16
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k1_mask_kernel launch at --batch 64 from the ncu --set full
# capture profiles/r1_ncu_k1_mask_b64.txt (79.6 MB read + 128.7 MB written; the algorithmic output is the
# 64 x 5000 x 160-word mask = 204.8 MB incl. row padding and the untouched lower triangle, inputs 10 MB)
K1_TRAFFIC_BYTES_B64 = 79_636_736 + 128_738_304
# the same at --batch 256 (profiles/r1_ncu_k1_mask_b256_final.txt: 336.4 MB read + 692.0 MB written; algorithmic
# output 4 x 204.8 MB): the default batch
# and at the default --batch 296 (profiles/r1_ncu_step_b296_final.txt: 472.1 MB read + 899.6 MB written; the
# algorithmic output is 296 x 5000 x 160 words = 947 MB incl. row padding and the untouched lower triangle)
K1_TRAFFIC_BYTES = {64: K1_TRAFFIC_BYTES_B64, 256: 336_378_368 + 691_986_688, 296: 472_137_000 + 899_623_000}
# (other batch sizes: scaled from the 296 capture -- the kernel's traffic is per registration)


def make_problems(rank: int, batch: int):
    import psulvsb_b200  # noqa: F401
    from psulvsb_b200 import synth

    return [synth.make_pair(N_CORR, OUTLIER_RATIO, 1_000_000 + rank * 100_000 + i, outliers="fpfh")
            for i in range(batch)]


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs (NVML every 5 ms; nvidia-smi, which
    needs ~100 ms per query, only as a fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 6:
            self.sm.append(float(f[0]))
            self.mx.append(float(f[1]))
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.005 if self._nvml is not None else 0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def _oracle_solve_one(args):
    from oracle import oracle as O

    src, dst, seed = args
    p = O.default_params(seed=seed, **PARAM_KW)
    sol, _ = O.solve(p, src, dst, trace_cap=1)
    return sol.final_inlier_count


def cpu_time_problems(pairs, seeds, workers: int) -> float:
    """Wall time of the oracle over `pairs` with `workers` processes (1 = in-process, scalar port)."""
    from oracle import oracle as O

    O.lib()
    jobs = [(p["src"], p["dst"], s) for p, s in zip(pairs, seeds)]
    if workers <= 1:
        t0 = time.perf_counter()
        for j in jobs:
            _oracle_solve_one(j)
        return time.perf_counter() - t0
    import multiprocessing as mp

    with mp.get_context("fork").Pool(workers) as pool:
        pool.map(_oracle_solve_one, jobs[:workers])  # warm the workers
        t0 = time.perf_counter()
        pool.map(_oracle_solve_one, jobs, chunksize=1)
        return time.perf_counter() - t0


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores * 2, 8)
    pairs = make_problems(0, per_step)
    seeds = [1000 + i for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_time_problems(pairs[:cores], seeds[:cores], cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_time_problems(pairs, seeds, cores)
    value = per_step * args.steps / t
    line = {
        "impl": "reference", "metric": "registrations/s", "value": value, "unit": "registrations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg-A: N={N_CORR} correspondences, {int(OUTLIER_RATIO * 100)}% FPFH-style outliers, "
                               f"independent fragment pairs", "pairs_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "registrations/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} cfg-A pairs per step on {cores} host processes (CPU oracle: a port; the "
                                   f"reference needs Eigen3/Boost/PCL/PMC and cannot be built in this image)"},
        "e2e": {"value": value, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=296,
                    help="fragment pairs per GPU per step (default: two per SM of a 148-SM B200 -- the GNC-TLS kernel runs "
                         "one CTA per registration, so multiples of the SM count leave no partial wave)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import psulvsb_b200  # noqa: F401
    from psulvsb_b200 import capi

    if not torch.cuda.is_available() or capi.lib().psulvsb_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    pairs = make_problems(rank, B)
    probs = [capi.HostProblem(p["src"], p["dst"]) for p in pairs]
    seeds = [1000 + rank * 100_000 + i for i in range(B)]
    params = capi.default_params(**PARAM_KW)
    h = capi.Handle(local_rank)
    W = max(args.warmup, 3)

    # ---------------- resident-input throughput (value) ----------------
    h.upload(probs)
    for _ in range(W):
        sols = h.solve_resident(params, seeds)
    bad = [s.status for s in sols if s.status != 0]
    if bad:
        raise SystemExit(f"bench.py: solve failed with status {bad[:4]}")
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    l0 = h.launch_count
    dev_ms, k1_ms, gnc_ms, ticks = 0.0, 0.0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sols = h.solve_resident(params, seeds)
        dev_ms += h.last_device_ms
        k1_ms += h.last_stage_ms(2)
        gnc_ms += h.last_stage_ms(3)
        ticks += h.last_ticks
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    launches = h.launch_count - l0
    clocks = sampler.stop()
    stage = {"stage1_ms": h.last_stage_ms(0), "ticks_ms": h.last_stage_ms(1), "k1_kernel_ms": h.last_stage_ms(2),
             "gnc_kernels_ms": h.last_stage_ms(3), "refine_ms": h.last_stage_ms(4), "ticks": h.last_ticks}
    # mean basic-subset size of the step (line vectors handed to one GNC-TLS solve): the algorithmic input of that kernel
    gnc_k_mean = float(np.mean([s.n_reduced for s in sols])) * 0.1 * 0.3

    # ---------------- end to end through the C ABI with host buffers (e2e) ----------------
    for _ in range(2):
        h.solve_batch(params, probs, seeds)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sols_e2e = h.solve_batch(params, probs, seeds)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1000.0
    h2d = world * sum(p.nbytes for p in probs)  # whole job, like `value`
    d2h = world * B * __import__("ctypes").sizeof(capi.Solution)

    # max over ranks (device-timed value, wall-timed e2e)
    t = torch.tensor([dev_ms, e2e_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, wall_ms_max = [float(x) for x in t.tolist()]
    total_regs = B * args.steps * world

    if rank == 0:
        inl = [s.final_inlier_count for s in sols]
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        prop = torch.cuda.get_device_properties(local_rank)
        sms = prop.multi_processor_count
        f_mhz = clocks["sm_mhz"] or peaks.get("clocks_under_load", {}).get("sm_mhz_median") or 1965.0
        pairs_per_launch = B * (N_CORR * (N_CORR - 1) // 2)
        k1_s = (k1_ms / args.steps) / 1000.0
        achieved = pairs_per_launch * K1_SLOTS_PER_PAIR / k1_s / 1e9 if k1_s > 0 else None
        peak = sms * 128 * f_mhz * 1e6 / 1e9
        stride = ((N_CORR + 31) // 32 + 3) // 4 * 4
        mask_bytes = B * N_CORR * stride * 4
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "registrations/s", "value": total_regs / (dev_ms_max / 1000.0), "unit": "registrations/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
            "config": {"workload": f"cfg-A: N={N_CORR} correspondences, {int(OUTLIER_RATIO * 100)}% FPFH-style "
                                   f"outliers, independent fragment pairs", "pairs_per_gpu_per_step": B,
                       "l2": "no flush: per-step working set (edge arena + masks) exceeds the 126 MB L2",
                       "params": "noise_bound 0.05, cbar2 1, known scale, GNC-TLS 1.4/100/0.005, replay mode",
                       "ticks_per_step": ticks / args.steps, "mean_inliers": float(np.mean(inl)),
                       "wall_ms_per_step": wall_ms_max / args.steps, "stage_ms_last_step": stage},
            "e2e": {"value": total_regs / (e2e_ms_max / 1000.0), "unit": "registrations/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp32-pipe", "kernel": "k1_mask_kernel (line-vector length-consistency bit mask)",
                         "achieved": achieved, "peak": peak, "unit": "Gslot/s (FP32-pipe issue slots, 16 per pair)",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": K1_TRAFFIC_BYTES.get(B, int(K1_TRAFFIC_BYTES[296] * B / 296) if B > 64 else None),
                         "algorithmic_bytes": mask_bytes // 2 + 2 * B * N_CORR * 16,
                         "pairs_per_s": pairs_per_launch / k1_s if k1_s > 0 else None,
                         "kernel_ms": k1_ms / args.steps, "share_of_step": k1_ms / dev_ms if dev_ms > 0 else None,
                         "peak_source": f"{sms} SMs x 128 lanes x {f_mhz:.0f} MHz (nvidia-smi median under load)",
                         # the kernel with the largest share of the step is not an FP32-pipe kernel and has no per-unit
                         # figure in SURVEY 8(d); reported beside K1: algorithmic bytes = every line vector of every
                         # GNC-TLS solve read once (48 B) -- what a launch would move if the whole solve stayed on chip
                         "largest_kernel": {
                             "kernel": "gnc_tls_kernel (GNC-TLS rotation, FP64; one launch per tick)",
                             "share_of_step": gnc_ms / dev_ms if dev_ms > 0 else None,
                             "ms_per_launch": gnc_ms / ticks if ticks else None, "bound": "hbm",
                             "algorithmic_bytes": int(B * gnc_k_mean * 48),
                             "achieved": (B * gnc_k_mean * 48) / (gnc_ms / ticks / 1e3) / 1e9 if ticks and gnc_ms > 0 else None,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": ((B * gnc_k_mean * 48) / (gnc_ms / ticks / 1e3) / 1e9 / hbm_peak)
                             if ticks and gnc_ms > 0 else None,
                             "traffic": int((2_886_296_000 + 692_751_000) * B / 296) if B >= 148 else None,
                             "note": "traffic = dram read + write of one launch at B = 296 "
                                     "(profiles/r1_ncu_step_b296_final.txt): 12x the algorithmic bytes -- the line vectors "
                                     "beyond the shared-memory cache are re-read every GNC iteration until they are "
                                     "parked; FP64 pipe 18 % busy, long-scoreboard bound"},
                         "hbm_mask_write": {"achieved": mask_bytes / k1_s / 1e9 if k1_s > 0 else None,
                                            "peak": hbm_peak, "unit": "GB/s",
                                            "frac": (mask_bytes / k1_s / 1e9 / hbm_peak) if k1_s > 0 else None,
                                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)"
                                            if "hbm_gbs" in peaks else "fallback 6650 GB/s"}},
        }
        if not args.no_cpu_baseline:
            n_cpu = min(B, 64)
            reps = max(1, int(round(120 / n_cpu)))  # ~120 registrations ~ 11 s of single-core work
            tcpu = sum(cpu_time_problems(pairs[:n_cpu], seeds[:n_cpu], 1) for _ in range(reps))
            line["cpu_baseline"] = {"value": reps * n_cpu / tcpu, "unit": "registrations/s", "cores": 1, "kind": "port",
                                    "sample": f"first {n_cpu} pairs of rank 0's batch x {reps}, CPU oracle (scalar port "
                                              f"of registration.cc:622-1535; the reference cannot be built here), "
                                              f"{tcpu:.1f} s"}
            # result agreement on the sample (the oracle as the checker, never as the thing measured)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
