"""TEST INFRASTRUCTURE.  One rank of a world-size-N run through the C ABI's own communicator (no torch.distributed):
    python tests/tools/sharded_worker.py RANK WORLD ID_FILE OUT_NPZ [N_CORR]
Rank 0 writes the NCCL unique id to ID_FILE; the others wait for it.  Every rank then
  * solves the SAME registration with psulvsb_solve_sharded (consistency rows split over the ranks) and alone,
  * scores its slice of a hypothesis batch with psulvsb_score_batch_sharded,
  * sums a vector of per-row counts over the communicator,
and writes what it got to OUT_NPZ for the parent test to compare."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import psulvsb_b200  # noqa: E402,F401
from psulvsb_b200 import capi, sharding, stages, synth  # noqa: E402


def main():
    rank, world, id_file, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 3000
    import torch

    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    h = capi.Handle(dev)
    if rank == 0:
        uid = capi.comm_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
        uid = open(id_file, "rb").read()
    h.comm_create(rank, world, uid)
    assert h.comm_world == world

    # ---- one registration, rows sharded
    pair = synth.make_pair(n, 0.9, 77, outliers="fpfh")
    kw = dict(noise_bound=0.05, cbar2=1.0, estimate_scaling=0, rotation_cost_threshold=0.005, wallclock_cap_s=0.0, seed=5)
    prob = capi.HostProblem(pair["src"], pair["dst"])
    params = capi.default_params(**kw)
    sh = h.solve_sharded(params, prob)
    alone, _ = capi.Handle(dev).solve(params, prob)

    # ---- hypothesis batch, sliced
    H = 4096
    (cs, cd), bound = stages.centre_and_bound(pair["src"], pair["dst"])
    d_src, d_dst = stages.to_device_points(pair["src"]), stages.to_device_points(pair["dst"])
    f_src, f_dst = stages.pack_points(d_src, cs), stages.pack_points(d_dst, cd)
    rng = np.random.default_rng(3)
    hyp = np.zeros((H, 12))
    for i in range(H):
        R, t = synth.random_rigid(rng)
        hyp[i, :9] = R.ravel(order="F")
        hyp[i, 9:] = t
    hyp[H // 3, :9] = pair["R"].ravel(order="F")
    hyp[H // 3, 9:] = pair["t"]
    hb, he = sharding.shard_range(H, rank, world)
    d_hyp = torch.from_numpy(hyp[hb:he].copy()).cuda()
    counts = torch.zeros(he - hb, dtype=torch.int32, device="cuda")
    best = torch.zeros(1, dtype=torch.int64, device="cuda")
    border = torch.zeros(1, dtype=torch.int64, device="cuda")
    csa = (ctypes.c_double * 3)(*[float(v) for v in cs])
    cda = (ctypes.c_double * 3)(*[float(v) for v in cd])
    st = torch.cuda.current_stream().cuda_stream
    capi.check(capi.lib().psulvsb_score_batch_sharded(h._h, st, f_src.data_ptr(), f_dst.data_ptr(), d_src.data_ptr(),
                                                      d_dst.data_ptr(), n, d_hyp.data_ptr(), he - hb, hb, 1.0, 0.04,
                                                      bound, csa, cda, counts.data_ptr(), best.data_ptr(),
                                                      border.data_ptr()))
    # ---- per-row counts: every rank contributes the rows it owns
    rows = torch.zeros(1000, dtype=torch.int32, device="cuda")
    rb, re = sharding.shard_range(1000, rank, world)
    rows[rb:re] = torch.arange(rb, re, dtype=torch.int32, device="cuda") + 1
    h.comm_allreduce_sum_u32(rows.data_ptr(), 1000, st)
    torch.cuda.synchronize()
    np.savez(out, sharded_R=np.array(sh.rotation[:]), sharded_t=np.array(sh.translation[:]),
             sharded_ints=np.array([sh.valid, sh.final_inlier_count, sh.n_reduced, sh.local_iters, sh.status]),
             alone_R=np.array(alone.rotation[:]), alone_t=np.array(alone.translation[:]),
             alone_ints=np.array([alone.valid, alone.final_inlier_count, alone.n_reduced, alone.local_iters, alone.status]),
             best=np.array([int(best.item())], dtype=np.int64), rows=rows.cpu().numpy(),
             local_best=np.array([int(counts.max().item())]))
    h.comm_destroy()
    h.close()


if __name__ == "__main__":
    main()
