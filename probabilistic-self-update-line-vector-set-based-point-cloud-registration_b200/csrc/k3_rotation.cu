// k3_rotation.cu -- stage 3: GNC-TLS rotation (thread-block cluster per hypothesis) and batched
// closed-form Kabsch (one warp per hypothesis).
//
// Reference: GNCTLSRotationSolver::solveForRotation (registration.cc:1563-1692) around
// teaser::utils::svdRot (utils.h:121-136).  All arithmetic is FP64: the rotation feeds
// discontinuous decisions downstream (w >= 0.5 inlier test, max-stabbing translation), so FP32
// here would break the 1e-5 parity bar (DESIGN.md "Why FP64 in stage 3").
//
// One GNC iteration is ONE pass over the K line vectors: with R_i known every thread computes
// r^2 = |tv - R_i sv|^2, adds w_{i-1} r^2 to the cost, updates the weight in closed form and
// accumulates H_{i+1} += w_i sv tv^T; the 9+1 partial sums are reduced warp -> CTA -> cluster
// through distributed shared memory in a fixed order (deterministic), and every CTA's thread 0
// turns H into R_{i+1} with a 3x3 Jacobi SVD (warm-started from the previous iteration's V).
// Line vectors (and weights) of a CTA live in its shared memory for the whole solve; the overflow
// beyond the smem capacity streams from a coalesced SoA scratch in HBM/L2 (GncJob::lv), and only
// what exceeds that too is recomputed from the points each pass.  The cluster size (1, 2, 4 or 8
// CTAs per hypothesis) is chosen by the launcher from the batch size: few registrations ->
// 8 SMs each (latency), many -> one SM each (throughput).
//
// Sleeping line vectors.  Once mu has grown, most outliers sit at weight 0 and stay there: w = 0 iff
// r >= sqrt(th1), th1 only shrinks (mu grows), and a rotation change moves a residual by at most
// |R_new - R_old|_2 |sv|.  A line vector found with w = 0 and margin m = (r - sqrt(th1)) / |sv| therefore keeps
// w = 0 -- contributing nothing to the cost (its previous weight is 0) nor to H -- until the accumulated
// drift sum |R_{i+1} - R_i|_F has grown by m.  Its weight slot then stores -(drift + m) ("asleep until the
// drift reaches this") and the pass skips it exactly; nothing is approximated.  When half of a CTA's active
// positions sleep deeply (remaining margin >= 0.02 rad) the CTA swaps them behind the active range
// [0, n_act) (a deterministic permutation, kept in GncJob::perm for the epilogue), so that later passes
// touch only the survivors, which by then fit in shared memory.  If the drift ever reaches the smallest
// parked wake-up value, the range is reopened to the whole slice.  On cfg-A (K = 22 000, 95 % outliers)
// 62 % of the line-vector evaluations of a solve disappear.
// FP64 with explicit fma(): reduction order already differs from a sequential CPU sum, so fusing
// adds no new class of deviation; the discrete decisions (r^2 vs th1/th2, w >= 0.5) are unaffected
// except within an ulp of their thresholds.
#include <cooperative_groups.h>
#include <cstdlib>
#include <cuda_runtime.h>

#include "common.cuh"
#include "engine.cuh"
#include "svd3.cuh"

namespace cg = cooperative_groups;

namespace psulvsb {

namespace {

// threads per CTA is a template parameter T: 256 (two CTAs per SM, so that one hypothesis' serial SVD phase
// overlaps another's pass) for clustered launches, 512 (one CTA per SM, twice the shared-memory cache) when
// every hypothesis runs on a single CTA (large batches)
constexpr int GNC_MAX_WARPS = 32;
#ifndef GNC_CTAS_PER_SM
#define GNC_CTAS_PER_SM 2
#endif
constexpr int GNC_NRED = 12;  // 9 H + cost + max/aux + count
constexpr double GNC_DEEP_MARGIN_DEFAULT = 0.005;  // remaining margin (rad) from which a sleeping line vector is parked

struct GncSmem {
  double part[2][GNC_NRED];           // this CTA's partial sums, double-buffered by iteration parity
  double warp_part[GNC_MAX_WARPS][GNC_NRED];
  double R[9];                        // row-major current rotation
  double total[GNC_NRED];
  double Vw[9];  // Jacobi warm start (right singular vectors of the previous solve), row-major
  // sleeping line vectors (see the kernel's header comment)
  double drift;     // sum of |R_new - R_old|_F over the rotation updates so far
  double min_wake;  // smallest wake-up drift among the line vectors parked behind n_act
  int wcnt2[2][4][GNC_MAX_WARPS];
  int n_act;        // this CTA's passes cover positions [0, n_act) of its slice
  int permuted;     // positions no longer are original indices: job.perm holds the map
  int flag;
};

__device__ __forceinline__ void load_lv(const double* __restrict__ src, const double* __restrict__ dst, uint2 e,
                                        double inv_scale, double sv[3], double tv[3]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    sv[r] = src[3 * (size_t)e.y + r] - src[3 * (size_t)e.x + r];
    // pruned_dst_tims_ *= (1 / solution_.scale)   (registration.cc:1102)
    tv[r] = (dst[3 * (size_t)e.y + r] - dst[3 * (size_t)e.x + r]) * inv_scale;
  }
}

// the same from 64-byte point records: six 16-byte read-only loads, four sectors
__device__ __forceinline__ void load_lv8(const double* __restrict__ pts8, uint2 e, double inv_scale, double sv[3],
                                         double tv[3]) {
  const double2* pa = reinterpret_cast<const double2*>(pts8 + 8 * (size_t)e.x);
  const double2* pb = reinterpret_cast<const double2*>(pts8 + 8 * (size_t)e.y);
  const double2 a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
  const double2 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
  sv[0] = b0.x - a0.x;
  sv[1] = b0.y - a0.y;
  sv[2] = b1.x - a1.x;
  tv[0] = (b1.y - a1.y) * inv_scale;  // pruned_dst_tims_ *= (1 / solution_.scale)   (registration.cc:1102)
  tv[1] = (b2.x - a2.x) * inv_scale;
  tv[2] = (b2.y - a2.y) * inv_scale;
}

// point-cache mode (one CTA per hypothesis, large batches): the CTA keeps the POINTS (48 B each, as many as
// fit) in shared memory and re-forms every line vector from its endpoint pair each pass -- 8 + 16 bytes of
// HBM traffic per line vector and pass (edge, old and new weight) instead of 48 + 16
__device__ __forceinline__ void load_lv_pc(const double* __restrict__ pc, unsigned p_cap, const double* __restrict__ src,
                                           const double* __restrict__ dst, uint2 e, double inv_scale, double sv[3],
                                           double tv[3]) {
  double sa[3], ta[3], sb[3], tb[3];
  if (e.x < p_cap) {
    const double2* p = reinterpret_cast<const double2*>(pc + 6 * (size_t)e.x);
    const double2 p0 = p[0], p1 = p[1], p2 = p[2];
    sa[0] = p0.x; sa[1] = p0.y; sa[2] = p1.x; ta[0] = p1.y; ta[1] = p2.x; ta[2] = p2.y;
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sa[r] = src[3 * (size_t)e.x + r];
      ta[r] = dst[3 * (size_t)e.x + r];
    }
  }
  if (e.y < p_cap) {
    const double2* p = reinterpret_cast<const double2*>(pc + 6 * (size_t)e.y);
    const double2 p0 = p[0], p1 = p[1], p2 = p[2];
    sb[0] = p0.x; sb[1] = p0.y; sb[2] = p1.x; tb[0] = p1.y; tb[1] = p2.x; tb[2] = p2.y;
  } else {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sb[r] = src[3 * (size_t)e.y + r];
      tb[r] = dst[3 * (size_t)e.y + r];
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    sv[r] = sb[r] - sa[r];
    tv[r] = (tb[r] - ta[r]) * inv_scale;  // pruned_dst_tims_ *= (1 / solution_.scale)   (registration.cc:1102)
  }
}

// line vector l of this CTA's slice (global index k): smem cache, else SoA scratch, else recompute
struct LvSrc {
  const double* lv_s;  // smem [7][cap]
  size_t cap;
  unsigned long long ncached;
  const double* lv_g;  // global scratch [6][lv_cap]
  unsigned long long lv_cap;
  const double* src;
  const double* dst;
  const uint2* edges;
  double inv_scale;
};
__device__ __forceinline__ void fetch_lv(const LvSrc& S, unsigned long long l, unsigned long long k, double sv[3],
                                         double tv[3]) {
  if (l < S.ncached) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sv[r] = S.lv_s[(size_t)r * S.cap + l];
      tv[r] = S.lv_s[(size_t)(3 + r) * S.cap + l];
    }
  } else if (k < S.lv_cap) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      sv[r] = S.lv_g[(size_t)r * S.lv_cap + k];
      tv[r] = S.lv_g[(size_t)(3 + r) * S.lv_cap + k];
    }
  } else {
    load_lv(S.src, S.dst, S.edges[k], S.inv_scale, sv, tv);
  }
}

__device__ __forceinline__ double residual2(const double R[9], const double sv[3], const double tv[3]) {
  const double d0 = fma(-R[2], sv[2], fma(-R[1], sv[1], fma(-R[0], sv[0], tv[0])));
  const double d1 = fma(-R[5], sv[2], fma(-R[4], sv[1], fma(-R[3], sv[0], tv[1])));
  const double d2 = fma(-R[8], sv[2], fma(-R[7], sv[1], fma(-R[6], sv[0], tv[2])));
  return fma(d2, d2, fma(d1, d1, d0 * d0));
}

// H = sm->total[0..8] -> sm->R, warm-started from / updating sm->Vw.  Out of line on purpose: one thread
// runs it, and inlining its ~600 instructions would set the register budget of the streaming loops
// around it; its operands travel through shared memory, not through local-memory arrays.
__device__ __noinline__ void svd_from_smem(GncSmem* sm) {
  double H[3][3], R[3][3], V[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      H[r][c] = sm->total[r * 3 + c];
      V[r][c] = sm->Vw[r * 3 + c];
    }
  kabsch_rotation(H, R, V);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sm->R[r * 3 + c] = R[r][c];
      sm->Vw[r * 3 + c] = V[r][c];
    }
}

// Inside the GNC loop: Newton update from the previous iteration's rotation, Jacobi SVD when it declines.
__device__ __noinline__ void rotation_from_smem(GncSmem* sm) {
  double H[3][3], R[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      H[r][c] = sm->total[r * 3 + c];
      R[r][c] = sm->R[r * 3 + c];
    }
  double d2 = 0.0;
  if (rotation_newton(H, R)) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double d = R[r][c] - sm->R[r * 3 + c];
        d2 = fma(d, d, d2);
        sm->R[r * 3 + c] = R[r][c];
      }
  } else {
    svd_from_smem(sm);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double d = R[r][c] - sm->R[r * 3 + c];
        d2 = fma(d, d, d2);
      }
  }
  // |R_new - R_old|_2 <= |.|_F, rounded up: every residual moved by at most this much times |sv|
  sm->drift += sqrt(d2) * (1.0 + 1e-9) + 1e-15;
}

// CTA-level then cluster-level sum (or max for index MAXI) of NRED values; result in sm->total.
template <int NC, int T>
__device__ __forceinline__ void cluster_reduce(GncSmem* sm, double vals[GNC_NRED], int parity, int max_index) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int i = 0; i < GNC_NRED; ++i) {
    double v = vals[i];
    if (i == max_index)
      v = warp_max(v);
    else
      v = warp_sum(v);
    if (lane == 0) sm->warp_part[wid][i] = v;
  }
  __syncthreads();
  if (tid < GNC_NRED) {
    double acc = sm->warp_part[0][tid];
    for (int w = 1; w < T / 32; ++w) {
      const double x = sm->warp_part[w][tid];
      acc = (tid == max_index) ? fmax(acc, x) : acc + x;
    }
    sm->part[parity][tid] = acc;
    if (NC == 1) sm->total[tid] = acc;  // single CTA: the partial sums are the totals (one barrier less per reduction)
  }
  if (NC > 1) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (tid < GNC_NRED) {
      double acc = 0.0;
      for (int r = 0; r < NC; ++r) {
        const GncSmem* peer = cluster.map_shared_rank(sm, r);
        const double x = peer->part[parity][tid];
        acc = (r == 0) ? x : ((tid == max_index) ? fmax(acc, x) : acc + x);
      }
      sm->total[tid] = acc;
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Tiny problems (K <= GNC_SERIAL_MAX line vectors): one thread replays GNCTLSRotationSolver::solveForRotation
// (registration.cc:1563-1692) with the reference's sequential sums and operation order (this file is built with
// -fmad=false; no fma() below), the two-sided Jacobi of svd3.cuh standing in for Eigen's.  A basic subset of ONE line
// vector makes H = w x y^T rank 1, where R = V U^T is decided by how the algorithm completes the null space from
// 1-ulp entries: only the same arithmetic reproduces the oracle's R there (and with it the rest of the run).
// ------------------------------------------------------------------------------------------
constexpr int GNC_SERIAL_MAX = 32;

__device__ __noinline__ void gnc_tls_serial(const GncJob& job_g) {
  const GncJob job = job_g;
  const int K = (int)job.K;
  double sv[GNC_SERIAL_MAX][3], tv[GNC_SERIAL_MAX][3], w[GNC_SERIAL_MAX], res[GNC_SERIAL_MAX];
  for (int k = 0; k < K; ++k) {
    load_lv(job.src, job.dst, job.edges[k], job.inv_scale, sv[k], tv[k]);
    w[k] = 1.0;
  }
  double nb2 = job.noise_bound * job.noise_bound;
  if (nb2 < 1e-16) nb2 = 1e-2;  // registration.cc:1592-1595
  double mu = 1.0, prev_cost = INFINITY, cost = INFINITY;
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  bool use_init = job.use_init != 0;
  int it_done = 0;
  for (int it = 0; it < job.max_iterations; ++it) {
    it_done = it + 1;
    if (use_init) {  // registration.cc:1617-1621
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = job.R_init[c * 3 + r];
      use_init = false;
    } else {  // svdRot, utils.h:121-136
      double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int k = 0; k < K; ++k)
        for (int r = 0; r < 3; ++r) {
          const double xw = sv[k][r] * w[k];
          for (int c = 0; c < 3; ++c) H[r][c] += xw * tv[k][c];
        }
      Svd3 d;
      svd3_two_sided(H, d);
      rotation_from_svd(d, R);
    }
    for (int k = 0; k < K; ++k) {
      const double d0 = tv[k][0] - ((R[0][0] * sv[k][0] + R[0][1] * sv[k][1]) + R[0][2] * sv[k][2]);
      const double d1 = tv[k][1] - ((R[1][0] * sv[k][0] + R[1][1] * sv[k][1]) + R[1][2] * sv[k][2]);
      const double d2 = tv[k][2] - ((R[2][0] * sv[k][0] + R[2][1] * sv[k][1]) + R[2][2] * sv[k][2]);
      res[k] = (d0 * d0 + d1 * d1) + d2 * d2;
    }
    if (it == 0) {  // registration.cc:1628-1639
      double max_residual = K > 0 ? res[0] : 0.0;
      for (int k = 1; k < K; ++k) max_residual = (max_residual < res[k]) ? res[k] : max_residual;
      mu = 1 / (2 * max_residual / nb2 - 1);
      if (mu <= 0) break;
    }
    const double th1 = (mu + 1) / mu * nb2, th2 = mu / (mu + 1) * nb2;
    cost = 0;
    for (int k = 0; k < K; ++k) {
      cost += w[k] * res[k];
      if (res[k] >= th1)
        w[k] = 0;
      else if (res[k] <= th2)
        w[k] = 1;
      else
        w[k] = sqrt(nb2 * mu * (mu + 1) / res[k]) - mu;
    }
    const double cost_diff = fabs(cost - prev_cost);
    mu = mu * job.gnc_factor;
    prev_cost = cost;
    if (cost_diff < job.cost_threshold) break;
  }
  int gf = 0;
  for (int k = 0; k < K; ++k) gf += (w[k] >= 0.5) ? 1 : 0;
  const bool all_in = gf <= 10;  // registration.cc:1685-1690
  for (int k = 0; k < K; ++k) {
    const bool in = all_in || w[k] >= 0.5;
    if (job.inliers) job.inliers[k] = in ? 1 : 0;
    if (in && job.point_flags) {
      const uint2 e = job.edges[k];
      job.point_flags[e.x] = 1;
      job.point_flags[e.y] = 1;
    }
  }
  if (job.R_out)
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) job.R_out[c * 3 + r] = R[r][c];
  if (job.info) {
    job.info[0] = it_done;
    job.info[1] = all_in ? K : gf;
    job.info[2] = 0;
    job.info[3] = 0;
  }
  if (job.cost) job.cost[0] = cost;
}

// one CTA per job; jobs above GNC_SERIAL_MAX line vectors belong to gnc_tls_kernel
__global__ void __launch_bounds__(128) gnc_tls_small_kernel(const GncJob* __restrict__ jobs) {
  const GncJob& job = jobs[blockIdx.x];
  if (!job.active || job.K > (unsigned long long)GNC_SERIAL_MAX) return;
  if (job.point_flags)
    for (int i = threadIdx.x; i < job.n_points; i += 128) job.point_flags[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) gnc_tls_serial(job);
}

template <int NC, int T, int CPS, bool PC>
__global__ void __launch_bounds__(T, CPS)
    gnc_tls_kernel(const GncJob* __restrict__ jobs, int cap_per_cta, const double GNC_DEEP_MARGIN, const int pf_steps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GncSmem* sm = reinterpret_cast<GncSmem*>(smem_raw);
  double* lv = reinterpret_cast<double*>(smem_raw + ((sizeof(GncSmem) + 15) & ~size_t(15)));
  // layout: lv[0..5][cap] = sv.xyz, tv.xyz ; lv[6][cap] = weight
  const GncJob job = jobs[blockIdx.y];
  if (!job.active) return;  // uniform over the cluster
  const long long t_kernel0 = clock64();
  const int tid = threadIdx.x;
  const unsigned rank = (NC > 1) ? cg::this_cluster().block_rank() : 0u;
  if (job.K <= (unsigned long long)GNC_SERIAL_MAX) return;  // gnc_tls_small_kernel's (uniform over the cluster)
  const unsigned long long K = job.K;
  // contiguous slice of the line vectors for this CTA
  const unsigned long long per = (K + NC - 1) / NC;
  const unsigned long long k_lo = (per * rank < K) ? per * rank : K;
  const unsigned long long k_hi = (k_lo + per < K) ? k_lo + per : K;
  const unsigned long long nloc = k_hi - k_lo;
  // PC: cap_per_cta counts cached POINTS and no line vector is cached
  const unsigned long long ncached =
      PC ? 0ull : (nloc < (unsigned long long)cap_per_cta ? nloc : (unsigned long long)cap_per_cta);
  const unsigned p_cap = PC ? (unsigned)min(cap_per_cta, job.n_points) : 0u;
  const double* __restrict__ src = job.src;
  const double* __restrict__ dst = job.dst;
  const uint2* __restrict__ edges = job.edges;
  const double* __restrict__ pts8 = job.pts8;
  double* __restrict__ gw = job.weights;  // weights of the overflow part live in global memory
  const size_t cap = (size_t)cap_per_cta;

  double nb2 = job.noise_bound * job.noise_bound;
  if (nb2 < 1e-16) nb2 = 1e-2;  // registration.cc:1592-1595

  // ---- prologue: stage line vectors, H_0 = sum sv tv^T with unit weights
  double acc[GNC_NRED];
#pragma unroll
  for (int i = 0; i < GNC_NRED; ++i) acc[i] = 0.0;
  double* __restrict__ lvg = PC ? nullptr : job.lv;
  const unsigned long long lv_cap = lvg ? job.lv_cap : 0ull;
  if (PC) {  // stage the points: (sx, sy, sz, tx, ty, tz) per point
    for (unsigned i = tid; i < p_cap; i += T) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        lv[6 * (size_t)i + r] = src[3 * (size_t)i + r];
        lv[6 * (size_t)i + 3 + r] = dst[3 * (size_t)i + r];
      }
    }
    __syncthreads();
  }
  LvSrc S;
  S.lv_s = lv;
  S.cap = cap;
  S.ncached = ncached;
  S.lv_g = lvg;
  S.lv_cap = lv_cap;
  S.src = src;
  S.dst = dst;
  S.edges = edges;
  S.inv_scale = job.inv_scale;
  // lv_ready: the sampler's emit pass already formed the line vectors k < lv_cap (inv_scale == 1 there)
  const unsigned long long n_ready = (!PC && job.lv_ready && lvg) ? ((k_hi <= lv_cap) ? nloc : (lv_cap > k_lo ? lv_cap - k_lo : 0ull)) : 0ull;
  // one line vector into its home (shared-memory cache / global SoA scratch), unit weight, H_0 += sv tv^T
  auto stage = [&](unsigned long long l, const double sv[3], const double tv[3]) {
    if (l < ncached) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        lv[(size_t)r * cap + l] = sv[r];
        lv[(size_t)(3 + r) * cap + l] = tv[r];
      }
      lv[6 * cap + l] = 1.0;
    } else {
      if (k_lo + l < lv_cap && l >= n_ready) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          __stcg(lvg + (size_t)r * lv_cap + k_lo + l, sv[r]);
          __stcg(lvg + (size_t)(3 + r) * lv_cap + k_lo + l, tv[r]);
        }
      }
      __stcg(gw + k_lo + l, 1.0);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[r * 3 + c] = fma(sv[r], tv[c], acc[r * 3 + c]);
  };
  {
    // the endpoint gathers are dependent loads (edge -> 4 points): two line vectors per step, and the edges of the
    // next step already in flight while this step's points arrive
    const uint2* __restrict__ el = edges + k_lo;
    // already formed: one coalesced read (two line vectors in flight, register-free look-ahead like the passes)
    for (unsigned long long l = tid; l < n_ready; l += T) {
      if ((tid & 3) == 0 && l + 2 * T < n_ready) {
#pragma unroll
        for (int r = 0; r < 6; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(lvg + (size_t)r * lv_cap + k_lo + l + 2 * T));
      }
      double sv[3], tv[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        sv[r] = __ldcg(lvg + (size_t)r * lv_cap + k_lo + l);
        tv[r] = __ldcg(lvg + (size_t)(3 + r) * lv_cap + k_lo + l);
      }
      stage(l, sv, tv);
    }
    unsigned long long l = n_ready + tid;
    if (l + T < nloc) {
      uint2 ea = el[l], eb = el[l + T];
      for (; l + T < nloc; l += 2 * T) {
        const unsigned long long ln = l + 2 * T;
        uint2 na = ea, nb = eb;
        if (ln + T < nloc) {
          na = el[ln];
          nb = el[ln + T];
        }
        double sa[3], ta[3], sb[3], tb[3];
        if (PC) {
          load_lv_pc(lv, p_cap, src, dst, ea, job.inv_scale, sa, ta);
          load_lv_pc(lv, p_cap, src, dst, eb, job.inv_scale, sb, tb);
        } else if (pts8) {
          load_lv8(pts8, ea, job.inv_scale, sa, ta);
          load_lv8(pts8, eb, job.inv_scale, sb, tb);
        } else {
          load_lv(src, dst, ea, job.inv_scale, sa, ta);
          load_lv(src, dst, eb, job.inv_scale, sb, tb);
        }
        stage(l, sa, ta);
        stage(l + T, sb, tb);
        ea = na;
        eb = nb;
      }
    }
    for (; l < nloc; l += T) {
      double sv[3], tv[3];
      if (PC)
        load_lv_pc(lv, p_cap, src, dst, el[l], job.inv_scale, sv, tv);
      else if (pts8)
        load_lv8(pts8, el[l], job.inv_scale, sv, tv);
      else
        load_lv(src, dst, el[l], job.inv_scale, sv, tv);
      stage(l, sv, tv);
    }
  }
  int parity = 0;
  if (tid < 9) sm->Vw[tid] = (tid % 4 == 0) ? 1.0 : 0.0;
  if (tid == 0) {
    sm->drift = 0.0;
    sm->min_wake = INFINITY;
    sm->n_act = (int)nloc;
    sm->permuted = 0;
  }
  // compaction needs every position's line vector in storage (no re-formed tail) and the index scratch
  uint32_t* __restrict__ perm = job.perm;
  const bool can_compact = !PC && perm != nullptr && k_hi <= lv_cap && nloc < 0x7FFFFFFFull;
  const bool can_sleep = can_compact && job.gnc_factor > 1.0;  // th1 must not grow; pointless without parking
  if (job.use_init) {
    if (tid < 9) sm->R[tid] = job.R_init[(tid % 3) * 3 + tid / 3];  // column-major -> row-major
    __syncthreads();
  } else {
    cluster_reduce<NC, T>(sm, acc, parity, -1);
    parity ^= 1;
    if (tid == 0) svd_from_smem(sm);
    __syncthreads();
  }

  double mu = 1.0, prev_cost = INFINITY, cost = INFINITY;
  int it_done = 0;
  long long t_svd = 0;               // cycles thread 0 spends in the 3x3 SVDs (diagnostic, info[2])
  long long t_stream = 0;            // cycles thread 0 spends in the line-vector passes (diagnostic, prof[0])
  long long sum_act = 0, n_compactions = 0, first_compaction = -1, t_compact = 0;  // diagnostics, prof[6..7]
  const long long t_start = clock64();
  bool weights_are_unit = true;
  for (int it = 0; it < job.max_iterations; ++it) {
    it_done = it + 1;
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = sm->R[i];
    if (it == 0) {
      // mu initialisation needs max r^2 first (registration.cc:1628-1639)
      double mx[GNC_NRED];
#pragma unroll
      for (int i = 0; i < GNC_NRED; ++i) mx[i] = 0.0;
      for (unsigned long long l = tid; l < nloc; l += T) {
        double sv[3], tv[3];
        if (!PC && pf_steps > 0 && (tid & 3) == 0) {  // same register-free look-ahead as in the streamed pass below
          const unsigned long long la = l + 2 * T;
          if (la >= ncached && la < nloc && k_lo + la < lv_cap) {
#pragma unroll
            for (int r = 0; r < 6; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(lvg + (size_t)r * lv_cap + k_lo + la));
          }
        }
        if (PC)
          load_lv_pc(lv, p_cap, src, dst, edges[k_lo + l], job.inv_scale, sv, tv);
        else
          fetch_lv(S, l, k_lo + l, sv, tv);
        mx[0] = fmax(mx[0], residual2(R, sv, tv));
      }
      cluster_reduce<NC, T>(sm, mx, parity, 0);
      parity ^= 1;
      const double max_residual = sm->total[0];
      mu = 1.0 / (2.0 * max_residual / nb2 - 1.0);
      if (mu <= 0.0) break;  // degenerate: residuals already tiny; weights stay 1, R stays
    }
    const double th1 = (mu + 1.0) / mu * nb2;
    const double th2 = mu / (mu + 1.0) * nb2;
    const double wnum = nb2 * mu * (mu + 1.0);
#pragma unroll
    for (int i = 0; i < GNC_NRED; ++i) acc[i] = 0.0;
    acc[10] = -INFINITY;  // reduced with max: this CTA's -(smallest wake-up drift among its deep sleepers)
    const double sqrt_wnum = sqrt(wnum);
    const double drift = sm->drift;
    const float sqrt_th1_up = sqrtf((float)th1) * 1.000001f;  // rounded up
    const int n_act = sm->n_act;
    sum_act += n_act;
    // one line vector: cost term with the previous weight, closed-form new weight, H += w sv tv^T.
    // slot >= 0: the previous weight; slot < 0: weight 0, asleep until the drift reaches -slot.  Positions inside
    // the active range are simply evaluated (a sleeper's weight comes out 0 again and its margin is refreshed).
    auto body = [&](const double sv[3], const double tv[3], double slot) -> double {
      const double r2 = residual2(R, sv, tv);
      acc[9] = fma(fmax(slot, 0.0), r2, acc[9]);  // cost uses the previous weights (registration.cc:1648)
      double wn;
      if (r2 >= th1)
        wn = 0.0;
      else if (r2 <= th2)
        wn = 1.0;
      else
        wn = fma(sqrt_wnum, rsqrt(r2), -mu);  // sqrt(eps^2 mu (mu + 1) / r^2) - mu  (registration.cc:1655)
      if (wn != 0.0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double xs = sv[r] * wn;
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[r * 3 + c] = fma(xs, tv[c], acc[r * 3 + c]);
        }
      }
      if (wn == 0.0 && can_sleep) {
        double out = slot;
        if (!(slot < 0.0 && (-slot - drift) >= GNC_DEEP_MARGIN)) {
          // not (or no longer) a deep sleeper: (re)compute the margin in units of rotation change.  A LOWER bound
          // is all that is needed, so it is formed in FP32 (two MUFU operations) and shaved by far more than the
          // FP32 rounding of r, sqrt(th1) and |sv|.  Early iterations have no zero weight and never enter here.
          const float s2f = (float)fma(sv[2], sv[2], fma(sv[1], sv[1], sv[0] * sv[0]));
          const float mf = (sqrtf((float)r2) - sqrt_th1_up) * rsqrtf(s2f) * 0.999f - 1e-6f;
          out = (mf > 0.0f && mf < 1e30f) ? -(drift + (double)mf) : 0.0;
        }
        if (out < 0.0 && (-out - drift) >= GNC_DEEP_MARGIN) {
          acc[11] += 1.0;
          acc[10] = fmax(acc[10], out);  // max of the negated wake-up drifts = -(the smallest one)
        }
        return out;
      }
      return wn;
    };
    const long long c_stream0 = clock64();
    // (a) shared-memory resident part (32-bit indices, one base pointer per component)
    {
      const int nc = n_act < (int)ncached ? n_act : (int)ncached;
      double* __restrict__ ws = lv + 6 * cap;
      int l = tid;
      for (; CPS < 3 && l + T < nc; l += 2 * T) {
        double sa[3], ta[3], sb[3], tb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          sa[r] = lv[r * (int)cap + l];
          ta[r] = lv[(3 + r) * (int)cap + l];
          sb[r] = lv[r * (int)cap + l + T];
          tb[r] = lv[(3 + r) * (int)cap + l + T];
        }
        const double wa = ws[l], wb = ws[l + T];
        ws[l] = body(sa, ta, wa);
        ws[l + T] = body(sb, tb, wb);
      }
      for (; l < nc; l += T) {
        double sa[3], ta[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          sa[r] = lv[r * (int)cap + l];
          ta[r] = lv[(3 + r) * (int)cap + l];
        }
        ws[l] = body(sa, ta, ws[l]);
      }
    }
    // (b) HBM/L2 scratch part: all loads of two line vectors in flight before the arithmetic
    const unsigned long long g_hi64 = (k_hi <= lv_cap) ? nloc : ((lv_cap > k_lo) ? lv_cap - k_lo : 0ull);
    {
      // (n_act < nloc only after a compaction, which requires g_hi64 == nloc)
      const int g_hi = n_act < (int)g_hi64 ? n_act : (int)g_hi64, nl = n_act < (int)nloc ? n_act : (int)nloc;
      const double* __restrict__ g0 = lvg + k_lo;  // component r of local line vector l: g0[r * lv_cap + l]
      double* __restrict__ gwl = gw + k_lo;
      const size_t st = (size_t)lv_cap;
      int l = (int)ncached + tid;
      // register-free look-ahead: the sectors this thread group reads `pf_ahead` steps from now are pulled into L2
      // (one lane per 32-byte sector issues the prefetch), so that the loads below see L2 latency instead of HBM's
      static_assert(T % 4 == 0, "");
      const int pf_ahead = pf_steps * (2 * T);
      for (; CPS < 3 && l + T < g_hi; l += 2 * T) {
        if (pf_steps > 0 && (tid & 3) == 0 && l + pf_ahead + T < g_hi) {
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g0 + r * st + l + pf_ahead));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g0 + r * st + l + pf_ahead + T));
          }
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gwl + l + pf_ahead));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gwl + l + pf_ahead + T));
        }
        double sa[3], ta[3], sb[3], tb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          sa[r] = __ldcg(g0 + r * st + l);
          ta[r] = __ldcg(g0 + (3 + r) * st + l);
          sb[r] = __ldcg(g0 + r * st + l + T);
          tb[r] = __ldcg(g0 + (3 + r) * st + l + T);
        }
        const double wa = __ldcg(gwl + l), wb = __ldcg(gwl + l + T);
        __stcg(gwl + l, body(sa, ta, wa));
        __stcg(gwl + l + T, body(sb, tb, wb));
      }
      for (; l < g_hi; l += T) {
        double sa[3], ta[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          sa[r] = __ldcg(g0 + r * st + l);
          ta[r] = __ldcg(g0 + (3 + r) * st + l);
        }
        __stcg(gwl + l, body(sa, ta, __ldcg(gwl + l)));
      }
      // (c) beyond the scratch capacity / point-cache mode: re-form the line vectors from the points
      if (PC) {
        // software pipeline: the (edge, weight) pairs of the NEXT two line vectors are already in flight from
        // HBM while the current two are re-formed from the cached points and processed
        const uint2* __restrict__ el = edges + k_lo;
        if (l + T < nl) {
          uint2 ea = el[l], eb = el[l + T];
          double wa = __ldcg(gwl + l), wb = __ldcg(gwl + l + T);
          for (; l + T < nl; l += 2 * T) {
            const int ln = l + 2 * T;
            const bool more = ln + T < nl;
            uint2 na = ea, nb = eb;
            double nwa = 0.0, nwb = 0.0;
            if (more) {
              na = el[ln];
              nb = el[ln + T];
              nwa = __ldcg(gwl + ln);
              nwb = __ldcg(gwl + ln + T);
            }
            double sa[3], ta[3], sb[3], tb[3];
            load_lv_pc(lv, p_cap, src, dst, ea, job.inv_scale, sa, ta);
            load_lv_pc(lv, p_cap, src, dst, eb, job.inv_scale, sb, tb);
            __stcg(gwl + l, body(sa, ta, wa));
            __stcg(gwl + l + T, body(sb, tb, wb));
            ea = na;
            eb = nb;
            wa = nwa;
            wb = nwb;
          }
        }
      }
      for (; l < nl; l += T) {
        double sa[3], ta[3];
        if (PC)
          load_lv_pc(lv, p_cap, src, dst, edges[k_lo + l], job.inv_scale, sa, ta);
        else
          load_lv(src, dst, edges[k_lo + l], job.inv_scale, sa, ta);
        gwl[l] = body(sa, ta, gwl[l]);
      }
    }
    weights_are_unit = false;
    t_stream += clock64() - c_stream0;
    cluster_reduce<NC, T>(sm, acc, parity, 10);
    parity ^= 1;
    cost = sm->total[9];
    const double cost_diff = fabs(cost - prev_cost);
    mu *= job.gnc_factor;
    prev_cost = cost;
    if (cost_diff < job.cost_threshold) break;
    // ---- park the deep sleepers behind the active range (see the header comment).  The pass counted them
    // (acc[11]) and took their smallest wake-up drift (acc[10]) with the predicate used again below, so this CTA's
    // own partial sums (before the cluster combined them) are exact.
    if (can_compact && n_act >= 512) {
      const int deep_total = (int)(sm->part[parity ^ 1][11] + 0.5);
      if (2 * deep_total >= n_act) {
        const long long c_comp0 = clock64();
        ++n_compactions;
        if (first_compaction < 0) first_compaction = it;
        const int lane = tid & 31, wid = tid >> 5;
        const unsigned lt_mask = (1u << lane) - 1u;
        double* __restrict__ ws = lv + 6 * cap;
        auto slot_at = [&](int l) -> double { return (l < (int)ncached) ? ws[l] : __ldcg(gw + k_lo + l); };
        auto is_deep = [&](double slot) -> bool { return slot < 0.0 && (-slot - drift) >= GNC_DEEP_MARGIN; };
        if (!sm->permuted)
          for (int l = tid; l < (int)nloc; l += T) perm[k_lo + l] = (uint32_t)l;
        const int new_act = n_act - deep_total;
        uint32_t* __restrict__ holes = perm + lv_cap + k_lo;
        uint32_t* __restrict__ movers = holes + (nloc + 1) / 2;
        int base_h = 0, base_m = 0;
        constexpr int E = 4;  // positions per thread and round: a quarter of the barriers
        for (int base = 0; base < n_act; base += E * T) {
          bool hole[E], mover[E];
          int nh = 0, nm = 0;
#pragma unroll
          for (int j = 0; j < E; ++j) {
            const int l = base + j * T + tid;
            const bool deep = (l < n_act) && is_deep(slot_at(l));
            hole[j] = deep && l < new_act;
            mover[j] = !deep && l >= new_act && l < n_act;
          }
          // rank order: round j before round j + 1, positions ascending inside a round
          unsigned bh[E], bm[E];
#pragma unroll
          for (int j = 0; j < E; ++j) {
            bh[j] = __ballot_sync(0xffffffffu, hole[j]);
            bm[j] = __ballot_sync(0xffffffffu, mover[j]);
            nh += __popc(bh[j]);
            nm += __popc(bm[j]);
          }
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < E; ++j) {
              sm->wcnt2[0][j][wid] = __popc(bh[j]);
              sm->wcnt2[1][j][wid] = __popc(bm[j]);
            }
          }
          __syncthreads();
          int off_h = base_h, off_m = base_m;
#pragma unroll
          for (int j = 0; j < E; ++j) {
            int ph = 0, pm = 0, th = 0, tm = 0;
            for (int w = 0; w < T / 32; ++w) {
              const int c0 = sm->wcnt2[0][j][w], c1 = sm->wcnt2[1][j][w];
              if (w < wid) {
                ph += c0;
                pm += c1;
              }
              th += c0;
              tm += c1;
            }
            const int l = base + j * T + tid;
            if (hole[j]) holes[off_h + ph + __popc(bh[j] & lt_mask)] = (uint32_t)l;
            if (mover[j]) movers[off_m + pm + __popc(bm[j] & lt_mask)] = (uint32_t)l;
            off_h += th;
            off_m += tm;
          }
          base_h = off_h;
          base_m = off_m;
          __syncthreads();
        }
        // the i-th deep sleeper inside the new range trades places with the i-th survivor behind it
        const int n_swap = base_h < base_m ? base_h : base_m;  // (equal by construction)
        // (every load of a swap is issued before its first store: with load -> store per array the seven
        // arrays cost seven global round trips per swap, 10 % of the kernel's warp time in the ncu source page)
        for (int i = tid; i < n_swap; i += T) {
          const int a = (int)holes[i], b = (int)movers[i];
          const bool a_sm = a < (int)ncached, b_sm = b < (int)ncached;
          const uint32_t qa = perm[k_lo + a], qb = perm[k_lo + b];
          double va[7], vb[7];
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            // (shared-memory cache, or the L2-level scratch the passes read with ld.cg / write with st.cg)
            const double* ga = (c < 6 ? lvg + (size_t)c * lv_cap : gw) + k_lo + a;
            const double* gb = (c < 6 ? lvg + (size_t)c * lv_cap : gw) + k_lo + b;
            va[c] = a_sm ? lv[(size_t)c * cap + a] : __ldcg(ga);
            vb[c] = b_sm ? lv[(size_t)c * cap + b] : __ldcg(gb);
          }
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            double* ga = (c < 6 ? lvg + (size_t)c * lv_cap : gw) + k_lo + a;
            double* gb = (c < 6 ? lvg + (size_t)c * lv_cap : gw) + k_lo + b;
            if (a_sm) lv[(size_t)c * cap + a] = vb[c]; else __stcg(ga, vb[c]);
            if (b_sm) lv[(size_t)c * cap + b] = va[c]; else __stcg(gb, va[c]);
          }
          perm[k_lo + a] = qb;
          perm[k_lo + b] = qa;
        }
        __syncthreads();
        if (tid == 0) {
          sm->min_wake = fmin(sm->min_wake, -sm->part[parity ^ 1][10]);
          sm->n_act = new_act;
          sm->permuted = 1;
        }
        __syncthreads();
        t_compact += clock64() - c_comp0;
      }
    }
    if (it + 1 < job.max_iterations) {
      if (tid == 0) {
        const long long c0 = clock64();
        rotation_from_smem(sm);
        if (sm->min_wake <= sm->drift) {  // a parked line vector may wake under the new rotation: reopen the range
          sm->n_act = (int)nloc;
          sm->min_wake = INFINITY;
        }
        t_svd += clock64() - c0;
      }
      __syncthreads();
    }
  }

  const long long t_loop_end = clock64();
  // ---- epilogue: inlier mask w >= 0.5 (all when <= 10), endpoint flags (registration.cc:1676-1691,
  // :1114-1155).  The stale-bit defect of the reference is resolved as "zero then set".
  double cntv[GNC_NRED];
#pragma unroll
  for (int i = 0; i < GNC_NRED; ++i) cntv[i] = 0.0;
  for (unsigned long long l = tid; l < nloc; l += T) {
    const double w = weights_are_unit ? 1.0 : ((l < ncached) ? lv[6 * cap + l] : gw[k_lo + l]);  // < 0: asleep, weight 0
    cntv[0] += (w >= 0.5) ? 1.0 : 0.0;
  }
  if (job.point_flags) {
    // zero this cluster's share of the flags before anyone sets them (cluster_reduce syncs)
    for (int i = rank * T + tid; i < job.n_points; i += NC * T) job.point_flags[i] = 0;
  }
  cluster_reduce<NC, T>(sm, cntv, parity, -1);
  parity ^= 1;
  const long long gf = (long long)(sm->total[0] + 0.5);
  const bool all_in = gf <= 10;
  const bool permuted = sm->permuted != 0;
  for (unsigned long long l = tid; l < nloc; l += T) {
    const double w = weights_are_unit ? 1.0 : ((l < ncached) ? lv[6 * cap + l] : gw[k_lo + l]);
    const bool in = all_in || (w >= 0.5);
    const unsigned long long k = k_lo + (permuted ? (unsigned long long)perm[k_lo + l] : l);  // original index
    if (job.inliers) job.inliers[k] = in ? 1 : 0;
    if (in && job.point_flags) {
      const uint2 e = edges[k];
      job.point_flags[e.x] = 1;
      job.point_flags[e.y] = 1;
    }
  }
  if (rank == 0 && tid == 0) {
    if (job.R_out) {
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) job.R_out[c * 3 + r] = sm->R[r * 3 + c];  // column-major out
    }
    if (job.info) {
      job.info[0] = it_done;
      job.info[1] = (int)(all_in ? (long long)K : gf);
      job.info[2] = (int)(t_svd >> 4);                   // SVD cycles / 16
      job.info[3] = (int)((clock64() - t_start) >> 4);  // GNC loop cycles / 16
    }
    if (job.cost) job.cost[0] = cost;
    if (job.prof) {
      job.prof[0] = t_stream;
      job.prof[1] = t_loop_end - t_start;
      job.prof[2] = t_svd;
      job.prof[3] = (long long)ncached;
      job.prof[4] = t_start - t_kernel0;     // prologue: line vectors from the points, H_0, first SVD
      job.prof[5] = clock64() - t_loop_end;  // epilogue: inlier mask, endpoint flags
      job.prof[6] = sum_act;                  // active positions of this CTA summed over the iterations
      job.prof[7] = n_compactions + 100 * (first_compaction + 1) + 10000 * t_compact;
    }
  }
  if (NC > 1) cg::this_cluster().sync();  // peers may still be reading this CTA's partial sums
}

// ------------------------------------------------------------------------------------------
// batched closed-form Kabsch: one warp per hypothesis
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    kabsch_batch_kernel(const double* __restrict__ src, const double* __restrict__ dst, const uint2* __restrict__ edges,
                        const uint32_t* __restrict__ sets, int k, unsigned long long n_hyp, double* __restrict__ Rout,
                        double* __restrict__ tout) {
  const unsigned long long h = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (h >= n_hyp) return;
  double H[9], cs[3] = {0, 0, 0}, cd[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = 0.0;
  for (int l = lane; l < k; l += 32) {
    const uint2 e = edges[sets[h * (unsigned long long)k + l]];
    double sv[3], tv[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double sa = src[3 * (size_t)e.x + r], sb = src[3 * (size_t)e.y + r];
      const double da = dst[3 * (size_t)e.x + r], db = dst[3 * (size_t)e.y + r];
      sv[r] = sb - sa;
      tv[r] = db - da;
      cs[r] += sa + sb;
      cd[r] += da + db;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) H[r * 3 + c] += sv[r] * tv[c];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] = warp_sum(H[i]);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    cs[r] = warp_sum(cs[r]);
    cd[r] = warp_sum(cd[r]);
  }
  if (lane == 0) {
    double Hm[3][3], R[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Hm[r][c] = H[r * 3 + c];
    kabsch_rotation(Hm, R);
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) Rout[h * 9 + c * 3 + r] = R[r][c];
    if (tout) {
      const double invn = 1.0 / (2.0 * (double)k);
      for (int r = 0; r < 3; ++r)
        tout[h * 3 + r] = cd[r] * invn - ((R[r][0] * cs[0] + R[r][1] * cs[1]) + R[r][2] * cs[2]) * invn;
    }
  }
}

int gnc_capacity_for(int ctas_per_sm);
size_t gnc_smem_bytes(int cap) { return ((sizeof(GncSmem) + 15) & ~size_t(15)) + (size_t)7 * cap * sizeof(double); }

template <int NC, int T, int CPS, bool PC>
int launch_gnc_nc(cudaStream_t st, const GncJob* d_jobs, int n_jobs, int cap_per_cta) {
  static bool attr_set = false;
  // shared-memory payload: 7 doubles per cached line vector, or (PC) 6 doubles per cached point
  const int max_cap = PC ? gnc_capacity_for(CPS) * 7 / 6 - 2 : gnc_capacity_for(CPS);
  if (cap_per_cta > max_cap) cap_per_cta = max_cap;
  const size_t smem = PC ? gnc_smem_bytes((cap_per_cta * 6 + 6) / 7 + 1) : gnc_smem_bytes(cap_per_cta);
  if (!attr_set) {
    PSU_CUDA(cudaFuncSetAttribute(gnc_tls_kernel<NC, T, CPS, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)gnc_smem_bytes(gnc_capacity_for(CPS))));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NC, (unsigned)n_jobs, 1);
  cfg.blockDim = dim3(T, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // (performance knobs only -- any value gives the same results; the tests shrink the margin to force wake-ups)
  const double deep_margin = debug_knobs().gnc_deep_margin > 0.0 ? debug_knobs().gnc_deep_margin : GNC_DEEP_MARGIN_DEFAULT;
  // look-ahead of the L2 prefetch in the streamed pass, in double-steps (0 = off; measured: 1 is best)
  const int pf_steps = debug_knobs().gnc_prefetch >= 0 ? debug_knobs().gnc_prefetch : 1;
  PSU_CUDA(cudaLaunchKernelEx(&cfg, gnc_tls_kernel<NC, T, CPS, PC>, d_jobs, cap_per_cta, deep_margin, pf_steps));
  return PSULVSB_OK;
}

}  // namespace

namespace {
int gnc_capacity_for(int ctas_per_sm) {
  // the CTAs resident on an SM share its 227 KB
  const size_t budget = (size_t)(220 / ctas_per_sm) * 1024;
  const size_t fixed = (sizeof(GncSmem) + 15) & ~size_t(15);
  int cap = (int)((budget - fixed) / (7 * sizeof(double)));
  cap &= ~31;
  return cap;
}
}  // namespace

int gnc_default_capacity() { return gnc_capacity_for(1); }

// CTAs per hypothesis for a batch of n_jobs: as many SMs per job as keeps the whole batch resident
int gnc_cluster_for(int n_jobs) {
  const int forced = debug_knobs().gnc_cluster;
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
  const int slots = sm_count() * GNC_CTAS_PER_SM;
  if (n_jobs * 8 <= slots) return 8;
  if (n_jobs * 4 <= slots) return 4;
  if (n_jobs * 2 <= slots) return 2;
  return 1;
}

int launch_gnc_tls(cudaStream_t st, const GncJob* d_jobs, int n_jobs, int cap_per_cta, int cluster, int max_points) {
  if (n_jobs <= 0) return PSULVSB_OK;
  (void)max_points;
  gnc_tls_small_kernel<<<n_jobs, 128, 0, st>>>(d_jobs);  // tiny subsets: the reference's arithmetic replayed by one thread
  PSU_CHECK_LAUNCH("gnc_tls_small_kernel");
  if (cap_per_cta < 32) cap_per_cta = 32;
  if (debug_knobs().gnc_cluster > 0) {
    switch (cluster) {
      case 8: return launch_gnc_nc<8, 512, 1, false>(st, d_jobs, n_jobs, cap_per_cta);
      case 4: return launch_gnc_nc<4, 512, 1, false>(st, d_jobs, n_jobs, cap_per_cta);
      case 2: return launch_gnc_nc<2, 512, 1, false>(st, d_jobs, n_jobs, cap_per_cta);
      default: return launch_gnc_nc<1, 512, 1, false>(st, d_jobs, n_jobs, cap_per_cta);
    }
  }
  switch (cluster) {
    // (measured: 512 threads x 2 CTAs per SM = 64 registers spills 1.3 KB per thread and loses 25 %)
    case 8:
      // few registrations (8 CTAs each still leave SMs idle): 512 threads, one CTA per SM -- half the line
      // vectors per thread in the latency-bound pass
      if (n_jobs * 8 <= sm_count()) return launch_gnc_nc<8, 512, 1, false>(st, d_jobs, n_jobs, cap_per_cta);
      return launch_gnc_nc<8, 256, 2, false>(st, d_jobs, n_jobs, cap_per_cta);
    case 4: return launch_gnc_nc<4, 256, 2, false>(st, d_jobs, n_jobs, cap_per_cta);
    case 2: return launch_gnc_nc<2, 256, 2, false>(st, d_jobs, n_jobs, cap_per_cta);
    // one CTA per hypothesis (large batches).  Measured and dropped: caching the points instead of the line vectors
    // (B = 256: 29.9 vs 28.1 ms per step), 2 x 256-thread CTAs per SM (-6 %), 640 / 768-thread CTAs (-1..2 %)
    case 1: return launch_gnc_nc<1, 512, 1, false>(st, d_jobs, n_jobs, cap_per_cta);
    default: return fail(PSULVSB_ERR_INVALID, "launch_gnc_tls: cluster must be 1, 2, 4 or 8");
  }
}

int launch_kabsch_batch(cudaStream_t st, const double* src, const double* dst, const uint2* edges,
                        const uint32_t* sets, int k, unsigned long long n_hyp, double* R, double* t) {
  if (n_hyp == 0) return PSULVSB_OK;
  if (k < 1) return fail(PSULVSB_ERR_INVALID, "kabsch_batch: k < 1");
  const unsigned long long threads = n_hyp * 32ull;
  kabsch_batch_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(src, dst, edges, sets, k, n_hyp, R, t);
  PSU_CHECK_LAUNCH("kabsch_batch_kernel");
  return PSULVSB_OK;
}

}  // namespace psulvsb
