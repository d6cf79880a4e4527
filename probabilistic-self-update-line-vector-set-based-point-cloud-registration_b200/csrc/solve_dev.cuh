// solve_dev.cuh -- block-wide FP64 device routines shared by the stage kernels (k4_score.cu) and the
// lock-step batch engine (engine_kernels.cu): block scan / reductions, residual, max-stabbing
// translation.
#pragma once

#include "common.cuh"

namespace psulvsb {

constexpr int BLK = 1024;  // all block-wide routines below assume blockDim.x == BLK

struct BlockScratch {
  double d[32][4];
  int i[32][2];
  double bd[4];
  int bi[2];
  unsigned long long u[32];
  unsigned long long bu;
};

// exclusive block scan of a pair of small ints (packed 32+32); returns exclusive prefixes and totals
__device__ __forceinline__ void block_scan2(BlockScratch* s, int a, int b, int& ea, int& eb, int& ta, int& tb) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned long long v = ((unsigned long long)(unsigned)a << 32) | (unsigned)b;
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s->u[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    unsigned long long t = s->u[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long x = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += x;
    }
    s->u[lane] = ti - t;
    if (lane == 31) s->bu = ti;
  }
  __syncthreads();
  const unsigned long long ex = s->u[wid] + (incl - v);
  const unsigned long long tot = s->bu;
  ea = (int)(ex >> 32);
  eb = (int)(ex & 0xffffffffull);
  ta = (int)(tot >> 32);
  tb = (int)(tot & 0xffffffffull);
  __syncthreads();
}

// block sum of up to 4 doubles (fixed order: lanes by xor-shuffle, warps ascending) -> every thread
template <int N>
__device__ __forceinline__ void block_sum(BlockScratch* s, double v[N]) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double x = warp_sum(v[k]);
    if (lane == 0) s->d[wid][k] = x;
  }
  __syncthreads();
  if (tid < N) {
    double acc = 0.0;
    for (int w = 0; w < BLK / 32; ++w) acc += s->d[w][tid];
    s->bd[tid] = acc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = s->bd[k];
  __syncthreads();
}

__device__ __forceinline__ int block_sum_int2(BlockScratch* s, int& a, int& b) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int xa = warp_sum_int(a), xb = warp_sum_int(b);
  if (lane == 0) {
    s->i[wid][0] = xa;
    s->i[wid][1] = xb;
  }
  __syncthreads();
  if (tid < 2) {
    int acc = 0;
    for (int w = 0; w < BLK / 32; ++w) acc += s->i[w][tid];
    s->bi[tid] = acc;
  }
  __syncthreads();
  a = s->bi[0];
  b = s->bi[1];
  __syncthreads();
  return a;
}

// | q - s (R p + t) | exactly as the reference forms it (registration.cc:1303-1308, :1417-1423):
// the entries of (s * TRANSFORM) first, then the 4-term row products, no fused multiply-add.
// R row-major.
__device__ __forceinline__ double residual_ref(const double* __restrict__ p, const double* __restrict__ q, double s,
                                               const double R[9], const double t[3]) {
  double d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double x = dadd(dadd(dadd(dmul(dmul(s, R[r * 3 + 0]), p[0]), dmul(dmul(s, R[r * 3 + 1]), p[1])),
                               dmul(dmul(s, R[r * 3 + 2]), p[2])),
                          dmul(dmul(s, t[r]), 1.0));
    d[r] = dsub(q[r], x);
  }
  return sqrt(sqnorm3(d[0], d[1], d[2]));
}

// ------------------------------------------------------------------------------------------
// Max-stabbing translation (TLSTranslationSolver, registration.cc:436-463, with the rewritten
// ScalarTLSEstimator translation branch, :121-203), one CTA.
//
// The reference sorts the 2N interval endpoints x_k -+ sigma and sweeps, keeping the mean of the
// open set at the first closing endpoint whose depth strictly exceeds every earlier one.  The
// depth seen at the closing endpoint of k is  #{ m : lo_m <= hi_k  and  hi_m >= hi_k }, so the
// sweep's answer is the candidate k maximising that depth (ties -> smallest hi_k) and the mean of
// its member set -- computed here without the sort: one warp per candidate, lanes over members.
//
// idx[0..P): compacted indices of the participating points (ascending), xs: scratch 3*(P+1)
// doubles (lo/hi are recomputed), last_best: NULL or the pseudo-measurement (registration.cc:136-161).
// Returns the estimate of each axis in t_out (unchanged for P == 0 without pseudo-measurement).
// ------------------------------------------------------------------------------------------
struct StabBest {
  int depth;
  double hi;
  int k;
};

// TLSScaleSolver::solveForScale -> ScalarTLSEstimator::estimate, scale branch (registration.cc:397-415, :66-120) on
// K ratios X with per-item bounds A: 1-D RANSAC (candidate = X[rand % K], consensus = |X_j - X_ran| <= A_j, stop when
// 1 - (1 - best / K)^it >= 0.99), four candidates evaluated per pass (their draws do not depend on the outcome; the
// confidence rule is then applied in order), then the inverse-variance weighted mean of the consensus set (:104-119).
// use_last: the last best scale is the first candidate (:75-86).  *est_out: the unrefined estimate (the pruning of
// :966-983 uses it), *scale_out: the refined one (fallback: init_scale semantics of the caller are kept by passing
// the value to keep when K == 0 -- not reached here, K > 0).  The draws are philox (seed; DOMAIN_SCALE, event, k).
__device__ inline void block_tls_scale(BlockScratch* scratch, const double* __restrict__ X, const double* __restrict__ A,
                                       int K, uint64_t seed, uint32_t event, bool use_last, double last_s,
                                       double init_scale, double* est_out, double* scale_out) {
  __shared__ double est_s;
  __shared__ int best_s, iter_s, done_s;
  __shared__ unsigned long long k_s;
  constexpr int G = 4;
  const int tid = threadIdx.x;
  if (tid == 0) {
    est_s = init_scale;
    best_s = 0;
    iter_s = 0;
    done_s = 0;
    k_s = 0ull;
  }
  __syncthreads();
  if (use_last) {
    const double s0 = last_s;
    int c = 0, dummy = 0;
    for (int j = tid; j < K; j += BLK) c += (fabs(dsub(X[j], s0)) <= A[j]) ? 1 : 0;
    block_sum_int2(scratch, c, dummy);
    if (tid == 0) {
      iter_s = 1;
      best_s = c;
      est_s = s0;
      const double conf = 1.0 - pow(1.0 - ((double)c / (double)K), 1);
      done_s = conf < 0.99 ? 0 : 1;
    }
    __syncthreads();
  }
  while (!done_s) {
    const unsigned long long k0 = k_s;
    double xr[G];
#pragma unroll
    for (int g = 0; g < G; ++g) xr[g] = X[philox_rand31(seed, PSULVSB_DOMAIN_SCALE, event, k0 + g) % (uint32_t)K];
    double cnt[4] = {0, 0, 0, 0};
    for (int j = tid; j < K; j += BLK) {
      const double xj = X[j], aj = A[j];
#pragma unroll
      for (int g = 0; g < G; ++g) cnt[g] += (fabs(dsub(xj, xr[g])) <= aj) ? 1.0 : 0.0;
    }
    block_sum<4>(scratch, cnt);
    if (tid == 0) {
      int used = 0;
      for (int g = 0; g < G && !done_s; ++g) {
        ++used;
        iter_s += 1;
        const int c = (int)(cnt[g] + 0.5);
        if (c > best_s) {
          best_s = c;
          est_s = xr[g];
        }
        const double conf = 1.0 - pow(1.0 - ((double)best_s / (double)K), iter_s);
        if (!(conf < 0.99) || iter_s > 100000) done_s = 1;
      }
      k_s = k0 + (unsigned long long)used;
    }
    __syncthreads();
  }
  const double est = est_s;
  double sums[4] = {0, 0, 0, 0};
  for (int i = tid; i < K; i += BLK)
    if (fabs(dsub(X[i], est)) <= A[i]) {
      const double a2 = dmul(A[i], A[i]);
      sums[0] += 1.0 / a2;
      sums[1] += X[i] / a2;
    }
  block_sum<4>(scratch, sums);
  double scale = est;
  if (sums[0] == sums[0] && sums[1] == sums[1]) scale = sums[1] / sums[0];
  __syncthreads();
  if (tid == 0) {
    *est_out = est;
    *scale_out = scale;
  }
  __syncthreads();
}


// Largest point sets (a rotation solve that found no consensus flags every endpoint; cfg-B) make the all-pairs count
// quadratic on ONE CTA: 5000 points took 3.6 ms, on which the whole lock-step batch waited.  Above BT_SORT_MIN
// measurements the axis is sorted instead (bitonic network in `sorted`, length a power of two >= N, padded with +inf):
// x -> fl(x - sigma) and x -> fl(x + sigma) are monotone, so the members of a candidate are a contiguous run of the
// sorted order, found by two binary searches with the SAME rounded predicates -- the same depths, the same winner
// (largest depth, ties -> smallest closing endpoint) in O(N log^2 N).
constexpr int BT_SORT_MIN = 1024;

__device__ inline void block_translation(BlockScratch* s, const double* __restrict__ src,
                                         const double* __restrict__ dst, const int* __restrict__ idx, int P,
                                         double scale, const double R[9], double sigma, const double* last_best,
                                         double* __restrict__ xs, double t_out[3], double* __restrict__ sorted = nullptr) {
  __shared__ StabBest warp_best[BLK / 32];
  __shared__ StabBest blk_best;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int N = P + (last_best ? 1 : 0);
  // x_m per axis: raw_translation = dst - (s*R) * src  (registration.cc:446 with :1248's argument)
  for (int m = tid; m < P; m += BLK) {
    const double* p = src + 3 * (size_t)idx[m];
    const double* q = dst + 3 * (size_t)idx[m];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double v1 = dadd(dadd(dmul(dmul(scale, R[r * 3 + 0]), p[0]), dmul(dmul(scale, R[r * 3 + 1]), p[1])),
                             dmul(dmul(scale, R[r * 3 + 2]), p[2]));
      xs[(size_t)r * (P + 1) + m] = dsub(q[r], v1);
    }
  }
  if (last_best && tid < 3) xs[(size_t)tid * (P + 1) + P] = last_best[tid];
  __syncthreads();
  for (int axis = 0; axis < 3; ++axis) {
    const double* x = xs + (size_t)axis * (P + 1);
    StabBest best;
    best.depth = 0;
    best.hi = 0.0;
    best.k = -1;
    if (sorted != nullptr && N > BT_SORT_MIN) {
      int M = 1;
      while (M < N) M <<= 1;
      for (int i = tid; i < M; i += BLK) sorted[i] = (i < N) ? x[i] : INFINITY;
      __syncthreads();
      for (int kk = 2; kk <= M; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < M; i += BLK) {
            const int l = i ^ j;
            if (l > i) {
              const double a = sorted[i], b = sorted[l];
              const bool up = (i & kk) == 0;
              if ((a > b) == up) {
                sorted[i] = b;
                sorted[l] = a;
              }
            }
          }
          __syncthreads();
        }
      // candidate = sorted position k (ascending closing endpoints: the first best is the one with the smallest)
      for (int k = tid; k < N; k += BLK) {
        const double hik = dadd(sorted[k], sigma);
        int lo = 0, hi = k;  // first position whose closing endpoint reaches hik (position k does)
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (dadd(sorted[mid], sigma) >= hik) hi = mid; else lo = mid + 1;
        }
        const int first = lo;
        lo = k;
        hi = N - 1;  // last position whose opening endpoint is not beyond hik (position k is not)
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (dsub(sorted[mid], sigma) <= hik) lo = mid; else hi = mid - 1;
        }
        const int cnt = lo - first + 1;
        if (cnt > best.depth) {  // (k ascends per thread: an equal depth later has a larger endpoint)
          best.depth = cnt;
          best.hi = hik;
          best.k = k;
        }
      }
      // warp-level combine, then the block-level code below
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        StabBest c;
        c.depth = __shfl_xor_sync(0xffffffffu, best.depth, o);
        c.hi = __shfl_xor_sync(0xffffffffu, best.hi, o);
        c.k = __shfl_xor_sync(0xffffffffu, best.k, o);
        if (c.k >= 0 && (best.k < 0 || c.depth > best.depth ||
                         (c.depth == best.depth && (c.hi < best.hi || (c.hi == best.hi && c.k < best.k)))))
          best = c;
      }
    } else
    for (int k = wid; k < N; k += BLK / 32) {
      const double hik = dadd(x[k], sigma);
      int cnt = 0;
      for (int m = lane; m < N; m += 32) {
        const double xm = x[m];
        cnt += (dsub(xm, sigma) <= hik && dadd(xm, sigma) >= hik) ? 1 : 0;
      }
      cnt = warp_sum_int(cnt);
      if (cnt > best.depth || (cnt == best.depth && best.k >= 0 && hik < best.hi)) {
        best.depth = cnt;
        best.hi = hik;
        best.k = k;
      }
    }
    if (lane == 0) warp_best[wid] = best;
    __syncthreads();
    if (tid == 0) {
      StabBest b = warp_best[0];
      for (int w = 1; w < BLK / 32; ++w) {
        const StabBest c = warp_best[w];
        if (c.k >= 0 && (b.k < 0 || c.depth > b.depth || (c.depth == b.depth && (c.hi < b.hi || (c.hi == b.hi && c.k < b.k)))))
          b = c;
      }
      blk_best = b;
    }
    __syncthreads();
    const StabBest b = blk_best;
    if (b.k >= 0) {
      double sum[1] = {0.0};
      for (int m = tid; m < N; m += BLK) {
        const double xm = x[m];
        if (dsub(xm, sigma) <= b.hi && dadd(xm, sigma) >= b.hi) sum[0] += xm;
      }
      block_sum<1>(s, sum);
      t_out[axis] = sum[0] / (double)b.depth;
    }
    __syncthreads();
  }
}

}  // namespace psulvsb
