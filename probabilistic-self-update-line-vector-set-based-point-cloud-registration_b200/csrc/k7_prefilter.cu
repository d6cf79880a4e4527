// k7_prefilter.cu -- the driver-side pre-filter as a device stage: normal-angle histogram -> keep_mask, fused with the
// reduced-set builder (keep_mask -> reduce_map + gathered src / tgt columns).
//
// What it replaces: the two host loops inside the reference driver's timed region,
// examples/teaser_cpp_ply/PSULVSB.cc:87-172 and :174-188.  The reference's version is a sequential scan whose
// peak bin is "whichever bin first exceeds the running maximum while correspondences are pushed in index order".
// That rule is restated here without the scan order:
//     peak = among the bins whose final height equals the maximum height H, the one whose LAST member has the
//            smallest correspondence index (its H-th member arrived first),
// so every phase is a data-parallel pass with integer atomics (order-free) and fixed-order reductions:
//   phase 1  angle of every correspondence (FP64), min / max / sum / count
//   phase 2  sum of squared deviations -> sigma -> bin width 3.49 sigma / cnt^(1/3) -> number of bins
//   phase 3  bin of every correspondence; height[bin] += 1, last[bin] = max(last[bin], i)
//   phase 4  maximum height, peak bin, mean / sigma of the heights -> threshold
//   phase 5  keep = 1 (bin taller than the threshold) else -1 (bin further than 2 from the peak) else unchanged
//   phase 6  exclusive scan of (keep == 1) -> reduce_map, gather of the kept columns
// One 1024-thread CTA per correspondence set (blockIdx.x), so a batch of fragment pairs is filtered in one launch,
// like the engine's control kernels.  Angles go through CUDA's acos, the reference's through libm: individual angles
// may differ in the last bit, which moves a correspondence to another bin only if it sits within ~1e-14 degrees of a
// bin edge (tests compare the masks with the numpy restatement on seeded inputs).
// Positions taken where the reference is undefined: sigma == 0 (all angles equal) -> one bin; an angle on the
// upper edge of the last bin goes into the last bin (the reference indexes one past the end there).
#include <cuda_runtime.h>

#include <cmath>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/psulvsb_io.h"
#include "common.cuh"
#include "solve_dev.cuh"

namespace psulvsb {

struct PrefilterJob {
  const double* src_normals;  // column-major 3 x n
  const double* tgt_normals;
  const double* src;  // optional (phase 6 gather): column-major 3 x n
  const double* tgt;
  int n;
  int run_histogram;  // 0: keep_mask is an input (mask_filter only)
  int* keep_mask;     // [n] in/out
  double* angle;      // [n] scratch
  int* bin;           // [n] scratch
  unsigned int* height;  // [n + 16] scratch
  int* last;             // [n + 16] scratch
  double* src_reduce;    // optional [3 n]
  double* tgt_reduce;
  int* reduce_map;  // optional [n]
  int* out;         // [2]: remain_count, C
};

namespace {

__device__ __forceinline__ void unit3(const double* __restrict__ v, double o[3]) {
  const double z = sqnorm3(v[0], v[1], v[2]);
  if (z > 0.0) {
    const double nrm = sqrt(z);
    o[0] = v[0] / nrm;
    o[1] = v[1] / nrm;
    o[2] = v[2] / nrm;
  } else {  // (a zero vector stays zero, a NaN stays NaN)
    o[0] = v[0];
    o[1] = v[1];
    o[2] = v[2];
  }
}

// block-wide min / max of one double each (fixed order)
__device__ __forceinline__ void block_minmax(BlockScratch* s, double& lo, double& hi) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) {
    s->d[wid][0] = lo;
    s->d[wid][1] = hi;
  }
  __syncthreads();
  if (tid == 0) {
    double a = s->d[0][0], b = s->d[0][1];
    for (int w = 1; w < BLK / 32; ++w) {
      a = fmin(a, s->d[w][0]);
      b = fmax(b, s->d[w][1]);
    }
    s->bd[0] = a;
    s->bd[1] = b;
  }
  __syncthreads();
  lo = s->bd[0];
  hi = s->bd[1];
  __syncthreads();
}

__global__ void __launch_bounds__(BLK) prefilter_kernel(const PrefilterJob* __restrict__ jobs) {
  const PrefilterJob J = jobs[blockIdx.x];
  __shared__ BlockScratch scratch;
  __shared__ int sh_i[4];
  const int tid = threadIdx.x;
  const int n = J.n;
  int remain = 0;
  if (J.run_histogram) {
    // ---- phase 1: angles (PSULVSB.cc:95-110)
    double lo = INFINITY, hi = -INFINITY, acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += BLK) {
      double a[3], b[3];
      unit3(J.src_normals + 3 * (size_t)i, a);
      unit3(J.tgt_normals + 3 * (size_t)i, b);
      double c = dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2]));
      c = (c < 1.0) ? c : 1.0;    // std::min(1.0, c): a NaN cosine becomes 1 (angle 0), it is not skipped
      c = (-1.0 < c) ? c : -1.0;  // std::max(-1.0, .)
      const double deg = acos(c) * 180.0 / 3.14159265358979323846;
      if (isnan(deg)) {
        J.angle[i] = -1.0;
      } else {
        J.angle[i] = deg;
        lo = fmin(lo, deg);
        hi = fmax(hi, deg);
        acc[0] += deg;
        acc[1] += 1.0;
      }
    }
    block_minmax(&scratch, lo, hi);
    block_sum<4>(&scratch, acc);
    const double cnt = acc[1];
    if (cnt > 0.0) {
      // the reference starts its running extrema at o_max = 0, o_min = INT_MAX (PSULVSB.cc:93)
      const double o_max = fmax(hi, 0.0), o_min = fmin(lo, 2147483647.0);
      const double mean = acc[0] / cnt;
      // ---- phase 2: spread -> bin width (PSULVSB.cc:112-121)
      double sq[4] = {0.0, 0.0, 0.0, 0.0};
      for (int i = tid; i < n; i += BLK) {
        const double d = J.angle[i];
        if (d != -1.0) sq[0] += (d - mean) * (d - mean);
      }
      block_sum<4>(&scratch, sq);
      const double sd = sqrt(sq[0] / cnt);
      const double width = 3.49 * sd / pow(cnt, 1.0 / 3.0);
      const bool binned = width > 0.0 && isfinite(width);
      int nbins = 1;
      if (binned) {
        const double q = ceil((o_max - o_min) / width);
        nbins = q < 1.0 ? 1 : (q > (double)(n + 15) ? n + 15 : (int)q);
      }
      for (int b = tid; b < nbins; b += BLK) {
        J.height[b] = 0u;
        J.last[b] = -1;
      }
      __syncthreads();
      // ---- phase 3: heights and last members (PSULVSB.cc:124-135, order-free)
      for (int i = tid; i < n; i += BLK) {
        const double d = J.angle[i];
        int b = -1;
        if (d != -1.0) {
          b = binned ? (int)((d - o_min) / width) : 0;
          b = b >= nbins ? nbins - 1 : (b < 0 ? 0 : b);
          atomicAdd(J.height + b, 1u);
          atomicMax(J.last + b, i);
        }
        J.bin[i] = b;
      }
      __syncthreads();
      // ---- phase 4: peak and threshold (PSULVSB.cc:130-133, :137-150)
      unsigned int hmax = 0u;
      for (int b = tid; b < nbins; b += BLK) hmax = max(hmax, J.height[b]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
      if ((tid & 31) == 0) scratch.i[tid >> 5][0] = (int)hmax;
      __syncthreads();
      if (tid == 0) {
        int m = 0;
        for (int w = 0; w < BLK / 32; ++w) m = max(m, scratch.i[w][0]);
        sh_i[0] = m;
        sh_i[1] = 0x7fffffff;  // smallest last-member index among the tallest bins
      }
      __syncthreads();
      hmax = (unsigned int)sh_i[0];
      for (int b = tid; b < nbins; b += BLK)
        if (J.height[b] == hmax) atomicMin(&sh_i[1], J.last[b]);
      __syncthreads();
      const int peak = J.bin[sh_i[1]];
      const double hmean = cnt / (double)nbins;  // the heights add up to cnt exactly
      double hv[4] = {0.0, 0.0, 0.0, 0.0};
      for (int b = tid; b < nbins; b += BLK) {
        const double d = (double)(int)J.height[b] - hmean;
        hv[0] += d * d;
      }
      block_sum<4>(&scratch, hv);
      const double threshold = hmean + sqrt(hv[0] / (double)nbins);
      // ---- phase 5: the mask (PSULVSB.cc:152-166)
      for (int i = tid; i < n; i += BLK) {
        const int b = J.bin[i];
        if (b < 0) continue;
        if ((double)J.height[b] > threshold) {
          J.keep_mask[i] = 1;
          ++remain;
        } else if (abs(b - peak) > 2) {
          J.keep_mask[i] = -1;
        }
      }
    }
    int dummy = 0;
    block_sum_int2(&scratch, remain, dummy);
  }
  // ---- phase 6: reduced set (PSULVSB.cc:174-188)
  int C = 0;
  if (J.reduce_map) {
    if (tid == 0) sh_i[2] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += BLK) {
      const int i = i0 + tid;
      const int f = (i < n && J.keep_mask[i] == 1) ? 1 : 0;
      int ea, eb, ta, tb;
      block_scan2(&scratch, f, 0, ea, eb, ta, tb);
      const int base = sh_i[2];
      if (i < n) J.reduce_map[i] = f ? base + ea : -1;
      if (f && J.src_reduce) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          J.src_reduce[3 * (size_t)(base + ea) + r] = J.src[3 * (size_t)i + r];
          J.tgt_reduce[3 * (size_t)(base + ea) + r] = J.tgt[3 * (size_t)i + r];
        }
      }
      __syncthreads();
      if (tid == 0) sh_i[2] = base + ta;
      __syncthreads();
    }
    C = sh_i[2];
  }
  if (tid == 0) {
    J.out[0] = remain;
    J.out[1] = C;
  }
}

struct DevMem {
  void* p = nullptr;
  ~DevMem() {
    if (p) cudaFree(p);
  }
  int alloc(size_t bytes) {
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(PSULVSB_ERR_CUDA, "pre-filter: cudaMalloc failed");
    }
    return PSULVSB_OK;
  }
};

inline size_t up(size_t x) { return (x + 255) & ~size_t(255); }

// B correspondence sets on host buffers: one arena, one launch (a CTA per set), results copied back.  Any of the
// optional groups may be absent (the same for every set).
int run_host_batch(int B, const double* const* src_normals, const double* const* tgt_normals, const double* const* src,
                   const double* const* tgt, const int* n_arr, bool histogram, int* const* keep_mask,
                   double* const* src_reduce, double* const* tgt_reduce, int* const* reduce_map, int* remain, int* C) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(PSULVSB_ERR_NO_DEVICE, "pre-filter: no CUDA device available (there is no CPU fallback)");
  }
  for (int b = 0; b < B; ++b) {
    if (remain) remain[b] = 0;
    if (C) C[b] = 0;
  }
  const bool gather = src_reduce != nullptr;
  // arena per set: [normals 6n] [points 6n] [angle n] [reduced 6n] | [keep n] [bin n] [height n+16] [last n+16] [map n] [out 2]
  struct Off {
    size_t sn, tn, s, t, ang, sr, tr, keep, bin, h, l, map, out;
  };
  std::vector<Off> offs((size_t)B);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += up(bytes);
    return o;
  };
  bool any = false;
  for (int b = 0; b < B; ++b) {
    const size_t nn = (size_t)n_arr[b];
    any = any || nn > 0;
    Off& o = offs[(size_t)b];
    o.sn = take(histogram ? 24 * nn : 0);
    o.tn = take(histogram ? 24 * nn : 0);
    o.s = take(gather ? 24 * nn : 0);
    o.t = take(gather ? 24 * nn : 0);
    o.ang = take(8 * nn);
    o.sr = take(gather ? 24 * nn : 0);
    o.tr = take(gather ? 24 * nn : 0);
    o.keep = take(4 * nn);
    o.bin = take(4 * nn);
    o.h = take(4 * (nn + 16));
    o.l = take(4 * (nn + 16));
    o.map = take(4 * nn);
    o.out = take(8);
  }
  if (!any) return PSULVSB_OK;
  const size_t o_jobs = take(sizeof(PrefilterJob) * (size_t)B);
  // grow-only arena shared by the calls of this process (cudaMalloc + cudaFree of a few hundred MB per call cost more
  // than the launch they serve); the calls are synchronous by contract, the mutex serialises concurrent callers
  static std::mutex arena_mutex;
  static DevMem arena;
  static size_t arena_cap = 0;
  std::lock_guard<std::mutex> lock(arena_mutex);
  if (off > arena_cap) {
    if (arena.p) cudaFree(arena.p);
    arena.p = nullptr;
    arena_cap = 0;
    if (int rc = arena.alloc(off + off / 4)) return rc;
    arena_cap = off + off / 4;
  }
  char* base = static_cast<char*>(arena.p);
  cudaStream_t st = nullptr;  // (the legacy stream)
  std::vector<PrefilterJob> jobs((size_t)B);
  for (int b = 0; b < B; ++b) {
    const size_t nn = (size_t)n_arr[b];
    const Off& o = offs[(size_t)b];
    if (nn > 0) {
      if (histogram) {
        PSU_CUDA(cudaMemcpyAsync(base + o.sn, src_normals[b], 24 * nn, cudaMemcpyHostToDevice, st));
        PSU_CUDA(cudaMemcpyAsync(base + o.tn, tgt_normals[b], 24 * nn, cudaMemcpyHostToDevice, st));
      }
      if (gather) {
        PSU_CUDA(cudaMemcpyAsync(base + o.s, src[b], 24 * nn, cudaMemcpyHostToDevice, st));
        PSU_CUDA(cudaMemcpyAsync(base + o.t, tgt[b], 24 * nn, cudaMemcpyHostToDevice, st));
      }
      PSU_CUDA(cudaMemcpyAsync(base + o.keep, keep_mask[b], 4 * nn, cudaMemcpyHostToDevice, st));
    }
    PrefilterJob& J = jobs[(size_t)b];
    J.src_normals = reinterpret_cast<const double*>(base + o.sn);
    J.tgt_normals = reinterpret_cast<const double*>(base + o.tn);
    J.src = reinterpret_cast<const double*>(base + o.s);
    J.tgt = reinterpret_cast<const double*>(base + o.t);
    J.n = n_arr[b];
    J.run_histogram = histogram ? 1 : 0;
    J.keep_mask = reinterpret_cast<int*>(base + o.keep);
    J.angle = reinterpret_cast<double*>(base + o.ang);
    J.bin = reinterpret_cast<int*>(base + o.bin);
    J.height = reinterpret_cast<unsigned int*>(base + o.h);
    J.last = reinterpret_cast<int*>(base + o.l);
    J.src_reduce = gather ? reinterpret_cast<double*>(base + o.sr) : nullptr;
    J.tgt_reduce = gather ? reinterpret_cast<double*>(base + o.tr) : nullptr;
    J.reduce_map = reduce_map ? reinterpret_cast<int*>(base + o.map) : nullptr;
    J.out = reinterpret_cast<int*>(base + o.out);
  }
  PSU_CUDA(cudaMemcpyAsync(base + o_jobs, jobs.data(), sizeof(PrefilterJob) * (size_t)B, cudaMemcpyHostToDevice, st));
  prefilter_kernel<<<B, BLK, 0, st>>>(reinterpret_cast<const PrefilterJob*>(base + o_jobs));
  PSU_CHECK_LAUNCH("prefilter_kernel");
  std::vector<int> out((size_t)2 * B, 0);
  for (int b = 0; b < B; ++b) {
    const size_t nn = (size_t)n_arr[b];
    const Off& o = offs[(size_t)b];
    PSU_CUDA(cudaMemcpyAsync(&out[(size_t)2 * b], base + o.out, 8, cudaMemcpyDeviceToHost, st));
    if (nn == 0) continue;
    if (histogram) PSU_CUDA(cudaMemcpyAsync(keep_mask[b], base + o.keep, 4 * nn, cudaMemcpyDeviceToHost, st));
    if (reduce_map) PSU_CUDA(cudaMemcpyAsync(reduce_map[b], base + o.map, 4 * nn, cudaMemcpyDeviceToHost, st));
  }
  PSU_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b < B; ++b) {
    const Off& o = offs[(size_t)b];
    const int kept = n_arr[b] > 0 ? out[(size_t)2 * b + 1] : 0;
    if (gather && kept > 0) {
      PSU_CUDA(cudaMemcpyAsync(src_reduce[b], base + o.sr, 24 * (size_t)kept, cudaMemcpyDeviceToHost, st));
      PSU_CUDA(cudaMemcpyAsync(tgt_reduce[b], base + o.tr, 24 * (size_t)kept, cudaMemcpyDeviceToHost, st));
    }
    if (remain) remain[b] = n_arr[b] > 0 ? out[(size_t)2 * b] : 0;
    if (C) C[b] = kept;
  }
  PSU_CUDA(cudaStreamSynchronize(st));
  return PSULVSB_OK;
}

// one correspondence set
int run_host(const double* src_normals, const double* tgt_normals, const double* src, const double* tgt, int n,
             bool histogram, int* keep_mask, double* src_reduce, double* tgt_reduce, int* reduce_map, int* remain,
             int* C) {
  return run_host_batch(1, &src_normals, &tgt_normals, &src, &tgt, &n, histogram, &keep_mask,
                        src_reduce ? &src_reduce : nullptr, tgt_reduce ? &tgt_reduce : nullptr,
                        reduce_map ? &reduce_map : nullptr, remain, C);
}

}  // namespace

// device-buffer entry for callers that keep correspondences resident (the bench's pre-filter + solve line)
int launch_prefilter(cudaStream_t st, const PrefilterJob* d_jobs, int n_jobs) {
  if (n_jobs <= 0) return PSULVSB_OK;
  prefilter_kernel<<<n_jobs, BLK, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("prefilter_kernel");
  return PSULVSB_OK;
}

}  // namespace psulvsb

using psulvsb::fail;

extern "C" {

int psulvsb_histogram_outlier_removal(const double* src_normals, const double* tgt_normals, int n, int* keep_mask,
                                      int* remain_count) {
  if (!src_normals || !tgt_normals || !keep_mask || n < 0)
    return fail(PSULVSB_ERR_INVALID, "psulvsb_histogram_outlier_removal: bad argument");
  return psulvsb::run_host(src_normals, tgt_normals, nullptr, nullptr, n, true, keep_mask, nullptr, nullptr, nullptr,
                           remain_count, nullptr);
}

int psulvsb_mask_filter(const double* src, const double* tgt, const int* keep_mask, int n, double* src_reduce,
                        double* tgt_reduce, int* reduce_map, int* C) {
  if (!src || !tgt || !keep_mask || !src_reduce || !tgt_reduce || !reduce_map || !C || n < 0)
    return fail(PSULVSB_ERR_INVALID, "psulvsb_mask_filter: bad argument");
  return psulvsb::run_host(nullptr, nullptr, src, tgt, n, false, const_cast<int*>(keep_mask), src_reduce, tgt_reduce,
                           reduce_map, nullptr, C);
}

int psulvsb_prefilter_reduce(const double* src_normals, const double* tgt_normals, const double* src, const double* tgt,
                             int n, int* keep_mask, double* src_reduce, double* tgt_reduce, int* reduce_map, int* C,
                             int* remain_count) {
  if (!src_normals || !tgt_normals || !src || !tgt || !keep_mask || !src_reduce || !tgt_reduce || !reduce_map || !C ||
      n < 0)
    return fail(PSULVSB_ERR_INVALID, "psulvsb_prefilter_reduce: bad argument");
  return psulvsb::run_host(src_normals, tgt_normals, src, tgt, n, true, keep_mask, src_reduce, tgt_reduce, reduce_map,
                           remain_count, C);
}

int psulvsb_prefilter_reduce_batch(int B, const double* const* src_normals, const double* const* tgt_normals,
                                   const double* const* src, const double* const* tgt, const int* n,
                                   int* const* keep_mask, double* const* src_reduce, double* const* tgt_reduce,
                                   int* const* reduce_map, int* C, int* remain_count) {
  if (B <= 0 || !src_normals || !tgt_normals || !src || !tgt || !n || !keep_mask || !src_reduce || !tgt_reduce ||
      !reduce_map || !C)
    return fail(PSULVSB_ERR_INVALID, "psulvsb_prefilter_reduce_batch: bad argument");
  for (int b = 0; b < B; ++b)
    if (n[b] < 0 || (n[b] > 0 && (!src_normals[b] || !tgt_normals[b] || !src[b] || !tgt[b] || !keep_mask[b] ||
                                  !src_reduce[b] || !tgt_reduce[b] || !reduce_map[b])))
      return fail(PSULVSB_ERR_INVALID, "psulvsb_prefilter_reduce_batch: set " + std::to_string(b) + " has a null array");
  return psulvsb::run_host_batch(B, src_normals, tgt_normals, src, tgt, n, true, keep_mask, src_reduce, tgt_reduce,
                                 reduce_map, remain_count, C);
}

}  // extern "C"
