// k2_sampler.cu -- stage 2: replayable sampling without replacement from a Philox stream.
//
// The reference draws `count` distinct indices with  do { r = rand() % n; } while (used[r]);
// (registration.cc:852-861, :916-932) -- inherently sequential, with a data-dependent number of
// consumed draws.  The same index sequence is produced in parallel here: draw k (a pure function
// of (seed, domain, event, k)) is accepted iff it is the FIRST occurrence of its value, i.e. iff
// first[v_k] == k where first[] is built with atomicMin; the r-th accepted draw is output r.
// That is exactly the rejection rule, so the output equals the sequential algorithm's for the same
// stream, and the stream position after the call (draws consumed) is reported for replay.
//
// Kernels take device-resident SampleJob arrays (one job per registration in the batch engine,
// whose control kernels rewrite n / count / event between ticks).
#include <cuda_runtime.h>

#include "common.cuh"
#include "engine.cuh"

namespace psulvsb {

namespace {

__global__ void __launch_bounds__(256) sample_mark_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.y];
  if (!job.active || job.identity) return;
  const unsigned long long n = job.n, max_draws = job.max_draws;
  uint32_t* __restrict__ first = job.first;
  // one Philox block (4 draws) per thread iteration
  const unsigned long long nblocks = (max_draws + 3) >> 2;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nblocks;
       q += (unsigned long long)gridDim.x * blockDim.x) {
    const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const unsigned long long k = (q << 2) + l;
      if (k < max_draws) {
        const unsigned long long v = (unsigned long long)(o.w[l] >> 1) % n;
        atomicMin(&first[v], (uint32_t)k);
      }
    }
  }
}

// one CTA per job: ordered emission of the accepted draws; restores first[] to 0xFFFFFFFF on the way
__global__ void __launch_bounds__(1024) sample_emit_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.x];
  if (!job.active) return;
  __shared__ unsigned int warp_tot[32];
  __shared__ unsigned long long base_s;
  __shared__ int flag_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned long long n = job.n, count = job.count, max_draws = job.max_draws;
  uint32_t* __restrict__ first = job.first;
  uint32_t* __restrict__ out = job.out;
  if (tid == 0) {
    base_s = 0ull;
    flag_total = 0;
  }
  if (job.post == 1)
    for (int i = tid; i < job.n_points; i += 1024) job.flags[i] = 0;
  __syncthreads();
  if (job.identity) {
    for (unsigned long long r = tid; r < count; r += 1024) out[r] = (uint32_t)r;
    if (tid == 0 && job.status) job.status[0] = 1ull;  // nothing drawn; non-zero = success
  } else {
    const unsigned long long nblocks = (max_draws + 3) >> 2;
    unsigned long long consumed = 0ull;  // meaningful in the thread that emits the last sample
    bool have_last = false;
    for (unsigned long long q0 = 0; q0 < nblocks; q0 += 1024) {
      const unsigned long long base = base_s;  // no early exit: every marked entry of first[] gets restored
      const unsigned long long q = q0 + tid;
      uint32_t v[4];
      bool acc[4] = {false, false, false, false};
      unsigned int c = 0;
      if (q < nblocks) {
        const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const unsigned long long k = (q << 2) + l;
          v[l] = (uint32_t)((unsigned long long)(o.w[l] >> 1) % n);
          if (k < max_draws) {
            acc[l] = first[v[l]] == (uint32_t)k;
            c += acc[l] ? 1u : 0u;
          }
        }
      }
      // block exclusive scan of c
      unsigned int incl = c;
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += t;
      }
      if (lane == 31) warp_tot[wid] = incl;
      __syncthreads();
      if (wid == 0) {
        unsigned int t = warp_tot[lane];
        unsigned int ti = t;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
          unsigned int u = __shfl_up_sync(0xffffffffu, ti, o2);
          if (lane >= o2) ti += u;
        }
        warp_tot[lane] = ti - t;  // exclusive
        if (lane == 31) base_s = base + ti;
      }
      __syncthreads();
      unsigned long long rank = base + warp_tot[wid] + (incl - c);
      if (q < nblocks) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const unsigned long long k = (q << 2) + l;
          if (acc[l]) {
            if (rank < count) {
              out[rank] = v[l];
              if (rank == count - 1) {
                consumed = k + 1;
                have_last = true;
              }
            }
            ++rank;
            first[v[l]] = 0xFFFFFFFFu;  // restore (every reader of this round passed the barrier above)
          }
        }
      }
      __syncthreads();
    }
    if (have_last && job.status) job.status[0] = consumed;
    __syncthreads();
    if (tid == 0 && base_s < count && job.status) job.status[0] = 0ull;  // not enough draws
  }
  __syncthreads();  // out[] complete (block-scope visibility of the global writes above)
  // ---- fused post-processing for the batch engine
  if (job.post == 1) {
    // src_sampled/dst_sampled = unique endpoints of the sampled line vectors (registration.cc:870-894);
    // only the SET matters downstream (inlier counts), so it is kept as per-point flags
    for (unsigned long long r = tid; r < count; r += 1024) {
      const uint2 e = job.edges[out[r]];
      job.flags[e.x] = 1;
      job.flags[e.y] = 1;
    }
    __syncthreads();
    int c = 0;
    for (int i = tid; i < job.n_points; i += 1024) c += job.flags[i] ? 1 : 0;
    c = warp_sum_int(c);
    if (lane == 0 && c) atomicAdd(&flag_total, c);
    __syncthreads();
    if (tid == 0 && job.flag_count) *job.flag_count = flag_total;
  } else if (job.post == 2) {
    // basic line vectors (registration.cc:922-925) as endpoint pairs
    for (unsigned long long r = tid; r < count; r += 1024) job.gathered[r] = job.edges[job.via[out[r]]];
  }
}

__global__ void philox_fill_kernel(uint64_t seed, uint32_t domain, uint32_t event, unsigned long long first_k,
                                   unsigned long long count, uint32_t* __restrict__ out) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = philox_rand31(seed, domain, event, first_k + i);
}

}  // namespace

unsigned long long sample_default_max_draws(unsigned long long n, unsigned long long count) {
  return sample_max_draws_formula(n, count);
}

int launch_sample(cudaStream_t st, const SampleJob* d_jobs, int n_jobs, unsigned long long max_draws_bound) {
  if (n_jobs <= 0) return PSULVSB_OK;
  const unsigned long long nblocks = (max_draws_bound + 3) >> 2;
  unsigned long long gx = (nblocks + 255) / 256;
  const unsigned long long cap = (unsigned long long)(148 * 8 / (n_jobs < 8 ? n_jobs : 8)) + 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)n_jobs);
  sample_mark_kernel<<<grid, 256, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("sample_mark_kernel");
  sample_emit_kernel<<<n_jobs, 1024, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("sample_emit_kernel");
  return PSULVSB_OK;
}

int launch_philox_fill(cudaStream_t st, uint64_t seed, uint32_t domain, uint32_t event, unsigned long long first_k,
                       unsigned long long count, uint32_t* out) {
  if (count == 0) return PSULVSB_OK;
  philox_fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(seed, domain, event, first_k, count, out);
  PSU_CHECK_LAUNCH("philox_fill_kernel");
  return PSULVSB_OK;
}

}  // namespace psulvsb
