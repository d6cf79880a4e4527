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
// Three fully parallel passes over the draw window, every one a grid of (chunks x jobs) CTAs:
//   mark  : first[v_k] = min(first[v_k], k)
//   count : accepted draws per 1024-draw chunk; the last CTA of a job to finish turns the chunk
//           counts into exclusive prefixes (threadfence reduction, no CTA ever waits on another)
//   emit  : rank = prefix[chunk] + position in chunk -> out[rank]; restores first[] on the way;
//           optional fused post-processing for the batch engine (endpoint flags / edge gather)
// Kernels take device-resident SampleJob arrays (one job per registration in the batch engine,
// whose control kernels rewrite n / count / event between ticks).
#include <cuda_runtime.h>

#include "common.cuh"
#include "engine.cuh"

namespace psulvsb {

namespace {

constexpr int SMP_THREADS = 256;
constexpr int SMP_CHUNK = SMP_THREADS * 4;  // draws per chunk (one Philox block per thread)

__device__ __forceinline__ uint32_t draw_value(uint32_t word, uint32_t n) { return (word >> 1) % n; }

__global__ void __launch_bounds__(SMP_THREADS) sample_mark_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.y];
  if (!job.active || job.identity) return;
  const uint32_t n = (uint32_t)job.n;
  const unsigned long long max_draws = job.max_draws;
  uint32_t* __restrict__ first = job.first;
  const unsigned long long nblocks = (max_draws + 3) >> 2;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < nblocks;
       q += (unsigned long long)gridDim.x * blockDim.x) {
    const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const unsigned long long k = (q << 2) + l;
      if (k < max_draws) atomicMin(&first[draw_value(o.w[l], n)], (uint32_t)k);
    }
  }
}

__device__ __forceinline__ unsigned int block_sum_u32(unsigned int v, unsigned int* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  unsigned int t = 0;
#pragma unroll
  for (int w = 0; w < SMP_THREADS / 32; ++w) t += sh[w];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(SMP_THREADS) sample_count_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  __shared__ unsigned int sh[SMP_THREADS / 32];
  __shared__ unsigned int ticket_s;
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x;
  if (job.post == 1)  // endpoint flags are rebuilt by the emit pass
    for (int i = blockIdx.x * SMP_THREADS + tid; i < job.n_points; i += gridDim.x * SMP_THREADS) job.flags[i] = 0;
  if (job.identity) return;
  const uint32_t n = (uint32_t)job.n;
  const unsigned long long max_draws = job.max_draws;
  const unsigned long long nchunks = (max_draws + SMP_CHUNK - 1) / SMP_CHUNK;
  const uint32_t* __restrict__ first = job.first;
  for (unsigned long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const unsigned long long q = ch * SMP_THREADS + tid;
    unsigned int c = 0;
    if ((q << 2) < max_draws) {
      const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const unsigned long long k = (q << 2) + l;
        if (k < max_draws && first[draw_value(o.w[l], n)] == (uint32_t)k) ++c;
      }
    }
    const unsigned int tot = block_sum_u32(c, sh);
    if (tid == 0) job.chunk_prefix[ch] = (unsigned long long)tot;
  }
  // last CTA of this job: counts -> exclusive prefixes
  __threadfence();
  if (tid == 0) ticket_s = atomicAdd(job.ticket, 1u);
  __syncthreads();
  if (ticket_s != gridDim.x - 1) return;
  __threadfence();
  if (tid == 0) {
    carry_s = 0ull;
    *job.ticket = 0u;  // ready for the next use
  }
  __syncthreads();
  volatile unsigned long long* pref = job.chunk_prefix;
  for (unsigned long long c0 = 0; c0 < nchunks; c0 += SMP_THREADS) {
    const unsigned long long ch = c0 + tid;
    const unsigned int v = (ch < nchunks) ? (unsigned int)pref[ch] : 0u;
    // block exclusive scan of v
    const int lane = tid & 31, wid = tid >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < SMP_THREADS / 32; ++w)
      if (w < wid) wbase += sh[w];
    const unsigned long long carry = carry_s;
    if (ch < nchunks) pref[ch] = carry + wbase + (incl - v);
    __syncthreads();
    if (tid == SMP_THREADS - 1) carry_s = carry + wbase + incl;
    __syncthreads();
  }
  if (tid == 0 && carry_s < job.count && job.status) job.status[0] = 0ull;  // draw budget too small
}

__global__ void __launch_bounds__(SMP_THREADS) sample_emit_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  __shared__ unsigned int sh[SMP_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned long long count = job.count;
  uint32_t* __restrict__ out = job.out;
  if (job.identity) {  // registration.cc:839-847: the sample is the whole set, in order
    for (unsigned long long r = (unsigned long long)blockIdx.x * SMP_THREADS + tid; r < count;
         r += (unsigned long long)gridDim.x * SMP_THREADS) {
      out[r] = (uint32_t)r;
      if (job.post == 1) {
        const uint2 e = job.edges[r];
        job.flags[e.x] = 1;
        job.flags[e.y] = 1;
      } else if (job.post == 2) {
        job.gathered[r] = job.edges[job.via[r]];
      }
    }
    if (blockIdx.x == 0 && tid == 0 && job.status) job.status[0] = 1ull;  // nothing drawn; non-zero = success
    return;
  }
  const uint32_t n = (uint32_t)job.n;
  const unsigned long long max_draws = job.max_draws;
  const unsigned long long nchunks = (max_draws + SMP_CHUNK - 1) / SMP_CHUNK;
  uint32_t* __restrict__ first = job.first;
  for (unsigned long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const unsigned long long q = ch * SMP_THREADS + tid;
    uint32_t v[4];
    bool acc[4] = {false, false, false, false};
    unsigned int c = 0;
    if ((q << 2) < max_draws) {
      const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const unsigned long long k = (q << 2) + l;
        v[l] = draw_value(o.w[l], n);
        if (k < max_draws) {
          acc[l] = first[v[l]] == (uint32_t)k;
          c += acc[l] ? 1u : 0u;
        }
      }
    }
    unsigned int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < SMP_THREADS / 32; ++w)
      if (w < wid) wbase += sh[w];
    unsigned long long rank = job.chunk_prefix[ch] + wbase + (incl - c);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      if (acc[l]) {
        if (rank < count) {
          out[rank] = v[l];
          if (job.post == 1) {
            // src_sampled/dst_sampled = unique endpoints of the sampled line vectors
            // (registration.cc:870-894); only the SET matters downstream, kept as per-point flags
            const uint2 e = job.edges[v[l]];
            job.flags[e.x] = 1;
            job.flags[e.y] = 1;
          } else if (job.post == 2) {
            // basic line vectors (registration.cc:922-925) as endpoint pairs
            job.gathered[rank] = job.edges[job.via[v[l]]];
          }
          if (rank == count - 1 && job.status) job.status[0] = (q << 2) + l + 1;  // draws consumed
        }
        ++rank;
        // restore the table; a rejected draw that still reads this entry sees a value != its own k either way
        first[v[l]] = 0xFFFFFFFFu;
      }
    }
    __syncthreads();
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

unsigned long long sample_chunk_slots(unsigned long long max_draws) { return (max_draws + SMP_CHUNK - 1) / SMP_CHUNK + 1; }

// jobs_per_group > 0: run the three passes group by group so that the random-access working set of a
// group (first-occurrence tables + the edge lists the emit pass gathers from) stays L2 resident.
int launch_sample(cudaStream_t st, const SampleJob* d_jobs, int n_jobs, unsigned long long max_draws_bound,
                  int jobs_per_group) {
  if (n_jobs <= 0) return PSULVSB_OK;
  if (jobs_per_group <= 0 || jobs_per_group > n_jobs) jobs_per_group = n_jobs;
  const unsigned long long nchunks = (max_draws_bound + SMP_CHUNK - 1) / SMP_CHUNK;
  for (int off = 0; off < n_jobs; off += jobs_per_group) {
    const int g = (n_jobs - off < jobs_per_group) ? n_jobs - off : jobs_per_group;
    unsigned long long gx = nchunks;
    const unsigned long long cap = (unsigned long long)(148 * 16) / (unsigned long long)(g < 64 ? g : 64) + 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)g);
    sample_mark_kernel<<<grid, SMP_THREADS, 0, st>>>(d_jobs + off);
    PSU_CHECK_LAUNCH("sample_mark_kernel");
    sample_count_kernel<<<grid, SMP_THREADS, 0, st>>>(d_jobs + off);
    PSU_CHECK_LAUNCH("sample_count_kernel");
    sample_emit_kernel<<<grid, SMP_THREADS, 0, st>>>(d_jobs + off);
    PSU_CHECK_LAUNCH("sample_emit_kernel");
  }
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
