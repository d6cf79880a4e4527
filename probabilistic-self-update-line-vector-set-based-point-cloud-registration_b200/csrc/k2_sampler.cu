// k2_sampler.cu -- stage 2: replayable sampling without replacement from a Philox stream.
//
// The reference draws `count` distinct indices with  do { r = rand() % n; } while (used[r]);
// (registration.cc:852-861, :916-932) -- inherently sequential, with a data-dependent number of
// consumed draws.  The same index sequence is produced in parallel here: draw k (a pure function
// of (seed, domain, event, k)) is accepted iff it is the FIRST occurrence of its value, i.e. iff
// first[v_k] == k where first[v] = min { k : v_k = v }; the r-th accepted draw is output r.
// That is exactly the rejection rule, so the output equals the sequential algorithm's for the same
// stream, and the stream position after the call (draws consumed) is reported for replay.
//
// Two passes, no global table:
//   bucket : the value range [0, n) is cut into buckets of SMP_BW values; CTA (b, job) walks the whole draw
//            window (Philox is cheap: the redundancy buys the removal of every random global access), keeps
//            first[v - lo] = min k for the values of ITS bucket in SHARED memory, and sets bit k of the
//            job's accept bitmask for every draw that is the first occurrence of its value.  The last
//            CTA of a job to finish turns the per-1024-draw popcounts into exclusive prefixes
//            (threadfence reduction, no CTA ever waits on another).
//   emit   : rank = prefix[chunk] + position in chunk -> out[rank]; clears the accept words on the way;
//            optional fused post-processing for the batch engine (endpoint flags / edge gather).
// Kernels take device-resident SampleJob arrays (one job per registration in the batch engine,
// whose control kernels rewrite n / count / event between ticks).
#include <cuda_runtime.h>

#include "common.cuh"
#include "engine.cuh"

namespace psulvsb {

namespace {

constexpr int SMP_THREADS = 256;            // emit pass
constexpr int SMP_CHUNK = SMP_THREADS * 4;  // draws per chunk (one Philox block per thread)
constexpr int SMB_THREADS = 1024;           // bucket pass (one CTA per SM: the table fills shared memory)
constexpr int SMP_BW = 49152;               // values per bucket (192 KB of shared memory)

// (word >> 1) % n with the division replaced by a multiply-high: magic = ceil(2^64 / n) gives the exact
// quotient for every 31-bit dividend (the error term v e / 2^64 < 2^-33 cannot reach the next integer,
// which is at least 1/n > 2^-31 away)
struct FastMod {
  unsigned long long magic;
  uint32_t n;
};
__device__ __forceinline__ FastMod make_fastmod(uint32_t n) {
  FastMod f;
  f.n = n;
  f.magic = n > 1u ? (~0ull / n) + 1ull : 0ull;  // ceil(2^64 / n) (n = 1: everything maps to 0)
  return f;
}
__device__ __forceinline__ uint32_t draw_value(uint32_t word, const FastMod& f) {
  const uint32_t v = word >> 1;
  if (f.n <= 1u) return 0u;
  const uint32_t q = (uint32_t)__umul64hi((unsigned long long)v, f.magic);
  return v - q * f.n;
}

// optional cache of the draw values (SampleJob::draws, one u32 per draw) so that the bucket CTAs stream
// them instead of re-running Philox; windows longer than draws_cap fall back to recomputation
__global__ void __launch_bounds__(256) sample_draws_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.y];
  if (!job.active || job.identity || !job.draws || job.max_draws > job.draws_cap) return;
  const FastMod fm = make_fastmod((uint32_t)job.n);
  const unsigned long long nblocks = (job.max_draws + 3) >> 2;
  uint4* __restrict__ out = reinterpret_cast<uint4*>(job.draws);
  for (unsigned long long q = (unsigned long long)blockIdx.x * 256 + threadIdx.x; q < nblocks;
       q += (unsigned long long)gridDim.x * 256) {
    const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
    out[q] = make_uint4(draw_value(o.w[0], fm), draw_value(o.w[1], fm), draw_value(o.w[2], fm), draw_value(o.w[3], fm));
  }
}

__global__ void __launch_bounds__(SMB_THREADS, 1) sample_bucket_kernel(const SampleJob* __restrict__ jobs) {
  const SampleJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw);  // [SMP_BW]
  __shared__ unsigned int ticket_s, sh[SMB_THREADS / 32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x;
  if (job.post == 1 && blockIdx.x == 0)  // endpoint flags are rebuilt by the emit pass
    for (int i = tid; i < job.n_points; i += SMB_THREADS) job.flags[i] = 0;
  if (job.identity) return;
  const uint32_t n = (uint32_t)job.n;
  const FastMod fm = make_fastmod(n);
  const unsigned int n_buckets = (n + SMP_BW - 1) / SMP_BW;
  if (blockIdx.x >= n_buckets) return;
  const unsigned int n_workers = min(gridDim.x, n_buckets);  // CTAs of this job that take buckets (and a ticket)
  const unsigned long long max_draws = job.max_draws;
  const unsigned long long nblocks = (max_draws + 3) >> 2;
  const bool cached = job.draws != nullptr && max_draws <= job.draws_cap;
  const uint4* __restrict__ dv = reinterpret_cast<const uint4*>(job.draws);
  uint32_t* __restrict__ accept = job.first;  // bit k set <=> draw k is accepted; all zero on entry and on exit
  auto draws_of = [&](uint32_t q) -> uint4 {
    if (cached) return __ldcg(dv + q);
    const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
    return make_uint4(draw_value(o.w[0], fm), draw_value(o.w[1], fm), draw_value(o.w[2], fm), draw_value(o.w[3], fm));
  };
  for (unsigned int b = blockIdx.x; b < n_buckets; b += gridDim.x) {
    const uint32_t lo = b * (uint32_t)SMP_BW;
    const uint32_t width = min((uint32_t)SMP_BW, n - lo);
    __syncthreads();  // the previous bucket's readers are done with the table
    for (uint32_t i = tid; i < width; i += SMB_THREADS) table[i] = 0xFFFFFFFFu;
    __syncthreads();
    // walk 1: first occurrence of every value of this bucket (four 16-byte loads in flight per thread:
    // the walk is bound by the latency of the cached draws coming from L2, not by arithmetic)
    const uint32_t nb32 = (uint32_t)nblocks, md32 = (uint32_t)max_draws;  // max_draws < 2^32 (launcher check)
    for (uint32_t q0 = tid; q0 < nb32; q0 += 4 * SMB_THREADS) {
      uint4 d4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t q = q0 + u * SMB_THREADS;
        d4[u] = q < nb32 ? draws_of(q) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t q = q0 + u * SMB_THREADS;
        const uint32_t vv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const uint32_t k = (q << 2) + l;
          const uint32_t off = vv[l] - lo;  // 0xFFFFFFFF (no draw) never lands in the bucket: lo + width <= n < 2^31
          if (off < width && k < md32) atomicMin(&table[off], k);
        }
      }
    }
    __syncthreads();
    // walk 2: a draw is accepted iff it is that first occurrence (the cached draw values stream from L2, so
    // walking them twice is cheaper than remembering the matches behind a shared counter)
    for (uint32_t q0 = tid; q0 < nb32; q0 += 4 * SMB_THREADS) {
      uint4 d4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t q = q0 + u * SMB_THREADS;
        d4[u] = q < nb32 ? draws_of(q) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t q = q0 + u * SMB_THREADS;
        const uint32_t vv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
        uint32_t bits = 0u;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const uint32_t k = (q << 2) + l;
          const uint32_t off = vv[l] - lo;
          if (off < width && k < md32 && table[off] == k) bits |= 1u << l;
        }
        if (bits) atomicOr(&accept[q >> 3], bits << ((q & 7) << 2));
      }
    }
  }
  // last CTA of this job: per-chunk popcounts of the accept bitmask -> exclusive prefixes
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket_s = atomicAdd(job.ticket, 1u);
  __syncthreads();
  if (ticket_s != n_workers - 1) return;
  __threadfence();
  if (tid == 0) {
    carry_s = 0ull;
    *job.ticket = 0u;  // ready for the next use
  }
  __syncthreads();
  const unsigned long long nchunks = (max_draws + SMP_CHUNK - 1) / SMP_CHUNK;
  constexpr int WPC = SMP_CHUNK / 32;  // accept words per chunk
  const unsigned long long nwords = (max_draws + 31) >> 5;
  for (unsigned long long c0 = 0; c0 < nchunks; c0 += SMB_THREADS) {
    const unsigned long long ch = c0 + tid;
    unsigned int v = 0;
    if (ch < nchunks)
      for (int w = 0; w < WPC; ++w) {
        const unsigned long long wi = ch * WPC + w;
        if (wi < nwords) v += __popc(__ldcg(accept + wi));
      }
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
    for (int w = 0; w < SMB_THREADS / 32; ++w)
      if (w < wid) wbase += sh[w];
    const unsigned long long carry = carry_s;
    if (ch < nchunks) job.chunk_prefix[ch] = carry + wbase + (incl - v);
    __syncthreads();
    if (tid == SMB_THREADS - 1) carry_s = carry + wbase + incl;
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
  const FastMod fm = make_fastmod((uint32_t)job.n);
  const unsigned long long max_draws = job.max_draws;
  const unsigned long long nchunks = (max_draws + SMP_CHUNK - 1) / SMP_CHUNK;
  const unsigned long long nwords = (max_draws + 31) >> 5;
  const bool cached = job.draws != nullptr && max_draws <= job.draws_cap;
  uint32_t* __restrict__ accept = job.first;
  for (unsigned long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const unsigned long long q = ch * SMP_THREADS + tid;  // Philox block: draws 4q .. 4q+3, bits of ONE accept word
    const unsigned long long wi = q >> 3;
    const uint32_t word = (wi < nwords) ? __ldcg(accept + wi) : 0u;
    const uint32_t bits = (word >> ((q & 7) << 2)) & 0xFu;
    const unsigned int c = __popc(bits);
    unsigned int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();  // also: every thread of the chunk has read its accept word
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < SMP_THREADS / 32; ++w)
      if (w < wid) wbase += sh[w];
    if ((q & 7) == 0 && wi < nwords && word != 0u) accept[wi] = 0u;  // leave the bitmask clear for the next use
    unsigned long long rank = job.chunk_prefix[ch] + wbase + (incl - c);
    if (bits) {
      uint32_t vv[4];
      if (cached) {
        const uint4 d4 = __ldcg(reinterpret_cast<const uint4*>(job.draws) + q);
        vv[0] = d4.x; vv[1] = d4.y; vv[2] = d4.z; vv[3] = d4.w;
      } else {
        const Philox4 o = philox4x32_10(job.seed, job.domain, job.event, q);
#pragma unroll
        for (int l = 0; l < 4; ++l) vv[l] = draw_value(o.w[l], fm);
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if ((bits >> l) & 1u) {
          if (rank < count) {
            const uint32_t v = vv[l];
            out[rank] = v;
            if (job.post == 1) {
              // src_sampled/dst_sampled = unique endpoints of the sampled line vectors
              // (registration.cc:870-894); only the SET matters downstream, kept as per-point flags
              const uint2 e = job.edges[v];
              job.flags[e.x] = 1;
              job.flags[e.y] = 1;
            } else if (job.post == 2) {
              // basic line vectors (registration.cc:922-925) as endpoint pairs
              job.gathered[rank] = job.edges[job.via[v]];
            }
            if (rank == count - 1 && job.status) job.status[0] = (q << 2) + l + 1;  // draws consumed
          }
          ++rank;
        }
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

// 32-bit words of the accept bitmask (SampleJob::first) a job with n values / max_draws draws needs
unsigned long long sample_table_words(unsigned long long n, unsigned long long max_draws) {
  const unsigned long long w = (max_draws + 31) / 32 + 1;
  return w > n ? w : n;
}

// n_bound: upper bound of SampleJob::n over the jobs (sizes the bucket grid)
int launch_sample(cudaStream_t st, const SampleJob* d_jobs, int n_jobs, unsigned long long max_draws_bound,
                  unsigned long long n_bound) {
  if (n_jobs <= 0) return PSULVSB_OK;
  static bool attr_set = false;
  const size_t smem = (size_t)SMP_BW * 4;
  if (!attr_set) {
    PSU_CUDA(cudaFuncSetAttribute(sample_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const unsigned long long nchunks = (max_draws_bound + SMP_CHUNK - 1) / SMP_CHUNK;
  unsigned long long gx = nchunks;
  const unsigned long long cap = (unsigned long long)(148 * 16) / (unsigned long long)(n_jobs < 64 ? n_jobs : 64) + 1;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  sample_draws_kernel<<<dim3((unsigned)gx, (unsigned)n_jobs), 256, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("sample_draws_kernel");
  // bucket CTAs: one per SM at a time; a CTA walks several buckets when the batch alone fills the GPU
  unsigned long long nb = (n_bound + SMP_BW - 1) / SMP_BW;
  if (nb < 1) nb = 1;
  unsigned long long bx = (148ull * 3ull + (unsigned long long)n_jobs - 1) / (unsigned long long)n_jobs;
  if (bx > nb) bx = nb;
  if (bx < 1) bx = 1;
  sample_bucket_kernel<<<dim3((unsigned)bx, (unsigned)n_jobs), SMB_THREADS, smem, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("sample_bucket_kernel");
  sample_emit_kernel<<<dim3((unsigned)gx, (unsigned)n_jobs), SMP_THREADS, 0, st>>>(d_jobs);
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
