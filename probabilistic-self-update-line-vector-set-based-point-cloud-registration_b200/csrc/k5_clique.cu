// k5_clique.cu -- inlier-graph clique for the last-resort escalation of the solver
// (registration.cc:1000-1085, reached only at the rate pair (1.0, 1.0)): vertices = correspondences,
// edges = the scale-consistent line vectors of the round, clique -> the points handed to the
// translation solver (registration.cc:1238-1244).
//
// The reference calls PMC (teaser/src/graph.cc:12-125; an un-vendored, unpinned dependency whose
// result is not unique), so parity for this branch is unpinned; what is built here is a
// deterministic greedy maximal clique on a bit-matrix adjacency: repeatedly take the candidate with
// the most neighbours among the remaining candidates (ties: lowest index) and intersect the
// candidate set with its adjacency row.  On registration graphs (one large planted clique of
// mutually consistent inliers over a sparse random background) this returns the planted clique.
#include <cuda_runtime.h>

#include "common.cuh"
#include "engine.cuh"

namespace psulvsb {

namespace {

__global__ void __launch_bounds__(256) clique_zero_kernel(const CliqueJob* __restrict__ jobs) {
  const CliqueJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  const size_t words = (size_t)job.n_vertices * job.stride;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < words; i += (size_t)gridDim.x * 256) job.adj[i] = 0u;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < job.n_vertices; i += gridDim.x * 256) job.flags[i] = 0;
}

__global__ void __launch_bounds__(256) clique_edges_kernel(const CliqueJob* __restrict__ jobs) {
  const CliqueJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  for (unsigned long long k = (unsigned long long)blockIdx.x * 256 + threadIdx.x; k < job.n_edges;
       k += (unsigned long long)gridDim.x * 256) {
    const uint2 e = job.edges[k];
    if (e.x == e.y) continue;
    if (job.filter) {
      // ScaleInliersSelector on the line vector (registration.cc:425-433), FP64, reference order
      const double* sa = job.src + 3 * (size_t)e.x;
      const double* sb = job.src + 3 * (size_t)e.y;
      const double* ta = job.dst + 3 * (size_t)e.x;
      const double* tb = job.dst + 3 * (size_t)e.y;
      const double a = sqrt(sqnorm3(dsub(sb[0], sa[0]), dsub(sb[1], sa[1]), dsub(sb[2], sa[2])));
      const double b = sqrt(sqnorm3(dsub(tb[0], ta[0]), dsub(tb[1], ta[1]), dsub(tb[2], ta[2])));
      if (!(fabs(dsub(a, b)) <= job.beta)) continue;
    }
    atomicOr(&job.adj[(size_t)e.x * job.stride + (e.y >> 5)], 1u << (e.y & 31));
    atomicOr(&job.adj[(size_t)e.y * job.stride + (e.x >> 5)], 1u << (e.x & 31));
  }
}

__global__ void __launch_bounds__(1024) clique_greedy_kernel(const CliqueJob* __restrict__ jobs) {
  const CliqueJob& job = jobs[blockIdx.x];
  if (!job.active) return;
  extern __shared__ uint32_t P[];  // candidate set, one bit per vertex
  __shared__ unsigned long long warp_best[32];
  __shared__ unsigned long long pick_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = job.n_vertices, W = job.stride;
  const uint32_t* __restrict__ adj = job.adj;
  for (int w = tid; w < W; w += 1024) P[w] = 0u;
  __syncthreads();
  // candidates: every vertex with at least one edge
  for (int v = wid; v < n; v += 32) {
    uint32_t any = 0u;
    for (int w = lane; w < W; w += 32) any |= adj[(size_t)v * W + w];
    any = __reduce_or_sync(0xffffffffu, any);
    if (lane == 0 && any) atomicOr(&P[v >> 5], 1u << (v & 31));
  }
  __syncthreads();
  int size = 0;
  while (true) {
    unsigned long long best = 0ull;  // (count + 1) << 32 | (0xFFFFFFFF - v): most neighbours, then lowest index
    for (int v = wid; v < n; v += 32) {
      if (!((P[v >> 5] >> (v & 31)) & 1u)) continue;  // warp-uniform
      int c = 0;
      for (int w = lane; w < W; w += 32) c += __popc(adj[(size_t)v * W + w] & P[w]);
      c = __reduce_add_sync(0xffffffffu, c);
      const unsigned long long key = ((unsigned long long)(c + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)v);
      best = key > best ? key : best;
    }
    if (lane == 0) warp_best[wid] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long b = 0ull;
      for (int w = 0; w < 32; ++w) b = warp_best[w] > b ? warp_best[w] : b;
      pick_s = b;
    }
    __syncthreads();
    const unsigned long long pick = pick_s;
    if (pick == 0ull) break;
    const int v = (int)(0xFFFFFFFFu - (uint32_t)(pick & 0xFFFFFFFFull));
    if (tid == 0) job.flags[v] = 1;
    ++size;
    for (int w = tid; w < W; w += 1024) P[w] &= adj[(size_t)v * W + w];  // v itself drops out: no self loops
    __syncthreads();
  }
  if (tid == 0) *job.size = size;
}

}  // namespace

int launch_greedy_clique(cudaStream_t st, const CliqueJob* d_jobs, int n_jobs, int max_vertices, int max_stride,
                         unsigned long long max_edges) {
  if (n_jobs <= 0 || max_vertices < 1) return PSULVSB_OK;
  const size_t smem = (size_t)max_stride * sizeof(uint32_t);
  if (smem > 200 * 1024) return fail(PSULVSB_ERR_UNSUPPORTED, "greedy clique: more than 1.6 M vertices");
  static bool attr_set = false;
  if (!attr_set) {
    PSU_CUDA(cudaFuncSetAttribute(clique_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const size_t words = (size_t)max_vertices * max_stride;
  unsigned long long gz = (words + 256 * 8 - 1) / (256 * 8);
  if (gz > 148 * 8) gz = 148 * 8;
  if (gz < 1) gz = 1;
  clique_zero_kernel<<<dim3((unsigned)gz, (unsigned)n_jobs), 256, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_zero_kernel");
  unsigned long long ge = (max_edges + 255) / 256;
  if (ge > 148 * 8) ge = 148 * 8;
  if (ge < 1) ge = 1;
  clique_edges_kernel<<<dim3((unsigned)ge, (unsigned)n_jobs), 256, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_edges_kernel");
  clique_greedy_kernel<<<n_jobs, 1024, smem, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_greedy_kernel");
  return PSULVSB_OK;
}

}  // namespace psulvsb
