// k5_clique.cu -- inlier-graph clique for the last-resort escalation of the solver
// (registration.cc:1000-1085, reached only at the rate pair (1.0, 1.0)): vertices = correspondences,
// edges = the scale-consistent line vectors of the round, clique -> the points handed to the
// translation solver (registration.cc:1238-1244).
//
// The reference calls PMC (teaser/src/graph.cc:12-125; an un-vendored, unpinned dependency) for an
// exact maximum clique; which maximum clique it returns is not pinned by anything in the reference.  Position
// taken (the same in the CPU oracle): of all cliques of the maximum size, the one whose ascending vertex list is
// lexicographically smallest -- so size AND members are comparable.  Three steps on a bit-matrix adjacency:
//  1. a deterministic greedy maximal clique (repeatedly take the candidate with the most neighbours
//     among the remaining candidates, ties: lowest index, and intersect the candidate set with its
//     adjacency row): the lower bound lb.  On registration graphs with consensus (one large planted
//     clique over a sparse random background) this already is the planted clique;
//  2. an exact improvement search: every vertex v of degree >= lb is the root of a branch and bound
//     over its later core neighbours (one warp per root, the neighbourhood relabelled into a local
//     <= 512-vertex bit matrix in shared memory, an explicit stack, the bound |clique| + |candidates|
//     <= best).  Roots only use lb and their own improvements, so the result does not depend on
//     scheduling: the largest size wins, ties go to the lowest root, and a second launch replays that
//     root to write the members.  Neighbourhoods above 512 vertices that the bound does not cut, or a
//     root that exhausts its node budget, leave the greedy answer in place and clear *proven;
//  3. the canonical members: with the size omega known, every vertex of degree >= omega - 1 is tried as the
//     SMALLEST member -- a depth-first search over its later neighbours, candidates ascending, cut only where
//     omega cannot be reached, stops at its first clique of size omega, which is the lexicographically smallest
//     one with that root; the lowest successful root wins (atomicMin) and is replayed to write the flags.
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

// ---- exact improvement ------------------------------------------------------------------------
constexpr int CX_D = 512;                    // local neighbourhood cap
constexpr int CX_W = CX_D / 32;              // words per local row (one per lane of a half warp)
constexpr unsigned int CX_BUDGET = 1u << 22; // search nodes per root
constexpr size_t CX_SMEM = sizeof(uint32_t) * ((size_t)CX_D + (size_t)CX_D * CX_W + (size_t)(CX_D + 1) * CX_W) + sizeof(uint16_t) * CX_D;

// work area behind the bit matrix: [0..1] best key (size << 32 | ~root), [2] unproven
__device__ __forceinline__ unsigned long long* cx_key(const CliqueJob& job) {
  return reinterpret_cast<unsigned long long*>(job.adj + (((size_t)job.n_vertices * job.stride + 3) & ~(size_t)3));
}

// the maximum size known after the improvement search: the greedy bound or what a root found beyond it
__device__ __forceinline__ int cx_omega(const CliqueJob& job) {
  const int found = (int)(cx_key(job)[0] >> 32);
  const int lb = *job.size;
  return found > lb ? found : lb;
}

// phase 0 -- flags bit 1 (value 2): degree >= lb (can belong to a clique larger than the greedy one); work area reset.
// phase 1 -- flags bit 2 (value 4): degree >= omega - 1 (can belong to a clique of the maximum size).
__global__ void __launch_bounds__(256) clique_core_kernel(const CliqueJob* __restrict__ jobs, int phase) {
  const CliqueJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  const int lane = threadIdx.x & 31;
  const int lb = *job.size;
  const int thr = phase == 0 ? lb : cx_omega(job) - 1;
  const int bit = phase == 0 ? 2 : 4;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long* key = cx_key(job);
    if (phase == 0) {
      key[0] = 0ull;
      key[1] = 0ull;
      *job.proven = 1;
    }
    key[2] = ~0ull;  // lowest root with a clique of the maximum size
  }
  for (int v = blockIdx.x * 8 + (threadIdx.x >> 5); v < job.n_vertices; v += gridDim.x * 8) {
    int deg = 0;
    for (int w = lane; w < job.stride; w += 32) deg += __popc(job.adj[(size_t)v * job.stride + w]);
    deg = __reduce_add_sync(0xffffffffu, deg);
    if (lane == 0) job.flags[v] = (uint8_t)((job.flags[v] & ~bit) | ((deg >= thr && lb > 0) ? bit : 0));
  }
}

struct CxSmem {
  uint32_t* list;    // [CX_D] global ids of the root's later core neighbours, ascending
  uint32_t* adjl;    // [CX_D][CX_W] local bit matrix
  uint32_t* stack;   // [CX_D + 1][CX_W] candidate sets per depth
  uint16_t* chosen;  // [CX_D]
};

// Branch and bound below root v.  Returns the best clique size found (> floor_size) or 0; target != 0: stop at the
// first clique of exactly that size and leave its local members in sm.chosen[0 .. target - 2].
__device__ int cx_search_root(const CliqueJob& job, const CxSmem& sm, int v, int floor_size, int target, int* status,
                              int core_bit) {
  const int lane = threadIdx.x;
  const int W = job.stride;
  const uint32_t* __restrict__ adj = job.adj;
  // later core neighbours of v, ascending
  int d = 0;
  for (int w0 = v >> 5; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = (w < W) ? adj[(size_t)v * W + w] : 0u;
    if (w == (v >> 5)) bits &= ~((2u << (v & 31)) - 1u);  // strictly above v
    // keep core vertices only
    uint32_t keep = 0u;
    for (uint32_t b = bits; b; b &= b - 1u) {
      const int u = w * 32 + (__ffs(b) - 1);
      if (job.flags[u] & core_bit) keep |= 1u << (u & 31);
    }
    const int c = __popc(keep);
    int pre = c;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    int at = d + pre - c;
    for (uint32_t b = keep; b; b &= b - 1u) {
      if (at < CX_D) sm.list[at] = (uint32_t)(w * 32 + (__ffs(b) - 1));
      ++at;
    }
    d += total;
  }
  __syncwarp();
  if (1 + d <= floor_size) return 0;
  if (d > CX_D) {
    *status = 1;
    return 0;
  }
  // local bit matrix: adjl[a][wl] bit k = adj[list[a]][list[wl * 32 + k]]
  const int WL = (d + 31) >> 5;
  for (int a = 0; a < d; ++a) {
    const uint32_t* __restrict__ row = adj + (size_t)sm.list[a] * W;
    for (int wl = 0; wl < WL; ++wl) {
      const int b = wl * 32 + lane;
      uint32_t bit = 0u;
      if (b < d) {
        const uint32_t u = sm.list[b];
        bit = (row[u >> 5] >> (u & 31)) & 1u;
      }
      const uint32_t word = __ballot_sync(0xffffffffu, bit != 0u);
      if (lane == 0) sm.adjl[a * CX_W + wl] = word;
    }
    if (lane >= WL && lane < CX_W) sm.adjl[a * CX_W + lane] = 0u;
  }
  // depth 0 candidates: everything
  if (lane < CX_W) {
    uint32_t w = 0u;
    if (lane < (d >> 5))
      w = 0xFFFFFFFFu;
    else if (lane == (d >> 5))
      w = (d & 31) ? ((1u << (d & 31)) - 1u) : 0u;
    sm.stack[lane] = w;
  }
  __syncwarp();
  int best = floor_size, depth = 0;
  unsigned int nodes = 0;
  while (true) {
    const uint32_t Pw = (lane < CX_W) ? sm.stack[depth * CX_W + lane] : 0u;
    const int cnt = __reduce_add_sync(0xffffffffu, __popc(Pw));
    if (cnt == 0 || 1 + depth + cnt <= best) {  // nothing below this node can beat best
      if (depth == 0) break;
      --depth;
      continue;
    }
    const uint32_t has = __ballot_sync(0xffffffffu, Pw != 0u);
    const int l0 = __ffs(has) - 1;
    const uint32_t w0 = __shfl_sync(0xffffffffu, Pw, l0);
    const int u = l0 * 32 + (__ffs(w0) - 1);
    uint32_t rest = Pw;
    if (lane == l0) {
      rest &= ~(1u << (u & 31));
      sm.stack[depth * CX_W + lane] = rest;  // siblings after u never see u again
    }
    const uint32_t Nw = (lane < CX_W) ? (rest & sm.adjl[u * CX_W + lane]) : 0u;
    const int ncnt = __reduce_add_sync(0xffffffffu, __popc(Nw));
    if (lane == 0) sm.chosen[depth] = (uint16_t)u;
    const int size = 2 + depth;  // root + chosen[0 .. depth]
    if (++nodes > CX_BUDGET) {
      *status = 1;
      break;
    }
    if (ncnt == 0) {
      if (size > best) {
        best = size;
        if (target && size == target) {
          __syncwarp();
          return size;
        }
      }
      continue;
    }
    if (size + ncnt <= best) continue;
    if (lane < CX_W) sm.stack[(depth + 1) * CX_W + lane] = Nw;
    ++depth;
    __syncwarp();
  }
  return best > floor_size ? best : 0;
}

__device__ __forceinline__ CxSmem cx_carve(uint32_t* base) {
  CxSmem sm;
  sm.list = base;
  sm.adjl = sm.list + CX_D;
  sm.stack = sm.adjl + CX_D * CX_W;
  sm.chosen = reinterpret_cast<uint16_t*>(sm.stack + (CX_D + 1) * CX_W);
  return sm;
}

__global__ void __launch_bounds__(32) clique_exact_kernel(const CliqueJob* __restrict__ jobs) {
  const CliqueJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  extern __shared__ uint32_t cx_raw[];
  const CxSmem sm = cx_carve(cx_raw);
  const int lb = *job.size;
  if (lb < 1) return;
  unsigned long long* key = cx_key(job);
  for (int v = blockIdx.x; v < job.n_vertices; v += gridDim.x) {
    if (!(job.flags[v] & 2)) continue;
    int status = 0;
    const int s = cx_search_root(job, sm, v, lb, 0, &status, 2);
    if (threadIdx.x == 0) {
      if (status) atomicExch(reinterpret_cast<unsigned int*>(key + 1), 1u);
      if (s > lb) atomicMax(key, ((unsigned long long)s << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)v));
    }
    __syncwarp();
  }
}

// step 3: the lowest root that is the smallest member of a clique of the maximum size
__global__ void __launch_bounds__(32) clique_canon_kernel(const CliqueJob* __restrict__ jobs) {
  const CliqueJob& job = jobs[blockIdx.y];
  if (!job.active) return;
  extern __shared__ uint32_t cx_raw[];
  const CxSmem sm = cx_carve(cx_raw);
  if (*job.size < 1) return;
  const int omega = cx_omega(job);
  unsigned long long* key = cx_key(job);
  for (int v = blockIdx.x; v < job.n_vertices; v += gridDim.x) {
    if (!(job.flags[v] & 4)) continue;
    // (a root above the current minimum cannot win; skipping it does not change the minimum)
    if ((unsigned long long)v > *reinterpret_cast<volatile unsigned long long*>(key + 2)) continue;
    int status = 0;
    const int s = cx_search_root(job, sm, v, omega - 1, omega, &status, 4);
    if (threadIdx.x == 0) {
      if (status) atomicExch(reinterpret_cast<unsigned int*>(key + 1), 1u);
      if (s == omega) atomicMin(key + 2, (unsigned long long)v);
    }
    __syncwarp();
  }
}

// replay the winning root and write the members
__global__ void __launch_bounds__(32) clique_record_kernel(const CliqueJob* __restrict__ jobs) {
  const CliqueJob& job = jobs[blockIdx.x];
  if (!job.active) return;
  extern __shared__ uint32_t cx_raw[];
  const CxSmem sm = cx_carve(cx_raw);
  const int lane = threadIdx.x;
  if (*job.size < 1) return;
  unsigned long long* key = cx_key(job);
  if (lane == 0 && (unsigned int)key[1] != 0u) *job.proven = 0;
  const int target = cx_omega(job);
  const unsigned long long r = key[2];
  if (r == ~0ull) {
    // no root reported a clique of the known size (only possible when a search gave up): the greedy clique stands
    for (int v = lane; v < job.n_vertices; v += 32) job.flags[v] &= 1;
    if (lane == 0) *job.proven = 0;
    return;
  }
  const int root = (int)r;
  int status = 0;
  const int s = cx_search_root(job, sm, root, target - 1, target, &status, 4);
  __syncwarp();
  for (int v = lane; v < job.n_vertices; v += 32) job.flags[v] = 0;
  __syncwarp();
  if (s == target) {
    if (lane == 0) {
      job.flags[root] = 1;
      *job.size = target;
    }
    for (int i = lane; i < target - 1; i += 32) job.flags[sm.list[sm.chosen[i]]] = 1;
  } else if (lane == 0) {
    *job.size = 0;  // cannot happen: the same deterministic search found it one launch earlier
    *job.proven = 0;
  }
}

}  // namespace

int launch_max_clique(cudaStream_t st, const CliqueJob* d_jobs, int n_jobs, int max_vertices, int max_stride,
                      unsigned long long max_edges, bool exact) {
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
  if (gz > sm_count() * 8) gz = sm_count() * 8;
  if (gz < 1) gz = 1;
  clique_zero_kernel<<<dim3((unsigned)gz, (unsigned)n_jobs), 256, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_zero_kernel");
  unsigned long long ge = (max_edges + 255) / 256;
  if (ge > sm_count() * 8) ge = sm_count() * 8;
  if (ge < 1) ge = 1;
  clique_edges_kernel<<<dim3((unsigned)ge, (unsigned)n_jobs), 256, 0, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_edges_kernel");
  clique_greedy_kernel<<<n_jobs, 1024, smem, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_greedy_kernel");
  if (!exact) return PSULVSB_OK;
  static bool attr2_set = false;
  if (!attr2_set) {
    PSU_CUDA(cudaFuncSetAttribute(clique_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CX_SMEM));
    PSU_CUDA(cudaFuncSetAttribute(clique_record_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CX_SMEM));
    PSU_CUDA(cudaFuncSetAttribute(clique_canon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CX_SMEM));
    attr2_set = true;
  }
  int gc = (max_vertices + 7) / 8;
  if (gc > sm_count() * 8) gc = sm_count() * 8;
  clique_core_kernel<<<dim3((unsigned)gc, (unsigned)n_jobs), 256, 0, st>>>(d_jobs, 0);
  PSU_CHECK_LAUNCH("clique_core_kernel");
  int gx = max_vertices < sm_count() * 3 ? max_vertices : sm_count() * 3;
  clique_exact_kernel<<<dim3((unsigned)gx, (unsigned)n_jobs), 32, CX_SMEM, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_exact_kernel");
  clique_core_kernel<<<dim3((unsigned)gc, (unsigned)n_jobs), 256, 0, st>>>(d_jobs, 1);
  PSU_CHECK_LAUNCH("clique_core_kernel");
  clique_canon_kernel<<<dim3((unsigned)gx, (unsigned)n_jobs), 32, CX_SMEM, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_canon_kernel");
  clique_record_kernel<<<n_jobs, 32, CX_SMEM, st>>>(d_jobs);
  PSU_CHECK_LAUNCH("clique_record_kernel");
  return PSULVSB_OK;
}

}  // namespace psulvsb
