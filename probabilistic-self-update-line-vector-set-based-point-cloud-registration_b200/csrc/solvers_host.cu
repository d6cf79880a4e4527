// solvers_host.cu -- the reference's public sub-solver calls as C entry points on HOST buffers
// (teaser/include/teaser/registration.h:107-317; used by the reference's own unit tests
// rotation-solver-test.cc, scale-solver-test.cc, translation-solver-test.cc and by
// RobustRegistrationSolver::solveForScale / solveForRotation / solveForTranslation / computeTIMs).
// Each call stages its arrays on the device, runs the same kernels the engine runs, and copies the result
// back; there is no CPU implementation behind them.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "common.cuh"
#include "engine.cuh"
#include "solve_dev.cuh"

namespace psulvsb {
namespace {

// all device scratch of one call; freed on scope exit
struct Scratch {
  void* p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n = 0;
  template <typename T>
  T* take(size_t count) {
    void* q = nullptr;
    if (n >= 8 || cudaMalloc(&q, sizeof(T) * (count ? count : 1)) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    p[n++] = q;
    return static_cast<T*>(q);
  }
  ~Scratch() {
    for (int i = 0; i < n; ++i) cudaFree(p[i]);
  }
};

int have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(PSULVSB_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
  }
  return PSULVSB_OK;
}

// registration.cc:471-505: column l = row_offset(i) + (j - i - 1) holds v_j - v_i, map = (i, j)
__global__ void __launch_bounds__(256) tims_kernel(const double* __restrict__ v, int n, double* __restrict__ tims,
                                                   int* __restrict__ map) {
  const int i = blockIdx.x;
  if (i >= n - 1) return;
  const unsigned long long base = (unsigned long long)i * (unsigned long long)n - (unsigned long long)i * (i + 1ull) / 2ull;
  const double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
  for (int j = i + 1 + threadIdx.x; j < n; j += 256) {
    const unsigned long long l = base + (unsigned long long)(j - i - 1);
    tims[3 * l + 0] = dsub(v[3 * j + 0], x);
    tims[3 * l + 1] = dsub(v[3 * j + 1], y);
    tims[3 * l + 2] = dsub(v[3 * j + 2], z);
    if (map) {
      map[2 * l + 0] = i;
      map[2 * l + 1] = j;
    }
  }
}

// registration.cc:418-434
__global__ void __launch_bounds__(256) scale_inliers_kernel(const double* __restrict__ s, const double* __restrict__ t,
                                                            unsigned long long n, double beta, uint8_t* __restrict__ out) {
  for (unsigned long long l = (unsigned long long)blockIdx.x * 256 + threadIdx.x; l < n;
       l += (unsigned long long)gridDim.x * 256) {
    const double a = sqrt(sqnorm3(s[3 * l], s[3 * l + 1], s[3 * l + 2]));
    const double b = sqrt(sqnorm3(t[3 * l], t[3 * l + 1], t[3 * l + 2]));
    out[l] = fabs(dsub(a, b)) <= beta ? 1 : 0;
  }
}

// registration.cc:397-415 + :66-120 on explicit line vectors
__global__ void __launch_bounds__(BLK) tls_scale_kernel(const double* __restrict__ s, const double* __restrict__ t, int K,
                                                        double beta, uint64_t seed, uint32_t event, int use_last,
                                                        double last_s, double* __restrict__ X, double* __restrict__ A,
                                                        double* __restrict__ scale_out, uint8_t* __restrict__ inliers) {
  __shared__ BlockScratch scratch;
  __shared__ double est_s, scale_s;
  for (int i = threadIdx.x; i < K; i += BLK) {
    const double v1 = sqrt(sqnorm3(s[3 * i], s[3 * i + 1], s[3 * i + 2]));
    const double v2 = sqrt(sqnorm3(t[3 * i], t[3 * i + 1], t[3 * i + 2]));
    X[i] = v2 / v1;
    A[i] = dmul(beta, 1.0 / v1);
  }
  block_tls_scale(&scratch, X, A, K, seed, event, use_last != 0, last_s, 1.0, &est_s, &scale_s);
  // registration.cc:104: the mask is taken against the UNREFINED estimate, the weighted mean follows it
  const double est = est_s;
  for (int i = threadIdx.x; i < K; i += BLK) inliers[i] = fabs(dsub(X[i], est)) <= A[i] ? 1 : 0;
  if (threadIdx.x == 0) {
    scale_out[0] = scale_s;
    scale_out[1] = est;
  }
}

// registration.cc:196-202 per axis, ANDed over the axes (:457-462)
__global__ void __launch_bounds__(256) translation_inliers_kernel(const double* __restrict__ s, const double* __restrict__ d,
                                                                  int n, const double* __restrict__ t, double noise,
                                                                  uint8_t* __restrict__ out) {
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    bool ok = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) ok = ok && (fabs(dsub(dsub(d[3 * i + r], s[3 * i + r]), t[r])) <= noise);
    out[i] = ok ? 1 : 0;
  }
}

__global__ void gnc_points_kernel(const double* __restrict__ tims, unsigned long long K, double* __restrict__ pts,
                                  uint2* __restrict__ edges) {
  // point 0 = origin, point k + 1 = line vector k; edge (0, k + 1): s[b] - s[a] == the line vector, exactly
  for (unsigned long long k = (unsigned long long)blockIdx.x * 256 + threadIdx.x; k < K;
       k += (unsigned long long)gridDim.x * 256) {
    pts[3 * (k + 1) + 0] = tims[3 * k + 0];
    pts[3 * (k + 1) + 1] = tims[3 * k + 1];
    pts[3 * (k + 1) + 2] = tims[3 * k + 2];
    if (edges) edges[k] = make_uint2(0u, (unsigned)(k + 1));
  }
  if (blockIdx.x == 0 && threadIdx.x < 3) pts[threadIdx.x] = 0.0;
}

unsigned grid_for(unsigned long long n) {
  unsigned long long g = (n + 255) / 256;
  if (g > (unsigned long long)sm_count() * 16) g = (unsigned long long)sm_count() * 16;
  return (unsigned)(g ? g : 1);
}

}  // namespace
}  // namespace psulvsb

using namespace psulvsb;

#define PSU_SYNC(what)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = cudaDeviceSynchronize();                                                           \
    if (e__ != cudaSuccess) return fail(PSULVSB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e__)); \
  } while (0)

extern "C" {

int psulvsb_compute_tims_host(const double* pts, int n, double* tims, int* map) {
  if (int rc = have_device()) return rc;
  if (!pts || !tims || n < 0) return fail(PSULVSB_ERR_INVALID, "psulvsb_compute_tims_host: bad argument");
  if (n < 2) return PSULVSB_OK;
  const unsigned long long L = (unsigned long long)n * (unsigned long long)(n - 1) / 2ull;
  Scratch sc;
  double* d_v = sc.take<double>(3 * (size_t)n);
  double* d_t = sc.take<double>(3 * (size_t)L);
  int* d_m = map ? sc.take<int>(2 * (size_t)L) : nullptr;
  if (!d_v || !d_t || (map && !d_m)) return fail(PSULVSB_ERR_CUDA, "psulvsb_compute_tims_host: device allocation failed");
  PSU_CUDA(cudaMemcpy(d_v, pts, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  tims_kernel<<<(unsigned)(n - 1), 256>>>(d_v, n, d_t, d_m);
  PSU_CHECK_LAUNCH("tims_kernel");
  PSU_CUDA(cudaMemcpy(tims, d_t, sizeof(double) * 3 * (size_t)L, cudaMemcpyDeviceToHost));
  if (map) PSU_CUDA(cudaMemcpy(map, d_m, sizeof(int) * 2 * (size_t)L, cudaMemcpyDeviceToHost));
  return PSULVSB_OK;
}

int psulvsb_scale_inliers_host(const double* src_tims, const double* dst_tims, unsigned long long n, double noise_bound,
                               double cbar2, unsigned char* inliers) {
  if (int rc = have_device()) return rc;
  if ((n && (!src_tims || !dst_tims || !inliers))) return fail(PSULVSB_ERR_INVALID, "psulvsb_scale_inliers_host: NULL array");
  if (n == 0) return PSULVSB_OK;
  Scratch sc;
  double* d_s = sc.take<double>(3 * (size_t)n);
  double* d_t = sc.take<double>(3 * (size_t)n);
  uint8_t* d_o = sc.take<uint8_t>((size_t)n);
  if (!d_s || !d_t || !d_o) return fail(PSULVSB_ERR_CUDA, "psulvsb_scale_inliers_host: device allocation failed");
  PSU_CUDA(cudaMemcpy(d_s, src_tims, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  PSU_CUDA(cudaMemcpy(d_t, dst_tims, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  const double beta = 2.0 * noise_bound * std::sqrt(cbar2);
  scale_inliers_kernel<<<grid_for(n), 256>>>(d_s, d_t, n, beta, d_o);
  PSU_CHECK_LAUNCH("scale_inliers_kernel");
  PSU_CUDA(cudaMemcpy(inliers, d_o, (size_t)n, cudaMemcpyDeviceToHost));
  return PSULVSB_OK;
}

int psulvsb_tls_scale_host(const double* src_tims, const double* dst_tims, int n, double noise_bound, double cbar2,
                           uint64_t seed, uint32_t event, const double* last_best_scale, double* scale,
                           unsigned char* inliers) {
  if (int rc = have_device()) return rc;
  if (!src_tims || !dst_tims || !scale || n < 1) return fail(PSULVSB_ERR_INVALID, "psulvsb_tls_scale_host: bad argument");
  Scratch sc;
  double* d_s = sc.take<double>(3 * (size_t)n);
  double* d_t = sc.take<double>(3 * (size_t)n);
  double* d_x = sc.take<double>((size_t)n);
  double* d_a = sc.take<double>((size_t)n);
  double* d_out = sc.take<double>(2);
  uint8_t* d_o = sc.take<uint8_t>((size_t)n);
  if (!d_s || !d_t || !d_x || !d_a || !d_out || !d_o)
    return fail(PSULVSB_ERR_CUDA, "psulvsb_tls_scale_host: device allocation failed");
  PSU_CUDA(cudaMemcpy(d_s, src_tims, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  PSU_CUDA(cudaMemcpy(d_t, dst_tims, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  const double beta = 2.0 * noise_bound * std::sqrt(cbar2);
  tls_scale_kernel<<<1, BLK>>>(d_s, d_t, n, beta, seed, event, last_best_scale ? 1 : 0,
                               last_best_scale ? *last_best_scale : 0.0, d_x, d_a, d_out, d_o);
  PSU_CHECK_LAUNCH("tls_scale_kernel");
  double out[2];
  PSU_CUDA(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
  *scale = out[0];
  if (inliers) PSU_CUDA(cudaMemcpy(inliers, d_o, (size_t)n, cudaMemcpyDeviceToHost));
  return PSULVSB_OK;
}

int psulvsb_gnc_tls_rotation_host(const double* src_tims, const double* dst_tims, unsigned long long n,
                                  double noise_bound, int max_iterations, double gnc_factor, double cost_threshold,
                                  const double* R_last_best, double* R, unsigned char* inliers, double* cost,
                                  int* iterations) {
  if (int rc = have_device()) return rc;
  if (!src_tims || !dst_tims || !R || n < 1) return fail(PSULVSB_ERR_INVALID, "psulvsb_gnc_tls_rotation_host: bad argument");
  if (n >= 0x7FFFFFF0ull) return fail(PSULVSB_ERR_UNSUPPORTED, "psulvsb_gnc_tls_rotation_host: n must fit 31 bits");
  Scratch sc;
  double* d_in = sc.take<double>(3 * (size_t)n);
  double* d_ps = sc.take<double>(3 * (size_t)(n + 1));
  double* d_pt = sc.take<double>(3 * (size_t)(n + 1));
  uint2* d_e = sc.take<uint2>((size_t)n);
  double* d_w = sc.take<double>((size_t)n);
  uint8_t* d_o = sc.take<uint8_t>((size_t)n);
  double* d_small = sc.take<double>(32);  // [0..8] R, [9] cost, [10..18] R_init, [20..] info (ints)
  if (!d_in || !d_ps || !d_pt || !d_e || !d_w || !d_o || !d_small)
    return fail(PSULVSB_ERR_CUDA, "psulvsb_gnc_tls_rotation_host: device allocation failed");
  PSU_CUDA(cudaMemcpy(d_in, src_tims, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  gnc_points_kernel<<<grid_for(n), 256>>>(d_in, n, d_ps, d_e);
  PSU_CHECK_LAUNCH("gnc_points_kernel");
  PSU_SYNC("psulvsb_gnc_tls_rotation_host");
  PSU_CUDA(cudaMemcpy(d_in, dst_tims, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  gnc_points_kernel<<<grid_for(n), 256>>>(d_in, n, d_pt, nullptr);
  PSU_CHECK_LAUNCH("gnc_points_kernel");
  PSU_CUDA(cudaMemset(d_small, 0, sizeof(double) * 32));
  if (R_last_best) PSU_CUDA(cudaMemcpy(d_small + 10, R_last_best, sizeof(double) * 9, cudaMemcpyHostToDevice));
  int* d_info = reinterpret_cast<int*>(d_small + 20);
  if (int rc = psulvsb_gnc_tls_rotation(nullptr, d_ps, d_pt, d_e, n, 1.0, noise_bound, max_iterations, gnc_factor,
                                        cost_threshold, R_last_best ? d_small + 10 : nullptr, d_w, d_small, d_o, d_info,
                                        d_small + 9))
    return rc;
  PSU_SYNC("psulvsb_gnc_tls_rotation_host");
  double out[10];
  int info[4];
  PSU_CUDA(cudaMemcpy(out, d_small, sizeof(out), cudaMemcpyDeviceToHost));
  PSU_CUDA(cudaMemcpy(info, d_info, sizeof(info), cudaMemcpyDeviceToHost));
  std::memcpy(R, out, sizeof(double) * 9);
  if (cost) *cost = out[9];
  if (iterations) *iterations = info[0];
  if (inliers) PSU_CUDA(cudaMemcpy(inliers, d_o, (size_t)n, cudaMemcpyDeviceToHost));
  return PSULVSB_OK;
}

int psulvsb_tls_translation_host(const double* src, const double* dst, int n, double noise_bound, double cbar2,
                                 const double* t_last_best, double* t, unsigned char* inliers) {
  if (int rc = have_device()) return rc;
  if (!src || !dst || !t || n < 1) return fail(PSULVSB_ERR_INVALID, "psulvsb_tls_translation_host: bad argument");
  Scratch sc;
  double* d_s = sc.take<double>(3 * (size_t)n);
  double* d_d = sc.take<double>(3 * (size_t)n);
  uint8_t* d_f = sc.take<uint8_t>((size_t)n);
  uint8_t* d_o = sc.take<uint8_t>((size_t)n);
  double* d_small = sc.take<double>(24);  // [0..8] identity, [9..11] t, [12..14] last best, [16] n_points (int)
  if (!d_s || !d_d || !d_f || !d_o || !d_small)
    return fail(PSULVSB_ERR_CUDA, "psulvsb_tls_translation_host: device allocation failed");
  PSU_CUDA(cudaMemcpy(d_s, src, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  PSU_CUDA(cudaMemcpy(d_d, dst, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  PSU_CUDA(cudaMemset(d_f, 1, (size_t)n));
  double small[24] = {0};
  small[0] = small[4] = small[8] = 1.0;
  if (t_last_best) std::memcpy(small + 12, t_last_best, sizeof(double) * 3);
  PSU_CUDA(cudaMemcpy(d_small, small, sizeof(small), cudaMemcpyHostToDevice));
  const double noise = noise_bound * std::sqrt(cbar2);
  if (int rc = launch_tls_translation(nullptr, d_s, d_d, d_f, n, 1.0, d_small, noise, t_last_best ? d_small + 12 : nullptr,
                                      d_small + 9, reinterpret_cast<int*>(d_small + 16)))
    return rc;
  translation_inliers_kernel<<<grid_for((unsigned long long)n), 256>>>(d_s, d_d, n, d_small + 9, noise, d_o);
  PSU_CHECK_LAUNCH("translation_inliers_kernel");
  PSU_CUDA(cudaMemcpy(t, d_small + 9, sizeof(double) * 3, cudaMemcpyDeviceToHost));
  if (inliers) PSU_CUDA(cudaMemcpy(inliers, d_o, (size_t)n, cudaMemcpyDeviceToHost));
  return PSULVSB_OK;
}

}  // extern "C"
