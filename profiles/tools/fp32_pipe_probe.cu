// fp32_pipe_probe.cu -- what the FP32 pipe of one sm_100a SM really sustains, scalar (FFMA) against packed (FFMA2,
// PTX fma.rn.f32x2), so that the K1 / K4 roofline fractions in DESIGN.md can be read against a measured ceiling.
//
// Every thread keeps CH independent accumulator chains and runs ITERS rounds of
//   scalar : acc[c] = fma(acc[c], a, b)                       CH FFMA per round
//   packed : acc2[c] = fma2(acc2[c], {a, a}, {b0, b1})         CH FFMA2 per round (the {a, a} operand as K1 / K4 use it)
//   mixed  : P packed chains and S scalar chains side by side
// Reported: lane-FMAs per clock and SM (128 = the nominal FP32 rate), from clock64() of the slowest CTA.
//
// Build (profiles/tools/Makefile-less): nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/fp32_pipe_probe fp32_pipe_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

template <int P, int S, bool BROADCAST>
__global__ void __launch_bounds__(1024) probe(float* out, long long* cycles, int iters, const float* __restrict__ consts) {
  // operands from memory: they live in ordinary registers, as the tile values of K1 / K4 do (a kernel parameter would
  // sit in a uniform register)
  const float a = consts[threadIdx.x & 1], b0 = consts[2 + (threadIdx.x & 1)], b1 = consts[4 + (threadIdx.x & 1)];
  float2 p[P > 0 ? P : 1];
  float s[S > 0 ? S : 1];
#pragma unroll
  for (int c = 0; c < P; ++c) p[c] = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f - c);
#pragma unroll
  for (int c = 0; c < S; ++c) s[c] = threadIdx.x * 3e-3f + c;
  const float2 av = BROADCAST ? make_float2(a, a) : make_float2(a, a * 1.0000001f);
  const float2 bv = make_float2(b0, b1);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int c = 0; c < P; ++c) p[c] = fma2(p[c], av, bv);
#pragma unroll
      for (int c = 0; c < S; ++c) s[c] = fmaf(s[c], a, b0);
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < P; ++c) acc += p[c].x + p[c].y;
#pragma unroll
  for (int c = 0; c < S; ++c) acc += s[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static float* d_consts = nullptr;
template <int P, int S, bool BC>
static void run(const char* name, int sms, int threads, int ctas_per_sm, float* d_out, long long* d_cyc) {
  const int iters = 4096;
  const int grid = sms * ctas_per_sm;
  for (int w = 0; w < 2; ++w) probe<P, S, BC><<<grid, threads>>>(d_out, d_cyc, iters, d_consts);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<P, S, BC><<<grid, threads>>>(d_out, d_cyc, iters, d_consts);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("%s: launch failed\n", name);
    return;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long c : cyc) mx = c > mx ? c : mx;
  const double lane_fma_per_cta = (double)threads * iters * 4.0 * (2.0 * P + S);
  const double per_clk_sm = lane_fma_per_cta * ctas_per_sm / (double)mx;
  const double issue_per_clk_smsp = (double)threads / 32.0 * iters * 4.0 * (P + S) * ctas_per_sm / 4.0 / (double)mx;
  printf("{\"case\": \"%s\", \"packed_chains\": %d, \"scalar_chains\": %d, \"threads\": %d, \"ctas_per_sm\": %d, "
         "\"lane_fma_per_clk_sm\": %.2f, \"frac_of_128\": %.4f, \"issue_per_clk_smsp\": %.3f, \"ms\": %.3f, "
         "\"mhz_effective\": %.0f}\n",
         name, P, S, threads, ctas_per_sm, per_clk_sm, per_clk_sm / 128.0, issue_per_clk_smsp, ms, mx / (ms * 1e3));
}


// K4-shaped operand traffic: NM "matrix" register pairs (loop invariant, like s*R of two hypotheses), NP "points" (scalar
// operands), one accumulator pair per (m, p).  ORDER 0 = the scalar operand is shared by consecutive instructions (what
// ptxas makes of k4's source: R.reuse.F32), ORDER 1 = the 64-bit matrix operand is shared by NP consecutive
// instructions.  Same arithmetic, different register-file traffic.
template <int NM, int NP, int ORDER, int NA, int KIND>
__global__ void __launch_bounds__(256) probe_k4(float* out, long long* cycles, int iters, const float* __restrict__ consts) {
  float2 M[NM], acc[NM][NP];
  float s[NP];
  int cnt = 0;
  float mnv[4] = {3e38f, 3e38f, 3e38f, 3e38f};
#pragma unroll
  for (int m = 0; m < NM; ++m) M[m] = make_float2(consts[(threadIdx.x + m) & 1] + m * 1e-4f, consts[(threadIdx.x + m + 1) & 1] - m * 1e-4f);
#pragma unroll
  for (int p = 0; p < NP; ++p) s[p] = consts[2 + ((threadIdx.x + p) & 1)] + p;
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[m][p] = make_float2(threadIdx.x * 1e-3f + m, p);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      if (ORDER == 0) {
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
          for (int m = 0; m < NM; ++m) acc[m][p] = fma2(M[m], make_float2(s[p], s[p]), acc[m][p]);
      } else {
#pragma unroll
        for (int m = 0; m < NM; ++m)
#pragma unroll
          for (int p = 0; p < NP; ++p) acc[m][p] = fma2(M[m], make_float2(s[p], s[p]), acc[m][p]);
      }
      // ALU-pipe instructions on fresh results.  KIND 0: NA x (LEA.HI count + FMNMX band tracking), as K4 has them;
      // 1: NA x LEA.HI; 2: NA x FMNMX (two-input); 3: NA x FMNMX3 (three-input: two new values per instruction);
      // 4: NA x SHF (sign bit funnelled into a mask word, as K1 packs its mask)
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const float v = (a & 1) ? acc[a % NM][a % NP].y : acc[a % NM][a % NP].x;
        const float w = (a & 1) ? acc[(a + 1) % NM][a % NP].x : acc[(a + 1) % NM][a % NP].y;
        if (KIND == 0 || KIND == 1) cnt += (int)(__float_as_uint(v) >> 31);
        if (KIND == 0 || KIND == 2) mnv[a & 3] = fminf(mnv[a & 3], fabsf(v));
        if (KIND == 3) mnv[a & 3] = fminf(mnv[a & 3], fminf(fabsf(v), fabsf(w)));
        if (KIND == 4) cnt = (int)__funnelshift_l(__float_as_uint(v), (unsigned)cnt, 1);
      }
    }
  }
  const long long t1 = clock64();
  float a = mnv[0] + mnv[1] + mnv[2] + mnv[3] + (float)cnt;
#pragma unroll
  for (int m = 0; m < NM; ++m)
#pragma unroll
    for (int p = 0; p < NP; ++p) a += acc[m][p].x + acc[m][p].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NM, int NP, int ORDER, int NA = 0, int KIND = 0>
static void run_k4(const char* name, int sms, int ctas_per_sm, float* d_out, long long* d_cyc) {
  const int iters = 2048, threads = 256;
  const int grid = sms * ctas_per_sm;
  for (int w = 0; w < 2; ++w) probe_k4<NM, NP, ORDER, NA, KIND><<<grid, threads>>>(d_out, d_cyc, iters, d_consts);
  probe_k4<NM, NP, ORDER, NA, KIND><<<grid, threads>>>(d_out, d_cyc, iters, d_consts);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("%s: launch failed\n", name);
    return;
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long c : cyc) mx = c > mx ? c : mx;
  const double per_clk_sm = (double)threads * iters * 2.0 * (2.0 * NM * NP) * ctas_per_sm / (double)mx;
  const int alu_instr = (KIND == 0 ? 2 : 1) * NA;
  const double cyc_per_ffma2 = (double)mx / ((double)threads / 32.0 / 4.0 * ctas_per_sm * iters * 2.0 * NM * NP);
  printf("{\"case\": \"%s\", \"matrix_pairs\": %d, \"points\": %d, \"order\": %d, \"alu_instr_per_ffma2\": %.3f, \"alu_kind\": %d, \"threads\": %d, "
         "\"ctas_per_sm\": %d, \"lane_fma_per_clk_sm\": %.2f, \"frac_of_128\": %.4f, \"smsp_cycles_per_ffma2\": %.3f}\n",
         name, NM, NP, ORDER, (double)alu_instr / (NM * NP), KIND, threads, ctas_per_sm, per_clk_sm, per_clk_sm / 128.0, cyc_per_ffma2);
}

// ---- the word loop of K1 (k1_consistency.cu eval_word), alone: R rows per thread as scalars against the 32
// pair-interleaved columns of a word, 8 words of a 256-column tile, over and over.  What the loop costs per pair with
// nothing else around it (no epilogue, no tile ring, no diagonal): the ceiling of the kernel's fast path.
__device__ __forceinline__ float2 bc2(const float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 k1_pair_fast2(const float4 ms, const float4 mt, const float4 sa, const float4 sb,
                                                const float4 ta, const float4 tb, const float n4b2) {
  const float2 A = fma2(bc2(ms.x), make_float2(sa.x, sa.y),
                        fma2(bc2(ms.y), make_float2(sa.z, sa.w), fma2(bc2(ms.z), make_float2(sb.x, sb.y), make_float2(sb.z, sb.w))));
  const float2 B = fma2(bc2(mt.x), make_float2(ta.x, ta.y),
                        fma2(bc2(mt.y), make_float2(ta.z, ta.w), fma2(bc2(mt.z), make_float2(tb.x, tb.y), make_float2(tb.z, tb.w))));
  const float2 u = fma2(fma2(B, bc2(-1.f), A), bc2(1.f), bc2(ms.w));
  const float2 w = fma2(bc2(n4b2), B, bc2(mt.w));
  return fma2(u, u, w);
}
template <int R, int UNROLL, int ALU>
__global__ void __launch_bounds__(256) probe_k1(float* out, long long* cycles, int iters, const float* __restrict__ consts) {
  __shared__ __align__(16) float4 cs[256], ct[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) {
    cs[k] = make_float4(k * 1e-3f, k * 2e-3f, 1.f - k * 1e-3f, 0.5f + k * 1e-4f);
    ct[k] = make_float4(k * 1.5e-3f, 0.3f - k * 2e-3f, k * 1e-3f, 0.25f + k * 1e-4f);
  }
  float4 ms[R], mt[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float a = consts[(threadIdx.x + r) & 1] * (1.f + r) + threadIdx.x * 1e-3f;
    ms[r] = make_float4(-2.f * a, a, 0.5f * a, a * a);
    mt[r] = make_float4(a, -2.f * a, 0.25f * a, -a * a);
  }
  const float n4b2 = consts[2] * -0.04f;
  uint32_t acc[R];
  float mv[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    acc[r] = 0u;
    mv[r] = 3e38f;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int wj = 0; wj < 8; ++wj) {
      const float4* cw = cs + wj * 32;
      const float4* tw = ct + wj * 32;
#pragma unroll UNROLL
      for (int q = 15; q >= 0; --q) {
        const float4 sa = cw[2 * q], sb = cw[2 * q + 1];
        const float4 ta = tw[2 * q], tb = tw[2 * q + 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 v = k1_pair_fast2(ms[r], mt[r], sa, sb, ta, tb, n4b2);
          if (ALU >= 1) {
            acc[r] = __funnelshift_l(__float_as_uint(v.y), acc[r], 1);
            acc[r] = __funnelshift_l(__float_as_uint(v.x), acc[r], 1);
          } else {
            acc[r] ^= __float_as_uint(v.x + v.y) * (q == 0 && wj == 0 && it == 0);
          }
          if (ALU >= 2) mv[r] = fminf(mv[r], fminf(fabsf(v.x), fabsf(v.y)));
        }
      }
    }
  }
  const long long t1 = clock64();
  float a = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r) a += mv[r] + (float)acc[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int R, int UNROLL, int ALU>
static void run_k1(const char* name, int sms, int ctas_per_sm, float* d_out, long long* d_cyc) {
  const int iters = 64, threads = 256;
  const int grid = sms * ctas_per_sm;
  for (int w = 0; w < 2; ++w) probe_k1<R, UNROLL, ALU><<<grid, threads>>>(d_out, d_cyc, iters, d_consts);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError()));
    return;
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (long long c : cyc) mx = c > mx ? c : mx;
  // pairs per lane and SMSP: (warps per SMSP) x iters x 256 columns x R rows
  const double pairs_per_lane_smsp = (double)threads / 32.0 / 4.0 * ctas_per_sm * iters * 256.0 * R;
  const double cyc_per_pair = (double)mx / pairs_per_lane_smsp;
  printf("{\"case\": \"%s\", \"rows_per_thread\": %d, \"colpair_unroll\": %d, \"alu\": %d, \"ctas_per_sm\": %d, "
         "\"smsp_cycles_per_pair\": %.3f, \"frac_of_16_slot_roofline\": %.4f}\n",
         name, R, UNROLL, ALU, ctas_per_sm, cyc_per_pair, 16.0 / cyc_per_pair);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, sizeof(float) * sms * 4 * 1024);
  cudaMalloc(&d_cyc, sizeof(long long) * sms * 4);
  const float hc[6] = {0.999f, 0.999f, 1e-3f, 1e-3f, 2e-3f, 2e-3f};
  cudaMalloc(&d_consts, sizeof(hc));
  cudaMemcpy(d_consts, hc, sizeof(hc), cudaMemcpyHostToDevice);
  for (int threads : {256, 512, 1024}) {
    run<0, 8, true>("scalar FFMA, 8 chains", sms, threads, 1, d_out, d_cyc);
    run<0, 16, true>("scalar FFMA, 16 chains", sms, threads, 1, d_out, d_cyc);
    run<8, 0, true>("packed FFMA2 {a,a}, 8 chains", sms, threads, 1, d_out, d_cyc);
    run<16, 0, true>("packed FFMA2 {a,a}, 16 chains", sms, threads, 1, d_out, d_cyc);
    run<8, 0, false>("packed FFMA2 {a,a'}, 8 chains", sms, threads, 1, d_out, d_cyc);
    run<8, 2, true>("mixed 8 FFMA2 : 2 FFMA", sms, threads, 1, d_out, d_cyc);
    run<8, 4, true>("mixed 8 FFMA2 : 4 FFMA", sms, threads, 1, d_out, d_cyc);
    run<8, 8, true>("mixed 8 FFMA2 : 8 FFMA", sms, threads, 1, d_out, d_cyc);
    run<4, 8, true>("mixed 4 FFMA2 : 8 FFMA", sms, threads, 1, d_out, d_cyc);
  }
  for (int cps : {1, 2}) {
    run_k4<12, 2, 0>("k4-shaped, scalar operand shared (12 x 2)", sms, cps, d_out, d_cyc);
    run_k4<12, 2, 1>("k4-shaped, matrix operand shared (12 x 2)", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0>("k4-shaped, scalar operand shared (6 x 4)", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 1>("k4-shaped, matrix operand shared (6 x 4)", sms, cps, d_out, d_cyc);
    run_k4<9, 4, 1>("k4-shaped, matrix operand shared (9 x 4)", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 3, 0>("scalar shared + 3 x (LEA.HI + FMNMX): 1 ALU per 4 FFMA2 (K4: 1 per 3.75)", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 1, 3, 0>("matrix shared + 3 x (LEA.HI + FMNMX)", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 6, 1>("scalar shared + 6 LEA.HI", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 6, 2>("scalar shared + 6 FMNMX", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 6, 3>("scalar shared + 6 FMNMX3", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 6, 4>("scalar shared + 6 SHF", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 12, 1>("scalar shared + 12 LEA.HI", sms, cps, d_out, d_cyc);
    run_k4<6, 4, 0, 12, 3>("scalar shared + 12 FMNMX3", sms, cps, d_out, d_cyc);
  }
  for (int cps : {2}) {
    run_k1<4, 4, 1>("K1 word loop, + SHF", sms, cps, d_out, d_cyc);
    run_k1<4, 4, 2>("K1 word loop, + SHF + FMNMX3 (the kernel's loop)", sms, cps, d_out, d_cyc);
    run_k1<4, 2, 2>("K1 word loop, unroll 2", sms, cps, d_out, d_cyc);
    run_k1<4, 8, 2>("K1 word loop, unroll 8", sms, cps, d_out, d_cyc);
    run_k1<2, 4, 2>("K1 word loop, 2 rows", sms, cps, d_out, d_cyc);
    run_k1<6, 4, 2>("K1 word loop, 6 rows", sms, cps, d_out, d_cyc);
  }
  return 0;
}
