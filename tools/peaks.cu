// peaks.cu -- measured FP64 / MUFU / DMMA issue peaks of the B200 the benchmarks run on, plus accuracy
// of the kx_math.cuh primitives.  Output: one JSON object on stdout (stored as profiles/peaks_rNN.json);
// these are the denominators of the FP64-pipe roofline in bench.py / DESIGN.md (SURVEY.md 8d asks for
// them to be measured, not assumed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I kinetix_b200/csrc -o /tmp/peaks tools/peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "kx_math.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b)
{
  double c0[4], c1[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// DFMA and DMMA interleaved: do the two pipes overlap?
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b)
{
  double c0[4], c1[4], x[8];
#pragma unroll
  for (int i = 0; i < 4; i++) { c0[i] = threadIdx.x; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
      x[2 * i] = fma(x[2 * i], a, b);
      x[2 * i + 1] = fma(x[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mufu(float* out, int iters, float a)
{
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  if (s == 12345.678f) out[0] = s + a;
}

__global__ void __launch_bounds__(256) k_rcp64h(double* out, int iters, double a)
{
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + 1.5 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[i]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i];
  if (s == 12345.678) out[0] = s + a;
}

// accuracy kernels
__global__ void k_acc(const double* x, double* seed, double* r3, double* r5, double* ex, double* lg, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a = x[i], s;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(a));
  seed[i] = s;
  double e = fma(-a, s, 1.0);
  double u = fma(e, e, e);
  r3[i] = fma(s, u, s);
  r5[i] = kx_rcp(a);
  ex[i] = kx_exp(a);
  lg[i] = kx_log(fabs(a) + 1e-300);
}

template <class F>
double time_ms(F launch, int reps = 5)
{
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  return best;
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* d; CK(cudaMalloc(&d, 1024));
  const int grid = sms * 8, block = 256, iters = 20000;
  const double lanes = (double)grid * block;
  printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);

  double ms;
  ms = time_ms([&] { k_dfma<8><<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_ilp8\": %.4e,\n", lanes * iters * 8 / (ms * 1e-3));
  ms = time_ms([&] { k_dfma<4><<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_ilp4\": %.4e,\n", lanes * iters * 4 / (ms * 1e-3));
  ms = time_ms([&] { k_dfma<2><<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_ilp2\": %.4e,\n", lanes * iters * 2 / (ms * 1e-3));
  ms = time_ms([&] { k_dfma<1><<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_ilp1\": %.4e,\n", lanes * iters * 1 / (ms * 1e-3));
  // low occupancy: 8 warps per SM (what a 255-register kernel gets)
  ms = time_ms([&] { k_dfma<8><<<sms, 256>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_8warps_ilp8\": %.4e,\n", (double)sms * 256 * iters * 8 / (ms * 1e-3));
  ms = time_ms([&] { k_dfma<2><<<sms, 256>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_8warps_ilp2\": %.4e,\n", (double)sms * 256 * iters * 2 / (ms * 1e-3));
  ms = time_ms([&] { k_dfma<8><<<sms, 128>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"dfma_lane_instr_per_s_4warps_ilp8\": %.4e,\n", (double)sms * 128 * iters * 8 / (ms * 1e-3));

  ms = time_ms([&] { k_dmma<<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  // one m8n8k4 warp instruction = 256 FMA
  printf(" \"dmma_fma_per_s\": %.4e,\n", (double)grid * (block / 32) * iters * 4 * 256.0 / (ms * 1e-3));
  double ms_mixed = time_ms([&] { k_mixed<<<grid, block>>>(d, iters, 1.0000001, 1e-9); });
  printf(" \"mixed_dmma_fma_per_s\": %.4e, \"mixed_dfma_lane_instr_per_s\": %.4e,\n",
         (double)grid * (block / 32) * iters * 4 * 256.0 / (ms_mixed * 1e-3), lanes * iters * 8 / (ms_mixed * 1e-3));

  ms = time_ms([&] { k_mufu<<<grid, block>>>((float*)d, iters, 1.f); });
  printf(" \"mufu_ex2_lane_instr_per_s\": %.4e,\n", lanes * iters * 8 / (ms * 1e-3));
  ms = time_ms([&] { k_rcp64h<<<grid, block>>>(d, iters, 1.0); });
  printf(" \"mufu_rcp64h_lane_instr_per_s\": %.4e,\n", lanes * iters * 8 / (ms * 1e-3));

  // sustained DFMA over ~3 s (power-capped clock)
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int launches = 0;
    float total = 0;
    while (total < 3000.f) {
      for (int i = 0; i < 10; i++) k_dfma<8><<<grid, block>>>(d, iters, 1.0000001, 1e-9);
      launches += 10;
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&total, e0, e1));
    }
    printf(" \"dfma_lane_instr_per_s_sustained_3s\": %.4e,\n", lanes * iters * 8 * launches / (total * 1e-3));
  }

  // accuracy
  const int n = 1 << 20;
  std::vector<double> hx(n), hs(n), h3(n), h5(n), he(n), hl(n);
  srand(7);
  for (int i = 0; i < n; i++) {
    double u = rand() / (double)RAND_MAX;
    hx[i] = (i & 1) ? (u * 1400.0 - 700.0) : exp((u - 0.5) * 200.0) * ((i & 2) ? 1 : -1);
  }
  double *dx, *ds, *d3, *d5, *de, *dl;
  CK(cudaMalloc(&dx, n * 8)); CK(cudaMalloc(&ds, n * 8)); CK(cudaMalloc(&d3, n * 8)); CK(cudaMalloc(&d5, n * 8));
  CK(cudaMalloc(&de, n * 8)); CK(cudaMalloc(&dl, n * 8));
  CK(cudaMemcpy(dx, hx.data(), n * 8, cudaMemcpyHostToDevice));
  k_acc<<<n / 256, 256>>>(dx, ds, d3, d5, de, dl, n);
  CK(cudaMemcpy(hs.data(), ds, n * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h3.data(), d3, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h5.data(), d5, n * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(he.data(), de, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hl.data(), dl, n * 8, cudaMemcpyDeviceToHost));
  double es = 0, e3 = 0, e5 = 0, ee = 0, el = 0;
  for (int i = 0; i < n; i++) {
    long double x = hx[i], inv = 1.0L / x;
    es = fmax(es, (double)fabsl((hs[i] - inv) / inv));
    e3 = fmax(e3, (double)fabsl((h3[i] - inv) / inv));
    e5 = fmax(e5, (double)fabsl((h5[i] - inv) / inv));
    if (fabs(hx[i]) < 700) { long double t = expl(x); ee = fmax(ee, (double)fabsl((he[i] - t) / t)); }
    long double lt = logl(fabsl(x) + 1e-300L);
    if (fabsl(lt) > 1e-3L) el = fmax(el, (double)fabsl((hl[i] - lt) / lt));
  }
  printf(" \"rcp64h_seed_max_rel_err\": %.3e, \"rcp_3dfma_max_rel_err\": %.3e, \"kx_rcp_max_rel_err\": %.3e,\n", es, e3, e5);
  printf(" \"kx_exp_max_rel_err\": %.3e, \"kx_log_max_rel_err\": %.3e\n}\n", ee, el);
  return 0;
}
