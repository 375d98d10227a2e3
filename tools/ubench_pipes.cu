// ubench_pipes.cu — issue-rate microbenchmark of the sm_100a pipes the march kernels live on:
// FFMA (3-reg), packed FFMA2/FMUL2/FADD2 (fma.rn.f32x2), FMNMX, MUFU.RCP and mixes of them.
// Prints warp-instructions per clock per SM and lane-flops per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ubench_pipes ubench_pipes.cu
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 4096
#define NCH 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float x[NCH];
  float2 y[NCH];
  for (int i = 0; i < NCH; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = make_float2(x[i], x[i] + 0.5f); }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) x[i] = __fmaf_rn(x[i], a, b);
      if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);
      if (MODE == 2) x[i] = fminf(fmaxf(x[i], a), b);                      // 2 FMNMX
      if (MODE == 3) { x[i] = fminf(fmaxf(x[i], a), b); x[i] = __fmaf_rn(x[i], a, b); }  // 2 FMNMX + 1 FFMA
      if (MODE == 4) y[i] = __fmul2_rn(y[i], a2);
      if (MODE == 5) y[i] = __fadd2_rn(y[i], b2);
      if (MODE == 6) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = fminf(fmaxf(x[i], a), b); }  // 1 FFMA2 + 2 FMNMX
      if (MODE == 7) x[i] = __fmul_rn(x[i], a);
      if (MODE == 8) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = __fmaf_rn(x[i], a, b); }  // FFMA2 + FFMA
      if (MODE == 9) { x[i] = __fmaf_rn(x[i], a, b); x[i] = fmaxf(x[i], a); }  // 1 FFMA + 1 FMNMX
      if (MODE == 10) { y[i] = __ffma2_rn(y[i], a2, b2); y[i] = __ffma2_rn(y[i], a2, b2); x[i] = fmaxf(x[i], a); }  // 2 FFMA2 + 1 FMNMX
      if (MODE == 11) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      if (MODE == 12) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = (x[i] > a) ? x[i] : b; }  // FFMA2 + FSETP/FSEL
    }
  }
  float s = 0;
  for (int i = 0; i < NCH; ++i) s += x[i] + y[i].x + y[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int instr_per_chain, double flops_per_chain, float* d, int nsm, double clk_hz) {
  const int blocks = nsm * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(d, 0.999f, 1e-3f, 64);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, 0.999f, 1e-3f, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double warps = (double)blocks * threads / 32, winstr = warps * ITERS * NCH * instr_per_chain;
  const double clocks = best * 1e-3 * clk_hz;
  printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM  %7.1f lane-flop/clk/SM\n", name, best, winstr / clocks / nsm,
         warps * 32 * ITERS * NCH * flops_per_chain / clocks / nsm);
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk = clk_khz * 1e3;
  printf("%s: %d SMs, clock %.0f MHz (rates assume the max clock)\n", p.name, p.multiProcessorCount, clk / 1e6);
  float* d;
  cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  const int n = p.multiProcessorCount;
  run<0>("FFMA", 1, 2, d, n, clk);
  run<7>("FMUL", 1, 1, d, n, clk);
  run<1>("FFMA2", 1, 4, d, n, clk);
  run<4>("FMUL2", 1, 2, d, n, clk);
  run<5>("FADD2", 1, 2, d, n, clk);
  run<2>("FMNMX x2", 2, 2, d, n, clk);
  run<3>("2 FMNMX + 1 FFMA", 3, 4, d, n, clk);
  run<9>("1 FFMA + 1 FMNMX", 2, 3, d, n, clk);
  run<6>("1 FFMA2 + 2 FMNMX", 3, 6, d, n, clk);
  run<8>("1 FFMA2 + 1 FFMA", 2, 6, d, n, clk);
  run<10>("2 FFMA2 + 1 FMNMX", 3, 9, d, n, clk);
  run<12>("1 FFMA2 + FSETP+FSEL", 3, 5, d, n, clk);
  run<11>("MUFU.RCP", 1, 1, d, n, clk);
  return 0;
}
