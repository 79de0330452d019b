// Micro-benchmark: issue rate of the instructions the activation role is made of
// (one warp per SM sub-partition, 8 independent dependency chains, clock64).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu
#include <cuda_runtime.h>
#include <cstdio>
#define REP 64
#define CH 8
template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float2 a[CH];
  for (int i = 0; i < CH; ++i) a[i] = make_float2(seed + i, seed * 0.5f + i);
  const float2 m = make_float2(seed * 0.999f, seed * 1.001f);
  const float2 c = make_float2(0.25f, 0.125f);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int r = 0; r < REP / CH; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if (OP == 0) a[i].x = fmaf(a[i].x, m.x, c.x);                       // FFMA reg
        if (OP == 1) a[i].x = fmaf(a[i].x, m.x, 0.25f);                     // FFMA imm
        if (OP == 2) a[i] = __ffma2_rn(a[i], m, c);                         // FFMA2 reg
        if (OP == 3) a[i] = __ffma2_rn(a[i], m, make_float2(0.25f, 0.25f)); // FFMA2 imm
        if (OP == 4) a[i] = __fadd2_rn(a[i], m);                            // FADD2
        if (OP == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x));
        if (OP == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x));
        if (OP == 7) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i].x));
        if (OP == 8) a[i].x = fmaxf(a[i].x, m.x);                           // FMNMX
        if (OP == 9) a[i].x = __uint_as_float(__float_as_uint(a[i].x) & 0xFFFFE000u | 1u);  // LOP3
        if (OP == 10) { a[i].x = fmaf(a[i].x, m.x, 0.25f); a[i].y = fmaf(a[i].y, m.y, 0.25f); }  // 2x FFMA imm
        if (OP == 11) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i].x));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"FFMA reg", "FFMA imm", "FFMA2 reg", "FFMA2 imm", "FADD2", "MUFU.EX2", "MUFU.LG2",
                         "MUFU.RCP", "FMNMX", "LOP3", "2xFFMA imm", "MUFU.TANH"};
  for (int warps = 1; warps <= 2; ++warps)
    for (int op = 0; op < 12; ++op) {
      long long h = 0;
#define RUN(OP) case OP: k<OP><<<148, 128 * warps>>>(out, cyc, 1.0001f); break;
      switch (op) { RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) }
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double n = 256.0 * REP * (op == 10 ? 2 : 1);
      printf("%d warp(s)/SMSP  %-10s  %.2f cycles per warp-instruction\n", warps, names[op], h / n);
    }
  return 0;
}
