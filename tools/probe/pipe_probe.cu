// Micro-probe: issue cost (cycles per warp instruction per SM sub-partition) of the instructions in the attention
// softmax loop on sm_100a.  Diagnostics only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_ITERS 512
#define CHAINS 8

template <int OP>
__global__ void probe(float* out, long long* cyc, float seed) {
  float x[CHAINS];
  unsigned u[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { x[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; u[i] = __float_as_uint(x[i]); }
  float2 p[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) p[i] = make_float2(x[i], x[i] * 0.5f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < N_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(__uint_as_float(u[i])));
      if (OP == 2) asm volatile("{ .reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; fma.rn.f32x2 a, a, b, b; mov.b64 {%0, %1}, a; }" : "+f"(p[i].x), "+f"(p[i].y) : "f"(seed));
      if (OP == 3) asm volatile("{ .reg .b64 a, b; mov.b64 a, {%0, %1}; mov.b64 b, {%2, %2}; add.rn.f32x2 a, a, b; mov.b64 {%0, %1}, a; }" : "+f"(p[i].x), "+f"(p[i].y) : "f"(seed));
      if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(p[i].x), "f"(p[i].y));
      if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(0x8000u + it));
      if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(seed));
      if (OP == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(p[i].x));
      if (OP == 9) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 10) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 11) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 12) asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed));
      if (OP == 13) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
      if (OP == 14) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(it));
      if (OP == 15) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(it));
      if (OP == 16) asm volatile("max.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 17) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
      if (OP == 18) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(__uint_as_float(u[i])));
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc += x[i] + __uint_as_float(u[i]) + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 8);
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    int threads = 128 * warps_per_smsp;
    probe<OP><<<1, threads>>>(out, cyc, 0.5f);
    probe<OP><<<1, threads>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (N_ITERS * CHAINS);
    printf("%-22s warps/SMSP %d: %.2f cycles per warp-instruction per warp, %.2f per SMSP-issue\n", name, warps_per_smsp, per,
           per / warps_per_smsp);
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx (MUFU)");
  run<1>("cvt.rn.bf16x2.f32");
  run<8>("max.f32 (2 inputs)");
  run<9>("max.bf16x2");
  run<17>("max.f16x2");
  run<10>("add.rn.bf16x2");
  run<11>("fma.rn.bf16x2");
  run<12>("add.f32");
  run<13>("shl.b32");
  run<14>("lop3");
  run<15>("mad.lo.u32");
  run<16>("max.u32");
  run<18>("cvt.rn.f16x2.f32");
  run<2>("fma.rn.f32x2");
  run<3>("add.rn.f32x2");
  run<7>("fma.rn.f32");
  run<4>("max3.f32");
  run<5>("prmt");
  run<6>("add.u32");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
