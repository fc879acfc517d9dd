// Micro-probe: cycles per tcgen05.mma (bf16, M=128, N=128, K=16) for SS / TS operand modes and K-major / MN-major B,
// issued back to back by one thread into one or two TMEM accumulators.  Diagnostics only (operands are garbage).
#include <cstdio>
#include "../../ltx-2-mlx_b200/csrc/common.cuh"
using namespace ltx2;

template <int MODE>   // 0: SS K-major B, 1: TS K-major B, 2: TS MN-major B, 3: SS N=256, 4: alternating SS(S) / TS(PV)
__global__ void __launch_bounds__(128, 1) mma_probe(long long* cyc, int n_groups) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t ad = umma_desc_k_sw128(smem_u32(smem));
    const uint64_t bd = umma_desc_k_sw128(smem_u32(smem + 32768));
    const uint64_t bmn = umma_desc_mn_sw128(smem_u32(smem + 32768), 128 * 128, 1024);
    constexpr uint32_t id128 = umma_idesc_bf16(128, 128), id128mn = umma_idesc_bf16(128, 128, true),
                       id256 = umma_idesc_bf16(128, 256);
    long long t0 = 0;
    uint32_t ph = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int g = 0; g < n_groups; ++g) {
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t off = ((ks / 4) * (128 * 128) >> 4) + 2 * (ks % 4);
            if (MODE == 0) umma_bf16_ss(tm + (g & 1) * 128, ad + off, bd + off, id128, ks != 0);
            if (MODE == 1) umma_bf16_ts(tm + (g & 1) * 128, tm + 256 + ks * 8, bd + off, id128, ks != 0);
            if (MODE == 2) umma_bf16_ts(tm + (g & 1) * 128, tm + 256 + ks * 8, bmn + ks * (2048 >> 4), id128mn, ks != 0);
            if (MODE == 3) umma_bf16_ss(tm + (g & 1) * 256, ad + off, bd + off, id256, ks != 0);
            if (MODE >= 16) umma_bf16_ss(tm + (g & 1) * 256, ad + off, bd + off, umma_idesc_bf16(128, MODE), ks != 0);
            if (MODE == 5) umma_bf16_ss(tm + (g & 3) * 64, ad + off, bd + off, umma_idesc_bf16(128, 64), ks != 0);
            if (MODE == 6) {   // sub-block step: P*V with K = 64 (4 TS MMAs, N = 128), then S with N = 64 (8 SS MMAs)
              if (ks < 4) umma_bf16_ts(tm + 256 + (g & 1) * 128, tm + (g & 3) * 64 + ks * 8, bmn + ks * (2048 >> 4), id128mn, 1);
            }
            if (MODE == 4) {
              if (g & 1) umma_bf16_ts(tm + 256 + (g & 2) * 64, tm + (g & 2) * 64 + ks * 8, bmn + ks * (2048 >> 4), id128mn, 1);
              else umma_bf16_ss(tm + (g & 2) * 64, ad + off, bd + off, id128, ks != 0);
            }
          }
        }
        if (MODE == 6 && leader) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t off = ((ks / 4) * (128 * 128) >> 4) + 2 * (ks % 4);
            umma_bf16_ss(tm + (g & 3) * 64, ad + off, bd + off + (8192 >> 4), umma_idesc_bf16(128, 64), ks != 0);
          }
        }
        __syncwarp();
      }
      if (leader) umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    if (leader) *cyc = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int MODE>
void run(const char* name, int sms) {
  long long* cyc;
  cudaMalloc(&cyc, 8);
  const int n = 256;
  cudaFuncSetAttribute(mma_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  mma_probe<MODE><<<sms, 128, 100 * 1024>>>(cyc, n);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s CTAs %3d: %.1f cycles per 8-MMA group (K=128), %.1f per MMA   [%s]\n", name, sms, (double)h / n,
         (double)h / n / 8, cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  for (int sms : {1, 148}) {
    run<0>("SS  A,B K-major smem, N=128", sms);
    run<1>("TS  A tmem, B K-major smem, N=128", sms);
    run<2>("TS  A tmem, B MN-major smem, N=128", sms);
    run<3>("SS  N=256", sms);
    run<4>("alternating SS (QK^T) / TS MN-major (PV)", sms);
    run<5>("SS  N=64", sms);
    run<6>("sub-block step: 4x TS N=128 + 8x SS N=64", sms);
    run<80>("SS  N=80", sms);
    run<112>("SS  N=112", sms);
    run<144>("SS  N=144", sms);
    run<176>("SS  N=176", sms);
    run<192>("SS  N=192", sms);
    run<208>("SS  N=208", sms);
    run<224>("SS  N=224", sms);
    run<240>("SS  N=240", sms);
  }
  return 0;
}
