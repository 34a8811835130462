// Micro-benchmark: cost of back-to-back tcgen05.mma (M=128, K=16, bf16) as a function of N and of accumulator
// dependence. One CTA per SM, one issuing thread, operands = whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I yolo_sam_inference_b200/csrc scripts/micro/mma_issue.cu -o gpurun_out/mma_issue
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ysi;

template <int N, int NACC, bool TS = false>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tptr;
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t a = umma_desc_sw128(sbase, 16, 1024), b = umma_desc_sw128(sbase + 16384, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (TS) umma_bf16_ts(tm + ((i * 8 + j) % NACC) * 128, tm + 384 + 8 * (j & 3), b + 2u * (j & 3), idesc, 1);
        else umma_bf16_ss(tm + ((i * 8 + j) % NACC) * (N < 128 ? 128 : N), a + 2u * (j & 3), b + 2u * (j & 3), idesc, 1);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int NACC, bool TS = false>
void run(const char* tag, long long* d) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<N, NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, NACC, TS><<<148, 128, 100 * 1024>>>(d, iters);
  cudaDeviceSynchronize();
  k<N, NACC, TS><<<148, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d acc=%d  %7.1f clk/MMA  (math floor %d)  %s\n", tag, N, NACC, double(h) / (iters * 8.0), N / 2, cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  run<64, 1>("dependent", d);   run<64, 2>("2 accumulators", d);   run<64, 4>("4 accumulators", d);
  run<80, 1>("dependent", d);   run<80, 2>("2 accumulators", d);
  run<128, 1>("dependent", d);  run<128, 2>("2 accumulators", d);
  run<256, 1>("dependent", d);  run<256, 2>("2 accumulators", d);
  run<64, 1, true>("A in TMEM, dependent", d); run<64, 2, true>("A in TMEM, 2 acc", d);
  run<80, 1, true>("A in TMEM, dependent", d); run<128, 1, true>("A in TMEM, dependent", d);
  run<16, 1>("dependent", d);   run<16, 4>("4 accumulators", d);
  return 0;
}
