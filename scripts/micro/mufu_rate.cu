// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition as a function of resident warps, alone and mixed with
// FADD2 / F2FP work (the softmax inner loop of csrc/attn.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/micro/mufu_rate.cu -o scripts/micro/mufu_rate.bin
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -1.0f - 0.01f * (threadIdx.x + i);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float e = ex2(v[i]);
      if (MODE >= 1) acc += e;                 // one FADD per exp
      if (MODE >= 2) v[i] = v[i] + 1e-3f * acc; // + dependent FFMA
      else v[i] = e - 1.5f;                    // keep the chain per element (FADD)
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(int warps, float* o, long long* c) {
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(o, c, iters);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(o, c, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  const double per_smsp = double(iters) * 32 * warps / 4.0;   // MUFU warp-instructions per sub-partition
  printf("mode %d  warps/SM %2d : %6.2f cycles per MUFU warp-instr per SMSP  (%s)\n", MODE, warps, double(h) / per_smsp, cudaGetErrorString(e));
}

int main() {
  float* o; long long* c;
  cudaMalloc(&o, 4); cudaMalloc(&c, 8);
  for (int w : {4, 8, 16, 32}) run<0>(w, o, c);
  for (int w : {4, 8, 16, 32}) run<1>(w, o, c);
  for (int w : {4, 8, 16, 32}) run<2>(w, o, c);
  return 0;
}
