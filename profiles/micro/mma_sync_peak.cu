// Peak rate of legacy mma.sync.m16n8k16 (fp16 in, fp32 accumulate) on this GPU: W warps per SM, A accumulators per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_peak mma_sync_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int ACC>
__global__ void k(float* out, int iters) {
  float c[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.f) out[0] = s;
}
template <int ACC>
void run(int warps, int ctas_per_sm) {
  int iters = 20000;
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ACC><<<148 * ctas_per_sm, warps * 32>>>(d, 100);
  cudaEventRecord(e0);
  k<ACC><<<148 * ctas_per_sm, warps * 32>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 16 * 8 * 16 * (double)ACC * iters * warps * ctas_per_sm * 148;
  printf("acc=%d warps/CTA=%d CTAs/SM=%d: %.3f ms, %.1f TFLOP/s, %.0f MAC/clk/SM at 1.9 GHz-equivalent (%.0f flop/ns/SM)\n", ACC, warps, ctas_per_sm, ms,
         flops / ms / 1e9, flops / 2 / (ms * 1e-3) / 148 / 1.9e9, flops / (ms * 1e6) / 148);
}
int main() {
  run<1>(4, 1); run<2>(4, 1); run<4>(4, 1); run<8>(4, 1); run<8>(8, 1); run<8>(4, 2); run<16>(4, 2); run<8>(8, 2); run<8>(16, 1); run<4>(16, 2);
  return 0;
}
