// Probe: what does the legacy tensor path (mma.sync m16n8k16 bf16 -> HMMA) sustain on this GPU?
// Register-only operands, ILP independent accumulators per warp; sweep warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int ILP>
__global__ void hmma_kernel(float* out, int iters, uint32_t seed) {
  float c[ILP][4];
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 * 11u, b1 = a0 * 13u;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int ILP>
void run(int warps_per_cta, int ctas_per_sm, int sms) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 20000;
  dim3 grid(sms * ctas_per_sm), block(warps_per_cta * 32);
  hmma_kernel<ILP><<<grid, block>>>(out, 100, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  hmma_kernel<ILP><<<grid, block>>>(out, iters, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 16 * 8 * 16 * (double)ILP * iters * warps_per_cta * grid.x;
  printf("ILP %d  warps/CTA %2d  CTAs/SM %d  -> %.1f TFLOP/s  (%.3f ms)  %s\n", ILP, warps_per_cta, ctas_per_sm, flop / ms * 1e-9, ms,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  for (int w : {4, 8, 16, 32}) run<1>(w, 1, sms);
  for (int w : {4, 8, 16, 32}) run<4>(w, 1, sms);
  for (int w : {4, 8, 16}) run<8>(w, 1, sms);
  run<8>(16, 2, sms);
  return 0;
}
