// Micro-benchmark: issue rate of warp-level mma.sync.m16n8k16 (bf16 x bf16 + fp32) on sm_100a -- the "legacy" tensor path that the fused
// FFN kernel uses for the depthwise 5x5 (block-diagonal taps).  Reports cycles per MMA per SM sub-partition and the dense-equivalent TFLOP/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

__device__ __forceinline__ void mma(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void k(float* out, const uint32_t* in, long long* cyc) {
  float acc[NACC][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = in[threadIdx.x % 32 + 32 * i];
  for (int i = 0; i < 2; ++i) b[i] = in[threadIdx.x % 32 + 32 * (4 + i)];
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma(acc[j], a[0], a[1], a[2], a[3], b[0], b[1]);
  }
  long long t1 = clock64();
  float r = 0.f;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) r += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int NACC>
void run(float* out, uint32_t* in, long long* cyc, double ghz) {
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    k<NACC><<<148, threads>>>(out, in, cyc);
    k<NACC><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double mma_per_smsp = (double)ITER * NACC * warps_per_smsp;
    const double cyc_per_mma = (double)h / mma_per_smsp;
    const double tflops = 2.0 * 16 * 8 * 16 / cyc_per_mma * 4 * 148 * ghz * 1e9 / 1e12;
    printf("independent accumulators %d, warps/SMSP %d: %.2f cycles per MMA per SMSP, %.0f dense-equivalent TFLOP/s at %.2f GHz\n", NACC, warps_per_smsp,
           cyc_per_mma, tflops, ghz);
  }
}

int main() {
  float* out; uint32_t* in; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 32 * 6 * 4); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 32 * 6 * 4);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  run<1>(out, in, cyc, ghz);
  run<4>(out, in, cyc, ghz);
  run<8>(out, in, cyc, ghz);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
