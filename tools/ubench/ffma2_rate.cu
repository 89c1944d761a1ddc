// Micro-benchmark: issue rate of packed FP32 FMA (FFMA2) on sm_100a as a function of operand sharing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
// Each variant runs ITER x 16 independent FMAs per thread with `warps` warps per SM sub-partition and reports
// cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int V>
__global__ void k(float2* out, const float2* in, long long* cyc) {
  float2 acc[16], a[8], b[8];
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * (i + 8)]; }
  for (int i = 0; i < 16; ++i) acc[i] = in[threadIdx.x + 32 * (i + 16)];
  float s0 = a[0].x, s1 = b[0].x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (V == 0) {            // FFMA2, three distinct register pairs
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(a[i & 7], b[(i + 3) & 7], acc[i]);
    } else if (V == 1) {     // FFMA2, multiplier shared by 4 consecutive instructions (weight reuse)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(a[i & 7], b[i >> 2], acc[i]);
    } else if (V == 2) {     // FFMA2, same multiplicand AND multiplier pattern (both shared by runs)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(a[i >> 2], b[i >> 2], acc[i]);
    } else if (V == 3) {     // scalar FFMA, three distinct registers (32 per iteration = same FMA count)
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc[i].x = fmaf(a[i & 7].x, b[(i + 3) & 7].x, acc[i].x); acc[i].y = fmaf(a[i & 7].y, b[(i + 3) & 7].y, acc[i].y); }
    } else if (V == 4) {     // scalar FFMA, shared multiplier
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc[i].x = fmaf(a[i & 7].x, s1, acc[i].x); acc[i].y = fmaf(a[i & 7].y, s1, acc[i].y); }
    } else if (V == 5) {     // scalar FFMA, shared multiplier and multiplicand per run of 4
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc[i].x = fmaf(a[i >> 2].x, b[i >> 2].x, acc[i].x); acc[i].y = fmaf(a[i >> 2].y, b[i >> 2].y, acc[i].y); }
    }
  }
  long long t1 = clock64();
  float2 r = make_float2(0.f, 0.f);
  for (int i = 0; i < 16; ++i) { r.x += acc[i].x; r.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  (void)s0;
}

template <int V>
void run(const char* name, float2* out, float2* in, long long* cyc) {
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
    k<V><<<148, threads>>>(out, in, cyc);
    k<V><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double fma_per_thread = 32.0 * ITER;                  // every variant: 32 FMAs per thread per iteration
    const double per_clk_sm = fma_per_thread * threads / (double)h;
    printf("%-44s warps/SMSP %d: %9lld cycles, %6.1f FMA/clk/SM\n", name, warps_per_smsp, h, per_clk_sm);
  }
}

int main() {
  float2 *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&in, 32 * 32 * 8); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 32 * 32 * 8);
  run<0>("FFMA2 3 distinct operands", out, in, cyc);
  run<1>("FFMA2 multiplier shared by runs of 4", out, in, cyc);
  run<2>("FFMA2 both inputs shared by runs of 4", out, in, cyc);
  run<3>("FFMA  3 distinct operands", out, in, cyc);
  run<4>("FFMA  one multiplier for all", out, in, cyc);
  run<5>("FFMA  both inputs shared by runs of 4", out, in, cyc);
  cudaError_t e = cudaGetLastError();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
