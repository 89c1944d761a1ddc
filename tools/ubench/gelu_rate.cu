// Micro-benchmark: throughput of candidate erf-GELU evaluations on sm_100a (values per clock per SM), to balance the FMA pipe
// against the MUFU pipe in the FFN epilogues (csrc/common.cuh gelu2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gelu_rate gelu_rate.cu && ./gelu_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int NV = 16;   // independent values per thread

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// current product form: packed, 2 MUFU per value
__device__ __forceinline__ float2 gelu2_mufu(float2 x) {
  float2 s = __fmul2_rn(x, x);
  s.x = fminf(s.x, 81.f); s.y = fminf(s.y, 81.f);
  float2 p = __ffma2_rn(s, make_float2(9.481962249e-04f, 9.481962249e-04f), make_float2(-1.064097551e-01f, -1.064097551e-01f));
  p = __ffma2_rn(p, s, make_float2(-2.301458255f, -2.301458255f));
  const float2 z = __fmul2_rn(p, x);
  const float2 e = __fadd2_rn(make_float2(ex2_approx(z.x), ex2_approx(z.y)), make_float2(1.f, 1.f));
  return __fmul2_rn(x, make_float2(rcp_approx(e.x), rcp_approx(e.y)));
}
// same maths, scalar instructions with immediate coefficients
__device__ __forceinline__ float gelu1_mufu(float x) {
  const float s = fminf(x * x, 81.f);
  float p = fmaf(s, 9.481962249e-04f, -1.064097551e-01f);
  p = fmaf(p, s, -2.301458255f);
  const float e = ex2_approx(p * x) + 1.f;
  return x * rcp_approx(e);
}
// polynomial only (degree 9 in x^2), scalar with immediates: no MUFU
__device__ __forceinline__ float gelu1_poly(float x) {
  const float xc = fminf(fmaxf(x, -4.242640687f), 4.242640687f);
  const float s = xc * xc;
  float p = -3.086732500e-12f;
  p = fmaf(p, s, 3.179519986e-10f);
  p = fmaf(p, s, -1.470231326e-08f);
  p = fmaf(p, s, 4.085154930e-07f);
  p = fmaf(p, s, -7.745199668e-06f);
  p = fmaf(p, s, 1.082166939e-04f);
  p = fmaf(p, s, -1.169120996e-03f);
  p = fmaf(p, s, 9.949907623e-03f);
  p = fmaf(p, s, -6.647990253e-02f);
  p = fmaf(p, s, 3.989525639e-01f);
  return x * fmaf(xc, p, 0.5f);
}
__device__ __forceinline__ float2 gelu2_poly(float2 x) {
  const float X = 4.242640687f;
  const float2 xc = make_float2(fminf(fmaxf(x.x, -X), X), fminf(fmaxf(x.y, -X), X));
  const float2 s = __fmul2_rn(xc, xc);
  float2 p = make_float2(-3.086732500e-12f, -3.086732500e-12f);
  p = __ffma2_rn(p, s, make_float2(3.179519986e-10f, 3.179519986e-10f));
  p = __ffma2_rn(p, s, make_float2(-1.470231326e-08f, -1.470231326e-08f));
  p = __ffma2_rn(p, s, make_float2(4.085154930e-07f, 4.085154930e-07f));
  p = __ffma2_rn(p, s, make_float2(-7.745199668e-06f, -7.745199668e-06f));
  p = __ffma2_rn(p, s, make_float2(1.082166939e-04f, 1.082166939e-04f));
  p = __ffma2_rn(p, s, make_float2(-1.169120996e-03f, -1.169120996e-03f));
  p = __ffma2_rn(p, s, make_float2(9.949907623e-03f, 9.949907623e-03f));
  p = __ffma2_rn(p, s, make_float2(-6.647990253e-02f, -6.647990253e-02f));
  p = __ffma2_rn(p, s, make_float2(3.989525639e-01f, 3.989525639e-01f));
  const float2 phi = __ffma2_rn(xc, p, make_float2(0.5f, 0.5f));
  return __fmul2_rn(x, phi);
}

// V: 0 packed mufu, 1 scalar mufu, 2 scalar poly, 3 packed poly, 10+k: of every 8 values k scalar-poly, 8-k packed-mufu; 20+k: k scalar-poly, 8-k scalar-mufu
template <int V>
__global__ void k(float* out, const float* in, long long* cyc) {
  float v[NV];
  for (int i = 0; i < NV; ++i) v[i] = in[threadIdx.x + 32 * i];
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if (V == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) { float2 g = gelu2_mufu(make_float2(v[i], v[i + 1])); v[i] = g.x + 0.25f; v[i + 1] = g.y + 0.25f; }
    } else if (V == 1) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = gelu1_mufu(v[i]) + 0.25f;
    } else if (V == 2) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = gelu1_poly(v[i]) + 0.25f;
    } else if (V == 3) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) { float2 g = gelu2_poly(make_float2(v[i], v[i + 1])); v[i] = g.x + 0.25f; v[i + 1] = g.y + 0.25f; }
    } else if (V >= 10 && V < 20) {
      constexpr int K = V - 10;
#pragma unroll
      for (int b = 0; b < NV; b += 8) {
#pragma unroll
        for (int i = 0; i < K; ++i) v[b + i] = gelu1_poly(v[b + i]) + 0.25f;
#pragma unroll
        for (int i = K; i < 8; i += 2) { float2 g = gelu2_mufu(make_float2(v[b + i], v[b + i + 1])); v[b + i] = g.x + 0.25f; v[b + i + 1] = g.y + 0.25f; }
      }
    } else {
      constexpr int K = V - 20;
#pragma unroll
      for (int b = 0; b < NV; b += 8) {
#pragma unroll
        for (int i = 0; i < K; ++i) v[b + i] = gelu1_poly(v[b + i]) + 0.25f;
#pragma unroll
        for (int i = K; i < 8; ++i) v[b + i] = gelu1_mufu(v[b + i]) + 0.25f;
      }
    }
  }
  long long t1 = clock64();
  float r = 0.f;
  for (int i = 0; i < NV; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
void run(const char* name, float* out, float* in, long long* cyc) {
  for (int warps_per_smsp : {4, 5}) {
    const int threads = 128 * warps_per_smsp;
    k<V><<<148, threads>>>(out, in, cyc);
    k<V><<<148, threads>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_clk_sm = (double)NV * ITER * threads / (double)h;
    printf("%-52s warps/SMSP %d: %9lld cycles, %6.2f GELU/clk/SM (%.2f cyc per warp-value per SMSP)\n", name, warps_per_smsp, h, per_clk_sm, 128.0 / per_clk_sm);
  }
}

int main() {
  float *out, *in; long long* cyc;
  cudaMalloc(&out, 148 * 640 * 4); cudaMalloc(&in, 32 * 32 * 4); cudaMalloc(&cyc, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = -3.f + 6.f * (float)((i * 37) % 1024) / 1024.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("packed logistic-quintic (2 MUFU / value) [product]", out, in, cyc);
  run<1>("scalar logistic-quintic (2 MUFU / value)", out, in, cyc);
  run<2>("scalar degree-9 polynomial (0 MUFU)", out, in, cyc);
  run<3>("packed degree-9 polynomial (0 MUFU)", out, in, cyc);
  run<12>("mix 2/8 scalar-poly + 6/8 packed-mufu", out, in, cyc);
  run<14>("mix 4/8 scalar-poly + 4/8 packed-mufu", out, in, cyc);
  run<22>("mix 2/8 scalar-poly + 6/8 scalar-mufu", out, in, cyc);
  run<23>("mix 3/8 scalar-poly + 5/8 scalar-mufu", out, in, cyc);
  run<24>("mix 4/8 scalar-poly + 4/8 scalar-mufu", out, in, cyc);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
