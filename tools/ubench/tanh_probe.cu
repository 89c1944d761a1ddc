// tanh.approx.f32 near saturation: does it reach exactly 1?  (build: nvcc -arch=sm_100a -o tanh_probe tanh_probe.cu)
#include <cstdio>
__global__ void k(float* out, const float* in, int n) {
  int i = threadIdx.x;
  if (i < n) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(in[i])); out[i] = y; }
}
int main() {
  const int n = 16;
  float h[n] = {0.5f, 1.f, 2.f, 3.f, 4.f, 4.5f, 5.f, 6.f, 8.f, 9.f, 10.f, 15.f, 20.f, 50.f, 88.f, 1000.f};
  float *di, *dout, o[n];
  cudaMalloc(&di, sizeof(h)); cudaMalloc(&dout, sizeof(h));
  cudaMemcpy(di, h, sizeof(h), cudaMemcpyHostToDevice);
  k<<<1, 32>>>(dout, di, n);
  cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i) printf("w %8.2f  tanh.approx %.9g  1-t %.3e  (exact 1-tanh %.3e)\n", h[i], o[i], 1.0 - (double)o[i], 1.0 - tanh((double)h[i]));
  return 0;
}
