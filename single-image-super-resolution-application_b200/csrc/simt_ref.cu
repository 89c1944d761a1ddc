// SIMT cross-check for the tcgen05 contraction kernel: one thread per output row, operands read
// straight from global memory, SAME fused epilogue (gemm.cuh).  Selected with
// HITSIR_GEMM=simt; exists so that a parity failure can be localised to either the tensor-core
// pipeline or the surrounding dataflow.  Never the default and far too slow for benchmarks.
#include "gemm.cuh"

namespace hitsir {

struct SimtAcc {
  const GemmParams* p;
  RowInfo ri;
  int n0;
  __device__ __forceinline__ float dot(int col) const {
    const GemmParams& P = *p;
    const bf16* wrow = P.Wp + (long long)(n0 + col) * P.ldw;
    float acc = 0.f;
    if (!P.conv) {
      if (!ri.valid) return 0.f;
      const bf16* arow = P.A + ri.grow * P.lda;
      const int K = P.num_kb * 64;
      for (int k = 0; k < K; ++k) acc += __bfloat162float(arow[k]) * __bfloat162float(wrow[k]);
    } else {
      const int cin = P.cblocks * 64;
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = ri.y + tap / 3 - 1, xx = ri.x + tap % 3 - 1;
        if (yy < 0 || yy >= P.H || xx < 0 || xx >= P.W) continue;
        const bf16* arow = P.A + (((long long)ri.b * P.H + yy) * P.W + xx) * P.lda;
        const bf16* wr = wrow + tap * cin;
        for (int c = 0; c < cin; ++c) acc += __bfloat162float(arow[c]) * __bfloat162float(wr[c]);
      }
    }
    return acc;
  }
  __device__ __forceinline__ void load16(int c0, float* v) {
#pragma unroll 1
    for (int i = 0; i < 16; ++i) v[i] = dot(c0 + i);
  }
};

template <int BN>
__global__ void __launch_bounds__(128) simt_gemm_kernel(const GemmParams p) {
  const int n_tile = blockIdx.x % p.n_tiles, m_tile = blockIdx.x / p.n_tiles;
  const int r = threadIdx.x;
  SimtAcc acc;
  acc.p = &p;
  acc.ri = row_info(p, m_tile, r);
  acc.n0 = n_tile * BN;
  epilogue_row<BN>(p, acc, acc.ri, n_tile);
}

int launch_simt_gemm(int BN, const GemmParams& p, cudaStream_t st) {
  const int grid = p.m_tiles * p.n_tiles;
  if (grid <= 0) return 0;
  switch (BN) {
    case 16: simt_gemm_kernel<16><<<grid, 128, 0, st>>>(p); break;
    case 32: simt_gemm_kernel<32><<<grid, 128, 0, st>>>(p); break;
    case 48: simt_gemm_kernel<48><<<grid, 128, 0, st>>>(p); break;
    case 64: simt_gemm_kernel<64><<<grid, 128, 0, st>>>(p); break;
    case 160: simt_gemm_kernel<160><<<grid, 128, 0, st>>>(p); break;
    case 192: simt_gemm_kernel<192><<<grid, 128, 0, st>>>(p); break;
    case 256: simt_gemm_kernel<256><<<grid, 128, 0, st>>>(p); break;
    default: set_error("launch_simt_gemm: unsupported N tile %d", BN); return 1;
  }
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace hitsir
