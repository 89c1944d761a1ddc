// Micro-benchmark: per-SM TMA load / store throughput as a function of box shape (row width) and boxes in flight, all 148 SMs
// streaming at once from a tensor larger than L2.  Question: is a box with many 128-byte rows bound by bytes or by rows?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda && ./tma_rate
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mode 0: 2-D loads; 1: 3-D loads (halo-style box: inner, w, h); 2: 2-D stores; 3: contiguous bulk loads of box_bytes
__global__ void __launch_bounds__(64, 1)
k(const __grid_constant__ CUtensorMap tm, const uint8_t* base, int mode, int stages, int box_bytes, int iters, int rows_per_box, int total_rows, int inner_tiles,
  long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    // box i of this CTA: rows [(blockIdx.x + i * gridDim.x) * rows_per_box ...) wrapped, inner tile cycles
    auto coords = [&](int i, int* c0, int* r0) {
      long long boxi = (long long)blockIdx.x + (long long)i * gridDim.x;
      *c0 = (int)(boxi % inner_tiles);
      *r0 = (int)(((boxi / inner_tiles) * rows_per_box) % (total_rows - rows_per_box));
    };
    if (mode == 2) {
      for (int i = 0; i < iters; ++i) {
        int c0, r0; coords(i, &c0, &r0);
        tma_store_2d(&tm, sb + (i % stages) * box_bytes, c0 * (box_bytes / rows_per_box / 2), r0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (stages == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else if (stages == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      for (int i = 0; i < iters + stages; ++i) {
        const int s = i % stages;
        if (i >= stages) mbar_wait(smem_u32(&bars[s]), ((i / stages) - 1) & 1);
        if (i < iters) {
          int c0, r0; coords(i, &c0, &r0);
          mbar_expect_tx(smem_u32(&bars[s]), box_bytes);
          if (mode == 0) tma_load_2d(sb + s * box_bytes, &tm, smem_u32(&bars[s]), c0 * (box_bytes / rows_per_box / 2), r0);
          else if (mode == 1) tma_load_3d(sb + s * box_bytes, &tm, smem_u32(&bars[s]), c0 * 64, (r0 % 200), (r0 / 256) % 3000);
          else bulk_load(sb + s * box_bytes, base + ((long long)r0 * 768 + (long long)c0 * 0) % (1LL << 30), box_bytes, smem_u32(&bars[s]));
        }
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cyc = t1 - t0;
  }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fp;
  const size_t bytes = 3ULL << 30;   // 3 GB >> L2
  uint8_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
  long long* cyc; cudaMalloc(&cyc, 8);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { const char* name; int mode; int inner_elems; int rows; CUtensorMapSwizzle sw; };
  // tensor: rows of 384 bf16 (768-byte pitch), like the hidden map H1; 3-D: (384, 256 w, 4096 h)
  Cfg cfgs[] = {
      {"load 2D 128 rows x 128 B (SW128)", 0, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B},
      {"load 2D 256 rows x  64 B (SW64) ", 0, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B},
      {"load 2D  64 rows x 256 B (none) ", 0, 128, 64, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"load 2D  32 rows x 512 B (none) ", 0, 256, 32, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"load 3D halo 12x20 x 128 B (SW128)", 1, 64, 240, CU_TENSOR_MAP_SWIZZLE_128B},
      {"store 2D 128 rows x 128 B (SW128)", 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B},
      {"bulk load 16 KB contiguous", 3, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    const int total_rows = (int)(bytes / 768);
    CUresult r;
    if (c.mode == 1) {
      cuuint64_t dims[3] = {384, 256, (cuuint64_t)(total_rows / 256)}; cuuint64_t strides[2] = {768, 768 * 256};
      cuuint32_t box[3] = {64, 20, 12}; cuuint32_t es[3] = {1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[2] = {384, (cuuint64_t)total_rows}; cuuint64_t strides[1] = {768};
      cuuint32_t box[2] = {(cuuint32_t)c.inner_elems, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int box_bytes = c.inner_elems * 2 * c.rows;
    const int inner_tiles = 384 / c.inner_elems;
    for (int stages : {2, 4, 6}) {
      if (stages * box_bytes > 190 * 1024) continue;
      const int iters = 2000;
      for (int rep = 0; rep < 2; ++rep) k<<<148, 64, 200 * 1024>>>(tm, buf, c.mode, stages, box_bytes, iters, c.mode == 1 ? 12 : c.rows, total_rows, inner_tiles, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double bpc = (double)box_bytes * iters / (double)h;
      printf("%-36s box %6d B, %d in flight: %7.1f B/clk/SM = %6.1f GB/s/SM = %5.2f TB/s chip, %6.1f cyc/box, %5.2f cyc/row  (%s)\n", c.name, box_bytes, stages, bpc,
             bpc * clk_khz * 1e3 / 1e9, bpc * clk_khz * 1e3 / 1e9 * 148 / 1e3, (double)h / iters, (double)h / iters / c.rows, cudaGetErrorString(e));
    }
  }
  return 0;
}
