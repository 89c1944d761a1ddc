// Attention tail and FFN head of a block in one persistent kernel (/root/reference/models/hit_sir_pro.py:597 proj, :700-703 norm1 +
// residual, :39-41 fc1 + GELU):
//     x  = x + LayerNorm(outsc Wp^T + bp)              (tcgen05, accumulator D1 in TMEM, LayerNorm in the epilogue)
//     h1 = GELU(x W1^T + b1)                           (tcgen05, the bf16 copy of x never leaves shared memory)
// As two launches the pair moved, per 128 tokens, 456 KB from L2 to the SM (both weight matrices re-streamed for every tile, the bf16
// shadow of x written and read back) and was bound by exactly that (ablation); chained, the shadow round trip (768 B
// per token of HBM traffic) and one pass over the activations disappear.
//
// Warp roles (640 threads): 0 = TMA producer (A k-blocks, then nine 24 KB weight units per tile: Wp k-blocks 0..2, W1 row chunks 0..5),
// 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4..19 = epilogue (TMEM lane quarter x 16-column slice of every 64-column group).
// The residual row slices are fetched with plain 16-byte loads before the accumulator is waited for (64 contiguous bytes per thread:
// whole sectors), x and h1 leave the same way; there is no staging buffer besides the A operand of the second contraction.
#include "gemm.cuh"
#include "kernels.cuh"

namespace hitsir {

namespace {

constexpr int kBlkA = 128 * 128;                           // [128 tokens x 64 K] bf16, SWIZZLE_128B
constexpr int kUnitW = 192 * 128;                          // weight unit: Wp k-block [192 N x 64 K] or W1 chunk 3 x [64 N x 64 K]
constexpr int kNW = 3;                                     // weight ring slots
constexpr int kOffA1 = 0;                                  // 3 k-blocks of the current tile's SCC output
constexpr int kOffW = kOffA1 + 3 * kBlkA;
constexpr int kOffA2 = kOffW + kNW * kUnitW;               // 3 k-blocks: bf16(x) of the current tile
constexpr int kOffPar = kOffA2 + 3 * kBlkA;                // bp | gamma | beta (3 x 192 fp32) | b1 (384 fp32)
constexpr int kOffPart = kOffPar + (3 * 192 + 384) * 4;    // LayerNorm partials [2][4][128] float2
constexpr int kOffBars = kOffPart + 2 * 4 * 128 * 8;
constexpr int kNumBars = 28;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16 + 1024;
static_assert(kSmemBytes <= 232448, "smem budget");
constexpr int kTmD1 = 0;                                   // proj accumulator [128 x 192] (its next use follows the fc1 chunks of the tile anyway)
constexpr int kTmD2 = 192;                                 // four fc1 chunk accumulators [128 x 64]: the MMA issuer runs up to four chunks ahead

struct Params {
  long long N;                // tokens
  int tiles;
  const float* bp; const float* gamma; const float* beta; const float* b1;
  const float* res;           // fp32 [N][180] residual stream in
  float* xout;                // fp32 [N][180] stream out
  bf16* h1;                   // bf16 [N][384]
};

__device__ __forceinline__ void epi_bar_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

__global__ void __launch_bounds__(640, 1)
proj_fc1_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_wp, const __grid_constant__ CUtensorMap tm_w1,
                const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sb - smem_u32(smem_raw));
  float* s_bp = reinterpret_cast<float*>(sp + kOffPar);
  float* s_gamma = s_bp + 192;
  float* s_beta = s_gamma + 192;
  float* s_b1 = s_beta + 192;
  float2* s_part = reinterpret_cast<float2*>(sp + kOffPart);
  const uint32_t bar0 = sb + kOffBars;
  auto a1_full = [&](int s) { return bar0 + 8u * s; };
  auto a1_empty = [&](int s) { return bar0 + 8u * (3 + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (6 + s); };
  auto w_empty = [&](int s) { return bar0 + 8u * (9 + s); };
  auto d2_full = [&](int s) { return bar0 + 8u * (12 + s); };
  auto d2_empty = [&](int s) { return bar0 + 8u * (16 + s); };
  const uint32_t d1_full = bar0 + 8u * 20, d1_empty = bar0 + 8u * 21, a2_full = bar0 + 8u * 22, a2_empty = bar0 + 8u * 23;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sp + kOffBars + kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_wp); tma_prefetch_desc(&tm_w1); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(a1_full(s), 1); mbar_init(a1_empty(s), 1); mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(d2_full(s), 1); mbar_init(d2_empty(s), 16); }
    mbar_init(d1_full, 1); mbar_init(d1_empty, 16); mbar_init(a2_full, 16); mbar_init(a2_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(tmem_ptr_smem), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 192; i += blockDim.x) {
    s_bp[i] = p.bp[i];
    s_gamma[i] = i < kC ? p.gamma[i] : 0.f;
    s_beta[i] = i < kC ? p.beta[i] : 0.f;
  }
  for (int i = threadIdx.x; i < 384; i += blockDim.x) s_b1[i] = p.b1[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== producer =====================
      int it = 0;
      uint32_t wu = 0;                                       // weight units issued so far (ring position)
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const int row0 = t * 128;
        for (int kb = 0; kb < 3; ++kb) {
          mbar_wait(a1_empty(kb), ((uint32_t)it & 1u) ^ 1u);
          mbar_expect_tx(a1_full(kb), kBlkA);
          tma_load_2d(sb + kOffA1 + kb * kBlkA, &tm_a, a1_full(kb), kb * 64, row0);
          const int ws = (int)(wu % kNW);
          mbar_wait(w_empty(ws), ((wu / kNW) & 1u) ^ 1u);
          mbar_expect_tx(w_full(ws), kUnitW);
          tma_load_2d(sb + kOffW + ws * kUnitW, &tm_wp, w_full(ws), kb * 64, 0);
          ++wu;
        }
        for (int c = 0; c < 6; ++c) {
          const int ws = (int)(wu % kNW);
          mbar_wait(w_empty(ws), ((wu / kNW) & 1u) ^ 1u);
          mbar_expect_tx(w_full(ws), kUnitW);
          for (int kb = 0; kb < 3; ++kb) tma_load_2d(sb + kOffW + ws * kUnitW + kb * (64 * 128), &tm_w1, w_full(ws), kb * 64, c * 64);
          ++wu;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc1 = umma_idesc_bf16(128, 192);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 64);
      int it = 0;
      uint32_t wu = 0, cu = 0;                               // weight units / fc1 chunks consumed so far
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        mbar_wait(d1_empty, ((uint32_t)it & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d1 = tmem_base + (uint32_t)kTmD1;
        for (int kb = 0; kb < 3; ++kb) {
          const int ws = (int)(wu % kNW);
          mbar_wait(a1_full(kb), (uint32_t)it & 1u);
          mbar_wait(w_full(ws), (wu / kNW) & 1u);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(sb + kOffA1 + kb * kBlkA);
          const uint64_t bdesc = umma_desc_sw128(sb + kOffW + ws * kUnitW);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d1, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc1, (kb | kk) != 0 ? 1u : 0u);
          umma_commit(a1_empty(kb));
          umma_commit(w_empty(ws));
          ++wu;
        }
        umma_commit(d1_full);
        mbar_wait(a2_full, (uint32_t)it & 1u);               // the epilogue has written bf16(x) of this tile
        tc_fence_after();
        for (int c = 0; c < 6; ++c, ++cu) {
          const int ws = (int)(wu % kNW), cs = (int)(cu & 3u);
          mbar_wait(d2_empty(cs), ((cu >> 2) & 1u) ^ 1u);
          mbar_wait(w_full(ws), (wu / kNW) & 1u);
          tc_fence_after();
          const uint32_t d2 = tmem_base + (uint32_t)(kTmD2 + cs * 64);
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) {
            const uint64_t adesc = umma_desc_sw128(sb + kOffA2 + kb * kBlkA);
            const uint64_t bdesc = umma_desc_sw128(sb + kOffW + ws * kUnitW + kb * (64 * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d2, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc2, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(w_empty(ws));
          umma_commit(d2_full(cs));
          ++wu;
        }
        umma_commit(a2_empty);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int cw = warp - 4;
    const int q = warp & 3, hs = cw >> 2;                    // TMEM lane quarter, 16-column slice of every 64-column group
    const int r = q * 32 + lane;
    const uint32_t rsw = (uint32_t)(r & 7);
    int it = 0;
    uint32_t cu = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const long long row = (long long)t * 128 + r;
      const bool row_ok = row < p.N;
      // ---------- residual slices of this row: 3 groups x 16 columns, fetched before the accumulator is waited for
      float4 res[3][4];
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int c0 = 64 * g + 16 * hs + 4 * ch;
          res[g][ch] = (row_ok && c0 < kC) ? __ldg(reinterpret_cast<const float4*>(p.res + row * kC + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      mbar_wait(d1_full, (uint32_t)it & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kTmD1;
      // LayerNorm statistics without cancellation: per 16-column group (mean, M2), merged with Chan's parallel update
      float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int c0 = 64 * g + 16 * hs;
        float v[16];
        tmem_ld16(tacc + c0, v);
        const int cnt = min(16, max(0, kC - c0));
        float sg = 0.f, qg = 0.f, mg;
        if (cnt == 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] += s_bp[c0 + i]; sg += v[i]; }
          mg = sg * (1.0f / 16.0f);
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = v[i] - mg; qg = fmaf(d, d, qg); }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] += s_bp[c0 + i]; if (i < cnt) sg += v[i]; }
          mg = sg / (float)max(cnt, 1);
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = v[i] - mg; if (i < cnt) qg = fmaf(d, d, qg); }
        }
        if (cnt > 0) {
          const float nn = n + (float)cnt, d = mg - mean;
          mean += d * ((float)cnt / nn);
          m2 += qg + d * d * (n * (float)cnt / nn);
          n = nn;
        }
      }
      float2* part = s_part + (it & 1) * 512;
      part[hs * 128 + r] = make_float2(mean, m2);
      epi_bar_sync(q);
#pragma unroll
      for (int o = 1; o < 4; ++o) {
        const int ho = (hs + o) & 3;
        const float2 tv = part[ho * 128 + r];
        const float cnt = ho == 3 ? 36.f : 48.f;             // real columns of slice ho: 3 x 16, the last slice ends at column 180
        const float nn = n + cnt, d = tv.x - mean;
        mean += d * (cnt / nn);
        m2 += tv.y + d * d * (n * cnt / nn);
        n = nn;
      }
      const float rstd = rsqrtf(m2 * (1.0f / (float)kC) + 1e-5f);
      mbar_wait(a2_empty, ((uint32_t)it & 1u) ^ 1u);         // the fc1 MMAs of the previous tile have read the A2 buffer
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int c0 = 64 * g + 16 * hs;
        float v[16];
        tmem_ld16(tacc + c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf((v[i] + s_bp[c0 + i] - mean) * rstd, s_gamma[c0 + i], s_beta[c0 + i]);   // gamma = beta = 0 beyond 180
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 rr = res[g][ch];
          v[4 * ch] += rr.x; v[4 * ch + 1] += rr.y; v[4 * ch + 2] += rr.z; v[4 * ch + 3] += rr.w;
          if (row_ok && c0 + 4 * ch < kC)                    // 180 = 45 chunks of 4: a chunk is entirely real or entirely padding
            *reinterpret_cast<float4*>(p.xout + row * kC + c0 + 4 * ch) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
        }
        // bf16(x) -> A operand of fc1: k-block g, row r, 16-byte chunks 2 hs and 2 hs + 1 (SWIZZLE_128B)
        uint8_t* arow = sp + kOffA2 + g * kBlkA + r * 128;
        *reinterpret_cast<uint4*>(arow + ((((uint32_t)(2 * hs)) ^ rsw) << 4)) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        *reinterpret_cast<uint4*>(arow + ((((uint32_t)(2 * hs + 1)) ^ rsw) << 4)) =
            make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(a2_full);
      tc_fence_before();
      mbar_arrive_warp(d1_empty);
      // ---------- fc1 chunks: bias + GELU -> bf16 h1, 32 contiguous bytes per thread and chunk
#pragma unroll 1
      for (int c = 0; c < 6; c += 2, cu += 2) {              // two chunks per TMEM wait
        const int cs0 = (int)(cu & 3u), cs1 = (int)((cu + 1) & 3u);
        uint32_t ra[16], rb[16];
        mbar_wait(d2_full(cs0), (cu >> 2) & 1u);
        tc_fence_after();
        tmem_ld16_nw(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kTmD2 + cs0 * 64 + 16 * hs), ra);
        mbar_wait(d2_full(cs1), ((cu + 1) >> 2) & 1u);
        tc_fence_after();
        tmem_ld16_nw(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kTmD2 + cs1 * 64 + 16 * hs), rb);
        tmem_ld_wait();
        reg_fence16(ra);
        reg_fence16(rb);
        tc_fence_before();
        mbar_arrive_warp(d2_empty(cs0));
        mbar_arrive_warp(d2_empty(cs1));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* rr = half ? rb : ra;
          const int n0 = (c + half) * 64 + 16 * hs;
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float2 gl = gelu2(make_float2(__uint_as_float(rr[i]) + s_b1[n0 + i], __uint_as_float(rr[i + 1]) + s_b1[n0 + i + 1]));
            o[i >> 1] = (n0 + i < kHid) ? pack_bf16x2(gl.x, gl.y) : 0u;        // 360 is even: a pair is entirely real or entirely padding
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.h1 + row * kHidp + n0);
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// outsc: bf16 [N][192] (head-padded SCC output); tm_wp: packed proj weights, box {64, 192}; w1: packed fc1 weights bf16 [384][192];
// res / xout: fp32 [N][180] (may alias); h1: bf16 [N][384]
int launch_proj_fc1(const bf16* outsc, const CUtensorMap& tm_wp, const float* bp, const float* gamma, const float* beta, const float* res,
                    float* xout, const bf16* w1, const float* b1, bf16* h1, long long N, int num_sms, cudaStream_t st) {
  static unsigned long long configured = 0;
  if (ensure_dynamic_smem(proj_fc1_kernel, kSmemBytes, &configured)) return 1;
  Params p;
  p.N = N;
  const long long tiles = (N + 127) / 128;
  if (tiles > 2147483647LL / 128) { set_error("launch_proj_fc1: too many tokens"); return 1; }
  p.tiles = (int)tiles;
  p.bp = bp; p.gamma = gamma; p.beta = beta; p.b1 = b1; p.res = res; p.xout = xout; p.h1 = h1;
  CUtensorMap tm_a, tm_w1;
  if (make_tmap_2d(&tm_a, outsc, kCp, (uint64_t)N, (uint64_t)kCp * 2, 64, 128)) return 1;
  if (make_tmap_2d(&tm_w1, w1, kCp, kHidp, (uint64_t)kCp * 2, 64, 64)) return 1;
  const int grid = p.tiles < num_sms ? p.tiles : num_sms;
  if (grid <= 0) return 0;
  proj_fc1_kernel<<<grid, 640, kSmemBytes, st>>>(tm_a, tm_wp, tm_w1, p);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace hitsir
