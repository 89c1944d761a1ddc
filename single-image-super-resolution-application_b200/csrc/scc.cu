// Spatial-channel self-correlation inside hierarchical windows (SCC.forward without the final
// proj, /root/reference/models/hit_sir_pro.py:542-596).  Round-1 formulation: fp32 SIMT math on
// bf16 tokens staged in shared memory, 64 tokens per step.
//
// Per window (w x w tokens, L = w*w; pooled grid base x base, Lb = base*base, r = w/base):
//   t  = qkv token (180) = [q | v], channel = half*90 + head*15 + j          (:569-570)
//   k  = (Wk1 q_h + bk1 + Wk2 v_h + bk2) / 2  per head                       (:572)
//   kp = pool(k), vp = pool(v):  cell(cy,cx) = sum_{i,j<r} wsl[i*r+j] * tok(cy*r+i, cx*r+j) + bsl   (:451-455)
//   S-SC: out_s[l,h,:] = sum_m (q_lh . kp[m,h] / 15 + bias[h,l,m]) * vp[m,h,:]       (:475,503,511)
//   C-SC: corr[c,c'] = sum_l q[l,c] k[l,c'] / L ;  out_c[l,c] = sum_c' corr[c,c'] v[l,c']   (:531,538)
//   out[l] = [out_s | out_c]                                                    (:596)
//
// Phase A (window reduction: corr, kp, vp) and phase B (per-token application) run in one CTA
// when the window has <= 256 tokens; larger windows are split over `parts` CTAs per phase with a
// deterministic partial-sum reduction in between (no float atomics -> bitwise reproducible).
#include "kernels.cuh"

namespace hitsir {

namespace {

constexpr int kChunk = 64;            // tokens per step
constexpr int kTS = 180;              // sT row stride (floats)
constexpr int kPS = 91;               // pooled row stride (odd -> conflict-free over cells)
constexpr int kWinFloats = kHalf * kHalf;   // corr 90x90

struct Smem {
  float T[kChunk * kTS];              // chunk tokens fp32 [64][180]
  float K[kChunk * kHalf];            // chunk keys [64][90]
  float KP[64 * kPS];                 // pooled keys  [Lb][91]
  float VP[64 * kPS];                 // pooled values
  float CorrT[kHalf * kHalf];         // corrT[c'][c] = corr[c][c'] / L
  float S[8 * kHeads * 64];           // per-warp score scratch
  float wk1[kHd * kHd], wk2[kHd * kHd], bk[kHd];
  float wsl[64];
  float bsl;
};

__device__ __forceinline__ long long tok_index(const SccGeom& g, int b, int wy, int wx, int l) {
  const int ly = l / g.w, lx = l - ly * g.w;
  return ((long long)b * g.pg.Hp + (wy * g.w + ly)) * g.pg.Wp + (wx * g.w + lx);
}

// load `n` tokens [l0, l0+n) of a window as fp32 into sm.T
__device__ __forceinline__ void load_chunk(Smem& sm, const bf16* __restrict__ t, const SccGeom& g, int b, int wy, int wx, int l0, int n) {
  // 180 bf16 = 90 bf16x2 words per token
  for (int idx = threadIdx.x; idx < n * 90; idx += blockDim.x) {
    const int i = idx / 90, c2 = idx - i * 90;
    const uint32_t u = *reinterpret_cast<const uint32_t*>(t + tok_index(g, b, wy, wx, l0 + i) * kCp + 2 * c2);
    const float2 f = unpack_bf16x2(u);
    sm.T[i * kTS + 2 * c2] = f.x;
    sm.T[i * kTS + 2 * c2 + 1] = f.y;
  }
}

// k for every (token, head) of the chunk
__device__ __forceinline__ void compute_k(Smem& sm, int n) {
  for (int idx = threadIdx.x; idx < n * kHeads; idx += blockDim.x) {
    const int i = idx / kHeads, h = idx - i * kHeads;
    const float* q = sm.T + i * kTS + h * kHd;
    const float* v = q + kHalf;
    float qv[kHd], vv[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) { qv[j] = q[j]; vv[j] = v[j]; }
#pragma unroll
    for (int o = 0; o < kHd; ++o) {
      float a = sm.bk[o];
#pragma unroll
      for (int j = 0; j < kHd; ++j) a += 0.5f * (sm.wk1[o * kHd + j] * qv[j] + sm.wk2[o * kHd + j] * vv[j]);
      sm.K[i * kHalf + h * kHd + o] = a;
    }
  }
}

// corr accumulation: thread (a, bb) owns the 6x6 tile corr[6a.., 6bb..]; 225 active threads
__device__ __forceinline__ void accum_corr(const Smem& sm, int n, float (&acc)[36]) {
  const int tid = threadIdx.x;
  if (tid >= 225) return;
  const int a = tid / 15, bb = tid - a * 15;
  for (int i = 0; i < n; ++i) {
    const float* q = sm.T + i * kTS + 6 * a;
    const float* k = sm.K + i * kHalf + 6 * bb;
    float qv[6], kv[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { qv[j] = q[j]; kv[j] = k[j]; }
#pragma unroll
    for (int x = 0; x < 6; ++x)
#pragma unroll
      for (int y = 0; y < 6; ++y) acc[x * 6 + y] += qv[x] * kv[y];
  }
}

// pooled sums: thread c (<180) owns column c of KP (c<90) or VP (c>=90)
__device__ __forceinline__ void accum_pool(Smem& sm, const SccGeom& g, int l0, int n) {
  const int c = threadIdx.x;
  if (c >= 2 * kHalf) return;
  float* dst = (c < kHalf) ? (sm.KP + c) : (sm.VP + (c - kHalf));
  for (int i = 0; i < n; ++i) {
    const int l = l0 + i;
    const int ly = l / g.w, lx = l - ly * g.w;
    const int cy = ly / g.r, cx = lx / g.r;
    const float wgt = sm.wsl[(ly - cy * g.r) * g.r + (lx - cx * g.r)];
    const float val = (c < kHalf) ? sm.K[i * kHalf + c] : sm.T[i * kTS + c];   // v channel = 90 + (c-90) = c
    dst[(cy * g.base + cx) * kPS] += wgt * val;
  }
}

// per-token application; one warp per token
__device__ __forceinline__ void apply_chunk(Smem& sm, const SccGeom& g, const SccW& w, int b, int wy, int wx, int l0, int n, bf16* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sS = sm.S + warp * kHeads * 64;
  const float inv_d = 1.0f / (float)kHd;
  for (int i = warp; i < n; i += 8) {
    const int l = l0 + i;
    const int ly = l / g.w, lx = l - ly * g.w;
    const int yp = wy * g.w + ly, xp = wx * g.w + lx;
    if (yp >= g.pg.H || xp >= g.pg.W) continue;          // reflect-padded token: result is cropped (:696)
    const float* tk = sm.T + i * kTS;
    // scores
    for (int m = lane; m < g.Lb; m += 32) {
#pragma unroll
      for (int h = 0; h < kHeads; ++h) {
        const float* q = tk + h * kHd;
        const float* kp = sm.KP + m * kPS + h * kHd;
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < kHd; ++j) a += q[j] * kp[j];
        sS[h * 64 + m] = a * inv_d + __ldg(w.bias_tbl + ((long long)h * g.L + l) * g.Lb + m);
      }
    }
    __syncwarp();
    const long long orow = (((long long)b * g.pg.H + yp) * g.pg.W + xp) * kCp;
    for (int c = lane; c < kHalf; c += 32) {
      const int h = c / kHd;
      float a = 0.f;
      for (int m = 0; m < g.Lb; ++m) a += sS[h * 64 + m] * sm.VP[m * kPS + c];
      // channel correlation
      float cc = 0.f;
      const float* v = tk + kHalf;
#pragma unroll 6
      for (int e = 0; e < kHalf; ++e) cc += sm.CorrT[e * kHalf + c] * v[e];
      out[orow + c] = __float2bfloat16(a);
      out[orow + kHalf + c] = __float2bfloat16(cc);
    }
    if (lane < kCp - kC) out[orow + kC + lane] = __float2bfloat16(0.f);
    __syncwarp();
  }
}

__device__ __forceinline__ void load_weights(Smem& sm, const SccGeom& g, const SccW& w) {
  for (int i = threadIdx.x; i < kHd * kHd; i += blockDim.x) { sm.wk1[i] = w.wk1[i]; sm.wk2[i] = w.wk2[i]; }
  if (threadIdx.x < kHd) sm.bk[threadIdx.x] = 0.5f * (w.bk1[threadIdx.x] + w.bk2[threadIdx.x]);
  if (threadIdx.x < g.r * g.r) sm.wsl[threadIdx.x] = w.wsl[threadIdx.x];
  if (threadIdx.x == 0) sm.bsl = *w.bsl_dev;
}

// mode 0: fused (one CTA per window); mode 1: phase A over this CTA's token range -> partials;
// mode 2: phase B over this CTA's token range, reading reduced window results from `finals`.
__global__ void __launch_bounds__(256, 1) scc_kernel(const bf16* __restrict__ t, const SccGeom g, const SccW w, int mode,
                                                     float* __restrict__ partials, const float* __restrict__ finals, bf16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int win = blockIdx.x / g.parts, part = blockIdx.x - win * g.parts;
  const int wx = win % g.nWx; const int t2 = win / g.nWx; const int wy = t2 % g.nWy; const int b = t2 / g.nWy;
  const int per = g.L / g.parts;               // tokens of this CTA
  const int l_begin = part * per, l_end = l_begin + per;
  const int win_floats = kWinFloats + 2 * g.Lb * kHalf;

  load_weights(sm, g, w);
  if (mode != 2) {
    for (int i = threadIdx.x; i < 64 * kPS; i += blockDim.x) { sm.KP[i] = 0.f; sm.VP[i] = 0.f; }
    float acc[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0.f;
    __syncthreads();
    for (int l0 = l_begin; l0 < l_end; l0 += kChunk) {
      const int n = min(kChunk, l_end - l0);
      load_chunk(sm, t, g, b, wy, wx, l0, n);
      __syncthreads();
      compute_k(sm, n);
      __syncthreads();
      accum_corr(sm, n, acc);
      accum_pool(sm, g, l0, n);
      __syncthreads();
    }
    if (mode == 0) {
      const float invL = 1.0f / (float)g.L;
      if (threadIdx.x < 225) {
        const int a = threadIdx.x / 15, bb = threadIdx.x - a * 15;
#pragma unroll
        for (int x = 0; x < 6; ++x)
#pragma unroll
          for (int y = 0; y < 6; ++y) sm.CorrT[(6 * bb + y) * kHalf + 6 * a + x] = acc[x * 6 + y] * invL;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < g.Lb * kHalf; i += blockDim.x) {
        const int m = i / kHalf, c = i - m * kHalf;
        sm.KP[m * kPS + c] += sm.bsl;
        sm.VP[m * kPS + c] += sm.bsl;
      }
      __syncthreads();
    } else {
      float* dst = partials + ((long long)win * g.parts + part) * win_floats;
      if (threadIdx.x < 225) {
        const int a = threadIdx.x / 15, bb = threadIdx.x - a * 15;
#pragma unroll
        for (int x = 0; x < 6; ++x)
#pragma unroll
          for (int y = 0; y < 6; ++y) dst[(6 * bb + y) * kHalf + 6 * a + x] = acc[x * 6 + y];
      }
      for (int i = threadIdx.x; i < g.Lb * kHalf; i += blockDim.x) {
        const int m = i / kHalf, c = i - m * kHalf;
        dst[kWinFloats + i] = sm.KP[m * kPS + c];
        dst[kWinFloats + g.Lb * kHalf + i] = sm.VP[m * kPS + c];
      }
      return;
    }
  } else {
    const float* src = finals + (long long)win * win_floats;
    for (int i = threadIdx.x; i < kWinFloats; i += blockDim.x) sm.CorrT[i] = src[i];
    for (int i = threadIdx.x; i < g.Lb * kHalf; i += blockDim.x) {
      const int m = i / kHalf, c = i - m * kHalf;
      sm.KP[m * kPS + c] = src[kWinFloats + i];
      sm.VP[m * kPS + c] = src[kWinFloats + g.Lb * kHalf + i];
    }
    __syncthreads();
  }
  // phase B
  for (int l0 = l_begin; l0 < l_end; l0 += kChunk) {
    const int n = min(kChunk, l_end - l0);
    if (!(mode == 0 && g.L <= kChunk)) {       // single-chunk windows still hold their tokens in sm.T
      __syncthreads();
      load_chunk(sm, t, g, b, wy, wx, l0, n);
    }
    __syncthreads();
    apply_chunk(sm, g, w, b, wy, wx, l0, n, out);
  }
}

// finals[win] = sum over parts (fixed order) ; corr scaled by 1/L, pooled + bsl
__global__ void scc_reduce_kernel(const float* __restrict__ partials, float* __restrict__ finals, const SccGeom g, const float* __restrict__ bsl_dev) {
  const int win_floats = kWinFloats + 2 * g.Lb * kHalf;
  const long long total = (long long)g.pg.B * g.nWy * g.nWx * win_floats;
  const float invL = 1.0f / (float)g.L;
  const float bsl = *bsl_dev;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long win = idx / win_floats; const int e = (int)(idx - win * win_floats);
    float s = 0.f;
    for (int p = 0; p < g.parts; ++p) s += partials[(win * g.parts + p) * win_floats + e];
    finals[idx] = (e < kWinFloats) ? s * invL : s + bsl;
  }
}

}  // namespace

int scc_workspace_floats(const SccGeom& g, long long* partial_floats, long long* final_floats) {
  const long long nwin = (long long)g.pg.B * g.nWy * g.nWx;
  const long long wf = kWinFloats + 2LL * g.Lb * kHalf;
  if (g.parts > 1) { *partial_floats = nwin * g.parts * wf; *final_floats = nwin * wf; }
  else { *partial_floats = 0; *final_floats = 0; }
  return 0;
}

int launch_scc(const bf16* t, const SccGeom& g, const SccW& w, float* partials, float* finals, bf16* out, cudaStream_t st) {
  static bool configured = false;
  const int smem = (int)sizeof(Smem);
  if (!configured) {
    HITSIR_CHECK(cudaFuncSetAttribute(scc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const long long nwin = (long long)g.pg.B * g.nWy * g.nWx;
  if (nwin * g.parts > 2147483647LL) { set_error("launch_scc: grid too large"); return 1; }
  if (g.parts == 1) {
    scc_kernel<<<(unsigned)nwin, 256, smem, st>>>(t, g, w, 0, nullptr, nullptr, out);
  } else {
    scc_kernel<<<(unsigned)(nwin * g.parts), 256, smem, st>>>(t, g, w, 1, partials, nullptr, out);
    HITSIR_CHECK(cudaGetLastError());
    scc_reduce_kernel<<<1184, 256, 0, st>>>(partials, finals, g, w.bsl_dev);
    HITSIR_CHECK(cudaGetLastError());
    scc_kernel<<<(unsigned)(nwin * g.parts), 256, smem, st>>>(t, g, w, 2, nullptr, finals, out);
  }
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace hitsir
