// Weight-only precomputation, run once per weight load (never in the forward):
//   * fp32 nn.Linear / nn.Conv2d tensors -> bf16 K-major GEMM operands padded for TMA/UMMA;
//   * the DynamicPosBias MLP table and its pooled relative-position bias (hit_sir_pro.py:477-503),
//     which the reference rebuilds on every forward of every block although it only depends on weights.
#include "kernels.cuh"

namespace hitsir {

namespace {

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  return (int)(g < 4096 ? (g > 0 ? g : 1) : 4096);
}

__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ wp, float* __restrict__ bp,
                                 int Co, int Ci, int taps, int Npad, int Cipad, int perm_k) {
  const long long K = (long long)taps * Cipad;
  const long long total = (long long)Npad * K;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / K); const int k = (int)(idx - (long long)n * K);
    const int tap = k / Cipad, ci = k - tap * Cipad;
    float v = 0.f;
    const int cs = perm_k ? scc_chan(ci) : (ci < Ci ? ci : -1);
    if (n < Co && cs >= 0) v = w[((long long)n * Ci + cs) * taps + tap];     // [Co][Ci][kh][kw], tap = kh*kw_size + kw
    wp[idx] = __float2bfloat16(v);
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < Npad; n += gridDim.x * blockDim.x) bp[n] = (n < Co && b != nullptr) ? b[n] : 0.f;
}

// 3x3 conv applied to a x2 nearest-upsampled map == four 2x2 convs on the ORIGINAL map, one per output phase (a, b) = (Y & 1, X & 1):
// HR row 2y + a + ky - 1 is LR row y + floor((a + ky - 1) / 2), so for a = 0 the taps ky = {0} | {1, 2} fall on rows y - 1 | y and for
// a = 1 the taps {0, 1} | {2} on rows y | y + 1 (same along x).  The taps that share a source pixel are summed in fp32 and rounded to
// bf16 once.  Layout [64 co][16 x 64]: column (phase * 4 + dyi * 2 + dxi) * 64 + ci, phase = 2a + b, LR offset dy = a - 1 + dyi.
// (interpolate(scale_factor=2, mode='nearest') + conv_up1 / conv_up2, /root/reference/models/hit_sir_pro.py:1331-1332)
__global__ void pack_subpixel_kernel(const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ wp, float* __restrict__ bp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 64 * 1024) {
    const int co = idx >> 10, k = idx & 1023;
    const int t = k >> 6, ci = k & 63;
    const int ph = t >> 2, dyi = (t >> 1) & 1, dxi = t & 1;
    const int a = ph >> 1, bb = ph & 1;
    // tap range of (phase bit, offset index): (0,0) -> {0}, (0,1) -> {1,2}, (1,0) -> {0,1}, (1,1) -> {2}
    const int ky0 = a == 0 ? (dyi == 0 ? 0 : 1) : (dyi == 0 ? 0 : 2), ky1 = a == 0 ? (dyi == 0 ? 0 : 2) : (dyi == 0 ? 1 : 2);
    const int kx0 = bb == 0 ? (dxi == 0 ? 0 : 1) : (dxi == 0 ? 0 : 2), kx1 = bb == 0 ? (dxi == 0 ? 0 : 2) : (dxi == 0 ? 1 : 2);
    float v = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) v += w[((co * 64 + ci) * 3 + ky) * 3 + kx];
    wp[idx] = __float2bfloat16(v);
  }
  if (idx < 64) bp[idx] = b[idx];
}

// conv_last 64 -> Co <= 4 with the taps folded into N: row tap * 4 + co of [48][64] holds w[co][:][tap] (rows with co >= Co and rows 36.. are zero)
__global__ void pack_fold_last_kernel(const float* __restrict__ w, bf16* __restrict__ wp, int Co) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 48 * 64) return;
  const int n = idx >> 6, ci = idx & 63;
  const int tap = n >> 2, co = n & 3;
  float v = 0.f;
  if (tap < 9 && co < Co) v = w[(co * 64 + ci) * 9 + tap];
  wp[idx] = __float2bfloat16(v);
}

// rows: tile j (0..5) x slot q (0..4: conv3,5,7,9,conv_x) x 32 embedding channels c = 32j + ci (zero rows for c >= 180);
// K = 9x9 footprint x in_ch.  One 160-row N tile therefore holds all five responses of 32 channels (EPI_MSGATE).
__global__ void pack_msconv_kernel(const float* __restrict__ w3, const float* __restrict__ w5, const float* __restrict__ w7, const float* __restrict__ w9,
                                   const float* __restrict__ wx, const float* __restrict__ b3, const float* __restrict__ b5, const float* __restrict__ b7,
                                   const float* __restrict__ b9, const float* __restrict__ bx, bf16* __restrict__ wp, float* __restrict__ bp, int in_ch, int Kp) {
  const long long total = 960LL * Kp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / Kp), k = (int)(idx - (long long)n * Kp);
    const int j = n / 160, rem = n - j * 160;
    const int q = rem / 32, c = 32 * j + (rem & 31);
    float v = 0.f;
    if (c < kC && k < 81 * in_ch) {
      const int tap = k / in_ch, ci = k - tap * in_ch;
      const int ky = tap / 9, kx = tap - ky * 9;
      const int s = (q < 4) ? (3 + 2 * q) : 1;        // filter size
      const int off = (9 - s) / 2;
      const int fy = ky - off, fx = kx - off;
      if (fy >= 0 && fy < s && fx >= 0 && fx < s) {
        const float* src = q == 0 ? w3 : q == 1 ? w5 : q == 2 ? w7 : q == 3 ? w9 : wx;
        v = src[(((long long)c * in_ch + ci) * s + fy) * s + fx];
      }
    }
    wp[idx] = __float2bfloat16(v);
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < 960; n += gridDim.x * blockDim.x) {
    const int j = n / 160, rem = n - j * 160;
    const int q = rem / 32, c = 32 * j + (rem & 31);
    float v = 0.f;
    if (c < kC) {
      const float* src = q == 0 ? b3 : q == 1 ? b5 : q == 2 ? b7 : q == 3 ? b9 : bx;
      v = src[c];
    }
    bp[n] = v;
  }
}

// conv_first.conv_last (1x1, 4C -> C) for the gated concat stored as [tile j][slot k][ci]: wp[n][j*128 + k*32 + ci] = w[n][k*C + 32j + ci]
__global__ void pack_mslast_kernel(const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ wp, float* __restrict__ bp, int Npad) {
  const long long total = (long long)Npad * 768;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / 768), kk = (int)(idx - (long long)n * 768);
    const int j = kk >> 7, k = (kk >> 5) & 3, c = 32 * j + (kk & 31);
    float v = 0.f;
    if (n < kC && c < kC) v = w[(long long)n * 4 * kC + k * kC + c];
    wp[idx] = __float2bfloat16(v);
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < Npad; n += gridDim.x * blockDim.x) bp[n] = n < kC ? b[n] : 0.f;
}

__global__ void pack_firstconv_kernel(const float* __restrict__ w, const float* __restrict__ b, bf16* __restrict__ wp, float* __restrict__ bp,
                                      int Co, int in_ch, int f, int Kp) {
  const long long total = 192LL * Kp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx / Kp), k = (int)(idx - (long long)n * Kp);
    float v = 0.f;
    if (n < Co && k < f * f * in_ch) {
      const int tap = k / in_ch, ci = k - tap * in_ch;
      v = w[((long long)n * in_ch + ci) * f * f + tap];
    }
    wp[idx] = __float2bfloat16(v);
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < 192; n += gridDim.x * blockDim.x) bp[n] = n < Co ? b[n] : 0.f;
}

__global__ void pack_tapmajor_kernel(const float* __restrict__ w, float* __restrict__ out, int C, int taps, int Cpad) {
  const int total = taps * Cpad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx / Cpad, c = idx - tap * Cpad;
    out[idx] = c < C ? w[c * taps + tap] : 0.f;
  }
}

// fc2 weights [192 rows][384 k] bf16 (K-major) -> six operand images of [192 rows x 64 k]: the byte image a SWIZZLE_128B TMA box would
// leave in shared memory (row r = 128 bytes, 16-byte chunk c stored at chunk position c ^ (r & 7)); thread = one 16-byte chunk
__global__ void pack_w2_image_kernel(const bf16* __restrict__ w, uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 6 * 192 * 8) return;
  const int c = idx & 7, r = (idx >> 3) % 192, k = idx / (192 * 8);
  const uint4 v = *reinterpret_cast<const uint4*>(w + (size_t)r * kHidp + k * 64 + c * 8);
  *reinterpret_cast<uint4*>(img + (size_t)k * (192 * 128) + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

// depthwise taps as MMA B-fragment words (ffn_tail.cu): one row of 28 words per channel, bf16(w) in the half selected by the channel parity,
// in the order in which the 13 MMAs of an output row pair the taps: words 4 ky + dx (dx = 0..3), 20 + ky for the taps (ky, 4) with ky < 4,
// 24 = tap (4,4), 26 = the bias as fp32 bits, 25 and 27 = 0.  tbl = fp32 tap-major table [26][384] (row 25 = bias).
__global__ void pack_dw_mma_kernel(const float* __restrict__ tbl, uint32_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kHidp * 28) return;
  const int c = idx / 28, wd = idx - c * 28;
  int tap = -1;
  if (wd < 20) tap = (wd >> 2) * 5 + (wd & 3);
  else if (wd < 24) tap = (wd - 20) * 5 + 4;
  else if (wd == 24) tap = 24;
  uint32_t v = 0u;
  if (tap >= 0) {
    const bf16 h = __float2bfloat16_rn(tbl[tap * kHidp + c]);
    v = (uint32_t)(*reinterpret_cast<const uint16_t*>(&h)) << (16 * (c & 1));
  } else if (wd == 26) {
    v = __float_as_uint(tbl[25 * kHidp + c]);
  }
  out[idx] = v;
}

__device__ __forceinline__ void ln_relu(float* v, int n, const float* g, const float* b) {
  float m = 0.f;
  for (int i = 0; i < n; ++i) m += v[i];
  m /= (float)n;
  float q = 0.f;
  for (int i = 0; i < n; ++i) { const float d = v[i] - m; q += d * d; }
  const float rstd = rsqrtf(q / (float)n + 1e-5f);
  for (int i = 0; i < n; ++i) v[i] = fmaxf((v[i] - m) * rstd * g[i] + b[i], 0.f);
}

// DynamicPosBias (residual=False): Linear 2->11, then 3 x (LayerNorm, ReLU, Linear); (:305-313)
__global__ void pos_table_kernel(PosW w, int win, float* __restrict__ tbl) {
  constexpr int D = 11;
  const int side = 2 * win - 1;
  const int total = side * side;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const float dy = (float)(idx / side - (win - 1)), dx = (float)(idx % side - (win - 1));
    float a[D], t[D];
    for (int o = 0; o < D; ++o) a[o] = w.proj_w[o * 2] * dy + w.proj_w[o * 2 + 1] * dx + w.proj_b[o];
    for (int s = 0; s < 2; ++s) {
      ln_relu(a, D, w.ln_w[s], w.ln_b[s]);
      for (int o = 0; o < D; ++o) { float acc = w.fc_b[s][o]; for (int i = 0; i < D; ++i) acc += w.fc_w[s][o * D + i] * a[i]; t[o] = acc; }
      for (int o = 0; o < D; ++o) a[o] = t[o];
    }
    ln_relu(a, D, w.ln_w[2], w.ln_b[2]);
    for (int o = 0; o < kHeads; ++o) {
      float acc = w.fc_b[2][o];
      for (int i = 0; i < D; ++i) acc += w.fc_w[2][o * D + i] * a[i];
      tbl[idx * kHeads + o] = acc;
    }
  }
}

// bias[h][l][cell] = mean over the r x r tokens m of the cell of tbl[(yl-ym+w-1)*(2w-1) + (xl-xm+w-1)][h]   (:486-501)
__global__ void pooled_bias_kernel(const float* __restrict__ tbl, int win, int base, float* __restrict__ out) {
  const int r = win / base, L = win * win, Lb = base * base, side = 2 * win - 1;
  const long long total = (long long)kHeads * L * Lb;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(idx % Lb); const long long t = idx / Lb; const int l = (int)(t % L); const int h = (int)(t / L);
    const int yl = l / win, xl = l - yl * win;
    const int cy = cell / base, cx = cell - cy * base;
    float s = 0.f;
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < r; ++j) {
        const int ym = cy * r + i, xm = cx * r + j;
        s += tbl[((yl - ym + win - 1) * side + (xl - xm + win - 1)) * kHeads + h];
      }
    out[idx] = s / (float)(r * r);          // a true division, like the reference's .mean(-1) (:501): a constant table stays exact
  }
}

}  // namespace

int launch_pack_conv(const float* w, const float* b, bf16* wp, float* bp, int Co, int Ci, int taps, int Npad, int Cipad, int perm_k, cudaStream_t st) {
  pack_conv_kernel<<<grid_for((long long)Npad * taps * Cipad, 256), 256, 0, st>>>(w, b, wp, bp, Co, Ci, taps, Npad, Cipad, perm_k);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_subpixel(const float* w, const float* b, bf16* wp, float* bp, cudaStream_t st) {
  pack_subpixel_kernel<<<64 * 1024 / 256, 256, 0, st>>>(w, b, wp, bp);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_fold_last(const float* w, bf16* wp, int Co, cudaStream_t st) {
  pack_fold_last_kernel<<<48 * 64 / 256, 256, 0, st>>>(w, wp, Co);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_msconv(const float* w3, const float* w5, const float* w7, const float* w9, const float* wx, const float* b3, const float* b5,
                       const float* b7, const float* b9, const float* bx, bf16* wp, float* bp, int in_ch, int Kp, cudaStream_t st) {
  pack_msconv_kernel<<<grid_for(960LL * Kp, 256), 256, 0, st>>>(w3, w5, w7, w9, wx, b3, b5, b7, b9, bx, wp, bp, in_ch, Kp);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_mslast(const float* w, const float* b, bf16* wp, float* bp, int Npad, cudaStream_t st) {
  pack_mslast_kernel<<<grid_for((long long)Npad * 768, 256), 256, 0, st>>>(w, b, wp, bp, Npad);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_firstconv(const float* w, const float* b, bf16* wp, float* bp, int Co, int in_ch, int f, int Kp, cudaStream_t st) {
  pack_firstconv_kernel<<<grid_for(192LL * Kp, 256), 256, 0, st>>>(w, b, wp, bp, Co, in_ch, f, Kp);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_tapmajor(const float* w, float* out, int C, int taps, int Cpad, cudaStream_t st) {
  pack_tapmajor_kernel<<<grid_for((long long)taps * Cpad, 256), 256, 0, st>>>(w, out, C, taps, Cpad);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_w2_image(const bf16* w2_packed, uint8_t* img, cudaStream_t st) {
  pack_w2_image_kernel<<<grid_for(6 * 192 * 8, 256), 256, 0, st>>>(w2_packed, img);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pack_dw_mma(const float* dw_tbl, uint32_t* out, cudaStream_t st) {
  pack_dw_mma_kernel<<<grid_for(28 * kHidp, 256), 256, 0, st>>>(dw_tbl, out);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pos_table(PosW w, int win, float* tbl, cudaStream_t st) {
  const int total = (2 * win - 1) * (2 * win - 1);
  pos_table_kernel<<<grid_for(total, 128), 128, 0, st>>>(w, win, tbl);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}
int launch_pooled_bias(const float* tbl, int win, int base, float* out, cudaStream_t st) {
  const long long total = (long long)kHeads * win * win * base * base;
  pooled_bias_kernel<<<grid_for(total, 256), 256, 0, st>>>(tbl, win, base, out);
  HITSIR_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace hitsir
