// Memory-bound kernels of the DINOSeg hot path (HBM-bound: coalesced 16-byte accesses,
// warp-shuffle reductions, no tensor cores).
#pragma once
#include "ptx.cuh"

namespace dsg {

// ---------------------------------------------------------------------------------------
// fp32 -> bf16 conversion (weight packing at load time)
// ---------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}

// "bf16x3" operand split.  A product x*w with both operands rounded to bf16 carries a 2^-9 relative
// error; writing x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) and using
//     x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// brings it to ~2^-16 while staying on the bf16 tensor cores: the three products are obtained from ONE
// GEMM of three-fold K over the operands A' = [hi | lo | hi], W' = [hi | hi | lo].  A' is stored as [hi | lo] only:
// the GEMM producer wraps its A column at 2*K (GemmParams::a_wrap), so the third segment re-reads hi (from L2).
//  Used for the
// segmentation head (0.7 % of the FLOPs), whose bf16 rounding otherwise dominates the log-prob error.
// W [N, K] fp32 -> W' [N, 3*Kp] bf16 (each part zero padded from K to Kp columns)
__global__ void split_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N, int K, int Kp) {
  const size_t total = size_t(N) * Kp;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int n = int(i / Kp), k = int(i - size_t(n) * Kp);
    float v = k < K ? w[size_t(n) * K + k] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + size_t(n) * 3 * Kp + k;
    o[0] = hi; o[Kp] = hi; o[2 * Kp] = lo;
  }
}

// ---------------------------------------------------------------------------------------
// Positional-embedding table for a g x g patch grid (reference vision_transformer.py:202-222):
// row 0 = pos[0] (cls); rows 1.. = bicubic resample of the G0 x G0 source grid with torch's
// F.interpolate(scale_factor=(g+0.1)/G0, mode='bicubic', align_corners=False) semantics:
//   src = rscale*(dst+0.5)-0.5 with rscale = float(1/scale_factor); taps floor(src)-1..+2 clamped
//   to [0,G0-1]; cubic-convolution weights with A = -0.75; inner sum over x, outer over y.
// Runs once per set_resolution(); the reference recomputes it every forward.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * t - (A + 3.0f)) * t * t + 1.0f;
  const float t1 = 1.0f - t;
  w[2] = ((A + 2.0f) * t1 - (A + 3.0f)) * t1 * t1 + 1.0f;
  const float x3 = 2.0f - t;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

__global__ void posembed_bicubic_kernel(const float* __restrict__ pos_src /*[G0*G0+1, D]*/,
                                        float* __restrict__ out /*[g*g+1, D]*/, int G0, int g, int D,
                                        float rscale) {
  const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = size_t(g * g + 1) * D;
  if (idx >= total) return;
  const int t = int(idx / D);
  const int d = int(idx - size_t(t) * D);
  if (t == 0) { out[idx] = pos_src[d]; return; }
  const int oy = (t - 1) / g, ox = (t - 1) - oy * g;
  const float sy = rscale * (float(oy) + 0.5f) - 0.5f;
  const float sx = rscale * (float(ox) + 0.5f) - 0.5f;
  int iy = min(int(floorf(sy)), G0 - 1);
  int ix = min(int(floorf(sx)), G0 - 1);
  const float ty = fminf(fmaxf(sy - float(iy), 0.f), 1.f);
  const float tx = fminf(fmaxf(sx - float(ix), 0.f), 1.f);
  float wy[4], wx[4];
  cubic_coeffs(ty, wy);
  cubic_coeffs(tx, wx);
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), G0 - 1);
    float rowacc = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int xx = min(max(ix - 1 + b, 0), G0 - 1);
      rowacc += wx[b] * pos_src[size_t(1 + yy * G0 + xx) * D + d];
    }
    acc += wy[a] * rowacc;
  }
  out[idx] = acc;
}

// ---------------------------------------------------------------------------------------
// im2col for the 8x8/stride-8 patch-embed conv (reference vision_transformer.py:153,157):
//   A[b*P + i*g + j][c*64 + ky*8 + kx] = frame[b][c][i*8+ky][j*8+kx]
// written as the bf16x3 operand [hi | lo] (2 x 192 columns, see split_weight_kernel): the rounding of
// the pixels / conv weights to plain bf16 is the largest single contribution to the final log-prob error.
// One thread moves two image rows of one patch (2 x 32 B in, 2 x 32 B out).
// ---------------------------------------------------------------------------------------
constexpr int IM2COL_K = 192;
constexpr int IM2COL_K3 = 3 * IM2COL_K;   // reduction length of the patch-embed GEMM
constexpr int IM2COL_KA = 2 * IM2COL_K;   // stored columns of its A operand [hi | lo]

__device__ __forceinline__ void split_pack8(const float4& a0, const float4& a1, uint4& hi, uint4& lo) {
  hi.x = pack_bf16x2(a0.x, a0.y); hi.y = pack_bf16x2(a0.z, a0.w);
  hi.z = pack_bf16x2(a1.x, a1.y); hi.w = pack_bf16x2(a1.z, a1.w);
  lo.x = pack_bf16x2(a0.x - __uint_as_float(hi.x << 16), a0.y - __uint_as_float(hi.x & 0xffff0000u));
  lo.y = pack_bf16x2(a0.z - __uint_as_float(hi.y << 16), a0.w - __uint_as_float(hi.y & 0xffff0000u));
  lo.z = pack_bf16x2(a1.x - __uint_as_float(hi.z << 16), a1.y - __uint_as_float(hi.z & 0xffff0000u));
  lo.w = pack_bf16x2(a1.z - __uint_as_float(hi.w << 16), a1.w - __uint_as_float(hi.w & 0xffff0000u));
}

__global__ void im2col_patch8_kernel(const float* __restrict__ frames, __nv_bfloat16* __restrict__ A, int B,
                                     int g) {
  const int r = g * 8;
  const size_t total = size_t(B) * 3 * g * 4 * g;  // (b, c, i, kyp, j)
  size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = int(idx % g); idx /= g;
  const int kyp = int(idx % 4); idx /= 4;
  const int i = int(idx % g); idx /= g;
  const int c = int(idx % 3);
  const int b = int(idx / 3);
  const float* src = frames + ((size_t(b) * 3 + c) * r + (i * 8 + kyp * 2)) * r + j * 8;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(src + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(src + r));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(src + r + 4));
  uint4 h0, l0, h1, l1;
  split_pack8(a0, a1, h0, l0);
  split_pack8(b0, b1, h1, l1);
  __nv_bfloat16* dst = A + (size_t(b) * g * g + size_t(i) * g + j) * IM2COL_KA + c * 64 + kyp * 16;
  *reinterpret_cast<uint4*>(dst) = h0;
  *reinterpret_cast<uint4*>(dst + 8) = h1;
  *reinterpret_cast<uint4*>(dst + IM2COL_K) = l0;
  *reinterpret_cast<uint4*>(dst + IM2COL_K + 8) = l1;
}

// ---------------------------------------------------------------------------------------
// Preprocessing of DINOSeg.predict fused into the im2col (reference pl_torch_modules.py:33-41, :291):
//   uint8 HWC RGB frame [H, W, 3] -> albumentations Resize(r, r) = cv2.resize(INTER_LINEAR) ->
//   Normalize: (pix - mean*255) * (1 / (std*255)) in fp32 -> CHW -> 8x8 patches (bf16x3 operand as above).
// The resize reproduces OpenCV's 8-bit bilinear path in integer arithmetic: source coordinate
// (d + 0.5) * src/dst - 0.5, taps clamped at the borders, 11-bit fixed-point weights
// (INTER_RESIZE_COEF_BITS), horizontal pass in int32, vertical pass
// ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2, i.e. the uint8 image cv2 would hand to
// Normalize.  Same-size frames (H = W = r) pass through exactly.  One thread = 8 output pixels of one row.
// ---------------------------------------------------------------------------------------
struct PreprocParams {
  int src_h, src_w;
  float mean255[3];     // float32(mean) * 255
  float denom[3];       // 1 / (float32(std) * 255)
};

// cv::resize INTER_LINEAR coefficients for destination index d: f = (d + 0.5) * scale - 0.5 (in fp32), s = floor(f),
// weights saturate_cast<short>(w * 2048) (round half to even).  Horizontally OpenCV zeroes the fraction at the
// borders, vertically it only clamps the two rows.
__device__ __forceinline__ void cv_linear_coef(int d, double scale, int src_size, bool horizontal, int& s0, int& s1,
                                               int& a0, int& a1) {
  float f = float((double(d) + 0.5) * scale - 0.5);
  int sx = int(floorf(f));
  f -= float(sx);
  if (horizontal) {
    if (sx < 0) { f = 0.f; sx = 0; }
    if (sx >= src_size - 1) { f = 0.f; sx = src_size - 1; }
  }
  s0 = min(max(sx, 0), src_size - 1);
  s1 = min(max(sx + 1, 0), src_size - 1);
  a0 = __float2int_rn((1.f - f) * 2048.f);
  a1 = __float2int_rn(f * 2048.f);
}

__global__ void im2col_u8_kernel(const uint8_t* __restrict__ frames /*[B, H, W, 3]*/, __nv_bfloat16* __restrict__ A,
                                 int B, int g, PreprocParams pp) {
  const int r = g * 8;
  const size_t total = size_t(B) * 3 * r * g;      // (b, c, y, j): 8 pixels each
  size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = int(idx % g); idx /= g;
  const int y = int(idx % r); idx /= r;
  const int c = int(idx % 3);
  const int b = int(idx / 3);
  const uint8_t* src = frames + size_t(b) * pp.src_h * pp.src_w * 3;
  const bool same = pp.src_h == r && pp.src_w == r;
  const double sy = double(pp.src_h) / double(r), sx = double(pp.src_w) / double(r);
  int y0, y1, b0, b1;
  cv_linear_coef(y, sy, pp.src_h, false, y0, y1, b0, b1);
  float v[8];
#pragma unroll
  for (int kx = 0; kx < 8; ++kx) {
    const int x = j * 8 + kx;
    int pix;
    if (same) {
      pix = src[(size_t(y) * pp.src_w + x) * 3 + c];
    } else {
      int x0, x1, a0, a1;
      cv_linear_coef(x, sx, pp.src_w, true, x0, x1, a0, a1);
      const int r0 = int(src[(size_t(y0) * pp.src_w + x0) * 3 + c]) * a0 + int(src[(size_t(y0) * pp.src_w + x1) * 3 + c]) * a1;
      const int r1 = int(src[(size_t(y1) * pp.src_w + x0) * 3 + c]) * a0 + int(src[(size_t(y1) * pp.src_w + x1) * 3 + c]) * a1;
      pix = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
      pix = pix < 0 ? 0 : (pix > 255 ? 255 : pix);
    }
    v[kx] = (float(pix) - pp.mean255[c]) * pp.denom[c];
  }
  uint4 hi, lo;
  split_pack8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
  const int i = y >> 3, ky = y & 7;
  __nv_bfloat16* dst = A + (size_t(b) * g * g + size_t(i) * g + j) * IM2COL_KA + c * 64 + ky * 8;
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + IM2COL_K) = lo;
}

// x[b*Ntok + 0, :] = cls + pos[0]      (reference vision_transformer.py:229-233)
__global__ void cls_row_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                               int B, int Ntok, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x[size_t(b) * Ntok * D + d] = cls[d] + pos[d];
}

// ---------------------------------------------------------------------------------------
// LayerNorm (eps inside the sqrt, biased variance; reference :114,:118,:183 with eps=1e-6 :303)
// fp32 in -> bf16 out (the bf16 copy is the A operand of the following GEMM).  One warp per row.
// ---------------------------------------------------------------------------------------
// Row statistics for D = 384 with EIGHT LANES PER ROW (a warp works on four rows at once): lane `sub` of a row's group
// holds the row's float4 number i*8 + sub, i = 0..11.  Two-pass statistics in registers, sums over the group with three
// xor-shuffles - one shuffle instruction serves four rows, so the dependent shuffle chain per row is 6 long instead of
// the 10 of a warp-per-row reduction (what bounds a single warp that has to normalise many rows in a row: the fused
// LayerNorm fill of mlp.cuh, which shares this function with the LayerNorm kernel so that both produce the same bits).
constexpr int LN384_V = 12;
__device__ __forceinline__ void ln384_stats(const float4 (&v)[LN384_V], float eps, float& mean, float& rstd) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN384_V; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  mean = sum * (1.0f / 384.0f);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN384_V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
  sq += __shfl_xor_sync(0xffffffffu, sq, 1);
  sq += __shfl_xor_sync(0xffffffffu, sq, 2);
  sq += __shfl_xor_sync(0xffffffffu, sq, 4);
  rstd = rsqrtf(sq * (1.0f / 384.0f) + eps);
}
// normalised, affine-transformed float4 number i*8 + sub of the row -> four bf16 (two packed words)
__device__ __forceinline__ uint2 ln384_out(const float4& v, float mean, float rstd, const float4& gm, const float4& bt) {
  const float a = (v.x - mean) * rstd * gm.x + bt.x, b = (v.y - mean) * rstd * gm.y + bt.y;
  const float c = (v.z - mean) * rstd * gm.z + bt.z, d = (v.w - mean) * rstd * gm.w + bt.w;
  uint2 o;
  o.x = pack_bf16x2(a, b);
  o.y = pack_bf16x2(c, d);
  return o;
}

// normalised float4 WITHOUT the affine transform (the consumer's weights carry gamma and beta: fold_ln_weight_kernel)
__device__ __forceinline__ uint2 ln384_out_plain(const float4& v, float mean, float rstd) {
  uint2 o;
  o.x = pack_bf16x2((v.x - mean) * rstd, (v.y - mean) * rstd);
  o.y = pack_bf16x2((v.z - mean) * rstd, (v.w - mean) * rstd);
  return o;
}

// LayerNorm folded into the Linear layer that consumes it (weight-load time):
//   Linear(LN(x)) = W (xhat * gamma + beta) + b = (W diag(gamma)) xhat + (b + W beta),   xhat = (x - mean) * rstd
// W [N, K] fp32 -> Wf [N, K] bf16 = W * gamma (per input column), bf [N] fp32 = b + W . beta (fp32 dot).  One warp per
// output row.  The kernel that consumes Wf then only has to produce xhat (mlp.cuh: fused LayerNorm2 -> fc1).
__global__ void __launch_bounds__(256)
fold_ln_weight_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ Wf, float* __restrict__ bf, int N, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* w = W + size_t(warp) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float v = w[k];
    Wf[size_t(warp) * K + k] = __float2bfloat16_rn(v * gamma[k]);
    acc = fmaf(v, beta[k], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bf[warp] = b[warp] + acc;
}

// The same for a bf16x3 consumer: Wf = W * gamma split into hi + lo, stored as [N, 2K] = [hi | lo] (head.cuh).
__global__ void __launch_bounds__(256)
fold_split_ln_weight_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ gamma,
                            const float* __restrict__ beta, __nv_bfloat16* __restrict__ Wf, float* __restrict__ bf, int N, int K) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* w = W + size_t(warp) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float v = w[k];
    const float f = v * gamma[k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(f);
    Wf[size_t(warp) * 2 * K + k] = hi;
    Wf[size_t(warp) * 2 * K + K + k] = __float2bfloat16_rn(f - __bfloat162float(hi));
    acc = fmaf(v, beta[k], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bf[warp] = b[warp] + acc;
}

// D = 384, plain bf16 output: four rows per warp (eight lanes per row), 32 rows per 256-thread block
__global__ void __launch_bounds__(256)
layernorm384_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         __nv_bfloat16* __restrict__ y, int M, float eps) {
  const int sub = threadIdx.x & 7;
  const int row = blockIdx.x * 32 + (threadIdx.x >> 3);
  const bool live = row < M;
  const float4* xr = reinterpret_cast<const float4*>(x + size_t(live ? row : M - 1) * 384);
  float4 v[LN384_V];
#pragma unroll
  for (int i = 0; i < LN384_V; ++i) v[i] = xr[i * 8 + sub];
  float mean, rstd;
  ln384_stats(v, eps, mean, rstd);
  if (!live) return;
  uint2* yr = reinterpret_cast<uint2*>(y + size_t(row) * 384);
#pragma unroll
  for (int i = 0; i < LN384_V; ++i)
    yr[i * 8 + sub] = ln384_out(v[i], mean, rstd, __ldg(reinterpret_cast<const float4*>(gamma) + i * 8 + sub),
                                __ldg(reinterpret_cast<const float4*>(beta) + i * 8 + sub));
}

// SPLIT: y is [M, 2*D] = [hi | lo] (bf16x3 operand of the head GEMM, see split_weight_kernel).
template <int D, bool SPLIT>
__global__ void __launch_bounds__(256)
layernorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ y, int M, float eps) {
  static_assert(D % 128 == 0, "D must be a multiple of 128");
  constexpr int V = D / 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + size_t(row) * D);
  float4 v[V];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = xr[i * 32 + lane];
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.0f / D) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + size_t(row) * (SPLIT ? 2 * D : D));
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    const float a = (v[i].x - mean) * rstd * gm.x + bt.x, b = (v[i].y - mean) * rstd * gm.y + bt.y;
    const float c = (v[i].z - mean) * rstd * gm.z + bt.z, d = (v[i].w - mean) * rstd * gm.w + bt.w;
    uint2 o;
    o.x = pack_bf16x2(a, b);
    o.y = pack_bf16x2(c, d);
    yr[i * 32 + lane] = o;
    if constexpr (SPLIT) {
      uint2 l;
      l.x = pack_bf16x2(a - __uint_as_float(o.x << 16), b - __uint_as_float(o.x & 0xffff0000u));
      l.y = pack_bf16x2(c - __uint_as_float(o.y << 16), d - __uint_as_float(o.y & 0xffff0000u));
      yr[D / 4 + i * 32 + lane] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Controller-side reduction on the label map (reference docs/index.html "Controller": the
// potential-field controller steers away from the image half with the most obstacle patches).  Per frame, the number
// of OUTPUT-MAP pixels of every class left (x < W/2) and right (x >= W/2) of the centre line, W = g*p, computed from
// the low-res map (patch (i, j) covers p rows and the columns [j*p, (j+1)*p)): counts[b][side][c], int32.
// One CTA per frame; integer work, exact.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
half_counts_kernel(const uint8_t* __restrict__ lowres, int32_t* __restrict__ counts, int g, int p, int C) {
  __shared__ int hist[2 * 256];
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int b = blockIdx.x;
  const int half = (g * p) / 2;
  const uint8_t* low = lowres + size_t(b) * g * g;
  for (int idx = threadIdx.x; idx < g * g; idx += blockDim.x) {
    const int j = idx % g;
    const int c = low[idx];
    if (c >= C) continue;
    int left = half - j * p;
    left = left < 0 ? 0 : (left > p ? p : left);
    if (left) atomicAdd(&hist[c], left * p);
    if (p - left) atomicAdd(&hist[C + c], (p - left) * p);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) counts[size_t(b) * 2 * C + i] = hist[i];
}

// ---------------------------------------------------------------------------------------
// argmax with torch semantics: first maximum wins, NaN counts as the maximum
// (reference pl_torch_modules.py:295 torch.argmax)
// ---------------------------------------------------------------------------------------
constexpr int HEAD_MAX_C = 16;

// v has HEAD_MAX_C entries (C <= MAXC valid); fully unrolled so that v stays in registers
template <int MAXC = HEAD_MAX_C>
__device__ __forceinline__ int argmax_first(const float (&v)[HEAD_MAX_C], int C) {
  float best = v[0];
  int idx = 0;
#pragma unroll
  for (int c = 1; c < MAXC; ++c) {
    const float x = v[c];
    if (c < C && ((x > best) || (x != x && best == best))) { best = x; idx = c; }
  }
  return idx;
}


// ---------------------------------------------------------------------------------------
// Head tail: layer_3 -> log_softmax -> argmax -> p x p replication   (reference pl_torch_modules.py:121-123 / :135-138,
// :295, :297-298 np.kron), fp32 CUDA cores, HBM-bound (reads 400 B per patch, writes 8*p*p B per patch).
//   LINEAR = false: in = h2 = relu(layer_2(relu(layer_1))) [B*Ntok, ld] fp32 (produced by the two tensor-core GEMMs),
//                   z = in . W3^T + b3;   LINEAR = true: in = the 'linear' head's logits [B*Ntok, ld], z = in[:C].
// Input rows are tokens INCLUDING the cls row of every frame; output rows drop it (reference :243).
// Eight lanes per patch row (four rows per warp per step): the lanes read the row as consecutive float4 (coalesced),
// keep partial dot products for all classes, all-reduce them with three xor-shuffles, and then every lane holds the
// row's logits.  The label is written to the low-res map and - fused, labels != nullptr - replicated into its p x p
// block of the int64 map: the eight lanes take the block's rows, 16-byte stores when p is even (the four patches of
// a warp are neighbours, so every output row receives 4*p*8 contiguous bytes per step).
// ---------------------------------------------------------------------------------------
constexpr int HT_MAX_H2 = 104;

// MAXC: compile-time bound of the class loops (8 or 16): every class slot costs instructions whether it is used or not
// (with MAXC = 16 and the reference's 7 classes the kernel was bound by instruction issue, not by HBM).
template <bool LINEAR, int MAXC>
__global__ void __launch_bounds__(256)
head_tail_kernel(const float* __restrict__ in /*[B*Ntok, ld]*/, int ld, const float* __restrict__ w3 /*[C, H2]*/,
                 const float* __restrict__ b3 /*[C]*/, float* __restrict__ logprobs /*[B*P, C] or null*/,
                 uint8_t* __restrict__ lowres /*[B*P] or null*/, long long* __restrict__ labels /*[B, g*p, g*p] or null*/,
                 int B, int g, int p, int Ntok, int H2, int C) {
  __shared__ float sW3[LINEAR ? 1 : MAXC * HT_MAX_H2];
  __shared__ float sB3[HEAD_MAX_C];
  if constexpr (!LINEAR) {
    for (int i = threadIdx.x; i < MAXC * HT_MAX_H2; i += blockDim.x) {
      const int c = i / HT_MAX_H2, k = i - c * HT_MAX_H2;
      sW3[i] = (c < C && k < H2) ? w3[c * H2 + k] : 0.f;
    }
    if (threadIdx.x < HEAD_MAX_C) sB3[threadIdx.x] = threadIdx.x < C ? b3[threadIdx.x] : 0.f;
    __syncthreads();
  }
  const int P = g * g;
  const int total_rows = B * P;
  const int sub = threadIdx.x & 7;                           // lane within the row's group of eight
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int ngrp = (gridDim.x * blockDim.x) >> 3;
  const int W = g * p;
  for (int r0 = 0; r0 < total_rows; r0 += ngrp) {            // uniform trip count: the shuffles below need full warps
    const int r = r0 + grp;
    const bool live = r < total_rows;
    const int rr = live ? r : total_rows - 1;
    const int b = rr / P, t = rr - b * P;
    const float* src = in + (size_t(b) * Ntok + 1 + t) * ld;
    float z[HEAD_MAX_C];
#pragma unroll
    for (int c = 0; c < HEAD_MAX_C; ++c) z[c] = 0.f;
    if constexpr (LINEAR) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) z[c] = __ldg(src + c);
    } else {
      for (int k4 = sub; k4 < H2 / 4; k4 += 8) {
        const float4 hv = __ldg(reinterpret_cast<const float4*>(src) + k4);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
          if (c < C) {
            const float* wr = sW3 + c * HT_MAX_H2 + 4 * k4;
            z[c] = fmaf(hv.x, wr[0], z[c]); z[c] = fmaf(hv.y, wr[1], z[c]);
            z[c] = fmaf(hv.z, wr[2], z[c]); z[c] = fmaf(hv.w, wr[3], z[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          z[c] += __shfl_xor_sync(0xffffffffu, z[c], 1);
          z[c] += __shfl_xor_sync(0xffffffffu, z[c], 2);
          z[c] += __shfl_xor_sync(0xffffffffu, z[c], 4);
          z[c] += sB3[c];
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) mx = fmaxf(mx, z[c]);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) se += expf(z[c] - mx);
    const float lse = logf(se);
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) z[c] = (z[c] - mx) - lse;
    if (!live) continue;
    if (logprobs != nullptr) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C && (c & 7) == sub) logprobs[size_t(r) * C + c] = z[c];
    }
    const int label = argmax_first<MAXC>(z, C);
    if (lowres != nullptr && sub == 0) lowres[r] = uint8_t(label);
    if (labels != nullptr) {
      const int i = t / g, j = t - i * g;
      long long* blk = labels + (size_t(b) * W + size_t(i) * p) * W + size_t(j) * p;
      if ((p & 1) == 0) {
        const longlong2 v = make_longlong2(label, label);
        for (int yy = sub; yy < p; yy += 8) {
          longlong2* row = reinterpret_cast<longlong2*>(blk + size_t(yy) * W);
          for (int x2 = 0; x2 < (p >> 1); ++x2) row[x2] = v;
        }
      } else {
        for (int yy = sub; yy < p; yy += 8) {
          long long* row = blk + size_t(yy) * W;
          for (int x = 0; x < p; ++x) row[x] = label;
        }
      }
    }
  }
}

// log-probs [rows, C] -> label per row (torch.argmax semantics)
__global__ void argmax_rows_kernel(const float* __restrict__ logprobs, uint8_t* __restrict__ lowres, int rows,
                                   int C) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float z[HEAD_MAX_C];
#pragma unroll
  for (int c = 0; c < HEAD_MAX_C; ++c) z[c] = c < C ? logprobs[size_t(r) * C + c] : 0.f;
  lowres[r] = uint8_t(argmax_first(z, C));
}

// ---------------------------------------------------------------------------------------
// CLS-query attention of one block: out[b, h, j] = softmax_j( q_cls(b,h) . k_j(b,h) )
// (q as the qkv GEMM wrote it, pre-scaled by dh^-0.5 * log2(e): the scores are in log2 units, attention.cuh)
// = row 0 of the attention matrix VisionTransformer.get_last_selfattention returns (reference
// vision_transformer.py:273-280, :85-101), the only row its caller uses (visualize_attention.py:46-54).
// qkv: [B, N, 3D] bf16 as written by the qkv GEMM.  One CTA per (b, h); N x 64 dot products are nothing next to
// the hot path, so this is a plain block-reduction kernel.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cls_attention_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ out, int N, int D, int H) {
  const int h = blockIdx.x, b = blockIdx.y;
  const __nv_bfloat16* base = qkv + size_t(b) * N * 3 * D;
  __shared__ float sq[64];
  __shared__ float red[8];
  __shared__ float bcast;
  if (threadIdx.x < 64) sq[threadIdx.x] = __bfloat162float(base[h * 64 + threadIdx.x]);   // token 0 = cls, q part
  __syncthreads();
  float* o = out + (size_t(b) * H + h) * N;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const uint4* kp = reinterpret_cast<const uint4*>(base + size_t(j) * 3 * D + D + h * 64);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 v = __ldg(kp + c);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc = fmaf(sq[c * 8 + 2 * e], __uint_as_float(w[e] << 16), acc);
        acc = fmaf(sq[c * 8 + 2 * e + 1], __uint_as_float(w[e] & 0xffff0000u), acc);
      }
    }
    o[j] = acc;
    mx = fmaxf(mx, acc);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const float e = exp2f(o[j] - mx);
    o[j] = e;
    sum += e;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int j = threadIdx.x; j < N; j += blockDim.x) o[j] *= inv;
}

// ---------------------------------------------------------------------------------------
// Nearest-neighbour block replication == np.kron(low_res, ones((p,p), int))
// (reference pl_torch_modules.py:297-298): out[b][y][x] = low[b][y/p][x/p], int64.
// ---------------------------------------------------------------------------------------
__global__ void replicate_labels_kernel(const uint8_t* __restrict__ lowres /*[B,g,g]*/,
                                        long long* __restrict__ out /*[B,g*p,g*p]*/, int B, int g, int p) {
  const int W = g * p;
  if ((W & 1) == 0) {
    const int W2 = W >> 1;
    const size_t total = size_t(B) * W * W2;
    size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x2 = int(idx % W2);
    const size_t t = idx / W2;
    const int y = int(t % W);
    const int b = int(t / W);
    const uint8_t* lr = lowres + (size_t(b) * g + y / p) * g;
    longlong2 v;
    v.x = lr[(2 * x2) / p];
    v.y = lr[(2 * x2 + 1) / p];
    reinterpret_cast<longlong2*>(out)[idx] = v;
  } else {
    const size_t total = size_t(B) * W * W;
    size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = int(idx % W);
    const size_t t = idx / W;
    const int y = int(t % W);
    const int b = int(t / W);
    out[idx] = lowres[(size_t(b) * g + y / p) * g + x / p];
  }
}

}  // namespace dsg
