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

// W2 [H2, H1] (nn.Linear layout) -> W2t [H1, H2P] fp32, zero padded columns
__global__ void transpose_pad_kernel(const float* __restrict__ in, float* __restrict__ out, int H2, int H1,
                                     int H2P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H1 * H2P) return;
  const int k = i / H2P, n = i - k * H2P;
  out[i] = n < H2 ? in[size_t(n) * H1 + k] : 0.f;
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
//   A[b*P + i*g + j][c*64 + ky*8 + kx] = bf16(frame[b][c][i*8+ky][j*8+kx])
// One thread moves two image rows of one patch (2 x 32 B in, one 32 B chunk out).
// ---------------------------------------------------------------------------------------
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
  uint4 o0, o1;
  o0.x = pack_bf16x2(a0.x, a0.y); o0.y = pack_bf16x2(a0.z, a0.w);
  o0.z = pack_bf16x2(a1.x, a1.y); o0.w = pack_bf16x2(a1.z, a1.w);
  o1.x = pack_bf16x2(b0.x, b0.y); o1.y = pack_bf16x2(b0.z, b0.w);
  o1.z = pack_bf16x2(b1.x, b1.y); o1.w = pack_bf16x2(b1.z, b1.w);
  __nv_bfloat16* dst = A + (size_t(b) * g * g + size_t(i) * g + j) * 192 + c * 64 + kyp * 16;
  *reinterpret_cast<uint4*>(dst) = o0;
  *reinterpret_cast<uint4*>(dst + 8) = o1;
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
template <int D>
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
  uint2* yr = reinterpret_cast<uint2*>(y + size_t(row) * D);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    uint2 o;
    o.x = pack_bf16x2((v[i].x - mean) * rstd * gm.x + bt.x, (v[i].y - mean) * rstd * gm.y + bt.y);
    o.y = pack_bf16x2((v[i].z - mean) * rstd * gm.z + bt.z, (v[i].w - mean) * rstd * gm.w + bt.w);
    yr[i * 32 + lane] = o;
  }
}

// ---------------------------------------------------------------------------------------
// argmax with torch semantics: first maximum wins, NaN counts as the maximum
// (reference pl_torch_modules.py:295 torch.argmax)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int argmax_first(const float* v, int C) {
  float best = v[0];
  int idx = 0;
  for (int c = 1; c < C; ++c) {
    const float x = v[c];
    if ((x > best) || (x != x && best == best)) { best = x; idx = c; }
  }
  return idx;
}

constexpr int HEAD_MAX_C = 16;

// ---------------------------------------------------------------------------------------
// Head tail: h1 (= relu(layer_1), fp32, produced by the tensor-core GEMM) -> layer_2 -> relu ->
// layer_3 -> log_softmax -> argmax   (reference pl_torch_modules.py:119-123, :295), fp32 CUDA cores.
// Input rows are tokens INCLUDING the cls row of every frame; output rows drop it (reference :243).
// Block = 256 threads = 64 patch rows; W2^T (padded) and the h1 tile live in shared memory.
// ---------------------------------------------------------------------------------------
constexpr int HT_ROWS = 64;
constexpr int HT_H2P = 112;  // layer-2 width padded to 7 x 16

__host__ __device__ constexpr size_t head_tail_smem_bytes(int H1) {
  return size_t(H1) * HT_H2P * 4 + size_t(HT_ROWS) * (H1 + 1) * 4 + HEAD_MAX_C * 104 * 4 + (HT_H2P + HEAD_MAX_C) * 4;
}

__global__ void __launch_bounds__(256, 1)
head_tail_kernel(const float* __restrict__ h1 /*[B*Ntok, H1]*/, const float* __restrict__ w2t /*[H1, HT_H2P]*/,
                 const float* __restrict__ b2 /*[H2]*/, const float* __restrict__ w3 /*[C, H2]*/,
                 const float* __restrict__ b3 /*[C]*/, float* __restrict__ logprobs /*[B*P, C] or null*/,
                 uint8_t* __restrict__ lowres /*[B*P] or null*/, int B, int P, int Ntok, int H1, int H2, int C) {
  extern __shared__ float sm[];
  float* sW2 = sm;                              // [H1][HT_H2P]
  float* sH = sW2 + size_t(H1) * HT_H2P;        // [64][H1+1]  (later reused as h2 [64][105])
  float* sW3 = sH + size_t(HT_ROWS) * (H1 + 1); // [C][104]
  float* sB2 = sW3 + HEAD_MAX_C * 104;          // [HT_H2P]
  float* sB3 = sB2 + HT_H2P;                    // [HEAD_MAX_C]
  const int tid = threadIdx.x;
  const int ldh = H1 + 1;
  const int total_rows = B * P;
  const int ntiles = (total_rows + HT_ROWS - 1) / HT_ROWS;

  // weights once per (persistent) block
  for (int i = tid; i < H1 * HT_H2P; i += 256) sW2[i] = w2t[i];
  for (int i = tid; i < C * H2; i += 256) sW3[(i / H2) * 104 + (i % H2)] = w3[i];
  for (int i = tid; i < HT_H2P; i += 256) sB2[i] = i < H2 ? b2[i] : 0.f;
  if (tid < C) sB3[tid] = b3[tid];

  const int ty = tid >> 4, tx = tid & 15;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * HT_ROWS;
    __syncthreads();  // previous tile's readers of sH are done (and weights are visible)
    for (int i = tid; i < HT_ROWS * H1; i += 256) {
      const int rr = i / H1, k = i - rr * H1;
      const int r = r0 + rr;
      float v = 0.f;
      if (r < total_rows) {
        const int b = r / P, t = r - b * P;
        v = h1[(size_t(b) * Ntok + 1 + t) * H1 + k];
      }
      sH[rr * ldh + k] = v;
    }
    __syncthreads();

    // layer 2: each thread -> 4 rows x 7 columns (tx + 16c)
    float acc[4][7];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 7; ++c) acc[a][c] = 0.f;
    for (int k = 0; k < H1; ++k) {
      float a[4], w[7];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sH[(ty * 4 + i) * ldh + k];
#pragma unroll
      for (int c = 0; c < 7; ++c) w[c] = sW2[k * HT_H2P + tx + 16 * c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[i][c] = fmaf(a[i], w[c], acc[i][c]);
    }
    __syncthreads();  // everyone is done reading sH (h1)
    float* sH2 = sH;  // [64][105]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 7; ++c) {
        const int n = tx + 16 * c;
        if (n < 104) sH2[(ty * 4 + i) * 105 + n] = fmaxf(acc[i][c] + sB2[n], 0.f);
      }
    __syncthreads();

    // layer 3 + log_softmax + argmax: one thread per row
    if (tid < HT_ROWS) {
      const int r = r0 + tid;
      if (r < total_rows) {
        float z[HEAD_MAX_C];
#pragma unroll
        for (int c = 0; c < HEAD_MAX_C; ++c) z[c] = 0.f;
        for (int k = 0; k < H2; ++k) {
          const float hv = sH2[tid * 105 + k];
#pragma unroll
          for (int c = 0; c < HEAD_MAX_C; ++c)
            if (c < C) z[c] = fmaf(hv, sW3[c * 104 + k], z[c]);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < HEAD_MAX_C; ++c)
          if (c < C) { z[c] += sB3[c]; mx = fmaxf(mx, z[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < HEAD_MAX_C; ++c)
          if (c < C) se += expf(z[c] - mx);
        const float lse = logf(se);
#pragma unroll
        for (int c = 0; c < HEAD_MAX_C; ++c)
          if (c < C) z[c] = (z[c] - mx) - lse;
        if (logprobs != nullptr) {
#pragma unroll
          for (int c = 0; c < HEAD_MAX_C; ++c)
            if (c < C) logprobs[size_t(r) * C + c] = z[c];
        }
        if (lowres != nullptr) lowres[r] = uint8_t(argmax_first(z, C));
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
  for (int c = 0; c < C; ++c) z[c] = logprobs[size_t(r) * C + c];
  lowres[r] = uint8_t(argmax_first(z, C));
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
