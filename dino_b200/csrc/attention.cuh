// Flash-style fused multi-head self-attention forward for the DINO ViT blocks
// (reference: vision_transformer.py:80-107 Attention.forward, no mask, dropout p=0).
//
//   S = (q*dh^-0.5) k^T ; P = softmax_rows(S) ; O = P v          per (frame b, head h)
//
// The reference materialises S and P ([B,H,N,N] fp32, 311 MB/frame/block at 480 px); here a
// CTA owns one 128-query tile of one (b,h), streams 128-key K/V tiles through shared memory
// with TMA, keeps S (128x128 fp32), P (bf16, aliasing S) and the running O (128x64 fp32) in
// TMEM, and never writes N x N data anywhere.
//
// q/k/v are read straight out of the qkv GEMM output [B, N, 3*D] bf16 (column =
// which*D + h*64 + d, exactly nn.Linear's output order, reference :82) through ONE 3-D tensor
// map {3D, N, B}; rows >= N of a frame are zero-filled by TMA.  q is pre-scaled by dh^-0.5
// (exact in bf16: 0.125) by the qkv epilogue.
//
// Warp roles (192 threads, 2 CTAs resident per SM so one CTA's softmax overlaps the other's MMAs):
//   warp 0     : TMA producer (Q once; K,V ring)
//   warp 1     : TMEM alloc + tcgen05.mma issuer:  S = Q K^T (SS, K-major both) ; O += P V (TS: P from
//                TMEM, V as an MN-major smem operand — no transpose of V is ever made)
//   warps 2..5 : online softmax, one query row per thread (tcgen05.ld 32x32b), O rescale, epilogue
#pragma once
#include "ptx.cuh"

namespace dsg {

struct AttnParams {
  int B, H, N;          // frames, heads, tokens per frame
  int D;                // embed dim (= H*64)
  __nv_bfloat16* out;   // [B*N, D], column = h*64 + d   (reference :104 transpose(1,2).reshape)
};

constexpr int ATT_BM = 128;     // queries per CTA
constexpr int ATT_BN = 128;     // keys per tile
constexpr int ATT_DH = 64;      // head dim
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * ATT_DH * 2;  // 16 KB (Q, K or V tile)

template <int KV_STAGES>
constexpr size_t attn_smem_bytes() {
  return size_t(1 + 2 * KV_STAGES) * ATT_TILE_BYTES + 1024 + 256;
}

template <int KV_STAGES>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  constexpr uint32_t TMEM_COLS = 256;
  constexpr uint32_t S_COL = 0;     // S: 128 fp32 columns; P (bf16x2) aliases columns [0,64)
  constexpr uint32_t O_COL = 128;   // O: 64 fp32 columns
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + ATT_TILE_BYTES;  // [stage][K 16KB | V 16KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(1 + 2 * KV_STAGES) * ATT_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int num_tiles = (p.N + ATT_BN - 1) / ATT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, h * ATT_DH, q0, b);
      for (int j = 0; j < num_tiles; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_expect_tx(&kv_full[s], 2 * ATT_TILE_BYTES);
        uint8_t* sk = sKV + size_t(s) * 2 * ATT_TILE_BYTES;
        tma_load_3d(sk, &tmQKV, &kv_full[s], p.D + h * ATT_DH, j * ATT_BN, b);
        tma_load_3d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[s], 2 * p.D + h * ATT_DH, j * ATT_BN, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0);  // K^T: K-major B
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_DH, 1);  // V  : MN-major B
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      for (int j = 0; j < num_tiles; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint32_t sk = smem_u32(sKV + size_t(s) * 2 * ATT_TILE_BYTES);
        const uint64_t kdesc = umma_desc_sw128(sk);
        const uint64_t vdesc = umma_desc_sw128(sk + ATT_TILE_BYTES);
        // S = Q K^T : 4 x (128x128x16).  Tensor-pipe ordering guarantees this does not overwrite
        // P(j-1) before the previously issued P(j-1) V(j-1) has read it.
#pragma unroll
        for (int k = 0; k < ATT_DH / 16; ++k)
          umma_ss(tmem_base + S_COL, qdesc + uint64_t(k * 2), kdesc + uint64_t(k * 2), idesc_qk, k != 0);
        tc_commit(s_full);
        // wait for the softmax warps to publish P(j) (and the rescaled O)
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        // O += P V : 8 x (128x64x16); P: 8 TMEM columns per step; V: 16 keys = 2 KB per step
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k)
          umma_ts(tmem_base + O_COL, tmem_base + S_COL + uint32_t(k * 8), vdesc + uint64_t(k * 128), idesc_pv,
                  (j | k) != 0);
        tc_commit(&kv_empty[s]);
      }
      tc_commit(o_full);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
    float m = -INFINITY;  // running row max of the raw (pre-log2e) scores
    float l = 0.f;        // running row sum

    for (int j = 0; j < num_tiles; ++j) {
      mbar_wait(s_full, j & 1);  // also implies P(j-1) V(j-1) retired: O is stable
      tc_fence_after();
      float s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_x32(lane_base + S_COL + uint32_t(c * 32), r);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(r[i]);
      }
      tmem_ld_wait();
      const int kbase = j * ATT_BN;
      if (kbase + ATT_BN > p.N) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kbase + i >= p.N) s[i] = -INFINITY;
      }
      float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        mx0 = fmaxf(mx0, s[i]); mx1 = fmaxf(mx1, s[i + 1]);
        mx2 = fmaxf(mx2, s[i + 2]); mx3 = fmaxf(mx3, s[i + 3]);
      }
      const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
      const float alpha = fast_exp2((m - m_new) * LOG2E);  // 0 on the first tile (m = -inf)
      if (j > 0 && __any_sync(0xffffffffu, m_new > m)) {
        // rescale the running O by alpha (per row)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_x32(lane_base + O_COL + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x32(lane_base + O_COL + uint32_t(c * 32), r);
        }
      }
      const float mb = m_new * LOG2E;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e0 = fast_exp2(fmaf(s[c * 64 + 2 * i], LOG2E, -mb));
          const float e1 = fast_exp2(fmaf(s[c * 64 + 2 * i + 1], LOG2E, -mb));
          if (i & 1) { sum2 += e0; sum3 += e1; } else { sum0 += e0; sum1 += e1; }
          pk[i] = pack_bf16x2(e0, e1);
        }
        tmem_st_x32(lane_base + S_COL + uint32_t(c * 32), pk);
      }
      l = l * alpha + ((sum0 + sum1) + (sum2 + sum3));
      m = m_new;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // epilogue: O / l -> bf16 -> out[b*N + q, h*64 + d]
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int q = q0 + row;
    __nv_bfloat16* o = p.out + (size_t(b) * p.N + q) * p.D + h * ATT_DH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld_x32(lane_base + O_COL + uint32_t(c * 32), r);
      tmem_ld_wait();
      if (q < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          v.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          v.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          v.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(o + c * 32 + i) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace dsg
