// Dense "TN" GEMM on tcgen05/TMEM fed by TMA:   C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// A is a row-major bf16 activation matrix, W is an nn.Linear weight in its native [out,in]
// layout (both K-major UMMA operands, so no transposes anywhere).  This one kernel, with
// different epilogues, carries every linear layer of the DINOSeg hot path:
//   patch-embed (reference vision_transformer.py:153-157 as an im2col GEMM), qkv (:75,:82),
//   proj (:77,:105), fc1/GELU (:54-55,:60-61), fc2 (:56,:63), head layer_1 (pl_torch_modules.py:113,118).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      : TMA producer  (A tile 128x64, W tile BNx64 per k-block, SWIZZLE_128B, STAGES-deep ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (accumulator 128 lanes x BN fp32 columns)
//   warps 2..5  : epilogue: tcgen05.ld (32x32b: one accumulator row per thread) -> bias/activation -> global
#pragma once
#include "ptx.cuh"

namespace dsg {

enum : int {
  EPI_BF16 = 0,       // out bf16 = (acc + bias) * (col < scale_cols ? col_scale : 1)
  EPI_GELU_BF16 = 1,  // out bf16 = gelu_erf(acc + bias)
  EPI_RESID_F32 = 2,  // out f32 += acc + bias        (in-place residual add)
  EPI_PATCH_F32 = 3,  // out f32[(r/P)*Ntok + 1 + r%P] = acc + bias + pos[1 + r%P]
  EPI_RELU_F32 = 4,   // out f32 = relu(acc + bias)
};

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] (may be null)
  void* out;
  int ldo;  // leading dimension of out, in elements
  float col_scale;
  int scale_cols;
  const float* pos;  // EPI_PATCH_F32: [Ntok, N] positional table (row 0 = cls)
  int P, Ntok;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

template <int BN, int STAGES>
constexpr size_t gemm_smem_bytes() {
  return size_t(STAGES) * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024 /*align slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN, int EPI, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const GemmParams p) {
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* sa = smem + size_t(s) * STAGE_BYTES;
        tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0);
        tma_load_2d(sa + A_BYTES, &tmW, &full_bar[s], kb * GEMM_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + size_t(s) * STAGE_BYTES);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >>4 -> +2)
          umma_ss(tmem_base, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb | k) != 0);
        }
        tc_commit(&empty_bar[s]);
      }
      tc_commit(acc_bar);
    }
  } else {
    // ---------------- epilogue: thread <-> accumulator row ----------------
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int grow = m0 + row;
    const bool row_ok = grow < p.M;
    mbar_wait(acc_bar, 0);
    tc_fence_after();

    size_t out_row = size_t(grow);
    const float* pos_row = nullptr;
    if constexpr (EPI == EPI_PATCH_F32) {
      const int b = grow / p.P;
      const int t = grow - b * p.P;
      out_row = size_t(b) * p.Ntok + 1 + t;
      pos_row = p.pos + size_t(1 + t) * p.N;
    }

#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the predicated stores
      tmem_ld_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(c * 32), r);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (col0 + i < p.N) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
          }
        }
      }
      if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + col0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          if (col0 + i < p.N) {
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = v[i + j];
              if constexpr (EPI == EPI_GELU_BF16) x = gelu_erf(x);
              else if (col0 + i + j < p.scale_cols) x *= p.col_scale;
              w[j] = x;
            }
            uint4 q;
            q.x = pack_bf16x2(w[0], w[1]); q.y = pack_bf16x2(w[2], w[3]);
            q.z = pack_bf16x2(w[4], w[5]); q.w = pack_bf16x2(w[6], w[7]);
            *reinterpret_cast<uint4*>(o + i) = q;
          }
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + col0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (col0 + i < p.N) {
            float4 q = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if constexpr (EPI == EPI_RESID_F32) {
              const float4 x = *reinterpret_cast<const float4*>(o + i);
              q.x += x.x; q.y += x.y; q.z += x.z; q.w += x.w;
            } else if constexpr (EPI == EPI_PATCH_F32) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(pos_row + col0 + i));
              q.x += x.x; q.y += x.y; q.z += x.z; q.w += x.w;
            } else if constexpr (EPI == EPI_RELU_F32) {
              q.x = fmaxf(q.x, 0.f); q.y = fmaxf(q.y, 0.f); q.z = fmaxf(q.z, 0.f); q.w = fmaxf(q.w, 0.f);
            }
            *reinterpret_cast<float4*>(o + i) = q;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace dsg
