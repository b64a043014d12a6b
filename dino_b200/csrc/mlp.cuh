// Fused transformer MLP for D = 384 (ViT-S):   x += fc2( gelu( fc1( LayerNorm2(x) ) ) )
// (reference vision_transformer.py:135 -> :59-65 Mlp.forward with :118 norm2; nn.GELU() = exact erf)
//
// One persistent CTA per SM walks 128-token row blocks.  Nothing of the 1536-wide hidden activation ever
// leaves the SM (the unfused path writes and re-reads 708 MB of it per block at batch 64), and the
// hidden bias / GELU / residual passes are fused:
//
//   A block     : LayerNorm2(x) as bf16 (written by the LayerNorm kernel) is TMA-loaded once per row block and
//                 stays resident in shared memory (6 k-blocks x 16 KB); the next block's A is requested as soon
//                 as the last MMA1 of the current block has been issued
//   per hidden chunk c of 128 columns (12 chunks):
//     MMA1(c)   : acc1[128 x 128] = A . W1[c]^T              (24 MMAs 128x128x16, W1 granules via TMA)
//     GELU warps: acc1 -> registers -> + b1 -> GELU -> bf16 -> shared memory G (again an A-operand layout)
//     MMA2(c)   : acc2[128 x 384] += G . W2[:, c]^T          (3 x 8 MMAs 128x128x16, W2 granules via TMA)
//   output warps: acc2 + b2 + x -> x   (x chunks TMA-loaded into staging, added in place, TMA-stored), running
//                 concurrently with the first chunks of the next row block
//
// TMEM: acc2 = 384 fp32 columns, acc1 = 128 columns -> exactly the 512 columns of an SM.
// The tensor pipe runs  MMA1(0) | MMA1(1) MMA2(0) | MMA1(2) MMA2(1) | ...  so the GELU of chunk c hides
// under MMA1(c+1).  All weight tiles are 16 KB "granules" ([128 rows x 64 k] bf16) that stream through
// one 4-deep TMA ring in exactly the order the MMAs consume them.
//
// Shared memory: A 96 KB + G 32 KB + output staging 32 KB + ring 4 x 16 KB = 224 KB (+ barriers).  The ring depth
// (bytes in flight against the L2 latency) is what bounds the kernel: it is as deep as shared memory allows.
#pragma once
#include "gemm.cuh"

namespace dsg {

struct MlpParams {
  int M;                      // token rows
  const float* x;             // [M, 384] fp32 residual stream (read by the output warps; written through tmX)
  const float* b1;            // [1536]
  const float* b2;            // [384]
  long long* timing;          // debug (DSG_MLP_TIMING): [grid][2 roles][8] cycle totals
};

#ifdef DSG_MLP_TIMING
#define MLP_T(i) do { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } while (0)
#define MLP_T_DECL long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tprev = clock64()
#define MLP_T_DUMP(role) do { if (p.timing) for (int _i = 0; _i < 8; ++_i) \
    p.timing[(size_t(blockIdx.x) * 2 + (role)) * 8 + _i] = tacc[_i]; } while (0)
#else
#define MLP_T(i) do { } while (0)
#define MLP_T_DECL do { } while (0)
#define MLP_T_DUMP(role) do { } while (0)
#endif

constexpr int MLP_D = 384;
constexpr int MLP_HID = 1536;
constexpr int MLP_BM = 128;
constexpr int MLP_CH = 128;                         // hidden columns per chunk
constexpr int MLP_NCH = MLP_HID / MLP_CH;           // 12
constexpr int MLP_KB = MLP_D / 64;                  // 6 k-blocks of A / W1 granules per chunk
constexpr int MLP_GRAN = 128 * 64 * 2;              // 16 KB
constexpr int MLP_RING = 4;
constexpr int MLP_LAG = 1;                          // MMA2(c) is issued right after MMA1(c + MLP_LAG) (a lag of 2 measured 25 % slower: single G buffer)
constexpr int MLP_THREADS = 64 + 256 + 128;         // producer + MMA, 8 GELU warps, 4 output warps
constexpr size_t MLP_SMEM = size_t(MLP_KB) * MLP_GRAN + 2 * MLP_GRAN + 2 * MLP_GRAN + size_t(MLP_RING) * MLP_GRAN + 1024 + 512;

// gelu(x) = 0.5 x + |x| (0.5 - Phi(-|x|)),  Phi(-|x|) = 2^-(q(|x|) + 1)   (same polynomial q as gelu_erf, evaluated
// for two values at once with packed fp32x2 FMAs: 9 FMA-pipe + 4 ALU + 2 MUFU instructions per pair)
__device__ __forceinline__ float2 gelu_erf_x2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 cx = make_float2(fminf(ax.x, 4.949747468f), fminf(ax.y, 4.949747468f));
  float2 q = ffma2(make_float2(-3.0103274184511974e-05f, -3.0103274184511974e-05f), cx,
                   make_float2(0.0007183064590208232f, 0.0007183064590208232f));
  q = ffma2(q, cx, make_float2(-0.007799314800649881f, -0.007799314800649881f));
  q = ffma2(q, cx, make_float2(0.05274621397256851f, 0.05274621397256851f));
  q = ffma2(q, cx, make_float2(0.45945441722869873f, 0.45945441722869873f));
  q = ffma2(q, cx, make_float2(1.150948166847229f, 1.150948166847229f));
  q = ffma2(q, cx, make_float2(1.0000104402643046f, 1.0000104402643046f));      // + 1: the factor 0.5 of Phi
  const float2 w = make_float2(fast_exp2(-q.x), fast_exp2(-q.y));              // Phi(-|x|)
  const float2 d = ffma2(w, make_float2(-1.f, -1.f), make_float2(0.5f, 0.5f));
  const float2 hx = ffma2(x, make_float2(0.5f, 0.5f), make_float2(0.f, 0.f));
  return ffma2(ax, d, hx);
}

__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX, const MlpParams p) {
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ACC2_COL = 0;                  // 384 columns
  constexpr uint32_t ACC1_COL = 384;                // 128 columns

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;                               // [6][16 KB]  LN2(x) bf16 row block (TMA), K-major SW128
  uint8_t* sG = sA + size_t(MLP_KB) * MLP_GRAN;     // [2 k-blocks][16 KB]  gelu(fc1) chunk, A operand of fc2
  uint8_t* sS = sG + 2 * MLP_GRAN;                  // [2][16 KB]  staging of the output warps
  uint8_t* sR = sS + 2 * MLP_GRAN;                  // [4][16 KB]  weight granule ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + size_t(MLP_RING) * MLP_GRAN);
  uint64_t* w_full = bars;                          // MLP_RING
  uint64_t* w_empty = w_full + MLP_RING;            // MLP_RING
  uint64_t* a_full = w_empty + MLP_RING;            // 1 (TMA transaction barrier: 6 granules)
  uint64_t* a_empty = a_full + 1;                   // 1 (commit after the last MMA1 of a row block)
  uint64_t* acc1_full = a_empty + 1;                // 1 (commit)
  uint64_t* acc1_empty = acc1_full + 1;             // 1 (8 arrivals)
  uint64_t* g_full = acc1_empty + 1;                // 1 (256 arrivals)
  uint64_t* g_empty = g_full + 1;                   // 1 (commit after MMA2 of the chunk)
  uint64_t* acc2_full = g_empty + 1;                // 1 (commit after the last MMA2)
  uint64_t* acc2_empty = acc2_full + 1;             // 1 (4 arrivals)
  uint64_t* add_bar = acc2_empty + 1;               // 2 (x chunks landed in staging)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(add_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.M + MLP_BM - 1) / MLP_BM;
  const int my_blocks = int(blockIdx.x) < m_blocks ? (m_blocks - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < MLP_RING; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 8);
    mbar_init(g_full, 256);
    mbar_init(g_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 4);
    for (int s = 0; s < 2; ++s) mbar_init(&add_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      // per row block: A (6 granules, own buffer) | W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W2(11)  through the ring
      uint32_t rc = 0;
      auto load_w1 = [&](int c) {
        for (int kb = 0; kb < MLP_KB; ++kb, ++rc) {
          const int s = rc % MLP_RING;
          mbar_wait(&w_empty[s], ((rc / MLP_RING) & 1) ^ 1);
          mbar_expect_tx(&w_full[s], MLP_GRAN);
          tma_load_2d(sR + size_t(s) * MLP_GRAN, &tmW1, &w_full[s], kb * 64, c * MLP_CH);   // W1[c*128.., kb*64..]
        }
      };
      auto load_w2 = [&](int c) {
        for (int kb = 0; kb < MLP_CH / 64; ++kb) {
          for (int nt = 0; nt < MLP_D / 128; ++nt, ++rc) {
            const int s = rc % MLP_RING;
            mbar_wait(&w_empty[s], ((rc / MLP_RING) & 1) ^ 1);
            mbar_expect_tx(&w_full[s], MLP_GRAN);
            tma_load_2d(sR + size_t(s) * MLP_GRAN, &tmW2, &w_full[s], c * MLP_CH + kb * 64, nt * 128);
          }
        }
      };
      auto load_a = [&](int bi) {
        const int r0 = (int(blockIdx.x) + bi * int(gridDim.x)) * MLP_BM;
        if (bi > 0) mbar_wait(a_empty, (bi - 1) & 1);   // last MMA1 of the previous block has read A
        mbar_expect_tx(a_full, MLP_KB * MLP_GRAN);
        for (int kb = 0; kb < MLP_KB; ++kb) tma_load_3d(sA + size_t(kb) * MLP_GRAN, &tmA, a_full, kb * 64, r0, 0);
      };
      if (my_blocks > 0) load_a(0);
      for (int bi = 0; bi < my_blocks; ++bi) {
        for (int i = 0; i < MLP_NCH + MLP_LAG; ++i) {
          if (i < MLP_NCH) load_w1(i);
          // the A block of the next row block: right after the weights of the last MMA1 have been requested
          if (i == MLP_NCH - 1 && bi + 1 < my_blocks) load_a(bi + 1);
          if (i >= MLP_LAG) load_w2(i - MLP_LAG);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0);
      uint32_t rc = 0;        // ring counter (same order as the producer)
      uint32_t c1 = 0;        // acc1 uses so far
      uint32_t gc = 0;        // G uses so far
      MLP_T_DECL;
      auto wait_gran = [&]() -> uint32_t {
        const int s = rc % MLP_RING;
        MLP_T(7);
        mbar_wait(&w_full[s], (rc / MLP_RING) & 1);
        MLP_T(0);
        tc_fence_after();
        return smem_u32(sR + size_t(s) * MLP_GRAN);
      };
      auto free_gran = [&]() {
        tc_commit(&w_empty[rc % MLP_RING]);
        ++rc;
      };
      auto mma1 = [&](int c) {
        // acc1 = A . W1[c]^T ; the GELU warps must have pulled the previous acc1 into registers
        MLP_T(7);
        mbar_wait(acc1_empty, (c1 & 1) ^ 1);
        MLP_T(1);
        tc_fence_after();
        for (int kb = 0; kb < MLP_KB; ++kb) {
          const uint32_t sw = wait_gran();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + size_t(kb) * MLP_GRAN));
          const uint64_t bdesc = umma_desc_sw128(sw);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tmem_base + ACC1_COL, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, (kb | k) != 0);
          free_gran();
        }
        tc_commit(acc1_full);
        ++c1;
        if (c == MLP_NCH - 1) tc_commit(a_empty);   // the next row block's A may be loaded
      };
      auto mma2 = [&](int c, bool first_of_block) {
        // acc2 += G . W2[:, c]^T
        MLP_T(7);
        mbar_wait(g_full, gc & 1);
        MLP_T(2);
        tc_fence_after();
        for (int kb = 0; kb < MLP_CH / 64; ++kb) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sG + size_t(kb) * MLP_GRAN));
          for (int nt = 0; nt < MLP_D / 128; ++nt) {
            const uint32_t sw = wait_gran();
            const uint64_t bdesc = umma_desc_sw128(sw);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss(tmem_base + ACC2_COL + uint32_t(nt * 128), adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc,
                      (!first_of_block || kb != 0 || k != 0) ? 1u : 0u);
            free_gran();
          }
        }
        tc_commit(g_empty);
        ++gc;
        (void)c;
      };
      for (int bi = 0; bi < my_blocks; ++bi) {
        MLP_T(7);
        mbar_wait(a_full, bi & 1);                  // LN2(x) row block has landed
        MLP_T(3);
        tc_fence_after();
        for (int i = 0; i < MLP_NCH + MLP_LAG; ++i) {
          if (i < MLP_NCH) mma1(i);
          if (i >= MLP_LAG) {
            const int c = i - MLP_LAG;
            if (c == 0) {
              MLP_T(7);
              mbar_wait(acc2_empty, (bi & 1) ^ 1);  // the output warps have drained the previous block's acc2
              MLP_T(4);
              tc_fence_after();
            }
            mma2(c, c == 0);
          }
        }
        tc_commit(acc2_full);
      }
      MLP_T_DUMP(0);
    }
  } else if (warp < 10) {
    // ---------------- GELU warps (8): acc1 -> + b1 -> GELU -> bf16 -> G ----------------
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;               // 64-column half of the chunk (= k-block of G)
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
    uint32_t cc = 0;                                // chunks so far (phases of acc1 / G barriers)
    MLP_T_DECL;
    for (int bi = 0; bi < my_blocks; ++bi) {
      for (int c = 0; c < MLP_NCH; ++c, ++cc) {
        MLP_T(7);
        mbar_wait(acc1_full, cc & 1);
        MLP_T(1);
        tc_fence_after();
        float v[64];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t r[32];
          tmem_ld_x32(lane_base + ACC1_COL + uint32_t(half * 64 + h2 * 32), r);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[h2 * 32 + i] = __uint_as_float(r[i]);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc1_empty);     // MMA1 of the next chunk may overwrite acc1
        const float* b1 = p.b1 + c * MLP_CH + half * 64;
        uint4 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(b1 + 8 * k));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + 8 * k + 4));
          const float2 g0 = gelu_erf_x2(make_float2(v[8 * k + 0] + ba.x, v[8 * k + 1] + ba.y));
          const float2 g1 = gelu_erf_x2(make_float2(v[8 * k + 2] + ba.z, v[8 * k + 3] + ba.w));
          const float2 g2 = gelu_erf_x2(make_float2(v[8 * k + 4] + bb.x, v[8 * k + 5] + bb.y));
          const float2 g3 = gelu_erf_x2(make_float2(v[8 * k + 6] + bb.z, v[8 * k + 7] + bb.w));
          q[k].x = pack_bf16x2(g0.x, g0.y); q[k].y = pack_bf16x2(g1.x, g1.y);
          q[k].z = pack_bf16x2(g2.x, g2.y); q[k].w = pack_bf16x2(g3.x, g3.y);
        }
        // the single G buffer was last read by MMA2 of the previous chunk
        MLP_T(2);
        mbar_wait(g_empty, (cc & 1) ^ 1);
        MLP_T(3);
        uint8_t* dst = sG + size_t(half) * MLP_GRAN + row * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(dst + ((k ^ (row & 7)) << 4)) = q[k];
        fence_proxy_async_smem();
        mbar_arrive(g_full);
        MLP_T(4);
      }
    }
    if (threadIdx.x == 64) { MLP_T_DUMP(1); }
  } else {
    // ---------------- output warps (4): x += acc2 + b2, 12 chunks of 32 fp32 columns ----------------
    // x chunks are TMA-loaded into a swizzled staging buffer (one chunk ahead, the first one long before acc2
    // is complete), the accumulator + bias is added in place and the buffer is TMA-stored back.
    // (Reading x with plain loads, one row per thread, measured 25 % slower for the whole kernel.)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool leader = threadIdx.x == 320;
    const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
    constexpr int NCH = MLP_D / 32;                 // 12
    uint32_t addc = 0;                              // staging chunks so far
    for (int bi = 0; bi < my_blocks; ++bi) {
      const int r0 = (int(blockIdx.x) + bi * int(gridDim.x)) * MLP_BM;
      auto issue_add = [&](int ch, uint32_t gidx) {
        const int b = gidx & 1;
        mbar_expect_tx(&add_bar[b], MLP_GRAN);
        tma_load_3d(sS + size_t(b) * MLP_GRAN, &tmX, &add_bar[b], ch * 32, r0, 0);
      };
      if (leader) {
        tma_store_wait_read<0>();
        issue_add(0, addc);
      }
      mbar_wait(acc2_full, bi & 1);                 // every MMA2 of this block has retired
      tc_fence_after();
      for (int ch = 0; ch < NCH; ++ch, ++addc) {
        const int b = addc & 1;
        float v[32];
        {
          uint32_t r[32];
          __syncwarp();
          tmem_ld_x32(lane_base + ACC2_COL + uint32_t(ch * 32), r);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        }
        tmem_ld_wait();
        if (ch == NCH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);   // the next block's MMA2 may overwrite acc2
        }
        const float* b2 = p.b2 + ch * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(b2 + i));
          v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
        }
        if (leader && ch + 1 < NCH) {
          tma_store_wait_read<0>();                 // the store that last read the other buffer is done
          issue_add(ch + 1, addc + 1);
        }
        mbar_wait(&add_bar[b], (addc >> 1) & 1);
        uint8_t* srow = sS + size_t(b) * MLP_GRAN + row * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4* ptr = reinterpret_cast<float4*>(srow + ((k ^ (row & 7)) << 4));
          float4 a = *ptr;
          a.x += v[4 * k]; a.y += v[4 * k + 1]; a.z += v[4 * k + 2]; a.w += v[4 * k + 3];
          *ptr = a;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (leader) {
          tma_store_3d(&tmX, sS + size_t(b) * MLP_GRAN, ch * 32, r0, 0);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace dsg
