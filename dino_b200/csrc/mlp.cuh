// Fused transformer MLP for D = 384 (ViT-S):   x += fc2( gelu( fc1( LayerNorm2(x) ) ) )
// (reference vision_transformer.py:135 -> :59-65 Mlp.forward with :118 norm2; nn.GELU() = exact erf)
//
// One persistent CTA per SM walks 128-token row blocks.  Nothing of the 1536-wide hidden activation ever
// leaves the SM (the unfused path writes and re-reads 708 MB of it per block at batch 64), and the
// LayerNorm kernel in front of fc1 disappears as well:
//
//   LN prologue : the 8 epilogue warps read the 128 x 384 fp32 rows of x, normalise them and write the bf16
//                 result straight into shared memory in the K-major SWIZZLE_128B layout of a UMMA A operand
//                 (6 k-blocks x 16 KB, resident for the whole row block)
//   per hidden chunk c of 128 columns (12 chunks):
//     MMA1(c)   : acc1[128 x 128] = A . W1[c]^T              (24 MMAs 128x128x16, W1 granules via TMA)
//     epilogue  : acc1 -> registers -> + b1 -> GELU -> bf16 -> shared memory G[c&1] (again an A-operand layout)
//     MMA2(c)   : acc2[128 x 384] += G[c&1] . W2[:, c]^T     (3 x 8 MMAs 128x128x16, W2 granules via TMA)
//   final       : acc2 + b2 + x -> x   (x chunks TMA-loaded into staging, added in place, TMA-stored)
//
// TMEM: acc2 = 384 fp32 columns, acc1 = 128 columns -> exactly the 512 columns of an SM.
// The tensor pipe runs  MMA1(0) | MMA1(1) MMA2(0) | MMA1(2) MMA2(1) | ...  so the GELU of chunk c hides
// under MMA1(c+1).  All weight tiles are 16 KB "granules" ([128 rows x 64 k] bf16) that stream through
// one 4-deep TMA ring in exactly the order the MMAs consume them.
//
// Shared memory: A 96 KB + G 2 x 32 KB + ring 4 x 16 KB = 224 KB (+ barriers); the staging buffers of the
// final epilogue alias G (free once the last MMA2 has retired).
#pragma once
#include "gemm.cuh"

namespace dsg {

struct MlpParams {
  int M;                      // token rows
  const float* x;             // [M, 384] fp32 residual stream (read by the LN prologue; updated through tmX)
  const float* ln_g;          // [384]
  const float* ln_b;          // [384]
  const float* b1;            // [1536]
  const float* b2;            // [384]
  float eps;
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
constexpr int MLP_G2 = (MLP_D / 128) * (MLP_CH / 64);   // 6 W2 granules per chunk: [n-third][k-block]
constexpr int MLP_GRAN = 128 * 64 * 2;              // 16 KB
constexpr int MLP_RING = 4;
constexpr int MLP_THREADS = 64 + 256;
constexpr size_t MLP_SMEM = size_t(MLP_KB) * MLP_GRAN + 2 * 2 * MLP_GRAN + size_t(MLP_RING) * MLP_GRAN + 1024 + 512;

__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmX, const MlpParams p) {
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ACC2_COL = 0;                  // 384 columns
  constexpr uint32_t ACC1_COL = 384;                // 128 columns

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;                               // [6][16 KB]  LN(x) as bf16, K-major SW128
  uint8_t* sG = sA + size_t(MLP_KB) * MLP_GRAN;     // [2 buffers][2 k-blocks][16 KB]  gelu(fc1) chunk
  uint8_t* sR = sG + 4 * MLP_GRAN;                  // [4][16 KB]  weight granule ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + size_t(MLP_RING) * MLP_GRAN);
  uint64_t* w_full = bars;                          // 4
  uint64_t* w_empty = w_full + MLP_RING;            // 4
  uint64_t* a_full = w_empty + MLP_RING;            // 1 (256 arrivals: LN threads)
  uint64_t* a_empty = a_full + 1;                   // 1 (commit after the last MMA1 of a row block)
  uint64_t* acc1_full = a_empty + 1;                // 1 (commit)
  uint64_t* acc1_empty = acc1_full + 1;             // 1 (8 arrivals)
  uint64_t* g_full = acc1_empty + 1;                // 2 (256 arrivals)
  uint64_t* g_empty = g_full + 2;                   // 2 (commit after MMA2 of the chunk)
  uint64_t* acc2_full = g_empty + 2;                // 1 (commit after the last MMA2)
  uint64_t* acc2_empty = acc2_full + 1;             // 1 (8 arrivals)
  uint64_t* add_bar = acc2_empty + 1;               // 4 (x chunks landed in staging)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(add_bar + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.M + MLP_BM - 1) / MLP_BM;
  const int my_blocks = int(blockIdx.x) < m_blocks ? (m_blocks - 1 - int(blockIdx.x)) / int(gridDim.x) + 1 : 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < MLP_RING; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a_full, 256);
    mbar_init(a_empty, 1);
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 8);
    for (int s = 0; s < 2; ++s) { mbar_init(&g_full[s], 256); mbar_init(&g_empty[s], 1); }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 8);
    for (int s = 0; s < 4; ++s) mbar_init(&add_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer: weight granules in consumption order ----------------
      // per row block:  W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W1(11) W2(10) | W2(11)
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
            tma_load_2d(sR + size_t(s) * MLP_GRAN, &tmW2, &w_full[s], c * MLP_CH + kb * 64, nt * 128);  // W2[nt*128.., c*128+kb*64..]
          }
        }
      };
      for (int bi = 0; bi < my_blocks; ++bi) {
        load_w1(0);
        for (int c = 0; c < MLP_NCH; ++c) {
          if (c + 1 < MLP_NCH) load_w1(c + 1);
          load_w2(c);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0);
      uint32_t rc = 0;        // ring counter (same order as the producer)
      uint32_t c1 = 0;        // acc1 uses so far (phase of acc1_full / acc1_empty)
      uint32_t gc[2] = {0, 0};  // uses of each G buffer
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
      auto mma1 = [&](int c, int bi) {
        // acc1 = A . W1[c]^T ; the epilogue must have pulled the previous acc1 into registers
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
        if (c == MLP_NCH - 1) tc_commit(a_empty);   // the LN prologue of the next row block may overwrite A
        (void)bi;
      };
      auto mma2 = [&](int c, bool first_of_block) {
        // acc2 += G[c&1] . W2[:, c]^T
        const int gb = c & 1;
        MLP_T(7);
        mbar_wait(&g_full[gb], gc[gb] & 1);
        MLP_T(2);
        tc_fence_after();
        for (int kb = 0; kb < MLP_CH / 64; ++kb) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sG + size_t(gb * 2 + kb) * MLP_GRAN));
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
        tc_commit(&g_empty[gb]);
        ++gc[gb];
      };
      for (int bi = 0; bi < my_blocks; ++bi) {
        MLP_T(7);
        mbar_wait(a_full, bi & 1);                  // LN(x) of this row block is in shared memory
        MLP_T(3);
        tc_fence_after();
        mma1(0, bi);
        for (int c = 0; c < MLP_NCH; ++c) {
          if (c + 1 < MLP_NCH) mma1(c + 1, bi);
          if (c == 0) {
            MLP_T(7);
            mbar_wait(acc2_empty, (bi & 1) ^ 1);    // the final epilogue of the previous block has drained acc2
            MLP_T(4);
            tc_fence_after();
          }
          mma2(c, c == 0);
        }
        tc_commit(acc2_full);
      }
      MLP_T_DUMP(0);
    }
  } else {
    // ---------------- LN prologue / GELU epilogue / final epilogue (8 warps) ----------------
    const int ew = warp - 2;                        // 0..7
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int half = ew >> 2;                       // column half
    const int row = quarter * 32 + lane;
    const bool leader = threadIdx.x == 64;
    const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
    uint32_t c1 = 0;                                // acc1 uses so far
    uint32_t gc[2] = {0, 0};
    uint32_t addc = 0;                              // staging chunks so far (add_bar phases)
    MLP_T_DECL;

    for (int bi = 0; bi < my_blocks; ++bi) {
      const int r0 = (int(blockIdx.x) + bi * int(gridDim.x)) * MLP_BM;
      MLP_T(7);

      // ---- LayerNorm prologue: rows r0 + ew*16 .. +15, one row per warp pass; lane l < 24 owns columns 16l..16l+15
      if (bi > 0) mbar_wait(a_empty, (bi - 1) & 1);  // last MMA1 of the previous block has read A
      {
        const int cbase = lane * 16;
        const bool active = lane < 24;
        float4 gam[4], bet[4];
        if (active) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            gam[i] = __ldg(reinterpret_cast<const float4*>(p.ln_g + cbase) + i);
            bet[i] = __ldg(reinterpret_cast<const float4*>(p.ln_b + cbase) + i);
          }
        }
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
          const int lr = ew * 16 + rr;              // row inside the block
          const int gr = r0 + lr;
          float4 v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active && gr < p.M) {
            const float4* src = reinterpret_cast<const float4*>(p.x + size_t(gr) * MLP_D + cbase);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = src[i];
          }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float mean = sum * (1.0f / MLP_D);
          float sq = 0.f;
          if (active) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
              sq += (a * a + b * b) + (c * c + d * d);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
          const float rstd = rsqrtf(sq * (1.0f / MLP_D) + p.eps);
          if (active) {
            // 16 columns = two 16-byte chunks of k-block (lane / 4), chunk index (lane % 4) * 2 + {0, 1}
            uint8_t* dst = sA + size_t(lane >> 2) * MLP_GRAN + lr * 128;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const float4 a = v[2 * h2], b = v[2 * h2 + 1];
              const float4 ga = gam[2 * h2], gb = gam[2 * h2 + 1], ba = bet[2 * h2], bb = bet[2 * h2 + 1];
              uint4 q;
              q.x = pack_bf16x2((a.x - mean) * rstd * ga.x + ba.x, (a.y - mean) * rstd * ga.y + ba.y);
              q.y = pack_bf16x2((a.z - mean) * rstd * ga.z + ba.z, (a.w - mean) * rstd * ga.w + ba.w);
              q.z = pack_bf16x2((b.x - mean) * rstd * gb.x + bb.x, (b.y - mean) * rstd * gb.y + bb.y);
              q.w = pack_bf16x2((b.z - mean) * rstd * gb.z + bb.z, (b.w - mean) * rstd * gb.w + bb.w);
              const int chunk = (lane & 3) * 2 + h2;
              *reinterpret_cast<uint4*>(dst + ((chunk ^ (lr & 7)) << 4)) = q;
            }
          }
        }
      }
      fence_proxy_async_smem();                     // generic-proxy writes -> visible to the UMMA reads
      mbar_arrive(a_full);
      MLP_T(0);

      // ---- GELU epilogue per hidden chunk: thread = (row, column half of 64)
      for (int c = 0; c < MLP_NCH; ++c, ++c1) {
        MLP_T(7);
        mbar_wait(acc1_full, c1 & 1);
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
        const int gb = c & 1;
        const float* b1 = p.b1 + c * MLP_CH + half * 64;
        uint4 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 ba = __ldg(reinterpret_cast<const float4*>(b1 + 8 * k));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(b1 + 8 * k + 4));
          q[k].x = pack_bf16x2(gelu_erf(v[8 * k + 0] + ba.x), gelu_erf(v[8 * k + 1] + ba.y));
          q[k].y = pack_bf16x2(gelu_erf(v[8 * k + 2] + ba.z), gelu_erf(v[8 * k + 3] + ba.w));
          q[k].z = pack_bf16x2(gelu_erf(v[8 * k + 4] + bb.x), gelu_erf(v[8 * k + 5] + bb.y));
          q[k].w = pack_bf16x2(gelu_erf(v[8 * k + 6] + bb.z), gelu_erf(v[8 * k + 7] + bb.w));
        }
        // G buffer gb was last read by MMA2 of chunk c-2
        MLP_T(2);
        mbar_wait(&g_empty[gb], (gc[gb] & 1) ^ 1);
        MLP_T(3);
        uint8_t* dst = sG + size_t(gb * 2 + half) * MLP_GRAN + row * 128;   // k-block = column half
#pragma unroll
        for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(dst + ((k ^ (row & 7)) << 4)) = q[k];
        fence_proxy_async_smem();
        mbar_arrive(&g_full[gb]);
        ++gc[gb];
        MLP_T(4);
      }

      // ---- final epilogue: x += acc2 + b2, 12 chunks of 32 fp32 columns through 4 staging buffers (alias of G)
      MLP_T(7);
      mbar_wait(acc2_full, bi & 1);                 // every MMA2 of this block has retired: acc2 complete, G free
      MLP_T(5);
      tc_fence_after();
      constexpr int NCH = MLP_D / 32;               // 12
      constexpr int PD = 2;
      auto issue_add = [&](int ch, uint32_t gidx) {
        const int b = gidx & 3;
        mbar_expect_tx(&add_bar[b], MLP_GRAN);
        tma_load_3d(sG + size_t(b) * MLP_GRAN, &tmX, &add_bar[b], ch * 32, r0, 0);
      };
      if (leader) {
        // the staging buffers were last read by the TMA stores of the previous block: all of them have been
        // waited for below (wait_group.read 0 at the end of the block)
        for (int ch = 0; ch < PD; ++ch) issue_add(ch, addc + ch);
      }
      for (int ch = 0; ch < NCH; ++ch, ++addc) {
        const int b = addc & 3;
        float v[16];
        {
          uint32_t r[16];
          __syncwarp();
          tmem_ld_x16(lane_base + ACC2_COL + uint32_t(ch * 32 + half * 16), r);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        tmem_ld_wait();
        if (ch == NCH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);   // the next block's MMA2 may overwrite acc2
        }
        const float* b2 = p.b2 + ch * 32 + half * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(b2 + i));
          v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
        }
        if (leader) {
          tma_store_wait_read<4 - PD - 1>();        // the store that last read buffer (ch + PD) & 3 is done
          if (ch + PD < NCH) issue_add(ch + PD, addc + PD);
        }
        mbar_wait(&add_bar[b], (addc >> 2) & 1);
        uint8_t* srow = sG + size_t(b) * MLP_GRAN + row * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4* ptr = reinterpret_cast<float4*>(srow + (((half * 4 + k) ^ (row & 7)) << 4));
          float4 a = *ptr;
          a.x += v[4 * k]; a.y += v[4 * k + 1]; a.z += v[4 * k + 2]; a.w += v[4 * k + 3];
          *ptr = a;
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        if (leader) {
          tma_store_3d(&tmX, sG + size_t(b) * MLP_GRAN, ch * 32, r0, 0);
          tma_store_commit();
        }
      }
      // G / staging is reused by the next block's GELU epilogue: its stores must have finished reading
      if (leader) tma_store_wait_read<0>();
      named_bar_sync(1, 256);
      MLP_T(6);
    }
    if (leader) { MLP_T_DUMP(1); }
    if (leader) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace dsg
