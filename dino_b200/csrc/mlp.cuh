// Fused transformer MLP for D = 384 (ViT-S):   x += fc2( gelu( fc1( LayerNorm2(x) ) ) )
// (reference vision_transformer.py:135 -> :59-65 Mlp.forward with :118 norm2; nn.GELU() = exact erf)
//
// One persistent CTA per SM walks 128-token row blocks.  Nothing of the 1536-wide hidden activation ever
// leaves the SM (the unfused path writes and re-reads 708 MB of it per block at batch 64), and the
// hidden bias / GELU / residual passes are fused:
//
//   A block     : LayerNorm2(x) as bf16, resident in shared memory (6 k-blocks x 16 KB) for the 12 hidden chunks of a row
//                 block.  Fused form (MlpParams::fuse_ln): the four output warps - idle between two drains - read the
//                 block's 128 rows of x (fp32, prefetched into L2 one block ahead), normalise them ((x - mean) * rstd;
//                 gamma / beta live in W1 / b1) and write the bf16 operand in the UMMA layout as soon as the last MMA1
//                 of the previous block has retired: no LayerNorm launch, no bf16 copy of the tokens in HBM.
//                 Unfused form: the block is TMA-loaded from the bf16 copy the LayerNorm kernel wrote.
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
#include "kernels.cuh"

namespace dsg {

struct MlpParams {
  int M;                      // token rows
  const float* x;             // [M, 384] fp32 residual stream (read by the LayerNorm fill; written through tmX)
  const float* b1;            // [1536]
  const float* b2;            // [384]
  int reverse;                // 1: row blocks are processed last to first
  // fused LayerNorm2 (reference vision_transformer.py:118, :135): with fuse_ln the A block is not TMA-loaded from a bf16
  // copy written by the LayerNorm kernel but produced in place from x by the output warps as xhat = (x - mean) * rstd;
  // gamma and beta are folded into W1 / b1 at weight-load time (fold_ln_weight_kernel)
  int fuse_ln;
  float ln_eps;
  long long* timing;          // debug (DSG_MLP_TIMING): [grid][2 roles][8] cycle totals
  int* hb;                    // diagnostic heartbeat (see hb_mark), may be null
};

#ifdef DSG_MLP_TIMING
#define MLP_T(i) do { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } while (0)
#define MLP_T_DECL long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tprev = clock64(); \
  const long long tclk0 = tprev; const long long tns0 = mlp_globaltimer()
// slots 5 / 6 of the MMA role: SM cycles and nanoseconds of the whole loop (-> the SM clock the kernel really ran at)
#define MLP_T_CLOCK() do { tacc[5] = clock64() - tclk0; tacc[6] = mlp_globaltimer() - tns0; } while (0)
#define MLP_T_DUMP(role) do { if (p.timing) for (int _i = 0; _i < 8; ++_i) \
    p.timing[(size_t(blockIdx.x) * 2 + (role)) * 8 + _i] = tacc[_i]; } while (0)
#else
#define MLP_T(i) do { } while (0)
#define MLP_T_DECL do { } while (0)
#define MLP_T_DUMP(role) do { } while (0)
#define MLP_T_CLOCK() do { } while (0)
#endif

__device__ __forceinline__ long long mlp_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int MLP_D = 384;
constexpr int MLP_HID = 1536;
constexpr int MLP_BM = 128;
constexpr int MLP_CH = 128;                         // hidden columns per chunk
constexpr int MLP_NCH = MLP_HID / MLP_CH;           // 12
constexpr int MLP_KB = MLP_D / 64;                  // 6 k-blocks of A / W1 granules per chunk
constexpr int MLP_GRAN = 128 * 64 * 2;              // 16 KB
#ifndef DSG_MLP_RING
#define DSG_MLP_RING 4
#endif
constexpr int MLP_RING = DSG_MLP_RING;
constexpr int MLP_LAG = 1;                          // MMA2(c) is issued right after MMA1(c + MLP_LAG) (a lag of 2 measured 25 % slower: single G buffer)
constexpr int MLP_THREADS = 64 + 256 + 128;         // producer + MMA, 8 GELU warps, 4 output warps
constexpr size_t MLP_SMEM = size_t(MLP_KB) * MLP_GRAN + 2 * MLP_GRAN + 2 * MLP_GRAN + size_t(MLP_RING) * MLP_GRAN + 1024 + 512 +
                            MLP_BM * sizeof(float2);   // + barriers (512) + LayerNorm statistics of the next row block (1 KB)

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

// PAIR = true: the kernel runs as CTA pairs (clusters of two, tcgen05 cta_group::2).  A pair owns two neighbouring row
// blocks (256 rows); every MMA is one M = 256 instruction issued by the leader CTA (cluster rank 0) and executed by both
// SMs, each on its own 128 rows of A / G and its own accumulators, with the B operand (a weight granule) SPLIT between
// the two CTAs' shared memories: each CTA streams only half of the weights (64 of the 128 rows of every granule).
// The kernel is bound by the L2 -> SM delivery of the weights (2.36 MB per row block at ~42 B/clk/SM); the pair halves
// it, and the ring holds 8 half granules instead of 4 whole ones.  (A TMA-multicast variant without cta_group::2 -
// same bytes delivered to every SM - measured no gain at all.)
// Barriers: everything the MMA issuer waits on lives in the leader CTA and collects both CTAs' arrivals (remote
// mbarrier arrives, TMA complete_tx from the peer's loads); everything it signals is a multicast commit to both CTAs.
template <bool PAIR>
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
  uint8_t* sR = sS + 2 * MLP_GRAN;                  // [4][16 KB]  weight granule ring  (PAIR: [8][8 KB] half granules)
  constexpr int RING = PAIR ? 2 * MLP_RING : MLP_RING;
  constexpr uint32_t SLOT = PAIR ? MLP_GRAN / 2 : MLP_GRAN;
  constexpr uint32_t NCTA = PAIR ? 2 : 1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + size_t(MLP_RING) * MLP_GRAN);
  uint64_t* w_full = bars;                          // RING
  uint64_t* w_empty = w_full + 2 * MLP_RING;        // RING
  uint64_t* a_full = w_empty + 2 * MLP_RING;        // 1 (TMA transaction barrier: 6 granules per CTA)
  uint64_t* a_empty = a_full + 1;                   // 1 (commit after the last MMA1 of a row block)
  uint64_t* acc1_full = a_empty + 1;                // 1 (commit)
  uint64_t* acc1_empty = acc1_full + 1;             // 1 (8 arrivals per CTA)
  uint64_t* g_full = acc1_empty + 1;                // 1 (8 arrivals per CTA)
  uint64_t* g_empty = g_full + 1;                   // 1 (commit after MMA2 of the chunk)
  uint64_t* acc2_full = g_empty + 1;                // 1 (commit after the last MMA2)
  uint64_t* acc2_empty = acc2_full + 1;             // 1 (4 arrivals per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);
  float2* sStats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [128] (mean, rstd) per row

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.M + MLP_BM - 1) / MLP_BM;
  // plain: row blocks bid, bid + G, ...; PAIR: cluster k takes the block pairs k, k + G/2, ... and rank r the r-th of
  // a pair (a block index past the end is harmless: TMA zero-fills its loads and clips its stores)
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0;
  const int units = PAIR ? (m_blocks + 1) / 2 : m_blocks;
  const int unit0 = PAIR ? int(blockIdx.x) / 2 : int(blockIdx.x);
  const int unit_stride = PAIR ? int(gridDim.x) / 2 : int(gridDim.x);
  const int my_blocks = unit0 < units ? (units - 1 - unit0) / unit_stride + 1 : 0;
  // p.reverse: row blocks last to first (the producer of A and x wrote its last rows most recently: L2 hits)
  auto block_row0 = [&](int bi) {
    int unit = unit0 + bi * unit_stride;
    if (p.reverse) unit = units - 1 - unit;
    return (unit * (PAIR ? 2 : 1) + int(cta_rank)) * MLP_BM;
  };
  // arrive on a barrier of the leader CTA (the MMA issuer's side)
  auto arrive_leader = [&](uint64_t* bar) {
    if constexpr (PAIR) mbar_arrive_cluster(mapa_rank(bar, 0));
    else mbar_arrive(bar);
  };

  constexpr int HB_CODE = 200 + (PAIR ? 1 : 0);
  hb_mark(p.hb, HB_CODE, 1);
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < RING; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a_full, p.fuse_ln ? 4 * NCTA : 1);   // fused LayerNorm: one arrival per output warp (both CTAs)
    mbar_init(a_empty, 1);
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 8 * NCTA);
    mbar_init(g_full, 8 * NCTA);
    mbar_init(g_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 4 * NCTA);
    fence_mbar_init();
  }
  // CTA pairs: BOTH CTAs must be running before either issues tcgen05.alloc.cta_group::2.  The two CTAs of a cluster
  // do not necessarily start at the same time: when kernels of other streams occupy the GPU, one SM of the pair can
  // drain much later than the other.  An allocation issued while the peer CTA is not resident yet returns, but the
  // peer's own allocation then blocks for ever (observed: rank 0 past the allocation, rank 1 stuck inside it; never
  // with one stream, where the two CTAs start together).  Hence a cluster barrier first - as CUTLASS' 2-SM kernels do
  // (cluster-wide pipeline-init barrier before the TMEM allocation) - which also publishes the barrier initialisation.
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();           // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  hb_mark(p.hb, HB_CODE, 2);

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      // per row block: A (6 granules, own buffer) | W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W2(11)  through the ring
      uint32_t rc = 0;
      // one ring slot: wait until the MMAs that read it have retired (both SMs), then fetch the next granule - in PAIR
      // mode my 64-row half of it, completing on the leader's barrier (which expects both halves)
      // (row0: first weight row of THIS CTA's box - 128 rows, or 64 in PAIR mode)
      auto load_gran = [&](const CUtensorMap* tm, int c0, int row0) {
        const int s = rc % RING;
        mbar_wait(&w_empty[s], ((rc / RING) & 1) ^ 1);
        if constexpr (PAIR) {
          if (cta_rank == 0) mbar_expect_tx(&w_full[s], 2 * SLOT);
          tma_load_2d_pair(sR + size_t(s) * SLOT, tm, mapa_rank(&w_full[s], 0), c0, row0);
        } else {
          mbar_expect_tx(&w_full[s], SLOT);
          tma_load_2d(sR + size_t(s) * SLOT, tm, &w_full[s], c0, row0);
        }
        ++rc;
      };
      auto load_w1 = [&](int c) {
        for (int kb = 0; kb < MLP_KB; ++kb)                                       // W1[c*128.., kb*64..]
          load_gran(&tmW1, kb * 64, c * MLP_CH + (PAIR ? int(cta_rank) * 64 : 0));
      };
      auto load_w2 = [&](int c) {
        // per k-block: output columns 0..255 as ONE N = 256 operand (two consecutive ring slots, even slot first) and
        // 256..383 as an N = 128 operand; the order alternates so that the pair always starts on an even slot
        for (int kb = 0; kb < MLP_CH / 64; ++kb) {
          const int k0 = c * MLP_CH + kb * 64;
          auto wide = [&]() {
            if constexpr (PAIR) {   // B rows of an N = 256 pair MMA: rank 0 holds 0..127, rank 1 holds 128..255
              load_gran(&tmW2, k0, int(cta_rank) * 128);
              load_gran(&tmW2, k0, int(cta_rank) * 128 + 64);
            } else {                // plain: two N = 128 operands (with 4 slots a two-slot operand stalls the ring)
              load_gran(&tmW2, k0, 0);
              load_gran(&tmW2, k0, 128);
            }
          };
          auto narrow = [&]() { load_gran(&tmW2, k0, 256 + (PAIR ? int(cta_rank) * 64 : 0)); };
          if (kb == 0) { wide(); narrow(); } else { narrow(); wide(); }
        }
      };
      auto load_a = [&](int bi) {
        const int r0 = block_row0(bi);
        if (bi > 0) mbar_wait(a_empty, (bi - 1) & 1);   // last MMA1 of the previous block has read A
        if constexpr (PAIR) {
          if (cta_rank == 0) mbar_expect_tx(a_full, 2 * MLP_KB * MLP_GRAN);
          const uint32_t bar = mapa_rank(a_full, 0);
          for (int kb = 0; kb < MLP_KB; ++kb) tma_load_3d_pair(sA + size_t(kb) * MLP_GRAN, &tmA, bar, kb * 64, r0, 0);
        } else {
          mbar_expect_tx(a_full, MLP_KB * MLP_GRAN);
          for (int kb = 0; kb < MLP_KB; ++kb) tma_load_3d(sA + size_t(kb) * MLP_GRAN, &tmA, a_full, kb * 64, r0, 0);
        }
      };
      const bool tma_a = !p.fuse_ln;             // fused LayerNorm: the output warps produce A
      if (my_blocks > 0 && tma_a) load_a(0);
      for (int bi = 0; bi < my_blocks; ++bi) {
        for (int i = 0; i < MLP_NCH + MLP_LAG; ++i) {
          if (i < MLP_NCH) load_w1(i);
          // the A block of the next row block: right after the weights of the last MMA1 have been requested
          if (i == MLP_NCH - 1 && bi + 1 < my_blocks && tma_a) load_a(bi + 1);
          if (i >= MLP_LAG) load_w2(i - MLP_LAG);
        }
      }
      // Producer tail: wait for the leader's last multicast commits on this CTA's w_empty / a_empty barriers before the
      // CTA may exit (an arrival still in flight would land in the shared memory of the next CTA on this SM).
      if constexpr (PAIR) {
        for (uint32_t i = 0; i < uint32_t(RING) && i < rc; ++i) {
          const uint32_t idx = rc - 1 - i;
          mbar_wait(&w_empty[idx % RING], (idx / RING) & 1);
        }
        if (my_blocks > 0 && tma_a) mbar_wait(a_empty, (my_blocks - 1) & 1);   // (fused: the output warps wait for it)
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0 && elect_one()) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, 128, 0);
      auto wait = [&](uint64_t* bar, uint32_t parity) {
        if constexpr (PAIR) mbar_wait_cluster(bar, parity);
        else mbar_wait(bar, parity);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) tc_commit_pair(bar, uint16_t(3));
        else tc_commit(bar);
      };
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
        if constexpr (PAIR) umma_ss_pair(d, ad, bd, idesc, acc);
        else umma_ss(d, ad, bd, idesc, acc);
      };
      constexpr uint32_t idesc_wide = umma_idesc_bf16(256, 256, 0);   // PAIR only: fc2 columns 0..255 in one instruction
      uint32_t rc = 0;        // ring counter (same order as the producer)
      uint32_t c1 = 0;        // acc1 uses so far
      uint32_t gc = 0;        // G uses so far
      MLP_T_DECL;
      auto wait_gran = [&]() -> uint32_t {
        const int s = rc % RING;
        MLP_T(7);
        wait(&w_full[s], (rc / RING) & 1);
        MLP_T(0);
        tc_fence_after();
        return smem_u32(sR + size_t(s) * SLOT);
      };
      auto free_gran = [&]() {
        commit(&w_empty[rc % RING]);
        ++rc;
      };
      auto mma1 = [&](int c) {
        // acc1 = A . W1[c]^T ; the GELU warps must have pulled the previous acc1 into registers
        MLP_T(7);
        wait(acc1_empty, (c1 & 1) ^ 1);
        MLP_T(1);
        tc_fence_after();
        for (int kb = 0; kb < MLP_KB; ++kb) {
          const uint32_t sw = wait_gran();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + size_t(kb) * MLP_GRAN));
          const uint64_t bdesc = umma_desc_sw128(sw);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma(tmem_base + ACC1_COL, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), (kb | k) != 0);
          free_gran();
        }
        commit(acc1_full);
        ++c1;
        if (c == MLP_NCH - 1) commit(a_empty);   // the next row block's A may be loaded
      };
      auto mma2 = [&](int c, bool first_of_block) {
        // acc2 += G . W2[:, c]^T
        MLP_T(7);
        wait(g_full, gc & 1);
        MLP_T(2);
        tc_fence_after();
        // PAIR: N = 256 costs 160 clk per K = 16 step against 2 x 96 for two N = 128 instructions (SS operands: 32 + N/2)
        for (int kb = 0; kb < MLP_CH / 64; ++kb) {
          const uint64_t adesc = umma_desc_sw128(smem_u32(sG + size_t(kb) * MLP_GRAN));
          auto wide = [&]() {
            if constexpr (PAIR) {
              const uint32_t sw = wait_gran();        // slot rc (even) ...
              ++rc;
              (void)wait_gran();                      // ... and rc + 1: 128 contiguous weight rows per CTA
              --rc;
              const uint64_t bdesc = umma_desc_sw128(sw);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_ss_pair(tmem_base + ACC2_COL, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc_wide,
                             (!first_of_block || kb != 0 || k != 0) ? 1u : 0u);
              free_gran();
              free_gran();
            } else {
              for (int nt = 0; nt < 2; ++nt) {
                const uint32_t sw = wait_gran();
                const uint64_t bdesc = umma_desc_sw128(sw);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma(tmem_base + ACC2_COL + uint32_t(nt * 128), adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2),
                      (!first_of_block || kb != 0 || k != 0) ? 1u : 0u);
                free_gran();
              }
            }
          };
          auto narrow = [&]() {
            const uint32_t sw = wait_gran();
            const uint64_t bdesc = umma_desc_sw128(sw);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(tmem_base + ACC2_COL + 256u, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2),
                  (!first_of_block || kb != 0 || k != 0) ? 1u : 0u);
            free_gran();
          };
          if (kb == 0) { wide(); narrow(); } else { narrow(); wide(); }
        }
        commit(g_empty);
        ++gc;
        (void)c;
      };
      for (int bi = 0; bi < my_blocks; ++bi) {
        MLP_T(7);
        wait(a_full, bi & 1);                       // LN2(x) row block has landed
        MLP_T(3);
        tc_fence_after();
        for (int i = 0; i < MLP_NCH + MLP_LAG; ++i) {
          if (i < MLP_NCH) mma1(i);
          if (i >= MLP_LAG) {
            const int c = i - MLP_LAG;
            if (c == 0) {
              MLP_T(7);
              wait(acc2_empty, (bi & 1) ^ 1);       // the output warps have drained the previous block's acc2
              MLP_T(4);
              tc_fence_after();
            }
            mma2(c, c == 0);
          }
        }
        commit(acc2_full);
      }
      MLP_T_CLOCK();
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
        if (lane == 0) arrive_leader(acc1_empty);   // MMA1 of the next chunk may overwrite acc1
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
        __syncwarp();
        if (lane == 0) arrive_leader(g_full);
        MLP_T(4);
      }
    }
    if constexpr (PAIR) {
      if (cc > 0) mbar_wait(g_empty, (cc - 1) & 1);   // the commit after the very last MMA2 (see the producer tail)
    }
    if (threadIdx.x == 64) { MLP_T_DUMP(1); }
  } else {
    // ---------------- output warps (4): x += acc2 + b2, 12 chunks of 32 fp32 columns ----------------
    // acc2 + b2 goes through a swizzled staging buffer into a TMA reduce-add on x: the fp32 add happens at L2 (one
    // add per element, so the result is the same single rounding as x + (acc2 + b2) in registers) and x is never
    // loaded into the SM.  Loading x chunks with TMA, adding in shared memory and storing back made this drain
    // latency-bound (~1500 clk per chunk, one load in flight) and cost the MMA issuer ~14k clk per row block waiting
    // for acc2; plain per-thread loads of x were 25 % slower still.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool leader = threadIdx.x == 320;
    const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
    constexpr int NCH = MLP_D / 32;                 // 12
    uint32_t addc = 0;                              // staging chunks so far
    // ---- fused LayerNorm2: A block of row block bi from x ----
    // Warp w of the four takes rows w*32 .. w*32+31, four rows per pass with the statistics of the LayerNorm kernel
    // (ln384_stats in kernels.cuh), and writes xhat = (x - mean) * rstd as bf16 into the k-block slots of sA in the
    // K-major SWIZZLE_128B layout the TMA load would have produced: column c of row r -> slot c/64, byte
    // r*128 + (((c%64)/8 ^ (r&7)) << 4) + (c%8)*2.  Rows >= M are written as zeros (what TMA's fill does).
    const bool fuse_ln = p.fuse_ln != 0;
    const int ow = warp - 10;                       // 0..3
    auto prefetch_block = [&](int bi) {             // the block's 128 rows of x -> L2 (192 KB: 12 lines per thread)
      const int r0 = block_row0(bi);
      const char* base = reinterpret_cast<const char*>(p.x + size_t(r0) * MLP_D);
      const size_t bytes = size_t(max(0, min(MLP_BM, p.M - r0))) * MLP_D * sizeof(float);
      for (size_t off = size_t(threadIdx.x - 320) * 128; off < bytes; off += 128 * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    };
    // Eight lanes per row, four rows per pass (ln384_stats): lane `sub` of group `grp` holds float4 number i*8 + sub of
    // row ow*32 + pass*4 + grp, i.e. columns (i*8 + sub)*4 ..+3 -> k-block i/2, 16-byte chunk ((i&1)*8 + sub)/2 of the
    // row's 128 bytes, 8-byte half sub&1.
    const int sub = lane & 7, grp = lane >> 3;
    auto ln_load = [&](float4 (&v)[LN384_V], int r0, int pass) {
      const int row = r0 + ow * 32 + pass * 4 + grp;
      const float4* xr = reinterpret_cast<const float4*>(p.x + size_t(row < p.M ? row : 0) * MLP_D);
#pragma unroll
      for (int i = 0; i < LN384_V; ++i) v[i] = row < p.M ? __ldg(xr + i * 8 + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // Two phases.  (1) Statistics: any time before the A buffer frees up - while the MMAs of the current block run and
    // these warps have nothing else to do - the block's rows are read once (HBM -> L2 -> registers), mean and rstd of
    // each row go to shared memory (1 KB).  (2) Write: once the last MMA1 of the previous block has retired, the rows are
    // read again (L2 hits), normalised with the stored statistics and written into sA.  Phase 2 is what the tensor pipe
    // may have to wait for, and it is pure streaming: no reductions, no shuffles, loads one pass ahead.
    auto ln_stats_pass = [&](int bi) {
      const int r0 = block_row0(bi);
      float4 va[LN384_V], vb[LN384_V];
      ln_load(va, r0, 0);
#pragma unroll 1
      for (int pass = 0; pass < 8; pass += 2) {
        ln_load(vb, r0, pass + 1);
        float mean, rstd;
        ln384_stats(va, p.ln_eps, mean, rstd);
        if (sub == 0) sStats[ow * 32 + pass * 4 + grp] = make_float2(mean, rstd);
        if (pass + 2 < 8) ln_load(va, r0, pass + 2);
        ln384_stats(vb, p.ln_eps, mean, rstd);
        if (sub == 0) sStats[ow * 32 + (pass + 1) * 4 + grp] = make_float2(mean, rstd);
      }
      __syncwarp();                                 // the statistics are read back by the lanes of this warp only
    };
    auto ln_store = [&](const float4 (&v)[LN384_V], int r0, int pass) {
      const int r = ow * 32 + pass * 4 + grp;       // row inside the block
      const float2 st = sStats[r];
      const bool live = r0 + r < p.M;
      uint8_t* rowp = sA + r * 128 + (sub & 1) * 8;
#pragma unroll
      for (int i = 0; i < LN384_V; ++i) {
        uint2 o = ln384_out_plain(v[i], st.x, st.y);
        if (!live) o = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(rowp + size_t(i >> 1) * MLP_GRAN + (((((i & 1) * 8 + sub) >> 1) ^ (r & 7)) << 4)) = o;
      }
    };
    auto ln_fill = [&](int bi) {
      const int r0 = block_row0(bi);
      ln_stats_pass(bi);
      float4 va[LN384_V], vb[LN384_V];
      ln_load(va, r0, 0);                           // (requested before the wait for the A buffer)
      if (bi > 0) mbar_wait(a_empty, (bi - 1) & 1); // the last MMA1 of the previous block has read A
#pragma unroll 1
      for (int pass = 0; pass < 8; pass += 2) {
        ln_load(vb, r0, pass + 1);
        ln_store(va, r0, pass);
        if (pass + 2 < 8) ln_load(va, r0, pass + 2);
        ln_store(vb, r0, pass + 1);
      }
      fence_proxy_async_smem();                     // generic-proxy writes -> visible to the UMMA reads of sA
      __syncwarp();
      if (lane == 0) arrive_leader(a_full);
    };
    if (fuse_ln && my_blocks > 0) {
      ln_fill(0);
      if (my_blocks > 1) prefetch_block(1);
    }
    for (int bi = 0; bi < my_blocks; ++bi) {
      const int r0 = block_row0(bi);
      if (fuse_ln && bi + 1 < my_blocks) {
        // the next block's A: as soon as this block's last MMA1 has retired (a_empty), i.e. while its last MMA2s run
        ln_fill(bi + 1);
        if (bi + 2 < my_blocks) prefetch_block(bi + 2);
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
          if (lane == 0) arrive_leader(acc2_empty); // the next block's MMA2 may overwrite acc2
        }
        const float* b2 = p.b2 + ch * 32;
        // staging buffer b was last read by the reduce of chunk addc - 2: the leader waited for that read before the
        // barrier of chunk addc - 1
        uint8_t* srow = sS + size_t(b) * MLP_GRAN + row * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(b2 + 4 * k));
          *reinterpret_cast<float4*>(srow + ((k ^ (row & 7)) << 4)) =
              make_float4(v[4 * k] + bv.x, v[4 * k + 1] + bv.y, v[4 * k + 2] + bv.z, v[4 * k + 3] + bv.w);
        }
        fence_proxy_async_smem();
        if (leader) tma_store_wait_read<0>();       // the other buffer is free again
        named_bar_sync(2, 128);
        if (leader) {
          tma_reduce_add_3d(&tmX, sS + size_t(b) * MLP_GRAN, ch * 32, r0, 0);
          tma_store_commit();
        }
      }
    }
    if constexpr (PAIR) {
      if (fuse_ln && my_blocks > 0) mbar_wait(a_empty, (my_blocks - 1) & 1);   // producer tail (see the TMA producer)
    }
    if (leader) tma_store_wait<0>();
  }

  hb_mark(p.hb, HB_CODE, 3);
  tc_fence_before();
  __syncthreads();
  hb_mark(p.hb, HB_CODE, 4);
  if constexpr (PAIR) cluster_sync_all();           // both CTAs are done with each other's barriers and TMEM
  hb_mark(p.hb, HB_CODE, 5);
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
    hb_mark_left(p.hb, HB_CODE);                     // by the deallocating warp: a CTA stuck in dealloc stays visible
  }
}

}  // namespace dsg
