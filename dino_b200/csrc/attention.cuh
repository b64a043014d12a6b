// Flash-style fused multi-head self-attention forward for the DINO ViT blocks
// (reference: vision_transformer.py:80-107 Attention.forward, no mask, dropout p=0).
//
//   S = (q*dh^-0.5) k^T ; P = softmax_rows(S) ; O = P v          per (frame b, head h)
//
// The reference materialises S and P ([B,H,N,N] fp32, 311 MB/frame/block at 480 px); here S, P and
// the running O only ever exist in TMEM and nothing N x N is written anywhere.
//
// q/k/v are read straight out of the qkv GEMM output [B, N, 3*D] bf16 (column = which*D + h*64 + d,
// exactly nn.Linear's output order, reference :82) through ONE 3-D tensor map {3D, N, B}; rows >= N
// of a frame are zero-filled by TMA.  q is pre-scaled by dh^-0.5 * log2(e) by the qkv GEMM epilogue (in fp32,
// before its one rounding to bf16), so S comes out of the tensor pipe in log2 units: P = 2^S.
//
// Work decomposition: a work item is a pair of 128-query tiles (256 queries) of one (b,h).  The
// kernel is persistent (grid = #SMs, item = blockIdx.x + k*gridDim.x) and streams the 128-key K/V
// tiles of the item through a shared-memory ring.
//
// Warp roles (640 threads, 1 CTA/SM, all 512 TMEM columns):
//   warp 0      : TMA producer (Q pair per item; K,V ring)
//   warps 1, 2  : one tcgen05.mma-issuing thread per query tile (warp 1 also allocates TMEM)
//                   S_t = Q_t K^T          (SS, both K-major)            t = 0,1
//                   O_t += P_t V           (TS: P from TMEM, V MN-major: V is never transposed)
//                 S_t, P_t and O_t have their OWN TMEM columns (2x128 + 2x64 + 2x64 = 512), so
//                 QK_t(j+1) is issued as soon as the softmax warps have pulled S_t(j) into registers
//                 (s_empty), long before P_t(j) exists: the softmax warps never wait for the tensor
//                 pipe.  Issue order per key tile j:  QK0(j+1) QK1(j+1) PV0(j) PV1(j).
//   warp 3      : idle (it only completes the producer warpgroup for setmaxnreg)
//   warps 4..11 / 12..19: softmax of query tile 0 / 1 (640 threads in all).  Each warp owns 16 rows and reads them in
//                 the 16x256b accumulator-fragment layout: a row lives in one quad (32 of its 128 scores per thread,
//                 two rows per thread), the row max is two quad shuffles.  Four softmax warps per scheduler: the
//                 tcgen05.ld / row-max / barrier phases of one warp hide under the exponentials of the other three.
//                 (Round 1 used one row per thread, 128 scores per thread, two warps per scheduler: 1.51 ms per
//                 launch against 1.39 ms)
// The exponentials are the bottleneck at head_dim 64 (16 MUFU.EX2 per clock per SM vs 8192 tensor
// flop per clock; 2*128*128 exps per key tile = 2048 MUFU clocks vs 1354 clocks of MMA):
//   * a quarter of the exponentials is evaluated with a degree-3 polynomial on the FMA/ALU pipes
//     (exp2_poly_x2, packed fp32x2 math) instead of MUFU.EX2;
//   * NO row maximum at all in the kernel that normally runs (attn_fwd_kernel<.., true>, att_softmax_unshifted):
//     P = 2^S straight from the accumulator, valid while the row sums stay within [2^-100, 2^100]; rows outside raise
//     a flag and the launch is redone by the classic kernel (attn_fwd_kernel<.., false>: running maximum, raised only
//     when it grows by more than 2^8 - lazy rescale - so that the O accumulator is almost never touched);
//   * the two softmax warpgroups run unsynchronised: their exponential phases overlap partially and
//     the loads / row-max / barrier phases of one hide under the MUFU work of the other (a strict
//     ping-pong between them measured 5 % slower, four half-row warpgroups 15 % slower).
#pragma once
#include "ptx.cuh"

namespace dsg {

struct AttnParams {
  int B, H, N;          // frames, heads, tokens per frame
  int D;                // embed dim (= H*64)
  __nv_bfloat16* out;   // [B*N, D], column = h*64 + d   (reference :104 transpose(1,2).reshape)
  int full_pairs;       // floor(ceil(N / 128) / 2): work items of two query tiles per (frame, head)
  int lone;             // 1: the number of query tiles is odd (one tail tile per (frame, head), see att_decode)
  int num_items;        // regular + tail items
  long long* timing;    // debug (DSG_ATTN_TIMING builds): [grid][2 warpgroups][8] phase cycle totals
  int* hb;              // diagnostic heartbeat (see hb_mark), may be null
  int* range_flag;      // device word shared by the two kernels of one attention launch (see attn_fwd_kernel)
  int launch_id;        // != 0; what the unshifted kernel writes to *range_flag when a row leaves its range
};

#ifdef DSG_ATTN_TIMING
#define ATT_T(i) do { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } while (0)
#else
#define ATT_T(i) do { } while (0)
#endif

constexpr int ATT_BM = 128;     // queries per tile (two tiles per work item)
constexpr int ATT_BN = 128;     // keys per tile
constexpr int ATT_DH = 64;      // head dim
constexpr int ATT_SMW = 8;                                  // softmax warps per query tile
constexpr int ATT_THREADS = 128 + 2 * 32 * ATT_SMW;         // producer warpgroup + 2 query tiles x 8 warps = 640
// bit i: pair i of every 16-pair chunk uses the polynomial exp2 instead of MUFU.EX2.  With four softmax warps per
// scheduler the kernel is bound by the exponentials AND by its instruction issue, and a polynomial pair costs 13
// instructions against 5: an eighth of the pairs is the optimum (1.437 / 1.404 / 1.390 / 1.399 / 1.425 / 1.520 ms per
// launch at 0 / 6 / 12.5 / 19 / 25 / 37.5 %; round 1's one-row-per-thread form, two warps per scheduler, wanted 25 %).
#ifndef DSG_ATTN_POLY_MASK
#define DSG_ATTN_POLY_MASK 0x8080
#endif
constexpr unsigned ATT_POLY_MASK_QUADS = DSG_ATTN_POLY_MASK;
#ifndef DSG_ATTN_POLY_MASK_UNSHIFTED
#define DSG_ATTN_POLY_MASK_UNSHIFTED 0xA4A4
#endif
constexpr unsigned ATT_POLY_MASK_UNSHIFTED = DSG_ATTN_POLY_MASK_UNSHIFTED;   // the same for the kernel without row maxima
constexpr int ATT_TILE_BYTES = 128 * ATT_DH * 2;  // 16 KB (Q, K or V tile)

// Work items.  A regular item is a PAIR of 128-query tiles of one (frame, head): both warpgroups share one K/V
// stream.  When the number of query tiles is odd (3601 tokens = 29 tiles) every (frame, head) has one tile left over;
// those are paired ACROSS heads in "dual" tail items: warpgroup t runs the last tile of head 2k + t, and the K/V ring
// carries the two heads' tiles alternately (stage order A0 B0 A1 B1 ...).  Without this the lone tile costs a whole
// item with one warpgroup idle: 15 instead of 14.5 items per head, +3.4 % attention time at 480 px.
// Items are numbered so that the dual item of a head pair directly follows the regular items of its two heads
// (group of 2 * full_pairs + 1 consecutive items): consecutive items run at the same time on neighbouring SMs, so the
// pair's K/V (2 x 0.9 MB) is still in L2 when the tail reads it.  (With all dual items at the end of the list ncu
// showed 886 MB of DRAM reads per launch against 531 MB of qkv: every K/V re-fetched once.)
struct AttItem {
  int bh[2];     // frame * H + head of warpgroup 0 / 1
  int q0[2];     // first query row of warpgroup 0 / 1
  bool act1;     // warpgroup 1 has work
  bool dual;     // two K/V streams
};
__host__ __device__ __forceinline__ AttItem att_decode(const AttnParams& p, int item) {
  AttItem I;
  I.act1 = true;
  I.dual = false;
  if (!p.lone) {                                   // even number of query tiles: regular items only
    const int qp = item % p.full_pairs, bh = item / p.full_pairs;
    I.bh[0] = I.bh[1] = bh;
    I.q0[0] = qp * 2 * 128;
    I.q0[1] = I.q0[0] + 128;
    return I;
  }
  const int G = 2 * p.full_pairs + 1;
  const int grp = item / G, r = item - grp * G;
  const int bh_a = 2 * grp, bh_b = 2 * grp + 1;
  const bool has_b = bh_b < p.B * p.H;             // the last group of an odd number of (frame, head)s has one head
  if (r < p.full_pairs || (has_b && r < 2 * p.full_pairs)) {
    const bool second = r >= p.full_pairs;
    const int qp = second ? r - p.full_pairs : r;
    I.bh[0] = I.bh[1] = second ? bh_b : bh_a;
    I.q0[0] = qp * 2 * 128;
    I.q0[1] = I.q0[0] + 128;
  } else {                                         // the lone last tiles of the group's heads
    I.bh[0] = bh_a;
    I.bh[1] = bh_b;
    I.q0[0] = I.q0[1] = p.full_pairs * 2 * 128;
    I.act1 = has_b;
    I.dual = has_b;
  }
  return I;
}

// (s_empty / p_full take one arrival per softmax THREAD: one arrival per warp - __syncwarp + elected lane - measured 6 %
// slower for the whole kernel: the extra convergence point costs more than the arrivals it saves.)

template <int KV_STAGES>
constexpr size_t attn_smem_bytes() {
  return size_t(2 + 2 * KV_STAGES) * ATT_TILE_BYTES + 1024 + 256;
}

template <int R> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

// TMEM columns (all 512): query tile t owns columns [t*256, t*256 + 256): S_t 128 fp32 columns, then P_t 64 columns (128
// bf16 keys, two per column), then O_t 64 fp32 columns - one base address per query tile, constant offsets from it
constexpr uint32_t ATT_T_COLS = 256, ATT_S_OFF = 0, ATT_P_OFF = 128, ATT_O_OFF = 192;
constexpr float ATT_LOG2E = 1.4426950408889634f;
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;   // log2 units: P stays <= 2^8

// register split (see the setmaxnreg comment in the kernel): 128 * producer + 512 * softmax <= 61440
#ifndef DSG_ATT_QUAD_SOFTMAX_REGS
#define DSG_ATT_QUAD_SOFTMAX_REGS 104
#define DSG_ATT_QUAD_PRODUCER_REGS 56
#endif
constexpr int ATT_QUAD_SOFTMAX_REGS = DSG_ATT_QUAD_SOFTMAX_REGS, ATT_QUAD_PRODUCER_REGS = DSG_ATT_QUAD_PRODUCER_REGS;
static_assert(128 * ATT_QUAD_PRODUCER_REGS + 512 * ATT_QUAD_SOFTMAX_REGS <= 640 * 96, "register pool of the 640-thread CTA");

struct AttBars {
  uint64_t *s_full, *s_empty, *p_full, *pv_done;
};

// 2^(s - m) for two scores (s already in log2 units): MUFU.EX2, or (a fixed share of the pairs) the polynomial on the
// FMA / ALU pipes.  SHIFT = false: the rows' reference maximum is 0 and the scores go to the exponential as they are
// (att_softmax_unshifted).
template <bool SHIFT>
__device__ __forceinline__ float2 att_exp_pair(float s0, float s1, float2 nmb, bool poly) {
  const float2 x = SHIFT ? fadd2(make_float2(s0, s1), nmb) : make_float2(s0, s1);
  return poly ? exp2_poly_x2(x) : make_float2(fast_exp2(x.x), fast_exp2(x.y));
}

// exponentials of one key tile: the thread's 64 scores -> 32 packed bf16 pairs + its share of the two row sums
__device__ __forceinline__ void att_exp_tile(const uint32_t (&sr)[64], uint32_t (&pk)[32], float n0, float n1, float& add0,
                                             float& add1) {
  const float2 nmb0 = make_float2(-n0, -n0), nmb1 = make_float2(-n1, -n1);
  float2 sum0 = make_float2(0.f, 0.f), sum1 = make_float2(0.f, 0.f);
  auto S = [&](int i) { return __uint_as_float(sr[i]); };
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float2 e0 = att_exp_pair<true>(S(4 * k), S(4 * k + 1), nmb0, (ATT_POLY_MASK_QUADS >> ((2 * k) & 15)) & 1);
    const float2 e1 = att_exp_pair<true>(S(4 * k + 2), S(4 * k + 3), nmb1, (ATT_POLY_MASK_QUADS >> ((2 * k + 1) & 15)) & 1);
    sum0 = fadd2(sum0, e0);
    sum1 = fadd2(sum1, e1);
    pk[2 * k] = pack_bf16x2(e0.x, e0.y);
    pk[2 * k + 1] = pack_bf16x2(e1.x, e1.y);
  }
  add0 = sum0.x + sum0.y;
  add1 = sum1.x + sum1.y;
}

// ---------------------------------------------------------------------------------------------------------------
// softmax, 16 rows per warp in the accumulator-fragment layout (warps 4..11 query tile 0, 12..19 tile 1).
// Warp w serves TMEM lanes 32*(w%4) + 16*((w-4)/4 % 2) .. +15.  Thread (r = lane/4, q = lane%4) holds rows r and r+8 of
// those 16 and, of every 8-key group k, the keys 8k + 2q, 8k + 2q + 1: 64 scores per key tile instead of 128, and
// twice as many warps per scheduler to overlap the load / max / barrier phases with the exponentials.
// The packed P words a thread produces (row r or r+8, keys 8k+2q, 8k+2q+1 -> P column 4k+q) are exactly its
// registers of a tcgen05.st 16x128b.x16.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void att_softmax_quads(const AttnParams& p, const uint32_t tmem_base, const AttBars bars,
                                                  const int warp, const int lane, const int num_tiles) {
  const int t = (warp - 4) >> 3;                   // query tile of this warp
  const int rbase = (warp & 3) * 32 + (((warp - 4) >> 2) & 1) * 16;   // first of the warp's 16 rows (= TMEM lanes)
  const int qd = lane & 3;                         // this thread's rows: rbase + lane / 4 and rbase + lane / 4 + 8
  // one TMEM base per warp (its 16 lanes, its query tile's columns); S / P / O are constant offsets from it
  const uint32_t t_addr = tmem_base + (uint32_t(rbase) << 16) + uint32_t(t) * ATT_T_COLS;
  const uint32_t s_addr = t_addr + ATT_S_OFF, p_addr = t_addr + ATT_P_OFF, o_addr = t_addr + ATT_O_OFF;
  // shared-memory address of this query tile's barriers, computed once (see mbar_wait_a): s_full[2], s_empty[2],
  // p_full[2], pv_done[2] are consecutive, so the four barriers of tile t are constant offsets from one register
  const uint32_t b_s_full = smem_u32(&bars.s_full[t]);
  const uint32_t b_s_empty = b_s_full + 16, b_p_full = b_s_full + 32, b_pv_done = b_s_full + 48;
  uint32_t sc = 0;                                 // key tiles processed so far by this query tile

  for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
    const AttItem I = att_decode(p, item);
    if (t == 1 && !I.act1) continue;               // unpaired tail item
    const int bh = t ? I.bh[1] : I.bh[0];
    const int h = bh % p.H, b = bh / p.H;
    const int q0 = t ? I.q0[1] : I.q0[0];
    float m0 = 0.f, m1 = 0.f;                      // reference maxima of this thread's two rows (log2 units)
    float l0 = 0.f, l1 = 0.f;                      // this thread's share of the running row sums of exp(s - m)

    for (int j = 0; j < num_tiles; ++j, ++sc) {
      mbar_wait_a(b_s_full, sc & 1);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_16x256b_x16(s_addr, sr);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_a(b_s_empty);               // the tensor pipe may overwrite S_t with the next scores
      const int kbase = j * ATT_BN;
      if (kbase + ATT_BN > p.N) {                  // last key tile of a ragged sequence: keys >= N do not exist
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int key = kbase + 8 * k + 2 * qd;
          if (key >= p.N) sr[4 * k] = sr[4 * k + 2] = 0xff800000u;          // -inf
          if (key + 1 >= p.N) sr[4 * k + 1] = sr[4 * k + 3] = 0xff800000u;
        }
      }
      auto S = [&](int i) { return __uint_as_float(sr[i]); };
      float mx0 = fmaxf(S(0), S(1)), mx1 = fmaxf(S(2), S(3));
#pragma unroll
      for (int k = 1; k < 16; ++k) {
        mx0 = fmaxf(mx0, fmaxf(S(4 * k), S(4 * k + 1)));
        mx1 = fmaxf(mx1, fmaxf(S(4 * k + 2), S(4 * k + 3)));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      // lazy rescale: the reference maximum of a row only moves when the row max has grown by more than 2^8
      const bool g0 = j == 0 || mx0 - m0 > ATT_RESCALE_THRESHOLD;
      const bool g1 = j == 0 || mx1 - m1 > ATT_RESCALE_THRESHOLD;
      const float n0 = g0 ? mx0 : m0, n1 = g1 ? mx1 : m1;
      uint32_t pk[32];
      float add0, add1;
      att_exp_tile(sr, pk, n0, n1, add0, add1);
      if (j > 0) {
        // PV_t(j-1) must have retired before P_t is overwritten or O_t is rescaled
        mbar_wait_a(b_pv_done, (sc - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, g0 || g1)) {
          const float a0 = fast_exp2(m0 - n0), a1 = fast_exp2(m1 - n1);   // 1 for rows that keep m
          uint32_t o[32];
          tmem_ld_16x256b_x8(o_addr, o);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            o[4 * k] = __float_as_uint(__uint_as_float(o[4 * k]) * a0);
            o[4 * k + 1] = __float_as_uint(__uint_as_float(o[4 * k + 1]) * a0);
            o[4 * k + 2] = __float_as_uint(__uint_as_float(o[4 * k + 2]) * a1);
            o[4 * k + 3] = __float_as_uint(__uint_as_float(o[4 * k + 3]) * a1);
          }
          tmem_st_16x256b_x8(o_addr, o);
          l0 *= a0;
          l1 *= a1;
        }
      }
      m0 = n0;
      m1 = n1;
      l0 += add0;
      l1 += add1;
      tmem_st_16x128b_x16(p_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(b_p_full);
    }

    // epilogue: O / l -> bf16 -> out[b*N + q, h*64 + d].  Row sums: add the quad's four shares; the O rows are read
    // as 16x32bx2 (lane i and i+16: row i % 16, columns 0..31 / 32..63), so each thread stores 64 contiguous bytes
    mbar_wait_a(b_pv_done, (sc - 1) & 1);
    tc_fence_after();
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int r = lane & 15, ch = lane >> 4;
    const float la = __shfl_sync(0xffffffffu, l0, 4 * (r & 7)), lb = __shfl_sync(0xffffffffu, l1, 4 * (r & 7));
    const float inv_l = 1.0f / (r < 8 ? la : lb);
    uint32_t o[32];
    tmem_ld_16x32bx2_x32(o_addr, o);
    tmem_ld_wait();
    const int q = q0 + rbase + r;
    if (q < p.N) {
      __nv_bfloat16* dst = p.out + (size_t(b) * p.N + q) * p.D + h * ATT_DH + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + i) = v;
      }
    }
    __syncwarp();
    // O_t is free again once every thread's tcgen05.ld has completed (wait::ld above); the next item's first PV_t is
    // ordered after this query tile's next p_full arrivals.
    tc_fence_before();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same softmax WITHOUT any row maximum (attn_fwd_kernel<.., true>): every row keeps the reference 0, P = 2^S as the
// tensor pipe wrote S.  No row-max pass, no shuffles, no subtraction: 2 MUFU + 1 packed add + 1 pack per pair of scores.
// The scores are pulled from TMEM in four 32-column chunks, each processed while the next one is in flight: the
// tcgen05.wait::ld between them keeps the instruction scheduler from hoisting all exponentials to the front (a warp
// that has queued a burst of MUFU.EX2 cannot issue its FMA-pipe work behind them; with the chunks the two interleave).
// Valid as long as every row sum stays inside [2^-100, 2^100] (row maximum within about +-88 log2 units, i.e. logits
// within +-60): checked per row at the end of the work item.  A row outside raises *range_flag and the launch is
// redone by the shifted kernel, which is enqueued behind this one and otherwise returns at once.
// ---------------------------------------------------------------------------------------------------------------
template <int CH, bool MASK>
__device__ __forceinline__ void att_exp_chunk(uint32_t (&c)[16], uint32_t* pk, float2& sum0, float2& sum1, int key0, int N) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (MASK) {                                    // last key tile of a ragged sequence: keys >= N do not exist
      const int key = key0 + 8 * k;
      if (key >= N) c[4 * k] = c[4 * k + 2] = 0xff800000u;          // -inf
      if (key + 1 >= N) c[4 * k + 1] = c[4 * k + 3] = 0xff800000u;
    }
    constexpr float2 z = {0.f, 0.f};
    const int kk = 4 * CH + k;
    const float2 e0 = att_exp_pair<false>(__uint_as_float(c[4 * k]), __uint_as_float(c[4 * k + 1]), z,
                                          (ATT_POLY_MASK_UNSHIFTED >> ((2 * kk) & 15)) & 1);
    const float2 e1 = att_exp_pair<false>(__uint_as_float(c[4 * k + 2]), __uint_as_float(c[4 * k + 3]), z,
                                          (ATT_POLY_MASK_UNSHIFTED >> ((2 * kk + 1) & 15)) & 1);
    sum0 = fadd2(sum0, e0);
    sum1 = fadd2(sum1, e1);
    pk[2 * k] = pack_bf16x2(e0.x, e0.y);
    pk[2 * k + 1] = pack_bf16x2(e1.x, e1.y);
  }
}

// a full key tile: four chunks, each processed while the next one is in flight
__device__ __forceinline__ void att_exp_tile_chunked(uint32_t s_addr, uint32_t b_s_empty, uint32_t (&pk)[32], float2& sum0,
                                                     float2& sum1) {
  uint32_t ca[16], cb[16];
  tmem_ld_16x256b_x4(s_addr, ca);
  tmem_ld_wait();
  tmem_ld_16x256b_x4(s_addr + 32, cb);
  att_exp_chunk<0, false>(ca, pk, sum0, sum1, 0, 0);
  tmem_ld_wait();
  tmem_ld_16x256b_x4(s_addr + 64, ca);
  att_exp_chunk<1, false>(cb, pk + 8, sum0, sum1, 0, 0);
  tmem_ld_wait();
  tmem_ld_16x256b_x4(s_addr + 96, cb);
  att_exp_chunk<2, false>(ca, pk + 16, sum0, sum1, 0, 0);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive_a(b_s_empty);                        // the tensor pipe may overwrite S_t with the next scores
  att_exp_chunk<3, false>(cb, pk + 24, sum0, sum1, 0, 0);
}

// The last key tile of a ragged sequence: only the 32-key chunks that hold existing keys are loaded and exponentiated
// (3601 tokens: 17 keys = one chunk of four); the PV MMA of this tile reads only the K steps that cover them
// (att_pv_ksteps), so the rest of P is never looked at.
__device__ __forceinline__ void att_exp_tile_ragged(uint32_t s_addr, uint32_t b_s_empty, uint32_t (&pk)[32], float2& sum0,
                                                    float2& sum1, int key0, int N, int nvc) {
  uint32_t ca[16], cb[16];
#pragma unroll
  for (int i = 8; i < 32; ++i) pk[i] = 0u;
  tmem_ld_16x256b_x4(s_addr, ca);
  if (nvc > 1) tmem_ld_16x256b_x4(s_addr + 32, cb);
  tmem_ld_wait();
  att_exp_chunk<0, true>(ca, pk, sum0, sum1, key0, N);
  if (nvc > 2) tmem_ld_16x256b_x4(s_addr + 64, ca);
  if (nvc > 1) att_exp_chunk<1, true>(cb, pk + 8, sum0, sum1, key0 + 32, N);
  if (nvc > 3) tmem_ld_16x256b_x4(s_addr + 96, cb);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive_a(b_s_empty);
  if (nvc > 2) att_exp_chunk<2, true>(ca, pk + 16, sum0, sum1, key0 + 64, N);
  if (nvc > 3) att_exp_chunk<3, true>(cb, pk + 24, sum0, sum1, key0 + 96, N);
}

// K = 16 steps of the PV MMA of key tile j: all eight, or, in the ragged last tile, those that cover existing keys
__host__ __device__ __forceinline__ int att_pv_ksteps(int N, int j) {
  const int left = N - j * ATT_BN;
  return left >= ATT_BN ? ATT_BN / 16 : (left + 15) / 16;
}

__device__ __forceinline__ void att_softmax_unshifted(const AttnParams& p, const uint32_t tmem_base, const AttBars bars,
                                                      const int warp, const int lane, const int num_tiles) {
  const int t = (warp - 4) >> 3;
  const int rbase = (warp & 3) * 32 + (((warp - 4) >> 2) & 1) * 16;
  const int qd = lane & 3;
  const uint32_t t_addr = tmem_base + (uint32_t(rbase) << 16) + uint32_t(t) * ATT_T_COLS;
  const uint32_t s_addr = t_addr + ATT_S_OFF, p_addr = t_addr + ATT_P_OFF, o_addr = t_addr + ATT_O_OFF;
  const uint32_t b_s_full = smem_u32(&bars.s_full[t]);
  const uint32_t b_s_empty = b_s_full + 16, b_p_full = b_s_full + 32, b_pv_done = b_s_full + 48;
  uint32_t sc = 0;
  bool out_of_range = false;

  for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
    const AttItem I = att_decode(p, item);
    if (t == 1 && !I.act1) continue;
    const int bh = t ? I.bh[1] : I.bh[0];
    const int h = bh % p.H, b = bh / p.H;
    const int q0 = t ? I.q0[1] : I.q0[0];
    float2 sum0 = make_float2(0.f, 0.f), sum1 = make_float2(0.f, 0.f);   // this thread's share of its two row sums

    if (q0 + rbase >= p.N) {
      // None of this warp's 16 query rows exists (last query tile of a ragged sequence: 17 of 128 rows at 3601 tokens).
      // Nothing to compute - whatever the PV MMA makes of these rows of P is never stored - but the barriers count
      // every thread of the query tile, phase by phase: s_empty after s_full (QK(j) is only issued once the previous
      // s_empty phase is complete), p_full after PV(j-1) has retired (which implies the previous p_full phase is complete).
      for (int j = 0; j < num_tiles; ++j, ++sc) {
        mbar_wait_a(b_s_full, sc & 1);
        mbar_arrive_a(b_s_empty);
        if (j > 0) mbar_wait_a(b_pv_done, (sc - 1) & 1);
        mbar_arrive_a(b_p_full);
      }
      mbar_wait_a(b_pv_done, (sc - 1) & 1);
      continue;
    }

    for (int j = 0; j < num_tiles; ++j, ++sc) {
      mbar_wait_a(b_s_full, sc & 1);
      tc_fence_after();
      uint32_t pk[32];
      const int kbase = j * ATT_BN;
      if (kbase + ATT_BN > p.N)
        att_exp_tile_ragged(s_addr, b_s_empty, pk, sum0, sum1, kbase + 2 * qd, p.N, (p.N - kbase + 31) >> 5);
      else att_exp_tile_chunked(s_addr, b_s_empty, pk, sum0, sum1);
      if (j > 0) {
        mbar_wait_a(b_pv_done, (sc - 1) & 1);      // PV_t(j-1) must have retired before P_t is overwritten
        tc_fence_after();
      }
      tmem_st_16x128b_x16(p_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(b_p_full);
    }

    mbar_wait_a(b_pv_done, (sc - 1) & 1);
    tc_fence_after();
    float l0 = sum0.x + sum0.y, l1 = sum1.x + sum1.y;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // (written so that a NaN sum counts as out of range)
    out_of_range |= !(l0 >= 0x1p-100f && l0 <= 0x1p100f) || !(l1 >= 0x1p-100f && l1 <= 0x1p100f);
    const int r = lane & 15, ch = lane >> 4;
    const float la = __shfl_sync(0xffffffffu, l0, 4 * (r & 7)), lb = __shfl_sync(0xffffffffu, l1, 4 * (r & 7));
    const float inv_l = 1.0f / (r < 8 ? la : lb);
    uint32_t o[32];
    tmem_ld_16x32bx2_x32(o_addr, o);
    tmem_ld_wait();
    const int q = q0 + rbase + r;
    if (q < p.N) {
      __nv_bfloat16* dst = p.out + (size_t(b) * p.N + q) * p.D + h * ATT_DH + ch * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + i) = v;
      }
    }
    __syncwarp();
    tc_fence_before();
  }
  if (out_of_range && p.range_flag != nullptr) *reinterpret_cast<volatile int*>(p.range_flag) = p.launch_id;
}

template <int KV_STAGES, bool UNSHIFTED>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t ATT_ARRIVALS = 32 * ATT_SMW;   // s_empty / p_full: one arrival per softmax thread of the query tile

  // The shifted kernel doubles as the redo of an unshifted launch whose scores left the range (att_softmax_unshifted):
  // it is enqueued behind it with the same launch_id and has nothing to do unless that launch raised the flag.
  if (!UNSHIFTED && p.range_flag != nullptr && *reinterpret_cast<volatile int*>(p.range_flag) != p.launch_id) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sQ = smem;                          // [2][16 KB]
  uint8_t* sKV = smem + 2 * ATT_TILE_BYTES;    // [stage][K 16 KB | V 16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + size_t(2 + 2 * KV_STAGES) * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                     // 1
  uint64_t* q_empty = bars + 1;                // 1
  uint64_t* kv_full = bars + 2;                // KV_STAGES
  uint64_t* kv_empty = kv_full + KV_STAGES;    // KV_STAGES
  uint64_t* s_full = kv_empty + KV_STAGES;     // 2: S_t(j) written by the tensor pipe
  uint64_t* s_empty = s_full + 2;              // 2 (ATT_ARRIVALS each): S_t(j) is in registers
  uint64_t* p_full = s_empty + 2;              // 2 (ATT_ARRIVALS each): P_t(j) written, O_t rescaled
  uint64_t* pv_done = p_full + 2;              // 2: PV_t(j) retired (P_t free again, O_t stable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int num_tiles = (p.N + ATT_BN - 1) / ATT_BN;

  constexpr int HB_CODE = 300 + ATT_SMW + (UNSHIFTED ? 10 : 0);
  hb_mark(p.hb, HB_CODE, 1);
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);                       // one commit per MMA issuer
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_empty[t], ATT_ARRIVALS);
      mbar_init(&p_full[t], ATT_ARRIVALS);
      mbar_init(&pv_done[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  hb_mark(p.hb, HB_CODE, 2);

  if (warp < 4) {
    // Registers move from the producer warpgroup to the softmax warps.  The budget is the CTA's register pool AT LAUNCH
    // (threads x the kernel's register count: 640 x 96 = 61440 >= 128*56 + 512*104), not the SM's 64 K: a
    // setmaxnreg.inc beyond the pool blocks for ever.
    setmaxnreg_dec<ATT_QUAD_PRODUCER_REGS>();
    if (warp == 0 && elect_one()) {
      // ------------------------------ TMA producer ------------------------------
      uint32_t kvc = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const AttItem I = att_decode(p, item);
        const int h0 = I.bh[0] % p.H, b0 = I.bh[0] / p.H;
        const int h1 = I.bh[1] % p.H, b1 = I.bh[1] / p.H;
        if (it > 0) mbar_wait(q_empty, (it - 1) & 1);
        mbar_expect_tx(q_full, I.act1 ? 2 * ATT_TILE_BYTES : ATT_TILE_BYTES);
        tma_load_3d(sQ, &tmQKV, q_full, h0 * ATT_DH, I.q0[0], b0);
        if (I.act1) tma_load_3d(sQ + ATT_TILE_BYTES, &tmQKV, q_full, h1 * ATT_DH, I.q0[1], b1);
        const int streams = I.dual ? 2 : 1;
        for (int j = 0; j < num_tiles; ++j) {
          for (int u = 0; u < streams; ++u, ++kvc) {
            const int h = u ? h1 : h0, b = u ? b1 : b0;
            const int s = kvc % KV_STAGES;
            mbar_wait(&kv_empty[s], ((kvc / KV_STAGES) & 1) ^ 1);
            mbar_expect_tx(&kv_full[s], 2 * ATT_TILE_BYTES);
            uint8_t* sk = sKV + size_t(s) * 2 * ATT_TILE_BYTES;
            tma_load_3d(sk, &tmQKV, &kv_full[s], p.D + h * ATT_DH, j * ATT_BN, b);
            tma_load_3d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[s], 2 * p.D + h * ATT_DH, j * ATT_BN, b);
          }
        }
      }
    } else if ((warp == 1 || warp == 2) && elect_one()) {
      // ------------------------------ MMA issuers ------------------------------
      // one issuing thread per query tile (warp 1: tile 0, warp 2: tile 1): with a single in-order issuer the
      // PV of one tile waits behind barrier waits that belong to the other tile
      const int t = warp - 1;
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0);  // K^T: K-major B
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_DH, 1);  // V  : MN-major B
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ + size_t(t) * ATT_TILE_BYTES));
      const uint32_t s_tmem = tmem_base + uint32_t(t) * ATT_T_COLS + ATT_S_OFF;
      const uint32_t p_tmem = tmem_base + uint32_t(t) * ATT_T_COLS + ATT_P_OFF;
      const uint32_t o_tmem = tmem_base + uint32_t(t) * ATT_T_COLS + ATT_O_OFF;
      uint32_t kvc = 0, ct = 0;           // ct: key tiles of this query tile processed so far (barrier phases)
      int it = 0;
#ifdef DSG_ATTN_TIMING
      long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long tprev = clock64();
#endif
      auto issue_qk = [&](uint32_t kv_counter) {
        const uint32_t sk = smem_u32(sKV + size_t(kv_counter % KV_STAGES) * 2 * ATT_TILE_BYTES);
        const uint64_t kdesc = umma_desc_sw128(sk);
#pragma unroll
        for (int k = 0; k < ATT_DH / 16; ++k)
          umma_ss(s_tmem, qdesc + uint64_t(k * 2), kdesc + uint64_t(k * 2), idesc_qk, k != 0);
        tc_commit(&s_full[t]);
      };
      auto issue_pv = [&](uint32_t kv_counter, bool accumulate, int ksteps) {
        const uint32_t sv = smem_u32(sKV + size_t(kv_counter % KV_STAGES) * 2 * ATT_TILE_BYTES) + ATT_TILE_BYTES;
        const uint64_t vdesc = umma_desc_sw128(sv);
        // P: 8 TMEM columns (16 bf16 keys) per K step; V: 16 keys = 2 KB per K step
        // (the ragged last key tile: only the K steps that cover existing keys; the softmax warps of the kernel without
        // row maxima do not produce the rest of P, the classic kernel writes zeros there)
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k)
          if (k < ksteps)
            umma_ts(o_tmem, p_tmem + uint32_t(k * 8), vdesc + uint64_t(k * 128), idesc_pv, (accumulate || k != 0) ? 1u : 0u);
        tc_commit(&pv_done[t]);
      };
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const AttItem I = att_decode(p, item);
        const bool active = t == 0 || I.act1;      // warpgroup 1 of an unpaired tail item has no tile
        // ring position of this issuer's K/V tile j; in a dual item the other head's tiles sit in between
        const uint32_t stride = I.dual ? 2u : 1u, off = I.dual ? uint32_t(t) : 0u;
        auto own = [&](int j) { return kvc + uint32_t(j) * stride + off; };
        mbar_wait(q_full, it & 1);
        mbar_wait(&kv_full[own(0) % KV_STAGES], (own(0) / KV_STAGES) & 1);
        tc_fence_after();
        if (active) issue_qk(own(0));
        if (num_tiles == 1) tc_commit(q_empty);
        for (int j = 0; j < num_tiles; ++j) {
          const bool more = j + 1 < num_tiles;
          if (I.dual) {
            // the other head's stage of this step is never read here: release it as soon as it has been filled (the
            // wait keeps this arrival in the right phase of kv_empty)
            const uint32_t o = kvc + uint32_t(j) * 2u + uint32_t(1 - t);
            mbar_wait(&kv_full[o % KV_STAGES], (o / KV_STAGES) & 1);
            mbar_arrive(&kv_empty[o % KV_STAGES]);
          }
          if (more) {
            // next scores as soon as the softmax warps hold the current ones in registers
            ATT_T(7);
            mbar_wait(&kv_full[own(j + 1) % KV_STAGES], (own(j + 1) / KV_STAGES) & 1);
            ATT_T(0);
            if (active) {
              mbar_wait(&s_empty[t], ct & 1);
              ATT_T(1);
              tc_fence_after();
              issue_qk(own(j + 1));
              ATT_T(6);
            }
            if (j + 2 == num_tiles) tc_commit(q_empty);  // this tile's last read of Q has been issued
          }
          if (active) {
            ATT_T(7);
            mbar_wait(&p_full[t], ct & 1); ++ct;
            ATT_T(3);
            tc_fence_after();
            issue_pv(own(j), j != 0, att_pv_ksteps(p.N, j));
            ATT_T(6);
          }
          tc_commit(&kv_empty[own(j) % KV_STAGES]);   // this tile's reads of K(j), V(j) have been issued
        }
        kvc += uint32_t(num_tiles) * stride;
      }
#ifdef DSG_ATTN_TIMING
      if (p.timing != nullptr && t == 0)
        for (int i = 0; i < 8; ++i) p.timing[(size_t(gridDim.x) * 2 + blockIdx.x) * 8 + i] = tacc[i];
#endif
    }
  } else {
    // ------------------------------ softmax warps ------------------------------
    const AttBars ab{s_full, s_empty, p_full, pv_done};
    setmaxnreg_inc<ATT_QUAD_SOFTMAX_REGS>();
    if (UNSHIFTED) att_softmax_unshifted(p, tmem_base, ab, warp, lane, num_tiles);
    else att_softmax_quads(p, tmem_base, ab, warp, lane, num_tiles);
  }

  hb_mark(p.hb, HB_CODE, 3);
  tc_fence_before();
  __syncthreads();
  hb_mark(p.hb, HB_CODE, 4);
  if (warp == 1) {
    tmem_dealloc(tmem_base, TMEM_COLS);
    hb_mark_left(p.hb, HB_CODE);                     // by the deallocating warp: a CTA stuck in dealloc stays visible
  }
}

}  // namespace dsg
