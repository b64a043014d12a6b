// Micro-benchmarks that decide the attention-kernel design on B200 (sm_100a):
//   per-SM throughput of MUFU.EX2 in f32 / f16x2 / bf16x2 form, the pack conversions, FFMA,
//   and tcgen05.ld.  One block per SM, clock64() inside the kernel -> ops per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench tools/ubench/ubench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define ITERS 2048
#define ILP 8

template <int MODE>
__global__ void mufu_kernel(float* out, long long* cycles, float seed) {
  float f[ILP];
  uint32_t u[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { f[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; u[i] = 0x3c003800u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(seed), "f"(f[(i + 1) % ILP]));
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, 0f3F8147AE, 0f3DCCCCCD;" : "+f"(f[i]));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) % ILP]));
      if (MODE == 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) % ILP]));
      if (MODE == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (MODE == 8) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) % ILP]));
      if (MODE == 9) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]));
      if (MODE == 10) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % ILP]));
      if (MODE == 11) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(*reinterpret_cast<unsigned long long*>(&f[i & ~1])) : "l"(*reinterpret_cast<unsigned long long*>(&f[(i + 2) % ILP & ~1])));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += f[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the softmax inner loop of the attention kernel: per pair FFMA2, 2 x MUFU.EX2, FADD2, F2FP
__global__ void softmax_body_kernel(float* out, long long* cycles, float seed, int poly_every) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
  unsigned long long sum01 = 0ull, sum23 = 0ull;
  const float2 l2e2 = make_float2(1.4426950408889634f, 1.4426950408889634f);
  const float2 nmb2 = make_float2(-seed, -seed);
  const unsigned long long l2e = *reinterpret_cast<const unsigned long long*>(&l2e2);
  const unsigned long long nmb = *reinterpret_cast<const unsigned long long*>(&nmb2);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      unsigned long long x, e;
      float2 sv = make_float2(v[2 * i], v[2 * i + 1]);
      asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(*reinterpret_cast<unsigned long long*>(&sv)), "l"(l2e), "l"(nmb));
      float2 xf = *reinterpret_cast<float2*>(&x);
      float e0, e1;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(xf.x));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(xf.y));
      float2 ef = make_float2(e0, e1);
      e = *reinterpret_cast<unsigned long long*>(&ef);
      if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(sum23) : "l"(e));
      else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(sum01) : "l"(e));
      uint32_t pk;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e1), "f"(e0));
      acc ^= pk;
      v[2 * i] = e0 - 1.5f;   // keep a loop-carried dependence so nothing is hoisted
    }
  }
  long long t1 = clock64();
  float2 a = *reinterpret_cast<float2*>(&sum01), b = *reinterpret_cast<float2*>(&sum23);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y + b.x + b.y + __uint_as_float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// softmax body variant 2: optional tcgen05.st of the packed probabilities and polynomial exp2 share
__device__ __forceinline__ float2 ub_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ub_fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 ub_poly(float2 x) {
  x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
  const float2 t = ub_fadd2(x, make_float2(12582912.f, 12582912.f));
  const float2 n = ub_fadd2(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = ub_ffma2(n, make_float2(-1.f, -1.f), x);
  float2 p = ub_ffma2(make_float2(0.0550886f, 0.0550886f), f, make_float2(0.242604f, 0.242604f));
  p = ub_ffma2(p, f, make_float2(0.693276f, 0.693276f));
  p = ub_ffma2(p, f, make_float2(0.999929f, 0.999929f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}
template <unsigned POLY, bool DO_ST>
__global__ void softmax_body2_kernel(float* out, long long* cycles, float seed) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
  float2 sum01 = make_float2(0.f, 0.f), sum23 = make_float2(0.f, 0.f);
  const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
  const float2 nmb = make_float2(-seed, -seed);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS / 4; ++it) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 x = ub_ffma2(make_float2(v[2 * i], v[2 * i + 1]), l2e, nmb);
      float2 e;
      if ((POLY >> i) & 1) e = ub_poly(x);
      else {
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(x.x));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(x.y));
      }
      if (i & 1) sum23 = ub_fadd2(sum23, e); else sum01 = ub_fadd2(sum01, e);
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk[i]) : "f"(e.y), "f"(e.x));
      v[2 * i] = e.x - 1.5f;
    }
    if (DO_ST) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
          "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + (it & 1) * 16),
          "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]), "r"(pk[8]),
          "r"(pk[9]), "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15]) : "memory");
    } else {
      v[1] += __uint_as_float(pk[0] ^ pk[5] ^ pk[10] ^ pk[15]) * 1e-30f;
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum01.x + sum01.y + sum23.x + sum23.y + v[1];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <unsigned POLY, bool DO_ST>
void run_body2(const char* name) {
  for (int w : {4, 8, 16}) {
    int nsm = 148;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * nsm * w * 32);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    softmax_body2_kernel<POLY, DO_ST><<<nsm, w * 32>>>(out, cyc, 0.37f);
    cudaDeviceSynchronize();
    softmax_body2_kernel<POLY, DO_ST><<<nsm, w * 32>>>(out, cyc, 0.37f);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double elems = double(w) * 32 * (ITERS / 4) * 32;
    printf("%-44s warps=%2d: %.2f exps/clk/SM (%s)\n", name, w, elems / double(h[nsm / 2]), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
  }
}

// tcgen05.ld throughput: W warps (W multiple of 4) each repeatedly load 32 lanes x 32 columns
__global__ void tmem_ld_kernel(float* out, long long* cycles, int reps, int x64) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
            "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
            "=r"(r[30]), "=r"(r[31])
          : "r"(base + (uint32_t)(c * 32 + (it & 1) * 128))
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int MODE>
void run(const char* name, int warps, double elems_per_op) {
  int nsm = 148;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * nsm * warps * 32);
  cudaMalloc(&cyc, sizeof(long long) * nsm);
  mufu_kernel<MODE><<<nsm, warps * 32>>>(out, cyc, 0.37f);
  cudaDeviceSynchronize();
  mufu_kernel<MODE><<<nsm, warps * 32>>>(out, cyc, 0.37f);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  double med = double(h[nsm / 2]);
  double ops = double(warps) * 32 * ITERS * ILP;
  printf("%-28s warps=%2d  %8.2f inst-lanes/clk/SM  %8.2f elems/clk/SM  (%s)\n", name, warps, ops / med,
         ops * elems_per_op / med, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("ex2.approx.ftz.f32", w, 1);
    run<1>("ex2.approx.f16x2", w, 2);
    run<2>("ex2.approx.ftz.bf16x2", w, 2);
    run<3>("fma.f32 (3 reg)", w, 1);
    run<4>("fma.f32 (imm)", w, 1);
    run<5>("cvt.rn.bf16x2.f32", w, 2);
    run<6>("cvt.rn.f16x2.f32", w, 2);
    run<7>("add.f32", w, 1);
    run<8>("max.f32", w, 1);
    run<9>("fma.f16x2", w, 2);
    run<10>("add.f16x2", w, 2);
    run<11>("fma.f32x2", w, 2);
  }
  run_body2<0x0000, false>("body2: all MUFU, no st");
  run_body2<0x0000, true>("body2: all MUFU, tcgen05.st");
  run_body2<0x8888, false>("body2: 25% poly, no st");
  run_body2<0x8888, true>("body2: 25% poly, tcgen05.st");
  run_body2<0xAAAA, true>("body2: 50% poly, tcgen05.st");
  for (int w : {4, 8, 12, 16}) {
    int nsm = 148;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * nsm * w * 32);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    softmax_body_kernel<<<nsm, w * 32>>>(out, cyc, 0.37f, 0);
    cudaDeviceSynchronize();
    softmax_body_kernel<<<nsm, w * 32>>>(out, cyc, 0.37f, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double elems = double(w) * 32 * (ITERS / 4) * 32;
    printf("softmax body (FFMA2+2MUFU+FADD2+F2FP per pair) warps=%2d: %.2f exps/clk/SM (%s)\n", w, elems / double(h[nsm / 2]),
           cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
  }
  for (int w : {4, 8}) {
    int nsm = 148, reps = 2000;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * nsm * w * 32);
    cudaMalloc(&cyc, sizeof(long long) * nsm);
    tmem_ld_kernel<<<nsm, w * 32>>>(out, cyc, reps, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double bytes = double(w) * reps * 4 * 4096.0;
    printf("tcgen05.ld 32x32b.x32 (ld+wait serial) warps=%d: %.1f B/clk/SM, %.1f clk per x32 load per warp (%s)\n", w,
           bytes / double(h[nsm / 2]), double(h[nsm / 2]) / (reps * 4.0), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
  }
  return 0;
}
