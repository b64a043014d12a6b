// tcgen05.mma issue/execution rate on B200: one CTA per SM, one thread issues REPS MMAs of a given
// shape back to back (operands = whatever is in shared memory / TMEM), then commits and waits.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dino_b200/csrc -o tools/ubench/mma_bench tools/ubench/mma_bench.cu
#include <cstdio>
#include <vector>
#include <algorithm>
#include "ptx.cuh"
using namespace dsg;

__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS (A,B smem K-major); mode 1: TS (A from TMEM, B smem MN-major)
__global__ void __launch_bounds__(128, 1) mma_kernel(long long* cycles, int M, int N, int mode, int reps, int nstage_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 196608 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(M, N, mode == 1 ? 1 : 0);
    const uint32_t sa = smem_u32(smem);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // rotate through several smem stages like a real mainloop (4 k-steps per 64-wide stage)
      const uint32_t st = sa + uint32_t((r >> 2) % 4) * uint32_t(nstage_bytes);
      const uint64_t adesc = umma_desc_sw128(st) + uint64_t((r & 3) * 2);
      const uint64_t bdesc = umma_desc_sw128(st + 16384) + uint64_t(mode == 1 ? (r & 3) * 128 : (r & 3) * 2);
      if (mode == 0) umma_ss(tmem + uint32_t((r >> 5) & 1) * 256, adesc, bdesc, idesc, 1);
      else umma_ts(tmem + 384, tmem + uint32_t((r & 7) * 8), bdesc, idesc, 1);
    }
    tc_commit(&bar);
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    cycles[blockIdx.x * 2] = t1 - t0;
    cycles[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, sizeof(long long) * 148 * 2);
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int M, N, mode; const char* name; } cfgs[] = {
      {128, 64, 0, "SS M128 N64"},   {128, 128, 0, "SS M128 N128"}, {128, 192, 0, "SS M128 N192"},
      {128, 256, 0, "SS M128 N256"}, {128, 64, 1, "TS M128 N64 (A in TMEM, B MN-major)"},
      {128, 128, 1, "TS M128 N128"}, {64, 128, 0, "SS M64 N128"}, {64, 256, 0, "SS M64 N256"}};
  for (auto& c : cfgs) {
    const int reps = 2048;
    mma_kernel<<<148, 128, 197 * 1024>>>(cyc, c.M, c.N, c.mode, reps, 16384 + c.N * 128);
    cudaDeviceSynchronize();
    mma_kernel<<<148, 128, 197 * 1024>>>(cyc, c.M, c.N, c.mode, reps, 16384 + c.N * 128);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(296);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * 296, cudaMemcpyDeviceToHost);
    std::vector<long long> iss, tot;
    for (int i = 0; i < 148; ++i) { iss.push_back(h[2 * i]); tot.push_back(h[2 * i + 1]); }
    std::sort(iss.begin(), iss.end()); std::sort(tot.begin(), tot.end());
    const double per = double(tot[74]) / reps;
    const double flop_clk = 2.0 * c.M * c.N * 16 / per;
    printf("%-40s issue %.1f clk/MMA, complete %.1f clk/MMA -> %.0f flop/clk/SM (%.0f%% of 8192) [%s]\n", c.name,
           double(iss[74]) / reps, per, flop_clk, flop_clk / 81.92, cudaGetErrorString(e));
  }
  return 0;
}
