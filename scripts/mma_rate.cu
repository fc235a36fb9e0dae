// Microbenchmark: issue rate of tcgen05.mma on one resident shared-memory stage (no TMA, no epilogue in the loop).
//   kind::i8 vs kind::f8f6f4, cta_group::1 (M=128) vs cta_group::2 (M=256 per CTA pair), N = 256.
// Answers VERDICT r1 item 3(i): is the ~300 clk per M128 N256 K32 MMA seen inside gram_u8_umma_kernel the rate of
// kind::i8 itself, or an operand-feed limit of that kernel?  Every CTA (or CTA pair) loops over the SAME smem stage,
// so the tensor pipe is the only consumer.  The accumulators are checked at the end (A = B = 1 -> D = K total).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/mma_rate scripts/mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; spins < (1u << 22); ++spins) if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t dst, uint32_t cols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int CG> __device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// cta_group::2 commit that arrives on the ISSUING CTA's barrier only (no multicast): used for the intermediate drains, so
// that the peer CTA cannot fall a barrier phase behind (the first version multicast every commit; the peer then missed
// phases and reported a failed check although every MMA had run: profiles/r2_mma_rate.jsonl, "bad_accumulators")
template <int CG> __device__ __forceinline__ void umma_commit_local(uint32_t bar) {
  if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int CG, int KIND> __device__ __forceinline__ void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 1 && KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  if (CG == 2 && KIND == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  if (CG == 1 && KIND == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  if (CG == 2 && KIND == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// KIND 0: i8 (u8 x u8 -> s32), KIND 1: f8f6f4 (e4m3 x e4m3 -> f32).  CG: cta_group.  N: MMA N.  M = 128 * CG.
template <int CG, int KIND, int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int mmas_per_commit, unsigned long long* clk_out, int* bad_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  constexpr int kABytes = 128 * 128;               // this CTA's 128 rows of A, 128 bytes of K
  constexpr int kBBytes = (N / CG) * 128;          // this CTA's share of B rows
  uint8_t* sA = base;
  uint8_t* sB = base + kABytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + kABytes + kBBytes);  // [0]: leader-local drains, [1]: final (both CTAs)
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const uint8_t one = KIND == 0 ? 1 : 0x38;        // u8 1, or e4m3 1.0
  for (int i = threadIdx.x; i < kABytes + kBBytes; i += blockDim.x) base[i] = one;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_init(smem_u32(bar + 1), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  if (warp == 0) tmem_alloc<CG>(smem_u32(slot), 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *slot;
  constexpr uint32_t cfmt = KIND == 0 ? 2u : 1u;
  constexpr uint32_t idesc = (cfmt << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  long long t0 = 0, t1 = 0;
  bool ok = true;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint64_t da = make_desc_sw128(smem_u32(sA)), db = make_desc_sw128(smem_u32(sB));
    // warm-up
    for (int k = 0; k < 4; ++k) umma<CG, KIND>(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit_local<CG>(smem_u32(bar));
    ok &= mbar_wait(smem_u32(bar), 0);
    uint32_t phase = 1;
    t0 = clock64();
    int since = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma<CG, KIND>(tmem_base + (it & 1) * N, da + 2 * k, db + 2 * k, idesc, (it > 1 || k > 0) ? 1u : 0u);
      since += 4;
      if (since >= mmas_per_commit || it == iters - 1) {
        umma_commit_local<CG>(smem_u32(bar));
        ok &= mbar_wait(smem_u32(bar), phase);  // drains the pipe once per `mmas_per_commit` MMAs
        if (!ok) break;                          // a commit that never arrives: give up instead of spinning 1250 times
        phase ^= 1;
        since = 0;
      }
    }
    t1 = clock64();
    umma_commit<CG>(smem_u32(bar + 1));          // everything has completed: tell both CTAs
    ok &= mbar_wait(smem_u32(bar + 1), 0);
    clk_out[blockIdx.x / CG] = (unsigned long long)(t1 - t0);
  } else if (CG == 2 && warp == 1 && lane == 0 && rank == 1) {
    // the peer only waits for the final multicast commit (its accumulators are complete then); generous spin budget
    for (int tries = 0; tries < 64 && !mbar_try_wait(smem_u32(bar + 1), 0); ++tries) mbar_wait(smem_u32(bar + 1), 0);
    ok &= mbar_try_wait(smem_u32(bar + 1), 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // check: accumulator stage 0 got ceil(iters/2) iterations (+ warm-up overwritten by it=0 with accumulate=0 for k=0)
  uint32_t v;
  tmem_ld1(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  const int its0 = (iters + 1) / 2;
  const int expect = its0 * 4 * 32;
  const int got = KIND == 0 ? (int)v : (int)__uint_as_float(v);
  if (got != expect || !ok) atomicAdd(bad_out, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 0) tmem_dealloc<CG>(tmem_base, 512);
}

template <int CG, int KIND, int N>
int run(const char* name, int grid, int iters, int per_commit) {
  unsigned long long* d_clk; int* d_bad;
  CK(cudaMalloc(&d_clk, 256 * 8)); CK(cudaMalloc(&d_bad, 4));
  CK(cudaMemset(d_clk, 0, 256 * 8)); CK(cudaMemset(d_bad, 0, 4));
  size_t smem = 128 * 128 + (N / CG) * 128 + 64 + 1024;  // tiles + 2 barriers + TMEM slot + alignment slack
  CK(cudaFuncSetAttribute(mma_rate_kernel<CG, KIND, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG, KIND, N>, iters, per_commit, d_clk, d_bad));  // warm-up launch
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(d_bad, 0, 4));
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG, KIND, N>, iters, per_commit, d_clk, d_bad));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned long long> clk(256); int bad;
  CK(cudaMemcpy(clk.data(), d_clk, 256 * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
  double mx = 0, mn = 1e30; int units = grid / CG;
  for (int i = 0; i < units; ++i) { double c = (double)clk[i] / (iters * 4.0); if (c > mx) mx = c; if (c < mn) mn = c; }
  const double macs = (double)units * iters * 4.0 * (128.0 * CG) * N * 32.0;
  printf("{\"variant\": \"%s\", \"grid\": %d, \"cta_group\": %d, \"N\": %d, \"iters\": %d, \"mmas_per_commit\": %d, \"clk_per_mma_min\": %.1f, "
         "\"clk_per_mma_max\": %.1f, \"kernel_ms\": %.4f, \"tops\": %.1f, \"bad_accumulators\": %d}\n",
         name, grid, CG, N, iters, per_commit, mn, mx, ms, 2.0 * macs / (ms * 1e-3) / 1e12, bad);
  fflush(stdout);
  cudaFree(d_clk); cudaFree(d_bad);
  return bad != 0;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int iters = 20000;
  int rc = 0;
  rc |= run<1, 0, 256>("i8  cta_group::1 M128 N256, one CTA", 1, iters, 64);
  rc |= run<1, 1, 256>("f8  cta_group::1 M128 N256, one CTA", 1, iters, 64);
  rc |= run<1, 0, 256>("i8  cta_group::1 M128 N256, all SMs", sms, iters, 64);
  rc |= run<1, 1, 256>("f8  cta_group::1 M128 N256, all SMs", sms, iters, 64);
  rc |= run<1, 0, 128>("i8  cta_group::1 M128 N128, all SMs", sms, iters, 64);
  rc |= run<1, 0, 256>("i8  cta_group::1 M128 N256, all SMs, commit+wait every 4 MMAs", sms, iters / 4, 4);
  rc |= run<2, 0, 256>("i8  cta_group::2 M256 N256, one pair", 2, iters, 64);
  rc |= run<2, 0, 256>("i8  cta_group::2 M256 N256, all SMs", sms / 2 * 2, iters, 64);
  rc |= run<2, 1, 256>("f8  cta_group::2 M256 N256, all SMs", sms / 2 * 2, iters, 64);
  return rc;
}
