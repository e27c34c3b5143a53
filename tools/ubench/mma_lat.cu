// mma_lat.cu -- tcgen05.mma (kind::tf32, M=128, K=8) timing on sm_100a: cycles per instruction as a function of N, of the
// A operand source (shared memory descriptor vs tensor memory) and of the number of independent accumulator chains.
// Operands are zeros; only timing matters.  One CTA, one issuing thread, clock64 around issue and around completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_lat mma_lat.cu && ./mma_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

struct Res { long long issue, total; };

template <int N, bool TS>
__device__ void run(uint32_t tmem, uint32_t sa, uint32_t sb, uint64_t* bar, uint32_t& phase, int nmma, int chains, int stride_cols, Res* out) {
  const uint32_t idesc = idesc_tf32(128, N);
  const uint64_t da = desc_sw128(sa), db = desc_sw128(sb);
  // warm
  for (int i = 0; i < 4; ++i) { if (TS) mma_ts(tmem, tmem + 448, db, idesc, 0); else mma_ss(tmem, da, db, idesc, 0); }
  commit(bar); mbar_wait(bar, phase); phase ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long t0 = clock64();
  for (int i = 0; i < nmma; ++i) {
    const uint32_t d = tmem + (uint32_t)((i % chains) * stride_cols);
    if (TS) mma_ts(d, tmem + 448 + (i & 3) * 8, db + 2 * (i & 3), idesc, 1); else mma_ss(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1);
  }
  commit(bar);
  const long long t1 = clock64();
  mbar_wait(bar, phase); phase ^= 1;
  const long long t2 = clock64();
  out->issue = t1 - t0; out->total = t2 - t0;
}

__global__ void __launch_bounds__(128) k(Res* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 0.f;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    const uint32_t sa = smem_u32(base), sb = smem_u32(base + 16384);
    int r = 0;
#define RUN(NN, TSV, NM, CH, ST) run<NN, TSV>(tmem, sa, sb, &bar, phase, NM, CH, ST, out + r); ++r;
    // N sweep, one dependent chain of 24
    RUN(32, false, 24, 1, 0) RUN(64, false, 24, 1, 0) RUN(96, false, 24, 1, 0) RUN(128, false, 24, 1, 0) RUN(192, false, 24, 1, 0) RUN(256, false, 24, 1, 0)
    RUN(32, true, 24, 1, 0) RUN(64, true, 24, 1, 0) RUN(96, true, 24, 1, 0) RUN(128, true, 24, 1, 0) RUN(192, true, 24, 1, 0) RUN(256, true, 24, 1, 0)
    // independent chains (distinct accumulator columns), N = 32 and 64
    RUN(32, true, 24, 2, 32) RUN(32, true, 24, 3, 32) RUN(32, true, 24, 4, 32) RUN(32, true, 24, 8, 32)
    RUN(64, true, 24, 2, 64) RUN(64, true, 24, 3, 64) RUN(64, true, 24, 4, 64)
    RUN(32, false, 24, 2, 32) RUN(32, false, 24, 4, 32) RUN(96, false, 24, 2, 96) RUN(96, false, 24, 4, 96)
    // chain length sweep (latency of the first MMA + commit + wake): 1, 2, 4, 8, 12 MMAs, N = 32 TS
    RUN(32, true, 1, 1, 0) RUN(32, true, 2, 1, 0) RUN(32, true, 4, 1, 0) RUN(32, true, 8, 1, 0) RUN(32, true, 12, 1, 0) RUN(32, true, 48, 1, 0)
    RUN(64, true, 12, 1, 0) RUN(128, true, 12, 1, 0) RUN(96, false, 12, 1, 0) RUN(192, false, 12, 1, 0)
#undef RUN
    out[63].issue = r;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  Res* d; cudaMalloc(&d, 64 * sizeof(Res)); cudaMemset(d, 0, 64 * sizeof(Res));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<<<1, 128, 64 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  Res h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {
    "SS N=32 x24 chain1", "SS N=64 x24", "SS N=96 x24", "SS N=128 x24", "SS N=192 x24", "SS N=256 x24",
    "TS N=32 x24 chain1", "TS N=64 x24", "TS N=96 x24", "TS N=128 x24", "TS N=192 x24", "TS N=256 x24",
    "TS N=32 x24 2 chains", "TS N=32 x24 3 chains", "TS N=32 x24 4 chains", "TS N=32 x24 8 chains",
    "TS N=64 x24 2 chains", "TS N=64 x24 3 chains", "TS N=64 x24 4 chains",
    "SS N=32 x24 2 chains", "SS N=32 x24 4 chains", "SS N=96 x24 2 chains", "SS N=96 x24 4 chains",
    "TS N=32 x1", "TS N=32 x2", "TS N=32 x4", "TS N=32 x8", "TS N=32 x12", "TS N=32 x48",
    "TS N=64 x12", "TS N=128 x12", "SS N=96 x12", "SS N=192 x12"};
  int n = (int)h[63].issue;
  for (int i = 0; i < n; ++i) printf("%-24s issue %6lld  total %6lld cycles\n", names[i], h[i].issue, h[i].total);
  return 0;
}
