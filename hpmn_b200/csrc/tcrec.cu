// tcrec.cu -- the hierarchical periodic GRU recurrence on the 5th-gen tensor cores, for the regime where a batch tile of
// 128 samples per SM exists (large batch per GPU) and for hidden sizes above one warp (H = 64).
//
// Reference: the two _Linear calls of the GRU cell, /root/reference/code/util.py:88-107 (gates = sigmoid([x,h] Wg + bg),
// c = tanh([x, r*h] Wc + bc), h' = u*h + (1-u)*c), driven step by step by /root/reference/code/rnn.py:780-793 and stacked
// with every-p-th subsampling by /root/reference/code/hpmn.py:113-131.
//
// One CTA owns 128 samples (one MMA M tile) of one layer and walks its S_k steps:
//
//   TMA warp     x_t tiles [128 x 32] fp32 (hi and lo halves of the 3xTF32 split, written by the producer of x)
//                global -> shared, cp.async.bulk.tensor (SASS UTMALDG), 128-byte swizzle, ring of K-block slots;
//                the weights [3H x (Din+H)] (hi, lo; transposed, K-major, scaled by log2e so that sigma / tanh need
//                no multiply) are loaded once per CTA the same way and stay resident in shared memory
//   MMA warp     one thread issues tcgen05.mma kind::tf32, M = 128:
//                  phase 1   acc[0,3H)  = x_t Wx                (A, B from shared memory)
//                            acc[0,2H) += h_{t-1} Wh_{r,u}      (A = h from TENSOR MEMORY, B from shared memory)
//                  phase 2   acc[2H,3H)+= (r*h_{t-1}) Wh_c      (A = r*h from tensor memory)
//                each product as hi*hi + lo*hi + hi*lo (fp32-level accuracy: 1e-4 parity through 1024 dependent steps)
//   8 epilogue   tcgen05.ld the accumulators (lane = sample, 2 warps per lane quadrant share the columns), add the bias,
//   warps        sigma / tanh / blend in registers, split r*h and h' into hi/lo and tcgen05.st them back into the
//                tensor-memory A operand -- the recurrent state never touches shared memory -- and store the h|r|u|c row
//                for the backward pass (plus, on firing steps, the hi/lo input rows of the layer above)
//
// Two dependent MMA -> epilogue round trips per step (TF applies the reset gate BEFORE the candidate matmul), so one tile
// is latency-bound (~2-3k cycles per step); throughput comes from 148 tiles in flight.  At B = 256 this is ~3x slower than
// the warp-per-sample FFMA wavefront kernels (wave.cu), which is why the library selects by batch size (api.cu).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace hpmn {

// ---- tcgen05 / TMA PTX wrappers -----------------------------------------------------------------------------------
namespace tcr {

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = row m, 32-bit column = k)
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                 "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
                 "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                 "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 3xTF32 split (see tc_gemm.cu): hi = x rounded to tf32 (ties away), lo = tf32-truncated remainder
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x, float hi) { return __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u); }

// shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes (32 tf32 along K), 8-row atoms of 1024 bytes
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// TMA tile loads (tensor maps live in kernel parameter space)
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// shared -> global tile store (completion tracked by the issuing thread's bulk groups); rows beyond the tensor are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

constexpr float LOG2E = 1.4426950408889634f;

}  // namespace tcr

constexpr int TCR_EPI_WARPS = 8;
constexpr int TCR_THREADS = 32 * (TCR_EPI_WARPS + 2);     // + MMA warp + TMA warp
constexpr int TCR_XSLOT = 2 * 128 * 128;                  // one K-block of x: hi tile + lo tile, 128 rows x 128 bytes each

// ---------------------------------------------------------------------------------------------------------------------
// packing: TF-layout GRU weights -> transposed K-major [6H rows: hi(3H) | lo(3H)][KW = DP + H] with the activation scale
// folded in (rows r,u: -log2e so that sigma(a) = 1/(1+2^a'); rows c: 2*log2e so that tanh(a) = 1 - 2/(1+2^a')), and the
// scaled bias [3H].  Backward: true (unscaled) recurrent weights transposed, [6H rows: hi | lo][H] with K = gate column.
// ---------------------------------------------------------------------------------------------------------------------
struct TcrPackArgs {
  int L, H;
  int Din[HPMN_MAX_LAYERS], DP[HPMN_MAX_LAYERS];
  int64_t Wg[HPMN_MAX_LAYERS], bg[HPMN_MAX_LAYERS], Wc[HPMN_MAX_LAYERS], bc[HPMN_MAX_LAYERS];
  int64_t wf[HPMN_MAX_LAYERS], bf[HPMN_MAX_LAYERS], wb[HPMN_MAX_LAYERS], wxt[HPMN_MAX_LAYERS];
};

__global__ void __launch_bounds__(256)
tcr_pack_kernel(const __grid_constant__ TcrPackArgs a, const float* __restrict__ params, float* __restrict__ pw) {
  const int k = blockIdx.y;
  const int H = a.H, Din = a.Din[k], DP = a.DP[k], KW = DP + H, N3 = 3 * H;
  const float* Wg = params + a.Wg[k];
  const float* bg = params + a.bg[k];
  const float* Wc = params + a.Wc[k];
  const float* bc = params + a.bc[k];
  auto weight = [&](int i_in, int n) -> float {     // i_in: row of the TF kernel (x rows then h rows); n: r|u|c column
    const int g = n / H, j = n % H;
    return g < 2 ? Wg[(int64_t)i_in * 2 * H + g * H + j] : Wc[(int64_t)i_in * H + j];
  };
  const int nf = N3 * KW;
  const int total = nf + N3 + nf + N3 * Din;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    if (e < nf) {                         // forward: row n, col kk
      const int n = e / KW, kk = e % KW;
      float w = 0.f;
      if (kk < DP) { if (kk < Din) w = weight(kk, n); } else w = weight(Din + (kk - DP), n);
      w *= (n < 2 * H) ? -tcr::LOG2E : 2.f * tcr::LOG2E;
      const float hi = tcr::tf32_hi(w);
      pw[a.wf[k] + e] = hi;
      pw[a.wf[k] + nf + e] = tcr::tf32_lo(w, hi);
    } else if (e < nf + N3) {
      const int n = e - nf, g = n / H, j = n % H;
      pw[a.bf[k] + n] = (g < 2 ? bg[g * H + j] : bc[j]) * ((n < 2 * H) ? -tcr::LOG2E : 2.f * tcr::LOG2E);
    } else {                              // backward: rows [0,H) Wc_h^T, [H,2H) Wu_h^T, [2H,3H) Wr_h^T; K = gate column j
      const int r = e - nf - N3;
      if (r >= nf) {                      // WxT [3H][Din]: dX = dA * WxT (plain fp32, split by the GEMM itself)
        const int q = r - nf, n = q / Din, i = q % Din;
        pw[a.wxt[k] + q] = weight(i, n);
      } else if (r < N3 * H) {
        const int n = r / H, j = r % H, blk = n / H, i = n % H;      // output i = hidden input index
        const int col = blk == 0 ? 2 * H + j : (blk == 1 ? H + j : j);
        const float w = weight(Din + i, col);
        const float hi = tcr::tf32_hi(w);
        pw[a.wb[k] + r] = hi;
        pw[a.wb[k] + N3 * H + r] = tcr::tf32_lo(w, hi);
      }
    }
  }
}

// x [rows, D] -> xh, xl [rows, DP] (hi / lo halves of the 3xTF32 split, columns >= D zero)
__global__ void __launch_bounds__(256)
tcr_split_kernel(const float* __restrict__ x, float* __restrict__ xh, float* __restrict__ xl, int64_t rows, int D, int DP) {
  const int64_t n = rows * (DP / 4);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const int64_t r = e / (DP / 4);
    const int c = (int)(e % (DP / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) v = ldg_nc_f4(reinterpret_cast<const float4*>(x + r * D + c));
    float4 hi, lo;
    hi.x = tcr::tf32_hi(v.x); hi.y = tcr::tf32_hi(v.y); hi.z = tcr::tf32_hi(v.z); hi.w = tcr::tf32_hi(v.w);
    lo.x = tcr::tf32_lo(v.x, hi.x); lo.y = tcr::tf32_lo(v.y, hi.y); lo.z = tcr::tf32_lo(v.z, hi.z); lo.w = tcr::tf32_lo(v.w, hi.w);
    *reinterpret_cast<float4*>(xh + r * DP + c) = hi;
    *reinterpret_cast<float4*>(xl + r * DP + c) = lo;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
struct TcrFwdArgs {
  const float* bias;          // [3H] scaled
  float* st;                  // [B][S][4H]   h | r | u | c
  float* nxt_h; float* nxt_l; // [B][S/period][H] hi / lo input rows of the layer above (nullptr for the top layer)
  float* memory;              // [B][L][H]
  int B, S, period, layer, L, nxb;
  long long* dbg;             // HPMN_TCR_DEBUG=1: per-step clock64 stamps of CTA 0 (epilogue warp 0: [t][0..7], MMA thread: [t][8..15])
  int a_first;                // HPMN_TCR_AFIRST=1: tensor-memory layout [A hi|lo][acc0][acc1] instead of [acc0][acc1][A hi|lo]
};

template <int H>
__device__ __forceinline__ float tcr_tanh_scaled(float a2) {
  // a2 = 2*log2e*a.  |a| >= 0.15: 1 - 2/(1+e^{2a});  |a| < 0.15: odd Taylor series (see tanh_f in common.cuh)
  const float e = ex2_ftz(a2);
  const float big = fmaf(-2.0f, rcp_ftz(1.0f + e), 1.0f);
  const float x = a2 * (0.5f / tcr::LOG2E);
  const float x2 = x * x;
  const float small = x * fmaf(x2, fmaf(x2, fmaf(x2, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  return fabsf(x) < 0.15f ? small : big;
}

// STG: the h|r|u|c rows leave through shared-memory staging tiles and TMA tile stores.  With lane = sample every direct store
// instruction touches 32 different rows (32 L2 requests of 16 bytes): at 16 instructions per thread and step the LSU, not the
// recurrence, set the step time (profiles/r2_tcrec_timeline.md).  Needs 4 x 16 KB of shared memory (H = 32).
template <int H, int DP, bool STG>
__global__ void __launch_bounds__(TCR_THREADS, 1)
tcrec_fwd_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                 const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_st, const TcrFwdArgs a) {
  constexpr int N3 = 3 * H, KW = DP + H, NKB = KW / 32, NKX = DP / 32, NKH = H / 32;
  constexpr int WKB = N3 * 128;                      // bytes of one weight K-block tile (3H rows x 128 B)
  constexpr int HC = H / 2;                          // hidden columns per epilogue thread
  // tensor memory: two accumulator buffers [3H] (step parity) so that the x half of step t+1 is computed while the
  // epilogue of step t still reads its accumulators, then the A operand: hi [AOFF, AOFF+H), lo [AOFF+H, AOFF+2H)
  const int AOFF = a.a_first ? 0 : 2 * N3;
  const int ACC0 = a.a_first ? 2 * H : 0;
  constexpr int TCOLS = (2 * N3 + 2 * H) <= 256 ? 256 : 512;
  static_assert(2 * N3 + 2 * H <= 512, "tensor memory budget");
  constexpr uint32_t ID_X = tcr::idesc_tf32(128, N3), ID_G = tcr::idesc_tf32(128, 2 * H), ID_C = tcr::idesc_tf32(128, H);
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still a shared pointer
  float* sBias = reinterpret_cast<float*>(base);     // [3H] (1 KB reserved)
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 768);
  unsigned char* sWh = base + 1024;                  // [NKB][3H x 128 B] hi
  unsigned char* sWl = sWh + NKB * WKB;              // lo
  unsigned char* sX = sWl + NKB * WKB;               // ring of a.nxb slots: [hi 16 KB | lo 16 KB]
  unsigned char* sStg = sX + (size_t)a.nxb * TCR_XSLOT;   // STG: 4 staging tiles [128 rows x 128 B] (h, r, u, c), 128-byte swizzle
  static_assert(!STG || H == 32, "staged stores: one 32-column block per gate");
  uint64_t* bar_w = bars;            // weights landed
  uint64_t* bar_r = bars + 1;        // r- and u-gate accumulators complete (h_{t-1} no longer needed as an operand)
  uint64_t* bar_2 = bars + 3;        // candidate accumulators complete
  uint64_t* bar_rh = bars + 4;       // r*h is in tensor memory (8 epilogue warps)
  uint64_t* bar_h = bars + 5;        // h_t is in tensor memory
  uint64_t* full = bars + 6;         // [8]
  uint64_t* empty = full + 8;        // [8]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(empty + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int S = a.S, nxb = a.nxb;

  if (warp == TCR_EPI_WARPS) tcr::tmem_alloc(tslot, TCOLS);
  if (tid == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_r, 1); mbar_init(bar_2, 1);
    mbar_init(bar_rh, TCR_EPI_WARPS); mbar_init(bar_h, TCR_EPI_WARPS);
    for (int i = 0; i < nxb; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_mbar_init();
  }
  for (int e = tid; e < N3; e += TCR_THREADS) sBias[e] = a.bias[e];
  tcr::fence_before();
  __syncthreads();
  tcr::fence_after();
  const uint32_t tmem = *tslot;

  if (warp < TCR_EPI_WARPS) {
    // =========================== epilogue warps ===========================
    const int q = warp & 3, ch = warp >> 2;
    const int64_t b = (int64_t)tile * 128 + q * 32 + lane;
    const bool live = b < a.B;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const int c0 = ch * HC;
    float hp[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i) hp[i] = 0.f;
    {   // h_{-1} = 0 (rnn.py:588 zero_state): clear this thread's slice of the A operand
      float z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0.f;
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) { tcr::st16(tl + AOFF + c0 + 16 * j, z); tcr::st16(tl + AOFF + H + c0 + 16 * j, z); }
      tcr::wait_st();
      tcr::fence_before();
      __syncwarp();
      if (lane == 0) tcr::arrive(bar_h);            // "h_{-1} is in tensor memory": phase 0 of bar_h
    }
    float* strow = a.st + (live ? b : 0) * (int64_t)S * 4 * H;
    const int srow = q * 32 + lane;                  // row of the staging tiles
    // 16 consecutive columns of this thread's row -> staging tile of gate g (swizzled 16-byte chunks, conflict free)
    auto stage16 = [&](int g, int col, const float (&v)[16]) {
      unsigned char* r = sStg + g * (128 * 128) + srow * 128;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd)
        *reinterpret_cast<float4*>(r + ((((col >> 2) + qd) ^ (srow & 7)) << 4)) = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
    };
    // every epilogue thread has written its part of gate tiles [g0, g1): hand them to the TMA engine.  The issuing thread first
    // waits until the engine has finished READING its earlier tiles, so that whatever is written after this barrier is free.
    auto flush = [&](int bar_id, int g0, int g1, int t) {
      fence_proxy_async();
      if (tid == 0) bulk_wait_read<0>();
      tcr::named_bar(bar_id, 32 * TCR_EPI_WARPS);
      if (tid == 0) {
        for (int g = g0; g < g1; ++g) tcr::tma_store_3d(&tm_st, sStg + g * (128 * 128), g * H, t, tile * 128);
        bulk_commit();
      }
    };
    const int Sn = a.period > 0 ? S / a.period : 0;
    int fire = a.period;                             // steps until the next firing step
    int64_t nrow = (live ? b : 0) * (int64_t)Sn * H;
    for (int t = 0; t < S; ++t) {
      const uint32_t acc = tl + ACC0 + (uint32_t)(t & 1) * N3;
      float u[HC];
      long long* dbg = (a.dbg != nullptr && tile == 0 && tid == 0 && t < 2048) ? a.dbg + 16 * t : nullptr;
      if (dbg) dbg[0] = clock64();
      mbar_wait(bar_r, t & 1);
      tcr::fence_after();
      if (dbg) dbg[1] = clock64();
      // ---- r gate -> r*h into the A operand
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float v[16], hi[16], lo[16];
        tcr::ld16(acc + 0 * H + c0 + 16 * j, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float r = rcp_ftz(1.0f + ex2_ftz(v[i] + sBias[c0 + 16 * j + i]));
          v[i] = r;
          const float rh = r * hp[16 * j + i];
          hi[i] = tcr::tf32_hi(rh);
          lo[i] = tcr::tf32_lo(rh, hi[i]);
        }
        tcr::st16(tl + AOFF + c0 + 16 * j, hi);
        tcr::st16(tl + AOFF + H + c0 + 16 * j, lo);
        if (STG) {
          stage16(1, c0 + 16 * j, v);
        } else if (live) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(strow + H + c0 + 16 * j + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
      tcr::wait_st();
      tcr::fence_before();
      __syncwarp();
      if (lane == 0) tcr::arrive(bar_rh);
      if (dbg) dbg[2] = clock64();
      if (STG) flush(1, 1, 2, t);
      // ---- u gate (kept in registers) while the tensor core computes the candidate's recurrent half
      if (dbg) dbg[3] = clock64();
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float v[16];
        tcr::ld16(acc + 1 * H + c0 + 16 * j, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) u[16 * j + i] = rcp_ftz(1.0f + ex2_ftz(v[i] + sBias[H + c0 + 16 * j + i]));
        if (STG) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = u[16 * j + i];
          stage16(2, c0 + 16 * j, v);
        } else if (live) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(strow + 2 * H + c0 + 16 * j + 4 * i) =
                make_float4(u[16 * j + 4 * i], u[16 * j + 4 * i + 1], u[16 * j + 4 * i + 2], u[16 * j + 4 * i + 3]);
        }
      }
      if (STG) flush(2, 2, 3, t);
      // ---- candidate, blend, h' into the A operand
      if (dbg) dbg[4] = clock64();
      mbar_wait(bar_2, t & 1);
      tcr::fence_after();
      if (dbg) dbg[5] = clock64();
      --fire;
      const bool firing = a.nxt_h != nullptr && fire == 0;
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float v[16], hi[16], lo[16], hn[16];
        tcr::ld16(acc + 2 * H + c0 + 16 * j, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float c = tcr_tanh_scaled<H>(v[i] + sBias[2 * H + c0 + 16 * j + i]);
          v[i] = c;
          const float h = fmaf(u[16 * j + i], hp[16 * j + i] - c, c);     // u*h + (1-u)*c
          hn[i] = h;
          hp[16 * j + i] = h;
          hi[i] = tcr::tf32_hi(h);
          lo[i] = tcr::tf32_lo(h, hi[i]);
        }
        tcr::st16(tl + AOFF + c0 + 16 * j, hi);
        tcr::st16(tl + AOFF + H + c0 + 16 * j, lo);
        if (j + 1 == HC / 16) {      // the recurrence continues as soon as h' is in tensor memory; the stores below trail it
          tcr::wait_st();
          tcr::fence_before();
          __syncwarp();
          if (lane == 0) tcr::arrive(bar_h);
          if (dbg) dbg[6] = clock64();
        }
        if (STG) { stage16(0, c0 + 16 * j, hn); stage16(3, c0 + 16 * j, v); }
        if (live) {
          if (!STG) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              *reinterpret_cast<float4*>(strow + c0 + 16 * j + 4 * i) = make_float4(hn[4 * i], hn[4 * i + 1], hn[4 * i + 2], hn[4 * i + 3]);
              *reinterpret_cast<float4*>(strow + 3 * H + c0 + 16 * j + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
          if (firing) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              *reinterpret_cast<float4*>(a.nxt_h + nrow + c0 + 16 * j + 4 * i) = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              *reinterpret_cast<float4*>(a.nxt_l + nrow + c0 + 16 * j + 4 * i) = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
        }
      }
      if (STG) {                                     // h tile (gate 0) and c tile (gate 3): two stores, one bulk group
        fence_proxy_async();
        if (tid == 0) bulk_wait_read<0>();
        tcr::named_bar(3, 32 * TCR_EPI_WARPS);
        if (tid == 0) {
          tcr::tma_store_3d(&tm_st, sStg, 0, t, tile * 128);
          tcr::tma_store_3d(&tm_st, sStg + 3 * (128 * 128), 3 * H, t, tile * 128);
          bulk_commit();
        }
      }
      if (fire == 0) { fire = a.period; nrow += H; }
      strow += 4 * H;
      if (dbg) dbg[7] = clock64();
    }
    if (STG && tid == 0) bulk_wait_read<0>();         // the staging tiles must outlive the last stores' reads
    if (live) {   // final state -> memory slot of this layer (hpmn.py:120,130)
      float* m = a.memory + (b * a.L + a.layer) * (int64_t)H + c0;
#pragma unroll
      for (int i = 0; i < HC / 4; ++i) *reinterpret_cast<float4*>(m + 4 * i) = make_float4(hp[4 * i], hp[4 * i + 1], hp[4 * i + 2], hp[4 * i + 3]);
    }
  } else if (warp == TCR_EPI_WARPS) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      // descriptors differ only in the start-address field: precompute the bases, add byte offsets >> 4
      const uint64_t dwh = tcr::desc_sw128(smem_u32(sWh)), dwl = tcr::desc_sw128(smem_u32(sWl));
      const uint64_t dx0 = tcr::desc_sw128(smem_u32(sX));
      const uint32_t a_hi = tmem + AOFF, a_lo = tmem + AOFF + H;
      int xi = 0;                                    // running x K-block index (ring position)
      // x half of one step into accumulator buffer `accb`: acc[0,3H) = x_t * Wx  (no dependence on the recurrence)
      auto issue_x = [&](uint32_t accb) {
        for (int kb = 0; kb < NKX; ++kb, ++xi) {
          const int slot = xi % nxb;
          mbar_wait(&full[slot], (uint32_t)(xi / nxb) & 1u);
          tcr::fence_after();
          const uint64_t dxh = dx0 + (uint64_t)((slot * TCR_XSLOT) >> 4), dxl = dxh + ((128 * 128) >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bo = (uint64_t)((kb * WKB + ks * 32) >> 4);
            tcr::mma_ss(accb, dxh + 2 * ks, dwh + bo, ID_X, (kb | ks) != 0);
            tcr::mma_ss(accb, dxl + 2 * ks, dwh + bo, ID_X, 1);
            tcr::mma_ss(accb, dxh + 2 * ks, dwl + bo, ID_X, 1);
          }
          tcr::commit(&empty[slot]);                 // slot reusable once these MMAs have read it
        }
      };
      // recurrent half: gates r,u together (acc[0,2H) += h * Wh_{r,u}) or the candidate (acc[2H,3H) += (r*h) * Wh_c).
      // An MMA instruction costs ~50 cycles to issue whatever its N below ~100 (profiles/r2_mma_issue_latency.md), so the
      // two gates that share the A operand go in one instruction.
      auto issue_h = [&](uint32_t accb, bool cand) {
#pragma unroll
        for (int kb = 0; kb < NKH; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bo = (uint64_t)(((NKX + kb) * WKB + (cand ? 2 * H * 128 : 0) + ks * 32) >> 4);   // rows [0,2H) or [2H,3H)
            const uint32_t ko = kb * 32 + ks * 8;
            const uint32_t d = accb + (cand ? 2 * H : 0);
            const uint32_t id = cand ? ID_C : ID_G;
            tcr::mma_ts(d, a_hi + ko, dwh + bo, id, 1);
            tcr::mma_ts(d, a_lo + ko, dwh + bo, id, 1);
            tcr::mma_ts(d, a_hi + ko, dwl + bo, id, 1);
          }
        }
      };
      issue_x(tmem + ACC0);
      for (int t = 0; t < S; ++t) {
        const uint32_t accb = tmem + ACC0 + (uint32_t)(t & 1) * N3;
        long long* dbg = (a.dbg != nullptr && tile == 0 && t < 2048) ? a.dbg + 16 * t + 8 : nullptr;
        if (dbg) dbg[0] = clock64();
        mbar_wait(bar_h, t & 1);                     // h_{t-1} is in tensor memory
        tcr::fence_after();
        if (dbg) dbg[1] = clock64();
        issue_h(accb, false); tcr::commit(bar_r);
        if (dbg) { dbg[2] = clock64(); dbg[3] = dbg[2]; }
        // the other buffer was drained before bar_h(t-1) completed: give the tensor core the next step's x half now, it
        // runs while the epilogue works on r
        if (t + 1 < S) issue_x(tmem + ACC0 + (uint32_t)((t + 1) & 1) * N3);
        if (dbg) dbg[4] = clock64();
        mbar_wait(bar_rh, t & 1);
        tcr::fence_after();
        if (dbg) dbg[5] = clock64();
        issue_h(accb, true); tcr::commit(bar_2);
        if (dbg) dbg[6] = clock64();
      }
    }
    __syncwarp();
  } else {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      tcr::prefetch_tmap(&tm_w); tcr::prefetch_tmap(&tm_xh); tcr::prefetch_tmap(&tm_xl);
      mbar_expect_tx(bar_w, 2u * NKB * WKB);
      for (int kb = 0; kb < NKB; ++kb) {
        tcr::tma_2d(sWh + kb * WKB, &tm_w, kb * 32, 0, bar_w);
        tcr::tma_2d(sWl + kb * WKB, &tm_w, kb * 32, N3, bar_w);
      }
      int xi = 0;
      for (int t = 0; t < S; ++t) {
        for (int kb = 0; kb < NKX; ++kb, ++xi) {
          const int slot = xi % nxb;
          if (xi >= nxb) mbar_wait(&empty[slot], (uint32_t)(xi / nxb - 1) & 1u);
          unsigned char* dst = sX + (size_t)slot * TCR_XSLOT;
          mbar_expect_tx(&full[slot], (uint32_t)TCR_XSLOT);
          tcr::tma_3d(dst, &tm_xh, kb * 32, t, tile * 128, &full[slot]);
          tcr::tma_3d(dst + 128 * 128, &tm_xl, kb * 32, t, tile * 128, &full[slot]);
        }
      }
    }
    __syncwarp();
  }
  tcr::fence_before();
  __syncthreads();
  if (warp == TCR_EPI_WARPS) { tcr::fence_after(); tcr::tmem_dealloc(tmem, TCOLS); }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: the reverse-time adjoint of one layer for 128 samples per CTA.  Only the RECURRENT
// contractions live here -- d(r*h) = da_c Wc_h^T and dh_g = [da_r | da_u] Wg_h^T, K = N = H -- because they are the
// dependent chain; dX = dA Wx^T and the weight gradients are plain GEMMs over the stored dA (tc_gemm.cu) and run afterwards.
//   TMA warp     per step the saved u, c, h_{t-1}, r blocks [128 x 32] of the state rows (128-byte swizzle, ring of 8 slots)
//   MMA warp     d(r*h) = da_c Wc_h^T ; dh_g = da_u Wu_h^T  (as soon as da_c, da_u are in tensor memory), then
//                dh_g += da_r Wr_h^T  (once the epilogue has turned d(r*h) into da_r)
//   epilogue     dh -> da_c, da_u (tcgen05.st as the next A operands) ; d(r*h) -> da_r ; dh_{t-1} = dh*u + d(r*h)*r + dh_g
// Tensor memory: d(r*h) [H] | dh_g [H] | da_c hi,lo | da_u hi,lo | da_r hi,lo = 8H columns.
// ---------------------------------------------------------------------------------------------------------------------
struct TcrBwdArgs {
  const float* dmemory;       // [B][L][H]
  const float* dx_up;         // [B][S/period][H] gradient wrt the input of the layer above (its dX), nullptr for the top layer
  float* da;                  // [B][S][3H]  da_r | da_u | da_c
  float* hr;                  // [B][S][2H]  h_prev | r*h_prev (weight-gradient operands), or nullptr
  int B, S, period, layer, L;
};

constexpr int TCR_BSLOTS = 8;

template <int H, bool STG>
__global__ void __launch_bounds__(TCR_THREADS, 1)
tcrec_bwd_kernel(const __grid_constant__ CUtensorMap tm_st, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_da, const TcrBwdArgs a) {
  constexpr int NKH = H / 32, WKB = 3 * H * 128, HC = H / 2, BLK = 128 * 128;
  constexpr int C_DRH = 0, C_DHG = H, C_AC = 2 * H, C_AU = 4 * H, C_AR = 6 * H;     // A operands: hi [C, C+H), lo [C+H, C+2H)
  constexpr int TCOLS = 8 * H;
  constexpr uint32_t ID_H = tcr::idesc_tf32(128, H);
  constexpr int PER_STEP = 4 * NKH;                  // blocks per step: u, c, h_prev, r (NKH each)
  static_assert(TCR_BSLOTS % PER_STEP == 0, "ring must hold whole steps");
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base);
  unsigned char* sWh = base + 1024;                  // [NKH][3H x 128 B] hi: rows [0,H) Wc_h^T, [H,2H) Wu_h^T, [2H,3H) Wr_h^T
  unsigned char* sWl = sWh + NKH * WKB;
  unsigned char* sR = sWl + NKH * WKB;               // ring of TCR_BSLOTS blocks of 16 KB
  unsigned char* sStg = sR + (size_t)TCR_BSLOTS * BLK;   // STG: staging tiles of da_r, da_u, da_c (see tcrec_fwd_kernel)
  static_assert(!STG || H == 32, "staged stores: one 32-column block per gate");
  uint64_t* bar_w = bars;
  uint64_t* bar_m1 = bars + 1;       // d(r*h) complete
  uint64_t* bar_m2 = bars + 2;       // dh_g complete
  uint64_t* bar_a1 = bars + 3;       // da_c, da_u in tensor memory
  uint64_t* bar_a2 = bars + 4;       // da_r in tensor memory
  uint64_t* full = bars + 5;         // [8]
  uint64_t* empty = full + 8;        // [8]
  uint32_t* tslot = reinterpret_cast<uint32_t*>(empty + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int S = a.S;

  if (warp == TCR_EPI_WARPS) tcr::tmem_alloc(tslot, TCOLS);
  if (tid == 0) {
    mbar_init(bar_w, 1); mbar_init(bar_m1, 1); mbar_init(bar_m2, 1);
    mbar_init(bar_a1, TCR_EPI_WARPS); mbar_init(bar_a2, TCR_EPI_WARPS);
    for (int i = 0; i < TCR_BSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], TCR_EPI_WARPS / NKH); }
    fence_mbar_init();
  }
  tcr::fence_before();
  __syncthreads();
  tcr::fence_after();
  const uint32_t tmem = *tslot;

  if (warp < TCR_EPI_WARPS) {
    // =========================== epilogue warps ===========================
    const int q = warp & 3, ch = warp >> 2;
    const int row = q * 32 + lane;
    const int64_t b = (int64_t)tile * 128 + row;
    const bool live = b < a.B;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    const int c0 = ch * HC;
    const int bi = c0 / 32;                          // which 32-column block of a gate this thread's columns live in
    const int co = (c0 % 32) / 4;                    // first 16-byte chunk inside that block
    const int64_t bb = live ? b : 0;
    float dh[HC], hp[HC];
    {
      const float* dm = a.dmemory + (bb * a.L + a.layer) * (int64_t)H + c0;
#pragma unroll
      for (int i = 0; i < HC / 4; ++i) {
        const float4 v = live ? __ldg(reinterpret_cast<const float4*>(dm) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        dh[4 * i] = v.x; dh[4 * i + 1] = v.y; dh[4 * i + 2] = v.z; dh[4 * i + 3] = v.w;
      }
    }
    const int Sn = a.period > 0 ? S / a.period : 0;
    float* darow = a.da + (bb * S + (S - 1)) * (int64_t)(3 * H);
    float* hrrow = a.hr ? a.hr + (bb * S + (S - 1)) * (int64_t)(2 * H) : nullptr;
    // swizzled read of 16 consecutive columns (4 chunks of 16 bytes) of this thread's row from a ring block
    auto rd16 = [&](const unsigned char* blk, int j, float (&v)[16]) {
      const unsigned char* r = blk + row * 128;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const int chunk = (co + 4 * j + qd) & 7;
        const float4 f = *reinterpret_cast<const float4*>(r + ((chunk ^ (row & 7)) << 4));
        v[4 * qd] = f.x; v[4 * qd + 1] = f.y; v[4 * qd + 2] = f.z; v[4 * qd + 3] = f.w;
      }
    };
    auto stage16 = [&](int g, int col, const float (&v)[16]) {
      unsigned char* r = sStg + g * BLK + row * 128;
#pragma unroll
      for (int qd = 0; qd < 4; ++qd)
        *reinterpret_cast<float4*>(r + ((((col >> 2) + qd) ^ (row & 7)) << 4)) = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
    };
    auto flush = [&](int bar_id, int g0, int g1, int t) {
      fence_proxy_async();
      if (tid == 0) bulk_wait_read<0>();
      tcr::named_bar(bar_id, 32 * TCR_EPI_WARPS);
      if (tid == 0) {
        for (int g = g0; g < g1; ++g) tcr::tma_store_3d(&tm_da, sStg + g * BLK, g * H, t, tile * 128);
        bulk_commit();
      }
    };
    auto release = [&](int idx) {                    // this warp has read everything it needs from ring block idx
      __syncwarp();
      if (lane == 0) tcr::arrive(&empty[idx % TCR_BSLOTS]);
    };
    int it = 0;
    for (int t = S - 1; t >= 0; --t, ++it) {
      if (a.dx_up != nullptr && (t + 1) % a.period == 0 && live) {       // the layer above consumed h_t: add its dX (hpmn.py:124-128)
        const float* dxu = a.dx_up + (bb * Sn + ((t + 1) / a.period - 1)) * (int64_t)H + c0;
#pragma unroll
        for (int i = 0; i < HC / 4; ++i) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(dxu) + i);
          dh[4 * i] += v.x; dh[4 * i + 1] += v.y; dh[4 * i + 2] += v.z; dh[4 * i + 3] += v.w;
        }
      }
      const int i_u = it * PER_STEP + 0 * NKH + bi, i_c = it * PER_STEP + 1 * NKH + bi, i_h = it * PER_STEP + 2 * NKH + bi,
                i_r = it * PER_STEP + 3 * NKH + bi;
      const uint32_t par = (uint32_t)((it * PER_STEP) / TCR_BSLOTS) & 1u;
      // ---- phase A: dh -> da_c, da_u (A operands), dh*u kept
      mbar_wait(&full[i_u % TCR_BSLOTS], par);
      mbar_wait(&full[i_c % TCR_BSLOTS], par);
      mbar_wait(&full[i_h % TCR_BSLOTS], par);
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float u[16], c[16], h[16], hi[16], lo[16], dac[16];
        rd16(sR + (size_t)(i_u % TCR_BSLOTS) * BLK, j, u);
        rd16(sR + (size_t)(i_c % TCR_BSLOTS) * BLK, j, c);
        rd16(sR + (size_t)(i_h % TCR_BSLOTS) * BLK, j, h);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float d = dh[16 * j + i];
          hp[16 * j + i] = h[i];
          const float dc = d * (1.0f - u[i]);
          const float du = d * (h[i] - c[i]);
          dh[16 * j + i] = d * u[i];                                  // dh_prev partial
          dac[i] = dc * fmaf(-c[i], c[i], 1.0f);                      // da_c = dc * (1 - c^2)
          u[i] = du * u[i] * (1.0f - u[i]);                           // da_u
          hi[i] = tcr::tf32_hi(dac[i]);
          lo[i] = tcr::tf32_lo(dac[i], hi[i]);
        }
        tcr::st16(tl + C_AC + c0 + 16 * j, hi);
        tcr::st16(tl + C_AC + H + c0 + 16 * j, lo);
#pragma unroll
        for (int i = 0; i < 16; ++i) { hi[i] = tcr::tf32_hi(u[i]); lo[i] = tcr::tf32_lo(u[i], hi[i]); }
        tcr::st16(tl + C_AU + c0 + 16 * j, hi);
        tcr::st16(tl + C_AU + H + c0 + 16 * j, lo);
        if (STG) {
          stage16(2, c0 + 16 * j, dac); stage16(1, c0 + 16 * j, u);
        } else if (live) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<float4*>(darow + 2 * H + c0 + 16 * j + 4 * i) = make_float4(dac[4 * i], dac[4 * i + 1], dac[4 * i + 2], dac[4 * i + 3]);
            *reinterpret_cast<float4*>(darow + H + c0 + 16 * j + 4 * i) = make_float4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
          }
        }
      }
      tcr::wait_st();
      tcr::fence_before();
      __syncwarp();
      if (lane == 0) tcr::arrive(bar_a1);
      release(i_u); release(i_c); release(i_h);
      if (STG) flush(1, 1, 3, t);
      // ---- phase B: d(r*h) -> da_r (A operand), dh_prev += d(r*h) * r
      mbar_wait(&full[i_r % TCR_BSLOTS], par);
      mbar_wait(bar_m1, it & 1);
      tcr::fence_after();
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float v[16], r[16], hi[16], lo[16];
        tcr::ld16(tl + C_DRH + c0 + 16 * j, v);
        rd16(sR + (size_t)(i_r % TCR_BSLOTS) * BLK, j, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float drh = v[i];
          const float dr = drh * hp[16 * j + i];
          dh[16 * j + i] = fmaf(drh, r[i], dh[16 * j + i]);
          v[i] = dr * r[i] * (1.0f - r[i]);                           // da_r
          hi[i] = tcr::tf32_hi(v[i]);
          lo[i] = tcr::tf32_lo(v[i], hi[i]);
        }
        tcr::st16(tl + C_AR + c0 + 16 * j, hi);
        tcr::st16(tl + C_AR + H + c0 + 16 * j, lo);
        if (STG) stage16(0, c0 + 16 * j, v);
        if (live) {
          if (!STG) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *reinterpret_cast<float4*>(darow + c0 + 16 * j + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          if (hrrow) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              *reinterpret_cast<float4*>(hrrow + c0 + 16 * j + 4 * i) =
                  make_float4(hp[16 * j + 4 * i], hp[16 * j + 4 * i + 1], hp[16 * j + 4 * i + 2], hp[16 * j + 4 * i + 3]);
              *reinterpret_cast<float4*>(hrrow + H + c0 + 16 * j + 4 * i) =
                  make_float4(r[4 * i] * hp[16 * j + 4 * i], r[4 * i + 1] * hp[16 * j + 4 * i + 1], r[4 * i + 2] * hp[16 * j + 4 * i + 2],
                              r[4 * i + 3] * hp[16 * j + 4 * i + 3]);
            }
          }
        }
      }
      tcr::wait_st();
      tcr::fence_before();
      __syncwarp();
      if (lane == 0) tcr::arrive(bar_a2);
      release(i_r);
      if (STG) flush(2, 0, 1, t);
      // ---- phase C: dh_{t-1} = dh*u + d(r*h)*r + dh_g
      mbar_wait(bar_m2, it & 1);
      tcr::fence_after();
#pragma unroll
      for (int j = 0; j < HC / 16; ++j) {
        float v[16];
        tcr::ld16(tl + C_DHG + c0 + 16 * j, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) dh[16 * j + i] += v[i];
      }
      tcr::fence_before();
      darow -= 3 * H;
      if (hrrow) hrrow -= 2 * H;
    }
    if (STG && tid == 0) bulk_wait_read<0>();
  } else if (warp == TCR_EPI_WARPS) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      mbar_wait(bar_w, 0);
      const uint64_t dwh = tcr::desc_sw128(smem_u32(sWh)), dwl = tcr::desc_sw128(smem_u32(sWl));
      auto issue = [&](uint32_t d, uint32_t a_col, int rowblk, uint32_t first_acc) {
#pragma unroll
        for (int kb = 0; kb < NKH; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bo = (uint64_t)((kb * WKB + rowblk * H * 128 + ks * 32) >> 4);
            const uint32_t ko = kb * 32 + ks * 8;
            tcr::mma_ts(tmem + d, tmem + a_col + ko, dwh + bo, ID_H, (kb | ks) ? 1u : first_acc);
            tcr::mma_ts(tmem + d, tmem + a_col + H + ko, dwh + bo, ID_H, 1);
            tcr::mma_ts(tmem + d, tmem + a_col + ko, dwl + bo, ID_H, 1);
          }
        }
      };
      for (int it = 0; it < S; ++it) {
        mbar_wait(bar_a1, it & 1);
        tcr::fence_after();
        issue(C_DRH, C_AC, 0, 0); tcr::commit(bar_m1);
        issue(C_DHG, C_AU, 1, 0);
        mbar_wait(bar_a2, it & 1);
        tcr::fence_after();
        issue(C_DHG, C_AR, 2, 1); tcr::commit(bar_m2);
      }
    }
    __syncwarp();
  } else {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      tcr::prefetch_tmap(&tm_w); tcr::prefetch_tmap(&tm_st);
      mbar_expect_tx(bar_w, 2u * NKH * WKB);
      for (int kb = 0; kb < NKH; ++kb) {
        tcr::tma_2d(sWh + kb * WKB, &tm_w, kb * 32, 0, bar_w);
        tcr::tma_2d(sWl + kb * WKB, &tm_w, kb * 32, 3 * H, bar_w);
      }
      int idx = 0;
      for (int t = S - 1; t >= 0; --t) {
        // order of use: u (gate cols 2H..), c (3H..), h of step t-1 (0..; step -1 is out of bounds -> zeros = the zero state), r (H..)
        const int gcol[4] = {2 * H, 3 * H, 0, H};
        const int gt[4] = {t, t, t - 1, t};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          for (int kb = 0; kb < NKH; ++kb, ++idx) {
            const int slot = idx % TCR_BSLOTS;
            if (idx >= TCR_BSLOTS) mbar_wait(&empty[slot], (uint32_t)(idx / TCR_BSLOTS - 1) & 1u);
            mbar_expect_tx(&full[slot], (uint32_t)BLK);
            tcr::tma_3d(sR + (size_t)slot * BLK, &tm_st, gcol[g] + kb * 32, gt[g], tile * 128, &full[slot]);
          }
        }
      }
    }
    __syncwarp();
  }
  tcr::fence_before();
  __syncthreads();
  if (warp == TCR_EPI_WARPS) { tcr::fence_after(); tcr::tmem_dealloc(tmem, TCOLS); }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }();
  return fn;
}

// fp32 tensor [d2][d1][d0] (d0 contiguous), box [b2][b1][32 floats], 128-byte swizzle; rank 2 when d2 == 0
static bool make_tmap(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b1, uint32_t b2) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  const cuuint32_t rank = d2 ? 3 : 2;
  cuuint64_t dims[3] = {d0, d1, d2 ? d2 : 1};
  cuuint64_t strides[2] = {d0 * sizeof(float), d0 * d1 * sizeof(float)};
  cuuint32_t box[3] = {32, b1, b2 ? b2 : 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// HPMN_TCR_DEBUG=1: 2048 steps x 16 clock64 stamps (layer 0, CTA 0), read back with hpmn_debug_tcr_stamps
long long* tcr_debug_buffer() {
  static long long* buf = [] {
    const char* e = getenv("HPMN_TCR_DEBUG");
    long long* p = nullptr;
    if (e && e[0] == '1') { if (cudaMalloc(&p, 2048 * 16 * sizeof(long long)) != cudaSuccess) p = nullptr; else cudaMemset(p, 0, 2048 * 16 * sizeof(long long)); }
    return p;
  }();
  return buf;
}

bool tcrec_supported(const Dims& d) {
  if (d.H != 32 && d.H != 64) return false;
  if (d.D > 64) return false;
  return true;
}

TcrLayout make_tcr_layout(const Dims& d) {
  TcrLayout t; memset(&t, 0, sizeof(t));
  if (!tcrec_supported(d)) return t;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) & ~(size_t)1023; return o; };
  const size_t f = sizeof(float);
  const int H = d.H;
  for (int k = 0; k < d.L; ++k) {
    t.DP[k] = k == 0 ? ((d.D + 31) / 32) * 32 : H;
    const size_t rows = (size_t)d.B * d.S[k];
    const size_t nf = (size_t)3 * H * (t.DP[k] + H);
    t.wf[k] = take(2 * nf * f);
    t.bf[k] = take((size_t)3 * H * f);
    t.wb[k] = take((size_t)2 * 3 * H * H * f);
    t.wxt[k] = take((size_t)3 * H * d.Din[k] * f);
    t.xh[k] = take(rows * t.DP[k] * f);
    t.xl[k] = take(rows * t.DP[k] * f);
    t.st[k] = take(rows * 4 * H * f);
    t.da[k] = take(rows * 3 * H * f);
    t.dx[k] = take(rows * t.DP[k] * f);
    t.hr[k] = take(rows * 2 * H * f);
  }
  t.total = off;
  t.ok = true;
  return t;
}

static size_t tcr_fwd_smem(int H, int DP, int nxb, bool stg) {   // alignment slack, bias + barriers, weights, x ring, staging tiles
  return 1024 + 1024 + (size_t)2 * ((DP + H) / 32) * (3 * H * 128) + (size_t)nxb * TCR_XSLOT + (stg ? 4 * 128 * 128 : 0);
}

void launch_tcr_pack(const Launch& L, const Dims& d, const ParamLayout& pl, const TcrLayout& tl, const float* params, char* ws,
                     cudaStream_t st) {
  TcrPackArgs a; memset(&a, 0, sizeof(a));
  a.L = d.L; a.H = d.H;
  for (int k = 0; k < d.L; ++k) {
    a.Din[k] = d.Din[k]; a.DP[k] = tl.DP[k];
    a.Wg[k] = pl.Wg[k]; a.bg[k] = pl.bg[k]; a.Wc[k] = pl.Wc[k]; a.bc[k] = pl.bc[k];
    a.wf[k] = (int64_t)(tl.wf[k] / sizeof(float)); a.bf[k] = (int64_t)(tl.bf[k] / sizeof(float)); a.wb[k] = (int64_t)(tl.wb[k] / sizeof(float));
    a.wxt[k] = (int64_t)(tl.wxt[k] / sizeof(float));
  }
  tcr_pack_kernel<<<dim3(16, d.L), 256, 0, st>>>(a, params, reinterpret_cast<float*>(ws));
  ++*L.counter;
}

void launch_tcr_split(const Launch& L, const float* x, float* xh, float* xl, int64_t rows, int D, int DP, cudaStream_t st) {
  int64_t blocks = (rows * (DP / 4) + 255) / 256;
  if (blocks > (int64_t)L.sms * 16) blocks = (int64_t)L.sms * 16;
  if (blocks < 1) blocks = 1;
  tcr_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, xh, xl, rows, D, DP);
  ++*L.counter;
}

template <int H, int DP>
static bool tcr_fwd_go(const Launch& L, const Dims& d, const TcrLayout& tl, int k, char* ws, float* memory, cudaStream_t st) {
  static const bool stg_env = [] { const char* e = getenv("HPMN_TCR_STAGE"); return !(e && e[0] == '0'); }();
  constexpr bool CAN_STAGE = H == 32;
  bool stg = CAN_STAGE && stg_env;
  int nxb = 2 * (DP / 32);                         // staged: two steps of x K-blocks beside the four staging tiles
  if (stg && tcr_fwd_smem(H, DP, nxb, true) > 227 * 1024) stg = false;
  if (!stg) {
    nxb = 4 * (DP / 32);
    while (nxb > 1 && tcr_fwd_smem(H, DP, nxb, false) > 227 * 1024) --nxb;
    if (tcr_fwd_smem(H, DP, nxb, false) > 227 * 1024) return false;
    if (nxb > 8) nxb = 8;
  }
  CUtensorMap tm_xh, tm_xl, tm_w, tm_st;
  const float* xh = reinterpret_cast<const float*>(ws + tl.xh[k]);
  const float* xl = reinterpret_cast<const float*>(ws + tl.xl[k]);
  if (!make_tmap(&tm_xh, xh, DP, (uint64_t)d.S[k], (uint64_t)d.B, 1, 128)) return false;
  if (!make_tmap(&tm_xl, xl, DP, (uint64_t)d.S[k], (uint64_t)d.B, 1, 128)) return false;
  if (!make_tmap(&tm_w, reinterpret_cast<const float*>(ws + tl.wf[k]), DP + H, 6 * H, 0, 3 * H, 0)) return false;
  if (!make_tmap(&tm_st, reinterpret_cast<const float*>(ws + tl.st[k]), 4 * H, (uint64_t)d.S[k], (uint64_t)d.B, 1, 128)) return false;
  TcrFwdArgs a;
  a.bias = reinterpret_cast<const float*>(ws + tl.bf[k]);
  a.st = reinterpret_cast<float*>(ws + tl.st[k]);
  const bool top = k == d.L - 1;
  a.nxt_h = top ? nullptr : reinterpret_cast<float*>(ws + tl.xh[k + 1]);
  a.nxt_l = top ? nullptr : reinterpret_cast<float*>(ws + tl.xl[k + 1]);
  a.memory = memory;
  a.B = d.B; a.S = d.S[k]; a.period = top ? 0 : d.P[k]; a.layer = k; a.L = d.L; a.nxb = nxb;
  a.dbg = k == 0 ? tcr_debug_buffer() : nullptr;
  { const char* e = getenv("HPMN_TCR_AFIRST"); a.a_first = (e && e[0] == '1') ? 1 : 0; }
  const size_t smem = tcr_fwd_smem(H, DP, nxb, stg);
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(d.B + 127) / 128, TCR_THREADS, smem, st>>>(tm_xh, tm_xl, tm_w, tm_st, a);
  };
  if (stg) go(tcrec_fwd_kernel<H, DP, CAN_STAGE>); else go(tcrec_fwd_kernel<H, DP, false>);
  ++*L.counter;
  return true;
}

static size_t tcr_bwd_smem(int H, bool stg) {
  return 1024 + 1024 + (size_t)2 * (H / 32) * (3 * H * 128) + (size_t)TCR_BSLOTS * 128 * 128 + (stg ? 3 * 128 * 128 : 0);
}

template <int H>
static bool tcr_bwd_go(const Launch& L, const Dims& d, const TcrLayout& tl, int k, char* ws, const float* dmemory, const float* dx_up,
                       bool write_hr, cudaStream_t st) {
  static const bool stg_env = [] { const char* e = getenv("HPMN_TCR_STAGE"); return !(e && e[0] == '0'); }();
  constexpr bool CAN_STAGE = H == 32;
  const bool stg = CAN_STAGE && stg_env;
  CUtensorMap tm_st, tm_w, tm_da;
  if (!make_tmap(&tm_st, reinterpret_cast<const float*>(ws + tl.st[k]), 4 * H, (uint64_t)d.S[k], (uint64_t)d.B, 1, 128)) return false;
  if (!make_tmap(&tm_w, reinterpret_cast<const float*>(ws + tl.wb[k]), H, 6 * H, 0, 3 * H, 0)) return false;
  if (!make_tmap(&tm_da, reinterpret_cast<const float*>(ws + tl.da[k]), 3 * H, (uint64_t)d.S[k], (uint64_t)d.B, 1, 128)) return false;
  TcrBwdArgs a;
  a.dmemory = dmemory; a.dx_up = dx_up;
  a.da = reinterpret_cast<float*>(ws + tl.da[k]);
  a.hr = write_hr ? reinterpret_cast<float*>(ws + tl.hr[k]) : nullptr;
  a.B = d.B; a.S = d.S[k]; a.period = k == d.L - 1 ? 1 : d.P[k]; a.layer = k; a.L = d.L;
  const size_t smem = tcr_bwd_smem(H, stg);
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(d.B + 127) / 128, TCR_THREADS, smem, st>>>(tm_st, tm_w, tm_da, a);
  };
  if (stg) go(tcrec_bwd_kernel<H, CAN_STAGE>); else go(tcrec_bwd_kernel<H, false>);
  ++*L.counter;
  return true;
}

// layer k of the memory, backward recurrence: st_k, dmemory, dx_up (dX of layer k+1, nullptr for the top layer) -> da_k
bool launch_tcrec_bwd(const Launch& L, const Dims& d, const TcrLayout& tl, int k, char* ws, const float* dmemory, const float* dx_up,
                      bool write_hr, cudaStream_t st) {
  if (d.H == 32) return tcr_bwd_go<32>(L, d, tl, k, ws, dmemory, dx_up, write_hr, st);
  if (d.H == 64) return tcr_bwd_go<64>(L, d, tl, k, ws, dmemory, dx_up, write_hr, st);
  return false;
}

// layer k of the memory, forward; inputs xh/xl of the layer must be in the workspace (layer 0: launch_tcr_split or the
// fused gather; layers above: written by the layer below)
bool launch_tcrec_fwd(const Launch& L, const Dims& d, const TcrLayout& tl, int k, char* ws, float* memory, cudaStream_t st) {
  const int DP = tl.DP[k];
  if (d.H == 32 && DP == 32) return tcr_fwd_go<32, 32>(L, d, tl, k, ws, memory, st);
  if (d.H == 32 && DP == 64) return tcr_fwd_go<32, 64>(L, d, tl, k, ws, memory, st);
  if (d.H == 64 && DP == 32) return tcr_fwd_go<64, 32>(L, d, tl, k, ws, memory, st);
  if (d.H == 64 && DP == 64) return tcr_fwd_go<64, 64>(L, d, tl, k, ws, memory, st);
  return false;
}

}  // namespace hpmn
