// tc_gemm.cu -- the non-recurrent halves of the GRU gate GEMMs on the 5th-gen tensor cores.
//
//   C[M,N] = A[M,K] * W[K,N] (+ bias)       M = B*S_k rows (up to 262 144), K,N in {32,48,64,96}
//
// used for the input projections P = X*W_x + b (the x-half of the two _Linear calls of
// /root/reference/code/util.py:88-107, hoisted out of the time loop) and for their adjoint dX = dA*W_x^T.
//
// tcgen05.mma kind::tf32, cta_group::1, M=128 rows per tile, accumulator in TMEM (128 lanes x N columns fp32),
// issued by one thread, completion through tcgen05.commit -> mbarrier, read back with tcgen05.ld.
// Precision: 3xTF32.  Each fp32 operand is split in registers into hi = rna_tf32(x), lo = rna_tf32(x - hi) and
// D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo accumulates in fp32 -- ~2^-20 relative per product, which keeps the
// 1e-4 parity budget through the recurrence that consumes P (single-pass tf32 would not).
// Operands reach shared memory from registers (the split needs a register pass anyway) in the canonical
// no-swizzle K-major core-matrix layout: 16-byte K-chunk c of row r lives at c*CHS + r*16, so a core matrix
// (8 rows x 16 B) is 128 contiguous bytes, SBO = 128 B, LBO = CHS (padded by 16 B to spread banks).
// The kernels are HBM-bound streams (A in, C out); several CTAs per SM overlap load / MMA / epilogue.
#include <stdlib.h>

#include "common.cuh"

namespace hpmn {

// ---- tcgen05 PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 3xTF32 operand split x = hi + lo.  cvt.rna.tf32.f32 has no native SASS on sm_100a -- it expands to a ~12-instruction
// LOP3/FSETP/SEL sequence, which made the operand producers of every kernel here issue-bound (r1: 416 instructions per
// 32-row stage and warp).  Integer form of the same rounding (nearest, ties away from zero: add half a tf32 ulp to the
// magnitude, clear the 13 low mantissa bits): 2 integer ops for hi, 1 exact FADD + 1 mask for lo (lo is truncated, its
// error is 2^-10 of a term that is already 2^-11 of x).
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x, float hi) { return __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u); }
// shared-memory matrix descriptor, no swizzle (layout_type 0), sm_100 version bit set
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// instruction descriptor: c=f32, a=b=tf32, K-major (0) or MN-major (1) operands, N>>3, M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------
// K is consumed in slabs of KS = 32 columns (one 128 x 32 operand buffer per CTA, so 3 CTAs fit an SM for every (K,N)
// and overlap each other's load / split / MMA / epilogue phases); the accumulator stays in TMEM across the slabs.
// Epilogue: TMEM -> registers (+bias) -> padded per-warp staging rows in the (now idle) operand buffer -> coalesced
// 512-byte global stores.  Storing straight from the TMEM register layout (thread = row) costs 32 L2 requests of
// 16 bytes per instruction and made the projection GEMM request-bound (r1: 2.2 TB/s).
template <int K, int N>
__global__ void __launch_bounds__(128, K % 32 == 0 ? 3 : 2)
tc_gemm_nn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ C, int64_t M) {
  constexpr int KS = K % 32 == 0 ? 32 : K;        // slab width (K = 48: one slab)
  constexpr int NSLAB = K / KS;
  static_assert(K % KS == 0 && KS % 8 == 0, "K must be a multiple of the slab width");
  constexpr int KC = KS / 4;                      // 16-byte chunks along K per slab
  constexpr int KCB = K / 4;                      // chunks of the whole weight operand
  constexpr int CHS_A = 128 * 16 + 16;            // bytes between K-chunks of the A slab (padded)
  constexpr int CHS_B = N * 16 + 16;
  constexpr int TCOLS = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  constexpr uint32_t IDESC = umma_idesc_tf32(128, N, 0, 0);
  constexpr int NH = N > 48 ? N / 2 : N;          // columns per epilogue pass
  constexpr int NPASS = N / NH;
  constexpr int SROW = NH + 4;                    // staging row stride (floats): an odd number of float4s
  static_assert(NH % 16 == 0 && ((SROW / 4) & 1) == 1, "epilogue pass width");
  static_assert(128 * SROW * 4 <= 2 * KC * CHS_A, "staging must fit in the operand buffer");
  pdl_trigger();                                  // a dependent wavefront kernel may set itself up while this grid drains
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sAhi = smem_raw;
  unsigned char* sAlo = sAhi + KC * CHS_A;
  unsigned char* sBhi = sAlo + KC * CHS_A;
  unsigned char* sBlo = sBhi + KCB * CHS_B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBlo + KCB * CHS_B);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(tslot, TCOLS);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  // weights: B[n][k] = W[k][n], split once per CTA
  for (int e = tid; e < K * N; e += 128) {
    const int k = e / N, n = e % N;
    const float w = __ldg(W + e);
    const float hi = tf32_rna(w), lo = tf32_lo(w, hi);
    const int off = (k >> 2) * CHS_B + n * 16 + (k & 3) * 4;
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = lo;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  const int64_t tiles = (M + 127) / 128;
  float4 areg[KC];                                // this thread's 16-byte chunks of the current slab
  auto load_slab = [&](int64_t tile, int s) {
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      const int e = tid + 128 * i;
      const int row = e / KC, c = e % KC;
      const int64_t m = tile * 128 + row;
      areg[i] = m < M ? ldg_nc_f4(reinterpret_cast<const float4*>(A + m * lda) + s * KC + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  int64_t tile = blockIdx.x;
  if (tile < tiles) load_slab(tile, 0);
  uint32_t phase = 0;
  for (; tile < tiles; tile += gridDim.x) {
#pragma unroll 1
    for (int s = 0; s < NSLAB; ++s) {
      // registers -> (hi, lo) -> shared, canonical K-major layout.  The previous user of the buffer (the MMAs of the
      // previous slab, or the staging reads of the previous tile's epilogue) is complete: see the waits below.
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        const int e = tid + 128 * i;
        const int row = e / KC, c = e % KC;
        const float4 v = areg[i];
        float4 hi, lo;
        hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
        lo.x = tf32_lo(v.x, hi.x); lo.y = tf32_lo(v.y, hi.y); lo.z = tf32_lo(v.z, hi.z); lo.w = tf32_lo(v.w, hi.w);
        *reinterpret_cast<float4*>(sAhi + c * CHS_A + row * 16) = hi;
        *reinterpret_cast<float4*>(sAlo + c * CHS_A + row * 16) = lo;
      }
      fence_proxy_async();                        // generic-proxy smem writes -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t ah = smem_u32(sAhi), al = smem_u32(sAlo);
        const uint32_t bh = smem_u32(sBhi) + s * KC * CHS_B, bl = smem_u32(sBlo) + s * KC * CHS_B;
#pragma unroll
        for (int q = 0; q < KS / 8; ++q) {
          const uint64_t dah = umma_desc(ah + q * 2 * CHS_A, CHS_A, 128), dal = umma_desc(al + q * 2 * CHS_A, CHS_A, 128);
          const uint64_t dbh = umma_desc(bh + q * 2 * CHS_B, CHS_B, 128), dbl = umma_desc(bl + q * 2 * CHS_B, CHS_B, 128);
          tc_mma_tf32(tmem, dah, dbh, IDESC, (s | q) != 0);
          tc_mma_tf32(tmem, dal, dbh, IDESC, 1);
          tc_mma_tf32(tmem, dah, dbl, IDESC, 1);
        }
        tc_commit(bar);                           // arrives when every MMA above has finished reading smem / writing TMEM
      }
      // the next slab's global loads fly during the MMA (and, after the last slab, the epilogue)
      if (s + 1 < NSLAB) load_slab(tile, s + 1);
      else if (tile + gridDim.x < tiles) load_slab(tile + gridDim.x, 0);
      mbar_wait(bar, phase);                      // every thread: the operand buffer is free, TMEM holds slabs 0..s
      phase ^= 1;
      tc_fence_after();
    }
    // epilogue: TMEM lane = row.  Each warp stages its own 32 rows (padded) and writes them back coalesced.
    float* stage = reinterpret_cast<float*>(smem_raw) + warp * 32 * SROW;
    const int64_t m0 = tile * 128 + warp * 32;
#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
#pragma unroll
      for (int cb = 0; cb < NH / 16; ++cb) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + pass * NH + cb * 16, v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (bias != nullptr) {
            const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + pass * NH + cb * 16) + q);
            o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
          }
          *reinterpret_cast<float4*>(stage + lane * SROW + cb * 16 + q * 4) = o;
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < NH / 4; ++it) {       // 32 rows x NH/4 float4s, 32 consecutive float4s per instruction
        const int e = it * 32 + lane;
        const int row = e / (NH / 4), c = e % (NH / 4);
        const float4 o = *reinterpret_cast<const float4*>(stage + row * SROW + c * 4);
        if (m0 + row < M) *reinterpret_cast<float4*>(C + (m0 + row) * N + pass * NH + c * 4) = o;
      }
      __syncwarp();                               // staging rows free for the next pass
    }
    tc_fence_before();
    __syncthreads();                              // TMEM drained and staging reads done before the next tile's writes / MMAs
  }
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// ---------------------------------------------------------------------------------------------------
// GRU weight gradients on tensor cores:  D[I,N] = sum_m At[m,I] * dA[m,N]   (reduction over all B*S_k rows)
//   I (MMA M = 128) = [x (DINP) | h_prev (32) | r*h_prev (32) | zero pad ... | ones]   N = 96 = da_r | da_u | da_c
//   dWg = rows {x,h} x cols {r,u};  dWc = rows {x,r*h} x cols {c};  the ones row (I = 127) yields the bias gradients.
// The reduction dimension (rows of global memory) is the MMA K dimension, so both operands are transposed on the
// way into shared memory: element (feature i, row k) lives at (k/4)*CHS + i*16 + (k%4)*4 -- the same K-major
// core-matrix layout as tc_gemm_nn (kind::tf32 with MN-major descriptors returned zeros on this part, so the
// transpose is done by the producers: lanes map to rows, and because CHS = 16 mod 128 bytes the 32 scalar stores of
// a warp hit 32 different banks).
// Warp-specialised: 4 producer warps (global -> registers -> hi/lo split, r*h -> shared, 3 stages of 32 rows),
// 1 MMA warp (single issuing thread, 3xTF32); the accumulator stays in TMEM for the CTA's whole row range and is
// scattered into the TF weight layout once at the end.
// ---------------------------------------------------------------------------------------------------
constexpr int WK = 32;                 // rows per stage (= MMA K per stage)
#ifndef HPMN_WNS
#define HPMN_WNS 3      // default number of stages (wgrad_stages())
#endif
constexpr int WNS = HPMN_WNS;          // stages
constexpr int WCHS_A = 128 * 16 + 16;  // bytes between 4-row K chunks of the feature tile (padded)
constexpr int WCHS_B = 96 * 16 + 16;
constexpr int W_AT = (WK / 4) * WCHS_A;
constexpr int W_B = (WK / 4) * WCHS_B;
constexpr int W_STAGE = 2 * W_AT + 2 * W_B;

// scatter 4 consecutive features of row k (hi and lo parts) into a K-major tile
__device__ __forceinline__ void put_t(unsigned char* hi_base, unsigned char* lo_base, int chs, int feat0, int k, float4 v) {
  const int off = (k >> 2) * chs + feat0 * 16 + (k & 3) * 4;
  const float hx = tf32_rna(v.x), hy = tf32_rna(v.y), hz = tf32_rna(v.z), hw = tf32_rna(v.w);
  *reinterpret_cast<float*>(hi_base + off) = hx;
  *reinterpret_cast<float*>(hi_base + off + 16) = hy;
  *reinterpret_cast<float*>(hi_base + off + 32) = hz;
  *reinterpret_cast<float*>(hi_base + off + 48) = hw;
  *reinterpret_cast<float*>(lo_base + off) = tf32_lo(v.x, hx);
  *reinterpret_cast<float*>(lo_base + off + 16) = tf32_lo(v.y, hy);
  *reinterpret_cast<float*>(lo_base + off + 32) = tf32_lo(v.z, hz);
  *reinterpret_cast<float*>(lo_base + off + 48) = tf32_lo(v.w, hw);
}

constexpr int WPW = 8;                 // producer warps (2 per SM sub-partition: the convert/transposing stores are issue-bound)

// one launch serves every layer that shares the padded input width: CTA -> (layer, row range)
// H <= 32: one problem per layer.  H = 64 (tensor-core recurrence, tcrec.cu): four sub-problems per layer -- hidden rows
// [32a, 32a+32) of h_prev / r*h_prev against gate columns [32b, 32b+32) of r, u, c -- each with the same 128 x 96 tile shape;
// the x rows and the bias row are taken from a = 0 only.
struct WgradProb {
  const float* xin; const float* st; const float* da;
  float* dWg; float* dbg; float* dWc; float* dbc;
  int64_t ldx, M, rows_per_cta;
  int S, Din, cta_begin, hrow0;        // Din: x rows of this problem's tile; hrow0: first h row of the TF kernel (the layer's real Din)
  int use_x, use_h, use_bias, pad_;    // which row groups of the tile this problem contributes
  int st_stride, h_off, r_off;         // floats: state row stride (4 Htot), offset of the h / r slice inside a row
  int da_stride, da_gate, da_off;      // floats: dA row stride (3 Htot), stride between gates (Htot), offset of the slice inside a gate
  int Htot, Hsub;                      // hidden size of the layer, width of this slice (<= 32)
};
struct WgradBatch { int n, H, producer_fence, pad_; WgradProb p[HPMN_MAX_LAYERS]; };

template <int DINP, int NS>
__global__ void __launch_bounds__(32 * (WPW + 1))
tc_wgrad_kernel(const __grid_constant__ WgradBatch batch) {
  static_assert(DINP + 64 < 128, "the ones row needs a free feature slot");
  int pi = 0;
  while (pi + 1 < batch.n && (int)blockIdx.x >= batch.p[pi + 1].cta_begin) ++pi;
  const WgradProb& P = batch.p[pi];
  const float* __restrict__ xin = P.xin; const float* __restrict__ st = P.st; const float* __restrict__ da = P.da;
  float* __restrict__ dWg = P.dWg; float* __restrict__ dbg = P.dbg; float* __restrict__ dWc = P.dWc; float* __restrict__ dbc = P.dbc;
  const int64_t ldx = P.ldx, M = P.M, rows_per_cta = P.rows_per_cta;
  const int S = P.S, Din = P.Din, H = P.Hsub, Htot = P.Htot;
  const int cta = (int)blockIdx.x - P.cta_begin;
  constexpr int XC = DINP / 4;           // feature chunks of x
  constexpr int NXU = XC;                // work units (16 rows x 2 chunks) in the x part
  constexpr int NX = (NXU + WPW - 1) / WPW;
  constexpr uint32_t IDESC = umma_idesc_tf32(128, 96, 0, 0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + NS * W_STAGE);
  uint64_t* empty = full + NS;
  uint64_t* done = empty + NS;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t mbeg = (int64_t)cta * rows_per_cta;
  const int64_t mend = mbeg + rows_per_cta < M ? mbeg + rows_per_cta : M;
  const int nst = mend > mbeg ? (int)((mend - mbeg + WK - 1) / WK) : 0;

  if (warp == WPW) tmem_alloc(tslot, 128);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 32 * WPW); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  // pad features [DINP+64, 128): zeros, except feature 127 = 1.0 (hi part) -> bias gradients
  for (int e = tid; e < NS * (64 - DINP) * WK; e += 32 * (WPW + 1)) {
    const int stg = e / ((64 - DINP) * WK), r = e % ((64 - DINP) * WK);
    const int feat = DINP + 64 + r / WK, k = r % WK;
    unsigned char* base = smem_raw + stg * W_STAGE;
    const int off = (k >> 2) * WCHS_A + feat * 16 + (k & 3) * 4;
    *reinterpret_cast<float*>(base + off) = feat == 127 ? 1.f : 0.f;   // hi
    *reinterpret_cast<float*>(base + W_AT + off) = 0.f;                // lo
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  if (warp < WPW) {
    // ================= producers =================
    struct Regs { float4 x[NX], h[1], r[1], d[3]; };
    Regs bufA, bufB;
    // work unit u = warp + 4*i covers 16 rows x 2 adjacent 16-byte chunks: every 32-byte sector a warp touches is
    // used completely (no reliance on L1 hits), and the 32 scalar stores of put_t land in 32 different banks
    const int lk = lane >> 1, lc = lane & 1;
    // Every unit of a thread sits on the same row of a stage (WPW is even): one running row counter, one validity test and
    // one running pointer per unit, bumped by a stage per load() call (the calls walk the stages in order).
    static_assert(WPW % 2 == 0, "units of a thread must share their row");
    const int nrows = (int)(mend > mbeg ? mend - mbeg : 0);
    int rrow = 16 * (warp & 1) + lk;                         // row inside this CTA's range of the next load
    int hrem = (int)((mbeg + rrow) % S);                     // step index inside the sample: 0 = no predecessor row (zero state)
    const float4* px[NX]; const float4* pd[3];
#pragma unroll
    for (int i = 0; i < NX; ++i) { const int u = warp + WPW * i; px[i] = reinterpret_cast<const float4*>(xin + (mbeg + rrow) * ldx) + 2 * (u >> 1) + lc; }
    const int sst4 = P.st_stride / 4, sda4 = P.da_stride / 4, r_rel4 = (P.r_off - P.h_off) / 4;
    const float4* ph = reinterpret_cast<const float4*>(st + (mbeg + rrow) * P.st_stride + P.h_off) + 2 * (warp >> 1) + lc;   // row m; h_prev is one row up
#pragma unroll
    for (int i = 0; i < 3; ++i) {       // 16-byte chunk c of the [r | u | c] slice: gate c / 8, chunk c % 8 inside the gate
      const int u = warp + WPW * i, c = 2 * (u >> 1) + lc;
      pd[i] = reinterpret_cast<const float4*>(da + (mbeg + rrow) * P.da_stride + (c >> 3) * P.da_gate + P.da_off) + (c & 7);
    }
    const int64_t sx = (int64_t)WK * ldx / 4;                // float4 strides of one stage
    auto load = [&](Regs& R, int) {
      const bool ok = rrow < nrows;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const int u = warp + WPW * i;
        R.x[i] = (ok && u < NXU) ? __ldg(px[i]) : z;
        px[i] += sx;
      }
      R.h[0] = (ok && hrem != 0) ? __ldg(ph - sst4) : z;
      R.r[0] = ok ? __ldg(ph + r_rel4) : z;
      ph += WK * sst4;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        R.d[i] = ok ? __ldg(pd[i]) : z;
        pd[i] += WK * sda4;
      }
      rrow += WK;
      hrem += WK;
      while (hrem >= S) hrem -= S;
    };
    auto store = [&](const Regs& R, int sidx) {
      const int slot = sidx % NS;
      if (sidx >= NS) mbar_wait(&empty[slot], (uint32_t)((sidx / NS) - 1) & 1u);     // MMAs that read this slot are done
      unsigned char* base = smem_raw + slot * W_STAGE;
      unsigned char* At_hi = base;
      unsigned char* At_lo = base + W_AT;
      unsigned char* B_hi = base + 2 * W_AT;
      unsigned char* B_lo = base + 2 * W_AT + W_B;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const int u = warp + WPW * i, c = 2 * (u >> 1) + lc, k = 16 * (u & 1) + lk;
        if (u < NXU) put_t(At_hi, At_lo, WCHS_A, 4 * c, k, R.x[i]);
      }
#pragma unroll
      for (int i = 0; i < 1; ++i) {
        const int u = warp + WPW * i, c = 2 * (u >> 1) + lc, k = 16 * (u & 1) + lk;
        put_t(At_hi, At_lo, WCHS_A, DINP + 4 * c, k, R.h[i]);
        put_t(At_hi, At_lo, WCHS_A, DINP + 32 + 4 * c, k,
              make_float4(R.h[i].x * R.r[i].x, R.h[i].y * R.r[i].y, R.h[i].z * R.r[i].z, R.h[i].w * R.r[i].w));
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int u = warp + WPW * i, c = 2 * (u >> 1) + lc, k = 16 * (u & 1) + lk;
        put_t(B_hi, B_lo, WCHS_B, 4 * c, k, R.d[i]);
      }
      // The generic -> async proxy fence sits on the consumer side of the (release) arrive / (acquire) wait pair: here it
      // compiles to MEMBAR.ALL.CTA, which would also wait for this thread's prefetched global loads of the NEXT three
      // stages -- every stage would then cost one full DRAM latency (r1: 1.3 us per stage, 2.8 TB/s).
      if (batch.producer_fence) fence_proxy_async();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[slot])) : "memory");
    };
    // four register buffers: ~80 KB of global loads in flight per SM (6.5 TB/s x ~1.5 us needs ~66 KB per SM)
    Regs bufC, bufD;
    load(bufA, 0);
    load(bufB, 1);
    load(bufC, 2);
    load(bufD, 3);
    for (int sidx = 0; sidx < nst; sidx += 4) {
      store(bufA, sidx);
      load(bufA, sidx + 4);
      if (sidx + 1 < nst) { store(bufB, sidx + 1); load(bufB, sidx + 5); }
      if (sidx + 2 < nst) { store(bufC, sidx + 2); load(bufC, sidx + 6); }
      if (sidx + 3 < nst) { store(bufD, sidx + 3); load(bufD, sidx + 7); }
    }
    // ================= epilogue (warps 0-3): TMEM lane = input feature =================
    if (warp < 4) {
    mbar_wait(done, 0);
    tc_fence_after();
    const int i = warp * 32 + lane;
    int kind, row;                       // kind 0: x row, 1: h row (gates only), 2: r*h row (candidate only), 3: bias, -1: pad
    if (i < DINP) { kind = (i < Din && P.use_x) ? 0 : -1; row = i; }
    else if (i < DINP + 32) { kind = ((i - DINP) < H && P.use_h) ? 1 : -1; row = P.hrow0 + P.h_off + (i - DINP); }
    else if (i < DINP + 64) { kind = ((i - DINP - 32) < H && P.use_h) ? 2 : -1; row = P.hrow0 + P.h_off + (i - DINP - 32); }
    else { kind = (i == 127 && P.use_bias) ? 3 : -1; row = 0; }
#pragma unroll
    for (int cb = 0; cb < 6; ++cb) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + cb * 16, v);
      if (nst == 0 || kind < 0) continue;
      const int g = cb >> 1;             // 0: r, 1: u, 2: c
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int j = (cb & 1) * 16 + q;
        if (j >= H) continue;
        const int col = P.da_off + j;                                    // column inside the gate
        if (kind == 3) {
          if (g < 2) atomicAdd(dbg + g * Htot + col, v[q]); else atomicAdd(dbc + col, v[q]);
        } else if (g < 2) {
          if (kind != 2) atomicAdd(dWg + (int64_t)row * 2 * Htot + g * Htot + col, v[q]);
        } else {
          if (kind != 1) atomicAdd(dWc + (int64_t)row * Htot + col, v[q]);
        }
      }
    }
    tc_fence_before();
    }
  } else {
    // ================= MMA warp =================
    if (lane == 0) {
      for (int sidx = 0; sidx < nst; ++sidx) {
        const int slot = sidx % NS;
        mbar_wait(&full[slot], (uint32_t)(sidx / NS) & 1u);
        fence_proxy_async();                  // producers' generic-proxy stores (ordered by the barrier) -> async proxy
        tc_fence_after();
        const uint32_t base = smem_u32(smem_raw + slot * W_STAGE);
        const uint32_t ah = base, al = base + W_AT, bh = base + 2 * W_AT, bl = base + 2 * W_AT + W_B;
#pragma unroll
        for (int ks = 0; ks < WK / 8; ++ks) {
          const uint64_t dah = umma_desc(ah + ks * 2 * WCHS_A, WCHS_A, 128), dal = umma_desc(al + ks * 2 * WCHS_A, WCHS_A, 128);
          const uint64_t dbh = umma_desc(bh + ks * 2 * WCHS_B, WCHS_B, 128), dbl = umma_desc(bl + ks * 2 * WCHS_B, WCHS_B, 128);
          tc_mma_tf32(tmem, dah, dbh, IDESC, (sidx | ks) != 0);
          tc_mma_tf32(tmem, dal, dbh, IDESC, 1);
          tc_mma_tf32(tmem, dah, dbl, IDESC, 1);
        }
        tc_commit(&empty[slot]);
      }
      tc_commit(done);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == WPW) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

// HPMN_WGRAD_FENCE=1: producers execute the proxy fence themselves (the r1 v1-v6 behaviour)
static int wgrad_producer_fence() { static int v = -1; if (v < 0) { const char* e = getenv("HPMN_WGRAD_FENCE"); v = (e && e[0] == '1') ? 1 : 0; } return v; }

static void wgrad_queue_add(WgradBatch& b, int& ctas, int sms, const float* xin, int64_t ldx, const float* st, const float* da,
                            float* dWg, float* dbg, float* dWc, float* dbc, int64_t M, int S, int Din, int64_t total_rows,
                            int Htot = 0, int a = 0, int bq = 0) {
  WgradProb& P = b.p[b.n++];
  P.xin = xin; P.st = st; P.da = da; P.dWg = dWg; P.dbg = dbg; P.dWc = dWc; P.dbc = dbc; P.ldx = ldx; P.M = M; P.S = S; P.Din = Din;
  P.hrow0 = Din; P.use_x = P.use_h = P.use_bias = 1; P.pad_ = 0;
  if (Htot <= HP) {                      // wavefront / per-layer layout: rows padded to 32 lanes, H real columns
    P.Htot = Htot > 0 ? Htot : b.H; P.Hsub = P.Htot; P.st_stride = ST; P.h_off = 0; P.r_off = HP; P.da_stride = G3; P.da_gate = HP;
    P.da_off = 0;
  } else {                               // tensor-core recurrence layout: unpadded rows, slice (a, bq)
    P.Htot = Htot; P.Hsub = 32; P.st_stride = 4 * Htot; P.h_off = 32 * a; P.r_off = Htot + 32 * a; P.da_stride = 3 * Htot;
    P.da_gate = Htot; P.da_off = 32 * bq;
  }
  const int64_t stages = (M + WK - 1) / WK;
  int64_t want = (int64_t)((double)sms * (double)M / (double)total_rows + 0.5);   // CTAs proportional to the layer's rows
  if (want > stages / 4) want = stages / 4;                                        // >= 4 stages per CTA
  if (want < 1) want = 1;
  P.rows_per_cta = ((stages + want - 1) / want) * WK;
  const int n = (int)((M + P.rows_per_cta - 1) / P.rows_per_cta);
  P.cta_begin = ctas;
  ctas += n;
}

// Stages of the producer -> MMA ring and the shared-memory carve-out the kernel asks for.  Default: 2 stages (116 KB) and the
// smallest carve-out that holds them (135 KB), which leaves ~120 KB of the SM's unified array to L1.  The producers keep
// ~80 KB of global loads in flight per SM, and every 128-byte line in flight holds a line of L1 (the L1 HIT rate is 0 % either
// way: nothing is re-read): with the maximal carve-out (233 KB) 23 KB of L1 cap the loads in flight -- ncu: DRAM throughput
// 33 % instead of 44 %, short-scoreboard stalls 2.7 instead of 1.05 per issue, 142 instead of 105 us (profiles/r2_timeline.md).
// In the step: 3 stages / driver-chosen split 121 us, 3 stages / maximal carve-out 162 us, 2 stages / minimal carve-out 117 us.
// HPMN_WGRAD_STAGES=3 and HPMN_WGRAD_CARVEOUT=<percent, -1 = driver default> override.
static int wgrad_stages() {
  static const int ns = [] { const char* e = getenv("HPMN_WGRAD_STAGES"); const int v = e ? atoi(e) : 2; return v == 3 ? 3 : 2; }();
  return ns;
}
template <int DINP>
static void wgrad_go(int ctas, cudaStream_t st, const WgradBatch& b) {
  static const int carve = [] { const char* e = getenv("HPMN_WGRAD_CARVEOUT"); return e ? atoi(e) : 50; }();
  const int ns = wgrad_stages();
  const size_t smem = (size_t)ns * W_STAGE + 128;
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (carve >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    kern<<<ctas, 32 * (WPW + 1), smem, st>>>(b);
  };
  if (ns == 2) go(tc_wgrad_kernel<DINP, 2>); else go(tc_wgrad_kernel<DINP, 3>);
}

// H = 64 (layouts of tcrec.cu: state rows [h|r|u|c] x 64, dA rows [r|u|c] x 64).  The kernel's tile is [x (32|48) | h (32) | r*h (32) |
// ones] x [r|u|c] (3 x 32), so a layer is cut into slices: hidden rows 32a.. against gate columns 32bq..; the x rows and the bias
// row are taken from the a = 0 slices; the layers above layer 0 have 64 input rows and get one more pair of slices for inputs [32,64).
bool launch_tc_wgrad_wide(const Launch& L, const Dims& d, const float* const* xin, const int64_t* ldx, const float* const* st,
                          const float* const* da, float* const* dWg, float* const* dbg, float* const* dWc, float* const* dbc,
                          cudaStream_t st_) {
  if (d.H != 64) return false;
  const int dp0 = ((d.D + 15) / 16) * 16;
  if (dp0 != 32 && dp0 != 48) return false;
  for (int pass = 0; pass < 2; ++pass) {          // pass 0: layer 0 (tile with 32 or 48 x rows); pass 1: the layers above (32 x rows)
    for (int k0 = pass == 0 ? 0 : 1; k0 < (pass == 0 ? 1 : d.L); k0 += 2) {      // at most 2 layers (12 slices) per launch
      WgradBatch b; b.n = 0; b.H = 32; b.producer_fence = wgrad_producer_fence();
      int ctas = 0; int64_t rows = 0;
      const int k1 = pass == 0 ? 1 : (k0 + 2 < d.L ? k0 + 2 : d.L);
      for (int k = k0; k < k1; ++k) rows += (pass == 0 ? 4 : 6) * (int64_t)d.B * d.S[k];
      for (int k = k0; k < k1; ++k) {
        const int64_t M = (int64_t)d.B * d.S[k];
        const int Dreal = d.Din[k];
        for (int a = 0; a < 2; ++a)
          for (int bq = 0; bq < 2; ++bq) {
            wgrad_queue_add(b, ctas, L.sms, xin[k], ldx[k], st[k], da[k], dWg[k], dbg[k], dWc[k], dbc[k], M, d.S[k],
                            k == 0 ? Dreal : 32, rows, 64, a, bq);
            WgradProb& P = b.p[b.n - 1];
            P.hrow0 = Dreal; P.use_x = a == 0; P.use_bias = a == 0;
          }
        if (k > 0)
          for (int bq = 0; bq < 2; ++bq) {        // inputs [32,64) of the layer below: x rows only
            wgrad_queue_add(b, ctas, L.sms, xin[k] + 32, ldx[k], st[k], da[k], dWg[k] + (int64_t)32 * 2 * 64, dbg[k],
                            dWc[k] + (int64_t)32 * 64, dbc[k], M, d.S[k], 32, rows, 64, 0, bq);
            WgradProb& P = b.p[b.n - 1];
            P.hrow0 = Dreal; P.use_h = 0; P.use_bias = 0;
          }
      }
      if (pass == 0 && dp0 == 48) {
        wgrad_go<48>(ctas, st_, b);
      } else {
        wgrad_go<32>(ctas, st_, b);
      }
      ++*L.counter;
    }
  }
  return true;
}

bool launch_tc_wgrad_all(const Launch& L, const Dims& d, const float* const* xin, const int64_t* ldx, const float* const* st,
                         const float* const* da, float* const* dWg, float* const* dbg, float* const* dWc, float* const* dbc,
                         cudaStream_t st_) {
  for (int k = 0; k < d.L; ++k)
    if (d.DinP[k] != 32 && d.DinP[k] != 48) return false;
  WgradBatch b32, b48; b32.n = b48.n = 0; b32.H = b48.H = d.H;
  b32.producer_fence = b48.producer_fence = wgrad_producer_fence();
  int c32 = 0, c48 = 0;
  int64_t rows32 = 0, rows48 = 0;
  for (int k = 0; k < d.L; ++k) (d.DinP[k] == 32 ? rows32 : rows48) += (int64_t)d.B * d.S[k];
  for (int k = 0; k < d.L; ++k) {
    const int64_t M = (int64_t)d.B * d.S[k];
    if (d.DinP[k] == 32) wgrad_queue_add(b32, c32, L.sms, xin[k], ldx[k], st[k], da[k], dWg[k], dbg[k], dWc[k], dbc[k], M, d.S[k], d.Din[k], rows32);
    else wgrad_queue_add(b48, c48, L.sms, xin[k], ldx[k], st[k], da[k], dWg[k], dbg[k], dWc[k], dbc[k], M, d.S[k], d.Din[k], rows48);
  }
  if (b48.n) {
    wgrad_go<48>(c48, st_, b48);
    ++*L.counter;
  }
  if (b32.n) {
    wgrad_go<32>(c32, st_, b32);
    ++*L.counter;
  }
  return true;
}

bool launch_tc_wgrad(const Launch& L, const Dims& d, int k, const float* xin, int64_t ldx, const float* st, const float* da,
                     float* dWg, float* dbg, float* dWc, float* dbc, cudaStream_t st_) {
  const int DinP = d.DinP[k];
  if (DinP != 32 && DinP != 48) return false;
  WgradBatch b; b.n = 0; b.H = d.H; b.producer_fence = wgrad_producer_fence();
  int ctas = 0;
  const int64_t M = (int64_t)d.B * d.S[k];
  wgrad_queue_add(b, ctas, L.sms, xin, ldx, st, da, dWg, dbg, dWc, dbc, M, d.S[k], d.Din[k], M);
  if (DinP == 32) wgrad_go<32>(ctas, st_, b); else wgrad_go<48>(ctas, st_, b);
  ++*L.counter;
  return true;
}

template <int K, int N>
static void launch_one(const Launch& L, const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t M,
                       cudaStream_t st) {
  constexpr int KC = (K % 32 == 0 ? 32 : K) / 4, KCB = K / 4;   // one A slab, the whole weight operand
  const size_t smem = (size_t)2 * KC * (128 * 16 + 16) + (size_t)2 * KCB * (N * 16 + 16) + 64;
  auto kern = tc_gemm_nn_kernel<K, N>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t tiles = (M + 127) / 128;
  int per_sm = (int)((227 * 1024) / (smem + 1024));   // resident CTAs: shared memory, then the 512-column TMEM budget
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  // grid = resident CTAs (a multiple of the SM count): tiles are dealt round-robin, so every SM ends within one tile of
  // the others even though CTAs differ by one tile
  const int64_t cap = (int64_t)L.sms * per_sm;
  const int grid = (int)(tiles < cap ? tiles : cap);
  kern<<<grid > 0 ? grid : 1, 128, smem, st>>>(A, lda, W, bias, C, M);
  ++*L.counter;
}

bool launch_tc_gemm_nn(const Launch& L, const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t M,
                       int N, int K, cudaStream_t st) {
#define HPMN_TC_CASE(KK, NN) if (K == KK && N == NN) { launch_one<KK, NN>(L, A, lda, W, bias, C, M, st); return true; }
  HPMN_TC_CASE(32, 96) HPMN_TC_CASE(48, 96) HPMN_TC_CASE(64, 96)
  HPMN_TC_CASE(96, 32) HPMN_TC_CASE(96, 48) HPMN_TC_CASE(96, 64)
  HPMN_TC_CASE(192, 32) HPMN_TC_CASE(192, 48) HPMN_TC_CASE(192, 64)          // dX at H = 64 (tensor-core recurrence)
#undef HPMN_TC_CASE
  return false;
}

}  // namespace hpmn
