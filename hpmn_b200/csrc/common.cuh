// common.cuh -- shared host/device helpers for libhpmn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hpmn_b200.h"

namespace hpmn {

constexpr int HP = 32;        // hidden width padded to one warp (warp-per-sample kernels: H <= 32; H = 64: tcrec.cu only)
constexpr int G3 = 3 * HP;    // r | u | c columns of a packed gate row
constexpr int ST = 4 * HP;    // h | r | u | c columns of a saved state row
constexpr int ATT1 = 80;      // code/hpmn.py:137
constexpr int ATT2 = 40;      // code/hpmn.py:138
constexpr int FC1 = 200;      // code/hpmn.py:191
constexpr int FC2 = 80;       // code/hpmn.py:193
constexpr float BN_EPS = 1e-3f;
constexpr float LOGLOSS_EPS = 1e-7f;

struct hpmn_ctx_impl;

// ---------------------------------------------------------------------------------------------
// derived dimensions
// ---------------------------------------------------------------------------------------------
struct Dims {
  int B, T, Tpad, F, E, D, H, L, hops, R;   // R = H + D (head input)
  int S[HPMN_MAX_LAYERS];                     // steps of layer k
  int Din[HPMN_MAX_LAYERS];                   // real input width of layer k
  int DinP[HPMN_MAX_LAYERS];                  // padded input width (D for k=0, HP above)
  int P[HPMN_MAX_LAYERS];                     // period of layer k (k < L-1), 1 for the top layer
  int64_t rows_total;                         // sum_k B*S_k
  bool ok;
};

inline Dims make_dims(const hpmn_shape* s) {
  Dims d; memset(&d, 0, sizeof(d));
  d.ok = false;
  if (!s) return d;
  if (s->B <= 0 || s->T <= 0 || s->F <= 0 || s->E <= 0 || (s->E & 3) || s->H <= 0 || (s->H > HP && s->H != 64)) return d;
  if (s->L <= 0 || s->L > HPMN_MAX_LAYERS || s->hops <= 0 || s->hops > HPMN_MAX_HOPS) return d;
  if (s->front_pad < 0 || s->last_offset < 1 || s->V <= 0) return d;
  d.B = s->B; d.T = s->T; d.Tpad = s->T + s->front_pad; d.F = s->F; d.E = s->E; d.D = s->F * s->E;
  d.H = s->H; d.L = s->L; d.hops = s->hops; d.R = d.H + d.D;
  if (s->last_offset > d.Tpad) return d;
  if (d.D > 256) return d;
  int steps = d.Tpad;
  for (int k = 0; k < d.L; ++k) {
    d.S[k] = steps;
    d.Din[k] = k == 0 ? d.D : d.H;
    d.DinP[k] = k == 0 ? d.D : HP;
    d.rows_total += (int64_t)d.B * steps;
    if (k < d.L - 1) {
      int p = s->periods[k];
      if (p <= 0 || steps % p) return d;
      d.P[k] = p;
      steps /= p;
    } else {
      d.P[k] = 1;
    }
  }
  d.ok = true;
  return d;
}

// ---------------------------------------------------------------------------------------------
// flat parameter layout (see include/hpmn_b200.h)
// ---------------------------------------------------------------------------------------------
struct ParamLayout {
  int64_t Wg[HPMN_MAX_LAYERS], bg[HPMN_MAX_LAYERS], Wc[HPMN_MAX_LAYERS], bc[HPMN_MAX_LAYERS];
  int64_t Wq, bq, Hmap;
  int64_t A1[HPMN_MAX_HOPS], a1[HPMN_MAX_HOPS], A2[HPMN_MAX_HOPS], a2[HPMN_MAX_HOPS], A3[HPMN_MAX_HOPS], a3[HPMN_MAX_HOPS];
  int64_t gamma, beta, F1, f1, F2, f2, F3, f3;
  int64_t total;
  int ntensors;
};

inline int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

// visit(offset, size) is called once per tensor in canonical order
template <class Visit>
inline ParamLayout make_param_layout(const Dims& d, Visit visit) {
  ParamLayout p; memset(&p, 0, sizeof(p));
  int64_t off = 0; int n = 0;
  auto take = [&](int64_t sz) { int64_t o = off; visit(o, sz); off = align4(off + sz); ++n; return o; };
  for (int k = 0; k < d.L; ++k) {
    int in = d.Din[k] + d.H;
    p.Wg[k] = take((int64_t)in * 2 * d.H); p.bg[k] = take(2 * d.H);
    p.Wc[k] = take((int64_t)in * d.H);     p.bc[k] = take(d.H);
  }
  p.Wq = take((int64_t)d.D * d.H); p.bq = take(d.H); p.Hmap = take((int64_t)d.H * d.H);
  for (int h = 0; h < d.hops; ++h) {
    p.A1[h] = take((int64_t)4 * d.H * ATT1); p.a1[h] = take(ATT1);
    p.A2[h] = take((int64_t)ATT1 * ATT2);    p.a2[h] = take(ATT2);
    p.A3[h] = take(ATT2);                    p.a3[h] = take(1);
  }
  p.gamma = take(d.R); p.beta = take(d.R);
  p.F1 = take((int64_t)d.R * FC1); p.f1 = take(FC1);
  p.F2 = take((int64_t)FC1 * FC2); p.f2 = take(FC2);
  p.F3 = take(FC2); p.f3 = take(1);
  p.total = off; p.ntensors = n;
  return p;
}
inline ParamLayout make_param_layout(const Dims& d) { return make_param_layout(d, [](int64_t, int64_t) {}); }

// ---------------------------------------------------------------------------------------------
// workspace layout (byte offsets, 256-B aligned)
// ---------------------------------------------------------------------------------------------
struct WsLayout {
  size_t ids, labels;                 // device staging for the *_host entry points
  size_t x;                           // [B,Tpad,D]
  size_t pw;                          // packed weights (see PackLayout), rebuilt every call
  size_t proj[HPMN_MAX_LAYERS];       // [B,S_k,3,HP]  input projections; reused as da in bwd
  size_t st[HPMN_MAX_LAYERS];         // [B,S_k,4,HP]  per-step state row h | r | u | c  (512 B, one bulk store)
  size_t dxk[HPMN_MAX_LAYERS];        // k>=1: [B,S_k,HP] gradient wrt layer input; k=0: [B,Tpad,D]
  size_t memory, dmemory;             // [B,L,H]
  size_t att_q, att_dq;               // [hops+1][B,H]  query before each hop (+ final) and its gradient
  size_t att_w, att_ds;               // [hops][B,L]    softmax weights, d score
  size_t att_inp;                     // [hops][B,L,4H] concat [q, m, q-m, q*m]
  size_t att_z1, att_dz1;             // [hops][B,L,80]
  size_t att_z2, att_dz2;             // [hops][B,L,40]
  size_t repre, drepre;               // [B,R]
  size_t dlast;                       // [B,D]
  size_t head_bn, head_dbn, head_dgt; // [B,R]  bn output, its gradient, gradient * x_hat (for gamma)
  size_t head_a1, head_act1, head_dl1;// [B,200] pre-activation, post-dropout activation, delta
  size_t head_a2, head_act2, head_dl2;// [B,80]
  size_t head_dlogit;                 // [B]
  size_t pred, logit, w_hop0, scalars;// outputs staging
  size_t tcr;                         // base of the tensor-core recurrence region (aliases proj/st/dxk: one path runs per call)
  size_t total;
};

// workspace of the tensor-core recurrence (tcrec.cu); byte offsets relative to WsLayout::tcr
struct TcrLayout {
  bool ok;
  int DP[HPMN_MAX_LAYERS];                              // padded input width of layer k (multiple of 32)
  size_t wf[HPMN_MAX_LAYERS], bf[HPMN_MAX_LAYERS];      // forward weights [6H][DP+H] (hi rows, lo rows), scaled bias [3H]
  size_t wb[HPMN_MAX_LAYERS];                           // backward weights [6H][H]: Wc_h^T | Wu_h^T | Wr_h^T, hi rows then lo rows
  size_t xh[HPMN_MAX_LAYERS], xl[HPMN_MAX_LAYERS];      // [B,S_k,DP] hi / lo halves of the layer input
  size_t st[HPMN_MAX_LAYERS];                           // [B,S_k,4H] h | r | u | c
  size_t da[HPMN_MAX_LAYERS];                           // [B,S_k,3H]
  size_t dx[HPMN_MAX_LAYERS];                           // [B,S_k,DP]
  size_t hr[HPMN_MAX_LAYERS];                           // [B,S_k,2H] h_prev | r*h_prev (operands of the FFMA weight-gradient fallback)
  size_t wxt[HPMN_MAX_LAYERS];                          // [3H][Din_k] = [Wg_x | Wc_x]^T (dX = dA * WxT), fp32
  size_t total;
};
struct Dims;
bool tcrec_supported(const Dims&);
TcrLayout make_tcr_layout(const Dims&);

struct PackLayout {                    // float offsets inside the packed-weights block
  int64_t Wx[HPMN_MAX_LAYERS];        // [DinP, 96]  input weights, cols g*32+j
  int64_t bx[HPMN_MAX_LAYERS];        // [96]
  int64_t Wh[HPMN_MAX_LAYERS];        // [3][32 i][32 j] recurrent weights (fwd, lane j coalesced)
  int64_t WhT[HPMN_MAX_LAYERS];       // [3][32 j][32 i] transposed (bwd, lane i coalesced)
  int64_t WxT[HPMN_MAX_LAYERS];       // [96, DinP]
  int64_t total;
};

inline PackLayout make_pack_layout(const Dims& d) {
  PackLayout p; memset(&p, 0, sizeof(p));
  int64_t off = 0;
  for (int k = 0; k < d.L; ++k) {
    p.Wx[k] = off;  off += (int64_t)d.DinP[k] * G3;
    p.bx[k] = off;  off += G3;
    p.Wh[k] = off;  off += 3 * HP * HP;
    p.WhT[k] = off; off += 3 * HP * HP;
    p.WxT[k] = off; off += (int64_t)G3 * d.DinP[k];
  }
  p.total = off;
  return p;
}

inline WsLayout make_ws_layout(const Dims& d) {
  WsLayout w; memset(&w, 0, sizeof(w));
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t f = sizeof(float);
  w.ids = take((size_t)d.B * d.T * d.F * sizeof(int32_t));
  w.labels = take((size_t)d.B * sizeof(int32_t));
  w.x = take((size_t)d.B * d.Tpad * d.D * f);
  w.pw = take((size_t)make_pack_layout(d).total * f);
  off = (off + 1023) & ~(size_t)1023;
  w.tcr = off;
  if (d.H <= HP) {
    for (int k = 0; k < d.L; ++k) {
      size_t rows = (size_t)d.B * d.S[k];
      w.proj[k] = take(rows * G3 * f);
      w.st[k] = take(rows * ST * f);
      w.dxk[k] = take(rows * (k == 0 ? d.D : HP) * f);
    }
  }
  {
    const TcrLayout t = make_tcr_layout(d);
    if (t.ok && w.tcr + t.total > off) off = w.tcr + t.total;
  }
  w.memory = take((size_t)d.B * d.L * d.H * f);
  w.dmemory = take((size_t)d.B * d.L * d.H * f);
  w.att_q = take((size_t)(d.hops + 1) * d.B * d.H * f);
  w.att_dq = take((size_t)(d.hops + 1) * d.B * d.H * f);
  w.att_w = take((size_t)d.hops * d.B * d.L * f);
  w.att_ds = take((size_t)d.hops * d.B * d.L * f);
  w.att_inp = take((size_t)d.hops * d.B * d.L * 4 * d.H * f);
  w.att_z1 = take((size_t)d.hops * d.B * d.L * ATT1 * f);
  w.att_dz1 = take((size_t)d.hops * d.B * d.L * ATT1 * f);
  w.att_z2 = take((size_t)d.hops * d.B * d.L * ATT2 * f);
  w.att_dz2 = take((size_t)d.hops * d.B * d.L * ATT2 * f);
  w.repre = take((size_t)d.B * d.R * f);
  w.drepre = take((size_t)d.B * d.R * f);
  w.dlast = take((size_t)d.B * d.D * f);
  w.head_bn = take((size_t)d.B * d.R * f);
  w.head_dbn = take((size_t)d.B * d.R * f);
  w.head_dgt = take((size_t)d.B * d.R * f);
  w.head_a1 = take((size_t)d.B * FC1 * f);
  w.head_act1 = take((size_t)d.B * FC1 * f);
  w.head_dl1 = take((size_t)d.B * FC1 * f);
  w.head_a2 = take((size_t)d.B * FC2 * f);
  w.head_act2 = take((size_t)d.B * FC2 * f);
  w.head_dl2 = take((size_t)d.B * FC2 * f);
  w.head_dlogit = take((size_t)d.B * f);
  w.pred = take((size_t)d.B * f);
  w.logit = take((size_t)d.B * f);
  w.w_hop0 = take((size_t)d.B * d.L * f);
  w.scalars = take(4 * f);
  w.total = off;
  return w;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
// exp via ex2.approx (<= 2 ulp) -- error ~1e-7 relative, far inside the 1e-4 parity budget.  The .ftz forms are used
// directly: __expf() wraps ex2 in a denormal-range fix-up (FSETP + two predicated FMULs) that sits on the dependent chain
// of every recurrent step, and a flushed denormal changes neither 1/(1+e) nor 1-2/(1+e).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)); }
__device__ __forceinline__ float tanh_f(float x) {
  // |x| >= 0.15: 1 - 2/(1+e^{2x}) (saturates cleanly: e^{2x} -> inf gives 1, -> 0 gives -1).
  // |x| <  0.15: odd Taylor series to x^7 (next term < 1e-9 relative) -- the closed form cancels there and
  // would carry ~1e-7 ABSOLUTE error into values of size 1e-3, i.e. 1e-4 relative.
  const float e = ex2_ftz(x * 2.8853900817779268f);
  const float big = fmaf(-2.0f, rcp_ftz(1.0f + e), 1.0f);
  const float x2 = x * x;
  const float small = x * fmaf(x2, fmaf(x2, fmaf(x2, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  return fabsf(x) < 0.15f ? small : big;
}
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// same with a 64-byte L2 fill: a 64-byte embedding row that misses must not drag its 128-byte line partner in from DRAM
// (ncu r1: 83 MB read for 36 MB of rows)
__device__ __forceinline__ float4 ldg_nc_f4_l2_64(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// ---- bulk async copies (TMA engine, SASS UBLKCP) and mbarriers -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  int spins = 0;
  do {
    // the suspend-time hint keeps a blocked warp parked in hardware (no spurious wake-ups stealing issue slots)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    if (!done && ++spins > (1 << 16)) __trap();      // a lost transaction must abort, never hang the GPU
  } while (!done);
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its predecessor in the stream is
// still running (as soon as every CTA of the predecessor has executed pdl_trigger() or exited); it must call pdl_wait()
// before it touches anything the predecessor writes.  Used to hide the prologue of the wavefront kernels (weights ->
// registers / shared memory, barrier setup) behind the tail of the GEMM / attention kernel in front of them.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

// ---------------------------------------------------------------------------------------------
// kernel launchers (defined in the .cu files; all asynchronous on `st`)
// ---------------------------------------------------------------------------------------------
struct Launch {   // launch bookkeeping shared with the ctx
  int64_t* counter;
  int sms;
};

#ifdef __CUDACC__
// <<<grid, block, smem, st>>> with the programmatic-stream-serialization attribute (HPMN_NO_PDL=1: plain launch)
template <class... Params, class... Args>
inline cudaError_t launch_pdl(void (*kern)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool off = [] { const char* e = getenv("HPMN_NO_PDL"); return e && e[0] == '1'; }();
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Params>(args)...);
}
#endif

void launch_gather_fwd(const Launch&, const Dims&, bool mask_id0, int front_pad, int64_t V, const int32_t* ids,
                       const float* table, float* x, float* iderr, cudaStream_t st);
void launch_gather_bwd(const Launch&, const Dims&, bool mask_id0, int front_pad, int last_offset, int64_t V,
                       const int32_t* ids, const float* dx, const float* dlast, float* dtable, cudaStream_t st);

// the same scatter-add over nsrc sources in one launch (local buffers, or peer mappings: peer-row gradient exchange)
void launch_gather_bwd_multi(const Launch&, const Dims&, bool mask_id0, int front_pad, int last_offset, int64_t V, int nsrc,
                             const int32_t* const* ids, const float* const* dx, const float* const* dlast, float* dtable,
                             cudaStream_t st);

void launch_pack(const Launch&, const Dims&, const ParamLayout&, const PackLayout&, const float* params, float* pw,
                 cudaStream_t st);
// C[M,N] = A[M,K](row stride lda) * W[K,N] (+ bias[N]); N, K multiples of 4
void launch_gemm_nn(const Launch&, const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t M,
                    int N, int K, cudaStream_t st);
// C[M,N](ldc) (+)= A[M, a0:a0+K](lda) * W[N,K](ldw)^T, fp32 FFMA (fallback where no tcgen05 instantiation exists)
void launch_gemm_nt(const Launch&, const float* A, int64_t lda, int a0, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                    int N, int K, bool accumulate, cudaStream_t st);
// same contract on tcgen05 tensor cores (3xTF32); returns false when (K,N) has no instantiation
bool launch_tc_gemm_nn(const Launch&, const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t M,
                       int N, int K, cudaStream_t st);
// GRU weight gradients on tcgen05 (same contract as launch_gru_wgrad); false when the shape has no instantiation
bool launch_tc_wgrad(const Launch&, const Dims&, int k, const float* xin, int64_t ldx, const float* st, const float* da,
                     float* dWg, float* dbg, float* dWc, float* dbc, cudaStream_t st_);
// every layer in one launch per padded input width (per-layer arrays of length d.L); false if a layer has no instantiation
bool launch_tc_wgrad_all(const Launch&, const Dims&, const float* const* xin, const int64_t* ldx, const float* const* st,
                         const float* const* da, float* const* dWg, float* const* dbg, float* const* dWc, float* const* dbc,
                         cudaStream_t st_);
// H = 64 with the row layouts of tcrec.cu (state rows 4 x 64, dA rows 3 x 64): the same kernel over 32 x 32 slices
bool launch_tc_wgrad_wide(const Launch&, const Dims&, const float* const* xin, const int64_t* ldx, const float* const* st,
                          const float* const* da, float* const* dWg, float* const* dbg, float* const* dWc, float* const* dbc,
                          cudaStream_t st_);
// batched C[I,N](ldc) += sum_m A[m,I](lda) * Bm[m,N](ldb) with atomic accumulation; A == nullptr -> ones
struct AtbProb {
  const float* A; const float* Bm; float* C;
  int64_t lda, ldb, ldc, M, rows_per_split;
  int I, N, block_begin, pad_;
};
constexpr int ATB_MAX = 40;
struct AtbBatch { int n; int blocks; AtbProb p[ATB_MAX]; };
void atb_add(AtbBatch& batch, int sms, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
             int64_t M, int I, int N);
void launch_atb_batch(const Launch&, const AtbBatch& batch, cudaStream_t st);

// tensor-core recurrence (tcrec.cu).  ws = base of the TcrLayout region.
void launch_tcr_pack(const Launch&, const Dims&, const ParamLayout&, const TcrLayout&, const float* params, char* ws, cudaStream_t st);
void launch_tcr_split(const Launch&, const float* x, float* xh, float* xl, int64_t rows, int D, int DP, cudaStream_t st);
long long* tcr_debug_buffer();
bool launch_tcrec_bwd(const Launch&, const Dims&, const TcrLayout&, int k, char* ws, const float* dmemory, const float* dx_up,
                      bool write_hr, cudaStream_t st);
bool launch_tcrec_fwd(const Launch&, const Dims&, const TcrLayout&, int k, char* ws, float* memory, cudaStream_t st);

void launch_rec_fwd(const Launch&, const Dims&, int k, const float* proj, const float* Wh, float* st, float* memory,
                    cudaStream_t st_);
void launch_rec_bwd(const Launch&, const Dims&, int k, const float* st, const float* WhT, const float* dmemory,
                    const float* dx_up, float* da, cudaStream_t st_);
void launch_gru_wgrad(const Launch&, const Dims&, int k, const float* xin, int64_t ldx, const float* st, const float* da,
                      float* dWg, float* dbg, float* dWc, float* dbc, cudaStream_t st_);

// all layers of the memory as one wavefront kernel (wave.cu); proj0 = layer-0 input projections; false if L is too large
bool launch_wave_fwd(const Launch&, const Dims&, const PackLayout&, const float* proj0, const float* pw, float* const* st,
                     float* memory, cudaStream_t st_);

// backward twin: consumes st[k], dmemory; emits da[k] for every layer (dx of layers >= 1 is handed down in-kernel)
bool launch_wave_bwd(const Launch&, const Dims&, const PackLayout&, const float* pw, const float* const* st, float* const* da,
                     const float* dmemory, cudaStream_t st_);

// resolved workspace pointers handed to the attention / head kernels
struct AttWs { float *q, *dq, *w, *ds, *inp, *z1, *dz1, *z2, *dz2; };
struct HeadWs { float *bn, *dbn, *dgt, *a1, *act1, *dl1, *a2, *act2, *dl2, *dlogit; };

void launch_attn_fwd(const Launch&, const Dims&, const ParamLayout&, int last_offset, const float* memory, const float* x,
                     const float* params, float* repre, float* w_hop0, float* scalars, const AttWs& ws, cudaStream_t st);
// per-sample deltas only; weight gradients are queued on `batch` (launch_atb_batch)
void launch_attn_bwd(const Launch&, const Dims&, const ParamLayout&, int last_offset, float memory_reg, const float* memory,
                     const float* x, const float* params, const float* drepre, float* dmemory, float* dlast, float* grads,
                     const AttWs& ws, AtbBatch& batch, cudaStream_t st);

// row0: first batch row of this row group (keeps the dropout hash independent of how the batch is grouped)
void launch_head_fwd(const Launch&, const Dims&, const ParamLayout&, const hpmn_hyper&, int row0, const float* repre,
                     const int32_t* labels, const float* params, float* pred, float* logit, float* scalars,
                     const HeadWs& ws, cudaStream_t st);
void launch_head_bwd(const Launch&, const Dims&, const ParamLayout&, const hpmn_hyper&, int row0, const float* repre,
                     const int32_t* labels, const float* params, const float* pred, float* drepre, float* grads,
                     const HeadWs& ws, AtbBatch& batch, cudaStream_t st);
// training step: attention fwd + head fwd + head bwd + attention bwd of one sample per CTA in ONE kernel (mid.cu)
void launch_mid_fused(const Launch&, const Dims&, const ParamLayout&, const hpmn_hyper&, int last_offset, int row0,
                      const float* memory, const float* x, const float* params, const int32_t* labels, float* repre,
                      float* w_hop0, float* pred, float* logit, float* scalars, float* drepre, float* dmemory, float* dlast,
                      float* grads, const AttWs& aws, const HeadWs& hws, AtbBatch& batch, cudaStream_t st);

// in-switch all-reduce of a symmetric buffer through its multicast mapping (comm.cu); ctas <= 0: one CTA per SM
void launch_nvls_allreduce(const Launch&, float* mc, int64_t n_floats, int rank, int world, int ctas, cudaStream_t st);

void launch_clip_adam(const Launch&, float* var, const float* grad, float* m, float* v, int64_t n, float lr_t, float b1,
                      float b2, float eps, float clip, cudaStream_t st);
void launch_axpy(const Launch&, float* y, const float* x, float a, int64_t n, cudaStream_t st);
// loss = logloss + memory_reg * covreg, and (optionally) the per-row results copied from the workspace staging to the caller's
// device buffers by the same launch (four device-to-device copies cost ~16 us of stream time at the end of every step)
void launch_zero(const Launch&, void* ptr, size_t bytes, int ctas, cudaStream_t st);
struct OutCopies { const float* src[4]; float* dst[4]; int64_t n[4]; };
void launch_finish_scalars(const Launch&, float* scalars, float memory_reg, const OutCopies* copies, cudaStream_t st);

// dropout keep-mask shared by head fwd/bwd (counter-based hash; deterministic in seed, sample, unit)
#ifdef __CUDACC__
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint32_t layer, uint32_t b, uint32_t unit, float keep_prob) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((uint64_t)layer * 0x100000000ull + ((uint64_t)b << 10) + unit + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);   // 24 bits -> [0,1)
  return u < keep_prob;
}
#endif

}  // namespace hpmn
