// wave.cu -- the hierarchical periodic memory as ONE persistent kernel per direction (forward here, backward
// below): all L layers of a sample run concurrently as a wavefront, layer k firing every prod(p[:k]) steps, so the
// critical path is the S_0 = Tpad steps of layer 0 instead of sum_k S_k (1024 vs 1984 at XLong).
// Replaces the per-layer tf.nn.dynamic_rnn calls of /root/reference/code/hpmn.py:113-131 (cell code/util.py:81-110);
// the equivalence of the layer-by-layer and the online formulation is the reference's own
// srnn.py:725-748 (`incremental_update`) and is pinned by tests/test_oracle.py.
//
// One CTA = NSPC samples x (L layer warps + 1 helper).  Warp (k, s) owns the recurrence of layer k of sample s with its
// recurrent weights in registers; state is exchanged between the lanes through shared memory, h|r|u|c rows leave through
// TMA bulk stores.  Layer 0 streams its input projections (tcgen05 GEMM output) in with cp.async.bulk + mbarrier; a layer
// k >= 1 receives every p-th hidden state of layer k-1 through a small shared-memory ring with full/empty mbarriers.
//
// Lane layout of every 32x32 mat-vec ("K-half"): lane l accumulates, over ITS HALF of the 32 inputs (k = 16 (l&1) ..
// +15), the partial dot products of output l and of output l^1; one SHFL.BFLY with the neighbour completes both.  The
// first version gave lane l the whole of output l, i.e. every lane read all 32 inputs (8 LDS.128 per mat-vec, 4 KB written
// back to the register file).  ncu showed that this write-back path (128 B/clk per SM, shared by the 12 warps of the CTA)
// was what the kernels were bound by once all layers were resident: 52 % (forward) and 77 % (backward) busy, layer 0's
// straight-line step 110 cycles slower with the upper layers running than alone (profiles/r1_v8_wave_step_timeline.md).
// K-half moves half the bytes (4 LDS.128) for the same 48 FFMA2 per step and one extra shuffle per mat-vec.
//
// Supported period patterns: L == 1, or p_0 == 2 (and p_1 == 2 when L > 2) -- every configuration of the reference
// (hpmn.py:576-662); other patterns return false and the caller runs the per-layer kernels (gru.cu).
#include <stdlib.h>

#include "common.cuh"

namespace hpmn {

constexpr int WCH = 8;        // steps per output chunk of layers >= 1 (one 4 KB bulk store)
constexpr int WIN = 16;       // steps per input / output chunk of layer 0
constexpr int WNS0 = 3;       // input ring stages of layer 0
constexpr int HRS = 8;        // hand-off ring slots between consecutive layers
constexpr int HG = HRS / 2;   // rows per hand-off group (group-granular protocol on the layer-0 / layer-1 side)
constexpr int BLK = 8;        // steps per block of the latency-critical loops = 2 * HG hand-offs at period 2
// Steps unrolled inside a block (even, divides BLK).  Fully unrolled blocks (8 steps = 17 KB of SASS per role) made the
// forward kernel's layer-0 step time swing between 405 and 625 cycles from build to build: with layer 1 running the same
// loop the hot code outgrew the 32 KB L1.5 instruction cache.  Two steps = 4.3 KB stay inside the 6 KB L0 I-cache.
#ifndef HPMN_FUNR
#define HPMN_FUNR 2
#endif
#ifndef HPMN_BUNR
#define HPMN_BUNR 2
#endif
#ifndef HPMN_GUNR
#define HPMN_GUNR 1
#endif
#ifndef HPMN_CSMEM_FWD
#define HPMN_CSMEM_FWD 1
#endif
#ifndef HPMN_CSMEM_BWD
#define HPMN_CSMEM_BWD 0
#endif
// layers >= 2 read the candidate gate's recurrent weights from shared memory instead of holding them in registers: 32 more
// registers for batching their other shared-memory weight loads.  Measured (XLong, B=256): forward 0.294 -> 0.292 ms; backward
// 0.266 -> 0.290 ms (layer 2's step gets 30 % shorter, but the extra LSU traffic slows the co-critical layers 0 and 1) => off.
constexpr bool CS_FWD = HPMN_CSMEM_FWD != 0, CS_BWD = HPMN_CSMEM_BWD != 0;
// Experimental: backward kernel with ONE warp per sample for all layers >= 2 (8 warps per CTA -> 255 registers per thread
// instead of 168), see wave_bwd_upper.  Parity-green (GPU test suite), but slower so far: rec_bwd 0.365 ms against 0.265 ms for
// one warp per layer (XLong, B=256).  Its arithmetic is fast (820 cycles per step against ~1400), the warp loses ~900 cycles
// per step outside it -- a first cut with two step variants in the loop (13 KB of SASS) ran at 0.408 ms, so instruction
// fetch is the suspect.  Off by default; -DHPMN_BWD_UPPER=1 builds it, HPMN_NO_BWD_UPPER=1 switches it off at run time.
#ifndef HPMN_BWD_UPPER
#define HPMN_BWD_UPPER 0
#endif
constexpr bool BWD_UPPER = HPMN_BWD_UPPER != 0;
constexpr int FUNR = HPMN_FUNR, BUNR = HPMN_BUNR, GUNR = HPMN_GUNR;   // GUNR: the loops of the layers off the critical path
constexpr int WAVE_MAX_L = 10;

static inline bool wave_supported(int L, const int* P) { return L == 1 || (P[0] == 2 && (L == 2 || P[1] == 2)); }

struct Handoff {                   // layer k -> k+1 (fwd) / k+1 -> k (bwd) of one sample: ring of HRS rows
  float ring[HRS][HP];
  uint64_t full[HRS], empty[HRS];  // per-slot protocol (layers off the critical path)
  uint64_t gfull[2], gempty[2];    // group protocol: one barrier round trip per HG rows
};
struct Handoff3 {                  // da rows of layer 1 on their way to the dx helper (bwd): 96 floats per slot, groups of HG rows
  float ring[HRS][G3];
  uint64_t gfull[2], gempty[2];
};

__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& acc, bool timed) {
  if (!timed) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ring-slot offset that aligns a producer's groups of HG rows with the consumer's BLK-step chunks when its top chunk is ragged
__device__ __forceinline__ unsigned group_vofs(int S_consumer) { return (unsigned)((HG - (S_consumer % BLK) / 2) & (HG - 1)); }

// ---- K-half mat-vec building blocks ------------------------------------------------------------------------------
struct HalfW { float2 m[8], o[8]; };   // (W[k][l], W[k+1][l]) and (W[k][l^1], W[k+1][l^1]) for k = 16 (l&1) + 2q
struct Part { float m, o; };           // partial sums of output l (mine) and l^1 (the partner's) over this lane's 16 inputs

// M: [32 in][32 out] row-major in global memory
__device__ __forceinline__ void load_half(HalfW& w, const float* M, int lane, float scale) {
  const int k0 = 16 * (lane & 1);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float* r0 = M + (k0 + 2 * q) * HP;
    w.m[q] = make_float2(scale * __ldg(r0 + lane), scale * __ldg(r0 + HP + lane));
    w.o[q] = make_float2(scale * __ldg(r0 + (lane ^ 1)), scale * __ldg(r0 + HP + (lane ^ 1)));
  }
}
// this lane's 16 inputs of a 32-vector in shared memory: 4 LDS.128, two addresses per warp
__device__ __forceinline__ void load_vec_half(float4 (&v)[4], const float* sh, int lane) {
  const float4* p = reinterpret_cast<const float4*>(sh + 16 * (lane & 1));
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = p[q];
}
__device__ __forceinline__ Part half_part(const float4 (&v)[4], const HalfW& w, float init) {
  float2 m0 = make_float2(init, 0.f), m1 = make_float2(0.f, 0.f), o0 = m1, o1 = m1;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 xa = make_float2(v[q].x, v[q].y), xb = make_float2(v[q].z, v[q].w);
    m0 = ffma2(xa, w.m[2 * q], m0); o0 = ffma2(xa, w.o[2 * q], o0);
    m1 = ffma2(xb, w.m[2 * q + 1], m1); o1 = ffma2(xb, w.o[2 * q + 1], o1);
  }
  Part p; p.m = (m0.x + m1.x) + (m0.y + m1.y); p.o = (o0.x + o1.x) + (o0.y + o1.y);
  return p;
}
// same with the weights in shared memory, one float4 per K-pair: (mine.x, mine.y, other.x, other.y) at wq[pair * pstride]
// (wq already includes the lane).  One LDS.128 feeds two FFMA2: with ~30 free registers next to the 96 of the recurrent
// weights, ptxas otherwise serialises load -> use pairs (r1 v9: 24 of them back to back in the backward step of layer 2).
__device__ __forceinline__ void half_part_smem(const float4 (&v)[4], const float4* wq, int pstride, float2& m, float2& o) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 wa = wq[(2 * q) * pstride], wb = wq[(2 * q + 1) * pstride];
    const float2 xa = make_float2(v[q].x, v[q].y), xb = make_float2(v[q].z, v[q].w);
    m = ffma2(xa, make_float2(wa.x, wa.y), m); o = ffma2(xa, make_float2(wa.z, wa.w), o);
    m = ffma2(xb, make_float2(wb.x, wb.y), m); o = ffma2(xb, make_float2(wb.z, wb.w), o);
  }
}
__device__ __forceinline__ float half_finish(float mine, float other) { return mine + __shfl_xor_sync(0xffffffffu, other, 1); }

// Warp -> role assignment.  Warps of a CTA land on SM sub-partition (warp id % 4) and the issue arbiter favours higher
// warp ids.  Roles are spread so that the estimated load per sub-partition is balanced and the critical layer-0 warps
// share theirs only with the lightest roles; inside a sub-partition the heaviest role gets the highest warp id.
struct WarpPlan { int n; signed char layer[16], sample[16], helper[16]; };
static WarpPlan plan_warps(int L, int nspc, bool with_helper, const char* env_weights = nullptr) {
  struct Role { int layer, sample, helper; double w; };
  Role roles[16]; int nr = 0;
  // experiment hook: HPMN_WAVE_WF / HPMN_WAVE_WB = "w0,w1,w2,w3,w4,whelper" override the per-layer load estimates
  double ow[6] = {-1, -1, -1, -1, -1, -1};
  if (const char* e = env_weights ? getenv(env_weights) : nullptr) sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf", ow, ow + 1, ow + 2, ow + 3, ow + 4, ow + 5);
  for (int s = 0; s < nspc; ++s)
    for (int k = 0; k < L; ++k) {
      double w = k == 0 ? 2.0 : 1.0;                     // layer 0 is the critical path: keep its sub-partition quiet
      for (int q = 0; q < k; ++q) w *= 0.5;              // layer k runs ~2^-k of layer 0's steps ...
      if (k >= 1) w *= (with_helper && k == 1) ? 1.1 : 1.9;   // ... each ~2x as expensive unless a helper takes the matvec
      if (k < 5 && ow[k] >= 0) w = ow[k];
      roles[nr++] = Role{k, s, 0, w};
    }
  if (with_helper && L > 1)
    for (int s = 0; s < nspc; ++s) roles[nr++] = Role{1, s, 1, ow[5] >= 0 ? ow[5] : 0.45};
  for (int a = 0; a < nr; ++a)                            // heaviest first
    for (int b = a + 1; b < nr; ++b)
      if (roles[b].w > roles[a].w) { Role t = roles[a]; roles[a] = roles[b]; roles[b] = t; }
  const int cap = (nr + 3) / 4;
  double load[4] = {0, 0, 0, 0}; int cnt[4] = {0, 0, 0, 0}; int slot[4][4];
  for (int a = 0; a < nr; ++a) {
    int best = -1;
    for (int q = 0; q < 4; ++q)
      if (cnt[q] < cap && (best < 0 || load[q] < load[best])) best = q;
    slot[best][cnt[best]++] = a; load[best] += roles[a].w;
  }
  // sub-partition q owns warp ids q, q+4, q+8, ...: heaviest role -> highest id.  Every id below the CTA size must be used,
  // so sub-partitions are filled from id q upward with their lightest role first.
  WarpPlan p; p.n = nr;
  for (int i = 0; i < 16; ++i) { p.layer[i] = -1; p.sample[i] = 0; p.helper[i] = 0; }
  // ids available to sub-partition q: q, q+4, ... < nr.  If a sub-partition got more roles than it has ids (uneven nr),
  // spill its lightest roles to any free id.
  bool used[16] = {false};
  int spill[16], nsp = 0;
  for (int q = 0; q < 4; ++q) {
    int ids[4], ni = 0;
    for (int id = q; id < nr; id += 4) ids[ni++] = id;
    for (int c = 0; c < cnt[q]; ++c) {                   // c = 0 is the heaviest
      const int a = slot[q][c];
      if (c < ni) { const int id = ids[ni - 1 - c]; p.layer[id] = roles[a].layer; p.sample[id] = roles[a].sample; p.helper[id] = roles[a].helper; used[id] = true; }
      else spill[nsp++] = a;
    }
  }
  for (int i = 0, id = 0; i < nsp; ++i) {
    while (used[id]) ++id;
    p.layer[id] = roles[spill[i]].layer; p.sample[id] = roles[spill[i]].sample; p.helper[id] = roles[spill[i]].helper; used[id] = true;
  }
  return p;
}

// =====================================================================================================
// Forward
// =====================================================================================================
struct WaveArgs {
  const float* proj0;              // [B,S_0,96]
  const float* pw;                 // packed weights
  float* st[HPMN_MAX_LAYERS];      // [B,S_k,128] per layer
  float* memory;                   // [B,L,H]
  int64_t Wh[HPMN_MAX_LAYERS], Wx[HPMN_MAX_LAYERS], bx[HPMN_MAX_LAYERS];   // offsets into pw
  int S[HPMN_MAX_LAYERS], P[HPMN_MAX_LAYERS];
  int B, L, H, nspc;
  int debug;                       // HPMN_WAVE_DEBUG=1: CTA 0 prints per-role cycle counts
  signed char wlayer[16], wsample[16], whelper[16];   // warp id -> (layer, sample slot, is helper); see plan_warps()
};

struct InRing {                    // helper -> layer 1: two chunks of BLK projected input rows (r|u|c, pre-scaled)
  float row[2][BLK][G3];
  uint64_t full[2], empty[2];
};

// shared-memory plan (bytes)
struct WaveSmem {
  int wx, whc, bx, hand, hand3, l0, lk, total;
  int r0, r1;                      // per-warp region sizes: layer 0 / layers >= 1
  __host__ __device__ WaveSmem(int L, int nspc) {
    int off = 0;
    wx = off; off += (L - 1) * 8 * 3 * HP * 16;            // float4 [q][g][lane] = (mine pair | other pair) per layer >= 1
    whc = off; off += (L > 2 ? L - 2 : 0) * 8 * HP * 16;   // candidate-gate recurrent weights of layers >= 2, float4 [q][lane]
    bx = off; off += (L - 1) * G3 * 4;
    off = (off + 127) & ~127;
    hand = off; off += (L - 1) * nspc * (int)sizeof(Handoff);
    off = (off + 127) & ~127;
    hand3 = off; off += (L > 1 ? nspc : 0) * (int)sizeof(InRing);     // helper warp of layer 1 -> layer 1
    off = (off + 127) & ~127;
    r0 = WNS0 * WIN * G3 * 4 + 2 * WIN * ST * 4 + 256 + 128;   // in ring | out ring (16-step chunks) | sh_rh + sh_h | mbarriers
    r1 = 2 * WCH * ST * 4 + 256 + 128;                         // out ring | sh_rh + sh_h
    l0 = off; off += nspc * r0;
    lk = off; off += (L - 1) * nspc * r1;
    total = off;
  }
};

// sigmoid(x) = 1 / (1 + 2^(-x log2 e)), tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e)): the log2(e) factors are folded into the
// register-resident recurrent weights and into the projected inputs, so a mat-vec result feeds MUFU.EX2 directly.
constexpr float kNegLog2e = -1.4426950408889634f;
constexpr float kTwoLog2e = 2.8853900817779268f;
constexpr float kInvTwoLog2e = 0.34657359027997264f;

struct FwdW { HalfW r, u, c; };
__device__ __forceinline__ void load_fwd_weights(FwdW& w, const float* Wh /*[3][32 i][32 j]*/, int lane) {
  load_half(w.r, Wh, lane, kNegLog2e);
  load_half(w.u, Wh + HP * HP, lane, kNegLog2e);
  load_half(w.c, Wh + 2 * HP * HP, lane, kTwoLog2e);
}

// One GRU step (util.py:81-110 minus :108) of the (layer, sample) this warp owns.  ar|au|ac: this lane's projected inputs,
// already multiplied by the log2(e) factors; sh_h holds h_{t-1} (all lanes), out is this lane's column of the state row.
// The caller issues the __syncwarp() that closes the step (after its own stores).
// C_SMEM (layers >= 2): the candidate gate's recurrent weights come from shared memory (wc_s: float4 [8][lane], pre-scaled)
// instead of registers -- 32 registers more for batching these layers' shared-memory weight loads.
template <bool C_SMEM>
__device__ __forceinline__ float gru_fwd_step(const FwdW& w, float ar, float au, float ac, float h, int j, float* sh_h,
                                              float* sh_rh, float* out, const float4* wc_s) {
  float4 v[4];
  load_vec_half(v, sh_h, j);
  const Part pr = half_part(v, w.r, ar), pu = half_part(v, w.u, au);
  const float sr = half_finish(pr.m, pr.o), su = half_finish(pu.m, pu.o);
  const float r = rcp_ftz(1.0f + ex2_ftz(sr));           // util.py:95-96
  sh_rh[j] = r * h;                                      // util.py:98
  __syncwarp();
  load_vec_half(v, sh_rh, j);
  const float u = rcp_ftz(1.0f + ex2_ftz(su));
  Part pc;
  if (C_SMEM) {
    float2 m = make_float2(ac, 0.f), o = make_float2(0.f, 0.f);
    half_part_smem(v, wc_s + j, HP, m, o);
    pc.m = m.x + m.y; pc.o = o.x + o.y;
  } else {
    pc = half_part(v, w.c, ac);
  }
  const float omu = 1.0f - u, uh = u * h;
  const float x2l = half_finish(pc.m, pc.o);             // 2 log2(e) x
  const float qv = rcp_ftz(1.0f + ex2_ftz(x2l));         // tanh(x) = 1 - 2 qv                  util.py:107
  // |x| < 0.15: odd Taylor series to x^7 (the closed form cancels there: ~1e-7 absolute on values of 1e-3); off the chain
  const float x = x2l * kInvTwoLog2e, xx = x * x;
  const float small = x * fmaf(xx, fmaf(xx, fmaf(xx, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
  const bool tiny = fabsf(x) < 0.15f;
  const float hs = fmaf(omu, small, uh);
  const float hbig = fmaf(-2.0f * omu, qv, uh + omu);    // u h + (1-u)(1 - 2 qv): one FFMA after the reciprocal   util.py:109
  const float hn = tiny ? hs : hbig;
  const float c = tiny ? small : fmaf(-2.0f, qv, 1.0f);
  sh_h[j] = hn;
  out[0] = hn; out[HP] = r; out[2 * HP] = u; out[3 * HP] = c;
  return hn;
}

// ---------------------------------------------------------------------------------------------------------------------
// Layer 0: the critical path of the kernel (S_0 dependent steps), written for the latency of ONE step.  Per 8-step block
// there is no branch and no address arithmetic (the broadcast buffers sit at fixed addresses, input / output rows are
// immediates off two per-block pointers) and the hand-offs to the helper warp happen at compile-time steps (period 2 =>
// odd t, slots 4g..4g+3 of group g = block & 1, one barrier round trip per block).  FIRE2 = false: no layer above.
// ---------------------------------------------------------------------------------------------------------------------
// TMA_IN (layer 0): projected inputs stream in from global memory, chunks of 16 steps.  Otherwise (layer 1): the helper
// warp delivers chunks of 8 pre-scaled projected rows through `inr` -- layer 1 runs the same loop as layer 0 at half its
// rate, so it never holds layer 0 back.
// Layer 0's chunk I/O as seen by the helper warp that serves it (wave_proj_helper): ncu / the cycle counters put the lane-0
// bulk store + TMA refill + bulk_wait_read at 29 k of layer 0's 568 k cycles, and the helper has the slack.
struct L0Io { float* s_in; float* s_out; uint64_t* full; const float* pp; float* so; int S; };

template <bool TMA_IN, bool FIRE2>
__device__ __forceinline__ float wave_layer_fast(const WaveArgs& a, int k, int b, int j, float* s_in, uint64_t* full,
                                                 InRing* inr, float* s_out, float* sh_rh, float* sh_h, Handoff* hout) {
  constexpr int CH = TMA_IN ? WIN : BLK;
  // SELF_IO = false (layer 0 with a helper): the helper warp issues this warp's bulk stores and TMA refills.  It learns that
  // chunk c is complete from the hand-off barrier of the chunk's last block (this warp fences its rows for the async proxy
  // before that arrive), stores chunk c and requests input chunk c+2; it makes sure the store has read its source before it
  // releases the NEXT hand-off group, which this warp waits for anyway before it rewrites output buffer c & 1 in chunk c+2.
  constexpr bool SELF_IO = !(TMA_IN && FIRE2);
  const int S = a.S[k];
  FwdW w;
  load_fwd_weights(w, a.pw + a.Wh[k], j);
  const float* pp = a.proj0 + (int64_t)b * S * G3;       // TMA_IN only
  float* so = a.st[k] + (int64_t)b * S * ST;
  const int nch = (S + CH - 1) / CH;
  if (TMA_IN) {
    if (j == 0) {
      for (int i = 0; i < WNS0; ++i) mbar_init(&full[i], 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (j == 0)
      for (int c = 0; c < (SELF_IO ? WNS0 : 2) && c < nch; ++c) {   // served by the helper: chunk c+2 is requested when chunk c is done
        const int len = min(CH, S - c * CH);
        mbar_expect_tx(&full[c], (uint32_t)len * G3 * 4);
        bulk_g2s(s_in + c * CH * G3, pp + (int64_t)c * CH * G3, (uint32_t)len * G3 * 4, &full[c]);
      }
    __syncwarp();
  }

  long long w_out = 0, w_tma = 0, w_steps = 0;           // debug: cycles blocked on the helper / on TMA, cycles inside the step blocks
  long long w_q[4] = {0, 0, 0, 0};                       // debug: bulk_wait_read | gfull arrive | proxy fence | lane-0 bulk store + TMA refill
  const bool dbg = a.debug != 0 && blockIdx.x == 0;
  const long long t_start = clock64();
  float h = 0.f;                                         // zero_state, code/rnn.py:588 (sh_h starts as the zero row)

  auto step = [&](const float* in, float* out) {
    if (TMA_IN) h = gru_fwd_step<false>(w, in[0] * kNegLog2e, in[HP] * kNegLog2e, in[2 * HP] * kTwoLog2e, h, j, sh_h, sh_rh, out, nullptr);
    else h = gru_fwd_step<false>(w, in[0], in[HP], in[2 * HP], h, j, sh_h, sh_rh, out, nullptr);
  };

  for (int c = 0; c < nch; ++c) {
    const int len = min(CH, S - c * CH);
    const int stage = TMA_IN ? c % WNS0 : c & 1;
    float* ob = s_out + (c & 1) * CH * ST;
    const float* ib;
    if (TMA_IN) {
      mbar_wait_t(&full[stage], (uint32_t)(c / WNS0) & 1u, w_tma, dbg);
      ib = s_in + stage * CH * G3;
    } else {
      mbar_wait_t(&inr->full[stage], (uint32_t)(c >> 1) & 1u, w_tma, dbg);
      ib = &inr->row[stage][0][0];
    }
    const long long tq0 = dbg ? clock64() : 0;
    if (SELF_IO && c >= 2) {                             // the bulk store that read this buffer two chunks ago
      if (j == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    if (dbg) w_q[0] += clock64() - tq0;
    const int nb = len / BLK;
    for (int half = 0; half < nb; ++half) {
      const int blk = c * (CH / BLK) + half, g = blk & 1;
      const float* in = ib + half * BLK * G3 + j;
      float* out = ob + half * BLK * ST + j;
      float* rg = nullptr;
      if (FIRE2) {
        rg = &hout->ring[g * HG][j];
        if (blk >= 2) mbar_wait_t(&hout->gempty[g], (uint32_t)(blk / 2 - 1) & 1u, w_out, dbg);
      }
      const long long tb0 = dbg ? clock64() : 0;
#pragma unroll 1
      for (int u = 0; u < BLK / FUNR; ++u) {             // the unrolled body must stay inside the instruction caches (see FUNR)
#pragma unroll
        for (int t = 0; t < FUNR; ++t) {
          step(in + t * G3, out + t * ST);
          if (FIRE2 && (t & 1)) rg[(t >> 1) * HP] = h;   // every 2nd state feeds the layer above (hpmn.py:124-128)
          __syncwarp();                                  // sh_h (next step's broadcast) and sh_rh settled
        }
        in += FUNR * G3; out += FUNR * ST;
        if (FIRE2) rg += (FUNR / 2) * HP;
      }
      if (dbg) w_steps += clock64() - tb0;
      const long long tq1 = dbg ? clock64() : 0;
      if (!SELF_IO && (half + 1) * BLK == len) { fence_proxy_async(); __syncwarp(); }   // last block of the chunk: rows -> async proxy
      if (FIRE2 && j == 0) mbar_arrive(&hout->gfull[g]);
      if (dbg) { __syncwarp(); w_q[1] += clock64() - tq1; }
    }
    const int t0 = nb * BLK;
    if (t0 < len) {                                      // ragged tail (< 8 steps, last chunk only): same step, runtime offsets
      const int blk = c * (CH / BLK) + nb, g = blk & 1;
      if (FIRE2 && blk >= 2 && len - t0 >= 2) mbar_wait_t(&hout->gempty[g], (uint32_t)(blk / 2 - 1) & 1u, w_out, dbg);
      for (int t = t0; t < len; ++t) {
        step(ib + t * G3 + j, ob + t * ST + j);
        if (FIRE2 && ((t - t0) & 1)) hout->ring[g * HG + ((t - t0) >> 1)][j] = h;
        __syncwarp();
      }
      if (!SELF_IO) { fence_proxy_async(); __syncwarp(); }
      if (FIRE2 && len - t0 >= 2 && j == 0) mbar_arrive(&hout->gfull[g]);
    }
    if (!SELF_IO) continue;                              // the helper stores the chunk and refills the input ring
    const long long tq2 = dbg ? clock64() : 0;
    fence_proxy_async();                                 // generic-proxy writes of ob -> visible to the bulk store
    __syncwarp();
    if (dbg) { w_q[2] += clock64() - tq2; }
    const long long tq3 = dbg ? clock64() : 0;
    if (j == 0) {
      bulk_s2g(so + (int64_t)c * CH * ST, ob, (uint32_t)len * ST * 4);
      bulk_commit();
      if (TMA_IN) {
        const int cn = c + WNS0;                         // refill the input stage every lane has finished reading
        if (cn < nch) {
          const int ln = min(CH, S - cn * CH);
          mbar_expect_tx(&full[stage], (uint32_t)ln * G3 * 4);
          bulk_g2s(s_in + stage * CH * G3, pp + (int64_t)cn * CH * G3, (uint32_t)ln * G3 * 4, &full[stage]);
        }
      } else {
        mbar_arrive(&inr->empty[stage]);                 // chunk consumed: the helper may refill it
      }
    }
    if (dbg) { __syncwarp(); w_q[3] += clock64() - tq3; }
  }
  if (SELF_IO && j == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && j == 0 && b == 0)
    printf("wave_fwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma %lld  in-blocks %lld  "
           "[wait_read %lld  arrive %lld  fence %lld  io %lld]\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, TMA_IN ? 0ll : w_tma, w_out, TMA_IN ? w_tma : 0ll, w_steps, w_q[0], w_q[1],
           w_q[2], w_q[3]);
  return h;
}

// ---------------------------------------------------------------------------------------------------------------------
// Layers >= 2: the layer applies its own input projection (W_x from shared memory, K-half layout, pre-scaled) one step
// ahead of the recurrent chain.  Any firing period towards the layer above.  GROUPED_IN (layer 2): layer 1 hands over in
// groups of HG rows (wave_layer_fast), the layers above slot by slot.
// ---------------------------------------------------------------------------------------------------------------------
template <bool GROUPED_IN>
__device__ __forceinline__ float wave_layer_up(const WaveArgs& a, int k, int b, int j, float* s_out, float* sh_rh, float* sh_h,
                                               const float4* myWx, const float* myBx, const float4* myWhc, Handoff* hin,
                                               Handoff* hout) {
  constexpr int CH = WCH;
  const int S = a.S[k], period = a.P[k];
  FwdW w;                                                // r and u in registers, c from shared memory (myWhc)
  load_half(w.r, a.pw + a.Wh[k], j, kNegLog2e);
  load_half(w.u, a.pw + a.Wh[k] + HP * HP, j, kNegLog2e);
  if (!CS_FWD) load_half(w.c, a.pw + a.Wh[k] + 2 * HP * HP, j, kTwoLog2e);
  float* so = a.st[k] + (int64_t)b * S * ST;
  const int nch = (S + CH - 1) / CH;

  float h = 0.f;                                         // zero_state, code/rnn.py:588
  long long w_in = 0, w_out = 0;                         // debug: cycles blocked on input / output barriers
  const bool dbg = a.debug != 0 && blockIdx.x == 0;
  const long long t_start = clock64();
  unsigned fired = 0, s_glob = 0;                        // hand-offs produced / consumed so far
  int to_fire = period;

  // input projection of hand-off `idx` (hidden state of layer k-1 at its step (idx+1)*p_{k-1} - 1, hpmn.py:124-128).  It
  // does not depend on this layer's state, so it is issued one step ahead and overlaps the bubbles of the recurrent chain.
  auto project = [&](unsigned idx, float& ar, float& au, float& ac) {
    const int slot_in = idx & (HRS - 1), gi = (idx / HG) & 1;
    if (GROUPED_IN) { if ((idx & (HG - 1)) == 0) mbar_wait_t(&hin->gfull[gi], (idx / HRS) & 1u, w_in, dbg); }
    else mbar_wait_t(&hin->full[slot_in], (idx / HRS) & 1u, w_in, dbg);   // hardware-suspended wait, no polling
    float4 v[4];
    load_vec_half(v, hin->ring[slot_in], j);
    float2 m0 = make_float2(myBx[j], 0.f), m1 = make_float2(myBx[HP + j], 0.f), m2 = make_float2(myBx[2 * HP + j], 0.f);
    float2 o0 = make_float2(0.f, 0.f), o1 = o0, o2 = o0;
    half_part_smem(v, myWx + 0 * HP + j, 3 * HP, m0, o0);
    half_part_smem(v, myWx + 1 * HP + j, 3 * HP, m1, o1);
    half_part_smem(v, myWx + 2 * HP + j, 3 * HP, m2, o2);
    ar = half_finish(m0.x + m0.y, o0.x + o0.y);
    au = half_finish(m1.x + m1.y, o1.x + o1.y);
    ac = half_finish(m2.x + m2.y, o2.x + o2.y);
    __syncwarp();
    if (GROUPED_IN) { if (j == 0 && ((idx & (HG - 1)) == HG - 1 || idx + 1 == (unsigned)S)) mbar_arrive(&hin->gempty[gi]); }
    else if (j == 0) mbar_arrive(&hin->empty[slot_in]);  // slot free again
  };
  float nar = 0.f, nau = 0.f, nac = 0.f;                 // projections of the NEXT step
  project(0, nar, nau, nac);

  auto step = [&](float* orow) {
    const float ar = nar, au = nau, ac = nac;
    ++s_glob;
    if (s_glob < (unsigned)S) project(s_glob, nar, nau, nac);
    h = gru_fwd_step<CS_FWD>(w, ar, au, ac, h, j, sh_h, sh_rh, orow + j, myWhc);
    if (hout != nullptr && --to_fire == 0) {             // this step feeds layer k+1
      to_fire = period;
      const int slot_out = fired & (HRS - 1);
      if (fired >= HRS) mbar_wait_t(&hout->empty[slot_out], (fired / HRS - 1) & 1u, w_out, dbg);
      hout->ring[slot_out][j] = h;
      ++fired;
      __syncwarp();
      if (j == 0) mbar_arrive(&hout->full[slot_out]);    // release: the row is visible to the waiting layer
    }
    __syncwarp();                                        // sh_h (next step's broadcast) and sh_rh settled
  };

  for (int c = 0; c < nch; ++c) {
    const int len = min(CH, S - c * CH);
    float* ob = s_out + (c & 1) * CH * ST;
    if (c >= 2) {                                        // the bulk store that read this buffer two chunks ago
      if (j == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    if (len == CH) {
#pragma unroll GUNR
      for (int t = 0; t < CH; ++t) step(ob + t * ST);
    } else {
      for (int t = 0; t < len; ++t) step(ob + t * ST);
    }
    fence_proxy_async();                                 // generic-proxy writes of ob -> visible to the bulk store
    __syncwarp();
    if (j == 0) {
      bulk_s2g(so + (int64_t)c * CH * ST, ob, (uint32_t)len * ST * 4);
      bulk_commit();
    }
  }
  if (j == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && j == 0 && b == 0)
    printf("wave_fwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma 0\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, w_in, w_out);
  return h;
}

// Helper warp of layer 1: applies W_x^(1) (pre-scaled, in registers) to every hand-off of layer 0 and delivers the projected
// rows to layer 1 in chunks of BLK, so that layer 1 runs layer 0's loop at half its rate.
__device__ __forceinline__ void wave_proj_helper(int n, int j, const float4* myWx, const float* myBx, Handoff* hin, InRing* inr,
                                                 const L0Io io) {
  HalfW w[3];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = myWx[(q * 3 + g) * HP + j];
      w[g].m[q] = make_float2(t.x, t.y); w[g].o[q] = make_float2(t.z, t.w);
    }
  const float b0 = myBx[j], b1 = myBx[HP + j], b2 = myBx[2 * HP + j];
  for (unsigned idx = 0; idx < (unsigned)n; ++idx) {
    const int slot = idx & (HRS - 1), g = (idx / HG) & 1;
    if ((idx & (HG - 1)) == 0) {
      mbar_wait(&hin->gfull[g], (idx / HRS) & 1u);
      // layer 0's I/O: group `grp` closing a chunk (odd group, or the last one) means its 16 (or fewer) rows are final and fenced
      const int grp = idx / HG, ngrp = (n + HG - 1) / HG;
      if (j == 0 && ((grp & 1) || grp == ngrp - 1)) {
        const int c = grp >> 1, nch = (io.S + WIN - 1) / WIN;
        const int len = min(WIN, io.S - c * WIN);
        bulk_s2g(io.so + (int64_t)c * WIN * ST, io.s_out + (c & 1) * WIN * ST, (uint32_t)len * ST * 4);
        bulk_commit();
        const int cn = c + 2;
        if (cn < nch) {
          const int ln = min(WIN, io.S - cn * WIN), stage = cn % WNS0;
          mbar_expect_tx(&io.full[stage], (uint32_t)ln * G3 * 4);
          bulk_g2s(io.s_in + stage * WIN * G3, io.pp + (int64_t)cn * WIN * G3, (uint32_t)ln * G3 * 4, &io.full[stage]);
        }
      }
      __syncwarp();
    }
    float4 v[4];
    load_vec_half(v, hin->ring[slot], j);
    const Part p0 = half_part(v, w[0], b0), p1 = half_part(v, w[1], b1), p2 = half_part(v, w[2], b2);
    __syncwarp();
    if (j == 0 && ((idx & (HG - 1)) == HG - 1 || idx + 1 == (unsigned)n)) {
      // group free again.  Layer 0 rewrites output buffer c & 1 two chunks after chunk c, behind this barrier for the group
      // that follows the one whose arrival triggered the store of chunk c: by now that store has long read its source
      bulk_wait_read<0>();
      mbar_arrive(&hin->gempty[g]);
    }
    const float ar = half_finish(p0.m, p0.o), au = half_finish(p1.m, p1.o), ac = half_finish(p2.m, p2.o);
    const int stage = (idx / BLK) & 1, rowi = idx & (BLK - 1);
    if (rowi == 0 && idx >= 2 * BLK) mbar_wait(&inr->empty[stage], (idx / (2 * BLK) - 1) & 1u);
    float* o = inr->row[stage][rowi];
    o[j] = ar; o[HP + j] = au; o[2 * HP + j] = ac;
    if (rowi == BLK - 1 || idx + 1 == (unsigned)n) {
      __syncwarp();
      if (j == 0) mbar_arrive(&inr->full[stage]);
    }
  }
}

__global__ void __launch_bounds__(384)
wave_fwd_kernel(const __grid_constant__ WaveArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, w = tid >> 5, j = tid & 31;
  const int L = a.L, nspc = a.nspc;
  const int k = a.wlayer[w], si = a.wsample[w];
  const bool helper = a.whelper[w] != 0;
  const int b = blockIdx.x * nspc + si;
  const WaveSmem sm(L, nspc);
  InRing* inr = reinterpret_cast<InRing*>(dsm + sm.hand3);
  float4* sWx = reinterpret_cast<float4*>(dsm + sm.wx);
  float* sBx = reinterpret_cast<float*>(dsm + sm.bx);
  Handoff* hand = reinterpret_cast<Handoff*>(dsm + sm.hand);

  // ---- CTA setup: W_x / b_x of layers >= 1 into shared memory, K-half layout [q][g][lane] x (mine | other), times the log2(e)
  // factor of their gate; hand-off barriers ----
  for (int e = tid; e < (L - 1) * 8 * 3 * HP; e += blockDim.x) {
    const int kk = 1 + e / (8 * 3 * HP), r = e % (8 * 3 * HP);
    const int q = r / (3 * HP), g = (r / HP) % 3, lane = r % HP;
    const int kr = 16 * (lane & 1) + 2 * q, cm = g * HP + lane, co = g * HP + (lane ^ 1);
    const float sc = g == 2 ? kTwoLog2e : kNegLog2e;
    const float* Wx = a.pw + a.Wx[kk];                   // [32][96]
    sWx[e] = make_float4(sc * __ldg(Wx + kr * G3 + cm), sc * __ldg(Wx + (kr + 1) * G3 + cm),
                         sc * __ldg(Wx + kr * G3 + co), sc * __ldg(Wx + (kr + 1) * G3 + co));
  }
  float4* sWhc = reinterpret_cast<float4*>(dsm + sm.whc);
  for (int e = tid; e < (L - 2) * 8 * HP; e += blockDim.x) {
    const int kk = 2 + e / (8 * HP), q = (e / HP) % 8, lane = e % HP;
    const int kr = 16 * (lane & 1) + 2 * q;
    const float* Wc = a.pw + a.Wh[kk] + 2 * HP * HP;     // [32 i][32 j]
    sWhc[e] = make_float4(kTwoLog2e * __ldg(Wc + kr * HP + lane), kTwoLog2e * __ldg(Wc + (kr + 1) * HP + lane),
                          kTwoLog2e * __ldg(Wc + kr * HP + (lane ^ 1)), kTwoLog2e * __ldg(Wc + (kr + 1) * HP + (lane ^ 1)));
  }
  for (int e = tid; e < (L - 1) * G3; e += blockDim.x)
    sBx[e] = (e % G3 >= 2 * HP ? kTwoLog2e : kNegLog2e) * __ldg(a.pw + a.bx[1 + e / G3] + e % G3);
  for (int e = tid; e < (L - 1) * nspc * HRS; e += blockDim.x) {
    mbar_init(&hand[e / HRS].full[e % HRS], 1);
    mbar_init(&hand[e / HRS].empty[e % HRS], 1);
    if (e % HRS < 2) { mbar_init(&hand[e / HRS].gfull[e % HRS], 1); mbar_init(&hand[e / HRS].gempty[e % HRS], 1); }
  }
  if (L > 1)
    for (int e = tid; e < nspc * 2; e += blockDim.x) {
      mbar_init(&inr[e / 2].full[e % 2], 1);
      mbar_init(&inr[e / 2].empty[e % 2], 1);
    }
  fence_mbar_init();
  __syncthreads();
  pdl_trigger();                                         // the attention / dX kernel behind may be scheduled as SMs free up
  pdl_wait();                                            // everything above only read weights written two launches ago
  if (b >= a.B || k < 0) return;                         // ragged last CTA: the whole warp leaves together

  if (helper) {                                          // layer 1's projection warp
    float* s_in0 = reinterpret_cast<float*>(dsm + sm.l0 + si * sm.r0);      // layer 0's region, same carve-up as below
    float* s_out0 = s_in0 + WNS0 * WIN * G3;
    uint64_t* full0 = reinterpret_cast<uint64_t*>(s_out0 + 2 * WIN * ST + 64);
    const L0Io io{s_in0, s_out0, full0, a.proj0 + (int64_t)b * a.S[0] * G3, a.st[0] + (int64_t)b * a.S[0] * ST, a.S[0]};
    wave_proj_helper(a.S[1], j, sWx, sBx, &hand[0 * nspc + si], &inr[si], io);
    return;
  }
  unsigned char* reg = k == 0 ? dsm + sm.l0 + si * sm.r0 : dsm + sm.lk + ((k - 1) * nspc + si) * sm.r1;
  Handoff* hin = k > 0 ? &hand[(k - 1) * nspc + si] : nullptr;
  Handoff* hout = k < L - 1 ? &hand[k * nspc + si] : nullptr;
  float h;
  if (k == 0) {
    float* s_in = reinterpret_cast<float*>(reg);
    float* s_out = s_in + WNS0 * WIN * G3;
    float* sh_rh = s_out + 2 * WIN * ST;
    float* sh_h = sh_rh + 32;
    uint64_t* full = reinterpret_cast<uint64_t*>(sh_h + 32);
    sh_h[j] = 0.f;
    __syncwarp();
    if (hout == nullptr) h = wave_layer_fast<true, false>(a, 0, b, j, s_in, full, nullptr, s_out, sh_rh, sh_h, nullptr);
    else h = wave_layer_fast<true, true>(a, 0, b, j, s_in, full, nullptr, s_out, sh_rh, sh_h, hout);
  } else {
    float* s_out = reinterpret_cast<float*>(reg);
    float* sh_rh = s_out + 2 * WCH * ST;
    float* sh_h = sh_rh + 32;
    sh_h[j] = 0.f;
    __syncwarp();
    const float4* myWx = sWx + (size_t)(k - 1) * 8 * 3 * HP;
    const float* myBx = sBx + (k - 1) * G3;
    if (k == 1) {
      if (hout == nullptr) h = wave_layer_fast<false, false>(a, 1, b, j, nullptr, nullptr, &inr[si], s_out, sh_rh, sh_h, nullptr);
      else h = wave_layer_fast<false, true>(a, 1, b, j, nullptr, nullptr, &inr[si], s_out, sh_rh, sh_h, hout);
    } else if (k == 2) {
      h = wave_layer_up<true>(a, k, b, j, s_out, sh_rh, sh_h, myWx, myBx, sWhc + (size_t)(k - 2) * 8 * HP, hin, hout);
    } else {
      h = wave_layer_up<false>(a, k, b, j, s_out, sh_rh, sh_h, myWx, myBx, sWhc + (size_t)(k - 2) * 8 * HP, hin, hout);
    }
  }
  if (j < a.H) a.memory[((int64_t)b * L + k) * a.H + j] = h;   // final state -> memory slot k, hpmn.py:121
}

bool launch_wave_fwd(const Launch& L, const Dims& d, const PackLayout& pk, const float* proj0, const float* pw,
                     float* const* st, float* memory, cudaStream_t st_) {
  if (d.L > WAVE_MAX_L || !wave_supported(d.L, d.P)) return false;
  const int nspc = d.L <= 5 ? 2 : 1;                     // register budget: 12 warps x 32 x 168
  const WaveSmem sm(d.L, nspc);
  if (sm.total > 227 * 1024) return false;   // opt-in maximum of dynamic shared memory per CTA on sm_100
  WaveArgs a; memset(&a, 0, sizeof(a));
  a.proj0 = proj0; a.pw = pw; a.memory = memory;
  a.B = d.B; a.L = d.L; a.H = d.H; a.nspc = nspc;
  { static int once = 0; const char* e = getenv("HPMN_WAVE_DEBUG"); a.debug = (e && e[0] == '1' && once++ == 3) ? 1 : 0; }   // 4th call only
  for (int k = 0; k < d.L; ++k) { a.st[k] = st[k]; a.Wh[k] = pk.Wh[k]; a.Wx[k] = pk.Wx[k]; a.bx[k] = pk.bx[k]; a.S[k] = d.S[k]; a.P[k] = d.P[k]; }
  cudaFuncSetAttribute(wave_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total);
  const int grid = (d.B + nspc - 1) / nspc;
  const WarpPlan wp = plan_warps(d.L, nspc, d.L > 1, "HPMN_WAVE_WF");
  for (int i = 0; i < 16; ++i) { a.wlayer[i] = wp.layer[i]; a.wsample[i] = wp.sample[i]; a.whelper[i] = wp.helper[i]; }
  launch_pdl(wave_fwd_kernel, dim3(grid), dim3(32 * wp.n), (size_t)sm.total, st_, a);   // prologue overlaps the projection GEMM's tail
  { cudaError_t e = cudaGetLastError();                  // resources: the caller falls back to the per-layer kernels
    if (e != cudaSuccess) { if (getenv("HPMN_VERBOSE")) fprintf(stderr, "wave_fwd launch failed: %s (threads %d smem %d)\n", cudaGetErrorString(e), 32 * wp.n, sm.total); return false; } }
  ++*L.counter;
  return true;
}

// =====================================================================================================
// Backward wavefront: every layer walks its steps in reverse concurrently.  Layer k needs, at each of its firing
// steps s = (j+1)*p_k - 1, the gradient dx_{k+1}[j] = da_{k+1}[j] * W_x^{(k+1)T} of the layer above, handed down through an
// mbarrier ring: layers >= 2 compute it right after their step j (W_x^T from shared memory); layer 1 -- half as often on
// duty as layer 0, but with that extra 96 x 32 mat-vec it was the bottleneck of the whole kernel -- hands its da row to a
// helper warp that holds W_x^T in registers (wave_bwd_dx_helper), mirroring the forward kernel's projection helper.
// Saved state rows stream in with cp.async.bulk (a chunk needs rows s0-1 .. s0+len-1), da rows leave with bulk stores for
// the weight-gradient kernel and, for layer 0, the dX GEMM that feeds the embedding scatter.
// =====================================================================================================
constexpr int BCH0 = 8, BNS0 = 3;   // layers 0 and 1: steps per chunk, state ring stages
constexpr int BCHK = 4, BNSK = 2;   // layers >= 3 (not on the critical path): smaller rings
constexpr int BCH2 = 8;             // layer 2: 8-step chunks (its per-chunk I/O was 150 of its ~2000 cycles per step)

struct WaveBwdArgs {
  const float* pw;
  const float* st[HPMN_MAX_LAYERS];   // [B,S_k,128]
  float* da[HPMN_MAX_LAYERS];         // [B,S_k,96]
  const float* dmemory;               // [B,L,H]
  int64_t WhT[HPMN_MAX_LAYERS], WxT[HPMN_MAX_LAYERS];   // offsets into pw: WhT [3][32 j][32 i], WxT [96][32]
  int S[HPMN_MAX_LAYERS], P[HPMN_MAX_LAYERS];
  int B, L, H, nspc;
  signed char wlayer[16], wsample[16], whelper[16];
};

__host__ __device__ constexpr int bwd_region_bytes(int ch, int ns) {
  return ns * (ch + 1) * ST * 4 + 2 * ch * G3 * 4 + 3 * 128 + 128 + 128;   // state ring | da ring | sh_c,sh_r,sh_u | mbarriers | dh (upper warp)
}

struct WaveBwdSmem {
  int wxt, whc, hand, hand3, l0, l1, l2, lk, total;
  int wxt_first;                   // first layer whose W_x^T is kept in shared memory
  __host__ __device__ WaveBwdSmem(int L, int nspc, bool up) {
    int off = 0;
    wxt_first = up ? 2 : 1;        // upper-warp mode: the helper reads layer 1's W_x^T straight from global memory
    wxt = off; off += (L > wxt_first ? L - wxt_first : 0) * 3 * 8 * HP * 16;   // float4 [g][q][lane] = (mine pair | other pair)
    whc = off;
    if (up) off += (L > 2 ? L - 2 : 0) * 3 * 8 * HP * 16;  // W_h^T (all gates) of layers >= 2, float4 [g][q][lane]
    else off += (CS_BWD && L > 2 ? L - 2 : 0) * 8 * HP * 16;   // candidate-gate W_h^T of layers >= 2, float4 [q][lane]
    off = (off + 127) & ~127;
    hand = off; off += (L - 1) * nspc * (int)sizeof(Handoff);
    off = (off + 127) & ~127;
    hand3 = off; off += (L > 1 ? nspc : 0) * (int)sizeof(Handoff3);   // layer 1 -> its dx helper
    off = (off + 127) & ~127;
    l0 = off; off += nspc * bwd_region_bytes(BCH0, BNS0);
    l1 = off; off += (L > 1 ? nspc : 0) * bwd_region_bytes(BCH0, BNS0);   // layer 1 runs the same loop as layer 0
    l2 = off; off += (!up && L > 2 ? nspc : 0) * bwd_region_bytes(BCH2, BNSK); // layer 2: next in line for the critical path
    lk = off; off += (up ? (L > 2 ? L - 2 : 0) : (L > 3 ? L - 3 : 0)) * nspc * bwd_region_bytes(BCHK, BNSK);
    total = off;
  }
};

struct BwdW { HalfW r, u, c; };        // W_h^T: lane i owns dh_prev[i], inputs are the 32 gate pre-activation gradients
__device__ __forceinline__ void load_bwd_weights(BwdW& w, const float* WhT /*[3][32 j][32 i]*/, int lane) {
  load_half(w.r, WhT, lane, 1.f);
  load_half(w.u, WhT + HP * HP, lane, 1.f);
  load_half(w.c, WhT + 2 * HP * HP, lane, 1.f);
}

// One reverse GRU step (hand-derived adjoint of util.py:81-110).
// The dependent chain is  dh -> da_c (one FMUL) -> broadcast -> (da_c Wc^T) -> da_r (one FMUL) -> broadcast -> (da_r Wr^T)
// -> dh';  every factor that does not depend on dh (1-u, 1-c^2, (h_prev-c) u (1-u), r (1-r) h_prev) is formed from the saved
// state before dh arrives, da_u is broadcast together with da_c so that its mat-vec fills the bubbles of the chain.
// row = this lane's column of buffer row t (h_{s-1}; row t+1 holds r|u|c of step s); orow = of the da row; add = gradient
// arriving from the layer above for the step BELOW this one (0 if that step did not fire).  WITH_DX (layers >= 2): the
// gradient handed down to the layer below, dx = da W_x^T, is accumulated from the same broadcast halves as they arrive
// (W_x^T from shared memory, [g][q][mine|other][lane]) -- nothing but six accumulators stays live for it.
// The caller closes the step with __syncwarp().
// ALL_SMEM (layers >= 3 of the upper warp): all three recurrent matrices come from shared memory (myWhT: [g][q][lane]).
template <bool WITH_DX, bool ALL_SMEM = false>
__device__ __forceinline__ float gru_bwd_step(const BwdW& w, const float* row, bool first_step, float dh, float add, int i,
                                              float* sh_c, float* sh_r, float* sh_u, float* orow, float& dar, float& dau,
                                              float& dac, const float4* myWxT, const float4* myWhcT, float& dx) {
  const float hp = first_step ? 0.f : row[0];            // zero state before step 0
  const float r = row[ST + HP], u = row[ST + 2 * HP], c = row[ST + 3 * HP];
  const float omu = 1.f - u;
  const float gc = omu * fmaf(-c, c, 1.f);               // d a_c / d h'
  const float gu = (hp - c) * u * omu;                   // d a_u / d h'
  const float gr = r * (1.f - r) * hp;                   // d a_r / d (r o h_prev)
  dac = dh * gc; dau = dh * gu;
  sh_c[i] = dac;
  sh_u[i] = dau;
  __syncwarp();
  float4 v[4];
  load_vec_half(v, sh_c, i);
  if (!WITH_DX) {                                        // latency-tuned order (layers 0 and 1): da_u's mat-vec fills the bubbles
    float4 vu[4];
    load_vec_half(vu, sh_u, i);
    const Part pc = half_part(v, w.c, 0.f);
    const float drh = half_finish(pc.m, pc.o);           // (da_c Wc^T)[Din + i]
    dar = drh * gr;
    sh_r[i] = dar;
    __syncwarp();
    load_vec_half(v, sh_r, i);
    const Part pu = half_part(vu, w.u, fmaf(dh, u, add));
    const Part pr = half_part(v, w.r, drh * r);
    orow[0] = dar; orow[HP] = dau; orow[2 * HP] = dac;
    return half_finish(pr.m + pu.m, pr.o + pu.o);        // gradient wrt h_{s-1} (+ the share of the layer above)
  }
  // layers >= 2: one broadcast vector live at a time, so that the W_x^T loads (shared memory) can be batched in the ~35
  // registers next to the recurrent weights
  float2 xm0 = make_float2(0.f, 0.f), xo0 = xm0, xm1 = xm0, xo1 = xm0, xm2 = xm0, xo2 = xm0;
  float2 cm = xm0, co = xm0;
  float drh;
  if (ALL_SMEM) {
    half_part_smem(v, myWhcT + (2 * 8) * HP + i, HP, cm, co);
    drh = half_finish(cm.x + cm.y, co.x + co.y);
  } else if (CS_BWD) {
    half_part_smem(v, myWhcT + i, HP, cm, co);           // candidate-gate W_h^T from shared memory (see gru_fwd_step, C_SMEM)
    drh = half_finish(cm.x + cm.y, co.x + co.y);
  } else {
    const Part pc = half_part(v, w.c, 0.f);
    drh = half_finish(pc.m, pc.o);
  }
  dar = drh * gr;
  sh_r[i] = dar;
  __syncwarp();
  half_part_smem(v, myWxT + (2 * 8) * HP + i, HP, xm2, xo2);                    // rows 64..95 of W_x^T: da_c
  load_vec_half(v, sh_u, i);
  Part pu, pr;
  if (ALL_SMEM) {
    float2 m = make_float2(fmaf(dh, u, add), 0.f), o = make_float2(0.f, 0.f);
    half_part_smem(v, myWhcT + (1 * 8) * HP + i, HP, m, o);
    pu.m = m.x + m.y; pu.o = o.x + o.y;
  } else {
    pu = half_part(v, w.u, fmaf(dh, u, add));
  }
  half_part_smem(v, myWxT + (1 * 8) * HP + i, HP, xm1, xo1);                    // rows 32..63: da_u
  load_vec_half(v, sh_r, i);
  if (ALL_SMEM) {
    float2 m = make_float2(drh * r, 0.f), o = make_float2(0.f, 0.f);
    half_part_smem(v, myWhcT + (0 * 8) * HP + i, HP, m, o);
    pr.m = m.x + m.y; pr.o = o.x + o.y;
  } else {
    pr = half_part(v, w.r, drh * r);
  }
  half_part_smem(v, myWxT + (0 * 8) * HP + i, HP, xm0, xo0);                    // rows 0..31: da_r
  orow[0] = dar; orow[HP] = dau; orow[2 * HP] = dac;
  const float dhn = half_finish(pr.m + pu.m, pr.o + pu.o);
  dx = half_finish(((xm0.x + xm0.y) + (xm1.x + xm1.y)) + (xm2.x + xm2.y), ((xo0.x + xo0.y) + (xo1.x + xo1.y)) + (xo2.x + xo2.y));
  return dhn;
}

// ---------------------------------------------------------------------------------------------------------------------
// Layers 0 and 1: latency-tuned loop.  One chunk = 8 steps = one hand-off group (4 rows of the layer above, one barrier
// round trip); no branch or address arithmetic inside a chunk.  FIRE2 = false: no layer above.  OUT_DA (layer 1): hand the
// da row of every step to the dx helper.
// ---------------------------------------------------------------------------------------------------------------------
// Layer 0's chunk I/O as seen by the dx helper that serves it (see wave_bwd_fast, SERVED)
struct L0BwdIo { float* s_st; float* s_da; uint64_t* full; const float* sb; float* dab; int S; };

// SERVED (layer 0 with a helper): the dx helper issues this warp's bulk stores and state-ring refills.  It learns that chunk
// `it` is consumed from the group-empty barrier (this warp fences its da rows for the async proxy before that arrive) and
// serves it before it produces group it+2 -- so the group-full barrier this warp waits on anyway for chunk it+2 also
// tells it that da buffer it & 1 is free again and that the state rows of chunk it+3 are on their way.
template <bool FIRE2, bool OUT_DA, bool SERVED, bool DBG>
__device__ __forceinline__ void wave_bwd_fast(const WaveBwdArgs& a, int k, int b, int i, unsigned char* reg, Handoff* hin,
                                              Handoff3* hda) {
  static_assert(!SERVED || FIRE2, "only a layer with a layer above has a helper");
  constexpr int CH = BCH0, NS = BNS0;
  static_assert(CH == BLK && CH == 2 * HG, "one chunk = one hand-off group of the layer above");
  long long w_in = 0, w_out = 0, w_tma = 0;               // DBG: cycles blocked on hand-off in / out and on the TMA ring
  const bool dbg = DBG && blockIdx.x == 0;
  const long long t_start = DBG ? clock64() : 0;
  const int S = a.S[k], H = a.H, L = a.L;
  unsigned sent = 0;                                     // OUT_DA: da rows handed to the helper so far
  float* s_st = reinterpret_cast<float*>(reg);           // [NS][(CH+1)*ST]
  float* s_da = s_st + NS * (CH + 1) * ST;               // [2][CH*G3]
  float* sh_c = s_da + 2 * CH * G3;
  float* sh_r = sh_c + 32;
  float* sh_u = sh_r + 32;
  uint64_t* full = reinterpret_cast<uint64_t*>(sh_u + 32);

  BwdW w;
  load_bwd_weights(w, a.pw + a.WhT[k], i);
  const float* sb = a.st[k] + (int64_t)b * S * ST;
  float* dab = a.da[k] + (int64_t)b * S * G3;
  const int nch = (S + CH - 1) / CH;

  auto issue = [&](int ci, int stage) {                  // lane 0: load state rows s0-1 .. s0+len-1 of chunk ci
    const int s0 = ci * CH;
    const int len = min(CH, S - s0);
    const uint32_t rows = (uint32_t)(s0 > 0 ? len + 1 : len);
    mbar_expect_tx(&full[stage], rows * ST * 4);
    if (s0 > 0) bulk_g2s(s_st + stage * (CH + 1) * ST, sb + (int64_t)(s0 - 1) * ST, rows * ST * 4, &full[stage]);
    else bulk_g2s(s_st + stage * (CH + 1) * ST + ST, sb, rows * ST * 4, &full[stage]);    // buffer row t+1 <-> step s0+t
  };
  if (i == 0) {
    for (int q = 0; q < NS; ++q) mbar_init(&full[q], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (i == 0)
    for (int it = 0; it < NS && it < nch; ++it) issue(nch - 1 - it, it);

  float dh = i < H ? __ldg(a.dmemory + ((int64_t)b * L + k) * H + i) : 0.f;   // memory-slot gradient enters at the last step

  auto step = [&](const float* row, float* orow, bool first_step, float add) {
    float dar, dau, dac, dx_unused;
    dh = gru_bwd_step<false>(w, row, first_step, dh, add, i, sh_c, sh_r, sh_u, orow, dar, dau, dac, nullptr, nullptr, dx_unused);
    if (OUT_DA) {                                        // every step of layer 1 fed a firing step of layer 0: da row -> helper
      const int slot = sent & (HRS - 1), g = (sent / HG) & 1;
      if ((sent & (HG - 1)) == 0 && sent >= HRS) mbar_wait_t(&hda->gempty[g], (sent / HRS - 1) & 1u, w_out, dbg);
      hda->ring[slot][i] = dar; hda->ring[slot][HP + i] = dau; hda->ring[slot][2 * HP + i] = dac;
      ++sent;
      if ((sent & (HG - 1)) == 0 || sent == (unsigned)S) {   // one barrier round trip per HG rows
        __syncwarp();
        if (i == 0) mbar_arrive(&hda->gfull[g]);
      }
    }
    __syncwarp();                                        // sh_c / sh_r / sh_u free for the next step
  };

  for (int it = 0; it < nch; ++it) {
    const int ci = nch - 1 - it;
    const int stage = it % NS;
    const int s0 = ci * CH;
    const int len = min(CH, S - s0);
    const int g = it & 1;
    mbar_wait_t(&full[stage], (uint32_t)(it / NS) & 1u, w_tma, dbg);
    float* ob = s_da + (it & 1) * CH * G3 + i;
    if (!SERVED && it >= 2) {
      if (i == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    const float* ib = s_st + stage * (CH + 1) * ST + i;
    const float* rg = nullptr;
    if (FIRE2) {                                         // the 4 (top chunk: len/2) rows the layer above hands down for this chunk
      rg = &hin->ring[g * HG][i];
      if (it == 0) {                                     // later groups are awaited at the end of the previous iteration
        mbar_wait_t(&hin->gfull[0], 0u, w_in, dbg);
        dh += rg[(HG - 1 - ((len - 1) >> 1)) * HP];      // the last step of the sequence fired
      }
    }
    // `add` of step t belongs to step t-1: row n = HG-1 - ((t-1)>>1) of the group if t-1 is odd (a firing step); the add for
    // this chunk's step 0 is the first row of the NEXT group and is applied at the end of this iteration.
    if (len == CH && s0 > 0) {
#pragma unroll 1
      for (int u = CH / BUNR - 1; u >= 0; --u) {         // the unrolled body must stay inside the instruction caches (see FUNR)
#pragma unroll
        for (int tt = BUNR - 1; tt >= 0; --tt) {
          const int t = u * BUNR + tt;                   // parity of t is tt's: which steps take an `add` is known at compile time
          float add = 0.f;
          if (FIRE2 && ((tt - 1) & 1) && (tt > 0 || u > 0)) add = rg[(HG - 1 - ((t - 1) >> 1)) * HP];
          step(ib + t * ST, ob + t * G3, false, add);
        }
      }
    } else {
      for (int t = len - 1; t >= 0; --t)
        step(ib + t * ST, ob + t * G3, s0 + t == 0, (FIRE2 && t > 0 && ((t - 1) & 1)) ? rg[(HG - 1 - ((t - 1) >> 1)) * HP] : 0.f);
    }
    if (SERVED) { fence_proxy_async(); __syncwarp(); }   // da rows -> async proxy before the helper is told the chunk is done
    if (FIRE2) {
      if (i == 0) mbar_arrive(&hin->gempty[g]);          // every lane's reads of the group precede the step's last __syncwarp
      if (it + 1 < nch) {                                // step s0-1 (odd) fired: its share is row 0 of the next group
        mbar_wait_t(&hin->gfull[g ^ 1], (uint32_t)((it + 1) / 2) & 1u, w_in, dbg);
        dh += hin->ring[(g ^ 1) * HG][i];
      }
    }
    if (SERVED) continue;
    fence_proxy_async();
    __syncwarp();
    if (i == 0) {
      bulk_s2g(dab + (int64_t)s0 * G3, ob - i, (uint32_t)len * G3 * 4);
      bulk_commit();
      if (it + NS < nch) issue(nch - 1 - (it + NS), stage);
    }
  }
  if (!SERVED && i == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && i == 0 && b == 0)
    printf("wave_bwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma %lld\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, w_in, w_out, w_tma);
}

// Helper warp of layer 1: dx = da W_x^(1)T for every step of layer 1, W_x^T in registers, handed to layer 0 in groups of
// HG rows aligned with layer 0's chunks.
__device__ __forceinline__ void wave_bwd_dx_helper(const WaveBwdArgs& a, int i, const float4* myWxT, Handoff3* hda, Handoff* hout,
                                                   const L0BwdIo io) {
  constexpr int CH = BCH0, NS = BNS0;
  const int nch0 = (io.S + CH - 1) / CH;
  auto serve = [&](int it) {                             // lane 0: layer 0 has consumed chunk `it` (reverse order) and fenced its da rows
    const int s0 = (nch0 - 1 - it) * CH, len = min(CH, io.S - s0);
    bulk_s2g(io.dab + (int64_t)s0 * G3, io.s_da + (it & 1) * CH * G3, (uint32_t)len * G3 * 4);
    bulk_commit();
    if (it + NS < nch0) {                                // refill the state stage the chunk has just released
      const int ci = nch0 - 1 - (it + NS), stage = it % NS, t0 = ci * CH, ln = min(CH, io.S - t0);
      const uint32_t rows = (uint32_t)(t0 > 0 ? ln + 1 : ln);
      mbar_expect_tx(&io.full[stage], rows * ST * 4);
      if (t0 > 0) bulk_g2s(io.s_st + stage * (CH + 1) * ST, io.sb + (int64_t)(t0 - 1) * ST, rows * ST * 4, &io.full[stage]);
      else bulk_g2s(io.s_st + stage * (CH + 1) * ST + ST, io.sb, rows * ST * 4, &io.full[stage]);
    }
  };
  HalfW w[3];                                            // gate g: rows g*32 .. g*32+31 of W_x^T
  if (myWxT != nullptr) {
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float4 t = myWxT[(g * 8 + q) * HP + i]; w[g].m[q] = make_float2(t.x, t.y); w[g].o[q] = make_float2(t.z, t.w); }
  } else {                                               // upper-warp mode: layer 1's W_x^T is not staged in shared memory
#pragma unroll
    for (int g = 0; g < 3; ++g) load_half(w[g], a.pw + a.WxT[1] + g * HP * HP, i, 1.f);   // rows g*32.. of [96][32] = [32 in][32 out]
  }
  const unsigned n = (unsigned)a.S[1], vofs = group_vofs(a.S[0]);
  for (unsigned idx = 0; idx < n; ++idx) {
    const int slot = idx & (HRS - 1), gi = (idx / HG) & 1;
    if ((idx & (HG - 1)) == 0) mbar_wait(&hda->gfull[gi], (idx / HRS) & 1u);
    float4 v0[4], v1[4], v2[4];
    load_vec_half(v0, hda->ring[slot], i);
    load_vec_half(v1, hda->ring[slot] + HP, i);
    load_vec_half(v2, hda->ring[slot] + 2 * HP, i);
    const Part p0 = half_part(v0, w[0], 0.f), p1 = half_part(v1, w[1], 0.f), p2 = half_part(v2, w[2], 0.f);
    const float dx = half_finish((p0.m + p1.m) + p2.m, (p0.o + p1.o) + p2.o);
    __syncwarp();
    if (i == 0 && ((idx & (HG - 1)) == HG - 1 || idx + 1 == n)) mbar_arrive(&hda->gempty[gi]);   // group consumed
    const unsigned v = idx + vofs;                       // virtual hand-off index: group = v / HG, ring slot = v % HRS
    const int g = (v / HG) & 1;
    if ((v & (HG - 1)) == 0 && v >= HRS) {
      mbar_wait(&hout->gempty[g], (v / HRS - 1) & 1u);   // layer 0 is done with group v/HG - 2 = its chunk of that index
      if (i == 0) serve((int)(v / HG) - 2);
      __syncwarp();
    }
    hout->ring[v & (HRS - 1)][i] = dx;
    if (((v + 1) & (HG - 1)) == 0 || idx + 1 == n) {
      __syncwarp();
      // layer 0 rewrites da buffer it & 1 in chunk it+2, behind this barrier: the store of chunk it (issued when this group
      // was started) must have read its source by then -- it has, four layer-1 steps later
      if (i == 0) { bulk_wait_read<0>(); mbar_arrive(&hout->gfull[g]); }
    }
  }
  for (int G = max(0, nch0 - 2); G < nch0; ++G) {        // the last two chunks of layer 0
    mbar_wait(&hout->gempty[G & 1], (uint32_t)(G / 2) & 1u);
    if (i == 0) serve(G);
    __syncwarp();
  }
  if (i == 0) bulk_wait_read<0>();
}

// Layers >= 2: any period; dx = da W_x^T in-warp with W_x^T from shared memory (K-half layout), per-slot hand-off below,
// except towards layer 1 (k == 2), which consumes groups.
template <int CH, int NS, bool DBG>
__device__ __forceinline__ void wave_bwd_layer(const WaveBwdArgs& a, int k, int b, int i, unsigned char* reg, Handoff* hin,
                                               Handoff* hout, const float4* myWxT, const float4* myWhcT) {
  long long w_in = 0, w_out = 0, w_tma = 0;               // DBG: cycles blocked on hand-off in / out and on the TMA ring
  const bool dbg = DBG && blockIdx.x == 0;
  const long long t_start = DBG ? clock64() : 0;
  const int S = a.S[k], H = a.H, L = a.L;
  const int period = k < L - 1 ? a.P[k] : 1;             // firing period towards layer k+1
  float* s_st = reinterpret_cast<float*>(reg);           // [NS][(CH+1)*ST]
  float* s_da = s_st + NS * (CH + 1) * ST;               // [2][CH*G3]
  float* sh_c = s_da + 2 * CH * G3;
  float* sh_r = sh_c + 32;
  float* sh_u = sh_r + 32;
  uint64_t* full = reinterpret_cast<uint64_t*>(sh_u + 32);

  BwdW w;
  load_half(w.r, a.pw + a.WhT[k], i, 1.f);              // r and u in registers, c from shared memory (myWhcT)
  load_half(w.u, a.pw + a.WhT[k] + HP * HP, i, 1.f);
  if (!CS_BWD) load_half(w.c, a.pw + a.WhT[k] + 2 * HP * HP, i, 1.f);
  const float* sb = a.st[k] + (int64_t)b * S * ST;
  float* dab = a.da[k] + (int64_t)b * S * G3;
  const int nch = (S + CH - 1) / CH;

  auto issue = [&](int ci, int stage) {                  // lane 0: load state rows s0-1 .. s0+len-1 of chunk ci
    const int s0 = ci * CH;
    const int len = min(CH, S - s0);
    const uint32_t rows = (uint32_t)(s0 > 0 ? len + 1 : len);
    mbar_expect_tx(&full[stage], rows * ST * 4);
    if (s0 > 0) bulk_g2s(s_st + stage * (CH + 1) * ST, sb + (int64_t)(s0 - 1) * ST, rows * ST * 4, &full[stage]);
    else bulk_g2s(s_st + stage * (CH + 1) * ST + ST, sb, rows * ST * 4, &full[stage]);    // buffer row t+1 <-> step s0+t
  };
  if (i == 0) {
    for (int q = 0; q < NS; ++q) mbar_init(&full[q], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (i == 0)
    for (int it = 0; it < NS && it < nch; ++it) issue(nch - 1 - it, it);

  float dh = i < H ? __ldg(a.dmemory + ((int64_t)b * L + k) * H + i) : 0.f;   // memory-slot gradient enters at the last step
  unsigned got = 0, sent = 0;                            // hand-offs consumed / produced
  int to_fire = 1;                                       // the last step is a firing step: S % p == 0
  const bool grouped_out = k == 2;                       // layer 1 consumes through wave_bwd_fast
  const unsigned vofs = grouped_out ? group_vofs(a.S[1]) : 0u;

  long long q_hin = 0, q_step = 0, q_out = 0;             // DBG: cycles in the hand-in block, the step proper, the hand-down block
  auto step = [&](const float* row, float* orow, bool first_step) {
    const long long c0 = dbg ? clock64() : 0;
    if (hin != nullptr && --to_fire == 0) {              // this step fed layer k+1 in the forward pass
      to_fire = period;
      const int slot = got & (HRS - 1);
      mbar_wait_t(&hin->full[slot], (got / HRS) & 1u, w_in, dbg);
      dh += hin->ring[slot][i];
      ++got;
      __syncwarp();
      if (i == 0) mbar_arrive(&hin->empty[slot]);
    }
    float dar, dau, dac, dx;
    const long long c1 = dbg ? clock64() : 0;
    dh = gru_bwd_step<true>(w, row, first_step, dh, 0.f, i, sh_c, sh_r, sh_u, orow, dar, dau, dac, myWxT, myWhcT, dx);
    const long long c2 = dbg ? clock64() + (long long)(dx == 12345.f) : 0;
    {                                                    // dx of this step -> layer k-1 (every step of layer k is one of its firing steps)
      if (grouped_out) {
        const unsigned v = sent + vofs;                  // virtual hand-off index: group = v / HG, ring slot = v % HRS
        const int g = (v / HG) & 1;
        if ((v & (HG - 1)) == 0 && v >= HRS) mbar_wait_t(&hout->gempty[g], (v / HRS - 1) & 1u, w_out, dbg);
        hout->ring[v & (HRS - 1)][i] = dx;
        ++sent;
        if (((v + 1) & (HG - 1)) == 0 || sent == (unsigned)S) {
          __syncwarp();
          if (i == 0) mbar_arrive(&hout->gfull[g]);
        }
      } else {
        const int slot = sent & (HRS - 1);
        if (sent >= HRS) mbar_wait_t(&hout->empty[slot], (sent / HRS - 1) & 1u, w_out, dbg);
        hout->ring[slot][i] = dx;
        ++sent;
        __syncwarp();
        if (i == 0) mbar_arrive(&hout->full[slot]);
      }
    }
    __syncwarp();                                        // sh_c / sh_r / sh_u free for the next step
    if (dbg) { const long long c3 = clock64(); q_hin += c1 - c0; q_step += c2 - c1; q_out += c3 - c2; }
  };

  for (int it = 0; it < nch; ++it) {
    const int ci = nch - 1 - it;
    const int stage = it % NS;
    const int s0 = ci * CH;
    const int len = min(CH, S - s0);
    mbar_wait_t(&full[stage], (uint32_t)(it / NS) & 1u, w_tma, dbg);
    float* ob = s_da + (it & 1) * CH * G3 + i;
    if (it >= 2) {
      if (i == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    const float* ib = s_st + stage * (CH + 1) * ST + i;
    if (len == CH && s0 > 0) {
#pragma unroll GUNR
      for (int t = CH - 1; t >= 0; --t) step(ib + t * ST, ob + t * G3, false);
    } else {
      for (int t = len - 1; t >= 0; --t) step(ib + t * ST, ob + t * G3, s0 + t == 0);
    }
    fence_proxy_async();
    __syncwarp();
    if (i == 0) {
      bulk_s2g(dab + (int64_t)s0 * G3, ob - i, (uint32_t)len * G3 * 4);
      bulk_commit();
      if (it + NS < nch) issue(nch - 1 - (it + NS), stage);
    }
  }
  if (i == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && i == 0 && b == 0)
    printf("wave_bwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma %lld  [hand-in %lld  step %lld  hand-down %lld]\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, w_in, w_out, w_tma, q_hin, q_step, q_out);
}


// ---------------------------------------------------------------------------------------------------------------------
// Upper warp (backward): ONE warp per sample walks layers 2 .. L-1.  With one warp per layer the CTA had 12 warps, i.e. 168
// registers per thread, and the layers >= 2 -- 96 registers of recurrent weights plus a 96 x 32 mat-vec with W_x^T from
// shared memory -- could keep only a handful of weight loads in flight: ~1400 cycles per step, which made layer 2 the
// slowest link of the whole backward pipeline (profiles/r1_v10_wave_ncu.md).  Together these layers fire 0.44 times per
// layer-0 step, so one warp has the time; 8 warps per CTA lift the cap to 255 registers, and the hand-offs between the
// layers >= 2 need no barriers any more (the dx of layer k+1 is simply a register of the same lane).
// For every step of layer 2 (reverse time) the layers above that fired at that step are processed first, top down
// (hpmn.py:124-128 in reverse).  Layer 2 keeps its W_h^T in registers, layers >= 3 read all weights from shared memory.
// ---------------------------------------------------------------------------------------------------------------------
template <bool DBG>
__device__ __forceinline__ void wave_bwd_upper(const WaveBwdArgs& a, int b, int i, unsigned char* dsm, const WaveBwdSmem& sm,
                                               int si, Handoff* hout, const float4* sWxT, const float4* sWhT3) {
  constexpr int NS = BNSK;
  const int L = a.L, nspc = a.nspc, H = a.H;
  const bool dbg = DBG && blockIdx.x == 0;
  const long long t_start = DBG ? clock64() : 0;
  long long w_out = 0, w_tma = 0, q_step = 0;
  BwdW w2;                                               // unused: every layer of this warp reads its weights from shared memory

  // per-layer bookkeeping lives in the layer's shared-memory block (no runtime-indexed register arrays, no integer divisions:
  // the first cut of this loop spent 1400 cycles per step on `/` and `%`): st[0] = next step to process, st[1] = steps until
  // the next firing step (0 = this one fired)
  struct Lay { float* s_st; float* s_da; float* sh_c; float* sh_r; float* sh_u; uint64_t* full; float* sh_dh; int* st; int CH, chs; };
  auto layer = [&](int k) {
    Lay l;
    l.chs = 2;
    static_assert(BCHK == 4, "chunk size is a power of two (shifts below)");
    l.CH = 1 << l.chs;
    unsigned char* reg = dsm + sm.lk + ((k - 2) * nspc + si) * bwd_region_bytes(BCHK, BNSK);
    l.s_st = reinterpret_cast<float*>(reg);
    l.s_da = l.s_st + NS * (l.CH + 1) * ST;
    l.sh_c = l.s_da + 2 * l.CH * G3;
    l.sh_r = l.sh_c + 32;
    l.sh_u = l.sh_r + 32;
    l.full = reinterpret_cast<uint64_t*>(l.sh_u + 32);
    l.st = reinterpret_cast<int*>(l.full + 8);
    l.sh_dh = l.sh_u + 64;
    return l;
  };
  auto issue = [&](int k, const Lay& l, int ci, int stage) {   // lane 0: state rows s0-1 .. s0+len-1 of chunk ci of layer k
    const int S = a.S[k], s0 = ci * l.CH, len = min(l.CH, S - s0);
    const float* sb = a.st[k] + (int64_t)b * S * ST;
    const uint32_t rows = (uint32_t)(s0 > 0 ? len + 1 : len);
    mbar_expect_tx(&l.full[stage], rows * ST * 4);
    if (s0 > 0) bulk_g2s(l.s_st + stage * (l.CH + 1) * ST, sb + (int64_t)(s0 - 1) * ST, rows * ST * 4, &l.full[stage]);
    else bulk_g2s(l.s_st + stage * (l.CH + 1) * ST + ST, sb, rows * ST * 4, &l.full[stage]);
  };
  for (int k = 2; k < L; ++k) {
    const Lay l = layer(k);
    if (i == 0) { for (int q = 0; q < NS; ++q) mbar_init(&l.full[q], 1); l.st[0] = a.S[k] - 1; l.st[1] = 0; }
    l.sh_dh[i] = i < H ? __ldg(a.dmemory + ((int64_t)b * L + k) * H + i) : 0.f;   // memory-slot gradient enters at the last step
  }
  if (i == 0) fence_mbar_init();
  __syncwarp();
  if (i == 0)
    for (int k = 2; k < L; ++k) {
      const Lay l = layer(k);
      const int nch = (a.S[k] + l.CH - 1) >> l.chs;
      for (int it = 0; it < NS && it < nch; ++it) issue(k, l, nch - 1 - it, it);
    }
  __syncwarp();

  // one reverse step of layer k at its step s; dx_in = gradient handed down by layer k+1 (0 if step s did not fire); returns the
  // dx for layer k-1
  auto step = [&](int k, const Lay& l, int s, float dx_in) -> float {
    const int S = a.S[k], CH = l.CH, nch = (S + CH - 1) >> l.chs;
    const int ci = s >> l.chs, t = s & (CH - 1), it = nch - 1 - ci, stage = it % NS, s0 = ci << l.chs, len = min(CH, S - s0);
    if (t == len - 1) {                                  // first step of this chunk (reverse order)
      mbar_wait_t(&l.full[stage], (uint32_t)(it / NS) & 1u, w_tma, dbg);
      if (it >= 2) {                                     // da buffer it & 1: its store two chunks ago has read its source.  Bulk groups are
        if (i == k - 2) bulk_wait_read<1>();             // per thread: lane k-2 issues only layer k's stores, so "all but the newest"
        __syncwarp();                                    // is exactly that store (one lane for all layers blocked ~2 us per chunk)
      }
    }
    const float* row = l.s_st + stage * (CH + 1) * ST + t * ST + i;
    float* ob = l.s_da + (it & 1) * CH * G3;
    float* orow = ob + t * G3 + i;
    const float dh = l.sh_dh[i] + dx_in;
    float dar, dau, dac, dx;
    const long long c0 = dbg ? clock64() : 0;
    float dhn;
    // ONE copy of the step for every layer (weights from shared memory): with a second, register-weight variant for layer 2
    // the loop body was ~13 KB of SASS and the warp lost ~900 cycles per step outside the arithmetic
    dhn = gru_bwd_step<true, true>(w2, row, s == 0, dh, 0.f, i, l.sh_c, l.sh_r, l.sh_u, orow, dar, dau, dac,
                                   sWxT + (size_t)(k - 2) * 3 * 8 * HP, sWhT3 + (size_t)(k - 2) * 3 * 8 * HP, dx);
    if (dbg) q_step += clock64() - c0 + (long long)(dx == 12345.f);
    l.sh_dh[i] = dhn;
    __syncwarp();                                        // sh_c / sh_r / sh_u free for this layer's next step
    if (t == 0) {                                        // chunk complete: da rows out, next state chunk in
      fence_proxy_async();
      __syncwarp();
      if (i == k - 2) {
        bulk_s2g(a.da[k] + ((int64_t)b * S + s0) * G3, ob, (uint32_t)len * G3 * 4);
        bulk_commit();
        if (it + NS < nch) issue(k, l, nch - 1 - (it + NS), stage);
      }
    }
    return dx;
  };

  const int S2 = a.S[2];
  const unsigned vofs = group_vofs(a.S[1]);
  unsigned sent = 0;                                     // rows handed to layer 1
  for (int s2 = S2 - 1; s2 >= 0; --s2) {
    int top = 2;                                         // layers 3 .. top fired at this step of layer 2
    while (top + 1 < L && layer(top).st[1] == 0) ++top;
    float dx = 0.f;
    for (int k = top; k >= 2; --k) {
      const Lay l = layer(k);
      const int sk = l.st[0], rem = l.st[1];
      dx = step(k, l, sk, dx);                           // (its __syncwarp()s order the bookkeeping reads above before the write below)
      if (i == 0) { l.st[0] = sk - 1; l.st[1] = rem == 0 ? a.P[k] - 1 : rem - 1; }
    }
    __syncwarp();
    // dx of layer 2 -> layer 1 (wave_bwd_fast): groups of HG rows aligned to ITS chunks
    const unsigned v = sent + vofs;
    const int g = (v / HG) & 1;
    if ((v & (HG - 1)) == 0 && v >= HRS) mbar_wait_t(&hout->gempty[g], (v / HRS - 1) & 1u, w_out, dbg);
    hout->ring[v & (HRS - 1)][i] = dx;
    ++sent;
    if (((v + 1) & (HG - 1)) == 0 || sent == (unsigned)S2) {
      __syncwarp();
      if (i == 0) mbar_arrive(&hout->gfull[g]);
    }
  }
  if (i < L - 2) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && i == 0 && b == 0)
    printf("wave_bwd upper warp (layers 2..%d): %d layer-2 steps total %lld cyc (%lld/step)  wait_out %lld  wait_tma %lld  in-steps %lld\n",
           L - 1, S2, clock64() - t_start, (clock64() - t_start) / S2, w_out, w_tma, q_step);
}

// Registers are partitioned per SM sub-partition (16 K each): 12 warps = 3 per SMSP = at most 168 per thread; the upper-warp
// variant (UP) runs 8 warps = 2 per SMSP = 255 per thread.
template <bool DBG, bool UP>
__global__ void __launch_bounds__(UP ? 256 : 384)
wave_bwd_kernel(const __grid_constant__ WaveBwdArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, w = tid >> 5, i = tid & 31;
  const int L = a.L, nspc = a.nspc;
  const int k = a.wlayer[w], si = a.wsample[w];         // see plan_warps()
  const int b = blockIdx.x * nspc + si;
  const WaveBwdSmem sm(L, nspc, UP);
  float4* sWxT = reinterpret_cast<float4*>(dsm + sm.wxt);
  Handoff* hand = reinterpret_cast<Handoff*>(dsm + sm.hand);
  Handoff3* hand3 = reinterpret_cast<Handoff3*>(dsm + sm.hand3);

  // ---- CTA setup: W_x^T of layers >= wxt_first in the K-half layout [g][q][lane] x (mine | other) ----
  const int nwx = L > sm.wxt_first ? L - sm.wxt_first : 0;
  for (int e = tid; e < nwx * 3 * 8 * HP; e += blockDim.x) {
    const int kk = sm.wxt_first + e / (3 * 8 * HP), r = e % (3 * 8 * HP);
    const int g = r / (8 * HP), q = (r / HP) % 8, lane = r % HP;
    const int n = g * HP + 16 * (lane & 1) + 2 * q;
    const float* WxT = a.pw + a.WxT[kk];                 // [96][32]
    sWxT[e] = make_float4(__ldg(WxT + n * HP + lane), __ldg(WxT + (n + 1) * HP + lane),
                          __ldg(WxT + n * HP + (lane ^ 1)), __ldg(WxT + (n + 1) * HP + (lane ^ 1)));
  }
  float4* sWhcT = reinterpret_cast<float4*>(dsm + sm.whc);
  if (UP) {                                              // W_h^T, all gates, of layers >= 2: [g][q][lane]
    for (int e = tid; e < (L - 2) * 3 * 8 * HP; e += blockDim.x) {
      const int kk = 2 + e / (3 * 8 * HP), r = e % (3 * 8 * HP);
      const int g = r / (8 * HP), q = (r / HP) % 8, lane = r % HP;
      const int n = 16 * (lane & 1) + 2 * q;
      const float* M = a.pw + a.WhT[kk] + g * HP * HP;   // [32 j][32 i]
      sWhcT[e] = make_float4(__ldg(M + n * HP + lane), __ldg(M + (n + 1) * HP + lane),
                             __ldg(M + n * HP + (lane ^ 1)), __ldg(M + (n + 1) * HP + (lane ^ 1)));
    }
  } else if (CS_BWD) {
    for (int e = tid; e < (L - 2) * 8 * HP; e += blockDim.x) {
      const int kk = 2 + e / (8 * HP), q = (e / HP) % 8, lane = e % HP;
      const int n = 16 * (lane & 1) + 2 * q;
      const float* WcT = a.pw + a.WhT[kk] + 2 * HP * HP;   // [32 j][32 i]
      sWhcT[e] = make_float4(__ldg(WcT + n * HP + lane), __ldg(WcT + (n + 1) * HP + lane),
                             __ldg(WcT + n * HP + (lane ^ 1)), __ldg(WcT + (n + 1) * HP + (lane ^ 1)));
    }
  }
  for (int e = tid; e < (L - 1) * nspc * HRS; e += blockDim.x) {
    mbar_init(&hand[e / HRS].full[e % HRS], 1);
    mbar_init(&hand[e / HRS].empty[e % HRS], 1);
    if (e % HRS < 2) { mbar_init(&hand[e / HRS].gfull[e % HRS], 1); mbar_init(&hand[e / HRS].gempty[e % HRS], 1); }
  }
  if (L > 1)
    for (int e = tid; e < nspc * 2; e += blockDim.x) {
      mbar_init(&hand3[e / 2].gfull[e % 2], 1);
      mbar_init(&hand3[e / 2].gempty[e % 2], 1);
    }
  fence_mbar_init();
  __syncthreads();
  pdl_trigger();                                         // the attention / dX kernel behind may be scheduled as SMs free up
  pdl_wait();                                            // everything above only read weights written two launches ago
  if (b >= a.B || k < 0) return;

  Handoff* hin = k < L - 1 ? &hand[k * nspc + si] : nullptr;        // from layer k+1
  Handoff* hout = k > 0 ? &hand[(k - 1) * nspc + si] : nullptr;     // to layer k-1
  if (a.whelper[w]) {                                    // layer 1's dx warp; also serves layer 0's chunk I/O
    float* s_st0 = reinterpret_cast<float*>(dsm + sm.l0 + si * bwd_region_bytes(BCH0, BNS0));   // same carve-up as wave_bwd_fast
    float* s_da0 = s_st0 + BNS0 * (BCH0 + 1) * ST;
    uint64_t* full0 = reinterpret_cast<uint64_t*>(s_da0 + 2 * BCH0 * G3 + 96);
    const L0BwdIo io{s_st0, s_da0, full0, a.st[0] + (int64_t)b * a.S[0] * ST, a.da[0] + (int64_t)b * a.S[0] * G3, a.S[0]};
    wave_bwd_dx_helper(a, i, UP ? nullptr : sWxT, &hand3[si], &hand[0 * nspc + si], io);
    return;
  }
  if (k == 0) {
    unsigned char* reg0 = dsm + sm.l0 + si * bwd_region_bytes(BCH0, BNS0);
    if (hin == nullptr) wave_bwd_fast<false, false, false, DBG>(a, 0, b, i, reg0, nullptr, nullptr);
    else wave_bwd_fast<true, false, true, DBG>(a, 0, b, i, reg0, hin, nullptr);
  } else if (k == 1) {
    unsigned char* reg1 = dsm + sm.l1 + si * bwd_region_bytes(BCH0, BNS0);
    if (hin == nullptr) wave_bwd_fast<false, true, false, DBG>(a, 1, b, i, reg1, nullptr, &hand3[si]);
    else wave_bwd_fast<true, true, false, DBG>(a, 1, b, i, reg1, hin, &hand3[si]);
  } else if (UP) {                                       // k == 2: the upper warp walks layers 2 .. L-1
    wave_bwd_upper<DBG>(a, b, i, dsm, sm, si, hout, sWxT, sWhcT);
  } else if (k == 2) {
    wave_bwd_layer<BCH2, BNSK, DBG>(a, k, b, i, dsm + sm.l2 + si * bwd_region_bytes(BCH2, BNSK), hin, hout,
                                    sWxT + (size_t)(k - 1) * 3 * 8 * HP, sWhcT + (size_t)(k - 2) * 8 * HP);
  } else {
    wave_bwd_layer<BCHK, BNSK, DBG>(a, k, b, i, dsm + sm.lk + ((k - 3) * nspc + si) * bwd_region_bytes(BCHK, BNSK), hin, hout,
                                    sWxT + (size_t)(k - 1) * 3 * 8 * HP, sWhcT + (size_t)(k - 2) * 8 * HP);
  }
}

template <bool DBG, bool UP>
static cudaError_t launch_wave_bwd_variant(const WaveBwdArgs& a, int grid, int threads, int smem, cudaStream_t st_) {
  cudaFuncSetAttribute(wave_bwd_kernel<DBG, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  return launch_pdl(wave_bwd_kernel<DBG, UP>, dim3(grid), dim3(threads), (size_t)smem, st_, a);   // prologue under the attention backward's tail
}

bool launch_wave_bwd(const Launch& L, const Dims& d, const PackLayout& pk, const float* pw, const float* const* st,
                     float* const* da, const float* dmemory, cudaStream_t st_) {
  if (d.L > WAVE_MAX_L || !wave_supported(d.L, d.P)) return false;
  const int nspc = d.L <= 5 ? 2 : 1;
  static const bool up_env_off = [] { const char* e = getenv("HPMN_NO_BWD_UPPER"); return e && e[0] == '1'; }();
  const bool up = BWD_UPPER && !up_env_off && d.L > 2;   // one warp for all layers >= 2 (8 warps per CTA, 255 registers)
  const WaveBwdSmem sm(d.L, nspc, up);
  if (sm.total > 227 * 1024) return false;   // opt-in maximum of dynamic shared memory per CTA on sm_100
  WaveBwdArgs a; memset(&a, 0, sizeof(a));
  a.pw = pw; a.dmemory = dmemory;
  a.B = d.B; a.L = d.L; a.H = d.H; a.nspc = nspc;
  for (int k = 0; k < d.L; ++k) { a.st[k] = st[k]; a.da[k] = da[k]; a.WhT[k] = pk.WhT[k]; a.WxT[k] = pk.WxT[k]; a.S[k] = d.S[k]; a.P[k] = d.P[k]; }
  const int grid = (d.B + nspc - 1) / nspc;
  // upper-warp mode plans three layer roles per sample (0, 1 and "2" = the upper warp) plus the helper
  const WarpPlan wp = plan_warps(up ? 3 : d.L, nspc, d.L > 1, "HPMN_WAVE_WB");
  for (int i = 0; i < 16; ++i) { a.wlayer[i] = wp.layer[i]; a.wsample[i] = wp.sample[i]; a.whelper[i] = wp.helper[i]; }
  bool debug = false;
  { static int once = 0; const char* e = getenv("HPMN_WAVE_DEBUG"); debug = e && e[0] == '1' && once++ == 3; }   // 4th call only
  if (debug) { if (up) launch_wave_bwd_variant<true, true>(a, grid, 32 * wp.n, sm.total, st_); else launch_wave_bwd_variant<true, false>(a, grid, 32 * wp.n, sm.total, st_); }
  else { if (up) launch_wave_bwd_variant<false, true>(a, grid, 32 * wp.n, sm.total, st_); else launch_wave_bwd_variant<false, false>(a, grid, 32 * wp.n, sm.total, st_); }
  { cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { if (getenv("HPMN_VERBOSE")) fprintf(stderr, "wave_bwd launch failed: %s (threads %d smem %d)\n", cudaGetErrorString(e), 32 * wp.n, sm.total); return false; } }
  ++*L.counter;
  return true;
}

}  // namespace hpmn
