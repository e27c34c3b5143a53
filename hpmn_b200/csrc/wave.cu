// wave.cu -- the hierarchical periodic memory as ONE persistent kernel per direction (forward here, backward
// below): all L layers of a sample run concurrently as a wavefront, layer k firing every prod(p[:k]) steps, so the
// critical path is the S_0 = Tpad steps of layer 0 instead of sum_k S_k (1024 vs 1984 at XLong).
// Replaces the per-layer tf.nn.dynamic_rnn calls of /root/reference/code/hpmn.py:113-131 (cell code/util.py:81-110);
// the equivalence of the layer-by-layer and the online formulation is the reference's own
// srnn.py:725-748 (`incremental_update`) and is pinned by tests/test_oracle.py.
//
// One CTA = NSPC samples x L warps.  Warp (k, s) owns the recurrence of layer k of sample s: recurrent weights in
// registers (96 per lane, FFMA2 pairs), state broadcast through shared memory, h|r|u|c rows leaving through TMA bulk
// stores.  Layer 0 streams its input projections (tcgen05 GEMM output) in with cp.async.bulk + mbarrier; a layer k >= 1
// receives the every-p-th hidden state of layer k-1 through a small shared-memory ring (producer / consumer counters)
// and applies its own input projection with W_x read from shared memory -- upper layers step at most every other
// layer-0 step, so they have the issue slots to spare.
#include <stdlib.h>

#include "common.cuh"

namespace hpmn {

constexpr int WCH = 8;        // steps per output chunk (one 4 KB bulk store)
constexpr int WIN = 16;       // steps per input chunk of layer 0
constexpr int WNS0 = 3;       // input ring stages of layer 0
constexpr int HRS = 8;        // hand-off ring slots between consecutive layers
constexpr int WAVE_MAX_L = 10;

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

struct Handoff {                   // layer k -> k+1 of one sample: ring of HRS rows, full / empty mbarrier per slot
  float ring[HRS][HP];
  uint64_t full[HRS], empty[HRS];
  // group-granular twin (HG rows per group, 2 groups) used on the layer-0 side of the layer-0 <-> helper interface:
  // the critical warp then touches a barrier once per HG hand-offs instead of twice per hand-off
  uint64_t gfull[2], gempty[2];
};
constexpr int HG = HRS / 2;

// dot(v[0..31], w) with v in shared memory in natural order and w as 16 (2q, 2q+1) register pairs
__device__ __forceinline__ float dotn(const float* v, const float2 (&w)[16], float init) {
  float2 a0 = make_float2(init, 0.f), a1 = make_float2(0.f, 0.f);
  const float4* s4 = reinterpret_cast<const float4*>(v);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 x = s4[q];
    a0 = ffma2(make_float2(x.x, x.y), w[2 * q], a0);
    a1 = ffma2(make_float2(x.z, x.w), w[2 * q + 1], a1);
  }
  return (a0.x + a1.x) + (a0.y + a1.y);
}

// same dot product as four chains of four FFMA2 (+ one more add level): for the dot products that sit alone on the
// dependent chain of a step (the candidate mat-vec, the first adjoint mat-vec) the chain latency is what counts
__device__ __forceinline__ float dotn4(const float* v, const float2 (&w)[16], float init) {
  float2 a0 = make_float2(init, 0.f), a1 = make_float2(0.f, 0.f), a2 = a1, a3 = a1;
  const float4* s4 = reinterpret_cast<const float4*>(v);
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    const float4 x = s4[q], y = s4[q + 1];
    a0 = ffma2(make_float2(x.x, x.y), w[2 * q], a0);
    a1 = ffma2(make_float2(x.z, x.w), w[2 * q + 1], a1);
    a2 = ffma2(make_float2(y.x, y.y), w[2 * q + 2], a2);
    a3 = ffma2(make_float2(y.z, y.w), w[2 * q + 3], a3);
  }
  return ((a0.x + a1.x) + (a2.x + a3.x)) + ((a0.y + a1.y) + (a2.y + a3.y));
}

// shared-memory load that keeps its program order among its kind (ptxas otherwise sorts broadcasts by consumer order)
__device__ __forceinline__ float4 lds128_ordered(const float4* p) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_u32(p)));
  return r;
}

struct Handoff3 {                  // projected input (fwd) / da row (bwd) of layer 1: 96 floats per slot
  float ring[HRS][G3];
  uint64_t full[HRS], empty[HRS];
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

// Warp -> role assignment.  Warps of a CTA land on SM sub-partition (warp id % 4) and the issue arbiter favours higher
// warp ids.  Roles are spread so that the estimated load per sub-partition is balanced and the critical layer-0 warps
// share theirs only with the lightest roles; inside a sub-partition the heaviest role gets the highest warp id.
struct WarpPlan { int n; signed char layer[16], sample[16], helper[16]; };
static WarpPlan plan_warps(int L, int nspc, bool with_helper) {
  struct Role { int layer, sample, helper; double w; };
  Role roles[16]; int nr = 0;
  for (int s = 0; s < nspc; ++s)
    for (int k = 0; k < L; ++k) {
      double w = k == 0 ? 2.0 : 1.0;                     // layer 0 is the critical path: keep its sub-partition quiet
      for (int q = 0; q < k; ++q) w *= 0.5;              // layer k runs ~2^-k of layer 0's steps ...
      if (k >= 1) w *= (with_helper && k == 1) ? 1.1 : 1.9;   // ... each ~2x as expensive unless a helper takes the matvec
      roles[nr++] = Role{k, s, 0, w};
    }
  if (with_helper && L > 1)
    for (int s = 0; s < nspc; ++s) roles[nr++] = Role{1, s, 1, 0.45};
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

// shared-memory plan (bytes)
struct WaveSmem {
  int wx, bx, hand, hand3, l0, lk, total;
  int r0, r1;                      // per-warp region sizes: layer 0 / layers >= 1
  __host__ __device__ WaveSmem(int L, int nspc) {
    int off = 0;
    wx = off; off += (L - 1) * 16 * 3 * HP * 8;            // float2 [q][g][j] per layer >= 1
    bx = off; off += (L - 1) * G3 * 4;
    off = (off + 127) & ~127;
    hand = off; off += (L - 1) * nspc * (int)sizeof(Handoff);
    off = (off + 127) & ~127;
    hand3 = off; off += (L > 1 ? nspc : 0) * (int)sizeof(Handoff3);   // helper warp of layer 1 -> layer 1
    off = (off + 127) & ~127;
    r0 = WNS0 * WIN * G3 * 4 + 2 * WIN * ST * 4 + 256 + 128;   // in ring | out ring (16-step chunks) | sh_rh + zero row | mbarriers
    r1 = 2 * WCH * ST * 4 + 256 + 128;                         // out ring | sh_rh + zero row | projected input row
    l0 = off; off += nspc * r0;
    lk = off; off += (L - 1) * nspc * r1;
    total = off;
  }
};

// The time loop of one (layer, sample) warp.  IS_L0 selects the input side at compile time: TMA-fed projection ring
// (layer 0) or hand-off ring + in-kernel projection (layers >= 1).  CH = steps per output chunk.  All bookkeeping is per
// chunk; the inner loop is unrolled so every shared-memory access has an immediate offset.
template <bool IS_L0, bool PRE, int CH>   // IS_L0 also selects the group-granular output protocol
__device__ __forceinline__ float wave_layer(const WaveArgs& a, int k, int b, int j, float* s_in, uint64_t* full,
                                            float* s_out, float* sh_rh, const float* zero_row, const float2* myWx,
                                            const float* myBx, Handoff* hin, Handoff3* hin3, Handoff* hout) {
  const int S = a.S[k], period = a.P[k];
  float2 wr[16], wu[16], wc[16];                         // recurrent weights, natural (2q, 2q+1) pairs
  {
    const float* Wh = a.pw + a.Wh[k];                    // [3][32 i][32 j]
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      wr[q] = make_float2(__ldg(Wh + (0 * HP + 2 * q) * HP + j), __ldg(Wh + (0 * HP + 2 * q + 1) * HP + j));
      wu[q] = make_float2(__ldg(Wh + (1 * HP + 2 * q) * HP + j), __ldg(Wh + (1 * HP + 2 * q + 1) * HP + j));
      wc[q] = make_float2(__ldg(Wh + (2 * HP + 2 * q) * HP + j), __ldg(Wh + (2 * HP + 2 * q + 1) * HP + j));
    }
  }
  const float* pp = a.proj0 + (int64_t)b * S * G3;       // layer 0 only
  float* so = a.st[k] + (int64_t)b * S * ST;
  const int nch = (S + CH - 1) / CH;
  if (IS_L0) {                                           // CH == WIN: one input chunk per output chunk
    if (j == 0) {
      for (int i = 0; i < WNS0; ++i) mbar_init(&full[i], 1);
      fence_mbar_init();
    }
    __syncwarp();
    if (j == 0)
      for (int c = 0; c < WNS0 && c < nch; ++c) {
        const int len = min(CH, S - c * CH);
        mbar_expect_tx(&full[c], (uint32_t)len * G3 * 4);
        bulk_g2s(s_in + c * CH * G3, pp + (int64_t)c * CH * G3, (uint32_t)len * G3 * 4, &full[c]);
      }
  }
  __syncwarp();

  float h = 0.f;                                         // zero_state, code/rnn.py:588
  const float* hprev = zero_row;                         // broadcast source of h_{t-1}: the previous output row
  long long w_in = 0, w_out = 0, w_tma = 0;              // debug: cycles blocked on input / output / TMA barriers
  const bool dbg = a.debug != 0 && blockIdx.x == 0;
  const long long t_start = clock64();
  unsigned fired = 0, s_glob = 0;                        // hand-offs produced / consumed so far
  const unsigned n_out = hout != nullptr ? (unsigned)(S / period) : 0u;
  int to_fire = period;

  // layers >= 1: input projection of hand-off `idx` (hidden state of layer k-1 at its step (idx+1)*p_{k-1} - 1,
  // hpmn.py:124-128).  It does not depend on this layer's state, so it is issued one step ahead and overlaps the
  // latency bubbles of the recurrent chain.
  auto project = [&](unsigned idx, float& ar, float& au, float& ac) {
    const int slot_in = idx & (HRS - 1);
    if (PRE) {                                           // layer 1: the helper warp already applied W_x
      mbar_wait_t(&hin3->full[slot_in], (idx / HRS) & 1u, w_in, dbg);
      ar = hin3->ring[slot_in][j]; au = hin3->ring[slot_in][HP + j]; ac = hin3->ring[slot_in][2 * HP + j];
      __syncwarp();
      if (j == 0) mbar_arrive(&hin3->empty[slot_in]);
      return;
    }
    mbar_wait_t(&hin->full[slot_in], (idx / HRS) & 1u, w_in, dbg);   // hardware-suspended wait, no polling
    const float4* x4 = reinterpret_cast<const float4*>(hin->ring[slot_in]);
    float2 p0 = make_float2(myBx[j], 0.f), p1 = make_float2(myBx[HP + j], 0.f), p2 = make_float2(myBx[2 * HP + j], 0.f);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 x = x4[q4];
      const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
      const float2* wq = myWx + (2 * q4) * 3 * HP + j;
      p0 = ffma2(xa, wq[0], p0); p1 = ffma2(xa, wq[HP], p1); p2 = ffma2(xa, wq[2 * HP], p2);
      p0 = ffma2(xb, wq[3 * HP], p0); p1 = ffma2(xb, wq[4 * HP], p1); p2 = ffma2(xb, wq[5 * HP], p2);
    }
    ar = p0.x + p0.y; au = p1.x + p1.y; ac = p2.x + p2.y;
    __syncwarp();
    if (j == 0) mbar_arrive(&hin->empty[slot_in]);       // slot free again
  };
  float nar = 0.f, nau = 0.f, nac = 0.f;                 // projections of the NEXT step (layers >= 1)
  if (!IS_L0) project(0, nar, nau, nac);

  auto step = [&](const float* ib, float* orow) {
    float ar, au, ac;
    if (IS_L0) {
      ar = ib[j]; au = ib[HP + j]; ac = ib[2 * HP + j];
    } else {
      ar = nar; au = nau; ac = nac;
      ++s_glob;
      if (s_glob < (unsigned)S) project(s_glob, nar, nau, nac);
    }
    const float r = sigmoid_f(dotn(hprev, wr, ar));      // util.py:95-96
    const float u = sigmoid_f(dotn(hprev, wu, au));
    sh_rh[j] = r * h;                                    // util.py:98
    __syncwarp();
    const float c = tanh_f(dotn4(sh_rh, wc, ac));        // util.py:107
    h = fmaf(u, h - c, c);                               // util.py:109
    orow[j] = h; orow[HP + j] = r; orow[2 * HP + j] = u; orow[3 * HP + j] = c;
    hprev = orow;
    if (hout != nullptr && --to_fire == 0) {             // this step feeds layer k+1
      to_fire = period;
      const int slot_out = fired & (HRS - 1);
      if (IS_L0) {                                       // -> helper warp, one barrier round trip per HG rows
        const int g = (fired / HG) & 1;
        if ((fired & (HG - 1)) == 0 && fired >= HRS) mbar_wait_t(&hout->gempty[g], (fired / HRS - 1) & 1u, w_out, dbg);
        hout->ring[slot_out][j] = h;
        ++fired;
        if ((fired & (HG - 1)) == 0 || fired == n_out) {
          __syncwarp();
          if (j == 0) mbar_arrive(&hout->gfull[g]);
        }
      } else {
        if (fired >= HRS) mbar_wait_t(&hout->empty[slot_out], (fired / HRS - 1) & 1u, w_out, dbg);
        hout->ring[slot_out][j] = h;
        ++fired;
        __syncwarp();
        if (j == 0) mbar_arrive(&hout->full[slot_out]);  // release: the row is visible to the waiting layer
      }
    }
    __syncwarp();                                        // orow (next step's broadcast source) and sh_rh settled
  };

  for (int c = 0; c < nch; ++c) {
    const int len = min(CH, S - c * CH);
    float* ob = s_out + (c & 1) * CH * ST;
    const float* ib = nullptr;
    if (IS_L0) {
      const int stage = c % WNS0;
      mbar_wait_t(&full[stage], (uint32_t)(c / WNS0) & 1u, w_tma, dbg);
      ib = s_in + stage * CH * G3;
    }
    if (c >= 2) {                                        // the bulk store that read this buffer two chunks ago
      if (j == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    if (len == CH) {
      if constexpr (IS_L0) {
#pragma unroll 4
        for (int t = 0; t < CH; ++t) step(ib + t * G3, ob + t * ST);
      } else {
#pragma unroll 2                                         // deeper unrolling spills at the 168-register cap
        for (int t = 0; t < CH; ++t) step(ib + t * G3, ob + t * ST);
      }
    } else {
      for (int t = 0; t < len; ++t) step(ib + t * G3, ob + t * ST);
    }
    fence_proxy_async();                                 // generic-proxy writes of ob -> visible to the bulk store
    __syncwarp();
    if (j == 0) {
      bulk_s2g(so + (int64_t)c * CH * ST, ob, (uint32_t)len * ST * 4);
      bulk_commit();
      if (IS_L0) {                                       // refill the input stage every lane has finished reading
        const int stage = c % WNS0, cn = c + WNS0;
        if (cn < nch) {
          const int ln = min(CH, S - cn * CH);
          mbar_expect_tx(&full[stage], (uint32_t)ln * G3 * 4);
          bulk_g2s(s_in + stage * CH * G3, pp + (int64_t)cn * CH * G3, (uint32_t)ln * G3 * 4, &full[stage]);
        }
      }
    }
  }
  if (j == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && j == 0 && b == 0)
    printf("wave_fwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma %lld\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, w_in, w_out, w_tma);
  return h;
}

// ---------------------------------------------------------------------------------------------------------------------
// Layer 0, forward: the critical path of the kernel (S_0 dependent steps), written for the latency of ONE step.
// What the ncu source view of the generic loop showed (profiles/r1_v8_wave_step_timeline.md): of ~485 cycles per step
// only ~270 were the dependent chain; ~90 went to the per-step hand-off branch (BSSY/BSYNC + refetch) and to address
// arithmetic rematerialised under register pressure, ~60 to broadcast loads that ptxas serialised behind their consumers.
// Here, per 8-step block: no branch and no address arithmetic inside the block (the broadcast buffers sh_h / sh_rh sit at
// fixed addresses, input / output rows are immediates off two per-block pointers), hand-offs to the helper warp at
// compile-time steps (period 2 => odd t, slots 4g..4g+3 of group g = block & 1, one barrier round trip per block), the
// log2(e) factors of sigmoid / tanh folded into the register-resident weights, and the blend h' = u h + (1-u) c folded
// into the tail of tanh (one FFMA after the reciprocal instead of three dependent ops).
// FIRE2 = false: no layer above (L == 1).  Layer-0 periods other than 2 take the generic wave_layer<true,...>.
// ---------------------------------------------------------------------------------------------------------------------
constexpr float kNegLog2e = -1.4426950408889634f;        // sigmoid(x) = 1 / (1 + 2^(-x log2 e))
constexpr float kTwoLog2e = 2.8853900817779268f;         // tanh(x) = 1 - 2 / (1 + 2^(2 x log2 e))
constexpr float kInvTwoLog2e = 0.34657359027997264f;
constexpr int BLK = 8;                                   // steps per unrolled block of layer 0

template <bool FIRE2>
__device__ __forceinline__ float wave_layer0(const WaveArgs& a, int b, int j, float* s_in, uint64_t* full, float* s_out,
                                             float* sh_rh, float* sh_h, Handoff* hout) {
  constexpr int CH = WIN;
  const int S = a.S[0];
  float2 wr[16], wu[16], wc[16];                         // recurrent weights x log2(e) factors, (2q, 2q+1) pairs
  {
    const float* Wh = a.pw + a.Wh[0];                    // [3][32 i][32 j]
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      wr[q] = make_float2(kNegLog2e * __ldg(Wh + (0 * HP + 2 * q) * HP + j), kNegLog2e * __ldg(Wh + (0 * HP + 2 * q + 1) * HP + j));
      wu[q] = make_float2(kNegLog2e * __ldg(Wh + (1 * HP + 2 * q) * HP + j), kNegLog2e * __ldg(Wh + (1 * HP + 2 * q + 1) * HP + j));
      wc[q] = make_float2(kTwoLog2e * __ldg(Wh + (2 * HP + 2 * q) * HP + j), kTwoLog2e * __ldg(Wh + (2 * HP + 2 * q + 1) * HP + j));
    }
  }
  const float* pp = a.proj0 + (int64_t)b * S * G3;
  float* so = a.st[0] + (int64_t)b * S * ST;
  const int nch = (S + CH - 1) / CH;
  if (j == 0) {
    for (int i = 0; i < WNS0; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (j == 0)
    for (int c = 0; c < WNS0 && c < nch; ++c) {
      const int len = min(CH, S - c * CH);
      mbar_expect_tx(&full[c], (uint32_t)len * G3 * 4);
      bulk_g2s(s_in + c * CH * G3, pp + (int64_t)c * CH * G3, (uint32_t)len * G3 * 4, &full[c]);
    }
  __syncwarp();

  long long w_out = 0, w_tma = 0, w_steps = 0;           // debug: cycles blocked on the helper / on TMA, cycles inside the step blocks
  const bool dbg = a.debug != 0 && blockIdx.x == 0;
  const long long t_start = clock64();
  float h = 0.f;                                         // zero_state, code/rnn.py:588 (sh_h starts as the zero row)
  const float4* hb4 = reinterpret_cast<const float4*>(sh_h);
  const float4* rb4 = reinterpret_cast<const float4*>(sh_rh);

  // one GRU step (util.py:81-110 minus :108): in = this lane's column of the projected input row, out = of the state row
  auto step = [&](const float* in, float* out) {
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = hb4[q];           // h_{t-1} broadcast
    const float ar = in[0] * kNegLog2e, au = in[HP] * kNegLog2e, ac = in[2 * HP] * kTwoLog2e;
    float2 r0 = make_float2(ar, 0.f), r1 = make_float2(0.f, 0.f), r2 = r1, r3 = r1;
    float2 u0 = make_float2(au, 0.f), u1 = r1, u2 = r1, u3 = r1;
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      const float2 x0 = make_float2(v[q].x, v[q].y), x1 = make_float2(v[q].z, v[q].w);
      const float2 x2 = make_float2(v[q + 1].x, v[q + 1].y), x3 = make_float2(v[q + 1].z, v[q + 1].w);
      r0 = ffma2(x0, wr[2 * q], r0); u0 = ffma2(x0, wu[2 * q], u0);
      r1 = ffma2(x1, wr[2 * q + 1], r1); u1 = ffma2(x1, wu[2 * q + 1], u1);
      r2 = ffma2(x2, wr[2 * q + 2], r2); u2 = ffma2(x2, wu[2 * q + 2], u2);
      r3 = ffma2(x3, wr[2 * q + 3], r3); u3 = ffma2(x3, wu[2 * q + 3], u3);
    }
    const float sr = ((r0.x + r1.x) + (r2.x + r3.x)) + ((r0.y + r1.y) + (r2.y + r3.y));
    const float su = ((u0.x + u1.x) + (u2.x + u3.x)) + ((u0.y + u1.y) + (u2.y + u3.y));
    const float r = rcp_ftz(1.0f + ex2_ftz(sr));         // util.py:95-96
    sh_rh[j] = r * h;                                    // util.py:98
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = rb4[q];           // r o h broadcast
    const float u = rcp_ftz(1.0f + ex2_ftz(su));
    float2 c0 = make_float2(ac, 0.f), c1 = make_float2(0.f, 0.f), c2 = c1, c3 = c1;
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      c0 = ffma2(make_float2(v[q].x, v[q].y), wc[2 * q], c0);
      c1 = ffma2(make_float2(v[q].z, v[q].w), wc[2 * q + 1], c1);
      c2 = ffma2(make_float2(v[q + 1].x, v[q + 1].y), wc[2 * q + 2], c2);
      c3 = ffma2(make_float2(v[q + 1].z, v[q + 1].w), wc[2 * q + 3], c3);
    }
    const float omu = 1.0f - u, uh = u * h;
    const float x2l = ((c0.x + c1.x) + (c2.x + c3.x)) + ((c0.y + c1.y) + (c2.y + c3.y));   // 2 log2(e) x
    const float qv = rcp_ftz(1.0f + ex2_ftz(x2l));       // tanh(x) = 1 - 2 qv                  util.py:107
    // |x| < 0.15: odd Taylor series (the closed form cancels there); off the dependent chain except for the select
    const float x = x2l * kInvTwoLog2e, xx = x * x;
    const float small = x * fmaf(xx, fmaf(xx, fmaf(xx, -17.0f / 315.0f, 2.0f / 15.0f), -1.0f / 3.0f), 1.0f);
    const bool tiny = fabsf(x) < 0.15f;
    const float hs = fmaf(omu, small, uh);
    const float hbig = fmaf(-2.0f * omu, qv, uh + omu);  // u h + (1-u)(1 - 2 qv)               util.py:109
    h = tiny ? hs : hbig;
    const float c = tiny ? small : fmaf(-2.0f, qv, 1.0f);
    sh_h[j] = h;
    out[0] = h; out[HP] = r; out[2 * HP] = u; out[3 * HP] = c;
  };

  for (int c = 0; c < nch; ++c) {
    const int len = min(CH, S - c * CH);
    const int stage = c % WNS0;
    float* ob = s_out + (c & 1) * CH * ST;
    mbar_wait_t(&full[stage], (uint32_t)(c / WNS0) & 1u, w_tma, dbg);
    const float* ib = s_in + stage * CH * G3;
    if (c >= 2) {                                        // the bulk store that read this buffer two chunks ago
      if (j == 0) bulk_wait_read<1>();
      __syncwarp();
    }
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
#pragma unroll
      for (int t = 0; t < BLK; ++t) {
        step(in + t * G3, out + t * ST);
        if (FIRE2 && (t & 1)) rg[(t >> 1) * HP] = h;     // every 2nd state feeds layer 1 (hpmn.py:124-128)
        __syncwarp();                                    // sh_h (next step's broadcast) and sh_rh settled
      }
      if (dbg) w_steps += clock64() - tb0;
      if (FIRE2 && j == 0) mbar_arrive(&hout->gfull[g]);
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
      if (FIRE2 && len - t0 >= 2 && j == 0) mbar_arrive(&hout->gfull[g]);
    }
    fence_proxy_async();                                 // generic-proxy writes of ob -> visible to the bulk store
    __syncwarp();
    if (j == 0) {
      bulk_s2g(so + (int64_t)c * CH * ST, ob, (uint32_t)len * ST * 4);
      bulk_commit();
      const int cn = c + WNS0;                           // refill the input stage every lane has finished reading
      if (cn < nch) {
        const int ln = min(CH, S - cn * CH);
        mbar_expect_tx(&full[stage], (uint32_t)ln * G3 * 4);
        bulk_g2s(s_in + stage * CH * G3, pp + (int64_t)cn * CH * G3, (uint32_t)ln * G3 * 4, &full[stage]);
      }
    }
  }
  if (j == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && j == 0 && b == 0)
    printf("wave_fwd layer 0: steps %d total %lld cyc (%lld/step)  wait_in 0  wait_out %lld  wait_tma %lld  in-blocks %lld\n", S,
           clock64() - t_start, (clock64() - t_start) / S, w_out, w_tma, w_steps);
  return h;
}

// Helper warp of layer 1 (forward): applies W_x^(1) to every hand-off of layer 0 so that layer 1's recurrent warp has the
// same per-step cost as layer 0's while running at half its rate.
__device__ __forceinline__ void wave_proj_helper(int n, int j, const float2* myWx, const float* myBx, Handoff* hin, Handoff3* hout3) {
  // W_x of layer 1 in registers (this warp has no other state): no shared-memory traffic competing with layer 0's broadcasts
  float2 w0[16], w1[16], w2[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) { w0[q] = myWx[(q * 3 + 0) * HP + j]; w1[q] = myWx[(q * 3 + 1) * HP + j]; w2[q] = myWx[(q * 3 + 2) * HP + j]; }
  const float b0 = myBx[j], b1 = myBx[HP + j], b2 = myBx[2 * HP + j];
  for (unsigned idx = 0; idx < (unsigned)n; ++idx) {
    const int slot = idx & (HRS - 1), g = (idx / HG) & 1;
    if ((idx & (HG - 1)) == 0) mbar_wait(&hin->gfull[g], (idx / HRS) & 1u);
    const float4* x4 = reinterpret_cast<const float4*>(hin->ring[slot]);
    float2 p0 = make_float2(b0, 0.f), p1 = make_float2(b1, 0.f), p2 = make_float2(b2, 0.f);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const float4 x = x4[q4];
      const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
      p0 = ffma2(xa, w0[2 * q4], p0); p1 = ffma2(xa, w1[2 * q4], p1); p2 = ffma2(xa, w2[2 * q4], p2);
      p0 = ffma2(xb, w0[2 * q4 + 1], p0); p1 = ffma2(xb, w1[2 * q4 + 1], p1); p2 = ffma2(xb, w2[2 * q4 + 1], p2);
    }
    __syncwarp();
    if (j == 0 && ((idx & (HG - 1)) == HG - 1 || idx + 1 == (unsigned)n)) mbar_arrive(&hin->gempty[g]);   // group free again
    if (idx >= HRS) mbar_wait(&hout3->empty[slot], (idx / HRS - 1) & 1u);
    hout3->ring[slot][j] = p0.x + p0.y;
    hout3->ring[slot][HP + j] = p1.x + p1.y;
    hout3->ring[slot][2 * HP + j] = p2.x + p2.y;
    __syncwarp();
    if (j == 0) mbar_arrive(&hout3->full[slot]);
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
  Handoff3* hand3 = reinterpret_cast<Handoff3*>(dsm + sm.hand3);
  float2* sWx = reinterpret_cast<float2*>(dsm + sm.wx);
  float* sBx = reinterpret_cast<float*>(dsm + sm.bx);
  Handoff* hand = reinterpret_cast<Handoff*>(dsm + sm.hand);

  // ---- CTA setup: W_x / b_x of layers >= 1 into shared memory as (2q, 2q+1) pairs; hand-off barriers ----
  for (int e = tid; e < (L - 1) * 16 * 3 * HP; e += blockDim.x) {
    const int kk = 1 + e / (16 * 3 * HP), r = e % (16 * 3 * HP);
    const int q = r / (3 * HP), g = (r / HP) % 3, jj = r % HP;
    const float* Wx = a.pw + a.Wx[kk];                   // [32][96]
    sWx[e] = make_float2(__ldg(Wx + (2 * q) * G3 + g * HP + jj), __ldg(Wx + (2 * q + 1) * G3 + g * HP + jj));
  }
  for (int e = tid; e < (L - 1) * G3; e += blockDim.x) sBx[e] = __ldg(a.pw + a.bx[1 + e / G3] + e % G3);
  for (int e = tid; e < (L - 1) * nspc * HRS; e += blockDim.x) {
    mbar_init(&hand[e / HRS].full[e % HRS], 1);
    mbar_init(&hand[e / HRS].empty[e % HRS], 1);
    if (e % HRS < 2) { mbar_init(&hand[e / HRS].gfull[e % HRS], 1); mbar_init(&hand[e / HRS].gempty[e % HRS], 1); }
  }
  if (L > 1)
    for (int e = tid; e < nspc * HRS; e += blockDim.x) {
      mbar_init(&hand3[e / HRS].full[e % HRS], 1);
      mbar_init(&hand3[e / HRS].empty[e % HRS], 1);
    }
  fence_mbar_init();
  __syncthreads();
  if (b >= a.B) return;                                  // ragged last CTA: the whole warp leaves together

  if (helper) {                                          // layer 1's projection warp
    wave_proj_helper(a.S[1], j, sWx, sBx, &hand[0 * nspc + si], &hand3[si]);
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
    float* zero_row = sh_rh + 32;
    uint64_t* full = reinterpret_cast<uint64_t*>(zero_row + 32);
    zero_row[j] = 0.f;
    if (hout == nullptr) h = wave_layer0<false>(a, b, j, s_in, full, s_out, sh_rh, zero_row, nullptr);
    else if (a.P[0] == 2) h = wave_layer0<true>(a, b, j, s_in, full, s_out, sh_rh, zero_row, hout);
    else h = wave_layer<true, false, WIN>(a, k, b, j, s_in, full, s_out, sh_rh, zero_row, nullptr, nullptr, nullptr, nullptr, hout);
  } else {
    float* s_out = reinterpret_cast<float*>(reg);
    float* sh_rh = s_out + 2 * WCH * ST;
    float* zero_row = sh_rh + 32;
    zero_row[j] = 0.f;
    if (k == 1)
      h = wave_layer<false, true, WCH>(a, k, b, j, nullptr, nullptr, s_out, sh_rh, zero_row, nullptr, nullptr, nullptr, &hand3[si], hout);
    else
      h = wave_layer<false, false, WCH>(a, k, b, j, nullptr, nullptr, s_out, sh_rh, zero_row, sWx + (size_t)(k - 1) * 16 * 3 * HP,
                                        sBx + (k - 1) * G3, hin, nullptr, hout);
  }
  if (j < a.H) a.memory[((int64_t)b * L + k) * a.H + j] = h;   // final state -> memory slot k, hpmn.py:121
}

bool launch_wave_fwd(const Launch& L, const Dims& d, const PackLayout& pk, const float* proj0, const float* pw,
                     float* const* st, float* memory, cudaStream_t st_) {
  if (d.L > WAVE_MAX_L) return false;
  const int nspc = d.L <= 5 ? 2 : 1;                     // register budget: 32 x ~180 per warp
  const WaveSmem sm(d.L, nspc);
  if (sm.total > 220 * 1024) return false;
  WaveArgs a; memset(&a, 0, sizeof(a));
  a.proj0 = proj0; a.pw = pw; a.memory = memory;
  a.B = d.B; a.L = d.L; a.H = d.H; a.nspc = nspc;
  { static int once = 0; const char* e = getenv("HPMN_WAVE_DEBUG"); a.debug = (e && e[0] == '1' && once++ == 3) ? 1 : 0; }   // 4th call only
  for (int k = 0; k < d.L; ++k) { a.st[k] = st[k]; a.Wh[k] = pk.Wh[k]; a.Wx[k] = pk.Wx[k]; a.bx[k] = pk.bx[k]; a.S[k] = d.S[k]; a.P[k] = d.P[k]; }
  cudaFuncSetAttribute(wave_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total);
  const int grid = (d.B + nspc - 1) / nspc;
  const WarpPlan wp = plan_warps(d.L, nspc, true);
  for (int i = 0; i < 16; ++i) { a.wlayer[i] = wp.layer[i]; a.wsample[i] = wp.sample[i]; a.whelper[i] = wp.helper[i]; }
  wave_fwd_kernel<<<grid, 32 * wp.n, sm.total, st_>>>(a);
  { cudaError_t e = cudaGetLastError();                  // resources: the caller falls back to the per-layer kernels
    if (e != cudaSuccess) { if (getenv("HPMN_VERBOSE")) fprintf(stderr, "wave_fwd launch failed: %s (threads %d smem %d)\n", cudaGetErrorString(e), 32 * d.L * nspc, sm.total); return false; } }
  ++*L.counter;
  return true;
}

// =====================================================================================================
// Backward wavefront: every layer walks its steps in reverse concurrently.  Layer k needs, at each of its firing
// steps s = (j+1)*p_k - 1, the gradient dx_{k+1}[j] = da_{k+1}[j] * W_x^{(k+1)T} of the layer above; the warp of
// layer k+1 computes it right after its step j (da_r|da_u|da_c are already broadcast in shared memory for the
// recurrent dot products; W_x^T comes from shared memory) and hands it down through an mbarrier ring.  Saved state
// rows stream in with cp.async.bulk (chunk = CHB steps needs rows s0-1 .. s0+len-1), da rows leave with bulk stores
// for the weight-gradient kernel and, for layer 0, the dX GEMM that feeds the embedding scatter.
// =====================================================================================================
constexpr int BCH0 = 8, BNS0 = 3;   // layer 0: steps per chunk, state ring stages
constexpr int BCHK = 4, BNSK = 2;   // layers >= 1 (not on the critical path): smaller rings

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
  return ns * (ch + 1) * ST * 4 + 2 * ch * G3 * 4 + 3 * 128 + 128;   // state ring | da ring | sh_c,sh_r,sh_u | mbarriers
}

struct WaveBwdSmem {
  int wxt, hand, hand3, l0, l1, lk, total;
  __host__ __device__ WaveBwdSmem(int L, int nspc) {
    int off = 0;
    wxt = off; off += (L - 1) * 48 * HP * 8;               // float2 [q (48 pairs of n)][i] per layer >= 1
    off = (off + 127) & ~127;
    hand = off; off += (L - 1) * nspc * (int)sizeof(Handoff);
    off = (off + 127) & ~127;
    hand3 = off; off += (L > 1 ? nspc : 0) * (int)sizeof(Handoff3);   // layer 1 -> its dx helper
    off = (off + 127) & ~127;
    l0 = off; off += nspc * bwd_region_bytes(BCH0, BNS0);
    l1 = off; off += (L > 1 ? nspc : 0) * bwd_region_bytes(BCH0, BNS0);   // layer 1 runs the same loop as layer 0
    lk = off; off += (L > 2 ? L - 2 : 0) * nspc * bwd_region_bytes(BCHK, BNSK);
    total = off;
  }
};

// Layers 0 and 1 run wave_bwd_fast (layer 1 with a helper warp that applies W_x^T and hands dx down to layer 0 in groups)
// when both fired every 2nd step -- every configuration of the reference (hpmn.py:576-662).  Otherwise every layer of an
// L > 1 stack runs the generic wave_bwd_layer.
__host__ __device__ __forceinline__ bool bwd_fast01(int L, const int* P) { return L > 1 && P[0] == 2 && (L == 2 || P[1] == 2); }
// ring-slot offset that aligns the producer's groups of HG rows with the consumer's chunks when its top chunk is ragged
__device__ __forceinline__ unsigned group_vofs(int S_consumer) { return (unsigned)((HG - (S_consumer % BLK) / 2) & (HG - 1)); }

// ---------------------------------------------------------------------------------------------------------------------
// Layer 0, backward: the critical path (S_0 dependent steps), written for the latency of one step like wave_layer0.
// Per step the dependent chain is  dh -> da_c (one FMUL) -> broadcast -> (da_c Wc^T) -> da_r (one FMUL) -> broadcast ->
// (da_r Wr^T) -> dh';  every factor that does not depend on dh (1-u, 1-c^2, (h_prev-c) u (1-u), r (1-r) h_prev) is
// formed from the saved state before dh arrives, da_u is broadcast together with da_c so that its mat-vec fills the
// bubbles of the chain, and the gradient handed down by layer 1 is added at the END of the previous step's chain.
// One chunk = BCH0 = 8 steps = one hand-off group (4 rows of the layer above, one barrier round trip); no branch or address
// arithmetic inside a chunk.  FIRE2 = false: no layer above.  The same loop serves layer 1 (OUT_DA): it is only half as
// often on duty as layer 0, but with its own dx = da W_x^T mat-vec (96 x 32, weights from shared memory) it was the
// bottleneck of the whole backward kernel (1450 cycles per step = 725 per layer-0 step); it now hands its da row to a
// helper warp that holds W_x^T in registers (wave_bwd_dx_helper), mirroring the forward kernel's projection helper.
// ---------------------------------------------------------------------------------------------------------------------
template <bool FIRE2, bool OUT_DA, bool DBG>
__device__ __forceinline__ void wave_bwd_fast(const WaveBwdArgs& a, int k, int b, int i, unsigned char* reg, Handoff* hin,
                                              Handoff3* hda) {
  constexpr int CH = BCH0, NS = BNS0;
  static_assert(CH == BLK && CH == 2 * HG, "one chunk = one hand-off group of the layer above");
  long long w_in = 0, w_out = 0, w_tma = 0;
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

  float2 wrT[16], wuT[16], wcT[16];                      // lane i: W[Din+i][g*H + j] over j, natural (2q, 2q+1) pairs
  {
    const float* WhT = a.pw + a.WhT[k];                  // [3][32 j][32 i]
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      wrT[q] = make_float2(__ldg(WhT + (0 * HP + 2 * q) * HP + i), __ldg(WhT + (0 * HP + 2 * q + 1) * HP + i));
      wuT[q] = make_float2(__ldg(WhT + (1 * HP + 2 * q) * HP + i), __ldg(WhT + (1 * HP + 2 * q + 1) * HP + i));
      wcT[q] = make_float2(__ldg(WhT + (2 * HP + 2 * q) * HP + i), __ldg(WhT + (2 * HP + 2 * q + 1) * HP + i));
    }
  }
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
  const float4* c4 = reinterpret_cast<const float4*>(sh_c);
  const float4* r4 = reinterpret_cast<const float4*>(sh_r);
  const float4* u4 = reinterpret_cast<const float4*>(sh_u);

  // one reverse step.  row = this lane's column of buffer row t (h_{s-1}; rows t+1 hold r|u|c of step s), orow = of the da row,
  // add = gradient arriving from layer 1 for the step BELOW this one (0 if that step did not fire)
  auto step = [&](const float* row, float* orow, bool first_step, float add) {
    const float hp = first_step ? 0.f : row[0];          // zero state before step 0
    const float r = row[ST + HP], u = row[ST + 2 * HP], c = row[ST + 3 * HP];
    const float omu = 1.f - u;
    const float gc = omu * fmaf(-c, c, 1.f);             // d a_c / d h'
    const float gu = (hp - c) * u * omu;                 // d a_u / d h'
    const float gr = r * (1.f - r) * hp;                 // d a_r / d (r o h_prev)
    const float dac = dh * gc, dau = dh * gu;
    sh_c[i] = dac;
    sh_u[i] = dau;
    __syncwarp();
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = c4[q];
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      a0 = ffma2(make_float2(v[q].x, v[q].y), wcT[2 * q], a0);
      a1 = ffma2(make_float2(v[q].z, v[q].w), wcT[2 * q + 1], a1);
      a2 = ffma2(make_float2(v[q + 1].x, v[q + 1].y), wcT[2 * q + 2], a2);
      a3 = ffma2(make_float2(v[q + 1].z, v[q + 1].w), wcT[2 * q + 3], a3);
    }
    const float drh = ((a0.x + a1.x) + (a2.x + a3.x)) + ((a0.y + a1.y) + (a2.y + a3.y));   // (da_c Wc^T)[Din + i]
    const float dar = drh * gr;
    sh_r[i] = dar;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = lds128_ordered(r4 + q);   // the chain's broadcast goes first ...
    float4 w[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) w[q] = lds128_ordered(u4 + q);   // ... da_u's mat-vec fills its latency
    float2 b0 = make_float2(fmaf(dh, u, add), 0.f), b1 = make_float2(0.f, 0.f), b2 = b1, b3 = b1;
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      b0 = ffma2(make_float2(w[q].x, w[q].y), wuT[2 * q], b0);
      b1 = ffma2(make_float2(w[q].z, w[q].w), wuT[2 * q + 1], b1);
      b2 = ffma2(make_float2(w[q + 1].x, w[q + 1].y), wuT[2 * q + 2], b2);
      b3 = ffma2(make_float2(w[q + 1].z, w[q + 1].w), wuT[2 * q + 3], b3);
    }
    const float part = fmaf(drh, r, ((b0.x + b1.x) + (b2.x + b3.x)) + ((b0.y + b1.y) + (b2.y + b3.y)));
    a0 = make_float2(0.f, 0.f); a1 = a0; a2 = a0; a3 = a0;
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      a0 = ffma2(make_float2(v[q].x, v[q].y), wrT[2 * q], a0);
      a1 = ffma2(make_float2(v[q].z, v[q].w), wrT[2 * q + 1], a1);
      a2 = ffma2(make_float2(v[q + 1].x, v[q + 1].y), wrT[2 * q + 2], a2);
      a3 = ffma2(make_float2(v[q + 1].z, v[q + 1].w), wrT[2 * q + 3], a3);
    }
    dh = (((a0.x + a1.x) + (a2.x + a3.x)) + ((a0.y + a1.y) + (a2.y + a3.y))) + part;   // gradient wrt h_{s-1} (+ layer 1's share)
    orow[0] = dar; orow[HP] = dau; orow[2 * HP] = dac;
    if (OUT_DA) {                                        // every step of layer k fed layer k-1's firing step: da row -> helper
      const int slot = sent & (HRS - 1);
      if (sent >= HRS) mbar_wait_t(&hda->empty[slot], (sent / HRS - 1) & 1u, w_out, dbg);
      hda->ring[slot][i] = dar; hda->ring[slot][HP + i] = dau; hda->ring[slot][2 * HP + i] = dac;
      ++sent;
      __syncwarp();
      if (i == 0) mbar_arrive(&hda->full[slot]);
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
    if (it >= 2) {
      if (i == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    const float* ib = s_st + stage * (CH + 1) * ST + i;
    const float* rg = nullptr;
    if (FIRE2) {                                         // the 4 (top chunk: len/2) rows layer 1 hands down for this chunk
      rg = &hin->ring[g * HG][i];
      if (it == 0) {                                     // later groups are awaited at the end of the previous iteration
        mbar_wait_t(&hin->gfull[0], 0u, w_in, dbg);
        dh += rg[(HG - 1 - ((len - 1) >> 1)) * HP];      // the last step of the sequence fired
      }
    }
    // `add` of step t belongs to step t-1: row n = HG-1 - ((t-1)>>1) of the group if t-1 is odd (a firing step); the add for
    // this chunk's step 0 is the first row of the NEXT group and is applied at the top of the next iteration.
    if (len == CH && s0 > 0) {
#pragma unroll
      for (int t = CH - 1; t >= 0; --t)
        step(ib + t * ST, ob + t * G3, false, (FIRE2 && t > 0 && ((t - 1) & 1)) ? rg[(HG - 1 - ((t - 1) >> 1)) * HP] : 0.f);
    } else {
      for (int t = len - 1; t >= 0; --t)
        step(ib + t * ST, ob + t * G3, s0 + t == 0, (FIRE2 && t > 0 && ((t - 1) & 1)) ? rg[(HG - 1 - ((t - 1) >> 1)) * HP] : 0.f);
    }
    if (FIRE2) {
      if (i == 0) mbar_arrive(&hin->gempty[g]);          // every lane's reads of the group precede the step's last __syncwarp
      if (it + 1 < nch) {                                // step s0-1 (odd) fired: its share is row 0 of the next group
        mbar_wait_t(&hin->gfull[g ^ 1], (uint32_t)((it + 1) / 2) & 1u, w_in, dbg);
        dh += hin->ring[(g ^ 1) * HG][i];
      }
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
    printf("wave_bwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma %lld\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, w_in, w_out, w_tma);
}

// Helper warp of layer 1 (backward): dx = da W_x^(1)T for every step of layer 1, W_x^T in registers (lane i owns column i),
// handed to layer 0 in groups of HG rows aligned with layer 0's chunks.
__device__ __forceinline__ void wave_bwd_dx_helper(const WaveBwdArgs& a, int i, const float2* myWxT, Handoff3* hda, Handoff* hout) {
  float2 w[48];                                          // (W_x^T[2q][i], W_x^T[2q+1][i]) over the 96 gate columns r|u|c
#pragma unroll
  for (int q = 0; q < 48; ++q) w[q] = myWxT[q * HP + i];
  const unsigned n = (unsigned)a.S[1], vofs = group_vofs(a.S[0]);
  for (unsigned idx = 0; idx < n; ++idx) {
    const int slot = idx & (HRS - 1);
    mbar_wait(&hda->full[slot], (idx / HRS) & 1u);
    const float4* d4 = reinterpret_cast<const float4*>(hda->ring[slot]);
    float2 x0 = make_float2(0.f, 0.f), x1 = x0, x2 = x0, x3 = x0;
#pragma unroll
    for (int q4 = 0; q4 < 24; q4 += 2) {
      const float4 va = d4[q4], vb = d4[q4 + 1];
      x0 = ffma2(make_float2(va.x, va.y), w[2 * q4], x0);
      x1 = ffma2(make_float2(va.z, va.w), w[2 * q4 + 1], x1);
      x2 = ffma2(make_float2(vb.x, vb.y), w[2 * q4 + 2], x2);
      x3 = ffma2(make_float2(vb.z, vb.w), w[2 * q4 + 3], x3);
    }
    const float dx = ((x0.x + x1.x) + (x2.x + x3.x)) + ((x0.y + x1.y) + (x2.y + x3.y));
    __syncwarp();
    if (i == 0) mbar_arrive(&hda->empty[slot]);          // row consumed
    const unsigned v = idx + vofs;                       // virtual hand-off index: group = v / HG, ring slot = v % HRS
    const int g = (v / HG) & 1;
    if ((v & (HG - 1)) == 0 && v >= HRS) mbar_wait(&hout->gempty[g], (v / HRS - 1) & 1u);
    hout->ring[v & (HRS - 1)][i] = dx;
    if (((v + 1) & (HG - 1)) == 0 || idx + 1 == n) {
      __syncwarp();
      if (i == 0) mbar_arrive(&hout->gfull[g]);
    }
  }
}

template <int CH, int NS, bool DBG>
__device__ __forceinline__ void wave_bwd_layer(const WaveBwdArgs& a, int k, int b, int i, unsigned char* reg, Handoff* hin,
                                               Handoff* hout, const float2* myWxT) {
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

  float2 wrT[16], wuT[16], wcT[16];                      // lane i: W[Din+i][g*H + j] over j, natural (2q, 2q+1) pairs
  {
    const float* WhT = a.pw + a.WhT[k];                  // [3][32 j][32 i]
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      wrT[q] = make_float2(__ldg(WhT + (0 * HP + 2 * q) * HP + i), __ldg(WhT + (0 * HP + 2 * q + 1) * HP + i));
      wuT[q] = make_float2(__ldg(WhT + (1 * HP + 2 * q) * HP + i), __ldg(WhT + (1 * HP + 2 * q + 1) * HP + i));
      wcT[q] = make_float2(__ldg(WhT + (2 * HP + 2 * q) * HP + i), __ldg(WhT + (2 * HP + 2 * q + 1) * HP + i));
    }
  }
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

  float dh_next = i < H ? __ldg(a.dmemory + ((int64_t)b * L + k) * H + i) : 0.f;   // memory-slot gradient enters at the last step
  unsigned got = 0, sent = 0;                            // hand-offs consumed / produced
  int to_fire = 1;                                       // the last step is a firing step: S % p == 0
  const bool grouped_out = k == 2 && bwd_fast01(L, a.P); // layer 1 consumes through wave_bwd_fast
  const unsigned vofs = grouped_out ? group_vofs(a.S[1]) : 0u;

  auto step = [&](const float* ib, float* orow, bool first_step) {
    const float hp = first_step ? 0.f : ib[i];           // h_{s-1}: state row s-1 (zero state before step 0)
    const float r = ib[ST + HP + i], u = ib[ST + 2 * HP + i], c = ib[ST + 3 * HP + i];
    float dh = dh_next;
    if (hin != nullptr && --to_fire == 0) {              // this step fed layer k+1 in the forward pass
      to_fire = period;
      const int slot = got & (HRS - 1);
      mbar_wait_t(&hin->full[slot], (got / HRS) & 1u, w_in, dbg);
      dh += hin->ring[slot][i];
      ++got;
      __syncwarp();
      if (i == 0) mbar_arrive(&hin->empty[slot]);
    }
    const float dc = dh * (1.f - u);
    const float du = dh * (hp - c);
    float dhp = dh * u;
    const float dac = dc * (1.f - c * c);
    sh_c[i] = dac;
    __syncwarp();
    const float drh = dotn4(sh_c, wcT, 0.f);             // (da_c * Wc^T)[Din + i]
    const float dr = drh * hp;
    dhp = fmaf(drh, r, dhp);
    const float dar = dr * r * (1.f - r);
    const float dau = du * u * (1.f - u);
    sh_r[i] = dar;
    sh_u[i] = dau;
    __syncwarp();
    const float dhg = dotn(sh_r, wrT, 0.f) + dotn(sh_u, wuT, 0.f);   // (da_g * Wg^T)[Din + i]
    dh_next = dhp + dhg;
    orow[i] = dar;
    orow[HP + i] = dau;
    orow[2 * HP + i] = dac;
    if (hout != nullptr) {                               // dx of this step -> layer k-1 (every step of layer k is one of its firing steps)
      float2 x0 = make_float2(0.f, 0.f), x1 = x0, x2 = x0;
      const float4* r4 = reinterpret_cast<const float4*>(sh_r);
      const float4* u4 = reinterpret_cast<const float4*>(sh_u);
      const float4* c4 = reinterpret_cast<const float4*>(sh_c);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const float4 vr = r4[q4], vu = u4[q4], vc = c4[q4];
        x0 = ffma2(make_float2(vr.x, vr.y), myWxT[(2 * q4) * HP + i], x0);
        x0 = ffma2(make_float2(vr.z, vr.w), myWxT[(2 * q4 + 1) * HP + i], x0);
        x1 = ffma2(make_float2(vu.x, vu.y), myWxT[(16 + 2 * q4) * HP + i], x1);
        x1 = ffma2(make_float2(vu.z, vu.w), myWxT[(16 + 2 * q4 + 1) * HP + i], x1);
        x2 = ffma2(make_float2(vc.x, vc.y), myWxT[(32 + 2 * q4) * HP + i], x2);
        x2 = ffma2(make_float2(vc.z, vc.w), myWxT[(32 + 2 * q4 + 1) * HP + i], x2);
      }
      const float dx = (x0.x + x0.y) + (x1.x + x1.y) + (x2.x + x2.y);
      if (grouped_out) {                                 // -> layer 1 (wave_bwd_fast): groups of HG rows aligned to ITS chunks
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
  };

  for (int it = 0; it < nch; ++it) {
    const int ci = nch - 1 - it;
    const int stage = it % NS;
    const int s0 = ci * CH;
    const int len = min(CH, S - s0);
    mbar_wait_t(&full[stage], (uint32_t)(it / NS) & 1u, w_tma, dbg);
    float* ob = s_da + (it & 1) * CH * G3;
    if (it >= 2) {
      if (i == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    const float* ib = s_st + stage * (CH + 1) * ST;
    if (len == CH && s0 > 0) {
#pragma unroll 2                                         // deeper unrolling spills at the 168-register cap
      for (int t = CH - 1; t >= 0; --t) step(ib + t * ST, ob + t * G3, false);
    } else {
      for (int t = len - 1; t >= 0; --t) step(ib + t * ST, ob + t * G3, s0 + t == 0);
    }
    fence_proxy_async();
    __syncwarp();
    if (i == 0) {
      bulk_s2g(dab + (int64_t)s0 * G3, ob, (uint32_t)len * G3 * 4);
      bulk_commit();
      if (it + NS < nch) issue(nch - 1 - (it + NS), stage);
    }
  }
  if (i == 0) bulk_wait_read<0>();
  __syncwarp();
  if (dbg && i == 0 && b == 0)
    printf("wave_bwd layer %d: steps %d total %lld cyc (%lld/step)  wait_in %lld  wait_out %lld  wait_tma %lld\n", k, S,
           clock64() - t_start, (clock64() - t_start) / S, w_in, w_out, w_tma);
}

// Registers are partitioned per SM sub-partition (16 K each): 10 warps = 3 on one SMSP = at most 168 per thread.
template <bool DBG>
__global__ void __launch_bounds__(384)
wave_bwd_kernel(const __grid_constant__ WaveBwdArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, w = tid >> 5, i = tid & 31;
  const int L = a.L, nspc = a.nspc;
  const int k = a.wlayer[w], si = a.wsample[w];         // see plan_warps()
  const int b = blockIdx.x * nspc + si;
  const WaveBwdSmem sm(L, nspc);
  float2* sWxT = reinterpret_cast<float2*>(dsm + sm.wxt);
  Handoff* hand = reinterpret_cast<Handoff*>(dsm + sm.hand);
  Handoff3* hand3 = reinterpret_cast<Handoff3*>(dsm + sm.hand3);

  // ---- CTA setup: W_x^T of layers >= 1 as pairs over n: sWxT[kk-1][q][i] = (WxT[2q][i], WxT[2q+1][i]) ----
  for (int e = tid; e < (L - 1) * 48 * HP; e += blockDim.x) {
    const int kk = 1 + e / (48 * HP), r = e % (48 * HP);
    const int q = r / HP, ii = r % HP;
    const float* WxT = a.pw + a.WxT[kk];                 // [96][32]
    sWxT[e] = make_float2(__ldg(WxT + (2 * q) * HP + ii), __ldg(WxT + (2 * q + 1) * HP + ii));
  }
  for (int e = tid; e < (L - 1) * nspc * HRS; e += blockDim.x) {
    mbar_init(&hand[e / HRS].full[e % HRS], 1);
    mbar_init(&hand[e / HRS].empty[e % HRS], 1);
    if (e % HRS < 2) { mbar_init(&hand[e / HRS].gfull[e % HRS], 1); mbar_init(&hand[e / HRS].gempty[e % HRS], 1); }
  }
  if (L > 1)
    for (int e = tid; e < nspc * HRS; e += blockDim.x) {
      mbar_init(&hand3[e / HRS].full[e % HRS], 1);
      mbar_init(&hand3[e / HRS].empty[e % HRS], 1);
    }
  fence_mbar_init();
  __syncthreads();
  if (b >= a.B || k < 0) return;

  const bool fast01 = bwd_fast01(L, a.P);
  Handoff* hin = k < L - 1 ? &hand[k * nspc + si] : nullptr;        // from layer k+1
  Handoff* hout = k > 0 ? &hand[(k - 1) * nspc + si] : nullptr;     // to layer k-1
  if (a.whelper[w]) {                                    // layer 1's dx warp (only planned when fast01)
    wave_bwd_dx_helper(a, i, sWxT, &hand3[si], &hand[0 * nspc + si]);
    return;
  }
  if (k == 0) {
    unsigned char* reg0 = dsm + sm.l0 + si * bwd_region_bytes(BCH0, BNS0);
    if (hin == nullptr) wave_bwd_fast<false, false, DBG>(a, 0, b, i, reg0, nullptr, nullptr);
    else if (fast01) wave_bwd_fast<true, false, DBG>(a, 0, b, i, reg0, hin, nullptr);
    else wave_bwd_layer<BCH0, BNS0, DBG>(a, k, b, i, reg0, hin, nullptr, nullptr);
  } else if (k == 1) {
    unsigned char* reg1 = dsm + sm.l1 + si * bwd_region_bytes(BCH0, BNS0);
    if (!fast01) wave_bwd_layer<BCHK, BNSK, DBG>(a, k, b, i, reg1, hin, hout, sWxT);
    else if (hin == nullptr) wave_bwd_fast<false, true, DBG>(a, 1, b, i, reg1, nullptr, &hand3[si]);
    else wave_bwd_fast<true, true, DBG>(a, 1, b, i, reg1, hin, &hand3[si]);
  } else {
    wave_bwd_layer<BCHK, BNSK, DBG>(a, k, b, i, dsm + sm.lk + ((k - 2) * nspc + si) * bwd_region_bytes(BCHK, BNSK), hin, hout,
                               sWxT + (size_t)(k - 1) * 48 * HP);
  }
}

bool launch_wave_bwd(const Launch& L, const Dims& d, const PackLayout& pk, const float* pw, const float* const* st,
                     float* const* da, const float* dmemory, cudaStream_t st_) {
  if (d.L > WAVE_MAX_L) return false;
  const int nspc = d.L <= 5 ? 2 : 1;
  const WaveBwdSmem sm(d.L, nspc);
  if (sm.total > 220 * 1024) return false;
  WaveBwdArgs a; memset(&a, 0, sizeof(a));
  a.pw = pw; a.dmemory = dmemory;
  a.B = d.B; a.L = d.L; a.H = d.H; a.nspc = nspc;
  for (int k = 0; k < d.L; ++k) { a.st[k] = st[k]; a.da[k] = da[k]; a.WhT[k] = pk.WhT[k]; a.WxT[k] = pk.WxT[k]; a.S[k] = d.S[k]; a.P[k] = d.P[k]; }
  const int grid = (d.B + nspc - 1) / nspc;
  const WarpPlan wp = plan_warps(d.L, nspc, bwd_fast01(d.L, d.P));
  for (int i = 0; i < 16; ++i) { a.wlayer[i] = wp.layer[i]; a.wsample[i] = wp.sample[i]; a.whelper[i] = wp.helper[i]; }
  bool debug = false;
  { static int once = 0; const char* e = getenv("HPMN_WAVE_DEBUG"); debug = e && e[0] == '1' && once++ == 3; }   // 4th call only
  if (debug) {
    cudaFuncSetAttribute(wave_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total);
    wave_bwd_kernel<true><<<grid, 32 * wp.n, sm.total, st_>>>(a);
  } else {
    cudaFuncSetAttribute(wave_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total);
    wave_bwd_kernel<false><<<grid, 32 * wp.n, sm.total, st_>>>(a);
  }
  { cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { if (getenv("HPMN_VERBOSE")) fprintf(stderr, "wave_bwd launch failed: %s (threads %d smem %d)\n", cudaGetErrorString(e), 32 * d.L * nspc, sm.total); return false; } }
  ++*L.counter;
  return true;
}

}  // namespace hpmn
