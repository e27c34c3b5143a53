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
#include "common.cuh"

namespace hpmn {

constexpr int WCH = 8;        // steps per output chunk (one 4 KB bulk store)
constexpr int WIN = 16;       // steps per input chunk of layer 0
constexpr int WNS0 = 3;       // input ring stages of layer 0
constexpr int HRS = 8;        // hand-off ring slots between consecutive layers
constexpr int WAVE_MAX_L = 11;

struct WaveArgs {
  const float* proj0;              // [B,S_0,96]
  const float* pw;                 // packed weights
  float* st[HPMN_MAX_LAYERS];      // [B,S_k,128] per layer
  float* memory;                   // [B,L,H]
  int64_t Wh[HPMN_MAX_LAYERS], Wx[HPMN_MAX_LAYERS], bx[HPMN_MAX_LAYERS];   // offsets into pw
  int S[HPMN_MAX_LAYERS], P[HPMN_MAX_LAYERS];
  int B, L, H, nspc;
};

struct Handoff {                   // layer k -> k+1 of one sample: ring of HRS rows, full / empty mbarrier per slot
  float ring[HRS][HP];
  uint64_t full[HRS], empty[HRS];
};

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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory plan (bytes)
struct WaveSmem {
  int wx, bx, hand, l0, lk, total;
  int r0, r1;                      // per-warp region sizes: layer 0 / layers >= 1
  __host__ __device__ WaveSmem(int L, int nspc) {
    int off = 0;
    wx = off; off += (L - 1) * 16 * 3 * HP * 8;            // float2 [q][g][j] per layer >= 1
    bx = off; off += (L - 1) * G3 * 4;
    off = (off + 127) & ~127;
    hand = off; off += (L - 1) * nspc * (int)sizeof(Handoff);
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
template <bool IS_L0, int CH>
__device__ __forceinline__ float wave_layer(const WaveArgs& a, int k, int b, int j, float* s_in, uint64_t* full,
                                            float* s_out, float* sh_rh, const float* zero_row, const float2* myWx,
                                            const float* myBx, Handoff* hin, Handoff* hout) {
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
  unsigned fired = 0, s_glob = 0;                        // hand-offs produced / consumed so far
  int to_fire = period;

  auto step = [&](const float* ib, float* orow) {
    float ar, au, ac;
    if (IS_L0) {
      ar = ib[j]; au = ib[HP + j]; ac = ib[2 * HP + j];
    } else {
      // input = hidden state of layer k-1 at its step (s+1)*p_{k-1} - 1 (hpmn.py:124-128), through the hand-off ring
      const int slot_in = s_glob & (HRS - 1);
      mbar_wait(&hin->full[slot_in], (s_glob / HRS) & 1u);           // hardware-suspended wait, no polling
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
      if (j == 0) mbar_arrive(&hin->empty[slot_in]);     // slot free again
      ++s_glob;
    }
    const float r = sigmoid_f(dotn(hprev, wr, ar));      // util.py:95-96
    const float u = sigmoid_f(dotn(hprev, wu, au));
    sh_rh[j] = r * h;                                    // util.py:98
    __syncwarp();
    const float c = tanh_f(dotn(sh_rh, wc, ac));         // util.py:107
    h = fmaf(u, h - c, c);                               // util.py:109
    orow[j] = h; orow[HP + j] = r; orow[2 * HP + j] = u; orow[3 * HP + j] = c;
    hprev = orow;
    if (hout != nullptr && --to_fire == 0) {             // this step feeds layer k+1
      to_fire = period;
      const int slot_out = fired & (HRS - 1);
      if (fired >= HRS) mbar_wait(&hout->empty[slot_out], (fired / HRS - 1) & 1u);
      hout->ring[slot_out][j] = h;
      ++fired;
      __syncwarp();
      if (j == 0) mbar_arrive(&hout->full[slot_out]);    // release: the row is visible to the waiting layer
    }
    __syncwarp();                                        // orow (next step's broadcast source) and sh_rh settled
  };

  for (int c = 0; c < nch; ++c) {
    const int len = min(CH, S - c * CH);
    float* ob = s_out + (c & 1) * CH * ST;
    const float* ib = nullptr;
    if (IS_L0) {
      const int stage = c % WNS0;
      mbar_wait(&full[stage], (uint32_t)(c / WNS0) & 1u);
      ib = s_in + stage * CH * G3;
    }
    if (c >= 2) {                                        // the bulk store that read this buffer two chunks ago
      if (j == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    if (len == CH) {
#pragma unroll 4
      for (int t = 0; t < CH; ++t) step(ib + t * G3, ob + t * ST);
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
  return h;
}

__global__ void __launch_bounds__(352)
wave_fwd_kernel(const __grid_constant__ WaveArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  const int tid = threadIdx.x, w = tid >> 5, j = tid & 31;
  const int L = a.L, nspc = a.nspc;
  const int k = L - 1 - w / nspc, si = w % nspc;        // layer 0 = highest warp ids = highest issue priority
  const int b = blockIdx.x * nspc + si;
  const WaveSmem sm(L, nspc);
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
  }
  fence_mbar_init();
  __syncthreads();
  if (b >= a.B) return;                                  // ragged last CTA: the whole warp leaves together

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
    h = wave_layer<true, WIN>(a, k, b, j, s_in, full, s_out, sh_rh, zero_row, nullptr, nullptr, nullptr, hout);
  } else {
    float* s_out = reinterpret_cast<float*>(reg);
    float* sh_rh = s_out + 2 * WCH * ST;
    float* zero_row = sh_rh + 32;
    zero_row[j] = 0.f;
    h = wave_layer<false, WCH>(a, k, b, j, nullptr, nullptr, s_out, sh_rh, zero_row, sWx + (size_t)(k - 1) * 16 * 3 * HP,
                               sBx + (k - 1) * G3, hin, hout);
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
  for (int k = 0; k < d.L; ++k) { a.st[k] = st[k]; a.Wh[k] = pk.Wh[k]; a.Wx[k] = pk.Wx[k]; a.bx[k] = pk.bx[k]; a.S[k] = d.S[k]; a.P[k] = d.P[k]; }
  cudaFuncSetAttribute(wave_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm.total);
  const int grid = (d.B + nspc - 1) / nspc;
  wave_fwd_kernel<<<grid, 32 * d.L * nspc, sm.total, st_>>>(a);
  ++*L.counter;
  return true;
}

}  // namespace hpmn
