// gru.cu -- K2/K4: the recurrent part of the hierarchical periodic memory.
// Replaces tf.nn.dynamic_rnn(GRUCell) of /root/reference/code/hpmn.py:118-120 (cell arithmetic
// code/util.py:81-110 without line 108; loop code/rnn.py:732-793, zero state, no length mask) and the
// tf.gradients adjoint of it.
//
// Layout: one warp owns one sample's recurrence for one layer; lane j owns hidden unit j (H <= 32,
// padded lanes carry zeros).  The 3*32 recurrent weights a lane needs live in registers as (i, i+16)
// pairs so every dot product is a chain of packed FFMA2; h (and r*h) are broadcast through 128 bytes of
// shared memory per warp in the matching pair order.  The x-half of both matmuls is precomputed by
// gemm_nn (it does not depend on h) and streamed in through a 4-deep register prefetch ring.
#include "common.cuh"

namespace hpmn {

__device__ __forceinline__ int pair_pos(int lane) { return ((lane & 15) << 1) | (lane >> 4); }

// dot(v[0..31], w) with v in shared memory in pair order and w as 16 (i, i+16) register pairs
__device__ __forceinline__ float dot32(const float* sh, const float2 (&w)[16], float init) {
  float2 a0 = make_float2(init, 0.f), a1 = make_float2(0.f, 0.f);
  const float4* s4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 v = s4[q];
    a0 = ffma2(make_float2(v.x, v.y), w[2 * q], a0);
    a1 = ffma2(make_float2(v.z, v.w), w[2 * q + 1], a1);
  }
  return (a0.x + a1.x) + (a0.y + a1.y);
}

// ---------------------------------------------------------------------------------------------
// forward: proj [B,S,3,32] (x*Wx + b), Wh [3][32][32] -> st [B,S,4,32] = h|r|u|c, memory[:,k,:]
//
// Data movement is chunked (16 steps): the TMA engine (cp.async.bulk, SASS UBLKCP) streams the next
// projection chunks into a 4-stage shared-memory ring, completion counted on mbarriers, and finished
// state chunks leave through bulk stores.  The time loop itself touches only registers and shared
// memory, so the only waits on the critical path are the recurrence's own dependencies.
// ---------------------------------------------------------------------------------------------
constexpr int CHF = 16;      // steps per chunk
constexpr int NSF = 4;       // input ring stages

__global__ void __launch_bounds__(32)
rec_fwd_kernel(const float* __restrict__ proj, const float* __restrict__ Wh, float* __restrict__ st,
               float* __restrict__ memory, int S, int H, int L, int k) {
  __shared__ __align__(128) float s_in[NSF][CHF * G3];
  __shared__ __align__(128) float s_out[2][CHF * ST];
  __shared__ __align__(16) float sh_h[32];
  __shared__ __align__(16) float sh_rh[32];
  __shared__ __align__(8) uint64_t full[NSF];
  const int b = blockIdx.x, j = threadIdx.x;
  const int pos = pair_pos(j);
  float2 wr[16], wu[16], wc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    wr[i] = make_float2(Wh[(0 * HP + i) * HP + j], Wh[(0 * HP + i + 16) * HP + j]);
    wu[i] = make_float2(Wh[(1 * HP + i) * HP + j], Wh[(1 * HP + i + 16) * HP + j]);
    wc[i] = make_float2(Wh[(2 * HP + i) * HP + j], Wh[(2 * HP + i + 16) * HP + j]);
  }
  const float* pp = proj + (int64_t)b * S * G3;
  float* so = st + (int64_t)b * S * ST;
  const int nch = (S + CHF - 1) / CHF;
  if (j == 0) {
#pragma unroll
    for (int i = 0; i < NSF; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (j == 0) {
    for (int c = 0; c < NSF && c < nch; ++c) {
      const int len = min(CHF, S - c * CHF);
      mbar_expect_tx(&full[c], (uint32_t)len * G3 * 4);
      bulk_g2s(s_in[c], pp + (int64_t)c * CHF * G3, (uint32_t)len * G3 * 4, &full[c]);
    }
  }
  float h = 0.f;                                          // zero_state, code/rnn.py:588
  auto step = [&](const float* ib, float* ob, int t) {
    const float ar = ib[t * G3 + j], au = ib[t * G3 + HP + j], ac = ib[t * G3 + 2 * HP + j];
    sh_h[pos] = h;
    __syncwarp();
    const float r = sigmoid_f(dot32(sh_h, wr, ar));      // util.py:95-96
    const float u = sigmoid_f(dot32(sh_h, wu, au));
    sh_rh[pos] = r * h;                                  // util.py:98
    __syncwarp();
    const float c = tanh_f(dot32(sh_rh, wc, ac));        // util.py:107
    h = fmaf(u, h - c, c);                               // u*h + (1-u)*c, util.py:109
    ob[t * ST + j] = h;
    ob[t * ST + HP + j] = r;
    ob[t * ST + 2 * HP + j] = u;
    ob[t * ST + 3 * HP + j] = c;
  };
  for (int c = 0; c < nch; ++c) {
    const int stage = c % NSF;
    const int len = min(CHF, S - c * CHF);
    mbar_wait(&full[stage], (uint32_t)(c / NSF) & 1u);
    float* ob = s_out[c & 1];
    if (c >= 2) {                                         // the bulk store that read this buffer two chunks ago
      if (j == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    const float* ib = s_in[stage];
    if (len == CHF) {
#pragma unroll 4
      for (int t = 0; t < CHF; ++t) step(ib, ob, t);
    } else {
      for (int t = 0; t < len; ++t) step(ib, ob, t);
    }
    fence_proxy_async();                                  // generic-proxy writes of ob -> visible to the bulk store
    __syncwarp();
    if (j == 0) {
      bulk_s2g(so + (int64_t)c * CHF * ST, ob, (uint32_t)len * ST * 4);
      bulk_commit();
      const int cn = c + NSF;                             // refill the stage every lane has finished reading
      if (cn < nch) {
        const int ln = min(CHF, S - cn * CHF);
        mbar_expect_tx(&full[stage], (uint32_t)ln * G3 * 4);
        bulk_g2s(s_in[stage], pp + (int64_t)cn * CHF * G3, (uint32_t)ln * G3 * 4, &full[stage]);
      }
    }
  }
  if (j == 0) bulk_wait_read<0>();                        // shared memory must outlive the last bulk store's reads
  __syncwarp();
  if (j < H) memory[((int64_t)b * L + k) * H + j] = h;    // final state -> memory slot k, hpmn.py:121
}

void launch_rec_fwd(const Launch& L, const Dims& d, int k, const float* proj, const float* Wh, float* st, float* memory,
                    cudaStream_t st_) {
  rec_fwd_kernel<<<d.B, 32, 0, st_>>>(proj, Wh, st, memory, d.S[k], d.H, d.L, k);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
// backward: reverse-time adjoint of one layer.  Emits da [B,S,3,32] = (da_r, da_u, da_c); the
// non-recurrent halves (dx, dW, db) are dense GEMMs over da done afterwards.
//   dh arriving at step s = dh_next + dmemory[b,k] (s == S-1) + dx_up[b,(s+1)/p-1] ((s+1)%p == 0)
// Same chunked TMA pipeline, walking the chunks from the end: a chunk needs state rows s0-1 .. s0+len-1
// (h_prev of step s is row s-1) and the len/p rows of dx_up that land on its firing steps.
// ---------------------------------------------------------------------------------------------
constexpr int CHB = 16;      // max steps per chunk (the launch picks a multiple of the period <= 16)
constexpr int NSB = 3;

__global__ void __launch_bounds__(32)
rec_bwd_kernel(const float* __restrict__ st, const float* __restrict__ WhT, const float* __restrict__ dmemory,
               const float* __restrict__ dx_up, float* __restrict__ da, int S, int H, int L, int k, int period, int chunk) {
  __shared__ __align__(128) float s_in[NSB][(CHB + 1) * ST];
  __shared__ __align__(128) float s_dx[NSB][CHB * HP];
  __shared__ __align__(128) float s_out[2][CHB * G3];
  __shared__ __align__(16) float sh_c[32];
  __shared__ __align__(16) float sh_r[32];
  __shared__ __align__(16) float sh_u[32];
  __shared__ __align__(8) uint64_t full[NSB];
  const int b = blockIdx.x, i = threadIdx.x;
  const int pos = pair_pos(i);
  float2 wrT[16], wuT[16], wcT[16];                       // lane i: W[Din+i][g*H + j] over j
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    wrT[jj] = make_float2(WhT[(0 * HP + jj) * HP + i], WhT[(0 * HP + jj + 16) * HP + i]);
    wuT[jj] = make_float2(WhT[(1 * HP + jj) * HP + i], WhT[(1 * HP + jj + 16) * HP + i]);
    wcT[jj] = make_float2(WhT[(2 * HP + jj) * HP + i], WhT[(2 * HP + jj + 16) * HP + i]);
  }
  const float* sb = st + (int64_t)b * S * ST;
  float* dab = da + (int64_t)b * S * G3;
  const int S_up = dx_up ? S / period : 0;
  const float* dxb = dx_up ? dx_up + (int64_t)b * S_up * HP : nullptr;
  const int nch = (S + chunk - 1) / chunk;
  uint32_t firemask = 0;                                  // bit t: step s0+t feeds the upper layer
  if (dx_up)
    for (int t = 0; t < chunk; ++t)
      if ((t + 1) % period == 0) firemask |= 1u << t;

  auto issue = [&](int ci, int stage) {                   // lane 0 only
    const int s0 = ci * chunk;
    const int len = min(chunk, S - s0);
    const int nj = dx_up ? len / period : 0;
    const uint32_t rows = (uint32_t)(s0 > 0 ? len + 1 : len);
    mbar_expect_tx(&full[stage], rows * ST * 4 + (uint32_t)nj * HP * 4);
    if (s0 > 0) bulk_g2s(s_in[stage], sb + (int64_t)(s0 - 1) * ST, rows * ST * 4, &full[stage]);
    else bulk_g2s(s_in[stage] + ST, sb, rows * ST * 4, &full[stage]);      // buffer row t+1 <-> step s0+t
    if (nj) bulk_g2s(s_dx[stage], dxb + (int64_t)(s0 / period) * HP, (uint32_t)nj * HP * 4, &full[stage]);
  };
  if (i == 0) {
#pragma unroll
    for (int q = 0; q < NSB; ++q) mbar_init(&full[q], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (i == 0)
    for (int it = 0; it < NSB && it < nch; ++it) issue(nch - 1 - it, it);

  float dh_next = i < H ? __ldg(dmemory + ((int64_t)b * L + k) * H + i) : 0.f;   // enters at the last step
  for (int it = 0; it < nch; ++it) {
    const int ci = nch - 1 - it;
    const int stage = it % NSB;
    const int s0 = ci * chunk;
    const int len = min(chunk, S - s0);
    mbar_wait(&full[stage], (uint32_t)(it / NSB) & 1u);
    float* ob = s_out[it & 1];
    if (it >= 2) {
      if (i == 0) bulk_wait_read<1>();
      __syncwarp();
    }
    const float* ib = s_in[stage];
    const float* xb = s_dx[stage];
    int xi = dx_up ? len / period - 1 : -1;               // dx_up row of the last firing step in the chunk
#pragma unroll 4
    for (int t = len - 1; t >= 0; --t) {
      const float hp = (s0 + t > 0) ? ib[t * ST + i] : 0.f;
      const float r = ib[(t + 1) * ST + HP + i], u = ib[(t + 1) * ST + 2 * HP + i], c = ib[(t + 1) * ST + 3 * HP + i];
      float dh = dh_next;
      if ((firemask >> t) & 1u) { dh += xb[xi * HP + i]; --xi; }
      const float dc = dh * (1.f - u);
      const float du = dh * (hp - c);
      float dhp = dh * u;
      const float dac = dc * (1.f - c * c);
      sh_c[pos] = dac;
      __syncwarp();
      const float drh = dot32(sh_c, wcT, 0.f);            // (da_c * Wc^T)[Din + i]
      const float dr = drh * hp;
      dhp = fmaf(drh, r, dhp);
      const float dar = dr * r * (1.f - r);
      const float dau = du * u * (1.f - u);
      sh_r[pos] = dar;
      sh_u[pos] = dau;
      __syncwarp();
      const float dhg = dot32(sh_r, wrT, 0.f) + dot32(sh_u, wuT, 0.f);   // (da_g * Wg^T)[Din + i]
      dh_next = dhp + dhg;
      ob[t * G3 + i] = dar;
      ob[t * G3 + HP + i] = dau;
      ob[t * G3 + 2 * HP + i] = dac;
    }
    fence_proxy_async();
    __syncwarp();
    if (i == 0) {
      bulk_s2g(dab + (int64_t)s0 * G3, ob, (uint32_t)len * G3 * 4);
      bulk_commit();
      if (it + NSB < nch) issue(nch - 1 - (it + NSB), stage);
    }
  }
  if (i == 0) bulk_wait_read<0>();
  __syncwarp();
}

void launch_rec_bwd(const Launch& L, const Dims& d, int k, const float* st, const float* WhT, const float* dmemory,
                    const float* dx_up, float* da, cudaStream_t st_) {
  const int p = d.P[k];
  const int chunk = dx_up ? (CHB / p) * p : CHB;          // multiple of the period (periods <= 16, checked in api.cu)
  rec_bwd_kernel<<<d.B, 32, 0, st_>>>(st, WhT, dmemory, dx_up, da, d.S[k], d.H, d.L, k, p, chunk);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
// weight gradients of one layer (reduction over all B*S rows):
//   dWg += [x | h_prev]^T da_g      dbg += sum da_g
//   dWc += [x | r*h_prev]^T da_c    dbc += sum da_c
// Thread (ig, ng) owns a 4 (input rows) x 8 (gate columns) block of the [DinP+32, 96] gradient; rows are
// staged 64 at a time in shared memory; the epilogue scatters into the TF layout with atomics.
// ---------------------------------------------------------------------------------------------
constexpr int WG_RC = 64;

__global__ void __launch_bounds__(320)
gru_wgrad_kernel(const float* __restrict__ xin, int64_t ldx, const float* __restrict__ st,
                 const float* __restrict__ da, float* __restrict__ dWg,
                 float* __restrict__ dbg, float* __restrict__ dWc, float* __restrict__ dbc, int64_t M, int S, int Din,
                 int DinP, int H, int64_t rows_per_block) {
  extern __shared__ __align__(16) float smem[];
  const int WA = DinP + 2 * HP;                 // x | h_prev | r*h_prev
  float* As = smem;                             // [WG_RC][WA]
  float* Ds = smem + WG_RC * WA;                // [WG_RC][96]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int IT = (DinP + HP) / 4;
  const bool active = tid < IT * 12;
  const int ig = tid % IT, ng = tid / IT;       // ng in [0,12): 8 columns each; ng >= 8 -> candidate
  int acol = ig * 4;
  if (acol >= DinP && ng >= 8) acol += HP;      // candidate columns pair with r*h_prev
  float2 acc[4][4];
  float bsum[8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) bsum[c] = 0.f;

  const int64_t mbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t mend = mbeg + rows_per_block < M ? mbeg + rows_per_block : M;
  const int X4 = DinP / 4;
  for (int64_t mc = mbeg; mc < mend; mc += WG_RC) {
    __syncthreads();
    // x part
    for (int e = tid; e < WG_RC * X4; e += nthr) {
      int r = e / X4, q = e % X4;
      int64_t m = mc + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < mend) v = ldg_nc_f4(reinterpret_cast<const float4*>(xin + m * ldx) + q);
      *reinterpret_cast<float4*>(As + r * WA + q * 4) = v;
    }
    // h_prev and r*h_prev
    for (int e = tid; e < WG_RC * (HP / 4); e += nthr) {
      int r = e / (HP / 4), q = e % (HP / 4);
      int64_t m = mc + r;
      float4 hp = make_float4(0.f, 0.f, 0.f, 0.f), rr = hp;
      if (m < mend) {
        if (m % S != 0) hp = ldg_nc_f4(reinterpret_cast<const float4*>(st + (m - 1) * ST) + q);
        rr = ldg_nc_f4(reinterpret_cast<const float4*>(st + m * ST + HP) + q);
      }
      *reinterpret_cast<float4*>(As + r * WA + DinP + q * 4) = hp;
      *reinterpret_cast<float4*>(As + r * WA + DinP + HP + q * 4) =
          make_float4(hp.x * rr.x, hp.y * rr.y, hp.z * rr.z, hp.w * rr.w);
    }
    // da
    for (int e = tid; e < WG_RC * (G3 / 4); e += nthr) {
      int r = e / (G3 / 4), q = e % (G3 / 4);
      int64_t m = mc + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < mend) v = ldg_nc_f4(reinterpret_cast<const float4*>(da + m * G3) + q);
      *reinterpret_cast<float4*>(Ds + r * G3 + q * 4) = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = 0; r < WG_RC; ++r) {
        const float4 a4 = *reinterpret_cast<const float4*>(As + r * WA + acol);
        const float4 b0 = *reinterpret_cast<const float4*>(Ds + r * G3 + ng * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(Ds + r * G3 + ng * 8 + 4);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                              make_float2(b1.z, b1.w)};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float2 aa = make_float2(av[a], av[a]);
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = ffma2(aa, bv[c], acc[a][c]);
        }
        if (ig == 0) {
          bsum[0] += b0.x; bsum[1] += b0.y; bsum[2] += b0.z; bsum[3] += b0.w;
          bsum[4] += b1.x; bsum[5] += b1.y; bsum[6] += b1.z; bsum[7] += b1.w;
        }
      }
    }
  }
  if (!active) return;
  // epilogue: scatter into TF layout.  logical input row li = ig*4 + a in [0, DinP+32)
  const int g = ng / 4;                         // 0: r, 1: u, 2: c
  const int j0 = (ng % 4) * 8;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int li = ig * 4 + a;
    int row;
    if (li < DinP) { if (li >= Din) continue; row = li; }
    else { if (li - DinP >= H) continue; row = Din + (li - DinP); }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = j0 + c;
      if (j >= H) continue;
      const float v = (c & 1) ? acc[a][c >> 1].y : acc[a][c >> 1].x;
      if (g < 2) atomicAdd(dWg + (int64_t)row * 2 * H + g * H + j, v);
      else atomicAdd(dWc + (int64_t)row * H + j, v);
    }
  }
  if (ig == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = j0 + c;
      if (j >= H) continue;
      if (g < 2) atomicAdd(dbg + g * H + j, bsum[c]);
      else atomicAdd(dbc + j, bsum[c]);
    }
  }
}

void launch_gru_wgrad(const Launch& L, const Dims& d, int k, const float* xin, int64_t ldx, const float* st,
                      const float* da, float* dWg, float* dbg, float* dWc, float* dbc, cudaStream_t st_) {
  const int DinP = d.DinP[k];
  const int IT = (DinP + HP) / 4;
  int threads = ((IT * 12 + 31) / 32) * 32;
  if (threads > 320) threads = 320;             // DinP <= 72 in this build (checked in api.cu)
  const int64_t M = (int64_t)d.B * d.S[k];
  int64_t chunks = (M + WG_RC - 1) / WG_RC;
  int64_t blocks = chunks < (int64_t)L.sms * 2 ? chunks : (int64_t)L.sms * 2;
  int64_t rpb = ((chunks + blocks - 1) / blocks) * WG_RC;
  blocks = (M + rpb - 1) / rpb;
  size_t smem = (size_t)WG_RC * (DinP + 2 * HP + G3) * sizeof(float);
  cudaFuncSetAttribute(gru_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gru_wgrad_kernel<<<(unsigned)blocks, threads, smem, st_>>>(xin, ldx, st, da, dWg, dbg, dWc, dbc, M, d.S[k],
                                                             d.Din[k], DinP, d.H, rpb);
  ++*L.counter;
}

}  // namespace hpmn
