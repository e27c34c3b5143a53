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
// forward: proj [B,S,3,32] (x*Wx + b), Wh [3][32][32] -> hs [B,S,32], gates [B,S,3,32], memory[:,k,:]
// ---------------------------------------------------------------------------------------------
constexpr int RING = 4;

__global__ void __launch_bounds__(32)
rec_fwd_kernel(const float* __restrict__ proj, const float* __restrict__ Wh, float* __restrict__ hs,
               float* __restrict__ gates, float* __restrict__ memory, int S, int H, int L, int k) {
  __shared__ __align__(16) float sh_h[32];
  __shared__ __align__(16) float sh_rh[32];
  const int b = blockIdx.x, j = threadIdx.x;
  const int pos = pair_pos(j);
  float2 wr[16], wu[16], wc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    wr[i] = make_float2(Wh[(0 * HP + i) * HP + j], Wh[(0 * HP + i + 16) * HP + j]);
    wu[i] = make_float2(Wh[(1 * HP + i) * HP + j], Wh[(1 * HP + i + 16) * HP + j]);
    wc[i] = make_float2(Wh[(2 * HP + i) * HP + j], Wh[(2 * HP + i + 16) * HP + j]);
  }
  const float* pp = proj + (int64_t)b * S * G3 + j;
  float* ho = hs + (int64_t)b * S * HP + j;
  float* go = gates + (int64_t)b * S * G3 + j;
  float pr[RING], pu[RING], pc[RING];
#pragma unroll
  for (int q = 0; q < RING; ++q) {
    pr[q] = pu[q] = pc[q] = 0.f;
    if (q < S) { pr[q] = __ldg(pp + q * G3); pu[q] = __ldg(pp + q * G3 + HP); pc[q] = __ldg(pp + q * G3 + 2 * HP); }
  }
  float h = 0.f;                                          // zero_state, code/rnn.py:588
  for (int s0 = 0; s0 < S; s0 += RING) {
#pragma unroll
    for (int q = 0; q < RING; ++q) {
      const int s = s0 + q;
      if (s < S) {
        const float ar = pr[q], au = pu[q], ac = pc[q];
        if (s + RING < S) {
          pr[q] = __ldg(pp + (int64_t)(s + RING) * G3);
          pu[q] = __ldg(pp + (int64_t)(s + RING) * G3 + HP);
          pc[q] = __ldg(pp + (int64_t)(s + RING) * G3 + 2 * HP);
        }
        sh_h[pos] = h;
        __syncwarp();
        const float r = sigmoid_f(dot32(sh_h, wr, ar));  // util.py:95-96
        const float u = sigmoid_f(dot32(sh_h, wu, au));
        sh_rh[pos] = r * h;                              // util.py:98
        __syncwarp();
        const float c = tanh_f(dot32(sh_rh, wc, ac));    // util.py:107
        h = fmaf(u, h - c, c);                           // u*h + (1-u)*c, util.py:109
        ho[(int64_t)s * HP] = h;
        go[(int64_t)s * G3] = r;
        go[(int64_t)s * G3 + HP] = u;
        go[(int64_t)s * G3 + 2 * HP] = c;
      }
    }
  }
  if (j < H) memory[((int64_t)b * L + k) * H + j] = h;    // final state -> memory slot k, hpmn.py:121
}

void launch_rec_fwd(const Launch& L, const Dims& d, int k, const float* proj, const float* Wh, float* hs, float* gates,
                    float* memory, cudaStream_t st) {
  rec_fwd_kernel<<<d.B, 32, 0, st>>>(proj, Wh, hs, gates, memory, d.S[k], d.H, d.L, k);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
// backward: reverse-time adjoint of one layer.  Emits da [B,S,3,32] = (da_r, da_u, da_c); the
// non-recurrent halves (dx, dW, db) are dense GEMMs over da done afterwards.
//   dh arriving at step s = dh_next + dmemory[b,k] (s == S-1) + dx_up[b,(s+1)/p-1] ((s+1)%p == 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
rec_bwd_kernel(const float* __restrict__ hs, const float* __restrict__ gates, const float* __restrict__ WhT,
               const float* __restrict__ dmemory, const float* __restrict__ dx_up, float* __restrict__ da, int S, int H,
               int L, int k, int period) {
  __shared__ __align__(16) float sh_c[32];
  __shared__ __align__(16) float sh_r[32];
  __shared__ __align__(16) float sh_u[32];
  const int b = blockIdx.x, i = threadIdx.x;
  const int pos = pair_pos(i);
  float2 wrT[16], wuT[16], wcT[16];                       // lane i: W[Din+i][g*H + j] over j
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    wrT[jj] = make_float2(WhT[(0 * HP + jj) * HP + i], WhT[(0 * HP + jj + 16) * HP + i]);
    wuT[jj] = make_float2(WhT[(1 * HP + jj) * HP + i], WhT[(1 * HP + jj + 16) * HP + i]);
    wcT[jj] = make_float2(WhT[(2 * HP + jj) * HP + i], WhT[(2 * HP + jj + 16) * HP + i]);
  }
  const float* hb = hs + (int64_t)b * S * HP + i;
  const float* gb = gates + (int64_t)b * S * G3 + i;
  float* dab = da + (int64_t)b * S * G3 + i;
  const int S_up = dx_up ? S / period : 0;
  const float* dxb = dx_up ? dx_up + (int64_t)b * S_up * HP + i : nullptr;
  const float dmem = i < H ? __ldg(dmemory + ((int64_t)b * L + k) * H + i) : 0.f;

  float fr[RING], fu[RING], fc[RING], fh[RING], fe[RING];
  auto fetch = [&](int s, float& r, float& u, float& c, float& hp, float& ext) {
    r = __ldg(gb + (int64_t)s * G3);
    u = __ldg(gb + (int64_t)s * G3 + HP);
    c = __ldg(gb + (int64_t)s * G3 + 2 * HP);
    hp = s > 0 ? __ldg(hb + (int64_t)(s - 1) * HP) : 0.f;
    ext = 0.f;
    if (dxb != nullptr && (s + 1) % period == 0) ext = __ldg(dxb + (int64_t)((s + 1) / period - 1) * HP);
  };
#pragma unroll
  for (int q = 0; q < RING; ++q) {
    fr[q] = fu[q] = fc[q] = fh[q] = fe[q] = 0.f;
    if (S - 1 - q >= 0) fetch(S - 1 - q, fr[q], fu[q], fc[q], fh[q], fe[q]);
  }
  float dh_next = dmem;                                   // memory-slot gradient enters at the last step
  for (int s0 = S - 1; s0 >= 0; s0 -= RING) {
#pragma unroll
    for (int q = 0; q < RING; ++q) {
      const int s = s0 - q;
      if (s >= 0) {
        const float r = fr[q], u = fu[q], c = fc[q], hp = fh[q];
        const float dh = dh_next + fe[q];
        if (s - RING >= 0) fetch(s - RING, fr[q], fu[q], fc[q], fh[q], fe[q]);
        const float dc = dh * (1.f - u);
        const float du = dh * (hp - c);
        float dhp = dh * u;
        const float dac = dc * (1.f - c * c);
        sh_c[pos] = dac;
        __syncwarp();
        const float drh = dot32(sh_c, wcT, 0.f);          // (da_c * Wc^T)[Din + i]
        const float dr = drh * hp;
        dhp = fmaf(drh, r, dhp);
        const float dar = dr * r * (1.f - r);
        const float dau = du * u * (1.f - u);
        sh_r[pos] = dar;
        sh_u[pos] = dau;
        __syncwarp();
        const float dhg = dot32(sh_r, wrT, 0.f) + dot32(sh_u, wuT, 0.f);   // (da_g * Wg^T)[Din + i]
        dh_next = dhp + dhg;
        dab[(int64_t)s * G3] = dar;
        dab[(int64_t)s * G3 + HP] = dau;
        dab[(int64_t)s * G3 + 2 * HP] = dac;
      }
    }
  }
}

void launch_rec_bwd(const Launch& L, const Dims& d, int k, const float* hs, const float* gates, const float* WhT,
                    const float* dmemory, const float* dx_up, float* da, cudaStream_t st) {
  rec_bwd_kernel<<<d.B, 32, 0, st>>>(hs, gates, WhT, dmemory, dx_up, da, d.S[k], d.H, d.L, k, d.P[k]);
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
gru_wgrad_kernel(const float* __restrict__ xin, int64_t ldx, const float* __restrict__ hs,
                 const float* __restrict__ gates, const float* __restrict__ da, float* __restrict__ dWg,
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
        if (m % S != 0) hp = ldg_nc_f4(reinterpret_cast<const float4*>(hs + (m - 1) * HP) + q);
        rr = ldg_nc_f4(reinterpret_cast<const float4*>(gates + m * G3) + q);
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

void launch_gru_wgrad(const Launch& L, const Dims& d, int k, const float* xin, int64_t ldx, const float* hs,
                      const float* gates, const float* da, float* dWg, float* dbg, float* dWc, float* dbc,
                      cudaStream_t st) {
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
  gru_wgrad_kernel<<<(unsigned)blocks, threads, smem, st>>>(xin, ldx, hs, gates, da, dWg, dbg, dWc, dbc, M, d.S[k],
                                                            d.Din[k], DinP, d.H, rpb);
  ++*L.counter;
}

}  // namespace hpmn
