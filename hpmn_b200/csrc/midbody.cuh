// midbody.cuh -- device bodies of the per-sample kernels between the two recurrences: covariance regulariser + multi-hop
// attention (forward, backward) and the prediction head (forward, backward).  Each body runs one sample on one CTA of 256
// threads.  attn.cu / head.cu wrap them as four kernels (the C ABI's K3 / head entry points, evaluation, the two-tower
// engine); mid.cu chains all four in ONE kernel for the training step (FUSED = true): the sample's memory slots, query
// maps and the last hop's score-MLP weights stay in shared memory from the forward to the backward half, three launches
// and two re-stagings disappear.  Replaces get_covreg / query_memory / attention / build_fc_net of
// /root/reference/code/hpmn.py:161-170, 172-182, 133-146, 184-202 and their tf.gradients adjoint.
#pragma once
#include "common.cuh"

namespace hpmn {

// A value another body of the same (fused) kernel wrote to global memory: a plain coherent load behind a __syncthreads();
// across kernels it is read-only data and may take the non-coherent path.
template <bool FUSED>
__device__ __forceinline__ float ldw(const float* p) {
  if constexpr (FUSED) return *p;
  else return __ldg(p);
}

constexpr int ML = HPMN_MAX_LAYERS;   // 16 slots max
constexpr int MD = 64;                // F*E <= 64 in this build

struct AttnArgs {
  const float* memory; const float* x; const float* params;
  float* repre; float* w_hop0; float* scalars;
  const float* drepre; float* dmemory; float* dlast;
  AttWs ws;
  int B, L, H, D, Tpad, hops, last_tp;
  float memory_reg;
  int64_t Wq, bq, Hmap;
  int64_t A1[HPMN_MAX_HOPS], a1[HPMN_MAX_HOPS], A2[HPMN_MAX_HOPS], a2[HPMN_MAX_HOPS], A3[HPMN_MAX_HOPS], a3[HPMN_MAX_HOPS];
};

constexpr int NT = 256;               // threads per CTA
// Row strides of the staged score-MLP weights: 16-byte aligned (float4 staging stores) and an odd number of float4s, so
// that both access patterns are bank-conflict free -- forward: scalar loads, lanes over the output unit (column);
// backward: float4 loads along a row, lanes over the row.
constexpr int A1S = ATT1 + 4;
constexpr int A2S = ATT2 + 4;
constexpr int PF2 = (ATT1 * ATT2 / 4 + NT - 1) / NT;     // float4s of A2 per thread (4)
// HPT: hidden width the kernels are compiled for -- 32 (one warp; every reference configuration) or 64 (tensor-core recurrence)

// One hop's score-MLP weights on their way from L2 to shared memory.  fetch() only issues the loads, so the hop that
// is being computed hides their latency; put() lands them once every reader of the previous weights has passed a barrier.
template <int HPT>
struct WPref {
  static constexpr int PF1 = (4 * HPT * ATT1 / 4 + NT - 1) / NT;   // float4s of A1 per thread (10 at 32, 20 at 64)
  float4 a1[PF1], a2[PF2];
  float a3, b1, b2;
  __device__ __forceinline__ void fetch(const float* __restrict__ P, const AttnArgs& a, int hop, int H4) {
    const int tid = threadIdx.x;
    const float4* g1 = reinterpret_cast<const float4*>(P + a.A1[hop]);
    const float4* g2 = reinterpret_cast<const float4*>(P + a.A2[hop]);
    const int n1 = H4 * (ATT1 / 4);
#pragma unroll
    for (int q = 0; q < PF1; ++q) { const int e = tid + NT * q; if (e < n1) a1[q] = __ldg(g1 + e); }
#pragma unroll
    for (int q = 0; q < PF2; ++q) { const int e = tid + NT * q; if (e < ATT1 * ATT2 / 4) a2[q] = __ldg(g2 + e); }
    a3 = tid < ATT2 ? __ldg(P + a.A3[hop] + tid) : 0.f;
    b1 = tid < ATT1 ? __ldg(P + a.a1[hop] + tid) : 0.f;
    b2 = tid < ATT2 ? __ldg(P + a.a2[hop] + tid) : 0.f;
  }
  __device__ __forceinline__ void put(int H4, float* sA1, float* sA2, float* sA3, float* sB1, float* sB2) const {
    const int tid = threadIdx.x;
    const int n1 = H4 * (ATT1 / 4);
#pragma unroll
    for (int q = 0; q < PF1; ++q) {
      const int e = tid + NT * q;
      if (e < n1) *reinterpret_cast<float4*>(sA1 + (e / (ATT1 / 4)) * A1S + (e % (ATT1 / 4)) * 4) = a1[q];
    }
#pragma unroll
    for (int q = 0; q < PF2; ++q) {
      const int e = tid + NT * q;
      if (e < ATT1 * ATT2 / 4) *reinterpret_cast<float4*>(sA2 + (e / (ATT2 / 4)) * A2S + (e % (ATT2 / 4)) * 4) = a2[q];
    }
    if (tid < ATT2) { sA3[tid] = a3; sB2[tid] = b2; }
    if (tid < ATT1) sB1[tid] = b1;
  }
};

// covariance pieces shared by fwd and bwd: centred memory mean per slot, off-diagonal C, Frobenius norm
template <int HPT>
__device__ __forceinline__ float covreg_block(const float (*sM)[HPT], float* sMean, float (*sC)[ML], float* sRed, int L,
                                              int H) {
  const int tid = threadIdx.x;
  if (tid < L) {
    float s = 0.f;
    for (int j = 0; j < H; ++j) s += sM[tid][j];
    sMean[tid] = s / (float)H;
  }
  __syncthreads();
  float part = 0.f;
  for (int e = tid; e < L * L; e += blockDim.x) {
    const int l = e / L, l2 = e % L;
    float c = 0.f;
    if (l != l2) {
      for (int j = 0; j < H; ++j) c = fmaf(sM[l][j] - sMean[l], sM[l2][j] - sMean[l2], c);
      c /= (float)H;
    }
    sC[l][l2] = c;
    part = fmaf(c, c, part);
  }
  part = warp_sum(part);
  if ((tid & 31) == 0) sRed[tid >> 5] = part;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sRed[w];
  return sqrtf(tot);
}

// dynamic shared memory carve-up (floats; every block starts on a 16-byte boundary)
template <int HPT>
struct AttSmem {
  float *A1, *A2, *A3, *B1, *B2, *Inp, *Z1, *Z2;
  __device__ AttSmem(float* base, int H4) {
    A1 = base; A2 = A1 + H4 * A1S; A3 = A2 + ATT1 * A2S; B1 = A3 + ATT2; B2 = B1 + ATT1;
    Inp = B2 + ATT2; Z1 = Inp + ML * 4 * HPT; Z2 = Z1 + ML * ATT1;
  }
  __host__ __device__ static size_t bytes(int H4) { return sizeof(float) * (size_t)(H4 * A1S + ATT1 * A2S + 2 * ATT2 + ATT1 + ML * 4 * HPT + ML * ATT1 + ML * ATT2); }
};

// Wq [D,H] and Hmap [H,H] -> shared memory (coalesced; rows padded to QS = HPT + 1: column and row access conflict free)
template <int QS>
__device__ __forceinline__ void stage_qmaps(const float* __restrict__ P, const AttnArgs& a, float* sWq, float* sHm) {
  const int tid = threadIdx.x, H = a.H, D = a.D;
  for (int e = tid; e < D * H; e += NT) sWq[(e / H) * QS + e % H] = __ldg(P + a.Wq + e);
  for (int e = tid; e < H * H; e += NT) sHm[(e / H) * QS + e % H] = __ldg(P + a.Hmap + e);
}


// per-sample state shared by the attention bodies (forward fields first; the backward half adds its adjoints)
template <int HPT>
struct AttSh {
  alignas(16) float M[ML][HPT];
  float C[ML][ML];
  float Mean[ML], Red[NT / 32], S[ML], W[ML];
  alignas(16) float Last[MD];
  alignas(16) float Q[HPT];
  float Qn[HPT];
  float Wq[MD * (HPT + 1)], Hm[HPT * (HPT + 1)];
  float Dm[ML][HPT], T[ML][HPT];
  float Dw[ML], Ds[ML], Mean2[ML];
  alignas(16) float Dlast[MD];
  alignas(16) float Dq[HPT];
  alignas(16) float Dqin[HPT];
};

template <int HPT, bool FUSED>
__device__ __forceinline__ void attn_fwd_body(const AttnArgs& a, AttSh<HPT>& sh, const AttSmem<HPT>& S) {
  constexpr int QS = HPT + 1;
  float (*sM)[HPT] = sh.M; float (*sC)[ML] = sh.C;
  float *sMean = sh.Mean, *sRed = sh.Red, *sS = sh.S, *sW = sh.W, *sLast = sh.Last, *sQ = sh.Q, *sQn = sh.Qn, *sWq = sh.Wq, *sHm = sh.Hm;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, H = a.H, D = a.D, B = a.B, H4 = 4 * H;
  const float* __restrict__ P = a.params;
  WPref<HPT> wp;
  wp.fetch(P, a, 0, H4);
  for (int e = tid; e < L * H; e += NT) sM[e / H][e % H] = __ldg(a.memory + (int64_t)b * L * H + e);
  for (int e = tid; e < D; e += NT) sLast[e] = __ldg(a.x + ((int64_t)b * a.Tpad + a.last_tp) * D + e);
  const float bqv = tid < H ? __ldg(P + a.bq + tid) : 0.f;
  stage_qmaps<QS>(P, a, sWq, sHm);
  wp.put(H4, S.A1, S.A2, S.A3, S.B1, S.B2);
  __syncthreads();
  const float nrm = covreg_block<HPT>(sM, sMean, sC, sRed, L, H);     // hpmn.py:161-170
  if (tid == 0) atomicAdd(a.scalars + HPMN_S_COVREG, nrm);
  if (tid < H) {                                                       // query = dense(last, H), hpmn.py:173
    float q0 = bqv, q1 = 0.f;
    int i = 0;
    for (; i + 2 <= D; i += 2) { q0 = fmaf(sLast[i], sWq[i * QS + tid], q0); q1 = fmaf(sLast[i + 1], sWq[(i + 1) * QS + tid], q1); }
    if (i < D) q0 = fmaf(sLast[i], sWq[i * QS + tid], q0);
    const float q = q0 + q1;
    sQ[tid] = q;
    a.ws.q[(int64_t)b * H + tid] = q;
  }
  __syncthreads();
  const int npair = (L + 1) >> 1;
  for (int hop = 0; hop < a.hops; ++hop) {
    if (hop + 1 < a.hops) wp.fetch(P, a, hop + 1, H4);                // lands after this hop's last weight read
    float* ginp = a.ws.inp + ((int64_t)hop * B + b) * L * H4;
    for (int e = tid; e < L * H4; e += NT) {                           // hpmn.py:135-136
      const int l = e / H4, c = e % H4, part = c / H, j = c % H;
      const float q = sQ[j], m = sM[l][j];
      const float v = part == 0 ? q : (part == 1 ? m : (part == 2 ? q - m : q * m));
      S.Inp[l * H4 + c] = v;
      ginp[e] = v;
    }
    __syncthreads();
    {                                                                  // fc1 (4H -> 80, relu), hpmn.py:137
      // thread = (slot pair, unit): one weight load feeds two slots, inputs come as broadcast float4s
      float* gz1 = a.ws.z1 + ((int64_t)hop * B + b) * L * ATT1;
      const int o = tid % ATT1;
      for (int lp = tid < (NT / ATT1) * ATT1 ? tid / ATT1 : npair; lp < npair; lp += NT / ATT1) {
        const int l0 = 2 * lp, l1 = min(2 * lp + 1, L - 1);
        const float4* in0 = reinterpret_cast<const float4*>(S.Inp + l0 * H4);
        const float4* in1 = reinterpret_cast<const float4*>(S.Inp + l1 * H4);
        const float* wcol = S.A1 + o;
        float p0 = S.B1[o], p1 = 0.f, r0 = p0, r1 = 0.f;
#pragma unroll 4
        for (int i4 = 0; i4 < H; ++i4) {                               // H4 / 4 float4s
          const float4 x = in0[i4], y = in1[i4];
          const float w0 = wcol[(4 * i4) * A1S], w1 = wcol[(4 * i4 + 1) * A1S], w2 = wcol[(4 * i4 + 2) * A1S],
                      w3 = wcol[(4 * i4 + 3) * A1S];
          p0 = fmaf(x.x, w0, p0); p1 = fmaf(x.y, w1, p1); p0 = fmaf(x.z, w2, p0); p1 = fmaf(x.w, w3, p1);
          r0 = fmaf(y.x, w0, r0); r1 = fmaf(y.y, w1, r1); r0 = fmaf(y.z, w2, r0); r1 = fmaf(y.w, w3, r1);
        }
        const float v0 = fmaxf(p0 + p1, 0.f), v1 = fmaxf(r0 + r1, 0.f);
        S.Z1[l0 * ATT1 + o] = v0; gz1[l0 * ATT1 + o] = v0;
        if (2 * lp + 1 < L) { S.Z1[l1 * ATT1 + o] = v1; gz1[l1 * ATT1 + o] = v1; }
      }
    }
    __syncthreads();
    {                                                                  // fc2 (80 -> 40, relu), hpmn.py:138
      float* gz2 = a.ws.z2 + ((int64_t)hop * B + b) * L * ATT2;
      const int o = tid % ATT2;
      for (int lp = tid < (NT / ATT2) * ATT2 ? tid / ATT2 : npair; lp < npair; lp += NT / ATT2) {
        const int l0 = 2 * lp, l1 = min(2 * lp + 1, L - 1);
        const float4* in0 = reinterpret_cast<const float4*>(S.Z1 + l0 * ATT1);
        const float4* in1 = reinterpret_cast<const float4*>(S.Z1 + l1 * ATT1);
        const float* wcol = S.A2 + o;
        float p0 = S.B2[o], p1 = 0.f, r0 = p0, r1 = 0.f;
#pragma unroll 4
        for (int i4 = 0; i4 < ATT1 / 4; ++i4) {
          const float4 x = in0[i4], y = in1[i4];
          const float w0 = wcol[(4 * i4) * A2S], w1 = wcol[(4 * i4 + 1) * A2S], w2 = wcol[(4 * i4 + 2) * A2S],
                      w3 = wcol[(4 * i4 + 3) * A2S];
          p0 = fmaf(x.x, w0, p0); p1 = fmaf(x.y, w1, p1); p0 = fmaf(x.z, w2, p0); p1 = fmaf(x.w, w3, p1);
          r0 = fmaf(y.x, w0, r0); r1 = fmaf(y.y, w1, r1); r0 = fmaf(y.z, w2, r0); r1 = fmaf(y.w, w3, r1);
        }
        const float v0 = fmaxf(p0 + p1, 0.f), v1 = fmaxf(r0 + r1, 0.f);
        S.Z2[l0 * ATT2 + o] = v0; gz2[l0 * ATT2 + o] = v0;
        if (2 * lp + 1 < L) { S.Z2[l1 * ATT2 + o] = v1; gz2[l1 * ATT2 + o] = v1; }
      }
    }
    __syncthreads();
    for (int l = warp; l < L; l += NT / 32) {                          // fc3 (40 -> 1), hpmn.py:139
      float s = 0.f;
      for (int o = lane; o < ATT2; o += 32) s = fmaf(S.Z2[l * ATT2 + o], S.A3[o], s);
      s = warp_sum(s);
      if (lane == 0) sS[l] = s + __ldg(P + a.a3[hop]);
    }
    __syncthreads();
    if (warp == 0) {                                                   // softmax over slots, hpmn.py:141
      const float v = lane < L ? sS[lane] : -INFINITY;
      const float mx = warp_max(v);
      const float e = lane < L ? expf(v - mx) : 0.f;
      const float sum = warp_sum(e);
      if (lane < L) {
        const float w = e / sum;
        sW[lane] = w;
        a.ws.w[((int64_t)hop * B + b) * L + lane] = w;
        if (hop == 0) a.w_hop0[(int64_t)b * L + lane] = w;             // weights[0], hpmn.py:182
      }
      __syncwarp();
      for (int j = lane; j < H; j += 32) {                             // query = query @ H + read, hpmn.py:179
        float qn = 0.f, qm = 0.f;
        for (int l = 0; l < L; ++l) qn = fmaf(sW[l], sM[l][j], qn);      // hpmn.py:143-144
        for (int i = 0; i < H; ++i) qm = fmaf(sQ[i], sHm[i * QS + j], qm);
        qn += qm;
        sQn[j] = qn;
        a.ws.q[((int64_t)(hop + 1) * B + b) * H + j] = qn;
      }
    }
    // every thread is past its last read of this hop's weights (barrier after fc3): land the next hop's
    if (hop + 1 < a.hops) wp.put(H4, S.A1, S.A2, S.A3, S.B1, S.B2);
    __syncthreads();
    if (tid < H) sQ[tid] = sQn[tid];
    __syncthreads();
  }
  if (tid < H) a.repre[(int64_t)b * (H + D) + tid] = sQ[tid];         // concat([query, last]), hpmn.py:442
  for (int e = tid; e < D; e += NT) a.repre[(int64_t)b * (H + D) + H + e] = sLast[e];
}

template <int HPT, bool FUSED>
__device__ __forceinline__ void attn_bwd_body(const AttnArgs& a, AttSh<HPT>& sh, const AttSmem<HPT>& S) {
  constexpr int QS = HPT + 1;
  float (*sM)[HPT] = sh.M; float (*sC)[ML] = sh.C; float (*sDm)[HPT] = sh.Dm; float (*sT)[HPT] = sh.T;
  float *sMean = sh.Mean, *sRed = sh.Red, *sW = sh.W, *sDw = sh.Dw, *sDs = sh.Ds, *sMean2 = sh.Mean2, *sDlast = sh.Dlast, *sQ = sh.Q,
        *sDq = sh.Dq, *sDqin = sh.Dqin, *sWq = sh.Wq, *sHm = sh.Hm;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, H = a.H, D = a.D, B = a.B, H4 = 4 * H;
  const float* __restrict__ P = a.params;
  WPref<HPT> wp;
  if (!FUSED) wp.fetch(P, a, a.hops - 1, H4);          // fused: the forward pass left the last hop's weights staged
  for (int e = tid; e < L * H; e += NT) {
    if (!FUSED) sM[e / H][e % H] = __ldg(a.memory + (int64_t)b * L * H + e);
    sDm[e / H][e % H] = 0.f;
  }
  for (int e = tid; e < D; e += NT) sDlast[e] = ldw<FUSED>(a.drepre + (int64_t)b * (H + D) + H + e);
  if (!FUSED) stage_qmaps<QS>(P, a, sWq, sHm);
  if (tid < H) {
    const float g = ldw<FUSED>(a.drepre + (int64_t)b * (H + D) + tid);
    sDq[tid] = g;
    a.ws.dq[((int64_t)a.hops * B + b) * H + tid] = g;
  }
  for (int hop = a.hops - 1; hop >= 0; --hop) {
    __syncthreads();                        // previous hop's readers of the staged weights / sDq writers are done
    if (!FUSED || hop != a.hops - 1) wp.put(H4, S.A1, S.A2, S.A3, S.B1, S.B2);
    if (hop > 0) wp.fetch(P, a, hop - 1, H4);
    if (tid < H) sQ[tid] = ldw<FUSED>(a.ws.q + ((int64_t)hop * B + b) * H + tid);
    if (tid < L) sW[tid] = ldw<FUSED>(a.ws.w + ((int64_t)hop * B + b) * L + tid);
    for (int e = tid; e < L * ATT1; e += NT) S.Z1[e] = ldw<FUSED>(a.ws.z1 + ((int64_t)hop * B + b) * L * ATT1 + e);
    for (int e = tid; e < L * ATT2; e += NT) S.Z2[e] = ldw<FUSED>(a.ws.z2 + ((int64_t)hop * B + b) * L * ATT2 + e);
    __syncthreads();
    // q_out = q_in @ Hmap + read ;  read = sum_l w_l m_l
    if (tid < H) {
      float s0 = 0.f;                                                  // d q_in = d q_out @ Hmap^T: row tid of Hmap
      for (int j = 0; j < H; ++j) s0 = fmaf(sDq[j], sHm[tid * QS + j], s0);
      sDqin[tid] = s0;
    }
    for (int l = warp; l < L; l += NT / 32) {
      float part = 0.f;
      for (int j = lane; j < H; j += 32) {
        const float dq = sDq[j], m = sM[l][j];
        sDm[l][j] = fmaf(dq, sW[l], sDm[l][j]);
        part = fmaf(m, dq, part);
      }
      const float dw = warp_sum(part);
      if (lane == 0) sDw[l] = dw;
    }
    __syncthreads();
    if (warp == 0) {                                                   // softmax adjoint
      const float w = lane < L ? sW[lane] : 0.f, dw = lane < L ? sDw[lane] : 0.f;
      const float dot = warp_sum(w * dw);
      if (lane < L) {
        const float ds = w * (dw - dot);
        sDs[lane] = ds;
        a.ws.ds[((int64_t)hop * B + b) * L + lane] = ds;
      }
    }
    __syncthreads();
    // dz2 = ds * A3 (z2 > 0)
    for (int e = tid; e < L * ATT2; e += NT) {
      const int l = e / ATT2, o = e % ATT2;
      const float v = S.Z2[e] > 0.f ? sDs[l] * S.A3[o] : 0.f;
      S.Z2[e] = v;
      a.ws.dz2[((int64_t)hop * B + b) * L * ATT2 + e] = v;
    }
    __syncthreads();
    // dz1[l][o] = (sum_o2 dz2[l][o2] A2[o][o2]) (z1 > 0): thread = (slot pair, unit o), row o of A2 as float4s
    {
      const int npair = (L + 1) >> 1;
      const int o = tid % ATT1;
      float* gdz1 = a.ws.dz1 + ((int64_t)hop * B + b) * L * ATT1;
      const float4* row = reinterpret_cast<const float4*>(S.A2 + o * A2S);
      for (int lp = tid < (NT / ATT1) * ATT1 ? tid / ATT1 : npair; lp < npair; lp += NT / ATT1) {
        const int l0 = 2 * lp, l1 = min(2 * lp + 1, L - 1);
        const float4* d0 = reinterpret_cast<const float4*>(S.Z2 + l0 * ATT2);
        const float4* d1 = reinterpret_cast<const float4*>(S.Z2 + l1 * ATT2);
        float p0 = 0.f, p1 = 0.f, r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int q = 0; q < ATT2 / 4; ++q) {
          const float4 w = row[q], x = d0[q], y = d1[q];
          p0 = fmaf(x.x, w.x, p0); p1 = fmaf(x.y, w.y, p1); p0 = fmaf(x.z, w.z, p0); p1 = fmaf(x.w, w.w, p1);
          r0 = fmaf(y.x, w.x, r0); r1 = fmaf(y.y, w.y, r1); r0 = fmaf(y.z, w.z, r0); r1 = fmaf(y.w, w.w, r1);
        }
        // Z1 is only overwritten after the barrier below (other threads still read nothing of it here: own elements only)
        const float v0 = S.Z1[l0 * ATT1 + o] > 0.f ? p0 + p1 : 0.f;
        S.Z1[l0 * ATT1 + o] = v0; gdz1[l0 * ATT1 + o] = v0;
        if (2 * lp + 1 < L) {
          const float v1 = S.Z1[l1 * ATT1 + o] > 0.f ? r0 + r1 : 0.f;
          S.Z1[l1 * ATT1 + o] = v1; gdz1[l1 * ATT1 + o] = v1;
        }
      }
    }
    __syncthreads();
    // dinp[l][i] = sum_o dz1[l][o] A1[i][o]: thread = (slot parity, input unit i), row i of A1 as float4s, 3 slots at a time
    {
      const int i = tid % H4;
      const float4* row = reinterpret_cast<const float4*>(S.A1 + i * A1S);
      const int nth = NT / H4;                                         // slot interleave (2 at H = 32)
      for (int lb = tid < nth * H4 ? tid / H4 : L; lb < L; lb += 3 * nth) {
        const int l0 = lb, l1 = min(lb + nth, L - 1), l2 = min(lb + 2 * nth, L - 1);
        const float4* d0 = reinterpret_cast<const float4*>(S.Z1 + l0 * ATT1);
        const float4* d1 = reinterpret_cast<const float4*>(S.Z1 + l1 * ATT1);
        const float4* d2 = reinterpret_cast<const float4*>(S.Z1 + l2 * ATT1);
        float p0 = 0.f, p1 = 0.f, r0 = 0.f, r1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll 5
        for (int q = 0; q < ATT1 / 4; ++q) {
          const float4 w = row[q], x = d0[q], y = d1[q], z = d2[q];
          p0 = fmaf(x.x, w.x, p0); p1 = fmaf(x.y, w.y, p1); p0 = fmaf(x.z, w.z, p0); p1 = fmaf(x.w, w.w, p1);
          r0 = fmaf(y.x, w.x, r0); r1 = fmaf(y.y, w.y, r1); r0 = fmaf(y.z, w.z, r0); r1 = fmaf(y.w, w.w, r1);
          t0 = fmaf(z.x, w.x, t0); t1 = fmaf(z.y, w.y, t1); t0 = fmaf(z.z, w.z, t0); t1 = fmaf(z.w, w.w, t1);
        }
        S.Inp[l0 * H4 + i] = p0 + p1;
        if (lb + nth < L) S.Inp[l1 * H4 + i] = r0 + r1;
        if (lb + 2 * nth < L) S.Inp[l2 * H4 + i] = t0 + t1;
      }
    }
    __syncthreads();
    // inp = [q, m, q-m, q*m]
    if (tid < H) {
      const float q = sQ[tid];
      float dQ = 0.f;
      for (int l = 0; l < L; ++l) {
        const float d0 = S.Inp[l * H4 + tid], d1 = S.Inp[l * H4 + H + tid], d2 = S.Inp[l * H4 + 2 * H + tid],
                    d3 = S.Inp[l * H4 + 3 * H + tid];
        const float m = sM[l][tid];
        dQ += d0 + d2 + d3 * m;
        sDm[l][tid] += d1 - d2 + d3 * q;
      }
      const float g = sDqin[tid] + dQ;
      sDq[tid] = g;                       // every other reader of sDq finished before the last barrier
      a.ws.dq[((int64_t)hop * B + b) * H + tid] = g;
    }
  }
  __syncthreads();
  // q0 = last @ Wq + bq
  for (int i = tid; i < D; i += NT) {
    float s = sDlast[i];
    for (int j = 0; j < H; ++j) s = fmaf(sDq[j], sWq[i * QS + j], s);
    a.dlast[(int64_t)b * D + i] = s;
  }
  // covreg adjoint: d||offdiag C||_F = C_off / ||.|| ;  C = mc mc^T / H ; mc = M - mean_j
  const float nrm = covreg_block<HPT>(sM, sMean, sC, sRed, L, H);
  const float scale = nrm > 0.f ? a.memory_reg * 2.f / ((float)H * nrm) : 0.f;   // TF yields NaN at nrm == 0; we yield 0
  for (int e = tid; e < L * H; e += NT) {
    const int l = e / H, j = e % H;
    float s = 0.f;
    for (int l2 = 0; l2 < L; ++l2) s = fmaf(sC[l][l2], sM[l2][j] - sMean[l2], s);
    sT[l][j] = s * scale;
  }
  __syncthreads();
  if (tid < L) {
    float s = 0.f;
    for (int j = 0; j < H; ++j) s += sT[tid][j];
    sMean2[tid] = s / (float)H;
  }
  __syncthreads();
  for (int e = tid; e < L * H; e += NT) {
    const int l = e / H, j = e % H;
    a.dmemory[(int64_t)b * L * H + e] = sDm[l][j] + sT[l][j] - sMean2[l];
  }
}

// ---- prediction head --------------------------------------------------------------------------
constexpr int MR = 2 * (HP + 64);   // one side: H + F*E <= 96; user + item sides concatenated (hpmn.py:452-456): <= 192

struct HeadArgs {
  const float* repre; const int32_t* labels; const float* params; const float* pred_in;
  float* pred; float* logit; float* scalars; float* drepre;
  HeadWs ws;
  int B, R, row0;
  float keep_prob, inv_bn, inv_lossB;
  uint64_t seed;
  int64_t gamma, beta, F1, f1, F2, f2, F3, f3;
};

__device__ __forceinline__ float elu_f(float a) { return a > 0.f ? a : expm1f(a); }
__device__ __forceinline__ float elu_grad_f(float a) { return a > 0.f ? 1.f : expf(a); }

// Latency notes: a CTA is one sample, so every dot product is a chain of L2-latency weight loads.  The loops are
// split over all 256 threads (K-split + shared-memory reduce where fewer outputs than threads exist) and unrolled
// deeply enough that each thread issues all its loads before the first FMA needs one.
constexpr int F2_KS = 3;                       // fc2: K = 200 split over 3 x 80 threads
constexpr int F2_KL = (FC1 + F2_KS - 1) / F2_KS;


struct HeadSh {
  float Bn[MR], Act1[FC1], Act2[FC2], Part[F2_KS][FC2];
  alignas(16) float Dl1[FC1];
  alignas(16) float Dl2[FC2];
  float Dlogit, Part2[2][MR];
};

template <bool FUSED>
__device__ __forceinline__ void head_fwd_body(const HeadArgs& a, HeadSh& hs) {
  float *sBn = hs.Bn, *sAct1 = hs.Act1, *sAct2 = hs.Act2; float (*sPart)[FC2] = hs.Part;
  const int b = blockIdx.x, tid = threadIdx.x, R = a.R;
  const float* __restrict__ P = a.params;
  const bool drop = a.keep_prob < 1.f;
  const float inv_keep = 1.f / a.keep_prob;
  if (tid < R) {
    const float v = ldw<FUSED>(a.repre + (int64_t)b * R + tid) * a.inv_bn * __ldg(P + a.gamma + tid) + __ldg(P + a.beta + tid);
    sBn[tid] = v;
    a.ws.bn[(int64_t)b * R + tid] = v;
  }
  __syncthreads();
  if (tid < FC1) {
    const float* __restrict__ W = P + a.F1 + tid;
    float s0 = __ldg(P + a.f1 + tid), s1 = 0.f;
    int i = 0;
    for (; i + 16 <= R; i += 16) {
      float w[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) w[q] = __ldg(W + (int64_t)(i + q) * FC1);
#pragma unroll
      for (int q = 0; q < 16; q += 2) { s0 = fmaf(sBn[i + q], w[q], s0); s1 = fmaf(sBn[i + q + 1], w[q + 1], s1); }
    }
    for (; i < R; ++i) s0 = fmaf(sBn[i], __ldg(W + (int64_t)i * FC1), s0);
    const float s = s0 + s1;
    a.ws.a1[(int64_t)b * FC1 + tid] = s;
    float act = elu_f(s);
    if (drop) act = dropout_keep(a.seed, 0, a.row0 + b, tid, a.keep_prob) ? act * inv_keep : 0.f;
    sAct1[tid] = act;
    a.ws.act1[(int64_t)b * FC1 + tid] = act;
  }
  __syncthreads();
  if (tid < F2_KS * FC2) {                     // partial sums of fc2 over one third of K
    const int g = tid / FC2, o = tid % FC2;
    const int i0 = g * F2_KL, i1 = min(FC1, i0 + F2_KL);
    const float* __restrict__ W = P + a.F2 + o;
    float s0 = 0.f, s1 = 0.f;
    int i = i0;
    for (; i + 17 <= i1; i += 17) {
      float w[17];
#pragma unroll
      for (int q = 0; q < 17; ++q) w[q] = __ldg(W + (int64_t)(i + q) * FC2);
#pragma unroll
      for (int q = 0; q < 16; q += 2) { s0 = fmaf(sAct1[i + q], w[q], s0); s1 = fmaf(sAct1[i + q + 1], w[q + 1], s1); }
      s0 = fmaf(sAct1[i + 16], w[16], s0);
    }
    for (; i < i1; ++i) s0 = fmaf(sAct1[i], __ldg(W + (int64_t)i * FC2), s0);
    sPart[g][o] = s0 + s1;
  }
  __syncthreads();
  if (tid < FC2) {
    float s = __ldg(P + a.f2 + tid);
#pragma unroll
    for (int g = 0; g < F2_KS; ++g) s += sPart[g][tid];
    a.ws.a2[(int64_t)b * FC2 + tid] = s;
    float act = elu_f(s);
    if (drop) act = dropout_keep(a.seed, 1, a.row0 + b, tid, a.keep_prob) ? act * inv_keep : 0.f;
    sAct2[tid] = act;
    a.ws.act2[(int64_t)b * FC2 + tid] = act;
  }
  __syncthreads();
  if (tid < 32) {
    float s = 0.f;
    for (int o = tid; o < FC2; o += 32) s = fmaf(sAct2[o], __ldg(P + a.F3 + o), s);
    s = warp_sum(s);
    if (tid == 0) {
      const float logit = s + __ldg(P + a.f3);
      const float p = 1.f / (1.f + expf(-logit));
      a.logit[b] = logit;
      a.pred[b] = p;
      const float y = (float)__ldg(a.labels + b);
      const float ll = -y * logf(p + LOGLOSS_EPS) - (1.f - y) * logf(1.f - p + LOGLOSS_EPS);
      atomicAdd(a.scalars + HPMN_S_LOGLOSS, ll * a.inv_lossB);
    }
  }
}

template <bool FUSED>
__device__ __forceinline__ void head_bwd_body(const HeadArgs& a, HeadSh& hs) {
  float *sDl1 = hs.Dl1, *sDl2 = hs.Dl2; float (*sPart2)[MR] = hs.Part2;
  const int b = blockIdx.x, tid = threadIdx.x, R = a.R;
  const float* __restrict__ P = a.params;
  const bool drop = a.keep_prob < 1.f;
  const float inv_keep = 1.f / a.keep_prob;
  // loads that do not depend on the deltas are issued first so their latency overlaps the chain below
  const float a2v = tid < FC2 ? ldw<FUSED>(a.ws.a2 + (int64_t)b * FC2 + tid) : 0.f;
  const float a1v = tid < FC1 ? ldw<FUSED>(a.ws.a1 + (int64_t)b * FC1 + tid) : 0.f;
  const float f3v = tid < FC2 ? __ldg(P + a.F3 + tid) : 0.f;
  float4 w2[FC2 / 4];                          // this thread's row of F2 (fc1 unit tid): 80 floats
  if (tid < FC1) {
    const float4* __restrict__ row = reinterpret_cast<const float4*>(P + a.F2 + (int64_t)tid * FC2);
#pragma unroll
    for (int q = 0; q < FC2 / 4; ++q) w2[q] = __ldg(row + q);
  }
  if (tid == 0) {
    const float p = ldw<FUSED>(a.pred_in + b);
    const float y = (float)__ldg(a.labels + b);
    const float dpred = (-y / (p + LOGLOSS_EPS) + (1.f - y) / (1.f - p + LOGLOSS_EPS)) * a.inv_lossB;
    const float dl = dpred * p * (1.f - p);
    hs.Dlogit = dl;
    a.ws.dlogit[b] = dl;
  }
  __syncthreads();
  if (tid < FC2) {
    float d = hs.Dlogit * f3v;
    if (drop) d = dropout_keep(a.seed, 1, a.row0 + b, tid, a.keep_prob) ? d * inv_keep : 0.f;
    d *= elu_grad_f(a2v);
    sDl2[tid] = d;
    a.ws.dl2[(int64_t)b * FC2 + tid] = d;
  }
  __syncthreads();
  if (tid < FC1) {
    const float4* d4 = reinterpret_cast<const float4*>(sDl2);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int q = 0; q < FC2 / 4; ++q) {
      const float4 v = d4[q];
      d0 = fmaf(v.x, w2[q].x, d0); d1 = fmaf(v.y, w2[q].y, d1);
      d0 = fmaf(v.z, w2[q].z, d0); d1 = fmaf(v.w, w2[q].w, d1);
    }
    float d = d0 + d1;
    if (drop) d = dropout_keep(a.seed, 0, a.row0 + b, tid, a.keep_prob) ? d * inv_keep : 0.f;
    d *= elu_grad_f(a1v);
    sDl1[tid] = d;
    a.ws.dl1[(int64_t)b * FC1 + tid] = d;
  }
  __syncthreads();
  for (int e = tid; e < 2 * R; e += 256) {     // row e%R of F1 (200 floats), half of it per work item
    const int r = e % R, half = e / R;
    const float4* __restrict__ row = reinterpret_cast<const float4*>(P + a.F1 + (int64_t)r * FC1) + half * (FC1 / 8);
    const float4* d4 = reinterpret_cast<const float4*>(sDl1) + half * (FC1 / 8);
    float4 w[FC1 / 8];
#pragma unroll
    for (int q = 0; q < FC1 / 8; ++q) w[q] = __ldg(row + q);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int q = 0; q < FC1 / 8; ++q) {
      const float4 v = d4[q];
      d0 = fmaf(v.x, w[q].x, d0); d1 = fmaf(v.y, w[q].y, d1);
      d0 = fmaf(v.z, w[q].z, d0); d1 = fmaf(v.w, w[q].w, d1);
    }
    sPart2[half][r] = d0 + d1;
  }
  __syncthreads();
  if (tid < R) {
    const float d = sPart2[0][tid] + sPart2[1][tid];
    const float xr = ldw<FUSED>(a.repre + (int64_t)b * R + tid);
    a.ws.dbn[(int64_t)b * R + tid] = d;
    a.ws.dgt[(int64_t)b * R + tid] = d * xr * a.inv_bn;
    a.drepre[(int64_t)b * R + tid] = d * a.inv_bn * __ldg(P + a.gamma + tid);
  }
}

// host side (attn.cu / head.cu)
AttnArgs make_attn_args(const Dims& d, const ParamLayout& pl, int last_offset, const float* memory, const float* x,
                        const float* params, const AttWs& ws);
HeadArgs make_head_args(const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
                        const int32_t* labels, const float* params, const HeadWs& ws);
void queue_attn_wgrads(const Launch& L, const Dims& d, const ParamLayout& pl, int last_offset, const float* x, float* grads,
                       const AttWs& ws, AtbBatch& batch, cudaStream_t st);
void queue_head_wgrads(const Launch& L, const Dims& d, const ParamLayout& pl, float* grads, const HeadWs& ws, AtbBatch& batch,
                       cudaStream_t st);

}  // namespace hpmn
