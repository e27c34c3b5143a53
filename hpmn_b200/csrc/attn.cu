// attn.cu -- K3: covariance regulariser + multi-hop attention of the target embedding over the L
// memory slots.  Replaces get_covreg / query_memory / attention of
// /root/reference/code/hpmn.py:161-170, 172-182, 133-146 and their tf.gradients adjoint.
//
// One CTA (256 threads) per sample; memory slots, query, the score-MLP activations AND the current hop's MLP
// weights (53 KB, staged from L2 into padded shared memory) live on chip, softmax over the L <= 16 slots is a
// warp-shuffle reduction, all hops are fused, one (slot, unit) output per thread.  The backward
// kernel emits per-sample deltas only; the weight gradients are reductions over the batch and are
// queued as one batched A^T*B launch (gemm.cu).
#include "common.cuh"

namespace hpmn {

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
  static size_t bytes(int H4) { return sizeof(float) * (size_t)(H4 * A1S + ATT1 * A2S + 2 * ATT2 + ATT1 + ML * 4 * HPT + ML * ATT1 + ML * ATT2); }
};

// Wq [D,H] and Hmap [H,H] -> shared memory (coalesced; rows padded to QS = HPT + 1: column and row access conflict free)
template <int QS>
__device__ __forceinline__ void stage_qmaps(const float* __restrict__ P, const AttnArgs& a, float* sWq, float* sHm) {
  const int tid = threadIdx.x, H = a.H, D = a.D;
  for (int e = tid; e < D * H; e += NT) sWq[(e / H) * QS + e % H] = __ldg(P + a.Wq + e);
  for (int e = tid; e < H * H; e += NT) sHm[(e / H) * QS + e % H] = __ldg(P + a.Hmap + e);
}

template <int HPT>
__global__ void __launch_bounds__(NT, HPT == 32 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ AttnArgs a) {
  constexpr int HP = HPT, QS = HPT + 1;
  extern __shared__ __align__(16) float dsm[];
  __shared__ __align__(16) float sM[ML][HP];
  __shared__ float sC[ML][ML];
  __shared__ float sMean[ML], sRed[NT / 32], sS[ML], sW[ML];
  __shared__ __align__(16) float sLast[MD];
  __shared__ __align__(16) float sQ[HP];
  __shared__ float sQn[HP];
  __shared__ float sWq[MD * QS], sHm[HP * QS];
  pdl_trigger();
  pdl_wait();                                           // launched early (launch_pdl): the wavefront kernel in front must be complete
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, H = a.H, D = a.D, B = a.B, H4 = 4 * H;
  AttSmem<HPT> S(dsm, H4);
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

template <int HPT>
__global__ void __launch_bounds__(NT, HPT == 32 ? 2 : 1)
attn_bwd_kernel(const __grid_constant__ AttnArgs a) {
  constexpr int HP = HPT, QS = HPT + 1;
  extern __shared__ __align__(16) float dsm[];
  __shared__ __align__(16) float sM[ML][HP];
  __shared__ float sDm[ML][HP];
  __shared__ float sT[ML][HP];
  __shared__ float sC[ML][ML];
  __shared__ float sMean[ML], sRed[NT / 32], sW[ML], sDw[ML], sDs[ML], sMean2[ML];
  __shared__ __align__(16) float sDlast[MD], sQ[HP], sDq[HP], sDqin[HP];
  __shared__ float sWq[MD * QS], sHm[HP * QS];
  pdl_trigger();                                        // the backward wavefront kernel may set itself up while this grid drains
  pdl_wait();                                           // launched early itself: the head kernel in front must be complete
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, H = a.H, D = a.D, B = a.B, H4 = 4 * H;
  AttSmem<HPT> S(dsm, H4);                  // S.Inp holds d(inp); S.Z1 / S.Z2 hold z then dz
  const float* __restrict__ P = a.params;
  WPref<HPT> wp;
  wp.fetch(P, a, a.hops - 1, H4);
  for (int e = tid; e < L * H; e += NT) { sM[e / H][e % H] = __ldg(a.memory + (int64_t)b * L * H + e); sDm[e / H][e % H] = 0.f; }
  for (int e = tid; e < D; e += NT) sDlast[e] = __ldg(a.drepre + (int64_t)b * (H + D) + H + e);
  stage_qmaps<QS>(P, a, sWq, sHm);
  if (tid < H) {
    const float g = __ldg(a.drepre + (int64_t)b * (H + D) + tid);
    sDq[tid] = g;
    a.ws.dq[((int64_t)a.hops * B + b) * H + tid] = g;
  }
  for (int hop = a.hops - 1; hop >= 0; --hop) {
    __syncthreads();                        // previous hop's readers of the staged weights / sDq writers are done
    wp.put(H4, S.A1, S.A2, S.A3, S.B1, S.B2);
    if (hop > 0) wp.fetch(P, a, hop - 1, H4);
    if (tid < H) sQ[tid] = __ldg(a.ws.q + ((int64_t)hop * B + b) * H + tid);
    if (tid < L) sW[tid] = __ldg(a.ws.w + ((int64_t)hop * B + b) * L + tid);
    for (int e = tid; e < L * ATT1; e += NT) S.Z1[e] = __ldg(a.ws.z1 + ((int64_t)hop * B + b) * L * ATT1 + e);
    for (int e = tid; e < L * ATT2; e += NT) S.Z2[e] = __ldg(a.ws.z2 + ((int64_t)hop * B + b) * L * ATT2 + e);
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

static AttnArgs make_args(const Dims& d, const ParamLayout& pl, int last_offset, const float* memory, const float* x,
                          const float* params, const AttWs& ws) {
  AttnArgs a; memset(&a, 0, sizeof(a));
  a.memory = memory; a.x = x; a.params = params; a.ws = ws;
  a.B = d.B; a.L = d.L; a.H = d.H; a.D = d.D; a.Tpad = d.Tpad; a.hops = d.hops; a.last_tp = d.Tpad - last_offset;
  a.Wq = pl.Wq; a.bq = pl.bq; a.Hmap = pl.Hmap;
  for (int h = 0; h < d.hops; ++h) {
    a.A1[h] = pl.A1[h]; a.a1[h] = pl.a1[h]; a.A2[h] = pl.A2[h]; a.a2[h] = pl.a2[h]; a.A3[h] = pl.A3[h]; a.a3[h] = pl.a3[h];
  }
  return a;
}

void launch_attn_fwd(const Launch& L, const Dims& d, const ParamLayout& pl, int last_offset, const float* memory,
                     const float* x, const float* params, float* repre, float* w_hop0, float* scalars, const AttWs& ws,
                     cudaStream_t st) {
  AttnArgs a = make_args(d, pl, last_offset, memory, x, params, ws);
  a.repre = repre; a.w_hop0 = w_hop0; a.scalars = scalars;
  if (d.H <= 32) {
    const size_t dsm = AttSmem<32>::bytes(4 * d.H);
    cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_fwd_kernel<32>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  } else {
    const size_t dsm = AttSmem<64>::bytes(4 * d.H);
    cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_fwd_kernel<64>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  }
  ++*L.counter;
}

void launch_attn_bwd(const Launch& L, const Dims& d, const ParamLayout& pl, int last_offset, float memory_reg,
                     const float* memory, const float* x, const float* params, const float* drepre, float* dmemory,
                     float* dlast, float* grads, const AttWs& ws, AtbBatch& batch, cudaStream_t st) {
  AttnArgs a = make_args(d, pl, last_offset, memory, x, params, ws);
  a.drepre = drepre; a.dmemory = dmemory; a.dlast = dlast; a.memory_reg = memory_reg;
  if (d.H <= 32) {
    const size_t dsm = AttSmem<32>::bytes(4 * d.H);
    cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_bwd_kernel<32>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  } else {
    const size_t dsm = AttSmem<64>::bytes(4 * d.H);
    cudaFuncSetAttribute(attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_bwd_kernel<64>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  }
  ++*L.counter;
  // weight gradients: reductions over the batch, queued for one batched launch
  auto add = [&](const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int64_t M, int I, int N) {
    if (batch.n == ATB_MAX) { launch_atb_batch(L, batch, st); batch.n = 0; batch.blocks = 0; }
    atb_add(batch, L.sms, A, lda, Bm, ldb, C, ldc, M, I, N);
  };
  const int64_t BL = (int64_t)d.B * d.L;
  const int H = d.H, H4 = 4 * d.H;
  for (int h = 0; h < d.hops; ++h) {
    const float* inp = ws.inp + (int64_t)h * BL * H4;
    const float* z1 = ws.z1 + (int64_t)h * BL * ATT1;
    const float* z2 = ws.z2 + (int64_t)h * BL * ATT2;
    const float* dz1 = ws.dz1 + (int64_t)h * BL * ATT1;
    const float* dz2 = ws.dz2 + (int64_t)h * BL * ATT2;
    const float* ds = ws.ds + (int64_t)h * BL;
    add(inp, H4, dz1, ATT1, grads + pl.A1[h], ATT1, BL, H4, ATT1);
    add(nullptr, 0, dz1, ATT1, grads + pl.a1[h], ATT1, BL, 1, ATT1);
    add(z1, ATT1, dz2, ATT2, grads + pl.A2[h], ATT2, BL, ATT1, ATT2);
    add(nullptr, 0, dz2, ATT2, grads + pl.a2[h], ATT2, BL, 1, ATT2);
    add(z2, ATT2, ds, 1, grads + pl.A3[h], 1, BL, ATT2, 1);
    add(nullptr, 0, ds, 1, grads + pl.a3[h], 1, BL, 1, 1);
    add(ws.q + (int64_t)h * d.B * H, H, ws.dq + (int64_t)(h + 1) * d.B * H, H, grads + pl.Hmap, H, d.B, H, H);
  }
  add(x + (int64_t)(d.Tpad - last_offset) * d.D, (int64_t)d.Tpad * d.D, ws.dq, H, grads + pl.Wq, H, d.B, d.D, H);
  add(nullptr, 0, ws.dq, H, grads + pl.bq, H, d.B, 1, H);
}

}  // namespace hpmn
