// attn.cu -- K3: covariance regulariser + multi-hop attention of the target embedding over the L
// memory slots.  Replaces get_covreg / query_memory / attention of
// /root/reference/code/hpmn.py:161-170, 172-182, 133-146 and their tf.gradients adjoint.
//
// One CTA (128 threads) per sample; memory slots, query and the score-MLP activations live in shared
// memory, softmax over the L <= 16 slots is a warp-shuffle reduction, all hops are fused.  The MLP
// weights (53 KB per hop) are read through L1/L2 -- every CTA reads the same addresses.  The backward
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

// y[l][o] = act(bias[o] + sum_i in[l][i] * W[i*NO + o]) for o = tid < NO, all l < L (8 slots at a time)
template <bool RELU>
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias, const float* in,
                                           int in_ld, int NI, int NO, int L, float* out, int out_ld, float* gout) {
  const int o = threadIdx.x;
  if (o >= NO) return;
  const float bo = __ldg(bias + o);
  for (int l0 = 0; l0 < L; l0 += 8) {
    float acc[8];
#pragma unroll
    for (int ll = 0; ll < 8; ++ll) acc[ll] = bo;
    for (int i = 0; i < NI; ++i) {
      const float w = __ldg(W + (int64_t)i * NO + o);
#pragma unroll
      for (int ll = 0; ll < 8; ++ll)
        if (l0 + ll < L) acc[ll] = fmaf(in[(l0 + ll) * in_ld + i], w, acc[ll]);
    }
#pragma unroll
    for (int ll = 0; ll < 8; ++ll)
      if (l0 + ll < L) {
        float v = RELU ? fmaxf(acc[ll], 0.f) : acc[ll];
        out[(l0 + ll) * out_ld + o] = v;
        gout[(l0 + ll) * out_ld + o] = v;
      }
  }
}

// covariance pieces shared by fwd and bwd: centred memory mean per slot, off-diagonal C, Frobenius norm
__device__ __forceinline__ float covreg_block(const float (*sM)[HP], float* sMean, float (*sC)[ML], float* sRed, int L,
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

__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __grid_constant__ AttnArgs a) {
  __shared__ float sM[ML][HP];
  __shared__ float sC[ML][ML];
  __shared__ float sMean[ML], sRed[4], sS[ML], sW[ML];
  __shared__ float sLast[MD], sQ[HP], sQn[HP];
  __shared__ float sInp[ML * 4 * HP];
  __shared__ float sZ1[ML * ATT1];
  __shared__ float sZ2[ML * ATT2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, H = a.H, D = a.D, B = a.B;
  const float* P = a.params;
  for (int e = tid; e < L * H; e += 128) sM[e / H][e % H] = __ldg(a.memory + (int64_t)b * L * H + e);
  for (int e = tid; e < D; e += 128) sLast[e] = __ldg(a.x + ((int64_t)b * a.Tpad + a.last_tp) * D + e);
  __syncthreads();
  const float nrm = covreg_block(sM, sMean, sC, sRed, L, H);          // hpmn.py:161-170
  if (tid == 0) atomicAdd(a.scalars + HPMN_S_COVREG, nrm);
  if (tid < H) {                                                       // query = dense(last, H), hpmn.py:173
    float q = __ldg(P + a.bq + tid);
    for (int i = 0; i < D; ++i) q = fmaf(sLast[i], __ldg(P + a.Wq + (int64_t)i * H + tid), q);
    sQ[tid] = q;
    a.ws.q[(int64_t)b * H + tid] = q;
  }
  __syncthreads();
  const int H4 = 4 * H;
  for (int hop = 0; hop < a.hops; ++hop) {
    float* ginp = a.ws.inp + ((int64_t)hop * B + b) * L * H4;
    for (int e = tid; e < L * H4; e += 128) {                          // hpmn.py:135-136
      const int l = e / H4, c = e % H4, part = c / H, j = c % H;
      const float q = sQ[j], m = sM[l][j];
      const float v = part == 0 ? q : (part == 1 ? m : (part == 2 ? q - m : q * m));
      sInp[l * H4 + c] = v;
      ginp[e] = v;
    }
    __syncthreads();
    dense_rows<true>(P + a.A1[hop], P + a.a1[hop], sInp, H4, H4, ATT1, L, sZ1, ATT1,
                     a.ws.z1 + ((int64_t)hop * B + b) * L * ATT1);     // hpmn.py:137
    __syncthreads();
    dense_rows<true>(P + a.A2[hop], P + a.a2[hop], sZ1, ATT1, ATT1, ATT2, L, sZ2, ATT2,
                     a.ws.z2 + ((int64_t)hop * B + b) * L * ATT2);     // hpmn.py:138
    __syncthreads();
    for (int l = warp; l < L; l += 4) {                                // hpmn.py:139
      float s = 0.f;
      for (int o = lane; o < ATT2; o += 32) s = fmaf(sZ2[l * ATT2 + o], __ldg(P + a.A3[hop] + o), s);
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
    }
    __syncthreads();
    if (tid < H) {                                                     // query = query @ H + read, hpmn.py:179
      float qn = 0.f;
      for (int l = 0; l < L; ++l) qn = fmaf(sW[l], sM[l][tid], qn);    // hpmn.py:143-144
      for (int i = 0; i < H; ++i) qn = fmaf(sQ[i], __ldg(P + a.Hmap + (int64_t)i * H + tid), qn);
      sQn[tid] = qn;
      a.ws.q[((int64_t)(hop + 1) * B + b) * H + tid] = qn;
    }
    __syncthreads();
    if (tid < H) sQ[tid] = sQn[tid];
    __syncthreads();
  }
  if (tid < H) a.repre[(int64_t)b * (H + D) + tid] = sQ[tid];         // concat([query, last]), hpmn.py:442
  for (int e = tid; e < D; e += 128) a.repre[(int64_t)b * (H + D) + H + e] = sLast[e];
}

__global__ void __launch_bounds__(128)
attn_bwd_kernel(const __grid_constant__ AttnArgs a) {
  __shared__ float sM[ML][HP];
  __shared__ float sDm[ML][HP];
  __shared__ float sT[ML][HP];
  __shared__ float sC[ML][ML];
  __shared__ float sMean[ML], sRed[4], sW[ML], sDw[ML], sDs[ML], sMean2[ML];
  __shared__ float sLast[MD], sDlast[MD], sQ[HP], sDq[HP], sDqin[HP];
  __shared__ float sDinp[ML * 4 * HP];
  __shared__ float sZ1[ML * ATT1];       // holds z1, then dz1
  __shared__ float sZ2[ML * ATT2];       // holds z2, then dz2
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, H = a.H, D = a.D, B = a.B, H4 = 4 * H;
  const float* P = a.params;
  for (int e = tid; e < L * H; e += 128) { sM[e / H][e % H] = __ldg(a.memory + (int64_t)b * L * H + e); sDm[e / H][e % H] = 0.f; }
  for (int e = tid; e < D; e += 128) {
    sLast[e] = __ldg(a.x + ((int64_t)b * a.Tpad + a.last_tp) * D + e);
    sDlast[e] = __ldg(a.drepre + (int64_t)b * (H + D) + H + e);
  }
  if (tid < H) {
    const float g = __ldg(a.drepre + (int64_t)b * (H + D) + tid);
    sDq[tid] = g;
    a.ws.dq[((int64_t)a.hops * B + b) * H + tid] = g;
  }
  __syncthreads();
  for (int hop = a.hops - 1; hop >= 0; --hop) {
    if (tid < H) sQ[tid] = __ldg(a.ws.q + ((int64_t)hop * B + b) * H + tid);
    if (tid < L) sW[tid] = __ldg(a.ws.w + ((int64_t)hop * B + b) * L + tid);
    for (int e = tid; e < L * ATT1; e += 128) sZ1[e] = __ldg(a.ws.z1 + ((int64_t)hop * B + b) * L * ATT1 + e);
    for (int e = tid; e < L * ATT2; e += 128) sZ2[e] = __ldg(a.ws.z2 + ((int64_t)hop * B + b) * L * ATT2 + e);
    __syncthreads();
    // q_out = q_in @ Hmap + read ;  read = sum_l w_l m_l
    if (tid < H) {
      float s = 0.f;
      for (int j = 0; j < H; ++j) s = fmaf(sDq[j], __ldg(P + a.Hmap + (int64_t)tid * H + j), s);
      sDqin[tid] = s;
    }
    for (int l = warp; l < L; l += 4) {
      const float dq = lane < H ? sDq[lane] : 0.f;
      const float m = lane < H ? sM[l][lane] : 0.f;
      if (lane < H) sDm[l][lane] = fmaf(dq, sW[l], sDm[l][lane]);
      const float dw = warp_sum(m * dq);
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
    for (int e = tid; e < L * ATT2; e += 128) {
      const int l = e / ATT2, o = e % ATT2;
      const float v = sZ2[e] > 0.f ? sDs[l] * __ldg(P + a.A3[hop] + o) : 0.f;
      sZ2[e] = v;
      a.ws.dz2[((int64_t)hop * B + b) * L * ATT2 + e] = v;
    }
    __syncthreads();
    // dz1[l][o] = (sum_o2 dz2[l][o2] A2[o][o2]) (z1 > 0)
    if (tid < ATT1) {
      const float* A2row = P + a.A2[hop] + (int64_t)tid * ATT2;
      for (int l0 = 0; l0 < L; l0 += 8) {
        float acc[8];
#pragma unroll
        for (int ll = 0; ll < 8; ++ll) acc[ll] = 0.f;
        for (int o2 = 0; o2 < ATT2; ++o2) {
          const float w = __ldg(A2row + o2);
#pragma unroll
          for (int ll = 0; ll < 8; ++ll)
            if (l0 + ll < L) acc[ll] = fmaf(sZ2[(l0 + ll) * ATT2 + o2], w, acc[ll]);
        }
#pragma unroll
        for (int ll = 0; ll < 8; ++ll)
          if (l0 + ll < L) {
            const int idx = (l0 + ll) * ATT1 + tid;
            const float v = sZ1[idx] > 0.f ? acc[ll] : 0.f;
            sZ1[idx] = v;                                              // own element only: no hazard
            a.ws.dz1[((int64_t)hop * B + b) * L * ATT1 + idx] = v;
          }
      }
    }
    __syncthreads();
    // dinp[l][i] = sum_o dz1[l][o] A1[i][o]
    if (tid < H4) {
      const float* A1row = P + a.A1[hop] + (int64_t)tid * ATT1;
      for (int l0 = 0; l0 < L; l0 += 8) {
        float acc[8];
#pragma unroll
        for (int ll = 0; ll < 8; ++ll) acc[ll] = 0.f;
        for (int o = 0; o < ATT1; ++o) {
          const float w = __ldg(A1row + o);
#pragma unroll
          for (int ll = 0; ll < 8; ++ll)
            if (l0 + ll < L) acc[ll] = fmaf(sZ1[(l0 + ll) * ATT1 + o], w, acc[ll]);
        }
#pragma unroll
        for (int ll = 0; ll < 8; ++ll)
          if (l0 + ll < L) sDinp[(l0 + ll) * H4 + tid] = acc[ll];
      }
    }
    __syncthreads();
    // inp = [q, m, q-m, q*m]
    if (tid < H) {
      const float q = sQ[tid];
      float dQ = 0.f;
      for (int l = 0; l < L; ++l) {
        const float d0 = sDinp[l * H4 + tid], d1 = sDinp[l * H4 + H + tid], d2 = sDinp[l * H4 + 2 * H + tid],
                    d3 = sDinp[l * H4 + 3 * H + tid];
        const float m = sM[l][tid];
        dQ += d0 + d2 + d3 * m;
        sDm[l][tid] += d1 - d2 + d3 * q;
      }
      const float g = sDqin[tid] + dQ;
      sDq[tid] = g;                       // every other reader of sDq finished before the last barrier
      a.ws.dq[((int64_t)hop * B + b) * H + tid] = g;
    }
    __syncthreads();
  }
  // q0 = last @ Wq + bq
  for (int i = tid; i < D; i += 128) {
    float s = sDlast[i];
    for (int j = 0; j < H; ++j) s = fmaf(sDq[j], __ldg(P + a.Wq + (int64_t)i * H + j), s);
    a.dlast[(int64_t)b * D + i] = s;
  }
  // covreg adjoint: d||offdiag C||_F = C_off / ||.|| ;  C = mc mc^T / H ; mc = M - mean_j
  const float nrm = covreg_block(sM, sMean, sC, sRed, L, H);
  const float scale = nrm > 0.f ? a.memory_reg * 2.f / ((float)H * nrm) : 0.f;   // TF yields NaN at nrm == 0; we yield 0
  for (int e = tid; e < L * H; e += 128) {
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
  for (int e = tid; e < L * H; e += 128) {
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
  attn_fwd_kernel<<<d.B, 128, 0, st>>>(a);
  ++*L.counter;
}

void launch_attn_bwd(const Launch& L, const Dims& d, const ParamLayout& pl, int last_offset, float memory_reg,
                     const float* memory, const float* x, const float* params, const float* drepre, float* dmemory,
                     float* dlast, float* grads, const AttWs& ws, AtbBatch& batch, cudaStream_t st) {
  AttnArgs a = make_args(d, pl, last_offset, memory, x, params, ws);
  a.drepre = drepre; a.dmemory = dmemory; a.dlast = dlast; a.memory_reg = memory_reg;
  attn_bwd_kernel<<<d.B, 128, 0, st>>>(a);
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
