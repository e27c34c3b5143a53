// head.cu -- prediction head and loss.  Replaces build_fc_net of /root/reference/code/hpmn.py:184-202:
//   bn1 (built with the default training=False, so it normalises with the never-updated moving stats
//   mean 0 / var 1: y = gamma*x/sqrt(1+1e-3)+beta) -> fc1 200 ELU -> dropout -> fc2 80 ELU -> dropout
//   -> fc3 1 sigmoid ; tf.losses.log_loss (eps 1e-7, mean over the batch).
// One CTA (256 threads) per sample, thread = output unit, weights through L1/L2 (read by every CTA).
// Backward emits per-sample deltas; weight gradients are queued as batched A^T*B problems (gemm.cu).
#include "common.cuh"

namespace hpmn {

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

__global__ void __launch_bounds__(256)
head_fwd_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ float sBn[MR], sAct1[FC1], sAct2[FC2], sPart[F2_KS][FC2];
  pdl_trigger();
  pdl_wait();                                           // launched early (launch_pdl): the kernel in front must be complete
  const int b = blockIdx.x, tid = threadIdx.x, R = a.R;
  const float* __restrict__ P = a.params;
  const bool drop = a.keep_prob < 1.f;
  const float inv_keep = 1.f / a.keep_prob;
  if (tid < R) {
    const float v = __ldg(a.repre + (int64_t)b * R + tid) * a.inv_bn * __ldg(P + a.gamma + tid) + __ldg(P + a.beta + tid);
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

__global__ void __launch_bounds__(256)
head_bwd_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ __align__(16) float sDl1[FC1];
  __shared__ __align__(16) float sDl2[FC2];
  __shared__ float sDlogit, sPart[2][MR];
  pdl_trigger();
  pdl_wait();                                           // launched early (launch_pdl): the kernel in front must be complete
  const int b = blockIdx.x, tid = threadIdx.x, R = a.R;
  const float* __restrict__ P = a.params;
  const bool drop = a.keep_prob < 1.f;
  const float inv_keep = 1.f / a.keep_prob;
  // loads that do not depend on the deltas are issued first so their latency overlaps the chain below
  const float a2v = tid < FC2 ? __ldg(a.ws.a2 + (int64_t)b * FC2 + tid) : 0.f;
  const float a1v = tid < FC1 ? __ldg(a.ws.a1 + (int64_t)b * FC1 + tid) : 0.f;
  const float f3v = tid < FC2 ? __ldg(P + a.F3 + tid) : 0.f;
  float4 w2[FC2 / 4];                          // this thread's row of F2 (fc1 unit tid): 80 floats
  if (tid < FC1) {
    const float4* __restrict__ row = reinterpret_cast<const float4*>(P + a.F2 + (int64_t)tid * FC2);
#pragma unroll
    for (int q = 0; q < FC2 / 4; ++q) w2[q] = __ldg(row + q);
  }
  if (tid == 0) {
    const float p = __ldg(a.pred_in + b);
    const float y = (float)__ldg(a.labels + b);
    const float dpred = (-y / (p + LOGLOSS_EPS) + (1.f - y) / (1.f - p + LOGLOSS_EPS)) * a.inv_lossB;
    const float dl = dpred * p * (1.f - p);
    sDlogit = dl;
    a.ws.dlogit[b] = dl;
  }
  __syncthreads();
  if (tid < FC2) {
    float d = sDlogit * f3v;
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
    sPart[half][r] = d0 + d1;
  }
  __syncthreads();
  if (tid < R) {
    const float d = sPart[0][tid] + sPart[1][tid];
    const float xr = __ldg(a.repre + (int64_t)b * R + tid);
    a.ws.dbn[(int64_t)b * R + tid] = d;
    a.ws.dgt[(int64_t)b * R + tid] = d * xr * a.inv_bn;
    a.drepre[(int64_t)b * R + tid] = d * a.inv_bn * __ldg(P + a.gamma + tid);
  }
}

static HeadArgs make_args(const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
                          const int32_t* labels, const float* params, const HeadWs& ws) {
  HeadArgs a; memset(&a, 0, sizeof(a));
  a.repre = repre; a.labels = labels; a.params = params; a.ws = ws;
  a.B = d.B; a.R = d.R; a.row0 = row0;
  a.keep_prob = hy.keep_prob > 0.f ? hy.keep_prob : 1.f;
  a.inv_bn = 1.0f / sqrtf(1.0f + BN_EPS);
  a.inv_lossB = 1.0f / (float)(hy.loss_batch > 0 ? hy.loss_batch : d.B);
  a.seed = hy.dropout_seed;
  a.gamma = pl.gamma; a.beta = pl.beta; a.F1 = pl.F1; a.f1 = pl.f1; a.F2 = pl.F2; a.f2 = pl.f2; a.F3 = pl.F3; a.f3 = pl.f3;
  return a;
}

void launch_head_fwd(const Launch& L, const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
                     const int32_t* labels, const float* params, float* pred, float* logit, float* scalars,
                     const HeadWs& ws, cudaStream_t st) {
  HeadArgs a = make_args(d, pl, hy, row0, repre, labels, params, ws);
  a.pred = pred; a.logit = logit; a.scalars = scalars;
  launch_pdl(head_fwd_kernel, dim3(d.B), dim3(256), (size_t)0, st, a);
  ++*L.counter;
}

void launch_head_bwd(const Launch& L, const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
                     const int32_t* labels, const float* params, const float* pred, float* drepre, float* grads,
                     const HeadWs& ws, AtbBatch& batch, cudaStream_t st) {
  HeadArgs a = make_args(d, pl, hy, row0, repre, labels, params, ws);
  a.pred_in = pred; a.drepre = drepre;
  launch_pdl(head_bwd_kernel, dim3(d.B), dim3(256), (size_t)0, st, a);
  ++*L.counter;
  auto add = [&](const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc, int64_t M, int I, int N) {
    if (batch.n == ATB_MAX) { launch_atb_batch(L, batch, st); batch.n = 0; batch.blocks = 0; }
    atb_add(batch, L.sms, A, lda, Bm, ldb, C, ldc, M, I, N);
  };
  const int R = d.R;
  add(ws.act2, FC2, ws.dlogit, 1, grads + pl.F3, 1, d.B, FC2, 1);
  add(nullptr, 0, ws.dlogit, 1, grads + pl.f3, 1, d.B, 1, 1);
  add(ws.act1, FC1, ws.dl2, FC2, grads + pl.F2, FC2, d.B, FC1, FC2);
  add(nullptr, 0, ws.dl2, FC2, grads + pl.f2, FC2, d.B, 1, FC2);
  add(ws.bn, R, ws.dl1, FC1, grads + pl.F1, FC1, d.B, R, FC1);
  add(nullptr, 0, ws.dl1, FC1, grads + pl.f1, FC1, d.B, 1, FC1);
  add(nullptr, 0, ws.dgt, R, grads + pl.gamma, R, d.B, 1, R);
  add(nullptr, 0, ws.dbn, R, grads + pl.beta, R, d.B, 1, R);
}

}  // namespace hpmn
