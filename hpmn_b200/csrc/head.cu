// head.cu -- prediction head and loss.  Replaces build_fc_net of /root/reference/code/hpmn.py:184-202:
//   bn1 (built with the default training=False, so it normalises with the never-updated moving stats
//   mean 0 / var 1: y = gamma*x/sqrt(1+1e-3)+beta) -> fc1 200 ELU -> dropout -> fc2 80 ELU -> dropout
//   -> fc3 1 sigmoid ; tf.losses.log_loss (eps 1e-7, mean over the batch).
// One CTA (256 threads) per sample, thread = output unit, weights through L1/L2 (read by every CTA).
// Backward emits per-sample deltas; weight gradients are queued as batched A^T*B problems (gemm.cu).
#include "common.cuh"

namespace hpmn {

constexpr int MR = HP + 64;   // H + F*E <= 96 in this build

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

__global__ void __launch_bounds__(256)
head_fwd_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ float sBn[MR], sAct1[FC1], sAct2[FC2];
  const int b = blockIdx.x, tid = threadIdx.x, R = a.R;
  const float* P = a.params;
  const bool drop = a.keep_prob < 1.f;
  const float inv_keep = 1.f / a.keep_prob;
  if (tid < R) {
    const float v = __ldg(a.repre + (int64_t)b * R + tid) * a.inv_bn * __ldg(P + a.gamma + tid) + __ldg(P + a.beta + tid);
    sBn[tid] = v;
    a.ws.bn[(int64_t)b * R + tid] = v;
  }
  __syncthreads();
  if (tid < FC1) {
    float s = __ldg(P + a.f1 + tid);
    for (int i = 0; i < R; ++i) s = fmaf(sBn[i], __ldg(P + a.F1 + (int64_t)i * FC1 + tid), s);
    a.ws.a1[(int64_t)b * FC1 + tid] = s;
    float act = elu_f(s);
    if (drop) act = dropout_keep(a.seed, 0, a.row0 + b, tid, a.keep_prob) ? act * inv_keep : 0.f;
    sAct1[tid] = act;
    a.ws.act1[(int64_t)b * FC1 + tid] = act;
  }
  __syncthreads();
  if (tid < FC2) {
    float s = __ldg(P + a.f2 + tid);
    for (int i = 0; i < FC1; ++i) s = fmaf(sAct1[i], __ldg(P + a.F2 + (int64_t)i * FC2 + tid), s);
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
  __shared__ float sDl1[FC1], sDl2[FC2], sDlogit;
  const int b = blockIdx.x, tid = threadIdx.x, R = a.R;
  const float* P = a.params;
  const bool drop = a.keep_prob < 1.f;
  const float inv_keep = 1.f / a.keep_prob;
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
    float d = sDlogit * __ldg(P + a.F3 + tid);
    if (drop) d = dropout_keep(a.seed, 1, a.row0 + b, tid, a.keep_prob) ? d * inv_keep : 0.f;
    d *= elu_grad_f(__ldg(a.ws.a2 + (int64_t)b * FC2 + tid));
    sDl2[tid] = d;
    a.ws.dl2[(int64_t)b * FC2 + tid] = d;
  }
  __syncthreads();
  if (tid < FC1) {
    const float* row = P + a.F2 + (int64_t)tid * FC2;
    float d = 0.f;
    for (int o = 0; o < FC2; ++o) d = fmaf(sDl2[o], __ldg(row + o), d);
    if (drop) d = dropout_keep(a.seed, 0, a.row0 + b, tid, a.keep_prob) ? d * inv_keep : 0.f;
    d *= elu_grad_f(__ldg(a.ws.a1 + (int64_t)b * FC1 + tid));
    sDl1[tid] = d;
    a.ws.dl1[(int64_t)b * FC1 + tid] = d;
  }
  __syncthreads();
  if (tid < R) {
    const float* row = P + a.F1 + (int64_t)tid * FC1;
    float d = 0.f;
    for (int o = 0; o < FC1; ++o) d = fmaf(sDl1[o], __ldg(row + o), d);
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
  head_fwd_kernel<<<d.B, 256, 0, st>>>(a);
  ++*L.counter;
}

void launch_head_bwd(const Launch& L, const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
                     const int32_t* labels, const float* params, const float* pred, float* drepre, float* grads,
                     const HeadWs& ws, AtbBatch& batch, cudaStream_t st) {
  HeadArgs a = make_args(d, pl, hy, row0, repre, labels, params, ws);
  a.pred_in = pred; a.drepre = drepre;
  head_bwd_kernel<<<d.B, 256, 0, st>>>(a);
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
