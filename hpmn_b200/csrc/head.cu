// head.cu -- prediction head and loss.  Replaces build_fc_net of /root/reference/code/hpmn.py:184-202:
//   bn1 (built with the default training=False, so it normalises with the never-updated moving stats
//   mean 0 / var 1: y = gamma*x/sqrt(1+1e-3)+beta) -> fc1 200 ELU -> dropout -> fc2 80 ELU -> dropout
//   -> fc3 1 sigmoid ; tf.losses.log_loss (eps 1e-7, mean over the batch).
// The kernel bodies live in midbody.cuh (shared with the fused training-step kernel of mid.cu).
// One CTA (256 threads) per sample, thread = output unit, weights through L1/L2 (read by every CTA).
// Backward emits per-sample deltas; weight gradients are queued as batched A^T*B problems (gemm.cu).
#include "midbody.cuh"

namespace hpmn {

__global__ void __launch_bounds__(256)
head_fwd_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ HeadSh hs;
  pdl_trigger();
  pdl_wait();                                           // launched early (launch_pdl): the kernel in front must be complete
  head_fwd_body<false>(a, hs);
}

__global__ void __launch_bounds__(256)
head_bwd_kernel(const __grid_constant__ HeadArgs a) {
  __shared__ HeadSh hs;
  pdl_trigger();
  pdl_wait();                                           // launched early (launch_pdl): the kernel in front must be complete
  head_bwd_body<false>(a, hs);
}

HeadArgs make_head_args(const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
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
  HeadArgs a = make_head_args(d, pl, hy, row0, repre, labels, params, ws);
  a.pred = pred; a.logit = logit; a.scalars = scalars;
  launch_pdl(head_fwd_kernel, dim3(d.B), dim3(256), (size_t)0, st, a);
  ++*L.counter;
}

void launch_head_bwd(const Launch& L, const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int row0, const float* repre,
                     const int32_t* labels, const float* params, const float* pred, float* drepre, float* grads,
                     const HeadWs& ws, AtbBatch& batch, cudaStream_t st) {
  HeadArgs a = make_head_args(d, pl, hy, row0, repre, labels, params, ws);
  a.pred_in = pred; a.drepre = drepre;
  launch_pdl(head_bwd_kernel, dim3(d.B), dim3(256), (size_t)0, st, a);
  ++*L.counter;
  queue_head_wgrads(L, d, pl, grads, ws, batch, st);
}

void queue_head_wgrads(const Launch& L, const Dims& d, const ParamLayout& pl, float* grads, const HeadWs& ws, AtbBatch& batch,
                       cudaStream_t st) {
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
