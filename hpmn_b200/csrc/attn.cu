// attn.cu -- K3: covariance regulariser + multi-hop attention of the target embedding over the L
// memory slots.  Replaces get_covreg / query_memory / attention of
// /root/reference/code/hpmn.py:161-170, 172-182, 133-146 and their tf.gradients adjoint.
//
// The kernel bodies live in midbody.cuh (shared with the fused training-step kernel of mid.cu).
// One CTA (256 threads) per sample; memory slots, query, the score-MLP activations AND the current hop's MLP
// weights (53 KB, staged from L2 into padded shared memory) live on chip, softmax over the L <= 16 slots is a
// warp-shuffle reduction, all hops are fused, one (slot, unit) output per thread.  The backward
// kernel emits per-sample deltas only; the weight gradients are reductions over the batch and are
// queued as one batched A^T*B launch (gemm.cu).
#include "midbody.cuh"

namespace hpmn {

// dynamic shared memory: [AttSh | staged score-MLP weights and activations (AttSmem)]
template <int HPT>
static size_t att_smem_bytes(int H4) { return sizeof(AttSh<HPT>) + AttSmem<HPT>::bytes(H4); }

template <int HPT>
__global__ void __launch_bounds__(NT, HPT == 32 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(16) unsigned char dsm_raw[];
  AttSh<HPT>& sh = *reinterpret_cast<AttSh<HPT>*>(dsm_raw);
  const AttSmem<HPT> S(reinterpret_cast<float*>(dsm_raw + sizeof(AttSh<HPT>)), 4 * a.H);
  pdl_trigger();
  pdl_wait();                                           // launched early (launch_pdl): the wavefront kernel in front must be complete
  attn_fwd_body<HPT, false>(a, sh, S);
}

template <int HPT>
__global__ void __launch_bounds__(NT, HPT == 32 ? 2 : 1)
attn_bwd_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ __align__(16) unsigned char dsm_raw[];
  AttSh<HPT>& sh = *reinterpret_cast<AttSh<HPT>*>(dsm_raw);
  const AttSmem<HPT> S(reinterpret_cast<float*>(dsm_raw + sizeof(AttSh<HPT>)), 4 * a.H);   // S.Inp holds d(inp); S.Z1 / S.Z2 hold z then dz
  pdl_trigger();                                        // the backward wavefront kernel may set itself up while this grid drains
  pdl_wait();                                           // launched early itself: the head kernel in front must be complete
  attn_bwd_body<HPT, false>(a, sh, S);
}

AttnArgs make_attn_args(const Dims& d, const ParamLayout& pl, int last_offset, const float* memory, const float* x,
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
  AttnArgs a = make_attn_args(d, pl, last_offset, memory, x, params, ws);
  a.repre = repre; a.w_hop0 = w_hop0; a.scalars = scalars;
  if (d.H <= 32) {
    const size_t dsm = att_smem_bytes<32>(4 * d.H);
    cudaFuncSetAttribute(attn_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_fwd_kernel<32>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  } else {
    const size_t dsm = att_smem_bytes<64>(4 * d.H);
    cudaFuncSetAttribute(attn_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_fwd_kernel<64>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  }
  ++*L.counter;
}

void launch_attn_bwd(const Launch& L, const Dims& d, const ParamLayout& pl, int last_offset, float memory_reg,
                     const float* memory, const float* x, const float* params, const float* drepre, float* dmemory,
                     float* dlast, float* grads, const AttWs& ws, AtbBatch& batch, cudaStream_t st) {
  AttnArgs a = make_attn_args(d, pl, last_offset, memory, x, params, ws);
  a.drepre = drepre; a.dmemory = dmemory; a.dlast = dlast; a.memory_reg = memory_reg;
  if (d.H <= 32) {
    const size_t dsm = att_smem_bytes<32>(4 * d.H);
    cudaFuncSetAttribute(attn_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_bwd_kernel<32>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  } else {
    const size_t dsm = att_smem_bytes<64>(4 * d.H);
    cudaFuncSetAttribute(attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    launch_pdl(attn_bwd_kernel<64>, dim3(d.B), dim3(NT), (size_t)dsm, st, a);
  }
  ++*L.counter;
  queue_attn_wgrads(L, d, pl, last_offset, x, grads, ws, batch, st);
}

// weight gradients: reductions over the batch, queued for one batched launch
void queue_attn_wgrads(const Launch& L, const Dims& d, const ParamLayout& pl, int last_offset, const float* x, float* grads,
                       const AttWs& ws, AtbBatch& batch, cudaStream_t st) {
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
