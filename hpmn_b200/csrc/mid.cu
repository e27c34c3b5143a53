// mid.cu -- the training step's per-sample middle section as ONE kernel: covariance regulariser + multi-hop attention,
// prediction head + log-loss, head adjoint, attention / covariance adjoint (midbody.cuh; reference:
// /root/reference/code/hpmn.py:133-146, 161-182, 184-202 and their tf.gradients adjoint).
//
// Nothing between the two recurrences crosses samples -- batch normalisation runs on its never-updated moving statistics
// (training=False, hpmn.py:190), the log-loss mean only scales the per-sample delta by 1 / batch -- so one CTA carries
// its sample from the memory slots to d(memory) without leaving the SM: the slots, the query maps Wq / Hmap and the last
// hop's score-MLP weights stay staged in shared memory for the backward half; the activations the batched weight-gradient
// GEMM needs afterwards (inp, z1, z2, their deltas, bn, act1, act2 ...) are written once and re-read by the same CTA
// through L1 / L2.  Against the four separate kernels (attn_fwd, head_fwd, head_bwd, attn_bwd -- still used for
// evaluation, the K3 / head entry points of the C ABI and the two-tower engine) this removes three launch boundaries
// on the critical path of the step and two of the four weight stagings.
#include "midbody.cuh"

namespace hpmn {

struct MidArgs { AttnArgs at; HeadArgs hd; };

template <int HPT>
static size_t mid_smem_bytes(int H4) { return sizeof(AttSh<HPT>) + ((AttSmem<HPT>::bytes(H4) + 15) & ~(size_t)15) + sizeof(HeadSh); }

template <int HPT>
__global__ void __launch_bounds__(NT, HPT == 32 ? 2 : 1)
mid_fused_kernel(const __grid_constant__ MidArgs m) {
  extern __shared__ __align__(16) unsigned char dsm_raw[];
  AttSh<HPT>& sh = *reinterpret_cast<AttSh<HPT>*>(dsm_raw);
  const int H4 = 4 * m.at.H;
  const AttSmem<HPT> S(reinterpret_cast<float*>(dsm_raw + sizeof(AttSh<HPT>)), H4);
  HeadSh& hs = *reinterpret_cast<HeadSh*>(dsm_raw + sizeof(AttSh<HPT>) + ((AttSmem<HPT>::bytes(H4) + 15) & ~(size_t)15));
  pdl_trigger();                                        // the backward recurrence may set itself up while this grid drains
  pdl_wait();                                           // launched early itself: the forward recurrence must be complete
  attn_fwd_body<HPT, true>(m.at, sh, S);
  __syncthreads();                                      // repre (global) is complete for this sample
  head_fwd_body<true>(m.hd, hs);
  __syncthreads();                                      // pred, a1, a2 (global) are complete
  head_bwd_body<true>(m.hd, hs);
  __syncthreads();                                      // drepre (global) is complete
  attn_bwd_body<HPT, true>(m.at, sh, S);
}

void launch_mid_fused(const Launch& L, const Dims& d, const ParamLayout& pl, const hpmn_hyper& hy, int last_offset, int row0,
                      const float* memory, const float* x, const float* params, const int32_t* labels, float* repre,
                      float* w_hop0, float* pred, float* logit, float* scalars, float* drepre, float* dmemory, float* dlast,
                      float* grads, const AttWs& aws, const HeadWs& hws, AtbBatch& batch, cudaStream_t st) {
  MidArgs m;
  m.at = make_attn_args(d, pl, last_offset, memory, x, params, aws);
  m.at.repre = repre; m.at.w_hop0 = w_hop0; m.at.scalars = scalars;
  m.at.drepre = drepre; m.at.dmemory = dmemory; m.at.dlast = dlast; m.at.memory_reg = hy.memory_reg;
  m.hd = make_head_args(d, pl, hy, row0, repre, labels, params, hws);
  m.hd.pred = pred; m.hd.logit = logit; m.hd.scalars = scalars; m.hd.pred_in = pred; m.hd.drepre = drepre;
  // A plain launch, not launch_pdl(): launched early, this grid would park 2 CTAs on each of the 20 SMs the forward wavefront
  // kernel leaves idle and keep its other 216 CTAs queued for 290 us -- and the block scheduler does not place CTAs of a
  // younger grid (the gradient-buffer zeroing of the side stream, which is meant to use exactly those SMs) past a grid
  // it cannot finish placing (tools/timeline.py).  HPMN_MID_PDL=1 restores the early launch.
  static const bool early = [] { const char* e = getenv("HPMN_MID_PDL"); return e && e[0] == '1'; }();
  auto go = [&](auto kern, size_t dsm) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);
    if (early) launch_pdl(kern, dim3(d.B), dim3(NT), dsm, st, m);
    else kern<<<d.B, NT, dsm, st>>>(m);
  };
  if (d.H <= 32) go(mid_fused_kernel<32>, mid_smem_bytes<32>(4 * d.H));
  else go(mid_fused_kernel<64>, mid_smem_bytes<64>(4 * d.H));
  ++*L.counter;
  queue_head_wgrads(L, d, pl, grads, hws, batch, st);
  queue_attn_wgrads(L, d, pl, last_offset, x, grads, aws, batch, st);
}

}  // namespace hpmn
