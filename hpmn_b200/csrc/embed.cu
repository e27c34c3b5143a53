// embed.cu -- K1 embedding gather / K5 scatter-add.
// Replaces tf.nn.embedding_lookup (+ mask_lookup_table multiply, + zero front padding) of
// /root/reference/code/hpmn.py:414-430, 266-282, 288-289 and its gradient.
//
// HBM-bound.  A row is E floats (64 B at E=16): E/4 lanes move one row with 128-bit accesses, so a
// warp reads 8 random rows and writes 512 contiguous bytes of x.  Each thread keeps 4 independent row
// loads in flight (grid-stride, unrolled) and the grid is a multiple of the SM count.
#include "common.cuh"

namespace hpmn {

__global__ void __launch_bounds__(256)
gather_fwd_kernel(const int32_t* __restrict__ ids, const float4* __restrict__ table, float4* __restrict__ x,
                  int64_t total, int T, int Tpad, int F, int E4, int front_pad, int mask_id0, int64_t V,
                  float* __restrict__ iderr) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 4
  for (; i < total; i += stride) {
    int q = (int)(i % E4);
    int64_t r = i / E4;
    int f = (int)(r % F); r /= F;
    int tp = (int)(r % Tpad);
    int64_t b = r / Tpad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tp >= front_pad) {
      int32_t id = __ldg(ids + (b * T + (tp - front_pad)) * F + f);
      if (id < 0 || (int64_t)id >= V) {
        if (q == 0) *iderr = 1.0f;          // reported as HPMN_EINVAL by the *_host entry points
      } else if (!(mask_id0 && id == 0)) {
        v = ldg_nc_f4(table + (int64_t)id * E4 + q);
      }
    }
    x[i] = v;
  }
}

__global__ void __launch_bounds__(256)
gather_bwd_kernel(const int32_t* __restrict__ ids, const float4* __restrict__ dx, const float4* __restrict__ dlast,
                  float* __restrict__ dtable, int64_t total, int T, int Tpad, int F, int E4, int front_pad,
                  int last_tp, int mask_id0, int64_t V) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 4
  for (; i < total; i += stride) {
    int q = (int)(i % E4);
    int64_t r = i / E4;
    int f = (int)(r % F); r /= F;
    int t = (int)(r % T);
    int64_t b = r / T;
    int32_t id = __ldg(ids + (b * T + t) * F + f);
    if (id < 0 || (int64_t)id >= V || (mask_id0 && id == 0)) continue;
    int tp = t + front_pad;
    float4 v = ldg_nc_f4(dx + ((b * Tpad + tp) * F + f) * E4 + q);
    if (dlast != nullptr && tp == last_tp) {
      float4 w = __ldg(dlast + (b * F + f) * E4 + q);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    red_add_f4(dtable + ((int64_t)id * E4 + q) * 4, v);
  }
}

static inline int grid_for(int64_t total, int threads, int sms, int per_sm) {
  int64_t need = (total + threads - 1) / threads;
  int64_t cap = (int64_t)sms * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

void launch_gather_fwd(const Launch& L, const Dims& d, bool mask_id0, int front_pad, int64_t V, const int32_t* ids,
                       const float* table, float* x, float* iderr, cudaStream_t st) {
  int E4 = d.E / 4;
  int64_t total = (int64_t)d.B * d.Tpad * d.F * E4;
  int grid = grid_for(total, 256 * 4, L.sms, 8);
  gather_fwd_kernel<<<grid, 256, 0, st>>>(ids, (const float4*)table, (float4*)x, total, d.T, d.Tpad, d.F, E4,
                                          front_pad, mask_id0 ? 1 : 0, V, iderr);
  ++*L.counter;
}

void launch_gather_bwd(const Launch& L, const Dims& d, bool mask_id0, int front_pad, int last_offset, int64_t V,
                       const int32_t* ids, const float* dx, const float* dlast, float* dtable, cudaStream_t st) {
  int E4 = d.E / 4;
  int64_t total = (int64_t)d.B * d.T * d.F * E4;
  int grid = grid_for(total, 256 * 4, L.sms, 8);
  gather_bwd_kernel<<<grid, 256, 0, st>>>(ids, (const float4*)dx, (const float4*)dlast, dtable, total, d.T, d.Tpad,
                                          d.F, E4, front_pad, d.Tpad - last_offset, mask_id0 ? 1 : 0, V);
  ++*L.counter;
}

}  // namespace hpmn
