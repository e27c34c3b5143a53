// comm.cu -- the gradient exchange of the data-parallel step over NVLink 5 / NVSwitch (the reference has no distributed code;
// batch rows sharded, table and dense parameters replicated, gradients summed over ranks).
//
// Two exchanges, both on buffers that live in SYMMETRIC memory (every rank maps every peer's copy; PyTorch's
// torch.distributed._symmetric_memory does the allocation / handle exchange -- plumbing -- and hands this library raw pointers):
//
//   * in-switch all-reduce (nvls_allreduce_kernel): rank r owns slice r of the flat [dense grads | table grad] buffer,
//     multimem.ld_reduce pulls the SUM of all ranks' copies of an element through the switch (one NVLink transfer instead
//     of N-1) and multimem.st broadcasts the result back into every copy.  Per GPU ~|buffer| in and ~|buffer| out, independent
//     of N -- against 2 (N-1)/N |buffer| each way for a ring.
//   * peer-row scatter (no kernel of its own): the embedding scatter-add (embed.cu) is pointed at a PEER's ids / dX rows, so the
//     rows cross NVLink as coalesced streaming loads inside the kernel that adds them into the local table gradient.  Per GPU
//     (N-1) x (ids + touched rows) in -- 36 MB per peer at XLong against 212 MB for the dense buffer.
//
// The caller brackets both with symmetric-memory barriers (all ranks' inputs final before, all ranks done reading after).
#include "common.cuh"

namespace hpmn {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mc: multicast address of the buffer; [q0, q1): this rank's range of float4 elements
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(float* __restrict__ mc, int64_t q0, int64_t q1) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < q1; i += 4 * stride) {         // four independent reductions in flight per thread
    const float4 a = multimem_ld_reduce_add(mc + 4 * i);
    const float4 b = multimem_ld_reduce_add(mc + 4 * (i + stride));
    const float4 c = multimem_ld_reduce_add(mc + 4 * (i + 2 * stride));
    const float4 d = multimem_ld_reduce_add(mc + 4 * (i + 3 * stride));
    multimem_st(mc + 4 * i, a);
    multimem_st(mc + 4 * (i + stride), b);
    multimem_st(mc + 4 * (i + 2 * stride), c);
    multimem_st(mc + 4 * (i + 3 * stride), d);
  }
  for (; i < q1; i += stride) multimem_st(mc + 4 * i, multimem_ld_reduce_add(mc + 4 * i));
}

void launch_nvls_allreduce(const Launch& L, float* mc, int64_t n_floats, int rank, int world, int ctas, cudaStream_t st) {
  const int64_t nq = n_floats / 4;                       // the buffer is padded to a multiple of 4 floats by the caller
  const int64_t per = (nq + world - 1) / world;
  const int64_t q0 = per * rank, q1 = q0 + per < nq ? q0 + per : nq;
  if (q1 <= q0) return;
  int grid = ctas > 0 ? ctas : L.sms;
  const int64_t need = (q1 - q0 + 511) / 512;
  if (grid > need) grid = (int)need;
  nvls_allreduce_kernel<<<grid, 512, 0, st>>>(mc, q0, q1);
  ++*L.counter;
}

}  // namespace hpmn
