// embed.cu -- K1 embedding gather / K5 scatter-add.
// Replaces tf.nn.embedding_lookup (+ mask_lookup_table multiply, + zero front padding) of
// /root/reference/code/hpmn.py:414-430, 266-282, 288-289 and its gradient.
//
// HBM-bound.  A row is E floats (64 B at E=16): E/4 lanes move one row with 128-bit accesses, so a
// warp reads 8 random rows and writes 512 contiguous bytes of x.  Each thread keeps 4 independent row
// loads in flight (grid-stride, unrolled) and the grid is a multiple of the SM count.
#include <stdlib.h>

#include "common.cuh"

namespace hpmn {

// One work item = one 16-byte quarter (q) of one (b, tp, f) row.  A thread owns U items a grid stride apart and
// runs them in phases -- U index computations, U id loads, U row loads, U stores -- so U dependent id -> row
// chains are in flight per thread (Little: ~5 MB of 64-byte requests must be outstanding to cover HBM latency).
// IdxT = uint32_t whenever the item count fits (32-bit div/mod instead of the 64-bit software routine).
template <int U, typename IdxT>
__global__ void __launch_bounds__(256)
gather_fwd_kernel(const int32_t* __restrict__ ids, const float4* __restrict__ table, float4* __restrict__ x,
                  int64_t total_, int T, int Tpad, int F, int E4, int front_pad, int mask_id0, int64_t V,
                  float* __restrict__ iderr, int l2_64) {
  const IdxT total = (IdxT)total_;
  const IdxT stride = (IdxT)gridDim.x * blockDim.x;
  for (IdxT i0 = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    IdxT src[U];
    int q[U];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const IdxT i = i0 + (IdxT)u * stride;
      live[u] = false; src[u] = 0; q[u] = 0;
      if (i < total && i >= i0) {                   // i >= i0: no wrap-around of the 32-bit index
        q[u] = (int)(i % (IdxT)E4);
        IdxT r = i / (IdxT)E4;
        const IdxT f = r % (IdxT)F; r /= (IdxT)F;
        const IdxT tp = r % (IdxT)Tpad;
        const IdxT b = r / (IdxT)Tpad;
        if (tp >= (IdxT)front_pad) { live[u] = true; src[u] = (b * (IdxT)T + (tp - (IdxT)front_pad)) * (IdxT)F + f; }
      }
    }
    int32_t id[U];
#pragma unroll
    for (int u = 0; u < U; ++u) id[u] = live[u] ? __ldg(ids + src[u]) : 0;
    float4 v[U];
    bool bad = false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool oob = id[u] < 0 || (int64_t)id[u] >= V;
      bad |= live[u] && oob;
      if (live[u] && !oob && !(mask_id0 && id[u] == 0))
        v[u] = l2_64 ? ldg_nc_f4_l2_64(table + (int64_t)id[u] * E4 + q[u]) : ldg_nc_f4(table + (int64_t)id[u] * E4 + q[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const IdxT i = i0 + (IdxT)u * stride;
      if (i < total && i >= i0) x[i] = v[u];
    }
    if (bad) *iderr = 1.0f;                          // reported as HPMN_EINVAL by the *_host entry points
  }
}

// Up to 15 sources per launch (blockIdx.y): the local (ids, dX, dlast) of the step, or -- peer-row gradient exchange, comm.cu --
// the symmetric-memory mappings of the PEER ranks' buffers, read over NVLink as coalesced streaming loads by the very kernel
// that adds them into the local table gradient.
struct ScatterSources {
  const int32_t* ids[15]; const float4* dx[15]; const float4* dlast[15];
};

// Hot rows.  With popularity-skewed ids (Zipf, or the one-uid-per-sample column of the real XLong feed) thousands of atomics hit the
// same 64-byte row and serialise in L2 (Zipf(1.05): 0.136 ms against 0.027 ms uniform).  Each CTA therefore keeps a small
// direct-mapped cache of row accumulators in shared memory: the first id to hash into a slot owns it for the lifetime of the
// CTA, later occurrences of that id are added in shared memory, everything else goes to global memory as before; the slots are
// flushed once at the end.  A hot id shows up within the first rows of every CTA with near certainty, so it is the hot rows that
// get the slots; for uniform ids the cache costs one shared-memory tag read per item.
constexpr int SC_SLOTS = 256;          // slots per CTA
constexpr int SC_E = 16;               // floats per slot (E <= 16; wider rows bypass the cache)

template <int U, typename IdxT>
__global__ void __launch_bounds__(256)
gather_bwd_kernel(const __grid_constant__ ScatterSources srcs, float* __restrict__ dtable, int64_t total_, int T, int Tpad, int F,
                  int E4, int front_pad, int last_tp, int mask_id0, int64_t V, int use_cache) {
  __shared__ int s_tag[SC_SLOTS];
  __shared__ int s_hits;
  __shared__ __align__(16) float s_acc[SC_SLOTS][SC_E];
  const int32_t* __restrict__ ids = srcs.ids[blockIdx.y];
  const float4* __restrict__ dx = srcs.dx[blockIdx.y];
  const float4* __restrict__ dlast = srcs.dlast[blockIdx.y];
  const int cache_on = use_cache;       // use_cache may be switched off per thread below; the flush follows the launch-time choice
  if (use_cache) {
    for (int e = threadIdx.x; e < SC_SLOTS; e += 256) s_tag[e] = -1;
    if (threadIdx.x == 0) s_hits = 0;
    for (int e = threadIdx.x; e < SC_SLOTS * SC_E; e += 256) (&s_acc[0][0])[e] = 0.f;
    __syncthreads();
  }
  const IdxT total = (IdxT)total_;
  const IdxT stride = (IdxT)gridDim.x * blockDim.x;
  int round = 0;
  for (IdxT i0 = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U, ++round) {
    // after two rounds (2048 items of this CTA) a cache that has seen almost no repeated id is switched off for this thread:
    // uniform ids then pay the tag reads for the first rounds only (what is already cached is still flushed at the end)
    if (use_cache && round == 2 && *reinterpret_cast<volatile int*>(&s_hits) < 16) use_cache = 0;
    IdxT isrc[U], src[U], lsrc[U];
    int q[U];
    bool live[U], has_last[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const IdxT i = i0 + (IdxT)u * stride;
      live[u] = i < total && i >= i0; has_last[u] = false;
      isrc[u] = 0; src[u] = 0; lsrc[u] = 0; q[u] = 0;
      if (live[u]) {
        q[u] = (int)(i % (IdxT)E4);
        IdxT r = i / (IdxT)E4;
        const IdxT f = r % (IdxT)F; r /= (IdxT)F;
        const IdxT t = r % (IdxT)T;
        const IdxT b = r / (IdxT)T;
        const IdxT tp = t + (IdxT)front_pad;
        isrc[u] = (b * (IdxT)T + t) * (IdxT)F + f;
        src[u] = ((b * (IdxT)Tpad + tp) * (IdxT)F + f) * (IdxT)E4 + (IdxT)q[u];   // < B*Tpad*F*E4, checked by the launcher
        has_last[u] = dlast != nullptr && tp == (IdxT)last_tp;
        lsrc[u] = (b * (IdxT)F + f) * (IdxT)E4 + (IdxT)q[u];
      }
    }
    int32_t id[U];
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      id[u] = live[u] ? __ldg(ids + isrc[u]) : -1;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live[u]) v[u] = ldg_nc_f4(dx + src[u]);    // streaming and independent of the id
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (id[u] < 0 || (int64_t)id[u] >= V || (mask_id0 && id[u] == 0)) continue;
      if (has_last[u]) {
        const float4 w = __ldg(dlast + lsrc[u]);
        v[u].x += w.x; v[u].y += w.y; v[u].z += w.z; v[u].w += w.w;
      }
      if (use_cache) {
        const int slot = (int)(((uint32_t)id[u] * 2654435761u) >> 24);      // SC_SLOTS = 256
        int tag = *reinterpret_cast<volatile int*>(&s_tag[slot]);
        const bool repeat = tag == id[u];                                   // the id was already cached by an earlier row
        if (tag == -1) { const int old = atomicCAS(&s_tag[slot], -1, id[u]); tag = old == -1 ? id[u] : old; }
        if (tag == id[u]) {
          if (repeat && round < 2 && q[u] == 0) atomicAdd(&s_hits, 1);      // one count per repeated row
          float* a = &s_acc[slot][4 * q[u]];
          atomicAdd(a, v[u].x); atomicAdd(a + 1, v[u].y); atomicAdd(a + 2, v[u].z); atomicAdd(a + 3, v[u].w);
          continue;
        }
      }
      red_add_f4(dtable + ((int64_t)id[u] * E4 + q[u]) * 4, v[u]);
    }
  }
  if (cache_on) {
    __syncthreads();
    for (int e = threadIdx.x; e < SC_SLOTS * E4; e += 256) {
      const int slot = e / E4, qq = e % E4;
      const int tag = s_tag[slot];
      if (tag < 0) continue;
      const float4 a = *reinterpret_cast<const float4*>(&s_acc[slot][4 * qq]);
      if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_f4(dtable + ((int64_t)tag * E4 + qq) * 4, a);
    }
  }
}

// resident 256-thread blocks per SM of a kernel on the CURRENT device (register-limited), queried once per device
template <class K>
static int resident_blocks(K kern) {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 256, 0) != cudaSuccess || n < 1) n = 4;
    cache[dev] = n;
  }
  return cache[dev];
}

// all blocks resident at once (per_sm per SM), every thread running the same number of full U-batches
static inline int grid_for(int64_t total, int threads, int sms, int per_sm) {
  int64_t need = (total + threads - 1) / threads;
  if (need < 1) need = 1;
  const int64_t cap = (int64_t)sms * per_sm;
  const int64_t iters = (need + cap - 1) / cap;
  return (int)((need + iters - 1) / iters);
}

void launch_gather_fwd(const Launch& L, const Dims& d, bool mask_id0, int front_pad, int64_t V, const int32_t* ids,
                       const float* table, float* x, float* iderr, cudaStream_t st) {
  int E4 = d.E / 4;
  int64_t total = (int64_t)d.B * d.Tpad * d.F * E4;
  const int per_sm = resident_blocks(gather_fwd_kernel<4, uint32_t>);
  int grid = grid_for(total, 256 * 4, L.sms, per_sm);
  // rows of at most 64 bytes: ask L2 for 64-byte fills (HPMN_GATHER_L2_64=0 restores the default 128-byte granularity)
  static const int l2_64_env = [] { const char* e = getenv("HPMN_GATHER_L2_64"); return e ? atoi(e) : 1; }();
  const int l2_64 = (l2_64_env && d.E * 4 <= 64) ? 1 : 0;
  if (total < (int64_t)1 << 31)
    gather_fwd_kernel<4, uint32_t><<<grid, 256, 0, st>>>(ids, (const float4*)table, (float4*)x, total, d.T, d.Tpad, d.F, E4,
                                                         front_pad, mask_id0 ? 1 : 0, V, iderr, l2_64);
  else
    gather_fwd_kernel<4, int64_t><<<grid, 256, 0, st>>>(ids, (const float4*)table, (float4*)x, total, d.T, d.Tpad, d.F, E4,
                                                        front_pad, mask_id0 ? 1 : 0, V, iderr, l2_64);
  ++*L.counter;
}

void launch_gather_bwd_multi(const Launch& L, const Dims& d, bool mask_id0, int front_pad, int last_offset, int64_t V, int nsrc,
                             const int32_t* const* ids, const float* const* dx, const float* const* dlast, float* dtable,
                             cudaStream_t st) {
  int E4 = d.E / 4;
  int64_t total = (int64_t)d.B * d.T * d.F * E4;
  const int per_sm = resident_blocks(gather_bwd_kernel<4, uint32_t>);
  static const int cache_env = [] { const char* e = getenv("HPMN_SCATTER_CACHE"); return e ? atoi(e) : 1; }();
  const int use_cache = (cache_env && d.E <= SC_E) ? 1 : 0;
  for (int s0 = 0; s0 < nsrc; s0 += 15) {
    const int n = nsrc - s0 < 15 ? nsrc - s0 : 15;
    ScatterSources S; memset(&S, 0, sizeof(S));
    for (int i = 0; i < n; ++i) {
      S.ids[i] = ids[s0 + i]; S.dx[i] = reinterpret_cast<const float4*>(dx[s0 + i]);
      S.dlast[i] = reinterpret_cast<const float4*>(dlast ? dlast[s0 + i] : nullptr);
    }
    // all sources share the machine: the grid of one source is sized for 1/n of the resident blocks (at least one per SM)
    int share = per_sm / n; if (share < 1) share = 1;
    dim3 grid(grid_for(total, 256 * 4, L.sms, share), n);
    if ((int64_t)d.B * d.Tpad * d.F * E4 < (int64_t)1 << 31)
      gather_bwd_kernel<4, uint32_t><<<grid, 256, 0, st>>>(S, dtable, total, d.T, d.Tpad, d.F, E4, front_pad, d.Tpad - last_offset,
                                                           mask_id0 ? 1 : 0, V, use_cache);
    else
      gather_bwd_kernel<4, int64_t><<<grid, 256, 0, st>>>(S, dtable, total, d.T, d.Tpad, d.F, E4, front_pad, d.Tpad - last_offset,
                                                          mask_id0 ? 1 : 0, V, use_cache);
    ++*L.counter;
  }
}

void launch_gather_bwd(const Launch& L, const Dims& d, bool mask_id0, int front_pad, int last_offset, int64_t V,
                       const int32_t* ids, const float* dx, const float* dlast, float* dtable, cudaStream_t st) {
  launch_gather_bwd_multi(L, d, mask_id0, front_pad, last_offset, V, 1, &ids, &dx, dlast ? &dlast : nullptr, dtable, st);
}

}  // namespace hpmn
