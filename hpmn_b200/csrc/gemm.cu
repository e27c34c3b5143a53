// gemm.cu -- weight packing and the small dense helpers around the recurrent kernels.
//   pack:     TF-layout GRU weights -> warp-friendly padded blocks (rebuilt every call, params move)
//   gemm_nn:  input projections  P = X * W_x + b   (the x-half of the two _Linear calls of
//             /root/reference/code/util.py:88-107, hoisted out of the time loop because it does not
//             depend on h) and their adjoint dX = dA * W_x^T
//   gemm_atb: C += A^T * B over the row (batch) dimension -- weight gradients of dense layers
//   colsum / axpy / clip_adam / finish_scalars: small elementwise helpers
#include "common.cuh"

namespace hpmn {

// ---------------------------------------------------------------------------------------------
struct PackArgs {
  int L, H;
  int Din[HPMN_MAX_LAYERS], DinP[HPMN_MAX_LAYERS];
  int64_t Wg[HPMN_MAX_LAYERS], bg[HPMN_MAX_LAYERS], Wc[HPMN_MAX_LAYERS], bc[HPMN_MAX_LAYERS];
  int64_t Wx[HPMN_MAX_LAYERS], bx[HPMN_MAX_LAYERS], Wh[HPMN_MAX_LAYERS], WhT[HPMN_MAX_LAYERS], WxT[HPMN_MAX_LAYERS];
};

__global__ void __launch_bounds__(256)
pack_kernel(const __grid_constant__ PackArgs a, const float* __restrict__ params, float* __restrict__ pw) {
  const int k = blockIdx.y;
  const int H = a.H, Din = a.Din[k], DinP = a.DinP[k];
  const float* Wg = params + a.Wg[k];
  const float* bg = params + a.bg[k];
  const float* Wc = params + a.Wc[k];
  const float* bc = params + a.bc[k];
  const int nWx = DinP * G3, nWh = 3 * HP * HP;
  const int total = nWx + G3 + nWh + nWh + nWx;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int r = e;
    if (r < nWx) {                      // Wx[i][g*32+j]
      int i = r / G3, n = r % G3, g = n / HP, j = n % HP;
      float v = 0.f;
      if (i < Din && j < H) v = g < 2 ? Wg[(int64_t)i * 2 * H + g * H + j] : Wc[(int64_t)i * H + j];
      pw[a.Wx[k] + r] = v;
      continue;
    }
    r -= nWx;
    if (r < G3) {                       // bx[g*32+j]
      int g = r / HP, j = r % HP;
      pw[a.bx[k] + r] = j < H ? (g < 2 ? bg[g * H + j] : bc[j]) : 0.f;
      continue;
    }
    r -= G3;
    if (r < nWh) {                      // Wh[g][i][j]
      int g = r / (HP * HP), i = (r / HP) % HP, j = r % HP;
      float v = 0.f;
      if (i < H && j < H) v = g < 2 ? Wg[(int64_t)(Din + i) * 2 * H + g * H + j] : Wc[(int64_t)(Din + i) * H + j];
      pw[a.Wh[k] + r] = v;
      continue;
    }
    r -= nWh;
    if (r < nWh) {                      // WhT[g][j][i]
      int g = r / (HP * HP), j = (r / HP) % HP, i = r % HP;
      float v = 0.f;
      if (i < H && j < H) v = g < 2 ? Wg[(int64_t)(Din + i) * 2 * H + g * H + j] : Wc[(int64_t)(Din + i) * H + j];
      pw[a.WhT[k] + r] = v;
      continue;
    }
    r -= nWh;
    {                                   // WxT[n][i]
      int n = r / DinP, i = r % DinP, g = n / HP, j = n % HP;
      float v = 0.f;
      if (i < Din && j < H) v = g < 2 ? Wg[(int64_t)i * 2 * H + g * H + j] : Wc[(int64_t)i * H + j];
      pw[a.WxT[k] + r] = v;
    }
  }
}

void launch_pack(const Launch& L, const Dims& d, const ParamLayout& pl, const PackLayout& pk, const float* params,
                 float* pw, cudaStream_t st) {
  PackArgs a; memset(&a, 0, sizeof(a));
  a.L = d.L; a.H = d.H;
  for (int k = 0; k < d.L; ++k) {
    a.Din[k] = d.Din[k]; a.DinP[k] = d.DinP[k];
    a.Wg[k] = pl.Wg[k]; a.bg[k] = pl.bg[k]; a.Wc[k] = pl.Wc[k]; a.bc[k] = pl.bc[k];
    a.Wx[k] = pk.Wx[k]; a.bx[k] = pk.bx[k]; a.Wh[k] = pk.Wh[k]; a.WhT[k] = pk.WhT[k]; a.WxT[k] = pk.WxT[k];
  }
  pack_kernel<<<dim3(8, d.L), 256, 0, st>>>(a, params, pw);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
// C[M,N] = A[M,K] * W[K,N] + bias.   TM rows per tile, 8 warps x 8 rows, lane owns columns lane+32c.
// ---------------------------------------------------------------------------------------------
constexpr int NN_TM = 64;

template <int NC>
__global__ void __launch_bounds__(256)
gemm_nn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, const float* __restrict__ bias,
               float* __restrict__ C, int64_t M, int N, int K) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;               // [K][N]
  float* As = smem + K * N;       // [NN_TM][K]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < K * N; e += 256) Ws[e] = W[e];
  float bcol[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = lane + 32 * c;
    bcol[c] = (bias != nullptr && col < N) ? bias[col] : 0.f;
  }
  const int K4 = K >> 2;
  const int64_t tiles = (M + NN_TM - 1) / NN_TM;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t m0 = tile * NN_TM;
    __syncthreads();               // previous tile's As fully consumed (and Ws visible on first pass)
    for (int e = tid; e < NN_TM * K4; e += 256) {
      int r = e / K4, kq = e % K4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = ldg_nc_f4(reinterpret_cast<const float4*>(A + (m0 + r) * lda) + kq);
      reinterpret_cast<float4*>(As)[r * K4 + kq] = v;
    }
    __syncthreads();
    float acc[8][NC];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[r][c] = bcol[c];
    const float4* Arow = reinterpret_cast<const float4*>(As) + (warp * 8) * K4;
    for (int kq = 0; kq < K4; ++kq) {
      float4 a[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = Arow[r * K4 + kq];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          int col = lane + 32 * c;
          w[c] = col < N ? Ws[(kq * 4 + q) * N + col] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float av = q == 0 ? a[r].x : (q == 1 ? a[r].y : (q == 2 ? a[r].z : a[r].w));
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int64_t m = m0 + warp * 8 + r;
      if (m < M) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          int col = lane + 32 * c;
          if (col < N) C[m * N + col] = acc[r][c];
        }
      }
    }
  }
}

void launch_gemm_nn(const Launch& L, const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t M,
                    int N, int K, cudaStream_t st) {
  size_t smem = (size_t)(K * N + NN_TM * K) * sizeof(float);
  int64_t tiles = (M + NN_TM - 1) / NN_TM;
  int grid = (int)(tiles < (int64_t)L.sms * 2 ? tiles : (int64_t)L.sms * 2);
  int nc = (N + 31) / 32;
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, 256, smem, st>>>(A, lda, W, bias, C, M, N, K);
  };
  if (nc <= 1) go(gemm_nn_kernel<1>);
  else if (nc == 2) go(gemm_nn_kernel<2>);
  else if (nc == 3) go(gemm_nn_kernel<3>);
  else if (nc == 4) go(gemm_nn_kernel<4>);
  else go(gemm_nn_kernel<8>);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
// C[M,N] (+)= A[M, a0 : a0+K] * W[N,K]^T   (W row-major [N rows, ldw], i.e. the x rows of a TF kernel used transposed).
// Plain FFMA, one thread per output element group; only used where no tcgen05 instantiation exists (H = 64 dX).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_nt_kernel(const float* __restrict__ A, int64_t lda, int a0, const float* __restrict__ W, int64_t ldw, float* __restrict__ C,
               int64_t ldc, int64_t M, int N, int K, int accumulate) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                  // [N][K+1]
  for (int e = threadIdx.x; e < N * K; e += 256) Ws[(e / K) * (K + 1) + e % K] = W[(int64_t)(e / K) * ldw + e % K];
  __syncthreads();
  const int64_t total = M * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t m = e / N; const int n = (int)(e % N);
    const float* a = A + m * lda + a0;
    const float* w = Ws + n * (K + 1);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(a[k], w[k], acc);
    if (accumulate) C[m * ldc + n] += acc; else C[m * ldc + n] = acc;
  }
}

void launch_gemm_nt(const Launch& L, const float* A, int64_t lda, int a0, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                    int N, int K, bool accumulate, cudaStream_t st) {
  const size_t smem = (size_t)N * (K + 1) * sizeof(float);
  cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t blocks = (M * N + 255) / 256;
  if (blocks > (int64_t)L.sms * 8) blocks = (int64_t)L.sms * 8;
  if (blocks < 1) blocks = 1;
  gemm_nt_kernel<<<(unsigned)blocks, 256, smem, st>>>(A, lda, a0, W, ldw, C, ldc, M, N, K, accumulate ? 1 : 0);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
// Batched  C[I,N] += sum_m A[m,I] * Bm[m,N]  (weight gradients of the dense layers: reductions over
// the batch rows).  One launch serves a list of problems; each CTA owns one 64x64 output tile of one
// problem over one slice of the rows and finishes with atomics.  A == nullptr means a column of ones
// (bias gradients).
// ---------------------------------------------------------------------------------------------
constexpr int ATB_T = 64, ATB_MC = 32;

__global__ void __launch_bounds__(256)
gemm_atb_batch_kernel(const __grid_constant__ AtbBatch batch) {
  __shared__ __align__(16) float As[ATB_MC][ATB_T];
  __shared__ __align__(16) float Bs[ATB_MC][ATB_T];
  const int tid = threadIdx.x;
  int pi = 0;
  while (pi + 1 < batch.n && (int)blockIdx.x >= batch.p[pi + 1].block_begin) ++pi;
  const AtbProb& P = batch.p[pi];
  int local = blockIdx.x - P.block_begin;
  const int tiles_n = (P.N + ATB_T - 1) / ATB_T;
  const int tiles_i = (P.I + ATB_T - 1) / ATB_T;
  const int split = local / (tiles_i * tiles_n);
  local -= split * tiles_i * tiles_n;
  const int i0 = (local / tiles_n) * ATB_T, n0 = (local % tiles_n) * ATB_T;
  const int64_t mbeg = (int64_t)split * P.rows_per_split;
  const int64_t mend = mbeg + P.rows_per_split < P.M ? mbeg + P.rows_per_split : P.M;
  const float* __restrict__ A = P.A;
  const float* __restrict__ Bm = P.Bm;
  const int I = P.I, N = P.N;
  const int64_t lda = P.lda, ldb = P.ldb;
  const int ti = tid >> 4, tn = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t mc = mbeg; mc < mend; mc += ATB_MC) {
    __syncthreads();
    for (int e = tid; e < ATB_MC * ATB_T; e += 256) {
      int r = e / ATB_T, c = e % ATB_T;
      int64_t m = mc + r;
      float av = 0.f, bv = 0.f;
      if (m < mend) {
        if (i0 + c < I) av = A ? __ldg(A + m * lda + i0 + c) : 1.0f;
        if (n0 + c < N) bv = __ldg(Bm + m * ldb + n0 + c);
      }
      As[r][c] = av;
      Bs[r][c] = bv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < ATB_MC; ++r) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[r][ti * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[r][tn * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int i = i0 + ti * 4 + a;
    if (i >= I) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int n = n0 + tn * 4 + b;
      if (n < N) atomicAdd(P.C + (int64_t)i * P.ldc + n, acc[a][b]);
    }
  }
}

void atb_add(AtbBatch& batch, int sms, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
             int64_t M, int I, int N) {
  AtbProb& P = batch.p[batch.n];
  P.A = A; P.Bm = Bm; P.C = C; P.lda = lda; P.ldb = ldb; P.ldc = ldc; P.M = M; P.I = I; P.N = N;
  int tiles = ((I + ATB_T - 1) / ATB_T) * ((N + ATB_T - 1) / ATB_T);
  int64_t chunks = (M + ATB_MC - 1) / ATB_MC;
  int64_t want = sms / (4 * tiles);             // a batch holds many problems; keep each modest
  if (want < 1) want = 1;
  int64_t splits = chunks < want ? chunks : want;
  P.rows_per_split = ((chunks + splits - 1) / splits) * ATB_MC;
  splits = (M + P.rows_per_split - 1) / P.rows_per_split;
  P.block_begin = batch.blocks;
  batch.blocks += (int)splits * tiles;
  ++batch.n;
}

void launch_atb_batch(const Launch& L, const AtbBatch& batch, cudaStream_t st) {
  if (batch.n == 0) return;
  gemm_atb_batch_kernel<<<batch.blocks, 256, 0, st>>>(batch);
  ++*L.counter;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void clip_adam_one(float& var, float g, float& m, float& v, float lr_t, float b1, float b2, float eps,
                                              float clip) {
  g = fminf(fmaxf(g, -clip), clip);                        // tf.clip_by_value, code/hpmn.py:212
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  var = var - lr_t * m / (sqrtf(v) + eps);                 // TF1.4 ApplyAdam
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ var, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr_t, float b1, float b2, float eps, float clip) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = var[i], mi = m[i], vi = v[i];
    clip_adam_one(w, grad[i], mi, vi, lr_t, b1, b2, eps, clip);
    m[i] = mi; v[i] = vi; var[i] = w;
  }
}

// The same update, 2 x 16 bytes per array per thread and iteration: the sweep is HBM-bound (4 reads + 3 writes per element) and
// the scalar loop (4 bytes per load) left a third of the bandwidth on the table.  Element-wise identical arithmetic.
__global__ void __launch_bounds__(256)
clip_adam_kernel_v4(float4* __restrict__ var, const float4* __restrict__ grad, float4* __restrict__ m, float4* __restrict__ v,
                    int64_t n4, float lr_t, float b1, float b2, float eps, float clip) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const int64_t j = i + stride;
    const bool two = j < n4;
    float4 g0 = grad[i], w0 = var[i], m0 = m[i], v0 = v[i];
    float4 g1 = g0, w1 = w0, m1 = m0, v1 = v0;
    if (two) { g1 = grad[j]; w1 = var[j]; m1 = m[j]; v1 = v[j]; }
    clip_adam_one(w0.x, g0.x, m0.x, v0.x, lr_t, b1, b2, eps, clip);
    clip_adam_one(w0.y, g0.y, m0.y, v0.y, lr_t, b1, b2, eps, clip);
    clip_adam_one(w0.z, g0.z, m0.z, v0.z, lr_t, b1, b2, eps, clip);
    clip_adam_one(w0.w, g0.w, m0.w, v0.w, lr_t, b1, b2, eps, clip);
    m[i] = m0; v[i] = v0; var[i] = w0;
    if (two) {
      clip_adam_one(w1.x, g1.x, m1.x, v1.x, lr_t, b1, b2, eps, clip);
      clip_adam_one(w1.y, g1.y, m1.y, v1.y, lr_t, b1, b2, eps, clip);
      clip_adam_one(w1.z, g1.z, m1.z, v1.z, lr_t, b1, b2, eps, clip);
      clip_adam_one(w1.w, g1.w, m1.w, v1.w, lr_t, b1, b2, eps, clip);
      m[j] = m1; v[j] = v1; var[j] = w1;
    }
  }
}

void launch_clip_adam(const Launch& L, float* var, const float* grad, float* m, float* v, int64_t n, float lr_t, float b1,
                      float b2, float eps, float clip, cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(var) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  if (n4 > 0) {
    int64_t blocks = (n4 + 511) / 512;
    if (blocks > (int64_t)L.sms * 8) blocks = (int64_t)L.sms * 8;
    clip_adam_kernel_v4<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float4*>(var), reinterpret_cast<const float4*>(grad),
                                                          reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4, lr_t, b1,
                                                          b2, eps, clip);
    ++*L.counter;
  }
  const int64_t done = n4 * 4;
  if (done < n) {
    int64_t blocks = (n - done + 255) / 256;
    if (blocks > (int64_t)L.sms * 8) blocks = (int64_t)L.sms * 8;
    clip_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(var + done, grad + done, m + done, v + done, n - done, lr_t, b1, b2, eps, clip);
    ++*L.counter;
  }
}

__global__ void __launch_bounds__(256)
axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fmaf(a, x[i], y[i]);
}

void launch_axpy(const Launch& L, float* y, const float* x, float a, int64_t n, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)L.sms * 8) blocks = (int64_t)L.sms * 8;
  if (blocks < 1) blocks = 1;
  axpy_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, x, a, n);
  ++*L.counter;
}

// Zeroing that is meant to run BESIDE the forward recurrence: a small fixed grid (the wavefront kernel leaves 20 SMs idle;
// cudaMemsetAsync would launch a machine-wide grid that the block scheduler serialises against the kernels around it).
__global__ void __launch_bounds__(512) zero_kernel(float4* p4, int64_t n4, float* tail, int ntail) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) p4[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}
void launch_zero(const Launch& L, void* ptr, size_t bytes, int ctas, cudaStream_t st) {
  if (!ptr || bytes == 0) return;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (bytes & 3)) { cudaMemsetAsync(ptr, 0, bytes, st); return; }
  const int64_t n4 = (int64_t)(bytes / 16);
  const int ntail = (int)((bytes % 16) / 4);
  zero_kernel<<<ctas < 1 ? 1 : ctas, 512, 0, st>>>(reinterpret_cast<float4*>(ptr), n4, reinterpret_cast<float*>(ptr) + n4 * 4, ntail);
  ++*L.counter;
}

// loss = logloss + memory_reg * covreg   (code/hpmn.py:202-207; the l2 term is added by the caller)
__global__ void finish_scalars_kernel(float* s, float memory_reg, const __grid_constant__ OutCopies c) {
  if (blockIdx.x == 0 && threadIdx.x == 0) s[HPMN_S_LOSS] = s[HPMN_S_LOGLOSS] + memory_reg * s[HPMN_S_COVREG];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    for (int64_t i = i0; i < c.n[k]; i += stride) c.dst[k][i] = c.src[k][i];
}
void launch_finish_scalars(const Launch& L, float* scalars, float memory_reg, const OutCopies* copies, cudaStream_t st) {
  OutCopies c; memset(&c, 0, sizeof(c));
  int64_t total = 0;
  if (copies) { c = *copies; for (int k = 0; k < 4; ++k) total += c.n[k]; }
  int64_t blocks = (total + 1023) / 1024;
  if (blocks < 1) blocks = 1;
  if (blocks > (int64_t)L.sms * 4) blocks = (int64_t)L.sms * 4;
  finish_scalars_kernel<<<(int)blocks, 256, 0, st>>>(scalars, memory_reg, c);
  ++*L.counter;
}

}  // namespace hpmn
