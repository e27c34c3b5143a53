// api.cu -- the C ABI of libhpmn_b200.so (include/hpmn_b200.h) and the step orchestration.
// Everything the reference does inside one `sess.run` (/root/reference/code/hpmn.py:336,365,482,511) --
// minus the Adam apply, which is hpmn_clip_adam -- is enqueued here on the caller's stream.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include <functional>

#include "common.cuh"

using namespace hpmn;

#define HPMN_MAX_GROUPS 4

struct hpmn_ctx {
  bool wave_now;    // decided per step: the wavefront kernels are used only when the whole batch is one wave of CTAs
  bool tc_now;      // decided per step on the WHOLE batch (row groups inherit it): memory on the tensor-core recurrence
  bool use_wave;    // fused wavefront kernels for the recurrence (HPMN_NO_WAVE=1 selects the layer-by-layer kernels)
  bool use_tc;      // tcgen05 path for the dense (non-recurrent) GEMMs; HPMN_NO_TC=1 selects the FFMA kernels
  int tcrec_mode;   // tensor-core recurrence (tcrec.cu): -1 auto (H > 32, or B >= tcrec_min_b), 0 never, 1 always (HPMN_TCREC)
  int tcrec_min_b;  // HPMN_TCREC_MIN_B
  int device;
  int sms;
  int64_t launches;
  char err[512];
  float* scratch;   // 256 B of device memory (sink for the id-range flag of the granular gather)
  // side stream: weight-gradient reductions run here, concurrently with the latency-bound recurrent chain on the
  // caller's stream (forked / joined with events, so the caller still sees plain stream order)
  cudaStream_t side;
  cudaStream_t comm;        // caller's collective stream (hpmn_set_comm_stream): told when the table gradient is final
  cudaEvent_t ev_dtable;
  cudaEvent_t ev_fork[HPMN_MAX_LAYERS + 2];
  cudaEvent_t ev_join, ev_zero;
  bool overlap, zero_pending;
  int wave_ctas;            // CTAs of the wavefront kernels for the step being queued (one SM each)
  void* zero_late; size_t zero_late_bytes; void* zero_grads; size_t zero_grads_bytes;   // buffers zero_behind_projection() clears
  void* fin_d2h_dst; size_t fin_d2h_bytes; cudaEvent_t fin_d2h_ev; bool fin_d2h_done;   // host path: result block copied out early
  OutCopies fin_oc; float* fin_scalars; float fin_mreg; bool fin_copies, fin_early;   // the step's finish kernel (see run_step)
  bool fuse_mid;            // training step: attention + head, forward and backward, as one kernel (mid.cu); HPMN_NO_FUSE_MID=1 disables
  bool fuse_now;            // ... for the step being queued
  bool dtable_late;         // this step touches dtable after the scatter of the whole-batch chain (row groups, l2 term)
  // feed double buffering (hpmn_prefetch_host): copy stream, completion event, what is staged where
  cudaStream_t copy;
  cudaEvent_t ev_copy, ev_consumed[2];
  // results of a *_host step: event behind its D2H copies, keyed by the host scalars pointer (two steps may be in flight)
  cudaEvent_t ev_out[2]; const void* out_key[2]; int out_next;
  const void* staged_ids; const void* staged_labels; int staged_slot, staged_B, cur_slot; bool consumed_valid[2];
  // row groups: the batch is cut into `groups` independent row ranges, each running its whole fwd+bwd chain on its own
  // stream, so one group's dense kernels fill the SMs while another group sits in its latency-bound recurrence
  int groups, group_min_rows;
  cudaStream_t gstream[HPMN_MAX_GROUPS];
  cudaEvent_t ev_gstart, ev_gdone[HPMN_MAX_GROUPS];
  // profiling (CUDA events on the caller's stream around each kernel family)
  bool profile;
  std::vector<cudaEvent_t> pool;
  size_t pool_used;
  struct Span { int fam; cudaEvent_t a, b; };
  std::vector<Span> spans;
  double ms[HPMN_K_COUNT];
  int64_t calls[HPMN_K_COUNT];
};

static char g_create_err[512] = "";


static int fail(hpmn_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_create_err;
  va_list ap; va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(ctx, HPMN_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static int check_launch(hpmn_ctx* ctx, const char* where) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, HPMN_ECUDA, "%s: kernel launch failed: %s", where, cudaGetErrorString(e));
  }
  return HPMN_OK;
}

// ---- profiling helpers ----------------------------------------------------------------------
static void prof_flush(hpmn_ctx* ctx) {
  for (auto& s : ctx->spans) {
    float ms = 0.f;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      ctx->ms[s.fam] += ms;
      ctx->calls[s.fam] += 1;
    }
  }
  ctx->spans.clear();
  ctx->pool_used = 0;
}

struct Bracket {          // RAII: records an event pair around a kernel family when profiling is on
  hpmn_ctx* ctx; cudaStream_t st; int fam; cudaEvent_t a, b; bool on;
  Bracket(hpmn_ctx* c, cudaStream_t s, int f) : ctx(c), st(s), fam(f), on(false) {
    if (!c->profile) return;
    if (c->pool_used + 2 > c->pool.size()) prof_flush(c);
    a = c->pool[c->pool_used++]; b = c->pool[c->pool_used++];
    cudaEventRecord(a, st);
    on = true;
  }
  ~Bracket() {
    if (!on) return;
    cudaEventRecord(b, st);
    ctx->spans.push_back({fam, a, b});
  }
};

static const char* kFamilyNames[HPMN_K_COUNT] = {"gather_fwd", "inproj_gemm", "rec_fwd", "attn_fwd", "head_fwd", "head_bwd",
                                                 "attn_bwd", "rec_bwd", "dx_gemm", "gru_wgrad", "scatter_add", "misc"};

// ---- shared plumbing ------------------------------------------------------------------------
// Workspace = header (whole-batch staging and per-row outputs) + G group regions (activations of one row group).
struct Hdr { size_t ids, labels, ids2, labels2, pred, logit, w_hop0, memory, scalars, out_end, total; };   // ids2/labels2: prefetch slot
static Hdr make_hdr(const Dims& d) {
  Hdr h; size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  h.ids = take((size_t)d.B * d.T * d.F * sizeof(int32_t));
  h.labels = take((size_t)d.B * sizeof(int32_t));
  h.ids2 = take((size_t)d.B * d.T * d.F * sizeof(int32_t));
  h.labels2 = take((size_t)d.B * sizeof(int32_t));
  // scalars | pred | logit | w_hop0 form one block: a host caller whose result buffers mirror it gets ONE D2H copy per step
  h.scalars = take(4 * sizeof(float));
  h.pred = take((size_t)d.B * sizeof(float));
  h.logit = take((size_t)d.B * sizeof(float));
  h.w_hop0 = take((size_t)d.B * d.L * sizeof(float));
  h.out_end = off;
  h.memory = take((size_t)d.B * d.L * d.H * sizeof(float));
  h.total = off;
  return h;
}

struct Plan {
  Dims d; ParamLayout pl; PackLayout pk; WsLayout wl; Hdr hdr;
  char* base;      // caller's workspace (header at offset 0)
  char* ws;        // this plan's group region
  int row0;        // first batch row of the group
  float* f(size_t off) const { return reinterpret_cast<float*>(ws + off); }
  float* hf(size_t off) const { return reinterpret_cast<float*>(base + off); }
  AttWs att() const { return AttWs{f(wl.att_q), f(wl.att_dq), f(wl.att_w), f(wl.att_ds), f(wl.att_inp), f(wl.att_z1),
                                   f(wl.att_dz1), f(wl.att_z2), f(wl.att_dz2)}; }
  HeadWs head() const { return HeadWs{f(wl.head_bn), f(wl.head_dbn), f(wl.head_dgt), f(wl.head_a1), f(wl.head_act1),
                                      f(wl.head_dl1), f(wl.head_a2), f(wl.head_act2), f(wl.head_dl2), f(wl.head_dlogit)}; }
};

static size_t group_stride(const hpmn_shape* s, int G) {
  hpmn_shape t = *s;
  t.B = (s->B + G - 1) / G;
  Dims d = make_dims(&t);
  return d.ok ? ((make_ws_layout(d).total + 255) & ~(size_t)255) : 0;
}

static int make_plan(hpmn_ctx* ctx, const hpmn_shape* s, void* workspace, Plan& p) {
  if (!ctx) return HPMN_EINVAL;
  p.d = make_dims(s);
  if (!p.d.ok) return fail(ctx, HPMN_EINVAL, "invalid hpmn_shape (need E%%4==0, H<=32 or H==64, L<=16, hops<=8, steps divisible by periods)");
  if (p.d.D > 64) return fail(ctx, HPMN_EINVAL, "F*E = %d > 64 is not supported by this build", p.d.D);
  for (int k = 0; k + 1 < p.d.L; ++k)
    if (p.d.P[k] > 16) return fail(ctx, HPMN_EINVAL, "period %d > 16 is not supported by this build", p.d.P[k]);
  p.pl = make_param_layout(p.d);
  p.pk = make_pack_layout(p.d);
  p.wl = make_ws_layout(p.d);
  p.hdr = make_hdr(p.d);
  p.base = static_cast<char*>(workspace);
  p.ws = p.base + p.hdr.total;
  p.row0 = 0;
  if (!workspace) return fail(ctx, HPMN_EINVAL, "workspace is NULL");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return fail(ctx, HPMN_EINVAL, "workspace must be 256-byte aligned");
  cudaSetDevice(ctx->device);
  return HPMN_OK;
}

// plan of row group g of G (rows [row0, row0+rows)) inside the whole-batch plan's workspace
static Plan group_plan(const Plan& whole, const hpmn_shape* s, int g, int G, int row0, int rows) {
  Plan gp = whole;
  hpmn_shape t = *s;
  t.B = rows;
  gp.d = make_dims(&t);
  gp.wl = make_ws_layout(gp.d);
  gp.ws = whole.base + whole.hdr.total + (size_t)g * group_stride(s, G);
  gp.row0 = row0;
  return gp;
}

static hpmn_hyper default_hyper() { hpmn_hyper h; memset(&h, 0, sizeof(h)); h.memory_reg = 1e-5f; h.keep_prob = 1.f; return h; }

// C = A*W (+bias): tcgen05 3xTF32 when an instantiation exists, fp32 FFMA otherwise
static void dense_gemm(hpmn_ctx* ctx, const Launch& L, const float* A, int64_t lda, const float* W, const float* bias, float* C,
                       int64_t M, int N, int K, cudaStream_t st) {
  if (ctx->use_tc && launch_tc_gemm_nn(L, A, lda, W, bias, C, M, N, K, st)) return;
  launch_gemm_nn(L, A, lda, W, bias, C, M, N, K, st);
}

// tensor-core recurrence: chosen when the hidden size needs it or the batch fills M = 128 tiles on most SMs
static bool want_tcrec(const hpmn_ctx* ctx, const Dims& d) {
  if (!tcrec_supported(d)) return false;
  if (d.H > HP) return true;
  if (ctx->tcrec_mode >= 0) return ctx->tcrec_mode != 0;
  return d.B >= ctx->tcrec_min_b;
}

// memory forward on tcgen05: pack, split x into the 3xTF32 halves, then one launch per layer
static bool run_memory_fwd_tc(hpmn_ctx* ctx, const Plan& p, const float* x, const float* params, float* memory, cudaStream_t st,
                              bool packed_old = false) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  const TcrLayout tl = make_tcr_layout(d);
  if (!tl.ok) return false;
  char* ws = p.ws + p.wl.tcr;
  { Bracket b(ctx, st, HPMN_K_MISC);
    launch_tcr_pack(L, d, p.pl, tl, params, ws, st);
    if (!packed_old && d.H <= HP) launch_pack(L, d, p.pl, p.pk, params, p.f(p.wl.pw), st); }   // dense kernels of the backward pass
  { Bracket b(ctx, st, HPMN_K_INPROJ);
    launch_tcr_split(L, x, reinterpret_cast<float*>(ws + tl.xh[0]), reinterpret_cast<float*>(ws + tl.xl[0]), (int64_t)d.B * d.S[0], d.D,
                     tl.DP[0], st); }
  Bracket b(ctx, st, HPMN_K_REC_FWD);
  for (int k = 0; k < d.L; ++k)
    if (!launch_tcrec_fwd(L, d, tl, k, ws, memory, st)) return false;
  return true;
}

// memory backward on tcgen05: per layer (top first) the recurrent adjoint -> da_k, then dX_k = da_k Wx_k^T as a dense GEMM
// (it is the dx_up of the layer below / the embedding gradient); the weight gradients of every layer in one launch at the end.
// H = 32 reuses the dense kernels of the wavefront path (same h|r|u|c and r|u|c row layouts); H = 64 uses the FFMA GEMMs.
static bool run_memory_bwd_tc(hpmn_ctx* ctx, const Plan& p, const float* x, const float* params, const float* dmemory, float* dx0,
                              float* grads, cudaStream_t st, bool packed = false) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  const TcrLayout tl = make_tcr_layout(d);
  if (!tl.ok) return false;
  char* ws = p.ws + p.wl.tcr;
  const int H = d.H;
  const bool small = H <= HP;                       // the tcgen05 dX / weight-gradient kernels cover H <= 32
  float* pw = p.f(p.wl.pw);
  if (!packed) {
    Bracket b(ctx, st, HPMN_K_MISC);
    launch_tcr_pack(L, d, p.pl, tl, params, ws, st);
    if (small) launch_pack(L, d, p.pl, p.pk, params, pw, st);
  }
  auto fp = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  // H = 64: the tcgen05 weight-gradient kernel covers input widths 32 / 48; otherwise the FFMA reductions need h_prev | r*h_prev rows
  const int dp0 = ((d.D + 15) / 16) * 16;
  const bool need_hr = !small && !(ctx->use_tc && (dp0 == 32 || dp0 == 48));
  for (int k = d.L - 1; k >= 0; --k) {
    const float* dx_up = k < d.L - 1 ? fp(tl.dx[k + 1]) : nullptr;
    { Bracket b(ctx, st, HPMN_K_REC_BWD);
      if (!launch_tcrec_bwd(L, d, tl, k, ws, dmemory, dx_up, need_hr, st)) return false; }
    { Bracket b(ctx, st, HPMN_K_DX);
      float* dxk = k == 0 ? dx0 : fp(tl.dx[k]);
      const int64_t rows = (int64_t)d.B * d.S[k];
      if (small) {
        dense_gemm(ctx, L, fp(tl.da[k]), G3, pw + p.pk.WxT[k], nullptr, dxk, rows, d.DinP[k], G3, st);
      } else if (!(ctx->use_tc && launch_tc_gemm_nn(L, fp(tl.da[k]), 3 * H, fp(tl.wxt[k]), nullptr, dxk, rows, d.Din[k], 3 * H, st))) {
        // no tcgen05 instantiation for this width: two FFMA GEMMs against the x rows of the TF-layout kernels, accumulated
        launch_gemm_nt(L, fp(tl.da[k]), 3 * H, 0, params + p.pl.Wg[k], 2 * H, dxk, d.Din[k], rows, d.Din[k], 2 * H, false, st);
        launch_gemm_nt(L, fp(tl.da[k]), 3 * H, 2 * H, params + p.pl.Wc[k], H, dxk, d.Din[k], rows, d.Din[k], H, true, st);
      } }
  }
  Bracket b(ctx, st, HPMN_K_WGRAD);
  if (small) {
    const float* xa[HPMN_MAX_LAYERS]; int64_t lx[HPMN_MAX_LAYERS]; const float* stp[HPMN_MAX_LAYERS]; const float* dap[HPMN_MAX_LAYERS];
    float *gWg[HPMN_MAX_LAYERS], *gbg[HPMN_MAX_LAYERS], *gWc[HPMN_MAX_LAYERS], *gbc[HPMN_MAX_LAYERS];
    for (int k = 0; k < d.L; ++k) {
      stp[k] = fp(tl.st[k]); dap[k] = fp(tl.da[k]);
      xa[k] = k == 0 ? x : fp(tl.st[k - 1]) + (int64_t)(d.P[k - 1] - 1) * ST;       // every p-th h row of the layer below
      lx[k] = k == 0 ? d.D : (int64_t)d.P[k - 1] * ST;
      gWg[k] = grads + p.pl.Wg[k]; gbg[k] = grads + p.pl.bg[k]; gWc[k] = grads + p.pl.Wc[k]; gbc[k] = grads + p.pl.bc[k];
    }
    if (!(ctx->use_tc && launch_tc_wgrad_all(L, d, xa, lx, stp, dap, gWg, gbg, gWc, gbc, st)))
      for (int k = 0; k < d.L; ++k) launch_gru_wgrad(L, d, k, xa[k], lx[k], stp[k], dap[k], gWg[k], gbg[k], gWc[k], gbc[k], st);
  } else {
    const float* xa[HPMN_MAX_LAYERS]; int64_t lx[HPMN_MAX_LAYERS]; const float* stp[HPMN_MAX_LAYERS]; const float* dap[HPMN_MAX_LAYERS];
    float *gWg[HPMN_MAX_LAYERS], *gbg[HPMN_MAX_LAYERS], *gWc[HPMN_MAX_LAYERS], *gbc[HPMN_MAX_LAYERS];
    for (int k = 0; k < d.L; ++k) {
      stp[k] = fp(tl.st[k]); dap[k] = fp(tl.da[k]);
      xa[k] = k == 0 ? x : fp(tl.st[k - 1]) + (int64_t)(d.P[k - 1] - 1) * 4 * H;
      lx[k] = k == 0 ? d.D : (int64_t)d.P[k - 1] * 4 * H;
      gWg[k] = grads + p.pl.Wg[k]; gbg[k] = grads + p.pl.bg[k]; gWc[k] = grads + p.pl.Wc[k]; gbc[k] = grads + p.pl.bc[k];
    }
    if (!(ctx->use_tc && launch_tc_wgrad_wide(L, d, xa, lx, stp, dap, gWg, gbg, gWc, gbc, st))) {
      // [x | h_prev]^T da_g, [x | r*h_prev]^T da_c, column sums: batched A^T B reductions (FFMA) over the rows of every layer
      AtbBatch batch; batch.n = 0; batch.blocks = 0;
      for (int k = 0; k < d.L; ++k) {
        const int64_t rows = (int64_t)d.B * d.S[k];
        const float* da = dap[k]; const float* hr = fp(tl.hr[k]);
        atb_add(batch, ctx->sms, xa[k], lx[k], da, 3 * H, gWg[k], 2 * H, rows, d.Din[k], 2 * H);
        atb_add(batch, ctx->sms, hr, 2 * H, da, 3 * H, gWg[k] + (int64_t)d.Din[k] * 2 * H, 2 * H, rows, H, 2 * H);
        atb_add(batch, ctx->sms, xa[k], lx[k], da + 2 * H, 3 * H, gWc[k], H, rows, d.Din[k], H);
        atb_add(batch, ctx->sms, hr + H, 2 * H, da + 2 * H, 3 * H, gWc[k] + (int64_t)d.Din[k] * H, H, rows, H, H);
        atb_add(batch, ctx->sms, nullptr, 0, da, 3 * H, gbg[k], 2 * H, rows, 1, 2 * H);
        atb_add(batch, ctx->sms, nullptr, 0, da + 2 * H, 3 * H, gbc[k], H, rows, 1, H);
        if (batch.n + 6 > ATB_MAX) { launch_atb_batch(L, batch, st); batch.n = 0; batch.blocks = 0; }
      }
      launch_atb_batch(L, batch, st);
    }
  }
  return true;
}

// memory forward: pack + per layer (projection GEMM, recurrence)
// Zero the gradient buffers of this step on the side stream, starting when everything queued on `st` so far is done.  Called
// right behind the layer-0 projection GEMM: the memsets then run beside the forward recurrence (which leaves 20 SMs and
// most of the HBM bandwidth idle) instead of in front of it.  The backward half waits for ev_zero.
static void zero_behind_projection(hpmn_ctx* ctx, cudaStream_t st) {
  if (!ctx->zero_grads && !ctx->zero_late) return;
  cudaEventRecord(ctx->ev_fork[HPMN_MAX_LAYERS + 1], st);
  cudaStreamWaitEvent(ctx->side, ctx->ev_fork[HPMN_MAX_LAYERS + 1], 0);
  Launch L{&ctx->launches, ctx->sms};
  static const int zctas = [] { const char* e = getenv("HPMN_ZERO_CTAS"); return e ? atoi(e) : 0; }();
  // 2 CTAs per SM the wavefront kernels leave idle (they need a whole SM per CTA: registers and shared memory), so that the
  // zeroing never sits on an SM the backward wavefront kernel is waiting for; with fewer than 8 idle SMs, or without the
  // wavefront kernels, whatever cudaMemsetAsync launches
  const int idle = ctx->wave_ctas > 0 ? ctx->sms - ctx->wave_ctas : 0;
  const int ctas = zctas > 0 ? zctas : (idle >= 8 ? 2 * idle : 0);
  if (ctx->zero_grads) cudaMemsetAsync(ctx->zero_grads, 0, ctx->zero_grads_bytes, ctx->side);
  if (ctx->zero_late) {
    if (ctas > 0) launch_zero(L, ctx->zero_late, ctx->zero_late_bytes, ctas, ctx->side);
    else cudaMemsetAsync(ctx->zero_late, 0, ctx->zero_late_bytes, ctx->side);
  }
  cudaEventRecord(ctx->ev_zero, ctx->side);
  ctx->zero_pending = true;
  ctx->zero_grads = nullptr; ctx->zero_late = nullptr;
}

static void run_memory_fwd(hpmn_ctx* ctx, const Plan& p, const float* x, const float* params, float* memory,
                           cudaStream_t st, bool packed = false) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  float* pw = p.f(p.wl.pw);
  if (!packed) { Bracket b(ctx, st, HPMN_K_MISC); launch_pack(L, d, p.pl, p.pk, params, pw, st); }
  if (ctx->use_wave && ctx->wave_now && d.L <= 10) {
    // layer-0 input projections (dense, tensor cores), then every layer of every sample as one wavefront kernel
    { Bracket b(ctx, st, HPMN_K_INPROJ);
      dense_gemm(ctx, L, x, d.D, pw + p.pk.Wx[0], pw + p.pk.bx[0], p.f(p.wl.proj[0]), (int64_t)d.B * d.S[0], G3, d.DinP[0], st); }
    zero_behind_projection(ctx, st);
    float* stp[HPMN_MAX_LAYERS];
    for (int k = 0; k < d.L; ++k) stp[k] = p.f(p.wl.st[k]);
    Bracket b(ctx, st, HPMN_K_REC_FWD);
    if (launch_wave_fwd(L, d, p.pk, p.f(p.wl.proj[0]), pw, stp, memory, st)) return;
  }
  for (int k = 0; k < d.L; ++k) {
    const float* A = k == 0 ? x : p.f(p.wl.st[k - 1]) + (int64_t)(d.P[k - 1] - 1) * ST;   // every p-th h row
    const int64_t lda = k == 0 ? d.D : (int64_t)d.P[k - 1] * ST;
    { Bracket b(ctx, st, HPMN_K_INPROJ);
      dense_gemm(ctx, L, A, lda, pw + p.pk.Wx[k], pw + p.pk.bx[k], p.f(p.wl.proj[k]), (int64_t)d.B * d.S[k], G3, d.DinP[k], st); }
    { Bracket b(ctx, st, HPMN_K_REC_FWD);
      launch_rec_fwd(L, d, k, p.f(p.wl.proj[k]), pw + p.pk.Wh[k], p.f(p.wl.st[k]), memory, st); }
  }
}

// memory backward: top layer first; da overwrites the projections.  The recurrent chain (rec_bwd -> dx GEMM -> next
// layer) stays on `st`; each layer's weight-gradient reduction only needs that layer's da and is forked to the side stream.
// after_dx (optional) is called once dX of layer 0 is queued on `st`: the caller's scatter then goes in FRONT of the weight-
// gradient kernel (launch order = dispatch order), so the table gradient is final ~60 us after the recurrence instead of
// ~175 us and a table all-reduce on another stream (hpmn_set_comm_stream) runs beside the weight-gradient reduction.
static void run_memory_bwd(hpmn_ctx* ctx, const Plan& p, const float* x, const float* dmemory, float* dx0, float* grads,
                           bool ov, cudaStream_t st, bool join = true, const std::function<void()>& after_dx = nullptr) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  float* pw = p.f(p.wl.pw);
  if (ctx->use_wave && ctx->wave_now && d.L <= 10) {
    // every layer's reverse-time recurrence as one wavefront kernel (dx of layers >= 1 handed down in-kernel); then the
    // dense work: layer-0 dX (feeds the embedding scatter) on `st`, all weight-gradient reductions on the side stream
    const float* stp[HPMN_MAX_LAYERS]; float* dap[HPMN_MAX_LAYERS];
    for (int k = 0; k < d.L; ++k) { stp[k] = p.f(p.wl.st[k]); dap[k] = p.f(p.wl.proj[k]); }
    bool done;
    { Bracket b(ctx, st, HPMN_K_REC_BWD);
      done = launch_wave_bwd(L, d, p.pk, pw, stp, dap, dmemory, st); }
    if (done) {
      cudaStream_t ws = st;
      if (ov) { cudaEventRecord(ctx->ev_fork[0], st); cudaStreamWaitEvent(ctx->side, ctx->ev_fork[0], 0); ws = ctx->side; }
      { Bracket b(ctx, st, HPMN_K_DX);
        dense_gemm(ctx, L, dap[0], G3, pw + p.pk.WxT[0], nullptr, dx0, (int64_t)d.B * d.S[0], d.DinP[0], G3, st); }
      if (after_dx) after_dx();
      // A collective stream is attached (multi-GPU): the table gradient must be final as early as possible and the exchange
      // kernels need SMs while the weight gradients are reduced.  Left alone, the persistent weight-gradient grid takes every
      // SM the moment the dX GEMM drains and the scatter waits 117 us behind it (tools/timeline.py).  So the weight-gradient
      // kernel is ordered behind the scatter and leaves HPMN_COMM_SMS (default 24) SMs to the exchange.
      Launch Lw = L;
      if (ov && ctx->comm && after_dx) {
        static const int reserve = [] { const char* e = getenv("HPMN_COMM_SMS"); return e ? atoi(e) : 24; }();
        cudaEventRecord(ctx->ev_fork[1], st);
        cudaStreamWaitEvent(ctx->side, ctx->ev_fork[1], 0);
        if (reserve > 0 && reserve < ctx->sms - 16) Lw.sms = ctx->sms - reserve;
      }
      { Bracket b(ctx, st, HPMN_K_WGRAD);
        const float* xa[HPMN_MAX_LAYERS]; int64_t lx[HPMN_MAX_LAYERS];
        float *gWg[HPMN_MAX_LAYERS], *gbg[HPMN_MAX_LAYERS], *gWc[HPMN_MAX_LAYERS], *gbc[HPMN_MAX_LAYERS];
        for (int k = 0; k < d.L; ++k) {
          xa[k] = k == 0 ? x : p.f(p.wl.st[k - 1]) + (int64_t)(d.P[k - 1] - 1) * ST;
          lx[k] = k == 0 ? d.D : (int64_t)d.P[k - 1] * ST;
          gWg[k] = grads + p.pl.Wg[k]; gbg[k] = grads + p.pl.bg[k]; gWc[k] = grads + p.pl.Wc[k]; gbc[k] = grads + p.pl.bc[k];
        }
        if (!(ctx->use_tc && launch_tc_wgrad_all(Lw, d, xa, lx, stp, dap, gWg, gbg, gWc, gbc, ws)))
          for (int k = 0; k < d.L; ++k)
            launch_gru_wgrad(L, d, k, xa[k], lx[k], stp[k], dap[k], gWg[k], gbg[k], gWc[k], gbc[k], ws);
      }
      if (ov && join) { cudaEventRecord(ctx->ev_join, ctx->side); cudaStreamWaitEvent(st, ctx->ev_join, 0); }
      return;
    }
  }
  for (int k = d.L - 1; k >= 0; --k) {
    float* da = p.f(p.wl.proj[k]);
    const float* dx_up = k < d.L - 1 ? p.f(p.wl.dxk[k + 1]) : nullptr;
    { Bracket b(ctx, st, HPMN_K_REC_BWD);
      launch_rec_bwd(L, d, k, p.f(p.wl.st[k]), pw + p.pk.WhT[k], dmemory, dx_up, da, st); }
    cudaStream_t ws = st;
    if (ov) { cudaEventRecord(ctx->ev_fork[k], st); cudaStreamWaitEvent(ctx->side, ctx->ev_fork[k], 0); ws = ctx->side; }
    const float* A = k == 0 ? x : p.f(p.wl.st[k - 1]) + (int64_t)(d.P[k - 1] - 1) * ST;   // every p-th h row
    const int64_t lda = k == 0 ? d.D : (int64_t)d.P[k - 1] * ST;
    { Bracket b(ctx, st, HPMN_K_WGRAD);
      if (!(ctx->use_tc && launch_tc_wgrad(L, d, k, A, lda, p.f(p.wl.st[k]), da, grads + p.pl.Wg[k], grads + p.pl.bg[k],
                                           grads + p.pl.Wc[k], grads + p.pl.bc[k], ws)))
        launch_gru_wgrad(L, d, k, A, lda, p.f(p.wl.st[k]), da, grads + p.pl.Wg[k], grads + p.pl.bg[k],
                         grads + p.pl.Wc[k], grads + p.pl.bc[k], ws); }
    float* dxk = k == 0 ? dx0 : p.f(p.wl.dxk[k]);
    { Bracket b(ctx, st, HPMN_K_DX);
      dense_gemm(ctx, L, da, G3, pw + p.pk.WxT[k], nullptr, dxk, (int64_t)d.B * d.S[k], d.DinP[k], G3, st); }
  }
  if (after_dx) after_dx();
  if (ov && join) { cudaEventRecord(ctx->ev_join, ctx->side); cudaStreamWaitEvent(st, ctx->ev_join, 0); }
}

// ---- sum of squares (only for l2_reg != 0) ---------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ v, int64_t n, float scale, float* out) {
  float s = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s = fmaf(v[i], v[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s * scale);
}

// =============================================================================================
extern "C" {

int hpmn_abi_version(void) { return HPMN_ABI_VERSION; }

int hpmn_create(hpmn_ctx** out, int device) {
  hpmn_ctx* ctx = nullptr;
  if (!out) return fail(nullptr, HPMN_EINVAL, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, HPMN_ECUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, HPMN_EINVAL, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, HPMN_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, HPMN_EARCH, "device %d is sm_%d%d; libhpmn_b200 is built for sm_100a only", device, prop.major, prop.minor);
  ctx = new (std::nothrow) hpmn_ctx();
  if (!ctx) return fail(nullptr, HPMN_ENOMEM, "out of host memory");
  ctx->device = device; ctx->sms = prop.multiProcessorCount; ctx->launches = 0; ctx->err[0] = 0;
  ctx->profile = false; ctx->pool_used = 0; ctx->wave_now = true; ctx->wave_ctas = 0;
  { const char* e_tc = getenv("HPMN_NO_TC"); ctx->use_tc = !(e_tc && e_tc[0] == '1'); }
  { const char* e_w = getenv("HPMN_NO_WAVE"); ctx->use_wave = !(e_w && e_w[0] == '1'); }
  { const char* e_t = getenv("HPMN_TCREC"); ctx->tcrec_mode = e_t ? atoi(e_t) : -1;
    const char* e_b = getenv("HPMN_TCREC_MIN_B"); ctx->tcrec_min_b = e_b ? atoi(e_b) : 8192; }
  memset(ctx->ms, 0, sizeof(ctx->ms)); memset(ctx->calls, 0, sizeof(ctx->calls));
  cudaSetDevice(device);
  e = cudaMalloc(&ctx->scratch, 256);
  if (e != cudaSuccess) { delete ctx; return fail(nullptr, HPMN_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
  cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
  for (auto& ev : ctx->ev_fork) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_zero, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_dtable, cudaEventDisableTiming);
  ctx->comm = nullptr;
  ctx->zero_pending = false;
  ctx->zero_late = nullptr; ctx->zero_grads = nullptr; ctx->zero_late_bytes = 0; ctx->zero_grads_bytes = 0;
  ctx->fin_d2h_dst = nullptr; ctx->fin_d2h_done = false; ctx->fin_early = false;
  cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
  for (auto& ev : ctx->ev_consumed) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  for (auto& ev : ctx->ev_out) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  ctx->out_key[0] = ctx->out_key[1] = nullptr; ctx->out_next = 0;
  ctx->staged_ids = nullptr; ctx->staged_labels = nullptr; ctx->staged_slot = 0; ctx->staged_B = 0; ctx->cur_slot = 0;
  ctx->consumed_valid[0] = ctx->consumed_valid[1] = false;
  { const char* e_ov = getenv("HPMN_NO_OVERLAP"); ctx->overlap = !(e_ov && e_ov[0] == '1'); }
  { const char* e_fm = getenv("HPMN_NO_FUSE_MID"); ctx->fuse_mid = !(e_fm && e_fm[0] == '1'); }
  for (auto& gs : ctx->gstream) cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking);
  for (auto& ev : ctx->ev_gdone) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_gstart, cudaEventDisableTiming);
  { const char* e_g = getenv("HPMN_GROUPS"); int g = e_g ? atoi(e_g) : 4; ctx->groups = g < 1 ? 1 : (g > HPMN_MAX_GROUPS ? HPMN_MAX_GROUPS : g);
    const char* e_m = getenv("HPMN_GROUP_MIN_ROWS"); ctx->group_min_rows = e_m ? atoi(e_m) : 256; if (ctx->group_min_rows < 16) ctx->group_min_rows = 16; }
  *out = ctx;
  return HPMN_OK;
}

void hpmn_destroy(hpmn_ctx* ctx) {
  if (!ctx) return;
  for (auto e : ctx->pool) cudaEventDestroy(e);
  for (auto& ev : ctx->ev_fork) cudaEventDestroy(ev);
  cudaEventDestroy(ctx->ev_join);
  cudaEventDestroy(ctx->ev_zero);
  cudaEventDestroy(ctx->ev_dtable);
  cudaEventDestroy(ctx->ev_copy); for (auto& ev : ctx->ev_consumed) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_out) cudaEventDestroy(ev);
  cudaStreamDestroy(ctx->copy);
  cudaStreamDestroy(ctx->side);
  for (auto& gs : ctx->gstream) cudaStreamDestroy(gs);
  for (auto& ev : ctx->ev_gdone) cudaEventDestroy(ev);
  cudaEventDestroy(ctx->ev_gstart);
  cudaFree(ctx->scratch);
  delete ctx;
}

const char* hpmn_last_error(hpmn_ctx* ctx) { return ctx ? ctx->err : g_create_err; }
int64_t hpmn_launch_count(hpmn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int hpmn_set_comm_stream(hpmn_ctx* ctx, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  ctx->comm = (cudaStream_t)stream;
  return HPMN_OK;
}

int hpmn_param_tensors(const hpmn_shape* s) {
  Dims d = make_dims(s);
  if (!d.ok) return HPMN_EINVAL;
  return make_param_layout(d).ntensors;
}
int64_t hpmn_param_count(const hpmn_shape* s) {
  Dims d = make_dims(s);
  if (!d.ok) return HPMN_EINVAL;
  return make_param_layout(d).total;
}
int hpmn_param_offsets(const hpmn_shape* s, int64_t* offsets, int64_t* sizes, int n) {
  Dims d = make_dims(s);
  if (!d.ok || !offsets || !sizes) return HPMN_EINVAL;
  int i = 0;
  ParamLayout pl = make_param_layout(d, [&](int64_t off, int64_t sz) { if (i < n) { offsets[i] = off; sizes[i] = sz; } ++i; });
  return pl.ntensors <= n ? pl.ntensors : HPMN_EINVAL;
}
size_t hpmn_workspace_bytes(const hpmn_shape* s, int for_bwd) {
  (void)for_bwd;   // forward keeps the same activations (eval and train share one layout)
  Dims d = make_dims(s);
  if (!d.ok) return 0;
  size_t best = 0;
  for (int G = 1; G <= HPMN_MAX_GROUPS; ++G) {
    size_t t = (size_t)G * group_stride(s, G);
    if (t > best) best = t;
  }
  return make_hdr(d).total + best;
}

const char* hpmn_kernel_family_name(int f) { return f >= 0 && f < HPMN_K_COUNT ? kFamilyNames[f] : "?"; }

int hpmn_profile_enable(hpmn_ctx* ctx, int on) {
  if (!ctx) return HPMN_EINVAL;
  cudaSetDevice(ctx->device);
  if (on && ctx->pool.empty()) {
    ctx->pool.resize(2048);
    for (auto& e : ctx->pool) CK(cudaEventCreate(&e));
  }
  if (!on) prof_flush(ctx);
  ctx->profile = on != 0;
  return HPMN_OK;
}
int hpmn_profile_read(hpmn_ctx* ctx, float* ms, int64_t* calls) {
  if (!ctx || !ms || !calls) return HPMN_EINVAL;
  prof_flush(ctx);
  for (int i = 0; i < HPMN_K_COUNT; ++i) { ms[i] = (float)ctx->ms[i]; calls[i] = ctx->calls[i]; ctx->ms[i] = 0; ctx->calls[i] = 0; }
  return HPMN_OK;
}

// ---- K1 / K5 ----------------------------------------------------------------------------------
int hpmn_gather_fwd(hpmn_ctx* ctx, const hpmn_shape* s, const int32_t* ids, const float* table, float* x, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  Dims d = make_dims(s);
  if (!d.ok) return fail(ctx, HPMN_EINVAL, "invalid hpmn_shape");
  if (!ids || !table || !x) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  cudaSetDevice(ctx->device);
  Launch L{&ctx->launches, ctx->sms};
  // out-of-range ids cannot be reported without a sync here: rows are zero-filled (see hpmn_step_host)
  launch_gather_fwd(L, d, s->mask_id0 != 0, s->front_pad, s->V, ids, table, x, ctx->scratch, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_gather_fwd");
}

int hpmn_gather_bwd(hpmn_ctx* ctx, const hpmn_shape* s, const int32_t* ids, const float* dx, const float* dlast,
                    float* dtable, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  Dims d = make_dims(s);
  if (!d.ok) return fail(ctx, HPMN_EINVAL, "invalid hpmn_shape");
  if (!ids || !dx || !dtable) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  cudaSetDevice(ctx->device);
  Launch L{&ctx->launches, ctx->sms};
  launch_gather_bwd(L, d, s->mask_id0 != 0, s->front_pad, s->last_offset, s->V, ids, dx, dlast, dtable, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_gather_bwd");
}

int hpmn_gather_bwd_multi(hpmn_ctx* ctx, const hpmn_shape* s, int nsrc, const int32_t* const* ids, const float* const* dx,
                          const float* const* dlast, float* dtable, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  Dims d = make_dims(s);
  if (!d.ok) return fail(ctx, HPMN_EINVAL, "invalid hpmn_shape");
  if (nsrc < 1 || nsrc > 64 || !ids || !dx || !dtable) return fail(ctx, HPMN_EINVAL, "bad argument");
  for (int i = 0; i < nsrc; ++i) if (!ids[i] || !dx[i]) return fail(ctx, HPMN_EINVAL, "NULL source %d", i);
  cudaSetDevice(ctx->device);
  Launch L{&ctx->launches, ctx->sms};
  launch_gather_bwd_multi(L, d, s->mask_id0 != 0, s->front_pad, s->last_offset, s->V, nsrc, ids, dx, dlast, dtable, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_gather_bwd_multi");
}

// ---- K2 / K4 ----------------------------------------------------------------------------------
int hpmn_memory_fwd(hpmn_ctx* ctx, const hpmn_shape* s, const float* x, const float* params, float* memory,
                    void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!x || !params || !memory) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  { const int nspc = p.d.L <= 5 ? 2 : 1; ctx->wave_now = (p.d.B + nspc - 1) / nspc <= ctx->sms; }
  if (want_tcrec(ctx, p.d)) {
    if (!run_memory_fwd_tc(ctx, p, x, params, memory, (cudaStream_t)stream))
      return fail(ctx, HPMN_ECUDA, "tensor-core recurrence could not be launched (tensor map / shared memory)");
  } else {
    run_memory_fwd(ctx, p, x, params, memory, (cudaStream_t)stream);
  }
  return check_launch(ctx, "hpmn_memory_fwd");
}

int hpmn_memory_bwd(hpmn_ctx* ctx, const hpmn_shape* s, const float* x, const float* params, const float* dmemory,
                    float* dx, float* grads, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!x || !params || !dmemory || !dx || !grads) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  Launch L{&ctx->launches, ctx->sms};
  if (p.d.H <= HP) launch_pack(L, p.d, p.pl, p.pk, params, p.f(p.wl.pw), st);
  { const int nspc = p.d.L <= 5 ? 2 : 1; ctx->wave_now = (p.d.B + nspc - 1) / nspc <= ctx->sms; }
  if (want_tcrec(ctx, p.d)) {
    if (!run_memory_bwd_tc(ctx, p, x, params, dmemory, dx, grads, st))
      return fail(ctx, HPMN_ECUDA, "tensor-core recurrence could not be launched (tensor map / shared memory)");
    return check_launch(ctx, "hpmn_memory_bwd");
  }
  run_memory_bwd(ctx, p, x, dmemory, dx, grads, ctx->overlap && !ctx->profile, st);
  return check_launch(ctx, "hpmn_memory_bwd");
}

// ---- K3 ---------------------------------------------------------------------------------------
int hpmn_attn_fwd(hpmn_ctx* ctx, const hpmn_shape* s, const float* memory, const float* x, const float* params,
                  float* repre, float* w_hop0, float* scalars, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!memory || !x || !params || !repre || !w_hop0 || !scalars) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  Launch L{&ctx->launches, ctx->sms};
  launch_attn_fwd(L, p.d, p.pl, s->last_offset, memory, x, params, repre, w_hop0, scalars, p.att(), (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_attn_fwd");
}

int hpmn_attn_bwd(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const float* memory, const float* x,
                  const float* params, const float* drepre, float* dmemory, float* dlast, float* grads, void* workspace,
                  void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!hy || !memory || !x || !params || !drepre || !dmemory || !dlast || !grads) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  Launch L{&ctx->launches, ctx->sms};
  AtbBatch batch; batch.n = 0; batch.blocks = 0;
  launch_attn_bwd(L, p.d, p.pl, s->last_offset, hy->memory_reg, memory, x, params, drepre, dmemory, dlast, grads, p.att(),
                  batch, (cudaStream_t)stream);
  launch_atb_batch(L, batch, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_attn_bwd");
}

// ---- head -------------------------------------------------------------------------------------
int hpmn_head_fwd(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const float* repre, const int32_t* labels,
                  const float* params, float* pred, float* logit, float* scalars, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!hy || !repre || !labels || !params || !pred || !logit || !scalars) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  Launch L{&ctx->launches, ctx->sms};
  // pred / logit are staged in the workspace (hpmn_head_bwd reads pred from there), then copied out
  launch_head_fwd(L, p.d, p.pl, *hy, 0, repre, labels, params, p.hf(p.hdr.pred), p.hf(p.hdr.logit), scalars, p.head(), (cudaStream_t)stream);
  CK(cudaMemcpyAsync(pred, p.hf(p.hdr.pred), (size_t)p.d.B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  CK(cudaMemcpyAsync(logit, p.hf(p.hdr.logit), (size_t)p.d.B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return check_launch(ctx, "hpmn_head_fwd");
}

int hpmn_head_bwd(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const float* repre, const int32_t* labels,
                  const float* params, float* drepre, float* grads, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!hy || !repre || !labels || !params || !drepre || !grads) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  Launch L{&ctx->launches, ctx->sms};
  AtbBatch batch; batch.n = 0; batch.blocks = 0;
  // pred of the preceding hpmn_head_fwd on this workspace
  launch_head_bwd(L, p.d, p.pl, *hy, 0, repre, labels, params, p.hf(p.hdr.pred), drepre, grads, p.head(), batch, (cudaStream_t)stream);
  launch_atb_batch(L, batch, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_head_bwd");
}

// ---- whole path -------------------------------------------------------------------------------
// forward chain of one row group (rows [p.row0, p.row0 + p.d.B)) on stream st
static void fwd_rows(hpmn_ctx* ctx, const Plan& p, const hpmn_shape* s, const hpmn_hyper& hy, const int32_t* ids,
                     const int32_t* labels, const float* params, const float* table, float* scalars, bool side_ok,
                     cudaStream_t st) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  const int r0 = p.row0;
  float* x = p.f(p.wl.x);
  float* memory = p.hf(p.hdr.memory) + (int64_t)r0 * d.L * d.H;
  const bool ov = side_ok;      // run_step already queued the weight repacking on the side stream (event ev_join)
  { Bracket b(ctx, st, HPMN_K_GATHER);
    launch_gather_fwd(L, d, s->mask_id0 != 0, s->front_pad, s->V, ids + (int64_t)r0 * d.T * d.F, table, x,
                      scalars + HPMN_S_IDERR, st); }
  if (ov) cudaStreamWaitEvent(st, ctx->ev_join, 0);
  if (!(ctx->tc_now && run_memory_fwd_tc(ctx, p, x, params, memory, st, ov)))
    run_memory_fwd(ctx, p, x, params, memory, st, ov);
  zero_behind_projection(ctx, st);     // no-op when run_memory_fwd already queued it behind the projection GEMM
  if (ctx->fuse_now) return;    // bwd_rows runs attention + head, both directions, as one kernel
  { Bracket b(ctx, st, HPMN_K_ATTN_FWD);
    launch_attn_fwd(L, d, p.pl, s->last_offset, memory, x, params, p.f(p.wl.repre), p.hf(p.hdr.w_hop0) + (int64_t)r0 * d.L,
                    scalars, p.att(), st); }
  { Bracket b(ctx, st, HPMN_K_HEAD_FWD);
    launch_head_fwd(L, d, p.pl, hy, r0, p.f(p.wl.repre), labels + r0, params, p.hf(p.hdr.pred) + r0, p.hf(p.hdr.logit) + r0,
                    scalars, p.head(), st); }
}

// backward chain of one row group; weight gradients accumulate into grads / dtable with atomics
static void bwd_rows(hpmn_ctx* ctx, const Plan& p, const hpmn_shape* s, const hpmn_hyper& hy, const int32_t* ids,
                     const int32_t* labels, const float* params, float* grads, float* dtable, float* scalars, bool side_ok,
                     cudaStream_t st) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  const int r0 = p.row0;
  float* x = p.f(p.wl.x);
  const float* memory = p.hf(p.hdr.memory) + (int64_t)r0 * d.L * d.H;
  const bool ov = side_ok && ctx->overlap && !ctx->profile;
  // the tensor-core recurrence keeps its own activations (they alias the wavefront path's: one of the two runs per call)
  const bool tc = ctx->tc_now;
  float* dx0 = tc ? reinterpret_cast<float*>(p.ws + p.wl.tcr + make_tcr_layout(d).dx[0]) : p.f(p.wl.dxk[0]);
  AtbBatch batch; batch.n = 0; batch.blocks = 0;
  if (!ctx->fuse_now) {
    Bracket b(ctx, st, HPMN_K_HEAD_BWD);
    launch_head_bwd(L, d, p.pl, hy, r0, p.f(p.wl.repre), labels + r0, params, p.hf(p.hdr.pred) + r0, p.f(p.wl.drepre), grads,
                    p.head(), batch, st); }
  { Bracket b(ctx, st, ctx->fuse_now ? HPMN_K_ATTN_FWD : HPMN_K_ATTN_BWD);   // fused: the whole section is booked on ATTN_FWD
    if (ctx->fuse_now)
      launch_mid_fused(L, d, p.pl, hy, s->last_offset, r0, memory, x, params, labels + r0, p.f(p.wl.repre),
                       p.hf(p.hdr.w_hop0) + (int64_t)r0 * d.L, p.hf(p.hdr.pred) + r0, p.hf(p.hdr.logit) + r0, scalars,
                       p.f(p.wl.drepre), p.f(p.wl.dmemory), p.f(p.wl.dlast), grads, p.att(), p.head(), batch, st);
    else
      launch_attn_bwd(L, d, p.pl, s->last_offset, hy.memory_reg, memory, x, params, p.f(p.wl.drepre), p.f(p.wl.dmemory),
                      p.f(p.wl.dlast), grads, p.att(), batch, st);
    if (ov) {     // joined at the end of run_memory_bwd
      cudaEventRecord(ctx->ev_fork[HPMN_MAX_LAYERS], st);
      cudaStreamWaitEvent(ctx->side, ctx->ev_fork[HPMN_MAX_LAYERS], 0);
      launch_atb_batch(L, batch, ctx->side);
      if (ctx->fin_early) {    // loss + output copies only need what the attention / head section produced: off the critical path
        launch_finish_scalars(L, ctx->fin_scalars, ctx->fin_mreg, ctx->fin_copies ? &ctx->fin_oc : nullptr, ctx->side);
        ctx->fin_early = false;
        if (ctx->fin_d2h_dst) {   // host path: the result block (scalars | pred | logit | w_hop0) is final here -- the host gets it
                                  // while the backward recurrence runs instead of behind a D2H at the end of the step
          cudaMemcpyAsync(ctx->fin_d2h_dst, ctx->fin_scalars, ctx->fin_d2h_bytes, cudaMemcpyDeviceToHost, ctx->side);
          cudaEventRecord(ctx->fin_d2h_ev, ctx->side);
          ctx->fin_d2h_done = true;
        }
      }
    } else {
      launch_atb_batch(L, batch, st);
    } }
  // the weight-gradient reductions on the side stream are joined behind the scatter: the dX GEMM and the scatter (61 us)
  // run beside the 114 us weight-gradient kernel instead of in front of / behind it
  auto scatter = [&]() {
    if (ctx->zero_pending) { cudaStreamWaitEvent(st, ctx->ev_zero, 0); ctx->zero_pending = false; }
    { Bracket b(ctx, st, HPMN_K_SCATTER);
      launch_gather_bwd(L, d, s->mask_id0 != 0, s->front_pad, s->last_offset, s->V, ids + (int64_t)r0 * d.T * d.F,
                        dx0, p.f(p.wl.dlast), dtable, st); }
    if (side_ok && ctx->comm && !ctx->dtable_late) {     // whole-batch step: the table gradient is final from here on
      cudaEventRecord(ctx->ev_dtable, st);
      cudaStreamWaitEvent(ctx->comm, ctx->ev_dtable, 0);
    }
  };
  if (tc) {
    // tensor-core recurrence: everything on `st` (its dense kernels fill the machine at the batch sizes that select it)
    if (ctx->zero_pending) { cudaStreamWaitEvent(st, ctx->ev_zero, 0); ctx->zero_pending = false; }   // grads are written on `st`
    run_memory_bwd_tc(ctx, p, x, params, p.f(p.wl.dmemory), dx0, grads, st, true);
    scatter();
  } else {
    run_memory_bwd(ctx, p, x, p.f(p.wl.dmemory), dx0, grads, ov, st, false, scatter);
  }
  if (ov) { cudaEventRecord(ctx->ev_join, ctx->side); cudaStreamWaitEvent(st, ctx->ev_join, 0); }
}

// One step: prologue on `st`, G concurrent row-group chains, epilogue on `st`.  scalars: device float[4].
static int run_step(hpmn_ctx* ctx, const Plan& p, const hpmn_shape* s, hpmn_hyper hy, const int32_t* ids, const int32_t* labels,
                    const float* params, const float* table, float* grads, float* dtable, int zero_dtable, bool with_backward,
                    float* scalars, cudaStream_t st, const hpmn_outputs* out_dev = nullptr) {
  Launch L{&ctx->launches, ctx->sms};
  const Dims& d = p.d;
  if (d.H > HP && !tcrec_supported(d))
    return fail(ctx, HPMN_EINVAL, "hidden_size %d: this build covers H <= 32 and H = 64 (F*E <= 64)", d.H);
  if (hy.loss_batch <= 0) hy.loss_batch = d.B;          // the groups must divide the log-loss by the whole batch
  // Wavefront kernels: 2 samples per CTA (1 for L > 5), one CTA per SM (shared memory).  They minimise the latency of
  // one wave; with more samples than one wave holds, the per-layer kernels (one warp per sample, ~12 resident per SM)
  // have the higher throughput (tools/microbench.py).
  { const int nspc = d.L <= 5 ? 2 : 1; ctx->wave_ctas = (d.B + nspc - 1) / nspc; ctx->wave_now = ctx->wave_ctas <= ctx->sms; }
  ctx->tc_now = want_tcrec(ctx, d);
  if (!(ctx->use_wave && ctx->wave_now && d.L <= 10) || ctx->tc_now) ctx->wave_ctas = 0;     // another memory path runs: no idle SMs to count on
  ctx->fuse_now = with_backward && ctx->fuse_mid;
  // co-running dense kernels steal issue slots from the latency-critical recurrent warps, so grouping only pays
  // once every group still fills the machine (measured: -4 % at B=256, +11 % at B=1024)
  int G = ctx->profile ? 1 : ctx->groups;
  while (G > 1 && d.B / G < ctx->group_min_rows) --G;
  ctx->zero_pending = false;
  // the comm stream is told "dtable is final" right behind the scatter only when nothing else writes dtable afterwards:
  // with row groups there are G scatters on G streams, and the l2 term adds l2_reg * table at the very end
  ctx->dtable_late = G > 1 || hy.l2_reg != 0.f;
  // the weight repacking does not depend on the gather: it runs beside it on the side stream (ahead of the table-gradient
  // zeroing queued there below) and is joined in front of the projection GEMM
  const bool side_pack = G == 1 && ctx->overlap && !ctx->profile;
  if (side_pack) {
    CK(cudaEventRecord(ctx->ev_fork[0], st));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork[0], 0));
    launch_pack(L, d, p.pl, p.pk, params, p.f(p.wl.pw), ctx->side);
    CK(cudaEventRecord(ctx->ev_join, ctx->side));
  }
  ctx->zero_late = nullptr; ctx->zero_late_bytes = 0; ctx->zero_grads = nullptr; ctx->zero_grads_bytes = 0;
  { Bracket b(ctx, st, HPMN_K_MISC);
    CK(cudaMemsetAsync(scalars, 0, 4 * sizeof(float), st));
    if (with_backward) {
      if (G == 1 && ctx->overlap && !ctx->profile) {
        // Neither gradient buffer is touched before the backward half: both are zeroed on the side stream once the forward
        // recurrence is under way (fwd_rows -> zero_behind_projection).  At the head of the step the 212 MB memset held
        // every SM while the projection GEMM was ready to run: 22 us of the critical path (tools/timeline.py).
        ctx->zero_grads = grads; ctx->zero_grads_bytes = (size_t)p.pl.total * sizeof(float);
        if (zero_dtable) { ctx->zero_late = dtable; ctx->zero_late_bytes = (size_t)s->V * d.E * sizeof(float); }
      } else {
        CK(cudaMemsetAsync(grads, 0, (size_t)p.pl.total * sizeof(float), st));
        if (zero_dtable) CK(cudaMemsetAsync(dtable, 0, (size_t)s->V * d.E * sizeof(float), st));
      }
    } }
  // The finish kernel (loss = logloss + memory_reg * covreg, and the per-row results copied to the caller's device buffers: one
  // launch instead of a cudaMemcpyAsync per output) needs nothing the backward recurrence produces: in the training step it
  // goes to the side stream behind the attention / head weight gradients (bwd_rows) and is joined with them.
  memset(&ctx->fin_oc, 0, sizeof(ctx->fin_oc));
  if (out_dev) {
    int n = 0;
    auto add = [&](float* dst, const float* src, int64_t cnt) {
      if (dst) { ctx->fin_oc.dst[n] = dst; ctx->fin_oc.src[n] = src; ctx->fin_oc.n[n] = cnt; ++n; } };
    add(out_dev->pred, p.hf(p.hdr.pred), d.B);
    add(out_dev->logit, p.hf(p.hdr.logit), d.B);
    add(out_dev->w_hop0, p.hf(p.hdr.w_hop0), (int64_t)d.B * d.L);
    add(out_dev->memory, p.hf(p.hdr.memory), (int64_t)d.B * d.L * d.H);
  }
  ctx->fin_scalars = scalars; ctx->fin_mreg = hy.memory_reg; ctx->fin_copies = out_dev != nullptr;
  const bool fin_on_side = with_backward && G == 1 && ctx->overlap && !ctx->profile && hy.l2_reg == 0.f;
  ctx->fin_early = fin_on_side;
  if (G == 1) {
    fwd_rows(ctx, p, s, hy, ids, labels, params, table, scalars, side_pack, st);
    if (with_backward) bwd_rows(ctx, p, s, hy, ids, labels, params, grads, dtable, scalars, true, st);
  } else {
    CK(cudaEventRecord(ctx->ev_gstart, st));
    const int base = d.B / G, rem = d.B % G;
    int row0 = 0;
    for (int g = 0; g < G; ++g) {
      const int rows = base + (g < rem ? 1 : 0);
      const Plan gp = group_plan(p, s, g, G, row0, rows);
      cudaStream_t gs = ctx->gstream[g];
      CK(cudaStreamWaitEvent(gs, ctx->ev_gstart, 0));
      fwd_rows(ctx, gp, s, hy, ids, labels, params, table, scalars, false, gs);
      if (with_backward) bwd_rows(ctx, gp, s, hy, ids, labels, params, grads, dtable, scalars, false, gs);
      CK(cudaEventRecord(ctx->ev_gdone[g], gs));
      CK(cudaStreamWaitEvent(st, ctx->ev_gdone[g], 0));
      row0 += rows;
    }
  }
  { Bracket b(ctx, st, HPMN_K_MISC);
    if (!fin_on_side) launch_finish_scalars(L, scalars, hy.memory_reg, out_dev ? &ctx->fin_oc : nullptr, st);
    if (hy.l2_reg != 0.f) {   // l2_reg * tf.nn.l2_loss(v) over every trainable, code/hpmn.py:204-205
      // Under data parallelism every rank holds the same parameters and the gradients / scalars are SUMMED over ranks:
      // each rank contributes its share B / loss_batch of the (batch-independent) l2 term, so the sum is exactly one l2 term.
      const float l2 = hy.l2_reg * (float)d.B / (float)hy.loss_batch;
      sumsq_kernel<<<ctx->sms * 4, 256, 0, st>>>(params, p.pl.total, 0.5f * l2, scalars + HPMN_S_LOSS);
      sumsq_kernel<<<ctx->sms * 4, 256, 0, st>>>(table, s->V * (int64_t)d.E, 0.5f * l2, scalars + HPMN_S_LOSS);
      ctx->launches += 2;
      if (with_backward) {
        launch_axpy(L, grads, params, l2, p.pl.total, st);
        launch_axpy(L, dtable, table, l2, s->V * (int64_t)d.E, st);
      }
    } }
  if (with_backward && ctx->comm && ctx->dtable_late) {   // every writer of dtable is joined into `st` by now
    CK(cudaEventRecord(ctx->ev_dtable, st));
    CK(cudaStreamWaitEvent(ctx->comm, ctx->ev_dtable, 0));
  }
  return HPMN_OK;
}

// copy the per-row results from the header staging to wherever the caller wants them
static int copy_outputs(hpmn_ctx* ctx, const Plan& p, const hpmn_outputs* out, cudaMemcpyKind kind, cudaStream_t st) {
  const Dims& d = p.d;
  if (out->pred) CK(cudaMemcpyAsync(out->pred, p.hf(p.hdr.pred), (size_t)d.B * sizeof(float), kind, st));
  if (out->logit) CK(cudaMemcpyAsync(out->logit, p.hf(p.hdr.logit), (size_t)d.B * sizeof(float), kind, st));
  if (out->w_hop0) CK(cudaMemcpyAsync(out->w_hop0, p.hf(p.hdr.w_hop0), (size_t)d.B * d.L * sizeof(float), kind, st));
  if (out->memory) CK(cudaMemcpyAsync(out->memory, p.hf(p.hdr.memory), (size_t)d.B * d.L * d.H * sizeof(float), kind, st));
  return HPMN_OK;
}

int hpmn_forward(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const int32_t* ids, const int32_t* labels,
                 const float* params, const float* table, const hpmn_outputs* out, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!ids || !labels || !params || !table || !out || !out->scalars) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  hpmn_hyper h = hy ? *hy : default_hyper();
  cudaStream_t st = (cudaStream_t)stream;
  rc = run_step(ctx, p, s, h, ids, labels, params, table, nullptr, nullptr, 0, false, out->scalars, st, out);
  if (rc) return rc;
  return check_launch(ctx, "hpmn_forward");
}

int hpmn_forward_backward(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const int32_t* ids,
                          const int32_t* labels, const float* params, const float* table, float* grads, float* dtable,
                          int zero_dtable, const hpmn_outputs* out, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!ids || !labels || !params || !table || !grads || !dtable || !out || !out->scalars)
    return fail(ctx, HPMN_EINVAL, "NULL buffer");
  hpmn_hyper h = hy ? *hy : default_hyper();
  cudaStream_t st = (cudaStream_t)stream;
  rc = run_step(ctx, p, s, h, ids, labels, params, table, grads, dtable, zero_dtable, true, out->scalars, st, out);
  if (rc) return rc;
  return check_launch(ctx, "hpmn_forward_backward");
}

int hpmn_step_host_begin(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const int32_t* ids_host,
                   const int32_t* labels_host, const float* params, const float* table, float* grads, float* dtable,
                   int zero_dtable, int with_backward, const hpmn_outputs* oh, void* workspace, void* stream) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!ids_host || !labels_host || !params || !table || !oh || !oh->scalars) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  if (with_backward && (!grads || !dtable)) return fail(ctx, HPMN_EINVAL, "NULL gradient buffer");
  hpmn_hyper h = hy ? *hy : default_hyper();
  cudaStream_t st = (cudaStream_t)stream;
  const Dims& d = p.d;
  // two feed slots: a batch staged by hpmn_prefetch_host is consumed in place, otherwise it is copied on `st` into the
  // slot no prefetch is pending for
  int slot;
  const bool staged = ctx->staged_ids == ids_host && ctx->staged_labels == labels_host && ctx->staged_B == d.B;
  if (staged) {
    slot = ctx->staged_slot;
    CK(cudaStreamWaitEvent(st, ctx->ev_copy, 0));
    ctx->staged_ids = nullptr; ctx->staged_labels = nullptr;
  } else {
    slot = ctx->staged_ids != nullptr ? 1 - ctx->staged_slot : 0;
  }
  int32_t* ids = reinterpret_cast<int32_t*>(p.base + (slot ? p.hdr.ids2 : p.hdr.ids));
  int32_t* labels = reinterpret_cast<int32_t*>(p.base + (slot ? p.hdr.labels2 : p.hdr.labels));
  float* scalars = p.hf(p.hdr.scalars);
  if (!staged) {
    CK(cudaMemcpyAsync(ids, ids_host, (size_t)d.B * d.T * d.F * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(labels, labels_host, (size_t)d.B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  ctx->cur_slot = slot;
  const char* hb = reinterpret_cast<const char*>(oh->scalars);
  const bool mirrored = oh->pred && oh->logit && oh->w_hop0 && !oh->memory &&
                        reinterpret_cast<const char*>(oh->pred) - hb == (ptrdiff_t)(p.hdr.pred - p.hdr.scalars) &&
                        reinterpret_cast<const char*>(oh->logit) - hb == (ptrdiff_t)(p.hdr.logit - p.hdr.scalars) &&
                        reinterpret_cast<const char*>(oh->w_hop0) - hb == (ptrdiff_t)(p.hdr.w_hop0 - p.hdr.scalars);
  const int k = ctx->out_next; ctx->out_next ^= 1;     // hpmn_step_host_end(oh) waits for ev_out[k]
  // a host caller whose result buffers mirror the staging block gets ONE D2H copy per step; in the training step run_step
  // issues it on the side stream as soon as the block is final (behind the attention / head section)
  ctx->fin_d2h_dst = mirrored ? oh->scalars : nullptr;
  ctx->fin_d2h_bytes = p.hdr.out_end - p.hdr.scalars;
  ctx->fin_d2h_ev = ctx->ev_out[k];
  ctx->fin_d2h_done = false;
  rc = run_step(ctx, p, s, h, ids, labels, params, table, grads, dtable, zero_dtable, with_backward != 0, scalars, st);
  ctx->fin_d2h_dst = nullptr;
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev_consumed[slot], st));       // the slot may be refilled once everything above has read it
  ctx->consumed_valid[slot] = true;
  if (!ctx->fin_d2h_done) {
    if (mirrored) {
      CK(cudaMemcpyAsync(oh->scalars, scalars, p.hdr.out_end - p.hdr.scalars, cudaMemcpyDeviceToHost, st));
    } else {
      CK(cudaMemcpyAsync(oh->scalars, scalars, 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
      rc = copy_outputs(ctx, p, oh, cudaMemcpyDeviceToHost, st);
      if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev_out[k], st));
  }
  ctx->out_key[k] = oh->scalars;
  return check_launch(ctx, "hpmn_step_host");
}

int hpmn_step_host_end(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_outputs* oh, void* stream) {
  if (!ctx || !s || !oh || !oh->scalars) return HPMN_EINVAL;
  int k = -1;
  for (int i = 0; i < 2; ++i) if (ctx->out_key[i] == oh->scalars) k = i;
  if (k >= 0) { CK(cudaEventSynchronize(ctx->ev_out[k])); ctx->out_key[k] = nullptr; }
  else CK(cudaStreamSynchronize((cudaStream_t)stream));
  if (oh->scalars[HPMN_S_IDERR] != 0.f)   // TF's GatherV2 raises InvalidArgumentError on CPU
    return fail(ctx, HPMN_EINVAL, "an id is outside [0, feature_size=%lld)", (long long)s->V);
  return HPMN_OK;
}

int hpmn_step_host(hpmn_ctx* ctx, const hpmn_shape* s, const hpmn_hyper* hy, const int32_t* ids_host,
                   const int32_t* labels_host, const float* params, const float* table, float* grads, float* dtable,
                   int zero_dtable, int with_backward, const hpmn_outputs* oh, void* workspace, void* stream) {
  int rc = hpmn_step_host_begin(ctx, s, hy, ids_host, labels_host, params, table, grads, dtable, zero_dtable, with_backward, oh,
                                workspace, stream);
  if (rc) return rc;
  return hpmn_step_host_end(ctx, s, oh, stream);
}

int hpmn_prefetch_host(hpmn_ctx* ctx, const hpmn_shape* s, const int32_t* ids_host, const int32_t* labels_host, void* workspace) {
  Plan p; int rc = make_plan(ctx, s, workspace, p);
  if (rc) return rc;
  if (!ids_host || !labels_host) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  const Dims& d = p.d;
  const int slot = 1 - ctx->cur_slot;                   // the slot the most recent step is NOT reading
  if (ctx->consumed_valid[slot]) CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_consumed[slot], 0));
  int32_t* ids = reinterpret_cast<int32_t*>(p.base + (slot ? p.hdr.ids2 : p.hdr.ids));
  int32_t* labels = reinterpret_cast<int32_t*>(p.base + (slot ? p.hdr.labels2 : p.hdr.labels));
  CK(cudaMemcpyAsync(ids, ids_host, (size_t)d.B * d.T * d.F * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->copy));
  CK(cudaMemcpyAsync(labels, labels_host, (size_t)d.B * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->copy));
  CK(cudaEventRecord(ctx->ev_copy, ctx->copy));
  ctx->staged_ids = ids_host; ctx->staged_labels = labels_host; ctx->staged_slot = slot; ctx->staged_B = d.B;
  return HPMN_OK;
}

int hpmn_debug_wgrad(hpmn_ctx* ctx, const hpmn_shape* s, int k, const float* xin, int64_t ldx, const float* st, const float* da,
                     float* grads, int use_tc, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  Dims d = make_dims(s);
  if (!d.ok || k < 0 || k >= d.L || !xin || !st || !da || !grads) return fail(ctx, HPMN_EINVAL, "bad argument");
  cudaSetDevice(ctx->device);
  ParamLayout pl = make_param_layout(d);
  Launch L{&ctx->launches, ctx->sms};
  float *dWg = grads + pl.Wg[k], *dbg = grads + pl.bg[k], *dWc = grads + pl.Wc[k], *dbc = grads + pl.bc[k];
  if (use_tc) {
    if (!launch_tc_wgrad(L, d, k, xin, ldx, st, da, dWg, dbg, dWc, dbc, (cudaStream_t)stream))
      return fail(ctx, HPMN_EINVAL, "no tcgen05 wgrad instantiation for this shape");
  } else {
    launch_gru_wgrad(L, d, k, xin, ldx, st, da, dWg, dbg, dWc, dbc, (cudaStream_t)stream);
  }
  return check_launch(ctx, "hpmn_debug_wgrad");
}

// ---- head over an arbitrary input width (user + item sides concatenated, code/hpmn.py:452-465) --------------------------
static ParamLayout head_only_layout(int R) {
  ParamLayout p; memset(&p, 0, sizeof(p));
  int64_t off = 0;
  auto take = [&](int64_t sz) { int64_t o = off; off = align4(off + sz); return o; };
  p.gamma = take(R); p.beta = take(R);
  p.F1 = take((int64_t)R * FC1); p.f1 = take(FC1);
  p.F2 = take((int64_t)FC1 * FC2); p.f2 = take(FC2);
  p.F3 = take(FC2); p.f3 = take(1);
  p.total = off; p.ntensors = 8;
  return p;
}
struct HeadWideWs { HeadWs w; float* pred; size_t total; };
static HeadWideWs head_wide_ws(char* base, int B, int R) {
  HeadWideWs h; size_t off = 0;
  auto take = [&](size_t n) { float* p = reinterpret_cast<float*>(base + off); off = (off + n * sizeof(float) + 255) & ~(size_t)255; return p; };
  h.w.bn = take((size_t)B * R); h.w.dbn = take((size_t)B * R); h.w.dgt = take((size_t)B * R);
  h.w.a1 = take((size_t)B * FC1); h.w.act1 = take((size_t)B * FC1); h.w.dl1 = take((size_t)B * FC1);
  h.w.a2 = take((size_t)B * FC2); h.w.act2 = take((size_t)B * FC2); h.w.dl2 = take((size_t)B * FC2);
  h.w.dlogit = take((size_t)B); h.pred = take((size_t)B);
  h.total = off;
  return h;
}
int64_t hpmn_head_wide_param_count(int R) { return R > 0 ? head_only_layout(R).total : HPMN_EINVAL; }
int hpmn_head_wide_param_offsets(int R, int64_t* offsets, int64_t* sizes) {
  if (R <= 0 || !offsets || !sizes) return HPMN_EINVAL;
  const ParamLayout p = head_only_layout(R);
  const int64_t o[8] = {p.gamma, p.beta, p.F1, p.f1, p.F2, p.f2, p.F3, p.f3};
  const int64_t z[8] = {R, R, (int64_t)R * FC1, FC1, (int64_t)FC1 * FC2, FC2, FC2, 1};
  for (int i = 0; i < 8; ++i) { offsets[i] = o[i]; sizes[i] = z[i]; }
  return 8;
}
size_t hpmn_head_wide_workspace_bytes(int B, int R) { return (B > 0 && R > 0) ? head_wide_ws(nullptr, B, R).total : 0; }

static int head_wide_dims(hpmn_ctx* ctx, int B, int R, Dims& d) {
  if (!ctx) return HPMN_EINVAL;
  if (B <= 0 || R <= 0 || R > 2 * (HP + 64)) return fail(ctx, HPMN_EINVAL, "head input width %d outside (0, %d]", R, 2 * (HP + 64));
  memset(&d, 0, sizeof(d));
  d.B = B; d.R = R; d.ok = true;
  cudaSetDevice(ctx->device);
  return HPMN_OK;
}

int hpmn_head_wide_fwd(hpmn_ctx* ctx, int B, int R, const hpmn_hyper* hy, const float* repre, const int32_t* labels, const float* hparams,
                       float* pred, float* logit, float* scalars, void* workspace, void* stream) {
  Dims d; int rc = head_wide_dims(ctx, B, R, d);
  if (rc) return rc;
  if (!hy || !repre || !labels || !hparams || !pred || !logit || !scalars || !workspace) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  Launch L{&ctx->launches, ctx->sms};
  HeadWideWs w = head_wide_ws(static_cast<char*>(workspace), B, R);
  launch_head_fwd(L, d, head_only_layout(R), *hy, 0, repre, labels, hparams, w.pred, logit, scalars, w.w, (cudaStream_t)stream);
  CK(cudaMemcpyAsync(pred, w.pred, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return check_launch(ctx, "hpmn_head_wide_fwd");
}

int hpmn_head_wide_bwd(hpmn_ctx* ctx, int B, int R, const hpmn_hyper* hy, const float* repre, const int32_t* labels, const float* hparams,
                       float* drepre, float* hgrads, void* workspace, void* stream) {
  Dims d; int rc = head_wide_dims(ctx, B, R, d);
  if (rc) return rc;
  if (!hy || !repre || !labels || !hparams || !drepre || !hgrads || !workspace) return fail(ctx, HPMN_EINVAL, "NULL buffer");
  Launch L{&ctx->launches, ctx->sms};
  HeadWideWs w = head_wide_ws(static_cast<char*>(workspace), B, R);
  AtbBatch batch; batch.n = 0; batch.blocks = 0;
  launch_head_bwd(L, d, head_only_layout(R), *hy, 0, repre, labels, hparams, w.pred, drepre, hgrads, w.w, batch, (cudaStream_t)stream);
  launch_atb_batch(L, batch, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_head_wide_bwd");
}

int hpmn_output_block(const hpmn_shape* s, size_t* offsets, size_t* total) {
  Dims d = make_dims(s);
  if (!d.ok || !offsets || !total) return HPMN_EINVAL;
  const Hdr h = make_hdr(d);
  offsets[0] = 0; offsets[1] = h.pred - h.scalars; offsets[2] = h.logit - h.scalars; offsets[3] = h.w_hop0 - h.scalars;
  *total = h.out_end - h.scalars;
  return HPMN_OK;
}

int hpmn_nvls_allreduce(hpmn_ctx* ctx, float* multicast_ptr, int64_t n_floats, int rank, int world, int ctas, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  if (!multicast_ptr || n_floats < 0 || (n_floats & 3) || world < 1 || rank < 0 || rank >= world)
    return fail(ctx, HPMN_EINVAL, "bad argument (the buffer must be a multiple of 4 floats)");
  if (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) return fail(ctx, HPMN_EINVAL, "multicast pointer must be 16-byte aligned");
  cudaSetDevice(ctx->device);
  Launch L{&ctx->launches, ctx->sms};
  launch_nvls_allreduce(L, multicast_ptr, n_floats, rank, world, ctas, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_nvls_allreduce");
}

int hpmn_table_grad_sources(hpmn_ctx* ctx, const hpmn_shape* s, size_t* ids_off, size_t* dx_off, size_t* dlast_off) {
  if (!ctx || !s || !ids_off || !dx_off || !dlast_off) return HPMN_EINVAL;
  Dims d = make_dims(s);
  if (!d.ok) return fail(ctx, HPMN_EINVAL, "invalid hpmn_shape");
  int G = ctx->groups;
  while (G > 1 && d.B / G < ctx->group_min_rows) --G;
  if (G != 1) return fail(ctx, HPMN_EINVAL, "the step runs as %d row groups: dX is not one contiguous block", G);
  const Hdr h = make_hdr(d);
  const WsLayout w = make_ws_layout(d);
  *ids_off = ctx->cur_slot ? h.ids2 : h.ids;
  *dx_off = h.total + (want_tcrec(ctx, d) ? w.tcr + make_tcr_layout(d).dx[0] : w.dxk[0]);
  *dlast_off = h.total + w.dlast;
  return HPMN_OK;
}

int hpmn_debug_tcr_stamps(long long* out_host, int n) {
  long long* buf = tcr_debug_buffer();
  if (!buf || !out_host || n <= 0) return HPMN_EINVAL;
  if (n > 2048 * 16) n = 2048 * 16;
  return cudaMemcpy(out_host, buf, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? HPMN_OK : HPMN_ECUDA;
}

int hpmn_clip_adam(hpmn_ctx* ctx, float* var, const float* grad, float* m, float* v, int64_t n, int64_t t, float lr,
                   float beta1, float beta2, float eps, float clip, void* stream) {
  if (!ctx) return HPMN_EINVAL;
  if (!var || !grad || !m || !v || n < 0 || t < 1) return fail(ctx, HPMN_EINVAL, "bad argument");
  cudaSetDevice(ctx->device);
  Launch L{&ctx->launches, ctx->sms};
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  launch_clip_adam(L, var, grad, m, v, n, (float)lr_t, beta1, beta2, eps, clip, (cudaStream_t)stream);
  return check_launch(ctx, "hpmn_clip_adam");
}

}  // extern "C"
