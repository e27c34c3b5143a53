/* hpmn_b200.h -- C ABI of libhpmn_b200.so: the HPMN forward/backward hot path on B200 (sm_100a).
 *
 * The reference (alimamarankgroup/HPMN) has no plugin / FFI interface: its hot path is the TF1.4
 * graph built by code/hpmn.py and executed by `sess.run` (code/hpmn.py:336,365,482,511).  This
 * header is the boundary a maintainer would bind instead of that `sess.run`: each entry point cites
 * the reference graph section it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns int (HPMN_OK or a negative HPMN_E*),
 *     never throws; the message for the last failure is hpmn_last_error(ctx).
 *   - the CALLER owns every buffer (device memory unless a parameter is called *_host); the library
 *     borrows it for the call and never allocates device memory after hpmn_create().
 *   - all work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed as void*;
 *     NULL = legacy default stream); no hidden synchronisation except in the *_host entry points,
 *     which synchronise the stream before returning (they hand results back in host memory).
 *   - fp32, row-major, 16-byte aligned; weights keep the TF kernel layout ([D_in+H, 2H] with gate
 *     columns ordered r then u, code/util.py:88-96) so TF checkpoints map 1:1.
 *   - a ctx is not thread-safe; distinct ctxs (one per rank / GPU) are independent.
 *   - there is NO CPU fallback: on a device that is not sm_100 hpmn_create() fails with HPMN_EARCH.
 */
#ifndef HPMN_B200_H
#define HPMN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPMN_ABI_VERSION 1
#define HPMN_MAX_LAYERS 16
#define HPMN_MAX_HOPS 8

enum {
  HPMN_OK = 0,
  HPMN_EINVAL = -1, /* bad shape / NULL pointer / id out of range */
  HPMN_EARCH = -2,  /* device is not sm_100 */
  HPMN_ECUDA = -3,  /* a CUDA call failed; see hpmn_last_error */
  HPMN_ENOMEM = -4
};

typedef struct hpmn_ctx hpmn_ctx;

/* Static configuration of one memory side (the `User` scope, code/hpmn.py:436-442 / 287-295). */
typedef struct hpmn_shape {
  int32_t B;           /* batch rows held by this rank */
  int32_t T;           /* id steps per sample as fed (user_maxlen, code/hpmn.py:249) */
  int32_t F;           /* id features per step (user_dim) */
  int32_t E;           /* embedding_size (multiple of 4) */
  int32_t H;           /* hidden_size: <= 32, or 64 (tensor-core recurrence; attention kernels compiled for 64 lanes) */
  int32_t L;           /* number of GRU layers (user_num_layers) */
  int32_t hops;        /* self.hop */
  int32_t front_pad;   /* zero steps prepended: Hpmn_Industry 23 (code/hpmn.py:288-289), Hpmn 0 */
  int32_t mask_id0;    /* 1: id 0 embeds to zeros (code/hpmn.py:417-423); 0: Hpmn_Industry */
  int32_t last_offset; /* target step from the end: Hpmn 1 (code/hpmn.py:439), Industry 2 (:292) */
  int32_t periods[HPMN_MAX_LAYERS]; /* li_layer[k], k < L-1 (code/hpmn.py:122-128) */
  int64_t V;           /* feature_size: rows of the embedding table */
} hpmn_shape;

/* Scalars of one step, written to `scalars[4]`: */
enum { HPMN_S_LOGLOSS = 0, HPMN_S_COVREG = 1, HPMN_S_LOSS = 2, HPMN_S_IDERR = 3 };

/* Loss / regularisation knobs (code/hpmn.py:202-207, 480, 509). */
typedef struct hpmn_hyper {
  float memory_reg;     /* weight of the covariance regulariser (memory_loss) */
  float l2_reg;         /* l2_reg * sum l2_loss(v) over all trainables (0 in every reference config) */
  float keep_prob;      /* dropout keep probability; 1 = eval */
  uint64_t dropout_seed;/* counter-based mask seed (TF's RNG stream cannot be matched) */
  int32_t loss_batch;   /* batch the log-loss mean divides by; 0 = B. Set to the GLOBAL batch when
                           B is one rank's shard so a SUM all-reduce gives the single-GPU gradient */
} hpmn_hyper;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int hpmn_abi_version(void);
int hpmn_create(hpmn_ctx** out, int device);
void hpmn_destroy(hpmn_ctx* ctx);
const char* hpmn_last_error(hpmn_ctx* ctx); /* ctx may be NULL: last error of a failed hpmn_create */
/* number of kernels this ctx has launched since create (bench.py's gpu_launches) */
int64_t hpmn_launch_count(hpmn_ctx* ctx);
/* Multi-GPU overlap hook (the reference has no distributed code; replaces nothing).  `stream` (a cudaStream_t, or NULL to
 * clear) is made to wait, inside every hpmn_forward_backward / hpmn_step_host call, for the point at which the embedding-table
 * gradient `dtable` is final -- the scatter is queued in front of the GRU weight-gradient reduction -- so the caller's
 * all-reduce of `dtable` on that stream runs beside the rest of the backward pass.  The dense gradients are final when the
 * call's own stream reaches the end of the call, as before. */
int hpmn_set_comm_stream(hpmn_ctx* ctx, void* stream);

/* ---- layouts (pure host functions, usable without a GPU) ---------------------------------- */
/* Dense parameters live in ONE flat fp32 buffer (gradients in a twin buffer of the same layout):
 * for k<L: gates/kernel [Din_k+H,2H], gates/bias [2H], candidate/kernel [Din_k+H,H], candidate/bias [H]
 * (Din_0 = F*E, Din_k = H); dense/kernel [D,H], dense/bias [H]; map [H,H]; per hop
 * dense_{3h+1..3h+3}: [4H,80],[80],[80,40],[40],[40,1],[1]; bn1 gamma,beta [H+D];
 * fc1 [H+D,200],[200]; fc2 [200,80],[80]; fc3 [80,1],[1].  Every tensor starts on a 4-float boundary. */
int hpmn_param_tensors(const hpmn_shape* s);                          /* number of tensors, <0 on bad shape */
int64_t hpmn_param_count(const hpmn_shape* s);                        /* floats in the flat buffer */
int hpmn_param_offsets(const hpmn_shape* s, int64_t* offsets, int64_t* sizes, int n);
size_t hpmn_workspace_bytes(const hpmn_shape* s, int for_bwd);        /* scratch + saved activations */

/* ---- K1 / K5: embedding gather and its adjoint (code/hpmn.py:414-430, 266-282, 288-289) ---- */
/* ids [B,T,F] int32, table [V,E] -> x [B,T+front_pad,F*E] (front pad rows and masked ids are zeros) */
int hpmn_gather_fwd(hpmn_ctx*, const hpmn_shape*, const int32_t* ids, const float* table, float* x,
                    void* stream);
/* dtable[ids] += dx rows (+ dlast [B,F*E] on the target step when dlast != NULL); dtable is accumulated */
int hpmn_gather_bwd(hpmn_ctx*, const hpmn_shape*, const int32_t* ids, const float* dx,
                    const float* dlast, float* dtable, void* stream);

/* the same adjoint over nsrc (<= 64) sources in one launch: ids[i], dx[i], dlast[i] (dlast may be NULL) are device pointers --
 * this rank's own buffers or the symmetric-memory mappings of peer ranks' buffers (peer-row gradient exchange, see below) */
int hpmn_gather_bwd_multi(hpmn_ctx*, const hpmn_shape*, int nsrc, const int32_t* const* ids, const float* const* dx,
                          const float* const* dlast, float* dtable, void* stream);

/* ---- K2 / K4: hierarchical periodic memory (code/hpmn.py:113-131; cell code/util.py:81-110) -- */
/* x [B,Tpad,D], params flat -> memory [B,L,H]; activations for the adjoint are kept in workspace */
int hpmn_memory_fwd(hpmn_ctx*, const hpmn_shape*, const float* x, const float* params, float* memory,
                    void* workspace, void* stream);
/* dmemory [B,L,H] -> dx [B,Tpad,D] (overwritten), GRU entries of grads (accumulated) */
int hpmn_memory_bwd(hpmn_ctx*, const hpmn_shape*, const float* x, const float* params,
                    const float* dmemory, float* dx, float* grads, void* workspace, void* stream);

/* ---- K3: covariance regulariser + multi-hop memory attention (code/hpmn.py:161-182, 133-146) - */
/* memory [B,L,H], x (for last = x[:, -last_offset]) -> repre [B,H+D] = [q | last], w_hop0 [B,L],
 * scalars[HPMN_S_COVREG] += sum_b ||offdiag cov||_F  (scalars must be zeroed by the caller) */
int hpmn_attn_fwd(hpmn_ctx*, const hpmn_shape*, const float* memory, const float* x, const float* params,
                  float* repre, float* w_hop0, float* scalars, void* workspace, void* stream);
/* drepre [B,H+D] -> dmemory [B,L,H] (overwritten, includes memory_reg * d covreg), dlast [B,D]
 * (overwritten), attention entries of grads (accumulated) */
int hpmn_attn_bwd(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const float* memory, const float* x,
                  const float* params, const float* drepre, float* dmemory, float* dlast, float* grads,
                  void* workspace, void* stream);

/* ---- head: BN(inference) -> 200 ELU -> 80 ELU -> 1 sigmoid, log-loss (code/hpmn.py:190-202) -- */
int hpmn_head_fwd(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const float* repre,
                  const int32_t* labels, const float* params, float* pred, float* logit, float* scalars,
                  void* workspace, void* stream);
int hpmn_head_bwd(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const float* repre,
                  const int32_t* labels, const float* params, float* drepre, float* grads,
                  void* workspace, void* stream);

/* The same head over an arbitrary input width R <= 192: `repre = concat([user_repre, item_repre])` when both memory sides are on
 * (item=True, code/hpmn.py:444-465).  hparams / hgrads: flat [gamma R | beta R | fc1 R x 200, 200 | fc2 200 x 80, 80 | fc3 80, 1]
 * (every tensor on a 4-float boundary: hpmn_head_wide_param_offsets); hgrads is accumulated.  _bwd uses the activations _fwd left
 * in `workspace` (hpmn_head_wide_workspace_bytes).  hpmn_b200/dual.py composes the two sides from the K1-K5 entry points. */
int64_t hpmn_head_wide_param_count(int R);
int hpmn_head_wide_param_offsets(int R, int64_t* offsets, int64_t* sizes); /* 8 tensors */
size_t hpmn_head_wide_workspace_bytes(int B, int R);
int hpmn_head_wide_fwd(hpmn_ctx*, int B, int R, const hpmn_hyper*, const float* repre, const int32_t* labels, const float* hparams,
                       float* pred, float* logit, float* scalars, void* workspace, void* stream);
int hpmn_head_wide_bwd(hpmn_ctx*, int B, int R, const hpmn_hyper*, const float* repre, const int32_t* labels, const float* hparams,
                       float* drepre, float* hgrads, void* workspace, void* stream);

/* ---- whole path: what `sess.run` does (code/hpmn.py:365-367 eval, :336/:482 train, minus Adam) */
typedef struct hpmn_outputs {  /* each pointer may be NULL except scalars */
  float* scalars;  /* [4]  HPMN_S_* */
  float* pred;     /* [B]  sigmoid probability = self.prediction (code/hpmn.py:197) */
  float* logit;    /* [B]  pre-sigmoid fc3 */
  float* w_hop0;   /* [B,L] self.user_weights (code/hpmn.py:182,293) */
  float* memory;   /* [B,L,H] */
} hpmn_outputs;

/* device buffers in, device buffers out, asynchronous on `stream` */
int hpmn_forward(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const int32_t* ids, const int32_t* labels,
                 const float* params, const float* table, const hpmn_outputs* out, void* workspace,
                 void* stream);
/* + gradients of loss = logloss + memory_reg*covreg (+ l2): grads (flat, OVERWRITTEN), dtable [V,E]
 * (OVERWRITTEN when zero_dtable != 0, else accumulated) -- tf.gradients at code/hpmn.py:211 before the clip */
int hpmn_forward_backward(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const int32_t* ids,
                          const int32_t* labels, const float* params, const float* table, float* grads,
                          float* dtable, int zero_dtable, const hpmn_outputs* out, void* workspace,
                          void* stream);
/* same, but ids / labels / outputs are HOST buffers (pinned for truly async copies): H2D of the feed,
 * compute, D2H of the fetched results, then a stream synchronise -- the feed_dict / fetches round trip
 * of code/hpmn.py:474-482.  with_backward = 0 is the eval fetch of code/hpmn.py:511-513. */
int hpmn_step_host(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const int32_t* ids_host,
                   const int32_t* labels_host, const float* params, const float* table, float* grads,
                   float* dtable, int zero_dtable, int with_backward, const hpmn_outputs* out_host,
                   void* workspace, void* stream);

/* The same call split in two so that other work can be queued between enqueue and wait: _begin enqueues H2D + compute +
 * D2H and returns immediately, _end waits until the results of THAT step (identified by out_host->scalars) have landed in host
 * memory and reports an out-of-range id.  Up to two steps may be in flight when they use distinct out_host buffers: _begin of
 * step i+1 may be called before _end of step i, so the host's enqueue time of step i+1 hides behind the device time of step i.
 * _end returns when the HOST results are there -- in the training step (with_backward, mirrored result block, l2_reg = 0) that is
 * as soon as the attention / head section has run, i.e. before the backward recurrence has finished.  The gradient buffers
 * (grads, dtable) are device memory and complete in stream order on `stream`, like after hpmn_forward_backward: anything queued
 * on `stream` afterwards sees them; a host read needs a stream synchronisation. */
int hpmn_step_host_begin(hpmn_ctx*, const hpmn_shape*, const hpmn_hyper*, const int32_t* ids_host,
                         const int32_t* labels_host, const float* params, const float* table, float* grads,
                         float* dtable, int zero_dtable, int with_backward, const hpmn_outputs* out_host,
                         void* workspace, void* stream);
int hpmn_step_host_end(hpmn_ctx*, const hpmn_shape*, const hpmn_outputs* out_host, void* stream);

/* Optional single result copy: byte offsets (scalars, pred, logit, w_hop0) and total size of the block in which the library stages
 * the fetched results on the device.  When the out_host pointers of a *_host call are laid out the same way inside one pinned
 * block (and out_host->memory is NULL) the results come back in ONE D2H copy instead of four. */
int hpmn_output_block(const hpmn_shape*, size_t* offsets /* [4] */, size_t* total);

/* Optional double buffering of the feed: start the H2D copy of the NEXT batch (pinned host memory) on the library's copy
 * stream while the current step computes.  A following hpmn_step_host() called with the same ids_host / labels_host pointers
 * and shape uses the staged copy instead of copying again.  Call it between hpmn_step_host_begin and hpmn_step_host_end of the
 * current step.  The host buffers must stay untouched until the consuming call returns. */
int hpmn_prefetch_host(hpmn_ctx*, const hpmn_shape*, const int32_t* ids_host, const int32_t* labels_host, void* workspace);

/* ---- multi-GPU gradient exchange over NVLink / NVSwitch (the reference is single-process) -------- */
/* In-switch SUM all-reduce, in place, of a buffer that lives in symmetric memory: `multicast_ptr` is the multicast address of
 * the buffer (every rank's copy mapped behind one address), n_floats a multiple of 4.  Rank r reduces slice r with
 * multimem.ld_reduce and broadcasts it with multimem.st.  The caller makes every rank's copy final before the launch and
 * lets no rank read the result before every rank's launch has finished (symmetric-memory barriers on `stream`). */
int hpmn_nvls_allreduce(hpmn_ctx*, float* multicast_ptr, int64_t n_floats, int rank, int world, int ctas, void* stream);
/* Where, inside `workspace`, the embedding scatter of the LAST hpmn_forward_backward / hpmn_step_host call took its inputs:
 * byte offsets of the host entry point's id slot [B,T,F] int32, of dX of layer 0 [B,Tpad,D] and of dlast [B,D].  A peer rank that
 * maps this workspace feeds them to its own hpmn_gather_bwd (peer-row exchange: 36 MB per peer instead of the dense table
 * gradient).  HPMN_EINVAL when the step ran as several row groups. */
int hpmn_table_grad_sources(hpmn_ctx*, const hpmn_shape*, size_t* ids_off, size_t* dx_off, size_t* dlast_off);

/* ---- update step: clip_by_value(g,-1,1) + dense Adam (code/hpmn.py:209-214) ---------------- */
/* var, m, v updated in place over n floats; t = 1-based step; TF1.4 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t) */
int hpmn_clip_adam(hpmn_ctx*, float* var, const float* grad, float* m, float* v, int64_t n, int64_t t,
                   float lr, float beta1, float beta2, float eps, float clip, void* stream);

/* ---- measurement support ------------------------------------------------------------------ */
/* When enabled, hpmn_forward_backward brackets each kernel family with CUDA events on `stream` and
 * accumulates their durations (read back with hpmn_profile_read, which synchronises the events). */
enum { HPMN_K_GATHER = 0, HPMN_K_INPROJ, HPMN_K_REC_FWD, HPMN_K_ATTN_FWD, HPMN_K_HEAD_FWD, HPMN_K_HEAD_BWD,
       HPMN_K_ATTN_BWD, HPMN_K_REC_BWD, HPMN_K_DX, HPMN_K_WGRAD, HPMN_K_SCATTER, HPMN_K_MISC, HPMN_K_COUNT };
int hpmn_profile_enable(hpmn_ctx*, int on);
/* ms[HPMN_K_COUNT] = accumulated milliseconds, calls[HPMN_K_COUNT] = brackets accumulated; resets */
int hpmn_profile_read(hpmn_ctx*, float* ms, int64_t* calls);
const char* hpmn_kernel_family_name(int family);

/* ---- kernel-level test hook ------------------------------------------------------------------ */
/* GRU weight gradients of layer k from caller buffers: xin rows [B*S_k] (row stride ldx floats), st [B*S_k,128]
 * (h|r|u|c), da [B*S_k,96]; accumulates into grads (flat layout).  use_tc: 1 = tcgen05 kernel, 0 = fp32 FFMA kernel.
 * Returns HPMN_EINVAL if the requested implementation does not cover the shape. */
int hpmn_debug_wgrad(hpmn_ctx*, const hpmn_shape*, int k, const float* xin, int64_t ldx, const float* st, const float* da,
                     float* grads, int use_tc, void* stream);

/* HPMN_TCR_DEBUG=1 in the environment: the tensor-core recurrence (layer 0, CTA 0) stamps clock64 at the hand-offs of
 * every step -- [t][0..7] epilogue warp 0, [t][8..15] MMA-issuing thread; copies n of the 2048 x 16 stamps to the host
 * (synchronises the device).  HPMN_EINVAL when the hook is off.  Profiling aid (profiles/r2_tcrec_*.md). */
int hpmn_debug_tcr_stamps(long long* out_host, int n);

#ifdef __cplusplus
}
#endif
#endif /* HPMN_B200_H */
