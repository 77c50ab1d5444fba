/* vfm_b200.h -- C ABI of the B200-native Variational Factorization Machine step.
 *
 * The reference (jilljenn/vae) has no FFI for this path: its boundary is the
 * PyTorch model API used by the training loops of vfm-torch.py and
 * vfm-tomasrch.py.  This library is what a binding for that path would call;
 * each entry point names the reference lines it replaces.  The host-side
 * mirror of the reference's `class CF` lives in vae_b200/ (Python, ctypes).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked host;
 *     the library allocates nothing persistent and frees nothing;
 *   - tables are row-major contiguous fp32, base pointers 16-byte aligned:
 *       bias   [R,2]  = [mean, raw scale]            (vfm-torch.py:152, vfm-tomasrch.py:229)
 *       entity [R,2d] = [mean(d) | raw scale(d)]     (vfm-torch.py:153, vfm-tomasrch.py:251)
 *   - all work is enqueued on the given stream; no call synchronises, so a
 *     whole step can be captured into a CUDA graph;
 *   - return 0 on success, otherwise a cudaError_t or a VFMB_E* code;
 *     vfmb_last_error() gives the message (thread-local).  Nothing throws.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with the CUDA error.
 */
#ifndef VFM_B200_H
#define VFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFMB_MAX_FIELDS 8

#define VFMB_EINVAL 10001 /* bad argument (message says which)              */
#define VFMB_ESHAPE 10002 /* unsupported width / field count                */
#define VFMB_ESPACE 10003 /* workspace too small                            */

enum { VFMB_GAUSSIAN = 0, VFMB_BERNOULLI = 1 };           /* vfm-torch.py:267-270 */
enum { VFMB_LINK_ABS = 0, VFMB_LINK_SOFTPLUS = 1 };       /* vfm-torch.py:125-126 */
enum { VFMB_ADAM_TOUCHED = 0, VFMB_GRAD_ONLY = 1 };       /* backward epilogue     */
/* factor interaction of the plan-free prediction entry points: the product over the fields
 * (vfm-torch.py:245 / vfm-tomasrch.py:343 `prod(axis)`; the scripts' own formula) or the pairwise FM
 * sum_{i<j} <v_i, v_j> (vfm-tomasrch.py:379-393, vfm.py:467-474).  Equal for F = 2.  The sampled
 * training step always scores F > 2 fields pairwise. */
enum { VFMB_INTER_PROD = 0, VFMB_INTER_PAIRWISE = 1 };

typedef void* vfmb_stream; /* cudaStream_t */

/* Problem description shared by every call. */
typedef struct vfmb_config {
    int32_t B;            /* samples in the batch                                  */
    int32_t F;            /* fields (columns of x), 2..VFMB_MAX_FIELDS             */
    int32_t d;            /* embedding size                                        */
    int32_t R;            /* table rows                                            */
    int32_t S;            /* variational samples (vfm-torch.py:19), 1 ... 8; S > 1: the
                             per-row scratch of vfmb_step_io is [S][u_cap,...], pred / mean
                             are [S,B], the noise arrays [S], [S,U], [S,U,d]              */
    int32_t likelihood;   /* VFMB_GAUSSIAN | VFMB_BERNOULLI                        */
    int32_t link;         /* VFMB_LINK_*                                           */
    int32_t n_classes;    /* KL weighting classes (<= F)                           */
    /* class(u) = #{i : class_bound[i] <= u}.  vfm-torch.py:316 is {N+1} with
     * sizes {N,M} (the `<= N` selection); vfm-tomasrch.py:574-587 is the field
     * offsets {G_0, G_0+G_1, ...} with sizes {G_g}.                              */
    int32_t class_bound[VFMB_MAX_FIELDS];
    float class_size[VFMB_MAX_FIELDS];
    float n_train;        /* nb_train_samples (vfm-torch.py:91,359)                */
    int32_t row_stride;   /* row-sharded tables: global id of local row r is        */
    uint64_t seed;        /* Philox key                                            */
    int32_t row_offset;   /*   r*row_stride + row_offset (0 / 0 mean 1 / 0); the    */
    int32_t interaction;  /*   noise and the KL class are keyed by the GLOBAL id.
                             interaction: VFMB_INTER_* of vfmb_predict_mean / vfmb_predict_sampled */
} vfmb_config;

/* Parameter tables and Adam state.  m/v may be NULL for forward-only use. */
typedef struct vfmb_tables {
    float* bias;      float* bias_m;   float* bias_v;    /* [R,2]  */
    float* entity;    float* entity_m; float* entity_v;  /* [R,2d] */
    const float* train_counts;                           /* [R] bincount(X_train) as fp32
                                                            (vfm-torch.py:89, vfm-tomasrch.py:182) */
    /* scalar parameter block + Adam state, layout per variant (see below) */
    float* scalars;   float* scalars_m; float* scalars_v;
    /* device step counter: int32[1], number of Adam steps already applied */
    int32_t* adam_step;
    /* device noise counter: int32[2] = {index the NEXT sampled forward draws with, index the LAST
     * forward drew with}.  It is the `step` word of the Philox counter.  Every sampled forward
     * (vfmb_sampled_forward / _score / _step) reads [0], records it in [1] and advances [0] when its
     * last kernel finishes, so consecutive forwards never share noise whether or not an Adam step
     * lies between them (the reference draws fresh noise in every forward, vfm-torch.py:238-241,
     * also at test time :402-403); backward-side kernels re-create the draws of their forward from
     * [1].  Required by the sampled entry points, unused by the closed form. */
    int32_t* noise_step;
} vfmb_tables;

/* sampled variant scalar block (vfm-torch.py:136-138): */
enum { VFMB_S_ALPHA = 0, VFMB_S_GB_MEAN = 1, VFMB_S_GB_SCALE = 2, VFMB_S_COUNT = 4 };
/* closed-form variant scalar block (vfm-tomasrch.py:194-248): five scalars,
 * then per group g: bias prior mean, bias prior scale, then entity prior
 * mean[d] and scale[d]; use vfmb_closed_scalar_count / offsets below. */
enum { VFMB_C_ALPHA = 0, VFMB_C_GB_MEAN = 1, VFMB_C_GB_SCALE = 2, VFMB_C_GB_PRIOR_MEAN = 3,
       VFMB_C_GB_PRIOR_SCALE = 4, VFMB_C_GROUP_BASE = 8 };
/* offsets inside the closed-form scalar block */
int32_t vfmb_closed_off_bias_prior_mean(int32_t G, int32_t d, int32_t g);
int32_t vfmb_closed_off_bias_prior_scale(int32_t G, int32_t d, int32_t g);
int32_t vfmb_closed_off_entity_prior_mean(int32_t G, int32_t d, int32_t g);
int32_t vfmb_closed_off_entity_prior_scale(int32_t G, int32_t d, int32_t g);
int32_t vfmb_closed_scalar_count(int32_t G, int32_t d);

typedef struct vfmb_adam {
    /* torch.optim.Adam hyper-parameters (vfm-torch.py:339, vfm-tomasrch.py:518; betas
     * (0.9, 0.999), eps 1e-8 by default).  Doubles, because torch forms 1-beta1, 1-beta2
     * and the bias corrections in double before rounding to fp32 -- (1 - 0.999f) is off
     * by 1.3e-5 relative. */
    double lr, beta1, beta2, eps;
} vfmb_adam;

/* The batch plan: replaces the torch.unique calls of vfm-torch.py:190-192 and
 * vfm-tomasrch.py:537-545.  All arrays are caller-allocated device memory of
 * the capacities given by vfmb_plan_capacity().  Integer results are
 * bit-identical to torch.unique(sorted=True, return_inverse, return_counts). */
typedef struct vfmb_plan {
    int32_t* uniq;       /* [U_cap]   sorted unique row ids                        */
    int32_t* inverse;    /* [B*F]     rank of x[n,f] in uniq                       */
    int32_t* seg_off;    /* [U_cap+1] segment offsets; counts[u]=seg_off[u+1]-seg_off[u] */
    int32_t* occ;        /* [B*F]     occurrence ids n*F+f grouped by rank, ascending inside */
    int32_t* pos_of;     /* [B*F]     inverse permutation of occ: sorted position of (n,f)   */
    int32_t* pos_rank;   /* [B*F]     unique rank of each sorted position                    */
    int32_t* partner;    /* [B*F]     per sorted position: F==2 the rank of the sample's other
                                      field; F>2 the sample index n                          */
    int32_t* urec;       /* [U_cap,4] per unique row {row id, segment length, seg_off, batch count} */
    int32_t* class_off;  /* [VFMB_MAX_FIELDS+1] first unique rank of each KL class (class_bound
                                      of the config); class_off[n_classes] = U               */
    float* z;            /* [VFMB_MAX_FIELDS] per-column normaliser Z_f             */
    int32_t* meta;       /* [8] 0:U 2:error flag (id out of range) 3:number of hot rows 5:number of
                                      other cut rows; rest reserved                           */
    int32_t* hot;        /* [cut_rows_cap of vfmb_plan_capacity] unique ranks of the rows cut by
                                      backward-tile boundaries: <= 32 tiles from the front, more
                                      ("hot" rows) from the back; order unspecified           */
} vfmb_plan;

typedef struct vfmb_plan_capacity_t {
    int64_t u_cap, n_tiles, workspace_bytes;
    int32_t tile;        /* sorted positions per backward tile */
    int32_t cut_rows_cap; /* entries of vfmb_plan.hot */
} vfmb_plan_capacity_t;

int vfmb_plan_capacity(int32_t B, int32_t F, int32_t R, vfmb_plan_capacity_t* out /*host*/);

/* Build the plan for one batch.  x: int64 [B,F] global row ids (the reference's
 * LongTensor batch, vfm-torch.py:88,351).  train_counts is only used for the
 * per-column normalisers Z_f = sum_{u in uniq(x[:,f])} cnt_f(u)/cnt_train(u)
 * (vfm-torch.py:305-306, vfm-tomasrch.py:578-581). */
int vfmb_plan_build(const vfmb_config* cfg /*host*/, const int64_t* x, const float* train_counts,
                    const vfmb_plan* plan /*host struct of device ptrs*/, void* workspace,
                    size_t workspace_bytes, vfmb_stream stream);

/* Scratch + outputs of one step.  Capacities: vs/es/grow [U_cap*d], ws/ebs/cq/gws [U_cap],
 * msg [B*d] (F>2 only), pred/mean/resid [B], rsorted [B*F]; the closed-form variant uses
 * vs [U_cap*2d] (mean | raw scale^2), grow [U_cap*3d], msg [B*3d],
 * partials [vfmb_partials_doubles()], stats [VFMB_STATS]. */
typedef struct vfmb_step_io {
    const float* y;          /* [B] targets (NULL: forward without likelihood terms)  */
    const float* eps_global; /* [S]     injected N(0,1) draws, or NULL -> Philox      */
    const float* eps_bias;   /* [S,U]   indexed by unique rank (vfm-torch.py:239)     */
    const float* eps_entity; /* [S,U,d] (vfm-torch.py:241)                            */
    float* vs;               /* scratch: sampled factor rows  [S][U_cap,d]            */
    float* ws;               /* scratch: sampled biases       [U_cap]                 */
    float* es;               /* scratch: the step's factor noise [U_cap,d] (Philox)   */
    float* ebs;              /* scratch: the step's bias noise   [U_cap]   (Philox)   */
    float* cq;               /* scratch: KL weight c_u per unique row [U_cap]         */
    float* grow;             /* scratch: dloss/dv per unique row [U_cap,d]            */
    float* gws;              /* scratch: dloss/dw per unique row [U_cap]              */
    float* msg;              /* scratch [B,d], only F>2 (may be NULL for F==2)        */
    float* pred;             /* [S,B] unscaled_pred (vfm-torch.py:265)                */
    float* mean;             /* [S,B] likelihood mean: pred or sigmoid(pred) (:363)   */
    float* resid;            /* [B] dloss/dpred (S > 1: mean over s of dloss/dpred[s,n]) */
    float* rsorted;          /* [B*F] dloss/dpred per sorted occurrence (scratch)     */
    double* partials;        /* scratch for deterministic reductions                  */
    int32_t* counters;       /* [8] zero-initialised once by the caller               */
    float* stats;            /* [VFMB_STATS] see enum                                 */
    float* kl_bias_out;      /* optional [U]   per-row bias KL   (closed form, kls[1])  */
    float* kl_entity_out;    /* optional [U,d] per-row factor KL (closed form, kls[2])  */
    float* grad_bias;        /* VFMB_GRAD_ONLY: dense [R,2]  (touched rows written)   */
    float* grad_entity;      /* VFMB_GRAD_ONLY: dense [R,2d]                          */
    float* grad_scalars;     /* VFMB_GRAD_ONLY: [scalar count]                        */
} vfmb_step_io;

enum { VFMB_ST_LOSS = 0,      /* n_train*mean(nll) + kl           (vfm-torch.py:359) */
       VFMB_ST_NLL_MEAN = 1,  /* mean over batch of -log_prob                        */
       VFMB_ST_KL = 2,        /* KL(global) + rescaled row KL     (vfm-torch.py:320-324) */
       VFMB_ST_SUM_RESID = 3, VFMB_ST_SUM_SQERR = 4, VFMB_ST_KL_ROWS = 5,
       VFMB_ST_W0 = 6, VFMB_ST_U = 7,
       VFMB_ST_RESID_S = 8,   /* [8] S > 1: sum_n dloss/dpred[s, n] per variational sample  */
       VFMB_ST_W0_S = 16,     /* [8] S > 1: sampled global bias per variational sample      */
       VFMB_STATS = 32 };

/* size (in doubles) of vfmb_step_io.partials for this problem */
int64_t vfmb_partials_doubles(const vfmb_config* cfg /*host*/);

/* Sampled ELBO, forward: gather of the unique rows, reparameterised draw,
 * FM interaction, likelihood and KL -- vfm-torch.py:200-324 (+ :359 when y given). */
int vfmb_sampled_forward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                         const vfmb_step_io* io, vfmb_stream stream);

/* Sampled ELBO, backward: deterministic segmented reduction over the sorted
 * occurrence list, chain rule to (mean, raw scale) + KL gradient, then either
 * Adam on the touched rows (mode VFMB_ADAM_TOUCHED) or dense gradient output
 * (VFMB_GRAD_ONLY).  Replaces loss.backward(); optimizer.step() of
 * vfm-torch.py:368-370.  io->resid must hold dloss/dpred (written by the
 * forward when y was given, or supplied by the caller's autograd). */
int vfmb_sampled_backward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                          const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                          float kl_grad_scale, vfmb_stream stream);

/* The phases of the sampled step as separate entry points (the row-sharded multi-GPU mode
 * runs them on different ranks with all-to-all exchanges in between; F may be 1 here):
 *   stage      k_stage       unique rows -> vs, ws, es, ebs, cq, KL sum       (needs plan->urec, z)
 *   score      k_score       samples -> pred, mean, resid, rsorted, msg       (needs vs, ws)
 *   gather     k_gather      sorted occurrences -> grow, gws (rows cut by tile boundaries are finished in
 *              the kernel).  unit_coef != 0: rows of `table` are added
 *              unscaled and gws sums io->rsorted -- the owner-side ordered sum of the
 *              gradient rows received from the other ranks (table [B*F, d])
 *   adam_rows  k_adam_rows   chain rule + KL gradient + Adam / dense gradients
 *   dp_final   scalar parameters + loss from the all-reduced tail (see mode A)            */
int vfmb_sampled_stage(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                       const vfmb_step_io* io, vfmb_stream stream);
int vfmb_sampled_score(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                       const vfmb_step_io* io, vfmb_stream stream);
int vfmb_sampled_gather(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io,
                        const float* table, int32_t unit_coef, vfmb_stream stream);
int vfmb_sampled_adam_rows(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                           const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                           float kl_grad_scale, vfmb_stream stream);
/* ---- mode B (row-sharded tables, owner = row mod P): glue between the plan / step kernels and the
 * three all-to-alls of a step.  No reference counterpart (the reference is single-process); the
 * contract is that the sharded step on the global batch equals the single-process step.
 * Slot layout of every exchange buffer: [P, CAP, w] -- slot q*CAP + j holds the j-th unique id
 * (ascending) this rank asks of owner q; w = 2 int32 {id, local batch count} for requests (-1 =
 * empty), d+1 floats {row, bias} for sampled rows and for row gradients (zeros = empty).         */
int64_t vfmb_shard_bucket_workspace(int32_t u_cap);
/* Every pack function takes an optional PEER TABLE: `peers` = host array of P device pointers, the
 * same exchange buffer of every rank mapped into this process (NVLink peer memory, e.g. torch
 * symmetric memory), `rank` = this process.  With a peer table the kernel stores slot (q, j) straight
 * into chunk `rank` of rank q's buffer -- the pack kernel IS the all-to-all; the caller only needs a
 * cross-rank barrier before the consumer.  With peers == NULL the slots go to the local `send` /
 * `reply` / `out` buffer and the caller runs a collective (NCCL all-to-all) on it.
 *
 * requester: unique ids of `plan` -> request slots [P*CAP,2]; dest[u] = slot of unique rank u
 * (P*CAP = none); *overflow |= 1 when an owner's bucket exceeds CAP                               */
int vfmb_shard_bucket(const vfmb_plan* plan, int32_t u_cap, int32_t P, int32_t CAP, int32_t* send,
                      int32_t* dest, int32_t* overflow, void* workspace, const void* const* peers,
                      int32_t rank, vfmb_stream stream);
/* peer mode all-reduce of a small vector: put `vec[n]` into slot `rank` (pitch floats apart) of
 * every rank; after the barrier every rank adds the P slots in rank order (bitwise equal results) */
int vfmb_shard_put_small(const float* vec, int32_t n, int32_t pitch, const void* const* peers, int32_t P,
                         int32_t rank, vfmb_stream stream);
int vfmb_shard_sum_small(const float* slots, int32_t P, int32_t n, int32_t pitch, float* out, vfmb_stream stream);
/* owner: received requests -> local row indices loc [M] (int64, the plan's input; padding -> R_loc);
 * recv_copy (optional) = private copy of the requests                                              */
int vfmb_shard_owner_ids(const int32_t* recv, int32_t M, int32_t P, int32_t R_loc, int64_t* loc,
                         int32_t* recv_copy, vfmb_stream stream);
/* owner: counts_only != 0 -> urec[u].w = batch count of row u summed over the requesters;
 *        counts_only == 0 -> reply slot s = {vs[inverse[s]], ws[inverse[s]]}                        */
int vfmb_shard_owner_pack(const vfmb_plan* plan_o, const int32_t* recv, int32_t M, int32_t CAP, int32_t d,
                          const float* vs, const float* ws, float* reply, int32_t counts_only,
                          const void* const* peers, int32_t rank, vfmb_stream stream);
/* requester: received sampled rows -> vs / ws in unique-rank order                                  */
int vfmb_shard_unpack_rows(const vfmb_plan* plan_l, const float* recv_rows, const int32_t* dest,
                           int32_t u_cap, int32_t M, int32_t d, float* vs, float* ws, vfmb_stream stream);
/* requester: row gradients -> slots; tail[tail_idx[0..3]] = {NLL sum, residual sum, squared-error
 * sum of this rank's samples, KL sum of this rank's owned rows} (tail_idx: host array of 4)         */
int vfmb_shard_pack_grads(const vfmb_plan* plan_l, const float* grow, const float* gws, const int32_t* dest,
                          int32_t u_cap, int32_t M, int32_t CAP, int32_t d, float* out, const float* stats_local,
                          const float* stats_owner, float n_local, float* tail, const int32_t* tail_idx,
                          int32_t n_tail, const void* const* peers, int32_t rank, vfmb_stream stream);
/* owner: received gradient slots -> gather table [M,d] + bias gradients in sorted-occurrence order;
 * recv_ids (peer mode) = the requests: empty slots were never written and read as zeros            */
int vfmb_shard_unpack_grads(const vfmb_plan* plan_o, const float* recv_g, const int32_t* recv_ids, int32_t M,
                            int32_t d, float* table, float* rsorted, vfmb_stream stream);

/* ---- mode B, fused over NVLink peer memory: the exchanges are written / read in place by the step
 * kernels themselves, two cross-rank barriers per step remain (after the rows, after the gradients).
 * Slot pitch SP = d + 4 floats ([row | scalar, pad]: 16-byte aligned for 128-bit peer stores); the
 * received-rows region has one spare slot M (zeros) that overflowed rows are routed to.
 *   vfmb_shard_route         requester, id-only: per occurrence / sorted position the slot its row is read
 *                            from, per unique rank the peer address its gradient row is stored to
 *   vfmb_shard_stage_put     owner: k_stage; every sampled row goes straight into its requesters' slots
 *                            (n_real = rows this rank owns; larger local indices are padding)
 *   vfmb_shard_score         requester: k_score on the received slots; the block that finishes last puts the
 *                            rank's additive scalars {NLL, residual, squared error, KL of its owned rows,
 *                            overflow flag} into slot `rank` of every rank's tail region
 *   vfmb_shard_gather_put    requester: k_gather on the received slots; finished gradient rows are stored
 *                            through gptr into their owners' slots
 *   vfmb_shard_owner_update  owner, ONE kernel: a row's gradient = the slots its requesters stored, added in
 *                            source-rank order; Adam on the owned rows; scalar parameters + global-batch
 *                            loss from the tail slots (NaN when any rank flagged a bucket overflow)      */
int vfmb_shard_route(const vfmb_plan* plan_l, const int32_t* dest, int32_t B, int32_t F, int32_t u_cap,
                     int32_t M, int32_t CAP, int32_t SP, const void* const* peers_grads, int32_t P,
                     int32_t rank, float* dump, int32_t* inv_slot, int32_t* partner_slot, float** gptr,
                     vfmb_stream stream);
int vfmb_shard_stage_put(const vfmb_config* cfg_o, const vfmb_tables* tab, const vfmb_plan* plan_o,
                         const vfmb_step_io* io_o, int32_t CAP, int32_t SP, int32_t n_real,
                         const void* const* peers_rows, int32_t P, int32_t rank, vfmb_stream stream);
int vfmb_shard_score(const vfmb_config* cfg_l, const vfmb_tables* tab, const vfmb_plan* plan_l,
                     const vfmb_step_io* io_l, const float* recv_rows, int32_t SP, const int32_t* inv_slot,
                     const void* const* peers_tail, int32_t P, int32_t rank, const float* stats_owner,
                     const int32_t* overflow, float n_local, int32_t tail_pitch, vfmb_stream stream);
int vfmb_shard_gather_put(const vfmb_config* cfg_l, const vfmb_plan* plan_l, const vfmb_step_io* io_l,
                          const float* recv_rows, int32_t SP, const int32_t* partner_slot,
                          float* const* gptr, const int32_t* own_slot, vfmb_stream stream);
int vfmb_shard_owner_update(const vfmb_config* cfg_o, const vfmb_tables* tab, const vfmb_plan* plan_o,
                            const vfmb_step_io* io_o, const vfmb_adam* adam, const float* recv_grads,
                            int32_t SP, int32_t n_real, const float* tail_slots, int32_t P, int32_t tail_pitch,
                            int32_t B_global, float n_train_global, float* stats_out,
                            const float* eps_global, vfmb_stream stream);

int vfmb_dp_final(const vfmb_config* cfg_global, const vfmb_tables* tab, const float* tail,
                  const float* eps_global, const vfmb_adam* adam, float* stats, vfmb_stream stream);

/* Forward (with targets) + backward + Adam on the touched rows in one call: the whole training
 * step of vfm-torch.py:351-370 after the plan.  Equivalent to vfmb_sampled_forward followed by
 * vfmb_sampled_backward(mode VFMB_ADAM_TOUCHED, kl_grad_scale 1). */
int vfmb_sampled_step(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                      const vfmb_step_io* io, const vfmb_adam* adam, vfmb_stream stream);

/* Dense Adam sweep over whole tables given dense gradients: the reference's
 * torch.optim.Adam semantics (every row moves every step). */
int vfmb_adam_dense(float* p, float* m, float* v, const float* g, int64_t n,
                    const vfmb_adam* adam, const int32_t* adam_step /*device*/,
                    vfmb_stream stream);
int vfmb_adam_step_advance(int32_t* adam_step, vfmb_stream stream);

/* ---- batch data-parallel mode (replicated small tables, dense gradient all-reduce) -------------
 * The reference is single-process; under DP the step on the GLOBAL batch must equal the
 * single-process step on that batch.  Per rank: forward/backward on the local slice
 * with kl_grad_scale 0 and n_train scaled by B_local/B_global (VFMB_GRAD_ONLY, dense gradients),
 * vfmb_dp_scatter_counts, one all-reduce(sum) over [grad_entity | grad_bias | counts | tail],
 * then vfmb_dp_apply_sampled on every rank: KL gradient with the global batch counts and
 * normalisers, Adam (dense = the reference's torch.optim.Adam over every row, or touched rows),
 * scalar parameters, step counter.  tail = {z[0..7], sum nll, sum resid, sum sq. err}. */
enum { VFMB_DP_TAIL = 16, VFMB_DP_T_NLL = 8, VFMB_DP_T_RESID = 9, VFMB_DP_T_SQERR = 10,
       VFMB_DP_T_KLROWS = 11,    /* mode B: KL of the rows a rank owns                         */
       VFMB_DP_T_OVERFLOW = 12   /* mode B: > 0 when any rank's request bucket overflowed       */ };
int vfmb_dp_scatter_counts(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io,
                           float* counts /*[R], zeroed*/, float* tail /*[VFMB_DP_TAIL]*/,
                           vfmb_stream stream);
int vfmb_dp_apply_sampled(const vfmb_config* cfg_global, const vfmb_tables* tab, const float* grad_entity,
                          const float* grad_bias, const float* counts, const float* tail,
                          const float* eps_global, const vfmb_adam* adam, int32_t dense_adam,
                          double* partials, int32_t* counters, float* stats, vfmb_stream stream);

/* Closed-form Gaussian variant (vfm-tomasrch.py:323-453 forward,
 * :569-594 loss/backward/Adam).  Same plan, same scratch. */
int vfmb_closed_forward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                        const vfmb_step_io* io, vfmb_stream stream);
int vfmb_closed_backward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                         const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                         vfmb_stream stream);
/* The same backward for a caller whose autograd assembles the loss itself (the loop of
 * vfm-tomasrch.py:569-588 run unchanged on the drop-in module): the upstream gradients arrive as
 *   kl_weight [U]  d loss / d (kls[1][u] + sum_k kls[2][u,k])  per unique row (NULL: the c_u of :574-587),
 *   data_scale     d loss / d (-partial_loss)                  (N_train / B in the script),
 *   kl0_scale      weight of kls[0] in the scalar-parameter gradients (0 when autograd differentiates
 *                  kls[0] itself).
 * vfmb_closed_backward is this call with (NULL, n_train / B, 1). */
int vfmb_closed_backward_weighted(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                  const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                  const float* kl_weight, float data_scale, float kl0_scale,
                                  vfmb_stream stream);

/* Posterior-mean prediction without a plan: global + sum of bias means + interaction of the factor
 * means -- cfg->interaction: product over the fields (vfm-torch.py:248-259 last_logits,
 * vfm-tomasrch.py:342-348) or pairwise (the model a sampled step with F > 2 fields optimises). */
int vfmb_predict_mean(const vfmb_config* cfg, const float* bias, const float* entity,
                      float global_bias, const int64_t* x, float* out, vfmb_stream stream);

/* Multi-sample predictive statistics without a plan (vfm.py:1047-1057 `predict_proba`, behind the
 * active-learning question selection of vfm.py:1024-1045): n_samples variational samples of the
 * logit of every row of x (int64 [cfg->B, cfg->F]) in ONE launch, nothing of size [S, ...] materialised:
 *   proba_mean[n] = mean_s likelihood.mean()[s, n]  (sigmoid(logit) for Bernoulli, the logit for Gaussian)
 *   logit_mean[n] = mean_s logit[s, n]              (optional, may be NULL)
 *   logit_var[n]  = var_s logit[s, n]               (population variance, as numpy .var)
 * Noise: Philox, sample index in the tag word, drawn at noise index tab->noise_step[0] (the caller
 * advances the counter afterwards, vfmb_adam_step_advance works on it).  per_occurrence = 0: one draw
 * per entity and sample, shared by all its occurrences (vfm-torch.py:238-245); 1: independent draws
 * per row of x (vfm.py:440-445).  Interaction per cfg->interaction.  Needs tab->bias, entity,
 * scalars (sampled layout), noise_step. */
int vfmb_predict_sampled(const vfmb_config* cfg, const vfmb_tables* tab, const int64_t* x, int32_t n_samples,
                         int32_t per_occurrence, float* proba_mean, float* logit_mean, float* logit_var,
                         vfmb_stream stream);

/* The N(0,1) draws the Philox path uses for `step` (for tests / reproducibility):
 * eps_bias [U], eps_entity [U,d] for the given unique row ids. */
int vfmb_philox_normals(const vfmb_config* cfg, const int32_t* uniq, int32_t U, int32_t step,
                        float* eps_global, float* eps_bias, float* eps_entity, vfmb_stream stream);

/* Measurement hook: when both are non-NULL (cudaEvent_t), the following backward calls of this
 * thread record them on the launch stream immediately before and after the step's dominant
 * kernel (k_adam_rows / k_cadam); pass NULLs to switch it off.  Used by bench.py for the
 * roofline of that kernel; no effect on results. */
int vfmb_profile_events(void* start_event, void* stop_event);

/* Number of kernels this library has enqueued (or captured into a CUDA graph) in this process so
 * far: bench.py derives its `gpu_launches` from it (per captured graph x replays + eager launches). */
int64_t vfmb_launch_count(void);

/* Grid sizing of the (persistent, one-resident-wave) step kernels: leave `blocks_per_sm` block
 * slots per SM unused.  Set to 1 while enqueueing / capturing steps that run concurrently with
 * vfmb_plan_build on another stream -- the plan's small blocks then start at once instead of
 * displacing step blocks at every kernel boundary (ml20m: 121 -> 116 us per step).  Process-wide;
 * affects launches made after the call.  Default 0. */
int vfmb_set_grid_reserve(int blocks_per_sm);

/* Other process-wide launch knobs (host side only).  Bit-identical results whatever the value
 * (tests/test_gpu_sampled.py): "grid_reserve" (as above), "adam_reserve" (0 / 1: the row update leaves the
 * reserved slot free too), "adam_pipe" (1 / 0: cp.async-pipelined or register-staged row update), "l2_keep"
 * (bit mask: rows of entity / m / v parked in L2 between the sampling kernel and the row update),
 * "prefetch_mv" (bit mask: k_stage prefetches the Adam moments), "gather_dyn" (1 / 0: block tiles handed
 * out by a counter or a fixed stride), "gather_fence" (1 / 0: fence.acq_rel or fence.sc in the cut-row
 * finisher), "stage_wide" (0 / 1: lane mapping of k_stage), "stage_chunk" / "score_chunk" (16 / 32 units
 * per warp pass), "pdl" (0 / 1: programmatic dependent launch along the kernel chains), "gather_keep"
 * (closed-form gather: rows up to this many occurrences are never cut).  Another fixed association of one
 * sum each (last-bit differences; set once per process): "gather_wide" (1 / 0), "score_wide" (-1 = when
 * F > 2 / 0 / 1).  Defaults are the measured best.  Unknown keys are an error. */
int vfmb_set_tuning(const char* key, int value);

const char* vfmb_last_error(void);
int vfmb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VFM_B200_H */
