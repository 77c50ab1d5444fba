// Shared by sampled.cu (stage / score / gather kernels, step entry points) and sampled_adam.cu
// (row-update kernels): noise helpers, the scalar-update argument block, launch preparation.
#pragma once
#include "step_common.cuh"

namespace vfmb {

constexpr int kMaxSamples = 8;      // variational samples per step (vfm-torch.py:19) the kernels take

// Noise of variational sample s (vfm-torch.py:238-241 draws [S,1], [S,U], [S,U,d]): injected arrays are
// indexed [s][unique rank] (U = number of unique rows of the batch), Philox carries s in the tag word.
template <int VEC>
__device__ __forceinline__ Vec<VEC> entity_eps(const float* __restrict__ eps_entity, const DevCfg& c,
                                              int u, int rowid, int k, uint32_t step, int s = 0, int U = 0) {
    Vec<VEC> e;
    if (eps_entity) {
        e = ld_vec_nc<VEC>(eps_entity + ((size_t)s * U + u) * c.d + k);
    } else {
        float n4[4];
        philox_normal4(c.seed, (uint32_t)rowid, (uint32_t)(k / VEC), step, philox_tag(kTagEntity, s), n4);
#pragma unroll
        for (int i = 0; i < VEC; ++i) e.v[i] = n4[i];
    }
    return e;
}
__device__ __forceinline__ float bias_eps(const float* __restrict__ eps_bias, const DevCfg& c, int u,
                                          int rowid, uint32_t step, int s = 0, int U = 0) {
    if (eps_bias) return __ldg(eps_bias + (size_t)s * U + u);
    float n4[4];
    philox_normal4(c.seed, (uint32_t)rowid, 0xFFFFFFFFu, step, philox_tag(kTagBias, s), n4);
    return n4[0];
}
__device__ __forceinline__ float global_eps(const float* __restrict__ eps_global, const DevCfg& c,
                                            uint32_t step, int s = 0) {
    if (eps_global) return __ldg(eps_global + s);
    float n4[4];
    philox_normal4(c.seed, 0xFFFFFFFFu, 0xFFFFFFFFu, step, philox_tag(kTagGlobal, s), n4);
    return n4[0];
}

// End of a sampled forward (thread 0 of the block that finishes the scoring kernel last -- every
// block has read the counter by then): record the noise index this forward drew with, advance the
// counter (see vfmb_tables.noise_step), and poison the loss when the plan saw a row id outside the
// table (the reference raises IndexError from nn.Embedding; here the id was remapped to row 0 and
// plan.meta[2] set -- a NaN loss is the device-side signal that needs no host synchronisation).
__device__ __forceinline__ void forward_done(int32_t* noise_step, uint32_t step, const int32_t* __restrict__ meta,
                                             float* __restrict__ stats) {
    if (noise_step) { noise_step[1] = (int32_t)step; noise_step[0] = (int32_t)step + 1; }
    if (meta && meta[2] != 0) {
        stats[VFMB_ST_LOSS] = __int_as_float(0x7fc00000);
        stats[VFMB_ST_NLL_MEAN] = __int_as_float(0x7fc00000);
    }
}

// arguments of the scalar-parameter update folded into the row kernels (final_scalars)
struct FinalArgs {
    float* scalars; float* sm; float* sv; float* stats; const float* eps_global;
    float* grad_scalars; double* partials; int32_t* counter; const float* gslot;
    const int32_t* noise_step;          // [1] = noise index of the forward this backward belongs to
    int likelihood;
    // mode B owner update (k_adam_rows_pipe<FLAVOR 3>): the ranks' additive scalars, one slot per rank
    const float* tail_slots; int tail_P, tail_pitch;
    int B_global; float n_train_global;
    const float* recv_grads; const int32_t* occ; int slot_pitch, n_real;   // the requesters' gradient slots
    int l2_demote;                      // bit 0 / 1 / 2: rows of entity / m / v were parked in L2 by k_stage (tuning l2_keep)
};

// Mode B / mode A: scalar parameters (replicated on every rank, identical results) and the global-batch
// loss from the ranks' additive scalars.  `tail` = the summed vector (VFMB_DP_T_*); stats[KL_ROWS] holds
// the KL of the rows.  Does not touch the step counter.
template <int LINK>
__device__ __forceinline__ void dp_scalar_update(int likelihood, int S, int B, float n_train, uint64_t seed,
                                                 float* __restrict__ scalars, float* __restrict__ sm, float* __restrict__ sv,
                                                 const float* tail, const float* __restrict__ eps_global, const AdamDev& h,
                                                 uint32_t step, uint32_t nstep, float* __restrict__ stats) {
    float alpha = scalars[VFMB_S_ALPHA], mu0 = scalars[VFMB_S_GB_MEAN], rho0 = scalars[VFMB_S_GB_SCALE];
    const float sig0 = link_fn<LINK>(rho0), ap = link_fn<LINK>(alpha);
    float e0;
    if (eps_global) e0 = eps_global[0];
    else {
        float n4[4];
        philox_normal4(seed, 0xFFFFFFFFu, 0xFFFFFFFFu, nstep, philox_tag(kTagGlobal, 0), n4);
        e0 = n4[0];
    }
    const double nll = (double)tail[VFMB_DP_T_NLL], sr = (double)tail[VFMB_DP_T_RESID], sq = (double)tail[VFMB_DP_T_SQERR];
    const float kl0 = kl_std_normal(mu0, sig0);
    const float kl = kl0 + stats[VFMB_ST_KL_ROWS];
    stats[VFMB_ST_KL] = kl;
    stats[VFMB_ST_NLL_MEAN] = (float)(nll / (double)B);
    stats[VFMB_ST_LOSS] = (float)((double)n_train * nll / (double)B + (double)kl);
    float g_mu0 = (float)(sr + (double)mu0);
    float g_rho0 = link_grad<LINK>(rho0) * (float)((double)e0 * sr + (double)(sig0 - 1.f / sig0));
    float ss, ib2;
    adam_coeffs(h, (int)step + 1, &ss, &ib2);
    adam_elem(mu0, sm[VFMB_S_GB_MEAN], sv[VFMB_S_GB_MEAN], g_mu0, h, ss, ib2);
    adam_elem(rho0, sm[VFMB_S_GB_SCALE], sv[VFMB_S_GB_SCALE], g_rho0, h, ss, ib2);
    scalars[VFMB_S_GB_MEAN] = mu0; scalars[VFMB_S_GB_SCALE] = rho0;
    if (likelihood == VFMB_GAUSSIAN) {
        double sc = (double)n_train / ((double)S * (double)B);
        float g_alpha = link_grad<LINK>(alpha) * (float)(sc * (0.5 * sq - 0.5 * (double)S * (double)B / (double)ap));
        adam_elem(alpha, sm[VFMB_S_ALPHA], sv[VFMB_S_ALPHA], g_alpha, h, ss, ib2);
        scalars[VFMB_S_ALPHA] = alpha;
    }
}

}  // namespace vfmb

using namespace vfmb;

// ---- shared host-side preparation of one phase launch
struct Prep {
    cudaStream_t stream;
    Layout L;
    DevCfg dc;
    vfmb_plan_capacity_t cap;
    int ch;
};
static inline int prep(const vfmb_config* cfg, const char* who, vfmb_stream stream_, int min_fields, Prep* p) {
    if (!cfg) return set_error(VFMB_EINVAL, "%s: null config", who);
    if (cfg->B <= 0 || cfg->R <= 0 || cfg->d <= 0) return set_error(VFMB_EINVAL, "%s: bad B/R/d", who);
    if (cfg->F < min_fields || cfg->F > VFMB_MAX_FIELDS) return set_error(VFMB_ESHAPE, "%s: F must be %d..%d", who, min_fields, VFMB_MAX_FIELDS);
    if (cfg->S < 1 || cfg->S > kMaxSamples) return set_error(VFMB_ESHAPE, "%s: S=%d variational samples (1..%d)", who, cfg->S, kMaxSamples);
    if (cfg->n_classes < 1 || cfg->n_classes > VFMB_MAX_FIELDS) return set_error(VFMB_EINVAL, "%s: bad n_classes", who);
    if (cfg->likelihood != VFMB_GAUSSIAN && cfg->likelihood != VFMB_BERNOULLI) return set_error(VFMB_EINVAL, "%s: bad likelihood", who);
    if (cfg->link != VFMB_LINK_ABS && cfg->link != VFMB_LINK_SOFTPLUS) return set_error(VFMB_EINVAL, "%s: bad link", who);
    p->stream = (cudaStream_t)stream_;
    if (!pick_layout(cfg->d, &p->L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    p->dc = make_dev(cfg);
    int rc = vfmb_plan_capacity(cfg->B, cfg->F, cfg->R, &p->cap);
    if (rc) return rc;
    p->ch = kRounds * (32 / p->L.lpr);
    return 0;
}


#define VFMB_PHASE_S1(who)                                                                             \
    if (cfg && cfg->S != 1) return set_error(VFMB_ESHAPE, who ": the phase entry points take S = 1 (use "    \
                                             "vfmb_sampled_forward / _backward / _step for S > 1)")


// mode B owner update: where the ranks' additive scalars sit and what the global batch is
struct DpTail {
    const float* tail_slots; int P, pitch;
    int B_global; float n_train_global;
    float* stats_out; const float* eps_global;
    const float* recv_grads; int slot_pitch, n_real;
};
// row update (sampled_adam.cu); flavor: see k_adam_rows (3: mode B owner, needs `dp`)
int launch_adam(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan, const vfmb_step_io* io,
                const vfmb_adam* adam, int32_t mode, float kl_grad_scale, int flavor, vfmb_stream stream_,
                const DpTail* dp = nullptr);
int launch_adam_multi(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan, const vfmb_step_io* io,
                      const vfmb_adam* adam, int32_t mode, float kl_grad_scale, vfmb_stream stream_);
