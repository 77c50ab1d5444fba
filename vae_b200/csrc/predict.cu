// Multi-sample predictive mean / variance: the inference step behind
// vfm.py:1047-1057 `predict_proba` -- S variational samples of the logit of every row of x, then
//   proba_means     = mean_s likelihood.mean()[s, n]   (sigmoid(logit) / logit for the Gaussian model)
//   logit_variances = var_s logits[s, n]               (numpy .var: population variance)
// -- used by the active-learning question selection (vfm.py:1024-1045).
//
// One launch for any S: a lane group owns a sample n, keeps its rows in L1 and loops over the
// variational samples in registers; nothing of size [S, ...] is ever materialised (the reference
// builds [S, B, d] tensors).  Noise is Philox4x32-10 with the sample index in the tag word, keyed
//   per entity      (row id)        one draw per entity and sample, shared by all its occurrences --
//                                   the estimator of vfm-torch.py:238-245, or
//   per occurrence  (n * F + f)     independent draws per row of x -- vfm.py:440-445.
// The per-row marginals (what predict_proba returns) are the same under both.
#include "sampled_common.cuh"

namespace vfmb {

constexpr uint32_t kTagOccurrence = 3u;     // keeps per-occurrence streams apart from the per-entity ones

template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_predict_sampled(DevCfg c, int S, int per_occ, int pairwise, const float* __restrict__ bias,
                  const float* __restrict__ entity, const float* __restrict__ scalars,
                  const int64_t* __restrict__ x, const int32_t* __restrict__ noise_step,
                  float* __restrict__ proba_mean, float* __restrict__ logit_mean, float* __restrict__ logit_var) {
    constexpr int GPW = kWarp / LPR;
    const int d = c.d, F = c.F, B = c.B, R = c.row_stride;    // row_stride carries the table size here
    const uint32_t step = noise_step ? (uint32_t)noise_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    const float mu0 = scalars[VFMB_S_GB_MEAN], sig0 = link_fn<LINK>(scalars[VFMB_S_GB_SCALE]);

    for (int n = group; n < B; n += ngroups) {
        // lane f < F of the group owns field f's row id and bias pair
        int id_l = 0;
        float a_l = 0.f, tau_l = 0.f;
        if (gl < F) {
            int64_t id = x[(size_t)n * F + gl];
            if (id < 0 || id >= R) id = 0;
            id_l = (int)id;
            const float2 ab = *reinterpret_cast<const float2*>(bias + (size_t)id_l * 2);
            a_l = ab.x; tau_l = link_fn<LINK>(ab.y);
        }
        float mean_l = 0.f, m2 = 0.f, pm = 0.f;              // Welford over the samples (lane 0 of the group)
        for (int s = 0; s < S; ++s) {
            // bias part: lanes 0..F-1, one draw each
            float bterm = 0.f;
            if (gl < F) {
                float n4[4];
                if (per_occ) philox_normal4(c.seed, (uint32_t)(n * F + gl), 0xFFFFFFFFu, step, philox_tag(kTagOccurrence, s), n4);
                else philox_normal4(c.seed, (uint32_t)id_l, 0xFFFFFFFFu, step, philox_tag(kTagBias, s), n4);
                bterm = fmaf(n4[0], tau_l, a_l);
            }
            bterm = group_sum<LPR>(bterm, gmask);
            // factor part
            Vec<VEC> prod[NV], ssum[NV], sq[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < VEC; ++j) { prod[i].v[j] = 1.f; ssum[i].v[j] = 0.f; sq[i].v[j] = 0.f; }
            for (int f = 0; f < F; ++f) {
                const int id = __shfl_sync(gmask, id_l, f, LPR);
                const float* erow = entity + (size_t)id * 2 * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        const Vec<VEC> mu = ld_vec_nc<VEC>(erow + k), rho = ld_vec_nc<VEC>(erow + d + k);   // L1 after s = 0
                        float n4[4];
                        if (per_occ) philox_normal4(c.seed, (uint32_t)(n * F + f), (uint32_t)(k / VEC), step, philox_tag(kTagOccurrence, s), n4);
                        else philox_normal4(c.seed, (uint32_t)id, (uint32_t)(k / VEC), step, philox_tag(kTagEntity, s), n4);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float v = fmaf(n4[j], link_fn<LINK>(rho.v[j]), mu.v[j]);
                            prod[i].v[j] *= v; ssum[i].v[j] += v; sq[i].v[j] = fmaf(v, v, sq[i].v[j]);
                        }
                    }
                }
            }
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int k = (gl + i * LPR) * VEC;
                if (k < d)
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        part += pairwise ? 0.5f * (ssum[i].v[j] * ssum[i].v[j] - sq[i].v[j]) : prod[i].v[j];
            }
            part = group_sum<LPR>(part, gmask);
            const float logit = fmaf(global_eps(nullptr, c, step, s), sig0, mu0) + bterm + part;
            // Welford: running mean / sum of squared deviations of the logit, running mean of the output
            const float delta = logit - mean_l;
            mean_l += delta / (float)(s + 1);
            m2 = fmaf(delta, logit - mean_l, m2);
            const float out = (LIK == VFMB_BERNOULLI) ? 1.f / (1.f + expf(-logit)) : logit;
            pm += (out - pm) / (float)(s + 1);
        }
        if (gl == 0) {
            proba_mean[n] = pm;
            if (logit_mean) logit_mean[n] = mean_l;
            logit_var[n] = m2 / (float)S;
        }
    }
}

}  // namespace vfmb

using namespace vfmb;

extern "C" int vfmb_predict_sampled(const vfmb_config* cfg, const vfmb_tables* tab, const int64_t* x, int32_t n_samples,
                                    int32_t per_occurrence, float* proba_mean, float* logit_mean, float* logit_var,
                                    vfmb_stream stream_) {
    if (!cfg || !tab || !x || !proba_mean || !logit_var || !tab->bias || !tab->entity || !tab->scalars || !tab->noise_step)
        return set_error(VFMB_EINVAL, "vfmb_predict_sampled: null argument");
    if (cfg->B <= 0 || cfg->F < 1 || cfg->F > VFMB_MAX_FIELDS || n_samples < 1)
        return set_error(VFMB_EINVAL, "vfmb_predict_sampled: bad B / F / n_samples");
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    if (cfg->F > L.lpr) return set_error(VFMB_ESHAPE, "vfmb_predict_sampled: F = %d fields need d >= %d", cfg->F, 4 * cfg->F);
    DevCfg dc = make_dev(cfg);
    dc.row_stride = cfg->R;                                   // table size for the id range check (not sharded here)
    cudaStream_t stream = (cudaStream_t)stream_;
    const int gpb = 8 * (32 / L.lpr);
    int64_t g = ((int64_t)cfg->B + gpb - 1) / gpb;
    if (g > 8 * kNumSMs) g = 8 * kNumSMs;
    const int pairwise = (cfg->interaction == VFMB_INTER_PAIRWISE && cfg->F != 2) ? 1 : 0;
#define LAUNCH_PS(LINK, LIK)                                                                              \
    k_predict_sampled<VEC, LPR, NV, LINK, LIK><<<(int)g, 256, 0, counted(stream)>>>(                      \
        dc, n_samples, per_occurrence ? 1 : 0, pairwise, tab->bias, tab->entity, tab->scalars, x, tab->noise_step, \
        proba_mean, logit_mean, logit_var)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_PS(0, VFMB_GAUSSIAN); else LAUNCH_PS(0, VFMB_BERNOULLI);
        } else {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_PS(1, VFMB_GAUSSIAN); else LAUNCH_PS(1, VFMB_BERNOULLI);
        }
    });
#undef LAUNCH_PS
    CUDA_TRY(cudaGetLastError());
    return 0;
}
