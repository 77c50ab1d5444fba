// Batch data-parallel mode for replicated (small) tables: the pieces that run after the dense
// gradient all-reduce (mode A).  The collective itself is NCCL through
// torch.distributed; there is no reference counterpart (the reference is single-process), so
// parity is defined against the single-process step on the global batch.
#include "sampled_common.cuh"

namespace vfmb {

// local batch counts -> dense [R] (the all-reduce turns them into global batch counts), and the
// additive scalars of the local forward into the tail of the reduced buffer
__global__ void __launch_bounds__(256)
k_dp_scatter_counts(const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
                    const float* __restrict__ z, const float* __restrict__ stats, int B, int F,
                    float* __restrict__ counts, float* __restrict__ tail) {
    const int U = meta[0];
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
        const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + u);
        counts[rec.x] = (float)rec.y;
    }
    if (blockIdx.x == 0 && threadIdx.x < VFMB_DP_TAIL) {
        float v = 0.f;
        if (threadIdx.x < F) v = z[threadIdx.x];
        else if (threadIdx.x == VFMB_DP_T_NLL) v = stats[VFMB_ST_NLL_MEAN] * (float)B;
        else if (threadIdx.x == VFMB_DP_T_RESID) v = stats[VFMB_ST_SUM_RESID];
        else if (threadIdx.x == VFMB_DP_T_SQERR) v = stats[VFMB_ST_SUM_SQERR];
        tail[threadIdx.x] = v;
    }
}

// dense sweep over the table after the all-reduce: KL gradient with the GLOBAL batch count and
// normalisers, Adam (every row, or only rows present in the global batch), KL value
template <int VEC, int LPR, int NV, int LINK>
__global__ void __launch_bounds__(256)
k_dp_apply(DevCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
           float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
           const float* __restrict__ train_counts, const float* __restrict__ g_entity,
           const float* __restrict__ g_bias, const float* __restrict__ counts, const float* __restrict__ tail,
           int R, AdamDev h, const int32_t* __restrict__ adam_step, int dense_adam,
           double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR;
    const int d = c.d;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    __shared__ float s_coef[2];
    if (threadIdx.x == 0) adam_coeffs(h, adam_step[0] + 1, &s_coef[0], &s_coef[1]);
    __syncthreads();
    const float step_size = s_coef[0], inv_bc2 = s_coef[1];
    float facc = 0.f;
    for (int r = group; r < R; r += ngroups) {
        const float cnt = __ldg(counts + r);
        if (cnt == 0.f && !dense_adam) continue;
        float cfac = 0.f;
        if (cnt > 0.f) {
            const int cls = class_of(c, r);
            float csz = 0.f, zc = 1.f;
#pragma unroll
            for (int i = 0; i < kMaxFields; ++i) if (i == cls) { csz = c.class_size[i]; zc = __ldg(tail + i); }
            cfac = (cnt / __ldg(train_counts + r)) * (csz / zc);
        }
        const size_t eoff = (size_t)r * 2 * d, boff = (size_t)r * 2;
        float kl = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                Vec<VEC> m1 = ld_vec<VEC>(entity_m + eoff + k), m2 = ld_vec<VEC>(entity_m + eoff + d + k);
                Vec<VEC> v1 = ld_vec<VEC>(entity_v + eoff + k), v2 = ld_vec<VEC>(entity_v + eoff + d + k);
                const Vec<VEC> g1 = ld_vec<VEC>(g_entity + eoff + k), g2 = ld_vec<VEC>(g_entity + eoff + d + k);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float gm = g1.v[j], gr = g2.v[j];
                    if (cnt > 0.f) {
                        const float sig = link_fn<LINK>(rho.v[j]);
                        kl += kl_std_normal(mu.v[j], sig);
                        gm = fmaf(cfac, mu.v[j], gm);
                        gr = fmaf(link_grad<LINK>(rho.v[j]) * cfac, sig - fast_rcp(sig), gr);
                    }
                    adam_elem(mu.v[j], m1.v[j], v1.v[j], gm, h, step_size, inv_bc2);
                    adam_elem(rho.v[j], m2.v[j], v2.v[j], gr, h, step_size, inv_bc2);
                }
                st_vec<VEC>(entity + eoff + k, mu);        st_vec<VEC>(entity + eoff + d + k, rho);
                st_vec<VEC>(entity_m + eoff + k, m1);      st_vec<VEC>(entity_m + eoff + d + k, m2);
                st_vec<VEC>(entity_v + eoff + k, v1);      st_vec<VEC>(entity_v + eoff + d + k, v2);
            }
        }
        kl = group_sum<LPR>(kl, gmask);
        if (gl == 0) {
            float2 ab = *reinterpret_cast<const float2*>(bias + boff);
            float2 bm = *reinterpret_cast<const float2*>(bias_m + boff);
            float2 bv = *reinterpret_cast<const float2*>(bias_v + boff);
            const float2 gb2 = *reinterpret_cast<const float2*>(g_bias + boff);
            float ga = gb2.x, gb = gb2.y;
            if (cnt > 0.f) {
                const float tau = link_fn<LINK>(ab.y);
                kl += kl_std_normal(ab.x, tau);
                ga = fmaf(cfac, ab.x, ga);
                gb = fmaf(link_grad<LINK>(ab.y) * cfac, tau - fast_rcp(tau), gb);
                facc = fmaf(cfac, kl, facc);
            }
            adam_elem(ab.x, bm.x, bv.x, ga, h, step_size, inv_bc2);
            adam_elem(ab.y, bm.y, bv.y, gb, h, step_size, inv_bc2);
            *reinterpret_cast<float2*>(bias + boff) = ab;
            *reinterpret_cast<float2*>(bias_m + boff) = bm;
            *reinterpret_cast<float2*>(bias_v + boff) = bv;
        }
    }
    double acc[1] = {(double)facc};
    if (block_partials<1>(acc, partials, counter)) {
        double tot[1];
        final_sums<1>(partials, tot);
        if (threadIdx.x == 0) { stats[VFMB_ST_KL_ROWS] = (float)tot[0]; *counter = 0; }
    }
}

// scalar parameters from the all-reduced sums; also the global-batch loss
template <int LINK, int LIK>
__global__ void k_dp_final(DevCfg c, float* __restrict__ scalars, float* __restrict__ sm, float* __restrict__ sv,
                           const float* __restrict__ tail, const float* __restrict__ eps_global, AdamDev h,
                           int32_t* __restrict__ adam_step, const int32_t* __restrict__ noise_step,
                           float* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t step = (uint32_t)adam_step[0];
    const uint32_t nstep = noise_step ? (uint32_t)noise_step[1] : step;    // noise index of the forward
    dp_scalar_update<LINK>(LIK, c.S, c.B, c.n_train, c.seed, scalars, sm, sv, tail, eps_global, h, step, nstep, stats);
    if (tail[VFMB_DP_T_OVERFLOW] > 0.f) stats[VFMB_ST_LOSS] = __int_as_float(0x7fc00000);   // a request bucket overflowed
    adam_step[0] = (int32_t)step + 1;
}

}  // namespace vfmb

using namespace vfmb;

extern "C" int vfmb_dp_scatter_counts(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io,
                                      float* counts, float* tail, vfmb_stream stream_) {
    if (!cfg || !plan || !io || !counts || !tail) return set_error(VFMB_EINVAL, "vfmb_dp_scatter_counts: null argument");
    int64_t n = (int64_t)cfg->B * cfg->F;
    int64_t u_cap = n < cfg->R ? n : cfg->R;
    int g = (int)((u_cap + 255) / 256);
    if (g > 4 * kNumSMs) g = 4 * kNumSMs;
    k_dp_scatter_counts<<<g, 256, 0, counted((cudaStream_t)stream_)>>>(plan->urec, plan->meta, plan->z, io->stats, cfg->B,
                                                               cfg->F, counts, tail);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_dp_apply_sampled(const vfmb_config* cfg, const vfmb_tables* tab, const float* grad_entity,
                                     const float* grad_bias, const float* counts, const float* tail,
                                     const float* eps_global, const vfmb_adam* adam, int32_t dense_adam,
                                     double* partials, int32_t* counters, float* stats, vfmb_stream stream_) {
    if (!cfg || !tab || !grad_entity || !grad_bias || !counts || !tail || !adam || !partials || !counters || !stats)
        return set_error(VFMB_EINVAL, "vfmb_dp_apply_sampled: null argument");
    if (cfg->S != 1) return set_error(VFMB_ESHAPE, "vfmb_dp_apply_sampled: S=1 only");
    cudaStream_t stream = (cudaStream_t)stream_;
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    DevCfg dc = make_dev(cfg);
    AdamDev h = make_adam(adam);
    const int gpb = 8 * (32 / L.lpr);
    int64_t g = ((int64_t)cfg->R + gpb - 1) / gpb;
    if (g > kMaxGrid) g = kMaxGrid;
    if (g < 1) g = 1;
#define LAUNCH_APPLY(LINK)                                                                              \
    k_dp_apply<VEC, LPR, NV, LINK><<<(int)g, 256, 0, counted(stream)>>>(                                         \
        dc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v,             \
        tab->train_counts, grad_entity, grad_bias, counts, tail, cfg->R, h, tab->adam_step, dense_adam, \
        partials, counters + 3, stats)
    VFMB_LAYOUT_SWITCH(L, { if (cfg->link == VFMB_LINK_ABS) LAUNCH_APPLY(0); else LAUNCH_APPLY(1); });
#undef LAUNCH_APPLY
    CUDA_TRY(cudaGetLastError());
    return vfmb_dp_final(cfg, tab, tail, eps_global, adam, stats, stream_);
}

extern "C" int vfmb_dp_final(const vfmb_config* cfg, const vfmb_tables* tab, const float* tail,
                             const float* eps_global, const vfmb_adam* adam, float* stats, vfmb_stream stream_) {
    if (!cfg || !tab || !tail || !adam || !stats || !tab->scalars || !tab->scalars_m || !tab->scalars_v || !tab->adam_step)
        return set_error(VFMB_EINVAL, "vfmb_dp_final: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    DevCfg dc = make_dev(cfg);
    AdamDev h = make_adam(adam);
#define LAUNCH_DPF(LINK, LIK)                                                                           \
    k_dp_final<LINK, LIK><<<1, 32, 0, counted(stream)>>>(dc, tab->scalars, tab->scalars_m, tab->scalars_v, tail,  \
                                                eps_global, h, tab->adam_step, tab->noise_step, stats)
    switch (cfg->link * 2 + cfg->likelihood) {
        case 0: LAUNCH_DPF(0, 0); break;
        case 1: LAUNCH_DPF(0, 1); break;
        case 2: LAUNCH_DPF(1, 0); break;
        default: LAUNCH_DPF(1, 1); break;
    }
#undef LAUNCH_DPF
    CUDA_TRY(cudaGetLastError());
    return 0;
}
