// Sampled-ELBO VFM step (vfm-torch.py:189-324 forward, :359 loss, :368-370 backward + Adam).
//
// Five launches per fused training step (S = 1), all HBM/L2-bound (no dense contraction on this path):
//   k_stage      one lane group per UNIQUE row: gather [mean|raw scale], draw eps (Philox or
//                injected), write the sampled row v = mu + eps*|rho| and bias w to an L2-resident
//                scratch (+ the count-rescaled KL when not fused).   (vfm-torch.py:207-241, 290-317)
//   k_score      one lane group per SAMPLE: FM interaction of the sampled rows, likelihood,
//                residual dloss/dpred.                                (vfm-torch.py:244-270, 359)
//   k_gather     ordered segmented sum of residual * partner row over the sorted occurrence list,
//                tiled by position; k_combine_cut finishes the rows cut by tile boundaries.
//   k_gather_score  F = 2, step running alone: k_score + k_gather in one pass (4 launches).
//   k_adam_rows  per unique row: chain rule to (mu, rho) + KL gradient + Adam; FLAVOR 2 also forms
//                the KL and recomputes the Philox draws; the last block updates the scalar
//                parameters (alpha, global bias) and the step counter.
//                Deterministic: fixed summation order, no floating-point atomics.    (:368-370)
// S > 1 variational samples: k_stage / k_gather once per sample, k_score_multi, k_adam_rows_multi.
// Host side: launch_* (internal) and the extern "C" entry points at the end of the file.
#include "step_common.cuh"

#include <cstdlib>

namespace vfmb {

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

// ------------------------------------------------------------------------------- k_stage
// LEAN = 1 (fused training step): no KL sum and no copy of the noise -- k_adam_rows, which holds
// the row anyway and has idle issue slots, recomputes both (same Philox counters => same bits).
template <int VEC, int LPR, int NV, int LINK, int LEAN>
__global__ void __launch_bounds__(256)
k_stage(DevCfg c, const float* __restrict__ bias, const float* __restrict__ entity,
        const float* __restrict__ train_counts, const int32_t* __restrict__ urec,
        const int32_t* __restrict__ meta, const float* __restrict__ z,
        const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
        const int32_t* __restrict__ adam_step, float* __restrict__ vs, float* __restrict__ ws,
        float* __restrict__ es, float* __restrict__ ebs, float* __restrict__ cq,
        double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats,
        int smp, int u_stride) {
    // smp: variational sample this launch draws (S > 1: one launch per sample, outputs [S][u_stride]);
    // the KL and the KL weights do not depend on the sample and are formed by sample 0
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW;
    const int U = meta[0];
    const int d = c.d;
    vs += (size_t)smp * u_stride * d; ws += (size_t)smp * u_stride;
    if (es) es += (size_t)smp * u_stride * d;
    if (ebs) ebs += (size_t)smp * u_stride;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const unsigned gmask = group_mask<LPR>();
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    float facc = 0.f;                                   // sum_u c_u * KL_u over this thread's rows

    for (int base = gwarp * CH; base < U; base += nwarps * CH) {
        // ---- lane-parallel: one unique row per lane (first CH lanes)
        const int ul = base + lane;
        const bool valid = lane < CH && ul < U;
        int rowid_l = 0;
        float klb = 0.f, cqv = 0.f;
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x;
            prefetch_row(entity + (size_t)rowid_l * 2 * d, 8 * d);
            const float2 ab = *reinterpret_cast<const float2*>(bias + (size_t)rowid_l * 2);
            const float tcnt = __ldg(train_counts + rowid_l);
            const int gid_l = rowid_l * c.row_stride + c.row_offset;     // global id (row-sharded tables)
            const float eb = bias_eps(eps_bias, c, ul, gid_l, step, smp, U);
            if (!eps_bias) ebs[ul] = eb;
            const float tau = link_fn<LINK>(ab.y);
            ws[ul] = fmaf(eb, tau, ab.x);
            if (!LEAN) klb = kl_std_normal(ab.x, tau);
            const int cls = class_of(c, gid_l);
            float csz = 0.f, zc = 1.f;
#pragma unroll
            for (int i = 0; i < kMaxFields; ++i) if (i == cls) { csz = c.class_size[i]; zc = __ldg(z + i); }
            cqv = ((float)rec.w / tcnt) * (csz / zc);        // c_u of SURVEY 8-Maths (rec.w: batch count)
            cq[ul] = cqv;
        }
        float klrow = 0.f;
        // ---- wide work: GPW rows per round, LPR lanes per row (rows were prefetched into L2 above)
#pragma unroll 1
        for (int it = 0; it < kRounds; ++it) {
            const int sel = it * GPW + gidx;
            const int rowid = bcast(rowid_l, sel);
            const int u = base + sel;
            float kl = 0.f;
            if (u < U) {
                const float* erow = entity + (size_t)rowid * 2 * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        const Vec<VEC> mu = ld_vec<VEC>(erow + k), rho = ld_vec<VEC>(erow + d + k);
                        Vec<VEC> e = entity_eps<VEC>(eps_entity, c, u, rowid * c.row_stride + c.row_offset, k, step, smp, U), out;
                        if (!LEAN && !eps_entity) st_vec<VEC>(es + (size_t)u * d + k, e);
                        if (LEAN) {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) out.v[j] = fmaf(e.v[j], link_fn<LINK>(rho.v[j]), mu.v[j]);
                            st_vec<VEC>(vs + (size_t)u * d + k, out);
                            continue;
                        }
                        // sum_k KL(N(mu,sig)||N(0,1)) = 0.5 (sum sig^2 + mu^2 - 1) - 0.5 log prod sig^2:
                        // one logarithm per lane instead of one per element
                        float quad = 0.f, prodv = 1.f;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float sig = link_fn<LINK>(rho.v[j]);
                            out.v[j] = fmaf(e.v[j], sig, mu.v[j]);
                            const float vr = sig * sig;
                            quad += vr + mu.v[j] * mu.v[j] - 1.f;
                            prodv *= vr;
                        }
                        float lg = __logf(prodv);
                        if (!(prodv > 1e-30f && prodv < 1e30f)) {       // tiny / huge scales: no product trick
                            lg = 0.f;
#pragma unroll
                            for (int j = 0; j < VEC; ++j) { const float sg = link_fn<LINK>(rho.v[j]); lg += logf(sg * sg); }
                        }
                        kl += 0.5f * (quad - lg);
                        st_vec<VEC>(vs + (size_t)u * d + k, out);
                    }
                }
                if (!LEAN) kl = group_sum<LPR>(kl, gmask);
            }
            if (!LEAN) hand_back<LPR>(klrow, kl, it, lane);
        }
        if (valid) facc = fmaf(cqv, klrow + klb, facc);
    }
    if (LEAN || smp > 0) return;
    double acc[1] = {(double)facc};
    if (block_partials<1>(acc, partials, counter)) {
        double tot[1];
        final_sums<1>(partials, tot);
        if (threadIdx.x == 0) {
            stats[VFMB_ST_KL_ROWS] = (float)tot[0];
            stats[VFMB_ST_U] = (float)U;
            *counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- k_score
template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_score(DevCfg c, const float* __restrict__ scalars, const int32_t* __restrict__ inverse,
        const int32_t* __restrict__ pos_of, const float* __restrict__ vs, const float* __restrict__ ws,
        const float* __restrict__ y, const float* __restrict__ eps_global,
        const int32_t* __restrict__ adam_step, float* __restrict__ pred, float* __restrict__ mean,
        float* __restrict__ resid, float* __restrict__ rsorted, float* __restrict__ msg,
        double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats,
        int defer_kl) {
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW;
    const int d = c.d, F = c.F, B = c.B;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const unsigned gmask = group_mask<LPR>();
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float mu0 = scalars[VFMB_S_GB_MEAN];
    const float sig0 = link_fn<LINK>(scalars[VFMB_S_GB_SCALE]);
    const float w0 = mu0 + global_eps(eps_global, c, step) * sig0;
    const float alpha = link_fn<LINK>(scalars[VFMB_S_ALPHA]);
    const float half_log_alpha = 0.5f * logf(alpha);
    const float scale = c.n_train / ((float)c.S * (float)B);
    double acc[3] = {0.0, 0.0, 0.0};     // nll, resid, squared error

    for (int base = gwarp * CH; base < B; base += nwarps * CH) {
        // ---- lane-parallel: one sample per lane (ranks, bias sum, target)
        const int nl = base + lane;
        const bool valid = lane < CH && nl < B;
        int2 rr = make_int2(0, 0);
        float bsum = 0.f, yn = 0.f;
        if (valid) {
            if (F == 2) {
                rr = __ldg(reinterpret_cast<const int2*>(inverse) + nl);
                bsum = __ldg(ws + rr.x) + __ldg(ws + rr.y);
            } else {
                for (int f = 0; f < F; ++f) bsum += __ldg(ws + __ldg(inverse + (size_t)nl * F + f));
            }
            if (y) yn = __ldg(y + nl);
        }
        float inter_l = 0.f;
        // ---- wide work: interaction of GPW samples per round
        if (F == 2) {
#pragma unroll 1
            for (int it = 0; it < kRounds; ++it) {
                const int sel = it * GPW + gidx;
                const int r0 = bcast(rr.x, sel), r1 = bcast(rr.y, sel);
                float part = 0.f;
                if (base + sel < B) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        int k = (gl + i * LPR) * VEC;
                        if (k < d) {
                            const Vec<VEC> a = ld_vec_nc<VEC>(vs + (size_t)r0 * d + k);
                            const Vec<VEC> b = ld_vec_nc<VEC>(vs + (size_t)r1 * d + k);
#pragma unroll
                            for (int j = 0; j < VEC; ++j) part = fmaf(a.v[j], b.v[j], part);
                        }
                    }
                }
                part = group_sum<LPR>(part, gmask);
                hand_back<LPR>(inter_l, part, it, lane);
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < kRounds; ++it) {
                const int sel = it * GPW + gidx;
                const int n = base + sel;
                float part = 0.f;
                if (n < B) {
                    Vec<VEC> ssum[NV], sq[NV];
#pragma unroll
                    for (int i = 0; i < NV; ++i)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) { ssum[i].v[j] = 0.f; sq[i].v[j] = 0.f; }
#pragma unroll 4
                    for (int f = 0; f < F; ++f) {
                        const int r = __ldg(inverse + (size_t)n * F + f);
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            int k = (gl + i * LPR) * VEC;
                            if (k < d) {
                                Vec<VEC> a = ld_vec_nc<VEC>(vs + (size_t)r * d + k);
#pragma unroll
                                for (int j = 0; j < VEC; ++j) { ssum[i].v[j] += a.v[j]; sq[i].v[j] = fmaf(a.v[j], a.v[j], sq[i].v[j]); }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        int k = (gl + i * LPR) * VEC;
                        if (k < d) {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) part += 0.5f * (ssum[i].v[j] * ssum[i].v[j] - sq[i].v[j]);
                            if (msg) st_vec<VEC>(msg + (size_t)n * d + k, ssum[i]);   // S_n = sum_f v_f (unscaled)
                        }
                    }
                }
                part = group_sum<LPR>(part, gmask);
                hand_back<LPR>(inter_l, part, it, lane);
            }
        }
        // ---- lane-parallel: likelihood, residual, outputs (coalesced)
        if (valid) {
            const float p = w0 + bsum + inter_l;
            float r = 0.f, mu_out = p;
            if (LIK == VFMB_BERNOULLI) mu_out = 1.f / (1.f + expf(-p));
            pred[nl] = p;
            mean[nl] = mu_out;
            if (y) {
                float nll, err = yn - p;
                if (LIK == VFMB_GAUSSIAN) {
                    nll = 0.5f * alpha * err * err - half_log_alpha + 0.9189385332046727f;
                    r = scale * alpha * (p - yn);
                } else {
                    nll = fmaxf(p, 0.f) - yn * p + log1pf(expf(-fabsf(p)));
                    r = scale * (mu_out - yn);
                }
                acc[0] += (double)nll; acc[1] += (double)r; acc[2] += (double)err * (double)err;
                resid[nl] = r;
                // residual in sorted-occurrence order: the backward reads it coalesced
                if (F == 2) {
                    const int2 pp = __ldg(reinterpret_cast<const int2*>(pos_of) + nl);
                    rsorted[pp.x] = r; rsorted[pp.y] = r;
                } else {
                    for (int f = 0; f < F; ++f) rsorted[__ldg(pos_of + (size_t)nl * F + f)] = r;
                }
            }
        }
    }
    if (block_partials<3>(acc, partials, counter)) {
        double tot[3];
        final_sums<3>(partials, tot);
        if (threadIdx.x == 0) {
            double nll = tot[0], sr = tot[1], sq = tot[2];
            stats[VFMB_ST_NLL_MEAN] = (float)(nll / (double)B);
            stats[VFMB_ST_SUM_RESID] = (float)sr;
            stats[VFMB_ST_SUM_SQERR] = (float)sq;
            if (defer_kl) {                 // fused step: k_adam_rows adds the KL (data term only here)
                stats[VFMB_ST_LOSS] = (float)((double)c.n_train * nll / (double)B);
            } else {
                float kl0 = kl_std_normal(mu0, sig0);
                float kl = kl0 + stats[VFMB_ST_KL_ROWS];
                stats[VFMB_ST_KL] = kl;
                stats[VFMB_ST_LOSS] = (float)((double)c.n_train * nll / (double)B + (double)kl);
            }
            stats[VFMB_ST_W0] = w0;
            *counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- k_score_multi
// S > 1 variational samples (N_VARIATIONAL_SAMPLES, vfm-torch.py:238-245, 264-270): the bias and FM
// terms are averaged over the samples BEFORE the likelihood, the global bias is not --
//   pred[s, n] = w0_s + mean_s'(sum_f w_{s'}) + mean_s'(FM_{s'}),   likelihood batch shape [S, B].
// One lane group per sample n walks the S sampled copies of its rows.  The residual handed to the
// backward is rho_n = (1/S) sum_s dloss/dpred[s, n] (every sampled copy of a row sees it, since
// dpred[s', n]/dv_{s,u} = partner_s / S).  Plain structure on purpose: S = 1 is the tuned path.
constexpr int kMaxSamples = 8;
template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_score_multi(DevCfg c, int u_stride, const float* __restrict__ scalars, const int32_t* __restrict__ inverse,
              const int32_t* __restrict__ pos_of, const float* __restrict__ vs, const float* __restrict__ ws,
              const float* __restrict__ y, const float* __restrict__ eps_global,
              const int32_t* __restrict__ adam_step, float* __restrict__ pred, float* __restrict__ mean,
              float* __restrict__ resid, float* __restrict__ rsorted, float* __restrict__ msg,
              double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR;
    const int d = c.d, F = c.F, B = c.B, S = c.S;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    const float mu0 = scalars[VFMB_S_GB_MEAN];
    const float sig0 = link_fn<LINK>(scalars[VFMB_S_GB_SCALE]);
    const float alpha = link_fn<LINK>(scalars[VFMB_S_ALPHA]);
    const float half_log_alpha = 0.5f * logf(alpha);
    const float scale = c.n_train / ((float)S * (float)B);
    const float inv_s = 1.f / (float)S;
    float w0s[kMaxSamples];
#pragma unroll
    for (int q = 0; q < kMaxSamples; ++q) w0s[q] = q < S ? mu0 + global_eps(eps_global, c, step, q) * sig0 : 0.f;
    double acc[2 + kMaxSamples];                          // nll, squared error, residual sum per sample
#pragma unroll
    for (int q = 0; q < 2 + kMaxSamples; ++q) acc[q] = 0.0;

    for (int n = group; n < B; n += ngroups) {
        float bsum = 0.f, fmsum = 0.f;
        for (int q = 0; q < S; ++q) {
            const float* vq = vs + (size_t)q * u_stride * d;
            const float* wq = ws + (size_t)q * u_stride;
            float b = 0.f, part = 0.f;
            if (F == 2) {
                const int2 rr = __ldg(reinterpret_cast<const int2*>(inverse) + n);
                b = __ldg(wq + rr.x) + __ldg(wq + rr.y);
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        const Vec<VEC> a = ld_vec_nc<VEC>(vq + (size_t)rr.x * d + k), bb = ld_vec_nc<VEC>(vq + (size_t)rr.y * d + k);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) part = fmaf(a.v[j], bb.v[j], part);
                    }
                }
            } else {
                Vec<VEC> ssum[NV], sq[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { ssum[i].v[j] = 0.f; sq[i].v[j] = 0.f; }
                for (int f = 0; f < F; ++f) {
                    const int r = __ldg(inverse + (size_t)n * F + f);
                    b += __ldg(wq + r);
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        int k = (gl + i * LPR) * VEC;
                        if (k < d) {
                            const Vec<VEC> a = ld_vec_nc<VEC>(vq + (size_t)r * d + k);
#pragma unroll
                            for (int j = 0; j < VEC; ++j) { ssum[i].v[j] += a.v[j]; sq[i].v[j] = fmaf(a.v[j], a.v[j], sq[i].v[j]); }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) part += 0.5f * (ssum[i].v[j] * ssum[i].v[j] - sq[i].v[j]);
                        if (msg) st_vec<VEC>(msg + ((size_t)q * B + n) * d + k, ssum[i]);   // S_n of sample q
                    }
                }
            }
            part = group_sum<LPR>(part, gmask);
            bsum += b; fmsum += part;
        }
        if (gl == 0) {
            const float mb = bsum * inv_s, mf = fmsum * inv_s;
            const float yn = y ? __ldg(y + n) : 0.f;
            float rsum = 0.f;
            for (int q = 0; q < S; ++q) {
                const float p = w0s[q] + mb + mf;
                float mu_out = p;
                if (LIK == VFMB_BERNOULLI) mu_out = 1.f / (1.f + expf(-p));
                pred[(size_t)q * B + n] = p;
                mean[(size_t)q * B + n] = mu_out;
                if (y) {
                    float nll, r;
                    const float err = yn - p;
                    if (LIK == VFMB_GAUSSIAN) {
                        nll = 0.5f * alpha * err * err - half_log_alpha + 0.9189385332046727f;
                        r = scale * alpha * (p - yn);
                    } else {
                        nll = fmaxf(p, 0.f) - yn * p + log1pf(expf(-fabsf(p)));
                        r = scale * (mu_out - yn);
                    }
                    acc[0] += (double)nll; acc[1] += (double)err * (double)err; acc[2 + q] += (double)r;
                    rsum += r;
                }
            }
            if (y) {
                const float rho = rsum * inv_s;
                resid[n] = rho;
                for (int f = 0; f < F; ++f) rsorted[__ldg(pos_of + (size_t)n * F + f)] = rho;
            }
        }
    }
    if (block_partials<2 + kMaxSamples>(acc, partials, counter)) {
        double tot[2 + kMaxSamples];
        final_sums<2 + kMaxSamples>(partials, tot);
        if (threadIdx.x == 0) {
            const double cnt = (double)S * (double)B;
            double sr = 0.0;
            for (int q = 0; q < S; ++q) { sr += tot[2 + q]; stats[VFMB_ST_RESID_S + q] = (float)tot[2 + q]; stats[VFMB_ST_W0_S + q] = w0s[q]; }
            const float kl = kl_std_normal(mu0, sig0) + stats[VFMB_ST_KL_ROWS];
            stats[VFMB_ST_NLL_MEAN] = (float)(tot[0] / cnt);
            stats[VFMB_ST_SUM_RESID] = (float)sr;
            stats[VFMB_ST_SUM_SQERR] = (float)tot[1];
            stats[VFMB_ST_KL] = kl;
            stats[VFMB_ST_LOSS] = (float)((double)c.n_train * tot[0] / cnt + (double)kl);
            stats[VFMB_ST_W0] = w0s[0];
            *counter = 0;
        }
    }
}

// residuals supplied by the caller (autograd path): scatter them into sorted-occurrence order
__global__ void __launch_bounds__(256)
k_scatter_resid(const float* __restrict__ resid, const int32_t* __restrict__ pos_of, int N, int F,
                float* __restrict__ rsorted) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < N; o += gridDim.x * blockDim.x)
        rsorted[pos_of[o]] = resid[o / F];
}

// ------------------------------------------------------------------------------- k_gather
// Backward, phase A: deterministic segmented reduction over the SORTED occurrence list, tiled by
// position so that every lane group does the same amount of work whatever the row popularity
// (a Zipf head row with thousands of occurrences is just many tiles).  A group walks the kTile
// positions of its tile in order, accumulating r_n * partner row; when the unique rank changes it
// flushes.  Rows that lie inside the tile are final (-> grow/gws).  A row cut by a tile boundary
// leaves a partial in the tile's head slot (row continues from the previous tile) or tail slot
// (row starts here and continues); k_adam_rows adds a row's partials in tile order.  Fixed
// summation order, no atomics => bitwise reproducible.
// Everything read here is L2-resident scratch written by k_stage / k_score.
template <int VEC, int LPR, int NV, int UNIT>
__global__ void __launch_bounds__(256)
k_gather(int d, int F, int N, const int32_t* __restrict__ partner, const int32_t* __restrict__ pos_rank,
         const float* __restrict__ vs, const float* __restrict__ msg, const float* __restrict__ rsorted,
         float* __restrict__ gslot, float* __restrict__ grow, float* __restrict__ gws) {
    constexpr int GPW = kWarp / LPR;
    const int dp = d + 4;                               // slot pitch (keeps 16 B alignment)
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    const int n_tiles = (N + kTile - 1) / kTile;
    const float* table = (F == 2) ? vs : msg;

    for (int tile = group; tile < n_tiles; tile += ngroups) {
        const int t0 = tile * kTile, t1 = min(N, t0 + kTile);
        // does the first row continue from the previous tile / the last row into the next one?
        const bool head_open = t0 > 0 && __ldg(pos_rank + t0 - 1) == __ldg(pos_rank + t0);
        const bool tail_open = t1 < N && __ldg(pos_rank + t1) == __ldg(pos_rank + t1 - 1);
        const int first_u = __ldg(pos_rank + t0), last_u = __ldg(pos_rank + t1 - 1);
        int cur = first_u;
        Vec<VEC> acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[i].v[j] = 0.f;
        float gw = 0.f;

        auto flush = [&](int u) {
            const bool open_h = head_open && u == first_u, open_t = tail_open && u == last_u;
            float* dst; float* dstw;
            if (open_h)      { dst = gslot + ((size_t)tile * 2) * dp;     dstw = dst + d; }
            else if (open_t) { dst = gslot + ((size_t)tile * 2 + 1) * dp; dstw = dst + d; }
            else             { dst = grow + (size_t)u * d;                dstw = gws + u; }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) {
                    if (F > 2 && !open_h && !open_t) {  // pairwise, complete row: sum r_n (S_n - v_u)
                        Vec<VEC> own = ld_vec_nc<VEC>(vs + (size_t)u * d + k);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) acc[i].v[j] = fmaf(-gw, own.v[j], acc[i].v[j]);
                    }
                    st_vec<VEC>(dst + k, acc[i]);
                }
            }
            if (gl == 0) *dstw = gw;
        };

        for (int b0 = t0; b0 < t1; b0 += LPR) {
            const int idx = b0 + gl;
            const bool ok = idx < t1;
            const float r = ok ? __ldg(rsorted + idx) : 0.f;
            const int src = ok ? __ldg(partner + idx) : 0;
            const int ur = ok ? __ldg(pos_rank + idx) : 0;
            const int cnt = min(LPR, t1 - b0);
            const int src0 = __shfl_sync(gmask, src, 0, LPR);
            constexpr int UNR = (NV == 1) ? 8 : 4;
            for (int j = 0; j < cnt; j += UNR) {               // UNR row gathers in flight
                float rj[UNR]; int uj[UNR]; Vec<VEC> t[UNR][NV];
#pragma unroll
                for (int e = 0; e < UNR; ++e) {
                    rj[e] = __shfl_sync(gmask, r, (j + e) & (LPR - 1), LPR);
                    uj[e] = __shfl_sync(gmask, ur, (j + e) & (LPR - 1), LPR);
                    int sj = __shfl_sync(gmask, src, (j + e) & (LPR - 1), LPR);
                    if (j + e >= cnt) sj = src0;
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        int k = (gl + i * LPR) * VEC;
                        if (k < d) t[e][i] = ld_vec_nc<VEC>(table + (size_t)sj * d + k);
                    }
                }
#pragma unroll
                for (int e = 0; e < UNR; ++e) {
                    if (j + e < cnt) {                         // group-uniform
                        if (uj[e] != cur) {
                            flush(cur);
                            cur = uj[e];
#pragma unroll
                            for (int i = 0; i < NV; ++i)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) acc[i].v[q] = 0.f;
                            gw = 0.f;
                        }
                        gw += rj[e];
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            int k = (gl + i * LPR) * VEC;
                            if (k < d)
#pragma unroll
                                for (int q = 0; q < VEC; ++q)
                                    acc[i].v[q] = UNIT ? acc[i].v[q] + t[e][i].v[q] : fmaf(rj[e], t[e][i].v[q], acc[i].v[q]);
                        }
                    }
                }
            }
        }
        flush(cur);
    }
}

// ------------------------------------------------------------------------------- k_gather_score
// F == 2 fused step: k_score and k_gather in one pass over the SORTED occurrence list.  An
// occurrence (row u, sample n) needs the residual r_n = dloss/dpred_n and the partner row; the
// partner row is being loaded anyway, so the group forms the score <v_u, v_partner> itself --
// every sample is scored twice, once from each of its rows, with bit-identical results (products
// and bias sums commute, the lane mapping and the shuffle tree are those of k_score).  The
// occurrence of field 0 writes the sample's outputs and its likelihood terms.  Saves a launch, the
// residual round trip through memory and one of the two passes over the sampled rows.
template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_gather_score(DevCfg c, const float* __restrict__ scalars, const int32_t* __restrict__ partner,
               const int32_t* __restrict__ pos_rank, const int32_t* __restrict__ occ,
               const float* __restrict__ vs, const float* __restrict__ ws, const float* __restrict__ y,
               const float* __restrict__ eps_global, const int32_t* __restrict__ adam_step,
               float* __restrict__ pred, float* __restrict__ mean, float* __restrict__ resid,
               float* __restrict__ gslot, float* __restrict__ grow, float* __restrict__ gws,
               double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR, UNR = 4;
    const int d = c.d, B = c.B, N = 2 * c.B;
    const int dp = d + 4;                               // slot pitch (keeps 16 B alignment)
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    const int n_tiles = (N + kTile - 1) / kTile;
    const float mu0 = scalars[VFMB_S_GB_MEAN];
    const float sig0 = link_fn<LINK>(scalars[VFMB_S_GB_SCALE]);
    const float w0 = mu0 + global_eps(eps_global, c, step) * sig0;
    const float alpha = link_fn<LINK>(scalars[VFMB_S_ALPHA]);
    const float half_log_alpha = 0.5f * logf(alpha);
    const float scale = c.n_train / ((float)c.S * (float)B);
    double acc3[3] = {0.0, 0.0, 0.0};                   // nll, resid, squared error (field-0 occurrences)

    for (int tile = group; tile < n_tiles; tile += ngroups) {
        const int t0 = tile * kTile, t1 = min(N, t0 + kTile);
        const bool head_open = t0 > 0 && __ldg(pos_rank + t0 - 1) == __ldg(pos_rank + t0);
        const bool tail_open = t1 < N && __ldg(pos_rank + t1) == __ldg(pos_rank + t1 - 1);
        const int first_u = __ldg(pos_rank + t0), last_u = __ldg(pos_rank + t1 - 1);
        int cur = first_u;
        Vec<VEC> acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[i].v[j] = 0.f;
        float gw = 0.f;

        auto flush = [&](int u) {
            const bool open_h = head_open && u == first_u, open_t = tail_open && u == last_u;
            float* dst; float* dstw;
            if (open_h)      { dst = gslot + ((size_t)tile * 2) * dp;     dstw = dst + d; }
            else if (open_t) { dst = gslot + ((size_t)tile * 2 + 1) * dp; dstw = dst + d; }
            else             { dst = grow + (size_t)u * d;                dstw = gws + u; }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) st_vec<VEC>(dst + k, acc[i]);
            }
            if (gl == 0) *dstw = gw;
        };

        for (int b0 = t0; b0 < t1; b0 += LPR) {
            // ---- lane-parallel: one position per lane of the group
            const int idx = b0 + gl;
            const bool ok = idx < t1;
            const int src = ok ? __ldg(partner + idx) : 0;
            const int ur = ok ? __ldg(pos_rank + idx) : 0;
            const int o = ok ? __ldg(occ + idx) : 1;
            // bias sum in field order (as k_score: ws[rank of field 0] + ws[rank of field 1])
            const float bs = ok ? ((o & 1) ? __ldg(ws + src) + __ldg(ws + ur) : __ldg(ws + ur) + __ldg(ws + src)) : 0.f;
            const float yl = ok ? __ldg(y + (o >> 1)) : 0.f;
            const int cnt = min(LPR, t1 - b0);
            const int src0 = __shfl_sync(gmask, src, 0, LPR), ur0 = __shfl_sync(gmask, ur, 0, LPR);
            for (int j = 0; j < cnt; j += UNR) {               // UNR row pairs in flight
                int uj[UNR], sj[UNR];
                Vec<VEC> t[UNR][NV], own[UNR][NV];
#pragma unroll
                for (int e = 0; e < UNR; ++e) {
                    uj[e] = __shfl_sync(gmask, ur, (j + e) & (LPR - 1), LPR);
                    sj[e] = __shfl_sync(gmask, src, (j + e) & (LPR - 1), LPR);
                    if (j + e >= cnt) { sj[e] = src0; uj[e] = ur0; }
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        int k = (gl + i * LPR) * VEC;
                        if (k < d) {
                            t[e][i] = ld_vec_nc<VEC>(vs + (size_t)sj[e] * d + k);
                            own[e][i] = ld_vec_nc<VEC>(vs + (size_t)uj[e] * d + k);   // L1 hit inside a segment
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < UNR; ++e) {
                    if (j + e < cnt) {                         // group-uniform
                        const int oj = __shfl_sync(gmask, o, (j + e) & (LPR - 1), LPR);
                        const float bsj = __shfl_sync(gmask, bs, (j + e) & (LPR - 1), LPR);
                        const float yn = __shfl_sync(gmask, yl, (j + e) & (LPR - 1), LPR);
                        // field-0 row first, as k_score multiplies them (a * b is commutative; kept for clarity)
                        float part = 0.f;
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            int k = (gl + i * LPR) * VEC;
                            if (k < d)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) part = fmaf(own[e][i].v[q], t[e][i].v[q], part);
                        }
                        part = group_sum<LPR>(part, gmask);
                        const float p = w0 + bsj + part;
                        float r, mu_out = p;
                        if (LIK == VFMB_BERNOULLI) { mu_out = 1.f / (1.f + expf(-p)); r = scale * (mu_out - yn); }
                        else r = scale * alpha * (p - yn);
                        if (!(oj & 1) && gl == 0) {            // the field-0 occurrence owns the sample's outputs
                            const int n = oj >> 1;
                            const float err = yn - p;
                            float nll;
                            if (LIK == VFMB_GAUSSIAN) nll = 0.5f * alpha * err * err - half_log_alpha + 0.9189385332046727f;
                            else nll = fmaxf(p, 0.f) - yn * p + log1pf(expf(-fabsf(p)));
                            pred[n] = p; mean[n] = mu_out; resid[n] = r;
                            acc3[0] += (double)nll; acc3[1] += (double)r; acc3[2] += (double)err * (double)err;
                        }
                        if (uj[e] != cur) {
                            flush(cur);
                            cur = uj[e];
#pragma unroll
                            for (int i = 0; i < NV; ++i)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) acc[i].v[q] = 0.f;
                            gw = 0.f;
                        }
                        gw += r;
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            int k = (gl + i * LPR) * VEC;
                            if (k < d)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) acc[i].v[q] = fmaf(r, t[e][i].v[q], acc[i].v[q]);
                        }
                    }
                }
            }
        }
        flush(cur);
    }
    if (block_partials<3>(acc3, partials, counter)) {
        double tot[3];
        final_sums<3>(partials, tot);
        if (threadIdx.x == 0) {     // the KL is added by k_adam_rows<FLAVOR 2> (data term only here)
            stats[VFMB_ST_NLL_MEAN] = (float)(tot[0] / (double)B);
            stats[VFMB_ST_SUM_RESID] = (float)tot[1];
            stats[VFMB_ST_SUM_SQERR] = (float)tot[2];
            stats[VFMB_ST_LOSS] = (float)((double)c.n_train * tot[0] / (double)B);
            stats[VFMB_ST_W0] = w0;
            *counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- k_adam_rows
// Backward, phase B (the HBM-bound kernel of the step): per unique row, chain rule from
// (g_v, g_w) to (mean, raw scale) + KL gradient, then Adam on the row -- parameters and both
// moments of every touched row are read and written exactly once.
//
// Row gradients are final in grow/gws (k_combine_cut finished the rows cut by tile boundaries).
// FLAVOR 0  plain.
// FLAVOR 1  + the block that finishes last updates the scalar parameters and the step counter
//             (what k_final did as a separate launch).
// FLAVOR 2  + the count-rescaled KL of the rows (the row is in registers, the kernel is DRAM-bound
//             with idle issue slots) and, without injected noise, the Philox draws recomputed
//             instead of read back -- k_stage<LEAN> wrote neither.
struct FinalArgs {
    float* scalars; float* sm; float* sv; float* stats; const float* eps_global;
    float* grad_scalars; double* partials; int32_t* counter; const float* gslot;
    int likelihood;
};

template <int LINK, int MODE>
__device__ __forceinline__ void final_scalars(const DevCfg& c, const FinalArgs& fa, const AdamDev& h,
                                              int32_t* __restrict__ adam_step, float kl_scale,
                                              bool with_kl, double kl_rows, int U) {
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    float* scalars = fa.scalars; float* stats = fa.stats;
    float alpha = scalars[VFMB_S_ALPHA], mu0 = scalars[VFMB_S_GB_MEAN], rho0 = scalars[VFMB_S_GB_SCALE];
    const float sig0 = link_fn<LINK>(rho0), ap = link_fn<LINK>(alpha);
    const double sr = (double)stats[VFMB_ST_SUM_RESID], sq = (double)stats[VFMB_ST_SUM_SQERR];
    // sum_s eps0_s * (sum_n dloss/dpred[s, n]); one sample: eps0 * sr
    double e0sr = 0.0;
    if (c.S > 1) {
        for (int q = 0; q < c.S; ++q) e0sr += (double)global_eps(fa.eps_global, c, step, q) * (double)stats[VFMB_ST_RESID_S + q];
    } else {
        e0sr = (double)global_eps(fa.eps_global, c, step) * sr;
    }
    if (with_kl) {                                      // loss terms of the pre-update parameters
        const float kl = kl_std_normal(mu0, sig0) + (float)kl_rows;
        stats[VFMB_ST_KL_ROWS] = (float)kl_rows;
        stats[VFMB_ST_KL] = kl;
        stats[VFMB_ST_LOSS] = (float)((double)stats[VFMB_ST_LOSS] + (double)kl);
        stats[VFMB_ST_U] = (float)U;
    }
    float g_mu0 = (float)(sr + (double)(kl_scale * mu0));
    float g_rho0 = link_grad<LINK>(rho0) * (float)(e0sr + (double)(kl_scale * (sig0 - 1.f / sig0)));
    float g_alpha = 0.f;
    if (fa.likelihood == VFMB_GAUSSIAN) {
        double sc = (double)c.n_train / ((double)c.S * (double)c.B);
        g_alpha = link_grad<LINK>(alpha) * (float)(sc * (0.5 * sq - 0.5 * (double)c.S * (double)c.B / (double)ap));
    }
    if (MODE == VFMB_ADAM_TOUCHED) {
        float ss, b2;
        adam_coeffs(h, (int)step + 1, &ss, &b2);
        adam_elem(mu0, fa.sm[VFMB_S_GB_MEAN], fa.sv[VFMB_S_GB_MEAN], g_mu0, h, ss, b2);
        adam_elem(rho0, fa.sm[VFMB_S_GB_SCALE], fa.sv[VFMB_S_GB_SCALE], g_rho0, h, ss, b2);
        scalars[VFMB_S_GB_MEAN] = mu0; scalars[VFMB_S_GB_SCALE] = rho0;
        if (fa.likelihood == VFMB_GAUSSIAN) {  // Bernoulli: alpha has no gradient, Adam skips it (N10)
            adam_elem(alpha, fa.sm[VFMB_S_ALPHA], fa.sv[VFMB_S_ALPHA], g_alpha, h, ss, b2);
            scalars[VFMB_S_ALPHA] = alpha;
        }
        adam_step[0] = (int32_t)step + 1;
    } else if (fa.grad_scalars) {
        fa.grad_scalars[VFMB_S_ALPHA] = g_alpha;
        fa.grad_scalars[VFMB_S_GB_MEAN] = g_mu0;
        fa.grad_scalars[VFMB_S_GB_SCALE] = g_rho0;
    }
}

#ifndef VFMB_ADAM_MINB
#define VFMB_ADAM_MINB 4          // resident blocks/SM of the fused flavours (64 registers)
#endif
// bias row of one unique row: chain rule + KL gradient + Adam (or the dense-gradient store).
// klw: in c_u, out c_u * KL(N(a, tau) || N(0,1)) of the pre-update row (when KLF)
template <int LINK, int MODE, bool KLF>
__device__ __forceinline__ void bias_update(float* __restrict__ bias, float* __restrict__ bias_m,
                                            float* __restrict__ bias_v, float* __restrict__ grad_bias,
                                            int rowid, float gw, float eb, float cfac, const AdamDev& h,
                                            float step_size, float inv_bc2, float& klw) {
    const size_t boff = (size_t)rowid * 2;
    float2 ab = *reinterpret_cast<const float2*>(bias + boff);
    const float tau = link_fn<LINK>(ab.y);
    if (KLF) klw *= kl_std_normal(ab.x, tau);
    const float ga = fmaf(cfac, ab.x, gw);
    const float gb = link_grad<LINK>(ab.y) * fmaf(gw, eb, cfac * (tau - fast_rcp(tau)));
    if (MODE == VFMB_ADAM_TOUCHED) {
        float2 bm = *reinterpret_cast<const float2*>(bias_m + boff);
        float2 bv = *reinterpret_cast<const float2*>(bias_v + boff);
        adam_elem(ab.x, bm.x, bv.x, ga, h, step_size, inv_bc2);
        adam_elem(ab.y, bm.y, bv.y, gb, h, step_size, inv_bc2);
        *reinterpret_cast<float2*>(bias + boff) = ab;
        *reinterpret_cast<float2*>(bias_m + boff) = bm;
        *reinterpret_cast<float2*>(bias_v + boff) = bv;
    } else {
        *reinterpret_cast<float2*>(grad_bias + boff) = make_float2(ga, gb);
    }
}

template <int VEC, int LPR, int NV, int LINK, int MODE, int FLAVOR>
__global__ void __launch_bounds__(256, NV > 1 ? 2 : (FLAVOR == 0 ? 4 : VFMB_ADAM_MINB))
k_adam_rows(DevCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
            float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
            const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
            const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
            const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
            AdamDev h, int32_t* __restrict__ adam_step, float kl_scale,
            float* __restrict__ grad_bias, float* __restrict__ grad_entity, FinalArgs fa) {
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW;
    // (summing the rows cut by tile boundaries in here was tried: +16 registers = one resident block
    // per SM less, which cost more than the separate k_combine_cut launch)
    constexpr bool KLF = FLAVOR == 2;
    const int U = meta[0];
    const int d = c.d;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const unsigned gmask = group_mask<LPR>();
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    // bias corrections: fp64 pow per thread (~100 cycles per warp on the FP64 pipe) -- cheaper than
    // a block barrier in front of the first loads
    float step_size = 0.f, inv_bc2 = 1.f;
    if (MODE == VFMB_ADAM_TOUCHED) adam_coeffs(h, (int)step + 1, &step_size, &inv_bc2);
    float facc = 0.f;                                     // sum_u c_u * KL_u over this thread's rows

    for (int base = gwarp * CH; base < U; base += nwarps * CH) {
        // ---- lane-parallel: record, prefetch of the row's parameter / moment lines
        const int ul = base + lane;
        const bool valid = lane < CH && ul < U;
        int rowid_l = 0;
        float cfac_l = 0.f, klw_l = 0.f;                  // KL weight c_u, and c_u * KL(bias) of the lane's row
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x;
            const size_t eoff = (size_t)rowid_l * 2 * d;
            prefetch_row(entity + eoff, 8 * d);
            if (MODE == VFMB_ADAM_TOUCHED) {
                prefetch_row(entity_m + eoff, 8 * d);
                prefetch_row(entity_v + eoff, 8 * d);
            }
            const float cq_l = __ldg(cq + ul);
            cfac_l = kl_scale * cq_l;
            klw_l = cq_l;
            // bias row now: nothing of it stays live across the wide work
            bias_update<LINK, MODE, KLF>(bias, bias_m, bias_v, grad_bias, rowid_l, __ldg(gws + ul),
                                         __ldg(eps_bias + ul), cfac_l, h, step_size, inv_bc2, klw_l);
        }
        float klrow = 0.f;
        // ---- wide work: GPW rows per round
#pragma unroll 1
        for (int it = 0; it < kRounds; ++it) {
            const int sel = it * GPW + gidx;
            const int rowid = bcast(rowid_l, sel);
            const float cfac = bcast(cfac_l, sel);
            const int u = base + sel;
            float kl = 0.f;
            if (u < U) {
                const size_t eoff = (size_t)rowid * 2 * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> e;                         // the noise k_stage used for this row
                        if (KLF) e = entity_eps<VEC>(eps_entity, c, u, rowid * c.row_stride + c.row_offset, k, step);
                        else e = ld_vec_nc<VEC>(eps_entity + (size_t)u * d + k);
                        Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                        Vec<VEC> m1, m2, v1, v2;
                        if (MODE == VFMB_ADAM_TOUCHED) {
                            m1 = ld_vec_cs<VEC>(entity_m + eoff + k); m2 = ld_vec_cs<VEC>(entity_m + eoff + d + k);
                            v1 = ld_vec_cs<VEC>(entity_v + eoff + k); v2 = ld_vec_cs<VEC>(entity_v + eoff + d + k);
                        }
                        const Vec<VEC> g = ld_vec_nc<VEC>(grow + (size_t)u * d + k);
                        Vec<VEC> gmu, grho;
                        float quad = 0.f, prodv = 1.f;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float sig = link_fn<LINK>(rho.v[j]);
                            gmu.v[j] = fmaf(cfac, mu.v[j], g.v[j]);
                            grho.v[j] = link_grad<LINK>(rho.v[j]) * fmaf(g.v[j], e.v[j], cfac * (sig - fast_rcp(sig)));
                            if (KLF) {
                                const float vr = sig * sig;
                                quad += vr + mu.v[j] * mu.v[j] - 1.f;
                                prodv *= vr;
                            }
                        }
                        if (KLF) {       // sum_k KL(N(mu,sig)||N(0,1)), one logarithm per lane (as k_stage)
                            float lg = __logf(prodv);
                            if (!(prodv > 1e-30f && prodv < 1e30f)) {
                                lg = 0.f;
#pragma unroll
                                for (int j = 0; j < VEC; ++j) { const float sg = link_fn<LINK>(rho.v[j]); lg += logf(sg * sg); }
                            }
                            kl += 0.5f * (quad - lg);
                        }
                        if (MODE == VFMB_ADAM_TOUCHED) {
#pragma unroll
                            for (int j = 0; j < VEC; ++j) {
                                adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, inv_bc2);
                                adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, inv_bc2);
                            }
                            st_vec<VEC>(entity + eoff + k, mu);        st_vec<VEC>(entity + eoff + d + k, rho);
                            st_vec_cs<VEC>(entity_m + eoff + k, m1);   st_vec_cs<VEC>(entity_m + eoff + d + k, m2);
                            st_vec_cs<VEC>(entity_v + eoff + k, v1);   st_vec_cs<VEC>(entity_v + eoff + d + k, v2);
                        } else {
                            st_vec<VEC>(grad_entity + eoff + k, gmu);  st_vec<VEC>(grad_entity + eoff + d + k, grho);
                        }
                    }
                }
                if (KLF) kl = group_sum<LPR>(kl, gmask);
            }
            if (KLF) hand_back<LPR>(klrow, kl, it, lane);
        }
        // ---- lane-parallel: KL of the rows -- klw_l is c_u * KL(bias row) (bias_update), klrow the
        // entity part of the row's KL
        if (KLF && valid) facc += fmaf(__ldg(cq + ul), klrow, klw_l);
    }
    if (FLAVOR >= 1) {
        // the block that finishes last owns the scalar parameters: every block has read the step
        // counter / Adam coefficients before it signalled, so updating them here is race-free
        double acc[1] = {(double)facc};
        if (block_partials<1>(acc, fa.partials, fa.counter)) {
            double tot[1] = {0.0};
            if (KLF) final_sums<1>(fa.partials, tot);
            if (threadIdx.x == 0) {
                final_scalars<LINK, MODE>(c, fa, h, adam_step, kl_scale, KLF, tot[0], U);
                *fa.counter = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------- k_adam_rows_multi
// S > 1: the row gradient is the sum over the S sampled copies of the row,
//   d/dmu = sum_s g_s + c_u mu,   d/drho = sign(rho) (sum_s g_s * eps_s + c_u (sigma - 1/sigma)),
//   d/da = S g_w + c_u a,         d/db = sign(b) (g_w sum_s eps^w_s + c_u (tau - 1/tau))
// (g_s = sum_n rho_n partner_s(n): k_gather on sample s; g_w = sum_n rho_n is the same for every s).
// One lane group per unique row, plain structure (S = 1 is the tuned path).  The block that finishes
// last updates the scalar parameters, as in k_adam_rows<FLAVOR 1>.
template <int VEC, int LPR, int NV, int LINK, int MODE>
__global__ void __launch_bounds__(256)
k_adam_rows_multi(DevCfg c, int u_stride, float* __restrict__ bias, float* __restrict__ bias_m,
                  float* __restrict__ bias_v, float* __restrict__ entity, float* __restrict__ entity_m,
                  float* __restrict__ entity_v, const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
                  const float* __restrict__ eps_bias, const float* __restrict__ eps_entity, int eps_stride,
                  const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
                  AdamDev h, int32_t* __restrict__ adam_step, float kl_scale,
                  float* __restrict__ grad_bias, float* __restrict__ grad_entity, FinalArgs fa) {
    constexpr int GPW = kWarp / LPR;
    const int U = meta[0], d = c.d, S = c.S;
    if (eps_stride < 0) eps_stride = U;                    // injected noise: [S, U, ...]; scratch: [S, u_stride, ...]
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    float step_size = 0.f, inv_bc2 = 1.f;
    if (MODE == VFMB_ADAM_TOUCHED) adam_coeffs(h, (int)step + 1, &step_size, &inv_bc2);

    for (int u = group; u < U; u += ngroups) {
        const int rowid = __ldg(urec + 4 * (size_t)u);
        const float cfac = kl_scale * __ldg(cq + u);
        const size_t eoff = (size_t)rowid * 2 * d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                Vec<VEC> gs, ges;
#pragma unroll
                for (int j = 0; j < VEC; ++j) { gs.v[j] = 0.f; ges.v[j] = 0.f; }
                for (int q = 0; q < S; ++q) {
                    const Vec<VEC> g = ld_vec_nc<VEC>(grow + ((size_t)q * u_stride + u) * d + k);
                    const Vec<VEC> e = ld_vec_nc<VEC>(eps_entity + ((size_t)q * eps_stride + u) * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { gs.v[j] += g.v[j]; ges.v[j] = fmaf(g.v[j], e.v[j], ges.v[j]); }
                }
                Vec<VEC> gmu, grho;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    const float sig = link_fn<LINK>(rho.v[j]);
                    gmu.v[j] = fmaf(cfac, mu.v[j], gs.v[j]);
                    grho.v[j] = link_grad<LINK>(rho.v[j]) * (ges.v[j] + cfac * (sig - fast_rcp(sig)));
                }
                if (MODE == VFMB_ADAM_TOUCHED) {
                    Vec<VEC> m1 = ld_vec<VEC>(entity_m + eoff + k), m2 = ld_vec<VEC>(entity_m + eoff + d + k);
                    Vec<VEC> v1 = ld_vec<VEC>(entity_v + eoff + k), v2 = ld_vec<VEC>(entity_v + eoff + d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, inv_bc2);
                        adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, inv_bc2);
                    }
                    st_vec<VEC>(entity + eoff + k, mu);     st_vec<VEC>(entity + eoff + d + k, rho);
                    st_vec<VEC>(entity_m + eoff + k, m1);   st_vec<VEC>(entity_m + eoff + d + k, m2);
                    st_vec<VEC>(entity_v + eoff + k, v1);   st_vec<VEC>(entity_v + eoff + d + k, v2);
                } else {
                    st_vec<VEC>(grad_entity + eoff + k, gmu);  st_vec<VEC>(grad_entity + eoff + d + k, grho);
                }
            }
        }
        if (gl == 0) {                                     // bias row
            const size_t boff = (size_t)rowid * 2;
            float2 ab = *reinterpret_cast<const float2*>(bias + boff);
            const float gw = __ldg(gws + u);
            float ebsum = 0.f;
            for (int q = 0; q < S; ++q) ebsum += __ldg(eps_bias + (size_t)q * eps_stride + u);
            const float tau = link_fn<LINK>(ab.y);
            const float ga = fmaf(cfac, ab.x, (float)S * gw);
            const float gb = link_grad<LINK>(ab.y) * fmaf(gw, ebsum, cfac * (tau - fast_rcp(tau)));
            if (MODE == VFMB_ADAM_TOUCHED) {
                float2 bm = *reinterpret_cast<const float2*>(bias_m + boff);
                float2 bv = *reinterpret_cast<const float2*>(bias_v + boff);
                adam_elem(ab.x, bm.x, bv.x, ga, h, step_size, inv_bc2);
                adam_elem(ab.y, bm.y, bv.y, gb, h, step_size, inv_bc2);
                *reinterpret_cast<float2*>(bias + boff) = ab;
                *reinterpret_cast<float2*>(bias_m + boff) = bm;
                *reinterpret_cast<float2*>(bias_v + boff) = bv;
            } else {
                *reinterpret_cast<float2*>(grad_bias + boff) = make_float2(ga, gb);
            }
        }
    }
    double acc[1] = {0.0};
    if (block_partials<1>(acc, fa.partials, fa.counter)) {
        if (threadIdx.x == 0) {
            final_scalars<LINK, MODE>(c, fa, h, adam_step, kl_scale, false, 0.0, U);
            *fa.counter = 0;
        }
    }
}

#ifdef VFMB_WITH_BULK      // experimental bulk-copy variant (slower on 512-byte rows): -DVFMB_WITH_BULK + VFMB_ADAM_BULK=1
// ------------------------------------------------------------------------------- k_adam_bulk
// The same row update as k_adam_rows<ADAM_TOUCHED, FLAVOR 2>, fed by the bulk-copy engine instead
// of register-staged loads.  k_adam_rows is latency-bound (80 % of the issue slots idle waiting on
// L1TEX scoreboards, DRAM at 55 %): the bytes a warp keeps in flight are limited by its registers.
// Here a CTA is a three-role pipeline over a ring of NS shared-memory stages:
//   warp 0 (loader)   per stage: cp.async.bulk of the parameter / moment rows of RPS unique rows
//                     (3 x 8d bytes each, gathered by row id) + one contiguous block of their
//                     gradient rows (+ injected noise) -> smem, completion on an mbarrier;
//   warps 2.. (consumers)  chain rule + KL + Adam in place in shared memory (LPR lanes per row);
//   warp 1 (storer)   cp.async.bulk smem -> global of the updated rows, then frees the stage.
// Bytes in flight per SM = NS x stage bytes x resident CTAs (~170 KB), independent of registers.
// Rows are split into equal contiguous ranges per CTA (all CTAs finish together).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kBulkConsumers = 4;                     // consumer warps per CTA
constexpr int kBulkThreads = 32 * (2 + kBulkConsumers);
constexpr int kBulkStages = 4;

// floats of one stage: P | M | V rows (2d each), gradient rows, noise rows (d each), then row ids
__host__ __device__ inline int bulk_stage_floats(int d, int rps) { return rps * 8 * d + ((rps + 3) / 4) * 4; }

template <int LPR, int NV, int LINK>
__global__ void __launch_bounds__(kBulkThreads, 3)
k_adam_bulk(DevCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
            float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
            const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
            const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
            const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
            AdamDev h, int32_t* __restrict__ adam_step, float kl_scale, FinalArgs fa) {
    constexpr int VEC = 4, GPW = kWarp / LPR, RPS = kBulkConsumers * GPW, NS = kBulkStages;
    extern __shared__ __align__(128) float s_ring[];
    __shared__ __align__(8) uint64_t s_full[NS], s_done[NS], s_empty[NS];
    const int U = meta[0];
    const int d = c.d, rowf = 2 * c.d;
    const int stage_f = bulk_stage_floats(d, RPS);
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < NS; ++s2) { mbar_init(&s_full[s2], 1); mbar_init(&s_done[s2], kBulkConsumers); mbar_init(&s_empty[s2], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // equal contiguous ranges of unique rows per CTA, in whole stages
    const int per = ((U + (int)gridDim.x - 1) / (int)gridDim.x + RPS - 1) / RPS * RPS;
    const int lo = min(U, (int)blockIdx.x * per), hi = min(U, lo + per);
    const int n_it = (hi - lo + RPS - 1) / RPS;
    float facc = 0.f;

    if (warp == 0) {
        // ------------------------------------------------------------ loader
        int rowid_next = (lane < RPS && lo + lane < hi) ? __ldg(urec + 4 * (size_t)(lo + lane)) : 0;
        for (int it = 0; it < n_it; ++it) {
            const int st = it % NS, ph = (it / NS) & 1;
            const int u0 = lo + it * RPS, nrows = min(RPS, hi - u0);
            const int rowid = rowid_next;
            const int un = u0 + RPS + lane;
            rowid_next = (lane < RPS && un < hi) ? __ldg(urec + 4 * (size_t)un) : 0;
            float* sp = s_ring + (size_t)st * stage_f;
            mbar_wait(&s_empty[st], ph ^ 1);
            if (lane < nrows) reinterpret_cast<int*>(sp + RPS * 8 * d)[lane] = rowid;
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)nrows * (uint32_t)(3 * rowf + d + (eps_entity ? d : 0)) * 4u;
                mbar_expect_tx(&s_full[st], bytes);
            }
            __syncwarp();
            if (lane == 0) {
                bulk_load(sp + 3 * RPS * rowf, grow + (size_t)u0 * d, nrows * d * 4, &s_full[st]);
                if (eps_entity)
                    bulk_load(sp + 3 * RPS * rowf + RPS * d, eps_entity + (size_t)u0 * d, nrows * d * 4, &s_full[st]);
            }
            if (lane < nrows) {
                const size_t eoff = (size_t)rowid * rowf;
                bulk_load(sp + (0 * RPS + lane) * rowf, entity + eoff, rowf * 4, &s_full[st]);
                bulk_load(sp + (1 * RPS + lane) * rowf, entity_m + eoff, rowf * 4, &s_full[st]);
                bulk_load(sp + (2 * RPS + lane) * rowf, entity_v + eoff, rowf * 4, &s_full[st]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ storer
        for (int it = 0; it < n_it; ++it) {
            const int st = it % NS, ph = (it / NS) & 1;
            const int u0 = lo + it * RPS, nrows = min(RPS, hi - u0);
            float* sp = s_ring + (size_t)st * stage_f;
            mbar_wait(&s_done[st], ph);
            if (lane < nrows) {
                const int rowid = reinterpret_cast<const int*>(sp + RPS * 8 * d)[lane];
                const size_t eoff = (size_t)rowid * rowf;
                bulk_store(entity + eoff, sp + (0 * RPS + lane) * rowf, rowf * 4);
                bulk_store(entity_m + eoff, sp + (1 * RPS + lane) * rowf, rowf * 4);
                bulk_store(entity_v + eoff, sp + (2 * RPS + lane) * rowf, rowf * 4);
            }
            bulk_commit();
            bulk_wait_read<1>();                              // the stores of stage it-1 have left smem
            __syncwarp();
            if (lane == 0 && it > 0) mbar_arrive(&s_empty[(it - 1) % NS]);
        }
        bulk_wait_read<0>();
        __syncwarp();
        if (lane == 0 && n_it > 0) mbar_arrive(&s_empty[(n_it - 1) % NS]);
        bulk_wait<0>();                                        // writes complete before the CTA retires
    } else {
        // ------------------------------------------------------------ consumers
        const int cw = warp - 2, gl = lane % LPR, gidx = lane / LPR;
        const unsigned gmask = group_mask<LPR>();
        float step_size, inv_bc2;
        adam_coeffs(h, (int)step + 1, &step_size, &inv_bc2);
        // scalars of the rows this warp handles: lane g < GPW owns row cw*GPW + g of every stage;
        // fetched one stage ahead
        auto fetch = [&](int it, int4& rec, float& cqv, float& gwv, float& ebv) {
            const int u = lo + it * RPS + cw * GPW + lane;
            rec = make_int4(0, 0, 0, 0); cqv = 0.f; gwv = 0.f; ebv = 0.f;
            if (lane < GPW && it < n_it && u < hi) {
                rec = __ldg(reinterpret_cast<const int4*>(urec) + u);
                cqv = __ldg(cq + u); gwv = __ldg(gws + u); ebv = __ldg(eps_bias + u);
            }
        };
        int4 rec_n; float cq_n, gw_n, eb_n;
        fetch(0, rec_n, cq_n, gw_n, eb_n);
        for (int it = 0; it < n_it; ++it) {
            const int st = it % NS, ph = (it / NS) & 1;
            const int u0 = lo + it * RPS;
            const int4 rec = rec_n; const float cq_l = cq_n, eb_l = eb_n; float gw_l = gw_n;
            fetch(it + 1, rec_n, cq_n, gw_n, eb_n);
            const int ul = u0 + cw * GPW + lane;
            const bool valid = lane < GPW && ul < hi;
            // bias row of this lane's row: in flight while the stage is awaited and processed
            float2 ab = make_float2(0.f, 1.f), bm = make_float2(0.f, 0.f), bv = make_float2(0.f, 0.f);
            if (valid) {
                const size_t boff = (size_t)rec.x * 2;
                ab = *reinterpret_cast<const float2*>(bias + boff);
                bm = *reinterpret_cast<const float2*>(bias_m + boff);
                bv = *reinterpret_cast<const float2*>(bias_v + boff);
            }                                              // (cut rows: k_combine_cut finished them)
            const float cfac_l = kl_scale * cq_l;
            float* sp = s_ring + (size_t)st * stage_f;
            mbar_wait(&s_full[st], ph);
            // ---- wide work: this warp's GPW rows, LPR lanes per row
            const int r = cw * GPW + gidx;                  // row slot in the stage
            const int u = u0 + r;
            const int rowid = bcast(rec.x, gidx);
            const float cfac = bcast(cfac_l, gidx);
            float kl = 0.f;
            if (u < hi) {
                float* pP = sp + (0 * RPS + r) * rowf;
                float* pM = sp + (1 * RPS + r) * rowf;
                float* pV = sp + (2 * RPS + r) * rowf;
                const float* pG = sp + 3 * RPS * rowf + r * d;
                const float* pE = pG + RPS * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> mu = ld_vec<VEC>(pP + k), rho = ld_vec<VEC>(pP + d + k);
                        Vec<VEC> m1 = ld_vec<VEC>(pM + k), m2 = ld_vec<VEC>(pM + d + k);
                        Vec<VEC> v1 = ld_vec<VEC>(pV + k), v2 = ld_vec<VEC>(pV + d + k);
                        const Vec<VEC> g = ld_vec<VEC>(pG + k);
                        Vec<VEC> e;
                        if (eps_entity) e = ld_vec<VEC>(pE + k);
                        else e = entity_eps<VEC>(nullptr, c, u, rowid * c.row_stride + c.row_offset, k, step);
                        float quad = 0.f, prodv = 1.f;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float sig = link_fn<LINK>(rho.v[j]);
                            const float gj = g.v[j];
                            const float gmu = fmaf(cfac, mu.v[j], gj);
                            const float grho = link_grad<LINK>(rho.v[j]) * fmaf(gj, e.v[j], cfac * (sig - fast_rcp(sig)));
                            const float vr = sig * sig;
                            quad += vr + mu.v[j] * mu.v[j] - 1.f;
                            prodv *= vr;
                            adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu, h, step_size, inv_bc2);
                            adam_elem(rho.v[j], m2.v[j], v2.v[j], grho, h, step_size, inv_bc2);
                        }
                        // KL of the pre-update row (quad / prodv were formed before adam_elem touched
                        // element j); one logarithm per lane as in k_stage
                        float lg = __logf(prodv);
                        if (!(prodv > 1e-30f && prodv < 1e30f)) {       // tiny / huge scales: no product trick
                            const Vec<VEC> rho0 = ld_vec<VEC>(pP + d + k);   // still the old row in smem
                            lg = 0.f;
#pragma unroll
                            for (int j = 0; j < VEC; ++j) { const float sg = link_fn<LINK>(rho0.v[j]); lg += logf(sg * sg); }
                        }
                        kl += 0.5f * (quad - lg);
                        st_vec<VEC>(pP + k, mu); st_vec<VEC>(pP + d + k, rho);
                        st_vec<VEC>(pM + k, m1); st_vec<VEC>(pM + d + k, m2);
                        st_vec<VEC>(pV + k, v1); st_vec<VEC>(pV + d + k, v2);
                    }
                }
                kl = group_sum<LPR>(kl, gmask);
            }
            fence_async_smem();                             // generic-proxy writes -> visible to the bulk store
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_done[st]);
            // ---- lane-parallel: bias row, KL of the row
            float klrow = 0.f;
#pragma unroll
            for (int g = 0; g < GPW; ++g) {
                const float kv = __shfl_sync(0xffffffffu, kl, g * LPR);
                if (lane == g) klrow = kv;
            }
            if (valid) {
                const size_t boff = (size_t)rec.x * 2;
                const float tau = link_fn<LINK>(ab.y);
                facc = fmaf(cq_l, klrow + kl_std_normal(ab.x, tau), facc);
                const float ga = fmaf(cfac_l, ab.x, gw_l);
                const float gb = link_grad<LINK>(ab.y) * fmaf(gw_l, eb_l, cfac_l * (tau - fast_rcp(tau)));
                adam_elem(ab.x, bm.x, bv.x, ga, h, step_size, inv_bc2);
                adam_elem(ab.y, bm.y, bv.y, gb, h, step_size, inv_bc2);
                *reinterpret_cast<float2*>(bias + boff) = ab;
                *reinterpret_cast<float2*>(bias_m + boff) = bm;
                *reinterpret_cast<float2*>(bias_v + boff) = bv;
            }
        }
    }
    // the block that finishes last owns the scalar parameters (see k_adam_rows)
    double acc[1] = {(double)facc};
    if (block_partials<1>(acc, fa.partials, fa.counter)) {
        double tot[1] = {0.0};
        final_sums<1>(fa.partials, tot);
        if (threadIdx.x == 0) {
            final_scalars<LINK, VFMB_ADAM_TOUCHED>(c, fa, h, adam_step, kl_scale, true, tot[0], U);
            *fa.counter = 0;
        }
    }
}

#endif  // VFMB_WITH_BULK

// ------------------------------------------------------------------------------- dense Adam
__global__ void __launch_bounds__(256)
k_adam_dense(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
             const float* __restrict__ g, int64_t n, AdamDev h, const int32_t* __restrict__ adam_step) {
    __shared__ float s_coef[2];
    if (threadIdx.x == 0) adam_coeffs(h, adam_step[0] + 1, &s_coef[0], &s_coef[1]);
    __syncthreads();
    const float ss = s_coef[0], b2 = s_coef[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_elem(pi, mi, vi, g[i], h, ss, b2);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}
__global__ void k_step_advance(int32_t* adam_step) { adam_step[0] += 1; }

// ------------------------------------------------------------------------------- philox export
template <int VEC>
__global__ void k_philox_export(DevCfg c, const int32_t* __restrict__ uniq, int U, uint32_t step,
                                float* eps_global, float* eps_bias, float* eps_entity) {
    // layout of the reference's draws (vfm-torch.py:238-241): [S,1], [S,U], [S,U,d]
    int nvec = (c.d + VEC - 1) / VEC;
    const int64_t per = (int64_t)U * nvec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per * c.S;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / per);
        const int64_t r = i - (int64_t)q * per;
        int u = (int)(r / nvec), j = (int)(r % nvec);
        int rowid = uniq[u];
        Vec<VEC> e = entity_eps<VEC>(nullptr, c, u, rowid, j * VEC, step, q, U);
        for (int t = 0; t < VEC; ++t)
            if (j * VEC + t < c.d) eps_entity[((size_t)q * U + u) * c.d + j * VEC + t] = e.v[t];
        if (j == 0) eps_bias[(size_t)q * U + u] = bias_eps(nullptr, c, u, rowid, step, q, U);
        if (r == 0) eps_global[q] = global_eps(nullptr, c, step, q);
    }
}

// ------------------------------------------------------------------------------- dispatch
bool pick_layout(int d, Layout* out) {
    int vec = (d % 4 == 0) ? 4 : 1;
    int nvec = (d + vec - 1) / vec;
    int lpr = 4;
    while (lpr < 32 && lpr < nvec) lpr *= 2;
    int nv = (nvec + lpr - 1) / lpr;
    if (nv > 2) return false;
    out->vec = vec; out->lpr = lpr; out->nv = nv;
    return true;
}

}  // namespace vfmb

using namespace vfmb;

extern "C" int64_t vfmb_partials_doubles(const vfmb_config* cfg) {
    if (!cfg) return 0;
    int64_t n = (int64_t)cfg->B * cfg->F;
    int64_t u_cap = n < cfg->R ? n : cfg->R;
    return (int64_t)scratch_map(cfg->B, cfg->F, cfg->d, u_cap).total_doubles;
}

// ---- shared host-side preparation of one phase launch
struct Prep {
    cudaStream_t stream;
    Layout L;
    DevCfg dc;
    vfmb_plan_capacity_t cap;
    int ch;
};
static int prep(const vfmb_config* cfg, const char* who, vfmb_stream stream_, int min_fields, Prep* p) {
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

// ---- internal launchers (the public entry points below are thin wrappers)
static int launch_stage(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                        const vfmb_step_io* io, vfmb_stream stream_, bool lean, int smp = 0) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_stage", stream_, 1, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: null argument");
    if (!lean && !io->eps_entity && !io->es) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: noise scratch (es) required");
    if (!io->eps_bias && !io->ebs) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: noise scratch (ebs) required");
    if (!io->vs || !io->ws || !io->cq || !io->partials || !io->counters || !io->stats)
        return set_error(VFMB_EINVAL, "vfmb_sampled_stage: scratch required");
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream; const int ch = P.ch; const auto& cap = P.cap;
#define LAUNCH_STAGE(LINK, LEAN)                                                                       \
    k_stage<VEC, LPR, NV, LINK, LEAN><<<grid_resident(k_stage<VEC, LPR, NV, LINK, LEAN>, cap.u_cap, ch), 256, 0, stream>>>( \
        dc, tab->bias, tab->entity, tab->train_counts, plan->urec, plan->meta, plan->z,                \
        io->eps_bias, io->eps_entity, tab->adam_step, io->vs, io->ws, io->es,                          \
        io->ebs, io->cq, io->partials, io->counters + 0, io->stats, smp, (int)cap.u_cap)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) { if (lean) LAUNCH_STAGE(0, 1); else LAUNCH_STAGE(0, 0); }
        else                            { if (lean) LAUNCH_STAGE(1, 1); else LAUNCH_STAGE(1, 0); }
    });
#undef LAUNCH_STAGE
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int launch_score(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                        const vfmb_step_io* io, vfmb_stream stream_, int defer_kl) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_score", stream_, 2, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_score: null argument");
    if (cfg->F > 2 && io->y && !io->msg) return set_error(VFMB_EINVAL, "vfmb_sampled_score: msg scratch required for F>2");
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream; const int ch = P.ch;
#define LAUNCH_SCORE(LINK, LIK)                                                                        \
    k_score<VEC, LPR, NV, LINK, LIK><<<grid_resident(k_score<VEC, LPR, NV, LINK, LIK>, cfg->B, ch), 256, 0, stream>>>( \
        dc, tab->scalars, plan->inverse, plan->pos_of, io->vs, io->ws, io->y, io->eps_global,          \
        tab->adam_step, io->pred, io->mean, io->resid, io->rsorted, io->msg, io->partials,             \
        io->counters + 1, io->stats, defer_kl)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SCORE(0, VFMB_GAUSSIAN); else LAUNCH_SCORE(0, VFMB_BERNOULLI);
        } else {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SCORE(1, VFMB_GAUSSIAN); else LAUNCH_SCORE(1, VFMB_BERNOULLI);
        }
    });
#undef LAUNCH_SCORE
    CUDA_TRY(cudaGetLastError());
    return 0;
}

#define VFMB_PHASE_S1(who)                                                                             \
    if (cfg && cfg->S != 1) return set_error(VFMB_ESHAPE, who ": the phase entry points take S = 1 (use "    \
                                             "vfmb_sampled_forward / _backward / _step for S > 1)")

extern "C" int vfmb_sampled_stage(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                  const vfmb_step_io* io, vfmb_stream stream) {
    VFMB_PHASE_S1("vfmb_sampled_stage");
    return launch_stage(cfg, tab, plan, io, stream, false);
}

extern "C" int vfmb_sampled_score(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                  const vfmb_step_io* io, vfmb_stream stream) {
    VFMB_PHASE_S1("vfmb_sampled_score");
    return launch_score(cfg, tab, plan, io, stream, 0);
}

// S > 1: one k_stage launch per variational sample (non-lean: the noise of every sample is kept in
// the [S][u_cap] scratch for the backward), then k_score_multi
static int forward_multi(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                         const vfmb_step_io* io, vfmb_stream stream_) {
    for (int q = 0; q < cfg->S; ++q) {
        int rc = launch_stage(cfg, tab, plan, io, stream_, false, q);
        if (rc) return rc;
    }
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_forward", stream_, 2, &P);
    if (rc) return rc;
    if (cfg->F > 2 && io->y && !io->msg) return set_error(VFMB_EINVAL, "vfmb_sampled_forward: msg scratch required for F>2");
    if (!io->pred || !io->mean || !io->resid || !io->rsorted) return set_error(VFMB_EINVAL, "vfmb_sampled_forward: outputs required");
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream;
    int grid = (int)((cfg->B + 8 * (32 / L.lpr) - 1) / (8 * (32 / L.lpr)));
    if (grid > kGridCap) grid = kGridCap;
#define LAUNCH_SM(LINK, LIK)                                                                             \
    k_score_multi<VEC, LPR, NV, LINK, LIK><<<grid, 256, 0, stream>>>(                                    \
        dc, (int)P.cap.u_cap, tab->scalars, plan->inverse, plan->pos_of, io->vs, io->ws, io->y, io->eps_global, \
        tab->adam_step, io->pred, io->mean, io->resid, io->rsorted, io->msg, io->partials, io->counters + 1, io->stats)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SM(0, VFMB_GAUSSIAN); else LAUNCH_SM(0, VFMB_BERNOULLI);
        } else {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SM(1, VFMB_GAUSSIAN); else LAUNCH_SM(1, VFMB_BERNOULLI);
        }
    });
#undef LAUNCH_SM
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_sampled_forward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                    const vfmb_step_io* io, vfmb_stream stream) {
    if (cfg && cfg->S > 1) {
        if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_forward: null argument");
        return forward_multi(cfg, tab, plan, io, stream);
    }
    int rc = launch_stage(cfg, tab, plan, io, stream, false);
    if (rc) return rc;
    return launch_score(cfg, tab, plan, io, stream, 0);
}

static int launch_gather(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io_,
                         const float* table, int32_t unit_coef, vfmb_stream stream_, int smp = 0) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_gather", stream_, 1, &P);
    if (rc) return rc;
    if (!plan || !io_) return set_error(VFMB_EINVAL, "vfmb_sampled_gather: null argument");
    vfmb_step_io shifted = *io_;                           // sample smp of the [S][u_cap] / [S][B] scratch
    if (smp > 0) {
        const size_t us = (size_t)P.cap.u_cap;
        shifted.vs += smp * us * cfg->d;
        if (shifted.grow) shifted.grow += smp * us * cfg->d;
        if (shifted.gws) shifted.gws += smp * us;
        if (shifted.msg) shifted.msg += (size_t)smp * cfg->B * cfg->d;
    }
    const vfmb_step_io* io = &shifted;
    if (!io->grow || !io->gws || !io->rsorted || !io->partials) return set_error(VFMB_EINVAL, "vfmb_sampled_gather: scratch required");
    const Layout& L = P.L; cudaStream_t stream = P.stream; const auto& cap = P.cap;
    float* gslot = (float*)io->partials + scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap).gslot_off;
    // F == 2: partner rows come from the sampled-row scratch; otherwise from `table`
    // (F > 2: the per-sample field sums written by k_score; unit_coef: received gradient rows)
    const float* tbl = table ? table : io->msg;
    if (cfg->F != 2 && !tbl) return set_error(VFMB_EINVAL, "vfmb_sampled_gather: table required for F != 2");
    const int N = cfg->B * cfg->F;
    VFMB_LAYOUT_SWITCH(L, {
        if (unit_coef)
            k_gather<VEC, LPR, NV, 1><<<grid_resident(k_gather<VEC, LPR, NV, 1>, cap.n_tiles, 32 / L.lpr), 256, 0, stream>>>(
                cfg->d, cfg->F, N, plan->partner, plan->pos_rank, io->vs, tbl, io->rsorted, gslot, io->grow, io->gws);
        else
            k_gather<VEC, LPR, NV, 0><<<grid_resident(k_gather<VEC, LPR, NV, 0>, cap.n_tiles, 32 / L.lpr), 256, 0, stream>>>(
                cfg->d, cfg->F, N, plan->partner, plan->pos_rank, io->vs, tbl, io->rsorted, gslot, io->grow, io->gws);
        const size_t smem = 8 * GPW_OF(LPR) * (cfg->d + 4) * sizeof(float);
        if (plan->hot)      // lists of the cut rows from the plan: no scan over the unique rows
            k_combine_cut<VEC, LPR, NV><<<grid_warps(cap.n_tiles, 1), 256, smem, stream>>>(
                cfg->d, unit_coef ? 2 : cfg->F, plan->urec, plan->meta, plan->hot, (int)cut_list_capacity(cap.n_tiles),
                gslot, io->vs, io->grow, io->gws);
        else
            k_combine<VEC, LPR, NV, 0><<<grid_warps(cap.u_cap, 32), 256, smem, stream>>>(
                cfg->d, unit_coef ? 2 : cfg->F, plan->urec, plan->meta, gslot, io->vs, io->grow, io->gws);
    });
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// F == 2 fused step: k_gather_score (score + ordered segmented sum) + k_combine_cut
static int launch_gather_score(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                               const vfmb_step_io* io, vfmb_stream stream_) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_step", stream_, 2, &P);
    if (rc) return rc;
    if (!io->grow || !io->gws || !io->partials || !io->pred || !io->mean || !io->resid || !plan->hot || !plan->occ)
        return set_error(VFMB_EINVAL, "vfmb_sampled_step: scratch required");
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream; const auto& cap = P.cap;
    float* gslot = (float*)io->partials + scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap).gslot_off;
    // one lane group per tile.  (Smaller blocks were tried for the plan-overlapped step -- 121.9 us with
    // 128 threads, 125.2 us with 64, against 113.4 us for the unfused pair: the fused kernel does not
    // lose to the plan through SM slots; VFMB_GS_BLOCK overrides for experiments.)
    static const int gs_block_env = [] { const char* e = getenv("VFMB_GS_BLOCK"); return e ? atoi(e) : 0; }();
    const int gs_block = gs_block_env ? gs_block_env : 256;
    const int64_t gs_warps = (cap.n_tiles + (32 / L.lpr) - 1) / (32 / L.lpr);
    int gs_grid = (int)((gs_warps + gs_block / 32 - 1) / (gs_block / 32));
    if (gs_grid > kGridCap) gs_grid = kGridCap;            // block partials are sized for kGridCap blocks
#define LAUNCH_GS(LINK, LIK)                                                                             \
    k_gather_score<VEC, LPR, NV, LINK, LIK><<<gs_grid, gs_block, 0, stream>>>(                           \
        dc, tab->scalars, plan->partner, plan->pos_rank, plan->occ, io->vs, io->ws, io->y, io->eps_global, \
        tab->adam_step, io->pred, io->mean, io->resid, gslot, io->grow, io->gws, io->partials,           \
        io->counters + 1, io->stats)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_GS(0, VFMB_GAUSSIAN); else LAUNCH_GS(0, VFMB_BERNOULLI);
        } else {
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_GS(1, VFMB_GAUSSIAN); else LAUNCH_GS(1, VFMB_BERNOULLI);
        }
        const size_t smem = 8 * GPW_OF(LPR) * (cfg->d + 4) * sizeof(float);
        k_combine_cut<VEC, LPR, NV><<<grid_warps(cap.n_tiles, 1), 256, smem, stream>>>(
            cfg->d, cfg->F, plan->urec, plan->meta, plan->hot, (int)cut_list_capacity(cap.n_tiles),
            gslot, io->vs, io->grow, io->gws);
    });
#undef LAUNCH_GS
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_sampled_gather(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io,
                                   const float* table, int32_t unit_coef, vfmb_stream stream) {
    VFMB_PHASE_S1("vfmb_sampled_gather");
    return launch_gather(cfg, plan, io, table, unit_coef, stream);
}

// flavor: see k_adam_rows
static int launch_adam(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                       const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode, float kl_grad_scale,
                       int flavor, vfmb_stream stream_) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_adam_rows", stream_, 1, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: null argument");
    if (mode == VFMB_ADAM_TOUCHED && (!adam || !tab->entity_m || !tab->entity_v || !tab->bias_m || !tab->bias_v || !tab->adam_step))
        return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: Adam state required");
    if (mode == VFMB_GRAD_ONLY && (!io->grad_bias || !io->grad_entity))
        return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: gradient outputs required");
    if (mode != VFMB_ADAM_TOUCHED && mode != VFMB_GRAD_ONLY) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: bad mode");
    if (!io->grow || !io->gws || !io->cq) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: scratch required");
    if (flavor >= 1 && mode == VFMB_ADAM_TOUCHED && (!tab->scalars_m || !tab->scalars_v))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: Adam state required");
    if (flavor >= 1 && (!tab->scalars || !io->stats || !io->partials || !io->counters))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: scalars / stats required");
    const Layout& L = P.L; cudaStream_t stream = P.stream; const int ch = P.ch; const auto& cap = P.cap;
    const DevCfg& dc = P.dc;
    AdamDev h = make_adam(adam);
    // the noise the forward used: injected arrays, or what k_stage wrote to scratch (Philox);
    // flavor 2 recomputes the Philox draws of the rows instead
    const float* eps_e = io->eps_entity ? io->eps_entity : (flavor == 2 ? nullptr : io->es);
    const float* eps_b = io->eps_bias ? io->eps_bias : io->ebs;
    if (!eps_b || (flavor != 2 && !eps_e)) return set_error(VFMB_EINVAL, "vfmb_sampled_adam_rows: noise of the forward required");
    FinalArgs fa{};
    fa.scalars = tab->scalars; fa.sm = tab->scalars_m; fa.sv = tab->scalars_v; fa.stats = io->stats;
    fa.eps_global = io->eps_global; fa.grad_scalars = io->grad_scalars; fa.partials = io->partials;
    fa.counter = io->counters ? io->counters + 2 : nullptr;
    fa.gslot = io->partials ? (const float*)io->partials + scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap).gslot_off : nullptr;
    fa.likelihood = cfg->likelihood;
#ifdef VFMB_WITH_BULK
    // measured on ml20m: 60.7 us vs 53.9 us for k_adam_rows -- 512-byte rows are too small for the
    // bulk-copy engine (per-copy overhead); kept selectable for wide rows (VFMB_ADAM_BULK=1)
    static const bool use_bulk = [] { const char* e = getenv("VFMB_ADAM_BULK"); return e && atoi(e) != 0; }();
    if (use_bulk && flavor == 2 && mode == VFMB_ADAM_TOUCHED && L.vec == 4) {
        // bulk-copy pipeline (k_adam_bulk): same arithmetic, rows staged through shared memory
        cudaEvent_t ev0, ev1;
        profile_events(&ev0, &ev1);
        if (ev0 && ev1) cudaEventRecord(ev0, stream);
#define LAUNCH_BULK(LPR_, NV_, LINK)                                                                     \
        do {                                                                                             \
            auto kern = k_adam_bulk<LPR_, NV_, LINK>;                                                    \
            const int rps = kBulkConsumers * (32 / LPR_);                                                \
            const size_t smem = (size_t)kBulkStages * bulk_stage_floats(cfg->d, rps) * sizeof(float);    \
            static int blocks = 0;                                                                       \
            static size_t smem_set = 0;                                                                  \
            if (smem_set != smem) {                                                                      \
                CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                int per_sm = 0;                                                                          \
                CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBulkThreads, smem)); \
                blocks = (per_sm < 1 ? 1 : per_sm) * kNumSMs;                                            \
                smem_set = smem;                                                                         \
            }                                                                                            \
            int64_t need = (cap.u_cap + rps - 1) / rps;                                                  \
            int grid = (int)(need < blocks ? (need < 1 ? 1 : need) : blocks);                            \
            kern<<<grid, kBulkThreads, smem, stream>>>(dc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, \
                tab->entity_m, tab->entity_v, plan->urec, plan->meta, eps_b, eps_e, io->cq, io->grow,    \
                io->gws, h, tab->adam_step, kl_grad_scale, fa);                                          \
        } while (0)
#define LAUNCH_BULK_L(LPR_, NV_) do { if (cfg->link == VFMB_LINK_ABS) LAUNCH_BULK(LPR_, NV_, 0); else LAUNCH_BULK(LPR_, NV_, 1); } while (0)
        if (L.lpr == 4) LAUNCH_BULK_L(4, 1);
        else if (L.lpr == 8) LAUNCH_BULK_L(8, 1);
        else if (L.lpr == 16) LAUNCH_BULK_L(16, 1);
        else if (L.nv == 1) LAUNCH_BULK_L(32, 1);
        else LAUNCH_BULK_L(32, 2);
#undef LAUNCH_BULK_L
#undef LAUNCH_BULK
        if (ev0 && ev1) cudaEventRecord(ev1, stream);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
#endif  // VFMB_WITH_BULK
    // the HBM-bound kernel keeps its full wave even next to the plan (measured: 113.8 vs 116.2 us/step)
    static const bool adam_reserve = [] { const char* e = getenv("VFMB_RESERVE_ADAM"); return e && atoi(e) != 0; }();
#define LAUNCH_ADAM(LINK, MODE, FLAVOR)                                                                  \
    k_adam_rows<VEC, LPR, NV, LINK, MODE, FLAVOR><<<grid_resident(k_adam_rows<VEC, LPR, NV, LINK, MODE, FLAVOR>, cap.u_cap, ch, 0, adam_reserve), 256, 0, stream>>>( \
        dc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v,              \
        plan->urec, plan->meta, eps_b, eps_e, io->cq, io->grow, io->gws, h, tab->adam_step,              \
        kl_grad_scale, io->grad_bias, io->grad_entity, fa)
#define LAUNCH_ADAM_F(LINK, MODE)                                                                        \
    do { if (flavor == 0) LAUNCH_ADAM(LINK, MODE, 0); else if (flavor == 1) LAUNCH_ADAM(LINK, MODE, 1);  \
         else LAUNCH_ADAM(LINK, MODE, 2); } while (0)
    VFMB_LAYOUT_SWITCH(L, {
        cudaEvent_t ev0, ev1;
        profile_events(&ev0, &ev1);
        if (ev0 && ev1) cudaEventRecord(ev0, stream);
        if (cfg->link == VFMB_LINK_ABS) {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_ADAM_F(0, VFMB_ADAM_TOUCHED); else LAUNCH_ADAM_F(0, VFMB_GRAD_ONLY);
        } else {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_ADAM_F(1, VFMB_ADAM_TOUCHED); else LAUNCH_ADAM_F(1, VFMB_GRAD_ONLY);
        }
        if (ev0 && ev1) cudaEventRecord(ev1, stream);
    });
#undef LAUNCH_ADAM_F
#undef LAUNCH_ADAM
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_sampled_adam_rows(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                      const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                      float kl_grad_scale, vfmb_stream stream) {
    VFMB_PHASE_S1("vfmb_sampled_adam_rows");
    return launch_adam(cfg, tab, plan, io, adam, mode, kl_grad_scale, 0, stream);
}

// S > 1: k_adam_rows_multi on the [S][u_cap] gradients / noise
static int launch_adam_multi(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                             const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode, float kl_grad_scale,
                             vfmb_stream stream_) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_backward", stream_, 2, &P);
    if (rc) return rc;
    if (mode == VFMB_ADAM_TOUCHED && (!adam || !tab->entity_m || !tab->entity_v || !tab->bias_m || !tab->bias_v ||
                                      !tab->adam_step || !tab->scalars_m || !tab->scalars_v))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: Adam state required");
    if (mode == VFMB_GRAD_ONLY && (!io->grad_bias || !io->grad_entity))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: gradient outputs required");
    if (mode != VFMB_ADAM_TOUCHED && mode != VFMB_GRAD_ONLY) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: bad mode");
    if (!io->grow || !io->gws || !io->cq || !tab->scalars || !io->stats || !io->partials || !io->counters)
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: scratch required");
    // the noise of the forward: injected arrays [S, U, ...] or what k_stage kept [S, u_cap, ...]
    const float* eps_e = io->eps_entity ? io->eps_entity : io->es;
    const float* eps_b = io->eps_bias ? io->eps_bias : io->ebs;
    if (!eps_e || !eps_b || (io->eps_entity != nullptr) != (io->eps_bias != nullptr))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: noise of the forward required (both arrays injected, or none)");
    const int eps_stride = io->eps_entity ? -1 : (int)P.cap.u_cap;
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream; const auto& cap = P.cap;
    AdamDev h = make_adam(adam);
    FinalArgs fa{};
    fa.scalars = tab->scalars; fa.sm = tab->scalars_m; fa.sv = tab->scalars_v; fa.stats = io->stats;
    fa.eps_global = io->eps_global; fa.grad_scalars = io->grad_scalars; fa.partials = io->partials;
    fa.counter = io->counters + 2; fa.gslot = nullptr; fa.likelihood = cfg->likelihood;
    int grid = (int)((cap.u_cap + 8 * (32 / L.lpr) - 1) / (8 * (32 / L.lpr)));
    if (grid > kGridCap) grid = kGridCap;
#define LAUNCH_AM(LINK, MODE)                                                                            \
    k_adam_rows_multi<VEC, LPR, NV, LINK, MODE><<<grid, 256, 0, stream>>>(                               \
        dc, (int)cap.u_cap, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v, \
        plan->urec, plan->meta, eps_b, eps_e, eps_stride, io->cq, io->grow, io->gws, h, tab->adam_step,  \
        kl_grad_scale, io->grad_bias, io->grad_entity, fa)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_AM(0, VFMB_ADAM_TOUCHED); else LAUNCH_AM(0, VFMB_GRAD_ONLY);
        } else {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_AM(1, VFMB_ADAM_TOUCHED); else LAUNCH_AM(1, VFMB_GRAD_ONLY);
        }
    });
#undef LAUNCH_AM
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int backward_impl(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                         const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                         float kl_grad_scale, int flavor, vfmb_stream stream_) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_backward", stream_, 2, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: null argument");
    cudaStream_t stream = P.stream;
    if (mode == VFMB_GRAD_ONLY) {
        // the residuals may come from the caller's autograd: (re)build their sorted-order copy
        if (!io->resid || !io->rsorted) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: resid required");
        const int N = cfg->B * cfg->F;
        int g = (N + 255) / 256;
        if (g > 4 * kNumSMs) g = 4 * kNumSMs;
        k_scatter_resid<<<g, 256, 0, stream>>>(io->resid, plan->pos_of, N, cfg->F, io->rsorted);
        CUDA_TRY(cudaGetLastError());
    }
    if (cfg->S > 1) {       // one ordered segmented sum per variational sample, then the summed row update
        for (int q = 0; q < cfg->S; ++q) {
            rc = launch_gather(cfg, plan, io, nullptr, 0, stream_, q);
            if (rc) return rc;
        }
        return launch_adam_multi(cfg, tab, plan, io, adam, mode, kl_grad_scale, stream_);
    }
    rc = launch_gather(cfg, plan, io, nullptr, 0, stream_);
    if (rc) return rc;
    return launch_adam(cfg, tab, plan, io, adam, mode, kl_grad_scale, flavor, stream_);
}

extern "C" int vfmb_sampled_backward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                     const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                     float kl_grad_scale, vfmb_stream stream) {
    return backward_impl(cfg, tab, plan, io, adam, mode, kl_grad_scale, 1, stream);
}

// The fused training step: 5 launches -- k_stage<LEAN>, k_score, k_gather, k_combine_cut,
// k_adam_rows<FLAVOR 2> (KL, scalar parameters and the step counter folded in).
extern "C" int vfmb_sampled_step(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                 const vfmb_step_io* io, const vfmb_adam* adam, vfmb_stream stream) {
    if (io && !io->y) return set_error(VFMB_EINVAL, "vfmb_sampled_step: targets required");
    if (cfg && cfg->S > 1) {
        if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_step: null argument");
        int rc = forward_multi(cfg, tab, plan, io, stream);
        if (rc) return rc;
        return backward_impl(cfg, tab, plan, io, adam, VFMB_ADAM_TOUCHED, 1.0f, 1, stream);
    }
    int rc = launch_stage(cfg, tab, plan, io, stream, true);
    if (rc) return rc;
    // measured (ml20m): alone the fused kernel saves 5 us per step (104 -> 99 us); next to a
    // concurrently running plan it loses 6 us -- it needs 111 registers per thread, which leaves no
    // room on an SM for the plan's blocks.  So: fused unless the caller reserved room for the plan.
    static const int fuse_env = [] { const char* e = getenv("VFMB_FUSE_SCORE"); return e ? atoi(e) : -1; }();
    const bool fuse_gs = fuse_env >= 0 ? fuse_env != 0 : grid_reserve() == 0;
    if (cfg && cfg->F == 2 && plan && plan->hot && fuse_gs) {
        // F == 2: 4 launches -- k_stage<LEAN>, k_gather_score, k_combine_cut, k_adam_rows<FLAVOR 2>
        rc = launch_gather_score(cfg, tab, plan, io, stream);
        if (rc) return rc;
        return launch_adam(cfg, tab, plan, io, adam, VFMB_ADAM_TOUCHED, 1.0f, 2, stream);
    }
    rc = launch_score(cfg, tab, plan, io, stream, 1);
    if (rc) return rc;
    return backward_impl(cfg, tab, plan, io, adam, VFMB_ADAM_TOUCHED, 1.0f, 2, stream);
}

extern "C" int vfmb_adam_dense(float* p, float* m, float* v, const float* g, int64_t n, const vfmb_adam* adam,
                               const int32_t* adam_step, vfmb_stream stream_) {
    if (!p || !m || !v || !g || !adam || !adam_step || n < 0) return set_error(VFMB_EINVAL, "vfmb_adam_dense: bad argument");
    if (n == 0) return 0;
    AdamDev h = make_adam(adam);
    int64_t grid = (n + 255) / 256;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    k_adam_dense<<<(int)grid, 256, 0, (cudaStream_t)stream_>>>(p, m, v, g, n, h, adam_step);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_adam_step_advance(int32_t* adam_step, vfmb_stream stream_) {
    if (!adam_step) return set_error(VFMB_EINVAL, "vfmb_adam_step_advance: null");
    k_step_advance<<<1, 1, 0, (cudaStream_t)stream_>>>(adam_step);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_philox_normals(const vfmb_config* cfg, const int32_t* uniq, int32_t U, int32_t step,
                                   float* eps_global, float* eps_bias, float* eps_entity, vfmb_stream stream_) {
    if (!cfg || !uniq || !eps_global || !eps_bias || !eps_entity || U < 0) return set_error(VFMB_EINVAL, "vfmb_philox_normals: bad argument");
    if (U == 0) return 0;
    DevCfg dc = make_dev(cfg);
    int vec = (cfg->d % 4 == 0) ? 4 : 1;
    if (cfg->S < 1 || cfg->S > kMaxSamples) return set_error(VFMB_ESHAPE, "vfmb_philox_normals: S=%d (1..%d)", cfg->S, kMaxSamples);
    int64_t work = (int64_t)U * ((cfg->d + vec - 1) / vec) * cfg->S;
    int grid = (int)((work + 255) / 256 > 4096 ? 4096 : (work + 255) / 256);
    if (vec == 4) k_philox_export<4><<<grid, 256, 0, (cudaStream_t)stream_>>>(dc, uniq, U, (uint32_t)step, eps_global, eps_bias, eps_entity);
    else k_philox_export<1><<<grid, 256, 0, (cudaStream_t)stream_>>>(dc, uniq, U, (uint32_t)step, eps_global, eps_bias, eps_entity);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
