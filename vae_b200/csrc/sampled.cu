// Sampled-ELBO VFM step (vfm-torch.py:189-324 forward, :359 loss, :368-370 backward + Adam).
//
// Four kernels per training step, all HBM/L2-bound (no dense contraction exists on this path):
//   k_stage  one lane group per UNIQUE row: gather [mean|raw scale], draw eps (Philox or
//            injected), write the sampled row v = mu + eps*|rho| and bias w to an L2-resident
//            scratch, accumulate the count-rescaled KL.          (vfm-torch.py:207-241, 290-317)
//   k_score  one lane group per SAMPLE: FM interaction of the sampled rows, likelihood,
//            residual dloss/dpred.                                (vfm-torch.py:244-270, 359)
//   k_rows   one lane group per (unique row, <=64-occurrence chunk): ordered segmented sum of
//            residual * partner row, chain rule to (mu, rho) + KL gradient, Adam on the row.
//            Deterministic: fixed summation order, no floating-point atomics. (:368-370)
//   k_final  scalar parameters (alpha, global bias) and the step counter.
#include "common.cuh"
#include "internal.h"

namespace vfmb {

struct DevCfg {
    int B, F, d, S, n_classes;
    int class_bound[kMaxFields];
    float class_size[kMaxFields];
    float n_train;
    uint64_t seed;
};

static DevCfg make_dev(const vfmb_config* c) {
    DevCfg r{};
    r.B = c->B; r.F = c->F; r.d = c->d; r.S = c->S; r.n_classes = c->n_classes;
    for (int i = 0; i < kMaxFields; ++i) { r.class_bound[i] = c->class_bound[i]; r.class_size[i] = c->class_size[i]; }
    r.n_train = c->n_train; r.seed = c->seed;
    return r;
}

__device__ __forceinline__ int class_of(const DevCfg& c, int row) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaxFields - 1; ++i) k += (i < c.n_classes - 1 && row >= c.class_bound[i]) ? 1 : 0;
    return k;
}

// eps for the VEC elements starting at k of row `rowid` (unique rank u)
template <int VEC>
__device__ __forceinline__ Vec<VEC> entity_eps(const float* __restrict__ eps_entity, const DevCfg& c,
                                              int u, int rowid, int k, uint32_t step) {
    Vec<VEC> e;
    if (eps_entity) {
        e = ld_vec_nc<VEC>(eps_entity + (size_t)u * c.d + k);
    } else {
        float n4[4];
        philox_normal4(c.seed, (uint32_t)rowid, (uint32_t)(k / VEC), step, philox_tag(kTagEntity, 0), n4);
#pragma unroll
        for (int i = 0; i < VEC; ++i) e.v[i] = n4[i];
    }
    return e;
}
__device__ __forceinline__ float bias_eps(const float* __restrict__ eps_bias, const DevCfg& c, int u,
                                          int rowid, uint32_t step) {
    if (eps_bias) return __ldg(eps_bias + u);
    float n4[4];
    philox_normal4(c.seed, (uint32_t)rowid, 0xFFFFFFFFu, step, philox_tag(kTagBias, 0), n4);
    return n4[0];
}
__device__ __forceinline__ float global_eps(const float* __restrict__ eps_global, const DevCfg& c,
                                            uint32_t step) {
    if (eps_global) return __ldg(eps_global);
    float n4[4];
    philox_normal4(c.seed, 0xFFFFFFFFu, 0xFFFFFFFFu, step, philox_tag(kTagGlobal, 0), n4);
    return n4[0];
}

// block-wide deterministic reduction of K per-thread doubles into partials[block][K]; returns
// true in every thread of the block that arrived last (which then owns the final reduction)
template <int K>
__device__ __forceinline__ bool block_partials(const double (&acc)[K], double* __restrict__ partials,
                                               int32_t* __restrict__ counter) {
    __shared__ double s_red[K][8];
    __shared__ bool s_last;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double s = warp_sum(acc[i]);
        if (lane == 0) s_red[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[threadIdx.x][w];
        partials[(size_t)blockIdx.x * K + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

template <int K>
__device__ __forceinline__ double final_sum(const double* __restrict__ partials, int i) {
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(partials + (size_t)b * K + i);
    return s;
}

// ------------------------------------------------------------------------------- k_stage
template <int VEC, int LPR, int NV, int LINK>
__global__ void __launch_bounds__(256)
k_stage(DevCfg c, const float* __restrict__ bias, const float* __restrict__ entity,
        const float* __restrict__ train_counts, const int32_t* __restrict__ uniq,
        const int32_t* __restrict__ seg_off, const int32_t* __restrict__ meta,
        const float* __restrict__ z, int32_t* __restrict__ heavy_done,
        const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
        const int32_t* __restrict__ adam_step, float* __restrict__ vs, float* __restrict__ ws,
        double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR;                  // groups (rows) per warp
    const int U = meta[0];
    const int d = c.d;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int group = (threadIdx.x >> 5) * GPW + lane / LPR;
    const int groups_per_block = (blockDim.x >> 5) * GPW;
    double acc[kMaxFields];
#pragma unroll
    for (int i = 0; i < kMaxFields; ++i) acc[i] = 0.0;

    for (int u = blockIdx.x * groups_per_block + group; u < U; u += gridDim.x * groups_per_block) {
        const int rowid = uniq[u];
        const float* erow = entity + (size_t)rowid * 2 * d;
        float kl = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                Vec<VEC> mu = ld_vec<VEC>(erow + k), rho = ld_vec<VEC>(erow + d + k);
                Vec<VEC> e = entity_eps<VEC>(eps_entity, c, u, rowid, k, step), out;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float sig = link_fn<LINK>(rho.v[j]);
                    out.v[j] = mu.v[j] + e.v[j] * sig;
                    kl += kl_std_normal(mu.v[j], sig);
                }
                st_vec<VEC>(vs + (size_t)u * d + k, out);
            }
        }
        kl = group_sum<LPR>(kl, group_mask<LPR>());
        if (gl == 0) {
            float a = bias[(size_t)rowid * 2], b = bias[(size_t)rowid * 2 + 1];
            float tau = link_fn<LINK>(b);
            ws[u] = a + bias_eps(eps_bias, c, u, rowid, step) * tau;
            kl += kl_std_normal(a, tau);
            float q = (float)(seg_off[u + 1] - seg_off[u]) / __ldg(train_counts + rowid);
            int cls = class_of(c, rowid);
            double t = (double)(q * kl);
#pragma unroll
            for (int i = 0; i < kMaxFields; ++i) if (i == cls) acc[i] += t;
            heavy_done[u] = 0;
        }
    }
    if (block_partials<kMaxFields>(acc, partials, counter)) {
        if (threadIdx.x == 0) {
            double kl_rows = 0.0;
            for (int i = 0; i < c.n_classes; ++i)
                kl_rows += (double)(c.class_size[i] / z[i]) * final_sum<kMaxFields>(partials, i);
            stats[VFMB_ST_KL_ROWS] = (float)kl_rows;
            stats[VFMB_ST_U] = (float)U;
            *counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- k_score
template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_score(DevCfg c, const float* __restrict__ scalars, const int32_t* __restrict__ inverse,
        const float* __restrict__ vs, const float* __restrict__ ws, const float* __restrict__ y,
        const float* __restrict__ eps_global, const int32_t* __restrict__ adam_step,
        float* __restrict__ pred, float* __restrict__ mean, float* __restrict__ resid,
        float* __restrict__ msg, double* __restrict__ partials, int32_t* __restrict__ counter,
        float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR;
    const int d = c.d, F = c.F, B = c.B;
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int group = (threadIdx.x >> 5) * GPW + lane / LPR;
    const int groups_per_block = (blockDim.x >> 5) * GPW;
    const float mu0 = scalars[VFMB_S_GB_MEAN];
    const float sig0 = link_fn<LINK>(scalars[VFMB_S_GB_SCALE]);
    const float w0 = mu0 + global_eps(eps_global, c, step) * sig0;
    const float alpha = link_fn<LINK>(scalars[VFMB_S_ALPHA]);
    const float scale = c.n_train / ((float)c.S * (float)B);
    double acc[3] = {0.0, 0.0, 0.0};     // nll, resid, squared error

    for (int n = blockIdx.x * groups_per_block + group; n < B; n += gridDim.x * groups_per_block) {
        float part = 0.f, bsum = 0.f;
        Vec<VEC> ssum[NV];
        if (F == 2) {
            int r0 = __ldg(inverse + 2 * n), r1 = __ldg(inverse + 2 * n + 1);
            bsum = __ldg(ws + r0) + __ldg(ws + r1);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) {
                    Vec<VEC> a = ld_vec_nc<VEC>(vs + (size_t)r0 * d + k);
                    Vec<VEC> b = ld_vec_nc<VEC>(vs + (size_t)r1 * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) part += a.v[j] * b.v[j];
                }
            }
        } else {
            Vec<VEC> sq[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < VEC; ++j) { ssum[i].v[j] = 0.f; sq[i].v[j] = 0.f; }
            for (int f = 0; f < F; ++f) {
                int r = __ldg(inverse + (size_t)n * F + f);
                bsum += __ldg(ws + r);
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> a = ld_vec_nc<VEC>(vs + (size_t)r * d + k);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) { ssum[i].v[j] += a.v[j]; sq[i].v[j] += a.v[j] * a.v[j]; }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < VEC; ++j) part += 0.5f * (ssum[i].v[j] * ssum[i].v[j] - sq[i].v[j]);
        }
        const float inter = group_sum<LPR>(part, group_mask<LPR>());
        const float p = w0 + bsum + inter;
        float r = 0.f, mu_out = p;
        if (LIK == VFMB_BERNOULLI) mu_out = 1.f / (1.f + expf(-p));
        if (y) {
            const float yn = __ldg(y + n);
            float nll, err = yn - p;
            if (LIK == VFMB_GAUSSIAN) {
                nll = 0.5f * alpha * err * err - 0.5f * logf(alpha) + 0.9189385332046727f;
                r = scale * alpha * (p - yn);
            } else {
                nll = fmaxf(p, 0.f) - yn * p + log1pf(expf(-fabsf(p)));
                r = scale * (mu_out - yn);
            }
            if (gl == 0) { acc[0] += (double)nll; acc[1] += (double)r; acc[2] += (double)err * (double)err; }
            if (F > 2 && msg) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> o;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) o.v[j] = r * ssum[i].v[j];
                        st_vec<VEC>(msg + (size_t)n * d + k, o);
                    }
                }
            }
        }
        if (gl == 0) {
            pred[n] = p;
            mean[n] = mu_out;
            if (y) resid[n] = r;
        }
    }
    if (block_partials<3>(acc, partials, counter)) {
        if (threadIdx.x == 0) {
            double nll = final_sum<3>(partials, 0), sr = final_sum<3>(partials, 1), sq = final_sum<3>(partials, 2);
            float kl0 = kl_std_normal(mu0, sig0);
            float kl = kl0 + stats[VFMB_ST_KL_ROWS];
            stats[VFMB_ST_NLL_MEAN] = (float)(nll / (double)B);
            stats[VFMB_ST_SUM_RESID] = (float)sr;
            stats[VFMB_ST_SUM_SQERR] = (float)sq;
            stats[VFMB_ST_KL] = kl;
            stats[VFMB_ST_LOSS] = (float)((double)c.n_train * nll / (double)B + (double)kl);
            stats[VFMB_ST_W0] = w0;
            *counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- k_rows
// hyper-parameters as torch rounds them: every derived constant is formed in double first
struct AdamDev {
    double lr, beta1, beta2;
    float b2, omb1, omb2, eps;      // beta2, 1-beta1, 1-beta2, eps as fp32
};
static AdamDev make_adam(const vfmb_adam* a) {
    AdamDev h{};
    double lr = a ? a->lr : 1e-3, b1 = a ? a->beta1 : 0.9, b2 = a ? a->beta2 : 0.999, e = a ? a->eps : 1e-8;
    h.lr = lr; h.beta1 = b1; h.beta2 = b2;
    h.b2 = (float)b2; h.omb1 = (float)(1.0 - b1); h.omb2 = (float)(1.0 - b2); h.eps = (float)e;
    return h;
}

__device__ __forceinline__ void adam_coeffs(const AdamDev& h, int t, float* step_size, float* bc2_sqrt) {
    double bc1 = 1.0 - pow(h.beta1, (double)t);
    double bc2 = 1.0 - pow(h.beta2, (double)t);
    *step_size = (float)(h.lr / bc1);
    *bc2_sqrt = (float)sqrt(bc2);
}

// torch _single_tensor_adam: m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
// p.addcdiv_(m, sqrt(v)/bc2_sqrt + eps, -step_size)
__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, const AdamDev& h,
                                          float step_size, float bc2_sqrt) {
    m = m + (g - m) * h.omb1;
    v = v * h.b2 + h.omb2 * g * g;
    float denom = sqrtf(v) / bc2_sqrt + h.eps;
    p = p - step_size * (m / denom);
}

template <int VEC, int LPR, int NV, int LINK, int MODE>
__global__ void __launch_bounds__(256)
k_rows(DevCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
       float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
       const float* __restrict__ train_counts, const int32_t* __restrict__ uniq,
       const int32_t* __restrict__ inverse, const int32_t* __restrict__ seg_off,
       const int32_t* __restrict__ occ, const int32_t* __restrict__ item_first,
       const int32_t* __restrict__ item_row, int32_t* __restrict__ heavy_done,
       const float* __restrict__ z, const int32_t* __restrict__ meta,
       const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
       const float* __restrict__ vs, const float* __restrict__ msg, const float* __restrict__ resid,
       float* __restrict__ gpart, AdamDev h, const int32_t* __restrict__ adam_step, float kl_scale,
       float* __restrict__ grad_bias, float* __restrict__ grad_entity) {
    constexpr int GPW = kWarp / LPR;
    const int W = meta[1];
    const int d = c.d, F = c.F;
    const int dp = d + 4;                               // partial-slot pitch (keeps 16 B alignment)
    const uint32_t step = adam_step ? (uint32_t)adam_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int gidx = lane / LPR;
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (gidx * LPR));
    const int group = (threadIdx.x >> 5) * GPW + gidx;
    const int groups_per_block = (blockDim.x >> 5) * GPW;
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (MODE == VFMB_ADAM_TOUCHED) adam_coeffs(h, (int)step + 1, &step_size, &bc2_sqrt);

    for (int w = blockIdx.x * groups_per_block + group; w < W; w += gridDim.x * groups_per_block) {
        const int u = item_row[w];
        const int first = item_first[u], nit = item_first[u + 1] - first, ci = w - first;
        const int seg0 = seg_off[u], seg1 = seg_off[u + 1];
        const int s0 = seg0 + ci * kChunk;
        const int s1 = min(seg1, s0 + kChunk);
        Vec<VEC> acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[i].v[j] = 0.f;
        float gw = 0.f;

        for (int base = s0; base < s1; base += LPR) {
            const int idx = base + gl;
            const bool valid = idx < s1;
            int o = valid ? __ldg(occ + idx) : 0;
            int n = o / F, f = o - n * F;
            float r = valid ? __ldg(resid + n) : 0.f;
            int src = n;                                 // F>2: message row of the sample
            if (F == 2) src = valid ? __ldg(inverse + 2 * n + (1 - f)) : 0;   // partner's rank
            const int cnt = min(LPR, s1 - base);
            const float* table = (F == 2) ? vs : msg;
            for (int j = 0; j < cnt; ++j) {
                float rj = __shfl_sync(gmask, r, j, LPR);
                int sj = __shfl_sync(gmask, src, j, LPR);
                gw += rj;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> t = ld_vec_nc<VEC>(table + (size_t)sj * d + k);
#pragma unroll
                        for (int e = 0; e < VEC; ++e)
                            acc[i].v[e] = (F == 2) ? fmaf(rj, t.v[e], acc[i].v[e]) : acc[i].v[e] + t.v[e];
                    }
                }
            }
        }

        if (nit > 1) {
            // multi-chunk row: publish the partial, the last arriver combines all of them in
            // chunk order (fixed order => bitwise reproducible)
            const int slot = 2 * (s0 / kChunk) + (ci > 0 ? 1 : 0);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) st_vec<VEC>(gpart + (size_t)slot * dp + k, acc[i]);
            }
            if (gl == 0) gpart[(size_t)slot * dp + d] = gw;
            __threadfence();
            __syncwarp(gmask);
            int old = 0;
            if (gl == 0) old = atomicAdd(heavy_done + u, 1);
            old = __shfl_sync(gmask, old, 0, LPR);
            if (old != nit - 1) continue;
            __threadfence();
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[i].v[j] = 0.f;
            gw = 0.f;
            for (int cc = 0; cc < nit; ++cc) {
                const int sl = 2 * ((seg0 + cc * kChunk) / kChunk) + (cc > 0 ? 1 : 0);
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        const float* q = gpart + (size_t)sl * dp + k;
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[i].v[e] += __ldcg(q + e);
                    }
                }
                gw += __ldcg(gpart + (size_t)sl * dp + d);
            }
        }

        // ---- epilogue: chain rule to (mu, rho), KL gradient, Adam / gradient store
        const int rowid = uniq[u];
        const float q = (float)(seg1 - seg0) / __ldg(train_counts + rowid);
        const int cls = class_of(c, rowid);
        float csz = 0.f, zc = 1.f;
#pragma unroll
        for (int i = 0; i < kMaxFields; ++i) if (i == cls) { csz = c.class_size[i]; zc = __ldg(z + i); }
        const float cfac = kl_scale * q * (csz / zc);
        const size_t eoff = (size_t)rowid * 2 * d;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                Vec<VEC> e = entity_eps<VEC>(eps_entity, c, u, rowid, k, step);
                Vec<VEC> gmu, grho;
                if (F > 2) {
                    Vec<VEC> own = ld_vec_nc<VEC>(vs + (size_t)u * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) acc[i].v[j] -= gw * own.v[j];
                }
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    float sig = link_fn<LINK>(rho.v[j]);
                    gmu.v[j] = acc[i].v[j] + cfac * mu.v[j];
                    grho.v[j] = link_grad<LINK>(rho.v[j]) * (acc[i].v[j] * e.v[j] + cfac * (sig - 1.f / sig));
                }
                if (MODE == VFMB_ADAM_TOUCHED) {
                    Vec<VEC> m1 = ld_vec<VEC>(entity_m + eoff + k), m2 = ld_vec<VEC>(entity_m + eoff + d + k);
                    Vec<VEC> v1 = ld_vec<VEC>(entity_v + eoff + k), v2 = ld_vec<VEC>(entity_v + eoff + d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, bc2_sqrt);
                        adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, bc2_sqrt);
                    }
                    st_vec<VEC>(entity + eoff + k, mu);        st_vec<VEC>(entity + eoff + d + k, rho);
                    st_vec<VEC>(entity_m + eoff + k, m1);      st_vec<VEC>(entity_m + eoff + d + k, m2);
                    st_vec<VEC>(entity_v + eoff + k, v1);      st_vec<VEC>(entity_v + eoff + d + k, v2);
                } else {
                    st_vec<VEC>(grad_entity + eoff + k, gmu);  st_vec<VEC>(grad_entity + eoff + d + k, grho);
                }
            }
        }
        if (gl == 0) {
            const size_t boff = (size_t)rowid * 2;
            float a = bias[boff], b = bias[boff + 1];
            float tau = link_fn<LINK>(b);
            float eb = bias_eps(eps_bias, c, u, rowid, step);
            float ga = gw + cfac * a;
            float gb = link_grad<LINK>(b) * (gw * eb + cfac * (tau - 1.f / tau));
            if (MODE == VFMB_ADAM_TOUCHED) {
                float m1 = bias_m[boff], m2 = bias_m[boff + 1], v1 = bias_v[boff], v2 = bias_v[boff + 1];
                adam_elem(a, m1, v1, ga, h, step_size, bc2_sqrt);
                adam_elem(b, m2, v2, gb, h, step_size, bc2_sqrt);
                bias[boff] = a; bias[boff + 1] = b;
                bias_m[boff] = m1; bias_m[boff + 1] = m2;
                bias_v[boff] = v1; bias_v[boff + 1] = v2;
            } else {
                grad_bias[boff] = ga; grad_bias[boff + 1] = gb;
            }
        }
    }
}

// ------------------------------------------------------------------------------- k_final
template <int LINK, int LIK, int MODE>
__global__ void k_final(DevCfg c, float* __restrict__ scalars, float* __restrict__ sm,
                        float* __restrict__ sv, const float* __restrict__ stats,
                        const float* __restrict__ eps_global, AdamDev h, int32_t* __restrict__ adam_step,
                        float kl_scale, float* __restrict__ grad_scalars) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t step = (uint32_t)adam_step[0];
    float alpha = scalars[VFMB_S_ALPHA], mu0 = scalars[VFMB_S_GB_MEAN], rho0 = scalars[VFMB_S_GB_SCALE];
    const float sig0 = link_fn<LINK>(rho0), ap = link_fn<LINK>(alpha);
    const float e0 = global_eps(eps_global, c, step);
    const double sr = (double)stats[VFMB_ST_SUM_RESID], sq = (double)stats[VFMB_ST_SUM_SQERR];
    float g_mu0 = (float)(sr + (double)(kl_scale * mu0));
    float g_rho0 = link_grad<LINK>(rho0) * (float)((double)e0 * sr + (double)(kl_scale * (sig0 - 1.f / sig0)));
    float g_alpha = 0.f;
    if (LIK == VFMB_GAUSSIAN) {
        double sc = (double)c.n_train / ((double)c.S * (double)c.B);
        g_alpha = link_grad<LINK>(alpha) * (float)(sc * (0.5 * sq - 0.5 * (double)c.S * (double)c.B / (double)ap));
    }
    if (MODE == VFMB_ADAM_TOUCHED) {
        float ss, b2;
        adam_coeffs(h, (int)step + 1, &ss, &b2);
        adam_elem(mu0, sm[VFMB_S_GB_MEAN], sv[VFMB_S_GB_MEAN], g_mu0, h, ss, b2);
        adam_elem(rho0, sm[VFMB_S_GB_SCALE], sv[VFMB_S_GB_SCALE], g_rho0, h, ss, b2);
        scalars[VFMB_S_GB_MEAN] = mu0; scalars[VFMB_S_GB_SCALE] = rho0;
        if (LIK == VFMB_GAUSSIAN) {          // Bernoulli: alpha has no gradient, Adam skips it (N10)
            adam_elem(alpha, sm[VFMB_S_ALPHA], sv[VFMB_S_ALPHA], g_alpha, h, ss, b2);
            scalars[VFMB_S_ALPHA] = alpha;
        }
        adam_step[0] = (int32_t)step + 1;
    } else {
        grad_scalars[VFMB_S_ALPHA] = g_alpha;
        grad_scalars[VFMB_S_GB_MEAN] = g_mu0;
        grad_scalars[VFMB_S_GB_SCALE] = g_rho0;
    }
}

// ------------------------------------------------------------------------------- dense Adam
__global__ void __launch_bounds__(256)
k_adam_dense(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
             const float* __restrict__ g, int64_t n, AdamDev h, const int32_t* __restrict__ adam_step) {
    float ss, b2;
    adam_coeffs(h, adam_step[0] + 1, &ss, &b2);
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
    int nvec = (c.d + VEC - 1) / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)U * nvec;
         i += (int64_t)gridDim.x * blockDim.x) {
        int u = (int)(i / nvec), j = (int)(i % nvec);
        int rowid = uniq[u];
        Vec<VEC> e = entity_eps<VEC>(nullptr, c, u, rowid, j * VEC, step);
        for (int t = 0; t < VEC; ++t)
            if (j * VEC + t < c.d) eps_entity[(size_t)u * c.d + j * VEC + t] = e.v[t];
        if (j == 0) eps_bias[u] = bias_eps(nullptr, c, u, rowid, step);
        if (i == 0) eps_global[0] = global_eps(nullptr, c, step);
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

#define VFMB_LAYOUT_SWITCH(L, ...)                                                     \
    do {                                                                               \
        if (L.vec == 4 && L.lpr == 4 && L.nv == 1) { constexpr int VEC = 4, LPR = 4, NV = 1; __VA_ARGS__; }        \
        else if (L.vec == 4 && L.lpr == 8 && L.nv == 1) { constexpr int VEC = 4, LPR = 8, NV = 1; __VA_ARGS__; }   \
        else if (L.vec == 4 && L.lpr == 16 && L.nv == 1) { constexpr int VEC = 4, LPR = 16, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 4 && L.lpr == 32 && L.nv == 1) { constexpr int VEC = 4, LPR = 32, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 4 && L.lpr == 32 && L.nv == 2) { constexpr int VEC = 4, LPR = 32, NV = 2; __VA_ARGS__; } \
        else if (L.vec == 1 && L.lpr == 4 && L.nv == 1) { constexpr int VEC = 1, LPR = 4, NV = 1; __VA_ARGS__; }   \
        else if (L.vec == 1 && L.lpr == 8 && L.nv == 1) { constexpr int VEC = 1, LPR = 8, NV = 1; __VA_ARGS__; }   \
        else if (L.vec == 1 && L.lpr == 16 && L.nv == 1) { constexpr int VEC = 1, LPR = 16, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 1 && L.lpr == 32 && L.nv == 1) { constexpr int VEC = 1, LPR = 32, NV = 1; __VA_ARGS__; } \
        else if (L.vec == 1 && L.lpr == 32 && L.nv == 2) { constexpr int VEC = 1, LPR = 32, NV = 2; __VA_ARGS__; } \
        else return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);   \
    } while (0)

static int check_cfg(const vfmb_config* cfg, const char* who) {
    if (!cfg) return set_error(VFMB_EINVAL, "%s: null config", who);
    if (cfg->B <= 0 || cfg->R <= 0 || cfg->d <= 0) return set_error(VFMB_EINVAL, "%s: bad B/R/d", who);
    if (cfg->F < 2 || cfg->F > VFMB_MAX_FIELDS) return set_error(VFMB_ESHAPE, "%s: F must be 2..%d", who, VFMB_MAX_FIELDS);
    if (cfg->S != 1) return set_error(VFMB_ESHAPE, "%s: S=%d variational samples not supported yet (S=1)", who, cfg->S);
    if (cfg->n_classes < 1 || cfg->n_classes > cfg->F) return set_error(VFMB_EINVAL, "%s: bad n_classes", who);
    if (cfg->likelihood != VFMB_GAUSSIAN && cfg->likelihood != VFMB_BERNOULLI) return set_error(VFMB_EINVAL, "%s: bad likelihood", who);
    if (cfg->link != VFMB_LINK_ABS && cfg->link != VFMB_LINK_SOFTPLUS) return set_error(VFMB_EINVAL, "%s: bad link", who);
    return 0;
}

static int grid_for(int64_t work_groups, int groups_per_block) {
    int64_t g = (work_groups + groups_per_block - 1) / groups_per_block;
    if (g < 1) g = 1;
    if (g > kMaxGrid) g = kMaxGrid;
    return (int)g;
}

}  // namespace vfmb

using namespace vfmb;

extern "C" int64_t vfmb_partials_doubles(const vfmb_config* cfg) {
    if (!cfg) return 0;
    // block partials of the reductions, then the heavy-row partial slots of k_rows (floats),
    // then (closed form) per-block prior-gradient partials
    int64_t n = (int64_t)cfg->B * cfg->F;
    int64_t slots = 2 * (n / kChunk + 2);
    int64_t gpart_doubles = (slots * (cfg->d + 4) + 1) / 2;
    int64_t prior = (int64_t)kPriorGrid * 2 * cfg->F * (1 + cfg->d);
    return (int64_t)kMaxGrid * 16 + gpart_doubles + prior;
}

extern "C" int vfmb_sampled_forward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                    const vfmb_step_io* io, vfmb_stream stream_) {
    int rc = check_cfg(cfg, "vfmb_sampled_forward");
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_forward: null argument");
    if (cfg->F > 2 && io->y && !io->msg) return set_error(VFMB_EINVAL, "vfmb_sampled_forward: msg scratch required for F>2");
    cudaStream_t stream = (cudaStream_t)stream_;
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    DevCfg dc = make_dev(cfg);
    vfmb_plan_capacity_t cap;
    rc = vfmb_plan_capacity(cfg->B, cfg->F, cfg->R, &cap);
    if (rc) return rc;
    const int gpb = 8 * (32 / L.lpr);
    const int grid_u = grid_for(cap.u_cap, gpb), grid_b = grid_for(cfg->B, gpb);
#define LAUNCH_STAGE(LINK)                                                                             \
    k_stage<VEC, LPR, NV, LINK><<<grid_u, 256, 0, stream>>>(                                           \
        dc, tab->bias, tab->entity, tab->train_counts, plan->uniq, plan->seg_off, plan->meta, plan->z, \
        plan->heavy_done, io->eps_bias, io->eps_entity, tab->adam_step, io->vs, io->ws, io->partials,  \
        io->counters + 0, io->stats)
#define LAUNCH_SCORE(LINK, LIK)                                                                        \
    k_score<VEC, LPR, NV, LINK, LIK><<<grid_b, 256, 0, stream>>>(                                      \
        dc, tab->scalars, plan->inverse, io->vs, io->ws, io->y, io->eps_global, tab->adam_step,        \
        io->pred, io->mean, io->resid, io->msg, io->partials, io->counters + 1, io->stats)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            LAUNCH_STAGE(0);
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SCORE(0, VFMB_GAUSSIAN); else LAUNCH_SCORE(0, VFMB_BERNOULLI);
        } else {
            LAUNCH_STAGE(1);
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SCORE(1, VFMB_GAUSSIAN); else LAUNCH_SCORE(1, VFMB_BERNOULLI);
        }
    });
#undef LAUNCH_STAGE
#undef LAUNCH_SCORE
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_sampled_backward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                     const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                     float kl_grad_scale, vfmb_stream stream_) {
    int rc = check_cfg(cfg, "vfmb_sampled_backward");
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: null argument");
    if (mode == VFMB_ADAM_TOUCHED && (!adam || !tab->entity_m || !tab->entity_v || !tab->bias_m || !tab->bias_v ||
                                      !tab->scalars_m || !tab->scalars_v || !tab->adam_step))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: Adam state required");
    if (mode == VFMB_GRAD_ONLY && (!io->grad_bias || !io->grad_entity))
        return set_error(VFMB_EINVAL, "vfmb_sampled_backward: gradient outputs required");
    if (mode != VFMB_ADAM_TOUCHED && mode != VFMB_GRAD_ONLY) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: bad mode");
    if (cfg->F > 2 && !io->msg) return set_error(VFMB_EINVAL, "vfmb_sampled_backward: msg scratch required for F>2");
    cudaStream_t stream = (cudaStream_t)stream_;
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    DevCfg dc = make_dev(cfg);
    vfmb_plan_capacity_t cap;
    rc = vfmb_plan_capacity(cfg->B, cfg->F, cfg->R, &cap);
    if (rc) return rc;
    AdamDev h = make_adam(adam);
    const int gpb = 8 * (32 / L.lpr);
    const int grid_w = grid_for(cap.w_cap, gpb);
    float* gpart = (float*)(io->partials + (size_t)kMaxGrid * 16);   // heavy-row partial slots follow
#define LAUNCH_ROWS(LINK, MODE)                                                                          \
    k_rows<VEC, LPR, NV, LINK, MODE><<<grid_w, 256, 0, stream>>>(                                        \
        dc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v,              \
        tab->train_counts, plan->uniq, plan->inverse, plan->seg_off, plan->occ, plan->item_first,        \
        plan->item_row, plan->heavy_done, plan->z, plan->meta, io->eps_bias, io->eps_entity, io->vs,     \
        io->msg, io->resid, gpart, h, tab->adam_step, kl_grad_scale, io->grad_bias, io->grad_entity)
    VFMB_LAYOUT_SWITCH(L, {
        if (cfg->link == VFMB_LINK_ABS) {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_ROWS(0, VFMB_ADAM_TOUCHED); else LAUNCH_ROWS(0, VFMB_GRAD_ONLY);
        } else {
            if (mode == VFMB_ADAM_TOUCHED) LAUNCH_ROWS(1, VFMB_ADAM_TOUCHED); else LAUNCH_ROWS(1, VFMB_GRAD_ONLY);
        }
    });
#undef LAUNCH_ROWS
    CUDA_TRY(cudaGetLastError());
    if (mode == VFMB_GRAD_ONLY && !io->grad_scalars) return 0;
#define LAUNCH_FINAL(LINK, LIK, MODE)                                                                    \
    k_final<LINK, LIK, MODE><<<1, 32, 0, stream>>>(dc, tab->scalars, tab->scalars_m, tab->scalars_v,     \
                                                   io->stats, io->eps_global, h, tab->adam_step,         \
                                                   kl_grad_scale, io->grad_scalars)
    const int key = cfg->link * 4 + cfg->likelihood * 2 + (mode == VFMB_GRAD_ONLY ? 1 : 0);
    switch (key) {
        case 0: LAUNCH_FINAL(0, 0, VFMB_ADAM_TOUCHED); break;
        case 1: LAUNCH_FINAL(0, 0, VFMB_GRAD_ONLY); break;
        case 2: LAUNCH_FINAL(0, 1, VFMB_ADAM_TOUCHED); break;
        case 3: LAUNCH_FINAL(0, 1, VFMB_GRAD_ONLY); break;
        case 4: LAUNCH_FINAL(1, 0, VFMB_ADAM_TOUCHED); break;
        case 5: LAUNCH_FINAL(1, 0, VFMB_GRAD_ONLY); break;
        case 6: LAUNCH_FINAL(1, 1, VFMB_ADAM_TOUCHED); break;
        default: LAUNCH_FINAL(1, 1, VFMB_GRAD_ONLY); break;
    }
#undef LAUNCH_FINAL
    CUDA_TRY(cudaGetLastError());
    return 0;
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
    int64_t work = (int64_t)U * ((cfg->d + vec - 1) / vec);
    int grid = (int)((work + 255) / 256 > 4096 ? 4096 : (work + 255) / 256);
    if (vec == 4) k_philox_export<4><<<grid, 256, 0, (cudaStream_t)stream_>>>(dc, uniq, U, (uint32_t)step, eps_global, eps_bias, eps_entity);
    else k_philox_export<1><<<grid, 256, 0, (cudaStream_t)stream_>>>(dc, uniq, U, (uint32_t)step, eps_global, eps_bias, eps_entity);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
