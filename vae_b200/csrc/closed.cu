// Closed-form Gaussian VFM step (vfm-tomasrch.py:323-453 forward, :569-594 loss/backward/Adam) and
// the plan-free posterior-mean prediction (vfm-torch.py:248-259, vfm-tomasrch.py:342-348).
//
// Same skeleton as the sampled step (plan -> stage -> score -> gather -> Adam), no RNG:
//   k_cstage   per unique row: stage [mean | raw scale^2] and the bias pair in L2-resident scratch,
//              KL against the learnable per-group priors, per-group sums feeding the prior-
//              parameter gradients.  Blocks are assigned to (group, row range) from the plan's
//              class offsets, so a block only ever sees one group's prior.   (:262-313, 363-367)
//   k_cscore   per sample: y_bar, T_n, partial_loss, product-form prediction.        (:342-348, 369-449)
//   k_cgather  position-tiled segmented sums A = sum delta*mu_partner, Bq = sum rho_partner^2,
//              C = sum mu_partner^2 per unique row (deterministic: fixed order, no floating-point
//              atomics; rows cut by tile boundaries are finished in tile order by finish_cut_row).
//   k_cadam    chain rule + KL gradient + Adam on the touched rows.                       (:592-594)
//   k_cfinal   prior / scalar parameter gradients (fixed-order reductions) and their Adam step.
#include "step_common.cuh"

namespace vfmb {

struct ClosedCfg {
    int B, F, d, n_train_pad;
    int cls_bound[kMaxFields];
    float cls_size[kMaxFields];
    float n_train;
    float data_scale;      // backward: d loss / d(-partial_loss); N_train / B for the script's loss (:571)
    float kl0_scale;       // backward: weight of the global-bias KL term (1; 0 when the caller's autograd owns it)
    int off_pbm, off_pbs, off_pem, off_pes, n_scalars;   // offsets inside the scalar block
};

__device__ __forceinline__ int cclass_of(const ClosedCfg& c, int row) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < kMaxFields - 1; ++i) k += (i < c.F - 1 && row >= c.cls_bound[i]) ? 1 : 0;
    return k;
}

// ------------------------------------------------------------------------------- k_cstage
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256)
k_cstage(ClosedCfg c, const float* __restrict__ bias, const float* __restrict__ entity,
         const float* __restrict__ train_counts, const float* __restrict__ scalars,
         const int32_t* __restrict__ urec, const int32_t* __restrict__ class_off,
         const float* __restrict__ z, float* __restrict__ vs2, float* __restrict__ as,
         float* __restrict__ b2s, float* __restrict__ cq, float* __restrict__ klb_out,
         float* __restrict__ kle_out, double* __restrict__ partials, int32_t* __restrict__ counter,
         float* __restrict__ pg_part, int32_t* __restrict__ blk_class, float* __restrict__ stats,
         const float* __restrict__ cq_in) {
    // cq_in (optional, [U]): KL weight of every unique row supplied by the caller (the upstream gradient
    // of the per-row KL terms under autograd) instead of c_u = cnt_b/cnt_train * G_g/Z_g (:574-587)
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW, RPB = 8 * CH;
    extern __shared__ float s_pg[];                      // [8 warps][2d+4]
    const int d = c.d, pgw = 2 * d + 4;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR, warp = threadIdx.x >> 5;
    const unsigned gmask = group_mask<LPR>();
    // block -> (group g, chunk j) from the class offsets of the plan
    int g = -1, r0 = 0, r1 = 0;
    {
        int b = blockIdx.x;
        for (int q = 0; q < c.F; ++q) {
            const int lo = class_off[q], hi = class_off[q + 1];
            const int nb = (hi - lo + RPB - 1) / RPB;
            if (b < nb) { g = q; r0 = lo + b * RPB; r1 = min(hi, r0 + RPB); break; }
            b -= nb;
        }
    }
    if (threadIdx.x == 0) blk_class[blockIdx.x] = g;
    float facc = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
    Vec<VEC> e1[NV], e2[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < VEC; ++j) { e1[i].v[j] = 0.f; e2[i].v[j] = 0.f; }
    if (g >= 0) {
        const float pbm = scalars[c.off_pbm + g], pbs = fabsf(scalars[c.off_pbs + g]);
        const float csz_over_z = c.cls_size[g] / __ldg(z + g);
        Vec<VEC> pm[NV], ps[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d) {
                pm[i] = ld_vec<VEC>(scalars + c.off_pem + g * d + k);
                ps[i] = ld_vec<VEC>(scalars + c.off_pes + g * d + k);
#pragma unroll
                for (int j = 0; j < VEC; ++j) ps[i].v[j] = fabsf(ps[i].v[j]);
            }
        }
        const int base = r0 + warp * CH;
        // ---- lane-parallel: one unique row per lane
        const int ul = base + lane;
        const bool valid = lane < CH && ul < r1;
        int rowid_l = 0;
        float klb = 0.f, cqv = 0.f;
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x;
            prefetch_row(entity + (size_t)rowid_l * 2 * d, 8 * d);
            const float2 ab = *reinterpret_cast<const float2*>(bias + (size_t)rowid_l * 2);
            const float tcnt = __ldg(train_counts + rowid_l);
            const float tau = fabsf(ab.y);
            as[ul] = ab.x;
            b2s[ul] = ab.y * ab.y;
            klb = kl_normal(ab.x, tau, pbm, pbs);
            if (klb_out) klb_out[ul] = klb;
            cqv = cq_in ? __ldg(cq_in + ul) : ((float)rec.w / tcnt) * csz_over_z;
            cq[ul] = cqv;
            s0 += cqv;
            s1 = fmaf(cqv, ab.x, s1);
            s2 = fmaf(cqv, tau * tau + (ab.x - pbm) * (ab.x - pbm), s2);
        }
        float klrow = 0.f;
#pragma unroll 1
        for (int it = 0; it < kRounds; ++it) {
            const int sel = it * GPW + gidx;
            const int rowid = bcast(rowid_l, sel);
            const float cu = bcast(cqv, sel);
            const int u = base + sel;
            float kl = 0.f;
            if (u < r1) {
                const float* erow = entity + (size_t)rowid * 2 * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> mu = ld_vec<VEC>(erow + k), rho = ld_vec<VEC>(erow + d + k), r2, klv;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            const float sig = fabsf(rho.v[j]);
                            r2.v[j] = rho.v[j] * rho.v[j];
                            klv.v[j] = kl_normal(mu.v[j], sig, pm[i].v[j], ps[i].v[j]);
                            kl += klv.v[j];
                            const float dm = mu.v[j] - pm[i].v[j];
                            e1[i].v[j] = fmaf(cu, mu.v[j], e1[i].v[j]);
                            e2[i].v[j] = fmaf(cu, r2.v[j] + dm * dm, e2[i].v[j]);
                        }
                        st_vec<VEC>(vs2 + (size_t)u * 2 * d + k, mu);
                        st_vec<VEC>(vs2 + (size_t)u * 2 * d + d + k, r2);
                        if (kle_out) st_vec<VEC>(kle_out + (size_t)u * d + k, klv);
                    }
                }
                kl = group_sum<LPR>(kl, gmask);
            }
            hand_back<LPR>(klrow, kl, it, lane);
        }
        if (valid) facc = fmaf(cqv, klrow + klb, facc);
    }
    // ---- block reduction of the prior-gradient sums, fixed order: groups of a warp, then warps
    // (lanes of different groups hold the same k-slice)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int k = (gl + i * LPR) * VEC;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int q = 0; q < GPW; ++q) {
                a1 += __shfl_sync(0xffffffffu, e1[i].v[j], q * LPR + gl);
                a2 += __shfl_sync(0xffffffffu, e2[i].v[j], q * LPR + gl);
            }
            if (gidx == 0 && k < d) { s_pg[warp * pgw + k + j] = a1; s_pg[warp * pgw + d + k + j] = a2; }
        }
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { s_pg[warp * pgw + 2 * d] = s0; s_pg[warp * pgw + 2 * d + 1] = s1; s_pg[warp * pgw + 2 * d + 2] = s2; }
    __syncthreads();
    for (int j = threadIdx.x; j < pgw - 1; j += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_pg[w * pgw + j];
        pg_part[(size_t)blockIdx.x * pgw + j] = t;
    }
    double acc[1] = {(double)facc};
    if (block_partials<1>(acc, partials, counter)) {
        double tot[1];
        final_sums<1>(partials, tot);
        if (threadIdx.x == 0) { stats[VFMB_ST_KL_ROWS] = (float)tot[0]; *counter = 0; }
    }
}

// ------------------------------------------------------------------------------- k_cscore
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256)
k_cscore(ClosedCfg c, const float* __restrict__ scalars, const int32_t* __restrict__ inverse,
         const int32_t* __restrict__ pos_of, const float* __restrict__ vs2, const float* __restrict__ as,
         const float* __restrict__ b2s, const float* __restrict__ y, float* __restrict__ pred,
         float* __restrict__ resid, float* __restrict__ rsorted, float* __restrict__ msg,
         double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW;
    const int d = c.d, F = c.F, B = c.B;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const unsigned gmask = group_mask<LPR>();
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float mu0 = scalars[VFMB_C_GB_MEAN], rho0 = scalars[VFMB_C_GB_SCALE];
    const float alpha = fabsf(scalars[VFMB_C_ALPHA]);
    const float half_log_alpha = 0.5f * logf(alpha);
    double acc[3] = {0.0, 0.0, 0.0};     // partial_loss, delta, delta^2 + T

    for (int base = gwarp * CH; base < B; base += nwarps * CH) {
        const int nl = base + lane;
        const bool valid = lane < CH && nl < B;
        float asum = 0.f, b2sum = 0.f, yn = 0.f;
        int2 rr = make_int2(0, 0);
        if (valid) {
            if (F == 2) {
                rr = __ldg(reinterpret_cast<const int2*>(inverse) + nl);
                asum = __ldg(as + rr.x) + __ldg(as + rr.y);
                b2sum = __ldg(b2s + rr.x) + __ldg(b2s + rr.y);
            } else {
                for (int f = 0; f < F; ++f) {
                    const int r = __ldg(inverse + (size_t)nl * F + f);
                    asum += __ldg(as + r); b2sum += __ldg(b2s + r);
                }
            }
            if (y) yn = __ldg(y + nl);
        }
        float dot_l = 0.f, t_l = 0.f, prod_l = 0.f;
#pragma unroll 1
        for (int it = 0; it < kRounds; ++it) {
            const int sel = it * GPW + gidx;
            const int n = base + sel;
            float pdot = 0.f, pt = 0.f, pprod = 0.f;
            const int r0 = bcast(rr.x, sel), r1 = bcast(rr.y, sel);
            if (n < B) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    int k = (gl + i * LPR) * VEC;
                    if (k < d) {
                        Vec<VEC> sm, sm2, sr2, sz2, sm4, pr;
#pragma unroll
                        for (int j = 0; j < VEC; ++j) { sm.v[j] = sm2.v[j] = sr2.v[j] = sz2.v[j] = sm4.v[j] = 0.f; pr.v[j] = 1.f; }
                        for (int f = 0; f < F; ++f) {
                            const int r = (F == 2) ? (f == 0 ? r0 : r1) : __ldg(inverse + (size_t)n * F + f);
                            const Vec<VEC> m = ld_vec_nc<VEC>(vs2 + (size_t)r * 2 * d + k);
                            const Vec<VEC> q = ld_vec_nc<VEC>(vs2 + (size_t)r * 2 * d + d + k);
#pragma unroll
                            for (int j = 0; j < VEC; ++j) {
                                const float m2 = m.v[j] * m.v[j], zz = m2 + q.v[j];
                                sm.v[j] += m.v[j]; sm2.v[j] += m2; sr2.v[j] += q.v[j];
                                sz2.v[j] = fmaf(zz, zz, sz2.v[j]); sm4.v[j] = fmaf(m2, m2, sm4.v[j]);
                                pr.v[j] *= m.v[j];
                            }
                        }
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            // e2(mu) ; e2(mu^2 + rho^2) - e2(mu^2)   
                            pdot += 0.5f * (sm.v[j] * sm.v[j] - sm2.v[j]);
                            const float zs = sm2.v[j] + sr2.v[j];
                            pt += 0.5f * (zs * zs - sz2.v[j]) - 0.5f * (sm2.v[j] * sm2.v[j] - sm4.v[j]);
                            pprod += pr.v[j];
                        }
                        if (F > 2 && msg) {                 // partner sums for the backward: [S_mu | S_rho2 | S_mu2]
                            st_vec<VEC>(msg + (size_t)n * 3 * d + k, sm);
                            st_vec<VEC>(msg + (size_t)n * 3 * d + d + k, sr2);
                            st_vec<VEC>(msg + (size_t)n * 3 * d + 2 * d + k, sm2);
                        }
                    }
                }
            }
            pdot = group_sum<LPR>(pdot, gmask);
            pt = group_sum<LPR>(pt, gmask);
            pprod = group_sum<LPR>(pprod, gmask);
            hand_back<LPR>(dot_l, pdot, it, lane);
            hand_back<LPR>(t_l, pt, it, lane);
            hand_back<LPR>(prod_l, pprod, it, lane);
        }
        if (valid) {
            const float lin = mu0 + asum;
            pred[nl] = lin + prod_l;                        // product over groups (vfm-tomasrch.py:342-348)
            if (y) {
                const float ybar = lin + dot_l;
                const float tn = rho0 * rho0 + b2sum + t_l;
                const float delta = yn - ybar;
                acc[0] += (double)(half_log_alpha - 0.5f * alpha * (delta * delta + tn));
                acc[1] += (double)delta;
                acc[2] += (double)delta * (double)delta + (double)tn;
                resid[nl] = delta;
                if (F == 2) {
                    const int2 pp = __ldg(reinterpret_cast<const int2*>(pos_of) + nl);
                    rsorted[pp.x] = delta; rsorted[pp.y] = delta;
                } else {
                    for (int f = 0; f < F; ++f) rsorted[__ldg(pos_of + (size_t)nl * F + f)] = delta;
                }
            }
        }
    }
    if (block_partials<3>(acc, partials, counter)) {
        double tot[3];
        final_sums<3>(partials, tot);
        if (threadIdx.x == 0) {
            const float m0 = scalars[VFMB_C_GB_PRIOR_MEAN], s0 = fabsf(scalars[VFMB_C_GB_PRIOR_SCALE]);
            const float kl0 = kl_normal(mu0, fabsf(rho0), m0, s0);
            const float kl = kl0 + stats[VFMB_ST_KL_ROWS];
            stats[VFMB_ST_NLL_MEAN] = (float)tot[0];       // partial_loss (vfm-tomasrch.py:445-449)
            stats[VFMB_ST_SUM_RESID] = (float)tot[1];
            stats[VFMB_ST_SUM_SQERR] = (float)tot[2];
            stats[VFMB_ST_KL] = kl;
            stats[VFMB_ST_LOSS] = (float)(-(double)c.n_train * tot[0] / (double)B + (double)kl);
            stats[VFMB_ST_W0] = mu0;
            *counter = 0;
        }
    }
}

// ------------------------------------------------------------------------------- k_cgather
// Same position-tiled segmented reduction as the sampled k_gather, three sums per unique row:
// A = sum delta_n * mu_src, Bq = sum rho_src^2, C = sum mu_src^2, where src is the partner row
// (F == 2, [mu | rho^2] in vs2) or the sample's field sums (F > 2, [S_mu | S_rho2 | S_mu2] in
// msg; the row's own terms are removed in k_cadam).  Output rows are [A | Bq | C], 3d wide.
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256)
k_cgather(int d, int F, int N, const int32_t* __restrict__ partner, const int32_t* __restrict__ pos_rank,
          const int32_t* __restrict__ urec, const float* __restrict__ vs2, const float* __restrict__ msg,
          const float* __restrict__ rsorted, float* gslot, float* __restrict__ grow, float* __restrict__ gws,
          int32_t* arrive, GatherKnobs kn) {
    constexpr int GPW = kWarp / LPR;
    const int dp = 3 * d + 4;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const unsigned gmask = group_mask<LPR>();
    const int group = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + lane / LPR;
    const int ngroups = gridDim.x * (blockDim.x >> 5) * GPW;
    const int n_tiles = (N + kTile - 1) / kTile;
    const float* table = (F == 2) ? vs2 : msg;
    const int pitch = (F == 2) ? 2 * d : 3 * d;

    for (int tile = group; tile < n_tiles; tile += ngroups) {
        const TileSpan ts = tile_span(tile, N, kn.keep, pos_rank, urec);    // short rows kept whole (step_common.cuh)
        const int t0 = ts.t0, t1 = ts.t1;
        if (t0 >= t1) continue;
        int cur = ts.head_u;                                // else set from the first position
        Vec<VEC> aA[NV], aB[NV], aC[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) { aA[i].v[j] = 0.f; aB[i].v[j] = 0.f; aC[i].v[j] = 0.f; }
        float gd = 0.f;

        auto flush = [&](int u) {
            const bool open_h = u == ts.head_u, open_t = u == ts.tail_u;
            float* dst; float* dstw;
            if (open_h)      { dst = gslot + ((size_t)tile * 2) * dp;     dstw = dst + 3 * d; }
            else if (open_t) { dst = gslot + ((size_t)tile * 2 + 1) * dp; dstw = dst + 3 * d; }
            else             { dst = grow + (size_t)u * 3 * d;            dstw = gws + u; }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) {
                    st_vec<VEC>(dst + k, aA[i]); st_vec<VEC>(dst + d + k, aB[i]); st_vec<VEC>(dst + 2 * d + k, aC[i]);
                }
            }
            if (gl == 0) *dstw = gd;
        };

        for (int b0 = t0; b0 < t1; b0 += LPR) {
            const int idx = b0 + gl;
            const bool ok = idx < t1;
            const float r = ok ? __ldg(rsorted + idx) : 0.f;
            const int src = ok ? __ldg(partner + idx) : 0;
            const int ur = ok ? __ldg(pos_rank + idx) : 0;
            const int cnt = min(LPR, t1 - b0);
            const int src0 = __shfl_sync(gmask, src, 0, LPR);
            if (b0 == t0) cur = __shfl_sync(gmask, ur, 0, LPR);
            for (int j = 0; j < cnt; j += 2) {                 // 2 x (2 or 3) row slices in flight
                float rj[2]; int uj[2]; Vec<VEC> tm[2][NV], tq[2][NV], tc[2][NV];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    rj[e] = __shfl_sync(gmask, r, (j + e) & (LPR - 1), LPR);
                    uj[e] = __shfl_sync(gmask, ur, (j + e) & (LPR - 1), LPR);
                    int sj = __shfl_sync(gmask, src, (j + e) & (LPR - 1), LPR);
                    if (j + e >= cnt) sj = src0;
                    const float* row = table + (size_t)sj * pitch;
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        int k = (gl + i * LPR) * VEC;
                        if (k < d) {
                            tm[e][i] = ld_vec_nc<VEC>(row + k);
                            tq[e][i] = ld_vec_nc<VEC>(row + d + k);
                            if (F > 2) tc[e][i] = ld_vec_nc<VEC>(row + 2 * d + k);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (j + e < cnt) {
                        if (uj[e] != cur) {
                            flush(cur);
                            cur = uj[e];
#pragma unroll
                            for (int i = 0; i < NV; ++i)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) { aA[i].v[q] = 0.f; aB[i].v[q] = 0.f; aC[i].v[q] = 0.f; }
                            gd = 0.f;
                        }
                        gd += rj[e];
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            int k = (gl + i * LPR) * VEC;
                            if (k < d)
#pragma unroll
                                for (int q = 0; q < VEC; ++q) {
                                    const float m = tm[e][i].v[q];
                                    aA[i].v[q] = fmaf(rj[e], m, aA[i].v[q]);
                                    aB[i].v[q] += tq[e][i].v[q];
                                    aC[i].v[q] = (F > 2) ? aC[i].v[q] + tc[e][i].v[q] : fmaf(m, m, aC[i].v[q]);
                                }
                        }
                    }
                }
            }
        }
        flush(cur);
        // the group storing the last partial of a cut row adds them in tile order (step_common.cuh)
        if (ts.head_u >= 0)
            finish_cut_row<VEC, LPR, NV, 3>(ts.head_u, tile, d, nullptr, urec, gslot, grow + (size_t)ts.head_u * 3 * d,
                                            gws + ts.head_u, arrive, n_tiles + 1, kn.light_fence);
        if (ts.tail_u >= 0 && ts.tail_u != ts.head_u)
            finish_cut_row<VEC, LPR, NV, 3>(ts.tail_u, tile, d, nullptr, urec, gslot, grow + (size_t)ts.tail_u * 3 * d,
                                            gws + ts.tail_u, arrive, n_tiles + 1, kn.light_fence);
    }
}

// ------------------------------------------------------------------------------- k_cadam
template <int VEC, int LPR, int NV, int MODE>
__global__ void __launch_bounds__(256)
k_cadam(ClosedCfg c, float* __restrict__ bias, float* __restrict__ bias_m, float* __restrict__ bias_v,
        float* __restrict__ entity, float* __restrict__ entity_m, float* __restrict__ entity_v,
        const float* __restrict__ scalars, const int32_t* __restrict__ urec, const int32_t* __restrict__ meta,
        const float* __restrict__ cq, const float* __restrict__ grow, const float* __restrict__ gws,
        AdamDev h, const int32_t* __restrict__ adam_step, float* __restrict__ grad_bias,
        float* __restrict__ grad_entity) {
    constexpr int GPW = kWarp / LPR, CH = kRounds * GPW;
    const int U = meta[0];
    const int d = c.d, F = c.F;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    __shared__ float s_coef[2];
    if (MODE == VFMB_ADAM_TOUCHED) {
        if (threadIdx.x == 0) adam_coeffs(h, adam_step[0] + 1, &s_coef[0], &s_coef[1]);
        __syncthreads();
    }
    const float step_size = (MODE == VFMB_ADAM_TOUCHED) ? s_coef[0] : 0.f;
    const float inv_bc2 = (MODE == VFMB_ADAM_TOUCHED) ? s_coef[1] : 1.f;
    const float kappa = c.data_scale * fabsf(scalars[VFMB_C_ALPHA]);

    for (int base = gwarp * CH; base < U; base += nwarps * CH) {
        const int ul = base + lane;
        const bool valid = lane < CH && ul < U;
        int rowid_l = 0, cls_l = 0, cnt_l = 0;
        float cu_l = 0.f, gd_l = 0.f;
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x; cnt_l = rec.y;
            const size_t eoff = (size_t)rowid_l * 2 * d;
            prefetch_row(entity + eoff, 8 * d);
            if (MODE == VFMB_ADAM_TOUCHED) { prefetch_row(entity_m + eoff, 8 * d); prefetch_row(entity_v + eoff, 8 * d); }
            cls_l = cclass_of(c, rowid_l);
            cu_l = __ldg(cq + ul);
            gd_l = __ldg(gws + ul);
            // bias pair: d/da = -kappa*sum(delta) + c(a-m)/s^2 ; d/db = kappa*b*cnt + sign(b) c (tau/s^2 - 1/tau)
            const float pbm = scalars[c.off_pbm + cls_l], pbs = fabsf(scalars[c.off_pbs + cls_l]);
            const float is2 = fast_rcp(pbs * pbs);
            const size_t boff = (size_t)rowid_l * 2;
            float2 ab = *reinterpret_cast<const float2*>(bias + boff);
            const float tau = fabsf(ab.y);
            const float sgn = ab.y > 0.f ? 1.f : (ab.y < 0.f ? -1.f : 0.f);
            const float ga = fmaf(cu_l * is2, ab.x - pbm, -kappa * gd_l);
            const float gb = fmaf(kappa * (float)cnt_l, ab.y, sgn * cu_l * (tau * is2 - fast_rcp(tau)));
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
#pragma unroll 1
        for (int it = 0; it < kRounds; ++it) {
            const int sel = it * GPW + gidx;
            const int rowid = bcast(rowid_l, sel), cls = bcast(cls_l, sel);
            const float cu = bcast(cu_l, sel), gd = bcast(gd_l, sel), fcnt = (float)bcast(cnt_l, sel);
            const int u = base + sel;
            if (u >= U) continue;
            const size_t eoff = (size_t)rowid * 2 * d;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) {
                    Vec<VEC> mu = ld_vec<VEC>(entity + eoff + k), rho = ld_vec<VEC>(entity + eoff + d + k);
                    Vec<VEC> m1, m2, v1, v2;
                    if (MODE == VFMB_ADAM_TOUCHED) {
                        m1 = ld_vec<VEC>(entity_m + eoff + k); m2 = ld_vec<VEC>(entity_m + eoff + d + k);
                        v1 = ld_vec<VEC>(entity_v + eoff + k); v2 = ld_vec<VEC>(entity_v + eoff + d + k);
                    }
                    const Vec<VEC> A = ld_vec_nc<VEC>(grow + (size_t)u * 3 * d + k);
                    const Vec<VEC> Bq = ld_vec_nc<VEC>(grow + (size_t)u * 3 * d + d + k);
                    const Vec<VEC> Cq = ld_vec_nc<VEC>(grow + (size_t)u * 3 * d + 2 * d + k);
                    const Vec<VEC> pm = ld_vec<VEC>(scalars + c.off_pem + cls * d + k);
                    const Vec<VEC> ps = ld_vec<VEC>(scalars + c.off_pes + cls * d + k);
                    Vec<VEC> gmu, grho;
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        float a = A.v[j], b = Bq.v[j], cc = Cq.v[j];
                        const float m = mu.v[j], r = rho.v[j], r2 = r * r;
                        if (F > 2) { a -= gd * m; b -= fcnt * r2; cc -= fcnt * m * m; }   // remove own terms
                        const float sig = fabsf(r), s = fabsf(ps.v[j]);
                        const float is2 = fast_rcp(s * s);
                        const float sgn = r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f);
                        gmu.v[j] = fmaf(kappa, fmaf(m, b, -a), cu * (m - pm.v[j]) * is2);
                        grho.v[j] = fmaf(kappa * r, cc + b, sgn * cu * (sig * is2 - fast_rcp(sig)));
                    }
                    if (MODE == VFMB_ADAM_TOUCHED) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            adam_elem(mu.v[j], m1.v[j], v1.v[j], gmu.v[j], h, step_size, inv_bc2);
                            adam_elem(rho.v[j], m2.v[j], v2.v[j], grho.v[j], h, step_size, inv_bc2);
                        }
                        st_vec<VEC>(entity + eoff + k, mu);        st_vec<VEC>(entity + eoff + d + k, rho);
                        st_vec<VEC>(entity_m + eoff + k, m1);      st_vec<VEC>(entity_m + eoff + d + k, m2);
                        st_vec<VEC>(entity_v + eoff + k, v1);      st_vec<VEC>(entity_v + eoff + d + k, v2);
                    } else {
                        st_vec<VEC>(grad_entity + eoff + k, gmu);  st_vec<VEC>(grad_entity + eoff + d + k, grho);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------- k_cfinal
// One warp per (group, column) of the prior-gradient sums: lanes stride over the blocks of
// k_cstage that worked on that group (contiguous block range), fixed shuffle tree.  The last
// block to finish turns the sums into gradients of every scalar parameter and applies Adam.
template <int MODE>
__global__ void __launch_bounds__(256)
k_cfinal(ClosedCfg c, float* __restrict__ scalars, float* __restrict__ sm, float* __restrict__ sv,
         const float* __restrict__ stats, const float* __restrict__ pg_part,
         const int32_t* __restrict__ blk_class, int nblk, float* __restrict__ pg_sum,
         int32_t* __restrict__ counter, AdamDev h, int32_t* __restrict__ adam_step,
         float* __restrict__ grad_scalars) {
    const int d = c.d, G = c.F, pgw = 2 * d + 4;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    __shared__ bool s_last;
    for (int item = gwarp; item < G * (pgw - 1); item += nwarps) {
        const int g = item / (pgw - 1), j = item % (pgw - 1);
        float t = 0.f;
        for (int b = lane; b < nblk; b += 32)
            if (__ldg(blk_class + b) == g) t += __ldg(pg_part + (size_t)b * pgw + j);
        t = warp_sum(t);
        if (lane == 0) pg_sum[g * pgw + j] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float ss = 0.f, ib2 = 1.f;
    __shared__ float s_coef[2];
    if (MODE == VFMB_ADAM_TOUCHED) {
        if (threadIdx.x == 0) adam_coeffs(h, adam_step[0] + 1, &s_coef[0], &s_coef[1]);
        __syncthreads();
        ss = s_coef[0]; ib2 = s_coef[1];
    }
    const float alpha = scalars[VFMB_C_ALPHA], mu0 = scalars[VFMB_C_GB_MEAN], rho0 = scalars[VFMB_C_GB_SCALE];
    const float m0 = scalars[VFMB_C_GB_PRIOR_MEAN], s0r = scalars[VFMB_C_GB_PRIOR_SCALE];
    const float ap = fabsf(alpha), as0 = fabsf(s0r);
    const double nb = (double)c.data_scale;
    const float kappa = (float)(nb * (double)ap);
    const float k0 = c.kl0_scale;
    const double sum_d = (double)stats[VFMB_ST_SUM_RESID], sum_q = (double)stats[VFMB_ST_SUM_SQERR];
    auto sgn = [](float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); };
    // every scalar parameter gets its gradient from one thread (block-stride loop)
    for (int idx = threadIdx.x; idx < c.n_scalars; idx += blockDim.x) {
        float g = 0.f;
        bool live = true;
        if (idx == VFMB_C_ALPHA) {
            g = -sgn(alpha) * (float)(nb * (0.5 * (double)c.B / (double)ap - 0.5 * sum_q));
        } else if (idx == VFMB_C_GB_MEAN) {
            g = (float)(-(double)kappa * sum_d) + k0 * (mu0 - m0) / (as0 * as0);
        } else if (idx == VFMB_C_GB_SCALE) {
            g = kappa * (float)c.B * rho0 + k0 * sgn(rho0) * (fabsf(rho0) / (as0 * as0) - 1.f / fabsf(rho0));
        } else if (idx == VFMB_C_GB_PRIOR_MEAN) {
            g = k0 * (m0 - mu0) / (as0 * as0);
        } else if (idx == VFMB_C_GB_PRIOR_SCALE) {
            g = k0 * sgn(s0r) * (1.f / as0 - (rho0 * rho0 + (mu0 - m0) * (mu0 - m0)) / (as0 * as0 * as0));
        } else if (idx >= c.off_pbm && idx < c.off_pbm + G) {
            const int q = idx - c.off_pbm;
            const float m = scalars[idx], s = fabsf(scalars[c.off_pbs + q]);
            const float S0 = __ldcg(pg_sum + q * pgw + 2 * d), S1 = __ldcg(pg_sum + q * pgw + 2 * d + 1);
            g = (m * S0 - S1) / (s * s);
        } else if (idx >= c.off_pbs && idx < c.off_pbs + G) {
            const int q = idx - c.off_pbs;
            const float s = fabsf(scalars[idx]);
            const float S0 = __ldcg(pg_sum + q * pgw + 2 * d), S2 = __ldcg(pg_sum + q * pgw + 2 * d + 2);
            g = sgn(scalars[idx]) * (S0 / s - S2 / (s * s * s));
        } else if (idx >= c.off_pem && idx < c.off_pem + G * d) {
            const int q = (idx - c.off_pem) / d, k = (idx - c.off_pem) % d;
            const float m = scalars[idx], s = fabsf(scalars[c.off_pes + q * d + k]);
            const float S0 = __ldcg(pg_sum + q * pgw + 2 * d), E1 = __ldcg(pg_sum + q * pgw + k);
            g = (m * S0 - E1) / (s * s);
        } else if (idx >= c.off_pes && idx < c.off_pes + G * d) {
            const int q = (idx - c.off_pes) / d, k = (idx - c.off_pes) % d;
            const float s = fabsf(scalars[idx]);
            const float S0 = __ldcg(pg_sum + q * pgw + 2 * d), E2 = __ldcg(pg_sum + q * pgw + d + k);
            g = sgn(scalars[idx]) * (S0 / s - E2 / (s * s * s));
        } else {
            live = false;                                   // padding
        }
        grad_scalars[idx] = live ? g : 0.f;
    }
    __syncthreads();                                        // every gradient is formed from the OLD values
    if (MODE == VFMB_ADAM_TOUCHED) {
        for (int idx = threadIdx.x; idx < c.n_scalars; idx += blockDim.x) {
            const bool live = idx <= VFMB_C_GB_PRIOR_SCALE || idx >= c.off_pbm;
            const bool pad = idx >= c.off_pbs + G && idx < c.off_pem;
            if (!live || pad) continue;
            float p = scalars[idx], m = sm[idx], v = sv[idx];
            adam_elem(p, m, v, grad_scalars[idx], h, ss, ib2);
            scalars[idx] = p; sm[idx] = m; sv[idx] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *counter = 0;
        if (MODE == VFMB_ADAM_TOUCHED) adam_step[0] += 1;
    }
}

// one lane group per sample; rows are read straight from the parameter table (means only)
template <int VEC, int LPR, int NV>
__global__ void __launch_bounds__(256)
k_predict_mean(int B, int F, int d, int R, int pairwise, const float* __restrict__ bias, const float* __restrict__ entity,
               float global_bias, const int64_t* __restrict__ x, float* __restrict__ out) {
    constexpr int GPW = kWarp / LPR;
    const int lane = threadIdx.x & 31, gl = lane % LPR;
    const int group = (threadIdx.x >> 5) * GPW + lane / LPR;
    const int groups_per_block = (blockDim.x >> 5) * GPW;
    for (int n = blockIdx.x * groups_per_block + group; n < B; n += gridDim.x * groups_per_block) {
        Vec<VEC> prod[NV], ssum[NV], sq[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < VEC; ++j) { prod[i].v[j] = 1.f; ssum[i].v[j] = 0.f; sq[i].v[j] = 0.f; }
        float bsum = 0.f;
        for (int f = 0; f < F; ++f) {
            int64_t id = x[(size_t)n * F + f];
            if (id < 0 || id >= R) id = 0;
            bsum += __ldg(bias + id * 2);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                int k = (gl + i * LPR) * VEC;
                if (k < d) {
                    Vec<VEC> a = ld_vec_nc<VEC>(entity + id * 2 * d + k);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) {
                        prod[i].v[j] *= a.v[j];
                        ssum[i].v[j] += a.v[j];
                        sq[i].v[j] = fmaf(a.v[j], a.v[j], sq[i].v[j]);
                    }
                }
            }
        }
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            int k = (gl + i * LPR) * VEC;
            if (k < d)
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    part += pairwise ? 0.5f * (ssum[i].v[j] * ssum[i].v[j] - sq[i].v[j]) : prod[i].v[j];
        }
        float inter = group_sum<LPR>(part, group_mask<LPR>());
        if (gl == 0) out[n] = global_bias + bsum + inter;
    }
}

}  // namespace vfmb

using namespace vfmb;

#define VFMB_LAYOUT_SWITCH3(L, ...)                                                    \
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

extern "C" int vfmb_predict_mean(const vfmb_config* cfg, const float* bias, const float* entity,
                                 float global_bias, const int64_t* x, float* out, vfmb_stream stream_) {
    if (!cfg || !bias || !entity || !x || !out) return set_error(VFMB_EINVAL, "vfmb_predict_mean: null argument");
    if (cfg->B <= 0 || cfg->F < 1 || cfg->F > VFMB_MAX_FIELDS) return set_error(VFMB_EINVAL, "vfmb_predict_mean: bad B/F");
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    const int gpb = 8 * (32 / L.lpr);
    int64_t g = ((int64_t)cfg->B + gpb - 1) / gpb;
    if (g > kMaxGrid) g = kMaxGrid;
    VFMB_LAYOUT_SWITCH3(L, {
        k_predict_mean<VEC, LPR, NV><<<(int)g, 256, 0, counted((cudaStream_t)stream_)>>>(
            cfg->B, cfg->F, cfg->d, cfg->R, (cfg->interaction == VFMB_INTER_PAIRWISE && cfg->F != 2) ? 1 : 0, bias, entity,
            global_bias, x, out);
    });
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int closed_check(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                        const vfmb_step_io* io, const char* who) {
    if (!cfg || !tab || !plan || !io) return set_error(VFMB_EINVAL, "%s: null argument", who);
    if (cfg->B <= 0 || cfg->R <= 0 || cfg->d <= 0) return set_error(VFMB_EINVAL, "%s: bad B/R/d", who);
    if (cfg->F < 2 || cfg->F > VFMB_MAX_FIELDS) return set_error(VFMB_ESHAPE, "%s: F must be 2..%d", who, VFMB_MAX_FIELDS);
    if (cfg->n_classes != cfg->F) return set_error(VFMB_EINVAL, "%s: closed form needs one KL class per group", who);
    if (cfg->likelihood != VFMB_GAUSSIAN) return set_error(VFMB_ESHAPE, "%s: the closed form is Gaussian only", who);
    if (!io->vs || !io->ws || !io->ebs || !io->cq || !io->grow || !io->gws || !io->rsorted || !io->partials)
        return set_error(VFMB_EINVAL, "%s: scratch required", who);
    if (cfg->F > 2 && !io->msg) return set_error(VFMB_EINVAL, "%s: msg scratch required for F>2", who);
    return 0;
}

static ClosedCfg make_closed(const vfmb_config* cfg) {
    ClosedCfg c{};
    c.B = cfg->B; c.F = cfg->F; c.d = cfg->d; c.n_train = cfg->n_train;
    c.data_scale = cfg->n_train / (float)cfg->B; c.kl0_scale = 1.f;
    for (int i = 0; i < kMaxFields; ++i) { c.cls_bound[i] = cfg->class_bound[i]; c.cls_size[i] = cfg->class_size[i]; }
    c.off_pbm = vfmb_closed_off_bias_prior_mean(cfg->F, cfg->d, 0);
    c.off_pbs = vfmb_closed_off_bias_prior_scale(cfg->F, cfg->d, 0);
    c.off_pem = vfmb_closed_off_entity_prior_mean(cfg->F, cfg->d, 0);
    c.off_pes = vfmb_closed_off_entity_prior_scale(cfg->F, cfg->d, 0);
    c.n_scalars = vfmb_closed_scalar_count(cfg->F, cfg->d);
    return c;
}

extern "C" int vfmb_closed_forward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                   const vfmb_step_io* io, vfmb_stream stream_) {
    int rc = closed_check(cfg, tab, plan, io, "vfmb_closed_forward");
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    vfmb_plan_capacity_t cap;
    rc = vfmb_plan_capacity(cfg->B, cfg->F, cfg->R, &cap);
    if (rc) return rc;
    ClosedCfg cc = make_closed(cfg);
    const ScratchMap sm = scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap);
    float* fbase = (float*)io->partials;
    float* pg_part = fbase + sm.pg_part_off;
    int32_t* blk_class = (int32_t*)(fbase + sm.blk_class_off);
    const int ch = kRounds * (32 / L.lpr);
    const int rpb = 8 * ch;
    int nblk = (int)(cap.u_cap / rpb) + cfg->F + 1;
    if (nblk > sm.nblk_max) nblk = sm.nblk_max;
    const int grid_b = grid_warps(cfg->B, ch);
    const size_t smem = (size_t)8 * (2 * cfg->d + 4) * sizeof(float);
    VFMB_LAYOUT_SWITCH(L, {
        k_cstage<VEC, LPR, NV><<<nblk, 256, smem, counted(stream)>>>(
            cc, tab->bias, tab->entity, tab->train_counts, tab->scalars, plan->urec, plan->class_off, plan->z,
            io->vs, io->ws, io->ebs, io->cq, io->kl_bias_out, io->kl_entity_out, io->partials,
            io->counters + 0, pg_part, blk_class, io->stats, nullptr);
        k_cscore<VEC, LPR, NV><<<grid_b, 256, 0, counted(stream)>>>(
            cc, tab->scalars, plan->inverse, plan->pos_of, io->vs, io->ws, io->ebs, io->y, io->pred,
            io->resid, io->rsorted, io->msg, io->partials, io->counters + 1, io->stats);
    });
    CUDA_TRY(cudaGetLastError());
    if (io->mean && io->mean != io->pred)
        CUDA_TRY(cudaMemcpyAsync(io->mean, io->pred, (size_t)cfg->B * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    return 0;
}

extern "C" int vfmb_closed_backward(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                    const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                    vfmb_stream stream_) {
    if (!cfg) return set_error(VFMB_EINVAL, "vfmb_closed_backward: null argument");
    return vfmb_closed_backward_weighted(cfg, tab, plan, io, adam, mode, nullptr, cfg->n_train / (float)cfg->B, 1.f, stream_);
}

extern "C" int vfmb_closed_backward_weighted(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                                             const vfmb_step_io* io, const vfmb_adam* adam, int32_t mode,
                                             const float* kl_weight, float data_scale, float kl0_scale,
                                             vfmb_stream stream_) {
    int rc = closed_check(cfg, tab, plan, io, "vfmb_closed_backward");
    if (rc) return rc;
    if (mode == VFMB_ADAM_TOUCHED && (!adam || !tab->entity_m || !tab->entity_v || !tab->bias_m || !tab->bias_v ||
                                      !tab->scalars_m || !tab->scalars_v || !tab->adam_step))
        return set_error(VFMB_EINVAL, "vfmb_closed_backward: Adam state required");
    if (mode == VFMB_GRAD_ONLY && (!io->grad_bias || !io->grad_entity))
        return set_error(VFMB_EINVAL, "vfmb_closed_backward: gradient outputs required");
    if (!io->grad_scalars) return set_error(VFMB_EINVAL, "vfmb_closed_backward: grad_scalars scratch required");
    if (mode != VFMB_ADAM_TOUCHED && mode != VFMB_GRAD_ONLY) return set_error(VFMB_EINVAL, "vfmb_closed_backward: bad mode");
    cudaStream_t stream = (cudaStream_t)stream_;
    Layout L;
    if (!pick_layout(cfg->d, &L)) return set_error(VFMB_ESHAPE, "unsupported embedding size %d", cfg->d);
    vfmb_plan_capacity_t cap;
    rc = vfmb_plan_capacity(cfg->B, cfg->F, cfg->R, &cap);
    if (rc) return rc;
    ClosedCfg cc = make_closed(cfg);
    cc.data_scale = data_scale; cc.kl0_scale = kl0_scale;
    AdamDev h = make_adam(adam);
    const ScratchMap sm = scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap);
    float* fbase = (float*)io->partials;
    float* gslot = fbase + sm.gslot_off;
    float* pg_part = fbase + sm.pg_part_off;
    float* pg_sum = fbase + sm.pg_sum_off;
    int32_t* blk_class = (int32_t*)(fbase + sm.blk_class_off);
    const int ch = kRounds * (32 / L.lpr);
    int nblk = (int)(cap.u_cap / (8 * ch)) + cfg->F + 1;
    if (nblk > sm.nblk_max) nblk = sm.nblk_max;
    const int grid_u = grid_warps(cap.u_cap, ch), grid_t = grid_warps(cap.n_tiles, 32 / L.lpr);
    if (kl_weight) {
        // caller-supplied per-row KL weights: redo the per-row pass of the forward with them (KL weights,
        // per-group sums behind the prior-parameter gradients); the staged rows are rewritten unchanged
        const size_t smem = (size_t)8 * (2 * cfg->d + 4) * sizeof(float);
        VFMB_LAYOUT_SWITCH(L, {
            k_cstage<VEC, LPR, NV><<<nblk, 256, smem, counted(stream)>>>(
                cc, tab->bias, tab->entity, tab->train_counts, tab->scalars, plan->urec, plan->class_off, plan->z,
                io->vs, io->ws, io->ebs, io->cq, nullptr, nullptr, io->partials, io->counters + 0, pg_part,
                blk_class, io->stats, kl_weight);
        });
    }
    int32_t* arrive = (int32_t*)(fbase + sm.arrive_off);
    VFMB_LAYOUT_SWITCH(L, {
        k_cgather<VEC, LPR, NV><<<grid_t, 256, 0, counted(stream)>>>(cfg->d, cfg->F, cfg->B * cfg->F, plan->partner,
                                                            plan->pos_rank, plan->urec, io->vs, io->msg, io->rsorted,
                                                            gslot, io->grow, io->gws, arrive,
                                                            GatherKnobs{tuning().gather_keep, 0, tuning().gather_fence});
    });
    cudaEvent_t ev0, ev1;                                   // measurement hook (vfmb_profile_events)
    profile_events(&ev0, &ev1);
    if (ev0 && ev1) record_profile_event(ev0, stream);
    VFMB_LAYOUT_SWITCH(L, {
        if (mode == VFMB_ADAM_TOUCHED)
            k_cadam<VEC, LPR, NV, VFMB_ADAM_TOUCHED><<<grid_u, 256, 0, counted(stream)>>>(
                cc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v, tab->scalars,
                plan->urec, plan->meta, io->cq, io->grow, io->gws, h, tab->adam_step, io->grad_bias, io->grad_entity);
        else
            k_cadam<VEC, LPR, NV, VFMB_GRAD_ONLY><<<grid_u, 256, 0, counted(stream)>>>(
                cc, tab->bias, tab->bias_m, tab->bias_v, tab->entity, tab->entity_m, tab->entity_v, tab->scalars,
                plan->urec, plan->meta, io->cq, io->grow, io->gws, h, tab->adam_step, io->grad_bias, io->grad_entity);
    });
    if (ev0 && ev1) record_profile_event(ev1, stream);
    const int items = cfg->F * (2 * cfg->d + 3);
    const int grid_f = (items + 7) / 8;
    if (mode == VFMB_ADAM_TOUCHED)
        k_cfinal<VFMB_ADAM_TOUCHED><<<grid_f, 256, 0, counted(stream)>>>(cc, tab->scalars, tab->scalars_m, tab->scalars_v, io->stats,
                                                                pg_part, blk_class, nblk, pg_sum, io->counters + 2, h,
                                                                tab->adam_step, io->grad_scalars);
    else
        k_cfinal<VFMB_GRAD_ONLY><<<grid_f, 256, 0, counted(stream)>>>(cc, tab->scalars, tab->scalars_m, tab->scalars_v, io->stats,
                                                             pg_part, blk_class, nblk, pg_sum, io->counters + 2, h,
                                                             tab->adam_step, io->grad_scalars);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
