// Sampled-ELBO VFM step (vfm-torch.py:189-324 forward, :359 loss, :368-370 backward + Adam).
//
// Four launches per fused training step (S = 1), all HBM/L2-bound (no dense contraction on this path):
//   k_stage      one lane group per UNIQUE row (16 / 32 rows per warp pass): gather [mean|raw scale], draw
//                eps (Philox or injected), write the sampled row v = mu + eps*|rho| and bias w to an L2-resident
//                scratch (+ the count-rescaled KL when not fused).   (vfm-torch.py:207-241, 290-317)
//   k_score      one lane group per SAMPLE (32 per warp pass): FM interaction of the sampled rows, likelihood,
//                residual dloss/dpred.                                (vfm-torch.py:244-270, 359)
//   k_gather     ordered segmented sum of residual * partner row over the sorted occurrence list,
//                in 512-position block tiles; rows cut inside a block are combined through shared memory,
//                rows cut by block-tile boundaries by the group that stores their last partial
//                (finish_cut_row), in tile order.
//   k_adam_rows  per unique row: chain rule to (mu, rho) + KL gradient + Adam; FLAVOR 2 also forms
//                the KL and recomputes the Philox draws; the last block updates the scalar
//                parameters (alpha, global bias) and the step counter.
//                Deterministic: fixed summation order, no floating-point atomics.    (:368-370)
// S > 1 variational samples: k_stage / k_gather once per sample, k_score_multi, k_adam_rows_multi.
// Host side: launch_* (internal) and the extern "C" entry points at the end of the file.
#include "sampled_common.cuh"

namespace vfmb {

// ------------------------------------------------------------------------------- k_stage
// LEAN = 1 (fused training step): no KL sum and no copy of the noise -- k_adam_rows, which holds
// the row anyway and has idle issue slots, recomputes both (same Philox counters => same bits).
template <int VEC, int LPR, int NV, int LINK, int LEAN>
__global__ void __launch_bounds__(256)
k_stage(DevCfg c, const float* __restrict__ bias, const float* __restrict__ entity,
        const float* __restrict__ train_counts, const int32_t* __restrict__ urec,
        const int32_t* __restrict__ meta, const float* __restrict__ z,
        const float* __restrict__ eps_bias, const float* __restrict__ eps_entity,
        const int32_t* __restrict__ noise_step, float* __restrict__ vs, float* __restrict__ ws,
        float* __restrict__ es, float* __restrict__ ebs, float* __restrict__ cq,
        double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats,
        int smp, int u_stride, const float* __restrict__ pf_m, const float* __restrict__ pf_v, int keep, int ch, RowPut put) {
    chain_wait();                                           // launched with launch_chained()
    // put (mode B, owner side): the sampled row also goes into the slot of every rank that asked for it
    // pf_m / pf_v (optional): Adam moment tables whose touched rows are pulled into L2 here, while
    // this kernel is issue-bound on Philox and the DRAM pipe idles -- k_adam_rows finds them there
    // smp: variational sample this launch draws (S > 1: one launch per sample, outputs [S][u_stride]);
    // the KL and the KL weights do not depend on the sample and are formed by sample 0
    constexpr int GPW = kWarp / LPR, ROUNDS = LPR;      // up to 32 unique rows per warp and pass, GPW per round
    constexpr int UR = (NV == 1) ? 4 : 2;               // rounds whose [mean | raw scale] loads are in flight
    const int U = meta[0];
    const int d = c.d;
    if (vs) vs += (size_t)smp * u_stride * d;
    if (ws) ws += (size_t)smp * u_stride;
    if (es) es += (size_t)smp * u_stride * d;
    if (ebs) ebs += (size_t)smp * u_stride;
    const uint32_t step = noise_step ? (uint32_t)noise_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    float facc = 0.f;                                   // sum_u c_u * KL_u over this thread's rows
    int kc[NV]; bool act[NV];                           // column of this lane; lanes past d load column 0, store nothing
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int k = (gl + i * LPR) * VEC; act[i] = k < d; kc[i] = act[i] ? k : 0; }

    // ch (16 or 32): unique rows per warp and pass -- fewer rows per warp = more warps in flight and a shorter chain
    for (int base = gwarp * ch; base < U; base += nwarps * ch) {
        // ---- lane-parallel: one unique row per lane
        const int ul = base + lane;
        const bool valid = lane < ch && ul < U;
        const int nvalid = min(ch, U - base);
        int rowid_l = 0, seg0_l = 0, len_l = 0;          // (lanes past U gather row 0: harmless)
        float klb = 0.f, cqv = 0.f, w_l = 0.f;
        if (valid) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(urec) + ul);
            rowid_l = rec.x; len_l = rec.y; seg0_l = rec.z;
            prefetch_row(entity + (size_t)rowid_l * 2 * d, 8 * d, keep & 1);     // keep: see prefetch_l2_keep
            if (pf_m) prefetch_row(pf_m + (size_t)rowid_l * 2 * d, 8 * d, keep & 2);
            if (pf_v) prefetch_row(pf_v + (size_t)rowid_l * 2 * d, 8 * d, keep & 4);
            const float2 ab = *reinterpret_cast<const float2*>(bias + (size_t)rowid_l * 2);
            const float tcnt = __ldg(train_counts + rowid_l);
            const int gid_l = rowid_l * c.row_stride + c.row_offset;     // global id (row-sharded tables)
            const float eb = bias_eps(eps_bias, c, ul, gid_l, step, smp, U);
            if (!eps_bias) ebs[ul] = eb;
            const float tau = link_fn<LINK>(ab.y);
            w_l = fmaf(eb, tau, ab.x);
            if (ws) ws[ul] = w_l;
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
        for (int it0 = 0; it0 < ROUNDS; it0 += UR) {
            if (it0 * GPW >= nvalid) break;                  // warp-uniform
            Vec<VEC> mu[UR][NV], rho[UR][NV];
#pragma unroll
            for (int q = 0; q < UR; ++q) {
                const float* erow = entity + (size_t)bcast(rowid_l, (it0 + q) * GPW + gidx) * 2 * d;
#pragma unroll
                for (int i = 0; i < NV; ++i) { mu[q][i] = ld_vec<VEC>(erow + kc[i]); rho[q][i] = ld_vec<VEC>(erow + d + kc[i]); }
            }
#pragma unroll
            for (int q = 0; q < UR; ++q) {
                const int it = it0 + q;
                const int sel = it * GPW + gidx;
                const int rowid = bcast(rowid_l, sel);
                const int u = base + sel;
                // requesters of the row (mode B): its sorted occurrences in the owner's plan are request slots
                int n_req = 0, seg0 = 0;
                float w_u = 0.f;
                if (put.pe.n) {
                    n_req = bcast(len_l, sel); seg0 = bcast(seg0_l, sel); w_u = bcast(w_l, sel);
                    if (sel >= nvalid || rowid >= put.n_real) n_req = 0;   // padding slots share a sentinel row
                }
                auto put_row = [&](int k, const Vec<VEC>& out) {
                    for (int i = 0; i < n_req; ++i) {
                        const int sl = __ldg(put.occ + seg0 + i);
                        const int pq = sl / put.CAP, j = sl - pq * put.CAP;
                        float* dst = reinterpret_cast<float*>(put.pe.base[pq]) + ((size_t)put.pe.rank * put.CAP + j) * put.SP;
                        st_vec<VEC>(dst + k, out);
                        if (k == 0) dst[d] = w_u;
                    }
                };
                float kl = 0.f;
                if (sel < nvalid) {                              // (a round can reach past the warp's chunk)
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const int k = kc[i];
                        if (act[i]) {
                            Vec<VEC> e = entity_eps<VEC>(eps_entity, c, u, rowid * c.row_stride + c.row_offset, k, step, smp, U), out;
                            if (!LEAN && !eps_entity) st_vec<VEC>(es + (size_t)u * d + k, e);
                            if (LEAN) {
#pragma unroll
                                for (int j = 0; j < VEC; ++j) out.v[j] = fmaf(e.v[j], link_fn<LINK>(rho[q][i].v[j]), mu[q][i].v[j]);
                                if (vs) st_vec<VEC>(vs + (size_t)u * d + k, out);
                                put_row(k, out);
                                continue;
                            }
                            // sum_k KL(N(mu,sig)||N(0,1)) = 0.5 (sum sig^2 + mu^2 - 1) - 0.5 log prod sig^2:
                            // one logarithm per lane instead of one per element
                            float quad = 0.f, prodv = 1.f;
#pragma unroll
                            for (int j = 0; j < VEC; ++j) {
                                const float sig = link_fn<LINK>(rho[q][i].v[j]);
                                out.v[j] = fmaf(e.v[j], sig, mu[q][i].v[j]);
                                const float vr = sig * sig;
                                quad += vr + mu[q][i].v[j] * mu[q][i].v[j] - 1.f;
                                prodv *= vr;
                            }
                            float lg = __logf(prodv);
                            if (!(prodv > 1e-30f && prodv < 1e30f)) {       // tiny / huge scales: no product trick
                                lg = 0.f;
#pragma unroll
                                for (int j = 0; j < VEC; ++j) { const float sg = link_fn<LINK>(rho[q][i].v[j]); lg += logf(sg * sg); }
                            }
                            kl += 0.5f * (quad - lg);
                            if (vs) st_vec<VEC>(vs + (size_t)u * d + k, out);
                            put_row(k, out);
                        }
                    }
                }
                if (!LEAN) { kl = group_sum<LPR>(kl, 0xffffffffu); hand_back<LPR>(klrow, kl, it, lane); }
            }
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
// Forward, phase B: per sample the logit (global bias + bias sum + FM interaction of the sampled rows),
// the likelihood terms, the residual r_n = dloss/dpred_n (also scattered into sorted-occurrence order
// for the backward) and, F > 2, the field sums S_n the backward gathers.
// A warp takes 32 consecutive samples per pass: the per-sample scalars are lane-parallel (one sample
// per lane, coalesced), the d-wide interaction runs GPW samples per round with the rows of UR rounds
// in flight.  The kernel is bound by L2 round trips per warp, so the passes are short (one per warp on
// the benchmark shapes) and every load of a round is issued before the first is used.
template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_score(DevCfg c, const float* __restrict__ scalars, const int32_t* __restrict__ inverse,
        const int32_t* __restrict__ pos_of, const float* __restrict__ vs, const float* __restrict__ ws,
        const float* __restrict__ y, const float* __restrict__ eps_global,
        int32_t* noise_step, const int32_t* __restrict__ meta, float* __restrict__ pred, float* __restrict__ mean,
        float* __restrict__ resid, float* __restrict__ rsorted, float* __restrict__ msg,
        double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats,
        int defer_kl, int vp, int ch, TailPut tp) {
    chain_wait();                                           // launched with launch_chained()
    // vp: pitch of the sampled rows in floats (d; mode B: the received slots, d + 4, bias sample at [d] and
    // `inverse` holding slot indices; ws == NULL then).  tp: mode B, see TailPut.
    constexpr int GPW = kWarp / LPR, ROUNDS = LPR;      // up to 32 samples per warp and pass, GPW per round
    constexpr int UR = (NV == 1) ? 4 : 2;               // F == 2: rounds in flight (two rows each)
    constexpr int FU = (NV == 1) ? 4 : 2;               // F > 2: rows of one sample in flight
    __shared__ int s_idx[8][kWarp * kMaxFields];        // F > 2: row ranks of the warp's 32 samples
    const int d = c.d, F = c.F, B = c.B;
    const uint32_t step = noise_step ? (uint32_t)noise_step[0] : 0u;
    const int lane = threadIdx.x & 31, gl = lane % LPR, gidx = lane / LPR;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const float mu0 = scalars[VFMB_S_GB_MEAN];
    const float sig0 = link_fn<LINK>(scalars[VFMB_S_GB_SCALE]);
    const float w0 = mu0 + global_eps(eps_global, c, step) * sig0;
    const float alpha = link_fn<LINK>(scalars[VFMB_S_ALPHA]);
    const float half_log_alpha = 0.5f * logf(alpha);
    const float scale = c.n_train / ((float)c.S * (float)B);
    double acc[3] = {0.0, 0.0, 0.0};     // nll, resid, squared error
    int kc[NV]; bool act[NV];            // column of this lane; lanes past d load column 0 and contribute nothing
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int k = (gl + i * LPR) * VEC; act[i] = k < d; kc[i] = act[i] ? k : 0; }

    // ch (16 or 32): samples per warp and pass
    for (int base = gwarp * ch; base < B; base += nwarps * ch) {
        // ---- lane-parallel: one sample per lane (ranks, bias sum, target)
        const int nl = base + lane;
        const bool valid = lane < ch && nl < B;
        const int nvalid = min(ch, B - base);
        int2 rr = make_int2(0, 0);                                    // (invalid samples gather row 0: harmless)
        float bsum = 0.f, yn = 0.f;
        int* si = s_idx[threadIdx.x >> 5];
        if (F == 2) {
            if (valid) {
                rr = __ldg(reinterpret_cast<const int2*>(inverse) + nl);
                bsum = ws ? __ldg(ws + rr.x) + __ldg(ws + rr.y)
                          : __ldg(vs + (size_t)rr.x * vp + d) + __ldg(vs + (size_t)rr.y * vp + d);
            }
        } else {
            for (int q = lane; q < nvalid * F; q += kWarp) si[q] = __ldg(inverse + (size_t)base * F + q);   // coalesced
            __syncwarp();
            if (valid)
                for (int f = 0; f < F; ++f) {
                    const int r = si[lane * F + f];
                    bsum += ws ? __ldg(ws + r) : __ldg(vs + (size_t)r * vp + d);
                }
        }
        if (valid && y) yn = __ldg(y + nl);
        float inter_l = 0.f;
        // ---- wide work: interaction of GPW samples per round
        if (F == 2) {
#pragma unroll 1
            for (int it0 = 0; it0 < ROUNDS; it0 += UR) {
                if (it0 * GPW >= nvalid) break;                       // warp-uniform
                Vec<VEC> ra[UR][NV], rb[UR][NV];
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const int sel = (it0 + u) * GPW + gidx;
                    const float* p0 = vs + (size_t)bcast(rr.x, sel) * vp;
                    const float* p1 = vs + (size_t)bcast(rr.y, sel) * vp;
#pragma unroll
                    for (int i = 0; i < NV; ++i) { ra[u][i] = ld_vec_nc<VEC>(p0 + kc[i]); rb[u][i] = ld_vec_nc<VEC>(p1 + kc[i]); }
                }
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    float part = 0.f;
#pragma unroll
                    for (int i = 0; i < NV; ++i)
                        if (act[i])
#pragma unroll
                            for (int j = 0; j < VEC; ++j) part = fmaf(ra[u][i].v[j], rb[u][i].v[j], part);
                    part = group_sum<LPR>(part, 0xffffffffu);
                    hand_back<LPR>(inter_l, part, it0 + u, lane);
                }
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < ROUNDS; ++it) {
                if (it * GPW >= nvalid) break;                        // warp-uniform
                const int sel = it * GPW + gidx;
                const int n = base + sel;
                const int* sidx = si + min(sel, nvalid - 1) * F;
                Vec<VEC> ssum[NV], sq[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i)
#pragma unroll
                    for (int j = 0; j < VEC; ++j) { ssum[i].v[j] = 0.f; sq[i].v[j] = 0.f; }
                for (int f0 = 0; f0 < F; f0 += FU) {
                    Vec<VEC> a[FU][NV];
#pragma unroll
                    for (int e = 0; e < FU; ++e) {
                        const float* row = vs + (size_t)sidx[min(f0 + e, F - 1)] * vp;
#pragma unroll
                        for (int i = 0; i < NV; ++i) a[e][i] = ld_vec_nc<VEC>(row + kc[i]);
                    }
#pragma unroll
                    for (int e = 0; e < FU; ++e)
                        if (f0 + e < F)
#pragma unroll
                            for (int i = 0; i < NV; ++i)
#pragma unroll
                                for (int j = 0; j < VEC; ++j) {
                                    ssum[i].v[j] += a[e][i].v[j];
                                    sq[i].v[j] = fmaf(a[e][i].v[j], a[e][i].v[j], sq[i].v[j]);
                                }
                }
                float part = 0.f;
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (act[i]) {
#pragma unroll
                        for (int j = 0; j < VEC; ++j) part += 0.5f * (ssum[i].v[j] * ssum[i].v[j] - sq[i].v[j]);
                        if (msg && sel < nvalid) st_vec<VEC>(msg + (size_t)n * d + kc[i], ssum[i]);   // S_n = sum_f v_f (unscaled)
                    }
                part = group_sum<LPR>(part, 0xffffffffu);
                hand_back<LPR>(inter_l, part, it, lane);
            }
            __syncwarp();                                             // s_idx is rewritten by the next pass
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
            forward_done(noise_step, step, meta, stats);
            *counter = 0;
        }
        if (tp.pe.n) {                       // mode B: this rank's additive scalars to every rank (peer stores)
            __syncthreads();
            if (threadIdx.x < VFMB_DP_TAIL) {
                const int t = threadIdx.x;
                float v = 0.f;
                if (t == VFMB_DP_T_NLL) v = stats[VFMB_ST_NLL_MEAN] * tp.n_local;
                else if (t == VFMB_DP_T_RESID) v = stats[VFMB_ST_SUM_RESID];
                else if (t == VFMB_DP_T_SQERR) v = stats[VFMB_ST_SUM_SQERR];
                else if (t == VFMB_DP_T_KLROWS) v = tp.stats_owner[VFMB_ST_KL_ROWS];
                else if (t == VFMB_DP_T_OVERFLOW) v = (tp.overflow && tp.overflow[0]) ? 1.f : 0.f;
                for (int q = 0; q < tp.pe.n; ++q)
                    reinterpret_cast<float*>(tp.pe.base[q])[(size_t)tp.pe.rank * tp.pitch + t] = v;
            }
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
template <int VEC, int LPR, int NV, int LINK, int LIK>
__global__ void __launch_bounds__(256)
k_score_multi(DevCfg c, int u_stride, const float* __restrict__ scalars, const int32_t* __restrict__ inverse,
              const int32_t* __restrict__ pos_of, const float* __restrict__ vs, const float* __restrict__ ws,
              const float* __restrict__ y, const float* __restrict__ eps_global,
              int32_t* noise_step, const int32_t* __restrict__ meta, float* __restrict__ pred, float* __restrict__ mean,
              float* __restrict__ resid, float* __restrict__ rsorted, float* __restrict__ msg,
              double* __restrict__ partials, int32_t* __restrict__ counter, float* __restrict__ stats) {
    constexpr int GPW = kWarp / LPR;
    const int d = c.d, F = c.F, B = c.B, S = c.S;
    const uint32_t step = noise_step ? (uint32_t)noise_step[0] : 0u;
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
            forward_done(noise_step, step, meta, stats);
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
// Backward, phase A: deterministic segmented reduction over the SORTED occurrence list,
//   g_u = sum over the occurrences (u, n) of  r_n * (partner row | field sums S_n),   g_w,u = sum r_n.
// Tiled by POSITION, so every lane group does the same amount of work whatever the row popularity
// (a Zipf head row with thousands of occurrences is just many tiles), in two levels:
//   block tile  512 consecutive positions per block and pass, handed out by a counter;
//   group tile  2 * LPR of them per lane group (two positions per lane: all index loads of a tile are
//               one round trip, the row gathers UNR at a time).
// A group walks its positions in order and flushes when the unique rank changes.  Rows inside the
// group tile are final (-> grow/gws, or the owner's slot over NVLink).  A row cut by a group-tile
// boundary leaves a partial in SHARED memory; after one __syncthreads the group holding the row's last
// piece in the block adds the pieces in tile order.  Only a row that crosses a BLOCK-tile boundary goes
// through global slots and finish_cut_row (fence + arrival counter; <= 2 per block tile -- in round-2
// profiles the per-tile fences and atomics of a flat 32-position tiling cost as much as the gather).
// The two groups of a warp run in lockstep (same trip counts, pads carry a zero coefficient), so the
// shuffles are full-mask and the compiler needs no convergence checks.  Fixed summation order, no
// floating-point atomics => bitwise reproducible whatever the schedule.
// Everything read here is L2-resident scratch written by k_stage / k_score.
#ifdef VFMB_TILE_TIMING                                     // measurement build (scripts/gather_tiles.py), not the product
__device__ int4 g_tile_dbg[1 << 16];                        // per block tile: main cycles, combine cycles, SM, start (ns)
__device__ __forceinline__ unsigned dbg_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (unsigned)t; }
__device__ __forceinline__ int dbg_sm() { int s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
extern "C" int vfmb_debug_tile_times(int32_t* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_tile_dbg, (size_t)n * sizeof(int4));
}
#endif

#ifndef VFMB_GATHER_MAXREG
#define VFMB_GATHER_MAXREG 0
#endif
template <int VEC, int LPR, int NV, int UNIT>
#if VFMB_GATHER_MAXREG
__global__ void __maxnreg__(VFMB_GATHER_MAXREG)
#else
__global__ void __launch_bounds__(256, 2)
#endif
k_gather(int d, int F, int N, const int32_t* __restrict__ partner, const int32_t* __restrict__ pos_rank,
         const int32_t* __restrict__ urec, const float* __restrict__ vs, const float* __restrict__ msg,
         const float* __restrict__ rsorted, float* gslot, float* __restrict__ grow, float* __restrict__ gws,
         int32_t* arrive, int vp, float* const* __restrict__ gptr, GatherKnobs kn) {
    chain_wait();                                           // launched with launch_chained()
    // F > 2 (pairwise): the sums are over the per-sample field sums S_n; the row's own term (sum r_n) v_u is
    // removed by the row kernel, which holds v_u anyway (DevCfg.pairwise).
    // mode B: vp = pitch of the gathered rows (received slots: d + 4); gptr[u] = where the finished gradient
    // row of unique rank u goes (its owner's slot over NVLink, bias gradient at [d]).
    // UNIT: unit coefficients (acc += row; the collective mode-B owner side), g_w still sums rsorted.
    constexpr int GPW = kWarp / LPR, G = 8 * GPW;       // lane groups per warp / per block
    constexpr int TP = 2 * LPR;                         // positions per group tile
    constexpr int BT = G * TP;                          // positions per block tile (512)
    constexpr int W = LPR * VEC * NV;                   // row width covered by a group
    constexpr int SP = W + 4;                           // pitch of a partial in shared memory (sum of r at [W])
    constexpr int UNR = LPR < 8 ? LPR : (NV == 1 ? 8 : 4);
    __shared__ __align__(16) float s_part[2][G][2][SP]; // [parity][group][head | tail]
    __shared__ int s_hu[2][G], s_tu[2][G], s_next[2];

    const int dp = d + 4;                               // global slot pitch (keeps 16 B alignment)
    const int lane = threadIdx.x & 31, gl = lane % LPR, gbase = lane - gl;
    const int g = (threadIdx.x >> 5) * GPW + lane / LPR;
    const unsigned gmask = group_mask<LPR>();
    const int n_bt = (N + BT - 1) / BT;
    const float* table = (F == 2) ? vs : msg;
    const int tpitch = (F == 2 || UNIT) ? vp : d;       // F > 2: the per-sample field sums are local, d wide
    int32_t* tctr = arrive + 3 * (n_bt + 1);            // [0] block tiles handed out, [1] blocks done; both end at zero
    int kc[NV]; bool act[NV];                           // column of this lane; lanes past d load column 0 and store nothing
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int k = (gl + i * LPR) * VEC; act[i] = k < d; kc[i] = act[i] ? k : 0; }
    auto row_out = [&](int u) { return gptr ? gptr[u] : grow + (size_t)u * d; };

    int it = 0;
    for (int bt = blockIdx.x; bt < n_bt; ++it) {
        const int par = it & 1;
        int fetched = 0;
        if (kn.dyn && threadIdx.x == 0) fetched = atomicAdd(tctr, 1);   // in flight during the tile
#ifdef VFMB_TILE_TIMING
        const long long dbg_c0 = clock64(); const unsigned dbg_t0 = dbg_ns();
#endif
        const int t0 = bt * BT + g * TP, t1 = min(N, t0 + TP), cnt = max(0, t1 - t0);
        const int wt0 = bt * BT + (threadIdx.x >> 5) * GPW * TP;        // first position of the warp
        // ---- lane-parallel: two positions per lane, plus the ranks just outside the tile
        int src[2], ur[2]; float rr[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = t0 + h * LPR + gl;
            const bool ok = idx < t1;
            src[h] = ok ? __ldg(partner + idx) : 0;
            rr[h] = ok ? __ldg(rsorted + idx) : 0.f;
            ur[h] = ok ? __ldg(pos_rank + idx) : -1;
        }
        int nb = -1;
        if (gl == 0 && cnt > 0 && t0 > 0) nb = __ldg(pos_rank + t0 - 1);
        if (gl == LPR - 1 && cnt > 0 && t1 < N) nb = __ldg(pos_rank + t1);
        const int prev_u = __shfl_sync(0xffffffffu, nb, gbase), next_u = __shfl_sync(0xffffffffu, nb, gbase + LPR - 1);
        const int first_u = __shfl_sync(0xffffffffu, ur[0], gbase);
        const int la = __shfl_sync(0xffffffffu, ur[0], gbase + min(max(cnt, 1), LPR) - 1);
        const int lb = __shfl_sync(0xffffffffu, ur[1], gbase + min(max(cnt - LPR, 1), LPR) - 1);
        const int last_u = cnt > LPR ? lb : la;                       // -1 when the tile has no position
#pragma unroll
        for (int h = 0; h < 2; ++h) if (ur[h] < 0) ur[h] = last_u;    // pads never open a segment
        const int head_u = (cnt > 0 && prev_u == first_u) ? first_u : -1;   // the row continues from the previous tile
        const int tail_u = (cnt > 0 && next_u == last_u) ? last_u : -1;     // ... into the next one

        int cur = first_u;
        Vec<VEC> acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[i].v[q] = 0.f;
        float gw = 0.f;
        auto flush = [&](int u) {
            if (u == head_u || u == tail_u) {                         // partial: shared memory
                float* dst = &s_part[par][g][u == head_u ? 0 : 1][0];
#pragma unroll
                for (int i = 0; i < NV; ++i) st_vec<VEC>(dst + (gl + i * LPR) * VEC, acc[i]);
                if (gl == 0) dst[W] = gw;
            } else {
                float* dst = row_out(u);
#pragma unroll
                for (int i = 0; i < NV; ++i) if (act[i]) st_vec<VEC>(dst + kc[i], acc[i]);
                if (gl == 0) *(gptr ? dst + d : gws + u) = gw;
            }
        };

#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll 1
            for (int j0 = 0; j0 < LPR; j0 += UNR) {
                if (wt0 + h * LPR + j0 >= N) break;                   // warp-uniform: nothing left for either group
                Vec<VEC> t[UNR][NV]; int uj[UNR]; float cj[UNR];
#pragma unroll
                for (int e = 0; e < UNR; ++e) {                       // UNR row gathers in flight
                    const int sj = __shfl_sync(0xffffffffu, src[h], gbase + j0 + e);
                    uj[e] = __shfl_sync(0xffffffffu, ur[h], gbase + j0 + e);
                    cj[e] = __shfl_sync(0xffffffffu, rr[h], gbase + j0 + e);
                    const float* row = table + (size_t)sj * tpitch;
#pragma unroll
                    for (int i = 0; i < NV; ++i) t[e][i] = ld_vec_nc<VEC>(row + kc[i]);
                }
#pragma unroll
                for (int e = 0; e < UNR; ++e) {
                    if (uj[e] != cur) {                               // group-uniform
                        flush(cur);
                        cur = uj[e];
#pragma unroll
                        for (int i = 0; i < NV; ++i)
#pragma unroll
                            for (int q = 0; q < VEC; ++q) acc[i].v[q] = 0.f;
                        gw = 0.f;
                    }
                    gw += cj[e];
                    const float ce = UNIT ? ((t0 + h * LPR + j0 + e < t1) ? 1.f : 0.f) : cj[e];   // pads: zero
#pragma unroll
                    for (int i = 0; i < NV; ++i)
#pragma unroll
                        for (int q = 0; q < VEC; ++q) acc[i].v[q] = fmaf(ce, t[e][i].v[q], acc[i].v[q]);
                }
            }
        }
        if (cur >= 0) flush(cur);
        if (gl == 0) { s_hu[par][g] = head_u; s_tu[par][g] = tail_u; }
        if (threadIdx.x == 0) s_next[par] = kn.dyn ? (int)gridDim.x + fetched : bt + (int)gridDim.x;
#ifdef VFMB_TILE_TIMING
        const long long dbg_c1 = clock64();
#endif
        __syncthreads();
        // ---- rows cut by group-tile boundaries inside the block: the group holding the last piece adds them
        const int g_last = (min(N, bt * BT + BT) - bt * BT - 1) / TP;     // last group with positions
        if (head_u >= 0 && (tail_u != head_u || g == g_last)) {
            const int u = head_u;
            int s0 = g;
            while (s0 > 0 && s_hu[par][s0] == u) --s0;                    // first piece: group s0
            const bool open_left = s_hu[par][s0] == u;                    // (s0 == 0) the row comes from the previous block tile
            const bool open_right = tail_u == u;                          // (g == g_last) ... runs on into the next one
            const float* p0 = &s_part[par][s0][open_left ? 0 : 1][0];
            Vec<VEC> tot[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) tot[i] = ld_vec<VEC>(p0 + (gl + i * LPR) * VEC);
            float gwt = p0[W];
            for (int k = s0 + 1; k <= g; ++k) {                           // head partials, in tile order
                const float* pk = &s_part[par][k][0][0];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const Vec<VEC> a = ld_vec<VEC>(pk + (gl + i * LPR) * VEC);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) tot[i].v[q] += a.v[q];
                }
                gwt += pk[W];
            }
            float* o = row_out(u);
            float* ow = gptr ? o + d : gws + u;
            const bool cut = open_left || open_right;
            float* dst = cut ? gslot + ((size_t)bt * 2 + (open_left ? 0 : 1)) * dp : o;
#pragma unroll
            for (int i = 0; i < NV; ++i) if (act[i]) st_vec<VEC>(dst + kc[i], tot[i]);
            if (gl == 0) *(cut ? dst + d : ow) = gwt;
            if (cut) finish_cut_row<VEC, LPR, NV, 1>(u, bt, d, nullptr, urec, gslot, o, ow, arrive, n_bt + 1, kn.light_fence, BT);
        }
        if (tail_u >= 0 && tail_u != head_u && g == g_last) {             // a row that starts in the last group tile and runs on
            const int u = tail_u;
            const float* p0 = &s_part[par][g][1][0];
            float* dst = gslot + ((size_t)bt * 2 + 1) * dp;
#pragma unroll
            for (int i = 0; i < NV; ++i) if (act[i]) st_vec<VEC>(dst + kc[i], ld_vec<VEC>(p0 + (gl + i * LPR) * VEC));
            if (gl == 0) dst[d] = p0[W];
            float* o = row_out(u);
            finish_cut_row<VEC, LPR, NV, 1>(u, bt, d, nullptr, urec, gslot, o, gptr ? o + d : gws + u, arrive, n_bt + 1, kn.light_fence, BT);
        }
#ifdef VFMB_TILE_TIMING
        if (threadIdx.x == 0) g_tile_dbg[bt & 0xFFFF] = make_int4((int)(dbg_c1 - dbg_c0), (int)(clock64() - dbg_c1), dbg_sm(), (int)dbg_t0);
#endif
        bt = s_next[par];          // (parity buffers: the next tile's partials do not touch what a slow combiner still reads)
    }
    if (kn.dyn && threadIdx.x == 0 && atomicAdd(tctr + 1, 1) == (int)gridDim.x - 1) { tctr[0] = 0; tctr[1] = 0; }   // last block out
}

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

// ---- internal launchers (the public entry points below are thin wrappers)
static int launch_stage(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                        const vfmb_step_io* io, vfmb_stream stream_, bool lean, int smp = 0, const RowPut* put_ = nullptr) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_stage", stream_, 1, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: null argument");
    if (!tab->noise_step) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: tables.noise_step required");
    if (!lean && !io->eps_entity && !io->es) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: noise scratch (es) required");
    if (!io->eps_bias && !io->ebs) return set_error(VFMB_EINVAL, "vfmb_sampled_stage: noise scratch (ebs) required");
    if (((!io->vs || !io->ws) && !put_) || !io->cq || !io->partials || !io->counters || !io->stats)
        return set_error(VFMB_EINVAL, "vfmb_sampled_stage: scratch required");
    RowPut put{};
    if (put_) put = *put_;
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream; const auto& cap = P.cap;
    // fused training step only: the row update follows within the same step
    const int keep = lean ? tuning().l2_keep : 0;
    const int ch = tuning().stage_chunk;
    const float* pf_m = (lean && ((tuning().prefetch_mv & 1) || (keep & 2))) ? tab->entity_m : nullptr;
    const float* pf_v = (lean && ((tuning().prefetch_mv & 2) || (keep & 4))) ? tab->entity_v : nullptr;
#define LAUNCH_STAGE(LINK, LEAN)                                                                       \
    CUDA_TRY(launch_chained(k_stage<VEC, LPR, NV, LINK, LEAN>, grid_resident(k_stage<VEC, LPR, NV, LINK, LEAN>, cap.u_cap, ch), 256, 0, stream, \
        dc, tab->bias, tab->entity, tab->train_counts, plan->urec, plan->meta, plan->z,                \
        io->eps_bias, io->eps_entity, tab->noise_step, io->vs, io->ws, io->es,                         \
        io->ebs, io->cq, io->partials, io->counters + 0, io->stats, smp, (int)cap.u_cap, pf_m, pf_v, keep, ch, put))
#define LAUNCH_STAGE_ALL()                                                                             \
    do {                                                                                               \
        if (cfg->link == VFMB_LINK_ABS) { if (lean) LAUNCH_STAGE(0, 1); else LAUNCH_STAGE(0, 0); }     \
        else                            { if (lean) LAUNCH_STAGE(1, 1); else LAUNCH_STAGE(1, 0); }     \
    } while (0)
    // stage_wide: half the lanes per row, two vectors per lane (the draws are keyed by row and column, so
    // the sampled rows are the same bits; only the KL sum of the non-lean forward changes its last bit)
    const bool wide = tuning().stage_wide && L.vec == 4 && L.nv == 1 && L.lpr >= 8;
    if (wide && L.lpr == 8) { constexpr int VEC = 4, LPR = 4, NV = 2; LAUNCH_STAGE_ALL(); }
    else if (wide && L.lpr == 16) { constexpr int VEC = 4, LPR = 8, NV = 2; LAUNCH_STAGE_ALL(); }
    else if (wide && L.lpr == 32) { constexpr int VEC = 4, LPR = 16, NV = 2; LAUNCH_STAGE_ALL(); }
    else VFMB_LAYOUT_SWITCH(L, { LAUNCH_STAGE_ALL(); });
#undef LAUNCH_STAGE_ALL
#undef LAUNCH_STAGE
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// mode B, requester side: the sampled rows are read in place from the received slots
struct ScoreB { const float* rows; int vp; const int32_t* inv_slot; TailPut tp; };

static int launch_score(const vfmb_config* cfg, const vfmb_tables* tab, const vfmb_plan* plan,
                        const vfmb_step_io* io, vfmb_stream stream_, int defer_kl, const ScoreB* sb = nullptr) {
    Prep P;
    int rc = prep(cfg, "vfmb_sampled_score", stream_, 2, &P);
    if (rc) return rc;
    if (!tab || !plan || !io) return set_error(VFMB_EINVAL, "vfmb_sampled_score: null argument");
    if (!tab->noise_step) return set_error(VFMB_EINVAL, "vfmb_sampled_score: tables.noise_step required");
    if (cfg->F > 2 && io->y && !io->msg) return set_error(VFMB_EINVAL, "vfmb_sampled_score: msg scratch required for F>2");
    const Layout& L = P.L; const DevCfg& dc = P.dc; cudaStream_t stream = P.stream;
    const float* rows = sb ? sb->rows : io->vs;
    const float* wsp = sb ? nullptr : io->ws;
    const int32_t* inv = sb ? sb->inv_slot : plan->inverse;
    const int vp = sb ? sb->vp : cfg->d;
    TailPut tp{};
    if (sb) tp = sb->tp;
    const int ch = tuning().score_chunk;
#define LAUNCH_SCORE(LINK, LIK)                                                                        \
    CUDA_TRY(launch_chained(k_score<VEC, LPR, NV, LINK, LIK>, grid_resident(k_score<VEC, LPR, NV, LINK, LIK>, cfg->B, ch), 256, 0, stream, \
        dc, tab->scalars, inv, plan->pos_of, rows, wsp, io->y, io->eps_global,                         \
        tab->noise_step, plan->meta, io->pred, io->mean, io->resid, io->rsorted, io->msg, io->partials, \
        io->counters + 1, io->stats, defer_kl, vp, ch, tp))
#define LAUNCH_SCORE_ALL()                                                                              \
    do {                                                                                                \
        if (cfg->link == VFMB_LINK_ABS) {                                                               \
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SCORE(0, VFMB_GAUSSIAN); else LAUNCH_SCORE(0, VFMB_BERNOULLI); \
        } else {                                                                                        \
            if (cfg->likelihood == VFMB_GAUSSIAN) LAUNCH_SCORE(1, VFMB_GAUSSIAN); else LAUNCH_SCORE(1, VFMB_BERNOULLI); \
        }                                                                                               \
    } while (0)
    // score_wide: half the lanes per row, two vectors per lane -- twice the samples per round.  The dot
    // product is then summed in another (fixed) order: a different last bit, so it is one knob for the process.
    const int sw = tuning().score_wide;                        // -1: for F > 2 (sideinfo: 49.8 -> 34.5 us; ml20m: no gain in the step)
    const bool wide = (sw < 0 ? cfg->F > 2 : sw != 0) && L.vec == 4 && L.nv == 1 && L.lpr >= 8;
    if (wide && L.lpr == 8) { constexpr int VEC = 4, LPR = 4, NV = 2; LAUNCH_SCORE_ALL(); }
    else if (wide && L.lpr == 16) { constexpr int VEC = 4, LPR = 8, NV = 2; LAUNCH_SCORE_ALL(); }
    else if (wide && L.lpr == 32) { constexpr int VEC = 4, LPR = 16, NV = 2; LAUNCH_SCORE_ALL(); }
    else VFMB_LAYOUT_SWITCH(L, { LAUNCH_SCORE_ALL(); });
#undef LAUNCH_SCORE_ALL
#undef LAUNCH_SCORE
    CUDA_TRY(cudaGetLastError());
    return 0;
}

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
    k_score_multi<VEC, LPR, NV, LINK, LIK><<<grid, 256, 0, counted(stream)>>>(                                    \
        dc, (int)P.cap.u_cap, tab->scalars, plan->inverse, plan->pos_of, io->vs, io->ws, io->y, io->eps_global, \
        tab->noise_step, plan->meta, io->pred, io->mean, io->resid, io->rsorted, io->msg, io->partials, io->counters + 1, io->stats)
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

// mode B: rows gathered in place from received slots (pitch vp), finished gradient rows stored through
// gptr (requester side) / coefficient read from the row itself (owner side)
struct GatherB { const float* rows; int vp; const int32_t* partner; float* const* gptr; };

static int launch_gather(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io_,
                         const float* table, int32_t unit_coef, vfmb_stream stream_, int smp = 0,
                         const GatherB* gb = nullptr) {
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
    if (!io->grow || !io->gws || !io->rsorted || !io->partials)
        return set_error(VFMB_EINVAL, "vfmb_sampled_gather: scratch required");
    const Layout& L = P.L; cudaStream_t stream = P.stream; const auto& cap = P.cap;
    float* gslot = (float*)io->partials + scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap).gslot_off;
    // F == 2: partner rows come from the sampled-row scratch; otherwise from `table`
    // (F > 2: the per-sample field sums written by k_score; unit_coef: received gradient rows)
    const float* tbl = table ? table : io->msg;
    if (cfg->F != 2 && !tbl) return set_error(VFMB_EINVAL, "vfmb_sampled_gather: table required for F != 2");
    const int N = cfg->B * cfg->F;
    int32_t* arrive = (int32_t*)((float*)io->partials + scratch_map(cfg->B, cfg->F, cfg->d, cap.u_cap).arrive_off);
    const float* vsp = gb && gb->rows && !unit_coef ? gb->rows : io->vs;
    const int vp = gb ? gb->vp : cfg->d;
    const int32_t* partner = gb && gb->partner ? gb->partner : plan->partner;
    float* const* gptr = gb ? gb->gptr : nullptr;
    const float* rs = io->rsorted;
    const GatherKnobs kn{tuning().gather_keep, tuning().gather_dyn, tuning().gather_fence};
    // one block per 512-position block tile, at most the resident blocks (further tiles come from the counter)
    const int64_t n_bt = ((int64_t)N + 511) / 512;
#define LAUNCH_GATHER(VEC, LPR, NV)                                                                          \
    do {                                                                                                     \
        if (unit_coef)                                                                                       \
            CUDA_TRY(launch_chained(k_gather<VEC, LPR, NV, 1>, grid_resident(k_gather<VEC, LPR, NV, 1>, n_bt * 8, 1), 256, 0, stream, \
                cfg->d, cfg->F, N, partner, plan->pos_rank, plan->urec, vsp, tbl, rs, gslot, io->grow, io->gws, arrive, \
                vp, gptr, kn));                                                                              \
        else                                                                                                 \
            CUDA_TRY(launch_chained(k_gather<VEC, LPR, NV, 0>, grid_resident(k_gather<VEC, LPR, NV, 0>, n_bt * 8, 1), 256, 0, stream, \
                cfg->d, cfg->F, N, partner, plan->pos_rank, plan->urec, vsp, tbl, rs, gslot, io->grow, io->gws, arrive, \
                vp, gptr, kn));                                                                              \
    } while (0)
    // The gather has no cross-lane arithmetic (every output element is a sequential sum over positions), so its
    // lane mapping is free: half as many lanes per row with two vectors each puts twice the positions behind
    // every warp instruction of the per-position bookkeeping, which is what bounds the kernel.
    if (L.vec == 4 && L.nv == 1 && L.lpr == 8 && tuning().gather_wide) LAUNCH_GATHER(4, 4, 2);
    else if (L.vec == 4 && L.nv == 1 && L.lpr == 16 && tuning().gather_wide) LAUNCH_GATHER(4, 8, 2);
    else if (L.vec == 4 && L.nv == 1 && L.lpr == 32 && tuning().gather_wide) LAUNCH_GATHER(4, 16, 2);
    else VFMB_LAYOUT_SWITCH(L, {
        LAUNCH_GATHER(VEC, LPR, NV);
    });
#undef LAUNCH_GATHER
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int vfmb_sampled_gather(const vfmb_config* cfg, const vfmb_plan* plan, const vfmb_step_io* io,
                                   const float* table, int32_t unit_coef, vfmb_stream stream) {
    VFMB_PHASE_S1("vfmb_sampled_gather");
    return launch_gather(cfg, plan, io, table, unit_coef, stream);
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
        k_scatter_resid<<<g, 256, 0, counted(stream)>>>(io->resid, plan->pos_of, N, cfg->F, io->rsorted);
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

// The fused training step: 4 launches -- k_stage<LEAN>, k_score, k_gather (the group that stores the
// last partial of a row cut by tile boundaries finishes it), k_adam_rows<FLAVOR 2> (KL, scalar
// parameters and the step counter folded in).
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
    // (a one-pass k_gather_score existed for F == 2 until the gather went hierarchical: the pair below then won,
    //  98.9 vs 105.4 us per ml20m step with cached plans, and the fused kernel was removed)
    rc = launch_score(cfg, tab, plan, io, stream, 1);
    if (rc) return rc;
    return backward_impl(cfg, tab, plan, io, adam, VFMB_ADAM_TOUCHED, 1.0f, 2, stream);
}

// ------------------------------------------------------------------------------- mode B, fused
// The phases of the row-sharded step with the exchanges written / read in place (NVLink peer memory):
//   owner      vfmb_shard_stage_put     k_stage, the sampled row stored into every requester's slot
//   requester  vfmb_shard_score         k_score on the received slots + the rank's additive scalars put
//                                       into every rank's tail region by the block that finishes last
//   requester  vfmb_shard_gather_put    k_gather on the received slots, finished gradient rows stored
//                                       into their owners' slots
//   owner      vfmb_shard_owner_update  k_adam_rows_pipe<3>: a row's gradient = its requesters' slots added
//                                       in source-rank order, Adam on the owned rows, scalar parameters
//                                       and loss from the ranks' tail slots
// Two cross-rank barriers remain per step (after the rows, after the gradients); the caller issues them.
static int make_peers_tab(const void* const* peers, int32_t P, int32_t rank, Peers* out) {
    Peers pe{};
    if (!peers || P < 1 || P > kMaxFields || rank < 0 || rank >= P) return set_error(VFMB_EINVAL, "peer table: bad P / rank");
    for (int q = 0; q < P; ++q) {
        if (!peers[q]) return set_error(VFMB_EINVAL, "peer table: null buffer of rank %d", q);
        pe.base[q] = const_cast<void*>(peers[q]);
    }
    pe.n = P; pe.rank = rank;
    *out = pe;
    return 0;
}

extern "C" int vfmb_shard_stage_put(const vfmb_config* cfg_o, const vfmb_tables* tab, const vfmb_plan* plan_o,
                                    const vfmb_step_io* io_o, int32_t CAP, int32_t SP, int32_t n_real,
                                    const void* const* peers_rows, int32_t P, int32_t rank, vfmb_stream stream) {
    if (!cfg_o || !plan_o || CAP < 1 || SP < cfg_o->d + 1 || (SP & 3)) return set_error(VFMB_EINVAL, "vfmb_shard_stage_put: bad argument");
    if (cfg_o->S != 1) return set_error(VFMB_ESHAPE, "vfmb_shard_stage_put: S = 1 only");
    RowPut put{};
    int rc = make_peers_tab(peers_rows, P, rank, &put.pe);
    if (rc) return rc;
    put.occ = plan_o->occ; put.CAP = CAP; put.SP = SP; put.n_real = n_real;
    return launch_stage(cfg_o, tab, plan_o, io_o, stream, false, 0, &put);
}

extern "C" int vfmb_shard_score(const vfmb_config* cfg_l, const vfmb_tables* tab, const vfmb_plan* plan_l,
                                const vfmb_step_io* io_l, const float* recv_rows, int32_t SP, const int32_t* inv_slot,
                                const void* const* peers_tail, int32_t P, int32_t rank, const float* stats_owner,
                                const int32_t* overflow, float n_local, int32_t tail_pitch, vfmb_stream stream) {
    if (!cfg_l || !recv_rows || !inv_slot || !stats_owner || tail_pitch < VFMB_DP_TAIL)
        return set_error(VFMB_EINVAL, "vfmb_shard_score: bad argument");
    if (cfg_l->S != 1) return set_error(VFMB_ESHAPE, "vfmb_shard_score: S = 1 only");
    ScoreB sb{};
    sb.rows = recv_rows; sb.vp = SP; sb.inv_slot = inv_slot;
    int rc = make_peers_tab(peers_tail, P, rank, &sb.tp.pe);
    if (rc) return rc;
    sb.tp.stats_owner = stats_owner; sb.tp.overflow = overflow; sb.tp.n_local = n_local; sb.tp.pitch = tail_pitch;
    return launch_score(cfg_l, tab, plan_l, io_l, stream, 0, &sb);
}

extern "C" int vfmb_shard_gather_put(const vfmb_config* cfg_l, const vfmb_plan* plan_l, const vfmb_step_io* io_l,
                                     const float* recv_rows, int32_t SP, const int32_t* partner_slot,
                                     float* const* gptr, const int32_t* own_slot, vfmb_stream stream) {
    if (!cfg_l || !recv_rows || !partner_slot || !gptr || !own_slot) return set_error(VFMB_EINVAL, "vfmb_shard_gather_put: bad argument");
    if (cfg_l->S != 1) return set_error(VFMB_ESHAPE, "vfmb_shard_gather_put: S = 1 only");
    GatherB gb{};
    gb.rows = recv_rows; gb.vp = SP; gb.partner = partner_slot; gb.gptr = gptr;
    return launch_gather(cfg_l, plan_l, io_l, nullptr, 0, stream, 0, &gb);
}

extern "C" int vfmb_shard_owner_update(const vfmb_config* cfg_o, const vfmb_tables* tab, const vfmb_plan* plan_o,
                                       const vfmb_step_io* io_o, const vfmb_adam* adam, const float* recv_grads,
                                       int32_t SP, int32_t n_real, const float* tail_slots, int32_t P, int32_t tail_pitch,
                                       int32_t B_global, float n_train_global, float* stats_out,
                                       const float* eps_global, vfmb_stream stream) {
    if (!cfg_o || !recv_grads || !tail_slots || !stats_out || P < 1 || !plan_o || !plan_o->occ)
        return set_error(VFMB_EINVAL, "vfmb_shard_owner_update: bad argument");
    if (cfg_o->S != 1) return set_error(VFMB_ESHAPE, "vfmb_shard_owner_update: S = 1 only");
    // one kernel: a row's gradient is the sum of the <= P slots its requesters stored, taken in source-rank
    // order straight from the received region (no separate ordered-sum pass over the slots)
    DpTail dp{};
    dp.tail_slots = tail_slots; dp.P = P; dp.pitch = tail_pitch; dp.B_global = B_global; dp.n_train_global = n_train_global;
    dp.stats_out = stats_out; dp.eps_global = eps_global;
    dp.recv_grads = recv_grads; dp.slot_pitch = SP; dp.n_real = n_real;
    return launch_adam(cfg_o, tab, plan_o, io_o, adam, VFMB_ADAM_TOUCHED, 1.0f, 3, stream, &dp);
}

extern "C" int vfmb_philox_normals(const vfmb_config* cfg, const int32_t* uniq, int32_t U, int32_t step,
                                   float* eps_global, float* eps_bias, float* eps_entity, vfmb_stream stream_) {
    if (!cfg || !uniq || !eps_global || !eps_bias || !eps_entity || U < 0) return set_error(VFMB_EINVAL, "vfmb_philox_normals: bad argument");
    if (U == 0) return 0;
    DevCfg dc = make_dev(cfg);
    int vec = (cfg->d % 4 == 0) ? 4 : 1;
    if (cfg->S < 1) return set_error(VFMB_ESHAPE, "vfmb_philox_normals: S=%d", cfg->S);
    int64_t work = (int64_t)U * ((cfg->d + vec - 1) / vec) * cfg->S;
    int grid = (int)((work + 255) / 256 > 4096 ? 4096 : (work + 255) / 256);
    if (vec == 4) k_philox_export<4><<<grid, 256, 0, counted((cudaStream_t)stream_)>>>(dc, uniq, U, (uint32_t)step, eps_global, eps_bias, eps_entity);
    else k_philox_export<1><<<grid, 256, 0, counted((cudaStream_t)stream_)>>>(dc, uniq, U, (uint32_t)step, eps_global, eps_bias, eps_entity);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
