// chain_kernel.cuh -- the persistent MC-chain kernel (one CTA per repetition).
//
// Replaces McSAS.mcFit (src/mcsas/mcsas/mcsas.py:287-439) including the retry loop of
// McSAS.analyse (:214-246).  Per step the reference draws a proposal for contribution
// ri (:358), evaluates new and old contribution intensity over all q (:360-362), forms
// ft - old + new (:366), solves scale/background (backgroundscalingfit.py:112-139) and
// accepts on strict chi-squared improvement (:379).
//
// B200 mapping
//  * data (q, I, sigma) is staged once with a TMA bulk copy (cp.async.bulk + mbarrier)
//    into shared memory; each thread then keeps its QPT q points, their weights and the
//    running model intensity ft in REGISTERS for the whole chain.
//  * the proposal stream is state independent (one draw per active parameter per step,
//    scatteringmodel.py:117-127), so a batch of K consecutive steps is evaluated
//    speculatively against the current state; the first improving one is committed,
//    later ones are discarded and re-evaluated.  This is bit-identical to the
//    sequential chain and turns the chain's latency-bound step into throughput work.
//  * LSQ sums (sum w M, sum w M^2, sum w M I) are accumulated in fp64, reduced with warp
//    shuffles + one shared-memory hop; chi-squared follows in closed form.
#pragma once
#include "models.cuh"
#include "philox.cuh"

#define KCAP MCSAS_MAX_SPEC
#define MAX_WARPS 16

// ---- small PTX helpers: mbarrier + 1-D TMA bulk copy -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Closed-form weighted LSQ for (scale a, background b) and reduced chi-squared
// (backgroundscalingfit.py:112-139; leastsq on a linear residual == normal equations).
struct LsqConst {
    double S, Si, Sii, inv_nq;
    int find_bg, pos_bg;
};
__device__ __forceinline__ void lsq_solve(const LsqConst &L, double Sm, double Smm, double Smi,
                                          double &a, double &b, double &chi)
{
    if (L.find_bg) {
        double det = L.S * Smm - Sm * Sm;
        a = (L.S * Smi - Sm * L.Si) / det;
        b = (Smm * L.Si - Sm * Smi) / det;
        if (L.pos_bg && b < 0.) { a = Smi / Smm; b = 0.; }   // |b| residual (:60-63)
    } else {
        a = Smi / Smm; b = 0.;                               // chiNoBg (:66-70)
    }
    chi = (L.Sii - 2. * a * Smi - 2. * b * L.Si + a * a * Smm + 2. * a * b * Sm + b * b * L.S)
          * L.inv_nq;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <class Real> __device__ __forceinline__ double model_sum(double ft, Real inew, Real iold);
// fp64: reference order (ft - old) + new (mcsas.py:366)
template <> __device__ __forceinline__ double model_sum<double>(double ft, double inew, double iold)
{
    return (ft - iold) + inew;
}
template <> __device__ __forceinline__ double model_sum<float>(double ft, float inew, float iold)
{
    return ft + (double)(inew - iold);
}

template <int MODEL, class Real, int QPT, int MAXT>
__global__ void __launch_bounds__(MAXT) chain_kernel(const __grid_constant__ ChainArgs a)
{
    using O = Ops<MODEL, Real>;
    constexpr int NARR = O::NARR;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5,
              nwarp = T >> 5;
    const int chain = blockIdx.x;
    const uint32_t gchain = (uint32_t)(a.first_chain + chain);
    const int set = chain / a.n_reps;
    const DatasetDesc ds = a.sets[set];
    const int nq = ds.nq, N = a.n_contribs, P = a.mp.n_active;
    const int D = NARR ? a.mp.quad_nodes : 0;

    // ---- shared memory carve-up -----------------------------------------------------
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    auto carve = [&](size_t bytes) { unsigned char *p = sp; sp += (bytes + 15) & ~size_t(15); return p; };
    uint64_t *s_bar = (uint64_t *)carve(16);
    double *s_data = (double *)carve(sizeof(double) * 3 * ds.nq_pad);
    double *s_rset = (double *)carve(sizeof(double) * N * P);
    double *s_prop = (double *)carve(sizeof(double) * KCAP * MCSAS_MAX_ACTIVE);
    double *s_red = (double *)carve(sizeof(double) * MAX_WARPS * KCAP * 3);
    double *s_chi = (double *)carve(sizeof(double) * KCAP * 3);
    double *s_misc = (double *)carve(sizeof(double) * 8);
    Contrib<Real> *s_cc = (Contrib<Real> *)carve(sizeof(Contrib<Real>) * KCAP * 2);
    Real *s_qx = (Real *)carve(sizeof(Real) * (D + 1));
    Real *s_qw = (Real *)carve(sizeof(Real) * (D + 1));
    Real *s_nodes = (Real *)sp;     // [kcap][2][NARR][D], sized by the host
    int *s_imisc = (int *)(s_misc + 6);

    // ---- stage the dataset once: TMA bulk copy global -> shared ----------------------
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = (uint32_t)(sizeof(double) * 3 * ds.nq_pad);
        mbar_expect_tx(s_bar, bytes);
        tma_bulk_g2s(s_data, a.data + ds.offset, bytes, s_bar);
    }
    for (int j = tid; j < D; j += T) {
        s_qx[j] = (Real)a.quad_x[j];
        s_qw[j] = (Real)a.quad_w[j];
    }
    mbar_wait(s_bar, 0);
    const QuadTab<Real> qt = { D, a.mp.quad_rule, s_qx, s_qw };

    // ---- per-thread resident state ---------------------------------------------------
    QRep<Real> Q[QPT];
    double w[QPT], wI[QPT], ft[QPT];
    double pS = 0., pSi = 0., pSii = 0.;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        int i = tid + j * T;
        bool valid = i < nq;
        int ii = valid ? i : nq - 1;
        double qv = s_data[ii], Iv = s_data[ds.nq_pad + ii], ev = s_data[2 * ds.nq_pad + ii];
        if (ev == 0.0) ev = 1.;                                  // backgroundscalingfit.py:117
        Q[j].set(qv);
        w[j] = valid ? 1. / (ev * ev) : 0.;
        wI[j] = w[j] * Iv;
        pS += w[j]; pSi += wI[j]; pSii += wI[j] * Iv;
    }
    pS = warp_sum(pS); pSi = warp_sum(pSi); pSii = warp_sum(pSii);
    if (lane == 0) { s_red[warp * 3] = pS; s_red[warp * 3 + 1] = pSi; s_red[warp * 3 + 2] = pSii; }
    __syncthreads();
    LsqConst L;
    L.S = 0.; L.Si = 0.; L.Sii = 0.;
    for (int wv = 0; wv < nwarp; ++wv) {
        L.S += s_red[wv * 3]; L.Si += s_red[wv * 3 + 1]; L.Sii += s_red[wv * 3 + 2];
    }
    L.inv_nq = 1. / (double)nq;
    L.find_bg = a.find_bg; L.pos_bg = a.pos_bg;
    __syncthreads();

    // ---- fp32 amplitude scale G: 1 / amplitude at the upper end of the active ranges ---
    double G = 1.;
    if (sizeof(Real) == 4) {
        if (tid == 0) {
            double p[MCSAS_MAX_PARAMS], pv[MCSAS_MAX_ACTIVE];
            for (int k = 0; k < P; ++k) pv[k] = a.mp.hi[k];
            patch_params(a.mp, pv, p);
            Contrib<Real> tmp;
            VWS o = O::prep(p, a.comp_exp, 1., qt, tmp, s_nodes);
            double g = 1. / o.amp_raw;
            if (!(g > 0.) || isinf(g)) g = 1.;
            s_misc[0] = g;
        }
        __syncthreads();
        G = s_misc[0];
        __syncthreads();
    }

    // block-wide LSQ of the resident ft: every thread gets (a, b, chi)
    auto lsq_block = [&](double &sa, double &sb, double &chi) {
        double sm = 0., smm = 0., smi = 0.;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            double t = w[j] * ft[j];
            sm += t; smm = fma(t, ft[j], smm); smi = fma(wI[j], ft[j], smi);
        }
        sm = warp_sum(sm); smm = warp_sum(smm); smi = warp_sum(smi);
        if (lane == 0) { s_red[warp * 3] = sm; s_red[warp * 3 + 1] = smm; s_red[warp * 3 + 2] = smi; }
        __syncthreads();
        sm = 0.; smm = 0.; smi = 0.;
        for (int wv = 0; wv < nwarp; ++wv) {
            sm += s_red[wv * 3]; smm += s_red[wv * 3 + 1]; smi += s_red[wv * 3 + 2];
        }
        lsq_solve(L, sm, smm, smi, sa, sb, chi);
        __syncthreads();
    };

    const int kcap = a.spec_max;        // host: 1..KCAP, already limited by shared memory
    const unsigned long long t_start = globaltimer_ns();
    long long step = 0, moves = 0, evals = 0;
    double conval = 0., conval0 = 0., sc_a = 1., sc_b = 0.;
    int attempt = 0;
    bool stopped = false;

    for (;; ++attempt) {
        // ---- initial contribution set (mcsas.py:310-317) -----------------------------
        for (int i = tid; i < N; i += T) {
            uint32_t o4[4];
            if (!a.start_min && !a.stream) philox4x32_10(a.seed, (uint64_t)i, gchain, 2u * attempt + 1u, o4);
            for (int k = 0; k < P; ++k) {
                double lo = a.mp.lo[k], hi = a.mp.hi[k], val;
                if (a.start_min) {
                    double mb = lo < hi ? lo : hi;
                    if (mb == 0.) mb = M_PI / ds.q_max;
                    val = 1. * mb * .5;
                } else {
                    double u = a.stream ? a.stream[(long long)chain * a.stream_stride + (long long)k * N + i]
                                        : generator_output(o4[k], a.mp.gen_kind[k]);
                    val = u * (hi - lo) + lo;                    // parameter.py:83-84
                }
                s_rset[i * P + k] = val;
            }
        }
#pragma unroll
        for (int j = 0; j < QPT; ++j) ft[j] = 0.;
        __syncthreads();
        // ---- ft = sum of all contributions, in contribution order (mcsas.py:319) -----
        for (int i0 = 0; i0 < N; i0 += kcap) {
            int nb = min(kcap, N - i0);
            if (tid < nb) {
                double p[MCSAS_MAX_PARAMS];
                patch_params(a.mp, s_rset + (i0 + tid) * P, p);
                O::prep(p, a.comp_exp, G, qt, s_cc[tid * 2], s_nodes + (size_t)(tid * 2) * NARR * D);
            }
            __syncthreads();
            for (int k = 0; k < nb; ++k) {
                const Contrib<Real> cc = s_cc[k * 2];
                const Real *nd = s_nodes + (size_t)(k * 2) * NARR * D;
#pragma unroll
                for (int j = 0; j < QPT; ++j) ft[j] += (double)O::eval(cc, nd, qt, Q[j]);
            }
            __syncthreads();
        }
        evals += N;
        lsq_block(sc_a, sc_b, conval);                           // mcsas.py:339-344
        conval0 = conval;
        step = 0; moves = 0;
        int ri = 0, K = 1;
        unsigned batch = 0;
        if (tid == 0) s_imisc[2] = 0;
        const long long sbase = (long long)chain * a.stream_stride + (a.start_min ? 0 : (long long)N * P);

        // ---- the MC loop (mcsas.py:354-404), K speculative steps per batch ------------
        while (N > 1 && conval > a.conv_crit && step < a.max_iter) {
            if (a.stop_flag && tid == 0 && (batch & 63) == 0) s_imisc[2] = *a.stop_flag;
            ++batch;
            int Kb = (int)min((long long)min(K, kcap), a.max_iter - step);
            if (tid < Kb) {
                long long t = step + tid;
                int rix = (ri + tid) % N;
                double pv[MCSAS_MAX_ACTIVE], p[MCSAS_MAX_PARAMS];
                uint32_t o4[4];
                if (!a.stream) philox4x32_10(a.seed, (uint64_t)t, gchain, 2u * attempt, o4);
                for (int k = 0; k < P; ++k) {
                    double u = a.stream ? a.stream[sbase + t * P + k]
                                        : generator_output(o4[k], a.mp.gen_kind[k]);
                    pv[k] = u * (a.mp.hi[k] - a.mp.lo[k]) + a.mp.lo[k];      // mcsas.py:358
                    s_prop[tid * MCSAS_MAX_ACTIVE + k] = pv[k];
                }
                patch_params(a.mp, pv, p);
                O::prep(p, a.comp_exp, G, qt, s_cc[tid * 2], s_nodes + (size_t)(tid * 2) * NARR * D);
                patch_params(a.mp, s_rset + rix * P, p);
                O::prep(p, a.comp_exp, G, qt, s_cc[tid * 2 + 1],
                        s_nodes + (size_t)(tid * 2 + 1) * NARR * D);
            }
            __syncthreads();
            if (a.stop_flag && s_imisc[2]) { stopped = true; break; }
            for (int k = 0; k < Kb; ++k) {
                const Contrib<Real> cn = s_cc[k * 2], co = s_cc[k * 2 + 1];
                const Real *nn = s_nodes + (size_t)(k * 2) * NARR * D, *no = nn + NARR * D;
                double sm = 0., smm = 0., smi = 0.;
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                    Real inew = O::eval(cn, nn, qt, Q[j]);       // mcsas.py:360
                    Real iold = O::eval(co, no, qt, Q[j]);       // mcsas.py:362
                    double M = model_sum<Real>(ft[j], inew, iold);   // :366
                    double t = w[j] * M;
                    sm += t; smm = fma(t, M, smm); smi = fma(wI[j], M, smi);
                }
                sm = warp_sum(sm); smm = warp_sum(smm); smi = warp_sum(smi);
                if (lane == 0) {
                    double *r = s_red + (warp * KCAP + k) * 3;
                    r[0] = sm; r[1] = smm; r[2] = smi;
                }
            }
            __syncthreads();
            if (warp == 0) {
                double chi = 0., sa = 0., sb = 0.;
                bool better = false;
                if (lane < Kb) {
                    double sm = 0., smm = 0., smi = 0.;
                    for (int wv = 0; wv < nwarp; ++wv) {
                        const double *r = s_red + (wv * KCAP + lane) * 3;
                        sm += r[0]; smm += r[1]; smi += r[2];
                    }
                    lsq_solve(L, sm, smm, smi, sa, sb, chi);     // mcsas.py:376
                    better = chi < conval;                       // :379 strict
                    s_chi[lane * 3] = chi; s_chi[lane * 3 + 1] = sa; s_chi[lane * 3 + 2] = sb;
                }
                unsigned mask = __ballot_sync(0xffffffffu, better);
                if (lane == 0) s_imisc[0] = mask ? (__ffs(mask) - 1) : -1;
            }
            __syncthreads();
            const int kstar = s_imisc[0];
            const int used = kstar >= 0 ? kstar + 1 : Kb;
            if (a.trace_accept) {
                for (int k = tid; k < used; k += T) {
                    a.trace_accept[(long long)chain * a.max_iter + step + k] = (k == kstar);
                    a.trace_chi[(long long)chain * a.max_iter + step + k] = s_chi[k * 3];
                }
            }
            if (kstar >= 0) {                                    // mcsas.py:381-384
                const Contrib<Real> cn = s_cc[kstar * 2], co = s_cc[kstar * 2 + 1];
                const Real *nn = s_nodes + (size_t)(kstar * 2) * NARR * D, *no = nn + NARR * D;
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                    Real inew = O::eval(cn, nn, qt, Q[j]);
                    Real iold = O::eval(co, no, qt, Q[j]);
                    ft[j] = model_sum<Real>(ft[j], inew, iold);
                }
                conval = s_chi[kstar * 3]; sc_a = s_chi[kstar * 3 + 1]; sc_b = s_chi[kstar * 3 + 2];
                int rix = (ri + kstar) % N;
                if (tid < P) s_rset[rix * P + tid] = s_prop[kstar * MCSAS_MAX_ACTIVE + tid];
                ++moves;
                evals += 2;
                K = max(1, min(kcap, 1 << (31 - __clz(kstar + 1))));
            } else {
                K = min(kcap, K * 2);
            }
            evals += 2 * Kb;
            step += used;
            ri = (int)((ri + used) % N);                         // :403-404
            __syncthreads();
        }
        // ---- final fit of the accepted state (mcsas.py:424-425) ------------------------
        lsq_block(sc_a, sc_b, conval);
        if (conval <= a.conv_crit || a.stream || stopped || attempt + 1 >= a.max_attempts
            || N <= 1)
            break;
    }

    // ---- results ------------------------------------------------------------------------
    for (int i = tid; i < N * P; i += T) a.rset[(long long)chain * N * P + i] = s_rset[i];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        int i = tid + j * T;
        if (i < nq)
            a.fit_meas[(long long)chain * a.nq_stride + i] = ft[j] * sc_a + sc_b;   // mcsas.py:429
    }
    if (tid == 0) {
        // internal amplitudes are w*G: scaling in the reference's (SI) units is a*G
        a.scaling[chain] = sc_a * G;
        a.background[chain] = sc_b;
        a.conval[chain] = conval;
        a.conval0[chain] = conval0;
        a.num_iter[chain] = step;
        a.num_moves[chain] = moves;
        a.evals[chain] = evals;
        a.attempts[chain] = attempt + 1;
        a.elapsed_s[chain] = 1e-9 * (double)(globaltimer_ns() - t_start);
    }
}

// ---- ScatteringModel.calc for single contributions (scatteringmodel.py:79-105) --------
template <int MODEL, class Real>
__global__ void __launch_bounds__(128) eval_kernel(const __grid_constant__ EvalArgs a)
{
    using O = Ops<MODEL, Real>;
    constexpr int NARR = O::NARR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = NARR ? a.mp.quad_nodes : 0;
    Real *s_qx = (Real *)smem_raw;
    Real *s_qw = s_qx + D + 1;
    Real *s_nodes = s_qw + D + 1;
    __shared__ Contrib<Real> s_cc;
    __shared__ double s_G;
    const int tid = threadIdx.x, c = blockIdx.x;
    const DatasetDesc ds = a.sets[a.set];
    for (int j = tid; j < D; j += blockDim.x) {
        s_qx[j] = (Real)a.quad_x[j];
        s_qw[j] = (Real)a.quad_w[j];
    }
    __syncthreads();
    const QuadTab<Real> qt = { D, a.mp.quad_rule, s_qx, s_qw };
    if (tid == 0) {
        double p[MCSAS_MAX_PARAMS];
        patch_params(a.mp, a.pset + (long long)c * a.mp.n_active, p);
        double G = 1.;
        if (sizeof(Real) == 4) {
            Contrib<Real> tmp;
            VWS o0 = O::prep(p, a.comp_exp, 1., qt, tmp, s_nodes);
            G = 1. / o0.amp_raw;                 // this contribution's own amplitude -> O(1)
            if (!(G > 0.) || isinf(G)) G = 1.;
        }
        VWS o = O::prep(p, a.comp_exp, G, qt, s_cc, s_nodes);
        s_G = G;
        a.v[c] = o.v; a.w[c] = o.w; a.s[c] = o.s;
    }
    __syncthreads();
    const double *q = a.data + ds.offset;
    for (int i = tid; i < ds.nq; i += blockDim.x) {
        QRep<Real> Q;
        Q.set(q[i]);
        a.intensity[(long long)c * ds.nq + i] = (double)O::eval(s_cc, s_nodes, qt, Q) / s_G;
    }
}
