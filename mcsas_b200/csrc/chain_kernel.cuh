// chain_kernel.cuh -- the persistent MC-chain kernel.
//
// Replaces McSAS.mcFit (src/mcsas/mcsas/mcsas.py:287-439) including the retry loop of
// McSAS.analyse (:214-246).  Per step the reference draws a proposal for contribution
// ri (:358), evaluates new and old contribution intensity over all q (:360-362), forms
// ft - old + new (:366), solves scale/background (backgroundscalingfit.py:112-139) and
// accepts on strict chi-squared improvement (:379).
//
// B200 mapping
//  * one thread-block CLUSTER of C CTAs per chain (C = 1,2,4,8, chosen by the host from
//    the number of live chains): the q axis is split over the cluster; every eval thread
//    keeps its QPT q points, their LSQ weights and the running model intensity ft in
//    REGISTERS for the whole launch.  Data is staged once by a TMA bulk copy.
//  * speculative batches: the proposal stream is state independent (one draw per active
//    parameter per step, scatteringmodel.py:117-127), so K consecutive steps are evaluated
//    against the current state; the first improving one is committed, later ones are
//    discarded and re-evaluated -- bit-identical to the sequential chain.
//  * warp specialisation: one PRODUCER warp per CTA runs ahead of the chain and fills a
//    shared-memory ring with (proposal, per-contribution constants of the new and of the
//    replaced contribution) for future steps (Philox + pow + node geometry), off the
//    critical path of the eval warps.
//  * LSQ sums (sum w M, sum w M^2, sum w M I) are accumulated in fp64; 8 proposals are
//    reduced at once with a select-free shuffle reduce-scatter (lanes evaluate the group
//    in a lane-permuted order), then over warps through shared memory and over the
//    cluster through distributed shared memory (one cluster barrier per batch); every
//    CTA then takes the identical accept decision redundantly.
//  * time slicing: a launch runs every chain for at most `slice_steps` MC steps and saves
//    (rset, counters); the host relaunches the unfinished chains with a larger cluster, so
//    SMs freed by converged chains shorten the tail of the slow ones.
#pragma once
#include "models.cuh"
#include "philox.cuh"

#define GRP 8                 // proposals per reduce-scatter group

// ---- small PTX helpers: mbarrier + 1-D TMA bulk copy, cluster, named barriers -----------
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
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// store a double into the same shared-memory offset of CTA `rank` of this cluster (DSMEM)
__device__ __forceinline__ void dsmem_store(double *local_ptr, uint32_t rank, double v)
{
    uint32_t la = smem_u32(local_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
// store a double into CTA `rank` and signal 8 bytes on that CTA's mbarrier (st.async)
__device__ __forceinline__ void dsmem_store_async(double *local_ptr, uint64_t *local_bar, uint32_t rank,
                                                  double v)
{
    uint32_t la = smem_u32(local_ptr), lb = smem_u32(local_bar), ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(ra), "l"(__double_as_longlong(v)), "r"(rb) : "memory");
}
__device__ __forceinline__ long long ld_volatile(const long long *p) { return *(const volatile long long *)p; }
__device__ __forceinline__ int ld_volatile_i(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void st_volatile(long long *p, long long v) { *(volatile long long *)p = v; }
__device__ __forceinline__ void st_volatile_i(int *p, int v) { *(volatile int *)p = v; }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    __syncwarp();
    asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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
        double inv = 1. / (L.S * Smm - Sm * Sm);
        a = (L.S * Smi - Sm * L.Si) * inv;
        b = (Smm * L.Si - Sm * Smi) * inv;
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

template <class Real> struct RingEntry {
    double prop[MCSAS_MAX_ACTIVE];
    Contrib<Real> cn, co;        // new contribution, replaced contribution
};

struct ChainCtrl {               // eval warps <-> producer warp (volatile accesses)
    long long step;              // first step of the current batch (published by eval thread 0)
    long long filled;            // ring holds steps [.., filled) (published by the producer)
    int done;                    // 1: leave the MC loop, 2: watchdog
    int batch;
};

template <int MODEL, class Real, int QPT>
__global__ void __launch_bounds__(288) chain_kernel(const __grid_constant__ ChainArgs a)
{
    using O = Ops<MODEL, Real>;
    constexpr int NARR = O::NARR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = blockDim.x - 32;                 // eval threads; the last warp produces
    const int W = T >> 5;
    const bool producer = tid >= T;
    const int C = a.cluster;
    const int rank = C > 1 ? (int)cluster_ctarank() : 0;
    const int slot = blockIdx.x / C;
    const int chain = a.chain_ids ? a.chain_ids[slot] : slot;
    const uint32_t gchain = (uint32_t)(a.first_chain + chain);
    const int set = chain / a.n_reps;
    const DatasetDesc ds = a.sets[set];
    const int nq = ds.nq, N = a.n_contribs, P = a.mp.n_active;
    const int D = NARR ? a.mp.quad_nodes : 0;
    const int RING = a.ring;                       // min(RING_MAX, N) (host)
    const int kmax = a.spec_max;                   // <= RING / 4 (host), >= 1

    // ---- shared memory carve-up (keep in sync with chain_smem_bytes) ----------------------
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    auto carve = [&](size_t bytes) { unsigned char *p = sp; sp += (bytes + 15) & ~size_t(15); return p; };
    uint64_t *s_bar = (uint64_t *)carve(16);
    double *s_data = (double *)carve(sizeof(double) * 3 * ds.nq_pad);
    double *s_rset = (double *)carve(sizeof(double) * N * (P > 0 ? P : 1));
    ChainCtrl *s_ctrl = (ChainCtrl *)carve(64);
    RingEntry<Real> *s_ring = (RingEntry<Real> *)carve((size_t)RING * sizeof(RingEntry<Real>));
    double *s_wsum = (double *)carve(sizeof(double) * (MAX_EVAL_WARPS + 1) * KCAP * 3);
    double *s_part = (double *)carve(sizeof(double) * 2 * MAX_CLUSTER * KCAP * 3);
    double *s_xflag = (double *)carve(sizeof(double) * 2 * MAX_CLUSTER);
    uint64_t *s_xbar = (uint64_t *)carve(16);      // [2] exchange mbarriers (tx-count based)
    double *s_chi = (double *)carve(sizeof(double) * KCAP * 3 + 64);
    int *s_dec = (int *)(s_chi + KCAP * 3);
    Real *s_qx = (Real *)carve(sizeof(Real) * (D + 1));
    Real *s_qw = (Real *)carve(sizeof(Real) * (D + 1));
    Real *s_nodes = (Real *)sp;                    // [RING][2][NARR][D]
    const size_t NSTR = (size_t)NARR * D;          // node reals per contribution

    // ---- stage the dataset once: TMA bulk copy global -> shared ---------------------------
    if (tid == 0) { mbar_init(s_bar, 1); mbar_init(s_xbar, 1); mbar_init(s_xbar + 1, 1); s_dec[1] = 0; }
    unsigned xphase = 0;                           // bit b = phase parity of s_xbar[b]
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = (uint32_t)(sizeof(double) * 3 * ds.nq_pad);
        mbar_expect_tx(s_bar, bytes);
        tma_bulk_g2s(s_data, a.data + ds.offset, bytes, s_bar);
    }
    for (int j = tid; j < D; j += blockDim.x) {
        s_qx[j] = (Real)a.quad_x[j];
        s_qw[j] = (Real)a.quad_w[j];
    }
    mbar_wait(s_bar, 0);
    __syncthreads();
    const QuadTab<Real> qt = { D, a.mp.quad_rule, s_qx, s_qw, &a.smear, a.smear.n };

    // ---- per-thread resident state: q point (j*C + rank)*T + tid -------------------------
    constexpr bool TWO_PASS = sizeof(Real) == 8;   // parity mode: residual-form chi-squared
    constexpr bool FAST = sizeof(Real) == 4;       // production: MUFU-based contribution amplitudes
    QRep<Real> Q[QPT];
    double w[QPT], wI[QPT], ft[QPT], Iq[QPT];
    double v3[3] = {0., 0., 0.};
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        int i = (j * C + rank) * T + tid;
        bool valid = !producer && i < nq;
        int ii = (i < nq) ? i : nq - 1;
        double qv = s_data[ii], Iv = s_data[ds.nq_pad + ii], ev = s_data[2 * ds.nq_pad + ii];
        if (ev == 0.0) ev = 1.;                                  // backgroundscalingfit.py:117
        Q[j].set(qv);
        w[j] = valid ? 1. / (ev * ev) : 0.;
        wI[j] = w[j] * Iv;
        Iq[j] = Iv;
        ft[j] = 0.;
        v3[0] += w[j]; v3[1] += wI[j]; v3[2] += wI[j] * Iv;
    }

    // Sum three per-thread partials over the whole cluster; every thread of every CTA gets
    // the identical totals (fixed summation order).  Called by ALL threads of the CTA.
    int xbuf = 0;
    auto cluster_allsum3 = [&](double (&v)[3]) {
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = warp_sum(v[i]);
        if (lane == 0) { s_wsum[warp * 3] = v[0]; s_wsum[warp * 3 + 1] = v[1]; s_wsum[warp * 3 + 2] = v[2]; }
        __syncthreads();
        double *part = s_part + (size_t)xbuf * MAX_CLUSTER * KCAP * 3;
        if (tid < 3) {
            double x = 0.;
            for (int wv = 0; wv <= W; ++wv) x += s_wsum[wv * 3 + tid];
            if (C > 1) { for (int r = 0; r < C; ++r) dsmem_store(part + rank * 3 + tid, (uint32_t)r, x); }
            else part[tid] = x;
        }
        if (C > 1) cluster_sync_all(); else __syncthreads();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double x = 0.;
            for (int r = 0; r < C; ++r) x += part[r * 3 + i];
            v[i] = x;
        }
        xbuf ^= 1;
    };

    cluster_allsum3(v3);
    LsqConst L;
    L.S = v3[0]; L.Si = v3[1]; L.Sii = v3[2];
    L.inv_nq = 1. / (double)nq;
    L.find_bg = a.find_bg; L.pos_bg = a.pos_bg;

    // ---- fp32 amplitude scale G: 1 / amplitude at the upper end of the active ranges ------
    double G = 1.;
    if (sizeof(Real) == 4) {
        if (tid == 0) {
            PVec p(a.mp, a.mp.hi);
            Contrib<Real> tmp;
            VWS o = O::template prep<false>(p, a.comp_exp, 1., qt, tmp, s_nodes);
            double g = 1. / o.amp_raw;
            if (!(g > 0.) || isinf(g)) g = 1.;
            s_chi[0] = g;
        }
        __syncthreads();
        G = s_chi[0];
        __syncthreads();
    }

    // LSQ of the resident ft over the cluster: every thread gets (a, b, chi)
    auto lsq_all = [&](double &sa, double &sb, double &chi) {
        double v[3] = {0., 0., 0.};
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            double t = w[j] * ft[j];
            v[0] += t; v[1] = fma(t, ft[j], v[1]); v[2] = fma(wI[j], ft[j], v[2]);
        }
        cluster_allsum3(v);
        lsq_solve(L, v[0], v[1], v[2], sa, sb, chi);
        if (TWO_PASS) {          // chiSqr as the reference forms it (backgroundscalingfit.py:73-77,:135)
            double r2[3] = {0., 0., 0.};
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                double r = Iq[j] - (ft[j] * sa + sb);
                r2[0] = fma(w[j] * r, r, r2[0]);
            }
            cluster_allsum3(r2);
            chi = r2[0] * L.inv_nq;
        }
    };

    // One ring entry: proposal of MC step t and the constants of the new / replaced contribution
    auto fill_entry = [&](long long t, int attempt, long long sbase) {
        const int e = (int)(t % RING);
        RingEntry<Real> &E = s_ring[e];
        uint32_t o4[4];
        if (!a.stream) philox4x32_10(a.seed, (uint64_t)t, gchain, 2u * attempt, o4);
#pragma unroll
        for (int k = 0; k < MCSAS_MAX_ACTIVE; ++k) {
            if (k < P) {
                double u = a.stream ? a.stream[sbase + t * P + k]
                                    : generator_output(o4[k], a.mp.gen_kind[k]);
                E.prop[k] = u * (a.mp.hi[k] - a.mp.lo[k]) + a.mp.lo[k];  // mcsas.py:358
            }
        }
        PVec pn(a.mp, E.prop);
        O::template prep<FAST>(pn, a.comp_exp, G, qt, E.cn, s_nodes + (size_t)(e * 2) * NSTR);
        PVec po(a.mp, s_rset + (int)(t % N) * P);
        O::template prep<FAST>(po, a.comp_exp, G, qt, E.co, s_nodes + (size_t)(e * 2 + 1) * NSTR);
    };

    const unsigned long long t_start = globaltimer_ns();
    long long step = 0, moves = 0, evals = 0, steps_all = 0;
    double conval = 0., conval0 = 0., sc_a = 1., sc_b = 0.;
    int attempt = 0, status = 1;
    // relaunch: status 0 = continue the saved attempt, 2 = start the next attempt afresh
    bool resume = a.resume != 0 && a.status[chain] == 0;
    if (a.resume) {
        evals = a.evals[chain]; steps_all = a.steps_all[chain];
        attempt = a.attempts[chain] - 1;
        if (resume) { step = a.num_iter[chain]; moves = a.num_moves[chain]; conval0 = a.conval0[chain]; }
    }
    const long long slice_end = a.slice_steps > 0 ? step + a.slice_steps : (1LL << 62);

    for (;; ++attempt) {
        // ---- contribution set: fresh draws (mcsas.py:310-317) or saved state ---------------
        for (int i = tid; i < N; i += blockDim.x) {
            uint32_t o4[4];
            if (!resume && !a.start_min && !a.stream)
                philox4x32_10(a.seed, (uint64_t)i, gchain, 2u * attempt + 1u, o4);
            for (int k = 0; k < P; ++k) {
                double lo = a.mp.lo[k], hi = a.mp.hi[k], val;
                if (resume) {
                    val = a.rset[((long long)chain * N + i) * P + k];
                } else if (a.start_min) {
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
        // ---- ft = sum of all contributions, in contribution order (mcsas.py:319) -----------
        for (int i0 = 0; i0 < N; i0 += RING) {
            int nb = min(RING, N - i0);
            for (int e = tid; e < nb; e += blockDim.x) {
                PVec p(a.mp, s_rset + (i0 + e) * P);
                O::template prep<FAST>(p, a.comp_exp, G, qt, s_ring[e].co, s_nodes + (size_t)(e * 2 + 1) * NSTR);
            }
            __syncthreads();
            if (!producer) {
                for (int e = 0; e < nb; ++e) {
                    const Contrib<Real> cc = s_ring[e].co;
                    const Real *nd = s_nodes + (size_t)(e * 2 + 1) * NSTR;
#pragma unroll
                    for (int j = 0; j < QPT; ++j) ft[j] += (double)eval_q<MODEL, Real>(cc, nd, qt, Q[j]);
                }
            }
            __syncthreads();
        }
        evals += N;
        lsq_all(sc_a, sc_b, conval);                             // mcsas.py:339-344
        if (!resume) { conval0 = conval; step = 0; moves = 0; }
        resume = false;
        const long long sbase = (long long)chain * a.stream_stride + (a.start_min ? 0 : (long long)N * P);
        // ---- first ring fill: steps [step, step + RING) -------------------------------------
        {
            long long target = min(step + (long long)RING, a.max_iter);
            for (long long t = step + tid; t < target; t += blockDim.x) fill_entry(t, attempt, sbase);
            if (tid == 0) {
                s_ctrl->step = step; s_ctrl->filled = target; s_ctrl->batch = 0;
                s_ctrl->done = !(N > 1 && conval > a.conv_crit && step < a.max_iter && step < slice_end);
            }
        }
        __syncthreads();
        int K = max(1, kmax / 2);
        int batch = 0;
        const int wd_batches = (int)min((long long)0x3fffffff, 2 * (min(a.max_iter, slice_end) - step) + 64);

        if (producer) {
            // ---- producer warp: runs ahead of the chain, decoupled from the eval warps.  It may
            // write entry t once t < (published batch start) + RING: that slot's previous content
            // (step t - RING) is already consumed, and the replaced contribution t % N cannot
            // change before step t is evaluated because RING <= N.
            long long filled = s_ctrl->filled;
            for (;;) {
                const long long pstep = ld_volatile(&s_ctrl->step);
                const int pdone = ld_volatile_i(&s_ctrl->done);
                __threadfence_block();                            // rset updates before `step`
                if (pdone) { step = ld_volatile(&s_ctrl->step); if (pdone == 2) status = -2; break; }
                const long long target = min(pstep + (long long)RING, a.max_iter);
                if (filled < target) {
                    for (long long t = filled + lane; t < target; t += 32) fill_entry(t, attempt, sbase);
                    __syncwarp();
                    __threadfence_block();
                    filled = target;
                    if (lane == 0) st_volatile(&s_ctrl->filled, filled);
                } else {
                    __nanosleep(64);
                }
            }
        } else {
        // ---- the MC loop (mcsas.py:354-404), K speculative steps per batch -------------------
        for (;;) {
            if (ld_volatile_i(&s_ctrl->done)) {
                if (s_ctrl->done == 2) status = -2;
                break;
            }
            const int Kb = (int)min((long long)min(K, kmax), a.max_iter - step);
            ++batch;
            bool wd = Kb < 1 || batch > wd_batches;               // watchdog: never spin forever
            if (!wd) {                                            // ring must hold [step, step+Kb)
                int spins = 0;
                while (ld_volatile(&s_ctrl->filled) < step + Kb) {
                    if (++spins > (1 << 24)) { wd = true; break; }
                    __nanosleep(32);
                }
                __threadfence_block();
            }
            wd = __any_sync(0xffffffffu, wd);
            // ---- evaluate Kb proposals, GRP at a time, lane-permuted order -------------------
            const int perm = (lane >> 2) & (GRP - 1);
            for (int k0 = 0; k0 < Kb && !wd; k0 += GRP) {
                double acc[GRP][3];
#pragma unroll
                for (int s = 0; s < GRP; ++s) {
                    acc[s][0] = 0.; acc[s][1] = 0.; acc[s][2] = 0.;
                    const int k = k0 + (s ^ perm);
                    if (k < Kb) {
                        const int e = (int)((step + k) % RING);
                        const Contrib<Real> cn = s_ring[e].cn, co = s_ring[e].co;
                        const Real *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
#pragma unroll
                        for (int j = 0; j < QPT; ++j) {
                            Real inew = eval_q<MODEL, Real>(cn, nn, qt, Q[j]);       // mcsas.py:360
                            Real iold = eval_q<MODEL, Real>(co, no, qt, Q[j]);       // mcsas.py:362
                            double M = model_sum<Real>(ft[j], inew, iold);   // :366
                            double t = w[j] * M;
                            acc[s][0] += t;
                            acc[s][1] = fma(t, M, acc[s][1]);
                            acc[s][2] = fma(wI[j], M, acc[s][2]);
                        }
                    }
                }
                // select-free reduce-scatter: slot s of lane l holds proposal s ^ perm(l)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) acc[s][i] += __shfl_xor_sync(0xffffffffu, acc[s + 4][i], 16);
#pragma unroll
                    for (int s = 0; s < 2; ++s) acc[s][i] += __shfl_xor_sync(0xffffffffu, acc[s + 2][i], 8);
                    acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[1][i], 4);
                    acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], 2);
                    acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], 1);
                }
                if ((lane & 3) == 0) {
                    double *r = s_wsum + ((size_t)warp * KCAP + k0 + perm) * 3;
                    r[0] = acc[0][0]; r[1] = acc[0][1]; r[2] = acc[0][2];
                }
            }
            if (wd && lane == 0) s_dec[1] = 1;                    // any warp may raise it
            named_bar_sync(1, T);                                // B1: eval warps only
            if (tid == 0) { __threadfence_block(); st_volatile(&s_ctrl->step, step); }
            wd = s_dec[1] != 0;                                   // uniform over the CTA ...
            // ---- exchange the partial sums over the cluster (st.async + mbarrier) -------------
            double *part = s_part + (size_t)xbuf * MAX_CLUSTER * KCAP * 3;
            if (C > 1) {
                const int nval = 3 * Kb + 1;                      // + watchdog flag of this CTA
                if (tid == 0) mbar_expect_tx(s_xbar + xbuf, (uint32_t)(C * nval * 8));
                for (int i = tid; i < nval; i += T) {
                    double x = 0.;
                    if (i < 3 * Kb) { for (int wv = 0; wv < W; ++wv) x += s_wsum[(size_t)wv * KCAP * 3 + i]; }
                    else x = wd ? 1. : 0.;
                    double *dst = (i < 3 * Kb) ? part + rank * KCAP * 3 + i : s_xflag + xbuf * MAX_CLUSTER + rank;
                    for (int r = 0; r < C; ++r) dsmem_store_async(dst, s_xbar + xbuf, (uint32_t)r, x);
                }
                if (warp == 0) {
                    mbar_wait(s_xbar + xbuf, (xphase >> xbuf) & 1);
                    bool anywd = false;                           // ... and over the cluster
                    for (int r = 0; r < C; ++r) anywd |= s_xflag[xbuf * MAX_CLUSTER + r] != 0.;
                    wd = anywd;
                }
                xphase ^= 1u << xbuf;
            }
            if (warp == 0) {
                double chi = 0., sa = 0., sb = 0.;
                bool better = false;
                if (lane < Kb && !wd) {
                    double sm = 0., smm = 0., smi = 0.;
                    if (C > 1) {
                        for (int r = 0; r < C; ++r) {
                            const double *x = part + ((size_t)r * KCAP + lane) * 3;
                            sm += x[0]; smm += x[1]; smi += x[2];
                        }
                    } else {
                        for (int wv = 0; wv < W; ++wv) {
                            const double *x = s_wsum + ((size_t)wv * KCAP + lane) * 3;
                            sm += x[0]; smm += x[1]; smi += x[2];
                        }
                    }
                    lsq_solve(L, sm, smm, smi, sa, sb, chi);     // mcsas.py:376
                    better = chi < conval;                       // :379 strict
                    s_chi[lane * 3] = chi; s_chi[lane * 3 + 1] = sa; s_chi[lane * 3 + 2] = sb;
                }
                unsigned mask = __ballot_sync(0xffffffffu, better);
                if (lane == 0) { s_dec[0] = mask ? (__ffs(mask) - 1) : -1; s_dec[1] = wd ? 2 : 0; }
            }
            xbuf ^= 1;
            named_bar_sync(1, T);                                // B2
            if (s_dec[1] == 2) {                                  // watchdog: identical in all CTAs
                if (tid == 0) { st_volatile(&s_ctrl->step, step); __threadfence_block(); st_volatile_i(&s_ctrl->done, 2); }
                status = -2;
                break;
            }
            if (TWO_PASS) {
                // second pass: chi-squared from the residuals with each proposal's own (a, b),
                // as the reference computes it -- immune to the cancellation of the expanded form
                for (int k0 = 0; k0 < Kb; k0 += GRP) {
                    double acc[GRP];
#pragma unroll
                    for (int s = 0; s < GRP; ++s) {
                        acc[s] = 0.;
                        const int k = k0 + (s ^ perm);
                        if (k < Kb) {
                            const int e = (int)((step + k) % RING);
                            const Contrib<Real> cn = s_ring[e].cn, co = s_ring[e].co;
                            const Real *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
                            const double ak = s_chi[k * 3 + 1], bk = s_chi[k * 3 + 2];
#pragma unroll
                            for (int j = 0; j < QPT; ++j) {
                                Real inew = eval_q<MODEL, Real>(cn, nn, qt, Q[j]);
                                Real iold = eval_q<MODEL, Real>(co, no, qt, Q[j]);
                                double M = model_sum<Real>(ft[j], inew, iold);
                                double r = Iq[j] - (M * ak + bk);
                                acc[s] = fma(w[j] * r, r, acc[s]);
                            }
                        }
                    }
#pragma unroll
                    for (int s = 0; s < 4; ++s) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s + 4], 16);
#pragma unroll
                    for (int s = 0; s < 2; ++s) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s + 2], 8);
                    acc[0] += __shfl_xor_sync(0xffffffffu, acc[1], 4);
                    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 2);
                    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
                    if ((lane & 3) == 0) s_wsum[((size_t)warp * KCAP + k0 + perm) * 3] = acc[0];
                }
                named_bar_sync(1, T);
                double *part2 = s_part + (size_t)xbuf * MAX_CLUSTER * KCAP * 3;
                if (C > 1) {
                    if (tid == 0) mbar_expect_tx(s_xbar + xbuf, (uint32_t)(C * Kb * 8));
                    for (int i = tid; i < Kb; i += T) {
                        double x = 0.;
                        for (int wv = 0; wv < W; ++wv) x += s_wsum[((size_t)wv * KCAP + i) * 3];
                        for (int r = 0; r < C; ++r)
                            dsmem_store_async(part2 + ((size_t)rank * KCAP + i) * 3, s_xbar + xbuf, (uint32_t)r, x);
                    }
                    if (warp == 0) mbar_wait(s_xbar + xbuf, (xphase >> xbuf) & 1);
                    xphase ^= 1u << xbuf;
                }
                if (warp == 0) {
                    bool better = false;
                    if (lane < Kb) {
                        double x = 0.;
                        if (C > 1) { for (int r = 0; r < C; ++r) x += part2[((size_t)r * KCAP + lane) * 3]; }
                        else { for (int wv = 0; wv < W; ++wv) x += s_wsum[((size_t)wv * KCAP + lane) * 3]; }
                        const double chi = x * L.inv_nq;
                        better = chi < conval;
                        s_chi[lane * 3] = chi;
                    }
                    unsigned mask = __ballot_sync(0xffffffffu, better);
                    if (lane == 0) s_dec[0] = mask ? (__ffs(mask) - 1) : -1;
                }
                xbuf ^= 1;
                named_bar_sync(1, T);
            }
            const int kstar = s_dec[0];
            const int used = kstar >= 0 ? kstar + 1 : Kb;
            if (a.trace_accept && rank == 0) {
                for (int k = tid; k < used; k += T) {
                    a.trace_accept[(long long)chain * a.max_iter + step + k] = (k == kstar);
                    a.trace_chi[(long long)chain * a.max_iter + step + k] = s_chi[k * 3];
                }
            }
            if (kstar >= 0) {                                    // mcsas.py:381-384
                const int e = (int)((step + kstar) % RING);
                const Contrib<Real> cn = s_ring[e].cn, co = s_ring[e].co;
                const Real *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                    Real inew = eval_q<MODEL, Real>(cn, nn, qt, Q[j]);
                    Real iold = eval_q<MODEL, Real>(co, no, qt, Q[j]);
                    ft[j] = model_sum<Real>(ft[j], inew, iold);
                }
                conval = s_chi[kstar * 3]; sc_a = s_chi[kstar * 3 + 1]; sc_b = s_chi[kstar * 3 + 2];
                const int rix = (int)((step + kstar) % N);
                if (tid < P) s_rset[rix * P + tid] = s_ring[e].prop[tid];
                ++moves;
                evals += 2;
                K = max(1, min(kmax, 1 << (31 - __clz(kstar + 1))));
            } else {
                K = min(kmax, K * 2);
            }
            evals += 2 * Kb;
            step += used;                                        // :403-404 (ri = step % N)
            steps_all += used;
            if (!(conval > a.conv_crit && step < a.max_iter && step < slice_end)) {
                if (tid == 0) {
                    st_volatile(&s_ctrl->step, step);
                    __threadfence_block();
                    st_volatile_i(&s_ctrl->done, 1);
                }
                break;
            }
            // (the new batch start is published to the producer after the next B1, when every
            //  eval thread has finished reading the ring entry and rset of this batch)
        }
        }
        __syncthreads();                                         // eval warps and producer rejoin
        // the producer learns the chain's progress from the control block
        step = s_ctrl->step;
        if (producer) { conval = 0.; }
        // ---- final fit of the accepted state (mcsas.py:424-425) --------------------------------
        lsq_all(sc_a, sc_b, conval);
        if (status == -2) break;                                 // watchdog fired (eval threads)
        const bool running = (N > 1 && conval > a.conv_crit && step < a.max_iter);
        if (running) { status = 0; break; }                      // slice ended: resume later
        status = 1;
        if (conval <= a.conv_crit || a.stream || attempt + 1 >= a.max_attempts || N <= 1)
            break;
        if (a.slice_steps > 0) {                                 // retry (mcsas.py:220-246) in a
            status = 2; ++attempt; step = 0; moves = 0;          // later slice, freshly re-packed
            break;
        }
        __syncthreads();
    }

    // ---- results / saved state ---------------------------------------------------------------
    if (rank == 0)
        for (int i = tid; i < N * P; i += blockDim.x) a.rset[(long long)chain * N * P + i] = s_rset[i];
    if (!producer) {
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            int i = (j * C + rank) * T + tid;
            if (i < nq)
                a.fit_meas[(long long)chain * a.nq_stride + i] = ft[j] * sc_a + sc_b;   // mcsas.py:429
        }
    }
    if (tid == 0 && rank == 0) {
        // internal amplitudes are w*G: scaling in the reference's (SI) units is a*G
        a.scaling[chain] = sc_a * G;
        a.background[chain] = sc_b;
        a.conval[chain] = conval;
        a.conval0[chain] = conval0;
        a.num_iter[chain] = step;
        a.num_moves[chain] = moves;
        a.evals[chain] = evals;
        a.steps_all[chain] = steps_all;
        a.attempts[chain] = attempt + 1;
        a.elapsed_s[chain] = (a.resume ? a.elapsed_s[chain] : 0.) + 1e-9 * (double)(globaltimer_ns() - t_start);
        a.status[chain] = status;
    }
    if (C > 1) cluster_sync_all();      // no CTA may exit while peers could still address its smem
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
    const QuadTab<Real> qt = { D, a.mp.quad_rule, s_qx, s_qw, &a.smear, a.smear.n };
    if (tid == 0) {
        PVec p(a.mp, a.pset + (long long)c * a.mp.n_active);
        double G = 1.;
        if (sizeof(Real) == 4) {
            Contrib<Real> tmp;
            VWS o0 = O::template prep<false>(p, a.comp_exp, 1., qt, tmp, s_nodes);
            G = 1. / o0.amp_raw;                 // this contribution's own amplitude -> O(1)
            if (!(G > 0.) || isinf(G)) G = 1.;
        }
        VWS o = O::template prep<false>(p, a.comp_exp, G, qt, s_cc, s_nodes);
        s_G = G;
        a.v[c] = o.v; a.w[c] = o.w; a.s[c] = o.s;
    }
    __syncthreads();
    const double *q = a.data + ds.offset;
    for (int i = tid; i < ds.nq; i += blockDim.x) {
        QRep<Real> Q;
        Q.set(q[i]);
        a.intensity[(long long)c * ds.nq + i] = (double)eval_q<MODEL, Real>(s_cc, s_nodes, qt, Q) / s_G;
    }
}
