// chain_fast.cuh -- the PRODUCTION (fp32) MC-chain kernel.
//
// Same contract as chain_kernel (chain_kernel.cuh): one greedy chain of McSAS.mcFit
// (src/mcsas/mcsas/mcsas.py:287-439) per thread-block cluster, speculative batches, a ring
// of prepared proposals, time slices.  What differs:
//
//  * RESIDUAL-DELTA least squares.  The reference re-solves scale/background for every
//    proposal (backgroundscalingfit.py:112-139).  chi^2(a,b) is an exact quadratic, so for
//    a proposal that changes the model by d = I_new - I_old (mcsas.py:366) its minimum
//    follows from FOUR sums over q of the small quantity u = d/sigma,
//        s1 = sum rho u,  s2 = sum u^2,  s3 = sum mu u,  s4 = sum omega u
//    (rho = current residual/sigma, mu = current model/sigma, omega = 1/sigma), expanded
//    around the current fit (a0,b0).  None of them cancels against a large base term, so
//    fp32 accumulation resolves the accept test `chi2_new < chi2_old` (mcsas.py:379) to
//    ~1e-6 of the CHANGE of chi2 -- no fp64 and no conversion in the inner loop.  The base
//    sums of the accepted state are re-accumulated exactly in fp64 inside the next batch.
//  * GRAM batches while moves are frequent: 8 steps per batch, all consumed; the Gram matrix
//    of the proposals' changes lets warp 0 re-derive the sums of later proposals after each
//    accepted one, so several moves are committed per batch (exact sequential semantics).
//  * CONTRIBUTION-ROW CACHE: the scaled intensity rows of the current contributions live in
//    a per-SM region of global memory (L2/HBM); the replaced contribution of a step is one
//    coalesced 16-byte load per thread instead of a second evaluation (mcsas.py:362).
//  * the 1/(qr)^3 of the sphere amplitude is split into a per-contribution factor (folded
//    into the amplitude when a ring entry is prepared) and a per-q factor (folded into
//    1/sigma), so an evaluation costs MUFU.SIN + MUFU.COS and no MUFU.RCP; the 32
//    evaluations of a chunk of 8 proposals form one basic block (one warp vote selects the
//    small-x series variant), which ptxas interleaves.
//  * no dedicated producer warp: all 8 warps evaluate; while warp 0 solves the batch, the
//    last warp prepares the ring entries of the next one.  When the q slice of a CTA needs
//    fewer warps, the spare warps form PROPOSAL GROUPS that evaluate other chunks.
//  * SCHEDULE-INDEPENDENT RESULTS: every warp's partial sums reach every CTA of the chain's
//    cluster directly (st.async + mbarrier transaction count, no CTA barrier in between) and
//    are added in one fixed order, and a launch that ends mid-chain saves everything the loop
//    carries (fp64 master intensity, centre of the expansion, running chi-squared, batch
//    size).  A chain's result is therefore a function of (seed, chain id) alone -- the same
//    bits for every cluster size, time slicing and GPU count.
#pragma once
#include "chain_kernel.cuh"

// ---- packed fp32 (f32x2: FFMA2 / FMUL2 / FADD2 of sm_100a) -------------------------------------------
// The evaluation loop is bound by issue slots, not by arithmetic: two q points of a thread share one
// instruction.  Every packed operation rounds each half exactly as the scalar one does.
__device__ __forceinline__ float2 pk2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 pk1(float a) { return make_float2(a, a); }       // (ptxas: broadcast operand)
__device__ __forceinline__ float2 pkneg(float2 a) { return make_float2(-a.x, -a.y); } // (operand modifier)
// a - b, each half rounded once: opaque to the compiler, so it is never contracted with a producing
// multiplication (the packed intrinsics are)
__device__ __forceinline__ float2 sub2_exact(float2 a, float2 b)
{
    float2 c;
    asm("{ .reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "sub.rn.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc; }"
        : "=f"(c.x), "=f"(c.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return c;
}

// The four delta sums of one proposal over the thread's q points (file header): u = cq * dd,
// s1 += rho u, s2 += u u, s3 += mu u, s4 += om u.  ONE definition for every batch kind, so a proposal's
// sums never depend on the path that evaluated it.  Even QPT: the even and the odd q points of the
// thread accumulate in the two halves of packed registers and are added at the end.
template <int QPT>
__device__ __forceinline__ void fast_sums(const float (&dd)[QPT], const float (&cq)[QPT], const float (&rho)[QPT],
                                          const float (&mu)[QPT], const float (&om)[QPT], float (&acc)[4],
                                          float (&u)[QPT])
{
    if constexpr (QPT % 2 == 0) {
        float2 A0, A1, A2, A3;
#pragma unroll
        for (int p = 0; p < QPT / 2; ++p) {
            const float2 u2 = __fmul2_rn(pk2(cq[2 * p], cq[2 * p + 1]), pk2(dd[2 * p], dd[2 * p + 1]));
            u[2 * p] = u2.x; u[2 * p + 1] = u2.y;
            const float2 r2 = pk2(rho[2 * p], rho[2 * p + 1]), m2 = pk2(mu[2 * p], mu[2 * p + 1]);
            const float2 o2 = pk2(om[2 * p], om[2 * p + 1]);
            if (p == 0) {
                A0 = __fmul2_rn(r2, u2); A1 = __fmul2_rn(u2, u2); A2 = __fmul2_rn(m2, u2); A3 = __fmul2_rn(o2, u2);
            } else {
                A0 = __ffma2_rn(r2, u2, A0); A1 = __ffma2_rn(u2, u2, A1);
                A2 = __ffma2_rn(m2, u2, A2); A3 = __ffma2_rn(o2, u2, A3);
            }
        }
        acc[0] = __fadd_rn(A0.x, A0.y); acc[1] = __fadd_rn(A1.x, A1.y);
        acc[2] = __fadd_rn(A2.x, A2.y); acc[3] = __fadd_rn(A3.x, A3.y);
    } else {
        acc[0] = 0.f; acc[1] = 0.f; acc[2] = 0.f; acc[3] = 0.f;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            u[j] = cq[j] * dd[j];
            acc[0] = fmaf(rho[j], u[j], acc[0]);
            acc[1] = fmaf(u[j], u[j], acc[1]);
            acc[2] = fmaf(mu[j], u[j], acc[2]);
            acc[3] = fmaf(om[j], u[j], acc[3]);
        }
    }
}

// ---- per-model fast evaluation: scaled intensity D with  I(q) = D * (q_ref/q)^QS_POW ------
template <int MODEL> struct FastOps {
    using O = Ops<MODEL, float>;
    static constexpr int QS_POW = 0;
    template <class PV>
    __device__ static void prep(const PV &p, double ce, double G, float, const QuadTab<float> &qt,
                                Contrib<float> &cc, float *nodes)
    {
        O::template prep<true>(p, ce, G, qt, cc, nodes);
    }
    __device__ __forceinline__ static float evalS(const Contrib<float> &cc, const float *nodes,
                                                  const QuadTab<float> &qt, const QRep<float> &Q)
    {
        return eval_q<MODEL, float>(cc, nodes, qt, Q);
    }
    // scaled intensity of one contribution for the thread's q points
    template <int QPT>
    __device__ __forceinline__ static void evalN(const Contrib<float> &cc, const float *nodes,
                                                 const QuadTab<float> &qt, const QRep<float> (&Q)[QPT],
                                                 float (&d)[QPT])
    {
        if constexpr ((MODEL == MCSAS_CYLINDERS_ISOTROPIC || MODEL == MCSAS_ELLIPSOIDS_ISOTROPIC
                       || MODEL == MCSAS_ELLIPSOIDAL_CORE_SHELL) && QPT > 1) {
            O::template evalV<QPT>(cc, nodes, qt, Q, d);       // q points interleaved in the node loop
            return;
        }
#pragma unroll
        for (int j = 0; j < QPT; ++j) d[j] = evalS(cc, nodes, qt, Q[j]);
    }
    // scaled intensity change of a proposal: D(new) - D(old) for the thread's q points
    template <int QPT>
    __device__ __forceinline__ static void diffS(const Contrib<float> &cn, const Contrib<float> &co,
                                                 const float *nn, const float *no, const QuadTab<float> &qt,
                                                 const QRep<float> (&Q)[QPT], float (&d)[QPT])
    {
#pragma unroll
        for (int j = 0; j < QPT; ++j) d[j] = evalS(cn, nn, qt, Q[j]) - evalS(co, no, qt, Q[j]);
    }
    // NS contributions (ring entries ent[s]) at once
    template <int QPT, int NS>
    __device__ __forceinline__ static void evalChunk(const RingEntry<float> *ring, const int (&ent)[NS],
                                                     const float *nodes, size_t nstr, const QuadTab<float> &qt,
                                                     const QRep<float> (&Q)[QPT], float (&d)[NS][QPT])
    {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const Contrib<float> cn = ring[ent[s]].cn;
            evalN<QPT>(cn, nodes + (size_t)(ent[s] * 2) * nstr, qt, Q, d[s]);
        }
    }
    // NS proposals against the cached rows `od` of the contributions they replace:
    // dn = D(new) (the new row), dd = dn - od
    template <int QPT, int NS>
    __device__ __forceinline__ static void diffChunk(const RingEntry<float> *ring, const int (&ent)[NS],
                                                     const float *nodes, size_t nstr, const QuadTab<float> &qt,
                                                     const QRep<float> (&Q)[QPT], const float (&od)[NS][QPT],
                                                     float (&dn)[NS][QPT], float (&dd)[NS][QPT])
    {
        evalChunk<QPT, NS>(ring, ent, nodes, nstr, qt, Q, dn);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int j = 0; j < QPT; ++j) dd[s][j] = __fsub_rn(dn[s][j], od[s][j]);
    }
};

// g(x) = sin x - x cos x = x^3 F(x)/3 (sphere.py:62) for the M (q, radius) pairs of a
// thread at once; SER: series below x = 1, where s - x c cancels.  Written stage by stage
// over the M independent chains: ptxas keeps that order, which gives M-way instruction
// level parallelism (an eval-by-eval formulation is scheduled serially, one dependent
// instruction every ~5 cycles).
template <int M, bool SER>
__device__ __forceinline__ void sph_g_vec(const float (&q)[M], const float (&th)[M], const float (&tl)[M],
                                          const float (&r)[M], float (&g)[M])
{
    float x[M], t[M], fr[M], sn[M], cs[M];
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = fmaf(th[i], r[i], 12582912.f);     // rint(th r) for |t| < 2^22
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = q[i] * r[i];
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = t[i] - 12582912.f;
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(th[i], r[i], -t[i]);         // exact up to one rounding
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(tl[i], r[i], fr[i]);         // low part of q / (2 pi)
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fr[i] * 6.2831853071795865f;
#pragma unroll
    for (int i = 0; i < M; ++i) sn[i] = __sinf(fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) cs[i] = __cosf(fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) g[i] = fmaf(-x[i], cs[i], sn[i]);
    if (SER) {
        float x2[M], p[M];
#pragma unroll
        for (int i = 0; i < M; ++i) x2[i] = x[i] * x[i];
#pragma unroll
        for (int i = 0; i < M; ++i) p[i] = fmaf(x2[i], 2.5052108e-7f, -2.2045855e-5f);
#pragma unroll
        for (int i = 0; i < M; ++i) p[i] = fmaf(x2[i], p[i], 1.1904762e-3f);
#pragma unroll
        for (int i = 0; i < M; ++i) p[i] = fmaf(x2[i], p[i], -0.033333333f);
#pragma unroll
        for (int i = 0; i < M; ++i) p[i] = fmaf(x2[i], p[i], 0.33333333f);
#pragma unroll
        for (int i = 0; i < M; ++i) g[i] = x[i] < 1.f ? (x[i] * x2[i]) * p[i] : g[i];
    }
}

// The same g(x), same operations and roundings, for MP PAIRS of q points (packed halves) and NS radii:
// stage by stage over all NS x MP packed chains.
template <int NS, int MP, bool SER>
__device__ __forceinline__ void sph_g_pack(const float2 (&q)[MP], const float2 (&th)[MP], const float2 (&tl)[MP],
                                           const float (&r)[NS], float2 (&g)[NS][MP])
{
    float2 x[NS][MP], t[NS][MP], fr[NS][MP], sn[NS][MP], cs[NS][MP];
    const float2 magic = pk1(12582912.f), nmagic = pk1(-12582912.f), twopi = pk1(6.2831853071795865f);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) t[s][i] = __ffma2_rn(th[i], pk1(r[s]), magic);   // rint(th r) for |t| < 2^22
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) x[s][i] = __fmul2_rn(q[i], pk1(r[s]));
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) t[s][i] = __fadd2_rn(t[s][i], nmagic);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) fr[s][i] = __ffma2_rn(th[i], pk1(r[s]), pkneg(t[s][i]));
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) fr[s][i] = __ffma2_rn(tl[i], pk1(r[s]), fr[s][i]);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) fr[s][i] = __fmul2_rn(fr[s][i], twopi);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) {
            sn[s][i] = pk2(__sinf(fr[s][i].x), __sinf(fr[s][i].y));
            cs[s][i] = pk2(__cosf(fr[s][i].x), __cosf(fr[s][i].y));
        }
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < MP; ++i) g[s][i] = __ffma2_rn(pkneg(x[s][i]), cs[s][i], sn[s][i]);
    if (SER) {
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                const float2 x2 = __fmul2_rn(x[s][i], x[s][i]);
                float2 p = __ffma2_rn(x2, pk1(2.5052108e-7f), pk1(-2.2045855e-5f));
                p = __ffma2_rn(x2, p, pk1(1.1904762e-3f));
                p = __ffma2_rn(x2, p, pk1(-0.033333333f));
                p = __ffma2_rn(x2, p, pk1(0.33333333f));
                const float2 sg = __fmul2_rn(__fmul2_rn(x[s][i], x2), p);
                g[s][i].x = x[s][i].x < 1.f ? sg.x : g[s][i].x;
                g[s][i].y = x[s][i].y < 1.f ? sg.y : g[s][i].y;
            }
    }
}

template <> struct FastOps<MCSAS_SPHERE> {
    using O = Ops<MCSAS_SPHERE, float>;
    static constexpr int QS_POW = 6;
    template <class PV>
    __device__ static void prep(const PV &p, double ce, double G, float qref, const QuadTab<float> &qt,
                                Contrib<float> &cc, float *nodes)
    {
        O::template prep<true>(p, ce, G, qt, cc, nodes);
        double xr = (double)cc.c[0] * (double)qref;          // r * q_ref
        double x3 = xr * xr * xr;
        cc.c[2] = (float)((double)cc.amp / (x3 * x3));       // amp / (r q_ref)^6
        cc.c[1] = cc.c[0];                                   // (r, A) = one aligned 8-byte load
    }
    __device__ __forceinline__ static float evalS(const Contrib<float> &cc, const float *,
                                                  const QuadTab<float> &qt, const QRep<float> &Q)
    {
        if (qt.nsm) return eval_q<MCSAS_SPHERE, float>(cc, nullptr, qt, Q);     // smeared: no q fold
        const float q[1] = {Q.q}, th[1] = {Q.th}, tl[1] = {Q.tl}, r[1] = {cc.c[0]};
        float g[1];
        sph_g_vec<1, true>(q, th, tl, r, g);
        return (g[0] * g[0]) * cc.c[2];
    }
    // smeared, tabulated locations, ONE radius (row build, the update after a move, batches shorter than a
    // chunk): four profile points per pass, so that 4 x MP independent packed evaluations are in flight where a
    // point-by-point loop has MP; the sum over the profile points is accumulated in their order, with the
    // operations of the chunk evaluator (smear_pack) -- the same bits
    template <int MP, bool SER>
    __device__ __forceinline__ static void smear_one(const QuadTab<float> &qt, float r, float2 (&acc)[MP])
    {
        constexpr int KU = 4;
        const float rr[1] = {r};
#pragma unroll
        for (int i = 0; i < MP; ++i) acc[i] = pk1(0.f);
        int k = 0;
#pragma unroll 1
        for (; k + KU <= qt.nsm; k += KU) {
            float2 q[KU * MP], th[KU * MP], tl[KU * MP], w[KU * MP], g[1][KU * MP];
#pragma unroll
            for (int u = 0; u < KU; ++u)
#pragma unroll
                for (int i = 0; i < MP; ++i) {
                    const float4 e0 = __ldg(qt.smloc + (size_t)(k + u) * qt.smstride + 2 * i);
                    const float4 e1 = __ldg(qt.smloc + (size_t)(k + u) * qt.smstride + 2 * i + 1);
                    q[u * MP + i] = pk2(e0.x, e1.x); th[u * MP + i] = pk2(e0.y, e1.y);
                    tl[u * MP + i] = pk2(e0.z, e1.z); w[u * MP + i] = pk2(e0.w, e1.w);
                }
            sph_g_pack<1, KU * MP, SER>(q, th, tl, rr, g);
#pragma unroll
            for (int u = 0; u < KU; ++u)
#pragma unroll
                for (int i = 0; i < MP; ++i)
                    acc[i] = __ffma2_rn(w[u * MP + i], __fmul2_rn(g[0][u * MP + i], g[0][u * MP + i]), acc[i]);
        }
#pragma unroll 1
        for (; k < qt.nsm; ++k) {
            float2 q[MP], th[MP], tl[MP], w[MP], g[1][MP];
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                const float4 e0 = __ldg(qt.smloc + (size_t)k * qt.smstride + 2 * i);
                const float4 e1 = __ldg(qt.smloc + (size_t)k * qt.smstride + 2 * i + 1);
                q[i] = pk2(e0.x, e1.x); th[i] = pk2(e0.y, e1.y); tl[i] = pk2(e0.z, e1.z); w[i] = pk2(e0.w, e1.w);
            }
            sph_g_pack<1, MP, SER>(q, th, tl, rr, g);
#pragma unroll
            for (int i = 0; i < MP; ++i) acc[i] = __ffma2_rn(w[i], __fmul2_rn(g[0][i], g[0][i]), acc[i]);
        }
    }
    // q ascends along the lanes, so "some x < 1" is warp-uniform for all but one warp per
    // radius: one vote per proposal selects the variant without the series.  Callers are
    // warp-convergent.
    template <int QPT>
    __device__ __forceinline__ static void evalN(const Contrib<float> &cc, const float *, const QuadTab<float> &qt,
                                                 const QRep<float> (&Q)[QPT], float (&d)[QPT])
    {
        if (qt.nsm) {
            if (qt.smloc) {
                // smeared (sasmodel.py:56-73): sum over the profile points of w g(loc r)^2 at the tabulated
                // locations (location, its split phase factor and weight * (q_ref/loc)^6 come from ONE
                // 16-byte load: no fp64 square root, no phase split in the loop)
                if constexpr (QPT % 2 == 0) {
                    float2 a2[QPT / 2];
                    smear_one<QPT / 2, true>(qt, cc.c[0], a2);
#pragma unroll
                    for (int i = 0; i < QPT / 2; ++i) {
                        const float2 v = __fmul2_rn(a2[i], pk1(cc.c[2]));
                        d[2 * i] = v.x; d[(2 * i + 1) % QPT] = v.y;
                    }
                    return;
                }
                // (the table of a 1000-point dataset is 26 x 16 KB: its entries come from the L2, so the next
                // profile point's entries are requested before this one's are evaluated)
                float acc[QPT];
                float4 nx[QPT];
#pragma unroll
                for (int j = 0; j < QPT; ++j) { acc[j] = 0.f; nx[j] = __ldg(qt.smloc + j); }
#pragma unroll 1
                for (int k = 0; k < qt.nsm; ++k) {
                    float q[QPT], th[QPT], tl[QPT], r[QPT], w[QPT], g[QPT];
#pragma unroll
                    for (int j = 0; j < QPT; ++j) {
                        const float4 e = nx[j];
                        q[j] = e.x; th[j] = e.y; tl[j] = e.z; w[j] = e.w; r[j] = cc.c[0];
                    }
                    if (k + 1 < qt.nsm) {
#pragma unroll
                        for (int j = 0; j < QPT; ++j) nx[j] = __ldg(qt.smloc + (size_t)(k + 1) * qt.smstride + j);
                    }
                    // (the series variant for every point: a vote per profile point costs more than it saves)
                    sph_g_vec<QPT, true>(q, th, tl, r, g);
#pragma unroll
                    for (int j = 0; j < QPT; ++j) acc[j] = fmaf(w[j], g[j] * g[j], acc[j]);
                }
#pragma unroll
                for (int j = 0; j < QPT; ++j) d[j] = acc[j] * cc.c[2];
                return;
            }
#pragma unroll
            for (int j = 0; j < QPT; ++j) d[j] = eval_q<MCSAS_SPHERE, float>(cc, nullptr, qt, Q[j]);
            return;
        }
        float q[QPT], th[QPT], tl[QPT], r[QPT], g[QPT];
        bool small = false;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            q[j] = Q[j].q; th[j] = Q[j].th; tl[j] = Q[j].tl; r[j] = cc.c[0];
            small |= Q[j].q * cc.c[0] < 1.f;
        }
        if (__any_sync(0xffffffffu, small)) sph_g_vec<QPT, true>(q, th, tl, r, g);
        else sph_g_vec<QPT, false>(q, th, tl, r, g);
#pragma unroll
        for (int j = 0; j < QPT; ++j) g[j] = g[j] * g[j];
#pragma unroll
        for (int j = 0; j < QPT; ++j) d[j] = g[j] * cc.c[2];
    }
    // smeared, tabulated locations: sum over the profile points of w g(loc r)^2 for NS radii and MP pairs of
    // q points; every table entry is loaded once for the NS proposals
    template <int NS, int MP, bool SER>
    __device__ __forceinline__ static void smear_pack(const QuadTab<float> &qt, const float (&rr)[NS], float2 (&acc)[NS][MP])
    {
        if constexpr (NS == 1) {
            smear_one<MP, SER>(qt, rr[0], acc[0]);
            return;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int i = 0; i < MP; ++i) acc[s][i] = pk1(0.f);
        // (table entries come from the L2 -- 26 x 16 KB per 1000-point dataset: the next profile point's entries are
        // requested before this one's are evaluated)
        float4 nx[2 * MP];
#pragma unroll
        for (int i = 0; i < 2 * MP; ++i) nx[i] = __ldg(qt.smloc + i);
#pragma unroll 1
        for (int k = 0; k < qt.nsm; ++k) {
            float2 q[MP], th[MP], tl[MP], w[MP], g[NS][MP];
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                const float4 e0 = nx[2 * i], e1 = nx[2 * i + 1];
                q[i] = pk2(e0.x, e1.x); th[i] = pk2(e0.y, e1.y); tl[i] = pk2(e0.z, e1.z); w[i] = pk2(e0.w, e1.w);
            }
            if (k + 1 < qt.nsm) {
#pragma unroll
                for (int i = 0; i < 2 * MP; ++i) nx[i] = __ldg(qt.smloc + (size_t)(k + 1) * qt.smstride + i);
            }
            sph_g_pack<NS, MP, SER>(q, th, tl, rr, g);
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int i = 0; i < MP; ++i)
                    acc[s][i] = __ffma2_rn(w[i], __fmul2_rn(g[s][i], g[s][i]), acc[s][i]);
        }
    }
    // NS contributions at once: ONE vote selects the series variant for the whole chunk, so
    // the NS x QPT evaluations form a single basic block that ptxas can interleave freely
    template <int QPT, int NS>
    __device__ __forceinline__ static void evalChunk(const RingEntry<float> *ring, const int (&ent)[NS],
                                                     const float *, size_t, const QuadTab<float> &qt,
                                                     const QRep<float> (&Q)[QPT], float (&d)[NS][QPT])
    {
        if (qt.nsm) {
            if (qt.smloc) {
                // smeared, tabulated locations: the NS proposals share every table entry (ONE 16-byte load per
                // (profile point, q point) and chunk instead of one per proposal -- the loads held the L1 for
                // as long as the evaluations hold the XU pipe) and run through the packed evaluation; per
                // (proposal, q point) the same operations in the same order as evalN above
                float rr[NS], aa[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const float2 ra = *(const float2 *)&ring[ent[s]].cn.c[1];
                    rr[s] = ra.x; aa[s] = ra.y;
                }
                if constexpr (QPT % 2 == 0) {
                    constexpr int MP = QPT / 2;
                    float2 acc[NS][MP];
                    // The series branch (x < 1) is needed by no element of the chunk when the smallest
                    // unsmeared q times the smallest radius is >= 1 and no location lies below its q point
                    // (qt.smmono: slit profiles, 2-D profiles without negative offsets): ONE vote per chunk,
                    // as in the unsmeared evaluator; the variants agree bit for bit where x >= 1.
                    bool ser = true;
                    if (qt.smmono) {
                        float rmin = rr[0], qmin = Q[0].q;
#pragma unroll
                        for (int s = 1; s < NS; ++s) rmin = fminf(rmin, rr[s]);
#pragma unroll
                        for (int j = 1; j < QPT; ++j) qmin = fminf(qmin, Q[j].q);
                        ser = __any_sync(0xffffffffu, qmin * rmin < 1.f);
                    }
                    if (ser) smear_pack<NS, MP, true>(qt, rr, acc);
                    else smear_pack<NS, MP, false>(qt, rr, acc);
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int i = 0; i < MP; ++i) {
                            const float2 v = __fmul2_rn(acc[s][i], pk1(aa[s]));
                            d[s][2 * i] = v.x; d[s][(2 * i + 1) % QPT] = v.y;
                        }
                } else {
                    float acc[NS][QPT];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int j = 0; j < QPT; ++j) acc[s][j] = 0.f;
                    float4 nx[QPT];
#pragma unroll
                    for (int j = 0; j < QPT; ++j) nx[j] = __ldg(qt.smloc + j);
#pragma unroll 1
                    for (int k = 0; k < qt.nsm; ++k) {
                        float q[QPT], th[QPT], tl[QPT], w[QPT];
#pragma unroll
                        for (int j = 0; j < QPT; ++j) {
                            const float4 e = nx[j];
                            q[j] = e.x; th[j] = e.y; tl[j] = e.z; w[j] = e.w;
                        }
                        if (k + 1 < qt.nsm) {
#pragma unroll
                            for (int j = 0; j < QPT; ++j) nx[j] = __ldg(qt.smloc + (size_t)(k + 1) * qt.smstride + j);
                        }
#pragma unroll
                        for (int s = 0; s < NS; ++s) {
                            float r[QPT], g[QPT];
#pragma unroll
                            for (int j = 0; j < QPT; ++j) r[j] = rr[s];
                            sph_g_vec<QPT, true>(q, th, tl, r, g);
#pragma unroll
                            for (int j = 0; j < QPT; ++j) acc[s][j] = fmaf(w[j], g[j] * g[j], acc[s][j]);
                        }
                    }
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int j = 0; j < QPT; ++j) d[s][j] = acc[s][j] * aa[s];
                }
                return;
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                const Contrib<float> cn = ring[ent[s]].cn;
                evalN<QPT>(cn, nullptr, qt, Q, d[s]);
            }
            return;
        }
        float rr[NS], aa[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float2 ra = *(const float2 *)&ring[ent[s]].cn.c[1];
            rr[s] = ra.x; aa[s] = ra.y;
        }
        float rmin = rr[0];
#pragma unroll
        for (int s = 1; s < NS; ++s) rmin = fminf(rmin, rr[s]);
        float qmin = Q[0].q;
#pragma unroll
        for (int j = 1; j < QPT; ++j) qmin = fminf(qmin, Q[j].q);
        float q[QPT], th[QPT], tl[QPT];
#pragma unroll
        for (int j = 0; j < QPT; ++j) { q[j] = Q[j].q; th[j] = Q[j].th; tl[j] = Q[j].tl; }
        if (__any_sync(0xffffffffu, qmin * rmin < 1.f)) {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                float r[QPT], g[QPT];
#pragma unroll
                for (int j = 0; j < QPT; ++j) r[j] = rr[s];
                sph_g_vec<QPT, true>(q, th, tl, r, g);
#pragma unroll
                for (int j = 0; j < QPT; ++j) d[s][j] = (g[j] * g[j]) * aa[s];
            }
        } else {
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                float r[QPT], g[QPT];
#pragma unroll
                for (int j = 0; j < QPT; ++j) r[j] = rr[s];
                sph_g_vec<QPT, false>(q, th, tl, r, g);
#pragma unroll
                for (int j = 0; j < QPT; ++j) d[s][j] = (g[j] * g[j]) * aa[s];
            }
        }
    }
    // NS proposals against their cached rows: dn = g^2 A (the new row, what evalN returns), dd = dn - od.
    // Packed over pairs of q points (even QPT); the series variant is selected by ONE vote for the
    // whole chunk, so the NS x QPT evaluations form a single basic block.
    template <int QPT, int NS>
    __device__ __forceinline__ static void diffChunk(const RingEntry<float> *ring, const int (&ent)[NS],
                                                     const float *nodes, size_t nstr, const QuadTab<float> &qt,
                                                     const QRep<float> (&Q)[QPT], const float (&od)[NS][QPT],
                                                     float (&dn)[NS][QPT], float (&dd)[NS][QPT])
    {
        if (QPT % 2 != 0 || qt.nsm) {
            evalChunk<QPT, NS>(ring, ent, nodes, nstr, qt, Q, dn);
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int j = 0; j < QPT; ++j) dd[s][j] = __fsub_rn(dn[s][j], od[s][j]);
            return;
        }
        constexpr int MP = QPT / 2 > 0 ? QPT / 2 : 1;
        float rr[NS], aa[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const float2 ra = *(const float2 *)&ring[ent[s]].cn.c[1];
            rr[s] = ra.x; aa[s] = ra.y;
        }
        float rmin = rr[0];
#pragma unroll
        for (int s = 1; s < NS; ++s) rmin = fminf(rmin, rr[s]);
        float qmin = Q[0].q;
#pragma unroll
        for (int j = 1; j < QPT; ++j) qmin = fminf(qmin, Q[j].q);
        float2 q[MP], th[MP], tl[MP], g[NS][MP];
#pragma unroll
        for (int i = 0; i < MP; ++i) {
            q[i] = pk2(Q[2 * i].q, Q[(2 * i + 1) % QPT].q);
            th[i] = pk2(Q[2 * i].th, Q[(2 * i + 1) % QPT].th);
            tl[i] = pk2(Q[2 * i].tl, Q[(2 * i + 1) % QPT].tl);
        }
        if (__any_sync(0xffffffffu, qmin * rmin < 1.f)) sph_g_pack<NS, MP, true>(q, th, tl, rr, g);
        else sph_g_pack<NS, MP, false>(q, th, tl, rr, g);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int i = 0; i < MP; ++i) {
                const float2 n2 = __fmul2_rn(__fmul2_rn(g[s][i], g[s][i]), pk1(aa[s]));
                dn[s][2 * i] = n2.x; dn[s][(2 * i + 1) % QPT] = n2.y;
                const float2 d2 = sub2_exact(n2, pk2(od[s][2 * i], od[s][(2 * i + 1) % QPT]));
                dd[s][2 * i] = d2.x; dd[s][(2 * i + 1) % QPT] = d2.y;
            }
    }
    template <int QPT>
    __device__ __forceinline__ static void diffS(const Contrib<float> &cn, const Contrib<float> &co,
                                                 const float *, const float *, const QuadTab<float> &qt,
                                                 const QRep<float> (&Q)[QPT], float (&d)[QPT])
    {
        if (qt.nsm) {
#pragma unroll
            for (int j = 0; j < QPT; ++j)
                d[j] = eval_q<MCSAS_SPHERE, float>(cn, nullptr, qt, Q[j]) - eval_q<MCSAS_SPHERE, float>(co, nullptr, qt, Q[j]);
            return;
        }
        float q[2 * QPT], th[2 * QPT], tl[2 * QPT], r[2 * QPT], g[2 * QPT];
        const float rmin = fminf(cn.c[0], co.c[0]);
        bool small = false;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            q[j] = Q[j].q; th[j] = Q[j].th; tl[j] = Q[j].tl; r[j] = cn.c[0];
            q[QPT + j] = Q[j].q; th[QPT + j] = Q[j].th; tl[QPT + j] = Q[j].tl; r[QPT + j] = co.c[0];
            small |= Q[j].q * rmin < 1.f;
        }
        if (__any_sync(0xffffffffu, small)) sph_g_vec<2 * QPT, true>(q, th, tl, r, g);
        else sph_g_vec<2 * QPT, false>(q, th, tl, r, g);
#pragma unroll
        for (int j = 0; j < 2 * QPT; ++j) g[j] = g[j] * g[j];
#pragma unroll
        for (int j = 0; j < QPT; ++j) g[QPT + j] = g[QPT + j] * co.c[2];
#pragma unroll
        for (int j = 0; j < QPT; ++j) d[j] = fmaf(g[j], cn.c[2], -g[QPT + j]);
    }
};

// one thread's slice of a cached contribution row: QPT contiguous floats, one vector access.
// `keep` selects the L2 eviction priority of the line (createpolicy encodings of evict_last / evict_first):
// with more resident chains than the L2 holds rows for, the cyclic sweep over a chain's rows is the LRU
// worst case (hit rate ~ 30 %); the scheduler marks the first a.keep_rows rows of every chain -- what fits --
// evict_last and streams the rest, so that the marked part stays resident.  The policy operand of a load is
// warp-uniform (ptxas moves it to a uniform register), and the lanes of a chunk load read FGRP consecutive rows:
// a chunk takes the policy of its first row.
__device__ __forceinline__ unsigned long long row_policy(bool keep)
{
    return (unsigned long long)(keep ? 0x14F00000u : 0x12F00000u) << 32;
}
template <int QPT> __device__ __forceinline__ void row_load(const float *p, float (&v)[QPT], unsigned long long pol)
{
    if (QPT == 4)
        asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=f"(v[0]), "=f"(v[1 % QPT]), "=f"(v[2 % QPT]), "=f"(v[3 % QPT]) : "l"(p), "l"(pol));
    else if (QPT == 2)
        asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
                     : "=f"(v[0]), "=f"(v[1 % QPT]) : "l"(p), "l"(pol));
    else {
#pragma unroll
        for (int j = 0; j < QPT; ++j)
            asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v[j]) : "l"(p + j), "l"(pol));
    }
}
template <int QPT> __device__ __forceinline__ void row_store(float *p, const float (&v)[QPT], unsigned long long pol)
{
    if (QPT == 4)
        asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                     :: "l"(p), "f"(v[0]), "f"(v[1 % QPT]), "f"(v[2 % QPT]), "f"(v[3 % QPT]), "l"(pol) : "memory");
    else if (QPT == 2)
        asm volatile("st.global.cg.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;"
                     :: "l"(p), "f"(v[0]), "f"(v[1 % QPT]), "l"(pol) : "memory");
    else {
#pragma unroll
        for (int j = 0; j < QPT; ++j)
            asm volatile("st.global.cg.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p + j), "f"(v[j]), "l"(pol) : "memory");
    }
}

// Per-phase cycle counters (tools/prof_phases.py) are compiled in only with -DMCSAS_B200_PROF
// (make PROF=1): their time stamps would otherwise occupy registers of the evaluation loop.
#ifdef MCSAS_B200_PROF
#define FAST_PROF (a.prof != nullptr)
#else
#define FAST_PROF false
#endif

// Loop state that only thread 0 and the final write-back need, kept in shared memory: the evaluation
// needs every register (a chunk's old rows, new rows and sums next to the q points' constants).
struct FastCold {
    long long moves, evals, steps_all;
    double conval0, G;
    unsigned long long t_start, deadline;        // of this launch (deadline 0: not set yet / no budget)
    LsqConst L;
};

struct FastDec {
    int kstar;
    int wd;
    int accmask;
    int used;                    // Gram batch: steps consumed (< FGRP when the criterion is reached inside it)
    double a, b, conval;
};

// index of the pair (t, s), t < s, in the packed upper triangle of an n x n matrix
template <int NN> __host__ __device__ constexpr int tri_idx(int t, int s)
{
    return t < s ? t * NN - t * (t + 1) / 2 + (s - t - 1) : s * NN - s * (s + 1) / 2 + (t - s - 1);
}
__device__ __forceinline__ int tri_idx_rt(int n, int t, int s)
{
    return t * n - t * (t + 1) / 2 + (s - t - 1);
}

// Minimum of the exact quadratic chi^2 of a proposal with delta sums (s1..s4) around the centre
// (a0,b0) of a state with base sums S, Sm, Smm, E0 = sum w e^2, E1 = sum w e M, E2 = sum w e (see
// file header).  fp64: the direct change of chi^2 and the gain of the re-fit cancel to the
// ~1e-8 level on high signal-to-noise data (an fp32 version of this solve was measured no
// faster -- the decision is bound by its cold code, not by arithmetic latency -- and flipped
// accept decisions on the 0.01 %-uncertainty golden set).
struct FastBase { double S, Sm, Smm, E0, E1, E2, a0, b0; };
__device__ __forceinline__ void fast_solve(const FastBase &B, int find_bg, int pos_bg, double s1, double s2,
                                           double s3, double s4, double &rss, double &an, double &bn,
                                           double &e1n, double &e2n)
{
    const double Smp = B.Sm + s4, Smmp = B.Smm + 2. * s3 + s2;
    const double f0 = B.E0 - 2. * B.a0 * s1 + B.a0 * B.a0 * s2;
    const double ga = -2. * (B.E1 + s1 - B.a0 * (s3 + s2));
    const double gb = -2. * (B.E2 - B.a0 * s4);
    const double Haa = 2. * Smmp, Hab = 2. * Smp, Hbb = 2. * B.S;
    double da, db = 0.;
    if (find_bg) {
        const double inv = 1. / (Haa * Hbb - Hab * Hab);
        da = -(Hbb * ga - Hab * gb) * inv;
        db = -(Haa * gb - Hab * ga) * inv;
        rss = f0 + .5 * (ga * da + gb * db);
        if (pos_bg && B.b0 + db < 0.) {                          // |b| residual -> b = 0 (:60-63)
            db = -B.b0;
            da = -(ga + Hab * db) / Haa;
            rss = f0 + ga * da + gb * db + .5 * (Haa * da * da + 2. * Hab * da * db + Hbb * db * db);
        }
    } else {                                                     // chiNoBg (:66-70), b0 == 0
        da = -ga / Haa;
        rss = f0 + .5 * ga * da;
    }
    an = B.a0 + da;
    bn = B.b0 + db;
    // sum w e M and sum w e at the new centre (zero at an unconstrained minimum)
    e1n = -.5 * (ga + Haa * da + Hab * db);
    e2n = -.5 * (gb + Hab * da + Hbb * db);
}

template <int NG> __device__ __forceinline__ int grp_perm(int lane)
{
    return NG == 8 ? ((lane >> 2) & 7) : ((lane >> 3) & 3);
}

// ---- exchange of per-warp partial sums -------------------------------------------------------
// Every CTA of a chain's cluster receives the partial sums of EVERY warp of the chain (global
// warp gw = rank * warps_q + wq covers the q points [32 QPT gw, 32 QPT (gw + 1))) and adds
// them in ascending gw order, so the totals -- and with them every accept decision -- are
// bit-identical for every cluster size (a chain's result does not depend on how the host
// scheduler packs it).  C > 1: st.async into all ranks (own CTA included), completion counted
// in bytes on the receiver's mbarrier; C == 1: plain shared-memory stores and one mbarrier
// arrival per warp.  No CTA-wide barrier between the evaluation and the decision.
__device__ __forceinline__ void xsend16(float4 *dst, uint64_t *bar, int C, const float4 v)
{
    if (C == 1) { *dst = v; return; }
    const uint32_t la = smem_u32(dst), lb = smem_u32(bar);
    for (int r = 0; r < C; ++r) {
        uint32_t ra, rb;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(r));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(r));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(ra), "r"(__float_as_int(v.x)), "r"(__float_as_int(v.y)), "r"(__float_as_int(v.z)),
                       "r"(__float_as_int(v.w)), "r"(rb) : "memory");
    }
}
__device__ __forceinline__ void xsend8(double *dst, uint64_t *bar, int C, const double v)
{
    if (C == 1) { *dst = v; return; }
    for (int r = 0; r < C; ++r) dsmem_store_async(dst, bar, (uint32_t)r, v);
}
__device__ __forceinline__ void mbar_arrive_release(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// CACHE: the scaled intensity rows D_i(q) of the N current contributions live in global
// memory (L2 resident: N x nq floats per chain), so the "old" half of every evaluation
// (mcsas.py:362 recomputes it) is one coalesced load; rows are rebuilt at the start of every
// launch and rewritten when a move is accepted.
// SMEAR: instantiation for data with a smearing profile (sasmodel.py:56-73); the unsmeared
// instantiation folds every smearing branch away at compile time.
template <int MODEL, int QPT, bool CACHE, bool SMEAR>
__global__ void __launch_bounds__(QPT == 1 ? 1024 : (QPT == 2 ? 512 : 256), 1)
chain_fast_kernel(const __grid_constant__ ChainArgs a)
{
    using FO = FastOps<MODEL>;
    using O = Ops<MODEL, float>;
    constexpr int NARR = O::NARR;
    constexpr int FGRP = QPT == 1 ? 4 : 8;         // proposals per reduce-scatter chunk
    constexpr int NGP = FGRP * (FGRP - 1) / 2;     // pairs of a chunk (Gram matrix, off diagonal)
    constexpr int NH = FGRP == 8 ? 4 : FGRP;       // proposals evaluated together (register budget)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = blockDim.x;                      // all warps evaluate; the last one also fills the ring
    const int W = T >> 5;
    const int WQ = a.warps_q;                      // warps that cover this CTA's q slice once
    const int PS = W / WQ;                         // proposal groups
    const int grp = warp / WQ, wq = warp % WQ;
    const int TQ = WQ * 32, qtid = wq * 32 + lane;
    const int C = a.cluster;
    const int rank = C > 1 ? (int)cluster_ctarank() : 0;
    const int GW = C * WQ, gw = rank * WQ + wq;    // warps of the whole chain / this warp's place among them
    const int slot = blockIdx.x / C;
    const int chain_id = a.chain_ids ? a.chain_ids[slot] : slot;
    const int chain = chain_id & ~CHAIN_ID_BACKUP;
    const bool is_backup = (chain_id & CHAIN_ID_BACKUP) != 0;   // a backup attempt: its launch ends with the primaries'
    // `chain` is a state slot: slots >= n_chains hold a chain's other attempt (retry prepared or
    // run ahead of time by the host scheduler, capi.cu)
    const int base_chain = chain % a.n_chains;
    const uint32_t gchain = (uint32_t)(a.first_chain + base_chain);
    const int set = base_chain / a.n_reps;
    const DatasetDesc ds = a.sets[set];
    const int nq = ds.nq, N = a.n_contribs, P = a.mp.n_active;
    const int D = NARR ? a.mp.quad_nodes : 0;
    const int RING = a.ring;
    const int kmax = a.spec_max;
    const int keep_rows = a.keep_rows;             // cached rows [0, keep_rows) are evict_last in the L2
    const int RL = QPT * TQ;                       // cached row length of this CTA
    // row regions: FAST_CACHE_SLOTS per SM (an 8-warp CTA fills the register file of its SM; two
    // 4-warp CTAs share one), taken under a lock word for the lifetime of the CTA
    unsigned smid = 0;
    if (CACHE) {
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        smid %= FAST_CACHE_SMS;                    // (the lock keeps aliased SMs apart; B200: 148 SMs)
    }

    // ---- shared memory carve-up (keep in sync with fast_smem_bytes) -----------------------
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *sp = smem_raw;
    auto carve = [&](size_t bytes) { unsigned char *p = sp; sp += (bytes + 15) & ~size_t(15); return p; };
    const int GWC = a.gw_cap;                      // >= GW of every layout the host may pick for this nq
    uint64_t *s_bar = (uint64_t *)carve(16);
    double *s_data = (double *)carve(sizeof(double) * 3 * ds.nq_pad);
    double *s_rset = (double *)carve(sizeof(double) * N * (P > 0 ? P : 1));
    int *s_ctrl = (int *)carve(64);
    RingEntry<float> *s_ring = (RingEntry<float> *)carve((size_t)RING * sizeof(RingEntry<float>));
    const int XPC = a.xp_cap;                      // proposals per batch the exchange buffers hold
    float4 *s_xp = (float4 *)carve(16 * 2 * (size_t)GWC * XPC);   // [buf][gw][proposal] (s1..s4)
    double *s_xb = (double *)carve(8 * 2 * (size_t)GWC * 8);            // [buf][gw][5 base sums, flags]
    float *s_xg = (float *)carve(4 * 2 * (size_t)GWC * 32);             // [buf][gw][Gram pairs]
    uint64_t *s_xbar = (uint64_t *)carve(16);
    float *s_gtot = (float *)carve(4 * 32);
    double *s_chi = (double *)carve(8 * (size_t)FAST_KCAP + 256);
    FastDec *s_dec = (FastDec *)(s_chi + FAST_KCAP);
    FastBase *s_B = (FastBase *)((char *)s_dec + 64);             // base sums of the current state (warp 0's)
    FastCold *s_cold = (FastCold *)((char *)s_dec + 128);
    static_assert(sizeof(FastDec) <= 64 && sizeof(FastBase) <= 64 && sizeof(FastCold) <= 128, "decision block layout");
    double *s_apart = (double *)carve(8 * 2 * (size_t)GWC * 3);         // [buf][gw][3] full-LSQ all-sum
    double *s_ft = (double *)carve(8 * 2 * FAST_QCAP);            // model intensity, double buffered
    double *s_In = (double *)carve(8 * FAST_QCAP);                // I / i_scale
    double *s_sw = (double *)carve(8 * FAST_QCAP);                // i_scale / sigma (0: padding)
    float *s_qs = (float *)carve(4 * FAST_QCAP);                  // (q_ref/q)^QS_POW
    float *s_qx = (float *)carve(sizeof(float) * (D + 1));
    float *s_qw = (float *)carve(sizeof(float) * (D + 1));
    float *s_nodes = (float *)sp;                  // [RING][2][NARR][D]
    const size_t NSTR = (size_t)NARR * D;

    if (CACHE && tid == 0) {
        int got = -1;
        while (got < 0) {
            for (int sl = 0; sl < FAST_CACHE_SLOTS && got < 0; ++sl)
                if (atomicCAS(a.cache_lock + smid * FAST_CACHE_SLOTS + sl, 0, 1) == 0) got = (int)smid * FAST_CACHE_SLOTS + sl;
            if (got < 0) __nanosleep(200);
        }
        __threadfence();
        s_ctrl[0] = got;
    }
    // ---- stage the dataset once: TMA bulk copy global -> shared ---------------------------
    if (tid == 0) {
        mbar_init(s_bar, 1);
        mbar_init(s_xbar, C > 1 ? 1 : W);
        mbar_init(s_xbar + 1, C > 1 ? 1 : W);
    }
    unsigned xphase = 0;
    __syncthreads();
    const int region = CACHE ? s_ctrl[0] : 0;
    float *const crow = CACHE ? a.cache + (size_t)region * a.cache_stride + qtid * QPT : nullptr;
    if (tid == 0) {
        uint32_t bytes = (uint32_t)(sizeof(double) * 3 * ds.nq_pad);
        mbar_expect_tx(s_bar, bytes);
        tma_bulk_g2s(s_data, a.data + ds.offset, bytes, s_bar);
    }
    for (int j = tid; j < D; j += blockDim.x) {
        s_qx[j] = (float)a.quad_x[j];
        s_qw[j] = (float)a.quad_w[j];
    }
    mbar_wait(s_bar, 0);
    if (C > 1) cluster_sync_all(); else __syncthreads();         // every rank's mbarriers are initialised
    const QuadTab<float> qt = { D, a.mp.quad_rule, s_qx, s_qw, &a.smear, SMEAR ? a.smear.n : 0,
                                SMEAR && a.smloc ? a.smloc + ((size_t)set * a.smear.n * a.smloc_row + (size_t)(rank * TQ + qtid) * QPT) : nullptr,
                                a.smloc_row, SMEAR ? a.smloc_mono : 0 };
    const float qref = (float)(ds.q_ref * 1e-9);   // 1/nm

    // ---- per-thread resident state of q points (rank*TQ + qtid)*QPT + j, j < QPT (contiguous:
    //      a thread's row slice is one 4/8/16-byte load) --------------------------------------
    // intensities are normalised by i_scale; sw = 1/sigma, In = I (normalised), ft = model (fp64)
    QRep<float> Q[QPT];
    float cq[QPT], rho[QPT], mu[QPT], om[QPT];
    int cur = 0;                                   // s_ft buffer holding the current state
    const double inorm = 1. / ds.i_scale;
    double v3[3] = {0., 0., 0.};
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
        int i = (rank * TQ + qtid) * QPT + j;
        bool valid = i < nq;
        int ii = (i < nq) ? i : nq - 1;
        double qv = s_data[ii], Iv = s_data[ds.nq_pad + ii] * inorm, ev = s_data[2 * ds.nq_pad + ii] * inorm;
        if (ev == 0.0) ev = 1. * inorm;                          // backgroundscalingfit.py:117
        Q[j].set(qv);
        double qsd = 1.;
        if (FO::QS_POW > 0 && qt.nsm == 0) {                  // (smeared: locations differ per profile point)
            double t = (double)qref / (qv * 1e-9);
            double t2 = t * t;
            qsd = FO::QS_POW == 6 ? t2 * t2 * t2 : (FO::QS_POW == 4 ? t2 * t2 : t2);
        }
        const float qsf = (float)qsd;
        const double swd = valid ? 1. / ev : 0.;
        cq[j] = (float)(swd * (double)qsf);
        om[j] = (float)swd;
        rho[j] = 0.f; mu[j] = 0.f;
        if (grp == 0) {
            const int ql = qtid * QPT + j;
            s_qs[ql] = qsf; s_sw[ql] = swd; s_In[ql] = Iv;
            double w = swd * swd;
            v3[0] += w; v3[1] += w * Iv; v3[2] += w * Iv * Iv;
        }
    }

    // Sum three per-thread partials (group 0 holds them) over the whole chain, in ascending
    // global-warp order on every rank (all threads call it)
    int sbuf = 0;
    auto cluster_allsum3 = [&](double (&v)[3]) {
#pragma unroll
        for (int i = 0; i < 3; ++i) v[i] = warp_sum(v[i]);
        double *part = s_apart + (size_t)sbuf * GWC * 3;
        if (grp == 0 && lane < 3) {
            const double x = lane == 0 ? v[0] : (lane == 1 ? v[1] : v[2]);
            if (C > 1) { for (int r = 0; r < C; ++r) dsmem_store(part + gw * 3 + lane, (uint32_t)r, x); }
            else part[gw * 3 + lane] = x;
        }
        if (C > 1) cluster_sync_all(); else __syncthreads();
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double x = 0.;
            for (int g = 0; g < GW; ++g) x += part[g * 3 + i];
            v[i] = x;
        }
        sbuf ^= 1;
    };

    cluster_allsum3(v3);
    if (tid == 0) {
        LsqConst L0;
        L0.S = v3[0]; L0.Si = v3[1]; L0.Sii = v3[2];
        L0.inv_nq = 1. / (double)nq;
        L0.find_bg = a.find_bg; L0.pos_bg = a.pos_bg;
        s_cold->L = L0;
        s_cold->t_start = globaltimer_ns();
        s_cold->deadline = 0;
    }
    const double inv_nq = 1. / (double)nq;

    // ---- fp32 amplitude scale G: 1 / amplitude at the upper end of the active ranges ------
    {
        if (tid == 0) {
            PVec p(a.mp, a.mp.hi);
            Contrib<float> tmp;
            VWS o = O::template prep<false>(p, a.comp_exp, 1., qt, tmp, s_nodes);
            double g = 1. / o.amp_raw;
            if (!(g > 0.) || isinf(g)) g = 1.;
            s_cold->G = g;
        }
        __syncthreads();
    }
#define FAST_G (*(volatile const double *)&s_cold->G)

    // exact fp64 fit of the resident ft over the cluster (initial and final fit)
    auto lsq_all = [&](double &sa, double &sb, double &chi) {
        double v[3] = {0., 0., 0.};
        if (grp == 0) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const int ql = qtid * QPT + j;
                const double f = s_ft[cur * FAST_QCAP + ql], sq = s_sw[ql];
                double w = sq * sq, t = w * f;
                v[0] += t; v[1] = fma(t, f, v[1]); v[2] = fma(t, s_In[ql], v[2]);
            }
        }
        cluster_allsum3(v);
        {
            const LsqConst L = s_cold->L;
            lsq_solve(L, v[0], v[1], v[2], sa, sb, chi);
        }
        double r2[3] = {0., 0., 0.};                             // residual form (:73-77,:135)
        if (grp == 0) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const int ql = qtid * QPT + j;
                double r = s_sw[ql] * (s_In[ql] - (s_ft[cur * FAST_QCAP + ql] * sa + sb));
                r2[0] = fma(r, r, r2[0]);
            }
        }
        cluster_allsum3(r2);
        chi = r2[0] * inv_nq;
    };

    auto fill_entry = [&](long long t, int attempt) {
        const int e = (int)t & (RING - 1);                       // RING is a power of two
        RingEntry<float> &E = s_ring[e];
        uint32_t o4[4];
        if (!a.stream) philox4x32_10(a.seed, (uint64_t)t, gchain, 2u * attempt, o4);
#pragma unroll
        for (int k = 0; k < MCSAS_MAX_ACTIVE; ++k) {
            if (k < P) {
                double u = a.stream ? a.stream[(long long)chain * a.stream_stride + (a.start_min ? 0 : (long long)N * P) + t * P + k]
                                    : generator_output(o4[k], a.mp.gen_kind[k]);
                E.prop[k] = u * (a.mp.hi[k] - a.mp.lo[k]) + a.mp.lo[k];  // mcsas.py:358
            }
        }
        PVec pn(a.mp, E.prop);
        const double G = FAST_G;
        FO::prep(pn, a.comp_exp, G, qref, qt, E.cn, s_nodes + (size_t)(e * 2) * NSTR);
        if (!CACHE) {                                            // cached rows replace the old contribution
            PVec po(a.mp, s_rset + (int)((uint32_t)t % (uint32_t)N) * P);   // max_iter < 2^31 (host)
            FO::prep(po, a.comp_exp, G, qref, qt, E.co, s_nodes + (size_t)(e * 2 + 1) * NSTR);
        }
    };

    long long step = 0;
    double conval = 0., sc_a = 1., sc_b = 0.;
    if (tid == 0) { s_cold->moves = 0; s_cold->evals = 0; s_cold->steps_all = 0; s_cold->conval0 = 0.; }
    double a0 = 1., b0 = 0.;                       // centre of the quadratic expansion (fit of the last accepted state)
    double cv_run = 0.;                            // chi-squared the accept test compares against
    // Batch kind and size.  Which batches are Gram batches follows from the positions of the
    // accepted steps alone (a chain-intrinsic rule: results must not depend on the schedule):
    // a move less than FGRP steps behind the previous one switches to Gram batches, a Gram batch
    // without a move switches back.  The size K of a speculative batch is free (it never changes
    // a result) and adapts to the distance between moves.
    long long last_acc = -(1LL << 40);             // step of the last accepted move
    bool gmode = true;
    int K = max(1, kmax / 2);
    int attempt = 0, status = 1;
    bool resume = a.resume != 0 && a.status[chain] == 0;
    if (a.resume) {
        if (tid == 0) { s_cold->evals = a.evals[chain]; s_cold->steps_all = a.steps_all[chain]; }
        attempt = a.attempts[chain] - 1;
        if (resume) {
            step = a.num_iter[chain];
            if (tid == 0) { s_cold->moves = a.num_moves[chain]; s_cold->conval0 = a.conval0[chain]; }
        }
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
        double ft0[QPT];
#pragma unroll
        for (int j = 0; j < QPT; ++j) ft0[j] = 0.;
        cur = 0;
        __syncthreads();
        // ---- ft = sum of all contributions, in contribution order (mcsas.py:319); the cached
        //      rows of this launch ------------------------------------------------------------
        for (int i0 = 0; i0 < N; i0 += RING) {
            int nb = min(RING, N - i0);
            for (int e = tid; e < nb; e += blockDim.x) {
                PVec p(a.mp, s_rset + (i0 + e) * P);
                FO::prep(p, a.comp_exp, FAST_G, qref, qt, s_ring[e].co, s_nodes + (size_t)(e * 2 + 1) * NSTR);
            }
            __syncthreads();
            {
                for (int e = 0; e < nb; ++e) {
                    const Contrib<float> cc = s_ring[e].co;
                    const float *nd = s_nodes + (size_t)(e * 2 + 1) * NSTR;
                    float dv[QPT];
                    FO::template evalN<QPT>(cc, nd, qt, Q, dv);
#pragma unroll
                    for (int j = 0; j < QPT; ++j) ft0[j] += (double)dv[j] * (double)s_qs[qtid * QPT + j];
                    if (CACHE && grp == 0) row_store<QPT>(crow + (size_t)(i0 + e) * RL, dv, row_policy(i0 + e < keep_rows));
                }
            }
            __syncthreads();
        }
        if (!resume && tid == 0) s_cold->evals += N;             // (rebuilding the rows of a resumed chain is not counted)
        // a launch's time budget is at least 16 x what building the rows cost it (orientation-averaged
        // and integral models: N evaluations are many MC steps' worth), so re-packing stays cheap
        if (a.slice_ns > 0 && tid == 0 && s_cold->deadline == 0) {
            const unsigned long long now = globaltimer_ns(), t_rows = now - s_cold->t_start;
            s_cold->deadline = s_cold->t_start + max((unsigned long long)a.slice_ns, 16ULL * t_rows);
        }
        if (resume) {
            // continue exactly where the previous launch stopped: the fp64 master intensity, the
            // centre of the expansion, the running chi-squared and the batch size are part of the
            // saved state, so a slice boundary leaves no trace in the chain
#pragma unroll
            for (int j = 0; j < QPT; ++j) {
                const int i = (rank * TQ + qtid) * QPT + j;
                ft0[j] = i < nq ? a.ft_save[(long long)chain * a.nq_stride + i] : 0.;
            }
        }
        if (grp == 0) {
#pragma unroll
            for (int j = 0; j < QPT; ++j) s_ft[qtid * QPT + j] = ft0[j];
        }
        if (resume) {
            a0 = a.cen_a[chain]; b0 = a.cen_b[chain]; cv_run = a.cv_run[chain];
            last_acc = (long long)a.last_acc[chain];
            K = a.batch_k[chain] & 0xffff;
            gmode = (a.batch_k[chain] >> 16) != 0;
            __syncthreads();
        } else {
            lsq_all(sc_a, sc_b, conval);                         // mcsas.py:339-344
            step = 0;
            if (tid == 0) { s_cold->conval0 = conval; s_cold->moves = 0; }
            a0 = sc_a; b0 = sc_b; cv_run = conval;
            K = max(1, kmax / 2);
            gmode = true; last_acc = -(1LL << 40);
        }
        resume = false;
        // residual / model in units of sigma around the centre
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            const int ql = qtid * QPT + j;
            const double sq = s_sw[ql];
            rho[j] = (float)(sq * (s_In[ql] - (a0 * ft0[j] + b0)));
            mu[j] = (float)(sq * ft0[j]);
        }
        bool dirty = true;                                       // base sums not yet accumulated
        // ---- first ring fill: the next two batches' worth of steps ------------------------------
        long long filled = min(step + 2LL * kmax, a.max_iter);   // ring holds steps [.., filled)
        for (long long t = step + tid; t < filled; t += blockDim.x) fill_entry(t, attempt);
        __syncthreads();
        int batch = 0;
        const int wd_batches = (int)min((long long)0x3fffffff, 2 * (min(a.max_iter, slice_end) - step) + 64);
        const int batch_end = a.slice_batches > 0 ? a.slice_batches : 0x7fffffff;
        bool slice_over = false;                                 // this launch's budget (steps, batches, time) is used up
        conval = cv_run;
        bool live = N > 1 && conval > a.conv_crit && step < a.max_iter && step < slice_end;

        // ---- the MC loop (mcsas.py:354-404) ---------------------------------------------------
        // Two kinds of batches: SPECULATIVE (Kb <= 64 steps evaluated against the current state,
        // committed up to the first accepted one) when moves are rare, and GRAM (exactly FGRP
        // steps: the Gram matrix of the proposals' changes lets warp 0 re-derive the sums of
        // later proposals after each accepted one) when moves are frequent.
        // (the base sums live in shared memory: only warp 0's decision uses them, and the evaluation
        // needs every register)
        if (tid == 0) { FastBase B0; B0.S = s_cold->L.S; B0.Sm = 0.; B0.Smm = 0.; B0.E0 = 0.; B0.E1 = 0.; B0.E2 = 0.; B0.a0 = a0; B0.b0 = b0; *s_B = B0; }
        int xb_ = 0;                                             // exchange buffer of the next batch
        for (;;) {
        while (live) {
            const long long pt0 = FAST_PROF ? clock64() : 0;
            // a speculative batch smaller than one chunk per proposal group would take as long
            int Kb = min(max(K, PS * FGRP), kmax);
            bool gram = a.gram && gmode && kmax >= FGRP;
            if (gram) Kb = FGRP;
            {
                long long rem = a.max_iter - step;
                if (rem < Kb) { Kb = rem >= 1 ? (1 << (31 - __clz((int)rem))) : 0; gram = false; }   // power of two
            }
            // The exact base sums after a move are accumulated by the LAST proposal group, so that warp 0
            // (which decides) is not the one that starts its evaluation late; a speculative batch that
            // would give every group the same number of chunks is one chunk shorter then.
            const int rgrp = PS - 1;
            if (dirty && PS >= 2 && !gram && Kb >= PS * FGRP && Kb % (PS * FGRP) == 0 && Kb > FGRP) Kb -= FGRP;
            ++batch;
            if (Kb < 1 || batch > wd_batches) { status = -2; break; }   // watchdog (uniform over the cluster)
            const int ri0 = (int)((uint32_t)step % (uint32_t)N);  // contribution replaced by the batch's first step
            const int xb = xb_;
            float4 *const xp = s_xp + (size_t)xb * GWC * XPC + (size_t)gw * XPC;
            double *const xbs = s_xb + (size_t)xb * GWC * 8 + (size_t)gw * 8;
            uint64_t *const xbar = s_xbar + xb;
            const long long pt1 = FAST_PROF ? clock64() : 0;
            // ---- exact base sums of the current state (after an accepted move) + this rank's flags
            if (grp == rgrp && dirty) {
                double x = 0.;
                {
                    double b5[5] = {0., 0., 0., 0., 0.};
#pragma unroll
                    for (int j = 0; j < QPT; ++j) {
                        const int ql = qtid * QPT + j;
                        const double sq = s_sw[ql];
                        double m = sq * s_ft[cur * FAST_QCAP + ql];
                        double r = sq * s_In[ql] - (a0 * m + b0 * sq);
                        b5[0] = fma(r, r, b5[0]);                    // E0 = sum w e^2
                        b5[1] = fma(r, m, b5[1]);                    // E1 = sum w e ft
                        b5[2] = fma(r, sq, b5[2]);                   // E2 = sum w e
                        b5[3] = fma(m, sq, b5[3]);                   // Sm
                        b5[4] = fma(m, m, b5[4]);                    // Smm
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) b5[i] = warp_sum(b5[i]);
                    x = b5[0];
                    if (lane == 1) x = b5[1];
                    if (lane == 2) x = b5[2];
                    if (lane == 3) x = b5[3];
                    if (lane == 4) x = b5[4];
                    if (lane < 5) xsend8(xbs + lane, xbar, C, x);
                }
            }
            if (grp == 0) {
                if (lane == 5) {                                 // time budget of this launch (any warp may end it)
                    bool over = a.slice_ns > 0 && globaltimer_ns() > *(volatile unsigned long long *)&s_cold->deadline;
                    if (is_backup && a.needed_left && *(volatile int *)a.needed_left <= 0) over = true;
                    if (a.repack_ctas > 0 && *(volatile int *)(a.needed_left + 1) >= a.repack_ctas) over = true;
                    xsend8(xbs + 5, xbar, C, over ? 1. : 0.);
                }
            }
            const long long pt1a = FAST_PROF ? clock64() : 0;
            long long ptc[4] = {0, 0, 0, 0};
            // ---- evaluate the batch ---------------------------------------------------------
            // pinned in a register: ptxas otherwise re-reads %tid (S2R) for every proposal slot
            int perm = grp_perm<FGRP>(lane);
            asm volatile("mov.b32 %0, %0;" : "+r"(perm));
            const unsigned long long gpol = row_policy(ri0 < keep_rows);   // L2 policy of a Gram batch's rows
            float uu[FGRP][QPT];                                 // GRAM: (new - old)/sigma of slot s
            float dnk[CACHE ? FGRP : 1][QPT];                    // GRAM + CACHE: new row of slot s
            if (gram) {
                // every group evaluates the chunk (all need uu for the update); group 0 reports
                const int e0 = (int)step & (RING - 1);
                float acc[FGRP][4], gq[NGP];
#pragma unroll
                for (int i = 0; i < NGP; ++i) gq[i] = 0.f;
#pragma unroll
                for (int h0 = 0; h0 < FGRP; h0 += NH) {
                    float dd[NH][QPT];
                    if (CACHE) {
                        float od[NH][QPT];
                        int ent[NH];
#pragma unroll
                        for (int s = 0; s < NH; ++s) {
                            int ri = ri0 + ((h0 + s) ^ perm);
                            if (ri >= N) ri -= N;
                            row_load<QPT>(crow + (size_t)ri * RL, od[s], gpol);
                            ent[s] = (e0 + ((h0 + s) ^ perm)) & (RING - 1);
                        }
                        FO::template diffChunk<QPT, NH>(s_ring, ent, s_nodes, NSTR, qt, Q, od,
                                                        *reinterpret_cast<float (*)[NH][QPT]>(&dnk[CACHE ? h0 : 0]), dd);
                    } else {
#pragma unroll
                        for (int s = 0; s < NH; ++s) {
                            const int e = (e0 + ((h0 + s) ^ perm)) & (RING - 1);
                            const Contrib<float> cn = s_ring[e].cn, co = s_ring[e].co;
                            const float *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
                            FO::template diffS<QPT>(cn, co, nn, no, qt, Q, dd[s]);
                        }
                    }
#pragma unroll
                    for (int s = h0; s < h0 + NH; ++s) {
                        fast_sums<QPT>(dd[s - h0], cq, rho, mu, om, acc[s], uu[s]);
#pragma unroll
                        for (int j = 0; j < QPT; ++j)
#pragma unroll
                            for (int t = 0; t < s; ++t)
                                gq[tri_idx<FGRP>(t, s)] = fmaf(uu[t][j], uu[s][j], gq[tri_idx<FGRP>(t, s)]);
                    }
                }
                if (grp == 0) {
#pragma unroll
                    for (int h = FGRP / 2, off = 16; h >= 1; h >>= 1, off >>= 1) {
#pragma unroll
                        for (int s = 0; s < h; ++s)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                acc[s][i] += __shfl_xor_sync(0xffffffffu, acc[s + h][i], off);
                        // Gram: all-reduce with the slot permutation of the partner lane
                        float tmp[NGP];
#pragma unroll
                        for (int t = 0; t < FGRP; ++t)
#pragma unroll
                            for (int s2 = t + 1; s2 < FGRP; ++s2)
                                tmp[tri_idx<FGRP>(t, s2)] =
                                    __shfl_xor_sync(0xffffffffu, gq[tri_idx<FGRP>(t ^ h, s2 ^ h)], off);
#pragma unroll
                        for (int i = 0; i < NGP; ++i) gq[i] += tmp[i];
                    }
#pragma unroll
                    for (int off = 32 / FGRP / 2; off >= 1; off >>= 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], off);
#pragma unroll
                        for (int i = 0; i < NGP; ++i) gq[i] += __shfl_xor_sync(0xffffffffu, gq[i], off);
                    }
                    if ((lane & (32 / FGRP - 1)) == 0)
                        xsend16(xp + perm, xbar, C, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
                    // every lane holds the warp's Gram pairs in natural proposal order only where
                    // perm == 0 (lanes 0..3): lane 0 publishes them, two pairs per 8-byte unit
                    if (lane == 0) {
                        double *xg = (double *)(s_xg + (size_t)xb * GWC * 32 + (size_t)gw * 32);
#pragma unroll
                        for (int i = 0; i < NGP; i += 2) {
                            const float g1 = i + 1 < NGP ? gq[(i + 1) % NGP] : 0.f;
                            xsend8(xg + (i >> 1), xbar, C, __hiloint2double(__float_as_int(g1), __float_as_int(gq[i])));
                        }
                    }
                }
            } else if (Kb >= FGRP) {
                const int e0 = (int)step & (RING - 1);
                for (int k0 = grp * FGRP; k0 < Kb; k0 += PS * FGRP) {
                    const long long ptk = FAST_PROF ? clock64() : 0;
                    float acc[FGRP][4];
                    // NH proposals at a time: the new rows and the evaluations in flight of a whole chunk do
                    // not fit the register file next to the loop's state (ptxas then serialises the
                    // evaluations at the end of the chunk); the old rows of the whole chunk are requested
                    // first (L2 / DRAM latency hides behind the evaluations)
                    float od[CACHE ? FGRP : 1][QPT];
                    int rb = ri0 + k0;
                    if (rb >= N) rb -= N;
                    const unsigned long long cpol = row_policy(rb < keep_rows);
                    if (CACHE) {
#pragma unroll
                        for (int s = 0; s < FGRP; ++s) {
                            int ri = ri0 + k0 + (s ^ perm);
                            if (ri >= N) ri -= N;
                            row_load<QPT>(crow + (size_t)ri * RL, od[CACHE ? s : 0], cpol);
                        }
                    }
#pragma unroll
                    for (int h0 = 0; h0 < FGRP; h0 += NH) {
                        float dd[NH][QPT];
                        if (CACHE) {
                            float dn[NH][QPT];
                            int ent[NH];
#pragma unroll
                            for (int s = 0; s < NH; ++s) ent[s] = (e0 + k0 + ((h0 + s) ^ perm)) & (RING - 1);
                            FO::template diffChunk<QPT, NH>(s_ring, ent, s_nodes, NSTR, qt, Q,
                                                            *reinterpret_cast<const float (*)[NH][QPT]>(&od[CACHE ? h0 : 0]), dn, dd);
                        } else {
#pragma unroll
                            for (int s = 0; s < NH; ++s) {
                                const int k = k0 + ((h0 + s) ^ perm);
                                const int e = (e0 + k) & (RING - 1);
                                const Contrib<float> cn = s_ring[e].cn, co = s_ring[e].co;
                                const float *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
                                FO::template diffS<QPT>(cn, co, nn, no, qt, Q, dd[s]);  // mcsas.py:360-362
                            }
                        }
#pragma unroll
                        for (int s = 0; s < NH; ++s) {
                            float u[QPT];                                // (new - old)/sigma, :366
                            fast_sums<QPT>(dd[s], cq, rho, mu, om, acc[h0 + s], u);
                        }
                    }
                    // select-free reduce-scatter: slot s of lane l holds proposal s ^ perm(l)
#pragma unroll
                    for (int h = FGRP / 2, off = 16; h >= 1; h >>= 1, off >>= 1) {
#pragma unroll
                        for (int s = 0; s < h; ++s)
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                acc[s][i] += __shfl_xor_sync(0xffffffffu, acc[s + h][i], off);
                    }
#pragma unroll
                    for (int off = 32 / FGRP / 2; off >= 1; off >>= 1)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], off);
                    if ((lane & (32 / FGRP - 1)) == 0)
                        xsend16(xp + k0 + perm, xbar, C, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
                    if (FAST_PROF) { ptc[k0 == 0 ? 0 : 1] += clock64() - ptk; ptc[k0 == 0 ? 2 : 3] += 1; }
                }
            } else if (grp == 0) {
                for (int k = 0; k < Kb; ++k) {
                    float acc[4];
                    const int e = ((int)step + k) & (RING - 1);
                    float dd[1][QPT];
                    if (CACHE) {
                        int ri = ri0 + k;
                        if (ri >= N) ri -= N;
                        float ov[1][QPT], dv[1][QPT];
                        row_load<QPT>(crow + (size_t)ri * RL, ov[0], row_policy(ri < keep_rows));
                        const int ent[1] = {e};
                        FO::template diffChunk<QPT, 1>(s_ring, ent, s_nodes, NSTR, qt, Q, ov, dv, dd);
                    } else {
                        const Contrib<float> cn = s_ring[e].cn, co = s_ring[e].co;
                        const float *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
                        FO::template diffS<QPT>(cn, co, nn, no, qt, Q, dd[0]);
                    }
                    float u[QPT];
                    fast_sums<QPT>(dd[0], cq, rho, mu, om, acc, u);
                    // the same shuffle tree as the chunk evaluator's (lane bits 4..0), so that a
                    // proposal's sums do not depend on the size of the batch it was evaluated in
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
                    if (lane == 0) xsend16(xp + k, xbar, C, make_float4(acc[0], acc[1], acc[2], acc[3]));
                }
            }
            if (C == 1) {                                        // this warp's part of the batch is in shared memory
                __syncwarp();
                if (lane == 0) mbar_arrive_release(xbar);
            }
            const long long pt2 = FAST_PROF ? clock64() : 0;
            // ---- the last warp prepares the ring entries of the NEXT batch while warp 0 decides:
            // steps [filled, step + Kb + kmax); their slots held steps <= step - RING/2 (consumed),
            // and the contributions they replace are not touched by this batch (Kb + kmax < N)
            const long long fill_from = filled, fill_to = min(step + Kb + (long long)kmax, a.max_iter);
            if (warp == W - 1 && W > 1)
                for (long long t = fill_from + lane; t < fill_to; t += 32) fill_entry(t, attempt);
            filled = max(filled, fill_to);
            long long pd[4] = {0, 0, 0, 0};
            const long long pt3 = FAST_PROF ? clock64() : 0;
            if (warp == 0) {
                if (C > 1 && lane == 0) {                        // bytes this rank receives per batch
                    const uint32_t per_gw = (uint32_t)(Kb * 16 + 8 + (dirty ? 40 : 0) + (gram ? ((NGP + 1) / 2) * 8 : 0));
                    mbar_expect_tx(xbar, (uint32_t)GW * per_gw);
                }
                mbar_wait(xbar, (xphase >> xb) & 1);             // every warp's partial sums have landed
                if (FAST_PROF) pd[0] = clock64();
                FastBase B = *s_B;
                const double *allb = s_xb + (size_t)xb * GWC * 8;
                {
                    // lane i < 5: base sum i (fresh only after a move), lane 5: launch budget flag;
                    // ascending global-warp order
                    double x = 0.;
                    if (lane < 6 && (dirty || lane == 5))
#pragma unroll 4
                        for (int g = 0; g < GW; ++g) x += allb[g * 8 + lane];
                    const double fl = __shfl_sync(0xffffffffu, x, 5);
                    if (fl != 0.) slice_over = true;
                    if (dirty) {
                        B.E0 = __shfl_sync(0xffffffffu, x, 0); B.E1 = __shfl_sync(0xffffffffu, x, 1);
                        B.E2 = __shfl_sync(0xffffffffu, x, 2); B.Sm = __shfl_sync(0xffffffffu, x, 3);
                        B.Smm = __shfl_sync(0xffffffffu, x, 4);
                        B.a0 = a0; B.b0 = b0;
                        // the centre is the fit of this state up to the fp32 error of the delta sums:
                        // chi2 at the centre exceeds the minimum by O(1e-14) relative
                        conval = B.E0 * inv_nq;
                    }
                }
                if (FAST_PROF) pd[1] = clock64();
                // lane l owns proposals l, l + 32, ...: the batch is decided 32 proposals at a time
                float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
                const bool mine = lane < Kb;
                const float4 *allp = s_xp + (size_t)xb * GWC * XPC;
                auto prop_sums = [&](int k, float &x1, float &x2, float &x3, float &x4) {
                    x1 = 0.f; x2 = 0.f; x3 = 0.f; x4 = 0.f;
                    if (k < Kb) {
#pragma unroll 4
                        for (int g = 0; g < GW; ++g) {           // ascending global-warp order
                            const float4 v = allp[(size_t)g * XPC + k];
                            x1 += v.x; x2 += v.y; x3 += v.z; x4 += v.w;
                        }
                    }
                };
                prop_sums(lane, s1, s2, s3, s4);
                if (FAST_PROF) pd[2] = clock64();
                int ks = -1, accmask = 0, gused = Kb;
                double ca = a0, cb = b0, cc = conval;
                if (!gram) {
                    double chi = 0., an = 0., bn = 0., e1, e2, rss;
                    unsigned mask = 0;
                    int koff = 0;
                    float c1 = s1, c2 = s2, c3 = s3, c4 = s4;
                    for (int r0 = 0; r0 < Kb; r0 += 32) {        // rounds of 32 proposals, up to the first move
                        float n1, n2, n3, n4;                    // (the next round's sums load under this round's solve)
                        prop_sums(lane + r0 + 32, n1, n2, n3, n4);
                        bool better = false;
                        if (lane + r0 < Kb) {
                            fast_solve(B, a.find_bg, a.pos_bg, c1, c2, c3, c4, rss, an, bn, e1, e2);   // mcsas.py:376
                            chi = rss * inv_nq;
                            better = chi < conval;               // :379 strict
                            s_chi[lane + r0] = chi;
                        }
                        mask = __ballot_sync(0xffffffffu, better);
                        koff = r0;
                        if (mask) break;
                        c1 = n1; c2 = n2; c3 = n3; c4 = n4;
                    }
                    const int kl = mask ? (__ffs(mask) - 1) : -1;
                    ks = kl >= 0 ? kl + koff : -1;
                    if (ks >= 0) {
                        ca = __shfl_sync(0xffffffffu, an, kl); cb = __shfl_sync(0xffffffffu, bn, kl);
                        cc = __shfl_sync(0xffffffffu, chi, kl);
                    }
                } else {
                    // Gram totals over the chain's warps (ascending global-warp order)
                    const float *allg = s_xg + (size_t)xb * GWC * 32;
                    for (int i = lane; i < NGP; i += 32) {
                        float v = 0.f;
                        for (int g = 0; g < GW; ++g) v += allg[g * 32 + i];
                        s_gtot[i] = v;
                    }
                    __syncwarp();
                    int pos = 0;
                    double g1 = s1, g2 = s2, g3 = s3, g4 = s4;      // fp64 copies updated per accepted move
                    for (int it = 0; it <= FGRP; ++it) {
                        double rss = 0., an = 0., bn = 0., e1 = 0., e2 = 0.;
                        bool better = false;
                        if (mine && lane >= pos) {
                            fast_solve(B, a.find_bg, a.pos_bg, g1, g2, g3, g4, rss, an, bn, e1, e2);
                            const double chi = rss * inv_nq;
                            better = chi < cc;                   // :379 against the state reached so far
                            s_chi[lane] = chi;
                        }
                        const unsigned mask = __ballot_sync(0xffffffffu, better);
                        if (!mask) break;
                        const int i = __ffs(mask) - 1;
                        const double ai = __shfl_sync(0xffffffffu, an, i), bi = __shfl_sync(0xffffffffu, bn, i);
                        const double ri = __shfl_sync(0xffffffffu, rss, i);
                        const double e1i = __shfl_sync(0xffffffffu, e1, i), e2i = __shfl_sync(0xffffffffu, e2, i);
                        const double s2i = __shfl_sync(0xffffffffu, g2, i), s3i = __shfl_sync(0xffffffffu, g3, i);
                        const double s4i = __shfl_sync(0xffffffffu, g4, i);
                        const double da = ai - B.a0, db = bi - B.b0;
                        if (mine && lane > i) {                  // sums of later proposals w.r.t. the new state
                            const double Gp = (double)s_gtot[tri_idx_rt(FGRP, i, lane)];
                            g1 = g1 - da * g3 - ai * Gp - db * g4;
                            g3 += Gp;
                        }
                        B.Sm += s4i; B.Smm += 2. * s3i + s2i;
                        B.E0 = ri; B.E1 = e1i; B.E2 = e2i; B.a0 = ai; B.b0 = bi;
                        ca = ai; cb = bi; cc = ri * inv_nq;
                        accmask |= 1 << i;
                        pos = i + 1;
                        // the reference tests the criterion before EVERY step (mcsas.py:354-357):
                        // the steps behind the move that reaches it are not made
                        if (cc <= a.conv_crit) { gused = i + 1; break; }
                    }
                    ks = accmask ? 0 : -1;
                }
                if (FAST_PROF) pd[3] = clock64();
                if (lane == 0) {
                    *s_B = B;
                    s_dec->kstar = ks;
                    s_dec->accmask = accmask;
                    s_dec->used = gused;
                    s_dec->wd = slice_over ? 1 : 0;
                    s_dec->a = ca;                               // centre of the state after this batch
                    s_dec->b = cb;
                    s_dec->conval = cc;
                }
            }
            if (W == 1)                                          // one-warp CTA: fill after deciding
                for (long long t = fill_from + lane; t < fill_to; t += 32) fill_entry(t, attempt);
            xphase ^= 1u << xb;
            xb_ = xb ^ 1;
            const long long pt4 = FAST_PROF ? clock64() : 0;
            named_bar_sync(1, T);                                // B2: the decision is published
            const long long pt5 = FAST_PROF ? clock64() : 0;
            const int kstar = s_dec->kstar;
            const int accmask = s_dec->accmask;
            const int used = gram ? s_dec->used : (kstar >= 0 ? kstar + 1 : Kb);
            conval = s_dec->conval;
            slice_over = s_dec->wd != 0;
            if (a.trace_accept && rank == 0) {
                for (int k = tid; k < used; k += T) {
                    a.trace_accept[(long long)chain * a.max_iter + step + k] = gram ? ((accmask >> k) & 1) : (k == kstar);
                    a.trace_chi[(long long)chain * a.max_iter + step + k] = s_chi[k];
                }
            }
            dirty = false;
            if (gram) {
                if (accmask) {                                   // mcsas.py:381-384 for every accepted step
                    a0 = s_dec->a; b0 = s_dec->b;
#pragma unroll
                    for (int j = 0; j < QPT; ++j) {
                        const int ql = qtid * QPT + j;
                        const double sq = s_sw[ql];
                        float du = 0.f;
#pragma unroll
                        for (int s = 0; s < FGRP; ++s)
                            if ((accmask >> (s ^ perm)) & 1) du += uu[s][j];
                        // uu = (new - old)/sigma in the model's units: undo the 1/sigma
                        const double f = s_ft[cur * FAST_QCAP + ql] + (om[j] > 0.f ? (double)__fdividef(du, om[j]) : 0.);
                        if (grp == PS - 1) s_ft[(cur ^ 1) * FAST_QCAP + ql] = f;   // the group that re-reads it (refresh)
                        rho[j] = (float)(sq * (s_In[ql] - (a0 * f + b0)));
                        mu[j] = (float)(sq * f);
                    }
                    cur ^= 1;
                    if (CACHE && grp == 0) {                     // rows of the accepted contributions
#pragma unroll
                        for (int s = 0; s < FGRP; ++s) {
                            if (!((accmask >> (s ^ perm)) & 1)) continue;
                            int ri = ri0 + (s ^ perm);
                            if (ri >= N) ri -= N;
                            row_store<QPT>(crow + (size_t)ri * RL, dnk[CACHE ? s : 0], gpol);
                        }
                    }
                    if (tid < FGRP * P) {
                        const int k = tid / P, p = tid - k * P;
                        if ((accmask >> k) & 1) {
                            const int e = ((int)step + k) & (RING - 1);
                            const int rix = (int)((uint32_t)(step + k) % (uint32_t)N);
                            s_rset[rix * P + p] = s_ring[e].prop[p];
                        }
                    }
                    if (tid == 0) s_cold->moves += __popc(accmask);
                    dirty = true;
                    last_acc = step + (31 - __clz(accmask));
                }
                if (__popc(accmask) < a.gram_out) {              // too few moves in a whole chunk: speculate
                    gmode = false;
                    K = min(kmax, 2 * FGRP);
                }
            } else if (kstar >= 0) {                             // mcsas.py:381-384
                a0 = s_dec->a; b0 = s_dec->b;
                const int e = ((int)step + kstar) & (RING - 1);
                const Contrib<float> cn = s_ring[e].cn, co = s_ring[e].co;
                const float *nn = s_nodes + (size_t)(e * 2) * NSTR, *no = nn + NSTR;
                const int rix = (int)((uint32_t)(step + kstar) % (uint32_t)N);
                float dd[QPT];
                if (CACHE) {
                    float dv[1][QPT], ov[1][QPT], d1[1][QPT];
                    row_load<QPT>(crow + (size_t)rix * RL, ov[0], row_policy(rix < keep_rows));   // in flight during the evaluation
                    const int ent[1] = {e};
                    FO::template diffChunk<QPT, 1>(s_ring, ent, s_nodes, NSTR, qt, Q, ov, dv, d1);
#pragma unroll
                    for (int j = 0; j < QPT; ++j) dd[j] = d1[0][j];
                    if (PS > 1) named_bar_sync(1, T);            // every group has read the old row
                    if (grp == 0) row_store<QPT>(crow + (size_t)rix * RL, dv[0], row_policy(rix < keep_rows));
                } else {
                    FO::template diffS<QPT>(cn, co, nn, no, qt, Q, dd);
                }
#pragma unroll
                for (int j = 0; j < QPT; ++j) {
                    const int ql = qtid * QPT + j;
                    const double f = s_ft[cur * FAST_QCAP + ql] + (double)dd[j] * (double)s_qs[ql];
                    const double sq = s_sw[ql];
                    if (grp == PS - 1) s_ft[(cur ^ 1) * FAST_QCAP + ql] = f; // other groups still read `cur`
                    rho[j] = (float)(sq * (s_In[ql] - (a0 * f + b0)));
                    mu[j] = (float)(sq * f);
                }
                cur ^= 1;
                if (tid < P) s_rset[rix * P + tid] = s_ring[e].prop[tid];
                if (tid == 0) { s_cold->moves += 1; s_cold->evals += 2; }
                dirty = true;
                gmode = step + kstar - last_acc <= a.gram_in;
                last_acc = step + kstar;
                K = max(1, min(kmax, 1 << (31 - __clz(kstar + 1))));
            } else {
                K = min(kmax, K * 2);
            }
            if (tid == 0) { s_cold->evals += 2 * Kb; s_cold->steps_all += used; }
            step += used;                                        // :403-404 (ri = step % N)
            if (FAST_PROF && tid == 0 && rank == 0) {               // phase cycles of warp 0 (tools/prof)
                long long *pp = a.prof + ((long long)chain * 2 + (gram ? 1 : 0)) * 24;
                pp[16] += pd[3] - pd[2]; pp[17] += pt4 - pd[3];
                pp[13] += pd[0] - pt3; pp[14] += pd[1] - pd[0]; pp[15] += pd[2] - pd[1];
                pp[8] += pt1a - pt1; pp[9] += ptc[0]; pp[10] += ptc[1]; pp[11] += ptc[2]; pp[12] += ptc[3];
                pp[0] += pt1 - pt0; pp[1] += pt2 - pt1; pp[2] += pt3 - pt2; pp[3] += pt4 - pt3;
                pp[4] += pt5 - pt4; pp[5] += clock64() - pt5; pp[6] += 1; pp[7] += used;
            }
            slice_over = slice_over || step >= slice_end || batch >= batch_end;
            live = conval > a.conv_crit && step < a.max_iter && !slice_over;
        }
        __syncthreads();
        // ---- exact fit of the accepted state (mcsas.py:424-425) --------------------------------
        cv_run = conval;                                         // what the accept test compares against
        lsq_all(sc_a, sc_b, conval);
        if (status == -2) break;
        // the running chi-squared crossed the criterion on fp32 delta sums but the exact fit has
        // not: the chain goes on from the exact value (whether or not a launch ends here)
        if (N > 1 && cv_run <= a.conv_crit && conval > a.conv_crit && step < a.max_iter) {
            cv_run = conval;
            dirty = true;
            if (!slice_over) { live = true; continue; }
        }
        break;
        }
        if (status == -2) break;
        const bool running = (N > 1 && conval > a.conv_crit && step < a.max_iter);
        if (running) { status = 0; break; }                      // this launch's budget ended: resume later
        status = 1;
        if (conval <= a.conv_crit || a.stream || attempt + 1 >= a.max_attempts || N <= 1)
            break;
        if (a.slice_steps > 0 || a.slice_ns > 0) {               // retry (mcsas.py:220-246) in a later launch
            status = 2; ++attempt; step = 0;
            if (tid == 0) s_cold->moves = 0;
            break;
        }
        __syncthreads();
    }

    // ---- results / saved state ---------------------------------------------------------------
    if (rank == 0)
        for (int i = tid; i < N * P; i += blockDim.x) a.rset[(long long)chain * N * P + i] = s_rset[i];
    if (grp == 0) {
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            int i = (rank * TQ + qtid) * QPT + j;
            if (i < nq) {
                const double f = s_ft[cur * FAST_QCAP + qtid * QPT + j];
                a.fit_meas[(long long)chain * a.nq_stride + i] = (f * sc_a + sc_b) * ds.i_scale;         // mcsas.py:429
                if (status == 0) a.ft_save[(long long)chain * a.nq_stride + i] = f;
            }
        }
    }
    if (tid == 0 && rank == 0) {
        // internal model = (reference model) * G / i_scale-normalised data: undo both
        a.scaling[chain] = sc_a * FAST_G * ds.i_scale;
        a.background[chain] = sc_b * ds.i_scale;
        a.conval[chain] = conval;
        a.conval0[chain] = s_cold->conval0;
        a.cen_a[chain] = a0; a.cen_b[chain] = b0; a.cv_run[chain] = cv_run;
        a.batch_k[chain] = K | (gmode ? 1 << 16 : 0);
        a.last_acc[chain] = (double)last_acc;
        a.num_iter[chain] = step;
        a.num_moves[chain] = s_cold->moves;
        a.evals[chain] = s_cold->evals;
        a.steps_all[chain] = s_cold->steps_all;
        a.attempts[chain] = attempt + 1;
        a.elapsed_s[chain] = (a.resume ? a.elapsed_s[chain] : 0.) + 1e-9 * (double)(globaltimer_ns() - s_cold->t_start);
        a.status[chain] = status;
        if (!is_backup && a.needed_left) atomicSub(a.needed_left, 1);
        if (a.repack_ctas > 0) atomicAdd(a.needed_left + 1, C);
    }
    if (CACHE) {
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicExch(a.cache_lock + region, 0); }
    }
    if (C > 1) cluster_sync_all();
}
