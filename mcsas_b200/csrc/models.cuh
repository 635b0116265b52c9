// models.cuh -- device form factors of the McSAS model plug-ins.
//
// Every model is split into
//   prep():  once per contribution (one thread): volume, weight w = V^(2c), per-node
//            geometry of the orientation average -> Contrib<Real> (+ node arrays in smem)
//   eval():  once per (contribution, q point): intensity  it = F(q)^2 * w
// Real = double: parity mode, SI units, reference operation order.
// Real = float : production mode, lengths in nm / q in 1/nm, sin/cos/rcp on the MUFU
//                pipe with an exact-FMA phase reduction; amplitudes are scaled by a
//                per-run constant G (kept in fp64) so fp32 never leaves its range.
#pragma once
#include "common.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

template <class Real> struct QRep;
template <> struct QRep<double> {
    double q;
    __device__ void set(double q_si) { q = q_si; }
    __device__ double q_si() const { return q; }
};
template <> struct QRep<float> {
    float q;        // 1/nm
    float th, tl;   // q/(2 pi) in turns per nm, split hi + lo
    __device__ void set(double q_si) {
        double qn = q_si * 1e-9;
        q = (float)qn;
        double t = qn * 0.15915494309189535;
        th = (float)t;
        tl = (float)(t - (double)th);
    }
    __device__ double q_si() const { return (double)q * 1e9; }
};

template <class Real> struct Units;
template <> struct Units<double> { static constexpr double LEN = 1.0; };   // metres
template <> struct Units<float> { static constexpr double LEN = 1e9; };    // nanometres

// x = q*r and sin/cos(x)
__device__ __forceinline__ void sincos_qr(const QRep<double> &Q, double r, double &s, double &c,
                                          double &x)
{
    x = Q.q * r;
    sincos(x, &s, &c);
}

__device__ __forceinline__ void sincos_qr(const QRep<float> &Q, float r, float &s, float &c,
                                          float &x)
{
    // phase in turns: n = rint(th*r); frac = fma(th, r, -n) is exact up to one rounding, the
    // low part of q/(2 pi) restores the bits a float q loses, so q*r ~ 1e3..1e4 rad keeps
    // ~1e-7 turn accuracy; then MUFU.SIN / MUFU.COS on [-pi, pi].
    x = Q.q * r;
    float t = Q.th * r;
    float n = (t + 12582912.f) - 12582912.f;           // rint for |t| < 2^22
    float fr = fmaf(Q.th, r, -n);
    fr = fmaf(Q.tl, r, fr);
    float a = fr * 6.2831853071795865f;
    s = __sinf(a);
    c = __cosf(a);
}

__device__ __forceinline__ float fast_rcp(float x)      // MUFU.RCP
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float fast_rsqrt(float x)    // MUFU.RSQ
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 3 (sin x - x cos x) / x^3     (sphere.py:62)
__device__ __forceinline__ double sph_ff(const QRep<double> &Q, double r)
{
    double s, c, x;
    sincos_qr(Q, r, s, c, x);
    return 3. * (s - x * c) / (x * x * x);
}

// float flavour returns F/3 (the factor 9 of F^2 is folded into the amplitude, SPH_NORM)
__device__ __forceinline__ float sph_ff(const QRep<float> &Q, float r)
{
    float s, c, x;
    sincos_qr(Q, r, s, c, x);
    float x2 = x * x;
    float big = fmaf(-x, c, s) * fast_rcp(x2 * x);
    // series below x = 1: MUFU's absolute error would be amplified by 1/x^3
    float ser = 0.33333333f + x2 * (-0.033333333f + x2 * (1.1904762e-3f + x2 * (-2.2045855e-5f
                + x2 * 2.5052108e-7f)));
    return x < 1.f ? ser : big;
}
// sph_ff(float) for the M q points of a thread, stage by stage (M-way instruction-level
// parallelism; same arithmetic as the scalar form)
// (iq3[i] = 1 / q_i^3: with it 1 / x^3 is a per-q factor times ONE reciprocal per radius instead of one per
// (q, radius) pair -- the orientation-averaged loops are bound by the XU pipe: 3 -> 2.25 MUFU per pair)
template <int M> __device__ __forceinline__ void sph_ff_vec(const QRep<float> (&Q)[M], const float (&iq3)[M], float r,
                                                            float (&f)[M])
{
    float x[M], t[M], fr[M], sn[M], cs[M], x2[M], big[M], ser[M];
    const float ir3 = fast_rcp(r * r * r);
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = Q[i].q * r;
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = Q[i].th * r;
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = (t[i] + 12582912.f) - 12582912.f;
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(Q[i].th, r, -t[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(Q[i].tl, r, fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fr[i] * 6.2831853071795865f;
#pragma unroll
    for (int i = 0; i < M; ++i) sn[i] = __sinf(fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) cs[i] = __cosf(fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) x2[i] = x[i] * x[i];
#pragma unroll
    for (int i = 0; i < M; ++i) big[i] = iq3[i] * ir3;
#pragma unroll
    for (int i = 0; i < M; ++i) ser[i] = fmaf(x2[i], 2.5052108e-7f, -2.2045855e-5f);
#pragma unroll
    for (int i = 0; i < M; ++i) big[i] = fmaf(-x[i], cs[i], sn[i]) * big[i];
#pragma unroll
    for (int i = 0; i < M; ++i) ser[i] = fmaf(x2[i], ser[i], 1.1904762e-3f);
#pragma unroll
    for (int i = 0; i < M; ++i) ser[i] = fmaf(x2[i], ser[i], -0.033333333f);
#pragma unroll
    for (int i = 0; i < M; ++i) ser[i] = fmaf(x2[i], ser[i], 0.33333333f);
#pragma unroll
    for (int i = 0; i < M; ++i) f[i] = x[i] < 1.f ? ser[i] : big[i];
}
template <class Real> struct SphNorm;
template <> struct SphNorm<double> { static constexpr double V = 1.0; };
template <> struct SphNorm<float> { static constexpr double V = 9.0; };

// log2 of a positive double: exponent by bit ops, mantissa on MUFU.LG2 (abs err ~2e-7)
__device__ __forceinline__ double fast_log2_pos(double v)
{
    int e;
    double m = frexp(v, &e);
    return (double)e + (double)__log2f((float)m);
}
// g * v^y for v, g > 0 with ~3e-7 relative error and no double pow: used for the fp32
// chain's contribution amplitudes w*G = volume^(2c)*G (sasmodel.py:44)
__device__ __forceinline__ float fast_amp(double v, double y, double g)
{
    double t = y * fast_log2_pos(v) + fast_log2_pos(g);
    double ti = rint(t);
    return ldexpf(exp2f((float)(t - ti)), (int)ti);
}

// Parameter vector view: fixed model parameters with the active values patched in
// (scatteringmodel.py:94-96  p.setValue(v) for each active parameter) -- no local array.
struct PVec {
    const ModelParams &mp;
    double v0, v1, v2, v3;
    __device__ __forceinline__ PVec(const ModelParams &m, const double *pv) : mp(m), v0(0.), v1(0.), v2(0.), v3(0.)
    {
        if (m.n_active > 0) v0 = pv[0];
        if (m.n_active > 1) v1 = pv[1];
        if (m.n_active > 2) v2 = pv[2];
        if (m.n_active > 3) v3 = pv[3];
    }
    __device__ __forceinline__ double operator[](int i) const
    {
        double v = mp.params[i];
        if (mp.n_active > 0 && mp.active_idx[0] == i) v = v0;
        if (mp.n_active > 1 && mp.active_idx[1] == i) v = v1;
        if (mp.n_active > 2 && mp.active_idx[2] == i) v = v2;
        if (mp.n_active > 3 && mp.active_idx[3] == i) v = v3;
        return v;
    }
};

template <class Real> struct Contrib {
    Real amp;        // w (fp64) or w*G*(model scale)^2 (fp32)
    Real c[7];
};

// Volume / weight / surface triple returned by prep (fp64, SI).
struct VWS { double v, w, s, amp_raw; };

// Amplitude of a contribution, amp = w * fac * G * norm with w = vol^(2c) (sasmodel.py:44):
// exact double pow (parity mode, eval and fraction kernels) or, FAST, the MUFU-based
// path of the fp32 chain (w and amp_raw are then not produced).
template <bool FAST, class Real>
__device__ __forceinline__ void set_amp(VWS &o, Contrib<Real> &cc, double vol, double ce, double fac,
                                        double G, double norm)
{
    if (FAST) {
        cc.amp = (Real)fast_amp(vol, 2. * ce, fac * G * norm);
    } else {
        o.w = pow(vol, 2. * ce);
        o.amp_raw = o.w * fac;
        cc.amp = (Real)(o.amp_raw * G * norm);
    }
}

template <class Real> struct QuadTab {
    int D;
    int rule;
    const Real *x;
    const Real *wt;
    const SmearTab *sm;      // smearing profile (n = 0: none)
    int nsm;                 // profile points in use: sm->n, or a compile-time 0 in the kernel
                             // instantiations for unsmeared data (their hot loops carry no smear code)
    const float4 *smloc;     // fp32 sphere kernel: this thread's column of the smeared-location table
    int smstride;            // (ChainArgs::smloc), entries between two profile points; NULL: none
    int smmono;              // 1: no tabulated location lies below its q point (slit profile / offsets >= 0)
};

template <int MODEL, class Real> struct Ops;

// ------------------------------------------------------------------ Sphere --
template <class Real> struct Ops<MCSAS_SPHERE, Real> {
    static constexpr int NARR = 0;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &,
                               Contrib<Real> &cc, Real *)
    {
        double r = p[0], sld = p[1];
        VWS o;
        double vol = (M_PI * 4. / 3.) * (r * r * r);    // sphere.py:44
        o.v = vol * sld * sld;                          // :53
        o.s = 4. * M_PI * r * r;                        // :37
        set_amp<FAST>(o, cc, vol, ce, 1., G, SphNorm<Real>::V);
        cc.c[0] = (Real)(r * Units<Real>::LEN);
        return o;
    }
    __device__ __forceinline__ static Real eval(const Contrib<Real> &cc, const Real *,
                                                const QuadTab<Real> &, const QRep<Real> &Q)
    {
        Real F = sph_ff(Q, cc.c[0]);
        return (F * F) * cc.amp;                        // sasmodel.py:77
    }
};

// --------------------------------------------------------- LMADenseSphere --
// The structure factor has x^-5 cancellations (lmadensesphere.py:76-86): it is kept
// in fp64 in both precision modes.
template <class Real> struct Ops<MCSAS_LMA_DENSE_SPHERE, Real> {
    static constexpr int NARR = 0;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &,
                               Contrib<Real> &cc, Real *)
    {
        double r = p[0], mu = p[1], mf = p[2], sld = p[3];
        if (mf == -1.) mf = FAST ? cbrt(0.634 / mu) : pow(0.634 / mu, 1. / 3);    // :73-74
        VWS o;
        double vol = (M_PI * 4. / 3.) * (r * r * r);
        o.v = vol * sld * sld;
        o.s = 0.;
        set_amp<FAST>(o, cc, vol, ce, 1., G, SphNorm<Real>::V);
        cc.c[0] = (Real)(r * Units<Real>::LEN);
        cc.c[1] = (Real)(mf * r * Units<Real>::LEN);
        cc.c[2] = (Real)mu;
        return o;
    }
    __device__ static double SFG(double A, double mu)
    {
        double om = 1. - mu, om4 = om * om * om * om;
        double alpha = (1. + 2. * mu) * (1. + 2. * mu) / om4;
        double beta = -6. * mu * (1. + mu / 2.) * (1. + mu / 2.) / om4;
        double gamma = mu * alpha / 2.;
        double sA, cA;
        sincos(A, &sA, &cA);
        double A2 = A * A, A3 = A2 * A, A4 = A2 * A2, A5 = A4 * A;
        return alpha * (sA - A * cA) / A2
             + beta * (2. * A * sA + (2. - A2) * cA - 2.) / A3
             + gamma * (-1. * A4 * cA + 4. * ((3. * A2 - 6.) * cA + (A3 - 6. * A) * sA + 6.)) / A5;
    }
    __device__ static Real eval(const Contrib<Real> &cc, const Real *, const QuadTab<Real> &,
                                const QRep<Real> &Q)
    {
        Real F = sph_ff(Q, cc.c[0]);
        double A = 2. * (double)Q.q * (double)cc.c[1];                    // :91
        double S = 1. / (1. + 24. * (double)cc.c[2] * SFG(A, (double)cc.c[2]) / A);   // :93
        return (Real)((double)(F * F) * S) * cc.amp;                      // :101 squared
    }
};

// ------------------------------------------------------ SphericalCoreShell --
template <class Real> struct Ops<MCSAS_SPHERICAL_CORE_SHELL, Real> {
    static constexpr int NARR = 0;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &,
                               Contrib<Real> &cc, Real *)
    {
        double r = p[0], t = p[1], ec = p[2], es = p[3], esol = p[4];
        double vc = 4. / 3 * M_PI * (r * r * r), vt = 4. / 3 * M_PI * ((r + t) * (r + t) * (r + t));
        double vr = vc / vt;                            // :61-63
        double d1 = es - esol, d2 = vr * (es - ec);     // ks, vRatio*kc prefactors :64-69
        VWS o;
        o.v = vt; o.s = 0.;                             // :72-76
        double nrm = 1.;
        if (sizeof(Real) == 4) {
            nrm = fmax(fabs(d1), fabs(d2));
            if (nrm == 0.) nrm = 1.;
        }
        set_amp<FAST>(o, cc, vt, ce, nrm * nrm, G, SphNorm<Real>::V);
        cc.c[0] = (Real)((r + t) * Units<Real>::LEN);
        cc.c[1] = (Real)(r * Units<Real>::LEN);
        cc.c[2] = (Real)(d1 / nrm);
        cc.c[3] = (Real)(d2 / nrm);
        return o;
    }
    __device__ __forceinline__ static Real eval(const Contrib<Real> &cc, const Real *,
                                                const QuadTab<Real> &, const QRep<Real> &Q)
    {
        Real F = cc.c[2] * sph_ff(Q, cc.c[0]) - cc.c[3] * sph_ff(Q, cc.c[1]);
        return (F * F) * cc.amp;
    }
};

// ---------------------------------------------------- EllipsoidalCoreShell --
template <class Real> struct Ops<MCSAS_ELLIPSOIDAL_CORE_SHELL, Real> {
    static constexpr int NARR = 2;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &qt,
                               Contrib<Real> &cc, Real *nodes)
    {
        double a = p[0], b = p[1], t = p[2], ec = p[3], es = p[4], esol = p[5];
        double vc = 4. / 3. * M_PI * a * (b * b);
        double vt = 4. / 3. * M_PI * (a + t) * ((b + t) * (b + t));
        double k1 = (ec - es) * (vc / vt), k2 = (es - esol) * 1.;   // :83-88
        VWS o;
        o.v = vt; o.s = 0.;                             // :93-97
        double nrm = 1.;
        if (sizeof(Real) == 4) {
            nrm = fmax(fabs(k1), fabs(k2));
            if (nrm == 0.) nrm = 1.;
        }
        set_amp<FAST>(o, cc, vt, ce, nrm * nrm, G, SphNorm<Real>::V);
        cc.c[0] = (Real)(k1 / nrm);
        cc.c[1] = (Real)(k2 / nrm);
        for (int j = 0; j < qt.D; ++j) {
            double mu = (double)qt.x[j];
            nodes[j] = (Real)(sqrt(a * a * mu * mu + b * b * (1. - mu * mu)) * Units<Real>::LEN);
            nodes[qt.D + j] = (Real)(sqrt((a + t) * (a + t) * mu * mu
                                          + (b + t) * (b + t) * (1. - mu * mu)) * Units<Real>::LEN);
        }
        return o;
    }
    __device__ static Real eval(const Contrib<Real> &cc, const Real *nodes,
                                const QuadTab<Real> &qt, const QRep<Real> &Q)
    {
        Real acc = 0;
#pragma unroll 2
        for (int j = 0; j < qt.D; ++j) {
            Real f = cc.c[0] * sph_ff(Q, nodes[j]) + cc.c[1] * sph_ff(Q, nodes[qt.D + j]);
            acc += qt.wt[j] * (f * f);
        }
        return acc * cc.amp;                            // sqrt(mean)^2 * w  :90
    }
    // fp32 production path: the thread's M q points together through the node loop
    template <int M>
    __device__ __forceinline__ static void evalV(const Contrib<float> &cc, const float *nodes,
                                                 const QuadTab<float> &qt, const QRep<float> (&Q)[M],
                                                 float (&out)[M])
    {
        float acc[M], iq3[M];
#pragma unroll
        for (int i = 0; i < M; ++i) { acc[i] = 0.f; iq3[i] = fast_rcp(Q[i].q * Q[i].q * Q[i].q); }
        for (int j = 0; j < qt.D; ++j) {
            const float wj = qt.wt[j];
            float f1[M], f2[M];
            sph_ff_vec<M>(Q, iq3, nodes[j], f1);
            sph_ff_vec<M>(Q, iq3, nodes[qt.D + j], f2);
#pragma unroll
            for (int i = 0; i < M; ++i) f1[i] = cc.c[0] * f1[i] + cc.c[1] * f2[i];
#pragma unroll
            for (int i = 0; i < M; ++i) acc[i] += wj * (f1[i] * f1[i]);
        }
#pragma unroll
        for (int i = 0; i < M; ++i) out[i] = acc[i] * cc.amp;
    }
};

// ----------------------------------------------------- EllipsoidsIsotropic --
template <class Real> struct Ops<MCSAS_ELLIPSOIDS_ISOTROPIC, Real> {
    static constexpr int NARR = 1;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &qt,
                               Contrib<Real> &cc, Real *nodes)
    {
        double Ra = p[0], Rc = (p[1] != 0.) ? p[0] * p[3] : p[2], sld = p[5];   // :60-63
        VWS o;
        double vol = 4. / 3. * M_PI * (Ra * Ra) * Rc;   // :80
        o.v = vol * sld * sld; o.s = 0.;                // :84
        set_amp<FAST>(o, cc, vol, ce, 1., G, SphNorm<Real>::V);
        for (int j = 0; j < qt.D; ++j) {
            double sa, ca;
            sincos((double)qt.x[j], &sa, &ca);
            nodes[j] = (Real)(sqrt(Ra * Ra * (sa * sa) + Rc * Rc * (ca * ca)) * Units<Real>::LEN);  // :56
        }
        return o;
    }
    __device__ static Real eval(const Contrib<Real> &cc, const Real *nodes,
                                const QuadTab<Real> &qt, const QRep<Real> &Q)
    {
        Real acc = 0;
#pragma unroll 4
        for (int j = 0; j < qt.D; ++j) {
            Real f = sph_ff(Q, nodes[j]);               // :67-68
            acc += qt.wt[j] * (f * f);                  // :71 (weights carry sin(alpha)/D)
        }
        return acc * cc.amp;
    }
    // fp32 production path: the thread's M q points together through the node loop
    template <int M>
    __device__ __forceinline__ static void evalV(const Contrib<float> &cc, const float *nodes,
                                                 const QuadTab<float> &qt, const QRep<float> (&Q)[M],
                                                 float (&out)[M])
    {
        float acc[M], iq3[M];
#pragma unroll
        for (int i = 0; i < M; ++i) { acc[i] = 0.f; iq3[i] = fast_rcp(Q[i].q * Q[i].q * Q[i].q); }
        for (int j = 0; j < qt.D; ++j) {
            const float wj = qt.wt[j];
            float f[M];
            sph_ff_vec<M>(Q, iq3, nodes[j], f);
#pragma unroll
            for (int i = 0; i < M; ++i) acc[i] += wj * (f[i] * f[i]);
        }
#pragma unroll
        for (int i = 0; i < M; ++i) out[i] = acc[i] * cc.amp;
    }
};

// ------------------------------------------------------ CylindersIsotropic --
__device__ __forceinline__ double bessel_j1(double x) { return j1(x); }
// fp32 J1 for x >= 0 without libdevice's slow paths: rational approximation below 8,
// Hankel asymptotic form above with the phase x - 3 pi / 4 reduced in turns and MUFU
// sin/cos (coefficients: Abramowitz & Stegun 9.4.4-9.4.6 as tabulated in Numerical Recipes
// `bessj1`); absolute error ~1e-6, far below the fp32 form-factor budget.
__device__ __forceinline__ float bessel_j1(float x)
{
    const float y = x * x;
    float n1 = fmaf(y, -30.16036606f, 15704.48260f);
    n1 = fmaf(y, n1, -2972611.439f);
    n1 = fmaf(y, n1, 242396853.1f);
    n1 = fmaf(y, n1, -7895059235.f);
    n1 = fmaf(y, n1, 72362614232.f);
    float d1 = y + 376.9991397f;
    d1 = fmaf(y, d1, 99447.43394f);
    d1 = fmaf(y, d1, 18583304.74f);
    d1 = fmaf(y, d1, 2300535178.f);
    d1 = fmaf(y, d1, 144725228442.f);
    float rd, small, z;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rd) : "f"(d1));
    small = x * n1 * rd;
    float rs;                                                    // 1/sqrt(x): gives 1/x and sqrt(2/(pi x)) without a second MUFU
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(x, 1e-20f)));
    z = rs * rs;
    const float z8 = 8.f * z, w = z8 * z8;
    float p = fmaf(w, -0.240337019e-6f, 0.2457520174e-5f);
    p = fmaf(w, p, -0.3516396496e-4f);
    p = fmaf(w, p, 0.183105e-2f);
    p = fmaf(w, p, 1.f);
    float qv = fmaf(w, 0.105787412e-6f, -0.88228987e-6f);
    qv = fmaf(w, qv, 0.8449199096e-5f);
    qv = fmaf(w, qv, -0.2002690873e-3f);
    qv = fmaf(w, qv, 0.04687499995f);
    float t = fmaf(x, 0.15915494309189535f, -0.375f);           // (x - 3 pi/4) in turns
    const float nn = (t + 12582912.f) - 12582912.f;
    t = (t - nn) * 6.2831853071795865f;
    const float big = (0.7978845608f * rs) * (__cosf(t) * p - z8 * __sinf(t) * qv);
    return x < 8.f ? small : big;
}

__device__ __forceinline__ double recip(double x) { return 1. / x; }
__device__ __forceinline__ float recip(float x) { return fast_rcp(x); }

// M independent J1 evaluations written stage by stage (same arithmetic as bessel_j1(float)):
// ptxas keeps program order, so this form gives M-way instruction-level parallelism where M
// scalar calls in a row are scheduled as M serial dependency chains.
template <int M> __device__ __forceinline__ void bessel_j1_vec(const float (&x)[M], float (&out)[M])
{
    float y[M], n1[M], d1[M], z[M], w[M], p[M], qv[M], t[M], nn[M], sn[M], cs[M], rt[M];
#pragma unroll
    for (int i = 0; i < M; ++i) y[i] = x[i] * x[i];
#pragma unroll
    for (int i = 0; i < M; ++i) n1[i] = fmaf(y[i], -30.16036606f, 15704.48260f);
#pragma unroll
    for (int i = 0; i < M; ++i) d1[i] = y[i] + 376.9991397f;
#pragma unroll
    for (int i = 0; i < M; ++i) rt[i] = fast_rsqrt(fmaxf(x[i], 1e-20f));   // 1/sqrt(x): 1/x and sqrt(2/(pi x)) from ONE MUFU
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = fmaf(x[i], 0.15915494309189535f, -0.375f);   // (x - 3 pi/4) in turns
#pragma unroll
    for (int i = 0; i < M; ++i) n1[i] = fmaf(y[i], n1[i], -2972611.439f);
#pragma unroll
    for (int i = 0; i < M; ++i) d1[i] = fmaf(y[i], d1[i], 99447.43394f);
#pragma unroll
    for (int i = 0; i < M; ++i) nn[i] = (t[i] + 12582912.f) - 12582912.f;
#pragma unroll
    for (int i = 0; i < M; ++i) n1[i] = fmaf(y[i], n1[i], 242396853.1f);
#pragma unroll
    for (int i = 0; i < M; ++i) d1[i] = fmaf(y[i], d1[i], 18583304.74f);
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = (t[i] - nn[i]) * 6.2831853071795865f;
#pragma unroll
    for (int i = 0; i < M; ++i) n1[i] = fmaf(y[i], n1[i], -7895059235.f);
#pragma unroll
    for (int i = 0; i < M; ++i) d1[i] = fmaf(y[i], d1[i], 2300535178.f);
#pragma unroll
    for (int i = 0; i < M; ++i) cs[i] = __cosf(t[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) sn[i] = __sinf(t[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) n1[i] = fmaf(y[i], n1[i], 72362614232.f);
#pragma unroll
    for (int i = 0; i < M; ++i) d1[i] = fmaf(y[i], d1[i], 144725228442.f);
#pragma unroll
    for (int i = 0; i < M; ++i) { z[i] = 8.f * (rt[i] * rt[i]); w[i] = z[i] * z[i]; }  // z8, w
#pragma unroll
    for (int i = 0; i < M; ++i) d1[i] = fast_rcp(d1[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fmaf(w[i], -0.240337019e-6f, 0.2457520174e-5f);
#pragma unroll
    for (int i = 0; i < M; ++i) qv[i] = fmaf(w[i], 0.105787412e-6f, -0.88228987e-6f);
#pragma unroll
    for (int i = 0; i < M; ++i) rt[i] = 0.7978845608f * rt[i];
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fmaf(w[i], p[i], -0.3516396496e-4f);
#pragma unroll
    for (int i = 0; i < M; ++i) qv[i] = fmaf(w[i], qv[i], 0.8449199096e-5f);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fmaf(w[i], p[i], 0.183105e-2f);
#pragma unroll
    for (int i = 0; i < M; ++i) qv[i] = fmaf(w[i], qv[i], -0.2002690873e-3f);
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = fmaf(w[i], p[i], 1.f);
#pragma unroll
    for (int i = 0; i < M; ++i) qv[i] = fmaf(w[i], qv[i], 0.04687499995f);
#pragma unroll
    for (int i = 0; i < M; ++i) n1[i] = x[i] * n1[i] * d1[i];                           // small-argument form
#pragma unroll
    for (int i = 0; i < M; ++i) p[i] = rt[i] * (cs[i] * p[i] - z[i] * sn[i] * qv[i]);   // Hankel form
#pragma unroll
    for (int i = 0; i < M; ++i) out[i] = x[i] < 8.f ? n1[i] : p[i];
}

// x_i = q_i r and sin(x_i) for M q points (the phase split of sincos_qr), stage by stage
template <int M> __device__ __forceinline__ void sin_qr_vec(const QRep<float> (&Q)[M], float r, float (&s)[M],
                                                            float (&x)[M])
{
    float t[M], fr[M];
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = Q[i].q * r;
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = Q[i].th * r;
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = (t[i] + 12582912.f) - 12582912.f;
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(Q[i].th, r, -t[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(Q[i].tl, r, fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = __sinf(fr[i] * 6.2831853071795865f);
}

template <class Real> struct Ops<MCSAS_CYLINDERS_ISOTROPIC, Real> {
    static constexpr int NARR = 2;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &qt,
                               Contrib<Real> &cc, Real *nodes)
    {
        double r = p[0], hL = (p[1] != 0.) ? p[0] * p[3] : 0.5 * p[2], sld = p[5];  // :65-68
        VWS o;
        double vol = M_PI * (r * r) * (hL * 2.);        // :93-98
        o.v = vol * sld * sld; o.s = 0.;                // :101
        set_amp<FAST>(o, cc, vol, ce, 1., G, 1.);
        cc.c[0] = (Real)(r * Units<Real>::LEN);
        cc.c[1] = (Real)(hL * Units<Real>::LEN);
        for (int j = 0; j < qt.D; ++j) {
            double x = (double)qt.x[j];
            nodes[j] = (Real)(r * sqrt(1. - x * x) * Units<Real>::LEN);    // :70
            nodes[qt.D + j] = (Real)(hL * x * Units<Real>::LEN);          // :71 (B/2 = q hL x)
        }
        return o;
    }
    __device__ static Real eval(const Contrib<Real> &cc, const Real *nodes,
                                const QuadTab<Real> &qt, const QRep<Real> &Q)
    {
        Real acc = 0;
        int j0 = 0, j1e = qt.D;
        if (qt.rule == MCSAS_QUAD_REFERENCE) {
            // endpoint columns replaced by their limits (:79-82), trapz half weights (:90)
            Real s, c, x;
            Real qr = Q.q * cc.c[0];
            Real f0 = (Real)0.5 * (bessel_j1(qr) / qr);
            sincos_qr(Q, cc.c[1], s, c, x);
            Real f1 = s / x;
            acc = qt.wt[0] * (f0 * f0) + qt.wt[qt.D - 1] * (f1 * f1);
            j0 = 1; j1e = qt.D - 1;
        }
        for (int j = j0; j < j1e; ++j) {
            Real s, c, hb;
            Real A = Q.q * nodes[j];
            sincos_qr(Q, nodes[qt.D + j], s, c, hb);    // hb = B/2
            Real f = bessel_j1(A) * s * recip(A * ((Real)2 * hb));   // :74-76
            acc += qt.wt[j] * (f * f);
        }
        return acc * cc.amp;                            // weights carry the factor 16 (:90)
    }
    // fp32 production path: the thread's M q points go through the node loop together (node
    // data loaded once, M independent chains in flight).  The loop is bound by the XU pipe, so it keeps
    // MUFU operations out of the (q, node) pairs where it can: 1 / (A B) is a per-q factor times ONE
    // reciprocal per node, and J1 takes 1/x and sqrt(2/(pi x)) from one reciprocal square root
    // (7 -> 5.25 MUFU per pair: configs[2] 2.5e6 -> 3.7e6 MC steps/s); eval() differs by rounding only
    template <int M>
    __device__ __forceinline__ static void evalV(const Contrib<float> &cc, const float *nodes,
                                                 const QuadTab<float> &qt, const QRep<float> (&Q)[M],
                                                 float (&out)[M])
    {
        float acc[M];
        int j0 = 0, j1e = qt.D;
#pragma unroll
        for (int i = 0; i < M; ++i) acc[i] = 0.f;
        if (qt.rule == MCSAS_QUAD_REFERENCE) {
            float qr[M], jv[M], s[M], x[M];
#pragma unroll
            for (int i = 0; i < M; ++i) qr[i] = Q[i].q * cc.c[0];
            bessel_j1_vec<M>(qr, jv);
            sin_qr_vec<M>(Q, cc.c[1], s, x);
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float f0 = 0.5f * (jv[i] / qr[i]);
                float f1 = s[i] / x[i];
                acc[i] = qt.wt[0] * (f0 * f0) + qt.wt[qt.D - 1] * (f1 * f1);
            }
            j0 = 1; j1e = qt.D - 1;
        }
        // 1 / (A B) = 1 / (2 q^2 nj mj): a per-q factor and ONE reciprocal per node instead of one per (q, node)
        float iq2[M];
#pragma unroll
        for (int i = 0; i < M; ++i) iq2[i] = 0.5f * fast_rcp(Q[i].q * Q[i].q);
        for (int j = j0; j < j1e; ++j) {
            const float nj = nodes[j], mj = nodes[qt.D + j], wj = qt.wt[j];
            const float inm = fast_rcp(nj * mj);
            float A[M], jv[M], s[M], hb[M], f[M];
#pragma unroll
            for (int i = 0; i < M; ++i) A[i] = Q[i].q * nj;
            sin_qr_vec<M>(Q, mj, s, hb);                   // hb = B/2
            bessel_j1_vec<M>(A, jv);
#pragma unroll
            for (int i = 0; i < M; ++i) f[i] = iq2[i] * inm;
#pragma unroll
            for (int i = 0; i < M; ++i) f[i] = jv[i] * s[i] * f[i];          // :74-76
#pragma unroll
            for (int i = 0; i < M; ++i) acc[i] += wj * (f[i] * f[i]);
        }
#pragma unroll
        for (int i = 0; i < M; ++i) out[i] = acc[i] * cc.amp;
    }
};

// ----------------------------------------------------------- GaussianChain --
template <class Real> struct Ops<MCSAS_GAUSSIAN_CHAIN, Real> {
    static constexpr int NARR = 0;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &,
                               Contrib<Real> &cc, Real *)
    {
        double rg = p[0], bp = p[1], etas = p[2], k = p[3];
        double beta = bp - (k * (rg * rg)) * etas;      // :56
        VWS o;
        double vol = k * (rg * rg);                     // :64
        o.v = vol; o.s = 0.;
        if (sizeof(Real) == 4) {
            set_amp<FAST>(o, cc, vol, ce, beta * beta, G, 1.);
            cc.c[1] = (Real)1;
        } else {
            set_amp<FAST>(o, cc, vol, ce, 1., G, 1.);
            cc.c[1] = (Real)beta;
        }
        cc.c[0] = (Real)(rg * Units<Real>::LEN);
        return o;
    }
    __device__ __forceinline__ static double shape(double u)
    {
        return sqrt(2.) * sqrt(expm1(-u) + u) / u;      // :58
    }
    __device__ __forceinline__ static float shape(float u)
    {
        // 2 (e^-u - 1 + u)/u^2 = 1 - u/3 + u^2/12 - ...  for small u (cancellation)
        float ser = 1.f + u * (-0.33333333f + u * (0.083333333f + u * (-0.016666667f
                    + u * (2.7777778e-3f + u * (-3.9682540e-4f + u * 4.9603175e-5f)))));
        float big = 2.f * (__expf(-u) - 1.f + u) * fast_rcp(u * u);
        return sqrtf(u < 0.5f ? ser : big);
    }
    __device__ __forceinline__ static Real eval(const Contrib<Real> &cc, const Real *,
                                                const QuadTab<Real> &, const QRep<Real> &Q)
    {
        Real x = Q.q * cc.c[0];
        Real u = x * x;                                 // :57
        Real F = shape(u) * cc.c[1];
        if (Q.q <= (Real)0) F = cc.c[1];                // :60
        return (F * F) * cc.amp;
    }
};

// -------------------------------------------------------------- Kholodenko --
// p0^2 = int_0^x core(z) dz (kholodenko.py:16-38): Gauss-Legendre near the origin plus a
// closed-form exponential series for the rest (kho_p0sq_series).  sinh ratios are written
// with decaying exponentials only (no overflow).  Evaluated in fp64 in both modes (the
// integrand needs ~1e-10 cancellation accuracy); fp32 mode only changes the LSQ side.
static __constant__ double GL16_X[8] = {
    0.0950125098376374401853193, 0.2816035507792589132304605, 0.4580167776572273863424194,
    0.6178762444026437484466718, 0.7554044083550030338951012, 0.8656312023878317438804679,
    0.9445750230732325760779884, 0.9894009349916499325961542 };
static __constant__ double GL16_W[8] = {
    0.1894506104550684962853967, 0.1826034150449235888667637, 0.1691565193950025381893121,
    0.1495959888165767320815017, 0.1246289712555338720524763, 0.0951585116824927848099251,
    0.0622535239386478928628438, 0.0271524594117540948517806 };

static __device__ __noinline__ double kho_core(double z, int regime, double ef, double x)
{
    if (z <= 0.0) return 1.0;
    double fz;
    double em2 = exp(-2. * z);
    if (regime < 0) {            // q < 3/lK: sinh(e z)/(e sinh z)
        fz = exp((ef - 1.) * z) * -expm1(-2. * ef * z) / (ef * (1. - em2));
    } else if (regime > 0) {     // q > 3/lK: sin(f z)/(f sinh z)
        fz = 2. * sin(ef * z) * exp(-z) / (ef * (1. - em2));
    } else {
        fz = 2. * z * exp(-z) / (1. - em2);
    }
    return fz * (2. / x) * (1.0 - z / x);
}

#ifndef MCSAS_KHO_SERIES
#define MCSAS_KHO_SERIES 1
#endif

// Composite-panel integration over the whole range: panels follow the oscillation sin(f z)
// and the decay 1/sinh z (up to ~45 k integrand evaluations at q lK/3 ~ 170).  Kept as the
// cross-check of kho_p0sq_series (compile with -DMCSAS_KHO_SERIES=0).
static __device__ __noinline__ double kho_p0sq_panels(int regime, double ef, double x)
{
    // the integrand is negligible beyond zc for the oscillating branch; the
    // sinh-ratio branch decays like exp(-(1-e) z) and may need the full range
    double rate = regime < 0 ? (1. - ef) : 1.;
    double zc = fmin(x, 42. / fmax(rate, 1e-3));
    double h = 1.0;                                 // panel width
    if (regime > 0) h = fmin(1.0, 2.5 / ef);
    int np = (int)ceil(zc / h);
    if (np < 4) np = 4;
    h = zc / np;
    double acc = 0.;
    for (int i = 0; i < np; ++i) {
        double c0 = (i + 0.5) * h, s = 0.;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            double dz = 0.5 * h * GL16_X[j];
            s += GL16_W[j] * (kho_core(c0 - dz, regime, ef, x) + kho_core(c0 + dz, regime, ef, x));
        }
        acc += s;
    }
    return acc * 0.5 * h;
}

// int_{z0}^{x} (1 - z/x) exp(-b z) dz for b > 0, stable for b (x - z0) -> 0.
// eb0 = exp(-b z0), ebL = exp(-b (x - z0)).
static __device__ __forceinline__ double kho_tail_exp(double b, double eb0, double ebL, double z0, double x)
{
    double L = x - z0, y = b * L, p1, p2;
    if (y < 0.5) {          // phi1 = sum (-y)^n/(n+1)!, phi2 = sum (-y)^n/(n! (n+2))
        double term = 1.;
        p1 = 0.; p2 = 0.;
#pragma unroll 1
        for (int n = 0; n < 15; ++n) {
            p1 += term / (n + 1);
            p2 += term / (n + 2);
            term *= -y / (n + 1);
        }
    } else {
        double iy = 1. / y;
        p1 = (1. - ebL) * iy;
        p2 = (1. - ebL * (1. + y)) * (iy * iy);
    }
    return eb0 * ((1. - z0 / x) * L * p1 - (L * L / x) * p2);
}

// p0^2 = int_0^x f(z) (2/x)(1 - z/x) dz without walking the oscillations: on [0, z0] two
// 16-point Gauss-Legendre panels (z0 <= 1 and f z0 <= 40, so at most ~6 periods); on
// [z0, x] the expansion 1/sinh z = 2 sum_k exp(-(2k+1) z), whose terms times sin(f z) or
// sinh(e z) and the linear factor integrate in closed form.  The term ratio is
// exp(-2 z0): ~13/z0 terms for 1e-11.  Agrees with QUADPACK (epsrel 1e-10, the reference's
// setting, kholodenko.py:31-38) to ~2e-11 over the model's parameter ranges.
static __device__ __noinline__ double kho_p0sq_series(int regime, double ef, double x)
{
    if (regime == 0) regime = -1;                   // z/sinh z = lim e->0 sinh(e z)/(e sinh z)
    ef = fmax(ef, 1e-6);                            // (e z)^2/6 < 1e-10 where the integrand matters
    double z0 = fmin(1.0, x);
    if (regime > 0) z0 = fmin(z0, 40. / ef);
    const double h = 0.5 * z0, ix = 1. / x;
    double acc = 0.;
    // z <= 1 here: sinh / sin directly (no overflow, no cancellation), four independent
    // nodes in flight per iteration
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
        const double dz = 0.5 * h * GL16_X[j];
        double zz[4] = {0.5 * h - dz, 0.5 * h + dz, 1.5 * h - dz, 1.5 * h + dz};
        double s = 0.;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            double num = regime > 0 ? sin(ef * zz[m]) : sinh(ef * zz[m]);
            s += num / sinh(zz[m]) * (1. - zz[m] * ix);
        }
        acc += GL16_W[j] * s;
    }
    acc *= 0.5 * h / ef;
    if (z0 >= x) return acc * (2. * ix);
    const double r = exp(-2. * z0);
    int K = (int)ceil(12.7 / z0);
    if (K > 8192) K = 8192;
    double ser = 0.;
    if (regime > 0) {
        // Im int (1 - z/x) exp(s z) dz,  s = -a + i f:  exp(s z) [ (1 - z/x)/s + 1/(x s^2) ]
        double s0, c0, sx, cx;
        sincos(ef * z0, &s0, &c0);
        sincos(ef * x, &sx, &cx);
        double ea0 = exp(-z0), eax = x < 700. ? exp(-x) : 0.;
        const double rx = x < 350. ? exp(-2. * x) : 0.;
        const double lin = 1. - z0 * ix, f2 = ef * ef;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            double a = 2. * k + 1.;
            double id = 1. / (a * a + f2);
            double isr = -a * id, isi = -ef * id;                       // 1/s
            double qr = (isr * isr - isi * isi) * ix, qi = 2. * isr * isi * ix;   // 1/(x s^2)
            double up = eax * (cx * qi + sx * qr);
            double lr = lin * isr + qr, li = lin * isi + qi;
            double lo = ea0 * (c0 * li + s0 * lr);
            ser += up - lo;
            ea0 *= r; eax *= rx;
        }
        ser *= 2. / ef;
    } else {
        // sinh(e z)/sinh z = sum_k exp(-(a - e) z) - exp(-(a + e) z),  a = 2k + 1; the
        // exponentials at both ends by recurrence (all factors <= 1)
        const double L = x - z0;
        double em = exp(-(1. - ef) * z0), ep = exp(-(1. + ef) * z0);
        double gm = exp(-(1. - ef) * L), gp = L * (1. + ef) < 700. ? exp(-(1. + ef) * L) : 0.;
        const double rL = L < 350. ? exp(-2. * L) : 0.;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            double a = 2. * k + 1.;
            ser += kho_tail_exp(a - ef, em, gm, z0, x) - kho_tail_exp(a + ef, ep, gp, z0, x);
            em *= r; ep *= r; gm *= rL; gp *= rL;
        }
        ser *= 1. / ef;
    }
    return (acc + ser) * (2. * ix);
}

template <class Real> struct Ops<MCSAS_KHOLODENKO, Real> {
    static constexpr int NARR = 0;
    template <bool FAST, class PV>
    __device__ static VWS prep(const PV &p, double ce, double G, const QuadTab<Real> &,
                               Contrib<Real> &cc, Real *)
    {
        double R = p[0], lk = p[1], L = p[2];
        VWS o;
        double vol = M_PI * L * (R * R);                // :93
        o.v = vol; o.s = 0.;
        set_amp<FAST>(o, cc, vol, ce, 1., G, 1.);
        cc.c[0] = (Real)(R * Units<Real>::LEN);
        cc.c[1] = (Real)(lk * Units<Real>::LEN);
        cc.c[2] = (Real)(3. * L / lk);                  // x, :86
        return o;
    }
    __device__ __noinline__ static Real eval(const Contrib<Real> &cc, const Real *,
                                             const QuadTab<Real> &, const QRep<Real> &Q)
    {
        double q = (double)Q.q, R = (double)cc.c[0], lk = (double)cc.c[1], x = (double)cc.c[2];
        double u = q * R;
        double pcs = (u <= 0.0) ? 1.0 : 2. * j1(u) / u; // :40-45
        double t = q * q * lk * lk / 9.;
        int regime = (q < 3.0 / lk) ? -1 : ((q > 3.0 / lk) ? 1 : 0);
        double ef = regime < 0 ? sqrt(1.0 - t) : (regime > 0 ? sqrt(t - 1.0) : 0.);
#if MCSAS_KHO_SERIES
        double acc = kho_p0sq_series(regime, ef, x);
#else
        double acc = kho_p0sq_panels(regime, ef, x);
#endif
        double F2 = acc * (pcs * pcs);                  // (p0 * pcs)^2, :90
        return (Real)F2 * cc.amp;
    }
};

// ---- smearing (sasmodel.py:56-73): only the reference's canSmear models -----------------
template <int MODEL> struct CanSmear { static constexpr bool value = MODEL == MCSAS_SPHERE || MODEL == MCSAS_LMA_DENSE_SPHERE; };

// Smeared intensity of one contribution at one q point.  Out of line on purpose: the loop
// (fp64 location + phase split per profile point) must not bloat the unsmeared hot loops.
template <int MODEL, class Real>
__device__ __noinline__ Real eval_q_smeared(const Contrib<Real> &cc, const Real *nodes, const QuadTab<Real> &qt,
                                            double q)
{
    using O = Ops<MODEL, Real>;
    Real acc = 0;
    for (int k = 0; k < qt.sm->n; ++k) {
        const double o = qt.sm->off[k];
        QRep<Real> Qk;
        Qk.set(qt.sm->two_d ? q + o : sqrt(q * q + o * o));      // sasconfig.py:326-337
        acc += (Real)qt.sm->tw[k] * O::eval(cc, nodes, qt, Qk);
    }
    return acc;
}

// Intensity of one contribution at one q point, smeared over the profile when one is set.
template <int MODEL, class Real>
__device__ __forceinline__ Real eval_q(const Contrib<Real> &cc, const Real *nodes, const QuadTab<Real> &qt,
                                       const QRep<Real> &Q)
{
    if (!CanSmear<MODEL>::value || qt.nsm == 0) return Ops<MODEL, Real>::eval(cc, nodes, qt, Q);
    return eval_q_smeared<MODEL, Real>(cc, nodes, qt, Q.q_si());
}
