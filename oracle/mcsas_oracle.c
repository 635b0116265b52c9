/*
 * oracle/mcsas_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or
 * executed by the product path; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it).
 *
 * Plain-C, float64 restatement of the Monte-Carlo fitting core of McSAS 1.3.1
 * (BAMresearch/McSAS, mounted at /root/reference in the build container).
 * Every function cites the reference file:line it follows.  Pinned against
 *   - the SASfit golden curves of the reference (testdata/sasfit_*.dat),
 *   - chain-level vectors recorded from the UNMODIFIED reference by
 *     oracle/gen_golden.py (npz files under tests/golden),
 * see tests/test_oracle_golden.py.
 *
 * Third-party arithmetic of the reference that is restated, not imported:
 *   scipy.optimize.leastsq / fmin (backgroundscalingfit.py:101,109): the
 *     residual is linear in (scale, background), so the minimum is the
 *     closed-form weighted least-squares solution (normal equations);
 *   scipy.special.j1 (cylindersisotropic.py:74, kholodenko.py:43): libm j1();
 *   scipy.integrate.quad (QUADPACK QAGS, kholodenko.py:32-34, epsrel 1e-10):
 *     adaptive Gauss-Kronrod (G10,K21) bisection to the same tolerance;
 *   numpy.random.uniform (numbergenerator.py:31,173): the generator OUTPUT
 *     stream is an input of the oracle (recorded from numpy, or Philox below).
 *
 * Build: make -C oracle   ->  oracle/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_PARAMS 16
#define ORC_MAX_ACTIVE 4

/* model ids: keep identical to include/mcsas_b200.h (tests assert it) */
enum {
    ORC_SPHERE = 0, ORC_LMA_DENSE_SPHERE = 1, ORC_SPHERICAL_CORE_SHELL = 2,
    ORC_ELLIPSOIDAL_CORE_SHELL = 3, ORC_ELLIPSOIDS_ISOTROPIC = 4,
    ORC_CYLINDERS_ISOTROPIC = 5, ORC_GAUSSIAN_CHAIN = 6, ORC_KHOLODENKO = 7
};

/* quadrature rule for the orientation-averaged models */
enum { ORC_QUAD_REFERENCE = 0, ORC_QUAD_GAUSS_LEGENDRE = 1 };

typedef struct {
    int model_id;
    int n_params;
    double params[ORC_MAX_PARAMS];    /* reference `parameters` tuple order, SI units */
    int n_active;
    int active_idx[ORC_MAX_ACTIVE];   /* index into params[], model.activeParams() order */
    double lo[ORC_MAX_ACTIVE], hi[ORC_MAX_ACTIVE]; /* activeRange ∩ valueRange */
    int quad_rule, quad_nodes;
} orc_model;

typedef struct {
    int n_contribs;            /* numContribs   mcsasparameters.json */
    double conv_crit;          /* convergenceCriterion */
    long long max_iter;        /* maxIterations */
    double comp_exp;           /* compensationExponent */
    int find_background;       /* findBackground */
    int positive_background;   /* positiveBackground */
    int start_from_minimum;    /* startFromMinimum */
    double q_limit_hi;         /* data.x0.limit[1]  (mcsas.py:313) */
} orc_settings;

/* ---------------------------------------------------------------- models -- */

/* sphere.py:62  F = 3 (sin qr - qr cos qr) / (qr)^3 */
static double sph_ff(double qr)
{
    return 3. * (sin(qr) - qr * cos(qr)) / pow(qr, 3.);
}

/* lmadensesphere.py:76-86 */
static double lma_G(double A, double mu)
{
    double alpha = pow(1. + 2. * mu, 2.) / pow(1. - mu, 4.);
    double beta = -6. * mu * pow(1. + mu / 2., 2.) / pow(1. - mu, 4.);
    double gamma = mu * alpha / 2.;
    double sA = sin(A), cA = cos(A);
    double A2 = A * A, A3 = A2 * A, A4 = A2 * A2, A5 = A4 * A;
    return alpha * (sA - A * cA) / A2
         + beta * (2. * A * sA + (2. - A2) * cA - 2.) / A3
         + gamma * (-1. * A4 * cA + 4. * ((3. * A2 - 6.) * cA
                    + (A3 - 6. * A) * sA + 6.)) / A5;
}

/* ellipsoidalcoreshell.py:60-61  j1(x) = (sin x - x cos x)/x^2 */
static double ecs_j1(double x) { return (sin(x) - x * cos(x)) / (x * x); }

/* kholodenko.py:16-29 */
typedef struct { double q, lk, x; } kho_args;
static double kho_core(double z, const kho_args *a)
{
    double q = a->q, lk = a->lk, x = a->x, fz;
    if (z <= 0.0 || x <= 0.0) return 1.0;
    double ratio = 3.0 / lk;
    if (q < ratio) {
        double e = sqrt(1.0 - q * q * lk * lk / 9.);
        fz = sinh(e * z) / (e * sinh(z));
    } else if (q > ratio) {
        double f = sqrt(q * q * lk * lk / 9. - 1.0);
        fz = sin(f * z) / (f * sinh(z));
    } else {
        fz = z / sinh(z);
    }
    return fz * (2. / x) * (1.0 - z / x);
}

/* Gauss-Kronrod (10,21) nodes/weights (QUADPACK qk21 constants, published). */
static const double GK21_X[11] = {
    0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
    0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
    0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
    0.0 };
static const double GK21_WK[11] = {
    0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
    0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
    0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
    0.149445554002916905664936468389821 };
static const double GK21_WG[5] = {
    0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
    0.295524224714752870173815619188769 };

static void gk21(const kho_args *a, double lo, double hi, double *res, double *err)
{
    double c = 0.5 * (lo + hi), h = 0.5 * (hi - lo);
    double fc = kho_core(c, a);
    double rk = GK21_WK[10] * fc, rg = 0.;
    for (int j = 0; j < 10; ++j) {
        double dx = h * GK21_X[j];
        double f1 = kho_core(c - dx, a), f2 = kho_core(c + dx, a);
        rk += GK21_WK[j] * (f1 + f2);
        if (j & 1) rg += GK21_WG[j / 2] * (f1 + f2);
    }
    *res = rk * h;
    *err = fabs((rk - rg) * h);
}

/* kholodenko.py:31-38: quad(core, 0, x, epsabs=0, epsrel=1e-10, limit=10000).
 * Globally adaptive bisection on the interval with the largest error. */
static double kho_integral(double q, double lk, double x)
{
    enum { MAXIV = 4096 };
    static __thread double lo[MAXIV], hi[MAXIV], r[MAXIV], e[MAXIV];
    kho_args a = { q, lk, x };
    int n = 1;
    lo[0] = 0.; hi[0] = x;
    gk21(&a, 0., x, &r[0], &e[0]);
    for (;;) {
        double tot = 0., terr = 0.;
        int worst = 0;
        for (int i = 0; i < n; ++i) {
            tot += r[i]; terr += e[i];
            if (e[i] > e[worst]) worst = i;
        }
        if (terr <= 1e-10 * fabs(tot) || n >= MAXIV - 1
            || terr <= 50. * 2.2e-16 * fabs(tot))
            return tot;
        double m = 0.5 * (lo[worst] + hi[worst]);
        if (m <= lo[worst] || m >= hi[worst]) return tot;
        lo[n] = m; hi[n] = hi[worst]; hi[worst] = m;
        gk21(&a, lo[worst], hi[worst], &r[worst], &e[worst]);
        gk21(&a, lo[n], hi[n], &r[n], &e[n]);
        ++n;
    }
}

/* Gauss-Legendre nodes on [-1,1] by Newton iteration (published algorithm). */
static void gauss_legendre(int n, double *x, double *w)
{
    for (int i = 0; i < n; ++i) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1., p2 = 0.;
            for (int j = 0; j < n; ++j) {
                double p3 = p2; p2 = p1;
                p1 = ((2. * j + 1.) * z * p2 - j * p3) / (j + 1.);
            }
            pp = n * (z * p1 - p2) / (z * z - 1.);
            double dz = p1 / pp;
            z -= dz;
            if (fabs(dz) < 1e-15) break;
        }
        x[i] = z;
        w[i] = 2. / ((1. - z * z) * pp * pp);
    }
}

/* Full parameter vector with the active values patched in
 * (scatteringmodel.py:94-96: p.setValue(v) for each active param). */
static void patch_params(const orc_model *m, const double *pvals, double *p)
{
    memcpy(p, m->params, sizeof(double) * ORC_MAX_PARAMS);
    for (int i = 0; i < m->n_active; ++i) p[m->active_idx[i]] = pvals[i];
}

/* One contribution: it[nq] = F(q)^2 * w, v = absVolume, w = volume^(2c), s = surface
 * (sasmodel.py:46-79 non-smeared branch, sasmodel.py:37-44, scatteringmodel.py:27-42). */
static int orc_eval_plain(const orc_model *m, int nq, const double *q, const double *pvals,
                          double comp_exp, double *it, double *v_out, double *w_out, double *s_out);

/* Smearing (sasmodel.py:56-73, sasconfig.py:122-149, 209-233, 308-339): for canSmear models
 * (Sphere, LMADenseSphere) the intensity at q_i is 2 * trapz(F^2 w weight, x = qOffset) over the
 * locations sqrt(q_i^2 + off_k^2) (slit) or q_i + off_k (2-D averaged).  The profile (qOffset,
 * weights) comes from SmearingConfig.prepared; it is process-global here like the reference's
 * data.config.smearing. */
#define ORC_MAX_SMEAR 64
static struct { int n, two_d; double off[ORC_MAX_SMEAR], wt[ORC_MAX_SMEAR]; } g_smear = {0, 0, {0}, {0}};

int orc_set_smearing(int n, const double *q_offset, const double *weights, int two_d_coll)
{
    if (n < 0 || n > ORC_MAX_SMEAR || n == 1) return -1;
    g_smear.n = n; g_smear.two_d = two_d_coll;
    for (int k = 0; k < n; ++k) { g_smear.off[k] = q_offset[k]; g_smear.wt[k] = weights[k]; }
    return 0;
}

int orc_eval(const orc_model *m, int nq, const double *q, const double *pvals,
             double comp_exp, double *it, double *v_out, double *w_out, double *s_out)
{
    const int n = g_smear.n;
    const int can_smear = m->model_id == ORC_SPHERE || m->model_id == ORC_LMA_DENSE_SPHERE;  /* canSmear */
    if (n == 0 || !can_smear) return orc_eval_plain(m, nq, q, pvals, comp_exp, it, v_out, w_out, s_out);
    double *locs = (double *)malloc(sizeof(double) * 2 * (size_t)nq * n), *il = locs + (size_t)nq * n;
    for (int i = 0; i < nq; ++i)
        for (int k = 0; k < n; ++k)                       /* sasconfig.py:326-337 */
            locs[(size_t)i * n + k] = g_smear.two_d ? q[i] + g_smear.off[k]
                                                    : sqrt(q[i] * q[i] + g_smear.off[k] * g_smear.off[k]);
    int rc = orc_eval_plain(m, nq * n, locs, pvals, comp_exp, il, v_out, w_out, s_out);
    for (int i = 0; i < nq && rc == 0; ++i) {            /* 2 * trapz(y, x = qOffset)  sasmodel.py:72-73 */
        double acc = 0.;
        for (int k = 0; k + 1 < n; ++k)
            acc += (g_smear.off[k + 1] - g_smear.off[k])
                   * (il[(size_t)i * n + k] * g_smear.wt[k] + il[(size_t)i * n + k + 1] * g_smear.wt[k + 1]) / 2.;
        it[i] = 2. * acc;
    }
    free(locs);
    return rc;
}

static int orc_eval_plain(const orc_model *m, int nq, const double *q, const double *pvals,
                          double comp_exp, double *it, double *v_out, double *w_out, double *s_out)
{
    double p[ORC_MAX_PARAMS];
    patch_params(m, pvals, p);
    double vol = 0., absvol = 0., surf = 0.;
    int D = m->quad_nodes > 0 ? m->quad_nodes : 100;
    double *gx = NULL, *gw = NULL;
    if (m->quad_rule == ORC_QUAD_GAUSS_LEGENDRE) {
        gx = (double *)malloc(sizeof(double) * 2 * D);
        gw = gx + D;
        gauss_legendre(D, gx, gw);
    }
    switch (m->model_id) {
    case ORC_SPHERE: {                                   /* sphere.py:12-63 */
        double r = p[0], sld = p[1];
        vol = (M_PI * 4. / 3.) * pow(r, 3.);             /* :44 */
        absvol = vol * sld * sld;                        /* :53 */
        surf = 4. * M_PI * r * r;                        /* :37 */
        for (int i = 0; i < nq; ++i) it[i] = sph_ff(q[i] * r);
        break;
    }
    case ORC_LMA_DENSE_SPHERE: {                         /* lmadensesphere.py:13-102 */
        double r = p[0], mu = p[1], mf = p[2], sld = p[3];
        if (mf == -1.) mf = pow(0.634 / mu, 1. / 3);     /* :73-74 */
        vol = (M_PI * 4. / 3.) * pow(r, 3.);
        absvol = vol * sld * sld;
        for (int i = 0; i < nq; ++i) {
            double F = sph_ff(q[i] * r);                 /* :90 */
            double A = 2. * q[i] * (mf * r);             /* :91 */
            double S = 1. / (1. + 24. * mu * lma_G(A, mu) / A);  /* :93 */
            it[i] = sqrt(F * F * S);                     /* :101 */
        }
        break;
    }
    case ORC_SPHERICAL_CORE_SHELL: {                     /* sphericalcoreshell.py:12-76 */
        double r = p[0], t = p[1], ec = p[2], es = p[3], esol = p[4];
        double vc = 4. / 3 * M_PI * pow(r, 3.), vt = 4. / 3 * M_PI * pow(r + t, 3.);
        double vr = vc / vt;                             /* :61-63 */
        vol = vt; absvol = vt;                           /* :72,:75-76 */
        for (int i = 0; i < nq; ++i) {
            double ks = (es - esol) * sph_ff(q[i] * (r + t));   /* :51-57 (dEta*3*(..)/qr^3) */
            double kc = (es - ec) * sph_ff(q[i] * r);
            it[i] = ks - vr * kc;                        /* :69 */
        }
        break;
    }
    case ORC_ELLIPSOIDAL_CORE_SHELL: {                   /* ellipsoidalcoreshell.py:14-97 */
        double a = p[0], b = p[1], t = p[2], ec = p[3], es = p[4], esol = p[5];
        double vc = 4. / 3. * M_PI * a * b * b;
        double vt = 4. / 3. * M_PI * (a + t) * (b + t) * (b + t);
        double vr = vc / vt;
        vol = vt; absvol = vt;                           /* :93,:96-97 */
        for (int i = 0; i < nq; ++i) {
            double acc = 0.;
            for (int j = 0; j < D; ++j) {
                double mu, wt;
                if (gx) { mu = 0.5 * (gx[j] + 1.); wt = 0.5 * gw[j]; }
                else { mu = (j == D - 1) ? 1. : j * (1. / (D - 1)); wt = 1. / D; } /* :75 linspace, :90 mean */
                double xc = q[i] * sqrt(a * a * mu * mu + b * b * (1. - mu * mu));
                double xt = q[i] * sqrt((a + t) * (a + t) * mu * mu
                                        + (b + t) * (b + t) * (1. - mu * mu));
                double f = (ec - es) * vr * (3. * ecs_j1(xc) / xc)
                         + (es - esol) * 1. * (3. * ecs_j1(xt) / xt);   /* :83-88 */
                acc += wt * f * f;
            }
            it[i] = sqrt(acc);
        }
        break;
    }
    case ORC_ELLIPSOIDS_ISOTROPIC: {                     /* ellipsoidsisotropic.py:15-84 */
        double Ra = p[0], useAspect = p[1], c = p[2], aspect = p[3], sld = p[5];
        double Rc = (useAspect != 0.) ? Ra * aspect : c; /* :60-63 */
        vol = 4. / 3. * M_PI * Ra * Ra * Rc;             /* :80 */
        absvol = vol * sld * sld;                        /* :84 */
        for (int i = 0; i < nq; ++i) {
            double acc = 0.;
            for (int j = 0; j < D; ++j) {
                double al, wt;
                if (gx) { al = 0.25 * M_PI * (gx[j] + 1.); wt = 0.25 * M_PI * gw[j] * sin(al); }
                else {  /* :65 linspace(0, pi/2, intDiv); :71 mean(fsplit^2 sin(alpha)) */
                    al = (j == D - 1) ? M_PI / 2. : j * ((M_PI / 2.) / (D - 1));
                    wt = sin(al) / D;
                }
                double sa = sin(al), ca = cos(al);
                double rp = sqrt(Ra * Ra * sa * sa + Rc * Rc * ca * ca);  /* :56 */
                double f = sph_ff(q[i] * rp);            /* :67-68 */
                acc += wt * f * f;
            }
            it[i] = sqrt(acc);
        }
        break;
    }
    case ORC_CYLINDERS_ISOTROPIC: {                      /* cylindersisotropic.py:16-101 */
        double r = p[0], useAspect = p[1], length = p[2], aspect = p[3], sld = p[5];
        double hL = (useAspect != 0.) ? r * aspect : 0.5 * length;   /* :65-68 */
        vol = M_PI * r * r * (hL * 2.);                  /* :93-98 */
        absvol = vol * sld * sld;                        /* :101 */
        for (int i = 0; i < nq; ++i) {
            double acc = 0.;
            if (gx) {
                for (int j = 0; j < D; ++j) {
                    double x = 0.5 * (gx[j] + 1.), wt = 0.5 * gw[j];
                    double A = q[i] * r * sqrt(1. - x * x), B = q[i] * 2. * hL * x;
                    double f = j1(A) * sin(B / 2.) / (A * B);
                    acc += wt * f * f;
                }
            } else {
                double step = 1. / (D - 1);              /* :56 linspace(0,1,intDiv) */
                for (int j = 0; j < D; ++j) {
                    double f;
                    if (j == 0) {                        /* :79-80 */
                        double qr = q[i] * r;
                        f = 0.5 * (j1(qr) / qr);
                    } else if (j == D - 1) {             /* :82 */
                        f = sin(q[i] * hL) / (q[i] * hL);
                    } else {
                        double x = (double)j * step;
                        double A = q[i] * (r * sqrt(1. - x * x));   /* :70 */
                        double B = q[i] * (2. * hL * x);            /* :71 */
                        f = j1(A) * sin(B / 2.) / (A * B);          /* :74-76 */
                    }
                    double wt = (j == 0 || j == D - 1) ? 0.5 * step : step;  /* trapz :90 */
                    acc += wt * f * f;
                }
            }
            it[i] = sqrt(16. * acc);                     /* :90 */
        }
        break;
    }
    case ORC_GAUSSIAN_CHAIN: {                           /* gaussianchain.py:12-65 */
        double rg = p[0], bp = p[1], etas = p[2], k = p[3];
        double beta = bp - (k * rg * rg) * etas;         /* :56 */
        vol = k * rg * rg; absvol = vol;                 /* :64 */
        for (int i = 0; i < nq; ++i) {
            double u = (q[i] * rg) * (q[i] * rg);        /* :57 */
            double F = sqrt(2.) * sqrt(expm1(-u) + u) / u;   /* :58 */
            F *= beta;
            if (q[i] <= 0.0) F = beta;                   /* :60 */
            it[i] = F;
        }
        break;
    }
    case ORC_KHOLODENKO: {                               /* kholodenko.py:16-94 */
        double R = p[0], lk = p[1], L = p[2];
        double x = 3. * L / lk;                          /* :86 */
        vol = M_PI * L * R * R; absvol = vol;            /* :93 */
        for (int i = 0; i < nq; ++i) {
            double u = q[i] * R;
            double pcs = (u <= 0.0) ? 1.0 : 2. * j1(u) / u;          /* :40-45 */
            double p0 = sqrt(kho_integral(q[i], lk, x));             /* :31-38 */
            it[i] = p0 * pcs;                            /* :90 */
        }
        break;
    }
    default:
        free(gx);
        return -1;
    }
    free(gx);
    double w = pow(vol, 2. * comp_exp);                  /* sasmodel.py:44 */
    for (int i = 0; i < nq; ++i) it[i] = it[i] * it[i] * w;   /* sasmodel.py:77 */
    *v_out = absvol; *w_out = w; *s_out = surf;
    return 0;
}

/* Bare form factor F(q) (ScatteringModel.formfactor) for golden-curve checks. */
int orc_formfactor(const orc_model *m, int nq, const double *q, const double *pvals,
                   double *ff)
{
    double v, w, s;
    int rc = orc_eval(m, nq, q, pvals, 0.0, ff, &v, &w, &s);  /* w = vol^0 = 1 */
    if (rc) return rc;
    for (int i = 0; i < nq; ++i) ff[i] = sqrt(ff[i]);   /* |F| */
    return 0;
}

/* ------------------------------------------- scale/background fit (a9) -- */

/* backgroundscalingfit.py:112-139.  Minimises sum(((I - a M - b)/sigma)^2):
 * closed form of what leastsq converges to (:94-103); sigma==0 -> 1 (:117);
 * findBackground False -> b = 0 (:130-131, residual chiNoBg :66-70);
 * positiveBackground -> |b| (:60-63): if the free optimum has b<0 the
 * constrained optimum lies on b=0.  chi2 = chiSqr(:73-77) in the reference's
 * two-pass residual form; aGoFs :80-84,:137-138. */
void orc_bgscale(int nq, const double *I, const double *sigma, const double *M,
                 int find_bg, int pos_bg, int n_active,
                 double *sc, double *conval, double *agofs)
{
    double S = 0, Sm = 0, Smm = 0, Si = 0, Smi = 0;
    if (nq == 0) { *conval = 1.; if (agofs) *agofs = 1.; return; }  /* :119-120 */
    for (int i = 0; i < nq; ++i) {
        double e = sigma[i] == 0.0 ? 1. : sigma[i];
        double w = 1. / (e * e);
        S += w; Sm += w * M[i]; Smm += w * M[i] * M[i];
        Si += w * I[i]; Smi += w * M[i] * I[i];
    }
    double a, b;
    if (find_bg) {
        double det = S * Smm - Sm * Sm;
        a = (S * Smi - Sm * Si) / det;
        b = (Smm * Si - Sm * Smi) / det;
        if (pos_bg && b < 0.) { a = Smi / Smm; b = 0.; }
    } else {
        a = Smi / Smm; b = 0.;
    }
    sc[0] = a; sc[1] = b;
    double chi = 0., num = 0., den = 0.;
    for (int i = 0; i < nq; ++i) {
        double e = sigma[i] == 0.0 ? 1. : sigma[i];
        double fit = M[i] * a + (find_bg ? b : 0.);      /* dataScaled :86-92 */
        double r = (I[i] - fit) / e;
        chi += r * r;
        num += (I[i] - fit) * (I[i] - fit);
        den += e * e;
    }
    *conval = chi / nq;
    if (agofs) *agofs = num / den * ((double)nq / (double)(nq - n_active));
}

/* ------------------------------------------------- the MC chain (a1-a3) -- */

/* McSAS.mcFit, mcsas.py:287-439.  `stream` holds number-generator OUTPUTS in
 * [0,1) (RandomUniform.get / RandomExponential.get, numbergenerator.py:28-31,
 * 168-175) in the reference's consumption order: N draws per active parameter
 * (param-major, scatteringmodel.py:121-126) unless startFromMinimum, then
 * n_active draws per step (mcsas.py:358).  Values are mapped to the active
 * range by generateValues, parameter.py:83-84: u*(hi-lo)+lo.
 * Returns 0, or -2 if the stream ran out before the chain stopped. */
int orc_chain(const orc_model *m, const orc_settings *st,
              int nq, const double *q, const double *I, const double *sigma,
              const double *stream, long long stream_len,
              double *rset /*[N][P]*/, double *fit_meas /*[nq]*/, double *sc_out /*[2]*/,
              double *conval_out, double *conval0_out,
              long long *num_iter_out, long long *num_moves_out,
              unsigned char *accept_trace /*[max_iter] or NULL*/,
              double *chi_trace /*[max_iter] or NULL: convalt of every step*/)
{
    const int N = st->n_contribs, P = m->n_active;
    long long pos = 0;
    double *ft = (double *)calloc((size_t)nq * 4, sizeof(double));
    double *inew = ft + nq, *iold = ft + 2 * nq, *test = ft + 3 * nq;
    double *wset = (double *)calloc((size_t)N * 3, sizeof(double));
    double *vset = wset + N, *sset = wset + 2 * N;
    int rc = 0;

    if (st->start_from_minimum) {                        /* mcsas.py:310-315 */
        for (int k = 0; k < P; ++k) {
            double mb = m->lo[k] < m->hi[k] ? m->lo[k] : m->hi[k];
            if (mb == 0.) mb = M_PI / st->q_limit_hi;
            for (int i = 0; i < N; ++i) rset[i * P + k] = 1. * mb * .5;
        }
    } else {                                             /* mcsas.py:317 */
        if (stream_len < (long long)N * P) { rc = -2; goto done; }
        for (int k = 0; k < P; ++k)
            for (int i = 0; i < N; ++i)
                rset[i * P + k] = stream[pos++] * (m->hi[k] - m->lo[k]) + m->lo[k];
    }
    /* mcsas.py:319  model.calc(data, rset): cumInt += it, contribution order */
    for (int i = 0; i < N; ++i) {
        orc_eval(m, nq, q, rset + i * P, st->comp_exp, inew, &vset[i], &wset[i], &sset[i]);
        for (int j = 0; j < nq; ++j) ft[j] += inew[j];
    }
    double sc[2] = { 1., 0. }, conval, agofs;
    /* mcsas.py:327-344: initial guess + simplex + LM == the same LSQ minimum */
    orc_bgscale(nq, I, sigma, ft, st->find_background, st->positive_background, P,
                sc, &conval, &agofs);
    if (conval0_out) *conval0_out = conval;

    long long num_iter = 0, num_moves = 0;
    int ri = 0;
    while (N > 1 && conval > st->conv_crit && num_iter < st->max_iter) {  /* :354-357 */
        double rt[ORC_MAX_ACTIVE], vn, wn, sn, vo, wo, so, sct[2], convalt;
        if (pos + P > stream_len) { rc = -2; break; }
        for (int k = 0; k < P; ++k)                      /* :358 */
            rt[k] = stream[pos++] * (m->hi[k] - m->lo[k]) + m->lo[k];
        orc_eval(m, nq, q, rt, st->comp_exp, inew, &vn, &wn, &sn);             /* :360 */
        orc_eval(m, nq, q, rset + ri * P, st->comp_exp, iold, &vo, &wo, &so);  /* :362 */
        for (int j = 0; j < nq; ++j) test[j] = ft[j] - iold[j] + inew[j];      /* :366 */
        orc_bgscale(nq, I, sigma, test, st->find_background, st->positive_background, P,
                    sct, &convalt, &agofs);              /* :376 */
        int acc = convalt < conval;                      /* :379 strict */
        if (acc) {                                       /* :381-384 */
            for (int k = 0; k < P; ++k) rset[ri * P + k] = rt[k];
            sc[0] = sct[0]; sc[1] = sct[1]; conval = convalt;
            memcpy(ft, test, sizeof(double) * nq);
            wset[ri] = wn; vset[ri] = vn; sset[ri] = sn;
            ++num_moves;
        }
        if (accept_trace) accept_trace[num_iter] = (unsigned char)acc;
        if (chi_trace) chi_trace[num_iter] = convalt;
        ri = (ri + 1) % N;                               /* :403 */
        ++num_iter;                                      /* :404 */
    }
    /* mcsas.py:424-429 */
    orc_bgscale(nq, I, sigma, ft, st->find_background, st->positive_background, P,
                sc, &conval, &agofs);
    for (int j = 0; j < nq; ++j) fit_meas[j] = ft[j] * sc[0] + sc[1];
    sc_out[0] = sc[0]; sc_out[1] = sc[1];
    *conval_out = conval;
    *num_iter_out = num_iter; *num_moves_out = num_moves;
done:
    free(ft); free(wset);
    return rc;
}

/* ------------------------------------------------------ Philox4x32-10 -- */
/* Published counter-based RNG (Salmon et al., SC'11); restated here so that
 * the oracle can reproduce the kernel's production proposal stream.
 * Stream convention shared with the CUDA kernel (csrc/philox.cuh):
 *   key = (seed_lo, seed_hi); counter = (idx_lo, idx_hi, chain_id, domain)
 *   domain 0 = per-step draws (idx = step), 1 = initial draws (idx = contribution)
 *   word k of the 4 outputs -> active parameter k; u = (word + 0.5) * 2^-32. */
static void philox_round(uint32_t c[4], const uint32_t k[2])
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void orc_philox4x32(uint64_t seed, uint64_t idx, uint32_t chain, uint32_t domain,
                    uint32_t out[4])
{
    uint32_t k[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t c[4] = { (uint32_t)idx, (uint32_t)(idx >> 32), chain, domain };
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(uint32_t) * 4);
}

/* Generator-output stream (values in [0,1)) the kernel's production mode uses
 * for one chain: gen_kind[k] 0 = RandomUniform, 1/2/3 = RandomExponential with
 * upper = 1/2/3 (numbergenerator.py:168-189). Layout = orc_chain's `stream`. */
void orc_philox_stream(uint64_t seed, uint32_t chain, int N, int P, const int *gen_kind,
                       int start_from_minimum, long long n_steps, double *stream)
{
    long long pos = 0;
    uint32_t o[4];
    if (!start_from_minimum)
        for (int k = 0; k < P; ++k)
            for (int i = 0; i < N; ++i) {
                orc_philox4x32(seed, (uint64_t)i, chain, 1u, o);
                double u = ((double)o[k] + 0.5) * (1.0 / 4294967296.0);
                if (gen_kind[k] > 0) {
                    double up = (double)gen_kind[k];
                    u = (pow(10., up * u) - 1.) / pow(10., up);
                }
                stream[pos++] = u;
            }
    for (long long t = 0; t < n_steps; ++t) {
        orc_philox4x32(seed, (uint64_t)t, chain, 0u, o);
        for (int k = 0; k < P; ++k) {
            double u = ((double)o[k] + 0.5) * (1.0 / 4294967296.0);
            if (gen_kind[k] > 0) {
                double up = (double)gen_kind[k];
                u = (pow(10., up * u) - 1.) / pow(10., up);
            }
            stream[pos++] = u;
        }
    }
}

/* ------------------------------------------- observability (next row f1) -- */
/* McSAS.histogram per-repetition fractions, mcsas.py:548-604.  Outputs are
 * [N] arrays for ONE repetition; totals[4] = vol, num, volsqr, surf totals. */
int orc_fractions(const orc_model *m, const orc_settings *st,
                  int nq, const double *q, const double *I, const double *sigma,
                  const double *rset, double *sc_out,
                  double *vol_frac, double *num_frac, double *volsqr_frac, double *surf_frac,
                  double *min_vol, double *min_num, double *min_volsqr, double *min_surf,
                  double *totals)
{
    const int N = st->n_contribs, P = m->n_active;
    double *ft = (double *)calloc((size_t)nq * 2, sizeof(double));
    double *it = ft + nq;
    double *wset = (double *)calloc((size_t)N * 3, sizeof(double));
    double *vset = wset + N, *sset = wset + 2 * N;
    for (int i = 0; i < N; ++i) {                        /* :552 */
        orc_eval(m, nq, q, rset + i * P, st->comp_exp, it, &vset[i], &wset[i], &sset[i]);
        for (int j = 0; j < nq; ++j) ft[j] += it[j];
    }
    double sc[2], conval, agofs;
    orc_bgscale(nq, I, sigma, ft, st->find_background, st->positive_background, P,
                sc, &conval, &agofs);                    /* :559 */
    sc_out[0] = sc[0]; sc_out[1] = sc[1];
    double tv = 0, tn = 0, tq = 0, ts = 0;
    for (int i = 0; i < N; ++i) {
        vol_frac[i] = wset[i] * sc[0] / vset[i];         /* :565, modeldata.py:57-61 */
        tv += vol_frac[i];
        num_frac[i] = vol_frac[i] / vset[i]; tn += num_frac[i];        /* :567 */
        volsqr_frac[i] = vol_frac[i] * vset[i]; tq += volsqr_frac[i];  /* :569 */
        surf_frac[i] = num_frac[i] * sset[i]; ts += surf_frac[i];      /* :571 */
    }
    for (int c = 0; c < N; ++c) {                        /* :575-594 */
        double v, w, s, mn = INFINITY;
        orc_eval(m, nq, q, rset + c * P, st->comp_exp, it, &v, &w, &s);
        for (int j = 0; j < nq; ++j) {
            double part = sc[0] * it[j];
            if (part != 0.) {
                double r = sigma[j] * vol_frac[c] / part;
                if (r < mn) mn = r;
            }
        }
        min_vol[c] = mn;
        min_num[c] = mn / vset[c];
        min_volsqr[c] = min_num[c] * mn * mn;
        min_surf[c] = min_num[c] * sset[c];
    }
    for (int i = 0; i < N; ++i) {                        /* :596-604 */
        if (tn != 0.) { num_frac[i] /= tn; min_num[i] /= tn; }
        if (tq != 0.) { volsqr_frac[i] /= tq; min_volsqr[i] /= tq; }
        if (ts != 0.) { surf_frac[i] /= ts; min_surf[i] /= ts; }
    }
    totals[0] = tv; totals[1] = tn; totals[2] = tq; totals[3] = ts;
    free(ft); free(wset);
    return 0;
}
