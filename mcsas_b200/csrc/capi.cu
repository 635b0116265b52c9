// capi.cu -- the C-ABI of libmcsas_b200.so (include/mcsas_b200.h): context, uploads,
// launch of the chain / eval / fraction kernels, result fetch.  No torch types, no CPU
// fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <chrono>
#include <thread>

#include "common.cuh"
#include "launch.cuh"

cudaError_t mcsas_launch_fractions(int prec, int model, int grid, size_t smem, cudaStream_t st,
                                   const FractionArgs &a);
cudaError_t mcsas_launch_peaks(int which, int grid, int threads, int iters, float *out,
                               cudaStream_t st);
struct PrepArgs {                 // dataprep.cu
    int n_sets, n_bin, cap, has_sigma;
    const long long *raw_off;
    const double *q, *I, *sigma;
    const double *edges;
    double fu_min, x0_low, x0_high;
    int mask_zero, mask_neg;
    int *scratch;
    int *nq_out;
    double *q_out, *I_out, *s_out;
};
cudaError_t mcsas_launch_prepare(const PrepArgs &a, cudaStream_t st);

static const chain_launch_fn CHAIN_FN[MCSAS_N_MODELS] = {
    mcsas_launch_chain_0, mcsas_launch_chain_1, mcsas_launch_chain_2, mcsas_launch_chain_3,
    mcsas_launch_chain_4, mcsas_launch_chain_5, mcsas_launch_chain_6, mcsas_launch_chain_7};
static const chain_query_fn QUERY_FN[MCSAS_N_MODELS] = {
    mcsas_query_chain_0, mcsas_query_chain_1, mcsas_query_chain_2, mcsas_query_chain_3,
    mcsas_query_chain_4, mcsas_query_chain_5, mcsas_query_chain_6, mcsas_query_chain_7};
static const eval_launch_fn EVAL_FN[MCSAS_N_MODELS] = {
    mcsas_launch_eval_0, mcsas_launch_eval_1, mcsas_launch_eval_2, mcsas_launch_eval_3,
    mcsas_launch_eval_4, mcsas_launch_eval_5, mcsas_launch_eval_6, mcsas_launch_eval_7};

struct mcsas_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // concurrent launches of a slice, one per cluster size
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_begin = nullptr;
    cudaEvent_t ev_side[3] = {nullptr, nullptr, nullptr};
    float span_ms = 0.f;          // first launch start -> last slice end (host gaps between slices included)
    cudaDeviceProp prop;
    std::string err;
    // datasets
    int n_sets = 0;
    int nq_max = 0;
    std::vector<DatasetDesc> sets;
    std::vector<double> h_q;                           // host copy of the q vectors (set after set): the smeared-location
    float4 *d_smloc = nullptr;                         // table of the fp32 sphere kernel is built from them
    size_t smloc_cap = 0;
    DatasetDesc *d_sets = nullptr;
    double *d_data = nullptr;
    size_t sets_cap = 0, data_cap = 0, quad_cap = 0;   // device buffers are kept between uploads (cudaMalloc / cudaFree
                                                       // synchronise the device: they do not belong in a per-fit path)
    double *h_stage = nullptr;                         // pinned staging of mcsas_fetch
    size_t stage_cap = 0;
    // model
    bool have_model = false;
    ModelParams mp;
    SmearTab smear;               // n = 0: no smearing
    double *d_quad = nullptr;     // [2][quad_nodes]
    // run state
    int n_chains = 0, n_contribs = 0, n_reps = 0, n_active = 0;
    int n_slots = 0;              // state slots on the device: 2 per chain (current attempt + backup attempt)
    bool backups = false;         // speculative next attempts of late chains on idle SMs
    int max_attempts = 1;
    std::vector<double> h_conval;                  // per slot: chi-squared after the last slice (mixed slices)
    std::vector<int> cur, bak, bak_state;          // per chain: slot of the current attempt, backup slot (-1), 1 running / 2 finished
    std::vector<char> done;
    std::vector<long long> carry_steps, carry_evals;   // work of the failed attempts left behind in the other slot
    int n_backups_started = 0, n_backups_used = 0;
    long long max_iter = 0;
    bool have_run = false, have_trace = false, running = false;
    double *d_stream = nullptr;
    size_t stream_cap = 0;
    double *d_rset = nullptr, *d_fit = nullptr, *d_scal = nullptr;   // scal: SCAL_N doubles per slot
    double *d_ftsave = nullptr;                                       // fp64 master intensity of unfinished chains
    long long *d_cnt = nullptr;                                       // 4 per slot
    int *d_att = nullptr;                                             // ATT_N ints per slot: attempts | status | batch size
    // pinned mirrors read back once per time slice (one async copy each, one synchronisation)
    int *h_att = nullptr;
    long long *h_cnt = nullptr;
    double *h_cv = nullptr;
    size_t pin_cap = 0;
    unsigned char *d_tacc = nullptr;
    double *d_tchi = nullptr;
    size_t rset_cap = 0, fit_cap = 0, slot_cap = 0, trace_cap = 0;   // elements allocated per output buffer
    int *h_stop = nullptr;        // mapped pinned host int
    int *d_stop = nullptr;
    // time-sliced launch state
    ChainArgs args;
    int precision = 0;
    size_t smem = 0;
    int forced_cluster = 0;
    std::vector<int> h_ids;
    int *d_ids = nullptr;
    long long *d_prof = nullptr;
    float *d_cache = nullptr;
    int *d_cache_lock = nullptr;
    size_t cache_cap = 0;
    int n_live = 0, n_launches = 0, n_slices = 0, last_cluster = 1, last_threads = 0, last_qpt = 0, last_resident = 0;
    int last_sizes[4] = {0, 0, 0, 0};   // chains launched with cluster size 1, 2, 4, 8 in the last slice
    int max_cluster = 8, forced_qpt = 0;
    bool spec_forced = false;                          // run_config.spec_max > 0
    mcsas_tuning tune;                  // experiment knobs (mcsas_set_tuning), all 0 = defaults
    float kernel_ms = 0.f;
    volatile int *stop_flag = nullptr;
    std::map<long long, int> occ_cache;
};

static int fail(mcsas_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define SCAL_N 9
#define ATT_N 3
#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(ctx, MCSAS_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                 \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                         \
    } while (0)

template <class T> static cudaError_t regrow(T **p, size_t n)
{
    cudaFree(*p);
    *p = nullptr;
    return cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
}

static thread_local std::string g_create_err;

extern "C" int mcsas_ctx_create(int device, mcsas_ctx **out)
{
    if (!out) return MCSAS_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        g_create_err = std::string("no usable CUDA device (") +
                       (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                       "); mcsas_b200 has no CPU fallback";
        return MCSAS_ERR_NO_DEVICE;
    }
    mcsas_ctx *ctx = new mcsas_ctx();
    ctx->device = device;
    memset(&ctx->smear, 0, sizeof ctx->smear);
    memset(&ctx->tune, 0, sizeof ctx->tune);
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess
        || cudaEventCreate(&ctx->ev_begin) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->side[0], cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->side[1], cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->side[2], cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&ctx->ev_side[0], cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&ctx->ev_side[1], cudaEventDisableTiming) != cudaSuccess
        || cudaEventCreateWithFlags(&ctx->ev_side[2], cudaEventDisableTiming) != cudaSuccess
        || cudaHostAlloc((void **)&ctx->h_stop, sizeof(int), cudaHostAllocMapped) != cudaSuccess
        || cudaHostGetDevicePointer((void **)&ctx->d_stop, ctx->h_stop, 0) != cudaSuccess) {
        g_create_err = std::string("CUDA context setup failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return MCSAS_ERR_CUDA;
    }
    *ctx->h_stop = 0;
    *out = ctx;
    return MCSAS_OK;
}

extern "C" void mcsas_ctx_destroy(mcsas_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_sets); cudaFree(ctx->d_data); cudaFree(ctx->d_quad); cudaFree(ctx->d_stream);
    cudaFree(ctx->d_rset); cudaFree(ctx->d_fit); cudaFree(ctx->d_scal); cudaFree(ctx->d_cnt);
    cudaFree(ctx->d_ftsave); cudaFreeHost(ctx->h_stage); cudaFreeHost(ctx->h_att); cudaFreeHost(ctx->h_cnt); cudaFreeHost(ctx->h_cv);
    cudaFree(ctx->d_att); cudaFree(ctx->d_tacc); cudaFree(ctx->d_tchi); cudaFree(ctx->d_ids); cudaFree(ctx->d_prof); cudaFree(ctx->d_cache); cudaFree(ctx->d_cache_lock); cudaFree(ctx->d_smloc);
    cudaFreeHost(ctx->h_stop);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev_begin);
    for (int i = 0; i < 3; ++i) { cudaEventDestroy(ctx->ev_side[i]); cudaStreamDestroy(ctx->side[i]); }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *mcsas_last_error(const mcsas_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int mcsas_device_info(mcsas_ctx *ctx, int *sm_count, int *clock_khz, char *name, int name_len)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *clock_khz = khz;
    }
    if (name && name_len > 0) snprintf(name, name_len, "%s", ctx->prop.name);
    return MCSAS_OK;
}

extern "C" int mcsas_upload_datasets(mcsas_ctx *ctx, int n_sets, const int *nq, const double *q,
                                     const double *intensity, const double *sigma)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (ctx->running) return fail(ctx, MCSAS_ERR_STATE, "upload_datasets: a run is in flight (call mcsas_wait first)");
    if (n_sets <= 0 || !nq || !q || !intensity || !sigma)
        return fail(ctx, MCSAS_ERR_ARG, "upload_datasets: null/empty input");
    CK(cudaSetDevice(ctx->device));
    std::vector<DatasetDesc> sets(n_sets);
    long long off = 0, src = 0;
    int nq_max = 0;
    for (int s = 0; s < n_sets; ++s) {
        if (nq[s] < 1 || nq[s] > MCSAS_MAX_NQ)
            return fail(ctx, MCSAS_ERR_ARG, "dataset %d: nq=%d outside [1,%d]", s, nq[s], MCSAS_MAX_NQ);
        sets[s].nq = nq[s];
        sets[s].nq_pad = (nq[s] + 1) & ~1;
        sets[s].offset = off;
        off += 3LL * sets[s].nq_pad;
        if (nq[s] > nq_max) nq_max = nq[s];
    }
    std::vector<double> packed((size_t)off, 0.0);
    for (int s = 0; s < n_sets; ++s) {
        double *dst = packed.data() + sets[s].offset;
        double qmax = -INFINITY, qmin = INFINITY, imax = 0.;
        for (int i = 0; i < nq[s]; ++i) {
            dst[i] = q[src + i];
            dst[sets[s].nq_pad + i] = intensity[src + i];
            dst[2 * sets[s].nq_pad + i] = sigma[src + i];
            if (q[src + i] > qmax) qmax = q[src + i];
            if (q[src + i] > 0. && q[src + i] < qmin) qmin = q[src + i];
            if (fabs(intensity[src + i]) > imax) imax = fabs(intensity[src + i]);
            // sigma = +inf is what DataObj._prepareUncertainty (dataobj.py:224-226) gives a point with an
            // invalid uncertainty: weight 0 in the fit, the point still counts in chi^2 / nq
            const bool sig_ok = std::isfinite(sigma[src + i]) || (std::isinf(sigma[src + i]) && sigma[src + i] > 0.);
            if (!std::isfinite(q[src + i]) || !std::isfinite(intensity[src + i]) || !sig_ok)
                return fail(ctx, MCSAS_ERR_ARG, "dataset %d: non-finite value at point %d", s, i);
        }
        sets[s].q_max = qmax;
        sets[s].q_ref = (qmin < INFINITY && qmax > 0.) ? sqrt(qmin * qmax) : 1e9;
        sets[s].i_scale = (imax > 0. && std::isfinite(imax)) ? imax : 1.;
        src += nq[s];
    }
    if ((size_t)n_sets > ctx->sets_cap) { CK(regrow(&ctx->d_sets, (size_t)n_sets)); ctx->sets_cap = (size_t)n_sets; }
    if (packed.size() > ctx->data_cap) { CK(regrow(&ctx->d_data, packed.size())); ctx->data_cap = packed.size(); }
    CK(cudaMemcpyAsync(ctx->d_sets, sets.data(), sizeof(DatasetDesc) * n_sets, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_data, packed.data(), sizeof(double) * packed.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sets = sets;
    ctx->h_q.assign(q, q + src);
    ctx->n_sets = n_sets;
    ctx->nq_max = nq_max;
    ctx->have_run = false;
    return MCSAS_OK;
}

// Gauss-Legendre nodes/weights on [-1,1] (Newton on P_n).
static void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w)
{
    x.resize(n); w.resize(n);
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

// numpy.linspace(start, stop, n): arange(n) * step + start, last element = stop
static double linspace_at(double start, double stop, int n, int j)
{
    if (n == 1) return start;
    if (j == n - 1) return stop;
    return j * ((stop - start) / (n - 1)) + start;
}

extern "C" int mcsas_set_model(mcsas_ctx *ctx, int model_id, const double *params, int n_params,
                               const int *active_idx, int n_active, const double *active_lo,
                               const double *active_hi, const int *generator_kind, int quad_rule,
                               int quad_nodes)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (ctx->running) return fail(ctx, MCSAS_ERR_STATE, "set_model: a run is in flight (call mcsas_wait first)");
    if (model_id < 0 || model_id >= MCSAS_N_MODELS) return fail(ctx, MCSAS_ERR_ARG, "unknown model id %d", model_id);
    if (n_params < 1 || n_params > MCSAS_MAX_PARAMS || !params) return fail(ctx, MCSAS_ERR_ARG, "bad n_params %d", n_params);
    if (n_active < 0 || n_active > MCSAS_MAX_ACTIVE) return fail(ctx, MCSAS_ERR_ARG, "n_active %d outside [0,%d]", n_active, MCSAS_MAX_ACTIVE);
    CK(cudaSetDevice(ctx->device));
    ModelParams mp;
    memset(&mp, 0, sizeof mp);
    mp.model_id = model_id; mp.n_params = n_params; mp.n_active = n_active;
    for (int i = 0; i < n_params; ++i) mp.params[i] = params[i];
    for (int i = 0; i < n_active; ++i) {
        if (active_idx[i] < 0 || active_idx[i] >= n_params) return fail(ctx, MCSAS_ERR_ARG, "active_idx[%d]=%d out of range", i, active_idx[i]);
        if (!(active_lo[i] <= active_hi[i])) return fail(ctx, MCSAS_ERR_ARG, "active range %d is empty", i);
        if (generator_kind[i] < 0 || generator_kind[i] > 3) return fail(ctx, MCSAS_ERR_ARG, "generator kind %d unknown", generator_kind[i]);
        mp.active_idx[i] = active_idx[i];
        mp.lo[i] = active_lo[i]; mp.hi[i] = active_hi[i];
        mp.gen_kind[i] = generator_kind[i];
    }
    int narr = MODEL_NARR[model_id];
    int D = 0;
    std::vector<double> qx, qw;
    if (narr) {
        D = quad_nodes;
        if (D < 2 || D > MCSAS_MAX_NODES) return fail(ctx, MCSAS_ERR_ARG, "quad_nodes %d outside [2,%d]", D, MCSAS_MAX_NODES);
        qx.resize(D); qw.resize(D);
        std::vector<double> gx, gw;
        if (quad_rule == MCSAS_QUAD_GAUSS_LEGENDRE) gauss_legendre(D, gx, gw);
        else if (quad_rule != MCSAS_QUAD_REFERENCE) return fail(ctx, MCSAS_ERR_ARG, "unknown quad rule %d", quad_rule);
        for (int j = 0; j < D; ++j) {
            if (model_id == MCSAS_ELLIPSOIDAL_CORE_SHELL) {       // ellipsoidalcoreshell.py:75,:90
                if (gx.empty()) { qx[j] = linspace_at(0., 1., D, j); qw[j] = 1. / D; }
                else { qx[j] = 0.5 * (gx[j] + 1.); qw[j] = 0.5 * gw[j]; }
            } else if (model_id == MCSAS_ELLIPSOIDS_ISOTROPIC) {  // ellipsoidsisotropic.py:65,:71
                if (gx.empty()) { qx[j] = linspace_at(0., M_PI / 2., D, j); qw[j] = sin(qx[j]) / D; }
                else { qx[j] = 0.25 * M_PI * (gx[j] + 1.); qw[j] = 0.25 * M_PI * gw[j] * sin(qx[j]); }
            } else {                                              // cylindersisotropic.py:56,:90
                if (gx.empty()) {
                    double step = 1. / (D - 1);
                    qx[j] = linspace_at(0., 1., D, j);
                    qw[j] = 16. * ((j == 0 || j == D - 1) ? 0.5 * step : step);
                } else { qx[j] = 0.5 * (gx[j] + 1.); qw[j] = 16. * 0.5 * gw[j]; }
            }
        }
    }
    mp.quad_rule = quad_rule;
    mp.quad_nodes = D;
    if ((size_t)2 * (D + 1) > ctx->quad_cap) { CK(regrow(&ctx->d_quad, (size_t)2 * (D + 1))); ctx->quad_cap = (size_t)2 * (D + 1); }
    if (D) {
        CK(cudaMemcpy(ctx->d_quad, qx.data(), sizeof(double) * D, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->d_quad + D, qw.data(), sizeof(double) * D, cudaMemcpyHostToDevice));
    }
    ctx->mp = mp;
    ctx->have_model = true;
    ctx->have_run = false;
    return MCSAS_OK;
}

extern "C" int mcsas_set_smearing(mcsas_ctx *ctx, int n_points, const double *q_offset, const double *weights,
                                  int two_d_coll)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (ctx->running) return fail(ctx, MCSAS_ERR_STATE, "set_smearing: a run is in flight (call mcsas_wait first)");
    if (n_points == 0) { memset(&ctx->smear, 0, sizeof ctx->smear); ctx->have_run = false; return MCSAS_OK; }
    if (n_points < 2 || n_points > MCSAS_MAX_SMEAR || !q_offset || !weights)
        return fail(ctx, MCSAS_ERR_ARG, "set_smearing: n_points %d outside [2,%d] or null profile", n_points, MCSAS_MAX_SMEAR);
    SmearTab t;
    memset(&t, 0, sizeof t);
    t.n = n_points; t.two_d = two_d_coll != 0;
    for (int k = 0; k < n_points; ++k) {
        if (!std::isfinite(q_offset[k]) || !std::isfinite(weights[k]))
            return fail(ctx, MCSAS_ERR_ARG, "set_smearing: non-finite profile value at %d", k);
        // 2 * numpy.trapz(y * weights, x = qOffset)  (sasmodel.py:72-73) as a weighted sum of y
        const double left = k > 0 ? q_offset[k] - q_offset[k - 1] : 0.;
        const double right = k + 1 < n_points ? q_offset[k + 1] - q_offset[k] : 0.;
        t.off[k] = q_offset[k];
        t.tw[k] = 2. * weights[k] * 0.5 * (left + right);
    }
    ctx->smear = t;
    ctx->have_run = false;
    return MCSAS_OK;
}

extern "C" int mcsas_eval(mcsas_ctx *ctx, int set, const double *pset, int m, double comp_exp,
                          int precision, double *intensity, double *v, double *w, double *s)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (!ctx->n_sets || !ctx->have_model) return fail(ctx, MCSAS_ERR_STATE, "eval: upload datasets and set a model first");
    if (set < 0 || set >= ctx->n_sets || m < 1 || !pset) return fail(ctx, MCSAS_ERR_ARG, "eval: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const int P = ctx->mp.n_active, nq = ctx->sets[set].nq, D = ctx->mp.quad_nodes;
    double *d_p = nullptr, *d_out = nullptr;
    size_t nout = (size_t)m * nq + 3 * (size_t)m;
    CK(cudaMalloc(&d_p, sizeof(double) * (size_t)m * (P > 0 ? P : 1)));
    CK(cudaMalloc(&d_out, sizeof(double) * nout));
    if (P) CK(cudaMemcpyAsync(d_p, pset, sizeof(double) * (size_t)m * P, cudaMemcpyHostToDevice, ctx->stream));
    EvalArgs a;
    a.mp = ctx->mp; a.smear = ctx->smear; a.sets = ctx->d_sets; a.data = ctx->d_data;
    a.quad_x = ctx->d_quad; a.quad_w = ctx->d_quad + D;
    a.set = set; a.m = m; a.comp_exp = comp_exp; a.pset = d_p;
    a.intensity = d_out; a.v = d_out + (size_t)m * nq; a.w = a.v + m; a.s = a.w + m;
    size_t real = precision == MCSAS_FP64 ? 8 : 4;
    size_t smem = real * (2 * (D + 1) + (size_t)MODEL_NARR[ctx->mp.model_id] * D) + 16;
    cudaError_t e = EVAL_FN[ctx->mp.model_id](precision, m, smem, ctx->stream, a);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    std::vector<double> h(nout);
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost);
    cudaFree(d_p); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, MCSAS_ERR_CUDA, "eval kernel: %s", cudaGetErrorString(e));
    if (intensity) memcpy(intensity, h.data(), sizeof(double) * (size_t)m * nq);
    if (v) memcpy(v, h.data() + (size_t)m * nq, sizeof(double) * m);
    if (w) memcpy(w, h.data() + (size_t)m * nq + m, sizeof(double) * m);
    if (s) memcpy(s, h.data() + (size_t)m * nq + 2 * m, sizeof(double) * m);
    return MCSAS_OK;
}

// Thread layout of one chain: cluster size C (q split), q points per eval thread, eval
// threads per CTA.  More CTAs per chain when few chains are alive, so that freed SMs
// shorten the tail of the slow chains.
struct ChainPlan { int cluster, layout, threads, resident, warps_q; };

static int query_resident(mcsas_ctx *ctx, int c, int layout, int block_threads, int *resident)
{
    const int smear = ctx->smear.n > 0;
    const long long key = (((((long long)ctx->mp.model_id * 2 + ctx->precision) * 2 + smear) * 16 + c) * 8 + layout) * 2048
                          + block_threads + ((long long)ctx->smem << 26);
    auto it = ctx->occ_cache.find(key);
    if (it != ctx->occ_cache.end()) { *resident = it->second; return 0; }
    int regs = 0;
    *resident = 0;
    cudaError_t e = QUERY_FN[ctx->mp.model_id](ctx->precision, smear, layout, block_threads, ctx->smem, c, resident, &regs);
    if (e != cudaSuccess) { cudaGetLastError(); *resident = 0; }
    ctx->occ_cache[key] = *resident;
    return 0;
}

// Candidate layouts of the fp64 parity kernel: cluster size C in {1,2,4,8} x eval threads T
// in {64,128,256}; QPT is the smallest instantiated value that covers nq.  Only layouts whose
// clusters are ALL co-resident (cudaOccupancyMaxActiveClusters) are considered, so a slice is
// one wave; among those the one that spreads the live chains over the most SMs wins, ties go
// to the smaller cluster (less exchange per batch).
static int plan_chain_fp64(mcsas_ctx *ctx, int nq, int n_live, ChainPlan *out)
{
    const int sms = ctx->prop.multiProcessorCount;
    const int forced_c = ctx->forced_cluster, forced_t = 0, forced_q = ctx->forced_qpt;
    ChainPlan best = {0, 0, 0, 0, 0};
    double best_score = -1.;
    for (int c = 1; c <= MAX_CLUSTER; c *= 2) {
        if (forced_c > 0 && c != forced_c && !(forced_c > MAX_CLUSTER && c == MAX_CLUSTER)) continue;
        const int n = (nq + c - 1) / c;                  // q points per CTA
        for (int tcap = 64; tcap <= CHAIN_MAX_EVAL_THREADS; tcap *= 2) {
            if (forced_t > 0 && tcap != forced_t) continue;
            int layout = -1;
            for (int l = 0; l < N_CHAIN_LAYOUTS; ++l)
                if (n <= tcap * CHAIN_QPT[l] && (forced_q <= 0 || CHAIN_QPT[l] >= forced_q)) { layout = l; break; }
            if (layout < 0) continue;
            const int qpt = CHAIN_QPT[layout];
            int threads = (((n + qpt - 1) / qpt) + 31) & ~31;
            if (threads < 32) threads = 32;
            if (forced_t <= 0 && tcap > 64 && threads * 2 <= tcap) continue;        // same as a smaller tcap
            int resident = 0;
            query_resident(ctx, c, layout, threads + 32, &resident);
            if (resident < 1) continue;
            const bool one_wave = n_live <= resident;
            const double ctas = (double)n_live * c;
            const double per_sm = ctas / sms;
            double busy = per_sm >= 1. ? 1. : per_sm;                 // fraction of SMs with work
            double warps = (per_sm > 1. ? per_sm : 1.) * (threads / 32);   // eval warps per busy SM
            double score = busy * (warps >= 8. ? 1. : 0.55 + 0.45 * warps / 8.);
            score *= 1. - 0.03 * (c > 1 ? (1 + (c > 2) + (c > 4)) : 0);   // exchange cost per batch
            if (!one_wave) score *= 0.4 * resident / (double)n_live;
            if (score > best_score) { best_score = score; best.cluster = c; best.layout = layout; best.threads = threads; best.resident = resident; }
        }
    }
    if (best.cluster == 0) return fail(ctx, MCSAS_ERR_ARG, "no launchable chain layout for nq=%d", nq);
    *out = best;
    return MCSAS_OK;
}

// Layout of the fp32 production kernel for ONE cluster size C (q split over the cluster) at
// q points per thread QPT = 4 (fewer only when forced): proposal groups PS = spare warps of the
// 8-warp CTA; eval threads = warps_q * PS * 32.
// Measured on B200 (tools/tune.py fast, nq = 1000): 8 eval warps with 4 q points per thread
// are the sweet spot at every cluster size.
static bool plan_fast_layout(mcsas_ctx *ctx, int nq, int c, int max_warps, ChainPlan *out)
{
    const int grp_of[3] = {4, 8, 8};                        // proposals per chunk of layout QPT = 1,2,4
    const int n = (nq + c - 1) / c;
    for (int l = N_CHAIN_LAYOUTS - 1; l >= 0; --l) {        // QPT = 4 first
        const int qpt = CHAIN_QPT[l];
        if (ctx->forced_qpt > 0 && qpt != ctx->forced_qpt) continue;
        const int wq = (n + 32 * qpt - 1) / (32 * qpt);
        if (wq * 32 > FAST_MAX_EVAL_THREADS[l] || wq > max_warps) continue;
        int ps = 1;
        while (wq * ps * 2 <= max_warps && wq * ps * 2 * 32 <= FAST_MAX_EVAL_THREADS[l]
               && ps * 2 * grp_of[l] <= ctx->args.spec_max) ps *= 2;
        const int threads = wq * ps * 32;
        if (threads > FAST_MAX_EVAL_THREADS[l] || threads / 32 > FAST_MAX_WARPS) continue;
        if (c * wq > ctx->args.gw_cap) continue;
        int resident = 0;
        query_resident(ctx, c, l, threads, &resident);
        if (resident < 1) continue;
        out->cluster = c; out->layout = l; out->threads = threads; out->resident = resident; out->warps_q = wq;
        return true;
    }
    return false;
}

static int launch_one_size(mcsas_ctx *ctx, const ChainArgs &base, const ChainPlan &pl, const int *d_ids, int n,
                           cudaStream_t st)
{
    ChainArgs a = base;
    a.cluster = pl.cluster;
    a.warps_q = pl.warps_q;
    // batches of more than FAST_KDEF proposals only where the SMs would otherwise idle (a chain on >= 4 CTAs
    // is one of few left): on a full machine the proposals evaluated behind an accepted move are lost work
    if (ctx->precision == MCSAS_FP32 && !ctx->spec_forced) {
        const int from = ctx->tune.big_batch > 0 ? ctx->tune.big_batch : (ctx->tune.big_batch < 0 ? 1 << 30 : 4);
        if (pl.cluster < from && a.spec_max > FAST_KDEF) a.spec_max = FAST_KDEF;
    }
    a.chain_ids = d_ids;
    const int extra = ctx->precision == MCSAS_FP64 ? 32 : 0;                          // fp64: + producer warp
    cudaError_t e = CHAIN_FN[a.mp.model_id](ctx->precision, pl.layout, n * pl.cluster, pl.threads + extra,
                                            ctx->smem, st, a);
    if (e != cudaSuccess) return fail(ctx, MCSAS_ERR_CUDA, "chain kernel launch (cluster %d): %s", pl.cluster, cudaGetErrorString(e));
    ctx->n_launches++;
    return MCSAS_OK;
}

// Launch one time slice for the state slots listed in ctx->h_ids[0..n_live).
//
// fp32 production kernel: a chain's result does not depend on the cluster size it runs with
// (chain_fast.cuh), so every slice re-packs the machine.  Fewer live chains than SMs: the
// chains furthest from the criterion (largest chi-squared: the likely critical path of the fit)
// are upgraded 1 -> 2 -> 4 -> 8 CTAs until no SM is left; the up to four cluster sizes run as
// concurrent launches, largest clusters first so that they find empty GPCs.  At least as many
// chains as SMs: one launch of 1-CTA chains, the hardware's CTA scheduler fills the SMs.
static int launch_slice(mcsas_ctx *ctx, int n_live, bool resume)
{
    ChainArgs a = ctx->args;
    a.resume = resume ? 1 : 0;
    const int sms = ctx->prop.multiProcessorCount;
    for (int i = 0; i < 4; ++i) ctx->last_sizes[i] = 0;
    if (ctx->precision == MCSAS_FP64) {
        ChainPlan pl;
        int prc = plan_chain_fp64(ctx, ctx->nq_max, n_live, &pl);
        if (prc) return prc;
        ctx->last_threads = pl.threads; ctx->last_qpt = CHAIN_QPT[pl.layout]; ctx->last_resident = pl.resident;
        const int *ids = nullptr;
        if (n_live != ctx->n_chains || resume) {
            CK(cudaMemcpyAsync(ctx->d_ids, ctx->h_ids.data(), sizeof(int) * n_live, cudaMemcpyHostToDevice, ctx->stream));
            ids = ctx->d_ids;
        }
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        int rc = launch_one_size(ctx, a, pl, ids, n_live, ctx->stream);
        if (rc) return rc;
        ctx->last_cluster = pl.cluster;
        ctx->n_slices++;
        return MCSAS_OK;
    }
    // ---- cluster size per chain --------------------------------------------------------------
    ChainPlan plans[4];
    bool have[4] = {false, false, false, false};
    int maxc = ctx->forced_cluster > 0 ? ctx->forced_cluster : ctx->max_cluster;
    if (maxc > MAX_CLUSTER) maxc = MAX_CLUSTER;
    // half = 1: 4-warp CTAs, two per SM (a chain's decision / update phases overlap with the other
    // chain's evaluation on the same SM); the machine then holds 2 x #SMs CTAs
    const bool half = ctx->tune.half > 0;
    const int max_warps = half ? 4 : 8;
    const int sms_eff = half ? 2 * sms : sms;
    for (int li = 0, c = 1; li < 4; ++li, c *= 2)
        if (c <= maxc && (ctx->forced_cluster <= 0 || c == ctx->forced_cluster))
            have[li] = plan_fast_layout(ctx, ctx->nq_max, c, max_warps, &plans[li]);
    int lowest = -1;
    for (int li = 0; li < 4; ++li) if (have[li]) { lowest = li; break; }
    if (lowest < 0) return fail(ctx, MCSAS_ERR_ARG, "no launchable fp32 chain layout for nq=%d", ctx->nq_max);
    std::vector<int> level(n_live, lowest);
    if ((int)ctx->h_conval.size() == ctx->n_slots && n_live > 1) {
        // primaries before backup attempts, then by distance from the criterion
        const std::vector<double> &cv = ctx->h_conval;
        const int NC = ctx->n_chains;
        const std::vector<int> &curv = ctx->cur;
        std::stable_sort(ctx->h_ids.begin(), ctx->h_ids.begin() + n_live, [&](int x, int y) {
            const bool px = curv[x % NC] == x, py = curv[y % NC] == y;
            if (px != py) return px;
            return cv[x] > cv[y];
        });
    }
    if (ctx->forced_cluster <= 0) {
        long long used = (long long)n_live << lowest;
        int cnt_at[4] = {0, 0, 0, 0};
        cnt_at[lowest] = n_live;
        // A cluster must sit inside one GPC, so a mix of cluster sizes cannot use every SM: a cluster of size
        // c takes 1 / resident(c) of the machine (resident = cudaOccupancyMaxActiveClusters of that size
        // alone).  Without this bound 13 chains on 4 and 12 on 8 CTAs -- 148 CTAs -- ran in TWO waves: every
        // slice of a 25-chain fit took twice its budget (one fit split over 4 GPUs: 29 instead of 20 ms).
        auto share = [&](int li) { return 1. / (double)(plans[li].resident > 0 ? plans[li].resident : 1); };
        double frac = n_live * share(lowest);
        const bool pack = ctx->tune.gpc_pack >= 0;
        auto fits = [&](int li) { return !pack || frac + share(li) - share(li - 1) <= 1.0 + 1e-9; };
        // Late chains first, as far as they can go: the chi-squared of a chain says little about its
        // length before ~20 000 steps (tools/predict_len.py: rank correlation 0.0 at 10 000, 0.4 at
        // 20 000, 0.7 at 40 000 steps); from then on the chains furthest from the criterion are the
        // likely critical path (the ones that will use all maxIterations and need another attempt)
        const long long late = ctx->tune.late_steps > 0 ? ctx->tune.late_steps : (1LL << 60);   // (measured: no gain on config 2, off by default)
        if (ctx->pin_cap >= (size_t)ctx->n_slots && !ctx->h_conval.empty()) {
            for (int i = 0; i < n_live; ++i) {
                if (ctx->h_cnt[ctx->h_ids[i]] < late) continue;
                for (int li = level[i] + 1; li < 4 && have[li]; ++li) {
                    const int cost = (1 << li) / 2;
                    if (used + cost > sms_eff || cnt_at[li] + 1 > plans[li].resident || !fits(li)) break;
                    frac += share(li) - share(li - 1);
                    cnt_at[li]++; cnt_at[level[i]]--; level[i] = li; used += cost;
                }
            }
        }
        for (int li = lowest + 1; li < 4 && used < sms_eff; ++li) {
            if (!have[li]) break;
            const int step_cost = (1 << li) / 2;
            for (int i = 0; i < n_live && used + step_cost <= sms_eff; ++i) {
                if (level[i] != li - 1) continue;
                if (cnt_at[li] + 1 > plans[li].resident) break;  // co-resident clusters of this size
                if (!fits(li)) break;                            // ... and of the mix
                frac += share(li) - share(li - 1);
                level[i] = li; cnt_at[li]++; cnt_at[li - 1]--;
                used += step_cost;
            }
        }
        // launches take contiguous id ranges: order by cluster size (stable: priority order within a size)
        std::vector<int> idx(n_live), ids(ctx->h_ids.begin(), ctx->h_ids.begin() + n_live), lv(level);
        for (int i = 0; i < n_live; ++i) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return lv[x] > lv[y]; });
        for (int i = 0; i < n_live; ++i) { ctx->h_ids[i] = ids[idx[i]]; level[i] = lv[idx[i]]; }
    }
    // Contribution-row cache and the L2: the chains resident at a time sweep N rows of nq floats each, cyclically
    // (row = step % N) -- with more bytes than the L2 holds, the LRU worst case.  The first keep_rows rows of every
    // chain are accessed evict_last, the others evict_first, so that what fits stays resident.
    // (only with a 1-CTA chain on every SM -- the saturated regime the rule was measured in: 3.9 -> 2.6 KB of DRAM
    // per MC step.  A slice of a single fit with fewer chains than SMs, e.g. 96 first attempts and 48 fresh backup
    // attempts, read 0.46 GB under plain LRU and 1.6 GB with the split: its chains do not sweep their rows in step.)
    a.keep_rows = ctx->args.n_contribs;
    if (ctx->tune.l2_keep_pct > 0 || (ctx->tune.l2_keep_pct == 0 && n_live >= sms_eff)) {   // (an explicit percentage: always)
        const double pct = ctx->tune.l2_keep_pct > 0 ? ctx->tune.l2_keep_pct : 50;
        const double bytes = (double)std::min(n_live, sms_eff) * ctx->args.n_contribs * (double)ctx->nq_max * sizeof(float);
        const double room = pct / 100. * (double)ctx->prop.l2CacheSize;
        if (bytes > room) a.keep_rows = (int)(ctx->args.n_contribs * room / bytes);
    }
    // A slice is over when its primaries are: backup attempts (flagged in the id list) leave their launch at the
    // next batch boundary once the count of primaries still in theirs -- d_ids[n_live], decremented by every
    // primary on its way out -- is zero.  Without it the LAST slice of a fit ran to its full budget whenever the
    // repetition that finished last had a backup attempt alive (1 ms of the fit on average).
    {
        std::vector<int> up((size_t)n_live + 2);
        int needed = 0;
        for (int i = 0; i < n_live; ++i) {
            const int slot = ctx->h_ids[i];
            const bool primary = ctx->cur[slot % ctx->n_chains] == slot;
            needed += primary ? 1 : 0;
            up[i] = primary ? slot : (slot | CHAIN_ID_BACKUP);
        }
        up[n_live] = needed;
        up[n_live + 1] = 0;                                  // CTAs whose chain has left its launch (re-pack trigger)
        CK(cudaMemcpyAsync(ctx->d_ids, up.data(), sizeof(int) * ((size_t)n_live + 2), cudaMemcpyHostToDevice, ctx->stream));
        a.needed_left = ctx->d_ids + n_live;
        // re-pack trigger: once chains holding this many CTAs have left, the others end their launch at the next
        // batch boundary, so that the freed SMs go to the chains that are left (larger clusters) at once
        a.repack_ctas = ctx->tune.repack_pct > 0 ? std::max(1, sms * ctx->tune.repack_pct / 100) : 0;
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // chains are sorted by level descending (upgrades are prefixes of the priority order)
    int pos = 0, nside = 0;
    for (int li = 3; li >= 0; --li) {
        int n = 0;
        while (pos + n < n_live && level[pos + n] == li) ++n;
        if (!n) continue;
        const bool last_group = pos + n == n_live;
        cudaStream_t st = ctx->stream;
        if (!last_group) { st = ctx->side[nside]; CK(cudaStreamWaitEvent(st, ctx->ev0, 0)); }
        int rc = launch_one_size(ctx, a, plans[li], ctx->d_ids + pos, n, st);
        if (rc) return rc;
        if (!last_group) { CK(cudaEventRecord(ctx->ev_side[nside], st)); ++nside; }
        ctx->last_sizes[li] = n;
        ctx->last_cluster = 1 << li;
        ctx->last_threads = plans[li].threads; ctx->last_qpt = CHAIN_QPT[plans[li].layout]; ctx->last_resident = plans[li].resident;
        pos += n;
    }
    for (int i = 0; i < nside; ++i) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side[i], 0));
    ctx->n_slices++;
    return MCSAS_OK;
}

// what the scheduler needs to know about every state slot after a slice: one async copy each into
// pinned memory behind the slice's launches, then the end-of-slice event
static int enqueue_readback(mcsas_ctx *ctx)
{
    const size_t NS = (size_t)ctx->n_slots;
    if (ctx->args.slice_steps > 0 || ctx->args.slice_ns > 0) {
        CK(cudaMemcpyAsync(ctx->h_att, ctx->d_att, sizeof(int) * ATT_N * NS, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(long long) * 4 * NS, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_cv, ctx->d_scal + 2 * NS, sizeof(double) * NS, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return MCSAS_OK;
}

extern "C" int mcsas_run_async(mcsas_ctx *ctx, const mcsas_run_config *cfg,
                               const double *proposal_stream, long long stream_stride,
                               volatile int *stop_flag)
{
    if (!ctx || !cfg) return MCSAS_ERR_ARG;
    if (!ctx->n_sets || !ctx->have_model) return fail(ctx, MCSAS_ERR_STATE, "run: upload datasets and set a model first");
    if (ctx->running) return fail(ctx, MCSAS_ERR_STATE, "run: previous run not waited for");
    const ModelParams &mp = ctx->mp;
    if (mp.n_active < 1) return fail(ctx, MCSAS_ERR_ARG, "run: model has no active fit parameter");
    if (cfg->n_contribs < 1 || cfg->n_reps < 1 || cfg->max_iter < 0)
        return fail(ctx, MCSAS_ERR_ARG, "run: n_contribs/n_reps/max_iter invalid");
    if (cfg->precision != MCSAS_FP64 && cfg->precision != MCSAS_FP32) return fail(ctx, MCSAS_ERR_ARG, "run: unknown precision");
    CK(cudaSetDevice(ctx->device));
    const int N = cfg->n_contribs, P = mp.n_active, D = mp.quad_nodes;
    const int n_chains = ctx->n_sets * cfg->n_reps;
    const int narr = MODEL_NARR[mp.model_id];
    const size_t real = cfg->precision == MCSAS_FP64 ? 8 : 4;
    int nq_pad_max = (ctx->nq_max + 1) & ~1;
    const size_t smem_limit = (size_t)ctx->prop.sharedMemPerBlockOptin;
    int ring = N < RING_MAX ? N : RING_MAX;
    if (cfg->precision == MCSAS_FP32) {                  // fp32 kernel indexes the ring with a mask
        ring = N < FAST_RING_MAX ? N : FAST_RING_MAX;
        int p2 = 1;
        while (p2 * 2 <= ring) p2 *= 2;
        ring = p2;
        // (the ring holds prepared proposals, not contributions: it may be larger than the set -- what a
        // batch needs is distinct rows, Kb <= N / 2 below -- so that sets of >= 257 contributions get
        // batches of FAST_KCAP)
        if (N > 2 * FAST_KCAP && ring < FAST_RING_MAX) ring = FAST_RING_MAX;
        if (cfg->max_iter >= (1LL << 31)) return fail(ctx, MCSAS_ERR_ARG, "run: max_iter must be < 2^31");
    }
    const int gw_cap = fast_gw_cap(ctx->nq_max);
    // (batches of FAST_KCAP proposals where the exchange buffers and a ring of 4 x FAST_KCAP prepared
    // proposals fit next to the chain's state; FAST_KDEF and the smaller ring otherwise)
    int xp_cap = FAST_KCAP;
    auto smem_of = [&](int rg) {
        return cfg->precision == MCSAS_FP64 ? chain_smem_bytes(nq_pad_max, N, P, D, narr, real, rg)
                                            : fast_smem_bytes(nq_pad_max, N, P, D, narr, rg, gw_cap, xp_cap);
    };
    if (cfg->precision == MCSAS_FP32 && smem_of(ring) > smem_limit) {
        xp_cap = FAST_KDEF;
        if (ring > 4 * FAST_KDEF) ring = 4 * FAST_KDEF;
    }
    while (ring > 4 && smem_of(ring) > smem_limit) ring /= 2;
    size_t smem = smem_of(ring);
    // (fp32 kernel: 255 registers x 256 threads fill an SM's register file, so 8-warp CTAs have their SM to
    // themselves; 4-warp CTAs share one in pairs and need their shared memory to fit twice)
    if (smem > smem_limit)
        return fail(ctx, MCSAS_ERR_ARG, "run: configuration needs %zu B shared memory (> %zu)", smem, smem_limit);
    const int kcap_max = cfg->precision == MCSAS_FP32 ? FAST_KCAP : MCSAS_MAX_SPEC;
    int kcap = cfg->spec_max <= 0 ? kcap_max : (cfg->spec_max > kcap_max ? kcap_max : cfg->spec_max);
    if (kcap > ring / 4) kcap = ring / 4;
    if (cfg->precision == MCSAS_FP32 && kcap > (N - 1) / 2) kcap = (N - 1) / 2;
    if (cfg->precision == MCSAS_FP32 && kcap > xp_cap) kcap = xp_cap;
    if (kcap < 1) kcap = 1;
    if (cfg->precision == MCSAS_FP32) {                  // fp32 kernel: batches are powers of two
        int p2 = 1;
        while (p2 * 2 <= kcap) p2 *= 2;
        kcap = p2;
    }

    // proposal stream (deterministic parity mode)
    const double *d_stream = nullptr;
    if (proposal_stream) {
        long long need = (cfg->start_from_minimum ? 0 : (long long)N * P);
        if (stream_stride < need) return fail(ctx, MCSAS_ERR_STREAM, "proposal stream shorter than the initial draws");
        long long steps_cover = (stream_stride - need) / P;
        if (steps_cover < cfg->max_iter)
            return fail(ctx, MCSAS_ERR_STREAM, "proposal stream covers %lld steps < max_iter %lld",
                        steps_cover, cfg->max_iter);
        size_t n = (size_t)n_chains * (size_t)stream_stride;
        if (n > ctx->stream_cap) { CK(regrow(&ctx->d_stream, n)); ctx->stream_cap = n; }
        CK(cudaMemcpyAsync(ctx->d_stream, proposal_stream, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
        d_stream = ctx->d_stream;
    }
    // outputs
    const int n_slots = 2 * n_chains;
    // every output buffer has its own capacity: a context may be reused with more q points but
    // fewer contributions (or the reverse) between fits
    if ((size_t)n_slots * N * P > ctx->rset_cap) {
        CK(regrow(&ctx->d_rset, (size_t)n_slots * N * P));
        ctx->rset_cap = (size_t)n_slots * N * P;
    }
    if ((size_t)n_slots * ctx->nq_max > ctx->fit_cap) {
        CK(regrow(&ctx->d_fit, (size_t)n_slots * ctx->nq_max));
        CK(regrow(&ctx->d_ftsave, (size_t)n_slots * ctx->nq_max));
        ctx->fit_cap = (size_t)n_slots * ctx->nq_max;
    }
    if ((size_t)n_slots > ctx->slot_cap) {
        CK(regrow(&ctx->d_scal, (size_t)n_slots * SCAL_N));
        CK(regrow(&ctx->d_cnt, (size_t)n_slots * 4));
        CK(regrow(&ctx->d_att, (size_t)n_slots * ATT_N));
        CK(regrow(&ctx->d_ids, (size_t)n_slots + 2));           // (+ the slice's counters, below)
        ctx->slot_cap = (size_t)n_slots;
    }
    if ((size_t)n_slots > ctx->pin_cap) {
        cudaFreeHost(ctx->h_att); cudaFreeHost(ctx->h_cnt); cudaFreeHost(ctx->h_cv);
        ctx->h_att = nullptr; ctx->h_cnt = nullptr; ctx->h_cv = nullptr;
        ctx->pin_cap = 0;
        CK(cudaHostAlloc((void **)&ctx->h_att, sizeof(int) * ATT_N * n_slots, cudaHostAllocDefault));
        CK(cudaHostAlloc((void **)&ctx->h_cnt, sizeof(long long) * 4 * n_slots, cudaHostAllocDefault));
        CK(cudaHostAlloc((void **)&ctx->h_cv, sizeof(double) * n_slots, cudaHostAllocDefault));
        ctx->pin_cap = (size_t)n_slots;
    }
    CK(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(long long) * (size_t)n_slots * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_scal, 0, sizeof(double) * (size_t)n_slots * SCAL_N, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_att, 0, sizeof(int) * (size_t)n_slots * ATT_N, ctx->stream));
    ctx->have_trace = false;
    if (cfg->record_trace) {
        size_t n = (size_t)n_chains * (size_t)cfg->max_iter;
        if (n > ctx->trace_cap) {
            CK(regrow(&ctx->d_tacc, n));
            CK(regrow(&ctx->d_tchi, n));
            ctx->trace_cap = n;
        }
        ctx->have_trace = true;
    }
    ChainArgs a;
    memset(&a, 0, sizeof a);
    a.mp = mp;
    a.smear = ctx->smear;
    a.n_contribs = N; a.n_reps = cfg->n_reps; a.n_chains = n_chains;
    // mcsas.py:220-246: `if nt > maxRetries: give up` is tested BEFORE each mcFit and nt counts the
    // finished ones, so a repetition makes at most maxRetries + 1 attempts (the "+ 2" of the reference's
    // log message over-counts; measured on the unmodified reference: 6 attempts for maxRetries = 5)
    a.max_attempts = proposal_stream ? 1 : (cfg->max_retries < 0 ? 1 : cfg->max_retries + 1);
    a.max_iter = cfg->max_iter; a.conv_crit = cfg->conv_crit; a.comp_exp = cfg->comp_exp;
    a.find_bg = cfg->find_background != 0; a.pos_bg = cfg->positive_background != 0;
    a.start_min = cfg->start_from_minimum != 0;
    a.spec_max = kcap; a.ring = ring;
    ctx->spec_forced = cfg->spec_max > 0;
    a.gw_cap = gw_cap;
    a.xp_cap = xp_cap;
    // fp32 sphere kernel on smeared data: per (profile point, q point) the smeared location
    // sqrt(q^2 + off^2) (slit) or q + off (2-D averaged; sasconfig.py:326-337) as the kernel wants it --
    // q in 1/nm, q/(2 pi) split hi + lo -- and weight * (q_ref/location)^6, so that the evaluation loop
    // has neither the fp64 square root nor the phase split (one 16-byte load per location instead)
    a.smloc = nullptr; a.smloc_row = 0; a.smloc_mono = 0;
    if (cfg->precision == MCSAS_FP32 && ctx->smear.n > 0 && mp.model_id == MCSAS_SPHERE) {
        const int row = gw_cap * 128, n = ctx->smear.n;
        std::vector<float4> tab((size_t)ctx->n_sets * n * row);
        size_t src = 0;
        for (int s = 0; s < ctx->n_sets; ++s) {
            const int nq = ctx->sets[s].nq;
            const double qref = (double)(float)(ctx->sets[s].q_ref * 1e-9);
            for (int k = 0; k < n; ++k)
                for (int i = 0; i < row; ++i) {
                    const double q = ctx->h_q[src + (i < nq ? i : nq - 1)], o = ctx->smear.off[k];
                    const double loc = (ctx->smear.two_d ? q + o : sqrt(q * q + o * o)) * 1e-9;
                    const double t = loc * 0.15915494309189535, ql = qref / loc, q2 = ql * ql;
                    float4 e;
                    e.x = (float)loc; e.y = (float)t; e.z = (float)(t - (double)e.y);
                    e.w = (float)(ctx->smear.tw[k] * q2 * q2 * q2);
                    tab[((size_t)s * n + k) * row + i] = e;
                }
            src += nq;
        }
        if (tab.size() > ctx->smloc_cap) { CK(regrow(&ctx->d_smloc, tab.size())); ctx->smloc_cap = tab.size(); }
        CK(cudaMemcpyAsync(ctx->d_smloc, tab.data(), sizeof(float4) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        a.smloc = ctx->d_smloc; a.smloc_row = row;
        a.smloc_mono = 1;                                    // sqrt(q^2 + off^2) >= q; q + off >= q for off >= 0
        if (ctx->smear.two_d)
            for (int k = 0; k < n; ++k) if (!(ctx->smear.off[k] >= 0.)) a.smloc_mono = 0;
    }
    a.gram = ctx->tune.gram < 0 ? 0 : 1;
    a.gram_in = ctx->tune.gram_in > 0 ? ctx->tune.gram_in : 8;
    a.gram_out = ctx->tune.gram_out > 0 ? ctx->tune.gram_out : 1;
    a.first_chain = cfg->first_chain; a.record_trace = cfg->record_trace;
    a.seed = cfg->seed;
    // Time slicing.  Deterministic (uploaded stream / traced) runs are one launch.  fp32 production
    // kernel, automatic: a launch gives every chain a TIME budget (the chain stops at the first batch
    // boundary behind it; results do not depend on where -- chain_fast.cuh), so all CTAs of a slice
    // end together whatever phase their chains are in.  More chains than twice the SMs run in waves
    // and get a long budget: the hardware's CTA scheduler balances them.
    long long slice = cfg->slice_steps;
    const bool one_launch = proposal_stream || cfg->record_trace || slice < 0;
    a.slice_steps = 0; a.slice_batches = 0; a.slice_ns = 0;
    if (!one_launch) {
        if (slice > 0) a.slice_steps = slice;                 // explicit step budget per launch
        else if (cfg->precision == MCSAS_FP32) {
            const int us = ctx->tune.slice_us > 0 ? ctx->tune.slice_us
                                                  : (n_chains > 2 * ctx->prop.multiProcessorCount ? 8000 : 2000);
            a.slice_ns = 1000LL * us;
        } else a.slice_steps = 8192;
    }
    a.sets = ctx->d_sets; a.data = ctx->d_data;
    a.quad_x = ctx->d_quad; a.quad_w = ctx->d_quad + D;
    a.stream = d_stream; a.stream_stride = stream_stride;
    a.rset = ctx->d_rset; a.fit_meas = ctx->d_fit; a.nq_stride = ctx->nq_max;
    a.ft_save = ctx->d_ftsave;
    a.scaling = ctx->d_scal; a.background = ctx->d_scal + n_slots; a.conval = ctx->d_scal + 2 * n_slots;
    a.conval0 = ctx->d_scal + 3 * n_slots; a.elapsed_s = ctx->d_scal + 4 * n_slots;
    a.cen_a = ctx->d_scal + 5 * (size_t)n_slots; a.cen_b = ctx->d_scal + 6 * (size_t)n_slots;
    a.cv_run = ctx->d_scal + 7 * (size_t)n_slots; a.last_acc = ctx->d_scal + 8 * (size_t)n_slots;
    a.num_iter = ctx->d_cnt; a.num_moves = ctx->d_cnt + n_slots; a.evals = ctx->d_cnt + 2 * n_slots;
    a.steps_all = ctx->d_cnt + 3 * (size_t)n_slots;
    a.attempts = ctx->d_att; a.status = ctx->d_att + n_slots; a.batch_k = ctx->d_att + 2 * (size_t)n_slots;
    a.trace_accept = cfg->record_trace ? ctx->d_tacc : nullptr;
    a.trace_chi = cfg->record_trace ? ctx->d_tchi : nullptr;
    a.cache = nullptr;
    if (cfg->precision == MCSAS_FP32) {
        // contribution-row cache: n_contribs rows of FAST_QCAP floats per SM (a CTA owns its SM)
        const size_t stride = (size_t)N * FAST_QCAP;
        const size_t need = stride * FAST_CACHE_SMS * FAST_CACHE_SLOTS;
        if (need > ctx->cache_cap) { CK(regrow(&ctx->d_cache, need)); ctx->cache_cap = need; }
        if (!ctx->d_cache_lock) CK(cudaMalloc(&ctx->d_cache_lock, sizeof(int) * FAST_CACHE_SMS * FAST_CACHE_SLOTS));
        CK(cudaMemsetAsync(ctx->d_cache_lock, 0, sizeof(int) * FAST_CACHE_SMS * FAST_CACHE_SLOTS, ctx->stream));
        a.cache = ctx->d_cache;
        a.cache_stride = (long long)stride;
        a.cache_lock = ctx->d_cache_lock;
    }
    if (ctx->tune.profile && cfg->precision == MCSAS_FP32) {
#ifndef MCSAS_B200_PROF
        fprintf(stderr, "[mcsas_b200] tuning.profile: this library was built without the phase counters (make PROF=1)\n");
#endif
        cudaFree(ctx->d_prof);
        ctx->d_prof = nullptr;
        CK(cudaMalloc(&ctx->d_prof, sizeof(long long) * 48 * n_slots));
        CK(cudaMemset(ctx->d_prof, 0, sizeof(long long) * 48 * n_slots));
        a.prof = ctx->d_prof;
    }
    ctx->args = a;
    ctx->precision = cfg->precision;
    ctx->smem = smem;
    ctx->forced_cluster = cfg->cluster_size;
    ctx->forced_qpt = ctx->tune.qpt;
    ctx->max_cluster = ctx->tune.max_cluster > 0 ? ctx->tune.max_cluster : 8;
    ctx->n_chains = n_chains; ctx->n_contribs = N; ctx->n_reps = cfg->n_reps; ctx->n_active = P;
    ctx->n_slots = n_slots;
    ctx->max_attempts = a.max_attempts;
    ctx->backups = cfg->precision == MCSAS_FP32 && !one_launch && a.max_attempts > 1 && ctx->tune.backups >= 0;
    ctx->cur.assign(n_chains, 0);
    for (int i = 0; i < n_chains; ++i) ctx->cur[i] = i;
    ctx->bak.assign(n_chains, -1);
    ctx->bak_state.assign(n_chains, 0);
    ctx->done.assign(n_chains, 0);
    ctx->carry_steps.assign(n_chains, 0);
    ctx->carry_evals.assign(n_chains, 0);
    ctx->n_backups_started = ctx->n_backups_used = 0;
    ctx->max_iter = cfg->max_iter;
    ctx->h_ids.resize(n_slots);
    for (int i = 0; i < n_chains; ++i) ctx->h_ids[i] = i;
    ctx->n_live = n_chains;
    ctx->h_conval.clear();
    ctx->n_launches = 0;
    ctx->n_slices = 0;
    ctx->kernel_ms = 0.f;
    ctx->stop_flag = stop_flag;
    *ctx->h_stop = 0;
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    int rc = launch_slice(ctx, n_chains, false);
    if (rc) return rc;
    rc = enqueue_readback(ctx);
    if (rc) return rc;
    ctx->running = true;
    ctx->have_run = false;
    return MCSAS_OK;
}

extern "C" int mcsas_request_stop(mcsas_ctx *ctx)
{
    if (!ctx) return MCSAS_ERR_ARG;
    *(volatile int *)ctx->h_stop = 1;
    return MCSAS_OK;
}

// Drives the remaining time slices: after each launch the chains that are still running
// are compacted and relaunched (with a larger cluster when fewer are alive).  McSAS.stop
// (mcsas.py:357) is honoured at slice boundaries.
extern "C" int mcsas_wait(mcsas_ctx *ctx, float *kernel_ms)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (!ctx->running) return fail(ctx, MCSAS_ERR_STATE, "wait: nothing running");
    ctx->running = false;
    const int NC = ctx->n_chains, NS = ctx->n_slots;
    const bool sliced = ctx->args.slice_steps > 0 || ctx->args.slice_ns > 0;
    for (;;) {
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->kernel_ms += ms;
        if (ctx->tune.verbose) {
            int nb = 0;                                      // backup attempts among the slots of this slice
            if (ctx->n_slices > 1)
                for (int i = 0; i < ctx->n_live; ++i)
                    if (ctx->cur[ctx->h_ids[i] % NC] != ctx->h_ids[i]) ++nb;
            fprintf(stderr, "[mcsas_b200] slice %d: live %d (%d backup attempts)  chains at cluster 1/2/4/8: %d/%d/%d/%d  %.3f ms\n",
                    ctx->n_slices, ctx->n_live, nb, ctx->last_sizes[0], ctx->last_sizes[1], ctx->last_sizes[2],
                    ctx->last_sizes[3], ms);
        }
        if (!sliced) break;
        int *att = ctx->h_att;
        long long *cnt = ctx->h_cnt;
        int *attempts = att, *status = att + NS;
        for (int i = 0; i < ctx->n_live; ++i) {
            int s = ctx->h_ids[i];
            if (status[s] < 0)
                return fail(ctx, MCSAS_ERR_CUDA, "chain %d: kernel watchdog fired (status %d)", s % NC, status[s]);
        }
        ctx->h_conval.assign(ctx->h_cv, ctx->h_cv + NS);
        const long long *num_iter = cnt, *evals = cnt + 2 * (size_t)NS, *steps_all = cnt + 3 * (size_t)NS;
        // ---- per chain: current attempt finished / failed / continues; backup bookkeeping --------
        int n = 0, n_primary = 0;
        for (int c = 0; c < NC; ++c) {
            if (ctx->done[c]) continue;
            const int pslot = ctx->cur[c], sp = status[pslot];
            int b = ctx->bak[c];
            if (b >= 0 && ctx->bak_state[c] == 1 && status[b] == 1) ctx->bak_state[c] = 2;   // backup attempt ended
            if (sp == 1) {                                   // converged, or out of attempts: final
                ctx->done[c] = 1;
                ctx->bak[c] = -1; ctx->bak_state[c] = 0;     // a running backup is dropped
                continue;
            }
            if (sp == 2 && b >= 0) {
                // this attempt failed and the NEXT attempt has been running in the backup slot:
                // it becomes the chain's current attempt (identical to a retry started now --
                // an attempt is a function of (seed, chain, attempt number) only)
                ctx->carry_steps[c] += steps_all[pslot];
                ctx->carry_evals[c] += evals[pslot];
                ctx->cur[c] = b;
                ctx->n_backups_used++;
                const bool finished = ctx->bak_state[c] == 2;
                ctx->bak[c] = -1; ctx->bak_state[c] = 0;
                if (finished) { ctx->done[c] = 1; continue; }
                ctx->h_ids[n++] = b;
                ++n_primary;
                continue;
            }
            ctx->h_ids[n++] = pslot;                         // 0: continue, 2: retry afresh in the same slot
            ++n_primary;
            if (b >= 0 && ctx->bak_state[c] == 1) ctx->h_ids[n++] = b;
        }
        // ---- start backups ------------------------------------------------------------------------
        // (a) few chains alive, SMs idle: every chain that is late in its attempt;
        // (b) earlier, while the machine is still full: only the few chains most likely to use all their
        //     iterations -- the largest chi-squared among those past `early_pct` % of maxIterations (from
        //     ~30 000 steps on the chi-squared rank predicts the length, tools/predict_len.py).  The second
        //     attempt of a failing repetition is what a fit waits for at the end; started this early it is
        //     finished when the first attempt gives up.
        bool new_backups = false;
        const int bdiv = ctx->tune.backup_div > 0 ? ctx->tune.backup_div : 2;
        // with at most #SMs/4 chains alive there are SMs to spare from the start (a fit partitioned over
        // several GPUs): hedge every repetition early, from `hedge_pct` % of maxIterations on
        // (measured, tools/tune.py hedge: 25 repetitions per GPU -15 %, 13 repetitions +4 %, 50 flat: off by default)
        const int hpct = ctx->tune.hedge_pct > 0 ? ctx->tune.hedge_pct : 100;
        const int bpct0 = ctx->tune.backup_pct > 0 ? ctx->tune.backup_pct : 40;
        const int bpct = n_primary * 4 <= ctx->prop.multiProcessorCount ? (hpct < bpct0 ? hpct : bpct0) : bpct0;
        const int epct = ctx->tune.early_pct > 0 ? ctx->tune.early_pct : 30;
        // a repetition whose current attempt is already a retry is a hard one (it failed before, and a fit waits for
        // the repetition that fails twice: 36 % of the config-2 fits have one): its next attempt starts at once
        // (tools/tune.py retry, 25 seeds: 29.9 -> 29.0 ms per fit, fits with a double failure -2.5 .. -4.6 ms)
        const int rpct = ctx->tune.retry_pct > 0 ? ctx->tune.retry_pct : (ctx->tune.retry_pct < 0 ? bpct : 0);
        const int btop = ctx->tune.backup_top > 0 ? ctx->tune.backup_top : 0;   // at most this many backups alive (largest chi-squared first), 0 = no bound
        const int etop = ctx->tune.early_top > 0 ? ctx->tune.early_top : 0;   // (measured on config 2: no gain at 8..24 chains from 25..35 %, off by default)
        auto start_backup = [&](int c) -> int {
            const int pslot = ctx->cur[c];
            const int b = pslot < NC ? pslot + NC : pslot - NC;
            // fresh state of the backup slot, uploaded below with the host copies
            att[b] = attempts[pslot] + 1; att[(size_t)NS + b] = 2;
            for (int k = 0; k < 4; ++k) cnt[(size_t)k * NS + b] = 0;
            ctx->h_conval[b] = 1e300;
            new_backups = true;
            cudaMemsetAsync(ctx->d_scal + 4 * (size_t)NS + b, 0, sizeof(double), ctx->stream);   // elapsed
            ctx->bak[c] = b; ctx->bak_state[c] = 1;
            ctx->h_ids[n++] = b;
            ctx->n_backups_started++;
            return b;
        };
        if (ctx->backups && n_primary > 0) {
            const bool idle = n_primary * bdiv <= ctx->prop.multiProcessorCount;
            std::vector<std::pair<double, int>> cand;
            int n_bak_alive = 0;
            for (int c = 0; c < NC; ++c) {
                if (ctx->done[c]) continue;
                if (ctx->bak[c] >= 0) { ++n_bak_alive; continue; }
                const int pslot = ctx->cur[c];
                if (status[pslot] != 0 || attempts[pslot] >= ctx->max_attempts) continue;
                const int thr = attempts[pslot] >= 1 ? rpct : bpct;   // a repetition that failed before is a hard one
                if (idle && 100 * num_iter[pslot] >= (long long)thr * ctx->max_iter) {
                    if (btop > 0) cand.push_back(std::make_pair(-ctx->h_conval[pslot], c));
                    else start_backup(c);
                } else if (!idle && etop > 0 && 100 * num_iter[pslot] >= (long long)epct * ctx->max_iter)
                    cand.push_back(std::make_pair(-ctx->h_conval[pslot], c));
            }
            if (!cand.empty()) {
                std::sort(cand.begin(), cand.end());
                const int cap = idle ? btop : etop;
                for (size_t i = 0; i < cand.size() && n_bak_alive < cap; ++i, ++n_bak_alive) start_backup(cand[i].second);
            }
        }
        if (new_backups) {                                  // nothing runs between slices: whole arrays
            CK(cudaMemcpyAsync(ctx->d_att, att, sizeof(int) * ATT_N * (size_t)NS, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_cnt, cnt, sizeof(long long) * 4 * (size_t)NS, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (ctx->n_slices > 1000000) return fail(ctx, MCSAS_ERR_CUDA, "time-slice loop did not terminate");
        ctx->n_live = n;
        bool stop = *(volatile int *)ctx->h_stop != 0 || (ctx->stop_flag && *ctx->stop_flag);
        if (n_primary == 0 || stop) break;
        int rc = launch_slice(ctx, n, true);
        if (rc) return rc;
        rc = enqueue_readback(ctx);
        if (rc) return rc;
    }
    if (ctx->tune.verbose && ctx->backups)
        fprintf(stderr, "[mcsas_b200] backup attempts started %d, used %d\n", ctx->n_backups_started, ctx->n_backups_used);
    if (ctx->args.prof) {                                 // phase profile (experiments only)
        std::vector<long long> h((size_t)48 * ctx->n_slots);
        CK(cudaMemcpy(h.data(), ctx->args.prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        static const char *nm[6] = {"ring-wait", "refresh+eval", "B1", "exchange+decide", "B2", "update"};
        for (int m = 0; m < 2; ++m) {
            long long tot[24] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = 0; c < ctx->n_slots; ++c)
                for (int i = 0; i < 24; ++i) tot[i] += h[((size_t)c * 2 + m) * 24 + i];
            if (!tot[6]) continue;
            fprintf(stderr, "[mcsas_b200 prof] %s batches %lld steps %lld (%.1f steps/batch), cycles/batch:",
                    m ? "GRAM" : "SPEC", tot[6], tot[7], (double)tot[7] / tot[6]);
            double sum = 0.;
            for (int i = 0; i < 6; ++i) { fprintf(stderr, " %s %.0f", nm[i], (double)tot[i] / tot[6]); sum += (double)tot[i] / tot[6]; }
            fprintf(stderr, " | total %.0f = %.0f cycles/step\n", sum, sum * tot[6] / tot[7]);
            fprintf(stderr, "[mcsas_b200 prof]    refresh %.0f/batch, first chunk %.0f (n %lld), later chunks %.0f each (n %lld)\n",
                    (double)tot[8] / tot[6], tot[11] ? (double)tot[9] / tot[11] : 0., tot[11],
                    tot[12] ? (double)tot[10] / tot[12] : 0., tot[12]);
            fprintf(stderr, "[mcsas_b200 prof]    decide: pre %.0f, base sums %.0f, proposal sums %.0f, solve+select %.0f, publish %.0f\n",
                    (double)tot[13] / tot[6], (double)tot[14] / tot[6], (double)tot[15] / tot[6], (double)tot[16] / tot[6],
                    (double)tot[17] / tot[6]);
        }
    }
    CK(cudaEventElapsedTime(&ctx->span_ms, ctx->ev_begin, ctx->ev1));
    if (kernel_ms) *kernel_ms = ctx->span_ms;
    ctx->have_run = true;
    return MCSAS_OK;
}

extern "C" int mcsas_run(mcsas_ctx *ctx, const mcsas_run_config *cfg, const double *proposal_stream,
                         long long stream_stride, volatile int *stop_flag)
{
    int rc = mcsas_run_async(ctx, cfg, proposal_stream, stream_stride, stop_flag);
    if (rc) return rc;
    return mcsas_wait(ctx, nullptr);
}

extern "C" int mcsas_run_info(mcsas_ctx *ctx, int *n_launches, int *last_cluster, float *kernel_ms)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (n_launches) *n_launches = ctx->n_launches;
    if (last_cluster) *last_cluster = ctx->last_cluster;
    if (kernel_ms) *kernel_ms = ctx->kernel_ms;
    return MCSAS_OK;
}

extern "C" int mcsas_set_tuning(mcsas_ctx *ctx, const mcsas_tuning *t)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (ctx->running) return fail(ctx, MCSAS_ERR_STATE, "set_tuning: a run is in flight");
    if (t) ctx->tune = *t; else memset(&ctx->tune, 0, sizeof ctx->tune);
    if (ctx->tune.max_cluster != 0 && ctx->tune.max_cluster != 1 && ctx->tune.max_cluster != 2
        && ctx->tune.max_cluster != 4 && ctx->tune.max_cluster != 8)
        return fail(ctx, MCSAS_ERR_ARG, "set_tuning: max_cluster must be 0, 1, 2, 4 or 8");
    if (ctx->tune.qpt != 0 && ctx->tune.qpt != 1 && ctx->tune.qpt != 2 && ctx->tune.qpt != 4)
        return fail(ctx, MCSAS_ERR_ARG, "set_tuning: qpt must be 0, 1, 2 or 4");
    return MCSAS_OK;
}

extern "C" int mcsas_scheduler_info(mcsas_ctx *ctx, int *backups_started, int *backups_used)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (backups_started) *backups_started = ctx->n_backups_started;
    if (backups_used) *backups_used = ctx->n_backups_used;
    return MCSAS_OK;
}

extern "C" int mcsas_fetch(mcsas_ctx *ctx, double *contribs, double *fit_meas, double *scaling,
                           double *background, double *conval, double *conval0, long long *num_iter,
                           long long *num_moves, int *attempts, double *elapsed_s, long long *evals,
                           long long *steps_all)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (!ctx->have_run) return fail(ctx, MCSAS_ERR_STATE, "fetch: no finished run");
    CK(cudaSetDevice(ctx->device));
    const int C = ctx->n_chains, N = ctx->n_contribs, P = ctx->n_active, R = ctx->n_reps, NS = ctx->n_slots;
    const std::vector<int> &fs = ctx->cur;                // slot holding each chain's final attempt
    // one pinned staging buffer, every copy asynchronous, one synchronisation
    const size_t n_rset = contribs ? (size_t)NS * N * P : 0, n_fit = fit_meas ? (size_t)NS * ctx->nq_max : 0;
    const size_t n_sc = (size_t)NS * 5, n_cnt = (size_t)NS * 4, n_att = ((size_t)NS + 1) / 2;   // (ints packed in doubles)
    const size_t need = n_rset + n_fit + n_sc + n_cnt + n_att;
    if (need > ctx->stage_cap) {
        cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->stage_cap = 0;
        CK(cudaHostAlloc((void **)&ctx->h_stage, sizeof(double) * need, cudaHostAllocDefault));
        ctx->stage_cap = need;
    }
    double *h_rset = ctx->h_stage, *h_fit = h_rset + n_rset, *sc = h_fit + n_fit;
    long long *cn = (long long *)(sc + n_sc);
    int *at = (int *)(cn + n_cnt);
    cudaStream_t st = ctx->stream;
    if (n_rset) CK(cudaMemcpyAsync(h_rset, ctx->d_rset, sizeof(double) * n_rset, cudaMemcpyDeviceToHost, st));
    if (n_fit) CK(cudaMemcpyAsync(h_fit, ctx->d_fit, sizeof(double) * n_fit, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sc, ctx->d_scal, sizeof(double) * n_sc, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cn, ctx->d_cnt, sizeof(long long) * n_cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(at, ctx->d_att, sizeof(int) * (size_t)NS, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (contribs) {
        // [slot][N][P] -> [set][N][P][R]
        for (int c = 0; c < C; ++c) {
            int s = c / R, r = c % R;
            const double *src = h_rset + (size_t)fs[c] * N * P;
            double *dst = contribs + (size_t)s * N * P * R + r;
            for (int i = 0; i < N * P; ++i) dst[(size_t)i * R] = src[i];
        }
    }
    if (fit_meas) {
        size_t base = 0;
        for (int s = 0; s < ctx->n_sets; ++s) {
            int nq = ctx->sets[s].nq;
            for (int r = 0; r < R; ++r) {
                const double *src = h_fit + (size_t)fs[s * R + r] * ctx->nq_max;
                for (int i = 0; i < nq; ++i) fit_meas[base + (size_t)i * R + r] = src[i];
            }
            base += (size_t)nq * R;
        }
    }
    for (int c = 0; c < C; ++c) {
        const size_t f = (size_t)fs[c];
        if (scaling) scaling[c] = sc[f];
        if (background) background[c] = sc[(size_t)NS + f];
        if (conval) conval[c] = sc[2 * (size_t)NS + f];
        if (conval0) conval0[c] = sc[3 * (size_t)NS + f];
        if (elapsed_s) elapsed_s[c] = sc[4 * (size_t)NS + f];
        if (num_iter) num_iter[c] = cn[f];
        if (num_moves) num_moves[c] = cn[(size_t)NS + f];
        if (evals) evals[c] = cn[2 * (size_t)NS + f] + ctx->carry_evals[c];
        if (steps_all) steps_all[c] = cn[3 * (size_t)NS + f] + ctx->carry_steps[c];
        if (attempts) attempts[c] = at[f];
    }
    return MCSAS_OK;
}

extern "C" int mcsas_fetch_trace(mcsas_ctx *ctx, int chain, unsigned char *accept, double *chi,
                                 long long capacity)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (!ctx->have_run || !ctx->have_trace) return fail(ctx, MCSAS_ERR_STATE, "fetch_trace: last run recorded no trace");
    if (chain < 0 || chain >= ctx->n_chains) return fail(ctx, MCSAS_ERR_ARG, "fetch_trace: chain out of range");
    CK(cudaSetDevice(ctx->device));
    long long nit = 0;
    CK(cudaMemcpy(&nit, ctx->d_cnt + chain, sizeof(long long), cudaMemcpyDeviceToHost));
    if (nit > capacity) nit = capacity;
    if (accept && nit) CK(cudaMemcpy(accept, ctx->d_tacc + (size_t)chain * ctx->max_iter, (size_t)nit, cudaMemcpyDeviceToHost));
    if (chi && nit) CK(cudaMemcpy(chi, ctx->d_tchi + (size_t)chain * ctx->max_iter, sizeof(double) * (size_t)nit, cudaMemcpyDeviceToHost));
    return MCSAS_OK;
}

extern "C" int mcsas_fractions(mcsas_ctx *ctx, const mcsas_run_config *cfg, const double *contribs,
                               double *vol_frac, double *num_frac, double *volsqr_frac,
                               double *surf_frac, double *min_vol, double *min_num,
                               double *min_volsqr, double *min_surf, double *totals, double *sc)
{
    if (!ctx || !cfg || !contribs) return MCSAS_ERR_ARG;
    if (!ctx->n_sets || !ctx->have_model) return fail(ctx, MCSAS_ERR_STATE, "fractions: upload datasets and set a model first");
    CK(cudaSetDevice(ctx->device));
    const int N = cfg->n_contribs, R = cfg->n_reps, P = ctx->mp.n_active, D = ctx->mp.quad_nodes;
    const int C = ctx->n_sets * R;
    std::vector<double> h((size_t)C * N * P);
    for (int c = 0; c < C; ++c) {                   // [set][N][P][R] -> [chain][N][P]
        int s = c / R, r = c % R;
        for (int i = 0; i < N * P; ++i) h[(size_t)c * N * P + i] = contribs[((size_t)s * N * P + i) * R + r];
    }
    double *d_r = nullptr, *d_o = nullptr;
    size_t nout = (size_t)C * N * 8 + (size_t)C * 6;
    CK(cudaMalloc(&d_r, sizeof(double) * h.size()));
    CK(cudaMalloc(&d_o, sizeof(double) * nout));
    CK(cudaMemcpyAsync(d_r, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    FractionArgs a;
    a.mp = ctx->mp; a.smear = ctx->smear; a.sets = ctx->d_sets; a.data = ctx->d_data;
    a.quad_x = ctx->d_quad; a.quad_w = ctx->d_quad + D;
    a.n_contribs = N; a.n_reps = R; a.n_chains = C; a.comp_exp = cfg->comp_exp;
    a.find_bg = cfg->find_background != 0; a.pos_bg = cfg->positive_background != 0;
    a.rset = d_r; a.out = d_o; a.totals = d_o + (size_t)C * N * 8; a.sc = a.totals + (size_t)C * 4;
    size_t real = cfg->precision == MCSAS_FP64 ? 8 : 4;
    size_t smem = real * (2 * (D + 1) + (size_t)MODEL_NARR[ctx->mp.model_id] * D) + 8 * (size_t)ctx->nq_max + 64 * 8 + 64;
    cudaError_t e = mcsas_launch_fractions(cfg->precision, ctx->mp.model_id, C, smem, ctx->stream, a);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    std::vector<double> o(nout);
    if (e == cudaSuccess) e = cudaMemcpy(o.data(), d_o, sizeof(double) * nout, cudaMemcpyDeviceToHost);
    cudaFree(d_r); cudaFree(d_o);
    if (e != cudaSuccess) return fail(ctx, MCSAS_ERR_CUDA, "fractions kernel: %s", cudaGetErrorString(e));
    double *outs[8] = {vol_frac, num_frac, volsqr_frac, surf_frac, min_vol, min_num, min_volsqr, min_surf};
    for (int k = 0; k < 8; ++k) {
        if (!outs[k]) continue;
        for (int c = 0; c < C; ++c) {               // [k][chain][N] -> [set][N][R]
            int s = c / R, r = c % R;
            for (int i = 0; i < N; ++i)
                outs[k][((size_t)s * N + i) * R + r] = o[((size_t)k * C + c) * N + i];
        }
    }
    const double *tot = o.data() + (size_t)C * N * 8, *scv = tot + (size_t)C * 4;
    for (int c = 0; c < C; ++c) {
        int s = c / R, r = c % R;
        if (totals) for (int k = 0; k < 4; ++k) totals[((size_t)s * 4 + k) * R + r] = tot[(size_t)c * 4 + k];
        if (sc) for (int k = 0; k < 2; ++k) sc[((size_t)s * 2 + k) * R + r] = scv[(size_t)c * 2 + k];
    }
    return MCSAS_OK;
}

// Data preparation of a batch of raw tables on the device (dataprep.cu); host buffers in and out.
extern "C" int mcsas_prepare_datasets(mcsas_ctx *ctx, int n_sets, const int *n_raw, const double *q,
                                      const double *intensity, const double *sigma, double fu_min, double x0_low,
                                      double x0_high, int mask_zero, int mask_neg, int n_bin, const double *edges,
                                      int cap, int *nq_out, double *q_out, double *i_out, double *sigma_out)
{
    if (!ctx) return MCSAS_ERR_ARG;
    if (n_sets < 1 || !n_raw || !q || !intensity || !nq_out || !q_out || !i_out || !sigma_out || cap < 1)
        return fail(ctx, MCSAS_ERR_ARG, "prepare_datasets: null/empty argument");
    if (n_bin > 0 && !edges) return fail(ctx, MCSAS_ERR_ARG, "prepare_datasets: n_bin > 0 needs the bin edges");
    if (n_bin > cap) return fail(ctx, MCSAS_ERR_ARG, "prepare_datasets: capacity %d < n_bin %d", cap, n_bin);
    CK(cudaSetDevice(ctx->device));
    std::vector<long long> off((size_t)n_sets + 1, 0);
    for (int s = 0; s < n_sets; ++s) {
        if (n_raw[s] < 0) return fail(ctx, MCSAS_ERR_ARG, "prepare_datasets: negative row count");
        if (n_bin <= 0 && n_raw[s] > cap) return fail(ctx, MCSAS_ERR_ARG, "prepare_datasets: capacity %d < %d rows", cap, n_raw[s]);
        off[s + 1] = off[s] + n_raw[s];
    }
    const size_t rows = (size_t)off[n_sets], nout = (size_t)n_sets * cap;
    double *d_in = nullptr, *d_edges = nullptr, *d_out = nullptr;
    long long *d_off = nullptr;
    int *d_scr = nullptr, *d_nq = nullptr;
    const int ncol = sigma ? 3 : 2;
    cudaError_t e = cudaMalloc(&d_in, sizeof(double) * (rows ? rows : 1) * ncol);
    if (e == cudaSuccess) e = cudaMalloc(&d_off, sizeof(long long) * (n_sets + 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_scr, sizeof(int) * (rows ? rows : 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_nq, sizeof(int) * n_sets);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * 3 * nout);
    if (e == cudaSuccess && n_bin > 0) e = cudaMalloc(&d_edges, sizeof(double) * (size_t)n_sets * (n_bin + 1));
    cudaStream_t st = ctx->stream;
    if (e == cudaSuccess && rows) {
        e = cudaMemcpyAsync(d_in, q, sizeof(double) * rows, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_in + rows, intensity, sizeof(double) * rows, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && sigma) e = cudaMemcpyAsync(d_in + 2 * rows, sigma, sizeof(double) * rows, cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, off.data(), sizeof(long long) * (n_sets + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_bin > 0)
        e = cudaMemcpyAsync(d_edges, edges, sizeof(double) * (size_t)n_sets * (n_bin + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        PrepArgs a;
        a.n_sets = n_sets; a.n_bin = n_bin; a.cap = cap; a.has_sigma = sigma != nullptr;
        a.raw_off = d_off; a.q = d_in; a.I = d_in + rows; a.sigma = sigma ? d_in + 2 * rows : nullptr;
        a.edges = d_edges; a.fu_min = fu_min; a.x0_low = x0_low; a.x0_high = x0_high;
        a.mask_zero = mask_zero != 0; a.mask_neg = mask_neg != 0;
        a.scratch = d_scr; a.nq_out = d_nq;
        a.q_out = d_out; a.I_out = d_out + nout; a.s_out = d_out + 2 * nout;
        e = mcsas_launch_prepare(a, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(nq_out, d_nq, sizeof(int) * n_sets, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_out, d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(i_out, d_out + nout, sizeof(double) * nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sigma_out, d_out + 2 * nout, sizeof(double) * nout, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_in); cudaFree(d_off); cudaFree(d_scr); cudaFree(d_nq); cudaFree(d_out); cudaFree(d_edges);
    if (e != cudaSuccess) return fail(ctx, MCSAS_ERR_CUDA, "prepare_datasets: %s", cudaGetErrorString(e));
    return MCSAS_OK;
}

extern "C" int mcsas_measure_peaks(mcsas_ctx *ctx, double *mufu_gops, double *ffma_tflops)
{
    if (!ctx) return MCSAS_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    float *d_out = nullptr;
    const int threads = 256, grid = ctx->prop.multiProcessorCount * 8, iters = 4096;
    CK(cudaMalloc(&d_out, sizeof(float) * (size_t)grid * threads));
    double res[2] = {0., 0.};
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(ctx->ev0, ctx->stream));
            cudaError_t e = mcsas_launch_peaks(which, grid, threads, iters, d_out, ctx->stream);
            if (e != cudaSuccess) { cudaFree(d_out); return fail(ctx, MCSAS_ERR_CUDA, "peaks kernel: %s", cudaGetErrorString(e)); }
            CK(cudaEventRecord(ctx->ev1, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
            if (rep > 0 && ms < best) best = ms;
        }
        // per thread and iteration: which 0 -> 24 MUFU ops (8 x sin,cos,rcp); 1 -> 64 FFMA = 128 flop
        double ops = (double)grid * threads * iters * (which == 0 ? 24. : 128.);
        res[which] = ops / (best * 1e-3);
    }
    cudaFree(d_out);
    if (mufu_gops) *mufu_gops = res[0] * 1e-9;
    if (ffma_tflops) *ffma_tflops = res[1] * 1e-12;
    return MCSAS_OK;
}
