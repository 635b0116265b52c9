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
    cudaStream_t stream = nullptr, stream2 = nullptr;     // stream2: the larger-cluster half of a mixed slice
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_begin = nullptr, ev_mix = nullptr;
    float span_ms = 0.f;          // first launch start -> last slice end (host gaps between slices included)
    cudaDeviceProp prop;
    std::string err;
    // datasets
    int n_sets = 0;
    int nq_max = 0;
    std::vector<DatasetDesc> sets;
    DatasetDesc *d_sets = nullptr;
    double *d_data = nullptr;
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
    double *d_rset = nullptr, *d_fit = nullptr, *d_scal = nullptr;   // scal: 5 doubles per chain
    long long *d_cnt = nullptr;                                       // 3 per chain
    int *d_att = nullptr;
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
    int n_live = 0, n_launches = 0, last_up = 0, last_cluster = 1, last_threads = 0, last_qpt = 0, last_resident = 0;
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

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(ctx, MCSAS_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                 \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                         \
    } while (0)

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
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess
        || cudaEventCreate(&ctx->ev_begin) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess
        || cudaEventCreateWithFlags(&ctx->ev_mix, cudaEventDisableTiming) != cudaSuccess
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
    cudaFree(ctx->d_att); cudaFree(ctx->d_tacc); cudaFree(ctx->d_tchi); cudaFree(ctx->d_ids); cudaFree(ctx->d_prof); cudaFree(ctx->d_cache); cudaFree(ctx->d_cache_lock);
    cudaFreeHost(ctx->h_stop);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev_begin); cudaEventDestroy(ctx->ev_mix);
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->stream2);
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
    cudaFree(ctx->d_sets); cudaFree(ctx->d_data);
    ctx->d_sets = nullptr; ctx->d_data = nullptr;
    CK(cudaMalloc(&ctx->d_sets, sizeof(DatasetDesc) * n_sets));
    CK(cudaMalloc(&ctx->d_data, sizeof(double) * packed.size()));
    CK(cudaMemcpyAsync(ctx->d_sets, sets.data(), sizeof(DatasetDesc) * n_sets, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_data, packed.data(), sizeof(double) * packed.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->sets = sets;
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
    cudaFree(ctx->d_quad);
    ctx->d_quad = nullptr;
    CK(cudaMalloc(&ctx->d_quad, sizeof(double) * 2 * (D + 1)));
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

template <class T> static cudaError_t regrow(T **p, size_t n)
{
    cudaFree(*p);
    *p = nullptr;
    return cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
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
    int forced_c = ctx->forced_cluster, forced_t = 0, forced_q = 0;
    const char *env;
    if ((env = getenv("MCSAS_B200_CLUSTER")) && atoi(env) > 0) forced_c = atoi(env);
    if ((env = getenv("MCSAS_B200_T")) && atoi(env) > 0) forced_t = atoi(env);
    if ((env = getenv("MCSAS_B200_QPT")) && atoi(env) > 0) forced_q = atoi(env);
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

// Layout of the fp32 production kernel: cluster size C (q split) x q points per thread
// (1,2,4) x proposal groups PS (1,2,4); eval threads = warps_q * PS * 32.
// Measured on B200 (tools/tune.py fast, nq = 1000): 8 eval warps with 4 q points per thread
// are the sweet spot at every cluster size; a cluster pays for its per-batch exchange only
// when its CTAs get SMs of their own, so C = largest of {1,2,4} with n_live * C <= #SMs.
static int plan_chain_fp32(mcsas_ctx *ctx, int nq, int n_live, ChainPlan *out)
{
    const int sms = ctx->prop.multiProcessorCount;
    int forced_c = ctx->forced_cluster, forced_q = 0, forced_ps = 0;
    const char *env;
    if ((env = getenv("MCSAS_B200_CLUSTER")) && atoi(env) > 0) forced_c = atoi(env);
    if ((env = getenv("MCSAS_B200_QPT")) && atoi(env) > 0) forced_q = atoi(env);
    if ((env = getenv("MCSAS_B200_PS")) && atoi(env) > 0) forced_ps = atoi(env);
    if (forced_c > MAX_CLUSTER) forced_c = MAX_CLUSTER;
    ChainPlan best = {0, 0, 0, 0, 0};
    // preference order of cluster sizes: the planned one first, then smaller, then larger
    int want_c = 1, max_c = 4;
    if ((env = getenv("MCSAS_B200_MAXC")) && atoi(env) > 0) max_c = atoi(env);
    while (want_c < max_c && (long long)n_live * want_c * 2 <= sms) want_c *= 2;
    int order[8], no = 0;
    if (forced_c > 0) order[no++] = forced_c;
    else {
        for (int c = want_c; c >= 1; c /= 2) order[no++] = c;
        for (int c = want_c * 2; c <= MAX_CLUSTER; c *= 2) order[no++] = c;
    }
    const int grp_of[3] = {4, 8, 8};                        // proposals per chunk of layout QPT = 1,2,4
    for (int oi = 0; oi < no && best.cluster == 0; ++oi) {
        const int c = order[oi];
        const int n = (nq + c - 1) / c;
        for (int l = N_CHAIN_LAYOUTS - 1; l >= 0 && best.cluster == 0; --l) {       // QPT = 4 first
            const int qpt = CHAIN_QPT[l];
            if (forced_q > 0 && qpt != forced_q) continue;
            const int wq = (n + 32 * qpt - 1) / (32 * qpt);
            if (wq * 32 > FAST_MAX_EVAL_THREADS[l]) continue;
            int ps = 1;
            while (ps < 4 && wq * ps * 2 <= 8 && wq * ps * 2 * 32 <= FAST_MAX_EVAL_THREADS[l]
                   && ps * 2 * grp_of[l] <= ctx->args.spec_max) ps *= 2;
            if (forced_ps > 0) ps = forced_ps;
            const int threads = wq * ps * 32;
            if (threads > FAST_MAX_EVAL_THREADS[l] || threads / 32 > FAST_MAX_WARPS) continue;
            int resident = 0;
            query_resident(ctx, c, l, threads, &resident);
            if (resident < 1) continue;                                            // more chains than fit: waves
            best.cluster = c; best.layout = l; best.threads = threads; best.resident = resident; best.warps_q = wq;
        }
    }
    if (best.cluster == 0) return fail(ctx, MCSAS_ERR_ARG, "no launchable fp32 chain layout for nq=%d", nq);
    *out = best;
    return MCSAS_OK;
}

static int plan_chain(mcsas_ctx *ctx, int nq, int n_live, ChainPlan *out)
{
    return ctx->precision == MCSAS_FP64 ? plan_chain_fp64(ctx, nq, n_live, out)
                                        : plan_chain_fp32(ctx, nq, n_live, out);
}

// Launch one time slice for the chains listed in ctx->h_ids[0..n_live).
//
// Mixed slices (fp32 production kernel, automatic layout): when the planned cluster size C
// leaves SMs idle (config 2: 100 chains x 1 CTA on 148 SMs), n_up = idle / C chains run this
// slice as 2C-CTA clusters in a second, concurrent launch: those with the largest chi-squared.  Their slice is 11/8 as many batches (a 2C cluster runs a batch ~1.4x faster,
// tools/tune.py), so both launches end together.
static int launch_slice(mcsas_ctx *ctx, int n_live, bool resume)
{
    ChainArgs a = ctx->args;
    ChainPlan pl;
    int prc = plan_chain(ctx, ctx->nq_max, n_live, &pl);
    if (prc) return prc;
    a.resume = resume ? 1 : 0;
    int n_up = 0, up_f = 2;
    ChainPlan pu = pl;
    const char *mixed_env = getenv("MCSAS_B200_MIXED");
    const bool mixed_on = mixed_env ? atoi(mixed_env) != 0 : true;
    if (mixed_on && ctx->precision == MCSAS_FP32 && ctx->forced_cluster == 0 && !getenv("MCSAS_B200_CLUSTER")
        && a.slice_batches > 0 && pl.cluster < 4 && n_live <= pl.resident) {
        const int spare = ctx->prop.multiProcessorCount - n_live * pl.cluster;
        // up_f: the upgraded chains run on up_f * C CTAs (4: fewer chains, but the few that bound
        // the time to solution go twice as fast; only from C = 1)
        if (const char *fe = getenv("MCSAS_B200_MIXED_FACTOR")) up_f = atoi(fe) == 4 ? 4 : 2;
        if (pl.cluster * up_f > 4) up_f = 2;
        n_up = spare / ((up_f - 1) * pl.cluster);
        if (n_up > n_live) n_up = n_live;
        if (n_up >= 1) {
            ctx->forced_cluster = up_f * pl.cluster;
            prc = plan_chain(ctx, ctx->nq_max, n_up, &pu);
            ctx->forced_cluster = 0;
            if (prc || pu.cluster != up_f * pl.cluster || pu.resident < n_up) n_up = 0;
        } else n_up = 0;
    }
    ctx->last_threads = pl.threads; ctx->last_qpt = CHAIN_QPT[pl.layout]; ctx->last_resident = pl.resident;
    a.chain_ids = nullptr;
    if (n_up > 0 && n_up < n_live && (int)ctx->h_conval.size() == ctx->n_slots) {
        // the chains that are furthest from the criterion get the larger clusters: they are the
        // likely critical path of the fit (deterministic: a function of the chains' states only)
        const std::vector<double> &cv = ctx->h_conval;
        std::stable_sort(ctx->h_ids.begin(), ctx->h_ids.begin() + n_live,
                         [&cv](int x, int y) { return cv[x] > cv[y]; });
    }
    if (n_live != ctx->n_chains || resume || n_up > 0) {
        CK(cudaMemcpyAsync(ctx->d_ids, ctx->h_ids.data(), sizeof(int) * n_live, cudaMemcpyHostToDevice, ctx->stream));
        a.chain_ids = ctx->d_ids;
    }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int extra = ctx->precision == MCSAS_FP64 ? 32 : 0;                          // fp64: + producer warp
    if (n_up > 0) {
        // the clusters go first, on the main stream, so that they are placed while the SMs are
        // still empty; the single-size CTAs follow on stream2 (behind ev0 only) and fill the rest
        ChainArgs u = a;
        u.cluster = pu.cluster;
        u.warps_q = pu.warps_q;
        int num8 = up_f == 4 ? 16 : 11;                      // batches of the upgraded chains, in eighths
        if (const char *ne = getenv("MCSAS_B200_MIXED_EIGHTHS")) { if (atoi(ne) >= 8) num8 = atoi(ne); }
        u.slice_batches = a.slice_batches * num8 / 8;
        u.slice_steps = a.slice_steps * num8 / 8;
        cudaError_t e = CHAIN_FN[a.mp.model_id](ctx->precision, pu.layout, n_up * pu.cluster, pu.threads + extra,
                                                ctx->smem, ctx->stream, u);
        if (e != cudaSuccess) return fail(ctx, MCSAS_ERR_CUDA, "chain kernel launch (mixed slice): %s", cudaGetErrorString(e));
        ctx->n_launches++;
        a.chain_ids = ctx->d_ids + n_up;
    }
    if (n_live - n_up > 0) {
        cudaStream_t st = ctx->stream;
        if (n_up > 0) { st = ctx->stream2; CK(cudaStreamWaitEvent(st, ctx->ev0, 0)); }
        a.cluster = pl.cluster;
        a.warps_q = pl.warps_q;
        cudaError_t e = CHAIN_FN[a.mp.model_id](ctx->precision, pl.layout, (n_live - n_up) * pl.cluster,
                                                pl.threads + extra, ctx->smem, st, a);
        if (e != cudaSuccess) return fail(ctx, MCSAS_ERR_CUDA, "chain kernel launch: %s", cudaGetErrorString(e));
        ctx->n_launches++;
        if (n_up > 0) {
            CK(cudaEventRecord(ctx->ev_mix, st));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_mix, 0));
        }
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->last_cluster = n_live - n_up > 0 ? pl.cluster : pu.cluster;
    ctx->last_up = n_up;
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
        if (cfg->max_iter >= (1LL << 31)) return fail(ctx, MCSAS_ERR_ARG, "run: max_iter must be < 2^31");
    }
    auto smem_of = [&](int rg) {
        return cfg->precision == MCSAS_FP64 ? chain_smem_bytes(nq_pad_max, N, P, D, narr, real, rg)
                                            : fast_smem_bytes(nq_pad_max, N, P, D, narr, rg);
    };
    while (ring > 4 && smem_of(ring) > smem_limit) ring /= 2;
    size_t smem = smem_of(ring);
    // fp32 kernel: one CTA per SM (its contribution-row cache region is indexed by the SM id):
    // ask for more than half of an SM's shared memory
    if (cfg->precision == MCSAS_FP32 && smem < (size_t)120 * 1024 && smem_limit >= (size_t)120 * 1024)
        smem = (size_t)120 * 1024;
    if (smem > smem_limit)
        return fail(ctx, MCSAS_ERR_ARG, "run: configuration needs %zu B shared memory (> %zu)", smem, smem_limit);
    const int kcap_max = cfg->precision == MCSAS_FP32 ? FAST_KCAP : MCSAS_MAX_SPEC;
    int kcap = cfg->spec_max <= 0 ? kcap_max : (cfg->spec_max > kcap_max ? kcap_max : cfg->spec_max);
    if (const char *ke = getenv("MCSAS_B200_KMAX")) { if (atoi(ke) > 0 && atoi(ke) < kcap) kcap = atoi(ke); }
    if (kcap > ring / 4) kcap = ring / 4;
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
        ctx->fit_cap = (size_t)n_slots * ctx->nq_max;
    }
    if ((size_t)n_slots > ctx->slot_cap) {
        CK(regrow(&ctx->d_scal, (size_t)n_slots * 5));
        CK(regrow(&ctx->d_cnt, (size_t)n_slots * 4));
        CK(regrow(&ctx->d_att, (size_t)n_slots * 2));
        CK(regrow(&ctx->d_ids, (size_t)n_slots));
        ctx->slot_cap = (size_t)n_slots;
    }
    CK(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(long long) * (size_t)n_slots * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_scal, 0, sizeof(double) * (size_t)n_slots * 5, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_att, 0, sizeof(int) * (size_t)n_slots * 2, ctx->stream));
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
    a.max_attempts = proposal_stream ? 1 : (cfg->max_retries < 0 ? 1 : cfg->max_retries + 2);
    a.max_iter = cfg->max_iter; a.conv_crit = cfg->conv_crit; a.comp_exp = cfg->comp_exp;
    a.find_bg = cfg->find_background != 0; a.pos_bg = cfg->positive_background != 0;
    a.start_min = cfg->start_from_minimum != 0;
    a.spec_max = kcap; a.ring = ring;
    {
        const char *ge = getenv("MCSAS_B200_GRAM");
        a.gram = ge ? atoi(ge) != 0 : 1;
    }
    a.first_chain = cfg->first_chain; a.record_trace = cfg->record_trace;
    a.seed = cfg->seed;
    // time slicing: deterministic (uploaded stream / traced) runs are one launch
    long long slice = cfg->slice_steps;
    if (slice == 0) slice = (proposal_stream || cfg->record_trace) ? -1 : 8192;
    const char *env = getenv("MCSAS_B200_SLICE");
    if (env && !(proposal_stream || cfg->record_trace)) slice = atoll(env);
    a.slice_steps = slice > 0 ? slice : 0;
    a.slice_batches = 0;
    if (cfg->precision == MCSAS_FP32 && cfg->slice_steps == 0 && !env && a.slice_steps > 0) {
        // auto: a slice is a number of BATCHES (similar duration whether a chain is in its
        // move-rich early phase or late), the step limit stays as an upper bound
        a.slice_batches = 256;                              // (tools/tune.py sched: 256 with backups from <= #SMs/4 live chains)
        if (const char *sb = getenv("MCSAS_B200_SLICE_BATCHES")) { if (atoi(sb) > 0) a.slice_batches = atoi(sb); }
        a.slice_steps = 16384;
    }
    a.sets = ctx->d_sets; a.data = ctx->d_data;
    a.quad_x = ctx->d_quad; a.quad_w = ctx->d_quad + D;
    a.stream = d_stream; a.stream_stride = stream_stride;
    a.rset = ctx->d_rset; a.fit_meas = ctx->d_fit; a.nq_stride = ctx->nq_max;
    a.scaling = ctx->d_scal; a.background = ctx->d_scal + n_slots; a.conval = ctx->d_scal + 2 * n_slots;
    a.conval0 = ctx->d_scal + 3 * n_slots; a.elapsed_s = ctx->d_scal + 4 * n_slots;
    a.num_iter = ctx->d_cnt; a.num_moves = ctx->d_cnt + n_slots; a.evals = ctx->d_cnt + 2 * n_slots;
    a.steps_all = ctx->d_cnt + 3 * (size_t)n_slots;
    a.attempts = ctx->d_att; a.status = ctx->d_att + n_slots;
    a.trace_accept = cfg->record_trace ? ctx->d_tacc : nullptr;
    a.trace_chi = cfg->record_trace ? ctx->d_tchi : nullptr;
    a.cache = nullptr;
    if (cfg->precision == MCSAS_FP32 && !(getenv("MCSAS_B200_CACHE") && atoi(getenv("MCSAS_B200_CACHE")) == 0)) {
        // contribution-row cache: n_contribs rows of FAST_QCAP floats per SM (a CTA owns its SM)
        const size_t stride = (size_t)N * FAST_QCAP;
        const size_t need = stride * FAST_CACHE_SMS;
        if (need > ctx->cache_cap) { CK(regrow(&ctx->d_cache, need)); ctx->cache_cap = need; }
        if (!ctx->d_cache_lock) CK(cudaMalloc(&ctx->d_cache_lock, sizeof(int) * FAST_CACHE_SMS));
        CK(cudaMemsetAsync(ctx->d_cache_lock, 0, sizeof(int) * FAST_CACHE_SMS, ctx->stream));
        a.cache = ctx->d_cache;
        a.cache_stride = (long long)stride;
        a.cache_lock = ctx->d_cache_lock;
    }
    if (getenv("MCSAS_B200_PROF") && cfg->precision == MCSAS_FP32) {
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
    ctx->n_chains = n_chains; ctx->n_contribs = N; ctx->n_reps = cfg->n_reps; ctx->n_active = P;
    ctx->n_slots = n_slots;
    ctx->max_attempts = a.max_attempts;
    {
        const char *be = getenv("MCSAS_B200_BACKUP");
        ctx->backups = cfg->precision == MCSAS_FP32 && !proposal_stream && !cfg->record_trace && a.slice_steps > 0
                       && a.max_attempts > 1 && !(be && atoi(be) == 0);
    }
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
    ctx->kernel_ms = 0.f;
    ctx->stop_flag = stop_flag;
    *ctx->h_stop = 0;
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    int rc = launch_slice(ctx, n_chains, false);
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
    std::vector<int> att(2 * (size_t)NS);                 // attempts | status
    std::vector<long long> cnt(4 * (size_t)NS);           // num_iter | num_moves | evals | steps_all
    for (;;) {
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->kernel_ms += ms;
        if (getenv("MCSAS_B200_DEBUG"))
            fprintf(stderr, "[mcsas_b200] slice %d: live %d cluster %d (%d chains at twice that) T %d qpt %d resident %d  %.3f ms\n",
                    ctx->n_launches, ctx->n_live, ctx->last_cluster, ctx->last_up, ctx->last_threads, ctx->last_qpt,
                    ctx->last_resident, ms);
        if (ctx->args.slice_steps <= 0) break;
        CK(cudaMemcpy(att.data(), ctx->d_att, sizeof(int) * att.size(), cudaMemcpyDeviceToHost));
        int *attempts = att.data(), *status = att.data() + NS;
        for (int i = 0; i < ctx->n_live; ++i) {
            int s = ctx->h_ids[i];
            if (status[s] < 0)
                return fail(ctx, MCSAS_ERR_CUDA, "chain %d: kernel watchdog fired (status %d)", s % NC, status[s]);
        }
        if (ctx->backups) CK(cudaMemcpy(cnt.data(), ctx->d_cnt, sizeof(long long) * cnt.size(), cudaMemcpyDeviceToHost));
        if (ctx->precision == MCSAS_FP32 && ctx->args.slice_batches > 0) {
            ctx->h_conval.resize(NS);
            CK(cudaMemcpy(ctx->h_conval.data(), ctx->d_scal + 2 * (size_t)NS, sizeof(double) * NS, cudaMemcpyDeviceToHost));
        }
        const long long *num_iter = cnt.data(), *evals = cnt.data() + 2 * (size_t)NS, *steps_all = cnt.data() + 3 * (size_t)NS;
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
        // ---- start backups: few chains alive, SMs idle, chain is late in its attempt -------------
        bool new_backups = false;
        int bdiv = 4;
        if (const char *bd = getenv("MCSAS_B200_BACKUP_DIV")) bdiv = atoi(bd) > 0 ? atoi(bd) : 4;
        if (ctx->backups && n_primary > 0 && n_primary * bdiv <= ctx->prop.multiProcessorCount) {
            for (int c = 0; c < NC; ++c) {
                if (ctx->done[c] || ctx->bak[c] >= 0) continue;
                const int pslot = ctx->cur[c];
                if (status[pslot] != 0 || attempts[pslot] >= ctx->max_attempts) continue;
                if (2 * num_iter[pslot] < ctx->max_iter) continue;
                const int b = pslot < NC ? pslot + NC : pslot - NC;
                // fresh state of the backup slot, uploaded below with the host copies
                att[b] = attempts[pslot] + 1; att[(size_t)NS + b] = 2;
                for (int k = 0; k < 4; ++k) cnt[(size_t)k * NS + b] = 0;
                new_backups = true;
                CK(cudaMemsetAsync(ctx->d_scal + 4 * (size_t)NS + b, 0, sizeof(double), ctx->stream));   // elapsed
                ctx->bak[c] = b; ctx->bak_state[c] = 1;
                ctx->h_ids[n++] = b;
                ctx->n_backups_started++;
            }
        }
        if (new_backups) {                                  // nothing runs between slices: whole arrays
            CK(cudaMemcpyAsync(ctx->d_att, att.data(), sizeof(int) * att.size(), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_cnt, cnt.data(), sizeof(long long) * cnt.size(), cudaMemcpyHostToDevice, ctx->stream));
        }
        if (ctx->n_launches > 100000) return fail(ctx, MCSAS_ERR_CUDA, "time-slice loop did not terminate");
        ctx->n_live = n;
        bool stop = *(volatile int *)ctx->h_stop != 0 || (ctx->stop_flag && *ctx->stop_flag);
        if (n_primary == 0 || stop) break;
        int rc = launch_slice(ctx, n, true);
        if (rc) return rc;
    }
    if (getenv("MCSAS_B200_DEBUG") && ctx->backups)
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
    if (contribs) {
        std::vector<double> h((size_t)NS * N * P);
        CK(cudaMemcpy(h.data(), ctx->d_rset, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        // [slot][N][P] -> [set][N][P][R]
        for (int c = 0; c < C; ++c) {
            int s = c / R, r = c % R;
            for (int i = 0; i < N * P; ++i)
                contribs[((size_t)s * N * P + i) * R + r] = h[(size_t)fs[c] * N * P + i];
        }
    }
    if (fit_meas) {
        std::vector<double> h((size_t)NS * ctx->nq_max);
        CK(cudaMemcpy(h.data(), ctx->d_fit, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
        size_t base = 0;
        for (int s = 0; s < ctx->n_sets; ++s) {
            int nq = ctx->sets[s].nq;
            for (int r = 0; r < R; ++r)
                for (int i = 0; i < nq; ++i)
                    fit_meas[base + (size_t)i * R + r] = h[(size_t)fs[s * R + r] * ctx->nq_max + i];
            base += (size_t)nq * R;
        }
    }
    std::vector<double> sc((size_t)NS * 5);
    CK(cudaMemcpy(sc.data(), ctx->d_scal, sizeof(double) * sc.size(), cudaMemcpyDeviceToHost));
    std::vector<long long> cn((size_t)NS * 4);
    CK(cudaMemcpy(cn.data(), ctx->d_cnt, sizeof(long long) * cn.size(), cudaMemcpyDeviceToHost));
    std::vector<int> at((size_t)NS);
    CK(cudaMemcpy(at.data(), ctx->d_att, sizeof(int) * at.size(), cudaMemcpyDeviceToHost));
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
