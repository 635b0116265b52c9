// aux_kernels.cu -- post-fit fractions/observability kernel (McSAS.histogram per
// repetition, mcsas.py:548-604) and the pipe-peak microbenchmarks used as roofline
// denominators (MUFU and FP32-FMA issue rates of this GPU).
#include "chain_kernel.cuh"
#include "launch.cuh"

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

#define FR_GROUP 8
#define FR_THREADS 256

// One CTA per (dataset, repetition).  Per contribution c: I_c(q) -> ft += I_c and
// min_q sigma/I_c (observability, mcsas.py:575-594); then scale/background (:559) and
// the four fraction kinds (:565-572) with their normalisation (:596-604).
template <int MODEL, class Real>
__global__ void __launch_bounds__(FR_THREADS) fractions_kernel(const __grid_constant__ FractionArgs a)
{
    using O = Ops<MODEL, Real>;
    constexpr int NARR = O::NARR;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int chain = blockIdx.x, set = chain / a.n_reps;
    const DatasetDesc ds = a.sets[set];
    const int nq = ds.nq, N = a.n_contribs, P = a.mp.n_active;
    const int D = NARR ? a.mp.quad_nodes : 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_ft = (double *)smem_raw;                   // [nq_max]
    double *s_red = s_ft + nq + (nq & 1);                // [8 warps][FR_GROUP] + scratch
    Real *s_qx = (Real *)(s_red + 64);
    Real *s_qw = s_qx + D + 1;
    Real *s_nodes = s_qw + D + 1;                        // one contribution at a time
    __shared__ Contrib<Real> s_cc;
    __shared__ double s_vws[3];
    __shared__ double s_G;
    const double *q = a.data + ds.offset, *I = q + ds.nq_pad, *sg = q + 2 * ds.nq_pad;
    for (int j = tid; j < D; j += T) { s_qx[j] = (Real)a.quad_x[j]; s_qw[j] = (Real)a.quad_w[j]; }
    for (int i = tid; i < nq; i += T) s_ft[i] = 0.;
    __syncthreads();
    const QuadTab<Real> qt = { D, a.mp.quad_rule, s_qx, s_qw, &a.smear, a.smear.n };
    if (tid == 0) {
        double G = 1.;
        if (sizeof(Real) == 4) {
            PVec p(a.mp, a.mp.hi);
            Contrib<Real> tmp;
            VWS o = O::template prep<false>(p, a.comp_exp, 1., qt, tmp, s_nodes);
            G = 1. / o.amp_raw;
            if (!(G > 0.) || isinf(G)) G = 1.;
        }
        s_G = G;
    }
    __syncthreads();
    const double G = s_G;
    double *o_vol = a.out + ((size_t)0 * a.n_chains + chain) * N;
    double *o_num = a.out + ((size_t)1 * a.n_chains + chain) * N;
    double *o_vsq = a.out + ((size_t)2 * a.n_chains + chain) * N;
    double *o_srf = a.out + ((size_t)3 * a.n_chains + chain) * N;
    double *o_mvol = a.out + ((size_t)4 * a.n_chains + chain) * N;
    double *o_mnum = a.out + ((size_t)5 * a.n_chains + chain) * N;
    double *o_mvsq = a.out + ((size_t)6 * a.n_chains + chain) * N;
    double *o_msrf = a.out + ((size_t)7 * a.n_chains + chain) * N;
    // pass 1: per contribution intensity, v/w/s, observability ratio
    for (int c = 0; c < N; ++c) {
        if (tid == 0) {
            PVec p(a.mp, a.rset + ((size_t)chain * N + c) * P);
            VWS o = O::template prep<false>(p, a.comp_exp, G, qt, s_cc, s_nodes);
            s_vws[0] = o.v; s_vws[1] = o.w; s_vws[2] = o.s;
        }
        __syncthreads();
        double mn = INFINITY;
        for (int i = tid; i < nq; i += T) {
            QRep<Real> Q;
            Q.set(q[i]);
            double it = (double)eval_q<MODEL, Real>(s_cc, s_nodes, qt, Q);
            s_ft[i] += it;
            if (it != 0.) mn = fmin(mn, sg[i] / it);
        }
        mn = warp_min(mn);
        if (lane == 0) s_red[warp] = mn;
        __syncthreads();
        if (tid == 0) {
            for (int wv = 1; wv < nwarp; ++wv) mn = fmin(mn, s_red[wv]);
            // stash: ratio w/v * G * min(sigma/I_int) == min(sigma*volFrac/(sc0*I_SI))
            o_mvol[c] = (s_vws[1] / s_vws[0]) * (mn * G);
            o_vol[c] = s_vws[1];      // w
            o_num[c] = s_vws[0];      // v
            o_srf[c] = s_vws[2];      // s
        }
        __syncthreads();
    }
    // scale / background of the summed intensity (:559)
    double pS = 0, pSi = 0, pSii = 0, sm = 0, smm = 0, smi = 0;
    for (int i = tid; i < nq; i += T) {
        double e = sg[i] == 0.0 ? 1. : sg[i];
        double w = 1. / (e * e), M = s_ft[i];
        pS += w; pSi += w * I[i]; pSii += w * I[i] * I[i];
        sm += w * M; smm += w * M * M; smi += w * M * I[i];
    }
    double vals[6] = {pS, pSi, pSii, sm, smm, smi};
    for (int k = 0; k < 6; ++k) {
        double v = warp_sum(vals[k]);
        if (lane == 0) s_red[warp * 6 + k] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (int wv = 0; wv < nwarp; ++wv) for (int k = 0; k < 6; ++k) t[k] += s_red[wv * 6 + k];
        LsqConst L;
        L.S = t[0]; L.Si = t[1]; L.Sii = t[2]; L.inv_nq = 1. / nq; L.find_bg = a.find_bg; L.pos_bg = a.pos_bg;
        double sa, sb, chi;
        lsq_solve(L, t[3], t[4], t[5], sa, sb, chi);
        const double sc0 = sa * G;               // SI scaling
        a.sc[(size_t)chain * 2] = sc0;
        a.sc[(size_t)chain * 2 + 1] = sb;
        double tv = 0, tn = 0, tq = 0, ts = 0;
        for (int c = 0; c < N; ++c) {
            double w = o_vol[c], v = o_num[c], s = o_srf[c];
            double vf = w * sc0 / v;             // modeldata.py:57-61
            double nf = vf / v, qf = vf * v, sf = nf * s;
            o_vol[c] = vf; o_num[c] = nf; o_vsq[c] = qf; o_srf[c] = sf;
            tv += vf; tn += nf; tq += qf; ts += sf;
            double mv = o_mvol[c];
            double mnum = mv / v;
            o_mnum[c] = mnum;
            o_mvsq[c] = mnum * mv * mv;          // :592
            o_msrf[c] = mnum * s;                // :594
        }
        for (int c = 0; c < N; ++c) {            // :596-604
            if (tn != 0.) { o_num[c] /= tn; o_mnum[c] /= tn; }
            if (tq != 0.) { o_vsq[c] /= tq; o_mvsq[c] /= tq; }
            if (ts != 0.) { o_srf[c] /= ts; o_msrf[c] /= ts; }
        }
        a.totals[(size_t)chain * 4] = tv; a.totals[(size_t)chain * 4 + 1] = tn;
        a.totals[(size_t)chain * 4 + 2] = tq; a.totals[(size_t)chain * 4 + 3] = ts;
    }
}

template <int MODEL>
static cudaError_t launch_fr(int prec, int grid, size_t smem, cudaStream_t st, const FractionArgs &a)
{
    if (prec == MCSAS_FP64) {
        auto k = fractions_kernel<MODEL, double>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, FR_THREADS, smem, st>>>(a);
    } else {
        auto k = fractions_kernel<MODEL, float>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, FR_THREADS, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t mcsas_launch_fractions(int prec, int model, int grid, size_t smem, cudaStream_t st,
                                   const FractionArgs &a)
{
    switch (model) {
    case 0: return launch_fr<0>(prec, grid, smem, st, a);
    case 1: return launch_fr<1>(prec, grid, smem, st, a);
    case 2: return launch_fr<2>(prec, grid, smem, st, a);
    case 3: return launch_fr<3>(prec, grid, smem, st, a);
    case 4: return launch_fr<4>(prec, grid, smem, st, a);
    case 5: return launch_fr<5>(prec, grid, smem, st, a);
    case 6: return launch_fr<6>(prec, grid, smem, st, a);
    case 7: return launch_fr<7>(prec, grid, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

// ---- pipe peaks ------------------------------------------------------------------------
__device__ __forceinline__ float mufu_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// 8 independent chains x (sin, cos, rcp) per iteration = 24 MUFU ops per thread-iteration
__global__ void __launch_bounds__(256) peak_mufu_kernel(int iters, float *out)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.1f * (threadIdx.x + 1) + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = __sinf(x[i]);
            float c = __cosf(s + 1.5f);
            x[i] = mufu_rcp(c + 2.5f);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// 8 independent accumulators x 8 FFMA per iteration = 64 FFMA = 128 flop per thread-iteration
__global__ void __launch_bounds__(256) peak_ffma_kernel(int iters, float *out)
{
    float x[8];
    const float a = 1.0000001f, b = 1e-7f * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.1f * i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

cudaError_t mcsas_launch_peaks(int which, int grid, int threads, int iters, float *out, cudaStream_t st)
{
    if (which == 0) peak_mufu_kernel<<<grid, threads, 0, st>>>(iters, out);
    else peak_ffma_kernel<<<grid, threads, 0, st>>>(iters, out);
    return cudaGetLastError();
}
