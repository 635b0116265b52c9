// dataprep.cu -- batched data preparation: what DataObj does between a raw
// (q, I, sigma) table and the vectors McSAS.mcFit reads, for MANY datasets in one launch:
//   DataObj._prepareUncertainty          src/mcsas/dataobj/dataobj.py:204-226   sigma <- max(sigma, fuMin I); invalid -> inf
//   _initMask/_applyFMasks/_applyLimits  :239-269                               finite I, zero / negative masks, q window
//   DataObj._reBin                       :288-345                               log-spaced bins, means, max(SEM, propagated)
// One CTA per dataset, one thread per bin.  The result is BIT-IDENTICAL to the reference's numpy
// arithmetic: the bin edges come from the host (numpy.logspace of the reference, :308-311 -- a
// device pow() could move a point across an edge), the sums follow numpy's pairwise summation
// (8 accumulators over blocks of <= 128 elements, halves above that; numpy/_core/src/umath/
// loops_utils.h.src: pairwise_sum) in the order of the input rows, and no multiply-add is
// contracted.  HBM-bound integer/compare work: O(n_raw) bytes per dataset, no tensor cores.
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>

#define PREP_THREADS 128

struct PrepArgs {
    int n_sets, n_bin, cap, has_sigma;
    const long long *raw_off;    // [n_sets + 1] row offsets into q / I / sigma
    const double *q, *I, *sigma; // concatenated raw tables, q in 1/m
    const double *edges;         // [n_sets][n_bin + 1] or NULL (n_bin <= 0: no rebinning)
    double fu_min, x0_low, x0_high;
    int mask_zero, mask_neg;
    int *scratch;                // [total rows] row indices grouped by bin
    int *nq_out;                 // [n_sets]
    double *q_out, *I_out, *s_out;   // [n_sets][cap]
};

__device__ __forceinline__ bool prep_valid(const PrepArgs &a, long long r)
{
    const double I = a.I[r], q = a.q[r];
    bool m = isfinite(I);
    if (a.mask_zero) m = m && I != 0.0;
    if (a.mask_neg) m = m && I > 0.0;
    return m && q >= a.x0_low && q <= a.x0_high;
}

// dataobj.py:204-226
__device__ __forceinline__ double prep_unc(const PrepArgs &a, long long r)
{
    const double fl = __dmul_rn(a.fu_min, a.I[r]);
    double u = fl;
    if (a.has_sigma) {
        const double s = a.sigma[r];
        u = (isnan(s) || isnan(fl)) ? NAN : (s > fl ? s : fl);      // numpy.maximum propagates NaN
    }
    return isfinite(u) ? u : INFINITY;
}

// numpy's pairwise summation of n values f(k), k = off .. off + n - 1: blocks of <= 128 values are
// summed with 8 interleaved accumulators, longer ranges are split in halves (first half rounded
// down to a multiple of 8) -- written with an explicit stack (depth <= log2(n / 128) + 1)
template <class F> __device__ __forceinline__ double pairwise_leaf(const F &f, int off, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, f(off + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(off + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(off + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(off + i));
    return res;
}

template <class F> __device__ double pairwise_sum(const F &f, int off0, int n0)
{
    if (n0 <= 128) return pairwise_leaf(f, off0, n0);
    int s_off[28], s_n[28], s_stage[28];
    double s_left[28];
    int sp = 0;
    s_off[0] = off0; s_n[0] = n0; s_stage[0] = 0; sp = 1;
    double ret = 0.;
    while (sp > 0) {
        const int t = sp - 1;
        const int off = s_off[t], n = s_n[t];
        int n2 = n / 2;
        n2 -= n2 % 8;
        if (s_stage[t] == 0) {
            if (n <= 128) { ret = pairwise_leaf(f, off, n); --sp; continue; }
            s_stage[t] = 1;
            s_off[sp] = off; s_n[sp] = n2; s_stage[sp] = 0; ++sp;
        } else if (s_stage[t] == 1) {
            s_left[t] = ret;
            s_stage[t] = 2;
            s_off[sp] = off + n2; s_n[sp] = n - n2; s_stage[sp] = 0; ++sp;
        } else {
            ret = __dadd_rn(s_left[t], ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(PREP_THREADS) prepare_kernel(const PrepArgs a)
{
    const int set = blockIdx.x, tid = threadIdx.x;
    const long long r0 = a.raw_off[set];
    const int n = (int)(a.raw_off[set + 1] - r0);
    double *qo = a.q_out + (size_t)set * a.cap, *Io = a.I_out + (size_t)set * a.cap, *so = a.s_out + (size_t)set * a.cap;
    __shared__ int s_scan[PREP_THREADS];
    __shared__ int s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    if (a.n_bin <= 0) {
        // no rebinning (dataobj.py:298-302): the sanitized rows in order
        for (int c0 = 0; c0 < n; c0 += PREP_THREADS) {
            const int i = c0 + tid;
            const bool v = i < n && prep_valid(a, r0 + i);
            s_scan[tid] = v ? 1 : 0;
            __syncthreads();
            for (int d = 1; d < PREP_THREADS; d <<= 1) {
                const int x = tid >= d ? s_scan[tid - d] : 0;
                __syncthreads();
                s_scan[tid] += x;
                __syncthreads();
            }
            const int pos = s_base + s_scan[tid] - 1;
            if (v && pos < a.cap) { qo[pos] = a.q[r0 + i]; Io[pos] = a.I[r0 + i]; so[pos] = prep_unc(a, r0 + i); }
            __syncthreads();
            if (tid == PREP_THREADS - 1) s_base += s_scan[tid];
            __syncthreads();
        }
        if (tid == 0) a.nq_out[set] = s_base;
        return;
    }
    const double *e = a.edges + (size_t)set * (a.n_bin + 1);
    int *scr = a.scratch + r0;
    for (int b0 = 0; b0 < a.n_bin; b0 += PREP_THREADS) {
        const int b = b0 + tid;
        const bool has_bin = b < a.n_bin;
        const double lo = has_bin ? e[b] : 0., hi = has_bin ? e[b + 1] : 0.;
        // rows of this bin, in input order: (sanX >= xEdges[b]) & (sanX < xEdges[b + 1])  (:317)
        int cnt = 0;
        if (has_bin)
            for (int i = 0; i < n; ++i) {
                const double qv = a.q[r0 + i];
                if (qv >= lo && qv < hi && prep_valid(a, r0 + i)) ++cnt;
            }
        s_scan[tid] = cnt;
        __syncthreads();
        for (int d = 1; d < PREP_THREADS; d <<= 1) {
            const int x = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += x;
            __syncthreads();
        }
        const int off = s_scan[tid] - cnt;                     // within this block of bins; bins are disjoint
        const int base = s_base;                               // rows already grouped by earlier bin blocks
        // (the scratch segment of a dataset has n ints; every row belongs to at most one bin)
        if (has_bin && cnt > 0) {
            int k = 0;
            for (int i = 0; i < n; ++i) {
                const double qv = a.q[r0 + i];
                if (qv >= lo && qv < hi && prep_valid(a, r0 + i)) scr[base + off + k++] = i;
            }
        }
        double bq = NAN, bf = NAN, bu = NAN;
        bool valid = false;
        if (has_bin && cnt == 1) {                             // :325-327
            const long long r = r0 + scr[base + off];
            bq = a.q[r]; bf = a.I[r]; bu = prep_unc(a, r);
            valid = true;
        } else if (has_bin && cnt > 1) {                       // :328-336
            const int *ix = scr + base + off;
            const double dn = (double)cnt;
            auto fI = [&](int k) { return a.I[r0 + ix[k]]; };
            auto fq = [&](int k) { return a.q[r0 + ix[k]]; };
            bf = __ddiv_rn(pairwise_sum(fI, 0, cnt), dn);      // fInBin.mean()
            bq = __ddiv_rn(pairwise_sum(fq, 0, cnt), dn);      // x0InBin.mean()
            const double mean = bf;
            auto fdev = [&](int k) { const double x = __dadd_rn(a.I[r0 + ix[k]], -mean); return __dmul_rn(x, x); };
            auto fu2 = [&](int k) { const double u = prep_unc(a, r0 + ix[k]); return __dmul_rn(u, u); };
            const double var = __ddiv_rn(pairwise_sum(fdev, 0, cnt), (double)(cnt - 1));     // std(ddof = 1) ** 2
            const double sem = __ddiv_rn(__dsqrt_rn(var), __dsqrt_rn(__dmul_rn(1., dn)));
            const double prop = __dsqrt_rn(__ddiv_rn(pairwise_sum(fu2, 0, cnt), dn));
            bu = (isnan(sem) || isnan(prop)) ? NAN : (sem > prop ? sem : prop);             // numpy.maximum
            valid = true;
        }
        const bool keep = valid && !isnan(bf);                 // :339-340
        __syncthreads();
        const int total_rows = s_scan[PREP_THREADS - 1];
        __syncthreads();
        // compact the kept bins of this block of bins behind the earlier ones
        s_scan[tid] = keep ? 1 : 0;
        __syncthreads();
        for (int d = 1; d < PREP_THREADS; d <<= 1) {
            const int x = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += x;
            __syncthreads();
        }
        __shared__ int s_nout;
        if (b0 == 0 && tid == 0) s_nout = 0;
        __syncthreads();
        const int pos = s_nout + s_scan[tid] - 1;
        if (keep && pos < a.cap) { qo[pos] = bq; Io[pos] = bf; so[pos] = bu; }
        __syncthreads();
        if (tid == PREP_THREADS - 1) { s_nout += s_scan[tid]; s_base += total_rows; }
        __syncthreads();
        if (b0 + PREP_THREADS >= a.n_bin && tid == 0) a.nq_out[set] = s_nout;
    }
}

cudaError_t mcsas_launch_prepare(const PrepArgs &a, cudaStream_t st)
{
    prepare_kernel<<<a.n_sets, PREP_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}
