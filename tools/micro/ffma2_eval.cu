// Micro-benchmark: the sphere evaluation loop of chain_fast.cuh (8 proposals x 4 q points per thread,
// MUFU.SIN + MUFU.COS per evaluation, four delta sums), scalar fp32 against packed f32x2 arithmetic.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_eval ffma2_eval.cu && ./ffma2_eval
#include <cstdio>
#include <cuda_runtime.h>

template <int M>
__device__ __forceinline__ void g_scalar(const float (&q)[M], const float (&th)[M], const float (&tl)[M], float r, float (&g)[M])
{
    float x[M], t[M], fr[M], sn[M], cs[M];
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = fmaf(th[i], r, 12582912.f);
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = q[i] * r;
#pragma unroll
    for (int i = 0; i < M; ++i) t[i] = t[i] - 12582912.f;
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(th[i], r, -t[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fmaf(tl[i], r, fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) fr[i] = fr[i] * 6.2831853071795865f;
#pragma unroll
    for (int i = 0; i < M; ++i) sn[i] = __sinf(fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) cs[i] = __cosf(fr[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) g[i] = fmaf(-x[i], cs[i], sn[i]);
}

template <int BPS> __global__ void __launch_bounds__(256, BPS) k_scalar(const float *in, float *out, int iters)
{
    const int tid = threadIdx.x + blockIdx.x * blockDim.x;
    float q[4], th[4], tl[4], cq[4], rho[4], mu[4], om[4];
    for (int j = 0; j < 4; ++j) {
        q[j] = in[tid * 4 + j]; th[j] = q[j] * 0.15915494f; tl[j] = q[j] * 1e-9f;
        cq[j] = 1.f + q[j]; rho[j] = 0.5f * q[j]; mu[j] = 0.25f + q[j]; om[j] = 2.f;
    }
    float tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        float acc[8][4];
        float od[8][4], dn[8][4];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float4 v = __ldcg((const float4 *)(in + ((it * 8 + s) & 1023) * 1024 + threadIdx.x * 4));
            od[s][0] = v.x; od[s][1] = v.y; od[s][2] = v.z; od[s][3] = v.w;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float r = in[(it * 8 + s) & 4095] * 10.f + 20.f, aa = r * 0.01f;
            float g[4];
            g_scalar<4>(q, th, tl, r, g);
#pragma unroll
            for (int j = 0; j < 4; ++j) dn[s][j] = (g[j] * g[j]) * aa;
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = cq[j] * (dn[s][j] - od[s][j]);
                acc[s][0] = fmaf(rho[j], u, acc[s][0]);
                acc[s][1] = fmaf(u, u, acc[s][1]);
                acc[s][2] = fmaf(mu[j], u, acc[s][2]);
                acc[s][3] = fmaf(om[j], u, acc[s][3]);
            }
        }
#pragma unroll
        for (int h = 4, off = 16; h >= 1; h >>= 1, off >>= 1)
#pragma unroll
            for (int s = 0; s < h; ++s)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[s][i] += __shfl_xor_sync(0xffffffffu, acc[s + h][i], off);
#pragma unroll
        for (int off = 2; off >= 1; off >>= 1)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], off);
        tot += acc[0][0] + acc[0][1] + acc[0][2] + acc[0][3];
    }
    out[tid] = tot;
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// MP pairs of q points of ONE radius
template <int MP>
__device__ __forceinline__ void g_packed(const float2 (&q)[MP], const float2 (&th)[MP], const float2 (&tl)[MP], float r, float2 (&g)[MP])
{
    const float2 r2 = f2(r, r), magic = f2(12582912.f, 12582912.f), nmagic = f2(-12582912.f, -12582912.f);
    const float2 twopi = f2(6.2831853071795865f, 6.2831853071795865f);
    float2 x[MP], t[MP], fr[MP], sn[MP], cs[MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) t[i] = __ffma2_rn(th[i], r2, magic);
#pragma unroll
    for (int i = 0; i < MP; ++i) x[i] = __fmul2_rn(q[i], r2);
#pragma unroll
    for (int i = 0; i < MP; ++i) t[i] = __fadd2_rn(t[i], nmagic);
#pragma unroll
    for (int i = 0; i < MP; ++i) fr[i] = __ffma2_rn(th[i], r2, f2(-t[i].x, -t[i].y));
#pragma unroll
    for (int i = 0; i < MP; ++i) fr[i] = __ffma2_rn(tl[i], r2, fr[i]);
#pragma unroll
    for (int i = 0; i < MP; ++i) fr[i] = __fmul2_rn(fr[i], twopi);
#pragma unroll
    for (int i = 0; i < MP; ++i) sn[i] = f2(__sinf(fr[i].x), __sinf(fr[i].y));
#pragma unroll
    for (int i = 0; i < MP; ++i) cs[i] = f2(__cosf(fr[i].x), __cosf(fr[i].y));
#pragma unroll
    for (int i = 0; i < MP; ++i) g[i] = __ffma2_rn(f2(-x[i].x, -x[i].y), cs[i], sn[i]);
}

template <int BPS> __global__ void __launch_bounds__(256, BPS) k_packed(const float *in, float *out, int iters)
{
    const int tid = threadIdx.x + blockIdx.x * blockDim.x;
    float2 q[2], th[2], tl[2], cq[2], rho[2], mu[2], om[2];
    for (int j = 0; j < 2; ++j) {
        q[j] = f2(in[tid * 4 + 2 * j], in[tid * 4 + 2 * j + 1]);
        th[j] = f2(q[j].x * 0.15915494f, q[j].y * 0.15915494f); tl[j] = f2(q[j].x * 1e-9f, q[j].y * 1e-9f);
        cq[j] = f2(1.f + q[j].x, 1.f + q[j].y); rho[j] = f2(0.5f * q[j].x, 0.5f * q[j].y);
        mu[j] = f2(0.25f + q[j].x, 0.25f + q[j].y); om[j] = f2(2.f, 2.f);
    }
    float tot = 0.f;
    for (int it = 0; it < iters; ++it) {
        float acc[8][4];
        float2 od[8][2], dn[8][2];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float4 v = __ldcg((const float4 *)(in + ((it * 8 + s) & 1023) * 1024 + threadIdx.x * 4));
            od[s][0] = f2(-v.x, -v.y); od[s][1] = f2(-v.z, -v.w);
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float r = in[(it * 8 + s) & 4095] * 10.f + 20.f, aa = r * 0.01f;
            float2 g[2];
            g_packed<2>(q, th, tl, r, g);
#pragma unroll
            for (int j = 0; j < 2; ++j) dn[s][j] = __fmul2_rn(__fmul2_rn(g[j], g[j]), f2(aa, aa));
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            float2 a0 = f2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 u = __fmul2_rn(cq[j], __fadd2_rn(dn[s][j], od[s][j]));
                a0 = __ffma2_rn(rho[j], u, a0);
                a1 = __ffma2_rn(u, u, a1);
                a2 = __ffma2_rn(mu[j], u, a2);
                a3 = __ffma2_rn(om[j], u, a3);
            }
            acc[s][0] = a0.x + a0.y; acc[s][1] = a1.x + a1.y; acc[s][2] = a2.x + a2.y; acc[s][3] = a3.x + a3.y;
        }
#pragma unroll
        for (int h = 4, off = 16; h >= 1; h >>= 1, off >>= 1)
#pragma unroll
            for (int s = 0; s < h; ++s)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[s][i] += __shfl_xor_sync(0xffffffffu, acc[s + h][i], off);
#pragma unroll
        for (int off = 2; off >= 1; off >>= 1)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[0][i] += __shfl_xor_sync(0xffffffffu, acc[0][i], off);
        tot += acc[0][0] + acc[0][1] + acc[0][2] + acc[0][3];
    }
    out[tid] = tot;
}

template <int BPS> void run(const float *in, float *out, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep)
        for (int v = 0; v < 2; ++v) {
            cudaEventRecord(e0);
            if (v == 0) k_scalar<BPS><<<148 * BPS, 256>>>(in, out, iters); else k_packed<BPS><<<148 * BPS, 256>>>(in, out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 8 * BPS * 32 / 4);
            printf("%d warps/SM %s %8.3f ms  -> %.1f cycles per warp-evaluation per SMSP  (%s)\n", 8 * BPS, v ? "packed" : "scalar", ms, cyc, cudaGetErrorString(cudaGetLastError()));
        }
}
int main()
{
    float *in, *out;
    cudaMalloc(&in, 4 << 20); cudaMalloc(&out, 4 << 20);
    cudaMemset(in, 0x3c, 4 << 20);
    run<1>(in, out, 4000);
    run<2>(in, out, 4000);
    run<4>(in, out, 4000);
    return 0;
}
