// Micro-benchmark: issue rate of FP32 FMA forms on sm_100a (cycles per warp instruction per SMSP, 16 warps/SM)
#include <cstdio>
#include <cuda_runtime.h>
#define N 16
template <int MODE> __global__ void __launch_bounds__(256, 2) k(float *out, int iters, float a0, float b0)
{
    float x[N], y[N], z[N];
    float2 X[N / 2], Y[N / 2], Z[N / 2];
    for (int i = 0; i < N; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = a0 + i * 1e-7f; z[i] = b0 + i * 1e-8f; }
    for (int i = 0; i < N / 2; ++i) { X[i] = make_float2(x[2 * i], x[2 * i + 1]); Y[i] = make_float2(y[2 * i], y[2 * i + 1]); Z[i] = make_float2(z[2 * i], z[2 * i + 1]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 0) {            // FFMA, 3 distinct registers per instruction
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = fmaf(x[i], y[i], z[i]);
            } else if (MODE == 1) {     // FFMA, one immediate
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = fmaf(x[i], y[i], 1.25f);
            } else if (MODE == 2) {     // FFMA2, 3 distinct register pairs
#pragma unroll
                for (int i = 0; i < N / 2; ++i) X[i] = __ffma2_rn(X[i], Y[i], Z[i]);
            } else if (MODE == 3) {     // FFMA2, broadcast scalar + immediate
#pragma unroll
                for (int i = 0; i < N / 2; ++i) X[i] = __ffma2_rn(X[i], make_float2(y[0], y[0]), make_float2(1.25f, 1.25f));
            } else if (MODE == 4) {     // FFMA, shared multiplier register (2 distinct + acc)
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = fmaf(x[i], y[0], z[i]);
            } else if (MODE == 5) {     // FMUL 2 regs
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] = x[i] * y[i];
            }
        }
    }
    float acc = 0.f;
    for (int i = 0; i < N; ++i) acc += x[i];
    for (int i = 0; i < N / 2; ++i) acc += X[i].x + X[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(float *out, const char *name, int per_iter)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<296, 256>>>(out, 100, 1.0000001f, 1e-7f);
    cudaEventRecord(e0);
    k<MODE><<<296, 256>>>(out, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst_per_smsp = (double)iters * per_iter * 16 / 4;     // 16 warps per SM
    printf("%-44s %7.3f ms  %.2f cycles per warp instruction per SMSP\n", name, ms, ms * 1e-3 * 1.965e9 / inst_per_smsp);
}
int main()
{
    float *out; cudaMalloc(&out, 1 << 22);
    run<0>(out, "FFMA  3 distinct registers", 8 * N);
    run<1>(out, "FFMA  2 registers + immediate", 8 * N);
    run<4>(out, "FFMA  shared multiplier register", 8 * N);
    run<5>(out, "FMUL  2 registers", 8 * N);
    run<2>(out, "FFMA2 3 distinct register pairs", 8 * N / 2);
    run<3>(out, "FFMA2 pair, broadcast scalar, immediate", 8 * N / 2);
    return 0;
}
