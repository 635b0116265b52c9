// model_inst.cu -- compiled once per model with -DMODEL_ID=<n>: instantiates the chain
// and eval kernels of that model for both precisions and all thread layouts.
#include "chain_kernel.cuh"
#include "launch.cuh"

#ifndef MODEL_ID
#error "compile with -DMODEL_ID=<mcsas_model_id>"
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

template <class Real, int QPT, int MAXT>
static cudaError_t launch_one(int grid, int threads, size_t smem, cudaStream_t st,
                              const ChainArgs &a)
{
    auto k = chain_kernel<MODEL_ID, Real, QPT, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

template <class Real>
static cudaError_t launch_prec(int layout, int grid, int threads, size_t smem, cudaStream_t st,
                               const ChainArgs &a)
{
    switch (layout) {
    case 0: return launch_one<Real, 1, 128>(grid, threads, smem, st, a);
    case 1: return launch_one<Real, 2, 128>(grid, threads, smem, st, a);
    case 2: return launch_one<Real, 4, 128>(grid, threads, smem, st, a);
    case 3: return launch_one<Real, 8, 128>(grid, threads, smem, st, a);
    case 4: return launch_one<Real, 4, 256>(grid, threads, smem, st, a);
    case 5: return launch_one<Real, 8, 256>(grid, threads, smem, st, a);
    case 6: return launch_one<Real, 8, 512>(grid, threads, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t CAT(mcsas_launch_chain_, MODEL_ID)(int prec, int layout, int grid, int threads,
                                               size_t smem, cudaStream_t st, const ChainArgs &a)
{
    return prec == MCSAS_FP64 ? launch_prec<double>(layout, grid, threads, smem, st, a)
                              : launch_prec<float>(layout, grid, threads, smem, st, a);
}

cudaError_t CAT(mcsas_launch_eval_, MODEL_ID)(int prec, int grid, size_t smem, cudaStream_t st,
                                              const EvalArgs &a)
{
    if (prec == MCSAS_FP64) {
        auto k = eval_kernel<MODEL_ID, double>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, 128, smem, st>>>(a);
    } else {
        auto k = eval_kernel<MODEL_ID, float>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, 128, smem, st>>>(a);
    }
    return cudaGetLastError();
}
