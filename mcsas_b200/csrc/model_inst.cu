// model_inst.cu -- compiled once per model with -DMODEL_ID=<n>: instantiates the chain
// and eval kernels of that model for both precisions and all thread layouts.
#include "chain_fast.cuh"
#include "launch.cuh"

#ifndef MODEL_ID
#error "compile with -DMODEL_ID=<mcsas_model_id>"
#endif
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

// kernel of one (precision, layout): fp64 -> parity kernel, fp32 -> production kernel
// SMEAR selects the production kernel's instantiation for smeared data (only for the models
// the reference can smear; the parity kernel decides at run time)
template <class Real, int QPT, bool SMEAR> struct KernelOf {
    static auto get() { return chain_kernel<MODEL_ID, Real, QPT>; }
};
struct CachedF {};            // tag: fp32 production kernel with the contribution-row cache
template <int QPT, bool SMEAR> struct KernelOf<CachedF, QPT, SMEAR> {
    static auto get() { return chain_fast_kernel<MODEL_ID, QPT, true, SMEAR && CanSmear<MODEL_ID>::value>; }
};

template <class Real, int QPT, bool SMEAR>
static cudaError_t launch_one(int grid, int threads, size_t smem, cudaStream_t st,
                              const ChainArgs &a)
{
    auto k = KernelOf<Real, QPT, SMEAR>::get();
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);          // eval threads + one producer warp
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, a);
}

template <class Real, int QPT, bool SMEAR>
static cudaError_t query_one(int threads, size_t smem, int cluster, int *max_clusters, int *regs)
{
    auto k = KernelOf<Real, QPT, SMEAR>::get();
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k);
    if (e != cudaSuccess) return e;
    *regs = fa.numRegs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster * 1024, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(max_clusters, k, &cfg);
}

template <class Real, bool SMEAR>
static cudaError_t query_prec(int layout, int threads, size_t smem, int cluster, int *mc, int *regs)
{
    switch (layout) {
    case 0: return query_one<Real, 1, SMEAR>(threads, smem, cluster, mc, regs);
    case 1: return query_one<Real, 2, SMEAR>(threads, smem, cluster, mc, regs);
    case 2: return query_one<Real, 4, SMEAR>(threads, smem, cluster, mc, regs);
    }
    return cudaErrorInvalidValue;
}

// co-resident clusters of this configuration on the current device
cudaError_t CAT(mcsas_query_chain_, MODEL_ID)(int prec, int smear, int layout, int threads, size_t smem,
                                              int cluster, int *max_clusters, int *regs)
{
    if (prec == MCSAS_FP64) return query_prec<double, false>(layout, threads, smem, cluster, max_clusters, regs);
    return smear ? query_prec<CachedF, true>(layout, threads, smem, cluster, max_clusters, regs)
                 : query_prec<CachedF, false>(layout, threads, smem, cluster, max_clusters, regs);
}

template <class Real, bool SMEAR>
static cudaError_t launch_prec(int layout, int grid, int threads, size_t smem, cudaStream_t st,
                               const ChainArgs &a)
{
    switch (layout) {
    case 0: return launch_one<Real, 1, SMEAR>(grid, threads, smem, st, a);
    case 1: return launch_one<Real, 2, SMEAR>(grid, threads, smem, st, a);
    case 2: return launch_one<Real, 4, SMEAR>(grid, threads, smem, st, a);
    }
    return cudaErrorInvalidValue;
}

cudaError_t CAT(mcsas_launch_chain_, MODEL_ID)(int prec, int layout, int grid, int threads,
                                               size_t smem, cudaStream_t st, const ChainArgs &a)
{
    if (prec == MCSAS_FP64) return launch_prec<double, false>(layout, grid, threads, smem, st, a);
    if (a.smear.n > 0)
            return launch_prec<CachedF, true>(layout, grid, threads, smem, st, a);
    return launch_prec<CachedF, false>(layout, grid, threads, smem, st, a);
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
