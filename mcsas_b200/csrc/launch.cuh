// launch.cuh -- host-side launch table shared by the per-model translation units.
#pragma once
#include "common.cuh"

// q points per eval thread the chain kernel is instantiated for (layout index -> QPT);
// eval threads per CTA <= 256 (+ one producer warp), cluster size 1/2/4/8 at launch.
static const int CHAIN_QPT[] = {1, 2, 4};
static const int N_CHAIN_LAYOUTS = 3;
#define CHAIN_MAX_EVAL_THREADS 256
// fp32 production kernel (chain_fast.cuh): max eval threads per CTA of layout QPT = 1, 2, 4
static const int FAST_MAX_EVAL_THREADS[] = {1024, 512, 256};

// node arrays kept per contribution (Ops<MODEL>::NARR)
static const int MODEL_NARR[MCSAS_N_MODELS] = {0, 0, 0, 2, 1, 2, 0, 0};

typedef cudaError_t (*chain_launch_fn)(int prec, int layout, int grid, int threads, size_t smem,
                                       cudaStream_t st, const ChainArgs &a);   // grid = chains * a.cluster
typedef cudaError_t (*eval_launch_fn)(int prec, int grid, size_t smem, cudaStream_t st,
                                      const EvalArgs &a);

typedef cudaError_t (*chain_query_fn)(int prec, int smear, int layout, int threads, size_t smem, int cluster,
                                      int *max_clusters, int *regs);

#define DECL_MODEL(N) \
    cudaError_t mcsas_launch_chain_##N(int, int, int, int, size_t, cudaStream_t, const ChainArgs &); \
    cudaError_t mcsas_query_chain_##N(int, int, int, int, size_t, int, int *, int *); \
    cudaError_t mcsas_launch_eval_##N(int, int, size_t, cudaStream_t, const EvalArgs &);
DECL_MODEL(0) DECL_MODEL(1) DECL_MODEL(2) DECL_MODEL(3)
DECL_MODEL(4) DECL_MODEL(5) DECL_MODEL(6) DECL_MODEL(7)
#undef DECL_MODEL
