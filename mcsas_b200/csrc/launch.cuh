// launch.cuh -- host-side launch table shared by the per-model translation units.
#pragma once
#include "common.cuh"

// Thread/q-point layouts: (q points per thread, max threads per CTA).
struct ChainLayout { int qpt; int max_threads; };
static const ChainLayout CHAIN_LAYOUTS[] = {
    {1, 128}, {2, 128}, {4, 128}, {8, 128}, {4, 256}, {8, 256}, {8, 512}
};
static const int N_CHAIN_LAYOUTS = 7;

// node arrays kept per contribution (Ops<MODEL>::NARR)
static const int MODEL_NARR[MCSAS_N_MODELS] = {0, 0, 0, 2, 1, 2, 0, 0};

typedef cudaError_t (*chain_launch_fn)(int prec, int layout, int grid, int threads, size_t smem,
                                       cudaStream_t st, const ChainArgs &a);
typedef cudaError_t (*eval_launch_fn)(int prec, int grid, size_t smem, cudaStream_t st,
                                      const EvalArgs &a);

#define DECL_MODEL(N) \
    cudaError_t mcsas_launch_chain_##N(int, int, int, int, size_t, cudaStream_t, const ChainArgs &); \
    cudaError_t mcsas_launch_eval_##N(int, int, size_t, cudaStream_t, const EvalArgs &);
DECL_MODEL(0) DECL_MODEL(1) DECL_MODEL(2) DECL_MODEL(3)
DECL_MODEL(4) DECL_MODEL(5) DECL_MODEL(6) DECL_MODEL(7)
#undef DECL_MODEL
