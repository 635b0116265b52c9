// common.cuh -- shared POD types of the sm_100a MC-chain kernels (host + device).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/mcsas_b200.h"

#define MCSAS_MAX_NODES 512      // orientation-average nodes kept in shared memory
#define MCSAS_MAX_SMEAR 64       // points of the smearing profile (sasconfig.py: nSteps + 1 or + 2)
#define KCAP MCSAS_MAX_SPEC       // speculative proposals per batch
#define RING_MAX 128              // ring entries (>= 4 * KCAP)
#define MAX_CLUSTER 8             // CTAs per chain
#define MAX_EVAL_WARPS 8          // eval warps per CTA (+ one producer warp)

// Model instance state (replaces the ScatteringModel object, scatteringmodel.py).
struct ModelParams {
    int model_id;
    int n_params;
    int n_active;
    int quad_rule;
    int quad_nodes;
    int active_idx[MCSAS_MAX_ACTIVE];
    int gen_kind[MCSAS_MAX_ACTIVE];
    double params[MCSAS_MAX_PARAMS];
    double lo[MCSAS_MAX_ACTIVE];
    double hi[MCSAS_MAX_ACTIVE];
};

// Smearing profile (sasmodel.py:56-73): intensity(q_i) = sum_k tw[k] I(loc_ik) with
// loc_ik = sqrt(q_i^2 + off[k]^2) (slit) or q_i + off[k] (2-D averaged); tw = 2 * trapezoid
// coefficients of qOffset * profile weights.  n = 0: no smearing.
struct SmearTab {
    int n, two_d;
    double off[MCSAS_MAX_SMEAR];
    double tw[MCSAS_MAX_SMEAR];
};

// One dataset inside the packed device array: [q | I | sigma], each nq_pad doubles.
struct DatasetDesc {
    long long offset;   // in doubles, 16-byte aligned
    int nq;
    int nq_pad;
    double q_max;       // data.x0.limit[1] (mcsas.py:313)
    double q_ref;       // sqrt(q_min q_max): centre of the fp32 kernel's per-q scale factors
    double i_scale;     // max |I|: the fp32 kernel normalises I and sigma by it
};

#define CHAIN_ID_BACKUP 0x40000000   // flag in ChainArgs::chain_ids: the slot runs a backup attempt (capi.cu)
struct ChainArgs {
    ModelParams mp;
    SmearTab smear;
    // settings (mcsasparameters.json)
    int n_contribs;
    int n_reps;
    int n_chains;
    int max_attempts;
    long long max_iter;
    double conv_crit;
    double comp_exp;
    int find_bg, pos_bg, start_min;
    int spec_max;                // max speculative proposals per batch (<= ring / 4)
    int ring;                    // ring entries = min(RING_MAX, n_contribs)
    int cluster;                 // CTAs per chain (q split)
    int gram;                    // fp32 kernel: 1 = Gram batches (multi-accept chunks) when moves are frequent
    int gram_in, gram_out;       // ... entered by a move at most gram_in steps behind the previous one, left by a
                                 // Gram batch with fewer than gram_out moves (part of the chain's definition)
    int warps_q;                 // fp32 kernel: warps covering one CTA's q slice (eval warps = warps_q * groups)
    int resume;                  // 1: continue from the saved state (rset, counters)
    long long slice_steps;       // MC steps per chain in this launch, 0 = run to the end
    int slice_batches;           // fp32 kernel: batches per chain in this launch (0 = no limit)
    long long slice_ns;          // fp32 kernel: time budget of this launch per chain in ns (0 = no limit)
    int gw_cap;                  // fp32 kernel: capacity (in warps of a whole chain) of the exchange buffers
    int xp_cap;                  // ... and in proposals per batch (>= spec_max: FAST_KCAP where it fits, else FAST_KDEF)
    const int *chain_ids;        // chains of this launch (grid = n * cluster) or NULL
    int *needed_left;            // fp32 kernel: NULL or the primaries (ids without CHAIN_ID_BACKUP) of this time slice that
                                 // are still in their launch; a backup attempt ends its launch when it is 0
    int repack_ctas;             // > 0: every chain ends its launch once chains holding this many CTAs have left theirs
                                 // (needed_left[1] counts them)
    int *status;                 // per chain: 1 finished, 0 needs another slice
    int first_chain;
    int record_trace;
    unsigned long long seed;
    // inputs
    const DatasetDesc *sets;
    const double *data;
    const double *quad_x;        // [quad_nodes] node positions (model specific meaning)
    const double *quad_w;        // [quad_nodes] weights
    const double *stream;        // NULL or [n_chains][stream_stride]
    long long stream_stride;
    // outputs, one slot per chain
    double *rset;                // [chain][N][P]
    double *fit_meas;            // [chain][nq_max]
    int nq_stride;
    double *scaling, *background, *conval, *conval0, *elapsed_s;
    long long *num_iter, *num_moves, *evals, *steps_all;
    int *attempts;
    // fp32 kernel: what a chain carries across launches besides rset and the counters
    double *ft_save;             // [chain][nq_stride] fp64 master copy of the model intensity
    double *cen_a, *cen_b;       // centre of the quadratic expansion (fit of the last accepted state)
    double *cv_run;              // chi-squared the accept test compares against
    double *last_acc;            // step of the last accepted move (decides which batches are Gram batches)
    int *batch_k;                // speculative batch size | Gram mode << 16
    unsigned char *trace_accept; // [chain][max_iter] or NULL
    double *trace_chi;           // [chain][max_iter] or NULL
    float *cache;                // fp32 kernel: NULL or [SM id][cache_stride] contribution rows
    long long cache_stride;      // floats per SM (n_contribs * FAST_QCAP)
    int *cache_lock;             // [SM id] 0 = free
    int keep_rows;               // cached rows [0, keep_rows) of every chain are loaded / stored evict_last, the
                                 // others evict_first (set per slice: what the L2 holds of the resident chains' rows)
    long long *prof;             // NULL or [chain][2][8] phase cycle counters (MCSAS_B200_PROF)
    const float4 *smloc;         // fp32 sphere kernel, smeared data: NULL or [set][profile point][smloc_row] table of
                                 // (q, split q/(2 pi), weight * (q_ref/q)^6) at the smeared locations of every q point
    int smloc_row;               // entries per profile point (>= q points of the largest chain layout)
    int smloc_mono;              // 1: every smeared location >= its q point (the series vote may use the unsmeared q)
};

struct EvalArgs {
    ModelParams mp;
    SmearTab smear;
    const DatasetDesc *sets;
    const double *data;
    const double *quad_x, *quad_w;
    int set;
    int m;                       // contributions
    double comp_exp;
    const double *pset;          // [m][P]
    double *intensity;           // [m][nq]
    double *v, *w, *s;           // [m]
};

struct FractionArgs {
    ModelParams mp;
    SmearTab smear;
    const DatasetDesc *sets;
    const double *data;
    const double *quad_x, *quad_w;
    int n_contribs, n_reps, n_chains;
    double comp_exp;
    int find_bg, pos_bg;
    const double *rset;          // [chain][N][P]
    double *out;                 // [8][chain][N]: vol,num,volsqr,surf,minvol,minnum,minvolsqr,minsurf
    double *totals;              // [chain][4]
    double *sc;                  // [chain][2]
};

// Shared-memory footprint of chain_kernel (mirrors its carve-up).
__host__ __device__ inline size_t chain_smem_bytes(int nq_pad, int N, int P, int D, int narr,
                                                   size_t real, int ring)
{
#define MCSAS_AL16(b) ((((size_t)(b)) + 15) & ~size_t(15))
    size_t s = 16;                                               // mbarrier
    s += MCSAS_AL16(8 * 3 * (size_t)nq_pad);                             // data
    s += MCSAS_AL16(8 * (size_t)N * (P > 0 ? P : 1));                    // rset replica
    s += MCSAS_AL16(64);                                                 // ctrl
    s += MCSAS_AL16((size_t)ring * (8 * MCSAS_MAX_ACTIVE + 2 * real * 8));   // ring entries
    s += MCSAS_AL16(8 * (MAX_EVAL_WARPS + 1) * KCAP * 3);                // per-warp sums
    s += MCSAS_AL16(8 * 2 * MAX_CLUSTER * KCAP * 3);                     // cluster partials (2 buffers)
    s += MCSAS_AL16(8 * 2 * MAX_CLUSTER) + 16;                           // exchange flags + mbarriers
    s += MCSAS_AL16(8 * KCAP * 3 + 64);                                  // chi/a/b per proposal + decision
    s += 2 * MCSAS_AL16(real * (D + 1));                                 // quadrature table
    s += MCSAS_AL16(real * (size_t)ring * 2 * narr * D) + 16;            // node arrays
    return s;
#undef MCSAS_AL16
}

// ---- shared-memory footprint of chain_fast_kernel (mirrors its carve-up) -----------------------------
#define FAST_MAX_WARPS 32
#define FAST_KCAP 128              // fp32 kernel: speculative proposals per batch (decided 32 at a time)
#define FAST_KDEF 64               // ... by default on one or two CTAs per chain (every SM has other work)
#define FAST_RING_MAX 512          // ... ring entries (>= 4 * FAST_KCAP)
#define FAST_CACHE_SMS 256         // SMs row regions are allocated for (>= %nsmid of a B200)
#define FAST_CACHE_SLOTS 2         // row regions per SM (co-resident CTAs of this kernel)
#define FAST_QCAP 1024            // q points per CTA (QPT * 32 * warps_q)
// warps of a whole chain (cluster size x warps per CTA slice) for the largest layout the planner
// may choose: the q slice of a CTA is a whole number of warps of 32 * 4 q points
__host__ __device__ inline int fast_gw_cap(int nq)
{
    int cap = 1;
    for (int c = 1; c <= MAX_CLUSTER; c *= 2) {
        const int n = (nq + c - 1) / c, wq = (n + 127) / 128;
        if (c * wq > cap) cap = c * wq;
    }
    return cap;
}
__host__ __device__ inline size_t fast_smem_bytes(int nq_pad, int N, int P, int D, int narr, int ring, int gw_cap,
                                                  int xp_cap)
{
#define AL16(b) ((((size_t)(b)) + 15) & ~size_t(15))
    size_t s = 16;                                               // mbarrier
    s += AL16(8 * 3 * (size_t)nq_pad);                           // data staging
    s += AL16(8 * (size_t)N * (P > 0 ? P : 1));                  // rset
    s += AL16(64);                                               // ctrl
    s += AL16((size_t)ring * (8 * MCSAS_MAX_ACTIVE + 2 * 4 * 8));// ring entries
    s += AL16(16 * 2 * (size_t)gw_cap * xp_cap);                 // exchange: per-warp proposal sums (float4), 2 buffers
    s += AL16(8 * 2 * (size_t)gw_cap * 8);                       // exchange: per-warp base sums + flags
    s += AL16(4 * 2 * (size_t)gw_cap * 32);                      // exchange: per-warp Gram pairs
    s += 16;                                                     // exchange mbarriers
    s += AL16(4 * 32);                                           // Gram totals
    s += AL16(8 * (size_t)FAST_KCAP + 256);                      // chi per proposal + decision, base sums, cold loop state
    s += AL16(8 * 2 * (size_t)gw_cap * 3);                       // full-LSQ all-sum
    s += AL16(8 * 2 * FAST_QCAP) + 2 * AL16(8 * FAST_QCAP) + AL16(4 * FAST_QCAP);   // ft (x2), I, 1/sigma, q scale
    s += 2 * AL16(4 * (size_t)(D + 1));                          // quadrature table
    s += AL16(4 * (size_t)ring * 2 * narr * D) + 16;             // node arrays
    return s;
#undef AL16
}
