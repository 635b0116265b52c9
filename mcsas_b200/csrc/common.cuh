// common.cuh -- shared POD types of the sm_100a MC-chain kernels (host + device).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/mcsas_b200.h"

#define MCSAS_MAX_NODES 512      // orientation-average nodes kept in shared memory

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

// One dataset inside the packed device array: [q | I | sigma], each nq_pad doubles.
struct DatasetDesc {
    long long offset;   // in doubles, 16-byte aligned
    int nq;
    int nq_pad;
    double q_max;       // data.x0.limit[1] (mcsas.py:313)
};

struct ChainArgs {
    ModelParams mp;
    // settings (mcsasparameters.json)
    int n_contribs;
    int n_reps;
    int n_chains;
    int max_attempts;
    long long max_iter;
    double conv_crit;
    double comp_exp;
    int find_bg, pos_bg, start_min;
    int spec_max;
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
    volatile int *stop_flag;     // mapped host int or NULL
    // outputs, one slot per chain
    double *rset;                // [chain][N][P]
    double *fit_meas;            // [chain][nq_max]
    int nq_stride;
    double *scaling, *background, *conval, *conval0, *elapsed_s;
    long long *num_iter, *num_moves, *evals;
    int *attempts;
    unsigned char *trace_accept; // [chain][max_iter] or NULL
    double *trace_chi;           // [chain][max_iter] or NULL
};

struct EvalArgs {
    ModelParams mp;
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
