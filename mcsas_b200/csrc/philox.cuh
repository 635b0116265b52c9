// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11) for in-kernel
// proposals.  Replaces numpy.random.uniform of RandomUniform.get / RandomExponential.get
// (numbergenerator.py:28-31,168-175) in production mode.
//
// Stream convention (shared with oracle/mcsas_oracle.c: orc_philox_stream):
//   key     = (seed_lo, seed_hi)
//   counter = (idx_lo, idx_hi, global chain id, domain)
//   domain  = 2*attempt + 0: per-step draws, idx = MC step
//             2*attempt + 1: initial draws,  idx = contribution index
//   output word k -> active parameter k;  u = (word + 0.5) * 2^-32  in (0,1)
#pragma once
#include <cstdint>

__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t idx, uint32_t chain,
                                              uint32_t domain, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = chain, c3 = domain;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Number-generator output in [0,1) from a raw 32-bit word
// (RandomUniform: u;  RandomExponential<k>: (10^(k u) - 1) / 10^k).
__device__ __forceinline__ double generator_output(uint32_t word, int gen_kind)
{
    double u = ((double)word + 0.5) * (1.0 / 4294967296.0);
    if (gen_kind > 0) {
        double up = (double)gen_kind;
        u = (pow(10., up * u) - 1.) / pow(10., up);
    }
    return u;
}
