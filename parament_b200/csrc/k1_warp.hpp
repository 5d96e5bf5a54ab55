// k1_warp.hpp -- host interface of the register-resident warp kernels (dim <= 16).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include "params.hpp"

namespace pb {

struct K1Plan {
    unsigned int grid;                 // CTAs of the chain kernel
    unsigned int chunks_per_pulse;     // warps cooperating on one pulse
    unsigned int partials_per_pulse;   // matrices the reduce kernel combines per pulse
    int reduce_in_cta;                 // 1: the warps of a CTA belong to one pulse and combine in shared memory
    int k3_warps;                      // warps per CTA of the reduce kernel
    size_t partial_elems;              // double2 elements of the partial buffer
};

K1Plan plan_k1(int npad, unsigned int batch, unsigned long long nsteps, int num_sms, bool horner);

// carr / out are device pointers in the context precision; Hfrag is the fragment-ordered matrix table.
cudaError_t launch_k1(int npad, bool fp64_io, const SeriesParams &p, const void *carr, const double2 *Hfrag,
                      double2 *partials, unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                      unsigned long long step_hi, void *out, cudaStream_t stream);

}  // namespace pb
