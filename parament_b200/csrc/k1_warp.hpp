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
    int ctas_per_sm;                   // occupancy the chain kernel variant is compiled for
    size_t partial_elems;              // double2 elements of the partial buffer
};

K1Plan plan_k1(int npad, unsigned int batch, unsigned long long nsteps, int num_sms, bool horner);
unsigned int k1_warp_slots(int npad, int num_sms, bool horner);   // co-resident warps of the chain kernel (one wave)

// carr / out are device pointers in the context precision; Hfrag is the fragment-ordered matrix table.
// The chain kernel (kernels 1+2) leaves plan.partials_per_pulse partial products per pulse at `partials`; the reduce
// kernel (kernel 3) combines `partials_per_pulse` consecutive partials of each pulse in order and writes the propagator.
cudaError_t launch_k1_chain(int npad, bool fp64_io, const SeriesParams &p, const void *carr, const double2 *Hfrag,
                            double2 *partials, unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                            unsigned long long step_hi, cudaStream_t stream);
cudaError_t launch_k3_reduce(int npad, bool fp64_io, const double2 *partials, unsigned int partials_per_pulse, int n,
                             void *out, unsigned int batch, double2 *mid, cudaStream_t stream);
size_t k3_mid_elems(int npad, unsigned int batch, unsigned int partials_per_pulse);   // scratch of the two-level reduction
int k3_launches(unsigned int partials_per_pulse);
int k3_warps_for(unsigned int partials_per_pulse);

}  // namespace pb
