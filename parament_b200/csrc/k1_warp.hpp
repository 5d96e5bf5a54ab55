// k1_warp.hpp -- host interface of the register-resident warp kernels (dim <= 16).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include "params.hpp"
#include "plan.hpp"

namespace pb {

// carr / out are device pointers in the context precision; Hfrag is the fragment-ordered matrix table.
// The chain kernel (kernels 1+2) leaves plan.partials_per_pulse partial products per pulse at `partials`; the reduce
// kernel (kernel 3) combines `partials_per_pulse` consecutive partials of each pulse in order and writes the propagator.
// Final stage of a chain launch (k1_common.cuh).  out == nullptr: partial products only, launch_k3_reduce follows.
struct K1Final {
    void *out;                 // propagators in the context precision, n x n row-major per pulse
    int n;
    unsigned int *counters;    // [pulse][groups + 1] arrival counters, zero on entry and zero again on exit
    double2 *mid;              // [pulse][groups] group products
    unsigned int groups;       // groups of K1_GROUP CTA partials per pulse (few-long-pulses mode)
};
constexpr unsigned int K1_GROUP = 16;

cudaError_t launch_k1_chain(int npad, bool fp64_io, const SeriesParams &p, const void *carr, const double2 *Hfrag,
                            double2 *partials, unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                            unsigned long long step_hi, const K1Final &fz, cudaStream_t stream);
cudaError_t launch_k3_reduce(int npad, bool fp64_io, const double2 *partials, unsigned int partials_per_pulse, int n,
                             void *out, unsigned int batch, double2 *mid, cudaStream_t stream);

// complex64 contexts, dim <= 8, degree-8 form: the chain kernel in FP32 arithmetic on the TF32 tensor path (k1_tf32.cu); same
// plan, partial layout and final stage as launch_k1_chain.
int k1_tf32_ctas_per_sm();
cudaError_t launch_k1_tf32_chain(const SeriesParams &p, const void *carr, const double2 *Hfrag, double2 *partials, unsigned int batch,
                                 const K1Plan &plan, unsigned long long step_lo, unsigned long long step_hi, const K1Final &fz,
                                 cudaStream_t stream);

// real products per complex matrix product in the chain kernel that launch_k1_chain selects (3 for the degree-8 form at dim 9..16)
int k1_real_products(int npad, bool fp64_io, int horner);

// Multi-GPU combine for dim <= 16 in one launch: out = parts[count-1] ... parts[0]; parts / out are dim x dim propagators in the
// context precision on the device.
cudaError_t launch_k3_combine(int npad, bool fp64_io, const void *parts, unsigned int count, int n, void *out, cudaStream_t stream);

}  // namespace pb
