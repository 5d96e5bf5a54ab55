// coef.cuh -- effective control coefficients of one time step, evaluated on the fly from the raw
// amplitude stream (replaces the reference's separate quadrature kernels + c2 array,
// control_expansion.cu:27-160 / parament.cpp:510-538).  Inputs are converted to double exactly; the
// arithmetic is double for both context precisions.
#pragma once
#include "params.hpp"

namespace pb {

// The amplitude stream is read once from HBM.  STREAM = true: evict-first loads keep it from displacing the L2-resident
// scratch of the kernels that have one.  STREAM = false (register kernel, no scratch): read-only cached loads, so the 16
// (8) points of a 128-byte line are served from L1 after the first touch -- measured 12 % faster on C2 than streaming.
template <bool STREAM>
__device__ __forceinline__ cplx ld_amp(const float2 *p) {
    const float2 v = STREAM ? __ldcs(p) : __ldg(p);
    return cplx{(double)v.x, (double)v.y};
}
template <bool STREAM>
__device__ __forceinline__ cplx ld_amp(const double2 *p) {
    const double2 v = STREAM ? __ldcs(p) : __ldg(p);
    return cplx{v.x, v.y};
}

// `c` points at the first control array of the pulse; arrays are `pts` apart.  j is the effective step.
template <typename IO, bool STREAM = true>
__device__ __forceinline__ cplx step_coefficient(const Term &t, const IO *__restrict__ c, unsigned int pts, int quad,
                                                 double magfac, unsigned long long j) {
    const IO *ca = c + (size_t)t.j * pts;
    if (t.type == TERM_PLAIN) {
        if (quad == QUAD_NONE) return ld_amp<STREAM>(ca + j);
        if (quad == QUAD_MIDPOINT) {
            const cplx u = ld_amp<STREAM>(ca + j), v = ld_amp<STREAM>(ca + j + 1);
            return cplx{0.5 * (u.re + v.re), 0.5 * (u.im + v.im)};
        }
        const cplx u = ld_amp<STREAM>(ca + 2 * j), v = ld_amp<STREAM>(ca + 2 * j + 1), w = ld_amp<STREAM>(ca + 2 * j + 2);
        // s / 6 as s * (r_hi + r_lo): a single pre-rounded 1/6 would bias every step the same way (relative 5.6e-17,
        // coherent over 1e6 steps), a true division costs ~30 FP64-pipe instructions.  r_hi + r_lo = 1/6 to 1e-33.
        constexpr double r_hi = 0.16666666666666666, r_lo = 9.251858538542970e-18;
        const double sr = (u.re + 4.0 * v.re) + w.re, si = (u.im + 4.0 * v.im) + w.im;
        return cplx{fma(sr, r_hi, sr * r_lo), fma(si, r_hi, si * r_lo)};
    }
    if (t.type == TERM_MAG_DRIFT) {
        const cplx u = ld_amp<STREAM>(ca + 2 * j), w = ld_amp<STREAM>(ca + 2 * j + 2);
        const double dr = w.re - u.re, di = w.im - u.im;
        return cplx{-di * magfac, dr * magfac};   // (w - u) * i * h/12
    }
    // TERM_MAG_PAIR
    const IO *cb = c + (size_t)t.k * pts;
    const cplx a0 = ld_amp<STREAM>(ca + 2 * j), a2 = ld_amp<STREAM>(ca + 2 * j + 2);
    const cplx b0 = ld_amp<STREAM>(cb + 2 * j), b2 = ld_amp<STREAM>(cb + 2 * j + 2);
    const double vr = (a0.re * b2.re - a0.im * b2.im) - (a2.re * b0.re - a2.im * b0.im);
    const double vi = (a0.re * b2.im + a0.im * b2.re) - (a2.re * b0.im + a2.im * b0.re);
    return cplx{-vi * magfac, vr * magfac};
}

}  // namespace pb
