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

// The up to four raw samples a term needs for effective step j, loaded without branches so that the loads can be issued
// ahead of the arithmetic (the register kernel issues them one product early):
//   PLAIN  NONE: c[j]            MIDPOINT: c[j], c[j+1]           SIMPSON: c[2j], c[2j+1], c[2j+2]
//   MAG_DRIFT:   c[2j], c[2j+2]  MAG_PAIR: a[2j], a[2j+2], b[2j], b[2j+2]
// Unused slots repeat an address that is read anyway.  `c` points at the first control array of the pulse; arrays are
// `pts` apart.
template <typename IO>
struct RawAmp { IO v[4]; };

template <typename IO, bool STREAM>
__device__ __forceinline__ IO ld_raw(const IO *p) { return STREAM ? __ldcs(p) : __ldg(p); }

template <typename IO, bool STREAM = true>
__device__ __forceinline__ RawAmp<IO> load_raw(const Term &t, const IO *__restrict__ c, unsigned int pts, int quad,
                                               unsigned long long j) {
    const bool plain = t.type == TERM_PLAIN, pair = t.type == TERM_MAG_PAIR;
    const IO *ca = c + (size_t)t.j * pts + ((plain && quad != QUAD_SIMPSON) ? j : 2 * j);
    const IO *cb = pair ? c + (size_t)t.k * pts + 2 * j : ca;
    const unsigned int o1 = plain ? (quad == QUAD_NONE ? 0u : 1u) : 2u;
    const unsigned int o2 = plain ? (quad == QUAD_NONE ? 0u : (quad == QUAD_MIDPOINT ? 1u : 2u)) : 0u;
    RawAmp<IO> r;
    r.v[0] = ld_raw<IO, STREAM>(ca);
    r.v[1] = ld_raw<IO, STREAM>(ca + o1);
    r.v[2] = ld_raw<IO, STREAM>(cb + o2);
    r.v[3] = ld_raw<IO, STREAM>(cb + (pair ? 2u : 0u));
    return r;
}

__device__ __forceinline__ cplx widen(const float2 v) { return cplx{(double)v.x, (double)v.y}; }
__device__ __forceinline__ cplx widen(const double2 v) { return cplx{v.x, v.y}; }

// ---- the arithmetic of the four term kinds (shared by the direct and the software-pipelined evaluation) ----
__device__ __forceinline__ cplx coef_midpoint(const cplx u, const cplx v) { return cplx{0.5 * (u.re + v.re), 0.5 * (u.im + v.im)}; }
__device__ __forceinline__ cplx coef_simpson(const cplx u, const cplx v, const cplx w) {
    // s / 6 as s * (r_hi + r_lo): a single pre-rounded 1/6 would bias every step the same way (relative 5.6e-17,
    // coherent over 1e6 steps), a true division costs ~30 FP64-pipe instructions.  r_hi + r_lo = 1/6 to 1e-33.
    constexpr double r_hi = 0.16666666666666666, r_lo = 9.251858538542970e-18;
    const double sr = (u.re + 4.0 * v.re) + w.re, si = (u.im + 4.0 * v.im) + w.im;
    return cplx{fma(sr, r_hi, sr * r_lo), fma(si, r_hi, si * r_lo)};
}
__device__ __forceinline__ cplx coef_mag_drift(const cplx u, const cplx w, double magfac) {
    const double dr = w.re - u.re, di = w.im - u.im;
    return cplx{-di * magfac, dr * magfac};   // (w - u) * i * h/12
}
__device__ __forceinline__ cplx coef_mag_pair(const cplx a0, const cplx a2, const cplx b0, const cplx b2, double magfac) {
    const double vr = (a0.re * b2.re - a0.im * b2.im) - (a2.re * b0.re - a2.im * b0.im);
    const double vi = (a0.re * b2.im + a0.im * b2.re) - (a2.re * b0.im + a2.im * b0.re);
    return cplx{-vi * magfac, vr * magfac};
}

// Effective coefficient of the term from its raw samples.
template <typename IO>
__device__ __forceinline__ cplx coef_from_raw(const Term &t, int quad, double magfac, const RawAmp<IO> &r) {
    if (t.type == TERM_PLAIN) {
        if (quad == QUAD_NONE) return widen(r.v[0]);
        if (quad == QUAD_MIDPOINT) return coef_midpoint(widen(r.v[0]), widen(r.v[1]));
        return coef_simpson(widen(r.v[0]), widen(r.v[1]), widen(r.v[2]));
    }
    if (t.type == TERM_MAG_DRIFT) return coef_mag_drift(widen(r.v[0]), widen(r.v[1]), magfac);
    return coef_mag_pair(widen(r.v[0]), widen(r.v[1]), widen(r.v[2]), widen(r.v[3]), magfac);
}

// Direct evaluation (loads only what the term kind needs).  j is the effective step.
template <typename IO, bool STREAM = true>
__device__ __forceinline__ cplx step_coefficient(const Term &t, const IO *__restrict__ c, unsigned int pts, int quad,
                                                 double magfac, unsigned long long j) {
    const IO *ca = c + (size_t)t.j * pts;
    auto ld = [](const IO *p) { return widen(ld_raw<IO, STREAM>(p)); };
    if (t.type == TERM_PLAIN) {
        if (quad == QUAD_NONE) return ld(ca + j);
        if (quad == QUAD_MIDPOINT) return coef_midpoint(ld(ca + j), ld(ca + j + 1));
        return coef_simpson(ld(ca + 2 * j), ld(ca + 2 * j + 1), ld(ca + 2 * j + 2));
    }
    if (t.type == TERM_MAG_DRIFT) return coef_mag_drift(ld(ca + 2 * j), ld(ca + 2 * j + 2), magfac);
    const IO *cb = c + (size_t)t.k * pts;
    return coef_mag_pair(ld(ca + 2 * j), ld(ca + 2 * j + 2), ld(cb + 2 * j), ld(cb + 2 * j + 2), magfac);
}

}  // namespace pb
