// params.hpp -- plain-data kernel parameter blocks shared by host and device code.
#pragma once
#include "frag.cuh"

namespace pb {

constexpr int kMaxTerms = 64;    // effective control terms A' per step (controls + Magnus commutators)
constexpr int kMaxDegree = 63;   // Chebyshev degree MMAX

enum TermType : int {
    TERM_PLAIN = 0,      // quadrature-averaged amplitude of control j          (control_expansion.cu:105-160)
    TERM_MAG_DRIFT = 1,  // (c_j(t2) - c_j(t0)) * i h/12   on [H0, H_j]          (control_expansion.cu:43-46)
    TERM_MAG_PAIR = 2    // (c_j(t0) c_k(t2) - c_j(t2) c_k(t0)) * i h/12 on [H_j, H_k], j < k (control_expansion.cu:49-59)
};

enum QuadKind : int { QUAD_NONE = 0, QUAD_MIDPOINT = 1, QUAD_SIMPSON = 2 };

struct Term {
    int type;   // TermType
    int mat;    // index into the device matrix table (0 = H0)
    int j, k;   // control indices
};

struct SeriesParams {
    int n;                       // Hamiltonian dimension
    int npad;                    // padded dimension used by the kernel family
    int quad;                    // QuadKind (Magnus implies SIMPSON)
    int nterms;                  // A'
    int M;                       // Chebyshev degree actually evaluated
    int horner;                  // 0: Clenshaw on the Chebyshev coefficients; 1: Horner in W = Y^2 on the monomial
                                 // coefficients of the SAME polynomial (a[m] = c_m, p(Y) - I = sum_m c_m Y^m)
    unsigned int pts;            // raw points per control array
    unsigned int amps_in;        // control arrays per pulse in `carr`
    double sigma;                // 2 / Hnorm as rounded to double; x is derived from THIS value on the host
    double magfac;               // h / 12
    cplx a[kMaxDegree + 1];      // a[0] = J_0(x) - 1,  a[k] = (-i)^k J_k(x)          (rounded to double)
    cplx a_lo[kMaxDegree + 1];   // long-double remainder of a[k]: added to the diagonal BEFORE a[k], while the
                                 // accumulator is still small, so that the rounding of the series constants does
                                 // not bias every step the same way (DESIGN.md "Numerics")
    int pack;                    // dim <= 4: diagonal blocks of the 8 x 8 tile that advance through different parts of the step range (k1_warp.cu); else 1
    int herm;                    // H0 and every H_k are exactly Hermitian and no Magnus terms: a step with real coefficients has a Hermitian X
    int mixed;                   // dim 9..16, complex64, degree-8 form: the two small products of the series at fp32 grade (k1_warp.cu MIXED)
    float fconst[16];            // degree-8 form, mixed-precision kernel (k1_warp.cu MIXED): c4, d2, then hi / lo pairs of c3, d1, e2, e0
    Term terms[kMaxTerms];
};

}  // namespace pb
