// frag.cuh -- complex FP64 matrix fragments on the DMMA (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4) pipe.
//
// All arithmetic of this library is done in double precision on the FP64 tensor pipe, for complex64 and
// complex128 contexts alike (DESIGN.md "Numerics": a float leaf cannot hold 1e-5 over 5e5 steps).
// A warp owns whole (8*NT)x(8*NT) complex matrices in registers.  Two register layouts exist:
//
//   AccFrag  (accumulator / left-operand layout).  lane = 4*g + q, g = lane>>2, q = lane&3.
//            element (mt, nt, i)  <->  row 8*mt + g, column 8*nt + 2*q + i.
//            This is the C/D layout of m8n8k4.  Read column-slot-wise it is ALSO a legal A operand:
//            the A fragment of k-tile kt = 2*nt + i is the register (mt, nt, i), i.e. the contraction
//            index is enumerated in the permuted order  slot q of k-tile kt  <->  column 8*(kt>>1)+2*q+(kt&1).
//   BFrag    (right-operand layout) built for that same permuted contraction order:
//            element (kt, nt)  <->  row 8*(kt>>1) + 2*q + (kt&1), column 8*nt + g.
//
// Consequences used everywhere: (1) the result of  C = A*B  can be fed back as the left operand of
// the next product with no data movement (Clenshaw recurrence  B_k = B_{k+1}*Y - B_{k+2} + a_k I);
// (2) the transpose of an AccFrag matrix IS a BFrag of the same registers
// (BFrag(E^T)(kt, nt) == AccFrag(E)(nt, kt>>1, kt&1)), so the running product Q <- Q * U^T needs no
// shuffles either.  Q accumulates the transposed propagator.
#pragma once
#include <cuda_runtime.h>

namespace pb {

struct cplx { double re, im; };

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// -x on the integer pipe: the FP64 pipe is shared with DMMA, a DADD/DMUL spent on a sign flip is taken from it.
__device__ __forceinline__ double neg(double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

template <int NT>
struct AccFrag {
    double re[NT][NT][2];
    double im[NT][NT][2];
};

template <int NT>
struct BFrag {
    double re[2 * NT][NT];
    double im[2 * NT][NT];
    double nim[2 * NT][NT];   // -im: DMMA has no operand negation
};

// row / column of AccFrag element (mt, nt, i) for this lane
__device__ __forceinline__ int acc_row(int lane, int mt) { return 8 * mt + (lane >> 2); }
__device__ __forceinline__ int acc_col(int lane, int nt, int i) { return 8 * nt + 2 * (lane & 3) + i; }
// row / column of BFrag element (kt, nt) for this lane
__device__ __forceinline__ int bf_row(int lane, int kt) { return 8 * (kt >> 1) + 2 * (lane & 3) + (kt & 1); }
__device__ __forceinline__ int bf_col(int lane, int nt) { return 8 * nt + (lane >> 2); }

// C += A * B   (complex, 4 real DMMA products per tile triple)
template <int NT>
__device__ __forceinline__ void cmma(AccFrag<NT> &C, const AccFrag<NT> &A, const BFrag<NT> &B) {
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt) {
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) {
            const double are = A.re[mt][kt >> 1][kt & 1];
            const double aim = A.im[mt][kt >> 1][kt & 1];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                dmma884(C.re[mt][nt][0], C.re[mt][nt][1], are, B.re[kt][nt]);
                dmma884(C.im[mt][nt][0], C.im[mt][nt][1], are, B.im[kt][nt]);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                dmma884(C.re[mt][nt][0], C.re[mt][nt][1], aim, B.nim[kt][nt]);
                dmma884(C.im[mt][nt][0], C.im[mt][nt][1], aim, B.re[kt][nt]);
            }
        }
    }
}

// C += A * B from THREE real products per tile triple (P1 = Ar Br, P2 = Ai Bi, P3 = (Ar + Ai)(Br + Bi); Re = P1 - P2,
// Im = P3 - P1 - P2): 12 instead of 16 DMMAs per complex 8x8x8 block for a handful of DADDs.  B.nim must hold Br + Bi here
// (bfrag_third<true>).  The three accumulators start from zero and are combined at the end, so C is rounded once.
template <int NT>
__device__ __forceinline__ void cmma3(AccFrag<NT> &C, const AccFrag<NT> &A, const BFrag<NT> &B) {
    double p1[NT][NT][2], p2[NT][NT][2], p3[NT][NT][2];
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) { p1[mt][nt][i] = 0.0; p2[mt][nt][i] = 0.0; p3[mt][nt][i] = 0.0; }
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt) {
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) {
            const double are = A.re[mt][kt >> 1][kt & 1];
            const double aim = A.im[mt][kt >> 1][kt & 1];
            const double asum = are + aim;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                dmma884(p1[mt][nt][0], p1[mt][nt][1], are, B.re[kt][nt]);
                dmma884(p2[mt][nt][0], p2[mt][nt][1], aim, B.im[kt][nt]);
                dmma884(p3[mt][nt][0], p3[mt][nt][1], asum, B.nim[kt][nt]);
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                C.re[mt][nt][i] += p1[mt][nt][i] - p2[mt][nt][i];
                C.im[mt][nt][i] += p3[mt][nt][i] - (p1[mt][nt][i] + p2[mt][nt][i]);
            }
}

// C = A * B for a product known to be HERMITIAN (16 x 16, NT = 2; A = B = a Hermitian X): the lower-left 8 x 8 tile is not computed
// (36 instead of 48 DMMAs) but taken as the conjugate transpose of the upper-right one, a two-round warp shuffle on the crossbar
// (the exchange of acc_to_bfrag).  C is overwritten; B.nim holds Br + Bi as for cmma3.
__device__ __forceinline__ void cmma3_herm16(AccFrag<2> &C, const AccFrag<2> &A, const BFrag<2> &B, int lane) {
    double p1[2][2][2], p2[2][2][2], p3[2][2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) { p1[mt][nt][i] = 0.0; p2[mt][nt][i] = 0.0; p3[mt][nt][i] = 0.0; }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const double are = A.re[mt][kt >> 1][kt & 1];
            const double aim = A.im[mt][kt >> 1][kt & 1];
            const double asum = are + aim;
#pragma unroll
            for (int nt = mt; nt < 2; ++nt) {   // tiles (0,0), (0,1), (1,1)
                dmma884(p1[mt][nt][0], p1[mt][nt][1], are, B.re[kt][nt]);
                dmma884(p2[mt][nt][0], p2[mt][nt][1], aim, B.im[kt][nt]);
                dmma884(p3[mt][nt][0], p3[mt][nt][1], asum, B.nim[kt][nt]);
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = mt; nt < 2; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                C.re[mt][nt][i] = p1[mt][nt][i] - p2[mt][nt][i];
                C.im[mt][nt][i] = p3[mt][nt][i] - (p1[mt][nt][i] + p2[mt][nt][i]);
            }
    // C[1][0](g, 2q + i) = conj(C[0][1](2q + i, g)): element (g & 1) of lane (g' = 2q + i, q' = g >> 1) -- the exchange of acc_to_bfrag:
    // round 1 serves i = g & 1 (source and requester have the same parity), round 2 the other one
    const int g = lane >> 2, q = lane & 3;
    const bool odd = g & 1;
    const int a1 = 4 * (2 * q + (g & 1)) + (g >> 1);
    const int a2 = 4 * (2 * q + 1 - (g & 1)) + (g >> 1);
    const double r1 = __shfl_sync(0xffffffffu, odd ? C.re[0][1][1] : C.re[0][1][0], a1);
    const double r2 = __shfl_sync(0xffffffffu, odd ? C.re[0][1][0] : C.re[0][1][1], a2);
    const double i1 = __shfl_sync(0xffffffffu, odd ? C.im[0][1][1] : C.im[0][1][0], a1);
    const double i2 = __shfl_sync(0xffffffffu, odd ? C.im[0][1][0] : C.im[0][1][1], a2);
    C.re[1][0][0] = odd ? r2 : r1;  C.re[1][0][1] = odd ? r1 : r2;
    C.im[1][0][0] = neg(odd ? i2 : i1);  C.im[1][0][1] = neg(odd ? i1 : i2);
}

// third component of a right operand: -Bi for the four-product form (DMMA has no operand negation), Br + Bi for cmma3
template <bool MUL3>
__device__ __forceinline__ double bfrag_third(double re, double im) { return MUL3 ? re + im : neg(im); }

// BFrag of E^T from the registers of AccFrag E (no data movement, see header comment)
template <int NT>
__device__ __forceinline__ void transpose_as_bfrag(BFrag<NT> &B, const AccFrag<NT> &E) {
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            B.re[kt][nt] = E.re[nt][kt >> 1][kt & 1];
            B.im[kt][nt] = E.im[nt][kt >> 1][kt & 1];
            B.nim[kt][nt] = neg(E.im[nt][kt >> 1][kt & 1]);
        }
}

// BFrag of E^H (conjugate transpose) from the registers of AccFrag E: for a Hermitian matrix this IS its own right-operand
// layout -- no data movement.  B.nim is left to the caller.
template <int NT>
__device__ __forceinline__ void conj_transpose_as_bfrag(BFrag<NT> &B, const AccFrag<NT> &E) {
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            B.re[kt][nt] = E.re[nt][kt >> 1][kt & 1];
            B.im[kt][nt] = neg(E.im[nt][kt >> 1][kt & 1]);
        }
}

// BFrag of X from the AccFrag of the SAME matrix X (a change of layout, unlike transpose_as_bfrag) by warp shuffles.
// Per 8x8 block, B[par](g, q) = X[2q + par][g] is accumulator element i = g & 1 of lane (g' = 2q + par, q' = g >> 1).
// Two exchange rounds serve both parities: in round 1 the even-g lanes fetch par 0 and the odd-g lanes par 1, so every
// source lane is asked for its element i = g' & 1 only; round 2 is the other way round.  Shuffles and selects run on the
// crossbar / integer pipes, not on the FP64 pipe the DMMAs need; B.nim is left to the caller.
template <int NT>
__device__ __forceinline__ void acc_to_bfrag(BFrag<NT> &B, const AccFrag<NT> &X, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const bool odd = g & 1;
    const int src1 = 4 * (2 * q + (g & 1)) + (g >> 1);
    const int src2 = 4 * (2 * q + 1 - (g & 1)) + (g >> 1);
#pragma unroll
    for (int kb = 0; kb < NT; ++kb)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double r1 = __shfl_sync(0xffffffffu, odd ? X.re[kb][nt][1] : X.re[kb][nt][0], src1);
            const double r2 = __shfl_sync(0xffffffffu, odd ? X.re[kb][nt][0] : X.re[kb][nt][1], src2);
            const double i1 = __shfl_sync(0xffffffffu, odd ? X.im[kb][nt][1] : X.im[kb][nt][0], src1);
            const double i2 = __shfl_sync(0xffffffffu, odd ? X.im[kb][nt][0] : X.im[kb][nt][1], src2);
            B.re[2 * kb][nt] = odd ? r2 : r1;
            B.re[2 * kb + 1][nt] = odd ? r1 : r2;
            B.im[2 * kb][nt] = odd ? i2 : i1;
            B.im[2 * kb + 1][nt] = odd ? i1 : i2;
        }
}

template <int NT>
__device__ __forceinline__ void set_identity(AccFrag<NT> &Q, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                Q.re[mt][nt][i] = (mt == nt && g == 2 * q + i) ? 1.0 : 0.0;
                Q.im[mt][nt][i] = 0.0;
            }
}

template <int NT>
__device__ __forceinline__ void set_zero(AccFrag<NT> &Q) {
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) { Q.re[mt][nt][i] = 0.0; Q.im[mt][nt][i] = 0.0; }
}

// S <- alpha * S + (beta + beta_lo) * I     (alpha real; beta_lo, the sub-ulp remainder of beta, is added first)
template <int NT>
__device__ __forceinline__ void scale_add_diag(AccFrag<NT> &S, double alpha, cplx beta, cplx beta_lo, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool diag = (mt == nt && g == 2 * q + i);
                S.re[mt][nt][i] = (alpha * S.re[mt][nt][i] + (diag ? beta_lo.re : 0.0)) + (diag ? beta.re : 0.0);
                S.im[mt][nt][i] = (alpha * S.im[mt][nt][i] + (diag ? beta_lo.im : 0.0)) + (diag ? beta.im : 0.0);
            }
}

// S <- a * Y + (b + b_lo) * I   (a, b complex)
template <int NT>
__device__ __forceinline__ void axpb_diag(AccFrag<NT> &S, cplx a, const AccFrag<NT> &Y, cplx b, cplx b_lo, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool diag = (mt == nt && g == 2 * q + i);
                const double yr = Y.re[mt][nt][i], yi = Y.im[mt][nt][i];
                S.re[mt][nt][i] = ((a.re * yr - a.im * yi) + (diag ? b_lo.re : 0.0)) + (diag ? b.re : 0.0);
                S.im[mt][nt][i] = ((a.re * yi + a.im * yr) + (diag ? b_lo.im : 0.0)) + (diag ? b.im : 0.0);
            }
}

// Horner addend  S <- i (ci + ci_lo) Y + (cr + cr_lo) I   for REAL scalars ci, cr (the monomial coefficients of the
// series are purely imaginary for odd and purely real for even powers).  The sub-ulp remainders go in first so the
// dominant term is rounded once, by the final FMA (unbiased rounding of the constants, DESIGN.md "Numerics").
template <int NT, bool LO>
__device__ __forceinline__ void horner_addend(AccFrag<NT> &S, double ci, double ci_lo, const AccFrag<NT> &Y, double cr,
                                              double cr_lo, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool diag = (mt == nt && g == 2 * q + i);
                const double yr = Y.re[mt][nt][i], yi = Y.im[mt][nt][i];
                if (LO) {
                    double tr = fma(-ci_lo, yi, diag ? cr_lo : 0.0) + (diag ? cr : 0.0);
                    S.re[mt][nt][i] = fma(-ci, yi, tr);
                    S.im[mt][nt][i] = fma(ci, yr, ci_lo * yr);
                } else {
                    S.re[mt][nt][i] = fma(-ci, yi, diag ? cr : 0.0);
                    S.im[mt][nt][i] = ci * yr;
                }
            }
}

// ---- memory <-> fragment, matrices stored row-major as interleaved complex doubles with pitch ld ----
template <int NT>
__device__ __forceinline__ void load_acc(AccFrag<NT> &Q, const double2 *m, int ld, int lane) {
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double2 v = m[acc_row(lane, mt) * ld + acc_col(lane, nt, i)];
                Q.re[mt][nt][i] = v.x;
                Q.im[mt][nt][i] = v.y;
            }
}

template <int NT>
__device__ __forceinline__ void store_acc(const AccFrag<NT> &Q, double2 *m, int ld, int lane) {
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i)
                m[acc_row(lane, mt) * ld + acc_col(lane, nt, i)] = make_double2(Q.re[mt][nt][i], Q.im[mt][nt][i]);
}

template <int NT>
__device__ __forceinline__ void load_bfrag(BFrag<NT> &B, const double2 *m, int ld, int lane) {
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double2 v = m[bf_row(lane, kt) * ld + bf_col(lane, nt)];
            B.re[kt][nt] = v.x;
            B.im[kt][nt] = v.y;
            B.nim[kt][nt] = neg(v.y);
        }
}

}  // namespace pb
