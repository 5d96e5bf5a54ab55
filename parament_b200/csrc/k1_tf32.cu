// k1_tf32.cu -- complex64 contexts, dim <= 8, short pulses: kernel (1) in FP32 arithmetic on the TF32 tensor path.
//
// The north_star asks for the complex64 branch to run below FP64 cost where the tolerance allows it.  Over the 5e5 steps
// of C2 no all-fp32 scheme holds 1e-5 (profiles/error_growth_tf32_r2.md: 1e-3 at 1e6 steps), but a GRAPE ensemble (C5) propagates 1e3 steps per pulse,
// and there the same three-product degree-8 evaluation as k1_warp.cu runs with
//   * every matrix product as a 3xTF32 split on the warp-level tensor path (mma.sync.m16n8k8.tf32, SASS HMMA.1688.F32.TF32;
//     measured 278 TFLOP/s on B200 = 92 TFLOP/s of fp32-grade products against 37 TFLOP/s of DMMA and 69 of FFMA):
//     a = a_hi + a_lo with both parts rounded to TF32, C += a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32 accumulation;
//     tcgen05.mma kind::tf32 cannot be used here: it needs M >= 64 operand tiles in shared memory, these matrices are
//     8 x 8 and live in registers;
//   * an 8 x 8 complex matrix Z held as the STACKED real 16 x 8 matrix [Re Z; Im Z] in the m16n8 accumulator layout
//     (four registers per lane): Z = A B is two real MMAs, [Ar; Ai] Br + [-Ai; Ar] Bi, and -- as in frag.cuh -- the
//     accumulator registers read column-slot-wise are a legal A operand (contraction index permuted: slot q <-> column 2q,
//     slot q + 4 <-> column 2q + 1) while the accumulator registers of Z^T are the B operand of Z for that same
//     permutation.  The running product Q <- Q + Q E^T therefore needs no data movement; the two right operands of the
//     series that are not transposes come from a 4-shuffle transposition;
//   * the low-order constants of the polynomial (r_0, r_1, r_2) carried as hi + lo pairs (their fp32 rounding would be a
//     coherent per-step error), every MMA started from a zero accumulator (the tensor core's truncating accumulation is a
//     coherent error too, proportional to the accumulator's magnitude), and optionally the running product kept as a
//     compensated pair Q_hi + Q_lo (TwoSum; measured without further effect, off by default).
// The host (api.cu tf32_candidate / use_tf32_path) selects this kernel only for complex64 contexts with dim <= 8, the degree-8
// form and an accumulated phase N h rho <= 128 -- the bound comes from the measured error law (linear in that phase), not from
// an argument.
// The warp's product is handed over in double (Q_hi + Q_lo); products across warps and CTAs and the fused final stage are the
// FP64 kernel's (k1_common.cuh).
#include <cstdlib>
#include "coef.cuh"
#include "k1_warp.hpp"
#include "k1_common.cuh"

namespace pb {

namespace {

struct F8 { float c[4]; };   // c0 = Re Z[g][2q], c1 = Re Z[g][2q+1], c2 = Im Z[g][2q], c3 = Im Z[g][2q+1]   (lane = 4 g + q)

// Round to TF32 (10 explicit mantissa bits), ties away from zero, on the integer pipe: add half an ulp of the kept part to the
// bit pattern and clear the 13 dropped bits.  Two instructions; `cvt.rna.tf32.f32` compiles to five on sm_100a (it also
// handles infinities, which a propagator never holds) and there are 52 roundings per step.
__device__ __forceinline__ unsigned tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float x, unsigned &hi, unsigned &lo) {
    hi = tf32_rna(x);
    // exact difference, |lo| <= 2^-11 |x|.  The tensor core reads the upper 19 bits of an operand register, i.e. truncates lo
    // to TF32: an error below 2^-21 |x| whose sign follows lo (random against x), so no rounding instruction is spent on it.
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct Op8 { unsigned h[4], l[4]; };   // TF32 split of the four registers of a matrix
__device__ __forceinline__ Op8 split8(const F8 &Z) {
    Op8 o;
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(Z.c[i], o.h[i], o.l[i]);
    return o;
}

// The tensor core accumulates with truncation (round toward zero), a systematic shrink of every result by ~2^-25 relative to
// the magnitude of the ACCUMULATOR -- over a pulse that acts like a coherent rescaling of the time axis (measured: the error of
// this kernel grows linearly in N, profiles/error_growth_tf32_r2.md).  So every MMA (or short chain of small cross terms) starts
// from a zero accumulator, keeping each truncation at the magnitude of its own product, and the pieces are combined with
// round-to-nearest FADDs; the independent accumulators also give the tensor pipe four MMAs in flight per warp.
// Both real MMAs of a product use the SAME A operand [Ar; Ai] (a0 = (g, slot q) = c0, a1 = (g + 8, slot q) = c2,
// a2 = (g, slot q + 4) = c1, a3 = c3), once against Br = (c0, c1) of B^T and once against Bi = (c2, c3):
//     M_re = [Ar Br; Ai Br],  M_im = [Ar Bi; Ai Bi]   =>   Re Z = top(M_re) - bottom(M_im),  Im Z = bottom(M_re) + top(M_im),
// and top / bottom of an accumulator are registers (c0, c1) / (c2, c3) of the same lane: the recombination is four FADDs and no
// operand is negated or permuted.
__device__ __forceinline__ void mma_re(float (&d)[4], const unsigned (&x)[4], const unsigned (&y)[4]) {
    mma_tf32(d, x[0], x[2], x[1], x[3], y[0], y[1]);
}
__device__ __forceinline__ void mma_im(float (&d)[4], const unsigned (&x)[4], const unsigned (&y)[4]) {
    mma_tf32(d, x[0], x[2], x[1], x[3], y[2], y[3]);
}

// A * B (complex 8 x 8) at fp32 grade: A = split of the left matrix, Bt = split of the TRANSPOSE of the right matrix.
// Six TF32 MMAs in four independent accumulators.
__device__ __forceinline__ F8 cmul(const Op8 &A, const Op8 &Bt) {
    float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f}, m3[4] = {0.f, 0.f, 0.f, 0.f}, m4[4] = {0.f, 0.f, 0.f, 0.f};
    mma_re(m1, A.h, Bt.h);
    mma_im(m2, A.h, Bt.h);
    mma_re(m3, A.l, Bt.h);
    mma_im(m4, A.l, Bt.h);
    mma_re(m3, A.h, Bt.l);
    mma_im(m4, A.h, Bt.l);
    F8 D;
    D.c[0] = (m1[0] - m2[2]) + (m3[0] - m4[2]);
    D.c[1] = (m1[1] - m2[3]) + (m3[1] - m4[3]);
    D.c[2] = (m1[2] + m2[0]) + (m3[2] + m4[0]);
    D.c[3] = (m1[3] + m2[1]) + (m3[3] + m4[1]);
    return D;
}
__device__ __forceinline__ F8 cmul(const F8 &A, const F8 &Bt) { return cmul(split8(A), split8(Bt)); }
// single-TF32 product (operands of ~1e-6 relative weight)
__device__ __forceinline__ F8 cmul_hi(const unsigned (&a)[4], const unsigned (&bt)[4]) {
    float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
    mma_re(m1, a, bt);
    mma_im(m2, a, bt);
    F8 D;
    D.c[0] = m1[0] - m2[2]; D.c[1] = m1[1] - m2[3]; D.c[2] = m1[2] + m2[0]; D.c[3] = m1[3] + m2[1];
    return D;
}

// Registers of Z^T from the registers of Z: element Z[2q + i][g] is register (g & 1) of lane (g' = 2q + i, q' = g >> 1); two
// exchange rounds per real block serve both i (same scheme as acc_to_bfrag, frag.cuh).
__device__ __forceinline__ F8 transpose8(const F8 &Z, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const bool odd = g & 1;
    const int src1 = 4 * (2 * q + (g & 1)) + (g >> 1);
    const int src2 = 4 * (2 * q + 1 - (g & 1)) + (g >> 1);
    F8 T;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float r1 = __shfl_sync(0xffffffffu, odd ? Z.c[2 * h + 1] : Z.c[2 * h], src1);
        const float r2 = __shfl_sync(0xffffffffu, odd ? Z.c[2 * h] : Z.c[2 * h + 1], src2);
        T.c[2 * h] = odd ? r2 : r1;
        T.c[2 * h + 1] = odd ? r1 : r2;
    }
    return T;
}

struct Split { float hi, lo; };
__device__ __forceinline__ Split split_const(double v) {
    Split s;
    s.hi = (float)v;
    s.lo = (float)(v - (double)s.hi);
    return s;
}

constexpr int KREG = 3;   // control terms whose fragment tables stay in registers (plus H0); further terms are read from L1 per step

}  // namespace

// Hfrag: the table of k1_warp.cu for NT = 1: [matrix][layout 0 = registers of Z, 1 = registers of Z^T][element i < 2][lane] double2
// (exact fp32 values: the context is complex64).  One partial (double, row-major 8 x 8, transposed propagator) per warp.
// COMP: the running product is a compensated pair Q_hi + Q_lo -- the fp32 rounding of Q_hi + Q_hi E^T is recovered exactly
// (TwoSum) into Q_lo, and Q_lo is propagated with the same step (a single-TF32 product suffices for a 1e-6-sized matrix:
// two more MMAs), because an error made at step j has to rotate with all later steps like the product itself.
// QK: QuadKind when every term is a plain control (no Magnus): the coefficient evaluation is then straight-line code;
// -1: generic (term kinds and quadrature decided at run time).
// HERM: H0 and every H_k are exactly Hermitian (api.cu set_hamiltonian) and there are no Magnus terms: the registers of Z^T are
// the conjugates of those of Z, so only one table per matrix is kept; a step with real coefficients has a Hermitian X and W, whose
// transposes cost a sign flip instead of four shuffles.
template <int OCC, bool COMP, int QK, bool HERM>
__global__ void __launch_bounds__(32 * K1_WARPS, OCC)
k1_tf32_chain_kernel(const SeriesParams p, const float2 *__restrict__ carr, const double2 *__restrict__ Hfrag,
                     double2 *__restrict__ partials, unsigned int batch, unsigned int chunks_per_pulse,
                     unsigned long long step_lo, unsigned long long step_hi, int reduce_in_cta, const K1Final fz) {
    __shared__ double2 smem[(K1_WARPS / 2) * 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const unsigned long long gw = (unsigned long long)blockIdx.x * K1_WARPS + warp;
    const unsigned int pulse = (unsigned int)(gw / chunks_per_pulse);
    const unsigned int chunk = (unsigned int)(gw % chunks_per_pulse);
    const bool active = pulse < batch;

    const unsigned long long nsteps = step_hi - step_lo;
    const unsigned long long lo = step_lo + nsteps * chunk / chunks_per_pulse;
    const unsigned long long hi = active ? step_lo + nsteps * (chunk + 1) / chunks_per_pulse : lo;   // inactive warps walk no step
    const float2 *c = carr + (size_t)(active ? pulse : 0) * p.amps_in * p.pts;
    const bool d0 = (g == 2 * q), d1 = (g == 2 * q + 1);   // this lane's c0 / c1 sits on the diagonal

    // constants of the degree-8 form (api.cu solve_degree8): p.a[k].re = c4 c3 d2 d1 e2 e0 r2' r1 r0
    const float c4 = (float)p.a[0].re, c3 = (float)p.a[1].re, d2 = (float)p.a[2].re, dd1 = (float)p.a[3].re;
    const float e2 = (float)p.a[4].re, e0 = (float)p.a[5].re;
    // r2' = r2 - e0 e2 has to cancel the e0 e2 W term of the product EXACTLY: rederived from e0 and e2 as rounded to float
    // (p.a[6] holds r2 minus the product of the double-precision parameters)
    const Split r2 = split_const(p.a[6].re + (p.a[5].re * p.a[4].re - (double)e0 * (double)e2));
    const Split r1 = split_const(p.a[7].re + p.a_lo[7].re), r0 = split_const(p.a[8].re + p.a_lo[8].re);

    // fragment tables of H0 and of the first KREG terms, converted once
    auto load_frag = [&](int mat, F8 &Z, F8 &Zt) {
        const double2 *H = Hfrag + (size_t)mat * 4 * 32 + lane;
        const double2 a0 = __ldg(H), a1 = __ldg(H + 32), b0 = __ldg(H + 64), b1 = __ldg(H + 96);
        Z.c[0] = (float)a0.x; Z.c[1] = (float)a1.x; Z.c[2] = (float)a0.y; Z.c[3] = (float)a1.y;
        Zt.c[0] = (float)b0.x; Zt.c[1] = (float)b1.x; Zt.c[2] = (float)b0.y; Zt.c[3] = (float)b1.y;
    };
    F8 H0, H0t, Hk[KREG], Hkt[HERM ? 1 : KREG];
    load_frag(0, H0, H0t);
#pragma unroll
    for (int t = 0; t < KREG; ++t)
        if (t < p.nterms) { if (HERM) { F8 unused; load_frag(p.terms[t].mat, Hk[t], unused); } else load_frag(p.terms[t].mat, Hk[t], Hkt[t]); }

    // effective coefficient of a term in single precision (its rounding is pseudo-random from step to step; control_expansion.cu
    // of the reference does the same arithmetic in float for complex64 contexts)
    auto coefficient = [&](const Term &t, unsigned long long j) -> float2 {
        if (QK >= 0) {
            const float2 *ca = c + (size_t)t.j * p.pts;
            if (QK == QUAD_NONE) return __ldg(ca + j);
            if (QK == QUAD_MIDPOINT) {
                const float2 u = __ldg(ca + j), v = __ldg(ca + j + 1);
                return make_float2(0.5f * (u.x + v.x), 0.5f * (u.y + v.y));
            }
            const float2 u = __ldg(ca + 2 * j), v = __ldg(ca + 2 * j + 1), w = __ldg(ca + 2 * j + 2);
            return make_float2(((u.x + 4.f * v.x) + w.x) * (1.f / 6.f), ((u.y + 4.f * v.y) + w.y) * (1.f / 6.f));
        }
        const cplx ct = step_coefficient<float2, false>(t, c, p.pts, p.quad, p.magfac, j);
        return make_float2((float)ct.re, (float)ct.im);
    };
    auto add_term = [&](F8 &X, F8 &Xt, const float2 ct, const F8 &Z, const F8 &Zt) {
        const float cr = ct.x, ci = ct.y;
        if (HERM) {   // X only; X^T is formed after the last term (conjugate of X, or from the conjugated tables)
            if (ci == 0.f) {
#pragma unroll
                for (int i = 0; i < 4; ++i) X.c[i] = fmaf(cr, Z.c[i], X.c[i]);
            } else {      // Z^T = conj(Z):  X^T += c conj(Z)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    X.c[i] = fmaf(cr, Z.c[i], fmaf(-ci, Z.c[2 + i], X.c[i]));
                    X.c[2 + i] = fmaf(cr, Z.c[2 + i], fmaf(ci, Z.c[i], X.c[2 + i]));
                    Xt.c[i] = fmaf(2.f * ci, Z.c[2 + i], Xt.c[i]);          // corrections to conj(X), applied below
                    Xt.c[2 + i] = fmaf(2.f * ci, Z.c[i], Xt.c[2 + i]);
                }
            }
        } else if (ci == 0.f) {   // real amplitude (warp-uniform)
#pragma unroll
            for (int i = 0; i < 4; ++i) { X.c[i] = fmaf(cr, Z.c[i], X.c[i]); Xt.c[i] = fmaf(cr, Zt.c[i], Xt.c[i]); }
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                X.c[i] = fmaf(cr, Z.c[i], fmaf(-ci, Z.c[2 + i], X.c[i]));
                X.c[2 + i] = fmaf(cr, Z.c[2 + i], fmaf(ci, Z.c[i], X.c[2 + i]));
                Xt.c[i] = fmaf(cr, Zt.c[i], fmaf(-ci, Zt.c[2 + i], Xt.c[i]));
                Xt.c[2 + i] = fmaf(cr, Zt.c[2 + i], fmaf(ci, Zt.c[i], Xt.c[2 + i]));
            }
        }
    };

    F8 Qh, Ql;   // running product (transposed propagator), compensated pair
    Qh.c[0] = d0 ? 1.f : 0.f; Qh.c[1] = d1 ? 1.f : 0.f; Qh.c[2] = Qh.c[3] = 0.f;
    Ql.c[0] = Ql.c[1] = Ql.c[2] = Ql.c[3] = 0.f;

    for (unsigned long long j = lo; j < hi; ++j) {
        // ---- X = H0 + sum_t c_t H_t, in the registers of X and of X^T ----
        F8 X = H0, Xt = H0t;
        bool xherm = HERM;   // X is Hermitian: Hermitian tables and real coefficients in this step (warp-uniform)
        if (HERM) Xt.c[0] = Xt.c[1] = Xt.c[2] = Xt.c[3] = 0.f;
#pragma unroll
        for (int t = 0; t < KREG; ++t)
            if (t < p.nterms) {
                const float2 ct = coefficient(p.terms[t], j);
                xherm = xherm && ct.y == 0.f;
                add_term(X, Xt, ct, Hk[t], Hkt[HERM ? 0 : t]);
            }
        for (int t = KREG; t < p.nterms; ++t) {
            F8 Z, Zt;
            load_frag(p.terms[t].mat, Z, Zt);
            const float2 ct = coefficient(p.terms[t], j);
            xherm = xherm && ct.y == 0.f;
            add_term(X, Xt, ct, Z, Zt);
        }
        if (HERM) {   // X^T = conj(X) (+ the corrections of complex coefficients collected in Xt)
            if (xherm) { Xt.c[0] = X.c[0]; Xt.c[1] = X.c[1]; Xt.c[2] = -X.c[2]; Xt.c[3] = -X.c[3]; }
            else { Xt.c[0] += X.c[0]; Xt.c[1] += X.c[1]; Xt.c[2] -= X.c[2]; Xt.c[3] -= X.c[3]; }
        }
        // ---- W = X X ----
        const F8 W = cmul(X, Xt);
        F8 Wt;
        if (xherm) { Wt.c[0] = W.c[0]; Wt.c[1] = W.c[1]; Wt.c[2] = -W.c[2]; Wt.c[3] = -W.c[3]; }   // W is Hermitian with X (to rounding)
        else Wt = transpose8(W, lane);
        // ---- y02 = (c4 W + i c3 X) W ----
        F8 T;
        T.c[0] = fmaf(-c3, X.c[2], c4 * W.c[0]); T.c[1] = fmaf(-c3, X.c[3], c4 * W.c[1]);
        T.c[2] = fmaf(c3, X.c[0], c4 * W.c[2]);  T.c[3] = fmaf(c3, X.c[1], c4 * W.c[3]);
        const F8 Y2 = cmul(T, Wt);
        // ---- E = (y02 - d2 W - i d1 X + e0 I)(y02 - e2 W) - r2' W - i r1 X + r0 I;  the sub-ulp parts of r go in first ----
        F8 E;
        E.c[0] = fmaf(-r2.lo, W.c[0], r1.lo * X.c[2]) + (d0 ? r0.lo : 0.f);
        E.c[1] = fmaf(-r2.lo, W.c[1], r1.lo * X.c[3]) + (d1 ? r0.lo : 0.f);
        E.c[2] = fmaf(-r2.lo, W.c[2], -r1.lo * X.c[0]);
        E.c[3] = fmaf(-r2.lo, W.c[3], -r1.lo * X.c[1]);
        E.c[0] = fmaf(r1.hi, X.c[2], fmaf(-r2.hi, W.c[0], E.c[0] + (d0 ? r0.hi : 0.f)));
        E.c[1] = fmaf(r1.hi, X.c[3], fmaf(-r2.hi, W.c[1], E.c[1] + (d1 ? r0.hi : 0.f)));
        E.c[2] = fmaf(-r1.hi, X.c[0], fmaf(-r2.hi, W.c[2], E.c[2]));
        E.c[3] = fmaf(-r1.hi, X.c[1], fmaf(-r2.hi, W.c[3], E.c[3]));
        F8 Rt = transpose8(Y2, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) Rt.c[i] = fmaf(-e2, Wt.c[i], Rt.c[i]);
        F8 L;
        L.c[0] = fmaf(dd1, X.c[2], fmaf(-d2, W.c[0], Y2.c[0])) + (d0 ? e0 : 0.f);
        L.c[1] = fmaf(dd1, X.c[3], fmaf(-d2, W.c[1], Y2.c[1])) + (d1 ? e0 : 0.f);
        L.c[2] = fmaf(-dd1, X.c[0], fmaf(-d2, W.c[2], Y2.c[2]));
        L.c[3] = fmaf(-dd1, X.c[1], fmaf(-d2, W.c[3], Y2.c[3]));
        {
            const F8 LR = cmul(L, Rt);
#pragma unroll
            for (int i = 0; i < 4; ++i) E.c[i] += LR.c[i];
        }
        // ---- running product  Q <- Q + Q E^T  (the registers of E are the B operand of E^T), compensated ----
        const Op8 Es = split8(E);
        if (COMP) {
            const F8 P = cmul(split8(Qh), Es);
            unsigned qlh[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qlh[i] = tf32_rna(Ql.c[i]);
            const F8 Pl = cmul_hi(qlh, Es.h);                  // Q_lo <- Q_lo + Q_lo E^T
#pragma unroll
            for (int i = 0; i < 4; ++i) Ql.c[i] += Pl.c[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float a = Qh.c[i], b = P.c[i];
                const float s = a + b;
                const float bb = s - a;
                const float err = (a - (s - bb)) + (b - bb);   // TwoSum: a + b = s + err exactly
                Qh.c[i] = s;
                Ql.c[i] += err;
            }
        } else {
            const F8 P = cmul(split8(Qh), Es);
#pragma unroll
            for (int i = 0; i < 4; ++i) Qh.c[i] += P.c[i];
        }
    }

    // ---- hand over in double: the ordered products across warps and CTAs run on the DMMA pipe like the FP64 kernel's ----
    AccFrag<1> Qd;
    Qd.re[0][0][0] = (double)Qh.c[0] + (double)Ql.c[0]; Qd.re[0][0][1] = (double)Qh.c[1] + (double)Ql.c[1];
    Qd.im[0][0][0] = (double)Qh.c[2] + (double)Ql.c[2]; Qd.im[0][0][1] = (double)Qh.c[3] + (double)Ql.c[3];
    k1_tail<1, float2>(Qd, active, pulse, chunk, chunks_per_pulse, reduce_in_cta, partials, fz, smem, warp, lane);
}

int k1_tf32_ctas_per_sm() { return 6; }

template <bool COMP, int QK, bool HERM>
static cudaError_t launch_tf32_t(const SeriesParams &p, const float2 *carr, const double2 *Hfrag, double2 *partials, unsigned int batch,
                                 const K1Plan &plan, unsigned long long step_lo, unsigned long long step_hi, const K1Final &fz, cudaStream_t stream) {
    k1_tf32_chain_kernel<6, COMP, QK, HERM><<<plan.grid, 32 * K1_WARPS, 0, stream>>>(p, carr, Hfrag, partials, batch, plan.chunks_per_pulse,
                                                                                step_lo, step_hi, plan.reduce_in_cta, fz);
    return cudaGetLastError();
}

cudaError_t launch_k1_tf32_chain(const SeriesParams &p, const void *carr, const double2 *Hfrag, double2 *partials, unsigned int batch,
                                 const K1Plan &plan, unsigned long long step_lo, unsigned long long step_hi, const K1Final &fz,
                                 cudaStream_t stream) {
    if (p.npad != 8 || p.horner != 3) return cudaErrorInvalidValue;
    // compensated running product: measured without effect on the error once every MMA starts from a zero accumulator
    // (profiles/error_growth_tf32_r2.md), so it is off unless $PARAMENT_TF32_COMP=1 asks for it (A/B runs)
    static const bool comp = getenv("PARAMENT_TF32_COMP") && atoi(getenv("PARAMENT_TF32_COMP")) == 1;
    bool plain = true;
    for (int t = 0; t < p.nterms; ++t) plain = plain && p.terms[t].type == TERM_PLAIN;
    const int qk = plain ? p.quad : -1;
    const float2 *cf = (const float2 *)carr;
    const bool herm = p.herm != 0 && plain;
#define PB_TF32_CASE(C, Q) if (comp == C && qk == Q) return launch_tf32_t<C, Q, false>(p, cf, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
#define PB_TF32_HERM(Q) if (!comp && herm && qk == Q) return launch_tf32_t<false, Q, true>(p, cf, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
    PB_TF32_HERM(QUAD_NONE); PB_TF32_HERM(QUAD_MIDPOINT); PB_TF32_HERM(QUAD_SIMPSON);
    PB_TF32_CASE(true, QUAD_NONE); PB_TF32_CASE(true, QUAD_MIDPOINT); PB_TF32_CASE(true, QUAD_SIMPSON); PB_TF32_CASE(true, -1);
    PB_TF32_CASE(false, QUAD_NONE); PB_TF32_CASE(false, QUAD_MIDPOINT); PB_TF32_CASE(false, QUAD_SIMPSON); PB_TF32_CASE(false, -1);
#undef PB_TF32_CASE
#undef PB_TF32_HERM
    return cudaErrorInvalidValue;
}

}  // namespace pb
