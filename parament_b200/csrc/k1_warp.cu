// k1_warp.cu -- kernels (1)-(3) of the north_star for dim <= 16: everything register-resident.
//
//  k1_chain_kernel   one warp walks a contiguous chunk of effective time steps.  Per step it assembles
//                    Y = sigma * (H0 + sum_t c_t H_t) from the raw amplitude stream (quadrature / Magnus
//                    coefficients evaluated on the fly, coef.cuh), evaluates the Chebyshev/Bessel series -- as a
//                    degree-8 / degree-12 polynomial in three / four matrix products, by Horner in Y^2, or by the
//                    reference's Clenshaw recurrence -- with every matrix product on the FP64 tensor pipe
//                    (frag.cuh: no shared memory, nothing written to HBM; complex products from three real ones
//                    in the degree-8 form), and multiplies the step into the warp's running product.   [kernel 1]
//                    MIXED (complex64, dim 9..16): the two small products of the degree-8 form run at fp32 grade
//                    as 3xTF32 on the other tensor sub-pipe (frag_tf32.cuh).
//                    dim <= 4: two or four systems share the 8 x 8 tile as diagonal blocks, each block advancing through its
//                    own part of the warp's step range (p.pack; unpacked by k1_common.cuh unpack_blocks after the loop).
//                    HERMK (dim 9..16, Hermitian input matrices): a step with real coefficients has a Hermitian X, whose
//                    right-operand layout -- and that of W = X X -- is a relabeling of registers instead of a warp shuffle,
//                    and whose square needs three of its four tiles from the tensor pipe (frag.cuh cmma3_herm16).
//                    The warps of a CTA then combine their chunk products in order through shared memory; the last
//                    CTAs to finish reduce the per-CTA partials and write the propagator (k1_common.cuh:
//                    ONE launch per call), or the partials are left to k3_reduce_kernel.         [kernels 2, 3]
//  k3_reduce_kernel  ordered reduction of stored partials of each pulse (host-pointer calls whose time axis is cut
//                    into copy groups), transposition to the row-major physical propagator and conversion to
//                    the context precision.                                                       [kernel 3]
//  k3_combine_kernel ordered product of the partial propagators of time slices (multi-GPU), one launch.
//
// Replaces parament.cpp:486-718 (equipropExpand / equipropPropagate / equipropReduce) and
// control_expansion.cu / diagonal_add.cu for these dimensions.  The running product is kept transposed,
// Q = (U_{hi-1} ... U_lo)^T = U_lo^T ... U_{hi-1}^T, see frag.cuh.
#include "coef.cuh"
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include "k1_warp.hpp"
#include "k1_common.cuh"
#include "frag_tf32.cuh"

namespace pb {

// Per-phase cycle counters (development aid, -DPB_PHASE_TIMING): warp 0 of block 0 prints its accumulated clock64() deltas.
#ifdef PB_PHASE_TIMING
#define K1_T_DECL long long pt_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pt_last_ = clock64();
#define K1_T(i) { const long long pt_now_ = clock64(); pt_[i] += pt_now_ - pt_last_; pt_last_ = pt_now_; }
#define K1_T_PRINT if (blockIdx.x == 0 && threadIdx.x == 0) printf("k1 phase cycles: assemble %lld  square %lld  horner %lld  chain %lld  other %lld  | mixed: S+convert+T %lld  W layout %lld  y02 %lld  R,L' %lld  steps %llu\n", pt_[0], pt_[1], pt_[2], pt_[3], pt_[4], pt_[5], pt_[6], pt_[7], pt_[8], hi - lo);
#else
#define K1_T_DECL
#define K1_T(i)
#define K1_T_PRINT
#endif


// Hfrag layout: [matrix][layout 0 = AccFrag order, 1 = BFrag order][element e < 2*NT*NT][lane] as double2.
// OCC: CTAs per SM the register allocation is bounded for (NT == 2: 2 -> 255 registers, 3 -> 168 registers with spills)
// MUL3 (degree-8 three-product form, complex64 contexts): complex products from three real ones (frag.cuh cmma3).
// MIXED (dim 9..16, complex64 contexts, degree-8 form): the two products of the series whose results are small -- y02 = T W
// (~1e-5) and L' R (~1e-4, L' = L - e0 I) -- run at fp32 grade as 3xTF32 on the warp-level tensor path (frag_tf32.cuh), which
// is a different pipe from the FP64 one; W = X X, every term of first and second order, and the running product stay in FP64.
// The FP64 pipe, which bounds this kernel, then carries two matrix products per step instead of four.
// HERMK: compiled with the Hermitian shortcuts; launched when api.cu set_hamiltonian found every matrix Hermitian (p.herm).
template <int NT, typename IO, int HORNER, int OCC, bool MUL3 = false, bool MIXED = false, bool HERMK = false>
__global__ void __launch_bounds__(32 * K1_WARPS, OCC)
k1_chain_kernel(const SeriesParams p, const IO *__restrict__ carr, const double2 *__restrict__ Hfrag,
                double2 *__restrict__ partials, unsigned int batch, unsigned int chunks_per_pulse,
                unsigned long long step_lo, unsigned long long step_hi, int reduce_in_cta, const K1Final fz) {
    constexpr int NP = 8 * NT;
    constexpr int NE = 2 * NT * NT;   // fragment elements per lane and layout
    __shared__ double2 smem[(K1_WARPS / 2) * NP * NP];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned long long gw = (unsigned long long)blockIdx.x * K1_WARPS + warp;
    const unsigned int pulse = (unsigned int)(gw / chunks_per_pulse);
    const unsigned int chunk = (unsigned int)(gw % chunks_per_pulse);
    const bool active = pulse < batch;

    AccFrag<NT> Q;
    set_identity<NT>(Q, lane);

    if (active) {
        const unsigned long long nsteps = step_hi - step_lo;
        const unsigned long long lo = step_lo + nsteps * chunk / chunks_per_pulse;
        const unsigned long long hi = step_lo + nsteps * (chunk + 1) / chunks_per_pulse;
        const IO *c = carr + (size_t)pulse * p.amps_in * p.pts;
        const double2 *HA = Hfrag + lane;             // + (mat * 2 + 0) * NE * 32 + e * 32
        const int M = p.M;

        // Coefficients of the first KPRE terms are software-pipelined: their raw samples for step j + 1 are requested
        // before the running-product DMMAs of step j and turned into coefficients after them, so the load -> convert ->
        // quadrature latency chain (measured: 13 % of the step at dim 16, 22 % at dim 8 when exposed) runs under tensor work.
        constexpr int KPRE = NT == 2 ? 2 : 0;   // (slower at dim <= 8, where six warps per scheduler hide the latency anyway)
        constexpr bool BOTH = HORNER != 3 || NT == 1;   // assemble the right-operand layout too; else it is shuffled (faster at dim 16 only)
        cplx ctn[KPRE > 0 ? KPRE : 1];
#pragma unroll
        for (int t = 0; t < KPRE; ++t) {
            ctn[t] = cplx{0.0, 0.0};
            if (t < p.nterms && lo < hi) ctn[t] = step_coefficient<IO, false>(p.terms[t], c, p.pts, p.quad, p.magfac, lo);
        }
        auto add_term = [&](AccFrag<NT> &Ya, BFrag<NT> &Yb, const cplx ct, const double2 *Ht) {
            if (ct.im == 0.0) {   // real amplitude (warp-uniform): half the FP64-pipe work of the assembly
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const double2 ha = __ldg(Ht + e * 32);
                    (&Ya.re[0][0][0])[e] = fma(ct.re, ha.x, (&Ya.re[0][0][0])[e]);
                    (&Ya.im[0][0][0])[e] = fma(ct.re, ha.y, (&Ya.im[0][0][0])[e]);
                    if (BOTH) {
                        const double2 hb = __ldg(Ht + (NE + e) * 32);
                        (&Yb.re[0][0])[e] = fma(ct.re, hb.x, (&Yb.re[0][0])[e]);
                        (&Yb.im[0][0])[e] = fma(ct.re, hb.y, (&Yb.im[0][0])[e]);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const double2 ha = __ldg(Ht + e * 32);
                    (&Ya.re[0][0][0])[e] += ct.re * ha.x - ct.im * ha.y;
                    (&Ya.im[0][0][0])[e] += ct.re * ha.y + ct.im * ha.x;
                    if (BOTH) {
                        const double2 hb = __ldg(Ht + (NE + e) * 32);
                        (&Yb.re[0][0])[e] += ct.re * hb.x - ct.im * hb.y;
                        (&Yb.im[0][0])[e] += ct.re * hb.y + ct.im * hb.x;
                    }
                }
            }
        };

        // Packed small systems (dim <= 4, p.pack = 4 or 2; api.cu upload_matrices stores kron(I_pack, H)): the 8 x 8 tile carries
        // `pack` diagonal blocks of nb = 8 / pack rows, and block b advances through the b-th quarter (half) of the warp's step
        // range -- every element a lane holds in any of the three register layouts lies in block (lane >> 2) / nb, so the time
        // index simply becomes lane-dependent.  Everything between the assembly and the running product is a polynomial in X
        // plus multiples of I, which keeps the blocks apart; a block that runs out of steps early gets E = 0 (its lanes hold all of it).
        const int pack = NT == 1 ? p.pack : 1;
        unsigned long long jb = lo, hb = hi, iters = hi - lo;
        if (NT == 1 && pack > 1) {
            const unsigned long long L = hi - lo;
            const unsigned int blk = (unsigned int)(lane >> 2) / (unsigned int)(8 / pack);
            jb = lo + L * blk / pack;
            hb = lo + L * (blk + 1) / pack;
            iters = (L + pack - 1) / pack;
        }
        K1_T_DECL
        for (unsigned long long it = 0; it < iters; ++it, ++jb) {
            const bool live = jb < hb;
            const unsigned long long j = live ? jb : hi - 1;
            K1_T(4)
            // ---- assemble X = H0 + sum_t c_t H_t in both register layouts, then Y = sigma X ----
            AccFrag<NT> Ya;
            BFrag<NT> Yb;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const double2 ha = __ldg(HA + e * 32);
                (&Ya.re[0][0][0])[e] = ha.x; (&Ya.im[0][0][0])[e] = ha.y;
                if (BOTH) {
                    const double2 hb = __ldg(HA + (NE + e) * 32);
                    (&Yb.re[0][0])[e] = hb.x;    (&Yb.im[0][0])[e] = hb.y;
                }
            }
            // Hermitian H0 / H_k (p.herm, checked on the host) and real coefficients make X Hermitian: its right-operand layout is
            // then the conjugate of the transpose-as-B-operand relabeling, with no shuffles (HERM_OK forms only)
            bool xherm = HERMK && p.herm != 0;   // HERMK: variant compiled with the Hermitian shortcuts (contexts whose matrices are Hermitian)
#pragma unroll
            for (int t = 0; t < KPRE; ++t)
                if (t < p.nterms) {
                    xherm = xherm && ctn[t].im == 0.0;
                    add_term(Ya, Yb, ctn[t], HA + (size_t)p.terms[t].mat * 2 * NE * 32);
                }
            for (int t = KPRE; t < p.nterms; ++t) {
                const cplx ct = step_coefficient<IO, false>(p.terms[t], c, p.pts, p.quad, p.magfac, j);
                xherm = xherm && ct.im == 0.0;
                add_term(Ya, Yb, ct, HA + (size_t)p.terms[t].mat * 2 * NE * 32);
            }
            if (!HORNER) {   // Horner form: sigma is folded into the monomial coefficients on the host, Y == X
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    (&Ya.re[0][0][0])[e] *= p.sigma; (&Ya.im[0][0][0])[e] *= p.sigma;
                    (&Yb.re[0][0])[e] *= p.sigma;    (&Yb.im[0][0])[e] *= p.sigma;
                    (&Yb.nim[0][0])[e] = neg((&Yb.im[0][0])[e]);
                }
            }

            K1_T(0)
            AccFrag<NT> S0, S1;
            if (HORNER == 4) {
                // ---- degree 12 in four products (api.cu solve_degree12; p.a[k].re = tV tW tY lV lW lY lI rV rW sV sW sY sI):
                //   W = X X, V = W X, T' = tV V + i tW W + tY X, y0 = T' V,
                //   E = (y0 + i lV V + lW W + i lY X + lI I)(y0 + i rV V + rW W) + i sV V + sW W + i sY X + sI I.
                // V and the right factor are needed as right operands: two layout shuffles.
                constexpr bool LO = sizeof(IO) == sizeof(double2);
                const int g = lane >> 2, q = lane & 3;
#pragma unroll
                for (int e = 0; e < NE; ++e) (&Yb.nim[0][0])[e] = neg((&Yb.im[0][0])[e]);
                AccFrag<NT> Wa, Va;
                set_zero<NT>(Wa);
                cmma<NT>(Wa, Ya, Yb);                       // W
                set_zero<NT>(Va);
                cmma<NT>(Va, Wa, Yb);                       // V = X^3
                BFrag<NT> Vb;
                acc_to_bfrag<NT>(Vb, Va, lane);
                {
                    const double tV = p.a[0].re, tW = p.a[1].re, tY = p.a[2].re;
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        (&Vb.nim[0][0])[e] = neg((&Vb.im[0][0])[e]);
                        (&S0.re[0][0][0])[e] = fma(tY, (&Ya.re[0][0][0])[e], fma(-tW, (&Wa.im[0][0][0])[e], tV * (&Va.re[0][0][0])[e]));
                        (&S0.im[0][0][0])[e] = fma(tY, (&Ya.im[0][0][0])[e], fma(tW, (&Wa.re[0][0][0])[e], tV * (&Va.im[0][0][0])[e]));
                    }
                }
                AccFrag<NT> Y0;
                set_zero<NT>(Y0);
                cmma<NT>(Y0, S0, Vb);                       // y0
                {   // right factor in accumulator layout (into S0), then shuffled
                    const double rV = p.a[7].re, rW = p.a[8].re;
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        (&S0.re[0][0][0])[e] = fma(rW, (&Wa.re[0][0][0])[e], fma(-rV, (&Va.im[0][0][0])[e], (&Y0.re[0][0][0])[e]));
                        (&S0.im[0][0][0])[e] = fma(rW, (&Wa.im[0][0][0])[e], fma(rV, (&Va.re[0][0][0])[e], (&Y0.im[0][0][0])[e]));
                    }
                }
                BFrag<NT> Rb;
                acc_to_bfrag<NT>(Rb, S0, lane);
#pragma unroll
                for (int e = 0; e < NE; ++e) (&Rb.nim[0][0])[e] = neg((&Rb.im[0][0])[e]);
                {
                    const double lV = p.a[3].re, lW = p.a[4].re, lY = p.a[5].re, lI = p.a[6].re;
                    const double sV = p.a[9].re, sW = p.a[10].re, sY = p.a[11].re, sI = p.a[12].re;
                    const double sV_lo = p.a_lo[9].re, sW_lo = p.a_lo[10].re, sY_lo = p.a_lo[11].re, sI_lo = p.a_lo[12].re;
#pragma unroll
                    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const bool diag = (mt == nt && g == 2 * q + i);
                                const double xr = Ya.re[mt][nt][i], xi = Ya.im[mt][nt][i];
                                const double wr = Wa.re[mt][nt][i], wi = Wa.im[mt][nt][i];
                                const double vr = Va.re[mt][nt][i], vi = Va.im[mt][nt][i];
                                double er, ei;
                                if (LO) {   // sub-ulp remainders first, the dominant X term last (DESIGN.md "Numerics")
                                    er = (sW_lo * wr - sV_lo * vi) - sY_lo * xi;
                                    ei = (sW_lo * wi + sV_lo * vr) + sY_lo * xr;
                                    if (diag) er = (er + sI_lo) + sI;
                                } else {
                                    er = diag ? sI : 0.0;
                                    ei = 0.0;
                                }
                                er = fma(-sV, vi, er); ei = fma(sV, vr, ei);
                                er = fma(sW, wr, er);  ei = fma(sW, wi, ei);
                                S1.re[mt][nt][i] = fma(-sY, xi, er);
                                S1.im[mt][nt][i] = fma(sY, xr, ei);
                                Y0.re[mt][nt][i] = fma(-lY, xi, fma(lW, wr, fma(-lV, vi, Y0.re[mt][nt][i]))) + (diag ? lI : 0.0);
                                Y0.im[mt][nt][i] = fma(lY, xr, fma(lW, wi, fma(lV, vr, Y0.im[mt][nt][i])));
                            }
                }
                cmma<NT>(S1, Y0, Rb);                       // E
            } else if (HORNER == 3) {
                // ---- degree 8 in three products (api.cu solve_degree8).  With A = -i X:  W = X^2 = -A^2,
                //   y02 = W (c4 W + i c3 X),   E = (y02 - d2 W - i d1 X + e0 I)(y02 - e2 W) - r2' W - i r1 X + r0 I
                // (the e0 y02 term of the published form is folded into the left factor: e0 y02 = e0 (y02 - e2 W) + e0 e2 W,
                // r2' = r2 - e0 e2 from the host).  W and y02 are needed on both sides of a product: their right-operand
                // layout comes from warp shuffles.  The addend only needs X and W, so it is formed under the y02 product.
                const double c4 = p.a[0].re, c3 = p.a[1].re, d2 = p.a[2].re, d1 = p.a[3].re, e2 = p.a[4].re, e0 = p.a[5].re;
                const double r2 = p.a[6].re, r1 = p.a[7].re, r0 = p.a[8].re;
                const int g = lane >> 2, q = lane & 3;
                if constexpr (MIXED && NT == 2) {
                    // E = -r2 W - i r1 X + r0 I  +  e0 y02 + L' R      with  L = L' + e0 I,  r2 = r2' + e0 e2  (api.cu solve_degree8):
                    // the first three terms in FP64, the rest -- fourth order and up, plus the part of the cubic term that the
                    // product carries -- at fp32 grade.  The constants that enter the cubic coefficient (c3, d1, e2, e0) are
                    // hi + lo pairs: a float-rounded constant would be the same relative error in every step.
                    // (split on the host, api.cu solve_degree8: kernel parameters are read straight from the constant bank)
                    const float c4f = p.fconst[0], d2f = p.fconst[1], c3h = p.fconst[2], c3l = p.fconst[3], d1h = p.fconst[4], d1l = p.fconst[5];
                    const float e2h = p.fconst[6], e2l = p.fconst[7], e0h = p.fconst[8], e0l = p.fconst[9];
                    const double r2full = p.a[9].re;
                    if (xherm) conj_transpose_as_bfrag<NT>(Yb, Ya); else acc_to_bfrag<NT>(Yb, Ya, lane);
#pragma unroll
                    for (int e = 0; e < NE; ++e) (&Yb.nim[0][0])[e] = (&Yb.re[0][0])[e] + (&Yb.im[0][0])[e];
                    AccFrag<NT> Wa;
                    set_zero<NT>(Wa);
                    if (xherm) cmma3_herm16(Wa, Ya, Yb, lane);   // Hermitian X: three of the four tiles, the fourth mirrored
                    else cmma3<NT>(Wa, Ya, Yb);             // W = X X in FP64
                    K1_T(1)
                    FAcc2 Xf, Wf;
#pragma unroll
                    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const bool diag = (mt == nt && g == 2 * q + i);
                                S1.re[mt][nt][i] = fma(-r2full, Wa.re[mt][nt][i], fma(r1, Ya.im[mt][nt][i], diag ? r0 : 0.0));
                                S1.im[mt][nt][i] = fma(-r2full, Wa.im[mt][nt][i], -r1 * Ya.re[mt][nt][i]);
                                Xf.re[mt][nt][i] = (float)Ya.re[mt][nt][i]; Xf.im[mt][nt][i] = (float)Ya.im[mt][nt][i];
                                Wf.re[mt][nt][i] = (float)Wa.re[mt][nt][i]; Wf.im[mt][nt][i] = (float)Wa.im[mt][nt][i];
                            }
                    FAcc2 Tf;                              // T = c4 W + i c3 X
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        (&Tf.re[0][0][0])[e] = fmaf(-c3h, (&Xf.im[0][0][0])[e], fmaf(-c3l, (&Xf.im[0][0][0])[e], c4f * (&Wf.re[0][0][0])[e]));
                        (&Tf.im[0][0][0])[e] = fmaf(c3h, (&Xf.re[0][0][0])[e], fmaf(c3l, (&Xf.re[0][0][0])[e], c4f * (&Wf.im[0][0][0])[e]));
                    }
                    K1_T(5)
                    FB2 Wbf;                               // W = X X is Hermitian with X (to rounding): same relabeling
                    if (xherm) fconj_transpose_as_fb(Wbf, Wf); else facc_to_fb(Wbf, Wf, lane);
                    K1_T(6)
                    FAcc2 Y2f;
                    tf32_cmul16(Y2f, Tf, Wbf);              // y02 = T W
                    K1_T(7)
                    FB2 Rbf;                                // R = y02 - e2 W, right-operand layout
                    facc_to_fb(Rbf, Y2f, lane);
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        (&Rbf.re[0][0])[e] = fmaf(-e2h, (&Wbf.re[0][0])[e], fmaf(-e2l, (&Wbf.re[0][0])[e], (&Rbf.re[0][0])[e]));
                        (&Rbf.im[0][0])[e] = fmaf(-e2h, (&Wbf.im[0][0])[e], fmaf(-e2l, (&Wbf.im[0][0])[e], (&Rbf.im[0][0])[e]));
                    }
                    FAcc2 Lf;                              // L' = y02 - d2 W - i d1 X
#pragma unroll
                    for (int e = 0; e < NE; ++e) {
                        (&Lf.re[0][0][0])[e] = fmaf(d1h, (&Xf.im[0][0][0])[e], fmaf(d1l, (&Xf.im[0][0][0])[e], fmaf(-d2f, (&Wf.re[0][0][0])[e], (&Y2f.re[0][0][0])[e])));
                        (&Lf.im[0][0][0])[e] = fmaf(-d1h, (&Xf.re[0][0][0])[e], fmaf(-d1l, (&Xf.re[0][0][0])[e], fmaf(-d2f, (&Wf.im[0][0][0])[e], (&Y2f.im[0][0][0])[e])));
                    }
                    K1_T(8)
                    FAcc2 LRf;
                    tf32_cmul16(LRf, Lf, Rbf);              // L' R
#pragma unroll
                    for (int e = 0; e < NE; ++e) {          // E = S + (e0 y02 + L' R)
                        (&S1.re[0][0][0])[e] += (double)fmaf(e0h, (&Y2f.re[0][0][0])[e], fmaf(e0l, (&Y2f.re[0][0][0])[e], (&LRf.re[0][0][0])[e]));
                        (&S1.im[0][0][0])[e] += (double)fmaf(e0h, (&Y2f.im[0][0][0])[e], fmaf(e0l, (&Y2f.im[0][0][0])[e], (&LRf.im[0][0][0])[e]));
                    }
                } else {
                if (!BOTH) { if (xherm) conj_transpose_as_bfrag<NT>(Yb, Ya); else acc_to_bfrag<NT>(Yb, Ya, lane); }
#pragma unroll
                for (int e = 0; e < NE; ++e) (&Yb.nim[0][0])[e] = bfrag_third<MUL3>((&Yb.re[0][0])[e], (&Yb.im[0][0])[e]);
                AccFrag<NT> Wa;
                set_zero<NT>(Wa);
                bool w_done = false;
                if constexpr (MUL3 && NT == 2) {
                    if (xherm) { cmma3_herm16(Wa, Ya, Yb, lane); w_done = true; }   // Hermitian X: three tiles, the fourth mirrored
                }
                if (!w_done) { if (MUL3) cmma3<NT>(Wa, Ya, Yb); else cmma<NT>(Wa, Ya, Yb); }   // W
                BFrag<NT> Wb;
                if (!BOTH && xherm) conj_transpose_as_bfrag<NT>(Wb, Wa); else acc_to_bfrag<NT>(Wb, Wa, lane);   // W = X X is Hermitian with X (to rounding)
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    (&Wb.nim[0][0])[e] = bfrag_third<MUL3>((&Wb.re[0][0])[e], (&Wb.im[0][0])[e]);
                    (&S0.re[0][0][0])[e] = fma(-c3, (&Ya.im[0][0][0])[e], c4 * (&Wa.re[0][0][0])[e]);
                    (&S0.im[0][0][0])[e] = fma(c3, (&Ya.re[0][0][0])[e], c4 * (&Wa.im[0][0][0])[e]);
                }
                K1_T(1)
                AccFrag<NT> Y2;
                set_zero<NT>(Y2);
                if (MUL3) cmma3<NT>(Y2, S0, Wb); else cmma<NT>(Y2, S0, Wb);   // y02
#pragma unroll
                for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const bool diag = (mt == nt && g == 2 * q + i);
                            S1.re[mt][nt][i] = fma(-r2, Wa.re[mt][nt][i], fma(r1, Ya.im[mt][nt][i], diag ? r0 : 0.0));
                            S1.im[mt][nt][i] = fma(-r2, Wa.im[mt][nt][i], -r1 * Ya.re[mt][nt][i]);
                        }
                BFrag<NT> Rb;
                acc_to_bfrag<NT>(Rb, Y2, lane);
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    (&Rb.re[0][0])[e] = fma(-e2, (&Wb.re[0][0])[e], (&Rb.re[0][0])[e]);
                    (&Rb.im[0][0])[e] = fma(-e2, (&Wb.im[0][0])[e], (&Rb.im[0][0])[e]);
                    (&Rb.nim[0][0])[e] = bfrag_third<MUL3>((&Rb.re[0][0])[e], (&Rb.im[0][0])[e]);
                }
#pragma unroll
                for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const bool diag = (mt == nt && g == 2 * q + i);
                            Y2.re[mt][nt][i] = fma(-d2, Wa.re[mt][nt][i], fma(d1, Ya.im[mt][nt][i], Y2.re[mt][nt][i])) + (diag ? e0 : 0.0);
                            Y2.im[mt][nt][i] = fma(-d2, Wa.im[mt][nt][i], fma(-d1, Ya.re[mt][nt][i], Y2.im[mt][nt][i]));
                        }
                if (MUL3) cmma3<NT>(S1, Y2, Rb); else cmma<NT>(S1, Y2, Rb);   // E
                }
            } else if (HORNER) {
                // ---- Horner in W = Y^2:  E = sum_i (c_2i I + c_2i+1 Y) W^i, 1 + floor(M/2) products instead of M - 1.
                // W is needed as a RIGHT operand.  (Y^T)^2 = (Y^2)^T is formed in accumulator layout from register
                // relabelings only -- AccFrag(Y^T) is Yb, BFrag(Y^T) is Ya (frag.cuh) -- and its transpose read as BFrag.
                AccFrag<NT> Zt;
                BFrag<NT> Bt;
#pragma unroll
                for (int mt = 0; mt < NT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            Zt.re[mt][nt][i] = Yb.re[2 * nt + i][mt];
                            Zt.im[mt][nt][i] = Yb.im[2 * nt + i][mt];
                        }
                transpose_as_bfrag<NT>(Bt, Ya);
                AccFrag<NT> Wt;
                set_zero<NT>(Wt);
                cmma<NT>(Wt, Zt, Bt);                       // (Y^T)(Y^T)
                BFrag<NT> Wb;
                transpose_as_bfrag<NT>(Wb, Wt);             // BFrag(Y^2)
                K1_T(1)
                const int L = M >> 1;
                horner_addend<NT, false>(S0, p.a[2 * L + 1].im, 0.0, Ya, p.a[2 * L].re, 0.0, lane);   // B_L (c_{M+1} = 0 for even M)
#pragma unroll 1
                for (int i = L - 1; i >= 0; --i) {
                    if (i <= 1 && sizeof(IO) == sizeof(double2))   // sub-ulp remainders only matter for complex128 contexts
                        horner_addend<NT, true>(S1, p.a[2 * i + 1].im, p.a_lo[2 * i + 1].im, Ya, p.a[2 * i].re, p.a_lo[2 * i].re, lane);
                    else        horner_addend<NT, false>(S1, p.a[2 * i + 1].im, 0.0, Ya, p.a[2 * i].re, 0.0, lane);
                    cmma<NT>(S1, S0, Wb);                   // R <- R W + B_i
                    const AccFrag<NT> T = S0; S0 = S1; S1 = T;
                }
                S1 = S0;                                    // E
            } else
            // ---- Clenshaw in E-form: E = U - I = a0' I + B_1 Y - 2 B_2 ----
            if (M == 1) {
                axpb_diag<NT>(S1, p.a[1], Ya, p.a[0], p.a_lo[0], lane);
            } else {
                axpb_diag<NT>(S0, p.a[M], Ya, p.a[M - 1], p.a_lo[M - 1], lane);      // B_{M-1}
                set_zero<NT>(S1);
                scale_add_diag<NT>(S1, 0.0, p.a[M], p.a_lo[M], lane);            // B_M
                for (int k = M - 2; k >= 1; --k) {
                    scale_add_diag<NT>(S1, -1.0, p.a[k], p.a_lo[k], lane);       // a_k I - B_{k+2}
                    cmma<NT>(S1, S0, Yb);                             // + B_{k+1} Y
                    const AccFrag<NT> T = S0; S0 = S1; S1 = T;
                }
                scale_add_diag<NT>(S1, -2.0, p.a[0], p.a_lo[0], lane);           // a0' I - 2 B_2
                cmma<NT>(S1, S0, Yb);                                 // + B_1 Y   -> E
            }

            K1_T(2)
            // ---- running product  Q <- Q (I + E)^T = Q + Q E^T, with the next step's raw samples in flight ----
            const unsigned long long jn = j + 1 < hi ? j + 1 : j;
            RawAmp<IO> raw[KPRE > 0 ? KPRE : 1];
#pragma unroll
            for (int t = 0; t < KPRE; ++t)
                if (t < p.nterms) raw[t] = load_raw<IO, false>(p.terms[t], c, p.pts, p.quad, jn);
            if (NT == 1 && pack > 1 && !live) set_zero<NT>(S1);   // this block has run out of steps: E = 0 exactly, Q_b stays
            BFrag<NT> Et;
            transpose_as_bfrag<NT>(Et, S1);
            AccFrag<NT> Qn = Q;
            if (MUL3) {
#pragma unroll
                for (int e = 0; e < NE; ++e) (&Et.nim[0][0])[e] = (&Et.re[0][0])[e] + (&Et.im[0][0])[e];
                cmma3<NT>(Qn, Q, Et);
            } else {
                cmma<NT>(Qn, Q, Et);
            }
            Q = Qn;
#pragma unroll
            for (int t = 0; t < KPRE; ++t)
                if (t < p.nterms) ctn[t] = coef_from_raw<IO>(p.terms[t], p.quad, p.magfac, raw[t]);
            K1_T(3)
        }
        K1_T_PRINT
        if constexpr (NT == 1) { if (pack > 1) unpack_blocks<NT>(Q, pack, lane); }
    }

    k1_tail<NT, IO>(Q, active, pulse, chunk, chunks_per_pulse, reduce_in_cta, partials, fz, smem, warp, lane);
}

// Ordered product of `count` propagators of time slices (multi-GPU combine, dim <= 16): out = parts[count-1] ... parts[0].
// ONE launch of one CTA: each warp multiplies a contiguous sub-range, then the in-CTA tree (kernel 2's) finishes.
// Replaces the padded E-form tree on the GEMM kernel (five launches for eight 16 x 16 partials in round 1).
template <int NT, typename IO>
__global__ void __launch_bounds__(256)
k3_combine_kernel(const IO *__restrict__ parts, unsigned int count, int n, IO *__restrict__ out) {
    constexpr int NP = 8 * NT;
    __shared__ double2 smem[4 * NP * NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned int b0 = (unsigned int)((unsigned long long)count * warp / nwarps);
    const unsigned int b1 = (unsigned int)((unsigned long long)count * (warp + 1) / nwarps);
    AccFrag<NT> Q;
    if (b0 < b1) {
        load_acc_of_transpose<NT, IO>(Q, parts + (size_t)b0 * n * n, n, lane);
        for (unsigned int b = b0 + 1; b < b1; ++b) {
            BFrag<NT> B;
            load_bfrag_of_transpose<NT, IO>(B, parts + (size_t)b * n * n, n, lane);
            AccFrag<NT> R;
            set_zero<NT>(R);
            cmma<NT>(R, Q, B);
            Q = R;
        }
    } else {
        set_identity<NT>(Q, lane);
    }
    cta_ordered_product<NT>(Q, smem, warp, nwarps, lane);
    if (warp == 0) store_propagator<NT, IO>(Q, out, n, lane);
}

// Ordered reduction of the partial products of each pulse.  grid = (pulses, groups): CTA (pulse, g) multiplies the
// contiguous group g of that pulse's nb partials (each warp a contiguous sub-range, the next partial prefetched while the
// current product runs, then the in-CTA tree).  With mid != nullptr the group product is written back as a partial
// (first level of a two-level reduction), otherwise it is transposed to the row-major physical propagator, converted to
// the IO precision and written to out[pulse].
template <int NT, typename IO>
__global__ void __launch_bounds__(256)
k3_reduce_kernel(const double2 *__restrict__ partials, unsigned int nb, unsigned int group, int n, double2 *__restrict__ mid,
                 IO *__restrict__ out) {
    constexpr int NP = 8 * NT;
    __shared__ double2 smem[4 * NP * NP];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const unsigned int pulse = blockIdx.x, g = blockIdx.y;
    const unsigned int g0 = g * group, g1 = min(nb, g0 + group);
    const double2 *P = partials + ((size_t)pulse * nb + g0) * NP * NP;
    const unsigned int cnt = g1 - g0;

    AccFrag<NT> Q;
    cta_reduce_range<NT>(Q, P, cnt, smem, warp, nwarps, lane);
    if (warp == 0) {
        if (mid) {
            store_acc<NT>(Q, mid + ((size_t)pulse * gridDim.y + g) * NP * NP, NP, lane);
        } else {
            store_propagator<NT, IO>(Q, out + (size_t)pulse * n * n, n, lane);   // transpose back: P = Q^T
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
template <int NT, typename IO, int HORNER, int OCC, bool MUL3 = false, bool MIXED = false, bool HERMK = false>
static cudaError_t launch_chain_ttt(const SeriesParams &p, const IO *carr, const double2 *Hfrag, double2 *partials,
                                   unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                                   unsigned long long step_hi, const K1Final &fz, cudaStream_t stream) {
    k1_chain_kernel<NT, IO, HORNER, OCC, MUL3, MIXED, HERMK><<<plan.grid, 32 * K1_WARPS, 0, stream>>>(p, carr, Hfrag, partials, batch,
                                                                                   plan.chunks_per_pulse, step_lo, step_hi,
                                                                                   plan.reduce_in_cta, fz);
    return cudaGetLastError();
}

// $PARAMENT_K1_3M=0 selects the four-product kernel of dim 9..16 (A/B runs); read once
static bool k1_mul3() {
    static const bool mul3 = !(getenv("PARAMENT_K1_3M") && atoi(getenv("PARAMENT_K1_3M")) == 0);
    return mul3;
}
int k1_real_products(int npad, bool fp64_io, int horner) { return (!fp64_io && horner == 3 && k1_mul3()) ? 3 : 4; }

template <int NT, typename IO, int HORNER>
static cudaError_t launch_chain_tt(const SeriesParams &p, const IO *carr, const double2 *Hfrag, double2 *partials,
                                   unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                                   unsigned long long step_hi, const K1Final &fz, cudaStream_t stream) {
    if constexpr (NT == 1 && HORNER == 3) {
        if (k1_mul3()) return launch_chain_ttt<NT, IO, HORNER, 6, true>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    }
    if constexpr (NT == 1) {   // `if constexpr` throughout: a plain `if` would instantiate six-CTA variants of NT == 2 (5 KB of spills, never launched)
        return launch_chain_ttt<NT, IO, HORNER, 6>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    } else {
    if constexpr (HORNER == 3 && sizeof(IO) == sizeof(float2)) {
        // mixed precision: the two small products of the series on the TF32 tensor path (api.cu use_mixed_path decides).
        // Measured at C2: 2.149 -> 1.818 ms per 5e5 steps (three CTAs per SM at 168 registers: 1.870 ms), error 2.3e-7.
        if (p.mixed) return p.herm ? launch_chain_ttt<NT, IO, HORNER, 2, true, true, true>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
                                   : launch_chain_ttt<NT, IO, HORNER, 2, true, true, false>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    }
    // three CTAs per SM (168 registers): the default of the plain Taylor form; for the degree-8 form only through $PARAMENT_K1_OCC=3.
    // The Horner forms spill 0.6-2.2 KB at 168 registers and always run two CTAs per SM.
    if constexpr (HORNER == 0 || HORNER == 3)
        if (plan.ctas_per_sm == 3) return launch_chain_ttt<NT, IO, HORNER, 3>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    if constexpr (HORNER == 3) {
        // complex products from three real ones: measured at C2 2.285 -> 2.149 ms per 5e5 steps, same error (2.65e-8)
        if (k1_mul3()) return p.herm ? launch_chain_ttt<NT, IO, HORNER, 2, true, false, true>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
                                     : launch_chain_ttt<NT, IO, HORNER, 2, true, false, false>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    }
    return launch_chain_ttt<NT, IO, HORNER, 2>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    }
}

template <int NT, typename IO>
static cudaError_t launch_chain_t(const SeriesParams &p, const IO *carr, const double2 *Hfrag, double2 *partials,
                                  unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                                  unsigned long long step_hi, const K1Final &fz, cudaStream_t stream) {
    if (p.horner == 4) return launch_chain_tt<NT, IO, 4>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    if (p.horner == 3) {   // complex64 contexts only (api.cu build_series)
        if constexpr (sizeof(IO) == sizeof(float2))
            return launch_chain_tt<NT, IO, 3>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
        else
            return cudaErrorInvalidValue;
    }
    return p.horner ? launch_chain_tt<NT, IO, 1>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
                    : launch_chain_tt<NT, IO, 0>(p, carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
}

cudaError_t launch_k1_chain(int npad, bool fp64_io, const SeriesParams &p, const void *carr, const double2 *Hfrag,
                            double2 *partials, unsigned int batch, const K1Plan &plan, unsigned long long step_lo,
                            unsigned long long step_hi, const K1Final &fz, cudaStream_t stream) {
    if (npad == 8) {
        return fp64_io ? launch_chain_t<1, double2>(p, (const double2 *)carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
                       : launch_chain_t<1, float2>(p, (const float2 *)carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
    }
    return fp64_io ? launch_chain_t<2, double2>(p, (const double2 *)carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream)
                   : launch_chain_t<2, float2>(p, (const float2 *)carr, Hfrag, partials, batch, plan, step_lo, step_hi, fz, stream);
}

template <int NT, typename IO>
static cudaError_t launch_k3_t(const double2 *partials, unsigned int nb, int n, double2 *mid, IO *out, unsigned int batch,
                               cudaStream_t stream) {
    constexpr unsigned int GROUP = 16;
    if (nb > 32 && mid) {   // two levels: groups of 16 in parallel CTAs, then the group products
        const unsigned int ng = (nb + GROUP - 1) / GROUP;
        k3_reduce_kernel<NT, IO><<<dim3(batch, ng), 256, 0, stream>>>(partials, nb, GROUP, n, mid, out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        k3_reduce_kernel<NT, IO><<<dim3(batch, 1), 32 * k3_warps_for(ng), 0, stream>>>(mid, ng, ng, n, nullptr, out);
        return cudaGetLastError();
    }
    k3_reduce_kernel<NT, IO><<<dim3(batch, 1), 32 * k3_warps_for(nb), 0, stream>>>(partials, nb, nb, n, nullptr, out);
    return cudaGetLastError();
}

// `mid`: scratch for k3_mid_elems(...) double2 (first-level group products); may be null for nb <= 32.
cudaError_t launch_k3_reduce(int npad, bool fp64_io, const double2 *partials, unsigned int partials_per_pulse, int n,
                             void *out, unsigned int batch, double2 *mid, cudaStream_t stream) {
    if (npad == 8)
        return fp64_io ? launch_k3_t<1, double2>(partials, partials_per_pulse, n, mid, (double2 *)out, batch, stream)
                       : launch_k3_t<1, float2>(partials, partials_per_pulse, n, mid, (float2 *)out, batch, stream);
    return fp64_io ? launch_k3_t<2, double2>(partials, partials_per_pulse, n, mid, (double2 *)out, batch, stream)
                   : launch_k3_t<2, float2>(partials, partials_per_pulse, n, mid, (float2 *)out, batch, stream);
}

template <int NT, typename IO>
static cudaError_t launch_combine_t(const IO *parts, unsigned int count, int n, IO *out, cudaStream_t stream) {
    k3_combine_kernel<NT, IO><<<1, 32 * k3_warps_for(count), 0, stream>>>(parts, count, n, out);
    return cudaGetLastError();
}

cudaError_t launch_k3_combine(int npad, bool fp64_io, const void *parts, unsigned int count, int n, void *out, cudaStream_t stream) {
    if (npad == 8)
        return fp64_io ? launch_combine_t<1, double2>((const double2 *)parts, count, n, (double2 *)out, stream)
                       : launch_combine_t<1, float2>((const float2 *)parts, count, n, (float2 *)out, stream);
    return fp64_io ? launch_combine_t<2, double2>((const double2 *)parts, count, n, (double2 *)out, stream)
                   : launch_combine_t<2, float2>((const float2 *)parts, count, n, (float2 *)out, stream);
}

}  // namespace pb
