// poly_solve.hpp -- parameters of the product-saving polynomial evaluation schemes (host only, no CUDA).
//
// J. Sastre, "Efficient evaluation of matrix polynomials", Linear Algebra Appl. 539 (2018): a matrix polynomial
// sum_m r_m A^m of degree 8 (12) can be evaluated with 3 (4) matrix products instead of the 4 (5) of the
// Paterson-Stockmeyer / Horner-in-A^2 forms, as nested products of linear combinations of low powers whose
// coefficients solve a small nonlinear system in the r_m.  The per-step series of this library, written in A = -i X,
// has REAL coefficients r_m (r_m = c_m / (-i)^m), so the parameters are real whenever the system has a real solution;
// the callers keep the Horner / Paterson-Stockmeyer form when it has none.  All arithmetic in long double.
#pragma once
#include <cmath>

namespace pb {

// Degree 8, three products:
//     A2  = A A
//     y02 = A2 (c4 A2 + c3 A)
//     E   = (y02 + d2 A2 + d1 A)(y02 + e2 A2) + e0 y02 + r2 A2 + r1 A + r0 I
// matches r_3 .. r_8 when
//     c4^2 = r8,  2 c3 c4 = r7,  c3^2 + (d2 + e2) c4 = r6,  (d2 + e2) c3 + d1 c4 = r5,
//     d1 c3 + d2 e2 + e0 c4 = r4,  d1 e2 + e0 c3 = r3.
// out = {c4, c3, d2, d1, e2, e0}.
inline bool solve_degree8_real(const long double r[9], long double out[6]) {
    if (!(r[8] > 0.0L) || r[7] == 0.0L) return false;
    const long double c4 = sqrtl(r[8]), c3 = r[7] / (2.0L * c4);
    const long double s = (r[6] - c3 * c3) / c4;            // d2 + e2
    const long double d1 = (r[5] - s * c3) / c4;
    // e2^2 - B e2 - C = 0
    const long double B = s - c4 * d1 / c3, C = c4 * r[3] / c3 - (r[4] - d1 * c3);
    const long double disc = B * B + 4.0L * C;
    if (!(disc >= 0.0L)) return false;
    const long double e2 = 0.5L * (B + sqrtl(disc));       // the root with the smaller |e0|
    const long double d2 = s - e2, e0 = (r[3] - d1 * e2) / c3;
    const long double v[6] = {c4, c3, d2, d1, e2, e0};
    for (int k = 0; k < 6; ++k) {
        if (!std::isfinite((double)v[k])) return false;
        out[k] = v[k];
    }
    return true;
}

// Degree 12, four products:
//     A2 = A A,  A3 = A2 A
//     y0 = A3 (c3 A3 + c2 A2 + c1 A)
//     E  = (y0 + d3 A3 + d2 A2 + d1 A)(y0 + e3 A3 + e2 A2) + f y0 + g3 A3 + r2 A2 + r1 A + r0 I
// Matching powers 12 .. 3 (s3 = d3 + e3, s2 = d2 + e2):
//     r12 = c3^2            r11 = 2 c2 c3                 r10 = c2^2 + 2 c1 c3
//     r9  = 2 c1 c2 + c3 s3                               r8  = c1^2 + c2 s3 + c3 s2
//     r7  = c1 s3 + c2 s2 + d1 c3
//     r6  = c1 s2 + d1 c2 + d3 e3 + f c3
//     r5  = d1 c1 + d3 e2 + d2 e3 + f c2
//     r4  = d2 e2 + d1 e3 + f c1
//     r3  = d1 e2 + g3
// c1..c3, s2, s3, d1 follow in closed form; r6 gives f as a quadratic in d3, r5 is then linear in d2, and r4 becomes a
// quartic in d3 whose real roots are bracketed on a grid and refined by bisection.  Among the real solutions the one
// with the smallest |f| is taken (ties: smallest |d2|).  out = {c1, c2, c3, d1, d2, d3, e2, e3, f}.
inline bool solve_degree12_real(const long double r[13], long double out[9]) {
    if (!(r[12] > 0.0L)) return false;
    const long double c3 = sqrtl(r[12]), c2 = r[11] / (2.0L * c3), c1 = (r[10] - c2 * c2) / (2.0L * c3);
    const long double s3 = (r[9] - 2.0L * c1 * c2) / c3;
    const long double s2 = (r[8] - c1 * c1 - c2 * s3) / c3;
    const long double d1 = (r[7] - c1 * s3 - c2 * s2) / c3;
    if (!(fabsl(s3) > 0.0L) || !std::isfinite((double)d1)) return false;
    const long double K6 = r[6] - c1 * s2 - d1 * c2;
    auto f_of = [&](long double d3) { return (K6 - d3 * (s3 - d3)) / c3; };
    // residual of the r4 equation times (s3 - 2 d3)^2, a quartic polynomial in d3
    auto resid = [&](long double d3, long double *d2_out) {
        const long double f = f_of(d3);
        const long double den = s3 - 2.0L * d3;
        const long double num = r[5] - d1 * c1 - s2 * d3 - f * c2;          // d2 * den = num
        if (d2_out) *d2_out = num / den;
        // r4 - [d2 (s2 - d2) + d1 (s3 - d3) + f c1], with d2 = num / den, multiplied by den^2
        return (r[4] - d1 * (s3 - d3) - f * c1) * den * den - num * (s2 * den - num);
    };
    long double best_f = 0.0L, best_d2 = 0.0L, best_d3 = 0.0L;
    bool found = false;
    const int G = 2000;
    const long double span = 5.0L * fabsl(s3);
    long double t_prev = -span, g_prev = resid(t_prev, nullptr);
    for (int i = 1; i <= G; ++i) {
        const long double t = -span + 2.0L * span * (long double)i / (long double)G;
        const long double g = resid(t, nullptr);
        if ((g_prev < 0.0L) != (g < 0.0L)) {
            long double a = t_prev, b = t, ga = g_prev;
            for (int it = 0; it < 90; ++it) {
                const long double m = 0.5L * (a + b), gm = resid(m, nullptr);
                if ((ga < 0.0L) != (gm < 0.0L)) b = m; else { a = m; ga = gm; }
            }
            const long double d3 = 0.5L * (a + b);
            if (fabsl(s3 - 2.0L * d3) > 1e-3L * fabsl(s3)) {     // away from the pole of d2(d3)
                long double d2;
                resid(d3, &d2);
                const long double f = f_of(d3);
                if (std::isfinite((double)d2) && std::isfinite((double)f) &&
                    (!found || fabsl(f) < fabsl(best_f) * (1.0L - 1e-9L) ||
                     (fabsl(f) <= fabsl(best_f) * (1.0L + 1e-9L) && fabsl(d2) < fabsl(best_d2)))) {
                    found = true; best_f = f; best_d2 = d2; best_d3 = d3;
                }
            }
        }
        t_prev = t; g_prev = g;
    }
    if (!found) return false;
    const long double d3 = best_d3, d2 = best_d2, f = best_f, e3 = s3 - d3, e2 = s2 - d2;
    // the solution must reproduce r4 .. r6 (r7 .. r12 hold by construction)
    const long double q6 = c1 * s2 + d1 * c2 + d3 * e3 + f * c3, q5 = d1 * c1 + d3 * e2 + d2 * e3 + f * c2,
                      q4 = d2 * e2 + d1 * e3 + f * c1;
    auto close = [](long double a, long double b) { return fabsl(a - b) <= 1e-13L * fabsl(b) + 1e-4000L; };
    if (!close(q6, r[6]) || !close(q5, r[5]) || !close(q4, r[4])) return false;
    const long double v[9] = {c1, c2, c3, d1, d2, d3, e2, e3, f};
    for (int k = 0; k < 9; ++k) {
        if (!std::isfinite((double)v[k])) return false;
        out[k] = v[k];
    }
    return true;
}

}  // namespace pb
