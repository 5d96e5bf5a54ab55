// series.hpp -- host-side scalar mathematics of the Chebyshev/Bessel series.
//
//   * iteration-count tables (API-visible data of Parament_selectIterationCycles_fp32/_fp64,
//     /root/reference/src/cuda/parament.cpp:723-766)
//   * Bessel coefficients a_k = (-i)^k J_k(x)   (reference: libm jn() in double, mathhelper.cpp:61-80).
//     Here J_k comes from Miller's backward recurrence in long double, normalised with
//     1 = J_0 + 2 sum_{m>=1} J_{2m}; the same identity yields a_0' = J_0 - 1 = -2 sum J_{2m} without
//     cancellation (E-form of the series, DESIGN.md "Numerics").
#pragma once
#include <cmath>
#include <vector>

namespace pb {

inline int select_cycles_fp32(double H_norm, double dt) {
    static const double thr[] = {0.032516793, 0.219062571, 0.619625593, 1.218059203, 1.979888284, 2.873301187,
                                 3.872963682, 4.959398466, 6.117657121, 7.336154907, 8.605792444, 9.919320831,
                                 11.27088616, 12.65570085};
    const double x = H_norm * dt;
    for (int i = 0; i < 14; ++i)
        if (x <= thr[i]) return 3 + 2 * i;
    return -1;
}

inline int select_cycles_fp64(double H_norm, double dt) {
    static const double thr[] = {0.000213616, 0.00768149, 0.0501474, 0.162592, 0.368382, 0.676861, 1.08784,
                                 1.59605, 2.19402, 2.87366, 3.62716, 4.44725, 5.3274, 6.26179,
                                 7.2453, 8.27338, 9.34206, 10.4478, 11.5875, 12.7584};
    const double x = H_norm * dt;
    for (int i = 0; i < 20; ++i)
        if (x <= thr[i]) return 3 + 2 * i;
    return -1;
}

// J_0(x) .. J_kmax(x) in long double, plus J_0(x) - 1 free of cancellation.
inline void bessel_j_table(long double x, int kmax, std::vector<long double> &J, long double &j0_minus_1) {
    J.assign(kmax + 1, 0.0L);
    if (x < 0) x = -x;   // only even/odd symmetry matters for callers with x >= 0
    if (x < 1e-30L) {
        J[0] = 1.0L;
        j0_minus_1 = -(x * x) / 4.0L;
        for (int k = 1; k <= kmax; ++k) J[k] = (k == 1) ? x / 2.0L : 0.0L;
        return;
    }
    int m0 = (int)std::ceil(1.5 * std::fmax((double)kmax, (double)x) + 40.0);
    if (m0 & 1) ++m0;
    std::vector<long double> j(m0 + 2, 0.0L);
    j[m0 + 1] = 0.0L;
    j[m0] = 1e-300L;
    for (int k = m0; k >= 1; --k) {
        j[k - 1] = (2.0L * k / x) * j[k] - j[k + 1];
        if (std::fabs(j[k - 1]) > 1e2000L) {   // rescale to stay inside the long double range
            for (int i = k - 1; i <= m0; ++i) j[i] *= 1e-2000L;
        }
    }
    long double even = 0.0L;
    for (int k = m0; k >= 2; k -= 2) even += j[k];   // small terms first
    const long double norm = j[0] + 2.0L * even;
    for (int k = 0; k <= kmax; ++k) J[k] = j[k] / norm;
    j0_minus_1 = -2.0L * even / norm;
}

}  // namespace pb
