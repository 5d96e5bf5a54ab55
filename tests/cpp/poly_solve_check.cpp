// CPU driver for parament_b200/csrc/poly_solve.hpp (tests/test_poly_solve.py):
//   poly_solve_check 8|12 r_0 ... r_deg   ->   prints the scheme parameters, one per line, or "none"
#include <cstdio>
#include <cstdlib>
#include "poly_solve.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const int deg = atoi(argv[1]);
    if ((deg != 8 && deg != 12) || argc != deg + 3) return 2;
    long double r[13];
    for (int m = 0; m <= deg; ++m) r[m] = strtold(argv[2 + m], nullptr);
    long double out[9];
    const bool ok = deg == 8 ? pb::solve_degree8_real(r, out) : pb::solve_degree12_real(r, out);
    if (!ok) { printf("none\n"); return 0; }
    for (int k = 0; k < (deg == 8 ? 6 : 9); ++k) printf("%.24Le\n", out[k]);
    return 0;
}
