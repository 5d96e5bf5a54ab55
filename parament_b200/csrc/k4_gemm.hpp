// k4_gemm.hpp -- host interface of the batched complex FP64 tensor-pipe contraction (dim > 16).
#pragma once
#include <cuda_runtime.h>
#include "params.hpp"

namespace pb {

// D_b = A_b * B_b + beta1 * C1_b + beta2 * C2_b + (gamma + gamma_lo) * I   for b < batch; all matrices n x n row-major
// interleaved complex double, n a multiple of 32 (n <= 32) or 64.  C1 / C2 may be null and may alias D.
struct GemmArgs {
    const double2 *A; long long strideA;
    const double2 *B; long long strideB;
    const double2 *C1; long long strideC1; double beta1;
    const double2 *C2; long long strideC2; double beta2;
    double2 *D; long long strideD;
    cplx gamma;
    cplx gamma_lo;   // sub-ulp remainder of gamma, added before gamma
    int n;
    int batch;
};

cudaError_t k4_gemm(const GemmArgs &g, cudaStream_t stream);
int k4_pad(int n);          // padded dimension used by this kernel family
int k4_tiles(int npad);     // CTA tiles per matrix
cudaError_t k4_assemble(bool fp64_io, const SeriesParams &p, const void *carr, const double2 *H, double2 *Y,
                        double2 *S0, double2 *S1, unsigned long long step0, int S, cudaStream_t stream);
cudaError_t k4_finish(bool fp64_io, const double2 *E, int n, int npad, void *out, bool add_identity, cudaStream_t stream);

}  // namespace pb
