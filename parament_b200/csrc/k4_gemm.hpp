// k4_gemm.hpp -- host interface of the complex FP64 tensor-pipe contraction kernels (dim > 16).
#pragma once
#include <cuda_runtime.h>
#include "params.hpp"

namespace pb {

// D_b = alpha * (A_b * B_b) + sum_j (beta[j] + beta_lo[j]) * C[j]_b + beta2 * C2_b + (gamma + gamma_lo) * I   for b < batch,
// and optionally Dprod_b = A_b * B_b and Dalt_b = A_b * B_b + sum_j beta_alt[j] * C[j]_b (a second combination of the same addends).  All matrices n x n row-major interleaved complex double, n a multiple of 32
// (n <= 32) or 64.  Addends may be null and may alias D.  The *_lo parts are sub-ulp remainders of the series constants;
// they are added before the leading parts (DESIGN.md "Numerics").
constexpr int kMaxAddends = 3;
struct GemmArgs {
    const double2 *A; long long strideA;
    const double2 *B; long long strideB;
    const double2 *C[kMaxAddends]; long long strideC[kMaxAddends]; cplx beta[kMaxAddends]; cplx beta_lo[kMaxAddends];
    const double2 *C2; long long strideC2; double beta2;
    double2 *D; long long strideD;
    double2 *Dprod; long long strideDprod;
    double2 *Dalt; long long strideDalt; cplx beta_alt[kMaxAddends];
    cplx alpha; int scaled;          // scaled != 0: the product is multiplied by alpha in D
    cplx gamma; cplx gamma_lo;
    int n;
    int batch;
    int herm;                        // A == B is Hermitian and so is every addend: the product is Hermitian.  Only the tiles that
                                     // touch the upper triangle are computed; their epilogue also writes the mirrored elements
                                     // (every output evaluated on the conjugated product and addends) into the skipped tiles
};

// The per-step series as a short program of fused GEMMs over matrix slots
//   0 = Y    1 = Y^2 (W)    2 = Y^3    3 = Y^4 (V)    4, 5 = recurrence registers
// (degree 12 in four products: 0 = Y, 1 = W, 2 = V = Y^3, 3 = T' then E, 4 = L, 5 = R; api.cu solve_degree12)
// `assemble` writes slot 0, slot 4 = u Y + (v + v_lo) I and, if init5, slot 5 = w I.
constexpr int kSeriesSlots = 6;
struct SeriesOp {
    int A, B, D, Dprod;          // slots; Dprod < 0: none
    int Dalt;                    // slot of the second combination (< 0: none), coefficients beta_alt
    cplx beta_alt[kMaxAddends];
    int scaled; cplx alpha;
    int C[kMaxAddends];          // addend slots, < 0: unused
    cplx beta[kMaxAddends], beta_lo[kMaxAddends];
    cplx gamma, gamma_lo;
};
constexpr int kMaxOps = kMaxDegree + 2;
struct SeriesProgram {
    int nops;
    int e_slot;                  // slot holding E = U - I after the last op
    int init5;
    cplx u, v, v_lo, w;
    SeriesOp ops[kMaxOps];
};

SeriesProgram build_program(const SeriesParams &p);

cudaError_t k4_gemm(const GemmArgs &g, cudaStream_t stream);
int k4_pad(int n);                            // padded dimension used by this kernel family
int k4_tiles(int npad);                       // CTA tiles per matrix
int k4_herm_tiles(int npad);                  // CTA tiles per matrix of a Hermitian-output launch (upper-triangular tiles only)
int k4_real_products(int npad);               // real matrix products per complex product of the batched GEMM (4, or 3)
int k4_wave_slots(int npad, int num_sms);     // co-resident CTAs of the batched GEMM kernel on the device
int k4_chain_slots(int npad, int num_sms);    // co-resident CTAs of the persistent chain kernel (npad <= 64)

// Batched path: slots[s] points at S consecutive matrices.
cudaError_t k4_assemble(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                        double2 *slot0, double2 *slot4, double2 *slot5, unsigned long long step0, int S, cudaStream_t stream);
cudaError_t k4_eform(bool fp64_io, const void *P, int n, int npad, int count, double2 *E, cudaStream_t stream);   // E = P - I, padded
cudaError_t k4_finish(bool fp64_io, const double2 *E, int n, int npad, void *out, bool add_identity, cudaStream_t stream);

// out_dev[k] <- bits of max |c_k|^2 (double) over `batch` pulses of `pts` points; control arrays are `stride` points apart.
cudaError_t k4_absmax(bool fp64_io, const void *carr, unsigned int batch, unsigned int amps, size_t stride, size_t pts,
                      unsigned long long *out_dev, cudaStream_t stream);

// Persistent single-tile chain kernel (npad == 32 or 64): `grid` CTAs each walk a contiguous range of the nsteps steps
// with their matrices in a private L2-resident scratch (kSeriesSlots + 2 matrices per CTA) and leave one E-form partial each.
cudaError_t k4_chain(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                     double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream);

// dim 17..64: persistent kernel with the step's operands resident in shared memory (k4_onchip.cu); scratch: 2 matrices per CTA.
int k4_onchip_slots(int npad, int num_sms);   // co-resident CTAs; 0 if the device cannot host it
cudaError_t k4_onchip(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                      double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream);

}  // namespace pb
