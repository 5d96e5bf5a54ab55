// k4_gemm.cu -- kernel (4) of the north_star: dense complex FP64 contraction for dim > 16.
//
// tile_gemm             one BM x BN complex output tile  D = A B + beta1 C1 + beta2 C2 + gamma I  on the FP64 tensor
//                       pipe (mma.sync.m8n8k4.f64 -> DMMA.8x8x4): 3-stage cp.async shared-memory pipeline,
//                       bank-conflict-free pitches for the LDS.128 fragment loads, fused epilogue (the Clenshaw
//                       "-B_{k+2} + a_k I", the Horner "+c_{2i+1} Y + c_{2i} I" and the E-form pair product
//                       "E_b + E_a + E_b E_a" of the ordered reduction).  Reference: cublasZgemmStridedBatched plus a
//                       separate diagonal_add launch per iteration (parament.cpp:596-643,681-690, diagonal_add.cu:21-51).
// k4_chain_kernel       dim 17..64: PERSISTENT kernel, one CTA per contiguous range of time steps.  Per step the CTA
//                       assembles Y into its private, L2-resident scratch, runs the whole series program and multiplies
//                       the step into its running product -- one launch for the whole pulse, no grid-wide dependency.
// k4_zgemm_kernel       dim > 64: the same tile code as a batched launch over the S steps of an L2-resident time chunk.
// k4_assemble_kernel    batched assembly Y_s = sigma (H0 + sum_t c_t(s) H_t) + series start values (fuses the reference's
//                       outer-product broadcast, quadrature kernels and rank-A' GEMM, parament.cpp:491-554).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "coef.cuh"
#include "k4_gemm.hpp"
#include "plan.hpp"

namespace pb {

constexpr int K4_BK = 16;       // contraction depth per pipeline stage
constexpr int K4_STAGES = 3;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Bulk asynchronous copies (the TMA engine's 1-D path, SASS UBLKCP): one instruction moves a whole tile row into the
// padded shared-memory layout and signals an mbarrier with the bytes it delivered.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();   // a lost copy must not hang the device
    }
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem, const void *gmem, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(bar) : "memory");
}

template <int BM, int BN, int STAGES = K4_STAGES>
struct K4Smem {
    static constexpr int PA = K4_BK + 4;   // pitch of the A tile in double2: 8 lanes of an LDS.128 phase hit 8 bank groups
    static constexpr int PB = BN + 2;      // pitch of the B tile
    static constexpr int A_ELEMS = BM * PA;
    static constexpr int B_ELEMS = K4_BK * PB;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double2);
};

struct TileArgs {
    const double2 *A, *B;             // matrix bases (n x n row-major)
    const double2 *C[kMaxAddends];
    const double2 *C2;
    double2 *D, *Dprod, *Dalt;
    cplx alpha; int scaled;
    cplx beta[kMaxAddends], beta_lo[kMaxAddends], beta_alt[kMaxAddends];
    double beta2;
    cplx gamma, gamma_lo;
    int n;
    int herm;                         // GemmArgs::herm
};

// Shared epilogue arithmetic for two horizontally adjacent elements (row r, columns c and c+1): on entry (vr, vi) hold the
// product, z[j][i] the addends; small terms are added first, the dominant ones last with a single-rounding FMA.
__device__ __forceinline__ void epilogue_pair(double (&vr)[2], double (&vi)[2], const double2 (&z)[kMaxAddends][2], const bool (&has)[kMaxAddends],
                                              int scaled, cplx alpha, const cplx (&beta)[kMaxAddends], const cplx (&beta_lo)[kMaxAddends],
                                              cplx gamma, cplx gamma_lo, bool diag0, bool diag1) {
    if (scaled) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double pr = vr[i], pi = vi[i];
            vr[i] = alpha.re * pr - alpha.im * pi;
            vi[i] = alpha.re * pi + alpha.im * pr;
        }
    }
#pragma unroll
    for (int j = 0; j < kMaxAddends; ++j)
        if (has[j]) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                vr[i] += beta_lo[j].re * z[j][i].x - beta_lo[j].im * z[j][i].y;
                vi[i] += beta_lo[j].re * z[j][i].y + beta_lo[j].im * z[j][i].x;
            }
        }
    if (diag0) { vr[0] = (vr[0] + gamma_lo.re) + gamma.re; vi[0] = (vi[0] + gamma_lo.im) + gamma.im; }
    if (diag1) { vr[1] = (vr[1] + gamma_lo.re) + gamma.re; vi[1] = (vi[1] + gamma_lo.im) + gamma.im; }
#pragma unroll
    for (int j = kMaxAddends - 1; j >= 0; --j)     // addend 0 carries the dominant (lowest-order) term: last
        if (has[j]) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                vr[i] = fma(beta[j].re, z[j][i].x, fma(-beta[j].im, z[j][i].y, vr[i]));
                vi[i] = fma(beta[j].re, z[j][i].y, fma(beta[j].im, z[j][i].x, vi[i]));
            }
        }
}

// All threads of the CTA call this; returns with every thread's part of D written (no trailing barrier).
// FEED = 0: every thread issues 16-byte cp.async copies (commit / wait groups).  FEED = 1: warp 0 issues one bulk copy per
// tile row (TMA engine) that completes on the stage's mbarrier `bars[stage]` (initialised by the caller, count 1; this
// variant is called once per CTA, so stage s is filled for the (ks / STAGES)-th time in iteration ks).
// MUL3 = 1: the complex product from THREE real products per fragment pair instead of four (P1 = Ar Br, P2 = Ai Bi,
// P3 = (Ar + Ai)(Br + Bi); Re = P1 - P2, Im = P3 - P1 - P2): a quarter fewer DMMAs for two additions per fragment and a
// third accumulator set.  The imaginary part loses the component-wise error bound but keeps the norm-wise one, which is
// what the tolerance of this path is stated in.
template <int BM, int BN, int WM, int WN, int FEED = 0, int MUL3 = 0, int STAGES = K4_STAGES>
__device__ __forceinline__ void tile_gemm(double2 *smem, const TileArgs &g, int tile_m, int tile_n, unsigned long long *bars = nullptr) {
    using SM = K4Smem<BM, BN, STAGES>;
    constexpr int NTHREADS = (BM / WM) * (BN / WN) * 32;
    constexpr int MT = WM / 8, NTL = WN / 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;
    const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
    const double2 *A = g.A + (size_t)tile_m * BM * g.n;   // rows tile_m*BM.., all k
    const double2 *B = g.B + (size_t)tile_n * BN;         // all k, cols tile_n*BN..

    auto load_stage = [&](int stage, int k0) {
        double2 *sA = smem + stage * SM::STAGE_ELEMS;
        double2 *sB = sA + SM::A_ELEMS;
        for (int e = tid; e < BM * K4_BK; e += NTHREADS) {
            const int r = e / K4_BK, c = e % K4_BK;
            cp_async16(sA + r * SM::PA + c, A + (size_t)r * g.n + k0 + c);
        }
        for (int e = tid; e < K4_BK * BN; e += NTHREADS) {
            const int r = e / BN, c = e % BN;
            cp_async16(sB + r * SM::PB + c, B + (size_t)(k0 + r) * g.n + c);
        }
    };

    double cre[MT][NTL][2], cim[MT][NTL][2];          // MUL3: P1 and P2 until the main loop is over
    double p3[MUL3 ? MT : 1][MUL3 ? NTL : 1][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            cre[mt][nt][0] = cre[mt][nt][1] = 0.0; cim[mt][nt][0] = cim[mt][nt][1] = 0.0;
            if (MUL3) p3[MUL3 ? mt : 0][MUL3 ? nt : 0][0] = p3[MUL3 ? mt : 0][MUL3 ? nt : 0][1] = 0.0;
        }

    const unsigned bar0 = FEED ? smem_u32(bars) : 0u;
    auto load_stage_bulk = [&](int stage, int k0) {   // warp 0, all lanes
        double2 *sA = smem + stage * SM::STAGE_ELEMS;
        double2 *sB = sA + SM::A_ELEMS;
        const unsigned bar = bar0 + 8u * stage;
        if (lane == 0) mbar_expect_tx(bar, (unsigned)((BM * K4_BK + K4_BK * BN) * sizeof(double2)));
        __syncwarp();
        for (int r = lane; r < BM; r += 32) bulk_copy_g2s(sA + r * SM::PA, A + (size_t)r * g.n + k0, K4_BK * sizeof(double2), bar);
        for (int r = lane; r < K4_BK; r += 32) bulk_copy_g2s(sB + r * SM::PB, B + (size_t)(k0 + r) * g.n, BN * sizeof(double2), bar);
    };

    const int nk = g.n / K4_BK;
    if (FEED) {
        if (warp == 0)
            for (int s = 0; s < STAGES - 1; ++s)
                if (s < nk) load_stage_bulk(s, s * K4_BK);
    } else {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < nk) load_stage(s, s * K4_BK);
            cp_async_commit();
        }
    }
    for (int ks = 0; ks < nk; ++ks) {
        if (FEED) mbar_wait(bar0 + 8u * (ks % STAGES), (unsigned)(ks / STAGES) & 1u);
        else cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // prefetch stage ks + STAGES - 1 into the buffer consumed in iteration ks - 1
            const int nxt = ks + STAGES - 1;
            if (FEED) {
                if (warp == 0 && nxt < nk) load_stage_bulk(nxt % STAGES, nxt * K4_BK);
            } else {
                if (nxt < nk) load_stage(nxt % STAGES, nxt * K4_BK);
                cp_async_commit();
            }
        }
        const double2 *sA = smem + (ks % STAGES) * SM::STAGE_ELEMS;
        const double2 *sB = sA + SM::A_ELEMS;
#pragma unroll
        for (int kt = 0; kt < K4_BK / 4; ++kt) {
            double2 af[MT], bf[NTL];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) af[mt] = sA[(wm0 + 8 * mt + gq) * SM::PA + 4 * kt + q];
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) bf[nt] = sB[(4 * kt + q) * SM::PB + wn0 + 8 * nt + gq];
            if (MUL3) {
                double as[MT], bs[NTL];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) as[mt] = af[mt].x + af[mt].y;
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) bs[nt] = bf[nt].x + bf[nt].y;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NTL; ++nt) {
                        dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].x, bf[nt].x);
                        dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].y, bf[nt].y);
                        dmma884(p3[MUL3 ? mt : 0][MUL3 ? nt : 0][0], p3[MUL3 ? mt : 0][MUL3 ? nt : 0][1], as[mt], bs[nt]);
                    }
            } else {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                    for (int nt = 0; nt < NTL; ++nt) {
                        dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].x, bf[nt].x);
                        dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].x, bf[nt].y);
                    }
#pragma unroll
                    for (int nt = 0; nt < NTL; ++nt) {
                        dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].y, -bf[nt].y);
                        dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].y, bf[nt].x);
                    }
                }
            }
        }
    }
    if (MUL3) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double a = cre[mt][nt][i], b = cim[mt][nt][i];
                    cre[mt][nt][i] = a - b;
                    cim[mt][nt][i] = p3[MUL3 ? mt : 0][MUL3 ? nt : 0][i] - (a + b);
                }
    }
    if (!FEED) cp_async_wait<0>();
    __syncthreads();   // every warp is done with the stage buffers: the next tile_gemm may refill them

    // ---- epilogue ----
    const size_t row0 = (size_t)tile_m * BM + wm0, col0 = (size_t)tile_n * BN + wn0;
    bool has[kMaxAddends];
#pragma unroll
    for (int j = 0; j < kMaxAddends; ++j) has[j] = g.C[j] != nullptr;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            const size_t r = row0 + 8 * mt + gq, c = col0 + 8 * nt + 2 * q;
            const size_t off = r * g.n + c;
            double vr[2] = {cre[mt][nt][0], cre[mt][nt][1]}, vi[2] = {cim[mt][nt][0], cim[mt][nt][1]};
            if (g.Dprod) {
                g.Dprod[off] = make_double2(vr[0], vi[0]);
                g.Dprod[off + 1] = make_double2(vr[1], vi[1]);
            }
            double2 z[kMaxAddends][2];
#pragma unroll
            for (int j = 0; j < kMaxAddends; ++j) {
                z[j][0] = z[j][1] = make_double2(0.0, 0.0);
                if (has[j]) { z[j][0] = g.C[j][off]; z[j][1] = g.C[j][off + 1]; }
            }
            if (g.Dalt) {   // second combination of the same addends (unscaled product)
                double ar[2] = {vr[0], vr[1]}, ai[2] = {vi[0], vi[1]};
#pragma unroll
                for (int j = kMaxAddends - 1; j >= 0; --j)
                    if (has[j]) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            ar[i] = fma(g.beta_alt[j].re, z[j][i].x, fma(-g.beta_alt[j].im, z[j][i].y, ar[i]));
                            ai[i] = fma(g.beta_alt[j].re, z[j][i].y, fma(g.beta_alt[j].im, z[j][i].x, ai[i]));
                        }
                    }
                g.Dalt[off] = make_double2(ar[0], ai[0]);
                g.Dalt[off + 1] = make_double2(ar[1], ai[1]);
            }
            double2 x0 = make_double2(0.0, 0.0), x1 = x0;
            if (g.C2) { x0 = g.C2[off]; x1 = g.C2[off + 1]; }
            if (g.herm && herm_tile_skipped(BM, BN, (int)(c / BM), (int)(r / BN))) {
                // the transposed position (c, r), (c + 1, r) lies in a tile nobody computes: every output there is the same
                // combination of the CONJUGATED product and addends (Hermitian operands; never on the diagonal)
                const size_t m0 = c * g.n + r, m1 = m0 + g.n;
                double wr[2] = {vr[0], vr[1]}, wi[2] = {-vi[0], -vi[1]};
                if (g.Dprod) { g.Dprod[m0] = make_double2(wr[0], wi[0]); g.Dprod[m1] = make_double2(wr[1], wi[1]); }
                double2 zc[kMaxAddends][2];
#pragma unroll
                for (int j = 0; j < kMaxAddends; ++j) { zc[j][0] = make_double2(z[j][0].x, -z[j][0].y); zc[j][1] = make_double2(z[j][1].x, -z[j][1].y); }
                if (g.Dalt) {
                    double ar[2] = {wr[0], wr[1]}, ai[2] = {wi[0], wi[1]};
#pragma unroll
                    for (int j = kMaxAddends - 1; j >= 0; --j)
                        if (has[j]) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                ar[i] = fma(g.beta_alt[j].re, zc[j][i].x, fma(-g.beta_alt[j].im, zc[j][i].y, ar[i]));
                                ai[i] = fma(g.beta_alt[j].re, zc[j][i].y, fma(g.beta_alt[j].im, zc[j][i].x, ai[i]));
                            }
                        }
                    g.Dalt[m0] = make_double2(ar[0], ai[0]);
                    g.Dalt[m1] = make_double2(ar[1], ai[1]);
                }
                epilogue_pair(wr, wi, zc, has, g.scaled, g.alpha, g.beta, g.beta_lo, g.gamma, g.gamma_lo, false, false);
                g.D[m0] = make_double2(wr[0], wi[0]);
                g.D[m1] = make_double2(wr[1], wi[1]);
            }
            epilogue_pair(vr, vi, z, has, g.scaled, g.alpha, g.beta, g.beta_lo, g.gamma, g.gamma_lo, r == c, r == c + 1);
            if (g.C2) { vr[0] += g.beta2 * x0.x; vi[0] += g.beta2 * x0.y; vr[1] += g.beta2 * x1.x; vi[1] += g.beta2 * x1.y; }
            g.D[off] = make_double2(vr[0], vi[0]);
            g.D[off + 1] = make_double2(vr[1], vi[1]);
        }
}

template <int BM, int BN, int WM, int WN, int FEED = 0, int MUL3 = 0, int STAGES = K4_STAGES>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
k4_zgemm_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) unsigned char k4_smem_raw[];
    double2 *smem = reinterpret_cast<double2 *>(k4_smem_raw);
    __shared__ __align__(8) unsigned long long bars[K4_STAGES];
    if (FEED) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < K4_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        __syncthreads();
    }
    const long long b = blockIdx.y;
    const int tiles_n = g.n / BN;
    TileArgs t;
    t.A = g.A + b * g.strideA;
    t.B = g.B + b * g.strideB;
#pragma unroll
    for (int j = 0; j < kMaxAddends; ++j) {
        t.C[j] = g.C[j] ? g.C[j] + b * g.strideC[j] : nullptr;
        t.beta[j] = g.beta[j]; t.beta_lo[j] = g.beta_lo[j];
    }
    t.C2 = g.C2 ? g.C2 + b * g.strideC2 : nullptr;
    t.D = g.D + b * g.strideD;
    t.Dprod = g.Dprod ? g.Dprod + b * g.strideDprod : nullptr;
    t.Dalt = g.Dalt ? g.Dalt + b * g.strideDalt : nullptr;
#pragma unroll
    for (int j = 0; j < kMaxAddends; ++j) t.beta_alt[j] = g.beta_alt[j];
    t.alpha = g.alpha; t.scaled = g.scaled;
    t.beta2 = g.beta2; t.gamma = g.gamma; t.gamma_lo = g.gamma_lo;
    t.n = g.n;
    t.herm = g.herm;
    int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
    if (g.herm) herm_tile_at(BM, BN, g.n, (int)blockIdx.x, tile_m, tile_n);   // blockIdx.x enumerates the tiles that are not skipped (plan.hpp)
    tile_gemm<BM, BN, WM, WN, FEED, MUL3, STAGES>(smem, t, tile_m, tile_n, bars);
}

// ------------------------------------------------------------------------------------------------
// persistent chain kernel (npad == BM): scratch per CTA: the kSeriesSlots series slots, then 2 slots of the running product
// ------------------------------------------------------------------------------------------------
template <int BM, int WM, int WN, typename IO>
__global__ void __launch_bounds__((BM / WM) * (BM / WN) * 32)
k4_chain_kernel(const SeriesParams p, const SeriesProgram prog, const IO *__restrict__ carr, const double2 *__restrict__ H,
                double2 *__restrict__ scratch, double2 *__restrict__ partials, unsigned long long nsteps) {
    constexpr int NTHREADS = (BM / WM) * (BM / WN) * 32;
    constexpr int NN = BM * BM;
    extern __shared__ __align__(16) unsigned char k4_smem_raw[];
    double2 *smem = reinterpret_cast<double2 *>(k4_smem_raw);
    __shared__ cplx coef[kMaxTerms];

    const int tid = threadIdx.x;
    double2 *slot = scratch + (size_t)blockIdx.x * (kSeriesSlots + 2) * NN;
    const unsigned long long lo = nsteps * blockIdx.x / gridDim.x, hi = nsteps * (blockIdx.x + 1) / gridDim.x;
    int f_cur = kSeriesSlots;
    bool have_f = false;

    for (unsigned long long j = lo; j < hi; ++j) {
        for (int t = tid; t < p.nterms; t += NTHREADS)
            coef[t] = step_coefficient<IO>(p.terms[t], carr, p.pts, p.quad, p.magfac, j);
        __syncthreads();
        // ---- assemble Y (slot 0) and the series start values (slots 2, 3) ----
        for (int e = tid; e < NN; e += NTHREADS) {
            double2 x = __ldg(H + e);
            for (int t = 0; t < p.nterms; ++t) {
                const double2 h = __ldg(H + (size_t)p.terms[t].mat * NN + e);
                const cplx ct = coef[t];
                x.x += ct.re * h.x - ct.im * h.y;
                x.y += ct.re * h.y + ct.im * h.x;
            }
            const double yr = x.x * p.sigma, yi = x.y * p.sigma;
            const bool diag = (e / BM == e % BM);
            slot[e] = make_double2(yr, yi);
            slot[4 * NN + e] = make_double2(((prog.u.re * yr - prog.u.im * yi) + (diag ? prog.v_lo.re : 0.0)) + (diag ? prog.v.re : 0.0),
                                            ((prog.u.re * yi + prog.u.im * yr) + (diag ? prog.v_lo.im : 0.0)) + (diag ? prog.v.im : 0.0));
            if (prog.init5) slot[5 * NN + e] = make_double2(diag ? prog.w.re : 0.0, diag ? prog.w.im : 0.0);
        }
        __syncthreads();
        // ---- series program ----
        for (int o = 0; o < prog.nops; ++o) {
            const SeriesOp &op = prog.ops[o];
            TileArgs t;
            t.A = slot + (size_t)op.A * NN;
            t.B = slot + (size_t)op.B * NN;
#pragma unroll
            for (int a = 0; a < kMaxAddends; ++a) {
                t.C[a] = op.C[a] >= 0 ? slot + (size_t)op.C[a] * NN : nullptr;
                t.beta[a] = op.beta[a]; t.beta_lo[a] = op.beta_lo[a];
            }
            t.C2 = nullptr;
            t.D = slot + (size_t)op.D * NN;
            t.Dprod = op.Dprod >= 0 ? slot + (size_t)op.Dprod * NN : nullptr;
            t.Dalt = op.Dalt >= 0 ? slot + (size_t)op.Dalt * NN : nullptr;
#pragma unroll
            for (int a = 0; a < kMaxAddends; ++a) t.beta_alt[a] = op.beta_alt[a];
            t.alpha = op.alpha; t.scaled = op.scaled;
            t.beta2 = 0.0; t.gamma = op.gamma; t.gamma_lo = op.gamma_lo;
            t.n = BM;
            t.herm = 0;
            tile_gemm<BM, BM, WM, WN>(smem, t, 0, 0);
            __syncthreads();
        }
        // ---- running product in E-form:  F <- E + F + E F  (later step on the left) ----
        const double2 *E = slot + (size_t)prog.e_slot * NN;
        if (!have_f) {
            for (int e = tid; e < NN; e += NTHREADS) slot[(size_t)f_cur * NN + e] = E[e];
            have_f = true;
        } else {
            const int f_nxt = (f_cur == kSeriesSlots) ? kSeriesSlots + 1 : kSeriesSlots;
            TileArgs t{};
            t.A = E; t.B = slot + (size_t)f_cur * NN; t.C[0] = E; t.C2 = slot + (size_t)f_cur * NN;
            t.D = slot + (size_t)f_nxt * NN;
            t.beta[0] = cplx{1.0, 0.0}; t.beta2 = 1.0;
            t.n = BM;
            tile_gemm<BM, BM, WM, WN>(smem, t, 0, 0);
            f_cur = f_nxt;
        }
        __syncthreads();
    }
    double2 *out = partials + (size_t)blockIdx.x * NN;
    for (int e = tid; e < NN; e += NTHREADS) out[e] = have_f ? slot[(size_t)f_cur * NN + e] : make_double2(0.0, 0.0);
}

// Batched assembly.  grid = (npad*npad/256, ceil(S/8)); H: [mat][npad*npad] row-major, zero padded.
template <typename IO>
__global__ void __launch_bounds__(256)
k4_assemble_kernel(const SeriesParams p, const SeriesProgram prog, const IO *__restrict__ carr, const double2 *__restrict__ H,
                   double2 *__restrict__ Y, double2 *__restrict__ S4, double2 *__restrict__ S5,
                   unsigned long long step0, int S) {
    __shared__ cplx coef[kMaxTerms][8];
    const int sg0 = blockIdx.y * 8;
    const int ns = min(8, S - sg0);
    for (int i = threadIdx.x; i < p.nterms * 8; i += blockDim.x) {
        const int t = i / 8, s = i % 8;
        if (s < ns) coef[t][s] = step_coefficient<IO>(p.terms[t], carr, p.pts, p.quad, p.magfac, step0 + sg0 + s);
    }
    __syncthreads();
    const size_t nn = (size_t)p.npad * p.npad;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nn) return;
    const int r = (int)(e / p.npad), c = (int)(e % p.npad);
    const double2 h0 = H[e];
    double xr[8], xi[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) { xr[s] = h0.x; xi[s] = h0.y; }
    for (int t = 0; t < p.nterms; ++t) {
        const double2 h = H[(size_t)p.terms[t].mat * nn + e];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const cplx ct = coef[t][s];
            xr[s] += ct.re * h.x - ct.im * h.y;
            xi[s] += ct.re * h.y + ct.im * h.x;
        }
    }
    const bool diag = (r == c);
    // slot 4 = u Y + v I is the start value of the Clenshaw / Horner recurrences only; the product-saving forms write slot 4 themselves
    // before they read it (build_program), so its 16 bytes per element and step -- as much as Y itself -- are not written for them
    const bool init4 = prog.u.re != 0.0 || prog.u.im != 0.0 || prog.v.re != 0.0 || prog.v.im != 0.0 || prog.v_lo.re != 0.0 || prog.v_lo.im != 0.0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        if (s < ns) {
            const size_t o = (size_t)(sg0 + s) * nn + e;
            const double yr = xr[s] * p.sigma, yi = xi[s] * p.sigma;
            Y[o] = make_double2(yr, yi);
            if (init4)
            S4[o] = make_double2(((prog.u.re * yr - prog.u.im * yi) + (diag ? prog.v_lo.re : 0.0)) + (diag ? prog.v.re : 0.0),
                                 ((prog.u.re * yi + prog.u.im * yr) + (diag ? prog.v_lo.im : 0.0)) + (diag ? prog.v.im : 0.0));
            if (prog.init5) S5[o] = make_double2(diag ? prog.w.re : 0.0, diag ? prog.w.im : 0.0);
        }
    }
}

// out[r][c] (n x n, IO precision) = I + E[r][c]  (E padded to npad)
template <typename IO>
__global__ void k4_finish_kernel(const double2 *__restrict__ E, int n, int npad, IO *__restrict__ out, int add_identity) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * n) return;
    const int r = e / n, c = e % n;
    const double2 v = E[(size_t)r * npad + c];
    IO o;
    o.x = v.x + ((add_identity && r == c) ? 1.0 : 0.0);
    o.y = v.y;
    out[e] = o;
}

// E[b] (npad x npad, zero padded) = P[b] - I  for `count` propagators in the IO precision (multi-GPU combine input)
template <typename IO>
__global__ void k4_eform_kernel(const IO *__restrict__ P, int n, int npad, int count, double2 *__restrict__ E) {
    const size_t nn = (size_t)npad * npad;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nn * count) return;
    const int b = (int)(e / nn), r = (int)((e % nn) / npad), c = (int)(e % npad);
    double2 v = make_double2(0.0, 0.0);
    if (r < n && c < n) {
        const IO x = P[((size_t)b * n + r) * n + c];
        v = make_double2((double)x.x - (r == c ? 1.0 : 0.0), (double)x.y);
    }
    E[e] = v;
}

// max_j |c_k(j)|^2 per control k over every pulse of the call (spectral bound of the step Hamiltonians, api.cu series_norm_for_call).
// grid = (blocks, amps); out[k] holds the bits of a non-negative double, which order like unsigned integers.
template <typename IO>
__global__ void __launch_bounds__(256)
k4_absmax_kernel(const IO *__restrict__ carr, unsigned int batch, unsigned int amps, size_t stride, size_t pts,
                 unsigned long long *__restrict__ out) {
    const unsigned int k = blockIdx.y;
    double best = 0.0;
    bool cplx_seen = false;   // any amplitude with a non-zero (or NaN) imaginary part -> out[amps] != 0
    for (unsigned int b = 0; b < batch; ++b) {
        const IO *c = carr + ((size_t)b * amps + k) * stride;
        for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < pts; j += (size_t)gridDim.x * blockDim.x) {
            const IO v = __ldg(c + j);
            const double m = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            best = m > best ? m : best;   // a NaN amplitude never wins: the propagator will be NaN anyway
            cplx_seen = cplx_seen || !(v.y == 0);
        }
    }
    for (int o = 16; o > 0; o >>= 1) { const double other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
    if ((threadIdx.x & 31) == 0 && best > 0.0) atomicMax(out + k, (unsigned long long)__double_as_longlong(best));
    if (__any_sync(0xffffffffu, cplx_seen) && (threadIdx.x & 31) == 0) atomicMax(out + amps, 1ull);
}

cudaError_t k4_absmax(bool fp64_io, const void *carr, unsigned int batch, unsigned int amps, size_t stride, size_t pts,
                      unsigned long long *out_dev, cudaStream_t stream) {
    if (amps == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(out_dev, 0, ((size_t)amps + 1) * sizeof(unsigned long long), stream);   // [amps]: complex amplitudes seen
    if (e != cudaSuccess) return e;
    const size_t work = pts * batch;
    const unsigned int bx = (unsigned int)std::max<size_t>(1, std::min<size_t>(592, (work + 2047) / 2048));
    dim3 grid(bx, amps);
    if (fp64_io) k4_absmax_kernel<double2><<<grid, 256, 0, stream>>>((const double2 *)carr, batch, amps, stride, pts, out_dev);
    else         k4_absmax_kernel<float2><<<grid, 256, 0, stream>>>((const float2 *)carr, batch, amps, stride, pts, out_dev);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// The series as a program over slots (k4_gemm.hpp).
//   p.horner == 0  Clenshaw: B_k = B_{k+1} Y - B_{k+2} + a_k I, E = B_1 Y - 2 B_2 + a0' I   (parament.cpp:569-652, E-form)
//   p.horner == 1  Horner in W = Y^2 on the monomial coefficients c_m of the same polynomial (p.a = c):
//                  R <- R W + c_{2i+1} Y + c_{2i} I                                   1 + floor(M/2) products
//   p.horner == 2  Paterson-Stockmeyer with V = Y^4: Y^2, Y^3, Y^4, then
//                  R <- R V + c_{4i+3} Y^3 + c_{4i+2} Y^2 + c_{4i+1} Y + c_{4i} I      3 + floor(M/4) products
//   p.horner == 3  degree 8 as (y02 + ...)(y02 + ...) + ... with y02 = Y^2 (...)            3 products (api.cu solve_degree8)
//   p.horner == 4  degree 12 as (y0 + ...)(y0 + ...) + ... with y0 = Y^3 (...)            4 products (api.cu solve_degree12)
SeriesProgram build_program(const SeriesParams &p) {
    SeriesProgram g{};
    const int M = p.M;
    const cplx zero{0.0, 0.0}, one{1.0, 0.0};
    auto push = [&](int A, int B, int D) -> SeriesOp & {
        SeriesOp &o = g.ops[g.nops++];
        o = SeriesOp{};
        o.A = A; o.B = B; o.D = D; o.Dprod = -1; o.Dalt = -1; o.scaled = 0; o.alpha = one;
        for (int j = 0; j < kMaxAddends; ++j) { o.C[j] = -1; o.beta[j] = zero; o.beta_lo[j] = zero; o.beta_alt[j] = zero; }
        o.gamma = zero; o.gamma_lo = zero;
        return o;
    };
    auto coef = [&](int m) { return m <= M ? p.a[m] : zero; };
    auto coef_lo = [&](int m) { return m <= M ? p.a_lo[m] : zero; };
    if (p.horner == 3) {
        // degree 8 in three products (api.cu solve_degree8): p.a[k].re = c4 c3 d2 d1 e2 e0 r2' r1 r0, A = -i Y, A2 = -W
        auto re = [&](int k, double sgn) { return cplx{sgn * p.a[k].re, 0.0}; };
        auto im = [&](int k, double sgn) { return cplx{0.0, sgn * p.a[k].re}; };
        auto re_lo = [&](int k, double sgn) { return cplx{sgn * p.a_lo[k].re, 0.0}; };
        auto im_lo = [&](int k, double sgn) { return cplx{0.0, sgn * p.a_lo[k].re}; };
        {   // W = Y Y -> slot 1;  T = c4 W + i c3 Y -> slot 3
            SeriesOp &o = push(0, 0, 3);
            o.Dprod = 1;
            o.scaled = 1; o.alpha = re(0, 1.0);
            o.C[0] = 0; o.beta[0] = im(1, 1.0);
        }
        {   // y02 = T W;  L = y02 - d2 W - i d1 Y + e0 I -> slot 4;  R = y02 - e2 W -> slot 5
            SeriesOp &o = push(3, 1, 4);
            o.C[0] = 0; o.beta[0] = im(3, -1.0);
            o.C[1] = 1; o.beta[1] = re(2, -1.0); o.beta_alt[1] = re(4, -1.0);
            o.gamma = re(5, 1.0);
            o.Dalt = 5;
        }
        {   // E = L R - r2' W - i r1 Y + r0 I -> slot 3
            SeriesOp &o = push(4, 5, 3);
            o.C[0] = 0; o.beta[0] = im(7, -1.0); o.beta_lo[0] = im_lo(7, -1.0);
            o.C[1] = 1; o.beta[1] = re(6, -1.0); o.beta_lo[1] = re_lo(6, -1.0);
            o.gamma = re(8, 1.0); o.gamma_lo = re_lo(8, 1.0);
        }
        g.u = zero; g.v = zero; g.v_lo = zero; g.w = zero; g.init5 = 0;
        g.e_slot = 3;
    } else if (p.horner == 4) {
        // degree 12 in four products (api.cu solve_degree12): p.a[k].re = tV tW tY lV lW lY lI rV rW sV sW sY sI
        auto re = [&](int k) { return cplx{p.a[k].re, 0.0}; };
        auto im = [&](int k) { return cplx{0.0, p.a[k].re}; };
        auto re_lo = [&](int k) { return cplx{p.a_lo[k].re, 0.0}; };
        auto im_lo = [&](int k) { return cplx{0.0, p.a_lo[k].re}; };
        push(0, 0, 1);                                             // W = Y Y
        {   // V = W Y  ->  slot 2;  T' = tV V + i tW W + tY Y  ->  slot 3
            SeriesOp &o = push(1, 0, 3);
            o.Dprod = 2;
            o.scaled = 1; o.alpha = re(0);
            o.C[0] = 0; o.beta[0] = re(2);
            o.C[1] = 1; o.beta[1] = im(1);
        }
        {   // y0 = T' V;  L = y0 + i lV V + lW W + i lY Y + lI I  ->  slot 4;  R = y0 + i rV V + rW W  ->  slot 5
            SeriesOp &o = push(3, 2, 4);
            o.C[0] = 0; o.beta[0] = im(5);
            o.C[1] = 1; o.beta[1] = re(4); o.beta_alt[1] = re(8);
            o.C[2] = 2; o.beta[2] = im(3); o.beta_alt[2] = im(7);
            o.gamma = re(6);
            o.Dalt = 5;
        }
        {   // E = L R + i sV V + sW W + i sY Y + sI I  ->  slot 3
            SeriesOp &o = push(4, 5, 3);
            o.C[0] = 0; o.beta[0] = im(11); o.beta_lo[0] = im_lo(11);
            o.C[1] = 1; o.beta[1] = re(10); o.beta_lo[1] = re_lo(10);
            o.C[2] = 2; o.beta[2] = im(9);  o.beta_lo[2] = im_lo(9);
            o.gamma = re(12); o.gamma_lo = re_lo(12);
        }
        g.u = zero; g.v = zero; g.v_lo = zero; g.w = zero; g.init5 = 0;
        g.e_slot = 3;
    } else if (p.horner == 2) {
        const int L = M >> 2;                                      // top block index
        push(0, 0, 1);                                             // Y^2
        // Y^3 = Y^2 Y, and the top block R_L = c_{4L+3} Y^3 + c_{4L+2} Y^2 + c_{4L+1} Y + c_{4L} I from the same product
        {
            SeriesOp &o = push(1, 0, 4);
            o.Dprod = 2;
            o.scaled = 1; o.alpha = coef(4 * L + 3);
            o.C[0] = 0; o.beta[0] = coef(4 * L + 1);
            o.C[1] = 1; o.beta[1] = coef(4 * L + 2);
            o.gamma = coef(4 * L);
        }
        push(1, 1, 3);                                             // Y^4
        int cur = 4;
        for (int i = L - 1; i >= 0; --i) {
            SeriesOp &o = push(cur, 3, cur ^ 1);
            o.C[0] = 0; o.beta[0] = coef(4 * i + 1); o.beta_lo[0] = coef_lo(4 * i + 1);
            o.C[1] = 1; o.beta[1] = coef(4 * i + 2); o.beta_lo[1] = coef_lo(4 * i + 2);
            o.C[2] = 2; o.beta[2] = coef(4 * i + 3); o.beta_lo[2] = coef_lo(4 * i + 3);
            o.gamma = coef(4 * i); o.gamma_lo = coef_lo(4 * i);
            cur ^= 1;
        }
        g.u = zero; g.v = zero; g.v_lo = zero; g.w = zero; g.init5 = 0;   // slot 4 is produced by the second op
        g.e_slot = cur;
    } else if (p.horner == 1) {
        const int L = M >> 1;
        g.u = coef(2 * L + 1); g.v = coef(2 * L); g.v_lo = coef_lo(2 * L); g.init5 = 0; g.w = zero;
        push(0, 0, 1);                                             // W = Y Y
        int cur = 4;
        for (int i = L - 1; i >= 0; --i) {
            SeriesOp &o = push(cur, 1, cur ^ 1);
            o.C[0] = 0; o.beta[0] = p.a[2 * i + 1]; o.beta_lo[0] = p.a_lo[2 * i + 1];
            o.gamma = p.a[2 * i]; o.gamma_lo = p.a_lo[2 * i];
            cur ^= 1;
        }
        g.e_slot = cur;
    } else if (M == 1) {
        g.u = p.a[1]; g.v = p.a[0]; g.v_lo = p.a_lo[0]; g.init5 = 0; g.w = zero;
        g.e_slot = 4;
    } else {
        g.u = p.a[M]; g.v = p.a[M - 1]; g.v_lo = p.a_lo[M - 1]; g.init5 = 1; g.w = p.a[M];
        int cur = 4;                                               // slot 4 = B_{M-1}, slot 5 = B_M
        for (int k = M - 2; k >= 0; --k) {
            SeriesOp &o = push(cur, 0, cur ^ 1);
            o.C[0] = cur ^ 1; o.beta[0] = cplx{k == 0 ? -2.0 : -1.0, 0.0};
            o.gamma = p.a[k]; o.gamma_lo = p.a_lo[k];
            cur ^= 1;
        }
        g.e_slot = cur;
    }
    return g;
}

template <typename K>
static cudaError_t opt_in_smem(K kern, size_t bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// operand feed of the batched GEMM: 0 = cp.async by all threads, 1 = bulk copies (TMA engine) by one warp + mbarriers
static int k4_feed_mode() {
    static const int mode = [] { const char *e = getenv("PARAMENT_K4_FEED"); return e ? (strcmp(e, "tma") == 0 ? 1 : 0) : 0; }();
    return mode;
}

// Complex product of the batched GEMM ($PARAMENT_K4_3M, read once).  Measured at dim 256 (C4, 2e4 points, four chunk streams):
//   0  four real products, 64x64 CTA tiles, 3 stages, 2 CTAs of 8 warps per SM                      4.50e4 steps/s
//   1  three real products, 64x64 tiles, one CTA per SM (the third accumulator set: 156 registers)  4.37e4
//   2  three real products, 64x32 tiles, 3 stages, 2 CTAs of 4 warps per SM                         4.75e4
//   3  three real products, 64x32 tiles, 2 stages, 3 CTAs of 4 warps per SM (default)               5.45e4
static int k4_mul3_mode() {
    static const int mode = [] { const char *e = getenv("PARAMENT_K4_3M"); const int v = e ? atoi(e) : 3; return v >= 0 && v <= 3 ? v : 3; }();
    return mode;
}

template <int BM, int BN, int WM, int WN, int FEED, int MUL3, int STAGES = K4_STAGES>
static cudaError_t launch_gemm_tt(const GemmArgs &g, cudaStream_t stream) {
    using SM = K4Smem<BM, BN, STAGES>;
    auto kern = k4_zgemm_kernel<BM, BN, WM, WN, FEED, MUL3, STAGES>;
    cudaError_t e = opt_in_smem(kern, SM::BYTES);   // per-device function attribute; cheap, so set on every launch
    if (e != cudaSuccess) return e;
    // (programmatic dependent launch was measured here: 2.78e4 -> 2.50e4 steps/s at dim 256, so plain stream order is kept)
    dim3 grid(g.herm ? herm_tile_count(BM, BN, g.n) : (g.n / BM) * (g.n / BN), g.batch);
    kern<<<grid, (BM / WM) * (BN / WN) * 32, SM::BYTES, stream>>>(g);
    return cudaGetLastError();
}

// four real products per complex product (mode 0), with either operand feed
template <int BM, int BN, int WM, int WN>
static cudaError_t launch_gemm_t(const GemmArgs &g, cudaStream_t stream) {
    return k4_feed_mode() ? launch_gemm_tt<BM, BN, WM, WN, 1, 0>(g, stream) : launch_gemm_tt<BM, BN, WM, WN, 0, 0>(g, stream);
}

cudaError_t k4_gemm(const GemmArgs &g, cudaStream_t stream) {
    if (g.batch <= 0) return cudaSuccess;
    if (g.n % 64 == 0) {
        if (k4_mul3_mode() == 1) return launch_gemm_tt<64, 64, 32, 16, 0, 1>(g, stream);
        if (k4_mul3_mode() == 2) return launch_gemm_tt<64, 32, 32, 16, 0, 1>(g, stream);
        if (k4_mul3_mode() == 3) return launch_gemm_tt<64, 32, 32, 16, 0, 1, 2>(g, stream);
        return launch_gemm_t<64, 64, 32, 16>(g, stream);
    }
    return launch_gemm_t<32, 32, 16, 16>(g, stream);
}

int k4_pad(int n) { return n <= 32 ? 32 : ((n + 63) / 64) * 64; }
int k4_real_products(int npad) { return (npad % 64 == 0 && k4_mul3_mode() != 0) ? 3 : 4; }
int k4_herm_tiles(int npad) {   // tiles of a Hermitian-output launch (GemmArgs::herm) in the tile shape k4_gemm selects
    if (npad % 64 != 0) return herm_tile_count(32, 32, npad);
    return k4_mul3_mode() >= 2 ? herm_tile_count(64, 32, npad) : herm_tile_count(64, 64, npad);
}
int k4_tiles(int npad) {
    if (npad % 64 != 0) return 1;
    return (npad / 64) * (npad / (k4_mul3_mode() >= 2 ? 32 : 64));
}

template <typename K>
static int k4_blocks_per_sm(K kern, int threads, size_t smem) {
    int per_sm = 0;
    opt_in_smem(kern, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    return per_sm;
}

int k4_wave_slots(int npad, int num_sms) {
    int per_sm = 0;
    if (npad % 64 == 0) {
        if (k4_mul3_mode() == 1) per_sm = k4_blocks_per_sm(k4_zgemm_kernel<64, 64, 32, 16, 0, 1>, 256, K4Smem<64, 64>::BYTES);
        else if (k4_mul3_mode() == 2) per_sm = k4_blocks_per_sm(k4_zgemm_kernel<64, 32, 32, 16, 0, 1>, 128, K4Smem<64, 32>::BYTES);
        else if (k4_mul3_mode() == 3) per_sm = k4_blocks_per_sm(k4_zgemm_kernel<64, 32, 32, 16, 0, 1, 2>, 128, K4Smem<64, 32, 2>::BYTES);
        else per_sm = k4_blocks_per_sm(k4_zgemm_kernel<64, 64, 32, 16>, 256, K4Smem<64, 64>::BYTES);
    } else {
        per_sm = k4_blocks_per_sm(k4_zgemm_kernel<32, 32, 16, 16>, 128, K4Smem<32, 32>::BYTES);
    }
    if (per_sm < 1) per_sm = 1;
    return per_sm * num_sms;
}

int k4_chain_slots(int npad, int num_sms) {
    int per_sm = 0;
    if (npad == 64) {
        opt_in_smem(k4_chain_kernel<64, 32, 16, double2>, K4Smem<64, 64>::BYTES);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4_chain_kernel<64, 32, 16, double2>, 256, K4Smem<64, 64>::BYTES);
    } else {
        opt_in_smem(k4_chain_kernel<32, 16, 16, double2>, K4Smem<32, 32>::BYTES);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4_chain_kernel<32, 16, 16, double2>, 128, K4Smem<32, 32>::BYTES);
    }
    if (per_sm < 1) per_sm = 1;
    return per_sm * num_sms;
}

template <int BM, int WM, int WN, typename IO>
static cudaError_t launch_chain_t(const SeriesParams &p, const SeriesProgram &prog, const IO *carr, const double2 *H,
                                  double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream) {
    using SM = K4Smem<BM, BM>;
    auto kern = k4_chain_kernel<BM, WM, WN, IO>;
    cudaError_t e = opt_in_smem(kern, SM::BYTES);
    if (e != cudaSuccess) return e;
    kern<<<grid, (BM / WM) * (BM / WN) * 32, SM::BYTES, stream>>>(p, prog, carr, H, scratch, partials, nsteps);
    return cudaGetLastError();
}

cudaError_t k4_chain(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                     double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream) {
    if (p.npad == 64)
        return fp64_io ? launch_chain_t<64, 32, 16, double2>(p, prog, (const double2 *)carr, H, scratch, partials, nsteps, grid, stream)
                       : launch_chain_t<64, 32, 16, float2>(p, prog, (const float2 *)carr, H, scratch, partials, nsteps, grid, stream);
    return fp64_io ? launch_chain_t<32, 16, 16, double2>(p, prog, (const double2 *)carr, H, scratch, partials, nsteps, grid, stream)
                   : launch_chain_t<32, 16, 16, float2>(p, prog, (const float2 *)carr, H, scratch, partials, nsteps, grid, stream);
}

cudaError_t k4_assemble(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                        double2 *slot0, double2 *slot4, double2 *slot5, unsigned long long step0, int S, cudaStream_t stream) {
    const size_t nn = (size_t)p.npad * p.npad;
    dim3 grid((unsigned)((nn + 255) / 256), (unsigned)((S + 7) / 8));
    if (fp64_io)
        k4_assemble_kernel<double2><<<grid, 256, 0, stream>>>(p, prog, (const double2 *)carr, H, slot0, slot4, slot5, step0, S);
    else
        k4_assemble_kernel<float2><<<grid, 256, 0, stream>>>(p, prog, (const float2 *)carr, H, slot0, slot4, slot5, step0, S);
    return cudaGetLastError();
}

cudaError_t k4_eform(bool fp64_io, const void *P, int n, int npad, int count, double2 *E, cudaStream_t stream) {
    const size_t total = (size_t)npad * npad * count;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (fp64_io) k4_eform_kernel<double2><<<blocks, 256, 0, stream>>>((const double2 *)P, n, npad, count, E);
    else         k4_eform_kernel<float2><<<blocks, 256, 0, stream>>>((const float2 *)P, n, npad, count, E);
    return cudaGetLastError();
}

cudaError_t k4_finish(bool fp64_io, const double2 *E, int n, int npad, void *out, bool add_identity, cudaStream_t stream) {
    const int blocks = (n * n + 255) / 256;
    if (fp64_io)
        k4_finish_kernel<double2><<<blocks, 256, 0, stream>>>(E, n, npad, (double2 *)out, add_identity ? 1 : 0);
    else
        k4_finish_kernel<float2><<<blocks, 256, 0, stream>>>(E, n, npad, (float2 *)out, add_identity ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace pb
