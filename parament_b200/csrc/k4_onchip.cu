// k4_onchip.cu -- dim 33..64: the whole per-step pipeline of one CTA with its operands RESIDENT in shared memory.
//
// One persistent CTA (256 threads, 8 warps of 32x16 DMMA tiles) per SM walks a contiguous range of time steps.  Three
// 64x64 complex128 buffers (pitch 68, 209 KB) hold, per step, Y / W = Y^2 / the recurrence matrices; every product of the
// series reads both operands from shared memory with LDS.128 straight into mma.sync.m8n8k4.f64 fragments and writes its
// result back to shared memory in the epilogue, so inside a step nothing but the Horner addend Y (epilogue, thread-private)
// and the running product F touch L2, and there is no cp.async pipeline to fill or drain between the ~7 dependent products.
// Replaces, for these dimensions, the MMAX batched cuBLAS GEMMs + diagonal_add launches of parament.cpp:569-652 and the
// reduction of parament.cpp:657-718; supersedes k4_chain_kernel (operands streamed from an L2 scratch) for npad == 64.
#include "coef.cuh"
#include "k4_gemm.hpp"

namespace pb {

constexpr int OC_N = 64;            // padded dimension
constexpr int OC_P = OC_N + 4;      // pitch in double2: rows 8 apart fall into different bank groups for the A-fragment loads
constexpr int OC_BUF = OC_N * OC_P; // elements per buffer
constexpr int OC_THREADS = 256;
constexpr size_t OC_SMEM = 3 * (size_t)OC_BUF * sizeof(double2);

struct OcEpilogue {
    const double2 *c1_smem;    // addend read from a shared-memory buffer (own elements), or null
    const double2 *c1_glob;    // addend read from a global row-major matrix, or null
    const double2 *c2_glob;    // second global addend, or null
    cplx beta1, beta1_lo;
    double beta2;
    cplx gamma, gamma_lo;
    double2 *d_smem;           // destination buffer in shared memory, or null
    double2 *d_glob;           // destination in global memory (row-major, pitch OC_N), or null
};

// D = A * B + addends, A and B in shared memory (pitch OC_P).  All 256 threads; no barrier inside.
__device__ __forceinline__ void oc_gemm(const double2 *__restrict__ sA, const double2 *__restrict__ sB, const OcEpilogue &ep) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 16;
    constexpr int MT = 4, NTL = 2;

    // the global addend is thread-private: fetch it before the main loop so its latency hides behind the products
    double2 y[MT][NTL][2];
    if (ep.c1_glob) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q;
                y[mt][nt][0] = ep.c1_glob[r * OC_N + c];
                y[mt][nt][1] = ep.c1_glob[r * OC_N + c + 1];
            }
    }
    double cre[MT][NTL][2], cim[MT][NTL][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) { cre[mt][nt][0] = cre[mt][nt][1] = 0.0; cim[mt][nt][0] = cim[mt][nt][1] = 0.0; }

#pragma unroll 4
    for (int kt = 0; kt < OC_N / 4; ++kt) {
        double2 af[MT], bf[NTL];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) af[mt] = sA[(wm0 + 8 * mt + gq) * OC_P + 4 * kt + q];
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) bf[nt] = sB[(4 * kt + q) * OC_P + wn0 + 8 * nt + gq];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].x, bf[nt].x);
                dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].x, bf[nt].y);
            }
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].y, neg(bf[nt].y));
                dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].y, bf[nt].x);
            }
        }
    }

    // epilogue (small terms first, DESIGN.md "Numerics")
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q;
            double vr[2] = {cre[mt][nt][0], cre[mt][nt][1]}, vi[2] = {cim[mt][nt][0], cim[mt][nt][1]};
            if (ep.c2_glob) {
                const double2 x0 = ep.c2_glob[r * OC_N + c], x1 = ep.c2_glob[r * OC_N + c + 1];
                vr[0] += ep.beta2 * x0.x; vi[0] += ep.beta2 * x0.y; vr[1] += ep.beta2 * x1.x; vi[1] += ep.beta2 * x1.y;
            }
            double2 z[2] = {make_double2(0, 0), make_double2(0, 0)};
            const bool has_c1 = ep.c1_glob || ep.c1_smem;
            if (ep.c1_glob) { z[0] = y[mt][nt][0]; z[1] = y[mt][nt][1]; }
            else if (ep.c1_smem) { z[0] = ep.c1_smem[r * OC_P + c]; z[1] = ep.c1_smem[r * OC_P + c + 1]; }
            if (has_c1) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    vr[i] += ep.beta1_lo.re * z[i].x - ep.beta1_lo.im * z[i].y;
                    vi[i] += ep.beta1_lo.re * z[i].y + ep.beta1_lo.im * z[i].x;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (r == c + i) { vr[i] = (vr[i] + ep.gamma_lo.re) + ep.gamma.re; vi[i] = (vi[i] + ep.gamma_lo.im) + ep.gamma.im; }
            if (has_c1) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    vr[i] = fma(ep.beta1.re, z[i].x, fma(-ep.beta1.im, z[i].y, vr[i]));
                    vi[i] = fma(ep.beta1.re, z[i].y, fma(ep.beta1.im, z[i].x, vi[i]));
                }
            }
            if (ep.d_smem) {
                ep.d_smem[r * OC_P + c] = make_double2(vr[0], vi[0]);
                ep.d_smem[r * OC_P + c + 1] = make_double2(vr[1], vi[1]);
            }
            if (ep.d_glob) {
                ep.d_glob[r * OC_N + c] = make_double2(vr[0], vi[0]);
                ep.d_glob[r * OC_N + c + 1] = make_double2(vr[1], vi[1]);
            }
        }
}

// scratch per CTA: 3 matrices (Y, F0, F1), row-major pitch 64 (the host allocates kSeriesSlots + 2).
template <typename IO>
__global__ void __launch_bounds__(OC_THREADS, 1)
k4_onchip_kernel(const SeriesParams p, const SeriesProgram prog, const IO *__restrict__ carr, const double2 *__restrict__ H,
                 double2 *__restrict__ scratch, double2 *__restrict__ partials, unsigned long long nsteps) {
    constexpr int NN = OC_N * OC_N;
    extern __shared__ __align__(16) unsigned char oc_smem_raw[];
    double2 *P0 = reinterpret_cast<double2 *>(oc_smem_raw), *P1 = P0 + OC_BUF, *P2 = P1 + OC_BUF;
    __shared__ cplx coef[kMaxTerms];

    const int tid = threadIdx.x;
    double2 *Yg = scratch + (size_t)blockIdx.x * 3 * NN;
    double2 *Fg[2] = {Yg + NN, Yg + 2 * NN};
    const unsigned long long lo = nsteps * blockIdx.x / gridDim.x, hi = nsteps * (blockIdx.x + 1) / gridDim.x;
    int f_cur = 0;
    bool have_f = false;
    const int M = p.M;
    const cplx zero{0.0, 0.0};

    for (unsigned long long j = lo; j < hi; ++j) {
        for (int t = tid; t < p.nterms; t += OC_THREADS)
            coef[t] = step_coefficient<IO>(p.terms[t], carr, p.pts, p.quad, p.magfac, j);
        __syncthreads();
        // ---- assemble Y into P0 (operand of the first product); Horner also keeps a global copy for the epilogues ----
        // 16 elements per thread; the table reads of 4 elements (up to 9 matrices each) are issued back to back so that
        // their L2 latency overlaps -- this phase has no other CTA on the SM to hide behind
#pragma unroll 4
        for (int it = 0; it < NN / OC_THREADS; ++it) {
            const int e = tid + it * OC_THREADS;
            double2 x = __ldg(H + e);
            double2 h[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (t < p.nterms) h[t] = __ldg(H + (size_t)p.terms[t].mat * NN + e);
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (t < p.nterms) {
                    const cplx ct = coef[t];
                    x.x += ct.re * h[t].x - ct.im * h[t].y;
                    x.y += ct.re * h[t].y + ct.im * h[t].x;
                }
            for (int t = 8; t < p.nterms; ++t) {
                const double2 hh = __ldg(H + (size_t)p.terms[t].mat * NN + e);
                const cplx ct = coef[t];
                x.x += ct.re * hh.x - ct.im * hh.y;
                x.y += ct.re * hh.y + ct.im * hh.x;
            }
            const double yr = x.x * p.sigma, yi = x.y * p.sigma;
            const int r = e >> 6, c = e & 63;
            const bool diag = (r == c);
            const double2 s2 = make_double2(((prog.u.re * yr - prog.u.im * yi) + (diag ? prog.v_lo.re : 0.0)) + (diag ? prog.v.re : 0.0),
                                            ((prog.u.re * yi + prog.u.im * yr) + (diag ? prog.v_lo.im : 0.0)) + (diag ? prog.v.im : 0.0));
            if (p.horner) {
                P0[r * OC_P + c] = make_double2(yr, yi);
                P2[r * OC_P + c] = s2;                                  // R_L
                Yg[e] = make_double2(yr, yi);
            } else {
                P2[r * OC_P + c] = make_double2(yr, yi);                // Y stays the right operand of every product
                P0[r * OC_P + c] = s2;                                  // B_{M-1}  (or E itself when M == 1)
                P1[r * OC_P + c] = make_double2(diag ? prog.w.re : 0.0, diag ? prog.w.im : 0.0);   // B_M
            }
        }
        __syncthreads();

        double2 *E;
        if (p.horner) {
            OcEpilogue ep{};
            ep.beta1 = ep.beta1_lo = ep.gamma = ep.gamma_lo = zero;
            ep.d_smem = P1;
            oc_gemm(P0, P0, ep);                                        // W = Y Y  -> P1
            __syncthreads();
            double2 *cur = P2, *oth = P0;
            for (int i = (M >> 1) - 1; i >= 0; --i) {                   // R <- R W + c_{2i+1} Y + c_{2i} I
                OcEpilogue eh{};
                eh.c1_glob = Yg;
                eh.beta1 = p.a[2 * i + 1]; eh.beta1_lo = p.a_lo[2 * i + 1];
                eh.gamma = p.a[2 * i]; eh.gamma_lo = p.a_lo[2 * i];
                eh.d_smem = oth;
                oc_gemm(cur, P1, eh);
                __syncthreads();
                double2 *t = cur; cur = oth; oth = t;
            }
            E = cur;
        } else if (M == 1) {
            E = P0;
        } else {
            double2 *cur = P0, *oth = P1;
            for (int k = M - 2; k >= 0; --k) {                          // B_k = B_{k+1} Y - B_{k+2} + a_k I ; last: E = B_1 Y - 2 B_2 + a0' I
                OcEpilogue ec{};
                ec.c1_smem = oth;
                ec.beta1 = cplx{k == 0 ? -2.0 : -1.0, 0.0}; ec.beta1_lo = zero;
                ec.gamma = p.a[k]; ec.gamma_lo = p.a_lo[k];
                ec.d_smem = oth;
                oc_gemm(cur, P2, ec);
                __syncthreads();
                double2 *t = cur; cur = oth; oth = t;
            }
            E = cur;
        }

        // ---- running product in E-form:  F <- E + F + E F  (later step on the left); F lives in L2 ----
        if (!have_f) {
            for (int e = tid; e < NN; e += OC_THREADS) Fg[f_cur][e] = E[(e >> 6) * OC_P + (e & 63)];
            have_f = true;
        } else {
            double2 *fb = (E == P0) ? P1 : P0;                          // any buffer that is not E
#pragma unroll
            for (int it = 0; it < NN / OC_THREADS; ++it) {
                const int e = tid + it * OC_THREADS;
                fb[(e >> 6) * OC_P + (e & 63)] = Fg[f_cur][e];
            }
            __syncthreads();
            OcEpilogue ef{};
            ef.c1_smem = E; ef.beta1 = cplx{1.0, 0.0}; ef.beta1_lo = zero;
            ef.c2_glob = Fg[f_cur]; ef.beta2 = 1.0;
            ef.gamma = ef.gamma_lo = zero;
            ef.d_glob = Fg[f_cur ^ 1];
            oc_gemm(E, fb, ef);
            f_cur ^= 1;
        }
        __syncthreads();
    }
    double2 *out = partials + (size_t)blockIdx.x * NN;
    for (int e = tid; e < NN; e += OC_THREADS) out[e] = have_f ? Fg[f_cur][e] : make_double2(0.0, 0.0);
}

int k4_onchip_slots(int num_sms) {
    int per_sm = 0;
    cudaFuncSetAttribute(k4_onchip_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OC_SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4_onchip_kernel<double2>, OC_THREADS, OC_SMEM);
    return per_sm < 1 ? 0 : per_sm * num_sms;
}

cudaError_t k4_onchip(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                      double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream) {
    if (fp64_io) {
        cudaError_t e = cudaFuncSetAttribute(k4_onchip_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OC_SMEM);
        if (e != cudaSuccess) return e;
        k4_onchip_kernel<double2><<<grid, OC_THREADS, OC_SMEM, stream>>>(p, prog, (const double2 *)carr, H, scratch, partials, nsteps);
    } else {
        cudaError_t e = cudaFuncSetAttribute(k4_onchip_kernel<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OC_SMEM);
        if (e != cudaSuccess) return e;
        k4_onchip_kernel<float2><<<grid, OC_THREADS, OC_SMEM, stream>>>(p, prog, (const float2 *)carr, H, scratch, partials, nsteps);
    }
    return cudaGetLastError();
}

}  // namespace pb
