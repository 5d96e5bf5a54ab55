// k4_onchip.cu -- dim 33..64: the whole per-step pipeline of one CTA with its operands RESIDENT in shared memory.
//
// One persistent CTA (256 threads, 8 warps of 32x16 DMMA tiles) per SM walks a contiguous range of time steps.  Three
// 64x64 complex128 buffers (pitch 68, 209 KB) hold, per step, Y / W = Y^2 / the recurrence matrices; every product of the
// series reads both operands from shared memory with LDS.128 straight into mma.sync.m8n8k4.f64 fragments and writes its
// result back to shared memory in the epilogue; the Horner addend Y is thread-private and stays in registers for the whole
// step, so inside a step only the running product F touches L2 (one read, one write), and there is no cp.async pipeline
// to fill or drain between the ~7 dependent products.
// Because one CTA owns the SM, every phase without DMMAs leaves the FP64 tensor pipe idle; therefore
//   * the assembly of Y_{j+1} = sigma (H0 + sum_t c_t H_t) is software-pipelined INTO the main loop of the last product
//     of step j (the running-product update): one matrix element per thread and k-tile, its table loads issued one
//     k-tile ahead, so their L2 latency hides behind the DMMAs;
//   * the start value of the Horner recurrence is folded into the first product
//     (R_{L-1} = c_{2L+1} (Y W) + c_{2L} W + c_{2L-1} Y + c_{2L-2} I), no separate elementwise pass;
//   * epilogues are specialised at compile time.
// Replaces, for these dimensions, the MMAX batched cuBLAS GEMMs + diagonal_add launches of parament.cpp:569-652 and the
// reduction of parament.cpp:657-718; supersedes k4_chain_kernel (operands streamed from an L2 scratch) for npad == 64.
#include <cstdio>
#include <cstdint>
#include "coef.cuh"
#include "k4_gemm.hpp"

namespace pb {

// L2 evict-last policy for the per-CTA scratch (running product F, pre-assembled steps): it is rewritten every few steps and
// must stay in L2; without the hint ~25 % of those writes were evicted to HBM (2.4 GB per 2e5 steps at dim 64).
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_keep(double2 *a, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;\n" ::"l"(a), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_keep(unsigned dst, const void *src, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "l"(pol));
}

// Per-phase cycle counters (development aid, -DPB_PHASE_TIMING): block 0 prints its accumulated clock64() deltas.
#ifdef PB_PHASE_TIMING
#define PB_T_DECL long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_last_ = clock64();
#define PB_T(i) { const long long pt_now_ = clock64(); pt_[i] += pt_now_ - pt_last_; pt_last_ = pt_now_; }
#define PB_T_PRINT if (blockIdx.x == 0 && threadIdx.x == 0) printf("phase cycles: assemble %lld  sq %lld  cube/first %lld  top %lld  horner %lld  fcopy %lld  chain %lld  other %lld\n", pt_[0], pt_[1], pt_[2], pt_[3], pt_[4], pt_[5], pt_[6], pt_[7]);
#else
#define PB_T_DECL
#define PB_T(i)
#define PB_T_PRINT
#endif

// Geometry for padded dimension N (64: 8 warps of 32x16 tiles; 32: 4 warps of 16x16 tiles).
template <int N>
struct Oc {
    static constexpr int P = N;                     // pitch in double2: no padding, the 16-byte column index is XOR-swizzled (at())
    static constexpr int BUF = N * P;               // elements per buffer
    static constexpr int WM = N / 2, WN = N / 4 < 16 ? 16 : N / 4;   // warp tile
    static constexpr int WARPS_N = N / WN, WARPS = (N / WM) * WARPS_N;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int MT = WM / 8, NTL = WN / 8;
    static constexpr int NN = N * N;
    static constexpr int EPT = NN / THREADS;        // matrix elements per thread in elementwise phases
    static constexpr int COEF = 3 * BUF;             // offset of the coefficient block of the fused assembly (OC_BLOCK x kMaxTerms)
    static constexpr size_t SMEM = (3 * (size_t)BUF + 4 * kMaxTerms) * sizeof(double2);   // 196 KB + 4 KB at N = 64
    static_assert(EPT == N / 4, "one assembled element per k-tile");
    // Element (r, c) of a buffer.  The row pitch is a multiple of 128 bytes, so the bank group (16-byte unit within a 128-byte
    // line) of an element is c & 7; XOR-ing it with s(r) = 0, 5, 2, 7 for r mod 4 = 0..3 makes every access pattern of this
    // kernel conflict-free per 8-lane LDS.128 / STS.128 phase (lane = 4 g + q):
    //   A fragments  rows R + g, cols 4 kt + q       : g = 2m, 2m + 1 differ in bit 2 of s -> the two rows use different halves
    //   B fragments  rows 4 kt + q, cols C + g       : s takes four values with distinct bits (2, 1) -> four disjoint unit pairs
    //   epilogue     rows R + g, cols C + 2 q (+ 1)  : bit 0 of s separates the even units of row 2m from those of row 2m + 1
    //   row-wise elementwise passes and cp.async     : eight consecutive units of one row stay a permutation of the line
    // (the padded pitch N + 4 of round 1 left the B-fragment loads and the epilogue two-way conflicted: 34 % of all
    // shared-memory wavefronts, profiles/ncu_prof_onchip_C3_r1.txt).
    __device__ static __forceinline__ int swz(int r) { return ((r & 1) * 5) | (r & 2); }
    __device__ static __forceinline__ int at(int r, int c) { return r * N + (c ^ swz(r)); }
};
constexpr int OC_MAXT = 8;                          // control terms the fused assembly keeps in registers
constexpr int OC_BLOCK = 4;                         // time steps assembled per pass over the Hamiltonian table

enum OcEpi : int {
    EPI_STORE = 0,     // D_smem = A B
    EPI_FIRST = 1,     // D_smem = alpha (A B) + bw * Wown(smem) + by * Y(global) + gamma I          (first Horner product)
    EPI_HORNER = 2,    // D_smem = A B + i ci Y(global) + cr I                                      (+ sub-ulp remainders if LO)
    EPI_CLENSHAW = 3,  // D_smem = A B + beta * Cown(smem, == D) + gamma I
    EPI_CHAIN = 4,     // D_glob = A B + Eown(smem) + F(global)
    EPI_KEEP = 5,      // D_smem = A B, and the thread keeps its own elements of the product in registers (y2)
    EPI_PS3 = 6,       // D_smem = A B + b2 * Y2own(regs) + b1 * Yown(regs) + gamma I        (+ sub-ulp remainders if LO)
    EPI_S12_LR = 7,    // D_smem = A B + i kV Vown(smem) + kW Wown(regs) + i kY Yown(regs) + kI I, then (after a barrier: the
                       // second destination is the A operand) D_smem2 = A B + i k2V Vown + k2W Wown     (degree 12, L and R)
    EPI_S12_E = 8,     // D_smem = A B + i kV Vown(smem) + kW Wown(regs) + i kY Yown(regs) + kI I   (+ sub-ulp remainders if LO)
    EPI_ADD = 9        // D_smem = A B + Sown(smem, == D): the addend of the last series product, precomputed by EPI_S12_LR (want_s)
};

// Shared-memory buffers are named by their element OFFSET into the dynamic shared array, never by pointer: a pointer that
// travels through a struct loses its address space and the compiler falls back to generic LD/ST (measured: 233 generic
// loads in the SASS and ~20 % of the kernel time).
extern __shared__ __align__(16) double2 oc_smem[];

struct OcArgs {
    int sA, sB;                 // operands (offsets into oc_smem)
    int d_smem;                 // destination buffer (all but EPI_CHAIN)
    double2 *d_glob;            // destination in global memory (EPI_CHAIN)
    uint64_t keep;              // L2 evict-last policy for d_glob
    int c_smem;                 // shared-memory addend, own elements (EPI_FIRST: W, EPI_CLENSHAW: B_{k+2}, EPI_CHAIN: E)
    int c_smem2;                // second shared-memory addend, own elements (EPI_CHAIN: F, which is also the B operand)
    cplx alpha, bw, by;         // EPI_FIRST (by also: coefficient of Y in EPI_PS3)
    cplx by_lo, b2, b2_lo;      // EPI_PS3
    double ci, ci_lo, cr, cr_lo;   // EPI_HORNER
    double beta;                // EPI_CLENSHAW
    cplx gamma, gamma_lo;
    int v_smem, d_smem2;        // EPI_S12_*: own elements of V = Y^3; second destination
    double kV, kW, kY, kI, kV_lo, kW_lo, kY_lo, kI_lo, k2V, k2W;
    // EPI_S12_LR with want_s: after the barrier the thread also writes S = i sV Vown + sW Wown + i sY Yown + sI I (+ sub-ulp
    // remainders if LO), the addend of the NEXT product (EPI_ADD), over its own elements of s_smem -- the last use of the own
    // elements of Y and W in registers, so the products after this one run with 128 registers less.
    int want_s, s_smem;
    double sV, sW, sY, sI, sV_lo, sW_lo, sY_lo, sI_lo;
};

// Fused assembly of the next step's Y (one element per thread and k-tile).
struct OcAssemble {
    const double2 *H;           // table, [mat][64*64]
    int coef_off;               // offset into oc_smem of the coefficients: [b * kMaxTerms + t] for step j + 1 + b
    int real_only;              // every coefficient of the pass is real (warp-uniform): half the multiply-adds
    const Term *terms;
    int nterms;
    double sigma;
    int y_smem;                 // destination buffer of the NEXT step's Y (offset into oc_smem, layout Oc<N>::at)
    int nblock;                 // steps assembled per pass of the table (1..OC_BLOCK)
    double2 *y_glob;            // row-major destinations of the later steps of the block: y_glob[(b - 1) * N*N + e], b >= 1
    uint64_t keep;              // L2 evict-last policy for y_glob
};

template <int N>
__device__ __forceinline__ void oc_issue_loads(const OcAssemble &as, int e, double2 &h0, double2 (&h)[OC_MAXT]) {
    h0 = __ldg(as.H + e);
#pragma unroll
    for (int t = 0; t < OC_MAXT; ++t)
        if (t < as.nterms) h[t] = __ldg(as.H + (size_t)as.terms[t].mat * Oc<N>::NN + e);
}

// Y element of block step b from the loaded table values
__device__ __forceinline__ double2 oc_combine(const OcAssemble &as, int b, double2 x, const double2 (&h)[OC_MAXT]) {
#pragma unroll
    for (int t = 0; t < OC_MAXT; ++t)
        if (t < as.nterms) {
            const double2 ct = oc_smem[as.coef_off + b * kMaxTerms + t];   // explicit shared-memory access (LDS, broadcast)
            if (as.real_only) {
                x.x = fma(ct.x, h[t].x, x.x);
                x.y = fma(ct.x, h[t].y, x.y);
            } else {
                x.x += ct.x * h[t].x - ct.y * h[t].y;
                x.y += ct.x * h[t].y + ct.y * h[t].x;
            }
        }
    return make_double2(x.x * as.sigma, x.y * as.sigma);
}

// D = A * B (+ epilogue), A and B in shared memory (layout Oc<N>::at).  All 256 threads; no barrier inside.
// a thread's own elements of a matrix, in epilogue order (sized for the larger geometry)
typedef double2 OcOwn[4][2][2];

// own elements of a shared-memory matrix -> registers
template <int N>
__device__ __forceinline__ void oc_load_own(OcOwn &y, int m) {
    using G = Oc<N>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;
    const int wm0 = (warp / G::WARPS_N) * G::WM, wn0 = (warp % G::WARPS_N) * G::WN;
#pragma unroll
    for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < G::NTL; ++nt) {
            const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q;
            y[mt][nt][0] = oc_smem[m + G::at(r, c)];
            y[mt][nt][1] = oc_smem[m + G::at(r, c + 1)];
        }
}

// `y`: the thread's own elements of Y (EPI_FIRST / EPI_HORNER), loaded once per step by oc_load_own.
// MUL3: the complex product from THREE real products per fragment pair (P1 = Ar Br, P2 = Ai Bi, P3 = (Ar + Ai)(Br + Bi);
// Re = P1 - P2, Im = P3 - P1 - P2): 24 instead of 32 DMMAs per k-tile for six DADDs and a third accumulator set.  Used where
// the thread's own elements of Y and W are not both live in registers (k4_gemm.cu tile_gemm has the norm-wise error argument).
template <int N, int EPI, bool LO, bool ASSEMBLE, bool MUL3 = false>
__device__ __forceinline__ void oc_gemm(const OcArgs &g, const OcAssemble &as, const OcOwn &y, OcOwn &y2) {
    using G = Oc<N>;
    constexpr int OC_P = G::P, OC_N = N, OC_THREADS = G::THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, q = lane & 3;
    const int wm0 = (warp / G::WARPS_N) * G::WM, wn0 = (warp % G::WARPS_N) * G::WN;
    constexpr int MT = G::MT, NTL = G::NTL;
    const double2 *sA = oc_smem + g.sA;
    const double2 *sB = oc_smem + g.sB;

    double cre[MT][NTL][2], cim[MT][NTL][2];          // MUL3: P1 and P2 until the main loop is over
    double p3[MUL3 ? MT : 1][MUL3 ? NTL : 1][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            cre[mt][nt][0] = cre[mt][nt][1] = 0.0; cim[mt][nt][0] = cim[mt][nt][1] = 0.0;
            if (MUL3) p3[MUL3 ? mt : 0][MUL3 ? nt : 0][0] = p3[MUL3 ? mt : 0][MUL3 ? nt : 0][1] = 0.0;
        }

    // swizzled column offsets (Oc<N>::at): the rows of this lane's A fragments are congruent to g mod 8, those of its B
    // fragments to q mod 4, so the XOR masks are per-thread constants
    const int a_col[2] = {q ^ G::swz(gq), (4 + q) ^ G::swz(gq)};
    const int b_col = gq ^ G::swz(q);

    double2 h0, h[OC_MAXT];
    if (ASSEMBLE) oc_issue_loads<N>(as, tid, h0, h);

    // (Double-buffering the fragments in registers -- k-tile kt + 1 requested before the DMMAs of k-tile kt -- was measured
    // without effect at dim 64: 278.5 vs 278.7 ms per 1e6 steps; the second warp of the scheduler already covers the LDS latency.)
#pragma unroll 2
    for (int kt = 0; kt < OC_N / 4; ++kt) {
        double2 af[MT], bf[NTL];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) af[mt] = sA[(wm0 + 8 * mt + gq) * OC_P + 8 * (kt >> 1) + a_col[kt & 1]];
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) bf[nt] = sB[(4 * kt + q) * OC_P + wn0 + 8 * nt + b_col];
        if (ASSEMBLE) {
            // element e = tid + 256 kt of the next step's Y: consume the loads issued one k-tile ago, issue the next ones
            const int e = tid + kt * OC_THREADS;
            oc_smem[as.y_smem + G::at(e / OC_N, e % OC_N)] = oc_combine(as, 0, h0, h);
            for (int b = 1; b < as.nblock; ++b) st_keep(as.y_glob + (size_t)(b - 1) * (OC_N * OC_N) + e, oc_combine(as, b, h0, h), as.keep);
            if (kt + 1 < OC_N / 4) oc_issue_loads<N>(as, e + OC_THREADS, h0, h);
        }
        if (MUL3) {
            double as_[MT], bs_[NTL];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) as_[mt] = af[mt].x + af[mt].y;
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) bs_[nt] = bf[nt].x + bf[nt].y;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].x, bf[nt].x);
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].y, bf[nt].y);
                    dmma884(p3[MUL3 ? mt : 0][MUL3 ? nt : 0][0], p3[MUL3 ? mt : 0][MUL3 ? nt : 0][1], as_[mt], bs_[nt]);
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
                    dmma884(cre[mt][nt][0], cre[mt][nt][1], af[mt].y, neg(bf[nt].y));
                    dmma884(cim[mt][nt][0], cim[mt][nt][1], af[mt].y, bf[nt].x);
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

    // ---- epilogue (small terms first, dominant term last with one FMA rounding: DESIGN.md "Numerics") ----
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q;
            double vr[2] = {cre[mt][nt][0], cre[mt][nt][1]}, vi[2] = {cim[mt][nt][0], cim[mt][nt][1]};
            if (EPI == EPI_FIRST) {
                const double2 w0 = oc_smem[g.c_smem + G::at(r, c)], w1 = oc_smem[g.c_smem + G::at(r, c + 1)];
                const double2 ws[2] = {w0, w1};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double pr = vr[i], pi = vi[i];
                    double tr = g.alpha.re * pr - g.alpha.im * pi + (g.bw.re * ws[i].x - g.bw.im * ws[i].y);
                    double ti = g.alpha.re * pi + g.alpha.im * pr + (g.bw.re * ws[i].y + g.bw.im * ws[i].x);
                    if (r == c + i) { tr += g.gamma.re; ti += g.gamma.im; }
                    vr[i] = fma(g.by.re, y[mt][nt][i].x, fma(-g.by.im, y[mt][nt][i].y, tr));
                    vi[i] = fma(g.by.re, y[mt][nt][i].y, fma(g.by.im, y[mt][nt][i].x, ti));
                }
            } else if (EPI == EPI_HORNER) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double yr = y[mt][nt][i].x, yi = y[mt][nt][i].y;
                    if (LO) {
                        vr[i] = fma(-g.ci_lo, yi, vr[i]);
                        vi[i] = fma(g.ci_lo, yr, vi[i]);
                        if (r == c + i) vr[i] = (vr[i] + g.cr_lo) + g.cr;
                    } else if (r == c + i) {
                        vr[i] += g.cr;
                    }
                    vr[i] = fma(-g.ci, yi, vr[i]);
                    vi[i] = fma(g.ci, yr, vi[i]);
                }
            } else if (EPI == EPI_CLENSHAW) {
                const double2 w0 = oc_smem[g.c_smem + G::at(r, c)], w1 = oc_smem[g.c_smem + G::at(r, c + 1)];
                vr[0] = fma(g.beta, w0.x, vr[0]); vi[0] = fma(g.beta, w0.y, vi[0]);
                vr[1] = fma(g.beta, w1.x, vr[1]); vi[1] = fma(g.beta, w1.y, vi[1]);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (r == c + i) { vr[i] = (vr[i] + g.gamma_lo.re) + g.gamma.re; vi[i] = (vi[i] + g.gamma_lo.im) + g.gamma.im; }
            } else if (EPI == EPI_KEEP) {
                y2[mt][nt][0] = make_double2(vr[0], vi[0]);
                y2[mt][nt][1] = make_double2(vr[1], vi[1]);
            } else if (EPI == EPI_PS3) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double2 a1 = y[mt][nt][i], a2 = y2[mt][nt][i];
                    if (LO) {
                        vr[i] += (g.b2_lo.re * a2.x - g.b2_lo.im * a2.y) + (g.by_lo.re * a1.x - g.by_lo.im * a1.y);
                        vi[i] += (g.b2_lo.re * a2.y + g.b2_lo.im * a2.x) + (g.by_lo.re * a1.y + g.by_lo.im * a1.x);
                        if (r == c + i) { vr[i] = (vr[i] + g.gamma_lo.re) + g.gamma.re; vi[i] = (vi[i] + g.gamma_lo.im) + g.gamma.im; }
                    } else if (r == c + i) {
                        vr[i] += g.gamma.re; vi[i] += g.gamma.im;
                    }
                    vr[i] = fma(g.b2.re, a2.x, fma(-g.b2.im, a2.y, vr[i]));
                    vi[i] = fma(g.b2.re, a2.y, fma(g.b2.im, a2.x, vi[i]));
                    vr[i] = fma(g.by.re, a1.x, fma(-g.by.im, a1.y, vr[i]));      // the Y term dominates: last
                    vi[i] = fma(g.by.re, a1.y, fma(g.by.im, a1.x, vi[i]));
                }
            } else if (EPI == EPI_S12_LR || EPI == EPI_S12_E) {
                const double2 vs[2] = {oc_smem[g.v_smem + G::at(r, c)], oc_smem[g.v_smem + G::at(r, c + 1)]};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double2 a1 = y[mt][nt][i], a2 = y2[mt][nt][i];
                    if (LO) {
                        vr[i] += (g.kW_lo * a2.x - g.kV_lo * vs[i].y) - g.kY_lo * a1.y;
                        vi[i] += (g.kW_lo * a2.y + g.kV_lo * vs[i].x) + g.kY_lo * a1.x;
                        if (r == c + i) vr[i] = (vr[i] + g.kI_lo) + g.kI;
                    } else if (r == c + i) {
                        vr[i] += g.kI;
                    }
                    vr[i] = fma(-g.kV, vs[i].y, vr[i]);
                    vi[i] = fma(g.kV, vs[i].x, vi[i]);
                    vr[i] = fma(g.kW, a2.x, vr[i]);
                    vi[i] = fma(g.kW, a2.y, vi[i]);
                    vr[i] = fma(-g.kY, a1.y, vr[i]);                             // the Y term dominates: last
                    vi[i] = fma(g.kY, a1.x, vi[i]);
                }
            } else if (EPI == EPI_ADD) {
                const double2 s0 = oc_smem[g.d_smem + G::at(r, c)], s1 = oc_smem[g.d_smem + G::at(r, c + 1)];
                vr[0] += s0.x; vi[0] += s0.y;
                vr[1] += s1.x; vi[1] += s1.y;
            } else if (EPI == EPI_CHAIN) {
                const double2 e0 = oc_smem[g.c_smem + G::at(r, c)], e1 = oc_smem[g.c_smem + G::at(r, c + 1)];
                const double2 f0 = oc_smem[g.c_smem2 + G::at(r, c)], f1 = oc_smem[g.c_smem2 + G::at(r, c + 1)];
                vr[0] += e0.x + f0.x; vi[0] += e0.y + f0.y;
                vr[1] += e1.x + f1.x; vi[1] += e1.y + f1.y;
            }
            if (EPI == EPI_CHAIN) {
                st_keep(g.d_glob + r * OC_N + c, make_double2(vr[0], vi[0]), g.keep);
                st_keep(g.d_glob + r * OC_N + c + 1, make_double2(vr[1], vi[1]), g.keep);
            } else {
                oc_smem[g.d_smem + G::at(r, c)] = make_double2(vr[0], vi[0]);
                oc_smem[g.d_smem + G::at(r, c + 1)] = make_double2(vr[1], vi[1]);
            }
        }
    if (EPI == EPI_S12_LR) {
        __syncthreads();   // every warp has read its A fragments: the A buffer may now receive the second combination
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double2 v = oc_smem[g.v_smem + G::at(r, c + i)], a2 = y2[mt][nt][i];
                    oc_smem[g.d_smem2 + G::at(r, c + i)] =
                        make_double2(fma(g.k2W, a2.x, fma(-g.k2V, v.y, cre[mt][nt][i])), fma(g.k2W, a2.y, fma(g.k2V, v.x, cim[mt][nt][i])));
                    if (g.want_s) {   // the next product's addend (small terms first, the dominant Y term last)
                        const double2 a1 = y[mt][nt][i];
                        double sr = 0.0, si = 0.0;
                        if (LO) {
                            sr = (g.sW_lo * a2.x - g.sV_lo * v.y) - g.sY_lo * a1.y;
                            si = (g.sW_lo * a2.y + g.sV_lo * v.x) + g.sY_lo * a1.x;
                            if (r == c + i) sr = (sr + g.sI_lo) + g.sI;
                        } else if (r == c + i) {
                            sr = g.sI;
                        }
                        sr = fma(-g.sV, v.y, sr); si = fma(g.sV, v.x, si);
                        sr = fma(g.sW, a2.x, sr); si = fma(g.sW, a2.y, si);
                        sr = fma(-g.sY, a1.y, sr); si = fma(g.sY, a1.x, si);
                        oc_smem[g.s_smem + G::at(r, c + i)] = make_double2(sr, si);
                    }
                }
            }
    }
}

// scratch per CTA: 2 matrices (F0, F1), row-major pitch 64 (the host allocates kSeriesSlots + 2).
template <int N, typename IO>
__global__ void __launch_bounds__(Oc<N>::THREADS, N == 64 ? 1 : 2)
k4_onchip_kernel(const __grid_constant__ SeriesParams p, const __grid_constant__ SeriesProgram prog, const IO *__restrict__ carr, const double2 *__restrict__ H,
                 double2 *__restrict__ scratch, double2 *__restrict__ partials, unsigned long long nsteps) {
    using G = Oc<N>;
    constexpr int OC_P = G::P, OC_N = N, OC_THREADS = G::THREADS, OC_NN = G::NN, OC_EPT = G::EPT, OC_BUF = G::BUF;
    static_assert(OC_BLOCK == 4, "Oc<N>::SMEM reserves 4 x kMaxTerms coefficients");
    cplx *coef = reinterpret_cast<cplx *>(oc_smem + G::COEF);   // writes only; the fused pass reads it by offset

    const int tid = threadIdx.x;
    // scratch per CTA: F0, F1 and OC_BLOCK - 1 pre-assembled Y matrices, packed (the host provides kSeriesSlots + 2 >= 5 per CTA)
    double2 *cta_scratch = scratch + (size_t)blockIdx.x * (2 + OC_BLOCK - 1) * OC_NN;
    double2 *Fg[2] = {cta_scratch, cta_scratch + OC_NN};
    double2 *Yq = cta_scratch + 2 * OC_NN;
    int ahead = 0, yq_slot = 0;     // pre-assembled future steps waiting in Yq
    const unsigned long long lo = nsteps * blockIdx.x / gridDim.x, hi = nsteps * (blockIdx.x + 1) / gridDim.x;
    int f_cur = 0;
    bool have_f = false;
    const int M = p.M;
    const bool horner = p.horner != 0;
    const bool fuse = p.nterms <= OC_MAXT;   // next step's assembly rides in the running-product update
    const bool ps3 = (M == 8 || M >= 10);    // degrees at which blocks of three need fewer products than the Y^2 form
    constexpr bool LO = sizeof(IO) == sizeof(double2);   // sub-ulp remainders of the constants: complex128 contexts only

    const uint64_t keep = l2_keep_policy();
    OcAssemble as{};
    as.H = H; as.coef_off = G::COEF; as.terms = p.terms; as.nterms = p.nterms; as.sigma = p.sigma; as.keep = keep;
    const OcAssemble none{};

    int iy = 0;              // buffer index holding Y of the current step
    bool y_ready = false;    // Y of the current step was assembled by the previous step's fused pass

    PB_T_DECL
    for (unsigned long long j = lo; j < hi; ++j) {
        PB_T(7)
        if (!y_ready) {
            // ---- standalone assembly (first step of the CTA, or more control terms than the fused pass keeps) ----
            for (int t = tid; t < p.nterms; t += OC_THREADS)
                coef[t] = step_coefficient<IO>(p.terms[t], carr, p.pts, p.quad, p.magfac, j);
            __syncthreads();
#pragma unroll 4
            for (int it = 0; it < OC_EPT; ++it) {
                const int e = tid + it * OC_THREADS;
                double2 x = __ldg(H + e);
                for (int t = 0; t < p.nterms; ++t) {
                    const double2 hh = __ldg(H + (size_t)p.terms[t].mat * OC_NN + e);
                    const cplx ct = coef[t];
                    x.x += ct.re * hh.x - ct.im * hh.y;
                    x.y += ct.re * hh.y + ct.im * hh.x;
                }
                const double2 v = make_double2(x.x * p.sigma, x.y * p.sigma);
                oc_smem[iy * OC_BUF + G::at(e / OC_N, e % OC_N)] = v;
            }
            __syncthreads();
        }
        PB_T(0)
        const int PY = iy * OC_BUF, PA = ((iy + 1) % 3) * OC_BUF, PB = ((iy + 2) % 3) * OC_BUF;   // buffer offsets
        int E;       // result of the series
        int Fb;      // buffer that will receive F for the running-product update
        int Yn;      // buffer that will receive the next step's Y

        OcOwn y, y2;
        if (p.horner == 3) {
            // ---- degree 8 in three products (api.cu solve_degree8; p.a[k].re = c4 c3 d2 d1 e2 e0 r2' r1 r0, A = -i Y):
            //      W = Y Y, T = c4 W + i c3 Y, y02 = T W, L = y02 - d2 W - i d1 Y + e0 I, R = y02 - e2 W,
            //      E = L R - r2' W - i r1 Y + r0 I.  Same epilogues as the degree-12 form with the V terms switched off ----
            oc_load_own<N>(y, PY);
            OcArgs a{};
            a.sA = PY; a.sB = PY; a.d_smem = PA;
            oc_gemm<N, EPI_KEEP, false, false, true>(a, none, y, y2);   // W -> PA, own elements -> y2
            PB_T(1)
            {   // T -> PB (free)
                const double c4 = p.a[0].re, c3 = p.a[1].re;
#pragma unroll
                for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < G::NTL; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
                            const int r = (warp / G::WARPS_N) * G::WM + 8 * mt + gq, c = (warp % G::WARPS_N) * G::WN + 8 * nt + 2 * q + i;
                            const double2 a1 = y[mt][nt][i], a2 = y2[mt][nt][i];
                            oc_smem[PB + G::at(r, c)] = make_double2(fma(-c3, a1.y, c4 * a2.x), fma(c3, a1.x, c4 * a2.y));
                        }
            }
            __syncthreads();
            PB_T(3)
            OcArgs lr{};
            lr.sA = PB; lr.sB = PA; lr.d_smem = PY; lr.d_smem2 = PB; lr.v_smem = PA;
            lr.kV = 0.0; lr.kW = -p.a[2].re; lr.kY = -p.a[3].re; lr.kI = p.a[5].re;
            lr.k2V = 0.0; lr.k2W = -p.a[4].re;
            // ... and the addend S = -r2' W - i r1 Y + r0 I of the last product over the own elements of W (dead as an operand)
            lr.want_s = 1; lr.s_smem = PA;
            lr.sV = 0.0; lr.sW = -p.a[6].re; lr.sY = -p.a[7].re; lr.sI = p.a[8].re;
            lr.sV_lo = 0.0; lr.sW_lo = -p.a_lo[6].re; lr.sY_lo = -p.a_lo[7].re; lr.sI_lo = p.a_lo[8].re;
            if (LO) oc_gemm<N, EPI_S12_LR, true, false>(lr, none, y, y2);   // L -> PY (Y is dead as an operand), R -> PB, S -> PA
            else    oc_gemm<N, EPI_S12_LR, false, false>(lr, none, y, y2);
            __syncthreads();
            OcArgs ee{};
            ee.sA = PY; ee.sB = PB; ee.d_smem = PA;                      // E = L R + S, in place over S
            oc_gemm<N, EPI_ADD, false, false, true>(ee, none, y, y2);
            __syncthreads();
            PB_T(4)
            E = PA; Fb = PY; Yn = PB;         // L and R are dead
        } else if (p.horner == 4) {
            // ---- degree 12 in four products (api.cu solve_degree12; p.a[k].re = tV tW tY lV lW lY lI rV rW sV sW sY sI):
            //      W = Y Y, V = W Y, T' = tV V + i tW W + tY Y, y0 = T' V,
            //      L = y0 + i lV V + lW W + i lY Y + lI I, R = y0 + i rV V + rW W, E = L R + i sV V + sW W + i sY Y + sI I.
            //      Y and W enter the combinations through the thread's own elements in registers, V through its own elements
            //      in shared memory ----
            oc_load_own<N>(y, PY);
            OcArgs a{};
            a.sA = PY; a.sB = PY; a.d_smem = PA;
            oc_gemm<N, EPI_KEEP, false, false, true>(a, none, y, y2);   // W -> PA, own elements -> y2
            __syncthreads();
            PB_T(1)
            OcArgs b{};
            b.sA = PA; b.sB = PY; b.d_smem = PB;
            oc_gemm<N, EPI_STORE, false, false>(b, none, y, y2);        // V -> PB   (Y and W own elements live: four real products)
            __syncthreads();
            PB_T(2)
            {   // T' -> PA (W is dead as an operand)
                const double tV = p.a[0].re, tW = p.a[1].re, tY = p.a[2].re;
                const int lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
                const int wm0 = (warp / G::WARPS_N) * G::WM, wn0 = (warp % G::WARPS_N) * G::WN;
#pragma unroll
                for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < G::NTL; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q + i;
                            const double2 v = oc_smem[PB + G::at(r, c)], a1 = y[mt][nt][i], a2 = y2[mt][nt][i];
                            oc_smem[PA + G::at(r, c)] = make_double2(fma(tY, a1.x, fma(-tW, a2.y, tV * v.x)), fma(tY, a1.y, fma(tW, a2.x, tV * v.y)));
                        }
            }
            __syncthreads();
            PB_T(3)
            OcArgs lr{};
            lr.sA = PA; lr.sB = PB; lr.d_smem = PY; lr.d_smem2 = PA; lr.v_smem = PB;
            lr.kV = p.a[3].re; lr.kW = p.a[4].re; lr.kY = p.a[5].re; lr.kI = p.a[6].re;
            lr.k2V = p.a[7].re; lr.k2W = p.a[8].re;
            // ... and the addend S = i sV V + sW W + i sY Y + sI I of the last product over the own elements of V (dead as an operand)
            lr.want_s = 1; lr.s_smem = PB;
            lr.sV = p.a[9].re; lr.sW = p.a[10].re; lr.sY = p.a[11].re; lr.sI = p.a[12].re;
            lr.sV_lo = p.a_lo[9].re; lr.sW_lo = p.a_lo[10].re; lr.sY_lo = p.a_lo[11].re; lr.sI_lo = p.a_lo[12].re;
            if (LO) oc_gemm<N, EPI_S12_LR, true, false>(lr, none, y, y2);   // L -> PY (Y is dead as an operand), R -> PA, S -> PB
            else    oc_gemm<N, EPI_S12_LR, false, false>(lr, none, y, y2);
            __syncthreads();
            OcArgs ee{};
            ee.sA = PY; ee.sB = PA; ee.d_smem = PB;                      // E = L R + S, in place over S
            oc_gemm<N, EPI_ADD, false, false, true>(ee, none, y, y2);
            __syncthreads();
            PB_T(4)
            E = PB; Fb = PY; Yn = PA;         // L and R are dead
        } else if (horner && ps3) {
            // ---- Paterson-Stockmeyer blocks of three:  E = sum_i (c_{3i} I + c_{3i+1} Y + c_{3i+2} Y^2) V^i,  V = Y^3:
            //      2 + floor(M/3) products; Y and Y^2 enter only through the thread's own elements, kept in registers ----
            const cplx zero{0.0, 0.0};
            auto cf = [&](int m) { return m <= M ? p.a[m] : zero; };
            auto cf_lo = [&](int m) { return m <= M ? p.a_lo[m] : zero; };
            oc_load_own<N>(y, PY);
            OcArgs a{};
            a.sA = PY; a.sB = PY; a.d_smem = PA;
            oc_gemm<N, EPI_KEEP, false, false>(a, none, y, y2);         // Y^2 -> PA, own elements -> y2
            __syncthreads();
            PB_T(1)
            OcArgs b{};
            b.sA = PA; b.sB = PY; b.d_smem = PB;
            oc_gemm<N, EPI_STORE, false, false>(b, none, y, y2);        // V = Y^2 Y -> PB
            // top block R_L = c_{3L+2} Y^2 + c_{3L+1} Y + c_{3L} I from the own elements -> PY (Y is dead as an operand
            // once every warp has left the product above)
            const int L = M / 3;
            __syncthreads();
            PB_T(2)
            {
                const cplx t2 = cf(3 * L + 2), t1 = cf(3 * L + 1), t0 = cf(3 * L);
                const int lane = tid & 31, warp = tid >> 5, gq = lane >> 2, q = lane & 3;
                const int wm0 = (warp / G::WARPS_N) * G::WM, wn0 = (warp % G::WARPS_N) * G::WN;
#pragma unroll
                for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < G::NTL; ++nt)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int r = wm0 + 8 * mt + gq, c = wn0 + 8 * nt + 2 * q + i;
                            const double2 a1 = y[mt][nt][i], a2 = y2[mt][nt][i];
                            double vr = t2.re * a2.x - t2.im * a2.y + (t1.re * a1.x - t1.im * a1.y);
                            double vi = t2.re * a2.y + t2.im * a2.x + (t1.re * a1.y + t1.im * a1.x);
                            if (r == c) { vr += t0.re; vi += t0.im; }
                            oc_smem[PY + G::at(r, c)] = make_double2(vr, vi);
                        }
            }
            __syncthreads();
            PB_T(3)
            int cur = PY, oth = PA;           // Y^2 is dead as an operand as well
            for (int i = L - 1; i >= 0; --i) {
                OcArgs hA{};
                hA.sA = cur; hA.sB = PB; hA.d_smem = oth;
                hA.b2 = cf(3 * i + 2); hA.b2_lo = cf_lo(3 * i + 2); hA.by = cf(3 * i + 1); hA.by_lo = cf_lo(3 * i + 1);
                hA.gamma = cf(3 * i); hA.gamma_lo = cf_lo(3 * i);
                if (LO && i == 0) oc_gemm<N, EPI_PS3, true, false>(hA, none, y, y2);
                else              oc_gemm<N, EPI_PS3, false, false>(hA, none, y, y2);
                __syncthreads();
                const int t = cur; cur = oth; oth = t;
            }
            PB_T(4)
            E = cur; Fb = oth; Yn = PB;       // V is dead
        } else if (horner) {
            oc_load_own<N>(y, PY);             // own elements of Y: addend of every Horner epilogue of this step
            // W = Y Y -> PA
            OcArgs a{};
            a.sA = PY; a.sB = PY; a.d_smem = PA;
            oc_gemm<N, EPI_STORE, false, false>(a, none, y, y2);
            __syncthreads();
            // R_{L-1} = c_{2L+1} (Y W) + c_{2L} W + c_{2L-1} Y + c_{2L-2} I -> PB          (L = M / 2 >= 1)
            const int L = M >> 1;
            OcArgs f{};
            f.sA = PY; f.sB = PA; f.d_smem = PB; f.c_smem = PA;
            f.alpha = (2 * L + 1 <= M) ? p.a[2 * L + 1] : cplx{0.0, 0.0};
            f.bw = p.a[2 * L]; f.by = p.a[2 * L - 1]; f.gamma = p.a[2 * L - 2];
            oc_gemm<N, EPI_FIRST, false, false>(f, none, y, y2);
            __syncthreads();
            int cur = PB, oth = PY;           // Y is dead as an operand from here on (its own elements are in registers)
            for (int i = L - 2; i >= 0; --i) {
                OcArgs hA{};
                hA.sA = cur; hA.sB = PA; hA.d_smem = oth;
                hA.ci = p.a[2 * i + 1].im; hA.ci_lo = p.a_lo[2 * i + 1].im; hA.cr = p.a[2 * i].re; hA.cr_lo = p.a_lo[2 * i].re;
                if (LO && i <= 1) oc_gemm<N, EPI_HORNER, true, false>(hA, none, y, y2);
                else              oc_gemm<N, EPI_HORNER, false, false>(hA, none, y, y2);
                __syncthreads();
                const int t = cur; cur = oth; oth = t;
            }
            E = cur; Fb = oth; Yn = PA;       // W is dead
        } else {
            // Clenshaw: PY keeps Y as the right operand; PA = B_{M-1} = a_M Y + a_{M-1} I, PB = B_M = a_M I
            for (int it = 0; it < OC_EPT; ++it) {
                const int e = tid + it * OC_THREADS;
                const int r = e / OC_N, c = e % OC_N;
                const bool diag = (r == c);
                const double2 v = oc_smem[PY + G::at(r, c)];
                oc_smem[PA + G::at(r, c)] = make_double2(((prog.u.re * v.x - prog.u.im * v.y) + (diag ? prog.v_lo.re : 0.0)) + (diag ? prog.v.re : 0.0),
                                                ((prog.u.re * v.y + prog.u.im * v.x) + (diag ? prog.v_lo.im : 0.0)) + (diag ? prog.v.im : 0.0));
                oc_smem[PB + G::at(r, c)] = make_double2(diag ? prog.w.re : 0.0, diag ? prog.w.im : 0.0);
            }
            __syncthreads();
            int cur = PA, oth = PB;
            for (int k = M - 2; k >= 0; --k) {       // B_k = B_{k+1} Y - B_{k+2} + a_k I ; k = 0: E = B_1 Y - 2 B_2 + a0' I
                OcArgs cA{};
                cA.sA = cur; cA.sB = PY; cA.d_smem = oth; cA.c_smem = oth;
                cA.beta = (k == 0) ? -2.0 : -1.0; cA.gamma = p.a[k]; cA.gamma_lo = p.a_lo[k];
                oc_gemm<N, EPI_CLENSHAW, false, false>(cA, none, y, y2);
                __syncthreads();
                const int t = cur; cur = oth; oth = t;
            }
            E = cur; Fb = oth; Yn = PY;      // for M == 1 the loop is empty and E = PA = a_1 Y + a0' I
        }

        // ---- running product in E-form:  F <- E + F + E F  (later step on the left); F lives in L2 ----
        const bool more = (j + 1 < hi);
        // The next step's Y: either waiting in Yq (pre-assembled by an earlier pass: fetched by cp.async under the product
        // below) or assembled now, together with up to OC_BLOCK - 1 further steps, by one pass over the table fused into
        // the product below -- the table (320 KB at dim 64 with 4 controls) is the dominant L2 traffic of this kernel.
        const bool fetch_now = more && have_f && ahead > 0;
        const bool fuse_now = more && fuse && have_f && ahead == 0;
        int nblock = 0, any_imag = 0;
        if (fuse_now) {
            nblock = (int)min((unsigned long long)OC_BLOCK, hi - (j + 1));
            for (int i = tid; i < p.nterms * nblock; i += OC_THREADS) {
                const int b = i / p.nterms, t = i % p.nterms;
                const cplx ct = step_coefficient<IO>(p.terms[t], carr, p.pts, p.quad, p.magfac, j + 1 + b);
                coef[b * kMaxTerms + t] = ct;
                any_imag |= (ct.im != 0.0);
            }
        }
        if (!have_f) {
#pragma unroll
            for (int it = 0; it < OC_EPT; ++it) {
                const int e = tid + it * OC_THREADS;
                st_keep(Fg[f_cur] + e, oc_smem[E + G::at(e / OC_N, e % OC_N)], keep);
            }
            have_f = true;
            __syncthreads();
            y_ready = false;
            iy = Yn / OC_BUF;
        } else {
            {   // F (L2) -> shared memory without staging through registers: 16 x 16-byte cp.async per thread in flight
                const double2 *fsrc = Fg[f_cur];
#pragma unroll
                for (int it = 0; it < OC_EPT; ++it) {
                    const int e = tid + it * OC_THREADS;
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(oc_smem + Fb + G::at(e / OC_N, e % OC_N));
                    cp_async16_keep(dst, fsrc + e, keep);
                }
                asm volatile("cp.async.commit_group;\n" ::);
                asm volatile("cp.async.wait_group 0;\n" ::);
            }
            as.real_only = !__syncthreads_or(any_imag);
            PB_T(5)
            OcArgs ch{};
            ch.sA = E; ch.sB = Fb; ch.c_smem = E; ch.c_smem2 = Fb; ch.d_glob = Fg[f_cur ^ 1]; ch.keep = keep;
            if (fuse_now) {
                as.y_smem = Yn; as.nblock = nblock; as.y_glob = Yq;
                oc_gemm<N, EPI_CHAIN, false, true, true>(ch, as, y, y2);
                ahead = nblock - 1; yq_slot = 0;
            } else {
                if (fetch_now) {   // Yq[yq_slot] -> Yn, in flight during the product
                    const double2 *ysrc = Yq + (size_t)yq_slot * OC_NN;
#pragma unroll
                    for (int it = 0; it < OC_EPT; ++it) {
                        const int e = tid + it * OC_THREADS;
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(oc_smem + Yn + G::at(e / OC_N, e % OC_N));
                        cp_async16_keep(dst, ysrc + e, keep);
                    }
                    asm volatile("cp.async.commit_group;\n" ::);
                    --ahead; ++yq_slot;
                }
                oc_gemm<N, EPI_CHAIN, false, false, true>(ch, none, y, y2);
                if (fetch_now) asm volatile("cp.async.wait_group 0;\n" ::);
            }
            f_cur ^= 1;
            __syncthreads();
            PB_T(6)
            y_ready = fuse_now || fetch_now;
            iy = Yn / OC_BUF;
        }
    }
    PB_T_PRINT
    double2 *out = partials + (size_t)blockIdx.x * OC_NN;
    for (int e = tid; e < OC_NN; e += OC_THREADS) out[e] = have_f ? Fg[f_cur][e] : make_double2(0.0, 0.0);
}

template <int N>
static int onchip_slots_t(int num_sms) {
    int per_sm = 0;
    cudaFuncSetAttribute(k4_onchip_kernel<N, double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Oc<N>::SMEM);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4_onchip_kernel<N, double2>, Oc<N>::THREADS, Oc<N>::SMEM);
    return per_sm < 1 ? 0 : per_sm * num_sms;
}

int k4_onchip_slots(int npad, int num_sms) { return npad == 64 ? onchip_slots_t<64>(num_sms) : onchip_slots_t<32>(num_sms); }

template <int N, typename IO>
static cudaError_t launch_onchip_t(const SeriesParams &p, const SeriesProgram &prog, const IO *carr, const double2 *H,
                                   double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k4_onchip_kernel<N, IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Oc<N>::SMEM);
    if (e != cudaSuccess) return e;
    k4_onchip_kernel<N, IO><<<grid, Oc<N>::THREADS, Oc<N>::SMEM, stream>>>(p, prog, carr, H, scratch, partials, nsteps);
    return cudaGetLastError();
}

cudaError_t k4_onchip(bool fp64_io, const SeriesParams &p, const SeriesProgram &prog, const void *carr, const double2 *H,
                      double2 *scratch, double2 *partials, unsigned long long nsteps, int grid, cudaStream_t stream) {
    if (p.npad == 64)
        return fp64_io ? launch_onchip_t<64, double2>(p, prog, (const double2 *)carr, H, scratch, partials, nsteps, grid, stream)
                       : launch_onchip_t<64, float2>(p, prog, (const float2 *)carr, H, scratch, partials, nsteps, grid, stream);
    return fp64_io ? launch_onchip_t<32, double2>(p, prog, (const double2 *)carr, H, scratch, partials, nsteps, grid, stream)
                   : launch_onchip_t<32, float2>(p, prog, (const float2 *)carr, H, scratch, partials, nsteps, grid, stream);
}

}  // namespace pb
