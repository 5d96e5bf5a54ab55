// k1_common.cuh -- device helpers shared by the register-resident chain kernels (k1_warp.cu: FP64 on the DMMA pipe, k1_tf32.cu:
// FP32 on the TF32 tensor path): in-CTA ordered products (kernel 2), the ordered reduction of stored partials, and the
// FUSED final stage -- the last CTA to finish a group of partials reduces it, the last group reducer of a pulse reduces the
// group products and writes the propagator, so a whole equiprop is ONE kernel launch (round 1: chain + one or two reduce launches).
#pragma once
#include "frag.cuh"
#include "k1_warp.hpp"

namespace pb {


// Ordered in-CTA product: afterwards warp 0 holds Q_0 Q_1 ... Q_{nwarps-1}.  smem: (nwarps/2) matrices.
// With k < nwarps (k a power of two dividing nwarps) the product runs inside every aligned group of k warps instead, and the
// first warp of each group holds its group's product.
template <int NT>
__device__ __forceinline__ void cta_ordered_product(AccFrag<NT> &Q, double2 *smem, int warp, int nwarps, int lane, int k = 0) {
    constexpr int NP = 8 * NT;
    if (k <= 0 || k > nwarps) k = nwarps;
    for (int stride = 1; stride < k; stride <<= 1) {
        const int mask = 2 * stride - 1;
        const int slot = warp / (2 * stride);
        if ((warp & mask) == stride) store_acc<NT>(Q, smem + slot * NP * NP, NP, lane);
        __syncthreads();
        if ((warp & mask) == 0 && (warp % k) + stride < k && warp + stride < nwarps) {
            BFrag<NT> B;
            load_bfrag<NT>(B, smem + slot * NP * NP, NP, lane);
            AccFrag<NT> R;
            set_zero<NT>(R);
            cmma<NT>(R, Q, B);
            Q = R;
        }
        __syncthreads();
    }
}

// out (n x n row-major, IO precision) = Q^T: the running products are kept transposed (frag.cuh).
// Packed small systems (k1_warp.cu, p.pack = 4 or 2): Q = diag(Q_0 .. Q_{pack-1}) with nb = 8 / pack rows per block, block b holding
// the (transposed) product of the b-th part of the warp's step range.  Returns diag(Q_0 Q_1 ... Q_{pack-1}, I): the warp's
// product in the layout every later stage expects.  log2(pack) products  Q <- Q * shift(Q)  where shift moves block b + s to
// block b (a lane permutation of the accumulator layout: row g + s nb, column pair q + s nb / 2, same register).
template <int NT>
__device__ __forceinline__ void unpack_blocks(AccFrag<NT> &Q, int pack, int lane) {
    static_assert(NT == 1, "block packing is a dim <= 4 (one 8 x 8 tile) feature");
    const int g = lane >> 2, q = lane & 3, nb = 8 / pack;
    for (int sh = nb; sh < 8; sh *= 2) {
        const int src = 4 * ((g + sh) & 7) + ((q + (sh >> 1)) & 3);
        AccFrag<NT> S;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            S.re[0][0][i] = __shfl_sync(0xffffffffu, Q.re[0][0][i], src);
            S.im[0][0][i] = __shfl_sync(0xffffffffu, Q.im[0][0][i], src);
        }
        BFrag<NT> Sb;
        acc_to_bfrag<NT>(Sb, S, lane);
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) Sb.nim[kt][0] = neg(Sb.im[kt][0]);
        AccFrag<NT> R;
        set_zero<NT>(R);
        cmma<NT>(R, Q, Sb);
        Q = R;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int col = 2 * q + i;
        if (g >= nb || col >= nb) {
            Q.re[0][0][i] = (g == col) ? 1.0 : 0.0;
            Q.im[0][0][i] = 0.0;
        }
    }
}

template <int NT, typename IO>
__device__ __forceinline__ void store_propagator(const AccFrag<NT> &Q, IO *__restrict__ o, int n, int lane) {
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = acc_row(lane, mt), cidx = acc_col(lane, nt, i);
                if (r < n && cidx < n) {
                    IO v;
                    v.x = Q.re[mt][nt][i];
                    v.y = Q.im[mt][nt][i];
                    o[(size_t)cidx * n + r] = v;
                }
            }
}

// Fragments of P^T from a propagator P stored n x n row-major in the IO precision (identity in the padding).
template <int NT, typename IO>
__device__ __forceinline__ void load_acc_of_transpose(AccFrag<NT> &Q, const IO *__restrict__ P, int n, int lane) {
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = acc_row(lane, mt), c = acc_col(lane, nt, i);
                double re = (r == c) ? 1.0 : 0.0, im = 0.0;
                if (r < n && c < n) { const IO v = P[(size_t)c * n + r]; re = v.x; im = v.y; }
                Q.re[mt][nt][i] = re;
                Q.im[mt][nt][i] = im;
            }
}
template <int NT, typename IO>
__device__ __forceinline__ void load_bfrag_of_transpose(BFrag<NT> &B, const IO *__restrict__ P, int n, int lane) {
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int r = bf_row(lane, kt), c = bf_col(lane, nt);
            double re = (r == c) ? 1.0 : 0.0, im = 0.0;
            if (r < n && c < n) { const IO v = P[(size_t)c * n + r]; re = v.x; im = v.y; }
            B.re[kt][nt] = re;
            B.im[kt][nt] = im;
            B.nim[kt][nt] = neg(im);
        }
}


// Coherent loads (L2) of matrices other CTAs stored earlier in the SAME launch.
template <int NT>
__device__ __forceinline__ void load_acc_cg(AccFrag<NT> &Q, const double2 *m, int lane) {
    constexpr int NP = 8 * NT;
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double2 v = __ldcg(m + acc_row(lane, mt) * NP + acc_col(lane, nt, i));
                Q.re[mt][nt][i] = v.x;
                Q.im[mt][nt][i] = v.y;
            }
}
template <int NT>
__device__ __forceinline__ void load_bfrag_cg(BFrag<NT> &B, const double2 *m, int lane) {
    constexpr int NP = 8 * NT;
#pragma unroll
    for (int kt = 0; kt < 2 * NT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const double2 v = __ldcg(m + bf_row(lane, kt) * NP + bf_col(lane, nt));
            B.re[kt][nt] = v.x;
            B.im[kt][nt] = v.y;
            B.nim[kt][nt] = neg(v.y);
        }
}

// Ordered product P[0] P[1] ... P[cnt-1] of stored matrices by all warps of the CTA (each a contiguous sub-range with the next
// operand prefetched, then the in-CTA tree); warp 0 returns with the product.  smem: (nwarps / 2) matrices.
template <int NT>
__device__ __forceinline__ void cta_reduce_range(AccFrag<NT> &Q, const double2 *P, unsigned int cnt, double2 *smem, int warp, int nwarps, int lane) {
    constexpr int NP = 8 * NT;
    const unsigned int b0 = (unsigned int)((unsigned long long)cnt * warp / nwarps);
    const unsigned int b1 = (unsigned int)((unsigned long long)cnt * (warp + 1) / nwarps);
    if (b0 < b1) {
        load_acc_cg<NT>(Q, P + (size_t)b0 * NP * NP, lane);
        BFrag<NT> B;
        if (b0 + 1 < b1) load_bfrag_cg<NT>(B, P + (size_t)(b0 + 1) * NP * NP, lane);
        for (unsigned int b = b0 + 1; b < b1; ++b) {
            BFrag<NT> Bn;
            if (b + 1 < b1) load_bfrag_cg<NT>(Bn, P + (size_t)(b + 1) * NP * NP, lane);   // prefetch
            AccFrag<NT> R;
            set_zero<NT>(R);
            cmma<NT>(R, Q, B);
            Q = R;
            if (b + 1 < b1) B = Bn;
        }
    } else {
        set_identity<NT>(Q, lane);
    }
    cta_ordered_product<NT>(Q, smem, warp, nwarps, lane);
}

// Fused final stage of the few-long-pulses mode.  On entry warp 0 of every CTA holds the CTA's ordered product Q of its warps'
// chunks; `cta` is the CTA's index within its pulse, `nb` the CTAs per pulse.  Every CTA stores its partial and takes a ticket
// of its group of K1_GROUP partials; the last one reduces the group, takes a ticket of the pulse, and the last group reducer
// reduces the group products and writes the propagator.  Counters are left at zero for the next launch.
template <int NT, typename IO>
__device__ __forceinline__ void k1_fused_final(AccFrag<NT> &Q, double2 *partials, const K1Final &fz, unsigned int pulse, unsigned int cta,
                                               unsigned int nb, double2 *smem, int warp, int nwarps, int lane) {
    constexpr int NP = 8 * NT;
    __shared__ unsigned int s_ticket;
    const unsigned int grp = cta / K1_GROUP, gcount = min(K1_GROUP, nb - grp * K1_GROUP);
    unsigned int *cnt = fz.counters + (size_t)pulse * (fz.groups + 1);
    double2 *mine = partials + ((size_t)pulse * nb + cta) * NP * NP;
    if (nb > 1) {
        if (warp == 0) { store_acc<NT>(Q, mine, NP, lane); __threadfence(); }
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(cnt + grp, 1u);
        __syncthreads();
        if (s_ticket != gcount - 1) return;             // CTA-uniform
        __threadfence();
        cta_reduce_range<NT>(Q, partials + ((size_t)pulse * nb + (size_t)grp * K1_GROUP) * NP * NP, gcount, smem, warp, nwarps, lane);
        if (threadIdx.x == 0) cnt[grp] = 0;
        if (fz.groups > 1) {
            if (warp == 0) { store_acc<NT>(Q, fz.mid + ((size_t)pulse * fz.groups + grp) * NP * NP, NP, lane); __threadfence(); }
            __syncthreads();
            if (threadIdx.x == 0) s_ticket = atomicAdd(cnt + fz.groups, 1u);
            __syncthreads();
            if (s_ticket != fz.groups - 1) return;
            __threadfence();
            cta_reduce_range<NT>(Q, fz.mid + (size_t)pulse * fz.groups * NP * NP, fz.groups, smem, warp, nwarps, lane);
            if (threadIdx.x == 0) cnt[fz.groups] = 0;
        }
    }
    if (warp == 0) store_propagator<NT, IO>(Q, (IO *)fz.out + (size_t)pulse * fz.n * fz.n, fz.n, lane);
}

// What a chain kernel does with its warps' running products Q (all warps of the CTA call this, active or not).
//   fz.out != nullptr  fused final stage: the launch writes the propagators itself
//       reduce_in_cta      the warps of a CTA belong to one pulse: in-CTA product, then k1_fused_final across the pulse's CTAs
//       otherwise          ensemble: the chunks_per_pulse (1, 2 or 4) warps of a pulse sit next to each other in this CTA
//   fz.out == nullptr  partial products only (one per CTA, or one per warp), a k3_reduce launch follows
template <int NT, typename IO>
__device__ __forceinline__ void k1_tail(AccFrag<NT> &Q, bool active, unsigned int pulse, unsigned int chunk, unsigned int chunks_per_pulse,
                                        int reduce_in_cta, double2 *partials, const K1Final &fz, double2 *smem, int warp, int lane) {
    constexpr int NP = 8 * NT;
    if (fz.out != nullptr) {
        if (reduce_in_cta) {
            cta_ordered_product<NT>(Q, smem, warp, K1_WARPS, lane);
            if (active)   // CTA-uniform
                k1_fused_final<NT, IO>(Q, partials, fz, pulse, chunk / K1_WARPS, chunks_per_pulse / K1_WARPS, smem, warp, K1_WARPS, lane);
        } else {
            cta_ordered_product<NT>(Q, smem, warp, K1_WARPS, lane, (int)chunks_per_pulse);
            if (active && chunk == 0) store_propagator<NT, IO>(Q, (IO *)fz.out + (size_t)pulse * fz.n * fz.n, fz.n, lane);
        }
        return;
    }
    if (reduce_in_cta) {
        cta_ordered_product<NT>(Q, smem, warp, K1_WARPS, lane);
        if (warp == 0 && active) {
            const unsigned int nb = chunks_per_pulse / K1_WARPS;
            store_acc<NT>(Q, partials + ((size_t)pulse * nb + chunk / K1_WARPS) * NP * NP, NP, lane);
        }
    } else if (active) {
        store_acc<NT>(Q, partials + ((size_t)pulse * chunks_per_pulse + chunk) * NP * NP, NP, lane);
    }
}

}  // namespace pb
