// frag_tf32.cuh -- complex FP32 matrix fragments for 16 x 16 matrices on the warp-level TF32 tensor path
// (mma.sync.m16n8k8.tf32 -> HMMA.1688.F32.TF32), element-compatible with the FP64 fragments of frag.cuh:
//   FAcc2  accumulator / left-operand layout, element (mt, nt, i) <-> row 8 mt + g, column 8 nt + 2 q + i   (== AccFrag<2>)
//          The m16n8 accumulator of column tile nt is (c0, c1, c2, c3) = ((0, nt, 0), (0, nt, 1), (1, nt, 0), (1, nt, 1)), and read
//          column-slot-wise it is the A operand of k-tile kt = nt for the permuted contraction order slot q <-> column 8 kt + 2 q,
//          slot q + 4 <-> column 8 kt + 2 q + 1:  (a0, a1, a2, a3) = ((0, kt, 0), (1, kt, 0), (0, kt, 1), (1, kt, 1)).
//   FB2    right-operand layout for that order, element (kt2, nt) <-> row 8 (kt2 >> 1) + 2 q + (kt2 & 1), column 8 nt + g
//          (== BFrag<2>): the B operand of k-tile kt and column tile nt is (b0, b1) = ((2 kt, nt), (2 kt + 1, nt)).
// So a matrix moves between the FP64 and the FP32 kernels' layouts by converting registers, nothing else.
// Products are 3xTF32 splits with zero-based accumulators (see k1_tf32.cu for why).
#pragma once
#include "frag.cuh"

namespace pb {

struct FAcc2 { float re[2][2][2], im[2][2][2]; };
struct FB2 { float re[4][2], im[4][2]; };

__device__ __forceinline__ unsigned tf32_rna_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void tf32_split(float x, unsigned &hi, unsigned &lo) {
    hi = tf32_rna_bits(x);
    lo = __float_as_uint(x - __uint_as_float(hi));   // the tensor core truncates it to TF32 (error below 2^-21 |x|, sign of lo)
}
__device__ __forceinline__ void hmma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// FB2 of X from the FAcc2 of the SAME matrix: the layout change of acc_to_bfrag (frag.cuh) with 32-bit shuffles.
__device__ __forceinline__ void facc_to_fb(FB2 &B, const FAcc2 &X, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const bool odd = g & 1;
    const int src1 = 4 * (2 * q + (g & 1)) + (g >> 1);
    const int src2 = 4 * (2 * q + 1 - (g & 1)) + (g >> 1);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const float r1 = __shfl_sync(0xffffffffu, odd ? X.re[kb][nt][1] : X.re[kb][nt][0], src1);
            const float r2 = __shfl_sync(0xffffffffu, odd ? X.re[kb][nt][0] : X.re[kb][nt][1], src2);
            const float i1 = __shfl_sync(0xffffffffu, odd ? X.im[kb][nt][1] : X.im[kb][nt][0], src1);
            const float i2 = __shfl_sync(0xffffffffu, odd ? X.im[kb][nt][0] : X.im[kb][nt][1], src2);
            B.re[2 * kb][nt] = odd ? r2 : r1;
            B.re[2 * kb + 1][nt] = odd ? r1 : r2;
            B.im[2 * kb][nt] = odd ? i2 : i1;
            B.im[2 * kb + 1][nt] = odd ? i1 : i2;
        }
}

// FB2 of X^H from the registers of FAcc2 X (conj_transpose_as_bfrag of frag.cuh): the right-operand layout of a Hermitian X.
__device__ __forceinline__ void fconj_transpose_as_fb(FB2 &B, const FAcc2 &X) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            B.re[kt][nt] = X.re[nt][kt >> 1][kt & 1];
            B.im[kt][nt] = -X.im[nt][kt >> 1][kt & 1];
        }
}

// D = A * B (complex 16 x 16, fp32 grade): 48 TF32 MMAs.  Per column tile four accumulators (Ar Br, Ai Bi, Ar Bi, Ai Br), each
// fed its three split terms for both k-tiles from zero, recombined with round-to-nearest FADDs.
__device__ __forceinline__ void tf32_cmul16(FAcc2 &D, const FAcc2 &A, const FB2 &B) {
    unsigned arh[2][4], arl[2][4], aih[2][4], ail[2][4];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
        tf32_split(A.re[0][kt][0], arh[kt][0], arl[kt][0]); tf32_split(A.re[1][kt][0], arh[kt][1], arl[kt][1]);
        tf32_split(A.re[0][kt][1], arh[kt][2], arl[kt][2]); tf32_split(A.re[1][kt][1], arh[kt][3], arl[kt][3]);
        tf32_split(A.im[0][kt][0], aih[kt][0], ail[kt][0]); tf32_split(A.im[1][kt][0], aih[kt][1], ail[kt][1]);
        tf32_split(A.im[0][kt][1], aih[kt][2], ail[kt][2]); tf32_split(A.im[1][kt][1], aih[kt][3], ail[kt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        float rr[4] = {0.f, 0.f, 0.f, 0.f}, ii[4] = {0.f, 0.f, 0.f, 0.f}, ri[4] = {0.f, 0.f, 0.f, 0.f}, ir[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            unsigned brh[2], brl[2], bih[2], bil[2];
            tf32_split(B.re[2 * kt][nt], brh[0], brl[0]); tf32_split(B.re[2 * kt + 1][nt], brh[1], brl[1]);
            tf32_split(B.im[2 * kt][nt], bih[0], bil[0]); tf32_split(B.im[2 * kt + 1][nt], bih[1], bil[1]);
            // small cross terms first
            hmma_tf32(rr, arl[kt], brh[0], brh[1]); hmma_tf32(ii, ail[kt], bih[0], bih[1]);
            hmma_tf32(ri, arl[kt], bih[0], bih[1]); hmma_tf32(ir, ail[kt], brh[0], brh[1]);
            hmma_tf32(rr, arh[kt], brl[0], brl[1]); hmma_tf32(ii, aih[kt], bil[0], bil[1]);
            hmma_tf32(ri, arh[kt], bil[0], bil[1]); hmma_tf32(ir, aih[kt], brl[0], brl[1]);
            hmma_tf32(rr, arh[kt], brh[0], brh[1]); hmma_tf32(ii, aih[kt], bih[0], bih[1]);
            hmma_tf32(ri, arh[kt], bih[0], bih[1]); hmma_tf32(ir, aih[kt], brh[0], brh[1]);
        }
        D.re[0][nt][0] = rr[0] - ii[0]; D.re[0][nt][1] = rr[1] - ii[1]; D.re[1][nt][0] = rr[2] - ii[2]; D.re[1][nt][1] = rr[3] - ii[3];
        D.im[0][nt][0] = ri[0] + ir[0]; D.im[0][nt][1] = ri[1] + ir[1]; D.im[1][nt][0] = ri[2] + ir[2]; D.im[1][nt][1] = ri[3] + ir[3];
    }
}

}  // namespace pb
