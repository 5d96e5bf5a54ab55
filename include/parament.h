/*
 * parament.h -- C-ABI of the B200-native Parament_equiprop library (libparament.so).
 *
 * Drop-in boundary: every entry point in section 1 has the name, argument order, argument meaning,
 * return codes and error strings of the reference's public header
 * (/root/reference/src/cuda/parament.h, cited per declaration as "ref parament.h:LINE"), so the
 * unchanged pyparament ctypes wrapper (/root/reference/src/python/pyparament/parament/paramentlib.py:57-72)
 * binds it without modification.  Section 2 are additive entry points (ensembles, device-resident
 * operands, timing, multi-GPU slices) that the reference does not have; nothing in section 1
 * depends on them.
 *
 * Conventions (same as the reference):
 *   - complex64 / complex128 operands are interleaved (re, im) pairs == cuComplex / cuDoubleComplex
 *     == numpy complex64 / complex128.
 *   - matrices are dim x dim, passed as the row-major (C-order) buffers the wrapper produces
 *     (parament.py:197-212); `out` receives the physical time-ordered propagator
 *     U = U_{N-1} ... U_1 U_0 in the same row-major layout (parament.py:273).
 *   - all pointer arguments of section 1 are HOST pointers owned by the caller and only read/written
 *     during the call; calls are synchronous.
 *   - every function returns a Parament_ErrorCode (C int) unless stated; the code of the last call is
 *     also kept in the context (Parament_peekAtLastError).
 *   - a context is not thread-safe; distinct contexts may be used from distinct threads.
 */
#ifndef PARAMENT_B200_PARAMENT_H_
#define PARAMENT_B200_PARAMENT_H_

#ifdef __cplusplus
extern "C" {
#else
#include <stdbool.h>
#endif

#if defined(__GNUC__)
#define PARAMENT_API __attribute__((visibility("default")))
#else
#define PARAMENT_API
#endif

/* Opaque contexts (ref parament.h:56-57). */
struct Parament_Context_f32;
struct Parament_Context_f64;

/* Interleaved complex operands; layout-compatible with cuComplex / cuDoubleComplex. */
#ifndef PARAMENT_NO_COMPLEX_TYPEDEFS
typedef struct Parament_c64 { float re, im; } Parament_c64;
typedef struct Parament_c128 { double re, im; } Parament_c128;
#endif

/* Quadrature rule selector (ref parament.h:64-69; values are part of the ABI, mirrored in constants.py:29-31). */
typedef enum Parament_QuadratureSpec {
    PARAMENT_QUADRATURE_NONE = 0x00000000,
    PARAMENT_QUADRATURE_MIDPOINT = 0x01000000,
    PARAMENT_QUADRATURE_SIMPSON = 0x02000000
} Parament_QuadratureSpec;

/* Error codes (ref parament.h:74-130; mirrored in constants.py:17-27).  Codes 30 and 60 keep their numeric
 * value although there is no cuBLAS in this library: 30 = CUDA runtime / device initialisation failed,
 * 60 = a kernel launch or device operation failed. */
typedef enum Parament_ErrorCode {
    PARAMENT_STATUS_SUCCESS = 0,
    PARAMENT_STATUS_HOST_ALLOC_FAILED = 10,
    PARAMENT_STATUS_DEVICE_ALLOC_FAILED = 20,
    PARAMENT_STATUS_CUBLAS_INIT_FAILED = 30,
    PARAMENT_STATUS_INVALID_VALUE = 50,
    PARAMENT_STATUS_CUBLAS_FAILED = 60,
    PARAMENT_STATUS_SELECT_SMALLER_DT = 70,
    PARAMENT_STATUS_NO_HAMILTONIAN = 80,
    PARAMENT_STATUS_INVALID_QUADRATURE_SELECTION = 90,
    PARAMENT_FAIL = 1000
} Parament_ErrorCode;

/* ===================================================================================================
 * Section 1 -- the reference's C-ABI (20 exported symbols of the reference build + device_info)
 * =================================================================================================== */

/* ref parament.h:155 (impl parament.cpp:52-128).  Creates a complex64 context on the current CUDA device.
 * 10 on host allocation failure, 30 if no usable CUDA device, 20 on device allocation failure. */
PARAMENT_API Parament_ErrorCode Parament_create(struct Parament_Context_f32 **handle_p);

/* ref parament.h:177 (impl parament.cpp:186-205).  NULL is accepted and ignored.  Waits for the device. */
PARAMENT_API Parament_ErrorCode Parament_destroy(struct Parament_Context_f32 *handle);

/* ref parament.h:219 (impl parament.cpp:211-368).  H0: dim*dim; H1: amps consecutive dim*dim matrices.
 * use_magnus requires quadrature_mode == SIMPSON, otherwise 90 is returned and the context is left with
 * NO Hamiltonian (parament.cpp:220-226,365-367).  The series norm is
 * Hnorm = rowsum(H0) + sum_k rowsum(H_k) (parament.cpp:280-284). */
PARAMENT_API Parament_ErrorCode Parament_setHamiltonian(struct Parament_Context_f32 *handle, const Parament_c64 *H0,
                                                        const Parament_c64 *H1, unsigned int dim, unsigned int amps,
                                                        bool use_magnus, enum Parament_QuadratureSpec quadrature_mode);

/* ref parament.h:261 (impl parament.cpp:791-851).  carr: amps arrays of pts amplitudes each, concatenated
 * (parament.h:225-226); amps may be smaller than the count given to setHamiltonian, missing controls have
 * zero amplitude (parament.h:228-230).  out: dim*dim.  80 without a Hamiltonian, 70 when the automatic
 * iteration count would exceed the table (parament.cpp:386-388). */
PARAMENT_API Parament_ErrorCode Parament_equiprop(struct Parament_Context_f32 *handle, const Parament_c64 *carr, double dt,
                                                  unsigned int pts, unsigned int amps, Parament_c64 *out);

/* ref parament.h:286 / :400 (impl parament.cpp:723-766).  Pure table look-up on H_norm*dt; -1 beyond the table. */
PARAMENT_API int Parament_selectIterationCycles_fp32(double H_norm, double dt);
PARAMENT_API int Parament_selectIterationCycles_fp64(double H_norm, double dt);

/* ref parament.h:308 (impl parament.cpp:772-776): fix the Chebyshev degree MMAX for subsequent calls.
 * Unlike the reference (correct only for odd MMAX >= 3) any cycles >= 1 is evaluated
 * correctly. */
PARAMENT_API Parament_ErrorCode Parament_setIterationCyclesManually(struct Parament_Context_f32 *handle, unsigned int cycles);

/* ref parament.h:327 (impl parament.cpp:782-786): back to the table-driven choice. */
PARAMENT_API Parament_ErrorCode Parament_automaticIterationCycles(struct Parament_Context_f32 *handle);

/* ref parament.h:347 (impl parament.cpp:854-857): code of the last call on this context; does not modify it. */
PARAMENT_API Parament_ErrorCode Parament_peekAtLastError(struct Parament_Context_f32 *handle);

/* ref parament.h:357 (impl parament.cpp:859-882): static, byte-identical message strings. */
PARAMENT_API const char *Parament_errorMessage(Parament_ErrorCode errorCode);

/* complex128 variants: ref parament.h:363,368,373,379,385,390,395 (impl parament.cpp:925-954). */
PARAMENT_API Parament_ErrorCode Parament_create_fp64(struct Parament_Context_f64 **handle_p);
PARAMENT_API Parament_ErrorCode Parament_destroy_fp64(struct Parament_Context_f64 *handle);
PARAMENT_API Parament_ErrorCode Parament_setHamiltonian_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *H0,
                                                             const Parament_c128 *H1, unsigned int dim, unsigned int amps,
                                                             bool use_magnus, Parament_QuadratureSpec quadrature_mode);
PARAMENT_API Parament_ErrorCode Parament_equiprop_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *carr,
                                                       double dt, unsigned int pts, unsigned int amps, Parament_c128 *out);
PARAMENT_API Parament_ErrorCode Parament_setIterationCyclesManually_fp64(struct Parament_Context_f64 *handle, unsigned int cycles);
PARAMENT_API Parament_ErrorCode Parament_automaticIterationCycles_fp64(struct Parament_Context_f64 *handle);
PARAMENT_API Parament_ErrorCode Parament_peekAtLastError_fp64(struct Parament_Context_f64 *handle);

/* ref mathhelper.h:67-68 (impl mathhelper.cpp:82-118): max over rows of the sum of |entries| of the
 * row-major buffer. */
PARAMENT_API double OneNorm(const Parament_c64 *mat, unsigned int dim);
PARAMENT_API double OneNorm_fp64(const Parament_c128 *mat, unsigned int dim);

/* ref deviceinfo.h:35 (impl deviceInfo.c:30-59): print the CUDA device table to stdout.  Resolved by the
 * wrapper at import time (paramentlib.py:72). */
PARAMENT_API void device_info(void);

/* The wrapper's Parament._get_error_message() calls this name (parament.py:284); the reference never
 * defined it.  Alias of Parament_peekAtLastError for either context type. */
PARAMENT_API Parament_ErrorCode Parament_getLastError(void *handle);

/* ===================================================================================================
 * Section 2 -- additive entry points (not in the reference)
 * =================================================================================================== */

/* Ensemble: `batch` independent pulses through the same Hamiltonian in one call (GRAPE sweeps,
 * BASELINE.json configs[4]).  carr: batch x amps x pts (pulse-major), out: batch x dim x dim. */
PARAMENT_API Parament_ErrorCode Parament_equipropBatch(struct Parament_Context_f32 *handle, const Parament_c64 *carr,
                                                       double dt, unsigned int pts, unsigned int amps, unsigned int batch,
                                                       Parament_c64 *out);
PARAMENT_API Parament_ErrorCode Parament_equipropBatch_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *carr,
                                                            double dt, unsigned int pts, unsigned int amps, unsigned int batch,
                                                            Parament_c128 *out);

/* Device-resident variant: carr_dev / out_dev are DEVICE pointers on the context's device; the work is
 * enqueued on `stream` (a cudaStream_t passed as void*, NULL = the context's own stream) and the call
 * returns without synchronising when `stream` is non-NULL.  Same layouts as Parament_equipropBatch.
 * The context's scratch memory is ordered by that stream (and by internal streams joined to it): successive calls on one
 * context must use the same stream or be separated by a synchronisation, as with a cuBLAS workspace. */
PARAMENT_API Parament_ErrorCode Parament_equipropDevice(struct Parament_Context_f32 *handle, const Parament_c64 *carr_dev,
                                                        double dt, unsigned int pts, unsigned int amps, unsigned int batch,
                                                        Parament_c64 *out_dev, void *stream);
PARAMENT_API Parament_ErrorCode Parament_equipropDevice_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *carr_dev,
                                                             double dt, unsigned int pts, unsigned int amps, unsigned int batch,
                                                             Parament_c128 *out_dev, void *stream);

/* Time-slice variant for multi-GPU runs: propagate only effective steps [step_lo, step_hi) of a pulse whose
 * full coefficient arrays (pts points per control) are given.  Partial propagators of consecutive slices
 * combine as P_total = P_last ... P_1 P_0 (Parament_combine).  Host pointers. */
PARAMENT_API Parament_ErrorCode Parament_equipropSlice(struct Parament_Context_f32 *handle, const Parament_c64 *carr, double dt,
                                                       unsigned int pts, unsigned int amps, unsigned long long step_lo,
                                                       unsigned long long step_hi, Parament_c64 *out);
PARAMENT_API Parament_ErrorCode Parament_equipropSlice_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *carr,
                                                            double dt, unsigned int pts, unsigned int amps,
                                                            unsigned long long step_lo, unsigned long long step_hi,
                                                            Parament_c128 *out);

/* The same with the partial propagator left on the GPU: carr is a HOST pointer, out_dev a DEVICE pointer (dim x dim, context
 * precision, on the context's device).  The last kernel writes the partial there and the call returns when it is complete --
 * the hand-off point to an NCCL exchange of the partials in a one-process-per-GPU run (bench.py --gpus N). */
PARAMENT_API Parament_ErrorCode Parament_equipropSliceToDevice(struct Parament_Context_f32 *handle, const Parament_c64 *carr,
                                                               double dt, unsigned int pts, unsigned int amps,
                                                               unsigned long long step_lo, unsigned long long step_hi,
                                                               Parament_c64 *out_dev);
PARAMENT_API Parament_ErrorCode Parament_equipropSliceToDevice_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *carr,
                                                                    double dt, unsigned int pts, unsigned int amps,
                                                                    unsigned long long step_lo, unsigned long long step_hi,
                                                                    Parament_c128 *out_dev);

/* Ordered product of `count` dim x dim partial propagators (host pointers, parts[0] is the earliest slice):
 * out = parts[count-1] ... parts[1] parts[0], evaluated on the device in complex128. */
PARAMENT_API Parament_ErrorCode Parament_combine(struct Parament_Context_f32 *handle, const Parament_c64 *parts,
                                                 unsigned int count, Parament_c64 *out);
PARAMENT_API Parament_ErrorCode Parament_combine_fp64(struct Parament_Context_f64 *handle, const Parament_c128 *parts,
                                                      unsigned int count, Parament_c128 *out);

/* Device-resident combine (either context type): parts_dev holds `count` dim x dim propagators in the context precision on
 * the context's device, earliest slice first; out_dev receives their ordered product.  Enqueued on `stream` (cudaStream_t
 * as void*) without synchronising, or on the context's stream followed by a synchronisation when stream is NULL. */
PARAMENT_API Parament_ErrorCode Parament_combineDevice(void *handle, const void *parts_dev, unsigned int count, void *out_dev,
                                                       void *stream);

/* Introspection of the last equiprop on this context (either context type).  Keys:
 *   0 device milliseconds between first and last kernel (CUDA events on the context stream)
 *   1 number of kernels launched          2 Chebyshev degree used (MMAX actually evaluated)
 *   3 degree the reference table selects  4 effective steps per pulse N
 *   5 kernel family used (1 = register-resident DMMA warp kernel, 2 = persistent CTA chain kernel,
 *                         3 = batched GEMM pipeline over L2-resident time chunks)
 *   6 H2D bytes copied                    7 D2H bytes copied
 *   8 Hnorm of the loaded Hamiltonian (the reference's bound, parament.cpp:280-284: sum of max-row-abs-sums)
 *  14 norm bound the series of the last call was built for: Hnorm for dim <= 16; for dim > 16 the spectral bound
 *     1.05 (s(H0) + sum_k max_t|c_k(t)| s(H_k)) (largest singular values, amplitude maxima measured per call), never above Hnorm.
 *     Key 3 (the reference table's degree) and error 70 always follow Hnorm; key 2 follows key 14.
 *   9 series evaluation (0 Clenshaw recurrence, 1 Horner in Y^2, 2 Paterson-Stockmeyer blocks of four,
 *                        3 degree 8 in three products, 4 degree 12 in four products -- all the same polynomial family)
 *  10 complex matrix products executed per effective step (series + ordered product); fractional for dim > 64 when the inputs are
 *     Hermitian and the amplitudes real: of the Hermitian square Y Y only the tiles touching the upper triangle are computed
 *     (20 of 32 at dim 256), the rest is mirrored by the epilogue (PARAMENT_K4_HERM=0 computes all tiles)
 *  13 real matrix products per complex product in the kernel family used, averaged over the products of a step (4; 3 for the
 *     batched GEMM of dim > 64 and for the degree-8 form of complex64 contexts at dim 9..16; 3.25 / 3.4 for the shared-memory-resident
 *     kernel of dim 17..64 in its degree-8 / degree-12 form)
 *  15 arithmetic of the last call: 0 = double precision on the FP64 tensor pipe (DMMA); 1 = single precision as 3xTF32 split
 *     products on the warp-level tensor path (complex64 contexts, dim <= 8, accumulated phase N h (s(H0) + sum_k s(H_k)) <= 128:
 *     the range in which the measured error stays below half the 1e-5 tolerance, profiles/error_growth_tf32_r2.md);
 *     2 = mixed: double precision for X^2, the first- and second-order terms and the running product, 3xTF32 for the two small
 *     products of the degree-8 form (complex64 contexts, dim 9..16, accumulated phase <= 5e5; PARAMENT_K1_MIXED=0 / 1 forces it)
 * Environment switches read at Parament_create (development / A-B testing): PARAMENT_SERIES=clenshaw forces the reference's
 * recurrence, PARAMENT_SERIES=horner the Horner / Paterson-Stockmeyer forms; PARAMENT_NO_ONCHIP=1 selects the L2-scratch
 * chain kernel for dim 17..64; PARAMENT_DEVICE the default CUDA device; PARAMENT_F3_STREAMS=1..4 the chunks in flight for dim > 64
 * (default 4); PARAMENT_K1_3M=0 the four-real-product chain kernel at dim 9..16 (read once per process); PARAMENT_K4_3M=0..3 the complex product of the batched GEMM (0: four real products on 64x64 tiles; 3, default: three real
 * products on 64x32 tiles); PARAMENT_K4_FEED=tma the bulk-copy (TMA engine) operand feed of that GEMM in mode 0 (measured slower than cp.async);
 * PARAMENT_COPY_GROUPS=1..8 the copy/compute groups of the host-pointer pipeline (default: up to six, sizes doubling); PARAMENT_NORM=reference builds the series for
 * Hnorm at every dimension (A/B of the spectral bound); PARAMENT_C64_MATH=f64|tf32 forces the arithmetic of complex64 contexts with
 * dim <= 8 (default: by step count, key 15), PARAMENT_TF32_MAX_PHASE moves that bound (both read per call), PARAMENT_TF32_COMP=1
 * switches the compensated running product of the TF32 kernel on (read once per process; measured without effect);
 * PARAMENT_K1_PACK=0 (read at Parament_setHamiltonian) keeps one system per 8 x 8 tensor-pipe tile at dim <= 4 (default: four
 * systems of dim <= 2 or two of dim 3..4 share the tile, each advancing through its own part of the step range; with it complex64
 * contexts of dim <= 4 compute in double precision, key 15 = 0); PARAMENT_K1_MIN_STEPS=1..64 the fewest steps per warp of a short
 * single pulse (default 8); PARAMENT_K1_HERM=0 (read at Parament_setHamiltonian) ignores that the loaded matrices are Hermitian
 * (dim 9..16, complex64: a Hermitian step Hamiltonian is its own right operand up to conjugation, which saves a register shuffle).
 * Limits: at most 64 effective control terms per step (controls + Magnus commutators: amps <= 64 without Magnus, amps <= 9
 * with it); Parament_setHamiltonian returns PARAMENT_STATUS_INVALID_VALUE beyond that (the reference has no stated limit but
 * its launch configurations break at amps > 16 with Magnus, control_expansion.cu:179).
 * Every entry point selects the context's device for its own duration and restores the caller's current device on return. */
PARAMENT_API double Parament_lastStat(void *handle, int key);

/* Select the CUDA device a context lives on.  Must be called before setHamiltonian; default is device 0
 * (reference: device 0 implicit, parament.cpp:108) or $PARAMENT_DEVICE. */
PARAMENT_API Parament_ErrorCode Parament_setDevice(void *handle, int device);

/* Single-process multi-GPU (the caller is ONE ctypes call, so one host process drives all devices).
 * Parament_setDevices(h, n): use n devices starting at the context's own (n <= 0 or n > visible: all visible devices;
 * n = 1: back to one device).  Parament_setDeviceList(h, devices, count): explicit list, devices[0] is where the context
 * lives (moving it drops the Hamiltonian, like Parament_setDevice); an entry may repeat.  $PARAMENT_NUM_GPUS at
 * Parament_create has the effect of Parament_setDevices, so the unchanged reference wrapper gets the mode as well.
 * Afterwards every host-pointer Parament_equiprop[_fp64] cuts the time axis into contiguous slices, one per device: each
 * device copies only its slice of the caller's arrays, reduces it to a partial propagator on its own stream (one host
 * thread per device), the dim x dim partials travel to the first device by peer copy (NVLink when peer access exists)
 * and are multiplied in order there.  Parament_equipropBatch[_fp64] cuts the pulses of the ensemble into contiguous
 * ranges instead (no exchange).  Calls too small to share (fewer than max(64, 2^26 / npad^3) effective steps per
 * device) run on the first device alone.  Parament_lastStat key 11 = devices that took part in the last call, key 12 =
 * devices configured; key 0 is then the host wall clock of the whole shared call in milliseconds.
 * The device-pointer entry points are not shared: they run on the context's own device. */
PARAMENT_API Parament_ErrorCode Parament_setDevices(void *handle, int ngpus);
PARAMENT_API Parament_ErrorCode Parament_setDeviceList(void *handle, const int *devices, int count);

/* Pipe-peak microbenchmark on the current CUDA device, used as roofline denominator by bench.py.
 * kind: 0 FP32 FFMA TFLOP/s, 1 FP64 DFMA TFLOP/s, 2 FP64 tensor-pipe DMMA TFLOP/s, 3 HBM copy GB/s,
 * 4 TF32 warp-level tensor path (mma.sync.m16n8k8) TFLOP/s. */
PARAMENT_API double Parament_measurePeak(int kind);

/* Library identification string, e.g. "parament-b200 0.1 (sm_100a)". */
PARAMENT_API const char *Parament_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PARAMENT_B200_PARAMENT_H_ */
