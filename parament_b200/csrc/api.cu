// api.cu -- C-ABI of libparament.so and the host orchestration of Parament_equiprop.
//
// Mirrors the reference's host layer (/root/reference/src/cuda/parament.cpp:51-954) function for function
// at the ABI, with its own context and a different device pipeline:
//   reference  coefficients -> transfer -> expand (X for all steps in HBM) -> MMAX batched cuBLAS GEMMs
//              + diagonal_add -> log2(N) batched GEMMs
//   here       coefficients -> transfer -> ONE fused chain kernel (+ one small ordered reduction) for
//              dim <= 16, or an L2-resident chunked tensor-pipe pipeline for larger dim.
// There is no CPU fallback: without a CUDA device Parament_create fails with code 30.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <new>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler injects its library

#include "context.hpp"
#include "k1_warp.hpp"
#include "k4_gemm.hpp"
#include "poly_solve.hpp"
#include "series.hpp"

using namespace pb;

namespace {

#define PB_CUDA_OK(expr) ((expr) == cudaSuccess)

// Handles are checked against the set of live contexts: a handle used after Parament_destroy (the wrapper's
// use-after-destroy test) is rejected without touching freed memory.
std::mutex g_live_mu;
std::unordered_set<const void *> g_live;

inline Context *as_ctx(void *h) {
    if (!h) return nullptr;
    std::lock_guard<std::mutex> lk(g_live_mu);
    return g_live.count(h) ? reinterpret_cast<Context *>(h) : nullptr;
}

// Every entry point works on the context's device and puts the caller's current device back on exit (the reference never
// changes it; Parament_equipropDevice is meant for torch / cupy pointers, whose owner expects its device to stay current).
struct DeviceGuard {
    int prev = -1, dev;
    explicit DeviceGuard(int device) : dev(device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// NVTX range for the phases of the path (visible in Nsight Systems / `ncu --nvtx`): the reference only has a commented-out
// nvtxMarkA (parament.cpp:361).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

Parament_ErrorCode fail(Context *c, Parament_ErrorCode code) {
    if (c) c->lastError = code;
    cudaGetLastError();   // clear a sticky-free error state
    return code;
}

void free_dev(DeviceBuffer &b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
}

bool ensure_dev(DeviceBuffer &b, size_t bytes) {
    if (b.bytes >= bytes && b.ptr) return true;
    free_dev(b);
    if (bytes == 0) bytes = 16;
    if (!PB_CUDA_OK(cudaMalloc(&b.ptr, bytes))) { b.ptr = nullptr; cudaGetLastError(); return false; }
    b.bytes = bytes;
    return true;
}

bool create_copy_events(Context *c) {
    for (int i = 0; i < kCopyEvents; ++i)
        if (!PB_CUDA_OK(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming))) return false;
    return true;
}
void destroy_copy_events(Context *c) {
    for (int i = 0; i < kCopyEvents; ++i)
        if (c->ev_copy[i]) { cudaEventDestroy(c->ev_copy[i]); c->ev_copy[i] = nullptr; }
}

// Streams, events and the pinned scalar buffer of a context on its (current) device.  destroy_device_objects nulls every handle
// it destroys, so it is safe on a partially created set and cannot destroy a handle twice.
bool create_device_objects(Context *c) {
    return PB_CUDA_OK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device)) &&
           PB_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) &&
           PB_CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) && create_copy_events(c) &&
           PB_CUDA_OK(cudaEventCreate(&c->ev_start)) && PB_CUDA_OK(cudaEventCreate(&c->ev_stop)) &&
           PB_CUDA_OK(cudaMallocHost(&c->h_absmax, (kMaxTerms + 1) * sizeof(double)));
}
void destroy_device_objects(Context *c) {
    if (c->stager) { c->stager->stop(); delete c->stager; c->stager = nullptr; }
    free_dev(c->d_gather); free_dev(c->d_absmax); free_dev(c->d_counters);
    free_dev(c->d_H); free_dev(c->d_carr); free_dev(c->d_out); free_dev(c->d_partials);
    free_dev(c->d_Y); free_dev(c->d_comb); free_dev(c->d_comb2); free_dev(c->d_pending); free_dev(c->d_tree);
    if (c->h_absmax) { cudaFreeHost(c->h_absmax); c->h_absmax = nullptr; }
    if (c->ev_start) { cudaEventDestroy(c->ev_start); c->ev_start = nullptr; }
    if (c->ev_stop) { cudaEventDestroy(c->ev_stop); c->ev_stop = nullptr; }
    destroy_copy_events(c);
    for (cudaStream_t &q : c->f3_streams) if (q) { cudaStreamDestroy(q); q = nullptr; }
    if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); c->copy_stream = nullptr; }
    if (c->stream) { cudaStreamDestroy(c->stream); c->stream = nullptr; }
}

// ---------------------------------------------------------------------------------------------------
// create / destroy                                                   (reference parament.cpp:51-205)
// ---------------------------------------------------------------------------------------------------
Parament_ErrorCode set_device_list(Context *c, const int *devices, int count);

Parament_ErrorCode create_ctx_on(Context **out, bool fp64, int dev) {
    if (!out) return PARAMENT_STATUS_INVALID_VALUE;
    *out = nullptr;
    Context *c = new (std::nothrow) Context();
    if (!c) return PARAMENT_STATUS_HOST_ALLOC_FAILED;
    c->fp64 = fp64;
    int ndev = 0;
    if (!PB_CUDA_OK(cudaGetDeviceCount(&ndev)) || ndev < 1) {
        cudaGetLastError();
        delete c;
        return PARAMENT_STATUS_CUBLAS_INIT_FAILED;   // code 30: device initialisation failed
    }
    if (dev < 0) {
        dev = 0;
        if (const char *e = getenv("PARAMENT_DEVICE")) dev = atoi(e);
    }
    if (dev < 0 || dev >= ndev) dev = 0;
    c->device = dev;
    DeviceGuard guard(dev);
    if (!create_device_objects(c)) {
        destroy_device_objects(c);   // whatever was created before the failing call
        cudaGetLastError();
        delete c;
        return PARAMENT_STATUS_CUBLAS_INIT_FAILED;
    }
    if (const char *e = getenv("PARAMENT_SERIES")) c->series_mode = (strcmp(e, "clenshaw") == 0) ? 1 : (strcmp(e, "horner") == 0 ? 2 : 0);
    if (const char *e = getenv("PARAMENT_NORM")) c->norm_mode = strcmp(e, "reference") == 0 ? 0 : 1;
    c->lastError = PARAMENT_STATUS_SUCCESS;
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live.insert(c);
    }
    *out = c;
    return PARAMENT_STATUS_SUCCESS;
}

Parament_ErrorCode set_device_count(Context *c, int ngpus);

Parament_ErrorCode create_ctx(Context **out, bool fp64) {
    Parament_ErrorCode ec = create_ctx_on(out, fp64, -1);
    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    // $PARAMENT_NUM_GPUS: the unchanged reference wrapper gets the single-process multi-GPU mode without a new call
    // (0 = every visible device; more than there are = every visible device)
    if (const char *e = getenv("PARAMENT_NUM_GPUS")) {
        const int want = atoi(e);
        if (want != 1 && set_device_count(*out, want) != PARAMENT_STATUS_SUCCESS) (*out)->lastError = PARAMENT_STATUS_SUCCESS;
    }
    return PARAMENT_STATUS_SUCCESS;
}

void destroy_peers(Context *c);

Parament_ErrorCode destroy_ctx(Context *c) {
    if (!c) return PARAMENT_STATUS_SUCCESS;   // NULL is a no-op (parament.cpp:190-191)
    destroy_peers(c);
    if (c->worker) { c->worker->stop(); delete c->worker; c->worker = nullptr; }
    {
        DeviceGuard guard(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        destroy_device_objects(c);
        cudaGetLastError();
    }
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live.erase(c);
    }
    c->magic = 0;
    delete c;
    return PARAMENT_STATUS_SUCCESS;
}

void destroy_peers(Context *c) {
    for (Context *p : c->peers) destroy_ctx(p);
    c->peers.clear();
}

// ---------------------------------------------------------------------------------------------------
// setHamiltonian                                                    (reference parament.cpp:211-368)
// ---------------------------------------------------------------------------------------------------
template <typename T>
double one_norm_t(const T *m, unsigned int dim);
template <>
double one_norm_t<Parament_c64>(const Parament_c64 *m, unsigned int dim) {
    double best = 0;   // mathhelper.cpp:82-97: float modulus, double accumulation
    for (unsigned i = 0; i < dim; ++i) {
        double s = 0;
        for (unsigned j = 0; j < dim; ++j) s += (double)hypotf(m[(size_t)dim * i + j].re, m[(size_t)dim * i + j].im);
        best = std::max(best, s);
    }
    return best;
}
template <>
double one_norm_t<Parament_c128>(const Parament_c128 *m, unsigned int dim) {
    double best = 0;   // mathhelper.cpp:99-114
    for (unsigned i = 0; i < dim; ++i) {
        double s = 0;
        for (unsigned j = 0; j < dim; ++j) s += hypot(m[(size_t)dim * i + j].re, m[(size_t)dim * i + j].im);
        best = std::max(best, s);
    }
    return best;
}

void commutator(const zc *A, const zc *B, zc *out, int n) {   // out = A B - B A
    std::vector<zc> t((size_t)n * n, zc(0, 0));
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) {
            const zc a = A[(size_t)i * n + k], b = B[(size_t)i * n + k];
            const zc *Bk = B + (size_t)k * n, *Ak = A + (size_t)k * n;
            zc *ti = t.data() + (size_t)i * n;
            for (int j = 0; j < n; ++j) ti[j] += a * Bk[j] - b * Ak[j];
        }
    std::copy(t.begin(), t.end(), out);
}

// Largest singular value of an n x n matrix: power iteration on A^H A from a fixed pseudo-random start (deterministic).
// The Rayleigh quotient converges from below; the caller adds a safety margin.  Cost O(iters n^2) on the host, once per
// Hamiltonian (the reference's norm, parament.cpp:280-284, is the looser max-row-abs-sum).
double spectral_norm(const zc *A, int n, int iters = 80) {
    std::vector<zc> v((size_t)n), w((size_t)n);
    unsigned long long seed = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < n; ++i) {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        const double a = (double)(seed >> 11) / 9007199254740992.0 - 0.5;
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        const double b = (double)(seed >> 11) / 9007199254740992.0 - 0.5;
        v[i] = zc(a, b);
    }
    double sigma = 0.0;
    for (int it = 0; it < iters; ++it) {
        double nv = 0.0;
        for (int i = 0; i < n; ++i) nv += std::norm(v[i]);
        nv = std::sqrt(nv);
        if (!(nv > 0.0)) return 0.0;
        for (int i = 0; i < n; ++i) v[i] /= nv;
        double nw = 0.0;
        for (int i = 0; i < n; ++i) {                 // w = A v
            zc acc(0, 0);
            const zc *Ai = A + (size_t)i * n;
            for (int j = 0; j < n; ++j) acc += Ai[j] * v[j];
            w[i] = acc;
            nw += std::norm(acc);
        }
        const double s_new = std::sqrt(nw);           // ||A v|| <= sigma_max, increasing with the iteration
        if (it > 8 && s_new <= sigma * (1.0 + 1e-7)) { sigma = std::max(sigma, s_new); break; }
        sigma = std::max(sigma, s_new);
        for (int j = 0; j < n; ++j) v[j] = zc(0, 0);  // v = A^H w
        for (int i = 0; i < n; ++i) {
            const zc *Ai = A + (size_t)i * n;
            const zc wi = w[i];
            for (int j = 0; j < n; ++j) v[j] += std::conj(Ai[j]) * wi;
        }
    }
    return sigma;
}

inline int pair_index(int j, int k, int A) { return j * A - j * (j + 1) / 2 + (k - j - 1); }   // j < k

// Upload the matrix table in the layout of the selected kernel family.
bool upload_matrices(Context *c) {
    const int n = c->dim;
    if (c->family == 1) {
        const int NT = c->npad / 8, NE = 2 * NT * NT;
        const int nb = c->npad / c->pack;   // packed small systems: kron(I_pack, H), one copy of H per diagonal block of nb rows
        std::vector<double2> frag((size_t)c->nmats * 2 * NE * 32, make_double2(0, 0));
        for (int m = 0; m < c->nmats; ++m) {
            const zc *M = c->mats.data() + (size_t)m * n * n;
            auto at = [&](int r, int col, double2 &dst) {
                if (r / nb != col / nb) return;
                const int rr = r % nb, cc = col % nb;
                if (rr < n && cc < n) dst = make_double2(M[(size_t)rr * n + cc].real(), M[(size_t)rr * n + cc].imag());
            };
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, q = lane & 3;
                for (int mt = 0; mt < NT; ++mt)            // AccFrag order (frag.cuh)
                    for (int nt = 0; nt < NT; ++nt)
                        for (int i = 0; i < 2; ++i) {
                            const int r = 8 * mt + g, col = 8 * nt + 2 * q + i, e = (mt * NT + nt) * 2 + i;
                            at(r, col, frag[(((size_t)m * 2 + 0) * NE + e) * 32 + lane]);
                        }
                for (int kt = 0; kt < 2 * NT; ++kt)        // BFrag order
                    for (int nt = 0; nt < NT; ++nt) {
                        const int r = 8 * (kt >> 1) + 2 * q + (kt & 1), col = 8 * nt + g, e = kt * NT + nt;
                        at(r, col, frag[(((size_t)m * 2 + 1) * NE + e) * 32 + lane]);
                    }
            }
        }
        if (!ensure_dev(c->d_H, frag.size() * sizeof(double2))) return false;
        return PB_CUDA_OK(cudaMemcpy(c->d_H.ptr, frag.data(), frag.size() * sizeof(double2), cudaMemcpyHostToDevice));
    }
    const int np = c->npad;
    std::vector<double2> tab((size_t)c->nmats * np * np, make_double2(0, 0));
    for (int m = 0; m < c->nmats; ++m)
        for (int r = 0; r < n; ++r)
            for (int col = 0; col < n; ++col) {
                const zc v = c->mats[((size_t)m * n + r) * n + col];
                tab[((size_t)m * np + r) * np + col] = make_double2(v.real(), v.imag());
            }
    if (!ensure_dev(c->d_H, tab.size() * sizeof(double2))) return false;
    if (!PB_CUDA_OK(cudaMemcpy(c->d_H.ptr, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice))) return false;
    if (c->enable_magnus) {
        // commutator slots on the device: T = B A, then C = A B - T (two launches of the tensor-pipe GEMM per commutator;
        // the reference does the same with cuBLAS, parament.cpp:300-355)
        const size_t nn = (size_t)np * np;
        const int A = c->amps;
        double2 *H = (double2 *)c->d_H.ptr;
        DeviceBuffer tmp;
        if (!ensure_dev(tmp, nn * sizeof(double2))) return false;
        auto comm = [&](int ia, int ib, int iout) -> bool {
            GemmArgs g{};
            g.n = np; g.batch = 1;
            g.A = H + (size_t)ib * nn; g.B = H + (size_t)ia * nn; g.D = (double2 *)tmp.ptr;        // T = B A
            if (k4_gemm(g, c->stream) != cudaSuccess) return false;
            GemmArgs h{};
            h.n = np; h.batch = 1;
            h.A = H + (size_t)ia * nn; h.B = H + (size_t)ib * nn; h.D = H + (size_t)iout * nn;    // A B - T
            h.C[0] = (const double2 *)tmp.ptr; h.beta[0] = cplx{-1.0, 0.0};
            return k4_gemm(h, c->stream) == cudaSuccess;
        };
        bool ok = true;
        for (int j = 0; j < A && ok; ++j) ok = comm(0, 1 + j, 1 + A + j);
        for (int j = 0; j < A && ok; ++j)
            for (int k = j + 1; k < A && ok; ++k) ok = comm(1 + j, 1 + k, 1 + 2 * A + pair_index(j, k, A));
        ok = ok && PB_CUDA_OK(cudaStreamSynchronize(c->stream));
        free_dev(tmp);
        if (!ok) return false;
    }
    return true;
}

template <typename T>
Parament_ErrorCode set_hamiltonian(Context *c, const T *H0, const T *H1, unsigned int dim, unsigned int amps,
                                   bool use_magnus, int quad) {
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    NvtxRange range("Parament_setHamiltonian");
    DeviceGuard guard(c->device);
    c->have_hamiltonian = false;   // a previous Hamiltonian is dropped first (parament.cpp:216)
    if (use_magnus && quad != PARAMENT_QUADRATURE_SIMPSON)
        return fail(c, PARAMENT_STATUS_INVALID_QUADRATURE_SELECTION);   // parament.cpp:220-226
    if (quad != PARAMENT_QUADRATURE_NONE && quad != PARAMENT_QUADRATURE_MIDPOINT && quad != PARAMENT_QUADRATURE_SIMPSON)
        return fail(c, PARAMENT_STATUS_INVALID_QUADRATURE_SELECTION);
    if (!H0 || (!H1 && amps > 0) || dim == 0) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    c->enable_magnus = use_magnus;
    c->quadrature = use_magnus ? PARAMENT_QUADRATURE_SIMPSON : quad;
    c->dim = (int)dim;
    c->amps = (int)amps;
    const size_t nn = (size_t)dim * dim;
    const int A = (int)amps;
    c->nmats = 1 + A + (use_magnus ? A + A * (A - 1) / 2 : 0);
    // effective control terms per step are limited to kMaxTerms = 64 (include/parament.h): rejected here, not at equiprop time
    if (c->nmats - 1 > kMaxTerms) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    try {
        c->mats.assign((size_t)c->nmats * nn, zc(0, 0));
        c->sigma_max.assign((size_t)1 + A, 0.0);
    } catch (const std::bad_alloc &) {
        return fail(c, PARAMENT_STATUS_HOST_ALLOC_FAILED);
    }
    for (size_t e = 0; e < nn; ++e) c->mats[e] = zc((double)H0[e].re, (double)H0[e].im);
    for (int a = 0; a < A; ++a)
        for (size_t e = 0; e < nn; ++e) c->mats[(size_t)(1 + a) * nn + e] = zc((double)H1[(size_t)a * nn + e].re, (double)H1[(size_t)a * nn + e].im);

    {   // exact Hermiticity of every input matrix (k1_warp.cu: the right-operand layout of a Hermitian X needs no shuffles)
        bool herm = !(getenv("PARAMENT_K1_HERM") && atoi(getenv("PARAMENT_K1_HERM")) == 0);
        for (int m = 0; m <= A && herm; ++m) {
            const zc *M = c->mats.data() + (size_t)m * nn;
            for (unsigned int r = 0; r < dim && herm; ++r)
                for (unsigned int col = r; col < dim; ++col)
                    if (M[(size_t)r * dim + col] != std::conj(M[(size_t)col * dim + r])) { herm = false; break; }
        }
        c->hermitian = herm;
    }

    // Series norm: sum of the max-row-abs-sums of the buffers as passed (parament.cpp:280-284)
    c->Hnorm = one_norm_t<T>(H0, dim);
    for (int a = 0; a < A; ++a) c->Hnorm += one_norm_t<T>(H1 + (size_t)a * nn, dim);

    c->pack = 1;
    if (dim <= 16) {
        c->family = 1; c->npad = dim <= 8 ? 8 : 16;
        // dim <= 4: two or four systems share the 8 x 8 tensor-pipe tile (k1_warp.cu "Packed small systems")
        const char *pk = getenv("PARAMENT_K1_PACK");
        if (dim <= 4 && !(pk && atoi(pk) == 0)) c->pack = dim <= 2 ? 4 : 2;
    }
    else if (dim <= 64) {
        c->family = 2; c->npad = k4_pad((int)dim);
        const int oc = !getenv("PARAMENT_NO_ONCHIP") ? k4_onchip_slots(c->npad, c->num_sms) : 0;
        c->onchip = oc > 0;
        c->k4_slots = c->onchip ? oc : k4_chain_slots(c->npad, c->num_sms);
    }
    else                { c->family = 3; c->npad = k4_pad((int)dim); c->k4_slots = k4_wave_slots(c->npad, c->num_sms); }
    c->series_cache.valid = false;
    // spectral bound of the step Hamiltonians (dim > 16, where it saves a matrix product per step; series_norm_for_call)
    // (also for complex64 contexts of dim <= 16: the accumulated phase decides between FP64 and TF32 / mixed arithmetic)
    if ((c->family != 1 && c->norm_mode == 1) || (!c->fp64 && c->family == 1))
        for (int m = 0; m <= A; ++m) c->sigma_max[m] = spectral_norm(c->mats.data() + (size_t)m * nn, (int)dim);

    // physical commutators [H0,H_j] and [H_j,H_k], j<k (parament.cpp:289-359): on the host for the
    // register-resident family (tiny matrices), on the device GEMM kernel otherwise (upload_matrices)
    if (use_magnus && c->family == 1) {
        const zc *h0 = c->mats.data();
        for (int j = 0; j < A; ++j) commutator(h0, h0 + (size_t)(1 + j) * nn, c->mats.data() + (size_t)(1 + A + j) * nn, (int)dim);
        for (int j = 0; j < A; ++j)
            for (int k = j + 1; k < A; ++k)
                commutator(h0 + (size_t)(1 + j) * nn, h0 + (size_t)(1 + k) * nn,
                           c->mats.data() + (size_t)(1 + 2 * A + pair_index(j, k, A)) * nn, (int)dim);
    }

    if (!upload_matrices(c)) return fail(c, PARAMENT_STATUS_DEVICE_ALLOC_FAILED);
    // single-process multi-GPU: every helper context gets the same Hamiltonian (constants are replicated)
    for (Context *p : c->peers) {
        const Parament_ErrorCode ec = set_hamiltonian<T>(p, H0, H1, dim, amps, use_magnus, quad);
        if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
    }
    c->have_hamiltonian = true;
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}

// The Hamiltonian of `c` loaded into the helper context `p` (helpers created after setHamiltonian): the host copies in
// double are exact conversions of the inputs, so converting back reproduces the caller's arrays bit for bit.
template <typename T>
Parament_ErrorCode replay_hamiltonian_t(const Context *c, Context *p) {
    const size_t nn = (size_t)c->dim * c->dim, cnt = (size_t)(1 + c->amps) * nn;
    std::vector<T> buf(cnt);
    for (size_t e = 0; e < cnt; ++e) {
        buf[e].re = static_cast<decltype(buf[e].re)>(c->mats[e].real());
        buf[e].im = static_cast<decltype(buf[e].im)>(c->mats[e].imag());
    }
    return set_hamiltonian<T>(p, buf.data(), buf.data() + nn, (unsigned int)c->dim, (unsigned int)c->amps, c->enable_magnus, c->quadrature);
}
Parament_ErrorCode replay_hamiltonian(const Context *c, Context *p) {
    p->MMAX = c->MMAX; p->MMAX_manual = c->MMAX_manual; p->series_mode = c->series_mode;
    if (!c->have_hamiltonian) return PARAMENT_STATUS_SUCCESS;
    try {
        return c->fp64 ? replay_hamiltonian_t<Parament_c128>(c, p) : replay_hamiltonian_t<Parament_c64>(c, p);
    } catch (const std::bad_alloc &) {
        return PARAMENT_STATUS_HOST_ALLOC_FAILED;
    }
}

// ---------------------------------------------------------------------------------------------------
// equiprop                                                          (reference parament.cpp:373-851)
// ---------------------------------------------------------------------------------------------------
struct CallSpec {
    unsigned int amps;              // control arrays per pulse present in the buffer
    unsigned int batch;             // pulses
    size_t stride;                  // points between consecutive control arrays in the device buffer
    unsigned long long nsteps;      // effective steps per pulse to process; step 0 starts at raw point 0
    unsigned long long total_steps; // steps of the whole pulse (degree policy of sliced runs)
    double dt;
    double series_norm;             // bound of ||H(t_j)||_2 the series is built for; 0: the reference's Hnorm
    bool real_amps;                 // every amplitude of the call has a zero imaginary part (measured with the maxima; dim > 16 only)
};

unsigned long long effective_steps(const Context *c, unsigned long long pts) {   // parament.cpp:820-831
    if (c->enable_magnus || c->quadrature == PARAMENT_QUADRATURE_SIMPSON) return pts >= 1 ? (pts - 1) / 2 : 0;
    if (c->quadrature == PARAMENT_QUADRATURE_MIDPOINT) return pts >= 1 ? pts - 1 : 0;
    return pts;
}
int points_per_step(const Context *c) {
    return (c->enable_magnus || c->quadrature == PARAMENT_QUADRATURE_SIMPSON) ? 2 : 1;
}
int point_overlap(const Context *c) {   // extra raw points a slice needs beyond r * nsteps
    if (c->enable_magnus || c->quadrature == PARAMENT_QUADRATURE_SIMPSON) return 1;
    if (c->quadrature == PARAMENT_QUADRATURE_MIDPOINT) return 1;
    return 0;
}

// Degree policy.  M_ref is what the reference's table selects for ITS norm bound Hnorm (parament.cpp:376-391); it decides
// SELECT_SMALLER_DT and is what Parament_lastStat reports as the reference degree.  The degree actually evaluated:
//   * series built for Hnorm (dim <= 16, or $PARAMENT_NORM=reference): complex128 contexts use the table; complex64 contexts
//     raise the degree until the accumulated truncation error N * 2|J_{M+1}(x)| is below 1e-6 (10 % of the 1e-5 tolerance),
//     because the fp32 table only bounds the error of ONE step (DESIGN.md "Numerics");
//   * series built for the tighter spectral bound Hs < Hnorm (dim > 16): the table is read at Hs * h (same per-step
//     semantics: the table is a function of norm bound x step), then raised until N * 2|J_{M+1}(Hs h)| is below 1e-6
//     (complex64) / 1e-13 (complex128), and never above the reference's own choice for complex128.
Parament_ErrorCode choose_degree(Context *c, double h, unsigned long long total_steps, double Hs, int &M_ref, int &M_used) {
    if (c->MMAX_manual) {
        M_ref = M_used = c->MMAX;
        if (M_used < 1 || M_used > kMaxDegree) return PARAMENT_STATUS_INVALID_VALUE;
        return PARAMENT_STATUS_SUCCESS;
    }
    const double ha = std::fabs(h);   // backward propagation (dt < 0) needs the degree of |dt|; the tables return 3 for x < 0
    M_ref = c->fp64 ? select_cycles_fp64(c->Hnorm, ha) : select_cycles_fp32(c->Hnorm, ha);
    if (M_ref < 3) return PARAMENT_STATUS_SELECT_SMALLER_DT;   // parament.cpp:386-388
    c->MMAX = M_ref;
    const bool tight = Hs > 0.0 && Hs < c->Hnorm;
    const double Hx = tight ? Hs : c->Hnorm;
    M_used = M_ref;
    if (tight) {
        const int mt = c->fp64 ? select_cycles_fp64(Hx, ha) : select_cycles_fp32(Hx, ha);
        if (mt >= 3 && mt < M_used) M_used = mt;
    }
    if (!c->fp64 || tight) {
        const int m64 = select_cycles_fp64(c->Hnorm, ha);
        const int cap = c->fp64 ? M_ref : std::max(M_ref, m64 > 0 ? m64 : kMaxDegree);
        const double budget = c->fp64 ? 1e-13 : 1e-6;
        std::vector<long double> J; long double j0m1;
        bessel_j_table((long double)(Hx * ha), cap + 1, J, j0m1);
        const double N = (double)std::max<unsigned long long>(total_steps, 1);
        while (M_used < cap && N * 2.0 * std::fabs((double)J[M_used + 1]) > budget) ++M_used;
    }
    return PARAMENT_STATUS_SUCCESS;
}

// r_m = c_m / (-i)^m: the series as a REAL polynomial in A = -i X (p.a holds the monomial coefficients c_m on entry).
static void real_coefficients(const SeriesParams &p, int deg, long double *r) {
    for (int m = 0; m <= deg; ++m) {
        const long double re = (long double)p.a[m].re + (long double)p.a_lo[m].re;
        const long double im = (long double)p.a[m].im + (long double)p.a_lo[m].im;
        switch (m & 3) {
            case 0: r[m] = re; break;
            case 1: r[m] = -im; break;
            case 2: r[m] = -re; break;
            default: r[m] = im; break;
        }
    }
}

static void store_split(SeriesParams &p, int k, long double v) {
    p.a[k] = cplx{(double)v, 0.0};
    p.a_lo[k] = cplx{(double)(v - (long double)p.a[k].re), 0.0};
}

// Degree 8 in three matrix products (poly_solve.hpp).  The kernel (k1_warp.cu) folds e0 y02 into the left factor,
//     E = (y02 + d2 A2 + d1 A + e0 I)(y02 + e2 A2) + (r2 - e0 e2) A2 + r1 A + r0 I,
// and gets p.a[0..8].re (+ a_lo) = c4, c3, d2, d1, e2, e0, r2 - e0 e2, r1, r0.  Returns false (keep the Horner form)
// when the system has no real solution.
bool solve_degree8(SeriesParams &p) {
    long double r[9], v[6];
    real_coefficients(p, 8, r);
    if (!solve_degree8_real(r, v)) return false;
    const long double e2 = (long double)(double)v[4], e0 = (long double)(double)v[5];   // as the kernel will see them
    for (int k = 0; k < 6; ++k) store_split(p, k, v[k]);
    store_split(p, 6, r[2] - e0 * e2);
    store_split(p, 7, r[1]);
    store_split(p, 8, r[0]);
    store_split(p, 9, r[2]);   // the full second-order coefficient: the mixed-precision kernel keeps the e0 e2 W term out of its fp32 product
    // constants of that kernel's fp32 part; those that enter the cubic coefficient as hi + lo pairs (a float-rounded constant
    // would be the same relative error in every step)
    auto hi = [](long double x) { return (float)x; };
    auto lo = [](long double x) { return (float)(x - (long double)(float)x); };
    p.fconst[0] = hi(v[0]); p.fconst[1] = hi(v[2]);
    p.fconst[2] = hi(v[1]); p.fconst[3] = lo(v[1]);
    p.fconst[4] = hi(v[3]); p.fconst[5] = lo(v[3]);
    p.fconst[6] = hi(v[4]); p.fconst[7] = lo(v[4]);
    p.fconst[8] = hi(v[5]); p.fconst[9] = lo(v[5]);
    return true;
}

// Degree 12 in four matrix products (poly_solve.hpp), for the shared-memory and batched families.  In terms of
// Y = X, W = X^2, V = X^3 (A = -i X, A2 = -W, A3 = i V) the kernels evaluate
//     T' = tV V + i tW W + tY Y                      (= i (c3 A3 + c2 A2 + c1 A))
//     y0 = T' V
//     L  = y0 + i lV V + lW W + i lY Y + lI I        (= y0 + d3 A3 + d2 A2 + d1 A + f I)
//     R  = y0 + i rV V + rW W                        (= y0 + e3 A3 + e2 A2)
//     E  = L R + i sV V + sW W + i sY Y + sI I       (f y0 is folded into L: f y0 = f R - f e3 A3 - f e2 A2)
// p.a[0..12].re = tV tW tY lV lW lY lI rV rW sV sW sY sI.  The low-order constants sV..sI are derived from the parameters AS
// ROUNDED to double, so that the polynomial the kernel evaluates has the exact r_0 .. r_3; their sub-ulp remainders go to
// p.a_lo (DESIGN.md "Numerics").
enum { S12_TV = 0, S12_TW, S12_TY, S12_LV, S12_LW, S12_LY, S12_LI, S12_RV, S12_RW, S12_SV, S12_SW, S12_SY, S12_SI };
bool solve_degree12(SeriesParams &p) {
    long double r[13], v[9];
    real_coefficients(p, 12, r);
    if (!solve_degree12_real(r, v)) return false;
    long double q[9];
    for (int k = 0; k < 9; ++k) q[k] = (long double)(double)v[k];
    const long double c1 = q[0], c2 = q[1], c3 = q[2], d1 = q[3], d2 = q[4], d3 = q[5], e2 = q[6], e3 = q[7], f = q[8];
    store_split(p, S12_TV, -c3); store_split(p, S12_TW, -c2); store_split(p, S12_TY, c1);
    store_split(p, S12_LV, d3);  store_split(p, S12_LW, -d2); store_split(p, S12_LY, -d1); store_split(p, S12_LI, f);
    store_split(p, S12_RV, e3);  store_split(p, S12_RW, -e2);
    store_split(p, S12_SV, (r[3] - d1 * e2) - f * e3);
    store_split(p, S12_SW, -(r[2] - f * e2));
    store_split(p, S12_SY, -r[1]);
    store_split(p, S12_SI, r[0]);
    return true;
}

// complex64 contexts, dim <= 8, degree-8 form, short pulses: FP32 arithmetic on the TF32 tensor path (k1_tf32.cu) instead of FP64.
// The tensor core accumulates with truncation, which acts as a coherent rescaling of the time axis: the measured error of that
// kernel is beta * (accumulated phase), accumulated phase = N h (s(H0) + sum_k s(H_k)) for amplitudes bounded by 1, with
// beta <= 4e-8 over the sweep of profiles/error_growth_tf32_r2.md.  The path is taken while that stays below half the 1e-5
// tolerance: kTf32MaxPhase = 128 (C5: 1e3 steps x 0.11 = 110).  $PARAMENT_C64_MATH = f64 | tf32 forces either path,
// $PARAMENT_TF32_MAX_PHASE moves the bound (A/B runs and the error sweep).
constexpr double kTf32MaxPhase = 128.0;
bool tf32_candidate(const Context *c, unsigned long long total_steps, double h) {
    if (c->fp64 || c->family != 1 || c->npad != 8) return false;
    const char *em = getenv("PARAMENT_C64_MATH"), *ep = getenv("PARAMENT_TF32_MAX_PHASE");   // read per call: tests toggle them
    const int mode = !em ? 0 : (strcmp(em, "f64") == 0 ? 1 : (strcmp(em, "tf32") == 0 ? 2 : 0));
    const double max_phase = ep ? atof(ep) : kTf32MaxPhase;
    if (mode == 1) return false;
    if (mode == 2) return true;
    if (c->pack > 1) return false;   // dim <= 4: the packed FP64 kernel does 2-4 steps per tile pass, more than 3xTF32 gains (and is exact to 1e-12)
    double rho = 0.0;
    for (double sg : c->sigma_max) rho += sg;
    if (!(rho > 0.0) || rho > c->Hnorm) rho = c->Hnorm;
    return (double)total_steps * std::fabs(h) * rho <= max_phase;
}
// complex64 contexts, dim 9..16, degree-8 form: mixed precision (k1_warp.cu MIXED).  W = X X, all first- and second-order terms
// and the running product stay in FP64; the two products of the series whose results are small run at fp32 grade on the TF32
// tensor path.  Their truncated accumulation is a coherent error of relative size ~1e-8 on terms of size <= 2e-4 per step:
// measured 2.3e-7 at C2 (5e5 steps, accumulated phase N h rho = 4e4), linear in the phase, so the path is
// taken while the phase is below 5e5 (predicted error 2.9e-6); $PARAMENT_K1_MIXED=0 / 1 forces the choice (read per call).
constexpr double kMixedMaxPhase = 5.0e5;
bool mixed_candidate(const Context *c, unsigned long long total_steps, double h) {
    if (c->fp64 || c->family != 1 || c->npad != 16) return false;
    if (const char *e = getenv("PARAMENT_K1_MIXED")) return atoi(e) == 1;
    double rho = 0.0;
    for (double sg : c->sigma_max) rho += sg;
    if (!(rho > 0.0) || rho > c->Hnorm) rho = c->Hnorm;
    return (double)total_steps * std::fabs(h) * rho <= kMixedMaxPhase;
}
bool use_mixed_path(const Context *c, const SeriesParams &p, unsigned long long total_steps, double h) {
    return p.horner == 3 && mixed_candidate(c, total_steps, h);   // the kernel implements the three-product degree-8 form only
}
bool use_tf32_path(const Context *c, const SeriesParams &p, const CallSpec &s) {
    const double h = (c->enable_magnus || c->quadrature == PARAMENT_QUADRATURE_SIMPSON) ? 2.0 * s.dt : s.dt;
    return p.horner == 3 && tf32_candidate(c, s.total_steps, h);   // the kernel implements the three-product degree-8 form only
}

Parament_ErrorCode build_series(Context *c, const CallSpec &s, SeriesParams &p) {
    const double h = (c->enable_magnus || c->quadrature == PARAMENT_QUADRATURE_SIMPSON) ? 2.0 * s.dt : s.dt;   // parament.cpp:800-802
    int M_ref = 0, M_used = 0;
    const double Hs = (s.series_norm > 0.0 && s.series_norm < c->Hnorm) ? s.series_norm : c->Hnorm;   // norm the series is built for
    Parament_ErrorCode ec = choose_degree(c, h, s.total_steps, Hs, M_ref, M_used);
    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    c->stat_series_norm = Hs;
    // Degrees 6..8 are evaluated as ONE degree-8 polynomial in three matrix products (below); the Y^2 Horner form needs four
    // for degree 6 or 7.  The register-resident family does so for complex64 contexts (its path has no compensated constants).
    // (The TF32 and mixed-precision kernels of complex64 contexts implement that form only; degrees 4 and 5 cost three products
    // as well.)
    const bool fp32_grade = tf32_candidate(c, s.total_steps, h) || mixed_candidate(c, s.total_steps, h);
    const bool want_s8 = (c->family != 1 || !c->fp64) && !c->MMAX_manual && c->series_mode == 0 &&
                         M_used >= (fp32_grade ? 4 : 6) && M_used <= 8 && Hs * std::fabs(h) <= 1.0;
    if (want_s8) M_used = 8;
    // degrees 9..12 as ONE degree-12 polynomial in four matrix products
    const bool want_s12 = !c->MMAX_manual && c->series_mode == 0 && M_used >= 9 && M_used <= 12 &&
                          Hs * std::fabs(h) <= 1.0;
    if (want_s12) M_used = 12;
    c->stat_M_ref = M_ref;
    c->stat_M_used = M_used;
    c->stat_horner = 0;
    memset(&p, 0, sizeof(p));
    p.n = c->dim;
    p.npad = c->npad;
    p.pack = c->pack;
    p.herm = (c->hermitian && !c->enable_magnus) ? 1 : 0;
    p.quad = c->enable_magnus ? QUAD_SIMPSON
             : (c->quadrature == PARAMENT_QUADRATURE_SIMPSON ? QUAD_SIMPSON
                : (c->quadrature == PARAMENT_QUADRATURE_MIDPOINT ? QUAD_MIDPOINT : QUAD_NONE));
    p.M = M_used;
    p.pts = (unsigned int)s.stride;
    p.amps_in = s.amps;
    p.magfac = h / 12.0;
    const int cache_flags = (want_s8 ? 1 : 0) | (want_s12 ? 2 : 0) | (c->series_mode << 2) | (c->fp64 ? 32 : 0);
    Context::SeriesCache &sc = c->series_cache;
    const bool cached = sc.valid && sc.h == h && sc.Hnorm == Hs && sc.M == M_used && sc.family == c->family &&
                        sc.onchip == (int)c->onchip && sc.flags == cache_flags;
    if (cached) {
        p.sigma = sc.sigma;
        p.horner = sc.horner;
        memcpy(p.a, sc.a, sizeof(p.a));
        memcpy(p.a_lo, sc.a_lo, sizeof(p.a_lo));
        memcpy(p.fconst, sc.fconst, sizeof(p.fconst));
    } else {
    // sigma is rounded to double FIRST and x is derived from the rounded value in long double, so that
    // sigma * x == 2 h holds to 1e-19: a relative error in sigma alone would stretch the time axis coherently.
    p.sigma = 2.0 / Hs;
    const long double x = 2.0L * (long double)h / (long double)p.sigma;
    std::vector<long double> J; long double j0m1;
    bessel_j_table(x, M_used, J, j0m1);
    if (x < 0)   // backward propagation: J_k(-x) = (-1)^k J_k(x)
        for (int k = 1; k <= M_used; k += 2) J[k] = -J[k];
    p.a[0] = cplx{(double)j0m1, 0.0};
    p.a_lo[0] = cplx{(double)(j0m1 - (long double)p.a[0].re), 0.0};
    for (int k = 1; k <= M_used; ++k) {
        const double v = (double)J[k];
        const double l = (double)(J[k] - (long double)v);   // remainder below the double rounding of J_k
        switch (k & 3) {   // (-i)^k, mathhelper.cpp:35-46
            case 0: p.a[k] = cplx{v, 0.0};  p.a_lo[k] = cplx{l, 0.0}; break;
            case 1: p.a[k] = cplx{0.0, -v}; p.a_lo[k] = cplx{0.0, -l}; break;
            case 2: p.a[k] = cplx{-v, 0.0}; p.a_lo[k] = cplx{-l, 0.0}; break;
            default: p.a[k] = cplx{0.0, v}; p.a_lo[k] = cplx{0.0, l}; break;
        }
    }
    // Horner-in-Y^2 evaluation of the same polynomial (register-resident family only): monomial coefficients
    // c_m = 2^-m sum_k alpha_k t_{k,m}, alpha_0 = J0 - 1, alpha_k = 2 (-i)^k J_k, t_{k,m} the integer coefficients of T_k,
    // converted in long double.  Restricted to x <= 1, where every term of the monomial sum is <= 1 (no cancellation).
    p.horner = 0;
    const double sigma0 = p.sigma;
    if (c->series_mode != 1 && M_used >= 3 && M_used <= 24 && std::fabs((double)x) <= 1.0) {
        p.horner = 1;
        const int d = M_used;
        std::vector<std::vector<long double>> t(d + 1, std::vector<long double>(d + 1, 0.0L));
        t[0][0] = 1.0L;
        if (d >= 1) t[1][1] = 1.0L;
        for (int k = 2; k <= d; ++k)
            for (int m = 0; m <= k; ++m) t[k][m] = (m > 0 ? 2.0L * t[k - 1][m - 1] : 0.0L) - t[k - 2][m];
        long double pow2 = 1.0L;
        for (int m = 0; m <= d + 1; ++m, pow2 *= 2.0L) {
            long double cr = 0.0L, ci = 0.0L;
            for (int k = d; k >= m && m <= d; --k) {          // small terms first
                if (((k - m) & 1) || t[k][m] == 0.0L) continue;
                const long double mag = (k == 0) ? j0m1 : 2.0L * J[k];
                switch (k & 3) {
                    case 0: cr += mag * t[k][m]; break;
                    case 1: ci -= mag * t[k][m]; break;
                    case 2: cr -= mag * t[k][m]; break;
                    default: ci += mag * t[k][m]; break;
                }
            }
            cr /= pow2; ci /= pow2;
            // fold sigma^m in: the kernels then work on the unscaled X = H0 + sum c_t H_t (p.sigma = 1)
            for (int q = 0; q < m; ++q) { cr *= (long double)sigma0; ci *= (long double)sigma0; }
            p.a[m] = cplx{(double)cr, (double)ci};
            p.a_lo[m] = cplx{(double)(cr - (long double)p.a[m].re), (double)(ci - (long double)p.a[m].im)};
        }
        p.sigma = 1.0;
        // Paterson-Stockmeyer blocks of four (3 + floor(M/4) products instead of 1 + floor(M/2)) where the kernels keep their
        // operands in global memory; the register- and shared-memory-resident kernels evaluate the Y^2 form.
        if (M_used >= 10 && (c->family == 3 || (c->family == 2 && !c->onchip))) p.horner = 2;
        if (want_s8 && solve_degree8(p)) p.horner = 3;
        if (want_s12 && solve_degree12(p)) p.horner = 4;
    }
    sc.valid = true; sc.h = h; sc.Hnorm = Hs; sc.M = M_used; sc.family = c->family; sc.onchip = (int)c->onchip; sc.flags = cache_flags;
    sc.horner = p.horner; sc.sigma = p.sigma;
    memcpy(sc.a, p.a, sizeof(p.a));
    memcpy(sc.a_lo, p.a_lo, sizeof(p.a_lo));
    memcpy(sc.fconst, p.fconst, sizeof(p.fconst));
    }
    c->stat_horner = p.horner;
    p.mixed = use_mixed_path(c, p, s.total_steps, h) ? 1 : 0;
    const int A = c->amps, Ain = (int)s.amps;
    int nt = 0;
    const int need = Ain + (c->enable_magnus ? Ain + Ain * (Ain - 1) / 2 : 0);
    if (need > kMaxTerms) return PARAMENT_STATUS_INVALID_VALUE;
    for (int j = 0; j < Ain; ++j) p.terms[nt++] = Term{TERM_PLAIN, 1 + j, j, 0};
    if (c->enable_magnus) {
        for (int j = 0; j < Ain; ++j) p.terms[nt++] = Term{TERM_MAG_DRIFT, 1 + A + j, j, 0};
        for (int j = 0; j < Ain; ++j)
            for (int k = j + 1; k < Ain; ++k) p.terms[nt++] = Term{TERM_MAG_PAIR, 1 + 2 * A + pair_index(j, k, A), j, k};
    }
    p.nterms = nt;
    return PARAMENT_STATUS_SUCCESS;
}

#define PB_LAUNCH(expr)                                         \
    do {                                                        \
        if ((expr) != cudaSuccess) return PARAMENT_STATUS_CUBLAS_FAILED; \
        ++c->stat_launches;                                     \
    } while (0)

// Ordered E-form tree level: dst[i] = src[2i] + src[2i+1] + src[2i+1] * src[2i]; odd leftover copied.
Parament_ErrorCode tree_level(Context *c, const double2 *src, int count, double2 *dst, int npad, cudaStream_t st, int &out_count) {
    const long long nn = (long long)npad * npad;
    const int pairs = count / 2;
    GemmArgs g{};
    g.A = src + nn; g.strideA = 2 * nn;
    g.B = src; g.strideB = 2 * nn;
    g.C[0] = src; g.strideC[0] = 2 * nn; g.beta[0] = cplx{1.0, 0.0};
    g.C2 = src + nn; g.strideC2 = 2 * nn; g.beta2 = 1.0;
    g.D = dst; g.strideD = nn;
    g.n = npad; g.batch = pairs;
    if (pairs > 0) PB_LAUNCH(k4_gemm(g, st));
    if (count & 1) {
        if (!PB_CUDA_OK(cudaMemcpyAsync(dst + (long long)pairs * nn, src + (long long)(count - 1) * nn, nn * sizeof(double2),
                                        cudaMemcpyDeviceToDevice, st)))
            return PARAMENT_STATUS_CUBLAS_FAILED;
    }
    out_count = pairs + (count & 1);
    return PARAMENT_STATUS_SUCCESS;
}

// Reduce `count` matrices at `buf` to one, ping-ponging with `scratch`; the result ends in buf[0].
Parament_ErrorCode tree_reduce_all(Context *c, double2 *buf, int count, double2 *scratch, int npad, cudaStream_t st) {
    const size_t nn = (size_t)npad * npad;
    double2 *src = buf, *dst = scratch;
    while (count > 1) {
        int nc = 0;
        Parament_ErrorCode ec = tree_level(c, src, count, dst, npad, st, nc);
        if (ec != PARAMENT_STATUS_SUCCESS) return ec;
        count = nc;
        std::swap(src, dst);
    }
    if (src != buf && !PB_CUDA_OK(cudaMemcpyAsync(buf, src, nn * sizeof(double2), cudaMemcpyDeviceToDevice, st)))
        return PARAMENT_STATUS_CUBLAS_FAILED;
    return PARAMENT_STATUS_SUCCESS;
}

struct F3Plan { int S; int cap; int NS; };   // chunk length, pending-list capacity, chunk streams (work sets) in flight
constexpr int kF3Streams = 4;
int f3_stream_count() {   // chunks in flight (one stream and one work set each); $PARAMENT_F3_STREAMS = 1..4 for A/B runs
    const char *env = getenv("PARAMENT_F3_STREAMS");
    const int v = env ? atoi(env) : 4;   // measured at dim 256: 1 -> 3.82e4, 2 -> 4.21e4, 3 -> 4.24e4, 4 -> 4.28e4 steps/s
    return v < 1 ? 1 : (v > kF3Streams ? kF3Streams : v);
}

// Chunk length S: a whole number of GEMM waves (S * tiles == k * co-resident CTAs: a launch that spills a few CTAs
// into an extra wave costs a full wave) with the six S x npad^2 work arrays (Y, Y^2, Y^3, Y^4, two recurrence
// registers; <= ~110 MB, mostly L2-resident).  `cap` bounds the buffer of pending partial products.
F3Plan plan_family3(const Context *c, const CallSpec &s) {
    const size_t nn = (size_t)c->npad * c->npad;
    const int tiles = k4_tiles(c->npad);
    const int slots = c->k4_slots > 0 ? c->k4_slots : 2 * c->num_sms;
    F3Plan f;
    const long long per_wave = std::max<long long>(1, slots / tiles);
    const long long budget = std::max<long long>(1, (long long)(110e6 / ((double)kSeriesSlots * nn * sizeof(double2))));
    long long waves = std::max<long long>(1, budget / per_wave);
    if (waves > 4) waves = 4;
    f.S = (int)std::max<long long>(2, std::min<long long>(per_wave * waves, std::max<long long>(budget, per_wave)));
    if ((unsigned long long)f.S > s.nsteps) f.S = (int)std::max<unsigned long long>(s.nsteps, 1);
    // pending partial products: enough for the first tree levels to run in whole waves (16 chunks), bounded by 1.5 GiB
    const size_t want = std::min<size_t>(std::max<size_t>(64, (size_t)16 * f.S), std::max<size_t>(64, ((size_t)3 << 29) / (nn * sizeof(double2))));
    f.cap = (int)std::max<size_t>(2, std::min<size_t>(want, (size_t)std::min<unsigned long long>(s.nsteps, 1ull << 30)));
    f.NS = (s.nsteps >= 4ull * f.S) ? f3_stream_count() : 1;   // short calls run on one stream and get one work set
    return f;
}

int chain_grid(const Context *c, const CallSpec &s) {
    return (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)c->k4_slots, s.nsteps));
}

bool alloc_family3(Context *c, const F3Plan &f) {
    const size_t nn = (size_t)c->npad * c->npad;
    return ensure_dev(c->d_Y, (size_t)f.NS * kSeriesSlots * f.S * nn * sizeof(double2)) &&   // one chunk work set per stream in use
           ensure_dev(c->d_pending, (size_t)2 * (f.cap + f.S) * nn * sizeof(double2)) &&   // two halves
           ensure_dev(c->d_tree, (size_t)((f.cap + f.S) / 2 + 1) * nn * sizeof(double2));
}

bool alloc_family2(Context *c, int grid) {
    const size_t nn = (size_t)c->npad * c->npad;
    return ensure_dev(c->d_Y, (size_t)grid * (kSeriesSlots + 2) * nn * sizeof(double2)) &&   // per-CTA scratch
           ensure_dev(c->d_pending, (size_t)grid * nn * sizeof(double2)) &&
           ensure_dev(c->d_tree, (size_t)(grid / 2 + 1) * nn * sizeof(double2));
}

// dim 17..64: one persistent launch per pulse + the ordered reduction of the per-CTA partials.
Parament_ErrorCode run_family2(Context *c, const SeriesParams &p, const void *carr_dev, const CallSpec &s, void *out_dev, cudaStream_t st) {
    const int np = c->npad;
    const size_t io = c->fp64 ? sizeof(double2) : sizeof(float2);
    const SeriesProgram prog = build_program(p);
    const int grid = chain_grid(c, s);
    double2 *pend = (double2 *)c->d_pending.ptr, *tree = (double2 *)c->d_tree.ptr;
    for (unsigned int b = 0; b < s.batch; ++b) {
        const char *cb = (const char *)carr_dev + (size_t)b * s.amps * s.stride * io;
        if (c->onchip)
            PB_LAUNCH(k4_onchip(c->fp64, p, prog, cb, (const double2 *)c->d_H.ptr, (double2 *)c->d_Y.ptr, pend, s.nsteps, grid, st));
        else
            PB_LAUNCH(k4_chain(c->fp64, p, prog, cb, (const double2 *)c->d_H.ptr, (double2 *)c->d_Y.ptr, pend, s.nsteps, grid, st));
        Parament_ErrorCode ec = tree_reduce_all(c, pend, grid, tree, np, st);
        if (ec != PARAMENT_STATUS_SUCCESS) return ec;
        PB_LAUNCH(k4_finish(c->fp64, pend, c->dim, np, (char *)out_dev + (size_t)b * c->dim * c->dim * io, true, st));
    }
    return PARAMENT_STATUS_SUCCESS;
}

// dim > 64: time chunks of one wave of CTAs, every series op one batched launch.  Consecutive chunks rotate over NS
// work sets on NS streams: a launch is exactly one wave, so on one stream the SMs drain at the end of every launch and
// refill at the start of the next (measured ~8 % of the time at dim 256); with other, independent chunks in flight the
// CTAs of their launches take the slots as they free up.  The ordered product only needs the chunks' results in the
// pending buffer in chunk order, which the host-side slot assignment fixes.  The pending buffer has two halves: while
// the reduction of a full half runs on the caller's stream (its last levels are far smaller than a wave), the chunk
// streams already fill the other half, whose slot 0 receives the reduced product as carry.
// $PARAMENT_F3_STREAMS=1 keeps everything on one stream (A/B).
Parament_ErrorCode run_family3(Context *c, const SeriesParams &p, const void *carr_dev, const CallSpec &s, void *out_dev, cudaStream_t st) {
    const int np = c->npad;
    const size_t nn = (size_t)np * np;
    const size_t io = c->fp64 ? sizeof(double2) : sizeof(float2);
    const int tiles = k4_tiles(np);
    const F3Plan f = plan_family3(c, s);
    const int S = f.S, cap = f.cap;
    const SeriesProgram prog = build_program(p);
    const int NS = f.NS;
    // Hermitian H0 / H_k, no Magnus terms and real amplitudes throughout the call (measured with the amplitude maxima): every Y is
    // Hermitian, so Y Y is too and only its upper-triangular tiles are computed (k4_gemm.hpp GemmArgs::herm); $PARAMENT_K4_HERM=0: A/B
    const bool herm_steps = c->hermitian && !c->enable_magnus && s.real_amps && !(getenv("PARAMENT_K4_HERM") && atoi(getenv("PARAMENT_K4_HERM")) == 0);
    c->stat_products_saved = 0.0;
    if (herm_steps)
        for (int o = 0; o < prog.nops; ++o)
            if (prog.ops[o].A == prog.ops[o].B && prog.ops[o].A <= 1) c->stat_products_saved += 1.0 - (double)k4_herm_tiles(np) / (double)tiles;
    cudaStream_t sx[kF3Streams] = {st, nullptr, nullptr, nullptr};
    if (NS > 1) {
        sx[0] = c->copy_stream;
        for (int k = 1; k < NS; ++k) {
            if (!c->f3_streams[k - 1] && !PB_CUDA_OK(cudaStreamCreateWithFlags(&c->f3_streams[k - 1], cudaStreamNonBlocking)))
                return PARAMENT_STATUS_CUBLAS_FAILED;
            sx[k] = c->f3_streams[k - 1];
        }
    }
    cudaEvent_t ev_main = c->ev_copy[0], *ev_chunk = c->ev_copy + 1, *ev_red = c->ev_copy + 1 + kF3Streams;
    double2 *slots[kF3Streams][kSeriesSlots];
    for (int k = 0; k < NS; ++k)
        for (int i = 0; i < kSeriesSlots; ++i) slots[k][i] = (double2 *)c->d_Y.ptr + ((size_t)k * kSeriesSlots + i) * S * nn;
    double2 *halves[2] = {(double2 *)c->d_pending.ptr, (double2 *)c->d_pending.ptr + (size_t)(cap + S) * nn};
    double2 *tree = (double2 *)c->d_tree.ptr;
    // the chunk streams start after everything already enqueued on the caller's stream (the H2D of the amplitudes, earlier pulses)
    auto chunks_after_main = [&]() {
        if (NS == 1) return true;
        if (!PB_CUDA_OK(cudaEventRecord(ev_main, st))) return false;
        for (int k = 0; k < NS; ++k)
            if (!PB_CUDA_OK(cudaStreamWaitEvent(sx[k], ev_main, 0))) return false;
        return true;
    };
    auto main_after_chunks = [&]() {
        if (NS == 1) return true;
        for (int k = 0; k < NS; ++k)
            if (!PB_CUDA_OK(cudaEventRecord(ev_chunk[k], sx[k])) || !PB_CUDA_OK(cudaStreamWaitEvent(st, ev_chunk[k], 0))) return false;
        return true;
    };
    for (unsigned int b = 0; b < s.batch; ++b) {
        const char *cb = (const char *)carr_dev + (size_t)b * s.amps * s.stride * io;
        if (!chunks_after_main()) return PARAMENT_STATUS_CUBLAS_FAILED;
        bool wait_red[2][kF3Streams] = {};   // stream k must see the last reduction of that half finished before writing into it
        int half = 0, pending = 0;
        unsigned long long chunk = 0;
        for (unsigned long long step0 = 0; step0 < s.nsteps; step0 += S, ++chunk) {
            const int Sc = (int)std::min<unsigned long long>(S, s.nsteps - step0);
            const int k = (int)(chunk % NS);
            cudaStream_t q = sx[k];
            double2 **sl = slots[k];
            double2 *pend = halves[half];
            if (wait_red[half][k]) {
                if (!PB_CUDA_OK(cudaStreamWaitEvent(q, ev_red[half], 0))) return PARAMENT_STATUS_CUBLAS_FAILED;
                wait_red[half][k] = false;
            }
            // no tree level inside the chunk (a level of a one-wave chunk does not fill a wave): the last series op writes
            // its result where the ordered product will read it, instead of a device-to-device copy of the chunk
            const bool direct = prog.nops > 0 && prog.ops[prog.nops - 1].D == prog.e_slot && !(Sc > 1 && (Sc / 2) * tiles >= c->k4_slots);
            PB_LAUNCH(k4_assemble(c->fp64, p, prog, cb, (const double2 *)c->d_H.ptr, sl[0], sl[4], sl[5], step0, Sc, q));
            for (int o = 0; o < prog.nops; ++o) {
                const SeriesOp &op = prog.ops[o];
                GemmArgs g{};
                g.A = sl[op.A]; g.B = sl[op.B]; g.D = sl[op.D];
                g.Dprod = op.Dprod >= 0 ? sl[op.Dprod] : nullptr;
                g.Dalt = op.Dalt >= 0 ? sl[op.Dalt] : nullptr;
                g.strideA = g.strideB = g.strideD = g.strideDprod = g.strideDalt = (long long)nn;
                for (int j = 0; j < kMaxAddends; ++j) {
                    g.C[j] = op.C[j] >= 0 ? sl[op.C[j]] : nullptr; g.strideC[j] = (long long)nn;
                    g.beta[j] = op.beta[j]; g.beta_lo[j] = op.beta_lo[j]; g.beta_alt[j] = op.beta_alt[j];
                }
                g.alpha = op.alpha; g.scaled = op.scaled; g.gamma = op.gamma; g.gamma_lo = op.gamma_lo;
                g.C2 = nullptr; g.beta2 = 0.0; g.n = np; g.batch = Sc;
                g.herm = (herm_steps && op.A == op.B && op.A <= 1) ? 1 : 0;   // Y Y and W W: Hermitian like their addends (powers of Y)
                if (direct && o == prog.nops - 1) g.D = pend + (size_t)pending * nn;   // E goes straight to the pending buffer
                PB_LAUNCH(k4_gemm(g, q));
            }
            if (direct) {
                pending += Sc;
            } else {
                // ordered product inside the chunk only while a level still fills a whole wave of CTAs; smaller levels are
                // deferred to the pending buffer, whose reduction runs on hundreds of matrices at a time (full waves)
                double2 *src = sl[prog.e_slot];
                double2 *other = sl[prog.e_slot == 4 ? 5 : 4];
                int count = Sc;
                while (count > 1 && (count / 2) * tiles >= c->k4_slots) {
                    int nc = 0;
                    Parament_ErrorCode ec = tree_level(c, src, count, other, np, q, nc);
                    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
                    std::swap(src, other);
                    count = nc;
                }
                if (!PB_CUDA_OK(cudaMemcpyAsync(pend + (size_t)pending * nn, src, (size_t)count * nn * sizeof(double2), cudaMemcpyDeviceToDevice, q)))
                    return PARAMENT_STATUS_CUBLAS_FAILED;
                pending += count;
            }
            if (pending >= cap && step0 + S < s.nsteps) {
                // this half is full: reduce it on the caller's stream, carry the product into slot 0 of the other half
                if (!main_after_chunks()) return PARAMENT_STATUS_CUBLAS_FAILED;
                Parament_ErrorCode ec = tree_reduce_all(c, pend, pending, tree, np, st);
                if (ec != PARAMENT_STATUS_SUCCESS) return ec;
                if (!PB_CUDA_OK(cudaMemcpyAsync(halves[half ^ 1], pend, nn * sizeof(double2), cudaMemcpyDeviceToDevice, st)))
                    return PARAMENT_STATUS_CUBLAS_FAILED;
                if (NS > 1) {
                    if (!PB_CUDA_OK(cudaEventRecord(ev_red[half], st))) return PARAMENT_STATUS_CUBLAS_FAILED;
                    for (int j = 0; j < NS; ++j) wait_red[half][j] = true;
                }
                half ^= 1;
                pending = 1;
            }
        }
        if (!main_after_chunks()) return PARAMENT_STATUS_CUBLAS_FAILED;
        Parament_ErrorCode ec = tree_reduce_all(c, halves[half], pending, tree, np, st);
        if (ec != PARAMENT_STATUS_SUCCESS) return ec;
        PB_LAUNCH(k4_finish(c->fp64, halves[half], c->dim, np, (char *)out_dev + (size_t)b * c->dim * c->dim * io, true, st));
    }
    return PARAMENT_STATUS_SUCCESS;
}

// Chain launch of the register-resident family in either arithmetic.
cudaError_t launch_family1_chain(const Context *c, const SeriesParams &p, bool tf32, const void *carr, double2 *partials, unsigned int batch,
                                 const K1Plan &plan, unsigned long long lo, unsigned long long hi, const K1Final &fz, cudaStream_t st) {
    if (tf32) return launch_k1_tf32_chain(p, carr, (const double2 *)c->d_H.ptr, partials, batch, plan, lo, hi, fz, st);
    return launch_k1_chain(c->npad, c->fp64, p, carr, (const double2 *)c->d_H.ptr, partials, batch, plan, lo, hi, fz, st);
}

// Scratch of a fused chain launch (plan_k1 with fuse = true): the CTA partials, the group products behind them, and the arrival
// counters (zeroed when the buffer grows; every launch leaves them at zero).  Returns the final-stage arguments.
bool prepare_fused(Context *c, const K1Plan &plan, unsigned int batch, void *out_dev, cudaStream_t st, K1Final &fz) {
    const size_t np2 = (size_t)c->npad * c->npad;
    const size_t mid_elems = (size_t)batch * plan.groups_per_pulse * np2;
    const size_t cnt_bytes = ((size_t)batch * (plan.groups_per_pulse + 1) + 1) * sizeof(unsigned int);
    if (!ensure_dev(c->d_partials, (plan.partial_elems + mid_elems) * sizeof(double2))) return false;
    if (c->d_counters.bytes < cnt_bytes || !c->d_counters.ptr) {
        if (!ensure_dev(c->d_counters, cnt_bytes) || !PB_CUDA_OK(cudaMemsetAsync(c->d_counters.ptr, 0, c->d_counters.bytes, st))) return false;
    }
    fz.out = out_dev; fz.n = c->dim; fz.counters = (unsigned int *)c->d_counters.ptr;
    fz.mid = (double2 *)c->d_partials.ptr + plan.partial_elems; fz.groups = plan.groups_per_pulse;
    return true;
}

// Spectral bound of the step Hamiltonians of THIS call:  ||H0 + sum_k c_k(t) H_k||_2 <= s_0 + sum_k max_t|c_k(t)| s_k, with s_m the
// largest singular values from setHamiltonian and the amplitude maxima measured on the device (one pass over the amplitude stream
// and a 64-byte read-back; dim > 16 only, where a step costs >= 50 us of tensor work per SM).  The quadrature averages cannot
// exceed the maxima; the Magnus commutator term adds at most (h/12) 2 rho^2.  5 % margin on the power-iteration estimates; never
// above the reference's bound Hnorm (parament.cpp:280-284), which stays in charge of the error semantics.
Parament_ErrorCode series_norm_for_call(Context *c, const void *carr_dev, const CallSpec &s, cudaStream_t st, double &Hs, bool &real_amps) {
    Hs = c->Hnorm;
    real_amps = false;
    if (c->family == 1 || c->norm_mode != 1 || c->MMAX_manual || c->sigma_max.size() < (size_t)c->amps + 1) return PARAMENT_STATUS_SUCCESS;
    double rho = c->sigma_max[0];
    if (s.amps > 0) {
        if (!ensure_dev(c->d_absmax, (kMaxTerms + 1) * sizeof(unsigned long long))) return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
        const size_t pts = (size_t)points_per_step(c) * s.nsteps + point_overlap(c);
        if (k4_absmax(c->fp64, carr_dev, s.batch, s.amps, s.stride, std::min(pts, s.stride), (unsigned long long *)c->d_absmax.ptr, st) != cudaSuccess ||
            !PB_CUDA_OK(cudaMemcpyAsync(c->h_absmax, c->d_absmax.ptr, (s.amps + 1) * sizeof(double), cudaMemcpyDeviceToHost, st)) ||
            !PB_CUDA_OK(cudaStreamSynchronize(st)))
            return PARAMENT_STATUS_CUBLAS_FAILED;
        for (unsigned int k = 0; k < s.amps; ++k) rho += std::sqrt(c->h_absmax[k]) * c->sigma_max[1 + k];
        unsigned long long flag;
        memcpy(&flag, &c->h_absmax[s.amps], sizeof(flag));
        real_amps = flag == 0;
    }
    if (c->enable_magnus) {
        const double h = 2.0 * std::fabs(s.dt);
        rho += (h / 6.0) * rho * rho;
    }
    rho *= 1.05;
    if (rho > 0.0 && rho < c->Hnorm) Hs = rho;
    return PARAMENT_STATUS_SUCCESS;
}

// Device-resident core shared by every equiprop entry point.
Parament_ErrorCode propagate_device(Context *c, const void *carr_dev, const CallSpec &s_in, void *out_dev, cudaStream_t st) {
    NvtxRange range("parament: propagate (device-resident core)");
    CallSpec s = s_in;
    Parament_ErrorCode ec = series_norm_for_call(c, carr_dev, s, st, s.series_norm, s.real_amps);
    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    SeriesParams p;
    ec = build_series(c, s, p);
    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    c->stat_steps = s.nsteps;
    c->stat_launches = 0;
    c->stat_products_saved = 0.0;
    // all allocations happen before the timed region (grow-only scratch, nothing is allocated in steady state)
    K1Plan plan{};
    const bool tf32 = use_tf32_path(c, p, s);
    c->stat_math = tf32 ? 1 : (p.mixed ? 2 : 0);
    K1Final fz{};
    if (c->family == 1) {
        plan = plan_k1(c->npad, s.batch, s.nsteps, c->num_sms, p.horner != 0, true);
        if (!prepare_fused(c, plan, s.batch, out_dev, st, fz)) return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
    } else if (c->family == 2) {
        if (!alloc_family2(c, chain_grid(c, s))) return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
    } else if (!alloc_family3(c, plan_family3(c, s))) {
        return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
    }
    if (!PB_CUDA_OK(cudaEventRecord(c->ev_start, st))) return PARAMENT_STATUS_CUBLAS_FAILED;
    if (c->family == 1) {
        // ONE launch: the chain kernel's last CTAs reduce the partials and write the propagators (k1_common.cuh)
        PB_LAUNCH(launch_family1_chain(c, p, tf32, carr_dev, (double2 *)c->d_partials.ptr, s.batch, plan, 0, s.nsteps, fz, st));
    } else {
        ec = c->family == 2 ? run_family2(c, p, carr_dev, s, out_dev, st) : run_family3(c, p, carr_dev, s, out_dev, st);
        if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    }
    if (!PB_CUDA_OK(cudaEventRecord(c->ev_stop, st))) return PARAMENT_STATUS_CUBLAS_FAILED;
    return PARAMENT_STATUS_SUCCESS;
}

// rows x width bytes from caller memory (row pitch spitch) to device memory (row pitch dpitch), asynchronous on `stream`.
// Large transfers from PAGEABLE memory go through the context's staging threads (context.hpp Stager); page-locked or registered
// caller buffers, small transfers, and $PARAMENT_STAGE_THREADS=0 use the plain asynchronous copy.
bool h2d_rows(Context *c, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, cudaStream_t stream) {
    const size_t total = width * rows;
    static const size_t min_bytes = getenv("PARAMENT_STAGE_MIN_KB") ? (size_t)atoi(getenv("PARAMENT_STAGE_MIN_KB")) << 10 : (size_t)256 << 10;
    bool staged = total >= min_bytes && !c->stager_failed;
    if (staged) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, src) != cudaSuccess) { cudaGetLastError(); staged = false; }
        else staged = attr.type == cudaMemoryTypeUnregistered;
    }
    if (staged && !c->stager) {
        static const int env_threads = getenv("PARAMENT_STAGE_THREADS") ? atoi(getenv("PARAMENT_STAGE_THREADS")) : -1;
        const int hw = (int)std::thread::hardware_concurrency();
        const int want = env_threads >= 0 ? env_threads : std::max(1, std::min(6, hw / 3));   // total copying threads, the caller included
        if (want == 0) { c->stager_failed = true; staged = false; }
        else {
            c->stager = new (std::nothrow) Stager();
            if (!c->stager || !c->stager->start(want - 1)) {
                if (c->stager) { c->stager->stop(); delete c->stager; c->stager = nullptr; }
                c->stager_failed = true;
                staged = false;
            }
        }
    }
    if (staged) {
        if (rows == 1 || (dpitch == width && spitch == width)) return c->stager->copy(dst, src, total, stream);
        for (size_t r = 0; r < rows; ++r)
            if (!c->stager->copy((char *)dst + r * dpitch, (const char *)src + r * spitch, width, stream)) return false;
        return true;
    }
    if (rows == 1 || (dpitch == width && spitch == width)) return PB_CUDA_OK(cudaMemcpyAsync(dst, src, total, cudaMemcpyHostToDevice, stream));
    return PB_CUDA_OK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, stream));
}

// Host-pointer path of the register-resident family with the H2D copy of the amplitude stream overlapped with the
// kernels: the work is cut into G groups (pulse ranges of an ensemble, time ranges of a single pulse); group g+1 is
// copied on the copy stream while group g is propagated.  Replaces the reference's one blocking cudaMemcpy of the
// whole array before any work starts (parament.cpp:477).
template <typename T>
Parament_ErrorCode pipelined_family1(Context *c, const T *carr, unsigned int pts, size_t p_lo, size_t seg, const CallSpec &s,
                                     int G, void *out_dev) {
    NvtxRange range("parament: copy / compute pipeline (dim <= 16)");
    SeriesParams p;
    Parament_ErrorCode ec = build_series(c, s, p);
    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    c->stat_steps = s.nsteps;
    c->stat_launches = 0;
    const int n = c->dim, NP2 = c->npad * c->npad;
    T *dcarr = (T *)c->d_carr.ptr;
    const bool horner = p.horner != 0;
    const bool tf32 = use_tf32_path(c, p, s);
    c->stat_math = tf32 ? 1 : (p.mixed ? 2 : 0);
    if (G > 8) G = 8;
    auto copy_arrays = [&](size_t a0, size_t a1, size_t pt0, size_t npts) -> bool {
        // arrays [a0, a1) of the device buffer (stride seg) <- host arrays (stride pts), points [pt0, pt0 + npts) of the slice
        if (seg == pts && npts == seg)
            return h2d_rows(c, dcarr + a0 * seg, 0, carr + a0 * pts, 0, (a1 - a0) * seg * sizeof(T), 1, c->copy_stream);
        // strided: rows = control arrays
        return h2d_rows(c, dcarr + a0 * seg + pt0, seg * sizeof(T), carr + a0 * pts + p_lo + pt0, (size_t)pts * sizeof(T), npts * sizeof(T), a1 - a0,
                        c->copy_stream);
    };
    if (s.batch > 1) {
        // Pulse groups sized in units of 1/8 wave of warps (a warp owns a pulse or 1/k of it, k <= 8: plan_k1), so that
        // every group's launch is a whole number of equally long waves; sizes grow 1, 2, 4, ... units: the first copy
        // is the only one no kernel hides, later groups are long enough to hide theirs behind the group before.
        unsigned int gb[9];
        const int ng = ensemble_copy_groups(s.batch, std::max(1u, k1_warp_slots(c->npad, c->num_sms, horner) / 8), G, gb);   // plan.hpp
        // every group is one fused launch that writes its pulses' propagators; scratch sized for the largest group up front
        {
            K1Plan big{};
            size_t need = 0;
            unsigned int bmax = 0;
            for (int g = 0; g < ng; ++g) {
                const K1Plan pg = plan_k1(c->npad, gb[g + 1] - gb[g], s.nsteps, c->num_sms, horner, true);
                const size_t e = pg.partial_elems + (size_t)(gb[g + 1] - gb[g]) * pg.groups_per_pulse * NP2;
                if (e >= need) { need = e; big = pg; bmax = gb[g + 1] - gb[g]; }
            }
            K1Final probe{};
            if (!ensure_dev(c->d_partials, need * sizeof(double2)) || !prepare_fused(c, big, std::max(bmax, s.batch), out_dev, c->stream, probe))
                return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
        }
        if (!PB_CUDA_OK(cudaEventRecord(c->ev_start, c->stream))) return PARAMENT_STATUS_CUBLAS_FAILED;
        for (int g = 0; g < ng; ++g) {
            const unsigned int b0 = gb[g], b1 = gb[g + 1];
            if (!copy_arrays((size_t)b0 * s.amps, (size_t)b1 * s.amps, 0, seg) ||
                !PB_CUDA_OK(cudaEventRecord(c->ev_copy[g], c->copy_stream)) || !PB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copy[g], 0)))
                return PARAMENT_STATUS_CUBLAS_FAILED;
            const K1Plan plan = plan_k1(c->npad, b1 - b0, s.nsteps, c->num_sms, horner, true);
            K1Final fz{};
            if (!prepare_fused(c, plan, b1 - b0, (T *)out_dev + (size_t)b0 * n * n, c->stream, fz)) return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
            PB_LAUNCH(launch_family1_chain(c, p, tf32, dcarr + (size_t)b0 * s.amps * seg, (double2 *)c->d_partials.ptr, b1 - b0, plan, 0,
                                           s.nsteps, fz, c->stream));
        }
    } else {
        const int r = points_per_step(c), ov = point_overlap(c);
        unsigned long long bound[9];
        K1Plan plans[8];
        size_t off[9];
        off[0] = 0;
        time_copy_groups(s.nsteps, G, bound);   // plan.hpp: a short first group, its copy is the only one no kernel hides
        for (int g = 0; g < G; ++g) {
            plans[g] = plan_k1(c->npad, 1, bound[g + 1] - bound[g], c->num_sms, horner);
            off[g + 1] = off[g] + plans[g].partials_per_pulse;
        }
        if (!ensure_dev(c->d_partials, (off[G] * NP2 + k3_mid_elems(c->npad, 1, (unsigned int)off[G])) * sizeof(double2)))
            return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
        if (!PB_CUDA_OK(cudaEventRecord(c->ev_start, c->stream))) return PARAMENT_STATUS_CUBLAS_FAILED;
        for (int g = 0; g < G; ++g) {
            const size_t pt0 = (size_t)r * bound[g] + (g == 0 ? 0 : ov);            // the overlap point came with the previous group
            const size_t pt1 = (size_t)r * bound[g + 1] + ov;
            if (!copy_arrays(0, s.amps, pt0, pt1 - pt0) ||
                !PB_CUDA_OK(cudaEventRecord(c->ev_copy[g], c->copy_stream)) || !PB_CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_copy[g], 0)))
                return PARAMENT_STATUS_CUBLAS_FAILED;
            PB_LAUNCH(launch_family1_chain(c, p, tf32, dcarr, (double2 *)c->d_partials.ptr + off[g] * NP2, 1, plans[g], bound[g], bound[g + 1],
                                           K1Final{}, c->stream));   // partials only: the groups' partials are reduced together below
        }
        PB_LAUNCH(launch_k3_reduce(c->npad, c->fp64, (const double2 *)c->d_partials.ptr, (unsigned int)off[G], n, out_dev, 1,
                                   (double2 *)c->d_partials.ptr + off[G] * NP2, c->stream));
        c->stat_launches += k3_launches((unsigned int)off[G]) - 1;
    }
    if (!PB_CUDA_OK(cudaEventRecord(c->ev_stop, c->stream))) return PARAMENT_STATUS_CUBLAS_FAILED;
    return PARAMENT_STATUS_SUCCESS;
}

template <typename T>
void write_identity(T *out, int n, unsigned int batch) {
    for (unsigned int b = 0; b < batch; ++b)
        for (int r = 0; r < n; ++r)
            for (int col = 0; col < n; ++col) {
                out[((size_t)b * n + r) * n + col].re = (r == col) ? 1 : 0;
                out[((size_t)b * n + r) * n + col].im = 0;
            }
}

// Identity propagators to a host array or (out_is_device) to device memory through the context's stream.
template <typename T>
Parament_ErrorCode deliver_identity(Context *c, T *out, int n, unsigned int batch, bool out_is_device) {
    if (!out_is_device) { write_identity(out, n, batch); return PARAMENT_STATUS_SUCCESS; }
    try {
        std::vector<T> id((size_t)batch * n * n);
        write_identity(id.data(), n, batch);
        if (!PB_CUDA_OK(cudaMemcpyAsync(out, id.data(), id.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream)) ||
            !PB_CUDA_OK(cudaStreamSynchronize(c->stream)))
            return PARAMENT_STATUS_CUBLAS_FAILED;
    } catch (const std::bad_alloc &) {
        return PARAMENT_STATUS_HOST_ALLOC_FAILED;
    }
    return PARAMENT_STATUS_SUCCESS;
}

// Host-pointer entry: stage the needed part of the amplitude stream, run, copy the result back.
// step range [lo, hi) of each pulse; carr holds batch * amps arrays of pts points.
template <typename T>
Parament_ErrorCode equiprop_multi(Context *c, const T *carr, double dt, unsigned int pts, unsigned int amps, unsigned int batch, T *out,
                                  bool &handled);

// gather_to != nullptr (single-process multi-GPU): the result is not copied to the host but to slot `slot` of
// gather_to->d_gather by a peer copy (NVLink when the devices have peer access), and `out` is not written.
// out_is_device: `out` is device memory on the context's device; the last kernel writes the result there (no D2H), and the
// call returns once it is complete (Parament_equipropSliceToDevice: the partial stays on the GPU for the NCCL exchange).
template <typename T>
Parament_ErrorCode equiprop_host(Context *c, const T *carr, double dt, unsigned int pts, unsigned int amps, unsigned int batch,
                                 unsigned long long lo, unsigned long long hi, bool whole, T *out, Context *gather_to = nullptr,
                                 unsigned int slot = 0, bool out_is_device = false) {
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    NvtxRange range(gather_to ? "parament: equiprop slice (helper device)" : "Parament_equiprop (host pointers)");
    if (whole && !gather_to && !c->peers.empty()) {
        bool handled = false;
        const Parament_ErrorCode mec = equiprop_multi<T>(c, carr, dt, pts, amps, batch, out, handled);
        if (handled) return mec;
    }
    DeviceGuard guard(c->device);
    c->stat_devices = 1;
    if (!c->have_hamiltonian) return fail(c, PARAMENT_STATUS_NO_HAMILTONIAN);   // parament.cpp:795-798
    if (!out || (!carr && amps > 0 && pts > 0) || (int)amps > c->amps || batch == 0) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    const unsigned long long N = effective_steps(c, pts);
    if (whole) { lo = 0; hi = N; }
    if (hi > N || lo > hi) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    const int n = c->dim;
    c->stat_ms = 0; c->stat_launches = 0; c->stat_h2d = 0; c->stat_d2h = 0; c->stat_steps = hi - lo;

    CallSpec s{};
    s.amps = amps; s.batch = batch; s.dt = dt; s.nsteps = hi - lo; s.total_steps = N;
    {   // the degree / dt check happens before any transfer, as in the reference (parament.cpp:804-807); it builds no series
        // (an all-zero Hamiltonian would give sigma = 2 / 0), so nothing is cached from here
        int M_ref = 0, M_used = 0;
        const double h = (c->enable_magnus || c->quadrature == PARAMENT_QUADRATURE_SIMPSON) ? 2.0 * dt : dt;
        Parament_ErrorCode ec = choose_degree(c, h, s.total_steps, c->Hnorm, M_ref, M_used);
        if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
    }
    if (s.nsteps == 0 || c->Hnorm == 0.0) {   // nothing to propagate: identity (reference returns stale memory, SURVEY A-7)
        if (Parament_ErrorCode ec = deliver_identity<T>(c, out, n, batch, out_is_device)) return fail(c, ec);
        c->lastError = PARAMENT_STATUS_SUCCESS;
        return PARAMENT_STATUS_SUCCESS;
    }
    const int r = points_per_step(c);
    const size_t p_lo = (size_t)r * lo;
    const size_t seg = (size_t)r * s.nsteps + point_overlap(c);   // raw points needed per control array
    s.stride = seg;
    const size_t arrays = (size_t)batch * amps;
    const size_t in_bytes = arrays * seg * sizeof(T);
    const size_t out_bytes = (size_t)batch * n * n * sizeof(T);
    if (!ensure_dev(c->d_carr, in_bytes) || (!out_is_device && !ensure_dev(c->d_out, out_bytes))) return fail(c, PARAMENT_STATUS_DEVICE_ALLOC_FAILED);
    // up to six groups of doubling size (>= 16k steps each along the time axis) for the copy / compute overlap of the register-resident family
    int G = (int)std::min<size_t>(6, in_bytes / ((size_t)2 << 20));
    if (const char *e = getenv("PARAMENT_COPY_GROUPS")) G = std::max(1, std::min(8, atoi(e)));   // A/B runs
    if (batch == 1) G = (int)std::min<unsigned long long>(G, s.nsteps / 16384);
    else G = (int)std::min<unsigned int>(G, batch);
    Parament_ErrorCode ec;
    c->stat_h2d = (double)in_bytes;
    // shared call: the last kernel of this device stores its partial straight into the gather buffer on the first device
    // (plain stores over NVLink peer memory) when that memory is addressable from here, else a peer copy follows
    const bool direct_store = gather_to && (gather_to == c || c->peer_store_ok);
    void *result_dev = direct_store ? (void *)((T *)gather_to->d_gather.ptr + (size_t)slot * n * n) : (out_is_device ? (void *)out : c->d_out.ptr);
    if (c->family == 1 && G >= 2) {
        ec = pipelined_family1<T>(c, carr, pts, p_lo, seg, s, G, result_dev);
    } else {
        if (seg == pts) {
            if (in_bytes && !h2d_rows(c, c->d_carr.ptr, 0, carr, 0, in_bytes, 1, c->stream)) return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
        } else {
            if (!h2d_rows(c, c->d_carr.ptr, seg * sizeof(T), carr + p_lo, (size_t)pts * sizeof(T), seg * sizeof(T), arrays, c->stream))
                return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
        }
        ec = propagate_device(c, c->d_carr.ptr, s, result_dev, c->stream);
    }
    if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
    if (gather_to) {
        if ((!direct_store && !PB_CUDA_OK(cudaMemcpyPeerAsync((T *)gather_to->d_gather.ptr + (size_t)slot * n * n, gather_to->device,
                                                               c->d_out.ptr, c->device, out_bytes, c->stream))) ||
            !PB_CUDA_OK(cudaStreamSynchronize(c->stream)))
            return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
        c->lastError = PARAMENT_STATUS_SUCCESS;
        return PARAMENT_STATUS_SUCCESS;
    }
    if ((!out_is_device && !PB_CUDA_OK(cudaMemcpyAsync(out, c->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, c->stream))) ||
        !PB_CUDA_OK(cudaStreamSynchronize(c->stream)))
        return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
    c->stat_d2h = out_is_device ? 0.0 : (double)out_bytes;
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}

template <typename T>
Parament_ErrorCode equiprop_device(Context *c, const T *carr_dev, double dt, unsigned int pts, unsigned int amps,
                                   unsigned int batch, T *out_dev, void *stream) {
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    DeviceGuard guard(c->device);
    if (!c->have_hamiltonian) return fail(c, PARAMENT_STATUS_NO_HAMILTONIAN);
    if (!out_dev || !carr_dev || (int)amps > c->amps || batch == 0) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    const unsigned long long N = effective_steps(c, pts);
    CallSpec s{};
    s.amps = amps; s.batch = batch; s.dt = dt; s.nsteps = N; s.total_steps = N; s.stride = pts;
    c->stat_ms = 0; c->stat_h2d = 0; c->stat_d2h = 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (N == 0 || c->Hnorm == 0.0) {
        std::vector<T> id((size_t)batch * c->dim * c->dim);
        write_identity(id.data(), c->dim, batch);
        if (!PB_CUDA_OK(cudaMemcpyAsync(out_dev, id.data(), id.size() * sizeof(T), cudaMemcpyHostToDevice, st)) ||
            !PB_CUDA_OK(cudaStreamSynchronize(st)))
            return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
        c->lastError = PARAMENT_STATUS_SUCCESS;
        return PARAMENT_STATUS_SUCCESS;
    }
    Parament_ErrorCode ec = propagate_device(c, carr_dev, s, out_dev, st);
    if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
    if (!stream && !PB_CUDA_OK(cudaStreamSynchronize(st))) return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}

// out = parts[count-1] ... parts[0]: ordered E-form tree on the GEMM kernel (multi-GPU combine of time slices).
// Device-resident core: parts_dev / out_dev in the IO precision, scratch from the context (grow-only), no synchronisation.
Parament_ErrorCode combine_device_core(Context *c, const void *parts_dev, unsigned int count, void *out_dev, cudaStream_t st) {
    NvtxRange range("parament: ordered combine of slice partials");
    const int n = c->dim;
    c->stat_launches = 0;
    if (c->family == 1) {   // register-resident family: one launch of one CTA
        PB_LAUNCH(launch_k3_combine(c->npad, c->fp64, parts_dev, count, n, out_dev, st));
        return PARAMENT_STATUS_SUCCESS;
    }
    const int gp = k4_pad(n);
    const size_t gnn = (size_t)gp * gp;
    if (!ensure_dev(c->d_comb, (size_t)count * gnn * sizeof(double2)) ||
        !ensure_dev(c->d_comb2, (size_t)(count / 2 + 1) * gnn * sizeof(double2)))
        return PARAMENT_STATUS_DEVICE_ALLOC_FAILED;
    c->stat_launches = 0;
    PB_LAUNCH(k4_eform(c->fp64, parts_dev, n, gp, (int)count, (double2 *)c->d_comb.ptr, st));
    Parament_ErrorCode ec = tree_reduce_all(c, (double2 *)c->d_comb.ptr, (int)count, (double2 *)c->d_comb2.ptr, gp, st);
    if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    PB_LAUNCH(k4_finish(c->fp64, (const double2 *)c->d_comb.ptr, n, gp, out_dev, true, st));
    return PARAMENT_STATUS_SUCCESS;
}

template <typename T>
Parament_ErrorCode combine_host(Context *c, const T *parts, unsigned int count, T *out) {
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    DeviceGuard guard(c->device);
    if (!c->have_hamiltonian) return fail(c, PARAMENT_STATUS_NO_HAMILTONIAN);
    if (!parts || !out || count == 0) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    const size_t in_bytes = (size_t)count * c->dim * c->dim * sizeof(T), out_bytes = (size_t)c->dim * c->dim * sizeof(T);
    if (!ensure_dev(c->d_carr, in_bytes) || !ensure_dev(c->d_out, out_bytes)) return fail(c, PARAMENT_STATUS_DEVICE_ALLOC_FAILED);
    if (!PB_CUDA_OK(cudaMemcpyAsync(c->d_carr.ptr, parts, in_bytes, cudaMemcpyHostToDevice, c->stream)))
        return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
    Parament_ErrorCode ec = combine_device_core(c, c->d_carr.ptr, count, c->d_out.ptr, c->stream);
    if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
    if (!PB_CUDA_OK(cudaMemcpyAsync(out, c->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, c->stream)) ||
        !PB_CUDA_OK(cudaStreamSynchronize(c->stream)))
        return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}

Parament_ErrorCode combine_device(Context *c, const void *parts_dev, unsigned int count, void *out_dev, void *stream) {
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    DeviceGuard guard(c->device);
    if (!c->have_hamiltonian) return fail(c, PARAMENT_STATUS_NO_HAMILTONIAN);
    if (!parts_dev || !out_dev || count == 0) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    Parament_ErrorCode ec = combine_device_core(c, parts_dev, count, out_dev, st);
    if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
    if (!stream && !PB_CUDA_OK(cudaStreamSynchronize(st))) return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}


// ---------------------------------------------------------------------------------------------------
// Single-process multi-GPU ("one host process drives all devices; the caller is a single ctypes call").
// A single pulse: the N effective steps are cut into contiguous slices, device g propagates slice g from the caller's
// host arrays (its own H2D of just that slice, its own stream, one host thread per device), the dim x dim partials travel
// to the first device by peer copy and are multiplied in order there (combine_device_core).  An ensemble: the pulses
// are cut into contiguous ranges, no exchange at all.  Work that is too small to share stays on one device.
// ---------------------------------------------------------------------------------------------------
template <typename T>
Parament_ErrorCode equiprop_multi(Context *c, const T *carr, double dt, unsigned int pts, unsigned int amps, unsigned int batch, T *out,
                                  bool &handled) {
    handled = false;
    if (!c->have_hamiltonian || !carr || !out || (int)amps > c->amps || batch == 0 || c->Hnorm == 0.0) return PARAMENT_STATUS_SUCCESS;
    const unsigned long long N = effective_steps(c, pts);
    const unsigned int G = devices_for_call((unsigned int)c->peers.size() + 1, batch, N, c->npad);   // plan.hpp
    if (G < 2) return PARAMENT_STATUS_SUCCESS;   // not worth sharing: the caller runs it on the first device
    handled = true;
    NvtxRange range("parament: single-process multi-GPU call");
    DeviceGuard guard(c->device);
    const int n = c->dim;
    const size_t nn = (size_t)n * n;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<Parament_ErrorCode> ecs(G, PARAMENT_STATUS_SUCCESS);
    auto ctx_of = [&](unsigned int g) { return g == 0 ? c : c->peers[g - 1]; };
    std::function<void(unsigned int)> run;
    if (batch == 1) {
        if (!ensure_dev(c->d_gather, (size_t)G * nn * sizeof(T))) return fail(c, PARAMENT_STATUS_DEVICE_ALLOC_FAILED);
        run = [&](unsigned int g) {   // runs on a helper thread for g > 0: no exception may leave it (std::terminate)
            try {
                ecs[g] = equiprop_host<T>(ctx_of(g), carr, dt, pts, amps, 1, N * g / G, N * (g + 1) / G, false, out, c, g);
            } catch (...) {
                ecs[g] = PARAMENT_STATUS_HOST_ALLOC_FAILED;
            }
        };
    } else {
        // whole = false with the full step range [0, N): neither this context nor a helper shares the work again
        run = [&](unsigned int g) {
            const size_t b0 = (size_t)batch * g / G, b1 = (size_t)batch * (g + 1) / G;
            try {
                ecs[g] = equiprop_host<T>(ctx_of(g), carr + b0 * amps * pts, dt, pts, amps, (unsigned int)(b1 - b0), 0, N, false, out + b0 * nn);
            } catch (...) {
                ecs[g] = PARAMENT_STATUS_HOST_ALLOC_FAILED;
            }
        };
    }
    for (unsigned int g = 1; g < G; ++g) ctx_of(g)->worker->submit([&run, g] { run(g); });   // fits std::function's inline storage
    run(0);
    for (unsigned int g = 1; g < G; ++g) ctx_of(g)->worker->wait();
    long long launches = 0;
    double h2d = 0;
    for (unsigned int g = 0; g < G; ++g) {
        if (ecs[g] != PARAMENT_STATUS_SUCCESS) return fail(c, ecs[g]);
        launches += ctx_of(g)->stat_launches;
        h2d += ctx_of(g)->stat_h2d;
    }
    if (batch == 1) {
        const size_t out_bytes = nn * sizeof(T);
        if (!ensure_dev(c->d_out, out_bytes)) return fail(c, PARAMENT_STATUS_DEVICE_ALLOC_FAILED);
        Parament_ErrorCode ec = combine_device_core(c, c->d_gather.ptr, G, c->d_out.ptr, c->stream);
        if (ec != PARAMENT_STATUS_SUCCESS) return fail(c, ec);
        launches += c->stat_launches;
        if (!PB_CUDA_OK(cudaMemcpyAsync(out, c->d_out.ptr, out_bytes, cudaMemcpyDeviceToHost, c->stream)) ||
            !PB_CUDA_OK(cudaStreamSynchronize(c->stream)))
            return fail(c, PARAMENT_STATUS_CUBLAS_FAILED);
    }
    c->stat_launches = launches;
    c->stat_h2d = h2d;
    c->stat_d2h = (double)(batch * nn * sizeof(T));
    c->stat_steps = N;
    c->stat_devices = (int)G;
    c->stat_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}

// devices[0] is where the context lives (moving it drops the Hamiltonian, like Parament_setDevice); devices[1..] get helper
// contexts.  A device may be listed more than once (its share of the work is then proportional; used by the tests on
// one-GPU machines).
Parament_ErrorCode move_to_device(Context *c, int device);

Parament_ErrorCode set_device_list(Context *c, const int *devices, int count) {
    if (!c || c->is_peer) return PARAMENT_STATUS_INVALID_VALUE;
    int ndev = 0;
    if (!devices || count < 1 || cudaGetDeviceCount(&ndev) != cudaSuccess) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    for (int i = 0; i < count; ++i)
        if (devices[i] < 0 || devices[i] >= ndev) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    {
        DeviceGuard guard(c->device);
        cudaStreamSynchronize(c->stream);
    }
    destroy_peers(c);
    if (devices[0] != c->device) {
        const Parament_ErrorCode ec = move_to_device(c, devices[0]);
        if (ec != PARAMENT_STATUS_SUCCESS) return ec;
    }
    for (int i = 1; i < count; ++i) {
        Context *p = nullptr;
        Parament_ErrorCode ec = create_ctx_on(&p, c->fp64, devices[i]);
        if (ec == PARAMENT_STATUS_SUCCESS) {
            p->is_peer = true;
            ec = replay_hamiltonian(c, p);
            if (ec == PARAMENT_STATUS_SUCCESS) {
                try {
                    p->worker = new DeviceWorker();
                    p->worker->start();
                } catch (...) {
                    delete p->worker;
                    p->worker = nullptr;
                    ec = PARAMENT_STATUS_HOST_ALLOC_FAILED;
                }
            }
            if (ec != PARAMENT_STATUS_SUCCESS) destroy_ctx(p);
        }
        if (ec != PARAMENT_STATUS_SUCCESS) {
            destroy_peers(c);
            return fail(c, ec);
        }
        c->peers.push_back(p);
        p->peer_store_ok = devices[i] == c->device;
        if (devices[i] != c->device) {   // direct NVLink path for the partials: the helper's kernels store into this device's memory
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], c->device) == cudaSuccess && can) {
                DeviceGuard guard(devices[i]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(c->device, 0);
                p->peer_store_ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
            }
            cudaGetLastError();
        }
    }
    c->lastError = PARAMENT_STATUS_SUCCESS;
    return PARAMENT_STATUS_SUCCESS;
}

Parament_ErrorCode set_device_count(Context *c, int ngpus) {
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
    std::vector<int> list(ngpus);
    for (int i = 0; i < ngpus; ++i) list[i] = (c->device + i) % ndev;
    return set_device_list(c, list.data(), ngpus);
}

// C++ exceptions must not cross the C ABI: every entry point that can allocate on the host runs through this.
template <typename F>
Parament_ErrorCode guarded(void *h, F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return fail(as_ctx(h), PARAMENT_STATUS_HOST_ALLOC_FAILED);
    } catch (...) {
        return fail(as_ctx(h), PARAMENT_FAIL);
    }
}

}  // namespace

// =====================================================================================================
// exported C symbols
// =====================================================================================================
extern "C" {

Parament_ErrorCode Parament_create(struct Parament_Context_f32 **h) { return create_ctx(reinterpret_cast<Context **>(h), false); }
Parament_ErrorCode Parament_create_fp64(struct Parament_Context_f64 **h) { return create_ctx(reinterpret_cast<Context **>(h), true); }
Parament_ErrorCode Parament_destroy(struct Parament_Context_f32 *h) { return destroy_ctx(h ? as_ctx(h) : nullptr); }
Parament_ErrorCode Parament_destroy_fp64(struct Parament_Context_f64 *h) { return destroy_ctx(h ? as_ctx(h) : nullptr); }

Parament_ErrorCode Parament_setHamiltonian(struct Parament_Context_f32 *h, const Parament_c64 *H0, const Parament_c64 *H1,
                                           unsigned int dim, unsigned int amps, bool use_magnus, enum Parament_QuadratureSpec q) {
    return guarded(h, [&] { return set_hamiltonian<Parament_c64>(as_ctx(h), H0, H1, dim, amps, use_magnus, (int)q); });
}
Parament_ErrorCode Parament_setHamiltonian_fp64(struct Parament_Context_f64 *h, const Parament_c128 *H0, const Parament_c128 *H1,
                                                unsigned int dim, unsigned int amps, bool use_magnus, Parament_QuadratureSpec q) {
    return guarded(h, [&] { return set_hamiltonian<Parament_c128>(as_ctx(h), H0, H1, dim, amps, use_magnus, (int)q); });
}

Parament_ErrorCode Parament_equiprop(struct Parament_Context_f32 *h, const Parament_c64 *carr, double dt, unsigned int pts,
                                     unsigned int amps, Parament_c64 *out) {
    return guarded(h, [&] { return equiprop_host<Parament_c64>(as_ctx(h), carr, dt, pts, amps, 1, 0, 0, true, out); });
}
Parament_ErrorCode Parament_equiprop_fp64(struct Parament_Context_f64 *h, const Parament_c128 *carr, double dt, unsigned int pts,
                                          unsigned int amps, Parament_c128 *out) {
    return guarded(h, [&] { return equiprop_host<Parament_c128>(as_ctx(h), carr, dt, pts, amps, 1, 0, 0, true, out); });
}

Parament_ErrorCode Parament_equipropBatch(struct Parament_Context_f32 *h, const Parament_c64 *carr, double dt, unsigned int pts,
                                          unsigned int amps, unsigned int batch, Parament_c64 *out) {
    return guarded(h, [&] { return equiprop_host<Parament_c64>(as_ctx(h), carr, dt, pts, amps, batch, 0, 0, true, out); });
}
Parament_ErrorCode Parament_equipropBatch_fp64(struct Parament_Context_f64 *h, const Parament_c128 *carr, double dt, unsigned int pts,
                                               unsigned int amps, unsigned int batch, Parament_c128 *out) {
    return guarded(h, [&] { return equiprop_host<Parament_c128>(as_ctx(h), carr, dt, pts, amps, batch, 0, 0, true, out); });
}

Parament_ErrorCode Parament_equipropDevice(struct Parament_Context_f32 *h, const Parament_c64 *carr_dev, double dt, unsigned int pts,
                                           unsigned int amps, unsigned int batch, Parament_c64 *out_dev, void *stream) {
    return guarded(h, [&] { return equiprop_device<Parament_c64>(as_ctx(h), carr_dev, dt, pts, amps, batch, out_dev, stream); });
}
Parament_ErrorCode Parament_equipropDevice_fp64(struct Parament_Context_f64 *h, const Parament_c128 *carr_dev, double dt, unsigned int pts,
                                                unsigned int amps, unsigned int batch, Parament_c128 *out_dev, void *stream) {
    return guarded(h, [&] { return equiprop_device<Parament_c128>(as_ctx(h), carr_dev, dt, pts, amps, batch, out_dev, stream); });
}

Parament_ErrorCode Parament_equipropSlice(struct Parament_Context_f32 *h, const Parament_c64 *carr, double dt, unsigned int pts,
                                          unsigned int amps, unsigned long long lo, unsigned long long hi, Parament_c64 *out) {
    return guarded(h, [&] { return equiprop_host<Parament_c64>(as_ctx(h), carr, dt, pts, amps, 1, lo, hi, false, out); });
}
Parament_ErrorCode Parament_equipropSlice_fp64(struct Parament_Context_f64 *h, const Parament_c128 *carr, double dt, unsigned int pts,
                                               unsigned int amps, unsigned long long lo, unsigned long long hi, Parament_c128 *out) {
    return guarded(h, [&] { return equiprop_host<Parament_c128>(as_ctx(h), carr, dt, pts, amps, 1, lo, hi, false, out); });
}

// host amplitudes in, partial propagator left on the device (include/parament.h section 2)
Parament_ErrorCode Parament_equipropSliceToDevice(struct Parament_Context_f32 *h, const Parament_c64 *carr, double dt, unsigned int pts,
                                                  unsigned int amps, unsigned long long lo, unsigned long long hi, Parament_c64 *out_dev) {
    return guarded(h, [&] { return equiprop_host<Parament_c64>(as_ctx(h), carr, dt, pts, amps, 1, lo, hi, false, out_dev, nullptr, 0, true); });
}
Parament_ErrorCode Parament_equipropSliceToDevice_fp64(struct Parament_Context_f64 *h, const Parament_c128 *carr, double dt, unsigned int pts,
                                                       unsigned int amps, unsigned long long lo, unsigned long long hi, Parament_c128 *out_dev) {
    return guarded(h, [&] { return equiprop_host<Parament_c128>(as_ctx(h), carr, dt, pts, amps, 1, lo, hi, false, out_dev, nullptr, 0, true); });
}

Parament_ErrorCode Parament_combine(struct Parament_Context_f32 *h, const Parament_c64 *parts, unsigned int count, Parament_c64 *out) {
    return guarded(h, [&] { return combine_host<Parament_c64>(as_ctx(h), parts, count, out); });
}
Parament_ErrorCode Parament_combine_fp64(struct Parament_Context_f64 *h, const Parament_c128 *parts, unsigned int count, Parament_c128 *out) {
    return guarded(h, [&] { return combine_host<Parament_c128>(as_ctx(h), parts, count, out); });
}

Parament_ErrorCode Parament_combineDevice(void *h, const void *parts_dev, unsigned int count, void *out_dev, void *stream) {
    return combine_device(as_ctx(h), parts_dev, count, out_dev, stream);
}

int Parament_selectIterationCycles_fp32(double H_norm, double dt) { return select_cycles_fp32(H_norm, dt); }
int Parament_selectIterationCycles_fp64(double H_norm, double dt) { return select_cycles_fp64(H_norm, dt); }

static Parament_ErrorCode set_cycles(Context *c, unsigned int cycles) {   // parament.cpp:772-776
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    c->MMAX = (int)cycles;
    c->MMAX_manual = true;
    for (Context *p : c->peers) { p->MMAX = (int)cycles; p->MMAX_manual = true; }
    return PARAMENT_STATUS_SUCCESS;
}
static Parament_ErrorCode auto_cycles(Context *c) {   // parament.cpp:782-786
    if (!c) return PARAMENT_STATUS_INVALID_VALUE;
    c->MMAX = 11;
    c->MMAX_manual = false;
    for (Context *p : c->peers) { p->MMAX = 11; p->MMAX_manual = false; }
    return PARAMENT_STATUS_SUCCESS;
}
Parament_ErrorCode Parament_setIterationCyclesManually(struct Parament_Context_f32 *h, unsigned int cycles) { return set_cycles(as_ctx(h), cycles); }
Parament_ErrorCode Parament_setIterationCyclesManually_fp64(struct Parament_Context_f64 *h, unsigned int cycles) { return set_cycles(as_ctx(h), cycles); }
Parament_ErrorCode Parament_automaticIterationCycles(struct Parament_Context_f32 *h) { return auto_cycles(as_ctx(h)); }
Parament_ErrorCode Parament_automaticIterationCycles_fp64(struct Parament_Context_f64 *h) { return auto_cycles(as_ctx(h)); }

Parament_ErrorCode Parament_peekAtLastError(struct Parament_Context_f32 *h) { Context *c = as_ctx(h); return c ? c->lastError : PARAMENT_STATUS_INVALID_VALUE; }
Parament_ErrorCode Parament_peekAtLastError_fp64(struct Parament_Context_f64 *h) { Context *c = as_ctx(h); return c ? c->lastError : PARAMENT_STATUS_INVALID_VALUE; }
Parament_ErrorCode Parament_getLastError(void *h) { Context *c = as_ctx(h); return c ? c->lastError : PARAMENT_STATUS_INVALID_VALUE; }

const char *Parament_errorMessage(Parament_ErrorCode errorCode) {   // strings byte-identical to parament.cpp:859-882
    switch (errorCode) {
        case PARAMENT_STATUS_SUCCESS: return "Success";
        case PARAMENT_STATUS_HOST_ALLOC_FAILED: return "Memory allocation on the host failed.";
        case PARAMENT_STATUS_DEVICE_ALLOC_FAILED: return "Memory allocation on the device failed.";
        case PARAMENT_STATUS_CUBLAS_INIT_FAILED: return "Failed to initialize the cuBLAS library.";
        case PARAMENT_STATUS_INVALID_VALUE: return "Invalid value.";
        case PARAMENT_STATUS_CUBLAS_FAILED: return "Failed to execute cuBLAS function.";
        case PARAMENT_STATUS_SELECT_SMALLER_DT: return "Timestep too large";
        case PARAMENT_STATUS_INVALID_QUADRATURE_SELECTION: return "Invalid quadrature selection.";
        case PARAMENT_STATUS_NO_HAMILTONIAN: return "No hamiltonian set";
        default: return "Unknown error code";
    }
}

double OneNorm(const Parament_c64 *mat, unsigned int dim) { return one_norm_t<Parament_c64>(mat, dim); }
double OneNorm_fp64(const Parament_c128 *mat, unsigned int dim) { return one_norm_t<Parament_c128>(mat, dim); }

void device_info(void) {   // same table as deviceInfo.c:30-59
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        printf("Failed to query the number of CUDA devices. Error code: %d\n", (int)e);
        cudaGetLastError();
        return;
    }
    printf("PARAMENT_INFO:\nTotal number of CUDA devices: %d\n-----------------------------------\n", n);
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) != cudaSuccess) {
            printf("Failed to query device properties.\n");
            return;
        }
        printf("Device Number: %d\n  Device name: %s\n", i, prop.name);
        printf("  Memory Clock Rate (KHz): %d\n  Memory Bus Width (bits): %d\n", prop.memoryClockRate, prop.memoryBusWidth);
        printf("  Peak Memory Bandwidth (GB/s): %f\n", 2.0 * prop.memoryClockRate * (prop.memoryBusWidth / 8) / 1.0e6);
        printf("  Total global memory: %zd MB\n\n", prop.totalGlobalMem / 1024 / 1024);
    }
    fflush(stdout);
}

double Parament_lastStat(void *h, int key) {
    Context *c = as_ctx(h);
    if (!c) return -1.0;
    switch (key) {
        case 0: {
            if (c->stat_devices > 1) return c->stat_ms;   // shared call: host wall clock around all devices and the combine
            float ms = 0;
            DeviceGuard guard(c->device);
            if (cudaEventSynchronize(c->ev_stop) == cudaSuccess && cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop) == cudaSuccess) return ms;
            cudaGetLastError();
            return -1.0;
        }
        case 1: return (double)c->stat_launches;
        case 2: return c->stat_M_used;
        case 3: return c->stat_M_ref;
        case 4: return (double)c->stat_steps;
        case 5: return c->family;
        case 6: return c->stat_h2d;
        case 7: return c->stat_d2h;
        case 8: return c->Hnorm;
        case 9: return c->stat_horner;
        case 10: {   // complex matrix products executed per effective step (series + ordered product)
            const int M = c->stat_M_used;
            if (M <= 0) return 0.0;
            if (c->family == 3 && c->stat_products_saved > 0.0) {   // Hermitian square: only the upper-triangular tiles were computed
                const double full = c->stat_horner == 2 ? 4.0 + (M >> 2) : (c->stat_horner == 3 ? 4.0 : (c->stat_horner == 4 ? 5.0 : 2.0 + (M >> 1)));
                return full - c->stat_products_saved;
            }
            if (c->stat_horner == 2) return 4.0 + (M >> 2);
            if (c->stat_horner == 3) return 4.0;   // degree 8 in three products + the ordered product
            if (c->stat_horner == 4) return 5.0;   // degree 12 in four products + the ordered product
            if (c->stat_horner == 1 && c->family == 2 && c->onchip && (M == 8 || M >= 10)) return 3.0 + M / 3;   // blocks of three
            return c->stat_horner ? 2.0 + (M >> 1) : (double)std::max(M, 1);
        }
        case 13:   // real matrix products per complex matrix product, averaged over the products of a step
            if (c->family == 3) return k4_real_products(c->npad);
            // shared-memory-resident kernel: three real products wherever the own elements of Y and W are not both live in
            // registers -- all but the L/R product of the degree-8 form, all but two products of the degree-12 form
            if (c->family == 2 && c->onchip && c->stat_horner == 3) return 13.0 / 4.0;
            if (c->family == 2 && c->onchip && c->stat_horner == 4) return 17.0 / 5.0;
            if (c->family == 2 && c->onchip) {   // other forms: only the running-product update
                const double np_ = Parament_lastStat(h, 10);
                return np_ > 0 ? (4.0 * (np_ - 1.0) + 3.0) / np_ : 4.0;
            }
            if (c->family == 1 && c->stat_math != 1) return k1_real_products(c->npad, c->fp64, c->stat_horner);
            return 4;
        case 14: return c->stat_series_norm;
        case 15: return c->stat_math;
        case 11: return c->stat_devices;
        case 12: return (double)c->peers.size() + 1.0;
        default: return -1.0;
    }
}

Parament_ErrorCode Parament_setDevice(void *h, int device) {
    Context *c = as_ctx(h);
    if (!c || c->is_peer) return PARAMENT_STATUS_INVALID_VALUE;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(c, PARAMENT_STATUS_INVALID_VALUE);
    destroy_peers(c);   // back to one device; Parament_setDevices adds helpers again
    if (device == c->device) return PARAMENT_STATUS_SUCCESS;
    return move_to_device(c, device);
}

Parament_ErrorCode Parament_setDevices(void *h, int ngpus) { return set_device_count(as_ctx(h), ngpus); }
Parament_ErrorCode Parament_setDeviceList(void *h, const int *devices, int count) { return set_device_list(as_ctx(h), devices, count); }

}  // extern "C"

namespace {
Parament_ErrorCode move_to_device(Context *c, int device) {
    // move: drop everything that lives on the old device (handles are nulled: a failed re-creation leaves nothing dangling)
    {
        DeviceGuard guard(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        destroy_device_objects(c);
    }
    c->device = device;
    c->have_hamiltonian = false;
    c->series_cache.valid = false;
    DeviceGuard guard(device);
    if (!create_device_objects(c)) {
        destroy_device_objects(c);
        return fail(c, PARAMENT_STATUS_CUBLAS_INIT_FAILED);
    }
    return PARAMENT_STATUS_SUCCESS;
}
}  // namespace

extern "C" {

const char *Parament_version(void) { return "parament-b200 0.1 (sm_100a, FP64 tensor pipe)"; }

}  // extern "C"
