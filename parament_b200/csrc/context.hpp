// context.hpp -- the opaque Parament context of this library (reference: parament_context.hpp:26-78, which
// holds a cuBLAS handle and three dim^2 x pts work arrays; none of that exists here).
#pragma once
#include <complex>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/parament.h"
#include "params.hpp"

namespace pb {

typedef std::complex<double> zc;

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

constexpr int kCopyEvents = 8;

// Host thread that drives one helper device of the single-process multi-GPU mode: it lives as long as the helper context,
// so a shared call costs two condition-variable hand-offs per device instead of a thread creation.
struct DeviceWorker {
    std::thread thread;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, done = false, quit = false;

    void start() { thread = std::thread([this] { run(); }); }
    void run() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [this] { return has_job || quit; });
            if (quit) return;
            lk.unlock();
            job();
            lk.lock();
            has_job = false;
            done = true;
            cv.notify_all();
        }
    }
    void submit(std::function<void()> f) {
        { std::lock_guard<std::mutex> lk(m); job = std::move(f); has_job = true; done = false; }
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [this] { return done; });
    }
    void stop() {
        if (!thread.joinable()) return;
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cv.notify_all();
        thread.join();
    }
};

struct Context {
    unsigned int magic = 0x50423230;   // "PB20"
    bool fp64 = false;                 // context precision of the I/O operands
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;        // H2D of the amplitude stream, overlapped with the kernels; second chunk stream of family 3
    cudaStream_t f3_streams[3] = {};           // further chunk streams of family 3 (created on first use)
    cudaEvent_t ev_copy[kCopyEvents] = {};     // copy groups of the host-pointer pipeline; fork / join / reduction events of family 3
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    Parament_ErrorCode lastError = PARAMENT_STATUS_SUCCESS;

    // options (reference: parament_context.hpp:72-77)
    int MMAX = 11;
    bool MMAX_manual = false;
    bool enable_magnus = false;
    int quadrature = PARAMENT_QUADRATURE_NONE;

    // Hamiltonian (host copies in double, exact conversions of the inputs)
    bool have_hamiltonian = false;
    int dim = 0;
    int amps = 0;        // controls given to setHamiltonian
    int nmats = 0;       // 1 + amps (+ amps + amps(amps-1)/2 commutators with Magnus)
    double Hnorm = 0.0;      // the reference's norm bound (parament.cpp:280-284): error semantics and the reported table degree
    std::vector<double> sigma_max;   // largest singular value of H0 and of every control matrix (power iteration, setHamiltonian)
    int norm_mode = 1;       // 1: series domain from the spectral bound below (dim > 16); 0: the reference's Hnorm ($PARAMENT_NORM=reference)
    double stat_series_norm = 0.0;   // norm the series of the last call was built for
    DeviceBuffer d_absmax;   // per control: max |c_k|^2 of the last call's amplitude stream (bits of a double, atomicMax)
    double *h_absmax = nullptr;      // pinned host copy
    std::vector<zc> mats;   // nmats * dim * dim, row-major
    int family = 0;      // 1 = register-resident warp kernels, 2 = persistent CTA chain kernel, 3 = batched GEMM pipeline
    int npad = 0;
    bool onchip = false; // family 2: operands resident in shared memory (npad == 64)
    int k4_slots = 0;    // co-resident CTAs of the GEMM kernel (family 3)
    DeviceBuffer d_H;    // family 1: fragment-ordered table; family 3: padded row-major table

    // per-call scratch (grow-only)
    DeviceBuffer d_carr, d_out, d_partials;
    DeviceBuffer d_Y, d_pending, d_tree;   // families 2 / 3: series slots, pending partial products, tree scratch
    DeviceBuffer d_comb, d_comb2;   // combine of time-slice partials
    DeviceBuffer d_counters;        // arrival counters of the fused ordered reduction (zero between launches)
    DeviceBuffer d_gather;          // single-process multi-GPU: partial propagators of all devices, slice order

    // Single-process multi-GPU (Parament_setDevices / $PARAMENT_NUM_GPUS): helper contexts on the other devices, owned
    // by this one.  The time axis of a pulse (or the pulses of an ensemble) is cut into contiguous shares, every
    // device reduces its share on its own stream, partials arrive by peer copy and are combined in order here.
    std::vector<Context *> peers;
    bool is_peer = false;
    bool peer_store_ok = false;     // helper: its kernels may store into the first device's memory (same device or peer access)
    DeviceWorker *worker = nullptr; // helper: the host thread that drives this device
    int stat_devices = 1;           // devices that took part in the last equiprop

    // statistics of the last equiprop (Parament_lastStat)
    double stat_ms = 0.0;
    long long stat_launches = 0;
    int stat_M_used = 0, stat_M_ref = 0, stat_horner = 0;
    int stat_math = 0;     // arithmetic of the last call: 0 = FP64 (DMMA), 1 = FP32 as 3xTF32 on the warp-level tensor path
    int series_mode = 0;   // 0 automatic, 1 force the Clenshaw recurrence ($PARAMENT_SERIES=clenshaw), 2 Horner / PS only
    // series constants of the last call: rebuilt only when the step size, the norm or the degree changes (the product-saving
    // forms solve a small nonlinear system per distinct step size)
    struct SeriesCache {
        bool valid = false;
        double h = 0.0, Hnorm = 0.0;
        int M = 0, family = 0, onchip = 0, flags = 0;
        int horner = 0;
        double sigma = 0.0;
        cplx a[kMaxDegree + 1], a_lo[kMaxDegree + 1];
    } series_cache;
    unsigned long long stat_steps = 0;
    double stat_h2d = 0.0, stat_d2h = 0.0;
};

}  // namespace pb
