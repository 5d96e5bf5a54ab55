// context.hpp -- the opaque Parament context of this library (reference: parament_context.hpp:26-78, which
// holds a cuBLAS handle and three dim^2 x pts work arrays; none of that exists here).
#pragma once
#include <algorithm>
#include <complex>
#include <cstring>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
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

// Host-to-device staging of PAGEABLE caller buffers (what the unchanged pyparament wrapper passes, parament.py:263-272).
// cudaMemcpyAsync from pageable memory goes through the driver's own single-threaded bounce buffer (~10 GB/s measured: the
// 160 MB of C5 took 16 ms against 5.5 ms from page-locked memory).  Here the transfer is cut into 2 MB chunks; a few persistent
// host threads (and the calling thread while it waits) copy chunks into a ring of page-locked slots with streaming stores, and
// the calling thread sends every finished chunk, in order, with an asynchronous DMA that overlaps the copies of the following
// chunks.  The reference does one blocking cudaMemcpy of the whole array (parament.cpp:477).
struct Stager {
    static constexpr size_t kSlotBytes = (size_t)2 << 20;
    static constexpr int kSlots = 12;
    struct Task { const char *src = nullptr; size_t bytes = 0; int state = 0; };   // state: 0 idle, 1 queued / being copied, 2 copied
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::vector<int> queue;                 // slots waiting for a copying thread (FIFO by position `qhead`)
    size_t qhead = 0;
    Task tasks[kSlots];
    bool quit = false;
    char *slots[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    bool ev_used[kSlots] = {};
    int next_slot = 0;

    // streaming copy: the destination (write-combined on its way to DRAM, read next by the DMA engine) bypasses the caches and costs
    // no read-for-ownership; dst is 64-byte aligned (slot bases are page aligned), src arbitrary
    static void copy_stream(char *dst, const char *src, size_t n) {
#if defined(__SSE2__)
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32)), d = _mm_loadu_si128((const __m128i *)(src + i + 48));
            _mm_stream_si128((__m128i *)(dst + i), a);
            _mm_stream_si128((__m128i *)(dst + i + 16), b);
            _mm_stream_si128((__m128i *)(dst + i + 32), c);
            _mm_stream_si128((__m128i *)(dst + i + 48), d);
        }
        _mm_sfence();
        if (i < n) memcpy(dst + i, src + i, n - i);
#else
        memcpy(dst, src, n);
#endif
    }
    // one queued chunk, if any (lock held on entry and exit)
    bool run_one(std::unique_lock<std::mutex> &lk) {
        if (qhead >= queue.size()) return false;
        const int s = queue[qhead++];
        if (qhead == queue.size()) { queue.clear(); qhead = 0; }
        Task t = tasks[s];
        lk.unlock();
        copy_stream(slots[s], t.src, t.bytes);
        lk.lock();
        tasks[s].state = 2;
        cv_done.notify_all();
        return true;
    }
    void worker() {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_work.wait(lk, [&] { return quit || qhead < queue.size(); });
            if (quit) return;
            run_one(lk);
        }
    }
    bool start(int helpers) {
        for (int i = 0; i < kSlots; ++i) {
            if (cudaHostAlloc((void **)&slots[i], kSlotBytes, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        try {
            queue.reserve(kSlots);
            for (int i = 0; i < helpers; ++i) threads.emplace_back([this] { worker(); });
        } catch (...) {
            // fewer helpers than asked for: the calling thread copies as well
        }
        return true;
    }
    // dst (device) <- src (pageable host), asynchronous on `stream` once this returns; src may be released on return
    bool copy(void *dst, const void *src, size_t bytes, cudaStream_t stream) {
        // chunks of 2 MB for long transfers; shorter ones are cut into about eight chunks (>= 128 KB) so that every copying
        // thread gets a share of them too
        const size_t chunk = bytes >= 8 * kSlotBytes ? kSlotBytes
                                                     : std::min(kSlotBytes, std::max<size_t>((size_t)128 << 10, ((bytes / 8 + 65535) >> 16) << 16));
        const size_t nchunks = (bytes + chunk - 1) / chunk;
        size_t queued = 0, sent = 0;
        const int base = next_slot;
        auto slot_of = [&](size_t i) { return (int)((base + i) % kSlots); };
        std::unique_lock<std::mutex> lk(m);
        while (sent < nchunks) {
            while (queued < nchunks && queued - sent < (size_t)kSlots) {
                const int s = slot_of(queued);
                if (ev_used[s]) {   // the slot's previous DMA has to have drained
                    lk.unlock();
                    const cudaError_t e = cudaEventSynchronize(ev[s]);
                    lk.lock();
                    if (e != cudaSuccess) return false;
                    ev_used[s] = false;
                }
                tasks[s].src = (const char *)src + queued * chunk;
                tasks[s].bytes = std::min(chunk, bytes - queued * chunk);
                tasks[s].state = 1;
                queue.push_back(s);
                ++queued;
            }
            cv_work.notify_all();
            const int s = slot_of(sent);
            while (tasks[s].state != 2)
                if (!run_one(lk)) cv_done.wait(lk, [&] { return tasks[s].state == 2 || qhead < queue.size(); });   // help, or wait
            tasks[s].state = 0;
            const size_t n = tasks[s].bytes;
            lk.unlock();
            const bool ok = cudaMemcpyAsync((char *)dst + sent * chunk, slots[s], n, cudaMemcpyHostToDevice, stream) == cudaSuccess &&
                            cudaEventRecord(ev[s], stream) == cudaSuccess;
            lk.lock();
            if (!ok) return false;
            ev_used[s] = true;
            ++sent;
        }
        next_slot = slot_of(nchunks);
        return true;
    }
    void stop() {   // with the owning context's device current
        {
            std::lock_guard<std::mutex> lk(m);
            quit = true;
        }
        cv_work.notify_all();
        for (std::thread &t : threads) if (t.joinable()) t.join();
        threads.clear();
        for (int i = 0; i < kSlots; ++i) {
            if (ev[i]) { if (ev_used[i]) cudaEventSynchronize(ev[i]); cudaEventDestroy(ev[i]); ev[i] = nullptr; }
            if (slots[i]) { cudaFreeHost(slots[i]); slots[i] = nullptr; }
        }
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
    bool hermitian = false;   // H0 and all H_k equal their conjugate transposes exactly (set_hamiltonian); $PARAMENT_K1_HERM=0 ignores it
    int pack = 1;        // family 1, dim <= 4: systems per 8 x 8 tile (4 for dim <= 2, 2 for dim 3..4), see k1_warp.cu; $PARAMENT_K1_PACK=0 disables
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
    Stager *stager = nullptr;       // staging threads + page-locked ring for pageable caller buffers (created on first use)
    bool stager_failed = false;
    int stat_devices = 1;           // devices that took part in the last equiprop

    // statistics of the last equiprop (Parament_lastStat)
    double stat_ms = 0.0;
    long long stat_launches = 0;
    int stat_M_used = 0, stat_M_ref = 0, stat_horner = 0;
    double stat_products_saved = 0.0;   // complex products per step NOT executed because only the upper-triangular tiles of a Hermitian square were computed
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
        float fconst[16];
    } series_cache;
    unsigned long long stat_steps = 0;
    double stat_h2d = 0.0, stat_d2h = 0.0;
};

}  // namespace pb
