// plan.hpp -- host-side work partitioning of the register-resident family and of the copy / device sharing (pure integer
// logic, no CUDA types: compiled into the library and, with g++, into tests/cpp/plan_check.cpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>

namespace pb {

constexpr int K1_WARPS = 4;   // warps per CTA of the chain kernel

struct K1Plan {
    unsigned int grid;                 // CTAs of the chain kernel
    unsigned int chunks_per_pulse;     // warps cooperating on one pulse
    unsigned int partials_per_pulse;   // matrices the reduce kernel combines per pulse
    int reduce_in_cta;                 // 1: the warps of a CTA belong to one pulse and combine in shared memory
    int k3_warps;                      // warps per CTA of the reduce kernel
    int ctas_per_sm;                   // occupancy the chain kernel variant is compiled for
    size_t partial_elems;              // double2 elements of the partial buffer
    unsigned int groups_per_pulse;     // fused final stage: groups of 16 CTA partials per pulse (few-long-pulses mode), else 0
};

inline int k3_warps_for(unsigned int partials_per_pulse) { return partials_per_pulse >= 16 ? 8 : (partials_per_pulse >= 4 ? 4 : 1); }
inline size_t k3_mid_elems(int npad, unsigned int batch, unsigned int partials_per_pulse) {   // scratch of the two-level reduction
    return partials_per_pulse > 32 ? (size_t)batch * ((partials_per_pulse + 15) / 16) * npad * npad : 0;
}
inline int k3_launches(unsigned int partials_per_pulse) { return partials_per_pulse > 32 ? 2 : 1; }

inline int k1_ctas_per_sm(int npad, bool horner) {
    static const int occ_env = getenv("PARAMENT_K1_OCC") ? atoi(getenv("PARAMENT_K1_OCC")) : 0;
    return (npad == 8) ? 6 : (occ_env == 2 || occ_env == 3 ? occ_env : (horner ? 2 : 3));
}

// co-resident warps of the chain kernel (one wave)
inline unsigned int k1_warp_slots(int npad, int num_sms, bool horner) {
    return (unsigned int)(num_sms * k1_ctas_per_sm(npad, horner) * K1_WARPS);
}

// How `batch` pulses of `nsteps` effective steps are spread over warps and CTAs.
// fuse: the chain launch writes the propagators itself (k1_common.cuh); the warps of an ensemble pulse must then share a CTA,
// i.e. a pulse is cut into 1, 2 or 4 chunks.
inline K1Plan plan_k1(int npad, unsigned int batch, unsigned long long nsteps, int num_sms, bool horner, bool fuse = false) {
    K1Plan plan{};
    const int ctas_per_sm = k1_ctas_per_sm(npad, horner);
    plan.ctas_per_sm = ctas_per_sm;
    const unsigned long long warps_total = (unsigned long long)num_sms * ctas_per_sm * K1_WARPS;
    if (batch >= warps_total / 2 || nsteps < 2ull * K1_WARPS) {
        // Ensemble: a warp owns a whole pulse or 1/k of it.  The kernel is bound by the FP64 pipe of the SM, not by latency, so
        // what matters is that every SM gets the same number of CTAs, and that CTAs are short: measured at dim 8 (3552 pulses
        // of 1000 steps = exactly six CTAs per SM with k = 1): k = 1 2.654 ms, 2 2.532, 4 2.475, 8 2.450 -- about 0.91 + 0.09 / k.
        unsigned int best_k = 1;
        double best = 1e300;
        for (unsigned int k = 1; k <= (fuse ? 4u : 8u); ++k) {
            if (k > 1 && nsteps / k < 64) break;
            if (fuse && k == 3) continue;
            const double ctas = std::ceil((double)batch * k / K1_WARPS);
            const double imbalance = std::ceil(ctas / num_sms) / (ctas / num_sms);
            const double cost = imbalance * (0.91 + 0.09 / k);
            if (cost < best * 0.999) { best = cost; best_k = k; }
        }
        static const int k_env = getenv("PARAMENT_K1_K") ? atoi(getenv("PARAMENT_K1_K")) : 0;   // A/B runs
        if (k_env >= 1 && k_env <= 8 && nsteps / k_env >= 64 && (!fuse || k_env == 1 || k_env == 2 || k_env == 4)) best_k = (unsigned int)k_env;
        plan.chunks_per_pulse = best_k;
        plan.reduce_in_cta = 0;
        plan.partials_per_pulse = best_k;
    } else {
        // Few long pulses: the CTAs of one wave are dealt to the pulses; more rounds were measured without effect (C2, 1..8 rounds:
        // 2.281 .. 2.306 ms).
        static const int waves_env = getenv("PARAMENT_K1_WAVES") ? atoi(getenv("PARAMENT_K1_WAVES")) : 0;   // A/B runs
        const unsigned long long rounds = waves_env >= 1 && waves_env <= 16 ? waves_env : 1;
        unsigned long long ctas_per_pulse = (rounds * (warps_total / K1_WARPS) + batch - 1) / batch;
        // keep at least ~8 steps per warp so the identity-start product stays a small fraction
        static const int min_env = getenv("PARAMENT_K1_MIN_STEPS") ? atoi(getenv("PARAMENT_K1_MIN_STEPS")) : 0;   // A/B runs
        const unsigned long long min_steps = min_env >= 1 && min_env <= 64 ? min_env : 8;
        const unsigned long long max_ctas = (nsteps / min_steps + K1_WARPS - 1) / K1_WARPS;
        if (ctas_per_pulse > max_ctas) ctas_per_pulse = max_ctas;
        if (ctas_per_pulse < 1) ctas_per_pulse = 1;
        plan.chunks_per_pulse = (unsigned int)(ctas_per_pulse * K1_WARPS);
        plan.reduce_in_cta = 1;
        plan.partials_per_pulse = (unsigned int)ctas_per_pulse;
    }
    const unsigned long long total_warps = (unsigned long long)batch * plan.chunks_per_pulse;
    plan.grid = (unsigned int)((total_warps + K1_WARPS - 1) / K1_WARPS);
    plan.k3_warps = k3_warps_for(plan.partials_per_pulse);
    plan.partial_elems = (size_t)batch * plan.partials_per_pulse * npad * npad;
    plan.groups_per_pulse = (fuse && plan.reduce_in_cta) ? (plan.partials_per_pulse + 15) / 16 : 0;
    return plan;
}

// Copy / compute groups of the host-pointer pipeline (at most 8: one event each).
// Ensemble: pulse boundaries gb[0..ng].  Groups grow 1, 2, 4, ... units of 1/8 wave of warps: the first copy is the only
// one no kernel hides, later groups are long enough to hide theirs behind the group before; the last allowed group, or a
// remainder not worth a launch of its own, takes everything that is left.  Small ensembles get equal shares.
inline int ensemble_copy_groups(unsigned int batch, unsigned int unit, int G, unsigned int *gb) {
    int ng = 0;
    gb[0] = 0;
    if (G < 1) G = 1;
    if (unit < 1) unit = 1;
    if (batch < 4 * unit) {   // less than half a wave: equal shares (the chain kernel then splits pulses over several warps)
        const unsigned int bg = (batch + G - 1) / G;
        for (unsigned int b0 = 0; b0 < batch; b0 += bg) gb[++ng] = std::min(batch, b0 + bg);
        return ng;
    }
    for (unsigned int size = unit; ng < G && gb[ng] < batch; size *= 2) {
        const unsigned int left = batch - gb[ng];
        const unsigned int take = (ng == G - 1 || left < size + unit) ? left : size;
        gb[ng + 1] = gb[ng] + take;
        ++ng;
    }
    return ng;
}

// Single pulse: step boundaries bound[0..G].  Group sizes double (1 : 2 : 4 : ...): the copy of the first group is the only one no
// kernel hides, so it is small (1 / (2^G - 1) of the pulse), and since the staged host-to-device copy runs at more than twice the
// kernels' consumption rate (measured: 22 GB/s against 16 MB per 1.8 ms at C2), the copy of group g + 1 -- twice the data -- still
// finishes under the kernel of group g.
inline void time_copy_groups(unsigned long long nsteps, int G, unsigned long long *bound) {
    bound[0] = 0;
    if (G <= 1) { bound[1] = nsteps; return; }
    const unsigned long long units = (1ull << G) - 1;
    for (int g = 1; g <= G; ++g) bound[g] = (unsigned long long)((long double)nsteps * (long double)((1ull << g) - 1) / (long double)units);
    bound[G] = nsteps;
}

// Single-process multi-GPU: fewest effective steps worth a device of its own, and how many of `configured` devices take
// part in a call of `batch` pulses with `nsteps` steps each.
inline unsigned long long min_steps_per_device(int npad) {
    const unsigned long long np3 = (unsigned long long)npad * npad * npad;
    return std::max<unsigned long long>(64, (1ull << 26) / std::max<unsigned long long>(np3, 1));
}
inline unsigned int devices_for_call(unsigned int configured, unsigned int batch, unsigned long long nsteps, int npad) {
    unsigned long long g = std::min<unsigned long long>(configured, nsteps * batch / min_steps_per_device(npad));
    if (batch > 1) g = std::min<unsigned long long>(g, batch);
    return (unsigned int)std::max<unsigned long long>(g, 1);
}

// Hermitian-output batched GEMM (k4_gemm.cu, GemmArgs::herm): tile (i, j) of BM x BN elements lies strictly below the diagonal -- and is
// left to the mirror writes of the tile that holds its transpose -- iff its first row is beyond its last column.  The launch
// enumerates the remaining tiles row by row.
#ifdef __CUDACC__
#define PB_HD __host__ __device__
#else
#define PB_HD
#endif
PB_HD inline bool herm_tile_skipped(int BM, int BN, int i, int j) { return BM * i >= BN * (j + 1); }
PB_HD inline int herm_tile_count(int BM, int BN, int n) {
    int cnt = 0;
    for (int i = 0; i < n / BM; ++i) cnt += n / BN - (BM * i) / BN;
    return cnt;
}
PB_HD inline void herm_tile_at(int BM, int BN, int n, int idx, int &i, int &j) {
    const int tiles_n = n / BN;
    for (i = 0;; ++i) {
        const int first = (BM * i) / BN, cnt = tiles_n - first;
        if (idx < cnt) { j = first + idx; return; }
        idx -= cnt;
    }
}

}  // namespace pb
