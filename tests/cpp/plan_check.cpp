// CPU checker for parament_b200/csrc/plan.hpp: prints the partitions as JSON lines for tests/test_plan.py.
//   plan_check k1 npad batch nsteps num_sms horner [fuse]
//   plan_check egroups batch unit G
//   plan_check tgroups nsteps G
//   plan_check devices configured batch nsteps npad
//   plan_check herm BM BN n
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "plan.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    using namespace pb;
    if (!strcmp(argv[1], "k1") && (argc == 7 || argc == 8)) {
        const bool fuse = argc == 8 && atoi(argv[7]) != 0;
        const K1Plan p = plan_k1(atoi(argv[2]), (unsigned)atoll(argv[3]), strtoull(argv[4], nullptr, 10), atoi(argv[5]), atoi(argv[6]) != 0, fuse);
        printf("{\"grid\": %u, \"chunks_per_pulse\": %u, \"partials_per_pulse\": %u, \"reduce_in_cta\": %d, \"k3_warps\": %d, "
               "\"ctas_per_sm\": %d, \"partial_elems\": %zu, \"groups_per_pulse\": %u, \"warp_slots\": %u, \"mid_elems\": %zu, \"k3_launches\": %d}\n",
               p.grid, p.chunks_per_pulse, p.partials_per_pulse, p.reduce_in_cta, p.k3_warps, p.ctas_per_sm, p.partial_elems, p.groups_per_pulse,
               k1_warp_slots(atoi(argv[2]), atoi(argv[5]), atoi(argv[6]) != 0),
               k3_mid_elems(atoi(argv[2]), (unsigned)atoll(argv[3]), p.partials_per_pulse), k3_launches(p.partials_per_pulse));
        return 0;
    }
    if (!strcmp(argv[1], "egroups") && argc == 5) {
        unsigned int gb[9];
        const int ng = ensemble_copy_groups((unsigned)atoll(argv[2]), (unsigned)atoll(argv[3]), atoi(argv[4]), gb);
        printf("[");
        for (int g = 0; g <= ng; ++g) printf("%s%u", g ? ", " : "", gb[g]);
        printf("]\n");
        return 0;
    }
    if (!strcmp(argv[1], "tgroups") && argc == 4) {
        unsigned long long b[9];
        const int G = atoi(argv[3]);
        time_copy_groups(strtoull(argv[2], nullptr, 10), G, b);
        printf("[");
        for (int g = 0; g <= (G < 1 ? 1 : G); ++g) printf("%s%llu", g ? ", " : "", b[g]);
        printf("]\n");
        return 0;
    }
    if (!strcmp(argv[1], "devices") && argc == 6) {
        printf("{\"devices\": %u, \"min_steps\": %llu}\n",
               devices_for_call((unsigned)atoll(argv[2]), (unsigned)atoll(argv[3]), strtoull(argv[4], nullptr, 10), atoi(argv[5])),
               min_steps_per_device(atoi(argv[5])));
        return 0;
    }
    if (!strcmp(argv[1], "herm") && argc == 5) {   // plan_check herm BM BN n -> the enumerated tiles and the skipped ones
        const int BM = atoi(argv[2]), BN = atoi(argv[3]), n = atoi(argv[4]);
        const int cnt = herm_tile_count(BM, BN, n);
        printf("{\"count\": %d, \"tiles\": [", cnt);
        for (int t = 0; t < cnt; ++t) {
            int i = -1, j = -1;
            herm_tile_at(BM, BN, n, t, i, j);
            printf("%s[%d, %d]", t ? ", " : "", i, j);
        }
        printf("], \"skipped\": [");
        bool first = true;
        for (int i = 0; i < n / BM; ++i)
            for (int j = 0; j < n / BN; ++j)
                if (herm_tile_skipped(BM, BN, i, j)) { printf("%s[%d, %d]", first ? "" : ", ", i, j); first = false; }
        printf("]}\n");
        return 0;
    }
    return 2;
}
