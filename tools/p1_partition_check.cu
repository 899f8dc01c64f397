// Host-side check of the first-step partition k_plan hands to k_poly1s (common.cuh: poly_first_step, poly_item_count,
// p1_item_blocks): for every frame length the items tile the whole blocks, fit the shared-memory key buffers, and together
// with poly_first_step_rest's share (left-over segments, Linear ends) cover every sample exactly once.  No GPU needed:
//   nvcc -std=c++17 -I atsc_b200/csrc -o /tmp/p1_partition_check tools/p1_partition_check.cu && /tmp/p1_partition_check
#include <cstdio>
#include <cstdint>
#include "common.cuh"
using namespace atsc;
int main() {
    long checked = 0;
    for (uint32_t N = 10000; N <= 131072; N += (N < 65536 ? 997 : 1)) {
        if (poly_first_step(N) != P1_STEP) { printf("N=%u: first step %u\n", N, poly_first_step(N)); return 1; }
        if (N < POLY_ITEM_MIN_LEN) continue;
        const uint32_t kreg = (N + 99) / 100, K = kreg + (((kreg - 1) * 100 != N - 1) ? 1 : 0), nblk = (K - 3) / 4;
        const uint32_t Q = poly_item_count(N);
        uint32_t expect = 0, covered = 0;
        for (uint32_t q = 0; q < Q; q++) {
            uint32_t lo, hi;
            p1_item_blocks(N, q, &lo, &hi);
            if (lo != expect || hi < lo) { printf("N=%u q=%u: blocks [%u, %u) do not continue at %u\n", N, q, lo, hi, expect); return 1; }
            expect = hi;
            const uint32_t nkeys = 4 * (hi - lo) + 1;
            if (nkeys + 2 > POLY_ITEM_KEYS + 2 || nkeys + 2 > (uint32_t)P1_T) { printf("N=%u q=%u: %u keys\n", N, q, nkeys); return 1; }
            if ((hi - lo) < P1_G) { printf("N=%u q=%u: fewer blocks than groups\n", N, q); return 1; }
            // keys 1 + 4 lo - 1 .. 4 hi + 2 must exist (the last may be the appended key K - 1)
            if (4 * hi + 2 > K - 1) { printf("N=%u q=%u: key %u beyond K=%u\n", N, q, 4 * hi + 2, K); return 1; }
            covered += 4 * (hi - lo) * 100;  // four regular segments per block
        }
        if (expect != nblk) { printf("N=%u: items end at block %u of %u\n", N, expect, nblk); return 1; }
        // the rest: left-over Catmull-Rom segments 4 nblk + 1 .. K - 3 (regular), segment 0, segments K - 2 .. and the last sample
        const uint32_t left = (K - 3 >= 4 * nblk + 1) ? (K - 3 - 4 * nblk) * 100 : 0;
        const uint32_t ends = 100 + (N - (K - 2) * 100);
        if (covered + left + ends != N) { printf("N=%u: %u + %u + %u samples\n", N, covered, left, ends); return 1; }
        checked++;
    }
    printf("ok %ld lengths\n", checked);
    return 0;
}
