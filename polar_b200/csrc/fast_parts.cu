// fast_parts.cu -- one quarter of the scl_fast_kernel variant table per compilation (-DPOLAR_PART=0..3); see
// fast_variants.cuh. Entry i of the concatenated table is variant i (POLAR_B200_FAST_VARIANT, POLAR_B200_INFO_KERNEL_KIND).
#include "fast_variants.cuh"

#if !POLAR_MINSUM
#ifndef POLAR_PART
#error "compile with -DPOLAR_PART=0..3"
#endif
#endif
#define POLAR_CAT2(a, b) a##b
#define POLAR_CAT(a, b) POLAR_CAT2(a, b)

// (log2 N, virtual top layers, first shared-memory layer, log2 lanes per codeword, warps/block, blocks/SM).
// pick_fast_variant() takes the first entry matching (n, lanes); POLAR_B200_FAST_VARIANT=<index> overrides.
#if POLAR_MINSUM
// opt-in fast arithmetic (min-sum check nodes, hardware-friendly path metric): the BASELINE.json block lengths only
extern const FastVariant kFastMsPart[] = {
    POLAR_FAST_TM(11, 3, 5, 5, 16, 1),  // N=2048 lists 17..32
    POLAR_FAST_TM(11, 3, 5, 4, 4, 4),   // lists 9..16
    POLAR_FAST_TM(11, 3, 5, 3, 4, 4),   // lists 5..8
    POLAR_FAST_TM(11, 3, 5, 2, 4, 4),   // lists 3..4
    POLAR_FAST_TM(11, 3, 5, 1, 4, 4),   // list 2
    POLAR_FAST_TM(11, 3, 5, 0, 4, 4),   // list 1
    POLAR_FAST_TM(9, 3, 4, 5, 20, 1),   // N=512 lists 17..32
    POLAR_FAST_TM(9, 3, 4, 4, 4, 5),
    POLAR_FAST_TM(9, 3, 4, 3, 4, 5),
    POLAR_FAST_TM(9, 3, 4, 2, 4, 5),
    POLAR_FAST_TM(9, 3, 4, 1, 4, 5),
    POLAR_FAST_TM(9, 3, 4, 0, 4, 5),
};
extern const int kFastMsPartN = (int)(sizeof(kFastMsPart) / sizeof(FastVariant));
#else
extern const FastVariant POLAR_CAT(kFastPart, POLAR_PART)[] = {
#if POLAR_PART == 0
    // N=2048: layer 3 in the HBM/L2 scratch, layer 4 in tensor memory, layers 5-6 shared, 7-11 registers; 16 warps/SM
    POLAR_FAST_TM(11, 3, 5, 5, 4, 4),  // 0: lists 17..32
    POLAR_FAST_TM(11, 3, 5, 4, 4, 4),  // 1: lists 9..16 (2 codewords per warp)
    POLAR_FAST_TM(11, 3, 5, 3, 4, 4),  // 2: lists 5..8  (4 codewords per warp)
    POLAR_FAST_TM(11, 3, 5, 2, 4, 4),  // 3: lists 3..4  (8 codewords per warp)
    POLAR_FAST_TM(11, 3, 5, 1, 4, 4),  // list 2 (16 codewords per warp)
    POLAR_FAST_TM(11, 3, 5, 0, 4, 4),  // list 1 = plain SC (32 codewords per warp, lane = codeword)
    // N=512: nothing per path leaves the SM: layer 3 in tensor memory, layer 4 shared, 5-9 registers; 20 warps/SM
    POLAR_FAST_TM(9, 3, 4, 5, 4, 5),   // 4: lists 17..32
    POLAR_FAST_TM(9, 3, 4, 4, 4, 5),   // 5: lists 9..16
    POLAR_FAST_TM(9, 3, 4, 3, 4, 5),   // 6: lists 5..8
    POLAR_FAST_TM(9, 3, 4, 2, 4, 5),   // 7: lists 3..4
    POLAR_FAST_TM(9, 3, 4, 1, 4, 5),   // list 2
    POLAR_FAST_TM(9, 3, 4, 0, 4, 5),   // list 1
#elif POLAR_PART == 1
    // other block lengths, lists 17..32
    POLAR_FAST_TM(10, 3, 5, 5, 4, 4),  // 8: N=1024: layer 3 scratch, layer 4 tensor memory, layer 5 shared
    POLAR_FAST_TM(12, 3, 6, 5, 4, 4),  // 9: N=4096: layers 3-4 scratch, layer 5 tensor memory, layers 6-7 shared
    POLAR_FAST(8, 3, 3, 5, 4, 4),      // 10: N=256
    // alternates without tensor memory (POLAR_B200_FAST_VARIANT=<index>)
    POLAR_FAST(11, 3, 5, 5, 4, 4),     // 11: N=2048 lists 17..32, layers 3-4 in the scratch
    POLAR_FAST(11, 3, 6, 5, 4, 5),     // 12: N=2048 lists 17..32, layers 3-5 in the scratch, 20 warps/SM
    POLAR_FAST(9, 3, 4, 5, 4, 5),      // 13: N=512 lists 17..32, layer 3 in the scratch
    // one block per SM; the warps of a sub-partition start every codeword together (shared L0 instruction cache).
    // Same placement as entries 0-5 / 6-11. Preferred by pick_fast_variant() unless POLAR_B200_SYNC=0.
    POLAR_FAST_TM(11, 3, 5, 5, 16, 1), // 18: N=2048 lists 17..32
    POLAR_FAST_TM(11, 3, 5, 4, 16, 1),
    POLAR_FAST_TM(11, 3, 5, 3, 16, 1),
    POLAR_FAST_TM(11, 3, 5, 2, 16, 1),
    POLAR_FAST_TM(11, 3, 5, 1, 16, 1),
    POLAR_FAST_TM(11, 3, 5, 0, 16, 1),
#elif POLAR_PART == 2
    POLAR_FAST_TM(9, 3, 4, 5, 20, 1),  // 24: N=512 lists 17..32
    POLAR_FAST_TM(9, 3, 4, 4, 20, 1),
    POLAR_FAST_TM(9, 3, 4, 3, 20, 1),
    POLAR_FAST_TM(9, 3, 4, 2, 20, 1),
    POLAR_FAST_TM(9, 3, 4, 1, 20, 1),
    POLAR_FAST_TM(9, 3, 4, 0, 20, 1),
    POLAR_FAST_TM(10, 3, 5, 5, 16, 1), // 30: N=1024 lists 17..32
    POLAR_FAST_TM(12, 3, 6, 5, 16, 1), // 31: N=4096 lists 17..32
    POLAR_FAST(8, 3, 3, 5, 16, 1),     // 32: N=256 lists 17..32
    // alternates (POLAR_B200_FAST_VARIANT=<index>)
    POLAR_FAST(11, 3, 6, 5, 20, 1),    // 33: N=2048 lists 17..32 without tensor memory, 20 warps/SM
    // other block lengths, lists 1..16 (2..32 codewords per warp), same placement as their list-32 entries
    POLAR_FAST_TM(10, 3, 5, 4, 4, 4),  // 34: N=1024 lists 9..16
    POLAR_FAST_TM(10, 3, 5, 3, 4, 4),
#else
    POLAR_FAST_TM(10, 3, 5, 2, 4, 4),
    POLAR_FAST_TM(10, 3, 5, 1, 4, 4),
    POLAR_FAST_TM(10, 3, 5, 0, 4, 4),
    POLAR_FAST_TM(12, 3, 6, 4, 4, 4),  // 39: N=4096 lists 9..16
    POLAR_FAST_TM(12, 3, 6, 3, 4, 4),
    POLAR_FAST_TM(12, 3, 6, 2, 4, 4),
    POLAR_FAST_TM(12, 3, 6, 1, 4, 4),
    POLAR_FAST_TM(12, 3, 6, 0, 4, 4),
    POLAR_FAST(8, 3, 3, 4, 4, 4),      // 44: N=256 lists 9..16
    POLAR_FAST(8, 3, 3, 3, 4, 4),
    POLAR_FAST(8, 3, 3, 2, 4, 4),
    POLAR_FAST(8, 3, 3, 1, 4, 4),
    POLAR_FAST(8, 3, 3, 0, 4, 4),
    // lists 1 and 2 at N=2048 with every per-path layer on the SM: layer 3 in tensor memory (256 columns per warp), layers
    // 4-6 in shared memory (28 KB per warp), 8 warps/SM (measured +12.6 % on list 1; the same placement lost 32 % at list 32
    // and changed nothing at list 4, profiles/r02_ab_round1_experiments.txt)
    POLAR_FAST_TM_SG(11, 3, 4, 0, 2, 8, 1),  // 49: list 1
    POLAR_FAST_TM_SG(11, 3, 4, 1, 2, 8, 1),  // 50: list 2
    // N=8192: layers 3-5 in the scratch, layer 6 in tensor memory, layers 7-8 shared, 9-13 registers
    POLAR_FAST_TM(13, 3, 7, 5, 16, 1), // 51: lists 17..32
    POLAR_FAST_TM(13, 3, 7, 2, 4, 4),  // 52: lists 3..4
    POLAR_FAST_TM(13, 3, 7, 0, 4, 4),  // 53: list 1
#endif
};
extern const int POLAR_CAT(kFastPartN, POLAR_PART) = (int)(sizeof(POLAR_CAT(kFastPart, POLAR_PART)) / sizeof(FastVariant));
#endif  // POLAR_MINSUM
