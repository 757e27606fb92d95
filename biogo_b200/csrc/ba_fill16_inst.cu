// ba_fill16_inst.cu -- one translation unit per (kind, linear/affine) family of the packed fill
// kernel, so that the ~45 template instances of a family compile in parallel with the others.
// Built nine times with -DBA_INST_KIND={0,1,2} -DBA_INST_LIN={0: affine, 1: linear, 2: affine with
// symmetric gap scores}; ba_api.cu picks kernels through
// the ba_pick_fill16_<kind>_<lin>() functions defined here.
#include "ba_fill16.cuh"

#ifndef BA_INST_KIND
#error "compile with -DBA_INST_KIND=0|1|2 -DBA_INST_LIN=0|1"
#endif

typedef void (*Fill16Fn)(Fill16Params);

namespace {
template <bool MULTI, bool GSUB>
Fill16Fn pick_c(int cols)
{
    constexpr int K = BA_INST_KIND;
    constexpr bool LIN = BA_INST_LIN == 1, SYM = BA_INST_LIN == 2;
    switch (cols) {
    case 1: return ba_fill16_kernel<K, 1, MULTI, GSUB, false, LIN, SYM>;
    case 2: return ba_fill16_kernel<K, 2, MULTI, GSUB, false, LIN, SYM>;
    case 3: return ba_fill16_kernel<K, 3, MULTI, GSUB, false, LIN, SYM>;
    case 4: return ba_fill16_kernel<K, 4, MULTI, GSUB, false, LIN, SYM>;
    case 5: return ba_fill16_kernel<K, 5, MULTI, GSUB, false, LIN, SYM>;
    case 6: return ba_fill16_kernel<K, 6, MULTI, GSUB, false, LIN, SYM>;
    case 7: return ba_fill16_kernel<K, 7, MULTI, GSUB, false, LIN, SYM>;
    case 10: return ba_fill16_kernel<K, 10, MULTI, GSUB, false, LIN, SYM>;   // 257..320 columns in one pass
    default: return ba_fill16_kernel<K, 8, MULTI, GSUB, false, LIN, SYM>;
    }
}
// LONG (inter-pass wavefront) instances exist for the two widest multi-pass column classes only: pairs with
// four or more passes always land there (ba_plan.h cols_for; the 10-column class is single-pass)
template <bool GSUB>
Fill16Fn pick_long(int cols)
{
    constexpr int K = BA_INST_KIND;
    constexpr bool LIN = BA_INST_LIN == 1, SYM = BA_INST_LIN == 2;
    if (cols == 7) return ba_fill16_kernel<K, 7, true, GSUB, true, LIN, SYM>;
    return ba_fill16_kernel<K, 8, true, GSUB, true, LIN, SYM>;
}
}  // namespace

#define BA_INST_CAT2(a, b, c) ba_pick_fill16_##a##_##b
#define BA_INST_CAT(a, b) BA_INST_CAT2(a, b, 0)
#define BA_INST_NAME BA_INST_CAT(BA_INST_KIND, BA_INST_LIN)

Fill16Fn BA_INST_NAME(int cols, bool multi, bool gsub, bool wavefront)
{
    if (wavefront) return gsub ? pick_long<true>(cols) : pick_long<false>(cols);
    if (gsub) return multi ? pick_c<true, true>(cols) : pick_c<false, true>(cols);
    return multi ? pick_c<true, false>(cols) : pick_c<false, false>(cols);
}
