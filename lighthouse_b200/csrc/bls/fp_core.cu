// fp_core.cu — the out-of-line Fp / Fp2 leaf primitives (fp_add, fp_sub, fp_mul, fp_pow_pm3d4, fp2_mul, fp2_sqr,
// fp2_add, fp2_sub, ...) in their own translation unit, linked as relocatable device code.
//
// Why a separate TU: with the bodies visible, nvcc 12.9's NVVM inter-procedural analysis mis-handles calls whose
// output aliases an input (fp_add(n, n, t) and friends) and merged live stack slots in map_to_curve_sswu
// (caught by tests/test_bls_stages_gpu.py; reproduction notes in DESIGN.md §7).  Keeping the leaf functions
// opaque to their callers forces conservative (correct) memory assumptions everywhere above.
#define LHB_FP_CORE_ONLY 1
#include "fp2.cuh"
