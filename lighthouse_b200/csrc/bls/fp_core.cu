// fp_core.cu — the out-of-line Fp primitives (fp_add, fp_sub, fp_mul, fp_pow_pm3d4) in their own translation
// unit, linked as relocatable device code.
//
// Why a separate TU: with the bodies visible, nvcc 12.9's NVVM inter-procedural analysis mis-handles calls whose
// output aliases an input (fp_add(n, n, t) and friends) and merged live stack slots in map_to_curve_sswu
// (caught by tests/test_bls_stages_gpu.py; reproduction notes in DESIGN.md §"Toolchain notes").  Keeping these
// four leaf functions opaque to the callers forces conservative (correct) memory assumptions everywhere above.
#include "fp.cuh"
