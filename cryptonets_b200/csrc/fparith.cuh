// Exact modular arithmetic on the FP64 pipe for moduli below 2^50 (see the note in ntt.cu and DESIGN.md section 4).
// All values are integers held in doubles; every routine is exact as long as its operands stay below 2^52 in magnitude.
#pragma once
#include "modarith.cuh"

namespace cnhe {

constexpr double FP_MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adding and subtracting it rounds to the nearest integer
constexpr double FP_TWO52 = 4503599627370496.0;
__device__ __forceinline__ double u2d(u64 x) { return __dsub_rn(__longlong_as_double((long long)(x | 0x4330000000000000ULL)), FP_TWO52); }
__device__ __forceinline__ u64 d2u(double r) { return (u64)__double_as_longlong(__dadd_rn(r, FP_TWO52)) & 0x000FFFFFFFFFFFFFULL; }
__device__ __forceinline__ double fmodmul(double a, double w, double p, double pinv) {
    const double h = __dmul_rn(a, w);
    const double l = __fma_rn(a, w, -h);
    const double q = __dsub_rn(__fma_rn(h, pinv, FP_MAGIC), FP_MAGIC);
    return __dadd_rn(__fma_rn(-q, p, h), l);
}
__device__ __forceinline__ double frecenter(double x, double p, double pinv) {
    const double q = __dsub_rn(__fma_rn(x, pinv, FP_MAGIC), FP_MAGIC);
    return __fma_rn(-q, p, x);
}
__device__ __forceinline__ double fcanon(double x, double p, double pinv) { // any |x| < 2^52 -> [0, p)
    double r = frecenter(x, p, pinv);
    r = r < 0.0 ? __dadd_rn(r, p) : r;
    return r >= p ? __dsub_rn(r, p) : r;
}


// canonical u64 residue of (sum) for a double holding any integer |x| < 2^52
__device__ __forceinline__ u64 fcanon_u(double x, double p, double pinv) { return d2u(fcanon(x, p, pinv)); }

} // namespace cnhe
