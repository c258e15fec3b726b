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


// exact signed integer value of a double holding an integer |x| < 2^51 (one FP64 add; the rest runs on the integer pipe)
__device__ __forceinline__ long long d2i(double r) { return __double_as_longlong(__dadd_rn(r, FP_MAGIC)) - 0x4338000000000000LL; }
// canonical u64 residue of a double holding any integer |x| < 2^52: re-centre on the FP64 pipe (3 instructions), fix the sign
// on the integer pipe -- half the FP64 work of fcanon() + d2u()
__device__ __forceinline__ u64 fcanon_u(double x, double p, double pinv) {
    long long v = d2i(frecenter(x, p, pinv)); // |v| <= p/2 (+1)
    v += (v >> 63) & (long long)p;
    return (u64)v;
}
// same for a value already known to lie in (-p, p)
__device__ __forceinline__ u64 fsmall_u(double r, u64 p) {
    long long v = d2i(r);
    v += (v >> 63) & (long long)p;
    return (u64)v;
}
// "lazy" residues: internal buffers between the kernels of one multiply / key switch hold IEEE doubles with integer values
// |x| <= A*p (A host-checked), so producers skip the canonicalisation and consumers skip the u64 -> double conversion
__device__ __forceinline__ double ld_lazy(const u64 *p) { return __longlong_as_double((long long)*p); }
__device__ __forceinline__ u64 lazy_bits(double x) { return (u64)__double_as_longlong(x); }

} // namespace cnhe
