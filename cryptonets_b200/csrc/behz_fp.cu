// FP64 variants of the BEHZ / key-switch element-wise kernels (behz.cu) for moduli below 2^50.
//
// Same functions, same canonical outputs, different pipe: every modular product here is by a constant or of two residues
// below 2^50, so it runs as the 6-instruction error-free FP64 product of fparith.cuh instead of a Barrett reduction of a
// 128-bit integer product (4 mul.hi.u64 + 6 mul.lo.u64, the slowest instructions on the B200 integer pipe).  Wherever SEAL's
// algorithm depends on the *representative* of a residue (the fast base conversions sum [x c]_p * c' over the integers),
// the canonical representative in [0,p) is formed first, exactly as in behz.cu.
#include "fparith.cuh"
#include "kernels.h"

namespace cnhe {

__device__ __forceinline__ void load_consts_f(BehzConstF *dst, const BehzConstF *src) {
    const int words = sizeof(BehzConstF) / 8;
    const u64 *s = reinterpret_cast<const u64 *>(src);
    u64 *d = reinterpret_cast<u64 *>(dst);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = s[i];
    __syncthreads();
}
static_assert(sizeof(BehzConstF) % 8 == 0, "BehzConstF must be a whole number of words");

__global__ void __launch_bounds__(256) k_behz_lift_fp(const u64 *const *__restrict__ ct_ptrs, u64 *__restrict__ out, int n_polys, int logn,
                                                     const BehzConstF *__restrict__ gf) {
    __shared__ BehzConstF F;
    load_consts_f(&F, gf);
    const int N = 1 << logn, k = F.k, kb = F.kb, kt = k + kb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n_polys << logn) return;
    const int x = (int)(gid & (N - 1)), poly = (int)(gid >> logn);
    const u64 *src = ct_ptrs[poly >> 1] + (size_t)(poly & 1) * k * N + x;
    u64 *dst = out + (size_t)poly * kt * N + x;
    double tmp[KMAX];
    u64 sm = 0;
#pragma unroll
    for (int i = 0; i < KMAX; i++)
        if (i < k) {
            const u64 v = src[(size_t)i * N];
            dst[(size_t)i * N] = v;
            tmp[i] = fcanon(fmodmul(u2d(v), F.mtilde_inv_qhat_mod_q[i], F.qd[i], F.qinv[i]), F.qd[i], F.qinv[i]);
            sm += d2u(tmp[i]) * F.qhat_mod_mtilde[i];
        }
    sm &= 0xffffffffULL;
    const u64 r = ((1ULL << 32) - ((sm * F.inv_q_mod_mtilde) & 0xffffffffULL)) & 0xffffffffULL;
    double rr = u2d(r);
    if (F.centered_mtilde && r >= (1ULL << 31)) rr -= 4294967296.0;
    for (int j = 0; j < kb; j++) {
        const double p = F.bd[j], pinv = F.binv[j];
        double acc = fmodmul(rr, F.q_mod_bsk[j], p, pinv);
#pragma unroll
        for (int i = 0; i < KMAX; i++)
            if (i < k) acc = __dadd_rn(acc, fmodmul(tmp[i], F.qhat_mod_bsk[j][i], p, pinv));
        dst[(size_t)(k + j) * N] = fcanon_u(fmodmul(acc, F.inv_mtilde_mod_bsk[j], p, pinv), p, pinv);
    }
}

__global__ void __launch_bounds__(256) k_behz_tensor_fp(const u64 *a, const u64 *b, u64 *__restrict__ d, int n, int logn,
                                                       const BehzConstF *__restrict__ gf) {
    __shared__ BehzConstF F;
    load_consts_f(&F, gf);
    const int N = 1 << logn, k = F.k, kt = k + F.kb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= ((size_t)n * kt) << logn) return;
    const int x = (int)(gid & (N - 1));
    const int l = (int)((gid >> logn) % kt), c = (int)((gid >> logn) / kt);
    const double p = l < k ? F.qd[l] : F.bd[l - k], pinv = l < k ? F.qinv[l] : F.binv[l - k];
    const size_t in0 = ((size_t)(c * 2 + 0) * kt + l) * N + x, in1 = ((size_t)(c * 2 + 1) * kt + l) * N + x;
    const double a0 = u2d(a[in0]), a1 = u2d(a[in1]);
    double d0, d1, d2;
    if (a == b) {
        d0 = fmodmul(a0, a0, p, pinv);
        d2 = fmodmul(a1, a1, p, pinv);
        const double cross = fmodmul(a0, a1, p, pinv);
        d1 = __dadd_rn(cross, cross);
    } else {
        const double b0 = u2d(b[in0]), b1 = u2d(b[in1]);
        d0 = fmodmul(a0, b0, p, pinv);
        d2 = fmodmul(a1, b1, p, pinv);
        d1 = __dadd_rn(fmodmul(a0, b1, p, pinv), fmodmul(a1, b0, p, pinv));
    }
    const size_t o = ((size_t)(c * 3) * kt + l) * N + x;
    d[o] = fcanon_u(d0, p, pinv);
    d[o + (size_t)kt * N] = fcanon_u(d1, p, pinv);
    d[o + (size_t)2 * kt * N] = fcanon_u(d2, p, pinv);
}

__global__ void __launch_bounds__(256) k_behz_floor_fp(const u64 *__restrict__ d, u64 *__restrict__ out, int n_polys, double t, int logn,
                                                      const BehzConstF *__restrict__ gf) {
    __shared__ BehzConstF F;
    load_consts_f(&F, gf);
    const int N = 1 << logn, k = F.k, kb = F.kb, kt = k + kb, na = kb - 1;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n_polys << logn) return;
    const int x = (int)(gid & (N - 1)), poly = (int)(gid >> logn);
    const u64 *src = d + (size_t)poly * kt * N + x;
    u64 *dst = out + (size_t)poly * k * N + x;
    double tmp[KBMAX], fl[KBMAX];
#pragma unroll
    for (int i = 0; i < KMAX; i++)
        if (i < k) {
            const double p = F.qd[i], pinv = F.qinv[i];
            const double v = fmodmul(u2d(src[(size_t)i * N]), t, p, pinv);
            tmp[i] = fcanon(fmodmul(v, F.inv_qhat_mod_q[i], p, pinv), p, pinv);
        }
#pragma unroll
    for (int j = 0; j < KBMAX; j++)
        if (j < kb) {
            const double p = F.bd[j], pinv = F.binv[j];
            double conv = 0.0;
#pragma unroll
            for (int i = 0; i < KMAX; i++)
                if (i < k) conv = __dadd_rn(conv, fmodmul(tmp[i], F.qhat_mod_bsk[j][i], p, pinv));
            const double xb = fmodmul(u2d(src[(size_t)(k + j) * N]), t, p, pinv);
            fl[j] = fmodmul(frecenter(__dsub_rn(xb, conv), p, pinv), F.inv_q_mod_bsk[j], p, pinv);
        }
    const double pm = F.bd[na], pminv = F.binv[na];
    double am = 0.0;
#pragma unroll
    for (int j = 0; j < KBMAX; j++)
        if (j < na) {
            tmp[j] = fcanon(fmodmul(fl[j], F.inv_bhat_mod_b[j], F.bd[j], F.binv[j]), F.bd[j], F.binv[j]);
            am = __dadd_rn(am, fmodmul(tmp[j], F.bhat_mod_msk[j], pm, pminv));
        }
    const double alpha = fcanon(fmodmul(frecenter(__dsub_rn(am, fl[na]), pm, pminv), F.inv_B_mod_msk, pm, pminv), pm, pminv);
    // centred alpha: alpha > m_sk/2 means alpha - m_sk (negative)
    const double alpha_c = alpha > F.msk_half ? __dsub_rn(alpha, pm) : alpha;
    for (int i = 0; i < k; i++) {
        const double p = F.qd[i], pinv = F.qinv[i];
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j < na) v = __dadd_rn(v, fmodmul(tmp[j], F.bhat_mod_q[i][j], p, pinv));
        v = __dsub_rn(v, fmodmul(alpha_c, F.B_mod_q[i], p, pinv));
        dst[(size_t)i * N] = fcanon_u(v, p, pinv);
    }
}

__global__ void __launch_bounds__(256) k_ks_mac_fp(const u64 *__restrict__ digits, const u64 *__restrict__ key, u64 *__restrict__ acc, int n, int D,
                                                  int logn, const BehzConstF *__restrict__ gf) {
    __shared__ BehzConstF F;
    load_consts_f(&F, gf);
    const int N = 1 << logn, k = F.k;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= ((size_t)n * k) << logn) return;
    const int x = (int)(gid & (N - 1));
    const int l = (int)((gid >> logn) % k), c = (int)((gid >> logn) / k);
    const double p = F.qd[l], pinv = F.qinv[l];
    const u64 *dg = digits + ((size_t)c * D * k + l) * N + x;
    const u64 *k0 = key + (size_t)l * N + x;
    const size_t dstride = (size_t)k * N, kstride = (size_t)2 * k * N;
    double a0 = 0.0, a1 = 0.0;
    for (int d0 = 0; d0 < D; d0 += 8) {
        const int dend = min(D, d0 + 8);
        for (int dd = d0; dd < dend; dd++) {
            const double v = u2d(dg[(size_t)dd * dstride]);
            a0 = __dadd_rn(a0, fmodmul(v, u2d(__ldg(k0 + (size_t)dd * kstride)), p, pinv));
            a1 = __dadd_rn(a1, fmodmul(v, u2d(__ldg(k0 + (size_t)dd * kstride + dstride)), p, pinv));
        }
        a0 = frecenter(a0, p, pinv);
        a1 = frecenter(a1, p, pinv);
    }
    const size_t o = ((size_t)(c * 2) * k + l) * N + x;
    acc[o] = fcanon_u(a0, p, pinv);
    acc[o + (size_t)k * N] = fcanon_u(a1, p, pinv);
}

static inline unsigned blocks_for(size_t threads) { return (unsigned)((threads + 255) / 256); }

cudaError_t launch_behz_lift_fp(const u64 *const *ct_ptrs, u64 *out, int n, int logn, const BehzConstF *f, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_behz_lift_fp<<<blocks_for((size_t)n * 2 << logn), 256, 0, s>>>(ct_ptrs, out, n * 2, logn, f);
    return cudaGetLastError();
}
cudaError_t launch_behz_tensor_fp(const u64 *a, const u64 *b, u64 *d, int n, int kt, int logn, const BehzConstF *f, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_behz_tensor_fp<<<blocks_for(((size_t)n * kt) << logn), 256, 0, s>>>(a, b, d, n, logn, f);
    return cudaGetLastError();
}
cudaError_t launch_behz_floor_fp(const u64 *d, u64 *out3, int n, u64 t, int logn, const BehzConstF *f, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_behz_floor_fp<<<blocks_for((size_t)n * 3 << logn), 256, 0, s>>>(d, out3, n * 3, (double)t, logn, f);
    return cudaGetLastError();
}
cudaError_t launch_ks_mac_fp(const u64 *digits, const u64 *key, u64 *acc, int n, int D, int k, int logn, const BehzConstF *f, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_ks_mac_fp<<<blocks_for(((size_t)n * k) << logn), 256, 0, s>>>(digits, key, acc, n, D, logn, f);
    return cudaGetLastError();
}

} // namespace cnhe
