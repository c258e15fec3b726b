// K5/K6 element-wise stages of the BFV ciphertext x ciphertext multiply and of key switching (sm_100a).
//
// Replaces, stage by stage, SEAL 3.2 Evaluator::bfv_multiply and util::BaseConverter::{fastbconv_mtilde, mont_rq,
// fast_floor, fastbconv_sk}, the 128-bit lazy inner product of Evaluator::relinearize_one_step / apply_galois, and the
// scale-and-round of Decryptor::decrypt -- reached from /root/reference "HE Wrapper/AtomicSealBfvVector.cs"
// :461-462,:546-547,:786-787,:839-840 (Multiply+Relinearize), every Rotate* call site, and :1042,:1085 (Decrypt).
// These kernels are per-coefficient (one thread owns one coefficient index across all RNS residues), fully
// coalesced along the coefficient axis, HBM-bound, and keep the per-context constants in shared memory.
#include "kernels.h"

namespace cnhe {

__device__ __forceinline__ void load_consts(BehzConst *dst, const BehzConst *src) {
    const int words = sizeof(BehzConst) / 8;
    const u64 *s = reinterpret_cast<const u64 *>(src);
    u64 *d = reinterpret_cast<u64 *>(dst);
    for (int i = threadIdx.x; i < words; i += blockDim.x) d[i] = s[i];
    __syncthreads();
}
static_assert(sizeof(BehzConst) % 8 == 0, "BehzConst must be a whole number of words");

constexpr u64 MT_MASK = 0xffffffffULL;
constexpr u64 M_TILDE = 1ULL << 32;

// ---- fastbconv_mtilde + mont_rq: q -> Bsk, with the q residues copied in front ("together" layout)
__global__ void __launch_bounds__(256) k_behz_lift(const u64 *const *__restrict__ ct_ptrs, u64 *__restrict__ out, int n_polys, int logn,
                                                  const BehzConst *__restrict__ gbc) {
    __shared__ BehzConst bc;
    load_consts(&bc, gbc);
    const int N = 1 << logn, k = bc.k, kb = bc.kb, kt = k + kb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n_polys << logn) return;
    const int x = (int)(gid & (N - 1)), poly = (int)(gid >> logn);
    const u64 *src = ct_ptrs[poly >> 1] + (size_t)(poly & 1) * k * N + x;
    u64 *dst = out + (size_t)poly * kt * N + x;
    u64 tmp[KMAX];
    u64 sm = 0;
#pragma unroll
    for (int i = 0; i < KMAX; i++) {
        if (i < k) {
            u64 v = src[(size_t)i * N];
            dst[(size_t)i * N] = v;
            tmp[i] = mulmod(v, bc.mtilde_inv_qhat_mod_q[i], bc.q[i]);
            sm += tmp[i] * bc.qhat_mod_mtilde[i];
        }
    }
    sm &= MT_MASK;
    const u64 r = (M_TILDE - ((sm * bc.inv_q_mod_mtilde) & MT_MASK)) & MT_MASK;
    for (int j = 0; j < kb; j++) {
        const DMod bj = bc.bsk[j];
        U128 acc = {0, 0};
#pragma unroll
        for (int i = 0; i < KMAX; i++)
            if (i < k) mac128(acc, tmp[i], bc.qhat_mod_bsk[j][i]);
        u64 xb = barrett128(acc, bj);
        u64 rr = r;
        if (bc.centered_mtilde && r >= (M_TILDE >> 1)) rr = r + (bj.p - M_TILDE);
        U128 t = mul64wide(bc.q_mod_bsk[j], rr);
        add128(t, xb);
        u64 v = barrett128(t, bj);
        dst[(size_t)(k + j) * N] = mulmod(v, bc.inv_mtilde_mod_bsk[j], bj);
    }
}

// ---- tensor product in the NTT domain, all 2k+1 residues
__global__ void __launch_bounds__(256) k_behz_tensor(const u64 *a, const u64 *b, u64 *__restrict__ d, int n, int logn,
                                                    const BehzConst *__restrict__ gbc) {
    __shared__ BehzConst bc;
    load_consts(&bc, gbc);
    const int N = 1 << logn, k = bc.k, kb = bc.kb, kt = k + kb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= ((size_t)n * kt) << logn) return;
    const int x = (int)(gid & (N - 1));
    const int l = (int)((gid >> logn) % kt), c = (int)((gid >> logn) / kt);
    const DMod m = l < k ? bc.q[l] : bc.bsk[l - k];
    const size_t in0 = ((size_t)(c * 2 + 0) * kt + l) * N + x, in1 = ((size_t)(c * 2 + 1) * kt + l) * N + x;
    const u64 a0 = a[in0], a1 = a[in1];
    u64 d0, d1, d2;
    if (a == b) {
        d0 = mulmod(a0, a0, m);
        d2 = mulmod(a1, a1, m);
        u64 cross = mulmod(a0, a1, m);
        d1 = addmod(cross, cross, m.p);
    } else {
        const u64 b0 = b[in0], b1 = b[in1];
        d0 = mulmod(a0, b0, m);
        d2 = mulmod(a1, b1, m);
        d1 = addmod(mulmod(a0, b1, m), mulmod(a1, b0, m), m.p);
    }
    const size_t o = ((size_t)(c * 3) * kt + l) * N + x;
    d[o] = d0;
    d[o + (size_t)kt * N] = d1;
    d[o + (size_t)2 * kt * N] = d2;
}

// ---- times t, fast_floor (q u Bsk -> Bsk), fastbconv_sk (Bsk -> q)
__global__ void __launch_bounds__(256) k_behz_floor(const u64 *__restrict__ d, u64 *__restrict__ out, int n_polys, u64 t, int logn,
                                                   const BehzConst *__restrict__ gbc) {
    __shared__ BehzConst bc;
    load_consts(&bc, gbc);
    const int N = 1 << logn, k = bc.k, kb = bc.kb, kt = k + kb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n_polys << logn) return;
    const int x = (int)(gid & (N - 1)), poly = (int)(gid >> logn);
    const u64 *src = d + (size_t)poly * kt * N + x;
    u64 *dst = out + (size_t)poly * k * N + x;
    const int na = kb - 1; // auxiliary primes (the base B); bsk[na] is m_sk
    u64 tmp[KBMAX], fl[KBMAX];
#pragma unroll
    for (int i = 0; i < KMAX; i++)
        if (i < k) {
            u64 v = mulmod(src[(size_t)i * N], t, bc.q[i]); // t < q_i is enforced at context creation
            tmp[i] = mulmod(v, bc.inv_qhat_mod_q[i], bc.q[i]);
        }
#pragma unroll
    for (int j = 0; j < KBMAX; j++)
        if (j < kb) {
            const DMod bj = bc.bsk[j];
            U128 acc = {0, 0};
#pragma unroll
            for (int i = 0; i < KMAX; i++)
                if (i < k) mac128(acc, tmp[i], bc.qhat_mod_bsk[j][i]);
            u64 conv = barrett128(acc, bj);
            u64 xb = mulmod(src[(size_t)(k + j) * N], t, bj);
            fl[j] = mulmod(xb + (bj.p - conv), bc.inv_q_mod_bsk[j], bj);
        }
    const DMod msk = bc.bsk[na];
#pragma unroll
    for (int j = 0; j < KBMAX; j++)
        if (j < na) tmp[j] = mulmod(fl[j], bc.inv_bhat_mod_b[j], bc.bsk[j]);
    U128 am = {0, 0};
#pragma unroll
    for (int j = 0; j < KBMAX; j++)
        if (j < na) mac128(am, tmp[j], bc.bhat_mod_msk[j]);
    const u64 alpha = mulmod(barrett128(am, msk) + (msk.p - fl[na]), bc.inv_B_mod_msk, msk);
    const bool neg = alpha > (msk.p >> 1);
    for (int i = 0; i < k; i++) {
        const DMod qi = bc.q[i];
        U128 acc = {0, 0};
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j < na) mac128(acc, tmp[j], bc.bhat_mod_q[i][j]);
        u64 v = barrett128(acc, qi);
        U128 c = neg ? mul64wide(bc.B_mod_q[i], msk.p - alpha) : mul64wide(qi.p - bc.B_mod_q[i], alpha);
        add128(c, v);
        dst[(size_t)i * N] = barrett128(c, qi);
    }
}

// ---- key-switch inner product: acc{0,1}[c][l][x] = sum_d digits[c][l][d][x] * key[d][{0,1}][l][x]
__global__ void __launch_bounds__(256) k_ks_mac(const u64 *__restrict__ digits, const u64 *__restrict__ key, u64 *__restrict__ acc, int n, int D,
                                               int logn, const BehzConst *__restrict__ gbc) {
    __shared__ BehzConst bc;
    load_consts(&bc, gbc);
    const int N = 1 << logn, k = bc.k;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= ((size_t)n * k) << logn) return;
    const int x = (int)(gid & (N - 1));
    const int l = (int)((gid >> logn) % k), c = (int)((gid >> logn) / k);
    const DMod m = bc.q[l];
    const u64 *dg = digits + ((size_t)c * k + l) * D * N + x;
    const u64 *k0 = key + (size_t)l * N + x;
    const size_t dstride = (size_t)N, kpoly = (size_t)k * N, kstride = (size_t)2 * k * N;
    U128 a0 = {0, 0}, a1 = {0, 0};
    for (int d0 = 0; d0 < D; d0 += 8) { // at most 8 products of 62x62 bits between reductions
        U128 s0 = {0, 0}, s1 = {0, 0};
        const int dend = min(D, d0 + 8);
        for (int dd = d0; dd < dend; dd++) {
            const u64 v = dg[(size_t)dd * dstride];
            mac128(s0, v, __ldg(k0 + (size_t)dd * kstride));
            mac128(s1, v, __ldg(k0 + (size_t)dd * kstride + kpoly));
        }
        add128(a0, barrett128(s0, m));
        add128(a1, barrett128(s1, m));
    }
    const size_t o = ((size_t)(c * 2) * k + l) * N + x;
    acc[o] = barrett128(a0, m);
    acc[o + (size_t)k * N] = barrett128(a1, m);
}

// ---- Decryptor::decrypt scale-and-round through {t, gamma}
__global__ void __launch_bounds__(256) k_decrypt_round(const u64 *__restrict__ xs, u64 *__restrict__ plain, int n, int logn,
                                                      const BehzConst *__restrict__ gbc, PlainConst pc) {
    __shared__ BehzConst bc;
    load_consts(&bc, gbc);
    const int N = 1 << logn, k = bc.k;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n << logn) return;
    const int x = (int)(gid & (N - 1)), c = (int)(gid >> logn);
    const u64 *src = xs + (size_t)c * k * N + x;
    U128 st = {0, 0}, sg = {0, 0};
    for (int i = 0; i < k; i++) {
        u64 v = mulmod(src[(size_t)i * N], pc.tgamma_mod_q[i], bc.q[i]);
        v = mulmod(v, bc.inv_qhat_mod_q[i], bc.q[i]);
        mac128(st, v, pc.qhat_mod_t[i]);
        mac128(sg, v, pc.qhat_mod_gamma[i]);
    }
    const u64 vt = mulmod(barrett128(st, pc.tmod), pc.neg_inv_q_mod_t, pc.tmod);
    const u64 vg = mulmod(barrett128(sg, pc.gmod), pc.neg_inv_q_mod_gamma, pc.gmod);
    u64 r;
    if (vg > (pc.gamma >> 1)) r = addmod(vt, reduce64(pc.gamma - vg, pc.tmod), pc.t);
    else r = submod(vt, reduce64(vg, pc.tmod), pc.t);
    plain[gid] = r ? mulmod(r, pc.inv_gamma_mod_t, pc.tmod) : 0;
}

static inline unsigned blocks_for(size_t threads) { return (unsigned)((threads + 255) / 256); }

cudaError_t launch_behz_lift(const u64 *const *ct_ptrs, u64 *out, int n, int logn, const BehzConst *bc, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_behz_lift<<<blocks_for((size_t)n * 2 << logn), 256, 0, s>>>(ct_ptrs, out, n * 2, logn, bc);
    return cudaGetLastError();
}
cudaError_t launch_behz_floor(const u64 *d, u64 *out3, int n, u64 t, int logn, const BehzConst *bc, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_behz_floor<<<blocks_for((size_t)n * 3 << logn), 256, 0, s>>>(d, out3, n * 3, t, logn, bc);
    return cudaGetLastError();
}
cudaError_t launch_behz_tensor(const u64 *a, const u64 *b, u64 *d, int n, int kt, int logn, const BehzConst *bc, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_behz_tensor<<<blocks_for(((size_t)n * kt) << logn), 256, 0, s>>>(a, b, d, n, logn, bc);
    return cudaGetLastError();
}
cudaError_t launch_ks_mac(const u64 *digits, const u64 *key, u64 *acc, int n, int D, int k, int logn, const BehzConst *bc, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_ks_mac<<<blocks_for(((size_t)n * k) << logn), 256, 0, s>>>(digits, key, acc, n, D, logn, bc);
    return cudaGetLastError();
}
cudaError_t launch_decrypt_round(const u64 *x, u64 *plain, int n, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s) {
    (void)k;
    if (n <= 0) return cudaSuccess;
    k_decrypt_round<<<blocks_for((size_t)n << logn), 256, 0, s>>>(x, plain, n, logn, bc, pc);
    return cudaGetLastError();
}

} // namespace cnhe
