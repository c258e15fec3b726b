// FP64 variants of the BEHZ / key-switch element-wise kernels (behz.cu) for moduli below 2^50.
//
// Same functions, same canonical outputs, different pipe: every modular product here is by a constant or of two residues
// below 2^50, so it runs as the 6-instruction error-free FP64 product of fparith.cuh instead of a Barrett reduction of a
// 128-bit integer product (4 mul.hi.u64 + 6 mul.lo.u64, the slowest instructions on the B200 integer pipe).  Wherever SEAL's
// algorithm depends on the *representative* of a residue (the fast base conversions sum [x c]_p * c' over the integers),
// the canonical representative in [0,p) is formed first, exactly as in behz.cu.
#include <cstdlib>
#include "fparith.cuh"
#include "kernels.h"

namespace cnhe {

// The conversion constants travel as a __grid_constant__ kernel parameter (3 KB, constant bank): with the loops over residues fully
// unrolled every constant is an immediate c[bank][offset] operand of its DFMA -- no shared-memory copy, no LDS per product.
static_assert(sizeof(BehzConstF) % 8 == 0 && sizeof(BehzConstF) <= 3584, "BehzConstF must fit the kernel parameter space");

// LAZY (all kernels below): buffers exchanged with the NTT kernels hold lazy doubles (fparith.cuh) instead of canonical words
template <bool LAZY>
__global__ void __launch_bounds__(256) k_behz_lift_fp(const u64 *const *__restrict__ ct_ptrs, u64 *__restrict__ out, int n_polys, int logn,
                                                     const __grid_constant__ BehzConstF F) {
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
            const double vd = u2d(v);
            dst[(size_t)i * N] = LAZY ? lazy_bits(vd) : v;
            tmp[i] = fcanon(fmodmul(vd, F.mtilde_inv_qhat_mod_q[i], F.qd[i], F.qinv[i]), F.qd[i], F.qinv[i]);
            sm += d2u(tmp[i]) * F.qhat_mod_mtilde[i];
        }
    sm &= 0xffffffffULL;
    const u64 r = ((1ULL << 32) - ((sm * F.inv_q_mod_mtilde) & 0xffffffffULL)) & 0xffffffffULL;
    double rr = u2d(r);
    if (F.centered_mtilde && r >= (1ULL << 31)) rr -= 4294967296.0;
#pragma unroll
    for (int j = 0; j < KBMAX; j++) {
        if (j >= kb) break;
        const double p = F.bd[j], pinv = F.binv[j];
        double acc = fmodmul(rr, F.q_mod_bsk[j], p, pinv);
#pragma unroll
        for (int i = 0; i < KMAX; i++)
            if (i < k) acc = __dadd_rn(acc, fmodmul(tmp[i], F.qhat_mod_bsk[j][i], p, pinv));
        const double r = fmodmul(acc, F.inv_mtilde_mod_bsk[j], p, pinv); // fresh product: |r| <= 0.51 p
        dst[(size_t)(k + j) * N] = LAZY ? lazy_bits(r) : fsmall_u(r, F.b_u[j]);
    }
}

template <bool LAZY>
__global__ void __launch_bounds__(256) k_behz_tensor_fp(const u64 *a, const u64 *b, u64 *__restrict__ d, int n, int logn,
                                                       const __grid_constant__ BehzConstF F) {
    const int N = 1 << logn, k = F.k, kt = k + F.kb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= ((size_t)n * kt) << logn) return;
    const int x = (int)(gid & (N - 1));
    const int l = (int)((gid >> logn) % kt), c = (int)((gid >> logn) / kt);
    const double p = l < k ? F.qd[l] : F.bd[l - k], pinv = l < k ? F.qinv[l] : F.binv[l - k];
    const size_t in0 = ((size_t)(c * 2 + 0) * kt + l) * N + x, in1 = ((size_t)(c * 2 + 1) * kt + l) * N + x;
    const double a0 = LAZY ? ld_lazy(a + in0) : u2d(a[in0]), a1 = LAZY ? ld_lazy(a + in1) : u2d(a[in1]);
    double d0, d1, d2;
    if (a == b) {
        d0 = fmodmul(a0, a0, p, pinv);
        d2 = fmodmul(a1, a1, p, pinv);
        const double cross = fmodmul(a0, a1, p, pinv);
        d1 = __dadd_rn(cross, cross);
    } else {
        const double b0 = LAZY ? ld_lazy(b + in0) : u2d(b[in0]), b1 = LAZY ? ld_lazy(b + in1) : u2d(b[in1]);
        d0 = fmodmul(a0, b0, p, pinv);
        d2 = fmodmul(a1, b1, p, pinv);
        d1 = __dadd_rn(fmodmul(a0, b1, p, pinv), fmodmul(a1, b0, p, pinv));
    }
    const size_t o = ((size_t)(c * 3) * kt + l) * N + x;
    if (LAZY) { // |d0|, |d2| <= 0.51 p, |d1| <= 1.02 p: within the inverse transform's input bound (1.25 p)
        d[o] = lazy_bits(d0);
        d[o + (size_t)kt * N] = lazy_bits(d1);
        d[o + (size_t)2 * kt * N] = lazy_bits(d2);
    } else {
        d[o] = fcanon_u(d0, p, pinv);
        d[o + (size_t)kt * N] = fcanon_u(d1, p, pinv);
        d[o + (size_t)2 * kt * N] = fcanon_u(d2, p, pinv);
    }
}

template <bool LAZY>
__global__ void __launch_bounds__(256) k_behz_floor_fp(const u64 *__restrict__ d, u64 *__restrict__ out, int n_polys, double t, int logn,
                                                      const __grid_constant__ BehzConstF F) {
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
            const double v = fmodmul(LAZY ? ld_lazy(src + (size_t)i * N) : u2d(src[(size_t)i * N]), t, p, pinv);
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
            const double xb = fmodmul(LAZY ? ld_lazy(src + (size_t)(k + j) * N) : u2d(src[(size_t)(k + j) * N]), t, p, pinv);
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
    double fl_sk = 0.0; // fl[na] without a runtime-indexed (local-memory) array access
#pragma unroll
    for (int j = 0; j < KBMAX; j++)
        if (j == na) fl_sk = fl[j];
    const double alpha = fcanon(fmodmul(frecenter(__dsub_rn(am, fl_sk), pm, pminv), F.inv_B_mod_msk, pm, pminv), pm, pminv);
    // centred alpha: alpha > m_sk/2 means alpha - m_sk (negative)
    const double alpha_c = alpha > F.msk_half ? __dsub_rn(alpha, pm) : alpha;
#pragma unroll
    for (int i = 0; i < KMAX; i++) {
        if (i >= k) break;
        const double p = F.qd[i], pinv = F.qinv[i];
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j < na) v = __dadd_rn(v, fmodmul(tmp[j], F.bhat_mod_q[i][j], p, pinv));
        v = __dsub_rn(v, fmodmul(alpha_c, F.B_mod_q[i], p, pinv));
        dst[(size_t)i * N] = fcanon_u(v, p, pinv);
    }
}

// fast_floor + fastbconv_sk with the constant factors folded into the conversion matrices (FloorConstF): -19 % FP64 instructions,
// -12 % time for the element-wise family.  Same outputs as k_behz_floor_fp: every folded product is the same residue class and the
// canonical representatives are formed at the same points.  ITERS > 1 walks several coefficients per thread with the next one's
// loads issued early -- measured 60 % slower (170 registers, one CTA per SM), so only ITERS = 1 is built.
template <int ITERS>
__global__ void __launch_bounds__(256) k_behz_floor_fold_fp(const u64 *__restrict__ d, u64 *__restrict__ out, size_t total, int logn,
                                                           const __grid_constant__ FloorConstF F) {
    const int N = 1 << logn, k = F.k, kb = F.kb, kt = k + kb, na = kb - 1;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double cq[KMAX], cb[KBMAX], nq[KMAX], nb[KBMAX]; // current / next coefficient: residues mod q_i and mod the Bsk primes
    auto load = [&](double (&vq)[KMAX], double (&vb)[KBMAX], size_t g) {
        const u64 *src = d + (g >> logn) * (size_t)kt * N + (g & (size_t)(N - 1));
#pragma unroll
        for (int i = 0; i < KMAX; i++)
            if (i < k) vq[i] = ld_lazy(src + (size_t)i * N);
        src += (size_t)k * N;
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j < kb) vb[j] = ld_lazy(src + (size_t)j * N);
    };
    if (gid < total) load(cq, cb, gid);
#pragma unroll 1
    for (int it = 0; it < ITERS; it++, gid += stride) {
        if (gid >= total) return;
        const bool more = ITERS > 1 && it + 1 < ITERS && gid + stride < total;
        if (more) load(nq, nb, gid + stride);
        double tmp[KBMAX];
        // [x t q-hat_i^-1]_{q_i}, canonical representative (the base conversion sums these integers)
#pragma unroll
        for (int i = 0; i < KMAX; i++)
            if (i < k) tmp[i] = fcanon(fmodmul(cq[i], F.xq[i], F.qd[i], F.qinv[i]), F.qd[i], F.qinv[i]);
        double fl[KBMAX];
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j < kb) {
                const double p = F.bd[j], pinv = F.binv[j];
                double acc = fmodmul(cb[j], F.xb[j], p, pinv);
#pragma unroll
                for (int i = 0; i < KMAX; i++)
                    if (i < k) acc = __dsub_rn(acc, fmodmul(tmp[i], F.conv[j][i], p, pinv));
                fl[j] = acc; // j < na: floor_j * B-hat_j^-1 mod p_j;  j = na: floor mod m_sk   (|acc| <= (k+1) * 0.51 p)
            }
        double pm = 0.0, pminv = 0.0, am = 0.0, fl_sk = 0.0;
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j == na) { pm = F.bd[j]; pminv = F.binv[j]; fl_sk = fl[j]; }
#pragma unroll
        for (int j = 0; j < KBMAX; j++)
            if (j < na) {
                tmp[j] = fcanon(fl[j], F.bd[j], F.binv[j]);
                am = __dadd_rn(am, fmodmul(tmp[j], F.bhat_mod_msk[j], pm, pminv));
            }
        const double alpha = fcanon(fmodmul(frecenter(__dsub_rn(am, fl_sk), pm, pminv), F.inv_B_mod_msk, pm, pminv), pm, pminv);
        const double alpha_c = alpha > F.msk_half ? __dsub_rn(alpha, pm) : alpha;
        u64 *dst = out + (gid >> logn) * (size_t)k * N + (gid & (size_t)(N - 1));
#pragma unroll
        for (int i = 0; i < KMAX; i++) {
            if (i >= k) break;
            const double p = F.qd[i], pinv = F.qinv[i];
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < KBMAX; j++)
                if (j < na) v = __dadd_rn(v, fmodmul(tmp[j], F.bhat_mod_q[i][j], p, pinv));
            v = __dsub_rn(v, fmodmul(alpha_c, F.B_mod_q[i], p, pinv));
            dst[(size_t)i * N] = fcanon_u(v, p, pinv);
        }
        if (more) {
#pragma unroll
            for (int i = 0; i < KMAX; i++) cq[i] = nq[i];
#pragma unroll
            for (int j = 0; j < KBMAX; j++) cb[j] = nb[j];
        }
    }
}

// One thread: two adjacent coefficients of residue l of one ciphertext, walking the D digit transforms (adjacent polynomials in the
// [c][l][d][N] layout) with 16-byte loads, UNR digits in flight: the kernel is bound by the digit stream out of HBM (8 MB per
// ciphertext at N=8192, D=25), the key (16 MB per channel) is re-read out of L2 by every ciphertext.
__device__ __forceinline__ ulonglong2 ldcs2(const u64 *p) { // streaming (evict-first): each digit word is read exactly once
    ulonglong2 v;
    asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
// CT ciphertexts per thread: the two key words of a (digit, residue, coefficient pair) are fetched from L2 once and used for all of
// them.  With one ciphertext per thread the kernel moved 48 bytes through L2 for every 16 bytes of the digit stream and sat on the L2
// bandwidth (3.9 TB/s of HBM reads = 11.7 TB/s out of L2); with four it is 24 bytes.
template <bool LAZY, int UNR, int CT>
__global__ void __launch_bounds__(256) k_ks_mac_fp(const u64 *__restrict__ digits, const u64 *__restrict__ key, u64 *__restrict__ acc, int n, int D,
                                                  int logn, const __grid_constant__ BehzConstF F) {
    const int N = 1 << logn, k = F.k;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_groups = (n + CT - 1) / CT;
    if (gid >= ((size_t)n_groups * k) << (logn - 1)) return;
    const int x = (int)(gid & (N / 2 - 1)) * 2;
    const int l = (int)((gid >> (logn - 1)) % k), c0 = (int)((gid >> (logn - 1)) / k) * CT;
    const double p = F.qd[l], pinv = F.qinv[l];
    const size_t kpoly = (size_t)k * N, kstride = (size_t)2 * k * N, cstride = (size_t)k * D * N;
    const u64 *dg = digits + ((size_t)c0 * k + l) * D * N + x;
    const u64 *k0 = key + (size_t)l * N + x;
    double a[CT][4]; // [ciphertext][key poly * 2 + coefficient]
#pragma unroll
    for (int ci = 0; ci < CT; ci++) a[ci][0] = a[ci][1] = a[ci][2] = a[ci][3] = 0.0;
#pragma unroll UNR
    for (int dd = 0; dd < D; dd++) {
        ulonglong2 vu[CT];
#pragma unroll
        for (int ci = 0; ci < CT; ci++)
            vu[ci] = c0 + ci < n ? ldcs2(dg + (size_t)ci * cstride + (size_t)dd * N) : make_ulonglong2(0, 0);
        const ulonglong2 w0u = __ldg(reinterpret_cast<const ulonglong2 *>(k0 + (size_t)dd * kstride));
        const ulonglong2 w1u = __ldg(reinterpret_cast<const ulonglong2 *>(k0 + (size_t)dd * kstride + kpoly));
        const double w00 = u2d(w0u.x), w01 = u2d(w0u.y), w10 = u2d(w1u.x), w11 = u2d(w1u.y);
#pragma unroll
        for (int ci = 0; ci < CT; ci++) {
            const double v0 = LAZY ? __longlong_as_double((long long)vu[ci].x) : u2d(vu[ci].x);
            const double v1 = LAZY ? __longlong_as_double((long long)vu[ci].y) : u2d(vu[ci].y);
            a[ci][0] = __dadd_rn(a[ci][0], fmodmul(v0, w00, p, pinv));
            a[ci][1] = __dadd_rn(a[ci][1], fmodmul(v1, w01, p, pinv));
            a[ci][2] = __dadd_rn(a[ci][2], fmodmul(v0, w10, p, pinv));
            a[ci][3] = __dadd_rn(a[ci][3], fmodmul(v1, w11, p, pinv));
        }
        if ((dd & 7) == 7) { // sums of 8 fresh products stay below 4.1 p; re-centre before they could leave the exact range
#pragma unroll
            for (int ci = 0; ci < CT; ci++)
#pragma unroll
                for (int j = 0; j < 4; j++) a[ci][j] = frecenter(a[ci][j], p, pinv);
        }
    }
#pragma unroll
    for (int ci = 0; ci < CT; ci++) {
        if (c0 + ci >= n) break;
#pragma unroll
        for (int j = 0; j < 4; j++) a[ci][j] = frecenter(a[ci][j], p, pinv);
        const size_t o = ((size_t)((c0 + ci) * 2) * k + l) * N + x;
        if (LAZY) {
            *reinterpret_cast<ulonglong2 *>(acc + o) = make_ulonglong2(lazy_bits(a[ci][0]), lazy_bits(a[ci][1]));
            *reinterpret_cast<ulonglong2 *>(acc + o + kpoly) = make_ulonglong2(lazy_bits(a[ci][2]), lazy_bits(a[ci][3]));
        } else {
            *reinterpret_cast<ulonglong2 *>(acc + o) = make_ulonglong2(fsmall_u(a[ci][0], F.q_u[l]), fsmall_u(a[ci][1], F.q_u[l]));
            *reinterpret_cast<ulonglong2 *>(acc + o + kpoly) = make_ulonglong2(fsmall_u(a[ci][2], F.q_u[l]), fsmall_u(a[ci][3], F.q_u[l]));
        }
    }
}

// ---- the same inner product with the operands staged by the copy engine (cp.async.bulk -> shared memory ring, mbarriers): the register
// version above is bound by load latency (ncu: long_scoreboard 10 stall cycles per issue, 49 % of DRAM bandwidth) -- it can keep only as
// many bytes in flight as it has registers for.  Here one producer thread streams, per digit, the tile's slice of CT digit polynomials
// (HBM) and of the two key polynomials (L2) into a four-stage ring; 256 consumer threads (2 coefficients each) run the identical
// arithmetic in the identical order out of shared memory.  CTA = (group of CT ciphertexts, residue l, 512 coefficients).
constexpr int KT_X = 512, KT_STAGES = 4, KT_CONSUMERS = 256, KT_THREADS = KT_CONSUMERS + 32;
__device__ __forceinline__ unsigned kt_sptr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kt_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = kt_sptr(bar);
    unsigned done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && spin > (1u << 28)) __trap(); // a protocol error fails the launch instead of hanging the GPU
    }
}
template <int CT>
__global__ void __launch_bounds__(KT_THREADS, 2) k_ks_mac_tma(const u64 *__restrict__ digits, const u64 *__restrict__ key, u64 *__restrict__ acc, int n, int D,
                                                             int logn, const __grid_constant__ BehzConstF F) {
    extern __shared__ __align__(128) unsigned char kt_smem[];
    constexpr int STAGE_BYTES = (CT + 2) * KT_X * 8;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(kt_smem + KT_STAGES * STAGE_BYTES);
    unsigned long long *full = bars, *empty = bars + KT_STAGES;
    const int N = 1 << logn, k = F.k, tid = threadIdx.x;
    const int tiles = N / KT_X;
    const int tile = blockIdx.x % tiles, l = (blockIdx.x / tiles) % k, c0 = (blockIdx.x / (tiles * k)) * CT;
    const int x0 = tile * KT_X;
    const size_t kpoly = (size_t)k * N, kstride = (size_t)2 * k * N, cstride = (size_t)k * D * N;
    const int n_here = min(CT, n - c0);
    if (tid == 0) {
        for (int i = 0; i < KT_STAGES; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_sptr(full + i)), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kt_sptr(empty + i)), "r"((unsigned)KT_CONSUMERS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= KT_CONSUMERS) {
        if (tid == KT_CONSUMERS) { // producer
            const u64 *dg = digits + ((size_t)c0 * k + l) * D * N + x0;
            const u64 *k0 = key + (size_t)l * N + x0;
            for (int dd = 0; dd < D; dd++) {
                const int s = dd % KT_STAGES;
                kt_wait(empty + s, (((unsigned)dd / KT_STAGES) & 1) ^ 1);
                unsigned char *st = kt_smem + s * STAGE_BYTES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kt_sptr(full + s)), "r"((unsigned)((n_here + 2) * KT_X * 8)) : "memory");
                for (int ci = 0; ci < n_here; ci++)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kt_sptr(st + ci * KT_X * 8)),
                                 "l"(dg + (size_t)ci * cstride + (size_t)dd * N), "r"((unsigned)(KT_X * 8)), "r"(kt_sptr(full + s))
                                 : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kt_sptr(st + CT * KT_X * 8)),
                             "l"(k0 + (size_t)dd * kstride), "r"((unsigned)(KT_X * 8)), "r"(kt_sptr(full + s))
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kt_sptr(st + (CT + 1) * KT_X * 8)),
                             "l"(k0 + (size_t)dd * kstride + kpoly), "r"((unsigned)(KT_X * 8)), "r"(kt_sptr(full + s))
                             : "memory");
            }
        }
        return;
    }
    const double p = F.qd[l], pinv = F.qinv[l];
    double a[CT][4]; // [ciphertext][key poly * 2 + coefficient]
#pragma unroll
    for (int ci = 0; ci < CT; ci++) a[ci][0] = a[ci][1] = a[ci][2] = a[ci][3] = 0.0;
    for (int dd = 0; dd < D; dd++) {
        const int s = dd % KT_STAGES;
        kt_wait(full + s, ((unsigned)dd / KT_STAGES) & 1);
        const ulonglong2 *st = reinterpret_cast<const ulonglong2 *>(kt_smem + s * STAGE_BYTES);
        const ulonglong2 w0u = st[CT * (KT_X / 2) + tid], w1u = st[(CT + 1) * (KT_X / 2) + tid];
        const double w00 = u2d(w0u.x), w01 = u2d(w0u.y), w10 = u2d(w1u.x), w11 = u2d(w1u.y);
#pragma unroll
        for (int ci = 0; ci < CT; ci++) {
            if (ci < n_here) {
                const ulonglong2 vu = st[ci * (KT_X / 2) + tid];
                const double v0 = __longlong_as_double((long long)vu.x), v1 = __longlong_as_double((long long)vu.y);
                a[ci][0] = __dadd_rn(a[ci][0], fmodmul(v0, w00, p, pinv));
                a[ci][1] = __dadd_rn(a[ci][1], fmodmul(v1, w01, p, pinv));
                a[ci][2] = __dadd_rn(a[ci][2], fmodmul(v0, w10, p, pinv));
                a[ci][3] = __dadd_rn(a[ci][3], fmodmul(v1, w11, p, pinv));
            }
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kt_sptr(empty + s)) : "memory"); // this thread is done with the stage
        if ((dd & 7) == 7) { // sums of 8 fresh products stay below 4.1 p; re-centre before they could leave the exact range
#pragma unroll
            for (int ci = 0; ci < CT; ci++)
#pragma unroll
                for (int j = 0; j < 4; j++) a[ci][j] = frecenter(a[ci][j], p, pinv);
        }
    }
    const int x = x0 + 2 * tid;
#pragma unroll
    for (int ci = 0; ci < CT; ci++) {
        if (ci >= n_here) break;
#pragma unroll
        for (int j = 0; j < 4; j++) a[ci][j] = frecenter(a[ci][j], p, pinv);
        const size_t o = ((size_t)((c0 + ci) * 2) * k + l) * N + x;
        *reinterpret_cast<ulonglong2 *>(acc + o) = make_ulonglong2(lazy_bits(a[ci][0]), lazy_bits(a[ci][1]));
        *reinterpret_cast<ulonglong2 *>(acc + o + kpoly) = make_ulonglong2(lazy_bits(a[ci][2]), lazy_bits(a[ci][3]));
    }
}

static inline unsigned blocks_for(size_t threads) { return (unsigned)((threads + 255) / 256); }

// `f` is the HOST copy of the constants (passed by value into the kernel's parameter space)
cudaError_t launch_behz_lift_fp(const u64 *const *ct_ptrs, u64 *out, int n, int logn, const BehzConstF *f, int lazy, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (lazy) k_behz_lift_fp<true><<<blocks_for((size_t)n * 2 << logn), 256, 0, s>>>(ct_ptrs, out, n * 2, logn, *f);
    else k_behz_lift_fp<false><<<blocks_for((size_t)n * 2 << logn), 256, 0, s>>>(ct_ptrs, out, n * 2, logn, *f);
    return cudaGetLastError();
}
cudaError_t launch_behz_tensor_fp(const u64 *a, const u64 *b, u64 *d, int n, int kt, int logn, const BehzConstF *f, int lazy, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (lazy) k_behz_tensor_fp<true><<<blocks_for(((size_t)n * kt) << logn), 256, 0, s>>>(a, b, d, n, logn, *f);
    else k_behz_tensor_fp<false><<<blocks_for(((size_t)n * kt) << logn), 256, 0, s>>>(a, b, d, n, logn, *f);
    return cudaGetLastError();
}
cudaError_t launch_behz_floor_fp(const u64 *d, u64 *out3, int n, u64 t, int logn, const BehzConstF *f, int lazy, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (lazy) k_behz_floor_fp<true><<<blocks_for((size_t)n * 3 << logn), 256, 0, s>>>(d, out3, n * 3, (double)t, logn, *f);
    else k_behz_floor_fp<false><<<blocks_for((size_t)n * 3 << logn), 256, 0, s>>>(d, out3, n * 3, (double)t, logn, *f);
    return cudaGetLastError();
}
cudaError_t launch_behz_floor_fold_fp(const u64 *d, u64 *out3, int n, int logn, const FloorConstF *f, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const size_t total = (size_t)n * 3 << logn;
    k_behz_floor_fold_fp<1><<<blocks_for(total), 256, 0, s>>>(d, out3, total, logn, *f);
    return cudaGetLastError();
}
cudaError_t launch_ks_mac_fp(const u64 *digits, const u64 *key, u64 *acc, int n, int D, int k, int logn, const BehzConstF *f, int lazy,
                             cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    static const int unr = getenv("CNHE_KSMAC") ? atoi(getenv("CNHE_KSMAC")) : 1;       // tuning knob: digits in flight per thread
    static const int cts = getenv("CNHE_KSMAC_CT") ? atoi(getenv("CNHE_KSMAC_CT")) : 4; // ciphertexts per thread (key reuse)
    // key reuse pays once the launch fills the GPU anyway: with few ciphertexts (LoLa: 1-32 per call) four per thread leaves SMs idle
    const int ct = n < 64 ? 1 : (cts == 1 || cts == 2 || (cts == 8 && lazy) ? cts : 4);
    // full waves of lazy digits: the copy-engine-staged kernel (CNHE_KSMAC_TMA=0 keeps the register version)
    const bool tma = getenv("CNHE_KSMAC_TMA") ? atoi(getenv("CNHE_KSMAC_TMA")) != 0 : true; // read per launch (tests compare both)
    if (lazy && tma && ct == 4 && (1 << logn) % KT_X == 0) {
        constexpr int smem = KT_STAGES * (4 + 2) * KT_X * 8 + 2 * KT_STAGES * 8;
        cudaError_t e = cudaFuncSetAttribute(k_ks_mac_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)(((1 << logn) / KT_X) * k * ((n + 3) / 4));
        k_ks_mac_tma<4><<<grid, KT_THREADS, smem, s>>>(digits, key, acc, n, D, logn, *f);
        return cudaGetLastError();
    }
    const unsigned blocks = blocks_for(((size_t)((n + ct - 1) / ct) * k) << (logn - 1));
    if (!lazy) {
        if (ct == 4) k_ks_mac_fp<false, 1, 4><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
        else if (ct == 2) k_ks_mac_fp<false, 2, 2><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
        else k_ks_mac_fp<false, 4, 1><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    } else if (ct == 8) {
        k_ks_mac_fp<true, 1, 8><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    } else if (ct == 4) {
        if (unr == 2) k_ks_mac_fp<true, 2, 4><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
        else k_ks_mac_fp<true, 1, 4><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    } else if (ct == 2) {
        if (unr == 2) k_ks_mac_fp<true, 2, 2><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
        else k_ks_mac_fp<true, 1, 2><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    } else if (unr == 1) k_ks_mac_fp<true, 1, 1><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    else if (unr == 2) k_ks_mac_fp<true, 2, 1><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    else if (unr == 8) k_ks_mac_fp<true, 8, 1><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    else k_ks_mac_fp<true, 4, 1><<<blocks, 256, 0, s>>>(digits, key, acc, n, D, logn, *f);
    return cudaGetLastError();
}

} // namespace cnhe
