// K1/K2: batched negacyclic NTT / inverse NTT over one residue polynomial per CTA  (sm_100a).
//
// Replaces SEAL 3.2 util::ntt_negacyclic_harvey(_lazy) / inverse_ntt_negacyclic_harvey(_lazy), reached from every
// Evaluator.Multiply / Relinearize / Rotate* / dense MultiplyPlain call site of
// /root/reference "HE Wrapper/AtomicSealBfvVector.cs" (map in SURVEY.md section 8a).
//
// Design: the whole residue polynomial (8N bytes: 32..128 KiB) lives in shared memory for the duration of the
// transform, so HBM sees exactly one read and one write of it (16N algorithmic bytes).  Each thread owns 16
// coefficients in registers and runs 2..4 radix-2 stages per pass (3..4 passes for log2 N = 10..14); twiddles and
// their Shoup quotients come through the read-only path (L1/L2-resident: 16N bytes per modulus shared by the batch).
// Butterflies are Harvey lazy butterflies on the integer pipe (values in [0,4p) forward, [0,2p) inverse); the result
// written back is canonical, which is what makes the kernel bit-comparable with the CPU oracle.
// Shared-memory layout: word i is stored at i ^ (((i>>4)&7)<<1) so that the unit-stride last pass (16 consecutive
// words per thread, 16-byte accesses) and the strided passes (gap >= 16 words) are both bank-conflict free.
#include "kernels.h"
#include "fparith.cuh"

namespace cnhe {

__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 7) << 1); }

__device__ __forceinline__ void ct_butterfly(u64 &X, u64 &Y, u64 W, u64 Ws, u64 p, u64 two_p) {
    u64 a = X;
    a = a >= two_p ? a - two_p : a;
    u64 t = mul_shoup_lazy(Y, W, Ws, p);
    X = a + t;
    Y = a - t + two_p;
}
__device__ __forceinline__ void gs_butterfly(u64 &X, u64 &Y, u64 W, u64 Ws, u64 p, u64 two_p) {
    u64 u = X, v = Y;
    u64 s = u + v;
    X = s >= two_p ? s - two_p : s;
    Y = mul_shoup_lazy(u - v + two_p, W, Ws, p);
}

struct FwdSrc {
    const u64 *src; // polynomial base (plain) or digit source polynomial
    int shift;      // digit mode
    u64 mask;
    bool digit, need_reduce;
};
__device__ __forceinline__ u64 fwd_load(const FwdSrc &s, int idx, const DMod &m) {
    u64 v = s.src[idx];
    if (s.digit) {
        v = (v >> s.shift) & s.mask;
        if (s.need_reduce) v = reduce64(v, m);
    }
    return v;
}

// Forward pass covering stages [S0, S0+R), gap of its last stage g = N >> (S0+R) >= 16.
template <int LOGN, int S0, int R, bool FROM_G>
__device__ __forceinline__ void fwd_pass(u64 *sm, const FwdSrc &src, const NttTab &tb, int tid) {
    constexpr int T = (1 << LOGN) / 16, G = 16 >> R, E = 1 << R, LG = LOGN - S0 - R;
    const u64 p = tb.mod.p, two_p = 2 * p;
#pragma unroll
    for (int gg = 0; gg < G; gg++) {
        const int gid = tid + gg * T;
        const int c = gid & ((1 << LG) - 1), j = gid >> LG;
        const int base = (j << (LG + R)) + c;
        u64 x[E];
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = FROM_G ? fwd_load(src, base + (e << LG), tb.mod) : sm[swz(base + (e << LG))];
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int h = E >> (u + 1);
#pragma unroll
            for (int e = 0; e < E; e++) {
                if (e & h) continue;
                const int tw = (1 << (S0 + u)) + (j << u) + (e >> (R - u));
                ct_butterfly(x[e], x[e + h], __ldg(tb.w + tw), __ldg(tb.ws + tw), p, two_p);
            }
        }
#pragma unroll
        for (int e = 0; e < E; e++) sm[swz(base + (e << LG))] = x[e];
    }
}
// Last forward pass: stages [LOGN-4, LOGN), 16 consecutive words per thread; canonical output left in smem.
template <int LOGN>
__device__ __forceinline__ void fwd_last(u64 *sm, const NttTab &tb, int tid) {
    constexpr int S0 = LOGN - 4;
    const u64 p = tb.mod.p, two_p = 2 * p;
    ulonglong2 *smv = reinterpret_cast<ulonglong2 *>(sm);
    const int j = tid, xr = j & 7;
    u64 x[16];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        ulonglong2 v = smv[j * 8 + (ch ^ xr)];
        x[2 * ch] = v.x;
        x[2 * ch + 1] = v.y;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 8 >> u;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const int tw = (1 << (S0 + u)) + (j << u) + (e >> (4 - u));
            ct_butterfly(x[e], x[e + h], __ldg(tb.w + tw), __ldg(tb.ws + tw), p, two_p);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        u64 a = x[2 * ch], b = x[2 * ch + 1];
        a = a >= two_p ? a - two_p : a;
        a = a >= p ? a - p : a;
        b = b >= two_p ? b - two_p : b;
        b = b >= p ? b - p : b;
        smv[j * 8 + (ch ^ xr)] = make_ulonglong2(a, b);
    }
}
template <int LOGN>
__device__ __forceinline__ void smem_to_global(const u64 *sm, u64 *dst, int tid) {
    constexpr int T = (1 << LOGN) / 16;
    const ulonglong2 *smv = reinterpret_cast<const ulonglong2 *>(sm);
    ulonglong2 *dv = reinterpret_cast<ulonglong2 *>(dst);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int ch = tid + i * T;
        dv[ch] = smv[ch ^ ((ch >> 3) & 7)];
    }
}
template <int LOGN>
__device__ __forceinline__ void global_to_smem(u64 *sm, const u64 *src, int tid) {
    constexpr int T = (1 << LOGN) / 16;
    ulonglong2 *smv = reinterpret_cast<ulonglong2 *>(sm);
    const ulonglong2 *sv = reinterpret_cast<const ulonglong2 *>(src);
    ulonglong2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = sv[tid + i * T];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int ch = tid + i * T;
        smv[ch ^ ((ch >> 3) & 7)] = v[i];
    }
}

template <int LOGN>
__device__ __forceinline__ void fwd_body(u64 *sm, const FwdSrc &src, const NttTab &tb, int tid) {
    if constexpr (LOGN == 10) {
        fwd_pass<10, 0, 2, true>(sm, src, tb, tid); __syncthreads();
        fwd_pass<10, 2, 4, false>(sm, src, tb, tid); __syncthreads();
    } else if constexpr (LOGN == 11) {
        fwd_pass<11, 0, 3, true>(sm, src, tb, tid); __syncthreads();
        fwd_pass<11, 3, 4, false>(sm, src, tb, tid); __syncthreads();
    } else if constexpr (LOGN == 12) {
        fwd_pass<12, 0, 4, true>(sm, src, tb, tid); __syncthreads();
        fwd_pass<12, 4, 4, false>(sm, src, tb, tid); __syncthreads();
    } else if constexpr (LOGN == 13) {
        fwd_pass<13, 0, 3, true>(sm, src, tb, tid); __syncthreads();
        fwd_pass<13, 3, 3, false>(sm, src, tb, tid); __syncthreads();
        fwd_pass<13, 6, 3, false>(sm, src, tb, tid); __syncthreads();
    } else {
        fwd_pass<14, 0, 3, true>(sm, src, tb, tid); __syncthreads();
        fwd_pass<14, 3, 3, false>(sm, src, tb, tid); __syncthreads();
        fwd_pass<14, 6, 4, false>(sm, src, tb, tid); __syncthreads();
    }
    fwd_last<LOGN>(sm, tb, tid);
    __syncthreads();
}

constexpr int min_blocks(int logn) { return logn >= 14 ? 1 : (logn == 13 ? 2 : (logn == 12 ? 3 : 2)); }

template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 16, min_blocks(LOGN))
k_ntt_forward(const u64 *src, u64 *dst, const NttTab *__restrict__ tabs, int mod_base, int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    FwdSrc fs;
    fs.src = src + (size_t)b * N;
    fs.digit = false; fs.need_reduce = false; fs.shift = 0; fs.mask = 0;
    fwd_body<LOGN>(sm, fs, tb, tid);
    smem_to_global<LOGN>(sm, dst + (size_t)b * N, tid);
}

template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 16, min_blocks(LOGN))
k_ntt_forward_digits(const u64 *target, u64 *dst, const NttTab *__restrict__ tabs, int k, DigitMap dm) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int l = b % k, d = (b / k) % dm.D, c = b / (k * dm.D);
    const NttTab tb = tabs[l];
    FwdSrc fs;
    fs.src = target + ((size_t)c * k + dm.src[d]) * N;
    fs.digit = true;
    fs.shift = dm.shift[d];
    fs.mask = dm.mask;
    fs.need_reduce = dm.mask >= tb.mod.p;
    fwd_body<LOGN>(sm, fs, tb, tid);
    smem_to_global<LOGN>(sm, dst + (size_t)b * N, tid);
}

// ---------------------------------------------------------------- inverse
template <int LOGN>
__device__ __forceinline__ void inv_first(u64 *sm, const NttTab &tb, int tid) {
    constexpr int N = 1 << LOGN;
    const u64 p = tb.mod.p, two_p = 2 * p;
    ulonglong2 *smv = reinterpret_cast<ulonglong2 *>(sm);
    const int j = tid, xr = j & 7;
    u64 x[16];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        ulonglong2 v = smv[j * 8 + (ch ^ xr)];
        x[2 * ch] = v.x;
        x[2 * ch + 1] = v.y;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 1 << u;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const int tw = (N >> (u + 1)) + (j << (3 - u)) + (e >> (u + 1));
            gs_butterfly(x[e], x[e + h], __ldg(tb.iw + tw), __ldg(tb.iws + tw), p, two_p);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ch++) smv[j * 8 + (ch ^ xr)] = make_ulonglong2(x[2 * ch], x[2 * ch + 1]);
}
// Inverse pass covering stages [V0, V0+R) (gap of its first stage g = 1<<V0 >= 16).  LAST: scale by N^-1,
// canonicalise and write straight to global (optionally adding `base`).
template <int LOGN, int V0, int R, bool LAST>
__device__ __forceinline__ void inv_pass(u64 *sm, u64 *dst, const u64 *base_add, const NttTab &tb, int tid) {
    constexpr int N = 1 << LOGN, T = N / 16, G = 16 >> R, E = 1 << R;
    const u64 p = tb.mod.p, two_p = 2 * p;
#pragma unroll
    for (int gg = 0; gg < G; gg++) {
        const int gid = tid + gg * T;
        const int c = gid & ((1 << V0) - 1), j = gid >> V0;
        const int base = (j << (V0 + R)) + c;
        u64 x[E];
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = sm[swz(base + (e << V0))];
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int h = 1 << u;
#pragma unroll
            for (int e = 0; e < E; e++) {
                if (e & h) continue;
                const int tw = (N >> (V0 + u + 1)) + (j << (R - 1 - u)) + (e >> (u + 1));
                gs_butterfly(x[e], x[e + h], __ldg(tb.iw + tw), __ldg(tb.iws + tw), p, two_p);
            }
        }
        if constexpr (LAST) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                u64 v = mul_shoup_lazy(x[e], tb.inv_n, tb.inv_n_s, p);
                v = v >= p ? v - p : v;
                const int idx = base + (e << V0);
                if (base_add) v = addmod(v, base_add[idx], p);
                dst[idx] = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; e++) sm[swz(base + (e << V0))] = x[e];
        }
    }
}

template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 16, min_blocks(LOGN))
k_ntt_inverse(const u64 *src, const u64 *base_add, u64 *dst, const NttTab *__restrict__ tabs, int mod_base, int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    global_to_smem<LOGN>(sm, src + (size_t)b * N, tid);
    __syncthreads();
    inv_first<LOGN>(sm, tb, tid);
    __syncthreads();
    u64 *d = dst + (size_t)b * N;
    const u64 *ba = base_add ? base_add + (size_t)b * N : nullptr;
    if constexpr (LOGN == 10) {
        inv_pass<10, 4, 4, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<10, 8, 2, true>(sm, d, ba, tb, tid);
    } else if constexpr (LOGN == 11) {
        inv_pass<11, 4, 4, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<11, 8, 3, true>(sm, d, ba, tb, tid);
    } else if constexpr (LOGN == 12) {
        inv_pass<12, 4, 4, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<12, 8, 4, true>(sm, d, ba, tb, tid);
    } else if constexpr (LOGN == 13) {
        inv_pass<13, 4, 3, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<13, 7, 3, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<13, 10, 3, true>(sm, d, ba, tb, tid);
    } else {
        inv_pass<14, 4, 4, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<14, 8, 3, false>(sm, d, ba, tb, tid); __syncthreads();
        inv_pass<14, 11, 3, true>(sm, d, ba, tb, tid);
    }
}

// ================================================================ FP64 butterfly path (p < 2^50)
// On B200 a 64x64->128-bit integer product costs ~9 IMAD-pipe slots (IMAD.WIDE issues at 0.77 and mul.hi.u64 at 0.23
// warp-instr/clk/SM, measured: profiles/r01_pipe_issue_rates.txt) while DFMA/DADD issue at 1.94 and overlap with the integer
// ALU.  For moduli below 2^50 -- all of SEAL's default coefficient primes and the 48-bit auxiliary base -- the butterfly
// is therefore done in double precision with error-free transformations:
//     h = a*w, l = fma(a,w,-h) (exact product h+l),  q = rint(h/p),  r = fma(-q,p,h) + l  ==  a*w - q*p  exactly,
// 6 DP ops for the modular product + 2 for the butterfly, no integer corrections at all: values stay centred and small
// (|r| <= (0.5 + 1.5|a|/2^53) p) and the host schedules a re-centring pass only where the bound could reach 2^52.
// The transform computed is the same function as the integer path (canonical output), so results are bit-identical.
template <int LOGN, int S0, int R, bool FROM_G, int PASS>
__device__ __forceinline__ void fwd_pass_fp(double *sm, const FwdSrc &src, const NttTab &tb, int tid) {
    constexpr int T = (1 << LOGN) / 16, G = 16 >> R, E = 1 << R, LG = LOGN - S0 - R;
    const double p = tb.pd, pinv = tb.pinv;
    const bool rc = (tb.fwd_recenter >> PASS) & 1;
#pragma unroll
    for (int gg = 0; gg < G; gg++) {
        const int gid = tid + gg * T;
        const int c = gid & ((1 << LG) - 1), j = gid >> LG;
        const int base = (j << (LG + R)) + c;
        double x[E];
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = FROM_G ? u2d(fwd_load(src, base + (e << LG), tb.mod)) : sm[swz(base + (e << LG))];
        if (rc) {
#pragma unroll
            for (int e = 0; e < E; e++) x[e] = frecenter(x[e], p, pinv);
        }
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int h = E >> (u + 1);
#pragma unroll
            for (int e = 0; e < E; e++) {
                if (e & h) continue;
                const double w = __ldg(tb.wd + ((1 << (S0 + u)) + (j << u) + (e >> (R - u))));
                const double t = fmodmul(x[e + h], w, p, pinv);
                const double a = x[e];
                x[e] = __dadd_rn(a, t);
                x[e + h] = __dsub_rn(a, t);
            }
        }
#pragma unroll
        for (int e = 0; e < E; e++) sm[swz(base + (e << LG))] = x[e];
    }
}
template <int LOGN, int PASS>
__device__ __forceinline__ void fwd_last_fp(double *sm, const NttTab &tb, int tid) {
    constexpr int S0 = LOGN - 4;
    const double p = tb.pd, pinv = tb.pinv;
    const bool rc = (tb.fwd_recenter >> PASS) & 1;
    double2 *smv = reinterpret_cast<double2 *>(sm);
    const int j = tid, xr = j & 7;
    double x[16];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        double2 v = smv[j * 8 + (ch ^ xr)];
        x[2 * ch] = v.x;
        x[2 * ch + 1] = v.y;
    }
    if (rc) {
#pragma unroll
        for (int e = 0; e < 16; e++) x[e] = frecenter(x[e], p, pinv);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 8 >> u;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const double w = __ldg(tb.wd + ((1 << (S0 + u)) + (j << u) + (e >> (4 - u))));
            const double t = fmodmul(x[e + h], w, p, pinv);
            const double a = x[e];
            x[e] = __dadd_rn(a, t);
            x[e + h] = __dsub_rn(a, t);
        }
    }
    ulonglong2 *smu = reinterpret_cast<ulonglong2 *>(sm);
#pragma unroll
    for (int ch = 0; ch < 8; ch++)
        smu[j * 8 + (ch ^ xr)] = make_ulonglong2(d2u(fcanon(x[2 * ch], p, pinv)), d2u(fcanon(x[2 * ch + 1], p, pinv)));
}
template <int LOGN>
__device__ __forceinline__ void fwd_body_fp(double *sm, const FwdSrc &src, const NttTab &tb, int tid) {
    if constexpr (LOGN == 10) {
        fwd_pass_fp<10, 0, 2, true, 0>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<10, 2, 4, false, 1>(sm, src, tb, tid); __syncthreads();
        fwd_last_fp<10, 2>(sm, tb, tid);
    } else if constexpr (LOGN == 11) {
        fwd_pass_fp<11, 0, 3, true, 0>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<11, 3, 4, false, 1>(sm, src, tb, tid); __syncthreads();
        fwd_last_fp<11, 2>(sm, tb, tid);
    } else if constexpr (LOGN == 12) {
        fwd_pass_fp<12, 0, 4, true, 0>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<12, 4, 4, false, 1>(sm, src, tb, tid); __syncthreads();
        fwd_last_fp<12, 2>(sm, tb, tid);
    } else if constexpr (LOGN == 13) {
        fwd_pass_fp<13, 0, 3, true, 0>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<13, 3, 3, false, 1>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<13, 6, 3, false, 2>(sm, src, tb, tid); __syncthreads();
        fwd_last_fp<13, 3>(sm, tb, tid);
    } else {
        fwd_pass_fp<14, 0, 3, true, 0>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<14, 3, 3, false, 1>(sm, src, tb, tid); __syncthreads();
        fwd_pass_fp<14, 6, 4, false, 2>(sm, src, tb, tid); __syncthreads();
        fwd_last_fp<14, 3>(sm, tb, tid);
    }
    __syncthreads();
}
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 16, min_blocks(LOGN))
k_ntt_forward_fp(const u64 *src, u64 *dst, const NttTab *__restrict__ tabs, int mod_base, int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    FwdSrc fs;
    fs.src = src + (size_t)b * N;
    fs.digit = false; fs.need_reduce = false; fs.shift = 0; fs.mask = 0;
    fwd_body_fp<LOGN>(reinterpret_cast<double *>(sm), fs, tb, tid);
    smem_to_global<LOGN>(sm, dst + (size_t)b * N, tid);
}
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 16, min_blocks(LOGN))
k_ntt_forward_digits_fp(const u64 *target, u64 *dst, const NttTab *__restrict__ tabs, int k, DigitMap dm) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int l = b % k, d = (b / k) % dm.D, c = b / (k * dm.D);
    const NttTab tb = tabs[l];
    FwdSrc fs;
    fs.src = target + ((size_t)c * k + dm.src[d]) * N;
    fs.digit = true;
    fs.shift = dm.shift[d];
    fs.mask = dm.mask;
    fs.need_reduce = dm.mask >= tb.mod.p;
    fwd_body_fp<LOGN>(reinterpret_cast<double *>(sm), fs, tb, tid);
    smem_to_global<LOGN>(sm, dst + (size_t)b * N, tid);
}

// ---- inverse, FP64
template <int LOGN>
__device__ __forceinline__ void inv_first_fp(u64 *smraw, const NttTab &tb, int tid) {
    constexpr int N = 1 << LOGN;
    const double p = tb.pd, pinv = tb.pinv;
    ulonglong2 *smu = reinterpret_cast<ulonglong2 *>(smraw);
    double2 *smv = reinterpret_cast<double2 *>(smraw);
    const int j = tid, xr = j & 7;
    const u64 half = tb.mod.p >> 1;
    double x[16];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) { // centred load: |x| <= p/2
        ulonglong2 v = smu[j * 8 + (ch ^ xr)];
        x[2 * ch] = v.x > half ? __dsub_rn(u2d(v.x), p) : u2d(v.x);
        x[2 * ch + 1] = v.y > half ? __dsub_rn(u2d(v.y), p) : u2d(v.y);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 1 << u;
        const bool rc = (tb.inv_recenter >> u) & 1;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const double w = __ldg(tb.iwd + ((N >> (u + 1)) + (j << (3 - u)) + (e >> (u + 1))));
            const double a = x[e], bq = x[e + h];
            const double sum = __dadd_rn(a, bq);
            x[e] = rc ? frecenter(sum, p, pinv) : sum;
            x[e + h] = fmodmul(__dsub_rn(a, bq), w, p, pinv);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ch++) smv[j * 8 + (ch ^ xr)] = make_double2(x[2 * ch], x[2 * ch + 1]);
}
template <int LOGN, int V0, int R, bool LAST, int PASS>
__device__ __forceinline__ void inv_pass_fp(double *sm, u64 *dst, const u64 *base_add, const NttTab &tb, int tid) {
    constexpr int N = 1 << LOGN, T = N / 16, G = 16 >> R, E = 1 << R;
    const double p = tb.pd, pinv = tb.pinv;
#pragma unroll
    for (int gg = 0; gg < G; gg++) {
        const int gid = tid + gg * T;
        const int c = gid & ((1 << V0) - 1), j = gid >> V0;
        const int base = (j << (V0 + R)) + c;
        double x[E];
#pragma unroll
        for (int e = 0; e < E; e++) x[e] = sm[swz(base + (e << V0))];
#pragma unroll
        for (int u = 0; u < R; u++) {
            const int h = 1 << u;
            const bool rc = (tb.inv_recenter >> (V0 + u)) & 1; // re-centre the sums of stage V0+u (host-scheduled)
#pragma unroll
            for (int e = 0; e < E; e++) {
                if (e & h) continue;
                const double w = __ldg(tb.iwd + ((N >> (V0 + u + 1)) + (j << (R - 1 - u)) + (e >> (u + 1))));
                const double a = x[e], bq = x[e + h];
                const double sum = __dadd_rn(a, bq);
                x[e] = rc ? frecenter(sum, p, pinv) : sum;
                x[e + h] = fmodmul(__dsub_rn(a, bq), w, p, pinv);
            }
        }
        if constexpr (LAST) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                double r = fmodmul(x[e], tb.inv_n_d, p, pinv); // |x| < 2^52 (host-checked); r in (-1.3p, 1.3p)
                r = r < 0.0 ? __dadd_rn(r, p) : r;
                r = r < 0.0 ? __dadd_rn(r, p) : r;
                r = r >= p ? __dsub_rn(r, p) : r;
                u64 v = d2u(r);
                const int idx = base + (e << V0);
                if (base_add) v = addmod(v, base_add[idx], tb.mod.p);
                dst[idx] = v;
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; e++) sm[swz(base + (e << V0))] = x[e];
        }
    }
}
template <int LOGN>
__global__ void __launch_bounds__((1 << LOGN) / 16, min_blocks(LOGN))
k_ntt_inverse_fp(const u64 *src, const u64 *base_add, u64 *dst, const NttTab *__restrict__ tabs, int mod_base, int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    global_to_smem<LOGN>(sm, src + (size_t)b * N, tid);
    __syncthreads();
    inv_first_fp<LOGN>(sm, tb, tid);
    __syncthreads();
    double *smd = reinterpret_cast<double *>(sm);
    u64 *d = dst + (size_t)b * N;
    const u64 *ba = base_add ? base_add + (size_t)b * N : nullptr;
    if constexpr (LOGN == 10) {
        inv_pass_fp<10, 4, 4, false, 1>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<10, 8, 2, true, 2>(smd, d, ba, tb, tid);
    } else if constexpr (LOGN == 11) {
        inv_pass_fp<11, 4, 4, false, 1>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<11, 8, 3, true, 2>(smd, d, ba, tb, tid);
    } else if constexpr (LOGN == 12) {
        inv_pass_fp<12, 4, 4, false, 1>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<12, 8, 4, true, 2>(smd, d, ba, tb, tid);
    } else if constexpr (LOGN == 13) {
        inv_pass_fp<13, 4, 3, false, 1>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<13, 7, 3, false, 2>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<13, 10, 3, true, 3>(smd, d, ba, tb, tid);
    } else {
        inv_pass_fp<14, 4, 4, false, 1>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<14, 8, 3, false, 2>(smd, d, ba, tb, tid); __syncthreads();
        inv_pass_fp<14, 11, 3, true, 3>(smd, d, ba, tb, tid);
    }
}

int ntt_pass_radices(int logn, int inverse, int *r) {
    static const int F[5][4] = {{2, 4, 4, 0}, {3, 4, 4, 0}, {4, 4, 4, 0}, {3, 3, 3, 4}, {3, 3, 4, 4}};
    static const int I[5][4] = {{4, 4, 2, 0}, {4, 4, 3, 0}, {4, 4, 4, 0}, {4, 3, 3, 3}, {4, 4, 3, 3}};
    if (logn < 10 || logn > 14) return 0;
    int n = 0;
    for (int i = 0; i < 4; i++) {
        r[i] = inverse ? I[logn - 10][i] : F[logn - 10][i];
        if (r[i]) n++;
    }
    return n;
}

// ---------------------------------------------------------------- launchers
int ntt_kernel_smem_bytes(int logn) { return (1 << logn) * 8; }

template <class K>
static cudaError_t prep(K kern, int logn) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ntt_kernel_smem_bytes(logn));
}

#define CNHE_DISPATCH_LOGN(logn, ...)                                                                                 \
    switch (logn) {                                                                                                     \
    case 10: { constexpr int L = 10; __VA_ARGS__; } break;                                                                     \
    case 11: { constexpr int L = 11; __VA_ARGS__; } break;                                                                     \
    case 12: { constexpr int L = 12; __VA_ARGS__; } break;                                                                     \
    case 13: { constexpr int L = 13; __VA_ARGS__; } break;                                                                     \
    case 14: { constexpr int L = 14; __VA_ARGS__; } break;                                                                     \
    default: return cudaErrorInvalidValue;                                                                              \
    }

cudaError_t launch_ntt_forward(const u64 *src, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base, int mod_count, int fp,
                               cudaStream_t s) {
    if (n_polys <= 0) return cudaSuccess;
    if (fp) {
        CNHE_DISPATCH_LOGN(logn, {
            cudaError_t e = prep(k_ntt_forward_fp<L>, L);
            if (e != cudaSuccess) return e;
            k_ntt_forward_fp<L><<<n_polys, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(src, dst, tabs, mod_base, mod_count);
        });
        return cudaGetLastError();
    }
    CNHE_DISPATCH_LOGN(logn, {
        cudaError_t e = prep(k_ntt_forward<L>, L);
        if (e != cudaSuccess) return e;
        k_ntt_forward<L><<<n_polys, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(src, dst, tabs, mod_base, mod_count);
    });
    return cudaGetLastError();
}
cudaError_t launch_ntt_forward_digits(const u64 *target, u64 *dst, int n_ct, int k, const DigitMap &dm, int logn, const NttTab *tabs, int fp,
                                      cudaStream_t s) {
    if (n_ct <= 0) return cudaSuccess;
    if (fp) {
        CNHE_DISPATCH_LOGN(logn, {
            cudaError_t e = prep(k_ntt_forward_digits_fp<L>, L);
            if (e != cudaSuccess) return e;
            k_ntt_forward_digits_fp<L><<<n_ct * dm.D * k, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(target, dst, tabs, k, dm);
        });
        return cudaGetLastError();
    }
    CNHE_DISPATCH_LOGN(logn, {
        cudaError_t e = prep(k_ntt_forward_digits<L>, L);
        if (e != cudaSuccess) return e;
        k_ntt_forward_digits<L><<<n_ct * dm.D * k, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(target, dst, tabs, k, dm);
    });
    return cudaGetLastError();
}
static cudaError_t launch_inv(const u64 *src, const u64 *base, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base,
                              int mod_count, int fp, cudaStream_t s) {
    if (n_polys <= 0) return cudaSuccess;
    if (fp) {
        CNHE_DISPATCH_LOGN(logn, {
            cudaError_t e = prep(k_ntt_inverse_fp<L>, L);
            if (e != cudaSuccess) return e;
            k_ntt_inverse_fp<L><<<n_polys, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(src, base, dst, tabs, mod_base, mod_count);
        });
        return cudaGetLastError();
    }
    CNHE_DISPATCH_LOGN(logn, {
        cudaError_t e = prep(k_ntt_inverse<L>, L);
        if (e != cudaSuccess) return e;
        k_ntt_inverse<L><<<n_polys, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(src, base, dst, tabs, mod_base, mod_count);
    });
    return cudaGetLastError();
}
cudaError_t launch_ntt_inverse(const u64 *src, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base, int mod_count, int fp,
                               cudaStream_t s) {
    return launch_inv(src, nullptr, dst, n_polys, logn, tabs, mod_base, mod_count, fp, s);
}
cudaError_t launch_ntt_inverse_add(const u64 *src, const u64 *base, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base,
                                   int mod_count, int fp, cudaStream_t s) {
    return launch_inv(src, base, dst, n_polys, logn, tabs, mod_base, mod_count, fp, s);
}

} // namespace cnhe
