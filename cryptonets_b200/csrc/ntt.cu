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
#include <cstdlib>

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
k_ntt_forward_digits(const u64 *target, size_t ct_stride, u64 *dst, const NttTab *__restrict__ tabs, int k, DigitMap dm) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int l = b % k, d = (b / k) % dm.D, c = b / (k * dm.D);
    const NttTab tb = tabs[l];
    FwdSrc fs;
    fs.src = target + (size_t)c * ct_stride + (size_t)dm.src[d] * N;
    fs.digit = true;
    fs.shift = dm.shift[d];
    fs.mask = dm.mask;
    fs.need_reduce = dm.mask >= tb.mod.p;
    fwd_body<LOGN>(sm, fs, tb, tid);
    smem_to_global<LOGN>(sm, dst + (((size_t)c * k + l) * dm.D + d) * N, tid); // [c][l][d]: the layout the key MAC streams
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
k_ntt_inverse(const u64 *src, const u64 *base_add, int base_group, size_t base_stride, u64 *dst, const NttTab *__restrict__ tabs, int mod_base,
              int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    global_to_smem<LOGN>(sm, src + (size_t)b * N, tid);
    __syncthreads();
    inv_first<LOGN>(sm, tb, tid);
    __syncthreads();
    u64 *d = dst + (size_t)b * N;
    const u64 *ba = base_add ? base_add + (size_t)(b / base_group) * base_stride + (size_t)(b % base_group) * N : nullptr;
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
// twiddles with table index < TWC are served from a per-CTA shared-memory copy (loaded once, under the first data loads)
constexpr int TWC = 512;
__device__ __forceinline__ void load_twiddle_cache(double *twc, const double *tw, int tid, int nthreads) {
    for (int i = tid; i < TWC; i += nthreads) twc[i] = __ldg(tw + i);
}
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a thread moves its 16 consecutive words as four full 32-byte sectors
__device__ __forceinline__ void ldg256(const u64 *p, u64 &a, u64 &b, u64 &c, u64 &d) {
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// Shared-memory round trips are what keeps the FP64 pipe idle (tools/dp_pass_bench.cu: 96 % utilisation on registers, ~62 %
// with an LDS/STS round trip every 3 stages), so the FP64 transform uses as few, as fat passes as the register file allows:
// N=8192 is 5+4+4 stages (was 3+3+3+4), the first pass reads HBM directly, the last one writes HBM directly.
// A pass over stages [S0, S0+R) is executed by "virtual threads": N/32 of them for R=5 (32 coefficients each), N/16 otherwise
// (16 coefficients: one radix-16 group, or two adjacent columns of radix-8 / four of radix-4 with 16-byte accesses).
// first-pass load: canonical u64 word, a digit of it, or (IN_F) a lazy double written by the producing kernel
template <bool IN_F>
__device__ __forceinline__ double fwd_load_fp(const FwdSrc &s, int idx, double p, double pinv) {
    if constexpr (IN_F) return ld_lazy(s.src + idx);
    u64 v = s.src[idx];
    if (s.digit) {
        v = (v >> s.shift) & s.mask; // source residue < 2^50, so every digit converts exactly
        const double x = u2d(v);
        return s.need_reduce ? frecenter(x, p, pinv) : x;
    }
    return u2d(v);
}
template <int LOGN, int S0, int R, bool FROM_G, int PASS, bool IN_F, int NV = 1>
__device__ __forceinline__ void fwd_pass_fp(double *sm, const double *twc, const FwdSrc &src, const NttTab &tb, int vt, int vstride = 0) {
    constexpr int E = 1 << R, LG = LOGN - S0 - R;
    constexpr bool CACHED = (S0 + R) <= 9; // every twiddle index of this pass is below TWC
    const double p = tb.pd, pinv = tb.pinv;
    const bool rc = (tb.fwd_recenter >> PASS) & 1;
    if constexpr (R <= 3) {
        constexpr int T = (1 << LOGN) / 16, G = 16 >> R;
#pragma unroll
        for (int gg = 0; gg < G / 2; gg++) {
            const int gid = vt + gg * T;
            const int c2 = gid & ((1 << (LG - 1)) - 1), j = gid >> (LG - 1);
            const int base = (j << (LG + R)) + 2 * c2;
            double x[E], y[E];
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int idx = base + (e << LG);
                if constexpr (FROM_G) {
                    x[e] = fwd_load_fp<IN_F>(src, idx, p, pinv);
                    y[e] = fwd_load_fp<IN_F>(src, idx + 1, p, pinv);
                } else {
                    const double2 v = *reinterpret_cast<const double2 *>(sm + swz(idx));
                    x[e] = v.x;
                    y[e] = v.y;
                }
            }
            if (rc) {
#pragma unroll
                for (int e = 0; e < E; e++) { x[e] = frecenter(x[e], p, pinv); y[e] = frecenter(y[e], p, pinv); }
            }
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int h = E >> (u + 1);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if (e & h) continue;
                    const int ti = (1 << (S0 + u)) + (j << u) + (e >> (R - u));
                    const double w = CACHED ? twc[ti] : __ldg(tb.wd + ti);
                    const double t0 = fmodmul(x[e + h], w, p, pinv), t1 = fmodmul(y[e + h], w, p, pinv);
                    const double a0 = x[e], a1 = y[e];
                    x[e] = __dadd_rn(a0, t0);
                    x[e + h] = __dsub_rn(a0, t0);
                    y[e] = __dadd_rn(a1, t1);
                    y[e + h] = __dsub_rn(a1, t1);
                }
            }
#pragma unroll
            for (int e = 0; e < E; e++) *reinterpret_cast<double2 *>(sm + swz(base + (e << LG))) = make_double2(x[e], y[e]);
        }
    } else {
        // NV independent groups per call (virtual threads vt, vt + vstride, ...): all their shared-memory loads are issued before
        // the first butterfly, so one group's LDS latency hides under the other's arithmetic (the compiler cannot hoist them itself
        // across the stores of the previous group)
        double x[NV][E];
        int jj[NV], bb[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const int v = vt + i * vstride;
            const int c = v & ((1 << LG) - 1);
            jj[i] = v >> LG;
            bb[i] = (jj[i] << (LG + R)) + c;
#pragma unroll
            for (int e = 0; e < E; e++) {
                if constexpr (FROM_G) x[i][e] = fwd_load_fp<IN_F>(src, bb[i] + (e << LG), p, pinv);
                else x[i][e] = sm[swz(bb[i] + (e << LG))];
            }
        }
        if (rc) {
#pragma unroll
            for (int i = 0; i < NV; i++)
#pragma unroll
                for (int e = 0; e < E; e++) x[i][e] = frecenter(x[i][e], p, pinv);
        }
#pragma unroll
        for (int i = 0; i < NV; i++) {
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int h = E >> (u + 1);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if (e & h) continue;
                    const int ti = (1 << (S0 + u)) + (jj[i] << u) + (e >> (R - u));
                    const double w = CACHED ? twc[ti] : __ldg(tb.wd + ti);
                    const double t = fmodmul(x[i][e + h], w, p, pinv);
                    const double a = x[i][e];
                    x[i][e] = __dadd_rn(a, t);
                    x[i][e + h] = __dsub_rn(a, t);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NV; i++)
#pragma unroll
            for (int e = 0; e < E; e++) sm[swz(bb[i] + (e << LG))] = x[i][e];
    }
}
// First forward pass split in two so that its HBM loads are in flight while the CTA fills its twiddle cache and waits at the
// barrier (ncu source view: 11 % of a transform's warp time sat in that fill with no data load outstanding).
// fwd_first_load: E = 2^R raw words of virtual thread vt (stride N/E); fwd_first_compute: convert, R stages, store to smem.
template <int LOGN, int R, bool IN_F>
__device__ __forceinline__ void fwd_first_load(u64 (&raw)[1 << R], const FwdSrc &src, int vt) {
    constexpr int E = 1 << R, LG = LOGN - R;
#pragma unroll
    for (int e = 0; e < E; e++) raw[e] = src.src[vt + (e << LG)];
}
template <int LOGN, int R, bool IN_F>
__device__ __forceinline__ void fwd_first_compute(double *sm, const double *twc, const u64 (&raw)[1 << R], const FwdSrc &src, const NttTab &tb, int vt) {
    constexpr int E = 1 << R, LG = LOGN - R;
    const double p = tb.pd, pinv = tb.pinv;
    double x[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        if constexpr (IN_F) x[e] = __longlong_as_double((long long)raw[e]);
        else {
            u64 v = raw[e];
            if (src.digit) v = (v >> src.shift) & src.mask; // source residue < 2^50, so every digit converts exactly
            x[e] = u2d(v);
            if (src.digit && src.need_reduce) x[e] = frecenter(x[e], p, pinv);
        }
    }
#pragma unroll
    for (int u = 0; u < R; u++) { // j = 0: the whole CTA uses twiddles [2^u, 2^(u+1)) in stage u
        const int h = E >> (u + 1);
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (e & h) continue;
            const double w = twc[(1 << u) + (e >> (R - u))];
            const double t = fmodmul(x[e + h], w, p, pinv);
            const double a = x[e];
            x[e] = __dadd_rn(a, t);
            x[e + h] = __dsub_rn(a, t);
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++) sm[swz(vt + (e << LG))] = x[e];
}

// Last forward pass: stages [LOGN-4, LOGN) on 16 consecutive words; canonical result goes straight to HBM.
template <int LOGN, int PASS, bool OUT_F>
__device__ __forceinline__ void fwd_last_fp(const double *sm, u64 *dst, const NttTab &tb, int j) {
    constexpr int S0 = LOGN - 4;
    const double p = tb.pd, pinv = tb.pinv;
    const bool rc = (tb.fwd_recenter >> PASS) & 1;
    const double2 *smv = reinterpret_cast<const double2 *>(sm);
    const int xr = j & 7;
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
    // twiddles of this pass: 1 + 2 + 4 + 8 consecutive doubles per thread, fetched as 16-byte loads up front
    double tw[15];
    tw[0] = __ldg(tb.wd + ((1 << S0) + j));
    {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(tb.wd + ((1 << (S0 + 1)) + (j << 1))));
        tw[1] = a.x; tw[2] = a.y;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double2 b = __ldg(reinterpret_cast<const double2 *>(tb.wd + ((1 << (S0 + 2)) + (j << 2) + 2 * i)));
            tw[3 + 2 * i] = b.x; tw[4 + 2 * i] = b.y;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const double2 c = __ldg(reinterpret_cast<const double2 *>(tb.wd + ((1 << (S0 + 3)) + (j << 3) + 2 * i)));
            tw[7 + 2 * i] = c.x; tw[8 + 2 * i] = c.y;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 8 >> u;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const double w = tw[(1 << u) - 1 + (e >> (4 - u))];
            const double t = fmodmul(x[e + h], w, p, pinv);
            const double a = x[e];
            x[e] = __dadd_rn(a, t);
            x[e + h] = __dsub_rn(a, t);
        }
    }
    u64 *o = dst + 16 * j;
    if constexpr (OUT_F) { // lazy doubles: |x| <= fwd bound * p; re-centred only where the consumer's product could overflow (host flag)
        if (tb.fwd_out_rc) {
#pragma unroll
            for (int e = 0; e < 16; e++) x[e] = frecenter(x[e], p, pinv);
        }
#pragma unroll
        for (int g = 0; g < 4; g++) stg256(o + 4 * g, lazy_bits(x[4 * g]), lazy_bits(x[4 * g + 1]), lazy_bits(x[4 * g + 2]), lazy_bits(x[4 * g + 3]));
    } else {
#pragma unroll
        for (int g = 0; g < 4; g++)
            stg256(o + 4 * g, fcanon_u(x[4 * g], p, pinv), fcanon_u(x[4 * g + 1], p, pinv), fcanon_u(x[4 * g + 2], p, pinv), fcanon_u(x[4 * g + 3], p, pinv));
    }
}

#define CNHE_VTN(COUNT, stmt) _Pragma("unroll") for (int vt = tid; vt < (COUNT); vt += TR) { stmt; }
__host__ __device__ constexpr int fp_threads(int logn) { return logn >= 13 ? (1 << logn) / 32 : (1 << logn) / 16; }
__host__ __device__ constexpr int fp_min_blocks(int logn) { return logn >= 14 ? 1 : (logn == 13 ? 2 : 3); }

// Ask L2 for the polynomial that the CTA taking this one's place will read (CTAs are dispatched in blockIdx order, so that
// is about `resident` blocks ahead): its first-pass loads then hit L2 instead of waiting on HBM with the FP64 pipe idle.
template <int LOGN>
__device__ __forceinline__ void prefetch_next_poly(const u64 *src_base, int b, int n_polys, int tid) {
    constexpr int N = 1 << LOGN, TR = fp_threads(LOGN);
    const int ahead = b + 148 * fp_min_blocks(LOGN);
    if (ahead < n_polys) {
        const char *p = reinterpret_cast<const char *>(src_base + (size_t)ahead * N);
        for (int i = tid * 128; i < N * 8; i += TR * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i));
    }
}
template <int LOGN, bool IN_F, bool OUT_F>
__device__ __forceinline__ void fwd_body_fp(double *sm, const FwdSrc &src, u64 *dst, const NttTab &tb, int tid) {
    constexpr int N = 1 << LOGN, TR = fp_threads(LOGN);
    double *twc = sm + N;
    if constexpr (LOGN == 12 || LOGN == 14) { // loads first, then the twiddle cache fill (+2..3 % at N=4096/16384; -3 % at N=8192, which keeps the plain order)
        constexpr int R1 = LOGN == 12 ? 4 : 5;
        static_assert((N >> R1) == TR, "first pass: one virtual thread per thread");
        u64 raw[1 << R1];
        fwd_first_load<LOGN, R1, IN_F>(raw, src, tid);
        load_twiddle_cache(twc, tb.wd, tid, TR);
        __syncthreads();
        fwd_first_compute<LOGN, R1, IN_F>(sm, twc, raw, src, tb, tid);
        __syncthreads();
        if constexpr (LOGN == 12) {
            CNHE_VTN(N / 16, (fwd_pass_fp<12, 4, 4, false, 1, false>(sm, twc, src, tb, vt))); __syncthreads();
            CNHE_VTN(N / 16, (fwd_last_fp<12, 2, OUT_F>(sm, dst, tb, vt)));
        } else {
            CNHE_VTN(N / 32, (fwd_pass_fp<14, 5, 5, false, 1, false>(sm, twc, src, tb, vt))); __syncthreads();
            CNHE_VTN(N / 16, (fwd_last_fp<14, 2, OUT_F>(sm, dst, tb, vt)));
        }
        return;
    }
    load_twiddle_cache(twc, tb.wd, tid, TR);
    __syncthreads();
    if constexpr (LOGN == 10) {
        CNHE_VTN(N / 16, (fwd_pass_fp<10, 0, 2, true, 0, IN_F>(sm, twc, src, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (fwd_pass_fp<10, 2, 4, false, 1, false>(sm, twc, src, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (fwd_last_fp<10, 2, OUT_F>(sm, dst, tb, vt)));
    } else if constexpr (LOGN == 11) {
        CNHE_VTN(N / 16, (fwd_pass_fp<11, 0, 3, true, 0, IN_F>(sm, twc, src, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (fwd_pass_fp<11, 3, 4, false, 1, false>(sm, twc, src, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (fwd_last_fp<11, 2, OUT_F>(sm, dst, tb, vt)));
    } else if constexpr (LOGN == 13) {
        CNHE_VTN(N / 32, (fwd_pass_fp<13, 0, 5, true, 0, IN_F>(sm, twc, src, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (fwd_pass_fp<13, 5, 4, false, 1, false>(sm, twc, src, tb, vt))); __syncthreads(); // NV=2 (both groups' loads first) measured 5% slower
        CNHE_VTN(N / 16, (fwd_last_fp<13, 2, OUT_F>(sm, dst, tb, vt)));
    }
}
template <int LOGN, bool IN_F, bool OUT_F>
__global__ void __launch_bounds__(fp_threads(LOGN), fp_min_blocks(LOGN))
k_ntt_forward_fp(const u64 *src, u64 *dst, const NttTab *__restrict__ tabs, int mod_base, int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    FwdSrc fs;
    fs.src = src + (size_t)b * N;
    fs.digit = false; fs.need_reduce = false; fs.shift = 0; fs.mask = 0;
    prefetch_next_poly<LOGN>(src, b, gridDim.x, tid);
    fwd_body_fp<LOGN, IN_F, OUT_F>(reinterpret_cast<double *>(sm), fs, dst + (size_t)b * N, tb, tid);
}
// target: ciphertext c's polynomial with `k` residues starts at target + c * ct_stride (words)
template <int LOGN, bool OUT_F>
__global__ void __launch_bounds__(fp_threads(LOGN), fp_min_blocks(LOGN))
k_ntt_forward_digits_fp(const u64 *target, size_t ct_stride, u64 *dst, const NttTab *__restrict__ tabs, int k, DigitMap dm) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int N = 1 << LOGN;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int l = b % k, d = (b / k) % dm.D, c = b / (k * dm.D);
    const NttTab tb = tabs[l];
    FwdSrc fs;
    fs.src = target + (size_t)c * ct_stride + (size_t)dm.src[d] * N;
    fs.digit = true;
    fs.shift = dm.shift[d];
    fs.mask = dm.mask;
    fs.need_reduce = dm.mask >= tb.mod.p;
    fwd_body_fp<LOGN, false, OUT_F>(reinterpret_cast<double *>(sm), fs, dst + (((size_t)c * k + l) * dm.D + d) * N, tb, tid); // [c][l][d]
}

// ---- inverse, FP64: first pass reads 16 consecutive words per virtual thread straight from HBM (256-bit loads)
template <int LOGN, bool IN_F>
__device__ __forceinline__ void inv_first_fp(double *sm, const u64 *src, const NttTab &tb, int j) {
    constexpr int N = 1 << LOGN;
    const double p = tb.pd, pinv = tb.pinv;
    double2 *smv = reinterpret_cast<double2 *>(sm);
    const int xr = j & 7;
    double x[16];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        u64 v0, v1, v2, v3;
        ldg256(src + 16 * j + 4 * g, v0, v1, v2, v3);
        if constexpr (IN_F) {
            x[4 * g] = __longlong_as_double((long long)v0);
            x[4 * g + 1] = __longlong_as_double((long long)v1);
            x[4 * g + 2] = __longlong_as_double((long long)v2);
            x[4 * g + 3] = __longlong_as_double((long long)v3);
        } else {
            x[4 * g] = u2d(v0);
            x[4 * g + 1] = u2d(v1);
            x[4 * g + 2] = u2d(v2);
            x[4 * g + 3] = u2d(v3);
        }
    }
    // stage u uses 8 >> u consecutive inverse twiddles: 8 + 4 + 2 + 1 doubles per thread, 16-byte loads
    double tw[15];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(tb.iwd + ((N >> 1) + (j << 3) + 2 * i)));
        tw[2 * i] = a.x; tw[2 * i + 1] = a.y;
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(tb.iwd + ((N >> 2) + (j << 2) + 2 * i)));
        tw[8 + 2 * i] = a.x; tw[9 + 2 * i] = a.y;
    }
    {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(tb.iwd + ((N >> 3) + (j << 1))));
        tw[12] = a.x; tw[13] = a.y;
        tw[14] = __ldg(tb.iwd + ((N >> 4) + j));
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 1 << u;
        const bool rc = (tb.inv_recenter >> u) & 1;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const double w = tw[(16 - (16 >> u)) + (e >> (u + 1))];
            const double a = x[e], bq = x[e + h];
            x[e] = __dadd_rn(a, bq);
            x[e + h] = fmodmul(__dsub_rn(a, bq), w, p, pinv);
        }
        if (rc) { // uniform branch: the host schedules a re-centring of the sums on very few stages
#pragma unroll
            for (int e = 0; e < 16; e++)
                if (!(e & h)) x[e] = frecenter(x[e], p, pinv);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ch++) smv[j * 8 + (ch ^ xr)] = make_double2(x[2 * ch], x[2 * ch + 1]);
}
// The last stage (one twiddle, iw[1]) carries N^-1: sums are multiplied by N^-1, differences by iw[1]*N^-1, so every output is a
// fresh modular product in (-0.51p, 0.51p): written as is (OUT_F, lazy double) or sign-fixed on the integer pipe (canonical).
template <int LOGN, int V0, int R, bool LAST, bool OUT_F, int NV = 1>
__device__ __forceinline__ void inv_pass_fp(double *sm, const double *twc, u64 *dst, const u64 *base_add, const NttTab &tb, int vt, int vstride = 0) {
    constexpr int N = 1 << LOGN, E = 1 << R;
    constexpr bool CACHED = (N >> V0) <= TWC; // stage v reads indices [N>>(v+1), N>>v)
    const double p = tb.pd, pinv = tb.pinv;
    auto finish = [&](double v, int idx) {
        if constexpr (OUT_F) return lazy_bits(v);
        u64 o = fsmall_u(v, tb.mod.p);
        if (base_add) o = addmod(o, base_add[idx], tb.mod.p);
        return o;
    };
    if constexpr (R <= 3) {
        constexpr int T = N / 16, G = 16 >> R;
#pragma unroll
        for (int gg = 0; gg < G / 2; gg++) {
            const int gid = vt + gg * T;
            const int c2 = gid & ((1 << (V0 - 1)) - 1), j = gid >> (V0 - 1);
            const int base = (j << (V0 + R)) + 2 * c2;
            double x[E], y[E];
#pragma unroll
            for (int e = 0; e < E; e++) {
                const double2 v = *reinterpret_cast<const double2 *>(sm + swz(base + (e << V0)));
                x[e] = v.x;
                y[e] = v.y;
            }
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int h = 1 << u;
                const bool rc = (tb.inv_recenter >> (V0 + u)) & 1;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if (e & h) continue;
                    const int ti = (N >> (V0 + u + 1)) + (j << (R - 1 - u)) + (e >> (u + 1));
                    const double w = CACHED ? twc[ti] : __ldg(tb.iwd + ti);
                    const double a0 = x[e], b0 = x[e + h], a1 = y[e], b1 = y[e + h];
                    if (LAST && u == R - 1) {
                        x[e] = fmodmul(__dadd_rn(a0, b0), tb.inv_n_d, p, pinv);
                        y[e] = fmodmul(__dadd_rn(a1, b1), tb.inv_n_d, p, pinv);
                        x[e + h] = fmodmul(__dsub_rn(a0, b0), tb.inv_n_w_d, p, pinv);
                        y[e + h] = fmodmul(__dsub_rn(a1, b1), tb.inv_n_w_d, p, pinv);
                    } else {
                        x[e] = __dadd_rn(a0, b0);
                        y[e] = __dadd_rn(a1, b1);
                        x[e + h] = fmodmul(__dsub_rn(a0, b0), w, p, pinv);
                        y[e + h] = fmodmul(__dsub_rn(a1, b1), w, p, pinv);
                    }
                }
                if (rc && !(LAST && u == R - 1)) {
#pragma unroll
                    for (int e = 0; e < E; e++)
                        if (!(e & h)) { x[e] = frecenter(x[e], p, pinv); y[e] = frecenter(y[e], p, pinv); }
                }
            }
            if constexpr (LAST) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const int idx = base + (e << V0);
                    *reinterpret_cast<ulonglong2 *>(dst + idx) = make_ulonglong2(finish(x[e], idx), finish(y[e], idx + 1));
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) *reinterpret_cast<double2 *>(sm + swz(base + (e << V0))) = make_double2(x[e], y[e]);
            }
        }
    } else {
        double x[NV][E];
        int jj[NV], bb[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) { // NV independent groups: loads first (see fwd_pass_fp)
            const int v = vt + i * vstride;
            const int c = v & ((1 << V0) - 1);
            jj[i] = v >> V0;
            bb[i] = (jj[i] << (V0 + R)) + c;
#pragma unroll
            for (int e = 0; e < E; e++) x[i][e] = sm[swz(bb[i] + (e << V0))];
        }
#pragma unroll
        for (int i = 0; i < NV; i++) {
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int h = 1 << u;
                const bool rc = (tb.inv_recenter >> (V0 + u)) & 1;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if (e & h) continue;
                    const int ti = (N >> (V0 + u + 1)) + (jj[i] << (R - 1 - u)) + (e >> (u + 1));
                    const double w = CACHED ? twc[ti] : __ldg(tb.iwd + ti);
                    const double a = x[i][e], bq = x[i][e + h];
                    if (LAST && u == R - 1) {
                        x[i][e] = fmodmul(__dadd_rn(a, bq), tb.inv_n_d, p, pinv);
                        x[i][e + h] = fmodmul(__dsub_rn(a, bq), tb.inv_n_w_d, p, pinv);
                    } else {
                        x[i][e] = __dadd_rn(a, bq);
                        x[i][e + h] = fmodmul(__dsub_rn(a, bq), w, p, pinv);
                    }
                }
                if (rc && !(LAST && u == R - 1)) {
#pragma unroll
                    for (int e = 0; e < E; e++)
                        if (!(e & h)) x[i][e] = frecenter(x[i][e], p, pinv);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (LAST) {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const int idx = bb[i] + (e << V0);
                    dst[idx] = finish(x[i][e], idx);
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) sm[swz(bb[i] + (e << V0))] = x[i][e];
            }
        }
    }
}
// base_add (optional): polynomial b is added to base_add[(b / base_group) * base_stride + (b % base_group) * N] (canonical output only)
template <int LOGN, bool IN_F, bool OUT_F>
__global__ void __launch_bounds__(fp_threads(LOGN), fp_min_blocks(LOGN))
k_ntt_inverse_fp(const u64 *src, const u64 *base_add, int base_group, size_t base_stride, u64 *dst, const NttTab *__restrict__ tabs, int mod_base,
                 int mod_count) {
    extern __shared__ __align__(16) u64 smraw[];
    constexpr int N = 1 << LOGN, TR = fp_threads(LOGN);
    const int b = blockIdx.x, tid = threadIdx.x;
    const NttTab tb = tabs[mod_base + b % mod_count];
    double *sm = reinterpret_cast<double *>(smraw);
    double *twc = sm + N;
    load_twiddle_cache(twc, tb.iwd, tid, TR);
    prefetch_next_poly<LOGN>(src, b, gridDim.x, tid);
    const u64 *s = src + (size_t)b * N;
    u64 *d = dst + (size_t)b * N;
    const u64 *ba = base_add ? base_add + (size_t)(b / base_group) * base_stride + (size_t)(b % base_group) * N : nullptr;
    if (ba) { // the base polynomial is consumed by the epilogue, ~20k cycles from now: have it waiting in L2
        const char *pb = reinterpret_cast<const char *>(ba);
        for (int i = tid * 128; i < N * 8; i += TR * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + i));
    }
    CNHE_VTN(N / 16, (inv_first_fp<LOGN, IN_F>(sm, s, tb, vt)));
    __syncthreads();
    if constexpr (LOGN == 10) {
        CNHE_VTN(N / 16, (inv_pass_fp<10, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (inv_pass_fp<10, 8, 2, true, OUT_F>(sm, twc, d, ba, tb, vt)));
    } else if constexpr (LOGN == 11) {
        CNHE_VTN(N / 16, (inv_pass_fp<11, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (inv_pass_fp<11, 8, 3, true, OUT_F>(sm, twc, d, ba, tb, vt)));
    } else if constexpr (LOGN == 12) {
        CNHE_VTN(N / 16, (inv_pass_fp<12, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); __syncthreads();
        CNHE_VTN(N / 16, (inv_pass_fp<12, 8, 4, true, OUT_F>(sm, twc, d, ba, tb, vt)));
    } else if constexpr (LOGN == 13) {
        CNHE_VTN(N / 16, (inv_pass_fp<13, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); __syncthreads();
        CNHE_VTN(N / 32, (inv_pass_fp<13, 8, 5, true, OUT_F>(sm, twc, d, ba, tb, vt)));
    } else {
        CNHE_VTN(N / 32, (inv_pass_fp<14, 4, 5, false, false>(sm, twc, d, ba, tb, vt))); __syncthreads();
        CNHE_VTN(N / 32, (inv_pass_fp<14, 9, 5, true, OUT_F>(sm, twc, d, ba, tb, vt)));
    }
}

int ntt_pass_radices(int logn, int inverse, int *r) {
    static const int F[5][4] = {{2, 4, 4, 0}, {3, 4, 4, 0}, {4, 4, 4, 0}, {5, 4, 4, 0}, {5, 5, 4, 0}};
    static const int I[5][4] = {{4, 4, 2, 0}, {4, 4, 3, 0}, {4, 4, 4, 0}, {4, 4, 5, 0}, {4, 5, 5, 0}};
    if (logn < 10 || logn > 14) return 0;
    int n = 0;
    for (int i = 0; i < 4; i++) {
        r[i] = inverse ? I[logn - 10][i] : F[logn - 10][i];
        if (r[i]) n++;
    }
    return n;
}

// ---------------------------------------------------------------- launchers
int ntt_kernel_smem_bytes(int logn) { return (1 << logn) * 8 + TWC * 8; } // polynomial + twiddle cache (FP64 path)

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

// `fp`: 0 = integer Harvey butterflies; NTT_FP = FP64 butterflies, optionally | NTT_IN_F (source holds lazy doubles) | NTT_OUT_F
// (destination receives lazy doubles instead of canonical words)
cudaError_t launch_ntt_forward(const u64 *src, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base, int mod_count, int fp,
                               cudaStream_t s) {
    if (n_polys <= 0) return cudaSuccess;
    if (fp & NTT_FP) {
        const bool lazy = (fp & (NTT_IN_F | NTT_OUT_F)) == (NTT_IN_F | NTT_OUT_F);
        if (!lazy && (fp & (NTT_IN_F | NTT_OUT_F))) return cudaErrorInvalidValue; // only canonical->canonical and lazy->lazy are built
        CNHE_DISPATCH_LOGN(logn, {
            if (lazy) {
                cudaError_t e = prep(k_ntt_forward_fp<L, true, true>, L);
                if (e != cudaSuccess) return e;
                k_ntt_forward_fp<L, true, true><<<n_polys, fp_threads(L), ntt_kernel_smem_bytes(L), s>>>(src, dst, tabs, mod_base, mod_count);
            } else {
                cudaError_t e = prep(k_ntt_forward_fp<L, false, false>, L);
                if (e != cudaSuccess) return e;
                k_ntt_forward_fp<L, false, false><<<n_polys, fp_threads(L), ntt_kernel_smem_bytes(L), s>>>(src, dst, tabs, mod_base, mod_count);
            }
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
cudaError_t launch_ntt_forward_digits(const u64 *target, size_t ct_stride, u64 *dst, int n_ct, int k, const DigitMap &dm, int logn,
                                      const NttTab *tabs, int fp, cudaStream_t s) {
    if (n_ct <= 0) return cudaSuccess;
    if (fp & NTT_FP) {
        if (fp & NTT_IN_F) return cudaErrorInvalidValue; // digits are cut from canonical words
        CNHE_DISPATCH_LOGN(logn, {
            if (fp & NTT_OUT_F) {
                cudaError_t e = prep(k_ntt_forward_digits_fp<L, true>, L);
                if (e != cudaSuccess) return e;
                k_ntt_forward_digits_fp<L, true><<<n_ct * dm.D * k, fp_threads(L), ntt_kernel_smem_bytes(L), s>>>(target, ct_stride, dst, tabs, k, dm);
            } else {
                cudaError_t e = prep(k_ntt_forward_digits_fp<L, false>, L);
                if (e != cudaSuccess) return e;
                k_ntt_forward_digits_fp<L, false><<<n_ct * dm.D * k, fp_threads(L), ntt_kernel_smem_bytes(L), s>>>(target, ct_stride, dst, tabs, k, dm);
            }
        });
        return cudaGetLastError();
    }
    CNHE_DISPATCH_LOGN(logn, {
        cudaError_t e = prep(k_ntt_forward_digits<L>, L);
        if (e != cudaSuccess) return e;
        k_ntt_forward_digits<L><<<n_ct * dm.D * k, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(target, ct_stride, dst, tabs, k, dm);
    });
    return cudaGetLastError();
}
template <int L, bool IN_F, bool OUT_F>
static cudaError_t launch_inv_fp(const u64 *src, const u64 *base, int base_group, size_t base_stride, u64 *dst, int n_polys, const NttTab *tabs,
                                 int mod_base, int mod_count, cudaStream_t s) {
    cudaError_t e = prep(k_ntt_inverse_fp<L, IN_F, OUT_F>, L);
    if (e != cudaSuccess) return e;
    k_ntt_inverse_fp<L, IN_F, OUT_F><<<n_polys, fp_threads(L), ntt_kernel_smem_bytes(L), s>>>(src, base, base_group, base_stride, dst, tabs, mod_base,
                                                                                            mod_count);
    return cudaSuccess;
}
static cudaError_t launch_inv(const u64 *src, const u64 *base, int base_group, size_t base_stride, u64 *dst, int n_polys, int logn,
                              const NttTab *tabs, int mod_base, int mod_count, int fp, cudaStream_t s) {
    if (n_polys <= 0) return cudaSuccess;
    if (base_group < 1) base_group = 1;
    if (fp & NTT_FP) {
        const bool in_f = fp & NTT_IN_F, out_f = fp & NTT_OUT_F;
        if (out_f && (!in_f || base)) return cudaErrorInvalidValue; // built: canonical->canonical, lazy->lazy, lazy->canonical(+base)
        CNHE_DISPATCH_LOGN(logn, {
            cudaError_t e = out_f  ? launch_inv_fp<L, true, true>(src, base, base_group, base_stride, dst, n_polys, tabs, mod_base, mod_count, s)
                            : in_f ? launch_inv_fp<L, true, false>(src, base, base_group, base_stride, dst, n_polys, tabs, mod_base, mod_count, s)
                                   : launch_inv_fp<L, false, false>(src, base, base_group, base_stride, dst, n_polys, tabs, mod_base, mod_count, s);
            if (e != cudaSuccess) return e;
        });
        return cudaGetLastError();
    }
    CNHE_DISPATCH_LOGN(logn, {
        cudaError_t e = prep(k_ntt_inverse<L>, L);
        if (e != cudaSuccess) return e;
        k_ntt_inverse<L><<<n_polys, (1 << L) / 16, ntt_kernel_smem_bytes(L), s>>>(src, base, base_group, base_stride, dst, tabs, mod_base, mod_count);
    });
    return cudaGetLastError();
}
cudaError_t launch_ntt_inverse(const u64 *src, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base, int mod_count, int fp,
                               cudaStream_t s) {
    return launch_inv(src, nullptr, 1, 0, dst, n_polys, logn, tabs, mod_base, mod_count, fp, s);
}
cudaError_t launch_ntt_inverse_add(const u64 *src, const u64 *base, int base_group, size_t base_stride, u64 *dst, int n_polys, int logn,
                                   const NttTab *tabs, int mod_base, int mod_count, int fp, cudaStream_t s) {
    return launch_inv(src, base, base_group, base_stride, dst, n_polys, logn, tabs, mod_base, mod_count, fp, s);
}

} // namespace cnhe
