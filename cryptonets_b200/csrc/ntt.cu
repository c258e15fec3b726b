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
#include <cstring>

#include <cuda.h> // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link dependency)

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
// MINB: resident CTAs per SM the register allocation aims for (0 = fp_min_blocks).  N = 8192 with three (80 registers, 6 doubles spilled)
// instead of two: a third CTA's memory phases fill the FP64 pipe's idle slots
template <int LOGN, bool IN_F, bool OUT_F, int MINB = 0>
__global__ void __launch_bounds__(fp_threads(LOGN), MINB ? MINB : fp_min_blocks(LOGN))
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
template <int LOGN, bool OUT_F, int MINB = 0>
__global__ void __launch_bounds__(fp_threads(LOGN), MINB ? MINB : fp_min_blocks(LOGN))
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
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// `after_load` runs once every input of the call sits in registers (the persistent kernels release / refill the shared-memory slot there)
template <int LOGN, int V0, int R, bool LAST, bool OUT_F, int NV = 1, class Hook = NoHook>
__device__ __forceinline__ void inv_pass_fp(double *sm, const double *twc, u64 *dst, const u64 *base_add, const NttTab &tb, int vt, int vstride = 0,
                                            Hook after_load = Hook()) {
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
        after_load();
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

// ================================================================ N = 16384 on CTA pairs ("split")
// A 16384-point polynomial is 128 KB of doubles: one CTA per SM, every warp of the SM at the same barrier, loads never overlapping
// butterflies (0.36 of the HBM roofline against 0.50 at N = 8192).  After the first Cooley-Tukey stage the two halves of the polynomial
// are independent 8192-point transforms with their own twiddle tables (NttTab::wd_hi holds them), so a cluster of two CTAs takes one
// polynomial: CTA h computes half h -- x[i] +- w x[i + N/2] on the way in from global memory (the pair reads the same lines at the same
// time: one trip to HBM, the second read is an L2 hit), then the 5+4+4 passes of the 8192-point kernel in 64 KB of shared memory, 2-3 CTAs
// per SM.  The inverse runs the 13 in-half stages first and the pair exchanges the halves through distributed shared memory for the last
// butterfly, which carries N^-1.  src == dst is allowed: the pair meets at a cluster barrier between its loads and its stores.
constexpr int SPLIT_LOGN = 13, SPLIT_H = 1 << SPLIT_LOGN, SPLIT_THREADS = SPLIT_H / 32;
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned dsmem_base(const void *p, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ double ld_dsmem(unsigned addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
// first pass of half `upper`: stage 0 of the 16384-point transform folded into the loads, then 5 stages (j = 0: low twiddles only)
template <bool IN_F>
__device__ __forceinline__ void fwd_split_first(double *sm, const double *twc, const FwdSrc &src, const NttTab &tb, int vt, double w0, bool upper) {
    constexpr int R = 5, E = 1 << R, LG = SPLIT_LOGN - R;
    const double p = tb.pd, pinv = tb.pinv;
    double x[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int idx = vt + (e << LG);
        const double a = fwd_load_fp<IN_F>(src, idx, p, pinv), c = fwd_load_fp<IN_F>(src, idx + SPLIT_H, p, pinv);
        const double t = fmodmul(c, w0, p, pinv);
        x[e] = upper ? __dsub_rn(a, t) : __dadd_rn(a, t);
    }
#pragma unroll
    for (int u = 0; u < R; u++) {
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
template <bool IN_F, bool OUT_F>
__device__ __forceinline__ void fwd_split_body(double *sm, const FwdSrc &src, u64 *dst, NttTab &tb, int half, int tid) {
    constexpr int TR = SPLIT_THREADS;
    double *twc = sm + SPLIT_H;
    const double w0 = __ldg(tb.wd + 1);
    tb.wd = tb.wd_hi + half * SPLIT_H;
    tb.fwd_recenter = tb.fwd_recenter_split;
    load_twiddle_cache(twc, tb.wd, tid, TR);
    __syncthreads();
    fwd_split_first<IN_F>(sm, twc, src, tb, tid, w0, half != 0);
    cluster_arrive(); // this CTA has read everything it needs from the source polynomial
    __syncthreads();
    CNHE_VTN(SPLIT_H / 16, (fwd_pass_fp<SPLIT_LOGN, 5, 4, false, 1, false>(sm, twc, src, tb, vt))); __syncthreads();
    cluster_wait(); // ... and so has its partner: the halves may be written in place
    CNHE_VTN(SPLIT_H / 16, (fwd_last_fp<SPLIT_LOGN, 2, OUT_F>(sm, dst, tb, vt)));
}
template <bool IN_F, bool OUT_F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SPLIT_THREADS, 2)
k_ntt_forward_split(const u64 *src, u64 *dst, const NttTab *__restrict__ tabs, int mod_base, int mod_count) {
    extern __shared__ __align__(16) u64 sm[];
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, tid = threadIdx.x;
    NttTab tb = tabs[mod_base + b % mod_count];
    FwdSrc fs;
    fs.src = src + (size_t)b * (2 * SPLIT_H);
    fs.digit = false; fs.need_reduce = false; fs.shift = 0; fs.mask = 0;
    fwd_split_body<IN_F, OUT_F>(reinterpret_cast<double *>(sm), fs, dst + (size_t)b * (2 * SPLIT_H) + half * SPLIT_H, tb, half, tid);
}
template <bool OUT_F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SPLIT_THREADS, 2)
k_ntt_forward_digits_split(const u64 *target, size_t ct_stride, u64 *dst, const NttTab *__restrict__ tabs, int k, DigitMap dm) {
    extern __shared__ __align__(16) u64 sm[];
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, tid = threadIdx.x;
    const int l = b % k, d = (b / k) % dm.D, c = b / (k * dm.D);
    NttTab tb = tabs[l];
    FwdSrc fs;
    fs.src = target + (size_t)c * ct_stride + (size_t)dm.src[d] * (2 * SPLIT_H);
    fs.digit = true;
    fs.shift = dm.shift[d];
    fs.mask = dm.mask;
    fs.need_reduce = dm.mask >= tb.mod.p;
    fwd_split_body<false, OUT_F>(reinterpret_cast<double *>(sm), fs, dst + (((size_t)c * k + l) * dm.D + d) * (2 * SPLIT_H) + half * SPLIT_H, tb, half, tid);
}
// last in-half pass (stages 8..12) and the cross-half butterfly: own results go to shared memory for the partner, the partner's come
// back through DSMEM; half 0 keeps the sums (times N^-1), half 1 the differences (times iw[1] N^-1)
template <bool OUT_F>
__device__ __forceinline__ void inv_split_last(double *sm, const double *twc, u64 *dst, const u64 *base_add, const NttTab &tb, int vt, int half) {
    constexpr int V0 = 8, R = 5, E = 1 << R, N = SPLIT_H;
    const double p = tb.pd, pinv = tb.pinv;
    double x[E];
#pragma unroll
    for (int e = 0; e < E; e++) x[e] = sm[swz(vt + (e << V0))];
#pragma unroll
    for (int u = 0; u < R; u++) {
        const int h = 1 << u;
        const bool rc = (tb.inv_recenter >> (V0 + u)) & 1;
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (e & h) continue;
            const double w = twc[(N >> (V0 + u + 1)) + (e >> (u + 1))];
            const double a = x[e], bq = x[e + h];
            x[e] = __dadd_rn(a, bq);
            x[e + h] = fmodmul(__dsub_rn(a, bq), w, p, pinv);
        }
        if (rc) {
#pragma unroll
            for (int e = 0; e < E; e++)
                if (!(e & h)) x[e] = frecenter(x[e], p, pinv);
        }
    }
#pragma unroll
    for (int e = 0; e < E; e++) sm[swz(vt + (e << V0))] = x[e];
    cluster_arrive();
    cluster_wait(); // both halves are complete and visible across the pair
    const unsigned peer = dsmem_base(sm, (unsigned)(half ^ 1));
    const double scale = half ? tb.inv_n_w_d : tb.inv_n_d;
#pragma unroll
    for (int e0 = 0; e0 < E; e0 += 8) { // the partner's values in batches of 8: all 32 next to x[] would spill
        double r[8];
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = ld_dsmem(peer + 8u * (unsigned)swz(vt + ((e0 + i) << V0)));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int idx = vt + ((e0 + i) << V0);
            const double v = fmodmul(half ? __dsub_rn(r[i], x[e0 + i]) : __dadd_rn(x[e0 + i], r[i]), scale, p, pinv);
            if constexpr (OUT_F) dst[idx] = lazy_bits(v);
            else {
                u64 o = fsmall_u(v, tb.mod.p);
                if (base_add) o = addmod(o, base_add[idx], tb.mod.p);
                dst[idx] = o;
            }
        }
    }
    cluster_arrive(); // done with the partner's memory: either CTA may exit once both have said so
    cluster_wait();
}
template <bool IN_F, bool OUT_F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SPLIT_THREADS, 2)
k_ntt_inverse_split(const u64 *src, const u64 *base_add, int base_group, size_t base_stride, u64 *dst, const NttTab *__restrict__ tabs, int mod_base,
                    int mod_count) {
    extern __shared__ __align__(16) u64 smraw[];
    constexpr int TR = SPLIT_THREADS, H = SPLIT_H;
    const int b = blockIdx.x >> 1, half = blockIdx.x & 1, tid = threadIdx.x;
    NttTab tb = tabs[mod_base + b % mod_count];
    tb.iwd = tb.iwd_hi + half * H;
    double *sm = reinterpret_cast<double *>(smraw);
    double *twc = sm + H;
    load_twiddle_cache(twc, tb.iwd, tid, TR);
    const u64 *s = src + (size_t)b * (2 * H) + half * H;
    u64 *d = dst + (size_t)b * (2 * H) + half * H;
    const u64 *ba = base_add ? base_add + (size_t)(b / base_group) * base_stride + (size_t)(b % base_group) * (2 * H) + half * H : nullptr;
    if (ba) {
        const char *pb = reinterpret_cast<const char *>(ba);
        for (int i = tid * 128; i < H * 8; i += TR * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + i));
    }
    CNHE_VTN(H / 16, (inv_first_fp<SPLIT_LOGN, IN_F>(sm, s, tb, vt)));
    __syncthreads();
    CNHE_VTN(H / 16, (inv_pass_fp<SPLIT_LOGN, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); __syncthreads();
    inv_split_last<OUT_F>(sm, twc, d, ba, tb, tid, half);
}

// ================================================================ persistent TMA-staged transforms, N = 4096 / 8192
// One persistent CTA per SM.  Each CTA is pinned to ONE modulus (CTA c serves the polynomials whose table index is c mod #moduli), so
// the twiddles it needs never change: the 15N/16 twiddles of the four unit-stride stages -- the ones that used to be fetched from L2 by
// every polynomial (as many bytes as the polynomial itself, and the top stall of the one-CTA-per-polynomial kernel) -- are staged into
// shared memory ONCE per CTA with cp.async.bulk, transposed so that lane j reads word m*T + j (conflict-free), next to the 512 low
// twiddles of the strided stages.  Polynomials arrive by cp.async.bulk.tensor (a 2-D tensor map over the source array, 128-byte rows,
// SWIZZLE_128B -- exactly the XOR layout swz() the butterfly passes use, so the hardware produces the bank-conflict-free layout on the
// way in), completing on an mbarrier.  Two groups of 8 warps each own one polynomial slot: three fat passes in place between named
// barriers of the group, results written straight from registers, then one elected thread refills the slot with the group's next
// polynomial.  No compute warp ever waits on HBM or L2 for data or twiddles, there is no CTA launch/exit (store drain) per polynomial,
// and the groups run out of phase so that one group's shared-memory round trips and slot refill hide under the other's FP64 work.
constexpr int WS_GROUPS = 2, WS_GROUP_THREADS = 256, WS_THREADS = WS_GROUPS * WS_GROUP_THREADS;
enum WsMode { WS_CANON = 0, WS_LAZY = 1, WS_DIGIT = 2 }; // how the first pass reads the staged words

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (a suspend-time hint on the try_wait -- the waiting warps sleep instead of spinning -- measured 3 % slower here: the wake-up latency
// costs more than the issue slots the spinning takes from the other group; mac_umma.cu, with ten mostly-waiting warps, keeps the hint)
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_rows(void *dst, const void *tmap, int row, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tmap), "r"(0), "r"(row), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ws_delay(unsigned ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < ns);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(WS_GROUP_THREADS) : "memory"); }

struct WsJob {
    long long row;  // first 128-byte row of the source polynomial in the tensor map
    size_t dst_off; // destination offset in words
    int shift;      // digit mode
};
struct WsArgs {
    int n_polys, mod_base, mod_count; // plain: polynomial b uses table mod_base + b % mod_count, source row b * N/16, destination b * N
    int D;                            // digit mode: b -> (c, d, l) as in k_ntt_forward_digits_fp, mod_count = k, mod_base = 0
    long long ct_stride_rows;
    unsigned char src[64], shift[64];
    u64 mask;
    // inverse: optional base added to the canonical result
    const u64 *base_add;
    int base_group;
    size_t base_stride;
    unsigned stagger_ns; // group 1 starts this much later than group 0 (the groups should not walk through the passes in lockstep)
};
template <int LOGN, int MODE>
__device__ __forceinline__ WsJob ws_job(const WsArgs &a, int b) {
    constexpr int N = 1 << LOGN;
    WsJob j;
    if constexpr (MODE == WS_DIGIT) {
        const int l = b % a.mod_count, d = (b / a.mod_count) % a.D, c = b / (a.mod_count * a.D);
        j.row = (long long)c * a.ct_stride_rows + (long long)a.src[d] * (N / 16);
        j.dst_off = (((size_t)c * a.mod_count + l) * a.D + d) * N;
        j.shift = a.shift[d];
    } else {
        j.row = (long long)b * (N / 16);
        j.dst_off = (size_t)b * N;
        j.shift = 0;
    }
    return j;
}
// stage polynomial `b` into a slot; called by one thread
template <int LOGN, int MODE>
__device__ __forceinline__ void ws_issue(const CUtensorMap *tmap, const WsArgs &a, double *slot, unsigned long long *full, int b) {
    constexpr int N = 1 << LOGN;
    constexpr int BOX_ROWS = 256, BOXES = (N / 16) / BOX_ROWS;
    const WsJob j = ws_job<LOGN, MODE>(a, b);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the slot's previous contents were read through the generic proxy
    mbar_expect_tx(full, (unsigned)(N * 8));
#pragma unroll
    for (int x = 0; x < BOXES; x++) tma_load_rows(slot + x * BOX_ROWS * 16, tmap, (int)(j.row + x * BOX_ROWS), full);
}
// first forward pass on the staged words: R stages on 2^R coefficients at stride N >> R, in place
template <int LOGN, int R, int MODE>
__device__ __forceinline__ void fwd_first_staged(double *sm, const double *twc, int shift, u64 mask, bool need_reduce, const NttTab &tb, int vt) {
    constexpr int E = 1 << R, LG = LOGN - R;
    const double p = tb.pd, pinv = tb.pinv;
    double x[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const double raw = sm[swz(vt + (e << LG))];
        if constexpr (MODE == WS_LAZY) x[e] = raw;
        else {
            u64 v = (u64)__double_as_longlong(raw);
            if constexpr (MODE == WS_DIGIT) v = (v >> shift) & mask;
            x[e] = u2d(v);
            if (MODE == WS_DIGIT && need_reduce) x[e] = frecenter(x[e], p, pinv);
        }
    }
#pragma unroll
    for (int u = 0; u < R; u++) {
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
// last forward pass with the unit-stride twiddles resident in shared memory (hi[m * T + j], m = 0..14): NV 16-coefficient groups per
// thread (j0, j0 + jstride, ..); all of them are read before `after_load` runs, so the slot can be refilled under the arithmetic
template <int LOGN, int PASS, bool OUT_F, int NV, class Hook>
__device__ __forceinline__ void fwd_last_staged(const double *sm, u64 *dst, const NttTab &tb, const double *hi, int j0, int jstride, Hook after_load) {
    constexpr int T = (1 << LOGN) / 16;
    const double p = tb.pd, pinv = tb.pinv;
    const bool rc = (tb.fwd_recenter >> PASS) & 1;
    const double2 *smv = reinterpret_cast<const double2 *>(sm);
    double xs[NV][16];
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const int j = j0 + i * jstride, xr = j & 7;
#pragma unroll
        for (int ch = 0; ch < 8; ch++) {
            double2 v = smv[j * 8 + (ch ^ xr)];
            xs[i][2 * ch] = v.x;
            xs[i][2 * ch + 1] = v.y;
        }
    }
    after_load();
#pragma unroll
    for (int i = 0; i < NV; i++) {
    const int j = j0 + i * jstride;
    double (&x)[16] = xs[i];
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
            const double w = hi[((1 << u) - 1 + (e >> (4 - u))) * T + j];
            const double t = fmodmul(x[e + h], w, p, pinv);
            const double a = x[e];
            x[e] = __dadd_rn(a, t);
            x[e + h] = __dsub_rn(a, t);
        }
    }
    u64 *o = dst + 16 * j;
    if constexpr (OUT_F) {
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
}
// first inverse pass on the staged words: stages 0..3 on 16 consecutive coefficients, in place; twiddles from the resident table
template <int LOGN, bool IN_F>
__device__ __forceinline__ void inv_first_staged(double *sm, const NttTab &tb, const double *hi, int j) {
    constexpr int T = (1 << LOGN) / 16;
    const double p = tb.pd, pinv = tb.pinv;
    double2 *smv = reinterpret_cast<double2 *>(sm);
    const int xr = j & 7;
    double x[16];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) {
        const double2 v = smv[j * 8 + (ch ^ xr)];
        if constexpr (IN_F) { x[2 * ch] = v.x; x[2 * ch + 1] = v.y; }
        else { x[2 * ch] = u2d((u64)__double_as_longlong(v.x)); x[2 * ch + 1] = u2d((u64)__double_as_longlong(v.y)); }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int h = 1 << u;
        const bool rc = (tb.inv_recenter >> u) & 1;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            if (e & h) continue;
            const double w = hi[((16 - (16 >> u)) + (e >> (u + 1))) * T + j];
            const double a = x[e], bq = x[e + h];
            x[e] = __dadd_rn(a, bq);
            x[e + h] = fmodmul(__dsub_rn(a, bq), w, p, pinv);
        }
        if (rc) {
#pragma unroll
            for (int e = 0; e < 16; e++)
                if (!(e & h)) x[e] = frecenter(x[e], p, pinv);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ch++) smv[j * 8 + (ch ^ xr)] = make_double2(x[2 * ch], x[2 * ch + 1]);
}

#define CNHE_WS_VT(COUNT, stmt) _Pragma("unroll") for (int vt = gt; vt < (COUNT); vt += WS_GROUP_THREADS) { stmt; }
template <int LOGN>
__host__ __device__ constexpr int ws_hi_words() { return 15 * (1 << LOGN) / 16; }
template <int LOGN>
__host__ __device__ constexpr int ws_smem_bytes() { return (WS_GROUPS * (1 << LOGN) + ws_hi_words<LOGN>() + TWC) * 8 + 64 + (int)sizeof(NttTab) + 1024; }

// shared prologue: carve shared memory, stage the CTA's twiddle tables and the first polynomial of each group
struct WsSmem {
    double *slots, *hi, *lo; // slot g = slots + g * N (plain pointer arithmetic on the __shared__ symbol keeps the address space)
    unsigned long long *full, *tbar, *empty;
    NttTab *tab; // the CTA's modulus record, copied once: per-polynomial reads are LDS instead of (L1-missing) global loads
};
template <int LOGN>
__device__ __forceinline__ WsSmem ws_carve(unsigned char *raw) {
    constexpr int N = 1 << LOGN;
    WsSmem w;
    // SWIZZLE_128B wants 1024-byte aligned slots.  The padding is added to the __shared__ symbol itself (no integer round trip), so the
    // compiler still knows every derived pointer is shared memory and emits LDS/STS instead of generic loads
    const unsigned pad = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
    double *base = reinterpret_cast<double *>(raw + pad);
    w.slots = base;
    w.hi = base + (size_t)WS_GROUPS * N;
    w.lo = w.hi + ws_hi_words<LOGN>();
    w.full = reinterpret_cast<unsigned long long *>(w.lo + TWC);
    w.tbar = w.full + WS_GROUPS;
    w.empty = w.tbar + 1;
    w.tab = reinterpret_cast<NttTab *>(w.empty + WS_GROUPS);
    return w;
}
// which polynomials this CTA serves: class m = blockIdx % mc (its modulus), the r-th CTA of the S CTAs of that class takes j = r, r + S, ...
struct WsWalk {
    int m, mc, r, S;
    __device__ __forceinline__ int poly(int s) const { return m + mc * (r + s * S); } // sequence number -> polynomial index (may run past n_polys)
};
__device__ __forceinline__ WsWalk ws_walk(int mc) {
    WsWalk w;
    w.mc = mc;
    w.m = blockIdx.x % mc;
    w.r = blockIdx.x / mc;
    w.S = ((int)gridDim.x - w.m + mc - 1) / mc;
    return w;
}

template <int LOGN, int MODE, bool OUT_F>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_ntt_forward_ws(const __grid_constant__ CUtensorMap tmap, u64 *dst, const NttTab *__restrict__ tabs, const WsArgs a) {
    static_assert(LOGN == 12 || LOGN == 13, "the staged transform covers N = 4096 and 8192");
    constexpr int N = 1 << LOGN;
    extern __shared__ unsigned char ws_raw[];
    const int tid = threadIdx.x;
    {
        const WsSmem sm_ = ws_carve<LOGN>(ws_raw);
        const WsWalk walk = ws_walk(a.mod_count);
        if (tid == 0) {
            const NttTab &tb0 = tabs[a.mod_base + walk.m];
            *sm_.tab = tb0;
            for (int g = 0; g < WS_GROUPS; g++) { mbar_init(&sm_.full[g], 1); mbar_init(&sm_.empty[g], WS_GROUP_THREADS); }
            mbar_init(sm_.tbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(sm_.tbar, (unsigned)((ws_hi_words<LOGN>() + TWC) * 8));
            bulk_load(sm_.hi, tb0.wd_hi, ws_hi_words<LOGN>() * 8, sm_.tbar);
            bulk_load(sm_.lo, tb0.wd, TWC * 8, sm_.tbar);
            for (int g = 0; g < WS_GROUPS; g++)
                if (walk.poly(g) < a.n_polys) ws_issue<LOGN, MODE>(&tmap, a, sm_.slots + (size_t)g * N, &sm_.full[g], walk.poly(g));
        }
        __syncthreads();
        mbar_wait(sm_.tbar, 0);
    }
    const int g = tid / WS_GROUP_THREADS, gt = tid % WS_GROUP_THREADS;
    if (g == 1 && a.stagger_ns) ws_delay(a.stagger_ns);
    FwdSrc unused;
    unused.src = nullptr; unused.digit = false; unused.need_reduce = false; unused.shift = 0; unused.mask = 0;
#pragma unroll 1
    for (int t = 0;; t++) {
        // everything below is re-derived per polynomial on purpose: nothing but t stays live across the register-hungry radix-32 pass
        const WsWalk walk = ws_walk(a.mod_count);
        const int b = walk.poly(g + WS_GROUPS * t);
        if (b >= a.n_polys) break;
        const WsSmem sm_ = ws_carve<LOGN>(ws_raw);
        const NttTab &tb = *sm_.tab;
        double *sm = sm_.slots + (size_t)g * N;
        const double *twc = sm_.lo, *hi = sm_.hi;
        const bool need_reduce = MODE == WS_DIGIT && a.mask >= tb.mod.p;
        const WsJob job = ws_job<LOGN, MODE>(a, b);
        // once the last pass holds its inputs in registers the slot is dead: refill it with the group's next polynomial under the arithmetic
        auto refill = [&]() { // every thread reports its reads done; only the elected thread waits for all of them before issuing the copy
            mbar_arrive(&sm_.empty[g]);
            const int nb = walk.poly(g + WS_GROUPS * (t + 1));
            if (gt == 0 && nb < a.n_polys) {
                mbar_wait(&sm_.empty[g], t & 1);
                ws_issue<LOGN, MODE>(&tmap, a, sm, &sm_.full[g], nb);
            }
        };
        mbar_wait(&sm_.full[g], t & 1);
        if constexpr (LOGN == 13) {
            CNHE_WS_VT(N / 32, (fwd_first_staged<13, 5, MODE>(sm, twc, job.shift, a.mask, need_reduce, tb, vt))); group_sync(g);
            CNHE_WS_VT(N / 16, (fwd_pass_fp<13, 5, 4, false, 1, false>(sm, twc, unused, tb, vt))); group_sync(g);
            fwd_last_staged<13, 2, OUT_F, 2>(sm, dst + job.dst_off, tb, hi, gt, WS_GROUP_THREADS, refill);
        } else {
            CNHE_WS_VT(N / 16, (fwd_first_staged<12, 4, MODE>(sm, twc, job.shift, a.mask, need_reduce, tb, vt))); group_sync(g);
            CNHE_WS_VT(N / 16, (fwd_pass_fp<12, 4, 4, false, 1, false>(sm, twc, unused, tb, vt))); group_sync(g);
            fwd_last_staged<12, 2, OUT_F, 1>(sm, dst + job.dst_off, tb, hi, gt, WS_GROUP_THREADS, refill);
        }
    }
}
template <int LOGN, bool IN_F, bool OUT_F>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_ntt_inverse_ws(const __grid_constant__ CUtensorMap tmap, u64 *dst, const NttTab *__restrict__ tabs, const WsArgs a) {
    static_assert(LOGN == 12 || LOGN == 13, "the staged transform covers N = 4096 and 8192");
    constexpr int N = 1 << LOGN;
    extern __shared__ unsigned char ws_raw[];
    const int tid = threadIdx.x;
    {
        const WsSmem sm_ = ws_carve<LOGN>(ws_raw);
        const WsWalk walk = ws_walk(a.mod_count);
        if (tid == 0) {
            const NttTab &tb0 = tabs[a.mod_base + walk.m];
            *sm_.tab = tb0;
            for (int g = 0; g < WS_GROUPS; g++) { mbar_init(&sm_.full[g], 1); mbar_init(&sm_.empty[g], WS_GROUP_THREADS); }
            mbar_init(sm_.tbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(sm_.tbar, (unsigned)((ws_hi_words<LOGN>() + TWC) * 8));
            bulk_load(sm_.hi, tb0.iwd_hi, ws_hi_words<LOGN>() * 8, sm_.tbar);
            bulk_load(sm_.lo, tb0.iwd, TWC * 8, sm_.tbar);
            for (int g = 0; g < WS_GROUPS; g++)
                if (walk.poly(g) < a.n_polys) ws_issue<LOGN, WS_CANON>(&tmap, a, sm_.slots + (size_t)g * N, &sm_.full[g], walk.poly(g));
        }
        __syncthreads();
        mbar_wait(sm_.tbar, 0);
    }
    const int g = tid / WS_GROUP_THREADS, gt = tid % WS_GROUP_THREADS;
    if (g == 1 && a.stagger_ns) ws_delay(a.stagger_ns);
#pragma unroll 1
    for (int t = 0;; t++) {
        const WsWalk walk = ws_walk(a.mod_count);
        const int b = walk.poly(g + WS_GROUPS * t);
        if (b >= a.n_polys) break;
        const WsSmem sm_ = ws_carve<LOGN>(ws_raw);
        const NttTab &tb = *sm_.tab;
        double *sm = sm_.slots + (size_t)g * N;
        const double *twc = sm_.lo, *hi = sm_.hi;
        u64 *d = dst + (size_t)b * N;
        const u64 *ba = a.base_add ? a.base_add + (size_t)(b / a.base_group) * a.base_stride + (size_t)(b % a.base_group) * N : nullptr;
        if (ba) { // consumed by the epilogue most of a transform from now: have it waiting in L2
            const char *pb = reinterpret_cast<const char *>(ba);
            for (int i = gt * 128; i < N * 8; i += WS_GROUP_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + i));
        }
        auto refill = [&]() {
            mbar_arrive(&sm_.empty[g]);
            const int nb = walk.poly(g + WS_GROUPS * (t + 1));
            if (gt == 0 && nb < a.n_polys) {
                mbar_wait(&sm_.empty[g], t & 1);
                ws_issue<LOGN, WS_CANON>(&tmap, a, sm, &sm_.full[g], nb);
            }
        };
        mbar_wait(&sm_.full[g], t & 1);
        CNHE_WS_VT(N / 16, (inv_first_staged<LOGN, IN_F>(sm, tb, hi, vt))); group_sync(g);
        if constexpr (LOGN == 13) {
            CNHE_WS_VT(N / 16, (inv_pass_fp<13, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); group_sync(g);
            inv_pass_fp<13, 8, 5, true, OUT_F, 1>(sm, twc, d, ba, tb, gt, 0, refill); // one radix-32 group per thread
        } else {
            CNHE_WS_VT(N / 16, (inv_pass_fp<12, 4, 4, false, false>(sm, twc, d, ba, tb, vt))); group_sync(g);
            inv_pass_fp<12, 8, 4, true, OUT_F, 1>(sm, twc, d, ba, tb, gt, 0, refill);
        }
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

// ---- host side of the staged transforms: tensor map over the source array, persistent grid of one CTA per SM
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// CNHE_NTT_WS=0 switches the persistent kernels off altogether; CNHE_NTT_WS_FWD / CNHE_NTT_WS_INV override one direction
static bool ws_flag(const char *name, bool dflt) {
    const char *all = getenv("CNHE_NTT_WS"), *one = getenv(name);
    if (encode_tiled() == nullptr) return false;
    if (one) return atoi(one) != 0;
    if (all) return atoi(all) != 0;
    return dflt;
}
// Measured on B200 (profiles/r02_ntt_ws_ab.txt): the persistent inverse transform is 20-25 % faster than one CTA per polynomial
// (0.53-0.60 of the HBM roofline against 0.43-0.49); the persistent forward transform only draws level (0.50 against 0.52, and 12 % slower
// in its digit-cutting form), so the forward direction keeps the per-polynomial kernel unless CNHE_NTT_WS_FWD=1 asks for the staged one.
static bool ws_enabled_fwd() { return ws_flag("CNHE_NTT_WS_FWD", false); } // read per launch: tests flip it inside one process
static bool ws_enabled_inv() { return ws_flag("CNHE_NTT_WS_INV", true); }
// N = 8192 forward transforms with three CTAs per SM (CNHE_NTT_FWD_BLOCKS=3) or two (=2); read per launch
static bool fwd_three_blocks() {
    const char *v = getenv("CNHE_NTT_FWD_BLOCKS");
    return v ? atoi(v) == 3 : false;
}
// N = 16384: CTA pairs unless CNHE_NTT_SPLIT=0 (read per launch, like the flags above)
static bool split_enabled() {
    const char *v = getenv("CNHE_NTT_SPLIT");
    return v ? atoi(v) != 0 : true;
}
constexpr int SPLIT_SMEM = SPLIT_H * 8 + TWC * 8;
template <class K>
static cudaError_t split_prep(K kern) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLIT_SMEM);
}
static int sm_count() {
    static int n = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}
// rows of 16 words (128 bytes) starting at `base`; boxes of 256 rows, hardware swizzle = swz()
static cudaError_t make_row_map(CUtensorMap *m, const u64 *base, size_t words) {
    const cuuint64_t dims[2] = {16, (cuuint64_t)(words / 16)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {16, 256}, estr[2] = {1, 1};
    if (words % 16 || dims[1] < 256 || (reinterpret_cast<uintptr_t>(base) & 15)) return cudaErrorInvalidValue;
    const CUresult r = encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<u64 *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
// General 2-D map over 64-bit words for other kernels (mac_umma.cu): `rows` rows of `inner_words` words, `row_stride` bytes apart, box of
// box_words x box_rows, optionally SWIZZLE_128B (box rows of 128 bytes), out-of-bounds rows read as zero.  `map` points at a CUtensorMap (128 bytes, 64-byte aligned).
cudaError_t make_word_map_2d(void *map, const u64 *base, size_t inner_words, size_t rows, size_t row_stride, unsigned box_words, unsigned box_rows,
                             int swizzle128) {
    if (encode_tiled() == nullptr) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)inner_words, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_stride};
    const cuuint32_t box[2] = {box_words, box_rows}, estr[2] = {1, 1};
    if ((row_stride & 15) || (reinterpret_cast<uintptr_t>(base) & 15) || box_words > 256 || box_rows > 256 || (swizzle128 && box_words != 16))
        return cudaErrorInvalidValue;
    const CUresult r = encode_tiled()(reinterpret_cast<CUtensorMap *>(map), CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<u64 *>(base), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
static unsigned ws_stagger_ns() {
    static const unsigned v = getenv("CNHE_WS_STAGGER") ? (unsigned)atoi(getenv("CNHE_WS_STAGGER")) : 5000u;
    return v;
}
template <class K>
static cudaError_t ws_launch(K kern, int smem, const CUtensorMap &map, u64 *dst, const NttTab *tabs, const WsArgs &a0, cudaStream_t s) {
    WsArgs a = a0;
    a.stagger_ns = ws_stagger_ns();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    int grid = sm_count();
    if (a.mod_count <= grid) grid -= grid % a.mod_count; // equally many CTAs per modulus: every CTA is pinned to one
    if (a.n_polys < grid) grid = a.n_polys;
    kern<<<grid, WS_THREADS, smem, s>>>(map, dst, tabs, a);
    return cudaGetLastError();
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
        if ((logn == 12 || logn == 13) && ws_enabled_fwd()) {
            CUtensorMap map;
            cudaError_t e = make_row_map(&map, src, (size_t)n_polys << logn);
            if (e != cudaSuccess) return e;
            WsArgs a;
            memset(&a, 0, sizeof(a));
            a.n_polys = n_polys; a.mod_base = mod_base; a.mod_count = mod_count;
            if (logn == 13) return lazy ? ws_launch(k_ntt_forward_ws<13, WS_LAZY, true>, ws_smem_bytes<13>(), map, dst, tabs, a, s)
                                        : ws_launch(k_ntt_forward_ws<13, WS_CANON, false>, ws_smem_bytes<13>(), map, dst, tabs, a, s);
            return lazy ? ws_launch(k_ntt_forward_ws<12, WS_LAZY, true>, ws_smem_bytes<12>(), map, dst, tabs, a, s)
                        : ws_launch(k_ntt_forward_ws<12, WS_CANON, false>, ws_smem_bytes<12>(), map, dst, tabs, a, s);
        }
        if (logn == 14 && split_enabled()) {
            if (lazy) {
                cudaError_t e = split_prep(k_ntt_forward_split<true, true>);
                if (e != cudaSuccess) return e;
                k_ntt_forward_split<true, true><<<2 * n_polys, SPLIT_THREADS, SPLIT_SMEM, s>>>(src, dst, tabs, mod_base, mod_count);
            } else {
                cudaError_t e = split_prep(k_ntt_forward_split<false, false>);
                if (e != cudaSuccess) return e;
                k_ntt_forward_split<false, false><<<2 * n_polys, SPLIT_THREADS, SPLIT_SMEM, s>>>(src, dst, tabs, mod_base, mod_count);
            }
            return cudaGetLastError();
        }
        if (logn == 13 && fwd_three_blocks()) {
            if (lazy) {
                cudaError_t e = prep(k_ntt_forward_fp<13, true, true, 3>, 13);
                if (e != cudaSuccess) return e;
                k_ntt_forward_fp<13, true, true, 3><<<n_polys, fp_threads(13), ntt_kernel_smem_bytes(13), s>>>(src, dst, tabs, mod_base, mod_count);
            } else {
                cudaError_t e = prep(k_ntt_forward_fp<13, false, false, 3>, 13);
                if (e != cudaSuccess) return e;
                k_ntt_forward_fp<13, false, false, 3><<<n_polys, fp_threads(13), ntt_kernel_smem_bytes(13), s>>>(src, dst, tabs, mod_base, mod_count);
            }
            return cudaGetLastError();
        }
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
        if ((logn == 12 || logn == 13) && ws_enabled_fwd() && ct_stride % 16 == 0) {
            CUtensorMap map;
            cudaError_t e = make_row_map(&map, target, (size_t)(n_ct - 1) * ct_stride + ((size_t)k << logn));
            if (e != cudaSuccess) return e;
            WsArgs a;
            memset(&a, 0, sizeof(a));
            a.n_polys = n_ct * dm.D * k; a.mod_base = 0; a.mod_count = k; a.D = dm.D; a.ct_stride_rows = (long long)(ct_stride / 16); a.mask = dm.mask;
            memcpy(a.src, dm.src, 64); memcpy(a.shift, dm.shift, 64);
            const bool of = fp & NTT_OUT_F;
            if (logn == 13) return of ? ws_launch(k_ntt_forward_ws<13, WS_DIGIT, true>, ws_smem_bytes<13>(), map, dst, tabs, a, s)
                                      : ws_launch(k_ntt_forward_ws<13, WS_DIGIT, false>, ws_smem_bytes<13>(), map, dst, tabs, a, s);
            return of ? ws_launch(k_ntt_forward_ws<12, WS_DIGIT, true>, ws_smem_bytes<12>(), map, dst, tabs, a, s)
                      : ws_launch(k_ntt_forward_ws<12, WS_DIGIT, false>, ws_smem_bytes<12>(), map, dst, tabs, a, s);
        }
        if (logn == 14 && split_enabled()) {
            if (fp & NTT_OUT_F) {
                cudaError_t e = split_prep(k_ntt_forward_digits_split<true>);
                if (e != cudaSuccess) return e;
                k_ntt_forward_digits_split<true><<<2 * n_ct * dm.D * k, SPLIT_THREADS, SPLIT_SMEM, s>>>(target, ct_stride, dst, tabs, k, dm);
            } else {
                cudaError_t e = split_prep(k_ntt_forward_digits_split<false>);
                if (e != cudaSuccess) return e;
                k_ntt_forward_digits_split<false><<<2 * n_ct * dm.D * k, SPLIT_THREADS, SPLIT_SMEM, s>>>(target, ct_stride, dst, tabs, k, dm);
            }
            return cudaGetLastError();
        }
        if (logn == 13 && fwd_three_blocks()) {
            if (fp & NTT_OUT_F) {
                cudaError_t e = prep(k_ntt_forward_digits_fp<13, true, 3>, 13);
                if (e != cudaSuccess) return e;
                k_ntt_forward_digits_fp<13, true, 3><<<n_ct * dm.D * k, fp_threads(13), ntt_kernel_smem_bytes(13), s>>>(target, ct_stride, dst, tabs, k, dm);
            } else {
                cudaError_t e = prep(k_ntt_forward_digits_fp<13, false, 3>, 13);
                if (e != cudaSuccess) return e;
                k_ntt_forward_digits_fp<13, false, 3><<<n_ct * dm.D * k, fp_threads(13), ntt_kernel_smem_bytes(13), s>>>(target, ct_stride, dst, tabs, k, dm);
            }
            return cudaGetLastError();
        }
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
template <bool IN_F, bool OUT_F>
static cudaError_t launch_inv_split(const u64 *src, const u64 *base, int base_group, size_t base_stride, u64 *dst, int n_polys, const NttTab *tabs,
                                    int mod_base, int mod_count, cudaStream_t s) {
    cudaError_t e = split_prep(k_ntt_inverse_split<IN_F, OUT_F>);
    if (e != cudaSuccess) return e;
    k_ntt_inverse_split<IN_F, OUT_F><<<2 * n_polys, SPLIT_THREADS, SPLIT_SMEM, s>>>(src, base, base_group, base_stride, dst, tabs, mod_base, mod_count);
    return cudaGetLastError();
}
static cudaError_t launch_inv(const u64 *src, const u64 *base, int base_group, size_t base_stride, u64 *dst, int n_polys, int logn,
                              const NttTab *tabs, int mod_base, int mod_count, int fp, cudaStream_t s) {
    if (n_polys <= 0) return cudaSuccess;
    if (base_group < 1) base_group = 1;
    if (fp & NTT_FP) {
        const bool in_f = fp & NTT_IN_F, out_f = fp & NTT_OUT_F;
        if (out_f && (!in_f || base)) return cudaErrorInvalidValue; // built: canonical->canonical, lazy->lazy, lazy->canonical(+base)
        if ((logn == 12 || logn == 13) && ws_enabled_inv()) {
            CUtensorMap map;
            cudaError_t e = make_row_map(&map, src, (size_t)n_polys << logn);
            if (e != cudaSuccess) return e;
            WsArgs a;
            memset(&a, 0, sizeof(a));
            a.n_polys = n_polys; a.mod_base = mod_base; a.mod_count = mod_count;
            a.base_add = base; a.base_group = base_group; a.base_stride = base_stride;
            if (logn == 13)
                return out_f  ? ws_launch(k_ntt_inverse_ws<13, true, true>, ws_smem_bytes<13>(), map, dst, tabs, a, s)
                       : in_f ? ws_launch(k_ntt_inverse_ws<13, true, false>, ws_smem_bytes<13>(), map, dst, tabs, a, s)
                              : ws_launch(k_ntt_inverse_ws<13, false, false>, ws_smem_bytes<13>(), map, dst, tabs, a, s);
            return out_f  ? ws_launch(k_ntt_inverse_ws<12, true, true>, ws_smem_bytes<12>(), map, dst, tabs, a, s)
                   : in_f ? ws_launch(k_ntt_inverse_ws<12, true, false>, ws_smem_bytes<12>(), map, dst, tabs, a, s)
                          : ws_launch(k_ntt_inverse_ws<12, false, false>, ws_smem_bytes<12>(), map, dst, tabs, a, s);
        }
        if (logn == 14 && split_enabled())
            return out_f  ? launch_inv_split<true, true>(src, base, base_group, base_stride, dst, n_polys, tabs, mod_base, mod_count, s)
                   : in_f ? launch_inv_split<true, false>(src, base, base_group, base_stride, dst, n_polys, tabs, mod_base, mod_count, s)
                          : launch_inv_split<false, false>(src, base, base_group, base_stride, dst, n_polys, tabs, mod_base, mod_count, s);
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
