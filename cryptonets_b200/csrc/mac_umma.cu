// Dense scalar-MAC layer on the 5th-generation tensor cores: tcgen05.mma kind::i8, accumulators in tensor memory (sm_100a).
//
// Same arithmetic as mac_imma.cu (the layer IS a matrix product over 8-bit limbs of the ciphertext words:
//     x = sum_a 2^(8a) x_a,  P_a[m][c] = sum_k W[m][k] x_a[k][c],  out = sum_a 2^(8a) P_a mod q_l,
// NeuralNetworks/PoolLayer.cs:196-227 for a window that covers the whole input), re-designed around what limited that kernel
// (profiles/r01_mac_layers_ncu.txt: 232 registers per thread for the 96 accumulators, one 8-warp CTA per SM, 12 % of HBM,
// long-scoreboard + barrier stalls -- latency bound, not tensor bound):
//   * accumulators live in TMEM (two 256-column buffers: the epilogue of one tile runs under the MMAs of the next), no thread
//     holds them;
//   * the ciphertext words arrive by TMA into a four-stage ring, one cp.async.bulk.tensor request per stage (a 2-D map over the
//     previous layer's output slab: 32 taps x 32 words = 8 KB): no registers, no address arithmetic in the consumers, as many bytes
//     in flight as HBM latency needs (taps that do not sit in the slab -- the W2 columns -- come one cp.async.bulk each);
//   * four "cutter" warps turn a raw stage into the B operand -- limb a of word n is row a*32 + n of a K-major, unswizzled
//     UMMA tile (8-row x 16-byte core matrices) -- with conflict-free 32-bit stores;
//   * one thread issues ONE tcgen05.mma (M = 128 outputs, N = 32 * limbs, K = 32 taps) per stage against the weight chunk that
//     has been resident in shared memory since the CTA started (A operand, packed by the host in core-matrix order), and
//     tcgen05.commit hands the B stage back;
//   * persistent CTAs (one per SM) walk the 32-word tiles of the ciphertext.
// Warp roles: 0-3 epilogue (TMEM lanes 32w..32w+31 = output rows), 4 bulk-copy producer, 5 MMA issuer, 6-9 cutters.
// Output words are bit-identical to k_mac_dense_imma / k_mac_layer_fp (tests/test_gpu_kernels.py::test_mac_layer_*).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h> // CUtensorMap
#include "fparith.cuh"
#include "kernels.h"
#include "plainops.cuh"

namespace cnhe {
namespace {

constexpr int UM_TN = 32;                            // ciphertext words per tile
constexpr int UM_M = 128;                            // MMA rows (outputs, zero padded)
constexpr int UM_CHUNK = 32;                         // taps per MMA (K of kind::i8)
constexpr int UM_RAW_STAGES = 4, UM_B_STAGES = 3;
constexpr int UM_RAW_ROW = 256;                      // bytes per tap row of a raw stage (dense: the tensor-map box lands this way)
constexpr int UM_RAW_BYTES = UM_CHUNK * UM_RAW_ROW;  // 8704
constexpr int UM_A_CHUNK = UM_M * UM_CHUNK;          // 4096 bytes of weights per chunk
constexpr int UM_OUT_ROW = 33;                       // words per row of the output staging tile (odd: conflict-free column writes)
constexpr int UM_THREADS = 320;
constexpr int UM_ACC_COLS = 256;                     // TMEM columns per accumulator buffer (two buffers = all 512)
constexpr int UM_CUTTERS = 128, UM_EPI = 128;

__device__ __forceinline__ unsigned sptr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(bar)) : "memory");
}
// bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mb_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = sptr(bar);
    unsigned done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
        if (!done && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sptr(dst)), "l"(src), "r"(bytes),
                 "r"(sptr(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_g2s_2d(void *dst, const void *tmap, int c0, int c1, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sptr(dst)), "l"(tmap),
                 "r"(c0), "r"(c1), "r"(sptr(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(bar)) : "memory");
}
// K-major operand without swizzle: 8-row x 16-byte core matrices; lbo = distance between the two 16-byte K halves, sbo = distance
// between 8-row groups (bytes); descriptor version 1 (sm_100)
__device__ __forceinline__ u64 umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (u64)((saddr & 0x3FFFFu) >> 4) | ((u64)(lbo >> 4) << 16) | ((u64)(sbo >> 4) << 32) | (1ULL << 46);
}
__device__ __forceinline__ void umma_i8(unsigned tmem_d, u64 adesc, u64 bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(UM_EPI) : "memory"); }

struct UmSmem { // offsets into the dynamic shared memory block (bytes)
    int w, raw, b, out, src, dst, mod, bars, total;
};
constexpr int UM_MAX_RES = 9; // coefficient moduli (KMAX)
__host__ __device__ inline UmSmem um_layout(int chunks, int limbs) {
    UmSmem s;
    s.w = 0;
    s.raw = s.w + chunks * UM_A_CHUNK;
    s.b = s.raw + UM_RAW_STAGES * UM_RAW_BYTES;
    s.out = s.b + UM_B_STAGES * limbs * UM_TN * UM_CHUNK;
    s.src = s.out + UM_M * UM_OUT_ROW * 8;          // source pointer of every tap (chunks x 32), read by the producer per stage
    s.dst = s.src + chunks * UM_CHUNK * 8;           // destination pointer of every output row
    s.mod = s.dst + UM_M * 8;                        // per residue: p, 1/p, 2^(8a) mod p (a = 3..6) as doubles
    s.bars = s.mod + UM_MAX_RES * 8 * 8;
    s.total = s.bars + 256;
    return s;
}

// wpack: chunks x 4096 bytes, weight (row r, tap kb of the chunk) at (r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15)
template <int LIMBS>
__global__ void __launch_bounds__(UM_THREADS, 1)
k_mac_dense_umma(const __grid_constant__ CUtensorMap tmap, const u64 *const *__restrict__ in_ptrs, const unsigned char *__restrict__ wpack,
                 const u64 *__restrict__ bias, int K, int n_extra, int M, u64 *const *__restrict__ out_ptrs, int k, int logn, const BehzConst *__restrict__ bc, PlainConst pc, unsigned long long *prof) {
    extern __shared__ __align__(128) unsigned char smem[];
    // CNHE_UMMA_PROF=1: CTA 0 reports, per role, the cycles spent in each of its waits and in its work (prof[role * 4 + i])
    const bool profiling = prof != nullptr && blockIdx.x == 0;
    long long t_a = 0, t_b = 0, t_c = 0, t0 = 0;
#define UM_T0() if (profiling) t0 = clock64()
#define UM_ACC(x) if (profiling) { const long long t1_ = clock64(); x += t1_ - t0; t0 = t1_; }
    constexpr int NB = LIMBS * UM_TN;            // MMA N: rows of the B operand / accumulator columns
    constexpr int B_BYTES = NB * UM_CHUNK;
    // taps 0..K-1 are rows of the tensor map (rows beyond K read as zero); the n_extra taps after them start a fresh chunk
    const int chunks_aff = (K + UM_CHUNK - 1) / UM_CHUNK;
    const int chunks = chunks_aff + (n_extra + UM_CHUNK - 1) / UM_CHUNK;
    const UmSmem L = um_layout(chunks, LIMBS);
    unsigned char *sw = smem + L.w, *sraw = smem + L.raw, *sb = smem + L.b;
    u64 *sout = reinterpret_cast<u64 *>(smem + L.out);
    const u64 **ssrc = reinterpret_cast<const u64 **>(smem + L.src);
    u64 **sdst = reinterpret_cast<u64 **>(smem + L.dst);
    double *smod = reinterpret_cast<double *>(smem + L.mod);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + L.bars);
    unsigned long long *raw_full = bars, *raw_empty = bars + 4, *b_full = bars + 8, *b_empty = bars + 11, *acc_full = bars + 14, *acc_empty = bars + 16,
                       *w_full = bars + 18;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + 20);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = 1 << logn;
    const int n_tiles = (int)(((size_t)2 * k << logn) / UM_TN);

    if (tid == 0) {
        for (int i = 0; i < UM_RAW_STAGES; i++) { mb_init(raw_full + i, 1); mb_init(raw_empty + i, UM_CUTTERS); }
        for (int i = 0; i < UM_B_STAGES; i++) { mb_init(b_full + i, UM_CUTTERS); mb_init(b_empty + i, 1); }
        for (int i = 0; i < 2; i++) { mb_init(acc_full + i, 1); mb_init(acc_empty + i, UM_EPI); }
        mb_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // pointer tables and modulus constants: fetched once, so that neither the producer (one dependent global load per stage cost it
    // ~0.6 us per chunk in the first version, ncu: long_scoreboard) nor the epilogue touches global memory for them again
    for (int i = tid; i < (chunks - chunks_aff) * UM_CHUNK; i += UM_THREADS) ssrc[i] = in_ptrs[K + min(i, n_extra - 1)]; // padding taps: zero weights
    for (int i = tid; i < UM_M; i += UM_THREADS) sdst[i] = i < M ? out_ptrs[i] : nullptr;
    if (tid < k) {
        const double p = (double)bc->q[tid].p, pinv = 1.0 / p;
        smod[tid * 8] = p;
        smod[tid * 8 + 1] = pinv;
        for (int a = 3; a < 8; a++) smod[tid * 8 + a - 1] = frecenter((double)(1ULL << (8 * a)), p, pinv); // slots 2..6 = a 3..7
    }
    if (warp == 0) { // the allocating warp also frees
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == 4) {
        // ---- producer: weights once, then 32 bulk copies (one per tap, 256 bytes = the tile's 32 words) per stage
        if (lane == 0) {
            mb_expect_tx(w_full, (unsigned)(chunks * UM_A_CHUNK));
            for (int c = 0; c < chunks; c++) bulk_g2s(sw + c * UM_A_CHUNK, wpack + (size_t)c * UM_A_CHUNK, UM_A_CHUNK, w_full);
        }
        unsigned it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const size_t col0 = (size_t)tile * UM_TN;
            for (int c = 0; c < chunks; c++, it++) {
                const unsigned s = it % UM_RAW_STAGES, ph = (it / UM_RAW_STAGES) & 1;
                UM_T0();
                mb_wait(raw_empty + s, ph ^ 1); // a fresh barrier passes the wait for the "previous" phase
                UM_ACC(t_a);
                if (lane == 0) mb_expect_tx(raw_full + s, UM_CHUNK * 256);
                __syncwarp();
                if (c < chunks_aff) { // one request: 32 taps x 32 words (issuing 32 separate copies cost ~60 cycles each: 94 % of the first version's time)
                    if (lane == 0) tma_g2s_2d(sraw + s * UM_RAW_BYTES, &tmap, (int)col0, c * UM_CHUNK, raw_full + s);
                } else
                    bulk_g2s(sraw + s * UM_RAW_BYTES + lane * UM_RAW_ROW, ssrc[(c - chunks_aff) * UM_CHUNK + lane] + col0, 256, raw_full + s);
                UM_ACC(t_b);
            }
        }
        if (profiling && lane == 0) { prof[0] = t_a; prof[1] = t_b; }
    } else if (warp == 5) {
        // ---- MMA issuer: one thread
        if (lane == 0) {
            // instruction descriptor: D = s32 (2 << 4), A = signed 8 bit (1 << 7), B = unsigned 8 bit (0 << 10), both K-major,
            // N >> 3 at bit 17, M >> 4 at bit 24
            const unsigned idesc = (2u << 4) | (1u << 7) | ((unsigned)(NB >> 3) << 17) | ((unsigned)(UM_M >> 4) << 24);
            mb_wait(w_full, 0);
            unsigned it = 0, ti = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ti++) {
                const unsigned as = ti & 1;
                UM_T0();
                mb_wait(acc_empty + as, ((ti >> 1) & 1) ^ 1);
                UM_ACC(t_a);
                tc_fence_after();
                for (int c = 0; c < chunks; c++, it++) {
                    const unsigned s = it % UM_B_STAGES, ph = (it / UM_B_STAGES) & 1;
                    UM_T0();
                    mb_wait(b_full + s, ph);
                    UM_ACC(t_b);
                    tc_fence_after();
                    umma_i8(tmem_base + as * UM_ACC_COLS, umma_desc(sptr(sw + c * UM_A_CHUNK), 128, 256), umma_desc(sptr(sb + s * B_BYTES), 128, 256), idesc,
                            c > 0);
                    tc_commit(b_empty + s); // the stage is free once this MMA (and everything before it) has read it
                    UM_ACC(t_c);
                }
                tc_commit(acc_full + as);
            }
            if (profiling) { prof[4] = t_a; prof[5] = t_b; prof[6] = t_c; }
        }
        __syncwarp();
    } else if (warp >= 6) {
        // ---- cutters: raw words -> limb bytes in UMMA order.  Thread (q = tap quad 0..7, n = word 0..31) packs limb a of taps 4q..4q+3
        // into one 32-bit store at row a*32+n, bytes 4q..4q+3.  Lane = (q & 3) + 4 * (n & 7): the 32 stores of a warp fill one 8-row core
        // matrix half (512 contiguous bytes, conflict free); the 64-bit loads touch 8 consecutive words of 4 tap rows (2 wavefronts)
        const int ct = tid - 6 * 32; // 0..127
        const int cw = ct >> 5;      // cutter warp: n >> 3
        unsigned it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int c = 0; c < chunks; c++, it++) {
                const unsigned rs = it % UM_RAW_STAGES, rph = (it / UM_RAW_STAGES) & 1;
                const unsigned bs = it % UM_B_STAGES, bph = (it / UM_B_STAGES) & 1;
                UM_T0();
                mb_wait(raw_full + rs, rph);
                UM_ACC(t_a);
                mb_wait(b_empty + bs, bph ^ 1);
                UM_ACC(t_b);
                const unsigned char *raw = sraw + rs * UM_RAW_BYTES;
                unsigned char *bst = sb + bs * B_BYTES;
                const int n = cw * 8 + (lane >> 2);
#pragma unroll
                for (int qh = 0; qh < 2; qh++) {
                    const int q = qh * 4 + (lane & 3);
                    unsigned lo[4], hi[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint2 t = *reinterpret_cast<const uint2 *>(raw + (4 * q + j) * UM_RAW_ROW + n * 8);
                        lo[j] = t.x;
                        hi[j] = t.y;
                    }
#pragma unroll
                    for (int a = 0; a < LIMBS; a++) { // byte a of the four words -> one 32-bit word, three byte permutes
                        const unsigned sel = (a & 3) | (((a & 3) + 4) << 4);
                        const unsigned t01 = __byte_perm(a < 4 ? lo[0] : hi[0], a < 4 ? lo[1] : hi[1], sel);
                        const unsigned t23 = __byte_perm(a < 4 ? lo[2] : hi[2], a < 4 ? lo[3] : hi[3], sel);
                        const unsigned w = __byte_perm(t01, t23, 0x5410);
                        const int row = a * UM_TN + n;
                        *reinterpret_cast<unsigned *>(bst + (row >> 3) * 256 + qh * 128 + (row & 7) * 16 + (lane & 3) * 4) = w;
                    }
                }
                fence_async_smem(); // generic-proxy stores -> visible to the tensor core's (async proxy) reads
                mb_arrive(b_full + bs);
                mb_arrive(raw_empty + rs);
                UM_ACC(t_c);
            }
        }
        if (profiling && ct == 0) { prof[8] = t_a; prof[9] = t_b; prof[10] = t_c; }
    } else {
        // ---- epilogue (warps 0-3): thread = output row m.  out = sum_a 2^(8a) P_a mod q_l, exact in FP64 for p < 2^50: the three low
        // limbs combine below 2^48 without reduction, every higher limb is a modular product with 2^(8a) mod p
        const int m = tid;
        unsigned ti = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ti++) {
            const unsigned as = ti & 1;
            const size_t col0 = (size_t)tile * UM_TN;
            const int l = (int)((col0 >> logn) % k);
            const double p = smod[l * 8], pinv = smod[l * 8 + 1];
            double cpow[LIMBS];
#pragma unroll
            for (int a = 3; a < LIMBS; a++) cpow[a] = smod[l * 8 + a - 1];
            UM_T0();
            mb_wait(acc_full + as, (ti >> 1) & 1);
            UM_ACC(t_a);
            tc_fence_after();
            const unsigned tbase = tmem_base + ((unsigned)(warp * 32) << 16) + as * UM_ACC_COLS;
#pragma unroll
            for (int n0 = 0; n0 < UM_TN; n0 += 8) {
                int acc[LIMBS][8];
#pragma unroll
                for (int a = 0; a < LIMBS; a++) tmem_ld8(tbase + a * UM_TN + n0, acc[a]);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    double rr = __fma_rn((double)acc[2][e], 65536.0, __fma_rn((double)acc[1][e], 256.0, (double)acc[0][e]));
#pragma unroll
                    for (int a = 3; a < LIMBS; a++) rr = __dadd_rn(rr, fmodmul((double)acc[a][e], cpow[a], p, pinv));
                    sout[m * UM_OUT_ROW + n0 + e] = fcanon_u(rr, p, pinv);
                }
            }
            tc_fence_before();
            mb_arrive(acc_empty + as); // the accumulator buffer may be overwritten
            if (bias && m < M && col0 < (size_t)k * N && (col0 & (size_t)(N - 1)) == 0) { // constant-plaintext bias: Delta*b on coefficient 0 of c0
                const u64 b = bias[m];
                const DMod q = bc->q[l];
                if (b) sout[m * UM_OUT_ROW] = addmod(sout[m * UM_OUT_ROW], scale_plain(b, l, q, pc), q.p);
            }
            UM_ACC(t_b);
            epi_sync();
            for (int r = warp; r < M; r += 4) sdst[r][col0 + lane] = sout[r * UM_OUT_ROW + lane]; // 256 contiguous bytes per row
            epi_sync();
            UM_ACC(t_c);
        }
        if (profiling && tid == 0) { prof[12] = t_a; prof[13] = t_b; prof[14] = t_c; }
    }
#undef UM_T0
#undef UM_ACC
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int sm_count_cached() {
    static int n = [] {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

template <int LIMBS>
cudaError_t umma_go(const CUtensorMap &map, const u64 *const *in_ptrs, const unsigned char *wpack, const u64 *bias, int K, int n_extra, int M,
                    u64 *const *out_ptrs, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s) {
    const int chunks = (K + UM_CHUNK - 1) / UM_CHUNK + (n_extra + UM_CHUNK - 1) / UM_CHUNK;
    const UmSmem L = um_layout(chunks, LIMBS);
    cudaError_t e = cudaFuncSetAttribute(k_mac_dense_umma<LIMBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return e;
    const int n_tiles = (int)(((size_t)2 * k << logn) / UM_TN);
    const int grid = std::min(sm_count_cached(), n_tiles);
    const bool want_prof = getenv("CNHE_UMMA_PROF") != nullptr; // read per launch: tests switch it on to see which kernel served a layer
    unsigned long long *prof = nullptr;
    if (want_prof) {
        static unsigned long long *buf = nullptr;
        if (!buf) cudaMalloc((void **)&buf, 16 * sizeof(unsigned long long));
        cudaMemsetAsync(buf, 0, 16 * sizeof(unsigned long long), s);
        prof = buf;
    }
    k_mac_dense_umma<LIMBS><<<grid, UM_THREADS, L.total, s>>>(map, in_ptrs, wpack, bias, K, n_extra, M, out_ptrs, k, logn, bc, pc, prof);
    if (want_prof) {
        unsigned long long h[16];
        cudaMemcpyAsync(h, prof, sizeof(h), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        fprintf(stderr, "[umma K=%d+%d M=%d tiles/cta=%.1f] producer: wait_empty %llu issue %llu | mma: wait_acc %llu wait_b %llu issue %llu | cutters: wait_raw %llu wait_b %llu work %llu | "
                        "epilogue: wait_acc %llu compute %llu store %llu (cycles, CTA 0)\n",
                K, n_extra, M, (double)n_tiles / grid, h[0], h[1], h[4], h[5], h[6], h[8], h[9], h[10], h[12], h[13], h[14]);
    }
    return cudaGetLastError();
}

} // namespace

// does the tcgen05 kernel take this layer?  (one weight byte per tap: |w| <= 127; the weight matrix has to fit in shared memory next to
// the rings: K <= 1088 taps for 6 limbs)
bool mac_dense_umma_fits(int K, int M, int limbs) {
    if (M < 1 || M > UM_M || K < 1 || limbs < 5 || limbs > 7) return false;
    const int chunks = (K + UM_CHUNK - 1) / UM_CHUNK;
    return um_layout(chunks, limbs).total <= 227 * 1024;
}
size_t mac_dense_umma_weight_bytes(int K) { return (size_t)((K + UM_CHUNK - 1) / UM_CHUNK) * UM_A_CHUNK; }
// host-side packing of the signed 8-bit weight matrix (w[m * K + kk], zero where a tap is padded) into the A-operand order
void mac_dense_umma_pack(const signed char *w, int M, int K, unsigned char *out) {
    const int chunks = (K + UM_CHUNK - 1) / UM_CHUNK;
    memset(out, 0, (size_t)chunks * UM_A_CHUNK);
    for (int r = 0; r < M; r++)
        for (int kk = 0; kk < K; kk++) {
            const int c = kk / UM_CHUNK, kb = kk % UM_CHUNK;
            out[(size_t)c * UM_A_CHUNK + (r >> 3) * 256 + (kb >> 4) * 128 + (r & 7) * 16 + (kb & 15)] = (unsigned char)w[(size_t)r * K + kk];
        }
}
cudaError_t launch_mac_dense_umma(const u64 *const *in_ptrs, const u64 *affine_base, size_t affine_stride_words, int K, int n_extra, const void *wpack,
                                  const u64 *bias, int M, int limbs, u64 *const *out_ptrs, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s) {
    const unsigned char *wp = reinterpret_cast<const unsigned char *>(wpack);
    alignas(64) CUtensorMap map;
    cudaError_t e = make_word_map_2d(&map, affine_base, (size_t)2 * k << logn, (size_t)K, affine_stride_words * 8, UM_TN, UM_CHUNK);
    if (e != cudaSuccess) return e;
    switch (limbs) {
    case 5: return umma_go<5>(map, in_ptrs, wp, bias, K, n_extra, M, out_ptrs, k, logn, bc, pc, s);
    case 6: return umma_go<6>(map, in_ptrs, wp, bias, K, n_extra, M, out_ptrs, k, logn, bc, pc, s);
    case 7: return umma_go<7>(map, in_ptrs, wp, bias, K, n_extra, M, out_ptrs, k, logn, bc, pc, s);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cnhe
