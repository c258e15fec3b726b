// Synthetic study of the FP64 butterfly pass structure: how close to the DFMA issue peak can a shared-memory resident
// radix-8 pass (LDS -> 3 stages of butterflies -> STS -> barrier) get, as a function of threads per CTA and CTAs per SM?
#include <cstdio>
#include <cuda_runtime.h>
constexpr double MAGIC = 6755399441055744.0;
__device__ __forceinline__ double fmodmul(double a, double w, double p, double pinv) {
    const double h = __dmul_rn(a, w), l = __fma_rn(a, w, -h);
    const double q = __dsub_rn(__fma_rn(h, pinv, MAGIC), MAGIC);
    return __dadd_rn(__fma_rn(-q, p, h), l);
}
__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) & 7) << 1); }
template <int MODE> // 0: full pass (smem + barrier), 1: no barrier, 2: registers only (no smem), 3: latency chain
__global__ void k(double *out, int iters, double p, double pinv, int n) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, T = blockDim.x;
    for (int i = tid; i < n; i += T) sm[i] = (double)((i * 2654435761u) & 0xfffff);
    __syncthreads();
    double acc = 0;
    const int groups = n / 16; // each "virtual thread" handles 16 elements (two column-adjacent radix-8 groups)
    if (MODE == 3) {
        double x = sm[tid];
        for (int it = 0; it < iters * 64; it++) x = __fma_rn(x, 1.0000001, 0.5);
        out[blockIdx.x * T + tid] = x;
        return;
    }
    double x[8], y[8];
    if (MODE == 2) for (int e = 0; e < 8; e++) { x[e] = sm[(tid * 16 + e) % n]; y[e] = sm[(tid * 16 + 8 + e) % n]; }
    for (int it = 0; it < iters; it++) {
        for (int vt = tid; vt < groups; vt += T) {
            const int LG = 7;
            const int c2 = vt & ((1 << (LG - 1)) - 1), j = vt >> (LG - 1);
            const int base = (j << (LG + 3)) + 2 * c2;
            if (MODE != 2) {
#pragma unroll
                for (int e = 0; e < 8; e++) { double2 v = *reinterpret_cast<double2 *>(sm + swz((base + (e << LG)) % n)); x[e] = v.x; y[e] = v.y; }
            }
#pragma unroll
            for (int u = 0; u < 3; u++) {
                const int h = 4 >> u;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    if (e & h) continue;
                    const double w = 12345.0 + (double)(u + (e >> (3 - u)));
                    const double t0 = fmodmul(x[e + h], w, p, pinv), t1 = fmodmul(y[e + h], w, p, pinv);
                    const double a0 = x[e], a1 = y[e];
                    x[e] = a0 + t0; x[e + h] = a0 - t0; y[e] = a1 + t1; y[e + h] = a1 - t1;
                }
            }
            if (MODE != 2) {
#pragma unroll
                for (int e = 0; e < 8; e++) *reinterpret_cast<double2 *>(sm + swz((base + (e << LG)) % n)) = make_double2(x[e] * 1e-3, y[e] * 1e-3);
            }
        }
        if (MODE == 0) __syncthreads();
    }
    for (int e = 0; e < 8; e++) acc += x[e] + y[e];
    out[blockIdx.x * T + tid] = acc + sm[tid];
}
template <int MODE> void run(const char *name, int threads, int ctas_per_sm, int n, int iters) {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int blocks = pr.multiProcessorCount * ctas_per_sm;
    double *out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    size_t smem = (size_t)n * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads, smem>>>(out, 2, 2.0e13, 1 / 2.0e13, n); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, threads, smem>>>(out, iters, 2.0e13, 1 / 2.0e13, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    if (MODE == 3) { printf("%-34s dependent DFMA latency ~ %.1f cycles\n", name, cyc / (iters * 64.0)); cudaFree(out); return; }
    // DP instructions per pass per CTA: (n/16) virtual threads * 12 butterfly pairs... = n/2*3 butterflies * 8 DP + 16 mults(store scale)
    double dp_warp_instr = (double)iters * ((n / 2) * 3 * 8 + n) / 32.0 * ctas_per_sm; // per SM
    double util = dp_warp_instr / (cyc * 4 * 0.5);
    printf("%-34s thr=%4d ctas/SM=%d n=%5d : %8.0f cycles/pass/SM-set, DP pipe utilisation %.2f\n", name, threads, ctas_per_sm, n, cyc / iters, util);
    cudaFree(out);
}
int main() {
    run<3>("latency", 32, 1, 8192, 200);
    for (int thr : {256, 512, 1024}) for (int c : {1, 2, 3}) { if (thr * c > 2048 || c * 8192 * 8 > 220000) continue; run<0>("full pass (smem+barrier)", thr, c, 8192, 200); }
    run<1>("no barrier", 512, 2, 8192, 200); run<1>("no barrier", 256, 3, 8192, 200);
    run<2>("registers only", 512, 2, 8192, 200); run<2>("registers only", 256, 3, 8192, 200); run<2>("registers only", 1024, 2, 8192, 200);
    run<0>("full pass n=4096", 256, 3, 4096, 200); run<0>("full pass n=4096", 256, 6, 4096, 200); run<0>("full pass n=4096 128thr", 128, 6, 4096, 200);
    return 0;
}
