// tools/ubench4.cu -- how many warps per scheduler does the packed window-tap loop need to fill the FP32 pipe?
// The compute loop of gauss_v_walk2 (38-entry float2 register window, 8 outputs x 15 symmetric taps, FADD2 + FFMA2) on
// register data only, at 1 / 2 / 4 / 8 warps per SM sub-partition, one CTA per SM.  Prints packed lane-ops per clk per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench4 tools/ubench4.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float2 v){ unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 up(unsigned long long r){ float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r)); return v; }
__device__ __forceinline__ float2 add2(float2 a, float2 b){ unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b){ unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c){ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(d); }
struct Taps { float k[16]; };
constexpr int MR = 15, RV = 8, NIN = RV + 2 * MR, ITERS = 512;

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, Taps t, float seed)
{
    float2 win[NIN];
#pragma unroll
    for (int r = 0; r < NIN; r++) win[r] = make_float2(threadIdx.x * 0.001f + r * seed, threadIdx.x * 0.002f + r);
    float2 tot = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS; it++) {
        float2 v[RV];
        if (MODE == 0) { // packed symmetric: FADD2 + FFMA2 (the kernel's relaxed arithmetic)
#pragma unroll
            for (int o = 0; o < RV; o++) v[o] = mul2(win[o + MR], make_float2(t.k[0], t.k[0]));
#pragma unroll
            for (int i = 1; i <= MR; i++) {
                const float2 kk = make_float2(t.k[i], t.k[i]);
#pragma unroll
                for (int o = 0; o < RV; o++) v[o] = fma2(add2(win[o + MR + i], win[o + MR - i]), kk, v[o]);
            }
        } else if (MODE == 1) { // packed, no symmetry: 31 FFMA2 per output
#pragma unroll
            for (int o = 0; o < RV; o++) v[o] = mul2(win[o], make_float2(t.k[15], t.k[15]));
#pragma unroll
            for (int i = 1; i <= 2 * MR; i++) {
                const float kf = t.k[i <= MR ? MR - i : i - MR];
                const float2 kk = make_float2(kf, kf);
#pragma unroll
                for (int o = 0; o < RV; o++) v[o] = fma2(win[o + i], kk, v[o]);
            }
        } else { // scalar symmetric: FADD + FFMA on both halves
#pragma unroll
            for (int o = 0; o < RV; o++) v[o] = make_float2(win[o + MR].x * t.k[0], win[o + MR].y * t.k[0]);
#pragma unroll
            for (int i = 1; i <= MR; i++) {
#pragma unroll
                for (int o = 0; o < RV; o++) {
                    v[o].x = fmaf(win[o + MR + i].x + win[o + MR - i].x, t.k[i], v[o].x);
                    v[o].y = fmaf(win[o + MR + i].y + win[o + MR - i].y, t.k[i], v[o].y);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < RV; o++) tot = add2(tot, v[o]);
        // slide the window by 8 like the kernel does (register renaming only) and refresh the tail from the outputs
#pragma unroll
        for (int r = 0; r < NIN - RV; r++) win[r] = win[r + RV];
#pragma unroll
        for (int r = 0; r < RV; r++) win[NIN - RV + r] = v[r];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot.x + tot.y;
}

template <int MODE>
void run(const char *name, int threads)
{
    float *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    Taps t;
    for (int i = 0; i < 16; i++) t.k[i] = 0.03f / (1 + i);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, threads>>>(out, t, 0.5f);
    cudaEventRecord(a);
    k<MODE><<<148, threads>>>(out, t, 0.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // lane-ops: per output 1 + 2*15 = 31 (symmetric) or 31 (direct) float ops x 2 lanes
    double laneops = 148.0 * threads * ITERS * RV * 31.0 * 2.0;
    printf("%-34s warps/SMSP %d  %8.3f ms  %7.1f lane-ops/clk/SM @1.965GHz\n", name, threads / 128, ms, laneops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main()
{
    for (int th : {128, 256, 384, 512}) run<0>("packed FADD2+FFMA2 (kernel, relaxed)", th);
    for (int th : {128, 256, 384, 512}) run<1>("packed 31 x FFMA2 (no symmetry)", th);
    for (int th : {128, 256, 384, 512}) run<2>("scalar FADD+FFMA", th);
    return 0;
}
