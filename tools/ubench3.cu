// tools/ubench3.cu -- packed tap variants: which instruction mix sustains the highest f32 lane-op rate?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ unsigned long long pk(float2 v){ unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 up(unsigned long long r){ float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r)); return v; }
__device__ __forceinline__ float2 add2(float2 a, float2 b){ unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b){ unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c){ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(d); }
template <int OP>
__global__ void k(float *out, float a, float b, float one)
{
    float2 x[8], acc[8];
    const float2 A = make_float2(a, a), Bv = make_float2(b, b), O = make_float2(one, one);
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); acc[i] = x[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) acc[i] = fma2(mul2(add2(x[i], x[(i + 1) & 7]), A), O, acc[i]);        // FADD2, FMUL2, FFMA2(one)  (faithful, current kernel)
            if (OP == 1) acc[i] = fma2(add2(x[i], x[(i + 1) & 7]), A, acc[i]);                  // FADD2, FFMA2              (gauss_fma)
            if (OP == 2) { float2 s = add2(x[i], x[(i + 1) & 7]);                               // scalar mul + packed add: FADD2, 2xFMUL, FADD2
                           float2 p = make_float2(__fmul_rn(s.x, a), __fmul_rn(s.y, a)); acc[i] = add2(acc[i], p); }
            if (OP == 3) { float2 s = add2(x[i], x[(i + 1) & 7]); float2 p = mul2(s, A);        // FADD2, FMUL2, 2x scalar FADD
                           acc[i] = make_float2(__fadd_rn(acc[i].x, p.x), __fadd_rn(acc[i].y, p.y)); }
            if (OP == 4) { acc[i].x = acc[i].x + (x[i].x + x[(i + 1) & 7].x) * a; acc[i].y = acc[i].y + (x[i].y + x[(i + 1) & 7].y) * a; } // all scalar
        }
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = add2(x[i], Bv);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name)
{
    float *out;
    cudaMalloc(&out, 148 * 2 * 512 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<148 * 2, 512>>>(out, 1.0001f, 0.0f, 1.0f);
    cudaEventRecord(a);
    k<OP><<<148 * 2, 512>>>(out, 1.0001f, 0.0f, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double taps = 148.0 * 2 * 512 * ITERS * 8 * 2; // scalar-equivalent taps
    printf("%-44s %8.3f ms   %6.2f taps/clk/SM @1.9GHz\n", name, ms, taps / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main()
{
    run<0>("FADD2 FMUL2 FFMA2(one)  [faithful, current]");
    run<1>("FADD2 FFMA2             [gauss_fma]");
    run<2>("FADD2 2xFMUL FADD2      [faithful]");
    run<3>("FADD2 FMUL2 2xFADD      [faithful]");
    run<4>("scalar FADD FMUL FADD x2 [faithful]");
    return 0;
}
