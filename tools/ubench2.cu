// tools/ubench2.cu -- packed f32x2 (sm_100 FADD2/FMUL2/FFMA2) throughput vs scalar, for the window-blur design.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP>
__global__ void k(float *out, float a, float b)
{
    float2 x[8];
    const float2 A = make_float2(a, a * 1.0001f), Bv = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) { x[i].x = x[i].x + a; x[i].y = x[i].y + b; }            // 2 scalar FADD
            if (OP == 1) x[i] = __fadd2_rn(x[i], A);                               // 1 FADD2
            if (OP == 2) x[i] = __fmul2_rn(x[i], A);                               // 1 FMUL2
            if (OP == 3) x[i] = __ffma2_rn(x[i], A, Bv);                           // 1 FFMA2
            if (OP == 4) x[i] = __fadd2_rn(x[i], __fmul2_rn(__fadd2_rn(x[(i + 1) & 7], Bv), A)); // faithful tap, packed
            if (OP == 5) { x[i].x = x[i].x + (x[(i + 1) & 7].x + b) * a; x[i].y = x[i].y + (x[(i + 1) & 7].y + b) * a; } // scalar
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, double lane_ops)
{
    float *out;
    cudaMalloc(&out, 148 * 2 * 1024 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<148 * 2, 1024>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(a);
    k<OP><<<148 * 2, 1024>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double total = 148.0 * 2 * 1024 * ITERS * 8 * lane_ops;
    printf("%-34s %8.3f ms   %6.1f f32 lane-ops/clk/SM @1.9GHz\n", name, ms, total / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main()
{
    run<0>("2x scalar FADD", 2); run<1>("FADD2", 2); run<2>("FMUL2", 2); run<3>("FFMA2", 2);
    run<4>("packed add,mul,add (6 lane-ops)", 6); run<5>("scalar add,mul,add x2 (6 lane-ops)", 6);
    return 0;
}
