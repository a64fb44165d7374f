// tools/ubench.cu -- instruction-throughput microbenchmarks that size the kernels (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
template <int OP>
__global__ void k(float *out, float a, float b, double da, double db)
{
    float x[8];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 0.001f + i; d[i] = threadIdx.x * 0.001 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) x[i] = x[i] + a;                       // FADD
            if (OP == 1) x[i] = x[i] * a;                       // FMUL
            if (OP == 2) x[i] = fmaf(x[i], a, b);               // FFMA
            if (OP == 3) x[i] = x[i] + (x[(i + 1) & 7] + b) * a; // add,mul,add (the faithful window tap)
            if (OP == 4) d[i] = d[i] + da;                       // DADD
            if (OP == 5) d[i] = d[i] * da;                       // DMUL
            if (OP == 6) d[i] = fma(d[i], da, db);               // DFMA
            if (OP == 7) { d[i] = d[i] + (double)x[i]; x[i] = x[i] + a; } // F2F.F64.F32 + DADD + FADD
            if (OP == 8) { x[i] = (float)d[i]; d[i] = d[i] + (double)x[i]; } // F2F both ways + DADD
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lds_k(float *out, int pitch)
{
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int base = (threadIdx.x & 31) * pitch;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float4 v = *reinterpret_cast<float4 *>(&sm[(base + i * 4 + (it & 7) * 32) & 8188]);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int OP>
void run(const char *name, double ops_per_iter)
{
    float *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = 148 * 2, threads = 1024;
    k<OP><<<blocks, threads>>>(out, 1.0001f, 0.5f, 1.0001, 0.5);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(out, 1.0001f, 0.5f, 1.0001, 0.5);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double total = (double)blocks * threads * ITERS * 8 * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Gop/s/SM-equivalent lanes/clk/SM @1.9GHz: %6.1f\n", name, ms, total / ms / 1e6 / 148,
           total / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main()
{
    run<0>("FADD", 1); run<1>("FMUL", 1); run<2>("FFMA", 1); run<3>("add,mul,add (3 ops)", 3);
    run<4>("DADD", 1); run<5>("DMUL", 1); run<6>("DFMA", 1); run<7>("F2F.64.32+DADD+FADD (3)", 3); run<8>("F2F x2 + DADD (3)", 3);
    float *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    for (int pitch : {132, 128, 36}) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        lds_k<<<148 * 2, 1024, 32768>>>(out, pitch);
        cudaEventRecord(a);
        lds_k<<<148 * 2, 1024, 32768>>>(out, pitch);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        double bytes = 148.0 * 2 * 1024 * ITERS * 8 * 16;
        printf("LDS.128 pitch %3d: %8.3f ms  %7.1f B/clk/SM @1.9GHz\n", pitch, ms, bytes / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
