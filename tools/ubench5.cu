// tools/ubench5.cu -- vertical pass of the window kernel IN ISOLATION, with its real global loads and shared-memory stores:
//   A  "window" form of the shipped kernel (gauss_v_walk2, direct-form taps): 38-row register window, 8 outputs per group,
//      128 registers -> 2 CTAs / SM (dynamic shared memory sized like the shipped kernel's, 108 KB)
//   B  "stream" form planned in DESIGN.md section 8 (tidal-wave_b200/csrc/experimental/gauss_stream_vpass.cu): 31 rotating
//      accumulators, 80 registers -> 3 CTAs / SM (66.5 KB)
// on the M layout of 16 pairs at 1920x1080 (row-interleaved planes, pitch 2048).  Prints ms per launch and FP32 lane-ops
// per clock per SM.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench5 tools/ubench5.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int MR = 15, TH = 32, NIN = TH + 2 * MR, NACC = 2 * MR + 1, RV = 8, SP = 130;
struct Taps { float k[MR + 1]; };

__device__ __forceinline__ unsigned long long pk(float2 v) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 up(unsigned long long r) { float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r)); return v; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(d); }

template <int PF>
__device__ __forceinline__ void v_stream(const float2 *__restrict__ src, int rstride, float2 *__restrict__ dst, const Taps &t)
{
    float2 acc[NACC], q[PF];
#pragma unroll
    for (int r = 0; r < PF; r++) q[r] = __ldg(src + (size_t)r * rstride);
#pragma unroll
    for (int r = 0; r < NIN; r++) {
        const float2 in = q[r % PF];
        if (r + PF < NIN) q[r % PF] = __ldg(src + (size_t)(r + PF) * rstride);
#pragma unroll
        for (int o = 0; o < TH; o++) {
            const int d = r - o - MR;
            if (d < -MR || d > MR) continue;
            const float kk = t.k[d < 0 ? -d : d];
            if (d == -MR) acc[o % NACC] = mul2(in, make_float2(kk, kk));
            else acc[o % NACC] = fma2(in, make_float2(kk, kk), acc[o % NACC]);
            if (d == MR) dst[o * SP] = acc[o % NACC];
        }
    }
}

__device__ __forceinline__ void v_window(const float2 *__restrict__ src, int rstride, float2 *__restrict__ dst, const Taps &t)
{
    constexpr int NW = RV + 2 * MR, NG = TH / RV;
    float2 win[NW];
#pragma unroll
    for (int r = 0; r < NW; r++) win[r] = __ldg(src + (size_t)r * rstride);
#pragma unroll
    for (int g = 0; g < NG; g++) {
        float2 nxt[RV];
        if (g + 1 < NG) {
#pragma unroll
            for (int r = 0; r < RV; r++) nxt[r] = __ldg(src + (size_t)(NW + RV * g + r) * rstride);
        }
        float2 v[RV];
#pragma unroll
        for (int o = 0; o < RV; o++) v[o] = mul2(win[o + MR], make_float2(t.k[0], t.k[0]));
#pragma unroll
        for (int i = 1; i <= MR; i++) {
            const float2 kk = make_float2(t.k[i], t.k[i]);
#pragma unroll
            for (int o = 0; o < RV; o++) v[o] = fma2(win[o + MR - i], kk, v[o]);
#pragma unroll
            for (int o = 0; o < RV; o++) v[o] = fma2(win[o + MR + i], kk, v[o]);
        }
#pragma unroll
        for (int o = 0; o < RV; o++) dst[(RV * g + o) * SP] = v[o];
        if (g + 1 < NG) {
#pragma unroll
            for (int r = 0; r < NW - RV; r++) win[r] = win[r + RV];
#pragma unroll
            for (int r = 0; r < RV; r++) win[NW - RV + r] = nxt[r];
        }
    }
}

template <int MODE>
__device__ __forceinline__ void body(const float *__restrict__ M, float *__restrict__ out, int pitch, size_t plane5, const Taps &t)
{
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x, pair = tid >> 7, j = tid & 127;
    const float *Mb = M + (size_t)blockIdx.z * plane5;
    const float2 *src = reinterpret_cast<const float2 *>(Mb + (size_t)blockIdx.y * TH * 5 * pitch + pair * 2 * pitch) + blockIdx.x * 96 + j;
    float2 *dst = sm + pair * TH * SP + j;
    if (MODE == 0) v_window(src, 5 * pitch / 2, dst, t);
    else if (MODE == 1) v_stream<4>(src, 5 * pitch / 2, dst, t);
    else v_stream<8>(src, 5 * pitch / 2, dst, t);
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 v = sm[(tid + 256 * i) % (2 * TH * SP)]; s += v.x + v.y; }
    out[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 256 + tid] = s;
}

__global__ void __launch_bounds__(256, 2) k_window(const float *M, float *out, int pitch, size_t plane5, Taps t) { body<0>(M, out, pitch, plane5, t); }
__global__ void __launch_bounds__(256, 3) k_stream4(const float *M, float *out, int pitch, size_t plane5, Taps t) { body<1>(M, out, pitch, plane5, t); }
__global__ void __launch_bounds__(256, 3) k_stream8(const float *M, float *out, int pitch, size_t plane5, Taps t) { body<2>(M, out, pitch, plane5, t); }
__global__ void fill(float *p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (float)((i * 2654435761u) & 1023) * 1e-3f; }

template <class K>
void run(const char *name, K kern, size_t smem, const float *M, float *out, int pitch, size_t plane5, Taps t, dim3 grid)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { printf("%s: attr %s\n", name, cudaGetErrorString(e)); return; }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<grid, 256, smem>>>(M, out, pitch, plane5, t);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    float best = 1e9f;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        kern<<<grid, 256, smem>>>(M, out, pitch, plane5, t);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double laneops = (double)grid.x * grid.y * grid.z * 256 * TH * 31.0 * 2.0;
    const double bytes = (double)grid.x * grid.y * grid.z * 256 * NIN * 8.0;
    printf("%-28s CTAs/SM %d  %8.3f ms  %6.1f lane-ops/clk/SM @1.965GHz  loads %.2f TB/s (L2->SM)\n", name, occ, best,
           laneops / (best * 1e-3) / 148 / 1.965e9, bytes / (best * 1e-3) / 1e12);
}

int main()
{
    const int W = 1920, H = 1080, B = 16, pitch = 2048;
    (void)W;
    const size_t plane5 = (size_t)H * 5 * pitch, n = plane5 * B;
    float *M, *out;
    if (cudaMalloc(&M, n * sizeof(float)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    dim3 grid(20, 32, B); // 32 tile rows: input rows 0 .. 32*32 + 61 < 1080
    cudaMalloc(&out, (size_t)grid.x * grid.y * grid.z * 256 * sizeof(float));
    fill<<<148 * 8, 256>>>(M, n);
    cudaDeviceSynchronize();
    Taps t;
    for (int i = 0; i <= MR; i++) t.k[i] = 0.03f / (1 + i);
    const size_t sm_v = sizeof(float2) * 2 * TH * SP;
    run("window (shipped), 108 KB", k_window, 108288, M, out, pitch, plane5, t, grid);
    run("window, 66.5 KB", k_window, sm_v, M, out, pitch, plane5, t, grid);
    run("stream PF=4, 66.5 KB", k_stream4, sm_v, M, out, pitch, plane5, t, grid);
    run("stream PF=8, 66.5 KB", k_stream8, sm_v, M, out, pitch, plane5, t, grid);
    run("stream PF=4, 108 KB (2/SM)", k_stream4, 108288, M, out, pitch, plane5, t, grid);
    // the same passes on 2 pairs only: the 71 MB they read stay L2-resident between the timed launches (DRAM out of the picture)
    dim3 g2(20, 32, 2);
    run("L2-resident: window, 108 KB", k_window, 108288, M, out, pitch, plane5, t, g2);
    run("L2-resident: window, 66.5 KB", k_window, sm_v, M, out, pitch, plane5, t, g2);
    run("L2-resident: stream PF=4", k_stream4, sm_v, M, out, pitch, plane5, t, g2);
    return 0;
}
