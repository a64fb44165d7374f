// EXPERIMENTAL -- not part of libtidalwave_b200.so, never launched by the product, not validated on a GPU yet.
// Register / shared-memory feasibility probe for the window-kernel rewrite planned in DESIGN.md (section 8): the vertical
// pass as a ROW STREAM.  A thread owns one column of one float2 channel pair; every input row it loads feeds the <= 31
// output rows whose window contains it, each through one FFMA2 into a rotating accumulator (31 live accumulators = 62
// registers instead of a 38-row window + outputs + refill = 108).  Tap order per output row = top to bottom (oracle relax
// bit 8, oracle/farneback_ref.c).  Compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -v -c gauss_stream_vpass.cu
// and read the register count: the plan needs <= 80 (three CTAs of 256 threads per SM).
#include <cuda_runtime.h>

namespace twx {

constexpr int MR = 15, TH = 32, NIN = TH + 2 * MR, NACC = 2 * MR + 1, PF = 4; // PF = rows requested ahead
struct Taps { float k[MR + 1]; };

__device__ __forceinline__ unsigned long long pk(float2 v) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ float2 up(unsigned long long r) { float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r)); return v; }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return up(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return up(d); }

// src: first input row (y0 - MR) of this thread's column, rstride in float2; dst: shared-memory column, dstride in float2
__device__ __forceinline__ void v_stream(const float2 *__restrict__ src, int rstride, float2 *__restrict__ dst, int dstride, const Taps &t)
{
    float2 acc[NACC];
    float2 q[PF];
#pragma unroll
    for (int r = 0; r < PF; r++) q[r] = __ldg(src + (size_t)r * rstride);
#pragma unroll
    for (int r = 0; r < NIN; r++) {
        const float2 in = q[r % PF];
        if (r + PF < NIN) q[r % PF] = __ldg(src + (size_t)(r + PF) * rstride);
#pragma unroll
        for (int o = 0; o < TH; o++) { // output row o covers input rows o .. o + 2 * MR
            const int d = r - o - MR;  // tap index of this row for output o (compile-time after unrolling)
            if (d < -MR || d > MR) continue;
            const float kk = t.k[d < 0 ? -d : d];
            if (d == -MR) acc[o % NACC] = mul2(in, make_float2(kk, kk));
            else acc[o % NACC] = fma2(in, make_float2(kk, kk), acc[o % NACC]);
            if (d == MR) dst[o * dstride] = acc[o % NACC];
        }
    }
}

__global__ void __launch_bounds__(256, 3) v_stream_probe(const float2 *__restrict__ M, float2 *__restrict__ out, int pitch2, Taps t)
{
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x, pair = tid >> 7, j = tid & 127;
    const float2 *src = M + (size_t)blockIdx.y * TH * 5 * pitch2 / 2 + pair * pitch2 + blockIdx.x * 96 + j;
    v_stream(src, 5 * pitch2 / 2, sm + pair * TH * 130 + j, 130, t);
    __syncthreads();
    for (int i = tid; i < 2 * TH * 130; i += 256) out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 2 * TH * 130 + i] = sm[i];
}

} // namespace twx
