// tw_kernels.cu -- hand-written sm_100a kernels of the Farneback hot path.
//
// Compiled with -fmad=false: the oracle's results depend on the rounding of every add / mul
// ("hard part 1"), so nothing may be contracted implicitly.  Fused multiply-adds
// appear only where the oracle itself uses them (pre-blur, App. A.2a) and are written as fmaf().
//
// The algorithm restated here is OpenCV's calcOpticalFlowFarneback -- the single library call at
// /root/reference/src/opticalflow.cpp:83-85 -- and the sampling loop of /root/reference/src/consumer.cpp:60-77.
#include "tw_kernels.cuh"

namespace tw {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ------------------------------------------------------------------------------------------------
// K1  level image: u8 -> float, GaussianBlur(ksize) REFLECT_101 (rows first, then columns), bilinear
//     resize to (w, h).  One block = one output tile; the u8 source tile (with halo) is staged in shared
//     memory, the row pass is evaluated only at the source columns the resize reads, the column pass
//     only at the source rows it reads.   SURVEY App. A.2 / A.2a / A.2b.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) level_image_kernel(LevelImageArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = a.ksize / 2;
    const int d0 = blockIdx.x * a.tile_w, e0 = blockIdx.y * a.tile_h;
    const int d1 = min(d0 + a.tile_w, a.d.w) - 1, e1 = min(e0 + a.tile_h, a.d.h) - 1;
    const int tw_ = d1 - d0 + 1, th_ = e1 - e0 + 1;
    const int W = a.W, H = a.H;
    // source extents
    const int xs0 = a.xi[d0], xs1 = min(a.xi[d1] + 1, W - 1);
    const int ys0 = a.yi[e0], ys1 = min(a.yi[e1] + 1, H - 1);
    const int cx_lo = max(0, xs0 - c), cx_hi = min(W - 1, xs1 + c);
    const int ry_lo = max(0, ys0 - c), ry_hi = min(H - 1, ys1 + c);
    const int SW = cx_hi - cx_lo + 1, SH = ry_hi - ry_lo + 1;
    const int SWp = (a.smem_w + 3) & ~3;
    const int NC = a.identity ? a.tile_w : a.tile_w * 2;
    unsigned char *stile = smem_raw;                                            // [smem_h][SWp]
    float *rp = reinterpret_cast<float *>(smem_raw + (((size_t)a.smem_h * SWp + 15) & ~(size_t)15)); // [smem_h][NC]
    float *ktab = rp + (size_t)a.smem_h * NC;                                   // [ksize]

    const uint8_t *src = a.src + (size_t)blockIdx.z * H * a.spitch;
    for (int i = threadIdx.x; i < a.ksize; i += blockDim.x) ktab[i] = a.taps[i];
    for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) {
        int r = i / SW, q = i - r * SW;
        stile[r * SWp + q] = src[(size_t)(ry_lo + r) * a.spitch + cx_lo + q];
    }
    __syncthreads();

    // row pass at the needed columns
    const int ncol = a.identity ? tw_ : tw_ * 2;
    for (int i = threadIdx.x; i < SH * ncol; i += blockDim.x) {
        int r = i / ncol, slot = i - r * ncol;
        int xcol;
        if (a.identity) xcol = d0 + slot;
        else xcol = min(a.xi[d0 + (slot >> 1)] + (slot & 1), W - 1);
        const unsigned char *row = stile + r * SWp - cx_lo;
        float o;
        if (a.ksize == 3) {
            float L = (float)row[reflect101(xcol - 1, W)], C = (float)row[xcol], Rr = (float)row[reflect101(xcol + 1, W)];
            o = fmaf(C, ktab[1], (L + Rr) * ktab[0]);
        } else {
            o = (float)row[reflect101(xcol - c, W)] * ktab[0];
            for (int j = 1; j < a.ksize; j++) o = fmaf((float)row[reflect101(xcol - c + j, W)], ktab[j], o);
        }
        rp[r * NC + slot] = o;
    }
    __syncthreads();

    // column pass at the needed rows + bilinear
    float *dst = a.dst + (size_t)blockIdx.z * a.d.plane;
    for (int i = threadIdx.x; i < tw_ * th_; i += blockDim.x) {
        int ty = i / tw_, tx = i - ty * tw_;
        int d = d0 + tx, e = e0 + ty;
        int y0 = a.yi[e], y1 = min(y0 + 1, H - 1);
        float res[2][2];
        const int nyy = a.identity ? 1 : 2, nxx = a.identity ? 1 : 2;
        for (int yy = 0; yy < nyy; yy++) {
            int y = yy ? y1 : y0;
            for (int xx = 0; xx < nxx; xx++) {
                int slot = a.identity ? tx : tx * 2 + xx;
                const float *col = rp + slot - (size_t)ry_lo * NC; // col[y*NC]
                float o;
                if (a.ksize == 3) {
                    float U = col[reflect101(y - 1, H) * NC], C = col[y * NC], D = col[reflect101(y + 1, H) * NC];
                    o = fmaf(U + D, ktab[0], C * ktab[1]);
                } else {
                    o = col[y * NC] * ktab[c];
                    for (int j = 1; j <= c; j++)
                        o = fmaf(col[reflect101(y - j, H) * NC] + col[reflect101(y + j, H) * NC], ktab[c + j], o);
                }
                res[yy][xx] = o;
            }
        }
        float out;
        if (a.identity) {
            out = res[0][0];
        } else {
            float fx = a.xf[d], gx = 1.f - fx, fy = a.yf[e], gy = 1.f - fy;
            float r0 = res[0][0] * gx + res[0][1] * fx;
            float r1 = res[1][0] * gx + res[1][1] * fx;
            out = r0 * gy + r1 * fy;
        }
        dst[(size_t)e * a.d.pitch + d] = out;
    }
}

cudaError_t launch_level_image(cudaStream_t s, const LevelImageArgs &a)
{
    int SWp = (a.smem_w + 3) & ~3;
    int NC = a.identity ? a.tile_w : a.tile_w * 2;
    size_t smem = (((size_t)a.smem_h * SWp + 15) & ~(size_t)15) + sizeof(float) * ((size_t)a.smem_h * NC + a.ksize);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(level_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid((a.d.w + a.tile_w - 1) / a.tile_w, (a.d.h + a.tile_h - 1) / a.tile_h, a.nimg);
    level_image_kernel<<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K2  polynomial expansion, SURVEY App. A.3.  Separable: vertical pass in float (no FMA) into shared
//     memory (r0, r1, r2), horizontal pass with double accumulators.
// ------------------------------------------------------------------------------------------------
constexpr int PE_TW = 64, PE_TH = 16;

__global__ void __launch_bounds__(256) polyexp_kernel(const float *__restrict__ I, float *__restrict__ R, LevelDims d,
                                                      PolyTables t)
{
    extern __shared__ float pe_smem[];
    const int n = t.n;
    const int SWD = PE_TW + 2 * n;
    float *s0 = pe_smem, *s1 = s0 + PE_TH * SWD, *s2 = s1 + PE_TH * SWD;
    const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
    const float *img = I + (size_t)blockIdx.z * d.plane;
    const int w = d.w, h = d.h, pitch = d.pitch;

    for (int i = threadIdx.x; i < SWD * PE_TH; i += blockDim.x) {
        int row = i / SWD, col = i - row * SWD;
        int gy = y0 + row;
        if (gy >= h) continue;
        int gx = clampi(x0 - n + col, 0, w - 1);
        float r0 = img[(size_t)gy * pitch + gx] * t.g[0], r1 = 0.f, r2 = 0.f;
        for (int k = 1; k <= n; k++) {
            float up = img[(size_t)max(gy - k, 0) * pitch + gx], dn = img[(size_t)min(gy + k, h - 1) * pitch + gx];
            float p = up + dn, q = dn - up;
            r0 = r0 + t.g[k] * p;
            r1 = r1 + t.xg[k] * q;
            r2 = r2 + t.xxg[k] * p;
        }
        s0[i] = r0; s1[i] = r1; s2[i] = r2;
    }
    __syncthreads();

    float *out = R + (size_t)blockIdx.z * 5 * d.plane;
    for (int i = threadIdx.x; i < PE_TW * PE_TH; i += blockDim.x) {
        int row = i / PE_TW, col = i - row * PE_TW;
        int gx = x0 + col, gy = y0 + row;
        if (gx >= w || gy >= h) continue;
        const float *p0 = s0 + row * SWD + col + n, *p1 = s1 + row * SWD + col + n, *p2 = s2 + row * SWD + col + n;
        float g0 = t.g[0];
        double b1 = (double)(p0[0] * g0), b2 = 0, b3 = (double)(p1[0] * g0), b4 = 0, b5 = (double)(p2[0] * g0), b6 = 0;
        for (int k = 1; k <= n; k++) {
            float a0 = p0[k], m0 = p0[-k], a1 = p1[k], m1 = p1[-k], a2 = p2[k], m2 = p2[-k];
            double tg = (double)(a0 + m0);
            b1 = b1 + tg * (double)t.g[k];
            b4 = b4 + tg * (double)t.xxg[k];
            b2 = b2 + (double)((a0 - m0) * t.xg[k]);
            b3 = b3 + (double)((a1 + m1) * t.g[k]);
            b6 = b6 + (double)((a1 - m1) * t.xg[k]);
            b5 = b5 + (double)((a2 + m2) * t.g[k]);
        }
        size_t o = (size_t)gy * pitch + gx;
        out[o] = (float)(b3 * t.ig11);
        out[o + d.plane] = (float)(b2 * t.ig11);
        out[o + 2 * d.plane] = (float)(b1 * t.ig03 + b5 * t.ig33);
        out[o + 3 * d.plane] = (float)(b1 * t.ig03 + b4 * t.ig33);
        out[o + 4 * d.plane] = (float)(b6 * t.ig55);
    }
}

cudaError_t launch_polyexp(cudaStream_t s, const float *I, float *R, const LevelDims &d, int nimg, const PolyTables &t)
{
    dim3 grid((d.w + PE_TW - 1) / PE_TW, (d.h + PE_TH - 1) / PE_TH, nimg);
    size_t smem = sizeof(float) * 3 * PE_TH * (PE_TW + 2 * t.n);
    polyexp_kernel<<<grid, 256, smem, s>>>(I, R, d, t);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// A.4  update matrices for one pixel.  All float, no FMA.  R0/R1 point at channel 0 of the pair's planes.
// ------------------------------------------------------------------------------------------------
// border damping {0.14, 0.14, 0.4472, 0.4472, 0.4472} indexed by the distance to the edge (App. A.4)
__device__ __forceinline__ float border_tab(int i) { return i < 2 ? 0.14f : 0.4472f; }

__device__ __forceinline__ void update_matrices_px(const float *__restrict__ R0, const float *__restrict__ R1, size_t plane,
                                                   int pitch, int w, int h, int x, int y, float dx, float dy, float m[5])
{
    const size_t o = (size_t)y * pitch + x;
    float fx = (float)x + dx, fy = (float)y + dy;
    int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
    fx = fx - (float)x1;
    fy = fy - (float)y1;
    float r2, r3, r4, r5, r6;
    const float q0 = R0[o], q1 = R0[o + plane], q2 = R0[o + 2 * plane], q3 = R0[o + 3 * plane], q4 = R0[o + 4 * plane];
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const float *p = R1 + (size_t)y1 * pitch + x1;
        r2 = a00 * p[0] + a01 * p[1] + a10 * p[pitch] + a11 * p[pitch + 1]; p += plane;
        r3 = a00 * p[0] + a01 * p[1] + a10 * p[pitch] + a11 * p[pitch + 1]; p += plane;
        r4 = a00 * p[0] + a01 * p[1] + a10 * p[pitch] + a11 * p[pitch + 1]; p += plane;
        r5 = a00 * p[0] + a01 * p[1] + a10 * p[pitch] + a11 * p[pitch + 1]; p += plane;
        r6 = a00 * p[0] + a01 * p[1] + a10 * p[pitch] + a11 * p[pitch + 1];
        r4 = (q2 + r4) * 0.5f;
        r5 = (q3 + r5) * 0.5f;
        r6 = (q4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = q2;
        r5 = q3;
        r6 = q4 * 0.5f;
    }
    r2 = (q0 - r2) * 0.5f;
    r3 = (q1 - r3) * 0.5f;
    r2 = r2 + (r4 * dy + r6 * dx);
    r3 = r3 + (r6 * dy + r5 * dx);
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        float sc = (x < 5 ? border_tab(x) : 1.f) * (x >= w - 5 ? border_tab(w - x - 1) : 1.f) * (y < 5 ? border_tab(y) : 1.f) *
                   (y >= h - 5 ? border_tab(h - y - 1) : 1.f);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

__device__ __forceinline__ void solve2x2(float g11f, float g12f, float g22f, float h1f, float h2f, float &fx, float &fy)
{
    double g11 = g11f, g12 = g12f, g22 = g22f, h1 = h1f, h2 = h2f;
    double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
    fx = (float)((g11 * h2 - g12 * h1) * idet);
    fy = (float)((g22 * h1 - g12 * h2) * idet);
}

__device__ __forceinline__ void solve2x2d(double g11, double g12, double g22, double h1, double h2, float &fx, float &fy)
{
    double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
    fx = (float)((g11 * h2 - g12 * h1) * idet);
    fy = (float)((g22 * h1 - g12 * h2) * idet);
}

// ------------------------------------------------------------------------------------------------
// K3  flow initialisation for a scale (zero, or bilinear upsample of the coarser flow * 1/pyrScale,
//     App. A.1 / A.2b) fused with the first update-matrices (A.4).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) first_update_kernel(FirstUpdateArgs a)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.d.w || y >= a.d.h) return;
    const int b = blockIdx.z;
    float dx = 0.f, dy = 0.f;
    if (a.coarse) {
        const float *cf = a.coarse + (size_t)b * 2 * a.cd.plane;
        int x0 = a.xi[x], x1 = min(x0 + 1, a.cd.w - 1), y0 = a.yi[y], y1 = min(y0 + 1, a.cd.h - 1);
        float fx = a.xf[x], gx = 1.f - fx, fy = a.yf[y], gy = 1.f - fy;
        const float *r0 = cf + (size_t)y0 * a.cd.pitch, *r1 = cf + (size_t)y1 * a.cd.pitch;
        float t0 = r0[x0] * gx + r0[x1] * fx, t1 = r1[x0] * gx + r1[x1] * fx;
        dx = (t0 * gy + t1 * fy) * a.inv_scale;
        r0 += a.cd.plane; r1 += a.cd.plane;
        t0 = r0[x0] * gx + r0[x1] * fx; t1 = r1[x0] * gx + r1[x1] * fx;
        dy = (t0 * gy + t1 * fy) * a.inv_scale;
    }
    const size_t o = (size_t)y * a.d.pitch + x;
    if (a.flow_out) {
        float *f = a.flow_out + (size_t)b * 2 * a.d.plane;
        f[o] = dx; f[o + a.d.plane] = dy;
    }
    if (a.M) {
        float m[5];
        const float *R0 = a.R + (size_t)b * 10 * a.d.plane, *R1 = R0 + 5 * a.d.plane;
        update_matrices_px(R0, R1, a.d.plane, a.d.pitch, a.d.w, a.d.h, x, y, dx, dy, m);
        float *M = a.M + (size_t)b * 5 * a.d.plane;
#pragma unroll
        for (int c = 0; c < 5; c++) M[o + c * a.d.plane] = m[c];
    }
}

cudaError_t launch_first_update(cudaStream_t s, const FirstUpdateArgs &a)
{
    dim3 grid((a.d.w + 31) / 32, (a.d.h + 7) / 8, a.batch);
    first_update_kernel<<<grid, dim3(32, 8), 0, s>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K4/K5 (generic radius)  Gaussian window blur (A.5) of the five M planes, 2x2 solve in double, then either
//     the next update-matrices (A.4) or, on the last iteration, the flow write.  One block = 64x16 tile;
//     per channel: vertical pass (global -> shared), horizontal pass (shared -> registers).
// ------------------------------------------------------------------------------------------------
constexpr int GI_TW = 64, GI_TH = 16;

__global__ void __launch_bounds__(256) gauss_iter_generic_kernel(IterArgs a, WinTaps t)
{
    extern __shared__ float gi_smem[];
    const int m = t.m, SWD = GI_TW + 2 * m;
    const int x0 = blockIdx.x * GI_TW, y0 = blockIdx.y * GI_TH, b = blockIdx.z;
    const int w = a.d.w, h = a.d.h, pitch = a.d.pitch;
    const size_t plane = a.d.plane;
    const float *Min = a.Min + (size_t)b * 5 * plane;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float acc[5][4];
    for (int c = 0; c < 5; c++) {
        const float *Mc = Min + c * plane;
        for (int i = threadIdx.x; i < SWD * GI_TH; i += blockDim.x) {
            int row = i / SWD, col = i - row * SWD;
            int gy = y0 + row;
            if (gy >= h) continue;
            int gx = clampi(x0 - m + col, 0, w - 1);
            float v = Mc[(size_t)gy * pitch + gx] * t.k[0];
            for (int k = 1; k <= m; k++)
                v = v + (Mc[(size_t)min(gy + k, h - 1) * pitch + gx] + Mc[(size_t)max(gy - k, 0) * pitch + gx]) * t.k[k];
            gi_smem[i] = v;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int col = tx + 32 * (j & 1), row = ty + 8 * (j >> 1);
            const float *p = gi_smem + row * SWD + col + m;
            float v = p[0] * t.k[0];
            for (int k = 1; k <= m; k++) v = v + t.k[k] * (p[-k] + p[k]);
            acc[c][j] = v;
        }
        __syncthreads();
    }
    const float *R0 = a.R + (size_t)b * 10 * plane, *R1 = R0 + 5 * plane;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int x = x0 + tx + 32 * (j & 1), y = y0 + ty + 8 * (j >> 1);
        if (x >= w || y >= h) continue;
        float fx, fy;
        solve2x2(acc[0][j], acc[1][j], acc[2][j], acc[3][j], acc[4][j], fx, fy);
        size_t o = (size_t)y * pitch + x;
        if (a.last) {
            float *f = a.flow + (size_t)b * 2 * plane;
            f[o] = fx; f[o + plane] = fy;
        } else {
            float mm[5];
            update_matrices_px(R0, R1, plane, pitch, w, h, x, y, fx, fy, mm);
            float *M = a.Mout + (size_t)b * 5 * plane;
#pragma unroll
            for (int c = 0; c < 5; c++) M[o + c * plane] = mm[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4/K5 (window radius MR = 15 or 7)  register-blocked Gaussian window blur + solve + update.
//   tile 96 x 32 outputs, 256 threads, 2 CTAs / SM.
//   phase V: thread = (column, 16-row group); 16+2*MR inputs in registers (coalesced global loads, row/column
//            replicate = clamped addresses), 16 outputs, taps summed in the oracle's order; -> shared Vb[5][32][132]
//   phase H: lane = row (pitch 132 = 4 mod 32 words: conflict-free LDS.128), warp = 12-column segment, 3 groups of
//            4 pixels x 5 channels; 2x2 solve in double; flow -> shared Fb[2][32][97]
//   phase U: lane = x (coalesced): next update-matrices (A.4) or, last iteration, the flow write.
// FMA = validated relaxation (fmaf in the tap sums, SURVEY App. B.5: <= 2e-4 px on the default options); the
// default build keeps the oracle's add-mul-add order bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int GK_TW = 96, GK_TH = 32, GK_VW = 128, GK_VP = 132, GK_FP = 97, GK_RV = 16;
constexpr size_t GK_SMEM = sizeof(float) * (5 * GK_TH * GK_VP + 2 * GK_TH * GK_FP);

template <int MR, bool FMA>
__global__ void __launch_bounds__(256, 2) gauss_iter_kernel(IterArgs a, WinTaps t)
{
    extern __shared__ __align__(16) float gk_smem[];
    float *Vb = gk_smem;
    float *Fb = gk_smem + 5 * GK_TH * GK_VP;
    const int tid = threadIdx.x, lane = tid & 31;
    const int x0 = blockIdx.x * GK_TW, y0 = blockIdx.y * GK_TH, b = blockIdx.z;
    const int w = a.d.w, h = a.d.h, pitch = a.d.pitch;
    const size_t plane = a.d.plane;
    const float *Min = a.Min + (size_t)b * 5 * plane;

    // ---- phase V ----
    {
        const int j = tid & (GK_VW - 1), g = tid >> 7;
        const int gx = clampi(x0 - 16 + j, 0, w - 1);
        const int ybase = y0 + g * GK_RV - MR;
        const bool interior = ybase >= 0 && ybase + GK_RV + 2 * MR - 1 <= h - 1;
#pragma unroll 1
        for (int c = 0; c < 5; c++) {
            const float *Mc = Min + c * plane + gx;
            float in[GK_RV + 2 * MR];
            if (interior) {
                const float *p = Mc + (size_t)ybase * pitch;
#pragma unroll
                for (int r = 0; r < GK_RV + 2 * MR; r++) in[r] = __ldg(p + (size_t)r * pitch);
            } else {
#pragma unroll
                for (int r = 0; r < GK_RV + 2 * MR; r++) in[r] = __ldg(Mc + (size_t)clampi(ybase + r, 0, h - 1) * pitch);
            }
            float *dst = Vb + (c * GK_TH + g * GK_RV) * GK_VP + j;
#pragma unroll
            for (int o = 0; o < GK_RV; o++) {
                float v = in[o + MR] * t.k[0];
#pragma unroll
                for (int i = 1; i <= MR; i++) {
                    if (FMA) v = fmaf(in[o + MR + i] + in[o + MR - i], t.k[i], v);
                    else v = v + (in[o + MR + i] + in[o + MR - i]) * t.k[i];
                }
                dst[o * GK_VP] = v;
            }
        }
    }
    __syncthreads();

    // ---- phase H + solve ----
    {
        const int seg = tid >> 5; // 12-column segment
        constexpr int LO = (16 - MR) & ~3;                // first needed index, 16B-aligned
        constexpr int NV = ((3 + 16 + MR) | 3) + 1 - LO;  // floats loaded per channel per group (multiple of 4)
#pragma unroll 1
        for (int grp = 0; grp < 3; grp++) {
            const int cbase = seg * 12 + grp * 4; // output column of pixel 0; V column = cbase + 16
            float res[5][4];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                const float4 *src = reinterpret_cast<const float4 *>(Vb + (c * GK_TH + lane) * GK_VP + cbase + LO);
                float v[NV];
#pragma unroll
                for (int q = 0; q < NV / 4; q++) {
                    float4 u = src[q];
                    v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
                }
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const int ctr = p + 16 - LO;
                    float s = v[ctr] * t.k[0];
#pragma unroll
                    for (int i = 1; i <= MR; i++) {
                        if (FMA) s = fmaf(t.k[i], v[ctr - i] + v[ctr + i], s);
                        else s = s + t.k[i] * (v[ctr - i] + v[ctr + i]);
                    }
                    res[c][p] = s;
                }
            }
#pragma unroll
            for (int p = 0; p < 4; p++) {
                float fx, fy;
                solve2x2(res[0][p], res[1][p], res[2][p], res[3][p], res[4][p], fx, fy);
                Fb[lane * GK_FP + cbase + p] = fx;
                Fb[(GK_TH + lane) * GK_FP + cbase + p] = fy;
            }
        }
    }
    __syncthreads();

    // ---- phase U: coalesced epilogue ----
    const float *R0 = a.R + (size_t)b * 10 * plane, *R1 = R0 + 5 * plane;
#pragma unroll 1
    for (int i = tid; i < GK_TW * GK_TH; i += 256) {
        const int row = i / GK_TW, col = i - row * GK_TW;
        const int x = x0 + col, y = y0 + row;
        if (x >= w || y >= h) continue;
        const float fx = Fb[row * GK_FP + col], fy = Fb[(GK_TH + row) * GK_FP + col];
        const size_t o = (size_t)y * pitch + x;
        if (a.last) {
            float *f = a.flow + (size_t)b * 2 * plane;
            f[o] = fx; f[o + plane] = fy;
        } else {
            float mm[5];
            update_matrices_px(R0, R1, plane, pitch, w, h, x, y, fx, fy, mm);
            float *M = a.Mout + (size_t)b * 5 * plane;
#pragma unroll
            for (int c = 0; c < 5; c++) M[o + c * plane] = mm[c];
        }
    }
}

template <int MR, bool FMA>
static cudaError_t launch_gauss_fast(cudaStream_t s, const IterArgs &a, const WinTaps &t)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gauss_iter_kernel<MR, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GK_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((a.d.w + GK_TW - 1) / GK_TW, (a.d.h + GK_TH - 1) / GK_TH, a.batch);
    gauss_iter_kernel<MR, FMA><<<grid, 256, GK_SMEM, s>>>(a, t);
    return cudaGetLastError();
}

cudaError_t launch_gauss_iter(cudaStream_t s, const IterArgs &a, const WinTaps &t)
{
    if (t.m == 15) return a.fma ? launch_gauss_fast<15, true>(s, a, t) : launch_gauss_fast<15, false>(s, a, t);
    if (t.m == 7) return a.fma ? launch_gauss_fast<7, true>(s, a, t) : launch_gauss_fast<7, false>(s, a, t);
    dim3 grid((a.d.w + GI_TW - 1) / GI_TW, (a.d.h + GI_TH - 1) / GI_TH, a.batch);
    size_t smem = sizeof(float) * GI_TH * (GI_TW + 2 * t.m);
    gauss_iter_generic_kernel<<<grid, 256, smem, s>>>(a, t);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Box window (flags == 0), App. A.6.
//   box_vsum: one thread per (column, plane) walks down the column keeping the oracle's running sum:
//             vs += double(float(M[y+m] - M[y-m-1])) -- the float difference is part of the answer.
//   box_iter: horizontal window of 2m+1 doubles (replicate border), * 1/winSize^2, solve, update/flow.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) box_vsum_kernel(const float *__restrict__ Min, double *__restrict__ V, LevelDims d, int m)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= d.w) return;
    const float *M = Min + (size_t)blockIdx.y * d.plane + x;
    double *v = V + (size_t)blockIdx.y * d.plane + x;
    const int h = d.h, pitch = d.pitch;
    double vs = (double)(M[0] * (float)(m + 2));
    for (int y = 1; y < m; y++) vs = vs + (double)M[(size_t)min(y, h - 1) * pitch];
    for (int y = 0; y < h; y++) {
        float diff = M[(size_t)min(y + m, h - 1) * pitch] - M[(size_t)max(y - m - 1, 0) * pitch];
        vs = vs + (double)diff;
        v[(size_t)y * pitch] = vs;
    }
}

cudaError_t launch_box_vsum(cudaStream_t s, const float *Min, double *V, const LevelDims &d, int batch, int m)
{
    dim3 grid((d.w + 127) / 128, batch * 5);
    box_vsum_kernel<<<grid, 128, 0, s>>>(Min, V, d, m);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) box_iter_kernel(const double *__restrict__ V, IterArgs a, int m, double scale)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const int w = a.d.w, h = a.d.h, pitch = a.d.pitch, b = blockIdx.z;
    if (x >= w || y >= h) return;
    const size_t plane = a.d.plane;
    double g[5];
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const double *row = V + ((size_t)b * 5 + c) * plane + (size_t)y * pitch;
        double s = 0;
        for (int j = -m; j <= m; j++) s = s + row[clampi(x + j, 0, w - 1)];
        g[c] = s * scale;
    }
    float fx, fy;
    solve2x2d(g[0], g[1], g[2], g[3], g[4], fx, fy);
    size_t o = (size_t)y * pitch + x;
    if (a.last) {
        float *f = a.flow + (size_t)b * 2 * plane;
        f[o] = fx; f[o + plane] = fy;
    } else {
        const float *R0 = a.R + (size_t)b * 10 * plane, *R1 = R0 + 5 * plane;
        float mm[5];
        update_matrices_px(R0, R1, plane, pitch, w, h, x, y, fx, fy, mm);
        float *M = a.Mout + (size_t)b * 5 * plane;
#pragma unroll
        for (int c = 0; c < 5; c++) M[o + c * plane] = mm[c];
    }
}

cudaError_t launch_box_iter(cudaStream_t s, const double *V, const IterArgs &a, int m, int winSize)
{
    dim3 grid((a.d.w + 31) / 32, (a.d.h + 7) / 8, a.batch);
    box_iter_kernel<<<grid, dim3(32, 8), 0, s>>>(V, a, m, 1. / ((double)winSize * winSize));
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Span sampling + threshold test + ordered compaction (reference src/consumer.cpp:60-77):
//   len = f32(f32(dx*dx) + f32(dy*dy));  keep iff double(len) > threshold*threshold (strict), row-major order.
// One block per pair; each thread owns a contiguous run of sample points so a block-wide exclusive scan of
// the per-thread counts yields row-major output positions.  Counts are reduced with warp shuffles.
// ------------------------------------------------------------------------------------------------
struct DevVector { int x, y; double dx, dy; };

__global__ void __launch_bounds__(1024) sample_kernel(SampleArgs a)
{
    __shared__ int warp_sums[32];
    __shared__ int total;
    const int b = blockIdx.x;
    const float *fxp = a.flow + (size_t)b * 2 * a.d.plane, *fyp = fxp + a.d.plane;
    const int nsx = (a.d.w + a.span - 1) / a.span, nsy = (a.d.h + a.span - 1) / a.span, ns = nsx * nsy;
    const int per = (ns + blockDim.x - 1) / blockDim.x;
    const int i0 = min(threadIdx.x * per, ns), i1 = min(i0 + per, ns);
    int cnt = 0;
    for (int i = i0; i < i1; i++) {
        int sy = i / nsx, sx = i - sy * nsx;
        size_t o = (size_t)(sy * a.span) * a.d.pitch + sx * a.span;
        float dx = fxp[o], dy = fyp[o];
        float len = (dx * dx) + (dy * dy);
        cnt += ((double)len > a.thr2) ? 1 : 0;
    }
    // block exclusive scan of cnt
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        int s = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += u;
        }
        warp_sums[lane] = s - v; // exclusive
        if (lane == 31) total = s;
    }
    __syncthreads();
    int pos = warp_sums[wid] + incl - cnt;
    if (threadIdx.x == 0) a.counts[b] = total;
    if (cnt == 0) return;
    DevVector *out = reinterpret_cast<DevVector *>(a.vectors) + (size_t)b * a.cap;
    for (int i = i0; i < i1; i++) {
        int sy = i / nsx, sx = i - sy * nsx;
        size_t o = (size_t)(sy * a.span) * a.d.pitch + sx * a.span;
        float dx = fxp[o], dy = fyp[o];
        float len = (dx * dx) + (dy * dy);
        if ((double)len > a.thr2) {
            if (pos < a.cap) {
                DevVector v;
                v.x = sx * a.span; v.y = sy * a.span; v.dx = (double)dx; v.dy = (double)dy;
                out[pos] = v;
            }
            pos++;
        }
    }
}

cudaError_t launch_sample(cudaStream_t s, const SampleArgs &a)
{
    sample_kernel<<<a.batch, 1024, 0, s>>>(a);
    return cudaGetLastError();
}

} // namespace tw
