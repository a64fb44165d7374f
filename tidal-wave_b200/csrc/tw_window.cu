// tw_window.cu -- K4 / K5 for the Gaussian window of radius 15 (winSize 30 / 31, the reference's default): ONE persistent,
// warp-specialised CTA per SM that slides down 96-column strips of the M planes.
//
//   Farneback stages A.5 (window blur of the five M planes + 2x2 solve) and A.4 (next update-matrices) of SURVEY App. A --
//   the per-iteration body of cv::calcOpticalFlowFarneback, /root/reference/src/opticalflow.cpp:83-85 -- fused into one
//   pass: per iteration M (20 B/px), R0 (20), R1 (20) are read once and M' (20) is written once.
//
// Roles of the 16 warps of a CTA (one CTA per SM, 128 registers per thread, 208 KB of shared memory):
//   warps 0-9    V walkers: thread = (column, channel pair) [warps 0-7] or (column pair, h2) [warps 8-9].  Each keeps a 40-row
//                sliding window of its column in REGISTERS for the whole strip, so an M row is read from L2 exactly once per
//                strip (the tile kernel re-read it 1.94x) and there is no cold start per tile; per 8-row group: 8 new rows
//                from the M ring, 8 outputs x 31 taps as packed f32x2 (FFMA2), taps in the oracle's order -> the P ring.
//   warps 10-15  H + solve + U: horizontal taps from the P ring (lane = row x 4-pixel quad, conflict-free LDS.128), 2x2 solve in
//                double -> flow tile in shared memory; then (not last) the next update matrices (A.4) with R0 / R1 read from
//                shared memory, two vertically adjacent pixels per packed f32x2 instruction on the warp-uniform fast path
//                (all pixels inside, within the staged R1 window, off the damped frame border), a scalar per-pixel path
//                otherwise (pixels whose displaced position leaves the staged window -- large motion -- gather from global
//                memory; same values, same arithmetic) -> M' stores; (last) flow stores + the fused span-grid threshold count.
//   TMA producers  tensor-map bulk copies (cp.async.bulk.tensor, UTMALDG), issued as far ahead as the rings allow by
//                lane 0 of warp 8 (the M rows: 8-row chunks, one 8-row box per plane; at the frame top / bottom one row per
//                copy with the row coordinate clamped = the replicate border of App. A.5; 3-slot ring) and lane 0 of warp 9
//                (the R0 tile of each group, 96 x 8 x 5 channels, 2 slots; the R1 rows its bilinear gather can touch, 104
//                columns x rows y-8 .. y+15, a 4-chunk row ring).  Out-of-frame parts are zero-filled by the TMA unit and
//                never read.  Both poll (mbarrier.test_wait) and never block, so no producer can stall a pipeline stage.
//   All hand-offs are mbarrier pipelines (full / empty per ring slot); the H -> U hand-off inside warps 10-15 is a named
//   barrier.  Every wait sleeps between probes and traps after ~2 s instead of hanging the device.
//
// Arithmetic per output is exactly that of gauss_iter2_kernel (tw_kernels.cu): FMA = 0 the oracle's add-mul-add order
// (bit-identical to oracle/farneback_ref.c), FMA = 2 the direct-form fmaf taps of the relaxed default (oracle relax bit 7).
#include "tw_device.cuh"

#include <cuda.h>

#include <algorithm>

namespace tw {

namespace {

constexpr int WS_SW = 96;                 // output columns per strip
constexpr int WS_XH = 16;                 // left halo of the V columns (15 needed, 16 keeps the copies 64-byte aligned)
constexpr int WS_VC = WS_SW + 2 * WS_XH;  // 128 V columns
constexpr int WS_G = 8;                   // rows per group
constexpr int WS_MR = 15;
constexpr int WS_WIN = 40;                // register window rows = 5 chunks
constexpr int WS_NM = 3, WS_NP = 2, WS_NR0 = 2, WS_NR1 = 4, WS_NF = 2;
constexpr int WS_MPROW = 2 * WS_VC * 4, WS_MHROW = WS_VC * 4;         // bytes per staged row of a channel-pair plane / of the h2 plane
constexpr int WS_MP1OFF = WS_G * WS_MPROW, WS_MHOFF = 2 * WS_MP1OFF;  // chunk layout: pair01[8 rows] | pair23[8 rows] | h2[8 rows]
constexpr int WS_MCHUNK = WS_MHOFF + WS_G * WS_MHROW;                 // 20480
constexpr int WS_P2 = 130, WS_P4 = 132;                               // float2 / float pitches of the P planes
constexpr int WS_P23OFF = WS_G * WS_P2 * 8, WS_P4OFF = 2 * WS_P23OFF; // 8320, 16640
constexpr int WS_PSLOT = WS_P4OFF + WS_G * WS_P4 * 4;                 // 20864
constexpr int WS_R0SLOT = WS_G * 5 * WS_SW * 4;                       // 15360
constexpr int WS_R1C = 104, WS_R1X = 4;                               // staged R1 columns: x0 - 4 .. x0 + 99
constexpr int WS_R1ROW = 5 * WS_R1C;                                  // floats per ring row
constexpr int WS_R1CHUNK = WS_G * WS_R1ROW * 4;                       // 16640
constexpr int WS_FP = 97;                                             // flow tile pitch
constexpr int WS_FSLOT = 2 * WS_G * WS_FP * 4;                        // 6208

constexpr int WS_OFF_M = 0;
constexpr int WS_OFF_P = WS_OFF_M + WS_NM * WS_MCHUNK;      // 61440
constexpr int WS_OFF_R0 = WS_OFF_P + WS_NP * WS_PSLOT;      // 103168
constexpr int WS_OFF_R1 = WS_OFF_R0 + WS_NR0 * WS_R0SLOT;   // 133888
constexpr int WS_OFF_F = WS_OFF_R1 + WS_NR1 * WS_R1CHUNK;   // 200448
constexpr int WS_OFF_BAR = WS_OFF_F + WS_NF * WS_FSLOT;     // 212864
constexpr int WS_NBAR = 2 * (WS_NM + WS_NP + WS_NR0 + WS_NR1);
constexpr int WS_SMEM = WS_OFF_BAR + WS_NBAR * 8;
static_assert(WS_OFF_R0 % 128 == 0 && WS_OFF_R1 % 128 == 0 && WS_R0SLOT % 128 == 0 && WS_R1CHUNK % 128 == 0 && WS_MPROW % 128 == 0 && WS_MHROW % 128 == 0,
              "TMA destinations are 128-byte aligned");
static_assert(WS_OFF_BAR % 8 == 0 && WS_SMEM <= 227 * 1024, "shared memory budget");

constexpr int WS_NVW = 10, WS_NHW = 6;                 // V walker warps, H / U warps
constexpr int WS_THREADS = (WS_NVW + WS_NHW) * 32;     // 512: the register file of an SM at 128 registers per thread
constexpr int WS_HU_THREADS = WS_NHW * 32;

// barrier indices
constexpr int B_FULLM = 0, B_EMPTYM = B_FULLM + WS_NM, B_FULLP = B_EMPTYM + WS_NM, B_EMPTYP = B_FULLP + WS_NP,
              B_FULLR0 = B_EMPTYP + WS_NP, B_EMPTYR0 = B_FULLR0 + WS_NR0, B_FULLR1 = B_EMPTYR0 + WS_NR0,
              B_EMPTYR1 = B_FULLR1 + WS_NR1;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned mbar_try(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done;
}
// try_wait with a suspend-time hint: the warp may sleep in hardware for up to `ns` before the instruction returns
__device__ __forceinline__ unsigned mbar_try_sleep(unsigned bar, unsigned parity, unsigned ns)
{
    unsigned done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity), "r"(ns)
                 : "memory");
    return done;
}
// non-blocking probe (the producers poll with it)
__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done;
}
// Waits for the phase with the given parity: try_wait suspends the warp in hardware for a short, implementation-defined time
// and is simply repeated.  (A version that slept between probes -- suspend-time hint + nanosleep -- woke the V walkers late
// enough that the H / U warps ran out of P slots 13 % of the time.)  A wait of more than ~2 s means a broken pipeline: trap
// (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    if (mbar_try(bar, parity)) return;
#pragma unroll 1
    for (unsigned it = 0; !mbar_try(bar, parity); ++it)
        if (it > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int x, int y, int z, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ float lds32(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void hu_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(WS_HU_THREADS) : "memory"); }

struct Unit {
    int b, x0, y0, ng;
};

} // namespace

struct StripArgs {
    const float *R; // [B][2][5] planes (global gathers of pixels that leave the staged R1 window)
    float *Mout;    // [B][5] planes
    float *flow;    // [B][2] planes (LAST)
    LevelDims d;
    int nsx, nsy, segh, nunits;
    int span;
    double thr2;
    int *counts;
};

namespace {

__device__ __forceinline__ Unit decode_unit(const StripArgs &a, int u)
{
    Unit U;
    const int sy = u % a.nsy, rest = u / a.nsy;
    const int sx = rest % a.nsx;
    U.b = rest / a.nsx;
    U.x0 = sx * WS_SW;
    U.y0 = sy * a.segh;
    U.ng = (min(a.segh, a.d.h - U.y0) + WS_G - 1) / WS_G;
    return U;
}

// ---- TMA producers: streams that run across the CTA's unit sequence; pump() issues while a ring slot is free ----
// M rows: every unit contributes ng + 4 chunks of 8 rows (rows y0 - 15 + 8c ..).
struct MStream {
    int u, c;
    Unit U;
    unsigned n; // chunks issued so far
};

__device__ __noinline__ void mstream_pump(MStream &ms, const StripArgs &a, unsigned smem, const CUtensorMap *mp1, const CUtensorMap *mh1,
                                          const CUtensorMap *mp8, const CUtensorMap *mh8)
{
    while (ms.u < a.nunits) {
        const unsigned s = ms.n % WS_NM;
        if (!mbar_test(smem + WS_OFF_BAR + 8 * (B_EMPTYM + s), ((ms.n / WS_NM) & 1) ^ 1)) return;
        const unsigned bar = smem + WS_OFF_BAR + 8 * (B_FULLM + s);
        mbar_expect_tx(bar, WS_MCHUNK);
        const unsigned dst = smem + WS_OFF_M + s * WS_MCHUNK;
        const int xs = ms.U.x0 - WS_XH, pitch = a.d.pitch, h = a.d.h;
        const int yf = ms.U.y0 - WS_MR + WS_G * ms.c;
        if (yf >= 0 && yf + WS_G - 1 <= h - 1) {
            tma_load_3d(dst, mp8, 2 * xs, yf, ms.U.b, bar);
            tma_load_3d(dst + WS_MP1OFF, mp8, 2 * pitch + 2 * xs, yf, ms.U.b, bar);
            tma_load_3d(dst + WS_MHOFF, mh8, 4 * pitch + xs, yf, ms.U.b, bar);
        } else { // frame top / bottom: replicate = clamped row coordinate, one row per copy
#pragma unroll 1
            for (int r = 0; r < WS_G; r++) {
                const int y = clampi(yf + r, 0, h - 1);
                tma_load_3d(dst + r * WS_MPROW, mp1, 2 * xs, y, ms.U.b, bar);
                tma_load_3d(dst + WS_MP1OFF + r * WS_MPROW, mp1, 2 * pitch + 2 * xs, y, ms.U.b, bar);
                tma_load_3d(dst + WS_MHOFF + r * WS_MHROW, mh1, 4 * pitch + xs, y, ms.U.b, bar);
            }
        }
        ms.n++;
        if (++ms.c == ms.U.ng + 4) {
            ms.c = 0;
            ms.u += gridDim.x;
            if (ms.u < a.nunits) ms.U = decode_unit(a, ms.u);
        }
    }
}

// R0 tiles (one per group) and R1 row chunks (ng + 2 per unit: rows y0 - 8 + 8c .. + 7 of the target image's expansion).
struct RStream {
    int u0, g0, u1, c1;
    Unit U0, U1;
    unsigned n0, n1;
};

__device__ __noinline__ void rstream_pump(RStream &rs, const StripArgs &a, unsigned smem, const CUtensorMap *r0map, const CUtensorMap *r1map)
{
    while (rs.u0 < a.nunits) {
        const unsigned s = rs.n0 % WS_NR0;
        if (!mbar_test(smem + WS_OFF_BAR + 8 * (B_EMPTYR0 + s), ((rs.n0 / WS_NR0) & 1) ^ 1)) break;
        const unsigned bar = smem + WS_OFF_BAR + 8 * (B_FULLR0 + s);
        mbar_expect_tx(bar, WS_R0SLOT);
        tma_load_3d(smem + WS_OFF_R0 + s * WS_R0SLOT, r0map, rs.U0.x0, 5 * (rs.U0.y0 + WS_G * rs.g0), 2 * rs.U0.b, bar);
        rs.n0++;
        if (++rs.g0 == rs.U0.ng) {
            rs.g0 = 0;
            rs.u0 += gridDim.x;
            if (rs.u0 < a.nunits) rs.U0 = decode_unit(a, rs.u0);
        }
    }
    while (rs.u1 < a.nunits) {
        const unsigned s = rs.n1 % WS_NR1;
        if (!mbar_test(smem + WS_OFF_BAR + 8 * (B_EMPTYR1 + s), ((rs.n1 / WS_NR1) & 1) ^ 1)) break;
        const unsigned bar = smem + WS_OFF_BAR + 8 * (B_FULLR1 + s);
        mbar_expect_tx(bar, WS_R1CHUNK);
        tma_load_3d(smem + WS_OFF_R1 + s * WS_R1CHUNK, r1map, rs.U1.x0 - WS_R1X, 5 * (rs.U1.y0 - WS_G + WS_G * rs.c1), 2 * rs.U1.b + 1, bar);
        rs.n1++;
        if (++rs.c1 == rs.U1.ng + 2) {
            rs.c1 = 0;
            rs.u1 += gridDim.x;
            if (rs.u1 < a.nunits) rs.U1 = decode_unit(a, rs.u1);
        }
    }
}

// The per-thread view of a V walker: two source words per staged row (srcA, srcB: the clamped columns) and one float2 output.
struct VLane {
    unsigned srcA, srcB, rstride, dst, dstride;
};

// 8 outputs x 31 taps of one walker from its register window (rows 8g .. 8g + 39 of the unit's M rows) -> P slot.
template <int FMA>
__device__ __forceinline__ void v_compute(const float2 (&win)[WS_WIN], unsigned pslot, const VLane &vl, const WinTaps &t)
{
    constexpr int J = 0;
    const float2 one2 = make_float2(t.one, t.one);
    float2 v[WS_G];
#pragma unroll
    for (int o = 0; o < WS_G; o++) v[o] = tw_mul2(win[(8 * J + o + WS_MR) % WS_WIN], make_float2(t.k[0], t.k[0]));
#pragma unroll
    for (int i = 1; i <= WS_MR; i++) {
        const float2 kk = make_float2(t.k[i], t.k[i]);
        if (FMA == 2) { // direct form (oracle relax bit 7): upper row first
#pragma unroll
            for (int o = 0; o < WS_G; o++) v[o] = tw_fma2(win[(8 * J + o + WS_MR - i) % WS_WIN], kk, v[o]);
#pragma unroll
            for (int o = 0; o < WS_G; o++) v[o] = tw_fma2(win[(8 * J + o + WS_MR + i) % WS_WIN], kk, v[o]);
        } else {
#pragma unroll
            for (int o = 0; o < WS_G; o++) {
                const float2 sum = tw_add2(win[(8 * J + o + WS_MR + i) % WS_WIN], win[(8 * J + o + WS_MR - i) % WS_WIN]);
                v[o] = tw_fma2(tw_mul2(sum, kk), one2, v[o]); // = v + round(sum * k): see tw_fma2 in tw_device.cuh
            }
        }
    }
    const unsigned pd = pslot + vl.dst;
#pragma unroll
    for (int o = 0; o < WS_G; o++) sts64(pd + o * vl.dstride, v[o]);
}

// Horizontal taps + solve for 4 adjacent pixels of one row (P slot pointers already offset to the row).
template <int FMA>
__device__ __forceinline__ void h_quad(const float2 *__restrict__ p01, const float2 *__restrict__ p23, const float *__restrict__ p4, const WinTaps &t,
                                       float (&fx)[4], float (&fy)[4])
{
    constexpr int MR = WS_MR;
    constexpr int LO = (16 - MR) & ~3;               // first needed V column relative to the quad, 16-byte aligned
    constexpr int NV = ((3 + 16 + MR) | 3) + 1 - LO; // values loaded per plane (multiple of 4)
    const float2 one2 = make_float2(t.one, t.one);
    float2 r01[4], r23[4];
    float r4[4];
#pragma unroll
    for (int pr = 0; pr < 2; pr++) {
        const float4 *src = reinterpret_cast<const float4 *>((pr ? p23 : p01) + LO);
        float2 v[NV];
#pragma unroll
        for (int q = 0; q < NV / 2; q++) {
            const float4 u = src[q];
            v[2 * q] = make_float2(u.x, u.y); v[2 * q + 1] = make_float2(u.z, u.w);
        }
        float2 sacc[4];
#pragma unroll
        for (int p = 0; p < 4; p++) sacc[p] = tw_mul2(v[p + 16 - LO], make_float2(t.k[0], t.k[0]));
#pragma unroll
        for (int i = 1; i <= MR; i++) {
            const float2 kk = make_float2(t.k[i], t.k[i]);
            if (FMA == 2) { // direct form: left tap first
#pragma unroll
                for (int p = 0; p < 4; p++) sacc[p] = tw_fma2(kk, v[p + 16 - LO - i], sacc[p]);
#pragma unroll
                for (int p = 0; p < 4; p++) sacc[p] = tw_fma2(kk, v[p + 16 - LO + i], sacc[p]);
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const int ctr = p + 16 - LO;
                    const float2 sum = tw_add2(v[ctr - i], v[ctr + i]);
                    sacc[p] = tw_fma2(tw_mul2(kk, sum), one2, sacc[p]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; p++) { if (pr) r23[p] = sacc[p]; else r01[p] = sacc[p]; }
    }
    {
        const float4 *src = reinterpret_cast<const float4 *>(p4 + LO);
        float v[NV];
#pragma unroll
        for (int q = 0; q < NV / 4; q++) {
            const float4 u = src[q];
            v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
        }
#pragma unroll
        for (int p = 0; p < 4; p += 2) { // pixels (p, p+1) packed
            const int ctr = p + 16 - LO;
            float2 sacc = tw_mul2(make_float2(v[ctr], v[ctr + 1]), make_float2(t.k[0], t.k[0]));
#pragma unroll
            for (int i = 1; i <= MR; i++) {
                const float2 kk = make_float2(t.k[i], t.k[i]);
                if (FMA == 2) {
                    sacc = tw_fma2(kk, make_float2(v[ctr - i], v[ctr + 1 - i]), sacc);
                    sacc = tw_fma2(kk, make_float2(v[ctr + i], v[ctr + 1 + i]), sacc);
                } else {
                    const float2 sum = tw_add2(make_float2(v[ctr - i], v[ctr + 1 - i]), make_float2(v[ctr + i], v[ctr + 1 + i]));
                    sacc = tw_fma2(tw_mul2(kk, sum), one2, sacc);
                }
            }
            r4[p] = sacc.x; r4[p + 1] = sacc.y;
        }
    }
#pragma unroll
    for (int p = 0; p < 4; p++) solve2x2(r01[p].x, r01[p].y, r23[p].x, r23[p].y, r4[p], fx[p], fy[p]);
}

template <int FMA, bool LAST>
__global__ void __launch_bounds__(WS_THREADS, 1)
gauss_strip_kernel(const __grid_constant__ CUtensorMap mapMp, const __grid_constant__ CUtensorMap mapMh, const __grid_constant__ CUtensorMap mapMp8,
                   const __grid_constant__ CUtensorMap mapMh8, const __grid_constant__ CUtensorMap mapR0, const __grid_constant__ CUtensorMap mapR1,
                   const StripArgs a, const WinTaps t)
{
    extern __shared__ __align__(1024) unsigned char ws_smem[];
    const unsigned smem = (unsigned)__cvta_generic_to_shared(ws_smem);
    const unsigned bars = smem + WS_OFF_BAR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int w = a.d.w, h = a.d.h, pitch = a.d.pitch;
    const size_t plane = a.d.plane;

    if (tid == 0) {
        for (int i = 0; i < WS_NM; i++) { mbar_init(bars + 8 * (B_FULLM + i), 1); mbar_init(bars + 8 * (B_EMPTYM + i), WS_NVW); }
        for (int i = 0; i < WS_NP; i++) { mbar_init(bars + 8 * (B_FULLP + i), WS_NVW); mbar_init(bars + 8 * (B_EMPTYP + i), WS_NHW); }
        for (int i = 0; i < WS_NR0; i++) { mbar_init(bars + 8 * (B_FULLR0 + i), 1); mbar_init(bars + 8 * (B_EMPTYR0 + i), WS_NHW); }
        for (int i = 0; i < WS_NR1; i++) { mbar_init(bars + 8 * (B_FULLR1 + i), 1); mbar_init(bars + 8 * (B_EMPTYR1 + i), WS_NHW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < WS_NVW) {
        // ================= V walkers; lane 0 of warps 8 / 9 carries the TMA producers =================
        const bool m_prod = tid == 8 * 32, r_prod = !LAST && tid == 9 * 32;
        MStream ms;
        RStream rs;
        if (m_prod) {
            ms.u = blockIdx.x; ms.c = 0; ms.n = 0;
            if (ms.u < a.nunits) ms.U = decode_unit(a, ms.u);
        }
        if (r_prod) {
            rs.u0 = rs.u1 = blockIdx.x; rs.g0 = rs.c1 = 0; rs.n0 = rs.n1 = 0;
            if (rs.u0 < a.nunits) rs.U0 = rs.U1 = decode_unit(a, rs.u0);
        }
        auto pump = [&]() {
            if (m_prod) mstream_pump(ms, a, smem, &mapMp, &mapMh, &mapMp8, &mapMh8);
            if (r_prod) rstream_pump(rs, a, smem, &mapR0, &mapR1);
        };
        // a producer lane keeps its stream going while its warp waits; the other warps sleep in mbar_wait
        auto wait = [&](unsigned bar, unsigned parity) {
            if (warp >= 8) {
                if (lane == 0) {
                    unsigned it = 0;
                    while (!mbar_test(bar, parity)) {
                        pump();
                        if (++it > (1u << 26)) __trap();
                    }
                }
                __syncwarp();
            }
            mbar_wait(bar, parity);
        };
        pump();

        unsigned nM = 0, nP = 0; // chunks consumed, groups blurred (global over the CTA's units)
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
            const Unit U = decode_unit(a, u);
            const int xs = U.x0 - WS_XH;
            VLane vl;
            if (tid < 2 * WS_VC) { // (column, channel pair)
                const int pair = tid >> 7, j = tid & (WS_VC - 1);
                const int cj = clampi(xs + j, 0, w - 1) - xs;
                vl.srcA = pair * WS_MP1OFF + cj * 8; vl.srcB = vl.srcA + 4; vl.rstride = WS_MPROW;
                vl.dst = (pair ? WS_P23OFF : 0) + j * 8; vl.dstride = WS_P2 * 8;
            } else { // (column pair, h2)
                const int jj = tid - 2 * WS_VC;
                const int ca = clampi(xs + 2 * jj, 0, w - 1) - xs, cb = clampi(xs + 2 * jj + 1, 0, w - 1) - xs;
                vl.srcA = WS_MHOFF + ca * 4; vl.srcB = WS_MHOFF + cb * 4; vl.rstride = WS_MHROW;
                vl.dst = WS_P4OFF + jj * 8; vl.dstride = WS_P4 * 4;
            }
            float2 win[WS_WIN];
#define WS_CONSUME_CHUNK(WIN_INDEX)                                                                                              \
    {                                                                                                                            \
        const unsigned s_ = nM % WS_NM;                                                                                          \
        wait(bars + 8 * (B_FULLM + s_), (nM / WS_NM) & 1);                                                                       \
        const unsigned base_ = smem + WS_OFF_M + s_ * WS_MCHUNK;                                                                 \
        _Pragma("unroll") for (int r = 0; r < WS_G; r++)                                                                         \
            win[WIN_INDEX] = make_float2(lds32(base_ + r * vl.rstride + vl.srcA), lds32(base_ + r * vl.rstride + vl.srcB));      \
        __syncwarp();                                                                                                            \
        if (lane == 0) mbar_arrive(bars + 8 * (B_EMPTYM + s_));                                                                  \
        nM++;                                                                                                                    \
    }
            // the first four chunks of the unit; every group then pulls one more into the last 8 window rows and, when it is
            // done, moves the window down by 8 rows.  (The moves cost 64 register copies per step on the otherwise idle ALU
            // pipe; rotating the window at compile time instead -- five unrolled copies of the tap loops, 26 KB of code -- made
            // the SM's instruction cache thrash: 28 % of all warp samples were instruction-fetch stalls.)
#pragma unroll
            for (int c = 0; c < 4; c++) WS_CONSUME_CHUNK(8 * c + r)
#pragma unroll 1
            for (int g = 0; g < U.ng; g++) {
                pump();
                const unsigned ps = nP % WS_NP;
                WS_CONSUME_CHUNK(32 + r)
                wait(bars + 8 * (B_EMPTYP + ps), ((nP / WS_NP) & 1) ^ 1);
                v_compute<FMA>(win, smem + WS_OFF_P + ps * WS_PSLOT, vl, t);
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (B_FULLP + ps));
                nP++;
#pragma unroll
                for (int r = 0; r < WS_WIN - WS_G; r++) win[r] = win[r + WS_G];
            }
#undef WS_CONSUME_CHUNK
        }
        if (r_prod) { // the R streams feed the U steps that run after the last V group
            unsigned it = 0;
            while (rs.u0 < a.nunits || rs.u1 < a.nunits) {
                pump();
                __nanosleep(128);
                if (++it > (1u << 24)) __trap();
            }
        }
        return;
    }

    // ================= H + solve + U warps =================
    {
        const int hw = warp - WS_NVW, ht = tid - WS_NVW * 32;
        const int hrow = lane & 7, cbase = 16 * hw + 4 * (lane >> 3);   // H: lane = (row, 4-pixel quad)
        const int urow0 = 4 * (hw & 1), ucol = 32 * (hw >> 1) + lane;   // U: warp = 4 rows x 32 columns
        unsigned nP = 0, nR0 = 0, nR1 = 0;
        for (int u = blockIdx.x; u < a.nunits; u += gridDim.x) {
            const Unit U = decode_unit(a, u);
            const unsigned r1base = nR1;
            const bool h_active = U.x0 + 16 * hw < w;        // warp-uniform: ragged last strip
            const bool u_active = U.x0 + 32 * (hw >> 1) < w;
            for (int g = 0; g < U.ng; g++) {
                const int yg = U.y0 + WS_G * g;
                const unsigned ps = nP % WS_NP;
                float *Fb = reinterpret_cast<float *>(ws_smem + WS_OFF_F + (nP & 1) * WS_FSLOT);
                mbar_wait(bars + 8 * (B_FULLP + ps), (nP / WS_NP) & 1);
                if (h_active) {
                    const unsigned char *P = ws_smem + WS_OFF_P + ps * WS_PSLOT;
                    const float2 *p01 = reinterpret_cast<const float2 *>(P) + hrow * WS_P2 + cbase;
                    const float2 *p23 = reinterpret_cast<const float2 *>(P + WS_P23OFF) + hrow * WS_P2 + cbase;
                    const float *p4 = reinterpret_cast<const float *>(P + WS_P4OFF) + hrow * WS_P4 + cbase;
                    float fx[4], fy[4];
                    h_quad<FMA>(p01, p23, p4, t, fx, fy);
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        Fb[hrow * WS_FP + cbase + p] = fx[p];
                        Fb[(WS_G + hrow) * WS_FP + cbase + p] = fy[p];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + 8 * (B_EMPTYP + ps));
                nP++;
                hu_barrier(); // the flow tile of this group is complete; every warp has finished U of the previous group
                if (LAST) {
                    float *f = a.flow + (size_t)U.b * 2 * plane;
#pragma unroll
                    for (int i = ht; i < WS_G * WS_SW; i += WS_HU_THREADS) {
                        const int row = i / WS_SW, col = i - row * WS_SW;
                        const int x = U.x0 + col, y = yg + row;
                        if (x < w && y < h) {
                            const size_t o = (size_t)y * pitch + x;
                            f[o] = Fb[row * WS_FP + col];
                            f[o + plane] = Fb[(WS_G + row) * WS_FP + col];
                        }
                    }
                    if (a.span > 0) { // span-grid threshold test on the tile (reference src/consumer.cpp:60-77), count by warp shuffles
                        const int sx0 = (U.x0 + a.span - 1) / a.span, sy0 = (yg + a.span - 1) / a.span;
                        const int nsx = max(0, (min(U.x0 + WS_SW, w) - 1) / a.span - sx0 + 1), nsy = max(0, (min(yg + WS_G, h) - 1) / a.span - sy0 + 1);
                        int hit = 0;
                        for (int i = ht; i < nsx * nsy; i += WS_HU_THREADS) {
                            const int j = i / nsx, col = (sx0 + i - j * nsx) * a.span - U.x0, row = (sy0 + j) * a.span - yg;
                            const float dx = Fb[row * WS_FP + col], dy = Fb[(WS_G + row) * WS_FP + col];
                            const float len = (dx * dx) + (dy * dy);
                            hit += ((double)len > a.thr2) ? 1 : 0;
                        }
                        if (nsx * nsy > 0) {
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) hit += __shfl_xor_sync(0xffffffffu, hit, off);
                            if (lane == 0 && hit > 0) atomicAdd(a.counts + U.b, hit);
                        }
                    }
                    continue;
                }
                // ---- U: next update matrices (A.4) from the staged R0 / R1 ----
                const unsigned rs = nR0 % WS_NR0;
                mbar_wait(bars + 8 * (B_FULLR0 + rs), (nR0 / WS_NR0) & 1);
                if (g == 0) {
                    mbar_wait(bars + 8 * (B_FULLR1 + (r1base % WS_NR1)), (r1base / WS_NR1) & 1);
                    mbar_wait(bars + 8 * (B_FULLR1 + ((r1base + 1) % WS_NR1)), ((r1base + 1) / WS_NR1) & 1);
                }
                {
                    const unsigned n = r1base + g + 2;
                    mbar_wait(bars + 8 * (B_FULLR1 + (n % WS_NR1)), (n / WS_NR1) & 1);
                }
                if (u_active) {
                    const float *R0s = reinterpret_cast<const float *>(ws_smem + WS_OFF_R0 + rs * WS_R0SLOT) + urow0 * 5 * WS_SW + ucol; // [8][5][96]
                    const float *R1s = reinterpret_cast<const float *>(ws_smem + WS_OFF_R1);                                             // [32][5][104]
                    const float *Fu = Fb + urow0 * WS_FP + ucol;
                    float *Mo = a.Mout + (size_t)U.b * 5 * plane;
                    const int x = U.x0 + ucol, y0u = yg + urow0;
                    const int xw = U.x0 - WS_R1X;
                    const int ring0 = WS_G - U.y0 + WS_G * (int)(r1base % WS_NR1); // ring row of frame row yy = (yy + ring0) & 31
                    // positions of the thread's four vertically adjacent pixels
                    float dxs[4], dys[4], fxr[4], fyr[4];
                    int x1s[4], y1s[4];
                    bool fast = (unsigned)(x - 5) < (unsigned)(w - 10) && y0u >= 5 && y0u + 3 < h - 5;
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        dxs[q4] = Fu[q4 * WS_FP]; dys[q4] = Fu[(WS_G + q4) * WS_FP];
                        const float pxf = (float)x + dxs[q4], pyf = (float)(y0u + q4) + dys[q4];
                        x1s[q4] = __float2int_rd(pxf); y1s[q4] = __float2int_rd(pyf);
                        fxr[q4] = pxf - (float)x1s[q4]; fyr[q4] = pyf - (float)y1s[q4];
                        // inside the frame (A.4) and inside the staged R1 window
                        fast = fast && (unsigned)x1s[q4] < (unsigned)(w - 1) && (unsigned)y1s[q4] < (unsigned)(h - 1) &&
                               (unsigned)(x1s[q4] - xw) <= (unsigned)(WS_R1C - 2) && (unsigned)(y1s[q4] - yg + WS_G) <= (unsigned)(3 * WS_G - 2);
                    }
                    if (__all_sync(0xffffffffu, fast)) {
#pragma unroll
                        for (int pr = 0; pr < 4; pr += 2) { // pixels (pr, pr + 1) packed
                            float2 q[5], pt[5][2], pb[5][2], m[5];
                            const int ria = (y1s[pr] + ring0) & (WS_NR1 * WS_G - 1), rja = (ria + 1) & (WS_NR1 * WS_G - 1);
                            const int rib = (y1s[pr + 1] + ring0) & (WS_NR1 * WS_G - 1), rjb = (rib + 1) & (WS_NR1 * WS_G - 1);
                            const float *a0 = R1s + ria * WS_R1ROW + (x1s[pr] - xw), *a1 = R1s + rja * WS_R1ROW + (x1s[pr] - xw);
                            const float *b0 = R1s + rib * WS_R1ROW + (x1s[pr + 1] - xw), *b1 = R1s + rjb * WS_R1ROW + (x1s[pr + 1] - xw);
#pragma unroll
                            for (int c = 0; c < 5; c++) {
                                q[c] = make_float2(R0s[(pr * 5 + c) * WS_SW], R0s[((pr + 1) * 5 + c) * WS_SW]);
                                pt[c][0] = make_float2(a0[c * WS_R1C], b0[c * WS_R1C]); pt[c][1] = make_float2(a0[c * WS_R1C + 1], b0[c * WS_R1C + 1]);
                                pb[c][0] = make_float2(a1[c * WS_R1C], b1[c * WS_R1C]); pb[c][1] = make_float2(a1[c * WS_R1C + 1], b1[c * WS_R1C + 1]);
                            }
                            upd_core2(q, pt, pb, make_float2(fxr[pr], fxr[pr + 1]), make_float2(fyr[pr], fyr[pr + 1]), make_float2(dxs[pr], dxs[pr + 1]),
                                      make_float2(dys[pr], dys[pr + 1]), t.one, m);
                            const float ma[5] = {m[0].x, m[1].x, m[2].x, m[3].x, m[4].x}, mb[5] = {m[0].y, m[1].y, m[2].y, m[3].y, m[4].y};
                            store_M(Mo, pitch, y0u + pr, x, ma);
                            store_M(Mo, pitch, y0u + pr + 1, x, mb);
                        }
                    } else { // frame borders, ragged tiles, large motion: one pixel at a time
                        const float *R1g = a.R + ((size_t)U.b * 10 + 5) * plane;
#pragma unroll 1
                        for (int q4 = 0; q4 < 4; q4++) {
                            const int y = y0u + q4;
                            if (x >= w || y >= h) continue;
                            const float dx = Fu[q4 * WS_FP], dy = Fu[(WS_G + q4) * WS_FP];
                            const float pxf = (float)x + dx, pyf = (float)y + dy;
                            const int x1 = __float2int_rd(pxf), y1 = __float2int_rd(pyf);
                            const float fx1 = pxf - (float)x1, fy1 = pyf - (float)y1;
                            const bool inside = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
                            float q[5], pt[5][2], pb[5][2];
#pragma unroll
                            for (int c = 0; c < 5; c++) q[c] = R0s[(q4 * 5 + c) * WS_SW];
                            if (inside) {
                                if ((unsigned)(x1 - xw) <= (unsigned)(WS_R1C - 2) && (unsigned)(y1 - yg + WS_G) <= (unsigned)(3 * WS_G - 2)) {
                                    const int ri = (y1 + ring0) & (WS_NR1 * WS_G - 1), rj = (ri + 1) & (WS_NR1 * WS_G - 1);
                                    const float *a0 = R1s + ri * WS_R1ROW + (x1 - xw), *a1 = R1s + rj * WS_R1ROW + (x1 - xw);
#pragma unroll
                                    for (int c = 0; c < 5; c++) {
                                        pt[c][0] = a0[c * WS_R1C]; pt[c][1] = a0[c * WS_R1C + 1];
                                        pb[c][0] = a1[c * WS_R1C]; pb[c][1] = a1[c * WS_R1C + 1];
                                    }
                                } else { // the same values from global memory
                                    const float *p = R1g + (size_t)y1 * 5 * pitch + x1;
#pragma unroll
                                    for (int c = 0; c < 5; c++) {
                                        pt[c][0] = __ldg(p + c * pitch); pt[c][1] = __ldg(p + c * pitch + 1);
                                        pb[c][0] = __ldg(p + (5 + c) * pitch); pb[c][1] = __ldg(p + (5 + c) * pitch + 1);
                                    }
                                }
                            }
                            float m[5];
                            upd_core<false, true>(q, pt, pb, inside, fx1, fy1, w, h, x, y, dx, dy, m);
                            store_M(Mo, pitch, y, x, m);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bars + 8 * (B_EMPTYR0 + rs));
                    mbar_arrive(bars + 8 * (B_EMPTYR1 + ((r1base + g) % WS_NR1)));
                }
                nR0++;
            }
            if (!LAST) { // the two chunks below the last group were staged for its gathers only
                if (lane == 0) {
                    mbar_arrive(bars + 8 * (B_EMPTYR1 + ((r1base + U.ng) % WS_NR1)));
                    mbar_arrive(bars + 8 * (B_EMPTYR1 + ((r1base + U.ng + 1) % WS_NR1)));
                }
                nR1 = r1base + U.ng + 2;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool encode3(CUtensorMap *m, const float *base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1_bytes, cuuint64_t s2_bytes, cuuint32_t b0,
             cuuint32_t b1)
{
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    const cuuint64_t dims[3] = {d0, d1, d2}, strides[2] = {s1_bytes, s2_bytes};
    const cuuint32_t box[3] = {b0, b1, 1}, estr[3] = {1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count()
{
    static int n[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    if (!n[d]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v < 1) v = 148;
        n[d] = v;
    }
    return n[d];
}

template <int FMA, bool LAST>
cudaError_t launch_strip(cudaStream_t s, const StripMaps &m, const StripArgs &sa, const WinTaps &t, int grid)
{
    static bool configured_dev[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) d = 0;
    if (!configured_dev[d]) {
        cudaError_t e = cudaFuncSetAttribute(gauss_strip_kernel<FMA, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
        if (e != cudaSuccess) return e;
        configured_dev[d] = true;
    }
    const CUtensorMap *maps = reinterpret_cast<const CUtensorMap *>(m.opaque);
    gauss_strip_kernel<FMA, LAST><<<grid, WS_THREADS, WS_SMEM, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], sa, t);
    return cudaGetLastError();
}

} // namespace

static_assert(sizeof(CUtensorMap) * 6 <= sizeof(StripMaps::opaque), "StripMaps holds six tensor maps");

// Tensor maps of one (M buffer, R buffer) pair at one scale: M as [B][h][5*pitch] floats (boxes 256 x {1, 8} and 128 x {1, 8}: rows
// of a channel-pair plane / of the h2 plane), R as [2B][5h][pitch] floats (boxes 96 x 40 and 104 x 40: 8 rows x 5 channels).
bool make_strip_maps(const float *M, const float *R, const LevelDims &d, int batch, StripMaps *out)
{
    CUtensorMap *maps = reinterpret_cast<CUtensorMap *>(out->opaque);
    const cuuint64_t pitch = (cuuint64_t)d.pitch, h = (cuuint64_t)d.h, plane = (cuuint64_t)d.plane;
    out->valid = 0;
    if (d.pitch % 4 != 0 || 5 * pitch < 2 * WS_VC) return false;
    if (!encode3(&maps[0], M, 5 * pitch, h, (cuuint64_t)batch, 5 * pitch * 4, 5 * plane * 4, 2 * WS_VC, 1)) return false;
    if (!encode3(&maps[1], M, 5 * pitch, h, (cuuint64_t)batch, 5 * pitch * 4, 5 * plane * 4, WS_VC, 1)) return false;
    if (!encode3(&maps[2], M, 5 * pitch, h, (cuuint64_t)batch, 5 * pitch * 4, 5 * plane * 4, 2 * WS_VC, WS_G)) return false;
    if (!encode3(&maps[3], M, 5 * pitch, h, (cuuint64_t)batch, 5 * pitch * 4, 5 * plane * 4, WS_VC, WS_G)) return false;
    if (!encode3(&maps[4], R, pitch, 5 * h, (cuuint64_t)batch * 2, pitch * 4, 5 * plane * 4, WS_SW, 5 * WS_G)) return false;
    if (!encode3(&maps[5], R, pitch, 5 * h, (cuuint64_t)batch * 2, pitch * 4, 5 * plane * 4, WS_R1C, 5 * WS_G)) return false;
    out->valid = 1;
    return true;
}

bool gauss_strip_ok(const IterArgs &a, const WinTaps &t)
{
    return t.m == WS_MR && !a.ufma && (a.fma == 0 || a.fma == 2) && !a.scalar && a.d.w >= 1 && a.d.h >= 1;
}

cudaError_t launch_gauss_strip(cudaStream_t s, const IterArgs &a, const WinTaps &t, const StripMaps &m)
{
    if (!m.valid || !gauss_strip_ok(a, t)) return cudaErrorInvalidValue;
    StripArgs sa{};
    sa.R = a.R; sa.Mout = a.Mout; sa.flow = a.flow; sa.d = a.d;
    sa.span = a.last ? a.span : 0; sa.thr2 = a.thr2; sa.counts = a.counts;
    const int sms = sm_count();
    sa.nsx = (a.d.w + WS_SW - 1) / WS_SW;
    // vertical segments: about 8 units per SM (tail of the last round), at least 64 rows each (the 40-row window refill of a unit)
    const int per = a.batch * sa.nsx;
    int nsy = (8 * sms + per - 1) / per;
    nsy = std::max(1, std::min(nsy, a.d.h / 64));
    sa.segh = ((a.d.h + nsy - 1) / nsy + WS_G - 1) & ~(WS_G - 1);
    sa.nsy = (a.d.h + sa.segh - 1) / sa.segh;
    sa.nunits = per * sa.nsy;
    const int grid = std::min(sa.nunits, sms);
    if (a.last) return a.fma == 2 ? launch_strip<2, true>(s, m, sa, t, grid) : launch_strip<0, true>(s, m, sa, t, grid);
    return a.fma == 2 ? launch_strip<2, false>(s, m, sa, t, grid) : launch_strip<0, false>(s, m, sa, t, grid);
}

} // namespace tw
