// tw_kernels.cuh -- launch wrappers of the sm_100a kernel families.
// Device layout: every tensor is PLANAR float32 with a row pitch that is a multiple of 32 floats (128 B):
//   I    [B][2][h][pitch]        level images (expected, target)
//   R    [B][2][5][h][pitch]     polynomial expansion coefficients (R0 = expected, R1 = target)
//   M    [B][5][h][pitch]        (G11, G12, G22, h1, h2)
//   flow [B][2][h][pitch]        (dx, dy)
// A "plane" is h*pitch floats.  Batches are the outermost dimension; kernels index them with blockIdx.z.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tw {

constexpr int kMaxPolyN = 15;    // radius of the polynomial expansion window
constexpr int kMaxWinRadius = 32; // winSize/2

struct PolyTables {
    float g[kMaxPolyN + 1], xg[kMaxPolyN + 1], xxg[kMaxPolyN + 1];
    double gd[kMaxPolyN + 1], xxgd[kMaxPolyN + 1]; // (double)g[k], (double)xxg[k]: no conversions in the tap loop
    double ig11, ig03, ig33, ig55;
    float fig11, fig55; // (float)ig11, (float)ig55: relaxed-arithmetic kernel
    float one;          // 1.0f, deliberately a runtime value (packed f32x2 accumulate: see tw_fma2 in tw_device.cuh)
    int n;
};

struct WinTaps {
    float k[kMaxWinRadius + 1];
    float one; // 1.0f, deliberately a runtime value (see tw_fma2 in tw_kernels.cu)
    int m;
};

struct LevelDims {
    int w, h, pitch;
    size_t plane; // h * pitch
};

// K1: u8 full-res frames -> float level image (pre-blur REFLECT_101 + bilinear downsample), SURVEY App. A.2.
struct LevelImageArgs {
    const uint8_t *src; // [nimg][H][spitch]
    int W, H, spitch;
    float *dst;         // [nimg][h][pitch]
    LevelDims d;
    const int *xi; const float *xf; const int *yi; const float *yf; // device tables, A.2b
    const float *taps;  // device, ksize floats
    int ksize;
    int nimg;
    int tile_w, tile_h; // output tile
    int smem_w, smem_h; // source tile bound (before clamping)
    int identity;       // w == W && h == H
    int small;          // ksize/2 >= min(W, H): multi-bounce reflect path
};
cudaError_t launch_level_image(cudaStream_t s, const LevelImageArgs &a);
// Fast paths (full-resolution level; exact integer down-scales 2/4/8/16).  int_scale = S if W == w*S, H == h*S and the
// resize tables are exactly (S*d + S/2 - 1, 0.5), else 0.  Returns cudaErrorNotSupported if none applies.
cudaError_t launch_level_image_fast(cudaStream_t s, const LevelImageArgs &a, const float *host_taps, int int_scale);
// Long pre-blur kernels (ksize >= 31, not an exact integer down-scale): separable blur through an intermediate T in global memory
// ([nimg][H][roundup(2w, 32)] floats, level_image_big_floats): two small launches instead of huge shared-memory tiles.
bool level_image_big_ok(const LevelImageArgs &a);
size_t level_image_big_floats(int H, int w, int nimg);
cudaError_t launch_level_image_big(cudaStream_t s, const LevelImageArgs &a, float *T);
// All four level images of the default pyramid (S = 8, 4, 2, 1 with 19/9/3/3-tap pre-blurs) from one staged source tile.
cudaError_t launch_level_fused(cudaStream_t s, const uint8_t *src, int W, int H, int spitch, float *const dst[4], const LevelDims d[4],
                               const float *k8, const float *k4, const float *k2, const float *k1, int nimg);
size_t level_image_smem_bytes(int smem_w, int smem_h, int tile_w, int ksize, int identity);

// K2: polynomial expansion I -> R (5 planes), SURVEY App. A.3.  nimg = 2*B images.
// relaxed = 1: the mixed double/float horizontal pass (polyN 5 / 7 only; otherwise the faithful kernels run).
// `imap` (optional): tensor map of I made by make_polyexp_map -- interior tiles of the relaxed kernel are then staged by ONE TMA tile copy
// (cp.async.bulk.tensor) instead of per-thread loads; border tiles (replicate clamp) keep the per-thread path.
struct TileMap {
    alignas(64) unsigned char opaque[128]; // one CUtensorMap
    int valid;
};
bool make_polyexp_map(const float *I, const LevelDims &d, int nimg, int polyN, TileMap *out);
cudaError_t launch_polyexp(cudaStream_t s, const float *I, float *R, const LevelDims &d, int nimg, const PolyTables &t, int relaxed,
                           const TileMap *imap = nullptr);

// K3: (coarse flow -> bilinear upsample * 1/pyrScale | zero) -> first update-matrices, App. A.1 + A.4.
struct FirstUpdateArgs {
    const float *flow_in; // [B][2] planes of THIS scale (box window path), or nullptr
    const float *coarse; // [B][2][ch][cpitch] or nullptr (coarsest scale: zero flow)
    LevelDims cd;
    const int *xi; const float *xf; const int *yi; const float *yf; // upsample tables
    float inv_scale;
    const float *R;  // [B][2][5] planes
    float *M;        // [B][5] planes (nullptr: skip)
    float *flow_out; // [B][2] planes (nullptr: skip; used when pyrIterations == 0)
    LevelDims d;
    int batch;
    int ufma; // validated relaxation: fmaf chains in the update-matrices arithmetic (oracle relax bit 6)
};
cudaError_t launch_first_update(cudaStream_t s, const FirstUpdateArgs &a);

// K4/K5: window blur of M + 2x2 solve, then (not last) next update-matrices or (last) flow write.
struct IterArgs {
    const float *Min; // [B][5] planes
    float *Mout;      // [B][5] planes (not last)
    const float *R;   // [B][2][5] planes
    float *flow;      // [B][2] planes (last)
    LevelDims d;
    int batch;
    int last;
    // span sampling fused into the last iteration of the finest scale (reference src/consumer.cpp:60-77): the
    // threshold test runs on the flow tile while it is still in shared memory; per-pair counts are reduced with
    // warp shuffles and one atomicAdd per warp.  span == 0: off.
    int span;
    double thr2;
    int *counts; // [B], zeroed by the host before the launch
    int scalar; // 1: scalar FP32 tap sums (v1 kernel) instead of packed f32x2
    int fma; // Gaussian tap sums: 0 = oracle order (add, mul, add); validated relaxations: 1 = fmaf(a + b, k, t) (oracle relax
             // bit 0), 2 = direct form fmaf(a, k, t), fmaf(b, k, t) (bit 7).  Never set for the box window.
    int ufma; // validated relaxation: fmaf chains in the update-matrices arithmetic (oracle relax bit 6)
};
cudaError_t launch_gauss_iter(cudaStream_t s, const IterArgs &a, const WinTaps &t);
// K4/K5 for window radius 15: the persistent, warp-specialised strip kernel (tw_window.cu).  Its operands arrive through
// TMA tensor maps (cuTensorMapEncodeTiled) built once per (scale, M buffer): make_strip_maps.
struct StripMaps {
    alignas(64) unsigned char opaque[6 * 128]; // six CUtensorMap: M channel-pair / h2 rows (1-row and 8-row boxes), R0 tiles, R1 row chunks
    int valid;
};
bool make_strip_maps(const float *M, const float *R, const LevelDims &d, int batch, StripMaps *out);
bool gauss_strip_ok(const IterArgs &a, const WinTaps &t);
cudaError_t launch_gauss_strip(cudaStream_t s, const IterArgs &a, const WinTaps &t, const StripMaps &m);
// K5s: last iteration of the finest scale evaluated only at the span-grid points (needs a.span, a.thr2, a.counts; writes the
// sampled (dx, dy) at their positions of a.flow and nothing else).  gauss_last_sparse_ok: whether this geometry is supported.
bool gauss_last_sparse_ok(const IterArgs &a, const WinTaps &t);
cudaError_t launch_gauss_last_sparse(cudaStream_t s, const IterArgs &a, const WinTaps &t);
// Box window (flags == 0), App. A.6: vertical float-difference running sums in double (VT = V transposed,
// [B*5][w][roundup(h,32)] doubles), then the horizontal running sum + solve -> flow; the next update-matrices
// is launch_first_update with flow_in.
cudaError_t launch_box_vsum(cudaStream_t s, const float *Min, double *VT, const LevelDims &d, int batch, int m);
cudaError_t launch_box_hscan(cudaStream_t s, const double *VT, float *flow, const LevelDims &d, int batch, int m, int winSize);
// Fused form (m <= 16): launch_box_ckpt keeps only the vertical running sums at the top of every band of box_band_height() rows (CK: [B][bands]
// [5 * pitch] doubles); launch_box_band re-runs them inside each band from its checkpoint, continues the horizontal running sums
// across the band and goes straight on to the solve and the next update-matrices (or the flow store): no V plane in HBM.
bool box_fused_ok(const LevelDims &d, int m);
int box_band_height(); // rows per band: CK holds ceil(h / box_band_height()) checkpoints per pair
cudaError_t launch_box_ckpt(cudaStream_t s, const float *Min, double *CK, const LevelDims &d, int batch, int m);
cudaError_t launch_box_band(cudaStream_t s, const float *Min, const double *CK, const float *R, float *Mout, float *flow, const LevelDims &d,
                            int batch, int m, int winSize, int last);

// +-5 px size-tolerance path: OpenCV's 8-bit fixed-point bilinear resize of the target to the expected size.
struct ResizeU8Args {
    const uint8_t *src; int W, H, spitch;   // device, target as uploaded
    uint8_t *dst; int w, h, dpitch;         // device, target slot of the pair
    const int *xi, *yi;                     // device tables
    const short *xa, *ya;                   // device, 2 coefficients per output column / row (x2048)
};
cudaError_t launch_resize_u8(cudaStream_t s, const ResizeU8Args &a);

// Span sampling + threshold classification + ordered compaction, reference src/consumer.cpp:60-77.
struct SampleArgs {
    const float *flow; // [B][2] planes, full-res
    LevelDims d;
    int batch;
    int span;
    double thr2; // threshold * threshold in double
    int *counts;       // [B]
    int counted;       // counts[] already hold the totals (fused into K5): pairs with 0 vectors exit at once
    void *vectors;     // [B][cap] tw_vector {int x, int y, double dx, double dy}
    int cap;
};
cudaError_t launch_sample(cudaStream_t s, const SampleArgs &a);

} // namespace tw
