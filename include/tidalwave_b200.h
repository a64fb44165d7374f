/*
 * tidalwave_b200.h -- C ABI of the B200-native hot path of arielnetworks/tidal-wave:
 * dense Farneback optical flow between an expected/target screenshot pair, span-strided
 * vector sampling and threshold classification into OK / SUSPICIOUS / ERROR.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * reference tree).  Plain pointers and sizes only; no exceptions cross this boundary
 * ("node-gyp can't use exceptions", src/opticalflow.h:16-19) -- errors are values.
 *
 * There is NO CPU fallback: every compute entry point fails with TW_CUDA_ERROR if no
 * sm_100-class device / kernel image is available.
 */
#ifndef TIDALWAVE_B200_H
#define TIDALWAVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ErrorCode, src/opticalflow.h:9-14 (values 0..3 identical) + two additions. */
enum tw_error_code {
    TW_OK = 0,
    TW_BAD_PARAMETER = 1,
    TW_BAD_IMAGE_FORMAT = 2,
    TW_DONT_MATCH_SIZE = 3,
    TW_CUDA_ERROR = 4,   /* reason carries cudaGetErrorString */
    TW_UNSUPPORTED = 5   /* a documented gap (removed experimental options) */
};

/* Response.status, src/consumer.cpp:77,86 */
enum tw_status { TW_STATUS_OK = 0, TW_STATUS_SUSPICIOUS = 1, TW_STATUS_ERROR = 2 };

/* OpticalFlowParameter, src/opticalflow.h:28-36 -- field for field. */
typedef struct tw_flow_param {
    double pyrScale;
    int pyrLevels;
    int winSize;
    int pyrIterations;
    int polyN;
    double polySigma;
    int flags; /* 256 = OPTFLOW_FARNEBACK_GAUSSIAN, 0 = box window; anything else -> TW_BAD_PARAMETER */
} tw_flow_param;

/* Vector, src/message_queue.h:20-25 */
typedef struct tw_vector {
    int x;
    int y;
    double dx;
    double dy;
} tw_vector;

/* The value part of Response (src/message_queue.h:27-42) + OpticalFlowStatus (src/opticalflow.h:20-26). */
typedef struct tw_result {
    int code;       /* tw_error_code */
    int status;     /* tw_status */
    int n_vectors;  /* number of vectors that passed the threshold (may exceed the caller's cap) */
    int width;      /* of the expected image, src/opticalflow.cpp:72-73 */
    int height;
    float time;     /* seconds of device compute for this pair (src/opticalflow.cpp:82-93,112-118) */
    char reason[128];
} tw_result;

typedef struct tw_ctx tw_ctx;
typedef struct tw_pool tw_pool;

/* Fills *p with the defaults of Broker::createInstance (src/broker.cpp:106-117):
 * pyrScale .5, pyrLevels 3, winSize 30, pyrIterations 3, polyN 7, polySigma 1.5, flags 256. */
void tw_default_param(tw_flow_param *p);

/* Library / build information: "tidalwave_b200 <version> sm_100a". */
const char *tw_version(void);

/* Number of CUDA devices visible (cv::gpu::getCudaEnabledDeviceCount, src/consumer.cpp:19-20). */
int tw_device_count(void);

/* cv::imread(path, IMREAD_GRAYSCALE) on an in-memory file (src/opticalflow.cpp:37,44): PNG (every colour type and bit depth, Adam7
 * or not, critical-chunk CRCs checked; colour -> gray exactly as OpenCV/libpng: (9797 R + 19234 G + 3737 B) >> 15), JPEG (baseline / progressive Huffman, gray or YCbCr: the
 * luma plane through libjpeg's ISLOW inverse DCT, as OpenCV's JCS_GRAYSCALE request does; EXIF orientation ignored like
 * OpenCV 2.4.9) and binary PGM.  Host code, no device needed.  Call with out == NULL to query *w, *h.  Anything else ->
 * TW_BAD_IMAGE_FORMAT, reported by callers as "Can't open <path>". */
int tw_decode_gray(const uint8_t *bytes, size_t n, uint8_t *out, size_t cap, int *w, int *h);

/* ---- operator seam: OpticalFlow instance, one per consumer thread (src/consumer.cpp:27-35) ----
 * A context is thread-confined, bound to one device (cv::gpu::setDevice(id), src/consumer.cpp:22),
 * owns all device memory, pinned staging and its stream.  max_batch = pairs processed per launch
 * sequence by the batch entry points (>= 1).  Returns NULL and fills err on failure. */
tw_ctx *tw_create(int device, int max_w, int max_h, int max_batch, char *err, int errlen);
void tw_destroy(tw_ctx *ctx);

/* OpticalFlow::calculateInternal (src/opticalflow.h:49, src/opticalflow.cpp:78-119):
 * equal-size 8-bit gray images in, two float planes out (either may be NULL), *seconds = device
 * compute time (uploads/downloads excluded, as in src/opticalflow.cpp:112-116).  flowx/flowy are
 * w*h floats, row-major, tightly packed.  Returns a tw_error_code. */
int tw_flow(tw_ctx *ctx, const uint8_t *expect, const uint8_t *target, int w, int h, int stride,
            const tw_flow_param *param, float *flowx, float *flowy, float *seconds);

/* OpticalFlow::calculate (src/opticalflow.cpp:20-76) from decoded images + the sampling loop of
 * Consumer::run (src/consumer.cpp:59-88).  Size rule of src/opticalflow.cpp:52-68: |dw|>5 or |dh|>5 ->
 * TW_DONT_MATCH_SIZE; unequal within 5 px -> target bilinearly resized to the expected size.
 * Vectors are written in row-major scan order (y outer, x inner); at most cap are stored,
 * res->n_vectors is the full count.  NULL / empty images -> TW_BAD_IMAGE_FORMAT.  Returns res->code. */
int tw_compare(tw_ctx *ctx, const uint8_t *expect, int ew, int eh, const uint8_t *target, int tw, int th,
               const tw_flow_param *param, double threshold, int span, tw_vector *out, int cap,
               tw_result *res);

/* The resize step of OpticalFlow::calculate alone (src/opticalflow.cpp:64-68, cv::resize INTER_LINEAR on 8-bit data), for
 * hosts that keep calculate() / calculateInternal() separate: target (tw x th) -> out (ew x eh), tightly packed. */
int tw_resize_target(tw_ctx *ctx, const uint8_t *target, int tw, int th, uint8_t *out, int ew, int eh);

/* n same-size pairs through one batched launch sequence (n <= max_batch).  out has n*cap entries
 * (pair i at out + i*cap), res has n entries.  Semantics per pair identical to tw_compare. */
int tw_compare_batch(tw_ctx *ctx, int n, const uint8_t *const *expect, const uint8_t *const *target, int w,
                     int h, int stride, const tw_flow_param *param, double threshold, int span,
                     tw_vector *out, int cap, tw_result *res);

/* ---- pipelined form of tw_compare_batch (what a dispatcher consumer uses: ONE host thread keeps its GPU busy, as the
 * reference's Consumer does with its device, src/consumer.cpp:18-24, 42-94).  tw_pipe_submit enqueues a batch and returns:
 * the host -> device copies run on a copy stream through a staging buffer (overlapping the previous batch's kernels), the
 * compute stream takes them over, runs the launch sequence and snapshots the compact results.  tw_pipe_collect waits for the
 * OLDEST submitted batch and fills res[0..n) / out exactly as tw_compare_batch does (n = that batch's size).  At most two batches
 * are in flight (tw_pipe_pending); the images of a submitted batch must stay valid until it has been collected; a change of
 * size or parameters needs an empty pipe (TW_BAD_PARAMETER otherwise).  Not to be interleaved with the synchronous entry
 * points while batches are pending. */
int tw_pipe_submit(tw_ctx *ctx, int n, const uint8_t *const *expect, const uint8_t *const *target, int w, int h, int stride,
                   const tw_flow_param *param, double threshold, int span);
int tw_pipe_collect(tw_ctx *ctx, tw_vector *out, int cap, tw_result *res);
int tw_pipe_pending(tw_ctx *ctx);
/* 1 if the oldest pending batch has finished on the device (tw_pipe_collect would not block), else 0. */
int tw_pipe_ready(tw_ctx *ctx);

/* ---- split phases of tw_compare_batch (same stream; used by the dispatcher to overlap, and by the
 * benchmark to time the device-resident pass separately from the PCIe legs) ---- */
int tw_batch_upload(tw_ctx *ctx, int n, const uint8_t *const *expect, const uint8_t *const *target, int w,
                    int h, int stride);                                     /* async H2D into the context */
int tw_batch_run(tw_ctx *ctx, int n, int w, int h, const tw_flow_param *param, double threshold,
                 int span);                                                  /* enqueue all kernels, no sync */
int tw_batch_fetch(tw_ctx *ctx, int n, tw_vector *out, int cap, tw_result *res); /* D2H compact results + sync */
int tw_sync(tw_ctx *ctx);
/* Copies the dense flow planes of pair `pair` of the last run back (either pointer may be NULL); TW_BAD_PARAMETER if that run
 * was classification-only ("sparse_last"). */
int tw_batch_flow(tw_ctx *ctx, int pair, float *flowx, float *flowy);
const char *tw_last_error(tw_ctx *ctx);
/* Per-context options.
 *   "arithmetic" = 0: every kernel keeps the oracle's operation order: results bit-identical to
 *                     oracle/farneback_ref.c on every configuration.
 *                = 1 (default; TW_ARITHMETIC=faithful in the environment or tw_set_default_arithmetic(0) flips it):
 *                     relaxed arithmetic where it was validated -- Gaussian window (flags 256) with winSize >= 30 and
 *                     polyN 7, i.e. the reference's default option family: direct-form fmaf window taps and a mixed
 *                     double / float horizontal pass in the polynomial expansion (restated bit for bit in the oracle:
 *                     twref_set_relax(144)).  Measured against the faithful oracle and cv2: <= 1.5e-4 px at 1920x1080,
 *                     2.3e-3 px max / 1.3e-4 px RMS on the reference's fixture, identical status and vector sets (bar:
 *                     1e-2 px max, 1e-3 px RMS; tools/relax_cases.py, tools/parity_report.py).  Every other option set
 *                     (box window, smaller windows, polyN != 7) runs the faithful kernels.
 *   "sparse_last" = 1: classification only -- the last iteration of the finest scale evaluates the window blur and the solve
 *                     only at the positions src/consumer.cpp:60-77 samples (Gaussian window of radius 15 / 7, span > 0).  Status
 *                     and vectors (positions and the float dx, dy) are bit-identical to the dense path; the dense flow field
 *                     is not produced (tw_batch_flow fails after such a run; tw_flow is always dense).  Default 0 for
 *                     tw_create, 1 for the dispatcher's contexts (tw_pool_*; TW_SPARSE_LAST=0 turns it off there).
 *   "graph"      = 1 (default): repeated runs of one (size, batch, options) replay a captured CUDA graph.
 *   "gauss_fma", "update_fma", "gauss_scalar": removed in round 2 (studied and rejected relaxations -- oracle relax bits 0 / 6 --
 *                     and the scalar v1 window kernel); setting one to 1 answers TW_UNSUPPORTED.
 *   "window_tiles" = 1 (default) / 0: the Gaussian window iterations of radius 15 run the tile-per-CTA kernel / the persistent
 *                     warp-specialised strip kernel (tw_window.cu: TMA tensor-map rings, register-resident column walkers,
 *                     setmaxnreg).  Same arithmetic, bit-identical results; the strip kernel measures 1-7 % slower (DESIGN.md
 *                     section 4.2), so it is opt-in; TW_WINDOW=strip|tiles in the environment sets the default.
 *   "plan_cache" = n (default 3; TW_PLAN_CACHE): plans -- the device buffers, tables, tensor maps and the captured graph of one
 *                     (size, parameters) -- kept besides the current one, least recently used evicted first, all dropped when
 *                     memory runs out: a dispatcher that sees mixed page sizes does not rebuild a plan on every switch.
 *   "polyexp_tma" = 1: interior input tiles of the relaxed polynomial expansion are staged by one TMA tensor-map copy instead of
 *                     per-thread loads (bit-identical; measured 5 % slower, hence opt-in; TW_POLY_TMA=1 sets the default).
 *   "level_generic", "level_unfused", "tight_pitch", "box_unfused": alternative code paths (also the fall-backs of unusual
 *                     options) that the parity tests force. */
int tw_set_option(tw_ctx *ctx, const char *name, int value);
/* Process-wide default of "arithmetic" for contexts created afterwards (the dispatcher's consumers included). */
int tw_set_default_arithmetic(int relaxed);
/* 1 if `param` would run the relaxed kernels on this context, 0 if the faithful ones, <0 on bad arguments. */
int tw_arithmetic_in_effect(tw_ctx *ctx, const tw_flow_param *param);

/* Pinned (page-locked) host memory for image buffers: H2D copies from it are true async DMA. */
void *tw_host_alloc(size_t bytes);
void tw_host_free(void *p);

/* ---- measurement hooks (CUDA events on the context's own stream) ---- */
/* Enqueues a write of a >L2 (256 MB) scratch buffer on the context's stream, evicting L2 between timed steps. */
int tw_l2_flush(tw_ctx *ctx);
int tw_timer_start(tw_ctx *ctx);
int tw_timer_stop(tw_ctx *ctx, float *ms);       /* records, synchronises, returns elapsed ms */
/* Per-kernel-family event timing: enable, run, then read n families back.  names[i] points to a static
 * string; ms[i] is the summed device time, launches[i] the launch count and alg_bytes[i] the ALGORITHMIC HBM
 * bytes (the family's term of B_alg) since enable/reset.  Returns the family count. */
int tw_profile_enable(tw_ctx *ctx, int on);
int tw_profile_read(tw_ctx *ctx, int max_n, const char **names, float *ms, int *launches, double *alg_bytes);
/* Number of kernel launches issued by this context since creation. */
long long tw_launch_count(tw_ctx *ctx);

/* ---- per-stage read-back for parity tests (planar float, c*h*w): name in
 * {"I","R0","R1","M","flow"}, scale = index in the coarse->fine schedule, pair < n of the last run.
 * Returns channel count written, <0 on error; *w,*h receive the plane size.  "M" holds whatever the last
 * kernel that wrote it left (run with pyrIterations = k to observe iteration k-1's input). ---- */
int tw_debug_read(tw_ctx *ctx, const char *name, int scale, int pair, float *out, int cap_floats, int *w, int *h);
int tw_debug_keep_levels(tw_ctx *ctx, int on); /* keep every scale's I/R/M (separate buffers) for read-back */

/* ---- dispatcher: Manager + Consumer pool (src/manager.cpp:40-98, src/consumer.cpp:12-94) ----
 * One host thread per entry of devices[] (consumer i binds GPU devices[i]), all blocking on one shared
 * request queue; options are fixed per pool (src/manager.cpp:72-73).  Images are caller-owned and must
 * stay valid until the result for that id has been taken.  No CPU fallback worker.
 * vector_cap > 0: at most that many vectors are kept per request (tw_result.n_vectors still holds the full count);
 * vector_cap == 0: all of them, whatever the image size (capacity = the request's own sampling grid).
 * max_w / max_h > 0 bound the image size a consumer accepts (larger requests answer TW_BAD_PARAMETER); 0 = no bound. */
tw_pool *tw_pool_create(const int *devices, int n_devices, int max_w, int max_h, int batch,
                        const tw_flow_param *param, double threshold, int span, int vector_cap,
                        char *err, int errlen);
/* Manager::request (src/manager.cpp:68-78): returns the request id (>= 0) or <0 if the pool is stopped. */
long long tw_pool_submit(tw_pool *pool, const uint8_t *expect, int ew, int eh, const uint8_t *target, int tw,
                         int th);
/* Manager::request exactly as the reference has it (src/manager.cpp:68-78, Request = two paths, src/message_queue.h:13-18): the
 * files are read and decoded -- cv::imread(IMREAD_GRAYSCALE), src/opticalflow.cpp:37,44 -> tw_decode_gray -- on the pool's decoder
 * threads (started on first use: one per host core, at most 32; tw_pool_set_decoders before the first call changes that), then join
 * the same request queue.  Empty path -> TW_BAD_PARAMETER, unreadable / undecodable file -> TW_BAD_IMAGE_FORMAT "Can't open <path>"
 * (src/opticalflow.cpp:20-49), sizes more than 5 px apart -> TW_DONT_MATCH_SIZE.  The pool owns the decoded images. */
long long tw_pool_submit_files(tw_pool *pool, const char *expect_path, const char *target_path);
int tw_pool_set_decoders(tw_pool *pool, int n);
/* Blocks until request `id` is answered; copies up to cap vectors.  Returns res->code, or <0 if dropped. */
int tw_pool_wait(tw_pool *pool, long long id, tw_vector *out, int cap, tw_result *res);
/* Blocks until request `id` is answered and copies its result WITHOUT taking it (res->n_vectors = the room tw_pool_wait needs).
 * Returns res->code, or <0 if dropped / unknown. */
int tw_pool_peek(tw_pool *pool, long long id, tw_result *res);
/* Non-blocking: 1 if answered (result copied), 0 if pending, <0 if dropped/unknown. */
int tw_pool_poll(tw_pool *pool, long long id, tw_vector *out, int cap, tw_result *res);
/* Report, src/message_queue.h:44-48: {request, data, error}. */
void tw_pool_report(tw_pool *pool, int *request_count, int *data_count, int *error_count);
/* Manager::stop (src/manager.cpp:63-66,93-97): drops pending requests, joins the consumers. */
void tw_pool_stop(tw_pool *pool);
void tw_pool_destroy(tw_pool *pool);

#ifdef __cplusplus
}
#endif
#endif /* TIDALWAVE_B200_H */
