/* b2resample.h -- C ABI of the B200-native FFT-upscale hot path (libb2resample.so).
 *
 * Drop-in boundary for the calls DTolm/VkResample's launchResample() makes between "planar host
 * buffer ready" and "planar host result ready" (reference file:line cited per entry point; paths
 * are relative to the reference checkout).  The reference has no FFI of its own -- it is a single
 * translation unit -- so these are the functions a maintainer would bind in place of the Vulkan
 * plumbing + VkFFT dispatch in VkResample.cpp (see INTEGRATION.md for the patch and for the
 * ctypes / cgo style stubs).
 *
 * Conventions (mirroring VkResample.cpp:1282-1780):
 *   - plain pointers and sizes only; the caller owns host buffers, the library owns device memory;
 *   - a plan is bound to one CUDA device and must be used from one thread at a time; distinct
 *     plans may be used concurrently from different threads and on different GPUs (the reference
 *     creates one private VkDevice + plan per std::thread, VkResample.cpp:1282, :1961-1965);
 *   - every function returning int returns B2R_SUCCESS (0) or a negative error code and sets a
 *     thread-local message readable through b2r_last_error() (the reference returns VkResult and
 *     printf()s, VkResample.cpp:1286-1320);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     B2R_ERR_CUDA.
 *
 * Data layouts (elements are float for precision 0, double for precision 1, IEEE binary16 for precision 2):
 *   input   planar [3][H][W], row stride W, plane stride (W+2)*H elements -- exactly the buffer
 *           launchResample fills at VkResample.cpp:1636-1685 (2*H unused pad elements per plane).
 *           Size = b2r_plan_input_bytes() = 3 * elem * (W+2) * H  (== 3*complexSize*(W/2+1)*H, :1437).
 *   output  planar [3][upH][upW] compact, plane stride upW*upH -- what transferDataToCPU returns
 *           at VkResample.cpp:1697-1700.  Size = b2r_plan_output_bytes() = 3 * elem * upW * upH.
 */
#ifndef B2RESAMPLE_H_
#define B2RESAMPLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_SUCCESS 0
#define B2R_ERR_INVALID_ARG (-1)   /* bad size / factor / precision / null pointer            */
#define B2R_ERR_CUDA (-2)          /* CUDA runtime error or no device (message has the detail) */
#define B2R_ERR_UNSUPPORTED (-3)   /* e.g. size not 2^a 3^b 5^c 7^d, precision 1 without NVRTC  */
#define B2R_ERR_NOMEM (-4)

#define B2R_PRECISION_FP32 0u      /* -p 0, VkResample.cpp:1860-1866 */
#define B2R_PRECISION_FP64 1u      /* -p 1: double storage and arithmetic; kernels compiled at plan time */
#define B2R_PRECISION_FP16 2u      /* -p 2: half storage, fp32 FFT arithmetic, half sharpen     */

/* flags for b2r_plan_create */
#define B2R_FLAG_NONE 0u
#define B2R_FLAG_NO_GRAPH 1u       /* launch kernels directly instead of replaying a CUDA graph */
#define B2R_FLAG_NO_SHARPEN_LITERAL_ROUNDING 2u /* use the raw floats instead of their "%f" text */
/* Reproduce the reference's C2C branch (performR2C == false, VkResample.cpp:1423-1424: what it runs
 * when upW > maxComputeSharedMemorySize/8, e.g. upW > 6144 on NVIDIA Vulkan): both Nyquist lines on the
 * negative side only (:527-546), complex result, sharpen on length(vec2) (:884-904), compact plane
 * below the sharpen (:1598).  Default (flag clear) is R2C/C2R semantics at every size. */
#define B2R_FLAG_C2C_PARITY 4u
/* Plan-time JIT (default ON): sizes without an ahead-of-time schedule get statically scheduled kernels
 * compiled with NVRTC when the plan is created (about a second; the cubin is cached on disk) -- the
 * counterpart of the reference JIT-compiling GLSL for every plan (vkFFT.h:7446-7521).  If NVRTC or the
 * driver library cannot be loaded the plan silently uses the any-size kernels (about half the speed;
 * b2r_plan_info.jit_note says why).  B2R_FLAG_NO_JIT (or B2R_JIT=0 in the environment) selects the
 * any-size kernels on purpose.  B2R_FLAG_JIT is accepted for compatibility and has no effect. */
#define B2R_FLAG_JIT 8u
#define B2R_FLAG_NO_JIT 16u
/* Sharpen arithmetic.  DEFAULT (since round 2): the tolerance-bound kernels (csrc/b2r_cas.cuh) -- the
 * reference's formula (VkResample.cpp:909-922) with one quotient instead of two (min(a,b) taken before the
 * monotone map z/(2-z)), rsqrt / rcp hardware approximations and FMA contraction; within 1e-5 (fp32) /
 * 1e-2 (fp16) of oracle.sharpen on the identical plane (north_star's bar; measured ~1e-6 / 4e-3), which is
 * also the precision class of the reference's own GLSL `/` and sqrt() (Vulkan allows 2.5 ulp).  Applies for
 * 0 <= sharpen <= 0.24, precision 0 / 2 and output widths that are a multiple of 4 (fp32) / 8 (fp16);
 * everything else runs the exact kernels.
 * B2R_FLAG_EXACT_SHARPEN: every operation individually rounded in the reference's order -- output
 * bit-identical to oracle.sharpen (numpy float32 / float16 arithmetic); about 2x the sharpen time.
 * B2R_FLAG_FAST_SHARPEN: accepted for compatibility (round 1's opt-in); on its own it has no effect any more,
 * together with B2R_FLAG_EXACT_SHARPEN it selects round 1's approximate-division variant of the exact kernels. */
#define B2R_FLAG_FAST_SHARPEN 32u
#define B2R_FLAG_EXACT_SHARPEN 64u
/* Keep the C2R rows and the sharpen as two kernels with the pre-sharpen plane in HBM between them (the
 * reference's tempBuffer round trip).  Default: ONE fused kernel per frame whenever the tolerance-bound sharpen
 * applies and the row schedule is one of the built-in sizes -- same output bit for bit, ~200 MB less HBM
 * traffic per 4096x2048 frame (csrc/b2r_fused.cuh).  (b2r_download_pre_sharpen on a fused plan rebuilds the
 * plane with the stand-alone C2R kernel from the resident column spectra of the last frame.) */
#define B2R_FLAG_SEPARATE_SHARPEN 128u

typedef struct b2r_plan b2r_plan;

typedef struct b2r_plan_info {
    uint32_t w, h, up_w, up_h;         /* VkResample.cpp:1409-1418                               */
    uint32_t precision;
    float upscale, sharpen;
    uint32_t zeropad_lo_y, zeropad_hi_y; /* rows the inverse reads as zero, VkResample.cpp:1494-1495 */
    uint32_t spectrum_row_stride;      /* complex elements per spectrum row (W/2+1 rounded up)    */
    size_t input_bytes, output_bytes;
    size_t device_bytes;               /* total device memory owned by the plan                  */
    uint32_t n_stages[4];              /* radix schedule of the W, H, upH, upW transforms        */
    uint32_t radices[4][8];
    uint32_t threads[4];               /* threads cooperating on one sequence                    */
    uint32_t column_tile;              /* spectrum columns per CTA in the fused column kernel    */
    uint32_t kernels_per_frame;        /* launches one b2r_execute iteration performs            */
    uint32_t static_kernels;           /* bit0 K1, bit1 columns, bit2 K7: ahead-of-time schedule  */
    uint32_t jit_kernels;              /* same bits: compiled at plan time (subset of static_kernels) */
    char jit_note[128];                /* why plan-time JIT was not used, "" otherwise                */
    uint32_t c2c_mode;                 /* 1 if created with B2R_FLAG_C2C_PARITY                   */
    size_t pre_sharpen_plane_stride;   /* elements between planes of the pre-sharpen buffer        */
    uint32_t fused_strips_per_plane;   /* > 0: C2R rows + sharpen run as ONE kernel, this many strip CTAs per plane */
    uint32_t sharpen_mode;             /* 0 exact kernels, 1 tolerance-bound kernels (b2r_cas.cuh)                 */
} b2r_plan_info;

/* devices_list(), VkResample.cpp:239-268; createInstance..createDevice, :1286-1320 */
int b2r_device_count(void);
int b2r_device_name(int device, char* buf, size_t buf_len);

/* Plan build.  Replaces the configuration fill + allocateFFTBuffer x3 (VkResample.cpp:1409-1448),
 * initializeVulkanFFT x2 (:1506-1509), createShiftApp (:1562) and createSharpenApp (:1617).
 * upscale: float, upW = (uint32)(upscale*W) as in the reference.  precision: 0, 1 or 2 (1 needs NVRTC). */
int b2r_plan_create(b2r_plan** out, int device, uint32_t w, uint32_t h, float upscale,
                    uint32_t precision, float sharpen, uint32_t flags);
/* deleteVulkanFFT x2, deleteShiftApp x2, vkDestroyBuffer/vkFreeMemory, VkResample.cpp:1762-1778 */
void b2r_plan_destroy(b2r_plan* plan);

size_t b2r_plan_input_bytes(const b2r_plan* plan);   /* inputBufferSize, VkResample.cpp:1437 */
size_t b2r_plan_output_bytes(const b2r_plan* plan);  /* transfer size,   VkResample.cpp:1698 */
int b2r_plan_get_info(const b2r_plan* plan, b2r_plan_info* info);

/* transferDataFromCPU(&vkGPU, host, &inputBuffer, inputBufferSize), VkResample.cpp:1688 (def :385) */
int b2r_upload(b2r_plan* plan, const void* host_in);
/* performVulkanUpscale(..., numIter), VkResample.cpp:1692 (def :1249-1279): runs the whole frame
 * num_iter times back to back on the device-resident input and returns the average device time
 * per iteration in milliseconds (host<->device transfers excluded, like the reference). */
int b2r_execute(b2r_plan* plan, uint32_t num_iter, double* ms_per_iter);
/* transferDataToCPU(&vkGPU, host, &buffer, 3*upW*upH*elem), VkResample.cpp:1697-1700 (def :430) */
int b2r_download(b2r_plan* plan, void* host_out);

/* Convenience for host-buffer callers: upload + execute(1) + download on the plan's stream with
 * one synchronisation; ms_total (optional) is the device time including both copies. */
int b2r_upscale_host(b2r_plan* plan, const void* host_in, void* host_out, double* ms_total);

/* Device-resident access (synthetic frames, zero-copy pipelines, stage-level parity tests).
 * Pointers are CUDA device pointers owned by the plan. */
void* b2r_device_input(b2r_plan* plan);
void* b2r_device_output(b2r_plan* plan);
/* C2R result before the sharpen: planar, plane stride (upW+2)*upH elements (the reference's
 * tempBuffer view, VkResample.cpp:1593-1596).  D2H copy of 3*(upW+2)*upH elements. */
int b2r_download_pre_sharpen(b2r_plan* plan, void* host_out);
size_t b2r_plan_pre_sharpen_bytes(const b2r_plan* plan);
/* Stage-level entry points for parity tests: run only the sharpen kernel on a caller-provided
 * pre-sharpen buffer (host pointer, layout above), result to host_out (compact layout). */
int b2r_sharpen_host(b2r_plan* plan, const void* host_pre, void* host_out);
/* Batch / streaming form of b2r_execute: enqueue ONE frame that reads a caller-provided
 * device-resident input (layout as b2r_upload) and writes a caller-provided device output (compact
 * layout), asynchronously on the plan's stream.  The launchResample file loop (VkResample.cpp:1627)
 * with the transfers hoisted out; used for device-resident frame streams. */
int b2r_enqueue_device(b2r_plan* plan, const void* device_in, void* device_out);
/* Same with HOST buffers: H2D copy + frame + D2H copy of one frame, asynchronously; host_out is valid
 * after b2r_synchronize().  Pinned host memory is needed for the copies to overlap. */
int b2r_enqueue_host(b2r_plan* plan, const void* host_in, void* host_out);
/* Byte-pixel forms (extension; the reference README's planned "reading data in uint8"): the host
 * loops around the hot path -- in[c][j][i] = u8[j][i][c]/255 (VkResample.cpp:1636-1685) and
 * u8[j][i][c] = (uchar)(255.0*out[c][j][i]) (VkResample.cpp:1708-1748) -- run as two small kernels, so a
 * frame crosses PCIe as 3*W*H + 3*upW*upH bytes instead of 4x (fp32) / 2x (fp16) that.  host buffers
 * are interleaved RGB (stb_image's layout).  upload_u8/download_u8 pair with b2r_execute on lane 0;
 * enqueue_host_u8 is the asynchronous per-lane form (results valid after b2r_synchronize). */
size_t b2r_plan_input_u8_bytes(const b2r_plan* plan);
size_t b2r_plan_output_u8_bytes(const b2r_plan* plan);
int b2r_upload_u8(b2r_plan* plan, const unsigned char* host_rgb);
int b2r_download_u8(b2r_plan* plan, unsigned char* host_rgb);
int b2r_enqueue_host_u8(b2r_plan* plan, const unsigned char* host_rgb_in, unsigned char* host_rgb_out);
/* Completion tickets for the asynchronous host forms (pipelined batch front-ends: decode / encode threads
 * around one submitting thread).  Every b2r_enqueue_host / b2r_enqueue_host_u8 call takes the next ticket
 * (1, 2, 3, ...); b2r_plan_last_ticket returns the ticket of the most recent call; b2r_wait_ticket blocks until
 * that frame's output is complete in host_out.  b2r_wait_ticket only waits on a CUDA event: it may be called
 * from a different thread than the one that enqueues (the one exception to "one thread per plan").  The lane
 * count must not change between an enqueue and its wait. */
uint64_t b2r_plan_last_ticket(const b2r_plan* plan);
int b2r_wait_ticket(b2r_plan* plan, uint64_t ticket);
/* Pinned (page-locked) host memory for the asynchronous forms -- what the staging buffer of
 * transferDataFromCPU / transferDataToCPU is to the reference (VkResample.cpp:385-473).  NULL on failure. */
void* b2r_host_alloc(size_t bytes);
void b2r_host_free(void* ptr);
/* Number of lanes (1..8, default 1) that b2r_enqueue_device / b2r_enqueue_host rotate over.  Each
 * lane owns a stream and a private set of working buffers, so consecutive frames of a stream overlap
 * on the GPU (and copies overlap kernels) -- the equivalent of running the reference with
 * -numthreads N, N private VkFFT applications on one device (VkResample.cpp:1959-1969).  Frames that
 * are in flight together must use distinct caller buffers.  b2r_execute / b2r_upload / b2r_download
 * always use lane 0. */
int b2r_plan_set_lanes(b2r_plan* plan, uint32_t lanes);
uint32_t b2r_plan_lanes(const b2r_plan* plan);
/* CUDA-event stopwatch on the plan's stream (what performVulkanUpscale's submit..fence clock is
 * to the reference, VkResample.cpp:1270-1274): start records, stop records + waits + returns ms. */
int b2r_timer_start(b2r_plan* plan);
int b2r_timer_stop(b2r_plan* plan, double* ms);
/* Per-kernel device time: runs num_iter frames with an event between kernels and returns the
 * average milliseconds of {R2C rows, fused columns, C2R rows, sharpen} in ms_per_kernel[4]. */
int b2r_profile_kernels(b2r_plan* plan, uint32_t num_iter, double* ms_per_kernel);
/* Block until all work queued on the plan's stream has finished. */
int b2r_synchronize(b2r_plan* plan);
/* The CUDA stream (cudaStream_t) the plan launches on, for callers that time with their own events. */
void* b2r_plan_stream(b2r_plan* plan);
/* Number of kernel launches issued by the library since the plan was created (bench bookkeeping). */
uint64_t b2r_plan_launch_count(const b2r_plan* plan);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* b2r_last_error(void);
const char* b2r_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2RESAMPLE_H_ */
