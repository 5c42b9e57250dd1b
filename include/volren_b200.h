/*
 * volren_b200.h -- C-ABI of the B200-native direct-volume raycaster (libvolren_b200.so).
 *
 * This is the drop-in boundary for ONE path of gallickgunner/Volume-Renderer: what the
 * reference does between RendererCore::readVolumeData's glTexImage3D and the RGBA32F image
 * RendererCore::render leaves in its FBO.  The reference has no FFI; RendererCore talks to
 * OpenGL directly.  Each entry point below names the reference GL call site it replaces
 * (paths relative to the reference tree).  Plain C types only: no torch, no C++, no CUDA
 * types in the signatures (a CUDA stream is passed as void*).
 *
 * Conventions
 *   - every function returns VR_OK (0) or a negative vr_status; vr_last_error() returns a
 *     human-readable message for the calling thread's last failure.  Nothing throws across
 *     the ABI, nothing aborts.
 *   - one caller thread per context (the reference is single threaded, RendererGUI.cpp:105).
 *   - images are W*H*4 float32, premultiplied RGBA with r=g=b, row 0 = BOTTOM of the view
 *     (GL image origin, VolumeRenderer.cs:58,96).
 *   - there is NO CPU fallback: every compute entry point fails with VR_ERR_CUDA when no
 *     CUDA device is usable.
 */
#ifndef VOLREN_B200_H
#define VOLREN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define VR_API __attribute__((visibility("default")))
#else
#define VR_API
#endif

typedef enum vr_status {
    VR_OK = 0,
    VR_ERR_INVALID = -1,        /* bad argument                                   */
    VR_ERR_CUDA = -2,           /* CUDA runtime/driver error (message has detail) */
    VR_ERR_NO_VOLUME = -3,      /* render before upload                           */
    VR_ERR_OOM = -4,
    VR_ERR_IO = -5,
    VR_ERR_FORMAT = -6,         /* malformed .raw.inf / .pvm                      */
    VR_ERR_TIMEOUT = -7         /* a peer-frame barrier wait gave up              */
} vr_status;

enum { VR_FILTER_NEAREST = 0, VR_FILTER_TRILINEAR = 1 };

/* kernel selection (all produce bit-identical images):
 *   DIRECT       one thread per ray, texels straight from the edge-replicated linear copy through L1/L2;
 *                handles every parameter combination, including the degenerate ones (min_val >= max_val,
 *                negative alpha_scale, non-finite LUT entries); also the instrumentation pass
 *   TEXPAIR_PIPE trilinear filter: layered array of (z, z+1) voxel pairs, ONE texture gather (tld4) returns the
 *                eight texels of a sample; software pipelined; forms for DVR / TF / MIP / view swizzles /
 *                opacity correction; optional result-identical empty-space skipping
 *   NEAREST_TEX  nearest filter (the reference's de-facto output): one integer-coordinate texel load (TLD) per
 *                sample from the source-type layered array; same pipeline, forms and skipping
 *   AUTO         TEXPAIR_PIPE or NEAREST_TEX by filter; DIRECT for frames they do not cover
 * A request the frame's parameters do not allow falls back to DIRECT.  (Values 2-6, 8, 9 were development
 * kernels of round 1 -- windowed TMA, LSU, two-gather, hybrids -- measured slower; they live in tools/lab.) */
enum { VR_KERNEL_AUTO = 0, VR_KERNEL_DIRECT = 1, VR_KERNEL_TEXPAIR_PIPE = 7, VR_KERNEL_NEAREST_TEX = 10 };

/* empty-space skipping (result-identical; see vr_cell_table_get) */
enum { VR_SKIP_AUTO = 0,   /* used when it pays: enough cells are empty under the current window */
       VR_SKIP_ON = 1,     /* whenever it is valid */
       VR_SKIP_OFF = 2 };

typedef struct vr_context vr_context;

/* The shader's plain uniforms (VolumeRenderer.cs:39-45; set by RendererCore.cpp:56-110 via
 * glUniform*) plus the documented extensions. */
typedef struct vr_params {
    float   alpha_scale;        /* location 0, RendererCore::setAlpha   RendererCore.cpp:56-60 */
    int32_t min_val;            /* location 2, the UNIFORM value, i.e. after the +1000 rule
                                   for 16-bit data, RendererCore.cpp:62-71                    */
    int32_t max_val;            /* location 3, RendererCore.cpp:73-82                          */
    int32_t is_mip;             /* location 4, RendererCore.cpp:84-88                          */
    int32_t view_top;           /* location 5, RendererCore.cpp:95                             */
    int32_t view_bottom;        /* location 6, RendererCore.cpp:96                             */
    /* ---- extensions; zero-initialised struct + vr_params_default() = reference behaviour */
    int32_t filter;             /* VR_FILTER_*; NEAREST is what an integer texture does        */
    float   step_scale;         /* march step = reference step (VolumeRenderer.cs:109) * this  */
    int32_t opacity_correction; /* a' = 1-(1-a)^step_scale (identity when step_scale == 1)     */
    int32_t use_tf;             /* src.a = tf_lut[round(v*255)] instead of src.a = v           */
    float   tf_lut[256];
    int32_t kernel;             /* VR_KERNEL_*                                                 */
    int32_t empty_skip;         /* VR_SKIP_*                                                   */
} vr_params;

typedef struct vr_render_stats {
    float    kernel_ms;         /* CUDA-event time of the march kernel(s) only; mirrors the
                                   GL_TIME_ELAPSED query, RendererCore.cpp:149-153             */
    float    total_ms;          /* whole call incl. copies (host-side clock)                   */
    uint32_t kernel_launches;   /* kernels of this library launched by the call                */
    uint32_t kernel_used;       /* VR_KERNEL_* actually run                                    */
    uint32_t skip_used;         /* 1: the empty-space skipping form of that kernel ran         */
} vr_render_stats;

/* footprint of the volume in this GPU's HBM, bytes (0 = that copy does not exist; the arrays are built on the
 * first frame that needs them) */
typedef struct vr_memory_info {
    uint64_t linear_bytes;          /* edge-replicated linear copy (DIRECT kernel, source of the others)  */
    uint64_t array_bytes;           /* source-type layered array (NEAREST_TEX)                            */
    uint64_t zpair_array_bytes;     /* z-pair layered array (TEXPAIR_PIPE): 2x the volume                 */
    uint64_t cell_table_bytes;      /* per-cell min/max + empty map                                       */
    uint64_t frame_bytes;
} vr_memory_info;

typedef struct vr_volume_stats {
    int32_t min_value, max_value;   /* RendererCore.cpp:362-384                                */
    float   histogram[256];         /* 0..100, RendererCore.cpp:386-405                        */
} vr_volume_stats;

/* ---- library ---- */
VR_API const char* vr_version(void);
VR_API const char* vr_last_error(void);
VR_API void        vr_params_default(vr_params* p);      /* RendererCore::RendererCore, RendererCore.cpp:13-26 */
VR_API int         vr_device_count(int* count);

/* ---- context: replaces RendererCore::setup/setupFBO (RendererCore.cpp:34-44,184-219):
 *      a W x H RGBA32F render target on CUDA device `device`. ---- */
VR_API int  vr_create(int device, int width, int height, vr_context** out);
VR_API void vr_destroy(vr_context* ctx);
VR_API int  vr_resize(vr_context* ctx, int width, int height);
VR_API int  vr_image_size(const vr_context* ctx, int* width, int* height);

/* ---- volume upload: replaces glTexImage3D(GL_R8UI|GL_R16UI ...) + the CLAMP_TO_EDGE state,
 *      RendererCore.cpp:408-419.  voxels: x fastest, then y, then z; tightly packed
 *      (GL_UNPACK_ALIGNMENT 1, RendererCore.cpp:417-418); host byte order as is.  The caller
 *      keeps ownership of `voxels` (the reference frees it right after, :420-433). ---- */
VR_API int vr_upload_volume(vr_context* ctx, const void* voxels, const uint64_t dims[3],
                            int bytes_per_voxel, const float voxel_size[3]);
/* same, source already resident in this device's HBM */
VR_API int vr_upload_volume_device(vr_context* ctx, const void* d_voxels, const uint64_t dims[3],
                                   int bytes_per_voxel, const float voxel_size[3]);
/* replaces glUniform3f(1, voxel_size...) RendererCore.cpp:103 */
VR_API int vr_set_voxel_size(vr_context* ctx, const float voxel_size[3]);

/* ---- min/max scan + 256-bin histogram on the GPU: RendererCore.cpp:360-405
 *      (64-bit safe; the reference's skip of index 8390640 is not replicated) ---- */
VR_API int vr_volume_stats_get(vr_context* ctx, vr_volume_stats* out);

/* ---- per-cell min/max table built at upload (the brick table of SURVEY.md 8f-2; no reference counterpart --
 *      the reference marches every sample, VolumeRenderer.cs:115-137).  Cells are cubes of 2^shift voxels;
 *      cell c covers, per axis, voxel indices [c*2^shift - 1, (c+1)*2^shift - 1] clamped to the volume (every
 *      voxel a sample based in the cell can touch); cells[a] = (dims[a] >> shift) + 1.  mins / maxs may be NULL;
 *      otherwise they receive cells[0]*cells[1]*cells[2] values, x fastest.  `empty_cells` = cells whose max
 *      is <= the current min_val, i.e. whose samples contribute exactly 0 (skipped by the SKIP kernels). ---- */
VR_API int vr_cell_table_get(vr_context* ctx, int* shift, int cells[3], uint16_t* mins, uint16_t* maxs,
                             uint64_t* empty_cells);
VR_API int vr_memory_info_get(vr_context* ctx, vr_memory_info* out);

/* ---- camera block: replaces glBufferData(GL_UNIFORM_BUFFER, 84 bytes) RendererCore.cpp:
 *      221-240; the same 21 floats Camera::setUBO emits (Camera.cpp:59-80): mat4 view_mat
 *      column-major, vec4 eye, float view_plane_dist. ---- */
VR_API int vr_set_camera(vr_context* ctx, const float cam21[21]);

/* ---- uniforms: replaces the glUniform* calls of RendererCore.cpp:56-110 ---- */
VR_API int vr_set_params(vr_context* ctx, const vr_params* p);
VR_API int vr_get_params(const vr_context* ctx, vr_params* p);

/* ---- multi-GPU screen-row-tile partition (no reference counterpart):
 *      this context renders only row tiles t with t % world == rank, tile = tile_rows rows.
 *      rank 0 / world 1 (default) renders everything. ---- */
VR_API int vr_set_partition(vr_context* ctx, int rank, int world, int tile_rows);
/* number of rows this context owns, and bytes of its compact (owned rows only) image */
VR_API int vr_owned_rows(const vr_context* ctx, int* rows);

/* ---- render: replaces glDispatchCompute + the timer query, RendererCore.cpp:147-155.
 *      vr_render writes the full W*H*4 float frame to HOST memory (rows this context does
 *      not own are zero).  vr_render_device writes into DEVICE memory on `cuda_stream`
 *      (a cudaStream_t, may be NULL; d_rgba == NULL renders into the context's own frame,
 *      the counterpart of the reference's FBO texture): `compact` = 0 -> full frame, rows not owned are left
 *      untouched; `compact` = 1 -> only the owned rows, packed in tile order (what the
 *      multi-GPU gather sends).  Both are synchronous like the reference (it blocks on the
 *      timer query every frame, RendererCore.cpp:152). ---- */
VR_API int vr_render(vr_context* ctx, float* host_rgba, vr_render_stats* stats);
/* copy of the context's own frame (the image the last vr_render / vr_render_device with
 * d_rgba == NULL left behind) to host memory; replaces glReadPixels on the FBO */
VR_API int vr_read_frame(vr_context* ctx, float* host_rgba);
VR_API int vr_render_device(vr_context* ctx, float* d_rgba, int compact, void* cuda_stream,
                            vr_render_stats* stats);
/* multi-GPU end to end (no reference counterpart): render this rank's row tiles (vr_set_partition) and copy
 * them into their rows of a FULL W x H host frame; with one host buffer shared by all rank processes
 * (shared memory, page-locked in each) every GPU's own PCIe link carries its share of the frame.  Rows
 * owned by other ranks are left untouched.  Synchronous. */
VR_API int vr_render_owned_to_host(vr_context* ctx, float* host_full_frame, vr_render_stats* stats);
/* pipelined form of the two calls above (no reference counterpart: RendererCore::render blocks every frame on its
 * timer query, RendererCore.cpp:152): vr_render_submit enqueues the frame -- this context's row tiles into their rows
 * of the W x H host frame, exactly what vr_render_owned_to_host writes; the whole frame when world == 1 -- and returns a
 * ticket at once; vr_render_wait blocks until that frame is complete in host memory and reports its stats (kernel_ms
 * = first band's start to last band's end, which overlaps the neighbouring frames' work).  At most TWO frames may be
 * in flight; the host buffers of frames in flight must differ and stay page-locked.  Setters that rewrite device state
 * (a new transfer function, window minimum, partition, volume) wait for the frames in flight by themselves; the
 * synchronous render calls refuse to run while a ticket is outstanding. */
VR_API int vr_render_submit(vr_context* ctx, float* host_frame, uint32_t* ticket);
VR_API int vr_render_wait(vr_context* ctx, uint32_t ticket, vr_render_stats* stats);
/* rank-major compact tiles [world][owned rows][W][4] -> full frame, on the device
 * (the de-interleave after the NCCL gather) */
VR_API int vr_assemble_tiles(vr_context* ctx, const float* d_gathered, float* d_frame,
                             int world, int tile_rows, void* cuda_stream);

/* ---- fused multi-GPU hand-off over NVLink peer memory (no reference counterpart): rank 0
 *      exports its frame buffer, every other rank (one process per GPU) maps it and renders
 *      its row tiles STRAIGHT into it -- vr_render_device(ctx, peer_frame, 0, ...) -- so the
 *      only per-frame collective left is a barrier.  The 64-byte handle is a cudaIpcMemHandle_t;
 *      ship it to the other processes any way you like (bench.py broadcasts it with NCCL). ---- */
VR_API int vr_frame_device_ptr(vr_context* ctx, float** d_frame);
VR_API int vr_frame_export_ipc(vr_context* ctx, unsigned char handle[64]);
VR_API int vr_frame_open_ipc(vr_context* ctx, const unsigned char handle[64], float** d_peer_frame);
VR_API int vr_frame_close_ipc(vr_context* ctx, float* d_peer_frame);
/* Frame barrier in the same peer memory (three words behind the owner's pixels), replacing the
 * per-frame NCCL barrier with two one-thread kernels in stream order:
 *   arrive:  every rank, after rendering frame `frame_no` (1, 2, 3 ...) into d_target_frame: a
 *            system-scope fence + atomic on the owner's arrival counter; on the owner additionally a
 *            bounded spin until all `world` ranks have arrived -> the frame is complete on the stream.
 *   release: owner, after consuming frame `frame_no`: publishes it; other ranks, before rendering
 *            frame_no + 1: bounded spin on the owner's word over NVLink.
 * Spins give up after 5 s (VR_PEER_TIMEOUT_MS overrides, e.g. when rank processes time-share one GPU) and set
 * the timed_out word (vr_peer_frame_status); they never hang the GPU. */
VR_API int vr_peer_frame_arrive(vr_context* ctx, float* d_target_frame, uint32_t frame_no, int world,
                                int is_owner, void* cuda_stream);
VR_API int vr_peer_frame_release(vr_context* ctx, float* d_target_frame, uint32_t frame_no, int is_owner,
                                 void* cuda_stream);
VR_API int vr_peer_frame_status(vr_context* ctx, float* d_target_frame, uint32_t* arrivals,
                                uint32_t* released, uint32_t* timed_out);
/* render + arrive in one call: this rank's row tiles are stored straight into `d_target_frame` (own or peer
 * frame) and the LAST CTA of the march kernel to finish publishes the arrival in the owner's barrier word, so no
 * separate signal kernel runs; on the owner the bounded wait for all `world` arrivals follows on the stream.
 * A wait that timed out in an earlier frame makes this call (and arrive / release) fail with VR_ERR_TIMEOUT until
 * vr_peer_frame_reset is called: a torn frame is never consumed silently. */
VR_API int vr_render_peer(vr_context* ctx, float* d_target_frame, uint32_t frame_no, int world, int is_owner,
                          void* cuda_stream, vr_render_stats* stats);
/* stats == NULL makes vr_render_peer ASYNCHRONOUS (nothing is waited for; frames run back to back in the stream);
 * vr_peer_kernel_ms then returns the march-kernel time of that call's frame_no (ring of the last 64 frames; it
 * waits for that kernel).  vr_peer_frame_wait_arrivals enqueues only the owner's wait for frame_no * world arrivals,
 * on a stream of the caller's choice -- on a consumer stream the owner's own march kernels are not held back by the
 * slowest rank (with two target frames every rank then runs one frame ahead of the consumer). */
VR_API int vr_peer_kernel_ms(vr_context* ctx, uint32_t frame_no, float* ms);
VR_API int vr_peer_frame_wait_arrivals(vr_context* ctx, float* d_target_frame, uint32_t frame_no, int world, void* cuda_stream);
VR_API int vr_peer_frame_reset(vr_context* ctx, float* d_target_frame);

/* ---- display/save step after the path: float RGBA -> RGB8 (clamp, no gamma), vertical
 *      flip; replaces glBlitFramebuffer/glReadPixels, RendererCore.cpp:158-171 ---- */
VR_API int vr_read_rgb8(vr_context* ctx, uint8_t* host_rgb, int flip_vertical);

/* ---- instrumentation (not on the timed path): exact number of distinct voxels referenced
 *      by the current frame and number of samples taken; feeds the roofline's algorithmic
 *      bytes. ---- */
VR_API int vr_count_frame(vr_context* ctx, uint64_t* distinct_voxels, uint64_t* samples,
                          uint64_t* rays_hit);

/* ---- deterministic synthetic volume `mix` generated straight into HBM and
 *      uploaded; with d_copy_out != NULL the raw x-fastest volume is also copied there
 *      (device pointer, dims[0]*dims[1]*dims[2]*bytes_per_voxel bytes). ---- */
VR_API int vr_upload_synthetic(vr_context* ctx, const uint64_t dims[3], int bytes_per_voxel,
                               const float voxel_size[3], uint32_t vmax, uint32_t seed,
                               int with_hash_noise, void* d_copy_out);
/* same generator into host memory through the GPU (for handing the oracle the same data) */
VR_API int vr_synthetic_to_host(int device, const uint64_t dims[3], int bytes_per_voxel,
                                uint32_t vmax, uint32_t seed, int with_hash_noise, void* host_out);

#ifdef __cplusplus
}
#endif
#endif /* VOLREN_B200_H */
