/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the reference hot path of gallickgunner/Volume-Renderer:
 * the GLSL 4.30 compute shader VolumeRenderer.cs, the orbit camera src/Camera.cpp, the
 * transfer-function spline src/CubicSpline.cpp and the DDS/PVM codec src/ddsbase.cpp.
 *
 * Nothing in the product path (volume-renderer_b200/, include/) may include, link or call
 * anything declared here.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py use it, and only as the checker.
 *
 * PARITY PINNING -- every part is pinned to the reference's OWN SOURCES, compiled where they lie under
 * /root/reference into oracle/_ref/ (oracle/Makefile; nothing is copied into the repo), and compared bit
 * for bit by tests/test_reference_pinning.py and tests/test_codec.py:
 *   - march: VolumeRenderer.cs itself -- the GLSL text, streamed through a 6-line sed (drops #version /
 *     layout qualifiers / forward declarations, `out T x` -> `T& x`) into g++ as the body of a C++ struct,
 *     with oracle/shim/glsl_compat.h supplying the GLSL types and built-ins -> _ref/libshader_ref.so.
 *     march_oracle.c equals it on every scenario that uses reference semantics, on random frames, and on
 *     the committed fixtures tests/golden/ref_*.npz that shader wrote.  What GLSL leaves implementation-
 *     defined (precision of normalize/length/division, texture filtering of an integer texture) is defined
 *     in glsl_compat.h as one correctly rounded IEEE-754 binary32 operation per operator -- a real GL
 *     driver may differ there in the last ulp; that boundary cannot be pinned without a GL stack.
 *     The extensions (step override, transfer function, opacity correction) do not exist in the shader and
 *     are pinned only through their identity settings.
 *   - camera / spline: src/Camera.cpp and src/CubicSpline.cpp, unmodified -> _ref/libhost_ref.so, built
 *     against a stand-in for GLM (oracle/shim/glm): GLM is a third-party dependency the reference neither
 *     vendors nor pins; the stand-in restates GLM 0.9.8/0.9.9's published generic definitions.
 *   - codec: src/ddsbase.cpp, unmodified -> _ref/libddsbase_ref.so, plus the golden .pvm fixtures it wrote.
 */
#ifndef VOLREN_ORACLE_H
#define VOLREN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FILTER_NEAREST = 0, ORC_FILTER_TRILINEAR = 1 };

/* Everything the shader reads (VolumeRenderer.cs:28-47) plus the documented extensions. */
typedef struct orc_params {
    int32_t width, height;          /* imageSize(render_texture)                      */
    int32_t dim[3];                 /* textureSize(vol_tex3D): Nx, Ny, Nz             */
    int32_t bytes_per_voxel;        /* 1 = R8UI, 2 = R16UI (RendererCore.cpp:419)     */
    float   voxel_size[3];          /* uniform location 1                             */
    float   cam[21];                /* UBO: mat4 view_mat (col-major), vec4 eye, vpd  */
    float   alpha_scale;            /* uniform location 0                             */
    int32_t min_val, max_val;       /* uniform locations 2,3 (after the +1000 rule)   */
    int32_t is_mip, view_top, view_bottom; /* uniform locations 4,5,6                 */
    /* ---- extensions (not in the shader) ---- */
    int32_t filter;                 /* ORC_FILTER_*                                    */
    float   step_scale;             /* step = reference step * step_scale (1 = ref)    */
    int32_t opacity_correction;     /* a' = 1-(1-a)^step_scale; identity at scale 1    */
    int32_t use_tf;                 /* src.a = tf_lut[iso] instead of src.a = v        */
    float   tf_lut[256];
    /* ---- which rows to render (row r is written iff r>=row_begin, r<row_end and
     *      (r-row_begin) % row_stride == 0); other rows of the output are untouched ---- */
    int32_t row_begin, row_end, row_stride;
} orc_params;

typedef struct orc_counters {
    uint64_t rays;        /* pixels processed                          */
    uint64_t rays_hit;    /* pixels whose ray hit the box              */
    uint64_t samples;     /* texture() calls executed                  */
} orc_counters;

/* Render.  voxels: x fastest, then y, then z (file order).  rgba: width*height*4 floats,
 * row 0 = bottom of the view.  touch (optional): one bit per voxel, bit (z*Ny+y)*Nx+x set
 * when a sample referenced that voxel (8 corners for trilinear).  nthreads<=1: scalar.
 * Returns 0, or -1 on bad arguments. */
int orc_render(const orc_params* p, const void* voxels, float* rgba,
               uint8_t* touch, orc_counters* counters, int nthreads);

/* The per-frame constants of VolumeRenderer.cs:62-83,109,146,122-124 as evaluated by the oracle:
 * pmin[3], pmax[3], half_len[3], denom[3], step_dvr, step_mip, fmin, fmax, frange. */
void orc_frame_consts(const orc_params* p, float out[17]);

/* Number of set bits in a touch bitmap of nvoxels bits. */
uint64_t orc_popcount(const uint8_t* touch, uint64_t nvoxels);

/* ---- synthetic volume `mix` on the host cores, x fastest; returns 0 or -1 ---- */
int orc_synth_mix(void* out, const int32_t dims[3], int bytes_per_voxel, uint32_t vmax, uint32_t seed,
                  int with_hash, int nthreads);

/* ---- Camera (src/Camera.cpp) ---- */
typedef struct orc_camera {
    float eye[4], side[4], up[4], look_at[4];
    float view2world[16];           /* column-major */
    float view_plane_dist, y_fov, rotation_speed, mov_speed;
    float zenith, azimuth, radius;
    int32_t is_changed;
} orc_camera;

void orc_camera_init(orc_camera* c, float y_fov, float rot_speed, float mov_speed);
void orc_camera_reset(orc_camera* c);
void orc_camera_set_orientation(orc_camera* c, float zoom, float zenith, float azimuth);
void orc_camera_ubo(orc_camera* c, float out21[21]);

/* ---- CubicSpline (src/CubicSpline.cpp) ---- */
#define ORC_SPLINE_MAX_KNOTS 64
typedef struct orc_spline {
    int32_t n_points;
    int32_t iso[ORC_SPLINE_MAX_KNOTS];
    float   color[ORC_SPLINE_MAX_KNOTS][4];
    float   coeffs[ORC_SPLINE_MAX_KNOTS][4];
    float   deriv[ORC_SPLINE_MAX_KNOTS][4];
    float   a[ORC_SPLINE_MAX_KNOTS][4], b[ORC_SPLINE_MAX_KNOTS][4],
            c[ORC_SPLINE_MAX_KNOTS][4], d[ORC_SPLINE_MAX_KNOTS][4];
} orc_spline;

int  orc_spline_calc(orc_spline* s, int n_points, const int32_t* iso, const float* color4);
void orc_spline_eval_iso(const orc_spline* s, int iso, float out4[4]);
void orc_spline_eval_t(const orc_spline* s, float t, int segment, float out4[4]);
/* 256-entry opacity LUT: lut[i] = clamp(spline(clamp(i, first knot, last knot)).w, 0, 1) */
void orc_spline_bake_alpha_lut(const orc_spline* s, float lut[256]);

/* ---- DDS / PVM codec restatement (src/ddsbase.cpp:394-452, 550-594, 768-858) ---- */
/* Decode a whole .pvm (or plain PVM) file image held in memory.  On success returns a
 * malloc'd payload (W*H*D*components bytes followed by the PVM3 strings, if any) and fills
 * the header fields; returns NULL on malformed input. */
uint8_t* orc_pvm_decode(const uint8_t* file, uint64_t file_bytes,
                        uint32_t* w, uint32_t* h, uint32_t* d, uint32_t* components,
                        float scale[3], uint64_t* payload_bytes);
/* ddsbase.cpp:872-893 */
uint32_t orc_checksum(const uint8_t* data, uint64_t bytes);
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
