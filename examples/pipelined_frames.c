/* pipelined_frames.c -- plain C against include/volren_b200.h (the drop-in boundary is a C ABI): an orbit of
 * frames rendered with two frames in flight (vr_render_submit / vr_render_wait).  The reference's render()
 * blocks on its timer query every frame (src/RendererCore.cpp:152); a front end that wants throughput submits
 * frame f + 1 before it collects frame f, and the host's per-frame work overlaps the GPU's.
 * Build:  gcc -std=c11 -O2 -Iinclude examples/pipelined_frames.c -Lvolume-renderer_b200/lib -lvolren_b200 \
 *             -Wl,-rpath,$PWD/volume-renderer_b200/lib -lm -o examples/pipelined_frames
 * Run (needs a B200; there is no CPU path):  examples/pipelined_frames [frames] */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "volren_b200.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != VR_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, vr_last_error()); return 1; } } while (0)

/* the 21-float camera block of Camera::setUBO (src/Camera.cpp:59-80): side | up | -look_at | eye, eye, view_plane_dist,
 * for an eye on a circle of radius r around the y axis (Camera.cpp:122-149 with zenith = pi/2) */
static void orbit_camera(float azimuth, float r, float cam[21])
{
    const float eye[3] = {r * sinf(azimuth), 0.0f, r * cosf(azimuth)};
    const float look[3] = {-eye[0] / r, 0.0f, -eye[2] / r};
    const float side[3] = {-look[2], 0.0f, look[0]};          /* cross(look, (0,1,0)) */
    const float up[3] = {0.0f, 1.0f, 0.0f};
    memset(cam, 0, 21 * sizeof(float));
    for (int i = 0; i < 3; ++i) { cam[i] = side[i]; cam[4 + i] = up[i]; cam[8 + i] = -look[i]; cam[12 + i] = eye[i]; cam[16 + i] = eye[i]; }
    cam[15] = 1.0f; cam[19] = 1.0f;
    cam[20] = 1.0f / tanf(30.0f * 3.14159265358979f / 360.0f); /* Camera.cpp:19, 30 degree field of view */
}

int main(int argc, char** argv)
{
    if (argc > 1 && (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help"))) { fprintf(stderr, "usage: %s [frames]\n", argv[0]); return 2; }
    const int frames = argc > 1 ? atoi(argv[1]) : 60;
    const int W = 1280, H = 720;
    vr_context* ctx = NULL;
    CHECK(vr_create(0, W, H, &ctx));
    const uint64_t dims[3] = {256, 256, 256};
    const float spacing[3] = {1.0f, 1.0f, 1.0f};
    CHECK(vr_upload_synthetic(ctx, dims, 2, spacing, 4095u, 0x5EED0003u, 1, NULL));
    vr_params p;
    vr_params_default(&p);
    p.alpha_scale = 0.05f; p.min_val = 0; p.max_val = 4095; p.filter = VR_FILTER_TRILINEAR;
    CHECK(vr_set_params(ctx, &p));
    float* buf[2] = {malloc((size_t)W * H * 16), malloc((size_t)W * H * 16)};   /* page-lock them (cudaHostRegister) for full speed */
    if (!buf[0] || !buf[1]) return 1;
    uint32_t ticket[2] = {0, 0};
    double kernel_ms = 0.0;
    for (int f = 0; f < frames; ++f) {
        float cam[21];
        orbit_camera(6.2831853f * (float)f / (float)frames, 3.0f, cam);
        CHECK(vr_set_camera(ctx, cam));
        CHECK(vr_render_submit(ctx, buf[f & 1], &ticket[f & 1]));            /* frame f is on its way ... */
        if (f > 0) {                                                          /* ... while frame f - 1 is collected */
            vr_render_stats st;
            CHECK(vr_render_wait(ctx, ticket[(f - 1) & 1], &st));
            kernel_ms += st.kernel_ms;                                        /* buf[(f - 1) & 1] now holds frame f - 1 */
        }
    }
    if (frames > 0) { vr_render_stats st; CHECK(vr_render_wait(ctx, ticket[(frames - 1) & 1], &st)); kernel_ms += st.kernel_ms; }
    printf("%d frames, %.3f ms of march per frame, centre pixel alpha of the last frame %.4f\n", frames,
           frames ? kernel_ms / frames : 0.0, buf[(frames - 1) & 1][((size_t)(H / 2) * W + W / 2) * 4 + 3]);
    free(buf[0]); free(buf[1]);
    vr_destroy(ctx);
    return 0;
}
