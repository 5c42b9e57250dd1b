/*
 * oracle/camera_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Restatement of the reference orbit camera /root/reference/src/Camera.cpp ("Cam:n").
 * The reference does its vector math through glm, which is NOT vendored under
 * /root/reference and whose version is pinned nowhere (no CMake / conan / submodule), so
 * the glm operators used at Cam:19,32-37,49-56,73,91-92,99,122-149 are restated from glm's
 * published generic (non-SIMD) definitions:
 *     dot(vec3)  = x*x + y*y + z*z          dot(vec4) = (x*x + y*y) + (z*z + w*w)
 *     inversesqrt(x) = 1/sqrt(x)            normalize(v) = v * inversesqrt(dot(v,v))
 *     length(v) = sqrt(dot(v,v))            clamp(x,a,b) = min(max(x,a),b)
 *     cross(a,b) = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
 *     rotate(I, angle, (0,1,0)) * (1,0,0,0) = (cos, 0, -sin, 0)
 * PINNED to the reference source: src/Camera.cpp is compiled unmodified into oracle/_ref/libhost_ref.so
 * (against the GLM stand-in oracle/shim/glm) and this restatement equals it bit for bit over random walks
 * (tests/test_reference_pinning.py); "parity unpinned" only at the GLM boundary itself.
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

static const float PI_F = 3.14159265358979323846264338327950288f;   /* glm::pi<float>() */

static void normalize4(const float v[4], float out[4])
{
    const float d = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
    const float inv = 1.0f / sqrtf(d);
    for (int i = 0; i < 4; ++i) out[i] = v[i] * inv;
}

static void cross3(const float a[4], const float b[3], float out[3])
{
    out[0] = a[1] * b[2] - b[1] * a[2];
    out[1] = a[2] * b[0] - b[2] * a[0];
    out[2] = a[0] * b[1] - b[0] * a[1];
}

/* Cam:46-57.  The matrix is built from the ARGUMENTS (not the normalised members). */
static void set_view_matrix(orc_camera* c, const float eye[4], const float side[4],
                            const float up[4], const float look_at[4])
{
    memcpy(c->eye, eye, sizeof(float) * 4);
    normalize4(side, c->side);
    normalize4(up, c->up);
    normalize4(look_at, c->look_at);
    for (int i = 0; i < 4; ++i) {
        c->view2world[0 + i]  = side[i];
        c->view2world[4 + i]  = up[i];
        c->view2world[8 + i]  = -look_at[i];
        c->view2world[12 + i] = eye[i];
    }
}

void orc_camera_reset(orc_camera* c)                     /* Cam:30-44 */
{
    const float eye[4] = {0, 0, 3, 1}, side[4] = {1, 0, 0, 0}, up[4] = {0, 1, 0, 0},
                look[4] = {0, 0, -1, 0};
    set_view_matrix(c, eye, side, up, look);
    c->zenith = (float)((double)PI_F / 2.0);
    c->azimuth = 0.0f;
    c->radius = 3.0f;
    /* NB: is_changed is NOT touched here (Cam:30-44) */
}

void orc_camera_init(orc_camera* c, float y_fov, float rot_speed, float mov_speed)   /* Cam:16-23 */
{
    memset(c, 0, sizeof *c);
    c->y_fov = y_fov; c->rotation_speed = rot_speed; c->mov_speed = mov_speed;
    c->view_plane_dist = 1.0f / tanf(y_fov * PI_F / 360.0f);
    c->is_changed = 1;
    orc_camera_reset(c);
}

void orc_camera_ubo(orc_camera* c, float out[21])        /* Cam:59-80 */
{
    memcpy(out, c->view2world, sizeof(float) * 16);
    out[16] = c->eye[0]; out[17] = c->eye[1]; out[18] = c->eye[2]; out[19] = 1.0f;
    out[20] = c->view_plane_dist;
    /* the reference returns from inside the loop (Cam:63-72) before reaching
     * `is_changed = false` (Cam:79): the flag is never cleared. */
}

void orc_camera_set_orientation(orc_camera* c, float zoom, float zenith, float azimuth)   /* Cam:83-151 */
{
    if (zenith == 0 && azimuth == 0) {
        for (int i = 0; i < 4; ++i)
            c->eye[i] = (zoom > 0) ? c->eye[i] + c->look_at[i] : c->eye[i] - c->look_at[i];
        c->radius = sqrtf(c->eye[0] * c->eye[0] + c->eye[1] * c->eye[1] + c->eye[2] * c->eye[2]);
        for (int i = 0; i < 4; ++i) c->view2world[12 + i] = c->eye[i];
        c->is_changed = 1;
        return;
    }
    const float pi2 = PI_F * 2;
    float new_zenith = c->zenith + zenith * c->rotation_speed;
    new_zenith = fminf(fmaxf(new_zenith, 0.0f), PI_F);
    float new_azimuth = c->azimuth + azimuth * c->rotation_speed;
    if (new_azimuth < 0) new_azimuth = pi2 - new_azimuth;            /* (sic) Cam:103-104 */
    else if (new_azimuth > pi2) new_azimuth = new_azimuth - pi2;
    if (new_zenith == c->zenith && new_azimuth == c->azimuth) return;
    c->zenith = new_zenith;
    c->azimuth = new_azimuth;

    c->eye[0] = c->radius * sinf(c->zenith) * sinf(c->azimuth);
    c->eye[1] = c->radius * cosf(c->zenith);
    c->eye[2] = c->radius * sinf(c->zenith) * cosf(c->azimuth);
    c->eye[3] = 1;

    float look[4] = { -c->eye[0], -c->eye[1], -c->eye[2], 0.0f };
    normalize4(look, c->look_at);

    float side[4], up[4];
    const float world_up[3] = {0, 1, 0};
    if (c->zenith == 0 || c->zenith == PI_F) {
        side[0] = cosf(c->azimuth); side[1] = 0.0f; side[2] = -sinf(c->azimuth); side[3] = 0.0f;
    } else {
        cross3(c->look_at, world_up, side);
        side[3] = 0.0f;
    }
    cross3(side, c->look_at, up);
    up[3] = 0.0f;
    normalize4(side, c->side);
    normalize4(up, c->up);
    for (int i = 0; i < 4; ++i) {
        c->view2world[0 + i]  = c->side[i];
        c->view2world[4 + i]  = c->up[i];
        c->view2world[8 + i]  = -c->look_at[i];
        c->view2world[12 + i] = c->eye[i];
    }
    c->is_changed = 1;
}
