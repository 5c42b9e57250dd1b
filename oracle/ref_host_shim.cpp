/*
 * oracle/ref_host_shim.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * C-linkage wrappers around the REFERENCE's own Camera and CubicSpline classes so that tests can drive
 * them through ctypes.  This file contains no reference code: it includes the reference headers from
 * /root/reference/include at build time (-I) and is linked with /root/reference/src/Camera.cpp and
 * /root/reference/src/CubicSpline.cpp compiled unmodified where they lie, against the GLM stand-in under
 * oracle/shim/glm (see oracle/Makefile, target _ref/libhost_ref.so).
 */
#include <vector>

#include "Camera.h"
#include "CubicSpline.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void* ref_camera_new(float y_fov, float rot_speed, float mov_speed) { return new Camera(y_fov, rot_speed, mov_speed); }
REF_API void ref_camera_delete(void* c) { delete static_cast<Camera*>(c); }
REF_API void ref_camera_reset(void* c) { static_cast<Camera*>(c)->resetCamera(); }
REF_API void ref_camera_set_orientation(void* c, float zoom, float zenith, float azimuth)
{
    static_cast<Camera*>(c)->setOrientation(zoom, zenith, azimuth);
}
/* the 21 floats of Camera::setUBO, then the public eye / side / up / look_at members: 37 floats */
REF_API int ref_camera_state(void* c_, float out[37], int* is_changed)
{
    Camera* c = static_cast<Camera*>(c_);
    const bool before = c->is_changed;
    std::vector<float> ubo;
    c->setUBO(ubo);
    if (ubo.size() != 21) return -1;
    for (int i = 0; i < 21; ++i) out[i] = ubo[i];
    const glm::vec4* v[4] = {&c->eye, &c->side, &c->up, &c->look_at};
    for (int k = 0; k < 4; ++k) { out[21 + 4 * k] = v[k]->x; out[22 + 4 * k] = v[k]->y; out[23 + 4 * k] = v[k]->z; out[24 + 4 * k] = v[k]->w; }
    if (is_changed) { is_changed[0] = before ? 1 : 0; is_changed[1] = c->is_changed ? 1 : 0; }
    return 0;
}

REF_API void* ref_spline_new(int n, const int* iso, const float* color4)
{
    std::vector<CubicSpline::TransferFuncControlPoint> cps((size_t)n);
    for (int i = 0; i < n; ++i) {
        cps[i].iso_value = iso[i];
        cps[i].color = glm::vec4(color4[4 * i], color4[4 * i + 1], color4[4 * i + 2], color4[4 * i + 3]);
    }
    CubicSpline* s = new CubicSpline();
    s->calcCubicSpline(cps);
    return s;
}
REF_API void ref_spline_delete(void* s) { delete static_cast<CubicSpline*>(s); }
REF_API void ref_spline_eval_iso(void* s, int iso, float out[4])
{
    const glm::vec4 v = static_cast<CubicSpline*>(s)->getPointOnSpline(iso);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
REF_API void ref_spline_eval_t(void* s, float t, int seg, float out[4])
{
    const glm::vec4 v = static_cast<CubicSpline*>(s)->getPointOnSpline(t, (float)seg);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
