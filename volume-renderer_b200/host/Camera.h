// Camera.h -- orbit camera with the reference's public surface (include/Camera.h:9-33):
// resetCamera / setOrientation / setViewMatrix / setUBO and the is_changed flag, fp32
// throughout, no glm.  Semantics follow src/Camera.cpp of the reference line by line,
// including its quirks (is_changed is never cleared; negative azimuth wraps to 2*pi - a).
#pragma once

#include <vector>

#include "vecmath.h"

class Camera
{
public:
    Camera();
    // y_FOV in degrees; view_plane_dist = 1 / tan(y_FOV * pi / 360)  (src/Camera.cpp:19)
    Camera(float y_FOV, float rot_speed = 0.7f, float mov_speed = 0.3f);
    ~Camera();

    // ---- what RendererGUI / GlfwManager call (RendererGUI.cpp:42-46,363; GlfwManager.cpp:179,213) ----
    void resetCamera();                                               // eye (0,0,3), looking down -z
    void setOrientation(float zoom, float zenith, float azimuth);     // one mouse / scroll event
    // ---- what RendererCore calls ----
    void setViewMatrix(vr::vec4 eye, vr::vec4 side, vr::vec4 up, vr::vec4 look_at);
    void setUBO(std::vector<float>& cam_data);                        // appends the 21 floats of the UBO
    // ---- extension (not in the reference): place the eye at (radius, zenith, azimuth) directly;
    //      same spherical -> cartesian and basis construction as setOrientation ----
    void setSpherical(float radius, float zenith, float azimuth);

    bool is_changed;            // polled by RendererCore::render (src/RendererCore.cpp:144)
    vr::vec4 look_at;
    vr::vec4 side;
    vr::vec4 up;
    vr::vec4 eye;
    vr::mat4 rot_mat;

private:
    void rebuildFromAngles();   // src/Camera.cpp:122-150

    float view_plane_dist;
    float y_FOV;
    float rotation_speed, mov_speed;
    float zenith, azimuth, radius;                     // spherical position of the eye
    float tot_zenith, tot_azimuth, tot2_azimuth;       // kept for parity with the reference; unused there too
    vr::mat4 view2world_mat;                           // [side | up | -look_at | eye], column major
};
