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
        Camera(float y_FOV, float rot_speed = 0.7f, float mov_speed = 0.3f);
        ~Camera();

        void resetCamera();
        void setOrientation(float zoom, float zenith, float azimuth);
        void setViewMatrix(vr::vec4 eye, vr::vec4 side, vr::vec4 up, vr::vec4 look_at);
        void setUBO(std::vector<float>& cam_data);

        // extension (not in the reference): place the eye at (radius, zenith, azimuth) directly;
        // same spherical -> cartesian and basis construction as setOrientation.
        void setSpherical(float radius, float zenith, float azimuth);

        bool is_changed;
        vr::vec4 look_at;
        vr::vec4 side;
        vr::vec4 up;
        vr::vec4 eye;
        vr::mat4 rot_mat;

    private:
        void rebuildFromAngles();
        float view_plane_dist, y_FOV,
        rotation_speed, mov_speed, zenith, azimuth, radius, tot_zenith, tot_azimuth, tot2_azimuth;
        vr::mat4 view2world_mat;
};
