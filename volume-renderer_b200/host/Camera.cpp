// Camera.cpp -- see Camera.h.  Reference: src/Camera.cpp ("Cam:n").
#include "Camera.h"

#include <cmath>

using vr::vec4;
using vr::mat4;

Camera::Camera()
    : is_changed(false), view_plane_dist(0), y_FOV(0), rotation_speed(0), mov_speed(0), zenith(0),
      azimuth(0), radius(0), tot_zenith(0), tot_azimuth(0), tot2_azimuth(0)
{
}

Camera::Camera(float y_FOV_, float rot_speed, float mov_speed_)                      // Cam:16-23
    : y_FOV(y_FOV_), rotation_speed(rot_speed), mov_speed(mov_speed_)
{
    view_plane_dist = 1 / std::tan(y_FOV * vr::kPi / 360);
    is_changed = true;
    resetCamera();
}

Camera::~Camera() {}

void Camera::resetCamera()                                                           // Cam:30-44
{
    view2world_mat = mat4();
    rot_mat = mat4();
    setViewMatrix(vec4(0, 0, 3, 1), vec4(1, 0, 0, 0), vec4(0, 1, 0, 0), vec4(0, 0, -1, 0));
    zenith = (float)(vr::kPi / 2.0);
    azimuth = 0;
    tot_zenith = 0;
    tot2_azimuth = tot_azimuth = 0;
    radius = 3;
    // like the reference, is_changed is left alone here
}

void Camera::setViewMatrix(vec4 eye_, vec4 side_, vec4 up_, vec4 look_at_)          // Cam:46-57
{
    eye = eye_;
    side = vr::normalize4(side_);
    up = vr::normalize4(up_);
    look_at = vr::normalize4(look_at_);
    // the matrix takes the arguments as passed, like the reference
    view2world_mat = mat4(side_, up_, -look_at_, eye_);
}

void Camera::setUBO(std::vector<float>& cam_data)                                    // Cam:59-80
{
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) cam_data.push_back(view2world_mat.col[c][r]);
    cam_data.push_back(eye.x);
    cam_data.push_back(eye.y);
    cam_data.push_back(eye.z);
    cam_data.push_back(1);
    cam_data.push_back(view_plane_dist);
    // Cam:63-72 returns before Cam:79 can clear is_changed, so the flag stays set
}

void Camera::rebuildFromAngles()                                                     // Cam:122-150
{
    eye.x = radius * std::sin(zenith) * std::sin(azimuth);
    eye.y = radius * std::cos(zenith);
    eye.z = radius * std::sin(zenith) * std::cos(azimuth);
    eye.w = 1;

    look_at = -eye;
    look_at.w = 0;
    look_at = vr::normalize4(look_at);

    if (zenith == 0 || zenith == vr::kPi) {
        // rotate(I, azimuth, +y) applied to (1,0,0,0)
        side = vec4(std::cos(azimuth), 0.0f, -std::sin(azimuth), 0.0f);
    } else {
        side = vr::cross3(look_at, vec4(0, 1, 0, 0));
    }
    up = vr::cross3(side, look_at);
    side = vr::normalize4(side);
    up = vr::normalize4(up);
    view2world_mat = mat4(side, up, -look_at, eye);
    is_changed = true;
}

void Camera::setOrientation(float zoom, float zenith_, float azimuth_)              // Cam:83-151
{
    if (zenith_ == 0 && azimuth_ == 0) {
        if (zoom > 0) eye = eye + look_at;
        else eye = eye - look_at;
        radius = vr::length3(eye.x, eye.y, eye.z);
        view2world_mat.col[3] = eye;
        is_changed = true;
        return;
    }
    const float pi2 = vr::kPi * 2;
    float new_zenith = this->zenith + zenith_ * rotation_speed;
    new_zenith = std::fmin(std::fmax(new_zenith, 0.0f), vr::kPi);
    float new_azimuth = this->azimuth + azimuth_ * rotation_speed;
    if (new_azimuth < 0) new_azimuth = pi2 - new_azimuth;
    else if (new_azimuth > pi2) new_azimuth = new_azimuth - pi2;
    if (new_zenith == this->zenith && new_azimuth == this->azimuth) return;
    this->zenith = new_zenith;
    this->azimuth = new_azimuth;
    rebuildFromAngles();
}

void Camera::setSpherical(float radius_, float zenith_, float azimuth_)
{
    radius = radius_;
    zenith = std::fmin(std::fmax(zenith_, 0.0f), vr::kPi);
    azimuth = azimuth_;
    rebuildFromAngles();
}
