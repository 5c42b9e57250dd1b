/* oracle/shim/glm/gtc/matrix_transform.hpp -- TEST INFRASTRUCTURE ONLY: see ../glm.hpp */
#pragma once
#include "../glm.hpp"
namespace glm {
/* gtc/matrix_transform.inl (ext/matrix_transform.inl in 0.9.9): rotate(m, angle, v) */
inline mat4 rotate(const mat4& m, float angle, const vec3& v)
{
    const float a = angle;
    const float c = cos(a);
    const float s = sin(a);
    const vec3 axis(normalize(v));
    const vec3 temp((1.0f - c) * axis);
    mat4 Rotate;
    Rotate[0][0] = c + temp[0] * axis[0];
    Rotate[0][1] = temp[0] * axis[1] + s * axis[2];
    Rotate[0][2] = temp[0] * axis[2] - s * axis[1];
    Rotate[1][0] = temp[1] * axis[0] - s * axis[2];
    Rotate[1][1] = c + temp[1] * axis[1];
    Rotate[1][2] = temp[1] * axis[2] + s * axis[0];
    Rotate[2][0] = temp[2] * axis[0] + s * axis[1];
    Rotate[2][1] = temp[2] * axis[1] - s * axis[0];
    Rotate[2][2] = c + temp[2] * axis[2];
    mat4 Result;
    Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
    Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
    Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
    Result[3] = m[3];
    return Result;
}
}
