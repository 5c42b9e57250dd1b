/* oracle/shim/glm/gtc/matrix_access.hpp -- TEST INFRASTRUCTURE ONLY: see ../glm.hpp */
#pragma once
#include "../glm.hpp"
namespace glm {
inline vec4 column(const mat4& m, int index) { return m[index]; }
inline mat4 column(const mat4& m, int index, const vec4& x) { mat4 r = m; r[index] = x; return r; }
}
