/* oracle/shim/glm/gtc/constants.hpp -- TEST INFRASTRUCTURE ONLY: see ../glm.hpp */
#pragma once
#include "../glm.hpp"
namespace glm {
/* gtc/constants.inl: pi<T>() = T(3.14159265358979323846264338327950288) */
template <class T> constexpr T pi() { return T(3.14159265358979323846264338327950288); }
}
