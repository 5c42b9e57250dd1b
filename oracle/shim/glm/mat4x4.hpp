/* oracle/shim/glm/mat4x4.hpp -- TEST INFRASTRUCTURE ONLY: see glm.hpp */
#pragma once
#include "glm.hpp"
