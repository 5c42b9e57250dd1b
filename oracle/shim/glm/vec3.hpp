/* oracle/shim/glm/vec3.hpp -- TEST INFRASTRUCTURE ONLY: see glm.hpp */
#pragma once
#include "glm.hpp"
