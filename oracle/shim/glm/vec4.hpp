/* oracle/shim/glm/vec4.hpp -- TEST INFRASTRUCTURE ONLY: see glm.hpp */
#pragma once
#include "glm.hpp"
