// TEST INFRASTRUCTURE ONLY (oracle/): glm::length2 lives in glm.hpp of this shim
// (random-utils.cpp:1,26,37 and common-model.cpp:1 include <glm/gtx/norm.hpp>).
#pragma once
#include "../glm.hpp"
