// scenes.h -- the scene builders that feed render(): the two the reference ships in main.cpp plus the
// mesh-on-ground scene BASELINE.json's config 3 describes (which the reference does not have).
#pragma once
#include "render.h"

namespace rtweekend {
Scene lots_of_balls(const Config& cfg);  // cover scene, main.cpp:23-83
Scene foo(const Config& cfg);            // OBJ mesh alone, main.cpp:85-136 (cfg.model must be set)
Scene mesh_on_ground(const Config& cfg); // ground sphere r=1000 + cfg.model mesh standing on it
}  // namespace rtweekend
