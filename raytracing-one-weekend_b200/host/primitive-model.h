// primitive-model.h -- selects the primitive container, exactly like the reference's selector (primitive-model.h:1-5):
// the std::variant model with -DRTWEEKEND_USE_VARIANT_PRIMITIVES, the virtual model otherwise.  Both expose the same
// names (Sphere, MovingSphere, Triangle, PrimitiveStore_t, MaterialStore_t, detail::flat, detail::material_of) and
// flatten to identical rtw_primitive arrays (tests/test_host.py::test_variant_and_oo_models_flatten_identically).
#pragma once
#ifdef RTWEEKEND_USE_VARIANT_PRIMITIVES
#include "variant-primitives.h"
#else
#include "oo-primitives.h"
#endif
