// glm/fwd.hpp for user code written against the reference: glm (1.0.1, cmake/glm.cmake:11-12) is a third-party
// dependency the reference downloads at configure time; the subset its hot-path kernels use (vec, cross, dot, normalize,
// length, length2, distance, distance2, pi) is provided by rxmesh/types.h in namespace glm.  With the real glm on the
// include path ahead of this directory, that one is used instead.
#pragma once
#include "rxmesh/types.h"
