// Empty stand-in (ours, not reference code) for the reference's
// include/rxmesh/util/report.h.  apps/VertexNormal/vertex_normal_ref.h includes
// that header but uses nothing from it; putting this
// directory first on the include path lets the reference header compile
// UNMODIFIED without spdlog / rapidjson / glm / CUDA.
#pragma once
