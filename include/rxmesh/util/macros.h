// rxmesh/util/macros.h -- the macros user code takes from the reference's util/macros.h (:22-89): DIVIDE_UP, INVALID*,
// STRINGIFY, CUDA_ERROR (log the CUDA error string and exit, macros.h:77-89).
#pragma once
#include <cuda_runtime_api.h>
#include <stdint.h>

#include "rxmesh/util/log.h"

namespace rxmesh {
typedef uint8_t flag_t;
}
#ifndef DIVIDE_UP
#define DIVIDE_UP(num, divisor) (((num) + (divisor) - 1) / (divisor))
#endif
#define ROUND_UP_TO_NEXT_MULTIPLE(num, mult) (DIVIDE_UP(num, mult) * (mult))
#ifndef INVALID64
#define INVALID64 0xFFFFFFFFFFFFFFFFu
#define INVALID32 0xFFFFFFFFu
#define INVALID16 0xFFFFu
#endif
#ifndef INVALID8
#define INVALID8 0xFFu
#define INVALID4 0xFu
#endif
#ifndef WARP_SIZE
#define WARP_SIZE 32
#endif
#define BYTES_TO_MEGABYTES(bytes) (double(bytes) / double(1024.0 * 1024.0))
#define STRINGIFY(x) TOSTRING(x)
#define TOSTRING(x) #x

#ifndef CUDA_ERROR
namespace rxmesh {
inline void HandleError(cudaError_t err, const char* file, int line)
{
    if (err != cudaSuccess) {
        RXMESH_ERROR("Line {} File {}", line, file);
        RXMESH_ERROR("CUDA ERROR: {}", cudaGetErrorString(err));
        exit(EXIT_FAILURE);
    }
}
}  // namespace rxmesh
#define CUDA_ERROR(err) (rxmesh::HandleError(err, __FILE__, __LINE__))
#endif
