// rxmesh/util/cuda_query.h -- cuda_query(device) (include/rxmesh/util/cuda_query.h:60-120): select the device and return
// its properties; the device checks of this library (sm_100) are in rxm_init.
#pragma once
#include "rxmesh/util/macros.h"
#include "../../rxmesh_b200.h"

namespace rxmesh {
inline cudaDeviceProp cuda_query(const int dev)
{
    int device_count = 0;
    cudaGetDeviceCount(&device_count);
    if (device_count == 0) {
        RXMESH_ERROR("cuda_query() device cannot be found!!");
        exit(EXIT_FAILURE);
    }
    if (dev < 0 || dev >= device_count) {
        RXMESH_ERROR("cuda_query() device id {} is out of range ({} devices)", dev, device_count);
        exit(EXIT_FAILURE);
    }
    if (rxm_init(dev) != RXM_OK) {
        RXMESH_ERROR("cuda_query() {}", rxm_last_error());
        exit(EXIT_FAILURE);
    }
    cudaDeviceProp prop;
    CUDA_ERROR(cudaGetDeviceProperties(&prop, dev));
    RXMESH_INFO("Device {}: {}, compute capability {}.{}, {} SMs", dev, prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    return prop;
}
}  // namespace rxmesh
