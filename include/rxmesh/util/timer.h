// rxmesh/util/timer.h -- GPUTimer (CUDA events on a stream) and CPUTimer (include/rxmesh/util/timer.h:10-69)
#pragma once
#include <chrono>

#include "rxmesh/util/macros.h"

namespace rxmesh {
struct GPUTimer
{
    GPUTimer(cudaStream_t stream = NULL) : m_stream(stream)
    {
        CUDA_ERROR(cudaEventCreate(&m_start));
        CUDA_ERROR(cudaEventCreate(&m_stop));
    }
    ~GPUTimer()
    {
        cudaEventDestroy(m_start);
        cudaEventDestroy(m_stop);
    }
    void  start() { CUDA_ERROR(cudaEventRecord(m_start, m_stream)); }
    void  stop() { CUDA_ERROR(cudaEventRecord(m_stop, m_stream)); }
    float elapsed_millis()
    {
        float elapsed = 0;
        CUDA_ERROR(cudaEventSynchronize(m_stop));
        CUDA_ERROR(cudaEventElapsedTime(&elapsed, m_start, m_stop));
        return elapsed;
    }

   private:
    cudaEvent_t  m_start, m_stop;
    cudaStream_t m_stream;
};
struct CPUTimer
{
    void  start() { m_start = std::chrono::high_resolution_clock::now(); }
    void  stop() { m_stop = std::chrono::high_resolution_clock::now(); }
    float elapsed_millis() { return std::chrono::duration<float, std::milli>(m_stop - m_start).count(); }

   private:
    std::chrono::high_resolution_clock::time_point m_start, m_stop;
};
}  // namespace rxmesh
