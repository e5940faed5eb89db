// rxmesh/util/timer.h -- the two stopwatch types user code takes from the reference (include/rxmesh/util/timer.h:10-69):
// GPUTimer measures between two CUDA events recorded on a stream, CPUTimer between two host clock readings; both report
// elapsed_millis().  One shared shape here: a stopwatch over a "now" policy.
#pragma once
#include <chrono>

#include "rxmesh/util/macros.h"

namespace rxmesh {
namespace detail {
// host clock readings
struct HostClock
{
    using Mark = std::chrono::steady_clock::time_point;
    void         open() {}
    void         close() {}
    void         mark(Mark& m) { m = std::chrono::steady_clock::now(); }
    static float millis(Mark& a, Mark& b) { return std::chrono::duration<float, std::milli>(b - a).count(); }
};
// CUDA events on one stream; the second mark is waited for when the time is read
struct EventClock
{
    using Mark = cudaEvent_t;
    explicit EventClock(cudaStream_t s = NULL) : stream(s) {}
    void         mark(Mark& m) { CUDA_ERROR(cudaEventRecord(m, stream)); }
    static float millis(Mark& a, Mark& b)
    {
        float ms = 0;
        CUDA_ERROR(cudaEventSynchronize(b));
        CUDA_ERROR(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
    cudaStream_t stream;
};
template <typename Clock>
class Stopwatch
{
   public:
    void  start() { m_clock.mark(m_begin); }
    void  stop() { m_clock.mark(m_end); }
    float elapsed_millis() { return Clock::millis(m_begin, m_end); }

   protected:
    explicit Stopwatch(Clock c) : m_clock(c) {}
    Clock                m_clock;
    typename Clock::Mark m_begin{}, m_end{};
};
}  // namespace detail

struct CPUTimer : detail::Stopwatch<detail::HostClock>
{
    CPUTimer() : Stopwatch(detail::HostClock()) {}
};
struct GPUTimer : detail::Stopwatch<detail::EventClock>
{
    GPUTimer(cudaStream_t stream = NULL) : Stopwatch(detail::EventClock(stream))
    {
        CUDA_ERROR(cudaEventCreate(&m_begin));
        CUDA_ERROR(cudaEventCreate(&m_end));
    }
    GPUTimer(const GPUTimer&)            = delete;
    GPUTimer& operator=(const GPUTimer&) = delete;
    ~GPUTimer()
    {
        cudaEventDestroy(m_begin);
        cudaEventDestroy(m_end);
    }
};
}  // namespace rxmesh
