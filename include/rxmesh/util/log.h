// rxmesh/util/log.h -- RXMESH_TRACE / INFO / WARN / ERROR / CRITICAL (include/rxmesh/util/log.h:18-77).  The reference logs
// through spdlog (third party, fmt-style "{}" placeholders); here the same call sites format into a line on stderr.
// INFO and below are silent unless RXMESH_LOG=1 (the reference prints them by default; tests and benchmarks stay quiet).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>

namespace rxmesh {
namespace detail {
inline void log_format(std::ostringstream& os, const char* fmt)
{
    os << fmt;
}
template <typename A, typename... Rest>
inline void log_format(std::ostringstream& os, const char* fmt, const A& a, const Rest&... rest)
{
    for (; *fmt; ++fmt) {
        if (fmt[0] == '{' && fmt[1] == '}') {
            os << a;
            log_format(os, fmt + 2, rest...);
            return;
        }
        os << *fmt;
    }
}
template <typename... Args>
inline void log_line(int level, const char* tag, const char* fmt, const Args&... args)
{
    static const bool verbose = getenv("RXMESH_LOG") != nullptr;
    if (level < 2 && !verbose) return;
    std::ostringstream os;
    log_format(os, fmt, args...);
    fprintf(stderr, "[RXMesh %s] %s\n", tag, os.str().c_str());
}
}  // namespace detail
}  // namespace rxmesh
#define RXMESH_TRACE(...) rxmesh::detail::log_line(0, "trace", __VA_ARGS__)
#define RXMESH_INFO(...) rxmesh::detail::log_line(1, "info", __VA_ARGS__)
#define RXMESH_WARN(...) rxmesh::detail::log_line(2, "warn", __VA_ARGS__)
#define RXMESH_ERROR(...) rxmesh::detail::log_line(3, "error", __VA_ARGS__)
#define RXMESH_CRITICAL(...) rxmesh::detail::log_line(4, "critical", __VA_ARGS__)
