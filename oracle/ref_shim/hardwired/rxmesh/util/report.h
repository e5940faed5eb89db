// stand-in for rxmesh/util/report.h (rapidjson-based run reports): declarations only, so that the UNINSTANTIATED host
// wrapper template in vertex_normal_hardwired.cuh parses; oracle/ref_hardwired.cu launches the kernel template itself
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
namespace rxmesh {
struct TestData
{
    std::string        test_name;
    std::vector<float> time_ms;
    std::vector<bool>  passed;
};
struct CustomReport
{
    explicit CustomReport(const std::string&) {}
    void command_line(int, char**) {}
    void device() {}
    void system() {}
    void model_data(const std::string&, uint32_t, uint32_t) {}
    template <typename T>
    void add_member(const std::string&, const T&) {}
    void add_test(const TestData&) {}
    void write(const std::string&, const std::string&) {}
};
struct GPUTimer
{
    void  start() {}
    void  stop() {}
    float elapsed_millis() { return 0.f; }
};
template <typename T>
bool compare(const T*, const T*, size_t, bool) { return true; }
inline std::string extract_file_name(const std::string& s) { return s; }
}  // namespace rxmesh
struct HardwiredArgs
{
    int         argc = 0;
    char**      argv = nullptr;
    std::string obj_file_name, output_folder;
    uint32_t    num_run = 1;
};
static HardwiredArgs Arg;
#define CUDA_ERROR(x) (x)
#define GPU_FREE(p) cudaFree(p)
#define DIVIDE_UP(a, b) (((a) + (b)-1) / (b))
