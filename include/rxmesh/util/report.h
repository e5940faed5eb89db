// rxmesh/util/report.h -- Report / TestData / CustomReport (include/rxmesh/util/report.h:36-471): the JSON record every
// reference app and test writes (command line, device, system, model and patch statistics, per-test timings) with THE SAME
// member names, so records of both implementations can be compared key by key.  The reference builds a rapidjson document
// (third party, not in the tree); here a record is an ordered list of (key, already-serialised JSON value) pairs.
#pragma once
#include <cuda_runtime.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <ctime>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "rxmesh/rxmesh_static.h"
#include "rxmesh/util/macros.h"

namespace rxmesh {

struct TestData  // report.h:36-46; values left at -1 / empty are not written
{
    std::vector<float> time_ms;
    int32_t            num_blocks  = -1;
    int32_t            num_threads = -1;
    std::vector<bool>  passed;
    std::string        test_name   = "";
    int32_t            dyn_smem    = -1;
    int32_t            static_smem = -1;
    int32_t            num_reg     = -1;
};

namespace detail {
struct JsonObject
{
    std::vector<std::pair<std::string, std::string>> members;
    static std::string quote(const std::string& s)
    {
        std::string o = "\"";
        for (char c : s) {
            if (c == '"' || c == '\\') o += '\\', o += c;
            else if (c == '\n') o += "\\n";
            else if ((unsigned char)c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
            else o += c;
        }
        return o + "\"";
    }
    static std::string num(double v)
    {
        char b[40];
        snprintf(b, sizeof b, "%.9g", v);
        return b;
    }
    void add(const std::string& k, const std::string& v) { members.emplace_back(k, quote(v)); }
    void add(const std::string& k, const char* v) { members.emplace_back(k, quote(v)); }
    void add(const std::string& k, bool v) { members.emplace_back(k, v ? "true" : "false"); }
    void add(const std::string& k, int32_t v) { members.emplace_back(k, std::to_string(v)); }
    void add(const std::string& k, uint32_t v) { members.emplace_back(k, std::to_string(v)); }
    void add(const std::string& k, size_t v) { members.emplace_back(k, std::to_string(v)); }
    void add(const std::string& k, double v) { members.emplace_back(k, num(v)); }
    void add(const std::string& k, float v) { members.emplace_back(k, num(v)); }
    template <typename T>
    void add(const std::string& k, const std::vector<T>& v)
    {
        std::string a = "[";
        for (size_t i = 0; i < v.size(); ++i) {
            JsonObject t;
            t.add("", (T)v[i]);
            a += (i ? ", " : "") + t.members[0].second;
        }
        members.emplace_back(k, a + "]");
    }
    void add_object(const std::string& k, const JsonObject& o, int indent) { members.emplace_back(k, o.str(indent)); }
    std::string str(int indent = 0) const
    {
        const std::string pad((size_t)indent + 4, ' ');
        std::string       o = "{\n";
        for (size_t i = 0; i < members.size(); ++i)
            o += pad + quote(members[i].first) + ": " + members[i].second + (i + 1 < members.size() ? ",\n" : "\n");
        return o + std::string((size_t)indent, ' ') + "}";
    }
};
}  // namespace detail

struct Report
{
    Report() {}
    // report.h:53-101: record name, git state of the build (not available here: the members are kept, empty), a time stamp
    // that also becomes the file-name suffix
    Report(const std::string& record_name)
    {
        m_doc.add("Record Name", record_name);
        m_doc.add("git_sha", "");
        m_doc.add("git_local_changes", "");
        m_doc.add("git_refspec", "");
        const std::time_t t = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
        std::tm           tm_{};
        localtime_r(&t, &tm_);
        char date[64], suffix[64];
        strftime(date, sizeof date, "%a %b %d %H:%M:%S %Y", &tm_);
        strftime(suffix, sizeof suffix, "_D%m_%d_%Y__T%H_%M_%S.json", &tm_);
        m_doc.add("date", date);
        m_output_name_suffix = suffix;
    }
    void command_line(int argc, char** argv)  // report.h:105-115
    {
        std::string cmd(argc > 0 ? argv[0] : "");
        for (int i = 1; i < argc; i++)
            cmd = cmd + " " + std::string(argv[i]);
        m_doc.add("command_line", cmd);
    }
    void device()  // report.h:118-189
    {
        detail::JsonObject sub;
        int                id = 0, driver = 0, runtime = 0;
        cudaDeviceProp     p{};
        CUDA_ERROR(cudaGetDevice(&id));
        CUDA_ERROR(cudaGetDeviceProperties(&p, id));
        cudaDriverGetVersion(&driver);
        cudaRuntimeGetVersion(&runtime);
        int clock_khz = 0, mem_clock_khz = 0;
        cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, id);
        cudaDeviceGetAttribute(&mem_clock_khz, cudaDevAttrMemoryClockRate, id);
        sub.add("ID", (int32_t)id);
        sub.add("Name", std::string(p.name));
        sub.add("Driver Version", std::to_string(driver / 1000) + "." + std::to_string((driver % 100) / 10));
        sub.add("Runtime Version", std::to_string(runtime / 1000) + "." + std::to_string((runtime % 100) / 10));
        sub.add("CUDA API Version", (int32_t)CUDART_VERSION);
        sub.add("Compute Capability", std::to_string(p.major) + "." + std::to_string(p.minor));
        sub.add("Total amount of global memory (MB)", (double)((float)p.totalGlobalMem / 1048576.0f));
        sub.add("Total amount of shared memory per block (Kb)", (double)((float)p.sharedMemPerBlock / 1024.0f));
        sub.add("Multiprocessors", (int32_t)p.multiProcessorCount);
        sub.add("GPU Max Clock rate (GHz)", (double)(clock_khz * 1e-6f));
        sub.add("Memory Clock rate (GHz)", (double)(mem_clock_khz * 1e-6f));
        sub.add("Memory Bus Width (bit)", (int32_t)p.memoryBusWidth);
        sub.add("Peak Memory Bandwidth (GB/s)", 2.0 * mem_clock_khz * (p.memoryBusWidth / 8.0) / 1.0E6);
        m_doc.add_object("GPU Device", sub, 4);
    }
    void system()  // report.h:192-259
    {
        detail::JsonObject sub;
        char               host[256] = "";
        gethostname(host, sizeof host - 1);
        sub.add("Hostname", std::string(host));
#if defined(__clang__)
        sub.add("compiler_name", "clang");
        sub.add("compiler_version", std::string(__clang_version__));
#elif defined(__GNUC__)
        sub.add("compiler_name", "gcc");
        sub.add("compiler_version", std::string(__VERSION__));
#endif
        sub.add("C++ version", (int32_t)__cplusplus);
#ifdef NDEBUG
        sub.add("Build Mode", "Release");
#else
        sub.add("Build Mode", "Debug");
#endif
        m_doc.add_object("System", sub, 4);
    }
    // report.h:262-283: <folder>/<file name without extension><time suffix | .json>
    void write(const std::string& output_folder, const std::string& output_filename, bool append_time_to_file_name = true)
    {
        const std::string stem = output_filename.substr(0, output_filename.find_last_of('.'));
        const std::string full = output_folder + "/" + stem + (append_time_to_file_name ? m_output_name_suffix : ".json");
        if (!std::filesystem::is_directory(output_folder) || !std::filesystem::exists(output_folder))
            std::filesystem::create_directories(output_folder);
        std::ofstream ofs(full);
        if (!ofs.is_open()) {
            fprintf(stderr, "Report::write() can not open %s\n", full.c_str());
            return;
        }
        ofs << m_doc.str(0) << "\n";
    }
    // report.h:286-319: model and patch statistics
    void model_data(const std::string& model_name, const RXMeshStatic& rx, const std::string json_member_name = "Model")
    {
        detail::JsonObject sub;
        sub.add("model_name", model_name);
        sub.add("num_vertices", rx.get_num_vertices());
        sub.add("num_edges", rx.get_num_edges());
        sub.add("num_faces", rx.get_num_faces());
        sub.add("max_valence", rx.get_input_max_valence());
        sub.add("is_edge_manifold", rx.is_edge_manifold());
        sub.add("is_closed", rx.is_closed());
        sub.add("patch_size", rx.get_patch_size());
        sub.add("num_patches", rx.get_num_patches());
        sub.add("num_components", rx.get_num_components());
        sub.add("num_lloyd_run", rx.get_num_lloyd_run());
        sub.add("patching_time", rx.get_patching_time());
        uint32_t min_p = 0, max_p = 0, avg_p = 0;
        rx.get_max_min_avg_patch_size(min_p, max_p, avg_p);
        sub.add("min_patch_size", min_p);
        sub.add("max_patch_size", max_p);
        sub.add("avg_patch_size", avg_p);
        sub.add("per_patch_max_vertices", rx.get_per_patch_max_vertices());
        sub.add("per_patch_max_edges", rx.get_per_patch_max_edges());
        sub.add("per_patch_max_faces", rx.get_per_patch_max_faces());
        sub.add("ribbon_overhead (%)", rx.get_ribbon_overhead());
        m_doc.add_object(json_member_name, sub, 4);
    }
    void add_test(const TestData& t)  // report.h:322-360
    {
        detail::JsonObject sub;
        if (t.num_blocks != -1) sub.add("num_blocks", t.num_blocks);
        if (t.num_threads != -1) sub.add("num_threads", t.num_threads);
        if (t.dyn_smem != -1) sub.add("dynamic_shared_memory (b)", t.dyn_smem);
        if (t.static_smem != -1) sub.add("static_shared_memory (b)", t.static_smem);
        if (t.num_reg != -1) sub.add("num_register_per_thread", t.num_reg);
        if (!t.passed.empty()) sub.add("passed", t.passed);
        if (!t.time_ms.empty()) sub.add("time (ms)", t.time_ms);
        m_doc.add_object(t.test_name, sub, 4);
    }
    template <typename T>
    void add_member(std::string member_key, const T member_val)  // report.h:363-367
    {
        m_doc.add(member_key, member_val);
    }

   protected:
    detail::JsonObject m_doc;
    std::string        m_output_name_suffix = ".json";
};

class CustomReport : public Report  // report.h:448-471: a record for a mesh held by another library
{
   public:
    CustomReport() : Report() {}
    CustomReport(const std::string& record_name) : Report(record_name) {}
    void model_data(const std::string& model_name, const uint32_t num_vertices, const uint32_t num_faces)
    {
        detail::JsonObject sub;
        sub.add("model_name", model_name);
        sub.add("num_vertices", num_vertices);
        sub.add("num_faces", num_faces);
        m_doc.add_object("Model", sub, 4);
    }
};

}  // namespace rxmesh
