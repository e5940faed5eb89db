// rxmesh/rxmesh_static.h -- RXMeshStatic (include/rxmesh/rxmesh_static.h:37-1245) for the static hot path,
// over librxmesh_b200 (C ABI in include/rxmesh_b200.h).  Same names / argument meaning; user kernels written
// against the reference (Query::dispatch, for_each<Op>, Attribute::operator()) compile against this header.
#pragma once
#include <algorithm>
#include <array>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <unordered_map>
#include <vector>

#include "rxmesh/attribute.h"
#include "rxmesh/context.h"
#include "rxmesh/kernels/for_each.cuh"
#include "rxmesh/launch_box.h"
#include "rxmesh/query.h"
#include "rxmesh/util/import_obj.h"
#include "rxmesh/util/macros.h"
#include "rxmesh/util/timer.h"

namespace rxmesh {

// rx_init(device_id, log level) (rxmesh.h:23-30): a negative device id skips the device selection, like the reference;
// the second argument is the reference's spdlog level (any integral / enum value), there is no logger to configure here
template <typename LevelT = int>
inline void rx_init(int device_id = 0, LevelT = LevelT())
{
    if (device_id >= 0) detail::rxm_check(rxm_init(device_id));
}

namespace detail {
inline void check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {  // CUDA_ERROR semantics (util/macros.h:77-89)
        fprintf(stderr, "rxmesh_b200: %s launch failed: %s\n", what, cudaGetErrorString(e));
        exit(EXIT_FAILURE);
    }
}
template <typename T>
__global__ static void flags_to_value(T* data, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] = reinterpret_cast<const uint32_t*>(data)[i] ? T(1) : T(0);
}
// detail::query_kernel (kernels/query_kernel.cuh:12-24)
template <uint32_t blockThreads, Op op, typename LambdaT>
__global__ static void query_kernel(const Context context, const bool oriented, LambdaT user_lambda)
{
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<op>(block, shrd_alloc, user_lambda, oriented);
}
// detail::for_each_vertex/edge/face (kernels/for_each.cuh:30-105): owned elements are the prefix [0, n_owned)
template <typename HandleT, typename LambdaT>
__global__ static void for_each_kernel(const Context context, LambdaT apply)
{
    const rxm::PatchDesc* d = context.view.desc + blockIdx.x;
    const uint32_t        n = d->n_owned[HandleT::elem], pid = d->patch_id;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        apply(HandleT(pid, typename HandleT::LocalT((uint16_t)i)));
}
}  // namespace detail

class RXMeshStatic
{
   public:
    RXMeshStatic(const RXMeshStatic&) = delete;

    // RXMeshStatic(fv, patcher_file, patch_size, ...) (rxmesh_static.h:61-100). patcher_file's role (a saved
    // patching) is played by an explicit face -> patch array.
    // extension: an explicit face -> patch array instead of a saved file
    RXMeshStatic(const std::vector<std::vector<uint32_t>>& fv, const std::vector<uint32_t>& face_patch,
                 const uint32_t patch_size)
    {
        std::vector<uint32_t> flat;
        flat.reserve(3 * fv.size());
        for (const auto& f : fv) {
            if (f.size() != 3) {  // rxmesh.cpp:590-597
                fprintf(stderr, "rxmesh_b200: non-triangular faces are not supported\n");
                exit(EXIT_FAILURE);
            }
            flat.insert(flat.end(), f.begin(), f.end());
        }
        init(flat.data(), (uint32_t)fv.size(), face_patch, patch_size);
    }
    RXMeshStatic(const uint32_t* fv, uint32_t num_faces, const std::vector<uint32_t>& face_patch = {},
                 const uint32_t patch_size = 512)
    {
        init(fv, num_faces, face_patch, patch_size);
    }
    // RXMeshStatic(fv, patcher_file, patch_size): replay a patching saved by Patcher::save / RXMesh::save
    // (rxmesh_static.h:61-66, patcher/patcher.h:154-182)
    // capacity_factor / patch_alloc_factor / lp_hashtable_load_factor size the reference's slack for DYNAMIC changes and
    // its cuckoo tables; a static mesh with direct owner tables has neither, the arguments are accepted and ignored
    explicit RXMeshStatic(const std::vector<std::vector<uint32_t>>& fv, const std::string patcher_file = "",
                          const uint32_t patch_size = 512, const float /*capacity_factor*/ = 1.0,
                          const float /*patch_alloc_factor*/ = 1.0, const float /*lp_hashtable_load_factor*/ = 0.8)
    {
        std::vector<uint32_t> flat;
        flat.reserve(3 * fv.size());
        for (const auto& f : fv) {
            if (f.size() != 3) {  // rxmesh.cpp:590-597
                fprintf(stderr, "rxmesh_b200: non-triangular faces are not supported\n");
                exit(EXIT_FAILURE);
            }
            flat.insert(flat.end(), f.begin(), f.end());
        }
        std::vector<uint32_t> face_patch;
        uint32_t              ps = patch_size;
        if (!patcher_file.empty()) {
            rxm_patcher_file pf;
            if (rxm_patcher_file_read(patcher_file.c_str(), &pf) == RXM_OK && pf.len[0] == fv.size()) {
                face_patch.assign(pf.vec[0], pf.vec[0] + pf.len[0]);
                ps = pf.header[0];
                rxm_patcher_file_free(&pf);
            } else {  // rxmesh.cpp:303-317: log and build fresh patches
                fprintf(stderr, "RXMesh::build patch file %s does not exist or does not match. Building unique patches.\n",
                        patcher_file.c_str());
            }
        }
        init(flat.data(), (uint32_t)fv.size(), face_patch, ps);
    }
    // RXMeshStatic(file_path, patcher_file, patch_size) (rxmesh_static.h:61-66): OBJ input, positions kept as the
    // input vertex coordinates (import_obj semantics of util/import_obj.h: "v x y z" and "f a b c" / "f a/b/c ..." lines,
    // 1-based or negative indices, triangles only)
    explicit RXMeshStatic(const std::string file_path, const std::string patcher_file = "", const uint32_t patch_size = 512,
                          const float /*capacity_factor*/ = 1.0, const float /*patch_alloc_factor*/ = 1.0,
                          const float /*lp_hashtable_load_factor*/ = 0.8)
        : RXMeshStatic(read_obj_faces(file_path), patcher_file, patch_size)
    {
        std::vector<std::vector<float>> verts;
        std::vector<std::vector<uint32_t>> faces;
        read_obj(file_path, verts, faces);
        add_vertex_coordinates(verts);
    }
    // RXMeshStatic(files_path, patch_size) (rxmesh_static.h:99-100, rxmesh_static.cu:66-161): several OBJ files loaded as
    // ONE mesh (each file's vertex ids offset by the vertices read so far); every face / vertex / edge carries the index of
    // the file it came from as its region label
    explicit RXMeshStatic(const std::vector<std::string> files_path, const uint32_t patch_size = 512)
    {
        std::vector<std::vector<float>>    verts;
        std::vector<std::vector<uint32_t>> faces;
        std::vector<int>                   region_num_faces, region_num_vertices;
        for (const auto& path : files_path) {
            read_obj(path, verts, faces);
            region_num_faces.push_back((int)faces.size());
            region_num_vertices.push_back((int)verts.size());
        }
        const std::vector<uint32_t> flat = flatten(faces);
        init(flat.data(), (uint32_t)faces.size(), {}, patch_size);
        m_num_regions = (int)files_path.size();
        add_vertex_coordinates(verts);
        m_face_label   = add_face_attribute<int>("rx:face_label", 1, LOCATION_ALL);
        m_edge_label   = add_edge_attribute<int>("rx:edge_label", 1, LOCATION_ALL);
        m_vertex_label = add_vertex_attribute<int>("rx:vertex_label", 1, LOCATION_ALL);
        auto label_of = [](const std::vector<int>& ends, int id) {  // first region whose end is past id
            return (int)std::distance(ends.begin(), std::upper_bound(ends.begin(), ends.end(), id));
        };
        for_each_face(HOST, [&](const FaceHandle fh) { (*m_face_label)(fh) = label_of(region_num_faces, (int)map_to_global(fh)); }, NULL, false);
        for_each_vertex(HOST, [&](const VertexHandle vh) { (*m_vertex_label)(vh) = label_of(region_num_vertices, (int)map_to_global(vh)); }, NULL, false);
        m_face_label->move(HOST, DEVICE);
        m_vertex_label->move(HOST, DEVICE);
        add_edge_labels(*m_face_label, *m_edge_label);
        m_edge_label->move(DEVICE, HOST);
    }
    // an edge takes the label of its faces (rxmesh_static.cu:716-727); public because nvcc wants extended lambdas in
    // public member functions
    void add_edge_labels(FaceAttribute<int>& face_label, EdgeAttribute<int>& edge_label)
    {
        for_each<Op::FE, 256>([face_label, edge_label] __device__(const FaceHandle fh, const EdgeIterator iter) {
            const int label = face_label(fh);
            edge_label(iter[0]) = label, edge_label(iter[1]) = label, edge_label(iter[2]) = label;
        });
        detail::rxm_check(cudaDeviceSynchronize() == cudaSuccess ? RXM_OK : RXM_ERR_CUDA);
    }
    // region labels (rxmesh_static.h:828-843): null for a mesh built from one input
    int                                   get_num_regions() const { return m_num_regions; }
    std::shared_ptr<FaceAttribute<int>>   get_face_region_label() { return m_face_label; }
    std::shared_ptr<EdgeAttribute<int>>   get_edge_region_label() { return m_edge_label; }
    std::shared_ptr<VertexAttribute<int>> get_vertex_region_label() { return m_vertex_label; }
    // bounding_box / scale of the input coordinates (rxmesh_static.cu:600-667)
    void bounding_box(glm::vec3& lower, glm::vec3& upper)
    {
        for (int i = 0; i < 3; ++i)
            lower[i] = std::numeric_limits<float>::max(), upper[i] = std::numeric_limits<float>::lowest();
        auto coord = *get_input_vertex_coordinates();
        for_each_vertex(HOST, [&](const VertexHandle vh) {
            for (int i = 0; i < 3; ++i)
                lower[i] = std::min(lower[i], coord(vh, i)), upper[i] = std::max(upper[i], coord(vh, i));
        }, NULL, false);
    }
    void scale(glm::fvec3 lower, glm::fvec3 upper)
    {
        if (lower[0] > upper[0] || lower[1] > upper[1] || lower[2] > upper[2]) {
            fprintf(stderr, "RXMeshStatic::scale() can not scale the mesh since the lower corner is higher than upper corner\n");
            return;
        }
        glm::vec3 bb_lower(0), bb_upper(0);
        bounding_box(bb_lower, bb_upper);
        float the_factor = std::numeric_limits<float>::max();
        for (int i = 0; i < 3; ++i)
            the_factor = std::min(the_factor, (upper[i] - lower[i]) / ((bb_upper[i] - bb_lower[i]) + std::numeric_limits<float>::epsilon()));
        auto coord = *get_input_vertex_coordinates();
        for_each_vertex(HOST, [&](const VertexHandle vh) {
            for (int i = 0; i < 3; ++i) {
                coord(vh, i) += (lower[i] - bb_lower[i]);
                coord(vh, i) *= the_factor;
            }
        }, NULL, false);
        coord.move(HOST, DEVICE);
    }
    // add_vertex_coordinates (rxmesh_static.h:110-112): attach positions to a mesh built from faces only
    void add_vertex_coordinates(std::vector<std::vector<float>>& vertices, std::string = "")
    {
        if (m_attrs.count("rx:vertices")) return;
        m_input_coords = add_vertex_attribute<float>(vertices, "rx:vertices");
    }
    // get_input_vertex_coordinates (rxmesh_static.h:823)
    std::shared_ptr<VertexAttribute<float>> get_input_vertex_coordinates()
    {
        if (!m_input_coords) {
            fprintf(stderr, "RXMeshStatic::get_input_vertex_coordinates input vertex was not initialized. Call RXMeshStatic "
                            "with a constructor to the obj file path\n");
            exit(EXIT_FAILURE);
        }
        return m_input_coords;
    }
    // RXMesh::save (rxmesh.h:326-329)
    void save(const std::string& filename) const { detail::rxm_check(rxm_mesh_save_patcher_file(m_mesh, filename.c_str())); }
    virtual ~RXMeshStatic()
    {
        m_attrs.clear();
        rxm_mesh_destroy(m_mesh);
    }

    // ---- getters (rxmesh.h:44-399) ----
    uint32_t get_num_vertices() const { return info(RXM_INFO_NUM_VERTICES); }
    uint32_t get_num_edges() const { return info(RXM_INFO_NUM_EDGES); }
    uint32_t get_num_faces() const { return info(RXM_INFO_NUM_FACES); }
    uint32_t get_num_patches() const { return info(RXM_INFO_NUM_PATCHES); }
    uint32_t get_patch_size() const { return info(RXM_INFO_PATCH_SIZE); }
    uint32_t get_input_max_valence() const { return info(RXM_INFO_MAX_VALENCE); }
    uint32_t get_input_max_edge_incident_faces() const { return info(RXM_INFO_MAX_EDGE_INCIDENT_FACES); }
    uint32_t get_input_max_face_adjacent_faces() const { return info(RXM_INFO_MAX_FACE_ADJACENT_FACES); }
    bool     is_closed() const { return info(RXM_INFO_IS_CLOSED) != 0; }
    bool     is_edge_manifold() const { return info(RXM_INFO_IS_EDGE_MANIFOLD) != 0; }
    uint32_t get_per_patch_max_vertices() const { return info(RXM_INFO_MAX_VERTICES_PER_PATCH); }
    uint32_t get_per_patch_max_edges() const { return info(RXM_INFO_MAX_EDGES_PER_PATCH); }
    uint32_t get_per_patch_max_faces() const { return info(RXM_INFO_MAX_FACES_PER_PATCH); }
    const Context& get_context() const { return m_context; }
    uint32_t get_max_num_patches() const { return get_num_patches(); }  // static meshes never add patches (rxmesh.h:209)
    // per-patch counts (rxmesh.h:355-399): owned elements, and all elements of the patch, ribbon included
    uint16_t get_num_owned_vertices(const uint32_t p) const { return (uint16_t)patch_view(p).n_owned[RXM_V]; }
    uint16_t get_num_owned_edges(const uint32_t p) const { return (uint16_t)patch_view(p).n_owned[RXM_E]; }
    uint16_t get_num_owned_faces(const uint32_t p) const { return (uint16_t)patch_view(p).n_owned[RXM_F]; }
    uint16_t get_num_vertices(const uint32_t p) const { return (uint16_t)patch_view(p).n[RXM_V]; }
    uint16_t get_num_edges(const uint32_t p) const { return (uint16_t)patch_view(p).n[RXM_E]; }
    uint16_t get_num_faces(const uint32_t p) const { return (uint16_t)patch_view(p).n[RXM_F]; }
    // faces per patch, ribbon included (patcher/patcher.h:115-133) and ribbon faces in percent of #faces (:139-142)
    void get_max_min_avg_patch_size(uint32_t& min_p, uint32_t& max_p, uint32_t& avg_p) const
    {
        max_p = 0, min_p = get_num_faces();
        uint64_t sum = 0;
        for (uint32_t p = 0; p < get_num_patches(); ++p) {
            const uint32_t n = get_num_faces(p);
            sum += n, max_p = std::max(max_p, n), min_p = std::min(min_p, n);
        }
        avg_p = (uint32_t)((float)sum / (float)get_num_patches());
    }
    // Patcher statistics the reference reports (rxmesh.h:242-262, patcher/patcher.h:95-113): connected components of the
    // input, assignment passes of the Lloyd patcher, patching time in milliseconds
    uint32_t get_num_components() const { return info(RXM_INFO_NUM_COMPONENTS); }
    uint32_t get_num_lloyd_run() const { return info(RXM_INFO_LLOYD_RUNS); }
    float    get_patching_time() const { return (float)(1e3 * rxm_mesh_build_seconds(m_mesh, 1)); }
    double get_ribbon_overhead() const
    {
        return 100.0 * (double(info(RXM_INFO_TOTAL_LOCAL_F)) - double(get_num_faces())) / double(get_num_faces());
    }
    // the linear-id prefix per patch (rxmesh.h:425-447)
    template <typename HandleT>
    const uint32_t* get_element_prefix(locationT location) const
    {
        return (location & DEVICE) ? rxm_mesh_device_lin_base(m_mesh, HandleT::elem) : rxm_mesh_lin_base(m_mesh, HandleT::elem);
    }
    // map_to_local_vertex / edge / face (rxmesh.h:336-350, rxmesh.cpp:998-1049): linear id -> handle of the owner
    const VertexHandle map_to_local_vertex(uint32_t i) const { return map_to_local<VertexHandle>(i); }
    const EdgeHandle   map_to_local_edge(uint32_t i) const { return map_to_local<EdgeHandle>(i); }
    const FaceHandle   map_to_local_face(uint32_t i) const { return map_to_local<FaceHandle>(i); }
    // get_edge_id (rxmesh.h:320, rxmesh.cpp:1052-1083): id of the edge between two INPUT vertex ids in the reference's
    // numbering (order of first appearance over the faces), INVALID32 when they share no edge
    uint32_t get_edge_id(const uint32_t v0, const uint32_t v1) const
    {
        if (m_edge_ids.empty()) {
            const uint32_t* ev = rxm_mesh_edges(m_mesh);
            for (uint32_t e = 0; e < get_num_edges(); ++e)
                m_edge_ids[((uint64_t)std::max(ev[2 * e], ev[2 * e + 1]) << 32) | std::min(ev[2 * e], ev[2 * e + 1])] = e;
        }
        auto it = m_edge_ids.find(((uint64_t)std::max(v0, v1) << 32) | std::min(v0, v1));
        return it == m_edge_ids.end() ? INVALID32 : it->second;
    }
    rxm_mesh*      c_handle() const { return m_mesh; }

    // ---- attributes (rxmesh_static.h:608-806) ----
    template <class T>
    std::shared_ptr<VertexAttribute<T>> add_vertex_attribute(const std::string& name, uint32_t num_attributes,
                                                             locationT location = LOCATION_ALL, layoutT layout = AoSoA)
    {
        return add<VertexAttribute<T>>(name, num_attributes, location, layout);
    }
    template <class T>
    std::shared_ptr<EdgeAttribute<T>> add_edge_attribute(const std::string& name, uint32_t num_attributes,
                                                         locationT location = LOCATION_ALL, layoutT layout = AoSoA)
    {
        return add<EdgeAttribute<T>>(name, num_attributes, location, layout);
    }
    template <class T>
    std::shared_ptr<FaceAttribute<T>> add_face_attribute(const std::string& name, uint32_t num_attributes,
                                                         locationT location = LOCATION_ALL, layoutT layout = AoSoA)
    {
        return add<FaceAttribute<T>>(name, num_attributes, location, layout);
    }
    // add_attribute<T, HandleT>: generic over the element type
    template <class T, class HandleT>
    std::shared_ptr<Attribute<T, HandleT>> add_attribute(const std::string& name, uint32_t num_attributes,
                                                         locationT location = LOCATION_ALL, layoutT layout = AoSoA)
    {
        return add<Attribute<T, HandleT>>(name, num_attributes, location, layout);
    }
    // add_vertex_attribute(Verts, name): filled from per-vertex input in GLOBAL order (rxmesh_static.inl:147-189)
    template <class T>
    std::shared_ptr<VertexAttribute<T>> add_vertex_attribute(const std::vector<std::vector<T>>& values, const std::string& name,
                                                             layoutT layout = AoSoA)
    {
        const uint32_t n = values.empty() ? 0 : (uint32_t)values[0].size();
        auto           a = add<VertexAttribute<T>>(name, n, LOCATION_ALL, layout);
        std::vector<T> flat;
        flat.reserve(values.size() * n);
        for (const auto& v : values)
            flat.insert(flat.end(), v.begin(), v.end());
        detail::rxm_check(rxm_attr_upload_global(a->c_handle(), flat.data(), nullptr));
        detail::rxm_check(cudaDeviceSynchronize() == cudaSuccess ? RXM_OK : RXM_ERR_CUDA);
        return a;
    }
    // add_*_attribute(values, name): one value per element in the order given to the constructor (rxmesh_static.h:608-806)
    template <class T>
    std::shared_ptr<VertexAttribute<T>> add_vertex_attribute(const std::vector<T>& values, const std::string& name,
                                                             layoutT layout = AoSoA)
    {
        return add_filled<VertexAttribute<T>>(values.data(), 1, name, layout);
    }
    template <class T>
    std::shared_ptr<FaceAttribute<T>> add_face_attribute(const std::vector<T>& values, const std::string& name, layoutT layout = AoSoA)
    {
        return add_filled<FaceAttribute<T>>(values.data(), 1, name, layout);
    }
    template <class T>
    std::shared_ptr<FaceAttribute<T>> add_face_attribute(const std::vector<std::vector<T>>& values, const std::string& name,
                                                         layoutT layout = AoSoA)
    {
        const uint32_t n = values.empty() ? 0 : (uint32_t)values[0].size();
        std::vector<T> flat;
        flat.reserve(values.size() * n);
        for (const auto& v : values)
            flat.insert(flat.end(), v.begin(), v.end());
        return add_filled<FaceAttribute<T>>(flat.data(), n, name, layout);
    }
    // add_*_attribute_like (rxmesh_static.h:643-806): same allocation, number of attributes and layout as `other`
    template <class T>
    std::shared_ptr<VertexAttribute<T>> add_vertex_attribute_like(const std::string& name, const VertexAttribute<T>& other)
    {
        return add<VertexAttribute<T>>(name, other.get_num_attributes(), other.get_allocated(), other.get_layout());
    }
    template <class T>
    std::shared_ptr<EdgeAttribute<T>> add_edge_attribute_like(const std::string& name, const EdgeAttribute<T>& other)
    {
        return add<EdgeAttribute<T>>(name, other.get_num_attributes(), other.get_allocated(), other.get_layout());
    }
    template <class T>
    std::shared_ptr<FaceAttribute<T>> add_face_attribute_like(const std::string& name, const FaceAttribute<T>& other)
    {
        return add<FaceAttribute<T>>(name, other.get_num_attributes(), other.get_allocated(), other.get_layout());
    }
    template <class T, class HandleT>
    std::shared_ptr<Attribute<T, HandleT>> add_attribute_like(const std::string& name, const Attribute<T, HandleT>& other)
    {
        return add<Attribute<T, HandleT>>(name, other.get_num_attributes(), other.get_allocated(), other.get_layout());
    }
    std::vector<std::string> get_attribute_names() const
    {
        std::vector<std::string> names;
        for (const auto& kv : m_attrs)
            names.push_back(kv.first);
        return names;
    }
    bool does_attribute_exist(const std::string& name) const { return m_attrs.count(name) != 0; }
    void remove_attribute(const std::string& name) { m_attrs.erase(name); }

    // ---- id maps ----
    template <typename HandleT>
    uint32_t map_to_global(const HandleT h) const  // rxmesh_static.cu:669-685
    {
        return rxm_mesh_slot_to_global(m_mesh, HandleT::elem)[rxm_mesh_slot_base(m_mesh, HandleT::elem)[h.patch_id()] + h.local_id()];
    }
    template <typename HandleT>
    uint32_t linear_id(const HandleT h) const  // context.h:275-290
    {
        return rxm_mesh_lin_base(m_mesh, HandleT::elem)[h.patch_id()] + h.local_id();
    }

    // get_owner_handle (rxmesh_static.h:905, context.h:219-270) on the host: a local (possibly not-owned) element of a
    // patch -> the handle of its owner
    template <typename HandleT>
    HandleT get_owner_handle(const HandleT input) const
    {
        rxm_patch_view v;
        detail::rxm_check(rxm_mesh_patch(m_mesh, input.patch_id(), &v));
        const uint32_t t = HandleT::elem, lid = input.local_id();
        if (lid < v.n_owned[t]) return input;
        const uint32_t o = v.owner[t][lid - v.n_owned[t]];
        return HandleT(v.stash[4 * (o >> 16)], typename HandleT::LocalT((uint16_t)(o & 0xFFFFu)));
    }
    // get_boundary_vertices (rxmesh_static.h:816-819, kernels/boundary.cuh:11-44): boundary_v(vh) = 1 on the boundary,
    // 0 elsewhere, for any single-component attribute type (the reference's test uses VertexAttribute<bool>)
    template <typename T>
    void get_boundary_vertices(VertexAttribute<T>& boundary_v, bool move_to_host = true, cudaStream_t stream = NULL)
    {
        if constexpr (sizeof(T) == 4) {
            detail::rxm_check(rxm_boundary_vertices(m_mesh, boundary_v.c_handle(), stream));
            if (!std::is_integral_v<T>)  // the kernel writes integer flags; convert in place for 32-bit floating point
                detail::flags_to_value<T><<<((uint32_t)boundary_v.storage_size() + 255) / 256, 256, 0, stream>>>(boundary_v.data(DEVICE), (uint32_t)boundary_v.storage_size());
        } else {
            // the kernel writes 32-bit integer flags: run it on a 32-bit attribute of the same layout, convert per vertex
            const std::string tmp_name = std::string("rx:boundary_flags_") + boundary_v.get_name();
            auto              flags    = add_vertex_attribute<int>(tmp_name, 1, DEVICE, boundary_v.get_layout());
            detail::rxm_check(rxm_boundary_vertices(m_mesh, flags->c_handle(), stream));
            auto f = *flags;
            auto b = boundary_v;
            for_each_vertex(DEVICE, [f, b] __device__(const VertexHandle vh) { b(vh) = f(vh) ? T(1) : T(0); }, stream);
            detail::rxm_check(cudaStreamSynchronize(stream) == cudaSuccess ? RXM_OK : RXM_ERR_CUDA);
            remove_attribute(tmp_name);
        }
        if (move_to_host) boundary_v.move(DEVICE, HOST, stream);
    }
    // export_obj (rxmesh_static.inl:365-397): vertices in linear-id order, faces in linear-id order with 1-based
    // linear vertex ids
    template <typename T>
    void export_obj(const std::string& filename, const VertexAttribute<T>& coords) const
    {
        std::fstream file(filename, std::ios::out);
        file.precision(30);
        std::vector<glm::vec3> v_list;
        create_vertex_list(v_list, coords);
        for (uint32_t v = 0; v < v_list.size(); ++v)
            file << "v " << v_list[v][0] << " " << v_list[v][1] << " " << v_list[v][2] << " \n";
        std::vector<glm::uvec3> f_list;
        create_face_list(f_list);
        for (uint32_t f = 0; f < f_list.size(); ++f)
            file << "f " << f_list[f][0] + 1 << " " << f_list[f][1] + 1 << " " << f_list[f][2] + 1 << " \n";
    }
    // export_vtk (rxmesh_static.h:934-1076): legacy ASCII VTK polydata, vertex and face attributes with 1, 2 or 3
    // components as SCALARS / COLOR_SCALARS / VECTORS in for_each order
    template <typename T, typename... AttributesT>
    void export_vtk(const std::string& filename, const VertexAttribute<T>& coords, AttributesT... attributes) const
    {
        std::fstream file(filename, std::ios::out);
        file.precision(30);
        std::string name = filename.substr(0, filename.find_last_of('.'));  // extract_file_name (util/util.h:334-341)
        name             = name.substr(name.find_last_of("/\\") + 1);
        file << "# vtk DataFile Version 3.0\n" << name << "\nASCII\nDATASET POLYDATA\n";
        file << "POINTS " << get_num_vertices() << " float\n ";
        std::vector<glm::vec3> v_list;
        create_vertex_list(v_list, coords);
        for (uint32_t v = 0; v < v_list.size(); ++v)
            file << v_list[v][0] << " " << v_list[v][1] << " " << v_list[v][2] << " \n";
        std::vector<glm::uvec3> f_list;
        create_face_list(f_list);
        file << "POLYGONS 3 " << 4 * f_list.size() << "\n";
        for (uint32_t f = 0; f < f_list.size(); ++f)
            file << "3 " << f_list[f][0] << " " << f_list[f][1] << " " << f_list[f][2] << " \n";
        bool first_v_attr = true, first_f_attr = true;
        ([&] { export_vtk_attribute(file, first_v_attr, first_f_attr, attributes); }(), ...);
    }
    // coordinates in linear-id order (rxmesh_static.inl:399-414)
    template <typename T>
    void create_vertex_list(std::vector<glm::vec3>& v_list, const VertexAttribute<T>& coords) const
    {
        v_list.resize(get_num_vertices());
        for_each_vertex(HOST, [&](const VertexHandle vh) {
            for (int i = 0; i < 3; ++i) v_list[linear_id(vh)][i] = (float)coords(vh, i);
        }, NULL, false);
    }
    // faces in linear-id order over linear vertex ids
    void create_face_list(std::vector<glm::uvec3>& f_list) const
    {
        const uint32_t* g2s = rxm_mesh_global_to_slot(m_mesh, RXM_V);
        const uint32_t* ep  = rxm_mesh_elem_patch(m_mesh, RXM_V);
        const uint32_t* sb  = rxm_mesh_slot_base(m_mesh, RXM_V);
        const uint32_t* lb  = rxm_mesh_lin_base(m_mesh, RXM_V);
        f_list.resize(get_num_faces());
        for_each_face(HOST, [&](const FaceHandle fh) {
            const uint32_t g = map_to_global(fh);
            for (int i = 0; i < 3; ++i) {
                const uint32_t gv = m_fv[3 * (size_t)g + i];
                f_list[linear_id(fh)][i] = lb[ep[gv]] + (g2s[gv] - sb[ep[gv]]);
            }
        }, NULL, false);
    }
    template <typename AttributeT>
    void export_vtk_attribute(std::fstream& file, bool& first_v_attr, bool& first_f_attr, const AttributeT& attribute) const
    {
        using HandleT = typename AttributeT::HandleType;
        static_assert(std::is_same_v<HandleT, FaceHandle> || std::is_same_v<HandleT, VertexHandle>,
                      "export_vtk supports vertex and face attributes (edge attributes are NOT supported)");
        bool& first = std::is_same_v<HandleT, FaceHandle> ? first_f_attr : first_v_attr;
        if (first) {
            if (std::is_same_v<HandleT, FaceHandle>)
                file << "CELL_DATA " << get_num_faces() << "\n";
            else
                file << "POINT_DATA " << get_num_vertices() << "\n";
            first = false;
        }
        const uint32_t num_attr = attribute.get_num_attributes();
        if (num_attr == 1)
            file << "SCALARS " << attribute.get_name() << " float 1\nLOOKUP_TABLE default\n";
        else if (num_attr == 2)
            file << "COLOR_SCALARS " << attribute.get_name() << " 2\n";
        else if (num_attr == 3)
            file << "VECTORS " << attribute.get_name() << " float \n";
        else {
            fprintf(stderr, "RXMeshStatic::export_vtk() The number of attributes (%u) is not support. Only 1, 2, or 3 attributes are supported\n", num_attr);
            return;
        }
        for_each<HandleT>(HOST, [&](const HandleT& h) {
            for (uint32_t i = 0; i < num_attr; ++i) file << attribute(h, i) << " ";
            file << "\n";
        }, NULL, false);
    }

    // ---- run_kernel (rxmesh_static.h:415-505) ----
    template <uint32_t blockThreads, typename KernelT, typename... ArgsT>
    void run_kernel(const LaunchBox<blockThreads>& lb, const KernelT kernel, cudaStream_t stream, ArgsT... args) const
    {
        kernel<<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn, stream>>>(get_context(), args...);
    }
    template <uint32_t blockThreads, typename KernelT, typename... ArgsT>
    void run_kernel(const LaunchBox<blockThreads>& lb, const KernelT kernel, ArgsT... args) const
    {
        kernel<<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn>>>(get_context(), args...);
    }
    template <uint32_t blockThreads, typename KernelT, typename... ArgsT>
    void run_kernel(const std::vector<Op> op, KernelT kernel, ArgsT... args) const
    {
        run_kernel<blockThreads>(kernel, op, false, false, false,
                                 [](uint32_t, uint32_t, uint32_t) -> size_t { return 0; }, NULL, args...);
    }
    template <uint32_t blockThreads, typename KernelT, typename... ArgsT>
    void run_kernel(KernelT kernel, const std::vector<Op> op, const bool oriented, const bool with_vertex_valence,
                    const bool is_concurrent, std::function<size_t(uint32_t, uint32_t, uint32_t)> user_shmem,
                    cudaStream_t stream, ArgsT... args) const
    {
        LaunchBox<blockThreads> lb;
        prepare_launch_box(op, lb, (void*)kernel, oriented, with_vertex_valence, is_concurrent, user_shmem);
        run_kernel(lb, kernel, stream, args...);
    }

    // ---- for_each_vertex / edge / face (rxmesh_static.h:205-379) ----
    template <typename LambdaT>
    void for_each_vertex(locationT location, LambdaT apply, cudaStream_t stream = NULL, bool with_omp = true) const
    {
        for_each_elem<VertexHandle>(location, apply, stream, with_omp);
    }
    template <typename LambdaT>
    void for_each_edge(locationT location, LambdaT apply, cudaStream_t stream = NULL, bool with_omp = true) const
    {
        for_each_elem<EdgeHandle>(location, apply, stream, with_omp);
    }
    template <typename LambdaT>
    void for_each_face(locationT location, LambdaT apply, cudaStream_t stream = NULL, bool with_omp = true) const
    {
        for_each_elem<FaceHandle>(location, apply, stream, with_omp);
    }

    // for_each<HandleT>(location, lambda) (rxmesh_static.h:387-410) and get_num_elements<HandleT>() (rxmesh.h)
    template <typename HandleT, typename LambdaT>
    void for_each(locationT location, LambdaT apply, cudaStream_t stream = NULL, bool with_omp = true) const
    {
        for_each_elem<HandleT>(location, apply, stream, with_omp);
    }
    template <typename HandleT>
    uint32_t get_num_elements() const
    {
        return info(RXM_INFO_NUM_VERTICES + HandleT::elem);
    }

    // ---- for_each<Op, blockThreads>(lambda) (rxmesh_static.h:524-566) ----
    template <Op op, uint32_t blockThreads, typename LambdaT>
    void for_each(const LambdaT user_lambda, const bool oriented = false, cudaStream_t stream = NULL) const
    {
        LaunchBox<blockThreads> lb;
        prepare_launch_box({op}, lb, (void*)detail::query_kernel<blockThreads, op, LambdaT>, oriented);
        detail::query_kernel<blockThreads, op><<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn, stream>>>(m_context, oriented, user_lambda);
        detail::check_launch("for_each<Op>");
    }

    // ---- prepare_launch_box (rxmesh_static.inl:443-496) ----
    template <uint32_t blockThreads>
    void prepare_launch_box(const std::vector<Op> op, LaunchBox<blockThreads>& launch_box, const void* kernel,
                            const bool oriented = false, const bool with_vertex_valence = false,
                            const bool is_concurrent = false,
                            std::function<size_t(uint32_t, uint32_t, uint32_t)> user_shmem =
                                [](uint32_t, uint32_t, uint32_t) { return 0; }) const
    {
        if (oriented && !info(RXM_INFO_FANS))
            for (Op o : op)
                if (o == Op::VV || o == Op::VE)
                    fprintf(stderr, "RXMeshStatic::prepare_launch_box() oriented %s needs an edge-manifold, consistently oriented input "
                                    "mesh: this mesh stores no one-ring fans and the kernel will stop\n", op_to_string(o).c_str());
        size_t   dyn = with_vertex_valence ? 4 * (size_t)get_per_patch_max_vertices() + 16 : 0;  // compute_vertex_valence
        dyn += 4 * ((size_t)std::max(get_per_patch_max_edges(), get_per_patch_max_faces()) / 32 + 8);  // prologue mask
        uint32_t blocks = get_num_patches(), threads = 0;
        for (Op o : op) {
            uint32_t b = 0;
            if (o == Op::V || o == Op::E || o == Op::F) continue;  // device for_each<Op::V|E|F>: one block per patch, no staging
            detail::rxm_check(rxm_mesh_launch_box(m_mesh, (int)o, &blocks, &threads, &b));
            dyn = is_concurrent ? dyn + b : std::max<size_t>(dyn, b);
        }
        dyn += user_shmem(get_per_patch_max_vertices(), get_per_patch_max_edges(), get_per_patch_max_faces());
        launch_box.blocks         = blocks;
        launch_box.smem_bytes_dyn = dyn;
        cudaFuncAttributes fa{};
        if (kernel && cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) {
            launch_box.smem_bytes_static        = fa.sharedSizeBytes;
            launch_box.num_registers_per_thread = (uint32_t)fa.numRegs;
            launch_box.local_mem_per_thread     = fa.localSizeBytes;
            if (dyn > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        }
    }

   private:
    void init(const uint32_t* fv, uint32_t nf, const std::vector<uint32_t>& face_patch, uint32_t patch_size)
    {
        m_fv.assign(fv, fv + 3 * (size_t)nf);  // kept for export_obj
        detail::rxm_check(rxm_mesh_create(fv, nf, face_patch.empty() ? nullptr : face_patch.data(), patch_size, 0, &m_mesh));
        detail::rxm_check(rxm_mesh_to_device(m_mesh));
        detail::rxm_check(rxm_mesh_view(m_mesh, &m_context.view, (uint32_t)sizeof(rxm::MeshView)));
    }
    uint32_t info(int k) const { return (uint32_t)rxm_mesh_info(m_mesh, k); }
    rxm_patch_view patch_view(uint32_t p) const
    {
        rxm_patch_view v;
        detail::rxm_check(rxm_mesh_patch(m_mesh, p, &v));
        return v;
    }
    template <typename HandleT>
    HandleT map_to_local(uint32_t i) const
    {
        const uint32_t* lb  = rxm_mesh_lin_base(m_mesh, HandleT::elem);
        const uint32_t* end = lb + get_num_patches() + 1;
        const uint32_t* p   = std::upper_bound(lb, end, i);  // first prefix past i
        if (p == end || p == lb) {
            fprintf(stderr, "RXMeshStatic::map_to_local can not its patch. Input is out of range!\n");
            return HandleT();
        }
        --p;
        return HandleT((uint32_t)(p - lb), typename HandleT::LocalT((uint16_t)(i - *p)));
    }
    static std::vector<uint32_t> flatten(const std::vector<std::vector<uint32_t>>& fv)
    {
        std::vector<uint32_t> flat;
        flat.reserve(3 * fv.size());
        for (const auto& f : fv) {
            if (f.size() != 3) {  // rxmesh.cpp:590-597
                fprintf(stderr, "rxmesh_b200: non-triangular faces are not supported\n");
                exit(EXIT_FAILURE);
            }
            flat.insert(flat.end(), f.begin(), f.end());
        }
        return flat;
    }

    template <typename AttrT, typename T>
    std::shared_ptr<AttrT> add_filled(const T* flat_global, uint32_t n, const std::string& name, layoutT layout)
    {
        auto a = add<AttrT>(name, n, LOCATION_ALL, layout);
        detail::rxm_check(rxm_attr_upload_global(a->c_handle(), flat_global, nullptr));
        detail::rxm_check(cudaDeviceSynchronize() == cudaSuccess ? RXM_OK : RXM_ERR_CUDA);
        return a;
    }
    // append semantics of import_obj (util/import_obj.h): positions and faces of `path` are added to what the vectors hold
    static void read_obj(const std::string& path, std::vector<std::vector<float>>& verts, std::vector<std::vector<uint32_t>>& faces)
    {
        if (!import_obj(path, verts, faces, true)) {
            fprintf(stderr, "RXMeshStatic::RXMeshStatic could not read the input file %s\n", path.c_str());
            exit(EXIT_FAILURE);
        }
    }
    static std::vector<std::vector<uint32_t>> read_obj_faces(const std::string& path)
    {
        std::vector<std::vector<float>>    v;
        std::vector<std::vector<uint32_t>> f;
        read_obj(path, v, f);
        return f;
    }

    template <typename AttrT>
    std::shared_ptr<AttrT> add(const std::string& name, uint32_t n, locationT location, layoutT layout)
    {
        if (m_attrs.count(name)) {
            fprintf(stderr, "rxmesh_b200: attribute %s already exists\n", name.c_str());  // RXMESH_ERROR: log, continue
            return std::dynamic_pointer_cast<AttrT>(m_attrs[name]);
        }
        // the container owns the storage (AttributeContainer, attribute.h:676-731): released when the last
        // shared_ptr goes, i.e. at remove_attribute / ~RXMeshStatic unless the user still holds one
        std::shared_ptr<AttrT> a(new AttrT(m_mesh, name.c_str(), n, location, layout), [](AttrT* p) {
            p->release();
            delete p;
        });
        m_attrs[name] = a;
        return a;
    }

    template <typename HandleT, typename LambdaT>
    void for_each_elem(locationT location, LambdaT apply, cudaStream_t stream, bool with_omp) const
    {
        constexpr bool is_d  = __nv_is_extended_device_lambda_closure_type(LambdaT);
        constexpr bool is_hd = __nv_is_extended_host_device_lambda_closure_type(LambdaT);
        if ((location & HOST) == HOST) {
            if constexpr (!is_d) {
                const uint32_t* lb = rxm_mesh_lin_base(m_mesh, HandleT::elem);
                const int       P  = (int)get_num_patches();
#pragma omp parallel for if (with_omp)
                for (int p = 0; p < P; ++p)
                    for (uint32_t i = 0; i < lb[p + 1] - lb[p]; ++i)
                        apply(HandleT((uint32_t)p, typename HandleT::LocalT((uint16_t)i)));
            }
        }
        if ((location & DEVICE) == DEVICE) {
            if constexpr (is_d || is_hd) {
                detail::for_each_kernel<HandleT><<<get_num_patches(), 256, 0, stream>>>(m_context, apply);
                detail::check_launch("for_each_vertex/edge/face");
            } else {
                fprintf(stderr, "RXMeshStatic::for_each_*() Input lambda function should be annotated with __device__ "
                                "for execution on device\n");
            }
        }
    }

    rxm_mesh*                                             m_mesh = nullptr;
    std::vector<uint32_t>                                 m_fv;
    std::shared_ptr<VertexAttribute<float>>               m_input_coords;
    mutable std::unordered_map<uint64_t, uint32_t>        m_edge_ids;  // get_edge_id, built on first use
    int                                                   m_num_regions = 1;
    std::shared_ptr<FaceAttribute<int>>                   m_face_label;
    std::shared_ptr<EdgeAttribute<int>>                   m_edge_label;
    std::shared_ptr<VertexAttribute<int>>                 m_vertex_label;
    Context                                               m_context;
    std::map<std::string, std::shared_ptr<AttributeBase>> m_attrs;
};
}  // namespace rxmesh
