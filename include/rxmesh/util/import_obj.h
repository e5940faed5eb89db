// rxmesh/util/import_obj.h -- import_obj(file, vertices, faces[, append]) (include/rxmesh/util/import_obj.h:26-222).  The
// reference parses with rapidobj (third party); this reader covers what its callers use: "v x y z" positions and
// "f a b c" / "f a/t/n ..." faces with 1-based or negative (relative) indices.  Texture coordinates and normals are skipped
// (the long overload returns them empty); append = true keeps what the vectors hold and offsets the new file's vertex ids.
#pragma once
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "rxmesh/util/log.h"

namespace rxmesh {
template <typename DataT, typename IndexT>
bool import_obj(const std::string file_name, std::vector<std::vector<DataT>>& vertices, std::vector<std::vector<IndexT>>& faces,
                bool append = false)
{
    RXMESH_INFO("Reading {}", file_name);
    if (!append) {
        vertices.clear();
        faces.clear();
    }
    std::ifstream in(file_name);
    if (!in) {
        RXMESH_ERROR("import_obj() can not open {}", file_name);
        return false;
    }
    const long  vertex_offset = (long)vertices.size();
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string        tag;
        ss >> tag;
        if (tag == "v") {
            std::vector<DataT> p(3);
            ss >> p[0] >> p[1] >> p[2];
            vertices.push_back(p);
        } else if (tag == "f") {
            std::vector<IndexT> f;
            std::string         tok;
            while (ss >> tok) {
                const long i = std::stol(tok.substr(0, tok.find('/')));
                f.push_back(i > 0 ? (IndexT)(vertex_offset + i - 1) : (IndexT)((long)vertices.size() + i));
            }
            faces.push_back(f);
        }
    }
    RXMESH_INFO("import_obj() #vertices= {} ", vertices.size());
    RXMESH_INFO("import_obj() #faces= {} ", faces.size());
    return true;
}
template <typename DataT, typename IndexT>
bool import_obj(const std::string file_name, std::vector<std::vector<DataT>>& vertices, std::vector<std::vector<IndexT>>& faces,
                std::vector<std::vector<DataT>>& tex, std::vector<std::vector<IndexT>>& face_tex,
                std::vector<std::vector<DataT>>& normals, std::vector<std::vector<IndexT>>& face_normal, bool append)
{
    if (!append) tex.clear(), face_tex.clear(), normals.clear(), face_normal.clear();
    return import_obj(file_name, vertices, faces, append);
}
}  // namespace rxmesh
