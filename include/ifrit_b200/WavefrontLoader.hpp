// WavefrontLoader.hpp — header-only mirror of the reference's OBJ loader, same method names and
// results as projects/softgraphics/src/utility/loader/WavefrontLoader.cpp:
//   loadObject   :82-133   "v" / "vn" / "vt" records appended in file order; every face corner of an "f" record (exactly the
//                           first three corners) appends THREE indices (position, texcoord, normal), 0-based; an empty
//                           field ("7//2") is stored as 0, a missing trailing field repeats the one before it ("7" -> 7, 7, 7)
//   RemapNormals :49-81    per position: sum of the normals of the corners that use it, normalised
//   RemapUVs     :24-48    per position: the texcoord of the last corner that uses it
// plus the interleaving the reference's demos do by hand (DefaultDemo.cpp:255-267): float4 position (w = 1) | float4
// normal (w = 0) | float4 uv, and the index buffer = every third loader index.
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace ifrit_b200 { namespace Utility { namespace Loader {

struct Vec2 { float x, y; };
struct Vec3 { float x, y, z; };

class WavefrontLoader {
public:
    static bool loadObject(const char* path, std::vector<Vec3>& vertices, std::vector<Vec3>& normals, std::vector<Vec2>& uvs,
                           std::vector<uint32_t>& indices) {
        std::ifstream file(path);
        if (!file) return false;
        std::string line;
        while (std::getline(file, line)) {
            std::istringstream iss(line);
            std::string type;
            iss >> type;
            if (type == "v") { Vec3 v{0, 0, 0}; iss >> v.x >> v.y >> v.z; vertices.push_back(v); }
            else if (type == "vn") { Vec3 n{0, 0, 0}; iss >> n.x >> n.y >> n.z; normals.push_back(n); }
            else if (type == "vt") { Vec2 t{0, 0}; iss >> t.x >> t.y; uvs.push_back(t); }
            else if (type == "f") {
                std::string corner;
                for (int i = 0; i < 3; i++) {
                    iss >> corner;                       // a missing corner keeps the previous one, as in the reference
                    std::istringstream vss(corner);
                    std::string field;
                    for (int j = 0; j < 3; j++) {
                        // getline on an exhausted stream leaves `field` unchanged: "7" -> (7, 7, 7), "7/3" -> (7, 3, 3);
                        // only an empty field ("7//2") is stored as 0 (WavefrontLoader.cpp:117-128)
                        std::getline(vss, field, '/');
                        indices.push_back(field.empty() ? 0u : (uint32_t)(std::stoi(field) - 1));
                    }
                }
            }
        }
        return true;
    }
    static std::vector<Vec3> RemapNormals(const std::vector<Vec3>& normals, const std::vector<uint32_t>& indices, int numVertices) {
        std::vector<Vec3> out((size_t)numVertices, Vec3{0, 0, 0});
        for (size_t i = 0; i + 2 < indices.size(); i += 3) {
            Vec3& d = out[indices[i]];
            const Vec3& n = normals[indices[i + 2]];
            d.x += n.x; d.y += n.y; d.z += n.z;
        }
        for (Vec3& d : out) {
            const float len = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
            d.x /= len; d.y /= len; d.z /= len;
        }
        return out;
    }
    static std::vector<Vec2> RemapUVs(const std::vector<Vec2>& uvs, const std::vector<uint32_t>& indices, int numVertices) {
        std::vector<Vec2> out((size_t)numVertices, Vec2{0, 0});
        for (size_t i = 0; i + 2 < indices.size(); i += 3) out[indices[i]] = uvs[indices[i + 1]];
        return out;
    }
    // interleaved FLOAT4 x 3 vertex buffer + int index buffer, ready for bindVertexBuffer / bindIndexBuffer
    static bool loadInterleaved(const char* path, std::vector<float>& vertexBuffer, std::vector<int>& indexBuffer) {
        std::vector<Vec3> pos, nrm;
        std::vector<Vec2> uv;
        std::vector<uint32_t> idx;
        if (!loadObject(path, pos, nrm, uv, idx)) return false;
        const std::vector<Vec3> pn = nrm.empty() ? std::vector<Vec3>(pos.size(), Vec3{0, 0, 0}) : RemapNormals(nrm, idx, (int)pos.size());
        const std::vector<Vec2> pu = uv.empty() ? std::vector<Vec2>(pos.size(), Vec2{0, 0}) : RemapUVs(uv, idx, (int)pos.size());
        vertexBuffer.resize(pos.size() * 12);
        for (size_t i = 0; i < pos.size(); i++) {
            float* v = &vertexBuffer[i * 12];
            v[0] = pos[i].x; v[1] = pos[i].y; v[2] = pos[i].z; v[3] = 1.0f;
            v[4] = pn[i].x; v[5] = pn[i].y; v[6] = pn[i].z; v[7] = 0.0f;
            v[8] = pu[i].x; v[9] = pu[i].y; v[10] = 0.0f; v[11] = 0.0f;
        }
        indexBuffer.resize(idx.size() / 3);
        for (size_t i = 0; i < idx.size(); i += 3) indexBuffer[i / 3] = (int)idx[i];
        return true;
    }
};

}}}   // namespace ifrit_b200::Utility::Loader
