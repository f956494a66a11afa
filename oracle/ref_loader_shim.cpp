// ref_loader_shim.cpp — thin C entry point over the reference's OWN vendored tinyobjloader.
//
// TEST INFRASTRUCTURE ONLY (same rules as oracle.cpp). This file contains no reference code:
// it #includes <tiny_obj_loader.h> from /root/reference/external/tinyobjloader *where it
// lies* (see oracle/Makefile, target _ref) and applies the loader semantics of the
// reference's loadFromFile (main.cpp:28-58): negate Y, de-index (one vertex per index,
// identity index buffer), one {Kd, Ke} record per triangle. The built library lives in
// oracle/_ref/ (git-ignored) and is used to pin the scene input: tests/golden/make_golden.py
// writes its output to tests/golden/cornell_scene.json, and tests compare the product's own
// OBJ loader and the fixture against it.
#define TINYOBJLOADER_IMPLEMENTATION
#include <tiny_obj_loader.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

// Two-call protocol: pass NULL buffers to get counts, then call again with storage.
// Returns 0 on success, -1 on load failure (message copied into err if non-NULL).
int ref_load_obj(const char* obj_path, const char* mtl_dir, float* verts, uint32_t* indices, float* faces,
                 uint32_t* nverts, uint32_t* nindices, uint32_t* nfaces, uint32_t* shape_tris, uint32_t* nshapes,
                 char* err, uint32_t errlen) {
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string warn, e;
    if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &e, obj_path, mtl_dir)) {
        if (err && errlen) { std::strncpy(err, (warn + e).c_str(), errlen - 1); err[errlen - 1] = 0; }
        return -1;
    }
    uint32_t nv = 0, nf = 0, ns = 0;
    for (const auto& shape : shapes) {
        for (const auto& index : shape.mesh.indices) {
            if (verts) {
                verts[3 * size_t(nv) + 0] = attrib.vertices[3 * index.vertex_index + 0];
                verts[3 * size_t(nv) + 1] = -attrib.vertices[3 * index.vertex_index + 1];
                verts[3 * size_t(nv) + 2] = attrib.vertices[3 * index.vertex_index + 2];
            }
            if (indices) indices[nv] = nv;
            ++nv;
        }
        for (int m : shape.mesh.material_ids) {
            if (faces) {
                for (int c = 0; c < 3; ++c) {
                    faces[6 * size_t(nf) + c] = materials[m].diffuse[c];
                    faces[6 * size_t(nf) + 3 + c] = materials[m].emission[c];
                }
            }
            ++nf;
        }
        if (shape_tris) shape_tris[ns] = uint32_t(shape.mesh.material_ids.size());
        ++ns;
    }
    *nverts = nv; *nindices = nv; *nfaces = nf;
    if (nshapes) *nshapes = ns;
    return 0;
}

}  // extern "C"
