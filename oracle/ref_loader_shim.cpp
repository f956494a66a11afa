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

// The RAW arrays tinyobj::LoadObj hands to loadFromFile (main.cpp:34-36), before lines :37-57 run: attrib.vertices,
// the vertex_index of every face corner (all shapes, in order), the material id of every face and {Kd, Ke} of every
// material. Same two-call protocol. Pins the input of the device-side front-end (bpt_upload_obj_arrays).
int ref_load_obj_raw(const char* obj_path, const char* mtl_dir, float* positions, int32_t* corner_vertex, int32_t* face_material,
                     float* materials_kd_ke, uint32_t* npositions, uint32_t* ncorners, uint32_t* nfaces, uint32_t* nmaterials) {
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string warn, e;
    if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &e, obj_path, mtl_dir)) return -1;
    if (positions) std::memcpy(positions, attrib.vertices.data(), attrib.vertices.size() * sizeof(float));
    uint32_t nc = 0, nf = 0;
    for (const auto& shape : shapes) {
        for (const auto& index : shape.mesh.indices) { if (corner_vertex) corner_vertex[nc] = index.vertex_index; ++nc; }
        for (int m : shape.mesh.material_ids) { if (face_material) face_material[nf] = m; ++nf; }
    }
    if (materials_kd_ke)
        for (size_t m = 0; m < materials.size(); ++m)
            for (int c = 0; c < 3; ++c) {
                materials_kd_ke[6 * m + c] = materials[m].diffuse[c];
                materials_kd_ke[6 * m + 3 + c] = materials[m].emission[c];
            }
    *npositions = uint32_t(attrib.vertices.size() / 3); *ncorners = nc; *nfaces = nf; *nmaterials = uint32_t(materials.size());
    return 0;
}

}  // extern "C"
