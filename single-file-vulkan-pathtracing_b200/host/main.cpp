// main.cpp — C++20 host application over the C ABI (include/bpt.h), shaped like the reference's main()
// (reference main.cpp:457-690): open the GLFW window, load the scene, upload it, build the acceleration structure, then
// the frame loop `while (!glfwWindowShouldClose) { glfwPollEvents(); push frame; trace; present; }`. What the reference
// does with Vulkan objects between main.cpp:496 and :641 (BLAS/TLAS descriptions, pipeline, SBT, descriptor set) has
// no counterpart here: it is bpt_upload_mesh + bpt_build_accel.
//
// Window: the reference's GLFW calls are kept one for one (glfwInit, the two window hints, glfwCreateWindow,
// glfwWindowShouldClose, glfwPollEvents, glfwDestroyWindow, glfwTerminate: main.cpp:77-80, 647-648, 688-689), built
// from the GLFW the reference vendors (external/glfw, compiled where it lies: see Makefile). The target machines have
// no display server, so the window lives on GLFW's null platform, which cannot create a Vulkan surface
// (external/glfw/src/null_window.c): "present" (main.cpp:661-682) hands the B8G8R8A8 bytes the reference's storage
// image would hold to present_frame(), which keeps the latest frame and writes it to a PPM file at exit (the float4
// accumulation goes to a PFM file). The read-back of frame f runs on the context's copy stream while frame f+1 is
// traced (bpt_read_image_bgra8_async), so the loop never stalls on the copy the way queue.waitIdle() (main.cpp:683) does.
// `--frames N` closes the window after N frames (glfwSetWindowShouldClose): the headless exit.
//
//   bpt_host --obj ../assets/CornellBox-Original.obj [--frames 8] [--width 1024 --height 1024 --spp 32 --depth 8]
//            [--rgba8-feedback] [--out out]            (tinyobjloader build only, see Makefile)
//   bpt_host --scene scene.bin ...                     (raw arrays: what loadFromFile produces, see read_scene_bin)
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "bpt.h"

#ifdef BPT_HOST_GLFW
#define GLFW_INCLUDE_NONE  // the reference includes vulkan.hpp first, which has the same effect (main.cpp:8-9)
#include <GLFW/glfw3.h>
#endif

#ifdef BPT_HOST_TINYOBJ
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"
#endif

namespace {

// same records as the reference (main.cpp:19-26): 12-byte vertices, 24-byte faces
struct Vertex { float position[3]; };
struct Face { float diffuse[3]; float emission[3]; };

struct Scene {
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<Face> faces;
};

#ifdef BPT_HOST_TINYOBJ
// Restated from the reference's loadFromFile (main.cpp:28-58), which SURVEY C12 keeps as the input contract:
// tinyobj::LoadObj, negate Y, one vertex per index (the index buffer becomes 0,1,2,...), one {Kd, Ke} record per
// triangle.
Scene load_obj(const std::string& path) {
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string warn, err;
    const std::string dir = path.substr(0, path.find_last_of("/\\") + 1);
    if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &warn, &err, path.c_str(), dir.c_str()))
        throw std::runtime_error("failed to load " + path + ": " + warn + err);
    Scene s;
    for (const auto& shape : shapes) {
        for (const auto& index : shape.mesh.indices) {
            Vertex v{};
            v.position[0] = attrib.vertices[3 * index.vertex_index + 0];
            v.position[1] = -attrib.vertices[3 * index.vertex_index + 1];
            v.position[2] = attrib.vertices[3 * index.vertex_index + 2];
            s.vertices.push_back(v);
            s.indices.push_back(static_cast<uint32_t>(s.indices.size()));
        }
        for (const auto& mat_id : shape.mesh.material_ids) {
            if (mat_id < 0 || mat_id >= static_cast<int>(materials.size()))
                throw std::runtime_error("face without a material in " + path);
            Face f{};
            for (int k = 0; k < 3; ++k) {
                f.diffuse[k] = materials[mat_id].diffuse[k];
                f.emission[k] = materials[mat_id].emission[k];
            }
            s.faces.push_back(f);
        }
    }
    return s;
}
#endif

// scene.bin: "BPTSCN1\0", u32 nverts, u32 nindices, u32 nfaces, then the three arrays
Scene read_scene_bin(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("failed to open " + path);
    char magic[8];
    uint32_t n[3];
    in.read(magic, 8);
    in.read(reinterpret_cast<char*>(n), sizeof(n));
    if (!in || std::memcmp(magic, "BPTSCN1", 8) != 0) throw std::runtime_error(path + " is not a BPTSCN1 file");
    Scene s;
    s.vertices.resize(n[0]); s.indices.resize(n[1]); s.faces.resize(n[2]);
    in.read(reinterpret_cast<char*>(s.vertices.data()), sizeof(Vertex) * n[0]);
    in.read(reinterpret_cast<char*>(s.indices.data()), sizeof(uint32_t) * n[1]);
    in.read(reinterpret_cast<char*>(s.faces.data()), sizeof(Face) * n[2]);
    if (!in) throw std::runtime_error(path + " is truncated");
    return s;
}

void check(bpt_context* c, int rc) {  // the reference throws std::runtime_error on every failed call
    if (rc != BPT_OK) throw std::runtime_error(std::string("bpt: ") + bpt_last_error(c));
}

void write_ppm_from_bgra(const std::string& path, const std::vector<uint8_t>& bgra, uint32_t w, uint32_t h) {
    std::ofstream out(path, std::ios::binary);
    out << "P6\n" << w << " " << h << "\n255\n";
    for (size_t i = 0; i < size_t(w) * h; ++i) {
        const char rgb[3] = {char(bgra[4 * i + 2]), char(bgra[4 * i + 1]), char(bgra[4 * i + 0])};
        out.write(rgb, 3);
    }
}
void write_pfm(const std::string& path, const std::vector<float>& rgba, uint32_t w, uint32_t h) {
    std::ofstream out(path, std::ios::binary);
    out << "PF\n" << w << " " << h << "\n-1.0\n";  // little endian, rows bottom-up
    for (uint32_t y = h; y-- > 0;)
        for (uint32_t x = 0; x < w; ++x) out.write(reinterpret_cast<const char*>(&rgba[4 * (size_t(y) * w + x)]), 12);
}

}  // namespace

int main(int argc, char** argv) {
    std::string obj, scene_bin, out = "bpt_out";
    int frames = 8, device = 0;
    bool want_display = false, fused = false;
    bpt_params p;
    bpt_params_default(&p);  // WIDTH/HEIGHT 1024 (main.cpp:16-17), 32 spp, depth 8, camera, sky: the shader constants
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "missing value after %s\n", a.c_str()); std::exit(2); }
            return argv[++i];
        };
        if (a == "--obj") obj = next();
        else if (a == "--scene") scene_bin = next();
        else if (a == "--out") out = next();
        else if (a == "--frames") frames = std::atoi(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--width") p.width = (uint32_t)std::atoi(next());
        else if (a == "--height") p.height = (uint32_t)std::atoi(next());
        else if (a == "--spp") p.spp_per_frame = (uint32_t)std::atoi(next());
        else if (a == "--depth") p.max_depth = (uint32_t)std::atoi(next());
        else if (a == "--rgba8-feedback") p.accum_mode = BPT_ACCUM_RGBA8;  // raygen.rgen:88-90 on the rgba8 image
        else if (a == "--display") want_display = true;  // let GLFW pick a real platform (none is compiled in here)
        else if (a == "--fused") fused = true;           // BPT_OPT_FUSED_PATHS: one path kernel per sample pass (small frames)
        else if (a == "--help" || a == "-h") {
            std::printf("usage: %s (--obj file.obj | --scene scene.bin) [--frames N] [--width W --height H --spp S --depth D]\n"
                        "          [--rgba8-feedback] [--fused] [--device I] [--out prefix] [--display]\n", argv[0]);
            return 0;
        } else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    try {
        Scene scene;
        if (!scene_bin.empty()) scene = read_scene_bin(scene_bin);
        else {
#ifdef BPT_HOST_TINYOBJ
            scene = load_obj(obj.empty() ? "../assets/CornellBox-Original.obj" : obj);  // main.cpp:34
#else
            throw std::runtime_error("built without tinyobjloader: pass --scene scene.bin");
#endif
        }
        std::printf("scene: %zu vertices, %zu triangles\n", scene.vertices.size(), scene.faces.size());

        bpt_context* pt = nullptr;
        check(nullptr, bpt_create(device, nullptr, &pt));  // Context (main.cpp:458): device + one stream
        // Buffer vertexBuffer / indexBuffer / faceBuffer (main.cpp:492-494)
        check(pt, bpt_upload_mesh(pt, &scene.vertices[0].position[0], (uint32_t)scene.vertices.size(), scene.indices.data(),
                                  (uint32_t)scene.indices.size(), &scene.faces[0].diffuse[0], (uint32_t)scene.faces.size()));
        check(pt, bpt_build_accel(pt));  // Accel bottomAccel + topAccel (main.cpp:512, :538)
        if (fused) check(pt, bpt_set_option(pt, BPT_OPT_FUSED_PATHS, 1));
        bpt_accel_info info;
        check(pt, bpt_accel_info_get(pt, &info));
        std::printf("accel: %u BVH8 nodes, depth %u, %s\n", info.num_nodes8, info.max_depth8,
                    info.top_nodes_smem ? "staged in shared memory" : "in global memory");

        // ---- window (Context::Context, main.cpp:77-80)
#ifdef BPT_HOST_GLFW
        if (!want_display) glfwInitHint(GLFW_PLATFORM, GLFW_PLATFORM_NULL);  // no display server: the null platform
        if (!glfwInit()) throw std::runtime_error("glfwInit failed");
        glfwWindowHint(GLFW_CLIENT_API, GLFW_NO_API);
        glfwWindowHint(GLFW_RESIZABLE, GLFW_FALSE);
        GLFWwindow* window = glfwCreateWindow((int)p.width, (int)p.height, "B200 Pathtracing", nullptr, nullptr);
        if (!window) throw std::runtime_error("glfwCreateWindow failed");
        std::printf("window: %ux%u on GLFW platform 0x%x%s\n", p.width, p.height, glfwGetPlatform(),
                    glfwGetPlatform() == GLFW_PLATFORM_NULL ? " (null: frames are kept in memory and dumped at exit)" : "");
#endif
        // present_frame: what copyImage -> swapchain + presentKHR (main.cpp:661-682) do with a finished frame
        std::vector<uint8_t> shown(size_t(p.width) * p.height * 4);
        int presented = 0;
        auto present_frame = [&](const std::vector<uint8_t>& f) { shown = f; ++presented; };

        std::vector<uint8_t> bgra[2] = {std::vector<uint8_t>(shown.size()), std::vector<uint8_t>(shown.size())};
        const auto t0 = std::chrono::steady_clock::now();
        int frame = 0;
#ifdef BPT_HOST_GLFW
        while (!glfwWindowShouldClose(window)) {        // main.cpp:647
            glfwPollEvents();                           // main.cpp:648
#else
        for (bool open = frames > 0; open;) {
#endif
            p.frame = frame;                            // pushConstants(frame) (main.cpp:658)
            check(pt, bpt_trace(pt, &p));               // traceRaysKHR(..., WIDTH, HEIGHT, 1) (main.cpp:659)
            if (frame > 0) {                            // frame - 1 was copied out while this frame was enqueued
                check(pt, bpt_read_wait(pt));
                present_frame(bgra[(frame - 1) & 1]);
            }
            check(pt, bpt_read_image_bgra8_async(pt, bgra[frame & 1].data(), bgra[frame & 1].size()));  // copyImage (:661-667)
            ++frame;                                    // main.cpp:684
            if (frame >= frames) {
#ifdef BPT_HOST_GLFW
                glfwSetWindowShouldClose(window, GLFW_TRUE);
#else
                open = false;
#endif
            }
        }
        check(pt, bpt_read_wait(pt));
        if (frame > 0) present_frame(bgra[(frame - 1) & 1]);
        check(pt, bpt_sync(pt));                        // context.device->waitIdle() (main.cpp:687)
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        bpt_stats st;
        check(pt, bpt_get_stats(pt, &st));
        std::printf("%d frames presented, %llu rays in %.3f s (%.1f Mray/s incl. per-frame read-back; device %.1f ms)\n", presented,
                    (unsigned long long)st.rays_traced, secs, st.rays_traced / secs / 1e6, st.frame_ms);
        std::vector<float> rgba(size_t(p.width) * p.height * 4);
        check(pt, bpt_read_image(pt, rgba.data(), rgba.size()));
        write_ppm_from_bgra(out + ".ppm", shown, p.width, p.height);
        write_pfm(out + ".pfm", rgba, p.width, p.height);
        std::printf("wrote %s.ppm and %s.pfm\n", out.c_str(), out.c_str());
        bpt_destroy(pt);
#ifdef BPT_HOST_GLFW
        glfwDestroyWindow(window);                      // main.cpp:688
        glfwTerminate();                                // main.cpp:689
#endif
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
