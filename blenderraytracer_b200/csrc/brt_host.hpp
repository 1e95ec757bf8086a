// Host-side scene model shared by the JSON loader (scene_loader.cpp) and the C ABI (brt_api.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/brt.h"

namespace brt {

struct HostScene {                       // World (reference js/world.js:9-18)
    std::vector<brt_object> objects;     // world.objects, in order: index = object ID
    std::vector<brt_material> materials;
    std::vector<brt_light> lights;       // world.lights
    std::vector<brt_texture> textures;   // textures.js instances referenced by brt_material.texture (1-based)
    std::vector<double> meshTris;        // 9 doubles per mesh triangle
};

struct HostBackground {                  // world.background / skyIntensity / cloudNoise.p (world.js:12-14)
    int kind = BRT_BG_GRADIENT;
    double color[3] = { 0.1, 0.1, 0.1 };
    double intensity = 1.0;
    uint8_t perm[512];
};

// new Camera(...) — reference js/camera.js:8-36, float64.  Fills the derived members of `c` from the raw ones.
void derive_camera(brt_camera& c);

// SceneLoader.loadFromJSON — reference js/scene-loader.js:20-84.  Returns BRT_OK or BRT_E_PARSE (with `err`).
int load_scene_json(const char* utf8, size_t len, int fallbackW, int fallbackH, HostScene& scene, HostBackground& bg,
                    brt_camera& cam, bool& hasCamera, int& outW, int& outH, std::string& err);

// Binary container "BRTSCN01": the same scene JSON with mesh arrays moved into a blob (see scene_loader.cpp).
int load_scene_binary(const unsigned char* bytes, size_t len, int fallbackW, int fallbackH, HostScene& scene, HostBackground& bg,
                      brt_camera& cam, bool& hasCamera, int& outW, int& outH, std::string& err);

}  // namespace brt
