// host_scene.h -- host-side scene store: loads the reference's scene description (XML subset,
// Mitsuba .serialized meshes, OBJ, pre-decoded images), flattens it into the arrays of
// core/scene.h, builds the SAH BVH2 that replaces Embree, and (de)serialises the result as a
// single "scene pack" file.
//
// Mirrors: ParseScene (src/parsescene.cpp:592-639), Scene::Scene (src/scene.cpp:8-46),
// Camera::Camera (src/camera.cpp:11-28), CreateEnvmapSampleInfo (src/envlight.cpp:24-71),
// LoadSerialized (src/loadserialized.cpp:239-325), TriangleMesh::SetAreaLight
// (src/trianglemesh.cpp:293-307), Phong::GetKsWeight (src/phong.cpp:159-169).
#pragma once
#include <string>
#include <vector>
#include "../core/scene.h"

namespace lmc_host {

struct SceneStore {
    std::vector<lmc::TriGeom> tris;
    std::vector<lmc::TriShade> shade;
    std::vector<lmc::BvhNode> nodes;
    std::vector<lmc::Material> mats;
    std::vector<lmc::Texture> textures;
    std::vector<float> texData;
    std::vector<lmc::Light> lights;
    std::vector<float> lightPickCdf;
    std::vector<float> lightCdf;
    std::vector<int> lightPrimTid;
    std::vector<float> envImage, envCdfRows, envCdfCols, envRowWeights;
    lmc::Scene head;          // scalar members valid; pointer members filled by view()
    std::string outputName;
    int spp, directSpp, numInitSamples;
    int reportIntervalSpp = 0;   // <dpt> reportintervalspp: progressive image dump every so many spp (src/mlt.cpp:171-193)
    std::string integrator;

    // Returns a Scene whose pointers reference this store's host arrays.
    lmc::Scene view() const;
};

// Default options == src/dptoptions.h:7-34 (+ the #define constants the reference compiles in).
lmc::Options default_options();

// Parse a reference-format scene xml from the reference's own scene directory: PNG / JPEG / OpenEXR images are decoded
// here (image_decode.h); any other encoding can be handed over as a "<file>.rawf" sibling (tools/stage_scenes.py:
// float32 or 8-bit RGB + an 8-bit-source flag).  Throws std::runtime_error.
void load_scene_xml(const std::string &xmlPath, SceneStore &out);

// Set / get one option by its reference name (the <dpt> names of src/parsescene.cpp:535-590
// plus the compile-time knobs of the reference: "malastddev", "discretestddev", "maxdervdepth",
// "pssminlength", "pssmaxlength", "lsratio").  Returns false for an unknown name.
bool set_option(lmc::Options &o, const std::string &name, double value);
bool get_option(const lmc::Options &o, const std::string &name, double &value);

// Scene pack I/O (the flattened store in one file).
void save_scene_pack(const std::string &path, const SceneStore &s);
void load_scene_pack(const std::string &path, SceneStore &out);

}  // namespace lmc_host
