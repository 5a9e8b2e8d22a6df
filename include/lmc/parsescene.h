// parsescene.h -- mirror of the reference's loader surface: `std::unique_ptr<Scene> ParseScene(const std::string &)`
// (src/parsescene.h:8, src/parsescene.cpp:627-639).  The returned Scene owns the flattened host scene
// (lmc_scene: meshes, BVH2, materials, textures, lights, camera -- what Scene::Scene builds with Embree in the
// reference, src/scene.cpp:8-46) and exposes the members the render loop reads: `options`, the film size and the
// output name.  Errors throw std::runtime_error with the loader's message, like the reference's Error()
// (src/flexception.h:23).
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include "dptoptions.h"
#include "lmc_abi.h"

namespace lmc {

struct Scene {
    lmc_scene *handle = nullptr;
    std::shared_ptr<DptOptions> options;          // Scene::options, src/scene.h:27
    int pixelWidth = 0, pixelHeight = 0;          // camera->film (src/camera.h, src/image.h)
    std::string outputName = "image.exr";         // Scene::outputName, src/scene.h:28
    lmc_scene_info info;

    Scene() : info() {}
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    ~Scene() { if (handle) lmc_scene_free(handle); }
};

inline std::unique_ptr<Scene> ParseScene(const std::string &filename) {
    std::unique_ptr<Scene> scene(new Scene());
    if (lmc_scene_load(filename.c_str(), &scene->handle) != LMC_OK) throw std::runtime_error(lmc_last_error());
    if (lmc_scene_get_info(scene->handle, &scene->info) != LMC_OK) throw std::runtime_error(lmc_last_error());
    scene->pixelWidth = scene->info.width;
    scene->pixelHeight = scene->info.height;
    scene->options = std::make_shared<DptOptions>();
    scene->options->FromScene(scene->handle);
    return scene;
}

}  // namespace lmc
