// bsdf.h -- mirror of the reference's scene-model plugin enums and serialized sizes (src/bsdf.h:6-68,
// src/light.h:7, src/shape.h:10) and the description of how a BSDF plugs into the device code.
//
// The reference's `struct BSDF` is a C++ class with virtual Evaluate / EvaluateAdjoint / Sample / SampleAdjoint /
// Roughness / Serialize (src/bsdf.h:10-68), selected per shape by ParseBSDF (src/parsescene.cpp:341-412).  A CUDA
// kernel cannot make a virtual call per mutation, so the same surface is an ENUM-KEYED TABLE on the device:
//
//   BSDFType value                     csrc/core/scene.h   enum BsdfType (same numbering as below)
//   row of LMC_BSDF_TABLE              csrc/core/bsdf.h    ROW(BSDF_<ID>, <name>) -> <name>_evaluate / <name>_sample /
//                                                          <name>_roughness, all with one fixed signature; bsdf_eval,
//                                                          bsdf_sample and bsdf_roughness are generated from the table
//   parameters                         csrc/core/bsdf.h    BsdfParams + serialize_bsdf: the 10-float record of the
//                                                          reference's BSDF::Serialize
//   differentiable twin                tools/adgen/pathfn.py   one row of BSDF_TABLE (evaluate, sample); the reverse
//                                                          sweep (adjoint) is GENERATED from it by tools/adgen/gen.py,
//                                                          nobody writes derivative code by hand
//   loader                             csrc/host/host_scene.cpp   parse_bsdf maps the xml `type` to the enum
//
// Adding a BSDF is therefore "enum value + (evaluate, sample, roughness) triple + twin row", the device analogue of
// deriving from `struct BSDF`.
#pragma once

namespace lmc {

enum class BSDFType { Lambertian, Phong, RoughDielectric };      // src/bsdf.h:6
enum class LightType { PointLight, AreaLight, EnvLight };         // src/light.h:7
enum class ShapeType { TriangleMesh };                            // src/shape.h:10

// serialized record sizes of the generated path functions' buffers
inline int GetLambertianSerializedSize() { return 4; }            // src/lambertian.cpp:5
inline int GetPhongSerializedSize() { return 9; }                 // src/phong.cpp:6
inline int GetRoughDielectricSerializedSize() { return 10; }      // src/roughdielectric.cpp:4
inline int GetMaxBSDFSerializedSize() { return 10; }              // src/bsdf.cpp:7-11
inline int GetMaxLightSerializedSize() { return 56; }             // src/light.cpp:7-10 (env light)
inline int GetMaxShapeSerializedSize() { return 46; }             // src/trianglemesh.cpp:3-10
inline int GetSceneSerializedSize() { return 38; }                // src/scene.cpp:160-162

}  // namespace lmc
