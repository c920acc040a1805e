// scene.h — C++ mirror of the reference's internal/scene package (scene.go:9-158, io.go:10-38).
//
// The reference host language is Go; no Go toolchain exists in this image, so the host side above the
// C ABI is written in C++ with the same type names, field meanings, JSON keys and Load/Save behaviour.
// (go/ holds the cgo binding a maintainer adds on the reference side; INTEGRATION.md explains both.)
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/ptb200.h"

namespace scene {

struct Vec3 { double X = 0, Y = 0, Z = 0; };                 // scene.go:9-13   json x,y,z
struct Color { double R = 0, G = 0, B = 0; };                // scene.go:16-20  json r,g,b

struct Camera {                                              // scene.go:23-32
    Vec3 Position, Target, Up;
    double FOV = 0, Aperture = 0, FocusDist = 0, AspectRatio = 0;
};

// MaterialType strings, scene.go:37-43
constexpr const char* MaterialLambert = "lambert";
constexpr const char* MaterialMetal = "metal";
constexpr const char* MaterialDielectric = "dielectric";
constexpr const char* MaterialEmissive = "emissive";
constexpr const char* MaterialMirror = "mirror";

struct Material {                                            // scene.go:46-68
    std::string ID, Type;
    Color Albedo;
    double Rough = 0, IOR = 0;
    Color Emit;
    double Power = 0;
    Color Absorption;
    double Smoothness = 0, Reflectivity = 0;
    Color Tint;
    double AbsorptionScale = 0;
};

// ObjectType strings, scene.go:73-78
constexpr const char* ObjectSphere = "sphere";
constexpr const char* ObjectPlane = "plane";
constexpr const char* ObjectBox = "box";
constexpr const char* ObjectSphereLight = "sphere_light";
constexpr const char* ObjectMesh = "mesh";   // EXTENSION (not in the reference): triangle mesh, see MeshData

// EXTENSION: triangle mesh attached to an Object of type "mesh" (north-star: "internal/scene gains a BVH builder").
// JSON:  "mesh": {"vertices": [x0,y0,z0, x1,...], "triangles": [a0,b0,c0, ...]}            inline, or
//        "mesh": {"heightfield": {"nx":..,"nz":..,"seed":..,"amplitude":..,"frequency":..,"octaves":..}}   generator:
//        nx x nz quads (2 triangles each) over [-0.5,0.5]^2 in x,z with y = amplitude * fractal value noise.
// World vertex = position + size * local vertex (a zero size component means 1).  The reference's Go decoder
// ignores the unknown "mesh" key and sceneToWorld drops the unknown object type (objects.go:237-266), so scene
// files with meshes still load there (without the mesh).
struct MeshData {
    std::vector<float> vertices;       // 3 per vertex, local space
    std::vector<uint32_t> triangles;   // 3 vertex indices per triangle
    // Identity of the vertex / index data for ptb_scene.mesh_generation: unique per MeshData object at creation.  Code that
    // edits vertices or triangles in place must call Touch() — the flattener folds it (with the owning object's position and
    // size) into the generation the CUDA library uses to recognise an unchanged mesh without reading its triangles.
    uint64_t generation = NextGeneration();
    void Touch() { generation = NextGeneration(); }
    static uint64_t NextGeneration();
    bool generated = false;            // true: came from the heightfield generator (Save writes the parameters back)
    int nx = 0, nz = 0, octaves = 0;
    uint32_t seed = 0;
    double amplitude = 0, frequency = 0;
};
void GenerateHeightfield(MeshData& m);

struct Object {                                              // scene.go:81-89
    std::string ID, Type;
    Vec3 Position, Size;
    std::string MaterialID;
    std::shared_ptr<MeshData> Mesh;                          // EXTENSION, only for Type == "mesh"
};

struct RenderSettings { int Width = 0, Height = 0, SamplesPerPx = 0, MaxDepth = 0; };   // scene.go:92-97

struct Fog {                                                 // scene.go:101-135 (GL back-end only; carried for Save)
    double Density = 0;
    Color FogColor;
    double Scatter = 0, SigmaS = 0, SigmaA = 0, G = 0, HeteroStrength = 0, NoiseScale = 0;
    int NoiseOctaves = 0;
    bool AffectSky = false, GPUVolumetric = false;
};

struct Sky {                                                 // scene.go:138-143
    std::string Type;
    Color SkyColor, Horizon, Zenith;
};

struct Scene {                                               // scene.go:146-158
    std::string Name;
    Camera Cam;
    std::vector<Object> Objects;
    std::vector<Material> Materials;
    RenderSettings Settings;
    Color Background;
    std::unique_ptr<Sky> SkyPtr;   // nil when absent/null
    std::unique_ptr<Fog> FogPtr;   // omitempty
};

// scene.Load (io.go:10-22).  Throws std::runtime_error("open scene: ...") / ("decode scene: ...").
std::unique_ptr<Scene> Load(const std::string& path);
std::unique_ptr<Scene> Parse(const std::string& json_text);
// scene.Save (io.go:25-38): 2-space indented JSON, Go field order.
void Save(const std::string& path, const Scene& sc);
std::string Marshal(const Scene& sc);

// SoA flattening for ptb_scene_upload.  The arrays live inside Flat; view() points into them.
struct Flat {
    std::vector<int32_t> obj_type, obj_mat, mat_type;
    std::vector<double> obj_pos, obj_size, mat_albedo, mat_rough, mat_ior, mat_emit, mat_power, mat_absorption, mat_smoothness;
    ptb_camera camera{};
    ptb_sky sky{};
    // EXTENSION: meshes of the "mesh" objects, world space, binary32
    std::vector<int32_t> obj_mesh;          // [n_obj] mesh index or -1
    std::vector<int64_t> mesh_tri_begin;    // [n_mesh+1] prefix of triangle counts
    std::vector<float> tri_vertices;        // 9 floats per triangle (v0, v1, v2), world space
    uint64_t mesh_generation = 0;           // ptb_scene.mesh_generation: hash of every mesh object's (MeshData::generation, position, size)
    ptb_scene view() const;
};
Flat Flatten(const Scene& sc);

}  // namespace scene
