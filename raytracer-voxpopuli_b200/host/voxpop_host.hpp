// voxpop_host.hpp - C++ host side above the C ABI (include/voxpop.h).
//
// Mirrors the reference's data model for the render path so existing host code keeps its
// shape: Tmpl8-style float3/mat4/quat value types, `Scene` (one voxel volume + transform,
// reference template/scene.h:173-270), `Material` (src/Materials/Material.h), the light
// PODs (src/Lighting/*.h), `Camera` (template/camera.h) and a `Renderer` whose
// Update()/Trace()/FindNearest()/IsOccluded()/ResetAccumulator() forward to
// libvoxpop_b200.so.  The authoring arithmetic (SetTransform, LoadModel, MaterialSetUp,
// camera frame) is restated here operation by operation so the uploaded bytes equal the
// reference's; tests/test_golden.py::test_host_authoring_matches_reference pins that against the reference build.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/voxpop.h"

namespace voxpop {

struct float3 {
  float x = 0, y = 0, z = 0;
  float3() = default;
  float3(float a, float b, float c) : x(a), y(b), z(c) {}
  explicit float3(float a) : x(a), y(a), z(a) {}
};
inline float3 operator+(const float3& a, const float3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(const float3& a, const float3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(const float3& a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(const float3& a, float b) { return {a.x * b, a.y * b, a.z * b}; }
inline float3 operator*(float b, const float3& a) { return {b * a.x, b * a.y, b * a.z}; }
float dot(const float3& a, const float3& b);
float3 cross(const float3& a, const float3& b);
float3 normalize(const float3& v);  // v * (1/sqrtf(dot)), tmpl8math.h:2350

struct mat4 {  // row-major cells, tmpl8math.h:2593-2860
  float cell[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  static mat4 Translate(const float3& p);
  static mat4 Scale(const float3& s);
  mat4 Inverted() const;
};
mat4 operator*(const mat4& a, const mat4& b);

struct quat {  // tmpl8math.h:2951-3120
  float w = 1, x = 0, y = 0, z = 0;
  void fromAxisAngle(const float3& axis, float theta);
  mat4 toMatrix() const;
  quat operator*(const quat& q) const;
};

float CosDegrees(float degrees);  // tmpl8math.cpp:29-33

namespace MaterialType {  // template/scene.h:36-58
enum MatType : uint8_t {
  NON_METAL_WHITE, NON_METAL_RED, NON_METAL_BLUE, NON_METAL_GREEN, NON_METAL_PINK, METAL_HIGH, METAL_MID, METAL_LOW, GLASS,
  SMOKE_LOW_DENSITY, SMOKE_LOW2_DENSITY, SMOKE_MID_DENSITY, SMOKE_MID2_DENSITY, SMOKE_HIGH_DENSITY, SMOKE_PLAYER, EMISSIVE,
  NONE = 255
};
}

struct Material {  // src/Materials/Material.h
  float3 albedo{1, 1, 1};
  float roughness = 0;
  float emissiveStrength = 0;
  float IOR = 1.5f;
  Material() = default;
  explicit Material(float3 a, float r = 1.0f) : albedo(a), roughness(r) {}
};

// Renderer::MaterialSetUp table (renderer.cpp:357-443), 256 entries
void default_materials(std::vector<std::shared_ptr<Material>>& materials);

struct Ray {  // the fields of template/scene.h Ray that cross the seam
  float3 O, D;
  float3 rayNormal;
  float t = 1e34f;
  bool isInsideGlass = false;
  MaterialType::MatType indexMaterial = MaterialType::NONE;
  Ray() = default;
  Ray(const float3& origin, const float3& direction, float rayLength = 1e34f);  // normalising ctor, scene.cpp:83-93
  float3 IntersectionPoint() const { return O + t * D; }
};

struct Cube { float3 b[2] = {float3(0), float3(1)}; };

class Renderer;

class Scene {  // template/scene.h:173-270
 public:
  Scene(const float3& position, uint32_t worldSize = 64);
  void SetCubeBoundaries(const float3& position);
  void SetTransform(const float3& rotation);
  void SetTransformPlayer(const mat4& rotation);
  void SetScale(const float3& scl) { scale = scl; }
  void ResetGrid(MaterialType::MatType type = MaterialType::NONE);
  void Set(uint32_t x, uint32_t y, uint32_t z, MaterialType::MatType v);
  void CreateEmmisiveSphere(MaterialType::MatType mat, float radiusEmissiveSphere);
  // returns false when the file cannot be read or parsed (the reference silently returns)
  bool LoadModel(Renderer& renderer, const char* filename);
  bool LoadModelPartial(const char* filename, uint32_t columns, uint32_t thickness = 13);
  uint64_t GetVoxel(uint32_t x, uint32_t y, uint32_t z) const { return x + y * gridsize + z * gridsize2; }

  uint32_t worldsize, gridsize, gridsize2, gridsize3;
  std::vector<uint8_t> grid;
  float3 scaleModel{1, 1, 1};
  float3 rotation{0, 0, 0};
  float3 position{0, 0, 0};
  float3 scale{1, 1, 1};
  mat4 invMatrix, matrix;
  Cube cube;
};

struct PointLightData { float3 position{0, 0, 0}; float3 color{1, 1, 1}; };
struct PointLight { PointLightData data{{0.5f, 0.5f, 3.50f}, {1, 1, 1}}; };
struct SpotLightData { float3 position, direction, color; float angle; };
struct SpotLight { SpotLightData data{{-1.0f, .5f, -1.0f}, {1, 0, 0}, {1.5f, 1.5f, 1.5f}, CosDegrees(45.0f)}; };
struct SphereAreaLightData { float3 position{0, 0, 0}; float3 color{0, 0, 0}; float colorMultiplier = 1; float radius = 1.0f; };
struct SphereAreaLight { SphereAreaLightData data{{-0.f, 0.5f, -3.50f}, {1, 1, 1}, 1.2f, 1.2f}; };
struct DirectionalLightData { float3 direction, color; };
struct DirectionalLight { DirectionalLightData data{{1, 0, 0}, {0, 0, 0}}; };

struct Sphere { float3 center{0, 0, 0}; float radius = 1; MaterialType::MatType material = MaterialType::NON_METAL_WHITE; };
struct Triangle { float3 vertex0, vertex1, vertex2, position; MaterialType::MatType material = MaterialType::NONE; };

class Camera {  // template/camera.h
 public:
  Camera(int scrwidth = 256, int scrheight = 212);
  // camPos/camTarget -> ahead/right/up/topLeft/topRight/bottomLeft exactly as the tail of
  // HandleInput does after a move (template/camera.h:121-170)
  void UpdateFrame();
  void LookAt(const float3& pos, const float3& target) { camPos = pos; camTarget = target; UpdateFrame(); }
  VpCamera Pack() const;
  // Camera::SetFrustumNormals (template/camera.h:53-66) and the part of the camera Renderer::CopyToPrevCamera keeps
  void SetFrustumNormals();
  VpPrevCamera PackPrev() const;
  float3 leftNormal, rightNormal, topNormal, bottomNormal;
  int width, height;
  float3 ahead, right, up;
  float3 camPos, camTarget;
  float3 topLeft, topRight, bottomLeft;
  float focalDistance = 1.0f;
  float focalTargetDistance = 1.0f;
  float defocusJitter = 2.0f;
};

// Renderer: the reference class's render-path surface (renderer.h:42-245) over the C ABI.
class Renderer {
 public:
  Renderer(int scrwidth, int scrheight, int device = -1);
  // the same renderer over several GPUs of the box (vp_create_multi): the library shards the rows of every frame
  Renderer(int scrwidth, int scrheight, const std::vector<int>& devices);
  ~Renderer();
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  void MaterialSetUp();         // renderer.cpp:357-443
  void CalculateLightCount();   // renderer.cpp:2290-2296
  void ResetAccumulator();      // renderer.cpp:343-346
  // push voxelVolumes / materials / lights / prims / sky to the device ("dirty" call, between frames)
  void UploadScene();
  // One accumulated frame, renderer.cpp:1646-1891.  The frame is ENQUEUED (vp_render_async): like the reference's Tick,
  // which only needs the pixels when it blits them (renderer.cpp:1972-2217), the host goes on - game logic, the next
  // Update() - while the GPU traces; Pixels() / Accumulator() / Finish() wait for the frames enqueued so far.  Read
  // screenPixels / accumulator directly only after one of them.
  void Update();
  void UpdateBlocking();                          // vp_render: returns with the frame and fills lastStats
  void Finish();
  const uint32_t* Pixels() { Finish(); return screenPixels.data(); }
  const float* Accumulator() { Finish(); return accumulator.data(); }
  // the static-camera branch of Tick (renderer.cpp:1997-2101): one reprojected frame into screenPixels; keeps
  // illuminationHistoryBuffer, and CopyToPrevCamera() (renderer.cpp:1893-1902) is the caller's move, as in the reference
  void UpdateReproject();
  void CopyToPrevCamera() { camera.SetFrustumNormals(); prevCamera = camera; }
  float3 Trace(Ray& ray, int depth, uint32_t seed);  // Renderer::Trace with an explicit RNG state
  int32_t FindNearest(Ray& ray);                  // renderer.cpp:946-1018
  bool IsOccluded(Ray& ray) const;                // renderer.cpp:209-243
  const char* LastError() const;

  int width, height;
  int32_t maxBounces = 14;
  float antiAliasingStrength = 1.0f;
  int32_t numCheckShadowsAreaLight = 3;
  uint64_t numRenderedFrames = 0;
  uint32_t rngSeed = 0x12345678u;
  bool activateSky = true;
  float HDRLightContribution = 1.0f;
  int skyWidth = 0, skyHeight = 0;
  std::vector<float> skyPixels;
  std::vector<float> accumulator;    // W*H float4
  std::vector<uint32_t> screenPixels;  // W*H RGB8 (screen->pixels)
  std::vector<float> illuminationHistoryBuffer, tempIlluminationBuffer;  // W*H float3 (renderer.h:242-243)
  Camera camera, prevCamera;
  std::vector<std::shared_ptr<Material>> materials;
  std::vector<PointLight> pointLights;
  std::vector<SphereAreaLight> areaLights;
  std::vector<SpotLight> spotLights;
  DirectionalLight dirLight;
  size_t lightCount = 0, pointCount = 0, areaCount = 0, spotCount = 0;
  std::vector<Sphere> spheres;
  std::vector<Triangle> triangles;
  std::vector<Scene> voxelVolumes;
  VpStats lastStats{};

 private:
  void Init();
  vp_ctx* ctx_ = nullptr;
  uint64_t ticket_ = 0;  // last frame enqueued by Update()
};

}  // namespace voxpop
