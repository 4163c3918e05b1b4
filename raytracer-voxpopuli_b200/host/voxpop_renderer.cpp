// voxpop_renderer.cpp - the reference's `Renderer` surface for the render path, forwarding to
// the C ABI (libvoxpop_b200.so).  Same member names and meaning as reference renderer.h:42-245:
// host code fills voxelVolumes / materials / lights / camera exactly as before, calls
// UploadScene() once after edits (the reference edits scene state only between frames and
// always calls ResetAccumulator(), renderer.cpp:2134,2201) and Update() per frame.
#include <cstring>
#include <stdexcept>

#include "voxpop_host.hpp"

namespace voxpop {

static void st3(float* d, const float3& v) { d[0] = v.x, d[1] = v.y, d[2] = v.z; }

Renderer::Renderer(int w, int h, int device) : width(w), height(h), camera(w, h)
{
  ctx_ = vp_create(device);
  if (!ctx_) throw std::runtime_error(std::string("vp_create failed: ") + vp_last_error());
  Init();
}
Renderer::Renderer(int w, int h, const std::vector<int>& devices) : width(w), height(h), camera(w, h)
{
  ctx_ = vp_create_multi(devices.data(), (int)devices.size());
  if (!ctx_) throw std::runtime_error(std::string("vp_create_multi failed: ") + vp_last_error());
  Init();
}
void Renderer::Init()
{
  accumulator.assign((size_t)width * height * 4, 0.0f);  // Init(): MALLOC64 + memset, renderer.cpp:699-702
  screenPixels.assign((size_t)width * height, 0u);
  // page-locked, so that the copies of Update() run asynchronously (failure is not fatal: the copies then block)
  vp_host_register(ctx_, accumulator.data(), accumulator.size() * sizeof(float));
  vp_host_register(ctx_, screenPixels.data(), screenPixels.size() * sizeof(uint32_t));
  MaterialSetUp();
  CalculateLightCount();
}
Renderer::~Renderer()
{
  vp_host_unregister(ctx_, accumulator.data());
  vp_host_unregister(ctx_, screenPixels.data());
  vp_destroy(ctx_);
}
const char* Renderer::LastError() const { return vp_last_error(); }

void Renderer::MaterialSetUp() { default_materials(materials); }
void Renderer::CalculateLightCount()
{
  pointCount = pointLights.size();
  spotCount = spotLights.size();
  areaCount = areaLights.size();
  lightCount = pointCount + spotCount + areaCount + 1;  // directional light is always counted
}
void Renderer::ResetAccumulator() { numRenderedFrames = 0; }
void Renderer::Finish()
{
  if (ticket_ && vp_render_wait(ctx_, ticket_)) throw std::runtime_error(vp_last_error());
  ticket_ = 0;
}

void Renderer::UploadScene()
{
  std::vector<VpVolume> vols(voxelVolumes.size());
  for (size_t i = 0; i < vols.size(); i++) {
    const Scene& s = voxelVolumes[i];
    vols[i].grid = s.grid.data();
    vols[i].n = s.gridsize;
    st3(vols[i].cube_min, s.cube.b[0]);
    st3(vols[i].cube_max, s.cube.b[1]);
    memcpy(vols[i].matrix, s.matrix.cell, 64);
    memcpy(vols[i].inv_matrix, s.invMatrix.cell, 64);
  }
  auto chk = [](int rc) { if (rc) throw std::runtime_error(vp_last_error()); };
  chk(vp_set_volumes(ctx_, vols.data(), (int)vols.size()));
  std::vector<VpMaterial> mats(materials.size());
  for (size_t i = 0; i < mats.size(); i++) {
    st3(mats[i].albedo, materials[i]->albedo);
    mats[i].roughness = materials[i]->roughness, mats[i].emissive = materials[i]->emissiveStrength, mats[i].ior = materials[i]->IOR;
  }
  chk(vp_set_materials(ctx_, mats.data(), (int)mats.size()));
  std::vector<VpPointLight> pl(pointLights.size());
  for (size_t i = 0; i < pl.size(); i++) st3(pl[i].position, pointLights[i].data.position), st3(pl[i].color, pointLights[i].data.color);
  std::vector<VpAreaLight> al(areaLights.size());
  for (size_t i = 0; i < al.size(); i++) {
    st3(al[i].position, areaLights[i].data.position), st3(al[i].color, areaLights[i].data.color);
    al[i].color_multiplier = areaLights[i].data.colorMultiplier, al[i].radius = areaLights[i].data.radius;
  }
  std::vector<VpSpotLight> sl(spotLights.size());
  for (size_t i = 0; i < sl.size(); i++) {
    st3(sl[i].position, spotLights[i].data.position), st3(sl[i].direction, spotLights[i].data.direction);
    st3(sl[i].color, spotLights[i].data.color), sl[i].angle = spotLights[i].data.angle;
  }
  VpDirLight dl;
  st3(dl.direction, dirLight.data.direction), st3(dl.color, dirLight.data.color);
  CalculateLightCount();
  chk(vp_set_lights(ctx_, pl.data(), (int)pl.size(), al.data(), (int)al.size(), sl.data(), (int)sl.size(), &dl, numCheckShadowsAreaLight));
  std::vector<VpSphere> sp(spheres.size());
  for (size_t i = 0; i < sp.size(); i++) st3(sp[i].center, spheres[i].center), sp[i].radius = spheres[i].radius, sp[i].material = spheres[i].material;
  std::vector<VpTriangle> tr(triangles.size());
  for (size_t i = 0; i < tr.size(); i++) {
    st3(tr[i].vertex0, triangles[i].vertex0), st3(tr[i].vertex1, triangles[i].vertex1), st3(tr[i].vertex2, triangles[i].vertex2);
    st3(tr[i].position, triangles[i].position), tr[i].material = triangles[i].material;
  }
  chk(vp_set_prims(ctx_, sp.data(), (int)sp.size(), tr.data(), (int)tr.size()));
  chk(vp_set_sky(ctx_, skyPixels.empty() ? nullptr : skyPixels.data(), skyWidth, skyHeight, HDRLightContribution,
                 activateSky && !skyPixels.empty()));
}

static VpFrameParams frame_params(const Renderer& r, int spp)
{
  VpFrameParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.width = r.width, fp.height = r.height, fp.max_bounces = r.maxBounces, fp.spp = spp;
  fp.frame_index = (uint32_t)r.numRenderedFrames, fp.seed = r.rngSeed, fp.aa_strength = r.antiAliasingStrength;
  fp.row_tile_stride = 1;
  return fp;
}

void Renderer::Update()  // renderer.cpp:1646-1891
{
  const VpCamera cam = camera.Pack();
  const VpFrameParams fp = frame_params(*this, 1);
  // the first frame after ResetAccumulator() ignores the accumulator (weight 1): nothing to upload; later frames
  // continue from the library's device copy of the previous frame's result
  const float* in = numRenderedFrames == 0 ? accumulator.data() : nullptr;
  if (vp_render_async(ctx_, &cam, &fp, in, accumulator.data(), screenPixels.data(), &ticket_)) throw std::runtime_error(vp_last_error());
  numRenderedFrames++;
}

void Renderer::UpdateBlocking()
{
  Finish();
  const VpCamera cam = camera.Pack();
  const VpFrameParams fp = frame_params(*this, 1);
  if (vp_render(ctx_, &cam, &fp, accumulator.data(), screenPixels.data(), &lastStats)) throw std::runtime_error(vp_last_error());
  numRenderedFrames++;
}

void Renderer::UpdateReproject()  // renderer.cpp:1997-2101
{
  Finish();
  const size_t n = (size_t)width * height * 3;
  if (illuminationHistoryBuffer.size() != n) illuminationHistoryBuffer.assign(n, 0.0f);
  tempIlluminationBuffer.resize(n);
  const VpCamera cam = camera.Pack();
  const VpPrevCamera prev = prevCamera.PackPrev();
  const VpFrameParams fp = frame_params(*this, 1);
  if (vp_render_reproject(ctx_, &cam, &prev, &fp, illuminationHistoryBuffer.data(), tempIlluminationBuffer.data(), screenPixels.data(), nullptr,
                          nullptr, &lastStats))
    throw std::runtime_error(vp_last_error());
  illuminationHistoryBuffer.swap(tempIlluminationBuffer);  // renderer.cpp:2100
  numRenderedFrames++;
}

static VpRay pack_ray(const Ray& ray)
{
  VpRay r;
  st3(r.o, ray.O), st3(r.d, ray.D);
  r.t = ray.t, r.inside = ray.isInsideGlass ? 1u : 0u;
  return r;
}

int32_t Renderer::FindNearest(Ray& ray)  // renderer.cpp:946-1018
{
  const VpRay r = pack_ray(ray);
  VpHit h;
  if (vp_find_nearest(ctx_, &r, 1, &h)) throw std::runtime_error(vp_last_error());
  ray.t = h.t;
  if (h.volume != -2) {
    ray.rayNormal = float3(h.normal[0], h.normal[1], h.normal[2]);
    ray.indexMaterial = (MaterialType::MatType)h.material;
    ray.isInsideGlass = h.inside != 0;
  }
  return h.volume;
}

bool Renderer::IsOccluded(Ray& ray) const  // renderer.cpp:209-243
{
  const VpRay r = pack_ray(ray);
  uint8_t out = 0;
  if (vp_is_occluded(ctx_, &r, 1, &out)) throw std::runtime_error(vp_last_error());
  return out != 0;
}

float3 Renderer::Trace(Ray& ray, int depth, uint32_t seed)  // renderer.cpp:1076-1328
{
  const VpRay r = pack_ray(ray);
  float rgb[3];
  if (vp_trace(ctx_, &r, &seed, 1, depth, rgb)) throw std::runtime_error(vp_last_error());
  return float3(rgb[0], rgb[1], rgb[2]);
}

}  // namespace voxpop

// ---- C test hook: drives the C++ mirror exactly like reference host code would ---------------
using namespace voxpop;
extern "C" int vphr_demo_render(const char* vox_path, const uint8_t* grid, uint32_t n, int width, int height, int frames, int max_bounces,
                                float* accum_out, uint32_t* rgb8_out, unsigned long long* rays_out)
{
  try {
    Renderer r(width, height);
    Scene s(float3(0, 0, 0), n);
    if (vox_path && vox_path[0]) { if (!s.LoadModel(r, vox_path)) return -2; }
    else memcpy(s.grid.data(), grid, s.grid.size());
    s.SetTransform(float3(0, 0.3f, 0));
    r.voxelVolumes.push_back(s);
    r.pointLights.resize(1);
    r.pointLights[0].data.position = float3(0.5f, 1.6f, -0.6f);
    r.pointLights[0].data.color = float3(4, 4, 4);
    r.maxBounces = max_bounces;
    r.camera.defocusJitter = 0.0f;
    r.camera.LookAt(float3(0.5f, 0.55f, -0.75f), float3(0.5f, 0.3f, 0.5f));
    r.UploadScene();
    unsigned long long rays = 0;
    for (int f = 0; f < frames; f++) {
      r.UpdateBlocking();
      rays += r.lastStats.primary_rays + r.lastStats.secondary_rays + r.lastStats.shadow_rays;
    }
    // the same frames again through the asynchronous Update(): enqueue all of them, then read
    std::vector<float> first(r.accumulator);
    std::vector<uint32_t> first_rgb(r.screenPixels);
    r.ResetAccumulator();
    for (int f = 0; f < frames; f++) r.Update();
    if (memcmp(first.data(), r.Accumulator(), first.size() * sizeof(float)) != 0) return -3;
    if (memcmp(first_rgb.data(), r.Pixels(), first_rgb.size() * sizeof(uint32_t)) != 0) return -4;
    memcpy(accum_out, r.accumulator.data(), r.accumulator.size() * sizeof(float));
    memcpy(rgb8_out, r.screenPixels.data(), r.screenPixels.size() * sizeof(uint32_t));
    *rays_out = rays;
    return 0;
  } catch (const std::exception& e) {
    return -1;
  }
}

// Test hook: the same scene through the mirror's reprojection mode with a camera that moves between frames
// (tests/test_gpu_parity.py::test_cpp_renderer_mirror_reproject replays it through the oracle).
extern "C" int vphr_demo_reproject(const uint8_t* grid, uint32_t n, int width, int height, int frames, int max_bounces, float* history_out,
                                   uint32_t* rgb8_out)
{
  try {
    Renderer r(width, height);
    Scene s(float3(0, 0, 0), n);
    memcpy(s.grid.data(), grid, s.grid.size());
    s.SetTransform(float3(0, 0.3f, 0));
    r.voxelVolumes.push_back(s);
    r.pointLights.resize(1);
    r.pointLights[0].data.position = float3(0.5f, 1.6f, -0.6f);
    r.pointLights[0].data.color = float3(4, 4, 4);
    r.maxBounces = max_bounces;
    r.UploadScene();
    for (int f = 0; f < frames; f++) {
      r.camera.LookAt(float3(0.5f + 0.01f * (float)f, 0.55f, -0.75f), float3(0.5f, 0.3f, 0.5f));
      if (f == 0) r.CopyToPrevCamera();
      r.UpdateReproject();
      r.CopyToPrevCamera();
    }
    memcpy(history_out, r.illuminationHistoryBuffer.data(), r.illuminationHistoryBuffer.size() * sizeof(float));
    memcpy(rgb8_out, r.screenPixels.data(), r.screenPixels.size() * sizeof(uint32_t));
    return 0;
  } catch (const std::exception& e) {
    return -1;
  }
}
