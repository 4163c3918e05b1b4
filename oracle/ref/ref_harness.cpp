// TEST INFRASTRUCTURE ONLY (never linked into the product).
// Headless driver for the REFERENCE's own render path, compiled by build_ref.py into
// oracle/_ref/libvpref.so.  It exposes the same calls as include/voxpop.h with a
// `vpref_` prefix so the parity tests can hand identical scene bytes to the reference,
// to the C restatement (oracle/vp_oracle.c) and to the CUDA library.
//
// Everything arithmetic below is the reference's code (included from the patched
// build directory).  The only restated piece is the scalar pixel loop of
// Renderer::Update (reference renderer.cpp:1851-1881), needed to insert the per-pixel
// reseed and OpenMP (std::execution::par is serial without TBB).
#include <atomic>
#include <omp.h>

struct VpRefCounters { unsigned long long nearest = 0, shadow = 0, exitray = 0, steps = 0; };
static thread_local VpRefCounters g_vp_cnt;
static thread_local unsigned g_vp_lastcell = 0xFFFFFFFFu;
static thread_local int g_vp_lastface = -1;
int g_vp_scrwidth = 256, g_vp_scrheight = 212;

#include "precomp.h"  // the shim (oracle/ref/shim/precomp.h)

static inline int vp_face_code(const float3& n)
{
  auto c = [](float v) { return v > 0 ? 1 : (v < 0 ? 2 : 0); };
  return c(n.x) + 3 * c(n.y) + 9 * c(n.z);
}
static inline __m256 vp_mm256_cos_ps(__m256 v) { float a[8]; _mm256_storeu_ps(a, v); for (float& f : a) f = cosf(f); return _mm256_loadu_ps(a); }
static inline __m256 vp_mm256_sin_ps(__m256 v) { float a[8]; _mm256_storeu_ps(a, v); for (float& f : a) f = sinf(f); return _mm256_loadu_ps(a); }

// Determinism pin for libm: the reference is an MSVC program (its CRT's sinf/cosf/powf/expf
// bits are not glibc's), so there is no canonical float libm.  All three implementations
// use the correctly-rounded value, obtained as float(f64 function).  Hidden visibility binds
// the reference's calls to these definitions without interposing libm for the process.
#define VP_HIDDEN __attribute__((visibility("hidden")))
extern "C" {
VP_HIDDEN float sinf(float x) noexcept { return (float)sin((double)x); }
VP_HIDDEN float cosf(float x) noexcept { return (float)cos((double)x); }
VP_HIDDEN float expf(float x) noexcept { return (float)exp((double)x); }
VP_HIDDEN float powf(float x, float y) noexcept { return (float)pow((double)x, (double)y); }
// g++ merges sinf(x)/cosf(x) pairs (RandomSphereSample, RandomPointInCircle) into one sincosf call
VP_HIDDEN void sincosf(float x, float* s, float* c) noexcept { *s = (float)sin((double)x); *c = (float)cos((double)x); }
}

#include "tmpl8math.cpp"
#include "scene.cpp"
#include "BVH/Shapes.cpp"
#include "BVH/BasicBVH.cpp"
#include "Materials/Material.cpp"
#include "Game/PlayerCharacter.cpp"
#include "Game/ModifyingProp.cpp"
#define STB_IMAGE_IMPLEMENTATION  // reference defines it in template/surface.cpp (not built)
#include "renderer.cpp"

#include "../../include/voxpop.h"

extern thread_local uint32_t seed;  // template/tmpl8math.cpp:16 (made visible by build_ref.py)

struct vpref_ctx
{
  Renderer* r = nullptr;
  std::vector<float> sky;
};

static inline float3 f3(const float* p) { return float3(p[0], p[1], p[2]); }

static inline uint32_t pixel_seed(uint32_t px, uint32_t frame, uint32_t npix, uint32_t s)
{
  uint32_t v = WangHash(px + frame * npix + s);
  return v ? v : 0x9E3779B9u;
}

static void set_camera(Renderer* r, const VpCamera* c)
{
  Camera& cam = r->camera;
  cam.camPos = f3(c->cam_pos);
  cam.topLeft = f3(c->top_left);
  cam.topRight = f3(c->top_right);
  cam.bottomLeft = f3(c->bottom_left);
  cam.right = f3(c->right);
  cam.up = f3(c->up);
  cam.focalDistance = c->focal_distance;
  cam.defocusJitter = c->defocus_jitter;
}

static inline bool row_selected(const VpFrameParams* fp, int y)
{
  int rb = fp->row_begin, re = fp->row_end;
  if (rb == 0 && re == 0) re = fp->height;
  if (y < rb || y >= re) return false;
  if (fp->row_tile > 0 && ((y / fp->row_tile) % fp->row_tile_stride) != fp->row_tile_phase) return false;
  return true;
}

static void fill_hit(VpHit* h, const Ray& ray, int vol)
{
  h->volume = vol;
  h->t = ray.t;
  h->normal[0] = ray.rayNormal.x, h->normal[1] = ray.rayNormal.y, h->normal[2] = ray.rayNormal.z;
  h->material = (uint8_t)ray.indexMaterial;
  h->inside = ray.isInsideGlass ? 1 : 0;
  h->pad = 0;
  if (vol >= 0) { h->cell = g_vp_lastcell; h->face = (int8_t)g_vp_lastface; }
  else { h->cell = 0xFFFFFFFFu; h->face = -1; }
}

extern "C" {

vpref_ctx* vpref_create(int)
{
  vpref_ctx* c = new vpref_ctx;
  c->r = new Renderer();
  c->r->MaterialSetUp();  // renderer.cpp:357-443 (the default table)
  c->r->pointLights.clear();
  c->r->spotLights.clear();
  c->r->areaLights.clear();
  c->r->CalculateLightCount();
  c->r->skyPixels = nullptr;
  c->r->skyWidth = c->r->skyHeight = 0;
  c->r->activateSky = false;
  c->r->accumulator = nullptr;
  return c;
}

void vpref_destroy(vpref_ctx* c)
{
  if (!c) return;
  delete c->r;
  delete c;
}

int vpref_threads() { return omp_get_max_threads(); }
// bench.py: the 1-thread CPU figure SURVEY.md 8(d) asks for, beside the all-cores one
void vpref_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

int vpref_default_materials(vpref_ctx* c, VpMaterial* out, int n)
{
  for (int i = 0; i < n && i < (int)c->r->materials.size(); i++)
  {
    const Material& m = *c->r->materials[i];
    out[i].albedo[0] = m.albedo.x, out[i].albedo[1] = m.albedo.y, out[i].albedo[2] = m.albedo.z;
    out[i].roughness = m.roughness, out[i].emissive = m.emissiveStrength, out[i].ior = m.IOR;
  }
  return 0;
}

int vpref_set_materials(vpref_ctx* c, const VpMaterial* m, int n)
{
  for (int i = 0; i < n && i < (int)c->r->materials.size(); i++)
  {
    Material& d = *c->r->materials[i];
    d.albedo = f3(m[i].albedo);
    d.roughness = m[i].roughness, d.emissiveStrength = m[i].emissive, d.IOR = m[i].ior;
  }
  return 0;
}

int vpref_set_volumes(vpref_ctx* c, const VpVolume* v, int n)
{
  c->r->voxelVolumes.clear();
  for (int i = 0; i < n; i++)
  {
    Scene s(f3(v[i].cube_min), v[i].n);
    memcpy(s.grid.data(), v[i].grid, (size_t)v[i].n * v[i].n * v[i].n);
    s.cube.b[0] = f3(v[i].cube_min), s.cube.b[1] = f3(v[i].cube_max);
    memcpy(s.matrix.cell, v[i].matrix, 64);
    memcpy(s.invMatrix.cell, v[i].inv_matrix, 64);
    c->r->voxelVolumes.push_back(s);
  }
  return 0;
}

// Scene ctor + SetScale + SetTransform (template/scene.cpp:213-217,373-404,430-437):
// pins the host-side transform restatement.
int vpref_make_transform(const float position[3], const float rotation[3], const float scale[3], uint32_t n,
                         float matrix[16], float inv_matrix[16], float cube_min[3], float cube_max[3])
{
  Scene s(f3(position), n);
  s.SetScale(f3(scale));
  s.SetTransform(f3(rotation));
  memcpy(matrix, s.matrix.cell, 64);
  memcpy(inv_matrix, s.invMatrix.cell, 64);
  for (int k = 0; k < 3; k++) cube_min[k] = s.cube.b[0].cell[k], cube_max[k] = s.cube.b[1].cell[k];
  return 0;
}

// Scene::LoadModel (template/scene.cpp:441-524) through the reference's ogt_vox decode:
// pins the .vox parser.  Also returns the palette-derived material edits.
int vpref_load_vox(vpref_ctx* c, const char* path, uint32_t n, uint8_t* grid_out, VpMaterial* mats_out)
{
  Scene s(float3(0), n);
  s.LoadModel(*c->r, path);
  memcpy(grid_out, s.grid.data(), (size_t)n * n * n);
  if (mats_out) vpref_default_materials(c, mats_out, 256);
  return 0;
}

int vpref_set_lights(vpref_ctx* c, const VpPointLight* p, int np, const VpAreaLight* a, int na, const VpSpotLight* s,
                     int ns, const VpDirLight* d, int area_samples)
{
  Renderer* r = c->r;
  r->pointLights.resize(np);
  for (int i = 0; i < np; i++) r->pointLights[i].data = PointLightData{f3(p[i].position), f3(p[i].color)};
  r->areaLights.resize(na);
  for (int i = 0; i < na; i++)
    r->areaLights[i].data = SphereAreaLightData{f3(a[i].position), f3(a[i].color), a[i].color_multiplier, a[i].radius};
  r->spotLights.resize(ns);
  for (int i = 0; i < ns; i++)
    r->spotLights[i].data = SpotLightData{f3(s[i].position), f3(s[i].direction), f3(s[i].color), s[i].angle};
  if (d) r->dirLight.data = DirectionalLightData{f3(d->direction), f3(d->color)};
  r->numCheckShadowsAreaLight = area_samples;
  r->CalculateLightCount();
  return 0;
}

int vpref_set_prims(vpref_ctx* c, const VpSphere* s, int ns, const VpTriangle* t, int nt)
{
  Renderer* r = c->r;
  r->spheres.clear();
  r->triangles.clear();
  for (int i = 0; i < ns; i++) r->spheres.push_back(Sphere(f3(s[i].center), s[i].radius, (MaterialType::MatType)s[i].material));
  for (int i = 0; i < nt; i++)
  {
    Triangle tr((MaterialType::MatType)t[i].material);
    tr.vertex0 = f3(t[i].vertex0), tr.vertex1 = f3(t[i].vertex1), tr.vertex2 = f3(t[i].vertex2);
    tr.position = f3(t[i].position);
    r->triangles.push_back(tr);
  }
  return 0;
}

int vpref_set_sky(vpref_ctx* c, const float* rgb, int w, int h, float contribution, int active)
{
  c->sky.assign(rgb, rgb + (size_t)w * h * 3);
  c->r->skyPixels = c->sky.data();
  c->r->skyWidth = w, c->r->skyHeight = h;
  c->r->HDRLightContribution = contribution;
  c->r->activateSky = active != 0;
  return 0;
}

// Renderer::Update, scalar branch (renderer.cpp:1851-1881), `spp` passes.
int vpref_render(vpref_ctx* c, const VpCamera* cam, const VpFrameParams* fp, float* accum_rgba, uint32_t* rgb8,
                 VpStats* stats)
{
  Renderer* r = c->r;
  const int W = fp->width, H = fp->height;
  g_vp_scrwidth = W, g_vp_scrheight = H;
  set_camera(r, cam);
  r->maxBounces = fp->max_bounces;
  r->antiAliasingStrength = fp->aa_strength;
  r->accumulator = (float4*)accum_rgba;
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
  std::atomic<unsigned long long> nearest{0}, shadow{0}, exitray{0}, steps{0}, primary{0};
  const double t0 = omp_get_wtime();
  for (int s = 0; s < fp->spp; s++)
  {
    const uint32_t frame = fp->frame_index + (uint32_t)s;
    r->numRenderedFrames = frame;
    r->weight = 1.0f / (static_cast<float>(r->numRenderedFrames) + 1.0f);  // renderer.cpp:1651
#pragma omp parallel
    {
      g_vp_cnt = VpRefCounters();
      unsigned long long prim = 0;
#pragma omp for schedule(dynamic, 1)
      for (int y = 0; y < H; y++)
      {
        if (!row_selected(fp, y)) continue;
        const uint32_t pitch = y * W;
        const float invWeight = 1.0f - r->weight;
        for (uint32_t x = 0; x < (uint32_t)W; x++)
        {
          seed = pixel_seed(pitch + x, frame, npix, fp->seed);
          float3 newPixel{0};
          const float randomXDir = RandomFloat() * r->antiAliasingStrength;
          const float randomYDir = RandomFloat() * r->antiAliasingStrength;
          Ray primaryRay = r->camera.GetPrimaryRay(static_cast<float>(x) + randomXDir, static_cast<float>(y) + randomYDir);
          newPixel = r->Trace(primaryRay, r->maxBounces);
          float4 pixel = newPixel;
          pixel = r->BlendColor(pixel, r->accumulator[x + pitch], invWeight);
          r->accumulator[x + pitch] = pixel;
          pixel = r->ApplyReinhardJodie(pixel);
          if (rgb8) rgb8[x + pitch] = RGBF32_to_RGB8(&pixel);
          if (r->maxBounces >= 0) prim++;
        }
      }
      nearest += g_vp_cnt.nearest, shadow += g_vp_cnt.shadow, exitray += g_vp_cnt.exitray, steps += g_vp_cnt.steps;
      primary += prim;
    }
  }
  const double t1 = omp_get_wtime();
  r->numRenderedFrames = fp->frame_index + fp->spp;
  if (stats)
  {
    stats->primary_rays = primary;
    stats->secondary_rays = nearest - primary;
    stats->shadow_rays = shadow;
    stats->exit_rays = exitray;
    stats->dda_steps = steps;
    stats->sky_samples = 0;  // not instrumented in the reference build
    stats->device_ms = (t1 - t0) * 1e3;
    stats->kernel_launches = 0;
  }
  return 0;
}

// Reprojection frame mode: the two pixel loops of Renderer::Tick's static-camera branch (renderer.cpp:2003-2100)
// re-stated under OpenMP with the per-pixel reseed; every call below is the reference's own function, the buffers are
// the reference's own members (std::vector after patch P6).
int vpref_render_reproject(vpref_ctx* c, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp, const float* hist_in,
                           float* hist_out, uint32_t* rgb8, float* albedo_out, float* illum_out, VpStats* stats)
{
  Renderer* r = c->r;
  const int W = fp->width, H = fp->height;
  g_vp_scrwidth = W, g_vp_scrheight = H;
  set_camera(r, cam);
  r->maxBounces = fp->max_bounces;
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
  r->illuminationBuffer.assign(npix, float3(0));
  r->albedoBuffer.assign(npix, float3(0));
  r->tempIlluminationBuffer.assign(npix, float3(0));
  r->illuminationHistoryBuffer.resize(npix);
  r->rayData.resize(npix);
  for (uint32_t i = 0; i < npix; i++) r->illuminationHistoryBuffer[i] = f3(hist_in + 3 * (size_t)i);
  r->prevCamera.camPos = f3(prev->cam_pos);
  r->prevCamera.leftNormal = f3(prev->left_normal);
  r->prevCamera.rightNormal = f3(prev->right_normal);
  r->prevCamera.topNormal = f3(prev->top_normal);
  r->prevCamera.bottomNormal = f3(prev->bottom_normal);
  std::atomic<unsigned long long> nearest{0}, shadow{0}, exitray{0}, steps{0}, primary{0};
  const double t0 = omp_get_wtime();
#pragma omp parallel
  {
    g_vp_cnt = VpRefCounters();
    unsigned long long prim = 0;
#pragma omp for schedule(dynamic, 1)
    for (int y = 0; y < H; y++)
    {
      const uint32_t pitch = y * W;
      for (uint32_t x = 0; x < (uint32_t)W; x++)
      {
        seed = pixel_seed(pitch + x, fp->frame_index, npix, fp->seed);
        Ray primaryRay = r->camera.GetPrimaryRayNoDOF(static_cast<float>(x), static_cast<float>(y));
        const AlbedoIlluminationData data = {r->TraceReproject(primaryRay, r->maxBounces)};
        const uint32_t index = pitch + x;
        r->albedoBuffer[index] = data.albedo;
        r->illuminationBuffer[index] = data.illumination;
        r->rayData[index] = RayDataReproject::GetRayInfo(primaryRay);
        if (r->maxBounces >= 0) prim++;
      }
    }
#pragma omp for schedule(dynamic, 1)
    for (int y = 0; y < H; y++)
    {
      const int32_t pitch = y * W;
      for (int32_t x = 0; x < W; x++)
      {
        const uint32_t index = x + pitch;
        float3 finalIllumination{0};
        const RayDataReproject rayInfo{r->rayData[index]};
        float3 intersectionP{rayInfo.intersectionPoint};
        const MaterialType::MatType matType = rayInfo.materialIndex;
        float3 newSample{r->illuminationBuffer[index]};
        const float2 halfAPixel = {HALF_PIXEL_W, HALF_PIXEL_H};
        const float2 uv{r->prevCamera.PointToUV(intersectionP) + halfAPixel};
        if (r->IsValid(uv) && !r->IsOccludedPrevFrame(intersectionP))
        {
          float3 historySample;
          historySample = r->SampleHistory(uv);
          r->ClampHistory(historySample, newSample, {x, y});
          float w = {0.9f};
          switch (matType)  // renderer.cpp:2051-2086
          {
          case MaterialType::NON_METAL_WHITE: case MaterialType::NON_METAL_RED: case MaterialType::NON_METAL_BLUE:
          case MaterialType::NON_METAL_GREEN: case MaterialType::NON_METAL_PINK: w = 0.8f; break;
          case MaterialType::METAL_HIGH: case MaterialType::METAL_MID: case MaterialType::METAL_LOW: w = 0.5f; break;
          case MaterialType::GLASS: w = 0.5f; break;
          case MaterialType::SMOKE_LOW_DENSITY: case MaterialType::SMOKE_LOW2_DENSITY: case MaterialType::SMOKE_MID_DENSITY:
          case MaterialType::SMOKE_MID2_DENSITY: case MaterialType::SMOKE_HIGH_DENSITY: case MaterialType::SMOKE_PLAYER: w = 0.9f; break;
          case MaterialType::EMISSIVE: w = 0; break;
          default: w = 0.9f; break;
          }
          finalIllumination = lerp(newSample, historySample, w);
        }
        else
        {
          finalIllumination = newSample;
        }
        r->tempIlluminationBuffer[index] = finalIllumination;
        float4 acesPixel = r->ApplyReinhardJodie(r->albedoBuffer[index] * finalIllumination);
        if (rgb8) rgb8[index] = RGBF32_to_RGB8(&acesPixel);
      }
    }
    nearest += g_vp_cnt.nearest, shadow += g_vp_cnt.shadow, exitray += g_vp_cnt.exitray, steps += g_vp_cnt.steps;
    primary += prim;
  }
  const double t1 = omp_get_wtime();
  for (uint32_t i = 0; i < npix; i++)
  {
    const float3 h = r->tempIlluminationBuffer[i], a = r->albedoBuffer[i], l = r->illuminationBuffer[i];
    hist_out[3 * (size_t)i] = h.x, hist_out[3 * (size_t)i + 1] = h.y, hist_out[3 * (size_t)i + 2] = h.z;
    if (albedo_out) albedo_out[3 * (size_t)i] = a.x, albedo_out[3 * (size_t)i + 1] = a.y, albedo_out[3 * (size_t)i + 2] = a.z;
    if (illum_out) illum_out[3 * (size_t)i] = l.x, illum_out[3 * (size_t)i + 1] = l.y, illum_out[3 * (size_t)i + 2] = l.z;
  }
  if (stats)
  {
    stats->primary_rays = primary;
    stats->secondary_rays = nearest - primary;
    stats->shadow_rays = shadow;
    stats->exit_rays = exitray;
    stats->dda_steps = steps;
    stats->sky_samples = 0;
    stats->device_ms = (t1 - t0) * 1e3;
    stats->kernel_launches = 0;
  }
  return 0;
}

int vpref_trace_primary(vpref_ctx* c, const VpCamera* cam, const VpFrameParams* fp, VpHit* hits)
{
  Renderer* r = c->r;
  const int W = fp->width, H = fp->height;
  g_vp_scrwidth = W, g_vp_scrheight = H;
  set_camera(r, cam);
  r->antiAliasingStrength = fp->aa_strength;
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
#pragma omp parallel for schedule(dynamic, 1)
  for (int y = 0; y < H; y++)
  {
    if (!row_selected(fp, y)) continue;
    for (int x = 0; x < W; x++)
    {
      const uint32_t px = (uint32_t)y * W + x;
      seed = pixel_seed(px, fp->frame_index, npix, fp->seed);
      const float randomXDir = RandomFloat() * r->antiAliasingStrength;
      const float randomYDir = RandomFloat() * r->antiAliasingStrength;
      Ray ray = r->camera.GetPrimaryRay(static_cast<float>(x) + randomXDir, static_cast<float>(y) + randomYDir);
      g_vp_lastcell = 0xFFFFFFFFu, g_vp_lastface = -1;
      const int vol = r->FindNearest(ray);
      fill_hit(&hits[px], ray, vol);
    }
  }
  return 0;
}

static inline Ray make_ray(const VpRay& v)
{
  Ray ray;
  ray.O = f3(v.o), ray.D = f3(v.d);
  ray.rD = float3(1 / ray.D.x, 1 / ray.D.y, 1 / ray.D.z);
  ray.Dsign = Ray::ComputeDsign(ray.D);
  ray.t = v.t;
  ray.isInsideGlass = v.inside != 0;
  return ray;
}

int vpref_find_nearest(vpref_ctx* c, const VpRay* rays, int n, VpHit* hits)
{
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++)
  {
    Ray ray;
    ray.O = f3(rays[i].o), ray.D = f3(rays[i].d), ray.t = rays[i].t, ray.isInsideGlass = rays[i].inside != 0;
    ray.rD = float3(1 / ray.D.x, 1 / ray.D.y, 1 / ray.D.z), ray.Dsign = Ray::ComputeDsign(ray.D);
    g_vp_lastcell = 0xFFFFFFFFu, g_vp_lastface = -1;
    const int vol = c->r->FindNearest(ray);
    fill_hit(&hits[i], ray, vol);
  }
  return 0;
}

// gameplay side channels: the reference's own members / methods
int vpref_in_light(vpref_ctx* c, int* out)
{
  *out = c->r->inLight ? 1 : 0;
  c->r->inLight = false;  // renderer.cpp:2206
  return 0;
}

int vpref_find_nearest_player(vpref_ctx* c, const VpRay* rays, int n, int current_chunk, VpHit* hits)
{
  c->r->currentChunk = (size_t)current_chunk;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++)
  {
    Ray ray;
    ray.O = f3(rays[i].o), ray.D = f3(rays[i].d), ray.t = rays[i].t, ray.isInsideGlass = rays[i].inside != 0;
    ray.rD = float3(1 / ray.D.x, 1 / ray.D.y, 1 / ray.D.z), ray.Dsign = Ray::ComputeDsign(ray.D);
    g_vp_lastcell = 0xFFFFFFFFu, g_vp_lastface = -1;
    const int vol = c->r->FindNearestPlayer(ray);
    fill_hit(&hits[i], ray, vol);
  }
  c->r->currentChunk = 0;
  return 0;
}

int vpref_focus_distance(vpref_ctx* c, const VpCamera* cam, int W, int H, uint32_t s, float* out)
{
  Renderer* r = c->r;
  g_vp_scrwidth = W, g_vp_scrheight = H;
  set_camera(r, cam);
  seed = s;
  Ray focusRay = r->camera.GetPrimaryRay(SCRWIDTH / 2, SCRHEIGHT / 2);  // renderer.cpp:1987-1991
  for (auto& scene : r->voxelVolumes)
    scene.FindNearest(focusRay);
  *out = clamp(focusRay.t, -1.0f, 1e4f);
  return 0;
}

int vpref_is_occluded(vpref_ctx* c, const VpRay* rays, int n, uint8_t* out)
{
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++)
  {
    Ray ray;
    ray.O = f3(rays[i].o), ray.D = f3(rays[i].d), ray.t = rays[i].t;
    ray.rD = float3(1 / ray.D.x, 1 / ray.D.y, 1 / ray.D.z), ray.Dsign = Ray::ComputeDsign(ray.D);
    out[i] = c->r->IsOccluded(ray) ? 1 : 0;
  }
  return 0;
}

int vpref_trace(vpref_ctx* c, const VpRay* rays, const uint32_t* seeds, int n, int depth, float* rgb_out)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++)
  {
    Ray ray;
    ray.O = f3(rays[i].o), ray.D = f3(rays[i].d), ray.t = rays[i].t, ray.isInsideGlass = rays[i].inside != 0;
    ray.rD = float3(1 / ray.D.x, 1 / ray.D.y, 1 / ray.D.z), ray.Dsign = Ray::ComputeDsign(ray.D);
    seed = seeds[i];
    const float3 col = c->r->Trace(ray, depth);
    rgb_out[3 * i] = col.x, rgb_out[3 * i + 1] = col.y, rgb_out[3 * i + 2] = col.z;
  }
  return 0;
}

// small pins for the math helpers the restatement re-derives
void vpref_offset_ray(const float p[3], const float n[3], float out[3])
{
  const float3 o = OffsetRay(f3(p), f3(n));
  out[0] = o.x, out[1] = o.y, out[2] = o.z;
}
void vpref_sample_sky(vpref_ctx* c, const float d[3], float out[3])
{
  const float3 s = c->r->SampleSky(f3(d));
  out[0] = s.x, out[1] = s.y, out[2] = s.z;
}
uint32_t vpref_wang_hash(uint32_t s) { return WangHash(s); }
void vpref_rng_floats(uint32_t s, int n, float* out)
{
  seed = s;
  for (int i = 0; i < n; i++) out[i] = RandomFloat();
}
// g++'s argument evaluation order for DiffuseReflection / RandomDirection etc.
void vpref_sampling(uint32_t s, const float nrm[3], float out[15], uint32_t* seed_after)
{
  seed = s;
  float3 a = RandomSphereSample();
  float3 b = RandomDirection();
  float3 d = DiffuseReflection(f3(nrm));
  float2 p = RandomPointInCircle();
  float3 l = RandomLambertianReflectionVector(f3(nrm));
  out[0] = a.x, out[1] = a.y, out[2] = a.z, out[3] = b.x, out[4] = b.y, out[5] = b.z;
  out[6] = d.x, out[7] = d.y, out[8] = d.z, out[9] = p.x, out[10] = p.y, out[11] = 0;
  out[12] = l.x, out[13] = l.y, out[14] = l.z;
  *seed_after = seed;
}
void vpref_tonemap(vpref_ctx* c, const float in[4], float out[3], uint32_t* rgb8)
{
  float4 px(in[0], in[1], in[2], in[3]);
  float4 o = c->r->ApplyReinhardJodie(px);
  out[0] = o.x, out[1] = o.y, out[2] = o.z;
  *rgb8 = RGBF32_to_RGB8(&o);
}
}  // extern "C"
