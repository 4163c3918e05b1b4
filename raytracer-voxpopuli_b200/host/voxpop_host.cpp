// voxpop_host.cpp - scene authoring side of the host mirror (no CUDA, no vp_* calls).
// Each function restates the cited reference code in its float operation order.
#include "voxpop_host.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace voxpop {

// correctly-rounded libm pins shared by all implementations (DESIGN.md "determinism pins")
static inline float p_sinf(float x) { return (float)std::sin((double)x); }
static inline float p_cosf(float x) { return (float)std::cos((double)x); }

float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
float3 cross(const float3& a, const float3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float3 normalize(const float3& v)
{
  const float invLen = 1.0f / sqrtf(dot(v, v));
  return v * invLen;
}

float CosDegrees(float degrees)  // tmpl8math.cpp:29-33
{
  const float radians = degrees * 3.14159265358979323846264f / 180.0f;
  return p_cosf(radians);
}

mat4 mat4::Translate(const float3& p)
{
  mat4 r;
  r.cell[3] = p.x, r.cell[7] = p.y, r.cell[11] = p.z;
  return r;
}
mat4 mat4::Scale(const float3& s)
{
  mat4 r;
  r.cell[0] = s.x, r.cell[5] = s.y, r.cell[10] = s.z;
  return r;
}
mat4 operator*(const mat4& a, const mat4& b)  // tmpl8math.cpp:262-275
{
  mat4 r;
  for (unsigned i = 0; i < 16; i += 4)
    for (unsigned j = 0; j < 4; ++j)
      r.cell[i + j] = (a.cell[i + 0] * b.cell[j + 0]) + (a.cell[i + 1] * b.cell[j + 4]) + (a.cell[i + 2] * b.cell[j + 8]) +
                      (a.cell[i + 3] * b.cell[j + 12]);
  return r;
}
mat4 mat4::Inverted() const  // tmpl8math.h:2810-2857 (MESA cofactor expansion, same term order)
{
  const float* c = cell;
  const float inv[16] = {
      c[5] * c[10] * c[15] - c[5] * c[11] * c[14] - c[9] * c[6] * c[15] + c[9] * c[7] * c[14] + c[13] * c[6] * c[11] - c[13] * c[7] * c[10],
      -c[1] * c[10] * c[15] + c[1] * c[11] * c[14] + c[9] * c[2] * c[15] - c[9] * c[3] * c[14] - c[13] * c[2] * c[11] + c[13] * c[3] * c[10],
      c[1] * c[6] * c[15] - c[1] * c[7] * c[14] - c[5] * c[2] * c[15] + c[5] * c[3] * c[14] + c[13] * c[2] * c[7] - c[13] * c[3] * c[6],
      -c[1] * c[6] * c[11] + c[1] * c[7] * c[10] + c[5] * c[2] * c[11] - c[5] * c[3] * c[10] - c[9] * c[2] * c[7] + c[9] * c[3] * c[6],
      -c[4] * c[10] * c[15] + c[4] * c[11] * c[14] + c[8] * c[6] * c[15] - c[8] * c[7] * c[14] - c[12] * c[6] * c[11] + c[12] * c[7] * c[10],
      c[0] * c[10] * c[15] - c[0] * c[11] * c[14] - c[8] * c[2] * c[15] + c[8] * c[3] * c[14] + c[12] * c[2] * c[11] - c[12] * c[3] * c[10],
      -c[0] * c[6] * c[15] + c[0] * c[7] * c[14] + c[4] * c[2] * c[15] - c[4] * c[3] * c[14] - c[12] * c[2] * c[7] + c[12] * c[3] * c[6],
      c[0] * c[6] * c[11] - c[0] * c[7] * c[10] - c[4] * c[2] * c[11] + c[4] * c[3] * c[10] + c[8] * c[2] * c[7] - c[8] * c[3] * c[6],
      c[4] * c[9] * c[15] - c[4] * c[11] * c[13] - c[8] * c[5] * c[15] + c[8] * c[7] * c[13] + c[12] * c[5] * c[11] - c[12] * c[7] * c[9],
      -c[0] * c[9] * c[15] + c[0] * c[11] * c[13] + c[8] * c[1] * c[15] - c[8] * c[3] * c[13] - c[12] * c[1] * c[11] + c[12] * c[3] * c[9],
      c[0] * c[5] * c[15] - c[0] * c[7] * c[13] - c[4] * c[1] * c[15] + c[4] * c[3] * c[13] + c[12] * c[1] * c[7] - c[12] * c[3] * c[5],
      -c[0] * c[5] * c[11] + c[0] * c[7] * c[9] + c[4] * c[1] * c[11] - c[4] * c[3] * c[9] - c[8] * c[1] * c[7] + c[8] * c[3] * c[5],
      -c[4] * c[9] * c[14] + c[4] * c[10] * c[13] + c[8] * c[5] * c[14] - c[8] * c[6] * c[13] - c[12] * c[5] * c[10] + c[12] * c[6] * c[9],
      c[0] * c[9] * c[14] - c[0] * c[10] * c[13] - c[8] * c[1] * c[14] + c[8] * c[2] * c[13] + c[12] * c[1] * c[10] - c[12] * c[2] * c[9],
      -c[0] * c[5] * c[14] + c[0] * c[6] * c[13] + c[4] * c[1] * c[14] - c[4] * c[2] * c[13] - c[12] * c[1] * c[6] + c[12] * c[2] * c[5],
      c[0] * c[5] * c[10] - c[0] * c[6] * c[9] - c[4] * c[1] * c[10] + c[4] * c[2] * c[9] + c[8] * c[1] * c[6] - c[8] * c[2] * c[5]};
  const float det = c[0] * inv[0] + c[1] * inv[4] + c[2] * inv[8] + c[3] * inv[12];
  mat4 ret;
  if (det != 0) {
    const float invdet = 1.0f / det;
    for (int i = 0; i < 16; i++) ret.cell[i] = inv[i] * invdet;
  }
  return ret;
}

void quat::fromAxisAngle(const float3& axis, float theta)  // tmpl8math.h:2980-2985
{
  w = p_cosf(theta / 2);
  const float s = p_sinf(theta / 2);
  x = axis.x * s, y = axis.y * s, z = axis.z * s;
}
mat4 quat::toMatrix() const  // tmpl8math.h:3025-3035
{
  mat4 ret;
  ret.cell[0] = 1 - 2 * y * y - 2 * z * z;
  ret.cell[1] = 2 * x * y - 2 * w * z, ret.cell[2] = 2 * x * z + 2 * w * y, ret.cell[4] = 2 * x * y + 2 * w * z;
  ret.cell[5] = 1 - 2 * x * x - 2 * z * z;
  ret.cell[6] = 2 * y * z - 2 * w * x, ret.cell[8] = 2 * x * z - 2 * w * y, ret.cell[9] = 2 * y * z + 2 * w * x;
  ret.cell[10] = 1 - 2 * x * x - 2 * y * y;
  return ret;
}
quat quat::operator*(const quat& q) const  // tmpl8math.h:3063-3069
{
  quat r;
  r.w = w * q.w - x * q.x - y * q.y - z * q.z;
  r.x = w * q.x + x * q.w + y * q.z - z * q.y;
  r.y = w * q.y - x * q.z + y * q.w + z * q.x;
  r.z = w * q.z + x * q.y - y * q.x + z * q.w;
  return r;
}

Ray::Ray(const float3& origin, const float3& direction, float rayLength) : O(origin), t(rayLength) { D = normalize(direction); }

// ---- Scene ------------------------------------------------------------------------------
Scene::Scene(const float3& pos, uint32_t worldSize)  // template/scene.cpp:427-437
    : worldsize(worldSize), gridsize(worldSize), gridsize2(worldSize * worldSize), gridsize3(worldSize * worldSize * worldSize)
{
  grid.resize(gridsize3);
  SetCubeBoundaries(pos);
  ResetGrid(MaterialType::NON_METAL_PINK);
}
void Scene::SetCubeBoundaries(const float3& p)  // scene.cpp:213-217
{
  cube.b[0] = p;
  cube.b[1] = p + float3(1, 1, 1);
}
void Scene::ResetGrid(MaterialType::MatType type) { std::fill(grid.begin(), grid.end(), (uint8_t)type); }
void Scene::Set(uint32_t x, uint32_t y, uint32_t z, MaterialType::MatType v) { grid[GetVoxel(x, y, z)] = v; }

static void set_transform(Scene& s, const mat4& rot)  // shared tail of scene.cpp:373-425
{
  const float3 centerCube = (s.cube.b[0] + s.cube.b[1]) * 0.5f;
  const mat4 translateToPivot = mat4::Translate(centerCube + s.position);
  const mat4 translateBack = mat4::Translate(-centerCube);
  const mat4 scl = mat4::Scale(s.scale);
  s.matrix = translateToPivot * scl * rot * translateBack;
  s.invMatrix = (translateToPivot * rot * scl * translateBack).Inverted();
}
void Scene::SetTransform(const float3& r)  // scene.cpp:373-404
{
  quat qRot;
  qRot.fromAxisAngle(float3{1, 0, 0}, r.x);
  quat qRotY;
  qRotY.fromAxisAngle(float3{0, 1, 0}, r.y);
  qRot = qRotY * qRot;
  quat qRotZ;
  qRotZ.fromAxisAngle(float3{0, 0, 1}, r.z);
  qRot = qRotZ * qRot;
  set_transform(*this, qRot.toMatrix());
}
void Scene::SetTransformPlayer(const mat4& r) { set_transform(*this, r); }  // scene.cpp:406-425

void Scene::CreateEmmisiveSphere(MaterialType::MatType mat, float radius)  // scene.cpp:685-709
{
  const float worldCenter{(static_cast<float>(worldsize) / 2.0f)};
  for (uint32_t z = 0; z < worldsize; ++z)
    for (uint32_t y = 0; y < worldsize; ++y)
      for (uint32_t x = 0; x < worldsize; ++x) {
        const float3 point{static_cast<float>(x), static_cast<float>(y), static_cast<float>(z)};
        const float3 d = float3(worldCenter) - point;
        if (sqrtf(dot(d, d)) < radius) Set(x, y, z, mat);
      }
}

// ---- MagicaVoxel .vox reader ---------------------------------------------------------------
// Restates what the reference obtains from its vendored ogt_vox (lib/ogt_vox.h, read with
// flags 0) for models[0] and the palette: dense x + y*sx + z*sx*sy colour indices, IMAP
// display-order remap (ogt_vox.h:2004-2038) and the palette rotation (:2041-2047) after
// which a voxel's colour index addresses palette.color[] directly.
namespace {
struct VoxModel {
  uint32_t sx = 0, sy = 0, sz = 0;
  std::vector<uint8_t> voxels;  // 0 = empty
  uint8_t rgba[256][4];
  bool ok = false;
};
uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// MagicaVoxel's built-in palette is only used by files without an RGBA chunk; none of the
// reference's assets is such a file, so a missing RGBA chunk is reported as a parse failure.
VoxModel read_vox(const char* filename)
{
  VoxModel m;
  FILE* fp = fopen(filename, "rb");
  if (!fp) return m;
  fseek(fp, 0, SEEK_END);
  const long size = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)(size > 0 ? size : 0));
  const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), fp);
  fclose(fp);
  if (got != buf.size() || buf.size() < 20 || memcmp(buf.data(), "VOX ", 4) != 0) return m;
  uint8_t palette[256][4];
  uint8_t index_map[256];
  bool have_palette = false, have_imap = false, have_model = false;
  uint32_t psx = 0, psy = 0, psz = 0;
  size_t pos = 8;
  // MAIN has no content; its children follow as a flat chunk list
  while (pos + 12 <= buf.size()) {
    const uint8_t* h = buf.data() + pos;
    const uint32_t n = rd32(h + 4), kids = rd32(h + 8);
    const uint8_t* body = h + 12;
    if (pos + 12 + (size_t)n > buf.size()) break;
    if (!memcmp(h, "MAIN", 4)) { pos += 12 + n; (void)kids; continue; }
    if (!memcmp(h, "SIZE", 4) && n >= 12) { psx = rd32(body); psy = rd32(body + 4); psz = rd32(body + 8); }
    else if (!memcmp(h, "XYZI", 4) && n >= 4 && !have_model) {
      const uint32_t nv = rd32(body);
      if (4 + (size_t)nv * 4 > n || psx == 0 || psy == 0 || psz == 0) return m;
      if (nv != 0) {  // ogt_vox skips empty models when flags == 0 (ogt_vox.h:1444)
        m.sx = psx; m.sy = psy; m.sz = psz;
        m.voxels.assign((size_t)psx * psy * psz, 0);
        for (uint32_t i = 0; i < nv; i++) {
          const uint8_t* v = body + 4 + (size_t)i * 4;
          if (v[0] < psx && v[1] < psy && v[2] < psz) m.voxels[v[0] + (size_t)v[1] * psx + (size_t)v[2] * psx * psy] = v[3];
        }
        have_model = true;
      }
    } else if (!memcmp(h, "RGBA", 4) && n == 1024) { memcpy(palette, body, 1024); have_palette = true; }
    else if (!memcmp(h, "IMAP", 4) && n == 256) { memcpy(index_map, body, 256); have_imap = true; }
    pos += 12 + (size_t)n;
  }
  if (!have_model || !have_palette) return m;
  if (have_imap) {
    uint8_t inverse[256];
    for (uint32_t i = 0; i < 256; i++) inverse[index_map[i]] = (uint8_t)i;
    uint8_t old[256][4];
    memcpy(old, palette, 1024);
    for (uint32_t i = 0; i < 256; i++) memcpy(palette[i], old[(index_map[i] + 255u) & 0xFFu], 4);
    for (auto& v : m.voxels) v = (uint8_t)(1 + inverse[v]);
  }
  // rotate so that colour index k addresses rgba[k]; entry 0 takes the last colour
  memcpy(m.rgba[0], palette[255], 4);
  for (uint32_t i = 1; i < 256; i++) memcpy(m.rgba[i], palette[i - 1], 4);
  m.ok = true;
  return m;
}
}  // namespace

// Scene::LoadModel, scene.cpp:441-524 (y/z swap, conditional rescale, palette index == material id)
static bool load_model(Scene& s, const char* filename, std::vector<std::shared_ptr<Material>>* materials, int columns, int thickness)
{
  s.ResetGrid();
  const VoxModel m = read_vox(filename);
  if (!m.ok) return false;
  if (m.sx > s.gridsize) {
    s.scaleModel.x *= static_cast<float>(s.gridsize) / static_cast<float>(m.sx);
    s.scaleModel.y *= static_cast<float>(s.gridsize) / static_cast<float>(m.sy);
    s.scaleModel.z *= static_cast<float>(s.gridsize) / static_cast<float>(m.sz);
  }
  for (uint32_t z = 0; z < m.sz; ++z)
    for (uint32_t y = 0; y < m.sy; ++y)
      for (uint32_t x = 0; x < m.sx; ++x) {
        const int scaledX = static_cast<int>(static_cast<float>(x) * s.scaleModel.x);
        const int scaledY = static_cast<int>(static_cast<float>(z) * s.scaleModel.y);
        const int scaledZ = static_cast<int>(static_cast<float>(y) * s.scaleModel.z);
        const uint8_t idx = m.voxels[x + (size_t)y * m.sx + (size_t)z * m.sx * m.sy];
        if (idx == 0) continue;
        if (columns >= 0 && !((int)x >= columns - thickness && (int)x <= columns + thickness)) continue;  // LoadModelPartial, scene.cpp:573
        if (materials) {
          Material& mat = *(*materials)[idx];
          mat.albedo = float3(static_cast<float>(m.rgba[idx][0]) / 255.0f, static_cast<float>(m.rgba[idx][1]) / 255.0f,
                              static_cast<float>(m.rgba[idx][2]) / 255.0f);
          mat.roughness = 1.0f;
        }
        // the reference indexes the grid unchecked; a model larger than the grid in y/z would
        // write out of bounds there - here such voxels are dropped
        if ((uint32_t)scaledX < s.gridsize && (uint32_t)scaledY < s.gridsize && (uint32_t)scaledZ < s.gridsize)
          s.Set((uint32_t)scaledX, (uint32_t)scaledY, (uint32_t)scaledZ, static_cast<MaterialType::MatType>(idx));
      }
  return true;
}
bool Scene::LoadModel(Renderer& renderer, const char* filename) { return load_model(*this, filename, &renderer.materials, -1, 0); }
bool Scene::LoadModelPartial(const char* filename, uint32_t columns, uint32_t thickness)
{
  return load_model(*this, filename, nullptr, (int)columns, (int)thickness);
}

// ---- Camera -----------------------------------------------------------------------------
Camera::Camera(int w, int h) : width(w), height(h)  // template/camera.h:16-26
{
  const float ASPECT = static_cast<float>(w) / static_cast<float>(h);
  camPos = float3(0, 0, -2);
  camTarget = float3(0, 0, -1);
  topLeft = float3(-ASPECT, 1, 0);
  topRight = float3(ASPECT, 1, 0);
  bottomLeft = float3(-ASPECT, -1, 0);
  ahead = normalize(camTarget - camPos);
  right = float3(1, 0, 0);
  up = float3(0, 1, 0);
}
void Camera::UpdateFrame()  // template/camera.h:121-123,150-170 with no key pressed
{
  const float ASPECT = static_cast<float>(width) / static_cast<float>(height);
  ahead = normalize(camTarget - camPos);
  const float3 tmpUp(0, 1, 0);
  right = normalize(cross(tmpUp, ahead));
  up = normalize(cross(ahead, right));
  ahead = normalize(camTarget - camPos);
  right = normalize(cross(tmpUp, ahead));
  up = normalize(cross(ahead, right));
  camTarget = camPos + ahead;
  ahead = normalize(camTarget - camPos);
  up = normalize(cross(ahead, right));
  right = normalize(cross(up, ahead));
  topLeft = camPos + 2 * ahead - ASPECT * right + up;
  topRight = camPos + 2 * ahead + ASPECT * right + up;
  bottomLeft = camPos + 2 * ahead - ASPECT * right - up;
}
void Camera::SetFrustumNormals()
{
  const float ASPECT = static_cast<float>(width) / static_cast<float>(height);
  const float3 leftFrustumDirection = (2 * ahead) - (ASPECT * right);
  const float3 rightFrustumDirection = (2 * ahead) + (ASPECT * right);
  const float3 topFrustumDirection = (2 * ahead) + up;
  const float3 bottomFrustumDirection = (2 * ahead) - up;
  leftNormal = cross(up, leftFrustumDirection);
  rightNormal = cross(rightFrustumDirection, up);
  topNormal = cross(right, topFrustumDirection);
  bottomNormal = cross(bottomFrustumDirection, right);
}
VpPrevCamera Camera::PackPrev() const
{
  VpPrevCamera c;
  auto st = [](float* d, const float3& v) { d[0] = v.x, d[1] = v.y, d[2] = v.z; };
  st(c.cam_pos, camPos), st(c.left_normal, leftNormal), st(c.right_normal, rightNormal), st(c.top_normal, topNormal),
      st(c.bottom_normal, bottomNormal);
  return c;
}
VpCamera Camera::Pack() const
{
  VpCamera c;
  auto st = [](float* d, const float3& v) { d[0] = v.x, d[1] = v.y, d[2] = v.z; };
  st(c.cam_pos, camPos), st(c.top_left, topLeft), st(c.top_right, topRight), st(c.bottom_left, bottomLeft);
  st(c.right, right), st(c.up, up);
  c.focal_distance = focalDistance, c.defocus_jitter = defocusJitter;
  return c;
}

// Renderer::MaterialSetUp table, renderer.cpp:357-443 (16 typed materials then white diffuse up to 256 entries)
void default_materials(std::vector<std::shared_ptr<Material>>& materials)
{
  materials.clear();
  auto push = [&](float3 a, float rough, float emis, float ior) {
    auto m = std::make_shared<Material>(a, rough);
    m->emissiveStrength = emis, m->IOR = ior;
    materials.push_back(m);
  };
  const float3 smokeColor{1.0f, 0.7f, 1.0f};
  push({1, 1, 1}, 1.0f, 0, 1.5f);          // NON_METAL_WHITE
  push({1, 0, 0}, 0.6f, 0, 1.5f);          // NON_METAL_RED
  push({0, 0, 1}, 0.25f, 0, 1.5f);         // NON_METAL_BLUE
  push({0, 1, 0}, 0.0f, 0, 1.5f);          // NON_METAL_GREEN
  push({1.0f, .6f, .8f}, 0.3f, 0, 1.5f);   // NON_METAL_PINK
  push({1, 1, 1}, 1.0f, 0, 1.5f);          // METAL_HIGH
  push({0, 1, 1}, 0.5f, 0, 1.5f);          // METAL_MID
  push({0.9f, 0.9f, 0.9f}, 0.01f, 0, 1.5f);  // METAL_LOW
  push({1, 0.5f, 1}, 1.0f, 0, 1.45f);      // GLASS
  push(smokeColor, 1.0f, 3.0f, 1.0f);
  push(smokeColor, 1.0f, 8.0f, 1.0f);
  push(smokeColor, 1.0f, 12.0f, 1.0f);
  push(smokeColor, 1.0f, 15.0f, 1.0f);
  push(smokeColor, 1.0f, 16.0f, 1.0f);
  push({0, 0, 0}, 1.0f, 22.0f, 1.0f);      // SMOKE_PLAYER
  push(smokeColor, 1.0f, 5.0f, 1.5f);      // EMISSIVE
  // the reference sizes the table to 255 and LoadModel writes entry 255 out of bounds
  // (scene.cpp:516-520); 256 entries make that write well defined (entry 255 is never read)
  while (materials.size() < 256) materials.push_back(std::make_shared<Material>(float3(1, 1, 1), 1.f));
}

bool load_model_into(Scene& s, const char* filename, std::vector<std::shared_ptr<Material>>& materials)
{
  return load_model(s, filename, &materials, -1, 0);
}

}  // namespace voxpop

// ---- C exports used by the Python layer (raytracer-voxpopuli_b200/host.py) ------------------
using namespace voxpop;
extern "C" {

void vph_make_transform(const float cube_pos[3], const float member_pos[3], const float rotation[3], const float scale[3],
                        float matrix[16], float inv_matrix[16], float cube_min[3], float cube_max[3])
{
  Scene s(float3(cube_pos[0], cube_pos[1], cube_pos[2]), 1);
  s.position = float3(member_pos[0], member_pos[1], member_pos[2]);
  s.SetScale(float3(scale[0], scale[1], scale[2]));
  s.SetTransform(float3(rotation[0], rotation[1], rotation[2]));
  memcpy(matrix, s.matrix.cell, 64);
  memcpy(inv_matrix, s.invMatrix.cell, 64);
  cube_min[0] = s.cube.b[0].x, cube_min[1] = s.cube.b[0].y, cube_min[2] = s.cube.b[0].z;
  cube_max[0] = s.cube.b[1].x, cube_max[1] = s.cube.b[1].y, cube_max[2] = s.cube.b[1].z;
}

void vph_default_materials(float out[256 * 6])
{
  std::vector<std::shared_ptr<Material>> m;
  default_materials(m);
  for (int i = 0; i < 256; i++) {
    out[i * 6 + 0] = m[i]->albedo.x, out[i * 6 + 1] = m[i]->albedo.y, out[i * 6 + 2] = m[i]->albedo.z;
    out[i * 6 + 3] = m[i]->roughness, out[i * 6 + 4] = m[i]->emissiveStrength, out[i * 6 + 5] = m[i]->IOR;
  }
}

// Scene(position 0, n) + LoadModel: grid bytes and the material table after the load. 0 on success.
int vph_load_vox(const char* path, uint32_t n, uint8_t* grid_out, float mats_out[256 * 6])
{
  std::vector<std::shared_ptr<Material>> m;
  default_materials(m);
  Scene s(float3(0, 0, 0), n);
  if (!load_model_into(s, path, m)) return -1;
  memcpy(grid_out, s.grid.data(), s.grid.size());
  if (mats_out)
    for (int i = 0; i < 256; i++) {
      mats_out[i * 6 + 0] = m[i]->albedo.x, mats_out[i * 6 + 1] = m[i]->albedo.y, mats_out[i * 6 + 2] = m[i]->albedo.z;
      mats_out[i * 6 + 3] = m[i]->roughness, mats_out[i * 6 + 4] = m[i]->emissiveStrength, mats_out[i * 6 + 5] = m[i]->IOR;
    }
  return 0;
}

void vph_camera_look_at(const float pos[3], const float target[3], int width, int height, float focal_distance, float defocus_jitter,
                        VpCamera* out)
{
  Camera c(width, height);
  c.focalDistance = focal_distance, c.defocusJitter = defocus_jitter;
  c.LookAt(float3(pos[0], pos[1], pos[2]), float3(target[0], target[1], target[2]));
  *out = c.Pack();
}

void vph_camera_prev(const float pos[3], const float target[3], int width, int height, VpPrevCamera* out)
{
  Camera c(width, height);
  c.LookAt(float3(pos[0], pos[1], pos[2]), float3(target[0], target[1], target[2]));
  c.SetFrustumNormals();
  *out = c.PackPrev();
}

float vph_cos_degrees(float d) { return CosDegrees(d); }
void vph_normalize(const float v[3], float out[3])
{
  const float3 n = normalize(float3(v[0], v[1], v[2]));
  out[0] = n.x, out[1] = n.y, out[2] = n.z;
}
}
