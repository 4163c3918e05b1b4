/* vp_oracle.c - TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the render hot path of Tycro-Games/Raytracer-VoxPopuli
 * (Renderer::Update -> Trace -> FindNearest / IsOccluded / Illumination -> accumulator).
 * It is the checker for the CUDA library: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product never does.
 *
 * Parity status: PINNED.  The reference holds no golden vectors or tests of its own
 * (SURVEY.md section 4), so this restatement is pinned against the reference ITSELF,
 * compiled headless into oracle/_ref/libvpref.so by oracle/ref/build_ref.py, on the same
 * scene bytes / camera / seeds (tests/test_oracle_vs_ref.py, bit-exact) and against the
 * committed fixtures under tests/golden/ that were generated from that build
 * (tools/make_golden.py).
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root).  Arithmetic is fp32, evaluated in the reference's operation order;
 * compile with -ffp-contract=off and without -ffast-math.  Determinism pins shared by
 * all three implementations (reference build, this file, CUDA):
 *   - per-pixel RNG reseed  s = WangHash(pixel + frame*W*H + seed), 0 -> 0x9E3779B9;
 *   - exact 1/x for the per-volume reciprocal (the reference's SlowReciprocal branch,
 *     renderer.cpp:966-971) instead of _mm_rcp_ps;
 *   - the scalar pixel loop of Update (renderer.cpp:1851-1881);
 *   - sinf/cosf/powf/expf = correctly rounded, computed as (float)f64fn((double)x);
 *   - function-call argument evaluation right-to-left as g++/MSVC do
 *     (DiffuseReflection: z gets the first draw, tmpl8math.h:2523).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/voxpop.h"

typedef struct { float x, y, z; } f3;

typedef struct {
  f3 O, D, rD, Dsign; /* template/scene.h:96-138 */
  f3 N;               /* rayNormal */
  float t;
  int inside;         /* isInsideGlass */
  int mat;            /* indexMaterial, 255 = NONE */
} Ray;

typedef struct {
  uint8_t* grid;
  uint32_t n;
  f3 bmin, bmax;
  float M[16], iM[16];
} Volume;

typedef struct vpo_ctx {
  Volume* vol; int nvol;
  VpMaterial mat[256];
  VpPointLight* pl; int npl;
  VpAreaLight* al; int nal;
  VpSpotLight* sl; int nsl;
  VpDirLight dl;
  int area_samples;
  VpSphere* sph; int nsph;
  VpTriangle* tri; int ntri;
  float* sky; int skyw, skyh; float sky_contrib; int sky_active;
  int in_light; /* Renderer::inLight, renderer.h:231 */
} vpo_ctx;

typedef struct { uint64_t nearest, shadow, exitray, steps, sky; uint32_t lastcell; int lastface; int in_light; } Counters;

typedef struct { vpo_ctx* c; uint32_t seed; Counters cnt; } Thread;

/* ---------- small vector helpers (tmpl8math.h operators, componentwise) ---------- */
static inline f3 v3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 v3s(float s) { return v3(s, s, s); }
static inline f3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
static inline f3 add(f3 a, f3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub(f3 a, f3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul(f3 a, f3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 divv(f3 a, f3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline f3 muls(f3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline f3 smul(float s, f3 a) { return v3(s * a.x, s * a.y, s * a.z); } /* tmpl8math.h:1760 b*a.x */
static inline f3 divs(f3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline f3 neg(f3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* tmpl8math.h:2251 */
static inline float stdmin(float a, float b) { return (b < a) ? b : a; }  /* std::min */
static inline float stdmax(float a, float b) { return (a < b) ? b : a; }  /* std::max */
/* tmpl8math.h:2350-2354 + :411-414: v * (1.0f / sqrtf(dot)) */
static inline f3 normalize3(f3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return muls(v, inv); }
/* static_cast<int>(float) as x86 cvttss2si does it: NaN / out of range -> INT_MIN */
static inline int f2i(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000u; }

/* correctly-rounded libm pins (see header) */
static inline float p_sinf(float x) { return (float)sin((double)x); }
static inline float p_cosf(float x) { return (float)cos((double)x); }
static inline float p_expf(float x) { return (float)exp((double)x); }
static inline float p_powf(float x, float y) { return (float)pow((double)x, (double)y); }

#define VP_PI 3.14159265358979323846264f /* template/common.h:8 */
#define VP_INVPI 0.31830988618379067153777f
#define VP_INV2PI 0.15915494309189533576888f

/* ---------- RNG: template/tmpl8math.cpp:16-38,119-157 ---------- */
static uint32_t wang_hash(uint32_t s) /* tmpl8math.cpp:20-27 */
{
  s = (s ^ 61) ^ (s >> 16);
  s *= 9, s = s ^ (s >> 4);
  s *= 0x27d4eb2d;
  s = s ^ (s >> 15);
  return s;
}
static inline uint32_t pixel_seed(uint32_t px, uint32_t frame, uint32_t npix, uint32_t seed)
{
  uint32_t v = wang_hash(px + frame * npix + seed);
  return v ? v : 0x9E3779B9u;
}
static inline uint32_t rnd_uint(Thread* T) /* tmpl8math.cpp:119-125 */
{
  uint32_t s = T->seed;
  s ^= s << 13; s ^= s >> 17; s ^= s << 5;
  T->seed = s;
  return s;
}
static inline float rnd(Thread* T) { return (float)rnd_uint(T) * 2.3283064365387e-10f; } /* tmpl8math.cpp:130-133 */

/* tmpl8math.cpp:76-92 RandomDirection: braced init -> x,y,z draw order */
static f3 random_direction(Thread* T)
{
  for (;;) {
    f3 p; p.x = rnd(T); p.y = rnd(T); p.z = rnd(T);
    if (dot3(p, p) < 1) return normalize3(p);
  }
}
/* tmpl8math.h:2502-2512 RandomSphereSample */
static f3 random_sphere_sample(Thread* T)
{
  const float theta = rnd(T) * 2 * VP_PI;
  const float phi = rnd(T) * VP_PI;
  const float r = rnd(T);
  const float x = r * p_sinf(phi) * p_cosf(theta);
  const float y = r * p_sinf(phi) * p_sinf(theta);
  const float z = r * p_cosf(phi);
  return v3(x, y, z);
}
/* tmpl8math.h:2519-2529 DiffuseReflection; make_float3(a,b,c) arguments evaluate right to left */
static f3 diffuse_reflection(Thread* T, f3 N)
{
  f3 R;
  do {
    R.z = rnd(T) * 2 - 1; R.y = rnd(T) * 2 - 1; R.x = rnd(T) * 2 - 1;
  } while (dot3(R, R) > 1);
  if (dot3(R, N) < 0) R = muls(R, -1.0f);
  return normalize3(R);
}
/* tmpl8math.cpp:137-142 RandomPointInCircle */
static void random_point_in_circle(Thread* T, float* px, float* py)
{
  const float r = sqrtf(rnd(T));
  const float theta = 2 * VP_PI * rnd(T);
  *px = r * p_cosf(theta); *py = r * p_sinf(theta);
}

/* ---------- transforms: tmpl8math.cpp:316-402 ---------- */
/* scalar: make_float4(a,w) * M, each row ((c0*x + c1*y) + c2*z) + c3*w  (tmpl8math.cpp:329-343) */
static inline f3 xform_scalar(f3 a, float w, const float* M)
{
  return v3(M[0] * a.x + M[1] * a.y + M[2] * a.z + M[3] * w,
            M[4] * a.x + M[5] * a.y + M[6] * a.z + M[7] * w,
            M[8] * a.x + M[9] * a.y + M[10] * a.z + M[11] * w);
}
/* SSE position: (x*c0 + y*c1) + (z*c2 + 1*c3)  (tmpl8math.cpp:369-380) */
static inline f3 xform_pos_sse(f3 a, const float* M)
{
  return v3((a.x * M[0] + a.y * M[1]) + (a.z * M[2] + 1.0f * M[3]),
            (a.x * M[4] + a.y * M[5]) + (a.z * M[6] + 1.0f * M[7]),
            (a.x * M[8] + a.y * M[9]) + (a.z * M[10] + 1.0f * M[11]));
}
/* SSE vector: (x*c0 + y*c1) + z*c2  (tmpl8math.cpp:393-402) */
static inline f3 xform_vec_sse(f3 a, const float* M)
{
  return v3((a.x * M[0] + a.y * M[1]) + a.z * M[2],
            (a.x * M[4] + a.y * M[5]) + a.z * M[6],
            (a.x * M[8] + a.y * M[9]) + a.z * M[10]);
}
/* template/scene.cpp:49-58 / :59-80 ComputeDsign: 1.0 where the sign bit is set */
static inline float dsign1(float v)
{
  uint32_t u; memcpy(&u, &v, 4);
  return (((float)(u >> 31) * 2 - 1) + 1) * 0.5f;
}
static inline f3 compute_dsign(f3 D) { return v3(dsign1(D.x), dsign1(D.y), dsign1(D.z)); }
static inline f3 rcp3(f3 D) { return v3(1 / D.x, 1 / D.y, 1 / D.z); }

/* template/scene.cpp:83-93 normalising Ray ctor */
static Ray make_ray(f3 O, f3 dir)
{
  Ray r;
  r.O = O;
  r.D = normalize3(dir);
  r.rD = rcp3(r.D);
  r.Dsign = compute_dsign(r.D);
  r.N = v3s(0); r.t = 1e34f; r.inside = 0; r.mat = VP_MAT_NONE;
  return r;
}

/* tmpl8math.cpp:446-487 OffsetRay (Ray Tracing Gems ch.6) */
static inline float ias(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline int32_t fai(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static f3 offset_ray(f3 p, f3 n)
{
  const int ofx = f2i(256.0f * n.x), ofy = f2i(256.0f * n.y), ofz = f2i(256.0f * n.z);
  /* two's complement wrap like the x86 add */
  const float pix = ias((int32_t)((uint32_t)fai(p.x) + (uint32_t)((p.x < 0) ? -ofx : ofx)));
  const float piy = ias((int32_t)((uint32_t)fai(p.y) + (uint32_t)((p.y < 0) ? -ofy : ofy)));
  const float piz = ias((int32_t)((uint32_t)fai(p.z) + (uint32_t)((p.z < 0) ? -ofz : ofz)));
  const float origin = 1.0f / 32.0f, fscale = 1.0f / 65536.0f;
  return v3(fabsf(p.x) < origin ? p.x + fscale * n.x : pix,
            fabsf(p.y) < origin ? p.y + fscale * n.y : piy,
            fabsf(p.z) < origin ? p.z + fscale * n.z : piz);
}

/* ---------- voxel volume traversal: template/scene.cpp ---------- */
typedef struct { int sx, sy, sz; uint32_t X, Y, Z; float t; f3 tdelta, tmax; } DDA;

/* template/scene.cpp:205-210 */
static inline int cube_contains(const Volume* v, f3 p)
{
  return p.x >= v->bmin.x && p.y >= v->bmin.y && p.z >= v->bmin.z &&
         p.x <= v->bmax.x && p.y <= v->bmax.y && p.z <= v->bmax.z;
}
/* template/scene.cpp:166-202 */
static float cube_intersect(const Volume* v, const Ray* r)
{
  const int sx = r->D.x < 0, sy = r->D.y < 0, sz = r->D.z < 0;
  const f3 b[2] = {v->bmin, v->bmax};
  float tmin_x = (b[sx].x - r->O.x) * r->rD.x;
  float tmax_x = (b[1 - sx].x - r->O.x) * r->rD.x;
  const float tmin_y = (b[sy].y - r->O.y) * r->rD.y;
  const float tmax_y = (b[1 - sy].y - r->O.y) * r->rD.y;
  if (tmin_x > tmax_y || tmin_y > tmax_x) return 1e34f;
  tmin_x = stdmax(tmin_x, tmin_y);
  tmax_x = stdmin(tmax_x, tmax_y);
  const float tmin_z = (b[sz].z - r->O.z) * r->rD.z;
  const float tmax_z = (b[1 - sz].z - r->O.z) * r->rD.z;
  if (tmin_x > tmax_z || tmin_z > tmax_x) return 1e34f;
  tmin_x = stdmax(tmin_x, tmin_z);
  if (tmin_x > 0) return tmin_x;
  return 1e34f;
}
static inline int clampi(int f, int a, int b) { int m = f < b ? f : b; return a > m ? a : m; } /* tmpl8math.h:2110 */

/* template/scene.cpp:719-749 Setup3DDDA */
static int setup_dda(const Volume* v, const Ray* ray, DDA* s)
{
  s->t = 0;
  if (!cube_contains(v, ray->O)) {
    s->t = cube_intersect(v, ray);
    if (s->t > 1e33f) return 0;
  }
  const f3 vmin = v->bmin;
  const f3 ext = sub(v->bmax, v->bmin);
  const float gs = (float)v->n;
  const float cell = 1.0f / gs;
  s->sx = f2i(1 - ray->Dsign.x * 2); s->sy = f2i(1 - ray->Dsign.y * 2); s->sz = f2i(1 - ray->Dsign.z * 2);
  /* gridsizeFloat * ((O - min) + (t + 0.00005f) * D) / ext */
  const f3 pos = divv(smul(gs, add(sub(ray->O, vmin), smul(s->t + 0.00005f, ray->D))), ext);
  const f3 planes = muls(sub(v3(ceilf(pos.x), ceilf(pos.y), ceilf(pos.z)), ray->Dsign), cell);
  const int nm1 = (int)(v->n - 1);
  s->X = (uint32_t)clampi(f2i(pos.x), 0, nm1);
  s->Y = (uint32_t)clampi(f2i(pos.y), 0, nm1);
  s->Z = (uint32_t)clampi(f2i(pos.z), 0, nm1);
  s->tdelta = mul(smul(cell, v3((float)s->sx, (float)s->sy, (float)s->sz)), ray->rD);
  s->tmax = mul(sub(mul(planes, ext), sub(ray->O, vmin)), ray->rD);
  return 1;
}

/* The shared Amanatides-Woo advance (template/scene.cpp:776-803 and the four copies):
 * returns 0 when the ray left the grid (unsigned wrap test).  `occl` selects the
 * IsOccluded ordering (:1021-1044) where s.t is only updated when the step stays inside. */
static inline int dda_advance(DDA* s, uint32_t n, int occl)
{
  if (s->tmax.x < s->tmax.y) {
    if (s->tmax.x < s->tmax.z) {
      if (occl) { if ((s->X += s->sx) >= n) return 0; s->t = s->tmax.x; s->tmax.x += s->tdelta.x; }
      else { s->t = s->tmax.x; s->X += s->sx; if (s->X >= n) return 0; s->tmax.x += s->tdelta.x; }
    } else {
      if (occl) { if ((s->Z += s->sz) >= n) return 0; s->t = s->tmax.z; s->tmax.z += s->tdelta.z; }
      else { s->t = s->tmax.z; s->Z += s->sz; if (s->Z >= n) return 0; s->tmax.z += s->tdelta.z; }
    }
  } else {
    if (s->tmax.y < s->tmax.z) {
      if (occl) { if ((s->Y += s->sy) >= n) return 0; s->t = s->tmax.y; s->tmax.y += s->tdelta.y; }
      else { s->t = s->tmax.y; s->Y += s->sy; if (s->Y >= n) return 0; s->tmax.y += s->tdelta.y; }
    } else {
      if (occl) { if ((s->Z += s->sz) >= n) return 0; s->t = s->tmax.z; s->tmax.z += s->tdelta.z; }
      else { s->t = s->tmax.z; s->Z += s->sz; if (s->Z >= n) return 0; s->tmax.z += s->tdelta.z; }
    }
  }
  return 1;
}
static inline uint32_t cell_index(const Volume* v, const DDA* s) { return s->X + s->Y * v->n + s->Z * v->n * v->n; }
static inline float fracf1(float v) { return v - floorf(v); }
static inline int face_code1(float v) { return v > 0 ? 1 : (v < 0 ? 2 : 0); }

/* template/scene.cpp:121-148 Ray::GetNormalVoxel (ray is in object space) */
static f3 get_normal_voxel(Thread* T, const Ray* ray, const Volume* v)
{
  const f3 I1 = muls(add(ray->O, smul(ray->t, ray->D)), (float)v->n);
  const f3 fG = v3(fracf1(I1.x), fracf1(I1.y), fracf1(I1.z));
  const f3 d = v3(stdmin(fG.x, 1.0f - fG.x), stdmin(fG.y, 1.0f - fG.y), stdmin(fG.z, 1.0f - fG.z));
  const float mind = stdmin(stdmin(d.x, d.y), d.z);
  const f3 sign = v3(ray->Dsign.x * 2 - 1, ray->Dsign.y * 2 - 1, ray->Dsign.z * 2 - 1);
  f3 n = v3(mind == d.x ? sign.x : 0.0f, mind == d.y ? sign.y : 0.0f, mind == d.z ? sign.z : 0.0f);
  T->cnt.lastface = face_code1(n.x) + 3 * face_code1(n.y) + 9 * face_code1(n.z);
  return normalize3(xform_scalar(n, 0.0f, v->M));
}

/* template/scene.cpp:751-811 Scene::FindNearest */
static int scene_find_nearest(Thread* T, const Volume* v, Ray* ray)
{
  DDA s;
  if (!setup_dda(v, ray, &s)) return 0;
  while (s.t < ray->t) {
    const uint8_t cell = v->grid[cell_index(v, &s)];
    T->cnt.steps++;
    if (cell != VP_MAT_NONE && s.t < ray->t) {
      ray->t = s.t;
      ray->N = get_normal_voxel(T, ray, v);
      ray->mat = cell;
      T->cnt.lastcell = cell_index(v, &s);
      return 1;
    }
    if (!dda_advance(&s, v->n, 0)) break;
  }
  return 0;
}
/* template/scene.cpp:1009-1047 Scene::IsOccluded */
static int scene_is_occluded(Thread* T, const Volume* v, const Ray* ray)
{
  DDA s;
  if (!setup_dda(v, ray, &s)) return 0;
  while (s.t < ray->t) {
    const uint8_t cell = v->grid[cell_index(v, &s)];
    T->cnt.steps++;
    if (cell != VP_MAT_NONE) return s.t < ray->t;
    if (!dda_advance(&s, v->n, 1)) return 0;
  }
  return 0;
}
/* template/scene.cpp:875-939 FindMaterialExit (smoke=0) and :941-1006 FindSmokeExit (smoke=1) */
static int scene_find_exit(Thread* T, const Volume* v, Ray* ray, int smoke, int mat)
{
  DDA s;
  T->cnt.exitray++;
  if (!setup_dda(v, ray, &s)) return 0;
  for (;;) {
    const uint8_t cell = v->grid[cell_index(v, &s)];
    T->cnt.steps++;
    const int out = smoke ? (cell > 14 || cell < 9) : (cell != mat);
    if (out) {
      ray->t = s.t;
      ray->N = get_normal_voxel(T, ray, v);
      ray->mat = cell;
      return 1;
    }
    if (!dda_advance(&s, v->n, 0)) break;
  }
  ray->t = s.t;
  return 0;
}

/* ---------- analytic primitives: src/BVH/Shapes.h ---------- */
static void sphere_hit(const VpSphere* sp, Ray* r) /* Shapes.h:12-42 */
{
  const f3 toRay = sub(r->O, ld3(sp->center));
  const float b = dot3(toRay, r->D);
  const float c = dot3(toRay, toRay) - (sp->radius * sp->radius);
  const float disc = b * b - c;
  if (c > 0.0f && b > 0.0f) return;
  if (disc < 0) return;
  const float len = -b - sqrtf(disc);
  if (len > r->t) return;
  if (len < 0) return;
  const f3 ip = add(r->O, smul(len, r->D));
  const f3 on = divs(sub(ip, ld3(sp->center)), sp->radius);
  const int outside = dot3(r->D, on) < 0;
  r->N = outside ? on : neg(on);
  r->inside = !outside;
  r->t = len;
  r->mat = (int)(sp->material & 255);
}
static int sphere_is_hit(const VpSphere* sp, const Ray* r) /* Shapes.h:44-62 */
{
  const f3 toRay = sub(r->O, ld3(sp->center));
  const float b = dot3(toRay, r->D);
  const float c = dot3(toRay, toRay) - (sp->radius * sp->radius);
  const float disc = b * b - c;
  if (c > 0.0f && b > 0.0f) return 0;
  if (disc < 0) return 0;
  const float len = -b - sqrtf(disc);
  if (len < 0) return 0;
  if (len > r->t) return 0;
  return 1;
}
static inline f3 cross3(f3 a, f3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* Shapes.h:79-110 (hit) / :112-138 (is_hit) */
static int triangle_test(const VpTriangle* tr, Ray* ray, int anyhit)
{
  const f3 pos = ld3(tr->position);
  const f3 p1 = add(pos, ld3(tr->vertex0)), p2 = add(pos, ld3(tr->vertex1)), p3 = add(pos, ld3(tr->vertex2));
  const f3 e1 = sub(p2, p1), e2 = sub(p3, p1);
  const f3 h = cross3(ray->D, e2);
  const float a = dot3(e1, h);
  if (a > -0.0001f && a < 0.0001f) return 0;
  const float f = 1 / a;
  const f3 s = sub(ray->O, p1);
  const float u = f * dot3(s, h);
  if (u < 0 || u > 1) return 0;
  const f3 q = cross3(s, e1);
  const float v = f * dot3(ray->D, q);
  if (v < 0 || u + v > 1) return 0;
  const float t = f * dot3(e2, q);
  if (anyhit) {
    if (t < 0.0001f) return 0;
    if (t > ray->t) return 0;
    return 1;
  }
  if (t > 0.0001f) {
    if (ray->t > t) {
      ray->t = t;
      ray->mat = (int)(tr->material & 255);
      const f3 n = normalize3(cross3(e1, e2));
      const int outside = dot3(ray->D, n) < 0;
      ray->N = outside ? n : neg(n);
    }
  }
  return 0;
}

/* ---------- scene query: renderer.cpp:946-1018 / :209-243 ---------- */
static int find_nearest(Thread* T, Ray* ray)
{
  vpo_ctx* c = T->c;
  int voxelIndex = -2;
  T->cnt.nearest++;
  for (int i = 0; i < c->nvol; i++) {
    const Volume* v = &c->vol[i];
    const f3 bO = ray->O, bD = ray->D, brD = ray->rD, bS = ray->Dsign; /* Ray copy ctor, scene.cpp:16-27 */
    ray->O = xform_pos_sse(ray->O, v->iM);
    ray->D = xform_vec_sse(ray->D, v->iM);
    ray->rD = rcp3(ray->D);              /* SlowReciprocal pin */
    ray->Dsign = compute_dsign(ray->D);
    if (scene_find_nearest(T, v, ray)) voxelIndex = i;
    ray->O = bO; ray->D = bD; ray->rD = brD; ray->Dsign = bS; /* CopyToPrevRay, scene.cpp:8-14 */
  }
  if (c->nsph + c->ntri > 0) {
    Ray sh = make_ray(ray->O, ray->D); /* renderer.cpp:997 */
    for (int i = 0; i < c->nsph; i++) sphere_hit(&c->sph[i], &sh);
    for (int i = 0; i < c->ntri; i++) triangle_test(&c->tri[i], &sh, 0);
    if (ray->t > sh.t) {
      ray->t = sh.t; ray->mat = sh.mat; ray->N = sh.N; ray->inside = sh.inside;
      voxelIndex = -1;
    }
  }
  return voxelIndex;
}
static int is_occluded(Thread* T, Ray* ray)
{
  vpo_ctx* c = T->c;
  T->cnt.shadow++;
  for (int i = 0; i < c->nvol; i++) {
    const Volume* v = &c->vol[i];
    const f3 bO = ray->O, bD = ray->D, brD = ray->rD, bS = ray->Dsign;
    ray->O = xform_scalar(ray->O, 1.0f, v->iM);
    ray->D = xform_scalar(ray->D, 0.0f, v->iM);
    ray->rD = rcp3(ray->D);
    ray->Dsign = compute_dsign(ray->D);
    const int occ = scene_is_occluded(T, v, ray);
    ray->O = bO; ray->D = bD; ray->rD = brD; ray->Dsign = bS;
    if (occ) return 1;
  }
  for (int i = 0; i < c->nsph; i++) if (sphere_is_hit(&c->sph[i], ray)) return 1;
  for (int i = 0; i < c->ntri; i++) if (triangle_test(&c->tri[i], ray, 1)) return 1;
  return 0;
}

/* ---------- direct lighting: renderer.cpp:102-207,315-338,738-764 ---------- */
static inline f3 albedo(const vpo_ctx* c, int m) { return ld3(c->mat[m & 255].albedo); }
static inline f3 isect(const Ray* r) { return add(r->O, smul(r->t, r->D)); } /* scene.h:80-83 */

static f3 point_light(Thread* T, const Ray* ray, const VpPointLight* L) /* renderer.cpp:102-132 */
{
  const f3 ip = isect(ray);
  const f3 dir = sub(ld3(L->position), ip);
  const float dst = sqrtf(dot3(dir, dir));
  const f3 dn = muls(dir, 1.0f / dst);
  const f3 n = ray->N;
  const float ct = dot3(dn, n);
  if (ct <= 0.0f) return v3s(0);
  const f3 li = muls(smul(stdmax(0.0f, ct), ld3(L->color)), 1.0f / (dst * dst));
  const f3 org = offset_ray(ip, n);
  const f3 k = albedo(T->c, ray->mat);
  Ray sh = make_ray(org, dn);
  sh.t = dst;
  if (is_occluded(T, &sh)) return v3s(0);
  return mul(li, k);
}
static f3 spot_light(Thread* T, const Ray* ray, const VpSpotLight* L) /* renderer.cpp:134-161 */
{
  const f3 ip = isect(ray);
  const f3 dir = sub(ld3(L->position), ip);
  const float dst = sqrtf(dot3(dir, dir));
  const f3 dn = divs(dir, dst);
  const f3 n = ray->N;
  const float ct = dot3(dn, ld3(L->direction));
  if (ct <= L->angle) return v3s(0);
  const float alpha = 1.0f - (1.0f - ct) * 1.0f / (1.0f - L->angle);
  const f3 li = divs(smul(stdmax(0.0f, ct), ld3(L->color)), dst * dst);
  const f3 k = albedo(T->c, ray->mat);
  Ray sh = make_ray(offset_ray(ip, n), dn);
  sh.t = dst;
  if (is_occluded(T, &sh)) return v3s(0);
  return muls(mul(li, k), alpha);
}
static f3 area_light(Thread* T, const Ray* ray, const VpAreaLight* L) /* renderer.cpp:163-207 */
{
  const f3 ip = isect(ray);
  const f3 n = ray->N;
  const f3 center = ld3(L->position);
  const float radius = L->radius;
  f3 inc = v3s(0);
  const f3 k = albedo(T->c, ray->mat);
  const f3 point = offset_ray(ip, n);
  const int ns = T->c->area_samples;
  for (int i = 0; i < ns; i++) {
    f3 rp = random_direction(T);
    rp = muls(rp, radius);
    rp = add(rp, center);
    const f3 dir = sub(rp, ip);
    const float dst = sqrtf(dot3(dir, dir));
    const f3 dn = muls(dir, 1 / dst);
    const float ct = dot3(dn, n);
    if (ct <= 0) continue;
    Ray sh = make_ray(point, dn);
    sh.t = dst;
    if (is_occluded(T, &sh)) continue;
    /* cosTheta * color * colorMultiplier * (radius*radius) * PI4 / (dst*dst); PI4 expands to PI*4.0f unparenthesised */
    f3 li = smul(ct, ld3(L->color));
    li = muls(li, L->color_multiplier);
    li = muls(li, radius * radius);
    li = muls(li, VP_PI);
    li = muls(li, 4.0f);
    li = divs(li, dst * dst);
    inc = add(inc, li);
  }
  inc = divs(inc, (float)ns);
  return mul(inc, k);
}
static f3 dir_light(Thread* T, const Ray* ray, const VpDirLight* L) /* renderer.cpp:315-338 */
{
  const f3 ip = isect(ray);
  const f3 dir = neg(ld3(L->direction));
  const f3 n = ray->N;
  const float ct = dot3(dir, n);
  if (ct <= 0) return v3s(0);
  const f3 li = smul(stdmax(0.0f, ct), ld3(L->color));
  const f3 k = albedo(T->c, ray->mat);
  Ray sh = make_ray(offset_ray(ip, n), dir);
  if (is_occluded(T, &sh)) return v3s(0);
  return mul(li, k);
}
static f3 illumination(Thread* T, const Ray* ray) /* renderer.cpp:738-764 */
{
  vpo_ctx* c = T->c;
  const size_t lightCount = (size_t)c->npl + c->nsl + c->nal + 1; /* renderer.cpp:2290-2296 */
  const float pick = rnd(T) * (float)lightCount;                  /* Rand(range), tmpl8math.cpp:149-152 */
  const size_t idx = (size_t)pick;
  f3 inc;
  if (idx < (size_t)c->npl) inc = point_light(T, ray, &c->pl[idx]);
  else if (idx < (size_t)c->nal + c->npl) inc = area_light(T, ray, &c->al[idx - c->npl]);
  else if (idx < (size_t)c->nal + c->nsl + c->npl) inc = spot_light(T, ray, &c->sl[idx - c->nal - c->npl]);
  else inc = dir_light(T, ray, &c->dl);
  return muls(inc, (float)lightCount);
}

/* ---------- BSDF helpers: renderer.cpp:913-925,1588-1616 ---------- */
static inline f3 reflect3(f3 d, f3 n) { return sub(d, muls(smul(2, n), dot3(n, d))); } /* direction - 2*normal*dot */
static f3 refract3(f3 d, f3 n, float ratio)
{
  const float ct = stdmin(dot3(neg(d), n), 1.0f);
  const f3 rper = smul(ratio, add(d, smul(ct, n)));
  const f3 rpar = smul(-sqrtf(fabsf(1.0f - dot3(rper, rper))), n);
  return add(rper, rpar);
}
static inline float schlick(float cosine, float ior)
{
  float r0 = (1 - ior) / (1 + ior);
  r0 = r0 * r0;
  return r0 + (1 - r0) * p_powf((1 - cosine), 5);
}
static inline float schlick_nonmetal(float cosine) { const float r0 = 0.04f; return r0 + (1 - r0) * p_powf((1 - cosine), 5); }
static f3 absorption(f3 color, float intensity, float dist) /* renderer.cpp:1596-1608 */
{
  const f3 flipped = v3(1.0f - color.x, 1.0f - color.y, 1.0f - color.z);
  const f3 e = smul(-dist * intensity, flipped);
  return v3(p_expf(e.x), p_expf(e.y), p_expf(e.z));
}

/* tmpl8math.cpp:405-443 */
static float atan2_approx2(float y, float x)
{
  const float ONEQTR_PI = (float)(VP_PI / 4.0);
  const float THRQTR_PI = (float)(3.0 * VP_PI / 4.0);
  float r, angle;
  float abs_y = fabsf(y) + 1e-10f;
  if (x < 0.0f) { r = (x + abs_y) / (abs_y - x); angle = THRQTR_PI; }
  else { r = (x - abs_y) / (x + abs_y); angle = ONEQTR_PI; }
  angle += (0.1963f * r * r - 0.9817f) * r;
  return (y < 0.0f) ? -angle : angle;
}
static float fast_acos(float x)
{
  float negate = (float)(x < 0);
  x = fabsf(x);
  float ret = (float)-0.0187293;
  ret = ret * x;
  ret = ret + 0.0742610f;
  ret = ret * x;
  ret = ret - 0.2121144f;
  ret = ret * x;
  ret = ret + 1.5707288f;
  ret = ret * sqrtf(1.0f - x);
  ret = ret - 2 * negate * ret;
  return negate * 3.14159265358979f + ret;
}
/* renderer.cpp:2308-2326 SampleSky.  The reference has no upper clamp; indices past the
 * array (unreachable for |d.y| <= 1) are clamped here instead of reading out of bounds. */
static f3 sample_sky(const vpo_ctx* c, f3 d)
{
  if (!c->sky_active) return v3(0.392f, 0.584f, 0.829f);
  const float uF = (float)c->skyw * atan2_approx2(d.z, d.x) * VP_INV2PI - 0.5f;
  const int u = f2i(uF);
  const float vF = (float)c->skyh * fast_acos(d.y) * VP_INVPI - 0.5f;
  const int v = f2i(vF);
  int32_t lin = (int32_t)((uint32_t)u + (uint32_t)v * (uint32_t)c->skyw);
  if (lin < 0) lin = 0;
  const int64_t maxtexel = (int64_t)c->skyw * c->skyh - 1;
  if (lin > maxtexel) lin = (int32_t)maxtexel;
  const size_t idx = (size_t)lin * 3;
  return smul(c->sky_contrib, v3(c->sky[idx], c->sky[idx + 1], c->sky[idx + 2]));
}

/* GLASS bounce: renderer.cpp:1146-1205 (Trace) == :1360-1418 (TraceDialectric), up to the recursive call.
 * Returns the attenuation colour and the next ray. */
static f3 glass_bounce(Thread* T, Ray* ray, int voxIndex, Ray* next)
{
  vpo_ctx* c = T->c;
  const int m = ray->mat;
  f3 color = v3s(1.0f);
  int inGlass = ray->inside;
  const float ior = c->mat[m].ior;
  const float ratio = inGlass ? ior : 1.0f / ior;
  int insideVolume = 1;
  if (inGlass) {
    color = albedo(c, m);
    if (voxIndex >= 0) { /* reference indexes voxelVolumes[voxIndex] unguarded (UB for -1) */
      const Volume* v = &c->vol[voxIndex];
      const f3 bO = ray->O, bD = ray->D, brD = ray->rD, bS = ray->Dsign;
      ray->O = xform_scalar(ray->O, 1.0f, v->iM);
      ray->D = xform_scalar(ray->D, 0.0f, v->iM);
      ray->rD = rcp3(ray->D);
      ray->Dsign = compute_dsign(ray->D);
      insideVolume = scene_find_exit(T, v, ray, 0, 8);
      ray->O = bO; ray->D = bD; ray->rD = brD; ray->Dsign = bS;
    }
  }
  if (!insideVolume) { ray->O = add(ray->O, muls(ray->D, ray->t)); ray->t = 0; }
  const float ct = stdmin(dot3(neg(ray->D), ray->N), 1.0f);
  const float st = sqrtf(1.0f - ct * ct);
  const int cannot = ratio * st > 1.0f;
  f3 rdir, rn;
  if (cannot || schlick(ct, ratio) > rnd(T)) { rdir = reflect3(ray->D, ray->N); rn = ray->N; }
  else { rdir = refract3(ray->D, ray->N, ratio); inGlass = !inGlass; rn = neg(ray->N); }
  *next = make_ray(offset_ray(isect(ray), rn), rdir);
  next->inside = inGlass;
  return color;
}

/* SMOKE bounce: renderer.cpp:1210-1310 (Trace) == :1425-1499 (TraceSmoke), up to the recursive call. */
static f3 smoke_bounce(Thread* T, Ray* ray, int voxIndex, Ray* next)
{
  vpo_ctx* c = T->c;
  const int m = ray->mat;
  f3 color = v3s(1.0f);
  int inGlass = ray->inside;
  const float ratio = 1.0f;
  int insideVolume = 1;
  float intensity = 0, dist = 0;
  if (voxIndex == 0) { /* player volume: inLight side channel :1228-1240 */
    const f3 inc = illumination(T, ray);
    if (dot3(inc, inc) > 16.0f) T->cnt.in_light = 1;
  }
  if (inGlass) {
    intensity = c->mat[m].emissive;
    color = albedo(c, m);
    if (voxIndex >= 0) {
      const Volume* v = &c->vol[voxIndex];
      const f3 bO = ray->O, bD = ray->D, brD = ray->rD, bS = ray->Dsign;
      ray->O = xform_scalar(ray->O, 1.0f, v->iM);
      ray->D = xform_scalar(ray->D, 0.0f, v->iM);
      ray->rD = rcp3(ray->D);
      ray->Dsign = compute_dsign(ray->D);
      insideVolume = scene_find_exit(T, v, ray, 1, 0);
      ray->O = bO; ray->D = bD; ray->rD = brD; ray->Dsign = bS;
    }
    dist = ray->t;
  }
  const float threshold = rnd(T) * 100 - intensity;
  if (rnd(T) * dist > threshold) {
    const float lo = ray->t * .45f;
    const float rr = lo + rnd(T) * (ray->t - lo); /* Rand(min,max), tmpl8math.cpp:154-157 */
    ray->O = add(ray->O, muls(ray->D, rr));
    ray->D = random_direction(T);
    ray->t = 0;
  }
  color = absorption(color, intensity, dist);
  if (!insideVolume) { ray->O = add(ray->O, muls(ray->D, ray->t)); ray->t = 0; }
  const f3 rdir = refract3(ray->D, ray->N, ratio);
  inGlass = !inGlass;
  const f3 rn = neg(ray->N);
  *next = make_ray(offset_ray(isect(ray), rn), rdir);
  next->inside = inGlass;
  return color;
}

/* ---------- Renderer::Trace: renderer.cpp:1076-1328 ---------- */
static f3 trace(Thread* T, Ray* ray, int depth)
{
  vpo_ctx* c = T->c;
  if (depth < 0) return v3s(0);
  const int voxIndex = find_nearest(T, ray);
  if (ray->mat == VP_MAT_NONE) { T->cnt.sky++; return sample_sky(c, ray->D); }
  const int m = ray->mat;
  if (m >= 5 && m <= 7) { /* METAL_HIGH/MID/LOW :1103-1114 */
    const f3 refl = reflect3(ray->D, ray->N);
    Ray nr = make_ray(offset_ray(isect(ray), ray->N), add(refl, smul(c->mat[m].roughness, random_sphere_sample(T))));
    return mul(trace(T, &nr, depth - 1), albedo(c, m));
  }
  if (m <= 4) { /* NON_METAL_* :1117-1144 */
    f3 color = v3s(0);
    if (rnd(T) > schlick_nonmetal(dot3(neg(ray->D), ray->N))) {
      const f3 rdir = add(ray->N, random_sphere_sample(T)); /* RandomLambertianReflectionVector */
      const f3 inc = illumination(T, ray);
      Ray nr = make_ray(offset_ray(isect(ray), ray->N), rdir);
      color = add(color, inc);
      color = add(color, mul(trace(T, &nr, depth - 1), albedo(c, m)));
    } else {
      const f3 refl = reflect3(ray->D, ray->N);
      Ray nr = make_ray(offset_ray(isect(ray), ray->N), add(refl, smul(c->mat[m].roughness, random_sphere_sample(T))));
      color = trace(T, &nr, depth - 1);
    }
    return color;
  }
  if (m == 8) { /* GLASS :1146-1208 */
    Ray nr;
    const f3 color = glass_bounce(T, ray, voxIndex, &nr);
    return mul(trace(T, &nr, depth - 1), color);
  }
  if (m >= 9 && m <= 14) { /* SMOKE_* :1210-1313 */
    Ray nr;
    const f3 color = smoke_bounce(T, ray, voxIndex, &nr);
    return mul(trace(T, &nr, depth - 1), color);
  }
  if (m == 15) /* EMISSIVE :1315-1316 */
    return muls(albedo(c, m), c->mat[m].emissive);
  { /* default: palette materials 16..254 :1319-1326 */
    const f3 rdir = diffuse_reflection(T, ray->N);
    const f3 inc = illumination(T, ray);
    Ray nr = make_ray(offset_ray(isect(ray), ray->N), rdir);
    return mul(add(trace(T, &nr, depth - 1), inc), albedo(c, m));
  }
}

/* ---------- Renderer::TraceReproject family: renderer.cpp:1330-1586 ---------- */
typedef struct { f3 albedo, illum; } AlbIll; /* AlbedoIlluminationData, renderer.h:10-19; GetColor() = albedo * illumination */
static inline f3 ai_color(AlbIll a) { return mul(a.albedo, a.illum); }
static AlbIll ai(f3 a, f3 i) { AlbIll r; r.albedo = a; r.illum = i; return r; }

static AlbIll sample_sky_reproject(const vpo_ctx* c, f3 d) /* renderer.cpp:2328-2346: same texel arithmetic as SampleSky */
{
  return ai(sample_sky(c, d), v3s(1.0f));
}

static AlbIll trace_reproject(Thread* T, Ray* ray, int depth)
{
  vpo_ctx* c = T->c;
  if (depth < 0) return ai(v3s(0.0f), v3s(0.0f)); /* {0.0f, 0} :1523-1526 */
  const int voxIndex = find_nearest(T, ray);
  if (ray->mat == VP_MAT_NONE) { T->cnt.sky++; return sample_sky_reproject(c, ray->D); }
  const int m = ray->mat;
  if (m >= 5 && m <= 7) { /* TraceMetal :1330-1340 */
    const f3 refl = reflect3(ray->D, ray->N);
    Ray nr = make_ray(offset_ray(isect(ray), ray->N), add(refl, smul(c->mat[m].roughness, random_sphere_sample(T))));
    const f3 ill = ai_color(trace_reproject(T, &nr, depth - 1));
    return ai(albedo(c, m), ill);
  }
  if (m <= 4) { /* TraceNonMetal :1342-1358: no Fresnel draw, always the Lambert branch */
    f3 ill = v3s(0);
    const f3 rdir = add(ray->N, random_sphere_sample(T)); /* RandomLambertianReflectionVector */
    const f3 inc = illumination(T, ray);
    Ray nr = make_ray(offset_ray(isect(ray), ray->N), rdir);
    ill = add(ill, inc);
    ill = add(ill, ai_color(trace_reproject(T, &nr, depth - 1)));
    return ai(albedo(c, m), ill);
  }
  if (m == 8) { /* TraceDialectric :1360-1423 */
    Ray nr;
    const f3 color = glass_bounce(T, ray, voxIndex, &nr);
    return ai(color, ai_color(trace_reproject(T, &nr, depth - 1)));
  }
  if (m >= 9 && m <= 14) { /* TraceSmoke :1425-1501 */
    Ray nr;
    const f3 color = smoke_bounce(T, ray, voxIndex, &nr);
    return ai(color, ai_color(trace_reproject(T, &nr, depth - 1)));
  }
  if (m == 15) /* TraceEmmision :1503-1506 */
    return ai(muls(albedo(c, m), c->mat[m].emissive), v3s(1.0f));
  { /* TraceModelMaterials :1508-1518 */
    const f3 rdir = diffuse_reflection(T, ray->N);
    const f3 inc = illumination(T, ray);
    Ray nr = make_ray(offset_ray(isect(ray), ray->N), rdir);
    f3 ill = inc;
    ill = add(ill, ai_color(trace_reproject(T, &nr, depth - 1)));
    return ai(albedo(c, m), ill);
  }
}

/* ---------- camera + frame: template/camera.h:85-101, renderer.cpp:1851-1881,2220-2240 ---------- */
static Ray primary_ray(Thread* T, const VpCamera* cam, float x, float y, int W, int H)
{
  const float u = x * (1.0f / (float)W);
  const float v = y * (1.0f / (float)H);
  const f3 TL = ld3(cam->top_left), TR = ld3(cam->top_right), BL = ld3(cam->bottom_left), cp = ld3(cam->cam_pos);
  const f3 P = add(add(TL, smul(u, sub(TR, TL))), smul(v, sub(BL, TL)));
  float cx, cy;
  random_point_in_circle(T, &cx, &cy);
  const float jx = cx * cam->defocus_jitter / (float)W; /* float2 * float / int */
  const float jy = cy * cam->defocus_jitter / (float)W;
  const f3 focal = add(cp, smul(cam->focal_distance, normalize3(sub(P, cp))));
  const f3 org = add(add(cp, smul(jx, ld3(cam->right))), smul(jy, ld3(cam->up)));
  return make_ray(org, sub(focal, org));
}
static inline float lerp1(float a, float b, float t) { return a + t * (b - a); } /* tmpl8math.h:2210 */
static f3 reinhard_jodie(f3 c) /* renderer.cpp:2220-2232 */
{
  const float lum = dot3(c, v3(0.2126f, 0.7152f, 0.0722f));
  const f3 rh = divv(c, v3(1.0f + c.x, 1.0f + c.y, 1.0f + c.z));
  const f3 la = divs(c, 1.0f + lum);
  return v3(lerp1(la.x, rh.x, rh.x), lerp1(la.y, rh.y, rh.y), lerp1(la.z, rh.z, rh.z));
}
static uint32_t rgb8(f3 v) /* template/precomp.h:382-385 */
{
  uint32_t r = (uint32_t)(int64_t)(255.0f * stdmin(1.0f, v.x));
  uint32_t g = (uint32_t)(int64_t)(255.0f * stdmin(1.0f, v.y));
  uint32_t b = (uint32_t)(int64_t)(255.0f * stdmin(1.0f, v.z));
  return (r << 16) + (g << 8) + b;
}
static inline int row_selected(const VpFrameParams* fp, int y)
{
  int rb = fp->row_begin, re = fp->row_end;
  if (rb == 0 && re == 0) re = fp->height;
  if (y < rb || y >= re) return 0;
  if (fp->row_tile > 0 && ((y / fp->row_tile) % fp->row_tile_stride) != fp->row_tile_phase) return 0;
  return 1;
}
static void fill_hit(VpHit* h, const Ray* ray, int vol, const Counters* cn)
{
  h->volume = vol; h->t = ray->t;
  h->normal[0] = ray->N.x; h->normal[1] = ray->N.y; h->normal[2] = ray->N.z;
  h->material = (uint8_t)ray->mat; h->inside = (uint8_t)(ray->inside != 0); h->pad = 0;
  if (vol >= 0) { h->cell = cn->lastcell; h->face = (int8_t)cn->lastface; }
  else { h->cell = 0xFFFFFFFFu; h->face = -1; }
}
static Ray ray_from_abi(const VpRay* v)
{
  Ray r;
  r.O = ld3(v->o); r.D = ld3(v->d); r.rD = rcp3(r.D); r.Dsign = compute_dsign(r.D);
  r.N = v3s(0); r.t = v->t; r.inside = v->inside != 0; r.mat = VP_MAT_NONE;
  return r;
}

/* ======================= exported API (vpo_*) ======================= */
vpo_ctx* vpo_create(int device)
{
  (void)device;
  vpo_ctx* c = (vpo_ctx*)calloc(1, sizeof(vpo_ctx));
  for (int i = 0; i < 256; i++) {
    c->mat[i].albedo[0] = c->mat[i].albedo[1] = c->mat[i].albedo[2] = 1.0f;
    c->mat[i].roughness = 1.0f; c->mat[i].emissive = 0.0f; c->mat[i].ior = 1.5f;
  }
  c->dl.direction[0] = 1;
  c->area_samples = 3;
  c->sky_contrib = 1.0f;
  return c;
}
static void free_volumes(vpo_ctx* c)
{
  for (int i = 0; i < c->nvol; i++) free(c->vol[i].grid);
  free(c->vol); c->vol = NULL; c->nvol = 0;
}
void vpo_destroy(vpo_ctx* c)
{
  if (!c) return;
  free_volumes(c);
  free(c->pl); free(c->al); free(c->sl); free(c->sph); free(c->tri); free(c->sky);
  free(c);
}
int vpo_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void vpo_set_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}
int vpo_set_volumes(vpo_ctx* c, const VpVolume* v, int n)
{
  free_volumes(c);
  c->vol = (Volume*)calloc((size_t)(n > 0 ? n : 1), sizeof(Volume));
  c->nvol = n;
  for (int i = 0; i < n; i++) {
    const size_t bytes = (size_t)v[i].n * v[i].n * v[i].n;
    c->vol[i].grid = (uint8_t*)malloc(bytes);
    memcpy(c->vol[i].grid, v[i].grid, bytes);
    c->vol[i].n = v[i].n;
    c->vol[i].bmin = ld3(v[i].cube_min); c->vol[i].bmax = ld3(v[i].cube_max);
    memcpy(c->vol[i].M, v[i].matrix, 64); memcpy(c->vol[i].iM, v[i].inv_matrix, 64);
  }
  return 0;
}
int vpo_set_materials(vpo_ctx* c, const VpMaterial* m, int n)
{
  if (n > 256) return VP_EINVAL;
  memcpy(c->mat, m, sizeof(VpMaterial) * (size_t)n);
  return 0;
}
#define VPO_COPY(dst, cnt, src, n, T) do { free(dst); dst = (T*)malloc(sizeof(T) * (size_t)((n) > 0 ? (n) : 1)); \
  if ((n) > 0) memcpy(dst, src, sizeof(T) * (size_t)(n)); cnt = (n); } while (0)
int vpo_set_lights(vpo_ctx* c, const VpPointLight* p, int np, const VpAreaLight* a, int na, const VpSpotLight* s, int ns,
                   const VpDirLight* d, int area_samples)
{
  VPO_COPY(c->pl, c->npl, p, np, VpPointLight);
  VPO_COPY(c->al, c->nal, a, na, VpAreaLight);
  VPO_COPY(c->sl, c->nsl, s, ns, VpSpotLight);
  if (d) c->dl = *d;
  c->area_samples = area_samples;
  return 0;
}
int vpo_set_prims(vpo_ctx* c, const VpSphere* s, int ns, const VpTriangle* t, int nt)
{
  VPO_COPY(c->sph, c->nsph, s, ns, VpSphere);
  VPO_COPY(c->tri, c->ntri, t, nt, VpTriangle);
  return 0;
}
int vpo_set_sky(vpo_ctx* c, const float* rgb, int w, int h, float contribution, int active)
{
  free(c->sky); c->sky = NULL;
  if (rgb && w > 0 && h > 0) {
    c->sky = (float*)malloc(sizeof(float) * 3 * (size_t)w * h);
    memcpy(c->sky, rgb, sizeof(float) * 3 * (size_t)w * h);
  }
  c->skyw = w; c->skyh = h; c->sky_contrib = contribution; c->sky_active = active && c->sky;
  return 0;
}

int vpo_render(vpo_ctx* c, const VpCamera* cam, const VpFrameParams* fp, float* accum, uint32_t* out8, VpStats* stats)
{
  const int W = fp->width, H = fp->height;
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
  uint64_t nearest = 0, shadow = 0, exitray = 0, steps = 0, primary = 0, sky = 0;
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
  for (int s = 0; s < fp->spp; s++) {
    const uint32_t frame = fp->frame_index + (uint32_t)s;
    const float weight = 1.0f / ((float)frame + 1.0f); /* renderer.cpp:1651 */
    const float invWeight = 1.0f - weight;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nearest, shadow, exitray, steps, primary, sky)
    for (int y = 0; y < H; y++) {
      if (!row_selected(fp, y)) continue;
      Thread T; memset(&T, 0, sizeof(T)); T.c = c;
      for (int x = 0; x < W; x++) {
        const uint32_t px = (uint32_t)y * (uint32_t)W + (uint32_t)x;
        T.seed = pixel_seed(px, frame, npix, fp->seed);
        const float rx = rnd(&T) * fp->aa_strength;
        const float ry = rnd(&T) * fp->aa_strength;
        Ray pr = primary_ray(&T, cam, (float)x + rx, (float)y + ry, W, H);
        const f3 col = trace(&T, &pr, fp->max_bounces);
        float* a = accum + 4 * (size_t)px;
        /* BlendColor(pixel, acc, invWeight) = pixel*(1-invWeight) + acc*invWeight, renderer.cpp:1630-1633 */
        const float w1 = 1.0f - invWeight;
        const float r = col.x * w1 + a[0] * invWeight;
        const float g = col.y * w1 + a[1] * invWeight;
        const float b = col.z * w1 + a[2] * invWeight;
        const float w = 0.0f * w1 + a[3] * invWeight;
        a[0] = r; a[1] = g; a[2] = b; a[3] = w;
        if (out8) out8[px] = rgb8(reinhard_jodie(v3(r, g, b)));
        if (fp->max_bounces >= 0) primary++;
      }
      nearest += T.cnt.nearest; shadow += T.cnt.shadow; exitray += T.cnt.exitray; steps += T.cnt.steps; sky += T.cnt.sky;
      if (T.cnt.in_light) {
#pragma omp atomic write
        c->in_light = 1;
      }
    }
  }
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  if (stats) {
    stats->primary_rays = primary; stats->secondary_rays = nearest - primary; stats->shadow_rays = shadow;
    stats->exit_rays = exitray; stats->dda_steps = steps; stats->sky_samples = sky; stats->device_ms = (t1 - t0) * 1e3; stats->kernel_launches = 0;
  }
  return 0;
}

/* ---------- reprojection frame: Renderer::Tick static-camera branch, renderer.cpp:1997-2101 ---------- */
static Ray primary_ray_nodof(const VpCamera* cam, float x, float y, int W, int H) /* template/camera.h:103-110 */
{
  const float u = x * (1.0f / (float)W);
  const float v = y * (1.0f / (float)H);
  const f3 TL = ld3(cam->top_left), TR = ld3(cam->top_right), BL = ld3(cam->bottom_left), cp = ld3(cam->cam_pos);
  const f3 P = add(add(TL, smul(u, sub(TR, TL))), smul(v, sub(BL, TL)));
  return make_ray(cp, sub(P, cp));
}
static inline float stdminf(float a, float b) { return a < b ? a : b; } /* fminf, tmpl8math.h:401 */
static inline float fmaxf8(float a, float b) { return a > b ? a : b; }  /* fmaxf, tmpl8math.h:406 */
static inline float clamp1(float f, float a, float b) { return fmaxf8(a, stdminf(f, b)); } /* tmpl8math.h:2105 */
static inline f3 rgb_to_ycocg(f3 c) /* renderer.cpp:832-838 */
{
  const float Y = dot3(c, v3(1, 2, 1)) * 0.25f;
  const float Co = dot3(c, v3(2, 0, -2)) * 0.25f + (0.5f * 256.0f / 255.0f);
  const float Cg = dot3(c, v3(-1, 2, -1)) * 0.25f + (0.5f * 256.0f / 255.0f);
  return v3(Y, Co, Cg);
}
static inline f3 ycocg_to_rgb(f3 c) /* renderer.cpp:841-850 */
{
  const float Y = c.x;
  const float Co = c.y - (0.5f * 256.0f / 255.0f);
  const float Cg = c.z - (0.5f * 256.0f / 255.0f);
  return v3(Y + Co - Cg, Y + Cg, Y - Co - Cg);
}
static inline int valid_screen(int x, int y, int W, int H) /* Renderer::IsValidScreen on int2 -> float2, renderer.cpp:1640-1643 */
{
  return (float)x >= 0.0f && (float)x < (float)W && (float)y >= 0.0f && (float)y < (float)H;
}
static f3 sample_history(const float* hist, float u, float v, int W, int H) /* renderer.cpp:777-830 */
{
  u -= 1.0f / (float)W / 2.0f; /* HALF_PIXEL_W, camera.h:7 */
  v -= 1.0f / (float)H / 2.0f;
  const float px = u * (float)W, py = v * (float)H;
  const int tlx = f2i(px), tly = f2i(py);
  const float fx = px - (float)tlx, fy = py - (float)tly;
  const float gx = 1 - fx, gy = 1 - fy;
  const int v1 = valid_screen(tlx, tly, W, H), v2 = valid_screen(tlx + 1, tly, W, H);
  const int v3_ = valid_screen(tlx, tly + 1, W, H), v4 = valid_screen(tlx + 1, tly + 1, W, H);
  float w1 = v1 ? gx * gy : 0.0f, w2 = v2 ? fx * gy : 0.0f, w3 = v3_ ? gx * fy : 0.0f, w4 = v4 ? fx * fy : 0.0f;
  const float total = w1 + w2 + w3 + w4;
  const float rcp = 1.0f / total;
  w1 *= rcp; w2 *= rcp; w3 *= rcp; w4 *= rcp;
  f3 avg = v3s(0.0f);
  if (v1) avg = add(avg, smul(w1, ld3(hist + 3 * ((size_t)tly * W + tlx))));
  if (v2) avg = add(avg, smul(w2, ld3(hist + 3 * ((size_t)tly * W + tlx + 1))));
  if (v3_) avg = add(avg, smul(w3, ld3(hist + 3 * ((size_t)(tly + 1) * W + tlx))));
  if (v4) avg = add(avg, smul(w4, ld3(hist + 3 * ((size_t)(tly + 1) * W + tlx + 1))));
  return avg;
}
static f3 clamp_history(f3 hist, f3 sample, const float* illum, int x, int y, int W, int H) /* renderer.cpp:855-910 */
{
  static const int off[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
  sample = rgb_to_ycocg(sample);
  hist = rgb_to_ycocg(hist);
  unsigned n = 1;
  f3 avg = sample, var = mul(sample, sample);
  for (int i = 0; i < 8; i++) {
    const int qx = x + off[i][0], qy = y + off[i][1];
    if (valid_screen(qx, qy, W, H)) {
      const f3 f = rgb_to_ycocg(ld3(illum + 3 * ((size_t)qx + (size_t)qy * W)));
      avg = add(avg, f);
      var = add(var, mul(f, f));
      n++;
    }
  }
  const float inv = 1.0f / (float)n;
  avg = muls(avg, inv);
  var = muls(var, inv);
  const float sig = 0.75f;
  const f3 sigma = v3(sqrtf(stdmax(0.0f, var.x - (avg.x * avg.x))), sqrtf(stdmax(0.0f, var.y - (avg.y * avg.y))),
                      sqrtf(stdmax(0.0f, var.z - (avg.z * avg.z))));
  const f3 lo = sub(avg, smul(sig, sigma)), hi = add(avg, smul(sig, sigma));
  hist = v3(clamp1(hist.x, lo.x, hi.x), clamp1(hist.y, lo.y, hi.y), clamp1(hist.z, lo.z, hi.z));
  hist = ycocg_to_rgb(hist);
  return v3(fmaxf8(hist.x, 0.0f), fmaxf8(hist.y, 0.0f), fmaxf8(hist.z, 0.0f));
}
static float history_weight(int m) /* renderer.cpp:2051-2086 */
{
  if (m <= 4) return 0.8f;             /* NON_METAL_* */
  if (m >= 5 && m <= 8) return 0.5f;   /* METAL_*, GLASS */
  if (m >= 9 && m <= 14) return 0.9f;  /* SMOKE_* */
  if (m == 15) return 0.0f;            /* EMISSIVE */
  return 0.9f;
}

int vpo_render_reproject(vpo_ctx* c, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp, const float* hist_in,
                         float* hist_out, uint32_t* out8, float* albedo_out, float* illum_out, VpStats* stats)
{
  const int W = fp->width, H = fp->height;
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
  if (!hist_in || !hist_out || hist_in == hist_out) return 1;
  float* alb = albedo_out ? albedo_out : (float*)malloc((size_t)npix * 12);
  float* ill = illum_out ? illum_out : (float*)malloc((size_t)npix * 12);
  float* rayP = (float*)malloc((size_t)npix * 12); /* RayDataReproject: intersectionPoint + materialIndex are what pass 2 reads */
  uint8_t* rayM = (uint8_t*)malloc(npix);
  uint64_t nearest = 0, shadow = 0, exitray = 0, steps = 0, primary = 0, sky = 0;
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nearest, shadow, exitray, steps, primary, sky)
  for (int y = 0; y < H; y++) { /* pass 1, renderer.cpp:2003-2022 */
    /* row sharding (not in the reference): a shard also traces the rows next to its own, pass 2 reads their illumination */
    if (!(row_selected(fp, y) || (y > 0 && row_selected(fp, y - 1)) || (y + 1 < H && row_selected(fp, y + 1)))) continue;
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    for (int x = 0; x < W; x++) {
      const uint32_t px = (uint32_t)y * (uint32_t)W + (uint32_t)x;
      T.seed = pixel_seed(px, fp->frame_index, npix, fp->seed);
      Ray pr = primary_ray_nodof(cam, (float)x, (float)y, W, H);
      const AlbIll d = trace_reproject(&T, &pr, fp->max_bounces);
      alb[3 * px] = d.albedo.x; alb[3 * px + 1] = d.albedo.y; alb[3 * px + 2] = d.albedo.z;
      ill[3 * px] = d.illum.x; ill[3 * px + 1] = d.illum.y; ill[3 * px + 2] = d.illum.z;
      const f3 P = isect(&pr); /* RayDataReproject::GetRayInfo, renderer.h:31-34 */
      rayP[3 * px] = P.x; rayP[3 * px + 1] = P.y; rayP[3 * px + 2] = P.z;
      rayM[px] = (uint8_t)pr.mat;
      if (fp->max_bounces >= 0) primary++;
    }
    nearest += T.cnt.nearest; shadow += T.cnt.shadow; exitray += T.cnt.exitray; steps += T.cnt.steps; sky += T.cnt.sky;
    if (T.cnt.in_light) {
#pragma omp atomic write
      c->in_light = 1;
    }
  }
  const f3 pc = ld3(prev->cam_pos);
  const f3 nl = ld3(prev->left_normal), nr_ = ld3(prev->right_normal), nt = ld3(prev->top_normal), nb = ld3(prev->bottom_normal);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : nearest, shadow, exitray, steps, sky)
  for (int y = 0; y < H; y++) { /* pass 2, renderer.cpp:2023-2099 */
    if (!row_selected(fp, y)) continue;
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    for (int x = 0; x < W; x++) {
      const uint32_t px = (uint32_t)x + (uint32_t)y * (uint32_t)W;
      const f3 P = ld3(rayP + 3 * px);
      const f3 newS = ld3(ill + 3 * px);
      /* Camera::PointToUV, template/camera.h:34-50 */
      const f3 delta = sub(P, pc);
      const float dl = dot3(nl, delta), dr = dot3(nr_, delta), dt = dot3(nt, delta), db = dot3(nb, delta);
      const float u = dl / (dl + dr) + 1.0f / (float)W / 2.0f;
      const float v = dt / (dt + db) + 1.0f / (float)H / 2.0f;
      f3 fin;
      int reuse = u >= 0.0f && u < 1.0f && v >= 0.0f && v < 1.0f; /* IsValid :1635-1638 */
      if (reuse) { /* IsOccludedPrevFrame :767-774 */
        const f3 dirTo = sub(P, pc);
        const f3 dn = normalize3(dirTo);
        const f3 pos = offset_ray(P, neg(dn));
        const f3 dd = sub(pos, pc);
        Ray occ = make_ray(pc, dn);
        occ.t = sqrtf(dot3(dd, dd)); /* length() */
        reuse = !is_occluded(&T, &occ);
      }
      if (reuse) {
        f3 h = sample_history(hist_in, u, v, W, H);
        h = clamp_history(h, newS, ill, x, y, W, H);
        const float w = history_weight(rayM[px]);
        fin = add(newS, smul(w, sub(h, newS))); /* lerp, tmpl8math.h:2220 */
      } else fin = newS;
      hist_out[3 * px] = fin.x; hist_out[3 * px + 1] = fin.y; hist_out[3 * px + 2] = fin.z;
      if (out8) out8[px] = rgb8(reinhard_jodie(mul(ld3(alb + 3 * px), fin)));
    }
    nearest += T.cnt.nearest; shadow += T.cnt.shadow; exitray += T.cnt.exitray; steps += T.cnt.steps; sky += T.cnt.sky;
  }
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  if (stats) {
    stats->primary_rays = primary; stats->secondary_rays = nearest - primary; stats->shadow_rays = shadow;
    stats->exit_rays = exitray; stats->dda_steps = steps; stats->sky_samples = sky; stats->device_ms = (t1 - t0) * 1e3; stats->kernel_launches = 0;
  }
  if (!albedo_out) free(alb);
  if (!illum_out) free(ill);
  free(rayP); free(rayM);
  return 0;
}

int vpo_trace_primary(vpo_ctx* c, const VpCamera* cam, const VpFrameParams* fp, VpHit* hits)
{
  const int W = fp->width, H = fp->height;
  const uint32_t npix = (uint32_t)W * (uint32_t)H;
#pragma omp parallel for schedule(dynamic, 1)
  for (int y = 0; y < H; y++) {
    if (!row_selected(fp, y)) continue;
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    for (int x = 0; x < W; x++) {
      const uint32_t px = (uint32_t)y * (uint32_t)W + (uint32_t)x;
      T.seed = pixel_seed(px, fp->frame_index, npix, fp->seed);
      const float rx = rnd(&T) * fp->aa_strength;
      const float ry = rnd(&T) * fp->aa_strength;
      Ray pr = primary_ray(&T, cam, (float)x + rx, (float)y + ry, W, H);
      T.cnt.lastcell = 0xFFFFFFFFu; T.cnt.lastface = -1;
      const int vol = find_nearest(&T, &pr);
      fill_hit(&hits[px], &pr, vol, &T.cnt);
    }
  }
  return 0;
}
int vpo_find_nearest(vpo_ctx* c, const VpRay* rays, int n, VpHit* hits)
{
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    Ray r = ray_from_abi(&rays[i]);
    T.cnt.lastcell = 0xFFFFFFFFu; T.cnt.lastface = -1;
    const int vol = find_nearest(&T, &r);
    fill_hit(&hits[i], &r, vol, &T.cnt);
  }
  return 0;
}
/* ---------- gameplay side channels ---------- */
int vpo_in_light(vpo_ctx* c, int* out) /* Renderer::inLight; cleared as Tick does once per frame (renderer.cpp:2206) */
{
  *out = c->in_light;
  c->in_light = 0;
  return 0;
}
/* Scene::FindNearestExcept, template/scene.cpp:813-873 */
static int scene_find_nearest_except(Thread* T, const Volume* v, Ray* ray, int lo, int hi)
{
  DDA s;
  if (!setup_dda(v, ray, &s)) return 0;
  while (s.t < ray->t) {
    const uint8_t cell = v->grid[cell_index(v, &s)];
    T->cnt.steps++;
    if (cell != VP_MAT_NONE && s.t < ray->t && (cell < lo || cell > hi)) {
      ray->t = s.t;
      ray->N = get_normal_voxel(T, ray, v);
      ray->mat = cell;
      T->cnt.lastcell = cell_index(v, &s);
      return 1;
    }
    if (!dda_advance(&s, v->n, 0)) break;
  }
  return 0;
}
/* Renderer::FindNearestPlayer, renderer.cpp:1020-1071 */
static int find_nearest_player(Thread* T, Ray* ray, int current_chunk)
{
  vpo_ctx* c = T->c;
  int voxelIndex = -2;
  for (int i = 1; i < c->nvol; i++) { /* skip player */
    const Volume* v = &c->vol[i];
    const f3 bO = ray->O, bD = ray->D, brD = ray->rD, bS = ray->Dsign;
    ray->O = xform_pos_sse(ray->O, v->iM);
    ray->D = xform_vec_sse(ray->D, v->iM);
    ray->rD = rcp3(ray->D); /* SlowReciprocal pin */
    ray->Dsign = compute_dsign(ray->D);
    if (scene_find_nearest_except(T, v, ray, 9, 14)) voxelIndex = i; /* SMOKE_LOW_DENSITY .. SMOKE_PLAYER */
    ray->O = bO; ray->D = bD; ray->rD = brD; ray->Dsign = bS;
  }
  if (current_chunk >= 2 && (uint64_t)(uint32_t)voxelIndex > (uint64_t)c->nvol - 4u) voxelIndex = -1; /* :1068-1069: uint32 against size_t */
  return voxelIndex;
}
int vpo_find_nearest_player(vpo_ctx* c, const VpRay* rays, int n, int current_chunk, VpHit* hits)
{
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    Ray r = ray_from_abi(&rays[i]);
    T.cnt.lastcell = 0xFFFFFFFFu; T.cnt.lastface = -1;
    const int vol = find_nearest_player(&T, &r, current_chunk);
    fill_hit(&hits[i], &r, vol, &T.cnt);
  }
  return 0;
}
/* the autofocus ray of Tick, renderer.cpp:1987-1991 */
int vpo_focus_distance(vpo_ctx* c, const VpCamera* cam, int W, int H, uint32_t seed, float* out)
{
  Thread T; memset(&T, 0, sizeof(T)); T.c = c; T.seed = seed;
  Ray fr = primary_ray(&T, cam, (float)(W / 2), (float)(H / 2), W, H);
  for (int i = 0; i < c->nvol; i++) scene_find_nearest(&T, &c->vol[i], &fr); /* scene.FindNearest(focusRay): untransformed */
  *out = clamp1(fr.t, -1.0f, 1e4f);
  return 0;
}
int vpo_is_occluded(vpo_ctx* c, const VpRay* rays, int n, uint8_t* out)
{
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    Ray r = ray_from_abi(&rays[i]);
    out[i] = (uint8_t)is_occluded(&T, &r);
  }
  return 0;
}
int vpo_trace(vpo_ctx* c, const VpRay* rays, const uint32_t* seeds, int n, int depth, float* rgb)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; i++) {
    Thread T; memset(&T, 0, sizeof(T)); T.c = c;
    Ray r = ray_from_abi(&rays[i]);
    T.seed = seeds[i];
    const f3 col = trace(&T, &r, depth);
    rgb[3 * i] = col.x; rgb[3 * i + 1] = col.y; rgb[3 * i + 2] = col.z;
  }
  return 0;
}

/* small pins, mirrored by vpref_* in oracle/ref/ref_harness.cpp */
void vpo_offset_ray(const float p[3], const float n[3], float out[3])
{
  const f3 o = offset_ray(ld3(p), ld3(n));
  out[0] = o.x; out[1] = o.y; out[2] = o.z;
}
void vpo_sample_sky(vpo_ctx* c, const float d[3], float out[3])
{
  const f3 s = sample_sky(c, ld3(d));
  out[0] = s.x; out[1] = s.y; out[2] = s.z;
}
uint32_t vpo_wang_hash(uint32_t s) { return wang_hash(s); }
void vpo_rng_floats(uint32_t s, int n, float* out)
{
  Thread T; memset(&T, 0, sizeof(T)); T.seed = s;
  for (int i = 0; i < n; i++) out[i] = rnd(&T);
}
void vpo_sampling(uint32_t s, const float nrm[3], float out[15], uint32_t* seed_after)
{
  Thread T; memset(&T, 0, sizeof(T)); T.seed = s;
  const f3 a = random_sphere_sample(&T);
  const f3 b = random_direction(&T);
  const f3 d = diffuse_reflection(&T, ld3(nrm));
  float px, py; random_point_in_circle(&T, &px, &py);
  const f3 l = add(ld3(nrm), random_sphere_sample(&T));
  out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = b.z;
  out[6] = d.x; out[7] = d.y; out[8] = d.z; out[9] = px; out[10] = py; out[11] = 0;
  out[12] = l.x; out[13] = l.y; out[14] = l.z;
  *seed_after = T.seed;
}
void vpo_tonemap(vpo_ctx* c, const float in[4], float out[3], uint32_t* out8)
{
  (void)c;
  const f3 o = reinhard_jodie(v3(in[0], in[1], in[2]));
  out[0] = o.x; out[1] = o.y; out[2] = o.z;
  *out8 = rgb8(o);
}
