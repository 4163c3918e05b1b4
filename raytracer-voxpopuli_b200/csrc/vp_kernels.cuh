// vp_kernels.cuh - kernels of the render path (sm_100a).
//
// K_render   : primary-gen + path walk + fold + accumulate + tonemap + RGB8, one lane per
//              pixel, persistent blocks pulling 8x4-pixel warp tiles from a global counter.
// K_primary  : primary-gen + nearest hit only (bit-exact hit dumps).
// K_nearest / K_occluded / K_trace : batch seams of Renderer::FindNearest / IsOccluded / Trace.
// K_fine / K_coarse : occupancy hierarchy build (per upload / dirty update).
#pragma once
#include "vp_device.cuh"

namespace vp {

struct RenderParams {
  DScene sc;
  VpCamera cam;
  int W, H;
  int max_bounces;
  int spp;
  uint32_t frame0;
  uint32_t seed;
  float aa;
  int row_begin, row_end, row_tile, row_tile_stride, row_tile_phase;
  float4* accum;
  uint32_t* rgb8;
  unsigned long long* counters;  // [0] nearest [1] shadow [2] exit [3] steps [4] primary [5] in_light [6] sky [7] launches
  unsigned int* tile_counter;
  VpHit* hits;
};

constexpr int kTileW = 8, kTileH = 4;  // one warp = one 8x4 pixel tile (coherent primary rays)
constexpr int kBlock = 128;

__device__ __forceinline__ bool row_selected(const RenderParams& p, int y)
{
  if (y < p.row_begin || y >= p.row_end) return false;
  if (p.row_tile > 0 && ((y / p.row_tile) % p.row_tile_stride) != p.row_tile_phase) return false;
  return true;
}

// stage the volume table and (when they fit) the region masks into shared memory
__device__ __forceinline__ Staged stage_scene(const DScene& sc, unsigned char* smem)
{
  DVolume* sv = reinterpret_cast<DVolume*>(smem);
  const int words = sc.nvol * (int)(sizeof(DVolume) / 4);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(sc.vol);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sv);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
  Staged st; st.vol = sv; st.regions = nullptr;
  if (sc.region_words > 0) {
    unsigned long long* rg = reinterpret_cast<unsigned long long*>(smem + (((size_t)words * 4 + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < sc.region_words; i += blockDim.x) rg[i] = __ldg(sc.regions + i);
    st.regions = rg;
  }
  __syncthreads();
  return st;
}

__device__ __forceinline__ void flush_counters(const RenderParams& p, const Counters& cn, unsigned int primary, int in_light)
{
  unsigned long long v[7] = {cn.nearest, cn.shadow, cn.exitray, cn.steps, primary, (unsigned long long)in_light, cn.sky};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned long long x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(p.counters + k, x);
  }
}

template <bool kCountSteps>
__global__ void __launch_bounds__(kBlock) k_render(const __grid_constant__ RenderParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  const int lane = threadIdx.x & 31;
  const int tiles_x = (p.W + kTileW - 1) / kTileW;
  const int tiles_y = (p.H + kTileH - 1) / kTileH;
  const unsigned int ntiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
  const uint32_t npix = (uint32_t)p.W * (uint32_t)p.H;
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  unsigned int primary = 0;
  int in_light = 0;
  Level stack[VP_MAX_LEVELS];
  for (;;) {
    unsigned int tile = 0;
    if (lane == 0) tile = atomicAdd(p.tile_counter, 1u);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= ntiles) break;
    const int ty = (int)(tile / (unsigned int)tiles_x), tx = (int)(tile % (unsigned int)tiles_x);
    const int x = tx * kTileW + (lane & (kTileW - 1));
    const int y = ty * kTileH + (lane / kTileW);
    if (x >= p.W || y >= p.H || !row_selected(p, y)) continue;
    const uint32_t px = (uint32_t)y * (uint32_t)p.W + (uint32_t)x;
    float4 acc = p.accum[px];
    for (int s = 0; s < p.spp; s++) {
      const uint32_t frame = p.frame0 + (uint32_t)s;
      uint32_t seed = pixel_seed(px, frame, npix, p.seed);
      // Renderer::Update scalar pixel loop, renderer.cpp:1851-1881
      const float rx = rnd(seed) * p.aa;
      const float ry = rnd(seed) * p.aa;
      const Ray pr = primary_ray(p.cam, (float)x + rx, (float)y + ry, p.W, p.H, seed);
      const f3 col = trace_path<kCountSteps>(p.sc, svol, pr, p.max_bounces, seed, cn, stack, &in_light);
      if (p.max_bounces >= 0) primary++;
      const float weight = 1.0f / ((float)frame + 1.0f);  // renderer.cpp:1651
      const float invWeight = 1.0f - weight;
      const float w1 = 1.0f - invWeight;                   // BlendColor, renderer.cpp:1630-1633
      acc.x = col.x * w1 + acc.x * invWeight;
      acc.y = col.y * w1 + acc.y * invWeight;
      acc.z = col.z * w1 + acc.z * invWeight;
      acc.w = 0.0f * w1 + acc.w * invWeight;
    }
    p.accum[px] = acc;
    if (p.rgb8) p.rgb8[px] = rgb8(reinhard_jodie(v3(acc.x, acc.y, acc.z)));
  }
  flush_counters(p, cn, primary, in_light);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.counters + 7, 1ull);
}

__global__ void __launch_bounds__(kBlock) k_primary(const __grid_constant__ RenderParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  const int lane = threadIdx.x & 31;
  const int tiles_x = (p.W + kTileW - 1) / kTileW;
  const int tiles_y = (p.H + kTileH - 1) / kTileH;
  const unsigned int ntiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
  const uint32_t npix = (uint32_t)p.W * (uint32_t)p.H;
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  for (;;) {
    unsigned int tile = 0;
    if (lane == 0) tile = atomicAdd(p.tile_counter, 1u);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= ntiles) break;
    const int ty = (int)(tile / (unsigned int)tiles_x), tx = (int)(tile % (unsigned int)tiles_x);
    const int x = tx * kTileW + (lane & (kTileW - 1));
    const int y = ty * kTileH + (lane / kTileW);
    if (x >= p.W || y >= p.H || !row_selected(p, y)) continue;
    const uint32_t px = (uint32_t)y * (uint32_t)p.W + (uint32_t)x;
    uint32_t seed = pixel_seed(px, p.frame0, npix, p.seed);
    const float rx = rnd(seed) * p.aa;
    const float ry = rnd(seed) * p.aa;
    Ray ray = primary_ray(p.cam, (float)x + rx, (float)y + ry, p.W, p.H, seed);
    NearestOut no; no.cell = 0xFFFFFFFFu; no.face = -1;
    const int vol = find_nearest<false>(p.sc, svol, ray, cn, &no);
    VpHit h;
    h.volume = vol; h.t = ray.t;
    h.normal[0] = ray.N.x; h.normal[1] = ray.N.y; h.normal[2] = ray.N.z;
    h.material = (uint8_t)ray.mat; h.inside = (uint8_t)(ray.inside != 0); h.pad = 0;
    h.cell = vol >= 0 ? no.cell : 0xFFFFFFFFu;
    h.face = (int8_t)(vol >= 0 ? no.face : -1);
    p.hits[px] = h;
  }
}

struct BatchParams {
  DScene sc;
  const VpRay* rays;
  const uint32_t* seeds;
  int n;
  int depth;
  VpHit* hits;
  uint8_t* occ;
  float* rgb;
};

__global__ void __launch_bounds__(kBlock) k_nearest(const __grid_constant__ BatchParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const VpRay r = p.rays[i];
    Ray ray; ray.O = ld3(r.o); ray.D = ld3(r.d); ray.N = v3s(0); ray.t = r.t; ray.inside = r.inside != 0; ray.mat = VP_MAT_NONE;
    NearestOut no; no.cell = 0xFFFFFFFFu; no.face = -1;
    const int vol = find_nearest<false>(p.sc, svol, ray, cn, &no);
    VpHit h;
    h.volume = vol; h.t = ray.t;
    h.normal[0] = ray.N.x; h.normal[1] = ray.N.y; h.normal[2] = ray.N.z;
    h.material = (uint8_t)ray.mat; h.inside = (uint8_t)(ray.inside != 0); h.pad = 0;
    h.cell = vol >= 0 ? no.cell : 0xFFFFFFFFu;
    h.face = (int8_t)(vol >= 0 ? no.face : -1);
    p.hits[i] = h;
  }
}

// Renderer::FindNearestPlayer per ray (p.depth carries currentChunk)
__global__ void __launch_bounds__(kBlock) k_nearest_player(const __grid_constant__ BatchParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const VpRay r = p.rays[i];
    Ray ray; ray.O = ld3(r.o); ray.D = ld3(r.d); ray.N = v3s(0); ray.t = r.t; ray.inside = r.inside != 0; ray.mat = VP_MAT_NONE;
    NearestOut no; no.cell = 0xFFFFFFFFu; no.face = -1;
    const int vol = find_nearest_player<false>(p.sc, svol, ray, p.depth, cn, &no);
    VpHit h;
    h.volume = vol; h.t = ray.t;
    h.normal[0] = ray.N.x; h.normal[1] = ray.N.y; h.normal[2] = ray.N.z;
    h.material = (uint8_t)ray.mat; h.inside = (uint8_t)(ray.inside != 0); h.pad = 0;
    h.cell = vol >= 0 ? no.cell : 0xFFFFFFFFu;
    h.face = (int8_t)(vol >= 0 ? no.face : -1);
    p.hits[i] = h;
  }
}

// the autofocus ray of Tick (renderer.cpp:1987-1991): one thread
__global__ void k_focus(const __grid_constant__ RenderParams p, float* out)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  uint32_t seed = p.seed;
  Ray fr = primary_ray(p.cam, (float)(p.W / 2), (float)(p.H / 2), p.W, p.H, seed);
  fr.t = 1e34f;
  unsigned long long steps = 0;
  for (int i = 0; i < p.sc.nvol; i++) {  // scene.FindNearest(focusRay) on the untransformed ray, as the reference does
    const DVolume& v = svol.vol[i];
    const ORay r = make_oray(fr.O, fr.D);
    float thit; uint32_t cell; int mat;
    if (scene_find_nearest<false>(v, region_of(svol, v), r, fr.t, thit, cell, mat, steps)) fr.t = thit;
  }
  const float lo = fr.t < 1e4f ? fr.t : 1e4f;  // clamp(t, -1, 1e4) = fmaxf(-1, fminf(t, 1e4)), tmpl8math.h:2105
  *out = -1.0f > lo ? -1.0f : lo;
}

__global__ void __launch_bounds__(kBlock) k_occluded(const __grid_constant__ BatchParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const VpRay r = p.rays[i];
    p.occ[i] = is_occluded_raw<false>(p.sc, svol, ld3(r.o), ld3(r.d), r.t, cn) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kBlock) k_trace(const __grid_constant__ BatchParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged svol = stage_scene(p.sc, smem);
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  Level stack[VP_MAX_LEVELS];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const VpRay r = p.rays[i];
    Ray ray; ray.O = ld3(r.o); ray.D = ld3(r.d); ray.N = v3s(0); ray.t = r.t; ray.inside = r.inside != 0; ray.mat = VP_MAT_NONE;
    uint32_t seed = p.seeds[i];
    const f3 c = trace_path<false>(p.sc, svol, ray, p.depth, seed, cn, stack, nullptr);
    p.rgb[3 * i] = c.x; p.rgb[3 * i + 1] = c.y; p.rgb[3 * i + 2] = c.z;
  }
}

// ---- occupancy hierarchy build ----------------------------------------------------------
// one thread per 4x4x4 brick in brick-major order (index = P >> 6, see struct DDA); cells outside
// the n^3 grid (padding up to the power-of-two domain) read as empty
__global__ void k_fine(const uint8_t* __restrict__ grid, uint32_t n, uint32_t jr, unsigned long long* __restrict__ fine)
{
  const uint32_t total = 64u << (3 * jr);
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < total; b += gridDim.x * blockDim.x) {
    const uint32_t bx = compact3(b), by = compact3(b >> 1), bz = compact3(b >> 2);  // b = Morton index of the brick (P >> 6)
    unsigned long long m = 0;
    for (uint32_t z = 0; z < 4; z++)
      for (uint32_t y = 0; y < 4; y++)
        for (uint32_t x = 0; x < 4; x++) {
          const uint32_t X = bx * 4 + x, Y = by * 4 + y, Z = bz * 4 + z;
          if (X < n && Y < n && Z < n && grid[X + (Y + Z * n) * n] != VP_MAT_NONE) m |= 1ull << (dilate3(x) | (dilate3(y) << 1) | (dilate3(z) << 2));
        }
    fine[b] = m;
  }
}
// one thread per 16^3 region: bit j <=> brick (region << 6 | j) holds a voxel
__global__ void k_region(const unsigned long long* __restrict__ fine, uint32_t jr, unsigned long long* __restrict__ region)
{
  const uint32_t total = 1u << (3 * jr);
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
    unsigned long long m = 0;
    for (uint32_t j = 0; j < 64; j++)
      if (fine[(r << 6) | j]) m |= 1ull << j;
    region[r] = m;
  }
}

// vp_check_sincos_pin: the inline main path of the fp64 sincos against the library routine, one argument per bit pattern.
// out[0] = differing arguments, out[1] = the smallest differing bit pattern (0xFFFFFFFF when none)
__global__ void k_check_sincos(uint32_t first, uint32_t last, unsigned long long* out)
{
  unsigned long long bad = 0;
  uint32_t first_bad = 0xFFFFFFFFu;
  const unsigned long long n = (unsigned long long)last - first + 1ull;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t bits = first + (uint32_t)i;
    const float x = __uint_as_float(bits);
    float s0, c0, s1, c1;
    p_sincosf_2pi(x, &s0, &c0);
    p_sincosf_lib(x, &s1, &c1);
    if (__float_as_uint(s0) != __float_as_uint(s1) || __float_as_uint(c0) != __float_as_uint(c1)) { bad++; first_bad = min(first_bad, bits); }
  }
  for (int o = 16; o; o >>= 1) {
    bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
    first_bad = min(first_bad, __shfl_xor_sync(0xFFFFFFFFu, first_bad, o));
  }
  if ((threadIdx.x & 31) == 0 && bad) { atomicAdd(out, bad); atomicMin(out + 1, (unsigned long long)first_bad); }
}

}  // namespace vp
