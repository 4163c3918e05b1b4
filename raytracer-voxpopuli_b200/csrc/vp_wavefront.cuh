// vp_wavefront.cuh - wavefront formulation of Renderer::Update for sm_100a.
//
// The single-kernel walk (vp_kernels.cuh k_render) keeps 5.8 of 32 lanes busy on room.vox
// (ncu, profiles/r1a_*): path lengths (1..15 segments), DDA trip counts, material branches
// and the optional shadow ray all diverge inside one warp.  Here every pixel-sample is a
// "path" whose state lives in HBM (SoA, 128-bit records) and each stage is its own kernel:
//
//   K1  wf_generate : seed, AA jitter, primary ray                     (1 lane per path)
//   K2a wf_prepare  : world -> object ray + Setup3DDDA of the first volume each ray enters, bounding-sphere
//                     candidates of the others                         (1 lane per ray; passes that overlap others)
//   K2  wf_extend   : Renderer::FindNearest over all volumes            (persistent warps, lane refill, brick walk)
//   K3  wf_shade    : material switch, RNG, next ray, light sample      (1 lane per active path)
//   K4a wf_shadow_mask : bounding-sphere candidates of the shadow rays  (1 lane per request; multi-volume scenes)
//   K4  wf_shadow   : Renderer::IsOccluded for the light samples        (persistent warps, lane refill, brick walk)
//   K6  wf_resolve  : exact back-to-front colour fold, accumulate, tonemap, RGB8
//
// K2/K4 are the hot loops: a warp keeps all lanes inside the DDA stepping loop and, when
// fewer than kRefill lanes are still walking, idle lanes pull the next ray from the stage
// queue (one warp-aggregated atomicAdd) instead of waiting for the slowest lane.
// Queues are plain arrays of path ids appended with ballot/popc warp aggregation.
// The arithmetic of every stage is the same device code the single-kernel path uses
// (vp_device.cuh), so results are bit-identical between the two formulations.
#pragma once
#include "vp_device.cuh"

namespace vp {

constexpr int kWfBlock = 128;
#ifndef VP_WF_MINBLOCKS
#define VP_WF_MINBLOCKS 7
#endif
#ifndef VP_HOIST_PS
#define VP_HOIST_PS 1
#endif
#ifndef VP_BW_UNROLL
#define VP_BW_UNROLL 2
#endif
constexpr int kBwUnroll = VP_BW_UNROLL;        // brick-walk steps between two votes on the number of walking lanes
#ifndef VP_WF_MINBLOCKS_SPARE
#define VP_WF_MINBLOCKS_SPARE VP_WF_MINBLOCKS
#endif
constexpr int kWfMinBlocksSpare = VP_WF_MINBLOCKS_SPARE;
#ifndef VP_WF_MINBLOCKS_SHADOW
#define VP_WF_MINBLOCKS_SHADOW VP_WF_MINBLOCKS
#endif
constexpr int kWfMinBlocksShadow = VP_WF_MINBLOCKS_SHADOW;
constexpr int kWfMinBlocks = VP_WF_MINBLOCKS;  // resident blocks per SM the traversal kernels are compiled for: 72 registers at 7 (no spills in
                                               // the brick-walk kernels); single-volume scenes: 8 blocks / 64 registers is 2.4 % slower, 6 / 80 slower again
#ifndef VP_WF_MINBLOCKS_SHADE
#define VP_WF_MINBLOCKS_SHADE 5
#endif
constexpr int kWfMinBlocksShade = VP_WF_MINBLOCKS_SHADE;
#ifndef VP_WF_MINBLOCKS_MULTI
#define VP_WF_MINBLOCKS_MULTI 7
#endif
constexpr int kWfMinBlocksMulti = VP_WF_MINBLOCKS_MULTI;          // multi-volume scenes: 7 since the round-2 step body (72 registers instead of 64 with spills: cfg 4 -3.7 %, cfg 5 -1.7 %)
constexpr int kRefill = 20;           // refill when fewer lanes than this are still stepping
constexpr int kBulkLanes = 12;        // with at most this many lanes walking, lanes in empty 8^3/16^3 blocks bulk-advance
constexpr int kMaskVolumes = 23;      // scenes up to this size carry the bounding-sphere candidates of a ray in its prepared record
constexpr int kMaxAreaSamples = 8;    // numCheckShadowsAreaLight supported by the wavefront path (reference default 3)

// Per-path state, structure of arrays over `cap` paths.
struct WfState {
  uint32_t cap;
  // The ray of a path lives at its POSITION in the bounce's queue (bounce parity): the shade stage writes the next ray where
  // its append put the path, so prepare / extend / shade read rays, hits and path ids of a bounce coalesced and without
  // waiting for the queue entry first (the path id is only needed for what outlives the bounce: levels, term, requests).
  float4* r0[2]; // [parity][pos]: O.xyz, w = bits: inside(1) | depth+1 (8 bits << 8) | top (8 bits << 16)
  float4* r1[2]; // [parity][pos]: D.xyz (normalised), w = RNG state bits
  float2* h0;    // [pos] hit: t, P bits (cell of the hit inside volume `vol`, brick-major)
  uint32_t* h1;  // [pos] hit: (vol + 2) << 8; primitive hits: mat (8) | 1 << 8 | inside << 30 | 1 << 31
  float4* hn;    // [pos] primitive hits only: normal
  float4* term;  // [path] terminal radiance (sky / emissive / 0), w = number of fold levels of the path
  // fold levels, ONE record per (level, path): w = type (bits 0..2) | material (bits 8..15); xyz = M for type 0 (the
  // multiplier is a computed colour), else A (the light term; M is then the material's albedo, looked up by the resolve
  // stage) - half the bytes of separate M / A arrays
  float4* lv;
  // Shadow requests of a bounce, indexed by POSITION in the bounce's request queue (bounce parity: shadow(k) runs on the
  // lane's side stream while the path is extended and shaded again; shade(k + 2) waits for it before it reuses the slots).
  // The shade stage gets the position from its aggregated append and writes the records there, so the shadow stage finds
  // everything a single-light request needs with ONE trip to memory after its ticket instead of four dependent ones
  // (queue entry -> request head -> light term -> direction: 22 % of its warp time, profiles/r2a):
  float4* rq0;      // [parity][pos]: origin.xyz, w = bits nsamp(8) | level(5) << 8 | type(3) << 13 | area(1) << 16 | mat(8) << 24
  float4* rq1;      // [parity][pos]: single-light requests: dir.xyz, w = dist
  float4* rq2;      // [parity][pos]: li.xyz (single lights: the final contribution), w = path id
  float4* sp2;      // [parity][pos]: k_wf_shadow_prepare: E.xyz, flags (rq0 / rq1 then hold tmax, t / tdelta, P')
  float4* sp3;      // [parity][pos]: k_wf_shadow_prepare: distance to the light, request bits
  // area-light requests keep their samples by path (laid out only when the scene has area lights)
  uint32_t nsamp;   // samples allocated per request (>= area_samples)
  float4* sqd;      // [parity][sample][path]: dir.xyz, w = dist
  float4* sql;      // [parity][sample][path]: li.xyz, w = valid
  uint32_t* smask;      // [sample][path]: volumes that pass the bounding-sphere pre-test for the shadow ray (k_wf_shadow_mask)
  float4* ps[3];        // prepared walks [queue position]: 0 = tmax.xyz, t; 1 = tdelta.xyz, P' bits; 2 = E.xyz, flags (k_wf_prepare)
  uint32_t* q[2];       // active-path queues (ping-pong)
  uint32_t* counts;     // [iter][4]: 0 = active count entering iter, 1 = shadow requests of iter, 2 = extend cursor, 3 = shadow cursor
};

struct WfParams {
  DScene sc;
  VpCamera cam;
  WfState st;
  int W, H;
  int max_bounces;
  uint32_t frame;
  uint32_t seed;
  float aa;
  int row_begin, row_end, row_tile, row_tile_stride, row_tile_phase;
  int iter;              // bounce iteration of this launch
  int refill;            // lanes still walking below which idle lanes refill from the queue
  int refill_spare;      // ... in k_wf_extend_spare (a lane only idles after its spare ray, too)
  int bulk;              // with at most this many lanes walking, lanes in empty 8^3/16^3 blocks bulk-advance
  // compact path ids: only the 4-row tile rows a pass can own get ids, so a row band or one rank's share of the frame
  // sizes (and scans) a quarter / an eighth of the path state.  cmode 0: tile row = j; 1: j + ctrow0 (row range);
  // 2: interleaved row tiles of ck tile rows each: tile row = ((j / ck) * row_tile_stride + row_tile_phase) * ck + j % ck
  int cmode, ctrow0, ck;
  int prepared;          // 1: k_wf_prepare has set up the first walk of every ray of the bounce
  int ps_fmt;            // 1: ... in the record format of the single-volume kernels (k_wf_extend_spare, k_wf_shadow_spare)
  int light_smem, light_smem_off;  // shade stage: the light lists are staged in shared memory at this byte offset
  int write_rgb8;
  float4* accum;
  uint32_t* rgb8;
  unsigned long long* counters;  // same slots as RenderParams
  // reprojection frame mode (Renderer::Tick static-camera branch, renderer.cpp:1997-2101)
  int halo;               // 1: pass 1 also renders the rows next to the owned ones (row-sharded reprojection)
  int reproject;          // 1: TraceReproject variant of the shade stage, resolve splits albedo / illumination
  float* rp_alb;          // albedoBuffer        [W*H] float3
  float* rp_ill;          // illuminationBuffer  [W*H] float3
  float4* rp_ray;         // rayData             [W*H]: intersection point, w = material byte
  const float* hist_in;   // illuminationHistoryBuffer [W*H] float3
  float* hist_out;        // tempIlluminationBuffer    [W*H] float3
  VpPrevCamera prev;
};

__device__ __forceinline__ bool wf_row_owned(const WfParams& p, int y)
{
  if (y < p.row_begin || y >= p.row_end) return false;
  if (p.row_tile > 0 && ((y / p.row_tile) % p.row_tile_stride) != p.row_tile_phase) return false;
  return true;
}
// rows a pass works on: the owned rows, plus - in pass 1 of a row-sharded reprojection frame - their direct neighbours,
// whose illumination pass 2 reads (ClampHistory's 3x3 box)
__device__ __forceinline__ bool wf_row_selected(const WfParams& p, int y)
{
  if (wf_row_owned(p, y)) return true;
  return p.halo && ((y > 0 && wf_row_owned(p, y - 1)) || (y + 1 < p.H && wf_row_owned(p, y + 1)));
}
// path id <-> pixel: 8x4 tiles in row-major tile order over the tile rows the pass can own (WfParams.cmode), lane =
// position inside the tile
__device__ __forceinline__ bool wf_pixel_of(const WfParams& p, uint32_t path, int& x, int& y)
{
  const uint32_t tiles_x = (uint32_t)(p.W + 7) / 8u;
  const uint32_t tile = path >> 5, lane = path & 31u;
  uint32_t trow = tile / tiles_x;
  if (p.cmode == 1) trow += (uint32_t)p.ctrow0;
  else if (p.cmode == 2) { const uint32_t k = (uint32_t)p.ck; trow = ((trow / k) * (uint32_t)p.row_tile_stride + (uint32_t)p.row_tile_phase) * k + trow % k; }
  x = (int)((tile % tiles_x) * 8u + (lane & 7u));
  y = (int)(trow * 4u + (lane >> 3));
  return x < p.W && y < p.H && wf_row_selected(p, y);
}
__device__ __forceinline__ Staged wf_stage_scene(const DScene& sc, unsigned char* smem)
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
// The light lists of the scene, staged behind the volume table / region masks by the shade stage (WfParams.light_smem:
// 24 / 32 / 40 B per point / area / spot light; every lane of a warp may pick a different light, so these are the
// stage's scattered small reads).  Falls back to the global arrays when they do not fit.
struct Lights { const VpPointLight* pl; const VpAreaLight* al; const VpSpotLight* sl; };
__device__ __forceinline__ Lights wf_stage_lights(const DScene& sc, unsigned char* smem, int offset, int enabled)
{
  Lights L; L.pl = sc.pl; L.al = sc.al; L.sl = sc.sl;
  if (!enabled) return L;
  uint32_t* dst = reinterpret_cast<uint32_t*>(smem + offset);
  const int wp = sc.npl * (int)(sizeof(VpPointLight) / 4), wa = sc.nal * (int)(sizeof(VpAreaLight) / 4), ws = sc.nsl * (int)(sizeof(VpSpotLight) / 4);
  const uint32_t* sp = reinterpret_cast<const uint32_t*>(sc.pl);
  const uint32_t* sa = reinterpret_cast<const uint32_t*>(sc.al);
  const uint32_t* ss = reinterpret_cast<const uint32_t*>(sc.sl);
  for (int i = threadIdx.x; i < wp; i += blockDim.x) dst[i] = __ldg(sp + i);
  for (int i = threadIdx.x; i < wa; i += blockDim.x) dst[wp + i] = __ldg(sa + i);
  for (int i = threadIdx.x; i < ws; i += blockDim.x) dst[wp + wa + i] = __ldg(ss + i);
  __syncthreads();
  L.pl = reinterpret_cast<const VpPointLight*>(dst);
  L.al = reinterpret_cast<const VpAreaLight*>(dst + wp);
  L.sl = reinterpret_cast<const VpSpotLight*>(dst + wp + wa);
  return L;
}
// Blocks that get scheduled after the stage queue ran dry (small shards, late bounces) leave before staging the scene
// into shared memory; `taken` is the queue's ticket counter, n its length.
__device__ __forceinline__ bool wf_block_has_work(const uint32_t* taken, uint32_t n)
{
  __shared__ int go;
  if (threadIdx.x == 0) go = *reinterpret_cast<const volatile uint32_t*>(taken) < n;
  __syncthreads();
  return go != 0;
}
// warp-aggregated append of one id per participating lane
// returns the queue position of the lane's id (only meaningful where `want`)
__device__ __forceinline__ uint32_t wf_push(uint32_t* queue, uint32_t* counter, bool want, uint32_t id)
{
  const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
  if (!m) return 0;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
  if (want) queue[slot] = id;
  return slot;
}
__device__ __forceinline__ void wf_add_counter(unsigned long long* slot, unsigned long long v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(slot, v);
}
__device__ __forceinline__ uint32_t pack_r0w(int inside, int depth, int top) { return (uint32_t)(inside & 1) | ((uint32_t)((depth + 1) & 0xFF) << 8) | ((uint32_t)(top & 0xFF) << 16); }

// ---- K1: primary rays (Renderer::Update pixel loop head, renderer.cpp:1858-1866) --------------------
__global__ void __launch_bounds__(kWfBlock) k_wf_generate(const __grid_constant__ WfParams p)
{
  const uint32_t npix = (uint32_t)p.W * (uint32_t)p.H;
  unsigned long long launched = 0;
  for (uint32_t path = blockIdx.x * blockDim.x + threadIdx.x; path < p.st.cap; path += gridDim.x * blockDim.x) {
    int x, y;
    const bool on = wf_pixel_of(p, path, x, y);
    bool active = false;
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
    if (on) {
      const uint32_t px = (uint32_t)y * (uint32_t)p.W + (uint32_t)x;
      uint32_t seed = pixel_seed(px, p.frame, npix, p.seed);
      Ray pr;
      if (p.reproject) {
        pr = primary_ray_nodof(p.cam, (float)x, (float)y, p.W, p.H);  // renderer.cpp:2010: no jitter, no draws
      } else {
        const float rx = rnd(seed) * p.aa;
        const float ry = rnd(seed) * p.aa;
        pr = primary_ray(p.cam, (float)x + rx, (float)y + ry, p.W, p.H, seed);
      }
      a0 = make_float4(pr.O.x, pr.O.y, pr.O.z, __uint_as_float(pack_r0w(0, p.max_bounces, 0)));
      a1 = make_float4(pr.D.x, pr.D.y, pr.D.z, __uint_as_float(seed));
      p.st.term[path] = make_float4(0, 0, 0, 0);
      active = p.max_bounces >= 0;  // Trace(depth < 0) returns before FindNearest, renderer.cpp:1080
      if (active) launched++;
    }
    const uint32_t slot = wf_push(p.st.q[0], p.st.counts + 0, active, path);
    if (active) { p.st.r0[0][slot] = a0; p.st.r1[0][slot] = a1; }
  }
  wf_add_counter(p.counters + 4, launched);
}

// ---- K2a: prepare = the part of Renderer::FindNearest that precedes the walk ------------------------------------------
// (renderer.cpp:958-975: world -> object ray, Scene::Setup3DDDA) for the first volume each ray of the bounce enters, one
// lane per ray at full width - inside the extend stage the same ~240 instructions run with a quarter of the lanes, because
// only the lanes that refill execute them.  The record is indexed by queue position, so writes here and reads by the
// refilling lanes are coalesced.  flags: bits 0..5 = index + 1 of the volume entered (0: the ray enters none),
// bit 6 = ray excluded from brick mode, bits 7..9 = mirrored axes, bits 10..31 = which of the next 22 volumes pass the
// bounding-sphere pre-test (scenes of up to kMaskVolumes volumes: the extend stage then jumps from candidate to
// candidate instead of repeating the test on a few lanes).
__global__ void __launch_bounds__(kWfBlock) k_wf_prepare(const __grid_constant__ WfParams p)
{
  const uint32_t n = p.st.counts[4 * p.iter];
  const int nvol = p.sc.nvol;
  for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n; qi += gridDim.x * blockDim.x) {
    const float4 a = p.st.r0[p.iter & 1][qi], b = p.st.r1[p.iter & 1][qi];
    const f3 O = v3(a.x, a.y, a.z), D = v3(b.x, b.y, b.z);
    DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = s.t = 0; s.P = 0;
    BW bw; bw.ex = bw.ey = bw.ez = 0; bw.plain = 0; bw.flip = 0; bw.F = 0;
    uint32_t flags = 0;
    for (int vi = 0; vi < nvol; vi++) {
      const DVolume& v = p.sc.vol[vi];
      if (nvol > 1 && !volume_may_hit(v, O, D)) continue;
      const ORay r = make_oray(xform_pos_sse(O, v.iM), xform_vec_sse(D, v.iM));
      if (setup_dda(v, r, s)) {
        bw_init(v, s, bw);
        flags = (uint32_t)(vi + 1) | (bw.plain << 6) | ((s.Sx != 1u ? 1u : 0u) << 7) | ((s.Sy != 2u ? 1u : 0u) << 8) | ((s.Sz != 4u ? 1u : 0u) << 9);
        if (nvol > 1 && nvol <= kMaskVolumes)
          for (int vj = vi + 1; vj < nvol; vj++)
            if (volume_may_hit(p.sc.vol[vj], O, D)) flags |= 1u << (10 + vj - vi - 1);
        break;
      }
    }
    if (p.ps_fmt) {
      // single-volume kernels (k_wf_extend_spare, k_wf_shadow_spare): the lane that takes the record does no decoding - the mirror
      // word itself, and a ray that enters no volume starts beyond every stopping distance
      flags = (flags & 63u) ? ((bw.flip << 2) | (bw.plain << 1) | 1u) : 0u;
      if (!flags) s.t = 3.0e38f;
    }
    p.st.ps[0][qi] = make_float4(s.tx, s.ty, s.tz, s.t);
    p.st.ps[1][qi] = make_float4(s.dx, s.dy, s.dz, __uint_as_float(s.P));
    p.st.ps[2][qi] = make_float4(bw.ex, bw.ey, bw.ez, __uint_as_float(flags));
  }
}

// ---- K2: extend = Renderer::FindNearest (renderer.cpp:946-1018), persistent warps with lane refill ----
// Writes per path: t, the winning volume and the hit cell (as P).  The face normal and the
// material byte are produced by the shade stage where all 32 lanes are busy, not here where
// a hit is a single-lane event inside the stepping loop.
template <bool CS, bool BWK, int MINB = kWfMinBlocks>
__global__ void __launch_bounds__(kWfBlock, MINB) k_wf_extend(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  if (!wf_block_has_work(p.st.counts + 4 * p.iter + 2, p.st.counts[4 * p.iter])) return;
  const Staged stg = wf_stage_scene(p.sc, smem);
  const DVolume* svol = stg.vol;
  const int lane = threadIdx.x & 31;
  uint32_t* cnt = p.st.counts + 4 * p.iter;
  const uint32_t n = cnt[0];
  const int nvol = p.sc.nvol;
  const bool need_world = nvol > 1 || p.sc.nsph + p.sc.ntri > 0 || !p.prepared;  // the world-space ray is only needed past the first volume
  const bool use_mask = BWK && p.prepared && nvol > 1 && nvol <= kMaskVolumes;
  uint32_t cand = 0;
  bool has_ray = false, in_dda = false, exhausted = false;
  uint32_t pos = 0;  // the ray's position in the bounce's queue: where its ray is read and its hit is written
  f3 O = v3s(0), D = v3s(0);
  float tbest = 0; uint32_t pbest = 0; int volbest = -2;
  int vi = 0;
  const unsigned long long* region = nullptr;
  const unsigned long long* fine = nullptr;
  DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = s.t = 0; s.P = s.R = 0; s.Ax = s.Ay = s.Az = s.Sx = s.Sy = s.Sz = 0;
  Occ o; occ_reset(o);
  BW bw; bw.ex = bw.ey = bw.ez = bw.tc = 0; bw.F = bw.flip = bw.ac = bw.plain = 0;
  unsigned long long steps = 0, steps_er = 0, steps_eb = 0;
  unsigned int nrays = 0;
  for (;;) {
    // 1. refill idle lanes from the queue
    const bool need = !has_ray && !exhausted;
    const unsigned nm = __ballot_sync(0xFFFFFFFFu, need);
    if (nm) {
      const int leader = __ffs(nm) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(cnt + 2, (uint32_t)__popc(nm));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (need) {
        const uint32_t qi = base + (uint32_t)__popc(nm & ((1u << lane) - 1u));
        if (qi < n) {
          pos = qi;
          tbest = 1e34f; pbest = 0; volbest = -2;
          vi = 0; has_ray = true; in_dda = false;
          nrays++;
          if (BWK && p.prepared) {
            // the walk through the first volume the ray enters was set up by k_wf_prepare.  The three records are taken
            // unconditionally (a ray that enters no volume has zeros there): loads that only feed the branch below would
            // be issued inside it, one memory round trip after the flags word arrived
            const float4 w0 = __ldcs(p.st.ps[0] + qi), w1 = __ldcs(p.st.ps[1] + qi), w2 = __ldcs(p.st.ps[2] + qi);
            const uint32_t fl = __float_as_uint(w2.w);
            vi = (fl & 63u) ? (int)(fl & 63u) : nvol;
            cand = fl >> 10;
#if VP_HOIST_PS
            s.tx = w0.x; s.ty = w0.y; s.tz = w0.z; s.t = w0.w;
            s.dx = w1.x; s.dy = w1.y; s.dz = w1.z; s.P = __float_as_uint(w1.w);
            bw.ex = w2.x; bw.ey = w2.y; bw.ez = w2.z;
#endif
            if (fl & 63u) {
              const DVolume& v = svol[vi - 1];
#if !VP_HOIST_PS
              s.tx = w0.x; s.ty = w0.y; s.tz = w0.z; s.t = w0.w;
              s.dx = w1.x; s.dy = w1.y; s.dz = w1.z; s.P = __float_as_uint(w1.w);
              bw.ex = w2.x; bw.ey = w2.y; bw.ez = w2.z;
#endif
              bw.F = axis_field(v.jr, 0);
              bw.flip = ((fl >> 7) & 1u ? bw.F : 0u) | ((fl >> 8) & 1u ? bw.F << 1 : 0u) | ((fl >> 9) & 1u ? bw.F << 2 : 0u);
              bw.plain = (fl >> 6) & 1u; bw.ac = kAcFirst; bw.tc = __uint_as_float(kTcFirst);
              occ_reset(o); region = region_of(stg, v); fine = v.fine;
              in_dda = true;
            }
          }
          if (!(BWK && p.prepared) || need_world) {
            const float4 a = p.st.r0[p.iter & 1][qi], b = p.st.r1[p.iter & 1][qi];
            O = v3(a.x, a.y, a.z); D = v3(b.x, b.y, b.z);
          }
        } else exhausted = true;
      }
    }
    if (!__any_sync(0xFFFFFFFFu, has_ray)) break;
    // 2. lanes between volumes: set up the next volume the ray enters, or finish the ray
    if (has_ray && !in_dda) {
      while (vi < nvol) {
        if (use_mask) {
          // bit 0 of cand <=> volume vi passes the bounding-sphere pre-test (k_wf_prepare): jump to the next candidate
          if (cand == 0u) { vi = nvol; break; }
          const int skip = __ffs(cand) - 1;
          vi += skip;
          cand = (cand >> skip) >> 1;
        } else if (nvol > 1 && !volume_may_hit(svol[vi], O, D)) { vi++; continue; }
        const DVolume& v = svol[vi];
        const ORay r = make_oray(xform_pos_sse(O, v.iM), xform_vec_sse(D, v.iM));
        vi++;
        if (setup_dda(v, r, s)) {
          in_dda = true; occ_reset(o); region = region_of(stg, v); fine = v.fine;
          if (BWK) bw_init(v, s, bw);
          break;
        }
      }
      if (!in_dda) {
        uint32_t hb = (uint32_t)(volbest + 2) << 8;
        if (p.sc.nsph + p.sc.ntri > 0) {
          PrimHit h; h.t = 1e34f; h.N = v3s(0); h.mat = VP_MAT_NONE; h.inside = 0;
          prims_nearest(p.sc, O, normalize3(D), h);
          if (tbest > h.t) {
            tbest = h.t;
            hb = (uint32_t)(h.mat & 255) | (1u << 8) | ((uint32_t)(h.inside & 1) << 30) | (1u << 31);  // vol = -1, bit 31: inside flag valid
            p.st.hn[pos] = make_float4(h.N.x, h.N.y, h.N.z, 0);
          }
        }
        p.st.h0[pos] = make_float2(tbest, __uint_as_float(pbest));
        p.st.h1[pos] = hb;
        has_ray = false;
      }
    }
    // 3. DDA stepping (Scene::FindNearest loop, scene.cpp:763-804) until too few lanes are walking.
    //    A lane walks while s.t < tbest: a hit sets tbest = s.t, leaving the grid sets s.t = inf.
    if (!in_dda) s.t = 3.0e38f;
    const unsigned thr = __any_sync(0xFFFFFFFFu, exhausted) ? 1u : (unsigned)p.refill;  // queue drained: run to completion
    if (BWK) {
      // one step body for lanes inside empty bricks (a whole brick per step) and lanes inside non-empty ones (the
      // reference's cell step with the occupancy bit test): vp_device.cuh "brick walk"
      for (;;) {
#pragma unroll
        for (int u = 0; u < kBwUnroll; u++) {
          if (s.t < tbest) {
            const float t0 = s.t;
            if (walk_step(region, fine, s, bw, o, __uint_as_float(kHitTag))) tbest = t0;  // NaN tag: stops the walk, recovered below
          }
        }
        if ((unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tbest)) < thr) break;
      }
    } else {
      const bool drained = thr == 1u;  // no more work to pull: latency of the remaining rays is all that matters
      bool sparse = drained || (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tbest)) <= (unsigned)p.bulk;
      for (;;) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (s.t < tbest) {
            if (CS) steps++;
            uint32_t skipped = 0;
            const bool solid = occ_visit(region, fine, o, s, tbest, sparse, skipped);
            if (CS) steps += skipped;
            if (CS) { if (o.rm == 0ull) steps_er += 1 + skipped; if (o.m == 0ull) steps_eb += 1 + skipped; }
            if (solid) { tbest = s.t; s.t = __uint_as_float(kHitTag); }  // NaN tag: stops the walk, recovered below
            else if (!dda_step(s)) s.t = 3.0e38f;
          }
        }
        const unsigned act = (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tbest));
        sparse = drained || act <= (unsigned)p.bulk;
        if (act < thr) break;
      }
    }
    if (__float_as_uint(s.t) == kHitTag) { pbest = BWK ? (s.P ^ bw.flip) : s.P; volbest = vi - 1; s.t = 3.0e38f; }  // the cell the walk stopped on
    in_dda = s.t < tbest;
  }
  wf_add_counter(p.counters + 0, nrays);
  if (CS) { wf_add_counter(p.counters + 3, steps); wf_add_counter(p.counters + 9, steps_er); wf_add_counter(p.counters + 10, steps_eb); wf_add_counter(p.counters + 11, steps); }
}

// ---- K2 for single-volume scenes: a spare ray per lane ---------------------------------------------------------------------
// In k_wf_extend a lane whose walk has ended idles until fewer than `refill` lanes of its warp walk: the stepping loop, 84 %
// of the stage's instructions, runs at 19 of 32 lanes (profiles/r2).  Refilling earlier does not pay there, because a refill
// is a ticket atomic plus loads from L2.  Here every lane holds a SPARE: the ticket of its next queue entry, taken together
// with the warp's other tickets in the same aggregated atomic as today, its records touched (prefetch to L1) when the
// ticket arrives.  A lane whose walk ends writes its hit and switches to its spare at the next vote of the stepping loop -
// the lanes that ended in the same round do it together, from L1, without an atomic - and only idles once the spare is
// used up too.  Single-volume scenes without primitives only (there the end of a walk is the end of the ray); towards the
// end of the queue no spares are handed out (no lane sits on an entry while others have run dry).  Same arithmetic,
// records and results as k_wf_extend<false, true>.
constexpr uint32_t kNoTicket = 0xFFFFFFFFu;
template <int MINB>
__global__ void __launch_bounds__(kWfBlock, MINB) k_wf_extend_spare(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  if (!wf_block_has_work(p.st.counts + 4 * p.iter + 2, p.st.counts[4 * p.iter])) return;
  const Staged stg = wf_stage_scene(p.sc, smem);
  const DVolume& v = stg.vol[0];
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t* cnt = p.st.counts + 4 * p.iter;
  const uint32_t n = cnt[0];
  const unsigned long long* const region = region_of(stg, v);
  const unsigned long long* const fine = v.fine;
  const uint32_t F = axis_field(v.jr, 0);
  const uint32_t spare_floor = (uint32_t)gridDim.x * (uint32_t)kWfBlock;  // entries left below which no spares are handed out
  bool has_ray = false, exhausted = false, nomore = false;
  uint32_t pos = 0, spare = kNoTicket;
  float tbest = 0;
  DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = 0; s.t = 3.0e38f; s.P = s.R = 0; s.Ax = s.Ay = s.Az = s.Sx = s.Sy = s.Sz = 0;
  Occ o; occ_reset(o);
  BW bw; bw.ex = bw.ey = bw.ez = bw.tc = 0; bw.F = F; bw.flip = bw.ac = bw.plain = 0;
  unsigned int nrays = 0;
  // the prepared walk of queue entry tk becomes the lane's ray (a ray that enters no volume ends at the next vote)
#define VP_LOAD_RAY(tk_)                                                                                              \
  {                                                                                                                   \
    const float4 w0_ = p.st.ps[0][tk_], w1_ = p.st.ps[1][tk_], w2_ = p.st.ps[2][tk_];                                 \
    const uint32_t fl_ = __float_as_uint(w2_.w);                                                                      \
    pos = (tk_); tbest = 1e34f; has_ray = true; nrays++;                                                              \
    s.tx = w0_.x; s.ty = w0_.y; s.tz = w0_.z; s.t = w0_.w;                                                            \
    s.dx = w1_.x; s.dy = w1_.y; s.dz = w1_.z; s.P = __float_as_uint(w1_.w);                                           \
    bw.ex = w2_.x; bw.ey = w2_.y; bw.ez = w2_.z;                                                                      \
    bw.flip = fl_ >> 2; bw.plain = (fl_ >> 1) & 1u; bw.ac = kAcFirst; bw.tc = __uint_as_float(kTcFirst);                 \
    occ_reset(o);                                                                                                     \
  }
  for (;;) {
    // 1. a lane without a ray takes a queue entry and a spare, a lane that walks but has used its spare takes a new spare:
    //    one aggregated atomic for all tickets of the warp
    if (!has_ray && spare != kNoTicket) { VP_LOAD_RAY(spare); spare = kNoTicket; }
    {
      const bool ask_ray = !has_ray && !exhausted;
      const bool ask_spare = !exhausted && !nomore && spare == kNoTicket;
      const unsigned m1 = __ballot_sync(0xFFFFFFFFu, ask_ray), m2 = __ballot_sync(0xFFFFFFFFu, ask_spare);
      if (m1 | m2) {
        const int leader = __ffs(m1 | m2) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(cnt + 2, (uint32_t)(__popc(m1) + __popc(m2)));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        uint32_t tk = base + (uint32_t)(__popc(m1 & lt) + __popc(m2 & lt));
        if (ask_ray) {
          if (tk < n) { VP_LOAD_RAY(tk); } else { exhausted = true; }
          tk++;
        }
        if (ask_spare) {
          // keep the spare only while the queue is long: at its end a held entry is a ray nobody walks
          if (tk < n && n - tk > spare_floor) {
            spare = tk;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.st.ps[0] + tk));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.st.ps[1] + tk));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.st.ps[2] + tk));
          } else if (tk < n) {
            spare = tk;  // too close to the end to ask for more, but the entry is ours
            nomore = true;
          } else nomore = true;
        }
      }
    }
    if (!__any_sync(0xFFFFFFFFu, has_ray)) break;
    // 2. the brick walk; at every vote the lanes whose walk has ended write their hit and go on with their spare
    const unsigned thr = __any_sync(0xFFFFFFFFu, exhausted | nomore) ? 1u : (unsigned)p.refill_spare;
    for (;;) {
#pragma unroll
      for (int u = 0; u < kBwUnroll; u++) {
        if (s.t < tbest) {
          const float t0 = s.t;
          if (walk_step(region, fine, s, bw, o, __uint_as_float(kHitTag))) tbest = t0;
        }
      }
      if (has_ray && !(s.t < tbest)) {
        const bool hit = __float_as_uint(s.t) == kHitTag;
        p.st.h0[pos] = make_float2(tbest, __uint_as_float(hit ? (s.P ^ bw.flip) : 0u));
        p.st.h1[pos] = (uint32_t)(hit ? 2 : 0) << 8;  // volume 0, or -2: no hit
        has_ray = false;
        s.t = 3.0e38f;
        if (spare != kNoTicket) { VP_LOAD_RAY(spare); spare = kNoTicket; }
      }
      if ((unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tbest)) < thr) break;
    }
  }
#undef VP_LOAD_RAY
  wf_add_counter(p.counters + 0, nrays);
}

// ---- K3: shade = the material switch of Renderer::Trace (renderer.cpp:1092-1327) -------------------
// One lane per path of the iteration's queue.  Emits the next ray, the fold level and, where
// Illumination() would run, a shadow request (the RNG draws of the light sample happen here, in
// the reference's order; only the occlusion tests are deferred to K4).
// Appends the lanes that call it with want == true (a subset of the lanes executing this code together) to the shadow-request
// queue of the bounce: one aggregated atomicAdd per converged group; returns the lane's position.
__device__ __forceinline__ uint32_t wf_request_slot(uint32_t* counter)
{
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
template <bool CS>
__device__ __forceinline__ void wf_light_request(const WfParams& p, const Lights& lights, uint32_t path, int level, int type, const Ray& ray,
                                                 uint32_t& seed, bool& queued, f3& A_now)
{
  const DScene& sc = p.sc;
  const uint32_t lightCount = (uint32_t)(sc.npl + sc.nsl + sc.nal + 1);
  const float pick = rnd(seed) * (float)lightCount;
  const uint32_t idx = (uint32_t)pick;
  const f3 ip = isect(ray);
  const f3 n = ray.N;
  const f3 k = albedo_of(sc, ray.mat);
  queued = false;
  A_now = muls(v3s(0), (float)lightCount);
  const uint32_t cap = p.st.cap;
  const size_t par = (size_t)(level & 1);  // in the staged pipeline the level of a request is the bounce index
  float4* const rq0 = p.st.rq0 + par * cap;
  float4* const rq1 = p.st.rq1 + par * cap;
  float4* const rq2 = p.st.rq2 + par * cap;
  uint32_t* const counter = p.st.counts + 4 * level + 1;
  const uint32_t tag = ((uint32_t)level << 8) | ((uint32_t)type << 13) | ((uint32_t)(ray.mat & 255) << 24);
  if (idx >= (uint32_t)sc.npl && idx < (uint32_t)(sc.nal + sc.npl)) {
    float4* const sqd = p.st.sqd + par * p.st.nsamp * cap;
    float4* const sql = p.st.sql + par * p.st.nsamp * cap;
    const VpAreaLight L = lights.al[idx - sc.npl];
    const f3 center = ld3(L.position);
    const float radius = L.radius;
    const f3 point = offset_ray(ip, n);
    int nvalid = 0;
    for (int i = 0; i < sc.area_samples; i++) {
      f3 rp = random_direction(seed);
      rp = muls(rp, radius);
      rp = add(rp, center);
      const f3 dir = sub(rp, ip);
      const float dst = sqrtf(dot3(dir, dir));
      const f3 dn = muls(dir, 1 / dst);
      const float ct = dot3(dn, n);
      f3 li = muls(ld3(L.color), ct);
      li = muls(li, L.color_multiplier);
      li = muls(li, radius * radius);
      li = muls(li, VP_PI);
      li = muls(li, 4.0f);
      li = divs(li, dst * dst);
      const bool valid = !(ct <= 0);
      nvalid += valid;
      sqd[(size_t)i * cap + path] = make_float4(dn.x, dn.y, dn.z, dst);
      sql[(size_t)i * cap + path] = make_float4(li.x, li.y, li.z, valid ? 1.0f : 0.0f);
    }
    if (nvalid > 0) {
      const uint32_t pos = wf_request_slot(counter);
      rq0[pos] = make_float4(point.x, point.y, point.z, __uint_as_float((uint32_t)sc.area_samples | (1u << 16) | tag));
      rq2[pos] = make_float4(0, 0, 0, __uint_as_float(path));
      queued = true;
    } else {
      // inc = (0 / n) * k * lightCount, as the reference computes it
      f3 inc = divs(v3s(0), (float)sc.area_samples);
      inc = mul(inc, k);
      A_now = muls(inc, (float)lightCount);
    }
  } else {
    // The light term is only evaluated where it is used (the sample faces the light / lies inside the cone): the reference
    // computes it before its early return, unobserved - and here an unlit spot sample is 0 / dst^2 three times, which the
    // IEEE division takes through its out-of-line path (7 % of the stage's instructions, profiles/r2d).
    f3 dirn, contrib = v3s(0); float tshadow; bool lit;
    if (idx < (uint32_t)sc.npl) {
      const VpPointLight L = lights.pl[idx];
      const f3 dir = sub(ld3(L.position), ip);
      const float dst = sqrtf(dot3(dir, dir));
      dirn = muls(dir, 1.0f / dst);
      const float ct = dot3(dirn, n);
      lit = !(ct <= 0.0f);
      if (lit) contrib = mul(muls(muls(ld3(L.color), stdmax(0.0f, ct)), 1.0f / (dst * dst)), k);
      tshadow = dst;
    } else if (idx < (uint32_t)(sc.nal + sc.nsl + sc.npl)) {
      const VpSpotLight L = lights.sl[idx - sc.nal - sc.npl];
      const f3 dir = sub(ld3(L.position), ip);
      const float dst = sqrtf(dot3(dir, dir));
      dirn = divs(dir, dst);
      const float ct = dot3(dirn, ld3(L.direction));
      lit = !(ct <= L.angle);
      if (lit) {
        const float alpha = 1.0f - (1.0f - ct) * 1.0f / (1.0f - L.angle);
        const f3 li = divs(muls(ld3(L.color), stdmax(0.0f, ct)), dst * dst);
        contrib = muls(mul(li, k), alpha);
      }
      tshadow = dst;
    } else {
      dirn = neg(ld3(sc.dl.direction));
      const float ct = dot3(dirn, n);
      lit = !(ct <= 0);
      if (lit) contrib = mul(muls(ld3(sc.dl.color), stdmax(0.0f, ct)), k);
      tshadow = 1e34f;
    }
    if (lit) {
      const f3 org = offset_ray(ip, n);
      const uint32_t pos = wf_request_slot(counter);
      rq0[pos] = make_float4(org.x, org.y, org.z, __uint_as_float(1u | tag));
      rq1[pos] = make_float4(dirn.x, dirn.y, dirn.z, tshadow);
      rq2[pos] = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(path));
      queued = true;
    }
  }
}

// Shade one path (the body of Trace's switch for the hit the extend stage recorded).  Outputs:
// cont = the path continues with the ray now stored in r0/r1; want_shadow = a light sample was
// written to the request slot [level][path] and needs its occlusion test.
template <bool CS>
__device__ __forceinline__ void wf_shade_one(const WfParams& p, const Staged& stg, const Lights& lights, uint32_t qi, uint32_t& path, Counters& cn,
                                             int& in_light, bool& cont, bool& want_shadow, int& level, float4& next0, float4& next1)
{
  const DVolume* svol = stg.vol;
  const DScene& sc = p.sc;
  const uint32_t cap = p.st.cap;
  cont = false; want_shadow = false; level = 0;
  // everything the bounce knows about the path sits at its queue position: one trip to memory
  const float4 a = __ldcs(p.st.r0[p.iter & 1] + qi), b = __ldcs(p.st.r1[p.iter & 1] + qi);
  const float2 h = __ldcs(p.st.h0 + qi);
  const uint32_t hb = __ldcs(p.st.h1 + qi), rb = __float_as_uint(a.w);
  path = p.st.q[p.iter & 1][qi];
  Ray ray;
  ray.O = v3(a.x, a.y, a.z); ray.D = v3(b.x, b.y, b.z);
  ray.t = h.x; ray.N = v3s(0); ray.mat = VP_MAT_NONE;
  ray.inside = (int)(rb & 1u);
  const int voxIndex = (int)((hb >> 8) & 0x3FFFFFu) - 2;
  if (voxIndex >= 0) {
    // finish Scene::FindNearest's hit block here (scene.cpp:766-773): material byte + Ray::GetNormalVoxel
    const DVolume& v = svol[voxIndex];
    const f3 oO = xform_pos_sse(ray.O, v.iM), oD = xform_vec_sse(ray.D, v.iM);
    ray.mat = __ldg(v.grid + cell_of(v, __float_as_uint(h.y)));
    ray.N = normal_voxel(v, oO, oD, neg_bits(oD), ray.t, nullptr);
  } else if (voxIndex == -1) {
    const float4 nn = __ldcs(p.st.hn + qi);
    ray.N = v3(nn.x, nn.y, nn.z);
    ray.mat = (int)(hb & 255u);
    ray.inside = (int)((hb >> 30) & 1u);  // FindNearest overwrites isInsideGlass on a sphere/triangle hit (renderer.cpp:1012)
  }
  uint32_t seed = __float_as_uint(b.w);
  int depth = (int)((rb >> 8) & 0xFFu) - 1;
  int top = (int)((rb >> 16) & 0xFFu);
  const int m = ray.mat;
  if (p.reproject && p.iter == 0) {  // RayDataReproject::GetRayInfo(primaryRay), renderer.h:31-34 (a primary ray is never inside a medium)
    int x, y;
    wf_pixel_of(p, path, x, y);
    const f3 P = isect(ray);
    __stcg(p.rp_ray + ((size_t)y * p.W + x), make_float4(P.x, P.y, P.z, __uint_as_float((uint32_t)m & 255u)));
  }
  if (m == VP_MAT_NONE) {
    cn.sky++;
    const f3 t = sample_sky(sc, ray.D);
    __stcg(p.st.term + path, make_float4(t.x, t.y, t.z, __uint_as_float((uint32_t)top)));
  } else if (m == 15) {
    const f3 t = muls(albedo_of(sc, m), sc.mat[m].emissive);
    __stcg(p.st.term + path, make_float4(t.x, t.y, t.z, __uint_as_float((uint32_t)top)));
  } else {
    // Every material class ends in "spawn the next ray from the hit along rdir" (normalising Ray ctor) and a fold level;
    // the classes only differ in how rdir, the side of the surface (rn), the medium flag and the level are made.  The
    // parts they share are executed once, with the lanes of all classes together.
    int type = 0;
    f3 M = v3s(1.0f);  // multiplier of a type-0 level
    f3 rn = ray.N, rdir = v3s(0);
    int ins = 0;
    bool illum = false;
    if (m <= 7) {
      // METAL_* (renderer.cpp:1103-1114) and NON_METAL_* (:1117-1144; TraceNonMetal :1342-1358 in the reprojection mode,
      // which draws no Fresnel number and always takes the Lambert branch): ONE sphere sample for the lanes of these
      // classes, after the Fresnel draw of the non-metals - the draw order of each class is unchanged
      bool mirror = m >= 5;
      if (m <= 4) {
        if (p.reproject) type = 4;
        else { mirror = !(rnd(seed) > schlick_nonmetal(dot3(neg(ray.D), ray.N))); type = mirror ? 3 : 1; }
      }
      const f3 rs = random_sphere_sample(seed);
      if (mirror) rdir = add(reflect3(ray.D, ray.N), muls(rs, sc.mat[m].roughness));
      else { rdir = add(ray.N, rs); illum = true; }
      if (type == 0) M = albedo_of(sc, m);
    } else if (m == 8) {
      f3 color = v3s(1.0f);
      int inGlass = ray.inside;
      const float ior = sc.mat[m].ior;
      const float ratio = inGlass ? ior : 1.0f / ior;
      bool insideVolume = true;
      if (inGlass) {
        color = albedo_of(sc, m);
        if (voxIndex >= 0) insideVolume = medium_exit<CS>(svol[voxIndex], ray, 0, cn);
      }
      if (!insideVolume) { ray.O = add(ray.O, muls(ray.D, ray.t)); ray.t = 0; }
      const float ct = stdmin(dot3(neg(ray.D), ray.N), 1.0f);
      const float st = sqrtf(1.0f - ct * ct);
      const bool cannot = ratio * st > 1.0f;
      if (cannot || schlick(ct, ratio) > rnd(seed)) { rdir = reflect3(ray.D, ray.N); rn = ray.N; }
      else { rdir = refract3(ray.D, ray.N, ratio); inGlass = !inGlass; rn = neg(ray.N); }
      ins = inGlass;
      M = color;
    } else if (m >= 9 && m <= 14) {
      f3 color = v3s(1.0f);
      int inGlass = ray.inside;
      bool insideVolume = true;
      float intensity = 0, dist = 0;
      if (voxIndex == 0) {  // player-volume side channel: evaluated in place (rare), result only feeds inLight
        const f3 inc = illumination<CS>(sc, stg, ray, seed, cn);
        if (dot3(inc, inc) > 16.0f) in_light = 1;
      }
      if (inGlass) {
        intensity = sc.mat[m].emissive;
        color = albedo_of(sc, m);
        if (voxIndex >= 0) insideVolume = medium_exit<CS>(svol[voxIndex], ray, 1, cn);
        dist = ray.t;
      }
      const float threshold = rnd(seed) * 100 - intensity;
      if (rnd(seed) * dist > threshold) {
        const float lo = ray.t * .45f;
        const float rr = lo + rnd(seed) * (ray.t - lo);
        ray.O = add(ray.O, muls(ray.D, rr));
        ray.D = random_direction(seed);
        ray.t = 0;
      }
      color = absorption(color, intensity, dist);
      if (!insideVolume) { ray.O = add(ray.O, muls(ray.D, ray.t)); ray.t = 0; }
      rdir = refract3(ray.D, ray.N, 1.0f);
      rn = neg(ray.N);
      ins = !inGlass;
      M = color;
    } else {
      rdir = diffuse_reflection(seed, ray.N);
      illum = true;
      type = 2;
    }
    f3 A = v3s(0);
    if (illum) { wf_light_request<CS>(p, lights, path, top, type, ray, seed, want_shadow, A); level = top; }
    // one record per level: the multiplier itself for type 0, else the light term (the shadow stage overwrites it when the
    // sample was queued) - M is then the material's albedo, which the resolve stage looks up again
    const f3 rec = type == 0 ? M : A;
    __stcg(p.st.lv + (size_t)top * cap + path, make_float4(rec.x, rec.y, rec.z, __uint_as_float((uint32_t)type | ((uint32_t)(m & 255) << 8))));
    top++; depth--;
    cont = depth >= 0 && top < VP_MAX_LEVELS;  // Trace(depth < 0) returns 0: term stays 0
    if (cont) {
      const Ray nr = spawn(offset_ray(isect(ray), rn), rdir, ins);
      next0 = make_float4(nr.O.x, nr.O.y, nr.O.z, __uint_as_float(pack_r0w(ins, depth, top)));  // stored where the append puts the path
      next1 = make_float4(nr.D.x, nr.D.y, nr.D.z, __uint_as_float(seed));
    } else {
      __stcg(p.st.term + path, make_float4(0, 0, 0, __uint_as_float((uint32_t)top)));  // the path ends here: Trace(depth < 0) returns 0
    }
  }

}

template <bool CS>
__global__ void __launch_bounds__(kWfBlock, kWfMinBlocksShade) k_wf_shade(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  if (blockIdx.x * blockDim.x >= p.st.counts[4 * p.iter]) return;  // grid-stride loop below would be empty
  const Staged stg = wf_stage_scene(p.sc, smem);
  const Lights lights = wf_stage_lights(p.sc, smem, p.light_smem_off, p.light_smem);
  uint32_t* qnext = p.st.q[(p.iter + 1) & 1];
  uint32_t* cnt = p.st.counts + 4 * p.iter;
  const uint32_t n = cnt[0];
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  int in_light = 0;
  const uint32_t nround = (n + 31u) & ~31u;  // whole warps so the aggregated pushes stay convergent
  for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < nround; qi += gridDim.x * blockDim.x) {
    bool cont = false, want_shadow = false;
    uint32_t path = 0;
    float4 next0 = make_float4(0, 0, 0, 0), next1 = next0;
    if (qi < n) {
      int level;
      wf_shade_one<CS>(p, stg, lights, qi, path, cn, in_light, cont, want_shadow, level, next0, next1);
    }
    const uint32_t slot = wf_push(qnext, cnt + 4, cont, path);          // counts[iter+1][0]
    if (cont) { __stcg(p.st.r0[(p.iter + 1) & 1] + slot, next0); __stcg(p.st.r1[(p.iter + 1) & 1] + slot, next1); }
  }
  wf_add_counter(p.counters + 6, cn.sky);
  wf_add_counter(p.counters + 2, cn.exitray);
  wf_add_counter(p.counters + 1, cn.shadow);
  wf_add_counter(p.counters + 5, (unsigned long long)in_light);
  if (CS) wf_add_counter(p.counters + 3, cn.steps);
}

// ---- K4a: bounding-sphere pre-test of every shadow ray of the bounce against every volume, one lane per request at full
// width (multi-volume scenes of up to 32 volumes): the shadow stage then jumps from candidate to candidate instead of
// repeating the test on the few lanes that are between volumes.
__global__ void __launch_bounds__(kWfBlock) k_wf_shadow_mask(const __grid_constant__ WfParams p)
{
  const uint32_t n = p.st.counts[4 * p.iter + 1];
  const uint32_t cap = p.st.cap;
  const int nvol = p.sc.nvol;
  const size_t par = (size_t)(p.iter & 1);
  const float4* const rq0 = p.st.rq0 + par * cap;
  const float4* const rq1 = p.st.rq1 + par * cap;
  const float4* const rq2 = p.st.rq2 + par * cap;
  const float4* const sqd = p.st.sqd + par * p.st.nsamp * cap;
  const float4* const sql = p.st.sql + par * p.st.nsamp * cap;
  for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n; qi += gridDim.x * blockDim.x) {
    const float4 q0 = rq0[qi];
    const f3 org = v3(q0.x, q0.y, q0.z);
    const uint32_t bits = __float_as_uint(q0.w);
    const int nsamp = (int)(bits & 255u);
    const bool area = (bits >> 16) & 1u;
    const uint32_t path = __float_as_uint(rq2[qi].w);
    for (int sample = 0; sample < nsamp; sample++) {
      if (area && sql[(size_t)sample * cap + path].w == 0.0f) continue;
      const float4 d4 = area ? sqd[(size_t)sample * cap + path] : rq1[qi];
      const f3 D = normalize3(v3(d4.x, d4.y, d4.z));  // as the shadow stage builds the ray
      uint32_t m = 0;
      for (int vi = 0; vi < nvol; vi++)
        if (volume_may_hit(p.sc.vol[vi], org, D)) m |= 1u << vi;
      p.st.smask[(size_t)sample * cap + qi] = m;  // by queue position
    }
  }
}

// ---- K4: shadow = Renderer::IsOccluded per light sample (renderer.cpp:209-243), lane refill -------
template <bool CS, bool BWK, int MINB = kWfMinBlocksShadow>
__global__ void __launch_bounds__(kWfBlock, MINB) k_wf_shadow(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  if (!wf_block_has_work(p.st.counts + 4 * p.iter + 3, p.st.counts[4 * p.iter + 1])) return;
  const Staged stg = wf_stage_scene(p.sc, smem);
  const DVolume* svol = stg.vol;
  const int lane = threadIdx.x & 31;
  const DScene& sc = p.sc;
  uint32_t* cnt = p.st.counts + 4 * p.iter;
  const uint32_t n = cnt[1];
  const uint32_t cap = p.st.cap;
  const int nvol = sc.nvol;
  const float lightCount = (float)(uint32_t)(sc.npl + sc.nsl + sc.nal + 1);
  const int par = p.iter & 1;
  const float4* const rq0 = p.st.rq0 + (size_t)par * cap;  // the requests of this bounce, by queue position
  const float4* const rq1 = p.st.rq1 + (size_t)par * cap;
  const float4* const rq2 = p.st.rq2 + (size_t)par * cap;
  const float4* const sqd = p.st.sqd + (size_t)par * p.st.nsamp * cap;  // area-light samples, by path
  const float4* const sql = p.st.sql + (size_t)par * p.st.nsamp * cap;
  const bool use_mask = BWK && p.prepared && nvol > 1 && nvol <= 32;  // k_wf_shadow_mask ran for this bounce
  uint32_t cand = 0;
  bool has_req = false, has_ray = false, in_dda = false, exhausted = false;
  uint32_t path = 0, bits = 0, pos = 0;
  int nsamp = 0, sample = 0, vi = 0;
  f3 org = v3s(0), D = v3s(0), inc = v3s(0), li = v3s(0);
  float tmaxray = 0;
  const unsigned long long* region = nullptr;
  const unsigned long long* fine = nullptr;
  DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = s.t = 0; s.P = s.R = 0; s.Ax = s.Ay = s.Az = s.Sx = s.Sy = s.Sz = 0;
  Occ o; occ_reset(o);
  BW bw; bw.ex = bw.ey = bw.ez = bw.tc = 0; bw.F = bw.flip = bw.ac = bw.plain = 0;
  unsigned long long steps = 0;
  unsigned int nrays = 0;
  for (;;) {
    const bool need = !has_req && !exhausted;
    const unsigned nm = __ballot_sync(0xFFFFFFFFu, need);
    if (nm) {
      const int leader = __ffs(nm) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(cnt + 3, (uint32_t)__popc(nm));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (need) {
        const uint32_t qi = base + (uint32_t)__popc(nm & ((1u << lane) - 1u));
        if (qi < n) {
          // one trip to memory: head, ray and light term of the request sit at its queue position
          const float4 q0 = __ldcs(rq0 + qi), q1 = __ldcs(rq1 + qi), q2 = __ldcs(rq2 + qi);
          org = v3(q0.x, q0.y, q0.z);
          bits = __float_as_uint(q0.w);
          path = __float_as_uint(q2.w);
          pos = qi;
          nsamp = (int)(bits & 255u);
          sample = 0; inc = v3s(0);
          has_req = true; has_ray = false; in_dda = false;
          if (!(bits & (1u << 16))) {  // single light: this is the ray
            D = normalize3(v3(q1.x, q1.y, q1.z));  // normalising Ray ctor of the shadow ray
            tmaxray = q1.w;
            li = v3(q2.x, q2.y, q2.z);
            if (use_mask) cand = p.st.smask[qi];
            vi = 0; has_ray = true;
            nrays++;
          }
        } else exhausted = true;
      }
    }
    if (!__any_sync(0xFFFFFFFFu, has_req)) break;
    if (has_req && !in_dda) {
      // finish the previous sample's volume loop / start the next sample
      for (;;) {
        if (has_ray) {
          bool entered = false;
          while (vi < nvol) {
            if (use_mask) {
              const uint32_t rest = cand >> vi;  // bit v of cand <=> volume v passes the bounding-sphere pre-test
              if (rest == 0u) { vi = nvol; break; }
              vi += __ffs(rest) - 1;
            } else if (nvol > 1 && !volume_may_hit(svol[vi], org, D)) { vi++; continue; }
            const DVolume& v = svol[vi];
            const ORay r = make_oray(xform_scalar(org, 1.0f, v.iM), xform_scalar(D, 0.0f, v.iM));
            vi++;
            if (setup_dda(v, r, s)) {
              entered = true; occ_reset(o); region = region_of(stg, v); fine = v.fine;
              if (BWK) bw_init(v, s, bw);
              break;
            }
          }
          if (entered) { in_dda = true; break; }
          // all volumes passed unoccluded: analytic primitives, then accept the sample
          bool occluded = false;
          if (sc.nsph + sc.ntri > 0) occluded = prims_any(sc, org, D, tmaxray);
          if (!occluded) inc = (bits & (1u << 16)) ? add(inc, li) : li;
          has_ray = false;
          sample++;
        }
        // next sample that needs a ray
        bool found = false;
        while (sample < nsamp) {
          const float4 l4 = sql[(size_t)sample * cap + path];
          if (l4.w != 0.0f) {
            const float4 d4 = sqd[(size_t)sample * cap + path];
            D = normalize3(v3(d4.x, d4.y, d4.z));  // normalising Ray ctor of the shadow ray
            tmaxray = d4.w;
            li = v3(l4.x, l4.y, l4.z);
            if (use_mask) cand = p.st.smask[(size_t)sample * cap + pos];
            vi = 0; has_ray = true; found = true;
            nrays++;
            break;
          }
          sample++;
        }
        if (!found) {
          // request complete: write A of the level it belongs to
          f3 A;
          if (bits & (1u << 16)) {
            f3 t = divs(inc, (float)nsamp);
            t = mul(t, albedo_of(sc, (int)(bits >> 24)));
            A = muls(t, lightCount);
          } else A = muls(inc, lightCount);
          const uint32_t level = (bits >> 8) & 31u;
          p.st.lv[(size_t)level * cap + path] = make_float4(A.x, A.y, A.z, __uint_as_float(((bits >> 13) & 7u) | ((bits >> 24) << 8)));
          has_req = false;
          break;
        }
      }
    }
    // walk while s.t < tlim: tlim = tmaxray while a volume is being traversed, -inf otherwise
    if (!in_dda) s.t = 3.0e38f;
    const unsigned thr = __any_sync(0xFFFFFFFFu, exhausted) ? 1u : (unsigned)p.refill;
    bool occluded = false;
    if (BWK) {
      // brick lanes / cell lanes as in k_wf_extend; an occupied cell ends the sample
      for (;;) {
#pragma unroll
        for (int u = 0; u < kBwUnroll; u++) {
          if (s.t < tmaxray) {
            if (walk_step(region, fine, s, bw, o, 3.0e38f)) occluded = true;
          }
        }
        if ((unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tmaxray)) < thr) break;
      }
    } else {
      const bool drained = thr == 1u;
      bool sparse = drained || (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tmaxray)) <= (unsigned)p.bulk;
      for (;;) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          if (s.t < tmaxray) {
            if (CS) steps++;
            uint32_t skipped = 0;
            const bool solid = occ_visit(region, fine, o, s, tmaxray, sparse, skipped);
            if (CS) steps += skipped;
            if (solid) { occluded = true; s.t = 3.0e38f; }
            else if (!dda_step(s)) s.t = 3.0e38f;
          }
        }
        const unsigned act = (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tmaxray));
        sparse = drained || act <= (unsigned)p.bulk;
        if (act < thr) break;
      }
    }
    if (in_dda) {
      in_dda = s.t < tmaxray;
      if (occluded) { has_ray = false; sample++; }  // occluded: the sample contributes nothing
    }
  }
  wf_add_counter(p.counters + 1, nrays);
  if (CS) wf_add_counter(p.counters + 3, steps);
}

// ---- K4 for single-volume scenes without area lights: prepared shadow rays and a spare request per lane --------------------
// The shadow stage spends 44 % of its warp time refilling lanes: the request's loads, then - at ~11 of 32 lanes - the part
// of Renderer::IsOccluded that precedes the walk (scalar world -> object transform, Setup3DDDA: ~240 instructions).  For
// scenes where a request is one ray into one volume, k_wf_shadow_prepare does that part one lane per request at full width
// and overwrites the request in place with the prepared walk (the records sit at the request's queue position, so this is
// coalesced both ways), and k_wf_shadow_spare walks them like k_wf_extend_spare walks bounce rays: every lane holds the
// ticket of a spare request and switches to it at the vote of the stepping loop.
//   rq0[pos] = tmax.xyz, t      rq1[pos] = tdelta.xyz, P'      sp2[pos] = E.xyz, flags (bit 0: the ray enters the volume,
//   bit 1: excluded from brick mode, bits 2..: the mirror word)    sp3[pos] = distance to the light, request bits
//   rq2[pos] = light term, path id (as the shade stage wrote it)
__global__ void __launch_bounds__(kWfBlock) k_wf_shadow_prepare(const __grid_constant__ WfParams p)
{
  const uint32_t n = p.st.counts[4 * p.iter + 1];
  const size_t par = (size_t)(p.iter & 1);
  float4* const rq0 = p.st.rq0 + par * p.st.cap;
  float4* const rq1 = p.st.rq1 + par * p.st.cap;
  float4* const sp2 = p.st.sp2 + par * p.st.cap;
  float4* const sp3 = p.st.sp3 + par * p.st.cap;
  const DVolume& v = p.sc.vol[0];
  for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n; qi += gridDim.x * blockDim.x) {
    const float4 q0 = rq0[qi], q1 = rq1[qi];
    const f3 org = v3(q0.x, q0.y, q0.z);
    const f3 D = normalize3(v3(q1.x, q1.y, q1.z));  // normalising Ray ctor of the shadow ray
    const ORay r = make_oray(xform_scalar(org, 1.0f, v.iM), xform_scalar(D, 0.0f, v.iM));
    DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = s.t = 0; s.P = 0;
    BW bw; bw.ex = bw.ey = bw.ez = 0; bw.plain = 0; bw.flip = 0; bw.F = 0;
    uint32_t flags = 0;
    if (setup_dda(v, r, s)) {
      bw_init(v, s, bw);
      flags = (bw.flip << 2) | (bw.plain << 1) | 1u;
    } else s.t = 3.0e38f;  // the request's ray enters no volume: beyond every stopping distance
    rq0[qi] = make_float4(s.tx, s.ty, s.tz, s.t);
    rq1[qi] = make_float4(s.dx, s.dy, s.dz, __uint_as_float(s.P));
    sp2[qi] = make_float4(bw.ex, bw.ey, bw.ez, __uint_as_float(flags));
    sp3[qi] = make_float4(q1.w, q0.w, 0, 0);
  }
}

template <int MINB>
__global__ void __launch_bounds__(kWfBlock, MINB) k_wf_shadow_spare(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  if (!wf_block_has_work(p.st.counts + 4 * p.iter + 3, p.st.counts[4 * p.iter + 1])) return;
  const Staged stg = wf_stage_scene(p.sc, smem);
  const DVolume& v = stg.vol[0];
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const DScene& sc = p.sc;
  uint32_t* cnt = p.st.counts + 4 * p.iter;
  const uint32_t n = cnt[1];
  const uint32_t cap = p.st.cap;
  const size_t par = (size_t)(p.iter & 1);
  const float4* const rq0 = p.st.rq0 + par * cap;
  const float4* const rq1 = p.st.rq1 + par * cap;
  const float4* const rq2 = p.st.rq2 + par * cap;
  const float4* const sp2 = p.st.sp2 + par * cap;
  const float4* const sp3 = p.st.sp3 + par * cap;
  const float lightCount = (float)(uint32_t)(sc.npl + sc.nsl + sc.nal + 1);
  const unsigned long long* const region = region_of(stg, v);
  const unsigned long long* const fine = v.fine;
  const uint32_t F = axis_field(v.jr, 0);
  const uint32_t spare_floor = (uint32_t)gridDim.x * (uint32_t)kWfBlock;
  bool has_req = false, exhausted = false, nomore = false, occluded = false;
  uint32_t path = 0, bits = 0, spare = kNoTicket;
  f3 li = v3s(0);
  float tmaxray = 0;
  DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = 0; s.t = 3.0e38f; s.P = s.R = 0; s.Ax = s.Ay = s.Az = s.Sx = s.Sy = s.Sz = 0;
  Occ o; occ_reset(o);
  BW bw; bw.ex = bw.ey = bw.ez = bw.tc = 0; bw.F = F; bw.flip = bw.ac = bw.plain = 0;
  unsigned int nrays = 0;
#define VP_LOAD_REQ(tk_)                                                                                              \
  {                                                                                                                   \
    const float4 w0_ = rq0[tk_], w1_ = rq1[tk_], w2_ = sp2[tk_], w3_ = sp3[tk_], w4_ = rq2[tk_];                      \
    const uint32_t fl_ = __float_as_uint(w2_.w);                                                                      \
    has_req = true; occluded = false; nrays++;                                                                        \
    tmaxray = w3_.x; bits = __float_as_uint(w3_.y);                                                                   \
    li = v3(w4_.x, w4_.y, w4_.z); path = __float_as_uint(w4_.w);                                                      \
    s.tx = w0_.x; s.ty = w0_.y; s.tz = w0_.z; s.t = w0_.w;                                                            \
    s.dx = w1_.x; s.dy = w1_.y; s.dz = w1_.z; s.P = __float_as_uint(w1_.w);                                           \
    bw.ex = w2_.x; bw.ey = w2_.y; bw.ez = w2_.z;                                                                      \
    bw.flip = fl_ >> 2; bw.plain = (fl_ >> 1) & 1u; bw.ac = kAcFirst; bw.tc = __uint_as_float(kTcFirst);                 \
    occ_reset(o);                                                                                                     \
  }
  for (;;) {
    if (!has_req && spare != kNoTicket) { VP_LOAD_REQ(spare); spare = kNoTicket; }
    {
      const bool ask_req = !has_req && !exhausted;
      const bool ask_spare = !exhausted && !nomore && spare == kNoTicket;
      const unsigned m1 = __ballot_sync(0xFFFFFFFFu, ask_req), m2 = __ballot_sync(0xFFFFFFFFu, ask_spare);
      if (m1 | m2) {
        const int leader = __ffs(m1 | m2) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(cnt + 3, (uint32_t)(__popc(m1) + __popc(m2)));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        uint32_t tk = base + (uint32_t)(__popc(m1 & lt) + __popc(m2 & lt));
        if (ask_req) {
          if (tk < n) { VP_LOAD_REQ(tk); } else { exhausted = true; }
          tk++;
        }
        if (ask_spare) {
          if (tk < n && n - tk > spare_floor) {
            spare = tk;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rq0 + tk));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rq1 + tk));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(sp2 + tk));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(sp3 + tk));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rq2 + tk));
          } else if (tk < n) {
            spare = tk;
            nomore = true;
          } else nomore = true;
        }
      }
    }
    if (!__any_sync(0xFFFFFFFFu, has_req)) break;
    const unsigned thr = __any_sync(0xFFFFFFFFu, exhausted | nomore) ? 1u : (unsigned)p.refill_spare;
    for (;;) {
#pragma unroll
      for (int u = 0; u < kBwUnroll; u++) {
        if (s.t < tmaxray) {
          if (walk_step(region, fine, s, bw, o, 3.0e38f)) occluded = true;
        }
      }
      if (has_req && !(s.t < tmaxray)) {
        // the ray is done: an unoccluded sample contributes its light term (Illumination: inc * lightCount)
        const f3 A = muls(occluded ? v3s(0) : li, lightCount);
        const uint32_t level = (bits >> 8) & 31u;
        p.st.lv[(size_t)level * cap + path] = make_float4(A.x, A.y, A.z, __uint_as_float(((bits >> 13) & 7u) | ((bits >> 24) << 8)));
        has_req = false;
        s.t = 3.0e38f; tmaxray = 0.0f;
        if (spare != kNoTicket) { VP_LOAD_REQ(spare); spare = kNoTicket; }
      }
      if ((unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tmaxray)) < thr) break;
    }
  }
#undef VP_LOAD_REQ
  wf_add_counter(p.counters + 1, nrays);
}

// ---- K6: resolve = fold the levels in the recursion's order, BlendColor, tonemap, RGB8 ------------
__global__ void __launch_bounds__(kWfBlock) k_wf_resolve(const __grid_constant__ WfParams p)
{
  const uint32_t cap = p.st.cap;
  for (uint32_t path = blockIdx.x * blockDim.x + threadIdx.x; path < cap; path += gridDim.x * blockDim.x) {
    int x, y;
    if (!wf_pixel_of(p, path, x, y)) continue;
    const uint32_t px = (uint32_t)y * (uint32_t)p.W + (uint32_t)x;
    const float4 t4 = p.st.term[path];
    const int top = (int)(__float_as_uint(t4.w) & 0xFFu);
    f3 R = v3(t4.x, t4.y, t4.z);
    for (int l = top - 1; l >= (p.reproject ? 1 : 0); l--) {
      const float4 v4 = __ldcs(p.st.lv + (size_t)l * cap + path);
      const uint32_t w = __float_as_uint(v4.w);
      const int type = (int)(w & 7u);
      const f3 V = v3(v4.x, v4.y, v4.z);  // M for type 0, A otherwise
      if (type == 0) R = mul(R, V);
      else if (type != 3) {
        const f3 M = albedo_of(p.sc, (int)((w >> 8) & 255u));
        if (type == 1) R = add(add(v3s(0), V), mul(R, M));
        else if (type == 2) R = mul(add(R, V), M);
        else R = mul(add(add(v3s(0), V), R), M);  // type 4, TraceNonMetal: albedo * ((0 + incLight) + child)
      }
    }
    if (p.reproject) {
      // the first hit keeps albedo and illumination apart (AlbedoIlluminationData of the primary ray, renderer.cpp:2013-2018)
      f3 alb = R, ill = v3s(1.0f);  // miss / emissive: {colour, 1}
      if (top > 0) {
        const float4 v4 = p.st.lv[path];
        const uint32_t w = __float_as_uint(v4.w);
        const int type = (int)(w & 7u);
        const f3 V = v3(v4.x, v4.y, v4.z);
        alb = type == 0 ? V : (type == 3 ? v3s(1.0f) : albedo_of(p.sc, (int)((w >> 8) & 255u)));
        ill = type == 0 ? R : (type == 4 ? add(add(v3s(0), V), R) : add(R, V));
      }
      p.rp_alb[3 * (size_t)px] = alb.x; p.rp_alb[3 * (size_t)px + 1] = alb.y; p.rp_alb[3 * (size_t)px + 2] = alb.z;
      p.rp_ill[3 * (size_t)px] = ill.x; p.rp_ill[3 * (size_t)px + 1] = ill.y; p.rp_ill[3 * (size_t)px + 2] = ill.z;
      continue;
    }
    float4 acc = p.accum[px];
    const float weight = 1.0f / ((float)p.frame + 1.0f);
    const float invWeight = 1.0f - weight;
    const float w1 = 1.0f - invWeight;
    acc.x = R.x * w1 + acc.x * invWeight;
    acc.y = R.y * w1 + acc.y * invWeight;
    acc.z = R.z * w1 + acc.z * invWeight;
    acc.w = 0.0f * w1 + acc.w * invWeight;
    p.accum[px] = acc;
    if (p.write_rgb8 && p.rgb8) p.rgb8[px] = rgb8(reinhard_jodie(v3(acc.x, acc.y, acc.z)));
  }
}

// ---- reprojection pass 2 (renderer.cpp:2023-2099): one lane per pixel -----------------------------------------
__device__ __forceinline__ bool rp_valid_screen(int x, int y, int W, int H)  // IsValidScreen on an int2 (renderer.cpp:1640-1643)
{
  return (float)x >= 0.0f && (float)x < (float)W && (float)y >= 0.0f && (float)y < (float)H;
}
__device__ __forceinline__ f3 rp_ycocg(f3 c)  // RGBToYCoCg, renderer.cpp:832-838
{
  const float k = 0.5f * 256.0f / 255.0f;
  return v3(dot3(c, v3(1, 2, 1)) * 0.25f, dot3(c, v3(2, 0, -2)) * 0.25f + k, dot3(c, v3(-1, 2, -1)) * 0.25f + k);
}
__device__ __forceinline__ f3 rp_rgb(f3 c)  // YCoCgToRGB, renderer.cpp:841-850
{
  const float k = 0.5f * 256.0f / 255.0f;
  const float Y = c.x, Co = c.y - k, Cg = c.z - k;
  return v3(Y + Co - Cg, Y + Cg, Y - Co - Cg);
}
__device__ __forceinline__ float rp_clamp(float f, float a, float b)  // clamp = fmaxf(a, fminf(f, b)), tmpl8math.h:401-409,2105
{
  const float lo = f < b ? f : b;
  return a > lo ? a : lo;
}
__device__ __forceinline__ f3 rp_sample_history(const float* hist, float u, float v, int W, int H)  // renderer.cpp:777-830
{
  u -= 1.0f / (float)W / 2.0f;
  v -= 1.0f / (float)H / 2.0f;
  const float px = u * (float)W, py = v * (float)H;
  const int tlx = f2i(px), tly = f2i(py);
  const float fx = px - (float)tlx, fy = py - (float)tly;
  const float gx = 1 - fx, gy = 1 - fy;
  const bool v1 = rp_valid_screen(tlx, tly, W, H), v2 = rp_valid_screen(tlx + 1, tly, W, H);
  const bool v3_ = rp_valid_screen(tlx, tly + 1, W, H), v4 = rp_valid_screen(tlx + 1, tly + 1, W, H);
  float w1 = v1 ? gx * gy : 0.0f, w2 = v2 ? fx * gy : 0.0f, w3 = v3_ ? gx * fy : 0.0f, w4 = v4 ? fx * fy : 0.0f;
  const float total = w1 + w2 + w3 + w4;
  const float rcp = 1.0f / total;
  w1 *= rcp; w2 *= rcp; w3 *= rcp; w4 *= rcp;
  f3 avg = v3s(0.0f);
  if (v1) avg = add(avg, muls(ld3(hist + 3 * ((size_t)tly * W + tlx)), w1));
  if (v2) avg = add(avg, muls(ld3(hist + 3 * ((size_t)tly * W + tlx + 1)), w2));
  if (v3_) avg = add(avg, muls(ld3(hist + 3 * ((size_t)(tly + 1) * W + tlx)), w3));
  if (v4) avg = add(avg, muls(ld3(hist + 3 * ((size_t)(tly + 1) * W + tlx + 1)), w4));
  return avg;
}
__device__ __forceinline__ f3 rp_clamp_history(f3 hist, f3 sample, const float* illum, int x, int y, int W, int H)  // renderer.cpp:855-910
{
  sample = rp_ycocg(sample);
  hist = rp_ycocg(hist);
  unsigned n = 1;
  f3 avg = sample, var = mul(sample, sample);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int k = i < 4 ? i : i + 1;  // the 3x3 ring in row-major order, centre skipped
    const int qx = x + (k % 3) - 1, qy = y + (k / 3) - 1;
    if (rp_valid_screen(qx, qy, W, H)) {
      const f3 f = rp_ycocg(ld3(illum + 3 * ((size_t)qx + (size_t)qy * W)));
      avg = add(avg, f);
      var = add(var, mul(f, f));
      n++;
    }
  }
  const float inv = 1.0f / (float)n;
  avg = muls(avg, inv);
  var = muls(var, inv);
  const f3 sigma = v3(sqrtf(stdmax(0.0f, var.x - (avg.x * avg.x))), sqrtf(stdmax(0.0f, var.y - (avg.y * avg.y))),
                      sqrtf(stdmax(0.0f, var.z - (avg.z * avg.z))));
  const f3 lo = sub(avg, muls(sigma, 0.75f)), hi = add(avg, muls(sigma, 0.75f));
  hist = v3(rp_clamp(hist.x, lo.x, hi.x), rp_clamp(hist.y, lo.y, hi.y), rp_clamp(hist.z, lo.z, hi.z));
  hist = rp_rgb(hist);
  return v3(hist.x > 0.0f ? hist.x : 0.0f, hist.y > 0.0f ? hist.y : 0.0f, hist.z > 0.0f ? hist.z : 0.0f);
}
__device__ __forceinline__ float rp_history_weight(int m)  // renderer.cpp:2051-2086
{
  if (m <= 4) return 0.8f;
  if (m <= 8) return 0.5f;
  if (m <= 14) return 0.9f;
  if (m == 15) return 0.0f;
  return 0.9f;
}

template <bool CS>
__global__ void __launch_bounds__(kWfBlock) k_wf_reproject(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged stg = wf_stage_scene(p.sc, smem);
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  const int W = p.W, H = p.H;
  const f3 pc = ld3(p.prev.cam_pos);
  const f3 nl = ld3(p.prev.left_normal), nr = ld3(p.prev.right_normal), nt = ld3(p.prev.top_normal), nb = ld3(p.prev.bottom_normal);
  for (uint32_t path = blockIdx.x * blockDim.x + threadIdx.x; path < p.st.cap; path += gridDim.x * blockDim.x) {
    int x, y;
    if (!wf_pixel_of(p, path, x, y) || !wf_row_owned(p, y)) continue;
    const size_t px = (size_t)x + (size_t)y * W;
    const float4 rd = p.rp_ray[px];
    const f3 P = v3(rd.x, rd.y, rd.z);
    const f3 newS = ld3(p.rp_ill + 3 * px);
    // Camera::PointToUV (template/camera.h:34-50) + half a pixel
    const f3 delta = sub(P, pc);
    const float dl = dot3(nl, delta), dr = dot3(nr, delta), dt = dot3(nt, delta), db = dot3(nb, delta);
    const float u = dl / (dl + dr) + 1.0f / (float)W / 2.0f;
    const float v = dt / (dt + db) + 1.0f / (float)H / 2.0f;
    bool reuse = u >= 0.0f && u < 1.0f && v >= 0.0f && v < 1.0f;  // IsValid
    if (reuse) {  // IsOccludedPrevFrame, renderer.cpp:767-774
      const f3 dn = normalize3(sub(P, pc));
      const f3 pos = offset_ray(P, neg(dn));
      const f3 dd = sub(pos, pc);
      reuse = !is_occluded<CS>(p.sc, stg, pc, dn, sqrtf(dot3(dd, dd)), cn);
    }
    f3 fin = newS;
    if (reuse) {
      f3 h = rp_sample_history(p.hist_in, u, v, W, H);
      h = rp_clamp_history(h, newS, p.rp_ill, x, y, W, H);
      const float w = rp_history_weight((int)(__float_as_uint(rd.w) & 255u));
      fin = add(newS, muls(sub(h, newS), w));  // lerp(a, b, t) = a + t * (b - a)
    }
    p.hist_out[3 * px] = fin.x; p.hist_out[3 * px + 1] = fin.y; p.hist_out[3 * px + 2] = fin.z;
    if (p.rgb8) p.rgb8[px] = rgb8(reinhard_jodie(mul(ld3(p.rp_alb + 3 * px), fin)));
  }
  wf_add_counter(p.counters + 1, cn.shadow);
  if (CS) wf_add_counter(p.counters + 3, cn.steps);
}

}  // namespace vp
