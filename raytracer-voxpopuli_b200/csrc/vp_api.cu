// vp_api.cu - extern "C" layer of libvoxpop_b200.so (include/voxpop.h).
// Plain CUDA runtime; no PyTorch, no CPU fallback: every compute entry point launches
// the sm_100a kernels of vp_kernels.cuh or fails with an error code.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vp_kernels.cuh"
#include "vp_wavefront.cuh"

using namespace vp;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(VP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct HostVolume {
  uint32_t n = 0, jr = 0;
  uint8_t* grid = nullptr;
  unsigned long long* fine = nullptr;  // 64 << 3*jr brick masks
  uint32_t region_off = 0;             // offset of this volume's 1 << 3*jr region masks in ctx->d_regions
  DVolume d{};
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n)
  {
    if (n <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) cap = n ? n : 1;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

// One frame in flight of the wavefront pipeline: its own stream pair, stage events and HBM-resident path state, so
// that the latency-bound tail of one Update() pass overlaps the wide first bounces of the next one.
// The launches of one pass up to (not including) its resolve, captured once per lane as a CUDA graph: ~75 kernel launches
// and ~60 event calls per pass become one cudaGraphLaunch plus one parameter update per kernel node.  What varies between
// frames (camera, frame index, rows, accumulator, grid sizes of the persistent stages) lives in the kernels' WfParams
// argument, which cudaGraphExecKernelNodeSetParams replaces; the topology only depends on what `key` holds.
struct PassGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  // what the captured topology depends on: bounce count, kernel variants (brick walk / prepare stage, single- or
  // multi-volume functions, the candidate-mask stage), the node that records ev_mid, dynamic shared memory, and the lane's buffers
  int max_bounces = -2, brickwalk = -1, nvol = -1, stagger = -99;
  size_t smem = 0;
  const void* counts = nullptr;  // the lane's buffers were (re)allocated: the memset node and every pointer changed
  // kind 0 = other, 1 = extend grid, 2 = shadow grid, 4 = shade grid, 3 = generate (the only graphed stage that reads per-frame fields)
  struct Node { cudaGraphNode_t node; cudaKernelNodeParams kp; int iter; int kind; WfParams last; unsigned grid; };
  std::vector<Node> nodes;
  unsigned long long updates = 0, replays = 0, captures = 0;  // node parameter updates issued / graph launches / captures (vp_info)
};

struct WfLane {
  PassGraph pg;
  cudaStream_t main = nullptr;  // generate, extend(k) -> shade(k), resolve
  cudaStream_t side = nullptr;  // shadow stage
  cudaEvent_t ev_done = nullptr;  // recorded after the lane's resolve
  cudaEvent_t ev_mid = nullptr;   // recorded after the shade stage of bounce c->tune_band_stagger (row bands start one after the other)
  cudaEvent_t ev_shade[VP_MAX_LEVELS] = {}, ev_shadow[VP_MAX_LEVELS] = {};
  WfState wf{};
  DevBuf<unsigned char> d_wf;
  uint32_t cap = 0;
  int levels = 0, samples = 0;
  bool prims = false, prepared = false, masks = false, area = false, sprep = false;  // optional parts laid out
  size_t bytes = 0;
};
constexpr int kLanes = 12;      // most frames in flight (c->nlanes are used: 6 unless VP_LANES says otherwise)
constexpr int kCallRing = 8;    // completion events kept for vp_render_fence

constexpr int kTickets = 16;     // vp_render_async calls whose completion events are kept
// One host thread per device of a multi-device context: a frame is ~25 runtime calls per device (copies, graph launch,
// events); issued from one thread they add up to more than a 1/8 frame takes on the GPUs.
struct KidWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = true, quit = false;
  int rc = 0;
  std::string err;
};
struct vp_ctx {
  // ---- multi-device context (vp_create_multi): owns one single-device context per GPU, the scene replicated on each;
  // render calls shard the frame as interleaved 8-row tiles (tile k -> kids[k mod n]) ----
  bool group = false;
  bool peer_ok = false;            // every device can store into kids[0]'s memory (NVLink / PCIe peer access)
  std::vector<vp_ctx*> kids;
  std::vector<KidWorker*> workers;  // one per kid (VP_GROUP_THREADS=0: issue from the calling thread)
  cudaEvent_t gev_in = nullptr;    // recorded on the caller's stream (device of kids[0]) when a device-buffer call enters
  unsigned long long gtickets[kTickets][16] = {};
  // ---- asynchronous host-buffer frames (vp_render_async / vp_render_wait) ----
  cudaEvent_t ev_ticket[kTickets] = {};
  unsigned long long nticket = 0;
  // ---- display interop (vp_render_to_gl): the registered GL texture ----
  cudaGraphicsResource* gl_res = nullptr;
  unsigned int gl_tex = 0;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  WfLane lane[kLanes];
  int next_lane = 0;
  int nlanes = 6;      // frames in flight (VP_LANES; 4 until the end of round 2: cfg 2 -2.5 %, cfg 4 -5.7 % with six and 32 hardware queues); lowered only when a lane's path state cannot be allocated
  bool lanes_fixed = false;
  cudaEvent_t ev_in = nullptr;      // recorded on the caller's stream when a render call enters
  cudaEvent_t ev_direct = nullptr;  // recorded after a launch made directly on the caller's stream (single-kernel formulations)
  bool direct_recorded = false;
  cudaEvent_t ev_last = nullptr;    // most recent resolve (any lane or direct): resolves are chained through it
  cudaEvent_t ev_call[kCallRing] = {};
  unsigned long long ncalls = 0;
  cudaStream_t copy = nullptr;  // vp_render: accumulator upload, overlapped with the trace (only the resolve stage reads it)
  cudaEvent_t ev_copy = nullptr;
  bool wait_copy = false;       // the next resolve / k_render must wait for ev_copy
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<HostVolume> vols;
  DevBuf<DVolume> d_vols;
  DevBuf<VpMaterial> d_mat;
  DevBuf<VpPointLight> d_pl;
  DevBuf<VpAreaLight> d_al;
  DevBuf<VpSpotLight> d_sl;
  DevBuf<VpSphere> d_sph;
  DevBuf<VpTriangle> d_tri;
  DevBuf<float> d_sky;
  DevBuf<unsigned long long> d_regions;  // region masks of all volumes, contiguous (staged to shared memory by the kernels)
  size_t region_words = 0;
  DevBuf<unsigned long long> d_counters;  // 8 counters + tile counter slot
  DevBuf<float4> d_accum;                 // library-owned resident accumulator / staging for vp_render
  DevBuf<uint32_t> d_rgb8;
  DevBuf<unsigned char> d_scratch_in, d_scratch_out, d_scratch_aux;
  // reprojection frame mode: pass-1 outputs (albedoBuffer, illuminationBuffer, rayData) and host-API staging of the history
  DevBuf<float> d_rp_alb, d_rp_ill, d_rp_hist_in, d_rp_hist_out;
  DevBuf<float4> d_rp_ray;
  int acc_w = 0, acc_h = 0;
  DScene sc{};
  bool have_volumes = false;
  int blocks_per_sm = 0;
  int wf_blocks_extend = 0, wf_blocks_shadow = 0;  // persistent grids at full load (occupancy x SMs)
  int wf_blocks_extend_mv = 0, wf_blocks_shadow_mv = 0;  // ... of the kernel variants multi-volume scenes run
  int wf_grid_extend = 0, wf_grid_shadow = 0;      // grids of the current call (smaller for small shards)
  bool wf_blocks_fixed = false;                    // VP_BLOCKS_* given: no per-call sizing
  unsigned long long host_launches = 0;
  unsigned long long in_light_seen = 0;  // value of the device's inLight counter at the previous vp_in_light
  int tune_spare = 1;         // VP_SPARE: single-volume scenes: every lane of the extend stage holds a spare ray (k_wf_extend_spare)
  int tune_refill_spare = 24; // VP_REFILL_SPARE (8: 3.83 ms, 16: 3.73, 20: 3.69, 24: 3.67, 28: 3.68 per cfg-2 frame)
  int wf_blocks_extend_spare = 0;
  uint32_t tune_sspare_min = 1u << 16;  // VP_SSPARE_MIN: calls of more paths than this run the prepared shadow stage (k_wf_shadow_prepare + k_wf_shadow_spare)
  int tune_shard_lanes = 2;   // VP_SHARD_LANES: passes in flight beyond nlanes for calls of at most 2^19 paths (a rank's share of a multi-GPU frame)
  int tune_ov_extend = 2, tune_ov_shadow = 1, tune_ov_shade = 1;  // VP_OV_EXTEND / _SHADOW / _SHADE: resident blocks per SM of a stage when passes of a single-volume scene overlap (0: no cap)
  int tune_refill = kRefill;  // VP_REFILL
  int tune_bulk = kBulkLanes; // VP_BULK_LANES
  int tune_brickwalk = 1;     // VP_BRICKWALK: 0 = cell walk only (A/B)
  int wf_grid_shade = 0;      // blocks of the shade stage of the current call (small shards: fewer blocks, each stages the scene once)
  int wf_blocks_shade = 0;    // ... at full load: what is resident at once (VP_BLOCKS_SHADE per SM fixes it)
  bool wf_shade_fixed = false;
  int tune_graph = 1;         // VP_GRAPH: 0 = launch every kernel of a pass from the host, 2 = graph + prepare stage for lone passes too
  int tune_prepare = 1;       // VP_PREPARE: 0 = the extend stage sets up every walk itself
  bool graph_now = false;     // this call's passes go through the lane graphs
  int tune_band_ramp = 0;     // VP_BAND_RAMP: band heights 1 : 2 : ... : n instead of equal
  int tune_band_stagger = -1; // VP_BAND_STAGGER: band k+1 of a vp_render call starts when band k has shaded this bounce (-1: all at once - the default since six passes are in flight: 4.25 -> 4.17 ms per synchronous cfg-2 frame; 0 until then)
  int band_last_lane = -1;    // lane the most recent pass went to
  int band_wait_lane = -1;    // lane whose ev_mid the next pass has to wait for (set by vp_render between its bands)
  int tune_bands = 6;         // VP_BANDS: row bands a vp_render call is cut into (1 = off; one per lane: 4 bands 4.50 ms, 6 bands 4.25 ms per synchronous cfg-2 frame)
  long long tune_band_min_pixels = 1 << 18;  // VP_BAND_MIN_PIXELS: smaller calls are not cut
  int tune_paths_per_thread = 8;  // VP_PATHS_PER_THREAD: paths of a call per thread of the persistent traversal grids (small shards: fewer, longer-lived warps)
  int tune_min_blocks = 2;        // VP_MIN_BLOCKS: ... but at least this many blocks per SM (1/8 frame shares: 0.636 ms with 1-2, 0.654 with 3, 0.669 with 4)
  // VP_FLAG_TIME_STAGES: CUDA-event pairs around every k_wf_extend launch, summed by vp_stage_times
  std::vector<cudaEvent_t> tev;
  size_t tev_used = 0;
  bool time_stages_now = false;
};

namespace {

int bind(vp_ctx* c)
{
  CU(cudaSetDevice(c->device));
  return 0;
}

void free_volume(HostVolume& v)
{
  if (v.grid) cudaFree(v.grid);
  if (v.fine) cudaFree(v.fine);
  v = HostVolume();
}

int build_masks(vp_ctx* c, HostVolume& v)
{
  const uint32_t nbricks = 64u << (3 * v.jr), nregions = 1u << (3 * v.jr);
  k_fine<<<(nbricks + 127) / 128, 128, 0, c->stream>>>(v.grid, v.n, v.jr, v.fine);
  k_region<<<(nregions + 127) / 128, 128, 0, c->stream>>>(v.fine, v.jr, c->d_regions.p + v.region_off);
  CU(cudaGetLastError());
  return 0;
}

// world-space bounding sphere of the cube: corners mapped through the inverse of inv_matrix (fp64), 1 % + 1e-4 margin
void bounding_sphere(const float* iM, const float* bmin, const float* bmax, float out[4])
{
  const double a = iM[0], b = iM[1], c = iM[2], d = iM[4], e = iM[5], f = iM[6], g = iM[8], h = iM[9], i = iM[10];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  out[0] = out[1] = out[2] = 0.0f;
  out[3] = INFINITY;
  if (!(fabs(det) > 1e-30) || !std::isfinite(det)) return;
  const double inv[9] = {(e * i - f * h) / det, (c * h - b * i) / det, (b * f - c * e) / det, (f * g - d * i) / det, (a * i - c * g) / det,
                         (c * d - a * f) / det, (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
  double P[8][3], ctr[3] = {0, 0, 0};
  for (int k = 0; k < 8; k++) {
    const double o[3] = {(k & 1 ? bmax[0] : bmin[0]) - iM[3], (k & 2 ? bmax[1] : bmin[1]) - iM[7], (k & 4 ? bmax[2] : bmin[2]) - iM[11]};
    for (int r = 0; r < 3; r++) { P[k][r] = inv[3 * r] * o[0] + inv[3 * r + 1] * o[1] + inv[3 * r + 2] * o[2]; ctr[r] += P[k][r] / 8.0; }
  }
  double r2 = 0;
  for (int k = 0; k < 8; k++) {
    const double dx = P[k][0] - ctr[0], dy = P[k][1] - ctr[1], dz = P[k][2] - ctr[2];
    r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
  }
  const double r = sqrt(r2) * 1.01 + 1e-4;
  if (!std::isfinite(r) || !std::isfinite(ctr[0] + ctr[1] + ctr[2])) return;
  out[0] = (float)ctr[0]; out[1] = (float)ctr[1]; out[2] = (float)ctr[2];
  out[3] = (float)(r * r);
}

void fill_dvolume(HostVolume& v, const float* M, const float* iM, const float* bmin, const float* bmax)
{
  bounding_sphere(iM, bmin, bmax, v.d.bsphere);
  memcpy(v.d.iM, iM, 12 * sizeof(float));
  memcpy(v.d.M, M, 12 * sizeof(float));
  memcpy(v.d.bmin, bmin, 12);
  memcpy(v.d.bmax, bmax, 12);
  v.d.n = v.n; v.d.jr = v.jr; v.d.region_off = v.region_off; v.d.pad = 0;
  v.d.grid = v.grid; v.d.fine = v.fine; v.d.region = nullptr;  // region pointer patched in upload_volume_table
}

int upload_volume_table(vp_ctx* c)
{
  std::vector<DVolume> tab(c->vols.size());
  for (size_t i = 0; i < c->vols.size(); i++) {
    tab[i] = c->vols[i].d;
    tab[i].region = c->d_regions.p + c->vols[i].region_off;
  }
  CU(c->d_vols.reserve(tab.size()));
  if (!tab.empty()) CU(cudaMemcpyAsync(c->d_vols.p, tab.data(), tab.size() * sizeof(DVolume), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->sc.vol = c->d_vols.p;
  c->sc.nvol = (int)tab.size();
  return 0;
}

template <class T>
int upload_array(vp_ctx* c, DevBuf<T>& buf, const T* src, int n)
{
  CU(buf.reserve((size_t)(n > 0 ? n : 1)));
  if (n > 0) CU(cudaMemcpyAsync(buf.p, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

size_t smem_bytes(const vp_ctx* c) { return (size_t)c->sc.nvol * sizeof(DVolume) + 16 + (size_t)c->sc.region_words * 8; }
// the shade stage also stages the light lists behind that, when they fit the 64 KB the kernels opt into
size_t light_bytes(const vp_ctx* c) { return (size_t)c->sc.npl * sizeof(VpPointLight) + (size_t)c->sc.nal * sizeof(VpAreaLight) + (size_t)c->sc.nsl * sizeof(VpSpotLight); }
size_t light_offset(const vp_ctx* c) { return (smem_bytes(c) + 15) & ~(size_t)15; }
bool lights_staged(const vp_ctx* c) { return light_bytes(c) > 0 && light_offset(c) + light_bytes(c) <= 64 * 1024 - 64; }
size_t smem_bytes_shade(const vp_ctx* c) { return lights_staged(c) ? light_offset(c) + light_bytes(c) : smem_bytes(c); }

int check_frame(const vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp)
{
  if (!c) return fail(VP_EINVAL, "null context");
  if (!cam || !fp) return fail(VP_EINVAL, "null camera / frame params");
  if (fp->width <= 0 || fp->height <= 0 || (long long)fp->width * fp->height > 0x7FFFFFFFll) return fail(VP_EINVAL, "bad resolution %dx%d", fp->width, fp->height);
  if (fp->max_bounces >= VP_MAX_LEVELS) return fail(VP_EINVAL, "max_bounces %d exceeds the supported %d", fp->max_bounces, VP_MAX_LEVELS - 1);
  if (fp->spp < 0) return fail(VP_EINVAL, "negative spp");
  if (!(fp->row_begin == 0 && fp->row_end == 0) && (fp->row_begin < 0 || fp->row_begin > fp->row_end || fp->row_end > fp->height))
    return fail(VP_EINVAL, "bad row range [%d,%d) for a frame of %d rows", fp->row_begin, fp->row_end, fp->height);
  if (fp->row_tile > 0 && (fp->row_tile_stride <= 0 || fp->row_tile_phase < 0 || fp->row_tile_phase >= fp->row_tile_stride))
    return fail(VP_EINVAL, "bad row tile interleave %d/%d/%d", fp->row_tile, fp->row_tile_stride, fp->row_tile_phase);
  if (!(c->group ? (!c->kids.empty() && c->kids[0]->have_volumes) : c->have_volumes)) return fail(VP_ESTATE, "vp_set_volumes has not been called");
  return 0;
}

void fill_params(const vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, RenderParams& p)
{
  p.sc = c->sc;
  p.cam = *cam;
  p.W = fp->width; p.H = fp->height;
  p.max_bounces = fp->max_bounces; p.spp = fp->spp;
  p.frame0 = fp->frame_index; p.seed = fp->seed; p.aa = fp->aa_strength;
  p.row_begin = fp->row_begin; p.row_end = fp->row_end;
  if (p.row_begin == 0 && p.row_end == 0) p.row_end = fp->height;
  p.row_tile = fp->row_tile; p.row_tile_stride = fp->row_tile_stride; p.row_tile_phase = fp->row_tile_phase;
  p.accum = nullptr; p.rgb8 = nullptr; p.hits = nullptr;
  p.counters = c->d_counters.p;
  p.tile_counter = reinterpret_cast<unsigned int*>(c->d_counters.p + 8);
}

int grid_blocks(const vp_ctx* c, long long work_warps)
{
  long long full = (long long)c->sm_count * c->blocks_per_sm;
  long long need = (work_warps + (kBlock / 32) - 1) / (kBlock / 32);
  if (need < 1) need = 1;
  return (int)(need < full ? need : full);
}

// path state of one lane (HBM-resident, sized for the largest pass / depth seen).  Optional parts are only laid out when
// a scene needs them: primitive normals (spheres / triangles), prepared walks (passes that overlap), candidate masks
// (multi-volume scenes).
struct WfNeed { uint32_t cap; int levels, samples; bool prims, prepared, masks, area, sprep; };
int wf_reserve(WfLane& L, const WfNeed& nd, size_t* bytes_only = nullptr)
{
  if (bytes_only) *bytes_only = 0;
  if (nd.cap <= L.cap && nd.levels <= L.levels && nd.samples <= L.samples && (!nd.prims || L.prims) && (!nd.prepared || L.prepared) &&
      (!nd.masks || L.masks) && (!nd.area || L.area) && (!nd.sprep || L.sprep) && L.d_wf.p)
    return 0;
  const uint32_t cap = nd.cap > L.cap ? nd.cap : L.cap;
  const int levels = nd.levels > L.levels ? nd.levels : L.levels;
  const int samples = nd.samples > L.samples ? nd.samples : L.samples;
  const bool prims = nd.prims || L.prims, prepared = nd.prepared || L.prepared, masks = nd.masks || L.masks, area = nd.area || L.area,
             sprep = nd.sprep || L.sprep;
  const size_t f4 = sizeof(float4);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_r0 = take(2 * cap * f4), o_r1 = take(2 * cap * f4), o_h0 = take(cap * 8), o_h1 = take(cap * 4), o_hn = take(prims ? cap * f4 : 0), o_term = take(cap * f4);
  const size_t o_lv = take((size_t)levels * cap * f4);
  const size_t o_rq0 = take((size_t)2 * cap * f4), o_rq1 = take((size_t)2 * cap * f4), o_rq2 = take((size_t)2 * cap * f4);
  const size_t o_sp2 = take(sprep ? (size_t)2 * cap * f4 : 0), o_sp3 = take(sprep ? (size_t)2 * cap * f4 : 0);
  const size_t o_sqd = take(area ? (size_t)2 * samples * cap * f4 : 0), o_sql = take(area ? (size_t)2 * samples * cap * f4 : 0);
  const size_t o_q0 = take(cap * 4), o_q1 = take(cap * 4), o_cnt = take((size_t)(VP_MAX_LEVELS + 2) * 4 * sizeof(uint32_t));
  const size_t o_ps = take(prepared ? 3 * cap * f4 : 0);  // prepared walks of the bounce being extended, by queue position
  const size_t o_sm = take(masks ? (size_t)samples * cap * 4 : 0);  // bounding-sphere candidates of the bounce's shadow rays
  if (bytes_only) { *bytes_only = off; return 0; }  // what growing this lane would take
  CU(cudaDeviceSynchronize());
  CU(L.d_wf.reserve(off));
  unsigned char* b = L.d_wf.p;
  WfState& w = L.wf;
  w.cap = cap;
  for (int k = 0; k < 2; k++) { w.r0[k] = (float4*)(b + o_r0) + (size_t)k * cap; w.r1[k] = (float4*)(b + o_r1) + (size_t)k * cap; }
  w.h0 = (float2*)(b + o_h0); w.h1 = (uint32_t*)(b + o_h1); w.hn = (float4*)(b + o_hn);
  w.term = (float4*)(b + o_term); w.lv = (float4*)(b + o_lv);
  w.rq0 = (float4*)(b + o_rq0); w.rq1 = (float4*)(b + o_rq1); w.rq2 = (float4*)(b + o_rq2);
  w.sp2 = (float4*)(b + o_sp2); w.sp3 = (float4*)(b + o_sp3);
  w.sqd = (float4*)(b + o_sqd); w.sql = (float4*)(b + o_sql); w.nsamp = (uint32_t)samples;
  w.q[0] = (uint32_t*)(b + o_q0); w.q[1] = (uint32_t*)(b + o_q1); w.counts = (uint32_t*)(b + o_cnt);
  for (int k = 0; k < 3; k++) w.ps[k] = (float4*)(b + o_ps) + (size_t)k * cap;
  w.smask = (uint32_t*)(b + o_sm);
  L.cap = cap; L.levels = levels; L.samples = samples; L.prims = prims; L.prepared = prepared; L.masks = masks; L.area = area; L.sprep = sprep;
  L.bytes = off;
  return 0;
}

// wavefront formulation: per pass  generate -> (extend, shade, shadow) x (max_bounces + 1) -> resolve
// One Update() pass on lane L.  Main chain on L.main: generate, then extend(k) -> shade(k) per bounce, resolve.
// The shadow stage of bounce k only feeds the fold (level k's light term), so it runs on the lane's
// side stream, overlapping the tails of the next bounce's extend/shade; shade(k+2) waits for shadow(k)
// before reusing its request buffer and resolve waits for all of them.  Resolves of successive passes are
// chained through c->ev_last (the accumulator is a running average: order matters); nothing else of a pass
// depends on the previous one, which is what lets kLanes passes be in flight.
// arguments of the reprojection frame mode (device pointers)
struct RpArgs {
  VpPrevCamera prev;
  const float* hist_in;
  float* hist_out;
};

// generate .. shadow(max_bounces) of one pass, the side stream joined back into L.main; issued directly or under capture
template <bool CS>
int enqueue_trace(vp_ctx* c, WfLane& L, WfParams& p, bool timing, bool capturing)
{
  cudaStream_t stream = L.main;
  const size_t sm = smem_bytes(c);
  CU(cudaMemsetAsync(p.st.counts, 0, (size_t)(VP_MAX_LEVELS + 2) * 4 * sizeof(uint32_t), stream));
  const int wide = c->sm_count * 8;
  k_wf_generate<<<wide, kWfBlock, 0, stream>>>(p);
  c->host_launches++;
  for (int it = 0; it <= p.max_bounces; it++) {
    p.iter = it;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) {
      if (c->tev_used + 2 > c->tev.size()) {
        for (int k = 0; k < 64; k++) { cudaEvent_t e; CU(cudaEventCreate(&e)); c->tev.push_back(e); }
      }
      e0 = c->tev[c->tev_used++]; e1 = c->tev[c->tev_used++];
    }
    if (!CS && p.prepared) {
      k_wf_prepare<<<wide, kWfBlock, 0, stream>>>(p);  // outside the extend stage's event pair: bench.py's roofline times k_wf_extend alone
      c->host_launches++;
    }
    if (e0) CU(cudaEventRecord(e0, stream));
    const bool spare = !CS && p.ps_fmt;  // single-volume scene without primitives: the records k_wf_prepare wrote are in that format
    if (spare) k_wf_extend_spare<kWfMinBlocksSpare><<<c->wf_grid_extend, kWfBlock, sm, stream>>>(p);
    else if (!CS && c->tune_brickwalk && p.sc.nvol > 1) k_wf_extend<false, true, kWfMinBlocksMulti><<<c->wf_grid_extend, kWfBlock, sm, stream>>>(p);
    else if (!CS && c->tune_brickwalk) k_wf_extend<false, true><<<c->wf_grid_extend, kWfBlock, sm, stream>>>(p);
    else k_wf_extend<CS, false><<<c->wf_grid_extend, kWfBlock, sm, stream>>>(p);
    if (e1) CU(cudaEventRecord(e1, stream));
    if (it >= 2) CU(cudaStreamWaitEvent(stream, L.ev_shadow[it - 2], 0));  // shade(k) reuses the id queue of shadow(k-2)
    k_wf_shade<CS><<<c->wf_grid_shade, kWfBlock, smem_bytes_shade(c), stream>>>(p);  // one resident wave: the stage is a grid-stride loop
    CU(cudaEventRecord(L.ev_shade[it], stream));
    if (it == c->tune_band_stagger || (it == p.max_bounces && p.max_bounces < c->tune_band_stagger))
      CU(cudaEventRecordWithFlags(L.ev_mid, stream, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));  // a node of its own when the pass is a graph
    CU(cudaStreamWaitEvent(L.side, L.ev_shade[it], 0));
    if (!CS && p.prepared && p.sc.nvol > 1 && p.sc.nvol <= 32) {
      k_wf_shadow_mask<<<wide, kWfBlock, 0, L.side>>>(p);
      c->host_launches++;
    }
    // (a rank's share of a multi-GPU frame is latency-bound: the extra kernel per bounce costs more than it saves there)
    const bool sspare = spare && p.sc.nal == 0 && !(c->tune_spare & 2) && p.st.cap > c->tune_sspare_min;
    if (sspare) {
      k_wf_shadow_prepare<<<wide, kWfBlock, 0, L.side>>>(p);
      c->host_launches++;
      k_wf_shadow_spare<kWfMinBlocksSpare><<<c->wf_grid_shadow, kWfBlock, sm, L.side>>>(p);
    }
    else if (!CS && c->tune_brickwalk && p.sc.nvol > 1) k_wf_shadow<false, true, kWfMinBlocksMulti><<<c->wf_grid_shadow, kWfBlock, sm, L.side>>>(p);
    else if (!CS && c->tune_brickwalk) k_wf_shadow<false, true><<<c->wf_grid_shadow, kWfBlock, sm, L.side>>>(p);
    else k_wf_shadow<CS, false><<<c->wf_grid_shadow, kWfBlock, sm, L.side>>>(p);
    CU(cudaEventRecord(L.ev_shadow[it], L.side));
    c->host_launches += 3;
  }
  for (int it = (p.max_bounces >= 1 ? p.max_bounces - 1 : 0); it <= p.max_bounces; it++) CU(cudaStreamWaitEvent(stream, L.ev_shadow[it], 0));
  return 0;
}

void drop_graph(PassGraph& G)
{
  if (G.exec) cudaGraphExecDestroy(G.exec);
  if (G.graph) cudaGraphDestroy(G.graph);
  G.exec = nullptr; G.graph = nullptr; G.nodes.clear(); G.max_bounces = -2;
}

// the trace part of a pass through the lane's graph; returns 1 when graphs are unusable (caller launches directly)
// Replay cost: of the ~77 kernel nodes only k_wf_generate reads what changes from frame to frame (camera, frame index,
// seed, rows); the traversal / shade nodes see the same arguments every frame unless the frame size or the grids of
// the persistent stages change.  A node's arguments are therefore compared with what it holds and only written on a
// difference: one cudaGraphExecKernelNodeSetParams per frame in steady state instead of one per node.
int graph_trace(vp_ctx* c, WfLane& L, WfParams& p)
{
  PassGraph& G = L.pg;
  const size_t sm = smem_bytes(c) + (smem_bytes_shade(c) << 20);
  const bool fresh = !(G.exec && G.max_bounces == p.max_bounces && G.brickwalk == c->tune_brickwalk * 2 + p.prepared + c->tune_spare * 4 + (p.sc.nal > 0 ? 64 : 0) + (p.sc.nsph + p.sc.ntri > 0 ? 128 : 0) + (p.st.cap > c->tune_sspare_min ? 256 : 0) && G.smem == sm &&
                       G.counts == (const void*)p.st.counts && G.nvol == p.sc.nvol && G.stagger == c->tune_band_stagger);
  if (fresh) {
    if (G.exec) CU(cudaStreamSynchronize(L.main));  // the old graph may still be running on this lane
    drop_graph(G);
    const unsigned long long before = c->host_launches;
    if (cudaStreamBeginCapture(L.main, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return 1; }
    const int rc = enqueue_trace<false>(c, L, p, false, true);
    const cudaError_t e = cudaStreamEndCapture(L.main, &G.graph);
    c->host_launches = before;
    if (rc || e != cudaSuccess || !G.graph || cudaGraphInstantiate(&G.exec, G.graph, 0) != cudaSuccess) { cudaGetLastError(); drop_graph(G); return 1; }
    size_t n = 0;
    cudaGraphGetNodes(G.graph, nullptr, &n);
    std::vector<cudaGraphNode_t> all(n);
    cudaGraphGetNodes(G.graph, all.data(), &n);
    for (cudaGraphNode_t nd : all) {
      cudaGraphNodeType ty;
      if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
      PassGraph::Node x; x.node = nd;
      if (cudaGraphKernelNodeGetParams(nd, &x.kp) != cudaSuccess) { cudaGetLastError(); drop_graph(G); return 1; }
      memcpy(&x.last, x.kp.kernelParams[0], sizeof(WfParams));
      x.iter = x.last.iter;
      x.grid = x.kp.gridDim.x;
      x.kind = (x.kp.func == (void*)k_wf_extend<false, true> || x.kp.func == (void*)k_wf_extend<false, false> ||
                x.kp.func == (void*)k_wf_extend<false, true, kWfMinBlocksMulti> || x.kp.func == (void*)k_wf_extend_spare<kWfMinBlocksSpare>) ? 1
             : (x.kp.func == (void*)k_wf_shadow<false, true> || x.kp.func == (void*)k_wf_shadow<false, false> ||
                x.kp.func == (void*)k_wf_shadow<false, true, kWfMinBlocksMulti> || x.kp.func == (void*)k_wf_shadow_spare<kWfMinBlocksSpare>) ? 2
             : (x.kp.func == (void*)k_wf_generate) ? 3 : (x.kp.func == (void*)k_wf_shade<false>) ? 4 : 0;
      G.nodes.push_back(x);
    }
    G.max_bounces = p.max_bounces; G.brickwalk = c->tune_brickwalk * 2 + p.prepared + c->tune_spare * 4 + (p.sc.nal > 0 ? 64 : 0) + (p.sc.nsph + p.sc.ntri > 0 ? 128 : 0) + (p.st.cap > c->tune_sspare_min ? 256 : 0); G.smem = sm; G.counts = (const void*)p.st.counts;
    G.nvol = p.sc.nvol; G.stagger = c->tune_band_stagger;
    G.captures++;
  } else {
    for (PassGraph::Node& x : G.nodes) {
      WfParams q;
      memcpy(&q, &p, sizeof(q));
      q.iter = x.iter;
      if (x.kind != 3 && !p.reproject) {
        // fields these stages never read keep the values the node already holds (see WfParams: per-frame fields)
        q.cam = x.last.cam; q.frame = x.last.frame; q.seed = x.last.seed; q.aa = x.last.aa;
        q.row_begin = x.last.row_begin; q.row_end = x.last.row_end; q.row_tile = x.last.row_tile;
        q.row_tile_stride = x.last.row_tile_stride; q.row_tile_phase = x.last.row_tile_phase;
        q.cmode = x.last.cmode; q.ctrow0 = x.last.ctrow0; q.ck = x.last.ck;
        q.accum = x.last.accum; q.rgb8 = x.last.rgb8; q.write_rgb8 = x.last.write_rgb8;
      }
      const unsigned grid = x.kind == 1 ? (unsigned)c->wf_grid_extend : x.kind == 2 ? (unsigned)c->wf_grid_shadow : x.kind == 4 ? (unsigned)c->wf_grid_shade : x.grid;
      if (grid == x.grid && memcmp(&q, &x.last, sizeof(q)) == 0) continue;
      void* args[1] = {&q};
      cudaKernelNodeParams kp = x.kp;
      kp.kernelParams = args;
      kp.gridDim = dim3(grid);
      CU(cudaGraphExecKernelNodeSetParams(G.exec, x.node, &kp));
      memcpy(&x.last, &q, sizeof(q));
      x.grid = grid;
      G.updates++;
    }
  }
  CU(cudaGraphLaunch(G.exec, L.main));
  G.replays++;
  c->host_launches += G.nodes.size();
  return 0;
}

template <bool CS>
int launch_wavefront_pass(vp_ctx* c, WfLane& L, WfParams& p)
{
  cudaStream_t stream = L.main;
  const size_t sm = smem_bytes(c);
  const int wide = c->sm_count * 8;
  bool traced = false;
  if (!CS && c->tune_graph && c->graph_now && !c->time_stages_now) {
    const int rc = graph_trace(c, L, p);
    if (rc < 0) return rc;
    if (rc == 1) c->tune_graph = 0;  // capture is not possible here: launch from the host from now on
    traced = rc == 0;
  }
  if (!traced)
    if (int rc = enqueue_trace<CS>(c, L, p, c->time_stages_now, false)) return rc;
  if (c->wait_copy) { CU(cudaStreamWaitEvent(stream, c->ev_copy, 0)); c->wait_copy = false; }
  if (c->ev_last && c->ev_last != L.ev_done) CU(cudaStreamWaitEvent(stream, c->ev_last, 0));
  k_wf_resolve<<<wide, kWfBlock, 0, stream>>>(p);
  c->host_launches++;
  if (p.reproject) {  // pass 2 reads the 3x3 neighbourhood of what resolve just wrote: a kernel boundary
    k_wf_reproject<CS><<<wide, kWfBlock, sm, stream>>>(p);
    c->host_launches++;
  }
  CU(cudaEventRecord(L.ev_done, stream));
  c->ev_last = L.ev_done;
  CU(cudaGetLastError());
  return 0;
}

// the caller's stream waits for every pass enqueued so far (resolves are chained, the last one covers all)
int join_lanes(vp_ctx* c, cudaStream_t stream)
{
  if (c->ev_last) CU(cudaStreamWaitEvent(stream, c->ev_last, 0));
  return 0;
}

// host-side wait for everything the lanes hold (scene edits, downloads, teardown)
int quiesce(vp_ctx* c)
{
  for (WfLane& L : c->lane)
    if (L.main) CU(cudaStreamSynchronize(L.main));
  if (c->copy) CU(cudaStreamSynchronize(c->copy));  // downloads of asynchronous frames
  return 0;
}

// Moves the rows of a [H][W] array of `elem`-byte pixels that a call with these frame parameters renders between host and
// device: a contiguous row range, or - with interleaved row tiles - one strided 2-D copy over the owned full tiles plus
// the ragged last tile when it is owned; tiles combined with a row range go tile by tile.  Rows a call does not own are
// never touched, so several devices can assemble one frame in one host buffer.
int copy_rows(const VpFrameParams* fp, void* dst, const void* src, size_t elem, cudaMemcpyKind kind, cudaStream_t st)
{
  const int H = fp->height;
  const size_t W = (size_t)fp->width;
  const bool whole = fp->row_begin == 0 && fp->row_end == 0;
  const int rb = whole ? 0 : fp->row_begin, re = whole ? H : fp->row_end;
  auto rows = [&](int y0, int y1) -> cudaError_t {
    if (y1 <= y0) return cudaSuccess;
    const size_t off = (size_t)y0 * W * elem;
    return cudaMemcpyAsync((char*)dst + off, (const char*)src + off, (size_t)(y1 - y0) * W * elem, kind, st);
  };
  if (fp->row_tile > 0 && fp->row_tile_stride > 1) {
    const int T = fp->row_tile, S = fp->row_tile_stride, ph = fp->row_tile_phase;
    if (rb == 0 && re == H) {
      const int nfull = H / T;
      const int owned = nfull > ph ? (nfull - ph + S - 1) / S : 0;
      const size_t chunk = (size_t)T * W * elem, pitch = chunk * (size_t)S, off = chunk * (size_t)ph;
      if (owned > 0) CU(cudaMemcpy2DAsync((char*)dst + off, pitch, (const char*)src + off, pitch, chunk, (size_t)owned, kind, st));
      if (H % T != 0 && nfull % S == ph) CU(rows(nfull * T, H));
      return 0;
    }
    for (int t = rb / T; t * T < re; t++)
      if (t % S == ph) CU(rows(t * T > rb ? t * T : rb, (t + 1) * T < re ? (t + 1) * T : re));
    return 0;
  }
  CU(rows(rb, re));
  return 0;
}

// a launch made directly on the caller's stream is finished: later lane passes order themselves after it
int mark_direct(vp_ctx* c, cudaStream_t stream)
{
  CU(cudaEventRecord(c->ev_direct, stream));
  c->direct_recorded = true;
  c->ev_last = c->ev_direct;
  return 0;
}

// tile rows (4 pixel rows each) that get path ids for this call, and how ids map to them (WfParams.cmode)
struct Compact { int mode, trow0, k; uint32_t trows; };
Compact compact_rows(const VpFrameParams* fp, int rb, int re, bool halo)
{
  Compact cp;
  const int H = fp->height;
  cp.mode = 0; cp.trow0 = 0; cp.k = 1; cp.trows = (uint32_t)((H + 3) / 4);
  if (halo || re <= rb) return cp;
  const bool whole = rb <= 0 && re >= H;
  if (fp->row_tile > 0 && fp->row_tile_stride > 1) {
    if (!whole || fp->row_tile % 4 != 0) return cp;
    const int ntile = (H + fp->row_tile - 1) / fp->row_tile;
    const int owned = ntile > fp->row_tile_phase ? (ntile - fp->row_tile_phase + fp->row_tile_stride - 1) / fp->row_tile_stride : 0;
    cp.mode = 2; cp.k = fp->row_tile / 4; cp.trows = (uint32_t)(owned > 0 ? owned : 1) * (uint32_t)cp.k;
    return cp;
  }
  if (whole) return cp;
  cp.mode = 1; cp.trow0 = rb / 4; cp.trows = (uint32_t)((re - 1) / 4 - rb / 4 + 1);
  return cp;
}

int launch_wavefront(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, float4* accum, uint32_t* rgb8, cudaStream_t stream, bool overlap,
                     const RpArgs* rp = nullptr)
{
  const int levels = fp->max_bounces + 1 > 1 ? fp->max_bounces + 1 : 1;
  const int samples = c->sc.area_samples > 1 ? c->sc.area_samples : 1;
  WfParams p;
  memset(&p, 0, sizeof(p));  // node arguments are compared bytewise (graph_trace): no indeterminate padding
  p.sc = c->sc; p.cam = *cam;
  p.W = fp->width; p.H = fp->height; p.max_bounces = fp->max_bounces;
  p.seed = fp->seed; p.aa = fp->aa_strength;
  p.row_begin = fp->row_begin; p.row_end = fp->row_end;
  if (p.row_begin == 0 && p.row_end == 0) p.row_end = fp->height;
  p.row_tile = fp->row_tile; p.row_tile_stride = fp->row_tile_stride; p.row_tile_phase = fp->row_tile_phase;
  p.accum = accum; p.rgb8 = rgb8; p.counters = c->d_counters.p; p.iter = 0;
  p.refill = c->tune_refill;
  p.refill_spare = c->tune_refill_spare;
  p.bulk = c->tune_bulk;
  p.light_smem = lights_staged(c) ? 1 : 0;
  p.light_smem_off = (int)light_offset(c);
  p.reproject = rp ? 1 : 0;
  p.halo = rp ? 1 : 0;  // no effect on whole frames
  p.rp_alb = c->d_rp_alb.p; p.rp_ill = c->d_rp_ill.p; p.rp_ray = c->d_rp_ray.p;
  p.hist_in = rp ? rp->hist_in : nullptr; p.hist_out = rp ? rp->hist_out : nullptr;
  if (rp) p.prev = rp->prev;
  const Compact cp = compact_rows(fp, p.row_begin, p.row_end, p.halo != 0);
  p.cmode = cp.mode; p.ctrow0 = cp.trow0; p.ck = cp.k;
  const uint32_t cap = (uint32_t)((fp->width + 7) / 8) * cp.trows * 32u;  // paths of this call
  c->time_stages_now = (fp->flags & VP_FLAG_TIME_STAGES) != 0;
  // Graph replay and the separate prepare stage pay off when passes overlap (deferred frames, several samples per call,
  // the row bands of vp_render): the host runs ahead and the extra kernel boundaries hide under other passes.  A pass the
  // caller waits for (a reprojection frame, a lone Update) is latency-bound: it is launched directly and its extend
  // stage sets up the walks itself (reprojection cfg 2: 5.94 against 6.32 ms per frame).
  const bool piped = overlap || fp->spp > 1 || c->tune_graph >= 2;
  c->graph_now = c->tune_graph != 0 && piped;
  // (a stage-timed pass is launched directly, but with the prepare stage and the kernel variants of the pipelined passes:
  // what bench.py times alone is what runs in the steady state)
  p.prepared = c->tune_prepare && (piped || c->time_stages_now) && c->tune_brickwalk && !(fp->flags & VP_FLAG_COUNT_STEPS) && c->sc.nvol <= 60;
  p.ps_fmt = p.prepared && c->tune_spare && c->sc.nvol == 1 && c->sc.nsph + c->sc.ntri == 0;
  WfNeed need;
  need.cap = cap; need.levels = levels; need.samples = samples;
  need.prims = c->sc.nsph + c->sc.ntri > 0; need.prepared = p.prepared != 0; need.masks = c->sc.nvol > 1 && c->sc.nvol <= 32; need.area = c->sc.nal > 0;
  need.sprep = p.prepared && c->tune_spare && c->tune_brickwalk && c->sc.nvol == 1 && c->sc.nsph + c->sc.ntri == 0 && c->sc.nal == 0 && cap > c->tune_sspare_min;
  // Persistent grids follow the number of paths this call can have: a warp that gets only ~1 ray per lane never refills
  // and ends at the SIMD efficiency of its longest ray, so small shards run fewer, longer-lived warps (>= 3 blocks per SM
  // to keep latency hidden; measured on 1/8 frame shares, profiles/README.md).
  {
    long long rows = 0;
    for (int y = p.row_begin; y < p.row_end && y < fp->height; y++)
      if (y >= 0 && (p.row_tile <= 0 || ((y / p.row_tile) % p.row_tile_stride) == p.row_tile_phase)) rows++;
    const long long paths = rows * (long long)fp->width;
    const long long per_sm = (long long)kWfBlock * c->tune_paths_per_thread * c->sm_count;
    long long want = (paths + per_sm - 1) / per_sm;
    if (want < c->tune_min_blocks) want = c->tune_min_blocks;
    const bool spare = c->tune_brickwalk && c->tune_spare && p.prepared && c->sc.nvol == 1 && c->sc.nsph + c->sc.ntri == 0 && !(fp->flags & VP_FLAG_COUNT_STEPS);
    const int be = spare ? c->wf_blocks_extend_spare : (c->sc.nvol > 1 ? c->wf_blocks_extend_mv : c->wf_blocks_extend);  // the kernel variant this scene runs
    const int bs = c->sc.nvol > 1 ? c->wf_blocks_shadow_mv : c->wf_blocks_shadow;
    const long long we = want * c->sm_count < be ? want * c->sm_count : be;
    const long long ws = want * c->sm_count < bs ? want * c->sm_count : bs;
    const bool full = c->wf_blocks_fixed || !(overlap || fp->spp > 1);  // a lone pass wants the lowest latency: full grid
    c->wf_grid_extend = full ? be : (int)we;
    c->wf_grid_shadow = full ? bs : (int)ws;
    const long long wh = want * c->sm_count < c->wf_blocks_shade ? want * c->sm_count : c->wf_blocks_shade;
    c->wf_grid_shade = (full || c->wf_shade_fixed) ? c->wf_blocks_shade : (int)wh;
    // Single-volume scenes whose passes overlap: the stages of the frames in flight (six) share the SMs, and a warp of a
    // smaller persistent grid sees more rays per lane (its lanes refill, its spares get used) - fewer resident blocks per
    // stage win there (cfg 2: 3.19 -> 2.9 ms per frame).  Multi-volume scenes lose with the same caps (cfg 5 +6 %): theirs stay.
    if (!full && spare) {
      const auto cap_to = [&](int& grid, int per_sm) { if (per_sm > 0 && grid > per_sm * c->sm_count) grid = per_sm * c->sm_count; };
      cap_to(c->wf_grid_extend, c->tune_ov_extend);
      cap_to(c->wf_grid_shadow, c->tune_ov_shadow);
      if (!c->wf_shade_fixed) cap_to(c->wf_grid_shade, c->tune_ov_shade);
    }
  }
  const bool cs = (fp->flags & VP_FLAG_COUNT_STEPS) != 0;
  CU(cudaEventRecord(c->ev_in, stream));
  bool entered[kLanes] = {};
  const bool rotate = overlap || fp->spp > 1;  // lanes rotate only when passes can overlap (several samples in this call, or a deferred call)
  // small shards (a rank's share of a multi-GPU frame) are latency-bound: two more passes in flight, their state is small
  const int lanes = (!c->lanes_fixed && cap <= (1u << 19) && c->nlanes + c->tune_shard_lanes <= kLanes) ? c->nlanes + c->tune_shard_lanes : c->nlanes;
  const int passes = rp ? 1 : fp->spp;
  for (int s = 0; s < passes; s++) {
    p.frame = fp->frame_index + (uint32_t)s;
    p.write_rgb8 = (s == passes - 1);
    int li = rotate ? (c->next_lane++ % lanes) : 0;
    if (int rc = wf_reserve(c->lane[li], need)) {  // grows the lane when this call needs more (synchronises the device then)
      if (li == 0) return rc;
      cudaGetLastError();  // out of memory for another lane: keep the frames in flight that fit
      c->nlanes = li; c->next_lane = 1; li = 0;
      if (int rc0 = wf_reserve(c->lane[0], need)) return rc0;
    }
    WfLane& L = c->lane[li];
    if (c->band_wait_lane >= 0 && c->band_wait_lane != li) CU(cudaStreamWaitEvent(L.main, c->lane[c->band_wait_lane].ev_mid, 0));
    c->band_last_lane = li;
    if (!entered[li]) {
      CU(cudaStreamWaitEvent(L.main, c->ev_in, 0));
      if (c->direct_recorded) CU(cudaStreamWaitEvent(L.main, c->ev_direct, 0));
      entered[li] = true;
    }
    p.st = L.wf;
    p.st.cap = cap;  // strides use the call's own cap
    if (int rc = cs ? launch_wavefront_pass<true>(c, L, p) : launch_wavefront_pass<false>(c, L, p)) return rc;
  }
  return 0;
}

// core: enqueue one render call (fp->spp Update() passes) for device buffers.  Work runs on the context's lanes; the
// caller's `stream` orders it: the passes start after what the stream holds now, and - unless the call is deferred
// (VP_FLAG_DEFER without stats) - the stream waits for the last resolve before anything enqueued later.
int launch_render(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, float4* accum, uint32_t* rgb8, cudaStream_t stream,
                  VpStats* stats, bool allow_defer, const RpArgs* rp = nullptr)
{
  RenderParams p;
  fill_params(c, cam, fp, p);
  p.accum = accum; p.rgb8 = rgb8;
  const bool defer = allow_defer && !stats && (fp->flags & VP_FLAG_DEFER) != 0;
  unsigned long long before[8] = {0};
  if (stats) {
    if (int rc = join_lanes(c, stream)) return rc;
    CU(cudaMemcpyAsync(before, c->d_counters.p, sizeof(before), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
  }
  const long long tiles = (long long)((p.W + kTileW - 1) / kTileW) * ((p.H + kTileH - 1) / kTileH);
  const int blocks = grid_blocks(c, tiles);
  const unsigned long long launches_before = c->host_launches;
  if (stats) CU(cudaEventRecord(c->ev0, stream));
  const bool mega = !rp && ((fp->flags & VP_FLAG_MEGAKERNEL) || c->sc.area_samples > kMaxAreaSamples || fp->max_bounces < 0);
  if (!mega) {
    if (int rc = launch_wavefront(c, cam, fp, accum, rgb8, stream, defer, rp)) return rc;
  } else {
    if (int rc = join_lanes(c, stream)) return rc;
    CU(cudaMemsetAsync(c->d_counters.p + 8, 0, sizeof(unsigned long long), stream));
    if (c->wait_copy) { CU(cudaStreamWaitEvent(stream, c->ev_copy, 0)); c->wait_copy = false; }
    if (fp->flags & VP_FLAG_COUNT_STEPS) k_render<true><<<blocks, kBlock, smem_bytes(c), stream>>>(p);
    else k_render<false><<<blocks, kBlock, smem_bytes(c), stream>>>(p);
    c->host_launches++;
    if (int rc = mark_direct(c, stream)) return rc;
  }
  CU(cudaGetLastError());
  // completion event of this call (vp_render_fence), then the join
  cudaEvent_t done = c->ev_call[c->ncalls % kCallRing];
  c->ncalls++;
  if (c->ev_last && c->ev_last != c->ev_direct) {
    // the last resolve sits on a lane: find it and record behind it
    cudaStream_t at = stream;
    for (WfLane& L : c->lane)
      if (L.ev_done == c->ev_last) at = L.main;
    CU(cudaEventRecord(done, at));
  } else {
    CU(cudaEventRecord(done, stream));
  }
  if (!defer) CU(cudaStreamWaitEvent(stream, done, 0));
  if (stats) {
    CU(cudaEventRecord(c->ev1, stream));
    unsigned long long h[8];
    CU(cudaMemcpyAsync(h, c->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    for (int k = 0; k < 8; k++) h[k] -= before[k];
    stats->primary_rays = h[4];
    stats->secondary_rays = h[0] - h[4];
    stats->shadow_rays = h[1];
    stats->exit_rays = h[2];
    stats->dda_steps = h[3];
    stats->sky_samples = h[6];
    stats->device_ms = ms;
    stats->kernel_launches = c->host_launches - launches_before;
  }
  return 0;
}

}  // namespace

#include "vp_multi.cuh"

// calls on a multi-device context: scene edits go to every device, queries of a single ray batch to the first one
#define KIDS(call)                          \
  if (c && c->group) {                      \
    for (vp_ctx* k : c->kids)               \
      if (int rc_ = (call)) return rc_;     \
    return 0;                               \
  }
#define KID0(call) \
  if (c && c->group) { vp_ctx* k = c->kids[0]; return (call); }

extern "C" {

const char* vp_last_error(void) { return g_err.c_str(); }

// Six passes in flight are twelve streams; with the default of 8 hardware queues independent kernels of different passes
// queue up behind each other.  The variable is read when the CUDA context is created, so a host that loads the library
// before it touches CUDA gets 32 queues; a value the host has set is kept.
__attribute__((constructor)) static void vp_default_hw_queues() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

vp_ctx* vp_create(int device)
{
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail(VP_ENODEVICE, "no CUDA device: %s (this library has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return nullptr;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev) { fail(VP_ENODEVICE, "device %d out of range (%d devices)", device, ndev); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { fail(VP_ENODEVICE, "cudaSetDevice(%d) failed", device); return nullptr; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail(VP_ENODEVICE, "cudaGetDeviceProperties failed"); return nullptr; }
  if (prop.major != 10) { fail(VP_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return nullptr; }
  vp_ctx* c = new vp_ctx;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  bool evok = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_direct, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < kCallRing && evok; i++) evok = cudaEventCreateWithFlags(&c->ev_call[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < kTickets && evok; i++) evok = cudaEventCreateWithFlags(&c->ev_ticket[i], cudaEventDisableTiming) == cudaSuccess;
  for (WfLane& L : c->lane) {
    evok = evok && cudaStreamCreateWithFlags(&L.main, cudaStreamNonBlocking) == cudaSuccess &&
           cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&L.ev_mid, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < VP_MAX_LEVELS && evok; i++)
      evok = cudaEventCreateWithFlags(&L.ev_shade[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&L.ev_shadow[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!evok || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
      cudaEventCreate(&c->ev1) != cudaSuccess || c->d_counters.reserve(16) != cudaSuccess ||
      cudaMemset(c->d_counters.p, 0, 16 * sizeof(unsigned long long)) != cudaSuccess) {
    fail(VP_ENODEVICE, "CUDA resource creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  k_fwd_table<<<1, 512, 0, c->stream>>>();  // forward boxes of the brick walk (vp_device.cuh); the same values for every context
  // default tables: 256 white diffuse materials (reference renderer.cpp:434-441), zero directional light
  std::vector<VpMaterial> mats(256);
  for (auto& m : mats) { m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.0f; m.roughness = 1.0f; m.emissive = 0.0f; m.ior = 1.5f; }
  if (upload_array(c, c->d_mat, mats.data(), 256) != 0) { delete c; return nullptr; }
  c->sc.mat = c->d_mat.p;
  c->sc.dl.direction[0] = 1; c->sc.dl.direction[1] = c->sc.dl.direction[2] = 0;
  c->sc.dl.color[0] = c->sc.dl.color[1] = c->sc.dl.color[2] = 0;
  c->sc.area_samples = 3;
  c->sc.sky_contrib = 1.0f;
  int occ = 0;
  cudaFuncSetAttribute(k_render<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_render<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_primary, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_nearest, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_nearest_player, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_focus, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_occluded, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_render<false>, kBlock, 1024) != cudaSuccess || occ < 1) occ = 4;
  c->blocks_per_sm = occ;
  cudaFuncSetAttribute(k_wf_extend<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_extend<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_extend<false, true, kWfMinBlocksMulti>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_extend<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shade<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shade<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shadow<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shadow<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shadow<false, true, kWfMinBlocksMulti>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shadow<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_reproject<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_reproject<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_extend_spare<kWfMinBlocksSpare>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k_wf_shadow_spare<kWfMinBlocksSpare>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wf_extend_spare<kWfMinBlocksSpare>, kWfBlock, 1024) != cudaSuccess || occ < 1) occ = 4;
  c->wf_blocks_extend_spare = occ * c->sm_count;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wf_extend<false, true>, kWfBlock, 1024) != cudaSuccess || occ < 1) occ = 4;
  c->wf_blocks_extend = occ * c->sm_count;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wf_shadow<false, true>, kWfBlock, 1024) != cudaSuccess || occ < 1) occ = 4;
  c->wf_blocks_shadow = occ * c->sm_count;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wf_extend<false, true, kWfMinBlocksMulti>, kWfBlock, 1024) != cudaSuccess || occ < 1) occ = 4;
  c->wf_blocks_extend_mv = occ * c->sm_count;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wf_shadow<false, true, kWfMinBlocksMulti>, kWfBlock, 1024) != cudaSuccess || occ < 1) occ = 4;
  c->wf_blocks_shadow_mv = occ * c->sm_count;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wf_shade<false>, kWfBlock, 1024) != cudaSuccess || occ < 1) occ = 6;
  c->wf_blocks_shade = c->wf_grid_shade = occ * c->sm_count;
  if (const char* e = getenv("VP_BLOCKS_SHADE")) { c->wf_blocks_shade = c->wf_grid_shade = atoi(e) * c->sm_count; c->wf_shade_fixed = true; }
  // tuning knobs for experiments (defaults are the shipped values)
  if (const char* e = getenv("VP_PATHS_PER_THREAD")) c->tune_paths_per_thread = atoi(e) < 1 ? 1 : atoi(e);
  if (const char* e = getenv("VP_MIN_BLOCKS")) c->tune_min_blocks = atoi(e) < 1 ? 1 : atoi(e);
  if (const char* e = getenv("VP_LANES")) { const int n = atoi(e); c->nlanes = n < 1 ? 1 : (n > kLanes ? kLanes : n); c->lanes_fixed = true; }
  if (const char* e = getenv("VP_BULK_LANES")) c->tune_bulk = atoi(e);
  if (const char* e = getenv("VP_REFILL")) c->tune_refill = atoi(e);
  if (const char* e = getenv("VP_SPARE")) c->tune_spare = atoi(e);
  if (const char* e = getenv("VP_REFILL_SPARE")) c->tune_refill_spare = atoi(e);
  if (const char* e = getenv("VP_BRICKWALK")) c->tune_brickwalk = atoi(e);
  if (const char* e = getenv("VP_GRAPH")) c->tune_graph = atoi(e);
  if (const char* e = getenv("VP_PREPARE")) c->tune_prepare = atoi(e);
  if (const char* e = getenv("VP_BAND_RAMP")) c->tune_band_ramp = atoi(e);
  if (const char* e = getenv("VP_BAND_STAGGER")) c->tune_band_stagger = atoi(e);
  if (const char* e = getenv("VP_BANDS")) c->tune_bands = atoi(e) < 1 ? 1 : atoi(e);
  if (const char* e = getenv("VP_BAND_MIN_PIXELS")) c->tune_band_min_pixels = atoll(e);
  if (const char* e = getenv("VP_SSPARE_MIN")) c->tune_sspare_min = (uint32_t)atoll(e);
  if (const char* e = getenv("VP_SHARD_LANES")) c->tune_shard_lanes = atoi(e) < 0 ? 0 : atoi(e);
  if (const char* e = getenv("VP_OV_EXTEND")) c->tune_ov_extend = atoi(e);
  if (const char* e = getenv("VP_OV_SHADOW")) c->tune_ov_shadow = atoi(e);
  if (const char* e = getenv("VP_OV_SHADE")) c->tune_ov_shade = atoi(e);
  if (const char* e = getenv("VP_BLOCKS_EXTEND")) { c->wf_blocks_extend = c->wf_blocks_extend_mv = c->wf_blocks_extend_spare = atoi(e) * c->sm_count; c->wf_blocks_fixed = true; }
  if (const char* e = getenv("VP_BLOCKS_SHADOW")) { c->wf_blocks_shadow = c->wf_blocks_shadow_mv = atoi(e) * c->sm_count; c->wf_blocks_fixed = true; }
  return c;
}

void vp_destroy(vp_ctx* c)
{
  if (!c) return;
  if (c->group) {
    for (KidWorker* w : c->workers) {
      { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
      w->cv.notify_all();
      if (w->th.joinable()) w->th.join();
      delete w;
    }
    for (vp_ctx* k : c->kids) vp_destroy(k);
    if (c->gev_in) { cudaSetDevice(c->device); cudaEventDestroy(c->gev_in); }
    delete c;
    return;
  }
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& v : c->vols) free_volume(v);
  c->d_vols.release(); c->d_mat.release(); c->d_pl.release(); c->d_al.release(); c->d_sl.release();
  c->d_sph.release(); c->d_tri.release(); c->d_sky.release(); c->d_regions.release(); c->d_counters.release(); c->d_accum.release();
  c->d_rp_alb.release(); c->d_rp_ill.release(); c->d_rp_ray.release(); c->d_rp_hist_in.release(); c->d_rp_hist_out.release();
  c->d_rgb8.release(); c->d_scratch_in.release(); c->d_scratch_out.release(); c->d_scratch_aux.release();
  for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
  for (WfLane& L : c->lane) {
    L.d_wf.release();
    for (int i = 0; i < VP_MAX_LEVELS; i++) {
      if (L.ev_shade[i]) cudaEventDestroy(L.ev_shade[i]);
      if (L.ev_shadow[i]) cudaEventDestroy(L.ev_shadow[i]);
    }
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.ev_mid) cudaEventDestroy(L.ev_mid);
    drop_graph(L.pg);
    if (L.side) cudaStreamDestroy(L.side);
    if (L.main) cudaStreamDestroy(L.main);
  }
  for (cudaEvent_t e : c->ev_call)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_ticket)
    if (e) cudaEventDestroy(e);
  if (c->gl_res) cudaGraphicsUnregisterResource(c->gl_res);
  if (c->ev_in) cudaEventDestroy(c->ev_in);
  if (c->ev_direct) cudaEventDestroy(c->ev_direct);
  if (c->copy) cudaStreamDestroy(c->copy);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int vp_set_volumes(vp_ctx* c, const VpVolume* vols, int n)
{
  KIDS(vp_set_volumes(k, vols, n))
  if (!c || n < 0 || (n > 0 && !vols)) return fail(VP_EINVAL, "bad arguments to vp_set_volumes");
  if ((size_t)n * sizeof(DVolume) + 16 > 64 * 1024) return fail(VP_EINVAL, "too many volumes (%d)", n);
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  for (int i = 0; i < n; i++)
    if (!vols[i].grid || vols[i].n == 0 || vols[i].n > 512) return fail(VP_EINVAL, "volume %d: bad grid pointer or size %u", i, vols[i].n);
  // the old scene is gone from here on: a failure below must not leave a table that points at freed grids
  c->have_volumes = false;
  c->sc.nvol = 0; c->sc.vol = nullptr; c->sc.region_words = 0;
  for (WfLane& L : c->lane) drop_graph(L.pg);  // the captured kernel variants and nodes depend on the volume count
  for (auto& v : c->vols) free_volume(v);
  c->vols.assign((size_t)n, HostVolume());
  size_t rwords = 0;
  for (int i = 0; i < n; i++) {
    HostVolume& v = c->vols[i];
    v.n = vols[i].n;
    uint32_t npad = 16;
    while (npad < v.n) npad <<= 1;
    v.jr = 0;
    while ((16u << v.jr) < npad) v.jr++;
    v.region_off = (uint32_t)rwords;
    rwords += (size_t)1 << (3 * v.jr);
  }
  c->region_words = rwords;
  auto upload = [&]() -> int {
    CU(c->d_regions.reserve(rwords ? rwords : 1));
    for (int i = 0; i < n; i++) {
      HostVolume& v = c->vols[i];
      const size_t bytes = (size_t)v.n * v.n * v.n;
      CU(cudaMalloc(&v.grid, bytes));
      CU(cudaMalloc(&v.fine, ((size_t)64 << (3 * v.jr)) * sizeof(unsigned long long)));
      CU(cudaMemcpyAsync(v.grid, vols[i].grid, bytes, cudaMemcpyHostToDevice, c->stream));
      if (build_masks(c, v)) return VP_ECUDA;
      fill_dvolume(v, vols[i].matrix, vols[i].inv_matrix, vols[i].cube_min, vols[i].cube_max);
    }
    return 0;
  };
  if (int rc = upload()) {  // partial allocations are released; the context stays without volumes (VP_ESTATE on render)
    cudaStreamSynchronize(c->stream);
    for (auto& v : c->vols) free_volume(v);
    c->vols.clear();
    return rc;
  }
  c->sc.regions = c->d_regions.p;
  // stage the region masks in shared memory when table + masks fit the 64 KB the kernels opt into
  const size_t budget = 64 * 1024 - 64;
  c->sc.region_words = ((size_t)n * sizeof(DVolume) + 16 + rwords * 8 <= budget) ? (int)rwords : 0;
  if (upload_volume_table(c)) return VP_ECUDA;
  c->have_volumes = true;
  return 0;
}

int vp_update_volume_grid(vp_ctx* c, int vi, uint32_t offset, uint32_t len, const uint8_t* bytes)
{
  KIDS(vp_update_volume_grid(k, vi, offset, len, bytes))
  if (!c || vi < 0 || vi >= (int)c->vols.size() || !bytes) return fail(VP_EINVAL, "bad arguments to vp_update_volume_grid");
  HostVolume& v = c->vols[vi];
  const size_t total = (size_t)v.n * v.n * v.n;
  if ((size_t)offset + len > total) return fail(VP_EINVAL, "dirty range [%u,+%u) exceeds grid of %zu bytes", offset, len, total);
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  CU(cudaMemcpyAsync(v.grid + offset, bytes, len, cudaMemcpyHostToDevice, c->stream));
  if (build_masks(c, v)) return VP_ECUDA;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_update_volume_transform(vp_ctx* c, int vi, const float M[16], const float iM[16], const float bmin[3], const float bmax[3])
{
  KIDS(vp_update_volume_transform(k, vi, M, iM, bmin, bmax))
  if (!c || vi < 0 || vi >= (int)c->vols.size() || !M || !iM || !bmin || !bmax) return fail(VP_EINVAL, "bad arguments to vp_update_volume_transform");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  fill_dvolume(c->vols[vi], M, iM, bmin, bmax);
  return upload_volume_table(c);
}

int vp_set_materials(vp_ctx* c, const VpMaterial* m, int n)
{
  KIDS(vp_set_materials(k, m, n))
  if (!c || !m || n < 0 || n > VP_MAX_MATERIALS) return fail(VP_EINVAL, "bad arguments to vp_set_materials (n=%d)", n);
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  if (n > 0) CU(cudaMemcpyAsync(c->d_mat.p, m, sizeof(VpMaterial) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_set_lights(vp_ctx* c, const VpPointLight* p, int np, const VpAreaLight* a, int na, const VpSpotLight* s, int ns,
                  const VpDirLight* d, int area_samples)
{
  KIDS(vp_set_lights(k, p, np, a, na, s, ns, d, area_samples))
  if (!c || np < 0 || na < 0 || ns < 0 || (np && !p) || (na && !a) || (ns && !s)) return fail(VP_EINVAL, "bad arguments to vp_set_lights");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  if (upload_array(c, c->d_pl, p, np) || upload_array(c, c->d_al, a, na) || upload_array(c, c->d_sl, s, ns)) return VP_ECUDA;
  c->sc.pl = c->d_pl.p; c->sc.npl = np;
  c->sc.al = c->d_al.p; c->sc.nal = na;
  c->sc.sl = c->d_sl.p; c->sc.nsl = ns;
  if (d) c->sc.dl = *d;
  c->sc.area_samples = area_samples;
  return 0;
}

int vp_set_prims(vp_ctx* c, const VpSphere* s, int ns, const VpTriangle* t, int nt)
{
  KIDS(vp_set_prims(k, s, ns, t, nt))
  if (!c || ns < 0 || nt < 0 || (ns && !s) || (nt && !t)) return fail(VP_EINVAL, "bad arguments to vp_set_prims");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  if (upload_array(c, c->d_sph, s, ns) || upload_array(c, c->d_tri, t, nt)) return VP_ECUDA;
  c->sc.sph = c->d_sph.p; c->sc.nsph = ns;
  c->sc.tri = c->d_tri.p; c->sc.ntri = nt;
  return 0;
}

int vp_set_sky(vp_ctx* c, const float* rgb, int w, int h, float contribution, int active)
{
  KIDS(vp_set_sky(k, rgb, w, h, contribution, active))
  if (!c || w < 0 || h < 0) return fail(VP_EINVAL, "bad arguments to vp_set_sky");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  const bool have = rgb && w > 0 && h > 0;
  if (have) {
    const size_t n = (size_t)w * h * 3;
    CU(c->d_sky.reserve(n));
    CU(cudaMemcpyAsync(c->d_sky.p, rgb, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  c->sc.sky = have ? c->d_sky.p : nullptr;
  c->sc.skyw = have ? w : 0; c->sc.skyh = have ? h : 0;
  c->sc.sky_contrib = contribution;
  c->sc.sky_active = (active && have) ? 1 : 0;
  return 0;
}

int vp_render_device(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, void* accum_dev, void* rgb8_dev, void* stream,
                     VpStats* stats)
{
  if (int rc = check_frame(c, cam, fp)) return rc;
  if (!accum_dev) return fail(VP_EINVAL, "null accumulator");
  if (c->group) return g_render_device(c, cam, fp, accum_dev, rgb8_dev, (cudaStream_t)stream, stats);
  if (bind(c)) return VP_ECUDA;
  return launch_render(c, cam, fp, (float4*)accum_dev, (uint32_t*)rgb8_dev, (cudaStream_t)stream, stats, true);
}

int vp_render_fence(vp_ctx* c, void* stream, int frames_back)
{
  if (!c || frames_back < 0 || frames_back >= kCallRing) return fail(VP_EINVAL, "bad arguments to vp_render_fence (frames_back %d, at most %d)", frames_back, kCallRing - 1);
  if (c->group) return g_render_fence(c, (cudaStream_t)stream, frames_back);
  if (bind(c)) return VP_ECUDA;
  if ((unsigned long long)frames_back >= c->ncalls) return 0;  // nothing that old was ever enqueued
  CU(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_call[(c->ncalls - 1 - (unsigned long long)frames_back) % kCallRing], 0));
  return 0;
}

static int check_reproject(const vp_ctx* c, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp)
{
  if (int rc = check_frame(c, cam, fp)) return rc;
  if (!prev) return fail(VP_EINVAL, "null previous camera");
  if (fp->max_bounces < 0) return fail(VP_EINVAL, "reprojection needs max_bounces >= 0");
  if (c->sc.area_samples > kMaxAreaSamples) return fail(VP_EINVAL, "reprojection supports at most %d area-light samples", kMaxAreaSamples);
  return 0;
}

int vp_render_reproject_device(vp_ctx* c, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp, const void* history_in_dev,
                               void* history_out_dev, void* rgb8_dev, void* stream, VpStats* stats)
{
  KID0(vp_render_reproject_device(k, cam, prev, fp, history_in_dev, history_out_dev, rgb8_dev, stream, stats))
  if (int rc = check_reproject(c, cam, prev, fp)) return rc;
  if (!history_in_dev || !history_out_dev || history_in_dev == history_out_dev) return fail(VP_EINVAL, "history in/out must be two distinct buffers");
  if (bind(c)) return VP_ECUDA;
  const size_t n = (size_t)fp->width * fp->height;
  if (c->d_rp_alb.cap < 3 * n || c->d_rp_ill.cap < 3 * n || c->d_rp_ray.cap < n) {
    if (quiesce(c)) return VP_ECUDA;
    CU(c->d_rp_alb.reserve(3 * n)); CU(c->d_rp_ill.reserve(3 * n)); CU(c->d_rp_ray.reserve(n));
  }
  RpArgs rp;
  rp.prev = *prev; rp.hist_in = (const float*)history_in_dev; rp.hist_out = (float*)history_out_dev;
  return launch_render(c, cam, fp, nullptr, (uint32_t*)rgb8_dev, (cudaStream_t)stream, stats, false, &rp);
}

int vp_render_reproject(vp_ctx* c, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp, const float* history_in,
                        float* history_out, uint32_t* rgb8, float* albedo, float* illumination, VpStats* stats)
{
  KID0(vp_render_reproject(k, cam, prev, fp, history_in, history_out, rgb8, albedo, illumination, stats))
  if (int rc = check_reproject(c, cam, prev, fp)) return rc;
  if (!history_in || !history_out || history_in == history_out) return fail(VP_EINVAL, "history in/out must be two distinct buffers");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  const size_t n = (size_t)fp->width * fp->height;
  CU(c->d_rp_hist_in.reserve(3 * n)); CU(c->d_rp_hist_out.reserve(3 * n)); CU(c->d_rgb8.reserve(n));
  CU(cudaMemcpyAsync(c->d_rp_hist_in.p, history_in, 3 * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  VpStats local;
  if (int rc = vp_render_reproject_device(c, cam, prev, fp, c->d_rp_hist_in.p, c->d_rp_hist_out.p, rgb8 ? c->d_rgb8.p : nullptr, c->stream,
                                          stats ? stats : &local))
    return rc;
  CU(cudaMemcpyAsync(history_out, c->d_rp_hist_out.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (rgb8) CU(cudaMemcpyAsync(rgb8, c->d_rgb8.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (albedo) CU(cudaMemcpyAsync(albedo, c->d_rp_alb.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (illumination) CU(cudaMemcpyAsync(illumination, c->d_rp_ill.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_accum_reset(vp_ctx* c, int w, int h)
{
  KID0(vp_accum_reset(k, w, h))
  if (!c || w <= 0 || h <= 0) return fail(VP_EINVAL, "bad arguments to vp_accum_reset");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  const size_t n = (size_t)w * h;
  CU(c->d_accum.reserve(n));
  CU(c->d_rgb8.reserve(n));
  CU(cudaMemsetAsync(c->d_accum.p, 0, n * sizeof(float4), c->stream));
  CU(cudaMemsetAsync(c->d_rgb8.p, 0, n * sizeof(uint32_t), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->acc_w = w; c->acc_h = h;
  return 0;
}

int vp_render_resident(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, VpStats* stats)
{
  if (int rc = check_frame(c, cam, fp)) return rc;
  if (c->group) {  // the resident accumulator lives on the first device; the others reach it through peer access
    vp_ctx* k = c->kids[0];
    if (k->acc_w != fp->width || k->acc_h != fp->height) return fail(VP_ESTATE, "vp_accum_reset(%d,%d) must precede vp_render_resident", fp->width, fp->height);
    return g_render_device(c, cam, fp, k->d_accum.p, k->d_rgb8.p, k->stream, stats);
  }
  if (c->acc_w != fp->width || c->acc_h != fp->height) return fail(VP_ESTATE, "vp_accum_reset(%d,%d) must precede vp_render_resident", fp->width, fp->height);
  if (bind(c)) return VP_ECUDA;
  return launch_render(c, cam, fp, c->d_accum.p, c->d_rgb8.p, c->stream, stats, false);
}

int vp_accum_download(vp_ctx* c, float* accum, uint32_t* rgb8)
{
  if (c && c->group) {
    for (vp_ctx* k : c->kids)
      if (bind(k) || quiesce(k)) return VP_ECUDA;
    return vp_accum_download(c->kids[0], accum, rgb8);
  }
  if (!c || c->acc_w <= 0) return fail(VP_ESTATE, "no resident accumulator");
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  const size_t n = (size_t)c->acc_w * c->acc_h;
  if (accum) CU(cudaMemcpyAsync(accum, c->d_accum.p, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (rgb8) CU(cudaMemcpyAsync(rgb8, c->d_rgb8.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_render(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, float* accum, uint32_t* rgb8, VpStats* stats)
{
  if (int rc = check_frame(c, cam, fp)) return rc;
  if (!accum) return fail(VP_EINVAL, "null accumulator");
  if (c->group) return g_render(c, cam, fp, accum, rgb8, stats);
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  const size_t n = (size_t)fp->width * fp->height;
  CU(c->d_accum.reserve(n));
  CU(c->d_rgb8.reserve(n));
  c->acc_w = fp->width; c->acc_h = fp->height;
  // Only the rows this call renders are moved (copy_rows)
  const int rb = (fp->row_begin == 0 && fp->row_end == 0) ? 0 : fp->row_begin;
  const int re = (fp->row_begin == 0 && fp->row_end == 0) ? fp->height : fp->row_end;
  const bool tiled = fp->row_tile > 0 && fp->row_tile_stride > 1;
  const size_t W = (size_t)fp->width;
  auto xfer = [&](void* dst, const void* src, size_t elem, cudaMemcpyKind kind, cudaStream_t st) -> int { return copy_rows(fp, dst, src, elem, kind, st); };
  // the accumulator is only read by the resolve stage: upload it on the copy stream while the frame is traced
  if (int rc = xfer(c->d_accum.p, accum, sizeof(float4), cudaMemcpyHostToDevice, c->copy)) return rc;
  CU(cudaEventRecord(c->ev_copy, c->copy));
  c->wait_copy = true;
  VpStats local;
  // Row bands: a synchronous call cannot keep several frames in flight, so the frame itself is cut into bands of rows that
  // run as deferred passes on the context's lanes (their latency-bound tails overlap like those of successive frames), and
  // the rows of a band travel back while the later bands are still being traced.  Pixels are independent and seeded per
  // pixel: the result does not depend on the cut.
  int nb = c->nlanes > 1 ? c->tune_bands : 1;  // more bands than lanes queue up on the lanes in turn
  if (nb > kCallRing) nb = kCallRing;
  const bool wavefront = fp->max_bounces >= 0 && c->sc.area_samples <= kMaxAreaSamples &&
                         !(fp->flags & (VP_FLAG_MEGAKERNEL | VP_FLAG_COUNT_STEPS | VP_FLAG_TIME_STAGES));
  if (tiled || !wavefront || fp->spp < 1 || (long long)(re - rb) * fp->width < c->tune_band_min_pixels || re - rb < 8 * nb) nb = 1;
  if (nb > 1) {
    VpStats* st = stats ? stats : &local;
    unsigned long long before[8] = {0}, after[8] = {0};
    if (int rc = join_lanes(c, c->stream)) return rc;
    if (stats) {
      CU(cudaMemcpyAsync(before, c->d_counters.p, sizeof(before), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    const unsigned long long launches_before = c->host_launches;
    CU(cudaEventRecord(c->ev0, c->stream));
    // Bands of growing height (1 : 2 : ... : nb): passes that start together advance bounce by bounce at the speed of
    // their own size, so the small bands finish first - their rows travel while the larger ones are still being traced,
    // and the latency-bound tails of the bands do not coincide.  Boundaries on multiples of 8 rows.
    int edge[kCallRing + 1];
    const int tri = nb * (nb + 1) / 2;
    edge[0] = rb;
    for (int k = 1, acc = 0; k <= nb; k++) {
      acc += c->tune_band_ramp ? k : 0;
      const long long cut = c->tune_band_ramp ? (long long)(re - rb) * acc / tri : (long long)(re - rb) * k / nb;
      edge[k] = k == nb ? re : rb + (int)((cut + 7) & ~7ll);
      if (edge[k] > re) edge[k] = re;
      if (edge[k] < edge[k - 1]) edge[k] = edge[k - 1];
    }
    cudaEvent_t done[kCallRing];
    for (int k = 0; k < nb; k++) {
      if (edge[k + 1] <= edge[k]) { done[k] = nullptr; continue; }
      VpFrameParams b = *fp;
      b.row_begin = edge[k]; b.row_end = edge[k + 1]; b.flags |= VP_FLAG_DEFER;
      done[k] = c->ev_call[c->ncalls % kCallRing];  // the completion event launch_render records for this call
      const int rc = launch_render(c, cam, &b, c->d_accum.p, rgb8 ? c->d_rgb8.p : nullptr, c->stream, nullptr, true);
      c->band_wait_lane = (c->tune_band_stagger >= 0 && b.spp == 1) ? c->band_last_lane : -1;  // the next band starts behind this one
      if (rc) { c->band_wait_lane = -1; return rc; }
    }
    c->band_wait_lane = -1;
    // copies after all launches: a copy into pageable memory blocks the host until its band is done
    for (int k = 0; k < nb; k++) {
      if (!done[k]) continue;
      CU(cudaStreamWaitEvent(c->copy, done[k], 0));
      const size_t off = (size_t)edge[k] * W, cnt = (size_t)(edge[k + 1] - edge[k]) * W;
      CU(cudaMemcpyAsync(accum + off * 4, (const char*)c->d_accum.p + off * sizeof(float4), cnt * sizeof(float4), cudaMemcpyDeviceToHost, c->copy));
      if (rgb8) CU(cudaMemcpyAsync(rgb8 + off, c->d_rgb8.p + off, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy));
    }
    if (int rc = join_lanes(c, c->stream)) return rc;
    CU(cudaEventRecord(c->ev1, c->stream));
    if (stats) CU(cudaMemcpyAsync(after, c->d_counters.p, sizeof(after), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    for (int k = 0; k < 8; k++) after[k] -= before[k];
    st->primary_rays = after[4];
    st->secondary_rays = after[0] - after[4];
    st->shadow_rays = after[1];
    st->exit_rays = after[2];
    st->dda_steps = after[3];
    st->sky_samples = after[6];
    st->device_ms = ms;
    st->kernel_launches = c->host_launches - launches_before;
    return 0;
  }
  if (int rc = launch_render(c, cam, fp, c->d_accum.p, rgb8 ? c->d_rgb8.p : nullptr, c->stream, stats ? stats : &local, false)) return rc;
  if (c->wait_copy) { CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0)); c->wait_copy = false; }  // spp == 0: nothing consumed the upload
  if (int rc = xfer(accum, c->d_accum.p, sizeof(float4), cudaMemcpyDeviceToHost, c->stream)) return rc;
  if (rgb8)
    if (int rc = xfer(rgb8, c->d_rgb8.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream)) return rc;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_trace_primary(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, VpHit* hits)
{
  KID0(vp_trace_primary(k, cam, fp, hits))
  if (int rc = check_frame(c, cam, fp)) return rc;
  if (!hits) return fail(VP_EINVAL, "null hits");
  if (bind(c)) return VP_ECUDA;
  const size_t n = (size_t)fp->width * fp->height;
  CU(c->d_scratch_out.reserve(n * sizeof(VpHit)));
  CU(cudaMemsetAsync(c->d_scratch_out.p, 0, n * sizeof(VpHit), c->stream));
  RenderParams p;
  fill_params(c, cam, fp, p);
  p.hits = reinterpret_cast<VpHit*>(c->d_scratch_out.p);
  CU(cudaMemsetAsync(c->d_counters.p + 8, 0, sizeof(unsigned long long), c->stream));
  const long long tiles = (long long)((p.W + kTileW - 1) / kTileW) * ((p.H + kTileH - 1) / kTileH);
  k_primary<<<grid_blocks(c, tiles), kBlock, smem_bytes(c), c->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(hits, c->d_scratch_out.p, n * sizeof(VpHit), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

static int batch_common(vp_ctx* c, const VpRay* rays, int n, BatchParams& p)
{
  if (!c || n < 0 || (n && !rays)) return fail(VP_EINVAL, "bad batch arguments");
  if (!c->have_volumes) return fail(VP_ESTATE, "vp_set_volumes has not been called");
  if (bind(c)) return VP_ECUDA;
  CU(c->d_scratch_in.reserve((size_t)(n ? n : 1) * sizeof(VpRay)));
  if (n) CU(cudaMemcpyAsync(c->d_scratch_in.p, rays, (size_t)n * sizeof(VpRay), cudaMemcpyHostToDevice, c->stream));
  memset(&p, 0, sizeof(p));
  p.sc = c->sc;
  p.rays = reinterpret_cast<const VpRay*>(c->d_scratch_in.p);
  p.n = n;
  return 0;
}

int vp_find_nearest(vp_ctx* c, const VpRay* rays, int n, VpHit* hits)
{
  KID0(vp_find_nearest(k, rays, n, hits))
  BatchParams p;
  if (int rc = batch_common(c, rays, n, p)) return rc;
  if (n == 0) return 0;
  if (!hits) return fail(VP_EINVAL, "null hits");
  CU(c->d_scratch_out.reserve((size_t)n * sizeof(VpHit)));
  p.hits = reinterpret_cast<VpHit*>(c->d_scratch_out.p);
  k_nearest<<<grid_blocks(c, (n + 31) / 32), kBlock, smem_bytes(c), c->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(hits, c->d_scratch_out.p, (size_t)n * sizeof(VpHit), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_find_nearest_player(vp_ctx* c, const VpRay* rays, int n, int current_chunk, VpHit* hits)
{
  KID0(vp_find_nearest_player(k, rays, n, current_chunk, hits))
  BatchParams p;
  if (int rc = batch_common(c, rays, n, p)) return rc;
  if (n == 0) return 0;
  if (!hits) return fail(VP_EINVAL, "null hits");
  CU(c->d_scratch_out.reserve((size_t)n * sizeof(VpHit)));
  p.hits = reinterpret_cast<VpHit*>(c->d_scratch_out.p);
  p.depth = current_chunk;
  k_nearest_player<<<grid_blocks(c, (n + 31) / 32), kBlock, smem_bytes(c), c->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(hits, c->d_scratch_out.p, (size_t)n * sizeof(VpHit), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_in_light(vp_ctx* c, int* in_light)
{
  if (!c || !in_light) return fail(VP_EINVAL, "bad arguments to vp_in_light");
  if (c->group) {  // any device may have shaded the player-volume smoke hit
    *in_light = 0;
    for (vp_ctx* k : c->kids) {
      int v = 0;
      if (int rc = vp_in_light(k, &v)) return rc;
      *in_light |= v;
    }
    return 0;
  }
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  CU(cudaDeviceSynchronize());
  unsigned long long now = 0;
  CU(cudaMemcpy(&now, c->d_counters.p + 5, sizeof(now), cudaMemcpyDeviceToHost));
  *in_light = now != c->in_light_seen ? 1 : 0;
  c->in_light_seen = now;
  return 0;
}

int vp_focus_distance(vp_ctx* c, const VpCamera* cam, int width, int height, uint32_t seed, float* focal_distance)
{
  KID0(vp_focus_distance(k, cam, width, height, seed, focal_distance))
  if (!c || !cam || !focal_distance || width <= 0 || height <= 0) return fail(VP_EINVAL, "bad arguments to vp_focus_distance");
  if (!c->have_volumes) return fail(VP_ESTATE, "vp_set_volumes has not been called");
  if (bind(c)) return VP_ECUDA;
  VpFrameParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.width = width; fp.height = height; fp.seed = seed; fp.row_tile_stride = 1;
  RenderParams p;
  fill_params(c, cam, &fp, p);
  CU(c->d_scratch_out.reserve(sizeof(float)));
  k_focus<<<1, 32, smem_bytes(c), c->stream>>>(p, reinterpret_cast<float*>(c->d_scratch_out.p));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(focal_distance, c->d_scratch_out.p, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_is_occluded(vp_ctx* c, const VpRay* rays, int n, uint8_t* out)
{
  KID0(vp_is_occluded(k, rays, n, out))
  BatchParams p;
  if (int rc = batch_common(c, rays, n, p)) return rc;
  if (n == 0) return 0;
  if (!out) return fail(VP_EINVAL, "null output");
  CU(c->d_scratch_out.reserve((size_t)n));
  p.occ = c->d_scratch_out.p;
  k_occluded<<<grid_blocks(c, (n + 31) / 32), kBlock, smem_bytes(c), c->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->d_scratch_out.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_trace(vp_ctx* c, const VpRay* rays, const uint32_t* seeds, int n, int depth, float* rgb)
{
  KID0(vp_trace(k, rays, seeds, n, depth, rgb))
  BatchParams p;
  if (int rc = batch_common(c, rays, n, p)) return rc;
  if (n == 0) return 0;
  if (!seeds || !rgb) return fail(VP_EINVAL, "null seeds / output");
  if (depth >= VP_MAX_LEVELS) return fail(VP_EINVAL, "depth %d exceeds the supported %d", depth, VP_MAX_LEVELS - 1);
  CU(c->d_scratch_aux.reserve((size_t)n * sizeof(uint32_t)));
  CU(cudaMemcpyAsync(c->d_scratch_aux.p, seeds, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CU(c->d_scratch_out.reserve((size_t)n * 3 * sizeof(float)));
  p.seeds = reinterpret_cast<const uint32_t*>(c->d_scratch_aux.p);
  p.depth = depth;
  p.rgb = reinterpret_cast<float*>(c->d_scratch_out.p);
  k_trace<<<grid_blocks(c, (n + 31) / 32), kBlock, smem_bytes(c), c->stream>>>(p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(rgb, c->d_scratch_out.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int vp_stats_begin(vp_ctx* c)
{
  KIDS(vp_stats_begin(k))
  if (!c) return fail(VP_EINVAL, "null context");
  if (bind(c)) return VP_ECUDA;
  CU(cudaDeviceSynchronize());
  CU(cudaMemset(c->d_counters.p, 0, 16 * sizeof(unsigned long long)));
  c->host_launches = 0;
  c->in_light_seen = 0;
  return 0;
}

int vp_stats_end(vp_ctx* c, VpStats* stats)
{
  if (!c || !stats) return fail(VP_EINVAL, "bad arguments to vp_stats_end");
  if (c->group) return g_stats_end(c, stats);
  if (bind(c)) return VP_ECUDA;
  CU(cudaDeviceSynchronize());
  unsigned long long h[8];
  CU(cudaMemcpy(h, c->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost));
  stats->primary_rays = h[4];
  stats->secondary_rays = h[0] - h[4];
  stats->shadow_rays = h[1];
  stats->exit_rays = h[2];
  stats->dda_steps = h[3];
  stats->sky_samples = h[6];
  stats->device_ms = 0;
  stats->kernel_launches = c->host_launches;
  return 0;
}

int vp_check_sincos_pin(vp_ctx* c, uint32_t first_bits, uint32_t last_bits, uint64_t* mismatches, uint32_t* first_bad_bits)
{
  KID0(vp_check_sincos_pin(k, first_bits, last_bits, mismatches, first_bad_bits))
  if (!c || !mismatches || last_bits < first_bits) return fail(VP_EINVAL, "bad arguments to vp_check_sincos_pin");
  if (bind(c)) return VP_ECUDA;
  DevBuf<unsigned long long> d;
  if (d.reserve(2)) return fail(VP_ECUDA, "vp_check_sincos_pin: out of device memory");
  const unsigned long long init[2] = {0ull, 0xFFFFFFFFull};
  unsigned long long out[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(d.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    k_check_sincos<<<c->sm_count * 16, 256, 0, c->stream>>>(first_bits, last_bits, d.p);
    e = cudaMemcpyAsync(out, d.p, sizeof(out), cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  d.release();
  CU(e);
  *mismatches = out[0];
  if (first_bad_bits) *first_bad_bits = (uint32_t)out[1];
  return VP_OK;
}

int vp_stage_times(vp_ctx* c, double* extend_ms, uint64_t* extend_launches)
{
  KID0(vp_stage_times(k, extend_ms, extend_launches))
  if (!c || !extend_ms || !extend_launches) return fail(VP_EINVAL, "bad arguments to vp_stage_times");
  if (bind(c)) return VP_ECUDA;
  CU(cudaDeviceSynchronize());
  double sum = 0;
  for (size_t i = 0; i + 1 < c->tev_used; i += 2) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->tev[i], c->tev[i + 1]));
    sum += ms;
  }
  *extend_ms = sum;
  *extend_launches = c->tev_used / 2;
  c->tev_used = 0;
  return 0;
}

int vp_info(vp_ctx* c, char* buf, int cap)
{
  if (!c || !buf || cap <= 0) return fail(VP_EINVAL, "bad arguments to vp_info");
  const int ndev = c->group ? (int)c->kids.size() : 1;
  const bool peer = c->group ? c->peer_ok : false;
  if (c->group) c = c->kids[0];  // the figures of the first device (all devices hold the same scene and lanes)
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess) return fail(VP_ECUDA, "cudaGetDeviceProperties failed");
  cudaFuncAttributes fa, fe, fs, fh;
  cudaFuncGetAttributes(&fa, k_render<false>);
  cudaFuncGetAttributes(&fe, k_wf_extend<false, true>);
  cudaFuncGetAttributes(&fs, k_wf_shadow<false, true>);
  cudaFuncGetAttributes(&fh, k_wf_shade<false>);
  unsigned long long dbg[16] = {0};
  if (bind(c)) return VP_ECUDA;
  cudaDeviceSynchronize();
  cudaMemcpy(dbg, c->d_counters.p, sizeof(dbg), cudaMemcpyDeviceToHost);
  // HBM held by this context on one device: scene, per-lane path state, frame buffers
  size_t scene = c->d_regions.cap * 8 + c->d_sky.cap * 4 + 256 * sizeof(VpMaterial), lanes = 0, frames = 0;
  for (const HostVolume& v : c->vols) scene += (size_t)v.n * v.n * v.n + ((size_t)64 << (3 * v.jr)) * 8;
  int lanes_held = 0;
  unsigned long long updates = 0, replays = 0, captures = 0;
  for (const WfLane& L : c->lane) { lanes += L.d_wf.p ? L.d_wf.cap : 0; lanes_held += L.d_wf.p ? 1 : 0; updates += L.pg.updates; replays += L.pg.replays; captures += L.pg.captures; }
  frames = c->d_accum.cap * sizeof(float4) + c->d_rgb8.cap * 4 + (c->d_rp_alb.cap + c->d_rp_ill.cap + c->d_rp_hist_in.cap + c->d_rp_hist_out.cap) * 4 +
           c->d_rp_ray.cap * sizeof(float4);
  snprintf(buf, (size_t)cap,
           "{\"device\": %d, \"devices\": %d, \"peer_access\": %s, \"name\": \"%s\", \"sm\": \"%d.%d\", \"sm_count\": %d, \"blocks_per_sm\": %d, \"block\": %d, "
           "\"k_render_regs\": %d, \"k_render_local_bytes\": %zu, \"k_wf_extend_regs\": %d, \"k_wf_shadow_regs\": %d, "
           "\"k_wf_shade_regs\": %d, \"k_wf_shade_local_bytes\": %zu, \"wf_blocks_extend\": %d, \"wf_blocks_shadow\": %d, \"volumes\": %d, \"max_levels\": %d, "
           "\"region_words_staged\": %d, \"dbg_extend_steps_in_empty_region\": %llu, \"dbg_extend_steps_in_empty_brick\": %llu, \"extend_steps\": %llu, "
           "\"hbm_scene_bytes\": %zu, \"hbm_lane_bytes\": %zu, \"hbm_frame_bytes\": %zu, \"hbm_in_use\": %zu, \"lanes\": %d, \"lanes_allocated\": %d, "
           "\"graph_replays\": %llu, \"graph_node_updates\": %llu, \"graph_captures\": %llu, \"graph_enabled\": %d, \"extend_kernel\": \"%s\"}",
           c->device, ndev, peer ? "true" : "false", prop.name, prop.major, prop.minor, c->sm_count, c->blocks_per_sm, kBlock, fa.numRegs, fa.localSizeBytes,
           fe.numRegs, fs.numRegs, fh.numRegs, fh.localSizeBytes, c->wf_blocks_extend, c->wf_blocks_shadow, c->sc.nvol, VP_MAX_LEVELS,
           c->sc.region_words, dbg[9], dbg[10], dbg[11], scene, lanes, frames, scene + lanes + frames, c->nlanes, lanes_held, replays, updates, captures, c->tune_graph,
           (c->tune_prepare && c->tune_brickwalk && c->tune_spare && c->sc.nvol == 1 && c->sc.nsph + c->sc.ntri == 0) ? "k_wf_extend_spare" : "k_wf_extend");
  return 0;
}

}  // extern "C"
