/* voxpop.h - C ABI of the B200-native VoxPopuli render hot path.
 *
 * Drop-in boundary for the path  Renderer::Update -> Trace -> FindNearest / IsOccluded /
 * Illumination -> accumulator  of Tycro-Games/Raytracer-VoxPopuli.  The reference has no
 * FFI or plugin interface; its seam is the `Renderer` class (reference renderer.h:42-245).
 * Every entry point below names the reference member it replaces.  Plain C types only:
 * no PyTorch, no C++ types, no exceptions cross this boundary.  All functions return 0 on
 * success and a negative VP_E* code on failure; vp_last_error() gives the message for the
 * calling thread.  There is NO CPU fallback behind these calls: without a CUDA device
 * vp_create fails with VP_ENODEVICE.
 *
 * The same POD structs describe a scene to the three implementations that must agree:
 *   libvoxpop_b200.so (this ABI, CUDA sm_100a)  - the product
 *   oracle/libvporacle.so (vpo_* , plain C)     - test infrastructure
 *   oracle/_ref/libvpref.so (vpref_*, the reference's own C++ built headless) - test infrastructure
 */
#ifndef VOXPOP_H
#define VOXPOP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_OK 0
#define VP_EINVAL (-1)    /* bad argument */
#define VP_ENODEVICE (-2) /* no usable CUDA device / CUDA runtime failure at create */
#define VP_ECUDA (-3)     /* CUDA error during upload or render */
#define VP_ENOMEM (-4)
#define VP_ESTATE (-5)    /* call order (e.g. render before set_volumes) */

#define VP_MAT_NONE 255   /* MaterialType::NONE, reference template/scene.h:56 */
#define VP_MAX_MATERIALS 256
#define VP_MAX_BOUNCES 31 /* Renderer::maxBounces limit of this build (reference default: 14) */

/* One voxel volume = reference `Scene` data members (template/scene.h:252-266).
 * grid is N*N*N bytes, index x + y*N + z*N*N (Scene::GetVoxel, template/scene.h:241-248),
 * 255 = empty.  matrix / inv_matrix are the row-major mat4 cells (tmpl8math.h:2596).
 * cube_min/cube_max = Scene::cube.b[0..1].  The library copies everything at upload. */
typedef struct VpVolume {
  const uint8_t* grid;
  uint32_t n;            /* gridsize (== worldsize), any value 1..512 (reference template/scene.h:253: "max 512") */
  float cube_min[3];
  float cube_max[3];
  float matrix[16];
  float inv_matrix[16];
} VpVolume;

/* reference src/Materials/Material.h:3-12 */
typedef struct VpMaterial {
  float albedo[3];
  float roughness;
  float emissive;
  float ior;
} VpMaterial;

/* reference src/Lighting/PointLight.h:3-7 */
typedef struct VpPointLight { float position[3]; float color[3]; } VpPointLight;
/* reference src/Lighting/SphereAreaLight.h:2-9 */
typedef struct VpAreaLight { float position[3]; float color[3]; float color_multiplier; float radius; } VpAreaLight;
/* reference src/Lighting/SpotLight.h:12-18; angle is the COSINE of the cone half angle */
typedef struct VpSpotLight { float position[3]; float direction[3]; float color[3]; float angle; } VpSpotLight;
/* reference src/Lighting/DirectionalLight.h:2-6; always counted in lightCount (renderer.cpp:2295) */
typedef struct VpDirLight { float direction[3]; float color[3]; } VpDirLight;

/* reference src/BVH/Shapes.h:4-66 */
typedef struct VpSphere { float center[3]; float radius; uint32_t material; } VpSphere;
/* reference src/BVH/Shapes.h:69-145: world vertices are position + vertexK */
typedef struct VpTriangle { float vertex0[3]; float vertex1[3]; float vertex2[3]; float position[3]; uint32_t material; } VpTriangle;

/* The fields of reference `Camera` (template/camera.h:175-183) that GetPrimaryRay reads. */
typedef struct VpCamera {
  float cam_pos[3];
  float top_left[3];
  float top_right[3];
  float bottom_left[3];
  float right[3];
  float up[3];
  float focal_distance;
  float defocus_jitter;
} VpCamera;

/* Per-call frame parameters = the Renderer members Update() reads
 * (renderer.h:175,186,199,205) plus the deterministic RNG contract:
 * every pixel of sample-frame f reseeds xorshift32 with
 *   s = WangHash(pixel_index + f * W*H + seed);  if (s == 0) s = 0x9E3779B9
 * (WangHash/xorshift32: reference template/tmpl8math.cpp:20-27,119-125). */
typedef struct VpFrameParams {
  int32_t width, height;
  int32_t max_bounces;     /* Renderer::maxBounces; Trace(depth = max_bounces) */
  int32_t spp;             /* number of Update() passes accumulated by this call */
  uint32_t frame_index;    /* Renderer::numRenderedFrames before the first pass */
  uint32_t seed;
  float aa_strength;       /* Renderer::antiAliasingStrength */
  int32_t row_begin, row_end; /* rows [row_begin,row_end) rendered; 0,0 = whole frame */
  int32_t row_tile, row_tile_stride, row_tile_phase; /* interleaved row tiles for multi-GPU:
                              row y is rendered iff ((y / row_tile) % row_tile_stride) == row_tile_phase;
                              row_tile == 0 disables the filter */
  uint32_t flags;          /* VP_FLAG_* */
} VpFrameParams;

#define VP_FLAG_COUNT_STEPS 1u /* also count DDA cell visits (instrumented kernel variant, slower) */
#define VP_FLAG_MEGAKERNEL 2u  /* single-kernel path walk instead of the wavefront pipeline (A/B measurement) */
#define VP_FLAG_TIME_STAGES 8u /* record CUDA events around every extend-stage launch (read with vp_stage_times) */
#define VP_FLAG_DEFER 16u      /* vp_render_device with stats == NULL only: frames in flight.  The call enqueues the frame on the
                                  library's own streams (after everything cuda_stream holds now) and returns WITHOUT making
                                  cuda_stream wait for it, so the latency-bound tail of this frame overlaps the start of the next
                                  call's.  Frames still resolve in call order (the accumulator is a running average).  Consume
                                  accum/rgb8 only behind vp_render_fence; any non-deferred call also waits for all earlier ones. */
#define VP_FLAG_PERSISTENT 4u  /* retired (round 2): the single persistent-kernel formulation lost its A/B against the staged
                                  launches at every frame size; the flag is accepted and ignored */

/* Result of one Renderer::FindNearest call (renderer.cpp:946-1018). */
typedef struct VpHit {
  int32_t volume;   /* index into volumes; -1 sphere/triangle; -2 miss */
  uint32_t cell;    /* x + y*N + z*N*N of the hit voxel; 0xFFFFFFFF if none */
  float t;          /* ray.t after the call (1e34f on miss) */
  float normal[3];  /* ray.rayNormal (world space, normalised) */
  uint8_t material; /* ray.indexMaterial; 255 on miss */
  int8_t face;      /* object-space face before the matrix: sx + 3*sy + 9*sz, s in {0:none,1:+1,2:-1}; -1 if none */
  uint8_t inside;   /* ray.isInsideGlass after the call */
  uint8_t pad;
} VpHit;

/* A ray as handed to Renderer::FindNearest / IsOccluded / Trace: O, D (used as given;
 * the reference always passes normalised D), t = max distance, inside = isInsideGlass. */
typedef struct VpRay { float o[3]; float d[3]; float t; uint32_t inside; } VpRay;

/* Ray counters, SURVEY section 8(d): one ray = one Renderer::FindNearest call
 * (primary or secondary) or one Renderer::IsOccluded call (shadow).  Medium-exit
 * traversals are reported separately and do not count toward Mrays/s. */
typedef struct VpStats {
  uint64_t primary_rays;
  uint64_t secondary_rays;
  uint64_t shadow_rays;
  uint64_t exit_rays;
  uint64_t dda_steps;       /* grid cell visits over all traversals (only with VP_FLAG_COUNT_STEPS) */
  uint64_t sky_samples;     /* paths that ended in Renderer::SampleSky */
  double device_ms;         /* device time of the render call (CUDA events) */
  uint64_t kernel_launches; /* kernels launched by this call */
} VpStats;

typedef struct vp_ctx vp_ctx;

/* Renderer::Init replacement (renderer.cpp:688-736, minus asset/game setup).
 * device < 0 selects the current CUDA device. */
vp_ctx* vp_create(int device);
/* The same for several GPUs of one box inside ONE process: the row fan-out of Renderer::Update
 * (renderer.cpp:1646-1665: rows handed to a thread pool) becomes interleaved 8-row tiles, tile k on devices[k mod ndev].
 * The context replicates the scene on every device (vp_set_* / vp_update_* go to all of them) and shards every
 * vp_render / vp_render_async / vp_render_device / vp_render_resident call itself (leave row_tile* unset):
 *  - host-buffer calls: every GPU moves only its own rows, over its own PCIe link, in and out of the caller's ONE
 *    accumulator and ONE image;
 *  - device-buffer calls: the buffers live on devices[0]; the other GPUs' resolve stage reads / writes the accumulator of
 *    its tiles and stores its RGB8 pixels straight into them over NVLink (peer access) - no pack buffer, no collective,
 *    no de-interleave.  cuda_stream must belong to devices[0].
 * Ray-batch seams, the reprojection mode and the gameplay queries run on devices[0].  ndev == 1 is vp_create. */
vp_ctx* vp_create_multi(const int* devices, int ndev);
int vp_device_count(vp_ctx* ctx);
void vp_destroy(vp_ctx* ctx);
const char* vp_last_error(void);

/* Renderer::voxelVolumes (renderer.h:217) - full replacement, uploaded once. */
int vp_set_volumes(vp_ctx* ctx, const VpVolume* volumes, int n);
/* Scene::Set / ResetGrid dirty range (template/scene.cpp:358-361,711-714): re-upload
 * bytes [offset, offset+len) of volume i's grid. */
int vp_update_volume_grid(vp_ctx* ctx, int volume, uint32_t offset, uint32_t len, const uint8_t* bytes);
/* Scene::SetTransform result (template/scene.cpp:373-404) for a moving volume. */
int vp_update_volume_transform(vp_ctx* ctx, int volume, const float matrix[16], const float inv_matrix[16],
                               const float cube_min[3], const float cube_max[3]);
/* Renderer::materials (renderer.h:196), n <= 256. */
int vp_set_materials(vp_ctx* ctx, const VpMaterial* materials, int n);
/* Renderer::pointLights/areaLights/spotLights/dirLight + numCheckShadowsAreaLight
 * (renderer.h:200-205); lightCount = np + na + ns + 1 (renderer.cpp:2290-2296). */
int vp_set_lights(vp_ctx* ctx, const VpPointLight* point, int np, const VpAreaLight* area, int na,
                  const VpSpotLight* spot, int ns, const VpDirLight* dir, int area_shadow_samples);
/* Renderer::spheres / triangles (renderer.h:213-214). */
int vp_set_prims(vp_ctx* ctx, const VpSphere* spheres, int ns, const VpTriangle* triangles, int nt);
/* Renderer::skyPixels/skyWidth/skyHeight/HDRLightContribution/activateSky
 * (renderer.h:225,232-234); rgb is w*h*3 floats. */
int vp_set_sky(vp_ctx* ctx, const float* rgb, int w, int h, float contribution, int active);

/* Renderer::Update (renderer.cpp:1646-1891), scalar pixel path (:1851-1881), `spp`
 * passes.  accum_rgba: W*H float4, in-out (Renderer::accumulator); rgb8: W*H uint32
 * (screen->pixels), may be NULL.  HOST pointers; copies are part of the call. */
int vp_render(vp_ctx* ctx, const VpCamera* cam, const VpFrameParams* fp, float* accum_rgba, uint32_t* rgb8,
              VpStats* stats);
/* Asynchronous Update(): the reference's Tick only needs frame k's pixels when it blits them (renderer.cpp:1972-2217), so
 * the call returns once the frame is enqueued - upload of accum_in (host, W*H float4; NULL = continue from the device copy
 * of the previous call's result), the trace, download into accum_out / rgb8_out (host; accum_out may equal accum_in) - and
 * vp_render_wait(ticket) blocks until that frame's downloads have landed.  Successive calls overlap like VP_FLAG_DEFER
 * frames: their copies are ordered (upload k+1 after download k), so the caller may pass the same accumulator every
 * frame, or alternate two (frame k: in A, out B; frame k+1: in B, out A) and read B while frame k+1 is in flight.  The
 * buffers of a call must stay untouched until its ticket has been waited for; at most 16 tickets are kept (the call
 * blocks on the oldest beyond that).  Page-locked buffers (vp_host_register) keep the copies asynchronous.  Results are
 * bit-identical to vp_render frame by frame. */
int vp_render_async(vp_ctx* ctx, const VpCamera* cam, const VpFrameParams* fp, const float* accum_in, float* accum_out,
                    uint32_t* rgb8_out, uint64_t* ticket);
int vp_render_wait(vp_ctx* ctx, uint64_t ticket);
/* cudaHostRegister / cudaHostUnregister for hosts that do not link the CUDA runtime (Renderer::accumulator is MALLOC64
 * memory, renderer.cpp:699). */
int vp_host_register(vp_ctx* ctx, void* ptr, uint64_t bytes);
int vp_host_unregister(vp_ctx* ctx, void* ptr);
/* Same work on DEVICE buffers owned by the caller (accumulator resident in HBM across
 * calls, e.g. a torch tensor that NCCL gathers afterwards).  accum_rgba_dev: W*H float4,
 * rgb8_dev: W*H uint32 or NULL.  cuda_stream: a cudaStream_t (NULL = default stream).
 * With stats == NULL the call only enqueues work and returns; otherwise it synchronises
 * the stream and fills the counters.  The samples of one call (spp > 1) are pipelined over
 * up to six internal stream pairs (eight for small calls); cuda_stream orders the call: it starts after what the
 * stream holds and (unless VP_FLAG_DEFER) the stream waits for its last resolve. */
int vp_render_device(vp_ctx* ctx, const VpCamera* cam, const VpFrameParams* fp, void* accum_rgba_dev, void* rgb8_dev,
                     void* cuda_stream, VpStats* stats);
/* Makes cuda_stream wait until the render call issued `frames_back` calls ago (0 = the most recent one,
 * at most 7) has resolved; with resolves in call order that covers every earlier call too.  Enqueue only,
 * no host synchronisation.  Needed after VP_FLAG_DEFER calls; harmless otherwise. */
int vp_render_fence(vp_ctx* ctx, void* cuda_stream, int frames_back);
/* ---- second frame mode: reprojection (Renderer::Tick, static-camera branch, renderer.cpp:1997-2101) -------------
 * Pass 1 per pixel: Camera::GetPrimaryRayNoDOF (template/camera.h:103-110), TraceReproject (renderer.cpp:1330-1586:
 * the Trace recursion with the first hit's albedo kept apart from the illumination and a Fresnel-free non-metal
 * branch) -> albedoBuffer, illuminationBuffer, rayData.  Pass 2 per pixel: project the hit into the previous frame
 * (Camera::PointToUV :34-50), IsOccludedPrevFrame (:767-774), SampleHistory (:777-830, bilinear), ClampHistory
 * (:855-910, 3x3 YCoCg variance box), lerp by a per-material weight (:2051-2086), tonemap albedo*illumination. */
typedef struct VpPrevCamera { /* what Renderer::CopyToPrevCamera keeps (renderer.cpp:1893-1902) */
  float cam_pos[3];
  float left_normal[3], right_normal[3], top_normal[3], bottom_normal[3]; /* Camera::SetFrustumNormals, camera.h:53-66 */
} VpPrevCamera;
/* history_in: W*H float3 = illuminationHistoryBuffer; history_out: W*H float3 = tempIlluminationBuffer (the caller
 * swaps them for the next frame, renderer.cpp:2100); must not alias.  albedo / illumination: optional W*H float3 dumps
 * of pass 1.  fp: width, height, max_bounces, frame_index, seed and the row selection are used (one pass).  With a row
 * selection (multi-GPU sharding) pass 1 also traces the rows adjacent to the owned ones - pass 2 reads their
 * illumination - and only the owned rows of history_out / rgb8 are written; history_in must hold the WHOLE previous
 * frame (gather it between frames).  HOST pointers. */
int vp_render_reproject(vp_ctx* ctx, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp,
                        const float* history_in, float* history_out, uint32_t* rgb8, float* albedo, float* illumination,
                        VpStats* stats);
/* Same on DEVICE buffers of the caller (history ping-pong resident in HBM); albedo/illumination dumps omitted. */
int vp_render_reproject_device(vp_ctx* ctx, const VpCamera* cam, const VpPrevCamera* prev, const VpFrameParams* fp,
                               const void* history_in_dev, void* history_out_dev, void* rgb8_dev, void* cuda_stream,
                               VpStats* stats);

/* ---- gameplay side channels of the frame loop ---------------------------------------------------
 * Renderer::inLight (renderer.cpp:1228-1240,1445-1449): set by Trace / TraceReproject when a smoke cell of volume 0 (the
 * player) is hit and the light sampled there exceeds sqrLength 16.  *in_light = 1 if that happened in any frame rendered
 * since the previous call, then the flag is cleared (Tick clears it once per frame, :2206).  Synchronises the device. */
int vp_in_light(vp_ctx* ctx, int* in_light);
/* Renderer::FindNearestPlayer (renderer.cpp:1020-1071) per ray: nearest voxel hit over volumes 1..V-1 (the player
 * volume 0 is skipped) that ignores the smoke materials SMOKE_LOW_DENSITY..SMOKE_PLAYER (Scene::FindNearestExcept,
 * template/scene.cpp:813-873); no spheres / triangles.  current_chunk reproduces the rule of :1068-1069 (chunk >= 2:
 * hits in the last three volumes, and misses, report volume -1); pass 0 to switch it off. */
int vp_find_nearest_player(vp_ctx* ctx, const VpRay* rays, int n, int current_chunk, VpHit* hits);
/* The autofocus ray of Tick (renderer.cpp:1987-1991): GetPrimaryRay(W/2, H/2) - its two lens draws taken from `seed` -
 * walked through every volume's Scene::FindNearest as the reference does (untransformed), focalDistance =
 * clamp(t, -1, 1e4). */
int vp_focus_distance(vp_ctx* ctx, const VpCamera* cam, int width, int height, uint32_t seed, float* focal_distance);

/* Library-owned resident accumulator: Renderer::accumulator kept in HBM between frames.
 * vp_accum_reset = allocate/zero (Renderer::ResetAccumulator, renderer.cpp:343-346, plus
 * the memset of Init :702); vp_render_resident = Update() on it; vp_accum_download
 * copies accumulator and/or RGB8 image to host (either pointer may be NULL). */
int vp_accum_reset(vp_ctx* ctx, int width, int height);
int vp_render_resident(vp_ctx* ctx, const VpCamera* cam, const VpFrameParams* fp, VpStats* stats);
int vp_accum_download(vp_ctx* ctx, float* accum_rgba, uint32_t* rgb8);

/* Primary-ray seam: Camera::GetPrimaryRay (template/camera.h:85-101) + Renderer::FindNearest
 * per pixel, RNG contract as above (2 AA draws then 2 disc draws). hits: W*H. */
int vp_trace_primary(vp_ctx* ctx, const VpCamera* cam, const VpFrameParams* fp, VpHit* hits);
/* Batch seams: Renderer::FindNearest (renderer.cpp:946), Renderer::IsOccluded (:209),
 * Renderer::Trace (:1076) with per-ray RNG seeds (xorshift32 state, must be non-zero). */
int vp_find_nearest(vp_ctx* ctx, const VpRay* rays, int n, VpHit* hits);
int vp_is_occluded(vp_ctx* ctx, const VpRay* rays, int n, uint8_t* out);
int vp_trace(vp_ctx* ctx, const VpRay* rays, const uint32_t* seeds, int n, int depth, float* rgb_out);

/* Cumulative counters over asynchronous launches: vp_stats_begin zeroes them, vp_stats_end
 * synchronises and reads them (rays of every vp_render* call in between; device_ms is not
 * filled - time with CUDA events on the stream you passed). */
int vp_stats_begin(vp_ctx* ctx);
int vp_stats_end(vp_ctx* ctx, VpStats* stats);

/* Sum of the device time of the extend-stage (k_wf_extend) launches made with VP_FLAG_TIME_STAGES since
 * the previous call, and their number; synchronises the device. */
int vp_stage_times(vp_ctx* ctx, double* extend_ms, uint64_t* extend_launches);

/* Self-check of a determinism pin: RandomSphereSample's sinf / cosf (reference template/tmpl8math.h:2502-2512)
 * are pinned to "fp64 sincos, one rounding to fp32"; the shade stage evaluates the fp64 routine's main path inline, because
 * its arguments all lie in [0, 2 pi].  The call runs both on every fp32 whose bit pattern lies in [first_bits, last_bits]
 * and reports how many results differ in either output (bit patterns compared) and the first such argument. */
int vp_check_sincos_pin(vp_ctx* ctx, uint32_t first_bits, uint32_t last_bits, uint64_t* mismatches, uint32_t* first_bad_bits);

/* ---- display interop ------------------------------------------------------------------------------
 * The template blits screen->pixels through an RGBA8 texture it re-uploads every frame (reference
 * template/opengl.cpp:144-150 GLTexture::CopyFrom with GL_BGRA, template/template.cpp:305-309).  vp_render_to_gl writes the
 * library's RGB8 image of the most recent vp_render / vp_render_async / vp_render_resident frame straight into that
 * texture (CUDA-GL interop: the texture is registered once, mapped per call, filled by a kernel through a surface, channel
 * order R, G, B, A = 255) - no device-to-host copy, no glTexImage2D.  gl_texture: the name of a GL_TEXTURE_2D with
 * internal format GL_RGBA8 and the frame's size; the OpenGL context that owns it must be current on the calling thread.
 * Headless build boxes have no GL: the call then fails with VP_ECUDA and a message (unmeasured on the B200 pool). */
int vp_render_to_gl(vp_ctx* ctx, uint32_t gl_texture);

/* ---- one image shared by several PROCESSES (one rank per GPU) ----------------------------------------
 * The display rank allocates the image in its GPU's HBM (vp_share_create) and hands the 128-byte VpShareInfo to the other
 * ranks by any transport; they map it (vp_share_open: CUDA IPC + peer access over NVLink) and pass vp_share_ptr() as the
 * rgb8_dev of vp_render_device, so their resolve stage stores its tiles straight into the display GPU's image - the
 * replacement for pack -> ncclAllGather -> de-interleave.  64 flag words in front of the payload order the ranks:
 * vp_share_signal(stream, slot, v) publishes v in flag `slot` (system scope) after everything the stream holds;
 * vp_share_wait(stream, first, n, v, timeout_s) makes the stream wait until the flags first..first+n-1 are >= v (it gives
 * up after timeout_s seconds - default 5 - and sets the word vp_share_error reads: a dead peer must not hang the GPU). */
typedef struct VpShareInfo { unsigned char handle[64]; uint64_t bytes; int32_t device; int32_t pad[13]; } VpShareInfo;
typedef struct vp_share vp_share;
vp_share* vp_share_create(vp_ctx* ctx, uint64_t bytes, VpShareInfo* info_out);
vp_share* vp_share_open(vp_ctx* ctx, const VpShareInfo* info);
void* vp_share_ptr(vp_share* share);
int vp_share_signal(vp_share* share, void* cuda_stream, int slot, uint32_t value);
int vp_share_wait(vp_share* share, void* cuda_stream, int first_slot, int nslots, uint32_t value, double timeout_s);
int vp_share_error(vp_share* share, uint32_t* error);
int vp_share_close(vp_share* share);

/* Build/runtime facts for the harness: fills up to `cap` bytes of a JSON string. */
int vp_info(vp_ctx* ctx, char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* VOXPOP_H */
