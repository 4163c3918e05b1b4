// vp_multi.cuh - the parts of the C ABI that go beyond one synchronous frame on one GPU (included by vp_api.cu):
//
//   * asynchronous host-buffer frames (vp_render_async / vp_render_wait): Renderer::Update() whose copies and
//     latency-bound tail overlap the next frame's trace, the way Tick only needs frame k's pixels when it blits them
//     (reference renderer.cpp:1972-2217);
//   * multi-device contexts (vp_create_multi): the row fan-out of Renderer::Update (reference renderer.cpp:1646-1665)
//     across the GPUs of one box inside ONE process - scene replicated, interleaved 8-row tiles, every GPU moves its own
//     rows over its own PCIe link into the caller's single host buffer, or stores them straight into the display GPU's
//     image over NVLink (peer stores from k_wf_resolve; no pack buffer, no collective, no de-interleave);
//   * shared images between PROCESSES (vp_share_*): the same peer stores when every GPU is driven by its own process
//     (one rank per GPU): CUDA IPC mapping of an image in the display GPU's HBM plus system-scope arrival flags.
#pragma once

namespace {

// ---- asynchronous host-buffer frames on one device ---------------------------------------------------------------
int single_render_async(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, const float* accum_in, float* accum_out, uint32_t* rgb8_out,
                        unsigned long long* ticket)
{
  if (bind(c)) return VP_ECUDA;
  const size_t n = (size_t)fp->width * fp->height;
  if (c->d_accum.cap < n || c->d_rgb8.cap < n || !c->d_accum.p || !c->d_rgb8.p) {
    if (quiesce(c)) return VP_ECUDA;
    CU(c->d_accum.reserve(n));
    CU(c->d_rgb8.reserve(n));
  }
  c->acc_w = fp->width; c->acc_h = fp->height;
  const int slot = (int)(c->nticket % kTickets);
  if (c->nticket >= (unsigned long long)kTickets) CU(cudaEventSynchronize(c->ev_ticket[slot]));  // the caller is a whole ring ahead
  // Copies of successive frames are ordered on the copy stream: the upload of frame k+1 follows the download of frame k,
  // which follows resolve(k) - so the same host accumulator can be handed from call to call (or two of them alternated)
  // while the traces, which do not read the accumulator, run ahead on the lanes.
  if (accum_in) {
    if (int rc = copy_rows(fp, c->d_accum.p, accum_in, sizeof(float4), cudaMemcpyHostToDevice, c->copy)) return rc;
    CU(cudaEventRecord(c->ev_copy, c->copy));
    c->wait_copy = true;
  }
  VpFrameParams f = *fp;
  f.flags |= VP_FLAG_DEFER;
  cudaEvent_t done = c->ev_call[c->ncalls % kCallRing];  // the completion event launch_render records for this call
  if (int rc = launch_render(c, cam, &f, c->d_accum.p, rgb8_out ? c->d_rgb8.p : nullptr, c->stream, nullptr, true)) return rc;
  c->wait_copy = false;  // spp == 0: nothing consumed the upload; the download below follows it on the copy stream anyway
  CU(cudaStreamWaitEvent(c->copy, done, 0));
  if (int rc = copy_rows(fp, accum_out, c->d_accum.p, sizeof(float4), cudaMemcpyDeviceToHost, c->copy)) return rc;
  if (rgb8_out)
    if (int rc = copy_rows(fp, rgb8_out, c->d_rgb8.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->copy)) return rc;
  CU(cudaEventRecord(c->ev_ticket[slot], c->copy));
  *ticket = ++c->nticket;
  return 0;
}

int single_render_wait(vp_ctx* c, unsigned long long ticket)
{
  if (ticket == 0 || ticket > c->nticket) return fail(VP_EINVAL, "unknown ticket %llu (last issued: %llu)", ticket, c->nticket);
  if (c->nticket - ticket >= (unsigned long long)kTickets) return 0;  // its ring slot was waited for before it was reused
  if (bind(c)) return VP_ECUDA;
  CU(cudaEventSynchronize(c->ev_ticket[(ticket - 1) % kTickets]));
  return 0;
}

int read_counters(vp_ctx* c, unsigned long long h[8])
{
  if (bind(c) || quiesce(c)) return VP_ECUDA;
  CU(cudaMemcpy(h, c->d_counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- multi-device context -----------------------------------------------------------------------------------------
constexpr int kGroupRowTile = 8;  // tile = 8 rows, tile k -> GPU k mod G

int g_shard(const vp_ctx* g, const VpFrameParams* fp, int d, VpFrameParams& f)
{
  if (fp->row_tile > 0 && fp->row_tile_stride > 1)
    return fail(VP_EINVAL, "a multi-device context shards the rows itself: leave row_tile / row_tile_stride unset");
  f = *fp;
  f.row_tile = kGroupRowTile; f.row_tile_stride = (int)g->kids.size(); f.row_tile_phase = d;
  return 0;
}

void worker_main(KidWorker* w)
{
  for (;;) {
    std::function<int()> job;
    {
      std::unique_lock<std::mutex> lk(w->mu);
      w->cv.wait(lk, [&] { return w->has_job || w->quit; });
      if (w->quit) return;
      job = std::move(w->job);
      w->has_job = false;
    }
    const int rc = job();
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->rc = rc;
      w->err = rc ? g_err : std::string();  // the message lives in the worker's thread-local slot: hand it over
      w->done = true;
    }
    w->cv.notify_all();
  }
}

// run one job per device (on the device's worker thread when the group has them) and wait for all of them
int g_parallel(vp_ctx* g, const std::function<int(size_t)>& per_kid)
{
  if (g->workers.empty()) {
    for (size_t d = 0; d < g->kids.size(); d++)
      if (int rc = per_kid(d)) return rc;
    return 0;
  }
  for (size_t d = 0; d < g->kids.size(); d++) {
    KidWorker* w = g->workers[d];
    {
      std::lock_guard<std::mutex> lk(w->mu);
      w->job = [&per_kid, d] { return per_kid(d); };
      w->has_job = true; w->done = false;
    }
    w->cv.notify_all();
  }
  int rc = 0;
  for (KidWorker* w : g->workers) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if (w->rc && !rc) { rc = w->rc; g_err = w->err; }
  }
  return rc;
}

int g_render_async(vp_ctx* g, const VpCamera* cam, const VpFrameParams* fp, const float* accum_in, float* accum_out, uint32_t* rgb8_out,
                   unsigned long long* ticket)
{
  const int slot = (int)(g->nticket % kTickets);
  VpFrameParams probe;
  if (int rc = g_shard(g, fp, 0, probe)) return rc;
  const int rc = g_parallel(g, [&](size_t d) -> int {
    VpFrameParams f;
    if (int r = g_shard(g, fp, (int)d, f)) return r;
    return single_render_async(g->kids[d], cam, &f, accum_in, accum_out, rgb8_out, &g->gtickets[slot][d]);
  });
  if (rc) return rc;
  *ticket = ++g->nticket;
  return 0;
}

int g_render_wait(vp_ctx* g, unsigned long long ticket)
{
  if (ticket == 0 || ticket > g->nticket) return fail(VP_EINVAL, "unknown ticket %llu (last issued: %llu)", ticket, g->nticket);
  if (g->nticket - ticket >= (unsigned long long)kTickets) return 0;
  const int slot = (int)((ticket - 1) % kTickets);
  for (size_t d = 0; d < g->kids.size(); d++)
    if (int rc = single_render_wait(g->kids[d], g->gtickets[slot][d])) return rc;
  return 0;
}

void fill_stats(VpStats* st, const unsigned long long a[8], const unsigned long long b[8])
{
  st->primary_rays = a[4] - b[4];
  st->secondary_rays = (a[0] - b[0]) - (a[4] - b[4]);
  st->shadow_rays = a[1] - b[1];
  st->exit_rays = a[2] - b[2];
  st->dda_steps = a[3] - b[3];
  st->sky_samples = a[6] - b[6];
}

int g_counters(vp_ctx* g, unsigned long long sum[8], unsigned long long* launches)
{
  memset(sum, 0, 8 * sizeof(unsigned long long));
  *launches = 0;
  for (vp_ctx* k : g->kids) {
    unsigned long long h[8];
    if (int rc = read_counters(k, h)) return rc;
    for (int i = 0; i < 8; i++) sum[i] += h[i];
    *launches += k->host_launches;
  }
  return 0;
}

// Update() with host buffers on every GPU of the group: each device uploads / renders / downloads its own row tiles, all
// of them into the caller's one accumulator and one image.
int g_render(vp_ctx* g, const VpCamera* cam, const VpFrameParams* fp, float* accum, uint32_t* rgb8, VpStats* stats)
{
  unsigned long long before[8], after[8], l0 = 0, l1 = 0;
  if (stats)
    if (int rc = g_counters(g, before, &l0)) return rc;
  for (vp_ctx* k : g->kids) {  // device time: first upload .. last download, the slowest device
    if (bind(k)) return VP_ECUDA;
    CU(cudaEventRecord(k->ev0, k->copy));
  }
  unsigned long long t = 0;
  if (int rc = g_render_async(g, cam, fp, accum, accum, rgb8, &t)) return rc;
  for (vp_ctx* k : g->kids) {
    if (bind(k)) return VP_ECUDA;
    CU(cudaEventRecord(k->ev1, k->copy));
  }
  if (int rc = g_render_wait(g, t)) return rc;
  if (stats) {
    if (int rc = g_counters(g, after, &l1)) return rc;
    fill_stats(stats, after, before);
    stats->kernel_launches = l1 - l0;
    stats->device_ms = 0;
    for (vp_ctx* k : g->kids) {
      float ms = 0;
      if (bind(k)) return VP_ECUDA;
      CU(cudaEventElapsedTime(&ms, k->ev0, k->ev1));
      if (ms > stats->device_ms) stats->device_ms = ms;
    }
  }
  return 0;
}

// Update() on device buffers that live on the group's first GPU: the other GPUs read and write the accumulator of their
// tiles and store their RGB8 pixels through peer-mapped pointers, i.e. k_wf_resolve itself is the exchange step.
int g_render_device(vp_ctx* g, const VpCamera* cam, const VpFrameParams* fp, void* accum_dev, void* rgb8_dev, cudaStream_t stream, VpStats* stats)
{
  if (!g->peer_ok) return fail(VP_ESTATE, "the devices of this context cannot access each other's memory: use the host-buffer calls");
  unsigned long long before[8], after[8], l0 = 0, l1 = 0;
  if (stats)
    if (int rc = g_counters(g, before, &l0)) return rc;
  vp_ctx* k0 = g->kids[0];
  if (bind(k0)) return VP_ECUDA;
  if (stats) CU(cudaEventRecord(k0->ev0, stream));
  CU(cudaEventRecord(g->gev_in, stream));
  const bool defer = !stats && (fp->flags & VP_FLAG_DEFER) != 0;
  cudaEvent_t done[16];
  VpFrameParams probe;
  if (int rc = g_shard(g, fp, 0, probe)) return rc;
  if (int rc = g_parallel(g, [&](size_t d) -> int {
        vp_ctx* k = g->kids[d];
        VpFrameParams f;
        if (int r = g_shard(g, fp, (int)d, f)) return r;
        f.flags |= VP_FLAG_DEFER;
        if (bind(k)) return VP_ECUDA;
        CU(cudaStreamWaitEvent(k->stream, g->gev_in, 0));  // the device's passes start after what the caller's stream held
        done[d] = k->ev_call[k->ncalls % kCallRing];
        return launch_render(k, cam, &f, (float4*)accum_dev, (uint32_t*)rgb8_dev, k->stream, nullptr, true);
      }))
    return rc;
  if (bind(k0)) return VP_ECUDA;
  if (!defer)
    for (size_t d = 0; d < g->kids.size(); d++) CU(cudaStreamWaitEvent(stream, done[d], 0));
  if (stats) {
    CU(cudaEventRecord(k0->ev1, stream));
    CU(cudaStreamSynchronize(stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, k0->ev0, k0->ev1));
    if (int rc = g_counters(g, after, &l1)) return rc;
    fill_stats(stats, after, before);
    stats->device_ms = ms;
    stats->kernel_launches = l1 - l0;
  }
  return 0;
}

int g_render_fence(vp_ctx* g, cudaStream_t stream, int frames_back)
{
  if (bind(g->kids[0])) return VP_ECUDA;
  for (vp_ctx* k : g->kids) {
    if ((unsigned long long)frames_back >= k->ncalls) continue;
    CU(cudaStreamWaitEvent(stream, k->ev_call[(k->ncalls - 1 - (unsigned long long)frames_back) % kCallRing], 0));
  }
  return 0;
}

int g_stats_end(vp_ctx* g, VpStats* stats)
{
  unsigned long long sum[8], zero[8] = {0}, launches = 0;
  if (int rc = g_counters(g, sum, &launches)) return rc;
  fill_stats(stats, sum, zero);
  stats->device_ms = 0;
  stats->kernel_launches = launches;
  return 0;
}

// ---- images shared between processes ---------------------------------------------------------------------------------
constexpr int kShareFlags = 64;          // arrival / release words, one 64-byte line each
constexpr size_t kShareHeader = 4096;    // flags live in front of the payload
__global__ void k_share_signal(unsigned int* flag, unsigned int value)
{
  // everything the stream held before this launch has completed; publish at system scope (the flag may live in a peer GPU's HBM)
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__global__ void k_share_wait(const unsigned int* flags, int first, int n, unsigned int value, unsigned long long timeout_ns, unsigned int* error)
{
  const int i = threadIdx.x;
  if (i >= n) return;
  const unsigned int* f = flags + (size_t)(first + i) * 16;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if ((int)(v - value) >= 0) return;  // wrap-safe >=
    __nanosleep(256);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) { atomicExch(error, 1u + (unsigned int)(first + i)); return; }  // a peer never arrived: report, do not hang the stream
  }
}

// ---- display interop -------------------------------------------------------------------------------
// The template uploads screen->pixels into an RGBA8 texture with format GL_BGRA (reference template/opengl.cpp:144-150,
// template/template.cpp:305-309): byte order B, G, R, x of every 0x00RRGGBB pixel.  Here the library's RGB8 image is
// written into the texture's CUDA array by a kernel (surface store, channels reordered to R, G, B, A = 255): the frame
// never leaves the GPU.
__global__ void k_to_gl(cudaSurfaceObject_t surf, const uint32_t* __restrict__ rgb8, int W, int H)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const uint32_t v = rgb8[(size_t)y * W + x];
  surf2Dwrite(make_uchar4((unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v, 255), surf, x * 4, y);
}

}  // namespace

// cuda_gl_interop.h needs <GL/gl.h>, which a headless build box does not have; the two entry points only take GL names
// (GLuint image, GLenum target) and are part of the CUDA runtime
extern "C" cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource** resource, unsigned int image, unsigned int target, unsigned int flags);

struct vp_share {
  vp_ctx* ctx = nullptr;
  unsigned char* base = nullptr;  // header + payload, in the owner's HBM
  size_t bytes = 0;
  bool owner = false;
  unsigned int* d_error = nullptr;  // local word the wait kernel sets when it gives up
};

extern "C" {

vp_ctx* vp_create_multi(const int* devices, int ndev)
{
  if (!devices || ndev < 1 || ndev > 16) { fail(VP_EINVAL, "vp_create_multi: 1..16 devices"); return nullptr; }
  if (ndev == 1) return vp_create(devices[0]);
  // test hook: several contexts on ONE device exercise the sharding, the copies and the peer-pointer path on a one-GPU box
  const bool dup_ok = getenv("VP_ALLOW_DUPLICATE_DEVICES") && atoi(getenv("VP_ALLOW_DUPLICATE_DEVICES")) != 0;
  for (int i = 0; i < ndev && !dup_ok; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) { fail(VP_EINVAL, "vp_create_multi: device %d listed twice", devices[i]); return nullptr; }
  vp_ctx* g = new vp_ctx;
  g->group = true;
  g->device = devices[0];
  for (int i = 0; i < ndev; i++) {
    vp_ctx* k = vp_create(devices[i]);
    if (!k) { vp_destroy(g); return nullptr; }
    g->kids.push_back(k);
  }
  g->peer_ok = true;
  for (int i = 1; i < ndev; i++) {
    if (devices[i] == devices[0]) continue;  // the same memory
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) != cudaSuccess || !can) { g->peer_ok = false; continue; }
    cudaSetDevice(devices[i]);
    const cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->peer_ok = false;
    cudaGetLastError();
  }
  bool distinct = true;  // contexts that share a device (the test hook) are issued from one thread: a graph capture on one of
  for (int i = 0; i < ndev; i++)  // them makes device-wide calls of the other fail
    for (int j = 0; j < i; j++) distinct = distinct && devices[i] != devices[j];
  if (distinct && !(getenv("VP_GROUP_THREADS") && atoi(getenv("VP_GROUP_THREADS")) == 0)) {
    for (int i = 0; i < ndev; i++) {
      KidWorker* w = new KidWorker;
      w->th = std::thread(worker_main, w);
      g->workers.push_back(w);
    }
  }
  cudaSetDevice(devices[0]);
  if (cudaEventCreateWithFlags(&g->gev_in, cudaEventDisableTiming) != cudaSuccess) {
    fail(VP_ENODEVICE, "vp_create_multi: event creation failed");
    vp_destroy(g);
    return nullptr;
  }
  return g;
}

int vp_device_count(vp_ctx* c) { return !c ? 0 : (c->group ? (int)c->kids.size() : 1); }

int vp_render_async(vp_ctx* c, const VpCamera* cam, const VpFrameParams* fp, const float* accum_in, float* accum_out, uint32_t* rgb8_out,
                    uint64_t* ticket)
{
  if (int rc = check_frame(c, cam, fp)) return rc;
  if (!accum_out || !ticket) return fail(VP_EINVAL, "null accumulator / ticket");
  unsigned long long t = 0;
  const int rc = c->group ? g_render_async(c, cam, fp, accum_in, accum_out, rgb8_out, &t) : single_render_async(c, cam, fp, accum_in, accum_out, rgb8_out, &t);
  *ticket = t;
  return rc;
}

int vp_render_wait(vp_ctx* c, uint64_t ticket)
{
  if (!c) return fail(VP_EINVAL, "null context");
  return c->group ? g_render_wait(c, ticket) : single_render_wait(c, ticket);
}

int vp_host_register(vp_ctx* c, void* ptr, uint64_t bytes)
{
  if (!c || !ptr || !bytes) return fail(VP_EINVAL, "bad arguments to vp_host_register");
  if (bind(c->group ? c->kids[0] : c)) return VP_ECUDA;
  CU(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  return 0;
}

int vp_host_unregister(vp_ctx* c, void* ptr)
{
  if (!c || !ptr) return fail(VP_EINVAL, "bad arguments to vp_host_unregister");
  if (bind(c->group ? c->kids[0] : c)) return VP_ECUDA;
  if (c->group) {
    for (vp_ctx* k : c->kids)
      if (bind(k) || quiesce(k)) return VP_ECUDA;
  } else if (quiesce(c)) return VP_ECUDA;
  CU(cudaHostUnregister(ptr));
  return 0;
}

int vp_render_to_gl(vp_ctx* c, uint32_t gl_texture)
{
  if (!c) return fail(VP_EINVAL, "null context");
  vp_ctx* k = c->group ? c->kids[0] : c;
  if (k->acc_w <= 0 || !k->d_rgb8.p) return fail(VP_ESTATE, "no image yet: render a frame with vp_render / vp_render_async / vp_render_resident first");
  if (c->group) {
    for (vp_ctx* q : c->kids)
      if (bind(q) || quiesce(q)) return VP_ECUDA;
  }
  if (bind(k) || quiesce(k)) return VP_ECUDA;
  if (k->gl_res && k->gl_tex != gl_texture) { cudaGraphicsUnregisterResource(k->gl_res); k->gl_res = nullptr; }
  if (!k->gl_res) {
    const unsigned int kGlTexture2D = 0x0DE1;  // GL_TEXTURE_2D
    const cudaError_t e = cudaGraphicsGLRegisterImage(&k->gl_res, gl_texture, kGlTexture2D, cudaGraphicsRegisterFlagsSurfaceLoadStore | cudaGraphicsRegisterFlagsWriteDiscard);
    if (e != cudaSuccess) {
      k->gl_res = nullptr;
      cudaGetLastError();
      return fail(VP_ECUDA, "cudaGraphicsGLRegisterImage(texture %u) failed: %s (an OpenGL context that owns the texture must be current on this thread)",
                  gl_texture, cudaGetErrorString(e));
    }
    k->gl_tex = gl_texture;
  }
  CU(cudaGraphicsMapResources(1, &k->gl_res, k->stream));
  cudaArray_t arr = nullptr;
  cudaError_t e = cudaGraphicsSubResourceGetMappedArray(&arr, k->gl_res, 0, 0);
  cudaSurfaceObject_t surf = 0;
  if (e == cudaSuccess) {
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    e = cudaCreateSurfaceObject(&surf, &rd);
  }
  if (e == cudaSuccess) {
    const dim3 blk(32, 8), grd((unsigned)(k->acc_w + 31) / 32, (unsigned)(k->acc_h + 7) / 8);
    k_to_gl<<<grd, blk, 0, k->stream>>>(surf, k->d_rgb8.p, k->acc_w, k->acc_h);
    e = cudaGetLastError();
  }
  cudaGraphicsUnmapResources(1, &k->gl_res, k->stream);
  const cudaError_t es = cudaStreamSynchronize(k->stream);
  if (surf) cudaDestroySurfaceObject(surf);
  if (e != cudaSuccess || es != cudaSuccess) return fail(VP_ECUDA, "vp_render_to_gl: %s", cudaGetErrorString(e != cudaSuccess ? e : es));
  return 0;
}

vp_share* vp_share_create(vp_ctx* c, uint64_t bytes, VpShareInfo* out)
{
  if (!c || !out || !bytes) { fail(VP_EINVAL, "bad arguments to vp_share_create"); return nullptr; }
  vp_ctx* k = c->group ? c->kids[0] : c;
  if (bind(k)) return nullptr;
  vp_share* s = new vp_share;
  s->ctx = k; s->owner = true; s->bytes = (size_t)bytes;
  cudaIpcMemHandle_t h;
  if (cudaMalloc(&s->base, kShareHeader + (size_t)bytes) != cudaSuccess || cudaMemset(s->base, 0, kShareHeader + (size_t)bytes) != cudaSuccess ||
      cudaMalloc(&s->d_error, sizeof(unsigned int)) != cudaSuccess || cudaMemset(s->d_error, 0, sizeof(unsigned int)) != cudaSuccess ||
      cudaIpcGetMemHandle(&h, s->base) != cudaSuccess) {
    fail(VP_ECUDA, "vp_share_create: %s", cudaGetErrorString(cudaGetLastError()));
    if (s->base) cudaFree(s->base);
    if (s->d_error) cudaFree(s->d_error);
    delete s;
    return nullptr;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(out->handle), "handle size");
  memset(out, 0, sizeof(*out));
  memcpy(out->handle, &h, sizeof(h));
  out->bytes = bytes;
  out->device = k->device;
  return s;
}

vp_share* vp_share_open(vp_ctx* c, const VpShareInfo* info)
{
  if (!c || !info || !info->bytes) { fail(VP_EINVAL, "bad arguments to vp_share_open"); return nullptr; }
  vp_ctx* k = c->group ? c->kids[0] : c;
  if (bind(k)) return nullptr;
  vp_share* s = new vp_share;
  s->ctx = k; s->owner = false; s->bytes = (size_t)info->bytes;
  cudaIpcMemHandle_t h;
  memcpy(&h, info->handle, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess || cudaMalloc(&s->d_error, sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(s->d_error, 0, sizeof(unsigned int)) != cudaSuccess) {
    fail(VP_ECUDA, "vp_share_open: %s (peer access between this device and device %d is required)", cudaGetErrorString(cudaGetLastError()), info->device);
    if (p) cudaIpcCloseMemHandle(p);
    if (s->d_error) cudaFree(s->d_error);
    delete s;
    return nullptr;
  }
  s->base = (unsigned char*)p;
  return s;
}

void* vp_share_ptr(vp_share* s) { return s ? (void*)(s->base + kShareHeader) : nullptr; }

int vp_share_signal(vp_share* s, void* stream, int slot, uint32_t value)
{
  if (!s || slot < 0 || slot >= kShareFlags) return fail(VP_EINVAL, "bad arguments to vp_share_signal");
  if (bind(s->ctx)) return VP_ECUDA;
  k_share_signal<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned int*>(s->base) + (size_t)slot * 16, value);
  CU(cudaGetLastError());
  return 0;
}

int vp_share_wait(vp_share* s, void* stream, int first_slot, int nslots, uint32_t value, double timeout_s)
{
  if (!s || first_slot < 0 || nslots < 1 || nslots > 32 || first_slot + nslots > kShareFlags) return fail(VP_EINVAL, "bad arguments to vp_share_wait");
  if (bind(s->ctx)) return VP_ECUDA;
  const unsigned long long ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 5.0) * 1e9);
  k_share_wait<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned int*>(s->base), first_slot, nslots, value, ns, s->d_error);
  CU(cudaGetLastError());
  return 0;
}

int vp_share_error(vp_share* s, uint32_t* error)
{
  if (!s || !error) return fail(VP_EINVAL, "bad arguments to vp_share_error");
  if (bind(s->ctx)) return VP_ECUDA;
  CU(cudaMemcpy(error, s->d_error, sizeof(unsigned int), cudaMemcpyDeviceToHost));  // synchronises with the default stream only: call after the streams were synchronised
  return 0;
}

int vp_share_close(vp_share* s)
{
  if (!s) return 0;
  if (bind(s->ctx)) return VP_ECUDA;
  cudaDeviceSynchronize();
  if (s->owner) cudaFree(s->base);
  else cudaIpcCloseMemHandle(s->base);
  if (s->d_error) cudaFree(s->d_error);
  delete s;
  return 0;
}

}  // extern "C"
