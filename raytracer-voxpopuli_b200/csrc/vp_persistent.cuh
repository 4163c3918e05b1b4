// vp_persistent.cuh - the wavefront stages as ONE persistent kernel with device-side queues.
//
// The staged pipeline (vp_wavefront.cuh) ends every stage of every bounce with a kernel
// boundary: 2 x 15 tails per frame in which a few long rays keep the whole GPU waiting
// (~1 ms of a 6.5 ms frame on room.vox, and the part of a frame that does not shrink when the
// frame is sharded over more GPUs).  Here the same stage code runs inside one launch:
//
//   three append-only queues in HBM (extend / shade / shadow), each a plain array of ids with
//   a reserved-tail and a claimed-head counter; a slot is published by storing the id (empty
//   slots hold 0xFFFFFFFF), so no per-batch commit protocol is needed;
//   every warp loops: pick the stage with work (shade when a full warp of hits is waiting,
//   else extend, else shade leftovers, else shadow), run a session of it, pick again;
//   extend / shadow sessions keep their lanes in the DDA loop and refill idle lanes from the
//   warp's private 32-slot window of the queue (one atomicAdd per 32 ids, contention-free polls);
//   `pending` counts paths + shadow requests that still need work; a warp that finds all
//   queues empty leaves when it reaches zero.
//
// Path state crosses SMs inside one launch, so it is read with ld.global.cg / written with
// st.global.cg (L2-coherent, no stale L1 lines) and a __threadfence() separates a producer's
// state writes from the id it publishes.  No warp ever blocks on another one; a warp that stays
// idle for ~1 s raises the error word and leaves, so a logic error cannot hang the device.
#pragma once
#include "vp_wavefront.cuh"

namespace vp {

constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;
// control block layout (32-bit words, one 128-byte line per hot counter)
enum : uint32_t { kQeTail = 0, kQeHead = 32, kQsTail = 64, kQsHead = 96, kQdTail = 128, kQdHead = 160, kPending = 192, kError = 224,
                  kPrevTails = 256, kCtlWords = 320 };

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
// release store: MEMBAR.ALL.GPU + ST.STRONG.GPU - orders the lane's earlier writes before the id without the L1
// invalidation a __threadfence() carries (CCTL.IVALL would throw away the cached occupancy masks every time)
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// A warp consumes a queue through a private window of 32 consecutive slots: one atomicAdd on the
// queue's head hands out a window, after that the warp only polls its own 32 slots (one coalesced
// volatile load, no contention) and takes the ids that have been published.  Nothing ever waits
// for a slot: unfilled tickets are simply looked at again later.
struct Win { uint32_t base; uint32_t consumed; };
__device__ __forceinline__ unsigned win_poll(const uint32_t* q, uint32_t* ctl, uint32_t head_w, Win& w, uint32_t& myval)
{
  const int lane = threadIdx.x & 31;
  if (w.consumed == 0xFFFFFFFFu) {
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(ctl + head_w, 32u);
    w.base = __shfl_sync(0xFFFFFFFFu, b, 0);
    w.consumed = 0;
  }
  myval = ld_volatile(q + w.base + (uint32_t)lane);
  return __ballot_sync(0xFFFFFFFFu, myval != kEmptySlot) & ~w.consumed;
}
// hand the first ids of `ready` to the lanes of `needers` (in lane order); returns the taken tickets
__device__ __forceinline__ unsigned win_take(unsigned ready, unsigned needers, uint32_t myval, bool& got, uint32_t& id)
{
  const int lane = threadIdx.x & 31;
  const int k = min(__popc(ready), __popc(needers));
  got = false;
  if (k == 0) return 0u;
  const int rank = __popc(needers & ((1u << lane) - 1u));
  const bool mine = ((needers >> lane) & 1u) && rank < k;
  const unsigned src = __fns(ready, 0, (mine ? rank : 0) + 1);
  const uint32_t v = __shfl_sync(0xFFFFFFFFu, myval, (int)(src & 31u));
  if (mine) { got = true; id = v; }
  const unsigned last = __fns(ready, 0, k);  // position of the k-th ready ticket
  return ready & (last >= 31u ? 0xFFFFFFFFu : ((2u << last) - 1u));
}

// all 32 lanes converged: append the ids of the lanes that `want`
__device__ __forceinline__ void pq_push(uint32_t* q, uint32_t* ctl, uint32_t tail_w, bool want, uint32_t id)
{
  const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(ctl + tail_w, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (want) st_release(q + base + (uint32_t)__popc(m & ((1u << lane) - 1u)), id);  // state writes of this lane first, then the id
}

// clear what the previous pass appended (the queues are all-empty again afterwards)
__global__ void __launch_bounds__(256) k_wf_pclear(const __grid_constant__ WfParams p)
{
  for (int qi = 0; qi < 3; qi++) {
    const uint32_t n = p.st.pctl[kPrevTails + qi];
    uint32_t* q = p.st.pq[qi];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) q[i] = kEmptySlot;
  }
}

// ---- extend session: Renderer::FindNearest for ids claimed from the extend queue ---------------------
template <bool CS>
__device__ __forceinline__ void p_extend_session(const WfParams& p, const Staged& stg, Win& win, uint32_t* ring, unsigned long long& steps,
                                                 unsigned int& nrays)
{
  uint32_t ring_w = 0, ring_n = 0;  // finished paths waiting in the warp's ring: entries [ring_w - ring_n, ring_w) mod 64
  const DVolume* svol = stg.vol;
  uint32_t* ctl = p.st.pctl;
  const int lane = threadIdx.x & 31;
  const int nvol = p.sc.nvol;
  bool has_ray = false, in_dda = false, dry = false;
  uint32_t path = 0;
  f3 O = v3s(0), D = v3s(0);
  float tbest = 0; uint32_t pbest = 0; int volbest = -2;
  int vi = 0;
  const unsigned long long* region = nullptr;
  const unsigned long long* fine = nullptr;
  DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = 0; s.t = 3.0e38f; s.P = s.R = 0; s.Ax = s.Ay = s.Az = s.Sx = s.Sy = s.Sz = 0;
  Occ o; occ_reset(o);
  for (;;) {
    // 1. refill idle lanes from the warp's window of the extend queue
    unsigned nm = __ballot_sync(0xFFFFFFFFu, !has_ray);
    if (nm) {
      bool any_taken = false, taken = false;
      for (int rep = 0; rep < 2 && nm; rep++) {
        uint32_t myval;
        const unsigned ready = win_poll(p.st.pq[0], ctl, kQeHead, win, myval);
        bool got; uint32_t id = 0;
        const unsigned tk = win_take(ready, nm, myval, got, id);
        win.consumed |= tk;
        if (got) { path = id; taken = true; }
        any_taken |= tk != 0u;
        nm &= ~__ballot_sync(0xFFFFFFFFu, got);
        if (win.consumed != 0xFFFFFFFFu) break;  // window not used up: the rest is not published yet
      }
      dry = nm != 0u;
      (void)any_taken;  // the state address depends on the id just read: no acquire fence needed
      if (taken) {
        const float4 a = __ldcg(p.st.r0 + path), b = __ldcg(p.st.r1 + path);
        O = v3(a.x, a.y, a.z); D = v3(b.x, b.y, b.z);
        tbest = 1e34f; pbest = 0; volbest = -2;
        vi = 0; has_ray = true; in_dda = false;
        nrays++;
      }
    }
    if (!__any_sync(0xFFFFFFFFu, has_ray)) break;  // nothing claimed and nothing walking: back to the scheduler
    // 2. lanes between volumes: enter the next volume or finish the ray
    bool finished = false;
    uint32_t done_path = 0;
    if (has_ray && !in_dda) {
      while (vi < nvol) {
        const DVolume& v = svol[vi];
        if (nvol > 1 && !volume_may_hit(v, O, D)) { vi++; continue; }
        const ORay r = make_oray(xform_pos_sse(O, v.iM), xform_vec_sse(D, v.iM));
        vi++;
        if (setup_dda(v, r, s)) { in_dda = true; occ_reset(o); region = region_of(stg, v); fine = v.fine; break; }
      }
      if (!in_dda) {
        uint32_t hb = (uint32_t)(volbest + 2) << 8;
        if (p.sc.nsph + p.sc.ntri > 0) {
          PrimHit h; h.t = 1e34f; h.N = v3s(0); h.mat = VP_MAT_NONE; h.inside = 0;
          prims_nearest(p.sc, O, normalize3(D), h);
          if (tbest > h.t) {
            tbest = h.t;
            hb = (uint32_t)(h.mat & 255) | (1u << 8) | ((uint32_t)(h.inside & 1) << 30) | (1u << 31);
            __stcg(p.st.hn + path, make_float4(h.N.x, h.N.y, h.N.z, 0));
          }
        }
        __stcg(p.st.h0 + path, make_float2(tbest, __uint_as_float(pbest)));
        __stcg(p.st.h1 + path, hb);
        has_ray = false; finished = true; done_path = path;
      }
    }
    // hit records go to the shade queue in batches of 32 (one release fence per batch); when the extend
    // queue runs dry they are published at once so that shading is never held back
    {
      const unsigned fm = __ballot_sync(0xFFFFFFFFu, finished);
      if (finished) ring[(ring_w + (uint32_t)__popc(fm & ((1u << lane) - 1u))) & 63u] = done_path;
      ring_w += (uint32_t)__popc(fm); ring_n += (uint32_t)__popc(fm);
      __syncwarp();
      while (ring_n >= 32u || (dry && ring_n > 0u)) {
        const uint32_t k = min(ring_n, 32u);
        const uint32_t id = ring[(ring_w - ring_n + (uint32_t)lane) & 63u];
        pq_push(p.st.pq[1], ctl, kQsTail, (uint32_t)lane < k, id);
        ring_n -= k;
      }
      __syncwarp();
    }
    // 3. DDA stepping
    if (!in_dda) s.t = 3.0e38f;
    const unsigned thr = dry ? 1u : (unsigned)p.refill;
    const bool drained = dry;
    bool sparse = drained || (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tbest)) <= (unsigned)p.bulk;
    for (;;) {
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (s.t < tbest) {
          if (CS) steps++;
          uint32_t skipped = 0;
          const bool solid = occ_visit(region, fine, o, s, tbest, sparse, skipped);
          if (CS) steps += skipped;
          if (solid) { tbest = s.t; s.t = __uint_as_float(kHitTag); }  // NaN tag: stops the walk, recovered below
          else if (!dda_step(s)) s.t = 3.0e38f;
        }
      }
      const unsigned act = (unsigned)__popc(__ballot_sync(0xFFFFFFFFu, s.t < tbest));
      sparse = drained || act <= (unsigned)p.bulk;
      if (act < thr) break;
    }
    if (__float_as_uint(s.t) == kHitTag) { pbest = s.P; volbest = vi - 1; s.t = 3.0e38f; }  // the cell the walk stopped on
    in_dda = s.t < tbest;
  }
  while (ring_n > 0u) {  // leaving the session: publish what is still in the ring
    const uint32_t k = min(ring_n, 32u);
    const uint32_t id = ring[(ring_w - ring_n + (uint32_t)lane) & 63u];
    pq_push(p.st.pq[1], ctl, kQsTail, (uint32_t)lane < k, id);
    ring_n -= k;
  }
  __syncwarp();
}

// ---- shadow session: Renderer::IsOccluded for requests claimed from the shadow queue -----------------
template <bool CS>
__device__ __forceinline__ void p_shadow_session(const WfParams& p, const Staged& stg, Win& win, unsigned long long& steps, unsigned int& nrays)
{
  const DVolume* svol = stg.vol;
  const DScene& sc = p.sc;
  uint32_t* ctl = p.st.pctl;
  const int lane = threadIdx.x & 31;
  const uint32_t cap = p.st.cap;
  const int nvol = sc.nvol;
  const float lightCount = (float)(uint32_t)(sc.npl + sc.nsl + sc.nal + 1);
  bool has_req = false, has_ray = false, in_dda = false, dry = false;
  uint32_t ndone = 0;
  uint32_t path = 0, level = 0, bits = 0;
  int nsamp = 0, sample = 0, vi = 0;
  f3 org = v3s(0), D = v3s(0), inc = v3s(0), li = v3s(0);
  float tmaxray = 0;
  const unsigned long long* region = nullptr;
  const unsigned long long* fine = nullptr;
  DDA s; s.tx = s.ty = s.tz = s.dx = s.dy = s.dz = 0; s.t = 3.0e38f; s.P = s.R = 0; s.Ax = s.Ay = s.Az = s.Sx = s.Sy = s.Sz = 0;
  Occ o; occ_reset(o);
  for (;;) {
    unsigned nm = __ballot_sync(0xFFFFFFFFu, !has_req);
    if (nm) {
      bool any_taken = false, taken = false;
      uint32_t rid = kEmptySlot;
      for (int rep = 0; rep < 2 && nm; rep++) {
        uint32_t myval;
        const unsigned ready = win_poll(p.st.pq[2], ctl, kQdHead, win, myval);
        bool got; uint32_t id = 0;
        const unsigned tk = win_take(ready, nm, myval, got, id);
        win.consumed |= tk;
        if (got) { rid = id; taken = true; }
        any_taken |= tk != 0u;
        nm &= ~__ballot_sync(0xFFFFFFFFu, got);
        if (win.consumed != 0xFFFFFFFFu) break;
      }
      dry = nm != 0u;
      (void)any_taken;
      if (taken) {
        level = rid / cap; path = rid - level * cap;
        const float4 q0 = __ldcg(p.st.sq0 + (size_t)level * cap + path);
        org = v3(q0.x, q0.y, q0.z);
        bits = __float_as_uint(q0.w);
        nsamp = (int)(bits & 255u);
        sample = 0; inc = v3s(0);
        has_req = true; has_ray = false; in_dda = false;
      }
    }
    if (!__any_sync(0xFFFFFFFFu, has_req)) break;
    unsigned int completed = 0;
    if (has_req && !in_dda) {
      for (;;) {
        if (has_ray) {
          bool entered = false;
          while (vi < nvol) {
            const DVolume& v = svol[vi];
            if (nvol > 1 && !volume_may_hit(v, org, D)) { vi++; continue; }
            const ORay r = make_oray(xform_scalar(org, 1.0f, v.iM), xform_scalar(D, 0.0f, v.iM));
            vi++;
            if (setup_dda(v, r, s)) { entered = true; occ_reset(o); region = region_of(stg, v); fine = v.fine; break; }
          }
          if (entered) { in_dda = true; break; }
          bool occluded = false;
          if (sc.nsph + sc.ntri > 0) occluded = prims_any(sc, org, D, tmaxray);
          if (!occluded) inc = (bits & (1u << 16)) ? add(inc, li) : li;
          has_ray = false;
          sample++;
        }
        bool found = false;
        while (sample < nsamp) {
          const size_t si = ((size_t)level * p.st.nsamp + (size_t)sample) * cap + path;
          const float4 l4 = __ldcg(p.st.sql + si);
          if (l4.w != 0.0f) {
            const float4 d4 = __ldcg(p.st.sqd + si);
            D = normalize3(v3(d4.x, d4.y, d4.z));
            tmaxray = d4.w;
            li = v3(l4.x, l4.y, l4.z);
            vi = 0; has_ray = true; found = true;
            nrays++;
            break;
          }
          sample++;
        }
        if (!found) {
          f3 A;
          if (bits & (1u << 16)) {
            f3 t = divs(inc, (float)nsamp);
            t = mul(t, albedo_of(sc, (int)(bits >> 24)));
            A = muls(t, lightCount);
          } else A = muls(inc, lightCount);
          __stcg(p.st.la + (size_t)level * cap + path, make_float4(A.x, A.y, A.z, 0));
          has_req = false;
          completed = 1;
          break;
        }
      }
    }
    {  // requests finished no longer count as pending work (batched: one atomic per 32, or at once when the queue is dry)
      ndone += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, completed != 0));
      if (ndone >= 32u || (dry && ndone)) { if (lane == 0) atomicSub(ctl + kPending, ndone); ndone = 0; }
    }
    if (!in_dda) s.t = 3.0e38f;
    const unsigned thr = dry ? 1u : (unsigned)p.refill;
    bool occluded = false;
    const bool drained = dry;
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
    if (in_dda) {
      in_dda = s.t < tmaxray;
      if (occluded) { has_ray = false; sample++; }
    }
  }
  if (ndone && lane == 0) atomicSub(ctl + kPending, ndone);
}

// ---- shade batch: the published hits of the warp's shade-queue window ------------------------------
template <bool CS>
__device__ __forceinline__ void p_shade_batch(const WfParams& p, const Staged& stg, Win& win, unsigned ready, uint32_t myval, Counters& cn,
                                              int& in_light)
{
  uint32_t* ctl = p.st.pctl;
  const int lane = threadIdx.x & 31;
  const uint32_t cap = p.st.cap;
  win.consumed |= ready;
  const bool valid = (ready >> lane) & 1u;
  const uint32_t path = myval;
  bool cont = false, want_shadow = false;
  int level = 0;
  if (valid) wf_shade_one<CS>(p, stg, path, cn, in_light, cont, want_shadow, level);
  // a shadow request is new pending work: count it before it can be seen; a path that ends stops counting after
  const unsigned sm = __ballot_sync(0xFFFFFFFFu, want_shadow);
  const unsigned tm = __ballot_sync(0xFFFFFFFFu, valid && !cont);
  if (lane == 0 && sm) atomicAdd(ctl + kPending, (uint32_t)__popc(sm));
  pq_push(p.st.pq[0], ctl, kQeTail, cont, path);
  pq_push(p.st.pq[2], ctl, kQdTail, want_shadow, (uint32_t)level * cap + path);
  if (lane == 0 && tm) atomicSub(ctl + kPending, (uint32_t)__popc(tm));
}

template <bool CS>
__global__ void __launch_bounds__(kWfBlock) k_wf_persistent(const __grid_constant__ WfParams p)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const Staged stg = wf_stage_scene(p.sc, smem);
  __shared__ uint32_t s_ring[kWfBlock / 32][64];
  uint32_t* ring = s_ring[threadIdx.x >> 5];
  uint32_t* ctl = p.st.pctl;
  const int lane = threadIdx.x & 31;
  Counters cn; cn.nearest = cn.shadow = cn.exitray = cn.sky = 0; cn.steps = 0;
  int in_light = 0;
  unsigned long long steps = 0;
  unsigned int n_extend = 0, n_shadow = 0;
  int idle = 0;
  Win wE, wS, wD;
  wE.base = wS.base = wD.base = 0; wE.consumed = wS.consumed = wD.consumed = 0xFFFFFFFFu;
  for (;;) {
    uint32_t vS, vE, vD;
    const unsigned rS = win_poll(p.st.pq[1], ctl, kQsHead, wS, vS);
    const unsigned rE = win_poll(p.st.pq[0], ctl, kQeHead, wE, vE);
    const unsigned rD = win_poll(p.st.pq[2], ctl, kQdHead, wD, vD);
    if (__popc(rS) >= 16 || (rS && !rE)) { p_shade_batch<CS>(p, stg, wS, rS, vS, cn, in_light); idle = 0; }
    else if (rE) { p_extend_session<CS>(p, stg, wE, ring, steps, n_extend); idle = 0; }
    else if (rD) { p_shadow_session<CS>(p, stg, wD, steps, n_shadow); idle = 0; }
    else {
      int stop = 0;
      if (lane == 0) stop = ((int)ld_volatile(ctl + kPending) <= 0 || ld_volatile(ctl + kError)) ? 1 : 0;
      if (__shfl_sync(0xFFFFFFFFu, stop, 0)) break;
      __nanosleep(100);
      if (++idle > (1 << 22)) { if (lane == 0) atomicExch(ctl + kError, 1u); break; }  // ~1 s without work: give up loudly
    }
  }
  wf_add_counter(p.counters + 0, n_extend);
  wf_add_counter(p.counters + 1, (unsigned long long)n_shadow + cn.shadow);
  wf_add_counter(p.counters + 2, cn.exitray);
  wf_add_counter(p.counters + 5, (unsigned long long)in_light);
  wf_add_counter(p.counters + 6, cn.sky);
  if (CS) wf_add_counter(p.counters + 3, steps + cn.steps);
}

}  // namespace vp
