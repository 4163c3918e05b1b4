/* TEST INFRASTRUCTURE (CPU model, never linked into the product).
 *
 * Exactness model of the two-level walk the CUDA traversal kernels use (csrc/vp_device.cuh, "brick walk"):
 * the reference's Amanatides-Woo loop (template/scene.cpp:763-804) visits cells in the order obtained by
 * MERGING three per-axis sequences  tmax_a, tmax_a + d_a, (tmax_a + d_a) + d_a, ...  (sequential float adds),
 * always consuming the smallest head, ties resolved z before y before x (the nested ifs of :776-803).
 * Every element is one cell-boundary crossing; every 4th element of an axis crosses a 4x4x4-brick boundary.
 * Because the merge order is a total order on (value, axis priority), the sequence of BRICKS the walk visits is
 * the merge of the brick-crossing elements alone.  Through bricks whose occupancy mask is 0 the walk therefore
 * only needs, per axis, the value of the next brick-crossing element E_a (the same sequential adds, 4 per
 * crossing) and no per-cell compare/select; when a non-empty brick is entered at key (T, a), the deferred
 * cell steps of the other axes are caught up (elements with key < (T, a), at most 3 per axis).
 * s.t at every cell test, the hit cell and the hit t are bit-identical to the cell walk.
 *
 * bw_compare() runs both walks on the same rays and reports mismatches plus the event counts that size the
 * kernel (brick steps, cell steps, mode switches).  Used by tests/test_brickwalk_model.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float F[3], d[3], t; int X[3], dir[3]; } St;

static inline int f2i(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? (int)v : (int)0x80000000u; }
static inline float stdminf_(float a, float b) { return (b < a) ? b : a; }
static inline float stdmaxf_(float a, float b) { return (a < b) ? b : a; }

/* Setup3DDDA (template/scene.cpp:719-749) for the unit cube [0,1]^3 and an n^3 grid; returns 0 on a miss */
static int setup(const float O[3], const float D[3], int n, St* s)
{
  float rD[3]; int neg[3];
  for (int a = 0; a < 3; a++) { rD[a] = 1.0f / D[a]; uint32_t u; memcpy(&u, &D[a], 4); neg[a] = (int)(u >> 31); }
  float t = 0;
  const int inside = O[0] >= 0 && O[1] >= 0 && O[2] >= 0 && O[0] <= 1 && O[1] <= 1 && O[2] <= 1;
  if (!inside) {
    const int sx = D[0] < 0, sy = D[1] < 0, sz = D[2] < 0;
    float tmin = ((sx ? 1.0f : 0.0f) - O[0]) * rD[0], tmax = ((sx ? 0.0f : 1.0f) - O[0]) * rD[0];
    const float tminy = ((sy ? 1.0f : 0.0f) - O[1]) * rD[1], tmaxy = ((sy ? 0.0f : 1.0f) - O[1]) * rD[1];
    if (tmin > tmaxy || tminy > tmax) return 0;
    tmin = stdmaxf_(tmin, tminy); tmax = stdminf_(tmax, tmaxy);
    const float tminz = ((sz ? 1.0f : 0.0f) - O[2]) * rD[2], tmaxz = ((sz ? 0.0f : 1.0f) - O[2]) * rD[2];
    if (tmin > tmaxz || tminz > tmax) return 0;
    tmin = stdmaxf_(tmin, tminz);
    if (!(tmin > 0)) return 0;
    t = tmin;
    if (t > 1e33f) return 0;
  }
  s->t = t;
  const float gs = (float)n, cell = 1.0f / gs, tt = t + 0.00005f;
  for (int a = 0; a < 3; a++) {
    const float pos = gs * (O[a] + tt * D[a]);
    const float plane = (ceilf(pos) - (float)neg[a]) * cell;
    int X = f2i(pos); if (X > n - 1) X = n - 1; if (X < 0) X = 0;
    s->X[a] = X;
    s->dir[a] = neg[a] ? -1 : 1;
    s->d[a] = cell * (neg[a] ? -1.0f : 1.0f) * rD[a];
    s->F[a] = (plane * 1.0f - O[a]) * rD[a];
  }
  return 1;
}

/* axis choice of template/scene.cpp:776-803 */
static inline int pick(const float v[3])
{
  if (v[0] < v[1]) return (v[0] < v[2]) ? 0 : 2;
  return (v[1] < v[2]) ? 1 : 2;
}

typedef struct { int hit; float t; int X[3]; uint32_t steps; } Res;

static Res cell_walk(const uint8_t* grid, int n, St s, float tlimit)
{
  Res r; memset(&r, 0, sizeof r);
  while (s.t < tlimit) {
    r.steps++;
    if (grid[s.X[0] + n * (s.X[1] + n * s.X[2])] != 255) { r.hit = 1; r.t = s.t; memcpy(r.X, s.X, sizeof r.X); return r; }
    const int a = pick(s.F);
    s.t = s.F[a];
    s.X[a] += s.dir[a];
    if ((unsigned)s.X[a] >= (unsigned)n) break;
    s.F[a] += s.d[a];
  }
  return r;
}

typedef struct { uint64_t brick_steps, cell_steps, to_cell, to_brick, rays, ref_steps, nan_rays; } Stats;

static inline uint64_t brick_mask(const uint64_t* fine, int nb, const int X[3])
{
  return fine[(X[0] >> 2) + nb * ((X[1] >> 2) + nb * (X[2] >> 2))];
}
static inline int mask_bit(uint64_t m, const int X[3]) { return (int)((m >> ((X[0] & 3) | ((X[1] & 3) << 2) | ((X[2] & 3) << 4))) & 1u); }

/* key (v, axis b) < key (T, axis a), b != a: ties go to the higher axis index */
static inline int key_before(float v, int b, float T, int a) { return v < T || (v == T && b > a); }

static Res brick_walk(const uint8_t* grid, const uint64_t* fine, int n, St s, float tlimit, Stats* st)
{
  (void)grid;
  Res r; memset(&r, 0, sizeof r);
  const int nb = n >> 2;
  const int plain = (s.F[0] != s.F[0]) || (s.F[1] != s.F[1]) || (s.F[2] != s.F[2]);  /* NaN tmax: cell walk only */
  if (plain) st->nan_rays++;
  uint64_t m = brick_mask(fine, nb, s.X);
  float E[3];
  int brick_mode = 0;
  for (;;) {
    if (!brick_mode) {
      /* ---- cell mode: the reference loop, occupancy from the brick mask ---- */
      int left_grid = 0, changed = 0;
      while (s.t < tlimit) {
        st->cell_steps++;
        if (mask_bit(m, s.X)) { r.hit = 1; r.t = s.t; memcpy(r.X, s.X, sizeof r.X); return r; }
        const int a = pick(s.F);
        s.t = s.F[a];
        const int before = s.X[a] >> 2;
        s.X[a] += s.dir[a];
        if ((unsigned)s.X[a] >= (unsigned)n) { left_grid = 1; break; }
        s.F[a] += s.d[a];
        if ((s.X[a] >> 2) != before) {
          m = brick_mask(fine, nb, s.X);
          if (m == 0 && !plain) { changed = 1; break; }
        }
      }
      if (left_grid || !changed) return r;
      /* -> brick mode: E_a = the element of axis a that crosses the brick boundary */
      st->to_brick++;
      for (int a = 0; a < 3; a++) {
        const int c = s.X[a] & 3, k = (s.dir[a] < 0) ? c : 3 - c;  /* cell steps before the crossing one */
        float e = s.F[a];
        for (int i = 0; i < k; i++) e += s.d[a];
        E[a] = e;
      }
      brick_mode = 1;
    } else {
      /* ---- brick mode: the current brick is empty; s.t = time it was entered ---- */
      int into_solid = 0, a = 0;
      while (s.t < tlimit) {
        st->brick_steps++;
        a = pick(E);
        const float T = E[a];
        s.t = T;
        /* move to the exit face, then across it */
        s.X[a] = (s.dir[a] < 0) ? (s.X[a] & ~3) : (s.X[a] | 3);
        s.X[a] += s.dir[a];
        if ((unsigned)s.X[a] >= (unsigned)n) return r;
        s.F[a] = T + s.d[a];
        E[a] = ((s.F[a] + s.d[a]) + s.d[a]) + s.d[a];
        m = brick_mask(fine, nb, s.X);   /* X of the other axes is stale but inside the same brick column */
        if (m) { into_solid = 1; break; }
      }
      if (!into_solid) return r;
      /* -> cell mode: catch up the deferred cell steps of the other axes */
      st->to_cell++;
      for (int b = 0; b < 3; b++) {
        if (b == a) continue;
        while (key_before(s.F[b], b, s.t, a)) { s.F[b] += s.d[b]; s.X[b] += s.dir[b]; }
      }
      brick_mode = 0;
    }
  }
}


/* The kernels' formulation: ONE step body for both kinds of lanes.  A lane inside an empty brick ("brick" set) merges
 * the crossing elements E_a, any other lane the cell elements F_a; E_a is kept current by cell steps too (a cell step
 * only changes E_a when it crosses a brick boundary), so entering brick mode costs nothing.  Leaving it needs no
 * explicit catch-up either: the deferred elements of the other two axes are the smallest heads, so the ordinary cell
 * steps that follow consume them first, in merge order; such a step is SILENT (no cell test, s.t unchanged) while its
 * key lies below (s.t, axis of the last real step). */
/* occupancy bits of the cells of a brick the ray can still reach from cell c: it never moves against its direction */
static inline uint64_t forward_box(const int X[3], const int dir[3])
{
  uint64_t m = 0;
  for (int z = 0; z < 4; z++) for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
    const int ok = (dir[0] > 0 ? x >= (X[0] & 3) : x <= (X[0] & 3)) && (dir[1] > 0 ? y >= (X[1] & 3) : y <= (X[1] & 3)) &&
                   (dir[2] > 0 ? z >= (X[2] & 3) : z <= (X[2] & 3));
    if (ok) m |= 1ull << (x | (y << 2) | (z << 4));
  }
  return m;
}

static Res unified_walk(const uint64_t* fine, int n, St s, float tlimit, Stats* st, int forward)
{
  Res r; memset(&r, 0, sizeof r);
  const int nb = n >> 2;
  const int plain = (s.F[0] != s.F[0]) || (s.F[1] != s.F[1]) || (s.F[2] != s.F[2]) || (n & 3);
  float E[3] = {0, 0, 0};
  if (!plain)
    for (int a = 0; a < 3; a++) {
      const int c = s.X[a] & 3, k = (s.dir[a] < 0) ? c : 3 - c;
      float e = s.F[a];
      for (int i = 0; i < k; i++) e += s.d[a];
      E[a] = e;
    }
  int brick = 0, ac = 0, have = 0;
  float tc = NAN;  /* key (tc, ac) of the lane's last brick step: cell steps below it are silent; NaN before the first (no element is below or equal: a tmax of -inf is possible, see kTcFirst in csrc/vp_device.cuh) */
  int key[3] = {-1, -1, -1};
  uint64_t m = 0;
  while (s.t < tlimit) {
    const int bx = s.X[0] >> 2, by = s.X[1] >> 2, bz = s.X[2] >> 2;
    if (!have || bx != key[0] || by != key[1] || bz != key[2]) {
      m = brick_mask(fine, nb, s.X); key[0] = bx; key[1] = by; key[2] = bz; have = 1;
      const int was = brick;
      brick = (m == 0) && !plain;
      if (was && !brick) st->to_cell++;
      if (!was && brick) st->to_brick++;
    }
    /* forward box: nothing the ray can still reach in this brick is occupied -> the rest of the brick is empty for it.
     * X may lag behind the true position on axes with deferred steps: the box from X contains the true one. */
    if (forward && !brick && !plain && (m & forward_box(s.X, s.dir)) == 0) { brick = 1; st->to_brick++; }
    const float* v = brick ? E : s.F;
    const int a = pick(v);
    const float T = v[a];
    const int silent = !brick && (T < tc || (T == tc && a > ac));
    if (brick) st->brick_steps++; else st->cell_steps++;
    if (!brick && !silent && mask_bit(m, s.X)) { r.hit = 1; r.t = s.t; memcpy(r.X, s.X, sizeof r.X); return r; }
    if (brick) { tc = T; ac = a; }
    const int c = s.X[a] & 3, k = (s.dir[a] < 0) ? c : 3 - c;
    const int cross = brick || k == 0;
    const float f = T + s.d[a];
    const float e = ((f + s.d[a]) + s.d[a]) + s.d[a];
    if (brick) s.X[a] = (s.dir[a] < 0) ? (s.X[a] & ~3) : (s.X[a] | 3);
    s.X[a] += s.dir[a];
    if (!silent) s.t = T;
    if ((unsigned)s.X[a] >= (unsigned)n) return r;
    s.F[a] = f;
    if (cross) E[a] = e;
  }
  return r;
}

/* ---- mode 3: the kernels' step at the BIT level (csrc/vp_device.cuh: setup of the walk word, bw_init, walk_step and uw_step
 * under VP_STEP_V2) ---------------------------------------------------------------------------------------------------------
 * Cell coordinates in one Morton word P (bit 3i + a = bit i of axis a), the walk mirrored on the axes the ray descends
 * (P' = P ^ flip: every step is +1, leaving the padded grid is the carry out of the masked add), the stepped axis carried
 * as its lowest field bit (1, 2, 4; BW::ac too), field masks as products, the forward-box table indexed by the mirrored
 * cell and holding the one-hot mask of the real cell, the brick key kept in mirrored coordinates.  finem: brick masks in
 * Morton order (index P >> 6, bit P & 63), the padded grid being (16 << jr)^3. */
static uint32_t dilate3_(uint32_t v)
{
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u; v = (v | (v << 2)) & 0x09249249u;
  return v;
}
static uint32_t compact3_(uint32_t v)
{
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u; v = (v | (v >> 4)) & 0x0300F00Fu; v = (v | (v >> 8)) & 0x030000FFu; v = (v | (v >> 16)) & 0x3FFu;
  return v;
}
static uint64_t fwd_box_[512], fwd_bit_[512];
static void fwd_table_(void)  /* k_fwd_table */
{
  for (uint32_t idx = 0; idx < 512; idx++) {
    const uint32_t sg = idx >> 6;
    const uint32_t mir = ((sg & 1u) ? 9u : 0u) | ((sg & 2u) ? 18u : 0u) | ((sg & 4u) ? 36u : 0u);
    const uint32_t cell = (idx & 63u) ^ mir;
    const uint32_t cx = compact3_(cell), cy = compact3_(cell >> 1), cz = compact3_(cell >> 2);
    uint64_t m = 0;
    for (uint32_t q = 0; q < 64; q++) {
      const uint32_t x = compact3_(q), y = compact3_(q >> 1), z = compact3_(q >> 2);
      const int ok = ((sg & 1u) ? x <= cx : x >= cx) && ((sg & 2u) ? y <= cy : y >= cy) && ((sg & 4u) ? z <= cz : z >= cz);
      if (ok) m |= 1ull << q;
    }
    fwd_box_[idx] = m; fwd_bit_[idx] = 1ull << cell;
  }
}
void bw_build_fine_morton(const uint8_t* grid, int n, int jr, uint64_t* finem)
{
  memset(finem, 0, sizeof(uint64_t) * ((size_t)64 << (3 * jr)));
  for (int z = 0; z < n; z++) for (int y = 0; y < n; y++) for (int x = 0; x < n; x++)
    if (grid[x + n * (y + n * z)] != 255) {
      const uint32_t P = dilate3_((uint32_t)x) | (dilate3_((uint32_t)y) << 1) | (dilate3_((uint32_t)z) << 2);
      finem[P >> 6] |= 1ull << (P & 63u);
    }
}
static Res bit_walk(const uint64_t* finem, int jr, St s0, float tlimit, Stats* st)
{
  Res r; memset(&r, 0, sizeof r);
  static int ready = 0;
  if (!ready) { fwd_table_(); ready = 1; }
  float tx = s0.F[0], ty = s0.F[1], tz = s0.F[2];
  const float dx = s0.d[0], dy = s0.d[1], dz = s0.d[2];
  float t = s0.t;
  uint32_t P = dilate3_((uint32_t)s0.X[0]) | (dilate3_((uint32_t)s0.X[1]) << 1) | (dilate3_((uint32_t)s0.X[2]) << 2);
  /* bw_init */
  uint32_t ac = 1u; float tc = NAN;
  const uint32_t plain = (tx == tx && ty == ty && tz == tz) ? 0u : 1u;
  const uint32_t F = 0x09249249u & ((1u << (3u * (4u + (uint32_t)jr))) - 1u);
  const uint32_t flip = (s0.dir[0] < 0 ? F : 0u) | (s0.dir[1] < 0 ? F << 1 : 0u) | (s0.dir[2] < 0 ? F << 2 : 0u);
  P ^= flip;
  const uint32_t nx = ~P & 9u, ny = ~P & 18u, nz = ~P & 36u;
  float e, ex, ey, ez;
  e = tx; if (nx & 1u) e += dx; if (nx & 8u) { e += dx; e += dx; } ex = e;
  e = ty; if (ny & 2u) e += dy; if (ny & 16u) { e += dy; e += dy; } ey = e;
  e = tz; if (nz & 4u) e += dz; if (nz & 32u) { e += dz; e += dz; } ez = e;
  uint32_t key = 0xFFFFFFFFu; uint64_t m = 0;
  if (plain) st->nan_rays++;
  while (t < tlimit) {
    const float t0 = t;
    /* walk_step: the brick mask is fetched when the (mirrored) brick changes */
    if ((P ^ key) >> 6) { key = P; m = finem[(P ^ flip) >> 6]; }
    /* uw_step */
    const uint32_t idx = ((flip & 7u) << 6) | (P & 63u);
    const int brick = (m & fwd_box_[idx]) == 0ull && plain == 0u;
    const float vx = brick ? ex : tx, vy = brick ? ey : ty, vz = brick ? ez : tz;
    const int xy = vx < vy;
    const int ax = xy & (vx < vz);
    const int ay = (!xy) & (vy < vz);
    const int az = !(ax | ay);
    const float T = ax ? vx : (ay ? vy : vz);
    const uint32_t lowb = ax ? 1u : (ay ? 2u : 4u);
    const int silent = !brick & ((T < tc) | ((T == tc) & (lowb > ac)));
    const int hit = !brick & !silent & ((m & fwd_bit_[idx]) != 0ull);
    if (brick) { tc = T; ac = lowb; st->brick_steps++; } else st->cell_steps++;
    const uint32_t M9 = lowb * 9u, Fa = lowb * F;
    if (brick) P |= M9;
    const int cross = (P & M9) == M9;
    if (ax) tx = vx + dx;
    if (ay) ty = vy + dy;
    if (az) tz = vz + dz;
    if (ax && cross) ex = ((tx + dx) + dx) + dx;
    if (ay && cross) ey = ((ty + dy) + dy) + dy;
    if (az && cross) ez = ((tz + dz) + dz) + dz;
    const uint32_t q = (P | ~Fa) + lowb;
    const int out = q < lowb;
    if (!hit) P = (q & Fa) | (P & ~Fa);
    if (!silent) t = T;
    if (out) t = 3.0e38f;
    if (hit) {
      const uint32_t Pr = P ^ flip;
      r.hit = 1; r.t = t0;
      r.X[0] = (int)compact3_(Pr); r.X[1] = (int)compact3_(Pr >> 1); r.X[2] = (int)compact3_(Pr >> 2);
      return r;
    }
  }
  return r;
}

/* fine[bx + nb*(by + nb*bz)] bit (x&3)|(y&3)<<2|(z&3)<<4 */
void bw_build_fine(const uint8_t* grid, int n, uint64_t* fine)
{
  const int nb = n >> 2;
  memset(fine, 0, sizeof(uint64_t) * (size_t)nb * nb * nb);
  for (int z = 0; z < n; z++) for (int y = 0; y < n; y++) for (int x = 0; x < n; x++)
    if (grid[x + n * (y + n * z)] != 255)
      fine[(x >> 2) + nb * ((y >> 2) + nb * (z >> 2))] |= 1ull << ((x & 3) | ((y & 3) << 2) | ((z & 3) << 4));
}

/* rays: O[3], D[3] per ray (object space of the unit cube), tlimit per ray.  Returns the number of mismatching rays;
 * first_bad = index of the first one (or -1).  out (optional): per ray {hit, t bits, cell} of the cell walk. */
int64_t bw_compare(const uint8_t* grid, const uint64_t* fine, int n, const float* O, const float* D, const float* tlimit, int64_t count,
                   int mode, Stats* st, int64_t* first_bad, uint32_t* out)  /* mode 3: fine = Morton-ordered masks (bw_build_fine_morton) */
{
  int64_t bad = 0;
  *first_bad = -1;
  memset(st, 0, sizeof *st);
  for (int64_t i = 0; i < count; i++) {
    St s;
    Res a, b;
    memset(&a, 0, sizeof a); memset(&b, 0, sizeof b);
    if (setup(O + 3 * i, D + 3 * i, n, &s)) {
      st->rays++;
      a = cell_walk(grid, n, s, tlimit[i]);
      if (mode == 3) { int jr = 0; while ((16 << jr) < n) jr++; b = bit_walk(fine, jr, s, tlimit[i], st); }
      else b = mode ? unified_walk(fine, n, s, tlimit[i], st, mode == 2) : brick_walk(grid, fine, n, s, tlimit[i], st);
      st->ref_steps += a.steps;
    }
    uint32_t ta, tb; memcpy(&ta, &a.t, 4); memcpy(&tb, &b.t, 4);
    const int same = a.hit == b.hit && (!a.hit || (ta == tb && a.X[0] == b.X[0] && a.X[1] == b.X[1] && a.X[2] == b.X[2]));
    if (!same) { if (*first_bad < 0) *first_bad = i; bad++; }
    if (out) { out[3 * i] = (uint32_t)a.hit; out[3 * i + 1] = ta; out[3 * i + 2] = a.hit ? (uint32_t)(a.X[0] + n * (a.X[1] + n * a.X[2])) : 0u; }
  }
  return bad;
}
