/*
 * aux_kernels.cu -- K2 (rasteriser), K3 (Verlet lists), K4 (DEM sub-step), K5 (density),
 * K6 (output fields), force post-processing and layout conversion for sm_100a.
 *
 * Compiled with -fmad=false: the obstacle map must be bit-exact with the reference
 * (src/main.c:1026-1029) and the DEM step reproduces the reference bit for bit when its inputs
 * do (gather form over a sorted full neighbour list == the reference's scatter order, see
 * dem_forces_kernel).
 */
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace lbmdem {

using namespace lbm;

/* ------------------------------------------------------------------------------------------
 * K2: obst_construction (src/main.c:991-1065) without the delta[] array; act[] is a bit of the map
 * ---------------------------------------------------------------------------------------- */
#ifndef LBMDEM_SWEEP_MINB
#define LBMDEM_SWEEP_MINB 4
#endif

/* ------------------------------------------------------------------------------------------
 * K2, tile form (the default).  obst_construction (src/main.c:991-1065) as ONE pass over the lattice in tiles of
 * RTX x RTY nodes: no clearing pass, no global atomic per node, no re-reading of the map through L2.
 * ---------------------------------------------------------------------------------------- */
/* a tile's part of the map depends on the owners of its nodes AND of the halo node all round: a node whose covering
 * set changed stamps the tiles of its 3 x 3 neighbourhood; a tile stamped for the first time at this step joins the
 * step's list of tiles to rebuild */
__device__ __forceinline__ void stamp_tile(const TileBins &T, int t, int step) {
  if (atomicMax(&T.stamp[t], step) < step) T.dirty[(step & 1) * T.ntx * T.nty + atomicAdd(&T.ndirty[step & 1], 1)] = t;
}
__device__ __forceinline__ void stamp_around(const TileBins &T, int x, int y, int x0, int nxl, int ly, int step) {
  const int ra = max(x - 1, x0) - x0, rb = min(x + 1, x0 + nxl - 1) - x0, ya = max(y - 1, 0), yb = min(y + 1, ly - 1);
  if (rb < ra || yb < ya) return;
  const int ta = ra / RTX, tb = rb / RTX, ua = ya / RTY, ub = yb / RTY;
  stamp_tile(T, ta * T.nty + ua, step);
  if (ub != ua) stamp_tile(T, ta * T.nty + ub, step);
  if (tb != ta) {
    stamp_tile(T, tb * T.nty + ua, step);
    if (ub != ua) stamp_tile(T, tb * T.nty + ub, step);
  }
}

/* The nodes of lattice row `row` that a grain's reduced disc covers (src/main.c:1016-1029: inside the clamped bounding
 * box and dist2 <= r2, dist2 <= R2) are an interval of y -- the reference's expression is monotone in |y - yc|,
 * roundings included.  Its ends [*lo, *hi] (empty: *lo > *hi), EXACTLY: a float square root puts each end within a
 * node, the reference's own expression settles it. */
template <typename real>
__device__ __forceinline__ void row_cover_interval(int row, real xc, real yc, real r2, real RR, const GrainBox &b, int *lo,
                                                   int *hi) {
  *lo = 1;
  *hi = 0;
  if (row < b.xi || row > b.xf || b.yf < b.yi) return;
  const real rm = r2 < RR ? r2 : RR;
  const real dxr = row - xc;
  const real h2 = rm - dxr * dxr;
  if (!(h2 >= 0)) return; /* not even the node nearest to the centre line: see grain_bin_kernel's note on monotony */
  const float h = sqrtf((float)h2);
  int l = max((int)ceilf((float)yc - h), b.yi), u = min((int)floorf((float)yc + h), b.yf);
  /* at most one node off, either way */
  if (l - 1 >= b.yi && disc_covers(xc, yc, r2, RR, row, l - 1)) --l;
  else if (l <= b.yf && !disc_covers(xc, yc, r2, RR, row, l)) ++l;
  if (u + 1 <= b.yf && disc_covers(xc, yc, r2, RR, row, u + 1)) ++u;
  else if (u >= b.yi && !disc_covers(xc, yc, r2, RR, row, u)) --u;
  *lo = l;
  *hi = u;
}

/* Which nodes of row `row` does the disc cover under one placement and not under the other?  The symmetric difference
 * of the two intervals; every such node stamps its tiles.  Returns true if there are implausibly many (the caller then
 * stamps every tile the grain touches instead). */
template <typename real>
__device__ __forceinline__ bool row_cover_diff(const TileBins &T, int x0, int nxl, int ly, int step, int row, real xa, real ya,
                                               real r2a, real RRa, const GrainBox &ba, real xb, real yb, real r2b, real RRb,
                                               const GrainBox &bb) {
  int la, ha, lb, hb;
  row_cover_interval<real>(row, xa, ya, r2a, RRa, ba, &la, &ha);
  row_cover_interval<real>(row, xb, yb, r2b, RRb, bb, &lb, &hb);
  if (la == lb && ha == hb) return false;
  if (la > ha && lb > hb) return false; /* both empty */
  const int ymin = min(la > ha ? lb : la, lb > hb ? la : lb), ymax = max(la > ha ? hb : ha, lb > hb ? ha : hb);
  if (ymax - ymin > 64) return true;
  for (int y = ymin; y <= ymax; ++y) {
    const bool ca = y >= la && y <= ha, cb = y >= lb && y <= hb;
    if (ca != cb) stamp_around(T, row, y, x0, nxl, ly, step);
  }
  return false;
}

/* First kernel of the step, one WARP per grain: the grain's record for this step, its tile bins, and -- the point of
 * it -- which lattice tiles hold a node that the grain covers now but did not cover in the previous step's map, or
 * the other way round.  A grain moves by 1e-3 .. 1e-2 node per LBM step, so nearly every tile of the map (and of the
 * two link lists) is the previous step's: the tile kernel rebuilds the stamped tiles alone.  Stamping errs on the safe
 * side: a grain that moved by half a node or more, or whose radius changed, stamps every tile it touches, before and
 * after.  first_run: no previous map (set-up). */
template <typename real>
__global__ void __launch_bounds__(128) grain_bin_kernel(RasterParams<real> P, int n, GrainArrays<real> g, GrainRec<real> *rec,
                                                        real *R2, GrainBox *boxes, const GrainRec<real> *rec_old,
                                                        const real *R2_old, const GrainBox *boxes_old, int x0, int nxl,
                                                        TileBins T, int step, int first_run, int *defer_count,
                                                        long long *facc) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *defer_count = 0;
  if (i >= n) return;
  GrainRec<real> r;
  GrainBox b;
  real R2i;
  grain_geometry(P, g.x1[i], g.x2[i], g.r[i], g.rLB[i], &r.xc, &r.yc, &r.r2, &R2i, &b);
  bool all_tiles = first_run != 0;
  GrainBox ob = b;
  if (!first_run) ob = boxes_old[i];
  /* strip-decomposed runs replicate the grains: most of them are nowhere near this rank's rows, now or a step ago */
  const bool far_now = b.xf + 1 < x0 || b.xi - 1 > x0 + nxl - 1 || b.xf < b.xi || b.yf < b.yi;
  const bool far_old = ob.xf + 1 < x0 || ob.xi - 1 > x0 + nxl - 1 || ob.xf < ob.xi || ob.yf < ob.yi;
  if (far_now && far_old) {
    if (lane == 0) {
      r.x1 = g.x1[i]; r.x2 = g.x2[i]; r.v1 = g.v1[i]; r.v2 = g.v2[i]; r.v3 = g.v3[i];
      rec[i] = r;
      R2[i] = R2i;
      boxes[i] = b;
    }
    if (facc != nullptr && lane < 3) facc[lane * n + i] = 0;
    return;
  }
  if (!all_tiles) {
    const GrainRec<real> o = rec_old[i];
    const real RRo = R2_old[i];
    if (!(fabs((double)(r.xc - o.xc)) < 0.5 && fabs((double)(r.yc - o.yc)) < 0.5) || o.r2 != r.r2 || RRo != R2i) {
      all_tiles = true;
    } else {
      const int ra = min(b.xi, ob.xi), rb = max(b.xf, ob.xf);
      bool d = false;
      for (int row = ra + lane; row <= rb; row += 32)
        d |= row_cover_diff<real>(T, x0, nxl, P.ly, step, row, o.xc, o.yc, o.r2, RRo, ob, r.xc, r.yc, r.r2, R2i, b);
      all_tiles = __any_sync(0xffffffffu, d);
    }
  }
  if (lane == 0) {
    r.x1 = g.x1[i]; r.x2 = g.x2[i]; r.v1 = g.v1[i]; r.v2 = g.v2[i]; r.v3 = g.v3[i];
    rec[i] = r;
    R2[i] = R2i;
    boxes[i] = b;
  }
  if (facc != nullptr && lane < 3) facc[lane * n + i] = 0;
  /* every tile whose nodes or halo the bounding box touches (local rows only): a lane per tile */
  {
    const int xa = max(b.xi - 1, x0), xb = min(b.xf + 1, x0 + nxl - 1);
    const int ya = max(b.yi - 1, 0), yb = min(b.yf + 1, P.ly - 1);
    if (!(b.xf < b.xi || b.yf < b.yi || xb < xa || yb < ya)) {
      const int tx0 = (xa - x0) / RTX, ntx = (xb - x0) / RTX - tx0 + 1, ty0 = ya / RTY, nty = yb / RTY - ty0 + 1;
      for (int k = lane; k < ntx * nty; k += 32) {
        const int t = (tx0 + k / nty) * T.nty + ty0 + k % nty;
        const int slot = atomicAdd(&T.count[t], 1);
        if (slot < T.cap) {
          TileEntry<real> en;
          en.id = i; en.xi = b.xi; en.xf = b.xf; en.yi = b.yi; en.yf = b.yf;
          en.xc = r.xc; en.yc = r.yc; en.r2 = r.r2; en.RR = R2i;
          static_cast<TileEntry<real> *>(T.list)[(size_t)t * T.cap + slot] = en;
        } else {
          *(volatile int *)T.overflow = 1; /* mapped host memory */
        }
        if (all_tiles) stamp_tile(T, t, step);
      }
    }
  }
  /* ... and the tiles the previous placement touched */
  if (all_tiles && !first_run) {
    const int xa = max(ob.xi - 1, x0), xb = min(ob.xf + 1, x0 + nxl - 1);
    const int ya = max(ob.yi - 1, 0), yb = min(ob.yf + 1, P.ly - 1);
    if (!(ob.xf < ob.xi || ob.yf < ob.yi || xb < xa || yb < ya)) {
      const int tx0 = (xa - x0) / RTX, ntx = (xb - x0) / RTX - tx0 + 1, ty0 = ya / RTY, nty = yb / RTY - ty0 + 1;
      for (int k = lane; k < ntx * nty; k += 32) stamp_tile(T, (tx0 + k / nty) * T.nty + ty0 + k % nty, step);
    }
  }
}

constexpr int RTC0 = 4;                      /* shared-memory column of tile column 0 (16-byte aligned rows of int4) */
constexpr int RTP = RTY + 8;                 /* shared-memory row pitch: 3 unused, left halo, RTY nodes, right halo, 3 unused */
constexpr int RT_THREADS = 256;
constexpr int RT_CHUNK = 64;                 /* grains staged in shared memory at a time */
static_assert(RTY == 64 && RTX == 32 && RT_THREADS == 256, "pass 3a: a warp takes 4 rows, a lane 4 consecutive columns");

template <typename real>
__global__ void __launch_bounds__(RT_THREADS) raster_tile_kernel(int n, int *cell, const int *cell_other, unsigned char *cls,
                                                                 const unsigned char *cls_other, unsigned short *own16,
                                                                 const unsigned short *own16_other, int x0, int nxl, int pitch,
                                                                 int lx, int ly, TileBins T, BoundaryList B, LinkList K,
                                                                 int step, int force_full) {
  /* region = tile + one halo node all round; region row r+1 / column c+RTC0 hold tile node (r, c) */
  __shared__ __align__(16) int own[RTX + 2][RTP]; /* owner: -1 fluid, n ring / outside the array */
  __shared__ __align__(16) int low[RTX + 2][RTP]; /* lowest covering index where more than one disc covers the node */
  __shared__ unsigned short cand[RTX * RTY]; /* tile nodes with a neighbour under another owner (r * RTY + c, bit 15: act) */
  __shared__ unsigned short msk[RTX * RTY];  /* per candidate: fluid neighbours | foreign non-fluid neighbours << 8 */
  __shared__ int s_ncand, s_nlinks, s_nnodes, s_last;
  __shared__ TileEntry<real> s_en[RT_CHUNK];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int ntiles = T.ntx * T.nty;
  /* Work items.  Tiles stamped at this step (some covered node changed in them or in their halo) are REBUILT.
   * Tiles stamped at the previous step only are COPIED from the other copy of the map: `cell` was last written two
   * steps ago, and their list segments -- rebuilt last step -- are still right.  Every other tile is right in both
   * copies already.  force_full: every tile is rebuilt. */
  const int nd_cur = force_full ? ntiles : T.ndirty[step & 1];
  const int nd_all = force_full ? ntiles : nd_cur + T.ndirty[(step - 1) & 1];
  for (int item = blockIdx.x; item < nd_all; item += gridDim.x) {
  int tile = item;
  bool copy_only = false;
  if (!force_full) {
    if (item < nd_cur) {
      tile = T.dirty[(step & 1) * ntiles + item];
    } else {
      tile = T.dirty[((step - 1) & 1) * ntiles + item - nd_cur];
      if (T.stamp[tile] == step) continue; /* rebuilt at this step as well: it is in the first list */
      copy_only = true;
    }
  }
  const int tile_x = tile / T.nty, tile_y = tile - tile_x * T.nty;
  const int tx0 = x0 + tile_x * RTX, ty0 = tile_y * RTY; /* lattice coordinates of tile node (0, 0) */
  if (copy_only) {
#pragma unroll
    for (int it = 0; it < RTX / 16; ++it) {
      const int r = w * (RTX / 8) + it * 2 + (lane >> 4), c0 = (lane & 15) * 4;
      const int x = tx0 + r;
      if (x <= x0 + nxl - 1 && ty0 + c0 + 3 < pitch) {
        const size_t k = (size_t)(x - x0) * pitch + ty0 + c0;
        *reinterpret_cast<int4 *>(&cell[k]) = *reinterpret_cast<const int4 *>(&cell_other[k]);
        *reinterpret_cast<uchar4 *>(&cls[k]) = *reinterpret_cast<const uchar4 *>(&cls_other[k]);
        *reinterpret_cast<ushort4 *>(&own16[k]) = *reinterpret_cast<const ushort4 *>(&own16_other[k]);
      }
    }
    continue;
  }
  /* the first list entries are fetched along with the count (entries past the count are ignored) */
  const TileEntry<real> *list = static_cast<const TileEntry<real> *>(T.list) + (size_t)tile * T.cap;
  if (tid < min(T.cap, RT_CHUNK)) s_en[tid] = list[tid];
  const int cnt = min(T.count[tile], T.cap);

  /* ---- 1. init_obst's frame (:674-687) for the region: ring (and beyond the array) = n, interior fluid ---- */
  if (tid == 0) { s_ncand = 0; s_nlinks = 0; s_nnodes = 0; }
  if (tx0 - 1 > 0 && tx0 + RTX < lx - 1 && ty0 - 1 > 0 && ty0 + RTY < ly - 1) {
    int4 *o4 = reinterpret_cast<int4 *>(&own[0][0]), *l4 = reinterpret_cast<int4 *>(&low[0][0]);
    for (int idx = tid; idx < (RTX + 2) * RTP / 4; idx += RT_THREADS) {
      o4[idx] = make_int4(-1, -1, -1, -1);
      l4[idx] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
    }
  } else {
    for (int idx = tid; idx < (RTX + 2) * RTP; idx += RT_THREADS) {
      const int r = idx / RTP, c = idx - r * RTP;
      const int gx = tx0 - 1 + r, gy = ty0 - RTC0 + c;
      own[r][c] = (gx <= 0 || gx >= lx - 1 || gy <= 0 || gy >= ly - 1) ? n : -1;
      low[r][c] = 0x7fffffff;
    }
  }

  /* ---- 2. the discs (:1016-1031).  The tile's grains are staged in shared memory (one round of global loads for
   * all of them), then painted, a warp per grain: lanes along y (up to three columns each), rows in turn.
   * Owner = highest covering index (the reference paints in index order), hence atomicMax; where a disc finds the
   * node taken the lower of the two indices is kept: the minimum over all such meetings is the lowest covering
   * index.  dist2 is the reference's expression (:1026), its two squares rounded separately. ---- */
  for (int base = 0; base < cnt; base += RT_CHUNK) {
    const int m = min(RT_CHUNK, cnt - base);
    if (base > 0 && tid < m) s_en[tid] = list[base + tid];
    __syncthreads();
    /* few grains: several warps per grain, each a share of its rows */
#if defined(LBMDEM_RT_NOSPLIT) /* measurement variant: always a warp per grain */
    const int parts = 1;
#else
    const int parts = m < RT_THREADS / 32 ? (RT_THREADS / 32) / m : 1;
#endif
    for (int k = w / parts; k < m; k += (RT_THREADS / 32) / parts) {
      const int i = s_en[k].id;
      const real xc = s_en[k].xc, yc = s_en[k].yc, r2 = s_en[k].r2, RR = s_en[k].RR;
      const int ra = max(s_en[k].xi, tx0 - 1), rb = min(s_en[k].xf, tx0 + RTX);
      const int ca = max(s_en[k].yi, ty0 - 1), cb = min(s_en[k].yf, ty0 + RTY);
      const int y0 = ca + lane, y1 = y0 + 32, y2 = y0 + 64; /* the region is RTY + 2 = 66 nodes wide */
      const int s0 = y0 - ty0 + RTC0, s1 = s0 + 32, s2 = s0 + 64; /* their shared-memory columns */
      const real d0 = (y0 - yc) * (y0 - yc), d1 = (y1 - yc) * (y1 - yc), d2 = (y2 - yc) * (y2 - yc);
      for (int x = ra + w % parts; x <= rb; x += parts) {
        const real dx2 = (x - xc) * (x - xc);
        int *orow = own[x - tx0 + 1], *lrow = low[x - tx0 + 1];
        const real e0 = dx2 + d0, e1 = dx2 + d1, e2 = dx2 + d2;
        /* the three atomics are issued together; their results are looked at afterwards */
        int o0 = -1, o1 = -1, o2 = -1;
        if (y0 <= cb && e0 <= RR && e0 <= r2) o0 = atomicMax(&orow[s0], i);
        if (y1 <= cb && e1 <= RR && e1 <= r2) o1 = atomicMax(&orow[s1], i);
        if (y2 <= cb && e2 <= RR && e2 <= r2) o2 = atomicMax(&orow[s2], i);
        if (o0 >= 0 && o0 != i) atomicMin(&lrow[s0], min(o0, i));
        if (o1 >= 0 && o1 != i) atomicMin(&lrow[s1], min(o1, i));
        if (o2 >= 0 && o2 != i) atomicMin(&lrow[s2], min(o2, i));
      }
    }
    __syncthreads();
  }
  if (cnt == 0) __syncthreads(); /* the frame of pass 1 */

  /* ---- 3a. every node of the tile: the owners go to `cell` in 16-byte pieces (a lane holds four consecutive
   * columns, a warp two rows), and the few nodes whose 3 x 3 neighbourhood is not under one owner are collected
   * for pass 3b ---- */
#pragma unroll 1
  for (int it = 0; it < RTX / 16; ++it) {
    const int r = w * (RTX / 8) + it * 2 + (lane >> 4), c0 = (lane & 15) * 4; /* tile coordinates */
    const int x = tx0 + r;
    const int *pm = &own[r][RTC0 + c0], *p0 = &own[r + 1][RTC0 + c0], *pp = &own[r + 2][RTC0 + c0];
    const int4 a = *reinterpret_cast<const int4 *>(pm), v = *reinterpret_cast<const int4 *>(p0),
               d = *reinterpret_cast<const int4 *>(pp);
    const int aL = pm[-1], aR = pm[4], vL = p0[-1], vR = p0[4], dL = pp[-1], dR = pp[4];
    unsigned hit = 0;
    if (x >= x0 + 1 && x <= x0 + nxl - 2) { /* rows whose neighbours are held locally */
      if (v.x >= 0 && v.x < n && !(aL == v.x && a.x == v.x && a.y == v.x && vL == v.x && v.y == v.x && dL == v.x && d.x == v.x && d.y == v.x)) hit |= 1u;
      if (v.y >= 0 && v.y < n && !(a.x == v.y && a.y == v.y && a.z == v.y && v.x == v.y && v.z == v.y && d.x == v.y && d.y == v.y && d.z == v.y)) hit |= 2u;
      if (v.z >= 0 && v.z < n && !(a.y == v.z && a.z == v.z && a.w == v.z && v.y == v.z && v.w == v.z && d.y == v.z && d.z == v.z && d.w == v.z)) hit |= 4u;
      if (v.w >= 0 && v.w < n && !(a.z == v.w && a.w == v.w && aR == v.w && v.z == v.w && vR == v.w && d.z == v.w && d.w == v.w && dR == v.w)) hit |= 8u;
    }
    if (x <= x0 + nxl - 1 && ty0 + c0 + 3 < pitch) {
      const size_t k = (size_t)(x - x0) * pitch + ty0 + c0;
      *reinterpret_cast<int4 *>(&cell[k]) = v;
      *reinterpret_cast<uchar4 *>(&cls[k]) = make_uchar4(cell_class(v.x, n), cell_class(v.y, n), cell_class(v.z, n), cell_class(v.w, n));
      *reinterpret_cast<ushort4 *>(&own16[k]) = make_ushort4(cell_own16(v.x), cell_own16(v.y), cell_own16(v.z), cell_own16(v.w));
    }
    if (__any_sync(0xffffffffu, hit != 0)) {
      const int mine = __popc(hit);
      int incl = mine;
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += t;
      }
      int at = 0;
      if (lane == 31) at = atomicAdd(&s_ncand, incl);
      at = __shfl_sync(0xffffffffu, at, 31) + incl - mine;
      while (hit) {
        const int bit = __ffs(hit) - 1;
        hit &= hit - 1;
        cand[at++] = (unsigned short)(r * RTY + c0 + bit);
      }
    }
  }
  __syncthreads();

  /* ---- 3b. the candidates, one per lane: act[x][y] (:1036-1052), rim bit, link masks; the two bits are added to
   * the node's entry of `cell` (the line is still in L2).  A thread takes a contiguous share of the list ---- */
  const int ncand = s_ncand;
  const int share = (ncand + RT_THREADS - 1) / RT_THREADS;
  const int e0 = min(tid * share, ncand), e1 = min(e0 + share, ncand);
  unsigned packed = 0; /* this thread's list entries: links | nodes << 16 */
  for (int e = e0; e < e1; ++e) {
    const int rc = cand[e], r = rc / RTY, c = rc % RTY;
    const int x = tx0 + r, y = ty0 + c;
    const int i = own[r + 1][c + RTC0];
    bool act = false;
    unsigned foreign = 0, fluid = 0;
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
      const int rr = r + 1 + ex_of(q), cc = c + RTC0 + ey_of(q);
      const int cn = own[rr][cc];
      if (cell_is_fluid(cn)) fluid |= 1u << (q - 1);
      if (cn != i) {
        foreign |= 1u << (q - 1);
        /* a neighbour owned by a LATER grain counted as fluid when grain i ran unless a grain j <= i lies
         * under it as well (lbm_node.cuh, fluid_when_grain_ran_exact) */
        int mo = -1;
        if (cn > i && cn < n) { const int lo = low[rr][cc]; mo = lo == 0x7fffffff ? -1 : lo; }
        if (fluid_when_grain_ran_exact(cn, i, n, mo)) act = true;
      }
    }
    const unsigned solid_foreign = foreign & ~fluid; /* != 0: the force kernel's share (CELL_RIM) */
    msk[e] = (unsigned short)(fluid | (solid_foreign << 8));
    if (act) {
      cand[e] = (unsigned short)(rc | 0x8000);
      const bool near_ring = !(x >= 2 && y >= 2 && x <= lx - 3 && y <= ly - 3);
      packed += (unsigned)__popc(near_ring ? 0xffu : fluid); /* bounce links; next to the ring the w-links too */
    }
    if (solid_foreign) packed += 1u << 16;
    if (act || solid_foreign) {
      cell[(size_t)(x - x0) * pitch + y] = i | (act ? CELL_ACT : 0) | (solid_foreign ? CELL_RIM : 0);
      cls[(size_t)(x - x0) * pitch + y] = (unsigned char)(CLS_SOLID | (act ? CLS_ACT : 0) | (solid_foreign ? CLS_RIM : 0));
    }
  }

  /* ---- 4. one slot range per list and WARP inside the tile's own segments of the two lists ---- */
  unsigned incl = packed;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
  int bl = 0, bb = 0;
  if (lane == 0 && (total & 0xffffu)) bl = atomicAdd(&s_nlinks, (int)(total & 0xffffu));
  if (lane == 1 && (total >> 16)) bb = atomicAdd(&s_nnodes, (int)(total >> 16));
  bl = __shfl_sync(0xffffffffu, bl, 0);
  bb = __shfl_sync(0xffffffffu, bb, 1);
  const unsigned excl = incl - packed;
  int kl = bl + (int)(excl & 0xffffu), kb = bb + (int)(excl >> 16);
  uint2 *Kseg = K.entry + (size_t)tile * K.cap, *Bseg = B.entry + (size_t)tile * B.cap;

  /* ---- 5. the entries ---- */
  for (int e = e0; e < e1; ++e) {
    const unsigned m = msk[e], rc = cand[e];
    const bool act = (rc & 0x8000u) != 0;
    const unsigned fluid = m & 0xffu, solid_foreign = m >> 8;
    if (!solid_foreign && !act) continue;
    const int r = (int)(rc & 0x7fffu) / RTY, c = (int)(rc & 0x7fffu) % RTY;
    const int x = tx0 + r, y = ty0 + c;
    const unsigned knode = (unsigned)((x - x0) * pitch + y);
    const unsigned i = (unsigned)own[r + 1][c + RTC0];
    if (solid_foreign) { /* the force kernel's share: all foreign neighbours, the kernel skips the fluid ones */
      if (kb < B.cap) Bseg[kb] = make_uint2(knode, ((fluid | solid_foreign) << 24) | (act ? BL_ACT : 0u) | i);
      else *(volatile int *)B.overflow = 1;
      ++kb;
    }
    if (!act) continue;
    const bool near_ring = !(x >= 2 && y >= 2 && x <= lx - 3 && y <= ly - 3);
    unsigned links = near_ring ? 0xffu : fluid; /* bounce links (fluid neighbour); next to the ring the others as w-links */
    while (links) {
      const int bit = __ffs(links) - 1;
      links &= links - 1;
      const bool isw = !((fluid >> bit) & 1u);
      /* the node two steps along the link, where the painted region (tile + one halo node) reaches it: fluid or wall
       * ring there means the link is not one across a one-node gap, and the sweep need not read the map to find out */
      const int r2 = r + 2 * ex_of(bit + 1), c2 = c + 2 * ey_of(bit + 1);
      bool clear = false;
      if (r2 >= -1 && r2 <= RTX && c2 >= -1 && c2 <= RTY) {
        const int o2 = own[r2 + 1][c2 + RTC0];
        clear = o2 < 0 || o2 >= n;
      }
      if (kl < K.cap) Kseg[kl] = make_uint2(knode, i | ((unsigned)(bit + 1) << 24) | (isw ? LL_W : 0u) | (clear ? LL_CLEAR : 0u));
      else *(volatile int *)K.overflow = 1;
      ++kl;
    }
  }
  __syncthreads();
  if (tid == 0) {
    K.tcount[tile] = min(s_nlinks, K.cap);
    B.tcount[tile] = min(s_nnodes, B.cap);
  }
  __syncthreads(); /* the next tile re-uses the shared arrays */
  } /* work items */
  /* the last CTA to finish empties the bins of every tile and the dirty list the next step will fill */
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(T.ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    for (int t = tid; t < ntiles; t += RT_THREADS) T.count[t] = 0;
    if (tid == 0) {
      T.ndirty[(step + 1) & 1] = 0;
      *T.ticket = 0;
    }
  }
}

template <typename real>
cudaError_t launch_raster_tiles(const RasterParams<real> &P, int n, const GrainArrays<real> &g, GrainRec<real> *rec, real *R2,
                                GrainBox *boxes, const GrainRec<real> *rec_old, const real *R2_old, const GrainBox *boxes_old,
                                int *cell, const int *cell_other, unsigned char *cls, const unsigned char *cls_other,
                                unsigned short *own16, const unsigned short *own16_other, int x0, int nxl, int pitch,
                                const TileBins &T, const BoundaryList &B, const LinkList &K,
                                int *defer_count, long long *facc, int step, int first_run, int force_full, cudaStream_t s) {
  grain_bin_kernel<real><<<(n + 3) / 4, 128, 0, s>>>(P, n, g, rec, R2, boxes, rec_old, R2_old, boxes_old, x0, nxl, T, step,
                                                     first_run, defer_count, facc);
  const int ntiles = T.ntx * T.nty;
  const int ctas = force_full ? ntiles : min(ntiles, T.resident_ctas);
  raster_tile_kernel<real><<<ctas, RT_THREADS, 0, s>>>(n, cell, cell_other, cls, cls_other, own16, own16_other, x0, nxl, pitch,
                                                       P.lx, P.ly, T, B, K, step, force_full);
  return cudaGetLastError();
}

/* init_obst's frame (:674-687): ring = nbgrains, interior = -1 */
__global__ void cell_frame_kernel(int *cell, unsigned char *cls, unsigned short *own16, int lx, int ly, int x0, int nxl,
                                  int pitch, int ring_value) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (y >= pitch || row >= nxl) return;
  const int x = x0 + row;
  int v = -1;
  if (x <= 0 || x >= lx - 1 || y <= 0 || y >= ly - 1) v = ring_value;
  cell[(size_t)row * pitch + y] = v;
  cls[(size_t)row * pitch + y] = cell_class(v, ring_value);
  own16[(size_t)row * pitch + y] = cell_own16(v);
}

cudaError_t launch_cell_frame(int *cell, unsigned char *cls, unsigned short *own16, int lx, int ly, int x0, int nxl, int pitch,
                              int ring_value, cudaStream_t s) {
  dim3 grid((pitch + 255) / 256, nxl);
  cell_frame_kernel<<<grid, 256, 0, s>>>(cell, cls, own16, lx, ly, x0, nxl, pitch, ring_value);
  return cudaGetLastError();
}

__global__ void cls_from_cell_kernel(const int *cell, unsigned char *cls, unsigned short *own16, size_t count, int ngrains) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) {
    cls[k] = cell_class(cell[k], ngrains);
    own16[k] = cell_own16(cell[k]);
  }
}
cudaError_t launch_cls_from_cell(const int *cell, unsigned char *cls, unsigned short *own16, int nxl, int pitch, int ngrains,
                                 cudaStream_t s) {
  const size_t count = (size_t)nxl * pitch;
  cls_from_cell_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(cell, cls, own16, count, ngrains);
  return cudaGetLastError();
}

template <typename real>
__global__ void act_map_kernel(Lattice<real> L, Stored<real> S, int xlo, int xhi, int *act_out) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = xlo + blockIdx.y;
  if (y >= L.ly || x >= xhi) return;
  int a = 0;
  if (!is_ring(L, x, y)) {
    const int c = S.cell[node_index(L, x, y)];
    a = cell_is_fluid(c) ? 1 : (node_act(L, S, x, y, c) ? 1 : 0); /* the reference clears act to 1 on fluid nodes */
  }
  act_out[(size_t)(x - xlo) * L.ly + y] = a;
}
template <typename real>
cudaError_t launch_act_map(const Lattice<real> &L, const Stored<real> &S, int xlo, int xhi, int *act_out, cudaStream_t s) {
  dim3 grid((L.ly + 255) / 256, xhi - xlo);
  act_map_kernel<real><<<grid, 256, 0, s>>>(L, S, xlo, xhi, act_out);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------
 * Sweep 3 in place: wall ring (src/main.c:1123-1145), two passes (lbm_node.cuh, ring_value), one
 * thread per ring node of the rows [xa, xb).
 * ---------------------------------------------------------------------------------------- */
template <typename real>
__global__ void ring_sweep_kernel(const __grid_constant__ Lattice<real> L, const __grid_constant__ Stored<real> S, real *A,
                                  int pass, int xa, int xb, int rows_lo, int rows_hi) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int x, y;
  if (pass == 0) { /* y = 0 and y = ly-1 of every row in range */
    if (t >= 2ll * (xb - xa)) return;
    x = xa + (int)(t >> 1);
    y = (t & 1) ? L.ly - 1 : 0;
  } else { /* the owned ring rows x = 0 / x = lx-1 in full (corners included) */
    if (t >= (long long)(rows_lo + rows_hi) * L.ly) return;
    const int r = (int)(t / L.ly);
    y = (int)(t - (long long)r * L.ly);
    x = (r < rows_lo) ? 0 : L.lx - 1;
  }
  const size_t k = node_index(L, x, y);
  real v[NQ];
#pragma unroll
  for (int q = 1; q < NQ; ++q) v[q] = ring_value(L, S, pass, x, y, q);
#pragma unroll
  for (int q = 1; q < NQ; ++q) A[q * L.plane + k] = v[q];
}
template <typename real>
cudaError_t launch_ring_sweep(const Lattice<real> &L, const Stored<real> &S, real *A, int xa, int xb, cudaStream_t s) {
  if (xb <= xa) return cudaSuccess;
  const int lo = (xa == 0) ? 1 : 0, hi = (xb == L.lx) ? 1 : 0;
  const long long t0 = 2ll * (xb - xa), t1 = (long long)(lo + hi) * L.ly;
  ring_sweep_kernel<real><<<(unsigned)((t0 + 127) / 128), 128, 0, s>>>(L, S, A, 0, xa, xb, lo, hi);
  if (t1 > 0) ring_sweep_kernel<real><<<(unsigned)((t1 + 127) / 128), 128, 0, s>>>(L, S, A, 1, xa, xb, lo, hi);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------
 * Sweep 4 in place: interpolated bounce-back on active solid nodes (src/main.c:1154-1222).
 * One warp per grain: the lanes scan the grain's bounding box, compact its active nodes into a
 * per-warp list, then share the (node, link) pairs out, one link per lane, so that the expensive
 * part (delta: one sqrt, the interpolation: divisions) runs with full lanes.
 * ---------------------------------------------------------------------------------------- */
/* Adds three fixed-point values to the sums of grain i (i < 0: nothing to add).  EVERY lane of the warp calls it.
 * Lanes that hold the same grain are summed first (exact: integers), so there is about one atomic per grain and
 * warp.  The link list is grouped by grain, so the lanes of a grain normally form one contiguous run: a segmented
 * shuffle reduction (five rounds); any other pattern takes the lane-by-lane loop. */
#if !defined(LBMDEM_SUMS_MATCH)
/* One segmented reduction over every maximal run of adjacent
 * lanes with the same grain, one atomic per run -- no lane-by-lane loop whatever the pattern.  The tile rasteriser
 * emits links tile row by tile row, so a warp usually sees A A B B A A B B ...; profiles/README.md: the lane-by-lane
 * loop below is then a third of this kernel's instructions. */
__device__ __forceinline__ void grain_sums_add(long long *facc, int n, int i, long long s1, long long s2, long long s3) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int prev = __shfl_up_sync(full, i, 1);
  const bool head = lane == 0 || prev != i;
  const unsigned heads = __ballot_sync(full, head);
  const unsigned above = lane == 31 ? 0u : heads >> (lane + 1);
  const int end = above ? lane + __ffs(above) : 32; /* the next run starts there */
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long t1 = __shfl_down_sync(full, s1, d), t2 = __shfl_down_sync(full, s2, d), t3 = __shfl_down_sync(full, s3, d);
    if (lane + d < end) { s1 += t1; s2 += t2; s3 += t3; }
  }
  if (head && i >= 0) {
    atomicAdd((unsigned long long *)&facc[i], (unsigned long long)s1);
    atomicAdd((unsigned long long *)&facc[n + i], (unsigned long long)s2);
    atomicAdd((unsigned long long *)&facc[2 * n + i], (unsigned long long)s3);
  }
}
#else
__device__ __forceinline__ void grain_sums_add(long long *facc, int n, int i, long long s1, long long s2, long long s3) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned peers = __match_any_sync(full, i);
  const int first = __ffs(peers) - 1, cnt = __popc(peers);
  const bool run = (peers >> first) == (cnt == 32 ? full : ((1u << cnt) - 1u));
  if (__all_sync(full, run)) {
    const int end = first + cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long t1 = __shfl_down_sync(full, s1, d), t2 = __shfl_down_sync(full, s2, d), t3 = __shfl_down_sync(full, s3, d);
      if (lane + d < end) { s1 += t1; s2 += t2; s3 += t3; }
    }
  } else {
    long long t1 = 0, t2 = 0, t3 = 0;
    for (unsigned m = peers; m; m &= m - 1) {
      const int src = __ffs(m) - 1;
      t1 += __shfl_sync(peers, s1, src);
      t2 += __shfl_sync(peers, s2, src);
      t3 += __shfl_sync(peers, s3, src);
    }
    s1 = t1; s2 = t2; s3 = t3;
  }
  if (lane == first && i >= 0) {
    atomicAdd((unsigned long long *)&facc[i], (unsigned long long)s1);
    atomicAdd((unsigned long long *)&facc[n + i], (unsigned long long)s2);
    atomicAdd((unsigned long long *)&facc[2 * n + i], (unsigned long long)s3);
  }
}
#endif

/* One thread per listed link.  Besides the sweep itself the thread holds both operands of the
 * link's momentum exchange (forces_fluid, :1313-1320: f_new[s][opp q] = A[n][opp q] and
 * f_new[n][q] = the value it just produced), so links into FLUID neighbours are added to the
 * grain's force sums here (facc != nullptr, owned rows only); force_links_kernel adds the rest. */
template <typename real>
__device__ __forceinline__ void bounce_tile_links(const Lattice<real> &L, const Stored<real> &S, real *A, int xa, int xb,
                                                  int xlo, int xhi, const LinkList &K, const DeferList<real> &D,
                                                  long long *facc, int tile) {
  /* the tile's own segment of the link list */
  const int items = K.tcount[tile];
  if (items == 0) return;
  const uint2 *seg = K.entry + (size_t)tile * K.cap;
  const int padded = (items + 31) & ~31; /* whole warps stay in the loop: grain_sums_add shuffles */
  for (int u = threadIdx.x; u < padded; u += blockDim.x) {
    int fi = -1;
    long long s1 = 0, s2 = 0, s3 = 0;
    if (u < items) {
      const uint2 en = seg[u];
      const int q = (int)((en.y >> 24) & 15u);
      const int row = (int)(en.x / (unsigned)L.pitch);
      const int x = L.x0 + row, y = (int)(en.x - (unsigned)row * (unsigned)L.pitch);
      if (x >= xa && x < xb) {
        const size_t e = q * L.plane + en.x;
        if (en.y & LL_W) { /* rest value next to the wall ring (the fused kernel does the others) */
          A[e] = L.w[q];
        } else {
          const int owner = (int)(en.y & BL_GRAIN);
          const GrainRec<real> g = S.grains[owner];
          real v, Fn_oq;
          bool gap;
          /* listed bounce links have a fluid neighbour; a link across a one-node gap is evaluated from the
           * pre-sweep state and filed in the deferred list (lbm_node.cuh, sweep_link) */
          const int r = sweep_link_core(L, S, g, x, y, q, true, &v, true, &gap, &Fn_oq, (en.y & LL_CLEAR) != 0);
          if (r == SWEEP_WRITE) {
            if (!gap) {
              A[e] = v;
            } else {
              /* one counter update per group of lanes that got here together */
              const unsigned grp = __activemask();
              const int leader = __ffs(grp) - 1, lane = threadIdx.x & 31;
              int slot = 0;
              if (lane == leader) slot = atomicAdd(D.count, __popc(grp));
              slot = __shfl_sync(grp, slot, leader) + __popc(grp & ((1u << lane) - 1));
              if (slot < D.capacity) { D.index[slot] = e; D.value[slot] = v; }
              else *(volatile int *)D.overflow = 1; /* mapped host memory */
            }
          }
          if (facc != nullptr && x >= xlo && x < xhi) {
            if (r == SWEEP_KEEP) v = A[e];
            real h1 = 0, h2 = 0, h3 = 0;
            force_link<real>(q, Fn_oq, v, x, y, g.xc, g.yc, &h1, &h2, &h3);
            fi = owner;
            if (!(fabs((double)h1) < 1024. && fabs((double)h2) < 1024. && fabs((double)h3) < 16384.))
              *(volatile int *)D.range_flag = 1; /* NaN or diverged populations: the sums would wrap */
            s1 = __double2ll_rn((double)h1 * FORCE_FIX);
            s2 = __double2ll_rn((double)h2 * FORCE_FIX);
            s3 = __double2ll_rn((double)h3 * TORQUE_FIX);
          }
        }
      }
    }
    if (facc != nullptr) grain_sums_add(facc, L.ngrains, fi, s1, s2, s3);
  }
}
template <typename real>
__global__ void __launch_bounds__(256, LBMDEM_SWEEP_MINB) bounce_sweep_kernel(const __grid_constant__ Lattice<real> L,
                                                           const __grid_constant__ Stored<real> S, real *A, int xa, int xb,
                                                           int xlo, int xhi, const LinkList K, const DeferList<real> D,
                                                           long long *facc, int tx_first, int nty) {
  /* one CTA per lattice tile (tile rows tx_first ..) */
  bounce_tile_links<real>(L, S, A, xa, xb, xlo, xhi, K, D, facc, (tx_first + blockIdx.y) * nty + blockIdx.x);
}
template <typename real>
__global__ void defer_apply_kernel(real *A, const DeferList<real> D) {
  const int n = min(*D.count, D.capacity);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) A[D.index[k]] = D.value[k];
}
/* The sweep in three stages, so that a strip-decomposed run can sweep the rows that need no ghost data while
 * the ghost rows are still in flight: begin (empty deferred list, zero force sums), any number of passes over
 * disjoint row ranges, end (apply the deferred links -- only after EVERY pass, they read pre-sweep values). */
template <typename real>
cudaError_t launch_bounce_pass(const Lattice<real> &L, const Stored<real> &S, real *A, int xa, int xb, int xlo, int xhi,
                               const LinkList &K, const DeferList<real> &D, long long *facc, cudaStream_t s) {
  if (xb <= xa || L.ngrains <= 0) return cudaSuccess;
  /* the tile rows that hold lattice rows [xa, xb) */
  const int t0 = max(xa - L.x0, 0) / RTX, t1 = min((xb - 1 - L.x0) / RTX, K.ntx - 1);
  if (t1 < t0) return cudaSuccess;
  bounce_sweep_kernel<real><<<dim3(K.nty, t1 - t0 + 1), 256, 0, s>>>(L, S, A, xa, xb, xlo, xhi, K, D, facc, t0, K.nty);
  return cudaGetLastError();
}
template <typename real>
cudaError_t launch_bounce_end(real *A, const DeferList<real> &D, cudaStream_t s) {
  defer_apply_kernel<real><<<296, 256, 0, s>>>(A, D);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------
 * K1f: forces_fluid (src/main.c:1285-1333) from the swept state.  For a node s of grain i and a
 * link q to a node n not owned by i the reference adds (f_new[s][opp q] + f_new[n][q]) e_{opp q}
 * AFTER streaming; by the pull identity these are A[n][opp q] and A[s][q] before streaming.
 * ---------------------------------------------------------------------------------------- */
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

/* one thread per (listed node, link) for the links into NON-fluid foreign neighbours (other grains,
 * the wall ring); every link is rounded to 64-bit fixed point before it is
 * added, so the result is independent of the order of the adds and of the strip decomposition.
 * The eight lanes of a node are summed with shuffles, then one lane adds to the grain's sums. */
/* the tile's own segment of the boundary-node list, eight lanes per node: links into NON-fluid foreign neighbours
 * (other grains, the wall ring); every link is rounded to 64-bit fixed point before it is added, so the result is
 * independent of the order of the adds and of the strip decomposition */
template <typename real>
__device__ __forceinline__ void force_tile_nodes(const Lattice<real> &L, const Stored<real> &S, int xlo, int xhi,
                                                 const BoundaryList &B, long long *facc, int tile, int *range_flag) {
  const int n = L.ngrains;
  const int items = 8 * B.tcount[tile];
  if (items == 0) return;
  const uint2 *seg = B.entry + (size_t)tile * B.cap;
  const int padded = (items + 31) & ~31; /* whole warps stay in the loop: shuffles below */
  for (int u = threadIdx.x; u < padded; u += blockDim.x) {
    long long s1 = 0, s2 = 0, s3 = 0;
    int i = -1;
    if (u < items) {
      const uint2 en = seg[u >> 3];
      const int q = 1 + (int)(u & 7);
      const int row = (int)(en.x / (unsigned)L.pitch);
      const int x = L.x0 + row, y = (int)(en.x - (unsigned)row * (unsigned)L.pitch);
      if (x >= xlo && x < xhi) {
        i = (int)(en.y & BL_GRAIN);
        const size_t kn = node_index(L, x + ex_of(q), y + ey_of(q));
        /* links into fluid neighbours were added by the sweep kernel */
        const int cn = S.cell[kn];
        if (((en.y >> 24) & (1u << (q - 1))) && !cell_is_fluid(cn)) {
          /* Both populations belong to links into NON-fluid neighbours.  Next to the wall ring the sweep gives such a
           * link of an active node the rest value (a listed w-link, possibly of another tile and still to be written
           * when this runs inside rim_kernel): take that value directly.  Everywhere else nothing in this launch
           * writes them (the fused kernel has put the rest value there already). */
          const int nx = x + ex_of(q), ny = y + ey_of(q), oq = opp_of(q);
          const real fn = (cell_is_act(cn) && !is_ring(L, nx, ny) && !w_links_with_collide(L, nx, ny)) ? L.w[oq] : S.A[oq * L.plane + kn];
          const real fs = ((en.y & BL_ACT) && !w_links_with_collide(L, x, y)) ? L.w[q] : S.A[q * L.plane + en.x];
          real h1 = 0, h2 = 0, h3 = 0;
          force_link<real>(q, fn, fs, x, y, S.grains[i].xc, S.grains[i].yc, &h1, &h2, &h3);
          if (!(fabs((double)h1) < 1024. && fabs((double)h2) < 1024. && fabs((double)h3) < 16384.)) *(volatile int *)range_flag = 1;
          s1 = __double2ll_rn((double)h1 * FORCE_FIX);
          s2 = __double2ll_rn((double)h2 * FORCE_FIX);
          s3 = __double2ll_rn((double)h3 * TORQUE_FIX);
        }
      }
    }
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      s2 += __shfl_xor_sync(0xffffffffu, s2, d);
      s3 += __shfl_xor_sync(0xffffffffu, s3, d);
    }
    if ((u & 7) == 0 && i >= 0) {
      atomicAdd((unsigned long long *)&facc[i], (unsigned long long)s1);
      atomicAdd((unsigned long long *)&facc[n + i], (unsigned long long)s2);
      atomicAdd((unsigned long long *)&facc[2 * n + i], (unsigned long long)s3);
    }
  }
}

template <typename real>
__global__ void __launch_bounds__(256) force_links_kernel(const __grid_constant__ Lattice<real> L,
                                                          const __grid_constant__ Stored<real> S, int xlo, int xhi,
                                                          const BoundaryList B, long long *facc, real *A,
                                                          const DeferList<real> D) {
  /* first the deferred bounce-back links of the sweep (defer_apply_kernel's job, one launch less).  They are links
   * into FLUID neighbours; the force links below read populations of links into NON-fluid neighbours only (A[q][s]
   * with s + e_q not fluid, A[opp q][n] with n + e_opp q = s solid): never the same location. */
  const int cta = blockIdx.y * gridDim.x + blockIdx.x, nctas = gridDim.x * gridDim.y;
  if (A != nullptr) {
    const int nd = min(*D.count, D.capacity);
    for (int k = cta * blockDim.x + threadIdx.x; k < nd; k += nctas * blockDim.x) A[D.index[k]] = D.value[k];
  }
  force_tile_nodes<real>(L, S, xlo, xhi, B, facc, cta, D.range_flag); /* one CTA per lattice tile */
}

/* Single-GPU form of sweep 4 + forces_fluid: ONE launch, one CTA per lattice tile -- the tile's bounce-back links
 * (with their momentum exchange), then its boundary nodes; the CTAs count themselves out and the last one applies
 * the deferred links (they read pre-sweep values: only after EVERY link has been evaluated). */
template <typename real>
__global__ void __launch_bounds__(256, LBMDEM_SWEEP_MINB) rim_kernel(const __grid_constant__ Lattice<real> L,
                                                                    const __grid_constant__ Stored<real> S, real *A, int xa,
                                                                    int xb, int xlo, int xhi, const LinkList K,
                                                                    const BoundaryList B, const DeferList<real> D,
                                                                    long long *facc, int *ticket, int apply_here) {
  __shared__ int s_done;
  const int tile = blockIdx.y * gridDim.x + blockIdx.x, nctas = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) s_done = 0;
  __syncthreads();
  bounce_tile_links<real>(L, S, A, xa, xb, xlo, xhi, K, D, facc, tile);
  if (facc != nullptr) force_tile_nodes<real>(L, S, xlo, xhi, B, facc, tile, D.range_flag);
  /* no CTA barrier at the end (a quarter of the warp samples sat at one): the warps count themselves out in shared
   * memory, the CTA's last warp counts the CTA out, and the last warp of the grid applies the deferred links */
  __syncwarp();
  int last = 0;
  if ((threadIdx.x & 31) == 0) {
    __threadfence();
    if (atomicAdd(&s_done, 1) == (int)(blockDim.x >> 5) - 1) last = atomicAdd(ticket, 1) == nctas - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  const int nd = min(*(volatile int *)D.count, D.capacity);
  /* a handful of deferred links (the usual case: grains of 15+ nodes radius) are applied here; a packing of small
   * grains has them by the hundred thousand (one-node gaps everywhere), and the host -- told the count through
   * mapped memory -- then follows this launch with defer_apply_kernel (apply_here == 0) */
  if (apply_here)
    for (int k = threadIdx.x & 31; k < nd; k += 32) A[__ldcg(&D.index[k])] = __ldcg(&D.value[k]);
  if ((threadIdx.x & 31) == 0) {
    *ticket = 0;
    *(volatile int *)D.seen = nd;
  }
}
template <typename real>
cudaError_t launch_rim(const Lattice<real> &L, const Stored<real> &S, real *A, int xa, int xb, int xlo, int xhi,
                       const LinkList &K, const BoundaryList &B, const DeferList<real> &D, long long *facc, int *ticket,
                       int apply_here, cudaStream_t s) {
  rim_kernel<real><<<dim3(K.nty, K.ntx), 256, 0, s>>>(L, S, A, xa, xb, xlo, xhi, K, B, D, facc, ticket, apply_here);
  if (!apply_here) defer_apply_kernel<real><<<296, 256, 0, s>>>(A, D);
  return cudaGetLastError();
}

template <typename real>
cudaError_t launch_force_links(const Lattice<real> &L, const Stored<real> &S, int xlo, int xhi, const BoundaryList &B,
                               long long *facc, real *A, const DeferList<real> &D, cudaStream_t s) {
  force_links_kernel<real><<<dim3(B.nty, B.ntx), 256, 0, s>>>(L, S, xlo, xhi, B, facc, A, D); /* adds to what the sweep kernel left */
  return cudaGetLastError();
}

/* fhf of one grain from the fixed-point sums (default build) or the fp64 sums in the reference's order (strict
 * build), scaled as src/main.c:1329-1331 */
template <typename real>
__device__ __forceinline__ void force_finish_one(const ForceFinish &fin, int n, int i, real *f1, real *f2, real *f3) {
  real h1, h2, h3;
  if (fin.fixed_point) {
    const long long *facc = static_cast<const long long *>(fin.sums);
    /* the 64-bit sums hold |fhf1|, |fhf2| < 2^11 and |fhf3| < 2^15 (lattice units): a sum in the top quarter of the
     * range has wrapped or is about to -- populations that diverged, or NaN (which converts to the most negative
     * value).  Reported instead of carried on silently. */
    const long long lim = 1ll << 61;
    if (fin.range_flag != nullptr && (llabs(facc[i]) >= lim || llabs(facc[n + i]) >= lim || llabs(facc[2 * n + i]) >= lim))
      *(volatile int *)fin.range_flag = 1; /* mapped host memory */
    h1 = (real)((double)facc[i] / FORCE_FIX);
    h2 = (real)((double)facc[n + i] / FORCE_FIX);
    h3 = (real)((double)facc[2 * n + i] / TORQUE_FIX);
  } else {
    const double *partial = static_cast<const double *>(fin.sums);
    h1 = (real)partial[i]; h2 = (real)partial[n + i]; h3 = (real)partial[2 * n + i];
  }
  *f1 = (real)((double)h1 * fin.k12);
  *f2 = (real)((double)h2 * fin.k12);
  *f3 = (real)((double)h3 * fin.k3);
}
template <typename real>
__global__ void force_finish_kernel(ForceFinish fin, int n, real *fhf1, real *fhf2, real *fhf3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  real f1, f2, f3;
  force_finish_one<real>(fin, n, i, &f1, &f2, &f3);
  fhf1[i] = f1; fhf2[i] = f2; fhf3[i] = f3;
}
template <typename real>
cudaError_t launch_force_finish(const ForceFinish &fin, int n, real *fhf1, real *fhf2, real *fhf3, cudaStream_t s) {
  force_finish_kernel<real><<<(n + 127) / 128, 128, 0, s>>>(fin, n, fhf1, fhf2, fhf3);
  return cudaGetLastError();
}

/* forces_fluid exactly as written (:1295-1325): the three sums of a grain are accumulated in the reference's order --
 * x outer, y inner, q inner -- which is what makes the strict build bit-identical.  One WARP per grain: the lanes load
 * the populations of 32 consecutive y of a row in parallel (the expensive part), then every lane replays the additions
 * in order, taking each term from the lane that holds it. */
template <typename real>
__global__ void __launch_bounds__(128) force_serial_kernel(const __grid_constant__ Lattice<real> L,
                                                           const __grid_constant__ Stored<real> S, int xlo, int xhi,
                                                           double *partial) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n = L.ngrains;
  if (i >= n) return;
  const GrainBox b = S.boxes[i];
  const real xc = S.grains[i].xc, yc = S.grains[i].yc;
  real h1 = 0, h2 = 0, h3 = 0;
  for (int x = max(b.xi, xlo); x <= min(b.xf, xhi - 1); ++x)
    for (int yb = b.yi; yb <= b.yf; yb += 32) {
      const int y = yb + lane;
      unsigned mask = 0; /* bit q-1: link q of node (x, y) leaves the grain */
      real t[NQ];        /* f_new[s][opp q] + f_new[n][q] of those links */
#pragma unroll
      for (int q = 1; q < NQ; ++q) t[q] = 0;
      if (y <= b.yf) {
        const size_t k = node_index(L, x, y);
        const int c = S.cell[k];
        /* a node with a neighbour outside the grain carries the act bit (fluid neighbour) or the rim bit (another
         * grain, the wall ring): the others have no link to add and their neighbours are not looked at */
        if (cell_obst(c) == i && (!S.act_folded || (c & (CELL_ACT | CELL_RIM)))) {
#pragma unroll
          for (int q = 1; q < NQ; ++q) {
            const size_t kn = node_index(L, x + ex_of(q), y + ey_of(q));
            if (cell_obst(S.cell[kn]) != i) {
              mask |= 1u << (q - 1);
              t[q] = S.A[opp_of(q) * L.plane + kn] + S.A[q * L.plane + k];
            }
          }
        }
      }
      unsigned nodes = __ballot_sync(0xffffffffu, mask != 0);
      while (nodes) {
        const int src = __ffs(nodes) - 1;
        nodes &= nodes - 1;
        const unsigned m = __shfl_sync(0xffffffffu, mask, src);
        const int ys = yb + src;
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
          const real v = __shfl_sync(0xffffffffu, t[q], src);
          if ((m >> (q - 1)) & 1u) { /* force_link (lbm_node.cuh) with the sum of the two populations already formed */
            const int oq = opp_of(q);
            const real fnx = v * ex_of(oq);
            const real fny = v * ey_of(oq);
            h1 = h1 + fnx;
            h2 = h2 + fny;
            h3 = h3 - fnx * (ys - yc) + fny * (x - xc);
          }
        }
      }
    }
  if (lane == 0) { partial[i] = h1; partial[n + i] = h2; partial[2 * n + i] = h3; }
}
template <typename real>
cudaError_t launch_force_serial(const Lattice<real> &L, const Stored<real> &S, int xlo, int xhi, double *partial,
                                cudaStream_t s) {
  force_serial_kernel<real><<<(L.ngrains + 3) / 4, 128, 0, s>>>(L, S, xlo, xhi, partial);
  return cudaGetLastError();
}
/* ------------------------------------------------------------------------------------------
 * K3: Verlet lists from a hashed uniform cell list.  The reference builds a half list with an
 * O(N^2) double loop (initVerlet, :1519-1543); here every grain gets its FULL list (both
 * directions), sorted by neighbour index, which is the order in which the reference's scatter
 * loop adds contributions to that grain.
 * ---------------------------------------------------------------------------------------- */
__device__ __forceinline__ unsigned cell_hash(int cx, int cy) {
  return ((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u);
}

template <typename real>
__global__ void verlet_cells_kernel(int n, const real *x1, const real *x2, real cell_size, VerletBuffers vb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = (int)floor((double)x1[i] / (double)cell_size), cy = (int)floor((double)x2[i] / (double)cell_size);
  vb.gcx[i] = cx;
  vb.gcy[i] = cy;
  atomicAdd(&vb.bucket_count[cell_hash(cx, cy) & (vb.nbuckets - 1)], 1);
}

/* exclusive scan of bucket_count[0..nbuckets] in place, one CTA */
__global__ void verlet_scan_kernel(VerletBuffers vb) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  const int total = vb.nbuckets + 1;
  for (int base = 0; base < total; base += blockDim.x) {
    const int idx = base + tid;
    const int v = (idx < vb.nbuckets) ? vb.bucket_count[idx] : 0;
    int incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int ws = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, d);
        if (lane >= d) ws += t;
      }
      warp_sums[lane] = ws; /* inclusive over warps */
    }
    __syncthreads();
    const int before = carry + (wid ? warp_sums[wid - 1] : 0) + incl - v;
    if (idx < total) vb.bucket_count[idx] = before;
    __syncthreads();
    if (tid == 0) carry += warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
}

__global__ void verlet_fill_kernel(int n, VerletBuffers vb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned b = cell_hash(vb.gcx[i], vb.gcy[i]) & (vb.nbuckets - 1);
  vb.sorted[vb.bucket_count[b] + atomicAdd(&vb.bucket_cursor[b], 1)] = i;
}

template <typename real>
__global__ void verlet_lists_kernel(dem::Params<real> P, int n, const real *x1, const real *x2, const real *r,
                                    VerletBuffers vb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = vb.gcx[i], cy = vb.gcy[i];
  const real xi1 = x1[i], xi2 = x2[i], ri = r[i];
  int *lst = vb.nbr + (size_t)i * vb.cap;
  int cnt = 0;
  bool overflow = false;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      const int tx = cx + dx, ty = cy + dy;
      const unsigned b = cell_hash(tx, ty) & (vb.nbuckets - 1);
      for (int k = vb.bucket_count[b]; k < vb.bucket_count[b + 1]; ++k) {
        const int j = vb.sorted[k];
        if (j == i || vb.gcx[j] != tx || vb.gcy[j] != ty) continue;
        /* the reference evaluates the criterion with the lower index first (:1525-1532) */
        const bool in = (i < j) ? dem::verlet_pair(P, xi1, xi2, ri, x1[j], x2[j], r[j])
                                : dem::verlet_pair(P, x1[j], x2[j], r[j], xi1, xi2, ri);
        if (!in) continue;
        if (cnt >= vb.cap) { overflow = true; continue; }
        int p = cnt++; /* insertion sort, ascending */
        while (p > 0 && lst[p - 1] > j) { lst[p] = lst[p - 1]; --p; }
        lst[p] = j;
      }
    }
  vb.nbr_count[i] = cnt;
  vb.wflags[i] = dem::wall_flags(P, xi1, xi2, ri);
  if (overflow) *(volatile int *)vb.error = 1; /* mapped host memory */
}

template <typename real>
cudaError_t launch_verlet(const dem::Params<real> &P, int n, const GrainArrays<real> &g, real cell_size,
                          const VerletBuffers &vb, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(vb.bucket_count, 0, sizeof(int) * (vb.nbuckets + 1), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(vb.bucket_cursor, 0, sizeof(int) * vb.nbuckets, s);
  if (e != cudaSuccess) return e;
  const int nb = (n + 127) / 128;
  verlet_cells_kernel<real><<<nb, 128, 0, s>>>(n, g.x1, g.x2, cell_size, vb);
  verlet_scan_kernel<<<1, 1024, 0, s>>>(vb);
  verlet_fill_kernel<<<nb, 128, 0, s>>>(n, vb);
  verlet_lists_kernel<real><<<nb, 128, 0, s>>>(P, n, g.x1, g.x2, g.r, vb);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------
 * K4: one DEM step = kick-drift (:1748-1753), forces (:1336-1516), kick (:1758-1763)
 * ---------------------------------------------------------------------------------------- */
template <typename real>
__global__ void dem_kick_drift_kernel(dem::Params<real> P, int n, GrainArrays<real> g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  real x, v;
  x = g.x1[i]; v = g.v1[i]; dem::kick_drift(P, &x, &v, g.a1[i]); g.x1[i] = x; g.v1[i] = v;
  x = g.x2[i]; v = g.v2[i]; dem::kick_drift(P, &x, &v, g.a2[i]); g.x2[i] = x; g.v2[i] = v;
  x = g.x3[i]; v = g.v3[i]; dem::kick_drift(P, &x, &v, g.a3[i]); g.x3[i] = x; g.v3[i] = v;
}

/* One warp per grain.  Lane k evaluates the contact with the k-th neighbour (the expensive
 * part: sqrt, divisions); the contributions are then added in neighbour order by every lane
 * redundantly, which is exactly the order in which the reference's half-list loop touches
 * this grain: lower-index partners first (their outer-loop turn comes earlier), then its own
 * list, both ascending (:1434-1450).  Walls follow in the order B, T, L, R (:1455-1508). */
template <typename real>
__global__ void dem_forces_kernel(dem::Params<real> P, int n, bool film, GrainArrays<real> g, VerletBuffers vb) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const real xi1 = g.x1[i], xi2 = g.x2[i], vi1 = g.v1[i], vi2 = g.v2[i], vi3 = g.v3[i], ri = g.r[i];
  real a1 = g.fhf1[i], a2 = g.fhf2[i], a3 = g.fhf3[i];
  const int cnt = vb.nbr_count[i];
  for (int base = 0; base < cnt; base += 32) {
    const int k = base + lane;
    real c1 = 0, c2 = 0, c3 = 0;
    bool touch = false;
    if (k < cnt) {
      const int j = vb.nbr[(size_t)i * vb.cap + k];
      const real xj1 = g.x1[j], xj2 = g.x2[j], vj1 = g.v1[j], vj2 = g.v2[j], vj3 = g.v3[j], rj = g.r[j];
      dem::Force<real> F;
      if (i < j) {
        touch = dem::pair_force(P, film, xi1, xi2, vi1, vi2, vi3, ri, xj1, xj2, vj1, vj2, vj3, rj, &F);
        if (touch) { c1 = F.f1; c2 = F.f2; c3 = F.f3; }
      } else {
        touch = dem::pair_force(P, film, xj1, xj2, vj1, vj2, vj3, rj, xi1, xi2, vi1, vi2, vi3, ri, &F);
        if (touch) { c1 = -F.f1; c2 = -F.f2; c3 = F.f3; }
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, touch);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      a1 = a1 + __shfl_sync(0xffffffffu, c1, src);
      a2 = a2 + __shfl_sync(0xffffffffu, c2, src);
      a3 = a3 + __shfl_sync(0xffffffffu, c3, src);
    }
  }
  if (lane != 0) return;
  dem::add_wall_forces(P, vb.wflags[i], xi1, xi2, vi1, vi2, vi3, ri, &a1, &a2, &a3);
  dem::finish_acceleration(P, g.m[i], g.It[i], &a1, &a2, &a3);
  g.a1[i] = a1; g.a2[i] = a2; g.a3[i] = a3;
}

template <typename real>
__global__ void dem_kick_kernel(dem::Params<real> P, int n, GrainArrays<real> g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  real v;
  v = g.v1[i]; dem::kick(P, &v, g.a1[i]); g.v1[i] = v;
  v = g.v2[i]; dem::kick(P, &v, g.a2[i]); g.v2[i] = v;
  v = g.v3[i]; dem::kick(P, &v, g.a3[i]); g.v3[i] = v;
}

/* the closing kick of one sub-step (:1758-1763) and the kick-drift that opens the next (:1748-1753): the same
 * operations on the same grain in the same order, one launch instead of two */
template <typename real>
__global__ void dem_kick_kick_drift_kernel(dem::Params<real> P, int n, GrainArrays<real> g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  real x, v, a;
  a = g.a1[i]; v = g.v1[i]; dem::kick(P, &v, a); x = g.x1[i]; dem::kick_drift(P, &x, &v, a); g.x1[i] = x; g.v1[i] = v;
  a = g.a2[i]; v = g.v2[i]; dem::kick(P, &v, a); x = g.x2[i]; dem::kick_drift(P, &x, &v, a); g.x2[i] = x; g.v2[i] = v;
  a = g.a3[i]; v = g.v3[i]; dem::kick(P, &v, a); x = g.x3[i]; dem::kick_drift(P, &x, &v, a); g.x3[i] = x; g.v3[i] = v;
}

template <typename real>
cudaError_t launch_dem_step(const dem::Params<real> &P, int n, bool film, const GrainArrays<real> &g,
                            const VerletBuffers &vb, real *mid, bool drift_done, bool drift_next, cudaStream_t s) {
  const int nb = (n + 127) / 128;
  if (!drift_done) dem_kick_drift_kernel<real><<<nb, 128, 0, s>>>(P, n, g);
  if (mid != nullptr) {
    const real *src[6] = {g.x1, g.x2, g.x3, g.v1, g.v2, g.v3};
    for (int k = 0; k < 6; ++k) {
      cudaError_t e = cudaMemcpyAsync(mid + (size_t)k * n, src[k], sizeof(real) * n, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return e;
    }
  }
  dem_forces_kernel<real><<<(n * 32 + 127) / 128, 128, 0, s>>>(P, n, film, g, vb);
  if (drift_next) dem_kick_kick_drift_kernel<real><<<nb, 128, 0, s>>>(P, n, g);
  else dem_kick_kernel<real><<<nb, 128, 0, s>>>(P, n, g);
  return cudaGetLastError();
}

/* The DEM sub-steps between two LBM steps in ONE launch, thread i = grain i: kick-drift, barrier, forces (gather over the
 * sorted full list), barrier, kick -- the barriers standing where the reference's loops end.  Up to 1024 grains: one
 * thread-block CLUSTER of 8 CTAs of 128 threads (8 SMs share the fp64 contact arithmetic -- a single CTA spent 8 us per
 * sub-step on one SM's fp64 pipe -- and meet at the hardware cluster barrier); above: a cooperative grid with grid
 * barriers.  Same arithmetic in the same order (neighbour contributions added in list order), hence the same bits as
 * the three-launch form.
 * fin.sums != nullptr: the launch follows an LBM step and first turns the fixed-point force sums into fhf
 * (force_finish_kernel's job).  film_first: the first sub-step is a film step (alternate contact law, :1342-1426). */
template <typename real, bool CLUSTER>
__global__ void __launch_bounds__(DEM_COOP_THREADS) dem_coop_kernel(dem::Params<real> P, int n, int nsub, bool film_first,
                                                                    GrainArrays<real> g, VerletBuffers vb, ForceFinish fin) {
  auto everybody = [] {
    if constexpr (CLUSTER) cg::this_cluster().sync();
    else cg::this_grid().sync();
  };
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n;
  real x1 = 0, x2 = 0, x3 = 0, v1 = 0, v2 = 0, v3 = 0, a1 = 0, a2 = 0, a3 = 0, ri = 0, mi = 1, Iti = 1, f1 = 0, f2 = 0, f3 = 0;
  int cnt = 0, wfl = 0;
  if (on) {
    x1 = g.x1[i]; x2 = g.x2[i]; x3 = g.x3[i]; v1 = g.v1[i]; v2 = g.v2[i]; v3 = g.v3[i];
    a1 = g.a1[i]; a2 = g.a2[i]; a3 = g.a3[i]; ri = g.r[i]; mi = g.m[i]; Iti = g.It[i];
    if (fin.sums != nullptr) {
      force_finish_one<real>(fin, n, i, &f1, &f2, &f3);
      g.fhf1[i] = f1; g.fhf2[i] = f2; g.fhf3[i] = f3;
    } else {
      f1 = g.fhf1[i]; f2 = g.fhf2[i]; f3 = g.fhf3[i];
    }
    cnt = vb.nbr_count[i]; wfl = vb.wflags[i];
  }
  for (int s = 0; s < nsub; ++s) {
    const bool film = film_first && s == 0;
    if (on) {
      dem::kick_drift(P, &x1, &v1, a1);
      dem::kick_drift(P, &x2, &v2, a2);
      dem::kick_drift(P, &x3, &v3, a3);
      g.x1[i] = x1; g.x2[i] = x2; g.v1[i] = v1; g.v2[i] = v2; g.v3[i] = v3; /* what the neighbours read */
    }
    everybody();
    if (on) {
      a1 = f1; a2 = f2; a3 = f3;
      for (int k = 0; k < cnt; ++k) {
        const int j = vb.nbr[(size_t)i * vb.cap + k];
        const real xj1 = g.x1[j], xj2 = g.x2[j], vj1 = g.v1[j], vj2 = g.v2[j], vj3 = g.v3[j], rj = g.r[j];
        dem::Force<real> F;
        if (i < j) {
          if (dem::pair_force(P, film, x1, x2, v1, v2, v3, ri, xj1, xj2, vj1, vj2, vj3, rj, &F)) {
            a1 = a1 + F.f1; a2 = a2 + F.f2; a3 = a3 + F.f3;
          }
        } else {
          if (dem::pair_force(P, film, xj1, xj2, vj1, vj2, vj3, rj, x1, x2, v1, v2, v3, ri, &F)) {
            a1 = a1 + (-F.f1); a2 = a2 + (-F.f2); a3 = a3 + F.f3;
          }
        }
      }
      dem::add_wall_forces(P, wfl, x1, x2, v1, v2, v3, ri, &a1, &a2, &a3);
      dem::finish_acceleration(P, mi, Iti, &a1, &a2, &a3);
    }
    everybody(); /* everybody has read the mid-step velocities before they move on */
    if (on) {
      dem::kick(P, &v1, a1);
      dem::kick(P, &v2, a2);
      dem::kick(P, &v3, a3);
    }
  }
  if (on) {
    g.x3[i] = x3; g.v1[i] = v1; g.v2[i] = v2; g.v3[i] = v3; g.a1[i] = a1; g.a2[i] = a2; g.a3[i] = a3;
  }
}
template <typename real>
cudaError_t launch_dem_coop(const dem::Params<real> &P, int n, int nsub, bool film_first, const GrainArrays<real> &g,
                            const VerletBuffers &vb, const ForceFinish &fin, cudaStream_t s) {
  if (n <= DEM_CLUSTER_MAX) { /* one cluster of 8 CTAs */
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(DEM_CLUSTER_MAX / DEM_COOP_THREADS);
    cfg.blockDim = dim3(DEM_COOP_THREADS);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = DEM_CLUSTER_MAX / DEM_COOP_THREADS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, dem_coop_kernel<real, true>, P, n, nsub, film_first, g, vb, fin);
  }
  dem::Params<real> Pc = P;
  GrainArrays<real> gc = g;
  VerletBuffers vc = vb;
  ForceFinish fc = fin;
  void *args[] = {&Pc, &n, &nsub, &film_first, &gc, &vc, &fc};
  const dim3 grid((n + DEM_COOP_THREADS - 1) / DEM_COOP_THREADS), block(DEM_COOP_THREADS);
  return cudaLaunchCooperativeKernel((const void *)dem_coop_kernel<real, false>, grid, block, args, 0, s);
}
template <typename real>
cudaError_t dem_coop_capacity(int *max_grains) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dem_coop_kernel<real, false>, DEM_COOP_THREADS, 0)) != cudaSuccess) return e;
  *max_grains = sms * per_sm * DEM_COOP_THREADS;
  return cudaSuccess;
}

/* In-process strip groups (sim.cu, LocalGroup): the force sums of the ranks are added by ONE kernel per rank that reads
 * every peer's partial sums from that peer's device memory (same device, or NVLink peer access) -- no collective
 * library.  Integer sums are exact in any order; the fp64 sums of the strict build are added in rank order. */
template <typename T>
__global__ void peer_sum_kernel(PeerPtrs pp, int len, T *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  T s = 0;
  for (int k = 0; k < pp.count; ++k) s += static_cast<const T *>(pp.p[k])[i];
  out[i] = s;
}
template <typename T>
cudaError_t launch_peer_sum(const PeerPtrs &pp, int len, T *out, cudaStream_t s) {
  peer_sum_kernel<T><<<(len + 255) / 256, 256, 0, s>>>(pp, len, out);
  return cudaGetLastError();
}
/* ---- the same across PROCESSES (one rank per GPU): peers' sums through CUDA IPC mappings ----
 * publish: runs behind this rank's last force kernel; tells every peer (a word in the PEER's memory) that this rank's
 * partial sums of step `step` are complete.  sum: waits until every peer has said so, then adds, per grain, the sums of
 * the ranks whose rows the grain's bounding box touches (+- one row) -- every force contribution comes from a solid node
 * of the grain in an owned row, so the other ranks hold zeros there -- which keeps the NVLink traffic at about one copy
 * of the sums instead of nranks - 1.  Integer adds: the result does not depend on who adds in which order. */
__global__ void ipc_publish_kernel(IpcPeers pp, unsigned step) {
  const int p = threadIdx.x;
  if (p < pp.nranks && p != pp.rank) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pp.flags[p] + pp.rank), "r"(step) : "memory");
  }
}
__global__ void __launch_bounds__(256) ipc_sum_kernel(IpcPeers pp, unsigned step, int n, const GrainBox *boxes, long long *out,
                                                      int *timeout_flag) {
  if ((int)threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank) {
    const unsigned *fl = pp.flags[pp.rank] + threadIdx.x;
    const long long t0 = clock64();
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(fl) : "memory");
    } while ((int)(v - step) < 0 && clock64() - t0 < 40000000000ll); /* ~20 s: a peer is gone, not late */
    if ((int)(v - step) < 0) *(volatile int *)timeout_flag = 1;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GrainBox b = boxes[i];
  long long s1 = 0, s2 = 0, s3 = 0;
  const int base = pp.lx / pp.nranks, rem = pp.lx % pp.nranks;
  for (int k = 0; k < pp.nranks; ++k) {
    const int klo = k * base + min(k, rem), khi = klo + base + (k < rem ? 1 : 0); /* rank k owns rows [klo, khi) */
    if (k == pp.rank || (b.xf + 1 >= klo && b.xi - 1 < khi && b.xf >= b.xi)) {
      const long long *f = pp.facc[k];
      s1 += __ldcv(f + i);
      s2 += __ldcv(f + n + i);
      s3 += __ldcv(f + 2 * (size_t)n + i);
    }
  }
  out[i] = s1;
  out[n + i] = s2;
  out[2 * (size_t)n + i] = s3;
}
cudaError_t launch_ipc_publish(const IpcPeers &pp, unsigned step, cudaStream_t s) {
  ipc_publish_kernel<<<1, 32, 0, s>>>(pp, step);
  return cudaGetLastError();
}
cudaError_t launch_ipc_sum(const IpcPeers &pp, unsigned step, int n, const GrainBox *boxes, long long *out, int *timeout_flag,
                           cudaStream_t s) {
  ipc_sum_kernel<<<(n + 255) / 256, 256, 0, s>>>(pp, step, n, boxes, out, timeout_flag);
  return cudaGetLastError();
}

template cudaError_t launch_peer_sum<long long>(const PeerPtrs &, int, long long *, cudaStream_t);
template cudaError_t launch_peer_sum<double>(const PeerPtrs &, int, double *, cudaStream_t);

/* ------------------------------------------------------------------------------------------
 * K5: check_density / final_density (src/main.c:1249-1273), fixed-shape two-stage sum in fp64
 * ---------------------------------------------------------------------------------------- */
template <typename real>
__global__ void density_partial_kernel(const real *f, int ly, int x0, int xlo, int xhi, int pitch, size_t plane,
                                       double *partials) {
  __shared__ double sh[256];
  const size_t per_plane = (size_t)(xhi - xlo) * ly;
  const size_t total = per_plane * NQ;
  double acc = 0;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int q = (int)(t / per_plane);
    const size_t rem = t - (size_t)q * per_plane;
    const int row = (int)(rem / ly), y = (int)(rem - (size_t)row * ly);
    acc += (double)f[q * plane + (size_t)(xlo - x0 + row) * pitch + y];
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}
__global__ void density_final_kernel(const double *partials, int np, double *out) {
  double acc = 0;
  for (int k = 0; k < np; ++k) acc += partials[k];
  *out = acc;
}
template <typename real>
cudaError_t launch_density(const real *f, int ly, int x0, int xlo, int xhi, int pitch, size_t plane, double *partials,
                           int npartials, double *out, cudaStream_t s) {
  density_partial_kernel<real><<<npartials, 256, 0, s>>>(f, ly, x0, xlo, xhi, pitch, plane, partials);
  density_final_kernel<<<1, 1, 0, s>>>(partials, npartials, out);
  return cudaGetLastError();
}

/* Exact, order-free fingerprint of the lattice state of the owned rows: sum mod 2^64 of
 *   bits(f[x][y][q] widened to double) * (2 k + 1),  k = (x * ly + y) * 9 + q      -> out[0]
 *   (obst[x][y] + 2) * (2 (x * ly + y) + 1)                                        -> out[1]
 * with GLOBAL coordinates: integer adds commute, so the fingerprints of the strips of a decomposed run add up
 * (mod 2^64) to the one-GPU value if and only if every population and node index is the same (up to 2^-64 odds). */
template <typename real>
__global__ void __launch_bounds__(256) checksum_kernel(const real *f, const int *cell, int ly, int x0, int xlo, int xhi,
                                                       int pitch, size_t plane, unsigned long long *out) {
  unsigned long long a = 0, b = 0;
  const size_t nodes = (size_t)(xhi - xlo) * ly;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nodes; t += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(t / ly), y = (int)(t - (size_t)row * ly);
    const size_t k = (size_t)(xlo - x0 + row) * pitch + y;
    const unsigned long long gnode = (unsigned long long)(xlo + row) * (unsigned long long)ly + (unsigned long long)y;
    b += (unsigned long long)(long long)(cell_obst(cell[k]) + 2) * (2ull * gnode + 1ull);
#pragma unroll
    for (int q = 0; q < NQ; ++q)
      a += (unsigned long long)__double_as_longlong((double)f[q * plane + k]) * (2ull * (gnode * NQ + q) + 1ull);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d);
    b += __shfl_xor_sync(0xffffffffu, b, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], a);
    atomicAdd(&out[1], b);
  }
}
template <typename real>
cudaError_t launch_checksum(const real *f, const int *cell, int ly, int x0, int xlo, int xhi, int pitch, size_t plane,
                            unsigned long long *out, int blocks, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  checksum_kernel<real><<<blocks, 256, 0, s>>>(f, cell, ly, x0, xlo, xhi, pitch, plane, out);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------
 * K6: write_vtk's five point fields (src/main.c:284-323) for the owned rows, [y][x] order
 * ---------------------------------------------------------------------------------------- */
template <typename real>
__global__ void fields_kernel(const real *f, const int *cell, GrainArrays<real> g, const real *gp, int n, int ly, int x0,
                              int xlo, int xhi, int pitch, size_t plane, real rho_moy, float *grain_p, float *grain_v,
                              float *grain_a, float *fluid_p, float *fluid_v) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = xlo + blockIdx.y;
  if (y >= ly || x >= xhi) return;
  const size_t k = (size_t)(x - x0) * pitch + y;
  const size_t o = (size_t)y * (xhi - xlo) + (x - xlo);
  const int i = cell_obst(cell[k]);
  if (i >= 0 && i < n) {
    grain_p[o] = gp ? (float)gp[i] : 0.f;
    grain_v[3 * o] = (float)g.v1[i]; grain_v[3 * o + 1] = (float)g.v2[i]; grain_v[3 * o + 2] = 0.f;
    grain_a[3 * o] = (float)g.a1[i]; grain_a[3 * o + 1] = (float)g.a2[i]; grain_a[3 * o + 2] = 0.f;
    fluid_p[o] = 0.f;
    fluid_v[3 * o] = 0.f; fluid_v[3 * o + 1] = 0.f; fluid_v[3 * o + 2] = 0.f;
  } else {
    real p[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) p[q] = f[q * plane + k];
    /* :308-313: the reference accumulates into float fields, one population at a time */
    float fp = 0.f, ux = 0.f, uy = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      fp = (float)(fp + p[q]);
      ux = (float)(ux + p[q] * ex_of(q));
      uy = (float)(uy + p[q] * ey_of(q));
    }
    grain_p[o] = -1.f;
    grain_v[3 * o] = 0.f; grain_v[3 * o + 1] = 0.f; grain_v[3 * o + 2] = 0.f;
    grain_a[3 * o] = 0.f; grain_a[3 * o + 1] = 0.f; grain_a[3 * o + 2] = 0.f;
    fluid_p[o] = (float)((1. / 3.) * rho_moy * (fp - 1.));
    fluid_v[3 * o] = ux; fluid_v[3 * o + 1] = uy; fluid_v[3 * o + 2] = 0.f;
  }
}
template <typename real>
cudaError_t launch_fields(const real *f, const int *cell, const GrainArrays<real> &g, const real *gp, int n, int ly,
                          int x0, int xlo, int xhi, int pitch, size_t plane, real rho_moy, float *grain_p, float *grain_v,
                          float *grain_a, float *fluid_p, float *fluid_v, cudaStream_t s) {
  dim3 grid((ly + 127) / 128, xhi - xlo);
  fields_kernel<real><<<grid, 128, 0, s>>>(f, cell, g, gp, n, ly, x0, xlo, xhi, pitch, plane, rho_moy, grain_p, grain_v,
                                           grain_a, fluid_p, fluid_v);
  return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------
 * layout conversion: reference f[x][y][q] (as double) <-> device f[q][x][y] (real)
 * ---------------------------------------------------------------------------------------- */
template <typename real>
__global__ void f_to_host_kernel(const real *f, int ly, int pitch, size_t plane, int row0, int nrows, double *out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nrows * ly * NQ;
  if (t >= total) return;
  const int q = (int)(t % NQ);
  const size_t node = t / NQ;
  const int row = (int)(node / ly), y = (int)(node - (size_t)row * ly);
  out[t] = (double)f[q * plane + (size_t)(row0 + row) * pitch + y];
}
template <typename real>
__global__ void f_from_host_kernel(real *f, int ly, int pitch, size_t plane, int row0, int nrows, const double *in) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nrows * ly * NQ;
  if (t >= total) return;
  const int q = (int)(t % NQ);
  const size_t node = t / NQ;
  const int row = (int)(node / ly), y = (int)(node - (size_t)row * ly);
  f[q * plane + (size_t)(row0 + row) * pitch + y] = (real)in[t];
}
template <typename real>
cudaError_t launch_f_to_host_layout(const real *f, int ly, int pitch, size_t plane, int row0, int nrows, double *out,
                                    cudaStream_t s) {
  const size_t total = (size_t)nrows * ly * NQ;
  f_to_host_kernel<real><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(f, ly, pitch, plane, row0, nrows, out);
  return cudaGetLastError();
}
template <typename real>
cudaError_t launch_f_from_host_layout(real *f, int ly, int pitch, size_t plane, int row0, int nrows, const double *in,
                                      cudaStream_t s) {
  const size_t total = (size_t)nrows * ly * NQ;
  f_from_host_kernel<real><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(f, ly, pitch, plane, row0, nrows, in);
  return cudaGetLastError();
}

template <typename real, typename H>
__global__ void grain_unpack_kernel(const H *rows, int n, int ncols, real *cols) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ncols) return;
  const int i = t / ncols, k = t - i * ncols;
  cols[(size_t)k * n + i] = (real)rows[t];
}
template <typename real, typename H>
__global__ void grain_pack_kernel(const real *cols, int n, int ncols, H *rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * ncols) return;
  const int i = t / ncols, k = t - i * ncols;
  rows[t] = (H)cols[(size_t)k * n + i];
}
/* two groups of columns -> [n][na] followed by [n][nb] (the kinematic state and fhf of lbmdem_step_host) */
template <typename real, typename H>
__global__ void grain_pack2_kernel(const real *cols_a, int na, const real *cols_b, int nb, int n, H *rows) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (na + nb)) return;
  if (t < n * na) {
    const int i = t / na, k = t - i * na;
    rows[t] = (H)cols_a[(size_t)k * n + i];
  } else {
    const int u = t - n * na, i = u / nb, k = u - i * nb;
    rows[t] = (H)cols_b[(size_t)k * n + i];
  }
}
template <typename real>
cudaError_t launch_grain_pack2(const real *cols_a, int na, const real *cols_b, int nb, int n, void *rows, bool rows_f32,
                               cudaStream_t s) {
  const int blocks = (n * (na + nb) + 255) / 256;
  if (rows_f32) grain_pack2_kernel<real, float><<<blocks, 256, 0, s>>>(cols_a, na, cols_b, nb, n, static_cast<float *>(rows));
  else grain_pack2_kernel<real, double><<<blocks, 256, 0, s>>>(cols_a, na, cols_b, nb, n, static_cast<double *>(rows));
  return cudaGetLastError();
}
template <typename real>
cudaError_t launch_grain_unpack(const void *rows, bool rows_f32, int n, int ncols, real *cols, cudaStream_t s) {
  if (rows_f32) grain_unpack_kernel<real, float><<<(n * ncols + 255) / 256, 256, 0, s>>>(static_cast<const float *>(rows), n, ncols, cols);
  else grain_unpack_kernel<real, double><<<(n * ncols + 255) / 256, 256, 0, s>>>(static_cast<const double *>(rows), n, ncols, cols);
  return cudaGetLastError();
}
template <typename real>
cudaError_t launch_grain_pack(const real *cols, int n, int ncols, void *rows, bool rows_f32, cudaStream_t s) {
  if (rows_f32) grain_pack_kernel<real, float><<<(n * ncols + 255) / 256, 256, 0, s>>>(cols, n, ncols, static_cast<float *>(rows));
  else grain_pack_kernel<real, double><<<(n * ncols + 255) / 256, 256, 0, s>>>(cols, n, ncols, static_cast<double *>(rows));
  return cudaGetLastError();
}

/* init_density (src/main.c:716-724): f = w everywhere */
template <typename real>
__global__ void fill_rest_kernel(real *f, size_t plane, Lattice<real> Lw) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= plane) return;
#pragma unroll
  for (int q = 0; q < NQ; ++q) f[q * plane + t] = Lw.w[q];
}
template <typename real>
cudaError_t launch_fill_rest(real *f, size_t plane, const Lattice<real> &Lw, cudaStream_t s) {
  fill_rest_kernel<real><<<(unsigned)((plane + 255) / 256), 256, 0, s>>>(f, plane, Lw);
  return cudaGetLastError();
}

#define INSTANTIATE(real)                                                                                               \
  template cudaError_t launch_raster_tiles<real>(const RasterParams<real> &, int, const GrainArrays<real> &,              \
                                                 GrainRec<real> *, real *, GrainBox *, const GrainRec<real> *,            \
                                                 const real *, const GrainBox *, int *, const int *, unsigned char *,     \
                                                 const unsigned char *, unsigned short *, const unsigned short *, int,    \
                                                 int, int,                                                                \
                                                 const TileBins &, const BoundaryList &, const LinkList &, int *,         \
                                                 long long *, int, int, int, cudaStream_t);                               \
  template cudaError_t launch_act_map<real>(const Lattice<real> &, const Stored<real> &, int, int, int *, cudaStream_t);  \
  template cudaError_t launch_ring_sweep<real>(const Lattice<real> &, const Stored<real> &, real *, int, int,             \
                                               cudaStream_t);                                                             \
  template cudaError_t launch_bounce_pass<real>(const Lattice<real> &, const Stored<real> &, real *, int, int, int, int,  \
                                                const LinkList &, const DeferList<real> &, long long *, cudaStream_t);    \
  template cudaError_t launch_rim<real>(const Lattice<real> &, const Stored<real> &, real *, int, int, int, int,          \
                                        const LinkList &, const BoundaryList &, const DeferList<real> &, long long *,     \
                                        int *, int, cudaStream_t);                                                        \
  template cudaError_t launch_bounce_end<real>(real *, const DeferList<real> &, cudaStream_t);                            \
  template cudaError_t launch_force_links<real>(const Lattice<real> &, const Stored<real> &, int, int,                    \
                                                const BoundaryList &, long long *, real *, const DeferList<real> &,       \
                                                cudaStream_t);                                                            \
  template cudaError_t launch_force_finish<real>(const ForceFinish &, int, real *, real *, real *,                        \
                                                 cudaStream_t);                                                           \
  template cudaError_t launch_force_serial<real>(const Lattice<real> &, const Stored<real> &, int, int, double *,         \
                                                 cudaStream_t);                                                           \
  template cudaError_t launch_verlet<real>(const dem::Params<real> &, int, const GrainArrays<real> &, real,               \
                                           const VerletBuffers &, cudaStream_t);                                         \
  template cudaError_t launch_dem_step<real>(const dem::Params<real> &, int, bool, const GrainArrays<real> &,             \
                                             const VerletBuffers &, real *, bool, bool, cudaStream_t);                   \
  template cudaError_t launch_dem_coop<real>(const dem::Params<real> &, int, int, bool, const GrainArrays<real> &,        \
                                             const VerletBuffers &, const ForceFinish &, cudaStream_t);                  \
  template cudaError_t dem_coop_capacity<real>(int *);                                                                    \
  template cudaError_t launch_density<real>(const real *, int, int, int, int, int, size_t, double *, int, double *,       \
                                            cudaStream_t);                                                                \
  template cudaError_t launch_checksum<real>(const real *, const int *, int, int, int, int, int, size_t,                  \
                                             unsigned long long *, int, cudaStream_t);                                    \
  template cudaError_t launch_fields<real>(const real *, const int *, const GrainArrays<real> &, const real *, int, int,  \
                                           int, int, int, int, size_t, real, float *, float *, float *, float *, float *, \
                                           cudaStream_t);                                                                 \
  template cudaError_t launch_f_to_host_layout<real>(const real *, int, int, size_t, int, int, double *, cudaStream_t);   \
  template cudaError_t launch_f_from_host_layout<real>(real *, int, int, size_t, int, int, const double *, cudaStream_t); \
  template cudaError_t launch_grain_unpack<real>(const void *, bool, int, int, real *, cudaStream_t);                     \
  template cudaError_t launch_grain_pack<real>(const real *, int, int, void *, bool, cudaStream_t);                       \
  template cudaError_t launch_grain_pack2<real>(const real *, int, const real *, int, int, void *, bool, cudaStream_t);   \
  template cudaError_t launch_fill_rest<real>(real *, size_t, const Lattice<real> &, cudaStream_t);
INSTANTIATE(float)
INSTANTIATE(double)

}  // namespace lbmdem
