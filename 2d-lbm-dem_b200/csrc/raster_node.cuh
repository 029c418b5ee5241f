/*
 * raster_node.cuh -- per-grain geometry of obst_construction() (src/main.c:991-1065), written
 * once for host and device.
 *
 * The obstacle map must be BIT-EXACT with the reference, so the disc test keeps the
 * reference's expression `(x-xc)*(x-xc) + (y-yc)*(y-yc) <= r2` in `real` arithmetic and this
 * header is only ever compiled without multiply-add contraction (nvcc -fmad=false, g++
 * -ffp-contract=off).
 *
 * Device-side map encoding ("cell"): -1 fluid, i | CELL_ACT*act for a node owned by grain i,
 * nbgrains for the wall ring (init_obst, :674-687).  The reference rasterises grains in index
 * order and lets later grains overwrite earlier ones, i.e. the owner of a node is the
 * HIGHEST-index grain that covers it: on the device that is an atomicMax.
 */
#pragma once
#include "lbm_node.cuh"

namespace lbm {

/* what the rasteriser keeps per grain besides the GrainRec */
struct GrainBox {
  int xi, xf, yi, yf;     /* clamped bounding box, :1016-1023 (empty when xi>xf or yi>yf) */
};

template <typename real>
struct RasterParams {
  int lx, ly;
  real dx, Mgx, Mby;
};

/* src/main.c:1009-1023 */
template <typename real>
LBM_HD void grain_geometry(const RasterParams<real> &P, real x1, real x2, real r, real rLB, real *xc, real *yc,
                           real *r2, real *R2, GrainBox *b) {
  *xc = (x1 - P.Mgx) / P.dx;
  *yc = (x2 - P.Mby) / P.dx;
  *r2 = rLB * rLB;
  const real rbl0 = r / P.dx;
  *R2 = rbl0 * rbl0;
  int xi = (int)(*xc - rbl0), xf = (int)(*xc + rbl0);
  if (xi < 1) xi = 1;
  if (xf >= P.lx - 1) xf = P.lx - 2;
  int yi = (int)(*yc - rbl0), yf = (int)(*yc + rbl0);
  if (yi < 1) yi = 1;
  if (yf >= P.ly - 1) yf = P.ly - 2;
  b->xi = xi; b->xf = xf; b->yi = yi; b->yf = yf;
}

/* src/main.c:1026-1029 */
template <typename real>
LBM_HD bool disc_covers(real xc, real yc, real r2, real R2, int x, int y) {
  const real dist2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
  return dist2 <= R2 && dist2 <= r2;
}

LBM_HD bool box_has(const GrainBox &b, int x, int y) { return x >= b.xi && x <= b.xf && y >= b.yi && y <= b.yf; }

/* Was node n "fluid" when the reference's grain loop reached grain i (:1047)?  At that moment
 * the map holds grains 0..i only.  n is fluid then iff no grain j <= i covers it: final map
 * -1, or final owner k > i while grain i itself does not cover n.  (A node covered by k > i
 * AND by some j < i but not by i -- three mutually overlapping reduced discs -- would be
 * misjudged; reduced discs are 0.85 r, so even a pair only overlaps at > 15 % interpenetration.)
 */
template <typename real>
LBM_HD bool fluid_when_grain_ran(int cell_n, int i, int ngrains, real xc, real yc, real r2, real R2,
                                 const GrainBox &b, int nx, int ny) {
  if (cell_is_fluid(cell_n)) return true;
  const int k = cell_obst(cell_n);
  if (k >= ngrains || k <= i) return false;
  return !(box_has(b, nx, ny) && disc_covers(xc, yc, r2, R2, nx, ny));
}

}  // namespace lbm
