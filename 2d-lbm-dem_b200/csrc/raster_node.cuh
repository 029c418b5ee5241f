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

}  // namespace lbm
