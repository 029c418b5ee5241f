/*
 * kernels.h -- launch interface between the host-side simulation object (sim.cu) and the
 * kernel translation units.  Internal to liblbmdem_gpu.so; the public boundary is
 * include/lbmdem_gpu.h.
 *
 * Kernel legend:
 *   K1  lbm_step        fused re-init + MRT collide + wall ring + grain bounce-back + pull
 *                       stream + momentum exchange      (src/main.c:966-986, :1071-1243, :1285-1325)
 *   K2  raster          grain records + obstacle map    (src/main.c:991-1065)
 *   K3  verlet          hash-grid cell list -> sorted full neighbour lists + wall flags
 *                                                       (src/main.c:1519-1594)
 *   K4  dem             kick-drift, contact forces, kick (src/main.c:1733-1763, :1336-1516)
 *   K5  density         sum of all populations          (src/main.c:1249-1273)
 *   K6  fields          rho / momentum / grain fields in VTK order (src/main.c:284-323)
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dem_node.cuh"
#include "lbm_node.cuh"
#include "raster_node.cuh"

namespace lbmdem {

/* tile of the fused LBM kernel (nodes); the TMA box adds a one-node halo on every side */
constexpr int TILE_X = 16;
constexpr int TILE_Y = 64;
template <typename real>
struct TileBox {
  /* The TMA unit wants the byte offset of the box origin along the contiguous dimension to be a
   * multiple of 16 (measured on B200: any other inner coordinate raises "illegal instruction",
   * tools/tma_probe.cu), so the y halo is HY = 16 / sizeof(real) nodes wide instead of one. */
  static constexpr int HY = 16 / (int)sizeof(real);
  static constexpr int BY = TILE_Y + 2 * HY;
  static constexpr int BX = TILE_X + 2;
  static constexpr size_t bytes = (size_t)BY * BX * lbm::NQ * sizeof(real);
};

/* momentum-exchange accumulators are 64-bit fixed point: integer adds commute, so the sum
 * does not depend on the order in which tiles / GPUs contribute */
constexpr double FORCE_FIX = 4503599627370496.0;   /* 2^52 : fhf1, fhf2 (|sum| < 2^11) */
constexpr double TORQUE_FIX = 281474976710656.0;   /* 2^48 : fhf3       (|sum| < 2^15) */

template <typename real>
struct StepArgs {
  lbm::Lattice<real> L;
  real *f_new;                  /* [q][x-x0][y] */
  long long *facc;              /* [3][ngrains] fixed-point accumulators, or nullptr */
  int xlo, xhi;                 /* owned global rows [xlo, xhi) */
};

template <typename real>
struct GrainArrays {
  real *x1, *x2, *x3, *v1, *v2, *v3, *a1, *a2, *a3, *r, *m, *It, *rLB;
  real *fhf1, *fhf2, *fhf3;
};

/* ---- K1 (two builds of the same source: contraction on = fast, off = strict) ---- */
#define LBMDEM_DECLARE_K1(NS)                                                                              \
  namespace NS {                                                                                           \
  template <typename real>                                                                                 \
  cudaError_t launch_lbm_tiled(const CUtensorMap &tmap, const StepArgs<real> &a, cudaStream_t s);          \
  template <typename real>                                                                                 \
  cudaError_t launch_lbm_generic(const StepArgs<real> &a, cudaStream_t s);                                 \
  }
LBMDEM_DECLARE_K1(k1_fast)
LBMDEM_DECLARE_K1(k1_strict)

/* ---- everything below lives in the contraction-free translation unit (aux_kernels.cu) ---- */
template <typename real>
cudaError_t launch_raster(const lbm::RasterParams<real> &P, int ngrains, const GrainArrays<real> &g,
                          lbm::GrainRec<real> *rec, real *R2, lbm::GrainBox *boxes, int *cell, int x0, int nxl, int pitch,
                          cudaStream_t s);
cudaError_t launch_cell_frame(int *cell, int lx, int ly, int x0, int nxl, int pitch, int ring_value, cudaStream_t s);
/* act[x][y] as the reference would hold it (tests / diagnostics); L.act_folded must be 0 */
template <typename real>
cudaError_t launch_act_map(const lbm::Lattice<real> &L, int xlo, int xhi, int *act_out, cudaStream_t s);
/* fixed-point accumulators -> fhf (scaled, src/main.c:1329-1331); zeroes the accumulators */
template <typename real>
cudaError_t launch_force_finish(long long *facc, int ngrains, double k12, double k3, real *fhf1, real *fhf2, real *fhf3,
                                cudaStream_t s);
/* forces_fluid in the reference's own summation order, one thread per grain (strict mode) */
template <typename real>
cudaError_t launch_force_serial(const lbm::Lattice<real> &L, const real *f_new, int xlo, int xhi,
                                double *partial /* [3][n] unscaled */, cudaStream_t s);
template <typename real>
cudaError_t launch_force_scale(const double *partial, int ngrains, double k12, double k3, real *fhf1, real *fhf2,
                               real *fhf3, cudaStream_t s);

struct VerletBuffers {
  int nbuckets;        /* power of two */
  int *bucket_count;   /* [nbuckets + 1] -> exclusive offsets after the scan */
  int *bucket_cursor;  /* [nbuckets] */
  int *sorted;         /* [n] grain ids grouped by bucket */
  int *gcx, *gcy;      /* [n] integer cell coordinates */
  int *nbr_count;      /* [n] */
  int *nbr;            /* [n][cap] ascending */
  int cap;
  int *wflags;         /* [n] */
  int *error;          /* device flag: 1 = neighbour capacity exceeded */
};
template <typename real>
cudaError_t launch_verlet(const dem::Params<real> &P, int n, const GrainArrays<real> &g, real cell_size,
                          const VerletBuffers &vb, cudaStream_t s);
template <typename real>
cudaError_t launch_dem_step(const dem::Params<real> &P, int n, bool film, const GrainArrays<real> &g,
                            const VerletBuffers &vb, cudaStream_t s);

template <typename real>
cudaError_t launch_density(const real *f, int ly, int x0, int xlo, int xhi, int pitch, size_t plane, double *partials,
                           int npartials, double *out, cudaStream_t s);
/* VTK point data of the owned rows, [y][x - xlo] order, float32 (src/main.c:284-323) */
template <typename real>
cudaError_t launch_fields(const real *f, const int *cell, const GrainArrays<real> &g, const real *gp, int ngrains,
                          int ly, int x0, int xlo, int xhi, int pitch, size_t plane, real rho_moy, float *grain_p,
                          float *grain_v, float *grain_a, float *fluid_p, float *fluid_v, cudaStream_t s);

/* layout conversion for get_f / set_f: reference [x][y][q] (double) <-> device [q][x][y] (real) */
template <typename real>
cudaError_t launch_f_to_host_layout(const real *f, int ly, int pitch, size_t plane, int row0, int nrows, double *out,
                                    cudaStream_t s);
template <typename real>
cudaError_t launch_f_from_host_layout(real *f, int ly, int pitch, size_t plane, int row0, int nrows, const double *in,
                                      cudaStream_t s);
/* init_density (src/main.c:716-724) */
template <typename real>
cudaError_t launch_fill_rest(real *f, size_t plane, const lbm::Lattice<real> &Lw, cudaStream_t s);

}  // namespace lbmdem
