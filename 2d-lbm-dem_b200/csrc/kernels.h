/*
 * kernels.h -- launch interface between the host-side simulation object (sim.cu) and the
 * kernel translation units.  Internal to liblbmdem_gpu.so; the public boundary is
 * include/lbmdem_gpu.h.
 *
 * Kernel legend:
 *   K1  lbm_rows        fused wall ring + grain bounce-back + pull stream of the stored step, then
 *                       re-init + MRT collide of this step  (src/main.c:966-986, :1071-1243)
 *   K1f force           momentum exchange per grain     (src/main.c:1285-1333)
 *   K2  raster          grain records + obstacle map + act bits + boundary-node list (src/main.c:991-1065)
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

/* Row pipeline of the fused LBM kernel.  A CTA owns TY consecutive y-columns and marches along
 * x; each TMA transaction brings ONE lattice row of the strip into a ring of NS shared-memory
 * slots: the nine population planes (TY nodes plus a halo of HY nodes per side), the matching row
 * of this step's obstacle map (halo HC) -- its class bytes only, lbm_node.cuh cell_class -- and of the
 * stored step's map (no halo): the owner in 16 bits (cell_own16; the re-initialised nodes need it, and
 * fetching it from global memory inside the loop, even a row ahead, was slower) or, for samples of 65 534
 * grains and more, the int32 map itself.
 * The TMA unit wants the byte offset of the box origin along the contiguous dimension to be a
 * multiple of 16 (measured on B200: any other inner coordinate raises "illegal instruction",
 * tools/tma_probe.cu), so the y halo is 16 / sizeof(real) nodes wide instead of one; a box is at
 * most 256 elements wide. */
template <typename real>
struct RowCfg {
  static constexpr int TY = 128;            /* nodes (= threads) per CTA row */
  static constexpr int HY = 16 / (int)sizeof(real);
#ifndef LBMDEM_K1_NS
#define LBMDEM_K1_NS 4   /* measured on B200: more resident CTAs beat a deeper ring (profiles/r01_k1_tuning.txt) */
#endif
  static constexpr int NS = LBMDEM_K1_NS;   /* ring slots */
  static constexpr int BY = TY + 2 * HY;
  static constexpr int HC = 16;             /* y halo of the class-byte row of this step's map (one byte per node) */
  static constexpr int BC = TY + 2 * HC;
  static constexpr int A_BYTES = lbm::NQ * BY * (int)sizeof(real);
  static constexpr int CN_BYTES = BC, CP_BYTES = TY * 4;
  static constexpr int A_PAD = (A_BYTES + 127) / 128 * 128;
  static constexpr int CN_PAD = (CN_BYTES + 127) / 128 * 128;
  static constexpr int CP_PAD = (CP_BYTES + 127) / 128 * 128;
  /* A CTA owns NB adjacent blocks of TY columns with TY consumer threads each; every block has its own boxes (halo
   * included: as if NB CTAs shared one ring, one producer and one set of barriers).  fp32 rows are half the bytes of fp64
   * rows, and what a row costs besides its bytes -- barrier round trips, TMA issue, one row of prefetch to hide a DRAM
   * latency behind -- is per row: two blocks per row bring fp32 to the bytes-per-row of fp64 (which runs at 0.95+ of the
   * HBM peak). */
#ifndef LBMDEM_K1_NB
#define LBMDEM_K1_NB 1 /* 2 (fp32): 0.2335 ms with 256 + 32 threads, 0.2558 ms with 128 + 32 looping over both blocks, against 0.2301 ms (r02l, r02m) */
#endif
  static constexpr int NB = LBMDEM_K1_NB;
  static constexpr int BLOCK = A_PAD + CN_PAD + CP_PAD;
  static constexpr int SLOT = NB * BLOCK;
  static constexpr int SMEM = NS * SLOT;
};

using lbm::FORCE_FIX;
using lbm::TORQUE_FIX;

/* One fused launch: sweep 5 of the stored step (streaming, a plain pull from the array the ring
 * and bounce-back sweeps left behind), then sweeps 1-2 of this step (re-init, collide). */
template <typename real>
struct FusedArgs {
  lbm::Lattice<real> L;
  const real *A;                           /* stored populations, sweeps 1-4 applied, [q][x-x0][y] */
  const int *cell_prev;                    /* obstacle map of the stored step (the row kernel reads owners from it) */
  const int *cell_new;                     /* obstacle map of this step (plain / ring kernel) */
  const lbm::GrainRec<real> *grains_new;   /* grain records of this step */
  real *out;                               /* [q][x-x0][y] */
  int xlo, xhi;                            /* owned global rows [xlo, xhi) */
  int stream_only;                         /* 1: sweep 5 alone (materialises the reference's f) */
  int prev16;                              /* 1: the row kernel's tensor map of the stored step's map is the 16-bit owner plane */
  const unsigned char *cls_prev;           /* class bytes of the stored step's map: which populations nobody pulls (LDGSTS rows) */
};

/* links of the bounce-back sweep that must not be written while others are evaluated */
template <typename real>
struct DeferList {
  int *count;          /* device counter */
  size_t *index;       /* flat element index into the population buffer */
  real *value;
  int capacity;
  int *overflow;       /* flag in mapped host memory */
  int *range_flag;     /* flag in mapped host memory: a link's momentum exchange does not fit the fixed-point sums */
  int *seen;           /* mapped host memory: how many links the last sweep deferred (the host sizes the next launch by it) */
};

template <typename real>
struct GrainArrays {
  real *x1, *x2, *x3, *v1, *v2, *v3, *a1, *a2, *a3, *r, *m, *It, *rLB;
  real *fhf1, *fhf2, *fhf3;
};

/* ---- K1 (two builds of the same source: contraction on = fast, off = strict) ---- */
#define LBMDEM_DECLARE_K1(NS)                                                                                          \
  namespace NS {                                                                                                       \
  /* interior nodes, TMA row pipeline */                                                                               \
  template <typename real>                                                                                             \
  cudaError_t launch_lbm_rows(const CUtensorMap &tmA, const CUtensorMap &tmCp, const CUtensorMap &tmCn,                \
                              const FusedArgs<real> &a, cudaStream_t s);                                               \
  /* same map from global memory, one thread per node: ring nodes only (ring_only = 1) or every owned node            \
   * (cross-check of the row kernel, params.kernel = 1) */                                                             \
  template <typename real>                                                                                             \
  cudaError_t launch_lbm_plain(const FusedArgs<real> &a, int ring_only, cudaStream_t s);                               \
  /* sweeps 1-2 alone, in place (first step, or after the populations were set from outside) */                        \
  template <typename real>                                                                                             \
  cudaError_t launch_lbm_h1(const lbm::Lattice<real> &L, real *f, const int *cell_prev, const int *cell_now,           \
                            const lbm::GrainRec<real> *grains_new, int xlo, int xhi, cudaStream_t s);                  \
  /* the populations the fused kernel leaves unwritten (lbm_node.cuh, node_is_dead), rows [xa, xb) of A */            \
  template <typename real>                                                                                             \
  cudaError_t launch_lbm_fill_dead(const lbm::Lattice<real> &L, real *A, const int *cell_prev, const int *cell_now,    \
                                   const lbm::GrainRec<real> *grains_new, int xa, int xb, cudaStream_t s);             \
  }
LBMDEM_DECLARE_K1(k1_fast)
LBMDEM_DECLARE_K1(k1_strict)

/* ---- everything below lives in the contraction-free translation unit (aux_kernels.cu) ---- */
/* The two sparse work lists of a step are kept PER LATTICE TILE (RTX rows x RTY columns): tile t owns the segment
 * entry[t * cap .. t * cap + tcount[t]).  The tile kernel rewrites a tile's segments only when the tile's part of the
 * obstacle map changed (see grain_bin_kernel); the sweep kernels run one CTA per tile.
 *
 * Boundary nodes: solid nodes with at least one neighbour that is not owned by the same grain and not fluid.
 * entry.x = local node index (row * pitch + y), entry.y = grain index | BL_ACT if act[x][y] == 1 | mask << 24 where
 * bit q-1 of mask marks link q as leaving the grain. */
struct BoundaryList {
  uint2 *entry;
  int *tcount;         /* [ntx * nty] entries per tile */
  int cap;             /* entries per tile segment */
  int ntx, nty;
  int *overflow;       /* flag in mapped host memory */
};
constexpr unsigned BL_ACT = 1u << 23;
constexpr unsigned BL_GRAIN = BL_ACT - 1;
/* Links of the bounce-back sweep that the fused kernel cannot do on its own: entry.x = local node
 * index of the active solid node, entry.y = grain index | q << 24 | LL_W if the link just takes
 * the rest value (a w-link next to the wall ring, lbm_node.cuh w_links_with_collide) | LL_CLEAR if
 * the rasteriser saw that the node two steps along the link is fluid or wall ring (it lies inside the
 * painted tile + halo for 95 % of the links): such a link cannot be one across a one-node gap, and
 * the sweep skips the map read that would tell. */
struct LinkList {
  uint2 *entry;
  int *tcount;
  int cap;
  int ntx, nty;
  int *overflow;       /* flag in mapped host memory */
};
constexpr unsigned LL_W = 1u << 28;
constexpr unsigned LL_CLEAR = 1u << 29; /* the node two steps along the link is known to be fluid or wall ring: no gap */

/* Grains binned by the lattice tiles (RTX rows x RTY columns, plus one halo node all round) that their bounding
 * box touches: filled by the first kernel of the rasteriser, consumed and emptied by the tile kernel. */
constexpr int RTX = 32, RTY = 64;
/* what the tile kernel needs to paint a grain: index, clamped bounding box (src/main.c:1016-1023), disc */
template <typename real>
struct alignas(16) TileEntry {
  int id, xi, xf, yi, yf;
  real xc, yc, r2, RR;   /* centre in lattice units, rLB^2, (r/dx)^2 */
};
struct TileBins {
  int *count;          /* [ntx * nty] */
  void *list;          /* [ntx * nty][cap] TileEntry<real>, any order */
  int cap;
  int ntx, nty;        /* tiles along x (local rows) and y */
  int *stamp;          /* [ntx * nty] rasteriser step at which some covered node of the tile (halo included) last changed */
  int *dirty;          /* [2][ntx * nty] tiles stamped at even / odd rasteriser steps, in the order they were stamped */
  int *ndirty;         /* [2] their numbers */
  int *ticket;         /* the tile kernel's CTAs count themselves out: the last one empties the bins */
  int resident_ctas;   /* CTAs of the tile kernel one wave holds on this device */
  int *overflow;       /* flag in mapped host memory */
};

/* K2: the obstacle map of the step (obst_construction, src/main.c:991-1065) with act / rim bits, and the two sparse
 * lists.  grain_bin_kernel: records, tile bins, and which tiles hold a node whose owner may have changed since the
 * previous step (rec_old ...: the previous step's records).  raster_tile_kernel, one CTA per tile, for the tiles
 * stamped at this step or the previous one (all of them when force_full): paints the reduced discs of the tile's
 * grains into shared memory (owner = highest covering index, lowest covering index kept beside it), derives act /
 * rim bits from the shared-memory neighbourhood, writes the tile to `cell` in full rows and rewrites the tile's
 * segments of the two lists.  Also empties the deferred list and zeroes the force sums. */
template <typename real>
cudaError_t launch_raster_tiles(const lbm::RasterParams<real> &P, int ngrains, const GrainArrays<real> &g,
                                lbm::GrainRec<real> *rec, real *R2, lbm::GrainBox *boxes,
                                const lbm::GrainRec<real> *rec_old, const real *R2_old, const lbm::GrainBox *boxes_old,
                                int *cell, const int *cell_other /* the previous step's map */,
                                unsigned char *cls, const unsigned char *cls_other /* their class bytes */,
                                unsigned short *own16, const unsigned short *own16_other /* their 16-bit owners */, int x0,
                                int nxl, int pitch,
                                const TileBins &T, const BoundaryList &B, const LinkList &K,
                                int *defer_count /* emptied as well */,
                                long long *facc /* nullptr, or [3][n] force sums to be zeroed */, int step,
                                int first_run /* no previous records */, int force_full /* rebuild every tile */,
                                cudaStream_t s);
cudaError_t launch_cell_frame(int *cell, unsigned char *cls, unsigned short *own16, int lx, int ly, int x0, int nxl, int pitch,
                              int ring_value, cudaStream_t s);
/* class bytes of a map that came from outside (lbmdem_set_obst): fluid / solid / ring only */
cudaError_t launch_cls_from_cell(const int *cell, unsigned char *cls, unsigned short *own16, int nxl, int pitch, int ngrains,
                                 cudaStream_t s);
/* act[x][y] as the reference would hold it (tests / diagnostics) */
template <typename real>
cudaError_t launch_act_map(const lbm::Lattice<real> &L, const lbm::Stored<real> &S, int xlo, int xhi, int *act_out,
                           cudaStream_t s);
/* sweep 3 in place: wall-ring copies (src/main.c:1123-1145), ring nodes of the rows [xa, xb) */
template <typename real>
cudaError_t launch_ring_sweep(const lbm::Lattice<real> &L, const lbm::Stored<real> &S, real *A, int xa, int xb,
                              cudaStream_t s);
/* sweep 4 in place: interpolated bounce-back on the active solid nodes (src/main.c:1154-1222), one thread per
 * listed link; links facing another grain across a one-node gap go through the deferred list (lbm_node.cuh,
 * sweep_link).  Two stages (the deferred list and facc[3][n] were emptied by launch_raster):
 * passes over disjoint row ranges [xa, xb) (each also adds the momentum exchange of its links into fluid
 * neighbours whose solid node lies in [xlo, xhi)), end (applies the deferred links; after every pass). */
template <typename real>
cudaError_t launch_bounce_pass(const lbm::Lattice<real> &L, const lbm::Stored<real> &S, real *A, int xa, int xb, int xlo,
                               int xhi, const LinkList &K, const DeferList<real> &D, long long *facc, cudaStream_t s);
template <typename real>
cudaError_t launch_bounce_end(real *A, const DeferList<real> &D, cudaStream_t s);
/* the rest of forces_fluid (src/main.c:1295-1325): links into non-fluid foreign neighbours (other
 * grains, the wall ring), one thread per (listed node, link); ADDS to facc[3][n] */
template <typename real>
cudaError_t launch_force_links(const lbm::Lattice<real> &L, const lbm::Stored<real> &S, int xlo, int xhi,
                               const BoundaryList &B, long long *facc,
                               real *A /* nullptr, or the populations: the deferred links D are applied first */,
                               const DeferList<real> &D, cudaStream_t s);
/* sums -> fhf, scaled as src/main.c:1329-1331: from the fixed-point sums (default build) or the fp64 sums the strict
 * build adds in the reference's order.  Passed to the first DEM launch after an LBM step (it does the conversion for
 * its own grain), or to launch_force_finish when something else wants fhf first. */
struct ForceFinish {
  const void *sums;    /* nullptr: fhf is up to date */
  int fixed_point;     /* 1: long long [3][n] scaled by FORCE_FIX / TORQUE_FIX; 0: double [3][n] */
  double k12, k3;
  int *range_flag;     /* mapped host memory: raised when a fixed-point sum is about to leave its range */
};
template <typename real>
cudaError_t launch_force_finish(const ForceFinish &fin, int ngrains, real *fhf1, real *fhf2, real *fhf3, cudaStream_t s);
/* forces_fluid in the reference's own summation order, one thread per grain (strict mode) */
template <typename real>
cudaError_t launch_force_serial(const lbm::Lattice<real> &L, const lbm::Stored<real> &S, int xlo, int xhi,
                                double *partial /* [3][n] unscaled */, cudaStream_t s);
/* single-GPU form of the bounce-back sweep + the rest of forces_fluid in ONE launch (aux_kernels.cu, rim_kernel) */
template <typename real>
cudaError_t launch_rim(const lbm::Lattice<real> &L, const lbm::Stored<real> &S, real *A, int xa, int xb, int xlo, int xhi,
                       const LinkList &K, const BoundaryList &B, const DeferList<real> &D, long long *facc, int *ticket,
                       int apply_here /* 0: a defer_apply launch follows (many deferred links) */, cudaStream_t s);

/* partial force sums of the ranks of an in-process strip group (device pointers, peer-accessible) */
constexpr int MAX_LOCAL_RANKS = 16;
struct PeerPtrs {
  const void *p[MAX_LOCAL_RANKS];
  int count;
};
template <typename T>
cudaError_t launch_peer_sum(const PeerPtrs &pp, int len, T *out, cudaStream_t s);

/* One process per GPU (the NCCL transport): the fixed-point force sums of the ranks added by a kernel that reads the
 * peers' partial sums through CUDA IPC mappings over NVLink instead of ncclAllReduce (which costs ~30 us on 8 GPUs
 * whatever the size, profiles/r02_k1_probes.txt).  facc[k]: rank k's partial sums of this step (3 n int64, rank k's
 * memory); flags[k]: rank k's arrival words (flags[k][j] = the last step whose sums rank j has completed, written by
 * rank j into rank k's memory). */
struct IpcPeers {
  const long long *facc[MAX_LOCAL_RANKS];
  unsigned *flags[MAX_LOCAL_RANKS];
  int nranks, rank, lx;
};
cudaError_t launch_ipc_publish(const IpcPeers &pp, unsigned step, cudaStream_t s);
cudaError_t launch_ipc_sum(const IpcPeers &pp, unsigned step, int n, const lbm::GrainBox *boxes, long long *out, int *timeout_flag,
                           cudaStream_t s);

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
                            const VerletBuffers &vb, real *mid /* nullptr, or [6][n]: x1 x2 x3 v1 v2 v3 after the
                            kick-drift, i.e. as acceleration_grains() sees them */,
                            bool drift_done /* the previous sub-step's last launch did this one's kick-drift */,
                            bool drift_next /* close with kick + the NEXT sub-step's kick-drift in one launch */,
                            cudaStream_t s);

/* nsub consecutive DEM sub-steps (fixed lists and fhf) in ONE launch, a thread per grain: one thread-block cluster of
 * 8 CTAs up to DEM_CLUSTER_MAX grains, a cooperative grid for anything that fits one co-resident wave
 * (dem_coop_capacity); the launch that follows an LBM step also turns the force sums into fhf (fin) */
constexpr int DEM_COOP_THREADS = 128;
constexpr int DEM_CLUSTER_MAX = 1024;
template <typename real>
cudaError_t launch_dem_coop(const dem::Params<real> &P, int n, int nsub, bool film_first, const GrainArrays<real> &g,
                            const VerletBuffers &vb, const ForceFinish &fin, cudaStream_t s);
template <typename real>
cudaError_t dem_coop_capacity(int *max_grains);

template <typename real>
cudaError_t launch_density(const real *f, int ly, int x0, int xlo, int xhi, int pitch, size_t plane, double *partials,
                           int npartials, double *out, cudaStream_t s);
/* exact order-free fingerprint of f[x][y][q] and obst[x][y] over the owned rows (aux_kernels.cu, checksum_kernel) */
template <typename real>
cudaError_t launch_checksum(const real *f, const int *cell, int ly, int x0, int xlo, int xhi, int pitch, size_t plane,
                            unsigned long long *out /* [2], device */, int blocks, cudaStream_t s);
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
/* grain rows [n][ncols] of doubles or floats (the C ABI's layouts) <-> the device's arrays of `real` [ncols][n] */
template <typename real>
cudaError_t launch_grain_unpack(const void *rows, bool rows_f32, int n, int ncols, real *cols, cudaStream_t s);
template <typename real>
cudaError_t launch_grain_pack(const real *cols, int n, int ncols, void *rows, bool rows_f32, cudaStream_t s);
/* two column groups in one launch: rows = [n][na] followed by [n][nb] */
template <typename real>
cudaError_t launch_grain_pack2(const real *cols_a, int na, const real *cols_b, int nb, int n, void *rows, bool rows_f32,
                               cudaStream_t s);
/* init_density (src/main.c:716-724) */
template <typename real>
cudaError_t launch_fill_rest(real *f, size_t plane, const lbm::Lattice<real> &Lw, cudaStream_t s);

}  // namespace lbmdem
