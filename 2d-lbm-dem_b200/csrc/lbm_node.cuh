/*
 * lbm_node.cuh -- node-level arithmetic of the coupled D2Q9 MRT step, written once for host
 * and device.
 *
 * The reference (cb-geo/2d-lbm-dem, src/main.c) does one LBM step as five in-place sweeps over
 * f[x][y][q]: (1) re-initialise solid nodes (:966-986), (2) MRT-collide fluid nodes
 * (:1077-1119), (3) wall-ring copies (:1123-1145), (4) interpolated bounce-back on active solid
 * nodes (:1154-1222) and (5) two swap passes that stream (:1224-1242).
 *
 * Sweeps 1-2 are node-local (reinit_collide below).  Sweeps 3-4 rewrite a sparse set of
 * populations -- ring nodes, the rim of the grains -- and run as small in-place kernels
 * (ring_value, sweep_link below).  After them the array "A" holds exactly what the reference
 * holds before its swap passes, and sweep 5 is a plain pull:
 *
 *      f_next[q](p) = A[q](p - e_q)        if p - e_q lies inside the array,
 *                   = A[opp q](p)          otherwise                      (swap passes, :1224-1242)
 *
 * The device therefore stores A between steps and one fused kernel does "sweep 5 of step n-1,
 * then sweeps 1-2 of step n" per node (lbm_kernels.cu); the hydrodynamic force of step n
 * (forces_fluid, :1285-1333) reads A as well (force_link).  Operand order and
 * int/float/double promotions follow the reference source expression by expression (they
 * decide the last bit, and in the -DSINGLE_PRECISION build the `1.`/`4.5`/`fabs`/`sqrt`
 * promotions to double are part of the result).
 *
 * Two users:
 *   - lbm_kernels.cu / aux_kernels.cu: the TMA row kernel uses reinit_collide on registers and
 *     pulls from shared-memory rows; the ring, sweep and force kernels call ring_value,
 *     sweep_link and force_link on global memory.
 *   - tests/hostcheck: the same header compiled by g++ to pin the formulation against the
 *     oracle on the CPU, bit for bit (test infrastructure only; not a product path).
 */
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define LBM_HD __host__ __device__ __forceinline__
#define LBM_HD_SLOW __host__ __device__ __noinline__ /* sparse paths: one copy, called */
#else
#define LBM_HD inline
#define LBM_HD_SLOW inline
#endif

namespace lbm {

constexpr int NQ = 9;
constexpr int CELL_FLUID = -1;          /* obst == -1 (src/main.c:999) */
constexpr int CELL_ACT = 1 << 30;       /* act[x][y] == 1 of a solid node, folded into the map */
constexpr int CELL_RIM = 1 << 29;       /* a solid node with a NON-fluid neighbour it does not share its owner with
                                           (another grain, the wall ring): forces_fluid reads its populations */
constexpr int CELL_IDX = CELL_RIM - 1;

/* src/main.c:70-71 */
LBM_HD int ex_of(int q) { return (q >= 1 && q <= 3) ? -1 : ((q >= 5 && q <= 7) ? 1 : 0); }
LBM_HD int ey_of(int q) { return (q == 1 || q == 7 || q == 8) ? 1 : ((q >= 3 && q <= 5) ? -1 : 0); }
LBM_HD int opp_of(int q) { return q == 0 ? 0 : (q <= 4 ? q + 4 : q - 4); }

LBM_HD bool cell_is_fluid(int c) { return c < 0; }
LBM_HD int cell_obst(int c) { return c < 0 ? -1 : (c & CELL_IDX); }
LBM_HD bool cell_is_act(int c) { return c >= 0 && (c & CELL_ACT) != 0; }

/* One byte per node beside the map: all the fused LBM kernel needs of a node's map entry -- except the owner index,
 * which it fetches from the int map for the few nodes that are re-initialised (solid under the stored step's map and
 * not dead).  0 = fluid. */
constexpr unsigned char CLS_SOLID = 1;  /* not fluid: a grain node or the wall ring */
constexpr unsigned char CLS_ACT = 2;    /* CELL_ACT */
constexpr unsigned char CLS_RIM = 4;    /* CELL_RIM */
constexpr unsigned char CLS_RING = 8;   /* the wall ring (obst == nbgrains): no grain record behind it */
LBM_HD unsigned char cell_class(int c, int ngrains) {
  if (c < 0) return 0;
  return (unsigned char)(CLS_SOLID | ((c & CELL_ACT) ? CLS_ACT : 0) | ((c & CELL_RIM) ? CLS_RIM : 0) |
                         ((c & CELL_IDX) >= ngrains ? CLS_RING : 0));
}

/* ... and the owner in 16 bits, for samples of fewer than 65 534 grains: what the fused kernel streams of the STORED
 * step's map (it needs nothing else of it: fluid or not, and whose equilibrium a re-initialised node takes) */
constexpr unsigned short OWN16_FLUID = 0xFFFF, OWN16_NONE = 0xFFFE; /* NONE: wall ring / index out of range */
LBM_HD unsigned short cell_own16(int c) {
  if (c < 0) return OWN16_FLUID;
  const int i = c & CELL_IDX;
  return i < (int)OWN16_NONE ? (unsigned short)i : OWN16_NONE;
}

/* what the LBM kernels need to know about one grain (filled by the rasteriser, K2) */
template <typename real>
struct GrainRec {
  real xc, yc, r2;        /* (x1-Mgx)/dx, (x2-Mby)/dx, rLB^2   src/main.c:1009-1011 */
  real x1, x2, v1, v2, v3;
};

/* clamped bounding box of a grain, src/main.c:1016-1023 (empty when xi > xf or yi > yf) */
struct GrainBox {
  int xi, xf, yi, yf;
};
LBM_HD bool box_has(const GrainBox &b, int x, int y) { return x >= b.xi && x <= b.xf && y >= b.yi && y <= b.yf; }

/* src/main.c:1026-1029 */
template <typename real>
LBM_HD bool disc_covers(real xc, real yc, real r2, real R2, int x, int y) {
  const real dist2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
  return dist2 <= R2 && dist2 <= r2;
}

/* Was node n "fluid" when the reference's grain loop reached grain i (:1047)?  At that moment
 * the map holds grains 0..i only.  n is fluid then iff no grain j <= i covers it: final map
 * -1, or final owner k > i while grain i itself does not cover n.  This form needs nothing but
 * the map and is used for maps that come from outside (lbmdem_set_obst); a node covered by k > i
 * AND by some j < i but not by i -- three mutually overlapping reduced discs -- is misjudged by
 * it (reduced discs are 0.85 r: even a pair only overlaps at > 15 % interpenetration).  The
 * device's own rasteriser uses fluid_when_grain_ran_exact below.
 */
template <typename real>
LBM_HD bool fluid_when_grain_ran(int cell_n, int i, int ngrains, real xc, real yc, real r2, real R2,
                                 const GrainBox &b, int nx, int ny) {
  if (cell_is_fluid(cell_n)) return true;
  const int k = cell_obst(cell_n);
  if (k >= ngrains || k <= i) return false;
  return !(box_has(b, nx, ny) && disc_covers(xc, yc, r2, R2, nx, ny));
}

/* geometry and constants of the lattice (strip-local storage: rows x0 .. x0+nxl-1) */
/* The exact form of the same question, for maps the rasteriser built itself: it also records,
 * for every node that more than one reduced disc covers, the LOWEST covering index (min_owner;
 * -1 = the node is covered by its owner alone).  n was fluid when grain i ran iff no grain
 * j <= i covers it, i.e. iff the lowest covering index is greater than i -- whatever the number
 * of mutually overlapping discs. */
LBM_HD bool fluid_when_grain_ran_exact(int cell_n, int i, int ngrains, int min_owner) {
  if (cell_is_fluid(cell_n)) return true;
  const int k = cell_obst(cell_n);
  if (k >= ngrains || k <= i) return false;
  return (min_owner >= 0 ? min_owner : k) > i;
}
/* min-owner map entries: (generation key << 23) | grain index, written with atomicMin; the key
 * DEcreases from step to step, so that entries of older steps lose against this step's and the
 * map needs clearing only when the key wraps (csrc/sim.cu) */
constexpr int MINOWNER_SHIFT = 23;
constexpr int MINOWNER_GRAIN = (1 << MINOWNER_SHIFT) - 1;
constexpr int MINOWNER_EMPTY = 0x7f7f7f7f;
LBM_HD int min_owner_decode(int entry, int genkey) {
  return (entry >> MINOWNER_SHIFT) == genkey ? (entry & MINOWNER_GRAIN) : -1;
}

template <typename real>
struct Lattice {
  int lx, ly;             /* global lattice size */
  int x0;                 /* global x of local row 0 (strip decomposition; 0 on one GPU) */
  int nxl;                /* rows held locally, ghost rows included */
  int pitch;              /* elements per row (>= ly) */
  size_t plane;           /* elements per population plane = nxl * pitch */
  int ngrains;
  real dx, c, Mgx, Mby, lid6;
  real s2, s3, s5, s7, s8, s9;
  real w[NQ];
};

/* The lattice state the device keeps between LBM steps: the populations of step n AFTER the
 * re-init and collide sweeps ("A"), together with the obstacle map and the grain records of
 * that same step.  Everything the remaining sweeps of step n produce (wall ring, grain
 * bounce-back, streaming, hydrodynamic forces) is a pure function of this. */
template <typename real>
struct Stored {
  const real *A;          /* [q][x-x0][y] */
  const int *cell;        /* obstacle map of the step (act bit folded in when act_folded) */
  const GrainRec<real> *grains;
  const GrainBox *boxes;  /* per grain, with R2 = (r/dx)^2: only the act rule needs them */
  const real *R2;
  int act_folded;         /* 1: cell carries CELL_ACT; 0: derive act on demand */
};

template <typename real>
LBM_HD size_t node_index(const Lattice<real> &L, int x, int y) {
  return (size_t)(x - L.x0) * L.pitch + y;
}
template <typename real>
LBM_HD bool in_array(const Lattice<real> &L, int x, int y) {
  return x >= 0 && y >= 0 && x < L.lx && y < L.ly;
}
template <typename real>
LBM_HD bool is_ring(const Lattice<real> &L, int x, int y) {
  return x == 0 || y == 0 || x == L.lx - 1 || y == L.ly - 1;
}

/* act[x][y] of an interior solid node (src/main.c:1038-1052): some neighbour was fluid when the
 * owner grain was rasterised.  Either read from the folded bit or derived from the map. */
template <typename real>
LBM_HD_SLOW bool node_act(const Lattice<real> &L, const Stored<real> &S, int x, int y, int c) {
  if (c < 0) return false;
  if (c & CELL_ACT) return true;
  if (S.act_folded) return false;
  const int i = cell_obst(c);
  if (i >= L.ngrains) return false;
  const GrainRec<real> &g = S.grains[i];
  for (int q = 1; q < NQ; ++q) {
    const int nx = x + ex_of(q), ny = y + ey_of(q);
    if (fluid_when_grain_ran(S.cell[node_index(L, nx, ny)], i, L.ngrains, g.xc, g.yc, g.r2, S.R2[i], S.boxes[i], nx, ny))
      return true;
  }
  return false;
}

/* rigid-body velocity of a grain at lattice node (x,y): the sub-expressions of :974-980 */
template <typename real>
LBM_HD real wall_ux(const Lattice<real> &L, const GrainRec<real> &g, int y) {
  return g.v1 - (y * L.dx + L.Mby - g.x2) * g.v3;
}
template <typename real>
LBM_HD real wall_uy(const Lattice<real> &L, const GrainRec<real> &g, int x) {
  return g.v2 + (x * L.dx + L.Mgx - g.x1) * g.v3;
}

/* src/main.c:974-981 */
template <typename real>
LBM_HD void equilibrium(const Lattice<real> &L, const GrainRec<real> &g, int x, int y, real *out) {
  const real ux = wall_ux(L, g, y), uy = wall_uy(L, g, x);
#if defined(LBM_RELAXED)
  /* default device build: one reciprocal instead of ten divisions, polynomial in `real`, and the two populations
   * of a direction pair (e, -e) share 1 + 4.5 (e.u)^2 - 1.5 u^2 */
  const real ic = 1 / L.c;
  const real X = ux * ic, Y = uy * ic;
  const real base = (real)1 - (real)1.5 * (X * X + Y * Y);
  const real P = X + Y, M = X - Y;
  const real tx = base + (real)4.5 * X * X, ty = base + (real)4.5 * Y * Y;
  const real tp = base + (real)4.5 * P * P, tm = base + (real)4.5 * M * M;
  out[0] = L.w[0] * base;
  out[6] = L.w[6] * (tx + 3 * X);   /* e = (+1, 0) */
  out[2] = L.w[2] * (tx - 3 * X);   /* e = (-1, 0) */
  out[8] = L.w[8] * (ty + 3 * Y);   /* e = (0, +1) */
  out[4] = L.w[4] * (ty - 3 * Y);   /* e = (0, -1) */
  out[7] = L.w[7] * (tp + 3 * P);   /* e = (+1, +1) */
  out[3] = L.w[3] * (tp - 3 * P);   /* e = (-1, -1) */
  out[5] = L.w[5] * (tm + 3 * M);   /* e = (+1, -1) */
  out[1] = L.w[1] * (tm - 3 * M);   /* e = (-1, +1) */
#else
  const real u_squ = (ux * ux + uy * uy) / (L.c * L.c);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const real eu = (ex_of(q) * ux + ey_of(q) * uy) / L.c;
    out[q] = L.w[q] * (1. + 3 * eu + 4.5 * eu * eu - 1.5 * u_squ);
  }
#endif
}

/* src/main.c:1082-1116.  The reference's expressions verbatim -- except in the default device build
 * (LBM_RELAXED, below), which evaluates the same nine-moment map with shared sub-sums. */
template <typename real>
LBM_HD void mrt_collide(const Lattice<real> &L, real *p) {
#if defined(LBM_RELAXED)
  /* Same linear maps, regrouped (differences at rounding level, well inside the parity tolerance):
   * the moments from the sums over the diagonal and over the axis populations, one reciprocal of rho
   * for the four equilibria, the back-transform from two shared combinations per population class. */
  const real sd = (p[1] + p[3]) + (p[5] + p[7]), sa = (p[2] + p[4]) + (p[6] + p[8]);
  const real rho = p[0] + sd + sa;
  const real e = -4 * p[0] + 2 * sd - sa;
  const real eps = 4 * p[0] + sd - 2 * sa;
  const real dxp = p[5] + p[7], dxm = p[1] + p[3], dyp = p[1] + p[7], dym = p[3] + p[5];
  const real j_x = (dxp - dxm) + (p[6] - p[2]);
  const real q_x = (dxp - dxm) - 2 * (p[6] - p[2]);
  const real j_y = (dyp - dym) + (p[8] - p[4]);
  const real q_y = (dyp - dym) - 2 * (p[8] - p[4]);
  const real p_xx = (p[2] + p[6]) - (p[4] + p[8]);
  const real p_xy = (p[3] + p[7]) - (p[1] + p[5]);
  const real ir = 1 / rho;
  const real j_x2 = j_x * j_x, j_y2 = j_y * j_y;
  const real jj = 3 * (j_x2 + j_y2) * ir;
  const real eO = e - L.s2 * (e + 2 * rho - jj);
  const real epsO = eps - L.s3 * (eps - rho + jj);
  const real q_xO = q_x - L.s5 * (q_x + j_x);
  const real q_yO = q_y - L.s7 * (q_y + j_y);
  const real p_xxO = p_xx - L.s8 * (p_xx - (j_x2 - j_y2) * ir);
  const real p_xyO = p_xy - L.s9 * (p_xy - j_x * j_y * ir);
  const real a = (real)(1. / 36);
  const real r4 = 4 * rho;
  const real A = a * (r4 - eO - 2 * epsO), B = a * (r4 + 2 * eO + epsO);
  const real cx = (6 * a) * (j_x - q_xO), cy = (6 * a) * (j_y - q_yO);
  const real gx = (3 * a) * (2 * j_x + q_xO), gy = (3 * a) * (2 * j_y + q_yO);
  const real pxx9 = (9 * a) * p_xxO, pxy9 = (9 * a) * p_xyO;
  p[0] = a * (r4 - 4 * eO + 4 * epsO);
  p[2] = (A + pxx9) - cx;
  p[6] = (A + pxx9) + cx;
  p[4] = (A - pxx9) - cy;
  p[8] = (A - pxx9) + cy;
  p[1] = (B - pxy9) - (gx - gy);
  p[5] = (B - pxy9) + (gx - gy);
  p[3] = (B + pxy9) - (gx + gy);
  p[7] = (B + pxy9) + (gx + gy);
#else
  const real a = 1. / 36;
  real rho = p[0] + p[1] + p[2] + p[3] + p[4] + p[5] + p[6] + p[7] + p[8];
  real e = -4 * p[0] + 2 * p[1] - p[2] + 2 * p[3] - p[4] + 2 * p[5] - p[6] + 2 * p[7] - p[8];
  real eps = 4 * p[0] + p[1] - 2 * p[2] + p[3] - 2 * p[4] + p[5] - 2 * p[6] + p[7] - 2 * p[8];
  real j_x = p[5] + p[6] + p[7] - p[1] - p[2] - p[3];
  real q_x = -p[1] + 2 * p[2] - p[3] + p[5] - 2 * p[6] + p[7];
  real j_y = p[1] + p[8] + p[7] - p[3] - p[4] - p[5];
  real q_y = p[1] - p[3] + 2 * p[4] - p[5] + p[7] - 2 * p[8];
  real p_xx = p[2] - p[4] + p[6] - p[8];
  real p_xy = -p[1] + p[3] - p[5] + p[7];
  real j_x2 = j_x * j_x, j_y2 = j_y * j_y;
  real eO = e - L.s2 * (e + 2 * rho - 3 * (j_x2 + j_y2) / rho);
  real epsO = eps - L.s3 * (eps - rho + 3 * (j_x2 + j_y2) / rho);
  real p_xxO = p_xx - L.s8 * (p_xx - (j_x2 - j_y2) / rho);
  real p_xyO = p_xy - L.s9 * (p_xy - j_x * j_y / rho);
  real q_xO = q_x - L.s5 * (q_x + j_x);
  real q_yO = q_y - L.s7 * (q_y + j_y);
  p[0] = a * (4 * rho - 4 * eO + 4 * epsO);
  p[2] = a * (4 * rho - eO - 2 * epsO - 6 * j_x + 6 * q_xO + 9 * p_xxO);
  p[4] = a * (4 * rho - eO - 2 * epsO - 6 * j_y + 6 * q_yO - 9 * p_xxO);
  p[6] = a * (4 * rho - eO - 2 * epsO + 6 * j_x - 6 * q_xO + 9 * p_xxO);
  p[8] = a * (4 * rho - eO - 2 * epsO + 6 * j_y - 6 * q_yO - 9 * p_xxO);
  p[1] = a * (4 * rho + 2 * eO + epsO - 6 * j_x - 3 * q_xO + 6 * j_y + 3 * q_yO - 9 * p_xyO);
  p[3] = a * (4 * rho + 2 * eO + epsO - 6 * j_x - 3 * q_xO - 6 * j_y - 3 * q_yO + 9 * p_xyO);
  p[5] = a * (4 * rho + 2 * eO + epsO + 6 * j_x + 3 * q_xO - 6 * j_y - 3 * q_yO - 9 * p_xyO);
  p[7] = a * (4 * rho + 2 * eO + epsO + 6 * j_x + 3 * q_xO + 6 * j_y + 3 * q_yO + 9 * p_xyO);
#endif
}

/* ---- arithmetic of the grain bounce-back links, immune to multiply-add contraction ----
 * Every product and sum of the link arithmetic is an explicitly rounded operation on the device, so that it gives the
 * same bits whichever translation unit it is compiled in (the sparse sweep kernels are built with -fmad=false, the
 * fused LBM kernel with contraction in the default build); the host build of tests/hostcheck uses -ffp-contract=off. */
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float x_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float x_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float x_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float x_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double x_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double x_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double x_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double x_div(double a, double b) { return __ddiv_rn(a, b); }
#else
inline float x_mul(float a, float b) { return a * b; }
inline float x_add(float a, float b) { return a + b; }
inline float x_sub(float a, float b) { return a - b; }
inline float x_div(float a, float b) { return a / b; }
inline double x_mul(double a, double b) { return a * b; }
inline double x_add(double a, double b) { return a + b; }
inline double x_sub(double a, double b) { return a - b; }
inline double x_div(double a, double b) { return a / b; }
#endif

/* e_q . u_wall at solid node (x,y) of grain g: `ex * wall_ux + ey * wall_uy` with the sub-expressions of :974-980 */
template <typename real>
LBM_HD real link_eu(const Lattice<real> &L, const GrainRec<real> &g, int x, int y, int q) {
  const real ux = x_sub(g.v1, x_mul(x_sub(x_add(x_mul((real)y, L.dx), L.Mby), g.x2), g.v3));
  const real uy = x_add(g.v2, x_mul(x_sub(x_add(x_mul((real)x, L.dx), L.Mgx), g.x1), g.v3));
  return x_add(x_mul((real)ex_of(q), ux), x_mul((real)ey_of(q), uy));
}

/* src/main.c:1053-1058: link fraction for the link from solid node (x,y) along q to its fluid
 * neighbour, measured from the fluid node.  C semantics: fabs/sqrt are the double functions. */
template <typename real>
LBM_HD real link_delta(const GrainRec<real> &g, int x, int y, int q) {
  const int ex = ex_of(q), ey = ey_of(q);
  const real aa = fabs((double)ex) + fabs((double)ey);
  const real dxn = x_sub((real)(x + ex), g.xc), dyn = x_sub((real)(y + ey), g.yc);
  const real bb = x_add(x_mul(dxn, (real)ex), x_mul(dyn, (real)ey));
  const real cc = x_sub(x_add(x_mul(dxn, dxn), x_mul(dyn, dyn)), g.r2);
  const real disc = x_sub(x_mul(bb, bb), x_mul(aa, cc));
  return (real)x_div(x_sub((double)bb, sqrt(fabs((double)disc))), (double)aa);
}

/* src/main.c:1166-1185 (and :1198-1217): the two interpolated bounce-back formulas.
 * Fn_oq = F[n][opp q], Fn_q = F[n][q], Xnn_oq = f[nn][opp q] as seen by the reference's sweep,
 * eu = ex*u_wall_x + ey*u_wall_y at the SOLID node.  `keep` is returned when delta <= 0.
 * (`2 * d`, `3 * (w / c)` ...: int literals, evaluated in `real`; the comparisons with 0.5 in double.) */
template <typename real>
LBM_HD real bounce_value(const Lattice<real> &L, int q, real d, real Fn_oq, real Fn_q, real Xnn_oq, real eu,
                         real keep) {
  real v = keep;
  const real d2 = x_mul((real)2, d), wc = x_div(L.w[q], L.c);
  if (d >= 0.5)
    v = x_add(x_add(x_div(Fn_oq, d2), x_div(x_mul(x_sub(d2, (real)1), Fn_q), d2)), x_div(x_mul(x_mul((real)3, wc), eu), d));
  if (d > 0. && d < 0.5)
    v = x_add(x_add(x_mul(d2, Fn_oq), x_mul(x_sub((real)1, d2), Xnn_oq)), x_mul(x_mul((real)6, wc), eu));
  return v;
}

/* Sweeps 1-2 of one LBM step at one node, in registers: reinit_obst_density (:966-986) where the
 * PREVIOUS step's map is solid, then the MRT collision (:1077-1119) where THIS step's map is
 * fluid.  Ring nodes are touched by neither sweep.  `grains` are this step's records. */
template <typename real>
LBM_HD void reinit_collide(const Lattice<real> &L, const GrainRec<real> *grains, int cell_prev, int cell_now, int x, int y,
                           real *p) {
  if (!cell_is_fluid(cell_prev)) equilibrium(L, grains[cell_obst(cell_prev)], x, y, p);
  if (cell_is_fluid(cell_now)) mrt_collide(L, p);
}

/* The part of sweep 4 that needs no neighbour population: at an active solid node every link
 * whose neighbour is not fluid gets the rest value, f[s][q] = w[q] (:1161-1162, :1192-1193).  The
 * fused kernel applies it while the node is in registers -- except next to the wall ring, where
 * the ring sweep (which the reference runs BEFORE the grain sweep) still has to read the old
 * value; those few links go through the sweep kernel's list instead. */
template <typename real>
LBM_HD bool w_links_with_collide(const Lattice<real> &L, int x, int y) {
  return x >= 2 && y >= 2 && x <= L.lx - 3 && y <= L.ly - 3;
}

/* Dead populations.  A node that is solid under BOTH maps, carries neither CELL_ACT nor CELL_RIM and does not
 * touch the wall ring holds the equilibrium of reinit_obst_density (a pure function of the two maps and this
 * step's grain records), and nothing on the path reads it before it is overwritten:
 *   - the next step's re-init overwrites the node itself (it is solid under what is then the old map);
 *   - only a node that is fluid under this step's map uses what it pulls from a neighbour, and a solid node with
 *     a fluid neighbour is active (the neighbour was fluid when the owner was rasterised, too);
 *   - the ring sweep reads rows / columns 1 and lx-2 / ly-2; the bounce-back sweep reads fluid, ring and active
 *     nodes; forces_fluid reads solid nodes with a foreign neighbour: active if that neighbour is fluid, CELL_RIM
 *     (on both sides of the link) if it is not.
 * The fused kernel therefore leaves such nodes unwritten (about a third of the lattice in a dense packing); the
 * reference's values are materialised on demand -- before anything OBSERVES the populations (get_f, fields,
 * density sum, checkpoint) -- by fill_dead (lbm_kernels.cu), from the same expression the fused kernel uses. */
template <typename real>
LBM_HD bool node_is_dead(const Lattice<real> &L, int cell_prev, int cell_now, int x, int y) {
#if defined(LBMDEM_DEAD_GROUP) && LBMDEM_DEAD_GROUP == 0 /* measurement variant: write every node */
  (void)L; (void)cell_prev; (void)cell_now; (void)x; (void)y;
  return false;
#else
  return cell_prev >= 0 && cell_now >= 0 && (cell_now & (CELL_ACT | CELL_RIM)) == 0 && w_links_with_collide(L, x, y);
#endif
}

/* ------------------------------------------------------------------------------------------
 * Sweeps 3-4 (wall ring, grain bounce-back) rewrite a SPARSE set of populations in place: ring
 * nodes and active solid nodes.  They run as separate small kernels on the stored state; after
 * them the array holds exactly what the reference holds before its swap passes, and sweep 5
 * (streaming) is a plain pull (pull_plain).
 * ---------------------------------------------------------------------------------------- */
template <typename real>
LBM_HD real A_value(const Lattice<real> &L, const Stored<real> &S, int x, int y, int q) {
  return S.A[q * L.plane + node_index(L, x, y)];
}

/* Sweep 3, the wall-ring copies (:1123-1145), in the reference's order and therefore in two
 * passes that each run in place, one thread per ring node:
 *   pass 0  rows y = 0 / y = ly-1, x = 1..lx-2  (reads interior nodes and, at the row ends,
 *           column ring nodes that pass 0 does not write);
 *   pass 1  columns x = 0 / x = lx-1, y = 1..ly-2 (reads interior nodes and, at the column
 *           ends, row ring nodes as pass 0 left them), and the four corners.
 * Returns the new content of population q of ring node (x,y); populations the sweep does not
 * assign come back unchanged. */
template <typename real>
LBM_HD real ring_value(const Lattice<real> &L, const Stored<real> &S, int pass, int x, int y, int q) {
  const int lx = L.lx, ly = L.ly;
  const bool xin = x >= 1 && x <= lx - 2, yin = y >= 1 && y <= ly - 2;
  if (pass == 0) {
    if (y == 0 && xin) {
      if (q == 8) return A_value(L, S, x, 1, 4);
      if (q == 7) return A_value(L, S, x + 1, 1, 3);
      if (q == 1) return A_value(L, S, x - 1, 1, 5);
    } else if (y == ly - 1 && xin) {
      if (q == 4) return A_value(L, S, x, ly - 2, 8);
      if (q == 3) return A_value(L, S, x - 1, ly - 2, 7) - L.lid6;
      if (q == 5) return A_value(L, S, x + 1, ly - 2, 1) + L.lid6;
    }
  } else {
    if (x == 0 && yin) {
      if (q == 6) return A_value(L, S, 1, y, 2);
      if (q == 7) return A_value(L, S, 1, y + 1, 3);
      if (q == 5) return A_value(L, S, 1, y - 1, 1);
    } else if (x == lx - 1 && yin) {
      if (q == 2) return A_value(L, S, lx - 2, y, 6);
      if (q == 3) return A_value(L, S, lx - 2, y - 1, 7);
      if (q == 1) return A_value(L, S, lx - 2, y + 1, 5);
    } else if (x == 0 && y == 0) {
      if (q == 7) return A_value(L, S, 1, 1, 3);
    } else if (x == lx - 1 && y == 0) {
      if (q == 1) return A_value(L, S, lx - 2, 1, 5);
    } else if (x == 0 && y == ly - 1) {
      if (q == 5) return A_value(L, S, 1, ly - 2, 1);
    } else if (x == lx - 1 && y == ly - 1) {
      if (q == 3) return A_value(L, S, lx - 2, ly - 2, 7);
    }
  }
  return A_value(L, S, x, y, q);
}

/* is (x,y) an interior solid node with act == 1 ? */
template <typename real>
LBM_HD bool is_active_solid(const Lattice<real> &L, const Stored<real> &S, int x, int y) {
  if (is_ring(L, x, y)) return false;
  const int c = S.cell[node_index(L, x, y)];
  return !cell_is_fluid(c) && node_act(L, S, x, y, c);
}

/* Sweep 4, one link (s = (x,y) an ACTIVE interior solid node, q = 1..8) of the grain bounce-back
 * sweep (:1154-1222).  The reference runs it serially, x outer, y inner, in place; the only read
 * that can see another link's result is X = f[nn][opp q] of a short link (delta < 1/2) whose
 * second fluid-side node nn = s + 2 e_q is itself an active solid node (a one-node gap between
 * two grains): the reference then reads the partner's NEW value if nn was swept earlier, its old
 * value otherwise.
 *
 *   SWEEP_KEEP   the link leaves f[s][q] untouched (delta <= 0)
 *   SWEEP_WRITE  *v is the new f[s][q]
 *   SWEEP_DEFER  (only when resolve == false) the link faces an active solid node across a
 *                one-node gap: its partner link (nn, opp q) may read f[s][q], so it must not be
 *                written while other links are still being evaluated.  Calling again with
 *                resolve == true evaluates it from the pre-sweep state alone (the partner's new
 *                value, where the reference would have seen it, is recomputed from that state).
 * Reads: A at fluid nodes, ring nodes (ring sweep already applied), and pre-sweep values of
 * deferred links -- none of which a concurrent in-place pass over the other links modifies. */
enum { SWEEP_KEEP = 0, SWEEP_WRITE = 1, SWEEP_DEFER = 2 };

/* One bounce-back link from its operands alone.  g: the grain that owns the solid node s = (x,y); Fn_oq, Fn_q: the
 * two populations of the fluid neighbour n = s + e_q; X: f[nn][opp q], nn = n + e_q, BEFORE the sweep; gap: nn is an
 * active solid node (of grain gp) across a one-node gap -- then the reference, sweeping x-outer y-inner, has already
 * rewritten X if nn comes before s, and that new value is recomputed here from the pre-sweep operands of the partner
 * link (nn, opp q): its fluid neighbour is n as well, its second fluid-side node is s itself, hence Fs_q = f[s][q]
 * before the sweep.  Returns SWEEP_KEEP (delta <= 0: f[s][q] stays) or SWEEP_WRITE (*v is the new f[s][q]). */
template <typename real>
LBM_HD int link_value(const Lattice<real> &L, const GrainRec<real> &g, int x, int y, int q, real Fn_oq, real Fn_q, real X,
                      bool gap, const GrainRec<real> *gp, real Fs_q, real *v) {
  const real d = link_delta(g, x, y, q);
  if (!(d > 0.)) return SWEEP_KEEP;
  const real eu = link_eu(L, g, x, y, q);
  if (d < 0.5 && gap) {
    const int oq = opp_of(q), nnx = x + 2 * ex_of(q), nny = y + 2 * ey_of(q);
    if (nnx < x || (nnx == x && nny < y)) {
      const real dp = link_delta(*gp, nnx, nny, oq);
      const real eup = link_eu(L, *gp, nnx, nny, oq);
      X = bounce_value(L, oq, dp, Fn_q, Fn_oq, Fs_q, eup, X);
    }
  }
  *v = bounce_value(L, q, d, Fn_oq, Fn_q, X, eu, (real)0);
  return SWEEP_WRITE;
}

/* The link with its operands read from the stored array (the sparse sweep kernels, tests/hostcheck).  `g` is the
 * record of the grain that owns (x,y).  *gap_out tells whether the link faces an active solid node across a one-node
 * gap (the caller of the bounce-back kernel passes resolve = true and files such links in the deferred list itself);
 * *Fn_oq_out is A[n][opp q], which the momentum exchange of the link needs as well (n is a fluid node: the sweep
 * never writes there). */
template <typename real>
LBM_HD int sweep_link_core(const Lattice<real> &L, const Stored<real> &S, const GrainRec<real> &g, int x, int y, int q,
                           bool resolve, real *v, bool n_is_fluid, bool *gap_out, real *Fn_oq_out,
                           bool nn_clear = false /* the caller knows that nn is fluid or wall ring */) {
  const int ex = ex_of(q), ey = ey_of(q), oq = opp_of(q);
  const int nx = x + ex, ny = y + ey;
  const int nnx = nx + ex, nny = ny + ey;
  const size_t ks = node_index(L, x, y), kn = node_index(L, nx, ny);
  /* the loads whose address depends on (x, y, q) alone are issued together */
  const real Fn_q = S.A[q * L.plane + kn], Fn_oq = S.A[oq * L.plane + kn];
  *Fn_oq_out = Fn_oq;
  *gap_out = false;
  if (!n_is_fluid && !cell_is_fluid(S.cell[kn])) { /* :1161-1162 */
    *v = L.w[q];
    return SWEEP_WRITE;
  }
  /* n fluid => n is an interior node => nn lies inside the array.  An interior solid nn is active:
   * its neighbour n is fluid */
  const size_t knn = node_index(L, nnx, nny);
  const int cnn = nn_clear ? CELL_FLUID : S.cell[knn];
  const bool gap = !nn_clear && !is_ring(L, nnx, nny) && !cell_is_fluid(cnn);
  *gap_out = gap;
  if (gap && !resolve) return SWEEP_DEFER;
  /* All three operands are loaded at once although bounce_value uses two (F[n][q] from delta = 1/2 up, f[nn][opp q]
   * below): loading only the one in use puts the grain record and delta in front of the loads, and the sweep is bound
   * by its chain of dependent DRAM latencies, not by sectors (r02o: 110.6 us against 95.5 us). */
  const real X = S.A[oq * L.plane + knn];
  real Fs_q = 0;
  GrainRec<real> gp = g;
  if (gap) {
    Fs_q = S.A[q * L.plane + ks];
    gp = S.grains[cell_obst(cnn)];
  }
  return link_value(L, g, x, y, q, Fn_oq, Fn_q, X, gap, &gp, Fs_q, v);
}

template <typename real>
LBM_HD_SLOW int sweep_link(const Lattice<real> &L, const Stored<real> &S, int x, int y, int q, bool resolve, real *v,
                           int grain = -1 /* owner of (x,y) if the caller knows it */,
                           bool n_is_fluid = false /* the caller knows that the neighbour is fluid */) {
  if (grain < 0) grain = cell_obst(S.cell[node_index(L, x, y)]);
  const GrainRec<real> g = S.grains[grain];
  bool gap;
  real Fn_oq;
  return sweep_link_core(L, S, g, x, y, q, resolve, v, n_is_fluid, &gap, &Fn_oq);
}

/* Sweep 5, the streamed value: what the two swap passes leave in f[x][y][q] (:1224-1242), from
 * the array as it stands after sweeps 1-4. */
template <typename real>
LBM_HD real pull_plain(const Lattice<real> &L, const real *A, int x, int y, int q) {
  const int sx = x - ex_of(q), sy = y - ey_of(q);
  if (q == 0 || !in_array(L, sx, sy)) return A[opp_of(q) * L.plane + node_index(L, x, y)];
  return A[q * L.plane + node_index(L, sx, sy)];
}

/* hydrodynamic-force sums of the default build are 64-bit fixed point: integer adds commute, so
 * the sum does not depend on the order in which lanes / GPUs contribute */
constexpr double FORCE_FIX = 4503599627370496.0;   /* 2^52 : fhf1, fhf2 (|sum| < 2^11) */
constexpr double TORQUE_FIX = 281474976710656.0;   /* 2^48 : fhf3       (|sum| < 2^15) */

/* One boundary link of forces_fluid (src/main.c:1313-1320).  fs_oq = f_new[s][opp q] and
 * fn_q = f_new[n][q] (both POST-stream), s = (x,y) a node owned by the grain, n = s + e_q a
 * node NOT owned by it.  Accumulates in the reference's expression order. */
template <typename real>
LBM_HD void force_link(int q, real fs_oq, real fn_q, int x, int y, real xc, real yc, real *fh1, real *fh2, real *fh3) {
  const int oq = opp_of(q);
  const real fnx = (fs_oq + fn_q) * ex_of(oq);
  const real fny = (fs_oq + fn_q) * ey_of(oq);
  *fh1 = *fh1 + fnx;
  *fh2 = *fh2 + fny;
  *fh3 = *fh3 - fnx * (y - yc) + fny * (x - xc);
}

}  // namespace lbm
