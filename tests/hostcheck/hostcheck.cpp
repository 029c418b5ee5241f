/*
 * tests/hostcheck/hostcheck.cpp -- TEST INFRASTRUCTURE, never linked into the product.
 *
 * Compiles the product's host+device node headers (csrc/lbm_node.cuh, raster_node.cuh,
 * dem_node.cuh) with g++ and drives them with plain serial loops, so that the *formulation*
 * the CUDA kernels implement -- the stored-state / on-demand restatement of the LBM step, the
 * atomicMax-style rasteriser with the act rule, the gather-form DEM step over a sorted full
 * neighbour list -- can be pinned against the oracle on a machine without a GPU
 * (tests/test_hostcheck.py).  It is not a CPU fallback: nothing under 2d-lbm-dem_b200/ loads it.
 *
 * Build: g++ -O2 -ffp-contract=off -shared -fPIC (tests/hostcheck/build.py).
 */
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "dem_node.cuh"
#include "lbm_node.cuh"
#include "raster_node.cuh"

using namespace lbm;

namespace {

int g_last_deferred = 0;
int g_last_dead = 0;
int g_act_folded = 1; /* 0: hand the on-demand path a bare obstacle map (what the device kernels get) */

/* K2 on the host: owner = highest-index covering grain, then the act rule. */
template <typename real>
void raster_host(int lx, int ly, int n, const RasterParams<real> &P, const real *x1, const real *x2, const real *r,
                 const real *rLB, const real *v1, const real *v2, const real *v3, std::vector<int> &cell,
                 std::vector<GrainRec<real>> &rec, std::vector<GrainBox> &box, std::vector<real> &R2) {
  cell.assign((size_t)lx * ly, -1);
  std::vector<int> min_owner((size_t)lx * ly, -1); /* lowest covering index of multiply covered nodes */
  for (int x = 0; x < lx; ++x) cell[(size_t)x * ly] = cell[(size_t)x * ly + ly - 1] = n;
  for (int y = 0; y < ly; ++y) cell[y] = cell[(size_t)(lx - 1) * ly + y] = n;
  rec.resize(n);
  box.resize(n);
  R2.resize(n);
  for (int i = 0; i < n; ++i) {
    GrainRec<real> &g = rec[i];
    grain_geometry(P, x1[i], x2[i], r[i], rLB[i], &g.xc, &g.yc, &g.r2, &R2[i], &box[i]);
    g.x1 = x1[i]; g.x2 = x2[i]; g.v1 = v1[i]; g.v2 = v2[i]; g.v3 = v3[i];
    for (int x = box[i].xi; x <= box[i].xf; ++x)
      for (int y = box[i].yi; y <= box[i].yf; ++y)
        if (disc_covers(g.xc, g.yc, g.r2, R2[i], x, y)) {
          int &c = cell[(size_t)x * ly + y];
          if (c >= 0) { int &m = min_owner[(size_t)x * ly + y]; m = (m < 0) ? std::min(c, i) : std::min(m, i); }
          c = std::max(c, i);
        }
  }
  for (int i = 0; i < n; ++i) {
    const GrainRec<real> &g = rec[i];
    for (int x = box[i].xi; x <= box[i].xf; ++x)
      for (int y = box[i].yi; y <= box[i].yf; ++y) {
        if (cell_obst(cell[(size_t)x * ly + y]) != i) continue;
        bool act = false;
        for (int q = 1; q < NQ; ++q) {
          const int nx = x + ex_of(q), ny = y + ey_of(q);
          if (fluid_when_grain_ran_exact(cell[(size_t)nx * ly + ny], i, n, min_owner[(size_t)nx * ly + ny])) act = true;
        }
        /* CELL_RIM: a NON-fluid neighbour under another owner (the boundary kernel's list criterion) */
        bool rim = false;
        for (int q = 1; q < NQ; ++q) {
          const int cn = cell[(size_t)(x + ex_of(q)) * ly + y + ey_of(q)];
          if (!cell_is_fluid(cn) && cell_obst(cn) != i) rim = true;
        }
        if (act || rim) cell[(size_t)x * ly + y] = i | (act ? CELL_ACT : 0) | (rim ? CELL_RIM : 0);
      }
  }
}

/* K2 in its TILE form (csrc/aux_kernels.cu: grain_bin_kernel + raster_tile_kernel), restated serially: grains binned
 * by the tiles their clamped box touches (one halo node all round), every tile painted on its own -- here in REVERSE
 * list order, the device's order being arbitrary -- with "highest index wins, the lower of two indices that meet is
 * kept", then act / rim bits from the tile's own copy of the neighbourhood.  Must give raster_host's map bit for bit
 * on the rows [x0+1, x0+nxl-2] of a strip holding the local rows [x0, x0+nxl), and the link / boundary-node sets that
 * follow from it.  Returns 0, or a negative code naming what differed. */
constexpr int HC_RTX = 32, HC_RTY = 64; /* kernels.h: RTX, RTY */

template <typename real>
int raster_tiles_check(int lx, int ly, int n, const double *scal, const double *grains, int x0, int nxl, long *counts) {
  std::vector<real> x1(n), x2(n), v1(n), v2(n), v3(n), r(n), rLB(n);
  for (int i = 0; i < n; ++i) {
    const double *g = grains + 7 * (size_t)i;
    x1[i] = (real)g[0]; x2[i] = (real)g[1]; v1[i] = (real)g[2]; v2[i] = (real)g[3]; v3[i] = (real)g[4];
    r[i] = (real)g[5]; rLB[i] = (real)g[6];
  }
  RasterParams<real> RP;
  RP.lx = lx; RP.ly = ly; RP.dx = (real)scal[0]; RP.Mgx = (real)scal[2]; RP.Mby = (real)scal[3];
  std::vector<int> ref;
  std::vector<GrainRec<real>> rec;
  std::vector<GrainBox> box;
  std::vector<real> R2v;
  raster_host(lx, ly, n, RP, x1.data(), x2.data(), r.data(), rLB.data(), v1.data(), v2.data(), v3.data(), ref, rec, box, R2v);

  /* grain_bin_kernel */
  const int ntx = (nxl + HC_RTX - 1) / HC_RTX, nty = (ly + HC_RTY - 1) / HC_RTY;
  std::vector<std::vector<int>> bins((size_t)ntx * nty);
  for (int i = 0; i < n; ++i) {
    const GrainBox &b = box[i];
    const int xa = std::max(b.xi - 1, x0), xb = std::min(b.xf + 1, x0 + nxl - 1);
    const int ya = std::max(b.yi - 1, 0), yb = std::min(b.yf + 1, ly - 1);
    if (b.xf < b.xi || b.yf < b.yi || xb < xa || yb < ya) continue;
    for (int tx = (xa - x0) / HC_RTX; tx <= (xb - x0) / HC_RTX; ++tx)
      for (int ty = ya / HC_RTY; ty <= yb / HC_RTY; ++ty) bins[(size_t)tx * nty + ty].push_back(i);
  }
  /* raster_tile_kernel */
  const int TR = HC_RTX + 2, TC = HC_RTY + 2;
  std::vector<int> own(TR * TC), low(TR * TC);
  long nlinks = 0, nrim = 0, nlinks_ref = 0, nrim_ref = 0, maxbin = 0;
  for (int tx = 0; tx < ntx; ++tx)
    for (int ty = 0; ty < nty; ++ty) {
      const int tx0 = x0 + tx * HC_RTX, ty0 = ty * HC_RTY;
      for (int rr = 0; rr < TR; ++rr)
        for (int cc = 0; cc < TC; ++cc) {
          const int gx = tx0 - 1 + rr, gy = ty0 - 1 + cc;
          own[rr * TC + cc] = (gx <= 0 || gx >= lx - 1 || gy <= 0 || gy >= ly - 1) ? n : -1;
          low[rr * TC + cc] = 0x7fffffff;
        }
      const std::vector<int> &lst = bins[(size_t)tx * nty + ty];
      maxbin = std::max<long>(maxbin, (long)lst.size());
      for (size_t k = lst.size(); k-- > 0;) { /* reverse order */
        const int i = lst[k];
        const GrainBox &b = box[i];
        const int ra = std::max(b.xi, tx0 - 1), rb = std::min(b.xf, tx0 + HC_RTX);
        const int ca = std::max(b.yi, ty0 - 1), cb = std::min(b.yf, ty0 + HC_RTY);
        for (int x = ra; x <= rb; ++x)
          for (int y = ca; y <= cb; ++y)
            if (disc_covers(rec[i].xc, rec[i].yc, rec[i].r2, R2v[i], x, y)) {
              int &o = own[(x - tx0 + 1) * TC + (y - ty0 + 1)], &l = low[(x - tx0 + 1) * TC + (y - ty0 + 1)];
              const int old = o;
              o = std::max(o, i);
              if (old >= 0 && old != i) l = std::min(l, std::min(old, i));
            }
      }
      for (int rr = 0; rr < HC_RTX; ++rr)
        for (int cc = 0; cc < HC_RTY; ++cc) {
          const int x = tx0 + rr, y = ty0 + cc;
          if (x > x0 + nxl - 1 || y >= ly) continue;
          const int v = own[(rr + 1) * TC + cc + 1];
          int out = v;
          if (v >= 0 && v < n && x >= x0 + 1 && x <= x0 + nxl - 2) {
            bool act = false;
            unsigned foreign = 0, fluid = 0;
            for (int q = 1; q < NQ; ++q) {
              const int at = (rr + 1 + ex_of(q)) * TC + cc + 1 + ey_of(q);
              const int cn = own[at];
              if (cell_is_fluid(cn)) fluid |= 1u << (q - 1);
              if (cn != v) {
                foreign |= 1u << (q - 1);
                int mo = -1;
                if (cn > v && cn < n) mo = low[at] == 0x7fffffff ? -1 : low[at];
                if (fluid_when_grain_ran_exact(cn, v, n, mo)) act = true;
              }
            }
            const unsigned solid_foreign = foreign & ~fluid;
            out = v | (act ? CELL_ACT : 0) | (solid_foreign ? CELL_RIM : 0);
            if (solid_foreign) ++nrim;
            if (act) {
              const bool near_ring = !(x >= 2 && y >= 2 && x <= lx - 3 && y <= ly - 3);
              for (unsigned m = near_ring ? 0xffu : fluid; m; m &= m - 1) ++nlinks;
            }
          }
          const int want = ref[(size_t)x * ly + y];
          const bool classified = x >= x0 + 1 && x <= x0 + nxl - 2;
          if (cell_obst(out) != cell_obst(want) || (out < 0) != (want < 0)) return -1; /* owner */
          if (classified && out != want) return (out ^ want) & CELL_ACT ? -2 : -3;     /* act / rim bit */
        }
    }
  /* the sets that follow from the reference map on the classified rows */
  for (int x = std::max(x0 + 1, 1); x <= std::min(x0 + nxl - 2, lx - 2); ++x)
    for (int y = 1; y <= ly - 2; ++y) {
      const int c = ref[(size_t)x * ly + y];
      if (c < 0 || cell_obst(c) >= n) continue;
      if (c & CELL_RIM) ++nrim_ref;
      if (!(c & CELL_ACT)) continue;
      const bool near_ring = !(x >= 2 && y >= 2 && x <= lx - 3 && y <= ly - 3);
      for (int q = 1; q < NQ; ++q)
        if (near_ring || cell_is_fluid(ref[(size_t)(x + ex_of(q)) * ly + y + ey_of(q)])) ++nlinks_ref;
    }
  counts[0] = nlinks; counts[1] = nrim; counts[2] = maxbin;
  if (nlinks != nlinks_ref) return -4;
  if (nrim != nrim_ref) return -5;
  return 0;
}

template <typename real>
int lbm_step_host(int lx, int ly, int n, const double *scal, const double *grains /* [n][7]: x1 x2 v1 v2 v3 r rLB */,
                  const double *f_in /* [x][y][q] */, const int *obst_old, double *f_out, int *obst_new, int *act_new,
                  double *fhf /* [n][3], unscaled */) {
  std::vector<real> x1(n), x2(n), v1(n), v2(n), v3(n), r(n), rLB(n);
  for (int i = 0; i < n; ++i) {
    const double *g = grains + 7 * (size_t)i;
    x1[i] = (real)g[0]; x2[i] = (real)g[1]; v1[i] = (real)g[2]; v2[i] = (real)g[3]; v3[i] = (real)g[4];
    r[i] = (real)g[5]; rLB[i] = (real)g[6];
  }
  RasterParams<real> RP;
  RP.lx = lx; RP.ly = ly; RP.dx = (real)scal[0]; RP.Mgx = (real)scal[2]; RP.Mby = (real)scal[3];
  std::vector<int> cell_new;
  std::vector<GrainRec<real>> rec;
  std::vector<GrainBox> box;
  std::vector<real> R2v;
  raster_host(lx, ly, n, RP, x1.data(), x2.data(), r.data(), rLB.data(), v1.data(), v2.data(), v3.data(), cell_new, rec,
              box, R2v);
  std::vector<int> cell_bare(cell_new);
  for (auto &c : cell_bare)
    if (c >= 0) c &= CELL_IDX;

  const size_t nn = (size_t)lx * ly;
  std::vector<real> fs(nn * NQ), fn(nn * NQ);
  for (size_t k = 0; k < nn; ++k)
    for (int q = 0; q < NQ; ++q) fs[q * nn + k] = (real)f_in[k * NQ + q];
  std::vector<int> cell_old(obst_old, obst_old + nn);

  Lattice<real> L;
  L.lx = lx; L.ly = ly; L.x0 = 0; L.nxl = lx; L.pitch = ly; L.plane = nn; L.ngrains = n;
  L.dx = (real)scal[0]; L.c = (real)scal[1]; L.Mgx = (real)scal[2]; L.Mby = (real)scal[3];
  const real lid = (real)scal[4];
  L.lid6 = lid / 6;
  L.s2 = 1.5; L.s3 = 1.4; L.s5 = 1.5; L.s7 = 1.5; L.s8 = 1.9841; L.s9 = 1.9841;
  const real w0[NQ] = {4. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9};
  for (int q = 0; q < NQ; ++q) L.w[q] = w0[q];

  /* sweeps 1-2, node-local: what the device stores between steps */
  std::vector<real> A(nn * NQ);
  for (int x = 0; x < lx; ++x)
    for (int y = 0; y < ly; ++y) {
      const size_t k = (size_t)x * ly + y;
      real p[NQ];
      for (int q = 0; q < NQ; ++q) p[q] = fs[q * nn + k];
      if (!is_ring(L, x, y)) {
        reinit_collide(L, rec.data(), cell_old[k], cell_new[k], x, y, p);
        /* the fused kernel also applies the w-links of active solid nodes away from the ring */
        if (cell_is_act(cell_new[k]) && w_links_with_collide(L, x, y))
          for (int q = 1; q < NQ; ++q)
            if (!cell_is_fluid(cell_new[(size_t)(x + ex_of(q)) * ly + y + ey_of(q)])) p[q] = L.w[q];
      }
      for (int q = 0; q < NQ; ++q) A[q * nn + k] = p[q];
    }
  /* the fused kernel does not write dead nodes (lbm_node.cuh, node_is_dead): poison them, so that any sweep or
   * force link that read one would show up in the comparison with the oracle */
  std::vector<size_t> dead;
  for (int x = 1; x < lx - 1; ++x)
    for (int y = 1; y < ly - 1; ++y) {
      const size_t k = (size_t)x * ly + y;
      if (!node_is_dead(L, cell_old[k], cell_new[k], x, y)) continue;
      dead.push_back(k);
      for (int q = 0; q < NQ; ++q) A[q * nn + k] = (real)NAN;
    }
  g_last_dead = (int)dead.size();
  Stored<real> S;
  S.A = A.data(); S.grains = rec.data();
  S.cell = g_act_folded ? cell_new.data() : cell_bare.data();
  S.boxes = box.data(); S.R2 = R2v.data(); S.act_folded = g_act_folded;

  /* sweep 3 in place, two passes, one "thread" per ring node, nodes in reverse order */
  for (int pass = 0; pass < 2; ++pass)
    for (int x = lx - 1; x >= 0; --x)
      for (int y = ly - 1; y >= 0; --y) {
        if (!is_ring(L, x, y)) continue;
        real v[NQ];
        for (int q = 1; q < NQ; ++q) v[q] = ring_value(L, S, pass, x, y, q);
        for (int q = 1; q < NQ; ++q) A[q * nn + (size_t)x * ly + y] = v[q];
      }
  /* sweep 4 in place, the way the device does it: links in ARBITRARY order (here: reversed, the
   * opposite of the reference's sweep), gap links through the deferred list */
  std::vector<std::pair<size_t, real>> deferred;
  for (int x = lx - 2; x >= 1; --x)
    for (int y = ly - 2; y >= 1; --y) {
      if (!is_active_solid(L, S, x, y)) continue;
      for (int q = NQ - 1; q >= 1; --q) {
        /* the link list holds bounce links, and w-links only next to the ring */
        if (!cell_is_fluid(S.cell[(size_t)(x + ex_of(q)) * ly + y + ey_of(q)]) && w_links_with_collide(L, x, y)) continue;
        real v;
        int r = sweep_link(L, S, x, y, q, false, &v);
        const size_t e = q * nn + (size_t)x * ly + y;
        if (r == SWEEP_WRITE) A[e] = v;
        else if (r == SWEEP_DEFER && sweep_link(L, S, x, y, q, true, &v) == SWEEP_WRITE) deferred.emplace_back(e, v);
      }
    }
  for (auto &d : deferred) A[d.first] = d.second;
  g_last_deferred = (int)deferred.size();

  /* what the path itself reads stays poisoned; the observable populations need the dead nodes materialised
   * (the fill_dead kernel): the re-init equilibrium of the previous owner */
  const std::vector<real> A_path(A);
  for (size_t k : dead) {
    const int x = (int)(k / ly), y = (int)(k % ly);
    real p[NQ];
    reinit_collide(L, rec.data(), cell_old[k], cell_new[k], x, y, p);
    for (int q = 0; q < NQ; ++q) A[q * nn + k] = p[q];
  }
  /* sweep 5: plain pull */
  for (int x = 0; x < lx; ++x)
    for (int y = 0; y < ly; ++y)
      for (int q = 0; q < NQ; ++q) fn[q * nn + (size_t)x * ly + y] = pull_plain(L, A.data(), x, y, q);

  for (size_t k = 0; k < nn; ++k) {
    for (int q = 0; q < NQ; ++q) f_out[k * NQ + q] = fn[q * nn + k];
    obst_new[k] = cell_obst(cell_new[k]);
    act_new[k] = cell_is_act(cell_new[k]) ? 1 : 0;
  }
  /* forces_fluid on the streamed field, reference order (src/main.c:1295-1325) */
  for (int i = 0; i < n; ++i) {
    real R2;
    GrainBox b;
    real xc, yc, r2;
    grain_geometry(RP, x1[i], x2[i], r[i], rLB[i], &xc, &yc, &r2, &R2, &b);
    real h1 = 0, h2 = 0, h3 = 0;
    for (int x = b.xi; x <= b.xf; ++x)
      for (int y = b.yi; y <= b.yf; ++y) {
        if (cell_obst(cell_new[(size_t)x * ly + y]) != i) continue;
        for (int q = 1; q < NQ; ++q) {
          const int ax = x + ex_of(q), ay = y + ey_of(q);
          if (cell_obst(cell_new[(size_t)ax * ly + ay]) == i) continue;
          /* f_new[s][opp q] = A[n][opp q] and f_new[n][q] = A[s][q]: read before streaming, as the device does */
          const real fs_oq = A_path[opp_of(q) * nn + (size_t)ax * ly + ay], fn_q = A_path[q * nn + (size_t)x * ly + y];
          if (fs_oq != fn[opp_of(q) * nn + (size_t)x * ly + y] || fn_q != fn[q * nn + (size_t)ax * ly + ay]) return -2;
          force_link<real>(q, fs_oq, fn_q, x, y, xc, yc, &h1, &h2, &h3);
        }
      }
    fhf[3 * (size_t)i] = h1; fhf[3 * (size_t)i + 1] = h2; fhf[3 * (size_t)i + 2] = h3;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Strip decomposition, staged the way Sim (csrc/sim.cu) stages it on one rank, so that the ghost
 * widths and sweep ranges can be tested with two CPU processes exchanging rows (gloo):
 *   stage 1  raster on the local rows, sweeps 1-2 (+ w-links) on the OWNED rows
 *   -- exchange GHOST rows of populations with the x-neighbours --
 *   stage 2  ring sweep on [xlo-3, xhi+3), bounce-back sweep on [xlo-1, xhi+1) with the deferred
 *            list, fixed-point force sums over links whose solid node is owned
 *   -- all-reduce (integer sum) of the force sums --
 *   stage 3  plain pull on the owned rows
 * Local arrays cover the rows x0 .. x0+nxl-1 (x0 = xlo - 4 on an inner edge), [q][row][y]. */
template <typename real>
struct StripCtx {
  Lattice<real> L;
  int xlo, xhi;
  std::vector<int> cell;
  std::vector<GrainRec<real>> rec;
  std::vector<GrainBox> box;
  std::vector<real> R2;
};

template <typename real>
void strip_lattice(Lattice<real> &L, int lx, int ly, int n, int x0, int nxl, const double *scal) {
  L.lx = lx; L.ly = ly; L.x0 = x0; L.nxl = nxl; L.pitch = ly; L.plane = (size_t)nxl * ly; L.ngrains = n;
  L.dx = (real)scal[0]; L.c = (real)scal[1]; L.Mgx = (real)scal[2]; L.Mby = (real)scal[3];
  const real lid = (real)scal[4];
  L.lid6 = lid / 6;
  L.s2 = 1.5; L.s3 = 1.4; L.s5 = 1.5; L.s7 = 1.5; L.s8 = 1.9841; L.s9 = 1.9841;
  const real w0[NQ] = {4. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9};
  for (int q = 0; q < NQ; ++q) L.w[q] = w0[q];
}

/* K2 on the local rows: frame, max-owner raster, act fold where the neighbours are local */
template <typename real>
void strip_raster(StripCtx<real> &C, const double *grains) {
  const Lattice<real> &L = C.L;
  const int n = L.ngrains, ly = L.ly;
  RasterParams<real> P;
  P.lx = L.lx; P.ly = L.ly; P.dx = L.dx; P.Mgx = L.Mgx; P.Mby = L.Mby;
  C.cell.assign(L.plane, -1);
  std::vector<int> min_owner(L.plane, -1);
  for (int r = 0; r < L.nxl; ++r)
    for (int y = 0; y < ly; ++y) {
      const int x = L.x0 + r;
      if (x <= 0 || x >= L.lx - 1 || y <= 0 || y >= ly - 1) C.cell[(size_t)r * ly + y] = n;
    }
  C.rec.resize(n); C.box.resize(n); C.R2.resize(n);
  for (int i = 0; i < n; ++i) {
    const double *g = grains + 7 * (size_t)i;
    GrainRec<real> &R = C.rec[i];
    grain_geometry(P, (real)g[0], (real)g[1], (real)g[5], (real)g[6], &R.xc, &R.yc, &R.r2, &C.R2[i], &C.box[i]);
    R.x1 = (real)g[0]; R.x2 = (real)g[1]; R.v1 = (real)g[2]; R.v2 = (real)g[3]; R.v3 = (real)g[4];
    for (int x = std::max(C.box[i].xi, L.x0); x <= std::min(C.box[i].xf, L.x0 + L.nxl - 1); ++x)
      for (int y = C.box[i].yi; y <= C.box[i].yf; ++y)
        if (disc_covers(R.xc, R.yc, R.r2, C.R2[i], x, y)) {
          int &c = C.cell[node_index(L, x, y)];
          if (c >= 0) { int &m = min_owner[node_index(L, x, y)]; m = (m < 0) ? std::min(c, i) : std::min(m, i); }
          c = std::max(c, i);
        }
  }
  for (int i = 0; i < n; ++i) {
    const GrainRec<real> &R = C.rec[i];
    for (int x = std::max(C.box[i].xi, L.x0 + 1); x <= std::min(C.box[i].xf, L.x0 + L.nxl - 2); ++x)
      for (int y = C.box[i].yi; y <= C.box[i].yf; ++y) {
        int &c = C.cell[node_index(L, x, y)];
        if (cell_obst(c) != i) continue;
        bool act = false;
        for (int q = 1; q < NQ; ++q) {
          const int nx = x + ex_of(q), ny = y + ey_of(q);
          if (fluid_when_grain_ran_exact(C.cell[node_index(L, nx, ny)], i, n, min_owner[node_index(L, nx, ny)])) act = true;
        }
        if (act) c |= CELL_ACT;
      }
  }
}

template <typename real>
int strip_stage1(int lx, int ly, int n, int x0, int nxl, int xlo, int xhi, const double *scal, const double *grains,
                 double *f /* in: pre-collision, out: A on the owned rows; [q][row][y] */, const int *cell_old /* local */,
                 int *cell_new_out /* local, act folded */) {
  StripCtx<real> C;
  strip_lattice(C.L, lx, ly, n, x0, nxl, scal);
  strip_raster(C, grains);
  const Lattice<real> &L = C.L;
  for (int x = xlo; x < xhi; ++x)
    for (int y = 0; y < ly; ++y) {
      if (is_ring(L, x, y)) continue;
      const size_t k = node_index(L, x, y);
      real p[NQ];
      for (int q = 0; q < NQ; ++q) p[q] = (real)f[q * L.plane + k];
      reinit_collide(L, C.rec.data(), cell_old[k], C.cell[k], x, y, p);
      if (cell_is_act(C.cell[k]) && w_links_with_collide(L, x, y))
        for (int q = 1; q < NQ; ++q)
          if (!cell_is_fluid(C.cell[node_index(L, x + ex_of(q), y + ey_of(q))])) p[q] = L.w[q];
      for (int q = 0; q < NQ; ++q) f[q * L.plane + k] = p[q];
    }
  for (size_t k = 0; k < L.plane; ++k) cell_new_out[k] = C.cell[k];
  return 0;
}

template <typename real>
int strip_stage2(int lx, int ly, int n, int x0, int nxl, int xlo, int xhi, int nranks, const double *scal,
                 const double *grains, double *A_io /* local, ghosts filled */, long long *facc /* [3][n] */) {
  StripCtx<real> C;
  strip_lattice(C.L, lx, ly, n, x0, nxl, scal);
  strip_raster(C, grains);
  const Lattice<real> &L = C.L;
  std::vector<real> A(L.plane * NQ);
  for (size_t k = 0; k < A.size(); ++k) A[k] = (real)A_io[k];
  Stored<real> S;
  S.A = A.data(); S.cell = C.cell.data(); S.grains = C.rec.data(); S.boxes = C.box.data(); S.R2 = C.R2.data();
  S.act_folded = 1;
  const bool multi = nranks > 1;
  const int ra = multi ? std::max(xlo - 3, 0) : 0, rb = multi ? std::min(xhi + 3, lx) : lx;
  for (int pass = 0; pass < 2; ++pass)
    for (int x = rb - 1; x >= ra; --x)
      for (int y = ly - 1; y >= 0; --y) {
        if (!is_ring(L, x, y)) continue;
        if (pass == 1 && !(x == 0 || x == lx - 1)) continue; /* the device's pass 1 covers the ring ROWS only */
        real v[NQ];
        for (int q = 1; q < NQ; ++q) v[q] = ring_value(L, S, pass, x, y, q);
        for (int q = 1; q < NQ; ++q) A[q * L.plane + node_index(L, x, y)] = v[q];
      }
  const int sa = std::max(multi ? xlo - 1 : xlo, 1), sb = std::min(multi ? xhi + 1 : xhi, lx - 1);
  std::vector<std::pair<size_t, real>> deferred;
  for (int x = sb - 1; x >= sa; --x)
    for (int y = ly - 2; y >= 1; --y) {
      if (!is_active_solid(L, S, x, y)) continue;
      for (int q = NQ - 1; q >= 1; --q) {
        const bool nb_fluid = cell_is_fluid(S.cell[node_index(L, x + ex_of(q), y + ey_of(q))]);
        if (!nb_fluid && w_links_with_collide(L, x, y)) continue;
        real v;
        int r = sweep_link(L, S, x, y, q, false, &v);
        const size_t e = q * L.plane + node_index(L, x, y);
        if (r == SWEEP_WRITE) A[e] = v;
        else if (r == SWEEP_DEFER && sweep_link(L, S, x, y, q, true, &v) == SWEEP_WRITE) deferred.emplace_back(e, v);
      }
    }
  for (auto &d : deferred) A[d.first] = d.second;
  /* fixed-point force sums over links whose solid node is owned */
  for (int k = 0; k < 3 * n; ++k) facc[k] = 0;
  for (int x = std::max(xlo, 1); x < std::min(xhi, lx - 1); ++x)
    for (int y = 1; y < ly - 1; ++y) {
      const size_t k = node_index(L, x, y);
      const int i = cell_obst(S.cell[k]);
      if (i < 0 || i >= n) continue;
      for (int q = 1; q < NQ; ++q) {
        const size_t kn = node_index(L, x + ex_of(q), y + ey_of(q));
        if (cell_obst(S.cell[kn]) == i) continue;
        real h1 = 0, h2 = 0, h3 = 0;
        force_link<real>(q, A[opp_of(q) * L.plane + kn], A[q * L.plane + k], x, y, C.rec[i].xc, C.rec[i].yc, &h1, &h2, &h3);
        facc[i] += llrint((double)h1 * FORCE_FIX);
        facc[n + i] += llrint((double)h2 * FORCE_FIX);
        facc[2 * n + i] += llrint((double)h3 * TORQUE_FIX);
      }
    }
  for (size_t k = 0; k < A.size(); ++k) A_io[k] = A[k];
  return 0;
}

template <typename real>
int strip_stage3(int lx, int ly, int x0, int nxl, int xlo, int xhi, const double *scal, const double *A_in,
                 double *f_out /* [xhi-xlo][ly][9], reference layout */) {
  Lattice<real> L;
  strip_lattice(L, lx, ly, 0, x0, nxl, scal);
  std::vector<real> A(L.plane * NQ);
  for (size_t k = 0; k < A.size(); ++k) A[k] = (real)A_in[k];
  for (int x = xlo; x < xhi; ++x)
    for (int y = 0; y < ly; ++y)
      for (int q = 0; q < NQ; ++q) f_out[((size_t)(x - xlo) * ly + y) * NQ + q] = pull_plain(L, A.data(), x, y, q);
  return 0;
}

/* gather-form DEM step over a sorted FULL neighbour list (what K4 does), serial on the host */
template <typename real>
int dem_step_host(int n, const double *par /* see below */, int film, double *state /* [n][9] x1 x2 x3 v1 v2 v3 a1 a2 a3 */,
                  const double *props /* [n][3] r m It */, const double *fhf /* [n][3] */, int rebuild, int *nbr_count,
                  int *nbr /* [n][cap] */, int cap, int *wflags) {
  dem::Params<real> P;
  P.kg = (real)par[0]; P.kt = (real)par[1]; P.km = (real)par[2]; P.ktm = (real)par[3]; P.nug = (real)par[4];
  P.num = (real)par[5]; P.numb = (real)par[6]; P.nugt = (real)par[7]; P.mu = (real)par[8]; P.mum = (real)par[9];
  P.mumb = (real)par[10]; P.murf = (real)par[11]; P.freq = (real)par[12]; P.amp = (real)par[13]; P.t = (real)par[14];
  P.distVerlet = (real)par[15]; P.dt = (real)par[16]; P.dt2 = (real)par[17]; P.xG = (real)par[18]; P.yG = (real)par[19];
  P.Mgx = (real)par[20]; P.Mdx = (real)par[21]; P.Mby = (real)par[22]; P.Mhy = (real)par[23];
  std::vector<real> s(9 * (size_t)n), pr(3 * (size_t)n), fh(3 * (size_t)n);
  for (size_t k = 0; k < 9 * (size_t)n; ++k) s[k] = (real)state[k];
  for (size_t k = 0; k < 3 * (size_t)n; ++k) { pr[k] = (real)props[k]; fh[k] = (real)fhf[k]; }
  if (rebuild) {
    for (int i = 0; i < n; ++i) {
      int cnt = 0;
      for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const int a = std::min(i, j), b = std::max(i, j);
        if (dem::verlet_pair(P, s[9 * a], s[9 * a + 1], pr[3 * a], s[9 * b], s[9 * b + 1], pr[3 * b])) {
          if (cnt >= cap) return -1;
          nbr[(size_t)i * cap + cnt++] = j; /* ascending j */
        }
      }
      nbr_count[i] = cnt;
      wflags[i] = dem::wall_flags(P, s[9 * i], s[9 * i + 1], pr[3 * i]);
    }
  }
  for (int i = 0; i < n; ++i) {
    real *g = &s[9 * (size_t)i];
    dem::kick_drift(P, &g[0], &g[3], g[6]);
    dem::kick_drift(P, &g[1], &g[4], g[7]);
    dem::kick_drift(P, &g[2], &g[5], g[8]);
  }
  std::vector<real> acc(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    const real *gi = &s[9 * (size_t)i];
    real a1 = fh[3 * i], a2 = fh[3 * i + 1], a3 = fh[3 * i + 2];
    for (int k = 0; k < nbr_count[i]; ++k) {
      const int j = nbr[(size_t)i * cap + k];
      const int a = std::min(i, j), b = std::max(i, j);
      const real *ga = &s[9 * (size_t)a], *gb = &s[9 * (size_t)b];
      dem::Force<real> F;
      if (!dem::pair_force(P, film != 0, ga[0], ga[1], ga[3], ga[4], ga[5], pr[3 * a], gb[0], gb[1], gb[3], gb[4], gb[5],
                           pr[3 * b], &F))
        continue;
      if (i == a) { a1 = a1 + F.f1; a2 = a2 + F.f2; a3 = a3 + F.f3; }
      else        { a1 = a1 - F.f1; a2 = a2 - F.f2; a3 = a3 + F.f3; }
    }
    dem::add_wall_forces(P, wflags[i], gi[0], gi[1], gi[3], gi[4], gi[5], pr[3 * i], &a1, &a2, &a3);
    dem::finish_acceleration(P, pr[3 * i + 1], pr[3 * i + 2], &a1, &a2, &a3);
    acc[3 * i] = a1; acc[3 * i + 1] = a2; acc[3 * i + 2] = a3;
  }
  for (int i = 0; i < n; ++i) {
    real *g = &s[9 * (size_t)i];
    g[6] = acc[3 * i]; g[7] = acc[3 * i + 1]; g[8] = acc[3 * i + 2];
    dem::kick(P, &g[3], g[6]);
    dem::kick(P, &g[4], g[7]);
    dem::kick(P, &g[5], g[8]);
  }
  for (size_t k = 0; k < 9 * (size_t)n; ++k) state[k] = s[k];
  return 0;
}

}  // namespace

extern "C" {
#define EXPORT __attribute__((visibility("default")))
EXPORT void hc_set_act_folded(int v) { g_act_folded = v; }
EXPORT int hc_last_deferred(void) { return g_last_deferred; }
EXPORT int hc_last_dead(void) { return g_last_dead; }
/* scal: dx c Mgx Mby lid; grains [n][7]: x1 x2 v1 v2 v3 r rLB; counts[3]: links, rim nodes, largest tile bin */
EXPORT int hc_raster_tiles_check_f64(int lx, int ly, int n, const double *scal, const double *grains, int x0, int nxl,
                                     long *counts) {
  return raster_tiles_check<double>(lx, ly, n, scal, grains, x0, nxl, counts);
}
EXPORT int hc_raster_tiles_check_f32(int lx, int ly, int n, const double *scal, const double *grains, int x0, int nxl,
                                     long *counts) {
  return raster_tiles_check<float>(lx, ly, n, scal, grains, x0, nxl, counts);
}
/* scal: dx c Mgx Mby lid */
EXPORT int hc_lbm_step_f64(int lx, int ly, int n, const double *scal, const double *grains, const double *f_in,
                           const int *obst_old, double *f_out, int *obst_new, int *act_new, double *fhf) {
  return lbm_step_host<double>(lx, ly, n, scal, grains, f_in, obst_old, f_out, obst_new, act_new, fhf);
}
EXPORT int hc_lbm_step_f32(int lx, int ly, int n, const double *scal, const double *grains, const double *f_in,
                           const int *obst_old, double *f_out, int *obst_new, int *act_new, double *fhf) {
  return lbm_step_host<float>(lx, ly, n, scal, grains, f_in, obst_old, f_out, obst_new, act_new, fhf);
}
EXPORT int hc_strip_stage1_f64(int lx, int ly, int n, int x0, int nxl, int xlo, int xhi, const double *scal,
                               const double *grains, double *f, const int *cell_old, int *cell_new) {
  return strip_stage1<double>(lx, ly, n, x0, nxl, xlo, xhi, scal, grains, f, cell_old, cell_new);
}
EXPORT int hc_strip_stage2_f64(int lx, int ly, int n, int x0, int nxl, int xlo, int xhi, int nranks, const double *scal,
                               const double *grains, double *A, long long *facc) {
  return strip_stage2<double>(lx, ly, n, x0, nxl, xlo, xhi, nranks, scal, grains, A, facc);
}
EXPORT int hc_strip_stage3_f64(int lx, int ly, int x0, int nxl, int xlo, int xhi, const double *scal, const double *A,
                               double *f_out) {
  return strip_stage3<double>(lx, ly, x0, nxl, xlo, xhi, scal, A, f_out);
}
EXPORT int hc_dem_step_f64(int n, const double *par, int film, double *state, const double *props, const double *fhf,
                           int rebuild, int *nbr_count, int *nbr, int cap, int *wflags) {
  return dem_step_host<double>(n, par, film, state, props, fhf, rebuild, nbr_count, nbr, cap, wflags);
}
EXPORT int hc_dem_step_f32(int n, const double *par, int film, double *state, const double *props, const double *fhf,
                           int rebuild, int *nbr_count, int *nbr, int cap, int *wflags) {
  return dem_step_host<float>(n, par, film, state, props, fhf, rebuild, nbr_count, nbr, cap, wflags);
}
}
