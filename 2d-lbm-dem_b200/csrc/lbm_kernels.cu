/*
 * lbm_kernels.cu -- K1, the fused LBM step for sm_100a.
 *
 * One launch replaces reinit_obst_density (src/main.c:966-986), the act/delta part of
 * obst_construction (:1036-1063), collision_streaming (:1071-1243) and the accumulation loop
 * of forces_fluid (:1295-1325).
 *
 * Tiled kernel, per CTA (TILE_X x TILE_Y nodes, 256 threads):
 *   1. one elected thread issues a 3-D TMA load (cp.async.bulk.tensor) of the nine population
 *      planes of the tile plus a one-node halo into shared memory; out-of-array box elements
 *      are zero-filled by the TMA unit and never read;
 *   2. meanwhile all threads stage the obstacle map of the tile plus a two-node halo and fold
 *      the reference's act[][] flag into it (a pure function of the map, :1036-1052);
 *   3. after the mbarrier flips, every node of tile+halo is brought to its state after the
 *      re-init and collide sweeps, in place in shared memory (halo nodes are recomputed here
 *      instead of being exchanged between CTAs);
 *   4. every tile node pulls its nine new populations: a fluid source gives its post-collision
 *      value, a solid source gives the interpolated bounce-back value evaluated from the fluid
 *      side of the link (delta computed on the fly, never stored), and the same link feeds the
 *      grain's momentum-exchange accumulators (64-bit fixed point, order-free);
 *   5. coalesced stores of the nine planes.
 * Nodes within two nodes of the array edge take the exact on-demand path of lbm_node.cuh
 * (wall-ring ordering rules); they are O(perimeter).
 *
 * The generic kernel evaluates EVERY node through that on-demand path; it is slow and exists
 * as the device-side cross-check of the tiled kernel and of the TMA plumbing.
 *
 * This file is compiled twice: with contraction (namespace k1_fast) and with -fmad=false
 * (namespace k1_strict), selected by -DK1_NS=...
 */
#include "kernels.h"

#ifndef K1_NS
#error "compile with -DK1_NS=k1_fast or -DK1_NS=k1_strict"
#endif

namespace lbmdem {
namespace K1_NS {

using namespace lbm;

constexpr int NTHREADS = 256;
constexpr int CELL_OUTSIDE = -2; /* beyond the array: neither fluid nor a grain */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

/* fixed-point momentum exchange for one link: s = (sx,sy) owned by grain i, population q
 * leaves s; Gs_q is what streams out of s, Gn_oq what streams into it (forces_fluid, :1316-1320) */
template <typename real>
__device__ __forceinline__ void add_link_force(const StepArgs<real> &a, int i, int q, real Gn_oq, real Gs_q, int sx,
                                               int sy) {
  const GrainRec<real> &g = a.L.grains[i];
  real h1 = 0, h2 = 0, h3 = 0;
  force_link<real>(q, Gn_oq, Gs_q, sx, sy, g.xc, g.yc, &h1, &h2, &h3);
  const int n = a.L.ngrains;
  if (h1 != 0) atomicAdd((unsigned long long *)&a.facc[i], (unsigned long long)__double2ll_rn((double)h1 * FORCE_FIX));
  if (h2 != 0) atomicAdd((unsigned long long *)&a.facc[n + i], (unsigned long long)__double2ll_rn((double)h2 * FORCE_FIX));
  if (h3 != 0)
    atomicAdd((unsigned long long *)&a.facc[2 * n + i], (unsigned long long)__double2ll_rn((double)h3 * TORQUE_FIX));
}

/* exact on-demand evaluation of one node from global memory (lbm_node.cuh) */
template <typename real>
__device__ __noinline__ void node_on_demand(const StepArgs<real> &a, int x, int y) {
  const Lattice<real> &L = a.L;
  const size_t k = node_index(L, x, y);
  const int cp = L.cell_new[k];
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    const real v = pull_value(L, x, y, q);
    a.f_new[q * L.plane + k] = v;
    if (q == 0 || a.facc == nullptr) continue;
    const int sx = x - ex_of(q), sy = y - ey_of(q);
    if (!in_array(L, sx, sy) || is_ring(L, sx, sy)) continue;
    const int cs = L.cell_new[node_index(L, sx, sy)];
    if (cell_is_fluid(cs)) continue;
    const int i = cell_obst(cs);
    if (cell_obst(cp) == i) continue;
    add_link_force(a, i, q, G_value(L, x, y, opp_of(q)), v, sx, sy);
  }
}

template <typename real>
__global__ void __launch_bounds__(NTHREADS) lbm_generic_kernel(const StepArgs<real> a) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = a.xlo + blockIdx.y;
  if (y >= a.L.ly || x >= a.xhi) return;
  node_on_demand(a, x, y);
}

/* act[x][y] for a solid node whose map entry and eight neighbours are in the shared tile */
template <typename real>
__device__ __forceinline__ bool act_from_tile(const StepArgs<real> &a, const int *sc, int CY, int cx, int cy, int gx,
                                              int gy, int c) {
  const int i = cell_obst(c);
  bool act = false;
#pragma unroll
  for (int q = 1; q < NQ; ++q) {
    const int cn = sc[(cx + ex_of(q)) * CY + cy + ey_of(q)];
    if (cn == CELL_OUTSIDE) continue;
    if (cell_is_fluid(cn)) {
      act = true;
    } else {
      const int k = cell_obst(cn);
      if (k > i && k < a.L.ngrains) { /* a later grain: fluid when grain i ran unless i covers it too */
        const GrainRec<real> &g = a.L.grains[i];
        if (fluid_when_grain_ran(cn, i, a.L.ngrains, g.xc, g.yc, g.r2, a.L.R2[i], a.L.boxes[i], gx + ex_of(q), gy + ey_of(q)))
          act = true;
      }
    }
  }
  return act;
}

template <typename real>
__global__ void __launch_bounds__(NTHREADS, (sizeof(real) == 8) ? 2 : 4)
    lbm_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const StepArgs<real> a) {
  constexpr int BY = TileBox<real>::BY, BX = TileBox<real>::BX, HY = TileBox<real>::HY;
  constexpr int CX = TILE_X + 4, CY = TILE_Y + 4;
  /* node (gx0 + rx, gy0 + ry): population tile [rx + 1][ry + HY], map tile [rx + 2][ry + 2] */
  extern __shared__ __align__(128) unsigned char smem_raw[];
  real *sA = reinterpret_cast<real *>(smem_raw);                       /* [NQ][BX][BY] */
  int *sc = reinterpret_cast<int *>(smem_raw + TileBox<real>::bytes);   /* [CX][CY] */
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + TileBox<real>::bytes + sizeof(int) * CX * CY);

  const Lattice<real> &L = a.L;
  const int tid = threadIdx.x;
  const int gx0 = a.xlo + blockIdx.y * TILE_X; /* global coordinates of the tile origin */
  const int gy0 = blockIdx.x * TILE_Y;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, (uint32_t)TileBox<real>::bytes);
    tma_load_3d(sA, &tmap, bar, gy0 - HY, gx0 - 1 - L.x0, 0);
  }

  /* obstacle map, tile + 2 halo */
  for (int idx = tid; idx < CX * CY; idx += NTHREADS) {
    const int cx = idx / CY, cy = idx - cx * CY;
    const int gx = gx0 - 2 + cx, gy = gy0 - 2 + cy;
    int c = CELL_OUTSIDE;
    if (in_array(L, gx, gy) && gx >= L.x0 && gx < L.x0 + L.nxl) c = L.cell_new[node_index(L, gx, gy)];
    sc[idx] = c;
  }
  __syncthreads();
  /* fold act into the map for tile + 1 halo (readers mask it off with cell_obst) */
  for (int idx = tid; idx < BX * (TILE_Y + 2); idx += NTHREADS) {
    const int bx = idx / (TILE_Y + 2), by = idx - bx * (TILE_Y + 2);
    const int cx = bx + 1, cy = by + 1;
    const int c = sc[cx * CY + cy];
    if (c < 0) continue; /* fluid or outside */
    const int gx = gx0 - 1 + bx, gy = gy0 - 1 + by;
    if (is_ring(L, gx, gy)) continue;
    if (act_from_tile(a, sc, CY, cx, cy, gx, gy, c)) sc[cx * CY + cy] = c | CELL_ACT;
  }
  mbar_wait(bar, 0);
  __syncthreads();

  /* sweeps 1-2 in place: re-init where the old map is solid, collide where the new one is fluid */
  for (int idx = tid; idx < BX * (TILE_Y + 2); idx += NTHREADS) {
    const int bx = idx / (TILE_Y + 2), hy = idx - bx * (TILE_Y + 2);
    const int by = hy + HY - 1;
    const int gx = gx0 - 1 + bx, gy = gy0 - 1 + hy;
    if (!in_array(L, gx, gy) || is_ring(L, gx, gy)) continue;
    if (gx < L.x0 || gx >= L.x0 + L.nxl) continue;
    const int cn = sc[(bx + 1) * CY + hy + 1];
    const int co = L.cell_old[node_index(L, gx, gy)];
    const bool reinit = !cell_is_fluid(co), coll = cell_is_fluid(cn);
    if (!reinit && !coll) continue;
    real p[NQ];
    if (reinit) {
      equilibrium(L, L.grains[cell_obst(co)], gx, gy, p);
    } else {
#pragma unroll
      for (int q = 0; q < NQ; ++q) p[q] = sA[(q * BX + bx) * BY + by];
    }
    if (coll) mrt_collide(L, p);
#pragma unroll
    for (int q = 0; q < NQ; ++q) sA[(q * BX + bx) * BY + by] = p[q];
  }
  __syncthreads();

  /* pull */
  for (int idx = tid; idx < TILE_X * TILE_Y; idx += NTHREADS) {
    const int tx = idx / TILE_Y, ty = idx - tx * TILE_Y;
    const int gx = gx0 + tx, gy = gy0 + ty;
    if (gx >= a.xhi || gy >= L.ly) continue;
    if (gx < 2 || gy < 2 || gx > L.lx - 3 || gy > L.ly - 3) { /* on or next to the wall ring */
      node_on_demand(a, gx, gy);
      continue;
    }
    const int bx = tx + 1, by = ty + HY; /* position in the population tile */
    const int cx = tx + 2, cy = ty + 2;  /* position in the map tile */
    const size_t k = node_index(L, gx, gy);
    const int cp = sc[cx * CY + cy];
    const bool p_fluid = cell_is_fluid(cp);
    a.f_new[k] = sA[(0 * BX + bx) * BY + by];
#pragma unroll
    for (int q = 1; q < NQ; ++q) {
      const int ex = ex_of(q), ey = ey_of(q), oq = opp_of(q);
      const int sbx = bx - ex, sby = by - ey;
      const int cs = sc[(cx - ex) * CY + cy - ey];
      const real As_q = sA[(q * BX + sbx) * BY + sby];
      real v = As_q;
      if (!cell_is_fluid(cs)) {
        const int i = cell_obst(cs);
        real Gp_oq;
        if (p_fluid) {
          /* interpolated bounce-back, evaluated by the fluid end of the link (:1166-1185) */
          const GrainRec<real> g = L.grains[i];
          const int sx = gx - ex, sy = gy - ey;
          const real d = link_delta(g, sx, sy, q);
          const real eu = ex * wall_ux(L, g, sy) + ey * wall_uy(L, g, sx);
          const real Fn_oq = sA[(oq * BX + bx) * BY + by], Fn_q = sA[(q * BX + bx) * BY + by];
          real X = 0;
          if (d > 0. && d < 0.5) {
            const int cnn = sc[(cx + ex) * CY + cy + ey];
            const int nnx = gx + ex, nny = gy + ey;
            if (cell_is_act(cnn) && (nnx < sx || (nnx == sx && nny < sy)))
              X = G_value<real, true>(L, nnx, nny, oq); /* serial-sweep look-back, ~1 link per step */
            else
              X = sA[(oq * BX + bx + ex) * BY + by + ey];
          }
          v = bounce_value(L, q, d, Fn_oq, Fn_q, X, eu, As_q);
          Gp_oq = Fn_oq;
        } else {
          v = cell_is_act(cs) ? L.w[q] : As_q;
          Gp_oq = cell_is_act(cp) ? L.w[oq] : sA[(oq * BX + bx) * BY + by];
        }
        if (a.facc != nullptr && cell_obst(cp) != i) add_link_force(a, i, q, Gp_oq, v, gx - ex, gy - ey);
      }
      a.f_new[q * L.plane + k] = v;
    }
  }
}

template <typename real>
cudaError_t launch_lbm_tiled(const CUtensorMap &tmap, const StepArgs<real> &a, cudaStream_t s) {
  constexpr size_t smem = TileBox<real>::bytes + sizeof(int) * (TILE_X + 4) * (TILE_Y + 4) + 16;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lbm_tiled_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((a.L.ly + TILE_Y - 1) / TILE_Y, (a.xhi - a.xlo + TILE_X - 1) / TILE_X);
  lbm_tiled_kernel<real><<<grid, NTHREADS, smem, s>>>(tmap, a);
  return cudaGetLastError();
}

template <typename real>
cudaError_t launch_lbm_generic(const StepArgs<real> &a, cudaStream_t s) {
  dim3 grid((a.L.ly + NTHREADS - 1) / NTHREADS, a.xhi - a.xlo);
  lbm_generic_kernel<real><<<grid, NTHREADS, 0, s>>>(a);
  return cudaGetLastError();
}

template cudaError_t launch_lbm_tiled<float>(const CUtensorMap &, const StepArgs<float> &, cudaStream_t);
template cudaError_t launch_lbm_tiled<double>(const CUtensorMap &, const StepArgs<double> &, cudaStream_t);
template cudaError_t launch_lbm_generic<float>(const StepArgs<float> &, cudaStream_t);
template cudaError_t launch_lbm_generic<double>(const StepArgs<double> &, cudaStream_t);

}  // namespace K1_NS
}  // namespace lbmdem
