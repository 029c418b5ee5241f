/*
 * lbm_kernels.cu -- K1, the fused LBM step for sm_100a.
 *
 * The device keeps, between LBM steps, the populations of the last step after its re-init and
 * collide sweeps ("A", lbm_node.cuh).  One fused launch then does, per node,
 *     sweeps 3-5 of the stored step  (wall ring :1123-1145, interpolated grain bounce-back
 *                                     :1154-1222, streaming :1224-1242)      -- a PULL from A,
 *     sweeps 1-2 of the new step     (reinit_obst_density :966-986, MRT collide :1077-1119)
 * and writes the new A.  Every value a node needs from its neighbours is already stored, so no
 * node is collided twice and the nine planes are read once and written once per step.
 *
 * lbm_rows_kernel (the hot kernel; nodes at least two away from the array edge).  A CTA owns TY
 * consecutive y-columns and a contiguous range of rows and marches along x.  One elected thread
 * feeds a ring of NS shared-memory slots with TMA (cp.async.bulk.tensor): per lattice row one
 * 3-D box of the nine population planes (TY nodes + halo), one 2-D box of the stored step's
 * obstacle map (+ halo) and one of this step's map, all completing on the slot's mbarrier.
 * Thread j computes node (x, y0 + j): it waits for row x+1, pulls its nine populations from
 * rows x-1, x, x+1 in shared memory (a solid source gives the interpolated bounce-back value,
 * evaluated from the fluid end of the link with delta computed on the fly), re-initialises /
 * collides in registers and stores nine coalesced values.  After a CTA barrier the slot of row
 * x-1 is refilled with row x-1+NS, so NS-3 rows per CTA are always in flight.
 *
 * lbm_slow_kernel evaluates nodes through the on-demand path of lbm_node.cuh from global memory:
 * the O(perimeter) nodes on or next to the wall ring every step (ring ordering rules), every
 * node when used as the cross-check of the row kernel (params.kernel = 1), and the stream-only
 * pass that materialises the reference's f[x][y][q] for output.
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

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

/* A link whose source node s = p - e_q is solid while p is fluid: the value that streams into p
 * is the interpolated bounce-back value of (s, q), :1166-1185, evaluated by p.  Rare (grain
 * surfaces only), so it is ONE out-of-line routine with a run-time q.  rowA/rowC address the
 * shared-memory rows x-1, x, x+1 through their slot bases. */
template <typename real>
__device__ __noinline__ real bounce_pull(const FusedArgs<real> &a, const unsigned char *smem, int slot_m, int slot_0,
                                         int slot_p, int gx, int gy, int jy, int q, int cs, real As_q) {
  using C = RowCfg<real>;
  const Lattice<real> &L = a.L;
  const int ex = ex_of(q), ey = ey_of(q), oq = opp_of(q);
  const int by = jy + C::HY, cy = jy + C::HC;
  const real *A0 = reinterpret_cast<const real *>(smem + (size_t)slot_0 * C::SLOT);
  const GrainRec<real> g = a.S.grains[cell_obst(cs)];
  const int sx = gx - ex, sy = gy - ey;
  const real d = link_delta(g, sx, sy, q);
  const real eu = ex * wall_ux(L, g, sy) + ey * wall_uy(L, g, sx);
  const real Fn_oq = A0[oq * C::BY + by], Fn_q = A0[q * C::BY + by];
  real X = 0;
  if (d > 0. && d < 0.5) {
    /* second fluid-side node nn = p + e_q, in row x + ex */
    const int slot_nn = ex > 0 ? slot_p : (ex < 0 ? slot_m : slot_0);
    const unsigned char *base = smem + (size_t)slot_nn * C::SLOT;
    const int cnn = reinterpret_cast<const int *>(base + C::A_PAD)[cy + ey];
    const int nnx = gx + ex, nny = gy + ey;
    if (cell_is_act(cnn) && (nnx < sx || (nnx == sx && nny < sy)))
      X = G_value<real, true>(L, a.S, nnx, nny, oq); /* serial-sweep look-back, ~1 link per step */
    else
      X = reinterpret_cast<const real *>(base)[oq * C::BY + by + ey];
  }
  return bounce_value(L, q, d, Fn_oq, Fn_q, X, eu, As_q);
}

template <typename real>
__global__ void __launch_bounds__(RowCfg<real>::TY) lbm_rows_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmCo,
                                                                    const __grid_constant__ CUtensorMap tmCn,
                                                                    const __grid_constant__ FusedArgs<real> a) {
  using C = RowCfg<real>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[C::NS];

  const Lattice<real> &L = a.L;
  const int jy = threadIdx.x;
  const int y0 = blockIdx.x * C::TY;
  const int gy = y0 + jy;
  /* rows of this CTA: a balanced share of the hot rows [R0, R1) */
  const int R0 = max(a.xlo, 2), R1 = min(a.xhi, L.lx - 2);
  const int r0 = R0 + (int)((long long)(R1 - R0) * blockIdx.y / gridDim.y);
  const int r1 = R0 + (int)((long long)(R1 - R0) * (blockIdx.y + 1) / gridDim.y);
  if (r1 <= r0) return;
  const int nload = r1 - r0 + 2; /* rows r0-1 .. r1; loaded row t is global row r0 - 1 + t */

  auto issue = [&](int t) {
    const int slot = t % C::NS;
    unsigned char *base = smem + (size_t)slot * C::SLOT;
    const int row = r0 - 1 + t - L.x0; /* local row */
    mbar_expect_tx(&full[slot], (uint32_t)(C::A_BYTES + C::CO_BYTES + C::CN_BYTES));
    tma_load_3d(base, &tmA, &full[slot], y0 - C::HY, row, 0);
    tma_load_2d(base + C::A_PAD, &tmCo, &full[slot], y0 - C::HC, row);
    tma_load_2d(base + C::A_PAD + C::CO_PAD, &tmCn, &full[slot], y0, row);
  };

  if (jy == 0) {
#pragma unroll
    for (int s = 0; s < C::NS; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int t = 0; t < C::NS && t < nload; ++t) issue(t);
  }
  __syncthreads();

  const bool active = gy >= 2 && gy <= L.ly - 3;
  const int by = jy + C::HY, cy = jy + C::HC;
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1 % C::NS], 0);

  int slot_m = 0, slot_0 = 1; /* slots of rows t-1 and t */
  for (int t = 1; t <= nload - 2; ++t) {
    const int slot_p = (slot_0 + 1 == C::NS) ? 0 : slot_0 + 1;
    mbar_wait(&full[slot_p], ((t + 1) / C::NS) & 1);
    if (active) {
      const int gx = r0 - 1 + t;
      const real *Am = reinterpret_cast<const real *>(smem + (size_t)slot_m * C::SLOT);
      const real *A0 = reinterpret_cast<const real *>(smem + (size_t)slot_0 * C::SLOT);
      const real *Ap = reinterpret_cast<const real *>(smem + (size_t)slot_p * C::SLOT);
      const int *Cm = reinterpret_cast<const int *>(smem + (size_t)slot_m * C::SLOT + C::A_PAD);
      const int *C0 = reinterpret_cast<const int *>(smem + (size_t)slot_0 * C::SLOT + C::A_PAD);
      const int *Cp = reinterpret_cast<const int *>(smem + (size_t)slot_p * C::SLOT + C::A_PAD);
      const int cp = C0[cy];
      const int cnow = reinterpret_cast<const int *>(smem + (size_t)slot_0 * C::SLOT + C::A_PAD + C::CO_PAD)[jy];
      real f[NQ];
      if (cell_is_fluid(cp)) {
        unsigned bb = 0; /* links whose source is a solid node */
        f[0] = A0[by];
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
          const int ex = ex_of(q), ey = ey_of(q);
          const real *As = ex > 0 ? Am : (ex < 0 ? Ap : A0); /* source row x - ex */
          const int *Cs = ex > 0 ? Cm : (ex < 0 ? Cp : C0);
          f[q] = As[q * C::BY + by - ey];
          if (!cell_is_fluid(Cs[cy - ey])) bb |= 1u << q;
        }
        while (bb) {
          const int q = __ffs(bb) - 1;
          bb &= bb - 1;
          const int ex = ex_of(q), ey = ey_of(q);
          const int *Cs = ex > 0 ? Cm : (ex < 0 ? Cp : C0);
          real old = f[1];
#pragma unroll
          for (int k = 2; k < NQ; ++k)
            if (k == q) old = f[k];
          const real v = bounce_pull<real>(a, smem, slot_m, slot_0, slot_p, gx, gy, jy, q, Cs[cy - ey], old);
#pragma unroll
          for (int k = 1; k < NQ; ++k)
            if (k == q) f[k] = v;
        }
      }
      /* a node that is solid under the stored step's map is overwritten by the re-init sweep,
       * whatever streamed into it */
      reinit_collide(L, a.grains_new, cp, cnow, gx, gy, f);
      const size_t k = node_index(L, gx, gy);
#pragma unroll
      for (int q = 0; q < NQ; ++q) a.out[q * L.plane + k] = f[q];
    }
    __syncthreads(); /* every thread is done with row t-1 */
    if (jy == 0 && t - 1 + C::NS < nload) issue(t - 1 + C::NS);
    slot_m = slot_0;
    slot_0 = slot_p;
  }
}

/* nodes the row kernel leaves out: within two nodes of the array edge */
template <typename real>
__device__ __forceinline__ bool is_edge_node(const Lattice<real> &L, int x, int y) {
  return x < 2 || y < 2 || x > L.lx - 3 || y > L.ly - 3;
}

template <typename real>
__global__ void __launch_bounds__(128) lbm_slow_kernel(const __grid_constant__ FusedArgs<real> a, int mode,
                                                       int n_edge_rows_lo, int n_edge_rows_hi) {
  const Lattice<real> &L = a.L;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = a.xhi - a.xlo;
  int x, y;
  if (mode == SLOW_EDGE) {
    /* full edge rows first (lo block, hi block), then four edge columns of the remaining rows */
    const long long full = (long long)(n_edge_rows_lo + n_edge_rows_hi) * L.ly;
    if (t < full) {
      const int r = (int)(t / L.ly);
      y = (int)(t - (long long)r * L.ly);
      x = (r < n_edge_rows_lo) ? a.xlo + r : a.xhi - (n_edge_rows_lo + n_edge_rows_hi - r);
    } else {
      const long long u = t - full;
      const int mid = rows - n_edge_rows_lo - n_edge_rows_hi;
      if (u >= 4ll * mid) return;
      const int r = (int)(u >> 2), c = (int)(u & 3);
      x = a.xlo + n_edge_rows_lo + r;
      y = (c < 2) ? c : L.ly - 4 + c;
    }
  } else {
    if (t >= (long long)rows * L.ly) return;
    const int r = (int)(t / L.ly);
    y = (int)(t - (long long)r * L.ly);
    x = a.xlo + r;
  }
  const size_t k = node_index(L, x, y);
  real f[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) f[q] = pull_value(L, a.S, x, y, q);
  if (mode != SLOW_STREAM_ONLY && !is_ring(L, x, y)) reinit_collide(L, a.grains_new, a.S.cell[k], a.cell_new[k], x, y, f);
#pragma unroll
  for (int q = 0; q < NQ; ++q) a.out[q * L.plane + k] = f[q];
}

template <typename real>
__global__ void __launch_bounds__(128) lbm_h1_kernel(const Lattice<real> L, real *f, const int *cell_prev,
                                                     const int *cell_now, const GrainRec<real> *grains_new, int xlo,
                                                     int xhi) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = xlo + blockIdx.y;
  if (y >= L.ly || x >= xhi || is_ring(L, x, y)) return;
  const size_t k = node_index(L, x, y);
  const int cprev = cell_prev[k], cnow = cell_now[k];
  if (cell_is_fluid(cprev) && !cell_is_fluid(cnow)) return; /* neither sweep touches the node */
  real p[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) p[q] = f[q * L.plane + k];
  reinit_collide(L, grains_new, cprev, cnow, x, y, p);
#pragma unroll
  for (int q = 0; q < NQ; ++q) f[q * L.plane + k] = p[q];
}

template <typename real>
cudaError_t launch_lbm_rows(const CUtensorMap &tmA, const CUtensorMap &tmCo, const CUtensorMap &tmCn,
                            const FusedArgs<real> &a, cudaStream_t s) {
  using C = RowCfg<real>;
  static int resident = 0;
  if (!resident) {
    cudaError_t e = cudaFuncSetAttribute(lbm_rows_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbm_rows_kernel<real>, C::TY, C::SMEM)) != cudaSuccess)
      return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    resident = sms * per_sm;
  }
  const int R0 = a.xlo > 2 ? a.xlo : 2, R1 = a.xhi < a.L.lx - 2 ? a.xhi : a.L.lx - 2;
  if (R1 <= R0) return cudaSuccess;
  const int strips = (a.L.ly + C::TY - 1) / C::TY;
  /* all CTAs co-resident (one wave), rows split evenly between the CTAs of a strip */
  int chunks = resident / strips;
  if (chunks < 1) chunks = 1;
  if (chunks > R1 - R0) chunks = R1 - R0;
  dim3 grid(strips, chunks);
  lbm_rows_kernel<real><<<grid, C::TY, C::SMEM, s>>>(tmA, tmCo, tmCn, a);
  return cudaGetLastError();
}

template <typename real>
cudaError_t launch_lbm_slow(const FusedArgs<real> &a, int mode, cudaStream_t s) {
  const int rows = a.xhi - a.xlo;
  if (rows <= 0) return cudaSuccess;
  int lo = 0, hi = 0;
  long long total;
  if (mode == SLOW_EDGE) {
    /* owned rows among global rows {0, 1} and {lx-2, lx-1} */
    for (int x = a.xlo; x < a.xhi && x < 2; ++x) ++lo;
    for (int x = a.xhi - 1; x >= a.xlo && x > a.L.lx - 3 && x >= 2; --x) ++hi;
    total = (long long)(lo + hi) * a.L.ly + 4ll * (rows - lo - hi);
  } else {
    total = (long long)rows * a.L.ly;
  }
  lbm_slow_kernel<real><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(a, mode, lo, hi);
  return cudaGetLastError();
}

template <typename real>
cudaError_t launch_lbm_h1(const Lattice<real> &L, real *f, const int *cell_prev, const int *cell_now,
                          const GrainRec<real> *grains_new, int xlo, int xhi, cudaStream_t s) {
  if (xhi <= xlo) return cudaSuccess;
  dim3 grid((L.ly + 127) / 128, xhi - xlo);
  lbm_h1_kernel<real><<<grid, 128, 0, s>>>(L, f, cell_prev, cell_now, grains_new, xlo, xhi);
  return cudaGetLastError();
}

#define INSTANTIATE_K1(real)                                                                                          \
  template cudaError_t launch_lbm_rows<real>(const CUtensorMap &, const CUtensorMap &, const CUtensorMap &,           \
                                             const FusedArgs<real> &, cudaStream_t);                                 \
  template cudaError_t launch_lbm_slow<real>(const FusedArgs<real> &, int, cudaStream_t);                             \
  template cudaError_t launch_lbm_h1<real>(const Lattice<real> &, real *, const int *, const int *,                   \
                                           const GrainRec<real> *, int, int, cudaStream_t);
INSTANTIATE_K1(float)
INSTANTIATE_K1(double)

}  // namespace K1_NS
}  // namespace lbmdem
