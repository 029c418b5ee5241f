/*
 * lbm_kernels.cu -- K1, the fused LBM step for sm_100a.
 *
 * Between LBM steps the device holds the population array exactly as the reference holds it
 * just before its swap passes: re-init, collide, wall-ring and grain bounce-back sweeps applied
 * (the last two as sparse in-place kernels, aux_kernels.cu).  One fused launch then does, per
 * node,
 *     sweep 5 of the stored step    (streaming, :1224-1242)                 -- a plain PULL,
 *     sweeps 1-2 of the new step    (reinit_obst_density :966-986, MRT collide :1077-1119)
 * and writes the new array: the nine planes are read once and written once per step and no
 * node is collided twice.
 *
 * lbm_rows_kernel (the hot kernel; every interior node).  A CTA owns TY consecutive y-columns
 * and a contiguous range of rows and marches along x.  A ring of NS shared-memory slots is fed
 * with TMA (cp.async.bulk.tensor): per lattice row one 3-D box of the nine
 * population planes (TY nodes + halo) and one 2-D box each of this step's (+ halo) and of the
 * stored step's obstacle map, all completing on the slot's mbarrier.  Thread j computes node
 * (x, y0 + j): it waits for row x+1, pulls its nine populations from rows x-1, x, x+1 in shared
 * memory, re-initialises / collides in registers and stores nine coalesced values.  The queue
 * is warp-specialised: a producer warp issues the TMA loads as slots are released (one "empty"
 * mbarrier per slot, one arrival per consumer warp), so consumer warps never meet at a CTA
 * barrier and up to NS-3 rows per CTA are in flight.
 *
 * lbm_plain_kernel is the same map from global memory, one thread per node: the ring nodes
 * every step (array-edge rule of the swap passes), or every node as the cross-check of the row
 * kernel (params.kernel = 1).
 *
 * This file is compiled twice: with contraction (namespace k1_fast) and with -fmad=false
 * (namespace k1_strict), selected by -DK1_NS=...
 */
#include "kernels.h"

#ifndef K1_NS
#error "compile with -DK1_NS=k1_fast or -DK1_NS=k1_strict"
#endif
#ifndef LBMDEM_K1_SMEM_PAD
#define LBMDEM_K1_SMEM_PAD 0 /* tuning knob: unused dynamic shared memory per CTA, i.e. fewer resident CTAs */
#endif
#ifndef LBMDEM_DEAD_GROUP
#define LBMDEM_DEAD_GROUP 8   /* lanes per skip decision of the row kernel: 8 floats = one 32-byte sector (0: never skip) */
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
#if defined(LBMDEM_K1_LD_EVICT_FIRST) /* tuning variant: the populations are read once per step -> L2 evict-first */
__device__ __forceinline__ void tma_load_3d_ef(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
#endif
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

/* 16 bytes global -> shared without a register round trip (LDGSTS, L2 only), and the arrival on an mbarrier that
 * fires when all of the calling thread's earlier copies have landed (the barrier's count already includes it) */
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

/* TY consumer threads (one node of the row each) + one producer warp that owns the TMA queue -- or, with
 * LBMDEM_K1_NOPROD, no producer warp: the first consumer lane issues the loads of row t+2 at the top of iteration t
 * (the same moment the producer could: when every warp has released row t-2), and the 1920 registers of the idle
 * producer lanes buy two more resident CTAs per SM */
#ifndef LBMDEM_K1_LDGSTS_ALL
#define LBMDEM_K1_LDGSTS_ALL 0 /* measurement: 1 = load every piece (the copy mechanism alone, no skipping) */
#endif
#if defined(LBMDEM_K1_LDGSTS) && !defined(LBMDEM_K1_NOPROD)
#define LBMDEM_K1_NOPROD 1 /* the consumer threads load the rows themselves */
#endif
#if defined(LBMDEM_K1_NOPROD)
#define K1_THREADS(C) (C::TY * C::NB)
#else
#define K1_THREADS(C) (C::TY * C::NB + 32)
#endif
#ifndef LBMDEM_K1_MINB_F32
#define LBMDEM_K1_MINB_F32 7   /* 56 registers: 7 CTAs per SM (r02W: 0.2294 ms against 0.2351 ms with 64 registers) */
#endif
#ifndef LBMDEM_K1_MINB_F64
#define LBMDEM_K1_MINB_F64 4   /* caps the fp64 build at 102 registers: 4 CTAs per SM (profiles/r01_k1_tuning.txt) */
#endif
template <typename real>
__global__ void __launch_bounds__(K1_THREADS(RowCfg<real>), (sizeof(real) == 8 ? LBMDEM_K1_MINB_F64 : LBMDEM_K1_MINB_F32 + RowCfg<real>::NB - 1) / RowCfg<real>::NB)
    lbm_rows_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmCp,
                                                                         const __grid_constant__ CUtensorMap tmCn,
                                                                         const __grid_constant__ FusedArgs<real> a) {
  using C = RowCfg<real>;
  constexpr int NCW = C::TY * C::NB / 32; /* consumer warps */
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[C::NS], empty[C::NS];

  const Lattice<real> &L = a.L;
  const int y0 = blockIdx.x * (C::TY * C::NB);
  /* rows of this CTA: a balanced share of the interior rows [R0, R1) of the strip */
  const int R0 = max(a.xlo, 1), R1 = min(a.xhi, L.lx - 1);
  const int r0 = R0 + (int)((long long)(R1 - R0) * blockIdx.y / gridDim.y);
  const int r1 = R0 + (int)((long long)(R1 - R0) * (blockIdx.y + 1) / gridDim.y);
  if (r1 <= r0) return;
  const int nload = r1 - r0 + 2; /* rows r0-1 .. r1; loaded row t is global row r0 - 1 + t */

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < C::NS; ++s) {
#if defined(LBMDEM_K1_LDGSTS)
      mbar_init(&full[s], 1 + C::TY); /* thread 0's expect_tx (the two map boxes) + one copy-completion arrival per thread */
#else
      mbar_init(&full[s], 1);
#endif
      mbar_init(&empty[s], NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  auto issue_row = [&](int t, int slot) { /* loaded row t is global row r0 - 1 + t */
    const int row = r0 - 1 + t - L.x0; /* local row */
    const uint32_t cp_bytes = a.prev16 ? C::CP_BYTES / 2 : C::CP_BYTES;
#if defined(LBMDEM_K1_PROBE) && LBMDEM_K1_PROBE == 3 /* 3 = writes alone: only the two map rows are loaded */
    mbar_expect_tx(&full[slot], (uint32_t)C::NB * (C::CN_BYTES + cp_bytes));
#else
    mbar_expect_tx(&full[slot], (uint32_t)C::NB * (C::A_BYTES + C::CN_BYTES + cp_bytes));
#endif
#pragma unroll
    for (int b = 0; b < C::NB; ++b) {
      unsigned char *base = smem + (size_t)slot * C::SLOT + (size_t)b * C::BLOCK;
      const int yb = y0 + b * C::TY;
#if !(defined(LBMDEM_K1_PROBE) && LBMDEM_K1_PROBE == 3)
#if defined(LBMDEM_K1_LD_EVICT_FIRST)
      tma_load_3d_ef(base, &tmA, &full[slot], yb - C::HY, row, 0);
#else
      tma_load_3d(base, &tmA, &full[slot], yb - C::HY, row, 0);
#endif
#endif
      tma_load_2d(base + C::A_PAD, &tmCn, &full[slot], yb - C::HC, row);
      tma_load_2d(base + C::A_PAD + C::CN_PAD, &tmCp, &full[slot], yb, row);
    }
  };
#if defined(LBMDEM_K1_LDGSTS)
  /* The population rows come in 16-byte pieces copied by the consumer threads themselves (LDGSTS) instead of one TMA
   * box: a piece whose nodes are all solid, not active and not wall ring under the STORED step's map is not loaded --
   * no fluid node neighbours it, so nothing pulls from it (solid nodes do not pull: the re-init sweep overwrites
   * them).  In a dense packing that is a quarter of the population reads, and the read stream is what bounds this
   * kernel (profiles/r02_k1_probes.txt).  The class bytes of a row's pieces are fetched one row ahead (kw). */
  static_assert(C::NB == 1, "LDGSTS rows: one block per CTA");
  constexpr int NPG = 16 / (int)sizeof(real);          /* nodes per piece */
  constexpr int GPR = C::BY / NPG;                     /* pieces per plane row */
  constexpr int NGR = NQ * GPR;                        /* pieces per row */
  constexpr int G = (NGR + C::TY - 1) / C::TY;         /* pieces per thread */
  static_assert(C::BY % NPG == 0, "row box is a whole number of 16-byte pieces");
  uint32_t kw[G];
  auto load_kw = [&](int t) { /* class bytes of this thread's pieces of loaded row t */
    const unsigned char *crow = a.cls_prev + (size_t)(r0 - 1 + t - L.x0) * L.pitch;
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int g = (int)threadIdx.x + k * C::TY, q = g / GPR, c = g - q * GPR;
      const int y = y0 - C::HY + c * NPG;
      kw[k] = 0;
      if (g < NGR && y >= 0 && y < L.pitch)
        kw[k] = sizeof(real) == 4 ? *reinterpret_cast<const uint32_t *>(crow + y) : *reinterpret_cast<const unsigned short *>(crow + y);
    }
  };
  auto issue_pieces = [&](int t, int slot) {
    if (threadIdx.x == 0) { /* the two map rows still travel as TMA boxes */
      const int row = r0 - 1 + t - L.x0;
      const uint32_t cp_bytes = a.prev16 ? C::CP_BYTES / 2 : C::CP_BYTES;
      unsigned char *base = smem + (size_t)slot * C::SLOT;
      mbar_expect_tx(&full[slot], (uint32_t)(C::CN_BYTES + cp_bytes));
      tma_load_2d(base + C::A_PAD, &tmCn, &full[slot], y0 - C::HC, row);
      tma_load_2d(base + C::A_PAD + C::CN_PAD, &tmCp, &full[slot], y0, row);
    }
    const real *rowp = a.A + (size_t)(r0 - 1 + t - L.x0) * L.pitch;
    real *dst = reinterpret_cast<real *>(smem + (size_t)slot * C::SLOT);
    constexpr uint32_t KEEP = sizeof(real) == 4 ? 0x0B0B0B0Bu : 0x0B0Bu; /* CLS_SOLID | CLS_ACT | CLS_RING per node */
    constexpr uint32_t DEAD = sizeof(real) == 4 ? 0x01010101u : 0x0101u; /* ... == CLS_SOLID for every node of the piece */
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int g = (int)threadIdx.x + k * C::TY, q = g / GPR, c = g - q * GPR;
      const int y = y0 - C::HY + c * NPG;
      if (g < NGR && y >= 0 && y < L.pitch && (a.stream_only || LBMDEM_K1_LDGSTS_ALL || (kw[k] & KEEP) != DEAD))
        cp_async16(dst + q * C::BY + c * NPG, rowp + q * L.plane + y);
    }
    cp_async_arrive(&full[slot]);
  };
  for (int t = 0; t < min(nload, C::NS); ++t) { /* fill the ring */
    load_kw(t);
    issue_pieces(t, t);
  }
  if (C::NS < nload) load_kw(C::NS);
#elif defined(LBMDEM_K1_NOPROD)
  if (threadIdx.x == 0) /* fill the ring */
    for (int t = 0; t < min(nload, C::NS); ++t) issue_row(t, t);
#else
  if (threadIdx.x >= C::TY * C::NB) {
    /* ---- producer warp: one lane keeps the ring full ---- */
    if (threadIdx.x == C::TY * C::NB) {
      int slot = 0;
      uint32_t round = 0; /* how many times the ring has wrapped */
      for (int t = 0; t < nload; ++t) {
        if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1); /* all consumer warps are done with row t - NS */
        issue_row(t, slot);
        if (++slot == C::NS) { slot = 0; ++round; }
      }
    }
    return;
  }
#endif

  /* ---- consumer warps ---- */
  const int b = threadIdx.x / C::TY;     /* the block of TY columns this warp works in (warp-uniform) */
  const int jy = threadIdx.x - b * C::TY;
  const int by = jy + C::HY;
  /* one 64-bit pointer + q * plane.  (A 32-bit offset from uniform plane bases needs fewer registers and gives more
   * CTAs per SM -- measured 4 % SLOWER, profiles/r01_k1_tuning.txt.) */
  real *out_row = a.out + node_index(L, r0, y0 + b * C::TY + jy);
  mbar_wait(&full[0], 0);
  mbar_wait(&full[1], 0);

#if defined(LBMDEM_K1_PROBE)
  real probe_acc = 0;
#endif
  int slot_m = 0, slot_0 = 1; /* slots of rows t-1 and t */
  uint32_t round_p = 0;       /* ring round of row t+1 */
  for (int t = 1; t <= nload - 2; ++t) {
    int slot_p = slot_0 + 1;
    if (slot_p == C::NS) { slot_p = 0; ++round_p; }
#if defined(LBMDEM_K1_LDGSTS)
    /* row t-2 was released by every warp at the end of iteration t-1 (or is about to be): its slot takes row t-2+NS */
    if (t >= 2 && t - 2 + C::NS < nload) {
      const int tl = t - 2 + C::NS, sl = (t - 2) % C::NS;
      mbar_wait(&empty[sl], ((t - 2) / C::NS) & 1);
      issue_pieces(tl, sl);
      if (tl + 1 < nload) load_kw(tl + 1);
    }
#elif defined(LBMDEM_K1_NOPROD)
    /* row t-2 was released by every warp at the end of iteration t-1 (or is about to be): its slot takes row
     * t-2+NS.  Loaded rows 0 .. NS-1 went in before the loop. */
    if (threadIdx.x == 0 && t >= 2 && t - 2 + C::NS < nload) {
      const int tl = t - 2 + C::NS, sl = (t - 2) % C::NS;
      mbar_wait(&empty[sl], ((t - 2) / C::NS) & 1);
      issue_row(tl, sl);
    }
#endif
    mbar_wait(&full[slot_p], round_p & 1);
    const int gx = r0 - 1 + t;
    { /* this thread's node */
    const int gy = y0 + b * C::TY + jy;
    const bool active = gy >= 1 && gy <= L.ly - 2;
    real *out = out_row;
    const size_t boff = (size_t)b * C::BLOCK;
    /* class bytes (lbm_node.cuh cell_class) of this step's map, row t with its y halo; behind them the stored step's map */
    const unsigned char *Kn0 = smem + (size_t)slot_0 * C::SLOT + boff + C::A_PAD;
    unsigned know = 0;
    int cprev = 0;
    bool work = active;
    if (active) {
      know = Kn0[jy + C::HC];
      if (a.prev16) {
        const unsigned short o = reinterpret_cast<const unsigned short *>(Kn0 + C::CN_PAD)[jy];
        cprev = o == OWN16_FLUID ? -1 : (int)o;
      } else {
        cprev = reinterpret_cast<const int *>(Kn0 + C::CN_PAD)[jy];
      }
    }
    {
      /* deep inside a grain under both maps: nothing reads what the re-init sweep would leave here (lbm_node.cuh,
       * node_is_dead; the fill_dead kernel materialises it when the populations are observed).  Skipped in whole
       * aligned groups of DEAD_GROUP lanes only, so that no 32-byte sector of the output is written in part. */
#if defined(LBMDEM_DEAD_GROUP) && LBMDEM_DEAD_GROUP == 0
      const bool dead = false;
#else
      const bool dead = cprev >= 0 && (know & CLS_SOLID) && !(know & (CLS_ACT | CLS_RIM)) && w_links_with_collide(L, gx, gy);
#endif
      bool skip = !active || (!a.stream_only && dead);
#if LBMDEM_DEAD_GROUP > 1
      const unsigned all = __ballot_sync(0xffffffffu, skip);
      const unsigned grp = (LBMDEM_DEAD_GROUP >= 32 ? 0xffffffffu : ((1u << (LBMDEM_DEAD_GROUP & 31)) - 1u))
                           << (threadIdx.x & 31 & ~(LBMDEM_DEAD_GROUP - 1));
      skip = (all & grp) == grp;
#endif
      if (skip) work = false;
    }
    if (work) {
      const real *Am = reinterpret_cast<const real *>(smem + (size_t)slot_m * C::SLOT + boff);
      const real *A0 = reinterpret_cast<const real *>(smem + (size_t)slot_0 * C::SLOT + boff);
      const real *Ap = reinterpret_cast<const real *>(smem + (size_t)slot_p * C::SLOT + boff);
      real f[NQ];
      /* a node that is solid under the stored step's map is overwritten by the re-init sweep,
       * whatever streams into it: skip the pull */
      if (a.stream_only || cell_is_fluid(cprev)) {
        f[0] = A0[by];
#pragma unroll
        for (int q = 1; q < NQ; ++q) {
          const int ex = ex_of(q), ey = ey_of(q);
          const real *As = ex > 0 ? Am : (ex < 0 ? Ap : A0); /* source row x - ex */
          f[q] = As[q * C::BY + by - ey];
        }
      }
#if defined(LBMDEM_K1_PROBE)
      /* TIMING PROBES of the row pipeline (results are wrong by construction; build with --tag, run bench.py only):
       * 1 = data movement alone (pull + store, no re-init / collide / w-links). */
      if (false) {
#else
      if (!a.stream_only) {
#endif
        /* sweeps 1-2 (lbm_node.cuh reinit_collide) */
        if (!cell_is_fluid(cprev)) equilibrium(L, a.grains_new[cell_obst(cprev)], gx, gy, f);
        if (know == 0) mrt_collide(L, f);
        if ((know & CLS_ACT) && w_links_with_collide(L, gx, gy)) {
          /* active solid node: links into non-fluid neighbours take the rest value (:1161-1162) */
          const unsigned char *Knm = smem + (size_t)slot_m * C::SLOT + boff + C::A_PAD;
          const unsigned char *Knp = smem + (size_t)slot_p * C::SLOT + boff + C::A_PAD;
#pragma unroll
          for (int q = 1; q < NQ; ++q) {
            const int ex = ex_of(q), ey = ey_of(q);
            const unsigned char *Ks = ex > 0 ? Knp : (ex < 0 ? Knm : Kn0); /* neighbour row x + ex */
            if (Ks[jy + C::HC + ey] != 0) f[q] = L.w[q];
          }
        }
      }
#if defined(LBMDEM_K1_PROBE) && LBMDEM_K1_PROBE == 2 /* 2 = reads alone: the row's values are folded into one register */
      {
        real acc = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc += f[q];
        probe_acc += acc;
      }
      if (false)
#endif
#pragma unroll
#if !defined(LBMDEM_K1_NO_STCS) /* streaming stores: the output is not read again before 600 MB of other traffic (measured: -2.5 %) */
      for (int q = 0; q < NQ; ++q) __stcs(&out[q * L.plane], f[q]);
#else
      for (int q = 0; q < NQ; ++q) out[q * L.plane] = f[q];
#endif
    }
    }
    out_row += L.pitch;
    /* this warp is done with row t-1 */
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[slot_m]);
    slot_m = slot_0;
    slot_0 = slot_p;
  }
#if defined(LBMDEM_K1_PROBE)
  if (probe_acc == (real)12345.678) a.out[0] = probe_acc; /* keeps the loads of probe 2 alive */
#endif
}

/* the w-links of an active solid node (see lbm_rows_kernel), map read from global memory */
template <typename real>
__device__ __forceinline__ void w_links_global(const Lattice<real> &L, const int *cell_now, int x, int y, real *f) {
  if (!cell_is_act(cell_now[node_index(L, x, y)]) || !w_links_with_collide(L, x, y)) return;
#pragma unroll
  for (int q = 1; q < NQ; ++q)
    if (!cell_is_fluid(cell_now[node_index(L, x + ex_of(q), y + ey_of(q))])) f[q] = L.w[q];
}

/* one thread per node from global memory: ring nodes (ring_only) or all owned nodes */
template <typename real>
__global__ void __launch_bounds__(128) lbm_plain_kernel(const __grid_constant__ FusedArgs<real> a, int ring_only,
                                                        int ring_rows_lo, int ring_rows_hi) {
  const Lattice<real> &L = a.L;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = a.xhi - a.xlo;
  int x, y;
  if (ring_only) {
    /* owned ring rows x = 0 / x = lx-1 in full, then the two ring columns of the other rows */
    const long long full = (long long)(ring_rows_lo + ring_rows_hi) * L.ly;
    if (t < full) {
      const int r = (int)(t / L.ly);
      y = (int)(t - (long long)r * L.ly);
      x = (r < ring_rows_lo) ? 0 : L.lx - 1;
    } else {
      const long long u = t - full;
      const int mid = rows - ring_rows_lo - ring_rows_hi;
      if (u >= 2ll * mid) return;
      x = a.xlo + ring_rows_lo + (int)(u >> 1);
      y = (u & 1) ? L.ly - 1 : 0;
    }
  } else {
    if (t >= (long long)rows * L.ly) return;
    const int r = (int)(t / L.ly);
    y = (int)(t - (long long)r * L.ly);
    x = a.xlo + r;
  }
  const size_t k = node_index(L, x, y);
  if (!a.stream_only && !is_ring(L, x, y) && node_is_dead(L, a.cell_prev[k], a.cell_new[k], x, y)) return;
  real f[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) f[q] = pull_plain(L, a.A, x, y, q);
  if (!a.stream_only && !is_ring(L, x, y)) {
    reinit_collide(L, a.grains_new, a.cell_prev[k], a.cell_new[k], x, y, f);
    w_links_global(L, a.cell_new, x, y, f);
  }
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
  if (cell_is_fluid(cprev) && !cell_is_act(cnow) && !cell_is_fluid(cnow)) return; /* nothing touches the node */
  real p[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) p[q] = f[q * L.plane + k];
  reinit_collide(L, grains_new, cprev, cnow, x, y, p);
  w_links_global(L, cell_now, x, y, p);
#pragma unroll
  for (int q = 0; q < NQ; ++q) f[q * L.plane + k] = p[q];
}

/* What the fused kernel left unwritten (lbm_node.cuh, node_is_dead), rows [xa, xb): the re-init equilibrium of
 * the node's previous owner, exactly as reinit_collide computes it. */
template <typename real>
__global__ void __launch_bounds__(128) lbm_fill_dead_kernel(const Lattice<real> L, real *A, const int *cell_prev,
                                                            const int *cell_now, const GrainRec<real> *grains_new, int xa,
                                                            int xb) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = xa + blockIdx.y;
  if (y >= L.ly || x >= xb || is_ring(L, x, y)) return;
  const size_t k = node_index(L, x, y);
  const int cprev = cell_prev[k], cnow = cell_now[k];
  if (!node_is_dead(L, cprev, cnow, x, y)) return;
  real p[NQ];
  reinit_collide(L, grains_new, cprev, cnow, x, y, p); /* both maps solid: the equilibrium alone */
#pragma unroll
  for (int q = 0; q < NQ; ++q) A[q * L.plane + k] = p[q];
}

template <typename real>
cudaError_t launch_lbm_rows(const CUtensorMap &tmA, const CUtensorMap &tmCp, const CUtensorMap &tmCn,
                            const FusedArgs<real> &a, cudaStream_t s) {
  using C = RowCfg<real>;
  static int resident = 0;
  if (!resident) {
    cudaError_t e = cudaFuncSetAttribute(lbm_rows_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + LBMDEM_K1_SMEM_PAD);
    if (e != cudaSuccess) return e;
    /* the whole 228 KB as shared memory: the kernel reads through TMA and barely uses L1, and with the driver's
     * default carve-out (164 KB, ncu: launch__occupancy_limit_shared_mem) a deeper ring costs resident CTAs */
    e = cudaFuncSetAttribute(lbm_rows_kernel<real>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbm_rows_kernel<real>, K1_THREADS(C), C::SMEM + LBMDEM_K1_SMEM_PAD)) != cudaSuccess)
      return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
#if defined(LBMDEM_K1_CTAS)
    per_sm = LBMDEM_K1_CTAS; /* tuning knob: trust the carve-out request instead of the occupancy calculator */
#endif
    resident = sms * per_sm;
  }
  const int R0 = a.xlo > 1 ? a.xlo : 1, R1 = a.xhi < a.L.lx - 1 ? a.xhi : a.L.lx - 1;
  if (R1 <= R0) return cudaSuccess;
  const int strips = (a.L.ly + C::TY * C::NB - 1) / (C::TY * C::NB);
  /* all CTAs co-resident (one wave), rows split evenly between the CTAs of a strip */
  int chunks = resident / strips;
  if (chunks < 1) chunks = 1;
  if (chunks > R1 - R0) chunks = R1 - R0;
  dim3 grid(strips, chunks);
  lbm_rows_kernel<real><<<grid, K1_THREADS(C), C::SMEM + LBMDEM_K1_SMEM_PAD, s>>>(tmA, tmCp, tmCn, a);
  return cudaGetLastError();
}

template <typename real>
cudaError_t launch_lbm_plain(const FusedArgs<real> &a, int ring_only, cudaStream_t s) {
  const int rows = a.xhi - a.xlo;
  if (rows <= 0) return cudaSuccess;
  int lo = 0, hi = 0;
  long long total;
  if (ring_only) {
    lo = (a.xlo == 0) ? 1 : 0;
    hi = (a.xhi == a.L.lx) ? 1 : 0;
    total = (long long)(lo + hi) * a.L.ly + 2ll * (rows - lo - hi);
  } else {
    total = (long long)rows * a.L.ly;
  }
  if (total <= 0) return cudaSuccess;
  lbm_plain_kernel<real><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(a, ring_only, lo, hi);
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

template <typename real>
cudaError_t launch_lbm_fill_dead(const Lattice<real> &L, real *A, const int *cell_prev, const int *cell_now,
                                 const GrainRec<real> *grains_new, int xa, int xb, cudaStream_t s) {
  if (xb <= xa) return cudaSuccess;
  dim3 grid((L.ly + 127) / 128, xb - xa);
  lbm_fill_dead_kernel<real><<<grid, 128, 0, s>>>(L, A, cell_prev, cell_now, grains_new, xa, xb);
  return cudaGetLastError();
}

#define INSTANTIATE_K1(real)                                                                                          \
  template cudaError_t launch_lbm_rows<real>(const CUtensorMap &, const CUtensorMap &, const CUtensorMap &,           \
                                             const FusedArgs<real> &, cudaStream_t);                                 \
  template cudaError_t launch_lbm_plain<real>(const FusedArgs<real> &, int, cudaStream_t);                            \
  template cudaError_t launch_lbm_h1<real>(const Lattice<real> &, real *, const int *, const int *,                   \
                                           const GrainRec<real> *, int, int, cudaStream_t);                           \
  template cudaError_t launch_lbm_fill_dead<real>(const Lattice<real> &, real *, const int *, const int *,            \
                                                  const GrainRec<real> *, int, int, cudaStream_t);
INSTANTIATE_K1(float)
INSTANTIATE_K1(double)

}  // namespace K1_NS
}  // namespace lbmdem
