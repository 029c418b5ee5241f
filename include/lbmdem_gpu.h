/*
 * lbmdem_gpu.h -- C ABI of liblbmdem_gpu.so, the B200 (sm_100a) implementation of the coupled
 * LBM-DEM hot path of cb-geo/2d-lbm-dem.
 *
 * The reference has no library boundary: the path sits behind `void f(void)` functions over
 * file-scope globals, called from renderScene() (src/main.c:1697-1777) and set up by main()
 * (src/main.c:1783-1879).  Each entry point below names the reference code it replaces; a
 * maintainer of the reference would call them from main()/renderScene() as shown in
 * INTEGRATION.md.  Plain pointers and sizes only; every function returns 0 on success and a
 * negative LBMDEM_E* code on failure (the message is available from lbmdem_last_error()).
 *
 * Conventions: one driving host thread per context; all calls are synchronous (they return
 * after the device work has completed) unless stated otherwise; host arrays are owned by the
 * caller and copied; device memory is owned by the context.  Array layouts are the
 * reference's: f[x][y][q] with q fastest (src/main.c:56), obst[x][y] (src/main.c:83),
 * grains as rows of doubles.  In a strip-decomposed run (nranks > 1) lattice arrays cover the
 * rows [xlo, xhi) this rank owns (lbmdem_get_strip).
 */
#ifndef LBMDEM_GPU_H
#define LBMDEM_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbmdem_ctx lbmdem_ctx;

enum {
  LBMDEM_OK = 0,
  LBMDEM_EINVAL = -1,   /* bad argument */
  LBMDEM_ECUDA = -2,    /* CUDA runtime / driver error */
  LBMDEM_ENOMEM = -3,
  LBMDEM_ESTATE = -4,   /* call out of order (e.g. step before set_grains) */
  LBMDEM_EIO = -5,      /* sample file could not be read */
  LBMDEM_ECAP = -6,     /* neighbour-list capacity exceeded (the reference only prints, :1535) */
  LBMDEM_ENCCL = -7,
  LBMDEM_ERANGE = -8    /* a hydrodynamic-force sum left the range of the fixed-point accumulators (diverged run) */
};

/* What the reference fixes at compile time (-Dlx -Dly -Dscale -DSINGLE_PRECISION,
 * src/main.c:24-40) or as initialised globals (src/main.c:74-165). */
typedef struct lbmdem_params {
  int lx, ly;              /* lattice size                            src/main.c:24-29  */
  double scale;            /* lattice refinement                      src/main.c:30-32  */
  int single_precision;    /* 1 = -DSINGLE_PRECISION build            src/main.c:34-40  */
  int device;              /* CUDA device ordinal                                          */
  int rank, nranks;        /* x-strip decomposition, one context per GPU; 0,1 = whole lattice */
  /* LBM constants, src/main.c:74-94 */
  double tau, nu, rho_moy, reductionR;
  double s2, s3, s5, s7, s8, s9;
  /* DEM constants, src/main.c:97-118, :162-164 */
  double G, angleG, kg, kt, km, ktm, nug, num, numb, nugt, mu, mum, mumb, murf;
  double rscale, distVerlet, dtt, iterDEM, freq, amp, rhoS;
  long UpdateVerlet;       /* 100  */
  long stepFilm;           /* 8000: steps on which the alternate contact law runs (:1342) */
  /* extensions (0 = reference behaviour) */
  double lid_u;            /* moving lid, the commented-out uw terms at src/main.c:1129-1130 */
  int strict_fp;           /* 1: LBM kernel built without multiply-add contraction and forces summed in
                              the reference's serial order -> bit-identical to the reference build */
  int kernel;              /* cross-check switches, 0 = default.  bit 0: plain one-thread-per-node LBM kernel instead of
                              the TMA row pipeline; bit 1: the rasteriser rebuilds every lattice tile every step instead
                              of the tiles in which a covered node changed; bit 2: three launches per DEM sub-step
                              instead of one launch for all the sub-steps between two LBM steps */
  int neighbour_capacity;  /* per-grain Verlet capacity, default 32 */
  int vib;                 /* 1: shake the left/right walls, src/main.c:162, :1701-1706 (default 0) */
} lbmdem_params;

/* Fills *p with the reference's defaults: lx=7826, ly=2325, scale=1, fp64 and every constant
 * of src/main.c:74-165. */
int lbmdem_default_params(lbmdem_params *p);
/* sizeof(lbmdem_params) as compiled into the library, for FFI bindings to check their struct */
int lbmdem_sizeof_params(void);

/* main():1802-1832 (allocations).  Fails with LBMDEM_ECUDA when no sm_100 device is usable:
 * there is no CPU path. */
int lbmdem_create(const lbmdem_params *p, lbmdem_ctx **out);
void lbmdem_destroy(lbmdem_ctx *ctx);
const char *lbmdem_last_error(const lbmdem_ctx *ctx); /* ctx may be NULL: error of the last failed create */

/* read_sample (src/main.c:609-639) followed by the rest of main()'s set-up (:1834-1861):
 * walls, gravity, dx, dtLB, npDEM, c, dt, rLB, init_density, init_obst.  Returns nbgrains. */
int lbmdem_load_sample(lbmdem_ctx *ctx, const char *path);
/* The same from arrays already in metres (values are rounded to the build's `real`). */
int lbmdem_set_grains(lbmdem_ctx *ctx, int n, const double *r, const double *x1, const double *x2);

/* renderScene() n times (src/main.c:1697-1765): LBM step every npDEM-th call, Verlet lists
 * every UpdateVerlet-th, kick-drift, forces, kick.  No file output. */
int lbmdem_step(lbmdem_ctx *ctx, long n_dem_steps);
/* ONE renderScene() call that also returns the grain positions and velocities as
 * acceleration_grains() saw them (after the kick-drift of src/main.c:1748-1753, before the
 * forces): mid[n][6] = x1 x2 x3 v1 v2 v3.  The host replays the reference's contact diagnostics
 * (p, s, slip, ... of write_DEM, src/main.c:340-438) from it on output steps. */
int lbmdem_step_capture(lbmdem_ctx *ctx, double *mid);
/* The LBM part of one renderScene() call (src/main.c:1711-1717): reinit_obst_density,
 * obst_construction, collision_streaming, forces_fluid. */
int lbmdem_lbm_step(lbmdem_ctx *ctx);
/* initVerlet + VerletWall (src/main.c:1519-1594) */
int lbmdem_build_verlet(lbmdem_ctx *ctx);
/* n LBM steps back to back without returning to the host in between, then one synchronise
 * (throughput measurements; same result as n calls of lbmdem_lbm_step). */
int lbmdem_lbm_steps(lbmdem_ctx *ctx, long n);

/* d[11]: dx dtLB dt dt2 c Mgx Mdx Mby Mhy xG yG ; l[4]: npDEM nbsteps nFile nbgrains */
int lbmdem_get_scalars(lbmdem_ctx *ctx, double *d, long *l);
int lbmdem_set_nbsteps(lbmdem_ctx *ctx, long nbsteps);
int lbmdem_get_strip(lbmdem_ctx *ctx, int *xlo, int *xhi);

/* check_density / final_density (src/main.c:1249-1273): sum of all populations of the owned rows */
int lbmdem_total_density(lbmdem_ctx *ctx, double *sum);

/* raw state, reference layouts, owned rows [xlo, xhi) */
int lbmdem_get_f(lbmdem_ctx *ctx, double *out /* [xhi-xlo][ly][9] */);
int lbmdem_set_f(lbmdem_ctx *ctx, const double *in);
int lbmdem_get_obst(lbmdem_ctx *ctx, int *out /* [xhi-xlo][ly] */);
int lbmdem_set_obst(lbmdem_ctx *ctx, const int *in); /* the map the NEXT LBM step treats as "old" */
int lbmdem_get_act(lbmdem_ctx *ctx, int *out);
int lbmdem_get_grains(lbmdem_ctx *ctx, double *out /* [n][13]: x1 x2 x3 v1 v2 v3 a1 a2 a3 r m It rLB */);
int lbmdem_set_grain_state(lbmdem_ctx *ctx, const double *in /* [n][9]: x1 x2 x3 v1 v2 v3 a1 a2 a3 */);
int lbmdem_get_fhf(lbmdem_ctx *ctx, double *out /* [n][3] */);
int lbmdem_set_fhf(lbmdem_ctx *ctx, const double *in);
/* full neighbour lists as built on the device: count[n], nbr[n][capacity] ascending, wall flags
 * (bit 0 B, 1 T, 2 L, 3 R).  The reference's half list is the j > i part. */
int lbmdem_get_verlet(lbmdem_ctx *ctx, int *count, int *nbr, int capacity, int *wall_flags);

/* Checkpoint / restart (the reference has none): everything the next renderScene()
 * call reads, in one file per rank ("<path>.rank<k>" when nranks > 1).  lbmdem_load_state replaces lbmdem_load_sample on a context
 * created with the same lattice, precision and decomposition; returns nbgrains.  A run continued
 * from the file is bit-identical to the uninterrupted one. */
int lbmdem_save_state(lbmdem_ctx *ctx, const char *path);
int lbmdem_load_state(lbmdem_ctx *ctx, const char *path);

/* write_vtk's five point fields (src/main.c:284-323) for the owned rows, float32, [y][x-xlo]
 * order (x fastest), vectors with 3 components; grain_p may be NULL on input side (pressure of
 * grains is a contact diagnostic kept by the caller): pass per-grain values or NULL for zeros. */
int lbmdem_get_fields(lbmdem_ctx *ctx, const double *grain_p_in, float *grain_pressure, float *grain_velocity,
                      float *grain_acceleration, float *fluid_pressure, float *fluid_velocity);

/* End-to-end form of one coupled step with HOST buffers: upload the grain kinematic state,
 * run n_dem_steps renderScene() calls, download the new state, fhf and the density checksum.
 * Any of the output pointers may be NULL.  Buffers in page-locked memory the device can address
 * (lbmdem_host_alloc, or anything cudaHostRegister'ed as mapped) cross PCIe without a copy operation:
 * the kernels that convert between grain rows and the device's columns read state_in and -- when fhf_out
 * directly follows state_out, i.e. fhf_out == state_out + 9 n -- write the outputs in place; pageable
 * buffers go through one staging copy each way. */
int lbmdem_step_host(lbmdem_ctx *ctx, const double *state_in /* [n][9] or NULL */, long n_dem_steps,
                     double *state_out /* [n][9] */, double *fhf_out /* [n][3] */, double *density_out);
/* the same with grain rows of float: what a -DSINGLE_PRECISION build of the reference holds (`real`,
 * src/main.c:24-40, :182-198); single-precision contexts only (LBMDEM_EINVAL otherwise) */
int lbmdem_step_host_f32(lbmdem_ctx *ctx, const float *state_in /* [n][9] or NULL */, long n_dem_steps,
                         float *state_out /* [n][9] */, float *fhf_out /* [n][3] */, double *density_out);

/* Strip-decomposed runs replicate the grains on every GPU, but only one copy has to cross the host boundary: in the
 * `share` form every rank uploads and downloads the rows of ITS share of the grains -- the contiguous index range
 * [i0, i1) of lbmdem_get_share, a balanced split over the ranks -- and the ranks pass the uploaded rows on to each other
 * over NVLink (one ncclAllGather when the shares are equal, else one ncclBroadcast per rank) before the step.  state_in / state_out: [i1-i0][9], fhf_out: [i1-i0][3].
 * With one rank it is lbmdem_step_host. */
int lbmdem_get_share(lbmdem_ctx *ctx, int *i0, int *i1);
int lbmdem_step_host_share(lbmdem_ctx *ctx, const double *state_in, long n_dem_steps, double *state_out, double *fhf_out,
                           double *density_out);
int lbmdem_step_host_share_f32(lbmdem_ctx *ctx, const float *state_in, long n_dem_steps, float *state_out, float *fhf_out,
                               double *density_out);

/* page-locked host memory for the buffers of lbmdem_step_host (the reference keeps its grain
 * array in plain malloc memory, src/main.c:612; a caller that wants the copies without staging
 * allocates it here instead) */
int lbmdem_host_alloc(size_t bytes, void **ptr);
int lbmdem_host_free(void *ptr);

/* ---- multi-GPU (one context per GPU / process; lattice split into x strips, grains replicated) ---- */
/* 128-byte NCCL unique id, produced on one rank and distributed by the caller */
int lbmdem_nccl_unique_id(void *id128);
/* joins the communicator (collective over all ranks of the run).  NCCL moves the ghost rows; at the first step the
 * ranks also map each other's force-sum buffers (CUDA IPC) and add the sums with a kernel that reads them over NVLink
 * -- same bits, a third of the latency of ncclAllReduce on 8 GPUs; LBMDEM_PEER_SUMS=0 in the environment, the strict
 * build, or a peer that cannot be mapped keep ncclAllReduce. */
int lbmdem_attach_nccl(lbmdem_ctx *ctx, const void *id128);

/* ---- in-process strip group (no collective library) ----
 * The ranks of a decomposed run as contexts of ONE process -- one per GPU, or several on the same GPU (the strip logic
 * exercised on a one-GPU machine) -- each driven by its own host thread.  A rank pulls its ghost rows from the
 * neighbour's device memory (peer copies over NVLink across devices) and adds the hydrodynamic-force sums with one
 * kernel that reads every peer's partial sums; the threads meet at a barrier inside the step, so every rank of the
 * group must make the same stepping calls concurrently.  Create the contexts with rank / nranks set, attach each to
 * the group, then drive them from nranks threads.  A rank that fails releases its peers (they return LBMDEM_ESTATE). */
typedef struct lbmdem_local_group lbmdem_local_group;
int lbmdem_local_group_create(int nranks, lbmdem_local_group **group);
int lbmdem_attach_local(lbmdem_ctx *ctx, lbmdem_local_group *group);
void lbmdem_local_group_destroy(lbmdem_local_group *group);   /* after the contexts are destroyed */

/* Exact fingerprint of the lattice state over the owned rows: sums[0] = sum mod 2^64 of the bit patterns of the
 * reference's f[x][y][q] (as double), each multiplied by 2 k + 1 with k its GLOBAL flat index (src/main.c:56);
 * sums[1] the same over obst[x][y] + 2 (src/main.c:83).  Integer adds commute: the fingerprints of the strips of a
 * decomposed run add up (mod 2^64) to the one-GPU value iff every population and node index is identical. */
int lbmdem_state_checksum(lbmdem_ctx *ctx, unsigned long long sums[2]);

/* ---- instrumentation ---- */
/* CUDA-event time (ms) and launch count of the fused LBM kernel accumulated since the last reset */
int lbmdem_get_kernel_timer(lbmdem_ctx *ctx, double *k1_ms, long *k1_launches, long *all_launches);
int lbmdem_reset_kernel_timer(lbmdem_ctx *ctx, int enable_events);
/* sizes of the sparse work lists of the LAST LBM step: counts[0] bounce-back links (active solid node, link into a fluid
 * neighbour; src/main.c:1163-1185), counts[1] boundary nodes with a non-fluid foreign neighbour (forces_fluid, :1313),
 * counts[2] links evaluated through the deferred list (the order-dependent one-node-gap case of :1176-1185),
 * counts[3] lattice tiles (32 x 64 nodes) whose part of the obstacle map and lists the last rasteriser run rebuilt: the
 * others were carried over unchanged from the previous step (no covered node changed in them) */
int lbmdem_get_list_counts(lbmdem_ctx *ctx, long counts[4]);
/* the CUDA stream all work of this context is issued on (a cudaStream_t) */
void *lbmdem_stream(lbmdem_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
