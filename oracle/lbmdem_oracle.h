/*
 * oracle/lbmdem_oracle.h -- TEST INFRASTRUCTURE (checker), never linked into the product.
 *
 * Plain-C, serial, phase-by-phase CPU restatement of the coupled LBM-DEM step of
 * cb-geo/2d-lbm-dem (src/main.c), with the lattice size, scale and precision chosen at
 * run time (the reference fixes them with -D macros).  Compiled twice from the same
 * source: real = double (suffix _f64) and real = float (-DORACLE_SINGLE, suffix _f32).
 *
 * PARITY PIN: tests/test_oracle_vs_reference.py requires this restatement to be
 * BIT-IDENTICAL to the compiled reference (oracle/_ref, gcc -O2 -ffp-contract=off) on f,
 * obst, act, delta, fhf, the Verlet/wall lists and the grain kinematics over multi-step
 * runs, in both precisions; tests/golden/ holds vectors produced by the compiled reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#ifndef LBMDEM_ORACLE_H
#define LBMDEM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle oracle; /* opaque; one per precision-specific entry point family */

#define ORACLE_DECLARE(SFX)                                                                  \
  oracle *oracle_create_##SFX(int lx, int ly, double scale);                                 \
  void oracle_destroy_##SFX(oracle *o);                                                      \
  /* src/main.c:609-639 read_sample: returns nbgrains or <0 */                               \
  int oracle_read_sample_##SFX(oracle *o, const char *path);                                 \
  /* same, from arrays already in metres (values are cast to real) */                        \
  int oracle_set_grains_##SFX(oracle *o, int n, const double *r, const double *x1,           \
                              const double *x2);                                             \
  /* src/main.c:1834-1861: walls, dx, dtLB, npDEM, c, dt, rLB, init_density, init_obst */    \
  void oracle_init_##SFX(oracle *o);                                                         \
  /* non-reference extension: moving lid, the commented-out uw terms at src/main.c:1129-1130 */\
  void oracle_set_lid_##SFX(oracle *o, double uw);                                           \
  void oracle_set_vib_##SFX(oracle *o, int vib);                                             \
  /* dormant switches of the reference: wall-removal time (:117, :1555-1561), gravity tilt (:98, :1841) */\
  void oracle_set_dtt_##SFX(oracle *o, double dtt);                                          \
  void oracle_set_angleG_##SFX(oracle *o, double angleG);                                    \
  /* src/main.c:1697-1765 renderScene, n times (no file output) */                           \
  void oracle_step_##SFX(oracle *o, long n);                                                 \
  /* src/main.c:1711-1717 without the density print */                                       \
  void oracle_lbm_step_##SFX(oracle *o);                                                     \
  void oracle_reinit_obst_density_##SFX(oracle *o); /* :966-986   */                         \
  void oracle_obst_construction_##SFX(oracle *o);   /* :991-1065  */                         \
  void oracle_collision_streaming_##SFX(oracle *o); /* :1071-1243 */                         \
  void oracle_forces_fluid_##SFX(oracle *o);        /* :1285-1333 */                         \
  void oracle_init_verlet_##SFX(oracle *o);         /* :1519-1594 */                         \
  double oracle_total_density_##SFX(oracle *o);     /* :1249-1258 */                         \
  /* d: dx dtLB dt dt2 c Mgx Mdx Mby Mhy xG yG ; l: npDEM nbsteps nFile nbgrains */          \
  void oracle_get_scalars_##SFX(oracle *o, double *d, long *l);                              \
  void oracle_set_nbsteps_##SFX(oracle *o, long n);                                          \
  void oracle_get_f_##SFX(oracle *o, double *out); /* [lx][ly][9] */                         \
  void oracle_set_f_##SFX(oracle *o, const double *in);                                      \
  void oracle_get_delta_##SFX(oracle *o, double *out);                                       \
  void oracle_get_obst_##SFX(oracle *o, int *out);                                           \
  void oracle_set_obst_##SFX(oracle *o, const int *in);                                      \
  void oracle_get_act_##SFX(oracle *o, int *out);                                            \
  void oracle_get_grains_##SFX(oracle *o, double *out);       /* [N][13] as ref_get_grains */\
  void oracle_set_grain_state_##SFX(oracle *o, const double *in); /* [N][9] */               \
  void oracle_get_grain_diag_##SFX(oracle *o, double *out);   /* [N][17] */                  \
  void oracle_get_fhf_##SFX(oracle *o, double *out);          /* [N][3] */                   \
  void oracle_set_fhf_##SFX(oracle *o, const double *in);                                    \
  int oracle_get_verlet_##SFX(oracle *o, int *cumul, int *neigh, int cap);                   \
  void oracle_get_wall_lists_##SFX(oracle *o, int *counts, int *b, int *t, int *l, int *r);  \
  double oracle_time_coupled_##SFX(oracle *o, long n_dem_steps, long *n_lbm_steps);          \
  double oracle_time_lbm_##SFX(oracle *o, long n_lbm_steps);

ORACLE_DECLARE(f64)
ORACLE_DECLARE(f32)

#ifdef __cplusplus
}
#endif
#endif
