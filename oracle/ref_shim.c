/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, never part of the product path.
 *
 * Wraps the UNMODIFIED reference translation unit (src/main.c of
 * cb-geo/2d-lbm-dem, compiled in place from /root/reference) into a shared
 * library with a small C interface so that tests and the CPU-baseline leg of
 * bench.py can drive the reference's own functions step by step and read its
 * state.  No reference source is copied: main.c is pulled in with #include
 * (path given on the command line as -DLBMDEM_REF_MAIN='"/root/reference/src/main.c"'),
 * its main() is renamed out of the way, and everything below only *calls*
 * reference functions / reads reference globals.
 *
 * The reference fixes lx, ly, scale and the precision at compile time
 * (src/main.c:24-40), so one library is built per configuration:
 *   oracle/_ref/libref_<lx>x<ly>_s<scale>_<f64|f32>[_omp].so      (see oracle/build.py)
 *
 * What ref_init() does is the start-up sequence of main() (src/main.c:1798-1861)
 * expressed as calls to the reference's own functions, with two deliberate
 * differences, both documented in SURVEY.md App. B:
 *   - g[i].mw, which the reference leaves uninitialised on the read_sample
 *     path (src/main.c:617-636), is set to 0 (what fresh malloc pages give);
 *   - the endless do/while over `duration` (src/main.c:1880-1890) is replaced
 *     by ref_step(n) = n calls of renderScene().
 */
#define _GNU_SOURCE
#define main lbmdem_reference_main__
#include LBMDEM_REF_MAIN
#undef main

#include <time.h>

#define REF_API __attribute__((visibility("default")))

static int ref_initialised = 0;

REF_API void ref_config(int *lx_, int *ly_, double *scale_, int *real_bytes) {
  *lx_ = lx;
  *ly_ = ly;
  *scale_ = (double)(scale);
  *real_bytes = (int)sizeof(real);
}

static void ref_free_all(void) {
  if (!ref_initialised) return;
  free(g); free(f); free(obst); free(act); free(delta); free(rLB); free(cumul);
  free(neighbours); free(neighbourWallB); free(neighbourWallR);
  free(neighbourWallL); free(neighbourWallT); free(fhf); free(fhf1); free(fhf2);
  free(fhf3);
  ref_initialised = 0;
}

/* main():1798-1861 and :1879, as calls into the reference. Returns nbgrains. */
REF_API int ref_init(const char *sample_path) {
  ref_free_all();
  /* globals that main() relies on being in their load-time state */
  nbsteps = 0; nFile = 0; start = 0; t = 0; vib = 0; dtt = 0.; angleG = 0.0;
  pf = 0.; pft = 0.; pff = 0.; ic = 0;
  TSE = 0.0; TBW = 0.0; INCE = 0.0; TSLIP = 0.0; TRW = 0.0;
  nNeighWallb = nNeighWallt = nNeighWallL = nNeighWallR = 0;

  c_squ = 1. / 3.;
  g = read_sample((char *)sample_path);
  for (int i = 0; i < nbgrains; ++i) g[i].mw = 0; /* App. B #2 */
  check_sample(nbgrains, g);

  f = malloc(sizeof(real) * lx * ly * Q);
  obst = malloc(sizeof(int) * lx * ly);
  act = malloc(sizeof(int) * lx * ly);
  delta = malloc(sizeof(real) * lx * ly * Q);
  rLB = malloc(sizeof(real) * nbgrains);
  cumul = calloc(nbgrains, sizeof(int));
  neighbours = calloc((size_t)nbgrains * 6, sizeof(int));
  neighbourWallB = calloc(nbgrains, sizeof(int));
  neighbourWallR = calloc(nbgrains, sizeof(int));
  neighbourWallL = calloc(nbgrains, sizeof(int));
  neighbourWallT = calloc(nbgrains, sizeof(int));
  fhf = malloc(sizeof(struct force) * nbgrains);
  fhf1 = calloc(nbgrains, sizeof(real));
  fhf2 = calloc(nbgrains, sizeof(real));
  fhf3 = calloc(nbgrains, sizeof(real));
  if (!f || !obst || !act || !delta) return -1;
  /* the reference never reads delta[..][0] / act on fresh pages before writing
     them, but give them a defined value so that dumps are reproducible */
  memset(delta, 0, sizeof(real) * lx * ly * Q);
  memset(act, 0, sizeof(int) * lx * ly);

  init_density(lx, ly, f);

  Mgx = 0.;
  Mdx = 1.e-3 * lx / 10;
  Mhy = 1.e-3 * ly / 10;
  Mby = 0.;
  xG = -G * sin(angleG);
  yG = -G * cos(angleG);
  dx = (1. / scale) * (Mdx - Mgx) / (lx - 1);

  real rMin = minimum_grain_radius(nbgrains, g);
  real dtmax = (1 / iterDEM) * pi * rMin * sqrt(pi * rhoS / kg);
  dtLB = dx * dx * (tau - 0.5) / (3 * nu);
  npDEM = (dtLB / dtmax + 1);
  c = dx / dtLB;
  dt = dtLB / npDEM;
  dt2 = dt * dt;
  for (int i = 0; i <= nbgrains - 1; i++) rLB[i] = reductionR * g[i].r / dx;
  init_obst();

  /* stats.data header is written by main(); write_DEM appends to it */
  s_stats = fopen("stats.data", "w");
  if (s_stats) { fprintf(s_stats, "#ref_shim\n"); fclose(s_stats); }

  start = 1;
  ref_initialised = 1;
  return nbgrains;
}

REF_API void ref_step(long n) { for (long k = 0; k < n; ++k) renderScene(); }

/* individual LBM phases, for single-phase parity tests (src/main.c:1711-1717) */
REF_API void ref_reinit_obst_density(void) { reinit_obst_density(); }
REF_API void ref_obst_construction(void) { obst_construction(); }
REF_API void ref_collision_streaming(void) { collision_streaming(); }
REF_API void ref_forces_fluid(void) { forces_fluid(lx, ly, f, nbgrains, g); }
REF_API void ref_init_verlet(void) { initVerlet(); VerletWall(); }
REF_API void ref_lbm_step(void) {
  reinit_obst_density(); obst_construction(); collision_streaming();
  forces_fluid(lx, ly, f, nbgrains, g);
}

/* scalars: dx dtLB dt dt2 c Mgx Mdx Mby Mhy xG yG  | npDEM nbsteps nFile nbgrains */
REF_API void ref_get_scalars(double *d, long *l) {
  d[0] = dx; d[1] = dtLB; d[2] = dt; d[3] = dt2; d[4] = c; d[5] = Mgx; d[6] = Mdx;
  d[7] = Mby; d[8] = Mhy; d[9] = xG; d[10] = yG;
  l[0] = npDEM; l[1] = nbsteps; l[2] = nFile; l[3] = nbgrains;
}
REF_API void ref_set_nbsteps(long n) { nbsteps = n; }
REF_API void ref_set_vib(int v) { vib = v; } /* src/main.c:162 */
/* dormant switches: the time at which VerletWall() lets the confining right / top walls go
 * (src/main.c:117, :1555-1561) and the tilt of gravity (src/main.c:98; main() derives xG, yG from it once, :1841-1842) */
REF_API void ref_set_dtt(double v) { dtt = (real)v; }
REF_API void ref_set_angleG(double v) {
  angleG = (real)v;
  xG = -G * sin(angleG);
  yG = -G * cos(angleG);
}

/* serial sum exactly as check_density (src/main.c:1249-1258), value returned */
REF_API double ref_total_density(void) {
  real sum = 0;
  for (int x = 0; x < lx; x++)
    for (int y = 0; y < ly; y++)
      for (int q = 0; q < Q; q++) sum = sum + f[x][y][q];
  return (double)sum;
}

/* lattice state, reference layout [x][y][q] / [x][y], widened to double */
REF_API void ref_get_f(double *out) {
  const real *p = &f[0][0][0];
  for (size_t k = 0; k < (size_t)lx * ly * Q; ++k) out[k] = p[k];
}
REF_API void ref_set_f(const double *in) {
  real *p = &f[0][0][0];
  for (size_t k = 0; k < (size_t)lx * ly * Q; ++k) p[k] = (real)in[k];
}
REF_API void ref_get_delta(double *out) {
  const real *p = &delta[0][0][0];
  for (size_t k = 0; k < (size_t)lx * ly * Q; ++k) out[k] = p[k];
}
REF_API void ref_get_obst(int *out) { memcpy(out, obst, sizeof(int) * lx * ly); }
REF_API void ref_set_obst(const int *in) { memcpy(obst, in, sizeof(int) * lx * ly); }
REF_API void ref_get_act(int *out) { memcpy(out, act, sizeof(int) * lx * ly); }

/* grains: 13 columns  x1 x2 x3 v1 v2 v3 a1 a2 a3 r m It rLB */
REF_API void ref_get_grains(double *out) {
  for (int i = 0; i < nbgrains; ++i) {
    double *o = out + (size_t)i * 13;
    o[0] = g[i].x1; o[1] = g[i].x2; o[2] = g[i].x3;
    o[3] = g[i].v1; o[4] = g[i].v2; o[5] = g[i].v3;
    o[6] = g[i].a1; o[7] = g[i].a2; o[8] = g[i].a3;
    o[9] = g[i].r; o[10] = g[i].m; o[11] = g[i].It; o[12] = rLB[i];
  }
}
/* kinematic state only: x1 x2 x3 v1 v2 v3 a1 a2 a3 */
REF_API void ref_set_grain_state(const double *in) {
  for (int i = 0; i < nbgrains; ++i) {
    const double *o = in + (size_t)i * 9;
    g[i].x1 = (real)o[0]; g[i].x2 = (real)o[1]; g[i].x3 = (real)o[2];
    g[i].v1 = (real)o[3]; g[i].v2 = (real)o[4]; g[i].v3 = (real)o[5];
    g[i].a1 = (real)o[6]; g[i].a2 = (real)o[7]; g[i].a3 = (real)o[8];
  }
}
/* per-step diagnostics: p s f1 f2 ifm fm fr ifr M11 M12 M21 M22 ice slip rw z zz  (17 columns) */
REF_API void ref_get_grain_diag(double *out) {
  for (int i = 0; i < nbgrains; ++i) {
    double *o = out + (size_t)i * 17;
    o[0] = g[i].p; o[1] = g[i].s; o[2] = g[i].f1; o[3] = g[i].f2; o[4] = g[i].ifm;
    o[5] = g[i].fm; o[6] = g[i].fr; o[7] = g[i].ifr; o[8] = g[i].M11; o[9] = g[i].M12;
    o[10] = g[i].M21; o[11] = g[i].M22; o[12] = g[i].ice; o[13] = g[i].slip;
    o[14] = g[i].rw; o[15] = g[i].z; o[16] = g[i].zz;
  }
}
REF_API void ref_get_fhf(double *out) {
  for (int i = 0; i < nbgrains; ++i) {
    out[3 * (size_t)i + 0] = fhf1[i]; out[3 * (size_t)i + 1] = fhf2[i];
    out[3 * (size_t)i + 2] = fhf3[i];
  }
}
REF_API void ref_set_fhf(const double *in) {
  for (int i = 0; i < nbgrains; ++i) {
    fhf1[i] = (real)in[3 * (size_t)i + 0]; fhf2[i] = (real)in[3 * (size_t)i + 1];
    fhf3[i] = (real)in[3 * (size_t)i + 2];
  }
}
/* Verlet half list: cumul[N] (end offsets, src/main.c:1539) and neighbours[cumul[N-2]] */
REF_API int ref_get_verlet(int *cumul_out, int *neigh_out, int neigh_cap) {
  int total = (nbgrains >= 2) ? cumul[nbgrains - 2] : 0;
  memcpy(cumul_out, cumul, sizeof(int) * nbgrains);
  if (total > neigh_cap) return -total;
  memcpy(neigh_out, neighbours, sizeof(int) * total);
  return total;
}
/* wall lists in the order B T L R; counts[4]; lists each sized nbgrains */
REF_API void ref_get_wall_lists(int *counts, int *b, int *tt, int *l, int *rr) {
  counts[0] = nNeighWallb; counts[1] = nNeighWallt; counts[2] = nNeighWallL;
  counts[3] = nNeighWallR;
  memcpy(b, neighbourWallB, sizeof(int) * nNeighWallb);
  memcpy(tt, neighbourWallT, sizeof(int) * nNeighWallt);
  memcpy(l, neighbourWallL, sizeof(int) * nNeighWallL);
  memcpy(rr, neighbourWallR, sizeof(int) * nNeighWallR);
}

static double ref_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
/* CPU-baseline timers (bench.py --impl reference / cpu_baseline).
 * ref_time_coupled: n DEM steps through renderScene(), returns seconds and the
 * number of LBM steps they contained.  ref_time_lbm: n LBM steps, LBM phases only. */
REF_API double ref_time_coupled(long n_dem_steps, long *n_lbm_steps) {
  long lbm = 0;
  double t0 = ref_now();
  for (long k = 0; k < n_dem_steps; ++k) {
    if (nbsteps % npDEM == 0) ++lbm;
    renderScene();
  }
  double t1 = ref_now();
  *n_lbm_steps = lbm;
  return t1 - t0;
}
REF_API double ref_time_lbm(long n_lbm_steps) {
  double t0 = ref_now();
  for (long k = 0; k < n_lbm_steps; ++k) ref_lbm_step();
  return ref_now() - t0;
}
/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the timed baseline asks for its threads explicitly */
REF_API void ref_set_omp_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
REF_API int ref_omp_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
