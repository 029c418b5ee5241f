/*
 * tools/integration/lbmdem_glue.h -- the binding INTEGRATION.md describes, as code that compiles INTO the reference.
 *
 * tools/integration/build.py patches a scratch copy of the reference's src/main.c in three places
 *   1. `#define duration 1.5` becomes overridable (-Dduration=...), so that a test can stop the run;
 *   2. this header is included just above renderScene() (src/main.c:1697), where every global it touches is
 *      already declared;
 *   3. renderScene() starts with `if (gpu_render()) return;` and main() calls gpu_setup(argv[1]) right after
 *      init_obst() (src/main.c:1861)
 * and builds it with -DLBMDEM_GPU against liblbmdem_gpu.so.  The reference's main(), its globals, its read_sample /
 * check_sample, its console lines and its OWN output writers (write_vtk, write_DEM, final_density) run unchanged; only
 * the body of renderScene() -- the coupled LBM + DEM step -- goes through the C ABI of include/lbmdem_gpu.h.
 * Everything here is plain C99 and only calls reference functions / assigns reference globals.
 */
#ifndef LBMDEM_GLUE_H
#define LBMDEM_GLUE_H
#ifdef LBMDEM_GPU
/* The reference fixes lx, ly, scale and rhoS as MACROS (src/main.c:24-32, :44), and the parameter block of the C ABI
 * has fields of the same names: the header is read, and the fields are assigned, with the macros out of the way. */
#pragma push_macro("lx")
#pragma push_macro("ly")
#pragma push_macro("scale")
#pragma push_macro("rhoS")
#undef lx
#undef ly
#undef scale
#undef rhoS
#include "lbmdem_gpu.h"
static void gpu_params(lbmdem_params *p, int lx_, int ly_, double scale_, double rhoS_, int single) {
  lbmdem_default_params(p); /* every constant of src/main.c:74-165 */
  p->lx = lx_; p->ly = ly_; p->scale = scale_; p->rhoS = rhoS_;
  p->single_precision = single;
}
#pragma pop_macro("rhoS")
#pragma pop_macro("scale")
#pragma pop_macro("ly")
#pragma pop_macro("lx")

static lbmdem_ctx *gpu;

static void gpu_die(const char *what) {
  fprintf(stderr, "%s: %s\n", what, lbmdem_last_error(gpu));
  exit(EXIT_FAILURE);
}

/* main(), after read_sample / check_sample and the set-up of :1834-1861 (kept: the writers read the host arrays) */
static void gpu_setup(const char *sample) {
  lbmdem_params p;
#ifdef SINGLE_PRECISION
  gpu_params(&p, lx, ly, scale, rhoS, 1); /* the -D macros of :24-32 */
#else
  gpu_params(&p, lx, ly, scale, rhoS, 0);
#endif
  if (getenv("LBMDEM_STRICT")) p.strict_fp = atoi(getenv("LBMDEM_STRICT"));
  if (lbmdem_create(&p, &gpu)) gpu_die("lbmdem_create");           /* fails loudly without an sm_100 GPU */
  if (lbmdem_load_sample(gpu, sample) < 0) gpu_die("lbmdem_load_sample");
  double d[11];
  long l[4];
  if (lbmdem_get_scalars(gpu, d, l)) gpu_die("lbmdem_get_scalars");
  /* the device derived the same constants from the same file: the reference's own values stay in force */
  if ((real)d[0] != dx || (real)d[2] != dt || (int)l[0] != npDEM || (int)l[3] != nbgrains) {
    fprintf(stderr, "device set-up differs from main()'s: dx %g/%g dt %g/%g npDEM %ld/%d\n", d[0], (double)dx, d[2], (double)dt,
            l[0], npDEM);
    exit(EXIT_FAILURE);
  }
}

/* what write_vtk / write_DEM / final_density read: f, obst, g[], fhf1..3 (src/main.c:237-438, :1263-1273) */
static void gpu_pull_state(void) {
  static double *fbuf, *tab, *fh;
  const size_t nodes = (size_t)lx * ly;
  if (!fbuf) {
    fbuf = malloc(sizeof(double) * nodes * Q);
    tab = malloc(sizeof(double) * 13 * nbgrains);
    fh = malloc(sizeof(double) * 3 * nbgrains);
    if (!fbuf || !tab || !fh) gpu_die("malloc");
  }
  if (lbmdem_get_f(gpu, fbuf) || lbmdem_get_obst(gpu, &obst[0][0]) || lbmdem_get_grains(gpu, tab) || lbmdem_get_fhf(gpu, fh))
    gpu_die("lbmdem_get_*");
  real *fp = &f[0][0][0];
  for (size_t k = 0; k < nodes * Q; ++k) fp[k] = (real)fbuf[k];
  for (int i = 0; i < nbgrains; ++i) {
    const double *o = tab + (size_t)13 * i;
    g[i].x1 = o[0]; g[i].x2 = o[1]; g[i].x3 = o[2]; g[i].v1 = o[3]; g[i].v2 = o[4]; g[i].v3 = o[5];
    g[i].a1 = o[6]; g[i].a2 = o[7]; g[i].a3 = o[8];
    fhf1[i] = fh[3 * (size_t)i]; fhf2[i] = fh[3 * (size_t)i + 1]; fhf3[i] = fh[3 * (size_t)i + 2];
  }
}

/* renderScene(): src/main.c:1708-1764 on the device, then the reference's own cadence and writers (:1764-1776) */
static int gpu_render(void) {
  if (lbmdem_step(gpu, 1)) gpu_die("lbmdem_step");
  nbsteps++;
  if (nbsteps % stepConsole == 0) { /* check_density, :1715 / :1259 */
    double sum;
    if (lbmdem_total_density(gpu, &sum)) gpu_die("lbmdem_total_density");
    printf("Iteration Number %ld, Total density in the system %f\n", nbsteps, sum);
  }
  const int film = nbsteps % stepFilm == 0, strob = nbsteps % stepStrob == 0;
  const int last = !((nbsteps * dt) <= duration); /* main()'s loop ends after this call: final_density() reads f */
  if (film || strob || last) gpu_pull_state();
  if (film) {
    write_vtk(lx, ly, f, nbgrains, g);
    nFile++;
  }
  if (strob) write_DEM();
  return 1;
}
#endif /* LBMDEM_GPU */
#endif
