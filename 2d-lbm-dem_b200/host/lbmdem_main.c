/*
 * lbmdem_main.c -- the drop-in executable: `lbmdem <inputfile>` as in the reference
 * (src/main.c:1783-1898), with the coupled step running on the GPU through liblbmdem_gpu.so.
 *
 * What the reference fixes at compile time is set at run time, by options AFTER the input file or
 * by LBMDEM_* environment variables (options win); without any of them the run is the reference's
 * default build (lx = 7826, ly = 2325, scale = 1, fp64, duration = 1.5 s):
 *   --lx N --ly N --scale S --single     -Dlx -Dly -Dscale -DSINGLE_PRECISION      LBMDEM_LX LBMDEM_LY LBMDEM_SCALE LBMDEM_SINGLE
 *   --duration T                         #define duration (src/main.c:47)          LBMDEM_DURATION
 *   --steps N                            stop after N renderScene() calls          LBMDEM_STEPS
 *   --strict                             bit-exact build (lbmdem_params.strict_fp) LBMDEM_STRICT
 *   --vib                                vibrating walls, int vib = 1 (:162)       LBMDEM_VIB
 *   --device D, --outdir DIR             CUDA device, directory of the output files (default: cwd)
 *   --gpus N                             x-strip decomposition over N GPUs of this box: the process forks N-1
 *                                        ranks (one process per GPU, devices D .. D+N-1), the ranks meet over
 *                                        NCCL inside the library; rank 0 prints and writes every file       LBMDEM_GPUS
 *   --restart FILE                       continue from a checkpoint instead of reading positions from <inputfile>
 *   --checkpoint FILE                    write a checkpoint when the run ends (lbmdem_save_state)
 * Outputs, as the reference writes them: stdout banner and progress lines, stderr
 * "final_density: %f", stats.data, DEM%06d.dat every 4000 calls, five VTK files every 8000 calls.
 * (The PostScript contact plot DEM%06d.ps of write_forces() is not produced.)
 */
#include <math.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "../../include/lbmdem_gpu.h"
#include "dem_output.h"
#include "vtk_writer.h"

#define STEP_CONSOLE 400  /* src/main.c:138-143 */
#define STEP_STROB 4000
#define STEP_FILM 8000

static lbmdem_ctx *ctx;

/* ---- ranks of a multi-GPU run: forked processes that share one anonymous mapping ---- */
typedef struct {
  volatile int id_ready;          /* rank 0 has published the NCCL id */
  char nccl_id[128];
  volatile int arrived, sense;    /* sense-reversing barrier */
  volatile int failed;            /* some rank died: everybody leaves */
  volatile int pids[64];          /* the ranks' process ids: a rank that fails takes down peers that may sit in a collective */
  double partial[64];             /* per-rank contribution to a sum (density checksum) */
} shared_t;
static shared_t *sh;
static float *sh_fields;          /* the five VTK point fields of the WHOLE lattice, [y][x] order */
static int rank, nranks = 1;

static void barrier(void) {
  if (nranks == 1) return;
  const int my = !sh->sense;
  if (__atomic_add_fetch(&sh->arrived, 1, __ATOMIC_ACQ_REL) == nranks) {
    sh->arrived = 0;
    __atomic_store_n(&sh->sense, my, __ATOMIC_RELEASE);
  } else {
    while (__atomic_load_n(&sh->sense, __ATOMIC_ACQUIRE) != my) {
      if (sh->failed) exit(EXIT_FAILURE);
      usleep(50);
    }
  }
}
/* sum over the ranks of one double each; every rank gets the result */
static double rank_sum(double v) {
  if (nranks == 1) return v;
  sh->partial[rank] = v;
  barrier();
  double s = 0;
  for (int r = 0; r < nranks; ++r) s += sh->partial[r];
  barrier();
  return s;
}

/* Any fatal path of a multi-rank run: tell the peers (those waiting at the barrier leave by themselves) and, after a
 * moment, terminate the ones that are blocked inside an NCCL collective waiting for this rank -- by process id, never
 * by group or pattern. */
static void fail_run(void) {
  if (!sh) exit(EXIT_FAILURE);
  sh->failed = 1;
  usleep(200000);
  for (int r = 0; r < nranks; ++r)
    if (r != rank && sh->pids[r] > 0) kill((pid_t)sh->pids[r], SIGTERM);
  exit(EXIT_FAILURE);
}
static void die(const char *what) {
  fprintf(stderr, "lbmdem: rank %d: %s: %s\n", rank, what, lbmdem_last_error(ctx));
  fail_run();
}
#define CK(call) do { if ((call) < 0) die(#call); } while (0)

static const char *opt_or_env(int argc, char **argv, const char *opt, const char *env) {
  for (int i = 2; i + 1 < argc; ++i)
    if (!strcmp(argv[i], opt)) return argv[i + 1];
  return getenv(env);
}
static int flag_or_env(int argc, char **argv, const char *opt, const char *env) {
  for (int i = 2; i < argc; ++i)
    if (!strcmp(argv[i], opt)) return 1;
  const char *e = getenv(env);
  return e && *e && strcmp(e, "0");
}

/* check_sample (src/main.c:640-658): extent, mass and packing fraction of the sample */
static void check_sample(int n, const double *g, double rhoS) {
  double xMax = g[0], xMin = g[0], yMax = g[1], yMin = g[1], mass = 0.;
  for (int i = 0; i < n; ++i) {
    const double *q = g + 13 * (size_t)i;
    mass += q[10];
    xMax = fmax(xMax, q[0] + q[9]);
    xMin = fmin(xMin, q[0] - q[9]);
    yMax = fmax(yMax, q[1] + q[9]);
    yMin = fmin(yMin, q[1] - q[9]);
  }
  const double L0 = xMax - xMin, H0 = yMax - yMin;
  printf("L0=%le H0=%le Mass of Grains=%le Phi=%le\n", L0, H0, mass, mass / (rhoS * (L0 * H0)));
}

int main(int argc, char **argv) {
  time_t now;
  printf("2D LBM-DEM code\n");
  if (argc < 2 || argv[1][0] == '-') {
    printf("usage: usage %s <filename>\n", argv[0]);
    exit(EXIT_FAILURE);
  }
  printf("Opening file : %s\n", argv[1]);

  lbmdem_params p;
  lbmdem_default_params(&p);
  const char *v;
  if ((v = opt_or_env(argc, argv, "--lx", "LBMDEM_LX"))) p.lx = atoi(v);
  if ((v = opt_or_env(argc, argv, "--ly", "LBMDEM_LY"))) p.ly = atoi(v);
  if ((v = opt_or_env(argc, argv, "--scale", "LBMDEM_SCALE"))) p.scale = atof(v);
  if ((v = opt_or_env(argc, argv, "--device", "LBMDEM_DEVICE"))) p.device = atoi(v);
  p.single_precision = flag_or_env(argc, argv, "--single", "LBMDEM_SINGLE");
  p.strict_fp = flag_or_env(argc, argv, "--strict", "LBMDEM_STRICT");
  p.vib = flag_or_env(argc, argv, "--vib", "LBMDEM_VIB"); /* int vib (src/main.c:162) */
  double duration = 1.5; /* src/main.c:47 */
  long max_steps = -1;
  if ((v = opt_or_env(argc, argv, "--duration", "LBMDEM_DURATION"))) duration = atof(v);
  if ((v = opt_or_env(argc, argv, "--steps", "LBMDEM_STEPS"))) max_steps = atol(v);
  const char *outdir = opt_or_env(argc, argv, "--outdir", "LBMDEM_OUTDIR");
  if (!outdir) outdir = "";
  if ((v = opt_or_env(argc, argv, "--gpus", "LBMDEM_GPUS"))) nranks = atoi(v);
  if (nranks < 1 || nranks > 64) { fprintf(stderr, "lbmdem: --gpus must be between 1 and 64\n"); return EXIT_FAILURE; }
  const size_t nn_all = (size_t)p.lx * p.ly;
  if (nranks > 1) {
    /* before any CUDA call: share a mapping, then fork one process per further GPU */
    fflush(stdout);
    sh = mmap(NULL, sizeof *sh, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    sh_fields = mmap(NULL, sizeof(float) * 11 * nn_all, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (sh == MAP_FAILED || sh_fields == MAP_FAILED) { fprintf(stderr, "lbmdem: mmap failed\n"); return EXIT_FAILURE; }
    memset(sh, 0, sizeof *sh);
    for (int r = 1; r < nranks; ++r) {
      const pid_t pid = fork();
      if (pid < 0) { fprintf(stderr, "lbmdem: fork failed\n"); return EXIT_FAILURE; }
      if (pid == 0) { rank = r; break; }
    }
    sh->pids[rank] = (int)getpid();
    p.rank = rank; p.nranks = nranks; p.device += rank;
    if (rank != 0) { /* only rank 0 talks */
      if (!freopen("/dev/null", "w", stdout)) return EXIT_FAILURE;
    }
  }

  if (lbmdem_create(&p, &ctx)) {
    fprintf(stderr, "lbmdem: rank %d: %s\n", rank, lbmdem_last_error(NULL));
    fail_run();
  }
  if (nranks > 1) {
    if (rank == 0) {
      if (lbmdem_nccl_unique_id(sh->nccl_id)) { fprintf(stderr, "lbmdem: %s\n", lbmdem_last_error(NULL)); fail_run(); }
      __atomic_store_n(&sh->id_ready, 1, __ATOMIC_RELEASE);
    } else {
      while (!__atomic_load_n(&sh->id_ready, __ATOMIC_ACQUIRE)) {
        if (sh->failed) exit(EXIT_FAILURE);
        usleep(100);
      }
    }
    char id[128];
    memcpy(id, sh->nccl_id, 128);
    CK(lbmdem_attach_nccl(ctx, id));
  }
  int xlo = 0, xhi = p.lx;
  CK(lbmdem_get_strip(ctx, &xlo, &xhi));
  /* read_sample prints the comment line and the grain count (src/main.c:613-619) */
  {
    FILE *fp = fopen(argv[1], "r");
    char com[256];
    int cnt = 0;
    if (!fp || !fgets(com, sizeof com, fp) || fscanf(fp, "%d\n", &cnt) != 1) {
      fprintf(stderr, "lbmdem: cannot read %s\n", argv[1]);
      fail_run();
    }
    fclose(fp);
    printf("%s\n", com);
    printf("Nb grains %d\n", cnt);
  }
  const char *restart = opt_or_env(argc, argv, "--restart", "LBMDEM_RESTART");
  const char *checkpoint = opt_or_env(argc, argv, "--checkpoint", "LBMDEM_CHECKPOINT");
  const int n = restart ? lbmdem_load_state(ctx, restart) : lbmdem_load_sample(ctx, argv[1]);
  if (n < 0) die(restart ? "restart" : "read_sample");

  double d11[11];
  long l4[4];
  CK(lbmdem_get_scalars(ctx, d11, l4));
  const double dx = d11[0], dtLB = d11[1], dt = d11[2], c = d11[4];
  const int npDEM = (int)l4[0];
  double *grains = malloc(sizeof(double) * 13 * (size_t)n), *fhf = malloc(sizeof(double) * 3 * (size_t)n);
  double *mid = malloc(sizeof(double) * 6 * (size_t)n), *diag = malloc(sizeof(double) * 17 * (size_t)n);
  double *gp = malloc(sizeof(double) * (size_t)n);
  const int cap = p.neighbour_capacity;
  int *cnt = malloc(sizeof(int) * (size_t)n), *nbr = malloc(sizeof(int) * (size_t)n * cap), *wfl = malloc(sizeof(int) * (size_t)n);
  if (!grains || !fhf || !mid || !diag || !gp || !cnt || !nbr || !wfl) fail_run();
  CK(lbmdem_get_grains(ctx, grains));
  check_sample(n, grains, p.rhoS);
  printf("no space %le\n", dx);
  {
    /* dtmax as main() prints it (src/main.c:1848-1856) */
    double rMin = grains[9];
    for (int i = 1; i < n; ++i) rMin = fmin(rMin, grains[13 * (size_t)i + 9]);
    const double dtmax = (1 / p.iterDEM) * 3.14159265358979 * rMin * sqrt(3.14159265358979 * p.rhoS / p.kg);
    printf("dtLB=%le,  dtmax=%le,   dt=%le,   npDEM=%d,   c=%lf\n", dtLB, dtmax, dt, npDEM, c);
  }
  time(&now);
  printf("Current local time and date: %s", asctime(localtime(&now)));
  if (rank == 0 && !restart && lbmdem_write_stats_header(outdir)) { fprintf(stderr, "lbmdem: cannot write stats.data\n"); fail_run(); }
  lbmdem_diag *dg = lbmdem_diag_create(n, &p);

  const size_t nn = (size_t)(xhi - xlo) * p.ly; /* nodes of this rank's strip */
  float *f_gp = NULL, *f_gv = NULL, *f_ga = NULL, *f_fp = NULL, *f_fv = NULL;
  long nbsteps = l4[1]; /* 0, or where the checkpoint was taken */
  int nFile = (int)l4[2];
  double summary[7] = {0, 0, 0, 0, 0, 0, 0};
  /* main loop: do { renderScene(); ... } while (nbsteps * dt <= duration)  (src/main.c:1880-1890) */
  int more = 1;
  while (more) {
    /* Run freely up to the next point where the host has something to do:
     *   - two calls before an output call (the diagnostics replay needs the state of both),
     *   - the console cadence (every UpdateVerlet calls, :1884),
     *   - right after a call that printed the density (:1715: nbsteps % stepConsole == 0 before the
     *     increment, on an LBM call),
     *   - the end of the run: the do/while makes call k+1 iff k * dt <= duration. */
    const long next_out = (nbsteps / STEP_STROB + 1) * STEP_STROB;
    long stop = next_out - 2;
    const long console = (nbsteps / p.UpdateVerlet + 1) * p.UpdateVerlet;
    if (stop > console) stop = console;
    const long dens = (nbsteps == 0) ? 1 : ((nbsteps - 1) / STEP_CONSOLE + 1) * STEP_CONSOLE + 1;
    if (stop > dens) stop = dens;
    const long last = (long)floor(duration / dt) + 1;
    if (stop > last) stop = last;
    if (max_steps >= 0 && stop > max_steps) stop = max_steps;
    if (stop > nbsteps) {
      CK(lbmdem_step(ctx, stop - nbsteps));
      nbsteps = stop;
    } else {
      /* one of the two calls before an output, or the output call itself */
      CK(lbmdem_step_capture(ctx, mid));
      ++nbsteps;
      CK(lbmdem_get_scalars(ctx, d11, l4));
      CK(lbmdem_get_verlet(ctx, cnt, nbr, cap, wfl));
      CK(lbmdem_get_grains(ctx, grains));
      CK(lbmdem_get_fhf(ctx, fhf)); /* an LBM step inside the call comes before its contact loop */
      lbmdem_diag_pass(dg, mid, grains, fhf, cnt, nbr, cap, wfl, d11);
    }
    if ((nbsteps - 1) % STEP_CONSOLE == 0 && (nbsteps - 1) % npDEM == 0) {
      double sum; /* check_density (:1249-1260) of the call that just ended */
      CK(lbmdem_total_density(ctx, &sum));
      sum = rank_sum(sum);
      printf("Iteration Number %ld, Total density in the system %f\n", nbsteps - 1, sum);
    }
    if (nbsteps % STEP_FILM == 0) { /* write_vtk, nFile++ (:1767-1772) */
      if (!f_gp) {
        f_gp = malloc(4 * nn); f_gv = malloc(12 * nn); f_ga = malloc(12 * nn); f_fp = malloc(4 * nn); f_fv = malloc(12 * nn);
        if (!f_gp || !f_gv || !f_ga || !f_fp || !f_fv) fail_run();
      }
      lbmdem_diag_get(dg, diag);
      for (int i = 0; i < n; ++i) gp[i] = diag[17 * (size_t)i];
      CK(lbmdem_get_fields(ctx, gp, f_gp, f_gv, f_ga, f_fp, f_fv));
      const float *w_gp = f_gp, *w_gv = f_gv, *w_ga = f_ga, *w_fp = f_fp, *w_fv = f_fv;
      if (nranks > 1) {
        /* every rank drops its columns [xlo, xhi) of each [y][x] field into the shared arrays */
        float *all[5] = {sh_fields, sh_fields + nn_all, sh_fields + 4 * nn_all, sh_fields + 7 * nn_all, sh_fields + 8 * nn_all};
        const float *mine[5] = {f_gp, f_gv, f_ga, f_fp, f_fv};
        const int comp[5] = {1, 3, 3, 1, 3};
        const int w = xhi - xlo;
        for (int k = 0; k < 5; ++k)
          for (int y = 0; y < p.ly; ++y)
            memcpy(all[k] + ((size_t)y * p.lx + xlo) * comp[k], mine[k] + (size_t)y * w * comp[k], sizeof(float) * w * comp[k]);
        barrier();
        w_gp = all[0]; w_gv = all[1]; w_ga = all[2]; w_fp = all[3]; w_fv = all[4];
      }
      if (rank == 0 && lbmdem_write_vtk_frame(outdir, nFile, p.lx, p.ly, w_gp, w_gv, w_ga, w_fp, w_fv)) {
        fprintf(stderr, "lbmdem: cannot write the VTK frame\n");
        fail_run();
      }
      barrier(); /* the shared arrays are free again */
      nFile++;
    }
    if (nbsteps % STEP_STROB == 0) { /* write_DEM (:1773-1776) */
      CK(lbmdem_get_fhf(ctx, fhf));
      if (rank == 0 && lbmdem_write_dem(dg, outdir, nFile, nbsteps, grains, fhf, d11, summary)) {
        fprintf(stderr, "lbmdem: cannot write DEM%06d.dat\n", nFile);
        fail_run();
      }
    }
    if (nbsteps % p.UpdateVerlet == 0) {
      time(&now);
      printf("steps %li steps %le KE %le PE %le SE %le WF %le INCE %le SLIP %le RW %le Time %s \n", nbsteps, nbsteps * dt,
             summary[0], summary[1], summary[2], summary[3], summary[4], summary[5], summary[6], asctime(localtime(&now)));
    }
    more = (nbsteps * dt <= duration) && (max_steps < 0 || nbsteps < max_steps);
  }
  {
    double sum; /* final_density (:1262-1273) */
    CK(lbmdem_total_density(ctx, &sum));
    sum = rank_sum(sum);
    if (rank == 0) fprintf(stderr, "final_density: %f\n", sum);
  }
  if (checkpoint) CK(lbmdem_save_state(ctx, checkpoint));
  time(&now);
  printf("End local time and date: %s", asctime(localtime(&now)));
  lbmdem_diag_destroy(dg);
  barrier();
  lbmdem_destroy(ctx);
  if (nranks > 1 && rank == 0) {
    int status = 0, bad = 0;
    while (wait(&status) > 0)
      if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad = 1;
    if (bad) return EXIT_FAILURE;
  }
  return 0;
}
