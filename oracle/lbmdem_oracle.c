/*
 * oracle/lbmdem_oracle.c -- TEST INFRASTRUCTURE (checker), never linked into the product.
 *
 * Serial CPU restatement of the coupled LBM-DEM step of cb-geo/2d-lbm-dem.  Every routine
 * names the lines of /root/reference/src/main.c it restates.  The arithmetic keeps the
 * reference's operand order and its int/float/double promotions (they matter in the
 * -DSINGLE_PRECISION build, where literals such as `1.` or `2.` silently widen an
 * expression to double), because the pin for this file is bit-identity with the compiled
 * reference -- see lbmdem_oracle.h.  What is different from the reference: run-time lattice
 * size, one heap object instead of ~80 globals, flat index arithmetic, no file I/O, and an
 * optional moving-lid term that the reference only carries as a comment.
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off [-DORACLE_SINGLE] (oracle/build.py).
 */
#define _GNU_SOURCE
#include "lbmdem_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef ORACLE_SINGLE
typedef float real;
#define SFX(name) name##_f32
#define SCAN3 "%e %e %e;\n"
#else
typedef double real;
#define SFX(name) name##_f64
#define SCAN3 "%le %le %le;\n"
#endif
#define API __attribute__((visibility("default")))

#define NQ 9
#define HALF 4
#define PI_REF 3.14159265358979 /* src/main.c:42 */
#define RHO_SOLID 2650          /* src/main.c:44 */

/* src/main.c:70-71 */
static const int EX[NQ] = {0, -1, -1, -1, 0, 1, 1, 1, 0};
static const int EY[NQ] = {0, 1, 0, -1, -1, -1, 0, 1, 1};

/* src/main.c:182-198 */
typedef struct {
  real x1, x2, x3, v1, v2, v3, a1, a2, a3;
  real r, m, mw, It;
  real p, s, f1, f2, ifm, fm, fr, ifr, M11, M12, M21, M22, ice, slip, rw;
  int z, zz;
} grain_t;

typedef struct { real f1, f2, f3; } force_t;

struct oracle {
  int lx, ly, n;
  double scale;
  /* lattice */
  real *f, *delta;
  int *obst, *act;
  real w[NQ];
  /* src/main.c:52,74-94 */
  real dx, dtLB, c, rho_moy, tau, s2, s3, s5, s7, s8, s9, nu, reductionR;
  /* src/main.c:97-118 */
  real G, angleG, xG, yG, dt, dt2, km, kg, kt, ktm, nug, num, numb, nugt, mu, mum, mumb, murf,
      rscale, distVerlet, dtt, iterDEM, freq, amp, t;
  long UpdateVerlet, nbsteps;
  int npDEM, stepFilm, nFile;
  real pf, pft, pff, ic; /* src/main.c:130-131 */
  real Mby, Mgx, Mhy, Mdx; /* src/main.c:201-204 */
  real lid; /* extension, 0 = reference behaviour */
  int vib;  /* src/main.c:162 */
  grain_t *g;
  real *rLB, *fhf1, *fhf2, *fhf3;
  int *cumul, *neigh, *wallB, *wallT, *wallL, *wallR;
  int nB, nT, nL, nR;
};

#define FI(o, x, y, q) ((o)->f[((size_t)(x) * (o)->ly + (y)) * NQ + (q)])
#define DI(o, x, y, q) ((o)->delta[((size_t)(x) * (o)->ly + (y)) * NQ + (q)])
#define OB(o, x, y) ((o)->obst[(size_t)(x) * (o)->ly + (y)])
#define AC(o, x, y) ((o)->act[(size_t)(x) * (o)->ly + (y)])

/* ------------------------------------------------------------------------------------------ */
API oracle *SFX(oracle_create)(int lx, int ly, double scale) {
  oracle *o = calloc(1, sizeof *o);
  if (!o) return NULL;
  o->lx = lx; o->ly = ly; o->scale = scale;
  size_t nn = (size_t)lx * ly;
  o->f = malloc(sizeof(real) * nn * NQ);
  o->delta = calloc(nn * NQ, sizeof(real));
  o->obst = malloc(sizeof(int) * nn);
  o->act = calloc(nn, sizeof(int));
  if (!o->f || !o->delta || !o->obst || !o->act) return NULL;
  /* src/main.c:53-54 */
  const real w0[NQ] = {4. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9};
  memcpy(o->w, w0, sizeof w0);
  o->rho_moy = 1000; o->tau = 0.504;
  o->s2 = 1.5; o->s3 = 1.4; o->s5 = 1.5; o->s7 = 1.5; o->s8 = 1.9841; o->s9 = 1.9841;
  o->nu = 1e-6; o->reductionR = 0.85;
  o->G = 9.81; o->angleG = 0.0;
  o->km = 3e+6; o->kg = 1.6e+6; o->kt = 1.0e+6; o->ktm = 2e+6;
  o->nug = 6.4e+1; o->num = 8.7e+1; o->numb = 8.7e+1; o->nugt = 5e-1;
  o->mu = .5317; o->mum = .466; o->mumb = .466; o->murf = 0.01;
  o->rscale = 1e-3; o->distVerlet = 5e-4; o->UpdateVerlet = 100; o->dtt = 0.; o->iterDEM = 100.;
  o->freq = 5; o->amp = 4.e-4; o->t = 0; o->stepFilm = 8000;
  return o;
}

static void free_grains(oracle *o) {
  free(o->g); free(o->rLB); free(o->fhf1); free(o->fhf2); free(o->fhf3); free(o->cumul);
  free(o->neigh); free(o->wallB); free(o->wallT); free(o->wallL); free(o->wallR);
  o->g = NULL;
}

API void SFX(oracle_destroy)(oracle *o) {
  if (!o) return;
  free_grains(o);
  free(o->f); free(o->delta); free(o->obst); free(o->act); free(o);
}

static int alloc_grains(oracle *o, int n) {
  free_grains(o);
  o->n = n;
  o->g = calloc(n, sizeof(grain_t)); /* mw = 0: SURVEY App. B #2 */
  o->rLB = calloc(n, sizeof(real));
  o->fhf1 = calloc(n, sizeof(real)); o->fhf2 = calloc(n, sizeof(real));
  o->fhf3 = calloc(n, sizeof(real));
  o->cumul = calloc(n, sizeof(int));
  o->neigh = calloc((size_t)n * 64 + 64, sizeof(int)); /* reference: 6n, print-only overflow */
  o->wallB = calloc(n, sizeof(int)); o->wallT = calloc(n, sizeof(int));
  o->wallL = calloc(n, sizeof(int)); o->wallR = calloc(n, sizeof(int));
  return (o->g && o->neigh) ? 0 : -1;
}

/* src/main.c:624-635: metres, mass and inertia of one freshly read grain */
static void finish_grain(oracle *o, grain_t *p) {
  p->m = RHO_SOLID * PI_REF * p->r * p->r;
  p->It = p->m * p->r * p->r / 2;
  p->x3 = 0.; p->v1 = 0.; p->v2 = 0.; p->v3 = 0.; p->a1 = 0.; p->a2 = 0.; p->a3 = 0.;
  (void)o;
}

/* src/main.c:609-639 */
API int SFX(oracle_read_sample)(oracle *o, const char *path) {
  FILE *fp = fopen(path, "r");
  if (!fp) return -1;
  char line[256];
  int n = 0;
  if (!fgets(line, sizeof line, fp) || fscanf(fp, "%d\n", &n) != 1 || n <= 0) { fclose(fp); return -2; }
  if (alloc_grains(o, n)) { fclose(fp); return -3; }
  for (int i = 0; i < n; ++i) {
    grain_t *p = &o->g[i];
    if (fscanf(fp, SCAN3, &p->r, &p->x1, &p->x2) != 3) { fclose(fp); return -4; }
    p->r = p->r * o->rscale;
    p->x1 = p->x1 * o->rscale;
    p->x2 = p->x2 * o->rscale;
    finish_grain(o, p);
  }
  fclose(fp);
  return n;
}

API int SFX(oracle_set_grains)(oracle *o, int n, const double *r, const double *x1, const double *x2) {
  if (alloc_grains(o, n)) return -3;
  for (int i = 0; i < n; ++i) {
    grain_t *p = &o->g[i];
    p->r = (real)r[i]; p->x1 = (real)x1[i]; p->x2 = (real)x2[i];
    finish_grain(o, p);
  }
  return n;
}

API void SFX(oracle_set_lid)(oracle *o, double uw) { o->lid = (real)uw; }
API void SFX(oracle_set_vib)(oracle *o, int vib) { o->vib = vib; }
/* src/main.c:117 (read by VerletWall, :1555-1561) and :98 (main() derives xG, yG from it, :1841-1842) */
API void SFX(oracle_set_dtt)(oracle *o, double dtt) { o->dtt = (real)dtt; }
API void SFX(oracle_set_angleG)(oracle *o, double a) {
  o->angleG = (real)a;
  o->xG = -o->G * sin(o->angleG);
  o->yG = -o->G * cos(o->angleG);
}

/* src/main.c:663-711 */
static void init_obst(oracle *o) {
  const int lx = o->lx, ly = o->ly;
  for (int x = 1; x < lx - 1; x++)
    for (int y = 1; y < ly - 1; y++) OB(o, x, y) = -1;
  for (int x = 0; x < lx; x++) {
    OB(o, x, 0) = OB(o, x, ly - 1) = o->n;
    AC(o, x, 0) = AC(o, x, ly - 1) = 0;
  }
  for (int y = 1; y < ly - 1; y++) {
    OB(o, 0, y) = OB(o, lx - 1, y) = o->n;
    AC(o, 0, y) = AC(o, lx - 1, y) = 0;
  }
  for (int i = 0; i < o->n; i++) {
    real xc = (o->g[i].x1 - o->Mgx) / o->dx;
    real yc = (o->g[i].x2 - o->Mby) / o->dx;
    real r2 = o->rLB[i] * o->rLB[i];
    real rbl0 = o->g[i].r / o->dx;
    real R2 = rbl0 * rbl0;
    int xi = (int)(xc - rbl0), xf = (int)(xc + rbl0);
    if (xi < 1) xi = 1;
    if (xf >= lx - 1) xf = lx - 2;
    int yi = (int)(yc - rbl0), yf = (int)(yc + rbl0);
    if (yi < 1) yi = 1;
    if (yf >= ly - 1) yf = ly - 2;
    for (int x = xi; x <= xf; x++)
      for (int y = yi; y <= yf; y++) {
        real dist2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
        if (dist2 <= R2 && dist2 <= r2) OB(o, x, y) = i;
      }
  }
}

/* src/main.c:1834-1861 */
API void SFX(oracle_init)(oracle *o) {
  const int lx = o->lx, ly = o->ly;
  for (size_t k = 0; k < (size_t)lx * ly; ++k) /* init_density :716-724 */
    for (int q = 0; q < NQ; ++q) o->f[k * NQ + q] = o->w[q];
  o->Mgx = 0.;
  o->Mdx = 1.e-3 * lx / 10;
  o->Mhy = 1.e-3 * ly / 10;
  o->Mby = 0.;
  o->xG = -o->G * sin(o->angleG);
  o->yG = -o->G * cos(o->angleG);
  o->dx = (1. / o->scale) * (o->Mdx - o->Mgx) / (lx - 1);
  real rMin = o->g[0].r; /* :219-225 */
  for (int i = 1; i <= o->n - 1; i++) rMin = fmin(rMin, o->g[i].r);
  real dtmax = (1 / o->iterDEM) * PI_REF * rMin * sqrt(PI_REF * RHO_SOLID / o->kg);
  o->dtLB = o->dx * o->dx * (o->tau - 0.5) / (3 * o->nu);
  o->npDEM = (o->dtLB / dtmax + 1);
  o->c = o->dx / o->dtLB;
  o->dt = o->dtLB / o->npDEM;
  o->dt2 = o->dt * o->dt;
  for (int i = 0; i <= o->n - 1; i++) o->rLB[i] = o->reductionR * o->g[i].r / o->dx;
  init_obst(o);
  o->nbsteps = 0; o->nFile = 0; o->t = 0;
  o->pf = o->pft = o->pff = 0.; o->ic = 0;
  o->nB = o->nT = o->nL = o->nR = 0;
}

/* ---------------------------------------------------------------------------------------------
 * LBM phases
 * ------------------------------------------------------------------------------------------ */

/* rigid-body velocity of grain i at lattice node (x,y): the sub-expressions of :974-980 */
#define UWX(o, gi, y) ((gi)->v1 - ((y) * (o)->dx + (o)->Mby - (gi)->x2) * (gi)->v3)
#define UWY(o, gi, x) ((gi)->v2 + ((x) * (o)->dx + (o)->Mgx - (gi)->x1) * (gi)->v3)

/* src/main.c:966-986 */
API void SFX(oracle_reinit_obst_density)(oracle *o) {
  const real c = o->c;
  for (int x = 1; x < o->lx - 1; x++)
    for (int y = 1; y < o->ly - 1; y++) {
      int i = OB(o, x, y);
      if (i == -1) continue;
      const grain_t *gi = &o->g[i];
      real u_squ = (UWX(o, gi, y) * UWX(o, gi, y) + UWY(o, gi, x) * UWY(o, gi, x)) / (c * c);
      for (int q = 0; q < NQ; q++) {
        real eu = (EX[q] * UWX(o, gi, y) + EY[q] * UWY(o, gi, x)) / c;
        FI(o, x, y, q) = o->w[q] * (1. + 3 * eu + 4.5 * eu * eu - 1.5 * u_squ);
      }
    }
}

/* src/main.c:991-1065 (serial semantics: grains in index order) */
API void SFX(oracle_obst_construction)(oracle *o) {
  const int lx = o->lx, ly = o->ly;
  for (int x = 1; x < lx - 1; x++)
    for (int y = 1; y < ly - 1; y++) {
      OB(o, x, y) = -1;
      AC(o, x, y) = 1;
      for (int q = 1; q < NQ; q++) DI(o, x, y, q) = 0;
    }
  for (int i = 0; i < o->n; i++) {
    real xc = (o->g[i].x1 - o->Mgx) / o->dx;
    real yc = (o->g[i].x2 - o->Mby) / o->dx;
    real r2 = o->rLB[i] * o->rLB[i];
    real rbl0 = o->g[i].r / o->dx;
    real R2 = rbl0 * rbl0;
    int xi = (int)(xc - rbl0), xf = (int)(xc + rbl0);
    if (xi < 1) xi = 1;
    if (xf >= lx - 1) xf = lx - 2;
    int yi = (int)(yc - rbl0), yf = (int)(yc + rbl0);
    if (yi < 1) yi = 1;
    if (yf >= ly - 1) yf = ly - 2;
    for (int y = yi; y <= yf; y++)
      for (int x = xi; x <= xf; x++) {
        real dist2 = (x - xc) * (x - xc) + (y - yc) * (y - yc);
        if (dist2 <= R2 && dist2 <= r2) OB(o, x, y) = i;
      }
    for (int y = yi; y <= yf; y++)
      for (int x = xi; x <= xf; x++) {
        if (OB(o, x, y) != i) continue;
        AC(o, x, y) = 0;
        for (int q = 1; q < NQ; q++) {
          int nx = x + EX[q], ny = y + EY[q];
          if (OB(o, nx, ny) != -1) continue;
          AC(o, x, y) = 1;
          real aa = fabs(EX[q]) + fabs(EY[q]);
          real bb = (x + EX[q] - xc) * EX[q] + (y + EY[q] - yc) * EY[q];
          real cc = (x + EX[q] - xc) * (x + EX[q] - xc) + (y + EY[q] - yc) * (y + EY[q] - yc) - r2;
          DI(o, x, y, q) = (bb - sqrt(fabs(bb * bb - aa * cc))) / aa;
        }
      }
  }
}

/* :1082-1116 */
static void mrt_collide(const oracle *o, real *p) {
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
  real eO = e - o->s2 * (e + 2 * rho - 3 * (j_x2 + j_y2) / rho);
  real epsO = eps - o->s3 * (eps - rho + 3 * (j_x2 + j_y2) / rho);
  real q_xO = q_x - o->s5 * (q_x + j_x);
  real q_yO = q_y - o->s7 * (q_y + j_y);
  real p_xxO = p_xx - o->s8 * (p_xx - (j_x2 - j_y2) / rho);
  real p_xyO = p_xy - o->s9 * (p_xy - j_x * j_y / rho);
  p[0] = a * (4 * rho - 4 * eO + 4 * epsO);
  p[2] = a * (4 * rho - eO - 2 * epsO - 6 * j_x + 6 * q_xO + 9 * p_xxO);
  p[4] = a * (4 * rho - eO - 2 * epsO - 6 * j_y + 6 * q_yO - 9 * p_xxO);
  p[6] = a * (4 * rho - eO - 2 * epsO + 6 * j_x - 6 * q_xO + 9 * p_xxO);
  p[8] = a * (4 * rho - eO - 2 * epsO + 6 * j_y - 6 * q_yO - 9 * p_xxO);
  p[1] = a * (4 * rho + 2 * eO + epsO - 6 * j_x - 3 * q_xO + 6 * j_y + 3 * q_yO - 9 * p_xyO);
  p[3] = a * (4 * rho + 2 * eO + epsO - 6 * j_x - 3 * q_xO - 6 * j_y - 3 * q_yO + 9 * p_xyO);
  p[5] = a * (4 * rho + 2 * eO + epsO + 6 * j_x + 3 * q_xO - 6 * j_y - 3 * q_yO - 9 * p_xyO);
  p[7] = a * (4 * rho + 2 * eO + epsO + 6 * j_x + 3 * q_xO + 6 * j_y + 3 * q_yO + 9 * p_xyO);
}

/* :1159-1186 / :1190-1218 -- one link q of active solid node (x,y) owned by grain gi */
static void grain_link(oracle *o, const grain_t *gi, int x, int y, int q) {
  const int oq = (q <= HALF) ? q + HALF : q - HALF;
  const int nx = x + EX[q], ny = y + EY[q];
  if (OB(o, nx, ny) != -1) {
    FI(o, x, y, q) = o->w[q];
    return;
  }
  const real d = DI(o, x, y, q);
  if (d >= 0.5) {
    FI(o, x, y, q) = FI(o, nx, ny, oq) / (2 * d) + (2 * d - 1) * FI(o, nx, ny, q) / (2 * d) +
                     3 * (o->w[q] / o->c) * (EX[q] * UWX(o, gi, y) + EY[q] * UWY(o, gi, x)) / d;
  }
  if (d > 0. && d < 0.5) {
    const int nnx = nx + EX[q], nny = ny + EY[q];
    FI(o, x, y, q) = 2 * d * FI(o, nx, ny, oq) + (1 - 2 * d) * FI(o, nnx, nny, oq) +
                     6 * (o->w[q] / o->c) * (EX[q] * UWX(o, gi, y) + EY[q] * UWY(o, gi, x));
  }
}

static inline void swap_real(real *a, real *b) { real t = *a; *a = *b; *b = t; }

/* src/main.c:1071-1243 */
API void SFX(oracle_collision_streaming)(oracle *o) {
  const int lx = o->lx, ly = o->ly;
  /* phase 1, :1077-1119 */
  for (int x = 1; x < lx - 1; x++)
    for (int y = 1; y < ly - 1; y++)
      if (OB(o, x, y) == -1) mrt_collide(o, &FI(o, x, y, 0));
  /* phase 2, :1123-1145 (the lid term is the commented uw_h/6 of :1129-1130) */
  const real lid6 = o->lid / 6;
  for (int x = 1; x < lx - 1; x++) {
    FI(o, x, 0, 8) = FI(o, x, 1, 4);
    FI(o, x, 0, 7) = FI(o, x + 1, 1, 3);
    FI(o, x, 0, 1) = FI(o, x - 1, 1, 5);
    FI(o, x, ly - 1, 4) = FI(o, x, ly - 2, 8);
    if (o->lid != 0) {
      FI(o, x, ly - 1, 3) = FI(o, x - 1, ly - 2, 7) - lid6;
      FI(o, x, ly - 1, 5) = FI(o, x + 1, ly - 2, 1) + lid6;
    } else {
      FI(o, x, ly - 1, 3) = FI(o, x - 1, ly - 2, 7);
      FI(o, x, ly - 1, 5) = FI(o, x + 1, ly - 2, 1);
    }
  }
  for (int y = 1; y < ly - 1; y++) {
    FI(o, 0, y, 6) = FI(o, 1, y, 2);
    FI(o, 0, y, 7) = FI(o, 1, y + 1, 3);
    FI(o, 0, y, 5) = FI(o, 1, y - 1, 1);
    FI(o, lx - 1, y, 2) = FI(o, lx - 2, y, 6);
    FI(o, lx - 1, y, 3) = FI(o, lx - 2, y - 1, 7);
    FI(o, lx - 1, y, 1) = FI(o, lx - 2, y + 1, 5);
  }
  FI(o, 0, 0, 7) = FI(o, 1, 1, 3);
  FI(o, lx - 1, 0, 1) = FI(o, lx - 2, 1, 5);
  FI(o, 0, ly - 1, 5) = FI(o, 1, ly - 2, 1);
  FI(o, lx - 1, ly - 1, 3) = FI(o, lx - 2, ly - 2, 7);
  /* phase 3, :1154-1222 */
  for (int x = 1; x < lx - 1; x++)
    for (int y = 1; y < ly - 1; y++) {
      int i = OB(o, x, y);
      if (i == -1 || AC(o, x, y) != 1) continue;
      for (int q = 1; q < NQ; q++) grain_link(o, &o->g[i], x, y, q);
    }
  /* phase 4, :1224-1242 */
  for (int x = 0; x < lx; x++)
    for (int y = 0; y < ly; y++)
      for (int q = 1; q <= HALF; q++) swap_real(&FI(o, x, y, q), &FI(o, x, y, q + HALF));
  for (int x = 0; x < lx; x++)
    for (int y = 0; y < ly; y++)
      for (int q = 1; q <= HALF; q++) {
        int nx = x + EX[q], ny = y + EY[q];
        if (nx >= 0 && ny >= 0 && nx < lx && ny < ly) swap_real(&FI(o, x, y, q + HALF), &FI(o, nx, ny, q));
      }
}

/* src/main.c:1249-1258 */
API double SFX(oracle_total_density)(oracle *o) {
  real sum = 0;
  for (size_t k = 0; k < (size_t)o->lx * o->ly * NQ; ++k) sum = sum + o->f[k];
  return (double)sum;
}

/* src/main.c:1285-1333 */
API void SFX(oracle_forces_fluid)(oracle *o) {
  const int nx = o->lx, ny = o->ly;
  for (int i = 0; i < o->n; ++i) {
    const grain_t *gi = &o->g[i];
    o->fhf1[i] = 0; o->fhf2[i] = 0; o->fhf3[i] = 0;
    const real xc = (gi->x1 - o->Mgx) / o->dx;
    const real yc = (gi->x2 - o->Mby) / o->dx;
    const real rbl0 = gi->r / o->dx;
    int xi = (int)(xc - rbl0), xf = (int)(xc + rbl0), yi = (int)(yc - rbl0), yf = (int)(yc + rbl0);
    if (xi < 1) xi = 1;
    if (xf > nx - 2) xf = nx - 2;
    if (yi < 1) yi = 1;
    if (yf > ny - 2) yf = ny - 2;
    for (int x = xi; x <= xf; ++x)
      for (int y = yi; y <= yf; ++y) {
        if (OB(o, x, y) != i) continue;
        for (int q = 1; q < NQ; ++q) {
          const int ax = x + EX[q], ay = y + EY[q];
          if (OB(o, ax, ay) == i) continue;
          const int oq = (q <= HALF) ? q + HALF : q - HALF;
          const real fnx = (FI(o, x, y, oq) + FI(o, ax, ay, q)) * EX[oq];
          const real fny = (FI(o, x, y, oq) + FI(o, ax, ay, q)) * EY[oq];
          o->fhf1[i] = o->fhf1[i] + fnx;
          o->fhf2[i] = o->fhf2[i] + fny;
          o->fhf3[i] = o->fhf3[i] - fnx * (y - (gi->x2 - o->Mby) / o->dx) + fny * (x - (gi->x1 - o->Mgx) / o->dx);
        }
      }
  }
  for (int i = 0; i < o->n; ++i) {
    o->fhf1[i] *= o->rho_moy * 9 * o->nu * o->nu / (o->dx * (o->tau - 0.5) * (o->tau - 0.5));
    o->fhf2[i] *= o->rho_moy * 9 * o->nu * o->nu / (o->dx * (o->tau - 0.5) * (o->tau - 0.5));
    o->fhf3[i] *= o->dx * o->rho_moy * 9 * o->nu * o->nu / (o->dx * (o->tau - 0.5) * (o->tau - 0.5));
  }
}

API void SFX(oracle_lbm_step)(oracle *o) {
  SFX(oracle_reinit_obst_density)(o);
  SFX(oracle_obst_construction)(o);
  SFX(oracle_collision_streaming)(o);
  SFX(oracle_forces_fluid)(o);
}

/* ---------------------------------------------------------------------------------------------
 * DEM
 * ------------------------------------------------------------------------------------------ */

/* src/main.c:211-216 */
static real maxt(real x, real y) { return (x < y) ? 0. : y; }

/* src/main.c:729-803 */
static force_t pair_force(oracle *o, long i, long j) {
  grain_t *gi = &o->g[i], *gj = &o->g[j];
  force_t F;
  double fn, ft;
  real xOiOj = gi->x1 - gj->x1, yOiOj = gi->x2 - gj->x2;
  real OiOj = sqrt(xOiOj * xOiOj + yOiOj * yOiOj);
  real dn = OiOj - gi->r - gj->r;
  if (dn >= 0) {
    F.f1 = 0; F.f2 = 0; F.f3 = 0;
    return F;
  }
  real vx = gi->v1 - gj->v1, vy = gi->v2 - gj->v2;
  real xn = xOiOj / OiOj, yn = yOiOj / OiOj;
  real vn = vx * xn + vy * yn;
  real vt = -vx * yn + vy * xn - gi->v3 * gi->r - gj->v3 * gj->r;
  fn = -o->kg * dn - o->nug * vn;
  if (fn < 0) fn = 0.0;
  ft = -o->kt * vt * o->dt;
  real ftest = o->mu * fn;
  if (fabs(ft) > ftest) ft = (ft < 0.0) ? ftest : -ftest;
  F.f1 = fn * xn - ft * yn;
  F.f2 = fn * yn + ft * xn;
  F.f3 = -maxt(ft * gi->r, fn * o->murf * gi->r * gj->r);
  gi->p += fn; gj->p += fn;
  gi->f1 += F.f1; gi->f2 += F.f2;
  gi->s += ft; gj->s += ft;
  gi->slip += fabs(ft) * (fabs(vt * o->dt) + (fabs(ft - o->pft)) / o->kt);
  o->pft = ft;
  gi->rw += fabs(F.f3) * (fabs(gi->v3 * o->dt) + (fabs(F.f3 - o->pff)) / o->kt);
  o->pff = F.f3;
  gi->z += 1; gi->zz += 1;
  gi->ice += o->ic;
  if (fn == 0) gi->ifm = 0;
  else gi->ifm += fabs(ft / (o->mu * fn));
  gi->M11 += F.f1 * xOiOj; gi->M12 += F.f1 * yOiOj;
  gi->M21 += F.f2 * xOiOj; gi->M22 += F.f2 * yOiOj;
  return F;
}

/* the in-lined law of src/main.c:1365-1417, used when nbsteps % stepFilm == 0 */
static force_t pair_force_film(oracle *o, long i, long j) {
  grain_t *gi = &o->g[i], *gj = &o->g[j];
  force_t F;
  real fn, ft, ftest;
  real xOiOj = gi->x1 - gj->x1, yOiOj = gi->x2 - gj->x2;
  real OiOj = sqrt(xOiOj * xOiOj + yOiOj * yOiOj);
  real dn = OiOj - gi->r - gj->r;
  if (dn >= 0) {
    F.f1 = 0; F.f2 = 0; F.f3 = 0;
    return F;
  }
  real vx = gi->v1 - gj->v1, vy = gi->v2 - gj->v2;
  real xn = xOiOj / OiOj, yn = yOiOj / OiOj;
  real vn = vx * xn + vy * yn;
  real vt = -vx * yn + vy * xn - gi->v3 * gi->r - gj->v3 * gj->r;
  fn = -o->kg * dn - o->nug * vn;
  if (fn < 0) fn = 0.0;
  ft = o->kt * vt * o->dt;
  ftest = o->mu * ft;
  if (fabs(ft) > ftest) ft = (ft > 0.0) ? ftest : -ftest;
  F.f1 = fn * xn - ft * yn;
  F.f2 = fn * yn + ft * xn;
  F.f3 = -ft * gi->r * o->murf;
  gi->p += fn; gj->p += fn;
  gi->s += ft; gj->s += ft;
  gi->slip += fabs(ft) * (fabs(vt * o->dt) + (fabs(ft - o->pft)) / o->kt);
  gj->slip += fabs(ft) * (fabs(vt * o->dt) + (fabs(ft - o->pft)) / o->kt);
  gi->rw += fabs(F.f3) * (fabs(gi->v3 * o->dt) + (fabs(F.f3 - o->pff)) / o->kt);
  gj->rw += fabs(F.f3) * (fabs(gi->v3 * o->dt) + (fabs(F.f3 - o->pff)) / o->kt);
  gi->z += 1;
  o->pff = F.f3; o->pft = ft;
  gi->M11 += F.f1 * xOiOj; gi->M12 += F.f1 * yOiOj;
  gi->M21 += F.f2 * xOiOj; gi->M22 += F.f2 * yOiOj;
  return F;
}

/* src/main.c:809-845 */
static force_t wall_bottom(oracle *o, long i, real dn) {
  grain_t *gi = &o->g[i];
  force_t F;
  real vn = gi->v2, vt = gi->v1;
  real fn = -o->km * dn - o->num * vn;
  if (fn < 0) fn = 0.;
  real ft = o->ktm * vt;
  real ftest = o->mumb * fn;
  if (fabs(ft) > ftest) ft = (ft < 0.0) ? ftest : -ftest;
  F.f1 = ft; F.f2 = fn; F.f3 = -(ft * gi->r * o->murf);
  gi->p += fn; gi->s += ft; gi->f1 += F.f1; gi->z += 1;
  gi->M11 += 0; gi->M12 += F.f1 * o->dt; gi->M21 += 0; gi->M22 += F.f2 * o->dt;
  gi->rw += fabs(F.f3) * (fabs(gi->v3 * o->dt) + (fabs(F.f3 - o->pff)) / o->kt);
  gi->fr += fabs(ft) * (fabs(vt * o->dt) + (fabs(ft - o->pft)) / o->kt);
  o->pff = F.f3; o->pft = ft;
  return F;
}

/* src/main.c:846-887 */
static force_t wall_top(oracle *o, long i, real dn) {
  grain_t *gi = &o->g[i];
  force_t F;
  real vn = gi->v2, ftmax;
  real fn = o->km * dn - o->num * vn;
  o->ic += o->num * vn * vn * o->dt;
  if (fn > 0.) fn = 0.;
  real vt = gi->v1 + gi->v3 * gi->r - o->amp * o->freq * cos(o->freq * o->t);
  real ft = fabs(o->ktm * vt);
  if (vt >= 0) ftmax = o->mumb * fn - o->nugt * vt;
  else ftmax = o->mumb * fn + o->nugt * vt;
  if (ft > ftmax) ft = ftmax;
  if (vt > 0) ft = -ft;
  F.f1 = ft; F.f2 = fn; F.f3 = ft * gi->r * o->murf;
  gi->M11 += 0; gi->M12 += F.f1 * fabs(o->dt); gi->M21 += 0; gi->M22 += F.f2 * fabs(o->dt);
  gi->p += fn; gi->s += ft; gi->z += 1;
  return F;
}

/* src/main.c:888-921 */
static force_t wall_left(oracle *o, long i, real dn) {
  grain_t *gi = &o->g[i];
  force_t F;
  real vn = gi->v1;
  real fn = -o->km * dn + o->num * vn;
  o->ic += o->num * vn * vn * o->dt;
  if (fn < 0.) fn = 0.;
  real vt = gi->v2;
  real ft = o->mum * fn; /* both branches of :896-899 are the same expression */
  if (vt > 0) ft = -ft;
  F.f1 = fn; F.f2 = ft; F.f3 = ft * gi->r * o->murf;
  gi->M11 += F.f1 * fabs(o->dt); gi->M12 += 0; gi->M21 += F.f2 * fabs(o->dt); gi->M22 += 0;
  gi->p += fn; gi->s += ft; gi->f1 += F.f1; gi->z += 1;
  gi->ice += o->ic;
  gi->rw += fabs(F.f3) * fabs(gi->v3 * o->dt);
  gi->fr += fabs(ft) * (fabs(vt * o->dt) + (fabs(ft - o->pft)) / o->kt);
  o->pft = ft;
  return F;
}

/* src/main.c:923-951 */
static force_t wall_right(oracle *o, long i, real dn) {
  grain_t *gi = &o->g[i];
  force_t F;
  real vn = gi->v1;
  real fn = o->km * dn - o->num * vn;
  real vt = gi->v2;
  real ft = o->mum * fn;
  if (vt > 0) ft = -ft;
  if (fn > 0.) fn = 0.;
  F.f1 = fn; F.f2 = -ft; F.f3 = ft * gi->r * o->murf;
  gi->p += fn; gi->f1 += F.f1;
  o->pft = ft;
  gi->M11 += F.f1 * fabs(o->dt); gi->M12 += 0; gi->M21 += F.f2 * fabs(o->dt); gi->M22 += 0;
  gi->z += 1;
  return F;
}

/* src/main.c:1336-1516 */
static void acceleration_grains(oracle *o) {
  grain_t *g = o->g;
  const int film = (o->nbsteps % o->stepFilm == 0);
  for (long i = 0; i <= o->n - 1; i++) { g[i].a1 = o->fhf1[i]; g[i].a2 = o->fhf2[i]; g[i].a3 = o->fhf3[i]; }
  for (long i = 0; i <= o->n - 1; i++) {
    int jdep = (i == 0) ? 0 : o->cumul[i - 1];
    for (long k = jdep; k < o->cumul[i]; k++) {
      long j = o->neigh[k];
      force_t F = film ? pair_force_film(o, i, j) : pair_force(o, i, j);
      g[i].a1 = g[i].a1 + F.f1; g[i].a2 = g[i].a2 + F.f2; g[i].a3 = g[i].a3 + F.f3;
      g[j].a1 = g[j].a1 - F.f1; g[j].a2 = g[j].a2 - F.f2; g[j].a3 = g[j].a3 + F.f3;
    }
  }
  /* :1455-1468 -- note g[k] (list position) inside the friction-work diagnostic */
  for (long k = 0; k < o->nB; k++) {
    long i = o->wallB[k];
    real dn = g[i].x2 - g[i].r - o->Mby;
    if (dn < 0) {
      force_t F = wall_bottom(o, i, dn);
      g[i].a1 = g[i].a1 + F.f1; g[i].a2 = g[i].a2 + F.f2; g[i].a3 = g[i].a3 + F.f3;
      g[i].fr += fabs(F.f1) * (fabs(o->dt * g[k].v1) + fabs(o->dt2 * g[k].a1) + (fabs(F.f1 - o->pf)) / o->kt);
      o->pf = F.f1;
    }
  }
  for (long k = 0; k < o->nT; k++) { /* :1472-1480 */
    long i = o->wallT[k];
    real dn = -g[i].x2 - g[i].r + o->Mhy;
    if (dn < 0) {
      force_t F = wall_top(o, i, dn);
      g[i].a1 = g[i].a1 + F.f1; g[i].a2 = g[i].a2 + F.f2; g[i].a3 = g[i].a3 + F.f3;
    }
  }
  for (long k = 0; k < o->nL; k++) { /* :1483-1496 */
    long i = o->wallL[k];
    real dn = g[i].x1 - g[i].r - o->Mgx;
    if (dn < 0) {
      force_t F = wall_left(o, i, dn);
      g[i].a1 = g[i].a1 + F.f1; g[i].a2 = g[i].a2 + F.f2; g[i].a3 = g[i].a3 + F.f3;
      g[i].fr += fabs(F.f2) * (fabs(o->dt * g[k].v1) + fabs(o->dt2 * g[k].a1) + (fabs(F.f2 - o->pf)) / o->kt);
      o->pf = F.f2;
    }
  }
  for (long k = 0; k < o->nR; k++) { /* :1500-1508 */
    long i = o->wallR[k];
    real dn = -g[i].x1 - g[i].r + o->Mdx;
    if (dn < 0) {
      force_t F = wall_right(o, i, dn);
      g[i].a1 = g[i].a1 + F.f1; g[i].a2 = g[i].a2 + F.f2; g[i].a3 = g[i].a3 + F.f3;
    }
  }
  for (long i = 0; i <= o->n - 1; i++) { /* :1511-1515 */
    g[i].a1 = g[i].a1 / g[i].m + ((g[i].m - g[i].mw) / g[i].m) * o->xG;
    g[i].a2 = (g[i].a2 / g[i].m) + ((g[i].m - g[i].mw) / g[i].m) * o->yG;
    g[i].a3 = g[i].a3 / g[i].It;
  }
}

/* src/main.c:1519-1594 */
API void SFX(oracle_init_verlet)(oracle *o) {
  const grain_t *g = o->g;
  int cnt = 0;
  const size_t cap = (size_t)o->n * 64 + 64;
  for (int i = 0; i < o->n; i++)
    for (int j = i + 1; j < o->n; j++) {
      real distx = g[i].x1 - g[j].x1, disty = g[i].x2 - g[j].x2;
      if (((fabs(distx) - g[i].r - g[j].r) <= o->distVerlet) && ((fabs(disty) - g[i].r - g[j].r) <= o->distVerlet))
        if ((sqrt(distx * distx + disty * disty) - g[i].r - g[j].r) <= o->distVerlet)
          if ((size_t)cnt < cap) o->neigh[cnt++] = j;
      o->cumul[i] = cnt;
    }
  o->nB = o->nL = o->nT = o->nR = 0;
  if (o->nbsteps * o->dt < o->dtt) {
    o->Mdx = 1.e-3 * o->lx / 10;
    o->Mhy = (1.e-3 * o->ly / 10);
  } else {
    o->Mdx = 1.e-3 * o->lx;
    o->Mhy = 1.e-3 * o->ly;
  }
  for (int i = 0; i < o->n; ++i) if (g[i].x2 - g[i].r - o->Mby < o->distVerlet) o->wallB[o->nB++] = i;
  for (int i = 0; i < o->n; ++i) if (-g[i].x2 - g[i].r + o->Mhy < o->distVerlet) o->wallT[o->nT++] = i;
  for (int i = 0; i < o->n; ++i) if (g[i].x1 - g[i].r - o->Mgx < o->distVerlet) o->wallL[o->nL++] = i;
  for (int i = 0; i < o->n; ++i) if (-g[i].x1 - g[i].r + o->Mdx < o->distVerlet) o->wallR[o->nR++] = i;
}

/* src/main.c:1697-1765 (outputs are the caller's business) */
static void render_scene(oracle *o) {
  grain_t *g = o->g;
  if (o->vib == 1) { /* :1701-1706 */
    o->t = o->t + o->dt;
    o->Mgx = o->Mgx + o->amp * sin(o->freq * o->t);
    o->Mdx = o->Mdx + o->amp * sin(o->freq * o->t);
  }
  if (o->nbsteps % o->npDEM == 0) SFX(oracle_lbm_step)(o);
  if (o->nbsteps % o->UpdateVerlet == 0) SFX(oracle_init_verlet)(o);
  for (long i = 0; i <= o->n - 1; i++) {
    g[i].p = 0; g[i].s = 0.; g[i].ifm = 0; g[i].f1 = 0.; g[i].f2 = 0.; g[i].ice = 0; g[i].fr = 0.;
    g[i].slip = 0; g[i].rw = 0.; o->ic = 0.;
    g[i].M11 = g[i].M12 = g[i].M21 = g[i].M22 = 0.;
    g[i].z = 0; g[i].zz = 0;
    g[i].x1 = g[i].x1 + o->dt * g[i].v1 + o->dt2 * g[i].a1 / 2.;
    g[i].x2 = g[i].x2 + o->dt * g[i].v2 + o->dt2 * g[i].a2 / 2.;
    g[i].x3 = g[i].x3 + o->dt * g[i].v3 + o->dt2 * g[i].a3 / 2.;
    g[i].v1 = g[i].v1 + o->dt * g[i].a1 / 2.;
    g[i].v2 = g[i].v2 + o->dt * g[i].a2 / 2.;
    g[i].v3 = g[i].v3 + o->dt * g[i].a3 / 2.;
  }
  acceleration_grains(o);
  for (long i = 0; i <= o->n - 1; i++) {
    g[i].v1 = g[i].v1 + o->dt * g[i].a1 / 2.;
    g[i].v2 = g[i].v2 + o->dt * g[i].a2 / 2.;
    g[i].v3 = g[i].v3 + o->dt * g[i].a3 / 2.;
  }
  o->nbsteps++;
  if (o->nbsteps % o->stepFilm == 0) o->nFile++;
}

API void SFX(oracle_step)(oracle *o, long n) { for (long k = 0; k < n; ++k) render_scene(o); }

/* ---------------------------------------------------------------------------------------------
 * accessors
 * ------------------------------------------------------------------------------------------ */
API void SFX(oracle_get_scalars)(oracle *o, double *d, long *l) {
  d[0] = o->dx; d[1] = o->dtLB; d[2] = o->dt; d[3] = o->dt2; d[4] = o->c; d[5] = o->Mgx; d[6] = o->Mdx;
  d[7] = o->Mby; d[8] = o->Mhy; d[9] = o->xG; d[10] = o->yG;
  l[0] = o->npDEM; l[1] = o->nbsteps; l[2] = o->nFile; l[3] = o->n;
}
API void SFX(oracle_set_nbsteps)(oracle *o, long n) { o->nbsteps = n; }
API void SFX(oracle_get_f)(oracle *o, double *out) {
  for (size_t k = 0; k < (size_t)o->lx * o->ly * NQ; ++k) out[k] = o->f[k];
}
API void SFX(oracle_set_f)(oracle *o, const double *in) {
  for (size_t k = 0; k < (size_t)o->lx * o->ly * NQ; ++k) o->f[k] = (real)in[k];
}
API void SFX(oracle_get_delta)(oracle *o, double *out) {
  for (size_t k = 0; k < (size_t)o->lx * o->ly * NQ; ++k) out[k] = o->delta[k];
}
API void SFX(oracle_get_obst)(oracle *o, int *out) { memcpy(out, o->obst, sizeof(int) * o->lx * o->ly); }
API void SFX(oracle_set_obst)(oracle *o, const int *in) { memcpy(o->obst, in, sizeof(int) * o->lx * o->ly); }
API void SFX(oracle_get_act)(oracle *o, int *out) { memcpy(out, o->act, sizeof(int) * o->lx * o->ly); }
API void SFX(oracle_get_grains)(oracle *o, double *out) {
  for (int i = 0; i < o->n; ++i) {
    const grain_t *p = &o->g[i];
    double *q = out + (size_t)i * 13;
    q[0] = p->x1; q[1] = p->x2; q[2] = p->x3; q[3] = p->v1; q[4] = p->v2; q[5] = p->v3;
    q[6] = p->a1; q[7] = p->a2; q[8] = p->a3; q[9] = p->r; q[10] = p->m; q[11] = p->It;
    q[12] = o->rLB[i];
  }
}
API void SFX(oracle_set_grain_state)(oracle *o, const double *in) {
  for (int i = 0; i < o->n; ++i) {
    grain_t *p = &o->g[i];
    const double *q = in + (size_t)i * 9;
    p->x1 = (real)q[0]; p->x2 = (real)q[1]; p->x3 = (real)q[2]; p->v1 = (real)q[3]; p->v2 = (real)q[4];
    p->v3 = (real)q[5]; p->a1 = (real)q[6]; p->a2 = (real)q[7]; p->a3 = (real)q[8];
  }
}
API void SFX(oracle_get_grain_diag)(oracle *o, double *out) {
  for (int i = 0; i < o->n; ++i) {
    const grain_t *p = &o->g[i];
    double *q = out + (size_t)i * 17;
    q[0] = p->p; q[1] = p->s; q[2] = p->f1; q[3] = p->f2; q[4] = p->ifm; q[5] = p->fm; q[6] = p->fr;
    q[7] = p->ifr; q[8] = p->M11; q[9] = p->M12; q[10] = p->M21; q[11] = p->M22; q[12] = p->ice;
    q[13] = p->slip; q[14] = p->rw; q[15] = p->z; q[16] = p->zz;
  }
}
API void SFX(oracle_get_fhf)(oracle *o, double *out) {
  for (int i = 0; i < o->n; ++i) {
    out[3 * (size_t)i] = o->fhf1[i]; out[3 * (size_t)i + 1] = o->fhf2[i]; out[3 * (size_t)i + 2] = o->fhf3[i];
  }
}
API void SFX(oracle_set_fhf)(oracle *o, const double *in) {
  for (int i = 0; i < o->n; ++i) {
    o->fhf1[i] = (real)in[3 * (size_t)i]; o->fhf2[i] = (real)in[3 * (size_t)i + 1];
    o->fhf3[i] = (real)in[3 * (size_t)i + 2];
  }
}
API int SFX(oracle_get_verlet)(oracle *o, int *cumul, int *neigh, int cap) {
  int total = (o->n >= 2) ? o->cumul[o->n - 2] : 0;
  memcpy(cumul, o->cumul, sizeof(int) * o->n);
  if (total > cap) return -total;
  memcpy(neigh, o->neigh, sizeof(int) * total);
  return total;
}
API void SFX(oracle_get_wall_lists)(oracle *o, int *counts, int *b, int *t, int *l, int *r) {
  counts[0] = o->nB; counts[1] = o->nT; counts[2] = o->nL; counts[3] = o->nR;
  memcpy(b, o->wallB, sizeof(int) * o->nB); memcpy(t, o->wallT, sizeof(int) * o->nT);
  memcpy(l, o->wallL, sizeof(int) * o->nL); memcpy(r, o->wallR, sizeof(int) * o->nR);
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
API double SFX(oracle_time_coupled)(oracle *o, long n_dem_steps, long *n_lbm_steps) {
  long lbm = 0;
  double t0 = now_s();
  for (long k = 0; k < n_dem_steps; ++k) {
    if (o->nbsteps % o->npDEM == 0) ++lbm;
    render_scene(o);
  }
  *n_lbm_steps = lbm;
  return now_s() - t0;
}
API double SFX(oracle_time_lbm)(oracle *o, long n_lbm_steps) {
  double t0 = now_s();
  for (long k = 0; k < n_lbm_steps; ++k) SFX(oracle_lbm_step)(o);
  return now_s() - t0;
}
