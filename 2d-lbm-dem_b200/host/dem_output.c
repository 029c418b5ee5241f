/* dem_output.c -- see dem_output.h.  Plain C99; build with -ffp-contract=off (the replay is meant to
 * give the reference's digits). */
#include "dem_output.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define real double
#define FN(name) name##_f64
#include "dem_output_impl.h"
#undef real
#undef FN
#define real float
#define FN(name) name##_f32
#include "dem_output_impl.h"
#undef real
#undef FN

struct lbmdem_diag {
  int single;
  int n;
  state_f64 d;
  state_f32 s;
};

#define ALLOC_ALL(S, T)                                                                                          \
  do {                                                                                                           \
    T **arr[] = {&S.x1, &S.x2, &S.x3, &S.v1, &S.v2, &S.v3, &S.r, &S.a1, &S.a2, &S.a3, &S.p, &S.s, &S.f1, &S.f2,  \
                 &S.ifm, &S.fm, &S.fr, &S.ifr, &S.M11, &S.M12, &S.M21, &S.M22, &S.ice, &S.slip, &S.rw};          \
    for (size_t k = 0; k < sizeof arr / sizeof arr[0]; ++k) *arr[k] = (T *)calloc((size_t)n, sizeof(T));         \
    S.z = (int *)calloc((size_t)n, sizeof(int));                                                                 \
    S.zz = (int *)calloc((size_t)n, sizeof(int));                                                                \
  } while (0)
#define FREE_ALL(S, T)                                                                                           \
  do {                                                                                                           \
    T *arr[] = {S.x1, S.x2, S.x3, S.v1, S.v2, S.v3, S.r, S.a1, S.a2, S.a3, S.p, S.s, S.f1, S.f2,                  \
                S.ifm, S.fm, S.fr, S.ifr, S.M11, S.M12, S.M21, S.M22, S.ice, S.slip, S.rw};                      \
    for (size_t k = 0; k < sizeof arr / sizeof arr[0]; ++k) free(arr[k]);                                        \
    free(S.z);                                                                                                   \
    free(S.zz);                                                                                                  \
  } while (0)
#define SET_CONST(S, T)                                                                                          \
  do {                                                                                                           \
    S.n = n;                                                                                                     \
    S.kg = (T)p->kg; S.kt = (T)p->kt; S.km = (T)p->km; S.ktm = (T)p->ktm; S.nug = (T)p->nug; S.num = (T)p->num;   \
    S.numb = (T)p->numb; S.nugt = (T)p->nugt; S.mu = (T)p->mu; S.mum = (T)p->mum; S.mumb = (T)p->mumb;            \
    S.murf = (T)p->murf; S.freq = (T)p->freq; S.amp = (T)p->amp; S.t = 0; S.G = (T)p->G; S.dtt = (T)p->dtt;        \
    S.pf = 0; S.pft = 0; S.pff = 0; S.ic = 0; S.TBW = 0; S.TSE = 0;                                               \
  } while (0)

lbmdem_diag *lbmdem_diag_create(int n, const lbmdem_params *p) {
  lbmdem_diag *d = (lbmdem_diag *)calloc(1, sizeof *d);
  if (!d) return NULL;
  d->single = p->single_precision;
  d->n = n;
  if (d->single) { ALLOC_ALL(d->s, float); SET_CONST(d->s, float); }
  else { ALLOC_ALL(d->d, double); SET_CONST(d->d, double); }
  return d;
}
void lbmdem_diag_destroy(lbmdem_diag *d) {
  if (!d) return;
  if (d->single) FREE_ALL(d->s, float);
  else FREE_ALL(d->d, double);
  free(d);
}

#define LOAD(S, T)                                                                                               \
  do {                                                                                                           \
    for (int i = 0; i < d->n; ++i) {                                                                             \
      S.x1[i] = (T)mid[6 * i]; S.x2[i] = (T)mid[6 * i + 1]; S.x3[i] = (T)mid[6 * i + 2];                         \
      S.v1[i] = (T)mid[6 * i + 3]; S.v2[i] = (T)mid[6 * i + 4]; S.v3[i] = (T)mid[6 * i + 5];                     \
      S.r[i] = (T)props[13 * i + 9];                                                                             \
    }                                                                                                            \
    S.dt = (T)d11[2]; S.dt2 = (T)d11[3]; S.Mgx = (T)d11[5]; S.Mdx = (T)d11[6]; S.Mby = (T)d11[7]; S.Mhy = (T)d11[8]; \
  } while (0)

void lbmdem_diag_pass(lbmdem_diag *d, const double *mid, const double *props, const double *fhf, const int *count,
                      const int *nbr, int cap, const int *wflags, const double *d11) {
  if (d->single) { LOAD(d->s, float); pass_f32(&d->s, fhf, count, nbr, cap, wflags); }
  else { LOAD(d->d, double); pass_f64(&d->d, fhf, count, nbr, cap, wflags); }
}

#define STORE(S)                                                                                                 \
  for (int i = 0; i < d->n; ++i) {                                                                               \
    double *o = out + (size_t)i * 17;                                                                            \
    o[0] = S.p[i]; o[1] = S.s[i]; o[2] = S.f1[i]; o[3] = S.f2[i]; o[4] = S.ifm[i]; o[5] = S.fm[i]; o[6] = S.fr[i]; \
    o[7] = S.ifr[i]; o[8] = S.M11[i]; o[9] = S.M12[i]; o[10] = S.M21[i]; o[11] = S.M22[i]; o[12] = S.ice[i];      \
    o[13] = S.slip[i]; o[14] = S.rw[i]; o[15] = S.z[i]; o[16] = S.zz[i];                                          \
  }
void lbmdem_diag_get(const lbmdem_diag *d, double *out) {
  if (d->single) { STORE(d->s) } else { STORE(d->d) }
}

int lbmdem_write_dem(lbmdem_diag *d, const char *dir, int nfile, long nbsteps, const double *grains, const double *fhf,
                     const double *d11, double *summary) {
  if (d->single) {
    d->s.dt = (float)d11[2]; d->s.dt2 = (float)d11[3];
    return write_dem_f32(&d->s, dir, nfile, nbsteps, grains, fhf, summary);
  }
  d->d.dt = d11[2]; d->d.dt2 = d11[3];
  return write_dem_f64(&d->d, dir, nfile, nbsteps, grains, fhf, summary);
}

int lbmdem_write_stats_header(const char *dir) {
  char path[1024];
  snprintf(path, sizeof path, "%s%sstats.data", dir ? dir : "", (dir && *dir) ? "/" : "");
  FILE *fp = fopen(path, "w");
  if (!fp) return -1;
  fprintf(fp,
          "#1_t 2_xfront 3_xgrainmax 4_height 5_zmean 6_energie_x 7_energie_y "
          "8_energie_teta 9_energie_cin 10_N0 11_N1 12_N2 13_N3 14_N4 15_N5 "
          "16_energy_Potential 17_Strain_Energy 18_Frictional_Work "
          "19_Internal_Friction 20_Inelastic_Collision 21_Slip "
          "22_Rotational_Work\n");
  fclose(fp);
  return 0;
}
