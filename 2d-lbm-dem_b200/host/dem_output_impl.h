/* dem_output_impl.h -- included twice by dem_output.c, with `real` = double and = float, so that
 * the replay keeps the promotions of the reference's two builds.  Expressions follow
 * src/main.c:729-951 and :1426-1516 operand by operand (diagnostics only: no force is fed back). */

typedef struct {
  real f1, f2, f3;
} FN(force);

typedef struct {
  int n;
  /* constants, as `real` globals of the reference (src/main.c:97-118) */
  real kg, kt, km, ktm, nug, num, numb, nugt, mu, mum, mumb, murf, freq, amp, t, G, dtt;
  real dt, dt2, Mgx, Mdx, Mby, Mhy;
  real pf, pft, pff, ic; /* carried across contacts AND across calls (never reset, :130-131) */
  real TBW, TSE;         /* accumulated by write_DEM (:399-401) */
  real *x1, *x2, *x3, *v1, *v2, *v3, *r, *a1, *a2, *a3;
  real *p, *s, *f1, *f2, *ifm, *fm, *fr, *ifr, *M11, *M12, *M21, *M22, *ice, *slip, *rw;
  int *z, *zz;
} FN(state);

static real FN(maxt)(real x, real y) { /* :211-216 */
  if (x < y) return 0.;
  return y;
}

static FN(force) FN(force_grains)(FN(state) *S, long i, long j) { /* :729-803 */
  real dn, xOiOj, yOiOj, OiOj, xn, yn, vn, vxOiOj, vyOiOj, vt, ftest;
  FN(force) f;
  double fn, ft;
  xOiOj = S->x1[i] - S->x1[j];
  yOiOj = S->x2[i] - S->x2[j];
  OiOj = sqrt(xOiOj * xOiOj + yOiOj * yOiOj);
  dn = OiOj - S->r[i] - S->r[j];
  if (dn >= 0) {
    f.f1 = 0; f.f2 = 0; f.f3 = 0;
  } else {
    vxOiOj = S->v1[i] - S->v1[j];
    vyOiOj = S->v2[i] - S->v2[j];
    xn = xOiOj / OiOj;
    yn = yOiOj / OiOj;
    vn = vxOiOj * xn + vyOiOj * yn;
    vt = -vxOiOj * yn + vyOiOj * xn - S->v3[i] * S->r[i] - S->v3[j] * S->r[j];
    fn = -S->kg * dn - S->nug * vn;
    if (fn < 0) fn = 0.0;
    ft = -S->kt * vt * S->dt;
    ftest = S->mu * fn;
    if (fabs(ft) > ftest) {
      if (ft < 0.0) ft = ftest;
      else ft = -ftest;
    }
    f.f1 = fn * xn - ft * yn;
    f.f2 = fn * yn + ft * xn;
    f.f3 = -FN(maxt)(ft * S->r[i], fn * S->murf * S->r[i] * S->r[j]);
    S->p[i] += fn;
    S->p[j] += fn;
    S->f1[i] += f.f1;
    S->f2[i] += f.f2;
    S->s[i] += ft;
    S->s[j] += ft;
    S->slip[i] += fabs(ft) * (fabs(vt * S->dt) + (fabs(ft - S->pft)) / S->kt);
    S->pft = ft;
    S->rw[i] += fabs(f.f3) * (fabs(S->v3[i] * S->dt) + (fabs(f.f3 - S->pff)) / S->kt);
    S->pff = f.f3;
    S->z[i] += 1;
    S->zz[i] += 1;
    S->ice[i] += S->ic;
    if (fn == 0) S->ifm[i] = 0;
    else S->ifm[i] += fabs(ft / (S->mu * fn));
    S->M11[i] += f.f1 * xOiOj;
    S->M12[i] += f.f1 * yOiOj;
    S->M21[i] += f.f2 * xOiOj;
    S->M22[i] += f.f2 * yOiOj;
  }
  return f;
}

static FN(force) FN(force_WallB)(FN(state) *S, long i, real dn) { /* :809-845 */
  real vn, vt, ftest, fn, ft;
  FN(force) f;
  vn = S->v2[i];
  vt = S->v1[i];
  fn = -S->km * dn - S->num * vn;
  if (fn < 0) fn = 0.;
  ft = S->ktm * vt;
  ftest = S->mumb * fn;
  if (fabs(ft) > ftest) {
    if (ft < 0.0) ft = ftest;
    else ft = -ftest;
  }
  f.f1 = ft;
  f.f2 = fn;
  f.f3 = -(ft * S->r[i] * S->murf);
  S->p[i] += fn;
  S->s[i] += ft;
  S->f1[i] += f.f1;
  S->z[i] += 1;
  S->M11[i] += 0;
  S->M12[i] += f.f1 * S->dt;
  S->M21[i] += 0;
  S->M22[i] += f.f2 * S->dt;
  S->rw[i] += fabs(f.f3) * (fabs(S->v3[i] * S->dt) + (fabs(f.f3 - S->pff)) / S->kt);
  S->fr[i] += fabs(ft) * (fabs(vt * S->dt) + (fabs(ft - S->pft)) / S->kt);
  S->pff = f.f3;
  S->pft = ft;
  return f;
}

static FN(force) FN(force_WallT)(FN(state) *S, long i, real dn) { /* :846-887 */
  real vn, vt, fn, ft, ftmax;
  FN(force) f;
  vn = S->v2[i];
  fn = S->km * dn - S->num * vn;
  S->ic += S->num * vn * vn * S->dt;
  if (fn > 0.) fn = 0.;
  vt = S->v1[i] + S->v3[i] * S->r[i] - S->amp * S->freq * cos(S->freq * S->t);
  ft = fabs(S->ktm * vt);
  if (vt >= 0) ftmax = S->mumb * fn - S->nugt * vt;
  else ftmax = S->mumb * fn + S->nugt * vt;
  if (ft > ftmax) ft = ftmax;
  if (vt > 0) ft = -ft;
  f.f1 = ft;
  f.f2 = fn;
  f.f3 = ft * S->r[i] * S->murf;
  S->M11[i] += 0;
  S->M12[i] += f.f1 * fabs(S->dt);
  S->M21[i] += 0;
  S->M22[i] += f.f2 * fabs(S->dt);
  S->p[i] += fn;
  S->s[i] += ft;
  S->z[i] += 1;
  return f;
}

static FN(force) FN(force_WallL)(FN(state) *S, long i, real dn) { /* :888-921 */
  real vn, fn, vt, ft;
  FN(force) f;
  vn = S->v1[i];
  fn = -S->km * dn + S->num * vn;
  S->ic += S->num * vn * vn * S->dt;
  if (fn < 0.) fn = 0.;
  vt = S->v2[i];
  if (vt > 0) ft = S->mum * fn;
  else ft = S->mum * fn;
  if (vt > 0) ft = -ft;
  f.f1 = fn;
  f.f2 = ft;
  f.f3 = ft * S->r[i] * S->murf;
  S->M11[i] += f.f1 * fabs(S->dt);
  S->M12[i] += 0;
  S->M21[i] += f.f2 * fabs(S->dt);
  S->M22[i] += 0;
  S->p[i] += fn;
  S->s[i] += ft;
  S->f1[i] += f.f1;
  S->z[i] += 1;
  S->ice[i] += S->ic;
  S->rw[i] += fabs(f.f3) * fabs(S->v3[i] * S->dt);
  S->fr[i] += fabs(ft) * (fabs(vt * S->dt) + (fabs(ft - S->pft)) / S->kt);
  S->pft = ft;
  return f;
}

static FN(force) FN(force_WallR)(FN(state) *S, long i, real dn) { /* :923-951 */
  real vn, fn, vt, ft;
  FN(force) f;
  vn = S->v1[i];
  fn = S->km * dn - S->num * vn;
  vt = S->v2[i];
  ft = S->mum * fn;
  if (vt > 0) ft = -ft;
  if (fn > 0.) fn = 0.;
  f.f1 = fn;
  f.f2 = -ft;
  f.f3 = ft * S->r[i] * S->murf;
  S->p[i] += fn;
  S->f1[i] += f.f1;
  S->pft = ft;
  S->M11[i] += f.f1 * fabs(S->dt);
  S->M12[i] += 0;
  S->M21[i] += f.f2 * fabs(S->dt);
  S->M22[i] += 0;
  S->z[i] += 1;
  return f;
}

static void FN(add)(FN(state) *S, long k, FN(force) f) {
  S->a1[k] = S->a1[k] + f.f1;
  S->a2[k] = S->a2[k] + f.f2;
  S->a3[k] = S->a3[k] + f.f3;
}

/* the reset of :1733-1746, then acceleration_grains :1426-1508 (force sums are kept only because
 * two wall diagnostics read a partially summed a1) */
static void FN(pass)(FN(state) *S, const double *fhf, const int *count, const int *nbr, int cap, const int *wflags) {
  const int n = S->n;
  for (int i = 0; i < n; ++i) {
    S->p[i] = 0; S->s[i] = 0.; S->ifm[i] = 0; S->f1[i] = 0.; S->f2[i] = 0.; S->ice[i] = 0; S->fr[i] = 0.;
    S->slip[i] = 0; S->rw[i] = 0.;
    S->ic = 0.;
    S->M11[i] = S->M12[i] = S->M21[i] = S->M22[i] = 0.;
    S->z[i] = 0; S->zz[i] = 0;
    S->a1[i] = (real)fhf[3 * i]; S->a2[i] = (real)fhf[3 * i + 1]; S->a3[i] = (real)fhf[3 * i + 2];
  }
  /* half list of the reference = the j > i part of the device's sorted full lists */
  for (long i = 0; i < n; ++i)
    for (int k = 0; k < count[i]; ++k) {
      const long j = nbr[(size_t)i * cap + k];
      if (j <= i) continue;
      FN(force) fji = FN(force_grains)(S, i, j);
      S->a1[i] = S->a1[i] + fji.f1; S->a2[i] = S->a2[i] + fji.f2; S->a3[i] = S->a3[i] + fji.f3;
      S->a1[j] = S->a1[j] - fji.f1; S->a2[j] = S->a2[j] - fji.f2; S->a3[j] = S->a3[j] + fji.f3;
    }
  /* wall loops: `i` below is the POSITION in the wall list, and the reference reads g[i].v1, g[i].a1
   * with it (:1462-1466, :1490-1494) */
  long i = 0;
  for (long w = 0; w < n; ++w) { /* bottom */
    if (!(wflags[w] & 1)) continue;
    const real dn = S->x2[w] - S->r[w] - S->Mby;
    if (dn < 0) {
      FN(force) fji = FN(force_WallB)(S, w, dn);
      FN(add)(S, w, fji);
      S->fr[w] += fabs(fji.f1) * (fabs(S->dt * S->v1[i]) + fabs(S->dt2 * S->a1[i]) + (fabs(fji.f1 - S->pf)) / S->kt);
      S->pf = fji.f1;
    }
    ++i;
  }
  for (long w = 0; w < n; ++w) { /* top */
    if (!(wflags[w] & 2)) continue;
    const real dn = -S->x2[w] - S->r[w] + S->Mhy;
    if (dn < 0) FN(add)(S, w, FN(force_WallT)(S, w, dn));
  }
  i = 0;
  for (long w = 0; w < n; ++w) { /* left */
    if (!(wflags[w] & 4)) continue;
    const real dn = S->x1[w] - S->r[w] - S->Mgx;
    if (dn < 0) {
      FN(force) fji = FN(force_WallL)(S, w, dn);
      FN(add)(S, w, fji);
      S->fr[w] += fabs(fji.f2) * (fabs(S->dt * S->v1[i]) + fabs(S->dt2 * S->a1[i]) + (fabs(fji.f2 - S->pf)) / S->kt);
      S->pf = fji.f2;
    }
    ++i;
  }
  for (long w = 0; w < n; ++w) { /* right */
    if (!(wflags[w] & 8)) continue;
    const real dn = -S->x1[w] - S->r[w] + S->Mdx;
    if (dn < 0) FN(add)(S, w, FN(force_WallR)(S, w, dn));
  }
}

/* write_DEM :340-438 */
static int FN(write_dem)(FN(state) *S, const char *dir, int nfile, long nbsteps, const double *grains, const double *fhf,
                         double *summary) {
  const int n = S->n;
  char path[1024];
  snprintf(path, sizeof path, "%s%sDEM%.6i.dat", dir ? dir : "", (dir && *dir) ? "/" : "", nfile);
  FILE *out = fopen(path, "w");
  if (!out) return -1;
#define GR(i, c) ((real)grains[(size_t)(i) * 13 + (c)]) /* x1 x2 x3 v1 v2 v3 a1 a2 a3 r m It rLB */
  real xfront = GR(0, 0) + GR(0, 9), height = GR(0, 1) + GR(0, 9);
  real energie_cin = 0., energie_x = 0., energie_y = 0., energie_teta = 0., energy_p = 0.;
  real SE = 0., ESE = 0., WF = 0., IFR = 0., INCE = 0., TSLIP = 0., TRW = 0., zmean = 0;
  real xgrainmax = GR(0, 0);
  real N0 = 0, N1 = 0, N2 = 0, N3 = 0, N4 = 0, N5 = 0;
  for (int i = 0; i < n; i++) {
    const real x1 = GR(i, 0), x2 = GR(i, 1), x3 = GR(i, 2), v1 = GR(i, 3), v2 = GR(i, 4), v3 = GR(i, 5);
    const real a1 = GR(i, 6), a2 = GR(i, 7), a3 = GR(i, 8), r = GR(i, 9), m = GR(i, 10), It = GR(i, 11);
    zmean += S->z[i];
    if (S->z[i] == 0) N0 += 1;
    if (S->z[i] == 1) N1 += 1;
    if (S->z[i] == 2) N2 += 1;
    if (S->z[i] == 3) N3 += 1;
    if (S->z[i] == 4) N4 += 1;
    if (S->z[i] == 5) N5 += 1;
    energie_x += 0.5 * m * v1 * v1;
    energie_y += 0.5 * m * v2 * v2;
    energie_teta += 0.5 * It * v3 * v3;
    energy_p += m * S->G * x2;
    SE += 0.5 * (((S->p[i] * S->p[i]) / S->kg) + ((S->s[i] * S->s[i]) / S->kt));
    WF += S->fr[i];
    S->ifr[i] = fabs(((m * S->G + S->f2[i]) * (S->dt * v2 + S->dt2 * a2 / 2.)) + (S->f1[i] * (S->dt * v1 + S->dt2 * a1 / 2.)));
    IFR += S->ifr[i];
    TSLIP += S->slip[i];
    TRW += S->rw[i];
    INCE += S->ice[i];
    S->TBW += S->ifr[i];
    ESE = 0.5 * (((S->p[i] * S->p[i]) / S->kg) + ((S->s[i] * S->s[i]) / S->kt));
    S->TSE += ESE;
    if (x1 + r > xgrainmax) xgrainmax = x1 + r;
    if (x2 + r > height) height = x2 + r;
    if (S->zz[i] > 0 && x1 + r >= xfront) xfront = x1 + r;
    if (S->z[i] == 0) S->fm[i] = 0;
    else S->fm[i] = S->ifm[i] / S->z[i];
    fprintf(out,
            "%i\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%"
            "le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%le\t%"
            "le\t%i\n",
            i, (double)r, (double)x1, (double)x2, (double)x3, (double)v1, (double)v2, (double)v3, (double)a1, (double)a2,
            (double)a3, (double)(real)fhf[3 * i], (double)(real)fhf[3 * i + 1], (double)(real)fhf[3 * i + 2], (double)S->p[i],
            (double)S->s[i], (double)ESE, (double)S->fr[i], (double)S->ifr[i], (double)S->ice[i], (double)S->slip[i],
            (double)S->rw[i], (double)S->fm[i], (double)S->M11[i], (double)S->M12[i], (double)S->M21[i], (double)S->M22[i],
            S->z[i]);
  }
#undef GR
  energie_cin = energie_x + energie_y + energie_teta;
  zmean = zmean / n;
  snprintf(path, sizeof path, "%s%sstats.data", dir ? dir : "", (dir && *dir) ? "/" : "");
  FILE *st = fopen(path, "a");
  if (!st) { fclose(out); return -1; }
  fprintf(st,
          "%le %le %le %le %le %le %le %le %le %le %le %le %le %le %le %le %le "
          "%le %le %le %le %le\n",
          (double)(nbsteps * S->dt - S->dtt), (double)xfront, (double)xgrainmax, (double)height, (double)zmean,
          (double)energie_x, (double)energie_y, (double)energie_teta, (double)energie_cin, (double)(N0 / n),
          (double)(N1 / n), (double)(N2 / n), (double)(N3 / n), (double)(N4 / n), (double)(N5 / n), (double)energy_p,
          (double)SE, (double)WF, (double)IFR, (double)INCE, (double)TSLIP, (double)TRW);
  fclose(st);
  fclose(out);
  if (summary) {
    summary[0] = energie_cin; summary[1] = energy_p; summary[2] = SE; summary[3] = WF; summary[4] = INCE;
    summary[5] = TSLIP; summary[6] = TRW;
  }
  return 0;
}
