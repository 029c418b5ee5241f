/*
 * dem_node.cuh -- contact laws and integrator of the DEM sub-step, written once for host and
 * device.  Restates, expression by expression (operand order, int/float/double promotions):
 *   force_grains           src/main.c:729-803     (pair law)
 *   in-lined "film" law    src/main.c:1365-1417   (steps with nbsteps % stepFilm == 0)
 *   force_WallB/T/L/R      src/main.c:809-951
 *   a = F/m + g            src/main.c:1511-1515
 *   velocity-Verlet halves src/main.c:1748-1753, :1758-1763
 *   Verlet-list criteria   src/main.c:1525-1532, :1563-1593
 * Only the trajectory-relevant outputs are produced here (force and torque); the per-contact
 * diagnostics the reference accumulates alongside (p, s, slip, rw, ...) do not feed back into
 * the motion.
 *
 * Compiled without multiply-add contraction (nvcc -fmad=false) so that, given the same inputs
 * and the same summation order, one DEM step is bit-identical to the reference build
 * (gcc -O2 -ffp-contract=off).
 */
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DEM_HD __host__ __device__ __forceinline__
#else
#define DEM_HD inline
#endif

namespace dem {

template <typename real>
struct Params {
  /* src/main.c:97-118 */
  real kg, kt, km, ktm, nug, num, numb, nugt, mu, mum, mumb, murf;
  real freq, amp, t, distVerlet;
  real dt, dt2, xG, yG;
  real Mgx, Mdx, Mby, Mhy;
};

template <typename real>
struct Force { real f1, f2, f3; };

/* src/main.c:211-216 */
template <typename real>
DEM_HD real maxt(real x, real y) { return (x < y) ? (real)0. : y; }

/* One contact of the ordered pair (i, j), i < j.  Returns false when the grains do not touch.
 * The force is the one applied to grain i; grain j receives (-f1, -f2, +f3)  (:1442-1448). */
template <typename real>
DEM_HD bool pair_force(const Params<real> &P, bool film, real xi1, real xi2, real vi1, real vi2, real vi3, real ri,
                       real xj1, real xj2, real vj1, real vj2, real vj3, real rj, Force<real> *F) {
  const real xOiOj = xi1 - xj1, yOiOj = xi2 - xj2;
  const real OiOj = sqrt((double)(xOiOj * xOiOj + yOiOj * yOiOj));
  const real dn = OiOj - ri - rj;
  if (dn >= 0) return false;
  const real vx = vi1 - vj1, vy = vi2 - vj2;
  const real xn = xOiOj / OiOj, yn = yOiOj / OiOj;
  const real vn = vx * xn + vy * yn;
  const real vt = -vx * yn + vy * xn - vi3 * ri - vj3 * rj;
  if (!film) {
    /* force_grains declares `double fn, ft` (:736) */
    double fn = -P.kg * dn - P.nug * vn;
    if (fn < 0) fn = 0.0;
    double ft = -P.kt * vt * P.dt;
    const real ftest = P.mu * fn;
    if (fabs(ft) > ftest) ft = (ft < 0.0) ? ftest : -ftest;
    F->f1 = fn * xn - ft * yn;
    F->f2 = fn * yn + ft * xn;
    F->f3 = -maxt<real>(ft * ri, fn * P.murf * ri * rj);
  } else {
    real fn = -P.kg * dn - P.nug * vn;
    if (fn < 0) fn = 0.0;
    real ft = P.kt * vt * P.dt;
    const real ftest = P.mu * ft;
    if (fabs((double)ft) > ftest) ft = (ft > 0.0) ? ftest : -ftest;
    F->f1 = fn * xn - ft * yn;
    F->f2 = fn * yn + ft * xn;
    F->f3 = -ft * ri * P.murf;
  }
  return true;
}

/* src/main.c:809-845 */
template <typename real>
DEM_HD Force<real> wall_bottom(const Params<real> &P, real v1, real v2, real r, real dn) {
  const real vn = v2, vt = v1;
  real fn = -P.km * dn - P.num * vn;
  if (fn < 0) fn = 0.;
  real ft = P.ktm * vt;
  const real ftest = P.mumb * fn;
  if (fabs((double)ft) > ftest) ft = (ft < 0.0) ? ftest : -ftest;
  Force<real> F;
  F.f1 = ft; F.f2 = fn; F.f3 = -(ft * r * P.murf);
  return F;
}

/* src/main.c:846-887 */
template <typename real>
DEM_HD Force<real> wall_top(const Params<real> &P, real v1, real v2, real v3, real r, real dn) {
  const real vn = v2;
  real ftmax;
  real fn = P.km * dn - P.num * vn;
  if (fn > 0.) fn = 0.;
  const real vt = v1 + v3 * r - P.amp * P.freq * cos((double)(P.freq * P.t));
  real ft = fabs((double)(P.ktm * vt));
  if (vt >= 0) ftmax = P.mumb * fn - P.nugt * vt;
  else ftmax = P.mumb * fn + P.nugt * vt;
  if (ft > ftmax) ft = ftmax;
  if (vt > 0) ft = -ft;
  Force<real> F;
  F.f1 = ft; F.f2 = fn; F.f3 = ft * r * P.murf;
  return F;
}

/* src/main.c:888-921 (both branches of :896-899 are the same expression) */
template <typename real>
DEM_HD Force<real> wall_left(const Params<real> &P, real v1, real v2, real r, real dn) {
  const real vn = v1;
  real fn = -P.km * dn + P.num * vn;
  if (fn < 0.) fn = 0.;
  const real vt = v2;
  real ft = P.mum * fn;
  if (vt > 0) ft = -ft;
  Force<real> F;
  F.f1 = fn; F.f2 = ft; F.f3 = ft * r * P.murf;
  return F;
}

/* src/main.c:923-951 (ft is taken from fn BEFORE fn is clipped) */
template <typename real>
DEM_HD Force<real> wall_right(const Params<real> &P, real v1, real v2, real r, real dn) {
  const real vn = v1;
  real fn = P.km * dn - P.num * vn;
  const real vt = v2;
  real ft = P.mum * fn;
  if (vt > 0) ft = -ft;
  if (fn > 0.) fn = 0.;
  Force<real> F;
  F.f1 = fn; F.f2 = -ft; F.f3 = ft * r * P.murf;
  return F;
}

/* src/main.c:1529-1532: is the pair in the Verlet list?  (symmetric in i, j) */
template <typename real>
DEM_HD bool verlet_pair(const Params<real> &P, real xi1, real xi2, real ri, real xj1, real xj2, real rj) {
  const real distx = xi1 - xj1, disty = xi2 - xj2;
  if (((fabs((double)distx) - ri - rj) <= P.distVerlet) && ((fabs((double)disty) - ri - rj) <= P.distVerlet))
    if ((sqrt((double)(distx * distx + disty * disty)) - ri - rj) <= P.distVerlet) return true;
  return false;
}

/* src/main.c:1563-1593: bit k set when the grain is in wall list k (0 B, 1 T, 2 L, 3 R) */
template <typename real>
DEM_HD int wall_flags(const Params<real> &P, real x1, real x2, real r) {
  int fl = 0;
  real dn = x2 - r - P.Mby;
  if (dn < P.distVerlet) fl |= 1;
  dn = -x2 - r + P.Mhy;
  if (dn < P.distVerlet) fl |= 2;
  dn = x1 - r - P.Mgx;
  if (dn < P.distVerlet) fl |= 4;
  dn = -x1 - r + P.Mdx;
  if (dn < P.distVerlet) fl |= 8;
  return fl;
}

/* the four wall loops of acceleration_grains (:1455-1508) for one grain, added onto A in the
 * reference's order B, T, L, R */
template <typename real>
DEM_HD void add_wall_forces(const Params<real> &P, int flags, real x1, real x2, real v1, real v2, real v3, real r,
                            real *a1, real *a2, real *a3) {
  if (flags & 1) {
    const real dn = x2 - r - P.Mby;
    if (dn < 0) {
      const Force<real> F = wall_bottom(P, v1, v2, r, dn);
      *a1 = *a1 + F.f1; *a2 = *a2 + F.f2; *a3 = *a3 + F.f3;
    }
  }
  if (flags & 2) {
    const real dn = -x2 - r + P.Mhy;
    if (dn < 0) {
      const Force<real> F = wall_top(P, v1, v2, v3, r, dn);
      *a1 = *a1 + F.f1; *a2 = *a2 + F.f2; *a3 = *a3 + F.f3;
    }
  }
  if (flags & 4) {
    const real dn = x1 - r - P.Mgx;
    if (dn < 0) {
      const Force<real> F = wall_left(P, v1, v2, r, dn);
      *a1 = *a1 + F.f1; *a2 = *a2 + F.f2; *a3 = *a3 + F.f3;
    }
  }
  if (flags & 8) {
    const real dn = -x1 - r + P.Mdx;
    if (dn < 0) {
      const Force<real> F = wall_right(P, v1, v2, r, dn);
      *a1 = *a1 + F.f1; *a2 = *a2 + F.f2; *a3 = *a3 + F.f3;
    }
  }
}

/* src/main.c:1511-1515, with mw = 0 (never initialised on the read_sample path) */
template <typename real>
DEM_HD void finish_acceleration(const Params<real> &P, real m, real It, real *a1, real *a2, real *a3) {
  const real mw = 0;
  *a1 = *a1 / m + ((m - mw) / m) * P.xG;
  *a2 = (*a2 / m) + ((m - mw) / m) * P.yG;
  *a3 = *a3 / It;
}

/* src/main.c:1748-1753: x <- x + dt v + dt2 a / 2. ; v <- v + dt a / 2. */
template <typename real>
DEM_HD void kick_drift(const Params<real> &P, real *x, real *v, real a) {
  *x = *x + P.dt * *v + P.dt2 * a / 2.;
  *v = *v + P.dt * a / 2.;
}
/* src/main.c:1758-1763 */
template <typename real>
DEM_HD void kick(const Params<real> &P, real *v, real a) { *v = *v + P.dt * a / 2.; }

}  // namespace dem
