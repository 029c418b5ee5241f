/*
 * dem_output.h -- host side of the reference's grain outputs: DEM%06d.dat and stats.data
 * (write_DEM, src/main.c:340-438).
 *
 * Most columns of those files are contact diagnostics (p, s, slip, rw, ice, fr, M11.., z) that the
 * reference accumulates inside force_grains / force_Wall* (src/main.c:729-951) while it loops
 * over the contacts SERIALLY, carrying the globals pf, pft, pff, ic from one contact to the next
 * (SURVEY.md App. B #7).  They never feed back into the motion, so the device kernels do not
 * compute them; on the two renderScene() calls before an output the host driver captures the
 * grain state as acceleration_grains() saw it (lbmdem_step_capture) and replays the contact loop
 * here, in the reference's order, for the diagnostics alone.
 */
#ifndef LBMDEM_DEM_OUTPUT_H
#define LBMDEM_DEM_OUTPUT_H
#include "../../include/lbmdem_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lbmdem_diag lbmdem_diag;

lbmdem_diag *lbmdem_diag_create(int n, const lbmdem_params *p);
void lbmdem_diag_destroy(lbmdem_diag *d);

/* One pass of acceleration_grains() (src/main.c:1426-1516, normal contact law) over the captured
 * state, diagnostics only.
 *   mid   [n][6]  x1 x2 x3 v1 v2 v3 after the kick-drift of this call (lbmdem_step_capture)
 *   props [n][13] rows of lbmdem_get_grains (r, m, It are read)
 *   fhf   [n][3]
 *   count/nbr/cap/wflags  the full neighbour lists and wall flags of lbmdem_get_verlet
 *   d11   dx dtLB dt dt2 c Mgx Mdx Mby Mhy xG yG (lbmdem_get_scalars) */
void lbmdem_diag_pass(lbmdem_diag *d, const double *mid, const double *props, const double *fhf, const int *count,
                      const int *nbr, int cap, const int *wflags, const double *d11);

/* per-grain diagnostics of the last pass: [n][17] p s f1 f2 ifm fm fr ifr M11 M12 M21 M22 ice slip rw z zz */
void lbmdem_diag_get(const lbmdem_diag *d, double *out);

/* write_DEM (src/main.c:340-438): <dir>/DEM%06d.dat and one row appended to <dir>/stats.data.
 *   grains [n][13] rows of lbmdem_get_grains AFTER the call, fhf [n][3]; nbsteps as after the call.
 * summary (may be NULL) receives energie_cin energy_p SE WF INCE TSLIP TRW for the console line. */
int lbmdem_write_dem(lbmdem_diag *d, const char *dir, int nfile, long nbsteps, const double *grains, const double *fhf,
                     const double *d11, double *summary);
/* truncates <dir>/stats.data and writes the header line (src/main.c:1867-1877) */
int lbmdem_write_stats_header(const char *dir);

#ifdef __cplusplus
}
#endif
#endif
