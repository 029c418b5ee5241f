/* vtk_writer.c -- see vtk_writer.h */
#include "vtk_writer.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int put_floats_be(FILE *fp, const float *v, size_t n) {
  enum { CHUNK = 16384 };
  uint32_t buf[CHUNK];
  while (n) {
    const size_t k = n < CHUNK ? n : CHUNK;
    for (size_t i = 0; i < k; ++i) {
      uint32_t u;
      memcpy(&u, &v[i], 4);
      buf[i] = (u >> 24) | ((u >> 8) & 0xff00u) | ((u << 8) & 0xff0000u) | (u << 24);
    }
    if (fwrite(buf, 4, k, fp) != k) return -1;
    v += k;
    n -= k;
  }
  return 0;
}

static int put_axis(FILE *fp, char axis, int n, float step) {
  /* "X_COORDINATES n float\n" then n big-endian floats i*step, no newline after binary data */
  float *c = (float *)malloc(sizeof(float) * (size_t)n);
  if (!c) return -1;
  for (int i = 0; i < n; ++i) c[i] = i * step;
  fprintf(fp, "%c_COORDINATES %d float\n", axis, n);
  const int rc = put_floats_be(fp, c, (size_t)n);
  free(c);
  return rc;
}

int lbmdem_write_vtk_field(const char *basename, const char *varname, int nx, int ny, int ncomp, const float *data) {
  char path[1024];
  snprintf(path, sizeof path, "%s.vtk", basename);
  FILE *fp = fopen(path, "wb");
  if (!fp) return -1;
  int rc = 0;
  fprintf(fp, "# vtk DataFile Version 2.0\nWritten using VisIt writer\nBINARY\nDATASET RECTILINEAR_GRID\n");
  fprintf(fp, "DIMENSIONS %d %d %d\n", nx, ny, 1);
  const float pas = 1. / nx; /* src/main.c:254-257: the SAME spacing 1/nx on both axes */
  rc |= put_axis(fp, 'X', nx, pas);
  rc |= put_axis(fp, 'Y', ny, pas);
  rc |= put_axis(fp, 'Z', 1, 0.f);
  /* an empty cell-data section is always announced (visit_writer.c:368-370) */
  fprintf(fp, "CELL_DATA %d\n", (nx - 1) * (ny - 1));
  fprintf(fp, "POINT_DATA %d\n", nx * ny);
  if (ncomp == 1) fprintf(fp, "SCALARS %s float\nLOOKUP_TABLE default\n", varname);
  else fprintf(fp, "VECTORS %s float\n", varname);
  rc |= put_floats_be(fp, data, (size_t)nx * ny * (size_t)ncomp);
  if (fclose(fp)) rc = -1;
  return rc ? -1 : 0;
}

int lbmdem_write_vtk_frame(const char *dir, int nfile, int nx, int ny, const float *grain_pressure,
                           const float *grain_velocity, const float *grain_acceleration, const float *fluid_pressure,
                           const float *fluid_velocity) {
  static const char *names[5] = {"grain_pressure", "grain_velocity", "grain_acceleration", "fluid_pressure",
                                 "fluid_velocity"};
  const int ncomp[5] = {1, 3, 3, 1, 3};
  const float *data[5] = {grain_pressure, grain_velocity, grain_acceleration, fluid_pressure, fluid_velocity};
  for (int k = 0; k < 5; ++k) {
    char base[1024];
    snprintf(base, sizeof base, "%s%s%s_%.6i", dir ? dir : "", (dir && *dir) ? "/" : "", names[k], nfile);
    if (lbmdem_write_vtk_field(base, names[k], nx, ny, ncomp[k], data[k])) return -1;
  }
  return 0;
}
