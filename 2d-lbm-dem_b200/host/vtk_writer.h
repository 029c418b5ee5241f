/*
 * vtk_writer.h -- the one legacy-VTK layout the reference writes (src/main.c:326-328 through
 * write_rectilinear_mesh, src/visit_writer.c:895-933): BINARY rectilinear grid, one point-data
 * variable per file, big-endian float32, no trailing newline.  Not a port of visit_writer: only
 * this byte layout is reproduced.
 */
#ifndef LBMDEM_VTK_WRITER_H
#define LBMDEM_VTK_WRITER_H
#ifdef __cplusplus
extern "C" {
#endif

/* data: nx*ny values (ncomp = 1, SCALARS) or nx*ny*3 (ncomp = 3, VECTORS), [y][x] order, x fastest.
 * Writes "<basename>.vtk".  Returns 0, or -1 if the file cannot be written. */
int lbmdem_write_vtk_field(const char *basename, const char *varname, int nx, int ny, int ncomp, const float *data);

/* the five files of one film frame (src/main.c:239-249): <dir>/grain_pressure_%06d.vtk, ... */
int lbmdem_write_vtk_frame(const char *dir, int nfile, int nx, int ny, const float *grain_pressure,
                           const float *grain_velocity, const float *grain_acceleration, const float *fluid_pressure,
                           const float *fluid_velocity);

#ifdef __cplusplus
}
#endif
#endif
