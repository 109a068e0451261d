/* gen.c -- synthetic input generators of the oracle (TEST INFRASTRUCTURE ONLY).
 * Restates, for a single partition (1x1x1 process grid, x-fastest lexicographic numbering),
 * the reference example generators:
 *   examples/src/C_laplacian/laplacian.c:720-921   (7-point, BuildLaplacianSystem_7pt)
 *   examples/src/C_laplacian/laplacian.c:1138-1356 (27-point, BuildLaplacianSystem_27pt)
 *   examples/src/C_convdif/convdif.c:782-990       (upwind FV convection-diffusion, wmax = 0)
 * Rows are emitted in the reference's column order; the caller applies the diagonal-first
 * swap that hypre's IJ assembly performs (ocsr_diag_first).
 */
#include "oracle.h"
#include <stdlib.h>

ocsr *ogen_laplace7(int nx, int ny, int nz, double cx, double cy, double cz)
{
   int64_t n = (int64_t)nx * ny * nz;
   ocsr   *A = ocsr_alloc((int)n, (int)n, 7 * n, 1);
   int     p = 0;
   for (int gz = 0; gz < nz; gz++)
      for (int gy = 0; gy < ny; gy++)
         for (int gx = 0; gx < nx; gx++)
         {
            int row    = (gz * ny + gy) * nx + gx;
            A->ia[row] = p;
            A->ja[p] = row; A->a[p++] = 2.0 * (cx + cy + cz);
            if (gz > 0) { A->ja[p] = row - nx * ny; A->a[p++] = -cz; }
            if (gy > 0) { A->ja[p] = row - nx; A->a[p++] = -cy; }
            if (gx > 0) { A->ja[p] = row - 1; A->a[p++] = -cx; }
            if (gx < nx - 1) { A->ja[p] = row + 1; A->a[p++] = -cx; }
            if (gy < ny - 1) { A->ja[p] = row + nx; A->a[p++] = -cy; }
            if (gz < nz - 1) { A->ja[p] = row + nx * ny; A->a[p++] = -cz; }
         }
   A->ia[n] = p;
   return A;
}

/* RHS = 1 on the plane gy == 0, else 0 (laplacian.c:898-906, 1337-1345) */
void ogen_rhs_yplane(int nx, int ny, int nz, double *b)
{
   for (int gz = 0; gz < nz; gz++)
      for (int gy = 0; gy < ny; gy++)
         for (int gx = 0; gx < nx; gx++) b[(gz * ny + gy) * nx + gx] = (gy == 0) ? 1.0 : 0.0;
}

ocsr *ogen_laplace27(int nx, int ny, int nz, double cx, double cy, double cz)
{
   int64_t n = (int64_t)nx * ny * nz;
   ocsr   *A = ocsr_alloc((int)n, (int)n, 27 * n, 1);
   int     p = 0;
   for (int gz = 0; gz < nz; gz++)
      for (int gy = 0; gy < ny; gy++)
         for (int gx = 0; gx < nx; gx++)
         {
            int    row    = (gz * ny + gy) * nx + gx;
            double center = 0.0;
            A->ia[row]    = p;
            for (int dz = -1; dz <= 1; dz++)
               for (int dy = -1; dy <= 1; dy++)
                  for (int dx = -1; dx <= 1; dx++)
                  {
                     if (!dx && !dy && !dz) continue;
                     int    ndiff = (dx != 0) + (dy != 0) + (dz != 0);
                     double adj   = 0.0;
                     if (dx) adj += cx / ndiff;
                     if (dy) adj += cy / ndiff;
                     if (dz) adj += cz / ndiff;
                     int x = gx + dx, y = gy + dy, z = gz + dz;
                     if (x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz)
                     {
                        A->ja[p]  = (z * ny + y) * nx + x;
                        A->a[p++] = -adj;
                     }
                     center += adj; /* in-domain and Dirichlet-truncated neighbours alike */
                  }
            A->ja[p]  = row; /* centre entry goes last; IJ assembly swaps it to the front */
            A->a[p++] = center;
         }
   A->ia[n] = p;
   return A;
}

static double axial_velocity(double y, double z, double Hy, double Hz, double umax)
{
   return 16.0 * umax * (y / Hy) * (1.0 - y / Hy) * (z / Hz) * (1.0 - z / Hz);
}

/* One backward-Euler step from c_old = 0 on the 4 x 1 x 1 duct (convdif.c defaults),
 * swirl off (wmax = 0).  Column order: diag, W, E, S, N, D, U. */
ocsr *ogen_convdif7(int nx, int ny, int nz, double kappa, double umax, double dt)
{
   int64_t      n  = (int64_t)nx * ny * nz;
   ocsr        *A  = ocsr_alloc((int)n, (int)n, 7 * n, 1);
   const double Lx = 4.0, Ly = 1.0, Lz = 1.0;
   const double hx = Lx / nx, hy = Ly / ny, hz = Lz / nz;
   const double vol = hx * hy * hz;
   const double Dx = kappa * hy * hz / hx, Dy = kappa * hx * hz / hy, Dz = kappa * hx * hy / hz;
   const double Ax = hy * hz;
   int          p  = 0;
   for (int gz = 0; gz < nz; gz++)
   {
      double z = ((double)gz + 0.5) * hz;
      for (int gy = 0; gy < ny; gy++)
      {
         double y      = ((double)gy + 0.5) * hy;
         double Cf     = axial_velocity(y, z, Ly, Lz, umax) * Ax;
         double Cf_pos = Cf > 0.0 ? Cf : 0.0, Cf_neg = Cf < 0.0 ? -Cf : 0.0;
         for (int gx = 0; gx < nx; gx++)
         {
            int    row  = (gz * ny + gy) * nx + gx;
            int    d    = p++;
            double diag = vol / dt;
            A->ia[row]  = d;
            A->ja[d]    = row;
            if (gx > 0) { diag += Dx + Cf_neg; A->ja[p] = row - 1; A->a[p++] = -(Dx + Cf_pos); }
            else diag += 2.0 * Dx + Cf_neg;
            if (gx < nx - 1) { diag += Dx + Cf_pos; A->ja[p] = row + 1; A->a[p++] = -(Dx + Cf_neg); }
            else diag += Cf_pos;
            if (gy > 0) { diag += Dy; A->ja[p] = row - nx; A->a[p++] = -Dy; }
            if (gy < ny - 1) { diag += Dy; A->ja[p] = row + nx; A->a[p++] = -Dy; }
            if (gz > 0) { diag += Dz; A->ja[p] = row - nx * ny; A->a[p++] = -Dz; }
            if (gz < nz - 1) { diag += Dz; A->ja[p] = row + nx * ny; A->a[p++] = -Dz; }
            A->a[d] = diag;
         }
      }
   }
   A->ia[n] = p;
   return A;
}

/* RHS of the convection-diffusion step above: inlet Dirichlet c_in = 1 through the west
 * boundary faces, c_old = 0 (convdif.c:905-916). */
void ogen_convdif7_rhs(int nx, int ny, int nz, double kappa, double umax, double dt, double *b)
{
   const double Lx = 4.0, Ly = 1.0, Lz = 1.0;
   const double hx = Lx / nx, hy = Ly / ny, hz = Lz / nz;
   const double Dx = kappa * hy * hz / hx, Ax = hy * hz;
   (void)dt;
   for (int gz = 0; gz < nz; gz++)
   {
      double z = ((double)gz + 0.5) * hz;
      for (int gy = 0; gy < ny; gy++)
      {
         double y      = ((double)gy + 0.5) * hy;
         double Cf     = axial_velocity(y, z, Ly, Lz, umax) * Ax;
         double Cf_pos = Cf > 0.0 ? Cf : 0.0;
         for (int gx = 0; gx < nx; gx++)
            b[(gz * ny + gy) * nx + gx] = (gx == 0) ? (2.0 * Dx + Cf_pos) * 1.0 : 0.0;
      }
   }
}
