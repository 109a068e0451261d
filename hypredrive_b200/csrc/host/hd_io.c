/* hd_io.c -- readers for hypre's IJ on-disk formats, one part per rank.
 *   ASCII  "<prefix>.%05d"      matrix: "ilower iupper jlower jupper" then "row col value" lines
 *                               vector: "jlower jupper" then "index value" lines
 *                               (hypre HYPRE_IJMatrixRead / HYPRE_IJVectorRead, used by the
 *                               reference at src/internal/linsys.c:973-976)
 *   binary "<prefix>.%05d.bin"  matrix: 11 x u64 header ([1] index bytes 4|8, [2] value bytes 4|8,
 *                               [6] nnz, [7..8] row range) then rows[], cols[], vals[]
 *                               vector: 8 x u64 header ([1] value bytes, [5] nrows) then values
 *                               (reference src/internal/matrix.c:153-470, vector.c:103-340; writer
 *                               layout tests/fuzz/tools/gen_ij{matrix,vector}_seed.py)
 * The containers produced are the host-side IJ objects of hd_ij.c. */
#include "hd_internal.h"
#include <stdlib.h>
#include <string.h>

static int read_idx(FILE *fp, uint64_t width, uint64_t n, HYPRE_BigInt *out)
{
   if (width == 8) return fread(out, 8, n, fp) == n ? 0 : 1;
   uint32_t *b = malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
   int       bad = fread(b, 4, n, fp) != n;
   for (uint64_t i = 0; !bad && i < n; i++) out[i] = (HYPRE_BigInt)b[i];
   free(b);
   return bad;
}

static int read_val(FILE *fp, uint64_t width, uint64_t n, double *out)
{
   if (width == 8) return fread(out, 8, n, fp) == n ? 0 : 1;
   float *b = malloc(sizeof(float) * (size_t)(n ? n : 1));
   int    bad = fread(b, 4, n, fp) != n;
   for (uint64_t i = 0; !bad && i < n; i++) out[i] = (double)b[i];
   free(b);
   return bad;
}

static FILE *open_part(const char *prefix, int rank, int *binary, char *name, size_t cap)
{
   snprintf(name, cap, "%s.%05d.bin", prefix, rank);
   FILE *fp = fopen(name, "rb");
   if (fp) { *binary = 1; return fp; }
   snprintf(name, cap, "%s.%05d", prefix, rank);
   fp = fopen(name, "r");
   *binary = 0;
   return fp;
}

int hd_read_ij_matrix(const char *prefix, int rank, HYPRE_IJMatrix *out)
{
   char  name[2048];
   int   binary = 0;
   FILE *fp = open_part(prefix, rank, &binary, name, sizeof(name));
   if (!fp) { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("matrix file not found: %s.%05d[.bin]", prefix, rank); return 1; }
   HYPRE_IJMatrix A = NULL;
   int            bad = 0;
   if (binary)
   {
      uint64_t h[11];
      if (fread(h, 8, 11, fp) != 11 || (h[1] != 4 && h[1] != 8) || (h[2] != 4 && h[2] != 8) || h[8] < h[7]) bad = 1;
      if (!bad)
      {
         uint64_t      nnz = h[6];
         HYPRE_BigInt *r = malloc(sizeof(HYPRE_BigInt) * (size_t)(nnz ? nnz : 1)), *c = malloc(sizeof(HYPRE_BigInt) * (size_t)(nnz ? nnz : 1));
         double       *v = malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
         bad = read_idx(fp, h[1], nnz, r) || read_idx(fp, h[1], nnz, c) || read_val(fp, h[2], nnz, v);
         if (!bad)
         {
            HYPRE_IJMatrixCreate(MPI_COMM_WORLD, (HYPRE_BigInt)h[7], (HYPRE_BigInt)h[8], (HYPRE_BigInt)h[7], (HYPRE_BigInt)h[8], &A);
            HYPRE_IJMatrixInitialize(A);
            HYPRE_Int one = 1;
            for (uint64_t k = 0; k < nnz && !bad; k++) bad = HYPRE_IJMatrixSetValues(A, 1, &one, &r[k], &c[k], &v[k]);
         }
         free(r); free(c); free(v);
      }
   }
   else
   {
      long long il, iu, jl, ju;
      if (fscanf(fp, "%lld %lld %lld %lld", &il, &iu, &jl, &ju) != 4 || iu < il) bad = 1;
      if (!bad)
      {
         HYPRE_IJMatrixCreate(MPI_COMM_WORLD, il, iu, jl, ju, &A);
         HYPRE_IJMatrixInitialize(A);
         long long i, j;
         double    v;
         HYPRE_Int one = 1;
         while (fscanf(fp, "%lld %lld %lf", &i, &j, &v) == 3)
         {
            HYPRE_BigInt ri = i, cj = j;
            if (HYPRE_IJMatrixSetValues(A, 1, &one, &ri, &cj, &v)) { bad = 1; break; }
         }
      }
   }
   fclose(fp);
   if (bad)
   {
      if (A) HYPRE_IJMatrixDestroy(A);
      hd_err_set(HYPREDRV_ERROR_FILE_UNEXPECTED_ENTRY);
      hd_err_msg("could not parse matrix file %s", name);
      return 1;
   }
   HYPRE_IJMatrixAssemble(A);
   *out = A;
   return 0;
}

int hd_read_ij_vector(const char *prefix, int rank, HYPRE_IJVector *out)
{
   char  name[2048];
   int   binary = 0;
   FILE *fp = open_part(prefix, rank, &binary, name, sizeof(name));
   if (!fp) { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("vector file not found: %s.%05d[.bin]", prefix, rank); return 1; }
   HYPRE_IJVector v = NULL;
   int            bad = 0;
   if (binary)
   {
      uint64_t h[8];
      if (fread(h, 8, 8, fp) != 8 || (h[1] != 4 && h[1] != 8)) bad = 1;
      if (!bad)
      {
         uint64_t n = h[5];
         /* the row offset of a part is the sum of the preceding parts; with one part per rank
          * the caller re-bases the vector onto the matrix row range */
         HYPRE_IJVectorCreate(MPI_COMM_WORLD, 0, (HYPRE_BigInt)n - 1, &v);
         HYPRE_IJVectorInitialize(v);
         bad = read_val(fp, h[1], n, v->data);
      }
   }
   else
   {
      long long jl, ju;
      if (fscanf(fp, "%lld %lld", &jl, &ju) != 2 || ju < jl) bad = 1;
      if (!bad)
      {
         HYPRE_IJVectorCreate(MPI_COMM_WORLD, jl, ju, &v);
         HYPRE_IJVectorInitialize(v);
         long long j;
         double    x;
         while (fscanf(fp, "%lld %lf", &j, &x) == 2)
            if (j >= jl && j <= ju) v->data[j - jl] = x;
      }
   }
   fclose(fp);
   if (bad)
   {
      if (v) HYPRE_IJVectorDestroy(v);
      hd_err_set(HYPREDRV_ERROR_FILE_UNEXPECTED_ENTRY);
      hd_err_msg("could not parse vector file %s", name);
      return 1;
   }
   *out = v;
   return 0;
}
