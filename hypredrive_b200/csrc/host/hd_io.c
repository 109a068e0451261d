/* hd_io.c -- readers for hypre's IJ on-disk formats; a rank reads its share of
 * consecutive parts (parts >= ranks) and concatenates them.
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

static FILE *open_part(const char *prefix, int part, int *binary, char *name, size_t cap)
{
   snprintf(name, cap, "%s.%05d.bin", prefix, part);
   FILE *fp = fopen(name, "rb");
   if (fp) { *binary = 1; return fp; }
   snprintf(name, cap, "%s.%05d", prefix, part);
   fp = fopen(name, "r");
   *binary = 0;
   return fp;
}

/* limits on what an (untrusted) part header may ask us to allocate: the reference caps them the same
 * way (IJMATRIX_MAX_PART_NNZ / IJVECTOR_MAX_PART_NROWS, src/internal/matrix.c, vector.c) and the
 * payload must also fit in the file */
#define HD_MAX_PART_NNZ   ((uint64_t)2147483647)
#define HD_MAX_PART_NROWS ((uint64_t)2147483647)

static uint64_t file_bytes_left(FILE *fp)
{
   long here = ftell(fp);
   if (here < 0 || fseek(fp, 0, SEEK_END)) return 0;
   long end = ftell(fp);
   fseek(fp, here, SEEK_SET);
   return end > here ? (uint64_t)(end - here) : 0;
}

/* A data set written with P parts can be read by any number of ranks R <= P: rank r takes
 * P/R consecutive parts (+1 for the first P%R ranks) and concatenates them -- the reference's
 * partition rule (src/internal/matrix.c:184-235, vector.c).  Returns 0 and the range [first, first+count). */
static int my_parts(const char *prefix, int rank, int *first, int *count)
{
   char name[2048];
   int  binary, nparts = 0, nprocs = 1;
   for (;;)
   {
      FILE *fp = open_part(prefix, nparts, &binary, name, sizeof(name));
      if (!fp) break;
      fclose(fp);
      nparts++;
      if (nparts > 1000000) break;
   }
   MPI_Comm_size(MPI_COMM_WORLD, &nprocs);
   if (nparts == 0) { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("file not found: %s.%05d[.bin]", prefix, 0); return 1; }
   if (nparts < nprocs)
   {
      hd_err_set(HYPREDRV_ERROR_FILE_UNEXPECTED_ENTRY);
      hd_err_msg("Invalid number of parts! (%d parts of %s for %d ranks)", nparts, prefix, nprocs);
      return 1;
   }
   int rem = nparts % nprocs;
   *count = nparts / nprocs + (rank < rem ? 1 : 0);
   *first = rank * (nparts / nprocs) + (rank < rem ? rank : rem);
   return 0;
}

static int parse_fail(const char *what, const char *name, HYPRE_IJMatrix A, HYPRE_IJVector v)
{
   if (A) HYPRE_IJMatrixDestroy(A);
   if (v) HYPRE_IJVectorDestroy(v);
   hd_err_set(HYPREDRV_ERROR_FILE_UNEXPECTED_ENTRY);
   hd_err_msg("could not parse %s file %s", what, name);
   return 1;
}

int hd_read_ij_matrix(const char *prefix, int rank, HYPRE_IJMatrix *out)
{
   char name[2048];
   int  first = 0, count = 0;
   if (my_parts(prefix, rank, &first, &count)) return 1;
   /* pass 1: the row range covered by my parts (headers only) */
   long long lo = 0, hi = -1;
   for (int p = 0; p < count; p++)
   {
      int   binary = 0;
      FILE *fp = open_part(prefix, first + p, &binary, name, sizeof(name));
      if (!fp) { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("matrix file not found: %s.%05d[.bin]", prefix, first + p); return 1; }
      long long il = 0, iu = -1;
      int       bad = 0;
      if (binary)
      {
         uint64_t h[11];
         if (fread(h, 8, 11, fp) != 11 || (h[1] != 4 && h[1] != 8) || (h[2] != 4 && h[2] != 8) || h[8] < h[7] ||
             h[8] - h[7] >= HD_MAX_PART_NROWS || h[6] > HD_MAX_PART_NNZ || h[6] * (2 * h[1] + h[2]) > file_bytes_left(fp)) bad = 1;
         il = (long long)h[7]; iu = (long long)h[8];
      }
      else
      {
         long long jl, ju;
         if (fscanf(fp, "%lld %lld %lld %lld", &il, &iu, &jl, &ju) != 4 || iu < il) bad = 1;
      }
      fclose(fp);
      if (bad) return parse_fail("matrix", name, NULL, NULL);
      if (p == 0) lo = il;
      else if (il != hi + 1)
      {
         hd_err_set(HYPREDRV_ERROR_FILE_UNEXPECTED_ENTRY);
         hd_err_msg("matrix part %s does not continue the row range of the previous part", name);
         return 1;
      }
      hi = iu;
   }
   /* pass 2: entries */
   HYPRE_IJMatrix A = NULL;
   HYPRE_IJMatrixCreate(MPI_COMM_WORLD, lo, hi, lo, hi, &A);
   HYPRE_IJMatrixInitialize(A);
   for (int p = 0; p < count; p++)
   {
      int   binary = 0, bad = 0;
      FILE *fp = open_part(prefix, first + p, &binary, name, sizeof(name));
      if (!fp) return parse_fail("matrix", name, A, NULL);
      HYPRE_Int one = 1;
      if (binary)
      {
         uint64_t h[11];
         if (fread(h, 8, 11, fp) != 11) bad = 1;
         if (!bad)
         {
            uint64_t      nnz = h[6];
            HYPRE_BigInt *r = malloc(sizeof(HYPRE_BigInt) * (size_t)(nnz ? nnz : 1)), *c = malloc(sizeof(HYPRE_BigInt) * (size_t)(nnz ? nnz : 1));
            double       *v = malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
            if (!r || !c || !v)
            {
               free(r); free(c); free(v); fclose(fp);
               HYPRE_IJMatrixDestroy(A);
               hd_err_set(HYPREDRV_ERROR_ALLOCATION);
               hd_err_msg("out of memory reading %s (%llu entries)", name, (unsigned long long)nnz);
               return 1;
            }
            bad = read_idx(fp, h[1], nnz, r) || read_idx(fp, h[1], nnz, c) || read_val(fp, h[2], nnz, v);
            for (uint64_t k = 0; k < nnz && !bad; k++) bad = HYPRE_IJMatrixSetValues(A, 1, &one, &r[k], &c[k], &v[k]);
            free(r); free(c); free(v);
         }
      }
      else
      {
         long long il, iu, jl, ju, i, j;
         double    v;
         if (fscanf(fp, "%lld %lld %lld %lld", &il, &iu, &jl, &ju) != 4) bad = 1;
         while (!bad && fscanf(fp, "%lld %lld %lf", &i, &j, &v) == 3)
         {
            HYPRE_BigInt ri = i, cj = j;
            if (HYPRE_IJMatrixSetValues(A, 1, &one, &ri, &cj, &v)) bad = 1;
         }
      }
      fclose(fp);
      if (bad) return parse_fail("matrix", name, A, NULL);
   }
   HYPRE_IJMatrixAssemble(A);
   *out = A;
   return 0;
}

int hd_read_ij_vector(const char *prefix, int rank, HYPRE_IJVector *out)
{
   char name[2048];
   int  first = 0, count = 0;
   if (my_parts(prefix, rank, &first, &count)) return 1;
   /* pass 1: sizes.  Binary parts carry only their length (the row offset of a part is the sum of the
    * preceding parts; the caller re-bases the vector onto the matrix row range), ASCII parts their range */
   long long lo = 0, total = 0;
   for (int p = 0; p < count; p++)
   {
      int   binary = 0, bad = 0;
      FILE *fp = open_part(prefix, first + p, &binary, name, sizeof(name));
      if (!fp) { hd_err_set(HYPREDRV_ERROR_FILE_NOT_FOUND); hd_err_msg("vector file not found: %s.%05d[.bin]", prefix, first + p); return 1; }
      if (binary)
      {
         uint64_t h[8];
         if (fread(h, 8, 8, fp) != 8 || (h[1] != 4 && h[1] != 8) || h[5] > HD_MAX_PART_NROWS || h[5] * h[1] > file_bytes_left(fp)) bad = 1;
         else total += (long long)h[5];
      }
      else
      {
         long long jl, ju;
         if (fscanf(fp, "%lld %lld", &jl, &ju) != 2 || ju < jl || (uint64_t)(ju - jl) >= HD_MAX_PART_NROWS) bad = 1;
         else { if (p == 0) lo = jl; total += ju - jl + 1; }
      }
      fclose(fp);
      if (bad) return parse_fail("vector", name, NULL, NULL);
   }
   HYPRE_IJVector v = NULL;
   HYPRE_IJVectorCreate(MPI_COMM_WORLD, lo, lo + total - 1, &v);
   HYPRE_IJVectorInitialize(v);
   if (!v || (total > 0 && !v->data)) { hd_err_set(HYPREDRV_ERROR_ALLOCATION); hd_err_msg("out of memory reading %s", prefix); return 1; }
   long long at = 0;
   for (int p = 0; p < count; p++)
   {
      int   binary = 0, bad = 0;
      FILE *fp = open_part(prefix, first + p, &binary, name, sizeof(name));
      if (!fp) return parse_fail("vector", name, NULL, v);
      if (binary)
      {
         uint64_t h[8];
         if (fread(h, 8, 8, fp) != 8 || at + (long long)h[5] > total) bad = 1;
         else { bad = read_val(fp, h[1], h[5], v->data + at); at += (long long)h[5]; }
      }
      else
      {
         long long jl, ju, j;
         double    x;
         if (fscanf(fp, "%lld %lld", &jl, &ju) != 2) bad = 1;
         while (!bad && fscanf(fp, "%lld %lf", &j, &x) == 2)
            if (j >= jl && j <= ju && j - lo >= 0 && j - lo < total) v->data[j - lo] = x;
         at += ju - jl + 1;
      }
      fclose(fp);
      if (bad) return parse_fail("vector", name, NULL, v);
   }
   *out = v;
   return 0;
}
