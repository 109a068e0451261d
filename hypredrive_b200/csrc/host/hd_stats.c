/* hd_stats.c -- per-solve statistics and the summary table.
 * Same columns and layout as the reference's STATISTICS SUMMARY (src/internal/stats.c:566-,
 * golden: examples/refOutput/ex1.txt:21-28): LS build / setup / solve times, initial residual
 * norm, final true relative residual norm, iterations.  Setup and solve times are measured
 * around device work that is synchronised (the reference uses MPI_Wtime without a device sync,
 * src/internal/stats.c:126-168). */
#include "hd_internal.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

double hd_wtime(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

hd_stats *hd_stats_create(void)
{
   hd_stats *s = calloc(1, sizeof(hd_stats));
   s->counter  = -1;
   s->ls_id    = -1;
   return s;
}

void hd_stats_print(const hd_stats *s, FILE *fp)
{
   /* header cells are right-aligned in their column width like the reference's PrintHeader
    * (src/internal/stats.c:564-595): "   times [s]" / "  times [ms]" */
   char         unit[32];
   const double f    = s->use_millisec ? 1000.0 : 1.0;
   snprintf(unit, sizeof(unit), "times %s", s->use_millisec ? "[ms]" : "[s]");
   const char  *div  = "+--------+-------------+-------------+-------------+------------+------------+--------+\n";
   for (int i = 0; i < 84; i++) fputc('=', fp); /* PRINT_EQUAL_LINE, stats.c:1231; width as in the golden outputs under examples/refOutput */
   fputc('\n', fp);
   if (s->name[0]) fprintf(fp, "\n\nSTATISTICS SUMMARY for %s:\n\n", s->name);
   else fprintf(fp, "\n\nSTATISTICS SUMMARY:\n\n");
   fprintf(fp, "%s", div);
   fprintf(fp, "|        |    LS build |       setup |       solve |    initial |   relative |        |\n");
   fprintf(fp, "|  Entry | %11s | %11s | %11s |  res. norm |  res. norm |  iters |\n", unit, unit, unit);
   fprintf(fp, "%s", div);
   int shown = 0;
   for (int i = 0; i <= s->counter && i < HD_STATS_MAX; i++)
   {
      if (!s->has_solve[i]) continue;
      if (s->has_build[i])
         fprintf(fp, "| %6d | %11.3f | %11.3f | %11.3f | %10.2e | %10.2e | %6d |\n", shown, f * s->build[i], f * s->setup[i],
                 f * s->solve[i], s->r0[i], s->rr[i], s->iters[i]);
      else
         fprintf(fp, "| %6d |             | %11.3f | %11.3f | %10.2e | %10.2e | %6d |\n", shown, f * s->setup[i],
                 f * s->solve[i], s->r0[i], s->rr[i], s->iters[i]);
      shown++;
   }
   fprintf(fp, "%s\n", div);
   fflush(fp);
}
