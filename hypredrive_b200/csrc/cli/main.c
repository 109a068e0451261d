/* hypredrive-cli for hypredrive_b200: same control flow as the reference driver
 * (src/internal/main.c:175-338) written against nothing but the HYPREDRV_* API:
 *   hypredrive-cli <config.yml> [-a --path:to:key value ...]
 * Build the system from the files named in the YAML, loop over repetitions
 * (reset guess, create, setup, apply, destroy), print the statistics table. */
#include <stdio.h>
#include <string.h>
#include "HYPREDRV.h"
#include "HYPREDRV_utils.h"

int main(int argc, char **argv)
{
   MPI_Comm comm = MPI_COMM_WORLD;
   int      myid = 0;
   MPI_Init(&argc, &argv);
   MPI_Comm_rank(comm, &myid);
   if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help"))
   {
      if (!myid) printf("Usage: %s <config.yml> [-a --path:to:key value ...]\n", argv[0]);
      MPI_Finalize();
      return argc < 2;
   }
   HYPREDRV_SAFE_CALL(HYPREDRV_Initialize());
   HYPREDRV_t obj = NULL;
   HYPREDRV_SAFE_CALL(HYPREDRV_Create(comm, &obj));
   HYPREDRV_SAFE_CALL(HYPREDRV_PrintLibInfo(comm, 1));
   HYPREDRV_SAFE_CALL(HYPREDRV_InputArgsParse(argc - 1, argv + 1, obj));
   int num_ls = 1, num_variants = 1, num_reps = 1;
   HYPREDRV_SAFE_CALL(HYPREDRV_InputArgsGetNumLinearSystems(obj, &num_ls));
   HYPREDRV_SAFE_CALL(HYPREDRV_InputArgsGetNumPreconVariants(obj, &num_variants));
   for (int k = 0; k < num_ls; k++)
   {
      HYPREDRV_SAFE_CALL(HYPREDRV_LinearSystemBuild(obj));
      for (int v = 0; v < num_variants; v++)
      {
         HYPREDRV_SAFE_CALL(HYPREDRV_InputArgsSetPreconVariant(obj, v));
         HYPREDRV_SAFE_CALL(HYPREDRV_InputArgsGetNumRepetitions(obj, &num_reps));
         for (int i = 0; i < num_reps; i++)
         {
            HYPREDRV_SAFE_CALL(HYPREDRV_AnnotateBegin(obj, "Run", i));
            HYPREDRV_SAFE_CALL(HYPREDRV_LinearSystemResetInitialGuess(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_PreconCreate(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_LinearSolverCreate(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_LinearSolverSetup(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_LinearSolverApply(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_PreconDestroy(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_LinearSolverDestroy(obj));
            HYPREDRV_SAFE_CALL(HYPREDRV_AnnotateEnd(obj, "Run", i));
         }
      }
   }
   if (!myid) HYPREDRV_SAFE_CALL(HYPREDRV_StatsPrint(obj));
   HYPREDRV_SAFE_CALL(HYPREDRV_Destroy(&obj));
   HYPREDRV_SAFE_CALL(HYPREDRV_PrintExitInfo(comm, argv[0]));
   HYPREDRV_SAFE_CALL(HYPREDRV_Finalize());
   MPI_Finalize();
   return 0;
}
