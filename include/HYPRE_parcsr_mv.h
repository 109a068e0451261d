/* HYPRE_parcsr_mv.h -- forwarding header of the hypre interface shim (see HYPRE.h). */
#include "HYPRE.h"
