/* HYPRE_parcsr_ls.h -- forwarding header of the hypre interface shim (see HYPRE.h). */
#include "HYPRE.h"
