"""hypredrive_b200 -- B200-native (sm_100a CUDA) implementation of the solve path hypredrive
drives through hypre: BoomerAMG-preconditioned PCG/GMRES on a ParCSR matrix, behind the
HYPREDRV_* C API (include/HYPREDRV.h).  `hdk` binds the thin device C-ABI, `driver` mirrors the
reference's Python driver interface on top of the HYPREDRV_* entry points."""
from .driver import (BIGINT_DTYPE, REAL_DTYPE, HypreDrive, HypreDriveError, SolveResult,  # noqa: F401
                     normalize_options, solve)

__all__ = ["HypreDrive", "HypreDriveError", "SolveResult", "solve", "normalize_options", "BIGINT_DTYPE", "REAL_DTYPE"]
