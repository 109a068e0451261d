// hdk_amg.cuh -- device BoomerAMG hierarchy (internal).
#pragma once
#include "hdk_internal.cuh"

namespace hdk {

struct AmgLevel
{
   hdk_csr_s *A = nullptr;  // level operator (level 0 is borrowed from the caller)
   bool       owns_A = false;
   hdk_csr_s *P = nullptr;  // n_l x n_{l+1}
   hdk_csr_s *R = nullptr;  // n_{l+1} x n_l  (P^T stored explicitly: keep_transpose)
   DevCSR     S;            // strength pattern (diag block)
   int       *cf = nullptr;
   int       *f2c = nullptr; // exclusive scan of (cf > 0), n+1 entries (kept for the N > 1 slicing)
   double    *measure = nullptr;
   double    *l1_down = nullptr, *l1_up = nullptr; // smoother diagonals (may alias)
   double    *u = nullptr, *f = nullptr, *t = nullptr;
   DevCSR     L;            // strict lower triangle (two-stage GS only)
   int        n = 0;
};

} // namespace hdk

struct hdk_amg_s
{
   hdk_amg_params             prm;
   std::vector<hdk::AmgLevel> lev;
   int                        nlev = 0;
   double                    *ge_inv = nullptr; // dense inverse of the coarsest operator
   int                        ge_n = 0;
   double                     op_complexity = 0.0;
   double                     vcycle_bytes = 0.0;
   bool                       keep_debug = false; // keep S / measure for introspection (tunable amg_keep_debug)
   bool                       prefilled_l0 = false; // next zero-guess cycle: level-0 first sweep already done by the caller
   bool                       keep_f2c = false;
   // N > 1: levels [0, nlev) are row-distributed; from level `tail_level` on, the hierarchy is
   // replicated on every rank (`tail`, a serial hierarchy of the GLOBAL problem) and the
   // restricted right-hand side is summed over ranks into `full_f`
   hdk_amg_s                 *tail = nullptr;
   int                        tail_level = 0;
   int64_t                    tail_off = 0, tail_cnt = 0, tail_n = 0; // my slice of the first replicated level
   double                    *full_f = nullptr, *full_u = nullptr;
};

namespace hdk {
// z = M^{-1} r (one V-cycle from a zero guess); if fin != FIN_NONE the last kernel also
// produces <r,z> and applies `fin`
int amg_precond(hdk_amg_s *M, const double *r, double *z, int fin, double *fin_out);
// V-cycle over levels [l0, nlev) of M; level l0 uses the caller's vectors
int amg_cycle(hdk_amg_s *M, const double *f, double *u, bool zero_guess, int fin, double *fin_out, int l0 = 0);
int exclusive_scan_int(const int *in, int *out, int n);
bool amg_prefill_target(hdk_amg_s *M, double *z, double **buf, const double **d, double *w);
} // namespace hdk
