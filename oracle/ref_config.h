/* TEST INFRASTRUCTURE ONLY -- build prelude for the reference oracle (oracle/_ref).
 *
 * The reference's compile-time switches live in src/ops_config.h (reference
 * src/ops_config.h:17-40), which every reference source reaches through
 * `#include "ops_config.h"` -- a quoted include that resolves next to the
 * including file, so neither -I nor -D can override it.  This prelude is
 * force-included (gcc -include) before every reference translation unit: it
 * claims the reference header's include guard, so that header expands to
 * nothing, and supplies the same set of switches with OpenMP enabled.  The
 * reference sources themselves are compiled where they lie, unmodified.
 *
 * The reference hard-codes OMP_NUM_THREADS as an integer literal
 * (src/ops_config.h:38-40).  Every use is inside an expression or a
 * num_threads() clause, so here it expands to a run-time variable that
 * oracle/ref_driver.c owns; the thread count can follow the host the oracle
 * runs on without rebuilding.
 */
#ifndef _OPS_CONFIG_H_
#define _OPS_CONFIG_H_

#define OPS_USE_HYPRE     0
#define OPS_USE_INTEL_MKL 0
#define OPS_USE_MATLAB    0
#define OPS_USE_MEMWATCH  0
#define OPS_USE_MPI       0
#define OPS_USE_MUMPS     0
#ifndef OPS_USE_OMP
#define OPS_USE_OMP       1
#endif
#define OPS_USE_PHG       0
#define OPS_USE_PETSC     0
#define OPS_USE_SLEPC     0
#define OPS_USE_UMFPACK   0
#define PRINT_RANK        0

#define FORTRAN_WRAPPER(x) x ## _

#if OPS_USE_OMP
extern int gcge_ref_omp_threads;
#define OMP_NUM_THREADS gcge_ref_omp_threads
#endif

#endif
