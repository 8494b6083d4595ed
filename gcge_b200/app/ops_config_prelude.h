/* Build prelude for compiling app_b200.c against the reference headers OUTSIDE the
 * reference tree.  The reference's src/ops_config.h is reached by a quoted include that
 * resolves next to the including header, so a standalone build claims its include guard
 * and supplies the switches of a plain serial, non-MPI configuration (the values of
 * reference src/ops_config.h:17-28).  Inside the reference tree this file is not needed:
 * app_b200.c is compiled like app_ccs.c.
 */
#ifndef _OPS_CONFIG_H_
#define _OPS_CONFIG_H_
#define OPS_USE_HYPRE     0
#define OPS_USE_INTEL_MKL 0
#define OPS_USE_MATLAB    0
#define OPS_USE_MEMWATCH  0
#define OPS_USE_MPI       0
#define OPS_USE_MUMPS     0
#define OPS_USE_OMP       0
#define OPS_USE_PHG       0
#define OPS_USE_PETSC     0
#define OPS_USE_SLEPC     0
#define OPS_USE_UMFPACK   0
#define PRINT_RANK        0
#define FORTRAN_WRAPPER(x) x ## _
#endif
