/* PETSc/SLEPc API shim header (test infrastructure, see petsc_shim.h) */
#include "petsc_shim.h"
