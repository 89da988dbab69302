/* oracle PETSc/MPI shim (test infrastructure): forwards to petsc.h */
#include "petsc.h"
