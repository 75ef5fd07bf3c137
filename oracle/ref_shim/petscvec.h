/* oracle shim: a sequential array-backed Vec, enough for the reference's src/jacobi/tests/ex1.c. */
#ifndef ORACLE_SHIM_PETSCVEC_H
#define ORACLE_SHIM_PETSCVEC_H
#include <petscsys.h>
struct _p_Vec { struct _p_PetscObject hdr; PetscInt n; PetscScalar *a; };
PetscErrorCode VecCreate(MPI_Comm,Vec*);
PetscErrorCode VecSetSizes(Vec,PetscInt,PetscInt);
PetscErrorCode VecSetFromOptions(Vec);
PetscErrorCode VecGetArray(Vec,PetscScalar**);
PetscErrorCode VecRestoreArray(Vec,PetscScalar**);
PetscErrorCode VecGetSize(Vec,PetscInt*);
PetscErrorCode VecGetLocalSize(Vec,PetscInt*);
PetscErrorCode VecDestroy(Vec*);
#endif
