#ifndef ORACLE_SHIM_PETSCMAT_H
#define ORACLE_SHIM_PETSCMAT_H
#include <petscsys.h>
#endif
