#ifndef ORACLE_SHIM_PETSCIMPL_H
#define ORACLE_SHIM_PETSCIMPL_H
#include <petscsys.h>
#endif
