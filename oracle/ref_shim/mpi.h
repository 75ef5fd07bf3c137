/* oracle shim: single-process stand-in for MPI. */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
typedef int MPI_Comm;
#define MPI_COMM_SELF  ((MPI_Comm)1)
#define MPI_COMM_WORLD ((MPI_Comm)2)
#endif
