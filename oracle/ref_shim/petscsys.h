/* TEST INFRASTRUCTURE (oracle): single-header stand-in for the handful of PETSc (~3.4-dev) facilities that the
 * reference's dJacobi/dQuadrature layer touches, so that the reference sources compile UNMODIFIED from
 * /root/reference (see oracle/Makefile).  Nothing in the product links or includes this file.
 *
 * Symbol list derived from a grep of the seven reference files named in SURVEY.md Appendix A.
 */
#ifndef ORACLE_SHIM_PETSCSYS_H
#define ORACLE_SHIM_PETSCSYS_H

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <mpi.h>

typedef int       PetscErrorCode;
typedef int       PetscInt;
typedef int       PetscMPIInt;
typedef double    PetscReal;
typedef double    PetscScalar;
typedef double    MatScalar;
typedef int       PetscClassId;
typedef int       PetscLogEvent;
typedef int       PetscEnum;
typedef enum {PETSC_FALSE, PETSC_TRUE} PetscBool;
typedef enum {NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES, MAX_VALUES} InsertMode;
typedef enum {PETSC_ERROR_INITIAL = 0, PETSC_ERROR_REPEAT = 1, PETSC_ERROR_IN_CXX = 2} PetscErrorType;

#define PETSC_NULL         NULL
#define PETSC_DECIDE       (-1)
#define PETSC_COMM_SELF    MPI_COMM_SELF
#define PETSC_COMM_WORLD   MPI_COMM_WORLD
#define PETSC_MAX_PATH_LEN 4096
#define PETSC_MEMALIGN     64
#define PETSC_UNUSED       __attribute__((unused))
#define PETSC_FUNCTION_NAME __func__
#define PETSC_PI           3.14159265358979323846264338327950288419716939937510582
#define __SDIR__           ""

#define PETSC_ERR_MEM            55
#define PETSC_ERR_SUP            56
#define PETSC_ERR_SIG            59
#define PETSC_ERR_LIB            76
#define PETSC_ERR_PLIB           77
#define PETSC_ERR_MEMC           78
#define PETSC_ERR_USER           83
#define PETSC_ERR_ARG_SIZ        60
#define PETSC_ERR_ARG_IDN        61
#define PETSC_ERR_ARG_WRONG      62
#define PETSC_ERR_ARG_CORRUPT    64
#define PETSC_ERR_ARG_OUTOFRANGE 63
#define PETSC_ERR_ARG_BADPTR     68
#define PETSC_ERR_ARG_NOTSAMETYPE 69
#define PETSC_ERR_ARG_WRONGSTATE 73
#define PETSC_ERR_ARG_INCOMP     75
#define PETSC_ERR_ARG_NULL       85
#define PETSC_ERR_ARG_UNKNOWN_TYPE 86

#define PetscUnlikely(c)   __builtin_expect(!!(c),0)
#define PetscLikely(c)     __builtin_expect(!!(c),1)
#define PetscRealPart(a)   (a)
#define PetscExpScalar(a)  exp(a)
#define PetscFunctionBegin do {} while (0)
#define PetscFunctionReturn(a) return (a)

PetscErrorCode PetscError(MPI_Comm,int,const char*,const char*,const char*,PetscErrorCode,PetscErrorType,const char*,...);

/* ---- objects --------------------------------------------------------------------------------------------------- */
typedef struct _p_PetscObject *PetscObject;
typedef struct _p_PetscViewer *PetscViewer;
typedef struct _p_Mat *Mat;
typedef struct _p_Vec *Vec;
typedef struct _n_PetscFunctionList *PetscFunctionList;

struct _p_PetscObject {
  PetscClassId classid;
  MPI_Comm     comm;
  char         *type_name;
  char         *prefix;
  int          refct;
  const char   *class_name;
  PetscFunctionList qlist;       /* composed functions */
};
#define PETSCHEADER(ObjectOps) struct _p_PetscObject hdr; ObjectOps ops[1]

struct _p_PetscViewer { struct _p_PetscObject hdr; FILE *fp; int tab; };
extern PetscClassId PETSC_VIEWER_CLASSID;
#define MAT_CLASSID 1212
#define VEC_CLASSID 1213
#define PETSCVIEWERASCII "ascii"
PetscViewer OracleShimStdoutViewer(void);
#define PETSC_VIEWER_STDOUT_WORLD OracleShimStdoutViewer()
#define PETSC_VIEWER_STDOUT_SELF  OracleShimStdoutViewer()

PetscErrorCode OracleShimHeaderCreate(void **obj,size_t bytes,PetscClassId classid,const char *class_name,MPI_Comm comm);
#define PetscHeaderCreate(h,tp,pops,cook,class_name,descr,mansec,com,des,vie) \
  OracleShimHeaderCreate((void**)&(h),sizeof(struct tp),(cook),(class_name),(com))
PetscErrorCode OracleShimHeaderDestroy(void **obj);
#define PetscHeaderDestroy(h) OracleShimHeaderDestroy((void**)(h))

PetscErrorCode PetscObjectTypeCompare(PetscObject,const char*,PetscBool*);
PetscErrorCode PetscObjectChangeTypeName(PetscObject,const char*);
PetscErrorCode OracleShimCompose(PetscObject,const char*,void (*)(void));
PetscErrorCode OracleShimQuery(PetscObject,const char*,void (**)(void));
#define PetscObjectComposeFunctionDynamic(obj,name,fname,fptr) OracleShimCompose((obj),(name),(void (*)(void))(fptr))
#define PetscTryMethod(obj,A,B,C) \
  0;{ PetscErrorCode (*_tm_f)B, _tm_err; \
    _tm_err = OracleShimQuery((PetscObject)(obj),(A),(void (**)(void))&_tm_f);if (_tm_err) return _tm_err; \
    if (_tm_f) {_tm_err = (*_tm_f)C;if (_tm_err) return _tm_err;} \
  }

#define PetscValidHeaderSpecific(h,ck,arg) do { \
    if (!(h)) return PetscError(PETSC_COMM_SELF,__LINE__,__func__,__FILE__,"",PETSC_ERR_ARG_NULL,PETSC_ERROR_INITIAL,"Null Object: Parameter # %d",(arg)); \
    if (((PetscObject)(h))->classid != (ck)) return PetscError(PETSC_COMM_SELF,__LINE__,__func__,__FILE__,"",PETSC_ERR_ARG_WRONG,PETSC_ERROR_INITIAL,"Wrong type of object: Parameter # %d",(arg)); \
  } while (0)
#define PetscValidCharPointer(p,arg) do {if (!(p)) return PetscError(PETSC_COMM_SELF,__LINE__,__func__,__FILE__,"",PETSC_ERR_ARG_NULL,PETSC_ERROR_INITIAL,"Null Pointer: Parameter # %d",(arg));} while (0)
#define PetscValidIntPointer(p,arg)  PetscValidCharPointer(p,arg)
#define PetscCheckSameComm(a,arga,b,argb) do {} while (0)

/* ---- memory ---------------------------------------------------------------------------------------------------- */
PetscErrorCode OracleShimMalloc(size_t,void**);
#define PetscMalloc(bytes,result)   OracleShimMalloc((size_t)(bytes),(void**)(result))
#define PetscNew(A,b)               (OracleShimMalloc(sizeof(A),(void**)(b)) || (memset(*(b),0,sizeof(A)),0))
#define PetscNewLog(o,A,b)          PetscNew(A,b)
#define PetscFree(a)                ((a) ? (free((void*)(a)),(a)=0,0) : 0)
#define PetscMemcpy(a,b,n)          (memcpy((a),(b),(size_t)(n)),0)
#define PetscMemzero(a,n)           (memset((a),0,(size_t)(n)),0)
#define PetscLogFlops(n)            0
PetscErrorCode PetscStrlen(const char*,size_t*);

/* ---- logging / registration ------------------------------------------------------------------------------------ */
PetscErrorCode PetscClassIdRegister(const char*,PetscClassId*);
#define PetscLogEventRegister(name,classid,ev) ((*(ev)) = 0, 0)
#define PetscLogEventBegin(e,a,b,c,d) 0
#define PetscLogEventEnd(e,a,b,c,d)   0
#define PetscRegisterFinalize(f)      0
#define PetscInfo(a,s)                0
PetscErrorCode PetscPrintf(MPI_Comm,const char*,...);

PetscErrorCode PetscFunctionListAdd(MPI_Comm,PetscFunctionList*,const char*,const char*,void (*)(void));
PetscErrorCode PetscFunctionListFind(MPI_Comm,PetscFunctionList,const char*,PetscBool,void (**)(void));
PetscErrorCode PetscFunctionListConcat(const char*,const char*,char*);

/* ---- options (argv-backed; see OracleShimSetArgs) ---------------------------------------------------------------- */
void OracleShimSetArgs(int argc,char **argv);
void OracleShimSetOption(const char *name,const char *value);   /* programmatic insertion (ctypes users) */
void OracleShimClearOptions(void);
#define PetscOptionsBegin(comm,prefix,title,mansec) 0
#define PetscOptionsEnd()  0
#define PetscOptionsHead(s) 0
#define PetscOptionsTail() 0
PetscErrorCode PetscOptionsGetInt(const char*,const char*,PetscInt*,PetscBool*);
PetscErrorCode PetscOptionsGetBool(const char*,const char*,PetscBool*,PetscBool*);
PetscErrorCode PetscOptionsGetEnum(const char*,const char*,const char *const*,PetscEnum*,PetscBool*);
PetscErrorCode PetscOptionsGetString(const char*,const char*,char*,size_t,PetscBool*);
#define PetscOptionsInt(opt,text,man,defaultv,value,set)   PetscOptionsGetInt(NULL,(opt),(value),(set))
#define PetscOptionsBool(opt,text,man,deflt,flg,set)       PetscOptionsGetBool(NULL,(opt),(flg),(set))
#define PetscOptionsEnum(opt,text,man,list,defaultv,value,set) \
  ((*(value) = (PetscEnum)(defaultv)), PetscOptionsGetEnum(NULL,(opt),(list),(value),(set)))
#define PetscOptionsList(opt,ltext,man,list,defaultv,value,len,set) \
  ((void)strncpy((value),(defaultv),(len)), PetscOptionsGetString(NULL,(opt),(value),(len),(set)))

/* ---- viewer ------------------------------------------------------------------------------------------------------ */
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer,const char*,...);
PetscErrorCode PetscViewerASCIIPushTab(PetscViewer);
PetscErrorCode PetscViewerASCIIPopTab(PetscViewer);
PetscErrorCode PetscViewerASCIIGetStdout(MPI_Comm,PetscViewer*);
PetscErrorCode PetscViewerASCIIUseTabs(PetscViewer,PetscBool);

/* ---- Mat (only MatSetValue is referenced, from the constraint builder which the oracle never calls) -------------- */
PetscErrorCode MatSetValue(Mat,PetscInt,PetscInt,PetscScalar,InsertMode);

#endif
