/* oracle shim: opaque ITAPS handles only (dohptype.h:74-76 needs the three typedefs). */
#ifndef ORACLE_SHIM_IMESH_H
#define ORACLE_SHIM_IMESH_H
typedef struct oracle_iBase_Entity_s     *iBase_EntityHandle;
typedef struct oracle_iBase_Tag_s        *iBase_TagHandle;
typedef struct oracle_iBase_EntitySet_s  *iBase_EntitySetHandle;
typedef struct oracle_iMesh_Instance_s   *iMesh_Instance;
typedef struct oracle_iBase_EntityIter_s *iBase_EntityArrIterator;
typedef struct oracle_iBase_EntityIter1_s *iBase_EntityIterator;
typedef iBase_EntityArrIterator iMesh_EntityArrIterator;
typedef iBase_EntityIterator iMesh_EntityIterator;
enum iBase_ErrorType {iBase_SUCCESS=0};
#endif
