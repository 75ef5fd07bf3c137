/* TEST INFRASTRUCTURE (oracle): run-time half of the PETSc/iMesh stand-in (see petscsys.h in this directory).
 * Provides just enough object/option/viewer plumbing for the reference's dJacobi layer to run; the product never
 * links this file.  Functions that restate tiny pieces of reference *utility* code cite the place they follow.
 */
#include <dohp.h>
#include <dohpjacimpl.h>
#include <dohpsys.h>
#include <ctype.h>

PetscClassId PETSC_VIEWER_CLASSID = 1211;
static PetscClassId next_classid = 1300;

/* ---- errors ------------------------------------------------------------------------------------------------------ */
PetscErrorCode PetscError(MPI_Comm comm,int line,const char *func,const char *file,const char *dir,PetscErrorCode n,PetscErrorType p,const char *mess,...)
{
  va_list ap;
  (void)comm; (void)dir;
  if (p == PETSC_ERROR_INITIAL) {
    fprintf(stderr,"[oracle-shim] error %d: ",n);
    va_start(ap,mess);
    vfprintf(stderr,mess,ap);
    va_end(ap);
    fputc('\n',stderr);
  }
  fprintf(stderr,"[oracle-shim]   at %s() line %d in %s\n",func,line,file);
  return n ? n : 1;
}

/* ---- objects ----------------------------------------------------------------------------------------------------- */
struct _n_PetscFunctionList {
  char *name;
  void (*fn)(void);
  struct _n_PetscFunctionList *next;
};

PetscErrorCode OracleShimMalloc(size_t bytes,void **result)
{
  void *p = NULL;
  if (!bytes) bytes = 1;
  if (posix_memalign(&p,PETSC_MEMALIGN,bytes)) return PETSC_ERR_MEM;
  *result = p;
  return 0;
}

PetscErrorCode OracleShimHeaderCreate(void **obj,size_t bytes,PetscClassId classid,const char *class_name,MPI_Comm comm)
{
  PetscObject o;
  if (OracleShimMalloc(bytes,(void**)&o)) return PETSC_ERR_MEM;
  memset(o,0,bytes);
  o->classid = classid;
  o->comm = comm;
  o->refct = 1;
  o->class_name = class_name;
  *obj = o;
  return 0;
}

PetscErrorCode OracleShimHeaderDestroy(void **obj)
{
  PetscObject o = (PetscObject)*obj;
  if (!o) return 0;
  free(o->type_name);
  free(o->prefix);
  while (o->qlist) {
    PetscFunctionList nx = o->qlist->next;
    free(o->qlist->name);
    free(o->qlist);
    o->qlist = nx;
  }
  free(o);
  *obj = NULL;
  return 0;
}

PetscErrorCode PetscObjectTypeCompare(PetscObject obj,const char *type,PetscBool *same)
{
  if (!obj) {*same = PETSC_FALSE; return 0;}
  if (obj->classid == PETSC_VIEWER_CLASSID) {*same = (PetscBool)!strcmp(type,PETSCVIEWERASCII); return 0;}
  *same = (PetscBool)(obj->type_name && type && !strcmp(obj->type_name,type));
  return 0;
}

PetscErrorCode PetscObjectChangeTypeName(PetscObject obj,const char *type)
{
  free(obj->type_name);
  obj->type_name = type ? strdup(type) : NULL;
  return 0;
}

PetscErrorCode PetscFunctionListAdd(MPI_Comm comm,PetscFunctionList *fl,const char *name,const char *path,void (*fn)(void))
{
  PetscFunctionList e;
  (void)comm; (void)path;
  for (e=*fl; e; e=e->next) if (!strcmp(e->name,name)) {e->fn = fn; return 0;}
  e = malloc(sizeof(*e));
  e->name = strdup(name);
  e->fn = fn;
  e->next = *fl;
  *fl = e;
  return 0;
}

PetscErrorCode PetscFunctionListFind(MPI_Comm comm,PetscFunctionList fl,const char *name,PetscBool searchlibs,void (**fn)(void))
{
  (void)comm; (void)searchlibs;
  *fn = NULL;
  for (; fl; fl=fl->next) if (!strcmp(fl->name,name)) {*fn = fl->fn; break;}
  return 0;
}

PetscErrorCode PetscFunctionListConcat(const char *path,const char *name,char *fullname)
{
  snprintf(fullname,PETSC_MAX_PATH_LEN,"%s%s%s",path ? path : "",path ? ":" : "",name ? name : "");
  return 0;
}

PetscErrorCode OracleShimCompose(PetscObject obj,const char *name,void (*fn)(void))
{ return PetscFunctionListAdd(0,&obj->qlist,name,NULL,fn); }
PetscErrorCode OracleShimQuery(PetscObject obj,const char *name,void (**fn)(void))
{ return PetscFunctionListFind(0,obj->qlist,name,PETSC_FALSE,fn); }

PetscErrorCode PetscClassIdRegister(const char *name,PetscClassId *id)
{ (void)name; *id = next_classid++; return 0; }

PetscErrorCode PetscStrlen(const char *s,size_t *len)
{ *len = s ? strlen(s) : 0; return 0; }

/* ---- options database -------------------------------------------------------------------------------------------- */
#define MAXOPT 128
static struct {char *name,*value;} optdb[MAXOPT];
static int nopt;

void OracleShimClearOptions(void)
{
  for (int i=0; i<nopt; i++) {free(optdb[i].name); free(optdb[i].value);}
  nopt = 0;
}

void OracleShimSetOption(const char *name,const char *value)
{
  for (int i=0; i<nopt; i++) {
    if (!strcmp(optdb[i].name,name)) {
      free(optdb[i].value);
      optdb[i].value = value ? strdup(value) : NULL;
      return;
    }
  }
  if (nopt >= MAXOPT) return;
  optdb[nopt].name = strdup(name);
  optdb[nopt].value = value ? strdup(value) : NULL;
  nopt++;
}

static int looks_like_option(const char *s)
{ return s[0] == '-' && s[1] && !isdigit((unsigned char)s[1]) && s[1] != '.'; }

void OracleShimSetArgs(int argc,char **argv)
{
  for (int i=1; i<argc; i++) {
    if (!looks_like_option(argv[i])) continue;
    if (i+1 < argc && !looks_like_option(argv[i+1])) {OracleShimSetOption(argv[i],argv[i+1]); i++;}
    else OracleShimSetOption(argv[i],NULL);
  }
}

static const char *findopt(const char *name,int *found)
{
  *found = 0;
  for (int i=0; i<nopt; i++) if (!strcmp(optdb[i].name,name)) {*found = 1; return optdb[i].value;}
  return NULL;
}

PetscErrorCode PetscOptionsGetInt(const char *pre,const char *name,PetscInt *val,PetscBool *set)
{
  int found; const char *v = findopt(name,&found);
  (void)pre;
  if (found && v) *val = atoi(v);
  if (set) *set = (PetscBool)(found && v);
  return 0;
}

PetscErrorCode PetscOptionsGetBool(const char *pre,const char *name,PetscBool *val,PetscBool *set)
{
  int found; const char *v = findopt(name,&found);
  (void)pre;
  if (found) {
    if (!v || !strcmp(v,"1") || !strcasecmp(v,"true") || !strcasecmp(v,"yes") || !strcasecmp(v,"on")) *val = PETSC_TRUE;
    else *val = PETSC_FALSE;
  }
  if (set) *set = (PetscBool)found;
  return 0;
}

/* PETSc enum lists are {"NAME0","NAME1",...,"TypeName","PREFIX_",0}: the count is (#entries - 2). */
PetscErrorCode PetscOptionsGetEnum(const char *pre,const char *name,const char *const *list,PetscEnum *val,PetscBool *set)
{
  int found,n = 0; const char *v = findopt(name,&found);
  (void)pre;
  if (set) *set = PETSC_FALSE;
  if (!found || !v) return 0;
  while (list[n]) n++;
  n -= 2;
  for (int i=0; i<n; i++) {
    if (!strcasecmp(list[i],v)) {*val = i; if (set) *set = PETSC_TRUE; return 0;}
  }
  return PetscError(PETSC_COMM_SELF,__LINE__,__func__,__FILE__,"",PETSC_ERR_USER,PETSC_ERROR_INITIAL,"Unknown enum value '%s' for %s",v,name);
}

PetscErrorCode PetscOptionsGetString(const char *pre,const char *name,char *val,size_t len,PetscBool *set)
{
  int found; const char *v = findopt(name,&found);
  (void)pre;
  if (found && v) {strncpy(val,v,len); val[len-1] = 0;}
  if (set) *set = (PetscBool)(found && v);
  return 0;
}

/* ---- viewer ------------------------------------------------------------------------------------------------------ */
static struct _p_PetscViewer stdout_viewer;
static int usetabs = 1;

PetscViewer OracleShimStdoutViewer(void)
{
  if (!stdout_viewer.fp) {
    stdout_viewer.hdr.classid = PETSC_VIEWER_CLASSID;
    stdout_viewer.hdr.refct = 1;
    stdout_viewer.fp = stdout;
  }
  return &stdout_viewer;
}

PetscErrorCode PetscViewerASCIIGetStdout(MPI_Comm comm,PetscViewer *v)
{ (void)comm; *v = OracleShimStdoutViewer(); return 0; }

/* PETSc's %D (PetscInt) conversion is not known to libc: rewrite to %d. */
static void fix_format(const char *in,char *out,size_t len)
{
  size_t j = 0;
  for (size_t i=0; in[i] && j+1<len; i++) {
    if (in[i] == '%') {
      size_t k = i+1;
      if (in[k] == '%') {out[j++] = '%'; if (j+1<len) out[j++] = '%'; i = k; continue;}
      out[j++] = in[i];
      while (in[k] && strchr("-+ #0123456789.*",in[k]) && j+1<len) out[j++] = in[k++];
      if (in[k] == 'D') {if (j+1<len) out[j++] = 'd'; i = k; continue;}
      i = k-1;
      continue;
    }
    out[j++] = in[i];
  }
  out[j] = 0;
}

PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v,const char *fmt,...)
{
  char f2[8192];
  va_list ap;
  FILE *fp = v && v->fp ? v->fp : stdout;
  if (v && usetabs) for (int i=0; i<v->tab; i++) fputs("  ",fp);
  fix_format(fmt,f2,sizeof f2);
  va_start(ap,fmt);
  vfprintf(fp,f2,ap);
  va_end(ap);
  return 0;
}
PetscErrorCode PetscViewerASCIIPushTab(PetscViewer v) { v->tab++; return 0; }
PetscErrorCode PetscViewerASCIIPopTab(PetscViewer v)  { if (v->tab > 0) v->tab--; return 0; }
PetscErrorCode PetscViewerASCIIUseTabs(PetscViewer v,PetscBool flg) { (void)v; usetabs = flg; return 0; }

PetscErrorCode PetscPrintf(MPI_Comm comm,const char *fmt,...)
{
  char f2[8192];
  va_list ap;
  (void)comm;
  fix_format(fmt,f2,sizeof f2);
  va_start(ap,fmt);
  vfprintf(stdout,f2,ap);
  va_end(ap);
  return 0;
}

PetscErrorCode MatSetValue(Mat m,PetscInt i,PetscInt j,PetscScalar v,InsertMode im)
{
  (void)m; (void)i; (void)j; (void)v; (void)im;
  return PetscError(PETSC_COMM_SELF,__LINE__,__func__,__FILE__,"",PETSC_ERR_SUP,PETSC_ERROR_INITIAL,"Mat is not available in the oracle shim");
}

/* ---- pieces of dohp outside the seven compiled files ------------------------------------------------------------- */

/* The modal basis is out of scope; registration needs the symbol. */
dErr dJacobiCreate_Modal(dJacobi jac)
{ (void)jac; return PetscError(PETSC_COMM_SELF,__LINE__,__func__,__FILE__,"",PETSC_ERR_SUP,PETSC_ERROR_INITIAL,"modal basis not built in the oracle"); }

/* follows src/fs/mesh/interface/mesh.c:46-63 (names only) */
const char *dMeshEntTopologyName(dEntTopology topo)
{
  static const char *const names[] = {"iMesh_POINT","iMesh_LINE_SEGMENT","iMesh_POLYGON","iMesh_TRIANGLE","iMesh_QUADRILATERAL",
                                      "iMesh_POLYHEDRON","iMesh_TETRAHEDRON","iMesh_HEXAHEDRON","iMesh_PRISM","iMesh_PYRAMID",
                                      "iMesh_SEPTAHEDRON","iMesh_ALL_TOPOLOGIES"};
  return ((unsigned)topo < sizeof(names)/sizeof(names[0])) ? names[topo] : "iMesh_TOPOLOGY_OUT_OF_RANGE";
}

/* output format follows src/util/util.c:50-108 */
static dErr table_view(int isreal,dInt m,dInt n,const void *mat,dViewer viewer,const char *format,va_list ap)
{
  char f2[4096],name[4096];
  fix_format(format,f2,sizeof f2);
  vsnprintf(name,sizeof name,f2,ap);
  if (!mat) {PetscViewerASCIIPrintf(viewer,"%10s: null %d*%d\n",name,m,n); return 0;}
  for (dInt i=0; i<m; i++) {
    PetscViewerASCIIPrintf(viewer,"%10s[%2d][%2d:%2d] ",name,i,0,n);
    PetscViewerASCIIUseTabs(viewer,PETSC_FALSE);
    for (dInt j=0; j<n; j++) {
      if (isreal) PetscViewerASCIIPrintf(viewer," % 9.5f",((const dReal*)mat)[i*n+j]);
      else        PetscViewerASCIIPrintf(viewer," %3d",((const dInt*)mat)[i*n+j]);
    }
    PetscViewerASCIIPrintf(viewer,"\n");
    PetscViewerASCIIUseTabs(viewer,PETSC_TRUE);
  }
  return 0;
}
dErr dRealTableView(dInt m,dInt n,const dReal mat[],dViewer viewer,const char *format,...)
{ va_list ap; dErr e; va_start(ap,format); e = table_view(1,m,n,mat,viewer,format,ap); va_end(ap); return e; }
dErr dIntTableView(dInt m,dInt n,const dInt mat[],dViewer viewer,const char *format,...)
{ va_list ap; dErr e; va_start(ap,format); e = table_view(0,m,n,mat,viewer,format,ap); va_end(ap); return e; }

dErr dInitialize(int *argc,char ***argv,const char *file,const char *help)
{
  (void)file; (void)help;
  if (argc && argv) OracleShimSetArgs(*argc,*argv);
  return 0;
}
dErr dFinalize(void) { fflush(stdout); return 0; }

/* ---- sequential Vec (used only by the reference's jacobi ex1 test program) --------------------------------------- */
#include <petscvec.h>
PetscErrorCode VecCreate(MPI_Comm comm,Vec *v)
{ return OracleShimHeaderCreate((void**)v,sizeof(struct _p_Vec),VEC_CLASSID,"Vec",comm); }
PetscErrorCode VecSetSizes(Vec v,PetscInt n,PetscInt N)
{
  if (n < 0) n = N;
  v->n = n;
  free(v->a);
  v->a = calloc((size_t)(n > 0 ? n : 1),sizeof(PetscScalar));
  return v->a ? 0 : PETSC_ERR_MEM;
}
PetscErrorCode VecSetFromOptions(Vec v) { (void)v; return 0; }
PetscErrorCode VecGetArray(Vec v,PetscScalar **a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec v,PetscScalar **a) { (void)v; *a = NULL; return 0; }
PetscErrorCode VecGetSize(Vec v,PetscInt *n) { *n = v->n; return 0; }
PetscErrorCode VecGetLocalSize(Vec v,PetscInt *n) { *n = v->n; return 0; }
PetscErrorCode VecDestroy(Vec *v)
{
  if (!*v) return 0;
  free((*v)->a);
  return OracleShimHeaderDestroy((void**)v);
}
