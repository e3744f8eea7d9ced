// petsc.h -- a ~400-line STAND-IN for the part of PETSc (>= 3.10) and MPI that the GenEO plug-in boundary touches
// (hdr/geneo_c.h, hdr/geneo.hpp, src/geneo.cpp:2516-2729 of the reference).  TEST INFRASTRUCTURE ONLY: it lets
// geneo4petsc_b200/csrc/petsc_adapter.cpp be compiled and exercised in an image that has neither PETSc nor MPI.
//
//   * "MPI" = threads of one process: stub_run_ranks(P, fn) starts P threads, each one a rank of PETSC_COMM_WORLD;
//     Gatherv / Scatterv / Allgather / Bcast / Allreduce / Barrier are real collectives over a shared barrier.
//   * Mat: MATIS (one local SeqAIJ per rank + a local-to-global mapping) and SeqAIJ with the CSR accessors;
//     Vec: contiguous ownership ranges by rank (PETSc's default layout); IS; ISLocalToGlobalMapping; options database;
//     PC with the PRIVATE layout members a plug-in uses (pc->data, pc->ops->{setup,apply,destroy,setfromoptions}).
// Names, argument orders and error conventions follow PETSc 3.10; nothing here is copied from PETSc.
#ifndef PETSC_STUB_H
#define PETSC_STUB_H
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define PETSC_STUB 1
#define PETSC_EXTERN extern "C" __attribute__((visibility("default")))
typedef int PetscErrorCode;
typedef int PetscInt;
typedef int PetscMPIInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef enum { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER } PetscCopyMode;
#define PETSC_ERR_ARG_WRONG 62
#define PETSC_ERR_ARG_NULL 85
#define PETSC_ERR_SUP 56
#define PETSC_DECIDE (-1)
#define PetscFunctionBegin
#define PetscFunctionReturn(x) return (x)
#define CHKERRQ(ierr) do { if (ierr) return (ierr); } while (0)

// ---- MPI over threads ---------------------------------------------------------------------------------------------
struct StubComm {
  int size = 1;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<const void*> slot;  // one pointer per rank, valid between two barriers
  std::vector<long> islot;
};
typedef StubComm* MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
enum { MPI_INT = 4, MPI_DOUBLE = 8, MPI_LONG_LONG = 9, MPI_CHAR = 1 };
enum { MPI_SUM = 1, MPI_MAX = 2 };
inline int stub_type_size(MPI_Datatype t) { return t == MPI_INT ? 4 : t == MPI_CHAR ? 1 : 8; }
inline StubComm stub_world_storage, stub_self_storage;
inline MPI_Comm PETSC_COMM_WORLD = &stub_world_storage;
inline MPI_Comm PETSC_COMM_SELF = &stub_self_storage;
inline thread_local int stub_rank = 0;
inline int stub_rank_of(MPI_Comm c) { return c == PETSC_COMM_SELF ? 0 : stub_rank; }
inline int MPI_Comm_size(MPI_Comm c, int* s) { *s = c->size; return 0; }
inline int MPI_Comm_rank(MPI_Comm c, int* r) { *r = stub_rank_of(c); return 0; }
inline int MPI_Barrier(MPI_Comm c) {
  if (c->size <= 1) return 0;
  std::unique_lock<std::mutex> lk(c->m);
  const long gen = c->generation;
  if (++c->arrived == c->size) { c->arrived = 0; c->generation++; c->cv.notify_all(); }
  else c->cv.wait(lk, [&] { return c->generation != gen; });
  return 0;
}
// every rank publishes one pointer (+ one integer); after the barrier everyone may read all of them; a second barrier
// (stub_done) ends the exchange
inline void stub_publish(MPI_Comm c, const void* p, long v) {
  { std::lock_guard<std::mutex> lk(c->m); if ((int)c->slot.size() < c->size) { c->slot.resize(c->size); c->islot.resize(c->size); } }
  MPI_Barrier(c);
  c->slot[stub_rank_of(c)] = p; c->islot[stub_rank_of(c)] = v;
  MPI_Barrier(c);
}
inline void stub_done(MPI_Comm c) { MPI_Barrier(c); }
inline int MPI_Gatherv(const void* sb, int sc, MPI_Datatype st, void* rb, const int* rc, const int* displs, MPI_Datatype rt, int root, MPI_Comm c) {
  stub_publish(c, sb, sc);
  if (stub_rank_of(c) == root)
    for (int r = 0; r < c->size; r++) std::memcpy((char*)rb + (size_t)displs[r] * stub_type_size(rt), c->slot[r], (size_t)rc[r] * stub_type_size(rt));
  stub_done(c);
  (void)st;
  return 0;
}
inline int MPI_Scatterv(const void* sb, const int* sc, const int* displs, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, int root, MPI_Comm c) {
  stub_publish(c, sb, 0);
  const int me = stub_rank_of(c);
  const char* src = (const char*)c->slot[root];
  // counts / displacements are significant at the root only: read them through the root's arrays
  stub_done(c);
  stub_publish(c, sc, 0);
  const int* rsc = (const int*)c->slot[root];
  stub_done(c);
  stub_publish(c, displs, 0);
  const int* rdp = (const int*)c->slot[root];
  std::memcpy(rb, src + (size_t)rdp[me] * stub_type_size(st), (size_t)rsc[me] * stub_type_size(st));
  stub_done(c);
  (void)rc; (void)rt;
  return 0;
}
inline int MPI_Allgather(const void* sb, int sc, MPI_Datatype st, void* rb, int rc, MPI_Datatype rt, MPI_Comm c) {
  stub_publish(c, sb, sc);
  for (int r = 0; r < c->size; r++) std::memcpy((char*)rb + (size_t)r * rc * stub_type_size(rt), c->slot[r], (size_t)sc * stub_type_size(st));
  stub_done(c);
  return 0;
}
inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  stub_publish(c, b, n);
  if (stub_rank_of(c) != root) std::memcpy(b, c->slot[root], (size_t)n * stub_type_size(t));
  stub_done(c);
  return 0;
}
inline int MPI_Allreduce(const void* sb, void* rb, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  stub_publish(c, sb, n);
  if (t == MPI_DOUBLE) {
    std::vector<double> acc(n);
    for (int i = 0; i < n; i++) {
      double v = ((const double*)c->slot[0])[i];
      for (int r = 1; r < c->size; r++) { const double w = ((const double*)c->slot[r])[i]; v = op == MPI_SUM ? v + w : (w > v ? w : v); }
      acc[i] = v;
    }
    stub_done(c);
    std::memcpy(rb, acc.data(), sizeof(double) * n);
  } else {
    std::vector<int> acc(n);
    for (int i = 0; i < n; i++) {
      int v = ((const int*)c->slot[0])[i];
      for (int r = 1; r < c->size; r++) { const int w = ((const int*)c->slot[r])[i]; v = op == MPI_SUM ? v + w : (w > v ? w : v); }
      acc[i] = v;
    }
    stub_done(c);
    std::memcpy(rb, acc.data(), sizeof(int) * n);
  }
  return 0;
}
// run fn(rank) on P threads = P ranks of PETSC_COMM_WORLD
inline void stub_run_ranks(int P, const std::function<void(int)>& fn) {
  PETSC_COMM_WORLD->size = P;
  PETSC_COMM_WORLD->arrived = 0;
  std::vector<std::thread> th;
  for (int r = 0; r < P; r++) th.emplace_back([&, r] { stub_rank = r; fn(r); });
  for (auto& t : th) t.join();
  PETSC_COMM_WORLD->size = 1;
}

// ---- objects ---------------------------------------------------------------------------------------------------------
struct _p_PetscObject { int refct = 1; MPI_Comm comm = PETSC_COMM_SELF; virtual ~_p_PetscObject() {} };
typedef _p_PetscObject* PetscObject;
inline MPI_Comm PetscObjectComm(PetscObject o) { return o->comm; }
inline PetscErrorCode PetscObjectReference(PetscObject o) { if (o) o->refct++; return 0; }
template <class T> inline PetscErrorCode stub_destroy(T** o) { if (o && *o) { if (--(*o)->refct == 0) delete *o; *o = nullptr; } return 0; }

struct _p_IS : _p_PetscObject { std::vector<PetscInt> idx; };
typedef _p_IS* IS;
inline PetscErrorCode ISCreateGeneral(MPI_Comm c, PetscInt n, const PetscInt* idx, PetscCopyMode, IS* is) {
  *is = new _p_IS(); (*is)->comm = c; (*is)->idx.assign(idx, idx + n); return 0;
}
inline PetscErrorCode ISGetLocalSize(IS is, PetscInt* n) { *n = (PetscInt)is->idx.size(); return 0; }
inline PetscErrorCode ISGetIndices(IS is, const PetscInt** p) { *p = is->idx.data(); return 0; }
inline PetscErrorCode ISRestoreIndices(IS, const PetscInt** p) { *p = nullptr; return 0; }
inline PetscErrorCode ISDestroy(IS* is) { return stub_destroy(is); }

struct _p_ISLocalToGlobalMapping : _p_PetscObject { std::vector<PetscInt> idx; };
typedef _p_ISLocalToGlobalMapping* ISLocalToGlobalMapping;
inline PetscErrorCode ISLocalToGlobalMappingCreate(MPI_Comm c, PetscInt bs, PetscInt n, const PetscInt* idx, PetscCopyMode, ISLocalToGlobalMapping* m) {
  (void)bs; *m = new _p_ISLocalToGlobalMapping(); (*m)->comm = c; (*m)->idx.assign(idx, idx + n); return 0;
}
inline PetscErrorCode ISLocalToGlobalMappingGetSize(ISLocalToGlobalMapping m, PetscInt* n) { *n = (PetscInt)m->idx.size(); return 0; }
inline PetscErrorCode ISLocalToGlobalMappingGetIndices(ISLocalToGlobalMapping m, const PetscInt** p) { *p = m->idx.data(); return 0; }
inline PetscErrorCode ISLocalToGlobalMappingRestoreIndices(ISLocalToGlobalMapping, const PetscInt** p) { *p = nullptr; return 0; }
inline PetscErrorCode ISLocalToGlobalMappingDestroy(ISLocalToGlobalMapping* m) { return stub_destroy(m); }

struct _p_Vec : _p_PetscObject { PetscInt N = 0, rstart = 0; std::vector<PetscScalar> a; };
typedef _p_Vec* Vec;
inline PetscErrorCode VecCreateMPI(MPI_Comm c, PetscInt n, PetscInt N, Vec* v) {
  *v = new _p_Vec(); (*v)->comm = c; (*v)->a.assign(n, 0.);
  std::vector<int> all(c->size);
  int nn = n;
  MPI_Allgather(&nn, 1, MPI_INT, all.data(), 1, MPI_INT, c);
  int tot = 0, me = stub_rank_of(c);
  for (int r = 0; r < c->size; r++) { if (r == me) (*v)->rstart = tot; tot += all[r]; }
  (*v)->N = N >= 0 ? N : tot;
  return 0;
}
inline PetscErrorCode VecGetLocalSize(Vec v, PetscInt* n) { *n = (PetscInt)v->a.size(); return 0; }
inline PetscErrorCode VecGetSize(Vec v, PetscInt* n) { *n = v->N; return 0; }
inline PetscErrorCode VecGetOwnershipRange(Vec v, PetscInt* lo, PetscInt* hi) { if (lo) *lo = v->rstart; if (hi) *hi = v->rstart + (PetscInt)v->a.size(); return 0; }
inline PetscErrorCode VecGetArray(Vec v, PetscScalar** p) { *p = v->a.data(); return 0; }
inline PetscErrorCode VecRestoreArray(Vec, PetscScalar** p) { *p = nullptr; return 0; }
inline PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar** p) { *p = v->a.data(); return 0; }
inline PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar** p) { *p = nullptr; return 0; }
inline PetscErrorCode VecSet(Vec v, PetscScalar s) { for (auto& x : v->a) x = s; return 0; }
inline PetscErrorCode VecDestroy(Vec* v) { return stub_destroy(v); }

struct _p_Mat : _p_PetscObject {
  std::string type;                // "seqaij" | "is"
  PetscInt m = 0, n = 0, M = 0, N = 0;
  std::vector<PetscInt> ia, ja;    // seqaij
  std::vector<PetscScalar> va;
  _p_Mat* local = nullptr;         // is
  ISLocalToGlobalMapping rmap = nullptr, cmap = nullptr;
  ~_p_Mat() override { if (local && --local->refct == 0) delete local; }
};
typedef _p_Mat* Mat;
typedef const char* MatType;
#define MATIS "is"
#define MATSEQAIJ "seqaij"
inline PetscErrorCode MatCreateSeqAIJWithArrays(MPI_Comm c, PetscInt m, PetscInt n, PetscInt* i, PetscInt* j, PetscScalar* a, Mat* A) {
  *A = new _p_Mat(); (*A)->comm = c; (*A)->type = MATSEQAIJ; (*A)->m = (*A)->M = m; (*A)->n = (*A)->N = n;
  (*A)->ia.assign(i, i + m + 1); (*A)->ja.assign(j, j + i[m]); (*A)->va.assign(a, a + i[m]);
  return 0;
}
inline PetscErrorCode MatCreateIS(MPI_Comm c, PetscInt bs, PetscInt m, PetscInt n, PetscInt M, PetscInt N, ISLocalToGlobalMapping rmap,
                                  ISLocalToGlobalMapping cmap, Mat* A) {
  (void)bs; *A = new _p_Mat(); (*A)->comm = c; (*A)->type = MATIS; (*A)->m = m; (*A)->n = n; (*A)->M = M; (*A)->N = N;
  (*A)->rmap = rmap; (*A)->cmap = cmap;
  return 0;
}
inline PetscErrorCode MatISSetLocalMat(Mat A, Mat loc) { if (loc) loc->refct++; if (A->local && --A->local->refct == 0) delete A->local; A->local = loc; return 0; }
inline PetscErrorCode MatISGetLocalMat(Mat A, Mat* loc) { *loc = A->local; return 0; }
inline PetscErrorCode MatGetType(Mat A, MatType* t) { *t = A->type.c_str(); return 0; }
inline PetscErrorCode MatGetSize(Mat A, PetscInt* M, PetscInt* N) { if (M) *M = A->M; if (N) *N = A->N; return 0; }
inline PetscErrorCode MatGetLocalSize(Mat A, PetscInt* m, PetscInt* n) { if (m) *m = A->m; if (n) *n = A->n; return 0; }
inline PetscErrorCode MatGetLocalToGlobalMapping(Mat A, ISLocalToGlobalMapping* r, ISLocalToGlobalMapping* c) { if (r) *r = A->rmap; if (c) *c = A->cmap; return 0; }
inline PetscErrorCode MatGetRowIJ(Mat A, PetscInt shift, PetscBool, PetscBool, PetscInt* n, const PetscInt** ia, const PetscInt** ja, PetscBool* done) {
  if (A->type != MATSEQAIJ || shift != 0) { if (done) *done = PETSC_FALSE; return 0; }
  *n = A->m; *ia = A->ia.data(); *ja = A->ja.data(); if (done) *done = PETSC_TRUE; return 0;
}
inline PetscErrorCode MatRestoreRowIJ(Mat, PetscInt, PetscBool, PetscBool, PetscInt*, const PetscInt** ia, const PetscInt** ja, PetscBool*) { *ia = *ja = nullptr; return 0; }
inline PetscErrorCode MatSeqAIJGetArray(Mat A, PetscScalar** a) { *a = A->va.data(); return 0; }
inline PetscErrorCode MatSeqAIJRestoreArray(Mat, PetscScalar** a) { *a = nullptr; return 0; }
inline PetscErrorCode MatDestroy(Mat* A) { return stub_destroy(A); }

typedef struct _p_VecScatter* VecScatter;  // (only named by geneoContext)
typedef struct _p_KSP* KSP;

// ---- options database --------------------------------------------------------------------------------------------------
typedef struct _p_PetscOptions* PetscOptions;
struct PetscOptionItems { int dummy; };
inline std::map<std::string, std::string>& stub_options() { static std::map<std::string, std::string> m; return m; }
inline PetscErrorCode PetscOptionsSetValue(PetscOptions, const char* name, const char* value) { stub_options()[name] = value ? value : ""; return 0; }
inline PetscErrorCode PetscOptionsClear(PetscOptions) { stub_options().clear(); return 0; }
inline PetscErrorCode PetscOptionsHasName(PetscOptions, const char*, const char* name, PetscBool* set) { *set = stub_options().count(name) ? PETSC_TRUE : PETSC_FALSE; return 0; }
inline PetscErrorCode PetscOptionsGetString(PetscOptions, const char*, const char* name, char* s, size_t len, PetscBool* set) {
  auto it = stub_options().find(name);
  if (it == stub_options().end()) { if (set) *set = PETSC_FALSE; return 0; }
  std::snprintf(s, len, "%s", it->second.c_str());
  if (set) *set = PETSC_TRUE;
  return 0;
}
inline PetscErrorCode PetscPrintf(MPI_Comm c, const char* fmt, ...) {
  if (stub_rank_of(c) != 0) return 0;
  va_list ap; va_start(ap, fmt); std::vprintf(fmt, ap); va_end(ap);
  return 0;
}
inline PetscErrorCode stub_error(MPI_Comm, int code, const char* msg) { std::fprintf(stderr, "[petsc_stub] error %d: %s\n", code, msg); return code; }
#define SETERRQ(comm, code, msg) return stub_error(comm, code, msg)
#define SETERRABORT(comm, code, msg) do { stub_error(comm, code, msg); std::abort(); } while (0)

// ---- PC (public handle + the private layout a plug-in uses, petsc/private/pcimpl.h) -------------------------------------
struct _p_PC;
typedef _p_PC* PC;
struct _PCOps {
  PetscErrorCode (*setup)(PC) = nullptr;
  PetscErrorCode (*apply)(PC, Vec, Vec) = nullptr;
  PetscErrorCode (*destroy)(PC) = nullptr;
  PetscErrorCode (*setfromoptions)(PetscOptionItems*, PC) = nullptr;
};
struct _p_PC : _p_PetscObject {
  _PCOps ops[1];
  void* data = nullptr;
  Mat mat = nullptr, pmat = nullptr;
  int setupcalled = 0;
  std::string type;
};
inline std::map<std::string, PetscErrorCode (*)(PC)>& stub_pc_registry() { static std::map<std::string, PetscErrorCode (*)(PC)> m; return m; }
inline PetscErrorCode PCRegister(const char* name, PetscErrorCode (*create)(PC)) { stub_pc_registry()[name] = create; return 0; }
inline PetscErrorCode PCCreate(MPI_Comm c, PC* pc) { *pc = new _p_PC(); (*pc)->comm = c; return 0; }
inline PetscErrorCode PCSetType(PC pc, const char* type) {
  auto it = stub_pc_registry().find(type);
  if (it == stub_pc_registry().end()) return stub_error(pc->comm, PETSC_ERR_SUP, "unknown PC type");
  pc->type = type;
  return it->second(pc);
}
inline PetscErrorCode PCGetType(PC pc, const char** t) { *t = pc->type.c_str(); return 0; }
inline PetscErrorCode PCSetOperators(PC pc, Mat A, Mat P) { pc->mat = A; pc->pmat = P; return 0; }
inline PetscErrorCode PCGetOperators(PC pc, Mat* A, Mat* P) { if (A) *A = pc->mat; if (P) *P = pc->pmat; return 0; }
inline PetscErrorCode PCSetFromOptions(PC pc) { PetscOptionItems it{0}; return pc->ops->setfromoptions ? pc->ops->setfromoptions(&it, pc) : 0; }
inline PetscErrorCode PCSetUp(PC pc) { if (pc->setupcalled) return 0; PetscErrorCode e = pc->ops->setup ? pc->ops->setup(pc) : 0; if (!e) pc->setupcalled = 1; return e; }
inline PetscErrorCode PCApply(PC pc, Vec x, Vec y) { PetscErrorCode e = PCSetUp(pc); if (e) return e; return pc->ops->apply(pc, x, y); }
inline PetscErrorCode PCDestroy(PC* pc) {
  if (!pc || !*pc) return 0;
  if ((*pc)->ops->destroy) (*pc)->ops->destroy(*pc);
  return stub_destroy(pc);
}
#endif
