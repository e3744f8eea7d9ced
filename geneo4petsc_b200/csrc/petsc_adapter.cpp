// petsc_adapter.cpp -- the PETSc-facing boundary of the reference, re-exported on top of libgeneob200's C ABI:
//
//     extern "C" PetscErrorCode createGenEOPC(PC);                      hdr/geneo_c.h:9   (callback of PCRegister("geneo", ...),
//                                                                                          src/geneo4PETSc.cpp:1331)
//     extern "C" PetscErrorCode PCGenEOSetup(PC, Mat, IS, IS*);         hdr/geneo_c.h:10  (src/geneo.cpp:2518-2573)
//     PetscErrorCode initGenEOPC(PC&, unsigned const&, ...);            hdr/geneo.hpp:30-35 (src/geneo.cpp:2591-2632)
//     std::string    usageGenEO(bool);                                  hdr/geneo.hpp:41  (src/geneo.cpp:2268-2327)
//     class geneoContext                                                hdr/geneo.hpp:46-138 -- the reference's OWN header is
//                                                                       included (-I <reference>/hdr), so pc->data has the
//                                                                       layout the reference driver reads (src/geneo4PETSc.cpp:
//                                                                       928-986, 1123-1225)
// and the four callbacks PETSc invokes through pc->ops (src/geneo.cpp:2717-2720): setup, apply, destroy, setfromoptions.
//
// Process model.  The reference runs ONE subdomain per MPI rank (src/geneo4PETSc.cpp:604).  Here the ranks of the PC's
// communicator hand their subdomain (local Neumann matrix of the MATIS, local-to-global map, optional Dirichlet matrix)
// to the rank that drives the GPU (rank 0 of the communicator): it builds ONE device-resident preconditioner holding all
// subdomains (geneo_problem_begin_subdomains / _set_subdomain / _end_subdomains, geneo_pc_setup).  apply gathers x to that
// rank, runs the persistent device kernels, scatters y back -- x is never modified (src/geneo.cpp:1965-1967).
// Several GPUs: one communicator (PC) per GPU group, or the native multi-GPU entry points (geneo_pc_setup_dist).
//
// Builds against real PETSc (>= 3.10) + MPI headers, or against the stand-in of tests/petsc_stub (this image has neither).
#include <geneo.hpp>  // the reference's header (pulls petsc.h, petsc/private/pcimpl.h, geneo_c.h)

#include <cstring>
#include <numeric>
#include <sstream>

#include "geneo_b200.h"

#define SETERRABT(msg) SETERRABORT(PETSC_COMM_WORLD, PETSC_ERR_ARG_NULL, msg)

namespace {

// geneoContext is what the driver sees through pc->data; the device-side handles ride behind it.
struct B200Context : public geneoContext {
  geneo_problem_t prob = nullptr;
  geneo_pc_t dev = nullptr;
  std::vector<std::string> argv;       // the -geneo_* / sub-solver options picked up from the PETSc options database
  std::vector<int> counts, displs;     // owned rows per rank of the PETSc vectors (root)
  std::vector<double> xg, yg;          // global staging vectors (root)
  int root = 0;
};

B200Context* ctx_of(PC pc) {
  if (!pc) SETERRABT("GenEO preconditioner is invalid");
  B200Context* c = static_cast<B200Context*>(static_cast<geneoContext*>(pc->data));
  if (!c) SETERRABT("GenEO preconditioner without context");
  return c;
}

PetscErrorCode fail(const char* what) {
  std::string m = std::string(what) + ": " + geneo_last_error();
  SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, m.c_str());
}

void build_name(geneoContext* g) {  // buildGenEOName, src/geneo.cpp:2245-2268
  std::string l1;
  if (g->lvl1ASM) l1 = "ASM";
  if (g->lvl1RAS) l1 = "RAS";
  if (g->lvl1SRAS) l1 = "SRAS";
  if (g->lvl1ORAS) l1 = "ORAS";
  if (g->lvl1SRAS && g->lvl1ORAS) l1 = "SORAS";
  g->name = std::string("geneo") + (g->lvl2 == 0 ? "0" : g->lvl2 == 1 ? "1" : "2") + (g->hybrid ? (g->effHybrid ? "E" : "H") : "") + l1;
}

// options database -> argv for geneo_pc_set_from_options, and the mirrored geneoContext fields (src/geneo.cpp:2329-2514)
PetscErrorCode set_from_options(PetscOptionItems*, PC pc) {
  B200Context* g = ctx_of(pc);
  static const char* with_value[] = {"-geneo_lvl", "-geneo_optim", "-geneo_tau", "-geneo_gamma", "-geneo_cut", "-geneo_dbg", "-geneo_chk",
                                     "-els2_eps_tol", "-els2_eps_block", "-els2_eps_ncv", "-geneo_nb", "-geneo_ordering"};
  static const char* flags[] = {"-geneo_cst", "-geneo_no_syl", "-geneo_offload", "-geneo_timing"};
  g->argv.clear();
  char buf[256];
  for (const char* o : with_value) {
    PetscBool set = PETSC_FALSE;
    PetscErrorCode e = PetscOptionsGetString(NULL, NULL, o, buf, sizeof(buf), &set); CHKERRQ(e);
    if (set) { g->argv.push_back(o); g->argv.push_back(buf); }
  }
  for (const char* o : flags) {
    PetscBool set = PETSC_FALSE;
    PetscErrorCode e = PetscOptionsHasName(NULL, NULL, o, &set); CHKERRQ(e);
    if (set) g->argv.push_back(o);
  }
  // validate with the library's own parser (same grammar and messages on every rank), then mirror into the context
  geneo_pc_t probe = nullptr;
  if (geneo_pc_create(&probe)) return fail("createGenEOPC");
  std::vector<const char*> av;
  for (auto& s : g->argv) av.push_back(s.c_str());
  const int rc = geneo_pc_set_from_options(probe, (int)av.size(), av.data());
  if (rc) { std::string m = geneo_last_error(); geneo_pc_destroy(probe); SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, m.c_str()); }
  int64_t ints[16];
  double reals[4];
  char nm[64];
  geneo_pc_info(probe, ints, reals);
  geneo_pc_name(probe, nm, sizeof(nm));
  geneo_pc_destroy(probe);
  const std::string name(nm);  // geneo{0,1,2}{,H,E}{ASM,RAS,SRAS,ORAS,SORAS}
  const std::string l1 = name.substr(name.find_first_of("ARSO", 6));
  g->lvl1ASM = true;  // the reference never clears it (src/geneo.cpp:2649, 2353-2375)
  g->lvl1RAS = l1 != "ASM";
  g->lvl1SRAS = l1 == "SRAS" || l1 == "SORAS";
  g->lvl1ORAS = l1 == "ORAS" || l1 == "SORAS";
  g->lvl2 = (int)ints[2]; g->hybrid = ints[3] != 0; g->effHybrid = ints[4] != 0;
  g->offload = ints[6] != 0; g->noSyl = ints[7] != 0;
  g->tau = reals[0]; g->gamma = reals[1]; g->optim = reals[2];
  for (size_t i = 0; i + 1 < g->argv.size(); i++) if (g->argv[i] == "-geneo_cut") g->cut = atoi(g->argv[i + 1].c_str());
  for (auto& s : g->argv) if (s == "-geneo_cst") g->cst = true;
  g->name = name;
  return 0;
}

// CSR of a SeqAIJ (MatGetRowIJ / MatSeqAIJGetArray)
PetscErrorCode local_csr(Mat A, std::vector<int64_t>& ptr, std::vector<int32_t>& idx, std::vector<double>& val) {
  PetscInt n = 0;
  const PetscInt *ia = nullptr, *ja = nullptr;
  PetscBool done = PETSC_FALSE;
  PetscErrorCode e = MatGetRowIJ(A, 0, PETSC_FALSE, PETSC_FALSE, &n, &ia, &ja, &done); CHKERRQ(e);
  if (!done) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_SUP, "GenEO: the local matrix of the MATIS must be SeqAIJ");
  PetscScalar* a = nullptr;
  e = MatSeqAIJGetArray(A, &a); CHKERRQ(e);
  ptr.assign(ia, ia + n + 1);
  idx.assign(ja, ja + ia[n]);
  val.assign(a, a + ia[n]);
  e = MatSeqAIJRestoreArray(A, &a); CHKERRQ(e);
  e = MatRestoreRowIJ(A, 0, PETSC_FALSE, PETSC_FALSE, &n, &ia, &ja, &done); CHKERRQ(e);
  return 0;
}

template <class T> MPI_Datatype mpi_type();
template <> MPI_Datatype mpi_type<int32_t>() { return MPI_INT; }
template <> MPI_Datatype mpi_type<int64_t>() { return MPI_LONG_LONG; }
template <> MPI_Datatype mpi_type<double>() { return MPI_DOUBLE; }

// variable-length gather to the root: out[r] = rank r's vector
template <class T>
void gather_to_root(MPI_Comm comm, int root, const std::vector<T>& mine, std::vector<std::vector<T>>& out) {
  int size = 1, rank = 0;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  int n = (int)mine.size();
  std::vector<int> cnt(size), dsp(size + 1, 0);
  MPI_Allgather(&n, 1, MPI_INT, cnt.data(), 1, MPI_INT, comm);
  for (int r = 0; r < size; r++) dsp[r + 1] = dsp[r] + cnt[r];
  std::vector<T> all(rank == root ? (size_t)dsp[size] : 1);
  MPI_Gatherv(mine.data(), n, mpi_type<T>(), all.data(), cnt.data(), dsp.data(), mpi_type<T>(), root, comm);
  out.clear();
  if (rank == root)
    for (int r = 0; r < size; r++) out.emplace_back(all.begin() + dsp[r], all.begin() + dsp[r + 1]);
}

// setUpGenEOPC (src/geneo.cpp:1672-1843): every factorization, eigen-solve, Z, E -- on the device
PetscErrorCode setup(PC pc) {
  B200Context* g = ctx_of(pc);
  if (!g->pcA) SETERRABT("GenEO preconditioner without A (initGenEOPC / PCGenEOSetup not called)");
  MPI_Comm comm = PetscObjectComm((PetscObject)pc);
  int size = 1, rank = 0;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  MatType type;
  PetscErrorCode e = MatGetType(g->pcA, &type); CHKERRQ(e);
  if (std::string(type) != MATIS) SETERRQ(comm, PETSC_ERR_ARG_WRONG, "GenEO: the A matrix must be MatIS");
  Mat loc = nullptr;
  e = MatISGetLocalMat(g->pcA, &loc); CHKERRQ(e);  // A_neu,i (src/geneo.cpp:1714)
  std::vector<int64_t> nptr, dptr;
  std::vector<int32_t> nidx, didx, ids;
  std::vector<double> nval, dval;
  e = local_csr(loc, nptr, nidx, nval); CHKERRQ(e);
  if (g->pcADirLoc) { e = local_csr(g->pcADirLoc, dptr, didx, dval); CHKERRQ(e); }
  {
    const PetscInt* l2g = nullptr;
    PetscInt n = 0;
    e = ISLocalToGlobalMappingGetSize(g->pcMap, &n); CHKERRQ(e);
    e = ISLocalToGlobalMappingGetIndices(g->pcMap, &l2g); CHKERRQ(e);
    ids.assign(l2g, l2g + n);
    e = ISLocalToGlobalMappingRestoreIndices(g->pcMap, &l2g); CHKERRQ(e);
    if ((unsigned)n != g->nbDOFLoc) SETERRQ(comm, PETSC_ERR_ARG_WRONG, "GenEO: local size and local-to-global map differ");
  }
  // the library numbers a subdomain by ascending global id (the reference's std::set order, src/geneo4PETSc.cpp:485-489):
  // a map that is not ascending is sorted here and the matrices are permuted with it
  std::vector<int32_t> ord(ids.size());
  std::iota(ord.begin(), ord.end(), 0);
  bool sorted = true;
  for (size_t i = 1; i < ids.size(); i++) sorted = sorted && ids[i - 1] < ids[i];
  if (!sorted) {
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return ids[a] < ids[b]; });
    std::vector<int32_t> inv(ids.size()), ids2(ids.size());
    for (size_t k = 0; k < ord.size(); k++) { inv[ord[k]] = (int32_t)k; ids2[k] = ids[ord[k]]; }
    auto permute = [&](std::vector<int64_t>& ptr, std::vector<int32_t>& idx, std::vector<double>& val) {
      if (ptr.empty()) return;
      const size_t n = ids.size();
      std::vector<int64_t> p2(n + 1, 0);
      for (size_t k = 0; k < n; k++) p2[k + 1] = p2[k] + (ptr[ord[k] + 1] - ptr[ord[k]]);
      std::vector<int32_t> i2(idx.size());
      std::vector<double> v2(val.size());
      for (size_t k = 0; k < n; k++) {
        int64_t q = p2[k];
        for (int64_t t = ptr[ord[k]]; t < ptr[ord[k] + 1]; t++, q++) { i2[q] = inv[idx[t]]; v2[q] = val[t]; }
      }
      ptr.swap(p2); idx.swap(i2); val.swap(v2);
    };
    permute(nptr, nidx, nval);
    permute(dptr, didx, dval);
    ids.swap(ids2);
  }
  std::vector<std::vector<int64_t>> aNptr, aDptr;
  std::vector<std::vector<int32_t>> aNidx, aDidx, aIds;
  std::vector<std::vector<double>> aNval, aDval;
  gather_to_root(comm, g->root, ids, aIds);
  gather_to_root(comm, g->root, nptr, aNptr); gather_to_root(comm, g->root, nidx, aNidx); gather_to_root(comm, g->root, nval, aNval);
  gather_to_root(comm, g->root, dptr, aDptr); gather_to_root(comm, g->root, didx, aDidx); gather_to_root(comm, g->root, dval, aDval);
  // owned rows of the PETSc vectors: contiguous blocks by rank (MatCreateIS layout, src/geneo4PETSc.cpp:755)
  PetscInt mloc = 0;
  e = MatGetLocalSize(g->pcA, &mloc, NULL); CHKERRQ(e);
  g->counts.assign(size, 0);
  int ml = (int)mloc;
  MPI_Allgather(&ml, 1, MPI_INT, g->counts.data(), 1, MPI_INT, comm);
  g->displs.assign(size + 1, 0);
  for (int r = 0; r < size; r++) g->displs[r + 1] = g->displs[r] + g->counts[r];
  if ((unsigned)g->displs[size] != g->nbDOF) SETERRQ(comm, PETSC_ERR_ARG_WRONG, "GenEO: the local row counts do not add up to the global size");

  int rc = 0;
  double stat[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // estimDimE, realDimE, nicolaides, 5 timers
  std::vector<int> per(3 * (size_t)size, 0);  // per subdomain: estim, nev, nicolaides
  if (rank == g->root) {
    if (g->dev) { geneo_pc_destroy(g->dev); g->dev = nullptr; }
    if (g->prob) { geneo_problem_destroy(g->prob); g->prob = nullptr; }
    rc = geneo_problem_create(&g->prob);
    if (!rc) rc = geneo_problem_begin_subdomains(g->prob, (int64_t)g->nbDOF, size);
    for (int s = 0; s < size && !rc; s++) {
      const bool haveDir = !aDptr[s].empty();
      rc = geneo_problem_set_subdomain(g->prob, s, (int64_t)aIds[s].size(), aIds[s].data(), aNptr[s].data(), aNidx[s].data(), aNval[s].data(),
                                       haveDir ? aDptr[s].data() : nullptr, haveDir ? aDidx[s].data() : nullptr, haveDir ? aDval[s].data() : nullptr);
    }
    if (!rc) rc = geneo_problem_end_subdomains(g->prob);  // multiplicities, intersections, missing A_dir = R A R^T (src/geneo.cpp:1692-1699)
    if (!rc) rc = geneo_pc_create(&g->dev);
    std::vector<const char*> av;
    for (auto& s : g->argv) av.push_back(s.c_str());
    if (!rc) rc = geneo_pc_set_from_options(g->dev, (int)av.size(), av.data());
    if (!rc) rc = geneo_pc_setup(g->dev, g->prob);
    if (!rc) {
      for (int s = 0; s < size; s++) {
        int64_t si[8];
        double sr[2];
        geneo_pc_sub_info(g->dev, s, si, sr);
        per[3 * s] = (int)si[2]; per[3 * s + 1] = (int)si[1]; per[3 * s + 2] = (int)si[3];
      }
      double t[25];
      geneo_pc_timers(g->dev, t, 25);
      stat[3] = t[0]; stat[4] = t[7]; stat[5] = t[8]; stat[6] = t[9]; stat[7] = t[10];
      g->xg.assign(g->nbDOF, 0.); g->yg.assign(g->nbDOF, 0.);
    }
  }
  MPI_Bcast(&rc, 1, MPI_INT, g->root, comm);
  if (rc) { if (rank == g->root) return fail("setUpGenEOPC"); SETERRQ(comm, PETSC_ERR_ARG_WRONG, "setUpGenEOPC failed on the GPU rank"); }
  MPI_Bcast(per.data(), 3 * size, MPI_INT, g->root, comm);
  MPI_Bcast(stat, 8, MPI_DOUBLE, g->root, comm);
  // the fields the driver reduces / prints per rank (src/geneo4PETSc.cpp:971-986, 1123-1225)
  g->estimDimELoc = per[3 * rank]; g->realDimELoc = per[3 * rank + 1]; g->nicolaidesLoc = per[3 * rank + 2];
  g->lvl1SetupMinvTimeLoc = stat[3]; g->lvl2SetupSylTimeLoc = stat[4]; g->lvl2SetupEigTimeLoc = stat[5];
  g->lvl2SetupZTimeLoc = stat[6]; g->lvl2SetupETimeLoc = stat[7];
  g->infoL2 = "blocklanczos ldlt (sm_100a)";
  if (g->pcX0) {  // x0 = Q b for the efficient hybrid variants, 0 otherwise (src/geneo.cpp:1601-1607)
    e = VecSet(g->pcX0, 0.); CHKERRQ(e);
    // (the device Krylov path computes Q b itself; a PETSc KSP on top of this PC starts from the zeroed guess)
  }
  return 0;
}

// applyGenEOPC (src/geneo.cpp:2051-2098): y = M^-1 x.  x is read only.
PetscErrorCode apply(PC pc, Vec x, Vec y) {
  B200Context* g = ctx_of(pc);
  MPI_Comm comm = PetscObjectComm((PetscObject)pc);
  int rank = 0;
  MPI_Comm_rank(comm, &rank);
  const PetscScalar* xa = nullptr;
  PetscScalar* ya = nullptr;
  PetscInt n = 0;
  PetscErrorCode e = VecGetLocalSize(x, &n); CHKERRQ(e);
  if (g->counts.empty() || n != g->counts[rank]) SETERRQ(comm, PETSC_ERR_ARG_WRONG, "GenEO apply: vector layout differs from the operator's");
  e = VecGetArrayRead(x, &xa); CHKERRQ(e);
  MPI_Gatherv(xa, n, MPI_DOUBLE, g->xg.data(), g->counts.data(), g->displs.data(), MPI_DOUBLE, g->root, comm);
  e = VecRestoreArrayRead(x, &xa); CHKERRQ(e);
  int rc = 0;
  if (rank == g->root) rc = geneo_pc_apply(g->dev, g->xg.data(), g->yg.data());
  MPI_Bcast(&rc, 1, MPI_INT, g->root, comm);
  if (rc) SETERRABT("GenEO apply failed on the device");  // solver failure inside the PC is fatal in the reference too (:1428-1431)
  e = VecGetArray(y, &ya); CHKERRQ(e);
  MPI_Scatterv(g->yg.data(), g->counts.data(), g->displs.data(), MPI_DOUBLE, ya, n, MPI_DOUBLE, g->root, comm);
  e = VecRestoreArray(y, &ya); CHKERRQ(e);
  return 0;
}

// destroyGenEOPC (src/geneo.cpp:2217-2243): pcA, pcMap, dofIdxMultLoc, intersectLoc are borrowed; pcADirLoc, pcB, pcX0 are
// reference counted; pcIS belongs to the plug-in
PetscErrorCode destroy(PC pc) {
  B200Context* g = ctx_of(pc);
  if (g->dev) geneo_pc_destroy(g->dev);
  if (g->prob) geneo_problem_destroy(g->prob);
  if (g->pcADirLoc) MatDestroy(&g->pcADirLoc);
  if (g->pcB) VecDestroy(&g->pcB);
  if (g->pcX0) VecDestroy(&g->pcX0);
  if (g->pcIS) ISDestroy(&g->pcIS);
  delete g;
  pc->data = nullptr;
  return 0;
}

}  // namespace

extern "C" {

PETSC_EXTERN PetscErrorCode createGenEOPC(PC pcPC) {
  if (!pcPC) SETERRABT("GenEO preconditioner is invalid");
  B200Context* g = new B200Context();
  // defaults of src/geneo.cpp:2649-2662
  g->lvl1ASM = true; g->lvl1RAS = g->lvl1SRAS = g->lvl1ORAS = false;
  g->lvl2 = 1; g->hybrid = g->effHybrid = false;
  g->optim = 0.; g->tau = 0.1; g->tauLoc = -1.; g->gamma = 10.; g->gammaLoc = -1.;
  g->cst = false; g->cut = -1; g->noSyl = false; g->offload = false;
  g->debug = 0; g->debugBin = g->debugMat = false;
  g->check = g->checkBin = g->checkMat = false;
  g->nbDOF = g->nbDOFLoc = 0;
  g->pcA = NULL; g->pcADirLoc = NULL; g->pcMap = NULL; g->pcIS = NULL; g->pcB = NULL; g->pcX0 = NULL;
  g->dofIdxMultLoc = NULL; g->intersectLoc = NULL;
  g->pcXLoc = NULL; g->pcScatCtx = NULL; g->pcX = NULL; g->pcXOld = NULL; g->pcKSPL1Loc = NULL; g->pcDLoc = NULL;
  g->pcKSPL2 = NULL; g->pcZE2G = NULL; g->pcEEig = NULL; g->pcYEig = NULL;
  g->estimDimELoc = g->realDimELoc = g->nicolaidesLoc = 0;
  g->pcZE2GOff = NULL; g->pcEEigOff = NULL; g->pcKSPL2Off = NULL; g->pcScatCtxOff = NULL; g->pcXOff = NULL; g->pcYEigOff = NULL;
  g->lvl1SetupMinvTimeLoc = 0.;
  g->lvl2SetupTauLocTimeLoc = g->lvl2SetupTauSylTimeLoc = g->lvl2SetupTauEigTimeLoc = 0.;
  g->lvl2SetupGammaLocTimeLoc = g->lvl2SetupGammaSylTimeLoc = g->lvl2SetupGammaEigTimeLoc = 0.;
  g->lvl2SetupSylTimeLoc = g->lvl2SetupEigTimeLoc = g->lvl2SetupZTimeLoc = g->lvl2SetupETimeLoc = 0.;
  g->lvl1ApplyTimeLoc = g->lvl1ApplyScatterTimeLoc = g->lvl1ApplyMinvTimeLoc = g->lvl1ApplyGatherTimeLoc = 0.;
  g->lvl1ApplyPrjFSTimeLoc = g->lvl1ApplyPrjFSZtTimeLoc = g->lvl1ApplyPrjFSEinvTimeLoc = g->lvl1ApplyPrjFSZTimeLoc = 0.;
  g->lvl2ApplyTimeLoc = g->lvl2ApplyZtTimeLoc = g->lvl2ApplyEinvTimeLoc = g->lvl2ApplyZTimeLoc = 0.;
  pcPC->data = static_cast<void*>(static_cast<geneoContext*>(g));
  pcPC->ops->setup = setup;
  pcPC->ops->apply = apply;
  pcPC->ops->destroy = destroy;
  pcPC->ops->setfromoptions = set_from_options;
  build_name(g);
  return 0;
}

PETSC_EXTERN PetscErrorCode PCGenEOSetup(PC pc, Mat pcADirLoc, IS dofMultiplicities, IS* dofIntersections) {
  PetscErrorCode ierr;
  Mat P;
  ISLocalToGlobalMapping rmap, cmap;
  PetscInt n, m, N, M;
  PetscFunctionBegin;
  ierr = PCGetOperators(pc, NULL, &P); CHKERRQ(ierr);
  ierr = MatGetLocalToGlobalMapping(P, &rmap, &cmap); CHKERRQ(ierr);
  if (rmap != cmap) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "Row and column LGMaps must match");
  ierr = MatGetSize(P, &N, &M); CHKERRQ(ierr);
  if (N != M) SETERRQ(PetscObjectComm((PetscObject)pc), PETSC_ERR_ARG_WRONG, "Matrix must be square");
  ierr = ISLocalToGlobalMappingGetSize(rmap, &n); CHKERRQ(ierr);
  // The multiplicities and intersections are RE-DERIVED on the device-owning rank from the gathered local-to-global maps
  // (geneo_problem_end_subdomains); the arguments are only checked for consistency, like the reference does (:2547-2548).
  if (dofMultiplicities) {
    ierr = ISGetLocalSize(dofMultiplicities, &m); CHKERRQ(ierr);
    if (n != m) SETERRQ(PETSC_COMM_SELF, PETSC_ERR_ARG_WRONG, "Mismatch in dof mult size and local size");
  }
  (void)dofIntersections;
  ierr = initGenEOPC(pc, (unsigned)N, (unsigned)n, rmap, P, pcADirLoc, NULL, NULL, NULL, NULL, NULL); CHKERRQ(ierr);
  PetscFunctionReturn(0);
}

}  // extern "C"

PetscErrorCode initGenEOPC(PC& pcPC, unsigned int const& nbDOF, unsigned int const& nbDOFLoc, ISLocalToGlobalMapping const& pcMap,
                           Mat const& pcA, Mat const& pcADirLoc, Vec const& pcB, Vec const& pcX0,
                           vector<unsigned int> const* const dofIdxDomLoc, vector<unsigned int> const* const dofIdxMultLoc,
                           vector<vector<unsigned int>> const* const intersectLoc) {
  PetscErrorCode ierr;
  B200Context* g = ctx_of(pcPC);
  g->nbDOF = nbDOF;
  g->nbDOFLoc = nbDOFLoc;
  g->pcMap = pcMap;   // borrowed
  g->pcA = pcA;       // borrowed
  g->pcB = pcB;
  if (pcADirLoc) { g->pcADirLoc = pcADirLoc; ierr = PetscObjectReference((PetscObject)pcADirLoc); CHKERRQ(ierr); }
  if (pcB) { ierr = PetscObjectReference((PetscObject)pcB); CHKERRQ(ierr); }
  g->pcX0 = pcX0;
  if (pcX0) { ierr = PetscObjectReference((PetscObject)pcX0); CHKERRQ(ierr); }
  g->pcIS = NULL;
  if (dofIdxDomLoc) {
    std::vector<PetscInt> dom(dofIdxDomLoc->begin(), dofIdxDomLoc->end());
    ierr = ISCreateGeneral(PETSC_COMM_WORLD, (PetscInt)nbDOFLoc, dom.data(), PETSC_COPY_VALUES, &g->pcIS); CHKERRQ(ierr);
  }
  g->dofIdxMultLoc = dofIdxMultLoc;  // borrowed (kept for the driver; the device side re-derives them)
  g->intersectLoc = intersectLoc;
  return 0;
}

string usageGenEO(bool const petscPrintf) {
  std::stringstream msg;
  msg << "\nusage: GenEO two-level Schwarz preconditioner, B200-native build (libgeneob200: block LDL^T on FP64 tensor cores,\n"
         "       block Lanczos eigen-solver, persistent streaming triangular solves) behind the geneo4PETSc plug-in surface\n\n"
         "  -geneo_lvl L1,L2 L1 = ASM | RAS | SRAS | ORAS | SORAS (level 1)\n"
         "                   L2 = 0 (one level) | 1 | H1 | E1 (GenEO-1: additive, hybrid, efficient hybrid) | 2 | H2 | E2 (GenEO-2)\n"
         "  -geneo_optim A   Robin parameter of ORAS / SORAS: robin = dirichlet + A * neumann on the border (default 0.)\n"
         "  -geneo_tau T     threshold of the GenEO eigenproblem, 0 < T < 1 (default 0.1)\n"
         "  -geneo_gamma G   threshold of the second GenEO-2 eigenproblem, G > 1 (default 10.)\n"
         "  -geneo_cst       GenEO-2: no local variation of tau and gamma\n"
         "  -geneo_cut C     at most C eigenvectors per subdomain and eigenproblem\n"
         "  -geneo_no_syl    do not count the eigenvalues with Sylvester's law of inertia before the eigen-solve\n"
         "  -geneo_offload   accepted (E^-1 is replicated on every GPU: nothing to offload)\n"
         "  -geneo_dbg F,D / -geneo_chk F   debug dumps / additional checks, F = log | bin | mat\n"
         "  sub-solver knobs (stand for the -dls1_/-syl2_/-els2_/-dcs2_ prefixes of the PETSc build):\n"
         "  -els2_eps_tol E  residual tolerance of the eigen-solver (default 1e-4)   -els2_eps_block B   Lanczos block (8 | 16)\n"
         "  -geneo_nb NB     panel width of the LDL^T factorization (default 128)    -geneo_ordering O   1 METIS, 0 natural\n\n";
  if (petscPrintf) PetscPrintf(PETSC_COMM_WORLD, "%s", msg.str().c_str());
  return msg.str();
}
