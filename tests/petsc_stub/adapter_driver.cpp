// adapter_driver.cpp -- exercises geneo4petsc_b200/csrc/petsc_adapter.cpp the way the reference driver uses the plug-in
// (src/geneo4PETSc.cpp:1328-1369): PCRegister("geneo", createGenEOPC) -> PCSetType -> PCSetFromOptions -> [PCGenEOSetup |
// initGenEOPC] -> PCSetUp -> PCApply, on the PETSc/MPI stand-in of tests/petsc_stub (ranks = threads).  TEST INFRASTRUCTURE.
//
//   adapter_driver cpu            no device needed: registration, option grammar, name strings, geneoContext defaults, usage
//   adapter_driver gpu P SIZE LVL P ranks, 3-D Laplacian SIZE^3 split into P METIS subdomains; y = M^-1 x through the PC callbacks
//                                 against geneo_pc_apply on the same decomposition; per-rank geneoContext fields against
//                                 geneo_pc_sub_info.  Exit code 0 when everything agrees.
#include <geneo.hpp>

#include <cmath>
#include <cstdlib>
#include <iostream>

#include "geneo_b200.h"

#define REQ(c) do { if (!(c)) { std::fprintf(stderr, "FAILED: %s (line %d): %s\n", #c, __LINE__, geneo_last_error()); std::exit(1); } } while (0)

static int run_cpu() {
  REQ(PCRegister("geneo", createGenEOPC) == 0);
  struct Case { const char* lvl; const char* name; bool ras, sras, oras, hybrid, eff; int lvl2; };
  const Case cases[] = {{"ASM,1", "geneo1ASM", false, false, false, false, false, 1},   {"RAS,0", "geneo0RAS", true, false, false, false, false, 0},
                        {"SRAS,H1", "geneo1HSRAS", true, true, false, true, false, 1},  {"ORAS,E1", "geneo1EORAS", true, false, true, true, true, 1},
                        {"SORAS,2", "geneo2SORAS", true, true, true, false, false, 2},  {"SORAS,E2", "geneo2ESORAS", true, true, true, true, true, 2}};
  for (const Case& c : cases) {
    PC pc;
    REQ(PCCreate(PETSC_COMM_WORLD, &pc) == 0);
    REQ(PCSetType(pc, "geneo") == 0);
    geneoContext* g = (geneoContext*)pc->data;
    REQ(g && g->name == "geneo1ASM" && g->lvl1ASM && !g->lvl1RAS && g->lvl2 == 1 && g->tau == 0.1 && g->gamma == 10. && g->cut == -1);  // defaults :2649-2662
    REQ(pc->ops->setup && pc->ops->apply && pc->ops->destroy && pc->ops->setfromoptions);
    PetscOptionsClear(NULL);
    PetscOptionsSetValue(NULL, "-geneo_lvl", c.lvl);
    PetscOptionsSetValue(NULL, "-geneo_tau", "0.25");
    PetscOptionsSetValue(NULL, "-geneo_gamma", "5");
    PetscOptionsSetValue(NULL, "-geneo_cut", "7");
    PetscOptionsSetValue(NULL, "-geneo_no_syl", NULL);
    REQ(PCSetFromOptions(pc) == 0);
    REQ(g->name == c.name);
    REQ(g->lvl1RAS == c.ras && g->lvl1SRAS == c.sras && g->lvl1ORAS == c.oras && g->hybrid == c.hybrid && g->effHybrid == c.eff && g->lvl2 == c.lvl2);
    REQ(g->tau == 0.25 && g->gamma == 5. && g->cut == 7 && g->noSyl);
    REQ(PCDestroy(&pc) == 0);
  }
  {  // bad values are rejected with a PETSc error code, not accepted silently (src/geneo.cpp:2486-2488)
    PC pc;
    PCCreate(PETSC_COMM_WORLD, &pc);
    PCSetType(pc, "geneo");
    PetscOptionsClear(NULL);
    PetscOptionsSetValue(NULL, "-geneo_tau", "1.5");
    REQ(PCSetFromOptions(pc) != 0);
    PetscOptionsClear(NULL);
    PetscOptionsSetValue(NULL, "-geneo_lvl", "FOO,1");
    REQ(PCSetFromOptions(pc) != 0);
    PCDestroy(&pc);
  }
  const std::string u = usageGenEO(false);
  REQ(u.find("-geneo_lvl") != std::string::npos && u.find("-geneo_tau") != std::string::npos && u.find("-geneo_no_syl") != std::string::npos);
  std::printf("adapter cpu ok\n");
  return 0;
}

static int run_gpu(int P, int size, const char* lvl, bool withDir, bool viaInit) {
  REQ(PCRegister("geneo", createGenEOPC) == 0);
  // the decomposition both paths share (driver half of the reference, src/geneo4PETSc.cpp:571-641)
  geneo_problem_t prob;
  REQ(geneo_problem_create(&prob) == 0);
  char args[128];
  std::snprintf(args, sizeof(args), "--dim 3 --size %d --inpEps 0.0001 --kappa 2. lin", size);
  REQ(geneo_problem_generate(prob, "laplacian", args) == 0);
  REQ(geneo_problem_decompose(prob, P, 1, 0, NULL, NULL) == 0);
  int64_t N64 = 0;
  REQ(geneo_problem_sizes(prob, &N64, NULL, NULL, NULL) == 0);
  const int N = (int)N64;
  // direct path
  // GenEO-2 scales tau by the maximal multiplicity (tauLoc = k tau): keep the coarse space a small part of the spectrum
  const bool g2 = std::string(lvl).find('2') != std::string::npos;
  const char* tau = g2 ? "0.1" : "0.3";
  const char* argv[] = {"-geneo_lvl", lvl, "-geneo_tau", tau, "-els2_eps_tol", "1e-10", "-geneo_optim", "0.5"};
  geneo_pc_t dev;
  REQ(geneo_pc_create(&dev) == 0);
  REQ(geneo_pc_set_from_options(dev, 8, argv) == 0);
  REQ(geneo_pc_setup(dev, prob) == 0);
  std::vector<double> x(N), yref(N), y(N, 0.);
  for (int i = 0; i < N; i++) x[i] = std::sin(0.37 * i) + 0.01 * (i % 7);
  REQ(geneo_pc_apply(dev, x.data(), yref.data()) == 0);
  // plug-in path
  PetscOptionsClear(NULL);
  PetscOptionsSetValue(NULL, "-geneo_lvl", lvl);
  PetscOptionsSetValue(NULL, "-geneo_tau", tau);
  PetscOptionsSetValue(NULL, "-els2_eps_tol", "1e-10");
  PetscOptionsSetValue(NULL, "-geneo_optim", "0.5");
  std::vector<int> failures(P, 0);
  std::vector<int> nevs(P, -1);
  std::vector<std::string> names(P);
  stub_run_ranks(P, [&](int r) {
    auto chk = [&](bool ok, int line) { if (!ok) { std::fprintf(stderr, "rank %d: check failed at line %d\n", r, line); failures[r]++; } };
    int64_t sz[4];
    geneo_problem_sub_sizes(prob, r, sz);
    const int n = (int)sz[0];
    std::vector<int32_t> nodes(n), mult(n);
    geneo_problem_sub_nodes(prob, r, nodes.data(), mult.data());
    std::vector<int64_t> p64(n + 1);
    std::vector<int32_t> ci(sz[2]);
    std::vector<double> cv(sz[2]);
    geneo_problem_sub_matrix(prob, r, 0, p64.data(), ci.data(), cv.data());
    std::vector<PetscInt> ia(p64.begin(), p64.end()), ja(ci.begin(), ci.end());
    Mat aloc, adir = NULL, A;
    chk(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, n, n, ia.data(), ja.data(), cv.data(), &aloc) == 0, __LINE__);
    if (withDir) {
      std::vector<int64_t> dp(n + 1);
      std::vector<int32_t> di(sz[3]);
      std::vector<double> dv(sz[3]);
      geneo_problem_sub_matrix(prob, r, 1, dp.data(), di.data(), dv.data());
      std::vector<PetscInt> dia(dp.begin(), dp.end()), dja(di.begin(), di.end());
      chk(MatCreateSeqAIJWithArrays(PETSC_COMM_SELF, n, n, dia.data(), dja.data(), dv.data(), &adir) == 0, __LINE__);
    }
    ISLocalToGlobalMapping map;
    std::vector<PetscInt> l2g(nodes.begin(), nodes.end());
    chk(ISLocalToGlobalMappingCreate(PETSC_COMM_WORLD, 1, n, l2g.data(), PETSC_COPY_VALUES, &map) == 0, __LINE__);
    const int lo = (int)((int64_t)N * r / P), hi = (int)((int64_t)N * (r + 1) / P);  // PETSc's contiguous row blocks
    chk(MatCreateIS(PETSC_COMM_WORLD, 1, hi - lo, hi - lo, N, N, map, map, &A) == 0, __LINE__);
    chk(MatISSetLocalMat(A, aloc) == 0, __LINE__);
    PC pc;
    chk(PCCreate(PETSC_COMM_WORLD, &pc) == 0, __LINE__);
    chk(PCSetType(pc, "geneo") == 0, __LINE__);
    chk(PCSetFromOptions(pc) == 0, __LINE__);
    chk(PCSetOperators(pc, A, A) == 0, __LINE__);
    IS multIS;
    std::vector<PetscInt> m32(mult.begin(), mult.end());
    ISCreateGeneral(PETSC_COMM_SELF, n, m32.data(), PETSC_COPY_VALUES, &multIS);
    std::vector<unsigned> dom(nodes.begin(), nodes.end()), mu(mult.begin(), mult.end());
    std::vector<std::vector<unsigned>> inter(P);
    if (viaInit) chk(initGenEOPC(pc, (unsigned)N, (unsigned)n, map, A, adir, NULL, NULL, &dom, &mu, &inter) == 0, __LINE__);  // C++ API, hdr/geneo.hpp:30-35
    else chk(PCGenEOSetup(pc, adir, multIS, NULL) == 0, __LINE__);                                                           // C API, hdr/geneo_c.h:10
    chk(PCSetUp(pc) == 0, __LINE__);
    Vec vx, vy;
    VecCreateMPI(PETSC_COMM_WORLD, hi - lo, N, &vx);
    VecCreateMPI(PETSC_COMM_WORLD, hi - lo, N, &vy);
    PetscScalar* a;
    VecGetArray(vx, &a);
    for (int i = lo; i < hi; i++) a[i - lo] = x[i];
    VecRestoreArray(vx, &a);
    chk(PCApply(pc, vx, vy) == 0, __LINE__);
    chk(PCApply(pc, vx, vy) == 0, __LINE__);  // x must not have been modified by the first apply
    VecGetArray(vy, &a);
    for (int i = lo; i < hi; i++) y[i] = a[i - lo];
    VecRestoreArray(vy, &a);
    geneoContext* g = (geneoContext*)pc->data;
    nevs[r] = g->realDimELoc;
    names[r] = g->name;
    VecDestroy(&vx); VecDestroy(&vy); ISDestroy(&multIS);
    chk(PCDestroy(&pc) == 0, __LINE__);
    MatDestroy(&A); MatDestroy(&aloc);
    if (adir) MatDestroy(&adir);
    ISLocalToGlobalMappingDestroy(&map);
  });
  int bad = 0;
  for (int r = 0; r < P; r++) bad += failures[r];
  double num = 0., den = 0.;
  for (int i = 0; i < N; i++) { num += (y[i] - yref[i]) * (y[i] - yref[i]); den += yref[i] * yref[i]; }
  const double rel = std::sqrt(num / den);
  char nm[64];
  geneo_pc_name(dev, nm, sizeof(nm));
  for (int r = 0; r < P; r++) {
    int64_t si[8];
    double sr[2];
    geneo_pc_sub_info(dev, r, si, sr);
    if (nevs[r] != (int)si[1]) { std::fprintf(stderr, "rank %d: realDimELoc %d != nev %d\n", r, nevs[r], (int)si[1]); bad++; }
    if (names[r] != nm) { std::fprintf(stderr, "rank %d: name %s != %s\n", r, names[r].c_str(), nm); bad++; }
  }
  std::printf("adapter gpu: %s, %d ranks, N %d, |y - y_direct| / |y_direct| = %.3e, failures %d\n", nm, P, N, rel, bad);
  geneo_pc_destroy(dev);
  geneo_problem_destroy(prob);
  return (bad == 0 && rel < 1e-9) ? 0 : 1;
}

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "cpu";
  if (mode == "cpu") return run_cpu();
  if (mode == "gpu") {
    const int P = argc > 2 ? atoi(argv[2]) : 4, size = argc > 3 ? atoi(argv[3]) : 10;
    const char* lvl = argc > 4 ? argv[4] : "ASM,1";
    const bool withDir = argc > 5 && atoi(argv[5]) != 0, viaInit = argc > 6 && atoi(argv[6]) != 0;
    return run_gpu(P, size, lvl, withDir, viaInit);
  }
  std::fprintf(stderr, "usage: adapter_driver cpu | gpu P SIZE LVL [withDir] [viaInit]\n");
  return 2;
}
