// cli.cpp -- geneo4PETSc, the driver of the reference (src/geneo4PETSc.cpp) re-implemented PETSc-free on top of the C ABI of
// include/geneo_b200.h.  Same command line (checkArguments, src/geneo4PETSc.cpp:1396-1500), same INFO: / TIME: lines
// (printIterativeGlobalSolveParameters :899-1017, ...Results :1052-1097, ...Timing :1099-1230) in the order tst/plot.py
// parses, same verbose dumps (PETSc ASCII_COMMON layout of MatView / VecView), same exit codes (0 = converged).
// Differences that cannot be avoided without MPI:  the number of partitions comes from --nbPart N (the reference takes it
// from `mpirun -n`, :604);  the solver names printed after "L1" / "L2" are this library's (ldlt, blocklanczos ldlt), not
// mumps / arpack;  -pc_type bjacobi | mg (PETSc built-ins used as comparison baselines by the test scripts) are rejected.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "geneo_b200.h"

namespace {

struct Options {
  std::string inpFileA, inpLibA, inpLibArg, inpFileB;
  double inpEps = 0.0001;
  bool metisDual = true;
  int addOverlap = 0;
  int verbose = 0;
  bool timing = false, shortRes = false, cmdLine = false, debug = false;
  int nbPart = 1;
  std::string userCmdLine;
  // PETSc options data base (what the reference reads through PCSetFromOptions / KSPSetFromOptions)
  std::string pcType = "geneo", kspType = "gmres";
  double rtol = 1e-5, atol = 1e-50, dtol = 1e5;
  int maxIt = 10000, restart = 30;
  std::vector<std::string> pcArgs;  // -geneo_* and -els2_* forwarded verbatim
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void usage() {
  std::cerr << "\nusage: geneo4PETSc (B200) is an implementation of the GenEO preconditioner on one NVIDIA B200 per process\n\n"
            << "  --help,         print help related to geneo4PETSc\n"
            << "  --inpFileA F,   input file F describing the A matrix: one element per line, a list of n degrees of freedom,\n"
            << "                  optionally followed by \"-\" and a dense row ordered nxn matrix (default matrix: --inpEps)\n"
            << "  --inpEps E,     epsilon used to tune the default elementary matrix (defaults to 0.0001)\n"
            << "  --inpLibA L A,  input provided by a library L (.so) exporting getInput(); A = its arguments, tokens joined by #\n"
            << "                  (L = laplacian | heat selects the built-in generators of tst/laplacian and tst/heat)\n"
            << "  --inpFileB F,   input file F describing the B vector: one degree of freedom per line, optional value (default 1.)\n"
            << "  --metisDual,    partition according to elements (default)\n"
            << "  --metisNodal,   partition according to nodes\n"
            << "  --addOverlap L, add L layers of overlap at each domain borders\n"
            << "  --nbPart N,     number of subdomains (the reference takes it from mpirun -n)\n"
            << "  --verbose V,    V = 1: dumps X, V = 2: dumps A, B and X\n"
            << "  --timing,       print timing\n"
            << "  --shortRes,     print short result status (makes output stable for test suite checks)\n"
            << "  --cmdLine,      print command line at the end of the log\n"
            << "  --debug F,      accepted and ignored (debug files are not written)\n"
            << "  -pc_type geneo  -ksp_type {gmres,cg}  -ksp_rtol  -ksp_atol  -ksp_max_it  -ksp_gmres_restart\n"
            << "  -geneo_lvl L1,L2  L1 = ASM,RAS,SRAS,ORAS,SORAS  L2 = 0,1,H1,E1,2,H2,E2 (defaults to ASM,1)\n"
            << "  -geneo_optim O  -geneo_tau T  -geneo_gamma G  -geneo_cst  -geneo_cut C  -geneo_no_syl  -geneo_offload\n"
            << "  -els2_eps_tol  -els2_eps_ncv  -els2_eps_block  (block Lanczos / Krylov-Schur eigen-solver)\n\n";
}

// src/geneo4PETSc.cpp:1396-1500.  Returns 0, 1 (error) or -1 (help).
int check_arguments(int argc, char** argv, Options& opt) {
  for (int a = 0; a < argc; a++) opt.userCmdLine += std::string(argv[a]) + " ";
  auto value = [&](int& a, const std::string& clo) -> const char* {
    a++;
    if (a >= argc) { std::cerr << "Error: invalid command line, " << clo << std::endl; return nullptr; }
    return argv[a];
  };
  auto number = [&](const char* s, double& v, const std::string& clo) -> bool {
    std::stringstream ss(s);
    ss >> v;
    if (!ss) { std::cerr << "Error: invalid command line, " << clo << std::endl; return false; }
    return true;
  };
  for (int a = 1; a < argc; a++) {
    const std::string clo = argv[a];
    double d = 0.;
    if (clo == "--help") return -1;
    else if (clo == "--inpFileA") { const char* v = value(a, clo); if (!v) return 1; opt.inpFileA = v; }
    else if (clo == "--inpEps") { const char* v = value(a, clo); if (!v || !number(v, opt.inpEps, clo)) return 1; }
    else if (clo == "--inpLibA") {
      const char* l = value(a, clo); if (!l) return 1;
      const char* g = value(a, clo); if (!g) return 1;
      opt.inpLibA = l; opt.inpLibArg = g;
    }
    else if (clo == "--inpFileB") { const char* v = value(a, clo); if (!v) return 1; opt.inpFileB = v; }
    else if (clo == "--metisDual") opt.metisDual = true;
    else if (clo == "--metisNodal") opt.metisDual = false;
    else if (clo == "--addOverlap") { const char* v = value(a, clo); if (!v || !number(v, d, clo)) return 1; opt.addOverlap = (int)d; }
    else if (clo == "--nbPart" || clo == "-n") { const char* v = value(a, clo); if (!v || !number(v, d, clo) || d < 1) return 1; opt.nbPart = (int)d; }
    else if (clo == "--debug") { const char* v = value(a, clo); if (!v) return 1; opt.debug = true; }
    else if (clo == "--verbose") { const char* v = value(a, clo); if (!v || !number(v, d, clo)) return 1; opt.verbose = (int)d; }
    else if (clo == "--timing") opt.timing = true;
    else if (clo == "--shortRes") opt.shortRes = true;
    else if (clo == "--cmdLine") opt.cmdLine = true;
    // ---- the PETSc options data base ----
    else if (clo == "-pc_type") { const char* v = value(a, clo); if (!v) return 1; opt.pcType = v; }
    else if (clo == "-ksp_type") { const char* v = value(a, clo); if (!v) return 1; opt.kspType = v; }
    else if (clo == "-ksp_rtol") { const char* v = value(a, clo); if (!v || !number(v, opt.rtol, clo)) return 1; }
    else if (clo == "-ksp_atol") { const char* v = value(a, clo); if (!v || !number(v, opt.atol, clo)) return 1; }
    else if (clo == "-ksp_divtol") { const char* v = value(a, clo); if (!v || !number(v, opt.dtol, clo)) return 1; }
    else if (clo == "-ksp_max_it") { const char* v = value(a, clo); if (!v || !number(v, d, clo)) return 1; opt.maxIt = (int)d; }
    else if (clo == "-ksp_gmres_restart") { const char* v = value(a, clo); if (!v || !number(v, d, clo)) return 1; opt.restart = (int)d; }
    else if (clo == "-geneo_cst" || clo == "-geneo_no_syl" || clo == "-geneo_offload" || clo == "-geneo_release_workspace" ||
             clo == "-geneo_kernel_timing" || clo == "-geneo_timing") opt.pcArgs.push_back(clo);
    else if (clo.rfind("-geneo_", 0) == 0 || clo == "-els2_eps_tol" || clo == "-els2_eps_ncv" || clo == "-els2_eps_block") {
      const char* v = value(a, clo); if (!v) return 1;
      opt.pcArgs.push_back(clo); opt.pcArgs.push_back(v);
    }
    else if (clo.size() > 1 && clo[0] == '-' && clo[1] != '-') {
      // any other PETSc-style option of the reference's scripts (-options_left no, -dls1_pc_factor_mat_solver_type mumps,
      // -els2_eps_type arpack, -mat_mumps_cntl_1 ..., -pc_mg_*): accepted, value skipped, no effect -- like an unused
      // entry of the PETSc options data base
      if (a + 1 < argc && !(argv[a + 1][0] == '-' && !(std::isdigit((unsigned char)argv[a + 1][1]) || argv[a + 1][1] == '.'))) a++;
    }
    else { std::cerr << "Error: invalid command line, " << clo << std::endl; return 1; }
  }
  if (opt.timing) opt.pcArgs.push_back("-geneo_timing");  // the per-phase timers of hdr/geneo.hpp:115-123 need their syncs
  if (opt.inpFileA.empty() && opt.inpLibA.empty()) { std::cerr << "Error: no input" << std::endl; return 1; }
  if (!opt.inpFileA.empty() && !opt.inpLibA.empty()) { std::cerr << "Error: several input" << std::endl; return 1; }
  return 0;
}

// PETSc prints a real with %g and appends a "." to a bare integer ("2." , "-1." , "0.5" , "1e-05")
std::string petsc_real(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%g", v);
  std::string s(buf);
  if (s.find_first_of(".einEIN") == std::string::npos) s += ".";
  return s;
}

#define CHK(call)                                                                    \
  do {                                                                               \
    if ((call) != 0) { std::cerr << "Error: " << geneo_last_error() << std::endl; return 1; } \
  } while (0)

// --inpLibA L A (src/geneo4PETSc.cpp:75-96, 1522-1543): the plug-in ABI is C++ (std::string / std::vector by reference)
typedef int (*GetInputFn)(std::string const&, unsigned int&, unsigned int&, std::vector<unsigned int>&, std::vector<unsigned int>&,
                          std::vector<std::vector<double>>&);

int load_input(const Options& opt, geneo_problem_t prob) {
  if (!opt.inpFileA.empty()) { CHK(geneo_problem_read_file(prob, opt.inpFileA.c_str(), opt.inpEps)); return 0; }
  std::string args = opt.inpLibArg;
  for (auto& c : args) if (c == '#') c = ' ';  // the tokens of A are joined by # on the command line
  if (opt.inpLibA == "laplacian" || opt.inpLibA == "heat") {  // built-in O(N) generators
    std::ostringstream eps; eps << " --inpEps " << opt.inpEps;
    if (args.find("--inpEps") == std::string::npos) args += eps.str();
    CHK(geneo_problem_generate(prob, opt.inpLibA.c_str(), args.c_str()));
    return 0;
  }
  void* h = dlopen(opt.inpLibA.c_str(), RTLD_NOW);
  if (!h) { std::cerr << "Error: can not open " << opt.inpLibA << " (" << dlerror() << ")" << std::endl; return 1; }
  GetInputFn fn = reinterpret_cast<GetInputFn>(dlsym(h, "getInput"));
  if (!fn) { std::cerr << "Error: no getInput in " << opt.inpLibA << std::endl; return 1; }
  unsigned int nbElem = 0, nbNode = 0;
  std::vector<unsigned int> elemPtr, elemIdx;
  std::vector<std::vector<double>> elemSubMat;
  if (fn(args, nbElem, nbNode, elemPtr, elemIdx, elemSubMat) != 0) { std::cerr << "Error: getInput KO" << std::endl; return 1; }
  std::vector<double> flat;
  for (auto& m : elemSubMat) flat.insert(flat.end(), m.begin(), m.end());
  CHK(geneo_problem_set_mesh(prob, nbNode, nbElem, elemPtr.data(), elemIdx.data(), flat.data()));
  return 0;
}

// createB (src/geneo4PETSc.cpp:807-862): b = A (1..N)^T, or the file (index [value], default value 1., missing rows 0)
int create_b(const Options& opt, geneo_pc_t pc, int64_t n, std::vector<double>& b) {
  b.assign((size_t)n, 0.);
  if (opt.inpFileB.empty()) { CHK(geneo_make_rhs(pc, b.data())); return 0; }
  std::ifstream inp(opt.inpFileB);
  if (!inp) { std::cerr << "Error: can not open " << opt.inpFileB << std::endl; return 1; }
  std::string line;
  while (std::getline(inp, line)) {
    size_t p = 0;
    while (p < line.size() && std::isspace((unsigned char)line[p])) p++;
    if (p == line.size() || line[p] == '%' || line[p] == '#') continue;
    std::stringstream ss(line.substr(p));
    long idx; ss >> idx;
    if (!ss || idx < 0 || idx >= n) { std::cerr << "Error: can not read " << opt.inpFileB << std::endl; return 1; }
    double val; ss >> val;
    if (!ss) val = 1.;
    b[(size_t)idx] = val;
  }
  return 0;
}

void view_vec(const std::vector<double>& v, int nbPart) {
  printf("Vec Object: %d MPI processes\n  type: %s\n", nbPart, nbPart > 1 ? "mpi" : "seq");
  for (double x : v) printf("%s\n", petsc_real(x).c_str());
}

int view_matis(geneo_problem_t prob, int nbPart) {
  printf("Mat Object: %d MPI processes\n  type: is\n", nbPart);
  for (int s = 0; s < nbPart; s++) {
    int64_t sz[4];
    CHK(geneo_problem_sub_sizes(prob, s, sz));
    std::vector<int64_t> ptr((size_t)sz[0] + 1);
    std::vector<int32_t> idx((size_t)sz[2]);
    std::vector<double> val((size_t)sz[2]);
    CHK(geneo_problem_sub_matrix(prob, s, 0, ptr.data(), idx.data(), val.data()));
    printf("  Mat Object: 1 MPI processes\n    type: seqaij\n");
    for (int64_t r = 0; r < sz[0]; r++) {
      printf("row %lld:", (long long)r);
      for (int64_t q = ptr[r]; q < ptr[r + 1]; q++)
        if (val[q] != 0.) printf(" (%d, %s) ", idx[q], petsc_real(val[q]).c_str());  // ASCII_COMMON skips zeros
      printf("\n");
    }
  }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  const double tStart = now_s();
  Options opt;
  const int rcArgs = check_arguments(argc, argv, opt);
  if (rcArgs != 0) { usage(); return rcArgs == -1 ? 0 : 1; }
  if (opt.pcType != "geneo") {
    std::cerr << "Error: -pc_type " << opt.pcType << " is a PETSc built-in (comparison baseline of the reference's test scripts): "
              << "only -pc_type geneo is implemented here" << std::endl;
    return 1;
  }
  if (opt.kspType != "gmres" && opt.kspType != "cg") { std::cerr << "Error: -ksp_type " << opt.kspType << " is not supported (gmres, cg)" << std::endl; return 1; }

  // ---- partitionAndDecompose (src/geneo4PETSc.cpp:571-641) ----
  geneo_problem_t prob = nullptr;
  CHK(geneo_problem_create(&prob));
  double t0 = now_s();
  if (load_input(opt, prob) != 0) return 1;
  const double readInpTime = now_s() - t0;
  t0 = now_s();
  CHK(geneo_problem_decompose(prob, opt.nbPart, opt.metisDual ? 1 : 0, opt.addOverlap, nullptr, nullptr));
  const double partDecompTime = now_s() - t0;
  int64_t nbNode = 0, nbElem = 0, nbPart = 0, nnz = 0;
  CHK(geneo_problem_sizes(prob, &nbNode, &nbElem, &nbPart, &nnz));

  // ---- solve (src/geneo4PETSc.cpp:1283-1395) ----
  if (opt.verbose >= 2) {
    printf("The matrix A is:\n");
    if (view_matis(prob, (int)nbPart) != 0) return 1;
    printf("\n");
  }
  geneo_pc_t pc = nullptr;
  CHK(geneo_pc_create(&pc));
  {
    std::vector<const char*> av;
    for (auto& s : opt.pcArgs) av.push_back(s.c_str());
    if (geneo_pc_set_from_options(pc, (int)av.size(), av.data()) != 0) {
      std::cerr << "Error: " << geneo_last_error() << std::endl;
      return 1;
    }
  }
  t0 = now_s();
  CHK(geneo_pc_setup(pc, prob));  // createA (operator assembly) + KSPSetUp (the GenEO setup)
  double timers[64] = {0.};
  CHK(geneo_pc_timers(pc, timers, 64));
  const double setupAll = now_s() - t0;
  // geneo_pc_timers: [0..19] the timers of hdr/geneo.hpp:115-123, then 20 symbolic, 21 operator, 22 setup, 23 upload, 24 numeric
  const double createATime = timers[21];
  const double kspSetUpTime = setupAll - createATime;

  std::vector<double> b, x((size_t)nbNode, 0.);
  if (create_b(opt, pc, nbNode, b) != 0) return 1;
  if (opt.verbose >= 2) {
    printf("The vector B is:\n");
    view_vec(b, (int)nbPart);
    printf("\n");
  }
  int64_t out[3] = {0, 0, 0};
  double rnorm = 0.;
  t0 = now_s();
  CHK(geneo_ksp_solve(pc, opt.kspType.c_str(), b.data(), x.data(), opt.rtol, opt.atol, opt.dtol, opt.maxIt, opt.restart, out, &rnorm,
                      nullptr, 0));
  const double kspItsTime = now_s() - t0;
  const int reason = (int)out[1];
  if (opt.verbose >= 1) {
    printf("The solution X is:\n");
    view_vec(x, (int)nbPart);
    printf("\n");
  }

  // ---- printIterativeGlobalSolveParameters (:899-1017) ----
  int64_t info[16];
  double reals[4];
  CHK(geneo_pc_info(pc, info, reals));
  char name[128];
  CHK(geneo_pc_name(pc, name, sizeof name));
  const bool hybrid = info[3] != 0, effHybrid = info[4] != 0, oras = info[5] != 0, offload = info[6] != 0, noSyl = info[7] != 0;
  const int lvl2 = (int)info[2];
  printf("INFO: nb DOFs %lld, nb elements %lld, nnz coefs %lld, nb partitions %lld, overlap %d, metis %s\n", (long long)nbNode,
         (long long)nbElem, (long long)nnz, (long long)nbPart, opt.addOverlap, opt.metisDual ? "dual" : "nodal");
  printf("INFO: %s ksp, eps rel %.1e, eps abs %.1e, max iterations %d\n", opt.kspType.c_str(), opt.rtol, opt.atol, opt.maxIt);
  printf("INFO: %s pc", name);
  if (oras) printf(", optim %.2f", reals[2]);
  if (effHybrid) printf(", initial guess");
  printf(", L1 ldlt %s", hybrid ? "proj-fine-space" : "no-proj-fine-space");
  if (lvl2) {
    printf(", tau %.2f", reals[0]);
    if (lvl2 >= 2) printf(", gamma %.2f", reals[1]);
    if (offload) printf(", offload");
    printf(", L2 blocklanczos ldlt\n");
    if (!opt.shortRes) {
      printf("INFO: setup - ");
      if (!noSyl) printf("estim dimE %lld (local: min %lld, max %lld), ", (long long)info[8], (long long)info[9], (long long)info[10]);
      printf(", real dimE %lld (local: min %lld, max %lld)", (long long)info[11], (long long)info[12], (long long)info[13]);
      printf(", nicolaides %lld\n", (long long)info[14]);
    }
  } else {
    printf("\n");
    if (!opt.shortRes) printf("INFO: setup - none\n");
  }

  // ---- printIterativeGlobalSolveResults (:1052-1097) ----
  printf("INFO: solve - %s", reason >= 0 ? "converged" : "diverged");
  if (!opt.shortRes) {
    std::vector<double> ax((size_t)nbNode);
    CHK(geneo_mult(pc, x.data(), ax.data()));
    double rn = 0., bn = 0.;
    for (int64_t i = 0; i < nbNode; i++) { rn += (ax[i] - b[i]) * (ax[i] - b[i]); bn += b[i] * b[i]; }
    printf(" (%s), %lld iteration(s), residual norm %.10f, || AX - B || / || B || %.10f", geneo_ksp_reason_name(reason),
           (long long)out[0], rnorm, std::sqrt(rn) / std::sqrt(bn));
  }
  printf("\n");

  // ---- printIterativeGlobalSolveTiming (:1099-1230) ----
  if (opt.timing) {
    CHK(geneo_pc_timers(pc, timers, 64));
    printf("\nTIME: read input %.5f s, part / decomp %.5f s, create A %.5f s, solver set up %.5f s, solver iterations %.5f s, solve %.5f s\n",
           readInpTime, partDecompTime, createATime, kspSetUpTime, kspItsTime, kspItsTime + kspSetUpTime);
    // order of geneo_pc_timers (abi.cpp): lvl1SetupMinv, lvl2SetupTauLoc, TauSyl, TauEig, GammaLoc, GammaSyl, GammaEig, lvl2SetupSyl,
    // lvl2SetupEig, lvl2SetupZ, lvl2SetupE, lvl1Apply, Scatter, Minv, Gather, PrjFS, lvl2Apply, Zt, Einv, Z
    printf("      L1       setup: Minv %.5f s\n", timers[0]);
    if (lvl2) {
      printf("      L2       setup: ");
      if (!noSyl) printf("sylvester %.5f s, ", timers[7]);
      printf("eigen solve %.5f s, Z %.5f s, E %.5f s\n", timers[8], timers[9], timers[10]);
      printf("      L2 tau   setup: tau   loc %.5f s", timers[1]);
      if (!noSyl) printf(", sylvester %.5f s", timers[2]);
      printf(", eigen solve %.5f s\n", timers[3]);
      if (lvl2 >= 2) {
        printf("      L2 gamma setup: gamma loc %.5f s", timers[4]);
        if (!noSyl) printf(", sylvester %.5f s", timers[5]);
        printf(", eigen solve %.5f s\n", timers[6]);
      }
    }
    printf("      L1       solve: apply %.5f s - scatter %.5f s, Minv %.5f s, gather %.5f s\n", timers[11], timers[12], timers[13], timers[14]);
    if (hybrid) printf("      L1       solve: prjFS %.5f s\n", timers[15]);
    if (lvl2) printf("      L2       solve: apply %.5f s - Zt %.5f s, Einv %.5f s, Z %.5f s\n", timers[16], timers[17], timers[18], timers[19]);
    printf("TIME: total time %.5f s\n", now_s() - tStart);
  }
  if (opt.cmdLine) printf("\nCMD: mpirun -n %lld %s\n", (long long)nbPart, opt.userCmdLine.c_str());
  fflush(stdout);
  geneo_pc_destroy(pc);
  geneo_problem_destroy(prob);
  return reason >= 0 ? 0 : 1;
}
