// eigen.hpp -- block shift-invert Lanczos for the GenEO pencils, replacing SLEPc EPS(arpack)+STSINVERT
// (reference: src/geneo.cpp:626-744 eigenLocalSolve, :746-780 buildEigenSolver, :842-893 eigenLocalProblem).
//
// Problem:  F x = lambda B x  with F factored (block LDL^T) and B an SPD CSR matrix on the same (permuted) pattern.
// Operator T = F^-1 B is self-adjoint in the B inner product; its largest eigenvalues theta = 1/lambda are the
// smallest lambda (tau problem: F = A_neu, B = D A_dir D or A_rob).  For the gamma problem of GenEO-2
// ((D A_dir D) x = lambda A_rob x, largest lambda) call it with F = A_rob, B = D A_dir D and invert = false.
// A whole block of right-hand sides goes through the factor per step, so the factor is streamed once per block
// (ARPACK streams it once per vector).
#pragma once
#include <functional>
#include <vector>
#include "kernels.hpp"
#include "ldlt.hpp"

namespace geneo {

// Device buffers of the block Lanczos iteration (basis Q, B Q, work blocks).  Grow-only: one instance serves every
// subdomain of every (re-)setup, so no multi-GB cudaMalloc / cudaFree sits between two eigen-solves.
struct EigWorkspace {
  DevBuf<double> Q, BQ, W, W2, BW, BW2, Xs, dC, Tmp;
  void release() { Tmp.release(); Q.release(); BQ.release(); W.release(); W2.release(); BW.release(); BW2.release(); Xs.release(); dC.release(); }
};

struct EigOptions {
  int block = 8;
  double tol = 1e-4;      // ||T x - theta x||_B <= tol * |theta|
  int maxDim = 0;         // 0: automatic
  bool invert = true;     // report lambda = 1/theta
  EigWorkspace* ws = nullptr;  // optional persistent buffers
  // Lock-step eigen-solves of several pencils (one host thread each) share ONE forest solve per step -- a single factor is
  // latency-bound in the level barriers of the solve kernel, four of them stream.  xsExt / wExt: this pencil's slices (n x bp,
  // ld = bp) of the group's right-hand-side / solution buffers; solve(j0, nr, st): true when the group solve wrote
  // wExt[:, j0:j0+nr] = F^-1 xsExt[:, j0:j0+nr] (ordered after `st`, `st` ordered after it), false = do it yourself;
  // leave(): called once when this pencil stops asking for solves.
  double* xsExt = nullptr;
  double* wExt = nullptr;
  std::function<bool(int j0, int nr, cudaStream_t st)> solve;
  std::function<void()> leave;
};

struct EigResult {
  int nconv = 0;               // converged wanted pairs (== nev on success)
  int steps = 0, dim = 0;
  std::vector<double> lambda;  // nev values (ascending lambda if invert, else descending)
  std::vector<double> resid;   // relative residual estimates
  DevBuf<double> vecs;         // n x nev row-major (ld = nev), permuted (solver) ordering, B-orthonormal
};

// ptr/idx/valB: device CSR of B in the solver ordering.
void block_lanczos(int n, const LdltFactor& F, const int64_t* ptr, const int* idx, const double* valB, int nev,
                   const EigOptions& opt, EigResult& res, cudaStream_t st);

}  // namespace geneo
