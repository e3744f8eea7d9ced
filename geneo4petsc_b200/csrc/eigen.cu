// eigen.cu -- see eigen.hpp.
#include "eigen.hpp"

#include <algorithm>
#include <cmath>

#include "dense_host.hpp"

namespace geneo {
namespace {

__global__ void k_fill_random(int64_t n, int ld, int ncols, uint64_t seed, double* __restrict__ x) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * ld; t += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % ld);
    uint64_t z = (uint64_t)t * 0x9E3779B97F4A7C15ull + seed;  // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    x[t] = j < ncols ? ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5) : 0.;
  }
}
// x[r*ld + j] = random for c0 <= j < c1 (other columns untouched)
__global__ void k_fill_random_cols(int64_t n, int ld, int c0, int c1, uint64_t seed, double* __restrict__ x) {
  const int nc = c1 - c0;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * nc; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / nc;
    const int j = c0 + (int)(t % nc);
    uint64_t z = (uint64_t)(r * ld + j) * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    x[r * ld + j] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}
// dst[r*ldd + j] = src[r*lds + j], j < ncols ; zero-fills dst columns ncols..ldfill-1
__global__ void k_copy_block(int64_t n, const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd,
                             int ncols, int ldfill) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * ldfill; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / ldfill;
    const int j = (int)(t % ldfill);
    dst[r * ldd + j] = j < ncols ? src[r * lds + j] : 0.;
  }
}
inline int gridn(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8)); }

}  // namespace

void block_lanczos(int n, const LdltFactor& F, const int64_t* ptr, const int* idx, const double* valB, int nev,
                   const EigOptions& opt, EigResult& res, cudaStream_t st) {
  res = EigResult();
  GENEO_CHECK(nev >= 1 && nev <= n, "block_lanczos: bad nev");
  int b, maxDim;
  if (n <= 64) { b = n; maxDim = n; }
  else {
    b = std::max(1, opt.block);
    // basis bound: 4 nev + 8 b columns (measured: 205 wanted pairs converge in a 590-column basis, 10 in 104-136); when it
    // is reached the iteration restarts thickly.  (6 nev before: 2 x 14 GB of basis for a 10^6-row pencil that wants 260 pairs.)
    maxDim = opt.maxDim > 0 ? opt.maxDim : std::max(4 * nev + 8 * b, 128);
    maxDim = std::min(maxDim, n);
    maxDim = std::max(b, maxDim / b * b);
  }
  const int bp = (b + 7) / 8 * 8;
  HostProfScope hpAll("lanczos: all");
  const double tAlloc0 = now_s();
  EigWorkspace localWs;
  EigWorkspace& E = opt.ws ? *opt.ws : localWs;
  auto need = [](DevBuf<double>& d, size_t cnt) { if (d.n < cnt) d.alloc(cnt); };
  need(E.Q, (size_t)n * maxDim); need(E.BQ, (size_t)n * maxDim);
  need(E.W, (size_t)n * bp); need(E.W2, (size_t)n * bp); need(E.BW, (size_t)n * bp); need(E.BW2, (size_t)n * bp);
  need(E.Xs, (size_t)n * bp); need(E.dC, (size_t)maxDim * std::max(bp, nev));
  DevBuf<double>&Q = E.Q, &BQ = E.BQ, &W = E.W, &W2 = E.W2, &BW = E.BW, &BW2 = E.BW2, &Xs = E.Xs, &dC = E.dC;
  double* w = W.p; double* w2 = W2.p; double* bw = BW.p; double* bw2 = BW2.p;
  host_prof_add("lanczos: alloc", now_s() - tAlloc0);

  auto spmmB = [&](const double* X, double* Y) {  // Y = B X for ld = bp blocks
    for (int j0 = 0; j0 < bp; j0 += 8) csr_spmm(n, ptr, idx, valB, X + j0, bp, Y + j0, bp, 8, st);
  };
  std::vector<double> hG((size_t)bp * bp), hR((size_t)b * b), hRinv((size_t)b * b);
  // B-orthonormalise the block in w (n x b, ld bp); on success w, bw hold W R^-1 and B W R^-1, hR holds R.
  auto orth_block = [&]() -> int {
    spmmB(w, bw);
    CUDA_CHECK(cudaMemsetAsync(dC.p, 0, sizeof(double) * b * b, st));
    ts_gram(n, w, bp, b, bw, bp, b, dC.p, b, st);
    CUDA_CHECK(cudaMemcpyAsync(hR.data(), dC.p, sizeof(double) * b * b, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(::geneo::sync_stream(st));
    for (int i = 0; i < b; i++)
      for (int j = i + 1; j < b; j++) hR[i * b + j] = hR[j * b + i] = 0.5 * (hR[i * b + j] + hR[j * b + i]);
    const int bad = chol_upper(b, hR.data(), 1e-13);
    if (bad >= 0) return bad + 1;
    triu_inverse(b, hR.data(), hRinv.data());
    CUDA_CHECK(cudaMemcpyAsync(dC.p, hRinv.data(), sizeof(double) * b * b, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemsetAsync(w2, 0, sizeof(double) * (size_t)n * bp, st));
    CUDA_CHECK(cudaMemsetAsync(bw2, 0, sizeof(double) * (size_t)n * bp, st));
    ts_update(n, w, bp, b, dC.p, b, b, w2, bp, 1., 0., st);
    ts_update(n, bw, bp, b, dC.p, b, b, bw2, bp, 1., 0., st);
    std::swap(w, w2);
    std::swap(bw, bw2);
    return 0;
  };

  k_fill_random<<<GENEO_TICK(gridn((int64_t)n * bp)), 256, 0, st>>>(n, bp, b, 0x1234567ull, w);
  std::vector<double> R0;
  for (int pass = 0; pass < 2; pass++) {  // CholQR2
    const int rc = orth_block();
    GENEO_CHECK(rc == 0, "block_lanczos: the B matrix of the pencil is not positive definite");
  }
  k_copy_block<<<GENEO_TICK(gridn((int64_t)n * b)), 256, 0, st>>>(n, w, bp, Q.p, maxDim, b, b);
  k_copy_block<<<GENEO_TICK(gridn((int64_t)n * b)), 256, 0, st>>>(n, bw, bp, BQ.p, maxDim, b, b);

  std::vector<double> Hm((size_t)maxDim * maxDim, 0.), T, Tsym, Ybot, theta, hC1, hC2, Rfirst;
  bool haveFullRR = false;
  auto full_rr = [&](int dimNow) {  // the whole eigenvector matrix of the projected problem (columns of T), ascending theta
    if (haveFullRR) return;
    HostProfScope hp("lanczos: rayleigh-ritz (full)");
    T = Tsym;
    theta.assign(dimNow, 0.);
    sym_eig(dimNow, T.data(), theta.data());
    haveFullRR = true;
  };
  std::vector<double> ritzVal, ritzRes;
  std::vector<double> Ysel;
  int dim = 0, steps = 0;
  bool done = false;
  bool prefetched = false;  // the solve of this step was already queued behind the previous step's device work
  int nextRR = 0, prevRRstep = -1;
  double prevWorst = 0.;
  double* xs = (opt.solve && opt.xsExt && opt.wExt) ? opt.xsExt : Xs.p;
  bool grouped = xs != Xs.p;  // this pencil's solves ride on the group's forest solve (until a pencil of the group is done)
  struct Leaver { const EigOptions& o; bool on; ~Leaver() { if (on && o.leave) o.leave(); } } leaver{opt, grouped};
  auto launch_solve = [&](int col0) {  // W = F^-1 (B Q[:, col0 : col0+b])
    HostProfScope hp("lanczos: launch_solve");
    k_copy_block<<<GENEO_TICK(gridn((int64_t)n * bp)), 256, 0, st>>>(n, BQ.p + (size_t)col0, maxDim, xs, bp, b, bp);
    for (int j0 = 0; j0 < bp;) {  // 16 right-hand sides per pass over the factor where the block allows it
      const int nr = (bp - j0 >= 16) ? 16 : 8;
      if (grouped && opt.solve(j0, nr, st))
        k_copy_block<<<GENEO_TICK(gridn((int64_t)n * nr)), 256, 0, st>>>(n, opt.wExt + j0, bp, w + j0, bp, nr, nr);
      else {
        grouped = false;  // the group dissolved: from here on this pencil solves alone
        F.solve_permuted(xs, w, bp, j0, nr, st);
      }
      j0 += nr;
    }
  };
  dim = b;           // basis columns, the newest block included
  int restarts = 0;
  while (!done) {
    const int c0 = dim - b;  // the newest block: columns c0 .. dim-1
    if (!prefetched) launch_solve(c0);
    prefetched = false;
    // two passes of block classical Gram-Schmidt against Q[:, 0:dim] in the B inner product
    hC1.assign((size_t)dim * b, 0.);
    for (int pass = 0; pass < 2; pass++) {
      CUDA_CHECK(cudaMemsetAsync(dC.p, 0, sizeof(double) * (size_t)dim * b, st));
      ts_gram(n, BQ.p, maxDim, dim, w, bp, b, dC.p, b, st);
      ts_update(n, Q.p, maxDim, dim, dC.p, b, b, w, bp, -1., 1., st);
      hC2.resize((size_t)dim * b);
      CUDA_CHECK(cudaMemcpyAsync(hC2.data(), dC.p, sizeof(double) * (size_t)dim * b, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(::geneo::sync_stream(st));
      for (size_t t = 0; t < hC1.size(); t++) hC1[t] += hC2[t];
    }
    for (int i = 0; i < dim; i++)
      for (int c = 0; c < b; c++) Hm[(size_t)i * maxDim + c0 + c] = hC1[(size_t)i * b + c];
    const bool room = (dim + b <= maxDim);
    // ---- next block: rank-revealing B-orthonormalisation of W (SVQB + random completion + CholQR) -----------------------
    // As Ritz pairs converge the residual block W loses numerical rank; plain CholQR would break down.  Directions
    // below the threshold are replaced by random vectors (orthogonalised against the basis); the coupling block is
    // then R = Q_{j+1}^T B W_orig, which is what the projected matrix and the residual estimates need.
    bool breakdown = false;
    {
      spmmB(w, bw);  // B W_orig (kept in bw for R)
      CUDA_CHECK(cudaMemsetAsync(dC.p, 0, sizeof(double) * b * b, st));
      ts_gram(n, w, bp, b, bw, bp, b, dC.p, b, st);
      std::vector<double> G((size_t)b * b), Gs((size_t)b * b), sv(b), dsc(b), M((size_t)b * b, 0.);
      CUDA_CHECK(cudaMemcpyAsync(G.data(), dC.p, sizeof(double) * b * b, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(::geneo::sync_stream(st));
      // a column whose B-norm fell below 1e-12 of its norm before the Gram-Schmidt sweep carries no information
      for (int i = 0; i < b; i++) {
        double ref = G[(size_t)i * b + i];
        for (int q = 0; q < dim; q++) ref += hC1[(size_t)q * b + i] * hC1[(size_t)q * b + i];
        const double gii = G[(size_t)i * b + i];
        dsc[i] = (gii > 1e-24 * ref && gii > 0.) ? 1. / std::sqrt(gii) : 0.;
      }
      for (int i = 0; i < b; i++)
        for (int c = 0; c < b; c++) Gs[(size_t)i * b + c] = 0.5 * (G[(size_t)i * b + c] + G[(size_t)c * b + i]) * dsc[i] * dsc[c];
      sym_eig(b, Gs.data(), sv.data());  // ascending
      int r = 0;
      for (int q = b - 1; q >= 0; q--) {  // keep directions with a significant scaled singular value
        if (!(sv[q] > 1e-10)) break;
        for (int i = 0; i < b; i++) M[(size_t)i * b + r] = dsc[i] * Gs[(size_t)i * b + q] / std::sqrt(sv[q]);
        r++;
      }
      CUDA_CHECK(cudaMemcpyAsync(dC.p, M.data(), sizeof(double) * b * b, cudaMemcpyHostToDevice, st));
      CUDA_CHECK(cudaMemsetAsync(w2, 0, sizeof(double) * (size_t)n * bp, st));
      ts_update(n, w, bp, b, dC.p, b, b, w2, bp, 1., 0., st);  // w2 = W M (columns >= r are zero)
      if (r < b) k_fill_random_cols<<<GENEO_TICK(gridn((int64_t)n * (b - r))), 256, 0, st>>>(n, bp, r, b, 0xabcdef12ull + (uint64_t)steps * 7919ull, w2);
      // re-orthogonalise against the basis (the 1/sqrt(s) scaling amplifies the loss of orthogonality)
      for (int pass = 0; pass < (r < b ? 2 : 1); pass++) {
        CUDA_CHECK(cudaMemsetAsync(dC.p, 0, sizeof(double) * (size_t)dim * b, st));
        ts_gram(n, BQ.p, maxDim, dim, w2, bp, b, dC.p, b, st);
        ts_update(n, Q.p, maxDim, dim, dC.p, b, b, w2, bp, -1., 1., st);
      }
      // CholQR on the (now well-conditioned) block
      double* worig_bw = bw;  // B W_orig
      double* wn = w2;        // new block
      int rc = 0;
      for (int pass = 0; pass < 2 && rc == 0; pass++) {
        spmmB(wn, bw2);
        CUDA_CHECK(cudaMemsetAsync(dC.p, 0, sizeof(double) * b * b, st));
        ts_gram(n, wn, bp, b, bw2, bp, b, dC.p, b, st);
        CUDA_CHECK(cudaMemcpyAsync(hR.data(), dC.p, sizeof(double) * b * b, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(::geneo::sync_stream(st));
        for (int i = 0; i < b; i++)
          for (int c = i + 1; c < b; c++) hR[i * b + c] = hR[c * b + i] = 0.5 * (hR[i * b + c] + hR[c * b + i]);
        if (chol_upper(b, hR.data(), 1e-13) >= 0) { rc = 1; break; }
        triu_inverse(b, hR.data(), hRinv.data());
        CUDA_CHECK(cudaMemcpyAsync(dC.p, hRinv.data(), sizeof(double) * b * b, cudaMemcpyHostToDevice, st));
        double* tmp = (wn == w2) ? w : w2;  // out-of-place target (w is free: W_orig only survives through bw)
        CUDA_CHECK(cudaMemsetAsync(tmp, 0, sizeof(double) * (size_t)n * bp, st));
        ts_update(n, wn, bp, b, dC.p, b, b, tmp, bp, 1., 0., st);
        wn = tmp;
      }
      if (rc != 0) breakdown = true;
      else {
        // R = Q_{j+1}^T B W_orig
        CUDA_CHECK(cudaMemsetAsync(dC.p, 0, sizeof(double) * b * b, st));
        ts_gram(n, wn, bp, b, worig_bw, bp, b, dC.p, b, st);
        CUDA_CHECK(cudaMemcpyAsync(hR.data(), dC.p, sizeof(double) * b * b, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(::geneo::sync_stream(st));
        // final block and its B image
        if (wn != w) std::swap(w, w2);  // make w point at the new block
        spmmB(w, bw);
      }
    }
    if (breakdown) std::fill(hR.begin(), hR.end(), 0.);
    if (room)  // T Q_j = Q_{0..j} C_j + Q_{j+1} R : the sub-diagonal block of the projected matrix
      for (int r = 0; r < b; r++)
        for (int c = 0; c < b; c++) Hm[(size_t)(dim + r) * maxDim + c0 + c] = hR[(size_t)r * b + c];
    // The next block is known: queue its copy into the basis and its solve NOW, so that the device streams the factor
    // while the host does the Rayleigh-Ritz below (if that says "converged" the extra solve is simply dropped).
    const bool canContinue = !breakdown && room && dim + b <= n;
    if (canContinue) {
      k_copy_block<<<GENEO_TICK(gridn((int64_t)n * b)), 256, 0, st>>>(n, w, bp, Q.p + (size_t)dim, maxDim, b, b);
      k_copy_block<<<GENEO_TICK(gridn((int64_t)n * b)), 256, 0, st>>>(n, bw, bp, BQ.p + (size_t)dim, maxDim, b, b);
      launch_solve(dim);
      prefetched = true;
    }
    // Rayleigh-Ritz on the symmetrised leading dim x dim block (skipped while the basis is too small for nev pairs to
    // have converged: the O(dim^3) host eigen-solve is the one serial piece of a step)
    const int want = std::min(nev, dim);
    int nconv = 0;
    // The O(dim^3) host eigen-solve is the one serial piece of a step and, once the factor streams at HBM speed, the
    // most expensive one: it runs when the basis can hold nev converged pairs and then only at the steps where the
    // residual of the slowest wanted pair is PREDICTED to reach the tolerance (geometric decay fitted through the last
    // two Rayleigh-Ritz steps, at most 6 steps ahead).  A step too many costs one block solve; the result is the same
    // Rayleigh-Ritz of the final basis either way.
    bool doRR = !(canContinue && dim < nev + b);
    if (doRR && canContinue && steps < nextRR) doRR = false;
    if (doRR) {
    HostProfScope hpRR("lanczos: rayleigh-ritz");
    Tsym.assign((size_t)dim * dim, 0.);
    for (int i = 0; i < dim; i++)
      for (int c = 0; c < dim; c++) Tsym[(size_t)i * dim + c] = 0.5 * (Hm[(size_t)i * maxDim + c] + Hm[(size_t)c * maxDim + i]);
    // Ritz values and the LAST b components of every Ritz vector: all the residual estimate ||R y_bottom|| needs.  The whole
    // eigenvector matrix (7 n^3 flops instead of 4/3 n^3; n reaches 1000+ when a subdomain wants hundreds of pairs) is only
    // computed when the basis is restarted or the iteration ends (full_rr below).
    T = Tsym;
    theta.assign(dim, 0.);
    std::vector<int> botRows(b);
    for (int c = 0; c < b; c++) botRows[c] = dim - b + c;
    Ybot.assign((size_t)b * dim, 0.);
    sym_eig_rows(dim, T.data(), botRows.data(), b, theta.data(), Ybot.data());
    haveFullRR = false;
    ritzVal.assign(want, 0.);
    ritzRes.assign(want, 0.);
    for (int q = 0; q < want; q++) {
      const int col = dim - 1 - q;  // largest theta first
      double s2 = 0.;
      for (int r = 0; r < b; r++) {
        double s = 0.;
        for (int c = 0; c < b; c++) s += hR[(size_t)r * b + c] * Ybot[(size_t)c * dim + col];
        s2 += s * s;
      }
      ritzVal[q] = theta[col];
      ritzRes[q] = std::sqrt(s2) / std::max(std::fabs(theta[col]), 1e-300);
      if (ritzRes[q] <= opt.tol) nconv++;
    }
    {
      double worst = 0.;
      for (int q = 0; q < want; q++) worst = std::max(worst, ritzRes[q]);
      int ahead = 1;
      if (want == nev && prevRRstep >= 0 && worst > opt.tol && worst < prevWorst && prevWorst > 0.) {
        const double rate = std::log(worst / prevWorst) / (double)(steps - prevRRstep);  // < 0 per step
        ahead = (int)std::floor(std::log(opt.tol / worst) / rate);
        ahead = std::max(1, std::min(ahead, 6));
      } else if (want == nev && prevRRstep < 0) ahead = 2;  // second sample for the fit
      prevWorst = worst;
      prevRRstep = steps;
      nextRR = steps + ahead;
    }
    }
    steps++;
    res.nconv = nconv;
    if ((nconv == want && want == nev) || breakdown || dim + b > n) done = true;
    else if (!room) {
      // ---- thick restart (block Krylov-Schur): the basis is full and wanted pairs are still missing.  Keep the leading
      //      Ritz vectors X = Q Y_k (they span the converged and the nearly converged directions) plus the block that was
      //      about to be appended; in the new basis the projected matrix is diag(theta_k) bordered by S = R Y_k[last b rows].
      const int keep = std::min(dim - b, std::max(nev + b, (maxDim - b) / 2));
      if (restarts >= 40 || keep + 2 * b > maxDim) done = true;
      else {
        restarts++;
        full_rr(dim);
        Ysel.assign((size_t)dim * keep, 0.);
        for (int q = 0; q < keep; q++)
          for (int i = 0; i < dim; i++) Ysel[(size_t)i * keep + q] = T[(size_t)i * dim + (dim - 1 - q)];
        need(E.dC, (size_t)maxDim * std::max(std::max(bp, nev), keep));
        need(E.Tmp, (size_t)n * keep);
        CUDA_CHECK(cudaMemcpyAsync(dC.p, Ysel.data(), sizeof(double) * Ysel.size(), cudaMemcpyHostToDevice, st));
        for (int which = 0; which < 2; which++) {  // Q <- Q Y_k, B Q <- B Q Y_k
          DevBuf<double>& src = which == 0 ? Q : BQ;
          CUDA_CHECK(cudaMemsetAsync(E.Tmp.p, 0, sizeof(double) * (size_t)n * keep, st));
          ts_update(n, src.p, maxDim, dim, dC.p, keep, keep, E.Tmp.p, keep, 1., 0., st);
          k_copy_block<<<GENEO_TICK(gridn((int64_t)n * keep)), 256, 0, st>>>(n, E.Tmp.p, keep, src.p, maxDim, keep, keep);
        }
        // the pending block (already B-orthonormal to everything) becomes columns keep .. keep+b-1
        k_copy_block<<<GENEO_TICK(gridn((int64_t)n * b)), 256, 0, st>>>(n, w, bp, Q.p + (size_t)keep, maxDim, b, b);
        k_copy_block<<<GENEO_TICK(gridn((int64_t)n * b)), 256, 0, st>>>(n, bw, bp, BQ.p + (size_t)keep, maxDim, b, b);
        CUDA_CHECK(::geneo::sync_stream(st));
        std::vector<double> Snew((size_t)b * keep, 0.);
        for (int r = 0; r < b; r++)
          for (int q = 0; q < keep; q++) {
            double sv = 0.;
            for (int c = 0; c < b; c++) sv += hR[(size_t)r * b + c] * T[(size_t)(dim - b + c) * dim + (dim - 1 - q)];
            Snew[(size_t)r * keep + q] = sv;
          }
        std::fill(Hm.begin(), Hm.end(), 0.);
        for (int q = 0; q < keep; q++) Hm[(size_t)q * maxDim + q] = theta[dim - 1 - q];
        for (int r = 0; r < b; r++)
          for (int q = 0; q < keep; q++) Hm[(size_t)(keep + r) * maxDim + q] = Hm[(size_t)q * maxDim + keep + r] = Snew[(size_t)r * keep + q];
        dim = keep;  // + b below
        nextRR = 0; prevRRstep = -1; prevWorst = 0.;
      }
    }
    if (!done) dim += b;
    if (done) {
      // Ritz vectors X = Q[:, 0:dim] Y
      full_rr(dim);
      const int got = want;
      Ysel.assign((size_t)dim * got, 0.);
      for (int q = 0; q < got; q++)
        for (int i = 0; i < dim; i++) Ysel[(size_t)i * got + q] = T[(size_t)i * dim + (dim - 1 - q)];
      CUDA_CHECK(cudaMemcpyAsync(dC.p, Ysel.data(), sizeof(double) * Ysel.size(), cudaMemcpyHostToDevice, st));
      res.vecs.alloc((size_t)n * got);
      res.vecs.zero(st);
      ts_update(n, Q.p, maxDim, dim, dC.p, got, got, res.vecs.p, got, 1., 0., st);
      CUDA_CHECK(::geneo::sync_stream(st));
      res.lambda.resize(got);
      res.resid = ritzRes;
      for (int q = 0; q < got; q++) res.lambda[q] = opt.invert ? 1. / ritzVal[q] : ritzVal[q];
      if (breakdown || dim >= n) res.nconv = got;  // exact invariant subspace
    }
  }
  res.steps = steps;
  res.dim = dim;
}

}  // namespace geneo
