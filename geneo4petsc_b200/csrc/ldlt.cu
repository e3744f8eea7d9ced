// ldlt.cu -- numeric block LDL^T multifrontal factorization and level-scheduled solves on sm_100a.
//
// Kernels (all batched over one schedule level, work lists precomputed on the host, see symbolic.hpp):
//   k_assemble      scatter the input CSR values (lower triangle, permuted) into the panels
//   k_extend_add    child update matrix -> parent panel / parent update matrix (relative indices)
//   k_diag_invert   one CTA per front: symmetric sweep inversion of the k x k pivot block held in REGISTERS
//                   (8x8 per thread, 256 threads), pivot signs give the inertia (Sylvester, src/geneo.cpp:452-500)
//   k_copy_panel    W = F21 (unscaled panel, needed by the Schur product)
//   k_panel / k_schur  64x64 tiles of  L21 = W * D^-1   and   U -= L21 * W^T   on the FP64 tensor pipe
//                   (mma.sync.m8n8k4.f64 -- FP64 has no tcgen05 path), smem double-buffered, 4 warps x (32x32)
//   k_fwd / k_dsolve / k_bwd   warp-per-(front,row block) rectangular GEMV sweeps; the factor is streamed exactly
//                   once per sweep with fully coalesced 256-byte warp loads (HBM-bound by design)
#include "ldlt.hpp"

#include <algorithm>

namespace geneo {

// =====================================================================================================================
// DMMA 64x64 tile:  C (+)= A * B^T,  A: M x K (col-major, lda), B: N x K (col-major, ldb), C: M x N (col-major, ldc)
// =====================================================================================================================
namespace {

constexpr int TS = 64;     // tile edge
constexpr int KC = 16;     // K chunk per stage
constexpr int SLD = 68;    // padded smem leading dimension (conflict-free 64-bit fragment loads)
constexpr int GEMM_THREADS = 128;

__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// All 128 threads of the CTA call this with identical arguments.
__device__ void gemm_tile_nt(const double* __restrict__ A, int lda, int M, const double* __restrict__ B, int ldb,
                             int N, int K, double* __restrict__ C, int ldc, int mode, double* sA, double* sB) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 1, wn = warp >> 1;
  const int lr = tid & 63;   // row loaded by this thread
  const int lk = tid >> 6;   // first k column loaded by this thread (0/1), stride 2
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.;

  double ra[8], rb[8];
  const int nch = (K + KC - 1) / KC;
  auto gload = [&](int ch) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int kk = ch * KC + lk + 2 * i;
      ra[i] = (lr < M && kk < K) ? __ldg(A + lr + (size_t)kk * lda) : 0.;
      rb[i] = (lr < N && kk < K) ? __ldg(B + lr + (size_t)kk * ldb) : 0.;
    }
  };
  auto sstore = [&](int buf) {
    double* a = sA + buf * KC * SLD;
    double* b = sB + buf * KC * SLD;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      a[(lk + 2 * i) * SLD + lr] = ra[i];
      b[(lk + 2 * i) * SLD + lr] = rb[i];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  for (int ch = 0; ch < nch; ch++) {
    if (ch + 1 < nch) gload(ch + 1);
    const double* a = sA + (ch & 1) * KC * SLD;
    const double* b = sB + (ch & 1) * KC * SLD;
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; k4++) {
      double fa[4], fb[4];
      const int krow = (k4 * 4 + (lane & 3)) * SLD;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        fa[i] = a[krow + wm * 32 + i * 8 + (lane >> 2)];
        fb[i] = b[krow + wn * 32 + i * 8 + (lane >> 2)];
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma8x8x4(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
    if (ch + 1 < nch) sstore((ch + 1) & 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int row = wm * 32 + i * 8 + (lane >> 2);
      const int col = wn * 32 + j * 8 + 2 * (lane & 3);
      if (row < M) {
#pragma unroll
        for (int e = 0; e < 2; e++)
          if (col + e < N) {
            double* p = C + row + (size_t)(col + e) * ldc;
            if (mode == 0) *p = acc[i][j][e];
            else *p -= acc[i][j][e];
          }
      }
    }
}

__global__ void __launch_bounds__(GEMM_THREADS) k_dgemm_nt(int M, int N, int K, const double* A, int lda,
                                                           const double* B, int ldb, double* C, int ldc, int mode) {
  __shared__ double sA[2 * KC * SLD], sB[2 * KC * SLD];
  const int ti = blockIdx.x, tj = blockIdx.y;
  gemm_tile_nt(A + ti * TS, lda, min(TS, M - ti * TS), B + tj * TS, ldb, min(TS, N - tj * TS), K,
               C + ti * TS + (size_t)tj * TS * ldc, ldc, mode, sA, sB);
}

// =====================================================================================================================
// Factorization kernels
// =====================================================================================================================
__global__ void k_assemble(int64_t cnt, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                           const double* __restrict__ vals, double* __restrict__ L) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cnt; t += (int64_t)gridDim.x * blockDim.x)
    L[dst[t]] = vals[src[t]];
}

constexpr int EADD_COLS = 8;
__global__ void __launch_bounds__(256) k_extend_add(const WorkItem* __restrict__ items, const FrontDev* __restrict__ fr,
                                                    const int* __restrict__ relArr, double* __restrict__ L,
                                                    const double* __restrict__ Uchild, double* __restrict__ Upar) {
  const WorkItem it = items[blockIdx.x];
  const FrontDev F = fr[it.f];
  const FrontDev P = fr[F.parent];
  const int m = F.h - F.k, pk = P.k, ph = P.h, pm = P.h - P.k;
  const double* Uc = Uchild + F.uOff;
  double* Up = Upar + P.uOff;
  double* Lp = L + P.lOff;
  const int* rel = F.relOff >= 0 ? relArr + F.relOff : nullptr;
  const bool atomic = P.nchild > 1;
  const int c1 = min(m, (it.a + 1) * EADD_COLS);
  for (int c = it.a * EADD_COLS; c < c1; c++) {
    const int pc = rel ? rel[c] : c;
    for (int r = c + threadIdx.x; r < m; r += blockDim.x) {
      const int pr = rel ? rel[r] : r;
      const double v = Uc[r + (size_t)c * m];
      double* dst = (pc < pk) ? (Lp + pr + (size_t)pc * ph) : (Up + (pr - pk) + (size_t)(pc - pk) * pm);
      if (atomic) atomicAdd(dst, v);
      else *dst += v;
    }
  }
}

// One CTA (256 threads) per front.  Thread (ti,tj) owns the 8x8 strided sub-block i = ti+16*ii, j = tj+16*jj of the
// (padded to 128x128) pivot block in registers.  Symmetric sweep operator: after sweeping every pivot the block holds
// -F11^-1; the pivots met on the way are the D of the LDL^T factorization (their signs give the inertia).
__global__ void __launch_bounds__(256, 1) k_diag_invert(const WorkItem* __restrict__ items,
                                                        const FrontDev* __restrict__ fr, double* __restrict__ L,
                                                        double pivTol, int* __restrict__ counters) {
  __shared__ double cbuf[2][128];
  const WorkItem it = items[blockIdx.x];
  const FrontDev F = fr[it.f];
  const int k = F.k, h = F.h;
  double* P = L + F.lOff;
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  double a[8][8];
#pragma unroll
  for (int ii = 0; ii < 8; ii++)
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const int i = ti + 16 * ii, j = tj + 16 * jj;
      double v = (i == j) ? 1. : 0.;
      if (i < k && j < k) v = (i >= j) ? P[i + (size_t)j * h] : P[j + (size_t)i * h];
      a[ii][jj] = v;
    }
  // publish column 0
  if (tj == 0) {
#pragma unroll
    for (int ii = 0; ii < 8; ii++) cbuf[0][ti + 16 * ii] = a[ii][0];
  }
  __syncthreads();
  int neg = 0, pert = 0;
  for (int p = 0; p < k; p++) {
    const double* cb = cbuf[p & 1];
    double d = cb[p];
    if (!(fabs(d) >= pivTol)) { d = (d < 0.) ? -pivTol : pivTol; pert++; }
    if (d < 0.) neg++;
    const double rinv = 1. / d;
    double ci[8], cj[8];
#pragma unroll
    for (int q = 0; q < 8; q++) { ci[q] = cb[ti + 16 * q]; cj[q] = cb[tj + 16 * q]; }
#pragma unroll
    for (int ii = 0; ii < 8; ii++) {
      const int i = ti + 16 * ii;
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        const int j = tj + 16 * jj;
        double v;
        if (i == p) v = (j == p) ? -rinv : cj[jj] * rinv;
        else if (j == p) v = ci[ii] * rinv;
        else v = a[ii][jj] - ci[ii] * cj[jj] * rinv;
        a[ii][jj] = v;
      }
    }
    // publish column p+1 for the next sweep
    const int pn = p + 1;
    if (pn < k && tj == (pn & 15)) {
      double* nb = cbuf[pn & 1];
      const int jj = pn >> 4;
#pragma unroll
      for (int ii = 0; ii < 8; ii++) {
        double v = 0.;
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (q == jj) v = a[ii][q];
        nb[ti + 16 * ii] = v;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int ii = 0; ii < 8; ii++)
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const int i = ti + 16 * ii, j = tj + 16 * jj;
      if (i < k && j < k) P[i + (size_t)j * h] = -a[ii][jj];
    }
  if (threadIdx.x == 0 && (neg | pert)) {
    if (neg) atomicAdd(&counters[0], neg);
    if (pert) atomicAdd(&counters[1], pert);
  }
}

constexpr int COPY_ROWS = 1024;
__global__ void __launch_bounds__(256) k_copy_panel(const WorkItem* __restrict__ items, const FrontDev* __restrict__ fr,
                                                    const double* __restrict__ L, double* __restrict__ W) {
  const WorkItem it = items[blockIdx.x];
  const FrontDev F = fr[it.f];
  const int m = F.h - F.k, k = F.k, h = F.h;
  const double* src = L + F.lOff + k;
  double* dst = W + F.wOff;
  const int r1 = min(m, (it.a + 1) * COPY_ROWS);
  for (int c = 0; c < k; c++)
    for (int r = it.a * COPY_ROWS + threadIdx.x; r < r1; r += blockDim.x) dst[r + (size_t)c * m] = src[r + (size_t)c * h];
}

__global__ void __launch_bounds__(GEMM_THREADS) k_panel(const WorkItem* __restrict__ items,
                                                        const FrontDev* __restrict__ fr, double* __restrict__ L,
                                                        const double* __restrict__ W) {
  __shared__ double sA[2 * KC * SLD], sB[2 * KC * SLD];
  const WorkItem it = items[blockIdx.x];
  const FrontDev F = fr[it.f];
  const int m = F.h - F.k, k = F.k, h = F.h;
  double* P = L + F.lOff;
  // L21[ti rows, tj cols] = W[ti rows, :] * Dinv[tj rows, :]^T   (Dinv symmetric)
  gemm_tile_nt(W + F.wOff + it.a * TS, m, min(TS, m - it.a * TS), P + it.b * TS, h, min(TS, k - it.b * TS), k,
               P + k + it.a * TS + (size_t)it.b * TS * h, h, 0, sA, sB);
}

__global__ void __launch_bounds__(GEMM_THREADS) k_schur(const WorkItem* __restrict__ items,
                                                        const FrontDev* __restrict__ fr, const double* __restrict__ L,
                                                        const double* __restrict__ W, double* __restrict__ U) {
  __shared__ double sA[2 * KC * SLD], sB[2 * KC * SLD];
  const WorkItem it = items[blockIdx.x];
  const FrontDev F = fr[it.f];
  const int m = F.h - F.k, k = F.k, h = F.h;
  // U[ti, tj] -= L21[ti rows, :] * W[tj rows, :]^T   (lower triangle of tiles only)
  gemm_tile_nt(L + F.lOff + k + it.a * TS, h, min(TS, m - it.a * TS), W + F.wOff + it.b * TS, m,
               min(TS, m - it.b * TS), k, U + F.uOff + it.a * TS + (size_t)it.b * TS * m, m, 1, sA, sB);
}

// =====================================================================================================================
// Solve kernels: one warp per (front, row block).
// =====================================================================================================================
constexpr int BWD_ROWS = 128;

template <int NR>
__global__ void __launch_bounds__(256) k_fwd(int nitems, const WorkItem* __restrict__ items,
                                             const FrontDev* __restrict__ fr, const int* __restrict__ rowIdx,
                                             const double* __restrict__ L, double* __restrict__ X, int ldx) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nitems) return;
  const WorkItem it = items[w];
  const FrontDev F = fr[it.f];
  const int k = F.k, h = F.h, m = h - k;
  const int r = it.a * 32 + lane;
  const bool ok = r < m;
  const double* Lp = L + F.lOff + k + (ok ? r : 0);
  const int col0 = rowIdx[F.rowOff];
  const double* x1 = X + (size_t)col0 * ldx;
  double acc[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) acc[j] = 0.;
  int c = 0;
  for (; c + 4 <= k; c += 4) {
    double l0 = Lp[(size_t)c * h], l1 = Lp[(size_t)(c + 1) * h], l2 = Lp[(size_t)(c + 2) * h], l3 = Lp[(size_t)(c + 3) * h];
#pragma unroll
    for (int j = 0; j < NR; j++) {
      acc[j] += l0 * x1[(size_t)c * ldx + j];
      acc[j] += l1 * x1[(size_t)(c + 1) * ldx + j];
      acc[j] += l2 * x1[(size_t)(c + 2) * ldx + j];
      acc[j] += l3 * x1[(size_t)(c + 3) * ldx + j];
    }
  }
  for (; c < k; c++) {
    const double l0 = Lp[(size_t)c * h];
#pragma unroll
    for (int j = 0; j < NR; j++) acc[j] += l0 * x1[(size_t)c * ldx + j];
  }
  if (ok) {
    const int row = rowIdx[F.rowOff + k + r];
#pragma unroll
    for (int j = 0; j < NR; j++) atomicAdd(&X[(size_t)row * ldx + j], -acc[j]);
  }
}

template <int NR>
__global__ void __launch_bounds__(256) k_dsolve(int nitems, const WorkItem* __restrict__ items,
                                                const FrontDev* __restrict__ fr, const int* __restrict__ rowIdx,
                                                const double* __restrict__ L, const double* __restrict__ X,
                                                double* __restrict__ Y, int ldx) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nitems) return;
  const WorkItem it = items[w];
  const FrontDev F = fr[it.f];
  const int k = F.k, h = F.h;
  const int r = it.a * 32 + lane;
  const bool ok = r < k;
  const double* Dp = L + F.lOff + (ok ? r : 0);
  const int col0 = rowIdx[F.rowOff];
  const double* x1 = X + (size_t)col0 * ldx;
  double acc[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) acc[j] = 0.;
  for (int c = 0; c < k; c++) {
    const double l0 = Dp[(size_t)c * h];
#pragma unroll
    for (int j = 0; j < NR; j++) acc[j] += l0 * x1[(size_t)c * ldx + j];
  }
  if (ok) {
#pragma unroll
    for (int j = 0; j < NR; j++) Y[(size_t)(col0 + r) * ldx + j] = acc[j];
  }
}

template <int NR>
__global__ void __launch_bounds__(256) k_bwd(int nitems, const WorkItem* __restrict__ items,
                                             const FrontDev* __restrict__ fr, const int* __restrict__ rowIdx,
                                             const double* __restrict__ L, double* __restrict__ Y, int ldx) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= nitems) return;
  const WorkItem it = items[w];
  const FrontDev F = fr[it.f];
  const int k = F.k, h = F.h, m = h - k;
  const int col0 = rowIdx[F.rowOff];
  const double* Lp = L + F.lOff + k;
  double xb[4][NR];
  int rr[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int r = it.a * BWD_ROWS + q * 32 + lane;
    rr[q] = r < m ? r : -1;
    const int row = r < m ? rowIdx[F.rowOff + k + r] : 0;
#pragma unroll
    for (int j = 0; j < NR; j++) xb[q][j] = r < m ? Y[(size_t)row * ldx + j] : 0.;
  }
  for (int c = 0; c < k; c++) {
    double p[NR];
#pragma unroll
    for (int j = 0; j < NR; j++) p[j] = 0.;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double l0 = rr[q] >= 0 ? Lp[rr[q] + (size_t)c * h] : 0.;
#pragma unroll
      for (int j = 0; j < NR; j++) p[j] += l0 * xb[q][j];
    }
#pragma unroll
    for (int j = 0; j < NR; j++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p[j] += __shfl_xor_sync(0xffffffffu, p[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NR; j++) atomicAdd(&Y[(size_t)(col0 + c) * ldx + j], -p[j]);
    }
  }
}

}  // namespace

void dgemm_nt_device(int M, int N, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                     int mode, cudaStream_t st) {
  dim3 grid((M + TS - 1) / TS, (N + TS - 1) / TS);
  k_dgemm_nt<<<grid, GEMM_THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, mode);
  CUDA_CHECK(cudaGetLastError());
}

// =====================================================================================================================
// Plan: symbolic + work lists
// =====================================================================================================================
LdltPlan::LdltPlan(int n, const int64_t* ptr, const int* idx, const SymbolicOptions& opt) {
  symbolic_analyze(n, ptr, idx, opt, sym);
  build_device();
}
LdltPlan::LdltPlan(Symbolic&& s) : sym(std::move(s)) { build_device(); }

void LdltPlan::build_device() {
  const int nf = (int)sym.fronts.size();
  std::vector<FrontDev> fd(nf);
  for (int f = 0; f < nf; f++) {
    const Front& F = sym.fronts[f];
    fd[f] = FrontDev{F.lOff, F.uOff, F.wOff, F.rowOff, F.relOff, F.k, F.h, F.parent, F.nchild};
  }
  std::vector<WorkItem> items;
  auto begin = [&](Range& r) { r.off = (int64_t)items.size(); };
  auto end = [&](Range& r) { r.cnt = (int)((int64_t)items.size() - r.off); };
  const int nl = sym.nlevels;
  eaddItems.resize(nl); diagItems.resize(nl); copyItems.resize(nl); panelItems.resize(nl); schurItems.resize(nl);
  fwdItems.resize(nl); bwdItems.resize(nl);
  levelU.assign(nl, 0);
  for (int l = 0; l < nl; l++) {
    const int* lf = &sym.levelFronts[sym.levelPtr[l]];
    const int cnt = sym.levelPtr[l + 1] - sym.levelPtr[l];
    for (int t = 0; t < cnt; t++) {
      const Front& F = sym.fronts[lf[t]];
      const int64_t m = F.m();
      if (m > 0) levelU[l] = std::max(levelU[l], F.uOff + m * m);
    }
    // extend-add: children are the fronts of level l-1 (their parents are all at level l)
    begin(eaddItems[l]);
    if (l > 0) {
      const int* cf = &sym.levelFronts[sym.levelPtr[l - 1]];
      const int cc = sym.levelPtr[l] - sym.levelPtr[l - 1];
      for (int t = 0; t < cc; t++) {
        const Front& C = sym.fronts[cf[t]];
        for (int cb = 0; cb * EADD_COLS < C.m(); cb++) items.push_back(WorkItem{cf[t], cb, 0});
      }
    }
    end(eaddItems[l]);
    begin(diagItems[l]);
    for (int t = 0; t < cnt; t++) items.push_back(WorkItem{lf[t], 0, 0});
    end(diagItems[l]);
    begin(copyItems[l]);
    for (int t = 0; t < cnt; t++)
      for (int rb = 0; rb * COPY_ROWS < sym.fronts[lf[t]].m(); rb++) items.push_back(WorkItem{lf[t], rb, 0});
    end(copyItems[l]);
    begin(panelItems[l]);
    for (int t = 0; t < cnt; t++) {
      const Front& F = sym.fronts[lf[t]];
      for (int ti = 0; ti * TS < F.m(); ti++)
        for (int tj = 0; tj * TS < F.k; tj++) items.push_back(WorkItem{lf[t], ti, tj});
    }
    end(panelItems[l]);
    begin(schurItems[l]);
    for (int t = 0; t < cnt; t++) {
      const Front& F = sym.fronts[lf[t]];
      for (int ti = 0; ti * TS < F.m(); ti++)
        for (int tj = 0; tj <= ti; tj++) items.push_back(WorkItem{lf[t], ti, tj});
    }
    end(schurItems[l]);
    begin(fwdItems[l]);
    for (int t = 0; t < cnt; t++)
      for (int rb = 0; rb * 32 < sym.fronts[lf[t]].m(); rb++) items.push_back(WorkItem{lf[t], rb, 0});
    end(fwdItems[l]);
    begin(bwdItems[l]);
    for (int t = 0; t < cnt; t++)
      for (int rb = 0; rb * BWD_ROWS < sym.fronts[lf[t]].m(); rb++) items.push_back(WorkItem{lf[t], rb, 0});
    end(bwdItems[l]);
  }
  begin(dsolveItems);
  for (int f = 0; f < nf; f++)
    for (int rb = 0; rb * 32 < sym.fronts[f].k; rb++) items.push_back(WorkItem{f, rb, 0});
  end(dsolveItems);

  dFronts.upload(fd);
  dRowIdx.upload(sym.rowIdx);
  if (!sym.rel.empty()) dRel.upload(sym.rel);
  dAsmSrc.upload(sym.asmSrc);
  dAsmDst.upload(sym.asmDst);
  dItems.upload(items);
  dPerm.upload(sym.perm);
  CUDA_CHECK(cudaStreamSynchronize(0));
  // the host copies of the big symbolic arrays are no longer needed
  std::vector<int64_t>().swap(sym.asmSrc);
  std::vector<int64_t>().swap(sym.asmDst);
}

size_t LdltPlan::plan_bytes() const {
  return dFronts.bytes() + dRowIdx.bytes() + dRel.bytes() + dAsmSrc.bytes() + dAsmDst.bytes() + dItems.bytes() +
         dPerm.bytes();
}

void LdltWorkspace::ensure(const Symbolic& s) {
  if ((int64_t)u0.n < s.uArena) { u0.alloc((size_t)s.uArena); u1.alloc((size_t)s.uArena); }
  if ((int64_t)w.n < s.wArena) w.alloc((size_t)s.wArena);
  if (counters.n < 2) counters.alloc(2);
}

FactorStats LdltFactor::factorize(const double* dVals, double pivTol, LdltWorkspace& ws, cudaStream_t st) {
  const LdltPlan& P = *plan_;
  const Symbolic& S = P.sym;
  FactorStats stats;
  const double t0 = now_s();
  ws.ensure(S);
  if ((int64_t)L.n != S.lSize) L.alloc((size_t)S.lSize);
  L.zero(st);
  ws.counters.zero(st);
  {
    const int64_t cnt = (int64_t)P.dAsmSrc.n;
    const int grid = (int)std::min<int64_t>((cnt + 255) / 256, 148 * 16);
    if (cnt) k_assemble<<<grid, 256, 0, st>>>(cnt, P.dAsmSrc.p, P.dAsmDst.p, dVals, L.p);
    CUDA_CHECK(cudaGetLastError());
  }
  const WorkItem* items = P.dItems.p;
  for (int l = 0; l < S.nlevels; l++) {
    double* Ucur = (l & 1) ? ws.u1.p : ws.u0.p;
    double* Uprev = (l & 1) ? ws.u0.p : ws.u1.p;
    if (P.levelU[l] > 0) CUDA_CHECK(cudaMemsetAsync(Ucur, 0, (size_t)P.levelU[l] * sizeof(double), st));
    if (P.eaddItems[l].cnt)
      k_extend_add<<<P.eaddItems[l].cnt, 256, 0, st>>>(items + P.eaddItems[l].off, P.dFronts.p, P.dRel.p, L.p, Uprev, Ucur);
    if (P.diagItems[l].cnt)
      k_diag_invert<<<P.diagItems[l].cnt, 256, 0, st>>>(items + P.diagItems[l].off, P.dFronts.p, L.p, pivTol, ws.counters.p);
    if (P.copyItems[l].cnt)
      k_copy_panel<<<P.copyItems[l].cnt, 256, 0, st>>>(items + P.copyItems[l].off, P.dFronts.p, L.p, ws.w.p);
    if (P.panelItems[l].cnt)
      k_panel<<<P.panelItems[l].cnt, GEMM_THREADS, 0, st>>>(items + P.panelItems[l].off, P.dFronts.p, L.p, ws.w.p);
    if (P.schurItems[l].cnt)
      k_schur<<<P.schurItems[l].cnt, GEMM_THREADS, 0, st>>>(items + P.schurItems[l].off, P.dFronts.p, L.p, ws.w.p, Ucur);
    CUDA_CHECK(cudaGetLastError());
  }
  int h[2] = {0, 0};
  CUDA_CHECK(cudaMemcpyAsync(h, ws.counters.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
  stats.neg = h[0];
  stats.perturbed = h[1];
  stats.seconds = now_s() - t0;
  return stats;
}

template <int NR>
static void solve_impl(const LdltPlan& P, const double* L, double* X, double* Y, int ldx, cudaStream_t st) {
  const Symbolic& S = P.sym;
  const WorkItem* items = P.dItems.p;
  for (int l = 0; l < S.nlevels; l++) {
    const int cnt = P.fwdItems[l].cnt;
    if (cnt) k_fwd<NR><<<(cnt + 7) / 8, 256, 0, st>>>(cnt, items + P.fwdItems[l].off, P.dFronts.p, P.dRowIdx.p, L, X, ldx);
  }
  {
    const int cnt = P.dsolveItems.cnt;
    k_dsolve<NR><<<(cnt + 7) / 8, 256, 0, st>>>(cnt, items + P.dsolveItems.off, P.dFronts.p, P.dRowIdx.p, L, X, Y, ldx);
  }
  for (int l = S.nlevels - 1; l >= 0; l--) {
    const int cnt = P.bwdItems[l].cnt;
    if (cnt) k_bwd<NR><<<(cnt + 7) / 8, 256, 0, st>>>(cnt, items + P.bwdItems[l].off, P.dFronts.p, P.dRowIdx.p, L, Y, ldx);
  }
  CUDA_CHECK(cudaGetLastError());
}

void LdltFactor::solve_permuted(double* X, double* Y, int ldx, int j0, int nr, cudaStream_t st) const {
  GENEO_CHECK(L.p != nullptr, "solve before factorize");
  switch (nr) {
    case 1: solve_impl<1>(*plan_, L.p, X + j0, Y + j0, ldx, st); break;
    case 2: solve_impl<2>(*plan_, L.p, X + j0, Y + j0, ldx, st); break;
    case 4: solve_impl<4>(*plan_, L.p, X + j0, Y + j0, ldx, st); break;
    case 8: solve_impl<8>(*plan_, L.p, X + j0, Y + j0, ldx, st); break;
    default: GENEO_CHECK(false, "nrhs chunk must be 1, 2, 4 or 8");
  }
}

}  // namespace geneo
