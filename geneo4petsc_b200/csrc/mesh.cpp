// mesh.cpp -- see mesh.hpp.  Host only.
#include "mesh.hpp"

#include <algorithm>
#include <cmath>
#include <atomic>
#include <fstream>
#include <functional>
#include <iostream>
#include <set>
#include <sstream>
#include <thread>

namespace geneo {

void Mesh::finalize() {
  const int ne = nbElem();
  matPtr.assign(ne + 1, 0);
  for (int e = 0; e < ne; e++) {
    int64_t k = elemPtr[e + 1] - elemPtr[e];
    matPtr[e + 1] = matPtr[e] + k * k;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Generators
// ---------------------------------------------------------------------------------------------------------------
int parse_gen_args(const std::string& args, GridGenOptions& o) {
  std::stringstream ss(args);
  while (ss) {
    std::string opt;
    ss >> opt;
    if (opt == "--size") { ss >> o.size; if (!ss) return 1; }
    if (opt == "--weakScaling") { ss >> o.weakScaling; if (!ss) return 1; }
    if (opt == "--dim") { ss >> o.dim; if (!ss || o.dim < 1 || o.dim > 3) return 1; }
    if (opt == "--inpEps") { ss >> o.inpEps; if (!ss) return 1; }
    if (opt == "--kappa") {
      ss >> o.kappaMax; if (!ss || o.kappaMax < 1.) return 1;
      ss >> o.kappaInterp; if (!ss) return 1;
      if (o.kappaInterp != "quad" && o.kappaInterp != "lin" && o.kappaInterp != "minmax") return 1;
    }
    if (opt == "--lbd") { ss >> o.lbd; if (!ss) return 1; }
    if (opt == "--dt") { ss >> o.dt; if (!ss) return 1; }
  }
  return 0;
}

static inline double kappa1d(int mode, double alpha, double beta, double x) {
  switch (mode) {
    case 1: return alpha * x * x + beta;  // quad
    case 2: return alpha * x + beta;      // lin
    case 3: {                             // minmax: a layer of alpha in the middle third
      double k = 1.;
      if (x >= beta) k = alpha;
      if (x >= 2. * beta) k = 1.;
      return k;
    }
    default: return 1.;
  }
}

int grid_edge(const GridGenOptions& o) {
  int n = 0;  // grid edge, with the reference's float truncation (laplacian.cpp:101-105)
  if (o.dim == 1) n = o.size * o.weakScaling;
  if (o.dim == 2) n = (int)std::sqrt((double)(o.size * o.size * o.weakScaling));
  if (o.dim == 3) n = (int)std::cbrt((double)(o.size * o.size * o.size * o.weakScaling));
  return n;
}

void generate_grid(const GridGenOptions& o, Mesh& m, std::vector<int>* elemPartOut) {
  const int n = grid_edge(o);
  const int n1 = n, n2 = (o.dim >= 2) ? n : 1, n3 = (o.dim >= 3) ? n : 1;
  int mode = 0;
  double alpha = 0., beta = 1.;
  const double xMax = (double)(n - 1);
  if (o.kappaInterp == "quad") { mode = 1; alpha = (o.kappaMax - beta) / (xMax * xMax); }
  else if (o.kappaInterp == "lin") { mode = 2; alpha = (o.kappaMax - beta) / xMax; }
  else if (o.kappaInterp == "minmax") { mode = 3; alpha = o.kappaMax; beta = xMax / 3.; }

  const int64_t npts = (int64_t)n1 * n2 * n3;
  GENEO_CHECK(npts < (int64_t)2147483647, "grid too large for 32-bit node ids");
  m.nbNode = (int)npts;
  m.grid[0] = n1; m.grid[1] = n2; m.grid[2] = n3;
  m.elemPtr.clear(); m.elemIdx.clear(); m.matVal.clear();
  if (elemPartOut) elemPartOut->clear();
  const bool sub = o.keepHi[0] >= 0;
  // sub-mesh: only the coordinates from one below the kept region can emit an element (a neighbour is +1 along one axis)
  int lo[3] = {0, 0, 0}, hi[3] = {n1, n2, n3};
  if (sub)
    for (int a = 0; a < 3; a++) {
      lo[a] = std::max(0, std::min(o.keepLo[a] - 1, hi[a]));
      hi[a] = std::max(lo[a], std::min(hi[a], o.keepHi[a]));
    }
  {
    const int64_t vol = (int64_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    m.elemPtr.reserve(vol * (o.dim + 0) + n1 * n2 + 1);
    m.elemIdx.reserve(vol * 2 * o.dim);
    m.matVal.reserve(vol * 4 * o.dim);
    if (elemPartOut) elemPartOut->reserve(vol * o.dim + n1 * n2);
  }
  m.elemPtr.push_back(0);
  const int nn[3] = {n1, n2, n3};
  auto inside = [&](int a, int b, int c) {
    return !sub || (a >= o.keepLo[0] && a < o.keepHi[0] && b >= o.keepLo[1] && b < o.keepHi[1] && c >= o.keepLo[2] && c < o.keepHi[2]);
  };
  auto boxOf = [&](int a, int b, int c) {
    const int K1 = std::max(1, o.boxK[0]), K2 = std::max(1, o.boxK[1]), K3 = std::max(1, o.boxK[2]);
    const int b1 = (int)(((int64_t)a * K1) / nn[0]), b2 = (int)(((int64_t)b * K2) / nn[1]), b3 = (int)(((int64_t)c * K3) / nn[2]);
    return b1 + K1 * (b2 + K2 * b3);
  };
  std::vector<double> k1(n1), k2(n2), k3(n3);
  for (int i = 0; i < n1; i++) k1[i] = kappa1d(mode, alpha, beta, (double)i);
  for (int i = 0; i < n2; i++) k2[i] = kappa1d(mode, alpha, beta, (double)i);
  for (int i = 0; i < n3; i++) k3[i] = kappa1d(mode, alpha, beta, (double)i);

  auto add = [&](int c, int nb, double kappa) {
    double dg = (1. + o.inpEps), og = -1.;
    if (nb >= 0) {
      // transform(elemMat *= kappa) then (heat) lbd*lap + inertia/dt -- same operation order as the reference
      double d = dg * kappa, f = og * kappa;
      if (o.heat) { d = o.lbd * d + (1. / 3.) / o.dt; f = o.lbd * f + (1. / 6.) / o.dt; }
      m.elemIdx.push_back(c); m.elemIdx.push_back(nb);
      m.matVal.push_back(d); m.matVal.push_back(f); m.matVal.push_back(f); m.matVal.push_back(d);
    } else {
      double d = dg * kappa;
      if (o.heat) d = o.lbd * d + (1. / 3.) / o.dt;
      m.elemIdx.push_back(c);
      m.matVal.push_back(d);
    }
    m.elemPtr.push_back((int64_t)m.elemIdx.size());
  };
  for (int d3 = lo[2]; d3 < hi[2]; d3++)
    for (int d2 = lo[1]; d2 < hi[1]; d2++)
      for (int d1 = lo[0]; d1 < hi[0]; d1++) {
        const int c = d1 + n1 * d2 + n1 * n2 * d3;
        const double kappa = k1[d1] * k2[d2] * k3[d3];
        const bool inC = inside(d1, d2, d3);
        const int box = elemPartOut ? boxOf(d1, d2, d3) : 0;
        auto emit = [&](int nb, bool keep) {
          if (!keep) return;
          add(c, nb, kappa);
          if (elemPartOut) elemPartOut->push_back(box);
        };
        if (o.dim == 1 && d1 == 0) emit(-1, inC);
        if (d1 + 1 < n1) emit(c + 1, inC || inside(d1 + 1, d2, d3));
        if (o.dim == 2 && d2 == 0) emit(-1, inC);
        if (d2 + 1 < n2) emit(c + n1, inC || inside(d1, d2 + 1, d3));
        if (o.dim == 3 && d3 == 0) emit(-1, inC);
        if (d3 + 1 < n3) emit(c + n1 * n2, inC || inside(d1, d2, d3 + 1));
      }
  m.finalize();
}

int parse_graph_args(const std::string& args, GraphGenOptions& o) {
  std::stringstream ss(args);
  while (ss) {
    std::string opt;
    ss >> opt;
    if (opt == "--size") { ss >> o.size; if (!ss) return 1; }
    if (opt == "--level") { ss >> o.level; if (!ss) return 1; }
    if (opt == "--weakScaling") { ss >> o.weakScaling; if (!ss) return 1; }
    if (opt == "--inpEps") { ss >> o.inpEps; if (!ss) return 1; }
    if (opt == "--noGround") o.noGround = true;
  }
  return (o.size >= 1 && o.level >= 0 && o.weakScaling >= 1) ? 0 : 1;
}

void generate_graph(const GraphGenOptions& o, Mesh& m) {
  const int B = (int)std::sqrt((double)(o.size * o.weakScaling));  // tst/graph/graph.cpp:153
  GENEO_CHECK(B >= 2, "graph generator: block size must be at least 2");
  const int64_t nblocks = 1 + 4 * (int64_t)o.level;
  const int64_t g0 = o.noGround ? 0 : 1;  // node 0 is the ground
  const int64_t nn = g0 + nblocks * B * B;
  GENEO_CHECK(nn < (int64_t)2147483647, "graph too large for 32-bit node ids");
  m = Mesh();
  m.nbNode = (int)nn;
  const int64_t perBlock = 2 * (int64_t)B * (B - 1) + (o.noGround ? 0 : 4 * B);
  const int64_t ne = nblocks * perBlock + (int64_t)o.level * 8 * B;
  m.elemPtr.reserve(ne + 1); m.elemIdx.reserve(2 * ne); m.matVal.reserve(4 * ne);
  m.elemPtr.push_back(0);
  auto edge = [&](int64_t a, int64_t b, double w) {
    m.elemIdx.push_back((int)a); m.elemIdx.push_back((int)b);
    m.matVal.push_back(w * (1. + o.inpEps)); m.matVal.push_back(w * -1.);
    m.matVal.push_back(w * -1.); m.matVal.push_back(w * (1. + o.inpEps));
    m.elemPtr.push_back((int64_t)m.elemIdx.size());
  };
  // block bi: nodes base + r B + c; its four borders in ascending node order: 0 up (first row), 1 right (last column),
  // 2 down (last row), 3 left (first column)
  auto base = [&](int64_t bi) { return g0 + bi * B * B; };
  auto border = [&](int64_t bi, int which, int i) -> int64_t {
    switch (which) {
      case 0: return base(bi) + i;
      case 1: return base(bi) + (int64_t)i * B + (B - 1);
      case 2: return base(bi) + (int64_t)(B - 1) * B + i;
      default: return base(bi) + (int64_t)i * B;
    }
  };
  auto block = [&](int64_t bi, double w) {
    const int64_t b0 = base(bi);
    for (int r = 0; r < B; r++)
      for (int c = 0; c + 1 < B; c++) edge(b0 + (int64_t)r * B + c, b0 + (int64_t)r * B + c + 1, w);
    for (int i = 0; i < B; i++)        // columns from the last one to the first, each walked upwards
      for (int j = 0; j + 1 < B; j++) edge(b0 + (int64_t)(B - 1 - j) * B + (B - 1 - i), b0 + (int64_t)(B - 2 - j) * B + (B - 1 - i), w);
    if (!o.noGround)
      for (int which = 0; which < 4; which++)
        for (int i = 0; i < B; i++) edge(border(bi, which, i), 0, w);
  };
  auto blk = [&](int l, int b) -> int64_t { return l == 0 ? 0 : 1 + 4 * (int64_t)(l - 1) + b; };  // level 0 = the central block, four times
  block(0, 1.);
  for (int l = 1; l <= o.level; l++) {
    for (int b = 0; b < 4; b++) block(blk(l, b), l + 1.);
    const double w = 0.5 * (l + 1.);
    for (int b = 0; b < 4; b++) {  // around the level: right->up, down->right, left->down, up->left of the next block
      const int from = (b + 1) % 4, to = b;
      for (int i = 0; i < B; i++) edge(border(blk(l, b), from, i), border(blk(l, (b + 1) % 4), to, i), w);
    }
    for (int b = 0; b < 4; b++)    // to the previous level: border b of the inner block -> the opposite border of the outer one
      for (int i = 0; i < B; i++) edge(border(blk(l - 1, b), b, i), border(blk(l, b), (b + 2) % 4, i), w);
  }
  m.finalize();
}

int read_input_file(const std::string& path, double inpEps, Mesh& m) {
  std::ifstream inp(path);
  if (!inp) { std::cerr << "Error: can not open " << path << std::endl; return 1; }
  m = Mesh();
  m.elemPtr.push_back(0);
  std::set<int> nodes;
  std::string line;
  while (std::getline(inp, line)) {
    size_t s = 0;
    while (s < line.size() && isspace((unsigned char)line[s])) s++;
    line = line.substr(s);
    if (line.empty() || line[0] == '%' || line[0] == '#') continue;
    std::stringstream ss(line);
    std::string tok;
    bool fillDof = true;
    std::vector<int> dofs;
    std::vector<double> vals;
    while (ss >> tok) {
      if (tok == "-") { fillDof = false; continue; }
      std::stringstream ts(tok);
      if (fillDof) { int d = 0; ts >> d; if (ts) dofs.push_back(d); }
      else { double a = 0.; ts >> a; if (ts) vals.push_back(a); }
    }
    const int nd = (int)dofs.size();
    if (vals.empty())
      for (int i = 0; i < nd; i++)
        for (int j = 0; j < nd; j++) vals.push_back(i == j ? 1. + inpEps : -1. / ((double)(nd - 1)));
    if ((int)vals.size() != nd * nd) { std::cerr << "Error: bad matrix (" << m.nbElem() + 1 << ") in file " << std::endl; return 1; }
    for (int d : dofs) { m.elemIdx.push_back(d); nodes.insert(d); }
    m.elemPtr.push_back((int64_t)m.elemIdx.size());
    m.matVal.insert(m.matVal.end(), vals.begin(), vals.end());
  }
  if (nodes.empty()) { std::cerr << "Error: empty input" << std::endl; return 1; }
  m.nbNode = (int)nodes.size();
  if (*nodes.rbegin() + 1 != m.nbNode || *nodes.begin() < 0) {
    std::cerr << "Error: bad node set (" << *nodes.rbegin() + 1 << "/" << m.nbNode << ") in file " << std::endl;
    return 1;
  }
  m.finalize();
  return 0;
}

int read_rhs_file(const std::string& path, int n, std::vector<double>& b) {
  std::ifstream inp(path);
  if (!inp) { std::cerr << "Error: can not open " << path << std::endl; return 1; }
  b.assign(n, 0.);
  std::string line;
  while (std::getline(inp, line)) {
    size_t s = 0;
    while (s < line.size() && isspace((unsigned char)line[s])) s++;
    line = line.substr(s);
    if (line.empty() || line[0] == '%' || line[0] == '#') continue;
    std::stringstream ss(line);
    long idx; double v;
    ss >> idx;
    if (!ss) { std::cerr << "Error: can not read " << path << std::endl; return 1; }
    ss >> v;
    if (!ss) v = 1.;
    if (idx < 0 || idx >= n) { std::cerr << "Error: bad index in " << path << std::endl; return 1; }
    b[idx] = v;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// METIS (CUDA-toolkit libmetis_static.a: 64-bit idx_t, 32-bit real_t, no header shipped)
// ---------------------------------------------------------------------------------------------------------------
typedef int64_t midx_t;
extern "C" {
int METIS_SetDefaultOptions(midx_t* options);
int METIS_PartMeshDual(midx_t* ne, midx_t* nn, midx_t* eptr, midx_t* eind, midx_t* vwgt, midx_t* vsize, midx_t* ncommon,
                       midx_t* nparts, float* tpwgts, midx_t* options, midx_t* objval, midx_t* epart, midx_t* npart);
int METIS_PartMeshNodal(midx_t* ne, midx_t* nn, midx_t* eptr, midx_t* eind, midx_t* vwgt, midx_t* vsize, midx_t* nparts,
                        float* tpwgts, midx_t* options, midx_t* objval, midx_t* epart, midx_t* npart);
}

int metis_partition(const Mesh& m, int nbPart, bool dual, std::vector<int>& elemPart, std::vector<int>& nodePart) {
  const int ne = m.nbElem(), nn = m.nbNode;
  elemPart.assign(ne, 0);
  nodePart.assign(nn, 0);
  if (nbPart == 1) return 0;  // the reference does not call METIS for one partition
  midx_t options[40];
  METIS_SetDefaultOptions(options);
  options[10] = 1;  // METIS_OPTION_MINCONN
  options[0] = 1;   // METIS_OPTION_PTYPE  = METIS_PTYPE_KWAY
  options[1] = 0;   // METIS_OPTION_OBJTYPE = METIS_OBJTYPE_CUT
  midx_t obj = 0, ncommon = 1, NE = ne, NN = nn, NP = nbPart;
  std::vector<midx_t> eptr(m.elemPtr.begin(), m.elemPtr.end()), eind(m.elemIdx.begin(), m.elemIdx.end());
  std::vector<midx_t> ep(ne), np(nn);
  int rc;
  if (dual) rc = METIS_PartMeshDual(&NE, &NN, eptr.data(), eind.data(), NULL, NULL, &ncommon, &NP, NULL, options, &obj, ep.data(), np.data());
  else      rc = METIS_PartMeshNodal(&NE, &NN, eptr.data(), eind.data(), NULL, NULL, &NP, NULL, options, &obj, ep.data(), np.data());
  if (rc != 1) { std::cerr << "Error: METIS partition KO" << std::endl; return 1; }
  for (int e = 0; e < ne; e++) elemPart[e] = (int)ep[e];
  for (int i = 0; i < nn; i++) nodePart[i] = (int)np[i];
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Decomposition
// ---------------------------------------------------------------------------------------------------------------
namespace {

// Sum-duplicate COO -> CSR with sorted columns.  rows/cols are local indices < n.
void coo_to_csr(int n, std::vector<int>& rows, std::vector<int>& cols, std::vector<double>& vals, CsrHost& a) {
  const size_t nz = rows.size();
  std::vector<int64_t> cnt(n + 1, 0);
  for (size_t t = 0; t < nz; t++) cnt[rows[t] + 1]++;
  for (int i = 0; i < n; i++) cnt[i + 1] += cnt[i];
  std::vector<int> c2(nz);
  std::vector<double> v2(nz);
  {
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (size_t t = 0; t < nz; t++) {
      int64_t q = pos[rows[t]]++;
      c2[q] = cols[t];
      v2[q] = vals[t];
    }
  }
  a.n = a.ncols = n;
  a.ptr.assign(n + 1, 0);
  a.idx.clear(); a.val.clear();
  a.idx.reserve(nz / 2 + n); a.val.reserve(nz / 2 + n);
  std::vector<std::pair<int, double>> tmp;
  for (int i = 0; i < n; i++) {
    tmp.clear();
    for (int64_t q = cnt[i]; q < cnt[i + 1]; q++) tmp.emplace_back(c2[q], v2[q]);
    std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
    for (size_t k = 0; k < tmp.size();) {
      int c = tmp[k].first;
      double s = 0.;
      while (k < tmp.size() && tmp[k].first == c) { s += tmp[k].second; k++; }  // insertion order, like ADD_VALUES
      a.idx.push_back(c);
      a.val.push_back(s);
    }
    a.ptr[i + 1] = (int64_t)a.idx.size();
  }
}

// One CSR row at a time: contributions are summed per column in ARRIVAL order (MatSetValues ADD_VALUES semantics, the
// same sums to the bit as a stable sort of the triplets), the columns of a finished row are sorted and appended.
struct RowAccumulator {
  std::vector<int> stamp, slot, cols;
  std::vector<double> vals;
  std::vector<std::pair<int, int>> order;
  int cur = -1;
  void reset(int ncols) { stamp.assign(ncols, -1); slot.resize(ncols); cur = -1; }
  void begin(int row) { cur = row; cols.clear(); vals.clear(); }
  void add(int c, double v) {
    if (stamp[c] != cur) { stamp[c] = cur; slot[c] = (int)cols.size(); cols.push_back(c); vals.push_back(0. + v); }
    else vals[slot[c]] += v;
  }
  void flush(CsrHost& a) {
    const int m = (int)cols.size();
    if (m <= 48) {  // the usual case (stencil rows): insertion sort of the column / value pairs in place
      for (int q = 1; q < m; q++) {
        const int c = cols[q];
        const double v = vals[q];
        int r = q - 1;
        for (; r >= 0 && cols[r] > c; r--) { cols[r + 1] = cols[r]; vals[r + 1] = vals[r]; }
        cols[r + 1] = c; vals[r + 1] = v;
      }
      a.idx.insert(a.idx.end(), cols.begin(), cols.end());
      a.val.insert(a.val.end(), vals.begin(), vals.end());
    } else {
      order.clear();
      for (int q = 0; q < m; q++) order.emplace_back(cols[q], q);
      std::sort(order.begin(), order.end());
      for (auto& o : order) { a.idx.push_back(o.first); a.val.push_back(vals[o.second]); }
    }
    a.ptr.push_back((int64_t)a.idx.size());
  }
};

// host threads of THIS process: the cores divided by the ranks sharing the node (torchrun exports LOCAL_WORLD_SIZE), at most
// 16 -- every thread of the passes below owns per-node work arrays
unsigned host_threads() {
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  if (const char* e = getenv("GENEO_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));
  else if (const char* e2 = getenv("LOCAL_WORLD_SIZE")) hw = std::max(1u, hw / (unsigned)std::max(1, atoi(e2)));
  return std::min(hw, 16u);
}

void run_threads(unsigned nthreads, const std::function<void(unsigned)>& fn) {
  std::vector<std::thread> pool;
  std::vector<std::string> errs(nthreads);
  for (unsigned tid = 0; tid < nthreads; tid++)
    pool.emplace_back([&, tid]() { try { fn(tid); } catch (std::exception& ex) { errs[tid] = ex.what(); } });
  for (auto& t : pool) t.join();
  for (auto& e : errs) GENEO_CHECK(e.empty(), e);
}

}  // namespace

void NodeIndex::build(int nbNode, const std::vector<int>& elemIdx) {
  *this = NodeIndex();
  nn = nbNode;
  const size_t words = ((size_t)nbNode + 63) / 64;
  bits.assign(words, 0);
  for (int g : elemIdx) {
    GENEO_CHECK(g >= 0 && g < nbNode, "mesh: node id out of range");
    bits[(size_t)g >> 6] |= 1ull << (g & 63);
  }
  pre.resize(words + 1);
  int c = 0;
  for (size_t w = 0; w < words; w++) { pre[w] = c; c += __builtin_popcountll(bits[w]); }
  pre[words] = c;
  nc = c;
  if (nc == nn) {  // every node is held: identity
    std::vector<uint64_t>().swap(bits);
    std::vector<int>().swap(pre);
    return;
  }
  active = true;
  present.resize(nc);
  int q = 0;
  for (size_t w = 0; w < words; w++)
    for (uint64_t x = bits[w]; x; x &= x - 1) present[q++] = (int)(w * 64 + (size_t)__builtin_ctzll(x));
}

void decompose(const Mesh& m, int nbPart, const std::vector<int>& elemPart, const std::vector<int>& nodePart,
               bool dual, int overlap, const std::vector<char>& owner, Decomposition& d) {
  const int ne = m.nbElem(), nn = m.nbNode;
  const bool tm = getenv("GENEO_DECOMP_TIMING") != nullptr;
  double tq = now_s();
  auto lap = [&](const char* what) { if (tm) { const double t = now_s(); fprintf(stderr, "decompose nn=%d ne=%d %s %.3fs\n", nn, ne, what, t - tq); tq = t; } };
  d = Decomposition();
  d.nbPart = nbPart; d.nbNode = nn; d.nbElem = ne;
  for (int a = 0; a < 3; a++) d.grid[a] = m.grid[a];
  // Every per-node array below is indexed by the DENSE id of the node among the nodes this mesh holds (ascending global
  // id): a rank of an N-GPU run holds a sub-mesh with 1/N of the nodes under global ids.
  d.index.build(nn, m.elemIdx);
  const NodeIndex& ix = d.index;
  const int nc = ix.size();
  std::vector<int> denseStore;
  const int* eidx = m.elemIdx.data();
  if (ix.active) {
    denseStore.resize(m.elemIdx.size());
    for (size_t t = 0; t < m.elemIdx.size(); t++) denseStore[t] = ix(m.elemIdx[t]);
    eidx = denseStore.data();
  }
  d.nodeMult.assign(nc, 0);
  d.elemMult.assign(ne, 0);
  d.subs.resize(nbPart);
  const unsigned hw = host_threads();

  // inverse topology node -> elements, ascending element ids (counting sort; start(c) = n2ePtr[c], end(c) = n2ePtr[c + 1])
  std::vector<int64_t> n2ePtr((size_t)nc + 2, 0);
  for (size_t t = 0; t < m.elemIdx.size(); t++) n2ePtr[(size_t)eidx[t] + 2]++;
  for (int c = 0; c < nc; c++) n2ePtr[(size_t)c + 2] += n2ePtr[(size_t)c + 1];
  std::vector<int> n2e(m.elemIdx.size());
  for (int e = 0; e < ne; e++)
    for (int64_t t = m.elemPtr[e]; t < m.elemPtr[e + 1]; t++) n2e[n2ePtr[(size_t)eidx[t] + 1]++] = e;
  n2ePtr.pop_back();
  lap("node -> elements");

  // -- element / node sets -----------------------------------------------------------------------------------------
  std::vector<std::vector<int>> partElems(nbPart);
  if (dual) {
    std::vector<int> cnt(nbPart, 0);
    for (int e = 0; e < ne; e++) {
      GENEO_CHECK(elemPart[e] >= 0 && elemPart[e] < nbPart, "bad element partition");
      cnt[elemPart[e]]++;
    }
    for (int p = 0; p < nbPart; p++) partElems[p].reserve(cnt[p]);
    for (int e = 0; e < ne; e++) partElems[elemPart[e]].push_back(e);
  } else {
    std::vector<std::vector<int>> partNodes(nbPart);
    for (int c = 0; c < nc; c++) {
      const int q = nodePart[ix.global(c)];
      GENEO_CHECK(q >= 0 && q < nbPart, "bad node partition");
      partNodes[q].push_back(c);
    }
    std::vector<int> stamp(ne, -1);
    for (int p = 0; p < nbPart; p++)  // an element belongs to p if one of its nodes does
      for (int c : partNodes[p])
        for (int64_t t = n2ePtr[c]; t < n2ePtr[c + 1]; t++) {
          int e = n2e[t];
          if (stamp[e] != p) { stamp[e] = p; partElems[p].push_back(e); }
        }
  }
  // dense node lists of the subdomains (subs[p].nodes holds the global ids)
  std::vector<std::vector<int>> denseNodes(ix.active ? nbPart : 0);
  auto nodesOf = [&](int p) -> const std::vector<int>& { return ix.active ? denseNodes[p] : d.subs[p].nodes; };
  {
    const unsigned nthreads = std::max(1u, std::min((unsigned)nbPart, hw));
    std::atomic<int> ticket(0);
    run_threads(nthreads, [&](unsigned) {
      std::vector<int> estamp, nstamp(nc, -1);
      if (overlap > 0) estamp.assign(ne, -1);
      for (;;) {
        const int p = ticket.fetch_add(1);
        if (p >= nbPart) break;
        std::vector<int>& el = partElems[p];
        if (overlap > 0) {
          for (int e : el) estamp[e] = p;
          size_t layerBegin = 0;
          for (int l = 0; l < overlap; l++) {  // one layer = every element sharing a node with the current set
            // (the reference rescans the whole set each time; scanning only the last layer + first pass is equivalent
            //  because older elements' neighbours were already added)
            size_t cur = el.size();
            size_t from = (l == 0) ? 0 : layerBegin;
            for (size_t q = from; q < cur; q++) {
              int e = el[q];
              for (int64_t t = m.elemPtr[e]; t < m.elemPtr[e + 1]; t++) {
                int c = eidx[t];
                for (int64_t u = n2ePtr[c]; u < n2ePtr[c + 1]; u++) {
                  int e2 = n2e[u];
                  if (estamp[e2] != p) { estamp[e2] = p; el.push_back(e2); }
                }
              }
            }
            layerBegin = cur;
          }
        }
        std::sort(el.begin(), el.end());
        Subdomain& s = d.subs[p];
        s.id = p;
        s.elems.swap(el);
        std::vector<int> dn;
        for (int e : s.elems)
          for (int64_t t = m.elemPtr[e]; t < m.elemPtr[e + 1]; t++) {
            int c = eidx[t];
            if (nstamp[c] != p) { nstamp[c] = p; dn.push_back(c); }
          }
        std::sort(dn.begin(), dn.end());
        if (ix.active) {
          s.nodes.resize(dn.size());
          for (size_t l = 0; l < dn.size(); l++) s.nodes[l] = ix.present[dn[l]];
          denseNodes[p].swap(dn);
        } else s.nodes.swap(dn);
      }
    });
    for (int p = 0; p < nbPart; p++) {
      for (int e : d.subs[p].elems) d.elemMult[e]++;
      for (int c : nodesOf(p)) d.nodeMult[c]++;
    }
  }
  lap("element / node sets");

  // -- multiplicities + intersections (local indices, ascending) -----------------------------------------------------
  d.nodeSubPtr.assign((size_t)nc + 1, 0);
  for (int c = 0; c < nc; c++) d.nodeSubPtr[c + 1] = d.nodeSubPtr[c] + d.nodeMult[c];
  d.nodeSub.resize((size_t)d.nodeSubPtr[nc]);
  {
    std::vector<int64_t> pos(d.nodeSubPtr.begin(), d.nodeSubPtr.end() - 1);
    for (int p = 0; p < nbPart; p++)
      for (int c : nodesOf(p)) d.nodeSub[pos[c]++] = p;
  }
  const std::vector<int64_t>& n2dPtr = d.nodeSubPtr;
  const std::vector<int>& n2d = d.nodeSub;
  {
    std::atomic<int> ticket(0);
    run_threads(std::max(1u, std::min((unsigned)nbPart, hw)), [&](unsigned) {
      for (;;) {
        const int p = ticket.fetch_add(1);
        if (p >= nbPart) break;
        Subdomain& s = d.subs[p];
        const std::vector<int>& dn = nodesOf(p);
        const int nl = (int)dn.size();
        s.mult.resize(nl);
        s.intersect.assign(nbPart, std::vector<int>());
        for (int l = 0; l < nl; l++) {
          const int c = dn[l];
          s.mult[l] = d.nodeMult[c];
          if (d.nodeMult[c] > 1)
            for (int64_t t = n2dPtr[c]; t < n2dPtr[c + 1]; t++)
              if (n2d[t] != p) s.intersect[n2d[t]].push_back(l);
        }
      }
    });
  }
  lap("multiplicities + intersections");

  // -- local matrices for owned subdomains: assembled row by row (no triplet lists) ------------------------------------
  std::vector<int> mine;
  for (int p = 0; p < nbPart; p++)
    if (owner.empty() || owner[p]) mine.push_back(p);
  {
    const unsigned nthreads = std::max(1u, std::min((unsigned)mine.size(), hw));
    std::atomic<int> ticket(0);
    run_threads(nthreads, [&](unsigned) {
      std::vector<int> g2l(nc, -1), emark(ne, -1);
      RowAccumulator acc;
      std::vector<std::pair<int, int>> touching;  // (first local node of the element, element)
      for (;;) {
        const int w = ticket.fetch_add(1);
        if (w >= (int)mine.size()) break;
        const int p = mine[w];
        Subdomain& s = d.subs[p];
        const std::vector<int>& dn = nodesOf(p);
        const int nl = (int)dn.size();
        for (int l = 0; l < nl; l++) g2l[dn[l]] = l;
        for (int e : s.elems) emark[e] = p;
        acc.reset(nl);
        // Neumann: own elements weighted by 1/elemMult (buildDomain :473-476, fillALoc :688-708); the contributions of a
        // row arrive in ascending element order
        s.aNeu = CsrHost();
        s.aNeu.n = s.aNeu.ncols = nl;
        s.aNeu.ptr.reserve((size_t)nl + 1);
        s.aNeu.ptr.push_back(0);
        {
          size_t cap = 0;  // upper bound: every (row, column) pair of every element once
          for (int e : s.elems) { const size_t k = (size_t)(m.elemPtr[e + 1] - m.elemPtr[e]); cap += k * k; }
          cap = std::min(cap, (size_t)nl * 32);
          s.aNeu.idx.reserve(cap);
          s.aNeu.val.reserve(cap);
        }
        for (int l = 0; l < nl; l++) {
          const int c = dn[l];
          acc.begin(l);
          for (int64_t u = n2ePtr[c]; u < n2ePtr[c + 1]; u++) {
            const int e = n2e[u];
            if (emark[e] != p) continue;
            const int64_t b = m.elemPtr[e];
            const int k = (int)(m.elemPtr[e + 1] - b);
            const double w8 = 1. / ((double)d.elemMult[e]);
            const double* K = &m.matVal[m.matPtr[e]];
            for (int i = 0; i < k; i++) {
              if (eidx[b + i] != c) continue;
              for (int j = 0; j < k; j++) acc.add(g2l[eidx[b + j]], K[i * k + j] * w8);
            }
          }
          acc.flush(s.aNeu);
        }
        // Dirichlet R A R^T: every element touching the subdomain, restricted to its nodes, full weight.  Elements
        // contribute in the order of their first local node, then of their id (the order a sweep over the nodes meets them)
        s.aDir = CsrHost();
        s.aDir.n = s.aDir.ncols = nl;
        s.aDir.ptr.reserve((size_t)nl + 1);
        s.aDir.ptr.push_back(0);
        s.aDir.idx.reserve(s.aNeu.idx.size());
        s.aDir.val.reserve(s.aNeu.idx.size());
        for (int l = 0; l < nl; l++) {
          const int c = dn[l];
          touching.clear();
          bool sorted = true;
          for (int64_t u = n2ePtr[c]; u < n2ePtr[c + 1]; u++) {
            const int e = n2e[u];
            int first = l;
            for (int64_t t = m.elemPtr[e]; t < m.elemPtr[e + 1]; t++) {
              const int lj = g2l[eidx[t]];
              if (lj >= 0 && lj < first) first = lj;
            }
            sorted = sorted && (touching.empty() || touching.back() < std::make_pair(first, e));
            touching.emplace_back(first, e);
          }
          if (!sorted) std::sort(touching.begin(), touching.end());
          acc.begin(l);
          for (auto& fe : touching) {
            const int e = fe.second;
            const int64_t b = m.elemPtr[e];
            const int k = (int)(m.elemPtr[e + 1] - b);
            const double* K = &m.matVal[m.matPtr[e]];
            for (int i = 0; i < k; i++) {
              if (eidx[b + i] != c) continue;
              for (int j = 0; j < k; j++) {
                const int lj = g2l[eidx[b + j]];
                if (lj >= 0) acc.add(lj, K[i * k + j]);
              }
            }
          }
          acc.flush(s.aDir);
        }
        for (int l = 0; l < nl; l++) g2l[dn[l]] = -1;
      }
    });
  }
  d.nnzNeuTotal = 0;
  for (int p : mine) d.nnzNeuTotal += d.subs[p].aNeu.nnz();
  lap("local matrices");
}

void finish_predecomposed(Decomposition& d) {
  const int nn = d.nbNode, P = d.nbPart;
  GENEO_CHECK((int)d.subs.size() == P && P >= 1 && nn >= 1, "pre-decomposed problem: bad sizes");
  d.index = NodeIndex();
  d.index.nn = nn;  // every DOF belongs to a subdomain (checked below): identity
  d.nodeMult.assign(nn, 0);
  for (int p = 0; p < P; p++) {
    Subdomain& s = d.subs[p];
    s.id = p;
    GENEO_CHECK(!s.nodes.empty() && s.aNeu.n == (int)s.nodes.size(), "pre-decomposed problem: subdomain without nodes / matrix");
    for (size_t l = 0; l < s.nodes.size(); l++) {
      GENEO_CHECK(s.nodes[l] >= 0 && s.nodes[l] < nn, "pre-decomposed problem: global id out of range");
      GENEO_CHECK(l == 0 || s.nodes[l - 1] < s.nodes[l], "pre-decomposed problem: global ids must be sorted ascending");
      d.nodeMult[s.nodes[l]]++;
    }
  }
  for (int g = 0; g < nn; g++) GENEO_CHECK(d.nodeMult[g] > 0, "pre-decomposed problem: a DOF belongs to no subdomain");
  d.nodeSubPtr.assign(nn + 1, 0);
  for (int g = 0; g < nn; g++) d.nodeSubPtr[g + 1] = d.nodeSubPtr[g] + d.nodeMult[g];
  d.nodeSub.assign(d.nodeSubPtr[nn], 0);
  {
    std::vector<int64_t> pos(d.nodeSubPtr.begin(), d.nodeSubPtr.end() - 1);
    for (int p = 0; p < P; p++)
      for (int g : d.subs[p].nodes) d.nodeSub[pos[g]++] = p;
  }
  d.nnzNeuTotal = 0;
  bool needDir = false;
  for (int p = 0; p < P; p++) {
    Subdomain& s = d.subs[p];
    const int nl = (int)s.nodes.size();
    s.mult.resize(nl);
    s.intersect.assign(P, std::vector<int>());
    for (int l = 0; l < nl; l++) {
      const int g = s.nodes[l];
      s.mult[l] = d.nodeMult[g];
      if (d.nodeMult[g] > 1)
        for (int64_t t = d.nodeSubPtr[g]; t < d.nodeSubPtr[g + 1]; t++)
          if (d.nodeSub[t] != p) s.intersect[d.nodeSub[t]].push_back(l);
    }
    d.nnzNeuTotal += s.aNeu.nnz();
    if (s.aDir.n != nl) needDir = true;
  }
  if (!needDir) return;
  // global operator rows, then the sub-blocks
  std::vector<int64_t> cnt(nn + 1, 0);
  for (auto& s : d.subs)
    for (int l = 0; l < s.aNeu.n; l++) cnt[s.nodes[l] + 1] += s.aNeu.ptr[l + 1] - s.aNeu.ptr[l];
  for (int i = 0; i < nn; i++) cnt[i + 1] += cnt[i];
  std::vector<int> ci((size_t)cnt[nn]);
  std::vector<double> cv((size_t)cnt[nn]);
  {
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (auto& s : d.subs)
      for (int l = 0; l < s.aNeu.n; l++) {
        int64_t& q = pos[s.nodes[l]];
        for (int64_t t = s.aNeu.ptr[l]; t < s.aNeu.ptr[l + 1]; t++) { ci[q] = s.nodes[s.aNeu.idx[t]]; cv[q] = s.aNeu.val[t]; q++; }
      }
  }
  std::vector<int> g2l(nn, -1);
  for (int p = 0; p < P; p++) {
    Subdomain& s = d.subs[p];
    const int nl = (int)s.nodes.size();
    if (s.aDir.n == nl) continue;
    for (int l = 0; l < nl; l++) g2l[s.nodes[l]] = l;
    std::vector<int> rows, cols;
    std::vector<double> vals;
    for (int l = 0; l < nl; l++) {
      const int g = s.nodes[l];
      for (int64_t t = cnt[g]; t < cnt[g + 1]; t++)
        if (g2l[ci[t]] >= 0) { rows.push_back(l); cols.push_back(g2l[ci[t]]); vals.push_back(cv[t]); }
    }
    coo_to_csr(nl, rows, cols, vals, s.aDir);
    for (int l = 0; l < nl; l++) g2l[s.nodes[l]] = -1;
  }
}

void build_rank_layout(const Mesh& m, const Decomposition& d, const std::vector<int>& subRank, int rank, int world,
                       RankLayout& L) {
  const int nn = m.nbNode, ne = m.nbElem();
  GENEO_CHECK((int)subRank.size() == d.nbPart, "layout: one rank per subdomain expected");
  GENEO_CHECK(rank >= 0 && rank < world, "layout: bad rank");
  const bool tm = getenv("GENEO_DECOMP_TIMING") != nullptr;
  double tq = now_s();
  auto lap = [&](const char* what) { if (tm) { const double t = now_s(); fprintf(stderr, "layout nn=%d ne=%d %s %.3fs\n", nn, ne, what, t - tq); tq = t; } };
  L = RankLayout();
  L.rank = rank; L.world = world; L.nbNode = nn; L.subRank = subRank;
  // dense node ids of the (sub-)mesh, as in decompose (d.nodeSubPtr is indexed by them)
  L.index = d.index;
  const NodeIndex& ix = L.index;
  const int nc = ix.size();
  GENEO_CHECK((int)d.nodeSubPtr.size() == nc + 1, "layout: the decomposition was built on another mesh");
  std::vector<int> denseStore;
  const int* eidx = m.elemIdx.data();
  if (ix.active) {
    denseStore.resize(m.elemIdx.size());
    for (size_t t = 0; t < m.elemIdx.size(); t++) {
      denseStore[t] = ix(m.elemIdx[t]);
      GENEO_CHECK(denseStore[t] >= 0, "layout: the decomposition was built on another mesh");
    }
    eidx = denseStore.data();
  }
  auto ownerRank = [&](int c) -> int {  // rank of the lowest-numbered subdomain containing the node (-1: in no subdomain)
    return d.nodeSubPtr[c] < d.nodeSubPtr[c + 1] ? subRank[d.nodeSub[d.nodeSubPtr[c]]] : -1;
  };
  // 1 = owned, 2 = ghost
  std::vector<char> flag(nc, 0);
  for (int p = 0; p < d.nbPart; p++) {
    if (subRank[p] != rank) continue;
    for (int g : d.subs[p].nodes) { const int c = ix(g); flag[c] = (ownerRank(c) == rank) ? 1 : 2; }
  }
  // node -> elements, for the owned rows only (ascending element ids)
  std::vector<int64_t> n2ePtr((size_t)nc + 2, 0);
  for (size_t t = 0; t < m.elemIdx.size(); t++)
    if (flag[eidx[t]] == 1) n2ePtr[(size_t)eidx[t] + 2]++;
  for (int c = 0; c < nc; c++) n2ePtr[(size_t)c + 2] += n2ePtr[(size_t)c + 1];
  std::vector<int> n2e((size_t)n2ePtr[(size_t)nc + 1]);
  for (int e = 0; e < ne; e++)
    for (int64_t t = m.elemPtr[e]; t < m.elemPtr[e + 1]; t++)
      if (flag[eidx[t]] == 1) n2e[n2ePtr[(size_t)eidx[t] + 1]++] = e;
  n2ePtr.pop_back();
  // columns of owned rows that are in none of the rank's subdomains become ghosts too
  for (int c = 0; c < nc; c++) {
    if (flag[c] != 1) continue;
    for (int64_t u = n2ePtr[c]; u < n2ePtr[c + 1]; u++) {
      const int e = n2e[u];
      for (int64_t t = m.elemPtr[e]; t < m.elemPtr[e + 1]; t++)
        if (flag[eidx[t]] == 0) flag[eidx[t]] = 2;
    }
  }
  std::vector<std::vector<int>> byOwner(world);
  std::vector<int> ownedDense;
  for (int c = 0; c < nc; c++) {
    if (flag[c] == 1) { L.owned.push_back(ix.global(c)); ownedDense.push_back(c); }
    else if (flag[c] == 2) {
      const int q = ownerRank(c);
      GENEO_CHECK(q >= 0 && q < world && q != rank, "layout: ghost node without a remote owner");
      byOwner[q].push_back(c);
    }
  }
  L.c2l.assign(nc, -1);
  for (int i = 0; i < L.nOwn(); i++) L.c2l[ownedDense[i]] = i;
  L.ghostPtr.assign(world + 1, 0);
  for (int q = 0; q < world; q++) {
    for (int c : byOwner[q]) { L.c2l[c] = L.nOwn() + (int)L.ghost.size(); L.ghost.push_back(ix.global(c)); }
    L.ghostPtr[q + 1] = (int64_t)L.ghost.size();
  }
  L.sendIdx.assign(world, std::vector<int>());
  lap("owned / ghost sets");
  // owned rows of A = sum over ALL elements touching the node, full weight (== sum_i R_i^T A_neu,i R_i, SURVEY.md 8a note 1),
  // assembled row by row in ascending element order; row blocks in parallel, concatenated afterwards
  const int nOwn = L.nOwn(), ncols = nOwn + L.nGhost();
  const unsigned nthreads = std::max(1u, std::min(host_threads(), (unsigned)std::max(1, nOwn / 4096)));
  std::vector<CsrHost> piece(nthreads);
  run_threads(nthreads, [&](unsigned tid) {
    const int i0 = (int)((int64_t)nOwn * tid / nthreads), i1 = (int)((int64_t)nOwn * (tid + 1) / nthreads);
    RowAccumulator acc;
    acc.reset(ncols);
    CsrHost& a = piece[tid];
    a.ptr.reserve((size_t)(i1 - i0) + 1);
    a.ptr.push_back(0);
    for (int i = i0; i < i1; i++) {
      const int c = ownedDense[i];
      acc.begin(i);
      for (int64_t u = n2ePtr[c]; u < n2ePtr[c + 1]; u++) {
        const int e = n2e[u];
        const int64_t b = m.elemPtr[e];
        const int k = (int)(m.elemPtr[e + 1] - b);
        const double* K = &m.matVal[m.matPtr[e]];
        for (int a2 = 0; a2 < k; a2++) {
          if (eidx[b + a2] != c) continue;
          for (int j = 0; j < k; j++) acc.add(L.c2l[eidx[b + j]], K[a2 * k + j]);
        }
      }
      acc.flush(a);
    }
  });
  L.A = CsrHost();
  L.A.n = nOwn;
  L.A.ncols = ncols;
  L.A.ptr.assign((size_t)nOwn + 1, 0);
  int64_t total = 0;
  for (auto& a : piece) total += a.nnz();
  L.A.idx.reserve((size_t)total);
  L.A.val.reserve((size_t)total);
  {
    int row = 0;
    for (auto& a : piece) {
      const int64_t base = (int64_t)L.A.idx.size();
      for (size_t r = 1; r < a.ptr.size(); r++) L.A.ptr[++row] = base + a.ptr[r];
      L.A.idx.insert(L.A.idx.end(), a.idx.begin(), a.idx.end());
      L.A.val.insert(L.A.val.end(), a.val.begin(), a.val.end());
      a = CsrHost();
    }
    GENEO_CHECK(row == nOwn, "layout: operator rows lost");
  }
  lap("owned rows of A");
}

}  // namespace geneo
