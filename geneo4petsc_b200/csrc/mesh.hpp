// mesh.hpp -- host side of the driver path: input mesh, METIS partition, overlapping decomposition,
// weighted Neumann / Dirichlet local matrices.  Mirrors (re-designed, not translated) the driver half of the
// reference: src/geneo4PETSc.cpp:98-194 (text input), :196-379 (decomposition), :381-445 (METIS),
// :447-494 (element weighting), :643-715 (local matrix assembly).  All of this stays on the host.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "common.hpp"

namespace geneo {

// Element mesh: elements in CSR form + one dense row-major n_e x n_e matrix per element
// (the reference's getInput() plug-in ABI, src/geneo4PETSc.cpp:81-85, flattened).
struct Mesh {
  int nbNode = 0;
  std::vector<int64_t> elemPtr;  // [nbElem+1]
  std::vector<int> elemIdx;
  std::vector<int64_t> matPtr;   // [nbElem+1]
  std::vector<double> matVal;
  int grid[3] = {0, 0, 0};       // structured generators: node g sits at (g % grid[0], (g / grid[0]) % grid[1], g / (grid[0] grid[1]))
  int nbElem() const { return (int)elemPtr.size() - 1; }
  void finalize();               // (re)build matPtr from elemPtr
};

// Structured generators producing EXACTLY the element sequence of the reference's tst/laplacian and tst/heat
// plug-ins (tst/laplacian/laplacian.cpp:56-188, tst/laplacian/laplacianServices.cpp:7-94, tst/heat/heat.cpp:24-261)
// without the multimap/set bookkeeping (O(N) time and memory).
struct GridGenOptions {
  int dim = 3, size = 4, weakScaling = 1;
  double inpEps = 1e-4, kappaMax = 1.0;
  std::string kappaInterp;  // "", "lin", "quad", "minmax"
  bool heat = false;
  double lbd = 1.0, dt = 0.1;
  // Box partition (multi-GPU runs, where METIS on the global mesh does not fit a rank): boxK[a] boxes along axis a,
  // element -> box of its first (lower) node.  0 = no box partition.
  int boxK[3] = {0, 0, 0};
  // Sub-mesh: keep only the elements with a node inside [keepLo, keepHi) (node coordinates); keepHi[0] < 0 = everything.
  // Node ids stay GLOBAL, so that every rank numbers nodes and subdomains identically.
  int keepLo[3] = {0, 0, 0}, keepHi[3] = {-1, -1, -1};
};
int parse_gen_args(const std::string& args, GridGenOptions& o);  // same "--size S --dim D ..." grammar
void generate_grid(const GridGenOptions& o, Mesh& m, std::vector<int>* elemPartOut = nullptr);
int grid_edge(const GridGenOptions& o);  // edge length incl. the reference's weak-scaling float truncation
// Graph Laplacian of tst/graph (BASELINE configs[3]): square blocks of blockSize^2 nodes (4-neighbour grids, weight l+1 at
// level l), one central block and 4 blocks per level, each level chained to itself and to the previous one through its
// borders (weight (l+1)/2), optionally every border node tied to node 0 ("ground").  Same element sequence and node
// numbering as the reference plug-in (tst/graph/graph.cpp:38-205), emitted in closed form: O(N), no std::set.
struct GraphGenOptions {
  int size = 4, level = 1, weakScaling = 1;
  double inpEps = 1e-4;
  bool noGround = false;
};
int parse_graph_args(const std::string& args, GraphGenOptions& o);
void generate_graph(const GraphGenOptions& o, Mesh& m);
int read_input_file(const std::string& path, double inpEps, Mesh& m);             // text format A
int read_rhs_file(const std::string& path, int n, std::vector<double>& b);        // text format B

// METIS_PartMeshDual / METIS_PartMeshNodal with the reference's options (MINCONN=1, KWAY, CUT, ncommon=1).
int metis_partition(const Mesh& m, int nbPart, bool dual, std::vector<int>& elemPart, std::vector<int>& nodePart);

// Global node id -> dense index over the nodes a (sub-)mesh actually holds, ascending (identity when it holds them all).
// A rank of a multi-GPU run holds 1/world of the nodes under GLOBAL ids: per-node work arrays are sized by size(), not by
// the global node count (one bit per global node + one counter per 64 is all that scales with the whole problem).
struct NodeIndex {
  bool active = false;
  int nn = 0, nc = 0;
  std::vector<uint64_t> bits;  // presence bitmap
  std::vector<int> pre;        // present nodes before each 64-bit word
  std::vector<int> present;    // dense -> global
  void build(int nbNode, const std::vector<int>& elemIdx);
  int size() const { return active ? nc : nn; }
  int operator()(int g) const {  // -1: not held
    if (!active) return g;
    const uint64_t w = bits[(size_t)g >> 6], b = 1ull << (g & 63);
    if (!(w & b)) return -1;
    return pre[(size_t)g >> 6] + __builtin_popcountll(w & (b - 1));
  }
  int global(int c) const { return active ? present[c] : c; }
};

struct Subdomain {
  int id = 0;
  std::vector<int> nodes;  // sorted global node ids == local numbering (rank in the sorted set)
  std::vector<int> elems;  // sorted global element ids
  std::vector<int> mult;   // node multiplicity, local order
  CsrHost aNeu;            // sum_e (1/elemMult[e]) K_e
  CsrHost aDir;            // (R A R^T): every element coupling between two nodes of the subdomain, full weight
  std::vector<std::vector<int>> intersect;  // [q] -> local indices shared with subdomain q (ascending)
};

struct Decomposition {
  int nbPart = 0, nbNode = 0, nbElem = 0;
  int grid[3] = {0, 0, 0};      // copy of Mesh::grid (0: unstructured) -- lets congruent box subdomains share one nested dissection
  NodeIndex index;              // nodeMult / nodeSubPtr below are indexed by index(global id)
  std::vector<int> nodeMult, elemMult;
  std::vector<Subdomain> subs;  // ALL subdomains' index sets; matrices only for the ones in `mine`
  int64_t nnzNeuTotal = 0;      // "nnz coefs" of the INFO line (sum over all local matrices)
  // node -> subdomains containing it (CSR, ascending subdomain id): ownership for the multi-GPU layout
  std::vector<int64_t> nodeSubPtr;
  std::vector<int> nodeSub;
};

// Pre-decomposed input (what the PETSc plug-in receives: initGenEOPC, hdr/geneo.hpp:30-35 -- the local Neumann matrices
// of a MATIS, the local-to-global maps, optionally the Dirichlet matrices).  Fills multiplicities (createPartitionOfUnity
// input, src/geneo.cpp:977-980), intersections, the node -> subdomain map, and, where a subdomain has no Dirichlet matrix
// yet, A_dir,i = R_i (sum_j R_j^T A_neu,j R_j) R_i^T  (MatConvert + MatCreateSubMatrices, src/geneo.cpp:1692-1699).
// Every subs[p].nodes must be sorted ascending (the reference's local numbering, src/geneo4PETSc.cpp:485-489).
void finish_predecomposed(Decomposition& d);

// ---- multi-GPU layout (one process per GPU) --------------------------------------------------------------
// Rank r holds the subdomains p with subRank[p] == r and OWNS the rows of the nodes whose lowest-numbered subdomain it
// holds (replaces the reference's block-by-index MatCreateIS layout, src/geneo4PETSc.cpp:755).  Local vectors are
// [owned (ascending global id) | ghosts (grouped by owner rank, ascending global id inside a group)]; ghosts are the
// non-owned nodes of the rank's subdomains plus the non-owned columns of its owned operator rows.
struct RankLayout {
  int rank = 0, world = 1, nbNode = 0;
  std::vector<int> subRank;
  std::vector<int> owned, ghost;
  std::vector<int64_t> ghostPtr;            // [world+1] into ghost, by owner rank
  std::vector<std::vector<int>> sendIdx;    // per peer: local OWNED indices it needs (set after the request exchange)
  CsrHost A;                                // owned rows of A = sum_e K_e, local column numbering
  NodeIndex index;                          // global -> local (-1 when absent) through the dense index of the sub-mesh
  std::vector<int> c2l;
  int local(int g) const { const int c = index(g); return c < 0 ? -1 : c2l[c]; }
  int nOwn() const { return (int)owned.size(); }
  int nGhost() const { return (int)ghost.size(); }
};
// The mesh may be a SUB-mesh (global node ids) as long as it holds every element touching a node of the rank's subdomains.
void build_rank_layout(const Mesh& m, const Decomposition& d, const std::vector<int>& subRank, int rank, int world,
                       RankLayout& L);

// Build node/element sets, multiplicities, intersections for every subdomain; assemble aNeu/aDir for
// subdomains p with owner[p] == true (all when owner is empty).
void decompose(const Mesh& m, int nbPart, const std::vector<int>& elemPart, const std::vector<int>& nodePart,
               bool dual, int overlap, const std::vector<char>& owner, Decomposition& d);

}  // namespace geneo
