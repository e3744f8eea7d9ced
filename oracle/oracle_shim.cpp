// oracle_shim.cpp -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// (a) oracle_get_input: dlopen a generator library exporting the reference plug-in ABI
//     `int getInput(std::string const&, unsigned&, unsigned&, vector<unsigned>&, vector<unsigned>&,
//                   vector<vector<double>>&)`   (reference: src/geneo4PETSc.cpp:75-96, documented :1522-1543)
//     and flatten the result into malloc'd C arrays for ctypes.
// (b) oracle_metis_part: METIS mesh partition with the reference's options
//     (reference: src/geneo4PETSc.cpp:381-421: MINCONN=1, PTYPE=KWAY, OBJTYPE=CUT, ncommon=1; nbPart==1 skips METIS).
#include <dlfcn.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <iostream>

typedef int64_t idx_t;  // the CUDA-toolkit libmetis_static.a is built with 64-bit idx_t
extern "C" {
int METIS_SetDefaultOptions(idx_t* options);
int METIS_PartMeshDual(idx_t* ne, idx_t* nn, idx_t* eptr, idx_t* eind, idx_t* vwgt, idx_t* vsize, idx_t* ncommon,
                       idx_t* nparts, float* tpwgts, idx_t* options, idx_t* objval, idx_t* epart, idx_t* npart);
int METIS_PartMeshNodal(idx_t* ne, idx_t* nn, idx_t* eptr, idx_t* eind, idx_t* vwgt, idx_t* vsize, idx_t* nparts,
                        float* tpwgts, idx_t* options, idx_t* objval, idx_t* epart, idx_t* npart);
}
enum { OPT_PTYPE = 0, OPT_OBJTYPE = 1, OPT_MINCONN = 10, PTYPE_KWAY = 1, OBJTYPE_CUT = 0, NOPTIONS = 40 };

typedef int (*getInputFn)(std::string const&, unsigned int&, unsigned int&, std::vector<unsigned int>&,
                          std::vector<unsigned int>&, std::vector<std::vector<double>>&);

extern "C" int oracle_get_input(const char* lib, const char* args, uint32_t* nbElem, uint32_t* nbNode,
                                uint32_t** elemPtr, uint32_t** elemIdx, uint64_t* nIdx, double** vals, uint64_t* nVals) {
  void* h = dlopen(lib, RTLD_LAZY | RTLD_LOCAL);
  if (!h) { std::cerr << "oracle_shim: dlopen KO - " << dlerror() << std::endl; return 1; }
  getInputFn fn = (getInputFn)dlsym(h, "getInput");
  if (!fn) { std::cerr << "oracle_shim: dlsym KO" << std::endl; dlclose(h); return 1; }
  unsigned int ne = 0, nn = 0;
  std::vector<unsigned int> ep, ei;
  std::vector<std::vector<double>> em;
  int rc = fn(std::string(args), ne, nn, ep, ei, em);
  if (rc != 0) { dlclose(h); return 1; }
  // NB: the laplacian/heat generators never push the final elemPtr sentinel twice; they keep a running pointer.
  if (ep.size() != (size_t)ne + 1) { std::cerr << "oracle_shim: bad elemPtr size" << std::endl; dlclose(h); return 1; }
  *nbElem = ne; *nbNode = nn;
  *elemPtr = (uint32_t*)malloc(sizeof(uint32_t) * ep.size());
  memcpy(*elemPtr, ep.data(), sizeof(uint32_t) * ep.size());
  *elemIdx = (uint32_t*)malloc(sizeof(uint32_t) * (ei.size() ? ei.size() : 1));
  memcpy(*elemIdx, ei.data(), sizeof(uint32_t) * ei.size());
  *nIdx = ei.size();
  uint64_t tot = 0;
  for (auto& m : em) tot += m.size();
  *vals = (double*)malloc(sizeof(double) * (tot ? tot : 1));
  uint64_t o = 0;
  for (auto& m : em) { memcpy(*vals + o, m.data(), sizeof(double) * m.size()); o += m.size(); }
  *nVals = tot;
  dlclose(h);
  return 0;
}

extern "C" void oracle_free(void* p) { free(p); }

extern "C" int oracle_metis_part(int dual, int64_t ne, int64_t nn, const int64_t* eptr, const int64_t* eind,
                                 int64_t nparts, int64_t* epart, int64_t* npart) {
  if (nparts == 1) {
    for (int64_t e = 0; e < ne; e++) epart[e] = 0;
    for (int64_t n = 0; n < nn; n++) npart[n] = 0;
    return 0;
  }
  idx_t options[NOPTIONS];
  METIS_SetDefaultOptions(options);
  options[OPT_MINCONN] = 1; options[OPT_PTYPE] = PTYPE_KWAY; options[OPT_OBJTYPE] = OBJTYPE_CUT;
  idx_t obj = 0, ncommon = 1, NE = ne, NN = nn, NP = nparts;
  std::vector<idx_t> ep(eptr, eptr + ne + 1), ei(eind, eind + eptr[ne]);
  int rc;
  if (dual) rc = METIS_PartMeshDual(&NE, &NN, ep.data(), ei.data(), NULL, NULL, &ncommon, &NP, NULL, options, &obj, epart, npart);
  else      rc = METIS_PartMeshNodal(&NE, &NN, ep.data(), ei.data(), NULL, NULL, &NP, NULL, options, &obj, epart, npart);
  return rc == 1 ? 0 : 1;  // METIS_OK == 1
}
