// FE assembly state shared by nm_pattern.cpp (host: topology, numbering, patterns) and
// nm_assembly.cu (device: element integration + scatter).  Mirrors the parts of the reference's globals
// unstrM / CGM (src/mod_datatype.f90, src/mod_cg_datatype.f90) the assembly needs.
#pragma once
#include "nm_internal.h"

enum NmMatId { NM_MAT_A = 0 /* A, or Ad in the fluid case */, NM_MAT_B = 1, NM_MAT_E = 2, NM_MAT_ET = 3, NM_MAT_AP = 4, NM_NMAT = 5 };

struct NmPattern {
  bool present = false;
  int nrow = 0;                            // local rows
  std::vector<int> rowdist, coldist;       // global offsets per rank (sizdist of rows / of columns)
  std::vector<int> ia, ja;                 // 0-based local row pointers, 0-based GLOBAL column ids
  std::vector<double> val;                 // filled by nm_fem_assemble
};

struct NmFem {
  int ntet = 0, nvert = 0, porder = 1, pNp = 4, nn = 0;
  int nproc = 1, rank = 0;
  bool fsexist = false, purefluid = false, fluidcase = false;
  std::vector<int> ele, neigh;             // [ntet][4], 0-based, -1 = boundary
  std::vector<double> node;                // [nvert][3]
  std::vector<int> t2n;                    // [ntet][pNp] global node ids, reference local order
  std::vector<int> edges;                  // P2: [nedge][2] endpoint vertices of edge node nvert+e
  std::vector<int> vstat;                  // [nn] 0 solid 1 fluid 2 interface (+3 for edge nodes)
  std::vector<unsigned char> efl;          // [ntet] fluid element (geometry test, vs < 1e-6)
  std::vector<int> n2e_ptr, n2e, v2v_ptr, v2v;
  std::vector<int> part, order, vnum, pnum, vstt, pstt, vtxdist, sizdist, psizdist;
  int N = 0, Np = 0;
  std::vector<int> lelist;                 // elements touching an owned node, ascending
  NmPattern pat[NM_NMAT];
};

NmFem* nm_fem_build(int ntet, int nvert, const int* ele, const int* neigh, const double* node, int porder,
                    const double* vs, int nproc, const int* part, int rank);
void nm_fem_assemble(NmFem& F, int job, const double* vp, const double* vs, const double* rho, const double* g0);
