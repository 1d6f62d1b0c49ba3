// CGR featurisation on the device: expands compact per-atom / per-bond attribute codes of the
// reactant and product sides of a reaction into the reference's feature rows -- reactant one-hots || (product - reactant),
// 78 atom and 14 bond columns (reference cgr_mpnn_3D/utils/graph_features.py:4-63, 177-195).  What needs a chemistry
// toolkit (SMILES -> atoms, bonds, atom maps) stays with the caller; everything after it is table lookups.
#include "../../include/cgr_b200.h"
#include "common.cuh"

namespace {

constexpr int ATOM_FDIM = 39, BOND_FDIM = 7;

struct Tables { cgr_feature_tables_t t; };

// one-hot slot of `v` in `choices` (n entries); unknown values light the extra last slot (graph_features.py:66-80)
__device__ __forceinline__ int slot_of(int v, const int16_t* choices, int n) {
  for (int i = 0; i < n; ++i)
    if (choices[i] == v) return i;
  return n;
}

// the 39 reactant-side (or product-side) atom features of graph_features.py:15-34 as slot indices + scalars
struct AtomSlots { int sym, deg, chg, nh, hyb, arom; double mass; };
__device__ __forceinline__ AtomSlots atom_slots(const int16_t* a, double mass, const cgr_feature_tables_t& t) {
  AtomSlots s;
  s.sym = slot_of(a[0], t.symbol_z, 11);
  s.deg = 12 + slot_of(a[1], t.degrees, 6);
  s.chg = 19 + slot_of(a[2], t.charges, 5);
  s.nh = 25 + slot_of(a[3], t.num_hs, 5);
  s.hyb = 31 + slot_of(a[4], t.hybridizations, 5);
  s.arom = a[5] != 0;
  s.mass = mass * 0.01;                     // Python float arithmetic: double, rounded to float32 on store
  return s;
}
__device__ __forceinline__ double atom_value(const AtomSlots& s, int k) {
  if (k < 37) return (k == s.sym || k == s.deg || k == s.chg || k == s.nh || k == s.hyb) ? 1.0 : 0.0;
  return k == 37 ? (s.arom ? 1.0 : 0.0) : s.mass;
}

__global__ void __launch_bounds__(128) featurize_atoms_kernel(const Tables tb, const int16_t* __restrict__ ar,
                                                              const int16_t* __restrict__ ap,
                                                              const double* __restrict__ mr, const double* __restrict__ mp,
                                                              int64_t n, float* __restrict__ x, int64_t ldx) {
  // one warp per atom: lanes sweep the 78 columns (coalesced row stores)
  const int64_t v = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (v >= n) return;
  const int lane = threadIdx.x & 31;
  const AtomSlots r = atom_slots(ar + v * 6, mr[v], tb.t), p = atom_slots(ap + v * 6, mp[v], tb.t);
  for (int c = lane; c < 2 * ATOM_FDIM; c += 32) {
    const int k = c < ATOM_FDIM ? c : c - ATOM_FDIM;
    const double fr = atom_value(r, k);
    x[v * ldx + c] = (float)(c < ATOM_FDIM ? fr : atom_value(p, k) - fr);      // graph_features.py:178-182
  }
}

// bond code: (type, conjugated, in_ring) with type -1 = the bond does not exist on this side (graph_features.py:50-51),
// 0 single, 1 double, 2 triple, 3 aromatic, anything else: a bond of another type
__device__ __forceinline__ float bond_value(const int8_t* b, int k) {
  if (b[0] < 0) return k == 0 ? 1.f : 0.f;
  if (k == 0) return 0.f;
  if (k <= 4) return b[0] == k - 1 ? 1.f : 0.f;
  return b[k - 4] != 0 ? 1.f : 0.f;           // k = 5: conjugated, 6: in ring
}

__global__ void __launch_bounds__(256) featurize_bonds_kernel(const int8_t* __restrict__ br, const int8_t* __restrict__ bp,
                                                              int64_t e, float* __restrict__ ea) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e * 2 * BOND_FDIM) return;
  const int64_t b = i / (2 * BOND_FDIM);
  const int c = (int)(i - b * 2 * BOND_FDIM);
  const int k = c < BOND_FDIM ? c : c - BOND_FDIM;
  const float fr = bond_value(br + b * 3, k);
  ea[i] = c < BOND_FDIM ? fr : bond_value(bp + b * 3, k) - fr;                 // graph_features.py:189-192
}

}  // namespace

extern "C" int cgr_featurize_cgr(const cgr_feature_tables_t* host_tables, const int16_t* atom_r, const int16_t* atom_p,
                                 const double* mass_r, const double* mass_p, int64_t n_atoms, const int8_t* bond_r,
                                 const int8_t* bond_p, int64_t n_bonds, float* x, int64_t ldx, float* edge_attr,
                                 void* stream) {
  CGR_CHECK_ARG(host_tables && n_atoms >= 0 && n_bonds >= 0 && ldx >= 2 * ATOM_FDIM, "cgr_featurize_cgr: bad argument");
  CGR_CHECK_ARG(n_atoms == 0 || (atom_r && atom_p && mass_r && mass_p && x), "cgr_featurize_cgr: null atom pointer");
  CGR_CHECK_ARG(n_bonds == 0 || (bond_r && bond_p && edge_attr), "cgr_featurize_cgr: null bond pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Tables tb;
  tb.t = *host_tables;
  if (n_atoms > 0) {
    cgr_note_launch("featurize", st, 1);
    featurize_atoms_kernel<<<(unsigned)cgr_ceil_div(n_atoms, 4), 128, 0, st>>>(tb, atom_r, atom_p, mass_r, mass_p, n_atoms, x, ldx);
  }
  if (n_bonds > 0) {
    cgr_note_launch("featurize", st, 1);
    featurize_bonds_kernel<<<(unsigned)cgr_ceil_div(n_bonds * 2 * BOND_FDIM, 256), 256, 0, st>>>(bond_r, bond_p, n_bonds, edge_attr);
  }
  CGR_LAUNCH_CHECK();
  return CGR_OK;
}
