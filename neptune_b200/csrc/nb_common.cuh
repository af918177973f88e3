// nb_common.cuh -- shared definitions for the sm_100a kernels of the NEPTUNE replan hot path.
//
// All device code is written as lane-strided SPMD phases over shared/global state so that the very
// same source also compiles as plain C++ with NL = 1 lane (tests/emul/, a CPU-side debugging
// harness only; the product never loads it).  On the device NL = 32 and a "group" is one warp.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
#define NB_DEV __device__ __forceinline__
#else
#define NB_HD inline
#define NB_DEV inline
#endif

#define NB_NPOL 8
#define NB_DOF_MAX 8
#define NB_NV_MAX 24     // 3 axes x dof_max
#define NB_NFEAT_AX 64   // features per axis: 8 intervals x (4 pos CP + 3 vel CP + 1 acc)
#define NB_SEP_EPS 1e-9

// ---------------------------------------------------------------- warp-group abstraction
template <int NL>
struct Group
{
  int lane;
  NB_HD Group(int l) : lane(l) {}
#if defined(__CUDA_ARCH__)
  NB_DEV void sync() const { __syncwarp(); }
  NB_DEV double sum(double v) const
  {
#pragma unroll
    for (int o = NL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  NB_DEV double max(double v) const
  {
#pragma unroll
    for (int o = NL / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  NB_DEV double min(double v) const
  {
#pragma unroll
    for (int o = NL / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  NB_DEV int any(int p) const { return __any_sync(0xffffffffu, p); }
#else
  void sync() const {}
  double sum(double v) const { return v; }
  double max(double v) const { return v; }
  double min(double v) const { return v; }
  int any(int p) const { return p; }
#endif
};

// ---------------------------------------------------------------- precomputed QP tables
// One table per (n intervals, mode) -- mode 0: terminal v/a equalities kept (first solve,
// solver_gurobi_poly.cpp:660-678), mode 1: dropped and penalised (fallback, :838-847).
// Per axis the equalities E1-E3 leave x_ax = Pm * init3 + Z * w with
// dof = n-2 (mode 0) or n (mode 1) free parameters; Z has orthonormal columns.
struct NbQpTable
{
  int n, mode, dof, has_resid;
  double Z[4 * NB_NPOL][NB_DOF_MAX];       // [interval*4 + coeff][dof]
  double Pm[4 * NB_NPOL][3];               // particular solution map on (b0, c0, d0)
  double C[NB_NFEAT_AX][NB_DOF_MAX];       // feature rows: [interval*8 + j], j: 0-3 pos CP, 4-6 vel CP, 7 acc
  double c0[NB_NFEAT_AX][3];
  double Hr[NB_DOF_MAX][NB_DOF_MAX];       // reduced objective Hessian (per axis)
  double Gr[NB_DOF_MAX][3];                // reduced gradient at w = 0: Gr * init3 + gpf * pf
  double gpf[NB_DOF_MAX];
  double tq[NB_DOF_MAX], tq0[3];           // terminal position error: tq.w + tq0.init3 - pf
  double Rres[2][3];                       // n = 1, mode 0: equality consistency residual map
};

struct NbConsts
{
  double T, W;
  double Ainv[16];   // A_rest_pos_basis_inverse_ (solver_gurobi_poly.cpp:93)
  double V[9];       // A_rest_vel_basis_inverse321_ (:94-97)
  double Ainv01[16]; // inverse MINVO position matrix on [0,1] (neptune.cpp:64)
  double lim_min[3], lim_max[3];
  double v_max, a_max;
  double long_length; // :173
  double drone_radius;
  int N, M, num_pol, S;
  int ent_cap, bp_max, ent_slots;
  int max_iter;
  double tol;
};
