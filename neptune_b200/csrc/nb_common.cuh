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
#define NB_SEP_EPS 1e-6   // slack of the row check of a candidate line (rows are normalised to +-1 at the closest pair); see nb_sep.cuh

// explicitly rounded FP64 operations: no FMA contraction where a sign or a comparison decides an integer result
#if defined(__CUDA_ARCH__)
#define NB_MUL(a, b) __dmul_rn((a), (b))
#define NB_ADD(a, b) __dadd_rn((a), (b))
#define NB_SUB(a, b) __dsub_rn((a), (b))
#else
#define NB_MUL(a, b) ((a) * (b))
#define NB_ADD(a, b) ((a) + (b))
#define NB_SUB(a, b) ((a) - (b))
#endif

// ---------------------------------------------------------------- thread-group abstraction
// NL = 1   : host emulation (one lane runs every phase sequentially)
// NL = 32  : one warp (sync = __syncwarp, reductions by shuffles)
// NL = 128 : one CTA of four warps (sync = __syncthreads, reductions through `red` in shared memory)
template <int NL>
struct Group
{
  int lane;     // thread index within the group
  double* red;  // shared scratch, >= 8 * (NL / 32) doubles when NL > 32
  static constexpr int SUB = (NL >= 256) ? 8 : ((NL >= 128) ? 4 : 1);  // adjacent lanes that may split one item
  NB_HD Group(int l, double* r = nullptr) : lane(l), red(r) {}
#if defined(__CUDA_ARCH__)
  NB_DEV void sync() const
  {
    if (NL <= 32)
      __syncwarp();
    else
      __syncthreads();
  }
  template <int OP>  // 0 sum, 1 max, 2 min
  NB_DEV static double comb(double a, double b)
  {
    return OP == 0 ? a + b : (OP == 1 ? fmax(a, b) : fmin(a, b));
  }
  template <int OP>
  NB_DEV double reduce(double v) const
  {
    constexpr int W = NL < 32 ? NL : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v = comb<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (NL > 32)
    {
      if ((lane & 31) == 0) red[lane >> 5] = v;
      __syncthreads();
      v = red[0];
#pragma unroll
      for (int w = 1; w < NL / 32; w++) v = comb<OP>(v, red[w]);
      __syncthreads();
    }
    return v;
  }
  NB_DEV double sum(double v) const { return reduce<0>(v); }
  NB_DEV double max(double v) const { return reduce<1>(v); }
  NB_DEV double min(double v) const { return reduce<2>(v); }
  // five values at once: (sum, max, min, sum, sum) -- one shared-memory exchange instead of five
  NB_DEV void reduce5(double& s0, double& mx, double& mn, double& s1, double& s2) const
  {
    constexpr int W = NL < 32 ? NL : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1)
    {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (NL > 32)
    {
      const int w = lane >> 5;
      if ((lane & 31) == 0)
      {
        red[5 * w] = s0, red[5 * w + 1] = mx, red[5 * w + 2] = mn, red[5 * w + 3] = s1, red[5 * w + 4] = s2;
      }
      __syncthreads();
      s0 = red[0], mx = red[1], mn = red[2], s1 = red[3], s2 = red[4];
#pragma unroll
      for (int q = 1; q < NL / 32; q++)
      {
        s0 += red[5 * q], mx = fmax(mx, red[5 * q + 1]), mn = fmin(mn, red[5 * q + 2]), s1 += red[5 * q + 3], s2 += red[5 * q + 4];
      }
      __syncthreads();
    }
  }
  // (sum, max) and (max, sum, sum) at once: as many shuffles as values, one shared-memory exchange
  NB_DEV void reduce_sum_max(double& s0, double& mx) const
  {
    constexpr int W = NL < 32 ? NL : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1)
    {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (NL > 32)
    {
      const int w = lane >> 5;
      if ((lane & 31) == 0) red[2 * w] = s0, red[2 * w + 1] = mx;
      __syncthreads();
      s0 = red[0], mx = red[1];
#pragma unroll
      for (int q = 1; q < NL / 32; q++) s0 += red[2 * q], mx = fmax(mx, red[2 * q + 1]);
      __syncthreads();
    }
  }
  NB_DEV void reduce_max_sum_sum(double& mx, double& s0, double& s1) const
  {
    constexpr int W = NL < 32 ? NL : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1)
    {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (NL > 32)
    {
      const int w = lane >> 5;
      if ((lane & 31) == 0) red[3 * w] = mx, red[3 * w + 1] = s0, red[3 * w + 2] = s1;
      __syncthreads();
      mx = red[0], s0 = red[1], s1 = red[2];
#pragma unroll
      for (int q = 1; q < NL / 32; q++) mx = fmax(mx, red[3 * q]), s0 += red[3 * q + 1], s1 += red[3 * q + 2];
      __syncthreads();
    }
  }
  // Split reduction of (sum, max, sum): put() folds the warp by shuffles and parks one partial per warp in half `buf` of
  // `red`; after ANY later barrier of the caller get() folds the partials.  No barrier of its own: consecutive uses
  // alternate the halves or are separated by a barrier of the caller.
  NB_DEV void put(int buf, double s0, double mx, double s1) const
  {
    constexpr int W = NL < 32 ? NL : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1)
    {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (NL > 32 && (lane & 31) == 0)
    {
      double* r = red + buf * 3 * (NL / 32) + 3 * (lane >> 5);
      r[0] = s0, r[1] = mx, r[2] = s1;
    }
    if (NL <= 32 && lane == 0 && red) red[buf * 3] = s0, red[buf * 3 + 1] = mx, red[buf * 3 + 2] = s1;
  }
  NB_DEV void get(int buf, double& s0, double& mx, double& s1) const
  {
    const double* r = red + buf * 3 * (NL > 32 ? NL / 32 : 1);
    s0 = r[0], mx = r[1], s1 = r[2];
#pragma unroll
    for (int q = 1; q < NL / 32; q++) s0 += r[3 * q], mx = fmax(mx, r[3 * q + 1]), s1 += r[3 * q + 2];
  }
  // sum over the SUB adjacent lanes that share an item
  NB_DEV double sub_sum(double v) const
  {
#pragma unroll
    for (int o = 1; o < SUB; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  NB_DEV int any(int p) const { return NL <= 32 ? __any_sync(0xffffffffu, p) : __syncthreads_or(p); }
#else
  void sync() const {}
  double sum(double v) const { return v; }
  double max(double v) const { return v; }
  double min(double v) const { return v; }
  void reduce5(double&, double&, double&, double&, double&) const {}
  void reduce_sum_max(double&, double&) const {}
  void reduce_max_sum_sum(double&, double&, double&) const {}
  double sub_sum(double v) const { return v; }
  void put(int, double, double, double) const {}
  void get(int, double&, double&, double&) const {}   // one lane: the values are the lane's own
  int any(int p) const { return p; }
#endif
};

// ---------------------------------------------------------------- precomputed QP tables
// One table per (n intervals, mode) -- mode 0: terminal v/a equalities kept (first solve,
// solver_gurobi_poly.cpp:660-678), mode 1: dropped and penalised (fallback, :838-847).
// Per axis the equalities E1-E3 leave x_ax = Pm * init3 + Z * w with
// dof = n-2 (mode 0) or n (mode 1) free parameters; Z has orthonormal columns.
#define NB_NPAIR (NB_DOF_MAX * (NB_DOF_MAX + 1) / 2)  // lower-triangle pairs (ca >= cb) of one axis block
struct NbQpTable
{
  int n, mode, dof, has_resid;
  double Z[4 * NB_NPOL][NB_DOF_MAX];       // [interval*4 + coeff][dof]
  double Pm[4 * NB_NPOL][3];               // particular solution map on (b0, c0, d0)
  double C[NB_NFEAT_AX][NB_DOF_MAX];       // feature rows: [interval*8 + j], j: 0-3 pos CP, 4-6 vel CP, 7 acc
  double Ct[NB_DOF_MAX][NB_NFEAT_AX];      // the same, transposed (lanes over features read it conflict-free)
  double c0[NB_NFEAT_AX][3];
  double Hr[NB_DOF_MAX][NB_DOF_MAX];       // reduced objective Hessian (per axis)
  double Gr[NB_DOF_MAX][3];                // reduced gradient at w = 0: Gr * init3 + gpf * pf
  double gpf[NB_DOF_MAX];
  double tq[NB_DOF_MAX], tq0[3];           // terminal position error: tq.w + tq0.init3 - pf
  double pad0;
  double Rres[2][3];                       // n = 1, mode 0: equality consistency residual map
  double PP[NB_NFEAT_AX][NB_NPAIR];        // C[f][ca] * C[f][cb], pair p = ca (ca + 1) / 2 + cb: normal-matrix assembly
};
static_assert(sizeof(NbQpTable) % 16 == 0, "NbQpTable is staged with cp.async.bulk (16-byte granules)");

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
