// nb_search.cuh -- K0: the front end of a replan, KinodynamicSearch (reference
// neptune/src/kinodynamic_search.cpp), one CTA per planning agent.  SURVEY.md section 8(f) #1.
//
// Replaces
//   KinodynamicSearch::setUp (entangle form)                   kinodynamic_search.cpp:190-249
//   KinodynamicSearch::run                                     :1629-1827
//   KinodynamicSearch::expandAndAddToQueue (root / node)       :1240-1385 / :1045-1228
//   KinodynamicSearch::entanglesWithOtherAgents                :805-895
//   eu::getTetherLength, eu::getBendPt2dwIdx                   entangle_utils.cpp:1724-1743, :1681-1707
//   KinodynamicSearch::collidesWithObstacles2dSolve            :1514-1580
//   KinodynamicSearch::collidesWithBases2d                     :1583-1627
//   recoverPwpOut, recoverEntStateVector                       :521-553, :582-603
//   getIz, power_int, CompareCost                              :2006-2031, kinodynamic_search.hpp:163-179
//
// Mapping onto the SM.  A best-first search is a sequential chain of pops, but every pop fans out.  One CTA of 32
// warps per agent:
//   * the 25 jerk primitives of the popped node (kinematics + admissibility) are evaluated one LANE per child from
//     a per-launch table of everything that depends only on the jerk value;
//   * each child's S-step entanglement chain runs on its own WARP, lanes over the N + M tethers: crossing lists by
//     ballot-ordered append (wedge products shared between steps and, for the base-side test, between all nodes of
//     the search), the list automaton on lane 0, then the tether-length test and the voxel key + node-map lookup;
//   * the collision tests of the popped node (GJK against the N - 1 window hulls, the M static obstacles and the
//     bases) run on the 7 remaining warps CONCURRENTLY with its speculatively evaluated children;
//   * one warp then replays the reference's sequential loop over the children in all_combinations_ order -- the
//     "better node" replacement rule with its ran_trigger parity, node allocation, push_heap -- with child ch in
//     the registers of lane ch, so node numbering, heap layout and therefore the pop order are those of a
//     sequential run; the accepted children's payloads are then copied to the node pool by the child warps.
// The open list (heap of node ids), the (f = g + bias h, h) keys its comparisons read, the children's scratch and
// the per-agent inputs of the chain live in shared memory while they fit (they do up to ~100 tethers) and in
// global memory otherwise: same code, generic pointers.  Node payloads (kinematics, entanglement lists) stay in
// global memory; a node's active_cases array is not stored: active = active_A - count_A(id) + count_node(id)
// because every change of active_cases is paired with an append/erase of alphas (entangle_utils.cpp:1402-1534).
//
// std::priority_queue is restated as libstdc++'s __push_heap / __adjust_heap (CompareCost is not a strict
// weak order, so the pop order depends on the sift algorithm).  Compiled with --fmad=false: every FP64
// result feeds a discrete decision (voxel keys, cost ties, admissibility tests), so products and sums are
// rounded separately, as in the CPU build of the reference.
//
// Non-deterministic reference state is injected: the order of the jerk samples (comb) and the number of
// pops that stands in for the wall-clock budget (max_exp).
#pragma once
#include <float.h>

#include "../../include/neptune_b200.h"
#include "nb_common.cuh"
#include "nb_entangle.cuh"
#include "nb_hull.cuh"

#define NB_SEARCH_MAXCHILD 25
#define NB_SEARCH_KIN 14  // end[6], coeff_x[4], coeff_y[4]

struct NbInt4
{
  int x, y, z, w;
};

struct NbSearchPar
{
  int N, M, num_pol, S, ns, nchild;
  double T, x_min, x_max, y_min, y_max, v_max, a_max, j_max, voxel, bias, goal_size, tether;
  int enable_entangle, use_not_reaching, max_nodes, max_exp, ecap, out_cap, es_cap, bp_max, hcap, tcap;
  double Ainv[16], V[9];
};

struct NbSearchArgs
{
  NbSearchPar p;
  // inputs
  const int* agent_id;
  const double* init;      // [B][6]
  const double* goal;      // [B][2]
  const double* coeffs_z;  // [B][8][4]
  const int* group;        // [B] or null
  const double* hull_xy;   // [G][N][8][24][2]
  const int* hull_cnt;     // [G][N][8]
  const double* samp;      // [G][N][num_pol][S+1][2]
  const uint8_t* known;    // [B][N]
  const int64_t* st_ptr;
  const double* st_xy;
  const double* strep;
  const double* st_longest;
  int strep_per_agent;   // 1: strep is [N][M][2][2] and st_longest [N][M][2], indexed by the planning agent's id - 1
  const double* pb;
  const int* bp_cnt;
  const double* bp_xy;
  nb_ent_state es;  // entangle_state_A, stride es_cap
  const uint8_t* comb;
  int comb_shared;
  // outputs
  int* status;
  int* solved;
  int* n_int;
  double* coeff;     // [B][3][8][4]
  nb_ent_state esv;  // [B][9], stride out_cap
  int* stats;        // [B][4]
  double* cost;      // [B]
  // workspace, per agent
  NbInt4* nd_meta;   // [max_nodes]: prev, index, state, n_alpha | n_bend << 16
  double* nd_kin;    // [max_nodes][14]
  int* nd_alpha;     // [max_nodes][ecap][2]
  double* nd_beta;   // [max_nodes][ecap]
  int* nd_bend;      // [max_nodes][ecap]
  NbInt4* hash;      // [hcap]: ix, iy, iz, id + 1
  int* heap_g;       // [max_nodes] when the open list does not fit in shared memory
  double* gh_g;      // [max_nodes][2] (f = g + bias h, h): the keys CompareCost reads
  double* nd_g;      // [max_nodes] g of every node
  int* ch_int;       // [nchild][2 tcap + 2 NA + 3 ecap] + [NA] parent active
  double* ch_dbl;    // [nchild][ecap + 18] + base squares
  uint8_t* fcode_g;  // [num_pol][8][N] per agent
  int* err;
  long long* prof;   // optional [B][16]: SM cycles thread 0 spent per phase (measurement hook)
};

// parameters of one launch from the ABI structs (host side; shared by nb_capi.cu and the test emulation)
inline void nb_search_fill_par(const nb_params& par, const nb_search_params& sp, const NbConsts& cs, NbSearchPar* p)
{
  const int NA = par.num_agents + par.num_static;
  p->N = par.num_agents, p->M = par.num_static, p->num_pol = par.num_pol, p->S = par.samples;
  p->ns = sp.num_samples, p->nchild = sp.num_samples * sp.num_samples;
  p->T = par.T_span;
  p->x_min = par.lim_min[0], p->x_max = par.lim_max[0], p->y_min = par.lim_min[1], p->y_max = par.lim_max[1];
  p->v_max = par.v_max, p->a_max = par.a_max, p->j_max = sp.j_max, p->voxel = sp.voxel_size, p->bias = sp.bias;
  p->goal_size = sp.goal_size, p->tether = par.tether_length;
  p->enable_entangle = sp.enable_entangle_check, p->use_not_reaching = sp.use_not_reaching_soln;
  p->max_nodes = sp.max_nodes, p->max_exp = sp.max_expansions, p->ecap = sp.ecap;
  p->out_cap = par.ent_cap, p->es_cap = par.ent_cap, p->bp_max = par.bp_max;
  p->hcap = 64;
  while (p->hcap < 2 * sp.max_nodes) p->hcap *= 2;
  p->tcap = NA;  // a crossing list longer than N + M entangles in any case (:844-848)
  for (int k = 0; k < 16; k++) p->Ainv[k] = cs.Ainv[k];
  for (int k = 0; k < 9; k++) p->V[k] = cs.V[k];
}

// shared-memory arena: per-agent working sets are placed in shared memory in priority order while they fit,
// and stay in global memory (same code, generic pointers) when they do not (large N + M)
struct NbArena
{
  unsigned char* p;
  size_t left;
  NB_HD void* take(size_t bytes)
  {
    bytes = (bytes + 15) & ~(size_t)15;
    if (!p || bytes > left) return nullptr;
    void* r = p;
    p += bytes;
    left -= bytes;
    return r;
  }
};

// ints of scratch per child: S crossing lists [S][2 tcap] | (id, old active) pairs [2 tcap] | active [NA] | alpha [2 ecap] | bend [ecap]
NB_HD int nb_search_ch_stride(const NbSearchPar& p) { return (2 * p.S + 2) * p.tcap + (p.N + p.M) + 3 * p.ecap; }

// doubles of scratch per agent: beta lists of the children, then the base squares [N][4][2]
#define NB_SEARCH_PTS 28  // per child: (S + 1) sample positions (S <= 8) [0..17], step lengths [18 + s], arc length [27]
// Strides of the per-tether arrays that lanes read side by side are padded to an ODD number of doubles in shared
// memory: with an even stride (64 doubles of samples, 16 of bend points, 48 of hull vertices) all 32 lanes of a
// warp hit the same bank.
#define NB_SEARCH_BSQ_STRIDE 9
#define NB_SEARCH_HSTAGE_STRIDE (NB_HMAX * 2 + 1)
NB_HD size_t nb_search_chd_stride(const NbSearchPar& p) { return (size_t)p.nchild * (p.ecap + NB_SEARCH_PTS) + (size_t)p.N * NB_SEARCH_BSQ_STRIDE; }
NB_HD int nb_search_odd(int n) { return n | 1; }
NB_HD size_t nb_search_fcode_bytes(const NbSearchPar& p) { return (size_t)p.num_pol * 8 * p.N; }

// one evaluated child of the node being expanded
struct NbChildRec
{
  int valid, ix, iy, iz, n_alpha, n_bend, accept_id, found, f_state, f_index;
  double kin[NB_SEARCH_KIN], g, h, f;
  int prim_ok, pad2;
};

struct NbSearchCtl
{
  int done, status, cur, n_path, closest, n_used, heap_n, pops, ran_trigger, goal_occupied, first_new, overflow;
  int n_acc, acc[NB_SEARCH_MAXCHILD];  // children accepted so far in the expansion being resolved
  int ovf_iter;                        // storage overflow seen while evaluating the children of the current node
  int cmax;                            // measurement: slowest child chain of the current expansion (cycles)
  int hit[2], invalid;                 // the popped node: collides (flag per iteration parity) / has an active case above 1
  double smallest;
  int flag[NB_SEARCH_MAXCHILD][12];  // per child: result, n_alpha, n_bend, -, entries of the S crossing lists
};

// per-launch constants that depend only on the parameters and the jerk order (built once by thread 0 with the
// very expressions of the reference, so that the per-child chains start from them)
struct NbSearchTab
{
  double j6[5], j2[5], jt[5], jc[5];  // ji*tau*tau*tau/6, ji*tau*tau/2, ji*tau, ji/6 for the ns jerk values
  double tt[8][3];                    // sampled_time_vector_ (:116-127): t, t*t, t*t*t for j = 1..S-1
  int cjx[NB_SEARCH_MAXCHILD], cjy[NB_SEARCH_MAXCHILD];
};

#define NB_SEARCH_NSTAGE 9

// per-agent view of the arguments
struct NbSearchCtx
{
  const NbSearchPar* p;
  int b, self, NA;
  double init[6], goal[2];
  const double* hull_xy;
  const int* hull_cnt;
  const double* samp;
  const uint8_t* known;
  const uint8_t* comb;
  const int64_t* st_ptr;
  const double* st_xy;
  const double* st_longest;
  NbEntCtx ecx;
  // entangle_state_A
  int a_na, a_nb;
  const int* a_alpha;
  const double* a_beta;
  const int* a_bend;
  const int* a_active;
  // pool
  NbInt4* meta;
  double* kin;
  int* alpha;
  double* beta;
  int* bend;
  NbInt4* hash;
  int* heap;
  double* gh;   // (f, h) per node
  double* ng;   // g per node
  int* ch_int;
  double* ch_dbl;
  int* par_act;
  double* base_sq;  // [N][4][2] squares around the bases (collidesWithBases2d :1610-1612)
  int samp_stride;     // doubles between the sample blocks of consecutive agents (padded to an odd count in shared memory)
  int hstage_stride, bsq_stride;
  uint8_t* fcode;      // [num_pol][8][N] base-crossing code of every (window, step, single-bend tether), built once
  int multi_bend;      // some known tether has bend points besides its base: the generic chain is used
  double* hull_stage;  // [N][24][2] shared-memory copy of the hulls of one window (null: read them from global)
  int ch_stride;
  long long* prof;
};

// Everything the threads of one search share lives here (shared memory on the device): parameters and the
// per-agent context are read through it instead of per-thread copies, so the kernel keeps no stack frame --
// with 800 threads per CTA any per-thread local memory would overflow the part of L1 left beside the arena.
struct NbSearchShared
{
  NbSearchCtl ctl;
  NbSearchTab tab;
  NbSearchPar par;
  NbSearchCtx cx;
  const void* stage_src[NB_SEARCH_NSTAGE];
  void* stage_dst[NB_SEARCH_NSTAGE];
  size_t stage_bytes[NB_SEARCH_NSTAGE];
  int stage_rows[NB_SEARCH_NSTAGE], stage_src_stride[NB_SEARCH_NSTAGE], stage_dst_stride[NB_SEARCH_NSTAGE];  // row-wise re-pack (doubles)
  int path[NB_NPOL];
  // the node being expanded (written by thread 0 after the pop): kinematics, control points, list views
  double par_kin[NB_SEARCH_KIN], par_cps[8], par_g, goal_hull[8];
  int par_index, par_na, par_nb;
  const int *par_alpha, *par_bend;
  const double* par_beta;
  NbChildRec rec[NB_SEARCH_MAXCHILD];
};

// The kernel is latency-bound on short sequential phases that run once per pop, so its code size matters more than
// call overhead: with everything inlined it was 277 KB of SASS, far beyond the instruction caches, and every phase
// started with a run of instruction-fetch misses.  The multi-site helpers below are therefore real functions on the
// device (one body each); on the host (single-lane emulation) they stay inline.
#if defined(__CUDA_ARCH__)
#define NB_OUTLINE __device__ __noinline__
#else
#define NB_OUTLINE inline
#endif

NB_OUTLINE double nb_norm2(double x, double y) { return sqrt(x * x + y * y); }
NB_OUTLINE bool nb_s_gjk(const double* v1, int n1, const double* v2, int n2) { return nb_gjk_collision(v1, n1, v2, n2); }
NB_OUTLINE int nb_s_add_alpha_beta(int* toadd, int nadd, NbEntState* es, const double* pk, const NbEntCtx* cx)
{
  return nb_add_alpha_beta(toadd, nadd, *es, pk, *cx);
}
NB_OUTLINE void nb_s_update_bend_pts(NbEntState* es, const double* pk1, const NbEntCtx* cx) { nb_update_bend_pts(*es, pk1, *cx); }

#if defined(__CUDA_ARCH__)
#define NB_TICK(slot)                                        \
  if (cta.tid == 0 && a.prof)                                \
  {                                                          \
    const long long now_ = clock64();                        \
    a.prof[(size_t)b * 16 + (slot)] += now_ - tick_;          \
    tick_ = now_;                                            \
  }
#define NB_TICK_INIT long long tick_ = clock64();
#define NB_CTICK(slot)                                       \
  if (c.prof && ch == 0 && g.lane == 0)                      \
  {                                                          \
    const long long now_ = clock64();                        \
    c.prof[(slot)] += now_ - ctick_;                         \
    ctick_ = now_;                                           \
  }
#define NB_CTICK_INIT long long ctick_ = clock64();
#else
#define NB_CTICK(slot)
#define NB_CTICK_INIT
#define NB_TICK(slot)
#define NB_TICK_INIT
#endif

NB_HD int nb_hull_count(const NbSearchCtx& c, int o, int i)
{
  if (o == c.self || !c.known[o]) return 0;
  return c.hull_cnt[o * NB_NPOL + i];
}

// CompareCost (kinodynamic_search.hpp:163-179): true when l has lower priority than r
// The cost g + bias h of a node is evaluated once, when (g, h) are written, and stored beside h: the same
// double the reference recomputes in every comparison.
NB_HD bool nb_cmp_cost(const NbSearchCtx& c, int l, int r)
{
  const double cl = c.gh[2 * l], cr = c.gh[2 * r];
  if (fabs(cl - cr) < 1e-5) return c.gh[2 * l + 1] > c.gh[2 * r + 1];
  return cl > cr;
}

// libstdc++ std::__push_heap
NB_HD void nb_heap_push_at(const NbSearchCtx& c, int hole, int top, int value)
{
  int parent = (hole - 1) / 2;
  while (hole > top && nb_cmp_cost(c, c.heap[parent], value))
  {
    c.heap[hole] = c.heap[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  c.heap[hole] = value;
}

// the same sift with the keys (f, h) of `value` already in registers (two shared-memory loads less per level)
NB_HD void nb_heap_push_at_keys(const NbSearchCtx& c, int hole, int top, int value, double fv, double hv)
{
  int parent = (hole - 1) / 2;
  while (hole > top)
  {
    const int pid = c.heap[parent];
    const double cl = c.gh[2 * pid];
    const bool lower = fabs(cl - fv) < 1e-5 ? c.gh[2 * pid + 1] > hv : cl > fv;  // CompareCost(parent, value)
    if (!lower) break;
    c.heap[hole] = pid;
    hole = parent;
    parent = (hole - 1) / 2;
  }
  c.heap[hole] = value;
}

// top() + pop(): std::pop_heap (= __pop_heap, __adjust_heap) + pop_back
NB_HD int nb_heap_pop(const NbSearchCtx& c, int& heap_n)
{
  const int top = c.heap[0];
  if (heap_n > 1)
  {
    const int last = heap_n - 1;
    const int value = c.heap[last];
    c.heap[last] = c.heap[0];
    const int len = last;
    int hole = 0, child = 0;
    while (child < (len - 1) / 2)
    {
      child = 2 * (child + 1);
      if (nb_cmp_cost(c, c.heap[child], c.heap[child - 1])) child--;
      c.heap[hole] = c.heap[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
      child = 2 * (child + 1);
      c.heap[hole] = c.heap[child - 1];
      hole = child - 1;
    }
    nb_heap_push_at_keys(c, hole, 0, value, c.gh[2 * value], c.gh[2 * value + 1]);
  }
  heap_n--;
  return top;
}

NB_HD uint32_t nb_hash3(int ix, int iy, int iz)
{
  uint32_t h = (uint32_t)ix * 0x9E3779B1u;
  h ^= (uint32_t)iy * 0x85EBCA77u + (h << 6) + (h >> 2);
  h ^= (uint32_t)iz * 0xC2B2AE3Du + (h << 6) + (h >> 2);
  return h;
}

NB_HD int nb_hash_find(const NbSearchCtx& c, int ix, int iy, int iz)
{
  const uint32_t mask = (uint32_t)(c.p->hcap - 1);
  uint32_t q = nb_hash3(ix, iy, iz) & mask;
  for (;;)
  {
    const NbInt4 e = c.hash[q];
    if (e.w == 0) return -1;
    if (e.x == ix && e.y == iy && e.z == iz) return e.w - 1;
    q = (q + 1) & mask;
  }
}

// unordered_map::insert keeps an existing mapping
NB_HD void nb_hash_insert(const NbSearchCtx& c, int ix, int iy, int iz, int id)
{
  const uint32_t mask = (uint32_t)(c.p->hcap - 1);
  uint32_t q = nb_hash3(ix, iy, iz) & mask;
  for (;;)
  {
    const NbInt4 e = c.hash[q];
    if (e.w == 0)
    {
      NbInt4 n;
      n.x = ix, n.y = iy, n.z = iz, n.w = id + 1;
      c.hash[q] = n;
      return;
    }
    if (e.x == ix && e.y == iy && e.z == iz) return;
    q = (q + 1) & mask;
  }
}

// KinodynamicSearch::power_int (:2016-2031)
NB_HD uint32_t nb_power_int(uint32_t base, uint32_t exponent)
{
  if (exponent == 0) return 1;
  if (base < 2) return base;
  uint32_t result = 1;
  for (uint32_t term = base;; term = term * term)
  {
    if (exponent % 2 != 0) result *= term;
    exponent /= 2;
    if (exponent == 0) break;
  }
  return result;
}

// unsigned int ix = round(x / voxel), then Eigen::Vector2i(ix, iy): two's-complement wrap
NB_HD int nb_voxel_index(double x, double voxel) { return (int)(uint32_t)(long long)round(x / voxel); }

// eu::getTetherLength (entangle_utils.cpp:1724-1743)
NB_HD double nb_tether_length(const NbSearchCtx& c, const NbEntState& es, const double* pk1)
{
  const int N = c.ecx.N;
  double length = 0.0;
  double prev[2] = { c.ecx.pb[2 * c.self], c.ecx.pb[2 * c.self + 1] };
  for (int i = 0; i < es.n_bend; i++)
  {
    const int q = es.bend[i];
    const int id = es.alpha[2 * q], cs = es.alpha[2 * q + 1];
    double bp[2], comp = 0.0;
    if (id <= N)
      bp[0] = c.ecx.pb[2 * (id - 1)], bp[1] = c.ecx.pb[2 * (id - 1) + 1];
    else
    {
      bp[0] = c.ecx.strep[4 * (id - N - 1) + 2 * cs], bp[1] = c.ecx.strep[4 * (id - N - 1) + 2 * cs + 1];
      comp = c.st_longest[2 * (id - N - 1) + cs];
    }
    length += nb_norm2(bp[0] - prev[0], bp[1] - prev[1]) + 2 * comp;
    prev[0] = bp[0], prev[1] = bp[1];
  }
  length += nb_norm2(pk1[0] - prev[0], pk1[1] - prev[1]);
  return length;
}

// KinodynamicSearch::entanglesWithOtherAgents (:805-895) for one child, group-cooperative.
// es works on the child's scratch lists / active array.  Returns 1 entangles, 0 fine, -1 storage overflow.
template <int NL>
NB_HD int nb_search_entangles(const Group<NL>& g, const NbSearchCtx& c, NbEntState& es, const double* kin, int index,
                              int* toadd, int* act_old, int* flag, double* arc_length, int ch, const double (*tt)[3])
{
  NB_CTICK_INIT
  const NbSearchPar& p = *c.p;
  const int N = p.N, NA = c.NA, S = p.S, num_pol = p.num_pol;
  const double* cx = kin + 6;
  const double* cy = kin + 10;
  double pk[2] = { cx[3], cy[3] }, pk1[2] = { cx[3], cy[3] };
  double arc = 0.0;
  const int stride = c.samp_stride;
  for (int j = 1; j <= S; j++)
  {
    if (j < S)
    {
      const double t = tt[j][0], t2 = tt[j][1], t3 = tt[j][2];
      pk1[0] = cx[0] * t3 + cx[1] * t2 + cx[2] * t + cx[3];
      pk1[1] = cy[0] * t3 + cy[1] * t2 + cy[2] * t + cy[3];
    }
    else
      pk1[0] = kin[0], pk1[1] = kin[1];
    arc += nb_norm2(pk1[0] - pk[0], pk1[1] - pk[1]);
    const double *pik, *pik1;
    if (index > num_pol)
    {
      pik = c.samp + ((size_t)(num_pol - 1) * (S + 1) + S) * 2;
      pik1 = pik;
    }
    else
    {
      pik = c.samp + ((size_t)(index - 1) * (S + 1) + (j - 1)) * 2;
      pik1 = c.samp + ((size_t)(index - 1) * (S + 1) + j) * 2;
    }
    NB_CTICK(12)
    const int nadd = nb_collect_toadd<NL>(g, c.ecx, pk, nullptr, pk1, pik, stride, pik1, stride, c.known, toadd, p.tcap);
    NB_CTICK(13)
    if (nadd < 0) return 1;  // more crossings than tcap >= N+M: over the list bound of :844-848 in any case
    if (g.lane == 0)
    {
      int r = 0;
      if (es.n_alpha + nadd > NA)
        r = 1;
      else
      {
        // active_cases_old (:813, :880) is only compared where active_cases can have grown, i.e. at the ids of
        // alphasToAdd: remember (id, value before) for those instead of copying the whole array every step
        for (int i = 0; i < nadd; i++) act_old[2 * i] = toadd[2 * i], act_old[2 * i + 1] = es.active[toadd[2 * i] - 1];
        if (nb_s_add_alpha_beta(toadd, nadd, &es, pk, &c.ecx))
          r = -1;
        else
        {
          for (int i = 0; i < nadd; i++)
          {
            const int a = act_old[2 * i] - 1, was = act_old[2 * i + 1];
            if (a >= N) continue;
            if (was < 2 && es.active[a] >= 2) r = 1;
            if (was >= 2 && es.active[a] > was) r = 1;
          }
          if (r == 0) nb_s_update_bend_pts(&es, pk1, &c.ecx);
        }
      }
      flag[0] = r;
      flag[1] = es.n_alpha;
      flag[2] = es.n_bend;
    }
    g.sync();
    const int r = flag[0];
    es.n_alpha = flag[1];
    es.n_bend = flag[2];
    g.sync();
    NB_CTICK(14)
    if (r != 0) return r;
    pk[0] = pk1[0], pk[1] = pk1[1];
  }
  *arc_length = arc;
  if (nb_tether_length(c, es, pk1) > p.tether) return 1;  // :884-891
  return 0;
}

// ---- the same chain with the wedge products shared between steps and between children (single-bend tethers).
// For a tether whose only bend point is its base b (nbend == 1, the normal case):
//   * the base-side test of eu::entangleHSigToAddAgentInd (:1145-1180) uses wedge(pb_self, sample, b): it does not
//     depend on the node at all, only on (window, step, tether) -> its outcome is the code c.fcode, built once per
//     search: 0 no crossing, 1 crossing without entry (a < 0), 2 entry case 1, 3 entry case 0;
//   * the robot-side test uses c1 = wedge(P_{s-1}, sample_{s-1}, b), c2 = wedge(P_s, sample_s, b): c2 of step s is c1
//     of step s+1 (same operands, same operation), so S + 1 products serve S steps.  Same for static obstacles.
// Values and their order of evaluation are those of the generic path, so the lists are bit-identical.
// Phase 1 (lanes over tethers) fills one crossing list per step; phase 2 (lane 0) runs the list automaton.
NB_HD uint8_t nb_search_fcode_one(const double* pb_self, const double* pik, const double* pik1, const double* b)
{
  double fb[2], fc[2];
  const double f1 = nb_wedge(pb_self, pik, b, fb, fc);
  const double f2 = nb_wedge(pb_self, pik1, b, nullptr, nullptr);
  if (!nb_neg_product(f1, f2)) return 0;
  const double a = nb_cross_ratio(fb, fc);
  return a < 0 ? 1 : (a < 1 ? 2 : 3);
}

template <int NL>
NB_HD int nb_search_entangles_fast(const Group<NL>& g, const NbSearchCtx& c, NbEntState& es, const double* kin, int index,
                                   int* toadd, int* act_old, int* flag, double* arc_length, int ch, const double (*tt)[3],
                                   double* pts)
{
  NB_CTICK_INIT
  const NbSearchPar& p = *c.p;
  const int N = p.N, M = p.M, NA = c.NA, S = p.S, num_pol = p.num_pol;
  const double* cx = kin + 6;
  const double* cy = kin + 10;
  // sample positions P_0 .. P_S of this child (:808, :818-819)
  if (g.lane == 0)
  {
    pts[0] = cx[3], pts[1] = cy[3];
    for (int j = 1; j < S; j++)
    {
      const double t = tt[j][0], t2 = tt[j][1], t3 = tt[j][2];
      pts[2 * j] = cx[0] * t3 + cx[1] * t2 + cx[2] * t + cx[3];
      pts[2 * j + 1] = cy[0] * t3 + cy[1] * t2 + cy[2] * t + cy[3];
    }
    pts[2 * S] = kin[0], pts[2 * S + 1] = kin[1];
    for (int j = 1; j <= S; j++) flag[3 + j] = 0;  // entries of the crossing list of step j (-1: over tcap)
  }
  g.sync();
  // lengths of the S steps (:820), one lane each; lane 0 adds them up in step order below
  for (int s2 = 1 + g.lane; s2 <= S; s2 += NL)
    pts[18 + s2] = nb_norm2(pts[2 * s2] - pts[2 * (s2 - 1)], pts[2 * s2 + 1] - pts[2 * (s2 - 1) + 1]);
  NB_CTICK(12)
  const int stride = c.samp_stride;
  const bool past = index > num_pol;
  const double* samp0 = past ? c.samp + ((size_t)(num_pol - 1) * (S + 1) + S) * 2 : c.samp + (size_t)(index - 1) * (S + 1) * 2;
  const uint8_t* fcode = c.fcode + (size_t)(past ? 0 : index - 1) * 8 * N;
  int* nlist = flag + 4;
  for (int base = 0; base < NA; base += NL)
  {
    const int j = base + g.lane;
    const bool agent = j < N && j != c.self && c.known[j];
    const bool stat = j >= N && j < NA;
    // per step: number of entries (0..2) and their cases, 6 bits per step
    unsigned long long code = 0;
    if (agent || stat)
    {
      const double* b = agent ? c.ecx.bp_xy + (size_t)c.ecx.bp_stride * j : c.ecx.strep + 4 * (j - N);
      const double* tgt0 = agent ? samp0 + (size_t)j * stride : c.ecx.strep + 4 * (j - N) + 2;
      double ab[2], ac[2];
      double prev = nb_wedge(pts, tgt0, b, ab, ac);
      for (int s = 1; s <= S; s++)
      {
        const double* tgt = (agent && !past) ? tgt0 + 2 * s : tgt0;
        double ab1[2], ac1[2];
        const double cur = nb_wedge(pts + 2 * s, tgt, b, ab1, ac1);
        int cnt = 0, e0 = 0, e1 = 0;
        bool base_add = false;
        if (agent && !past)
        {
          const int fcd = fcode[(size_t)(s - 1) * N + j];
          if (fcd)
          {
            base_add = true;
            if (fcd == 2) e0 = 1, cnt = 1;
            if (fcd == 3) e0 = 0, cnt = 1;
          }
        }
        if (nb_neg_product(prev, cur))
        {
          const double a = nb_cross_ratio(ab, ac);
          int cs = -1;
          if (agent)
            cs = a < 0 ? 2 : (a < 1 ? 1 : (a >= 1 ? 0 : -1));
          else if (!(a < 0))
            cs = a < 1 ? 1 : 0;
          if (cs >= 0)
          {
            if (cnt == 0)
              e0 = cs;
            else
              e1 = cs;
            cnt++;
          }
        }
        if (base_add && cnt == 2 && e0 == e1) cnt = 0;  // the cancellation rule (:1221-1227)
        code |= (unsigned long long)(cnt | (e0 << 2) | (e1 << 4)) << (6 * (s - 1));
        prev = cur, ab[0] = ab1[0], ab[1] = ab1[1], ac[0] = ac1[0], ac[1] = ac1[1];
      }
    }
    if (!g.any(code != 0)) continue;
    for (int s = 1; s <= S; s++)
    {
      const int f6 = (int)((code >> (6 * (s - 1))) & 63);
      const int cnt = f6 & 3;
      if (!g.any(cnt > 0)) continue;
      int total;
      const int off = nb_excl_scan<NL>(g, cnt, total);
      const int nl = nlist[s - 1];
      const bool over = nl < 0 || nl + total > p.tcap;
      if (!over)
      {
        int* out = toadd + (size_t)(s - 1) * 2 * p.tcap + 2 * (nl + off);
        if (cnt > 0) out[0] = j + 1, out[1] = (f6 >> 2) & 3;
        if (cnt > 1) out[2] = j + 1, out[3] = (f6 >> 4) & 3;
      }
      g.sync();
      if (g.lane == 0) nlist[s - 1] = over ? -1 : nl + total;
      g.sync();
    }
  }
  g.sync();
  NB_CTICK(13)
  // phase 2: the list automaton, step by step (lane 0)
  if (g.lane == 0)
  {
    int r = 0;
    double arc = 0.0;
    for (int s = 1; s <= S && r == 0; s++)
    {
      const double* pk = pts + 2 * (s - 1);
      const double* pk1 = pts + 2 * s;
      arc += pts[18 + s];
      const int nadd = nlist[s - 1];
      int* ta = toadd + (size_t)(s - 1) * 2 * p.tcap;
      if (nadd < 0 || es.n_alpha + nadd > NA)
        r = 1;  // :844-848
      else if (nadd == 0 && M == 0 && es.n_bend == 0)
      {  // no crossing in this step, no static obstacle whose beta could change sign, no bend point to release:
         // addAlphaBetaToList and updateBendPts leave the state as it is
      }
      else
      {
        for (int i = 0; i < nadd; i++) act_old[2 * i] = ta[2 * i], act_old[2 * i + 1] = es.active[ta[2 * i] - 1];
        if (nb_s_add_alpha_beta(ta, nadd, &es, pk, &c.ecx))
          r = -1;
        else
        {
          for (int i = 0; i < nadd; i++)
          {
            const int a = act_old[2 * i] - 1, was = act_old[2 * i + 1];
            if (a >= N) continue;
            if (was < 2 && es.active[a] >= 2) r = 1;
            if (was >= 2 && es.active[a] > was) r = 1;
          }
          if (r == 0) nb_s_update_bend_pts(&es, pk1, &c.ecx);
        }
      }
    }
    if (r == 0 && nb_tether_length(c, es, pts + 2 * S) > p.tether) r = 1;  // :884-891
    flag[0] = r, flag[1] = es.n_alpha, flag[2] = es.n_bend;
    pts[NB_SEARCH_PTS - 1] = arc;
  }
  g.sync();
  const int r = flag[0];
  es.n_alpha = flag[1], es.n_bend = flag[2];
  *arc_length = pts[NB_SEARCH_PTS - 1];
  NB_CTICK(14)
  return r;
}

NB_HD void nb_search_build_tab(const NbSearchPar& p, const uint8_t* comb, NbSearchTab& tb)
{
  const double tau = p.T, j_max = p.j_max, j_min = -p.j_max;
  const double delta_x = (j_max - j_min) / (p.ns - 1);  // :1053
  for (int k = 0; k < p.ns; k++)
  {
    const double ji = j_min + k * delta_x;  // :1074
    tb.j6[k] = ji * tau * tau * tau / 6, tb.j2[k] = ji * tau * tau / 2, tb.jt[k] = ji * tau, tb.jc[k] = ji / 6;
  }
  for (int ch = 0; ch < p.nchild; ch++) tb.cjx[ch] = comb[ch] / p.ns, tb.cjy[ch] = comb[ch] % p.ns;
  for (int j = 1; j < p.S && j < 8; j++)
  {
    const double t = p.T * j / p.S;
    tb.tt[j][0] = t, tb.tt[j][1] = t * t, tb.tt[j][2] = t * t * t;
  }
}

// kinematics and admissibility of one jerk sample (:1071-1155 node form, :1260-1339 root form).
// Everything that depends only on the jerk sample comes from the per-launch table (same expressions, evaluated
// once); the tests are pure, so they are all evaluated and combined at the end (no branches between dependent
// FP64 chains).
NB_HD bool nb_search_primitive(const NbSearchCtx& c, const NbSearchTab& tb, const double* ist, int ch, bool root,
                               double* kin)
{
  const NbSearchPar& p = *c.p;
  const double tau = p.T, j_max = p.j_max, j_min = -p.j_max, a_max = p.a_max, a_min = -p.a_max;
  const double v_max = p.v_max, v_min = -p.v_max;
  const int jx = tb.cjx[ch], jy = tb.cjy[ch];
  const double i0 = ist[0], i1 = ist[1], i2 = ist[2], i3 = ist[3], i4 = ist[4], i5 = ist[5];
  const double e0 = i0 + i2 * tau + i4 * tau * tau / 2 + tb.j6[jx];
  const double e1 = i1 + i3 * tau + i5 * tau * tau / 2 + tb.j6[jy];
  const double e2 = i2 + i4 * tau + tb.j2[jx];
  const double e3 = i3 + i5 * tau + tb.j2[jy];
  const double e4 = i4 + tb.jt[jx];
  const double e5 = i5 + tb.jt[jy];
  double n2 = 0;
  n2 += (e0 - i0) * (e0 - i0), n2 += (e1 - i1) * (e1 - i1), n2 += (e2 - i2) * (e2 - i2);
  n2 += (e3 - i3) * (e3 - i3), n2 += (e4 - i4) * (e4 - i4), n2 += (e5 - i5) * (e5 - i5);
  bool ok = !(sqrt(n2) < 0.00001);
  ok &= !(e5 > a_max || e5 < a_min || e4 > a_max || e4 < a_min);
  const double cx0 = tb.jc[jx], cx1 = i4 / 2, cx2 = i2, cx3 = i0;
  const double cy0 = tb.jc[jy], cy1 = i5 / 2, cy2 = i3, cy3 = i1;
  const double bx = c.ecx.pb[2 * c.self], by = c.ecx.pb[2 * c.self + 1];
#pragma unroll
  for (int i = 0; i < 4; i++)
  {
    double qx = 0, qy = 0;
    qx += cx0 * p.Ainv[i], qx += cx1 * p.Ainv[4 + i], qx += cx2 * p.Ainv[8 + i], qx += cx3 * p.Ainv[12 + i];
    qy += cy0 * p.Ainv[i], qy += cy1 * p.Ainv[4 + i], qy += cy2 * p.Ainv[8 + i], qy += cy3 * p.Ainv[12 + i];
    ok &= !(qx < p.x_min || qx > p.x_max || qy < p.y_min || qy > p.y_max || nb_norm2(qx - bx, qy - by) > p.tether);
  }
  if (!root)
  {
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      double vx = 0, vy = 0;
      vx += cx0 * p.V[i], vx += cx1 * p.V[3 + i], vx += cx2 * p.V[6 + i];
      vy += cy0 * p.V[i], vy += cy1 * p.V[3 + i], vy += cy2 * p.V[6 + i];
      ok &= !(vx < v_min || vx > v_max || vy < v_min || vy > v_max);
    }
  }
  // future violation of the velocity bound under maximum braking jerk (:1136-1155)
  const double fx = e2 - 0.5 * e4 * e4 / (e4 > 0 ? j_min : j_max);
  const double fy = e3 - 0.5 * e5 * e5 / (e5 > 0 ? j_min : j_max);
  ok &= !((e4 > 0 && fx > v_max) || (e4 < 0 && fx < v_min));
  ok &= !((e5 > 0 && fy > v_max) || (e5 < 0 && fy < v_min));
  kin[0] = e0, kin[1] = e1, kin[2] = e2, kin[3] = e3, kin[4] = e4, kin[5] = e5;
  kin[6] = cx0, kin[7] = cx1, kin[8] = cx2, kin[9] = cx3;
  kin[10] = cy0, kin[11] = cy1, kin[12] = cy2, kin[13] = cy3;
  return ok;
}

// position control points of a node from its coefficients (Q = P * A_rest_pos_basis_inverse_, :1108)
NB_HD void nb_search_ctrl(const NbSearchPar& p, const double* kin, double* cps /*[4][2]*/)
{
  const double* cx = kin + 6;
  const double* cy = kin + 10;
  for (int i = 0; i < 4; i++)
  {
    double qx = 0, qy = 0;
    for (int k = 0; k < 4; k++) qx += cx[k] * p.Ainv[k * 4 + i], qy += cy[k] * p.Ainv[k * 4 + i];
    cps[2 * i] = qx, cps[2 * i + 1] = qy;
  }
}

// child c of the node `cur` (cur < 0: root), evaluated by one group; result in rec
template <int NL>
NB_HD void nb_search_child(const Group<NL>& g, const NbSearchCtx& c, NbSearchShared* sh, int cur, int ch)
{
  const double* ist = sh->par_kin;
  const int par_index = sh->par_index, par_na = sh->par_na, par_nb = sh->par_nb;
  const int *par_alpha = sh->par_alpha, *par_bend = sh->par_bend;
  const double* par_beta = sh->par_beta;
  const NbSearchPar& p = *c.p;
  NbChildRec& rec = sh->rec[ch];
  const bool root = cur < 0;
  NB_CTICK_INIT
  if (!rec.prim_ok) return;  // primitive evaluated by nb_search_primitives (one lane per child)
  const double* kin = rec.kin;

  const int index = par_index + 1;
  int* ci = c.ch_int + (size_t)ch * c.ch_stride;
  int* toadd = ci;
  int* act_old = ci + 2 * p.S * p.tcap;
  int* act = act_old + 2 * p.tcap;
  NbEntState es;
  es.alpha = act + c.NA;
  es.bend = es.alpha + 2 * p.ecap;
  es.beta = c.ch_dbl + (size_t)ch * (p.ecap + NB_SEARCH_PTS);
  es.active = act;
  es.n_alpha = par_na, es.n_bend = par_nb;
  for (int q = g.lane; q < par_na; q += NL)
  {
    es.alpha[2 * q] = par_alpha[2 * q], es.alpha[2 * q + 1] = par_alpha[2 * q + 1];
    es.beta[q] = par_beta[q];
  }
  for (int q = g.lane; q < par_nb; q += NL) es.bend[q] = par_bend[q];
  for (int q = g.lane; q < c.NA; q += NL) act[q] = c.par_act[q];
  g.sync();
  NB_CTICK(9)
  double arc = 0.0;
  if (p.enable_entangle)
  {
    const int r = c.multi_bend ? nb_search_entangles<NL>(g, c, es, kin, index, toadd, act_old, sh->ctl.flag[ch], &arc, ch, sh->tab.tt)
                               : nb_search_entangles_fast<NL>(g, c, es, kin, index, toadd, act_old, sh->ctl.flag[ch], &arc, ch, sh->tab.tt,
                                                              es.beta + p.ecap);
    NB_CTICK(10)
    if (r < 0 && g.lane == 0) sh->ctl.ovf_iter = 1;  // counts only if this node is really expanded (speculation)
    if (r != 0) return;
  }
  else
    arc = nb_norm2(kin[0] - ist[0], kin[1] - ist[1]);
  // getIz (:2006-2014): a sum of 32-bit terms (wrap-around), one term per lane
  uint32_t iz = 0;
  for (int i = g.lane; i < es.n_alpha; i += NL) iz += (uint32_t)(i + 1) * nb_power_int((uint32_t)es.alpha[2 * i], (uint32_t)es.alpha[2 * i + 1]);
#if defined(__CUDA_ARCH__)
  if (NL > 1) iz = __reduce_add_sync(0xffffffffu, iz);
#endif
  if (g.lane == 0)
  {
    rec.iz = (int)iz;
    rec.n_alpha = es.n_alpha, rec.n_bend = es.n_bend;
    rec.g = sh->par_g + arc;
    rec.h = nb_norm2(kin[0] - c.goal[0], kin[1] - c.goal[1]) + 0.3 * (double)es.n_alpha + 1.0 * (double)es.n_bend;
    rec.f = rec.g + p.bias * rec.h;
    // node-map lookup against the map as it is BEFORE this expansion (all children in parallel); the
    // sequential pass adds the siblings accepted ahead of this child
    rec.found = root ? -1 : nb_hash_find(c, rec.ix, rec.iy, rec.iz);
    if (rec.found >= 0)
    {
      const NbInt4 m = c.meta[rec.found];
      rec.f_state = m.z, rec.f_index = m.y;
    }
    rec.valid = 1;
  }
  NB_CTICK(11)
#if defined(__CUDA_ARCH__)
  if (c.prof && g.lane == 0) atomicMax(&sh->ctl.cmax, (int)(clock64() - ctick_));
#endif
}

// the sequential half of expandAndAddToQueue: children in all_combinations_ order (one thread)
NB_HD void nb_search_resolve(const NbSearchCtx& c, NbSearchShared* sh, int cur, int par_index)
{
  const NbSearchPar& p = *c.p;
  NbSearchCtl& ctl = sh->ctl;
  const bool root = cur < 0;
  ctl.first_new = ctl.n_used;
  ctl.n_acc = 0;
  for (int ch = 0; ch < p.nchild; ch++)
  {
    if (!root && ctl.n_used == p.max_nodes - 1) break;  // "run out of memory" (:1060-1064)
    if (root && ctl.n_used >= p.max_nodes) break;        // storage guard (the reference would overrun its pool)
    NbChildRec& rec = sh->rec[ch];
    if (!rec.valid) continue;
    if (!root)
    {
      int f = rec.found, f_state = rec.f_state, f_index = rec.f_index;
      if (f < 0)
        for (int k = 0; k < ctl.n_acc; k++)
        {
          const NbChildRec& sr = sh->rec[ctl.acc[k]];
          if (sr.ix == rec.ix && sr.iy == rec.iy && sr.iz == rec.iz)
          {  // a sibling accepted a moment ago holds this voxel
            f = sr.accept_id, f_state = 1, f_index = par_index + 1;
            break;
          }
        }
      if (f >= 0)
      {
        if (f_state == 1 && f_index == par_index + 1)
        {
          if (rec.f < c.gh[2 * f] && ctl.ran_trigger % 2 == 0)
          {  // :1193-1205: kinematics replaced; entangle state, index and heap position kept
            c.gh[2 * f] = rec.f, c.gh[2 * f + 1] = rec.h, c.ng[f] = rec.g;
            if (f >= ctl.first_new)
            {
              NbChildRec& sr = sh->rec[ctl.acc[f - ctl.first_new]];
              for (int k = 0; k < NB_SEARCH_KIN; k++) sr.kin[k] = rec.kin[k];
            }
            else
            {
              c.meta[f].x = cur;
              for (int k = 0; k < NB_SEARCH_KIN; k++) c.kin[(size_t)f * NB_SEARCH_KIN + k] = rec.kin[k];
            }
          }
          ctl.ran_trigger++;
        }
        continue;
      }
    }
    const int id = ctl.n_used;
    rec.accept_id = id;
    ctl.acc[ctl.n_acc++] = ch;
    c.gh[2 * id] = rec.f, c.gh[2 * id + 1] = rec.h, c.ng[id] = rec.g;
    c.heap[ctl.heap_n++] = id;
    nb_heap_push_at(c, ctl.heap_n - 1, 0, id);
    nb_hash_insert(c, rec.ix, rec.iy, rec.iz, id);
    ctl.n_used++;
  }
}

#if defined(__CUDA_ARCH__)
// The same sequential pass run by one warp: lane ch holds child ch in registers, the loop over the children is
// warp-uniform (broadcasts by shuffle, sibling voxel matches by ballot), only the open-list sift stays on lane 0,
// and the node-map inserts of the accepted children go in parallel afterwards (their keys are distinct).
__device__ __forceinline__ void nb_search_resolve_warp(const NbSearchCtx& c, NbSearchShared* sh, int cur, int par_index, int lane)
{
  const NbSearchPar& p = *c.p;
  NbSearchCtl& ctl = sh->ctl;
  const unsigned FULL = 0xffffffffu;
  const bool root = cur < 0;
  const bool has = lane < p.nchild;
  NbChildRec& mine = sh->rec[has ? lane : 0];
  const int valid = has ? mine.valid : 0;
  const int ix = mine.ix, iy = mine.iy, iz = mine.iz;
  const int found = (valid && !root) ? mine.found : -1, fst = mine.f_state, fidx = mine.f_index;
  const double fm = mine.f;
  int n_used = ctl.n_used, heap_n = ctl.heap_n, ran = ctl.ran_trigger;
  const int first_new = n_used;
  int my_id = -1;
  unsigned todo = __ballot_sync(FULL, valid != 0);  // the invalid children only cost a `continue` in the reference
  while (todo)
  {
    const int ch = __ffs(todo) - 1;
    todo &= todo - 1;
    if (!root && n_used == p.max_nodes - 1) break;  // "run out of memory" (:1060-1064)
    if (root && n_used >= p.max_nodes) break;
    int f = -1, f_state = 0, f_index = 0, sib = -1;
    if (!root)
    {
      f = __shfl_sync(FULL, found, ch), f_state = __shfl_sync(FULL, fst, ch), f_index = __shfl_sync(FULL, fidx, ch);
      if (f < 0)
      {
        const int kx = __shfl_sync(FULL, ix, ch), ky = __shfl_sync(FULL, iy, ch), kz = __shfl_sync(FULL, iz, ch);
        const unsigned m = __ballot_sync(FULL, my_id >= 0 && ix == kx && iy == ky && iz == kz);
        if (m)
        {  // a sibling accepted a moment ago holds this voxel
          sib = __ffs(m) - 1;
          f = __shfl_sync(FULL, my_id, sib), f_state = 1, f_index = par_index + 1;
        }
      }
    }
    if (f >= 0)
    {
      if (f_state == 1 && f_index == par_index + 1)
      {
        const double fc = __shfl_sync(FULL, fm, ch);
        if (fc < c.gh[2 * f] && ran % 2 == 0)
        {  // :1193-1205: kinematics replaced; entangle state, index and heap position kept
          __syncwarp();
          if (lane == 0) c.gh[2 * f] = fc, c.gh[2 * f + 1] = sh->rec[ch].h, c.ng[f] = sh->rec[ch].g;
          if (sib >= 0)
          {
            if (lane < NB_SEARCH_KIN) sh->rec[sib].kin[lane] = sh->rec[ch].kin[lane];
          }
          else
          {
            if (lane == 0) c.meta[f].x = cur;
            if (lane < NB_SEARCH_KIN) c.kin[(size_t)f * NB_SEARCH_KIN + lane] = sh->rec[ch].kin[lane];
          }
          __syncwarp();
        }
        ran++;
      }
      continue;
    }
    const int id = n_used;
    if (lane == ch) my_id = id;
    if (lane == 0)
    {
      const double fv = sh->rec[ch].f, hv = sh->rec[ch].h;
      c.gh[2 * id] = fv, c.gh[2 * id + 1] = hv, c.ng[id] = sh->rec[ch].g;
      nb_heap_push_at_keys(c, heap_n, 0, id, fv, hv);
    }
    __syncwarp();
    heap_n++, n_used++;
  }
  if (has) mine.accept_id = my_id;
  if (root)
  {  // the root form has no lookup: duplicates are pushed, the map keeps the first (:1377)
    __syncwarp();
    if (lane == 0)
      for (int ch = 0; ch < p.nchild; ch++)
        if (sh->rec[ch].accept_id >= 0) nb_hash_insert(c, sh->rec[ch].ix, sh->rec[ch].iy, sh->rec[ch].iz, sh->rec[ch].accept_id);
  }
  // (node form: the accepted children's voxels go into the node map during the payload copy, in parallel)
  if (lane == 0) ctl.n_used = n_used, ctl.heap_n = heap_n, ctl.ran_trigger = ran, ctl.first_new = first_new;
}
#endif

// the jerk primitives of the node published in sh->par_*: one LANE per child (a warp per child would run the same
// scalar FP64 chain on 32 lanes 25 times over), into the child records
template <class Cta>
NB_HD void nb_search_primitives(Cta& cta, const NbSearchCtx& c, NbSearchShared* sh, bool root)
{
  const NbSearchPar& p = *c.p;
  for (int ch = cta.tid; ch < p.nchild; ch += cta.nthreads)
  {
    NbChildRec& rec = sh->rec[ch];
    double kin[NB_SEARCH_KIN];
    rec.prim_ok = nb_search_primitive(c, sh->tab, sh->par_kin, ch, root, kin) ? 1 : 0;
#pragma unroll
    for (int k = 0; k < NB_SEARCH_KIN; k++) rec.kin[k] = kin[k];
    rec.valid = 0, rec.accept_id = -1;
    rec.ix = nb_voxel_index(kin[0], p.voxel), rec.iy = nb_voxel_index(kin[1], p.voxel);  // :1172-1173
  }
}

// the CTA-wide search of one agent.  Cta: tid, nthreads, warp, nwarps, lane, sync(), any(int)
// bytes of shared memory that hold every per-agent working set (the launcher clamps to what the SM has)
inline size_t nb_search_arena_wanted(const NbSearchPar& p)
{
  const size_t NA = (size_t)p.N + p.M, mn = (size_t)p.max_nodes;
  auto r16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
  size_t t = 0;
  t += r16(((size_t)p.nchild * nb_search_ch_stride(p) + NA) * 4) + r16(nb_search_chd_stride(p) * 8);
  t += r16(mn * 16) + r16(mn * 4);
  t += r16((size_t)p.N * 16) + r16((size_t)p.N * 4) + r16((size_t)p.N * nb_search_odd(2 * p.bp_max) * 8) + r16((size_t)p.N) + r16(NA * 4);
  t += r16((size_t)p.M * 32) + r16((size_t)p.M * 16) + r16((size_t)p.N * NB_NPOL * 4);
  t += r16((size_t)p.N * nb_search_odd(p.num_pol * (p.S + 1) * 2) * 8);
  t += r16(nb_search_fcode_bytes(p)) + r16((size_t)p.N * NB_SEARCH_HSTAGE_STRIDE * 8);
  return t;
}

template <class Cta, int NL>
NB_HD void nb_search_task(Cta& cta, const NbSearchArgs& a, int b, int wb, NbSearchShared* sh, unsigned char* arena, size_t arena_bytes)
{  // b: agent of the batch; wb: which per-agent workspace block to use (= b on the device)
  // ---- thread 0 builds the shared context: parameters, per-agent pointers, shared-memory placement
  if (cta.tid == 0)
  {
    sh->par = a.p;
    const NbSearchPar& p = sh->par;
    const int N = p.N, M = p.M, NA = N + M, S = p.S;
    NbSearchCtx& c = sh->cx;
    c.p = &sh->par, c.b = b, c.self = a.agent_id[b] - 1, c.NA = NA;
    c.prof = a.prof ? a.prof + (size_t)b * 16 : nullptr;
    for (int k = 0; k < 6; k++) c.init[k] = a.init[(size_t)b * 6 + k];
    c.goal[0] = a.goal[2 * b], c.goal[1] = a.goal[2 * b + 1];
    const int grp = a.group ? a.group[b] : b;
    c.hull_xy = a.hull_xy + (size_t)grp * N * NB_NPOL * NB_HMAX * 2;
    c.hull_cnt = a.hull_cnt + (size_t)grp * N * NB_NPOL;
    c.samp = a.samp + (size_t)grp * N * p.num_pol * (S + 1) * 2;
    c.known = a.known + (size_t)b * N;
    c.comb = a.comb + (a.comb_shared ? 0 : (size_t)b * p.nchild);
    c.st_ptr = a.st_ptr, c.st_xy = a.st_xy, c.st_longest = a.st_longest + (a.strep_per_agent ? (size_t)c.self * 2 * M : 0);
    c.ecx.N = N, c.ecx.M = M, c.ecx.self = c.self, c.ecx.cap = p.ecap, c.ecx.bp_max = p.bp_max, c.ecx.bp_stride = 2 * p.bp_max;
    c.samp_stride = p.num_pol * (S + 1) * 2, c.bsq_stride = NB_SEARCH_BSQ_STRIDE, c.hstage_stride = NB_SEARCH_HSTAGE_STRIDE;
    c.ecx.pb = a.pb, c.ecx.strep = a.strep + (a.strep_per_agent ? (size_t)c.self * 4 * M : 0), c.ecx.bp_cnt = a.bp_cnt, c.ecx.bp_xy = a.bp_xy;
    c.ecx.use_alt = nullptr, c.ecx.bp_cnt_alt = nullptr, c.ecx.bp_xy_alt = nullptr;
    c.a_na = a.es.cnt[2 * b], c.a_nb = a.es.cnt[2 * b + 1];
    c.a_alpha = a.es.alpha + (size_t)b * p.es_cap * 2, c.a_beta = a.es.beta + (size_t)b * p.es_cap;
    c.a_bend = a.es.bend + (size_t)b * p.es_cap, c.a_active = a.es.active + (size_t)b * NA;
    c.meta = a.nd_meta + (size_t)wb * p.max_nodes;
    c.kin = a.nd_kin + (size_t)wb * p.max_nodes * NB_SEARCH_KIN;
    c.alpha = a.nd_alpha + (size_t)wb * p.max_nodes * p.ecap * 2;
    c.beta = a.nd_beta + (size_t)wb * p.max_nodes * p.ecap;
    c.bend = a.nd_bend + (size_t)wb * p.max_nodes * p.ecap;
    c.hash = a.hash + (size_t)wb * p.hcap;
    c.ng = a.nd_g + (size_t)wb * p.max_nodes;
    c.ch_stride = nb_search_ch_stride(p);
    NbArena ar;
    ar.p = arena, ar.left = arena_bytes;
    // priority 1: the children's scratch (crossing lists, working entanglement state); 2: the open list and
    // its keys; 3: read-only per-agent inputs of the entanglement chain (copied in by all threads below)
    int* ci = (int*)ar.take(((size_t)p.nchild * c.ch_stride + NA) * sizeof(int));
    c.ch_int = ci ? ci : a.ch_int + (size_t)wb * (p.nchild * c.ch_stride + NA);
    double* cd = (double*)ar.take(nb_search_chd_stride(p) * sizeof(double));
    c.ch_dbl = cd ? cd : a.ch_dbl + (size_t)wb * nb_search_chd_stride(p);
    c.base_sq = c.ch_dbl + (size_t)p.nchild * (p.ecap + NB_SEARCH_PTS);
    double* gh = (double*)ar.take((size_t)p.max_nodes * 2 * sizeof(double));
    c.gh = gh ? gh : a.gh_g + (size_t)wb * p.max_nodes * 2;
    int* hp = (int*)ar.take((size_t)p.max_nodes * sizeof(int));
    c.heap = hp ? hp : a.heap_g + (size_t)wb * p.max_nodes;
    c.par_act = c.ch_int + (size_t)p.nchild * c.ch_stride;
    sh->stage_src[0] = c.ecx.pb, sh->stage_bytes[0] = (size_t)N * 16;
    sh->stage_src[1] = c.ecx.bp_cnt, sh->stage_bytes[1] = (size_t)N * 4;
    for (int k = 0; k < NB_SEARCH_NSTAGE; k++) sh->stage_rows[k] = 0;
    sh->stage_src[2] = c.ecx.bp_xy, sh->stage_bytes[2] = (size_t)N * nb_search_odd(2 * p.bp_max) * 8;
    sh->stage_rows[2] = N, sh->stage_src_stride[2] = 2 * p.bp_max, sh->stage_dst_stride[2] = nb_search_odd(2 * p.bp_max);
    sh->stage_src[3] = c.known, sh->stage_bytes[3] = (size_t)N;
    sh->stage_src[4] = c.a_active, sh->stage_bytes[4] = (size_t)NA * 4;
    sh->stage_src[5] = c.ecx.strep, sh->stage_bytes[5] = (size_t)M * 32;
    sh->stage_src[6] = c.st_longest, sh->stage_bytes[6] = (size_t)M * 16;
    sh->stage_src[7] = c.hull_cnt, sh->stage_bytes[7] = (size_t)N * NB_NPOL * 4;
    sh->stage_src[8] = c.samp, sh->stage_bytes[8] = (size_t)N * nb_search_odd(c.samp_stride) * 8;
    sh->stage_rows[8] = N, sh->stage_src_stride[8] = c.samp_stride, sh->stage_dst_stride[8] = nb_search_odd(c.samp_stride);
    for (int k = 0; k < NB_SEARCH_NSTAGE; k++)
      sh->stage_dst[k] = (sh->stage_src[k] && sh->stage_bytes[k]) ? ar.take(sh->stage_bytes[k]) : nullptr;
    uint8_t* fcd = (uint8_t*)ar.take(nb_search_fcode_bytes(p));
    c.fcode = fcd ? fcd : a.fcode_g + (size_t)wb * nb_search_fcode_bytes(p);
    c.multi_bend = 0;
    c.hull_stage = (double*)ar.take((size_t)N * NB_SEARCH_HSTAGE_STRIDE * sizeof(double));
    {  // goal hull of setUp (:215-220)
      const double r = 0.5, gx = c.goal[0], gy = c.goal[1];
      double* gq = sh->goal_hull;
      gq[0] = gx + r, gq[1] = gy + r, gq[2] = gx + r, gq[3] = gy - r, gq[4] = gx - r, gq[5] = gy + r, gq[6] = gx - r, gq[7] = gy - r;
    }
  }
  cta.sync();
  for (int k = 0; k < NB_SEARCH_NSTAGE; k++)
  {
    if (!sh->stage_dst[k]) continue;
    const size_t nb = sh->stage_bytes[k];
    if (sh->stage_rows[k] > 0)
    {
      const double* s8 = (const double*)sh->stage_src[k];
      double* d8 = (double*)sh->stage_dst[k];
      const int ss = sh->stage_src_stride[k], ds = sh->stage_dst_stride[k];
      for (size_t q = cta.tid; q < (size_t)sh->stage_rows[k] * ss; q += cta.nthreads) d8[(q / ss) * ds + q % ss] = s8[q];
    }
    else if ((nb & 3) == 0)
    {
      const int* s4 = (const int*)sh->stage_src[k];
      int* d4 = (int*)sh->stage_dst[k];
      for (size_t q = cta.tid; q < nb / 4; q += cta.nthreads) d4[q] = s4[q];
    }
    else
    {
      const unsigned char* s1 = (const unsigned char*)sh->stage_src[k];
      unsigned char* d1 = (unsigned char*)sh->stage_dst[k];
      for (size_t q = cta.tid; q < nb; q += cta.nthreads) d1[q] = s1[q];
    }
  }
  cta.sync();
  if (cta.tid == 0)
  {
    NbSearchCtx& c = sh->cx;
    void** d = sh->stage_dst;
    if (d[0]) c.ecx.pb = (const double*)d[0];
    if (d[1]) c.ecx.bp_cnt = (const int*)d[1];
    if (d[2]) c.ecx.bp_xy = (const double*)d[2], c.ecx.bp_stride = sh->stage_dst_stride[2];
    if (d[3]) c.known = (const uint8_t*)d[3];
    if (d[4]) c.a_active = (const int*)d[4];
    if (d[5]) c.ecx.strep = (const double*)d[5];
    if (d[6]) c.st_longest = (const double*)d[6];
    if (d[7]) c.hull_cnt = (const int*)d[7];
    if (d[8]) c.samp = (const double*)d[8], c.samp_stride = sh->stage_dst_stride[8];
  }
  cta.sync();
  const NbSearchPar& p = sh->par;
  const NbSearchCtx& c = sh->cx;
  const int N = p.N, M = p.M, NA = N + M;
  for (int ag = cta.tid; ag < N; ag += cta.nthreads)
  {  // squares around the bases (:1610-1612)
    const double radius = 0.7, bx = c.ecx.pb[2 * ag], by = c.ecx.pb[2 * ag + 1];
    double* sq = c.base_sq + c.bsq_stride * ag;
    sq[0] = bx + radius, sq[1] = by + radius, sq[2] = bx + radius, sq[3] = by - radius;
    sq[4] = bx - radius, sq[5] = by - radius, sq[6] = bx - radius, sq[7] = by + radius;
  }
  {  // tethers with bend points besides the base use the generic chain; otherwise build the base-crossing codes
    int mb = 0;
    for (int j = cta.tid; j < N; j += cta.nthreads)
      if (j != c.self && c.known[j] && c.ecx.bp_cnt[j] != 1) mb = 1;
    mb = cta.any(mb);
    if (cta.tid == 0) sh->cx.multi_bend = mb;
    if (!mb && p.enable_entangle)
    {
      const int S = p.S;
      const double* pb_self = c.ecx.pb + 2 * c.self;
      for (int q = cta.tid; q < p.num_pol * S * N; q += cta.nthreads)
      {
        const int i = q / (S * N), s0 = (q / N) % S, j = q % N;
        uint8_t code = 0;
        if (j != c.self && c.known[j])
        {
          const double* sp = c.samp + (size_t)j * c.samp_stride + ((size_t)i * (S + 1) + s0) * 2;
          code = nb_search_fcode_one(pb_self, sp, sp + 2, c.ecx.bp_xy + (size_t)c.ecx.bp_stride * j);
        }
        c.fcode[((size_t)i * 8 + s0) * N + j] = code;
      }
    }
    cta.sync();
  }
  NB_TICK_INIT
  NbSearchCtl& ctl = sh->ctl;
  Group<NL> g(cta.lane);

  // ---- setUp: clear the node map; is the goal occupied by another agent's last hull? (:213-229)
  for (int q = cta.tid; q < p.hcap; q += cta.nthreads)
  {
    NbInt4 z;
    z.x = z.y = z.z = z.w = 0;
    c.hash[q] = z;
  }
  int occ = 0;
  for (int o = cta.tid; o < N; o += cta.nthreads)
  {
    const int hn = nb_hull_count(c, o, p.num_pol - 1);
    if (hn > 0 && nb_s_gjk(c.hull_xy + ((size_t)(o * NB_NPOL + p.num_pol - 1) * NB_HMAX) * 2, hn, sh->goal_hull, 4)) occ = 1;
  }
  occ = cta.any(occ);
  if (cta.tid == 0)
  {
    ctl.done = 0, ctl.status = 2, ctl.cur = -1, ctl.n_path = 0, ctl.closest = -1, ctl.n_used = 0, ctl.heap_n = 0;
    ctl.pops = 0, ctl.ran_trigger = 0, ctl.goal_occupied = occ, ctl.first_new = 0, ctl.overflow = 0;
    ctl.smallest = DBL_MAX;
    nb_search_build_tab(p, c.comb, sh->tab);
  }
  for (int q = cta.tid; q < NA; q += cta.nthreads) c.par_act[q] = c.a_active[q];
  cta.sync();
  if (c.a_na > p.ecap || c.a_nb > p.ecap)
  {  // entangle_state_A does not fit a search node
    if (cta.tid == 0)
    {
      *a.err = 5;
      a.status[b] = 2, a.solved[b] = 0, a.n_int[b] = 0;
      a.stats[4 * b] = a.stats[4 * b + 1] = a.stats[4 * b + 2] = 0, a.stats[4 * b + 3] = occ;
      a.cost[b] = 0.0;
    }
    return;
  }

  int cur = -1;
  if (cta.tid == 0)
  {
    for (int k = 0; k < NB_SEARCH_KIN; k++) sh->par_kin[k] = k < 6 ? c.init[k] : 0.0;
    sh->par_index = 0, sh->par_g = 0.0, sh->par_na = c.a_na, sh->par_nb = c.a_nb;
    sh->par_alpha = c.a_alpha, sh->par_beta = c.a_beta, sh->par_bend = c.a_bend;
    ctl.hit[0] = ctl.hit[1] = 0, ctl.invalid = 0, ctl.cmax = 0, ctl.ovf_iter = 0;
  }
  cta.sync();
  nb_search_primitives(cta, c, sh, true);
  cta.sync();
  NB_TICK(6)
  int par = 0;  // iteration parity: the collision flag of one iteration is cleared during the next
  // Every iteration handles one popped node `cur` (the first one: the root).  The reference tests the node
  // (:1669-1737) and then expands it (:1739); neither depends on the other, so here the child warps evaluate the
  // 25 children SPECULATIVELY while the remaining warps run the collision tests of `cur`; the sequential pass that
  // creates nodes runs only if `cur` survived, so the open list, the node numbering and the result are unchanged.
  for (;;)
  {
    // one WARP per child for the entanglement chain (the primitives were evaluated when the node was published)
    if (cta.child)
      for (int ch = cta.warp; ch < p.nchild; ch += cta.nwarps) nb_search_child<NL>(g, c, sh, cur, ch);
    if (cta.aux && cur >= 0)
    {  // collidesWithObstacles2dSolve (:1514-1580) and collidesWithBases2d (:1583-1627) of the popped node
#if defined(__CUDA_ARCH__)
      const long long aux_t0 = clock64();
#endif
      const double* cps = sh->par_cps;
      const int node_index = sh->par_index;
      const int hi = node_index > p.num_pol ? p.num_pol : node_index;
      const double safe_dist = p.T * p.v_max * 2;
      int hit = 0;
      if (c.hull_stage)
      {  // the hulls of this window, copied once (coalesced) so that the GJK loops read shared memory
        for (int q = cta.aux_tid; q < N * NB_HMAX * 2; q += cta.aux_n)
        {
          const int o = q / (NB_HMAX * 2), r = q % (NB_HMAX * 2);
          if (r < 2 * nb_hull_count(c, o, hi - 1))
            c.hull_stage[(size_t)o * c.hstage_stride + r] = c.hull_xy[((size_t)(o * NB_NPOL + hi - 1) * NB_HMAX) * 2 + r];
        }
        cta.sync_aux();
      }
#if defined(__CUDA_ARCH__)
      if (c.prof && cta.aux_tid == 0) c.prof[9] += clock64() - aux_t0;   // staging part of the auxiliary path
#endif
      for (int it = cta.aux_tid; it < 2 * N + M && !hit; it += cta.aux_n)
      {
        if (it < N)
        {  // other agents' hulls of window index-1
          const int hn = nb_hull_count(c, it, hi - 1);
          const double* hv = c.hull_stage ? c.hull_stage + (size_t)it * c.hstage_stride
                                          : c.hull_xy + ((size_t)(it * NB_NPOL + hi - 1) * NB_HMAX) * 2;
          if (hn > 0 && nb_s_gjk(hv, hn, cps, 4)) hit = 1;
        }
        else if (it < N + M)
        {  // static obstacles
          const int64_t p0 = c.st_ptr[it - N], p1 = c.st_ptr[it - N + 1];
          if (nb_s_gjk(c.st_xy + 2 * p0, (int)(p1 - p0), cps, 4)) hit = 1;
        }
        else if (p.enable_entangle)
        {  // bases of the other agents
          const int ag = it - N - M;
          if (ag == c.self) continue;
          const double bx = c.ecx.pb[2 * ag], by = c.ecx.pb[2 * ag + 1];
          if (nb_norm2(cps[0] - bx, cps[1] - by) > safe_dist) continue;
          if (nb_s_gjk(c.base_sq + c.bsq_stride * ag, 4, cps, 4)) hit = 1;
        }
      }
      if (hit) ctl.hit[par] = 1;
#if defined(__CUDA_ARCH__)
      if (c.prof && cta.aux_tid == 0) c.prof[8] += clock64() - aux_t0;
#endif
    }
    cta.sync();
    NB_TICK(0)
#if defined(__CUDA_ARCH__)
    if (c.prof && cta.tid == 0) c.prof[15] += ctl.cmax, ctl.cmax = 0;
#endif
    bool expand = true;
    if (cur >= 0)
    {
      if (ctl.hit[par])
        expand = false;  // constraintViolated: continue (:1673, :1676)
      else
      {
        if (cta.tid == 0)
        {  // closest safe node so far and the goal test (:1694-1737)
          const double* nk = sh->par_kin;
          const int invalid = ctl.invalid;
          const double dist = nb_norm2(nk[0] - c.goal[0], nk[1] - c.goal[1]);
          const double dist_init = nb_norm2(nk[0] - c.init[0], nk[1] - c.init[1]);
          const double dcmp = ctl.goal_occupied ? dist * dist : dist_init;
          const double dti = dcmp * (double)sh->par_index;
          if (dti < ctl.smallest && !invalid)
          {
            ctl.smallest = dti;
            ctl.closest = cur;
          }
          if (dist < p.goal_size && !invalid) ctl.done = 1, ctl.status = 1;
        }
        cta.sync();
        if (ctl.done) break;
      }
    }
    NB_TICK(5)
    if (expand)
    {
      if (cta.tid == 0 && ctl.ovf_iter) ctl.overflow = 1;
      // ---- the sequential half of expandAndAddToQueue(cur), then the accepted children's payload
      const int par_index = sh->par_index;
#if defined(__CUDA_ARCH__)
      if (cta.warp == 0) nb_search_resolve_warp(c, sh, cur, par_index, cta.lane);
#else
      if (cta.tid == 0) nb_search_resolve(c, sh, cur, par_index);
#endif
      NB_TICK(1)
      cta.sync();
      for (int ch = cta.warp; ch < p.nchild; ch += cta.nwarps)
      {
        const NbChildRec& rec = sh->rec[ch];
        const int id = rec.accept_id;
        if (id < 0) continue;
        const int* ci = c.ch_int + (size_t)ch * c.ch_stride + (2 * p.S + 2) * p.tcap + NA;
        const double* cb = c.ch_dbl + (size_t)ch * (p.ecap + NB_SEARCH_PTS);
        for (int q = g.lane; q < rec.n_alpha; q += NL)
        {
          c.alpha[((size_t)id * p.ecap + q) * 2] = ci[2 * q], c.alpha[((size_t)id * p.ecap + q) * 2 + 1] = ci[2 * q + 1];
          c.beta[(size_t)id * p.ecap + q] = cb[q];
        }
        for (int q = g.lane; q < rec.n_bend; q += NL) c.bend[(size_t)id * p.ecap + q] = ci[2 * p.ecap + q];
        for (int q = g.lane; q < NB_SEARCH_KIN; q += NL) c.kin[(size_t)id * NB_SEARCH_KIN + q] = rec.kin[q];
        if (g.lane == 0)
        {
          NbInt4 m;
          m.x = cur, m.y = par_index + 1, m.z = 1, m.w = rec.n_alpha | (rec.n_bend << 16);
          c.meta[id] = m;
#if defined(__CUDA_ARCH__)
          if (cur >= 0)
          {  // expanded_nodes_.insert (:1219): the accepted siblings have distinct voxels, so they insert concurrently
            const uint32_t mask = (uint32_t)(p.hcap - 1);
            uint32_t q = nb_hash3(rec.ix, rec.iy, rec.iz) & mask;
            for (;;)
            {
              if (atomicCAS(&c.hash[q].w, 0, id + 1) == 0)
              {
                c.hash[q].x = rec.ix, c.hash[q].y = rec.iy, c.hash[q].z = rec.iz;
                break;
              }
              q = (q + 1) & mask;
            }
          }
#endif
        }
      }
      cta.sync();
      NB_TICK(2)
    }

    // ---- next node of the open list (:1642-1656), published for every thread
    par ^= 1;
    if (cta.tid == 0)
    {
      ctl.hit[par] = 0, ctl.ovf_iter = 0;
      if (ctl.heap_n == 0)
        ctl.done = 1, ctl.status = 2;
      else if (ctl.pops >= p.max_exp)
        ctl.done = 1, ctl.status = 0;
      else
      {
        ctl.pops++;
        const int nc = nb_heap_pop(c, ctl.heap_n);
        ctl.cur = nc;
        const NbInt4 m = c.meta[nc];
        c.meta[nc].z = -1;
        for (int k = 0; k < NB_SEARCH_KIN; k++) sh->par_kin[k] = c.kin[(size_t)nc * NB_SEARCH_KIN + k];
        nb_search_ctrl(p, sh->par_kin, sh->par_cps);
        sh->par_index = m.y, sh->par_na = m.w & 0xffff, sh->par_nb = m.w >> 16, sh->par_g = c.ng[nc];
        sh->par_alpha = c.alpha + (size_t)nc * p.ecap * 2, sh->par_beta = c.beta + (size_t)nc * p.ecap;
        sh->par_bend = c.bend + (size_t)nc * p.ecap;
      }
    }
    NB_TICK(3)
    cta.sync();
    if (ctl.done) break;
    cur = ctl.cur;
    {  // active_cases of the node: active_A - count_A + count_node; valid_endpoint (:1694-1702)
      const int na = sh->par_na;
      const int* al = sh->par_alpha;
      int invalid = 0;
      // the first warp evaluates the node's 25 jerk primitives meanwhile (they need neither list nor active cases)
      const int t_first = cta.nthreads > 32 ? 32 : 0;
      if (cta.tid < 32) nb_search_primitives(cta, c, sh, false);
      for (int q = cta.tid - t_first; q >= 0 && q < NA; q += cta.nthreads - t_first)
      {
        int v = c.a_active[q];
        for (int k = 0; k < c.a_na; k++) v -= (c.a_alpha[2 * k] == q + 1);
        for (int k = 0; k < na; k++) v += (al[2 * k] == q + 1);
        c.par_act[q] = v;
        if (q < N && v > 1) invalid = 1;
      }
      invalid = cta.any(invalid);
      if (cta.tid == 0) ctl.invalid = invalid;
    }
    NB_TICK(4)
  }

  // ---- choose the result (:1753-1826), recoverPwpOut (:521-553), recoverEntStateVector (:582-603)
  int best = -1;
  if (ctl.status == 1)
    best = ctl.cur;
  else if (ctl.closest >= 0 && p.use_not_reaching)
    best = ctl.closest;
  int* path = sh->path;
  if (cta.tid == 0)
  {
    int nn = 0;
    if (best >= 0)
      for (int t = best; t >= 0; t = c.meta[t].x)
      {
        const int idx = c.meta[t].y;
        if (idx <= p.num_pol)
        {
          path[idx - 1] = t;
          if (idx > nn) nn = idx;
        }
      }
    ctl.n_path = nn;
  }
  cta.sync();
  const int n = ctl.n_path;
  if (cta.tid == 0)
  {
    a.status[b] = ctl.status;
    a.solved[b] = best >= 0;
    a.n_int[b] = n;
    a.stats[4 * b] = ctl.n_used, a.stats[4 * b + 1] = ctl.pops, a.stats[4 * b + 2] = best >= 0 ? c.meta[best].y : 0;
    a.stats[4 * b + 3] = ctl.goal_occupied;
    a.cost[b] = best >= 0 ? c.ng[best] : 0.0;
    if (ctl.overflow) *a.err = 5;
  }
  double* co = a.coeff + (size_t)b * 3 * NB_NPOL * 4;
  for (int q = cta.tid; q < 3 * NB_NPOL * 4; q += cta.nthreads)
  {
    const int ax = q / (NB_NPOL * 4), i = (q / 4) % NB_NPOL, k = q % 4;
    double v = 0.0;
    if (i < n) v = ax == 2 ? a.coeffs_z[((size_t)b * NB_NPOL + i) * 4 + k] : c.kin[(size_t)path[i] * NB_SEARCH_KIN + 6 + 4 * ax + k];
    co[q] = v;
  }
  if (best < 0) return;
  const int ocap = p.out_cap;
  for (int i = 0; i <= NB_NPOL; i++)
  {
    const int src = i == 0 ? -1 : path[(i <= n ? i : n) - 1];
    const int na = src < 0 ? c.a_na : (c.meta[src].w & 0xffff), nbd = src < 0 ? c.a_nb : (c.meta[src].w >> 16);
    const int* al = src < 0 ? c.a_alpha : c.alpha + (size_t)src * p.ecap * 2;
    const double* be = src < 0 ? c.a_beta : c.beta + (size_t)src * p.ecap;
    const int* bd = src < 0 ? c.a_bend : c.bend + (size_t)src * p.ecap;
    const size_t o = (size_t)b * 9 + i;
    if (cta.tid == 0)
    {
      a.esv.cnt[2 * o] = na, a.esv.cnt[2 * o + 1] = nbd;
      if (na > ocap || nbd > ocap) *a.err = 5;
    }
    for (int q = cta.tid; q < ocap; q += cta.nthreads)
    {
      const bool in = q < na && na <= ocap;
      a.esv.alpha[(o * ocap + q) * 2] = in ? al[2 * q] : 0;
      a.esv.alpha[(o * ocap + q) * 2 + 1] = in ? al[2 * q + 1] : 0;
      a.esv.beta[o * ocap + q] = in ? be[q] : 0.0;
      a.esv.bend[o * ocap + q] = (q < nbd && nbd <= ocap) ? bd[q] : 0;
    }
    for (int q = cta.tid; q < NA; q += cta.nthreads)
    {
      int v = c.a_active[q];
      if (src >= 0)
      {
        for (int k = 0; k < c.a_na; k++) v -= (c.a_alpha[2 * k] == q + 1);
        for (int k = 0; k < na; k++) v += (al[2 * k] == q + 1);
      }
      a.esv.active[o * NA + q] = v;
    }
  }
}
