// nb_entangle.cuh -- K3: tether entanglement-signature chain, one agent per warp-group.
//
// Replaces (reference neptune/src/entangle_utils.cpp) eu::vectorWedge2 :16-27,
// eu::entangleHSigToAddAgentInd (8-arg) :1129-1228, eu::entangleHSigToAddStatic :1231-1277,
// eu::addAlphaBetaToList :1402-1534, eu::updateBendPts :1536-1604, eu::breakcondition :1608-1647,
// eu::getBendPt2d :1649-1679, eu::calculateBetaForCase :1709-1722, and their drivers
// Neptune::PredictAlphasBetas (neptune.cpp:976-1008), KinodynamicSearch::entangleCheckGivenPwp
// (kinodynamic_search.cpp:897-985) and the per-interval chain of entanglesWithOtherAgents (:707-895).
//
// The crossing tests against the N other tethers and M static obstacles are independent: lanes take
// one tether each, keep their (id, case) entries locally and a warp prefix sum places them in the
// shared list in agent order, exactly the order the reference's sequential loop produces.  The
// variable-length signature word is then reduced by lane 0 (it is a short, inherently sequential
// integer automaton).  All outputs are integers decided by IEEE comparisons of FP64 wedges.
#pragma once
#include "../../include/neptune_b200.h"
#include "nb_common.cuh"
#include "nb_hull.cuh"

#define NB_ENT_LOCAL 24  // entries one tether can add in one step: bp_max + 2 <= 24 pairs... (ints = 2x)

struct NbEntState
{
  int n_alpha, n_bend;
  int* alpha;    // [cap][2]
  double* beta;  // [cap]
  int* bend;     // [cap]
  int* active;   // [N+M]
};

struct NbEntCtx
{
  int N, M, self, cap, bp_max;
  int bp_stride;        // doubles between the bend-point lists of consecutive agents (2 bp_max, or padded in shared memory)
  const double* pb;     // [N][2]
  const double* strep;  // [M][2][2]
  const int* bp_cnt;    // [N]
  const double* bp_xy;  // [N][bp_max][2]
  // post-check only (Neptune::safetyCheckAfterReplan): agents whose trajectory arrived late are seen with the bend
  // points of that late message (trajCB, neptune_ros.cpp:418-424); nullptr elsewhere
  const unsigned char* use_alt;  // [N]
  const int* bp_cnt_alt;
  const double* bp_xy_alt;
};

NB_HD int nb_bp_cnt(const NbEntCtx& cx, int j) { return (cx.use_alt && cx.use_alt[j]) ? cx.bp_cnt_alt[j] : cx.bp_cnt[j]; }
NB_HD const double* nb_bp_xy(const NbEntCtx& cx, int j)
{
  return ((cx.use_alt && cx.use_alt[j]) ? cx.bp_xy_alt : cx.bp_xy) + (size_t)cx.bp_stride * j;
}

// eu::vectorWedge2 (:16-27): (b-a) x (c-a); ab, ac returned when asked for
NB_HD double nb_wedge(const double* a, const double* b, const double* c, double* ab, double* ac)
{
  const double abx = b[0] - a[0], aby = b[1] - a[1], acx = c[0] - a[0], acy = c[1] - a[1];
  if (ab)
  {
    ab[0] = abx, ab[1] = aby, ac[0] = acx, ac[1] = acy;
  }
#if defined(__CUDA_ARCH__)
  return __dsub_rn(__dmul_rn(abx, acy), __dmul_rn(acx, aby));  // no FMA contraction: sign must match the CPU
#else
  return abx * acy - acx * aby;
#endif
}

// classification ratio on the dominant coordinate (:1164-1172, :1194-1202)
NB_HD double nb_cross_ratio(const double* ab, const double* ac)
{
#if defined(__CUDA_ARCH__)
  if (fabs(__dmul_rn(ab[1], ac[1])) > fabs(__dmul_rn(ab[0], ac[0]))) return __ddiv_rn(ab[1], ac[1]);
  return __ddiv_rn(ab[0], ac[0]);
#else
  if (fabs(ab[1] * ac[1]) > fabs(ab[0] * ac[0])) return ab[1] / ac[1];
  return ab[0] / ac[0];
#endif
}

NB_HD bool nb_neg_product(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b) < 0;
#else
  return a * b < 0;
#endif
}

// eu::entangleHSigToAddAgentInd, 8-arg (:1129-1228): entries of ONE tether into loc; returns count
NB_HD int nb_hsig_agent(int* loc, const double* pk, const double* pk1, const double* pik, const double* pik1,
                        const double* pb_self, const double* bend, int nbend, int agent_id)
{
  int nadd = 0;
  bool base_add = false;
  for (int i = 0; i < nbend; i++)
  {
    double ab[2], ac[2], c1, c2;
    const double* bi = bend + 2 * i;
    const bool last = (i == nbend - 1);
    if (!last)
    {
      c1 = nb_wedge(pk, bend + 2 * (i + 1), bi, ab, ac);
      c2 = nb_wedge(pk1, bend + 2 * (i + 1), bi, nullptr, nullptr);
    }
    else
    {
      c1 = nb_wedge(pk, pik, bi, ab, ac);
      c2 = nb_wedge(pk1, pik1, bi, nullptr, nullptr);
    }
    if (last)
    {
      double fb[2], fc[2];
      const double f1 = nb_wedge(pb_self, pik, bi, fb, fc);
      const double f2 = nb_wedge(pb_self, pik1, bi, nullptr, nullptr);
      if (nb_neg_product(f1, f2))
      {
        const double a = nb_cross_ratio(fb, fc);
        if (a < 0)
        {
        }
        else if (a < 1)
        {
          loc[2 * nadd] = agent_id, loc[2 * nadd + 1] = 1;
          nadd++;
        }
        else if (i == 0)
        {
          loc[2 * nadd] = agent_id, loc[2 * nadd + 1] = 0;
          nadd++;
        }
        base_add = true;
      }
    }
    if (nb_neg_product(c1, c2))
    {
      const double a = nb_cross_ratio(ab, ac);
      if (a < 0)
      {
        loc[2 * nadd] = agent_id, loc[2 * nadd + 1] = i + 2;
        nadd++;
      }
      else if (a < 1 && last)
      {
        loc[2 * nadd] = agent_id, loc[2 * nadd + 1] = 1;
        nadd++;
      }
      else if (a >= 1 && i == 0)
      {
        loc[2 * nadd] = agent_id, loc[2 * nadd + 1] = 0;
        nadd++;
      }
    }
  }
  if (base_add && nadd >= 2 && loc[2 * (nadd - 1)] == loc[2 * (nadd - 2)] && loc[2 * (nadd - 1) + 1] == loc[2 * (nadd - 2) + 1])
    nadd -= 2;
  return nadd;
}

// Same function without a per-thread entry buffer: the entries go straight to `out` (pairs) when out != nullptr,
// otherwise they are only counted; the last two entries stay in registers for the cancellation rule (:1221-1227).
// nb_collect_toadd calls it twice (count, then write at the scanned offset) so no thread keeps a stack array.
// limit: entries the caller reserved room for (the final count of the counting pass): the cancellation rule drops
// the LAST two entries, so on the writing pass everything from index `limit` on must not be stored.
NB_HD int nb_hsig_agent_stream(int* out, const double* pk, const double* pk1, const double* pik, const double* pik1,
                               const double* pb_self, const double* bend, int nbend, int agent_id, int limit)
{
  int nadd = 0, last_cs = -1, prev_cs = -2;  // entries all carry agent_id: only the case numbers can differ
  bool base_add = false;
#define NB_HSIG_PUSH(cs_)                                                   \
  {                                                                         \
    if (out && nadd < limit) out[2 * nadd] = agent_id, out[2 * nadd + 1] = (cs_); \
    prev_cs = last_cs, last_cs = (cs_);                                     \
    nadd++;                                                                 \
  }
  for (int i = 0; i < nbend; i++)
  {
    double ab[2], ac[2], c1, c2;
    const double* bi = bend + 2 * i;
    const bool last = (i == nbend - 1);
    if (!last)
    {
      c1 = nb_wedge(pk, bend + 2 * (i + 1), bi, ab, ac);
      c2 = nb_wedge(pk1, bend + 2 * (i + 1), bi, nullptr, nullptr);
    }
    else
    {
      c1 = nb_wedge(pk, pik, bi, ab, ac);
      c2 = nb_wedge(pk1, pik1, bi, nullptr, nullptr);
    }
    if (last)
    {
      double fb[2], fc[2];
      const double f1 = nb_wedge(pb_self, pik, bi, fb, fc);
      const double f2 = nb_wedge(pb_self, pik1, bi, nullptr, nullptr);
      if (nb_neg_product(f1, f2))
      {
        const double a = nb_cross_ratio(fb, fc);
        if (a < 0)
        {
        }
        else if (a < 1)
          NB_HSIG_PUSH(1)
        else if (i == 0)
          NB_HSIG_PUSH(0)
        base_add = true;
      }
    }
    if (nb_neg_product(c1, c2))
    {
      const double a = nb_cross_ratio(ab, ac);
      if (a < 0)
        NB_HSIG_PUSH(i + 2)
      else if (a < 1 && last)
        NB_HSIG_PUSH(1)
      else if (a >= 1 && i == 0)
        NB_HSIG_PUSH(0)
    }
  }
#undef NB_HSIG_PUSH
  if (base_add && nadd >= 2 && last_cs == prev_cs) nadd -= 2;
  return nadd;
}

// eu::entangleHSigToAddAgentInd, 9-arg (:820-1127): the online tracker's form, aware of a bend point of the other
// tether having been added or released since the last check.  Streams its entries like nb_hsig_agent_stream.
// *stop = k where the reference prints "stop k" and calls exit(-1) (:917, :981, :1073, :1107).
NB_HD int nb_hsig_agent9_stream(int* out, const double* pk, const double* pk1, const double* pik, const double* pik1,
                                const double* pb_self, const double* bend, int nbend, const double* prev, int nprev,
                                int agent_id, int* stop, int limit)
{
  if (nprev == nbend) return nb_hsig_agent_stream(out, pk, pk1, pik, pik1, pb_self, bend, nbend, agent_id, limit);
  if (nbend == 0 || nprev == 0) return 0;  // "bendpts empty!" (:833-836)
  int nadd = 0, last_cs = -1, prev_cs = -2;
  bool base_add = false;
#define NB_HSIG_PUSH(cs_)                                                         \
  {                                                                               \
    if (out && nadd < limit) out[2 * nadd] = agent_id, out[2 * nadd + 1] = (cs_); \
    prev_cs = last_cs, last_cs = (cs_);                                           \
    nadd++;                                                                       \
  }
  const double* pback = prev + 2 * (nprev - 1);
  if (nbend < nprev)
  {  // released from a bend point (:837-986)
    const double* pback2 = prev + 2 * (nprev - 2);
    for (int i = 0; i < nbend; i++)
    {
      double ab[2], ac[2], abp[2] = { 0, 0 }, acp[2] = { 0, 0 }, c1, c2, c1p = 0.0;
      const double* bi = bend + 2 * i;
      const bool last = (i == nbend - 1);
      if (!last)
      {
        c1 = nb_wedge(pk, bend + 2 * (i + 1), bi, ab, ac);
        c2 = nb_wedge(pk1, bend + 2 * (i + 1), bi, nullptr, nullptr);
      }
      else
      {
        c1 = nb_wedge(pk, pik, pback, ab, ac);
        c2 = nb_wedge(pk1, pik1, bi, nullptr, nullptr);
        c1p = nb_wedge(pk, pback, pback2, abp, acp);
      }
      if (last)
      {
        double fb[2], fc[2];
        const double f1 = nb_wedge(pb_self, pik, pback, nullptr, nullptr);
        const double f2 = nb_wedge(pb_self, pik1, bi, fb, fc);
        double f1p = 0.0;
        if (i == 0) f1p = nb_wedge(pb_self, pback, pback2, nullptr, nullptr);
        if (nb_neg_product(f1, f2))
        {
          const double a = nb_cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
            NB_HSIG_PUSH(1)
          base_add = true;
        }
        if (i == 0 && nb_neg_product(f1p, f2))
        {
          const double a = nb_cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
          {
          }
          else
          {
            NB_HSIG_PUSH(0)
            *stop = 4;
          }
          base_add = true;
        }
      }
      bool added_inbtw = false;
      if (nb_neg_product(c1, c2))
      {
        const double a = nb_cross_ratio(ab, ac);
        if (a < 0)
        {
          NB_HSIG_PUSH(i + 2)
          added_inbtw = true;
        }
        else if (a < 1 && last)
          NB_HSIG_PUSH(1)
      }
      if (last && nb_neg_product(c1p, c2))
      {
        const double a = nb_cross_ratio(abp, acp);
        if (a < 0 && !added_inbtw)
          NB_HSIG_PUSH(i + 2)
        else if (a < 1 && last)
        {
        }
        else if (a >= 1 && i == 0)
        {
          NB_HSIG_PUSH(0)
          *stop = 3;
        }
      }
    }
  }
  else
  {  // a bend point was added (:987-1116)
    for (int i = 0; i < nbend; i++)
    {
      double ab[2], ac[2], c1, c2;
      const double* bi = bend + 2 * i;
      if (i == nbend - 1)
      {
        c1 = nb_wedge(pk, pik, pback, nullptr, nullptr);
        c2 = nb_wedge(pk1, pik1, bi, ab, ac);
      }
      else if (i == nbend - 2)
      {
        c1 = nb_wedge(pk, pik, pback, nullptr, nullptr);
        c2 = nb_wedge(pk1, bend + 2 * (i + 1), bi, ab, ac);
      }
      else
      {
        c1 = nb_wedge(pk, bend + 2 * (i + 1), bi, ab, ac);
        c2 = nb_wedge(pk1, bend + 2 * (i + 1), bi, nullptr, nullptr);
      }
      if (i == nbend - 1)
      {
        double fb[2], fc[2];
        const double f1 = nb_wedge(pb_self, pik, pback, nullptr, nullptr);
        const double f2 = nb_wedge(pb_self, pik1, bi, fb, fc);
        if (nb_neg_product(f1, f2))
        {
          const double a = nb_cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
            NB_HSIG_PUSH(1)
          base_add = true;
        }
      }
      if (i == 0 && nbend == 2)
      {
        double fb[2], fc[2];
        const double f1 = nb_wedge(pb_self, pik, pback, nullptr, nullptr);
        const double f2 = nb_wedge(pb_self, bend + 2 * (i + 1), bi, fb, fc);
        if (nb_neg_product(f1, f2))
        {
          const double a = nb_cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
          {
          }
          else
          {
            NB_HSIG_PUSH(0)
            *stop = 2;
          }
          base_add = true;
        }
      }
      if (nb_neg_product(c1, c2))
      {
        const double a = nb_cross_ratio(ab, ac);
        if (a < 0)
          NB_HSIG_PUSH(i + 2)
        else if (a < 1 && i == nbend - 1)
          NB_HSIG_PUSH(1)
        else if (a >= 1 && i == 0)
        {
          NB_HSIG_PUSH(0)
          *stop = 1;
        }
      }
    }
  }
#undef NB_HSIG_PUSH
  if (base_add && nadd >= 2 && last_cs == prev_cs) nadd -= 2;
  return nadd;
}

// eu::entangleHSigToAddStatic (:1231-1277) for ONE static obstacle
NB_HD int nb_hsig_static_one(int* loc, const double* pk, const double* pk1, const double* rep /*[2][2]*/, int id)
{
  const double* pbi = rep;      // col(0)
  const double* pik = rep + 2;  // col(1)
  double ab[2], ac[2];
  const double c1 = nb_wedge(pk, pik, pbi, ab, ac);
  const double c2 = nb_wedge(pk1, pik, pbi, nullptr, nullptr);
  if (!nb_neg_product(c1, c2)) return 0;
  const double a = nb_cross_ratio(ab, ac);
  if (a < 0) return 0;
  loc[0] = id;
  loc[1] = (a < 1) ? 1 : 0;
  return 1;
}

// warp-exclusive prefix sum of per-lane counts; total returned to every lane
template <int NL>
NB_HD int nb_excl_scan(const Group<NL>& g, int v, int& total)
{
#if defined(__CUDA_ARCH__)
  constexpr int W = NL < 32 ? NL : 32;
  int x = v;
#pragma unroll
  for (int o = 1; o < W; o <<= 1)
  {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((g.lane & 31) >= o) x += y;
  }
  if (NL <= 32)
  {
    total = __shfl_sync(0xffffffffu, x, W - 1);
    return x - v;
  }
  // several warps (a CTA per agent in worlds with hundreds of tethers): warp totals through shared memory
  __shared__ int wtot[NL > 32 ? NL / 32 : 1];
  const int w = g.lane >> 5;
  if ((g.lane & 31) == 31) wtot[w] = x;
  __syncthreads();
  int before = 0, all = 0;
#pragma unroll
  for (int q = 0; q < NL / 32; q++)
  {
    if (q < w) before += wtot[q];
    all += wtot[q];
  }
  __syncthreads();
  total = all;
  return before + x - v;
#else
  total = v;
  return 0;
#endif
}

// All crossing tests of one step pk -> pk+1: tethers of the known agents (positions pik -> pik+1 per
// agent: pik_all[j], pik1_all[j]) then the static obstacles.  toadd: shared list [tcap][2].
// Returns the number of entries, or -1 if tcap would be exceeded.  Two passes per chunk of NL tethers: count,
// then (only when the chunk has a crossing at all) a prefix sum and a second evaluation that writes in place.
template <int NL>
NB_HD int nb_collect_toadd(const Group<NL>& g, const NbEntCtx& cx, const double* pk, const double* pk_agents,
                           const double* pk1, const double* pik_all, int pik_stride, const double* pik1_all,
                           int pik1_stride, const unsigned char* known, int* toadd, int tcap)
{
  int nadd = 0;
  const double* pb_self = cx.pb + 2 * cx.self;
  for (int base = 0; base < cx.N + cx.M; base += NL)
  {
    const int j = base + g.lane;
    int cnt = 0, nb = 0, st[2];
    const bool agent = j < cx.N && j != cx.self && known[j];
    if (agent)
    {
      nb = nb_bp_cnt(cx, j);
      if (nb > NB_ENT_LOCAL - 2) nb = NB_ENT_LOCAL - 2;
      cnt = nb_hsig_agent_stream(nullptr, pk_agents ? pk_agents + 2 * j : pk, pk1, pik_all + (size_t)j * pik_stride,
                                 pik1_all + (size_t)j * pik1_stride, pb_self, nb_bp_xy(cx, j), nb, j + 1, 0);
    }
    else if (j >= cx.N && j < cx.N + cx.M)
      cnt = nb_hsig_static_one(st, pk, pk1, cx.strep + 4 * (j - cx.N), j + 1);
    if (!g.any(cnt > 0)) continue;  // no crossing in this chunk of tethers (the common case): nothing to place
    int total;
    const int off = nb_excl_scan<NL>(g, cnt, total);
    if (nadd + total > tcap) return -1;
    if (cnt > 0)
    {
      int* out = toadd + 2 * (nadd + off);
      if (agent)
        nb_hsig_agent_stream(out, pk_agents ? pk_agents + 2 * j : pk, pk1, pik_all + (size_t)j * pik_stride,
                             pik1_all + (size_t)j * pik1_stride, pb_self, nb_bp_xy(cx, j), nb, j + 1, cnt);
      else
        out[0] = st[0], out[1] = st[1];
    }
    nadd += total;
  }
  g.sync();
  return nadd;
}

// Crossing tests of one tracker tick (NeptuneRos::updateEntStateStaticObs, neptune_ros.cpp:806-822): per known
// agent its own previous checking position and the 9-argument test; then the static obstacles from
// previousCheckingPos_[N].  Also advances previousCheckingPos_[i] / previousCheckingPosAgent_[i] (:814-815).
template <int NL>
NB_HD int nb_collect_track(const Group<NL>& g, const NbEntCtx& cx, const int* bp_cnt_prev, const double* bp_xy_prev,
                           double* prev_pos, double* prev_pos_agent, const double* latest, const double* cur, int* toadd,
                           int tcap, int* stop_out)
{
  int nadd = 0;
  const double* pb_self = cx.pb + 2 * cx.self;
  const double pkN[2] = { prev_pos[2 * cx.N], prev_pos[2 * cx.N + 1] };
  for (int base = 0; base < cx.N + cx.M; base += NL)
  {
    const int j = base + g.lane;
    int cnt = 0, nb = 0, np = 0, st[2], stop = 0;
    double pk[2] = { 0, 0 }, pik[2] = { 0, 0 };
    bool agent = j < cx.N && j != cx.self;
    if (agent)
    {
      nb = cx.bp_cnt[j], np = bp_cnt_prev[j];
      pik[0] = prev_pos_agent[2 * j], pik[1] = prev_pos_agent[2 * j + 1];
      if (pik[0] < -900 || nb == 0) agent = false;  // agent msg not received yet (:811)
    }
    if (agent)
    {
      pk[0] = prev_pos[2 * j], pk[1] = prev_pos[2 * j + 1];
      if (nb > NB_ENT_LOCAL - 2) nb = NB_ENT_LOCAL - 2;
      cnt = nb_hsig_agent9_stream(nullptr, pk, cur, pik, latest + 2 * j, pb_self, cx.bp_xy + (size_t)cx.bp_stride * j, nb,
                                  bp_xy_prev + (size_t)cx.bp_stride * j, np, j + 1, &stop, 0);
    }
    else if (j >= cx.N && j < cx.N + cx.M)
      cnt = nb_hsig_static_one(st, pkN, cur, cx.strep + 4 * (j - cx.N), j + 1);
    if (g.any(stop != 0))
    {  // the sequential loop keeps the value assigned last: the highest tether index wins
#if defined(__CUDA_ARCH__)
      const unsigned m = __ballot_sync(0xffffffffu, stop != 0);  // (the tracker tick always runs one warp per agent)
      if (g.lane == 31 - __clz(m)) *stop_out = stop;
#else
      *stop_out = stop;
#endif
    }
    if (g.any(cnt > 0))
    {
      int total;
      const int off = nb_excl_scan<NL>(g, cnt, total);
      if (nadd + total > tcap) return -1;
      if (cnt > 0)
      {
        int* out = toadd + 2 * (nadd + off);
        int dummy = 0;
        if (agent)
          nb_hsig_agent9_stream(out, pk, cur, pik, latest + 2 * j, pb_self, cx.bp_xy + (size_t)cx.bp_stride * j, nb,
                                bp_xy_prev + (size_t)cx.bp_stride * j, np, j + 1, &dummy, cnt);
        else
          out[0] = st[0], out[1] = st[1];
      }
      nadd += total;
    }
    if (agent)
    {
      prev_pos[2 * j] = cur[0], prev_pos[2 * j + 1] = cur[1];
      prev_pos_agent[2 * j] = latest[2 * j], prev_pos_agent[2 * j + 1] = latest[2 * j + 1];
    }
  }
  g.sync();
  return nadd;
}

// eu::getBendPt2d (:1649-1679)
NB_HD void nb_bend_pt(double* bp, const NbEntState& es, const NbEntCtx& cx)
{
  if (es.n_bend == 0)
  {
    bp[0] = cx.pb[2 * cx.self], bp[1] = cx.pb[2 * cx.self + 1];
    return;
  }
  const int q = es.bend[es.n_bend - 1];
  const int id = es.alpha[2 * q], cs = es.alpha[2 * q + 1];
  if (id <= cx.N && id >= 1)
    bp[0] = cx.pb[2 * (id - 1)], bp[1] = cx.pb[2 * (id - 1) + 1];
  else if (id > cx.N)
    bp[0] = cx.strep[4 * (id - cx.N - 1) + 2 * cs], bp[1] = cx.strep[4 * (id - cx.N - 1) + 2 * cs + 1];
}

NB_HD void nb_bend_coord(double* bp, int id, int cs, const NbEntCtx& cx)
{
  if (id <= cx.N)
    bp[0] = cx.pb[2 * (id - 1)], bp[1] = cx.pb[2 * (id - 1) + 1];
  else
    bp[0] = cx.strep[4 * (id - cx.N - 1) + 2 * cs], bp[1] = cx.strep[4 * (id - cx.N - 1) + 2 * cs + 1];
}

// eu::calculateBetaForCase (:1709-1722)
NB_HD double nb_beta_for_case(int id, int cs, const double* pk, const double* bp, const NbEntCtx& cx)
{
  if (id <= cx.N) return 0.0;
  return nb_wedge(pk, cx.strep + 4 * (id - cx.N - 1) + 2 * cs, bp, nullptr, nullptr);
}

// eu::breakcondition (:1608-1647)
NB_HD bool nb_break_condition(int aid, int acs, int lid, int N, int idx_to_check, int idx_last_bend)
{
  if (aid <= N && acs >= 2) return idx_to_check <= idx_last_bend;
  if (aid <= N) return false;
  return lid > N || idx_to_check <= idx_last_bend;
}

// eu::addAlphaBetaToList (:1402-1534); single lane.  Returns 0, or -1 on storage overflow.
NB_HD int nb_add_alpha_beta(int* toadd, int nadd, NbEntState& es, const double* pk, const NbEntCtx& cx)
{
  const int N = cx.N;
  bool have = true;
  while (have)
  {
    have = false;
    const int b = es.n_bend == 0 ? -1 : es.bend[es.n_bend - 1];
    for (int i = 0; i < nadd && !have; i++)
    {
      const int aid = toadd[2 * i], acs = toadd[2 * i + 1];
      const int nb = (aid <= N) ? nb_bp_cnt(cx, aid - 1) : 0;
      for (int j = es.n_alpha - 1; j >= 0; j--)
      {
        const int lid = es.alpha[2 * j], lcs = es.alpha[2 * j + 1];
        const int diff = acs > lcs ? acs - lcs : lcs - acs;
        const bool cond = (lid == aid && lcs == acs) || (aid <= N && lid == aid && acs >= nb + 1 && acs < lcs) ||
                          (aid <= N && lid == aid && lcs >= 2 && acs >= 2 && diff == 1 && j > b);
        if (cond)
        {
          es.active[aid - 1] -= 1;
          for (int q = i; q < nadd - 1; q++) toadd[2 * q] = toadd[2 * (q + 1)], toadd[2 * q + 1] = toadd[2 * (q + 1) + 1];
          nadd--;
          for (int q = j; q < es.n_alpha - 1; q++)
          {
            es.alpha[2 * q] = es.alpha[2 * (q + 1)], es.alpha[2 * q + 1] = es.alpha[2 * (q + 1) + 1];
            es.beta[q] = es.beta[q + 1];
          }
          es.n_alpha--;
          if (j == b)
          {
            es.n_bend--;
            double bp[2];
            nb_bend_pt(bp, es, cx);
            for (int k = j; k < es.n_alpha; k++) es.beta[k] = nb_beta_for_case(es.alpha[2 * k], es.alpha[2 * k + 1], pk, bp, cx);
          }
          else if (j < b)
          {
            es.bend[es.n_bend - 1] = b - 1;
            for (int k = es.n_bend - 2; k >= 0; k--)
            {
              if (es.bend[k] > j)
                es.bend[k] -= 1;
              else
                break;
            }
          }
          have = true;
          break;
        }
        if (nb_break_condition(aid, acs, lid, N, j, b)) break;
      }
    }
  }
  if (nadd == 0) return 0;
  double bp[2];
  nb_bend_pt(bp, es, cx);
  for (int i = 0; i < nadd; i++)
  {
    if (es.n_alpha >= cx.cap) return -1;
    es.alpha[2 * es.n_alpha] = toadd[2 * i];
    es.alpha[2 * es.n_alpha + 1] = toadd[2 * i + 1];
    es.active[toadd[2 * i] - 1] += 1;
    es.beta[es.n_alpha] = nb_beta_for_case(toadd[2 * i], toadd[2 * i + 1], pk, bp, cx);
    es.n_alpha++;
  }
  return 0;
}

NB_HD bool nb_lt_prod(double a, double b, double thr)
{
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b) < thr;
#else
  return a * b < thr;
#endif
}

NB_HD bool nb_gt_prod(double a, double b, double thr)
{
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b) > thr;
#else
  return a * b > thr;
#endif
}

// eu::updateBendPts (:1536-1604); single lane
NB_HD void nb_update_bend_pts(NbEntState& es, const double* pk1, const NbEntCtx& cx)
{
  double bp[2];
  nb_bend_pt(bp, es, cx);
  int idx_new = -1;
  const int start = es.n_bend == 0 ? -1 : es.bend[es.n_bend - 1];
  // agents: calculateBetaForCase returns 0, and 0 * beta < -1e-7 is never true -- without static obstacles no entry
  // can become a bend point, so the scan is skipped altogether
  for (int i = start + 1; i < es.n_alpha && cx.M > 0; i++)
  {
    if (es.alpha[2 * i] <= cx.N) continue;
    const double beta = nb_beta_for_case(es.alpha[2 * i], es.alpha[2 * i + 1], pk1, bp, cx);
    if (nb_lt_prod(beta, es.beta[i], -1e-7)) idx_new = i;
  }
  if (idx_new > -1)
  {
    if (es.n_bend < cx.cap) es.bend[es.n_bend++] = idx_new;
    double nbp[2];
    nb_bend_coord(nbp, es.alpha[2 * idx_new], es.alpha[2 * idx_new + 1], cx);
    for (int i = idx_new + 1; i < es.n_alpha; i++)
      es.beta[i] = nb_beta_for_case(es.alpha[2 * i], es.alpha[2 * i + 1], pk1, nbp, cx);
    return;
  }
  while (es.n_bend > 0)
  {
    double prev[2];
    if (es.n_bend == 1)
      prev[0] = cx.pb[2 * cx.self], prev[1] = cx.pb[2 * cx.self + 1];
    else
    {
      const int q = es.bend[es.n_bend - 2];
      nb_bend_coord(prev, es.alpha[2 * q], es.alpha[2 * q + 1], cx);
    }
    const int lb = es.bend[es.n_bend - 1];
    const double beta = nb_beta_for_case(es.alpha[2 * lb], es.alpha[2 * lb + 1], pk1, prev, cx);
    if (nb_gt_prod(beta, es.beta[lb], 1e-7))
    {
      for (int k = lb + 1; k < es.n_alpha; k++)
        es.beta[k] = nb_beta_for_case(es.alpha[2 * k], es.alpha[2 * k + 1], pk1, prev, cx);
      es.n_bend--;
    }
    else
      break;
  }
}

NB_HD void nb_eval_xy(const double* cxy /*[3][8][4] of agent*/, int i, double t, double* p)
{
  const double* x = cxy + 4 * i;
  const double* y = cxy + 32 + 4 * i;
  const double t3 = t * t * t, t2 = t * t;
#if defined(__CUDA_ARCH__)
  p[0] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x[0], t3), __dmul_rn(x[1], t2)), __dmul_rn(x[2], t)), x[3]);
  p[1] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(y[0], t3), __dmul_rn(y[1], t2)), __dmul_rn(y[2], t)), y[3]);
#else
  p[0] = x[0] * t3 + x[1] * t2 + x[2] * t + x[3];
  p[1] = y[0] * t3 + y[1] * t2 + y[2] * t + y[3];
#endif
}

// One S-step pass over interval ii: shared by the front-end chain (kinodynamic_search.cpp:813-881)
// and the post-check (:909-972).  samp: [N][num_pol][S+1][2] of this planning agent.
// Returns 1 entangling, 0 fine, -1 capacity overflow.  Group-uniform result.
template <int NL>
NB_HD int nb_ent_interval_pass(const Group<NL>& g, NbEntState& es, const NbEntCtx& cx, const double* cxy, int ii,
                               const double* samp, const unsigned char* known, int num_pol, int S, double T, int limit,
                               int* toadd, int tcap, int* pairs /*[tcap][2] shared*/, int* flag /*shared int*/)
{
  double pk[2] = { cxy[4 * ii + 3], cxy[32 + 4 * ii + 3] }, pk1[2];
  for (int j = 1; j <= S; j++)
  {
    const double t = (j < S) ? T * j / S : T;
    nb_eval_xy(cxy, ii, t, pk1);
    const double *pik, *pik1;
    const int stride = num_pol * (S + 1) * 2;
    if (ii > num_pol - 1)
    {
      pik = samp + ((size_t)(num_pol - 1) * (S + 1) + S) * 2;
      pik1 = pik;
    }
    else
    {
      pik = samp + ((size_t)ii * (S + 1) + (j - 1)) * 2;
      pik1 = samp + ((size_t)ii * (S + 1) + j) * 2;
    }
    const int nadd = nb_collect_toadd<NL>(g, cx, pk, nullptr, pk1, pik, stride, pik1, stride, known, toadd, tcap);
    if (nadd < 0) return -1;
    if (g.lane == 0)
    {
      int r = 0;
      if (es.n_alpha + nadd > limit)
        r = 1;
      else
      {
        // active_cases_old (kinodynamic_search.cpp:813, :880 / :909, :971) is only compared where active_cases can
        // have grown, i.e. at the ids of alphasToAdd: remember (id, value before) for those instead of copying and
        // scanning the whole array (N + M entries in global memory) every step
        for (int i = 0; i < nadd; i++) pairs[2 * i] = toadd[2 * i], pairs[2 * i + 1] = es.active[toadd[2 * i] - 1];
        if (nb_add_alpha_beta(toadd, nadd, es, pk, cx))
          r = -1;
        else
        {
          for (int i = 0; i < nadd; i++)
          {
            const int a = pairs[2 * i] - 1, was = pairs[2 * i + 1];
            if (a >= cx.N) continue;
            if (was < 2 && es.active[a] >= 2) r = 1;
            if (was >= 2 && es.active[a] > was) r = 1;
          }
          if (r == 0) nb_update_bend_pts(es, pk1, cx);
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
    if (r != 0) return r;
    pk[0] = pk1[0];
    pk[1] = pk1[1];
  }
  return 0;
}

// ---- K3 task: one agent per warp-group.  mode 0 predict, 1 rollout, 2 check-given-pwp
// shared memory of the staged work state (nb_entangle_task, mode 4), after the crossing lists
inline size_t nb_ent_stage_bytes(int NA, int cap) { return (size_t)cap * 20 + (size_t)NA * 4 + 16; }

struct NbEntArgs
{
  int mode, N, M, cap, bp_max, num_pol, S, tcap;
  double T;
  const int* agent_id;
  const uint8_t* known;  // [B][N]
  const int* bp_cnt;
  const double* bp_xy;
  const double* pb;
  const double* strep;
  nb_ent_state st;       // in (and out for predict / check)
  nb_ent_state out;      // rollout: [B][9][...]
  const int* n_int;
  const double* coeff;   // [B][3][8][4]
  const double* samp;    // [B or 1][N][8][S+1][2], or [G][...] with samp_group
  int samp_shared;
  const int* samp_group; // optional [B]: agent b reads block samp_group[b] of samp (shared-window groups)
  int samp0_stride;      // doubles between the first samples of consecutive agents in samp0 (2: packed [B][N][2];
                         // num_pol (S+1) 2: samp0 points into a samples array and is group-indexed like samp)
  int strep_per_agent;   // 1: strep is [N][M][2][2], indexed by the planning agent's id - 1
  // post-check (mode 4, Neptune::safetyCheckAfterReplan neptune.cpp:735-752)
  const unsigned char* late;     // [B][N] trajectory of j received after time_init_opt_
  const double* late_recs;       // [N][NB_REC] those trajectories
  const double* t_start;         // [B]
  const int* bp_cnt_late;        // [N] bend points carried by the late messages
  const double* bp_xy_late;      // [N][bp_max][2]
  double* psamp;                 // scratch [B][N][S+1][2]: interval-0 samples per planning agent
  unsigned char* pknown;         // scratch [B][N]
  int stage;                     // mode 4: the work state lives in shared memory (host: it fits, nb_ent_stage_bytes)
  int phase;                     // mode 4: 0 whole post-check, 1 up to PredictAlphasBetas (independent of the optimised
                                 // trajectory: runs beside the QP), 2 entangleCheckGivenPwp on the state phase 1 left
  const double* prev_pos;        // predict
  const double* prev_pos_agent;
  const double* cur;
  const double* samp0;
  // tracker (mode 3)
  const int* bp_cnt_prev;
  const double* bp_xy_prev;
  double* prev_pos_rw;           // [B][N+1][2] in/out
  double* prev_pos_agent_rw;     // [B][N][2] in/out
  const double* latest;          // [B][N][2]
  const double* elapsed_ms;      // [B]
  int* result;           // done / entangled / tracker result
  int* err;
};

template <int NL>
NB_HD void nb_entangle_task(const Group<NL>& g, int b, const NbEntArgs& a, int* toadd /*[2 tcap][2]: list, then pairs*/, int* flag /*[4]*/)
{
  const int NA = a.N + a.M;
  NbEntCtx cx;
  cx.N = a.N, cx.M = a.M, cx.self = a.agent_id[b] - 1, cx.cap = a.cap, cx.bp_max = a.bp_max, cx.bp_stride = 2 * a.bp_max;
  cx.pb = a.pb, cx.strep = a.strep + (a.strep_per_agent ? (size_t)cx.self * 4 * a.M : 0), cx.bp_cnt = a.bp_cnt, cx.bp_xy = a.bp_xy;
  cx.use_alt = nullptr, cx.bp_cnt_alt = nullptr, cx.bp_xy_alt = nullptr;
  const uint8_t* known = a.known + (size_t)b * a.N;
  int* pairs = toadd + 2 * a.tcap;  // shared: (id, active before) of the entries of one step
  NbEntState es;
  if (a.mode == 4)
  {  // post-check: nothing to do without a late trajectory (need_to_rerun_entanglecheck, neptune.cpp:722, :743)
    int mine = 0;
    for (int j = g.lane; j < a.N; j += NL) mine |= (j != cx.self && a.late[(size_t)b * a.N + j]) ? 1 : 0;
    if (!g.any(mine))
    {
      if (g.lane == 0) a.result[b] = 0;
      return;
    }
  }
  const bool resume = a.mode == 4 && a.phase == 2;   // the work state is where phase 1 left it
  if (a.mode == 1 || a.mode == 4)
  {  // work on slot 0 of the output (mode 4: on a scratch copy, the reference's local ent_state_begin)
    const size_t o = (size_t)b * (a.mode == 1 ? 9 : 1);
    es.alpha = a.out.alpha + o * a.cap * 2, es.beta = a.out.beta + o * a.cap, es.bend = a.out.bend + o * a.cap;
    es.active = a.out.active + o * NA;
  }
  bool staged = false;
#if defined(__CUDA_ARCH__)
  if (a.mode == 4 && a.stage)
  {  // The work state in SHARED memory: the list automaton (one lane) walks alpha / beta / bend / active entry by entry, a
     // chain of dependent loads that costs ~300 cycles per step from L2 and 29 from shared memory.  Only the used
     // entries are copied.  Layout after the crossing lists: beta [cap] f64 | alpha [cap][2] | bend [cap] | active [NA].
    double* s_beta = reinterpret_cast<double*>(flag + 4);
    int* s_alpha = reinterpret_cast<int*>(s_beta + a.cap);
    int* s_bend = s_alpha + 2 * a.cap;
    int* s_active = s_bend + a.cap;
    const int* g_alpha = resume ? es.alpha : a.st.alpha + (size_t)b * a.cap * 2;
    const double* g_beta = resume ? es.beta : a.st.beta + (size_t)b * a.cap;
    const int* g_bend = resume ? es.bend : a.st.bend + (size_t)b * a.cap;
    const int* g_active = resume ? es.active : a.st.active + (size_t)b * NA;
    const int na = resume ? a.out.cnt[2 * b] : a.st.cnt[2 * b], nbd = resume ? a.out.cnt[2 * b + 1] : a.st.cnt[2 * b + 1];
    for (int q = g.lane; q < na; q += NL) s_alpha[2 * q] = g_alpha[2 * q], s_alpha[2 * q + 1] = g_alpha[2 * q + 1], s_beta[q] = g_beta[q];
    for (int q = g.lane; q < nbd; q += NL) s_bend[q] = g_bend[q];
    for (int q = g.lane; q < NA; q += NL) s_active[q] = g_active[q];
    es.alpha = s_alpha, es.beta = s_beta, es.bend = s_bend, es.active = s_active;
    staged = true;
  }
#endif
  if (staged)
  {
  }
  else if ((a.mode == 1 || a.mode == 4) && !resume)
  {
    for (int q = g.lane; q < a.cap; q += NL)
    {
      es.alpha[2 * q] = a.st.alpha[((size_t)b * a.cap + q) * 2], es.alpha[2 * q + 1] = a.st.alpha[((size_t)b * a.cap + q) * 2 + 1];
      es.beta[q] = a.st.beta[(size_t)b * a.cap + q];
      es.bend[q] = a.st.bend[(size_t)b * a.cap + q];
    }
    for (int q = g.lane; q < NA; q += NL) es.active[q] = a.st.active[(size_t)b * NA + q];
  }
  else if (!(a.mode == 4 && resume))
  {
    es.alpha = a.st.alpha + (size_t)b * a.cap * 2, es.beta = a.st.beta + (size_t)b * a.cap;
    es.bend = a.st.bend + (size_t)b * a.cap, es.active = a.st.active + (size_t)b * NA;
  }
  es.n_alpha = resume ? a.out.cnt[2 * b] : a.st.cnt[2 * b];
  es.n_bend = resume ? a.out.cnt[2 * b + 1] : a.st.cnt[2 * b + 1];
  g.sync();
  const size_t samp_blk = (size_t)a.N * a.num_pol * (a.S + 1) * 2;
  const double* samp = a.samp ? a.samp + (a.samp_group ? (size_t)a.samp_group[b] * samp_blk : (a.samp_shared ? 0 : (size_t)b * samp_blk))
                              : nullptr;
  const double* cxy = a.coeff ? a.coeff + (size_t)b * 96 : nullptr;
  int bad = 0;
  if (a.mode == 0)
  {  // Neptune::PredictAlphasBetas neptune.cpp:976-1008
    const double* pp = a.prev_pos + (size_t)b * (a.N + 1) * 2;
    const double* cur = a.cur + 2 * b;
    const int s0 = a.samp0_stride > 0 ? a.samp0_stride : 2;
    const double* samp0 = a.samp0 + (a.samp_group && s0 != 2 ? (size_t)a.samp_group[b] * samp_blk : (size_t)b * a.N * s0);
    const int nadd = nb_collect_toadd<NL>(g, cx, pp + 2 * a.N, pp, cur, a.prev_pos_agent + (size_t)b * a.N * 2, 2, samp0, s0,
                                          known, toadd, a.tcap);
    if (nadd < 0)
      bad = 1;
    else if (g.lane == 0)
    {
      if (nb_add_alpha_beta(toadd, nadd, es, pp + 2 * a.N, cx))
        flag[3] = 1;
      else
      {
        flag[3] = 0;
        nb_update_bend_pts(es, cur, cx);
      }
      a.st.cnt[2 * b] = es.n_alpha;
      a.st.cnt[2 * b + 1] = es.n_bend;
    }
    g.sync();
    if (!bad && flag[3]) bad = 1;
  }
  else if (a.mode == 3)
  {  // NeptuneRos::updateEntStateStaticObs neptune_ros.cpp:798-850
    double* pp = a.prev_pos_rw + (size_t)b * (a.N + 1) * 2;
    double* ppa = a.prev_pos_agent_rw + (size_t)b * a.N * 2;
    const double* cur = a.cur + 2 * b;
    const double dx = NB_SUB(pp[2 * a.N], cur[0]), dy = NB_SUB(pp[2 * a.N + 1], cur[1]);
    if (sqrt(NB_ADD(NB_MUL(dx, dx), NB_MUL(dy, dy))) < 0.05 && a.elapsed_ms[b] < 100)
    {  // the gate of :803-804 (group-uniform)
      if (g.lane == 0) a.result[b] = 1;
      return;
    }
    if (g.lane == 0) flag[2] = 0;
    g.sync();
    const double pkN[2] = { pp[2 * a.N], pp[2 * a.N + 1] };
    const int nadd = nb_collect_track<NL>(g, cx, a.bp_cnt_prev, a.bp_xy_prev, pp, ppa, a.latest + (size_t)b * a.N * 2, cur, toadd,
                                          a.tcap, &flag[2]);
    g.sync();
    if (nadd < 0)
      bad = 1;
    else if (g.lane == 0)
    {
      if (nb_add_alpha_beta(toadd, nadd, es, pkN, cx))
        flag[3] = 1;
      else
      {
        flag[3] = 0;
        nb_update_bend_pts(es, cur, cx);
      }
      pp[2 * a.N] = cur[0], pp[2 * a.N + 1] = cur[1];
      a.st.cnt[2 * b] = es.n_alpha;
      a.st.cnt[2 * b + 1] = es.n_bend;
      a.result[b] = flag[2] ? -flag[2] : 0;
    }
    g.sync();
    if (!bad && flag[3]) bad = 1;
  }
  else if (a.mode == 4)
  {  // Neptune::safetyCheckAfterReplan, entanglement half (neptune.cpp:735-752)
    const int S1 = a.S + 1, n = a.n_int[b];
    double* ps = a.psamp + (size_t)b * a.N * S1 * 2;
    unsigned char* pk = a.pknown + (size_t)b * a.N;
    const unsigned char* late = a.late + (size_t)b * a.N;
    // SampledPointsForAll[j]: re-sampled over [times.front(), times.back()] of pwp_optimized for the late agents
    // (SamplePointsOfIntervals, :737-738: deltaT = n T / num_pol), the planning-time samples for the others; only
    // interval 0 is ever read (PredictAlphasBetas :986, entangleCheckGivenPwp :899 / :982)
    const double t0 = a.t_start[b], t1 = NB_ADD(t0, NB_MUL((double)n, a.T));
    cx.use_alt = late, cx.bp_cnt_alt = a.bp_cnt_late, cx.bp_xy_alt = a.bp_xy_late;
    const double* pp = a.prev_pos + (size_t)b * (a.N + 1) * 2;
    const double* cur = a.cur + 2 * b;
    int r = 0;
    if (!resume)
    {
      for (int j = g.lane; j < a.N; j += NL)
      {
        const bool lt = j != cx.self && late[j];
        pk[j] = (known[j] || lt) ? 1 : 0;
        if (lt)
          nb_sample_interval(a.late_recs + (size_t)j * NB_REC, t0, t1, a.num_pol, a.S, 0, ps + (size_t)j * S1 * 2);
        else
          for (int q = 0; q < S1 * 2; q++) ps[(size_t)j * S1 * 2 + q] = samp ? samp[(size_t)j * a.num_pol * S1 * 2 + q] : 0.0;
      }
      g.sync();
      // PredictAlphasBetas on the updated samples (:747-749)
      const int nadd = nb_collect_toadd<NL>(g, cx, pp + 2 * a.N, pp, cur, a.prev_pos_agent + (size_t)b * a.N * 2, 2, ps, S1 * 2,
                                            pk, toadd, a.tcap);
      if (nadd < 0)
        bad = 1;
      else
      {
        if (g.lane == 0)
        {
          if (nb_add_alpha_beta(toadd, nadd, es, pp + 2 * a.N, cx))
            flag[3] = 1;
          else
          {
            flag[3] = 0;
            nb_update_bend_pts(es, cur, cx);
          }
        }
        g.sync();
        if (flag[3]) bad = 1;
      }
      if (a.phase == 1)
      {  // the state PredictAlphasBetas left: phase 2 resumes from it
        if (staged)
        {  // back to the scratch slot in global memory
          int* o_alpha = a.out.alpha + (size_t)b * a.cap * 2;
          double* o_beta = a.out.beta + (size_t)b * a.cap;
          int* o_bend = a.out.bend + (size_t)b * a.cap;
          int* o_active = a.out.active + (size_t)b * NA;
          if (g.lane == 0) flag[1] = es.n_alpha, flag[2] = es.n_bend;   // only the automaton's lane knows the new lengths
          g.sync();
          const int na = flag[1], nbd = flag[2];
          for (int q = g.lane; q < na; q += NL) o_alpha[2 * q] = es.alpha[2 * q], o_alpha[2 * q + 1] = es.alpha[2 * q + 1], o_beta[q] = es.beta[q];
          for (int q = g.lane; q < nbd; q += NL) o_bend[q] = es.bend[q];
          for (int q = g.lane; q < NA; q += NL) o_active[q] = es.active[q];
        }
        if (g.lane == 0) a.out.cnt[2 * b] = es.n_alpha, a.out.cnt[2 * b + 1] = es.n_bend;
        if (bad && g.lane == 0) *a.err = 2;
        return;
      }
    }
    if (!bad && n > 0)  // entangleCheckGivenPwp (:750), interval 0 only; the scratch samples hold one interval per agent
      r = nb_ent_interval_pass<NL>(g, es, cx, cxy, 0, ps, pk, 1, a.S, a.T, 3 * NA, toadd, a.tcap, pairs, flag);
    if (r < 0) bad = 1;
    if (g.lane == 0) a.result[b] = r > 0 ? 1 : 0;
  }
  else if (a.mode == 2)
  {  // entangleCheckGivenPwp: interval 0 only (kinodynamic_search.cpp:899, :982-983)
    int r = 0;
    if (a.n_int[b] > 0)
      r = nb_ent_interval_pass<NL>(g, es, cx, cxy, 0, samp, known, a.num_pol, a.S, a.T, 3 * NA, toadd, a.tcap, pairs, flag);
    if (r < 0) bad = 1;
    if (g.lane == 0)
    {
      a.result[b] = r > 0 ? 1 : 0;
      a.st.cnt[2 * b] = es.n_alpha;
      a.st.cnt[2 * b + 1] = es.n_bend;
    }
  }
  else
  {  // rollout: states after 0..n intervals
    const int n = a.n_int[b];
    int done = n;
    for (int i = 0; i <= NB_NPOL; i++)
    {
      const size_t o = (size_t)b * 9 + i;
      if (i > 0)
      {  // copy state i-1 forward, then advance it when i <= n
        int* al = a.out.alpha + o * a.cap * 2;
        double* be = a.out.beta + o * a.cap;
        int* bd = a.out.bend + o * a.cap;
        int* ac = a.out.active + o * NA;
        for (int q = g.lane; q < a.cap; q += NL)
        {
          al[2 * q] = es.alpha[2 * q], al[2 * q + 1] = es.alpha[2 * q + 1];
          be[q] = es.beta[q];
          bd[q] = es.bend[q];
        }
        for (int q = g.lane; q < NA; q += NL) ac[q] = es.active[q];
        g.sync();
        es.alpha = al, es.beta = be, es.bend = bd, es.active = ac;
        if (i <= n && !bad)
        {
          const int r = nb_ent_interval_pass<NL>(g, es, cx, cxy, i - 1, samp, known, a.num_pol, a.S, a.T, NA, toadd,
                                                 a.tcap, pairs, flag);
          if (r < 0) bad = 1;
          if (r > 0 && done == n) done = i - 1;
        }
      }
      if (g.lane == 0)
      {
        a.out.cnt[2 * o] = es.n_alpha;
        a.out.cnt[2 * o + 1] = es.n_bend;
      }
    }
    if (g.lane == 0) a.result[b] = bad ? -1 : done;
  }
  if (bad && g.lane == 0) *a.err = 2;
}

