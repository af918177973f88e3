// nb_lines.cuh -- generation of every separating-line LP of one (agent, interval).
//
// Replaces the separation part of PolySolverGurobi::addConstraints (reference
// neptune/src/solver_gurobi_poly.cpp:475-495 other agents, :521-553 bases, :556-615 static
// obstacles, :620-642 + addEntangleConstraintForIJCase :715-784 non-entangling), i.e. every
// separator_solver_->solveModel call of a replan.  LPs are solved on the INITIAL control points
// ctrlPtsInit_ (:232-243).  Results land in fixed slots so the order is deterministic:
//   [0,NH) other agents' hulls | [NH,NH+N) bases | [NH+N,NH+N+M) static | [..,+ent_slots) tether.
#pragma once
#include "nb_common.cuh"
#include "nb_prune.cuh"
#include "nb_sep.cuh"

struct NbLinesIn
{
  const int* agent_id;       // [B]
  const int* n_int;          // [B]
  const double* coeff_init;  // [B][3][8][4]
  int NH;
  const int64_t* hull_ptr;   // [B*NH*8+1] (or [B*NH*8] with hull_cnt)
  const int* hull_cnt;       // optional [B*NH*8]
  const double* hull_xy;
  const double* nih0;        // [B][N][8][2] (or [G][N][8][2] with nih0_group)
  const int* nih0_group;     // optional [B]
  const uint8_t* hull_known; // optional [B][N]: shared-window hulls (group-shaped hull_xy / hull_cnt)
  const int64_t* st_ptr;     // [M+1]
  const double* st_xy;
  const int* esv_cnt;        // [B][9][2]
  const int* esv_alpha;      // [B][9][cap][2]
  const int* esv_active;     // [B][9][N+M]
  const int* bp_cnt;         // [N]
  const double* bp_xy;       // [N][bp_max][2]
  const double* pb;          // [N][2]
};

NB_HD double nb_dist(const double* a, const double* b)
{
  const double dx = a[0] - b[0], dy = a[1] - b[1];
  return sqrt(dx * dx + dy * dy);
}

// control points of interval i of the initial path: [cx; cy] * Ainv  (:232-243)
NB_HD void nb_ctrl_pts(const NbConsts& cs, const double* ci /*[3][8][4] of agent*/, int i, double cp[8])
{
  for (int k = 0; k < 4; k++)
  {
    double x = 0, y = 0;
    for (int r = 0; r < 4; r++)
    {
      x += ci[4 * i + r] * cs.Ainv[r * 4 + k];
      y += ci[32 + 4 * i + r] * cs.Ainv[r * 4 + k];
    }
    cp[2 * k] = x;
    cp[2 * k + 1] = y;
  }
}

// Threads tid, tid+NT, ... of the group handle slots; tid 0 additionally walks the (few)
// non-entangling constraints.  lines: [LS][3], ok: [LS] of this (b, i).  Sets *err on overflow.
// Ordered compaction of the lines the QP must see: cl_out[q] = (n0, n1, 1 - d) of the q-th slot with flag[s] == 1,
// *ncl_out = their number.  This is the list K4 reads (one contiguous run per interval, no slot scan there).
template <int NT>
NB_HD void nb_emit_kept(const Cta<NT>& cta, int LS, const double* lines, const uint8_t* flag, int* red, double* cl_out,
                        int* ncl_out)
{
  int total = 0;
  for (int base = 0; base < LS; base += NT)
  {
    const int s = base + cta.tid;
    const int mine = (s < LS) && (flag[s] == 1);
    int pos = 0, cnt = mine;
#if defined(__CUDA_ARCH__)
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    const int lane = cta.tid & 31, w = cta.tid >> 5;
    if (lane == 0) red[w] = __popc(bal);
    cta.sync();
    pos = __popc(bal & ((1u << lane) - 1u));
    cnt = 0;
    for (int q = 0; q < (NT + 31) / 32; q++)
    {
      if (q < w) pos += red[q];
      cnt += red[q];
    }
#endif
    if (mine)
    {
      const double* l = lines + (size_t)3 * s;
      double* o = cl_out + (size_t)3 * (total + pos);
      o[0] = l[0];
      o[1] = l[1];
      o[2] = 1.0 - l[2];
    }
    total += cnt;
    cta.sync();
  }
  if (cta.tid == 0) *ncl_out = total;
}

// prune == false keeps every solved line (test hook of the host emulation: pruned == unpruned optimum)
template <int NT>
NB_HD void nb_lines_task(int tid, int b, int i, const NbConsts& cs, const NbLinesIn& in, double* lines,
                         uint8_t* ok, uint8_t* keep, const NbPruneShared& ps, int* err, double* cl_out, int* ncl_out,
                         bool prune = true)
{
  const int N = cs.N, M = cs.M, NH = in.NH;
  const int LS = NH + N + M + cs.ent_slots;
  const int n = in.n_int[b];
  if (i >= n)
  {
    for (int s = tid; s < LS; s += NT)
    {
      ok[s] = 0;
      keep[s] = 0;
    }
    if (tid == 0) *ncl_out = 0;
    return;
  }
  double cp[8];
  nb_ctrl_pts(cs, in.coeff_init + (size_t)b * 96, i, cp);
  const double base_radius = 0.7;
  for (int s = tid; s < NH + N + M; s += NT)
  {
    uint8_t res = 0;
    double l[3] = { 0, 0, 0 };
    if (s < NH)
    {  // other agents :477-495
      int64_t o0;
      int cnt;
      if (in.hull_known)
      {  // shared windows: hull (group, s, i) of nb_hulls_batch, fixed stride; own slot / unknown agents empty
        const size_t src = ((size_t)in.nih0_group[b] * NH + s) * 8 + i;
        o0 = (int64_t)src * 24;
        cnt = (s == in.agent_id[b] - 1 || !in.hull_known[(size_t)b * NH + s]) ? 0 : in.hull_cnt[src];
      }
      else
      {
        const size_t hk = ((size_t)b * NH + s) * 8 + i;
        o0 = in.hull_ptr[hk];
        cnt = in.hull_cnt ? in.hull_cnt[hk] : (int)(in.hull_ptr[hk + 1] - o0);
      }
      if (cnt > 0) res = nb_separate(in.hull_xy + 2 * o0, cnt, true, cp, 4, l) ? 1 : 2;
    }
    else if (s < NH + N)
    {  // bases :521-553 (own base included)
      const int j = s - NH;
      const double* pbj = in.pb + 2 * j;
      bool close = false;
      for (int k = 0; k < 4; k++)
        if (nb_dist(cp + 2 * k, pbj) < base_radius * 3) close = true;
      if (close)
      {
        const double hull[8] = { pbj[0] + base_radius, pbj[1] + base_radius, pbj[0] + base_radius, pbj[1] - base_radius,
                                 pbj[0] - base_radius, pbj[1] + base_radius, pbj[0] - base_radius, pbj[1] - base_radius };
        res = nb_separate(hull, 4, false, cp, 4, l) ? 1 : 2;
      }
    }
    else
    {  // static obstacles :556-593
      const int j = s - NH - N;
      const double* sv = in.st_xy + 2 * in.st_ptr[j];
      const int cnt = (int)(in.st_ptr[j + 1] - in.st_ptr[j]);
      bool close = false;
      double dist = nb_dist(cp, sv);
      for (int k = 0; k < 3 && !close; k++)
      {
        dist -= nb_dist(cp + 2 * (k + 1), cp + 2 * k);
        if (dist < 0) close = true;
      }
      for (int k = 0; k < cnt - 1 && !close; k++)
      {
        dist -= nb_dist(sv + 2 * (k + 1), sv + 2 * k);
        if (dist < 0) close = true;
      }
      if (close) res = nb_separate(sv, cnt, true, cp, 4, l) ? 1 : 2;
    }
    ok[s] = res;
    lines[3 * s] = l[0];
    lines[3 * s + 1] = l[1];
    lines[3 * s + 2] = l[2];
  }
  {  // non-entangling :620-642 -> :715-784
    const int cap = cs.ent_cap, NA = N + M;
    const int* alpha = in.esv_alpha + ((size_t)b * 9 + i) * cap * 2;
    const int n_alpha = in.esv_cnt[((size_t)b * 9 + i) * 2];
    const int* active = in.esv_active + ((size_t)b * 9 + i) * NA;
    const int self = in.agent_id[b] - 1;
    for (int e = tid; e < cs.ent_slots; e += NT) ok[NH + N + M + e] = 0;
    Cta<NT>(tid).sync();
    // an agent with active_cases == 1 appears in the alphas list: walk the (short) list instead of all N
    // agents, in increasing agent order as the reference's loop over j does
    if (tid == 0)
    {
      int eslot = 0;
      double hulldist = 0.0;
      for (int k = 0; k < 3; k++) hulldist += nb_dist(cp + 2 * (k + 1), cp + 2 * k);
      int prev_j = -1;
      while (true)
      {
        // next agent id (> prev_j) present in the list
        int j = N;
        for (int jj = 0; jj < n_alpha; jj++)
        {
          const int id = alpha[2 * jj] - 1;
          if (id > prev_j && id < j) j = id;
        }
        if (j >= N) break;
        prev_j = j;
        if (j == self || active[j] != 1) continue;
        int case_id = 0;
        for (int jj = 0; jj < n_alpha; jj++)
          if (alpha[2 * jj] == j + 1) case_id = alpha[2 * jj + 1];
        if (case_id == 0) continue;
        const int nb = in.bp_cnt[j];
        const double* bend = in.bp_xy + (size_t)2 * cs.bp_max * j;
        const double* posj = in.nih0 + (((size_t)(in.nih0_group ? in.nih0_group[b] : b) * N + j) * 8 + i) * 2;
        if (nb < 1 || posj[0] != posj[0]) continue;
        bool over = false;
        for (int k = 1; k < nb + 1; k++)
        {
          if (k == case_id) continue;
          double pA[2], pB[2];
          if (k == 1)
          {  // :719-724 ray beyond agent j
            pA[0] = (1 - cs.long_length) * bend[2 * (nb - 1)] + cs.long_length * posj[0];
            pA[1] = (1 - cs.long_length) * bend[2 * (nb - 1) + 1] + cs.long_length * posj[1];
            pB[0] = posj[0];
            pB[1] = posj[1];
          }
          else
          {  // :725-730 tether segment k-2 -> k-1
            pA[0] = bend[2 * (k - 2)];
            pA[1] = bend[2 * (k - 2) + 1];
            pB[0] = bend[2 * (k - 1)];
            pB[1] = bend[2 * (k - 1) + 1];
          }
          if (nb_dist(pA, cp) - hulldist > 0 && nb_dist(pB, cp) - hulldist > 0) continue;  // :743-745
          if (eslot >= cs.ent_slots)
          {
            *err = 1;
            over = true;
            break;
          }
          const double Aset[4] = { pA[0], pA[1], pB[0], pB[1] };
          double l[3];
          const int sl = NH + N + M + eslot;
          ok[sl] = nb_separate(Aset, 2, false, cp, 4, l) ? 1 : 2;  // :751
          lines[3 * sl] = l[0];
          lines[3 * sl + 1] = l[1];
          lines[3 * sl + 2] = l[2];
          eslot++;
        }
        if (over) break;
      }
    }
  }
  Cta<NT> cta(tid);
  cta.sync();
  nb_prune_lines<NT>(cta, LS, lines, ok, cp, ps, keep);
  cta.sync();
  nb_emit_kept<NT>(cta, LS, lines, prune ? keep : ok, ps.red, cl_out, ncl_out);
}
