// nb_emul.cpp -- SINGLE-LANE HOST EMULATION of the device code in neptune_b200/csrc/*.cuh.
//
// CPU-side debugging/test harness only: it compiles the very same lane-strided kernels with one
// lane (NL = NT = 1) so that algorithmic regressions can be caught by `pytest -m "not gpu"` in a
// container without a GPU.  It is never loaded by the neptune_b200 package and is not a fallback:
// the product path is libneptune_b200.so on an sm_100a device only.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/neptune_b200.h"
#include "../../neptune_b200/csrc/nb_common.cuh"
#include "../../neptune_b200/csrc/nb_entangle.cuh"
#include "../../neptune_b200/csrc/nb_lines.cuh"
#include "../../neptune_b200/csrc/nb_qp.cuh"
#include "../../neptune_b200/csrc/nb_search.cuh"
#include "../../neptune_b200/csrc/nb_sep.cuh"
#include "../../neptune_b200/csrc/nb_tables.h"

extern "C" int emul_separate(const double* A, int nA, int a_polygon, const double* B, int nB, double* out)
{
  return nb_separate(A, nA, a_polygon != 0, B, nB, out) ? 1 : 0;
}

extern "C" int emul_table(const nb_params* par, int n, int mode, NbQpTable* out)
{
  NbConsts cs;
  nb_build_consts(par, &cs);
  return nb_build_table(&cs, n, mode, out) ? 0 : -1;
}

extern "C" int emul_replan_batch(const nb_params* par, const double* pb, const int64_t* st_ptr, const double* st_xy,
                                 const nb_replan_args* a, int prune, int* n_lines_out)
{
  NbConsts cs;
  nb_build_consts(par, &cs);
  std::vector<NbQpTable> tabs(2 * NB_NPOL);
  for (int mode = 0; mode < 2; mode++)
    for (int n = 1; n <= NB_NPOL; n++)
      if (!nb_build_table(&cs, n, mode, &tabs[mode * NB_NPOL + n - 1])) return -2;
  const int B = a->B, N = par->num_agents, M = par->num_static, NH = a->n_hull_slots;
  const int LS = NH + N + M + par->ent_slots;
  const int RS = NB_ROW_LINE0 + 4 * NB_NPOL * LS;
  NbLinesIn in;
  in.agent_id = a->agent_id, in.n_int = a->n_int, in.coeff_init = a->coeff_init, in.NH = NH;
  in.hull_ptr = a->hull_ptr, in.hull_cnt = a->hull_cnt, in.hull_xy = a->hull_xy, in.nih0 = a->nih0, in.nih0_group = a->nih0_group, in.hull_known = a->hull_known, in.st_ptr = st_ptr, in.st_xy = st_xy;
  in.esv_cnt = a->esv_cnt, in.esv_alpha = a->esv_alpha, in.esv_active = a->esv_active;
  in.bp_cnt = a->bp_cnt, in.bp_xy = a->bp_xy, in.pb = pb;
  std::vector<double> lines((size_t)NB_NPOL * LS * 3), cl((size_t)NB_NPOL * LS * 3), rows((size_t)5 * RS);
  std::vector<uint8_t> ok((size_t)NB_NPOL * LS), keep((size_t)NB_NPOL * LS), valid(LS + 1);
  std::vector<double> px(LS + 1), py(LS + 1);
  int red[1], hull[NB_PRUNE_KMAX + 1], misc[8];
  NbPruneShared ps;
  ps.px = px.data(), ps.py = py.data(), ps.valid = valid.data(), ps.red = red, ps.hull = hull, ps.misc = misc;
  int lstart[9], ncl[NB_NPOL], err = 0;
  const double* clb[NB_NPOL];
  NbQpShared* sh = new NbQpShared();
  Group<1> g(0);
  for (int b = 0; b < B; b++)
  {
    for (int i = 0; i < NB_NPOL; i++)
      nb_lines_task<1>(0, b, i, cs, in, lines.data() + (size_t)i * LS * 3, ok.data() + (size_t)i * LS,
                       keep.data() + (size_t)i * LS, ps, &err, cl.data() + (size_t)i * LS * 3, &ncl[i], prune != 0);
    if (a->lines) memcpy(a->lines + (size_t)b * NB_NPOL * LS * 3, lines.data(), sizeof(double) * lines.size());
    if (a->line_ok) memcpy(a->line_ok + (size_t)b * NB_NPOL * LS, ok.data(), ok.size());
    const int n = a->n_int[b];
    const double* ci = a->coeff_init + (size_t)b * 96;
    int nl = 0;
    for (int i = 0; i < NB_NPOL; i++)
    {  // one run of kept lines per interval, as the kernel's global-row path addresses them
      lstart[i] = nl;
      clb[i] = cl.data() + (size_t)i * LS * 3 - (size_t)3 * nl;
      if (i < n) nl += ncl[i];
    }
    for (int i = n; i <= NB_NPOL; i++) lstart[i] = nl;
    if (n_lines_out) n_lines_out[b] = nl;
    NbQpRows R;
    R.s = rows.data(), R.lam = R.s + RS, R.dsa = R.lam + RS, R.dla = R.dsa + RS, R.inv = R.dla + RS, R.clb = clb, R.lstart = lstart, R.l2i = nullptr;
    double xout[96], obj = 0;
    int it0 = 0, it1 = 0, status = NB_STATUS_FAILED;
    bool okq = nb_qp_solve<1>(g, cs, &tabs[n - 1], sh, R, ci, nl, xout, &it0, &obj);
    if (okq)
      status = NB_STATUS_OK;
    else
    {
      okq = nb_qp_solve<1>(g, cs, &tabs[NB_NPOL + n - 1], sh, R, ci, nl, xout, &it1, &obj);
      if (okq) status = NB_STATUS_FALLBACK;
    }
    const double T = cs.T, qp[4] = { T * T * T, T * T, T, 1.0 };
    double pfx = 0, pfy = 0;
    for (int r = 0; r < 4; r++)
    {
      pfx += qp[r] * ci[4 * (n - 1) + r];
      pfy += qp[r] * ci[32 + 4 * (n - 1) + r];
    }
    const double dx = ci[3] - pfx, dy = ci[32 + 3] - pfy;
    const bool keep_z = sqrt(dx * dx + dy * dy) < 1.0;
    double* co = a->coeff_out + (size_t)b * 96;
    for (int q = 0; q < 96; q++)
    {
      const int ax = q / 32, r = q % 32;
      double v = ci[q];
      if (okq && r < 4 * n && !(ax == 2 && keep_z)) v = xout[q];
      co[q] = v;
    }
    a->obj[b] = okq ? obj : 0.0;
    a->status[b] = status;
    a->iters[2 * b] = it0;
    a->iters[2 * b + 1] = it1;
  }
  delete sh;
  return err ? NB_ERR_CAPACITY : 0;
}

#include "../../neptune_b200/csrc/nb_entangle.cuh"

extern "C" int emul_entangle(const nb_params* par, const double* pb, const double* strep, int mode, int B,
                             const int32_t* agent_id, const uint8_t* known, const int32_t* bp_cnt, const double* bp_xy,
                             nb_ent_state st, nb_ent_state out, const int32_t* n_int, const double* coeff,
                             const double* samp, int samp_shared, const double* prev_pos, const double* prev_pos_agent,
                             const double* cur, const double* samp0, int32_t* result)
{
  NbEntArgs a;
  memset(&a, 0, sizeof(a));
  const int N = par->num_agents, M = par->num_static;
  a.mode = mode, a.N = N, a.M = M, a.cap = par->ent_cap, a.bp_max = par->bp_max, a.num_pol = par->num_pol;
  a.S = par->samples, a.T = par->T_span;
  int tcap = 4 * (N + M) + 16;
  a.tcap = tcap > 1024 ? 1024 : tcap;
  a.agent_id = agent_id, a.known = known, a.bp_cnt = bp_cnt, a.bp_xy = bp_xy, a.pb = pb, a.strep = strep;
  a.st = st, a.out = out, a.n_int = n_int, a.coeff = coeff, a.samp = samp, a.samp_shared = samp_shared;
  a.prev_pos = prev_pos, a.prev_pos_agent = prev_pos_agent, a.cur = cur, a.samp0 = samp0, a.result = result;
  std::vector<int> toadd(4 * a.tcap + 8);
  int err = 0;
  a.err = &err;
  Group<1> g(0);
  for (int b = 0; b < B; b++) nb_entangle_task<1>(g, b, a, toadd.data(), toadd.data() + 4 * a.tcap);
  return err ? NB_ERR_CAPACITY : 0;
}

// k_entangle mode 4 (Neptune::safetyCheckAfterReplan, entanglement half) on one host lane
extern "C" int emul_postcheck_entangle(const nb_params* par, const double* pb, const double* strep, int B, const int32_t* agent_id,
                                       const uint8_t* known, const uint8_t* late, const int32_t* bp_cnt, const double* bp_xy,
                                       const int32_t* bp_cnt_late, const double* bp_xy_late, nb_ent_state st,
                                       const double* prev_pos, const double* prev_pos_agent, const double* cur,
                                       const int32_t* n_int, const double* coeff, const double* t_start, const double* samp,
                                       const double* late_recs, int32_t* entangled)
{
  NbEntArgs a;
  memset(&a, 0, sizeof(a));
  const int N = par->num_agents, M = par->num_static, NA = N + M, cap = par->ent_cap, S = par->samples;
  a.mode = 4, a.N = N, a.M = M, a.cap = cap, a.bp_max = par->bp_max, a.num_pol = par->num_pol, a.S = S, a.T = par->T_span;
  int tcap = 4 * NA + 16;
  a.tcap = tcap > 1024 ? 1024 : tcap;
  a.agent_id = agent_id, a.known = known, a.bp_cnt = bp_cnt, a.bp_xy = bp_xy, a.pb = pb, a.strep = strep;
  a.st = st, a.n_int = n_int, a.coeff = coeff, a.samp = samp, a.prev_pos = prev_pos, a.prev_pos_agent = prev_pos_agent, a.cur = cur;
  a.late = late, a.late_recs = late_recs, a.t_start = t_start, a.bp_cnt_late = bp_cnt_late, a.bp_xy_late = bp_xy_late;
  a.result = entangled;
  std::vector<int32_t> wc((size_t)B * 2), wa((size_t)B * cap * 2), wb((size_t)B * cap), wact((size_t)B * NA);
  std::vector<double> wbeta((size_t)B * cap), ps((size_t)B * N * (S + 1) * 2);
  std::vector<uint8_t> pk((size_t)B * N);
  a.out.cnt = wc.data(), a.out.alpha = wa.data(), a.out.beta = wbeta.data(), a.out.bend = wb.data(), a.out.active = wact.data();
  a.psamp = ps.data(), a.pknown = pk.data();
  std::vector<int> toadd(4 * a.tcap + 8);
  int err = 0;
  a.err = &err;
  Group<1> g(0);
  for (int b = 0; b < B; b++) nb_entangle_task<1>(g, b, a, toadd.data(), toadd.data() + 4 * a.tcap);
  return err ? NB_ERR_CAPACITY : 0;
}

#include "../../neptune_b200/csrc/nb_hull.cuh"

extern "C" int emul_hulls(const nb_params* par, int B, const double* t_start, const double* recs, const uint8_t* known,
                          double delta, double* hull_xy, int32_t* hull_cnt, int64_t* hull_ptr, double* nih0, double* samp,
                          int32_t* idx)
{
  NbConsts cs;
  nb_build_consts(par, &cs);
  int err = 0;
  for (int b = 0; b < B; b++)
    for (int j = 0; j < cs.N; j++)
    {
      for (int i = 0; i < NB_NPOL; i++)
      {
        const size_t k = ((size_t)b * cs.N + j) * NB_NPOL + i;
        hull_ptr[k] = (int64_t)k * NB_HMAX;
        int cnt = 0, id2[2] = { -1, -1 };
        double n0[2] = { NAN, NAN };
        if (i < cs.num_pol && known[(size_t)b * cs.N + j])
        {
          const double t0 = t_start[b] + (double)i * cs.T, t1 = t_start[b] + (double)(i + 1) * cs.T;
          cnt = nb_hull_of_window(cs, recs + (size_t)j * NB_REC, t0, t1, delta, hull_xy + k * NB_HMAX * 2, n0, id2);
          if (cnt < 0 || cnt > NB_HMAX) err = 3, cnt = 0;
        }
        hull_cnt[k] = cnt;
        nih0[2 * k] = n0[0], nih0[2 * k + 1] = n0[1];
        if (idx) idx[2 * k] = id2[0], idx[2 * k + 1] = id2[1];
      }
      double* out = samp + ((size_t)b * cs.N + j) * cs.num_pol * (cs.S + 1) * 2;
      if (known[(size_t)b * cs.N + j])
        nb_sample_points(cs, recs + (size_t)j * NB_REC, t_start[b], t_start[b] + cs.T * (double)cs.num_pol, out, nullptr);
    }
  return err ? NB_ERR_CAPACITY : 0;
}

extern "C" int emul_postcheck(const nb_params* par, int B, const int32_t* n_int, const double* coeff, const double* t_start,
                              const double* recs, const uint8_t* late, double delta, int32_t* collide)
{
  NbConsts cs;
  nb_build_consts(par, &cs);
  for (int b = 0; b < B; b++)
  {
    collide[b] = 0;
    for (int j = 0; j < cs.N; j++)
      if (late[(size_t)b * cs.N + j] &&
          nb_pwp_collides(cs, coeff + (size_t)b * 96, n_int[b], t_start[b], recs + (size_t)j * NB_REC, delta) > 0)
        collide[b] = 1;
  }
  return 0;
}

extern "C" int emul_compose(int B, const double* t, const uint8_t* has_prev, const double* prev, const double* now,
                            double* out, int32_t* n_pieces)
{
  for (int b = 0; b < B; b++)
  {
    if (!has_prev[b])
    {
      memcpy(out + (size_t)b * NB_REC, now + (size_t)b * NB_REC, sizeof(double) * NB_REC);
      n_pieces[b] = (int)now[(size_t)b * NB_REC];
    }
    else
      n_pieces[b] = nb_compose_records(t[b], prev + (size_t)b * NB_REC, now + (size_t)b * NB_REC, out + (size_t)b * NB_REC);
  }
  return 0;
}

// ---- K0 front-end search with one host lane: the same nb_search_task the device runs with 25 warps
struct EmulCta
{
  int tid = 0, nthreads = 1, warp = 0, nwarps = 1, lane = 0, aux_tid = 0, aux_n = 1;
  bool child = true, aux = true;
  void sync() const {}
  void sync_aux() const {}
  int any(int p) const { return p; }
};

extern "C" int emul_search_batch(const nb_params* par, const nb_search_params* sp, const double* pb, const int64_t* st_ptr,
                                 const double* st_xy, const double* strep, const double* st_longest, const nb_search_args* u)
{
  NbConsts cs;
  nb_build_consts(par, &cs);
  NbSearchArgs a;
  memset(&a, 0, sizeof(a));
  nb_search_fill_par(*par, *sp, cs, &a.p);
  const NbSearchPar& p = a.p;
  const int NA = p.N + p.M;
  a.agent_id = u->agent_id, a.init = u->init, a.goal = u->goal, a.coeffs_z = u->coeffs_z, a.group = u->group;
  a.hull_xy = u->hull_xy, a.hull_cnt = u->hull_cnt, a.samp = u->samp, a.known = u->known;
  a.st_ptr = st_ptr, a.st_xy = st_xy, a.strep = strep, a.st_longest = st_longest, a.pb = pb;
  a.bp_cnt = u->bp_cnt, a.bp_xy = u->bp_xy, a.es = u->es, a.comb = u->comb, a.comb_shared = u->comb_shared;
  a.status = u->status, a.solved = u->solved, a.n_int = u->n_int, a.coeff = u->coeff, a.esv = u->esv;
  a.stats = u->stats, a.cost = u->cost;
  const size_t mn = (size_t)p.max_nodes, chs = (size_t)nb_search_ch_stride(p);
  std::vector<NbInt4> meta(mn), hash((size_t)p.hcap);
  std::vector<double> kin(mn * NB_SEARCH_KIN), beta(mn * p.ecap), gh(mn * 2), ng(mn), chd(nb_search_chd_stride(p));
  std::vector<uint8_t> fcode(nb_search_fcode_bytes(p));
  std::vector<int> alpha(mn * p.ecap * 2), bend(mn * p.ecap), heap(mn), chi((size_t)p.nchild * chs + NA);
  int err = 0;
  a.nd_meta = meta.data(), a.nd_kin = kin.data(), a.nd_alpha = alpha.data(), a.nd_beta = beta.data(), a.nd_bend = bend.data();
  a.hash = hash.data(), a.heap_g = heap.data(), a.gh_g = gh.data(), a.nd_g = ng.data(), a.ch_int = chi.data(), a.ch_dbl = chd.data(), a.fcode_g = fcode.data(), a.err = &err;
  NbSearchShared* sh = new NbSearchShared();
  // one workspace block, reused by every agent in turn (workspace index 0)
  for (int b = 0; b < u->B; b++)
  {
    EmulCta cta;
    nb_search_task<EmulCta, 1>(cta, a, b, 0, sh, nullptr, 0);
  }
  delete sh;
  return err ? NB_ERR_CAPACITY : 0;
}

// ---- one tick of the online tracker (k_entangle mode 3) with one host lane
extern "C" int emul_track_batch(const nb_params* par, const double* pb, const double* strep, int B, const int32_t* agent_id,
                                const int32_t* bp_cnt, const double* bp_xy, const int32_t* bp_cnt_prev, const double* bp_xy_prev,
                                nb_ent_state st, double* prev_pos, double* prev_pos_agent, const double* latest,
                                const double* cur, const double* elapsed_ms, int32_t* result)
{
  NbEntArgs a;
  memset(&a, 0, sizeof(a));
  const int N = par->num_agents, M = par->num_static;
  a.mode = 3, a.N = N, a.M = M, a.cap = par->ent_cap, a.bp_max = par->bp_max, a.num_pol = par->num_pol, a.S = par->samples;
  a.T = par->T_span, a.tcap = 4 * (N + M) + 16;
  a.agent_id = agent_id, a.bp_cnt = bp_cnt, a.bp_xy = bp_xy, a.pb = pb, a.strep = strep, a.st = st;
  a.bp_cnt_prev = bp_cnt_prev, a.bp_xy_prev = bp_xy_prev, a.prev_pos_rw = prev_pos, a.prev_pos_agent_rw = prev_pos_agent;
  a.latest = latest, a.cur = cur, a.elapsed_ms = elapsed_ms, a.result = result;
  std::vector<uint8_t> known((size_t)B * N, 1);
  std::vector<int> toadd((size_t)4 * a.tcap);
  int err = 0, flag[4];
  a.known = known.data(), a.err = &err;
  Group<1> g(0);
  for (int b = 0; b < B; b++) nb_entangle_task<1>(g, b, a, toadd.data(), flag);
  return err ? NB_ERR_CAPACITY : 0;
}
