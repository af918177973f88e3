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
#include "../../neptune_b200/csrc/nb_lines.cuh"
#include "../../neptune_b200/csrc/nb_qp.cuh"
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
                                 const nb_replan_args* a)
{
  NbConsts cs;
  nb_build_consts(par, &cs);
  std::vector<NbQpTable> tabs(2 * NB_NPOL);
  for (int mode = 0; mode < 2; mode++)
    for (int n = 1; n <= NB_NPOL; n++)
      if (!nb_build_table(&cs, n, mode, &tabs[mode * NB_NPOL + n - 1])) return -2;
  const int B = a->B, N = par->num_agents, M = par->num_static, NH = a->n_hull_slots;
  const int LS = NH + N + M + par->ent_slots;
  const int RS = 6 * NB_NFEAT_AX + 4 * NB_NPOL * LS;
  NbLinesIn in;
  in.agent_id = a->agent_id, in.n_int = a->n_int, in.coeff_init = a->coeff_init, in.NH = NH;
  in.hull_ptr = a->hull_ptr, in.hull_xy = a->hull_xy, in.nih0 = a->nih0, in.st_ptr = st_ptr, in.st_xy = st_xy;
  in.esv_cnt = a->esv_cnt, in.esv_alpha = a->esv_alpha, in.esv_active = a->esv_active;
  in.bp_cnt = a->bp_cnt, in.bp_xy = a->bp_xy, in.pb = pb;
  std::vector<double> lines((size_t)NB_NPOL * LS * 3), cl((size_t)NB_NPOL * LS * 3), rows((size_t)4 * RS);
  std::vector<uint8_t> ok((size_t)NB_NPOL * LS);
  int lstart[9], err = 0;
  NbQpShared* sh = new NbQpShared();
  Group<1> g(0);
  for (int b = 0; b < B; b++)
  {
    for (int i = 0; i < NB_NPOL; i++)
      nb_lines_task<1>(0, b, i, cs, in, lines.data() + (size_t)i * LS * 3, ok.data() + (size_t)i * LS, &err);
    if (a->lines) memcpy(a->lines + (size_t)b * NB_NPOL * LS * 3, lines.data(), sizeof(double) * lines.size());
    if (a->line_ok) memcpy(a->line_ok + (size_t)b * NB_NPOL * LS, ok.data(), ok.size());
    const int n = a->n_int[b];
    const double* ci = a->coeff_init + (size_t)b * 96;
    const int nl = nb_compact_lines<1>(g, n, LS, lines.data(), ok.data(), cl.data(), lstart);
    NbQpRows R;
    R.s = rows.data(), R.lam = R.s + RS, R.dsa = R.lam + RS, R.dla = R.dsa + RS, R.cl = cl.data(), R.lstart = lstart;
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
