// GRBModel::optimize of the Gurobi stand-in (oracle/ref_stubs/gurobi_c++.h; TEST INFRASTRUCTURE, oracle/_ref build):
// flattens the model the reference's solver_gurobi_poly.cpp has built into dense arrays over the ACTIVE variables and
// hands it to the solver callback the test installed (HiGHS through scipy).  Every optimize() call is therefore also
// a recording of the reference's own addObjective / addConstraints output.
#include <vector>

#include "gurobi_c++.h"

static ref_qp_solver g_solver = nullptr;
static void* g_user = nullptr;
extern "C" void ref_set_qp_solver(ref_qp_solver f, void* user)
{
  g_solver = f;
  g_user = user;
}

void GRBModel::optimize()
{
  std::vector<int> col(vars.size(), -1);
  int nv = 0;
  for (size_t i = 0; i < vars.size(); i++)
    if (vars[i].state == ACTIVE) col[i] = nv++;
  std::vector<double> lb(nv), ub(nv), Q((size_t)nv * nv, 0.0), c(nv, 0.0);
  for (size_t i = 0; i < vars.size(); i++)
    if (col[i] >= 0) lb[col[i]] = vars[i].lb, ub[col[i]] = vars[i].ub;
  for (auto& t : objective.lin.t) c[col[t.var]] += t.coef;
  for (auto& e : objective.q) Q[(size_t)col[e.v1] * nv + col[e.v2]] += e.coef;
  std::vector<double> A, rhs, Qc, qc, qrhs;
  std::vector<char> sense, qsense;
  int nl = 0, nq = 0;
  for (auto& r : lin)
    if (r.state == ACTIVE)
    {
      A.resize((size_t)(nl + 1) * nv, 0.0);
      for (auto& t : r.e.lin.t) A[(size_t)nl * nv + col[t.var]] += t.coef;
      rhs.push_back(-r.e.lin.cst);
      sense.push_back(r.sense);
      nl++;
    }
  for (auto& r : quad)
    if (r.state == ACTIVE)
    {
      Qc.resize((size_t)(nq + 1) * nv * nv, 0.0);
      qc.resize((size_t)(nq + 1) * nv, 0.0);
      for (auto& t : r.e.lin.t) qc[(size_t)nq * nv + col[t.var]] += t.coef;
      for (auto& e : r.e.q) Qc[((size_t)nq * nv + col[e.v1]) * nv + col[e.v2]] += e.coef;
      qrhs.push_back(-r.e.lin.cst);
      qsense.push_back(r.sense);
      nq++;
    }
  ref_qp_model m;
  m.nvar = nv, m.lb = lb.data(), m.ub = ub.data(), m.Q = Q.data(), m.c = c.data(), m.c0 = objective.lin.cst;
  m.nlin = nl, m.A = A.data(), m.sense = sense.data(), m.rhs = rhs.data();
  m.nquad = nq, m.Qc = Qc.data(), m.qc = qc.data(), m.qsense = qsense.data(), m.qrhs = qrhs.data();
  m.time_limit = time_limit, m.non_convex = non_convex;
  std::vector<double> xs(nv, 0.0);
  status = GRB_LOADED, sol_count = 0;
  if (g_solver) status = g_solver(&m, xs.data(), &sol_count, g_user);
  x.assign(vars.size(), 0.0);
  for (size_t i = 0; i < vars.size(); i++)
    if (col[i] >= 0) x[i] = xs[col[i]];
}
