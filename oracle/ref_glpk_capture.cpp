// Implementation of the GLPK stand-in declared in ref_stubs/glpk.h (TEST INFRASTRUCTURE): records the model, defers the
// solve to the installed callback.
#include <glpk.h>
#include <cstddef>
#include <vector>

struct glp_prob
{
  int dir = GLP_MIN;
  std::vector<int> row_type, col_type;
  std::vector<double> row_lb, row_ub, col_lb, col_ub, obj, x;
  std::vector<int> ia, ja;
  std::vector<double> ar;
  int status = GLP_UNDEF;
};

static ref_lp_solver g_solver = nullptr;

extern "C" {
void ref_set_lp_solver(ref_lp_solver f) { g_solver = f; }
glp_prob* glp_create_prob(void) { return new glp_prob; }
void glp_delete_prob(glp_prob* p) { delete p; }
int glp_free_env(void) { return 0; }
void glp_set_prob_name(glp_prob*, const char*) {}
void glp_set_obj_dir(glp_prob* p, int dir) { p->dir = dir; }
int glp_add_rows(glp_prob* p, int n)
{
  const int first = (int)p->row_type.size() + 1;
  p->row_type.resize(p->row_type.size() + n, GLP_FR), p->row_lb.resize(p->row_type.size(), 0.0), p->row_ub.resize(p->row_type.size(), 0.0);
  return first;
}
int glp_add_cols(glp_prob* p, int n)
{
  const int first = (int)p->col_type.size() + 1;
  p->col_type.resize(p->col_type.size() + n, GLP_FX), p->col_lb.resize(p->col_type.size(), 0.0), p->col_ub.resize(p->col_type.size(), 0.0);
  p->obj.resize(p->col_type.size(), 0.0), p->x.resize(p->col_type.size(), 0.0);
  return first;
}
void glp_set_row_name(glp_prob*, int, const char*) {}
void glp_set_col_name(glp_prob*, int, const char*) {}
void glp_set_row_bnds(glp_prob* p, int i, int type, double lb, double ub) { p->row_type[i - 1] = type, p->row_lb[i - 1] = lb, p->row_ub[i - 1] = ub; }
void glp_set_col_bnds(glp_prob* p, int j, int type, double lb, double ub) { p->col_type[j - 1] = type, p->col_lb[j - 1] = lb, p->col_ub[j - 1] = ub; }
void glp_set_obj_coef(glp_prob* p, int j, double c) { p->obj[j - 1] = c; }
void glp_load_matrix(glp_prob* p, int ne, const int ia[], const int ja[], const double ar[])
{
  p->ia.assign(ia + 1, ia + 1 + ne), p->ja.assign(ja + 1, ja + 1 + ne), p->ar.assign(ar + 1, ar + 1 + ne);
}
int glp_init_smcp(glp_smcp* parm)
{
  parm->msg_lev = GLP_MSG_ALL;
  return 0;
}
int glp_simplex(glp_prob* p, const glp_smcp*)
{
  p->status = GLP_UNDEF;
  if (g_solver)
    p->status = g_solver((int)p->row_type.size(), (int)p->col_type.size(), p->row_type.data(), p->row_lb.data(), p->row_ub.data(), p->col_type.data(),
                         p->col_lb.data(), p->col_ub.data(), p->obj.data(), p->dir, (int)p->ar.size(), p->ia.data(), p->ja.data(), p->ar.data(),
                         p->x.data());
  return 0;
}
int glp_get_status(glp_prob* p) { return p->status; }
double glp_get_obj_val(glp_prob* p)
{
  double z = 0;
  for (std::size_t j = 0; j < p->obj.size(); j++) z += p->obj[j] * p->x[j];
  return z;
}
double glp_get_col_prim(glp_prob* p, int j) { return p->x[j - 1]; }
int glp_write_lp(glp_prob*, const void*, const char*) { return 0; }
}
