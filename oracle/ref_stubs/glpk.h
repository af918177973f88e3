// Stand-in for <glpk.h> (TEST INFRASTRUCTURE, oracle/_ref build): the handful of GLPK calls the reference's
// separator_glpk.cpp makes.  The model is recorded as the reference builds it; glp_simplex hands it to a solver callback
// installed by the test (HiGHS through scipy) and keeps the status and the primal values it returns.  Nothing of GLPK's
// algorithm is restated: with a zero objective the vertex an LP solver returns is its own business, only the
// feasible / infeasible answer is defined by the model.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef struct glp_prob glp_prob;
typedef struct
{
  int msg_lev;
  int reserved[40];
} glp_smcp;
enum
{
  GLP_MIN = 1,
  GLP_MAX = 2
};
enum
{
  GLP_FR = 1,
  GLP_LO = 2,
  GLP_UP = 3,
  GLP_DB = 4,
  GLP_FX = 5
};
enum
{
  GLP_UNDEF = 1,
  GLP_FEAS = 2,
  GLP_INFEAS = 3,
  GLP_NOFEAS = 4,
  GLP_OPT = 5,
  GLP_UNBND = 6
};
enum
{
  GLP_MSG_OFF = 0,
  GLP_MSG_ERR = 1,
  GLP_MSG_ON = 2,
  GLP_MSG_ALL = 3
};
glp_prob* glp_create_prob(void);
void glp_delete_prob(glp_prob* p);
int glp_free_env(void);
void glp_set_prob_name(glp_prob* p, const char* name);
void glp_set_obj_dir(glp_prob* p, int dir);
int glp_add_rows(glp_prob* p, int n);
int glp_add_cols(glp_prob* p, int n);
void glp_set_row_name(glp_prob* p, int i, const char* name);
void glp_set_col_name(glp_prob* p, int j, const char* name);
void glp_set_row_bnds(glp_prob* p, int i, int type, double lb, double ub);
void glp_set_col_bnds(glp_prob* p, int j, int type, double lb, double ub);
void glp_set_obj_coef(glp_prob* p, int j, double c);
void glp_load_matrix(glp_prob* p, int ne, const int ia[], const int ja[], const double ar[]);
int glp_init_smcp(glp_smcp* parm);
int glp_simplex(glp_prob* p, const glp_smcp* parm);
int glp_get_status(glp_prob* p);
double glp_get_obj_val(glp_prob* p);
double glp_get_col_prim(glp_prob* p, int j);
int glp_write_lp(glp_prob* p, const void* parm, const char* fname);

// the solver hook: rows / cols, row bounds (type, lb, ub) [rows], col bounds [cols], objective [cols] and direction, the
// matrix as triplets (1-based, as GLPK takes them); writes x [cols] and returns a GLP_* status
typedef int (*ref_lp_solver)(int rows, int cols, const int* row_type, const double* row_lb, const double* row_ub, const int* col_type,
                             const double* col_lb, const double* col_ub, const double* obj, int dir, int ne, const int* ia, const int* ja,
                             const double* ar, double* x);
void ref_set_lp_solver(ref_lp_solver f);
#ifdef __cplusplus
}
#endif
