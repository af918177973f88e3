// nb_tables.h -- host-side precomputation of solver constants and reduced-space QP tables.
#pragma once
#include "../../include/neptune_b200.h"
#include "nb_common.cuh"

void nb_build_consts(const nb_params* p, NbConsts* c);
// table for n intervals (1..NB_NPOL), mode 0 (terminal v/a equalities) or 1 (fallback); false on a
// degenerate basis (cannot happen for T_span > 0)
bool nb_build_table(const NbConsts* cs, int n, int mode, NbQpTable* t);
