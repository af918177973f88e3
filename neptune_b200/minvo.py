"""MINVO basis constants and the derived solver constants (host side, numpy).

Numbers: ``mt::basisConverter`` ``A_pos_mv_rest`` / ``A_vel_mv_rest``, reference
``neptune/include/mader_types.hpp:152-162``.  Derived matrices follow
``PolySolverGurobi::PolySolverGurobi`` (``neptune/src/solver_gurobi_poly.cpp:35-62, :93-97, :126-129``).
"""
from __future__ import annotations

import numpy as np

A_POS_MV = np.array([
    [-3.4416308968564117698463178385282, 6.9895481477801393310755884158425,
     -4.4622887507045296828778191411402, 0.91437149978080234369315348885721],
    [6.6792587327074839365081970754545, -11.845989901556746914934592496138,
     5.2523596690684613008670567069203, 0.0],
    [-6.6792587327074839365081970754545, 8.1917862965657040064115790301003,
     -1.5981560640774179482548333908198, 0.085628500219197656306846511142794],
    [3.4416308968564117698463178385282, -3.3353445427890959784633650997421,
     0.80808514571348655231020075007109, -0.0000000000000000084567769453869345852581318467855]])

A_VEL_MV = np.array([[1.5, -2.36602540378444, 0.933012701892219],
                     [-3.0, 3.0, 0.0],
                     [1.5, -0.633974596215561, 0.0669872981077807]])


def solver_basis(T: float):
    """(Ainv, V, Ainv01): ``A_rest_pos_basis_inverse_`` (4x4), ``A_rest_vel_basis_inverse321_`` (3x3)
    of the back end for interval length T, and the [0,1] position inverse used by hull generation
    (``neptune.cpp:64``)."""
    Ainv = np.linalg.inv(A_POS_MV @ np.diag([T ** -3, T ** -2, 1.0 / T, 1.0]))
    V = np.linalg.inv(A_VEL_MV @ np.diag([T ** -2, 1.0 / T, 1.0]))
    V[0] *= 3.0
    V[1] *= 2.0
    return Ainv, V, np.linalg.inv(A_POS_MV)
