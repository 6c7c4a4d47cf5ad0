"""steps_b200 -- B200-native direct-summation gravity engine behind StePS's force-call boundary.

The package holds only what the hot path needs: ``csrc/`` (CUDA kernels + the C ABI of
``include/steps_b200.h``), the ctypes binding (``_lib``), the host-side mirror of the reference's
force/step interface (``api``) and the synthetic IC shapes of BASELINE.json (``ic``).
"""
from ._lib import StepsError, TOPO_R3, TOPO_S1R2_LOOKUP, TOPO_S1R2_NOLOOKUP, TOPO_T3, LIB_PATH  # noqa: F401
from .api import (  # noqa: F401
    CALCULATE_Hubble_param,
    Engine,
    Globals,
    UNIT_T,
    UNIT_V,
    calculate_softening_length,
    calculate_S1R2ewald_correction_table,
    calculate_t3_ewald_lookup_table,
    fma_peak,
    fma_peak_sustained,
    force_entry,
    forces,
    forces_periodic,
    forces_periodic_z,
    friedmann_solver_step,
    get_cylindrical_force_table,
    partition,
    s1r2_ewald_defaults,
    t3_ewald_defaults,
)
