"""entity_b200 -- B200-native PIC timestep hot path behind a C ABI.

The product is ``libentity_b200.so`` (hand-written sm_100a CUDA, see ``csrc/`` and
``include/entity_b200.h``). This package is the thin Python side used by the tests and the
benchmark: ctypes bindings plus torch tensors as device memory. It never falls back to a CPU
implementation: importing :mod:`entity_b200.lib` raises if the shared library is missing.
"""
from .lib import (  # noqa: F401
    Context, Config, Grid, Prtls, Pusher, LIB_PATH, load, nghosts_for,
    PUSHER_NONE, PUSHER_PHOTON, PUSHER_BORIS, PUSHER_VAY, PUSHER_GCA,
    DRAG_NONE, DRAG_SYNCHROTRON, DRAG_COMPTON,
    PBC_NONE, PBC_PERIODIC, PBC_ABSORB, PBC_REFLECT, PBC_AXIS,
    FBC_NONE, FBC_PERIODIC, FBC_CONDUCTOR, FBC_AXIS, FBC_SYNC,
    DEPOSIT_ATOMIC, DEPOSIT_ORDERED, DEPOSIT_AGGREGATED, EB200Error,
    STATS_NPART, STATS_N, STATS_RHO, STATS_CHARGE, STATS_T,
    SDIST_UNIFORM, SDIST_TABLE, SDIST_REPLENISH, SDIST_REPLENISH_TABLE, SDIST_ATMOSPHERE,
)
