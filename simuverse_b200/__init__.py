"""simuverse_b200 — B200-native D2Q9 lattice-Boltzmann step behind the interface of
jinleili/simuverse's LBM player (``D2Q9Node`` / ``FluidSimulator``).

The compute path is hand-written CUDA for sm_100a in ``csrc/`` behind the C ABI of
``include/lbm_b200.h``; this package is the thin host-side mirror of the reference's Rust
interface over that ABI.  Importing it loads ``_native/liblbm_b200.so`` and fails if the
library has not been built: there is no CPU / PyTorch fallback.
"""
from . import wire
from ._capi import (FLAG_KERNEL_GENERIC, FLAG_MACRO_EVERY_STEP, FLAG_NO_GRAPH, FLAG_AA, FLAG_NO_FUSE, LIB_PATH, MACRO_F32_PLANES, MACRO_RGBA16F,
                    PRESET_POROUS, LbmError, lib)
from .d2q9_node import D2Q9Node, SettingObj, init_lattice_material, init_porous_material, lbm_uniform_new
from .fluid_simulator import FluidSimulator, init_trajectory_particles, particle_grid

__all__ = [
    "D2Q9Node", "FluidSimulator", "SettingObj", "LbmError", "init_lattice_material", "init_porous_material",
    "lbm_uniform_new", "init_trajectory_particles", "particle_grid", "wire", "lib", "LIB_PATH",
    "FLAG_KERNEL_GENERIC", "FLAG_MACRO_EVERY_STEP", "FLAG_NO_GRAPH", "FLAG_AA", "FLAG_NO_FUSE", "MACRO_F32_PLANES", "MACRO_RGBA16F", "PRESET_POROUS",
]
