"""lattice_qcd_rs_b200 -- B200-native (sm_100a CUDA) pure-gauge SU(3) update path behind the trait surface of
lattice-qcd-rs v0.2.1.

  _capi   : ctypes binding of the C ABI (include/lqcd_b200.h)
  state   : host-side mirror of the reference's traits for this path (LatticeStateDefault,
            LatticeStateEFSyncDefault, SymplecticEulerCuda, HybridMonteCarloDiagnostic, HeatBathSweep, ...)
  serde_io: serde_json / bincode layouts of the reference's serde-derived state structs (checkpoints)
  dist    : one-process-per-GPU domain decomposition plumbing (torch.distributed: NCCL on GPUs)

The compute path is the CUDA library only; importing works without a GPU (so that the build can be checked),
creating a state does not.
"""
from ._capi import (INTEGRATOR_OMELYAN, INTEGRATOR_SYMPLECTIC_EULER, OMELYAN_LAMBDA, FLAG_GAUSS_FUSED, FLAG_GENERIC_KERNELS, FLAG_UNIFORM_DIRECTION, FLAG_NO_KICK_MERGE, FLAG_PAULI3_FIXED, LEAP_LEAP, LEAP_SYNC, OR_REVERSE, OR_ROTATION, OR_SU2_SUBGROUPS, SYMPLECTIC,
                    SYNC_LEAP, SYNC_SYNC, Context, LqError, load)

__all__ = ["Context", "LqError", "load", "SYNC_SYNC", "LEAP_LEAP", "SYNC_LEAP", "LEAP_SYNC", "SYMPLECTIC",
           "OR_ROTATION", "OR_REVERSE", "OR_SU2_SUBGROUPS", "FLAG_PAULI3_FIXED", "FLAG_NO_KICK_MERGE", "FLAG_GAUSS_FUSED", "FLAG_GENERIC_KERNELS", "FLAG_UNIFORM_DIRECTION",
           "INTEGRATOR_SYMPLECTIC_EULER", "INTEGRATOR_OMELYAN", "OMELYAN_LAMBDA"]
