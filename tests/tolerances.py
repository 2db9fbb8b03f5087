"""Parity tolerances, stated once.

BASELINE.json (north_star): forward state trajectories match the reference within 1e-5 relative
(here: max |a - b| / max |b| over a state vector).  That bound is met for the well-conditioned
articulations (CartPole, Ant, Hopper, Cheetah: cond(H) < 1e3).  For Humanoid and SNUHumanoid the
joint-space inertia has cond(H + armature) ~ 1e4 (tests/test_emu_golden.py::
test_reference_solve_is_conditioning_limited measures it from the golden H): the REFERENCE's own
fp32 Cholesky solve of H q'' = tau deviates from an fp64 solve by up to ~2e-5 relative, so two
correct fp32 implementations (different summation order, FMA contraction on the GPU) cannot agree
better than that over 48 substeps.  Those two models are held to 3e-5.  Gradients (no tolerance is
stated by the north star) are held to 5e-5 of the largest component.
"""
FWD_RTOL = 1e-5
GRAD_RTOL = 5e-5
_ILL_CONDITIONED = {"HumanoidEnv": 3e-5, "SNUHumanoidEnv": 3e-5}


def fwd_rtol(env_name):
    return _ILL_CONDITIONED.get(env_name, FWD_RTOL)


# bf16 tape (dfx_set_tape_dtype(1), config C2 "bf16 states"): v, a, f_tot of every tape row stored as bf16, arithmetic fp32.
# The forward pass is untouched (bit-identical); gradients against the fp32 reference goldens, relative to the largest
# component (measured on the host emulation: Humanoid 2.5e-3, SNU 2.7e-2 through the muscle wrenches; actions <= 4e-4).
BF16_TAPE_STATE_GRAD_RTOL = {"SNUHumanoidEnv": 5e-2}
BF16_TAPE_STATE_GRAD_RTOL_DEFAULT = 5e-3
BF16_TAPE_ACTION_GRAD_RTOL = 1e-3
