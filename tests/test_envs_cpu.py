"""Host-side logic of the env layer that needs no GPU: the tiled Model equals the reference's per-env
parse (bit for bit), and stepping without CUDA fails loudly instead of falling back."""
import numpy as np
import pytest
import torch

from conftest import ENVS
from emu_util import load_golden


@pytest.mark.parametrize("name", ENVS)
def test_tiled_model_equals_reference_model(name):
    import diffrl_b200.dflex_api as df
    d, ref = load_golden(name)
    n = int(d["meta/num_envs"])
    import os
    from diffrl_b200.envs.base import ASSET_DIR
    arrays = dict(np.load(os.path.join(ASSET_DIR, name + ".npz")))
    m = df.model_from_articulation(arrays, n, "cpu", ground=bool(d["meta/ground"]))
    for field, want in ref.items():
        if field in ("ground",) or not hasattr(m, field):
            continue
        got = getattr(m, field)
        if not torch.is_tensor(got):
            continue
        got = got.numpy()
        if got.size == 0 and np.asarray(want).size == 0:
            continue
        if field == "joint_target" and want.shape[0] > got.shape[0]:
            want = want[: got.shape[0]]     # reference quirk: envs/hopper.py:119 grows the list
        assert got.shape == np.asarray(want).shape and np.array_equal(got, want), field
    for cnt in ("link_count", "joint_coord_count", "joint_dof_count", "shape_count", "contact_count", "muscle_count"):
        assert int(getattr(m, cnt)) == int(d["meta/" + cnt]), cnt


def test_no_cpu_fallback():
    """The product must refuse to simulate without the CUDA path."""
    from diffrl_b200 import _capi
    from diffrl_b200.envs import AntEnv
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    env = AntEnv(num_envs=2, device="cpu", no_grad=False)
    env.clear_grad()
    env.reset()
    with pytest.raises(_capi.DfxError):
        env.step(torch.zeros(2, 8))


def test_c_abi_exports_every_declared_symbol():
    import ctypes, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "dfx.h")).read()
    names = sorted(set(re.findall(r"\b(dfx_[a-z_]+)\s*\(", header)))
    assert "dfx_step_forward" in names and "dfx_step_backward" in names
    lib = ctypes.CDLL(os.path.join(root, "diffrl_b200", "libdfx.so"))
    for n in names:
        assert hasattr(lib, n), n


def test_c_abi_rejects_invalid_arguments_before_touching_the_gpu():
    """Error behaviour of the boundary: every entry point validates its arguments first and returns
    cudaErrorInvalidValue (1) -- no launch, no exception, nothing that needs a device -- and the ctypes mirrors of the parameter
    structs have the layout the header declares (a size mismatch would shift every pointer of DfxEnvTransition)."""
    import ctypes, os
    from diffrl_b200 import env_ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = ctypes.CDLL(os.path.join(root, "diffrl_b200", "libdfx.so"))
    env_ops._bind(lib)
    INVALID = 1
    null = ctypes.c_void_p(None)
    tr, tra = env_ops.DfxEnvTransition(kind=1), env_ops.DfxEnvTransitionAdj(kind=1)
    assert lib.dfx_env_step_forward(null, 4, 16, 16, 0.01, null, null, null, null, null, null, null, null, null, ctypes.byref(tr), null) == INVALID
    assert lib.dfx_env_step_backward(null, 4, 16, 16, 0.01, null, null, null, null, ctypes.byref(tra), null, null, null, null) == INVALID
    wp = env_ops.DfxWalkerParams(num_q=15, num_qd=14, num_act=8, num_obs=37)
    assert lib.dfx_walker_transition_forward(ctypes.byref(wp), 0, *([null] * 15)) == INVALID        # n <= 0
    assert lib.dfx_walker_transition_forward(ctypes.byref(wp), 4, *([null] * 15)) == INVALID        # null rows
    pp = env_ops.DfxPlanarParams(num_q=2, num_qd=2, num_act=1, num_obs=5, kind=7)
    assert lib.dfx_planar_transition_forward(ctypes.byref(pp), 4, *([null] * 15)) == INVALID
    # struct layouts: the ctypes mirrors against sizeof() as the library was compiled, and against the header read by hand
    from diffrl_b200.modelpack import DfxActionMap, DfxDerived, DfxModelDesc
    for which, mirror in enumerate((DfxModelDesc, DfxDerived, env_ops.DfxWalkerParams, env_ops.DfxPlanarParams, DfxActionMap,
                                    env_ops.DfxEnvTransition, env_ops.DfxEnvTransitionAdj)):
        assert lib.dfx_abi_sizeof(which) == ctypes.sizeof(mirror), (which, mirror.__name__)
    assert lib.dfx_abi_sizeof(99) == -1
    assert ctypes.sizeof(env_ops.DfxWalkerParams) == 4 * (11 + 5 + 3 + 4 + 3 + 3)
    assert ctypes.sizeof(env_ops.DfxPlanarParams) == 4 * (8 + 9)
    head = 4 + ctypes.sizeof(env_ops.DfxWalkerParams) + ctypes.sizeof(env_ops.DfxPlanarParams)
    head += (-head) % 8
    assert ctypes.sizeof(env_ops.DfxEnvTransition) == head + 8 * 11
    assert ctypes.sizeof(env_ops.DfxEnvTransitionAdj) == head + 8 * 13


@pytest.mark.parametrize("name", ["AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "CartPoleSwingUpEnv", "HopperEnv", "CheetahEnv"])
def test_tile_kernels_are_compiled_for_the_joint_types_of_the_six_articulations(name):
    """The size-specialised tile kernels carry the articulation's joint types as a compile-time constant (Pack::jmask,
    csrc/dfx_launch.h tile_joint_mask): branches of absent types are not in the binary.  dfx_pack_create keeps a pack with any
    other joint type on the run-time-generic lane-group kernels -- correct, but several times slower -- so the masks compiled
    into the library must cover what the six DiffRL articulations really contain (pack built on the host emulation: no GPU)."""
    import ctypes, os
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path)
    from emu_util import EmuSim, load_golden
    root = os.path.dirname(sys_path)
    lib = ctypes.CDLL(os.path.join(root, "diffrl_b200", "libdfx.so"))
    d, model = load_golden(name)
    sim = EmuSim(model, int(d["meta/num_envs"]))
    q = lambda what: sim.lib.emu_pack_query(sim.pack, what)
    L, D, Q, C, M = q(0), q(1), q(2), q(3), q(4)
    have = q(12)                                                    # DFX_QUERY_JOINT_MASK
    n = int(d["meta/num_envs"])
    types = set(int(t) for t in model["joint_type"][: len(model["joint_type"]) // n])
    assert have == sum(1 << t for t in types), (have, types)
    compiled = lib.dfx_tile_joint_mask(L, D, Q, C, M)
    assert compiled != 31, "no size-specialised tile kernel for %s (%d, %d, %d, %d, %d)" % (name, L, D, Q, C, M)
    assert have & ~compiled == 0, (name, bin(have), bin(compiled))
    assert lib.dfx_tile_joint_mask(5, 5, 5, 5, 5) == 31             # unknown sizes: every type
