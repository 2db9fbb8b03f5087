

def set_tape_dtype(dtype="fp32"):
    """Storage of the adjoint tape for packs created FROM NOW ON (tile kernels): "fp32" (default) or "bf16" -- link
    velocities, bias accelerations and link wrenches (18 L floats of every row) as bf16, everything that carries
    positions and all arithmetic in fp32 (include/dfx.h dfx_set_tape_dtype; tolerance: tests/tolerances.py)."""
    from . import _capi
    if dtype not in ("fp32", "bf16"):
        raise ValueError("tape dtype must be 'fp32' or 'bf16'")
    _capi.check(_capi.lib().dfx_set_tape_dtype(1 if dtype == "bf16" else 0), "dfx_set_tape_dtype")
