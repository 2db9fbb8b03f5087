"""Harness-level torch compatibility for the UNCHANGED reference trainers (``algorithms/shac.py``, ``bptt.py``).

The reference was written for torch 1.8: it keeps per-env counters on the CPU (``shac.py:399``
``self.episode_length = torch.zeros(self.num_envs, dtype=int)``) and indexes them with CUDA index tensors
(``shac.py:277-281``), which torch >= 2 rejects ("indices should be either on cpu or on the same device as the indexed
tensor").  Importing this module restores the old behaviour -- a CUDA index applied to a CPU tensor is moved to the CPU
first -- so that the trainer files themselves stay byte-identical.  Test / bench harness only.
"""
import torch

_getitem, _setitem = torch.Tensor.__getitem__, torch.Tensor.__setitem__


def _host_index(t, idx):
    if isinstance(idx, torch.Tensor) and idx.is_cuda and not t.is_cuda:
        return idx.cpu()
    return idx


def _compat_getitem(self, idx):
    return _getitem(self, _host_index(self, idx))


def _compat_setitem(self, idx, value):
    return _setitem(self, _host_index(self, idx), value)


torch.Tensor.__getitem__ = _compat_getitem
torch.Tensor.__setitem__ = _compat_setitem
