"""``tensorboardX`` is absent from this image; algorithms/shac.py only needs SummaryWriter."""
try:
    from torch.utils.tensorboard import SummaryWriter  # noqa: F401
except Exception:  # tensorboard itself missing -> inert writer
    class SummaryWriter:  # type: ignore
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def flush(self):
            pass

        def close(self):
            pass
