"""Default device of the simulation / ZeroMQ backends: first available of cuda → mps → cpu (reference ``murmura/utils/device.py:6-17``).
The B200 engine does not use it — there every rank owns ``cuda:LOCAL_RANK``."""
import torch


def get_device() -> torch.device:
    mps = getattr(torch.backends, "mps", None)
    candidates = (("cuda", torch.cuda.is_available()), ("mps", bool(mps is not None and mps.is_available())))
    return torch.device(next((kind for kind, usable in candidates if usable), "cpu"))
