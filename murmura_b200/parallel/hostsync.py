"""CPU-only rendezvous of the ranks through the process group's key-value store (no GPU work, no NCCL kernel).

Why it exists: a spin-wait kernel on GPU A (``wait_epoch``, or any NCCL kernel) keeps A busy until GPU B publishes — and while A is
busy a ``cudaMalloc`` on B can block, because with peer mappings enabled a new allocation has to be mapped on the peers too.  If B still
has to allocate before its publish is enqueued (first rounds: workspaces, CUDA-graph capture), B waits for A and A waits for B until
the flag timeout fires and B's nodes are dropped for the round.  During warm-up rounds every rank therefore enqueues its publish FIRST,
meets the others here on the host, and only then launches the wait: whatever a peer still allocates afterwards, the flag this GPU
spins on is already on its way.
"""
from __future__ import annotations

import time


def store_barrier(store, key: str, world: int, timeout_s: float = 30.0, poll_s: float = 2e-4) -> bool:
    """Every caller adds 1 to ``key`` and polls until the counter reaches ``world``.  Returns False on timeout (a dead rank: the
    caller falls through to the device-side flag timeout, which reports it)."""
    store.add(key, 1)
    deadline = time.time() + max(0.0, timeout_s)
    while int(store.add(key, 0)) < world:
        if time.time() >= deadline:
            return False
        time.sleep(poll_s)
    return True
