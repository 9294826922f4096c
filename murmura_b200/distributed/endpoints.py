"""ZeroMQ endpoint naming (parity: reference ``murmura/distributed/endpoints.py:19-69``)."""
from __future__ import annotations

import os

from murmura_b200.config.schema import DistributedConfig


class Endpoints:
    def __init__(self, dist_cfg: DistributedConfig, num_nodes: int, run_id: str):
        self.cfg, self.num_nodes, self.run_id = dist_cfg, num_nodes, run_id

    def _ipc(self, leaf: str) -> str:
        return f"ipc://{self.cfg.ipc_dir}/{self.run_id}/{leaf}"

    @staticmethod
    def _bind_host(own: str) -> str:
        """PULL sockets bind to the node's own configured host, not to every interface (payloads are pickles, as in the reference);
        ``MURMURA_BIND_HOST=0.0.0.0`` restores the reference's bind-all behaviour on a trusted network."""
        return os.environ.get("MURMURA_BIND_HOST") or own

    def monitor_pull_bind(self) -> str:
        return self._ipc("monitor_pull") if self.cfg.transport == "ipc" else f"tcp://{self._bind_host(self.cfg.host)}:{self.cfg.coordinator_pull_port}"

    def monitor_pull_connect(self) -> str:
        return self._ipc("monitor_pull") if self.cfg.transport == "ipc" else f"tcp://{self.cfg.host}:{self.cfg.coordinator_pull_port}"

    def node_pull_bind(self, node_id: int) -> str:
        return self._ipc(f"node_{node_id}") if self.cfg.transport == "ipc" else f"tcp://{self._bind_host((self.cfg.node_hosts or {}).get(node_id, self.cfg.host))}:{self.cfg.base_port + node_id}"

    def node_pull_connect(self, node_id: int) -> str:
        if self.cfg.transport == "ipc":
            return self._ipc(f"node_{node_id}")
        host = (self.cfg.node_hosts or {}).get(node_id, self.cfg.host)
        return f"tcp://{host}:{self.cfg.base_port + node_id}"

    def ensure_dirs(self) -> None:
        if self.cfg.transport == "ipc":
            os.makedirs(f"{self.cfg.ipc_dir}/{self.run_id}", exist_ok=True)
