"""Pydantic config schema — a superset of the reference's (``murmura/config/schema.py:7-202``).

Every reference block/field/default is kept (so reference YAMLs load unchanged; unknown
top-level keys are rejected, unknown keys inside blocks ignored).  Additions are purely
additive: ``backend: "b200"`` and the optional ``b200:`` block that configures the
Blackwell engine (GPU count, transport, CUDA graphs, compute dtype, sketch precision…).
"""
from __future__ import annotations

from typing import Union, Any, Dict, Literal, Optional

from pydantic import BaseModel, ConfigDict, Field


class DistributedConfig(BaseModel):
    """ZeroMQ wall-clock backend settings (reference ``schema.py:7-51``)."""
    transport: Literal["ipc", "tcp"] = Field(default="ipc", description="ipc (one host) or tcp (multi-host)")
    ipc_dir: str = Field(default="/tmp/murmura", description="base directory of the IPC socket files")
    host: str = Field(default="127.0.0.1", description="coordinator host for tcp transport")
    coordinator_pub_port: int = Field(default=5500, description="kept for schema parity (unused)")
    coordinator_pull_port: int = Field(default=5501, description="monitor PULL port (tcp)")
    base_port: int = Field(default=5550, description="node i listens on base_port + i (tcp)")
    node_hosts: Optional[Dict[int, str]] = Field(default=None, description="per-node host overrides (tcp)")
    round_duration_s: float = Field(default=60.0, description="wall-clock budget of one round")
    startup_grace_s: float = Field(default=5.0, description="delay between launch and round 0")


class ExperimentConfig(BaseModel):
    name: str = Field(description="experiment name")
    seed: int = Field(default=42)
    rounds: int = Field(default=20)
    verbose: bool = Field(default=False)


class TopologyConfig(BaseModel):
    type: Literal["ring", "fully", "erdos", "k-regular"]
    num_nodes: int
    p: Optional[float] = Field(default=None, description="edge probability (erdos)")
    k: Optional[int] = Field(default=None, description="degree (k-regular)")
    seed: int = Field(default=12345)


class AggregationConfig(BaseModel):
    algorithm: Literal["fedavg", "krum", "balance", "sketchguard", "ubar", "evidential_trust"]
    params: Dict[str, Any] = Field(default_factory=dict)


class AttackConfig(BaseModel):
    enabled: bool = False
    type: Optional[Literal["gaussian", "directed_deviation", "topology_liar"]] = None
    percentage: float = 0.0
    params: Dict[str, Any] = Field(default_factory=dict)


class MobilityConfig(BaseModel):
    area_size: float = 100.0
    comm_range: float = 30.0
    max_speed: float = 5.0
    seed: int = 42
    ensure_connected: bool = True


class DMTTConfig(BaseModel):
    """DMTT hyper-parameters (reference ``schema.py:114-139``)."""
    budget_B: int = 5
    rho: float = 0.1
    lambda_forget: float = 0.9
    w_d: float = 1.0
    w_c: float = 0.5
    w_x: float = 1.0
    tau_U: float = 0.3
    eta: float = 5.0
    w_a: float = 0.7
    tau_u: float = 0.5
    lambda1: float = 0.4
    lambda2: float = 0.3
    lambda3: float = 0.2
    lambda4: float = 0.1


class TrainingConfig(BaseModel):
    local_epochs: int = 1
    batch_size: int = 64
    lr: float = 0.01
    max_samples: Optional[int] = None


class DataConfig(BaseModel):
    adapter: str
    params: Dict[str, Any] = Field(default_factory=dict)


class ModelConfig(BaseModel):
    factory: str
    params: Dict[str, Any] = Field(default_factory=dict)


class B200Config(BaseModel):
    """Blackwell engine knobs (new; all optional)."""
    gpus: Optional[int] = Field(default=None, description="GPUs to use; default WORLD_SIZE or 1")
    transport: Literal["auto", "p2p", "nvls", "nccl"] = Field(
        default="auto", description="p2p = in-kernel peer loads; nvls = multimem.ld_reduce (in-switch sum) for full-mesh FedAvg; "
        "auto = nvls whenever the round is a full-mesh FedAvg over several GPUs, p2p otherwise; nccl = baseline: NCCL exchange + "
        "PyTorch aggregation")
    cuda_graphs: bool = Field(default=True, description="capture per-node train/eval steps in CUDA graphs")
    compute_dtype: Literal["fp32", "tf32", "bf16"] = Field(
        default="fp32", description="matmul/conv math mode for local training (params stay fp32)")
    sketch_dtype: Literal["fp32", "fp8"] = Field(
        default="fp32", description="published Count-Sketch precision (fp8 = e4m3 + ue8m0 per 32)")
    krum_gram: Literal["auto", "tcgen05", "fp32"] = Field(
        default="auto", description="Krum distances: fp32 = exact Σ(a−b)² (the reference's definition; auto picks it); tcgen05 = TF32 Gram "
        "on tensor cores with an exact recomputation of every pair whose distance is below krum_refine_tau·(‖a‖²+‖b‖²)")
    krum_refine_tau: float = Field(default=1e-2, description="cancellation guard of the tcgen05 Gram path (relative to ‖a‖²+‖b‖²)")
    channels_last: bool = Field(default=True, description="store 4-D weights / image shards NHWC (tensor-core conv path); "
                                "aggregation is element-wise so the physical order is irrelevant to it")
    split_backward: Union[bool, Literal["auto"]] = Field(
        default="auto", description="compute weight gradients on a side stream (parallel graph branch) so only the data-gradient chain is on "
                                    "the critical path; auto = only when a GPU hosts at most 2 nodes (more streams than hardware queues serialise)")
    fused_bn: bool = Field(default=True, description="fused BatchNorm(+residual)(+ReLU) training kernels (cluster/DSMEM reduction) in the bundled models")
    gather_impl: Literal["auto", "ldg", "tma"] = Field(
        default="auto", description="weighted_gather variant: ldg = 128-bit streaming loads (best when every source is local: L2 reuse); "
                                    "tma = cp.async.bulk ring through shared memory (best over NVLink); auto = tma iff more than one GPU")
    placement: Literal["balanced", "contiguous"] = Field(
        default="balanced", description="virtual-node → GPU map: balanced = longest-shard-first onto the least-loaded GPU")
    unroll_round: bool = Field(default=True, description="capture ALL local steps of a node's round (epochs x batches) in one CUDA "
                               "graph instead of one graph per step (fewer graph launches; matters for tiny models)")
    fused_train: Union[bool, Literal["auto"]] = Field(
        default="auto", description="train all nodes of a GPU with the fused tcgen05 program (parallel/fused_trainer.py: grouped implicit-GEMM "
                                    "conv / linear kernels with the SGD step in the wgrad epilogue, one CUDA graph per round); auto = whenever the "
                                    "model family, loss and layout are supported, otherwise the per-node autograd graphs")
    lazy_metrics: bool = Field(default=True, description="evaluated rounds write their metric table to a pinned host ring and `history` is filled "
                               "lazily (no host synchronisation per round; verbose / profile runs stay round-by-round)")
    seed_parity: bool = Field(default=False, description="draw model initialisation and the per-round shuffles from the SAME host RNG stream as the "
                              "simulation backend / the reference (all N models built in node order from the caller's torch seed, one stock "
                              "DataLoader-style permutation per node and epoch), so attack-free runs can be compared round by round; "
                              "attack noise and dropout stay on the device Philox streams")
    score_tma: bool = Field(default=True, description="foreign-weight scoring (UBAR / EvidentialTrust / DMTT): feed the candidates' weights by TMA from "
                            "their (peer-mapped) arenas, one launch per source GPU; false = cp.async gather through per-group row pointers")
    fullmesh_rank_sum: bool = Field(default=True, description="fully connected FedAvg: exchange one per-rank sum row (publish_sum → fedavg_fullmesh) "
                                    "instead of every node's row; false = the general edge-list gather")
    fullmesh_two_shot: Union[bool, Literal["auto"]] = Field(default="auto", description="several GPUs: all-reduce the per-rank sum rows with the fused reduce-scatter + all-gather "
                                    "kernel (each GPU reduces 1/G of the row and scatters it to every peer); false = every GPU reduces the whole row; "
                                    "auto = two-shot from 4 GPUs (measured: one-shot peer loads win at 2 GPUs)")
    fused_eval_rows: int = Field(default=256, description="fused evaluation: samples per node and launch (nodes are sorted by shard size, so a "
                                 "chunk only runs the nodes that still have samples: no padding to the largest shard)")
    fused_side_stream: bool = Field(default=True, description="fused_train: weight-gradient launches on a parallel graph branch")
    batched_mlp_train: bool = Field(default=False, description="train all MLP-family nodes of a GPU in one batched step (strided-batched GEMMs "
                                    "over arena-row views, masked per-node BatchNorm/SGD); opt-in, falls back to per-node graphs")
    grouped_mlp: bool = Field(default=True, description="score foreign MLP weights (UBAR stage 2 / EvidentialTrust / DMTT) with the "
                              "grouped tcgen05 forward (TF32) instead of per-candidate graph replays (fp32)")
    streams: int = Field(default=0, description="concurrent CUDA streams for virtual-node training (0 = auto: min(16, nodes on this GPU))")
    eval_batch: int = Field(default=1024, description="evaluation micro-batch (results are batch-size independent)")
    host_barrier_rounds: int = Field(default=3, description="rounds after (re)configuration in which the ranks meet on the host (key-value store, no GPU "
                                     "work) between publishing and spinning on the peers' flags: while workspaces / CUDA graphs are still being "
                                     "allocated, a cudaMalloc can block behind a peer GPU's spin-wait kernel")
    flag_timeout_ms: float = Field(default=30000.0, description="device-side wait budget for a peer's publish flag; a rank that stays silent "
                                   "longer is treated as missing for the round (and reported): the reference's deadline-driven partial aggregation")
    fault_drop_edges: Dict[int, list] = Field(
        default_factory=dict, description="fault injection: {round: [[src, dst], ...]} edges to drop")
    checkpoint_every: int = Field(default=0, description="save arena checkpoint every k rounds (0 = off)")
    checkpoint_dir: str = Field(default="checkpoints")
    profile: bool = Field(default=False, description="emit NVTX ranges + per-phase CUDA-event timings")
    stream_inputs: bool = Field(
        default=False, description="keep shards in pinned host memory and copy them H2D every round "
        "(end-to-end mode); default keeps shards resident in HBM")


class Config(BaseModel):
    model_config = ConfigDict(extra="forbid")

    experiment: ExperimentConfig
    topology: TopologyConfig
    aggregation: AggregationConfig
    attack: AttackConfig = Field(default_factory=AttackConfig)
    training: TrainingConfig
    data: DataConfig
    model: ModelConfig
    backend: Literal["simulation", "distributed", "b200"] = Field(default="simulation")
    distributed: DistributedConfig = Field(default_factory=DistributedConfig)
    mobility: Optional[MobilityConfig] = None
    dmtt: Optional[DMTTConfig] = None
    b200: B200Config = Field(default_factory=B200Config)
