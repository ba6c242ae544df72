"""The score-model training step as ONE CUDA graph (`train_step` orchestration
without host syncs; reference `agents/state_agent.py:143-156`, `core/active_inference.py:533-636`).

One replay = zero the gradients -> `compute_diffusion_elbo` forward (two score-net evaluations,
decoder, reward head, all random draws) -> backward, including the gradient penalty's double
backward -> the sequential time-importance EMA -> gradients left in `p.grad`.  Eagerly this is
~1,500 kernel launches (≈300 GEMM calls of 5-6 launches each plus the element-wise glue); at the
per-GPU batches of the data-parallel configuration (4,096-8,192 rows) the step is launch-bound, and
the graph removes every launch gap and all Python/autograd dispatch from it.  There is no host read
inside the step: the loss and the eight metrics stay on the device until the caller asks.

Under torchrun the gradient exchange is part of the step (and of the graph): the trunk's backward
all-reduces each stage's gradients on a side stream while the next stage computes
(train_native._Trunk.backward), and the remaining parameters go out as one flat collective.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch

import torch.distributed as dist

from . import distributed, train_native


class GraphedElboStep:
    """`step = GraphedElboStep(ai, batch_size)`; `loss, metrics = step(obs, rewards, latents)`.

    `ai` is a `DiffusionActiveInference`; `params` defaults to what the reference's score optimizer
    owns (`latent_score_network` + `latent_diffusion`, agents/state_agent.py:248-253).  Gradients
    of every other parameter reached by the loss (decoder, reward head) are zeroed inside the
    graph and dropped, as the reference's `dynamics_optimizer.zero_grad()` does (:225).
    The first call runs `warmup` eager steps on a side stream (this also creates the
    time-importance weights, so the captured step is the reference's steady-state call: importance-
    sampled t), then captures; later calls copy the inputs into the static buffers and replay.
    """

    def __init__(self, ai, batch_size: int, params: Optional[Iterable[torch.nn.Parameter]] = None,
                 warmup: int = 3, allreduce: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedElboStep needs a CUDA device (there is no CPU fallback)")
        self.ai = ai
        self.batch_size = int(batch_size)
        self.params: List[torch.nn.Parameter] = list(params) if params is not None else (
            list(ai.latent_score_network.parameters()) + list(ai.latent_diffusion.parameters()))
        self._owned = {id(p) for p in self.params}
        self._others = [p for p in ai.parameters() if id(p) not in self._owned]
        # decoder / reward-head gradients are discarded below: when none of their parameters is owned,
        # their loss terms are evaluated without recording a graph (no wasted backward GEMMs)
        side = list(ai.observation_decoder.parameters()) + list(ai.reward_predictor.parameters())
        self.score_only = not any(id(p) in self._owned for p in side)
        self.warmup = max(1, int(warmup))
        self.allreduce = allreduce
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        dev = ai.device
        L = ai.latent_dim
        O = ai.observation_dim
        self.obs = torch.zeros(self.batch_size, O, device=dev)
        self.rewards = torch.zeros(self.batch_size, device=dev)
        self.latents = torch.zeros(self.batch_size, L, device=dev)
        self.loss: Optional[torch.Tensor] = None
        self.metrics: Optional[torch.Tensor] = None
        # data parallel (one process per GPU): the trunk's backward averages its own gradients stage by
        # stage, overlapped with the following stages (train_native); the remaining parameters
        # (conditioning path, diffusion parameters) accumulate into one flat buffer that is exchanged as
        # a single collective after backward.  All of it is enqueued inside the step, i.e. inside the
        # captured graph.
        self.group = dist.group.WORLD if (allreduce and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.flat_rest: Optional[distributed.FlatGrads] = None
        if self.group is not None:
            native = ai._native_training()
            done = {id(p) for p in train_native.reduced_parameters(ai.latent_score_network)} if native else set()
            self._reduced = [p for p in self.params if id(p) in done]
            self.flat_rest = distributed.FlatGrads([p for p in self.params if id(p) not in done])
            ai.data_parallel_group = self.group

    def _eager(self) -> Tuple[torch.Tensor, torch.Tensor]:
        for p in self.params:
            p.grad = None
        for p in self._others:
            p.grad = None
        if self.flat_rest is not None:
            self.flat_rest.zero_and_attach()
            train_native.DATA_PARALLEL_GROUP = self.group if self.ai._native_training() else None
        prev = getattr(self.ai, "elbo_score_only", False)
        self.ai.elbo_score_only = self.score_only
        try:
            loss, vals = self.ai.elbo_device(self.obs, self.rewards, self.latents)
        finally:
            self.ai.elbo_score_only = prev
        try:
            loss.backward()
        finally:
            train_native.DATA_PARALLEL_GROUP = None
        if self.flat_rest is not None:
            self.flat_rest.allreduce(self.group)
        self.ai._join_time_importance()   # the EMA ran beside the backward on its side stream
        for p in self._others:        # decoder / reward-head gradients are discarded (reference :225)
            p.grad = None
        return loss.detach(), vals

    def _capture(self) -> None:
        side = torch.cuda.Stream(device=self.ai.device)
        side.wait_stream(torch.cuda.current_stream(self.ai.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._eager()
        torch.cuda.current_stream(self.ai.device).wait_stream(side)
        torch.cuda.synchronize(self.ai.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.metrics = self._eager()

    def __call__(self, observations: torch.Tensor, rewards: torch.Tensor, latents: torch.Tensor
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (loss, metrics[8]) as device tensors owned by the graph (overwritten by the next
        call); `.grad` of `params` holds this step's (all-reduced) gradients."""
        if observations.shape[0] != self.batch_size:
            raise ValueError(f"GraphedElboStep was built for batch {self.batch_size}, got {observations.shape[0]}")
        self.obs.copy_(observations, non_blocking=True)
        self.rewards.copy_(rewards.reshape(-1), non_blocking=True)
        self.latents.copy_(latents, non_blocking=True)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        return self.loss, self.metrics

    def metrics_dict(self) -> Dict[str, float]:
        """One device->host read of the last step's metrics (the reference's info dict, :625-636)."""
        vals = self.metrics.cpu()
        return {k: float(v) for k, v in zip(self.ai.ELBO_KEYS, vals)}
