"""CUDA-graph capture of the per-minibatch train step (the default on CUDA; ``agent.cuda_graphs = False`` or
``CUSRL_B200_CUDA_GRAPHS=0`` selects the eager step).

Validated on a B200 in round 2 (tests/test_graphs_gpu.py: replay == eager training over several iterations for the MLP,
RND and LSTM agents, learning-rate changes between replays, re-capture on a schedule change).

Why.  One train step of the MLP preset is ~60 kernel launches of this library plus ~100 small PyTorch launches, issued
from Python at ~10-20 us each.  At 65536 environments on one GPU the kernels are long enough to hide that; when the
environments are split over 8 ranks (8192 per rank, BASELINE.json's scale-out configuration) or the rollout is small
(4096 environments: 41.7 ms per iteration measured against ~12 ms of kernel time) the host is the bottleneck.  The step
has static shapes and static buffers -- the sampler gathers every minibatch into the same destination tensors -- so
it is captured once and replayed.

What is captured.  ``ActorCritic._train_step`` is split at the one point that must stay outside a graph to keep the
multi-GPU path identical to the eager one, the gradient allreduce:

    graph A : hook.pre_objective -> hook.objective -> loss = sum(objectives) -> zero_grad -> backward
    eager   : distributed.reduce_gradients(optimizer)          (one in-place NCCL allreduce; no-op on one rank)
    graph B : hook.pre_optim (clip) -> optimizer.step -> hook.post_optim -> record(objectives) -> hook.post_objective

Host-side state that changes from step to step and would otherwise be frozen into kernel arguments:

* Adam's step count and learning rate  -> device memory (``FlatAdam.use_device_scalars``, ``adam_step_dev`` kernel);
* the running means of the metrics (their weights ``count / total`` are host numbers) -> recorded in deferred mode
  during capture (``Metrics.begin_deferred``) and merged eagerly after every replay from the graph's static outputs;
* hook hyper-parameters a schedule may change (``register_mutable`` attributes), optimizer betas / eps / weight decay,
  the GEMM precision -> part of the cache key: a change discards the graphs and re-captures.

The first ``WARMUP`` steps of a given (batch tensors, hyper-parameters) key run eagerly (library one-time setup, cuBLAS /
allocator warm-up), the next one is captured (capture executes nothing) and immediately replayed.
"""

from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import torch

from .. import distributed, ops

__all__ = ["TrainStepGraphs"]


def _flat_leaves(tree: Any, prefix: str = ""):
    if isinstance(tree, Mapping):
        for key, value in tree.items():
            yield from _flat_leaves(value, f"{prefix}{key}.")
    elif isinstance(tree, (tuple, list)):
        for i, value in enumerate(tree):
            yield from _flat_leaves(value, f"{prefix}{i}.")
    elif isinstance(tree, torch.Tensor):
        yield prefix[:-1], tree


class TrainStepGraphs:
    WARMUP = 2

    def __init__(self, agent):
        self.agent = agent
        self._entries: dict[tuple, dict[str, Any]] = {}
        self._hyper: tuple | None = None
        self.replays = 0
        self.captures = 0
        self._side_stream = None

    # ---- cache keys --------------------------------------------------------------------------------------------------
    def _hyper_key(self) -> tuple:
        agent = self.agent
        hooks = tuple((hook.name, tuple((m, repr(getattr(hook, m, None))) for m in sorted(getattr(hook, "_mutable", ()))))
                      for hook in agent.hook.active_hooks())
        groups = tuple((tuple(g["betas"]), g["eps"], g["weight_decay"]) for g in agent.optimizer.param_groups)
        return hooks, groups, ops.GEMM_PRECISION, distributed.world_size()

    @staticmethod
    def _batch_key(metadata: Mapping[str, Any], batch: Mapping[str, Any]) -> tuple:
        leaves = tuple((name, t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype) for name, t in _flat_leaves(batch))
        return leaves, bool(metadata.get("temporal", False))

    # ---- the step ----------------------------------------------------------------------------------------------------
    def __call__(self, metadata: dict[str, Any], batch: dict[str, Any]) -> None:
        agent = self.agent
        hyper = self._hyper_key()
        if hyper != self._hyper:
            self._entries.clear()  # a schedule changed a captured constant: capture again
            self._hyper = hyper
        key = self._batch_key(metadata, batch)
        entry = self._entries.setdefault(key, {"seen": 0})
        if entry.get("eager_only"):
            agent._train_step_eager(metadata, batch)
            return
        if entry["seen"] < self.WARMUP:
            entry["seen"] += 1
            self._warmup_step(metadata, batch)
            return
        if "graph_a" not in entry:
            if not self._capture(entry, metadata, batch):
                agent._train_step_eager(metadata, batch)
                return
        self._replay(entry, batch)

    def _warmup_step(self, metadata: dict[str, Any], batch: dict[str, Any]) -> None:
        """An eager step on a SIDE stream.  Autograd's AccumulateGrad node of a parameter remembers the stream it was
        created on and is shared by every graph that is alive at the same time; a node born on the legacy default stream
        makes the backward pass inside a capture synchronise the legacy stream with the capturing one, which CUDA refuses
        (cudaErrorStreamCaptureImplicit -- observed on a B200).  Warming up off the default stream is the documented
        recipe for whole-network capture."""
        if self.agent.device.type != "cuda":  # the CPU control-flow harness (tools/host_overhead_cpu.py)
            self.agent._train_step_eager(metadata, batch)
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        current = torch.cuda.current_stream()
        self._side_stream.wait_stream(current)
        with torch.cuda.stream(self._side_stream):
            self.agent._train_step_eager(metadata, batch)
        current.wait_stream(self._side_stream)

    def _capture(self, entry: dict[str, Any], metadata: dict[str, Any], batch: dict[str, Any]) -> bool:
        agent, optimizer = self.agent, self.agent.optimizer
        agent.last_objectives = None
        optimizer.use_device_scalars()
        ops.invalidate_weight_cache()  # the operand copies of the weights must be rebuilt INSIDE the graph
        keys_before = set(batch)
        step_before = optimizer.step_count
        torch.cuda.synchronize()
        graph_a, graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        agent.metrics.begin_deferred()
        try:
            with torch.cuda.graph(graph_a):
                objectives = agent._train_step_forward_backward(metadata, batch)
            if objectives is None:
                # nothing to optimise for this batch layout: keep it eager for good (capture executed nothing)
                agent.metrics.end_deferred()
                entry["eager_only"] = True
                for name in set(batch) - keys_before:
                    del batch[name]
                return False
            with torch.cuda.graph(graph_b, pool=graph_a.pool()):
                agent._train_step_optimize(metadata, batch, objectives)
        except BaseException:
            agent.metrics.end_deferred()
            ops.invalidate_weight_cache()
            raise
        finally:
            # the host-side weight-operand cache recorded "fresh" copies that no kernel has produced yet
            ops.invalidate_weight_cache()
        entry["deferred"] = agent.metrics.end_deferred()
        entry["outputs"] = {name: batch[name] for name in batch if name not in keys_before}
        entry["objectives"] = {name: value.detach() for name, value in objectives.items()}
        entry["graph_a"], entry["graph_b"] = graph_a, graph_b
        # the capture ran the host side of a step but no kernel: undo the host bookkeeping, the replay redoes it
        optimizer.step_count = step_before
        self.captures += 1
        return True

    def _replay(self, entry: dict[str, Any], batch: dict[str, Any]) -> None:
        agent, optimizer = self.agent, self.agent.optimizer
        optimizer.sync_device_scalars()
        agent.actor.clear_intermediate_repr()
        agent.critic.clear_intermediate_repr()
        entry["graph_a"].replay()
        distributed.reduce_gradients(optimizer)
        entry["graph_b"].replay()
        # host-side effects of the step
        optimizer.step_count += 1
        optimizer._clip_pending = False
        ops.invalidate_weight_cache()
        agent.metrics.apply(entry["deferred"])
        batch.update(entry["outputs"])
        agent.last_objectives = entry["objectives"]
        self.replays += 1
