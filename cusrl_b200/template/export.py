"""Export of a trained policy: ``ActorCritic.export`` with the reference's arguments and
artefacts (cusrl/template/actor_critic.py:332-418, cusrl/nn/layer/export.py:20-230).

The B200 modules run hand-written kernels through a C ABI and cannot be traced, but every one of them OWNS the plain torch
modules that hold its parameters under the reference's names (``backbone.layers.{0,2,4}`` Linear / activation stack or
``backbone.rnn`` nn.LSTM, ``distribution.mean_head``), so the deployed policy -- observation (-> normalisation) -> backbone
-> mean head -> action, i.e. the reference's ``forward_type="act_deterministic"`` -- is rebuilt from those modules as a
plain-torch twin that SHARES the parameters, and that twin is what gets exported:

* with the reference importable (``python -m cusrl export -m cusrl_b200.plugin``) the twin goes through the
  reference's own ``FlowGraph`` exporter: same node / input / output names (``observation``, ``memory_in`` -> ``action``,
  ``memory_out``), same files (``actor.onnx`` / ``actor.pt`` + ``actor_stateless.pt``, ``actor.yml``), same optimisation;
* standalone, ``target_format="jit"`` traces the twin with TorchScript (``actor.pt`` + ``actor.yml``); ONNX needs the
  reference's exporter (and the ``onnx`` package), so it is refused with a clear message instead of approximated.
"""

from __future__ import annotations

import os
from typing import Any

import torch
from torch import Tensor, nn

from ..nn import modules as M
from ..nn.recurrent import Rnn

__all__ = ["PlainActor", "export_agent"]


class _Affine(nn.Module):
    """(x - mean) / std -- the exported form of a running normaliser / the environment's observation normalisation."""

    def __init__(self, mean: Tensor, std: Tensor, clamp: float | None = None, denormalize: bool = False):
        super().__init__()
        self.register_buffer("mean", mean.detach().clone().float())
        self.register_buffer("std", std.detach().clone().float())
        self.clamp, self.denormalize = clamp, denormalize

    def forward(self, input: Tensor) -> Tensor:
        if self.denormalize:
            return input * self.std + self.mean
        out = (input - self.mean) / self.std
        return out if self.clamp is None else out.clamp(-self.clamp, self.clamp)


class PlainActor(nn.Module):
    """Plain-torch twin of a B200 :class:`~cusrl_b200.nn.modules.Actor`: deterministic action (and next memory) from the
    torch modules the B200 actor owns -- same parameters (shared, not copied), same state_dict keys under ``backbone`` /
    ``distribution``.  Call signature of the reference's actor node: ``(observation, memory=None, forward_type=...)``."""

    def __init__(self, actor: M.Actor):
        super().__init__()
        from ..hook.symmetry import SymmetricActor

        if isinstance(actor, SymmetricActor):
            raise ValueError("export of a SymmetricActor is not implemented (export the wrapped actor of a converged policy)")
        self.backbone = actor.backbone
        self.distribution = actor.distribution
        self.is_recurrent = bool(actor.is_recurrent)
        self.intermediate_repr: dict[str, Any] = {}

    def reset_memory(self, memory, done=None) -> None:
        self.backbone.reset_memory(memory, done)

    def forward(self, observation: Tensor, memory=None, forward_type: str = "act_deterministic", **kwargs):
        if forward_type not in ("act_deterministic", "forward"):
            raise ValueError(f"PlainActor implements the deployed (deterministic) policy only, got '{forward_type}'")
        backbone = self.backbone
        if isinstance(backbone, Rnn):
            lstm = backbone.rnn
            L, H = lstm.num_layers, lstm.hidden_size
            lead = observation.shape[:-1]
            x = observation.reshape(-1, observation.shape[-1]).unsqueeze(0) if observation.dim() == 2 else observation
            n = x.shape[1]
            state = None
            if memory is not None:
                to_lnh = lambda m: m.reshape(n, L, H).transpose(0, 1).contiguous()  # noqa: E731   "n (k c) -> k n c"
                state = (to_lnh(memory["hidden"]), to_lnh(memory["cell"]))
            latent, (h_n, c_n) = lstm(x, state)
            latent = latent.reshape(*lead, H)
            to_flat = lambda m: m.transpose(0, 1).reshape(n, L * H)  # noqa: E731          "k n c -> n (k c)"
            memory_out = {"hidden": to_flat(h_n), "cell": to_flat(c_n)}
        else:
            latent, memory_out = backbone.layers(observation), None
        action = self.distribution.mean_head(latent)
        if forward_type == "forward":
            return {"mean": action, "std": self.distribution.std(action)}, memory_out
        return action, memory_out


def _spec_nodes(agent, with_environment_normalization: bool) -> tuple[list, list]:
    """Normalisation the ENVIRONMENT declares (actor_critic.py:353-362,392-401): (nodes before the actor, nodes after it)."""
    spec = agent.environment_spec
    pre, post = [], []
    obs_norm = getattr(spec, "observation_normalization", None)
    if with_environment_normalization and obs_norm is not None:
        pre.append((_Affine(agent.to_tensor(obs_norm[1]), agent.to_tensor(obs_norm[0])), "observation_normalization"))
    act_denorm = getattr(spec, "action_denormalization", None)
    if with_environment_normalization and act_denorm is not None:
        post.append((_Affine(agent.to_tensor(act_denorm[1]), agent.to_tensor(act_denorm[0]), denormalize=True),
                     "action_denormalization"))
    return pre, post


class _ChainGraph:
    """Stand-in for the reference's ``FlowGraph`` when the reference is not importable: records the nodes the hooks' export
    callbacks add.  Only what a linear chain can express is accepted -- a node that maps ``observation`` to ``observation``
    (before the actor) or ``action`` to ``action`` (after it); anything else needs the reference's exporter."""

    def __init__(self):
        self.pre: list = []
        self.post: list = []
        self.actor_added = False

    def add_node(self, module, module_name: str, input_names, output_names, **kwargs) -> None:
        sources = list(input_names.values()) if isinstance(input_names, dict) else list(input_names)
        outputs = [output_names] if isinstance(output_names, str) else list(output_names)
        if sources == ["observation"] and outputs == ["observation"] and not self.actor_added:
            self.pre.append((module, module_name))
        elif sources == ["action"] and outputs == ["action"] and self.actor_added:
            self.post.append((module, module_name))
        else:
            raise RuntimeError(f"export node '{module_name}' ({sources} -> {outputs}) cannot be expressed without the reference's "
                               "FlowGraph exporter: run the export with the reference package importable")


class _Deployed(nn.Module):
    """Standalone chain: normalisation nodes -> actor -> denormalisation."""

    def __init__(self, pre, actor: PlainActor, post):
        super().__init__()
        self.pre = nn.ModuleList([m for m, _ in pre])
        self.actor = actor
        self.post = nn.ModuleList([m for m, _ in post])

    def forward(self, observation: Tensor, hidden: Tensor | None = None, cell: Tensor | None = None):
        for module in self.pre:
            observation = module(observation)
        memory = None if hidden is None else {"hidden": hidden, "cell": cell}
        action, memory_out = self.actor(observation, memory)
        for module in self.post:
            action = module(action)
        if memory_out is None:
            return action
        return action, memory_out["hidden"], memory_out["cell"]


def export_agent(agent, output_dir: str, *, target_format: str = "onnx", with_environment_normalization: bool = True,
                 optimize: bool = True, sequence_len: int = 1, batch_size: int = 1, opset_version: int | None = None,
                 dynamo: bool = False, verbose: bool = True, **kwargs) -> None:
    if target_format not in ("onnx", "jit"):
        raise ValueError(f"Unsupported export format '{target_format}'")
    os.makedirs(output_dir, exist_ok=True)
    actor = PlainActor(agent.actor).eval()
    pre, post = _spec_nodes(agent, with_environment_normalization)
    obs_dim = agent.environment_spec.observation_dim
    inputs: dict[str, Any] = {"observation": torch.zeros(sequence_len, batch_size, obs_dim, device=agent.device)}
    try:
        from cusrl.nn.layer.export import FlowGraph   # the REFERENCE's exporter
    except ImportError:
        FlowGraph = None
    if FlowGraph is not None:
        graph = FlowGraph(graph_name="actor")
        for module, name in pre:
            graph.add_node(module, module_name=name, input_names={"input": "observation"}, output_names="observation",
                           expose_outputs=False)
        agent.hook.pre_export(graph)     # e.g. ObservationNormalization adds its running statistics (observation.py:248-255)
        names_in, names_out = {"observation": "observation"}, ["action"]
        if actor.is_recurrent:
            with torch.no_grad():
                _, init_memory = actor(inputs["observation"])
            actor.reset_memory(init_memory)
            inputs["memory_in"] = init_memory
            names_in["memory"] = "memory_in"
            names_out.append("memory_out")
        graph.add_node(actor, module_name="actor", input_names=names_in, output_names=names_out,
                       extra_kwargs={"forward_type": "act_deterministic"},
                       info={"observation_dim": obs_dim, "action_dim": agent.action_dim, "is_recurrent": actor.is_recurrent},
                       expose_outputs=True)
        agent.hook.post_export(graph)
        for module, name in post:
            graph.add_node(module, module_name=name, input_names={"input": "action"}, output_names="action", expose_outputs=False)
        if target_format == "onnx":
            graph.export_onnx(inputs, output_dir, optimize=optimize, dynamo=dynamo, verbose=verbose, opset_version=opset_version)
        else:
            graph.export_jit(inputs, output_dir, optimize=optimize)
    else:
        if target_format == "onnx":
            raise RuntimeError("ONNX export goes through the reference's exporter (cusrl.nn.layer.export.FlowGraph, needs the "
                               "'onnx' package): run `python -m cusrl export ... -m cusrl_b200.plugin`, or use "
                               "target_format='jit'")
        import yaml

        chain = _ChainGraph()
        agent.hook.pre_export(chain)
        chain.actor_added = True
        agent.hook.post_export(chain)
        deployed = _Deployed(pre + chain.pre, actor, chain.post + post).eval()
        example = [inputs["observation"]]
        if actor.is_recurrent:
            with torch.no_grad():
                _, init_memory = actor(inputs["observation"])
            example += [torch.zeros_like(init_memory["hidden"]), torch.zeros_like(init_memory["cell"])]
        with torch.no_grad():
            traced = torch.jit.trace(deployed, tuple(example), strict=False)
            outputs = deployed(*example)
        torch.jit.save(traced, f"{output_dir}/actor.pt")
        outputs = outputs if isinstance(outputs, tuple) else (outputs,)
        info = {"observation_dim": obs_dim, "action_dim": agent.action_dim, "is_recurrent": actor.is_recurrent,
                "inputs": [{n: list(t.shape)} for n, t in zip(("observation", "memory_in.hidden", "memory_in.cell"), example)],
                "outputs": [{n: list(t.shape)} for n, t in zip(("action", "memory_out.hidden", "memory_out.cell"), outputs)]}
        with open(f"{output_dir}/actor.yml", "w") as f:
            yaml.safe_dump(info, f)
    if verbose:
        print(f"Agent exported to \033[4m{output_dir}\033[0m in '{target_format}' format.")
