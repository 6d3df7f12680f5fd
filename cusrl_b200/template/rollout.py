"""Fused rollout step: ``ActorCritic.act`` / ``ActorCritic.step`` for the standard
feed-forward PPO agent, writing every transition leaf STRAIGHT into its time-major buffer slot.

The generic path (template/actor_critic.py, the reference's control flow actor_critic.py:227-291) costs, per environment
step, two full copies of the observation and of the next observation (``to_tensor`` clone + ``Buffer.push``), ~40 small
elementwise launches for sampling / log-probability / bookkeeping and 14 indexed copies.  Here a step is:

  act :  observation -> its 16-byte-padded slot (one copy kernel; for host inputs one H2D into a staging tensor first, or
         NOTHING when the array is the very object that ``step`` received as ``next_observation`` one call earlier: the
         slot of the previous step already holds it on the device);
         actor trunk (3 tcgen05 GEMMs) -> mean head -> ``action_dist.mean`` slot;  critic trunk -> value head -> ``value`` slot;
         ONE draw of standard-normal noise from torch's generator (the draw Normal.rsample makes, distribution.py:203) and
         ONE kernel for std / action / action_logp into their slots;
  step:  ONE kernel for next_observation (+ next_state), reward, terminated, truncated, done into their slots.

It applies when nothing on the agent needs the generic flow: feed-forward ``Mlp`` actor and critic with SIMT-sized heads,
no hook other than ``ValueComputation`` overriding ``pre_act`` / ``post_act`` / ``post_step``, training mode, fp32 inputs
of the declared shapes.  Anything else (recurrent nets, observation normalisation, user hooks, extra transition fields,
numpy inputs) takes the generic path -- call by call, both write the same storage.  The first step of a run always goes
through the generic path: its ``Buffer.push`` is what allocates the leaves.
"""

from __future__ import annotations

import os
from typing import Any

import torch

from .. import _lib, distributed, ops
from ..nn import functional as F
from ..nn import modules as M
from .hook import Hook

__all__ = ["FusedRecurrentRollout", "FusedRollout", "make_fused_rollout"]

_ROLLOUT_CALLBACKS = ("pre_act", "post_act", "post_step", "should_update")


def _overrides(hook: Any, name: str) -> bool:
    fn = getattr(type(hook), name, None)
    return fn is not None and fn is not getattr(Hook, name)


def _fingerprint(x: torch.Tensor) -> tuple:
    return (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, x.device, x._version)


def _same_array(x: torch.Tensor, prev: tuple) -> bool:
    """`x` is the array remembered in `prev` = (tensor kept alive, its fingerprint at that time): same memory, layout and
    version counter, i.e. the Trainer's ``observation = next_observation`` hand-over (reference trainer.py:313) or an equal
    view of the same storage.  The remembered tensor is kept alive, so its memory cannot have been recycled."""
    return isinstance(x, torch.Tensor) and _fingerprint(x) == prev[1] == _fingerprint(prev[0])


class _Net:
    """Forward-only launch plan of one trunk + head with persistent activation buffers.

    On the 3xTF32 path (fewer rows than ``ops.F16X3_MIN_ROWS``: the regime in which a rollout step is bound by the time the
    host needs to ISSUE its ~11 launches, not by the GPU) the launches are *bound*: every argument that does not depend on
    the buffer step -- operand copies of the weights, biases, the activation buffers, shapes -- is resolved to a plain
    integer once, and a step passes only the input rows, the output slot and the stream.  The binding is re-made when a
    parameter moved or its operand copies were re-allocated; stale operand copies are re-split exactly as before
    (``ops.prepared_weight``), detected by the weights' version counters and the optimizer's epoch."""

    def __init__(self, backbone: M.Mlp, head: torch.nn.Linear, rows: int, device: torch.device):
        self.linears = backbone.linears()
        self.act = F.ACTIVATIONS[backbone.activation]
        self.head = head
        self.rows = rows
        self.acts = [torch.empty(rows, lin.out_features, device=device) for lin in self.linears]
        self._weights = [lin.weight for lin in self.linears]
        self._params = [p for lin in self.linears for p in (lin.weight, lin.bias) if p is not None]
        self._params += [p for p in (head.weight, head.bias) if p is not None]
        self._f16_ok = F.f16x3_supported(self._weights, [lin.bias for lin in self.linears])
        self._stamp = self._ptrs = self._wps = None
        self._first = self._rest = self._head_args = None

    def uses_f16x3(self) -> bool:
        return ops.GEMM_PRECISION == 2 and self.rows >= ops.F16X3_MIN_ROWS and self._f16_ok

    def forward(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """f16x3 trunk (large row counts: the step is bound by the GPU)."""
        pairs = [(lin.weight, lin.bias) for lin in self.linears]
        _, acts = F.f16_trunk_forward(x, pairs, self.act, True)
        self.acts[-1] = acts[-1]
        return ops.head_fwd(acts[-1], self.head.weight, self.head.bias, out=out)

    # ---- bound 3xTF32 launches ------------------------------------------------------------------------------------------------
    def _bind(self, wps) -> None:
        lib = _lib.load()
        self._linear_fn, self._head_fn = lib.cusrl_b200_linear_fwd_tf32, lib.cusrl_b200_head_fwd_f32
        precision = ops.tf32_passes()
        f32 = torch.float32
        calls = []
        src = None
        for lin, wp, buf in zip(self.linears, wps, self.acts):
            yp, ldy = ops._rows(buf, "y")
            tail = (wp["hi"].data_ptr(), wp["lo"].data_ptr(), wp["hi"].stride(0), ops._ptr(lin.bias, f32, "bias"), yp, ldy,
                    self.rows, lin.out_features, lin.in_features, self.act, precision)
            calls.append(tail if src is None else src + tail)
            src = (yp, ldy)
        self._first, self._rest = calls[0], calls[1:]
        head = self.head
        No, K = head.weight.shape
        self._head_args = (src[0], src[1], ops._ptr(head.weight.detach(), f32, "weight"),
                           ops._ptr(None if head.bias is None else head.bias.detach(), f32, "bias"))
        self._head_tail = (self.rows, K, No)
        self._wps = wps

    def refresh(self) -> None:
        """Re-split stale operand copies (on the CURRENT stream) and re-bind the launches if anything moved."""
        stamp = (ops._weights_epoch, *[w._version for w in self._weights])
        ptrs = [p.data_ptr() for p in self._params]
        if stamp != self._stamp or ptrs != self._ptrs:
            wps = [ops.prepared_weight(w) for w in self._weights]   # re-splits the stale ones
            if ptrs != self._ptrs or self._wps is None or any(a is not b for a, b in zip(wps, self._wps)):
                self._bind(wps)
            self._stamp, self._ptrs = stamp, ptrs

    def launch(self, x_ptr: int, ldx: int, out_ptr: int, stream: int, refresh: bool = True) -> None:
        """Trunk + head into the dense ``[rows, No]`` slot at `out_ptr`, on `stream`; `x_ptr` / `ldx`: fp32 input rows
        (validated by the caller once per slot).  `refresh=False`: the caller has called :meth:`refresh` itself (it must run on
        the stream that owns the operand copies, which need not be `stream`)."""
        if refresh:
            self.refresh()
        fn = self._linear_fn
        code = fn(x_ptr, ldx, *self._first, stream)
        if code:
            _lib.check(code, "linear_fwd")
        for args in self._rest:
            code = fn(*args, stream)
            if code:
                _lib.check(code, "linear_fwd")
        code = self._head_fn(*self._head_args, out_ptr, *self._head_tail, stream)
        if code:
            _lib.check(code, "head_fwd")
        _lib.KERNEL_LAUNCHES += 2 + len(self._rest)


class _StepSlots:
    """The slots of ONE buffer step with their addresses resolved (views kept alive: they also go into the transition)."""

    __slots__ = ("t", "views", "obs", "state", "mean", "std", "action", "logp", "value",
                 "store_obs", "store_state", "store_tail", "transition_act", "transition_step")


class FusedRollout:
    REQUIRE_CUDA = True   # tools/host_overhead_cpu.py (kernels stubbed out) clears it to drive this control flow on the CPU

    _NARROW = ("action_dist.mean", "action_dist.std", "action", "action_logp", "value", "reward", "terminated", "truncated", "done")

    def __init__(self, agent):
        self.agent = agent
        self.enabled = self._supported()
        self._nets: tuple[_Net, _Net] | None = None
        self._stage: dict[str, torch.Tensor] = {}
        self._eps: torch.Tensor | None = None
        self._prev_next_obs: Any = None      # (tensor, fingerprint) step() received last as next_observation
        self._prev_next_state: Any = None
        self._acted_fast = False
        self.fast_steps = 0
        # CUSRL_B200_ROLLOUT_STREAMS=1: actor and critic of a step on one stream (the default puts the critic on a second one
        # when the step is latency-bound, see act).  Single-process runs only: that is the configuration the second stream was
        # measured and validated in (profiles/r02_rollout_host.md); a multi-rank job keeps the one-stream step unless the knob
        # is set explicitly.
        knob = os.environ.get("CUSRL_B200_ROLLOUT_STREAMS")
        self._two_streams = agent.device.type == "cuda" and (knob == "2" or (knob is None and not distributed.enabled()))
        self._streams: Any = None
        self._layout_key: Any = None         # (buffer.layout_version, number of leaves) the cached decisions below belong to
        self._layout_ok = False
        self._slots: dict[int, _StepSlots] = {}

    # ---- applicability ---------------------------------------------------------------------------------------------------
    def _supported(self) -> bool:
        agent = self.agent
        if agent.device.type != "cuda" and self.REQUIRE_CUDA:
            return False
        actor, critic = agent.actor, agent.critic
        if not (type(actor) is M.Actor and type(critic) is M.Value):   # wrappers (SymmetricActor) run their own forward
            return False
        if not (isinstance(actor.backbone, M.Mlp) and isinstance(critic.backbone, M.Mlp)):
            return False
        if not (actor.backbone.ends_with_activation and critic.backbone.ends_with_activation):
            return False
        if not isinstance(actor.distribution, M.NormalDist):
            return False
        if not (F.simt_head_supported(*actor.distribution.mean_head.weight.shape)
                and F.simt_head_supported(*critic.value_head.weight.shape)):
            return False
        from ..hook.on_policy import ValueComputation

        for hook in agent.hook:
            if isinstance(hook, ValueComputation):
                continue
            if any(_overrides(hook, name) for name in _ROLLOUT_CALLBACKS):
                return False
        return any(isinstance(h, ValueComputation) for h in agent.hook)

    def _needed_leaves(self) -> list[str]:
        need = ["observation", "action_dist.mean", "action_dist.std", "action", "action_logp", "value", "next_observation",
                "reward", "terminated", "truncated", "done"]
        if self.agent.has_state:
            need += ["state", "next_state"]
        return need

    def _layout_supported(self) -> bool:
        storage = self.agent.buffer.storage
        need = self._needed_leaves()
        if not all(k in storage for k in need):
            return False  # not allocated yet: the generic path's first push does that
        extra = set(storage) - set(need) - {"next_value", "advantage", "return"}
        if extra:
            return False  # leaves this path does not know how to fill (user transition fields)
        return all(storage[k].is_contiguous() for k in self._NARROW)   # their kernels write dense rows

    def _sync_layout(self) -> bool:
        """Decisions and resolved slot addresses are cached per buffer layout (leaves are allocated by the first pushes and
        then stay put); any allocation / removal of a leaf drops them."""
        buffer = self.agent.buffer
        key = (buffer.layout_version, len(buffer.storage))
        if key != self._layout_key:
            self._layout_key = key
            self._slots = {}
            self._layout_ok = self._layout_supported()
        return self._layout_ok

    def _ready(self) -> bool:
        return self.enabled and not self.agent.inference_mode and self._sync_layout()

    def _input_ok(self, x, width: int, dtype=torch.float32) -> bool:
        return (isinstance(x, torch.Tensor) and x.dtype == dtype and x.dim() == 2 and x.shape[0] == self.agent.parallelism
                and x.shape[1] == width and x.stride(1) == 1 and (x.is_cuda or x.device.type == "cpu"))

    # ---- helpers -----------------------------------------------------------------------------------------------------------
    def _device_rows(self, x: torch.Tensor, key: str) -> torch.Tensor:
        """`x` on the device: itself, or its asynchronous H2D copy in a persistent staging tensor."""
        if x.is_cuda:
            return x
        stage = self._stage.get(key)
        if stage is None or stage.shape != x.shape or stage.dtype != x.dtype:
            stage = torch.empty(x.shape, dtype=x.dtype, device=self.agent.device)
            self._stage[key] = stage
        stage.copy_(x, non_blocking=True)
        return stage

    def _slot(self, key: str, t: int) -> torch.Tensor:
        return self.agent.buffer.storage[key][t]

    def _stream_objects(self):
        """(fork event, join event, side stream) of the two-stream act, created on first use."""
        if self._streams is None:
            device = self.agent.device
            with torch.cuda.device(device):
                self._streams = (torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Stream(device=device))
        return self._streams

    def _step_slots(self, t: int) -> _StepSlots:
        """Views and addresses of buffer step `t`, resolved once per buffer layout (was: ~25 ``storage[key][t]`` views and as
        many pointer / stride / dtype look-ups per environment step)."""
        slots = self._slots.get(t)
        if slots is not None:
            return slots
        agent = self.agent
        storage = agent.buffer.storage
        f32 = torch.float32
        s = _StepSlots()
        v = s.views = {k: storage[k][t] for k in self._needed_leaves()}
        s.t = t
        s.obs = ops._rows(v["observation"], "observation slot")
        s.state = ops._rows(v["state"], "state slot") if agent.has_state else None
        s.mean, s.std, s.action, s.logp, s.value = (ops._ptr(v[k], f32, k + " slot") for k in (
            "action_dist.mean", "action_dist.std", "action", "action_logp", "value"))
        s.store_obs = (*ops._rows(v["next_observation"], "next_observation slot"), agent.observation_dim)
        s.store_state = ((*ops._rows(v["next_state"], "next_state slot"), agent.state_dim) if agent.has_state
                         else (None, 0, 0))
        s.store_tail = (ops._flag_ptr(v["terminated"], "terminated slot"), ops._flag_ptr(v["truncated"], "truncated slot"),
                        ops._flag_ptr(v["done"], "done slot"), agent.parallelism)
        s.transition_act = {"observation": v["observation"], "action_dist": {"mean": v["action_dist.mean"], "std": v["action_dist.std"]},
                            "action": v["action"], "action_logp": v["action_logp"], "value": v["value"]}
        if agent.has_state:
            s.transition_act["state"] = v["state"]
        s.transition_step = {k: v[k] for k in ("next_observation", "reward", "terminated", "truncated", "done")}
        if agent.has_state:
            s.transition_step["next_state"] = v["next_state"]
        self._slots[t] = s
        return s

    def _fill_wide(self, key: str, t: int, value: torch.Tensor, prev_obj: Any, prev_key: str) -> torch.Tensor:
        slots = self._step_slots(t)
        wide_state = key == "state"
        dst = slots.state if wide_state else slots.obs
        if prev_obj is not None and prev_obj[2] is self._slots.get(prev_obj[2].t) and _same_array(value, prev_obj):
            # the array the caller passed to step() one call ago: already on the device in the slot that step wrote (the slots
            # object must still be the current one of its step: a re-allocated buffer starts empty)
            held = prev_obj[2]
            src = held.store_state[:2] if wide_state else held.store_obs[:2]
        else:
            rows = self._device_rows(value, key)
            src = (rows.data_ptr(), rows.stride(0))
        code = _lib.load().cusrl_b200_copy_rows_padded_f32(*src, *dst, value.shape[0], value.shape[1], ops._stream())
        _lib.check(code, "copy_rows_padded")
        return slots.views[key]

    # ---- act ------------------------------------------------------------------------------------------------------------------
    def act(self, observation, state):
        """Returns the action, or ``None`` when this call must take the generic path."""
        agent = self.agent
        self._acted_fast = False
        if not self._ready() or not self._input_ok(observation, agent.observation_dim):
            return None
        if agent.has_state != (state is not None) or (state is not None and not self._input_ok(state, agent.state_dim)):
            return None
        if self._nets is None:
            dev, n = agent.device, agent.parallelism
            self._nets = (_Net(agent.actor.backbone, agent.actor.distribution.mean_head, n, dev),
                          _Net(agent.critic.backbone, agent.critic.value_head, n, dev))
        t = agent.buffer.cursor
        slots = self._step_slots(t)
        views = slots.views
        tr = agent.transition
        tr.clear()
        obs_slot = self._fill_wide("observation", t, observation, self._prev_next_obs, "next_observation")
        critic_in, critic_rows = obs_slot, slots.obs
        if state is not None:
            critic_in, critic_rows = self._fill_wide("state", t, state, self._prev_next_state, "next_state"), slots.state
        actor_net, critic_net = self._nets
        mean = views["action_dist.mean"]
        join = None
        if self._two_streams and not critic_net.uses_f16x3():
            # small row counts: a step is a chain of ~11 launch-latency-bound kernels, and the critic's four do not depend on
            # the actor's six -- they run on a second stream, forked after the observation reached its slot and joined before
            # act() returns (every later consumer of the value slot / the critic's activations is ordered after the join)
            critic_net.refresh()   # on the main stream: other consumers of the operand copies live there
            main = torch.cuda.current_stream()
            fork, join, side = self._stream_objects()
            fork.record(main)
            side.wait_event(fork)
            critic_net.launch(*critic_rows, slots.value, side.cuda_stream, refresh=False)
            join.record(side)
        if actor_net.uses_f16x3():
            actor_net.forward(obs_slot, mean)
        else:
            actor_net.launch(*slots.obs, slots.mean, ops._stream())
        agent.actor.intermediate_repr["backbone.output"] = actor_net.acts[-1]
        eps = None if agent.deterministic else M.standard_normal_like(mean)
        code = _lib.load().cusrl_b200_sample_logp_f32(
            slots.mean, ops._ptr(agent.actor.distribution.std.param.detach(), torch.float32, "sigma"),
            ops._ptr(eps, torch.float32, "eps"), mean.shape[0], mean.shape[1], int(agent.deterministic), slots.std, slots.action,
            slots.logp, ops._stream())
        _lib.check(code, "sample_logp")
        if join is not None:
            main.wait_event(join)
        elif critic_net.uses_f16x3():
            critic_net.forward(critic_in, views["value"])
        else:
            critic_net.launch(*critic_rows, slots.value, ops._stream())
        agent.critic.intermediate_repr["backbone.output"] = critic_net.acts[-1]
        tr.update(slots.transition_act)
        tr["action_dist"] = dict(slots.transition_act["action_dist"])   # a private dict per step, like the generic flow
        self._acted_fast = True
        action = views["action"]
        if observation.is_cuda:
            return action.clone()   # the slot is overwritten one rollout later: hand out a private copy
        return action.to(device=observation.device)

    # ---- step -----------------------------------------------------------------------------------------------------------------
    def step(self, next_observation, reward, terminated, truncated, next_state, kwargs) -> bool:
        """True when the transition was stored by the fused kernel (the caller then only advances the counters)."""
        agent = self.agent
        self._prev_next_obs, self._prev_next_state = None, None
        if not self._acted_fast:
            return False
        self._acted_fast = False
        ok = (not any(v is not None for v in kwargs.values()) and self._input_ok(next_observation, agent.observation_dim)
              and self._input_ok(reward, agent.value_dim) and self._input_ok(terminated, 1, torch.bool)
              and self._input_ok(truncated, 1, torch.bool)
              and (agent.has_state == (next_state is not None))
              and (next_state is None or self._input_ok(next_state, agent.state_dim)))
        if not ok or not self._sync_layout():
            # the generic step must find the act outputs in the transition dict: they are the slot views, and push()
            # recognises tensors that already live in their slot
            return False
        t = agent.buffer.cursor
        slots = self._step_slots(t)
        views = slots.views
        obs_rows = self._device_rows(next_observation, "next_observation")
        if next_state is None:
            state_src = (None, 0)
        else:
            state_rows = self._device_rows(next_state, "next_state")
            state_src = (state_rows.data_ptr(), state_rows.stride(0))
        f32 = torch.float32
        code = _lib.load().cusrl_b200_rollout_store_step_f32(
            obs_rows.data_ptr(), obs_rows.stride(0), *slots.store_obs, *state_src, *slots.store_state,
            ops._ptr(self._device_rows(reward.contiguous(), "reward"), f32, "reward"), ops._ptr(views["reward"], f32, "reward slot"),
            reward.shape[-1],
            ops._flag_ptr(self._device_rows(terminated.contiguous(), "terminated"), "terminated"),
            ops._flag_ptr(self._device_rows(truncated.contiguous(), "truncated"), "truncated"), *slots.store_tail, ops._stream())
        _lib.check(code, "rollout_store_step")
        agent.transition.update(slots.transition_step)
        # remembered WITH the slots of the step that holds the copy: the hand-over reads that step's slot, wherever the cursor
        # is by then (wrap-around, or a caller that moved it)
        self._prev_next_obs = (next_observation, _fingerprint(next_observation), slots)
        self._prev_next_state = None if next_state is None else (next_state, _fingerprint(next_state), slots)
        agent.buffer.advance()
        self.fast_steps += 1
        return True


class FusedRecurrentRollout(FusedRollout):
    """The same idea for the recurrent PPO agent (LSTM actor and critic, ``RecurrentPpoAgentFactory``): a rollout step of
    the generic flow costs ~0.9 ms of Python (some 80 small torch operations: memory bookkeeping, distribution arithmetic,
    ~20 leaf copies of ``Buffer.push``) against ~0.25 ms of GPU work, so the rollout of BASELINE.json's config 3 was bound by
    the host.  Here a step is

      act :  observation -> padded slot;  per network: per layer one projection GEMM + one sequence-kernel launch (T = 1,
             nothing saved, the layer's slice of the flat memory read and written in place) -> head -> ``action_dist.mean`` /
             ``value`` slot;  ONE noise draw + ONE kernel for std / action / log-prob;
      step:  ONE kernel for next_observation / reward / flags / done, and ONE kernel that zeroes the memories of finished
             episodes in place (``reset_memory``) while writing the copies the buffer stores: ``actor_memory`` /
             ``critic_memory`` of the NEXT step and ``next_critic_memory`` of this one.

    Applies when actor and critic are the plain ``Actor`` / ``Value`` over ``Rnn`` backbones whose hidden size the sequence
    kernels cover, with a ``NormalDist`` head, once both memories exist and every leaf is allocated (i.e. from the third
    step of a run on); everything else takes the generic path, call by call."""

    def _supported(self) -> bool:
        agent = self.agent
        if agent.device.type != "cuda" and self.REQUIRE_CUDA:
            return False
        actor, critic = agent.actor, agent.critic
        if not (type(actor) is M.Actor and type(critic) is M.Value):
            return False
        from ..nn.recurrent import Rnn

        if not (type(actor.backbone) is Rnn and type(critic.backbone) is Rnn):
            return False
        if not isinstance(actor.distribution, M.NormalDist) or actor.distribution.mean_head.weight.shape[0] > 64:
            return False
        for net, head in ((actor, actor.distribution.mean_head), (critic, critic.value_head)):
            H = net.backbone.rnn.hidden_size
            if not ops.lstm_seq_supported(H):
                return False
            if not (F.simt_head_supported(*head.weight.shape) or head.weight.shape[0] % 4 == 0):
                return False
        from ..hook.on_policy import ValueComputation

        self._value_hook = None
        for hook in agent.hook:
            if isinstance(hook, ValueComputation):
                self._value_hook = hook
                continue
            if any(_overrides(hook, name) for name in _ROLLOUT_CALLBACKS):
                return False
        return self._value_hook is not None

    _MEMORY_LEAVES = ("actor_memory.hidden", "actor_memory.cell", "critic_memory.hidden", "critic_memory.cell",
                      "next_critic_memory.hidden", "next_critic_memory.cell")

    def _needed_leaves(self) -> list[str]:
        return [*super()._needed_leaves(), *self._MEMORY_LEAVES]

    def _ready(self) -> bool:
        agent = self.agent
        if not self.enabled or agent.inference_mode:
            return False
        if agent.actor_memory is None or self._value_hook._critic_memory is None:
            return False
        return self._sync_layout()

    @staticmethod
    def _head(latent: torch.Tensor, head: torch.nn.Linear, out: torch.Tensor) -> torch.Tensor:
        if F.simt_head_supported(*head.weight.shape):
            return ops.head_fwd(latent, head.weight, head.bias, out=out)
        return ops.tc_linear_fwd(latent, ops.prepared_weight(head.weight), head.bias, head.weight.shape[0], 0, ops.tf32_passes(),
                                 out=out)

    def act(self, observation, state):
        agent = self.agent
        self._acted_fast = False
        if not self._ready() or not self._input_ok(observation, agent.observation_dim):
            return None
        if agent.has_state != (state is not None) or (state is not None and not self._input_ok(state, agent.state_dim)):
            return None
        hook = self._value_hook
        actor_mem, critic_mem = agent.actor_memory, hook._critic_memory
        if not all(isinstance(m, dict) and m["hidden"].dim() == 2 and m["hidden"].is_contiguous() and m["cell"].is_contiguous()
                   for m in (actor_mem, critic_mem)):
            return None
        t = agent.buffer.cursor
        tr = agent.transition
        tr.clear()
        obs_slot = self._fill_wide("observation", t, observation, self._prev_next_obs, "next_observation")
        tr["observation"] = obs_slot
        critic_in = obs_slot
        if state is not None:
            critic_in = self._fill_wide("state", t, state, self._prev_next_state, "next_state")
            tr["state"] = critic_in
        if self._slots_hold_memory != t:
            # the memories entering this step are not in their slots yet (first fused step of a rollout: the last step of the
            # previous rollout must not overwrite slot 0 before the update has read it)
            for name, mem in (("actor_memory", actor_mem), ("critic_memory", critic_mem)):
                for leaf in ("hidden", "cell"):
                    self._slot(f"{name}.{leaf}", t).copy_(mem[leaf])
        tr["actor_memory"] = {leaf: self._slot(f"actor_memory.{leaf}", t) for leaf in ("hidden", "cell")}
        tr["critic_memory"] = {leaf: self._slot(f"critic_memory.{leaf}", t) for leaf in ("hidden", "cell")}
        # actor
        latent, next_actor_mem = agent.actor.backbone(obs_slot, actor_mem, sequential=False)
        agent.actor.intermediate_repr["backbone.output"] = latent
        mean = self._head(latent, agent.actor.distribution.mean_head, self._slot("action_dist.mean", t))
        std, action, logp = self._slot("action_dist.std", t), self._slot("action", t), self._slot("action_logp", t)
        eps = None if agent.deterministic else M.standard_normal_like(mean)
        ops.sample_logp(mean, agent.actor.distribution.std.param.detach(), eps, std, action, logp, agent.deterministic)
        tr["action_dist"] = {"mean": mean, "std": std}
        tr["action"], tr["action_logp"] = action, logp
        # critic (ValueComputation.post_act, value.py:42-54)
        latent_c, next_critic_mem = agent.critic.backbone(critic_in, critic_mem, sequential=False)
        agent.critic.intermediate_repr["backbone.output"] = latent_c
        tr["value"] = self._head(latent_c, agent.critic.value_head, self._slot("value", t))
        tr["next_critic_memory"] = next_critic_mem
        agent.actor_memory, hook._critic_memory = next_actor_mem, next_critic_mem
        self._acted_fast = True
        if observation.is_cuda:
            return action.clone()
        return action.to(device=observation.device)

    _slots_hold_memory = -1   # buffer step whose actor_memory / critic_memory slots already hold the memories entering it

    def step(self, next_observation, reward, terminated, truncated, next_state, kwargs) -> bool:
        agent = self.agent
        acted_fast = self._acted_fast
        t = agent.buffer.cursor
        if not super().step(next_observation, reward, terminated, truncated, next_state, kwargs):
            if acted_fast:
                # the generic step finishes this transition (its push copies every leaf, its hooks reset the memories): the
                # slots of the next step are not prepared
                self._slots_hold_memory = -1
            return False
        # super().step stored next_observation / reward / flags / done and advanced the cursor
        T = agent.buffer.capacity
        hook = self._value_hook
        actor_mem, critic_mem = agent.actor_memory, hook._critic_memory
        done = agent.buffer.storage["done"][t]
        nxt = t + 1
        slot = (lambda key: agent.buffer.storage[key][nxt]) if nxt < T else (lambda key: None)
        mems = [actor_mem["hidden"], actor_mem["cell"], critic_mem["hidden"], critic_mem["cell"]]
        dst_a = [slot("actor_memory.hidden"), slot("actor_memory.cell"), slot("critic_memory.hidden"), slot("critic_memory.cell")]
        dst_b = [None, None, agent.buffer.storage["next_critic_memory.hidden"][t], agent.buffer.storage["next_critic_memory.cell"][t]]
        if actor_mem["hidden"].shape == critic_mem["hidden"].shape:
            ops.memory_reset_store(mems, done, dst_a=dst_a, dst_b=dst_b)
        else:   # networks of different sizes: one launch per network
            ops.memory_reset_store(mems[:2], done, dst_a=dst_a[:2], dst_b=dst_b[:2])
            ops.memory_reset_store(mems[2:], done, dst_a=dst_a[2:], dst_b=dst_b[2:])
        self._slots_hold_memory = nxt if nxt < T else -1
        agent.transition["next_critic_memory"] = {leaf: agent.buffer.storage[f"next_critic_memory.{leaf}"][t] for leaf in ("hidden", "cell")}
        return True


def make_fused_rollout(agent) -> FusedRollout:
    """The fused rollout implementation that applies to `agent` (feed-forward or recurrent); a disabled FusedRollout (every
    call takes the generic path) when neither does."""
    fused = FusedRollout(agent)
    if fused.enabled:
        return fused
    recurrent = FusedRecurrentRollout(agent)
    return recurrent if recurrent.enabled else fused
