"""The plugin ABI: ``Hook`` lifecycle callbacks and ``HookComposite`` dispatch.

Interface-compatible with the reference (cusrl/template/hook.py:21-485): same callback names and
arguments, hook identity = snake_case of the class name (:36), module / stateful / mutable registries
(:74-141), ``state_dict`` round trip (:167-196), composite dispatch in list order skipping inactive or
training-only hooks in inference mode (:480-485).  ``torch.compile`` of the joint objective is not
offered: the B200 path is hand-written kernels, and ``compile != False`` is refused by the agent.
"""

from __future__ import annotations

import itertools
import re
from collections.abc import Iterable, Iterator, Mapping
from typing import Any

import torch
from torch import nn

from .. import distributed
from .buffer import Buffer

__all__ = ["Hook", "HookComposite", "camel_to_snake", "is_hook"]

_MISSING = object()


def camel_to_snake(name: str) -> str:
    name = re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", name).lower()


class Hook:
    """A component executed at fixed points of the agent's lifecycle."""

    agent: Any

    def __init__(self, training_only: bool = False):
        self._modules: dict[str, nn.Module | None] = {}
        self._statefuls: dict[str, Any] = {}
        self._mutable: set[str] = set()
        self._name: str = camel_to_snake(self.__class__.__name__)
        self._active: bool = True
        self._training_only: bool = training_only

    # ---- identity ------------------------------------------------------------------------------
    @property
    def name(self) -> str:
        return self._name

    @property
    def active(self) -> bool:
        return self._active

    @property
    def training_only(self) -> bool:
        return self._training_only

    def name_(self, name: str):
        self._name = name
        return self

    def active_(self, active: bool):
        self._active = active
        return self

    # ---- registries ----------------------------------------------------------------------------
    def register_module(self, name: str, module: nn.Module | None) -> None:
        if module is not None:
            module = self.agent.setup_module(module)
        if name in self._statefuls:
            raise RuntimeError(f"Cannot register module '{name}': a stateful with the same name already exists")
        setattr(self, name, module)
        self._modules[name] = module

    def register_stateful(self, name: str, value: Any) -> None:
        if name in self._modules:
            raise RuntimeError(f"Cannot register stateful '{name}': a module with the same name already exists")
        setattr(self, name, value)
        self._statefuls[name] = value

    def register_mutable(self, name: str, value: Any = _MISSING) -> None:
        if value is not _MISSING:
            setattr(self, name, value)
        self._mutable.add(name)

    def update_attribute(self, name: str, value: Any) -> None:
        if name not in self._mutable:
            raise ValueError(f"Attribute '{name}' is not mutable on hook '{self.name}'")
        setattr(self, name, value)

    # ---- parameters / state ----------------------------------------------------------------------
    def named_parameters(self, prefix: str = "") -> Iterator[tuple[str, nn.Parameter]]:
        if prefix:
            prefix += "."
        for module_name, module in self._modules.items():
            if module is not None:
                yield from module.named_parameters(prefix=f"{prefix}{module_name}")

    def parameters(self):
        for _name, param in self.named_parameters():
            yield param

    def state_dict(self) -> dict[str, Any]:
        return {
            name: obj.state_dict()
            for name, obj in itertools.chain(self._modules.items(), self._statefuls.items())
            if obj is not None
        }

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        keys = set(state_dict.keys())
        for name, obj in itertools.chain(self._modules.items(), self._statefuls.items()):
            if obj is None:
                continue
            if name not in keys:
                self.warn(f"No state_dict entry was found for '{name}'.")
                continue
            keys.discard(name)
            try:
                obj.load_state_dict(state_dict[name])
            except (RuntimeError, ValueError) as error:
                self.warn(f"State dict for '{name}' is incompatible: {error}")
        if keys:
            self.warn(f"Unused state_dict keys: {keys}.")

    def train(self, mode: bool = True) -> None:
        for module in self._modules.values():
            if module is not None and hasattr(module, "train"):
                module.train(mode)

    def eval(self) -> None:
        self.train(False)

    # ---- lifecycle (all no-ops by default) --------------------------------------------------------
    def pre_init(self, agent) -> None:
        self.agent = agent

    def init(self) -> None: ...

    def post_init(self) -> None: ...

    def pre_act(self, transition: dict[str, Any]) -> None: ...

    def post_act(self, transition: dict[str, Any]) -> None: ...

    def post_step(self, transition: dict[str, Any]) -> None: ...

    def should_update(self, transition: dict[str, Any]) -> bool:
        return True

    def pre_update(self, buffer: Buffer) -> None: ...

    def pre_objective(self, metadata: dict[str, Any], batch: dict[str, Any]) -> None: ...

    def objective(self, metadata: dict[str, Any], batch: dict[str, Any]) -> dict[str, torch.Tensor] | None:
        return None

    def pre_optim(self, optimizer) -> None: ...

    def post_optim(self) -> None: ...

    def post_objective(self, metadata: dict[str, Any], batch: dict[str, Any]) -> None: ...

    def post_update(self) -> None: ...

    def apply_schedule(self, iteration: int) -> None: ...

    def pre_export(self, graph) -> None:
        """Before the actor node is added to the export graph (reference template/hook.py:344-349): a hook that transforms
        the observation at rollout time adds the same transformation here (``graph.add_node``)."""

    def post_export(self, graph) -> None:
        """After the actor node was added (template/hook.py:351-356)."""

    @classmethod
    def warn(cls, message: str) -> None:
        if distributed.is_main_process():
            print(f"\033[1;31m{cls.__name__}: {message}\033[0m")


_LIFECYCLE = ("pre_init", "init", "post_init", "pre_act", "post_act", "post_step", "should_update", "pre_update",
              "pre_objective", "objective", "pre_optim", "post_optim", "post_objective", "post_update", "apply_schedule",
              "named_parameters", "state_dict", "load_state_dict", "train", "update_attribute", "active_", "name_")


def is_hook(obj: Any) -> bool:
    """A :class:`Hook`, or any object with the reference's hook protocol -- in particular instances of the REFERENCE's
    own ``cusrl.template.Hook`` subclasses (``HookParameterSchedule`` / ``HookActivationSchedule`` of
    cusrl/hook/control/schedule.py:12-77, ``AdvantageReduction``, user hooks), which address the other hooks by name
    through ``agent.hook[name]`` and must keep working next to the B200 hooks."""
    if isinstance(obj, Hook):
        return True
    return (all(callable(getattr(obj, name, None)) for name in _LIFECYCLE)
            and all(hasattr(obj, attr) for attr in ("name", "active", "training_only")))


class HookComposite(Hook):
    """Runs a list of hooks in order; addressable by hook name."""

    def __init__(self, hooks: Iterable[Hook]):
        super().__init__()
        self._hooks = tuple(hooks)
        self._named_hooks: dict[str, Hook] = {}
        for hook in self._hooks:
            if not is_hook(hook):
                raise TypeError(f"Expected a Hook instance, but got '{type(hook).__name__}'")
            if hook.name in self._named_hooks:
                raise RuntimeError(f"Hook '{hook.name}' already exists")
            self._named_hooks[hook.name] = hook
        self._statefuls.update(self._named_hooks)

    def __getitem__(self, name: str) -> Hook:
        if "." in name:
            head, rest = name.split(".", 1)
            return self._named_hooks[head][rest]
        return self._named_hooks[name]

    def __contains__(self, name: str) -> bool:
        return name in self._named_hooks

    def __iter__(self) -> Iterator[Hook]:
        yield from self._hooks

    def named_parameters(self, prefix: str = ""):
        if prefix and not prefix.endswith("."):
            prefix += "."
        for hook_name, hook in self._named_hooks.items():
            yield from hook.named_parameters(prefix=f"{prefix}{hook_name}")

    def train(self, mode: bool = True) -> None:
        for hook in self:
            hook.train(mode)

    def active_hooks(self) -> Iterator[Hook]:
        for hook in self:
            if hook.active and not (self.agent.inference_mode and hook.training_only):
                yield hook

    def pre_init(self, agent) -> None:
        super().pre_init(agent)
        for hook in self.active_hooks():
            hook.pre_init(agent)

    def init(self) -> None:
        for hook in self.active_hooks():
            hook.init()

    def post_init(self) -> None:
        for hook in self.active_hooks():
            hook.post_init()

    def pre_act(self, transition) -> None:
        for hook in self.active_hooks():
            hook.pre_act(transition)

    def post_act(self, transition) -> None:
        for hook in self.active_hooks():
            hook.post_act(transition)

    def post_step(self, transition) -> None:
        for hook in self.active_hooks():
            hook.post_step(transition)

    def should_update(self, transition) -> bool:
        return all(hook.should_update(transition) for hook in self.active_hooks())

    def pre_update(self, buffer) -> None:
        for hook in self.active_hooks():
            hook.pre_update(buffer)

    def pre_objective(self, metadata, batch) -> None:
        for hook in self.active_hooks():
            hook.pre_objective(metadata, batch)

    def objective(self, metadata, batch) -> dict[str, torch.Tensor] | None:
        objectives: dict[str, torch.Tensor] = {}
        for hook in self.active_hooks():
            if (obj := hook.objective(metadata, batch)) is not None:
                objectives.update(obj)
        return objectives or None

    def pre_optim(self, optimizer) -> None:
        for hook in self.active_hooks():
            hook.pre_optim(optimizer)

    def post_optim(self) -> None:
        for hook in self.active_hooks():
            hook.post_optim()

    def post_objective(self, metadata, batch) -> None:
        for hook in self.active_hooks():
            hook.post_objective(metadata, batch)

    def post_update(self) -> None:
        for hook in self.active_hooks():
            hook.post_update()

    def apply_schedule(self, iteration: int) -> None:
        for hook in self.active_hooks():
            hook.apply_schedule(iteration)

    def pre_export(self, graph) -> None:
        """Every hook, active or not, like the reference (template/hook.py:473-479)."""
        for hook in self:
            if callable(fn := getattr(hook, "pre_export", None)):
                fn(graph)

    def post_export(self, graph) -> None:
        for hook in self:
            if callable(fn := getattr(hook, "post_export", None)):
                fn(graph)
