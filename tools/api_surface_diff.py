"""Audit of the reference-facing surface: for every class this package mirrors, list what the REFERENCE's class has and the
mirror lacks -- public and private attribute names, constructor parameters, constructor defaults -- and the state-dict /
instance-attribute differences of a default PPO agent.  Needs the reference package (baseline/_ref or the build
container's /root/reference, see tools/install_reference.py); CPU only.

    python tools/api_surface_diff.py

What remains in its output is deliberate and listed in DESIGN.md section 1 (options the mirror refuses loudly -- bijectors,
action-aware critics, RNN output projections --, the reference's Trainer / logging / inference wrappers, private helpers)."""
import inspect
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from install_reference import reference_path  # noqa: E402

sys.path[:0] = reference_path()
import cusrl  # noqa: E402

import cusrl_b200 as C  # noqa: E402
import cusrl_b200.metrics  # noqa: E402
import cusrl_b200.nn.rms  # noqa: E402

pairs = []
for name in C.hook.__all__:
    ref = getattr(cusrl.hook, name, None) or getattr(cusrl.hook.auxiliary.symmetry, name, None)
    pairs.append((name, ref, getattr(C.hook, name)))
pairs += [
    ("Buffer", cusrl.template.Buffer, C.Buffer), ("Sampler", cusrl.template.Sampler, C.Sampler), ("Hook", cusrl.template.Hook, C.Hook),
    ("HookComposite", cusrl.template.hook.HookComposite, C.HookComposite), ("ActorCritic", cusrl.template.ActorCritic, C.ActorCritic),
    ("ActorCriticFactory", cusrl.template.ActorCritic.Factory, C.ActorCriticFactory), ("MiniBatchSampler", cusrl.MiniBatchSampler, C.MiniBatchSampler),
    ("TemporalMiniBatchSampler", cusrl.TemporalMiniBatchSampler, C.TemporalMiniBatchSampler),
    ("AutoMiniBatchSampler", cusrl.AutoMiniBatchSampler, C.AutoMiniBatchSampler), ("Metrics", cusrl.utils.metrics.Metrics, C.metrics.Metrics),
    ("NormalDist", cusrl.NormalDist, C.NormalDist), ("Actor", cusrl.Actor, C.Actor), ("Value", cusrl.Value, C.Value), ("Mlp", cusrl.Mlp, C.Mlp),
    ("Rnn", cusrl.Rnn, C.Rnn), ("RunningMeanStd", cusrl.nn.layer.rms.RunningMeanStd, C.nn.rms.RunningMeanStd),
    ("PpoAgentFactory", cusrl.preset.ppo.PpoAgentFactory, C.PpoAgentFactory),
    ("RecurrentPpoAgentFactory", cusrl.preset.ppo.RecurrentPpoAgentFactory, C.RecurrentPpoAgentFactory),
    ("EnvironmentSpec", cusrl.EnvironmentSpec, C.EnvironmentSpec)]
base = set(dir(object)) | set(dir(torch.nn.Module))
print("== attributes / constructor parameters the reference class has and the mirror lacks")
for name, ref, ours in pairs:
    if ref is None:
        print(f"{name}: no reference class of this name")
        continue
    missing = sorted(({n for n in dir(ref) if not n.startswith("__")} - base) - set(dir(ours)))
    ref_params = [p for p in inspect.signature(ref.__init__).parameters if p not in ("self", "args", "kwargs")]
    our_sig = inspect.signature(ours.__init__)
    takes_kwargs = any(p.kind is p.VAR_KEYWORD for p in our_sig.parameters.values())
    params = [] if takes_kwargs else [p for p in ref_params if p not in our_sig.parameters]
    if missing or params:
        print(f"{name}: attributes {missing}; constructor parameters {params}")
print("== constructor defaults that differ")
for name, ref, ours in pairs:
    if ref is None:
        continue
    ours_params = inspect.signature(ours.__init__).parameters
    for pname, rp in inspect.signature(ref.__init__).parameters.items():
        op = ours_params.get(pname)
        if op is None or pname == "self" or rp.default is inspect._empty or op.default is inspect._empty:
            continue
        if repr(rp.default) != repr(op.default):
            print(f"{name}.{pname}: reference {rp.default!r}, here {op.default!r}")
print("== a default PPO agent: instance attributes / state-dict entries the reference has and the mirror lacks")
ref_agent = cusrl.preset.ppo.PpoAgentFactory(device="cpu")(cusrl.EnvironmentSpec(19, 5, num_instances=4))
our_agent = C.PpoAgentFactory(device="cpu")(C.EnvironmentSpec(4, 19, 5))
print("agent attributes:", sorted(set(vars(ref_agent)) - set(vars(our_agent))))
for ref_hook, our_hook in zip(ref_agent.hook, our_agent.hook):
    if lacking := sorted(set(vars(ref_hook)) - set(vars(our_hook))):
        print(f"hook {ref_hook.name}: {lacking}")
print("state dict:", sorted(set(ref_agent.state_dict()) - set(our_agent.state_dict())),
      "| hook entries:", sorted(set(ref_agent.state_dict()["hook"]) - set(our_agent.state_dict()["hook"])))
