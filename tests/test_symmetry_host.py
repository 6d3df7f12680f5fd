"""CPU: host logic of the symmetry hooks (cusrl_b200/hook/symmetry.py) against the reference's own unit test
(cusrl_test/hook/auxiliary/test_symmetry.py:10-36) and the live-reference golden ``symmetry.npz``."""

from __future__ import annotations

import sys
from pathlib import Path
from types import SimpleNamespace

import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import recipes as R  # noqa: E402

from cusrl_b200.hook.symmetry import (MirrorDef, MirrorSymmetryLoss, SymmetricDataAugmentation, TransitionMirroring,
                                      _as_mirror_def)


def test_mirror_def_matches_the_index_and_sign_formula():
    dest, flipped = R.mirror_tables(19, seed=21)
    mirror = MirrorDef(dest, flipped)
    x = torch.randn(7, 3, 19)
    mult = torch.ones(19)
    mult[flipped] = -1.0
    assert torch.equal(mirror(x), x[..., dest] * mult)
    assert torch.equal(mirror(mirror(x)), x)            # the recipe is self-inverse, like the reference helper's
    mask = [i in flipped for i in range(19)]
    assert torch.equal(MirrorDef(dest, mask)(x), mirror(x))   # boolean-mask form (cusrl_test/_helpers.py:28)
    with pytest.raises(IndexError):
        MirrorDef([0, 5], [])


def test_transition_mirroring_rewrites_transition_with_selected_variant():
    # the reference's own test, against this implementation
    def stacked_self_inverse_mirror(tensor):
        return torch.stack([tensor.flip(-1), -tensor], dim=0)

    hook = TransitionMirroring(index=1)
    hook.mirror_observation = hook.mirror_state = hook.mirror_action = stacked_self_inverse_mirror
    transition = {"observation": torch.tensor([[1.0, 2.0, 3.0]]), "state": torch.tensor([[4.0, 5.0, 6.0]])}
    hook.pre_act(transition)
    torch.testing.assert_close(transition["observation"], torch.tensor([[-1.0, -2.0, -3.0]]))
    torch.testing.assert_close(transition["state"], torch.tensor([[-4.0, -5.0, -6.0]]))
    transition["action"] = torch.tensor([[7.0, 8.0, 9.0]])
    hook.post_act(transition)
    torch.testing.assert_close(transition["action"], torch.tensor([[-7.0, -8.0, -9.0]]))
    transition["next_observation"] = torch.tensor([[10.0, 11.0, 12.0]])
    transition["next_state"] = torch.tensor([[13.0, 14.0, 15.0]])
    hook.post_step(transition)
    torch.testing.assert_close(transition["next_observation"], torch.tensor([[-10.0, -11.0, -12.0]]))
    torch.testing.assert_close(transition["next_state"], torch.tensor([[-13.0, -14.0, -15.0]]))
    with pytest.raises(IndexError):
        TransitionMirroring._select_mirrored_tensor(torch.zeros(1, 3), stacked_self_inverse_mirror, 2)
    with pytest.raises(TypeError):
        TransitionMirroring(index=1.0)


def test_augmented_tensors_match_the_reference_golden(golden):
    g = golden("symmetry")
    mirror_obs = MirrorDef(*R.mirror_tables(R.SYMMETRY_SHAPE["obs"], seed=21))
    mirror_act = MirrorDef(*R.mirror_tables(R.SYMMETRY_SHAPE["act"], seed=22))
    obs = g.t("augmentation/buffer/observation")
    act = g.t("augmentation/buffer/action")
    for t in range(obs.shape[0]):
        mirrored, augmented = SymmetricDataAugmentation._build_augmented_tensor(obs[t], mirror_obs)
        assert torch.equal(augmented, g.t("augmentation/buffer/augmented_observation")[t])
        assert torch.equal(mirrored, augmented[:, 1:])
        assert torch.equal(SymmetricDataAugmentation._build_augmented_tensor(act[t], mirror_act)[1],
                           g.t("augmentation/buffer/augmented_action")[t])
    # transition mirroring stores the mirrored observation
    plain = R.anymal_stream(R.SYMMETRY_SHAPE["T"], R.SYMMETRY_SHAPE["N"], seed=24, obs_dim=R.SYMMETRY_SHAPE["obs"],
                            p_term=0.1, p_trunc=0.05)["obs"]
    assert torch.equal(mirror_obs(plain[:-1]), g.t("transition_mirroring/buffer/observation"))


def test_custom_callable_mirrors_and_shape_errors():
    stack2 = lambda x: torch.stack([x, x.flip(-1)])  # noqa: E731  (two variants, as in the reference test :57-59)
    x = torch.randn(4, 6)
    mirrored, augmented = SymmetricDataAugmentation._build_augmented_tensor(x, stack2)
    assert mirrored.shape == (4, 2, 6) and augmented.shape == (4, 3, 6)
    assert torch.equal(augmented[:, 0], x) and torch.equal(augmented[:, 2], x.flip(-1))
    flat2 = lambda x: torch.cat([x, -x])  # noqa: E731  ([V * N, C] form)
    assert SymmetricDataAugmentation._build_mirrored(x, flat2).shape == (2, 4, 6)
    with pytest.raises(ValueError, match="incompatible shape"):
        SymmetricDataAugmentation._build_mirrored(x, lambda t: t[:, :3])


def test_augmentation_objective_doubles_the_per_sample_leaves():
    hook = SymmetricDataAugmentation()
    hook.agent = SimpleNamespace(has_state=False)
    B = 5
    batch = {"observation": torch.randn(B, 4), "augmented_observation": torch.randn(B, 2, 4), "action": torch.randn(B, 3),
             "augmented_action": torch.randn(B, 2, 3), "augmented_next_observation": torch.randn(B, 2, 4),
             "action_logp": torch.randn(B, 1), "advantage": torch.randn(B, 1), "value": torch.randn(B, 1),
             "return": torch.randn(B, 1)}
    adv = batch["advantage"].clone()
    assert hook.objective({"temporal": False}, batch) is None
    assert batch["observation"].shape == (B, 2, 4) and batch["action"].shape == (B, 2, 3)
    for key in ("action_logp", "advantage", "value", "return"):
        assert batch[key].shape == (B, 2, 1)
    assert torch.equal(batch["advantage"][:, 0], adv) and torch.equal(batch["advantage"][:, 1], adv)
    # temporal batches carry the variants on dim 2
    tb = {"augmented_observation": torch.randn(3, B, 2, 4), "augmented_action": torch.randn(3, B, 2, 3),
          "action_logp": torch.randn(3, B, 1), "advantage": torch.randn(3, B, 1), "value": torch.randn(3, B, 1),
          "return": torch.randn(3, B, 1)}
    hook.objective({"temporal": True}, tb)
    assert tb["advantage"].shape == (3, B, 2, 1)


def test_reference_style_mirror_objects_are_recognised_by_duck_typing():
    class Foreign:   # the shape of the reference's MirrorDef: the two tensors
        def __init__(self):
            self.destination = torch.tensor([1, 0, 2])
            self.multiplier = torch.tensor([1.0, 1.0, -1.0])

        def __call__(self, x):
            return x[..., self.destination] * self.multiplier

    foreign = Foreign()
    twin = _as_mirror_def(foreign)
    x = torch.randn(5, 3)
    assert twin is not None and torch.equal(twin(x), foreign(x))
    assert _as_mirror_def(foreign) is twin
    assert _as_mirror_def(lambda t: t) is None


def test_symmetry_hooks_validate_their_arguments_and_spec():
    with pytest.raises(ValueError):
        MirrorSymmetryLoss(-1.0)
    hook = MirrorSymmetryLoss(0.1)
    assert hook.name == "mirror_symmetry_loss" and "weight" in hook._mutable
    hook.agent = SimpleNamespace(environment_spec=SimpleNamespace(mirror_observation=None, mirror_action=None, mirror_state=None),
                                 has_state=False, sampler=None)
    with pytest.raises(ValueError, match="mirror_observation"):
        hook.init()
    assert SymmetricDataAugmentation().training_only and not TransitionMirroring().training_only


def test_mirror_adjoint_tables_are_the_transpose_of_the_transform():
    """The differentiable CUDA path applies `inverse_tables` in its backward: it must be the exact adjoint of the forward
    index-permute + sign-flip (checked here against autograd through torch indexing), and absent for non-permutations."""
    mirror = MirrorDef(*R.mirror_tables(12, seed=22))
    x = torch.randn(6, 12, requires_grad=True)
    g = torch.randn(6, 12)
    (mirror(x) * g).sum().backward()
    inv_dest, inv_mult = mirror.inverse_tables(torch.device("cpu"))
    assert torch.equal(x.grad, g[..., inv_dest[0].long()] * inv_mult[0])
    assert MirrorDef([0, 0, 1], []).inverse_tables(torch.device("cpu")) is None


def test_expanded_done_flags_and_std_vector_helpers():
    from cusrl_b200.hook.on_policy import _std_vector
    from cusrl_b200.nn.recurrent import _expand_done

    done = torch.rand(5, 4, 1) < 0.5
    assert torch.equal(_expand_done(done, 5, 4, (4,)), done.reshape(5, 4))
    wide = _expand_done(done, 5, 12, (4, 3))          # [T, N, 1] flags for a [T, N, 3 variants, C] input
    assert wide.shape == (5, 12) and torch.equal(wide.reshape(5, 4, 3)[:, :, 2], done.reshape(5, 4))
    param = torch.nn.Parameter(torch.rand(7) + 0.5)
    assert _std_vector(param.expand(9, 7), param) is param               # the plain NormalDist: the parameter itself
    assert _std_vector(param.expand(9, 2, 7), param) is param
    mixed = (param.expand(9, 7) + param.flip(0).expand(9, 7)) / 2         # a wrapper that post-processes the std
    vec = _std_vector(mixed, param)
    assert vec.shape == (7,) and torch.equal(vec, mixed[0])
    vec.sum().backward()
    assert torch.allclose(param.grad, torch.ones(7))                     # gradient reaches the parameter through row 0
