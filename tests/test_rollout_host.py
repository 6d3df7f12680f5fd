"""CPU: host logic of the fused rollout step (template/rollout.py) with every kernel launch stubbed out -- which buffer slot,
which source rows and which tensor-core operand copy each launch of a step is handed.  The arithmetic itself is covered on a
B200 by tests/test_rollout_gpu.py (bit-identical buffers against the generic act / step flow, the reference's control flow
actor_critic.py:227-291); this test pins the ADDRESSING, which is resolved once per buffer layout and reused every step."""

from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("state_dim", [0, 48])
def test_fused_rollout_hands_every_launch_the_slots_of_the_current_step(state_dim):
    # a child process: the tool replaces the library binding of its process with stubs
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "trace_rollout_calls.py"), "--check", "--state-dim", str(state_dim)],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "addressing rules hold" in out.stdout


def test_buffer_layout_version_tracks_allocations():
    """Holders of resolved slot addresses re-resolve when (and only when) a leaf is allocated or dropped."""
    import torch

    from cusrl_b200.template.buffer import Buffer

    buffer = Buffer(4, 3, device="cpu")
    v0 = buffer.layout_version
    buffer.push({"observation": torch.zeros(3, 19), "reward": torch.zeros(3, 1)})
    v1 = buffer.layout_version
    assert v1 > v0
    buffer.push({"observation": torch.ones(3, 19), "reward": torch.ones(3, 1)})
    assert buffer.layout_version == v1                      # writes into existing leaves do not move anything
    buffer["advantage"] = torch.zeros(4, 3, 1)
    v2 = buffer.layout_version
    assert v2 > v1
    buffer["advantage"] = torch.ones(4, 3, 1)
    assert buffer.layout_version == v2
    del buffer["advantage"]
    assert buffer.layout_version > v2
    v3 = buffer.layout_version
    buffer.clear()
    assert buffer.layout_version > v3
