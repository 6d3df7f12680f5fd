"""CPU, world_size 2, Gloo: the cross-rank host logic of the hot path (collectives C1-C5 of SURVEY.md section 2.1)."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent


def test_two_rank_gloo_collectives(tmp_path):
    out = tmp_path / "result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(HERE / "_dist_worker.py"), str(out)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert proc.returncode == 0, proc.stderr[-3000:]
    res = json.loads(out.read_text())
    assert res["mean"] == pytest.approx([1.5, 15.0])              # reduce_mean_ (lr_schedule.py:61-62)
    assert res["mean_var_ok"]                                      # reduce_mean_var_ == oracle merge
    assert res["params_equal_after_broadcast"]                     # broadcast_parameters (actor_critic.py:224)
    assert res["grad_mean"] == pytest.approx(1.5) and res["grad_uniform"]  # reduce_gradients: mean, not sum
    assert res["avg_dict"]["a"] == pytest.approx(0.5) and res["avg_dict"]["only0"] == pytest.approx(5.0)
    # observation normalisation's running statistics: per-step and deferred cross-rank synchronisation give the pooled statistics
    assert res["rms_every_step"] and res["rms_deferred"] and res["rms_deferred_is_local_before_sync"]
