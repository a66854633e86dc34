"""train.py on the GPU with the real drop-in model (small configuration): the loop runs the CUDA filter blocks forward and
backward under the reference's three-term loss, checkpoints in the reference's layout and resumes where it stopped."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("isolated_rng")]


def _conf(root, total):
    return {"name": "gpu_unit", "manual_seed": 2204, "path": {"root_dir": str(root)},
            "datasets": {"train": {"type": "SyntheticNoisyPatches", "dataset_args": {"patch_size": 64, "lambda_noise": 25.0, "max_num_patchs": 64},
                                   "dataloader_args": {"batch_size": 2}}},
            "model": {"type": "AbtractMultiScaleGraphFilter",
                      "args": dict(dims=[12, 24, 24, 48], hidden_dims=[24, 48, 48, 96], ngraphs=[2, 4, 2, 4], num_blocks=[1, 1, 1, 1], num_blocks_out=1)},
            "train": {"total_iters": total, "checkpoint_every": 2, "log_every": 0}}


def test_train_checkpoint_resume(tmp_path):
    from imagerestoration_development_unrolling_b200 import ops, train as T
    n0 = ops.launch_count
    losses = {}
    straight = T.train(_conf(tmp_path / "a", 4), on_step=lambda i, l: losses.setdefault(i, l))
    assert ops.launch_count > n0                                   # the filter blocks ran on libglrgtv
    assert sorted(losses) == [0, 1, 2, 3] and all(l == l and l < 10 for l in losses.values())
    T.train(_conf(tmp_path / "b", 2))
    folder = T.checkpoints_folder(_conf(tmp_path / "b", 2))
    state = torch.load(T.latest_checkpoint(folder), weights_only=False)
    assert state["i"] == 1 and len(state["model"]) == len(straight.state_dict())
    seen = []
    resumed = T.train(_conf(tmp_path / "b", 4), on_step=lambda i, l: seen.append(i))
    assert seen == [2, 3]
    # parameter-gradient sums use atomics: equal up to summation order, amplified by four Adam steps
    for (k, a), b in zip(straight.state_dict().items(), resumed.state_dict().values()):
        assert float((a - b).abs().max()) <= 1e-3 * max(float(a.abs().max()), 1e-3), k
